// orb.cu -- the dmORB detector on the device: cv::ORB::detectAndCompute as the reference calls it
// (stage2_detect.cpp:458-496: nfeatures = 1.5 x orb_nfeats with NMS, scale 1.2, orb_nlevels levels, edge 31, HARRIS_SCORE,
// patch 31, fastThreshold = m_current_fast_th), followed by the reference's own m_non_max_sup (S2:296-370) and
// m_update_indexes (S2:65-130) on the float keypoints.
//
//   orb_resize_kernel    level l from level l-1, cv::resize(INTER_LINEAR_EXACT) (8.8 fixed point, tables from the host)
//   fast_blur_kernel     (detect.cu) per level: FAST + 3x3 NMS + 31-px border filter as tile lists, 7x7 sigma-2 blur plane
//   orb_select_kernel    one CTA per image, level after level: row-major corner list -> retainBest(2 N_l) by FAST response
//                        -> Harris -> retainBest(N_l) -> IC angle -> pt * scale          (computeKeyPoints)
//   orb_nms_sort_kernel  one CTA per image: std::sort by response + occupancy grid + std::sort by row (emulated libstdc++)
//   orb_describe_kernel  one warp per surviving keypoint: steered BRIEF on the blurred level   (computeOrbDescriptors)
//
// retainBest is std::nth_element + std::partition on the keypoint vector; the order they leave decides which equal
// responses survive and the order of the output, so both are emulated (orb_core.cuh), the large partitions of
// nth_element with the block- / warp-cooperative partition of introsort_emul.cuh.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "orb.cuh"
#include "corner_lists.cuh"

namespace svo {

namespace {

constexpr int OT = 256;                   // threads of the per-image kernels

// ---- level l from level l-1 ------------------------------------------------------------------------------------------
// src: by batch item (level 0 = the caller's images, levels >= 1 = the level planes); 4 output pixels per thread
__global__ void __launch_bounds__(128)
orb_resize_kernel(const uint8_t* __restrict__ srcL, const uint8_t* __restrict__ srcR, size_t sstride, size_t simage_pitch, OrbLevelDev D) {
    const int item = blockIdx.z >> 1, side = blockIdx.z & 1;
    const int y = blockIdx.y, x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (x4 >= D.pitch) return;
    const uint8_t* src = (side ? srcR : srcL) + (size_t)item * simage_pitch;
    uint8_t* dst = (side ? D.imgR : D.imgL) + (size_t)item * D.pitch * D.h + (size_t)y * D.pitch;
    const int yo = D.yofs[y], yc = D.yc1[y];
    const uint8_t* r0 = src + (size_t)yo * sstride;
    const uint8_t* r1 = r0 + sstride;
    uint32_t w = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int x = x4 + q;
        uint32_t v = 0;
        if (x < D.w) {
            const int xo = D.xofs[x], xc = D.xc1[x];
            const uint32_t h0 = r0[xo] * (256 - xc) + r0[xo + 1] * xc;          // ufixedpoint16 (8.8)
            const uint32_t h1 = r1[xo] * (256 - xc) + r1[xo + 1] * xc;
            v = (h0 * (256 - yc) + h1 * yc + (1u << 15)) >> 16;                 // ufixedpoint32 -> uint8, round half up
        }
        w |= v << (8 * q);
    }
    *reinterpret_cast<uint32_t*>(dst + x4) = w;
}

// ---- std::nth_element with the block ------------------------------------------------------------------------------------
// a[0..n): indices; all threads call.  tmp: n words, s_blk: 64 * (OT / 32) ints (shared), s_cut: one int (shared)
template <class C>
__device__ void nth_element_block(const C& c, uint32_t* a, int n, int nth, uint32_t* tmp, int* s_blk, int* s_cut) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (n == 0 || nth == n) return;
    int first = 0, last = n;
    int depth = 2 * floor_log2(n);
    uint16_t* pos16 = reinterpret_cast<uint16_t*>(tmp);
    while (last - first > 64 && depth > 0 && n < 65536) {
        --depth;
        if (tid == 0) median_to_first_(c, a, first, first + 1, first + (last - first) / 2, last - 1);
        __syncthreads();
        int cut;
        if (last - first >= kBlockPartMin) {
            cut = block_unguarded_partition(c, a, first, last, pos16, pos16 + n, s_blk);
        } else {
            if (warp == 0) {
                const int cw = warp_unguarded_partition(c, a, first, last, s_blk);
                if (lane == 0) *s_cut = cw;
            }
            __syncthreads();
            cut = *s_cut;
            __syncthreads();
        }
        if (cut <= nth) first = cut; else last = cut;
    }
    __syncthreads();
    if (tid == 0) introselect_seq(c, a, first, last, nth, depth);
    __syncthreads();
}

// std::partition(idx + first, idx + n, resp >= v) as retainBest uses it: the vector is resized to the partition point right
// after, so only the kept prefix [first, first + T) matters.  The bidirectional std::partition leaves the trues of that
// prefix in place and fills its k-th false position (ascending) with the k-th true of the rest taken from the END
// (descending) -- computed here by all threads with two ordered compactions.  tmp: n words.  Returns first + T.
__device__ int partition_tail_block(uint32_t* idx, int first, int n, const float* resp, float v, uint32_t* tmp, int* s_warp) {
    const int tid = threadIdx.x;
    int c = 0;
    for (int i = first + tid; i < n; i += OT) c += resp[idx[i]] >= v ? 1 : 0;
    int T;
    block_exclusive_scan<OT>(c, s_warp, T);
    const int mid = first + T;
    int K, K2;
    {   // false positions of [first, mid), ascending
        const int per = (mid - first + OT - 1) / OT;
        const int lo = min(mid, first + tid * per), hi = min(mid, lo + per);
        int cf = 0;
        for (int i = lo; i < hi; ++i) cf += resp[idx[i]] >= v ? 0 : 1;
        int off = block_exclusive_scan<OT>(cf, s_warp, K);
        for (int i = lo; i < hi; ++i) if (!(resp[idx[i]] >= v)) tmp[off++] = (uint32_t)i;
    }
    {   // the trues of [mid, n), ascending by position
        const int per = (n - mid + OT - 1) / OT;
        const int lo = min(n, mid + tid * per), hi = min(n, lo + per);
        int ct = 0;
        for (int i = lo; i < hi; ++i) ct += resp[idx[i]] >= v ? 1 : 0;
        int off = K + block_exclusive_scan<OT>(ct, s_warp, K2);
        for (int i = lo; i < hi; ++i) if (resp[idx[i]] >= v) tmp[off++] = idx[i];
    }
    __syncthreads();
    for (int k = tid; k < K; k += OT) idx[tmp[k]] = tmp[K + (K - 1 - k)];
    __syncthreads();
    return mid;
}

// KeyPointsFilter::retainBest on idx[0..n) (block-wide); returns the new size on every thread
__device__ int retain_best_block(uint32_t* idx, int n, const float* resp, int n_points, uint32_t* tmp, int* s_blk, int* s_cut, int* s_warp) {
    if (!(n_points >= 0 && n > n_points)) return n;
    if (n_points == 0) return 0;
    nth_element_block(OrbRespGreater{resp}, idx, n, n_points - 1, tmp, s_blk, s_cut);
    const float ambiguous = resp[idx[n_points - 1]];
    return partition_tail_block(idx, n_points, n, resp, ambiguous, tmp, s_warp);
}

// HarrisResponses / ICAngles of one keypoint by a warp: the sums are integers, so the lanes may add in any order
__device__ __forceinline__ float harris_warp(const uint8_t* __restrict__ img, int step, int x0, int y0, int lane) {
    int a = 0, b = 0, c = 0;
    for (int k = lane; k < kOrbHarrisBlock * kOrbHarrisBlock; k += 32) {
        const int i = k / kOrbHarrisBlock, j = k - i * kOrbHarrisBlock;
        const uint8_t* p = img + (y0 - kOrbHarrisBlock / 2 + i) * step + (x0 - kOrbHarrisBlock / 2 + j);
        const int Ix = (p[1] - p[-1]) * 2 + (p[-step + 1] - p[-step - 1]) + (p[step + 1] - p[step - 1]);
        const int Iy = (p[step] - p[-step]) * 2 + (p[step - 1] - p[-step - 1]) + (p[step + 1] - p[-step + 1]);
        a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
    }
    a = __reduce_add_sync(0xffffffffu, a); b = __reduce_add_sync(0xffffffffu, b); c = __reduce_add_sync(0xffffffffu, c);
    return orb_harris_from_sums(a, b, c);
}

__device__ __forceinline__ float ic_angle_warp(const uint8_t* __restrict__ img, int step, int x0, int y0, int lane) {
    const int umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
    const uint8_t* center = img + y0 * step + x0;
    const int v = lane & 15, half = lane >> 4;           // lane = row v of the half-disc, columns u < 0 (half 0) or u >= 0 (half 1)
    const int d = umax[v];
    const int u_lo = half ? 0 : -d, u_hi = half ? d : -1;
    int m_01 = 0, m_10 = 0;
    if (v == 0) {
        for (int u = u_lo; u <= u_hi; ++u) m_10 += u * center[u];
    } else {
        int v_sum = 0;
        for (int u = u_lo; u <= u_hi; ++u) {
            const int val_plus = center[u + v * step], val_minus = center[u - v * step];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 = v * v_sum;
    }
    m_01 = __reduce_add_sync(0xffffffffu, m_01); m_10 = __reduce_add_sync(0xffffffffu, m_10);
    return orb_fast_atan2((float)m_01, (float)m_10);
}

// ---- computeKeyPoints -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(OT)
orb_select_kernel(Pools P, OrbDev D, const uint8_t* __restrict__ img0L, const uint8_t* __restrict__ img0R, size_t stride0,
                  size_t image_pitch0, int item0, int smem_cap) {
    extern __shared__ __align__(16) uint32_t osm[];
    uint32_t* s_idx = osm;                                           // [smem_cap]
    float* s_resp = reinterpret_cast<float*>(osm + smem_cap);       // [smem_cap]
    __shared__ int s_warp[OT / 32];
    __shared__ int s_blk[64 * (OT / 32)];
    __shared__ int s_cut;
    const int item = blockIdx.x >> 1, side = blockIdx.x & 1;
    const int slot = P.slots[item];
    const int tid = threadIdx.x;
    const size_t im = img_idx(slot, side);
    const int l = blockIdx.y;                                        // one CTA per (image, level)
    const size_t il = im * D.n_levels + l;
    uint32_t* g_idx = D.idx + il * D.cap;
    float* g_resp = D.resp + il * D.cap;
    uint32_t* pos = D.pos + il * D.cap;
    uint32_t* tmp = D.tmp + il * D.cap;
    float4* raw = D.lraw + il * D.lvl_cap;
    int n_out = 0;
    bool overflow = false;

    {
        const OrbLevelDev& Lv = D.L[l];
        const int W = Lv.w, H = Lv.h;
        const uint8_t* img;
        int step;
        if (l == 0) { img = (side ? img0R : img0L) + (size_t)(item0 + item) * image_pitch0; step = (int)stride0; }
        else { img = (side ? Lv.imgR : Lv.imgL) + (size_t)(item0 + item) * Lv.pitch * H; step = Lv.pitch; }
        // A. the level's FAST corners in row-major order (cv::FAST's keypoint order) from fast_blur_kernel's tile lists
        const int tiles_x = tiles_x_of(W), tiles_y = tiles_y_of(H);
        const uint16_t* __restrict__ seg = Lv.tile_seg + im * (size_t)tiles_x * tiles_y * kTileH;
        const uint16_t* __restrict__ tlst = Lv.tile_list + im * (size_t)tiles_x * tiles_y * kTileCap;
        const CornerRun run = corner_run(seg, H * tiles_x, OT, tid);
        int total;
        const int o = block_exclusive_scan<OT>(run.cnt, s_warp, total);
        if (total > D.cap) overflow = true;
        const int n = min(total, D.cap);
        uint32_t* idx = n <= smem_cap ? s_idx : g_idx;
        float* resp = n <= smem_cap ? s_resp : g_resp;
        corner_codes(seg, run, tiles_x, o, n, idx);
        __syncthreads();
        for (int i = tid; i < n; i += OT) {
            const Corner c = corner_fetch(tlst, idx[i], tiles_x);
            idx[i] = (uint32_t)i;
            resp[i] = (float)c.score;                                    // cv::FAST: response = score
            pos[i] = ((uint32_t)c.y << 16) | (uint32_t)c.x;
        }
        __syncthreads();
        // B. retainBest(2 * featuresNum) by FAST response (HARRIS_SCORE keeps twice as many for the second cull)
        const int n2 = retain_best_block(idx, n, resp, 2 * Lv.nfeat, tmp, s_blk, &s_cut, s_warp);
        // C. HarrisResponses of the survivors
        for (int k = tid >> 5; k < n2; k += OT / 32) {
            const uint32_t i = idx[k], p = pos[i];
            const float h = harris_warp(img, step, (int)(p & 0xFFFFu), (int)(p >> 16), tid & 31);
            if ((tid & 31) == 0) resp[i] = h;
        }
        __syncthreads();
        // D. retainBest(featuresNum) by Harris response
        const int n3 = retain_best_block(idx, n2, resp, Lv.nfeat, tmp, s_blk, &s_cut, s_warp);
        // E. ICAngles, pt *= scale
        for (int k = tid >> 5; k < min(n3, D.lvl_cap); k += OT / 32) {
            const uint32_t i = idx[k], p = pos[i];
            const int x = (int)(p & 0xFFFFu), y = (int)(p >> 16);
            const float ang = ic_angle_warp(img, step, x, y, tid & 31);
            if ((tid & 31) == 0) raw[k] = make_float4(__fmul_rn((float)x, Lv.scale), __fmul_rn((float)y, Lv.scale), resp[i], ang);
        }
        if (n3 > D.lvl_cap) overflow = true;
        n_out = min(n3, D.lvl_cap);
    }
    if (tid == 0) {
        D.n_lraw[il] = n_out;
        if (overflow) P.state[slot].status = -3;       // SVO_ERR_CAPACITY
    }
}

// ---- m_non_max_sup + m_update_indexes on detectAndCompute's keypoints ---------------------------------------------------
struct IndexedAscComp {                 // KpRowSorter on float rows: comp(a, b) = y[a] < y[b]  (libstereo-odometry.h:120-127)
    const float* key;
    __device__ __forceinline__ bool lt(uint32_t a, uint32_t b) const { return key[a] < key[b]; }
};

__global__ void __launch_bounds__(OT)
orb_nms_sort_kernel(Pools P, OrbDev D, uint32_t* __restrict__ occ_scratch, int occ_words_per_img, int occ_smem_words,
                    uint32_t* __restrict__ sort_scratch, uint2* __restrict__ list_scratch) {
    extern __shared__ __align__(16) uint32_t osm[];
    const int rc = D.raw_cap;
    uint32_t* s_a = osm;                                             // [raw_cap] indices sorted by response
    uint32_t* s_b = osm + rc;                                        // [raw_cap] positions sorted by row
    float* s_key = reinterpret_cast<float*>(osm + 2 * rc);          // [raw_cap] response, then pt.y of the selection
    __shared__ int s_cnt[3];
    __shared__ int s_pos[64 * (OT / 32)];
    __shared__ int s_nsel;
    const int item = blockIdx.x >> 1, side = blockIdx.x & 1;
    const int slot = P.slots[item];
    const int tid = threadIdx.x;
    const size_t im = img_idx(slot, side);
    // allKeypoints: the levels' keypoints one after the other (computeKeyPoints appends level by level)
    float4* __restrict__ raw = D.raw + im * D.raw_cap;
    int32_t* __restrict__ raw_oct = D.raw_oct + im * D.raw_cap;
    int n = 0;
    bool overflow = false;
    for (int l = 0; l < D.n_levels; ++l) {
        const int nl = D.n_lraw[im * D.n_levels + l];
        const float4* __restrict__ src = D.lraw + (im * D.n_levels + l) * D.lvl_cap;
        for (int k = tid; k < nl; k += OT)
            if (n + k < D.raw_cap) { raw[n + k] = src[k]; raw_oct[n + k] = l; }
        if (n + nl > D.raw_cap) overflow = true;
        n = min(n + nl, D.raw_cap);
    }
    if (tid == 0) {
        D.n_raw[im] = n;
        if (overflow) P.state[slot].status = -3;       // SVO_ERR_CAPACITY
    }
    __syncthreads();
    const int list_cap = P.p.max_corners / 16 + 2;
    uint2* listA = list_scratch + (size_t)blockIdx.x * 3 * list_cap;
    uint2* listB = listA + list_cap;
    uint2* listS = listB + list_cap;
    uint32_t* tmp = sort_scratch + (size_t)blockIdx.x * P.p.max_corners;
    const int W = P.p.W, H = P.p.H;

    for (int i = tid; i < n; i += OT) { s_a[i] = (uint32_t)i; s_key[i] = raw[i].z; }
    __syncthreads();
    int n_sel;
    if (P.p.nms) {
        introsort_block(IndexedDescComp{s_key}, s_a, n, listA, listB, listS, s_cnt, tmp, s_pos, n);      // S2:322-324
        const unsigned cell = (unsigned)(P.p.min_distance / 2.0);     // S2:331
        const float cell_inv = 1.0f / (float)cell;                     // S2:332
        const unsigned glx = (unsigned)(1 + W * cell_inv), gly = (unsigned)(1 + H * cell_inv);
        const unsigned occ_words = (glx * gly + 31) / 32;
        // the occupancy grid lives in shared memory when it fits (the walk below is one thread's chain of dependent accesses)
        uint32_t* occ = (int)occ_words <= occ_smem_words ? osm + 3 * rc : occ_scratch + (size_t)blockIdx.x * occ_words_per_img;
        for (unsigned i = tid; i < occ_words; i += OT) occ[i] = 0;
        for (int i = tid; i < n; i += OT) {            // the grid cell of every keypoint (s_b is free until the row sort)
            const float4 kp = raw[i];
            s_b[i] = ((unsigned)__fmul_rn(kp.x, cell_inv) << 16) | (unsigned)__fmul_rn(kp.y, cell_inv);
        }
        __syncthreads();
        if (tid == 0) {
            auto get = [&](unsigned x, unsigned y) { const unsigned b = x * gly + y; return (occ[b >> 5] >> (b & 31)) & 1u; };
            auto set = [&](unsigned x, unsigned y) { const unsigned b = x * gly + y; occ[b >> 5] |= 1u << (b & 31); };
            int c = 0, k = 0;
            while (c < P.p.quota && c < P.p.kp_cap && k < n) {
                const uint32_t id = s_a[k++];
                const unsigned sx = s_b[id] >> 16, sy = s_b[id] & 0xFFFFu;
                if (get(sx, sy)) continue;
                set(sx, sy);
                if (sx > 0) set(sx - 1, sy);
                if (sy > 0) set(sx, sy - 1);
                if (sx < glx - 1) set(sx + 1, sy);
                if (sy < gly - 1) set(sx, sy + 1);
                s_a[c++] = id;
            }
            s_nsel = c;
        }
        __syncthreads();
        n_sel = s_nsel;
    } else {
        n_sel = min(n, P.p.kp_cap);                   // feats_vector.swap(...)  S2:613
    }
    __syncthreads();
    // m_update_indexes: std::sort of the positions by pt.y
    for (int k = tid; k < n_sel; k += OT) { s_b[k] = (uint32_t)k; s_key[k] = 0.f; }
    __syncthreads();
    for (int k = tid; k < n_sel; k += OT) s_key[k] = raw[s_a[k]].y;
    __syncthreads();
    introsort_block(IndexedAscComp{s_key}, s_b, n_sel, listA, listB, listS, s_cnt, tmp, s_pos, n_sel);   // S2:87-89
    const int cur = P.state[slot].cur;
    const size_t f = frame_idx(slot, cur, side);
    float4* __restrict__ kpf = P.kpf + f * (size_t)P.p.kp_cap;
    int32_t* __restrict__ kpo = P.kpo + f * (size_t)P.p.kp_cap;
    int2* __restrict__ kp = P.kp + f * (size_t)P.p.kp_cap;
    for (int i = tid; i < n_sel; i += OT) {
        const uint32_t id = s_a[s_b[i]];
        const float4 v = raw[id];
        kpf[i] = v;
        kpo[i] = raw_oct[id];
        kp[i] = make_int2((int)(((unsigned)(int)v.y << 16) | ((unsigned)(int)v.x & 0xFFFFu)), 0);   // truncated copy for index-only readers
    }
    if (tid == 0) {
        P.n_kp[f] = n_sel;
        P.oct_stats[stat_idx(slot, P.octave, side)] = n_sel;
        if (side == 0) P.result[slot].detected_left = n_sel; else P.result[slot].detected_right = n_sel;
    }
}

// ---- computeOrbDescriptors ------------------------------------------------------------------------------------------------
struct OrbPattern { int8_t v[256][4]; };
constexpr OrbPattern make_orb_pattern() {
    constexpr int pat[256][4] = {
#include "orb_pattern.inc"
    };
    OrbPattern t{};
    for (int i = 0; i < 256; ++i)
        for (int k = 0; k < 4; ++k) t.v[i][k] = (int8_t)pat[i][k];
    return t;
}
__device__ const OrbPattern g_orb_pattern = make_orb_pattern();

constexpr int OD_WARPS = 8;
__global__ void __launch_bounds__(OD_WARPS * 32)
orb_describe_kernel(Pools P, OrbDev D) {
    const int item = blockIdx.y >> 1, side = blockIdx.y & 1;
    const int slot = P.slots[item];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cur = P.state[slot].cur;
    const size_t f = frame_idx(slot, cur, side);
    const int n = P.n_kp[f];
    const int i = blockIdx.x * OD_WARPS + warp;
    if (i >= n) return;
    const float4 kp = P.kpf[f * (size_t)P.p.kp_cap + i];
    const int oct = P.kpo[f * (size_t)P.p.kp_cap + i];
    const OrbLevelDev& Lv = D.L[oct];
    const float inv = __fdiv_rn(1.f, Lv.scale);
    const int cy = __float2int_rn(__fmul_rn(kp.y, inv)), cx = __float2int_rn(__fmul_rn(kp.x, inv));
    float a, b;
    orb_cos_sin(kp.w, a, b);
    const uint8_t* __restrict__ center = Lv.blur + img_idx(slot, side) * (size_t)Lv.pitch * Lv.h + (size_t)cy * Lv.pitch + cx;
    uint32_t byte = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int8_t* p = g_orb_pattern.v[lane * 8 + t];
        int x0, y0, x1, y1;
        orb_steer(p[0], p[1], a, b, x0, y0);
        orb_steer(p[2], p[3], a, b, x1, y1);
        const int t0 = center[y0 * Lv.pitch + x0], t1 = center[y1 * Lv.pitch + x1];
        byte |= (t0 < t1 ? 1u : 0u) << t;
    }
    // 32 descriptor bytes of the keypoint: assemble words with shuffles, lanes 0..7 store
    uint32_t w = byte;
    w |= __shfl_down_sync(0xffffffffu, byte, 1) << 8;
    w |= __shfl_down_sync(0xffffffffu, byte, 2) << 16;
    w |= __shfl_down_sync(0xffffffffu, byte, 3) << 24;
    uint32_t* out = reinterpret_cast<uint32_t*>(P.desc + (f * (size_t)P.p.kp_cap + i) * 2);
    if ((lane & 3) == 0) out[lane >> 2] = w;
}

}  // namespace

// ---- host side ------------------------------------------------------------------------------------------------------------
// words of the occupancy grid of m_non_max_sup that orb_nms_sort_kernel keeps in shared memory (0: the grid stays in global memory)
static int orb_nms_occ_smem_words(const Params& p) {
    const unsigned cell = (unsigned)(p.min_distance / 2.0);
    const float inv = 1.0f / (float)cell;
    const unsigned glx = (unsigned)(1 + p.W * inv), gly = (unsigned)(1 + p.H * inv);
    const unsigned words = (glx * gly + 31) / 32;
    return words * 4 <= 96 * 1024 ? (int)words : 0;
}

static int orb_select_smem_cap(const OrbDev& D) { return D.cap < 8192 ? D.cap : 8192; }   // 64 KB: three CTAs per SM

cudaError_t alloc_orb(OrbDev& D, const Params& p, int n_slots, int n_levels, int n_extract, std::vector<void*>& allocs) {
    memset(&D, 0, sizeof(D));
    if (n_levels < 1 || n_levels > kOrbMaxLevels) return cudaErrorInvalidValue;
    const OrbLevels G = orb_levels(p.W, p.H, n_levels, n_extract);
    D.n_levels = n_levels;
    D.n_extract = n_extract;
    D.cap = p.max_corners;
    D.raw_cap = 2 * n_extract + 64;
    D.lvl_cap = 2 * G.nfeat[0] + 64;            // retainBest(N_l) keeps N_l + the ties of the boundary response
    const size_t S = (size_t)n_slots;
    cudaError_t e = cudaSuccess;
    auto A = [&](void** ptr, size_t bytes) { if (e != cudaSuccess) return; e = cudaMalloc(ptr, bytes ? bytes : 1); if (e == cudaSuccess) { allocs.push_back(*ptr); e = cudaMemset(*ptr, 0, bytes ? bytes : 1); } };
    auto upload = [&](const std::vector<uint16_t>& v, const uint16_t** out) {
        void* d = nullptr;
        A(&d, v.size() * sizeof(uint16_t));
        if (e == cudaSuccess) e = cudaMemcpy(d, v.data(), v.size() * sizeof(uint16_t), cudaMemcpyHostToDevice);
        *out = (const uint16_t*)d;
    };
    for (int l = 0; l < n_levels; ++l) {
        OrbLevelDev& L = D.L[l];
        L.w = G.w[l]; L.h = G.h[l]; L.pitch = (G.w[l] + 15) & ~15; L.nfeat = G.nfeat[l]; L.scale = G.scale[l];
        if (L.w < 64 || L.h < 64) return cudaErrorInvalidValue;          // FAST + the 31-px border need some image
        const size_t plane = (size_t)L.pitch * L.h;
        if (l > 0) {
            A((void**)&L.imgL, S * plane); A((void**)&L.imgR, S * plane);
            // uniform tables: out-of-range outputs replicate the border pixel with weights (256, 0) / (0, 256)
            auto tables = [&](int src, int dst, const uint16_t** dofs, const uint16_t** dc1) {
                const OrbAxis ax = orb_resize_axis(src, dst);
                std::vector<uint16_t> ofs(dst), c1(dst);
                for (int v = 0; v < dst; ++v) {
                    if (v < ax.lo) { ofs[v] = 0; c1[v] = 0; }
                    else if (v >= ax.hi) { ofs[v] = (uint16_t)(src - 2); c1[v] = 256; }
                    else { ofs[v] = (uint16_t)ax.ofs[v]; c1[v] = (uint16_t)ax.c1[v]; }
                }
                upload(ofs, dofs); upload(c1, dc1);
            };
            tables(G.w[l - 1], L.w, &L.xofs, &L.xc1);
            tables(G.h[l - 1], L.h, &L.yofs, &L.yc1);
        }
        A((void**)&L.blur, S * 2 * plane);
        const size_t tiles = (size_t)tiles_x_of(L.w) * tiles_y_of(L.h);
        A((void**)&L.tile_seg, S * 2 * tiles * kTileH * sizeof(uint16_t));
        A((void**)&L.tile_list, S * 2 * tiles * kTileCap * sizeof(uint16_t));
    }
    A((void**)&D.idx, S * 2 * n_levels * D.cap * sizeof(uint32_t));
    A((void**)&D.resp, S * 2 * n_levels * D.cap * sizeof(float));
    A((void**)&D.pos, S * 2 * n_levels * D.cap * sizeof(uint32_t));
    A((void**)&D.tmp, S * 2 * n_levels * D.cap * sizeof(uint32_t));
    A((void**)&D.lraw, S * 2 * n_levels * D.lvl_cap * sizeof(float4));
    A((void**)&D.n_lraw, S * 2 * n_levels * sizeof(int32_t));
    A((void**)&D.raw, S * 2 * D.raw_cap * sizeof(float4));
    A((void**)&D.raw_oct, S * 2 * D.raw_cap * sizeof(int32_t));
    A((void**)&D.n_raw, S * 2 * sizeof(int32_t));
    if (e != cudaSuccess) return e;
    e = SVO_RAISE_DYN_SMEM(orb_select_kernel, orb_select_smem_cap(D) * 8);
    if (e != cudaSuccess) return e;
    return SVO_RAISE_DYN_SMEM(orb_nms_sort_kernel, D.raw_cap * 12 + 4 * orb_nms_occ_smem_words(p));
}

cudaError_t launch_orb_detect(const Pools& P, const OrbDev& D, int n, const uint8_t* d_left, const uint8_t* d_right, size_t stride,
                              size_t image_pitch, int item0, uint32_t* occ_scratch, int occ_words_per_img, uint32_t* sort_scratch,
                              uint2* list_scratch, cudaStream_t st, int* launches) {
    // the pyramid, level after level (each level is resized from the previous one), and FAST + blur of every level
    const uint8_t *sL = d_left + (size_t)item0 * image_pitch, *sR = d_right + (size_t)item0 * image_pitch;
    size_t sstride = stride, spitch = image_pitch;
    int L = 0;
    for (int l = 0; l < D.n_levels; ++l) {
        const OrbLevelDev& Lv = D.L[l];
        if (l > 0) {
            OrbLevelDev Dl = Lv;
            Dl.imgL += (size_t)item0 * Lv.pitch * Lv.h; Dl.imgR += (size_t)item0 * Lv.pitch * Lv.h;
            const dim3 g((Lv.pitch / 4 + 127) / 128, Lv.h, 2 * n);
            orb_resize_kernel<<<g, 128, 0, st>>>(sL, sR, sstride, spitch, Dl);
            sL = Dl.imgL; sR = Dl.imgR; sstride = (size_t)Lv.pitch; spitch = (size_t)Lv.pitch * Lv.h;
            ++L;
        }
        Pools PL = P;
        PL.p.W = Lv.w; PL.p.H = Lv.h;
        PL.blur = Lv.blur; PL.tile_seg = Lv.tile_seg; PL.tile_list = Lv.tile_list;
        cudaError_t e = launch_fast_blur_only(PL, n, sL, sR, sstride, spitch, Lv.pitch, st);
        if (e != cudaSuccess) return e;
        ++L;
    }
    const int cap = orb_select_smem_cap(D);
    orb_select_kernel<<<dim3(2 * n, D.n_levels), OT, cap * 8, st>>>(P, D, d_left, d_right, stride, image_pitch, item0, cap);
    const int occ_smem = orb_nms_occ_smem_words(P.p);
    orb_nms_sort_kernel<<<2 * n, OT, D.raw_cap * 12 + 4 * occ_smem, st>>>(P, D, occ_scratch, occ_words_per_img, occ_smem, sort_scratch, list_scratch);
    const dim3 gd((P.p.kp_cap + OD_WARPS - 1) / OD_WARPS, 2 * n);
    orb_describe_kernel<<<gd, OD_WARPS * 32, 0, st>>>(P, D);
    if (launches) *launches += L + 3;
    return cudaGetLastError();
}

}  // namespace svo
