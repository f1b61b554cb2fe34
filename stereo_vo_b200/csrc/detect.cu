// detect.cu -- stage 2 of the hot path on sm_100a: FAST-9/16 + 7x7 blur + top-K selection +
// row sort + 256-bit BRIEF.  Replaces, for the dmFAST_ORB branch,
//   cv::FastFeatureDetector::detect + cv::ORB::compute   stage2_detect.cpp:510-512
//   m_non_max_sup (8-arg)                                stage2_detect.cpp:296-370 (called :590-597)
//   m_update_indexes                                     stage2_detect.cpp:65-130  (called :618)
//
// Order of work differs from the reference (which describes every corner and then throws most
// descriptors away, see the authors' TODO at stage2_detect.cpp:507): corners are selected first and
// only the survivors are described.  A descriptor depends on its own keypoint only and the border
// filter of ORB::compute is applied before the selection, so the results are identical.
#include "detect.cuh"
#include "introsort_emul.cuh"
#include "corner_lists.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cstdlib>
#include <cstring>

namespace svo {

// ------------------------------------------------------------------------------------------------
// K1  fast_blur_kernel: one CTA per 64x32 tile.
//   in : 8-bit image (HBM, read once through shared memory, halo 4)
//   out: blur  (7x7 sigma=2, OpenCV float32 FMA order, rint, saturate)        1 B / px
//        score (FAST-9/16 corner score after strict 3x3 NMS and the 31-px ORB border filter) 1 B / px
// Roofline: algorithmic HBM traffic 3 B / px; the kernel is instruction-issue bound (ring pre-test, arc
// min/max network, float blur), not HBM bound.  Stages, all on shared memory:
//   1  stage the tile (+ halo 4) as 32-bit words
//   2  blur row pass (float32)
//   3a FAST pre-test, 4 pixels per work item on 16-bit SIMD lanes (DPX VIMNMX.S16x2), candidates -> list
//   3b full score for the listed candidates (dense), corners -> second list
//   3c strict 3x3 NMS + ORB border filter over the corner list (dense)
//   4  blur column pass + stores, 4 pixels per work item
// ------------------------------------------------------------------------------------------------
constexpr int TW = kTileW, TH = kTileH, HALO = 4;
constexpr int SIN_W = TW + 2 * HALO;   // 72 bytes per staged row ...
constexpr int SIN_WW = SIN_W / 4;      // ... = 18 words
// TMA wants the box to start at a multiple of 16 bytes in x and to be a multiple of 16 bytes wide: the staged box begins
// SIN_XOFF = 12 columns left of the halo (x0 - 16) and is 96 bytes wide; logical column c lives at byte SIN_XOFF + c.
constexpr int SIN_P = 96;              // row pitch of the staged tile and of the score plane in bytes
constexpr int SIN_PW = SIN_P / 4;      // 24 words
constexpr int SIN_XOFF = 12;           // bytes; a multiple of 4, so staged words stay aligned
constexpr int SIN_H = TH + 2 * HALO;   // 40
constexpr int SC_H = TH + 2;           // score rows: tile rows -1 .. TH (row r <-> tile row r-1); columns as in s_in
constexpr int K1_THREADS = 256;
// bit 7 of every non-zero byte
__device__ __forceinline__ uint32_t nonzero_bytes(uint32_t w) { return ((w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu | w) & 0x80808080u; }

constexpr int K1_NCAND = SC_H * (TW + 2);
constexpr int K1_ITEMS3 = SC_H * SIN_WW;   // pre-test work items (row, word)

// one-step REFLECT_101 (|offset| <= 4 < n)
__device__ __forceinline__ int reflect1(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

// cv::getGaussianKernel(7, 2, CV_32F) bit patterns
__device__ __forceinline__ float gk(int i) {
    const uint32_t b[7] = {0x3D8FAFB1u, 0x3E06387Eu, 0x3E434A39u, 0x3E5D4AE0u, 0x3E434A39u, 0x3E06387Eu, 0x3D8FAFB1u};
    return __uint_as_float(b[i]);
}

// FAST-9/16 score of the pixel at s (shared memory, row pitch SIN_W); 0 if it is not a corner.
//   score = max over the 16 arcs of length 9 of min(d) and of min(-d), minus 1, d_k = I(p) - I(ring_k);
//   the pixel is a corner iff score >= th (cv::FAST_t<16> + cornerScore<16>).
// Both polarities ride in one register as 16-bit lanes {256 + d, 256 - d} (always positive, so one IMAD
// per ring pixel builds the pair: r * -K + (v * K + C) with K = 1 - 65536); the sliding 9-window minima
// then need 16-bit SIMD min/max only (DPX: VIMNMX.S16x2 / VIMNMX3.S16x2).
// NOTE: the obvious scalar form `best = max(best, max(lo9, -hi9))` is miscompiled by ptxas 12.9
// (-O1 and above, also by the driver JIT): wrong scores on sm_100a, correct with -Xptxas -O0.
__device__ __forceinline__ int fast9_score(const uint8_t* __restrict__ s, int th) {
    constexpr unsigned K = 0xFFFF0001u;                      // 1 - 65536
    const unsigned cv = (unsigned)s[0] * K + 0x01000100u;    // {256 + v, 256 - v}
    unsigned p[16];
    p[0] = cv - (unsigned)s[3 * SIN_P + 0] * K;   p[1] = cv - (unsigned)s[3 * SIN_P + 1] * K;
    p[2] = cv - (unsigned)s[2 * SIN_P + 2] * K;   p[3] = cv - (unsigned)s[1 * SIN_P + 3] * K;
    p[4] = cv - (unsigned)s[3] * K;               p[5] = cv - (unsigned)s[-1 * SIN_P + 3] * K;
    p[6] = cv - (unsigned)s[-2 * SIN_P + 2] * K;  p[7] = cv - (unsigned)s[-3 * SIN_P + 1] * K;
    p[8] = cv - (unsigned)s[-3 * SIN_P + 0] * K;  p[9] = cv - (unsigned)s[-3 * SIN_P - 1] * K;
    p[10] = cv - (unsigned)s[-2 * SIN_P - 2] * K; p[11] = cv - (unsigned)s[-1 * SIN_P - 3] * K;
    p[12] = cv - (unsigned)s[-3] * K;             p[13] = cv - (unsigned)s[1 * SIN_P - 3] * K;
    p[14] = cv - (unsigned)s[2 * SIN_P - 2] * K;  p[15] = cv - (unsigned)s[3 * SIN_P - 1] * K;
    // minima over the 16 arcs of 9: lo3[k] = min(p[k..k+2]), lo9[k] = min(lo3[k], lo3[k+3], lo3[k+6]); 3-input DPX ops
    unsigned lo3[16], lo9[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) lo3[k] = __vimin3_s16x2(p[k], p[(k + 1) & 15], p[(k + 2) & 15]);
#pragma unroll
    for (int k = 0; k < 16; ++k) lo9[k] = __vimin3_s16x2(lo3[k], lo3[(k + 3) & 15], lo3[(k + 6) & 15]);
    unsigned best = 0;
#pragma unroll
    for (int k = 0; k < 16; k += 2) best = __vimax3_s16x2(best, lo9[k], lo9[k + 1]);
    const int sc = max((int)(best & 0xFFFFu), (int)(best >> 16)) - 257;
    return sc >= th ? sc : 0;
}

// FAST pre-test for two horizontally adjacent pixels held in the 16-bit lanes of one register.
// Necessary condition for a FAST-9 corner: every arc of 9 contains two adjacent compass points (ring 0, 4, 8, 12),
// so two adjacent compass points must both be darker (or both brighter) than the centre by more than th.
//   c2        {256 + v} per lane;  down/right/up/left: the compass pixels, one per lane
//   t_hi = 257 + th, t_lo = 255 - th (replicated): with d = 256 + v - ring,  v - ring > th <=> d >= t_hi,
//   v - ring < -th <=> d <= t_lo
__device__ __forceinline__ void fast_pretest2(unsigned c2, unsigned down, unsigned right, unsigned up, unsigned left,
                                              unsigned t_hi, unsigned t_lo, bool& cand_lo, bool& cand_hi) {
    const unsigned d0 = c2 - down, d4 = c2 - right, d8 = c2 - up, d12 = c2 - left;   // lanes stay in 1..511: no borrow
    // an adjacent compass pair is (0 or 8) with (4 or 12): "both beyond the threshold" for some adjacent pair
    // <=> (d0 or d8 beyond) and (d4 or d12 beyond) <=> min(max(d0,d8), max(d4,d12)) >= t_hi, resp. max(min, min) <= t_lo
    const unsigned mn = __vmins2(__vmaxs2(d0, d8), __vmaxs2(d4, d12));
    const unsigned mx = __vmaxs2(__vmins2(d0, d8), __vmins2(d4, d12));
    bool ph, pl, qh, ql;
    (void)__vibmax_s16x2(mn, t_hi, &ph, &pl);    // pred = (mn >= t_hi) per lane
    (void)__vibmin_s16x2(mx, t_lo, &qh, &ql);    // pred = (mx <= t_lo) per lane
    cand_lo = pl || ql;
    cand_hi = ph || qh;
}

__device__ __forceinline__ unsigned widen_lo(unsigned w) { return __byte_perm(w, 0u, 0x4140); }   // bytes 0,1 -> 16-bit lanes
__device__ __forceinline__ unsigned widen_hi(unsigned w) { return __byte_perm(w, 0u, 0x4342); }   // bytes 2,3 -> 16-bit lanes

// a staged word that straddles the image border (or an unaligned image): byte loads with REFLECT_101; kept out of
// line so that the common word path of border tiles stays lean
__device__ __noinline__ uint32_t load_border_word(const uint8_t* __restrict__ row, int gx0, int W) {
    return (uint32_t)__ldg(row + reflect1(gx0, W)) | ((uint32_t)__ldg(row + reflect1(gx0 + 1, W)) << 8) |
           ((uint32_t)__ldg(row + reflect1(gx0 + 2, W)) << 16) | ((uint32_t)__ldg(row + reflect1(gx0 + 3, W)) << 24);
}

// ---- TMA helpers (sm_100a): one thread arms an mbarrier with the byte count and issues a 3-D tiled bulk copy
//      global -> shared; every thread waits on the barrier's phase before reading the tile -------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // make the initialised barrier visible to the async proxy
}
__device__ __forceinline__ void tma_load_tile3d(void* smem_dst, const CUtensorMap* tm, int x, int y, int z, uint64_t* bar, unsigned bytes) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(z), "r"(b)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    unsigned done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(b), "r"(parity) : "memory");
    } while (!done);
}

// one lane's add on a shared counter, as the single ATOMS it is (atomicAdd() under `lane == 0` makes the compiler wrap it in
// its own vote / find-leader / popc aggregation, ~17 instructions per call; the warps here have aggregated by ballot already)
__device__ __forceinline__ int smem_add(int* p, int v) {
    int old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
    return old;
}

// USE_TMA: the tile (+ halo) comes in with one cp.async.bulk.tensor per CTA from a tensor map over the batch's images
// (zero fill outside the image; the few halo bytes that REFLECT_101 needs are patched afterwards, border tiles only).
// Needs 16-byte aligned images / strides; launch_detect falls back to the load loop otherwise.
template <bool USE_TMA>
__global__ void __launch_bounds__(K1_THREADS, 6)
fast_blur_kernel(Pools P, const uint8_t* __restrict__ imgL, const uint8_t* __restrict__ imgR,
                 size_t stride, size_t image_pitch, int pitch_out, const __grid_constant__ CUtensorMap tmL,
                 const __grid_constant__ CUtensorMap tmR) {
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ __align__(128) uint32_t s_inw[SIN_H * SIN_PW];
    __shared__ __align__(16) float s_row[(TH + 6) * TW];
    __shared__ __align__(16) uint32_t s_zero[SC_H * SIN_PW + TH * (TW / 4)];   // zero-initialised planes:
    uint32_t* s_scw = s_zero;                                     //   scores of the candidates (tile + 1-px ring)
    uint32_t* s_nmsw = s_zero + SC_H * SIN_PW;                    //   scores after NMS + border filter (tile only)
    static_assert((SC_H * SIN_PW + TH * (TW / 4)) % 4 == 0 && (SC_H * SIN_PW) % 4 == 0, "uint4 zeroing / aligned planes");
    __shared__ uint16_t s_list[K1_NCAND];                         // pre-test survivors  (row * SIN_P + column)
    __shared__ uint16_t s_corner[K1_NCAND];                       // corners (score > 0)
    __shared__ int s_n[2];
    uint32_t* s_inl = s_inw + SIN_XOFF / 4;                       // logical origin of the staged tile (column x0 - HALO)
    const uint8_t* s_in = reinterpret_cast<const uint8_t*>(s_inl);
    uint8_t* s_sc = reinterpret_cast<uint8_t*>(s_scw);

    const int W = P.p.W, H = P.p.H;
    const int item = blockIdx.z >> 1, side = blockIdx.z & 1;
    const int slot = P.slots[item];
    const uint8_t* __restrict__ img = (side ? imgR : imgL) + (size_t)item * image_pitch;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned lt_mask = (1u << lane) - 1u;

    // 1. stage the tile with halo 4 (REFLECT_101 outside the image: only the blur looks there)
    if (USE_TMA) {
        if (tid == 0) mbar_init(&s_bar, 1);
        __syncthreads();
        if (tid == 0) {      // (no pointer select between the two __grid_constant__ maps: each branch names its parameter)
            if (side) tma_load_tile3d(s_inw, &tmR, x0 - HALO - SIN_XOFF, y0 - HALO, item, &s_bar, SIN_H * SIN_P);
            else tma_load_tile3d(s_inw, &tmL, x0 - HALO - SIN_XOFF, y0 - HALO, item, &s_bar, SIN_H * SIN_P);
        }
        for (int w = tid; w < (SC_H * SIN_PW + TH * (TW / 4)) / 4; w += K1_THREADS)      // overlaps the copy
            reinterpret_cast<uint4*>(s_zero)[w] = make_uint4(0, 0, 0, 0);
        if (tid < 2) s_n[tid] = 0;
        mbar_wait(&s_bar, 0);
        // staged rows / columns outside the image arrived as zeros: patch them with the reflected pixels
        const int c_lo = max(0, HALO - x0), c_hi = min(SIN_W, W - (x0 - HALO));
        const int r_lo = max(0, HALO - y0), r_hi = min(SIN_H, H - (y0 - HALO));
        if (c_lo > 0 || c_hi < SIN_W || r_lo > 0 || r_hi < SIN_H) {
            __syncthreads();
            uint8_t* s_inb = reinterpret_cast<uint8_t*>(s_inl);
            // rows above / below the image (whole staged rows), then columns left / right of it (the corners twice: harmless)
            const int n_rows = r_lo + (SIN_H - r_hi), n_cols = c_lo + (SIN_W - c_hi);
            for (int i = tid; i < n_rows * SIN_W; i += K1_THREADS) {
                int r = i / SIN_W;
                const int c = i - r * SIN_W;
                if (r >= r_lo) r += r_hi - r_lo;
                s_inb[r * SIN_P + c] = __ldg(img + (size_t)reflect1(y0 - HALO + r, H) * stride + reflect1(x0 - HALO + c, W));
            }
            for (int i = tid; i < n_cols * SIN_H; i += K1_THREADS) {
                const int r = i / n_cols;
                int c = i - r * n_cols;
                if (c >= c_lo) c += c_hi - c_lo;
                s_inb[r * SIN_P + c] = __ldg(img + (size_t)reflect1(y0 - HALO + r, H) * stride + reflect1(x0 - HALO + c, W));
            }
        }
    } else {
        const bool aligned = ((reinterpret_cast<uintptr_t>(img) | stride) & 3) == 0;
        if (aligned && x0 >= HALO && x0 + TW + HALO <= W && y0 >= HALO && y0 + TH + HALO <= H) {
            const uint8_t* base = img + (size_t)(y0 - HALO) * stride + (x0 - HALO);      // interior tile: plain word copies
            for (int w = tid; w < SIN_H * SIN_WW; w += K1_THREADS) {
                const int r = w / SIN_WW, cw = w - r * SIN_WW;
                s_inl[r * SIN_PW + cw] = __ldg(reinterpret_cast<const uint32_t*>(base + (size_t)r * stride) + cw);
            }
        } else {
            for (int w = tid; w < SIN_H * SIN_WW; w += K1_THREADS) {
                const int r = w / SIN_WW, cw = w - r * SIN_WW;
                const int gy = reflect1(y0 - HALO + r, H), gx0 = x0 - HALO + 4 * cw;
                const uint8_t* row = img + (size_t)gy * stride;
                s_inl[r * SIN_PW + cw] = (aligned && gx0 >= 0 && gx0 + 3 < W) ? __ldg(reinterpret_cast<const uint32_t*>(row + gx0))
                                                                               : load_border_word(row, gx0, W);
            }
        }
        for (int w = tid; w < (SC_H * SIN_PW + TH * (TW / 4)) / 4; w += K1_THREADS)
            reinterpret_cast<uint4*>(s_zero)[w] = make_uint4(0, 0, 0, 0);
        if (tid < 2) s_n[tid] = 0;
    }
    __syncthreads();

    // 2. blur, row pass: rows y0-3 .. y0+TH+2, four outputs per work item (float32, sequential FMA left to right)
    for (int i = tid; i < (TH + 6) * (TW / 4); i += K1_THREADS) {
        const int r = i / (TW / 4), c = (i - r * (TW / 4)) * 4;
        const uint32_t* pw = s_inl + (r + 1) * SIN_PW + (c >> 2);   // bytes c .. c+11
        const uint32_t w0 = pw[0], w1 = pw[1], w2 = pw[2];
        float f[10];                                   // columns c+1 .. c+10  (x-3 .. x+6 of the first output)
        f[0] = (float)((w0 >> 8) & 0xFF); f[1] = (float)((w0 >> 16) & 0xFF); f[2] = (float)(w0 >> 24);
        f[3] = (float)(w1 & 0xFF); f[4] = (float)((w1 >> 8) & 0xFF); f[5] = (float)((w1 >> 16) & 0xFF); f[6] = (float)(w1 >> 24);
        f[7] = (float)(w2 & 0xFF); f[8] = (float)((w2 >> 8) & 0xFF); f[9] = (float)((w2 >> 16) & 0xFF);
        float4 o;
        float* op = &o.x;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float acc = __fmul_rn(gk(0), f[q]);
#pragma unroll
            for (int t = 1; t < 7; ++t) acc = __fmaf_rn(gk(t), f[q + t], acc);
            op[q] = acc;
        }
        *reinterpret_cast<float4*>(s_row + r * TW + c) = o;
    }

    // 3a. FAST pre-test on the tile plus a 1-px ring (needed by the 3x3 NMS): one work item = one staged word
    //     (4 pixels) of rows -1 .. TH; survivors are appended to s_list (order is irrelevant)
    const int th = slot_fast_th(P.p, P.state[slot]);     // m_current_fast_th of this estimator (S2:510)
    {
        const unsigned t_hi = (unsigned)(257 + th) * 0x00010001u;
        const unsigned t_lo = ((unsigned)(255 - th) & 0xFFFFu) * 0x00010001u;
        // candidate columns: FAST is evaluated for 3 <= x < W-3, the tile needs x0-1 .. x0+TW
        const int cx_lo = max(3, x0 - 1) - (x0 - HALO), cx_hi = min(W - 3, x0 + TW + 1) - (x0 - HALO);   // staged columns [cx_lo, cx_hi)
        for (int i0 = 0; i0 < K1_ITEMS3; i0 += K1_THREADS) {
            if (i0 + (tid & ~31) >= K1_ITEMS3) break;      // whole warp beyond the work items
            const int i = i0 + tid;
            bool c0 = false, c1 = false, c2 = false, c3 = false;
            int r = 0, cw = 0;
            if (i < K1_ITEMS3) {
                r = i / SIN_WW; cw = i - r * SIN_WW;
                const int gy = y0 - 1 + r;
                const int jlo = cx_lo - 4 * cw, jhi = cx_hi - 4 * cw;      // valid pixels of this word: jlo <= j < jhi
                if (gy >= 3 && gy < H - 3 && jlo < 4 && jhi > 0) {
                    const uint32_t* rp = s_inl + (r + 3) * SIN_PW + cw;
                    const uint32_t C = rp[0], Cm = rp[-1], Cp = rp[1], U = rp[-3 * SIN_PW], D = rp[3 * SIN_PW];
                    const uint32_t L = __funnelshift_r(Cm, C, 8);      // pixels x-3
                    const uint32_t R = __funnelshift_r(C, Cp, 24);     // pixels x+3
                    fast_pretest2(widen_lo(C) + 0x01000100u, widen_lo(D), widen_lo(R), widen_lo(U), widen_lo(L), t_hi, t_lo, c0, c1);
                    fast_pretest2(widen_hi(C) + 0x01000100u, widen_hi(D), widen_hi(R), widen_hi(U), widen_hi(L), t_hi, t_lo, c2, c3);
                    c0 = c0 && jlo <= 0 && jhi > 0;
                    c1 = c1 && jlo <= 1 && jhi > 1;
                    c2 = c2 && jlo <= 2 && jhi > 2;
                    c3 = c3 && jlo <= 3 && jhi > 3;
                }
            }
            const unsigned b0 = __ballot_sync(0xffffffffu, c0), b1 = __ballot_sync(0xffffffffu, c1);
            const unsigned b2 = __ballot_sync(0xffffffffu, c2), b3 = __ballot_sync(0xffffffffu, c3);
            const int n0 = __popc(b0), n1 = __popc(b1), n2 = __popc(b2), n3 = __popc(b3);
            int base = 0;
            if (lane == 0 && (n0 + n1 + n2 + n3) > 0) base = smem_add(&s_n[0], n0 + n1 + n2 + n3);
            base = __shfl_sync(0xffffffffu, base, 0);
            const int e = r * SIN_P + 4 * cw;
            if (c0) s_list[base + __popc(b0 & lt_mask)] = (uint16_t)e;
            if (c1) s_list[base + n0 + __popc(b1 & lt_mask)] = (uint16_t)(e + 1);
            if (c2) s_list[base + n0 + n1 + __popc(b2 & lt_mask)] = (uint16_t)(e + 2);
            if (c3) s_list[base + n0 + n1 + n2 + __popc(b3 & lt_mask)] = (uint16_t)(e + 3);
        }
    }
    __syncthreads();
    // 3b. full score for the candidates only (dense over the CTA); corners go to a second list
    {
        const int nlist = s_n[0];
        for (int q0 = 0; q0 < nlist; q0 += K1_THREADS) {
            if (q0 + (tid & ~31) >= nlist) break;          // whole warp beyond the list
            const int q = q0 + tid;
            int sc = 0, e = 0;
            if (q < nlist) {
                e = s_list[q];
                sc = fast9_score(s_in + e + 3 * SIN_P, th);
                if (sc > 0) s_sc[e] = (uint8_t)sc;
            }
            const unsigned b = __ballot_sync(0xffffffffu, sc > 0);
            int base = 0;
            if (lane == 0 && b) base = smem_add(&s_n[1], __popc(b));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (sc > 0) s_corner[base + __popc(b & lt_mask)] = (uint16_t)e;
        }
    }
    __syncthreads();
    // 3c. strict 3x3 NMS (cv::FAST, nonmaxSuppression = true) + the 31-px border filter of ORB::compute, tile pixels only
    {
        const int ncorner = s_n[1];
        uint8_t* s_nms = reinterpret_cast<uint8_t*>(s_nmsw);
        for (int q = tid; q < ncorner; q += K1_THREADS) {
            const int e = s_corner[q];
            const int r = e / SIN_P, c = e - r * SIN_P;           // r: 1 .. TH inside the tile, c: HALO .. HALO+TW-1
            const int gx = x0 - HALO + c, gy = y0 - 1 + r;
            if (r < 1 || r > TH || c < HALO || c >= HALO + TW) continue;
            if (gx < kOrbEdge || gx >= W - kOrbEdge || gy < kOrbEdge || gy >= H - kOrbEdge) continue;
            const uint8_t* sp = s_sc + e;
            const int sv = sp[0];
            const bool keep = sv > sp[-SIN_P - 1] && sv > sp[-SIN_P] && sv > sp[-SIN_P + 1] && sv > sp[-1] && sv > sp[1] &&
                              sv > sp[SIN_P - 1] && sv > sp[SIN_P] && sv > sp[SIN_P + 1];
            if (keep) s_nms[(r - 1) * TW + (c - HALO)] = (uint8_t)sv;
        }
    }
    __syncthreads();

    // 4. blur column pass + stores; 4 pixels per work item, 32-bit stores
    uint8_t* __restrict__ blur_out = P.blur + img_idx(slot, side) * (size_t)pitch_out * H;
    for (int i = tid; i < TH * TW / 4; i += K1_THREADS) {
        const int r = i / (TW / 4), c4 = (i - r * (TW / 4)) * 4;
        const int gy = y0 + r, gx = x0 + c4;
        if (gy >= H || gx >= pitch_out) continue;
        const float* rr0 = s_row + (r + 3) * TW + c4;
        const float4 m0 = *reinterpret_cast<const float4*>(rr0);
        const float4 p1 = *reinterpret_cast<const float4*>(rr0 + TW), n1 = *reinterpret_cast<const float4*>(rr0 - TW);
        const float4 p2 = *reinterpret_cast<const float4*>(rr0 + 2 * TW), n2 = *reinterpret_cast<const float4*>(rr0 - 2 * TW);
        const float4 p3 = *reinterpret_cast<const float4*>(rr0 + 3 * TW), n3 = *reinterpret_cast<const float4*>(rr0 - 3 * TW);
        const float c0[4] = {m0.x, m0.y, m0.z, m0.w};
        const float a1[4] = {p1.x, p1.y, p1.z, p1.w}, b1[4] = {n1.x, n1.y, n1.z, n1.w};
        const float a2[4] = {p2.x, p2.y, p2.z, p2.w}, b2[4] = {n2.x, n2.y, n2.z, n2.w};
        const float a3[4] = {p3.x, p3.y, p3.z, p3.w}, b3[4] = {n3.x, n3.y, n3.z, n3.w};
        uint32_t bw = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float acc = __fmul_rn(gk(3), c0[q]);
            acc = __fmaf_rn(gk(4), __fadd_rn(a1[q], b1[q]), acc);
            acc = __fmaf_rn(gk(5), __fadd_rn(a2[q], b2[q]), acc);
            acc = __fmaf_rn(gk(6), __fadd_rn(a3[q], b3[q]), acc);
            uint32_t v;                                  // round half to even, saturate to 0..255
            asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(v) : "f"(acc));
            bw |= v << (8 * q);
        }
        *reinterpret_cast<uint32_t*>(blur_out + (size_t)gy * pitch_out + gx) = bw;
    }

    // 5. the tile's corners as compact lists in row-major order: thread t owns pixels 8t .. 8t+7 of the tile (row t / 8), a
    //    warp owns a band of kBandH rows with its own kBandCap entries of the tile's list (strict 3x3 maxima: at most one per
    //    2x2 block, so a band cannot overflow) - a warp scan places the corners, no block-wide step.  Every image row of
    //    the tile (a "segment") records where its corners start inside the band's list and how many they are.
    {
        static_assert(K1_THREADS * 8 == TW * TH && 32 * 8 == kBandH * TW, "one thread per 8 pixels, one warp per band");
        const uint32_t w0 = s_nmsw[2 * tid], w1 = s_nmsw[2 * tid + 1];
        uint32_t m0 = nonzero_bytes(w0), m1 = nonzero_bytes(w1);
        const int cnt = __popc(m0) + __popc(m1);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const size_t tile = ((size_t)img_idx(slot, side) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        const int row_lo = __shfl_up_sync(0xffffffffu, incl, 8);      // inclusive count at the end of the previous row
        if ((tid & 7) == 7) {
            const int lo = lane >= 8 ? row_lo : 0;
            const size_t y = (size_t)blockIdx.y * TH + (tid >> 3);
            P.tile_seg[((size_t)img_idx(slot, side) * gridDim.y * TH + y) * gridDim.x + blockIdx.x] = (uint16_t)((lo << 8) | (incl - lo));
        }
        uint16_t* __restrict__ lst = P.tile_list + tile * kTileCap + (tid >> 5) * kBandCap + (incl - cnt);
        const int xb = (tid & 7) * 8;
        while (m0) {
            const int bit = __ffs(m0) - 1;               // 7, 15, 23 or 31
            m0 &= m0 - 1;
            *lst++ = (uint16_t)((((w0 >> (bit - 7)) & 0xFFu) << 8) | (uint32_t)(xb + (bit >> 3)));
        }
        while (m1) {
            const int bit = __ffs(m1) - 1;
            m1 &= m1 - 1;
            *lst++ = (uint16_t)((((w1 >> (bit - 7)) & 0xFFu) << 8) | (uint32_t)(xb + 4 + (bit >> 3)));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K2  select_kernel: one CTA per image.
//   A. row-major ordered compaction of the score map  -> candidate words  score<<24 | ordinal
//   B. std::sort emulation by response (descending)                       S2:322-324
//   C. greedy grid occupancy, first `quota` survivors                     S2:331-369
//   D. std::sort emulation by row (ascending)                             S2:87-89
//   E. write keypoints in final order
// ------------------------------------------------------------------------------------------------
constexpr int K2_THREADS = 256;
constexpr int K2_SEGS = 4;          // 16-pixel score segments per thread per compaction round

// rank-based stable sort (diagnostic sort_mode = 1): O(n^2 / threads)
template <bool Desc, int Shift>
__device__ void block_stable_sort(uint32_t* a, uint32_t* tmp, int n) {
    const KeyComp<Desc, Shift> C;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t v = a[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const uint32_t w = a[j];
            rank += (C.lt(w, v) || (!C.lt(v, w) && j < i)) ? 1 : 0;
        }
        tmp[rank] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) a[i] = tmp[i];
    __syncthreads();
}

__global__ void __launch_bounds__(K2_THREADS)
select_kernel(Pools P, int pitch, uint32_t* __restrict__ occ_scratch, int occ_words_per_img,
              uint32_t* __restrict__ sort_scratch, uint2* __restrict__ list_scratch, uint32_t* __restrict__ key_scratch,
              int smem_corners) {
    extern __shared__ __align__(16) uint32_t smem[];
    // The candidate words are compacted into global scratch first; they are sorted in shared memory when they fit
    // into smem_corners words (4 CTAs per SM at the default sizes), in the global array otherwise (same code, slower).
    uint32_t* g_a = key_scratch + (size_t)blockIdx.x * P.p.max_corners;
    uint32_t* s_b = smem + smem_corners;           // [kp_cap]
    // segment lists of the sort emulation live in global memory (one read + one write per segment): keeps the
    // shared-memory footprint at 4 B per candidate so that three CTAs fit on an SM
    const int list_cap = P.p.max_corners / 16 + 2;
    uint2* s_listA = list_scratch + (size_t)blockIdx.x * 3 * list_cap;
    uint2* s_listB = s_listA + list_cap;
    uint2* s_listS = s_listB + list_cap;
    __shared__ int s_warp[K2_THREADS / 32];
    __shared__ int s_nsel;
    __shared__ int s_cnt2[3];
    __shared__ int s_pos[64 * (K2_THREADS / 32)];
    uint32_t* tmp = sort_scratch + (size_t)blockIdx.x * P.p.max_corners;

    const int W = P.p.W, H = P.p.H;
    const int item = blockIdx.x >> 1, side = blockIdx.x & 1;
    const int slot = P.slots[item];
    const int tid = threadIdx.x;
    uint32_t* __restrict__ cand_pos = P.cand_pos + img_idx(slot, side) * (size_t)P.p.max_corners;

    // ---- A. the row-major candidate list from fast_blur_kernel's per-tile lists --------------------------
    // a segment = the corners of one image row inside one tile; segments in (row, tile column) order are the row-major
    // order of cv::FAST's keypoint list.  Every thread takes a contiguous run of segments, a block scan places them.
    const int tiles_x = tiles_x_of(W), tiles_y = tiles_y_of(H);
    const uint16_t* __restrict__ seg = P.tile_seg + img_idx(slot, side) * (size_t)tiles_x * tiles_y * kTileH;
    const uint16_t* __restrict__ tlst = P.tile_list + img_idx(slot, side) * (size_t)tiles_x * tiles_y * kTileCap;
    const CornerRun run = corner_run(seg, H * tiles_x, K2_THREADS, tid);
    int total;
    const int o = block_exclusive_scan<K2_THREADS>(run.cnt, s_warp, total);
    const bool overflow = total > P.p.max_corners;
    const int n = min(total, P.p.max_corners);
    // sorted in shared memory when the candidates fit into smem_corners words, in the global array otherwise
    uint32_t* s_a = n <= smem_corners ? smem : g_a;
    {
        uint32_t* stage = n <= smem_corners ? smem : tmp;
        corner_codes(seg, run, tiles_x, o, n, stage);
        __syncthreads();
        for (int i = tid; i < n; i += K2_THREADS) {
            const Corner c = corner_fetch(tlst, stage[i], tiles_x);
            s_a[i] = ((uint32_t)c.score << 24) | (uint32_t)i;
            cand_pos[i] = ((uint32_t)c.y << 16) | (uint32_t)c.x;
        }
    }
    if (tid == 0) {
        P.n_cand[img_idx(slot, side)] = n;
        if (overflow) P.state[slot].status = -3;  // SVO_ERR_CAPACITY
    }
    __syncthreads();

    const int cur = P.state[slot].cur;
    int2* __restrict__ kp_out = P.kp + frame_idx(slot, cur, side) * (size_t)P.p.kp_cap;
    int n_sel;

    if (P.p.nms) {
        // ---- B. sort by response, descending (KeypointResponseSorter) ------------------------------
        const unsigned cell = (unsigned)(P.p.min_distance / 2.0);     // S2:331
        // with 1-px cells the selection is the first `quota` entries of the sorted order (see C below): only that prefix
        // of std::sort's result is computed; the greedy grid walks further down the list and needs all of it
        const int keep = cell == 1 ? min(n, P.p.quota) : n;
        if (P.p.sort_mode == 0) introsort_block(KeyComp<true, 24>(), s_a, n, s_listA, s_listB, s_listS, s_cnt2, tmp, s_pos, keep);
        else block_stable_sort<true, 24>(s_a, tmp, n);
        // ---- C. greedy occupancy grid ---------------------------------------------------------------
        if (cell == 1) {
            // FAST's strict 3x3 NMS leaves no two corners in 8-adjacent pixels, and with 1-px cells a
            // corner only blocks its own pixel and the 4 neighbours: nothing is ever rejected.
            n_sel = min(n, P.p.quota);
        } else {
            const float cell_inv = 1.0f / (float)cell;                 // S2:332
            const unsigned glx = (unsigned)(1 + W * cell_inv), gly = (unsigned)(1 + H * cell_inv);
            uint32_t* occ = occ_scratch + (size_t)blockIdx.x * occ_words_per_img;
            for (unsigned i = tid; i < (glx * gly + 31) / 32; i += K2_THREADS) occ[i] = 0;
            __syncthreads();
            if (tid == 0) {
                auto get = [&](unsigned x, unsigned y) { const unsigned b = x * gly + y; return (occ[b >> 5] >> (b & 31)) & 1u; };
                auto set = [&](unsigned x, unsigned y) { const unsigned b = x * gly + y; occ[b >> 5] |= 1u << (b & 31); };
                int c = 0, k = 0;
                while (c < P.p.quota && k < n) {
                    const uint32_t wd = s_a[k++];
                    const uint32_t pos = cand_pos[wd & 0xFFFFFFu];
                    const unsigned sx = (unsigned)((float)(pos & 0xFFFFu) * cell_inv);
                    const unsigned sy = (unsigned)((float)(pos >> 16) * cell_inv);
                    if (get(sx, sy)) continue;
                    set(sx, sy);
                    if (sx > 0) set(sx - 1, sy);
                    if (sy > 0) set(sx, sy - 1);
                    if (sx < glx - 1) set(sx + 1, sy);
                    if (sy < gly - 1) set(sx, sy + 1);
                    s_a[c++] = wd;   // c <= k-1: in-place compaction of the survivors
                }
                s_nsel = c;
            }
            __syncthreads();
            n_sel = s_nsel;
        }
    } else {
        n_sel = min(n, P.p.kp_cap);   // feats_vector.swap(...)  S2:613
    }

    // ---- D. sort by row (KpRowSorter), ascending ----------------------------------------------------
    for (int k = tid; k < n_sel; k += K2_THREADS) {
        const uint32_t pos = cand_pos[s_a[k] & 0xFFFFFFu];
        s_b[k] = (pos & 0xFFFF0000u) | (uint32_t)k;   // y << 16 | position in the selection order
    }
    __syncthreads();
    if (P.p.sort_mode == 0) introsort_block(KeyComp<false, 16>(), s_b, n_sel, s_listA, s_listB, s_listS, s_cnt2, tmp, s_pos, n_sel);
    else block_stable_sort<false, 16>(s_b, tmp, n_sel);
    // ---- E. emit ------------------------------------------------------------------------------------
    for (int i = tid; i < n_sel; i += K2_THREADS) {
        const uint32_t wd = s_a[s_b[i] & 0xFFFFu];
        kp_out[i] = make_int2((int)cand_pos[wd & 0xFFFFFFu], (int)(wd >> 24));
    }
    if (tid == 0) {
        P.n_kp[frame_idx(slot, cur, side)] = n_sel;
        P.oct_stats[stat_idx(slot, P.octave, side)] = n_sel;
        if (P.octave == 0) {                     // svo_result mirrors detected_feats[0]
            if (side == 0) P.result[slot].detected_left = n_sel; else P.result[slot].detected_right = n_sel;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K3  brief_kernel: one warp per keypoint, lane l produces descriptor byte l (tests 8l..8l+7).
//   bit i = blur[y+ay_i][x+ax_i] < blur[y+by_i][x+bx_i]      (cv::ORB::compute, S2:512)
// ------------------------------------------------------------------------------------------------
constexpr int K3_WARPS = 8;
constexpr int K3_ITERS = 8;         // keypoints per warp
constexpr int K3_PITCH = 48;        // patch row pitch in bytes: a 48-byte TMA box row (x origin rounded down to 16 bytes)
constexpr int K3_PBYTES = 31 * K3_PITCH;             // 1488 bytes per staged patch
constexpr int K3_PSLOT = (K3_PBYTES + 127) & ~127;   // 1536: TMA destinations are 128-byte aligned

// Sample offsets inside the staged patch, built at compile time from the ORB pattern and stored transposed
// ([test-in-byte][lane]) so that the 32 lanes of a warp read consecutive words.
struct BriefOffsets { uint32_t v[8 * 32]; };
constexpr BriefOffsets make_brief_offsets() {
    constexpr int pat[256][4] = {
#include "orb_pattern.inc"
    };
    BriefOffsets t{};
    for (int lane = 0; lane < 32; ++lane)
        for (int k = 0; k < 8; ++k) {
            const int i = lane * 8 + k;
            const uint32_t a = (uint32_t)((pat[i][1] + 15) * K3_PITCH + pat[i][0] + 15);
            const uint32_t b = (uint32_t)((pat[i][3] + 15) * K3_PITCH + pat[i][2] + 15);
            t.v[k * 32 + lane] = a | (b << 16);
        }
    return t;
}
__device__ const BriefOffsets g_brief_offsets = make_brief_offsets();

// The 512 samples of one keypoint lie within |offset| <= 13 of it.  The 31-row patch around the keypoint is staged in
// shared memory and every lane reads its 16 samples from there (instead of 16 global byte gathers per lane).
//   USE_TMA : lane 0 issues one cp.async.bulk.tensor.3d per keypoint (48 x 31 box of the blur plane, x origin rounded
//             down to 16 bytes) into one of two buffers, one keypoint ahead of the warp; all lanes wait on the mbarrier.
//   else    : the warp copies the patch with aligned 32-bit loads (3 rows x 9 words per instruction).
template <bool USE_TMA>
__global__ void __launch_bounds__(K3_WARPS * 32)
brief_kernel(Pools P, int pitch, const __grid_constant__ CUtensorMap tmBlur) {
    __shared__ __align__(128) uint8_t s_patch[K3_WARPS][2][K3_PSLOT];
    __shared__ __align__(8) uint64_t s_bar[K3_WARPS][2];
    const int item = blockIdx.y >> 1, side = blockIdx.y & 1;
    const int slot = P.slots[item];
    const int cur = P.state[slot].cur;
    const size_t f = frame_idx(slot, cur, side);
    const int n = P.n_kp[f];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t off[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) off[t] = g_brief_offsets.v[t * 32 + lane];
    const int k0 = (blockIdx.x * K3_WARPS + w) * K3_ITERS;
    const int2* __restrict__ kps = P.kp + f * P.p.kp_cap;
    uint8_t* __restrict__ desc = reinterpret_cast<uint8_t*>(P.desc + f * P.p.kp_cap * 2);
    if (USE_TMA) {
        const int img = (int)img_idx(slot, side);
        if (lane == 0) { mbar_init(&s_bar[w][0], 1); mbar_init(&s_bar[w][1], 1); }
        __syncwarp();
        auto issue = [&](int k, int buf) {              // lane 0: patch of keypoint k -> buffer buf
            const int2 kp = kps[k];
            const int xs = (kp.x & 0xFFFF) - 15, y = (int)((unsigned)kp.x >> 16);
            tma_load_tile3d(s_patch[w][buf], &tmBlur, xs & ~15, y - 15, img, &s_bar[w][buf], K3_PBYTES);
        };
        if (lane == 0 && k0 < n) issue(k0, 0);
        for (int it = 0; it < K3_ITERS; ++it) {
            const int k = k0 + it;
            if (k >= n) break;
            const int buf = it & 1;
            if (lane == 0 && it + 1 < K3_ITERS && k + 1 < n) issue(k + 1, buf ^ 1);     // its last readers passed the __syncwarp below
            const int xs = (kps[k].x & 0xFFFF) - 15;
            mbar_wait(&s_bar[w][buf], (unsigned)(it >> 1) & 1u);
            const uint8_t* patch = s_patch[w][buf] + (xs & 15);
            unsigned v = 0;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int pa = patch[off[t] & 0xFFFFu], pb = patch[off[t] >> 16];
                v |= (unsigned)(pa < pb) << t;
            }
            desc[(size_t)k * 32 + lane] = (uint8_t)v;
            __syncwarp();
        }
    } else {
        const int lr = lane / 9, lw = lane - lr * 9;
        uint32_t* patchw = reinterpret_cast<uint32_t*>(s_patch[w][0]);
        const uint8_t* __restrict__ plane = P.blur + img_idx(slot, side) * (size_t)pitch * P.p.H;
        for (int it = 0; it < K3_ITERS; ++it) {
            const int k = k0 + it;
            if (k >= n) break;
            const int2 kp = kps[k];
            const int x = kp.x & 0xFFFF, y = (unsigned)kp.x >> 16;
            // rows y-15 .. y+15; columns from (x-15) rounded down to a multiple of 4 (planes are 16-byte aligned with a
            // pitch that is a multiple of 16; keypoints are >= 31 px inside the image, so the patch never leaves it)
            const int xs = x - 15, a = xs & 3;
            const uint8_t* __restrict__ src = plane + (size_t)(y - 15) * pitch + (xs - a);
            __syncwarp();
            if (lane < 27) {
#pragma unroll
                for (int r3 = 0; r3 < 11; ++r3) {
                    const int row = r3 * 3 + lr;
                    if (row < 31) patchw[row * (K3_PITCH / 4) + lw] = __ldg(reinterpret_cast<const uint32_t*>(src + (size_t)row * pitch) + lw);
                }
            }
            __syncwarp();
            const uint8_t* patch = s_patch[w][0] + a;
            unsigned v = 0;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int pa = patch[off[t] & 0xFFFFu], pb = patch[off[t] >> 16];
                v |= (unsigned)(pa < pb) << t;
            }
            desc[(size_t)k * 32 + lane] = (uint8_t)v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K0  begin_step_kernel: the prev/cur shift rule of processNewImagePair (P:86-100)
// ------------------------------------------------------------------------------------------------
__global__ void begin_step_kernel(Pools P, int n, const uint8_t* __restrict__ repeat) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int slot = P.slots[i];
    SlotState& s = P.state[slot];
    const int rep = repeat ? repeat[i] : 0;
    if (!rep && s.m_error != 5 /*voecBadTracking*/ && s.m_error != 1 /*voecBadCondNumber*/) {
        // m_prev_imgpair = m_current_imgpair: the buffers swap roles, the old "prev" gets overwritten
        s.has_prev = s.has_cur;
        s.cur ^= 1;
    }
    s.repeat = rep;
    s.m_error = 0;
    s.status = 0;
    s.has_cur = 1;
    Result& r = P.result[slot];
    r.error_code = 0;
    r.valid = 0;
    r.num_it = r.num_it_final = 0;
    r.n_outliers = 0;
    r.n_residuals = 0;
    r.tracked_frame = 0;
    r.tracked_kf = 0;
    r.status = 0;
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
// candidate words sorted in shared memory: at most 12288 (48 KB) so that four CTAs share an SM
int detect_smem_corners(const Params& p) { return p.max_corners < 12288 ? p.max_corners : 12288; }
int detect_smem_bytes(const Params& p) {
    return ((detect_smem_corners(p) + p.kp_cap + 1) & ~1) * (int)sizeof(uint32_t);
}

cudaError_t configure_detect(const Params& p) {
    return SVO_RAISE_DYN_SMEM(select_kernel, detect_smem_bytes(p));
}

cudaError_t launch_begin_step(const Pools& P, int n, const uint8_t* d_repeat, cudaStream_t st) {
    begin_step_kernel<<<(n + 127) / 128, 128, 0, st>>>(P, n, d_repeat);
    return cudaGetLastError();
}

// Tensor map over the n images of one side: dims (x, y, image) of 8-bit pixels, box = one staged tile (SIN_P x SIN_H x 1),
// no swizzle, out-of-bounds elements read as 0.  false if TMA's alignment rules do not hold (or the driver entry point
// is missing): the caller then uses the load-loop kernel.
static bool make_plane_map(CUtensorMap* tm, const uint8_t* d_img, int W, int H, int n, size_t stride, size_t image_pitch, int box_w,
                           int box_h) {
    static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    static bool looked = false;
    if (!looked) {
        looked = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }
    if (!encode || getenv("SVO_NO_TMA")) return false;
    if ((reinterpret_cast<uintptr_t>(d_img) & 15) || (stride & 15) || (image_pitch & 15) || n < 1) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n};
    const cuuint64_t strides[2] = {(cuuint64_t)stride, (cuuint64_t)image_pitch};      // bytes, dims 1 and 2
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(d_img), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static bool make_image_map(CUtensorMap* tm, const uint8_t* d_img, int W, int H, int n, size_t stride, size_t image_pitch) {
    return make_plane_map(tm, d_img, W, H, n, stride, image_pitch, SIN_P, SIN_H);
}
// the pitched internal planes (blur): all (slot, side) images of the context
static bool make_plane_map(CUtensorMap* tm, const uint8_t* d_plane, int pitch, int H, int n_images, int box_w, int box_h) {
    return make_plane_map(tm, d_plane, pitch, H, n_images, (size_t)pitch, (size_t)pitch * H, box_w, box_h);
}

cudaError_t launch_fast_blur_only(const Pools& PL, int n, const uint8_t* d_left, const uint8_t* d_right, size_t stride,
                                  size_t image_pitch, int pitch_out, cudaStream_t st) {
    const dim3 g1((PL.p.W + TW - 1) / TW, (PL.p.H + TH - 1) / TH, 2 * n);
    CUtensorMap tmL, tmR;
    if (make_image_map(&tmL, d_left, PL.p.W, PL.p.H, n, stride, image_pitch) && make_image_map(&tmR, d_right, PL.p.W, PL.p.H, n, stride, image_pitch))
        fast_blur_kernel<true><<<g1, K1_THREADS, 0, st>>>(PL, d_left, d_right, stride, image_pitch, pitch_out, tmL, tmR);
    else {
        memset(&tmL, 0, sizeof(tmL));
        fast_blur_kernel<false><<<g1, K1_THREADS, 0, st>>>(PL, d_left, d_right, stride, image_pitch, pitch_out, tmL, tmL);
    }
    return cudaGetLastError();
}

cudaError_t launch_detect(const Pools& P, int n, const uint8_t* d_left, const uint8_t* d_right, size_t stride,
                          size_t image_pitch, int pitch_out, uint32_t* occ_scratch, int occ_words_per_img,
                          uint32_t* sort_scratch, uint2* list_scratch, uint32_t* key_scratch, cudaStream_t st, int* launches,
                          StepTimer* tm) {
    const dim3 g1((P.p.W + TW - 1) / TW, (P.p.H + TH - 1) / TH, 2 * n);
    CUtensorMap tmL, tmR;
    if (make_image_map(&tmL, d_left, P.p.W, P.p.H, n, stride, image_pitch) && make_image_map(&tmR, d_right, P.p.W, P.p.H, n, stride, image_pitch))
        fast_blur_kernel<true><<<g1, K1_THREADS, 0, st>>>(P, d_left, d_right, stride, image_pitch, pitch_out, tmL, tmR);
    else {
        memset(&tmL, 0, sizeof(tmL));
        fast_blur_kernel<false><<<g1, K1_THREADS, 0, st>>>(P, d_left, d_right, stride, image_pitch, pitch_out, tmL, tmL);
    }
    if (tm) tm->mark(st);
    const int smem = detect_smem_bytes(P.p);
    select_kernel<<<2 * n, K2_THREADS, smem, st>>>(P, pitch_out, occ_scratch, occ_words_per_img, sort_scratch, list_scratch,
                                                   key_scratch, detect_smem_corners(P.p));
    if (tm) tm->mark(st);
    const dim3 g3((P.p.kp_cap + K3_WARPS * K3_ITERS - 1) / (K3_WARPS * K3_ITERS), 2 * n);
    CUtensorMap tmB;
    if (!getenv("SVO_NO_TMA_BRIEF") && make_plane_map(&tmB, P.blur, pitch_out, P.p.H, 2 * (P.p.max_seqs + 1), K3_PITCH, 31))
        brief_kernel<true><<<g3, K3_WARPS * 32, 0, st>>>(P, pitch_out, tmB);
    else {
        memset(&tmB, 0, sizeof(tmB));
        brief_kernel<false><<<g3, K3_WARPS * 32, 0, st>>>(P, pitch_out, tmB);
    }
    if (tm) tm->mark(st);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

}  // namespace svo
