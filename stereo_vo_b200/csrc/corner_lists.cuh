// corner_lists.cuh -- readers of fast_blur_kernel's compact corner lists (svo_internal.cuh: Pools::tile_seg / tile_list),
// shared by select_kernel (detect.cu) and orb_select_kernel (orb.cu).
//
// A segment = the corners of one image row inside one tile column; segments in (row, tile column) order are the row-major
// order of cv::FAST's keypoint list.  The list is rebuilt in three block-wide passes:
//   1. every thread takes a contiguous run of segments and counts its corners (corner_run), a block scan places the runs;
//   2. one word per corner that says where its list entry lives (corner_codes);
//   3. the entries are fetched with independent loads, one corner per thread (corner_fetch) -- instead of a dependent
//      chain of loads per run of segments.
#pragma once
#include <stdint.h>
#include "svo_internal.cuh"

namespace svo {

template <int THREADS>
__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp /* THREADS / 32 ints, shared */, int& total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    int wsum = (lane < THREADS / 32) ? s_warp[lane] : 0;
    int wincl = wsum;
#pragma unroll
    for (int o = 1; o < THREADS / 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, wincl, o);
        if (lane >= o) wincl += t;
    }
    const int wexcl = __shfl_sync(0xffffffffu, wincl - wsum, wid);
    total = __shfl_sync(0xffffffffu, wincl, THREADS / 32 - 1);
    __syncthreads();
    return wexcl + incl - v;
}

struct CornerRun { int g_lo, g_hi, cnt; };

// pass 1: this thread's run of the H * tiles_x segments of an image (seg = the image's tile_seg words) and its corner count
__device__ __forceinline__ CornerRun corner_run(const uint16_t* __restrict__ seg, int total_segs, int threads, int tid) {
    CornerRun r;
    const int spt = (total_segs + threads - 1) / threads;
    r.g_lo = min(total_segs, tid * spt);
    r.g_hi = min(total_segs, r.g_lo + spt);
    r.cnt = 0;
    for (int g = r.g_lo; g < r.g_hi; ++g) r.cnt += (int)(seg[g] & 0xFFu);
    return r;
}

// pass 2: stage[o ..] = one word per corner of the run: entry index inside the image's tile_list (tile << 9 | band << 6 | k)
// in the low 30 bits, the row inside the band in the top two.  o = the run's first position, n = capacity of stage.
__device__ __forceinline__ void corner_codes(const uint16_t* __restrict__ seg, const CornerRun& r, int tiles_x, int o, int n, uint32_t* stage) {
    static_assert(kTileCap == 512 && kBandCap == 64 && kBandH == 4, "entry index: tile << 9 | band << 6 | k");
    if (!r.cnt) return;
    for (int g = r.g_lo; g < r.g_hi; ++g) {
        const uint32_t sg = seg[g];
        const int c = (int)(sg & 0xFFu);
        if (!c) continue;
        const int y = g / tiles_x, tx = g - y * tiles_x;
        const uint32_t code = ((uint32_t)(y & (kBandH - 1)) << 30) |
                              (((uint32_t)(y / kTileH) * tiles_x + tx) << 9 | (uint32_t)((y & (kTileH - 1)) / kBandH) << 6) + (sg >> 8);
        for (int k = 0; k < c; ++k, ++o)
            if (o < n) stage[o] = code + k;
    }
}

struct Corner { int x, y, score; };

// pass 3: the corner behind one word of pass 2 (tlst = the image's tile_list)
__device__ __forceinline__ Corner corner_fetch(const uint16_t* __restrict__ tlst, uint32_t code, int tiles_x) {
    const uint32_t ent = code & 0x3FFFFFFFu;
    const uint32_t v = tlst[ent];
    const int tile = (int)(ent >> 9), ty = tile / tiles_x, tx = tile - ty * tiles_x;
    Corner c;
    c.y = ty * kTileH + (int)((ent >> 6) & 7u) * kBandH + (int)(code >> 30);
    c.x = tx * kTileW + (int)(v & 0xFFu);
    c.score = (int)(v >> 8);
    return c;
}

}  // namespace svo
