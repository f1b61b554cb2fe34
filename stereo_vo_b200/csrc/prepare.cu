// prepare.cu -- stage 1 of the pipeline on sm_100a (stage1_rectify.cpp:37-93 with nOctaves = 1):
//   * grayscale conversion: mrpt::utils::CImage(img, FAST_REF_OR_CONVERT_TO_GRAY) = cv::cvtColor(BGR2GRAY)
//     (stage1_rectify.cpp:51-52); 8-bit fixed point  (B*3735 + G*19235 + R*9798 + 2^14) >> 15
//   * stereo rectification: mrpt::vision::CStereoRectifyMap::rectify (stage1_rectify.cpp:68-74) =
//     cv::remap(src, dst, map_xy (CV_16SC2), map_a (CV_16UC1), INTER_LINEAR, BORDER_CONSTANT, 0) with the maps
//     cv::initUndistortRectifyMap produced once from the camera parameters (setFromCamParams, :69-70).
//     Fixed point, INTER_BITS = 5: a = map_a & 1023, fx = a & 31, fy = a >> 5,
//       dst = (S00*(32-fx)*(32-fy)*32 + S01*fx*(32-fy)*32 + S10*(32-fx)*fy*32 + S11*fx*fy*32 + 2^14) >> 15,
//     taps outside the source image read the border value 0.
// Both are pinned bit-exactly against cv2 (tests/test_oracle_pins.py, tests/golden/golden_stage1.npz).
// The reference hands the ORIGINAL image to the rectifier (stage1_rectify.cpp:72), which only works for
// gray input; here colour input is converted first and the gray image is rectified.
//
// Roofline: pure streaming / gather, HBM bound.  Algorithmic bytes per output pixel: CH (source) + 1 (output);
// the 6 B/px of maps are shared by the whole batch and stay in L2.
#include "prepare.cuh"

namespace svo {

__device__ __forceinline__ uint32_t gray_bgr(uint32_t b, uint32_t g, uint32_t r) {
    return (b * 3735u + g * 19235u + r * 9798u + 16384u) >> 15;
}

template <int CH>
__device__ __forceinline__ uint32_t tap(const uint8_t* __restrict__ src, size_t stride, int W, int H, int x, int y) {
    if ((unsigned)x >= (unsigned)W || (unsigned)y >= (unsigned)H) return 0u;      // BORDER_CONSTANT, value 0
    const uint8_t* p = src + (size_t)y * stride + (size_t)x * CH;
    if (CH == 1) return __ldg(p);
    return gray_bgr(__ldg(p), __ldg(p + 1), __ldg(p + 2));
}

// ---- rectification (optionally fused with the gray conversion of the taps) ----------------------
// one thread = 4 consecutive output pixels (one 32-bit store); blockIdx.z = 2 * item + side.
// generic pixel: per-tap border test; kept out of line so that the common path below stays lean
template <int CH>
__device__ __noinline__ uint32_t remap_px(const uint8_t* __restrict__ src, int stride, int W, int H, int sx, int sy, uint32_t a) {
    const uint32_t fx = a & 31u, fy = (a >> 5) & 31u;
    const uint32_t s00 = tap<CH>(src, (size_t)stride, W, H, sx, sy);
    const uint32_t s01 = tap<CH>(src, (size_t)stride, W, H, sx + 1, sy);
    const uint32_t s10 = tap<CH>(src, (size_t)stride, W, H, sx, sy + 1);
    const uint32_t s11 = tap<CH>(src, (size_t)stride, W, H, sx + 1, sy + 1);
    // weights (32-fx)(32-fy)*32 ... sum to 2^15; (32*X + 2^14) >> 15 == (X + 512) >> 10
    const uint32_t gx = 32u - fx, gy = 32u - fy;
    return (s00 * (gx * gy) + s01 * (fx * gy) + s10 * (gx * fy) + s11 * (fx * fy) + 512u) >> 10;
}

constexpr int RK_BX = 32, RK_BY = 8, RK_ROWS = 4;      // CTA = 128 x 32 output pixels, 4 rows per thread

template <int CH>
__global__ void __launch_bounds__(RK_BX * RK_BY)
rectify_kernel(PrepareArgs A) {
    const int x4 = (blockIdx.x * RK_BX + threadIdx.x) * 4;
    if (x4 >= A.pitch_out) return;
    const int item = blockIdx.z >> 1, side = blockIdx.z & 1;
    const uint8_t* __restrict__ src = (side ? A.srcR : A.srcL) + (size_t)item * A.src_image_pitch;
    uint8_t* __restrict__ dst = (side ? A.dstR : A.dstL) + (size_t)item * A.dst_image_pitch;
    const int stride = (int)A.src_stride;
    const bool words_ok = CH == 1 && ((stride | (int)reinterpret_cast<uintptr_t>(src)) & 3) == 0;
    const bool maps4 = x4 + 3 < A.W && (A.W & 3) == 0;      // maps of 4 pixels: 16 + 8 bytes, aligned
#pragma unroll
    for (int rr = 0; rr < RK_ROWS; ++rr) {
        const int y = (blockIdx.y * RK_ROWS + rr) * RK_BY + threadIdx.y;
        if (y >= A.H) break;
        const short2* __restrict__ mxy = A.map_xy[side] + (size_t)y * A.W + x4;
        const uint16_t* __restrict__ ma = A.map_a[side] + (size_t)y * A.W + x4;
        uint32_t out = 0;
        if (maps4) {
            const uint4 m = __ldg(reinterpret_cast<const uint4*>(mxy));
            const uint2 av = __ldg(reinterpret_cast<const uint2*>(ma));
            const int sx[4] = {(short)(m.x & 0xFFFF), (short)(m.y & 0xFFFF), (short)(m.z & 0xFFFF), (short)(m.w & 0xFFFF)};
            const int sy[4] = {(int)m.x >> 16, (int)m.y >> 16, (int)m.z >> 16, (int)m.w >> 16};
            const uint32_t as[4] = {av.x, av.x >> 16, av.y, av.y >> 16};
            // Common case (gray source, smooth map): the four 2x2 footprints lie inside the image, on at most three
            // consecutive source rows and within the 8 bytes of an aligned word pair -> 3 x 2 word loads instead of
            // 16 byte gathers; the two taps of a row are picked with PRMT and blended with IDP.2A
            const int base = sx[0] & ~3;
            const int ylo = min(min(sy[0], sy[1]), min(sy[2], sy[3])), yhi = max(max(sy[0], sy[1]), max(sy[2], sy[3]));
            const unsigned o0 = (unsigned)(sx[0] - base), o1 = (unsigned)(sx[1] - base), o2 = (unsigned)(sx[2] - base),
                           o3 = (unsigned)(sx[3] - base);
            if (words_ok && yhi - ylo <= 1 && ylo >= 0 && ylo + 2 < A.H && max(max(o0, o1), max(o2, o3)) <= 6u && base >= 0 &&
                base + 8 <= A.W) {
                const uint32_t* r0 = reinterpret_cast<const uint32_t*>(src + (ylo * stride + base));
                const uint32_t* r1 = reinterpret_cast<const uint32_t*>(src + ((ylo + 1) * stride + base));
                const uint32_t* r2 = reinterpret_cast<const uint32_t*>(src + ((ylo + 2) * stride + base));
                const uint32_t t0 = __ldg(r0), t1 = __ldg(r0 + 1), m0 = __ldg(r1), m1 = __ldg(r1 + 1), b0 = __ldg(r2), b1 = __ldg(r2 + 1);
                const unsigned os[4] = {o0, o1, o2, o3};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t sel = os[q] * 0x11u + 0x4410u;            // bytes o, o+1 of a word pair -> bytes 0, 1
                    const bool lower = sy[q] != ylo;                         // footprint on rows ylo+1, ylo+2
                    const uint32_t tp = __byte_perm(lower ? m0 : t0, lower ? m1 : t1, sel);
                    const uint32_t bp = __byte_perm(lower ? b0 : m0, lower ? b1 : m1, sel);
                    const uint32_t fx = as[q] & 31u, fy = (as[q] >> 5) & 31u;
                    const uint32_t wx = 32u + fx * 0xFFFFu;                  // {32 - fx, fx} as 16-bit lanes
                    // (32*X + 2^14) >> 15 == (X + 512) >> 10 with X = sum of tap * wx * wy
                    const uint32_t acc = __dp2a_lo(fy * wx, bp, __dp2a_lo((32u - fy) * wx, tp, 512u));
                    out |= (acc >> 10) << (8 * q);
                }
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) out |= remap_px<CH>(src, stride, A.W, A.H, sx[q], sy[q], as[q]) << (8 * q);
            }
        } else {
            for (int q = 0; q < 4; ++q) {
                if (x4 + q >= A.W) break;
                const short2 sxy = mxy[q];
                out |= remap_px<CH>(src, stride, A.W, A.H, sxy.x, sxy.y, ma[q]) << (8 * q);
            }
        }
        *reinterpret_cast<uint32_t*>(dst + (size_t)y * A.pitch_out + x4) = out;
    }
}

// ---- gray conversion / plain copy into the pitched plane (images already rectified) ------------------
// one thread = 16 consecutive output pixels: 3 x 16-byte loads (BGR) or one (gray), one 16-byte store
template <int CH>
__global__ void __launch_bounds__(256)
gray_kernel(PrepareArgs A) {
    const int x16 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x16 >= A.pitch_out || y >= A.H) return;
    const int item = blockIdx.z >> 1, side = blockIdx.z & 1;
    const uint8_t* __restrict__ src = (side ? A.srcR : A.srcL) + (size_t)item * A.src_image_pitch;
    uint8_t* __restrict__ dst = (side ? A.dstR : A.dstL) + (size_t)item * A.dst_image_pitch;
    const uint8_t* __restrict__ row = src + (size_t)y * A.src_stride + (size_t)x16 * CH;
    uint32_t o[4] = {0, 0, 0, 0};
    if (x16 + 15 < A.W && (reinterpret_cast<uintptr_t>(row) & 15) == 0) {
        if (CH == 1) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(row));
            o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
        } else {
            const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(row)), v1 = __ldg(reinterpret_cast<const uint4*>(row) + 1),
                        v2 = __ldg(reinterpret_cast<const uint4*>(row) + 2);
            const uint32_t w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
            for (int g = 0; g < 4; ++g) {            // 4 pixels = 3 words: B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
                const uint32_t w0 = w[3 * g], w1 = w[3 * g + 1], w2 = w[3 * g + 2];
                o[g] = gray_bgr(w0 & 0xFF, (w0 >> 8) & 0xFF, (w0 >> 16) & 0xFF) |
                       (gray_bgr(w0 >> 24, w1 & 0xFF, (w1 >> 8) & 0xFF) << 8) |
                       (gray_bgr((w1 >> 16) & 0xFF, w1 >> 24, w2 & 0xFF) << 16) |
                       (gray_bgr((w2 >> 8) & 0xFF, (w2 >> 16) & 0xFF, w2 >> 24) << 24);
            }
        }
    } else {
#pragma unroll
        for (int q = 0; q < 16; ++q)
            if (x16 + q < A.W) o[q >> 2] |= tap<CH>(src, A.src_stride, A.W, A.H, x16 + q, y) << (8 * (q & 3));
    }
    *reinterpret_cast<uint4*>(dst + (size_t)y * A.pitch_out + x16) = make_uint4(o[0], o[1], o[2], o[3]);
}

cudaError_t launch_prepare(const PrepareArgs& A, int n, int channels, bool rectify, cudaStream_t st, int* launches) {
    if (n <= 0) return cudaSuccess;
    const dim3 rblock(RK_BX, RK_BY), gblock(32, 8);
    const dim3 rgrid((A.pitch_out / 4 + RK_BX - 1) / RK_BX, (A.H + RK_BY * RK_ROWS - 1) / (RK_BY * RK_ROWS), 2 * n);
    const dim3 ggrid((A.pitch_out / 16 + gblock.x - 1) / gblock.x, (A.H + gblock.y - 1) / gblock.y, 2 * n);
    if (!rectify) {
        if (channels == 1) gray_kernel<1><<<ggrid, gblock, 0, st>>>(A);
        else gray_kernel<3><<<ggrid, gblock, 0, st>>>(A);
        if (launches) *launches += 1;
    } else if (channels == 1) {
        rectify_kernel<1><<<rgrid, rblock, 0, st>>>(A);
        if (launches) *launches += 1;
    } else if (A.tmpL && A.tmpR) {
        // colour + rectification: convert into the scratch planes with wide loads, then rectify the gray planes with
        // the word-load path (cheaper than 12 byte gathers per pixel in one fused pass)
        PrepareArgs G = A;
        G.dstL = A.tmpL; G.dstR = A.tmpR;
        gray_kernel<3><<<ggrid, gblock, 0, st>>>(G);
        PrepareArgs R = A;
        R.srcL = A.tmpL; R.srcR = A.tmpR; R.src_stride = (size_t)A.pitch_out; R.src_image_pitch = A.dst_image_pitch;
        rectify_kernel<1><<<rgrid, rblock, 0, st>>>(R);
        if (launches) *launches += 2;
    } else {
        rectify_kernel<3><<<rgrid, rblock, 0, st>>>(A);
        if (launches) *launches += 1;
    }
    return cudaGetLastError();
}

// ---- image pyramid (params_rectify.nOctaves > 1) -------------------------------------------------
// CImagePyramid::buildPyramidFast(img, nOctaves) halves with CImage::scaleHalfSmooth (smooth_halves defaults to true):
//   * MRPT's SSE2 routine when the source is 8-bit gray with width % 16 == 0: rows y and y+1 averaged with
//     _mm_avg_epu8, then horizontal neighbours with _mm_avg_epu16, both rounding half up  (restated from MRPT 1.x:
//     its source is not in this image, UNPINNED);
//   * cvPyrDown otherwise: separable [1 4 6 4 1], (sum + 128) >> 8, BORDER_REFLECT_101 (pinned against cv2.pyrDown).
// One thread = 8 output pixels (two 16-byte loads, one 8-byte store); blockIdx.z = 2 * item + side.  Streaming and
// HBM-bound: algorithmic bytes = sw*sh read + a quarter of it written per image; < 2 % of a step.
__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

template <int METHOD>
__global__ void __launch_bounds__(256)
pyramid_kernel(PyramidArgs A) {
    const int x8 = (blockIdx.x * 8 + threadIdx.x) * 8;       // CTA = 64 x 32 output pixels: a warp covers 4 rows of 64
    const int y = blockIdx.y * 32 + threadIdx.y;
    const int dw = A.sw >> 1, dh = A.sh >> 1;
    if (x8 >= (int)A.dst_stride || y >= dh) return;
    const int item = blockIdx.z >> 1, side = blockIdx.z & 1;
    const uint8_t* __restrict__ src = (side ? A.srcR : A.srcL) + (size_t)item * A.src_image_pitch;
    uint8_t* __restrict__ dst = (side ? A.dstR : A.dstL) + (size_t)item * A.dst_image_pitch;
    uint32_t out[2] = {0u, 0u};
    if (METHOD == PYR_AVG2X2) {
        const uint8_t* r0 = src + (size_t)(2 * y) * A.src_stride + 2 * x8;
        const uint8_t* r1 = r0 + A.src_stride;
        if (x8 + 7 < dw && ((reinterpret_cast<uintptr_t>(r0) | A.src_stride) & 15) == 0) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(r0)), b = __ldg(reinterpret_cast<const uint4*>(r1));
            // vertical, (p + q + 1) >> 1 per byte; then even / odd columns apart and the horizontal average
            const uint32_t v0 = __vavgu4(a.x, b.x), v1 = __vavgu4(a.y, b.y), v2 = __vavgu4(a.z, b.z), v3 = __vavgu4(a.w, b.w);
            out[0] = __vavgu4(__byte_perm(v0, v1, 0x6420), __byte_perm(v0, v1, 0x7531));
            out[1] = __vavgu4(__byte_perm(v2, v3, 0x6420), __byte_perm(v2, v3, 0x7531));
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (x8 + j >= dw) break;
                const uint32_t va = (r0[2 * j] + r1[2 * j] + 1u) >> 1, vb = (r0[2 * j + 1] + r1[2 * j + 1] + 1u) >> 1;
                out[j >> 2] |= ((va + vb + 1u) >> 1) << (8 * (j & 3));
            }
        }
    } else {
        int ry[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) ry[j] = reflect101(2 * y + j - 2, A.sh);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int x = x8 + j;
            if (x >= dw) break;
            int cx[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) cx[i] = reflect101(2 * x + i - 2, A.sw);
            uint32_t acc = 0;
#pragma unroll
            for (int r = 0; r < 5; ++r) {
                const uint8_t* row = src + (size_t)ry[r] * A.src_stride;
                const uint32_t h = row[cx[0]] + 4u * row[cx[1]] + 6u * row[cx[2]] + 4u * row[cx[3]] + row[cx[4]];
                acc += h * (r == 0 || r == 4 ? 1u : (r == 2 ? 6u : 4u));
            }
            out[j >> 2] |= ((acc + 128u) >> 8) << (8 * (j & 3));
        }
    }
    *reinterpret_cast<uint2*>(dst + (size_t)y * A.dst_stride + x8) = make_uint2(out[0], out[1]);   // padding columns: 0
}

cudaError_t launch_pyramid_level(const PyramidArgs& A, int n, int method, cudaStream_t st, int* launches) {
    if (n <= 0) return cudaSuccess;
    const dim3 block(8, 32);
    const dim3 grid(((int)A.dst_stride / 8 + 7) / 8, ((A.sh >> 1) + 31) / 32, 2 * n);
    if (method == PYR_AVG2X2) pyramid_kernel<PYR_AVG2X2><<<grid, block, 0, st>>>(A);
    else pyramid_kernel<PYR_PYRDOWN><<<grid, block, 0, st>>>(A);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace svo
