// orb_host.cpp -- host build of stereo_vo_b200/csrc/orb_core.cuh for tests/test_orb_cpu.py: the arithmetic the dmORB device
// kernels evaluate (level geometry, INTER_LINEAR_EXACT resize, retainBest = nth_element + partition, Harris response,
// IC angle / fastAtan2, steered BRIEF), driven against cv2.ORB_create(...).detectAndCompute on the CPU.
// TEST INFRASTRUCTURE: FAST corners and the blurred levels are handed in by the test (from the oracle / cv2).
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../stereo_vo_b200/csrc/orb_core.cuh"

using namespace svo;

static const int kPattern[256][4] = {
#include "../../stereo_vo_b200/csrc/orb_pattern.inc"
};

extern "C" {

void orb_host_levels(int W, int H, int nlevels, int nfeatures, float* scale, int32_t* w, int32_t* h, int32_t* nfeat) {
    const OrbLevels L = orb_levels(W, H, nlevels, nfeatures);
    for (int l = 0; l < nlevels; ++l) { scale[l] = L.scale[l]; w[l] = L.w[l]; h[l] = L.h[l]; nfeat[l] = L.nfeat[l]; }
}

// cv::resize(src, dst, Size(dw, dh), 0, 0, INTER_LINEAR_EXACT), 8-bit single channel, with the uniform tables of alloc_orb
void orb_host_resize(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride) {
    auto tables = [](int s, int d, std::vector<int>& ofs, std::vector<int>& c1) {
        const OrbAxis ax = orb_resize_axis(s, d);
        ofs.resize(d); c1.resize(d);
        for (int v = 0; v < d; ++v) {
            if (v < ax.lo) { ofs[v] = 0; c1[v] = 0; }
            else if (v >= ax.hi) { ofs[v] = s - 2; c1[v] = 256; }
            else { ofs[v] = ax.ofs[v]; c1[v] = ax.c1[v]; }
        }
    };
    std::vector<int> xo, xc, yo, yc;
    tables(sw, dw, xo, xc);
    tables(sh, dh, yo, yc);
    for (int y = 0; y < dh; ++y) {
        const uint8_t* r0 = src + (size_t)yo[y] * sstride;
        const uint8_t* r1 = r0 + sstride;
        for (int x = 0; x < dw; ++x) {
            const uint32_t h0 = r0[xo[x]] * (256 - xc[x]) + r0[xo[x] + 1] * xc[x];
            const uint32_t h1 = r1[xo[x]] * (256 - xc[x]) + r1[xo[x] + 1] * xc[x];
            dst[(size_t)y * dstride + x] = (uint8_t)((h0 * (256 - yc[y]) + h1 * yc[y] + (1u << 15)) >> 16);
        }
    }
}

// computeKeyPoints behind the per-level FAST detection: fast[l] = n_fast[l] x {x, y, score} (row-major order, already
// inside the 31-px border).  out: rows of {pt.x, pt.y, size, angle, response, octave}; returns the number of keypoints.
int orb_host_select(const uint8_t* const* level_img, const int32_t* level_stride, int W, int H, int nlevels, int nfeatures,
                    const int32_t* n_fast, const int32_t* const* fast, float* out, int cap) {
    const OrbLevels L = orb_levels(W, H, nlevels, nfeatures);
    int n_out = 0;
    for (int l = 0; l < nlevels; ++l) {
        const int n = n_fast[l];
        std::vector<uint32_t> idx(n);
        std::vector<float> resp(n);
        for (int i = 0; i < n; ++i) { idx[i] = (uint32_t)i; resp[i] = (float)fast[l][3 * i + 2]; }
        const int n2 = orb_retain_best(idx.data(), n, resp.data(), 2 * L.nfeat[l]);
        for (int k = 0; k < n2; ++k) {
            const int i = (int)idx[k];
            resp[i] = orb_harris(level_img[l], level_stride[l], fast[l][3 * i], fast[l][3 * i + 1]);
        }
        const int n3 = orb_retain_best(idx.data(), n2, resp.data(), L.nfeat[l]);
        for (int k = 0; k < n3 && n_out < cap; ++k, ++n_out) {
            const int i = (int)idx[k];
            const int x = fast[l][3 * i], y = fast[l][3 * i + 1];
            float* o = out + 6 * n_out;
            o[0] = (float)x * L.scale[l]; o[1] = (float)y * L.scale[l]; o[2] = (float)kOrbPatch * L.scale[l];
            o[3] = orb_ic_angle(level_img[l], level_stride[l], x, y);
            o[4] = resp[i]; o[5] = (float)l;
        }
    }
    return n_out;
}

// computeOrbDescriptors on the blurred levels
void orb_host_describe(const uint8_t* const* level_blur, const int32_t* level_stride, int W, int H, int nlevels,
                       const float* kps /* rows of 6 as above */, int n, uint8_t* desc) {
    const OrbLevels L = orb_levels(W, H, nlevels, 1);
    for (int j = 0; j < n; ++j) {
        const float* k = kps + 6 * j;
        const int oct = (int)k[5];
        const float inv = 1.f / L.scale[oct];
        const int cy = (int)nearbyintf(k[1] * inv), cx = (int)nearbyintf(k[0] * inv);
        float a, b;
        orb_cos_sin(k[3], a, b);
        const int step = level_stride[oct];
        const uint8_t* center = level_blur[oct] + (size_t)cy * step + cx;
        for (int i = 0; i < 32; ++i) {
            int val = 0;
            for (int t = 0; t < 8; ++t) {
                const int* p = kPattern[i * 8 + t];
                int x0, y0, x1, y1;
                orb_steer(p[0], p[1], a, b, x0, y0);
                orb_steer(p[2], p[3], a, b, x1, y1);
                val |= (center[y0 * step + x0] < center[y1 * step + x1] ? 1 : 0) << t;
            }
            desc[32 * j + i] = (uint8_t)val;
        }
    }
}

}  // extern "C"
