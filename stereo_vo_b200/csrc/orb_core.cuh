// orb_core.cuh -- the arithmetic of cv::ORB::detectAndCompute (OpenCV 4.13 features2d/src/orb.cpp) that the dmORB branch
// of the reference calls (stage2_detect.cpp:458-496), restated so that host and device evaluate the same text:
//   * the level geometry (scales, sizes, features per level) and the INTER_LINEAR_EXACT coefficient tables   (host)
//   * KeyPointsFilter::retainBest = libstdc++ std::nth_element + std::partition on the keypoint vector       (host + device)
//   * HarrisResponses (integer sums, float formula), ICAngles (integer moments) + cv::fastAtan2              (host + device)
//   * the steered BRIEF sample position (float rotation, cvRound)                                             (host + device)
// Pinned against cv2: tests/test_orb_cpu.py drives the host build (tests/cpp/orb_host.cpp) against
// cv2.ORB_create(...).detectAndCompute, every keypoint field and descriptor byte.
#pragma once
#include <math.h>
#include <stdint.h>
#include <vector>
#include "introsort_emul.cuh"

namespace svo {

constexpr int kOrbMaxLevels = 8;          // params_detect.orb_nlevels (reference default 8, libstereo-odometry.h:397)
constexpr int kOrbPatch = 31, kOrbHalfPatch = 15;
constexpr int kOrbHarrisBlock = 7;

// ---- level geometry (ORB_Impl::detectAndCompute / computeKeyPoints) -------------------------------------------
struct OrbLevels {
    int n;
    float scale[kOrbMaxLevels];           // getScale(level, 0, scaleFactor) = (float)pow(scaleFactor, level)
    int w[kOrbMaxLevels], h[kOrbMaxLevels];
    int nfeat[kOrbMaxLevels];             // nfeaturesPerLevel
};

inline OrbLevels orb_levels(int W, int H, int nlevels, int nfeatures) {
    OrbLevels L;
    L.n = nlevels;
    const double sf = (double)1.2f;       // ORB::create takes a float, ORB_Impl stores a double
    for (int l = 0; l < nlevels; ++l) {
        L.scale[l] = (float)pow(sf, (double)l);
        const float inv = 1.0f / L.scale[l];
        L.w[l] = (int)nearbyintf((float)W * inv);      // cvRound(image.cols * inv_scale)
        L.h[l] = (int)nearbyintf((float)H * inv);
    }
    const float factor = (float)(1.0 / sf);
    float nd = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        L.nfeat[l] = (int)nearbyintf(nd);
        sum += L.nfeat[l];
        nd *= factor;
    }
    L.nfeat[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
    return L;
}

// cv::resize(INTER_LINEAR_EXACT) for 8-bit, one axis: interpolationLinear<ufixedpoint16>::getCoeffs.
// ofs[v] = left tap (the right one is ofs + 1), c1[v] = weight of the right tap in 1/256 (c0 = 256 - c1);
// v < lo reads src[0], v >= hi reads src[src - 1].
struct OrbAxis { std::vector<int32_t> ofs; std::vector<int32_t> c1; int lo, hi; };
inline OrbAxis orb_resize_axis(int src, int dst) {
    OrbAxis A;
    A.ofs.assign(dst, 0); A.c1.assign(dst, 0); A.lo = 0; A.hi = dst;
    const double inv_scale = (double)dst / (double)src;
    const double scale = 1.0 / inv_scale;
    for (int v = 0; v < dst; ++v) {
        const double fval = scale * ((double)v + 0.5) - 0.5;
        const int ival = (int)floor(fval);
        if (ival >= 0 && src > 1) {
            if (ival < src - 1) {
                A.ofs[v] = ival;
                A.c1[v] = (int)nearbyint((fval - ival) * 256.0);
            } else {
                A.ofs[v] = src - 1;
                if (v < A.hi) A.hi = v;
            }
        } else if (v + 1 > A.lo) A.lo = v + 1;
    }
    return A;
}

// ---- KeyPointsFilter::retainBest --------------------------------------------------------------------------------
// std::__introselect (libstdc++): nth_element(a, a + nth, a + n, comp) on 32-bit elements
template <class C>
SVO_HD void introselect_seq(const C& c, uint32_t* a, int first, int last, int nth, int depth) {
    while (last - first > 3) {
        if (depth == 0) {
            // std::__heap_select(first, nth + 1, last) + iter_swap(first, nth)
            uint32_t* h = a + first;
            const int len = nth + 1 - first;
            if (len >= 2) {
                int parent = (len - 2) / 2;
                while (true) {
                    const uint32_t v = h[parent];
                    adjust_heap_(c, h, parent, len, v);
                    if (parent == 0) break;
                    --parent;
                }
            }
            for (int i = nth + 1; i < last; ++i)
                if (c.lt(a[i], h[0])) {
                    const uint32_t v = a[i];
                    a[i] = h[0];
                    adjust_heap_(c, h, 0, len, v);
                }
            const uint32_t t = a[first]; a[first] = a[nth]; a[nth] = t;
            return;
        }
        --depth;
        const int mid = first + (last - first) / 2;
        median_to_first_(c, a, first, first + 1, mid, last - 1);
        const int cut = unguarded_partition_(c, a, first + 1, last, a[first]);
        if (cut <= nth) first = cut; else last = cut;
    }
    insertion_sort_(c, a, first, last);
}
template <class C>
SVO_HD void nth_element_seq(const C& c, uint32_t* a, int n, int nth) {
    if (n == 0 || nth == n) return;
    introselect_seq(c, a, 0, n, nth, 2 * floor_log2(n));
}

// std::partition (bidirectional version) of a[first..last) by pred; returns the partition point
template <class Pred>
SVO_HD int partition_seq(const Pred& pred, uint32_t* a, int first, int last) {
    while (true) {
        while (true) {
            if (first == last) return first;
            else if (pred(a[first])) ++first;
            else break;
        }
        --last;
        while (true) {
            if (first == last) return first;
            else if (!pred(a[last])) --last;
            else break;
        }
        const uint32_t t = a[first]; a[first] = a[last]; a[last] = t;
        ++first;
    }
}

// retainBest on an index vector idx[0..n) with responses resp[idx]: returns the new size
struct OrbRespGreater {                 // KeypointResponseGreater
    const float* r;
    SVO_HD bool lt(uint32_t a, uint32_t b) const { return r[a] > r[b]; }
};
struct OrbRespGE {                      // KeypointResponseGreaterThanOrEqualToThreshold
    const float* r; float v;
    SVO_HD bool operator()(uint32_t a) const { return r[a] >= v; }
};
SVO_HD int orb_retain_best(uint32_t* idx, int n, const float* resp, int n_points) {
    if (n_points >= 0 && n > n_points) {
        if (n_points == 0) return 0;
        nth_element_seq(OrbRespGreater{resp}, idx, n, n_points - 1);
        const float ambiguous = resp[idx[n_points - 1]];
        return partition_seq(OrbRespGE{resp, ambiguous}, idx, n_points, n);
    }
    return n;
}

// ---- HarrisResponses (blockSize 7, k = 0.04) on the unblurred level image ----------------------------------------
// the float formula on the three integer sums: (a*b - c*c - k*(a+b)*(a+b)) * scale^4, every operation rounded to float
SVO_HD float orb_harris_from_sums(int a, int b, int c) {
    const float scale = 1.f / ((1 << 2) * kOrbHarrisBlock * 255.f);
    const float s2 = scale * scale, s3 = s2 * scale, s4 = s3 * scale;        // scale * scale * scale * scale, left to right
    const float fa = (float)a, fb = (float)b, fc = (float)c;
#if defined(__CUDA_ARCH__)
    const float ab = __fmul_rn(fa, fb), cc = __fmul_rn(fc, fc), s = __fadd_rn(fa, fb);
    const float t = __fmul_rn(__fmul_rn(0.04f, s), s);
    return __fmul_rn(__fsub_rn(__fsub_rn(ab, cc), t), s4);
#else
    const float ab = fa * fb, cc = fc * fc, s = fa + fb;
    const float t = (0.04f * s) * s;
    return ((ab - cc) - t) * s4;
#endif
}
SVO_HD float orb_harris(const uint8_t* img, int step, int x0, int y0) {
    const int r = kOrbHarrisBlock / 2;
    int a = 0, b = 0, c = 0;
    for (int i = 0; i < kOrbHarrisBlock; ++i)
        for (int j = 0; j < kOrbHarrisBlock; ++j) {
            const uint8_t* p = img + (y0 - r + i) * step + (x0 - r + j);
            const int Ix = (p[1] - p[-1]) * 2 + (p[-step + 1] - p[-step - 1]) + (p[step + 1] - p[step - 1]);
            const int Iy = (p[step] - p[-step]) * 2 + (p[step - 1] - p[-step - 1]) + (p[step + 1] - p[-step + 1]);
            a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
        }
    return orb_harris_from_sums(a, b, c);
}

// ---- cv::fastAtan2 (degrees) ---------------------------------------------------------------------------------------
SVO_HD float orb_fast_atan2(float y, float x) {
    const float p1 = 0.9997878412794807f * (float)(180 / 3.14159265358979323846);
    const float p3 = -0.3258083974640975f * (float)(180 / 3.14159265358979323846);
    const float p5 = 0.1555786518463281f * (float)(180 / 3.14159265358979323846);
    const float p7 = -0.04432655554792128f * (float)(180 / 3.14159265358979323846);
    const float ax = fabsf(x), ay = fabsf(y);
    const float eps = (float)2.2204460492503131e-16;
    float a;
#if defined(__CUDA_ARCH__)
    if (ax >= ay) {
        const float c = __fdiv_rn(ay, __fadd_rn(ax, eps)), c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        const float c = __fdiv_rn(ax, __fadd_rn(ay, eps)), c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
#else
    if (ax >= ay) {
        const float c = ay / (ax + eps), c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        const float c = ax / (ay + eps), c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
#endif
    return a;
}

// ---- ICAngles: intensity centroid over the circular patch of radius 15 ----------------------------------------------
SVO_HD float orb_ic_angle(const uint8_t* img, int step, int x0, int y0) {
    const int umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
    const uint8_t* center = img + y0 * step + x0;
    int m_01 = 0, m_10 = 0;
    for (int u = -kOrbHalfPatch; u <= kOrbHalfPatch; ++u) m_10 += u * center[u];
    for (int v = 1; v <= kOrbHalfPatch; ++v) {
        int v_sum = 0;
        const int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            const int val_plus = center[u + v * step], val_minus = center[u - v * step];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return orb_fast_atan2((float)m_01, (float)m_10);
}

// ---- steered BRIEF: sample offset of pattern point (px, py) under rotation (a = cos, b = sin), computeOrbDescriptors ---
SVO_HD void orb_steer(int px, int py, float a, float b, int& ix, int& iy) {
#if defined(__CUDA_ARCH__)
    const float x = __fsub_rn(__fmul_rn((float)px, a), __fmul_rn((float)py, b));
    const float y = __fadd_rn(__fmul_rn((float)px, b), __fmul_rn((float)py, a));
    ix = __float2int_rn(x); iy = __float2int_rn(y);
#else
    const float x = (float)px * a - (float)py * b;
    const float y = (float)px * b + (float)py * a;
    ix = (int)nearbyintf(x); iy = (int)nearbyintf(y);
#endif
}

// angle in degrees -> (cos, sin) as computeOrbDescriptors takes them: angle *= (float)(CV_PI/180.f); (float)cos(angle)
SVO_HD void orb_cos_sin(float angle_deg, float& a, float& b) {
    const float k = (float)(3.14159265358979323846 / 180.f);
#if defined(__CUDA_ARCH__)
    const float ang = __fmul_rn(angle_deg, k);
#else
    const float ang = angle_deg * k;
#endif
    a = (float)cos((double)ang);
    b = (float)sin((double)ang);
}

}  // namespace svo
