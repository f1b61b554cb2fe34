// Host harness (test only): sequential findFundamentalMat(FM_RANSAC, 1.0, 0.99) built from
// stereo_vo_b200/csrc/fundmat_core.cuh, used to pin that header against cv2 on the CPU.
// The CUDA kernel (ransac.cu) evaluates the same iterations in parallel batches.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../stereo_vo_b200/csrc/fundmat_core.cuh"
using namespace svo;

extern "C" int fundmat_host(const float* p1, const float* p2, int n, uint8_t* mask, double* bestF) {
    const Pt2f* m1 = reinterpret_cast<const Pt2f*>(p1);
    const Pt2f* m2 = reinterpret_cast<const Pt2f*>(p2);
    if (n < 7) return -1;                       // findFundamentalMat returns an empty Mat, no mask
    if (n == 7) { std::memset(mask, 1, 7); return 7; }
    CvRng rng;
    rng.seed();
    int idx[7];
    Pt2f ms1[7], ms2[7];
    double F[27], best[9] = {0};
    if (n >= 15) {   // RANSACPointSetRegistrator::run
        int niters = kFmMaxIters, maxGood = 0;
        const float thr = (float)(kFmThreshold * kFmThreshold);
        for (int iter = 0; iter < niters; ++iter) {
            if (!get_subset7(m1, m2, n, rng, 10000, idx, ms1, ms2)) { if (iter == 0) return 0; break; }
            const int nm = run_7point(ms1, ms2, F);
            for (int i = 0; i < nm; ++i) {
                int good = 0;
                for (int k = 0; k < n; ++k) good += fm_error(F + 9 * i, m1[k], m2[k]) <= thr;
                if (good > std::max(maxGood, kFmModelPoints - 1)) {
                    maxGood = good;
                    std::memcpy(best, F + 9 * i, sizeof(best));
                    niters = ransac_update_num_iters(kFmConfidence, (double)(n - good) / n, kFmModelPoints, niters);
                }
            }
        }
        std::memset(mask, 0, n);
        if (maxGood > 0) for (int k = 0; k < n; ++k) mask[k] = fm_error(best, m1[k], m2[k]) <= thr;
        if (bestF) std::memcpy(bestF, best, sizeof(best));
        return maxGood;
    }
    // LMeDSPointSetRegistrator::run (8 <= n < 15)
    int niters = ransac_update_num_iters(kFmConfidence, 0.45, kFmModelPoints, kFmMaxIters);
    niters = std::max(niters, 3);
    double minMedian = DBL_MAX;
    std::vector<float> err(n);
    for (int iter = 0; iter < niters; ++iter) {
        if (!get_subset7(m1, m2, n, rng, 1000, idx, ms1, ms2)) { if (iter == 0) return 0; break; }
        const int nm = run_7point(ms1, ms2, F);
        for (int i = 0; i < nm; ++i) {
            for (int k = 0; k < n; ++k) err[k] = fm_error(F + 9 * i, m1[k], m2[k]);
            std::nth_element(err.begin(), err.begin() + n / 2, err.end());
            const double median = err[n / 2];
            if (median < minMedian) { minMedian = median; std::memcpy(best, F + 9 * i, sizeof(best)); }
        }
    }
    std::memset(mask, 0, n);
    if (!(minMedian < DBL_MAX)) return 0;
    double sigma = 2.5 * 1.4826 * (1 + 5. / (n - kFmModelPoints)) * std::sqrt(minMedian);
    sigma = std::max(sigma, 0.001);
    const float thr = (float)(sigma * sigma);
    int good = 0;
    for (int k = 0; k < n; ++k) { mask[k] = fm_error(best, m1[k], m2[k]) <= thr; good += mask[k]; }
    if (bestF) std::memcpy(bestF, best, sizeof(best));
    return good;
}

// run7Point alone (cv2.findFundamentalMat(p1[:7], p2[:7], FM_7POINT) returns its stacked 3x3 models)
extern "C" int fundmat_host_7point(const float* p1, const float* p2, double* F27) {
    return run_7point(reinterpret_cast<const Pt2f*>(p1), reinterpret_cast<const Pt2f*>(p2), F27);
}

// the SVD step alone: At = 9 x 9 row-major, rows 0..6 = A; cv2.SVDecomp(A, flags=FULL_UV)[2] is the result
extern "C" void fundmat_host_svd_rows(double* At) { fm7_svd_rows(reinterpret_cast<double (*)[9]>(At)); }

// the correctly rounded libm replacements the device build of solve_cubic uses (csrc/cr_math.cuh)
// (cr_math.cuh comes in through fundmat_core.cuh)
extern "C" double cr_pow13(double x) { return svo::cr::pow13(x); }
extern "C" double cr_cos(double x) { return svo::cr::cos_cr(x); }
extern "C" double cr_acos(double x) { return svo::cr::acos_cr(x); }
extern "C" int fundmat_host_cubic(const double* c, double* r) { return solve_cubic(c, r); }
