// fundmat_core.cuh -- building blocks of cv::findFundamentalMat(p1, p2, FM_RANSAC, 1.0, 0.99, mask)
// as called at stage4_match_consecutive.cpp:202,237, restated from OpenCV's published algorithm
// (calib3d: RANSACPointSetRegistrator / LMeDSPointSetRegistrator / FMEstimatorCallback):
//   * cv::RNG (multiply-with-carry, state 0xFFFFFFFFFFFFFFFF) and RNG::uniform(0, count)
//   * getSubset: 7 distinct indices, re-drawn while the 7th point is collinear with two others
//   * 7-point algorithm: null space of the 7x9 epipolar system on normalised points, cubic in lambda
//   * symmetric epipolar distance max(d1^2 s1, d2^2 s2) in double, compared as float32
//   * RANSACUpdateNumIters
// The integer parts (RNG, subsets, iteration counts) are exact.  The null space of the 7x9 system is OpenCV's
// own one-sided Jacobi SVD (core/lapack.cpp JacobiSVDImpl_<double>, which cv::SVDecomp runs for every matrix with
// fewer than 25 rows even in a LAPACK build), operation for operation: sequential sums, OpenCV's scaled hypot, and
// the two extra right singular vectors built from the fixed RNG(0x12345678) sign vectors by two Gram-Schmidt
// passes.  Checked bit for bit against cv2.SVDecomp and cv2.findFundamentalMat(FM_7POINT) (tests/test_host_cpu.py).
// Host + device; the host build is used by tests/cpp/fundmat_host.cpp to pin this file against cv2.
#pragma once
#include <float.h>
#include <math.h>
#include <stdint.h>
#include "cr_math.cuh"

#if defined(__CUDACC__)
#define SVO_HD2 __host__ __device__ inline
#define SVO_HD2_CALL __host__ __device__ __noinline__ inline   // kept out of line on the device (instruction-cache footprint)
#else
#define SVO_HD2 inline
#define SVO_HD2_CALL inline
#endif

namespace svo {

struct CvRng {
    uint64_t state;
    SVO_HD2 void seed() { state = 0xFFFFFFFFFFFFFFFFull; }   // RNG rng((uint64)-1)
    SVO_HD2 unsigned next() {
        state = (uint64_t)(unsigned)state * 4164903690ull + (unsigned)(state >> 32);
        return (unsigned)state;
    }
    SVO_HD2 int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

struct Pt2f { float x, y; };

// haveCollinearPoints(ms, count): only the last point is tested against all pairs of earlier ones
SVO_HD2_CALL bool last_point_collinear(const Pt2f* p, int count) {
    const int i = count - 1;
    for (int j = 0; j < i; j++) {
        const double dx1 = p[j].x - p[i].x, dy1 = p[j].y - p[i].y;
        for (int k = 0; k < j; k++) {
            const double dx2 = p[k].x - p[i].x, dy2 = p[k].y - p[i].y;
            if (fabs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2))) return true;
        }
    }
    return false;
}

// getSubset(m1, m2, ms1, ms2, rng, maxAttempts): returns false when no valid subset was found
SVO_HD2_CALL bool get_subset7(const Pt2f* m1, const Pt2f* m2, int count, CvRng& rng, int max_attempts, int* idx,
                         Pt2f* ms1, Pt2f* ms2) {
    for (int iters = 0; iters < max_attempts; ++iters) {
        for (int i = 0; i < 7; ++i) {
            int idx_i;
            while (true) {
                idx_i = rng.uniform(0, count);
                bool dup = false;
                for (int j = 0; j < i; ++j) dup = dup || (idx[j] == idx_i);
                if (!dup) break;
            }
            idx[i] = idx_i;
            ms1[i] = m1[idx_i];
            ms2[i] = m2[idx_i];
        }
        if (!last_point_collinear(ms1, 7) && !last_point_collinear(ms2, 7)) return true;
    }
    return false;
}

SVO_HD2 int cv_round(double v) { return (int)nearbyint(v); }   // cvRound: round half to even

// RANSACUpdateNumIters(p, ep, modelPoints, maxIters)
SVO_HD2_CALL int ransac_update_num_iters(double p, double ep, int model_points, int max_iters) {
    p = fmax(p, 0.); p = fmin(p, 1.);
    ep = fmax(ep, 0.); ep = fmin(ep, 1.);
    double num = fmax(1. - p, DBL_MIN);
    double denom = 1. - pow(1. - ep, (double)model_points);
    if (denom < DBL_MIN) return 0;
    num = log(num);
    denom = log(denom);
    return (denom >= 0 || -num >= max_iters * (-denom)) ? max_iters : cv_round(num / denom);
}

// cv::solveCubic (core/mathfuncs.cpp, OpenCV 4.13) for c[0] x^3 + c[1] x^2 + c[2] x + c[3] = 0; returns the number of
// real roots.  4.13 rescales the coefficients by their largest magnitude before the monic normalisation, decides the
// degree against DBL_EPSILON * that magnitude, and evaluates the discriminant in its expanded, regrouped form; the
// operation order below is the one of the shipped binary (checked bit for bit against cv2.solveCubic).  libm calls:
// SVO_POW13 / SVO_ACOS / SVO_COS (glibc on the host; correctly rounded double-double versions on the device).
#if defined(__CUDA_ARCH__) || defined(SVO_FORCE_CR_LIBM)
#define SVO_POW13(x) cr::pow13(x)
#define SVO_ACOS(x) cr::acos_cr(x)
#define SVO_COS(x) cr::cos_cr(x)
#else
#define SVO_POW13(x) pow((x), 1. / 3)
#define SVO_ACOS(x) acos(x)
#define SVO_COS(x) cos(x)
#endif
SVO_HD2_CALL int solve_cubic(const double* c, double* r) {
    double a0 = c[0], a1 = c[1], a2 = c[2], a3 = c[3];
    double x0 = 0., x1 = 0., x2 = 0.;
    int n = 0;
    const double scale = fmax(fabs(a3), fmax(fabs(a2), fmax(fabs(a1), fabs(a0))));
    if (DBL_EPSILON > scale) { r[0] = r[1] = r[2] = 0; return -1; }
    const double thr = DBL_EPSILON * scale;
    if (thr > fabs(a0)) {
        if (thr > fabs(a1)) {
            if (thr > fabs(a2)) n = thr > fabs(a3) ? -1 : 0;
            else { x0 = -a3 / a2; n = 1; }
        } else {
            double d = a2 * a2 - 4 * a1 * a3;
            if (d >= 0) {
                d = sqrt(d);
                const double q1 = (d - a2) * 0.5, q2 = (a2 + d) * -0.5;
                if (fabs(q1) > fabs(q2)) { x0 = q1 / a1; x1 = a3 / q1; }
                else { x0 = q2 / a1; x1 = a3 / q2; }
                n = d > 0 ? 2 : 1;
            }
        }
    } else {
        const double inv = 1. / scale;
        a0 *= inv; a1 *= inv; a2 *= inv; a3 *= inv;
        a0 = 1. / a0; a1 *= a0; a2 *= a0; a3 *= a0;
        const double Q = (a1 * a1 - 3 * a2) * (1. / 9);
        const double R = ((a1 + a1) * a1 - 9 * a2) * a1;
        const double R2 = (R + 27 * a3) * (1. / 54);
        double d = ((a2 * a2 - 4 * a1 * a3) * (a1 * a1) + (9 * a1 * a3 - a2 * (a2 + a2)) * (a2 + a2) - 27 * a3 * a3) * (1. / 108);
        const double kPi = 3.14159265358979323846;
        if (d > 0) {
            const double Qcubed = Q * Q * Q;
            const double theta = SVO_ACOS(R2 / sqrt(Qcubed));
            const double sqrtQ = sqrt(Q);
            const double t0 = -2 * sqrtQ, t1 = theta * (1. / 3), t2 = a1 * (1. / 3);
            x0 = t0 * SVO_COS(t1) - t2;
            x1 = t0 * SVO_COS(t1 + (2. * kPi / 3)) - t2;
            x2 = t0 * SVO_COS(t1 + (4. * kPi / 3)) - t2;
            n = 3;
        } else if (d == 0) {
            if (R2 >= 0) { x0 = -2 * SVO_POW13(R2) - a1 / 3; x1 = SVO_POW13(R2) - a1 / 3; }
            else { x0 = 2 * SVO_POW13(-R2) - a1 / 3; x1 = -SVO_POW13(-R2) - a1 / 3; }
            x2 = 0;
            n = x0 == x1 ? 1 : 2;
            x1 = x0 == x1 ? 0 : x1;
        } else {
            d = sqrt(-d);
            double e = SVO_POW13(d + fabs(R2));
            if (R2 > 0) e = -e;
            x0 = (e + Q / e) - a1 * (1. / 3);
            n = 1;
        }
    }
    r[0] = x0; r[1] = x1; r[2] = x2;
    return n;
}

// run7Point: up to three candidate fundamental matrices (row-major 3x3 each) from 7 correspondences.
// ---- the 7-point solver in three steps (run_7point below chains them; the CUDA kernel runs the middle one
//      cooperatively in a warp, see run_7point_warp in ransac.cu) -------------------------------------------------
struct Fm7Norm { double m1cx, m1cy, m2cx, m2cy, scale1, scale2; };

// Hartley normalisation of the 7 correspondences; false if the points coincide
SVO_HD2 bool fm7_normalise(const Pt2f* m1, const Pt2f* m2, Fm7Norm& nm) {
    double m1cx = 0, m1cy = 0, m2cx = 0, m2cy = 0;
    for (int i = 0; i < 7; ++i) { m1cx += m1[i].x; m1cy += m1[i].y; m2cx += m2[i].x; m2cy += m2[i].y; }
    const double t = 1. / 7;
    m1cx *= t; m1cy *= t; m2cx *= t; m2cy *= t;
    double scale1 = 0, scale2 = 0;
    for (int i = 0; i < 7; ++i) {
        const double ax = m1[i].x - m1cx, ay = m1[i].y - m1cy, bx = m2[i].x - m2cx, by = m2[i].y - m2cy;
        scale1 += sqrt(ax * ax + ay * ay);
        scale2 += sqrt(bx * bx + by * by);
    }
    scale1 *= t; scale2 *= t;
    if (scale1 < FLT_EPSILON || scale2 < FLT_EPSILON) return false;
    nm.m1cx = m1cx; nm.m1cy = m1cy; nm.m2cx = m2cx; nm.m2cy = m2cy;
    nm.scale1 = sqrt(2.) / scale1;
    nm.scale2 = sqrt(2.) / scale2;
    return true;
}

// entry (i, j) of the 7x9 system: rows [x1*x0, x1*y0, x1, y1*x0, y1*y0, y1, x0, y0, 1]
SVO_HD2 double fm7_entry(const Pt2f* m1, const Pt2f* m2, const Fm7Norm& nm, int i, int j) {
    const double x0 = (m1[i].x - nm.m1cx) * nm.scale1, y0 = (m1[i].y - nm.m1cy) * nm.scale1;
    const double x1 = (m2[i].x - nm.m2cx) * nm.scale2, y1 = (m2[i].y - nm.m2cy) * nm.scale2;
    const double p = j < 3 ? x1 : (j < 6 ? y1 : 1.0);
    const int jj = j % 3;
    const double q = jj == 0 ? x0 : (jj == 1 ? y0 : 1.0);
    return (j < 6 && jj < 2) ? p * q : (j < 6 ? p : q);          // the products only where the reference form has them
}

// cv::hypot of core/lapack.cpp (not libm's)
SVO_HD2 double cv_hypot(double a, double b) {
    a = fabs(a);
    b = fabs(b);
    if (a > b) { b /= a; return a * sqrt(1 + b * b); }
    if (b > 0) { a /= b; return b * sqrt(1 + a * a); }
    return 0;
}

// SVDecomp(A, W, U, Vt, MODIFY_A + FULL_UV) of run7Point for the 7x9 matrix A: SVD::compute works on the transposed
// problem (m = 9, n = 7, n1 = 9), i.e. JacobiSVDImpl_ orthogonalises the 7 rows of A in place and appends two rows;
// rows 7 and 8 of the result are f1 and f2.  U (= Vt of the transposed problem) is not needed and not formed: it
// never feeds back into the rows.  At: 9 rows of 9 doubles, rows 0..6 = A on entry.
SVO_HD2 void fm7_svd_rows(double (*At)[9]) {
    const int m = 9, n = 7, n1 = 9;
    const double minval = DBL_MIN, eps = DBL_EPSILON * 10;
    double W[7];
    for (int i = 0; i < n; i++) {
        double sd = 0;
        for (int k = 0; k < m; k++) { const double t = At[i][k]; sd += t * t; }
        W[i] = sd;
    }
    const int max_iter = 30;                              // std::max(m, 30)
    for (int iter = 0; iter < max_iter; iter++) {
        bool changed = false;
        for (int i = 0; i < n - 1; i++)
            for (int j = i + 1; j < n; j++) {
                double* Ai = At[i];
                double* Aj = At[j];
                double a = W[i], p = 0, b = W[j];
                for (int k = 0; k < m; k++) p += Ai[k] * Aj[k];
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                const double beta = a - b, gamma = cv_hypot(p, beta);
                double c, s;
                if (beta < 0) {
                    const double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
                for (int k = 0; k < m; k++) {
                    const double t0 = c * Ai[k] + s * Aj[k];
                    const double t1 = -s * Ai[k] + c * Aj[k];
                    Ai[k] = t0; Aj[k] = t1;
                    a += t0 * t0; b += t1 * t1;
                }
                W[i] = a; W[j] = b;
                changed = true;
            }
        if (!changed) break;
    }
    for (int i = 0; i < n; i++) {
        double sd = 0;
        for (int k = 0; k < m; k++) { const double t = At[i][k]; sd += t * t; }
        W[i] = sqrt(sd);
    }
    for (int i = 0; i < n - 1; i++) {                     // selection sort, descending
        int j = i;
        for (int k = i + 1; k < n; k++) if (W[j] < W[k]) j = k;
        if (i != j) {
            const double tw = W[i]; W[i] = W[j]; W[j] = tw;
            for (int k = 0; k < m; k++) { const double t = At[i][k]; At[i][k] = At[j][k]; At[j][k] = t; }
        }
    }
    CvRng rng;
    rng.state = 0x12345678;
    for (int i = 0; i < n1; i++) {
        double sd = i < n ? W[i] : 0;
        for (int ii = 0; ii < 100 && sd <= minval; ii++) {
            // a zero singular value: random +-1/m vector, projected off the previous rows (twice), normalised
            const double val0 = 1. / m;
            for (int k = 0; k < m; k++) At[i][k] = (rng.next() & 256) != 0 ? val0 : -val0;
            for (int iter = 0; iter < 2; iter++)
                for (int j = 0; j < i; j++) {
                    sd = 0;
                    for (int k = 0; k < m; k++) sd += At[i][k] * At[j][k];
                    double asum = 0;
                    for (int k = 0; k < m; k++) {
                        const double t = At[i][k] - sd * At[j][k];
                        At[i][k] = t;
                        asum += fabs(t);
                    }
                    asum = asum > eps * 100 ? 1 / asum : 0;
                    for (int k = 0; k < m; k++) At[i][k] *= asum;
                }
            sd = 0;
            for (int k = 0; k < m; k++) { const double t = At[i][k]; sd += t * t; }
            sd = sqrt(sd);
        }
        const double s = sd > minval ? 1 / sd : 0.;
        for (int k = 0; k < m; k++) At[i][k] *= s;
    }
}

// cubic in the pencil lambda*f1 + (1-lambda)*f2, roots -> up to 3 de-normalised F matrices.  f1, f2: rows 7, 8 of Vt.
SVO_HD2_CALL int fm7_finish(double* f1, double* f2, const Fm7Norm& nm, double* F /* 27 */) {
    // f ~ lambda*f1 + (1-lambda)*f2, det(f) = 0 gives a cubic in lambda
    for (int i = 0; i < 9; ++i) f1[i] -= f2[i];
    double c[4], r[3] = {0, 0, 0};
    double t0 = f2[4] * f2[8] - f2[5] * f2[7];
    double t1 = f2[3] * f2[8] - f2[5] * f2[6];
    double t2 = f2[3] * f2[7] - f2[4] * f2[6];
    c[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2;
    c[2] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2 -
           f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) + f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) -
           f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) + f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) -
           f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) + f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
    t0 = f1[4] * f1[8] - f1[5] * f1[7];
    t1 = f1[3] * f1[8] - f1[5] * f1[6];
    t2 = f1[3] * f1[7] - f1[4] * f1[6];
    c[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;
    c[1] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 -
           f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) + f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) -
           f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) + f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) -
           f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) + f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
    const int n = solve_cubic(c, r);
    if (n < 1 || n > 3) return n < 0 ? 0 : n;
    // T1 = [s1 0 -s1*c1x; 0 s1 -s1*c1y; 0 0 1], T2 likewise; F = T2^T * F * T1 evaluated as cv::gemm does for 3x3
    // operands: (T2^T * F) first, every entry the left-to-right sum of its three products (zeros included)
    const double T1[9] = {nm.scale1, 0, -nm.scale1 * nm.m1cx, 0, nm.scale1, -nm.scale1 * nm.m1cy, 0, 0, 1};
    const double T2t[9] = {nm.scale2, 0, 0, 0, nm.scale2, 0, -nm.scale2 * nm.m2cx, -nm.scale2 * nm.m2cy, 1};
    for (int k = 0; k < n; ++k) {
        double* fm = F + 9 * k;
        double lambda = r[k], mu = 1.;
        const double s = f1[8] * r[k] + f2[8];
        if (fabs(s) > DBL_EPSILON) { mu = 1. / s; lambda *= mu; fm[8] = 1.; }
        else fm[8] = 0.;
        for (int i = 0; i < 8; ++i) fm[i] = f1[i] * lambda + f2[i] * mu;
        double G[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                G[3 * i + j] = T2t[3 * i] * fm[j] + T2t[3 * i + 1] * fm[3 + j] + T2t[3 * i + 2] * fm[6 + j];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                fm[3 * i + j] = G[3 * i] * T1[j] + G[3 * i + 1] * T1[3 + j] + G[3 * i + 2] * T1[6 + j];
        if (fabs(fm[8]) > FLT_EPSILON) { const double inv = 1. / fm[8]; for (int i = 0; i < 9; ++i) fm[i] *= inv; }
    }
    return n;
}

SVO_HD2 int run_7point(const Pt2f* m1, const Pt2f* m2, double* F /* 27 */) {
    Fm7Norm nm;
    if (!fm7_normalise(m1, m2, nm)) return 0;
    double a[9][9];
    for (int i = 0; i < 7; ++i)
        for (int j = 0; j < 9; ++j) a[i][j] = fm7_entry(m1, m2, nm, i, j);
    for (int j = 0; j < 9; ++j) a[7][j] = a[8][j] = 0;
    fm7_svd_rows(a);
    return fm7_finish(a[7], a[8], nm, F);
}

// FMEstimatorCallback::computeError for one correspondence (float32 result)
SVO_HD2 float fm_error(const double* F, Pt2f m1, Pt2f m2) {
    double a = F[0] * m1.x + F[1] * m1.y + F[2];
    double b = F[3] * m1.x + F[4] * m1.y + F[5];
    double c = F[6] * m1.x + F[7] * m1.y + F[8];
    const double s2 = 1. / (a * a + b * b);
    const double d2 = m2.x * a + m2.y * b + c;
    a = F[0] * m2.x + F[3] * m2.y + F[6];
    b = F[1] * m2.x + F[4] * m2.y + F[7];
    c = F[2] * m2.x + F[5] * m2.y + F[8];
    const double s1 = 1. / (a * a + b * b);
    const double d1 = m1.x * a + m1.y * b + c;
    return (float)fmax(d1 * d1 * s1, d2 * d2 * s2);
}

constexpr int kFmModelPoints = 7;
constexpr int kFmMaxIters = 1000;          // findFundamentalMat's default maxIters
constexpr double kFmConfidence = 0.99;     // stage4_match_consecutive.cpp:202
constexpr double kFmThreshold = 1.0;       // ransacReprojThreshold

}  // namespace svo
