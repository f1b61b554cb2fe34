// fundmat_core.cuh -- building blocks of cv::findFundamentalMat(p1, p2, FM_RANSAC, 1.0, 0.99, mask)
// as called at stage4_match_consecutive.cpp:202,237, restated from OpenCV's published algorithm
// (calib3d: RANSACPointSetRegistrator / LMeDSPointSetRegistrator / FMEstimatorCallback):
//   * cv::RNG (multiply-with-carry, state 0xFFFFFFFFFFFFFFFF) and RNG::uniform(0, count)
//   * getSubset: 7 distinct indices, re-drawn while the 7th point is collinear with two others
//   * 7-point algorithm: null space of the 7x9 epipolar system on normalised points, cubic in lambda
//   * symmetric epipolar distance max(d1^2 s1, d2^2 s2) in double, compared as float32
//   * RANSACUpdateNumIters
// The integer parts (RNG, subsets, iteration counts) are exact.  The 7-point solve uses a Gauss-Jordan
// null space instead of LAPACK's SVD: the set of candidate matrices is the same, their floating-point
// rounding differs at the 1e-12 level, which can flip an inlier only if its error lies within that
// distance of the threshold.
// Host + device; the host build is used by tests/cpp/fundmat_host.cpp to pin this file against cv2.
#pragma once
#include <float.h>
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SVO_HD2 __host__ __device__ inline
#else
#define SVO_HD2 inline
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
SVO_HD2 bool last_point_collinear(const Pt2f* p, int count) {
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
SVO_HD2 bool get_subset7(const Pt2f* m1, const Pt2f* m2, int count, CvRng& rng, int max_attempts, int* idx,
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
SVO_HD2 int ransac_update_num_iters(double p, double ep, int model_points, int max_iters) {
    p = fmax(p, 0.); p = fmin(p, 1.);
    ep = fmax(ep, 0.); ep = fmin(ep, 1.);
    double num = fmax(1. - p, DBL_MIN);
    double denom = 1. - pow(1. - ep, (double)model_points);
    if (denom < DBL_MIN) return 0;
    num = log(num);
    denom = log(denom);
    return (denom >= 0 || -num >= max_iters * (-denom)) ? max_iters : cv_round(num / denom);
}

// cv::solveCubic for c[0] x^3 + c[1] x^2 + c[2] x + c[3] = 0; returns the number of real roots
SVO_HD2 int solve_cubic(const double* c, double* r) {
    double a0 = c[0], a1 = c[1], a2 = c[2], a3 = c[3];
    double x0 = 0., x1 = 0., x2 = 0.;
    int n = 0;
    if (a0 == 0) {
        if (a1 == 0) {
            if (a2 == 0) n = a3 == 0 ? -1 : 0;
            else { x0 = -a3 / a2; n = 1; }
        } else {
            double d = a2 * a2 - 4 * a1 * a3;
            if (d >= 0) {
                d = sqrt(d);
                const double q1 = (-a2 + d) * 0.5, q2 = (a2 + d) * -0.5;
                if (fabs(q1) > fabs(q2)) { x0 = q1 / a1; x1 = a3 / q1; }
                else { x0 = q2 / a1; x1 = a3 / q2; }
                n = d > 0 ? 2 : 1;
            }
        }
    } else {
        a0 = 1. / a0; a1 *= a0; a2 *= a0; a3 *= a0;
        const double Q = (a1 * a1 - 3 * a2) * (1. / 9);
        const double R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1. / 54);
        const double Qcubed = Q * Q * Q;
        double d = Qcubed - R * R;
        const double kPi = 3.14159265358979323846;
        if (d > 0) {
            const double theta = acos(R / sqrt(Qcubed));
            const double sqrtQ = sqrt(Q);
            const double t0 = -2 * sqrtQ, t1 = theta * (1. / 3), t2 = a1 * (1. / 3);
            x0 = t0 * cos(t1) - t2;
            x1 = t0 * cos(t1 + (2. * kPi / 3)) - t2;
            x2 = t0 * cos(t1 + (4. * kPi / 3)) - t2;
            n = 3;
        } else if (d == 0) {
            if (R >= 0) { x0 = -2 * pow(R, 1. / 3) - a1 / 3; x1 = pow(R, 1. / 3) - a1 / 3; }
            else { x0 = 2 * pow(-R, 1. / 3) - a1 / 3; x1 = -pow(-R, 1. / 3) - a1 / 3; }
            x2 = 0;
            n = x0 == x1 ? 1 : 2;
            x1 = x0 == x1 ? 0 : x1;
        } else {
            d = sqrt(-d);
            double e = pow(d + fabs(R), 1. / 3);
            if (R > 0) e = -e;
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

// cubic in the pencil f1 + (1 - lambda) ... , roots -> up to 3 de-normalised F matrices.  f1, f2: null-space basis.
SVO_HD2 int fm7_finish(double* f1, double* f2, const Fm7Norm& nm, double* F /* 27 */) {
    const double m1cx = nm.m1cx, m1cy = nm.m1cy, m2cx = nm.m2cx, m2cy = nm.m2cy, scale1 = nm.scale1, scale2 = nm.scale2;
    // orthonormalise (the parametrisation below is basis independent; this only helps conditioning)
    double n1 = 0;
    for (int j = 0; j < 9; ++j) n1 += f1[j] * f1[j];
    n1 = 1. / sqrt(n1);
    for (int j = 0; j < 9; ++j) f1[j] *= n1;
    double dot = 0;
    for (int j = 0; j < 9; ++j) dot += f1[j] * f2[j];
    double n2 = 0;
    for (int j = 0; j < 9; ++j) { f2[j] -= dot * f1[j]; n2 += f2[j] * f2[j]; }
    n2 = 1. / sqrt(n2);
    for (int j = 0; j < 9; ++j) f2[j] *= n2;

    // f ~ lambda*f1 + (1-lambda)*f2, det(f) = 0 gives a cubic in lambda
    for (int i = 0; i < 9; ++i) f1[i] -= f2[i];
    double c[4], r[3];
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
    for (int k = 0; k < n; ++k) {
        double* fm = F + 9 * k;
        double lambda = r[k], mu = 1.;
        const double s = f1[8] * r[k] + f2[8];
        if (fabs(s) > DBL_EPSILON) { mu = 1. / s; lambda *= mu; fm[8] = 1.; }
        else fm[8] = 0.;
        for (int i = 0; i < 8; ++i) fm[i] = f1[i] * lambda + f2[i] * mu;
        // de-normalise: F = T2^T * F * T1 with T = [s 0 -s*cx; 0 s -s*cy; 0 0 1]
        double G[9];
        // G = F * T1
        for (int i = 0; i < 3; ++i) {
            G[3 * i + 0] = fm[3 * i + 0] * scale1;
            G[3 * i + 1] = fm[3 * i + 1] * scale1;
            G[3 * i + 2] = -fm[3 * i + 0] * scale1 * m1cx - fm[3 * i + 1] * scale1 * m1cy + fm[3 * i + 2];
        }
        // F = T2^T * G
        for (int j = 0; j < 3; ++j) {
            fm[0 + j] = scale2 * G[0 + j];
            fm[3 + j] = scale2 * G[3 + j];
            fm[6 + j] = -scale2 * m2cx * G[0 + j] - scale2 * m2cy * G[3 + j] + G[6 + j];
        }
        if (fabs(fm[8]) > FLT_EPSILON) { const double inv = 1. / fm[8]; for (int i = 0; i < 9; ++i) fm[i] *= inv; }
    }
    return n;
}

SVO_HD2 int run_7point(const Pt2f* m1, const Pt2f* m2, double* F /* 27 */) {
    Fm7Norm nm;
    if (!fm7_normalise(m1, m2, nm)) return 0;
    double a[7][9];
    for (int i = 0; i < 7; ++i)
        for (int j = 0; j < 9; ++j) a[i][j] = fm7_entry(m1, m2, nm, i, j);
    // null space of the 7x9 system: Gauss-Jordan with complete pivoting
    int pivcol[7];
    bool used[9];
    for (int j = 0; j < 9; ++j) used[j] = false;
    int rank = 0;
    for (int r = 0; r < 7; ++r) {
        int pr = -1, pc = -1;
        double best = 0;
        for (int i = r; i < 7; ++i)
            for (int j = 0; j < 9; ++j)
                if (!used[j] && fabs(a[i][j]) > best) { best = fabs(a[i][j]); pr = i; pc = j; }
        if (pr < 0 || best < 1e-13) break;
        if (pr != r) for (int j = 0; j < 9; ++j) { const double tmp = a[r][j]; a[r][j] = a[pr][j]; a[pr][j] = tmp; }
        const double inv = 1. / a[r][pc];
        for (int j = 0; j < 9; ++j) a[r][j] *= inv;
        for (int i = 0; i < 7; ++i) {
            if (i == r) continue;
            const double f = a[i][pc];
            if (f != 0) for (int j = 0; j < 9; ++j) a[i][j] -= f * a[r][j];
        }
        used[pc] = true;
        pivcol[r] = pc;
        ++rank;
    }
    // basis vectors of the null space for the two (or more) free columns; use the first two free columns
    double f1[9], f2[9];
    int freec[9], nfree = 0;
    for (int j = 0; j < 9; ++j) if (!used[j]) freec[nfree++] = j;
    if (nfree < 2) return 0;
    for (int v = 0; v < 2; ++v) {
        double* f = v == 0 ? f1 : f2;
        const int fc = freec[v];
        for (int j = 0; j < 9; ++j) f[j] = 0;
        f[fc] = 1;
        for (int r = 0; r < rank; ++r) f[pivcol[r]] = -a[r][fc];
    }
    return fm7_finish(f1, f2, nm, F);
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
