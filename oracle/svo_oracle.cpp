// svo_oracle.cpp -- CPU restatement of the stereo-VO hot path (TEST INFRASTRUCTURE).
//
// This file is the checker, never the product: only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may build, load or call it.  The shipped
// library (stereo_vo_b200/csrc) contains no CPU path and never links against this file.
//
// What it restates (alias -> file under /root/reference/libstereo-odometry):
//   S2 = src/stage2_detect.cpp   S3 = src/stage3_match_left_right.cpp
//   S4 = src/stage4_match_consecutive.cpp   S5 = src/stage5_optimization.cpp
//   H  = include/libstereo-odometry.h       P  = src/process_new_image_pair.cpp
//
// The reference's own code on this path (NMS/top-K, row sort, L/R filter, tracking filter,
// consistency check, Gauss-Newton) is restated line range by line range; it is compiled with
// g++ so that `std::sort` is libstdc++'s introsort, whose tie order decides keypoint indices
// (S2:89, S2:324, S2:242).
//
// The third-party arithmetic the reference calls (OpenCV, un-vendored, un-pinned:
// README.md:15-16 "OpenCV >= 2.4.8, recommended >= 3.0.0"; Eigen JacobiSVD; MRPT 1.x pose
// classes) is restated from the published algorithms and pinned against the real OpenCV in
// this image (cv2 4.13.0) by tests/test_oracle_pins.py and tests/golden/.
//
// Parity status: PARITY UNPINNED BY THE REFERENCE -- its tests hold no vectors for this path (only
// compute_SAD8, tests/computeSAD8_unittest.cpp:20-41) and it cannot be built in this image (MRPT,
// OpenCV C++ headers, Eigen absent).  What this oracle IS pinned against: goldens generated from the
// real OpenCV (cv2 4.13.0) for every third-party call, libstdc++'s std::sort itself, and the
// survey's independently derived Gauss-Newton known-answer values.  The restatement
// of the reference's own glue is checked only against those and against independent restatements
// in tests/test_oracle_pins.py, not against outputs of the reference.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <utility>
#include <vector>

extern "C" {

// Same field order as cv::KeyPoint (28 bytes) and cv::DMatch (16 bytes).
struct OrcKeyPoint { float x, y, size, angle, response; int32_t octave, class_id; };
struct OrcDMatch   { int32_t queryIdx, trainIdx, imgIdx; float distance; };

struct OrcCamera { double fxl, fyl, cxl, cyl, fxr, fyr, cxr, cyr, baseline; };

struct OrcLSParams {               // TLeastSquaresParams H:313-350, defaults common.cpp:69-82
    int32_t use_robust_kernel;     // true
    double  kernel_param;          // 3
    int32_t max_iters;             // 100
    int32_t initial_max_iters;     // 10
    double  min_mod_out_vector;    // 1e-3
    int32_t max_incr_cost;         // 3
    double  residual_threshold;    // 10
};

struct OrcGNResult {
    int32_t valid;
    int32_t error_code;            // VOErrorCode H:142
    int32_t num_it, num_it_final;
    double  delta_pose[6];         // (w1,w2,w3,t1,t2,t3) after the last update (S5:576-577,678-679)
    double  out_pose_rotvec[3];    // inverse: -w
    double  out_pose_t[3];         // inverse: -R(w)^T t
    double  out_pose_mat[16];      // row-major 4x4 of the inverse
    double  out_pose_ypr[3];       // yaw, pitch, roll of the inverse (MRPT convention)
    int32_t n_outliers_field;      // size of result.outliers (which holds the INLIERS, S5:601-610)
    int32_t n_residuals;           // size of result.out_residual
    int32_t reached_end;           // 1 if the function ran to S5:717 (pose written)
    int32_t bad_cond;              // 1 if m_evalRGN set m_error = voecBadCondNumber (S5:384)
    double  last_cost;
};

}  // extern "C"

namespace {

enum { voecNone = 0, voecBadCondNumber, voecIncrFuncCostStg1, voecIncrFuncCostStg2,
       voecFirstIteration, voecBadTracking };

// ---------------------------------------------------------------------------------------
// FAST-9/16  (OpenCV FAST_t<16> + cornerScore<16>, called at S2:510-511)
// ---------------------------------------------------------------------------------------
const int kRing[16][2] = {{0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3},
                          {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

inline int fast_score(const uint8_t* p, int stride) {
    int d[16 + 9];
    const int v = p[0];
    for (int k = 0; k < 16; ++k) d[k] = v - p[kRing[k][1] * stride + kRing[k][0]];
    for (int k = 0; k < 9; ++k) d[16 + k] = d[k];
    int best = INT32_MIN;
    for (int s = 0; s < 16; ++s) {
        int mn = d[s], mx = d[s];
        for (int j = 1; j < 9; ++j) { mn = std::min(mn, d[s + j]); mx = std::max(mx, d[s + j]); }
        best = std::max(best, std::max(mn, -mx));
    }
    return best - 1;
}

// ---------------------------------------------------------------------------------------
// 7x7 sigma=2 blur with the float32 FMA order of cv::sepFilter2D as used inside
// cv::ORB::compute (S2:512).  Kernel = cv::getGaussianKernel(7, 2, CV_32F).
// ---------------------------------------------------------------------------------------
const uint32_t kGaussBits[7] = {0x3D8FAFB1u, 0x3E06387Eu, 0x3E434A39u, 0x3E5D4AE0u,
                                0x3E434A39u, 0x3E06387Eu, 0x3D8FAFB1u};
inline float gauss_k(int i) { float f; std::memcpy(&f, &kGaussBits[i], 4); return f; }
inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) { if (i < 0) i = -i; else i = 2 * (n - 1) - i; }
    return i;
}

const int8_t kOrbPattern[256][4] = {
#include "orb_pattern.inc"
};

// libstdc++ std::sort or std::stable_sort over an index vector
template <class Cmp>
void sort_indices(std::vector<size_t>& idx, Cmp cmp, int sort_mode) {
    if (sort_mode == 0) std::sort(idx.begin(), idx.end(), cmp);
    else std::stable_sort(idx.begin(), idx.end(), cmp);
}

// ---------------------------------------------------------------------------------------
// 6x6 SVD (restates what Eigen::JacobiSVD delivers at S5:375-388: singular values sorted
// descending, solve() = V * diag(1/sigma_i for sigma_i above threshold) * U^T * rhs)
// One-sided Jacobi (Hestenes) on the columns of A.
// ---------------------------------------------------------------------------------------
struct Svd6 {
    double U[6][6], V[6][6], s[6];
    bool finite;
};

void svd6(const double Hm[6][6], Svd6& out) {
    double A[6][6];
    out.finite = true;
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            A[i][j] = Hm[i][j];
            out.V[i][j] = (i == j) ? 1.0 : 0.0;
            if (!std::isfinite(Hm[i][j])) out.finite = false;
        }
    if (!out.finite) { for (int i = 0; i < 6; ++i) out.s[i] = std::numeric_limits<double>::quiet_NaN(); return; }
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < 5; ++p)
            for (int q = p + 1; q < 6; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < 6; ++i) {
                    alpha += A[i][p] * A[i][p];
                    beta += A[i][q] * A[i][q];
                    gamma += A[i][p] * A[i][q];
                }
                if (gamma == 0.0 || std::fabs(gamma) <= 1e-300) continue;
                if (std::fabs(gamma) <= DBL_EPSILON * std::sqrt(alpha * beta)) continue;
                rotated = true;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
                for (int i = 0; i < 6; ++i) {
                    const double ap = A[i][p], aq = A[i][q];
                    A[i][p] = c * ap - s * aq;
                    A[i][q] = s * ap + c * aq;
                    const double vp = out.V[i][p], vq = out.V[i][q];
                    out.V[i][p] = c * vp - s * vq;
                    out.V[i][q] = s * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
    int order[6];
    double nrm[6];
    for (int j = 0; j < 6; ++j) {
        double n2 = 0;
        for (int i = 0; i < 6; ++i) n2 += A[i][j] * A[i][j];
        nrm[j] = std::sqrt(n2);
        order[j] = j;
    }
    std::stable_sort(order, order + 6, [&](int a, int b) { return nrm[a] > nrm[b]; });
    double Vs[6][6];
    for (int k = 0; k < 6; ++k) {
        const int j = order[k];
        out.s[k] = nrm[j];
        for (int i = 0; i < 6; ++i) {
            out.U[i][k] = nrm[j] > 0 ? A[i][j] / nrm[j] : 0.0;
            Vs[i][k] = out.V[i][j];
        }
    }
    std::memcpy(out.V, Vs, sizeof(Vs));
}

void svd6_solve(const Svd6& S, const double g[6], double x[6]) {
    // Eigen SVDBase::rank(): sigma_i > max(sigma_0 * diagSize * eps, realmin)
    const double thr = std::max(S.s[0] * 6.0 * DBL_EPSILON, DBL_MIN);
    double tmp[6];
    for (int k = 0; k < 6; ++k) {
        double acc = 0;
        for (int i = 0; i < 6; ++i) acc += S.U[i][k] * g[i];
        tmp[k] = (S.s[k] > thr) ? acc / S.s[k] : 0.0;
    }
    for (int i = 0; i < 6; ++i) {
        double acc = 0;
        for (int k = 0; k < 6; ++k) acc += S.V[i][k] * tmp[k];
        x[i] = acc;
    }
}

// ---------------------------------------------------------------------------------------
// m_pinhole_stereo_projection  S5:35-257 (rotation, derivatives incl. the typo at S5:162)
// ---------------------------------------------------------------------------------------
struct RotDeriv {
    double r[3][3];
    double dr[3][3][3];  // dr[row][col][k] = d r(row,col) / d w_k
    bool small;
};

void rot_and_derivs(const double* dp, RotDeriv& o) {
    const double w1 = dp[0], w2 = dp[1], w3 = dp[2];
    const double w12 = w1 * w1, w22 = w2 * w2, w32 = w3 * w3;
    const double tt = std::sqrt(w1 * w1 + w2 * w2 + w3 * w3);
    const double tt2 = tt * tt, tt3 = tt2 * tt, tt4 = tt3 * tt;
    const double sin_tt = std::sin(tt), cos_tt = std::cos(tt);
    std::memset(o.dr, 0, sizeof(o.dr));
    o.small = tt < 1e-5;
    if (o.small) {  // S5:65-97
        o.r[0][0] = 1;   o.r[0][1] = -w3; o.r[0][2] = w2;
        o.r[1][0] = w3;  o.r[1][1] = 1;   o.r[1][2] = -w1;
        o.r[2][0] = -w2; o.r[2][1] = w1;  o.r[2][2] = 1;
        o.dr[0][1][2] = -1; o.dr[0][2][1] = 1;
        o.dr[1][0][2] = 1;  o.dr[1][2][0] = -1;
        o.dr[2][0][1] = -1; o.dr[2][1][0] = 1;
        return;
    }
    const double u = (cos_tt - 1) / tt2;                                            // S5:102
    const double dudw1 = ((-sin_tt * w1 / tt) * tt2 - (cos_tt - 1) * 2 * w1) / tt4;  // S5:103-105
    const double dudw2 = ((-sin_tt * w2 / tt) * tt2 - (cos_tt - 1) * 2 * w2) / tt4;
    const double dudw3 = ((-sin_tt * w3 / tt) * tt2 - (cos_tt - 1) * 2 * w3) / tt4;
    const double v = sin_tt / tt;                                                   // S5:107
    const double dvdw1 = w1 * (tt * cos_tt - sin_tt) / tt3;                         // S5:108-110
    const double dvdw2 = w2 * (tt * cos_tt - sin_tt) / tt3;
    const double dvdw3 = w3 * (tt * cos_tt - sin_tt) / tt3;
    // S5:113-123
    o.r[0][0] = (w22 + w32) * u + 1;    o.r[0][1] = -w3 * v - w1 * w2 * u;  o.r[0][2] = w2 * v - w1 * w3 * u;
    o.r[1][0] = w3 * v - w1 * w2 * u;   o.r[1][1] = (w12 + w32) * u + 1;    o.r[1][2] = -w1 * v - w2 * w3 * u;
    o.r[2][0] = -w2 * v - w1 * w3 * u;  o.r[2][1] = w1 * v - w2 * w3 * u;   o.r[2][2] = (w12 + w22) * u + 1;
    // S5:126-136
    o.dr[0][0][0] = (w22 + w32) * dudw1;
    o.dr[0][0][1] = 2 * w2 * u + (w22 + w32) * dudw2;
    o.dr[0][0][2] = 2 * w3 * u + (w22 + w32) * dudw3;
    o.dr[0][1][0] = -w3 * dvdw1 - (w2 * u + w1 * w2 * dudw1);
    o.dr[0][1][1] = -w3 * dvdw2 - (w1 * u + w1 * w2 * dudw2);
    o.dr[0][1][2] = -(v + w3 * dvdw3) - w1 * w2 * dudw3;
    o.dr[0][2][0] = w2 * dvdw1 - (w3 * u + w1 * w3 * dudw1);
    o.dr[0][2][1] = (v + w2 * dvdw2) - w1 * w3 * dudw2;
    o.dr[0][2][2] = w2 * dvdw3 - (w1 * u + w1 * w3 * dudw3);
    // S5:139-149
    o.dr[1][0][0] = w3 * dvdw1 - (w2 * u + w1 * w2 * dudw1);
    o.dr[1][0][1] = w3 * dvdw2 - (w1 * u + w1 * w2 * dudw2);
    o.dr[1][0][2] = (v + w3 * dvdw3) - w1 * w2 * dudw3;
    o.dr[1][1][0] = 2 * w1 * u + (w12 + w32) * dudw1;
    o.dr[1][1][1] = (w12 + w32) * dudw2;
    o.dr[1][1][2] = 2 * w3 * u + (w12 + w32) * dudw3;
    o.dr[1][2][0] = -(v + w1 * dvdw1) - w2 * w3 * dudw1;
    o.dr[1][2][1] = -w1 * dvdw2 - (w3 * u + w2 * w3 * dudw2);
    o.dr[1][2][2] = -w1 * dvdw3 - (w2 * u + w2 * w3 * dudw3);
    // S5:152-162
    o.dr[2][0][0] = -w2 * dvdw1 - (w3 * u + w1 * w3 * dudw1);
    o.dr[2][0][1] = -(v + w2 * dvdw2) - w1 * w3 * dudw2;
    o.dr[2][0][2] = -w2 * dvdw3 - (w1 * u + w1 * w3 * dudw3);
    o.dr[2][1][0] = (v + w1 * dvdw1) - w2 * w3 * dudw1;
    o.dr[2][1][1] = w1 * dvdw2 - (w3 * u + w2 * w3 * dudw2);
    o.dr[2][1][2] = w1 * dvdw3 - (w2 * u + w2 * w3 * dudw3);
    o.dr[2][2][0] = 2 * w1 * u + (w12 + w22) * dudw1;
    o.dr[2][2][1] = 2 * w2 * u + (w12 + w22) * dudw2;
    o.dr[2][2][2] = (w22 + w32) * dudw3;  // sic: reference typo at S5:162 (should be w12+w22)
}

struct Lmk { double x, y, z; };

// projection + 4x6 jacobian of one landmark, S5:169-256
inline void project_one(const Lmk& L, const OrcCamera& cam, const double* dp, const RotDeriv& R,
                        float px[4], double J[4][6]) {
    const double X1p = L.x, Y1p = L.y, Z1p = L.z;
    const double X1c = R.r[0][0] * X1p + R.r[0][1] * Y1p + R.r[0][2] * Z1p + dp[3];
    const double Y1c = R.r[1][0] * X1p + R.r[1][1] * Y1p + R.r[1][2] * Z1p + dp[4];
    const double Z1c = R.r[2][0] * X1p + R.r[2][1] * Y1p + R.r[2][2] * Z1p + dp[5];
    const double X2c = X1c - cam.baseline;
    px[0] = (float)(cam.fxl * X1c / Z1c + cam.cxl);  // TPixelCoordf is float32 (S5:188-193)
    px[1] = (float)(cam.fyl * Y1c / Z1c + cam.cyl);
    px[2] = (float)(cam.fxr * X2c / Z1c + cam.cxr);
    px[3] = (float)(cam.fyr * Y1c / Z1c + cam.cyr);
    for (int j = 0; j < 6; ++j) {
        double X1cd, Y1cd, Z1cd;
        if (j < 3) {
            if (R.small) {  // S5:207-210, 219-222, 231-234
                if (j == 0)      { X1cd = 0;    Y1cd = -Z1p; Z1cd = Y1p; }
                else if (j == 1) { X1cd = Z1p;  Y1cd = 0;    Z1cd = -X1p; }
                else             { X1cd = -Y1p; Y1cd = X1p;  Z1cd = 0; }
            } else {
                X1cd = R.dr[0][0][j] * X1p + R.dr[0][1][j] * Y1p + R.dr[0][2][j] * Z1p;
                Y1cd = R.dr[1][0][j] * X1p + R.dr[1][1][j] * Y1p + R.dr[1][2][j] * Z1p;
                Z1cd = R.dr[2][0][j] * X1p + R.dr[2][1][j] * Y1p + R.dr[2][2][j] * Z1p;
            }
        } else {
            X1cd = (j == 3); Y1cd = (j == 4); Z1cd = (j == 5);
        }
        J[0][j] = cam.fxl * (X1cd * Z1c - X1c * Z1cd) / (Z1c * Z1c);  // S5:251-254
        J[1][j] = cam.fyl * (Y1cd * Z1c - Y1c * Z1cd) / (Z1c * Z1c);
        J[2][j] = cam.fxr * (X1cd * Z1c - X2c * Z1cd) / (Z1c * Z1c);
        J[3][j] = cam.fyr * (Y1cd * Z1c - Y1c * Z1cd) / (Z1c * Z1c);
    }
}

// m_evalRGN  S5:275-390
bool eval_rgn(const float* l2l, const float* l2r, size_t nL, const std::vector<char>& mask,
              const std::vector<Lmk>& lmks, const double* deltaPose, const OrcCamera& cam,
              const OrcLSParams& ls, double newPose[6], double grad[6],
              std::vector<double>& out_residual, double& out_cost, int& out_error_code) {
    out_residual.resize(nL, std::numeric_limits<double>::max());  // S5:296 (only NEW elements get DBL_MAX)
    out_cost = 0;
    for (int i = 0; i < 6; ++i) grad[i] = 0;
    out_error_code = voecNone;
    double Hm[6][6] = {{0}};
    RotDeriv R;
    rot_and_derivs(deltaPose, R);
    const double b2 = ls.use_robust_kernel ? ls.kernel_param * ls.kernel_param : 0;  // S5:316-317
    const double b2_1 = ls.use_robust_kernel ? 1. / b2 : 0;
    for (size_t m = 0, i = 0; m < nL; ++m) {
        if (!mask[m]) continue;
        float px[4];
        double J[4][6];
        project_one(lmks[i], cam, deltaPose, R, px, J);
        bool good = true;  // m_jacobian_is_good H:919-928
        for (int r = 0; r < 4 && good; ++r)
            for (int c = 0; c < 6; ++c)
                if (!std::isfinite(J[r][c])) { good = false; break; }
        if (!good) { ++i; continue; }  // S5:322
        // residual: float32 subtraction widened to double (S5:335-338)
        const double ri[4] = {(double)(l2l[2 * m] - px[0]), (double)(l2l[2 * m + 1] - px[1]),
                              (double)(l2r[2 * m] - px[2]), (double)(l2r[2 * m + 1] - px[3])};
        const double s = ri[0] * ri[0] + ri[1] * ri[1] + ri[2] * ri[2] + ri[3] * ri[3];  // S5:344
        out_residual[m] = s;
        double rho_p = 1, fi;
        if (ls.use_robust_kernel) {  // S5:351-356
            const double n = std::sqrt(1 + (s * b2_1));
            rho_p = 1 / n;
            fi = b2 * (n - 1);
        } else {
            fi = 0.5 * s;
        }
        out_cost += fi;
        for (int a = 0; a < 6; ++a) {  // S5:364-369: g weighted, H NOT weighted
            double jtr = 0;
            for (int r = 0; r < 4; ++r) jtr += J[r][a] * ri[r];
            grad[a] += rho_p * jtr;
            for (int b = 0; b < 6; ++b) {
                double hab = 0;
                for (int r = 0; r < 4; ++r) hab += J[r][a] * J[r][b];
                Hm[a][b] += hab;
            }
        }
        ++i;
    }
    Svd6 S;  // S5:375-388
    svd6(Hm, S);
    const double cond = S.s[0] / S.s[5];
    if (std::isnan(cond)) { out_error_code = voecBadCondNumber; return false; }
    svd6_solve(S, grad, newPose);
    return true;
}

void rodrigues(const double w[3], double Rm[3][3]) {
    const double th = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    double a, b;  // R = I + a*K + b*K^2
    if (th < 1e-12) { a = 1.0; b = 0.5; }
    else { a = std::sin(th) / th; b = (1 - std::cos(th)) / (th * th); }
    const double K[3][3] = {{0, -w[2], w[1]}, {w[2], 0, -w[0]}, {-w[1], w[0], 0}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double k2 = 0;
            for (int k = 0; k < 3; ++k) k2 += K[i][k] * K[k][j];
            Rm[i][j] = (i == j ? 1.0 : 0.0) + a * K[i][j] + b * k2;
        }
}

// aux m_non_max_sup  S2:225-283
void aux_non_max_sup(const float* xy, const float* response, size_t n_feats, std::vector<char>& survivors,
                     size_t imgH, size_t imgW, size_t num_out_points, int min_distance, int sort_mode) {
    survivors.assign(n_feats, 0);
    std::vector<size_t> sorted_indices(n_feats);
    for (size_t i = 0; i < n_feats; i++) sorted_indices[i] = i;
    sort_indices(sorted_indices, [&](size_t a, size_t b) { return response[a] > response[b]; }, sort_mode);
    const unsigned int cell = (unsigned int)(min_distance / 2.0);   // S2:249
    const float cell_inv = 1.0f / cell;                             // S2:250
    const unsigned int grid_lx = (unsigned int)(1 + imgW * cell_inv);
    const unsigned int grid_ly = (unsigned int)(1 + imgH * cell_inv);
    std::vector<char> occ((size_t)grid_lx * grid_ly, 0);
    auto at = [&](size_t x, size_t y) -> char& { return occ[x * grid_ly + y]; };
    size_t k = 0, c_feats = 0;
    while (c_feats < num_out_points && k < n_feats) {
        const size_t idx = sorted_indices[k++];
        const size_t sx = size_t(xy[2 * idx] * cell_inv);
        const size_t sy = size_t(xy[2 * idx + 1] * cell_inv);
        if (at(sx, sy)) continue;
        at(sx, sy) = 1;
        if (sx > 0) at(sx - 1, sy) = 1;
        if (sy > 0) at(sx, sy - 1) = 1;
        if (sx < grid_lx - 1) at(sx + 1, sy) = 1;
        if (sy < grid_ly - 1) at(sx, sy + 1) = 1;
        survivors[idx] = 1;
        ++c_feats;
    }
}

}  // namespace

extern "C" {

// ---- FAST: cv::FastFeatureDetector::create(th)->detect(img)  S2:510-511 -----------------
// returns number of keypoints (row-major order); writes at most cap of them.
int orc_fast9(const uint8_t* img, int w, int h, int stride, int th, OrcKeyPoint* out, int cap) {
    if (w < 7 || h < 7) return 0;
    std::vector<int> score((size_t)w * h, 0);
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            const int s = fast_score(img + (size_t)y * stride + x, stride);
            if (s >= th) score[(size_t)y * w + x] = s;
        }
    int n = 0;
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            const int s = score[(size_t)y * w + x];
            if (s <= 0) continue;
            bool keep = true;
            for (int dy = -1; dy <= 1 && keep; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    if (!dx && !dy) continue;
                    if (score[(size_t)(y + dy) * w + x + dx] >= s) { keep = false; break; }
                }
            if (!keep) continue;
            if (n < cap) { OrcKeyPoint k = {(float)x, (float)y, 7.f, -1.f, (float)s, 0, -1}; out[n] = k; }
            ++n;
        }
    return n;
}

// ---- blur used inside cv::ORB::compute  S2:512 ------------------------------------------
void orc_blur7(const uint8_t* img, int w, int h, int stride, uint8_t* out, int out_stride) {
    std::vector<float> rows((size_t)w * (h + 6));
    for (int yy = -3; yy < h + 3; ++yy) {
        const uint8_t* src = img + (size_t)reflect101(yy, h) * stride;
        float* dst = &rows[(size_t)(yy + 3) * w];
        for (int x = 0; x < w; ++x) {
            float r = gauss_k(0) * (float)src[reflect101(x - 3, w)];
            for (int t = 1; t < 7; ++t) r = std::fmaf(gauss_k(t), (float)src[reflect101(x - 3 + t, w)], r);
            dst[x] = r;
        }
    }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            const float* r0 = &rows[(size_t)(y + 3) * w + x];
            float c = gauss_k(3) * r0[0];
            for (int j = 1; j <= 3; ++j) c = std::fmaf(gauss_k(3 + j), r0[(ptrdiff_t)j * w] + r0[-(ptrdiff_t)j * w], c);
            float v = std::nearbyintf(c);  // round-half-even in the default rounding mode
            out[(size_t)y * out_stride + x] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
}

// ---- cv::ORB::create()->compute(img, kps, desc)  S2:512 ---------------------------------
// border filter (31 px), order preserved, in-place compaction of kps; desc is n_out x 32.
int orc_orb_compute(const uint8_t* img, int w, int h, int stride, OrcKeyPoint* kps, int n, uint8_t* desc) {
    std::vector<uint8_t> blur((size_t)w * h);
    orc_blur7(img, w, h, stride, blur.data(), w);
    int m = 0;
    for (int i = 0; i < n; ++i) {
        const OrcKeyPoint k = kps[i];
        if (!(k.x >= 31 && k.x < w - 31 && k.y >= 31 && k.y < h - 31)) continue;
        const int x = (int)k.x, y = (int)k.y;
        uint8_t* d = desc + (size_t)m * 32;
        for (int b = 0; b < 32; ++b) {
            unsigned v = 0;
            for (int t = 0; t < 8; ++t) {
                const int8_t* p = kOrbPattern[b * 8 + t];
                const int va = blur[(size_t)(y + p[1]) * w + x + p[0]];
                const int vb = blur[(size_t)(y + p[3]) * w + x + p[2]];
                v |= (unsigned)(va < vb) << t;
            }
            d[b] = (uint8_t)v;
        }
        kps[m++] = k;
    }
    return m;
}

// ---- m_non_max_sup (8-arg)  S2:296-370 ---------------------------------------------------
// sort_mode 0 = std::sort (faithful), 1 = std::stable_sort.  Returns number of output kps.
// sel_idx (optional) receives, for every output, its index in the input list.
int orc_non_max_sup(int num_out_points, const OrcKeyPoint* kps, int n_feats, const uint8_t* desc,
                    int min_distance, int imgH, int imgW, OrcKeyPoint* out_kp, uint8_t* out_desc,
                    int32_t* sel_idx, int sort_mode) {
    std::vector<size_t> sorted_indices(n_feats);
    for (int i = 0; i < n_feats; i++) sorted_indices[i] = i;
    // mrpt::vision::KeypointResponseSorter: response[a] > response[b]  (S2:322-324)
    sort_indices(sorted_indices, [&](size_t a, size_t b) { return kps[a].response > kps[b].response; }, sort_mode);
    const unsigned int cell = (unsigned int)(min_distance / 2.0);  // S2:331
    const float cell_inv = 1.0f / cell;                            // S2:332
    const unsigned int grid_lx = (unsigned int)(1 + imgW * cell_inv);
    const unsigned int grid_ly = (unsigned int)(1 + imgH * cell_inv);
    std::vector<char> occ((size_t)grid_lx * grid_ly, 0);
    auto at = [&](size_t x, size_t y) -> char& { return occ[x * grid_ly + y]; };
    size_t k = 0, c_feats = 0;
    while (c_feats < (size_t)num_out_points && k < (size_t)n_feats) {  // S2:342
        const size_t idx = sorted_indices[k++];
        const OrcKeyPoint& kp = kps[idx];
        const size_t sx = size_t(kp.x * cell_inv), sy = size_t(kp.y * cell_inv);
        if (at(sx, sy)) continue;
        at(sx, sy) = 1;
        if (sx > 0) at(sx - 1, sy) = 1;
        if (sy > 0) at(sx, sy - 1) = 1;
        if (sx < grid_lx - 1) at(sx + 1, sy) = 1;
        if (sy < grid_ly - 1) at(sx, sy + 1) = 1;
        out_kp[c_feats] = kp;
        if (desc && out_desc) std::memcpy(out_desc + c_feats * 32, desc + idx * 32, 32);
        if (sel_idx) sel_idx[c_feats] = (int32_t)idx;
        ++c_feats;
    }
    return (int)c_feats;
}

// ---- m_update_indexes  S2:65-130 (KpRowSorter H:120-127) --------------------------------
// sorts kps+desc in place by pt.y ascending; perm (optional) receives sorted_indices;
// row_index (optional, imgH entries) receives pyr_feats_index with the reference's quirks.
void orc_update_indexes(OrcKeyPoint* kps, uint8_t* desc, int n, int imgH, int32_t* perm,
                        int64_t* row_index, int sort_mode) {
    std::vector<size_t> sorted_indices(n);
    for (int i = 0; i < n; i++) sorted_indices[i] = i;
    sort_indices(sorted_indices, [&](size_t a, size_t b) { return kps[a].y < kps[b].y; }, sort_mode);
    std::vector<OrcKeyPoint> aux(n);
    std::vector<uint8_t> auxd((size_t)n * 32);
    for (int i = 0; i < n; ++i) {
        aux[i] = kps[sorted_indices[i]];
        if (desc) std::memcpy(&auxd[(size_t)i * 32], desc + sorted_indices[i] * 32, 32);
        if (perm) perm[i] = (int32_t)sorted_indices[i];
    }
    if (n) std::memcpy(kps, aux.data(), sizeof(OrcKeyPoint) * n);
    if (desc && n) std::memcpy(desc, auxd.data(), (size_t)n * 32);
    if (!row_index) return;
    // S2:103-129: vector is resized (new entries zero); ranges [fromRow,toRow) are filled, the
    // tail after the last feature's row keeps its zeros.
    std::fill(row_index, row_index + imgH, 0);
    int64_t feats_till_now = 0;
    size_t current_row = 0, from = 0;
    for (int i = 0; i < n; ++i) {
        if (i == 0) {
            current_row = (size_t)kps[i].y;
            from = current_row;
            continue;
        }
        if (kps[i].y == (float)int(current_row)) { ++feats_till_now; continue; }
        current_row = (size_t)kps[i].y;
        ++feats_till_now;
        for (size_t r = from; r < current_row && r < (size_t)imgH; ++r) row_index[r] = feats_till_now;
        from = current_row;
    }
}

// ---- cv::BFMatcher(NORM_HAMMING,false).match(query, train)  S3:88-94, S4:137-142 ---------
int orc_bf_match(const uint8_t* q, int nq, const uint8_t* t, int nt, OrcDMatch* out) {
    if (nq == 0 || nt == 0) return 0;
    for (int i = 0; i < nq; ++i) {
        const uint64_t* a = (const uint64_t*)(q + (size_t)i * 32);
        uint64_t a0, a1, a2, a3;
        std::memcpy(&a0, a, 8); std::memcpy(&a1, a + 1, 8); std::memcpy(&a2, a + 2, 8); std::memcpy(&a3, a + 3, 8);
        int best = INT32_MAX, bi = -1;
        for (int j = 0; j < nt; ++j) {
            uint64_t b[4];
            std::memcpy(b, t + (size_t)j * 32, 32);
            const int d = __builtin_popcountll(a0 ^ b[0]) + __builtin_popcountll(a1 ^ b[1]) +
                          __builtin_popcountll(a2 ^ b[2]) + __builtin_popcountll(a3 ^ b[3]);
            if (d < best) { best = d; bi = j; }  // strict: lowest index wins ties
        }
        OrcDMatch m = {i, bi, 0, (float)best};
        out[i] = m;
    }
    return nq;
}

// ---- stage3 BF branch post-filter  S3:124-175 -------------------------------------------
// matches: in/out (compacted in place).  Returns number kept.
int orc_lr_filter(OrcDMatch* matches, int n, const OrcKeyPoint* leftKps, const OrcKeyPoint* rightKps,
                  int n_right, int enable_robust_1to1, double max_y_diff, int orb_th, int img_width) {
    std::vector<OrcDMatch> v(matches, matches + n);
    if (enable_robust_1to1) {  // S3:124-148
        std::vector<std::pair<double, size_t> > right_cand(n_right, std::make_pair(-1.0, (size_t)0));
        for (size_t k = 0; k < v.size(); ++k) {
            const size_t idR = v[k].trainIdx;
            if (right_cand[idR].first < 0 || right_cand[idR].first > v[k].distance) {
                right_cand[idR].first = v[k].distance;
                right_cand[idR].second = v[k].queryIdx;
            }
        }
        std::vector<OrcDMatch> w;
        for (size_t k = 0; k < v.size(); ++k)
            if (v[k].queryIdx == int(right_cand[v[k].trainIdx].second)) w.push_back(v[k]);
        v.swap(w);
    }
    const double min_disp = 1;           // S3:155
    const double max_disp = img_width;   // S3:156
    int m = 0;
    for (size_t k = 0; k < v.size(); ++k) {
        const int diff = leftKps[v[k].queryIdx].y - rightKps[v[k].trainIdx].y;  // float -> int truncation S3:162
        const int disp = leftKps[v[k].queryIdx].x - rightKps[v[k].trainIdx].x;  // S3:163
        if (std::abs(diff) > max_y_diff || v[k].distance > orb_th || disp < min_disp || disp > max_disp) continue;
        matches[m++] = v[k];
    }
    return m;
}

// ---- matches_lr_row_index  S3:425-445 ---------------------------------------------------
void orc_lr_row_index(const OrcDMatch* matches, int n, const OrcKeyPoint* leftKps, int n_left, int imgH,
                      int64_t* row_index /* imgH+1 */) {
    size_t idx = 0;
    for (int y = 0; y < imgH; y++) {
        row_index[y] = (int64_t)idx;
        while (idx < (size_t)n && leftKps[matches[idx].queryIdx].y <= int(y)) idx++;
    }
    row_index[imgH] = n_left;  // sic: number of left FEATURES (S3:443)
}

// ---- stage 3, row-by-row descriptor branch (smDescRbR)  S3:185-419 ----------------------------------------
// Hamming distance the way that branch computes it: byte-wise Kernighan popcount summed in a uint8_t (S3:321-331),
// i.e. modulo 256 (identical to the true distance except for 256 -> 0).
static inline unsigned rbr_hamming_u8(const uint8_t* a, const uint8_t* b) {
    uint8_t d = 0;
    for (int k = 0; k < 32; ++k) d = (uint8_t)(d + __builtin_popcount((unsigned)(a[k] ^ b[k])));
    return d;
}

// idxL / idxR: pyr_feats_index of the two images as m_update_indexes leaves them (orc_update_indexes), imgH entries.
// out: at most nL matches in left-index order.  Returns the number of matches.
// Notes kept from the reference: candidate rows come from the quirky row index (so the loop at row y works on the
// left features of row y+1 and on the right rows y-r+1 .. y+r; the last feature row of either image is never
// matched), disparity window [1, int(0.7 W)] on int-truncated coordinates, threshold size_t(orb_max_distance) (the
// parameter, not m_current_orb_th), ratio test inert for descriptors (max_ratio = 1), no epipolar post-filter.
int orc_match_rbr(const OrcKeyPoint* kpsL, const uint8_t* descL, int nL, const int64_t* idxL,
                  const OrcKeyPoint* kpsR, const uint8_t* descR, int nR, const int64_t* idxR,
                  int imgH, int imgW, double max_y_diff, double orb_max_distance, int robust_1to1, OrcDMatch* out) {
    const size_t max_distance = (size_t)orb_max_distance;                  // S3:206
    const int INVALID = -1;
    const uint32_t MAX_D = 0xFFFFFFFFu;
    std::vector<int> left_match(nL, INVALID);                             // S3:231
    std::vector<std::pair<int, uint32_t> > right_assign(nR, std::make_pair(INVALID, MAX_D));
    const int max_disparity = (int)(imgW * 0.7);                           // S3:249
    const size_t nRowsMax = (size_t)imgH;                                  // idxL.size()
    const int tol = (int)std::round(max_y_diff);
    for (size_t y = 0; y + 1 < nRowsMax; ++y) {                            // S3:252
        const size_t L0 = (size_t)idxL[y], L1 = (size_t)idxL[y + 1];
        const size_t rmin = (size_t)std::max(0, (int)y - tol);
        const size_t rmax = std::min((size_t)(imgH - 1), y + (size_t)tol);
        const size_t R0 = (size_t)idxR[rmin], R1 = (size_t)idxR[rmax];
        if (L1 - L0 == 0 || R1 - R0 == 0) continue;                        // size_t differences (S3:261-265)
        for (size_t i = L0; i < L1; ++i) {
            uint32_t min_1 = MAX_D, min_2 = MAX_D;
            int min_idx = INVALID;
            for (size_t j = R0; j < R1; ++j) {
                const int disparity = (int)(kpsL[i].x - kpsR[j].x);        // S3:285
                if (disparity < 1 || disparity > max_disparity) continue;
                const size_t dist = rbr_hamming_u8(descL + i * 32, descR + j * 32);
                if (dist > max_distance) continue;                         // S3:335
                if (dist < min_1) { min_2 = min_1; min_1 = (uint32_t)dist; min_idx = (int)j; }
                else if (dist < min_2) min_2 = (uint32_t)dist;
            }
            if (min_idx == INVALID) continue;
            std::pair<int, uint32_t>& ra = right_assign[min_idx];
            if (ra.first == INVALID) {                                     // both modes: first claim
                left_match[i] = min_idx; ra.first = (int)i; ra.second = min_1;
            } else if (robust_1to1 && min_1 < ra.second) {                 // S3:370-377: keep the BEST, undo the previous one
                left_match[ra.first] = INVALID;
                left_match[i] = min_idx; ra.first = (int)i; ra.second = min_1;
            }
        }
    }
    int m = 0;
    for (int i = 0; i < nL; ++i)                                           // S3:398-410
        if (left_match[i] != INVALID) {
            OrcDMatch d = {i, left_match[i], -1, (float)right_assign[left_match[i]].second};   // cv::DMatch(q, t, d): imgIdx = -1 (S3:404)
            out[m++] = d;
        }
    return m;
}

// ---- stage 4, windowed descriptor branch (ifmDescWin)  S4:435-640 -------------------------------------------
// prev/cur: stereo matches, left/right keypoints, LEFT descriptors and matches_lr_row_index (imgH+1 entries) of the
// two frames.  out_prev / out_cur: candidate pairs (index into the previous / current match list) in current-match
// order, before the F-matrix filter.  Returns their number.
// Notes kept from the reference: only the left descriptors are compared (S4:585-603), distance in a uint8_t, a
// candidate must beat 255 strictly, no distance threshold, windows |dy| <= ifm_win_w (sic) and |dx| <= ifm_win_h on
// both the left and the right keypoint.  One deviation: matches_lr_row_index[imgH] holds the number of left FEATURES
// (S3:443), so a window that reaches the last image row makes the reference read past the match list; the window end
// is clamped to the list here.
int orc_track_win(const OrcDMatch* pm, int nPm, const OrcKeyPoint* pL, const OrcKeyPoint* pR, const uint8_t* pDescL,
                  const int64_t* pRowIdx, const OrcDMatch* cm, int nCm, const OrcKeyPoint* cL, const OrcKeyPoint* cR,
                  const uint8_t* cDescL, const int64_t* cRowIdx, int imgH, int imgW, int win_w, int win_h,
                  int32_t* out_prev, int32_t* out_cur) {
    const int INVALID = -1;
    std::vector<std::pair<int, uint32_t> > current(nCm, std::make_pair(INVALID, 0xFFFFFFFFu));
    const int abs_wx_max = imgW - 1, abs_wy_max = imgH - 1;                // S4:491-492 (descriptor variant)
    for (int y = 0; y < imgH - 1; ++y) {
        const size_t p0 = (size_t)pRowIdx[y], p1 = std::min((size_t)pRowIdx[y + 1], (size_t)nPm);
        if (p1 - p0 == 0 || p1 < p0) continue;
        const int wy_min = std::max(0, y - win_w), wy_max = std::min(abs_wy_max, y + win_w);
        const size_t c0 = (size_t)cRowIdx[wy_min], c1 = std::min((size_t)cRowIdx[wy_max + 1], (size_t)nCm);
        if (c1 - c0 == 0 || c1 < c0) continue;
        for (size_t pi = p0; pi < p1; ++pi) {
            const OrcKeyPoint& a = pL[pm[pi].queryIdx];
            const OrcKeyPoint& b = pR[pm[pi].trainIdx];
            const int xl0 = std::max(0, (int)(a.x - win_h)), xl1 = std::min(abs_wx_max, (int)(a.x + win_h));
            const int xr0 = std::max(0, (int)(b.x - win_h)), xr1 = std::min(abs_wx_max, (int)(b.x + win_h));
            int best_ci = INVALID;
            uint8_t best = 255;
            for (size_t ci = c0; ci < c1; ++ci) {
                const OrcKeyPoint& fl = cL[cm[ci].queryIdx];
                const OrcKeyPoint& fr = cR[cm[ci].trainIdx];
                if (fl.x < xl0 || fl.x > xl1 || fr.x < xr0 || fr.x > xr1) continue;
                const unsigned orb = rbr_hamming_u8(pDescL + (size_t)pm[pi].queryIdx * 32, cDescL + (size_t)cm[ci].queryIdx * 32);
                if (orb < best) { best = (uint8_t)orb; best_ci = (int)ci; }
            }
            if (best_ci == INVALID) continue;
            std::pair<int, uint32_t>& cur = current[best_ci];              // S4:622-634
            if (cur.first == INVALID) { cur.first = (int)pi; cur.second = best; }
            if (cur.first != INVALID && best < cur.second) { cur.first = (int)pi; cur.second = best; }
        }
    }
    int n = 0;
    for (int ci = 0; ci < nCm; ++ci)
        if (current[ci].first != INVALID) { out_prev[n] = current[ci].first; out_cur[n] = ci; ++n; }
    return n;
}

// ---- stage4 greedy distance / collision filter  S4:145-160 ------------------------------
// matL / matR compacted in place in lockstep; returns number kept.
int orc_track_filter(OrcDMatch* matL, OrcDMatch* matR, int n, int curNMatches, int orb_th) {
    std::vector<char> left_train_matched(curNMatches, 0), right_train_matched(curNMatches, 0);
    int m = 0;
    for (int k = 0; k < n; ++k) {
        if (matL[k].distance > orb_th || matR[k].distance > orb_th ||
            left_train_matched[matL[k].trainIdx] || right_train_matched[matR[k].trainIdx]) continue;
        left_train_matched[matL[k].trainIdx] = right_train_matched[matR[k].trainIdx] = 1;
        matL[m] = matL[k];
        matR[m] = matR[k];
        ++m;
    }
    return m;
}

// ---- stage4 F-matrix mask application + consistency check  S4:243-294 -------------------
// inliersL / inliersR may be NULL (mask not available => counts as 0 inliers => filter skipped).
// tracked_pairs: out, pairs (pre_match_idx, cur_match_idx); returns their number.
int orc_track_consistency(const OrcDMatch* matL_in, const OrcDMatch* matR_in, int n,
                          const uint8_t* inliersL, const uint8_t* inliersR, int64_t* tracked_pairs) {
    std::vector<OrcDMatch> matL(matL_in, matL_in + n), matR(matR_in, matR_in + n);
    int nl = 0, nr = 0;
    if (inliersL) for (int k = 0; k < n; ++k) nl += inliersL[k] != 0;
    if (inliersR) for (int k = 0; k < n; ++k) nr += inliersR[k] != 0;
    if (nl >= 8 && nr >= 8) {  // S4:243-255
        std::vector<OrcDMatch> l2, r2;
        for (int k = 0; k < n; ++k)
            if (!(inliersL[k] == 0 || inliersR[k] == 0)) { l2.push_back(matL[k]); r2.push_back(matR[k]); }
        matL.swap(l2);
        matR.swap(r2);
    }
    int m = 0;
    for (size_t k = 0; k < matL.size(); ++k)  // S4:276-294
        if (matL[k].queryIdx == matR[k].queryIdx && matL[k].trainIdx == matR[k].trainIdx) {
            tracked_pairs[2 * m] = matL[k].queryIdx;
            tracked_pairs[2 * m + 1] = matL[k].trainIdx;
            ++m;
        }
    return m;
}

// ---- one Gauss-Newton evaluation (test hook for m_evalRGN S5:275-390) -------------------
int orc_eval_rgn(const float* l1l, const float* l1r, const float* l2l, const float* l2r, int n,
                 const OrcCamera* cam, const OrcLSParams* ls, const double* deltaPose,
                 double* newPose, double* grad, double* residual, double* cost) {
    std::vector<char> mask(n, 1);
    std::vector<Lmk> lmks(n);
    for (int m = 0; m < n; ++m) {
        const double ul = l1l[2 * m], vl = l1l[2 * m + 1], ur = l1r[2 * m];
        const double b_d = cam->baseline / (cam->fxl * (cam->cxr - ur) + cam->fxr * (ul - cam->cxl));
        lmks[m] = Lmk{b_d * cam->fxr * (ul - cam->cxl), b_d * cam->fxr * (vl - cam->cyl), b_d * cam->fxl * cam->fxr};
    }
    std::vector<double> res;
    int ec;
    const bool ok = eval_rgn(l2l, l2r, n, mask, lmks, deltaPose, *cam, *ls, newPose, grad, res, *cost, ec);
    for (int m = 0; m < n; ++m) residual[m] = res[m];
    return ok ? 0 : ec;
}

// ---- stage5_optimize  S5:392-736 (single octave, lists already gathered S5:419-461) ------
// l1l,l1r,l2l,l2r: n x 2 float32 pixel coordinates (prev L/R, cur L/R); resp1l: response of the
// prev-left keypoints (drives the aux NMS order).  cur_match_idx: tracked_pairs[i].second.
// initial_pose: the deltaPose to start from (custom / m_last_computed_pose / zeros, S5:500-507).
// out_residual: n doubles (DBL_MAX where never evaluated); out_inlier_cur_idx: n int64 (result.outliers).
void orc_stage5(const float* l1l, const float* l1r, const float* l2l, const float* l2r, const float* resp1l,
                const int64_t* cur_match_idx, int n, int imgH, int imgW, int min_distance,
                const OrcCamera* camp, const OrcLSParams* lsp, const double* initial_pose, int sort_mode,
                const uint8_t* survivors_override /* test hook: NULL = run the aux NMS */,
                OrcGNResult* res, double* out_residual, int64_t* out_inlier_cur_idx) {
    const OrcCamera& cam = *camp;
    const OrcLSParams& ls = *lsp;
    std::memset(res, 0, sizeof(*res));
    res->error_code = voecNone;
    const size_t nL = (size_t)n;
    std::vector<char> survivors;
    if (survivors_override) survivors.assign(survivors_override, survivors_override + nL);
    else aux_non_max_sup(l1l, resp1l, nL, survivors, imgH, imgW, nL, min_distance, sort_mode);  // S5:465-474

    std::vector<double> residual;
    double pCost = 0, cCost = 0;
    bool done = false, abort = false;
    double deltaPose[6];
    for (int i = 0; i < 6; ++i) deltaPose[i] = initial_pose[i];
    double newPose[6] = {0}, grad[6];

    auto finish_residuals = [&]() {
        res->n_residuals = (int)residual.size();
        for (size_t i = 0; i < residual.size(); ++i) out_residual[i] = residual[i];
    };
    auto triangulate = [&](std::vector<Lmk>& lmks) {  // S5:528-544 / S5:623-638
        size_t cnt = 0;
        for (size_t m = 0; m < nL; ++m) cnt += survivors[m] != 0;
        lmks.resize(cnt);
        for (size_t m = 0, i = 0; m < nL; ++m) {
            if (!survivors[m]) continue;
            const double ul = l1l[2 * m], vl = l1l[2 * m + 1], ur = l1r[2 * m];
            const double b_d = cam.baseline / (cam.fxl * (cam.cxr - ur) + cam.fxr * (ul - cam.cxl));
            lmks[i++] = Lmk{b_d * cam.fxr * (ul - cam.cxl), b_d * cam.fxr * (vl - cam.cyl), b_d * cam.fxl * cam.fxr};
        }
        return cnt;
    };

    std::vector<Lmk> lmks;
    size_t n_non_masked = triangulate(lmks);
    if (n_non_masked < 8) {  // S5:521-526
        finish_residuals();
        res->valid = 0;
        return;
    }
    unsigned int timesInc = 0;  // shared by both loops (S5:547)
    res->num_it = 0;
    while (res->num_it < int(ls.initial_max_iters) && !done && !abort) {  // S5:549-598
        pCost = cCost;
        int ec;
        const bool cond = eval_rgn(l2l, l2r, nL, survivors, lmks, deltaPose, cam, ls, newPose, grad, residual, cCost, ec);
        res->error_code = ec;  // S5:564 writes result.error_code
        if (!cond) { res->valid = 0; res->bad_cond = 1; res->last_cost = cCost; return; }  // S5:569-573 (residuals NOT swapped out)
        for (int k = 0; k < 6; ++k) { deltaPose[k] += newPose[k]; res->delta_pose[k] = deltaPose[k]; }
        if (res->num_it > 0) {
            double m = 0;
            for (int c = 0; c < 6; ++c) m += newPose[c] * newPose[c];
            done = std::sqrt(m) < ls.min_mod_out_vector;
            if (pCost < cCost) {
                if (++timesInc > (unsigned)ls.max_incr_cost) { res->error_code = voecIncrFuncCostStg1; abort = true; }
            }
        }
        res->num_it++;
    }
    // S5:601-611: result.outliers receives the cur-match index of every INLIER
    int n_in = 0;
    for (size_t i = 0; i < residual.size(); ++i) {
        if (residual[i] > ls.residual_threshold) survivors[i] = 0;
        else out_inlier_cur_idx[n_in++] = cur_match_idx[i];
    }
    res->n_outliers_field = n_in;
    n_non_masked = triangulate(lmks);
    if (n_non_masked < 8) {  // S5:616-621
        finish_residuals();
        res->valid = 0;
        return;
    }
    done = false; abort = false;
    res->num_it_final = 0;
    while (res->num_it_final < int(ls.max_iters) && !done && !abort) {  // S5:650-700
        pCost = cCost;
        int ec;  // local out_error_code: result.error_code is NOT updated here (S5:665)
        const bool cond = eval_rgn(l2l, l2r, nL, survivors, lmks, deltaPose, cam, ls, newPose, grad, residual, cCost, ec);
        if (!cond) {
            finish_residuals();
            res->valid = 0;
            res->last_cost = cCost;
            res->bad_cond = 1;  // m_error = voecBadCondNumber is set inside m_evalRGN (S5:384)
            return;
        }
        for (int k = 0; k < 6; ++k) { deltaPose[k] += newPose[k]; res->delta_pose[k] = deltaPose[k]; }
        if (res->num_it_final > 0) {
            double m = 0;
            for (int c = 0; c < 6; ++c) m += newPose[c] * newPose[c];
            done = std::sqrt(m) < ls.min_mod_out_vector;
            if (pCost < cCost) {
                if (++timesInc > (unsigned)ls.max_incr_cost) { abort = true; res->error_code = voecIncrFuncCostStg2; }
            }
        }
        res->num_it_final++;
    }
    // S5:717-718: outPose = CPose3D( CPose3DRotVec(deltaPose).getInverse() )
    for (int k = 0; k < 6; ++k) res->delta_pose[k] = deltaPose[k];
    double Rm[3][3];
    rodrigues(deltaPose, Rm);
    for (int i = 0; i < 3; ++i) {
        res->out_pose_rotvec[i] = -deltaPose[i];
        res->out_pose_t[i] = -(Rm[0][i] * deltaPose[3] + Rm[1][i] * deltaPose[4] + Rm[2][i] * deltaPose[5]);
    }
    double Ri[3][3];
    rodrigues(res->out_pose_rotvec, Ri);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) res->out_pose_mat[4 * i + j] = Ri[i][j];
        res->out_pose_mat[4 * i + 3] = res->out_pose_t[i];
    }
    res->out_pose_mat[12] = res->out_pose_mat[13] = res->out_pose_mat[14] = 0;
    res->out_pose_mat[15] = 1;
    // MRPT CPose3D::getYawPitchRoll
    res->out_pose_ypr[1] = std::atan2(-Ri[2][0], std::hypot(Ri[0][0], Ri[1][0]));
    res->out_pose_ypr[0] = std::atan2(Ri[1][0], Ri[0][0]);
    res->out_pose_ypr[2] = std::atan2(Ri[2][1], Ri[2][2]);
    finish_residuals();
    res->valid = !abort;  // S5:727
    res->reached_end = 1;
    res->last_cost = cCost;
}

// ---- getProjectedCoords  C:415-470 --------------------------------------------------------------------------
// The matches flagged -1 are triangulated (C:428-455, same formula as S5:537-542) and projected with
// m_pinhole_stereo_projection (S5:165-193) for delta_pose = [w t] of the inverse of change_pose (C:459-464; that MRPT
// pose algebra is the caller's).  out: 4 floats per projected match.  Returns their number.
int orc_projected_coords(const OrcDMatch* pre_matches, int n_matches, const OrcKeyPoint* pre_left, const OrcKeyPoint* pre_right,
                         const int32_t* other_matches_tracked, const OrcCamera* cam, const double* delta_pose, float* out) {
    RotDeriv R;
    rot_and_derivs(delta_pose, R);
    int n = 0;
    for (int m = 0; m < n_matches; ++m) {
        if (other_matches_tracked[m] != -1) continue;
        const double ul = pre_left[pre_matches[m].queryIdx].x, vl = pre_left[pre_matches[m].queryIdx].y;
        const double ur = pre_right[pre_matches[m].trainIdx].x;
        const double disparity = cam->fxl * (cam->cxr - ur) + cam->fxr * (ul - cam->cxl);
        const double b_d = cam->baseline / disparity;
        const double X1p = b_d * cam->fxr * (ul - cam->cxl), Y1p = b_d * cam->fxr * (vl - cam->cyl), Z1p = b_d * cam->fxl * cam->fxr;
        const double X1c = R.r[0][0] * X1p + R.r[0][1] * Y1p + R.r[0][2] * Z1p + delta_pose[3];
        const double Y1c = R.r[1][0] * X1p + R.r[1][1] * Y1p + R.r[1][2] * Z1p + delta_pose[4];
        const double Z1c = R.r[2][0] * X1p + R.r[2][1] * Y1p + R.r[2][2] * Z1p + delta_pose[5];
        const double X2c = X1c - cam->baseline;
        out[4 * n + 0] = (float)(cam->fxl * X1c / Z1c + cam->cxl);
        out[4 * n + 1] = (float)(cam->fyl * Y1c / Z1c + cam->cyl);
        out[4 * n + 2] = (float)(cam->fxr * X2c / Z1c + cam->cxr);
        out[4 * n + 3] = (float)(cam->fyr * Y1c / Z1c + cam->cyr);
        ++n;
    }
    return n;
}

// ---- test hooks for tests/test_ref_cpu.py (restatement vs the reference's compiled text) ----
// auxiliary m_non_max_sup S2:225-283 alone
void orc_aux_non_max_sup(const float* xy, const float* response, int n, int imgH, int imgW, int num_out, int min_distance,
                         int sort_mode, uint8_t* survivors_out) {
    std::vector<char> s;
    aux_non_max_sup(xy, response, (size_t)n, s, (size_t)imgH, (size_t)imgW, (size_t)num_out, min_distance, sort_mode);
    for (int i = 0; i < n; ++i) survivors_out[i] = (uint8_t)s[i];
}
// m_pinhole_stereo_projection S5:35-257 alone: pixels (4 floats) and the 4x6 Jacobian per landmark
void orc_pinhole_projection(const double* lmks, int n, const OrcCamera* cam, const double* dp, float* px4, double* jac24) {
    RotDeriv R;
    rot_and_derivs(dp, R);
    for (int i = 0; i < n; ++i) {
        double J[4][6];
        project_one(Lmk{lmks[3 * i], lmks[3 * i + 1], lmks[3 * i + 2]}, *cam, dp, R, px4 + 4 * i, J);
        std::memcpy(jac24 + 24 * i, J, sizeof J);
    }
}

int orc_sizeof_keypoint() { return (int)sizeof(OrcKeyPoint); }
int orc_sizeof_dmatch() { return (int)sizeof(OrcDMatch); }

}  // extern "C"
