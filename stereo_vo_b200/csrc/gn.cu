// gn.cu -- stage 5 of the hot path on sm_100a: robust Gauss-Newton SE(3) refinement of the stereo
// reprojection error, one persistent CTA per stereo pair, no host synchronisation between
// iterations.  Replaces
//   stage5_optimize                stage5_optimization.cpp:392-736
//   m_evalRGN                      stage5_optimization.cpp:275-390
//   m_pinhole_stereo_projection    stage5_optimization.cpp:35-257
//   m_non_max_sup (aux, 5-arg)     stage2_detect.cpp:225-283  (called stage5_optimization.cpp:470-474)
//   Eigen::JacobiSVD solve         stage5_optimization.cpp:375-388
//   CPose3DRotVec::getInverse / CPose3D(rotvec)   stage5_optimization.cpp:717-718
// and keeps the reference's quirks: unweighted Hessian with weighted gradient (:364-365), the
// Jacobian typo at :162, float32 residual subtraction (:335-338), `timesInc` shared by both loops
// (:547), result.outliers holding the inliers (:601-610).
//
// Per correspondence and iteration: 4 residuals + 4x6 Jacobian in FP64, pseudo-Huber weight,
// accumulation of the 21 upper-triangle entries of J^T J, the 6 entries of J^T r and the cost with
// warp shuffles, then a cross-warp sum in shared memory; thread 0 solves the 6x6 system.
// HBM traffic is ~40 B per correspondence (L2 resident across iterations): the kernel is bound by
// FP64 latency and the block-wide reduction, not by bandwidth.
#include "gn.cuh"
#include "introsort_emul.cuh"
#include <cfloat>

namespace svo {

constexpr int GN_THREADS = 64;    // 254 registers/thread: 4 CTAs (stereo pairs) per SM overlap each other's serial 6x6 solves
constexpr int NACC = 28;   // 21 (H upper triangle) + 6 (g) + 1 (cost)

struct RotDeriv {
    double r[3][3];
    double dr[3][3][3];   // dr[row][col][k] = d r(row,col) / d w_k
    int small;
};

// S5:45-163
__device__ void rot_and_derivs(const double* dp, RotDeriv& o) {
    const double w1 = dp[0], w2 = dp[1], w3 = dp[2];
    const double w12 = w1 * w1, w22 = w2 * w2, w32 = w3 * w3;
    const double tt = sqrt(w1 * w1 + w2 * w2 + w3 * w3);
    const double tt2 = tt * tt, tt3 = tt2 * tt, tt4 = tt3 * tt;
    const double sin_tt = sin(tt), cos_tt = cos(tt);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) o.dr[i][j][k] = 0;
    o.small = tt < 1e-5;
    if (o.small) {
        o.r[0][0] = 1;   o.r[0][1] = -w3; o.r[0][2] = w2;
        o.r[1][0] = w3;  o.r[1][1] = 1;   o.r[1][2] = -w1;
        o.r[2][0] = -w2; o.r[2][1] = w1;  o.r[2][2] = 1;
        o.dr[0][1][2] = -1; o.dr[0][2][1] = 1;
        o.dr[1][0][2] = 1;  o.dr[1][2][0] = -1;
        o.dr[2][0][1] = -1; o.dr[2][1][0] = 1;
        return;
    }
    const double u = (cos_tt - 1) / tt2;
    const double dudw1 = ((-sin_tt * w1 / tt) * tt2 - (cos_tt - 1) * 2 * w1) / tt4;
    const double dudw2 = ((-sin_tt * w2 / tt) * tt2 - (cos_tt - 1) * 2 * w2) / tt4;
    const double dudw3 = ((-sin_tt * w3 / tt) * tt2 - (cos_tt - 1) * 2 * w3) / tt4;
    const double v = sin_tt / tt;
    const double dvdw1 = w1 * (tt * cos_tt - sin_tt) / tt3;
    const double dvdw2 = w2 * (tt * cos_tt - sin_tt) / tt3;
    const double dvdw3 = w3 * (tt * cos_tt - sin_tt) / tt3;
    o.r[0][0] = (w22 + w32) * u + 1;    o.r[0][1] = -w3 * v - w1 * w2 * u;  o.r[0][2] = w2 * v - w1 * w3 * u;
    o.r[1][0] = w3 * v - w1 * w2 * u;   o.r[1][1] = (w12 + w32) * u + 1;    o.r[1][2] = -w1 * v - w2 * w3 * u;
    o.r[2][0] = -w2 * v - w1 * w3 * u;  o.r[2][1] = w1 * v - w2 * w3 * u;   o.r[2][2] = (w12 + w22) * u + 1;
    o.dr[0][0][0] = (w22 + w32) * dudw1;
    o.dr[0][0][1] = 2 * w2 * u + (w22 + w32) * dudw2;
    o.dr[0][0][2] = 2 * w3 * u + (w22 + w32) * dudw3;
    o.dr[0][1][0] = -w3 * dvdw1 - (w2 * u + w1 * w2 * dudw1);
    o.dr[0][1][1] = -w3 * dvdw2 - (w1 * u + w1 * w2 * dudw2);
    o.dr[0][1][2] = -(v + w3 * dvdw3) - w1 * w2 * dudw3;
    o.dr[0][2][0] = w2 * dvdw1 - (w3 * u + w1 * w3 * dudw1);
    o.dr[0][2][1] = (v + w2 * dvdw2) - w1 * w3 * dudw2;
    o.dr[0][2][2] = w2 * dvdw3 - (w1 * u + w1 * w3 * dudw3);
    o.dr[1][0][0] = w3 * dvdw1 - (w2 * u + w1 * w2 * dudw1);
    o.dr[1][0][1] = w3 * dvdw2 - (w1 * u + w1 * w2 * dudw2);
    o.dr[1][0][2] = (v + w3 * dvdw3) - w1 * w2 * dudw3;
    o.dr[1][1][0] = 2 * w1 * u + (w12 + w32) * dudw1;
    o.dr[1][1][1] = (w12 + w32) * dudw2;
    o.dr[1][1][2] = 2 * w3 * u + (w12 + w32) * dudw3;
    o.dr[1][2][0] = -(v + w1 * dvdw1) - w2 * w3 * dudw1;
    o.dr[1][2][1] = -w1 * dvdw2 - (w3 * u + w2 * w3 * dudw2);
    o.dr[1][2][2] = -w1 * dvdw3 - (w2 * u + w2 * w3 * dudw3);
    o.dr[2][0][0] = -w2 * dvdw1 - (w3 * u + w1 * w3 * dudw1);
    o.dr[2][0][1] = -(v + w2 * dvdw2) - w1 * w3 * dudw2;
    o.dr[2][0][2] = -w2 * dvdw3 - (w1 * u + w1 * w3 * dudw3);
    o.dr[2][1][0] = (v + w1 * dvdw1) - w2 * w3 * dudw1;
    o.dr[2][1][1] = w1 * dvdw2 - (w3 * u + w2 * w3 * dudw2);
    o.dr[2][1][2] = w1 * dvdw3 - (w2 * u + w2 * w3 * dudw3);
    o.dr[2][2][0] = 2 * w1 * u + (w12 + w22) * dudw1;
    o.dr[2][2][1] = 2 * w2 * u + (w12 + w22) * dudw2;
    o.dr[2][2][2] = (w22 + w32) * dudw3;   // sic, reference typo S5:162
}

// Symmetric 6x6 eigen-decomposition by cyclic Jacobi; stands in for Eigen::JacobiSVD on the
// (symmetric positive semi-definite) Hessian: singular values = |eigenvalues|.
// Returns false when the condition number sigma_max/sigma_min would be NaN (S5:379-386).
// The rotation is applied with the classic symmetric update (only the upper triangle is live):
//   a_pp -= t a_pq, a_qq += t a_pq, a_pq = 0, a_rp -= s (a_rq + tau a_rp), a_rq += s (a_rp - tau a_rq).
// A two-pass "columns then rows" update on a double[6][6] was observed to be miscompiled by nvcc 12.9
// for sm_100a (the column pass acted on rows); this form was validated on the device.
__device__ __forceinline__ double& sym(double* A, int i, int j) { return i <= j ? A[6 * i + j] : A[6 * j + i]; }

// (out of line, loops kept as loops: the rarely taken fallback was 60 KB of the kernel's code when unrolled and inlined twice)
__device__ __noinline__ bool solve6(const double* Hu /* 21 upper-triangle entries, row-major */, const double* g, double* x) {
    double A[36], V[36];
    bool finite = true;
    {
        int k = 0;
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j) {
                V[6 * i + j] = (i == j) ? 1.0 : 0.0;
                if (j >= i) { A[6 * i + j] = Hu[k++]; finite = finite && isfinite(A[6 * i + j]); }
                else A[6 * i + j] = 0.0;
            }
    }
    if (!finite) return false;
#pragma unroll 1
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0, diag = 0;
        for (int i = 0; i < 6; ++i) {
            diag += A[6 * i + i] * A[6 * i + i];
            for (int j = i + 1; j < 6; ++j) off += A[6 * i + j] * A[6 * i + j];
        }
        if (off <= diag * 1e-32) break;
#pragma unroll 1
        for (int p = 0; p < 5; ++p)
#pragma unroll 1
            for (int q = p + 1; q < 6; ++q) {
                const double apq = A[6 * p + q];
                if (apq == 0.0) continue;
                const double theta = (A[6 * q + q] - A[6 * p + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c, tau = sn / (1.0 + c);
                A[6 * p + p] -= t * apq;
                A[6 * q + q] += t * apq;
                A[6 * p + q] = 0.0;
                for (int r = 0; r < 6; ++r) {
                    if (r == p || r == q) continue;
                    double& arp = sym(A, r, p);
                    double& arq = sym(A, r, q);
                    const double vp = arp, vq = arq;
                    arp = vp - sn * (vq + tau * vp);
                    arq = vq + sn * (vp - tau * vq);
                }
                for (int r = 0; r < 6; ++r) {
                    const double vp = V[6 * r + p], vq = V[6 * r + q];
                    V[6 * r + p] = vp - sn * (vq + tau * vp);
                    V[6 * r + q] = vq + sn * (vp - tau * vq);
                }
            }
    }
    double smax = 0, smin = DBL_MAX;
    for (int i = 0; i < 6; ++i) { const double sv = fabs(A[6 * i + i]); smax = fmax(smax, sv); smin = fmin(smin, sv); }
    if (smax == 0.0 && smin == 0.0) return false;          // 0/0 = NaN
    const double thr = fmax(smax * 6.0 * DBL_EPSILON, DBL_MIN);   // Eigen SVDBase::rank() threshold
    for (int i = 0; i < 6; ++i) x[i] = 0;
    for (int k = 0; k < 6; ++k) {
        const double lam = A[6 * k + k];
        if (!(fabs(lam) > thr)) continue;
        double proj = 0;
        for (int i = 0; i < 6; ++i) proj += V[6 * i + k] * g[i];
        proj /= lam;
        for (int i = 0; i < 6; ++i) x[i] += V[6 * i + k] * proj;
    }
    return true;
}

// Fast path of the 6x6 solve: Cholesky in registers (fully unrolled, static indices).  Taken only when every
// pivot is far above the rank threshold Eigen's JacobiSVD::solve applies (sigma_i > 6 eps sigma_max), i.e. when the
// pseudo-inverse is the inverse; otherwise (near-singular or non-finite H) the Jacobi path below decides, so the
// degenerate cases keep the reference's semantics.  ~150 flops instead of ~5000 on local-memory arrays.
__device__ __forceinline__ bool chol6_solve(const double* Hu, const double* g, double* x) {
    double L[6][6];
    double dmax = 0;
    {
        int k = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = i; j < 6; ++j) { L[j][i] = Hu[k++]; }   // lower triangle = H
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) dmax = fmax(dmax, L[i][i]);
    if (!(dmax > 0) || !isfinite(dmax)) return false;
    const double tiny = dmax * 1e-9;
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double d = L[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
        ok = ok && (d > tiny);
        const double r = sqrt(d), inv = 1.0 / r;
        L[j][j] = r;
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
            double v = L[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k];
            L[i][j] = v * inv;
        }
    }
    if (!ok) return false;
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double v = g[i];
#pragma unroll
        for (int k = 0; k < i; ++k) v -= L[i][k] * y[k];
        y[i] = v / L[i][i];
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        double v = y[i];
#pragma unroll
        for (int k = i + 1; k < 6; ++k) v -= L[k][i] * x[k];
        x[i] = v / L[i][i];
    }
    bool fin = true;
#pragma unroll
    for (int i = 0; i < 6; ++i) fin = fin && isfinite(x[i]);
    return fin;
}

__device__ __noinline__ void rodrigues(const double* w, double R[3][3]) {
    const double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    double a, b;
    if (th < 1e-12) { a = 1.0; b = 0.5; }
    else { a = sin(th) / th; b = (1 - cos(th)) / (th * th); }
    const double K[3][3] = {{0, -w[2], w[1]}, {w[2], 0, -w[0]}, {-w[1], w[0], 0}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double k2 = 0;
            for (int k = 0; k < 3; ++k) k2 += K[i][k] * K[k][j];
            R[i][j] = (i == j ? 1.0 : 0.0) + a * K[i][j] + b * k2;
        }
}

struct GNShared {
    RotDeriv R;
    double acc[GN_THREADS / 32][NACC];
    double pose[6];
    double step[6];
    double cCost, pCost;
    int done, abort, fail;
#ifdef SVO_GN_DEBUG
    int dbg_once; Result* dbg_res;
#endif
    unsigned timesInc;
    int n_non_masked;
};

// one m_evalRGN (S5:275-390); returns through sh: step, cCost, fail
__device__ void eval_rgn(const GnScratch& G, size_t base, int T, const Camera& cam, const Params& p, GNShared& sh) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) rot_and_derivs(sh.pose, sh.R);
    __syncthreads();
    const RotDeriv& R = sh.R;
    double dp[6];
    for (int i = 0; i < 6; ++i) dp[i] = sh.pose[i];
    const double b2 = p.use_robust_kernel ? p.kernel_param * p.kernel_param : 0;
    const double b2_1 = p.use_robust_kernel ? 1. / b2 : 0;
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0;
    for (int m = tid; m < T; m += GN_THREADS) {
        if (!G.mask[base + m]) continue;
        const double X1p = G.lmk[(base + m) * 3 + 0], Y1p = G.lmk[(base + m) * 3 + 1], Z1p = G.lmk[(base + m) * 3 + 2];
        const double X1c = R.r[0][0] * X1p + R.r[0][1] * Y1p + R.r[0][2] * Z1p + dp[3];
        const double Y1c = R.r[1][0] * X1p + R.r[1][1] * Y1p + R.r[1][2] * Z1p + dp[4];
        const double Z1c = R.r[2][0] * X1p + R.r[2][1] * Y1p + R.r[2][2] * Z1p + dp[5];
        const double X2c = X1c - cam.baseline;
        const float px0 = (float)(cam.fxl * X1c / Z1c + cam.cxl);
        const float px1 = (float)(cam.fyl * Y1c / Z1c + cam.cyl);
        const float px2 = (float)(cam.fxr * X2c / Z1c + cam.cxr);
        const float px3 = (float)(cam.fyr * Y1c / Z1c + cam.cyr);
        double J[4][6];
        bool good = true;
        const double iz2 = 1.0 / (Z1c * Z1c);
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            double X1cd, Y1cd, Z1cd;
            if (j < 3) {
                if (R.small) {
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
            // the reference divides each of the 24 entries by Z1c^2 (S5:251-254); one reciprocal and 24 products differ
            // from that by at most 2 ulp per entry (observed pose differences vs the oracle stay below 1e-12), and
            // FP64 division was 40 % of this kernel
            J[0][j] = cam.fxl * (X1cd * Z1c - X1c * Z1cd) * iz2;
            J[1][j] = cam.fyl * (Y1cd * Z1c - Y1c * Z1cd) * iz2;
            J[2][j] = cam.fxr * (X1cd * Z1c - X2c * Z1cd) * iz2;
            J[3][j] = cam.fyr * (Y1cd * Z1c - Y1c * Z1cd) * iz2;
            good = good && isfinite(J[0][j]) && isfinite(J[1][j]) && isfinite(J[2][j]) && isfinite(J[3][j]);
        }
        if (!good) continue;   // m_jacobian_is_good, S5:322
        const float4 ob = G.obs[base + m];
        const double r0 = (double)(ob.x - px0), r1 = (double)(ob.y - px1);   // float32 subtraction, S5:335-338
        const double r2 = (double)(ob.z - px2), r3 = (double)(ob.w - px3);
        const double s = r0 * r0 + r1 * r1 + r2 * r2 + r3 * r3;
        G.residual[base + m] = s;
        double rho_p = 1, fi;
        if (p.use_robust_kernel) {
            const double n = sqrt(1 + (s * b2_1));
            rho_p = 1 / n;
            fi = b2 * (n - 1);
        } else {
            fi = 0.5 * s;
        }
        acc[27] += fi;
        int k = 0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            acc[21 + a] += rho_p * (J[0][a] * r0 + J[1][a] * r1 + J[2][a] * r2 + J[3][a] * r3);
#pragma unroll
            for (int b = a; b < 6; ++b)
                acc[k++] += J[0][a] * J[0][b] + J[1][a] * J[1][b] + J[2][a] * J[2][b] + J[3][a] * J[3][b];
        }
    }
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        double v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sh.acc[wid][i] = v;
    }
    __syncthreads();
    if (tid == 0) {
        double tot[NACC];
        for (int i = 0; i < NACC; ++i) {
            double v = 0;
            for (int w = 0; w < GN_THREADS / 32; ++w) v += sh.acc[w][i];
            tot[i] = v;
        }
        sh.cCost = tot[27];
        double stepv[6];
        if (chol6_solve(tot, tot + 21, stepv)) {
            for (int i = 0; i < 6; ++i) sh.step[i] = stepv[i];
            sh.fail = 0;
        } else {
            sh.fail = solve6(tot, tot + 21, sh.step) ? 0 : 1;
        }
#ifdef SVO_GN_DEBUG
        if (sh.dbg_once == 0) {
            sh.dbg_once = 1;
            Result& rr = *sh.dbg_res;
            for (int i = 0; i < 16; ++i) rr.out_pose[i] = tot[i];
            for (int i = 0; i < 3; ++i) { rr.out_pose_rotvec[i] = tot[16 + i]; rr.out_pose_t[i] = tot[19 + i]; rr.out_pose_ypr[i] = tot[22 + i]; }
        }
#endif
    }
    __syncthreads();
}

// S5:528-544 / S5:623-638; returns the number of survivors
__device__ int triangulate(const GnScratch& G, size_t base, int T, const Camera& cam, int* s_cnt) {
    int cnt = 0;
    for (int m = threadIdx.x; m < T; m += GN_THREADS) {
        if (!G.mask[base + m]) continue;
        ++cnt;
        const float4 pp = G.prev[base + m];   // l1l.x, l1l.y, l1r.x, response
        const double ul = pp.x, vl = pp.y, ur = pp.z;
        const double b_d = cam.baseline / (cam.fxl * (cam.cxr - ur) + cam.fxr * (ul - cam.cxl));
        G.lmk[(base + m) * 3 + 0] = b_d * cam.fxr * (ul - cam.cxl);
        G.lmk[(base + m) * 3 + 1] = b_d * cam.fxr * (vl - cam.cyl);
        G.lmk[(base + m) * 3 + 2] = b_d * cam.fxl * cam.fxr;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
    __syncthreads();
    int tot = 0;
    for (int w = 0; w < GN_THREADS / 32; ++w) tot += s_cnt[w];
    __syncthreads();
    return tot;
}

__global__ void __launch_bounds__(GN_THREADS)
gn_kernel(Pools P, GnScratch G, int external, int use_custom_pose, int assume_fast_kps) {
    extern __shared__ __align__(16) unsigned char s_dyn[];   // [cap] uint32 index + [cap] float key (aux NMS)
    __shared__ GNShared sh;
    __shared__ int s_cnt[GN_THREADS / 32];
    const int slot = P.slots[blockIdx.x];
    SlotState& st = P.state[slot];
    Result& res = P.result[slot];
    const Params& p = P.p;
    const int tid = threadIdx.x;
    const size_t cap = (size_t)p.kp_cap;
    const size_t base = (size_t)slot * cap;
    const Camera cam = P.cam[slot];

    if (external != GN_EXTERNAL) {
        if (!st.has_prev) return;               // voecFirstIteration already reported (P:348-352)
        if (st.m_error == 5) {                  // voecBadTracking: stage 5 is skipped (P:332) ...
            if (tid == 0) res.tracked_frame = res.tracked_kf = 0;   // ... and with it S5:724-725 (see the end of this kernel)
            return;
        }
    }
    const int T = P.n_tracked[slot];
    const int cur = st.cur, prev = cur ^ 1;

    // ---- S5:419-461: gather the four point lists ----------------------------------------------------
    if (external == GN_FROM_TRACKED) {
        const int4* mp = P.match + pair_idx(slot, prev) * cap;
        const int4* mc = P.match + pair_idx(slot, cur) * cap;
        const size_t fpl = frame_idx(slot, prev, 0) * cap, fpr = frame_idx(slot, prev, 1) * cap;
        const size_t fcl = frame_idx(slot, cur, 0) * cap, fcr = frame_idx(slot, cur, 1) * cap;
        for (int i = tid; i < T; i += GN_THREADS) {
            const int2 tp = P.tracked[base + i];
            const int4 a = mp[tp.x], b = mc[tp.y];
            const float2 pl = kp_xy(P, fpl, a.x), pr = kp_xy(P, fpr, a.y), cl = kp_xy(P, fcl, b.x), cr = kp_xy(P, fcr, b.y);
            const float resp = P.kpf ? P.kpf[fpl + a.x].z : (float)P.kp[fpl + a.x].y;     // KeyPoint::response of the previous-left point
            G.prev[base + i] = make_float4(pl.x, pl.y, pr.x, resp);
            G.obs[base + i] = make_float4(cl.x, cl.y, cr.x, cr.y);
        }
    }
    for (int i = tid; i < T; i += GN_THREADS) G.residual[base + i] = DBL_MAX;   // S5:296
    __syncthreads();

    // ---- S5:465-474: aux non-max-suppression on the previous-left points ----------------------------
    const unsigned cell = (unsigned)(p.min_distance / 2.0);
    if (assume_fast_kps == GN_MASK_READY) {
        // G.mask was written by aux_nms_kernel (launch_gn runs it first for lists that do not come from stage 2)
    } else if (assume_fast_kps && cell == 1) {
        // keypoints produced by stage 2 are distinct and never 8-adjacent: with 1-px cells all survive
        for (int i = tid; i < T; i += GN_THREADS) G.mask[base + i] = 1;
    } else {
        uint32_t* s_idx = reinterpret_cast<uint32_t*>(s_dyn);
        float* s_key = reinterpret_cast<float*>(s_dyn) + cap;
        for (int i = tid; i < T; i += GN_THREADS) { s_idx[i] = i; s_key[i] = G.prev[base + i].w; G.mask[base + i] = 0; }
        __syncthreads();
        if (tid == 0) {
            IndexedDescComp c{s_key};
            if (p.sort_mode == 0) introsort_seq_c(c, s_idx, T);
            else insertion_sort_(c, s_idx, 0, T);
            const float cell_inv = 1.0f / (float)cell;
            const unsigned glx = (unsigned)(1 + p.W * cell_inv), gly = (unsigned)(1 + p.H * cell_inv);
            uint32_t* occ = G.occ + (size_t)slot * G.occ_words;
            for (unsigned i = 0; i < (glx * gly + 31) / 32; ++i) occ[i] = 0;
            auto get = [&](unsigned x, unsigned y) { const unsigned b = x * gly + y; return (occ[b >> 5] >> (b & 31)) & 1u; };
            auto set = [&](unsigned x, unsigned y) { const unsigned b = x * gly + y; occ[b >> 5] |= 1u << (b & 31); };
            int c_feats = 0, k = 0;
            while (c_feats < T && k < T) {
                const uint32_t idx = s_idx[k++];
                const float4 pp = G.prev[base + idx];
                const unsigned sx = (unsigned)(pp.x * cell_inv), sy = (unsigned)(pp.y * cell_inv);
                if (sx >= glx || sy >= gly) continue;   // outside the image: the reference would index out of bounds
                if (get(sx, sy)) continue;
                set(sx, sy);
                if (sx > 0) set(sx - 1, sy);
                if (sy > 0) set(sx, sy - 1);
                if (sx < glx - 1) set(sx + 1, sy);
                if (sy < gly - 1) set(sx, sy + 1);
                G.mask[base + idx] = 1;
                ++c_feats;
            }
        }
    }
    __syncthreads();
    // the survivors as stage 5 starts with them (what the reference lists in optimization.txt, S5:477-492); G.mask
    // itself loses the outliers of the first loop later (S5:601-605)
    for (int i = tid; i < T; i += GN_THREADS) G.mask0[base + i] = G.mask[base + i];

    // ---- initial pose (S5:500-507) ------------------------------------------------------------------
    if (tid == 0) {
        for (int i = 0; i < 6; ++i) {
            double v = 0;
            if (use_custom_pose) v = G.ini_pose[(size_t)slot * 6 + i];
            else if (p.use_prev_pose) v = st.last_pose[i];
            sh.pose[i] = v;
        }
        sh.cCost = 0; sh.pCost = 0; sh.done = 0; sh.abort = 0; sh.fail = 0; sh.timesInc = 0;
#ifdef SVO_GN_DEBUG
        sh.dbg_once = 0; sh.dbg_res = &res;
#endif
        res.n_outliers = 0;
        res.n_residuals = 0;
    }
    __syncthreads();

    int n_non_masked = triangulate(G, base, T, cam, s_cnt);
    // result.tracked_feats_from_last_frame / _KF are assigned at S5:724-725 only, i.e. when stage 5 runs to its end;
    // every early return leaves the caller's fresh TStereoOdometryResult at 0 (svo_get_octave_stats keeps the counts)
    if (n_non_masked < 8) {   // S5:521-526
        if (tid == 0) { res.valid = 0; res.n_residuals = 0; if (external != GN_EXTERNAL) res.tracked_frame = res.tracked_kf = 0; }
        return;
    }

    // ---- first loop (S5:549-598) --------------------------------------------------------------------
    int num_it = 0;
    while (num_it < p.initial_max_iters && !sh.done && !sh.abort) {
        if (tid == 0) sh.pCost = sh.cCost;
        __syncthreads();
        eval_rgn(G, base, T, cam, p, sh);
        if (tid == 0) res.error_code = sh.fail ? 1 : 0;   // S5:564: result.error_code = out_error_code
        if (sh.fail) {
            if (tid == 0) { st.m_error = 1; res.valid = 0; res.num_it = num_it; if (external != GN_EXTERNAL) res.tracked_frame = res.tracked_kf = 0; }   // S5:384, S5:569-573
            return;
        }
        if (tid == 0) {
            double m = 0;
            for (int k = 0; k < 6; ++k) { sh.pose[k] += sh.step[k]; m += sh.step[k] * sh.step[k]; res.delta_pose[k] = sh.pose[k]; }
            if (num_it > 0) {
                sh.done = sqrt(m) < p.min_mod_out_vector;
                if (sh.pCost < sh.cCost) {
                    if (++sh.timesInc > (unsigned)p.max_incr_cost) { res.error_code = 2; sh.abort = 1; }
                }
            }
        }
        ++num_it;
        __syncthreads();
    }
    if (tid == 0) res.num_it = num_it;

    // ---- S5:601-611: inliers; result.outliers receives their current-match indices ------------------
    {
        int basec = 0;
        int32_t* outl = P.outliers + base;
        for (int i0 = 0; i0 < T; i0 += GN_THREADS) {
            const int i = i0 + tid;
            int in = 0;
            if (i < T) {
                if (G.residual[base + i] > p.residual_threshold) G.mask[base + i] = 0;
                else in = 1;
            }
            // block exclusive scan of `in`
            const int lane = tid & 31, wid = tid >> 5;
            const unsigned bal = __ballot_sync(0xffffffffu, in);
            if (lane == 0) s_cnt[wid] = __popc(bal);
            __syncthreads();
            int woff = 0, tot = 0;
            for (int w = 0; w < GN_THREADS / 32; ++w) { if (w < wid) woff += s_cnt[w]; tot += s_cnt[w]; }
            if (in) outl[basec + woff + __popc(bal & ((1u << lane) - 1))] = external ? G.cur_idx[base + i] : P.tracked[base + i].y;
            basec += tot;
            __syncthreads();
        }
        if (tid == 0) res.n_outliers = basec;
    }
    __syncthreads();
    n_non_masked = triangulate(G, base, T, cam, s_cnt);
    if (n_non_masked < 8) {   // S5:616-621
        if (tid == 0) { res.valid = 0; res.n_residuals = T; if (external != GN_EXTERNAL) res.tracked_frame = res.tracked_kf = 0; }
        return;
    }

    // ---- refinement loop (S5:650-700) ----------------------------------------------------------------
    if (tid == 0) { sh.done = 0; sh.abort = 0; }
    __syncthreads();
    int num_it_final = 0;
    while (num_it_final < p.max_iters && !sh.done && !sh.abort) {
        if (tid == 0) sh.pCost = sh.cCost;
        __syncthreads();
        eval_rgn(G, base, T, cam, p, sh);
        if (sh.fail) {   // S5:670-675 (result.error_code is not touched here; m_error is, S5:384)
            if (tid == 0) { st.m_error = 1; res.valid = 0; res.n_residuals = T; res.num_it_final = num_it_final; if (external != GN_EXTERNAL) res.tracked_frame = res.tracked_kf = 0; }
            return;
        }
        if (tid == 0) {
            double m = 0;
            for (int k = 0; k < 6; ++k) { sh.pose[k] += sh.step[k]; m += sh.step[k] * sh.step[k]; res.delta_pose[k] = sh.pose[k]; }
            if (num_it_final > 0) {
                sh.done = sqrt(m) < p.min_mod_out_vector;
                if (sh.pCost < sh.cCost) {
                    if (++sh.timesInc > (unsigned)p.max_incr_cost) { sh.abort = 1; res.error_code = 3; }
                }
            }
        }
        ++num_it_final;
        __syncthreads();
    }

    // ---- S5:715-727: outPose = CPose3D( CPose3DRotVec(deltaPose).getInverse() ) ----------------------
    if (tid == 0) {
        res.num_it_final = num_it_final;
        double dp[6], Rm[3][3], Ri[3][3];
        for (int k = 0; k < 6; ++k) { dp[k] = sh.pose[k]; res.delta_pose[k] = dp[k]; }
        rodrigues(dp, Rm);
        for (int i = 0; i < 3; ++i) {
            res.out_pose_rotvec[i] = -dp[i];
            res.out_pose_t[i] = -(Rm[0][i] * dp[3] + Rm[1][i] * dp[4] + Rm[2][i] * dp[5]);
        }
        rodrigues(res.out_pose_rotvec, Ri);
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) res.out_pose[4 * i + j] = Ri[i][j];
            res.out_pose[4 * i + 3] = res.out_pose_t[i];
        }
        res.out_pose[12] = res.out_pose[13] = res.out_pose[14] = 0;
        res.out_pose[15] = 1;
        res.out_pose_ypr[1] = atan2(-Ri[2][0], hypot(Ri[0][0], Ri[1][0]));
        res.out_pose_ypr[0] = atan2(Ri[1][0], Ri[0][0]);
        res.out_pose_ypr[2] = atan2(Ri[2][1], Ri[2][2]);
        if (!use_custom_pose && p.use_prev_pose)
            for (int k = 0; k < 6; ++k) st.last_pose[k] = dp[k];   // S5:720-721
        res.n_residuals = T;
        res.valid = sh.abort ? 0 : 1;   // S5:727
    }
}

// ------------------------------------------------------------------------------------------------
// aux_nms_kernel: the auxiliary non-maximal suppression of stage 5 (S5:465-474 -> m_non_max_sup, S2:225-283) for point
// lists that are not stage 2's own keypoints: the flattened octaves (rescaled points of different octaves collide) and
// getChangeInPose's caller data.  One CTA per slot: the indices are ordered by response with the block-cooperative
// emulation of std::sort (introsort_block: same permutation as the reference's std::sort, ties included), the occupancy
// grid lives in shared memory when it fits, one thread then walks the sorted list (the greedy rule is sequential).
// gn_kernel keeps a one-thread version of the same steps for the remaining case (stage-2 keypoints with cells > 1 px).
// ------------------------------------------------------------------------------------------------
constexpr int AUX_THREADS = 128;
struct AuxLayout { int tmp, lists, pos, occ, words; };      // offsets in 32-bit words of the dynamic shared memory
static __host__ __device__ AuxLayout aux_layout(int cap, int occ_words_smem) {
    AuxLayout L;
    L.tmp = 2 * cap;                                   // [cap] idx | [cap] key | [cap] tmp
    L.lists = 3 * cap + (cap & 1);                     // 8-byte aligned: 3 lists of (cap/16 + 2) uint2
    L.pos = L.lists + 6 * (cap / 16 + 2);              // 64 ints per warp
    L.occ = L.pos + 64 * (AUX_THREADS / 32);
    L.words = L.occ + occ_words_smem;
    return L;
}

__global__ void __launch_bounds__(AUX_THREADS)
aux_nms_kernel(Pools P, GnScratch G, int occ_in_smem) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ int s_sortcnt[3];
    const int slot = P.slots[blockIdx.x];
    const Params& p = P.p;
    const int tid = threadIdx.x;
    const int cap = p.kp_cap;
    const size_t base = (size_t)slot * cap;
    const int T = min(P.n_tracked[slot], cap);
    const unsigned cell = (unsigned)(p.min_distance / 2.0);
    const float cell_inv = 1.0f / (float)cell;
    const unsigned glx = (unsigned)(1 + p.W * cell_inv), gly = (unsigned)(1 + p.H * cell_inv);
    const unsigned occ_n = (glx * gly + 31) / 32;
    const AuxLayout L = aux_layout(cap, 0);
    uint32_t* s_idx = reinterpret_cast<uint32_t*>(s_dyn);
    float* s_key = reinterpret_cast<float*>(s_dyn) + cap;
    uint32_t* s_tmp = s_idx + L.tmp;
    uint2* s_list = reinterpret_cast<uint2*>(s_idx + L.lists);
    int* s_pos = reinterpret_cast<int*>(s_idx + L.pos);
    uint32_t* occ = occ_in_smem ? s_idx + L.occ : G.occ + (size_t)slot * G.occ_words;
    for (int i = tid; i < T; i += AUX_THREADS) { s_idx[i] = i; s_key[i] = G.prev[base + i].w; G.mask[base + i] = 0; }
    for (unsigned i = tid; i < occ_n; i += AUX_THREADS) occ[i] = 0;
    __syncthreads();
    IndexedDescComp c{s_key};
    if (p.sort_mode == 0) {
        const int nl = cap / 16 + 2;
        introsort_block(c, s_idx, T, s_list, s_list + nl, s_list + 2 * nl, s_sortcnt, s_tmp, s_pos, T);
    } else {
        if (tid == 0) insertion_sort_(c, s_idx, 0, T);
        __syncthreads();
    }
    if (tid == 0) {
        auto get = [&](unsigned x, unsigned y) { const unsigned b = x * gly + y; return (occ[b >> 5] >> (b & 31)) & 1u; };
        auto set = [&](unsigned x, unsigned y) { const unsigned b = x * gly + y; occ[b >> 5] |= 1u << (b & 31); };
        for (int k = 0; k < T; ++k) {                  // num_out == T: the quota never ends the walk early (S5:468-473)
            const uint32_t idx = s_idx[k];
            const float4 pp = G.prev[base + idx];
            const unsigned sx = (unsigned)(pp.x * cell_inv), sy = (unsigned)(pp.y * cell_inv);
            if (sx >= glx || sy >= gly) continue;      // outside the image: the reference would index out of bounds
            if (get(sx, sy)) continue;
            set(sx, sy);
            if (sx > 0) set(sx - 1, sy);
            if (sy > 0) set(sx, sy - 1);
            if (sx < glx - 1) set(sx + 1, sy);
            if (sy < gly - 1) set(sx, sy + 1);
            G.mask[base + idx] = 1;
        }
    }
}

// dynamic shared memory of aux_nms_kernel; *occ_in_smem = whether the occupancy grid fits beside the sort scratch
static int aux_smem_bytes(const Params& p, int* occ_in_smem) {
    const unsigned cell = (unsigned)(p.min_distance / 2.0);
    const float cell_inv = 1.0f / (float)cell;
    const unsigned glx = (unsigned)(1 + p.W * cell_inv), gly = (unsigned)(1 + p.H * cell_inv);
    const int occ_n = (int)((glx * gly + 31) / 32);
    const int without = aux_layout(p.kp_cap, 0).words * 4, with = aux_layout(p.kp_cap, occ_n).words * 4;
    *occ_in_smem = with <= 160 * 1024 ? 1 : 0;
    return *occ_in_smem ? with : without;
}
static bool aux_kernel_usable(const Params& p) { int o; return aux_smem_bytes(p, &o) <= 200 * 1024; }

cudaError_t configure_gn(const Params& p) {
    if (aux_kernel_usable(p)) {
        int o;
        cudaError_t e = SVO_RAISE_DYN_SMEM(aux_nms_kernel, aux_smem_bytes(p, &o));
        if (e != cudaSuccess) return e;
    }
    return SVO_RAISE_DYN_SMEM(gn_kernel, (int)(p.kp_cap * 8));
}

// getProjectedCoords (common.cpp:415-470): triangulate the stereo matches flagged "not tracked" (flag == -1) in the
// previous frame (S5:537-542 formula) and project them with the pose [w1 w2 w3 t1 t2 t3] (m_pinhole_stereo_projection,
// S5:165-193, pixels as float32 TPixelCoordf).  One thread per flagged match; `order` lists their match indices.
__global__ void project_kernel(const int4* __restrict__ matches, const float2* __restrict__ left, const float2* __restrict__ right,
                               const int32_t* __restrict__ order, int n, Camera cam, const double* __restrict__ pose,
                               float4* __restrict__ out) {
    __shared__ RotDeriv R;
    __shared__ double dp[6];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 6; ++i) dp[i] = pose[i];
        rot_and_derivs(dp, R);
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 m = matches[order[i]];
    const double ul = left[m.x].x, vl = left[m.x].y, ur = right[m.y].x;
    const double disparity = cam.fxl * (cam.cxr - ur) + cam.fxr * (ul - cam.cxl);
    const double b_d = cam.baseline / disparity;
    const double X1p = b_d * cam.fxr * (ul - cam.cxl), Y1p = b_d * cam.fxr * (vl - cam.cyl), Z1p = b_d * cam.fxl * cam.fxr;
    const double X1c = R.r[0][0] * X1p + R.r[0][1] * Y1p + R.r[0][2] * Z1p + dp[3];
    const double Y1c = R.r[1][0] * X1p + R.r[1][1] * Y1p + R.r[1][2] * Z1p + dp[4];
    const double Z1c = R.r[2][0] * X1p + R.r[2][1] * Y1p + R.r[2][2] * Z1p + dp[5];
    const double X2c = X1c - cam.baseline;
    out[i] = make_float4((float)(cam.fxl * X1c / Z1c + cam.cxl), (float)(cam.fyl * Y1c / Z1c + cam.cyl),
                         (float)(cam.fxr * X2c / Z1c + cam.cxr), (float)(cam.fyr * Y1c / Z1c + cam.cyr));
}

cudaError_t launch_project(const int4* d_matches, const float2* d_left, const float2* d_right, const int32_t* d_order, int n,
                           const Camera& cam, const double* d_pose, float4* d_out, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    project_kernel<<<(n + 127) / 128, 128, 0, st>>>(d_matches, d_left, d_right, d_order, n, cam, d_pose, d_out);
    return cudaGetLastError();
}

// S5:419-461 for nOctaves > 1: one CTA per slot walks the octaves in order and appends their tracked pairs, keypoint
// coordinates multiplied by scale_norm = 2^octave (float x size_t, exact for FAST's integer coordinates).
__global__ void __launch_bounds__(256)
flatten_octaves_kernel(Pools PG, GnScratch G, OctaveList O) {
    const int slot = PG.slots[blockIdx.x];
    const SlotState& st = PG.state[slot];
    if (!st.has_prev) { if (threadIdx.x == 0) PG.n_tracked[slot] = 0; return; }
    const int cur = st.cur, prev = cur ^ 1;
    const size_t base = (size_t)slot * PG.p.kp_cap;
    int off = 0;
    for (int k = 0; k < O.n; ++k) {
        const size_t cap = (size_t)O.cap[k];
        const int T = O.n_tracked[k][slot];
        const int2* __restrict__ tr = O.tracked[k] + (size_t)slot * cap;
        const int4* __restrict__ mp = O.match[k] + pair_idx(slot, prev) * cap;
        const int4* __restrict__ mc = O.match[k] + pair_idx(slot, cur) * cap;
        const int2* __restrict__ kpl = O.kp[k] + frame_idx(slot, prev, 0) * cap;
        const int2* __restrict__ kpr = O.kp[k] + frame_idx(slot, prev, 1) * cap;
        const int2* __restrict__ kcl = O.kp[k] + frame_idx(slot, cur, 0) * cap;
        const int2* __restrict__ kcr = O.kp[k] + frame_idx(slot, cur, 1) * cap;
        const float s = (float)(1u << k);
        for (int i = threadIdx.x; i < T; i += blockDim.x) {
            const int2 tp = tr[i];
            const int4 a = mp[tp.x], b = mc[tp.y];
            const int2 pl = kpl[a.x], pr = kpr[a.y], cl = kcl[b.x], cr = kcr[b.y];
            G.prev[base + off + i] = make_float4((float)(pl.x & 0xFFFF) * s, (float)((unsigned)pl.x >> 16) * s,
                                                 (float)(pr.x & 0xFFFF) * s, (float)pl.y);
            G.obs[base + off + i] = make_float4((float)(cl.x & 0xFFFF) * s, (float)((unsigned)cl.x >> 16) * s,
                                                (float)(cr.x & 0xFFFF) * s, (float)((unsigned)cr.x >> 16) * s);
            G.cur_idx[base + off + i] = tp.y;   // result.outliers receives tracked_pairs[octave][i].second (S5:601-610)
        }
        off += T;
    }
    if (threadIdx.x == 0) PG.n_tracked[slot] = off;
}

cudaError_t launch_flatten_octaves(const Pools& PG, const GnScratch& G, const OctaveList& O, int n, cudaStream_t st, int* launches) {
    flatten_octaves_kernel<<<n, 256, 0, st>>>(PG, G, O);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_gn(const Pools& P, const GnScratch& G, int n, int external, int use_custom_pose,
                      int assume_fast_kps, cudaStream_t st, int* launches, StepTimer* tm) {
    // point lists that are not stage 2's keypoints: the aux NMS runs in its own block-parallel kernel first
    if (external != GN_FROM_TRACKED && !assume_fast_kps && aux_kernel_usable(P.p)) {
        int occ_in_smem = 0;
        const int smem = aux_smem_bytes(P.p, &occ_in_smem);
        aux_nms_kernel<<<n, AUX_THREADS, smem, st>>>(P, G, occ_in_smem);
        if (launches) *launches += 1;
        assume_fast_kps = GN_MASK_READY;
    }
    gn_kernel<<<n, GN_THREADS, P.p.kp_cap * 8, st>>>(P, G, external, use_custom_pose, assume_fast_kps);
    if (tm) tm->mark(st);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace svo
