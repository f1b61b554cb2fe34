// ransac.cu -- the two cv::findFundamentalMat(p1, p2, FM_RANSAC, 1.0, 0.99, mask) calls of
// stage4_match_consecutive.cpp:171-241 on sm_100a, one CTA per (stereo pair, side).
//
// OpenCV's estimator is sequential (adaptive iteration count, one RNG stream), but iteration k's
// sample does not depend on the models of earlier iterations: thread 0 draws the next RS_BATCH samples
// from the exact cv::RNG stream (including the duplicate and collinearity re-draws), every warp solves
// one of them (its Jacobi rotations as a wavefront) and the CTA scores their models, and thread 0 then replays OpenCV's "best so far /
// update niters" logic in iteration order.  The result is the inlier mask OpenCV returns (see fundmat_core.cuh for the one
// floating-point caveat).  With >= 15 points this is RANSAC, with 8..14 LMedS, with 7 an all-ones mask
// and with < 7 no mask (zero inliers), as in cv::findFundamentalMat.
#include "fundmat_core.cuh"
#include "ransac.cuh"

namespace svo {

constexpr int RS_WARPS = 6;                      // warps per (stereo pair, side) problem
constexpr int RS_GROUPS = 1;                     // 7-point samples per warp (run_7point_wave: one, its row pairs side by side)
constexpr int RS_BATCH = RS_WARPS * RS_GROUPS;   // RANSAC iterations evaluated per round
constexpr int RS_THREADS = RS_WARPS * 32;

struct RsShared {
    CvRng rng;
    double F[RS_BATCH][27];
    double best[9];
    double minMedian;
    Pt2f ms1[RS_BATCH][7], ms2[RS_BATCH][7];
    int found[RS_BATCH];
    int nmodels[RS_BATCH];
    int good[RS_BATCH][3];
    float median[RS_BATCH][3];
    int iter_base, niters, maxGood, stop, have_model, n_scored;
    int wsum[RS_WARPS];
    CvRng rng_after[RS_BATCH];     // speculative sampling: the RNG state behind every sample of the round
    int idx[RS_BATCH][7];
    int collinear[RS_BATCH];
    double svd[RS_WARPS][96];       // run_7point_wave's rows and norms, per warp
};

// S4:181-232: point sets of the left-left and right-right problems from the greedy-filtered lists
__global__ void __launch_bounds__(256)
ransac_prepare_kernel(Pools P, RansacScratch R) {
    const int slot = P.slots[blockIdx.x];
    const SlotState& st = P.state[slot];
    const size_t cap = (size_t)P.p.kp_cap;
    const int n = st.has_prev ? P.n_tcand[slot] : 0;
    const int cur = st.cur, prev = cur ^ 1;
    const int4* __restrict__ tc = P.tcand + (size_t)slot * cap;
    const int4* __restrict__ mp = P.match + pair_idx(slot, prev) * cap;
    const int4* __restrict__ mc = P.match + pair_idx(slot, cur) * cap;
    for (int side = 0; side < 2; ++side) {
        const size_t fp = frame_idx(slot, prev, side) * cap, fc = frame_idx(slot, cur, side) * cap;
        float2* __restrict__ o1 = R.p1 + ((size_t)slot * 2 + side) * cap;
        float2* __restrict__ o2 = R.p2 + ((size_t)slot * 2 + side) * cap;
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            const int4 c = tc[k];
            const int4 a = mp[c.x], b = mc[side ? c.z : c.y];
            o1[k] = kp_xy(P, fp, side ? a.y : a.x);
            o2[k] = kp_xy(P, fc, side ? b.y : b.x);
        }
        if (threadIdx.x == 0) R.n_pts[slot * 2 + side] = n;
    }
}


// ((v0 + v1) + v2) + ... + v8 over the 9 lanes of this lane's group (base = first lane of the group), on every lane
// of the group: the left-to-right sum of OpenCV's scalar loops
__device__ __forceinline__ double ordered_sum9(double v, int base) {
    const unsigned full = 0xffffffffu;
    double s = __shfl_sync(full, v, base);
#pragma unroll
    for (int k = 1; k < 9; ++k) s += __shfl_sync(full, v, base + k);
    return s;
}

// run_7point with the SVD step (fm7_svd_rows: OpenCV's one-sided Jacobi on the rows of the 7x9 system, then the two
// extra right singular vectors) by ONE WARP PER SAMPLE, the rotations of a sweep as a wavefront.
// OpenCV walks the row pairs (i, j), i < j, in order; pair (i, j) only needs row i as (i, j-1) left it and row j as
// (i-1, j) left it, so all pairs with the same i + j are independent: 11 wavefront steps of up to three pairs replace the
// 21 sequential pair steps of a sweep (the chain of IEEE divisions and square roots per pair is what the kernel waits
// for).  Lanes 9p .. 9p+8 (p = 0..2) own the nine columns of the p-th pair of the step.  Every rotation touches its
// elements with exactly the operations of the sequential loops, every dot product is summed in the sequential order
// (ordered_sum9), a pair sees exactly the rows and norms the sequential order would hand it, and "changed" is an OR:
// f1 / f2 and the models are bit-identical to the host build of fundmat_core.cuh that is pinned against cv2.
// OpenCV's data-dependent branches ("rows already orthogonal: skip", "beta < 0", hypot's operand order) are selects.
// The rows live in shared memory (svd: 96 doubles of this warp: rows[9][9], then W[7]); loops stay loops and the function
// out of line: fully unrolled on register arrays it was 250 KB of code and the kernel waited for instructions.
// ms1 / ms2: the sample's 7 points, valid[0] != 1 marks an unused warp; F: 27 doubles, nm_out: the model count.
// All 32 lanes must call it.
__device__ __noinline__ void run_7point_wave(const Pt2f (*ms1)[7], const Pt2f (*ms2)[7], const int* valid, double (*F)[27],
                                             int* nm_out, double* svd) {
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    const int pg = lane < 27 ? lane / 9 : 2;           // pair of the wavefront step this lane works for
    const int base = 9 * pg;
    const int k = lane - base;                         // 0..8: a column; 9..13 (lanes 27..31): no column
    const bool colv = k < 9;
    const int kc = colv ? k : 0;
    const Pt2f* m1 = ms1[0];
    const Pt2f* m2 = ms2[0];
    Fm7Norm nm;
    bool ok = valid[0] == 1;
    if (ok) ok = fm7_normalise(m1, m2, nm);
    const int m = 9, n = 7;
    const double minval = DBL_MIN, eps = DBL_EPSILON * 10;
    double* const rows = svd;                          // rows[9 * i + k], i < 9
    double* const W = svd + 81;                        // W[i], i < 7
    // an unused / degenerate sample works on orthonormal dummy rows: no rotations, no zero singular values
    for (int i = pg; i < 7; i += 3)
        if (colv) rows[9 * i + k] = ok ? fm7_entry(m1, m2, nm, i, k) : (k == i ? 1.0 : 0.0);
    if (pg == 0 && colv) { rows[63 + k] = 0.0; rows[72 + k] = 0.0; }
    __syncwarp();
#pragma unroll 1
    for (int i0 = 0; i0 < 7; i0 += 3) {
        const int i = min(i0 + pg, 6);
        const double c = rows[9 * i + kc];
        const double w = ordered_sum9(c * c, base);
        if (k == 0 && i0 + pg < 7) W[i] = w;
    }
    __syncwarp();
#pragma unroll 1
    for (int iter = 0; iter < 30; ++iter) {
        bool changed = false;
#pragma unroll 1
        for (int t = 1; t <= 2 * n - 3; ++t) {         // pairs (i, t - i), max(0, t - 6) <= i <= (t - 1) / 2
            const int imin = t > n - 1 ? t - (n - 1) : 0, imax = (t - 1) / 2;
            const bool has = colv && imin + pg <= imax;
            const int i = has ? imin + pg : 0, j = has ? t - i : 1;
            const double ci = rows[9 * i + kc], cj = rows[9 * j + kc];
            const double a = W[i], b = W[j];
            double p = ordered_sum9(ci * cj, base);
            const bool rot = has && !(fabs(p) <= eps * sqrt(a * b));
            if (!__any_sync(full, rot)) continue;      // warp-uniform; nothing was written
            p *= 2;
            const double beta = a - b;
            // gamma = cv_hypot(p, beta)
            const double ap = fabs(p), ab = fabs(beta);
            const double hi = ap > ab ? ap : ab, lo = ap > ab ? ab : ap;
            const double q = lo / hi;
            const double gamma = hi > 0 ? hi * sqrt(1 + q * q) : 0.0;
            // beta < 0: s = sqrt(((gamma - beta)*0.5)/gamma), c = p/(gamma*s*2); else c = sqrt((gamma+beta)/(gamma*2)), s = p/(gamma*c*2)
            const bool neg = beta < 0;
            const double num = neg ? (gamma - beta) * 0.5 : (gamma + beta);
            const double den = neg ? gamma : gamma * 2;
            const double first = sqrt(num / den);
            const double second = p / (gamma * first * 2);
            const double c = neg ? second : first, s = neg ? first : second;
            const double t0 = c * ci + s * cj;
            const double t1 = -s * ci + c * cj;
            const double wa = ordered_sum9(t0 * t0, base), wb = ordered_sum9(t1 * t1, base);
            __syncwarp();                              // every lane holds its operands: the pairs of a step share no row
            if (rot) {
                rows[9 * i + kc] = t0; rows[9 * j + kc] = t1;
                if (k == 0) { W[i] = wa; W[j] = wb; }
                changed = true;
            }
            __syncwarp();
        }
        if (!__any_sync(full, changed)) break;
    }
#pragma unroll 1
    for (int i0 = 0; i0 < 7; i0 += 3) {
        const int i = min(i0 + pg, 6);
        const double c = rows[9 * i + kc];
        const double w = sqrt(ordered_sum9(c * c, base));
        __syncwarp();
        if (k == 0 && i0 + pg < 7) W[i] = w;
    }
    __syncwarp();
#pragma unroll 1
    for (int i = 0; i < n - 1; ++i) {                  // selection sort, descending; first maximum wins (uniform over the warp)
        int j = i;
        double wj = W[i];
        for (int kk = i + 1; kk < n; ++kk) if (wj < W[kk]) { j = kk; wj = W[kk]; }
        __syncwarp();
        if (j != i) {
            if (lane == 0) { const double tw = W[i]; W[i] = W[j]; W[j] = tw; }
            if (pg == 0 && colv) { const double tc = rows[9 * i + k]; rows[9 * i + k] = rows[9 * j + k]; rows[9 * j + k] = tc; }
        }
        __syncwarp();
    }
    // the rows of the null space: every lane group repeats the same computation on the same rows (all 32 lanes take part
    // in the shuffles), group 0 writes
    CvRng rng;
    rng.state = 0x12345678;
#pragma unroll 1
    for (int i = 0; i < 9; ++i) {
        double sd = i < n ? W[i] : 0;
        double ci_cur = rows[9 * i + kc];
#pragma unroll 1
        for (int ii = 0; ii < 100; ++ii) {
            const bool need = sd <= minval;            // draw another vector
            if (!need) break;                          // uniform: sd is the same on every lane
            const double val0 = 1. / m;
            double v = 0;
            for (int kk = 0; kk < m; ++kk) {
                const double vk = (rng.next() & 256) != 0 ? val0 : -val0;
                if (kk == kc) v = vk;
            }
            double ci = v;
#pragma unroll 1
            for (int it2 = 0; it2 < 2; ++it2) {
#pragma unroll 1
                for (int j = 0; j < i; ++j) {
                    const double cj = rows[9 * j + kc];
                    const double d = ordered_sum9(ci * cj, base);
                    const double tt = ci - d * cj;
                    double asum = ordered_sum9(fabs(tt), base);
                    asum = asum > eps * 100 ? 1 / asum : 0;
                    ci = tt * asum;
                }
            }
            sd = sqrt(ordered_sum9(ci * ci, base));
            ci_cur = ci;
        }
        // lanes without a column (27..31) ran on group 2's sums: sd and the loop trip count are the same everywhere
        const double sc = sd > minval ? 1 / sd : 0.;
        __syncwarp();
        if (pg == 0 && colv) rows[9 * i + k] = ci_cur * sc;
        __syncwarp();
    }
    if (lane == 0) {
        double f1[9], f2[9];
        for (int j = 0; j < 9; ++j) { f1[j] = rows[63 + j]; f2[j] = rows[72 + j]; }
        nm_out[0] = ok ? fm7_finish(f1, f2, nm, F[0]) : 0;
    }
}

__global__ void __launch_bounds__(RS_THREADS)
fundmat_kernel(RansacScratch R) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ RsShared sh;
    const int item = blockIdx.x >> 1, side = blockIdx.x & 1;
    const int slot = R.slots[item];
    const size_t prob = (size_t)slot * 2 + side;
    const int n = R.n_pts[prob];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    Pt2f* m1 = reinterpret_cast<Pt2f*>(s_dyn);
    Pt2f* m2 = m1 + R.kp_cap;
    uint8_t* __restrict__ mask = R.inl + prob * R.kp_cap;
    const float2* __restrict__ g1 = R.p1 + prob * R.kp_cap;
    const float2* __restrict__ g2 = R.p2 + prob * R.kp_cap;

    if (n < 7) {                       // findFundamentalMat returns an empty Mat and no mask
        if (tid == 0) { R.n_inl[prob] = 0; R.n_iter[prob] = 0; }
        return;
    }
    if (n == 7) {                      // mask.setTo(1)
        if (tid < 7) mask[tid] = 1;
        if (tid == 0) { R.n_inl[prob] = 7; R.n_iter[prob] = 0; }
        return;
    }
    for (int i = tid; i < n; i += RS_THREADS) {
        const float2 a = g1[i], b = g2[i];
        m1[i].x = a.x; m1[i].y = a.y; m2[i].x = b.x; m2[i].y = b.y;
    }
    const bool lmeds = n < 15;
    if (tid == 0) {
        sh.rng.seed();
        sh.iter_base = 0;
        sh.maxGood = 0;
        sh.stop = 0;
        sh.have_model = 0;
        sh.n_scored = 0;
        sh.minMedian = DBL_MAX;
        if (lmeds) { const int it = ransac_update_num_iters(kFmConfidence, 0.45, kFmModelPoints, kFmMaxIters); sh.niters = it > 3 ? it : 3; }
        else sh.niters = kFmMaxIters;
    }
    __syncthreads();
    const float thr = (float)(kFmThreshold * kFmThreshold);

    while (!sh.stop) {
        // ---- the next RS_BATCH samples from the cv::RNG stream.  Drawing is sequential (one stream, re-draws on duplicate
        //      indices and on collinear samples), but the collinearity test -- the expensive part -- only rarely fails: thread
        //      0 draws all samples assuming they pass, RS_BATCH threads test them in parallel, and from the first failing
        //      sample on thread 0 falls back to the exact sequential getSubset ------------------------------------------------
        if (tid == 0) {
            for (int w = 0; w < RS_BATCH; ++w) sh.found[w] = -1;
            CvRng rg = sh.rng;
            for (int w = 0; w < RS_BATCH; ++w) {
                if (sh.iter_base + w >= sh.niters) break;
                for (int i = 0; i < 7; ++i) {
                    int idx_i;
                    while (true) {
                        idx_i = rg.uniform(0, n);
                        bool dup = false;
                        for (int j = 0; j < i; ++j) dup = dup || (sh.idx[w][j] == idx_i);
                        if (!dup) break;
                    }
                    sh.idx[w][i] = idx_i;
                    sh.ms1[w][i] = m1[idx_i];
                    sh.ms2[w][i] = m2[idx_i];
                }
                sh.rng_after[w] = rg;
                sh.found[w] = 1;
            }
        }
        __syncthreads();
        if (tid < RS_BATCH) sh.collinear[tid] = sh.found[tid] == 1 && (last_point_collinear(sh.ms1[tid], 7) || last_point_collinear(sh.ms2[tid], 7));
        __syncthreads();
        if (tid == 0) {
            int bad = -1;
            for (int w = 0; w < RS_BATCH && bad < 0; ++w) if (sh.found[w] == 1 && sh.collinear[w]) bad = w;
            if (bad < 0) {
                for (int w = RS_BATCH - 1; w >= 0; --w) if (sh.found[w] == 1) { sh.rng = sh.rng_after[w]; break; }
            } else {                                   // exact sequential path from the first collinear sample on
                sh.rng = bad == 0 ? sh.rng : sh.rng_after[bad - 1];
                for (int w = bad; w < RS_BATCH; ++w) {
                    sh.found[w] = -1;
                    if (sh.iter_base + w >= sh.niters) continue;
                    int idx[7];
                    const bool ok = get_subset7(m1, m2, n, sh.rng, lmeds ? 1000 : 10000, idx, sh.ms1[w], sh.ms2[w]);
                    sh.found[w] = ok ? 1 : 0;
                    if (!ok) { for (int v = w + 1; v < RS_BATCH; ++v) sh.found[v] = -1; break; }
                }
            }
        }
        __syncthreads();
        // ---- every warp: solve its sample -----------------------------------------------------------------
        const int w0 = wid * RS_GROUPS;
        if (sh.found[w0] == 1)                         // samples are drawn in order: none of the later ones without the first
            run_7point_wave(sh.ms1 + w0, sh.ms2 + w0, sh.found + w0, sh.F + w0, sh.nmodels + w0, sh.svd[wid]);
        __syncthreads();
        // ---- OpenCV's sequential loop over the iterations of this round: the whole CTA scores one sample's models,
        //      thread 0 does the "best so far / update niters" bookkeeping; samples beyond the (shrinking) iteration
        //      count are never scored ---------------------------------------------------------------------------
        for (int w = 0; w < RS_BATCH; ++w) {
            const int iter = sh.iter_base + w;
            if (iter >= sh.niters || sh.found[w] != 1) {       // uniform: shared state, read after a barrier
                if (tid == 0) sh.stop = 1;                     // getSubset failed: return false (iter 0) or break
                break;
            }
            const int nm = sh.nmodels[w];
            if (tid == 0) ++sh.n_scored;                       // an iteration OpenCV's loop executes
            for (int i = 0; i < nm; ++i) {
                const double* F = sh.F[w] + 9 * i;
                if (!lmeds) {
                    int good = 0;
                    for (int k = tid; k < n; k += RS_THREADS) good += fm_error(F, m1[k], m2[k]) <= thr ? 1 : 0;
                    for (int o = 16; o > 0; o >>= 1) good += __shfl_xor_sync(0xffffffffu, good, o);
                    if (lane == 0) sh.wsum[wid] = good;
                } else if (tid == 0) {
                    float e[14];
                    for (int k = 0; k < n; ++k) {       // n <= 14: insertion sort, element n/2 is nth_element's answer
                        const float v = fm_error(F, m1[k], m2[k]);
                        int j = k;
                        while (j > 0 && e[j - 1] > v) { e[j] = e[j - 1]; --j; }
                        e[j] = v;
                    }
                    sh.median[w][i] = e[n / 2];
                }
                __syncthreads();
                if (tid == 0) {
                    if (!lmeds) {
                        int good = 0;
                        for (int q = 0; q < RS_WARPS; ++q) good += sh.wsum[q];
                        if (good > max(sh.maxGood, kFmModelPoints - 1)) {
                            sh.maxGood = good;
                            for (int q = 0; q < 9; ++q) sh.best[q] = F[q];
                            sh.have_model = 1;
                            sh.niters = ransac_update_num_iters(kFmConfidence, (double)(n - good) / n, kFmModelPoints, sh.niters);
                        }
                    } else {
                        const double med = (double)sh.median[w][i];
                        if (med < sh.minMedian) {
                            sh.minMedian = med;
                            for (int q = 0; q < 9; ++q) sh.best[q] = F[q];
                            sh.have_model = 1;
                        }
                    }
                }
                __syncthreads();
            }
        }
        __syncthreads();
        if (tid == 0) {
            sh.iter_base += RS_BATCH;
            if (sh.iter_base >= sh.niters) sh.stop = 1;
        }
        __syncthreads();
    }
    // ---- final mask -------------------------------------------------------------------------------------
    float t = thr;
    if (lmeds) {
        double sigma = 2.5 * 1.4826 * (1 + 5. / (n - kFmModelPoints)) * sqrt(sh.minMedian);
        sigma = fmax(sigma, 0.001);
        t = (float)(sigma * sigma);
    }
    int cnt = 0;
    for (int k = tid; k < n; k += RS_THREADS) {
        const int in = sh.have_model ? (fm_error(sh.best, m1[k], m2[k]) <= t ? 1 : 0) : 0;
        mask[k] = (uint8_t)in;
        cnt += in;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) sh.wsum[wid] = cnt;
    __syncthreads();
    if (tid == 0) {
        int tot = 0;
        for (int w = 0; w < RS_WARPS; ++w) tot += sh.wsum[w];
        R.n_inl[prob] = tot;
        R.n_iter[prob] = sh.n_scored;
    }
}

cudaError_t alloc_ransac(RansacScratch& R, int n_slots, int kp_cap, std::vector<void*>& allocs) {
    R.kp_cap = kp_cap;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)n_slots * 2 * kp_cap * sizeof(float2));
    if (e != cudaSuccess) return e;
    allocs.push_back(p); R.p1 = (float2*)p;
    e = cudaMalloc(&p, (size_t)n_slots * 2 * kp_cap * sizeof(float2));
    if (e != cudaSuccess) return e;
    allocs.push_back(p); R.p2 = (float2*)p;
    e = cudaMalloc(&p, (size_t)n_slots * 2 * sizeof(int32_t));
    if (e != cudaSuccess) return e;
    allocs.push_back(p); R.n_pts = (int32_t*)p;
    if ((e = cudaMemset(R.n_pts, 0, (size_t)n_slots * 2 * sizeof(int32_t))) != cudaSuccess) return e;
    e = cudaMalloc(&p, (size_t)n_slots * 2 * sizeof(int32_t));
    if (e != cudaSuccess) return e;
    allocs.push_back(p); R.n_iter = (int32_t*)p;
    return cudaMemset(R.n_iter, 0, (size_t)n_slots * 2 * sizeof(int32_t));
}

static int ransac_smem(int kp_cap) { return 2 * kp_cap * (int)sizeof(Pt2f); }

cudaError_t configure_ransac(int kp_cap) {
    return SVO_RAISE_DYN_SMEM(fundmat_kernel, ransac_smem(kp_cap));
}

cudaError_t launch_ransac(const Pools& P, const RansacScratch& R, int n, cudaStream_t st, int* launches,
                          StepTimer* tm) {
    ransac_prepare_kernel<<<n, 256, 0, st>>>(P, R);
    if (tm) tm->mark(st);
    fundmat_kernel<<<2 * n, RS_THREADS, ransac_smem(R.kp_cap), st>>>(R);
    if (tm) tm->mark(st);
    if (launches) *launches += 2;
    return cudaGetLastError();
}

cudaError_t ransac_host_points(const RansacScratch& R, int scratch_slot, const float* p1, const float* p2, int n,
                               uint8_t* mask, int* n_inliers, cudaStream_t st) {
    const size_t prob = (size_t)scratch_slot * 2;
    cudaError_t e = cudaMemcpyAsync(R.slots, &scratch_slot, sizeof(int), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    if (n > 0) {
        if ((e = cudaMemcpyAsync(R.p1 + prob * R.kp_cap, p1, (size_t)n * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(R.p2 + prob * R.kp_cap, p2, (size_t)n * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
    }
    const int zero = 0;
    if ((e = cudaMemcpyAsync(R.n_pts + prob, &n, sizeof(int), cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(R.n_pts + prob + 1, &zero, sizeof(int), cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;   // the host sources above are stack / caller memory
    fundmat_kernel<<<2, RS_THREADS, ransac_smem(R.kp_cap), st>>>(R);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(n_inliers, R.n_inl + prob, sizeof(int), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    if (n > 0) {
        if (n < 7) memset(mask, 0, n);
        else if ((e = cudaMemcpy(mask, R.inl + prob * R.kp_cap, n, cudaMemcpyDeviceToHost)) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// parity seam: run7Point alone, RS_GROUPS samples per warp (svo_fundmat_7point)
__global__ void __launch_bounds__(RS_THREADS)
fm7_kernel(const float2* __restrict__ p1, const float2* __restrict__ p2, int n_sets, double* __restrict__ F,
           int32_t* __restrict__ n_models) {
    __shared__ Pt2f s1[RS_BATCH][7], s2[RS_BATCH][7];
    __shared__ double sF[RS_BATCH][27];
    __shared__ int sv[RS_BATCH], sn[RS_BATCH];
    __shared__ double s_svd[RS_WARPS][96];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int w0 = wid * RS_GROUPS;
    const int set0 = (blockIdx.x * RS_WARPS + wid) * RS_GROUPS;
    if (set0 >= n_sets) return;
    for (int g = 0; g < RS_GROUPS; ++g) {
        const int set = set0 + g;
        if (lane == 0) { sv[w0 + g] = set < n_sets ? 1 : 0; sn[w0 + g] = 0; }
        if (lane < 7 && set < n_sets) {
            const float2 a = p1[set * 7 + lane], b = p2[set * 7 + lane];
            s1[w0 + g][lane].x = a.x; s1[w0 + g][lane].y = a.y; s2[w0 + g][lane].x = b.x; s2[w0 + g][lane].y = b.y;
        }
        if (lane < 27) sF[w0 + g][lane] = 0;
    }
    __syncwarp();
    run_7point_wave(s1 + w0, s2 + w0, sv + w0, sF + w0, sn + w0, s_svd[wid]);
    __syncwarp();
    for (int g = 0; g < RS_GROUPS; ++g) {
        const int set = set0 + g;
        if (set >= n_sets) break;
        if (lane < 27) F[(size_t)set * 27 + lane] = sF[w0 + g][lane];
        if (lane == 0) n_models[set] = sn[w0 + g];
    }
}

cudaError_t fm7_host_points(const float* p1, const float* p2, int n_sets, double* F, int32_t* n_models, cudaStream_t st) {
    if (n_sets <= 0) return cudaSuccess;
    float2 *d1 = nullptr, *d2 = nullptr;
    double* dF = nullptr;
    int32_t* dn = nullptr;
    cudaError_t e;
    const size_t pb = (size_t)n_sets * 7 * sizeof(float2);
    if ((e = cudaMalloc(&d1, pb)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&d2, pb)) == cudaSuccess && (e = cudaMalloc(&dF, (size_t)n_sets * 27 * sizeof(double))) == cudaSuccess &&
        (e = cudaMalloc(&dn, (size_t)n_sets * sizeof(int32_t))) == cudaSuccess &&
        (e = cudaMemcpyAsync(d1, p1, pb, cudaMemcpyHostToDevice, st)) == cudaSuccess &&
        (e = cudaMemcpyAsync(d2, p2, pb, cudaMemcpyHostToDevice, st)) == cudaSuccess) {
        fm7_kernel<<<(n_sets + RS_BATCH - 1) / RS_BATCH, RS_THREADS, 0, st>>>(d1, d2, n_sets, dF, dn);
        if ((e = cudaGetLastError()) == cudaSuccess &&
            (e = cudaMemcpyAsync(F, dF, (size_t)n_sets * 27 * sizeof(double), cudaMemcpyDeviceToHost, st)) == cudaSuccess &&
            (e = cudaMemcpyAsync(n_models, dn, (size_t)n_sets * sizeof(int32_t), cudaMemcpyDeviceToHost, st)) == cudaSuccess)
            e = cudaStreamSynchronize(st);
    }
    cudaFree(d1); cudaFree(d2); cudaFree(dF); cudaFree(dn);
    return e;
}

}  // namespace svo
