// ransac.cu -- the two cv::findFundamentalMat(p1, p2, FM_RANSAC, 1.0, 0.99, mask) calls of
// stage4_match_consecutive.cpp:171-241 on sm_100a, one CTA per (stereo pair, side).
//
// OpenCV's estimator is sequential (adaptive iteration count, one RNG stream), but iteration k's
// sample does not depend on the models of earlier iterations: thread 0 draws the next 8 samples
// from the exact cv::RNG stream (including the duplicate and collinearity re-draws), the 8 warps solve
// and score them concurrently, and thread 0 then replays OpenCV's "best so far / update niters" logic in
// iteration order.  The result is the inlier mask OpenCV returns (see fundmat_core.cuh for the one
// floating-point caveat).  With >= 15 points this is RANSAC, with 8..14 LMedS, with 7 an all-ones mask
// and with < 7 no mask (zero inliers), as in cv::findFundamentalMat.
#include "fundmat_core.cuh"
#include "ransac.cuh"

namespace svo {

constexpr int RS_WARPS = 8;
constexpr int RS_THREADS = RS_WARPS * 32;

struct RsShared {
    CvRng rng;
    double F[RS_WARPS][27];
    double best[9];
    double minMedian;
    Pt2f ms1[RS_WARPS][7], ms2[RS_WARPS][7];
    int found[RS_WARPS];
    int nmodels[RS_WARPS];
    int good[RS_WARPS][3];
    float median[RS_WARPS][3];
    int iter_base, niters, maxGood, stop, have_model;
    int wsum[RS_WARPS];
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
        const int2* __restrict__ kp = P.kp + frame_idx(slot, prev, side) * cap;
        const int2* __restrict__ kc = P.kp + frame_idx(slot, cur, side) * cap;
        float2* __restrict__ o1 = R.p1 + ((size_t)slot * 2 + side) * cap;
        float2* __restrict__ o2 = R.p2 + ((size_t)slot * 2 + side) * cap;
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            const int4 c = tc[k];
            const int4 a = mp[c.x], b = mc[side ? c.z : c.y];
            const int2 p = kp[side ? a.y : a.x], q = kc[side ? b.y : b.x];
            o1[k] = make_float2((float)(p.x & 0xFFFF), (float)((unsigned)p.x >> 16));
            o2[k] = make_float2((float)(q.x & 0xFFFF), (float)((unsigned)q.x >> 16));
        }
        if (threadIdx.x == 0) R.n_pts[slot * 2 + side] = n;
    }
}


// run_7point with the null-space step (Gauss-Jordan with complete pivoting on the 7x9 system) spread over the warp:
// lane j < 9 owns column j in registers, the pivot search is a warp arg-max with the sequential scan's tie order
// (largest |a|, then smallest row, then smallest column), row swaps / scaling / elimination touch each element with
// exactly the operations of the sequential loop, so the models are bit-identical to run_7point's.  All 32 lanes must
// call it; F and the return value are valid on lane 0.  (On one lane the dynamically indexed 7x9 array lives in local
// memory and this step was more than half of fundmat_kernel's time.)
__device__ int run_7point_warp(const Pt2f* m1, const Pt2f* m2, double* F /* 27, lane 0 */) {
    const int lane = threadIdx.x & 31;
    const unsigned full = 0xffffffffu;
    Fm7Norm nm;
    if (!fm7_normalise(m1, m2, nm)) return 0;          // same inputs on every lane: uniform
    const bool colv = lane < 9;
    double col[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) col[i] = colv ? fm7_entry(m1, m2, nm, i, lane) : 0.0;
    bool used = !colv;                                 // lanes >= 9 never offer a pivot
    int pivcol[7];
    int rank = 0;
#pragma unroll
    for (int r = 0; r < 7; ++r) {
        // this column's best candidate among rows r..6 (strict '>' keeps the smallest row among equals)
        double bv = 0;
        int bi = -1;
        if (!used) {
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                if (i < r) continue;
                const double v = fabs(col[i]);
                if (v > bv) { bv = v; bi = i; }
            }
        }
        int bl = lane;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {              // arg-max over lanes 0..15: value, then smaller row, then smaller column
            const double ov = __shfl_xor_sync(full, bv, o);
            const int oi = __shfl_xor_sync(full, bi, o);
            const int ol = __shfl_xor_sync(full, bl, o);
            const bool take = oi >= 0 && (bi < 0 || ov > bv || (ov == bv && (oi < bi || (oi == bi && ol < bl))));
            if (take) { bv = ov; bi = oi; bl = ol; }
        }
        const double best = __shfl_sync(full, bv, 0);
        const int pr = __shfl_sync(full, bi, 0), pc = __shfl_sync(full, bl, 0);
        if (pr < 0 || best < 1e-13) break;
        if (pr != r) {                                 // swap rows r and pr (pr > r)
            double tmp = col[r];
#pragma unroll
            for (int i = 0; i < 7; ++i) if (i == pr) { col[r] = col[i]; col[i] = tmp; }
        }
        const double inv = 1. / __shfl_sync(full, col[r], pc);
        col[r] *= inv;
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            if (i == r) continue;
            const double f = __shfl_sync(full, col[i], pc);
            if (f != 0) col[i] -= f * col[r];
        }
        if (lane == pc) used = true;
        pivcol[r] = pc;
        ++rank;
    }
    // the first two free columns span the null-space basis: f[fc] = 1, f[pivcol[r]] = -a[r][fc]
    const unsigned freem = __ballot_sync(full, !used);
    if (__popc(freem) < 2) return 0;
    const int fc0 = __ffs(freem) - 1, fc1 = __ffs(freem & (freem - 1)) - 1;
    double e1 = lane == fc0 ? 1.0 : 0.0, e2 = lane == fc1 ? 1.0 : 0.0;      // this lane's entries of f1 / f2
#pragma unroll
    for (int r = 0; r < 7; ++r) {
        const double v1 = __shfl_sync(full, col[r], fc0), v2 = __shfl_sync(full, col[r], fc1);
        if (r < rank && pivcol[r] == lane) { e1 = -v1; e2 = -v2; }
    }
    double f1[9], f2[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) { f1[j] = __shfl_sync(full, e1, j); f2[j] = __shfl_sync(full, e2, j); }
    int n = 0;
    if (lane == 0) n = fm7_finish(f1, f2, nm, F);
    return __shfl_sync(full, n, 0);
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
        if (tid == 0) R.n_inl[prob] = 0;
        return;
    }
    if (n == 7) {                      // mask.setTo(1)
        if (tid < 7) mask[tid] = 1;
        if (tid == 0) R.n_inl[prob] = 7;
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
        sh.minMedian = DBL_MAX;
        if (lmeds) { const int it = ransac_update_num_iters(kFmConfidence, 0.45, kFmModelPoints, kFmMaxIters); sh.niters = it > 3 ? it : 3; }
        else sh.niters = kFmMaxIters;
    }
    __syncthreads();
    const float thr = (float)(kFmThreshold * kFmThreshold);

    while (!sh.stop) {
        // ---- thread 0: the next RS_WARPS samples from the cv::RNG stream ------------------------------
        if (tid == 0) {
            for (int w = 0; w < RS_WARPS; ++w) {
                sh.found[w] = -1;   // not needed
                if (sh.iter_base + w >= sh.niters) break;
                int idx[7];
                const bool ok = get_subset7(m1, m2, n, sh.rng, lmeds ? 1000 : 10000, idx, sh.ms1[w], sh.ms2[w]);
                sh.found[w] = ok ? 1 : 0;
                if (!ok) break;
            }
        }
        __syncthreads();
        // ---- every warp: solve and score its sample -------------------------------------------------
        if (sh.found[wid] == 1) {
            {
                const int nm7 = run_7point_warp(sh.ms1[wid], sh.ms2[wid], sh.F[wid]);
                if (lane == 0) sh.nmodels[wid] = nm7;
            }
            __syncwarp();
            const int nm = sh.nmodels[wid];
            for (int i = 0; i < nm; ++i) {
                const double* F = sh.F[wid] + 9 * i;
                if (!lmeds) {
                    int good = 0;
                    for (int k = lane; k < n; k += 32) good += fm_error(F, m1[k], m2[k]) <= thr ? 1 : 0;
                    for (int o = 16; o > 0; o >>= 1) good += __shfl_xor_sync(0xffffffffu, good, o);
                    if (lane == 0) sh.good[wid][i] = good;
                } else if (lane == 0) {
                    float e[14];
                    for (int k = 0; k < n; ++k) {   // n <= 14: insertion sort, element n/2 is nth_element's answer
                        const float v = fm_error(F, m1[k], m2[k]);
                        int j = k;
                        while (j > 0 && e[j - 1] > v) { e[j] = e[j - 1]; --j; }
                        e[j] = v;
                    }
                    sh.median[wid][i] = e[n / 2];
                }
            }
        }
        __syncthreads();
        // ---- thread 0: OpenCV's sequential bookkeeping, in iteration order ------------------------------
        if (tid == 0) {
            for (int w = 0; w < RS_WARPS; ++w) {
                const int iter = sh.iter_base + w;
                if (iter >= sh.niters) { sh.stop = 1; break; }
                if (sh.found[w] != 1) { sh.stop = 1; break; }   // getSubset failed: return false (iter 0) or break
                const int nm = sh.nmodels[w];
                for (int i = 0; i < nm; ++i) {
                    if (!lmeds) {
                        const int good = sh.good[w][i];
                        if (good > max(sh.maxGood, kFmModelPoints - 1)) {
                            sh.maxGood = good;
                            for (int q = 0; q < 9; ++q) sh.best[q] = sh.F[w][9 * i + q];
                            sh.have_model = 1;
                            sh.niters = ransac_update_num_iters(kFmConfidence, (double)(n - good) / n, kFmModelPoints, sh.niters);
                        }
                    } else {
                        const double med = (double)sh.median[w][i];
                        if (med < sh.minMedian) {
                            sh.minMedian = med;
                            for (int q = 0; q < 9; ++q) sh.best[q] = sh.F[w][9 * i + q];
                            sh.have_model = 1;
                        }
                    }
                }
            }
            sh.iter_base += RS_WARPS;
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
    return cudaMemset(R.n_pts, 0, (size_t)n_slots * 2 * sizeof(int32_t));
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

}  // namespace svo
