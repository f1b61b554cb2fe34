// match.cu -- stages 3 and 4 of the hot path on sm_100a: brute-force 256-bit Hamming matching,
// the left/right filter, the greedy inter-frame filter and the consistency check.  Replaces
//   cv::BFMatcher(NORM_HAMMING,false).match      stage3_match_left_right.cpp:88-94,
//                                                stage4_match_consecutive.cpp:137-142
//   1-to-1 + epipolar/threshold/disparity filter stage3_match_left_right.cpp:124-175
//   descriptor gather                            stage4_match_consecutive.cpp:97-131
//   greedy distance/collision filter             stage4_match_consecutive.cpp:145-160
//   F-matrix mask application + consistency + ID management   stage4_match_consecutive.cpp:243-305
//   tracking statistics and the bad-tracking gate             stage4_match_consecutive.cpp:743-752,
//                                                             process_new_image_pair.cpp:326-330
#include "match.cuh"

namespace svo {

// ------------------------------------------------------------------------------------------------
// K4  hamming_kernel: for every query row the packed minimum (distance << 16 | trainIdx) over all
// train rows, i.e. smallest distance, lowest index among ties -- BFMatcher's rule.
//   CTA = 128 threads, QPT queries per thread held in registers (8 words each); the train set streams
//   through shared memory in tiles of TILE rows (32 B each, read as two broadcast LDS.128).
// Roofline: 8 XOR + 8 POPC per pair-distance; the POPC pipe (16 lanes/clk/SM) binds, HBM traffic is
// 32 B per descriptor once (L2 serves the re-reads of the train tile by the other CTAs).
// ------------------------------------------------------------------------------------------------
constexpr int HM_THREADS = 128;
constexpr int HM_QPT = 2;     // queries per thread (hamming_body dispatches on 1 or 2 live groups per warp)
constexpr int HM_TILE = 512;   // train rows per shared-memory tile (16 KB)

// 256-bit Hamming distance with a carry-save adder tree: four LOP3 full adders compress the eight XOR words
// into two weight-1 planes (s3, w7), one weight-2 plane (t1) and one weight-4 plane (f1), so 4 POPC are needed
// instead of 8 (POPC issues at a quarter of the LOP3 rate) for 16 LOP3 in total; the weighted sum and the
// (distance << 16 | index) packing are IMADs (FMA pipe), leaving the ALU pipe 16 LOP3 + 1 min per pair-distance.
// The full adders are written as explicit LOP3s (sum = 0x96, carry = 0xE8): left to itself nvcc re-associates the
// query XORs into the adders and ends up with 19 LOP3 per pair-distance instead of 16.
__device__ __forceinline__ uint32_t lop3_xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r; asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ uint32_t lop3_maj3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r; asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
__device__ __forceinline__ uint32_t hamming256_packed(const uint4& qa, const uint4& qb, const uint4& ta, const uint4& tb,
                                                      uint32_t idx) {
    const uint32_t w0 = qa.x ^ ta.x, w1 = qa.y ^ ta.y, w2 = qa.z ^ ta.z, w3 = qa.w ^ ta.w;
    const uint32_t w4 = qb.x ^ tb.x, w5 = qb.y ^ tb.y, w6 = qb.z ^ tb.z, w7 = qb.w ^ tb.w;
    const uint32_t s1 = lop3_xor3(w0, w1, w2), c1 = lop3_maj3(w0, w1, w2);
    const uint32_t s2 = lop3_xor3(w3, w4, w5), c2 = lop3_maj3(w3, w4, w5);
    const uint32_t s3 = lop3_xor3(s1, s2, w6), c3 = lop3_maj3(s1, s2, w6);
    const uint32_t t1 = lop3_xor3(c1, c2, c3), f1 = lop3_maj3(c1, c2, c3);
    const uint32_t x = (uint32_t)__popc(t1) * 2u + (uint32_t)__popc(s3);
    const uint32_t y = (uint32_t)__popc(f1) * 4u + (uint32_t)__popc(w7);
    return y * 65536u + (x * 65536u + idx);
}

struct HammingProblem {
    const uint4* q; const uint4* t; uint32_t* out; int nq, nt;
};

__device__ __forceinline__ HammingProblem pipeline_problem(const Pools& P, int mode, int item) {
    HammingProblem h;
    const int slot = P.slots[item];
    const SlotState& s = P.state[slot];
    const int cur = s.cur, prev = cur ^ 1;
    const size_t cap = (size_t)P.p.kp_cap;
    if (mode == 0) {          // left -> right of the current frame (S3:91-94)
        h.q = P.desc + frame_idx(slot, cur, 0) * cap * 2;
        h.t = P.desc + frame_idx(slot, cur, 1) * cap * 2;
        h.nq = P.n_kp[frame_idx(slot, cur, 0)];
        h.nt = P.n_kp[frame_idx(slot, cur, 1)];
        h.out = P.lr_best + (size_t)slot * cap;
    } else {                  // prev -> cur, mode 1 = left-left, 2 = right-right (S4:141-142)
        const int side = mode - 1;
        h.q = P.mdesc + frame_idx(slot, prev, side) * cap * 2;
        h.t = P.mdesc + frame_idx(slot, cur, side) * cap * 2;
        h.nq = s.has_prev ? P.n_match[pair_idx(slot, prev)] : 0;
        h.nt = P.n_match[pair_idx(slot, cur)];
        h.out = P.tr_best + ((size_t)slot * 2 + side) * cap;
    }
    return h;
}

// 128-bit prefix of the distance (the first four words): a lower bound of the full distance, 3 POPC + 6 LOP3
__device__ __forceinline__ uint32_t hamming128(const uint4& q, const uint4& t) {
    const uint32_t w0 = q.x ^ t.x, w1 = q.y ^ t.y, w2 = q.z ^ t.z, w3 = q.w ^ t.w;
    const uint32_t s1 = lop3_xor3(w0, w1, w2), c1 = lop3_maj3(w0, w1, w2);
    return (uint32_t)__popc(c1) * 2u + (uint32_t)__popc(s1) + (uint32_t)__popc(w3);
}

// inner loop over one staged train tile for NQ queries per thread (NQ is warp-uniform).
// PRUNE (pipeline only): every consumer of the pipeline's matches discards distances above orb_th (S3:164, S4:148),
// so `best` starts at (orb_th + 1) << 16 and a pair is finished only when its 128-bit prefix distance is below the
// best distance so far -- otherwise the full distance cannot win (train indices only grow, so an equal distance loses
// the lowest-index tie-break too).  On ORB descriptors the prefix of a non-matching pair is below 41 in 0.4 % of the
// pairs: the second half of the distance is computed for a few percent of the warp-instructions.
template <int NQ, bool PRUNE>
__device__ __forceinline__ void hamming_tile(const uint4* __restrict__ s_t, int rows, int t0, const uint4* qa, const uint4* qb,
                                             uint32_t* best) {
#pragma unroll 4
    for (int j = 0; j < rows; ++j) {
        const uint4 ta = s_t[2 * j], tb = s_t[2 * j + 1];
#pragma unroll
        for (int r = 0; r < NQ; ++r) {
            if (PRUNE) {
                const uint32_t d1 = hamming128(qa[r], ta);
                if (d1 < (best[r] >> 16)) {
                    const uint32_t d = d1 + hamming128(qb[r], tb);
                    best[r] = min(best[r], (d << 16) | (uint32_t)(t0 + j));
                }
            } else {
                best[r] = min(best[r], hamming256_packed(qa[r], qb[r], ta, tb, (uint32_t)(t0 + j)));
            }
        }
    }
}

template <bool PRUNE>
__device__ __forceinline__ void hamming_body(const HammingProblem& h, uint4* s_t, uint32_t best_init) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q0 = blockIdx.x * (HM_THREADS * HM_QPT);
    if (q0 >= h.nq) return;
    // a warp owns 32 * HM_QPT consecutive queries; groups of 32 that lie beyond nq are skipped by the whole warp
    // (a ragged last CTA then costs what its queries cost, not a full CTA)
    const int wq0 = q0 + warp * (32 * HM_QPT);
    int nq_w = 0;                                   // query groups of this warp with at least one valid query
    uint4 qa[HM_QPT], qb[HM_QPT];
    uint32_t best[HM_QPT];
#pragma unroll
    for (int r = 0; r < HM_QPT; ++r) {
        const int qi = wq0 + r * 32 + lane;
        best[r] = best_init;
        if (wq0 + r * 32 < h.nq) nq_w = r + 1;
        if (qi < h.nq) { qa[r] = h.q[2 * (size_t)qi]; qb[r] = h.q[2 * (size_t)qi + 1]; }
        else { qa[r] = make_uint4(0, 0, 0, 0); qb[r] = qa[r]; }
    }
    // gridDim.z > 1 (small batches only): the train set is split over the z CTAs of a query block, every CTA keeps the
    // packed minimum of its share and the shares meet in an atomicMin on the output word, which hamming_init_kernel has set
    // to best_init -- min over (distance << 16 | index) is associative, so the result is the single-pass one, ties included
    const int splits = (int)gridDim.z;
    const int chunk = splits > 1 ? ((h.nt + splits - 1) / splits + 31) & ~31 : h.nt;
    const int t_lo = splits > 1 ? (int)blockIdx.z * chunk : 0, t_hi = min(h.nt, t_lo + chunk);
    for (int t0 = t_lo; t0 < t_hi; t0 += HM_TILE) {
        const int rows = min(HM_TILE, t_hi - t0);
        __syncthreads();
        for (int i = tid; i < rows * 2; i += HM_THREADS) s_t[i] = h.t[2 * (size_t)t0 + i];
        __syncthreads();
        if (nq_w == HM_QPT) hamming_tile<HM_QPT, PRUNE>(s_t, rows, t0, qa, qb, best);
        else if (nq_w == 1) hamming_tile<1, PRUNE>(s_t, rows, t0, qa, qb, best);
    }
#pragma unroll
    for (int r = 0; r < HM_QPT; ++r) {
        const int qi = wq0 + r * 32 + lane;
        if (qi < h.nq) {
            if (splits > 1) atomicMin(&h.out[qi], best[r]);
            else h.out[qi] = best[r];
        }
    }
}

__global__ void __launch_bounds__(HM_THREADS)
hamming_pipeline_kernel(Pools P, int mode_base, int n_modes) {
    __shared__ uint4 s_t[HM_TILE * 2];
    const int item = blockIdx.y / n_modes, mode = mode_base + blockIdx.y % n_modes;
    const HammingProblem h = pipeline_problem(P, mode, item);
    // distances above orb_th are dropped by lr_filter_kernel / track_filter_kernel: (orb_th + 1, index 0) stands for them
    const uint32_t cut = (uint32_t)min(max(slot_orb_th(P.p, P.state[P.slots[item]]), 0) + 1, 0x7FFF);
    hamming_body<true>(h, s_t, cut << 16);
}

// split-train launches only: every output word starts at the value an empty scan would leave
__global__ void __launch_bounds__(256)
hamming_init_kernel(Pools P, int mode_base, int n_modes) {
    const int item = blockIdx.y / n_modes, mode = mode_base + blockIdx.y % n_modes;
    const HammingProblem h = pipeline_problem(P, mode, item);
    const uint32_t cut = (uint32_t)min(max(slot_orb_th(P.p, P.state[P.slots[item]]), 0) + 1, 0x7FFF);
    const int qi = blockIdx.x * 256 + threadIdx.x;
    if (qi < h.nq) h.out[qi] = cut << 16;
}

__global__ void __launch_bounds__(HM_THREADS)
hamming_raw_kernel(const uint4* __restrict__ q, int nq, const uint4* __restrict__ t, int nt, uint32_t* __restrict__ out) {
    __shared__ uint4 s_t[HM_TILE * 2];
    HammingProblem h;
    h.q = q + (size_t)blockIdx.y * nq * 2;
    h.t = t + (size_t)blockIdx.y * nt * 2;
    h.out = out + (size_t)blockIdx.y * nq;
    h.nq = nq; h.nt = nt;
    hamming_body<false>(h, s_t, 0xFFFFFFFFu);
}

// ------------------------------------------------------------------------------------------------
// block-wide exclusive scan helper (256 threads)
// ------------------------------------------------------------------------------------------------
constexpr int FT = 256;
__device__ __forceinline__ int scan256(int v, int* s_warp, int& total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    int wsum = (lane < FT / 32) ? s_warp[lane] : 0;
    int wincl = wsum;
#pragma unroll
    for (int o = 1; o < FT / 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, wincl, o);
        if (lane >= o) wincl += t;
    }
    const int wexcl = __shfl_sync(0xffffffffu, wincl - wsum, wid);
    total = __shfl_sync(0xffffffffu, wincl, FT / 32 - 1);
    __syncthreads();
    return wexcl + incl - v;
}

// ------------------------------------------------------------------------------------------------
// K5  lr_filter_kernel: one CTA per slot.  S3:124-175 + descriptor gather of S4:105-131.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FT)
lr_filter_kernel(Pools P) {
    extern __shared__ uint32_t s_rc[];   // right_cand: (distance<<16 | queryIdx) per right keypoint
    __shared__ int s_warp[FT / 32];
    const int slot = P.slots[blockIdx.x];
    SlotState& st = P.state[slot];
    const int cur = st.cur;
    const size_t cap = (size_t)P.p.kp_cap;
    const size_t fL = frame_idx(slot, cur, 0), fR = frame_idx(slot, cur, 1);
    const int nL = P.n_kp[fL], nR = P.n_kp[fR];
    const uint32_t* __restrict__ best = P.lr_best + (size_t)slot * cap;
    const int tid = threadIdx.x;
    const int n_in = (nR > 0) ? nL : 0;   // BFMatcher returns no matches for an empty train set
    const int orb_th = slot_orb_th(P.p, st);          // m_current_orb_th (S3:164)

    // P:254-267 (resetIds()): before stage 3 the previous frame's IDs are renumbered 0..N-1, the ID generator
    // restarts behind them and the previous frame becomes the key frame.  (The reference dereferences a null
    // m_prev_imgpair when there is no previous frame; here the flag simply stays pending.)
    if (st.reset_ids && st.has_prev) {
        const int Mp = P.p.use_ids ? P.n_match[pair_idx(slot, cur ^ 1)] : 0;   // matches_IDs is empty without ID control
        long long* idp = P.match_id + pair_idx(slot, cur ^ 1) * cap;
        for (int m = tid; m < Mp; m += FT) idp[m] = m;
        __syncthreads();
        if (tid == 0) { st.last_match_id = Mp; st.last_kf_max_id = (long long)Mp - 1; st.reset_ids = 0; }
        __syncthreads();
    }

    if (P.p.robust_1to1) {   // S3:124-148: per right keypoint the left one with the smallest distance
        for (int i = tid; i < nR; i += FT) s_rc[i] = 0xFFFFFFFFu;
        __syncthreads();
        for (int k = tid; k < n_in; k += FT) {
            const uint32_t b = best[k];
            atomicMin(&s_rc[b & 0xFFFFu], (b & 0xFFFF0000u) | (uint32_t)k);  // strict '>' in the reference: first k wins ties
        }
        __syncthreads();
    }
    int4* __restrict__ mout = P.match + pair_idx(slot, cur) * cap;
    uint4* __restrict__ mdL = P.mdesc + frame_idx(slot, cur, 0) * cap * 2;
    uint4* __restrict__ mdR = P.mdesc + frame_idx(slot, cur, 1) * cap * 2;
    const uint4* __restrict__ dL = P.desc + fL * cap * 2;
    const uint4* __restrict__ dR = P.desc + fR * cap * 2;
    int base = 0;
    for (int k0 = 0; k0 < n_in; k0 += FT) {
        const int k = k0 + tid;
        int keep = 0, t = 0, dist = 0;
        if (k < n_in) {
            const uint32_t b = best[k];
            t = (int)(b & 0xFFFFu);
            dist = (int)(b >> 16);
            keep = 1;
            if (P.p.robust_1to1 && (int)(s_rc[t] & 0xFFFFu) != k) keep = 0;
            const float2 l = kp_xy(P, fL * cap, k), r = kp_xy(P, fR * cap, t);
            // S3:162-165: float differences truncated to int
            const int diff = (int)__fsub_rn(l.y, r.y);
            const int disp = (int)__fsub_rn(l.x, r.x);
            if ((double)abs(diff) > P.p.max_y_diff || (float)dist > (float)orb_th ||
                (double)disp < 1.0 || (double)disp > (double)P.p.W) keep = 0;
        }
        int total;
        const int off = scan256(keep, s_warp, total);
        if (keep) {
            const int m = base + off;
            mout[m] = make_int4(k, t, dist, 0);
            mdL[2 * m] = dL[2 * k]; mdL[2 * m + 1] = dL[2 * k + 1];
            mdR[2 * m] = dR[2 * t]; mdR[2 * m + 1] = dR[2 * t + 1];
        }
        base += total;
    }
    if (tid == 0) {
        P.n_match[pair_idx(slot, cur)] = base;
        P.oct_stats[stat_idx(slot, P.octave, 2)] = base;
        if (P.octave == 0) P.result[slot].stereo_matches = base;   // svo_result mirrors stereo_matches[0]
    }
    // S3:67,172-173: on the first frame every kept match receives a fresh ID
    if (P.p.use_ids && !st.has_prev) {
        long long* ids = P.match_id + pair_idx(slot, cur) * cap;
        const long long b0 = st.last_match_id;
        for (int m = tid; m < base; m += FT) ids[m] = b0 + m;
        __syncthreads();
        if (tid == 0) st.last_match_id = b0 + base;
    }
}

// ------------------------------------------------------------------------------------------------
// K7  track_filter_kernel: one CTA per slot.  S4:145-160: pair k is erased if one of its two distances exceeds the
// threshold or if its left or right train index was already claimed by an EARLIER pair that was itself kept.  The
// reference walks the list sequentially; here the same fixed point is reached in parallel rounds:
//   minL[v] / minR[v] = earliest pair not (yet) rejected among those whose left / right train index is v;
//   a pair that is the earliest in both of its chains is KEPT (everything before it in those chains is rejected for
//   good, so at its turn the sequential walk finds both indices unclaimed); a pair whose chain head is an earlier KEPT
//   pair is REJECTED.  The smallest undecided pair is always decided, so the rounds end; real lists need 2-3 rounds
//   (collisions are rare).  After kTfMaxRounds a single thread finishes the undecided rest sequentially.
// ------------------------------------------------------------------------------------------------
constexpr int kTfMaxRounds = 12;
enum : uint8_t { TF_UNDECIDED = 0, TF_KEPT = 1, TF_REJECTED = 2 };

__global__ void __launch_bounds__(FT)
track_filter_kernel(Pools P) {
    extern __shared__ uint32_t s_mem[];   // [cap] packed (trainL << 16 | trainR) | [cap] minL | [cap] minR | [cap] states (bytes)
    __shared__ int s_warp[FT / 32];
    __shared__ int s_undecided;
    const int slot = P.slots[blockIdx.x];
    const SlotState& st = P.state[slot];
    const int cur = st.cur, prev = cur ^ 1;
    const int cap = P.p.kp_cap;
    const int tid = threadIdx.x;
    const int Mp = st.has_prev ? P.n_match[pair_idx(slot, prev)] : 0;
    const int Mc = P.n_match[pair_idx(slot, cur)];
    const int n_in = (Mc > 0) ? Mp : 0;
    uint32_t* s_ab = s_mem;
    uint32_t* minL = s_mem + cap;
    uint32_t* minR = s_mem + 2 * cap;
    uint8_t* state = reinterpret_cast<uint8_t*>(s_mem + 3 * cap);
    const uint32_t* __restrict__ bl = P.tr_best + ((size_t)slot * 2 + 0) * cap;
    const uint32_t* __restrict__ br = P.tr_best + ((size_t)slot * 2 + 1) * cap;
    const uint32_t th = (uint32_t)slot_orb_th(P.p, st);          // m_current_orb_th (S4:149)
    for (int k = tid; k < n_in; k += FT) {
        const uint32_t a = bl[k], b = br[k];
        s_ab[k] = (a << 16) | (b & 0xFFFFu);
        state[k] = ((a >> 16) > th || (b >> 16) > th) ? TF_REJECTED : TF_UNDECIDED;
    }
    __syncthreads();
    for (int round = 0; round < kTfMaxRounds; ++round) {
        for (int v = tid; v < Mc; v += FT) { minL[v] = 0xFFFFFFFFu; minR[v] = 0xFFFFFFFFu; }
        if (tid == 0) s_undecided = 0;
        __syncthreads();
        for (int k = tid; k < n_in; k += FT)
            if (state[k] != TF_REJECTED) { atomicMin(&minL[s_ab[k] >> 16], (uint32_t)k); atomicMin(&minR[s_ab[k] & 0xFFFFu], (uint32_t)k); }
        __syncthreads();
        for (int k = tid; k < n_in; k += FT)
            if (state[k] == TF_UNDECIDED && minL[s_ab[k] >> 16] == (uint32_t)k && minR[s_ab[k] & 0xFFFFu] == (uint32_t)k) state[k] = TF_KEPT;
        __syncthreads();
        for (int k = tid; k < n_in; k += FT)
            if (state[k] == TF_UNDECIDED) {
                const uint32_t hl = minL[s_ab[k] >> 16], hr = minR[s_ab[k] & 0xFFFFu];     // chain heads, <= k
                if ((hl < (uint32_t)k && state[hl] == TF_KEPT) || (hr < (uint32_t)k && state[hr] == TF_KEPT)) state[k] = TF_REJECTED;
                else s_undecided = 1;
            }
        __syncthreads();
        const int undecided = s_undecided;
        __syncthreads();                       // thread 0 clears the flag at the top of the next round
        if (!undecided) break;
    }
    if (s_undecided && tid == 0) {            // pathological collision chains: finish sequentially (decided states are final)
        for (int v = 0; v < Mc; ++v) { minL[v] = 0; minR[v] = 0; }      // reused as "claimed" flags
        for (int k = 0; k < n_in; ++k) {
            const uint32_t tl = s_ab[k] >> 16, tr = s_ab[k] & 0xFFFFu;
            if (state[k] == TF_UNDECIDED) state[k] = (minL[tl] || minR[tr]) ? TF_REJECTED : TF_KEPT;
            if (state[k] == TF_KEPT) { minL[tl] = 1; minR[tr] = 1; }
        }
    }
    __syncthreads();
    int4* __restrict__ out = P.tcand + (size_t)slot * cap;
    int base = 0;
    for (int k0 = 0; k0 < n_in; k0 += FT) {
        const int k = k0 + tid;
        const int keep = (k < n_in && state[k] == TF_KEPT) ? 1 : 0;
        int total;
        const int off = scan256(keep, s_warp, total);
        if (keep) out[base + off] = make_int4(k, (int)(s_ab[k] >> 16), (int)(s_ab[k] & 0xFFFFu), 0);
        base += total;
    }
    if (tid == 0) {
        P.n_tcand[slot] = base;
        // masks default to "not computed": zero inliers => filter skipped (S4:243-259)
        P.n_inl[slot * 2 + 0] = 0;
        P.n_inl[slot * 2 + 1] = 0;
    }
}

// ------------------------------------------------------------------------------------------------
// Alternative matchers of the reference, restated with their quirks.
// Both are band-limited: a few candidates per query instead of the full N x N brute force.
// ------------------------------------------------------------------------------------------------
// Hamming distance as those branches compute it: byte-wise popcounts summed in a uint8_t (S3:321-331, S4:590-603),
// i.e. the true distance modulo 256.
__device__ __forceinline__ unsigned hamming256_u8(const uint4* __restrict__ a, const uint4* __restrict__ b) {
    const uint4 a0 = a[0], a1 = a[1], b0 = b[0], b1 = b[1];
    const int d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                  __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
    return (unsigned)d & 0xFFu;
}

// number of entries of the ascending row list rows[0..n) that are <= r  (rows(i) reads the row of entry i)
template <class RowOf>
__device__ __forceinline__ int count_rows_le(RowOf rows, int n, int r) {
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (rows(mid) <= r) lo = mid + 1; else hi = mid; }
    return lo;
}
// the same for float rows (dmORB keypoints): the reference compares pt.y <= int(y) in float (S3:436)
template <class RowOf>
__device__ __forceinline__ int count_rows_le_f(RowOf rows, int n, int r) {
    int lo = 0, hi = n;
    const float rf = (float)r;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (rows(mid) <= rf) lo = mid + 1; else hi = mid; }
    return lo;
}

// pyr_feats_index as m_update_indexes leaves it (S2:103-129): rows before the first feature row and from the last
// feature row on hold 0, the others the number of features with row <= r.
__device__ void build_feats_index(const int2* __restrict__ kp, int n, int H, int* idx) {
    const int y_first = n ? (int)((unsigned)kp[0].x >> 16) : 0, y_last = n ? (int)((unsigned)kp[n - 1].x >> 16) : 0;
    for (int r = threadIdx.x; r < H; r += blockDim.x)
        idx[r] = (n && r >= y_first && r < y_last) ? count_rows_le([&](int i) { return (int)((unsigned)kp[i].x >> 16); }, n, r) : 0;
}

// K5b  rbr_match_kernel: stage 3, smDescRbR branch (S3:185-419), one CTA per slot.
//   shared: idxL[H] idxR[H] | cand[kp_cap] (min_1 << 16 | right index, or ~0) | assign[kp_cap]
__global__ void __launch_bounds__(FT)
rbr_match_kernel(Pools P) {
    extern __shared__ uint32_t s_dyn[];
    __shared__ int s_warp[FT / 32];
    const int H = P.p.H, W = P.p.W;
    int* idxL = reinterpret_cast<int*>(s_dyn);
    int* idxR = idxL + H;
    uint32_t* cand = s_dyn + 2 * H;
    uint32_t* assign = cand + P.p.kp_cap;
    const int slot = P.slots[blockIdx.x];
    SlotState& st = P.state[slot];
    const int cur = st.cur;
    const size_t cap = (size_t)P.p.kp_cap;
    const size_t fL = frame_idx(slot, cur, 0), fR = frame_idx(slot, cur, 1);
    const int nL = P.n_kp[fL], nR = P.n_kp[fR];
    const int2* __restrict__ kpL = P.kp + fL * cap;
    const int2* __restrict__ kpR = P.kp + fR * cap;
    const uint4* __restrict__ dL = P.desc + fL * cap * 2;
    const uint4* __restrict__ dR = P.desc + fR * cap * 2;
    const int tid = threadIdx.x;

    if (st.reset_ids && st.has_prev) {           // P:254-267, see lr_filter_kernel
        const int Mp = P.p.use_ids ? P.n_match[pair_idx(slot, cur ^ 1)] : 0;
        long long* idp = P.match_id + pair_idx(slot, cur ^ 1) * cap;
        for (int m = tid; m < Mp; m += FT) idp[m] = m;
        __syncthreads();
        if (tid == 0) { st.last_match_id = Mp; st.last_kf_max_id = (long long)Mp - 1; st.reset_ids = 0; }
    }
    build_feats_index(kpL, nL, H, idxL);
    build_feats_index(kpR, nR, H, idxR);
    for (int j = tid; j < nR; j += FT) assign[j] = 0xFFFFFFFFu;
    __syncthreads();

    // best right candidate of every left feature (S3:252-352): the loop at row y works on the left features
    // [idxL[y], idxL[y+1]) and the right features [idxR[max(0,y-tol)], idxR[min(H-1,y+tol)])
    const int tol = (int)round(P.p.max_y_diff);
    const int max_disparity = (int)(W * 0.7);
    const unsigned max_distance = (unsigned)P.p.rbr_max_distance;
    const bool robust = P.p.robust_1to1 != 0;
    for (int i = tid; i < nL; i += FT) {
        uint32_t best = 0xFFFFFFFFu;
        const int2 l = kpL[i];
        const int y = (int)((unsigned)l.x >> 16) - 1;
        if (y >= 0 && y + 1 < H) {
            const int L0 = idxL[y], L1 = idxL[y + 1];
            const int R0 = idxR[max(0, y - tol)], R1 = idxR[min(H - 1, y + tol)];
            if (i >= L0 && i < L1 && R1 != R0) {
                const float xl = kp_xy(P, fL * cap, i).x;
                for (int j = R0; j < R1; ++j) {
                    const int disparity = (int)__fsub_rn(xl, kp_xy(P, fR * cap, j).x);      // S3:285: float difference, truncated
                    if (disparity < 1 || disparity > max_disparity) continue;
                    const unsigned d = hamming256_u8(dL + 2 * (size_t)i, dR + 2 * (size_t)j);
                    if (d > max_distance) continue;
                    best = min(best, (d << 16) | (unsigned)j);        // strict '<' in the reference: first index wins ties
                }
            }
        }
        cand[i] = best;
        // S3:355-392: a right feature keeps its FIRST claimant, or with enable_robust_1to1_match the BEST one
        // (strictly smaller distance replaces): both are a minimum over the claimants, walked in index order
        if (best != 0xFFFFFFFFu) atomicMin(&assign[best & 0xFFFFu], robust ? ((best & 0xFFFF0000u) | (unsigned)i) : (unsigned)i);
    }
    __syncthreads();

    // S3:398-410: matches in left-index order; distance = the winning claimant's
    int4* __restrict__ mout = P.match + pair_idx(slot, cur) * cap;
    uint4* __restrict__ mdL = P.mdesc + frame_idx(slot, cur, 0) * cap * 2;
    uint4* __restrict__ mdR = P.mdesc + frame_idx(slot, cur, 1) * cap * 2;
    int base = 0;
    for (int i0 = 0; i0 < nL; i0 += FT) {
        const int i = i0 + tid;
        int keep = 0, t = 0, dist = 0;
        if (i < nL && cand[i] != 0xFFFFFFFFu) {
            t = (int)(cand[i] & 0xFFFFu);
            dist = (int)(cand[i] >> 16);
            keep = (int)(assign[t] & 0xFFFFu) == i;
        }
        int total;
        const int off = scan256(keep, s_warp, total);
        if (keep) {
            const int m = base + off;
            mout[m] = make_int4(i, t, dist, 0);
            mdL[2 * m] = dL[2 * i]; mdL[2 * m + 1] = dL[2 * i + 1];
            mdR[2 * m] = dR[2 * t]; mdR[2 * m + 1] = dR[2 * t + 1];
        }
        base += total;
    }
    if (tid == 0) {
        P.n_match[pair_idx(slot, cur)] = base;
        P.oct_stats[stat_idx(slot, P.octave, 2)] = base;
        if (P.octave == 0) P.result[slot].stereo_matches = base;   // svo_result mirrors stereo_matches[0]
    }
    if (P.p.use_ids && !st.has_prev) {           // S3:67,406-407
        long long* ids = P.match_id + pair_idx(slot, cur) * cap;
        const long long b0 = st.last_match_id;
        for (int m = tid; m < base; m += FT) ids[m] = b0 + m;
        __syncthreads();
        if (tid == 0) st.last_match_id = b0 + base;
    }
}

// K7b  win_track_kernel: stage 4, ifmDescWin branch up to the candidate list (S4:435-640), one CTA per slot.
//   shared: pidx[H+1] cidx[H+1] (matches_lr_row_index of the previous / current frame, S3:425-445) | cur[kp_cap]
// Deviation: matches_lr_row_index[H] holds the number of left FEATURES (S3:443), so a window reaching the last image
// row makes the reference read past the current match list; the window end is clamped to the list here.
__global__ void __launch_bounds__(FT)
win_track_kernel(Pools P) {
    extern __shared__ uint32_t s_dyn[];
    __shared__ int s_warp[FT / 32];
    const int H = P.p.H, W = P.p.W;
    int* pidx = reinterpret_cast<int*>(s_dyn);
    int* cidx = pidx + (H + 1);
    uint32_t* curm = s_dyn + 2 * (H + 1);
    const int slot = P.slots[blockIdx.x];
    const SlotState& st = P.state[slot];
    const int cur = st.cur, prev = cur ^ 1;
    const size_t cap = (size_t)P.p.kp_cap;
    const int tid = threadIdx.x;
    const int Mc = P.n_match[pair_idx(slot, cur)];
    const int Mp = st.has_prev ? P.n_match[pair_idx(slot, prev)] : 0;
    const int4* __restrict__ mp = P.match + pair_idx(slot, prev) * cap;
    const int4* __restrict__ mc = P.match + pair_idx(slot, cur) * cap;
    const size_t fpL = frame_idx(slot, prev, 0) * cap, fpR = frame_idx(slot, prev, 1) * cap;
    const size_t fcL = frame_idx(slot, cur, 0) * cap, fcR = frame_idx(slot, cur, 1) * cap;
    const uint4* __restrict__ pD = P.desc + frame_idx(slot, prev, 0) * cap * 2;
    const uint4* __restrict__ cD = P.desc + frame_idx(slot, cur, 0) * cap * 2;

    // matches_lr_row_index[y] = number of matches whose left keypoint lies above row y; [H] = number of left features
    for (int y = tid; y <= H; y += FT) {
        pidx[y] = y == H ? P.n_kp[frame_idx(slot, prev, 0)]
                         : count_rows_le_f([&](int k) { return kp_xy(P, fpL, mp[k].x).y; }, Mp, y - 1);
        cidx[y] = y == H ? P.n_kp[frame_idx(slot, cur, 0)]
                         : count_rows_le_f([&](int k) { return kp_xy(P, fcL, mc[k].x).y; }, Mc, y - 1);
    }
    for (int k = tid; k < Mc; k += FT) curm[k] = 0xFFFFFFFFu;
    __syncthreads();

    const int win_w = P.p.ifm_win_w, win_h = P.p.ifm_win_h;
    for (int pi = tid; pi < Mp; pi += FT) {
        const int4 m = mp[pi];
        const float2 a = kp_xy(P, fpL, m.x), b = kp_xy(P, fpR, m.y);
        // the loop row that owns this match: pidx[y] <= pi < pidx[y + 1], i.e. the first y with pt.y <= y (S3:436) -- the row
        // itself for integer coordinates, the next one for dmORB's fractional ones
        const int y = (int)ceilf(a.y);
        if (y >= H - 1) continue;                                         // S4:512: rows 0 .. H-2
        const int p0 = pidx[y], p1 = pidx[y + 1];
        if (pi < p0 || pi >= p1) continue;
        const int wy_min = max(0, y - win_w), wy_max = min(H - 1, y + win_w);     // sic: the vertical window uses win_w
        const int c0 = cidx[wy_min], c1 = min(cidx[wy_max + 1], Mc);
        if (c1 <= c0) continue;
        const float axf = a.x, bxf = b.x;
        const int xl0 = max(0, (int)__fsub_rn(axf, (float)win_h)), xl1 = min(W - 1, (int)__fadd_rn(axf, (float)win_h));
        const int xr0 = max(0, (int)__fsub_rn(bxf, (float)win_h)), xr1 = min(W - 1, (int)__fadd_rn(bxf, (float)win_h));
        unsigned best = 255u, best_ci = 0xFFFFu;
        for (int ci = c0; ci < c1; ++ci) {
            const int4 c = mc[ci];
            const float fx = kp_xy(P, fcL, c.x).x, gx = kp_xy(P, fcR, c.y).x;
            if (fx < (float)xl0 || fx > (float)xl1 || gx < (float)xr0 || gx > (float)xr1) continue;
            const unsigned d = hamming256_u8(pD + 2 * (size_t)m.x, cD + 2 * (size_t)c.x);   // left descriptors only (S4:585-603)
            if (d < best) { best = d; best_ci = (unsigned)ci; }
        }
        // S4:620-634: a current match keeps the previous match with the smallest distance, the first one among equals
        if (best_ci != 0xFFFFu) atomicMin(&curm[best_ci], (best << 16) | (unsigned)pi);
    }
    __syncthreads();

    // S4:638-670: candidates in current-match order
    int4* __restrict__ out = P.tcand + (size_t)slot * cap;
    int base = 0;
    for (int c0 = 0; c0 < Mc; c0 += FT) {
        const int ci = c0 + tid;
        const int keep = (ci < Mc && curm[ci] != 0xFFFFFFFFu) ? 1 : 0;
        int total;
        const int off = scan256(keep, s_warp, total);
        if (keep) out[base + off] = make_int4((int)(curm[ci] & 0xFFFFu), ci, ci, 0);
        base += total;
    }
    if (tid == 0) {
        P.n_tcand[slot] = base;
        P.n_inl[slot * 2 + 0] = 0;     // masks default to "not computed" (S4:693-713)
        P.n_inl[slot * 2 + 1] = 0;
    }
}

// ------------------------------------------------------------------------------------------------
// K9  consistency_kernel: one CTA per slot.  S4:243-305, S4:743-752, P:326-330, P:348-352.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FT)
consistency_kernel(Pools P) {
    __shared__ int s_warp[FT / 32];
    extern __shared__ uint32_t s_tr[];   // tracked flag per current match (ID management)
    const int slot = P.slots[blockIdx.x];
    SlotState& st = P.state[slot];
    Result& res = P.result[slot];
    const int cur = st.cur, prev = cur ^ 1;
    const size_t cap = (size_t)P.p.kp_cap;
    const int tid = threadIdx.x;
    if (!st.has_prev) {                      // first iteration (P:348-352)
        if (tid == 0) { res.error_code = 4; res.valid = 0; P.n_tracked[slot] = 0; P.oct_stats[stat_idx(slot, P.octave, 3)] = 0; }
        return;
    }
    const int n = P.n_tcand[slot];
    const int4* __restrict__ tc = P.tcand + (size_t)slot * cap;
    const uint8_t* __restrict__ inlL = P.inl + ((size_t)slot * 2 + 0) * cap;
    const uint8_t* __restrict__ inlR = P.inl + ((size_t)slot * 2 + 1) * cap;
    const bool use_masks = P.p.fundmat_filter && P.n_inl[slot * 2] >= 8 && P.n_inl[slot * 2 + 1] >= 8;  // S4:205,240,243
    int2* __restrict__ out = P.tracked + (size_t)slot * cap;
    const int Mc = P.n_match[pair_idx(slot, cur)];
    const bool ids = P.p.use_ids != 0;
    long long* idc = P.match_id + pair_idx(slot, cur) * cap;
    const long long* idp = P.match_id + pair_idx(slot, prev) * cap;
    if (ids) {
        for (int i = tid; i < Mc; i += FT) s_tr[i] = 0;
        __syncthreads();
    }
    int base = 0;
    for (int k0 = 0; k0 < n; k0 += FT) {
        const int k = k0 + tid;
        int keep = 0;
        int4 c = make_int4(0, 0, 0, 0);
        if (k < n) {
            c = tc[k];
            keep = 1;
            if (use_masks && (inlL[k] == 0 || inlR[k] == 0)) keep = 0;
            if (c.y != c.z) keep = 0;        // S4:282 (the query indices are equal by construction)
        }
        int total;
        const int off = scan256(keep, s_warp, total);
        if (keep) {
            out[base + off] = make_int2(c.x, c.y);
            if (ids) { idc[c.y] = idp[c.x]; s_tr[c.y] = 1; }   // S4:288-292
        }
        base += total;
    }
    __syncthreads();
    int kf = 0;
    if (ids) {
        // S4:300-305: untracked current matches get new IDs in index order
        long long b0 = st.last_match_id;
        int nb = 0;
        for (int k0 = 0; k0 < Mc; k0 += FT) {
            const int k = k0 + tid;
            const int fresh = (k < Mc && !s_tr[k]) ? 1 : 0;
            int total;
            const int off = scan256(fresh, s_warp, total);
            if (fresh) idc[k] = b0 + nb + off;
            nb += total;
        }
        __syncthreads();
        // S4:747-751
        int cnt = 0;
        for (int k = tid; k < Mc; k += FT) cnt += (idc[k] <= st.last_kf_max_id) ? 1 : 0;
        int total;
        scan256(cnt, s_warp, total);
        kf = total;
        if (tid == 0) st.last_match_id = b0 + nb;
    }
    if (tid == 0) {
        P.n_tracked[slot] = base;
        P.oct_stats[stat_idx(slot, P.octave, 3)] = base;
        // S4:743-752: both counters run over all octaves; begin_step_kernel zeroed them
        const int total = res.tracked_frame + base;
        res.tracked_frame = total;
        res.tracked_kf += kf;
        if (P.octave == P.n_octaves - 1 && total < P.p.bad_tracking_th) {    // P:326-330, after the last octave
            st.m_error = 5;
            res.error_code = 5;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K4b  reset_ids_octaves_kernel (nOctaves > 1 only): P:254-267.  After resetIds() the previous frame's match IDs of
// ALL octaves are renumbered 0..N-1 in octave order, the generator restarts behind them and the previous frame becomes
// the key frame.  Runs before the per-octave stage 3, so lr_filter_kernel never sees the flag in a multi-octave step.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FT)
reset_ids_octaves_kernel(Pools P, OctaveList O) {
    const int slot = P.slots[blockIdx.x];
    SlotState& st = P.state[slot];
    if (!(st.reset_ids && st.has_prev)) return;          // the whole CTA decides alike: the flag is written after the barrier
    const int prev = st.cur ^ 1;
    long long base = 0;
    for (int k = 0; k < O.n; ++k) {
        const int Mp = P.p.use_ids ? O.n_match[k][pair_idx(slot, prev)] : 0;   // matches_IDs is empty without ID control
        long long* idp = O.match_id[k] + pair_idx(slot, prev) * (size_t)O.cap[k];
        for (int m = threadIdx.x; m < Mp; m += FT) idp[m] = base + m;
        base += Mp;
    }
    __syncthreads();
    if (threadIdx.x == 0) { st.last_match_id = base; st.last_kf_max_id = base - 1; st.reset_ids = 0; }
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
static int tf_smem(const Params& p) { return 3 * p.kp_cap * (int)sizeof(uint32_t) + ((p.kp_cap + 3) & ~3); }
static int rbr_smem(const Params& p) { return (2 * p.H + 2 * p.kp_cap) * (int)sizeof(uint32_t); }
static int win_smem(const Params& p) { return (2 * (p.H + 1) + p.kp_cap) * (int)sizeof(uint32_t); }

// a batch too small to fill the GPU with one CTA per query block (the literal one-estimator seam): split the train set so
// that a few hundred CTAs share the scan; the three Hamming problems of one pair at batch 1: 0.19 -> 0.04 ms (tools/latency_breakdown.py)
static int hamming_splits(int ctas) {
    int s = 1;
    while (s < 16 && ctas * s * 2 <= 296) s *= 2;
    return s;
}

cudaError_t launch_match_lr(const Pools& P, int n, cudaStream_t st, int* launches, StepTimer* tm) {
    if (P.p.match_method == 1) {            // smDescRbR: one band-limited kernel instead of brute force + filter
        rbr_match_kernel<<<n, FT, rbr_smem(P.p), st>>>(P);
        if (tm) { tm->mark(st); tm->mark(st); }
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    const int per = HM_THREADS * HM_QPT;
    dim3 g((P.p.kp_cap + per - 1) / per, n, hamming_splits((P.p.kp_cap + per - 1) / per * n));
    if (g.z > 1) hamming_init_kernel<<<dim3((P.p.kp_cap + 255) / 256, n), 256, 0, st>>>(P, 0, 1);
    hamming_pipeline_kernel<<<g, HM_THREADS, 0, st>>>(P, 0, 1);
    if (tm) tm->mark(st);
    lr_filter_kernel<<<n, FT, P.p.kp_cap * sizeof(uint32_t), st>>>(P);
    if (tm) tm->mark(st);
    if (launches) *launches += 2 + (g.z > 1 ? 1 : 0);
    return cudaGetLastError();
}

cudaError_t launch_track_match(const Pools& P, int n, cudaStream_t st, int* launches, StepTimer* tm) {
    if (P.p.ifm_method == 1) {              // ifmDescWin
        win_track_kernel<<<n, FT, win_smem(P.p), st>>>(P);
        if (tm) { tm->mark(st); tm->mark(st); }
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    const int per = HM_THREADS * HM_QPT;
    dim3 g((P.p.kp_cap + per - 1) / per, 2 * n, hamming_splits((P.p.kp_cap + per - 1) / per * 2 * n));
    if (g.z > 1) hamming_init_kernel<<<dim3((P.p.kp_cap + 255) / 256, 2 * n), 256, 0, st>>>(P, 1, 2);
    hamming_pipeline_kernel<<<g, HM_THREADS, 0, st>>>(P, 1, 2);
    if (tm) tm->mark(st);
    track_filter_kernel<<<n, FT, tf_smem(P.p), st>>>(P);
    if (tm) tm->mark(st);
    if (launches) *launches += 2 + (g.z > 1 ? 1 : 0);
    return cudaGetLastError();
}

cudaError_t launch_consistency(const Pools& P, int n, cudaStream_t st, int* launches, StepTimer* tm) {
    consistency_kernel<<<n, FT, P.p.kp_cap * sizeof(uint32_t), st>>>(P);
    if (tm) tm->mark(st);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_reset_ids_octaves(const Pools& P, const OctaveList& O, int n, cudaStream_t st, int* launches) {
    reset_ids_octaves_kernel<<<n, FT, 0, st>>>(P, O);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_hamming_raw(const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int reps, uint32_t* d_out,
                               cudaStream_t st) {
    if (nq <= 0 || reps <= 0) return cudaSuccess;
    const int per = HM_THREADS * HM_QPT;
    dim3 g((nq + per - 1) / per, reps);
    hamming_raw_kernel<<<g, HM_THREADS, 0, st>>>(reinterpret_cast<const uint4*>(d_q), nq,
                                                  reinterpret_cast<const uint4*>(d_t), nt, d_out);
    return cudaGetLastError();
}

cudaError_t configure_match(const Params& p) {
    cudaError_t e = SVO_RAISE_DYN_SMEM(track_filter_kernel, tf_smem(p));
    if (e != cudaSuccess) return e;
    e = SVO_RAISE_DYN_SMEM(lr_filter_kernel, (int)(p.kp_cap * sizeof(uint32_t)));
    if (e != cudaSuccess) return e;
    e = SVO_RAISE_DYN_SMEM(rbr_match_kernel, rbr_smem(p));
    if (e != cudaSuccess) return e;
    e = SVO_RAISE_DYN_SMEM(win_track_kernel, win_smem(p));
    if (e != cudaSuccess) return e;
    return SVO_RAISE_DYN_SMEM(consistency_kernel, (int)(p.kp_cap * sizeof(uint32_t)));
}

}  // namespace svo
