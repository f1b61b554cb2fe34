// introsort_emul.cuh -- reproduces the permutation produced by GNU libstdc++'s std::sort.
//
// Why: the reference orders keypoints with std::sort on tie-heavy keys
// (stage2_detect.cpp:89 by pt.y, :242/:324 by response).  std::sort is not stable, so the order
// of equal keys -- and with it every keypoint index, match index and BFMatcher tie-break
// downstream -- is whatever libstdc++'s introsort happens to produce.  "Bit-exact keypoint and
// match indices" therefore needs the same algorithm, restated here from its published
// structure: median-of-3 pivot moved to the front, unguarded Hoare partition, recursion on the
// right part / loop on the left with a depth limit of 2*floor(log2 n), heap-sort fallback when the
// limit is hit, threshold 16, one final insertion sort.
//
// Elements are 32-bit words  key << KEY_SHIFT | payload ; only the key takes part in comparisons:
//   Desc = true  : comp(a,b) = key(a) > key(b)   (KeypointResponseSorter, response descending)
//   Desc = false : comp(a,b) = key(a) < key(b)   (KpRowSorter, pt.y ascending)
// The payload (original index) rides along, exactly like the index vector the reference sorts.
//
// Two implementations with identical results:
//   introsort_seq      one thread, host or device (also the unit-test reference against std::sort)
//   introsort_warp     one warp cooperates on each partition step (device only, see below)
#pragma once
#include <stdint.h>

#ifndef SVO_THREAD_SEG_MAX
#define SVO_THREAD_SEG_MAX 32
#endif

#if defined(__CUDACC__)
#define SVO_HD __host__ __device__ __forceinline__
#else
#define SVO_HD inline
#endif

namespace svo {

template <bool Desc, int KeyShift>
struct KeyComp {
    SVO_HD bool lt(uint32_t a, uint32_t b) const {
        return Desc ? ((a >> KeyShift) > (b >> KeyShift)) : ((a >> KeyShift) < (b >> KeyShift));
    }
};

// elements are plain indices into a float key array; comp(a,b) = key[a] > key[b]
// (KeypointResponseSorter on arbitrary cv::KeyPoint::response values, stage2_detect.cpp:242)
struct IndexedDescComp {
    const float* key;
    SVO_HD bool lt(uint32_t a, uint32_t b) const { return key[a] > key[b]; }
};

constexpr int kSortThreshold = 16;

SVO_HD int floor_log2(int n) {
    int l = 0;
    while (n > 1) { n >>= 1; ++l; }
    return l;
}

// ---- heap helpers (only reached when the depth limit is exhausted) ---------------------------
template <class C>
SVO_HD void push_heap_(const C& c, uint32_t* a, int hole, int top, uint32_t value) {
    int parent = (hole - 1) / 2;
    while (hole > top && c.lt(a[parent], value)) {
        a[hole] = a[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    a[hole] = value;
}

template <class C>
SVO_HD void adjust_heap_(const C& c, uint32_t* a, int hole, int len, uint32_t value) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (c.lt(a[child], a[child - 1])) child--;
        a[hole] = a[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        a[hole] = a[child - 1];
        hole = child - 1;
    }
    push_heap_(c, a, hole, top, value);
}

template <class C>
SVO_HD void heap_sort_(const C& c, uint32_t* a, int len) {   // partial_sort(first, last, last)
    if (len >= 2) {
        int parent = (len - 2) / 2;
        while (true) {
            uint32_t v = a[parent];
            adjust_heap_(c, a, parent, len, v);
            if (parent == 0) break;
            parent--;
        }
    }
    int last = len;
    while (last > 1) {
        --last;
        uint32_t v = a[last];
        a[last] = a[0];
        adjust_heap_(c, a, 0, last, v);
    }
}

template <class C>
SVO_HD void median_to_first_(const C& c, uint32_t* a, int result, int ia, int ib, int ic) {
    int pick;
    if (c.lt(a[ia], a[ib])) {
        if (c.lt(a[ib], a[ic])) pick = ib;
        else if (c.lt(a[ia], a[ic])) pick = ic;
        else pick = ia;
    } else if (c.lt(a[ia], a[ic])) pick = ia;
    else if (c.lt(a[ib], a[ic])) pick = ic;
    else pick = ib;
    uint32_t t = a[result];
    a[result] = a[pick];
    a[pick] = t;
}

template <class C>
SVO_HD int unguarded_partition_(const C& c, uint32_t* a, int first, int last, uint32_t pivot) {
    while (true) {
        while (c.lt(a[first], pivot)) ++first;
        --last;
        while (c.lt(pivot, a[last])) --last;
        if (!(first < last)) return first;
        uint32_t t = a[first];
        a[first] = a[last];
        a[last] = t;
        ++first;
    }
}

// final insertion sort == a stable sort of the partially ordered array; written as such
template <class C>
SVO_HD void insertion_sort_(const C& c, uint32_t* a, int first, int last) {
    for (int i = first + 1; i < last; ++i) {
        uint32_t v = a[i];
        int j = i;
        while (j > first && c.lt(v, a[j - 1])) {
            a[j] = a[j - 1];
            --j;
        }
        a[j] = v;
    }
}

// std::__introsort_loop on a[first..last) with the given remaining depth, one thread: explicit stack of
// (first, last, depth) replacing the recursion on the right part.  Leaves segments of <= 16 elements unsorted.
// `keep`: only the final positions [0, keep) are wanted (top-K selection).  A segment that starts at or beyond `keep`
// can be left unsorted: partitions never move an element across a segment boundary again, and the final insertion sort
// cannot either (everything right of a boundary is not smaller than everything left of it), so the first `keep`
// positions end up exactly as in the full sort.
template <class C>
SVO_HD void introsort_loop_seq(const C& c, uint32_t* a, int first0, int last0, int depth0, int keep = 0x7FFFFFFF) {
    int st_first[64], st_last[64], st_depth[64];
    int sp = 1;
    st_first[0] = first0; st_last[0] = last0; st_depth[0] = depth0;
    if (first0 >= keep) return;
    while (sp > 0) {
        --sp;
        int first = st_first[sp], last = st_last[sp], depth = st_depth[sp];
        while (last - first > kSortThreshold) {
            if (depth == 0) {
                heap_sort_(c, a + first, last - first);
                break;
            }
            --depth;
            const int mid = first + (last - first) / 2;
            median_to_first_(c, a, first, first + 1, mid, last - 1);
            const int cut = unguarded_partition_(c, a, first + 1, last, a[first]);
            // recurse on [cut, last) (unless it lies beyond the wanted prefix), continue with [first, cut)
            if (cut < keep) { st_first[sp] = cut; st_last[sp] = last; st_depth[sp] = depth; ++sp; }
            last = cut;
        }
    }
}

// Sequential emulation of std::sort.
template <class C>
SVO_HD void introsort_seq_c(const C& c, uint32_t* a, int n) {
    if (n <= 1) return;
    introsort_loop_seq(c, a, 0, n, 2 * floor_log2(n));
    insertion_sort_(c, a, 0, n);
}

// The first `keep` positions of std::sort's result (the rest of the array is left in an unspecified order).
template <class C>
SVO_HD void introsort_prefix_seq_c(const C& c, uint32_t* a, int n, int keep) {
    if (n <= 1) return;
    introsort_loop_seq(c, a, 0, n, 2 * floor_log2(n), keep);
    // final insertion sort restricted to the leaves that reach into the prefix: an element moves at most 15 places
    const int lim = keep < n - kSortThreshold ? keep + kSortThreshold : n;
    insertion_sort_(c, a, 0, lim);
}

// Rank-table formulation of the unguarded Hoare partition of a[first..last) around the pivot at a[first]
// (what the warp- and block-cooperative versions below compute, kept sequential here so that the host unit
// test can pin it against unguarded_partition_):
//   L = positions x in (first, last), ascending,  with !comp(a[x], pivot)       ("left stoppers")
//   R = positions x in [first, last), descending, with !comp(pivot, a[x])       ("right stoppers"; x = first always is)
//   K = #{k : L[k] < R[k]}  (a prefix, L ascends and R descends);  the pairs k < K are swapped;
//   cut = min(L[K], R[K-1])  (the left scan of the last round stops at the next original stopper or at the
//   stopper the previous swap moved to R[K-1], whichever comes first).
template <class C>
SVO_HD int ranked_partition_seq(const C& c, uint32_t* a, int first, int last, int* posL, int* posR) {
    const uint32_t pivot = a[first];
    int nl = 0, nr = 0;
    for (int x = first + 1; x < last; ++x) if (!c.lt(a[x], pivot)) posL[nl++] = x;
    for (int x = last - 1; x >= first; --x) if (!c.lt(pivot, a[x])) posR[nr++] = x;
    int K = 0;
    const int kmax = nl < nr ? nl : nr;
    while (K < kmax && posL[K] < posR[K]) {
        const uint32_t t = a[posL[K]]; a[posL[K]] = a[posR[K]]; a[posR[K]] = t;
        ++K;
    }
    int cut = (K < nl) ? posL[K] : last;
    if (K > 0 && posR[K - 1] < cut) cut = posR[K - 1];
    return cut;
}

// introsort_seq_c with the partition in its rank-table formulation (host unit test only)
template <class C>
SVO_HD void introsort_seq_ranked_c(const C& c, uint32_t* a, int n, int* posL, int* posR) {
    if (n <= 1) return;
    int st_first[64], st_last[64], st_depth[64];
    int sp = 0;
    st_first[0] = 0; st_last[0] = n; st_depth[0] = 2 * floor_log2(n);
    sp = 1;
    while (sp > 0) {
        --sp;
        int first = st_first[sp], last = st_last[sp], depth = st_depth[sp];
        while (last - first > kSortThreshold) {
            if (depth == 0) { heap_sort_(c, a + first, last - first); break; }
            --depth;
            median_to_first_(c, a, first, first + 1, first + (last - first) / 2, last - 1);
            const int cut = ranked_partition_seq(c, a, first, last, posL, posR);
            st_first[sp] = cut; st_last[sp] = last; st_depth[sp] = depth;
            ++sp;
            last = cut;
        }
    }
    insertion_sort_(c, a, 0, n);
}

template <bool Desc, int KeyShift>
SVO_HD void introsort_seq(uint32_t* a, int n) {
    introsort_seq_c(KeyComp<Desc, KeyShift>(), a, n);
}

#if defined(__CUDACC__)
// ------------------------------------------------------------------------------------------------
// Block-cooperative emulation (device only): identical result to introsort_seq.
//
//  * The recursion tree is walked level by level: all segments of one level are independent, each is
//    handled by one warp, the children (> 16 elements) go to the next level's list.  Both children of a
//    partition inherit depth-1, exactly like the recursion on [cut,last) + loop on [first,cut).
//  * The unguarded Hoare partition is executed 32 elements at a time: the warp loads a chunk from each
//    end, ballots the "stoppers" (left: !comp(x,pivot), right: !comp(pivot,x)), pairs the k-th left
//    stopper with the k-th right stopper and swaps the pairs with i < j in one step.  The pairing
//    sequence, and therefore every swap and the returned cut, is the one the sequential loop produces
//    (validated exhaustively against the sequential partition, tests/test_host_cpu.py + GPU tests).
//  * The final insertion sort is a stable sort that only moves elements inside their <= 16-element
//    leaf segment, so an element's final position is  i - #{j in [i-15,i): comp(a[i],a[j])}
//    + #{j in (i,i+15]: comp(a[j],a[i])}: computed by all threads at once.
// ------------------------------------------------------------------------------------------------
// Partition a[seg_first .. seg_last) around the pivot already placed at a[seg_first]; all 32 lanes call it.
// s_pos: 64 ints of shared scratch private to the calling warp (rank -> position tables of one chunk pair).
template <class C>
__device__ int warp_unguarded_partition(const C& c, uint32_t* a, int seg_first, int seg_last, int* s_pos) {
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t pivot = a[seg_first];
    int* posL = s_pos;          // posL[k] = position of the k-th pending left stopper
    int* posR = s_pos + 32;     // posR[k] = position of the k-th pending right stopper
    int nl = 0, nr = 0;         // pending stoppers; the tables are consumed from index kl / kr upwards
    int kl = 0, kr = 0;
    int lb = seg_first + 1 - 32, rb = seg_last - 1 + 32;
    int lastR = 0x7FFFFFFF;
    while (true) {
        if (kl == nl) {          // left table exhausted: next chunk
            lb += 32;
            const int x = lb + lane;
            const bool stop = (x >= seg_last) ? true : !c.lt(a[x], pivot);
            const unsigned m = __ballot_sync(0xffffffffu, stop);
            if (stop) posL[__popc(m & lt_mask)] = x;
            nl = __popc(m);
            kl = 0;
        }
        if (kr == nr) {
            rb -= 32;
            const int x = rb - lane;
            const bool stop = (x < seg_first) ? true : !c.lt(pivot, a[x]);
            const unsigned m = __ballot_sync(0xffffffffu, stop);
            if (stop) posR[__popc(m & lt_mask)] = x;
            nr = __popc(m);
            kr = 0;
        }
        __syncwarp();
        if (nl == 0 || nr == 0) continue;
        const int cnt = min(nl - kl, nr - kr);
        int i = 0, j = 0;
        bool ok = false;
        if (lane < cnt) {
            i = posL[kl + lane];
            j = posR[kr + lane];
            ok = i < j;
        }
        const unsigned okmask = __ballot_sync(0xffffffffu, ok);
        const int m = __popc(okmask);            // the pairs with i < j form a prefix
        if (ok) { const uint32_t t = a[i]; a[i] = a[j]; a[j] = t; }
        if (m > 0) lastR = posR[kr + m - 1];
        if (m < cnt) {
            const int Lm = posL[kl + m];
            __syncwarp();
            return min(Lm, lastR);
        }
        kl += cnt;
        kr += cnt;
        __syncwarp();
    }
}


// Block-wide partition for the few large segments at the top of the recursion (one warp would walk them 32
// elements at a time while the other warps wait): every thread classifies elements, the stoppers are ranked with
// ballots + per-warp counts, the rank -> position tables go to scratch (16-bit positions) and the pairs k < K are
// swapped in parallel.  Same result as ranked_partition_seq.  All threads of the block must call it.
//   posL / posR: room for (last - first) uint16 each;  s_blk: 2 * nwarps + 1 ints of shared memory.
constexpr int kBlockPartMin = 1024;      // segments at least this long are partitioned by the whole block
constexpr int kBlockPartMaxList = 16;    // ... on levels with at most this many segments
constexpr int kThreadSegMax = SVO_THREAD_SEG_MAX;   // segments up to this long are finished by ONE thread each (sequential
                                         // std::__introsort_loop): near the leaves there are hundreds of independent
                                         // segments, and a warp per 20..100 elements wastes most of its lanes

template <class C>
__device__ int block_unguarded_partition(const C& c, uint32_t* a, int first, int last, uint16_t* posL, uint16_t* posR,
                                         int* s_blk) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    int* s_cl = s_blk;
    int* s_cr = s_blk + nwarps;
    int* s_K = s_blk + 2 * nwarps;
    const uint32_t pivot = a[first];
    const int per = ((last - first + nwarps * 32 - 1) / (nwarps * 32)) * 32;   // elements per warp, multiple of 32
    const int wbeg = first + warp * per, wend = min(last, wbeg + per);
    // A. count the stoppers of this warp's range
    int cl = 0, cr = 0;
    for (int x0 = wbeg; x0 < wend; x0 += 32) {
        const int x = x0 + lane;
        const bool in = x < wend;
        const uint32_t v = in ? a[x] : 0u;
        cl += __popc(__ballot_sync(0xffffffffu, in && x > first && !c.lt(v, pivot)));
        cr += __popc(__ballot_sync(0xffffffffu, in && !c.lt(pivot, v)));
    }
    if (lane == 0) { s_cl[warp] = cl; s_cr[warp] = cr; }
    if (tid == 0) *s_K = 0;
    __syncthreads();
    int offL = 0, offR = 0, nl = 0, nr = 0;
    for (int w = 0; w < nwarps; ++w) {
        const int l = s_cl[w], r = s_cr[w];
        if (w < warp) offL += l;
        if (w > warp) offR += r;
        nl += l; nr += r;
    }
    // B. rank -> position tables: left stoppers ascending, right stoppers descending
    for (int x0 = wbeg; x0 < wend; x0 += 32) {
        const int x = x0 + lane;
        const bool stop = x < wend && x > first && !c.lt(a[x], pivot);
        const unsigned m = __ballot_sync(0xffffffffu, stop);
        if (stop) posL[offL + __popc(m & lt_mask)] = (uint16_t)x;
        offL += __popc(m);
    }
    for (int x0 = wend - 1; x0 >= wbeg; x0 -= 32) {
        const int x = x0 - lane;
        const bool stop = x >= wbeg && !c.lt(pivot, a[x]);
        const unsigned m = __ballot_sync(0xffffffffu, stop);
        if (stop) posR[offR + __popc(m & lt_mask)] = (uint16_t)x;
        offR += __popc(m);
    }
    __syncthreads();
    // C. swap the pairs with L[k] < R[k] (a prefix of the pair list); no position takes part in two swaps
    const int kmax = min(nl, nr);
    int cnt = 0;
    for (int k0 = 0; k0 < kmax; k0 += blockDim.x) {
        const int k = k0 + tid;
        bool ok = false;
        int i = 0, j = 0;
        if (k < kmax) { i = posL[k]; j = posR[k]; ok = i < j; }
        if (ok) { const uint32_t t = a[i]; a[i] = a[j]; a[j] = t; }
        cnt += __popc(__ballot_sync(0xffffffffu, ok));
    }
    if (lane == 0 && cnt) atomicAdd(s_K, cnt);
    __syncthreads();
    const int K = *s_K;
    int cut = (K < nl) ? (int)posL[K] : last;
    if (K > 0) cut = min(cut, (int)posR[K - 1]);
    __syncthreads();       // s_blk is reused by the next call
    return cut;
}

// children of a partition: short ones go to the per-thread list, the others to the next level
__device__ __forceinline__ void push_segment(int first, int last, int depth, uint2* out, int* cnt_out, uint2* small, int* cnt_small,
                                             int keep) {
    if (last - first <= kSortThreshold || first >= keep) return;     // beyond the wanted prefix: see introsort_loop_seq
    const uint2 e = make_uint2((unsigned)first, (unsigned)last | ((unsigned)depth << 24));
    if (last - first <= kThreadSegMax) small[atomicAdd(cnt_small, 1)] = e;
    else out[atomicAdd(cnt_out, 1)] = e;
}

// Sort a[0..n) (shared memory) with all threads of the block.  listA/listB/listS: room for n/16 + 2 entries each,
// s_cnt: 3 ints (shared), tmp: n words of scratch (global or shared), s_pos: 64 ints per warp.
// keep: only positions [0, keep) of the result are needed (pass n for a full sort); the rest of a[] is left unspecified.
template <class C>
__device__ void introsort_block(const C& c, uint32_t* a, int n, uint2* listA, uint2* listB, uint2* listS, int* s_cnt, uint32_t* tmp,
                                int* s_pos /* 64 ints per warp, shared */, int keep) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    if (tid == 0) {
        s_cnt[0] = s_cnt[1] = s_cnt[2] = 0;
        push_segment(0, n, 2 * floor_log2(max(n, 1)), listA, &s_cnt[0], listS, &s_cnt[2], keep);
    }
    __syncthreads();
    int cur = 0;
    while (true) {
        const int cntA = s_cnt[cur];
        if (cntA == 0) break;
        uint2* in = cur ? listB : listA;
        uint2* out = cur ? listA : listB;
        // the few large segments at the top of the recursion: whole block, one after the other
        const bool top = cntA <= kBlockPartMaxList && n < 65536;
        if (top) {
            for (int sidx = 0; sidx < cntA; ++sidx) {
                const uint2 e = in[sidx];
                const int first = (int)e.x, last = (int)(e.y & 0xFFFFFFu);
                const int depth = (int)(e.y >> 24);
                if (depth == 0 || last - first < kBlockPartMin) continue;
                if (tid == 0) median_to_first_(c, a, first, first + 1, first + (last - first) / 2, last - 1);
                __syncthreads();
                uint16_t* pos16 = reinterpret_cast<uint16_t*>(tmp);
                const int cut = block_unguarded_partition(c, a, first, last, pos16, pos16 + n, s_pos);
                if (tid == 0) {
                    push_segment(cut, last, depth - 1, out, &s_cnt[cur ^ 1], listS, &s_cnt[2], keep);
                    push_segment(first, cut, depth - 1, out, &s_cnt[cur ^ 1], listS, &s_cnt[2], keep);
                }
            }
        }
        int small_idx = 0;
        for (int sidx = top ? 0 : warp; sidx < cntA; sidx += top ? 1 : nwarps) {
            const uint2 e = in[sidx];
            const int first = (int)e.x, last = (int)(e.y & 0xFFFFFFu);
            int depth = (int)(e.y >> 24);
            if (top) {
                if (depth != 0 && last - first >= kBlockPartMin) continue;      // done by the block above
                if ((small_idx++ % nwarps) != warp) continue;
            }
            if (depth == 0) {                          // depth limit hit: heap sort, no children
                if (lane == 0) heap_sort_(c, a + first, last - first);
                __syncwarp();
                continue;
            }
            --depth;
            if (lane == 0) median_to_first_(c, a, first, first + 1, first + (last - first) / 2, last - 1);
            __syncwarp();
            const int cut = warp_unguarded_partition(c, a, first, last, s_pos + 64 * warp);
            if (lane == 0) {
                push_segment(cut, last, depth, out, &s_cnt[cur ^ 1], listS, &s_cnt[2], keep);
                push_segment(first, cut, depth, out, &s_cnt[cur ^ 1], listS, &s_cnt[2], keep);
            }
        }
        __syncthreads();
        if (tid == 0) s_cnt[cur] = 0;
        cur ^= 1;
        __syncthreads();
    }
    // the short segments: one thread each, sequential std::__introsort_loop
    {
        const int ns = s_cnt[2];
        for (int sidx = tid; sidx < ns; sidx += blockDim.x) {
            const uint2 e = listS[sidx];
            introsort_loop_seq(c, a, (int)e.x, (int)(e.y & 0xFFFFFFu), (int)(e.y >> 24));
        }
        __syncthreads();
    }
    // final insertion sort == stable placement inside the <= 16-element leaves (those that reach into the prefix)
    const int lim = keep < n - kSortThreshold ? keep + kSortThreshold : n;
    for (int i = tid; i < lim; i += blockDim.x) {
        const uint32_t v = a[i];
        int pos = i;
        const int lo = max(0, i - (kSortThreshold - 1)), hi = min(n - 1, i + (kSortThreshold - 1));
        for (int j = lo; j < i; ++j) pos -= c.lt(v, a[j]) ? 1 : 0;
        for (int j = i + 1; j <= hi; ++j) pos += c.lt(a[j], v) ? 1 : 0;
        tmp[pos] = v;
    }
    __syncthreads();
    for (int i = tid; i < min(lim, keep); i += blockDim.x) a[i] = tmp[i];
    __syncthreads();
}
#endif  // __CUDACC__

// Stable order (sort_mode = 1): plain insertion-free reference used for diagnostics.
template <bool Desc, int KeyShift>
SVO_HD void stable_sort_seq(uint32_t* a, int n) {
    insertion_sort_(KeyComp<Desc, KeyShift>(), a, 0, n);
}

}  // namespace svo
