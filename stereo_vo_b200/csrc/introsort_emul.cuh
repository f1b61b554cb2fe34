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

// Sequential emulation.  `stack` needs room for 2*64 ints (ranges pending on the explicit stack).
template <class C>
SVO_HD void introsort_seq_c(const C& c, uint32_t* a, int n) {
    if (n <= 1) return;
    // explicit stack of (first, last, depth) replacing the recursion on the right part
    int st_first[64], st_last[64], st_depth[64];
    int sp = 0;
    st_first[0] = 0; st_last[0] = n; st_depth[0] = 2 * floor_log2(n);
    sp = 1;
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
            // recurse on [cut, last), continue with [first, cut)
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

// Stable order (sort_mode = 1): plain insertion-free reference used for diagnostics.
template <bool Desc, int KeyShift>
SVO_HD void stable_sort_seq(uint32_t* a, int n) {
    insertion_sort_(KeyComp<Desc, KeyShift>(), a, 0, n);
}

}  // namespace svo
