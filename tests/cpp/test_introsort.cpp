// Host unit test: stereo_vo_b200/csrc/introsort_emul.cuh (sequential emulation) must reproduce the
// permutation of libstdc++'s std::sort for the comparators the reference uses
// (stage2_detect.cpp:89 KpRowSorter, :242/:324 KeypointResponseSorter), including the
// heap-sort fallback (forced with a median-of-3 killer sequence).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>
#include "../../stereo_vo_b200/csrc/introsort_emul.cuh"

static long g_fail = 0, g_cases = 0;

template <bool Desc, int Shift>
static void check(const std::vector<uint32_t>& keys, const char* what) {
    const int n = (int)keys.size();
    std::vector<size_t> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    if (Desc) std::sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return keys[a] > keys[b]; });
    else      std::sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return keys[a] < keys[b]; });
    std::vector<uint32_t> w(n);
    for (int i = 0; i < n; ++i) w[i] = (keys[i] << Shift) | (uint32_t)i;
    svo::introsort_seq<Desc, Shift>(w.data(), n);
    // the rank-table formulation of the partition (what the warp / block kernels compute) must agree too
    std::vector<uint32_t> w2(n);
    for (int i = 0; i < n; ++i) w2[i] = (keys[i] << Shift) | (uint32_t)i;
    std::vector<int> pl(n + 1), pr(n + 1);
    svo::introsort_seq_ranked_c(svo::KeyComp<Desc, Shift>(), w2.data(), n, pl.data(), pr.data());
    if (w2 != w) {
        if (g_fail < 5) std::printf("MISMATCH (ranked partition) %s n=%d\n", what, n);
        ++g_fail;
        return;
    }
    // top-K variant: the first `keep` positions must equal the full sort's for every keep
    for (int keep : {1, 7, 16, 17, n / 3, n / 2, n - 1, n}) {
        if (keep < 1 || keep > n) continue;
        std::vector<uint32_t> w3(n);
        for (int i = 0; i < n; ++i) w3[i] = (keys[i] << Shift) | (uint32_t)i;
        svo::introsort_prefix_seq_c(svo::KeyComp<Desc, Shift>(), w3.data(), n, keep);
        if (!std::equal(w3.begin(), w3.begin() + keep, w.begin())) {
            if (g_fail < 5) std::printf("MISMATCH (prefix %d) %s n=%d\n", keep, what, n);
            ++g_fail;
            return;
        }
    }
    ++g_cases;
    for (int i = 0; i < n; ++i)
        if ((w[i] & ((1u << Shift) - 1)) != idx[i]) {
            if (g_fail < 5) std::printf("MISMATCH %s n=%d at %d: got %u want %zu\n", what, n, i, w[i] & ((1u << Shift) - 1), idx[i]);
            ++g_fail;
            return;
        }
}

// median-of-3 killer (Musser): drives quicksort to quadratic depth -> heap-sort fallback
static std::vector<uint32_t> killer(int n) {
    std::vector<uint32_t> v(n);
    const int k = n / 2;
    for (int i = 1; i <= k; ++i) {
        if (i % 2 == 1) { v[i - 1] = i; v[i] = k + i; }
        v[k + i - 1] = 2 * i;
    }
    return v;
}

int main() {
    std::mt19937 rng(1234);
    for (int rep = 0; rep < 400; ++rep) {
        const int n = (rep < 40) ? rep : (int)(rng() % 9000) + 1;
        const int distinct = 1 + (int)(rng() % (rep % 3 == 0 ? 6 : 60));
        std::vector<uint32_t> k(n);
        for (auto& x : k) x = 20 + rng() % distinct;
        check<true, 24>(k, "resp-desc");
        for (auto& x : k) x = 31 + rng() % (distinct * 8);
        check<false, 16>(k, "row-asc");
    }
    // sorted / reversed / all-equal / organ-pipe inputs
    for (int n : {17, 33, 100, 1000, 5000}) {
        std::vector<uint32_t> k(n);
        for (int i = 0; i < n; ++i) k[i] = i % 251;
        check<true, 24>(k, "saw"); check<false, 16>(k, "saw");
        for (int i = 0; i < n; ++i) k[i] = 7;
        check<true, 24>(k, "equal"); check<false, 16>(k, "equal");
        for (int i = 0; i < n; ++i) k[i] = std::min(i, n - i) % 250;
        check<true, 24>(k, "pipe"); check<false, 16>(k, "pipe");
    }
    for (int n : {64, 256, 1024, 4096, 8192}) {
        std::vector<uint32_t> k = killer(n);
        check<false, 16>(k, "killer-asc");
        for (auto& x : k) x = 65535 - x;
        check<false, 16>(k, "killer-rev");
        std::vector<uint32_t> kk = killer(n);
        for (auto& x : kk) x = x % 200;  // 8-bit keys, ties
        check<true, 24>(kk, "killer-desc-ties");
    }
    std::printf("cases=%ld failures=%ld\n", g_cases, g_fail);
    return g_fail ? 1 : 0;
}
