"""CPU checks of two exactness arguments the CUDA kernels rely on (the kernels themselves are checked against the
oracle in the -m gpu tests; here the ARGUMENT is checked on adversarial inputs that real images rarely produce):

  * track_filter_kernel reaches the reference's sequential greedy collision filter (stage4_match_consecutive.cpp:145-160)
    as a fixed point of parallel rounds;
  * the pipeline's Hamming kernels finish a distance only if its 128-bit prefix beats the best so far, with the best
    initialised at orb_th + 1: every match at or below the threshold is unchanged (distance and lowest-index tie-break).
"""
import numpy as np


def _greedy_sequential(a, b, ok):
    claimed_l, claimed_r, kept = set(), set(), []
    for k in range(len(a)):
        if not ok[k] or a[k] in claimed_l or b[k] in claimed_r:
            continue
        claimed_l.add(a[k])
        claimed_r.add(b[k])
        kept.append(k)
    return kept


def _greedy_rounds(a, b, ok, max_rounds=12):
    """the rounds of track_filter_kernel (csrc/match.cu), one list at a time"""
    n = len(a)
    U, K, R = 0, 1, 2
    st = np.where(ok, U, R)
    undecided = False
    for _ in range(max_rounds):
        head_l, head_r = {}, {}
        for k in range(n):
            if st[k] != R:
                head_l[a[k]] = min(head_l.get(a[k], n), k)
                head_r[b[k]] = min(head_r.get(b[k], n), k)
        for k in range(n):
            if st[k] == U and head_l[a[k]] == k and head_r[b[k]] == k:
                st[k] = K
        new = st.copy()
        undecided = False
        for k in range(n):
            if st[k] == U:
                hl, hr = head_l[a[k]], head_r[b[k]]
                if (hl < k and st[hl] == K) or (hr < k and st[hr] == K):
                    new[k] = R
                else:
                    undecided = True
        st = new
        if not undecided:
            break
    if undecided:                      # the kernel's sequential tail
        cl, cr = set(), set()
        for k in range(n):
            if st[k] == U:
                st[k] = R if (a[k] in cl or b[k] in cr) else K
            if st[k] == K:
                cl.add(a[k])
                cr.add(b[k])
    return [k for k in range(n) if st[k] == K]


def test_parallel_greedy_filter_equals_sequential_walk():
    rng = np.random.default_rng(0)
    for trial in range(600):
        n = int(rng.integers(1, 160))
        m = int(rng.integers(1, max(2, n // int(rng.integers(1, 6)))))
        a, b, ok = rng.integers(0, m, n), rng.integers(0, m, n), rng.random(n) < 0.8
        if trial % 10 == 0:            # alternating collision chain: the worst case for the number of rounds
            a = np.repeat(np.arange(n // 2 + 1), 2)[:n]
            b = np.repeat(np.arange(n // 2 + 1), 2)[1:n + 1]
            ok = np.ones(n, bool)
        assert _greedy_rounds(a, b, ok) == _greedy_sequential(a, b, ok), trial


def _popcount_rows(x):
    return np.unpackbits(x, axis=-1).sum(-1)


def test_prefix_pruned_argmin_keeps_every_match_below_the_threshold():
    rng = np.random.default_rng(1)
    th = 40
    for trial in range(20):
        nq, nt = int(rng.integers(1, 60)), int(rng.integers(1, 80))
        t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
        q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
        for i in range(nq):            # plant near matches, duplicates and threshold cases
            j = int(rng.integers(0, nt))
            nflip = int(rng.choice([0, 1, 10, th - 1, th, th + 1, 60]))
            d = t[j].copy()
            for bit in rng.choice(256, nflip, replace=False):
                d[bit >> 3] ^= np.uint8(1 << (bit & 7))
            q[i] = d
            if rng.random() < 0.3 and j + 1 < nt:
                t[j + 1] = t[j]        # exact duplicate with a higher index: must lose the tie
        full = _popcount_rows(q[:, None, :] ^ t[None, :, :])
        ref_idx, ref_d = full.argmin(1), full.min(1)
        pre = _popcount_rows(q[:, None, :16] ^ t[None, :, :16])
        for i in range(nq):
            best = ((th + 1) << 16) | 0
            for j in range(nt):
                if pre[i, j] < (best >> 16):
                    best = min(best, (int(full[i, j]) << 16) | j)
            if ref_d[i] <= th:
                assert best == ((int(ref_d[i]) << 16) | int(ref_idx[i])), (trial, i)
            else:
                assert best >> 16 == th + 1      # "no match": dropped by the threshold filters downstream
