#!/usr/bin/env python
"""Device time of run7Point (the SVD-dominated 7-point solve) per wave of samples: svo_fundmat_7point on realistic samples
(three depth layers, sideways motion, integer pixel coordinates).  usage: python tools/fm7_timing.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stereo_vo_b200 import capi        # noqa: E402


def main():
    rng = np.random.default_rng(0)
    ctx = capi.Context(width=640, height=480, orb_nfeats=500, max_seqs=1)
    for n in (6, 6 * 148, 6 * 148 * 8, 6 * 148 * 32):
        p1 = np.round(rng.uniform(31, 609, (n, 7, 2))).astype(np.float32)
        d = rng.choice([4.0, 8.0, 15.0], (n, 7)).astype(np.float32)
        p2 = p1.copy()
        p2[:, :, 0] -= d
        p2[:, :, 1] += np.round(rng.normal(0, 0.3, (n, 7))).astype(np.float32)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            nm, F = ctx.fundmat_7point(p1, p2)
            ts.append(time.perf_counter() - t0)
        print(f"{n} samples ({n // 6} CTAs): wall min {1e3 * min(ts):.3f} ms (incl. malloc + copies), models/sample {nm.mean():.2f}")
    ctx.close()


if __name__ == "__main__":
    main()
