#!/usr/bin/env python
"""Per-kernel device time of svo_process_pair at batch 1 (the literal processNewImagePair seam), 1 and 3 octaves.
usage: python tools/latency_breakdown.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stereo_vo_b200 import capi, synth            # noqa: E402


def main():
    frames, cam = synth.make_sequence("c2", seed=0, n_frames=8)
    gc = capi.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    for n_oct in (1, 3):
        ctx = capi.Context(width=640, height=480, orb_nfeats=500, max_seqs=1, n_octaves=n_oct)
        ts = []
        for t in range(40):
            L, R = frames[t % 8 if (t // 8) % 2 == 0 else 7 - t % 8]
            t0 = time.perf_counter()
            ctx.process_pair(L, R, gc)
            ts.append(time.perf_counter() - t0)
        ctx.enable_stage_timing(True)
        for t in range(40, 48):
            L, R = frames[t % 8 if (t // 8) % 2 == 0 else 7 - t % 8]
            ctx.process_pair(L, R, gc)
        st = ctx.stage_times()
        print(f"{n_oct} octave(s): wall median {1e3 * np.median(ts[5:]):.3f} ms; kernels sum {sum(st.values()):.3f} ms:",
              {k: round(v, 4) for k, v in st.items()})
        ctx.close()


if __name__ == "__main__":
    main()
