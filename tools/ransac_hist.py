#!/usr/bin/env python
"""Distribution of the RANSAC iterations OpenCV's loop executes per findFundamentalMat call over all slots of bench.py's
workload (device-resident steps), and the kernel time of fundmat_kernel per step.  usage: python tools/ransac_hist.py [steps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                # noqa: E402
import bench                                # noqa: E402
from stereo_vo_b200 import capi             # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    two = len(sys.argv) > 2 and sys.argv[2] == "two"
    B = 512
    seqs, cam = bench.make_frames(two_motions=two)
    gc = capi.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    dev = torch.device("cuda", 0)
    pb = bench.Playback(torch, seqs, B, 0, dev, pinned=False)
    pb.build_device_batches(torch)
    ctx = capi.Context(width=640, height=480, orb_nfeats=500, max_seqs=B)
    ctx.set_concurrency(1)
    ctx.enable_stage_timing(True)
    for t in range(steps):
        L, R = pb.dev[t % pb.period]
        ctx.process_batch_device(L.data_ptr(), R.data_ptr(), B, 640, 640 * 480, [gc] * B)
        ctx.synchronize()
        if t == 0:
            continue
        it = np.array([ctx.ransac_stats(s) for s in range(B)])[:, :, 2].ravel()
        h = np.bincount(np.minimum(it, 40), minlength=41)
        print(f"step {t}: iterations mean {it.mean():.2f} max {it.max()}  histogram (0..40+): {h.tolist()}")
    print({k: round(v, 4) for k, v in ctx.stage_times().items()})
    ctx.close()


if __name__ == "__main__":
    main()
