#!/usr/bin/env python
"""A few steps of the dmORB pipeline (svo_config.detect_method = 0) on device-resident synthetic c2 sequences: the workload
for the ncu launch list of that path, e.g.

  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_orb.csv \
      python tools/run_orb.py --steps 3

Prints one JSON line with the CUDA-event time per step (not a bench value when run under ncu)."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from stereo_vo_b200 import capi, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    cfg = synth.CONFIGS["c2"]
    W, H = cfg["W"], cfg["H"]
    n_frames = args.steps + args.warmup
    seqs = [synth.make_sequence("c2", seed=s, n_frames=n_frames) for s in range(4)]
    cam = seqs[0][1]
    gc = capi.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    B = args.batch
    dev = []
    for t in range(n_frames):
        L = np.stack([seqs[i % 4][0][t][0] for i in range(B)])
        R = np.stack([seqs[i % 4][0][t][1] for i in range(B)])
        dev.append((torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda()))
    torch.cuda.synchronize()
    ctx = capi.Context(width=W, height=H, orb_nfeats=500, max_seqs=B, detect_method=0, max_y_diff=2.0)
    stream = torch.cuda.ExternalStream(ctx.stream())
    cams = [gc] * B
    for t in range(args.warmup):
        ctx.process_batch_device(dev[t][0].data_ptr(), dev[t][1].data_ptr(), B, W, W * H, cams)
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for t in range(args.warmup, n_frames):
        ctx.process_batch_device(dev[t][0].data_ptr(), dev[t][1].data_ptr(), B, W, W * H, cams)
    e1.record(stream)
    e1.synchronize()
    res = ctx.fetch_results(B)
    print(json.dumps({"batch": B, "ms_per_step": e0.elapsed_time(e1) / args.steps, "valid": sum(int(r.valid) for r in res),
                      "stats_slot0": ctx.octave_stats(0)[0].tolist()}))
    ctx.close()


if __name__ == "__main__":
    main()
