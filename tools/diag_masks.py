#!/usr/bin/env python
"""Diagnostic for a soak mismatch: on one sequence, compare the two F-matrix RANSAC masks of octave 0 from (a) the device
kernel, (b) cv2.findFundamentalMat, (c) the host build of csrc/fundmat_core.cuh on the same point sets.
usage: python tools/diag_masks.py <config> <seed> <n_frames> <n_octaves>     (test infrastructure, GPU box)"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc          # noqa: E402
from stereo_vo_b200 import capi, synth    # noqa: E402


def main():
    cfg, seed, n_frames, n_oct = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    so = "/tmp/libfundmat_host.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", so,
                           os.path.join(ROOT, "tests", "cpp", "fundmat_host.cpp")])
    lib = C.CDLL(so)

    def host_mask(p1, p2):
        p1 = np.ascontiguousarray(p1, np.float32)
        p2 = np.ascontiguousarray(p2, np.float32)
        mask = np.zeros(max(len(p1), 1), np.uint8)
        lib.fundmat_host(p1.ctypes.data_as(C.c_void_p), p2.ctypes.data_as(C.c_void_p), len(p1), mask.ctypes.data_as(C.c_void_p), None)
        return mask[:len(p1)]
    W, H, nf = dict(c1=(320, 240, 100), c2=(640, 480, 500), c3=(1242, 375, 1000))[cfg]
    frames, cam = synth.make_sequence(cfg, seed=seed, n_frames=n_frames)
    ctx = capi.Context(width=W, height=H, orb_nfeats=nf, max_seqs=1, max_corners=32768, n_octaves=n_oct)
    est = orc.OracleEstimatorOctaves(orc.Params(orb_nfeats=nf, nOctaves=n_oct), backend="port", ransac="cv2")
    gc = capi.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    oc = orc.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    for t, (L, R) in enumerate(frames):
        ctx.process_batch([L], [R], [gc])
        try:
            o = est.process_new_image_pair(L, R, oc)
        except Exception as e:      # noqa: BLE001
            print(f"frame {t}: oracle stopped: {str(e).splitlines()[0][:100]}")
            break
        if t == 0:
            continue
        d = o["track_dbg"][0]
        q, tl, tr, il, ir = ctx.track_candidates(0)
        same_lists = np.array_equal(q, d["matL"]["queryIdx"]) and np.array_equal(tl, d["matL"]["trainIdx"])
        for side, dev, cv, p1, p2 in (("L", il, d["inlL"], d["p1L"], d["p2L"]), ("R", ir, d["inlR"], d["p1R"], d["p2R"])):
            if cv is None:
                continue
            hm = host_mask(p1, p2)
            again = ctx.fundmat_ransac(p1, p2)[1]
            print(f"frame {t} {side}: n={len(p1)} lists_equal={same_lists} inliers cv2={int(cv.sum())} device={int(dev.sum())} host={int(hm.sum())} "
                  f"device_entry={int(again.sum())} | dev!=cv2: {int((dev != cv).sum())} host!=cv2: {int((hm != cv).sum())} dev!=host: {int((dev != hm).sum())} "
                  f"dev_pipeline!=dev_entry: {int((dev != again).sum())}")
        print(f"frame {t}: tracked equal = {np.array_equal(ctx.tracked_pairs(0, octave=0), o['tracked'][0])}")
    ctx.close()


if __name__ == "__main__":
    main()
