#!/usr/bin/env python
"""CPU-only census of the F-matrix RANSAC masks: the host build of stereo_vo_b200/csrc/fundmat_core.cuh (the arithmetic the
device kernel evaluates, ransac.cu is compiled without FMA contraction) against cv2.findFundamentalMat on every point set
the oracle meets in seeded sequences.  Test infrastructure.  usage: python tools/ransac_host_vs_cv2.py [n_seeds] [n_frames]"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc          # noqa: E402
from stereo_vo_b200 import synth          # noqa: E402


def main():
    n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    so = "/tmp/libfundmat_host_census%s.so" % os.environ.get("SVO_CENSUS_CR", "")
    # SVO_CENSUS_CR=1: build with the correctly rounded pow / acos / cos of csrc/cr_math.cuh (what the device evaluates)
    # instead of glibc's (what OpenCV on the host evaluates)
    extra = ["-DSVO_FORCE_CR_LIBM", "-mfma"] if os.environ.get("SVO_CENSUS_CR") == "1" else []
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", *extra, "-o", so,
                           os.path.join(ROOT, "tests", "cpp", "fundmat_host.cpp")])
    lib = C.CDLL(so)
    stats = {}

    def make_fn(key):
        st = stats.setdefault(key, dict(masks=0, differ=0, points=0, flipped=0, asserts=0))

        def fn(p1, p2):
            a = orc.cv2_fmat_mask(p1, p2)          # may raise: cv2 4.13's internal assertion (DESIGN.md limit (d))
            p1c = np.ascontiguousarray(p1, np.float32)
            p2c = np.ascontiguousarray(p2, np.float32)
            if a is not None and len(p1c) >= 7:
                m = np.zeros(len(p1c), np.uint8)
                lib.fundmat_host(p1c.ctypes.data_as(C.c_void_p), p2c.ctypes.data_as(C.c_void_p), len(p1c), m.ctypes.data_as(C.c_void_p), None)
                st["masks"] += 1
                st["points"] += len(p1c)
                if not np.array_equal(a, m):
                    st["differ"] += 1
                    st["flipped"] += int((a != m).sum())
            return a
        return fn, st
    for cfg, nf, n_oct in (("c1", 100, 1), ("c2", 500, 1), ("c3", 1000, 1), ("c1", 100, 2), ("c2", 500, 3)):
        fn, st = make_fn(f"{cfg} x {n_oct} octave(s)")
        for sd in range(100, 100 + n_seeds):
            frames, cam = synth.make_sequence(cfg, seed=sd, n_frames=n_frames)
            oc = orc.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
            est = orc.OracleEstimatorOctaves(orc.Params(orb_nfeats=nf, nOctaves=n_oct), backend="port", ransac=fn)
            for (L, R) in frames:
                try:
                    est.process_new_image_pair(L, R, oc)
                except Exception:                  # noqa: BLE001
                    st["asserts"] += 1
                    break
        print(f"{cfg} x {n_oct} octave(s): {st['masks']} masks ({st['points']} points), {st['differ']} differ from cv2 "
              f"({st['flipped']} points), {st['asserts']} sequences stopped by cv2's assertion", flush=True)


if __name__ == "__main__":
    main()
