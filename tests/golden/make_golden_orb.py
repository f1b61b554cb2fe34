#!/usr/bin/env python
"""Generate tests/golden/golden_orb.npz from the REAL OpenCV (cv2) in this image: the dmORB branch of the reference
(stage2_detect.cpp:458-496) -- cv2.ORB_create(1.5 * orb_nfeats, 1.2, 8, 31, 0, 2, HARRIS_SCORE, 31, fast_th).detectAndCompute
on seeded images (raw keypoints + descriptors), and a 3-frame config-1 sequence through the whole pipeline with that
detector (real OpenCV calls for ORB / BFMatcher / findFundamentalMat, the reference's glue from the C++ restatement).

    python tests/golden/make_golden_orb.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from stereo_vo_b200 import synth  # noqa: E402

CASES = (("c1", 3, 100, 20), ("c2", 2, 300, 20), ("c1", 8, 150, 9))     # config, seed, orb_nfeats, FAST threshold


def kp_rows(k):
    return np.stack([k["x"], k["y"], k["size"], k["angle"], k["response"], k["octave"].astype(np.float32)], 1).astype(np.float32)


def main():
    out = {"cv2_version": np.array(cv2.__version__), "cases": np.array([f"{c} {s} {n} {t}" for c, s, n, t in CASES])}
    for ci, (cfg, seed, nfeats, th) in enumerate(CASES):
        frames, _ = synth.make_sequence(cfg, seed=seed, n_frames=1)
        img = frames[0][0]
        k, d = orc.cv2_orb_detect_and_compute(img, int(1.5 * nfeats), 8, th)
        out[f"case{ci}_raw_kps"] = kp_rows(k)
        out[f"case{ci}_raw_desc"] = d
        est = orc.OracleEstimator(orc.Params(orb_nfeats=nfeats, detect_method=0, initial_FAST_threshold=th), backend="cv2", ransac="off")
        fk, fd, _ = est.detect(img)
        out[f"case{ci}_final_kps"] = kp_rows(fk)
        out[f"case{ci}_final_desc"] = fd
    frames, cam = synth.make_sequence("c1", seed=4, n_frames=3)
    oc = orc.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    est = orc.OracleEstimator(orc.Params(orb_nfeats=150, detect_method=0), backend="cv2", ransac="cv2")
    for i, (L, R) in enumerate(frames):
        o = est.process_new_image_pair(L, R, oc)
        f = o["frame"]
        out[f"seq_f{i}_kpsL"] = kp_rows(f.kpsL)
        out[f"seq_f{i}_descL"] = f.descL
        out[f"seq_f{i}_matches"] = np.stack([f.matches["queryIdx"], f.matches["trainIdx"], f.matches["distance"].astype(np.int32)], 1)
        out[f"seq_f{i}_status"] = np.array([o["error_code"], int(o["valid"]), o["num_it"], o["num_it_final"]])
        if i > 0:
            out[f"seq_f{i}_tracked"] = o["tracked"]
            out[f"seq_f{i}_inlL"] = o["track_dbg"]["inlL"]
            out[f"seq_f{i}_inlR"] = o["track_dbg"]["inlR"]
            out[f"seq_f{i}_delta"] = np.array(o["pose"]["delta"]) if o["pose"] is not None else np.zeros(6)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_orb.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
