#!/usr/bin/env python
"""Generate tests/golden/golden_octaves.npz from the REAL OpenCV (cv2) in this image: the multi-octave path
(params_rectify.nOctaves = 2) on a seeded config-1 sequence -- cv2's FAST / ORB::compute / BFMatcher on every pyramid
level, the reference's glue and Gauss-Newton from the C++ restatement, F-matrix filter off (the masks are pinned
separately in golden_v1.npz) -- plus cv2.pyrDown of a config-3 image (the cvPyrDown branch of the pyramid).

    python tests/golden/make_golden_octaves.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from stereo_vo_b200 import synth  # noqa: E402


def kp_xys(k):
    return np.stack([k["x"], k["y"], k["response"]], 1).astype(np.int32)


def main():
    out = {"cv2_version": np.array(cv2.__version__)}
    frames, cam = synth.make_sequence("c1", seed=3, n_frames=3)
    oc = orc.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    est = orc.OracleEstimatorOctaves(orc.Params(orb_nfeats=100, nOctaves=2, vo_use_matches_ids=True), backend="cv2", ransac="off")
    for i, (L, R) in enumerate(frames):
        o = est.process_new_image_pair(L, R, oc)
        for k, f in enumerate(o["frames"]):
            out[f"f{i}_o{k}_kpsL"] = kp_xys(f.kpsL)
            out[f"f{i}_o{k}_kpsR"] = kp_xys(f.kpsR)
            out[f"f{i}_o{k}_descL"] = f.descL
            out[f"f{i}_o{k}_descR"] = f.descR
            out[f"f{i}_o{k}_matches"] = np.stack([f.matches["queryIdx"], f.matches["trainIdx"], f.matches["distance"].astype(np.int32)], 1)
            out[f"f{i}_o{k}_ids"] = f.ids
            if "track_dbg" in o:
                out[f"f{i}_o{k}_tracked"] = o["tracked"][k]
        out[f"f{i}_status"] = np.array([o["error_code"], int(o["valid"]), o["num_it"], o["num_it_final"]])
        if o["pose"] is not None:
            out[f"f{i}_delta"] = o["pose"]["delta"]
            out[f"f{i}_outliers"] = o["outliers"]
            out[f"f{i}_residual"] = o["out_residual"]
    # the cvPyrDown branch (width not a multiple of 16): level 1 of a config-3 image straight from cv2
    f3, _ = synth.make_sequence("c3", seed=2, n_frames=1)
    out["c3_pyrdown_L"] = cv2.pyrDown(f3[0][0], dstsize=(621, 187))
    out["c3_pyrdown_R"] = cv2.pyrDown(f3[0][1], dstsize=(621, 187))
    path = os.path.join(ROOT, "tests", "golden", "golden_octaves.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
