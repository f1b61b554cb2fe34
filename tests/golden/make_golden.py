#!/usr/bin/env python
"""Generate tests/golden/golden_v1.npz from the REAL OpenCV (cv2) in this image.

Run in the build container (needs /root/reference for the two fixture images and cv2):
    python tests/golden/make_golden.py
The vectors pin the oracle's restatement of the third-party arithmetic the reference calls
(FAST, ORB::compute, BFMatcher, findFundamentalMat) and are what the -m "not gpu" tests and
the GPU parity tests compare against where /root/reference and cv2 are not available.
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from stereo_vo_b200 import synth  # noqa: E402

REF_TESTS = "/root/reference/libstereo-odometry/tests"


def kp_xys(k):
    return np.stack([k["x"], k["y"], k["response"]], 1).astype(np.int32)


def main():
    out = {"cv2_version": np.array(cv2.__version__)}
    # --- the reference's own test fixtures (tests/0L.png, tests/0R.png: 800x600 gray, rectified) ---
    for name in ("0L", "0R"):
        img = cv2.imread(os.path.join(REF_TESTS, name + ".png"), cv2.IMREAD_GRAYSCALE)
        assert img is not None and img.shape == (600, 800)
        out["img_" + name] = img
        for th in (20,):
            k = orc.cv2_kps_to_array(cv2.FastFeatureDetector_create(th).detect(img))
            out[f"fast_{name}_th{th}"] = kp_xys(k)
            k2, d2 = orc.cv2_fast_orb(img, th)
            out[f"orbkp_{name}_th{th}"] = kp_xys(k2)
            out[f"orbdesc_{name}_th{th}"] = d2
    # BFMatcher + the S3 filter on the fixture descriptors
    m = orc.cv2_bf_match(out["orbdesc_0L_th20"], out["orbdesc_0R_th20"])
    out["bf_0L_0R"] = np.stack([m["trainIdx"], m["distance"].astype(np.int32)], 1)
    # --- synthetic c1 sequence: every stage boundary from the cv2 back-end ---
    frames, cam = synth.make_sequence("c1", seed=3, n_frames=3)
    camo = orc.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    est = orc.OracleEstimator(orc.Params(orb_nfeats=100), backend="cv2", ransac="cv2")
    for i, (L, R) in enumerate(frames):
        o = est.process_new_image_pair(L, R, camo)
        f = o["frame"]
        out[f"c1_f{i}_kpsL"] = kp_xys(f.kpsL)
        out[f"c1_f{i}_kpsR"] = kp_xys(f.kpsR)
        out[f"c1_f{i}_descL"] = f.descL
        out[f"c1_f{i}_descR"] = f.descR
        out[f"c1_f{i}_matches"] = np.stack([f.matches["queryIdx"], f.matches["trainIdx"],
                                            f.matches["distance"].astype(np.int32)], 1)
        out[f"c1_f{i}_tracked"] = o["tracked"]
        out[f"c1_f{i}_status"] = np.array([o["error_code"], int(o["valid"]), o["num_it"], o["num_it_final"]])
        if o["pose"] is not None:
            out[f"c1_f{i}_delta"] = o["pose"]["delta"]
            out[f"c1_f{i}_pose_t"] = o["pose"]["t"]
            out[f"c1_f{i}_pose_rotvec"] = o["pose"]["rotvec"]
            out[f"c1_f{i}_residual"] = o["out_residual"]
        if "track_dbg" in o:
            t = o["track_dbg"]
            for key in ("p1L", "p2L", "p1R", "p2R"):
                out[f"c1_f{i}_{key}"] = t[key]
            out[f"c1_f{i}_inlL"] = t["inlL"] if t["inlL"] is not None else np.zeros(0, np.uint8)
            out[f"c1_f{i}_inlR"] = t["inlR"] if t["inlR"] is not None else np.zeros(0, np.uint8)
    # --- BFMatcher on random descriptors with duplicated train rows (tie-break = lowest index) ---
    rng = np.random.default_rng(7)
    q = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (257, 32), dtype=np.uint8)
    t[100:130] = t[0:30]
    q[:50] = t[100:150]
    m = orc.cv2_bf_match(q, t)
    out["bf_rand_q"] = q
    out["bf_rand_t"] = t
    out["bf_rand"] = np.stack([m["trainIdx"], m["distance"].astype(np.int32)], 1)
    # --- blur anchor: cv2.sepFilter2D float path == blur inside ORB::compute ---
    img = synth.make_texture(np.random.default_rng(11), 96, 128)
    k = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    out["blur_in"] = img
    out["blur_out"] = cv2.sepFilter2D(img, cv2.CV_8U, k, k, borderType=cv2.BORDER_REFLECT_101)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"), **out)
    print("wrote golden_v1.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
