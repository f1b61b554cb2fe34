#!/usr/bin/env python
"""Parity soak of the dmORB pipeline (svo_config.detect_method = 0) against the oracle with the real OpenCV calls
(cv::ORB::detectAndCompute, BFMatcher, findFundamentalMat) on seeded sequences; run on a GPU box:

  python tools/soak_orb.py > profiles/r02_parity_soak_dmorb.log
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc                     # noqa: E402
from stereo_vo_b200 import capi, synth               # noqa: E402

FIELDS = ("x", "y", "size", "angle", "response", "octave")


def run(cfg_name, nfeats, seed, n_frames, **params):
    frames, cam = synth.make_sequence(cfg_name, seed=seed, n_frames=n_frames)
    H, W = frames[0][0].shape
    ctx = capi.Context(width=W, height=H, orb_nfeats=nfeats, max_seqs=1, detect_method=0, max_corners=32768, **params)
    est = orc.OracleEstimator(orc.Params(orb_nfeats=nfeats, detect_method=0, **params), backend="cv2", ransac="cv2_tolerant")
    oc = orc.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    gc = capi.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    bad, valid, maxd = [], 0, 0.0
    try:
        for i, (L, R) in enumerate(frames):
            o = est.process_new_image_pair(L, R, oc)
            r = ctx.process_batch([L], [R], [gc], slots=[0])[0]
            f = o["frame"]
            why = []
            for side, ref in ((0, f.kpsL), (1, f.kpsR)):
                k = ctx.keypoints(0, 0, side)
                if len(k) != len(ref) or any(not np.array_equal(k[q], ref[q]) for q in FIELDS):
                    why.append("keypoints")
            if not (np.array_equal(ctx.descriptors(0, 0, 0), f.descL) and np.array_equal(ctx.descriptors(0, 0, 1), f.descR)):
                why.append("descriptors")
            m = ctx.matches(0, 0)
            if not (np.array_equal(m["queryIdx"], f.matches["queryIdx"]) and np.array_equal(m["trainIdx"], f.matches["trainIdx"])):
                why.append("stereo matches")
            if i > 0 and "track_dbg" in o:
                q, tl, tr, il, ir = ctx.track_candidates(0)
                t = o["track_dbg"]
                if t["inlL"] is not None and not (np.array_equal(il, t["inlL"]) and np.array_equal(ir, t["inlR"])):
                    why.append("RANSAC masks")
                if not np.array_equal(ctx.tracked_pairs(0), o["tracked"]):
                    why.append("tracked pairs")
            if (r.error_code, bool(r.valid), r.num_it, r.num_it_final) != (o["error_code"], o["valid"], o["num_it"], o["num_it_final"]):
                why.append("status / iterations")
            if o["pose"] is not None:
                d = float(np.abs(np.array(r.delta_pose[:]) - np.array(o["pose"]["delta"])).max())
                maxd = max(maxd, d)
                if d > 1e-6 or not np.array_equal(ctx.outliers(0), o["outliers"]):
                    why.append("pose / inliers")
                valid += bool(r.valid)
            if why:
                bad.append((i, why))
    finally:
        ctx.close()
    return len(frames), valid, maxd, bad, orc.cv2_fmat_mask_tolerant.raised


def main():
    t0 = time.time()
    total = nbad = 0
    for cfg, nfeats, seeds, nf, kw in (("c1", 150, range(200, 212), 8, {}), ("c2", 500, range(300, 308), 6, dict(max_y_diff=2.0)),
                                       ("c3", 800, range(400, 404), 4, dict(max_y_diff=2.0))):
        for seed in seeds:
            n, valid, maxd, bad, raised = run(cfg, nfeats, seed, nf, **kw)
            total += n
            nbad += len(bad)
            print(f"{cfg} seed {seed}: {n} frames, {valid} valid poses, max |delta pose diff| {maxd:.2e}, mismatching frames {bad}", flush=True)
    print(f"TOTAL {total} stereo pairs, {nbad} mismatching frames, cv2 findFundamentalMat assertions caught so far {orc.cv2_fmat_mask_tolerant.raised}, {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
