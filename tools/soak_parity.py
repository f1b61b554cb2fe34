#!/usr/bin/env python
"""Parity soak on a GPU box: many seeded config-2 sequences, every frame compared with the CPU oracle (real OpenCV
calls for FAST / ORB / BFMatcher / findFundamentalMat + the C++ restatement of the reference's glue).
Test infrastructure, not product code.  usage: python tools/soak_parity.py [n_seeds] [n_frames] [config] [n_octaves] [cv2_tolerant|cv2]
With n_octaves > 1 the octave-aware oracle runs the real OpenCV calls on every pyramid level."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc          # noqa: E402
from stereo_vo_b200 import capi, synth    # noqa: E402


def soak_octaves(n_seeds, n_frames, cfg, n_oct, ransac="cv2_tolerant", seed0=100):
    W, H, nf = dict(c1=(320, 240, 100), c2=(640, 480, 500), c3=(1242, 375, 1000))[cfg]
    ctx = capi.Context(width=W, height=H, orb_nfeats=nf, max_seqs=n_seeds, max_corners=32768, n_octaves=n_oct)
    seqs = [synth.make_sequence(cfg, seed=seed0 + s, n_frames=n_frames) for s in range(n_seeds)]
    cam = seqs[0][1]
    gc = capi.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    oc = orc.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    ests = [orc.OracleEstimatorOctaves(orc.Params(orb_nfeats=nf, nOctaves=n_oct), backend="cv2", ransac=ransac) for _ in range(n_seeds)]
    raised0 = orc.cv2_fmat_mask_tolerant.raised
    bad = n_cmp = 0
    worst_rot = worst_t = 0.0
    dead = set()
    t0 = time.time()
    for t in range(n_frames):
        res = ctx.process_batch([s[0][t][0] for s in seqs], [s[0][t][1] for s in seqs], [gc] * n_seeds)
        for k in range(n_seeds):
            if k in dead:
                continue
            try:
                o = ests[k].process_new_image_pair(seqs[k][0][t][0], seqs[k][0][t][1], oc)
            except Exception as e:          # noqa: BLE001
                print(f"ORACLE FAILED seed {100 + k} frame {t}: {str(e).splitlines()[0][:120]}")
                dead.add(k)
                continue
            n_cmp += 1
            checks = {"status": res[k].error_code == o["error_code"] and bool(res[k].valid) == o["valid"]}
            for j in range(n_oct):
                f = o["frames"][j]
                kp = ctx.keypoints(k, 0, 0, octave=j)
                checks[f"keypoints o{j}"] = (np.array_equal(kp["x"], f.kpsL["x"]) and np.array_equal(kp["y"], f.kpsL["y"])
                                             and np.array_equal(kp["response"], f.kpsL["response"]))
                checks[f"descriptors o{j}"] = np.array_equal(ctx.descriptors(k, 0, 1, octave=j), f.descR)
                m = ctx.matches(k, 0, octave=j)
                checks[f"matches o{j}"] = np.array_equal(m["trainIdx"], f.matches["trainIdx"]) and np.array_equal(m["queryIdx"], f.matches["queryIdx"])
                if "track_dbg" in o:
                    checks[f"tracked o{j}"] = np.array_equal(ctx.tracked_pairs(k, octave=j), o["tracked"][j])
            if "track_dbg" in o:
                checks["iterations"] = (res[k].num_it, res[k].num_it_final) == (o["num_it"], o["num_it_final"])
                if o["pose"] is not None:
                    d = np.array(res[k].delta_pose[:]) - o["pose"]["delta"]
                    worst_rot = max(worst_rot, float(np.abs(d[:3]).max()))
                    worst_t = max(worst_t, float(np.abs(d[3:]).max()))
                    checks["pose"] = bool(np.abs(d[:3]).max() < 1e-4 and np.abs(d[3:]).max() < 1e-3)
            failed = [n for n, v in checks.items() if not v]
            if failed:
                bad += 1
                print(f"MISMATCH seed {100 + k} frame {t}: {failed}")
    raised = orc.cv2_fmat_mask_tolerant.raised - raised0
    print(f"soak {cfg} with {n_oct} octaves: {n_seeds} sequences x {n_frames} frames, {n_cmp} stereo pairs compared, {bad} mismatching "
          f"frames, {len(dead)} sequences cut short by an OpenCV assertion in the oracle, {raised} findFundamentalMat calls where "
          f"cv2 raised and the host build of the same arithmetic answered; "
          f"max |d rot| = {worst_rot:.2e} rad, max |d t| = {worst_t:.2e} m; {time.time() - t0:.1f} s")
    ctx.close()
    return dict(compared=n_cmp, bad=bad, dead=len(dead), raised=raised, worst_rot=worst_rot, worst_t=worst_t)


def soak(n_seeds, n_frames, cfg, n_oct=1, ransac="cv2_tolerant", seed0=100):
    """Every frame of n_seeds seeded sequences: device (C-ABI) against the oracle with the real OpenCV calls."""
    if n_oct > 1:
        return soak_octaves(n_seeds, n_frames, cfg, n_oct, ransac, seed0)
    W, H, nf = dict(c1=(320, 240, 100), c2=(640, 480, 500), c3=(1242, 375, 1000))[cfg]
    ctx = capi.Context(width=W, height=H, orb_nfeats=nf, max_seqs=n_seeds, max_corners=32768 if cfg == "c3" else 16384)
    seqs = [synth.make_sequence(cfg, seed=seed0 + s, n_frames=n_frames) for s in range(n_seeds)]
    cam = seqs[0][1]
    gc = capi.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    oc = orc.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    ests = [orc.OracleEstimator(orc.Params(orb_nfeats=nf), backend="cv2", ransac=ransac) for _ in range(n_seeds)]
    raised0 = orc.cv2_fmat_mask_tolerant.raised
    bad = 0
    worst_rot = worst_t = 0.0
    t0 = time.time()
    dead = set()            # sequences whose oracle stopped (ransac="cv2": cv2.findFundamentalMat raises on some inputs)
    n_cmp = 0
    for t in range(n_frames):
        res = ctx.process_batch([s[0][t][0] for s in seqs], [s[0][t][1] for s in seqs], [gc] * n_seeds)
        for k in range(n_seeds):
            if k in dead:
                continue
            try:
                o = ests[k].process_new_image_pair(seqs[k][0][t][0], seqs[k][0][t][1], oc)
            except Exception as e:          # noqa: BLE001
                print(f"ORACLE FAILED seed {seed0 + k} frame {t}: {str(e).splitlines()[0][:120]}")
                dead.add(k)
                continue
            n_cmp += 1
            f = o["frame"]
            kp = ctx.keypoints(k, 0, 0)
            checks = {"keypoints": np.array_equal(kp["x"], f.kpsL["x"]) and np.array_equal(kp["y"], f.kpsL["y"])
                      and np.array_equal(kp["response"], f.kpsL["response"]),
                      "descriptors": np.array_equal(ctx.descriptors(k, 0, 1), f.descR)}
            m = ctx.matches(k, 0)
            checks["matches"] = (np.array_equal(m["trainIdx"], f.matches["trainIdx"]) and np.array_equal(m["queryIdx"], f.matches["queryIdx"])
                                 and np.array_equal(m["distance"], f.matches["distance"]))
            checks["status"] = res[k].error_code == o["error_code"] and bool(res[k].valid) == o["valid"]
            if t > 0:
                q, tl, tr, il, ir = ctx.track_candidates(k)
                dbg = o["track_dbg"]
                checks["greedy list"] = np.array_equal(q, dbg["matL"]["queryIdx"]) and np.array_equal(tl, dbg["matL"]["trainIdx"])
                if dbg["inlL"] is not None:
                    checks["mask L"] = np.array_equal(il, dbg["inlL"])
                    checks["mask R"] = np.array_equal(ir, dbg["inlR"])
                checks["tracked"] = np.array_equal(ctx.tracked_pairs(k), o["tracked"])
                checks["iterations"] = (res[k].num_it, res[k].num_it_final) == (o["num_it"], o["num_it_final"])
                checks["inliers"] = np.array_equal(ctx.outliers(k), o["outliers"]) if o["pose"] is not None else True
                if o["pose"] is not None:
                    d = np.array(res[k].delta_pose[:]) - o["pose"]["delta"]
                    worst_rot = max(worst_rot, float(np.abs(d[:3]).max()))
                    worst_t = max(worst_t, float(np.abs(d[3:]).max()))
                    checks["pose"] = bool(np.abs(d[:3]).max() < 1e-4 and np.abs(d[3:]).max() < 1e-3)
            failed = [n for n, v in checks.items() if not v]
            if failed:
                bad += 1
                extra = ""
                if "mask L" in failed or "mask R" in failed:
                    extra = f" (inliers L {int(il.sum())}/{int(np.sum(dbg['inlL']))} R {int(ir.sum())}/{int(np.sum(dbg['inlR']))} of {len(il)})"
                print(f"MISMATCH seed {seed0 + k} frame {t}: {failed}{extra}")
    raised = orc.cv2_fmat_mask_tolerant.raised - raised0
    print(f"soak {cfg}: {n_seeds} sequences x {n_frames} frames, {n_cmp} stereo pairs compared, {bad} mismatching frames, "
          f"{len(dead)} sequences cut short by an OpenCV assertion in the oracle, {raised} findFundamentalMat calls where cv2 raised "
          f"and the host build of the same arithmetic answered; "
          f"max |d rot| = {worst_rot:.2e} rad, max |d t| = {worst_t:.2e} m; {time.time() - t0:.1f} s")
    ctx.close()
    return dict(compared=n_cmp, bad=bad, dead=len(dead), raised=raised, worst_rot=worst_rot, worst_t=worst_t)


def main():
    n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    cfg = sys.argv[3] if len(sys.argv) > 3 else "c2"
    n_oct = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    ransac = sys.argv[5] if len(sys.argv) > 5 else "cv2_tolerant"
    return 1 if soak(n_seeds, n_frames, cfg, n_oct, ransac)["bad"] else 0


if __name__ == "__main__":
    sys.exit(main())
