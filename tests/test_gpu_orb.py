"""-m gpu: the dmORB detector (svo_config.detect_method = 0, stage2_detect.cpp:458-496) against the real OpenCV.

cv::ORB::detectAndCompute is third-party code the reference calls; cv2 4.13 (the OpenCV of this image) IS the oracle here,
called with exactly the reference's arguments (oracle.cv2_orb_detect_and_compute).  The reference's own glue behind it
(m_non_max_sup, m_update_indexes, stages 3-5 on float keypoints) is checked against the restated oracle, which
tests/test_ref_cpu.py pins against the reference's compiled text (incl. the dmORB branch)."""
import numpy as np
import pytest

from oracle import oracle as orc
from stereo_vo_b200 import capi, synth

cv2 = pytest.importorskip("cv2")
pytestmark = pytest.mark.gpu

FIELDS = ("x", "y", "size", "angle", "response", "octave")


def _images():
    out = []
    for cfg, seed in (("c1", 3), ("c2", 2), ("c3", 1)):
        frames, _ = synth.make_sequence(cfg, seed=seed, n_frames=1)
        out.append((cfg, frames[0][0]))
        out.append((cfg, frames[0][1]))
    return out


def _assert_kps_equal(a, b, what):
    assert len(a) == len(b), f"{what}: {len(a)} vs {len(b)} keypoints"
    for f in FIELDS:
        assert np.array_equal(a[f], b[f]), f"{what}: field {f} differs at {np.nonzero(a[f] != b[f])[0][:5]}"


@pytest.mark.parametrize("nfeats,fast_th", [(200, 20), (500, 20), (100, 7)])
def test_orb_detect_and_compute_equals_opencv(nfeats, fast_th):
    """raw output of detectAndCompute (all keypoint fields, order) and the final keypoints + descriptors of stage 2"""
    for cfg, img in _images():
        H, W = img.shape
        ctx = capi.Context(width=W, height=H, orb_nfeats=nfeats, max_seqs=1, detect_method=0, fast_threshold=fast_th, max_corners=60000)
        try:
            kps, desc = ctx.detect(img)
            raw = ctx.orb_raw_keypoints(ctx.cfg.max_seqs, 0)                 # svo_detect works on the hidden slot
            ref_k, ref_d = orc.cv2_orb_detect_and_compute(img, int(1.5 * nfeats), 8, fast_th)
            _assert_kps_equal(raw, ref_k, f"{cfg} {W}x{H} raw")
            est = orc.OracleEstimator(orc.Params(orb_nfeats=nfeats, detect_method=0, initial_FAST_threshold=fast_th),
                                      backend="cv2", ransac="off")
            ok, od, _ = est.detect(img)
            _assert_kps_equal(kps, ok, f"{cfg} {W}x{H} after m_non_max_sup + m_update_indexes")
            assert np.array_equal(desc, od), f"{cfg}: descriptors differ in rows {np.nonzero((desc != od).any(axis=1))[0][:8]}"
        finally:
            ctx.close()


def test_orb_pyramid_levels_equal_opencv_resize_and_blur():
    from oracle import pyoracle
    frames, _ = synth.make_sequence("c2", seed=5, n_frames=1)
    img = frames[0][0]
    ctx = capi.Context(width=640, height=480, orb_nfeats=300, max_seqs=1, detect_method=0)
    try:
        ctx.detect(img)
        hs = ctx.cfg.max_seqs
        prev = img
        for level in range(1, 8):
            got = ctx.orb_level(hs, 0, level, 0)
            want = cv2.resize(prev, (got.shape[1], got.shape[0]), interpolation=cv2.INTER_LINEAR_EXACT)
            assert np.array_equal(got, want), f"level {level} image"
            assert np.array_equal(ctx.orb_level(hs, 0, level, 1), pyoracle.blur7_exact(want)), f"level {level} blur"
            prev = want
        assert ctx.orb_level(hs, 0, 1, 0).shape == (400, 533)
    finally:
        ctx.close()


def test_orb_without_nms_keeps_opencvs_list():
    frames, _ = synth.make_sequence("c1", seed=8, n_frames=1)
    img = frames[0][0]
    ctx = capi.Context(width=320, height=240, orb_nfeats=150, max_seqs=1, detect_method=0, non_maximal_suppression=0, max_corners=4096)
    try:
        kps, desc = ctx.detect(img)
        est = orc.OracleEstimator(orc.Params(orb_nfeats=150, detect_method=0, non_maximal_suppression=False), backend="cv2", ransac="off")
        ok, od, _ = est.detect(img)
        _assert_kps_equal(kps, ok, "no NMS")
        assert np.array_equal(desc, od)
    finally:
        ctx.close()


def _run(cfg_name, nfeats, seed, n_frames, **params):
    frames, cam = synth.make_sequence(cfg_name, seed=seed, n_frames=n_frames)
    H, W = frames[0][0].shape
    ctx = capi.Context(width=W, height=H, orb_nfeats=nfeats, max_seqs=2, detect_method=0, **params)
    est = orc.OracleEstimator(orc.Params(orb_nfeats=nfeats, detect_method=0, **params), backend="cv2", ransac="cv2_tolerant")
    oc = orc.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    gc = capi.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    n_valid = 0
    try:
        for i, (L, R) in enumerate(frames):
            o = est.process_new_image_pair(L, R, oc)
            r = ctx.process_batch([L], [R], [gc], slots=[1])[0]
            f = o["frame"]
            _assert_kps_equal(ctx.keypoints(1, 0, 0), f.kpsL, f"frame {i} left")
            _assert_kps_equal(ctx.keypoints(1, 0, 1), f.kpsR, f"frame {i} right")
            assert np.array_equal(ctx.descriptors(1, 0, 0), f.descL) and np.array_equal(ctx.descriptors(1, 0, 1), f.descR)
            m = ctx.matches(1, 0)
            assert np.array_equal(np.stack([m["queryIdx"], m["trainIdx"], m["distance"].astype(np.int32)], 1),
                                  np.stack([f.matches["queryIdx"], f.matches["trainIdx"], f.matches["distance"].astype(np.int32)], 1)), f"frame {i}: stereo matches"
            if i > 0 and "track_dbg" in o:
                q, tl, tr, il, ir = ctx.track_candidates(1)
                t = o["track_dbg"]
                assert np.array_equal(q, t["matL"]["queryIdx"]) and np.array_equal(tl, t["matL"]["trainIdx"])
                if t["inlL"] is not None:
                    assert np.array_equal(il, t["inlL"]) and np.array_equal(ir, t["inlR"]), f"frame {i}: RANSAC masks"
                assert np.array_equal(ctx.tracked_pairs(1), o["tracked"]), f"frame {i}: tracked pairs"
            assert (r.error_code, bool(r.valid), r.num_it, r.num_it_final) == (o["error_code"], o["valid"], o["num_it"], o["num_it_final"]), f"frame {i}"
            if o["pose"] is not None:
                np.testing.assert_allclose(r.delta_pose[:3], o["pose"]["delta"][:3], atol=1e-4)
                np.testing.assert_allclose(r.delta_pose[3:], o["pose"]["delta"][3:], atol=1e-3)
                assert np.array_equal(ctx.outliers(1), o["outliers"]), f"frame {i}: inlier list"
                n_valid += bool(r.valid)
    finally:
        ctx.close()
    return n_valid


def test_pipeline_orb_c1_equals_oracle():
    assert _run("c1", 150, seed=4, n_frames=5) >= 3


def test_pipeline_orb_c2_equals_oracle():
    assert _run("c2", 500, seed=1, n_frames=4, max_y_diff=2.0) >= 2


def test_orb_other_level_counts_and_c3_size():
    """orb_nlevels = 4 on a c3-sized image (1242x375, 600 features)"""
    frames, _ = synth.make_sequence("c3", seed=2, n_frames=1)
    img = frames[0][0]
    ctx = capi.Context(width=1242, height=375, orb_nfeats=400, max_seqs=1, detect_method=0, orb_nlevels=4, max_corners=32768)
    try:
        kps, desc = ctx.detect(img)
        raw = ctx.orb_raw_keypoints(ctx.cfg.max_seqs, 0)
        ref_k, _ = orc.cv2_orb_detect_and_compute(img, 600, 4, 20)
        _assert_kps_equal(raw, ref_k, "raw, 4 levels")
        est = orc.OracleEstimator(orc.Params(orb_nfeats=400, detect_method=0, orb_nlevels=4), backend="cv2", ransac="off")
        ok, od, _ = est.detect(img)
        _assert_kps_equal(kps, ok, "final, 4 levels")
        assert np.array_equal(desc, od) and set(kps["octave"]) <= {0, 1, 2, 3}
    finally:
        ctx.close()


@pytest.mark.parametrize("match_method,ifm_method", [(1, 0), (0, 1), (1, 1)])
def test_pipeline_orb_with_the_band_limited_matchers(match_method, ifm_method):
    """dmORB keypoints (float coordinates) through smDescRbR (S3:185-419) and ifmDescWin (S4:435-738): the float
    comparisons of those branches (disparity, row lists pt.y <= int(y), column windows) against the oracle, which
    tests/test_ref_cpu.py pins against the reference's compiled text for these combinations"""
    frames, cam = synth.make_sequence("c1", seed=61, n_frames=5)
    ctx = capi.Context(width=320, height=240, orb_nfeats=150, max_seqs=1, detect_method=0, match_method=match_method,
                       if_match_method=ifm_method, max_y_diff=2.0)
    est = orc.OracleEstimator(orc.Params(orb_nfeats=150, detect_method=0, match_method=match_method, ifm_method=ifm_method, max_y_diff=2.0),
                              backend="cv2", ransac="cv2_tolerant")
    oc = orc.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    gc = capi.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    try:
        n_tracked = 0
        for i, (L, R) in enumerate(frames):
            o = est.process_new_image_pair(L, R, oc)
            r = ctx.process_batch([L], [R], [gc], slots=[0])[0]
            f = o["frame"]
            _assert_kps_equal(ctx.keypoints(0, 0, 0), f.kpsL, f"frame {i} left")
            m = ctx.matches(0, 0)
            assert np.array_equal(np.stack([m["queryIdx"], m["trainIdx"], m["distance"].astype(np.int32), m["imgIdx"]], 1),
                                  np.stack([f.matches["queryIdx"], f.matches["trainIdx"], f.matches["distance"].astype(np.int32), f.matches["imgIdx"]], 1)), f"frame {i}: stereo matches"
            if i > 0:
                assert np.array_equal(ctx.tracked_pairs(0), o["tracked"]), f"frame {i}: tracked pairs"
                n_tracked += len(o["tracked"])
            assert (r.error_code, bool(r.valid), r.num_it, r.num_it_final) == (o["error_code"], o["valid"], o["num_it"], o["num_it_final"]), f"frame {i}"
            if o["pose"] is not None:
                np.testing.assert_allclose(r.delta_pose[:], o["pose"]["delta"], atol=1e-4)
        assert n_tracked > 40
    finally:
        ctx.close()


def test_orb_config_is_validated():
    with pytest.raises(Exception):
        capi.Context(width=640, height=480, max_seqs=1, detect_method=0, n_octaves=2)
    with pytest.raises(Exception):
        capi.Context(width=640, height=480, max_seqs=1, detect_method=2)


def test_host_shim_runs_dmorb_from_the_ini(tmp_path):
    """rso::CStereoOdometryEstimator (host shim) with detect_method = dmORB selected in the INI file (H:554-663): the demo
    loop's composed camera poses equal the oracle's with the same detector; nOctaves of the INI is ignored (S1:80)."""
    import os
    import struct
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    demo = os.path.join(root, "stereo_vo_b200", "host", "svo_demo")
    assert os.path.exists(demo), "host shim not built (run __graft_entry__.build())"
    frames, cam = synth.make_sequence("c1", seed=12, n_frames=4)
    seq = tmp_path / "seq.bin"
    with open(seq, "wb") as f:
        f.write(struct.pack("<3i5d", 320, 240, len(frames), cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"]))
        for (L, R) in frames:
            f.write(L.tobytes())
            f.write(R.tobytes())
    text = open(os.path.join(root, "stereo_vo_b200", "host", "hot_path.ini")).read()
    assert "detect_method = 1" in text and "orb_nfeats = 500" in text and "nOctaves = 1" in text
    text = text.replace("detect_method = 1", "detect_method = 0").replace("orb_nfeats = 500", "orb_nfeats = 150").replace("nOctaves = 1", "nOctaves = 3")
    ini = tmp_path / "cfg.ini"
    ini.write_text(text)
    out = tmp_path / "camera_pose.txt"
    p = subprocess.run([demo, str(seq), str(out), str(ini)], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    got = np.loadtxt(out).reshape(-1, 6)
    est = orc.OracleEstimator(orc.Params(orb_nfeats=150, detect_method=0, nOctaves=3), backend="cv2", ransac="cv2_tolerant")
    oc = orc.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    pose = np.eye(4)
    exp = []
    for (L, R) in frames:
        o = est.process_new_image_pair(L, R, oc)
        if o["valid"]:
            pose = pose @ o["pose"]["mat"]
        ypr = [np.arctan2(pose[1, 0], pose[0, 0]), np.arctan2(-pose[2, 0], np.hypot(pose[0, 0], pose[1, 0])), np.arctan2(pose[2, 1], pose[2, 2])]
        exp.append([pose[0, 3], pose[1, 3], pose[2, 3]] + ypr)
    np.testing.assert_allclose(got, np.array(exp), atol=1.5e-3)     # the file holds %.3f values
    assert abs(got[-1, 0]) > 0.05                                   # the sequence moves


def test_orb_soak_many_sequences():
    """bounded soak of the dmORB pipeline: 7 seeded sequences (c1, c2), every frame compared with the oracle"""
    valid = 0
    for seed in range(30, 35):
        valid += _run("c1", 120, seed=seed, n_frames=6)
    for seed in (40, 41):
        valid += _run("c2", 400, seed=seed, n_frames=4, max_y_diff=2.0)
    assert valid >= 20


def test_orb_flat_image_and_single_blob():
    flat = np.full((240, 320), 128, np.uint8)
    blob = flat.copy()
    blob[100:140, 150:190] = 220
    ctx = capi.Context(width=320, height=240, orb_nfeats=100, max_seqs=1, detect_method=0)
    try:
        for img in (flat, blob):
            kps, desc = ctx.detect(img)
            raw = ctx.orb_raw_keypoints(ctx.cfg.max_seqs, 0)
            ref_k, ref_d = orc.cv2_orb_detect_and_compute(img, 150, 8, 20)
            _assert_kps_equal(raw, ref_k, "raw")
            est = orc.OracleEstimator(orc.Params(orb_nfeats=100, detect_method=0), backend="cv2", ransac="off")
            ok, od, _ = est.detect(img)
            _assert_kps_equal(kps, ok, "final")
            assert np.array_equal(desc, od)
        assert len(ctx.detect(flat)[0]) == 0
    finally:
        ctx.close()


def test_orb_match_lr_seam_takes_float_keypoints():
    """svo_match_lr on caller data in a dmORB context: cv::ORB keypoints (float coordinates) go through the stage-3 filter"""
    frames, _ = synth.make_sequence("c1", seed=21, n_frames=1)
    L, R = frames[0]
    ctx = capi.Context(width=320, height=240, orb_nfeats=150, max_seqs=1, detect_method=0, max_y_diff=2.0)
    try:
        kl, dl = ctx.detect(L)
        kr, dr = ctx.detect(R)
        got = ctx.match_lr(kl, dl, kr, dr)
        want = orc.lr_filter(orc.cv2_bf_match(dl, dr), kl, kr, False, 2.0, 40, 320)
        assert len(got) == len(want) > 20
        assert np.array_equal(got["queryIdx"], want["queryIdx"]) and np.array_equal(got["trainIdx"], want["trainIdx"])
        assert (kl["x"] != np.floor(kl["x"])).any()          # the coarser levels give non-integer coordinates
    finally:
        ctx.close()


def _rows(k):
    return np.stack([k["x"], k["y"], k["size"], k["angle"], k["response"], k["octave"].astype(np.float32)], 1).astype(np.float32)


def test_orb_against_the_committed_opencv_golden():
    """tests/golden/golden_orb.npz (generated from cv2 by make_golden_orb.py): raw + final keypoints and descriptors of three
    cases, and a 3-frame c1 sequence through the whole pipeline (matches, RANSAC masks, tracked pairs, status, pose)"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_orb.npz"))
    for ci, case in enumerate(g["cases"]):
        cfg, seed, nfeats, th = str(case).split()
        frames, _ = synth.make_sequence(cfg, seed=int(seed), n_frames=1)
        img = frames[0][0]
        ctx = capi.Context(width=img.shape[1], height=img.shape[0], orb_nfeats=int(nfeats), max_seqs=1, detect_method=0,
                           fast_threshold=int(th), max_corners=32768)
        try:
            kps, desc = ctx.detect(img)
            assert np.array_equal(_rows(ctx.orb_raw_keypoints(ctx.cfg.max_seqs, 0)), g[f"case{ci}_raw_kps"]), f"case {ci}: raw keypoints"
            assert np.array_equal(_rows(kps), g[f"case{ci}_final_kps"]) and np.array_equal(desc, g[f"case{ci}_final_desc"]), f"case {ci}: final"
        finally:
            ctx.close()
    frames, cam = synth.make_sequence("c1", seed=4, n_frames=3)
    gc = capi.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    ctx = capi.Context(width=320, height=240, orb_nfeats=150, max_seqs=1, detect_method=0)
    try:
        for i, (L, R) in enumerate(frames):
            r = ctx.process_batch([L], [R], [gc], slots=[0])[0]
            assert np.array_equal(_rows(ctx.keypoints(0, 0, 0)), g[f"seq_f{i}_kpsL"]) and np.array_equal(ctx.descriptors(0, 0, 0), g[f"seq_f{i}_descL"])
            m = ctx.matches(0, 0)
            assert np.array_equal(np.stack([m["queryIdx"], m["trainIdx"], m["distance"].astype(np.int32)], 1), g[f"seq_f{i}_matches"])
            assert [r.error_code, int(r.valid), r.num_it, r.num_it_final] == g[f"seq_f{i}_status"].tolist()
            if i > 0:
                _, _, _, il, ir = ctx.track_candidates(0)
                assert np.array_equal(il, g[f"seq_f{i}_inlL"]) and np.array_equal(ir, g[f"seq_f{i}_inlR"])
                assert np.array_equal(ctx.tracked_pairs(0), g[f"seq_f{i}_tracked"])
                np.testing.assert_allclose(r.delta_pose[:], g[f"seq_f{i}_delta"], atol=1e-9)
    finally:
        ctx.close()


def test_orb_batch_of_slots_thresholds_and_repeat():
    """three estimators on one context with dmORB: slot 1 changes its FAST / ORB thresholds at frame 2 (the FAST threshold is
    cv::ORB's fastThreshold, S2:491), slot 2 repeats frame 3 (request.repeat, P:86-100); every slot equals its own oracle"""
    frames, cam = synth.make_sequence("c1", seed=77, n_frames=6)
    oc = orc.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    gc = capi.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    ctx = capi.Context(width=320, height=240, orb_nfeats=120, max_seqs=3, detect_method=0, fundmat_filter=0)
    ests = [orc.OracleEstimator(orc.Params(orb_nfeats=120, detect_method=0), backend="cv2", ransac="off") for _ in range(3)]
    try:
        for i, (L, R) in enumerate(frames):
            if i == 2:
                ctx.set_thresholds(11, 50, slot=1)
                ests[1].m_current_fast_th, ests[1].m_current_orb_th = 11, 50
            rep = [False, False, i == 3]
            outs = [e.process_new_image_pair(L, R, oc, repeat=rp) if rp else e.process_new_image_pair(L, R, oc) for e, rp in zip(ests, rep)]
            r = ctx.process_batch([L, L, L], [R, R, R], [gc] * 3, repeat=rep)
            for slot, o in enumerate(outs):
                _assert_kps_equal(ctx.keypoints(slot, 0, 0), o["frame"].kpsL, f"frame {i} slot {slot}")
                assert np.array_equal(ctx.descriptors(slot, 0, 1), o["frame"].descR)
                m = ctx.matches(slot, 0)
                assert np.array_equal(m["trainIdx"], o["frame"].matches["trainIdx"]), (i, slot)
                assert (r[slot].error_code, bool(r[slot].valid)) == (o["error_code"], o["valid"]), (i, slot)
                if i > 0:
                    assert np.array_equal(ctx.tracked_pairs(slot), o["tracked"]), (i, slot)
        assert len(ctx.keypoints(1, 0, 0)) != len(ctx.keypoints(0, 0, 0)) or not np.array_equal(ctx.keypoints(1, 0, 0)["x"], ctx.keypoints(0, 0, 0)["x"])
    finally:
        ctx.close()
