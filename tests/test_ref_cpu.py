"""restated oracle == the reference's own compiled text (CPU only).

oracle/_ref/libsvo_ref.so is built by oracle/Makefile.ref from the UNMODIFIED sources under
/root/reference/libstereo-odometry/src (common.cpp, process_new_image_pair.cpp, stage1_rectify.cpp ..
stage5_optimization.cpp) against stand-in third-party headers (oracle/ref_stubs/); the OpenCV calls inside go to the real
OpenCV (cv2) through callbacks.  Every test below drives rso::CStereoOdometryEstimator itself and the restated oracle
(oracle/svo_oracle.cpp + oracle/oracle.py, the checker of the -m gpu parity tests) on the same inputs.
Integer / index results must be identical; poses agree to 1e-12 (the two evaluate H^-1 g with different 6x6 solvers)."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from oracle import ref
from stereo_vo_b200 import synth

pytestmark = pytest.mark.skipif(not ref.available(), reason="neither oracle/_ref/libsvo_ref.so nor /root/reference is present")

POSE_TOL = 1e-12


def ocam(cam):
    return orc.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])


def _frames_of(out, n_oct):
    return out["frames"] if "frames" in out else [out["frame"]]


def _tracked_of(out, n_oct):
    t = out["tracked"]
    return t if isinstance(t, list) else [t]


def compare_step(a, b, n_oct, ids=False, where="", delta=True):
    """a: RefEstimator output, b: oracle output"""
    assert (a["error_code"], a["valid"]) == (b["error_code"], b["valid"]), where
    assert (a["num_it"], a["num_it_final"]) == (b["num_it"], b["num_it_final"]), where
    fb, tb = _frames_of(b, n_oct), _tracked_of(b, n_oct)
    for k in range(n_oct):
        f, g = a["frames"][k], fb[k]
        for name in ("kpsL", "kpsR", "descL", "descR", "matches"):
            assert np.array_equal(getattr(f, name), getattr(g, name)), f"{where} octave {k}: {name}"
        if "track_dbg" in b or len(tb[k]):
            assert np.array_equal(a["tracked"][k], np.asarray(tb[k]).reshape(-1, 2)), f"{where} octave {k}: tracked pairs"
        if ids:
            assert np.array_equal(f.ids, g.ids), f"{where} octave {k}: match IDs"
    assert np.array_equal(a["outliers"], b["outliers"]), where
    assert len(a["out_residual"]) == len(b["out_residual"]), where
    if len(a["out_residual"]):
        big = a["out_residual"] > 1e300                 # never evaluated: numeric_limits<double>::max() (S5:296)
        assert np.array_equal(big, np.asarray(b["out_residual"]) > 1e300), where
        np.testing.assert_allclose(a["out_residual"][~big], np.asarray(b["out_residual"])[~big], rtol=1e-9, atol=1e-12)
    if b["pose"] is not None:
        if delta:                                       # the reference keeps deltaPose only as m_last_computed_pose (S5:720-721)
            np.testing.assert_allclose(a["pose"]["delta"], b["pose"]["delta"], atol=POSE_TOL)
        np.testing.assert_allclose(a["pose"]["mat"], b["pose"]["mat"], atol=POSE_TOL)
        np.testing.assert_allclose(a["pose"]["ypr"], b["pose"]["ypr"], atol=1e-10)
    if "tracked_feats_from_last_frame" in b:
        assert a["tracked_feats_from_last_frame"] == b["tracked_feats_from_last_frame"], where
        assert a["tracked_feats_from_last_KF"] == b["tracked_feats_from_last_KF"], where


def run_both(cfg, nf, seed, n_frames, n_oct=1, backend="cv2", ransac="cv2_tolerant", hook=None, **kw):
    frames, cam = synth.make_sequence(cfg, seed=seed, n_frames=n_frames)
    oc = ocam(cam)
    P = orc.Params(orb_nfeats=nf, nOctaves=n_oct, **kw)
    r = ref.RefEstimator(P, backend=backend, ransac=ransac)
    o = (orc.OracleEstimatorOctaves if n_oct > 1 else orc.OracleEstimator)(P, backend=backend, ransac=ransac)
    compared = 0
    try:
        for i, (L, R) in enumerate(frames):
            rep = False
            if hook:
                rep = bool(hook(i, r, o))
            try:
                a = r.process_new_image_pair(L, R, oc, repeat=rep)
            except RuntimeError as e:                     # cv2 4.13's own assertion inside findFundamentalMat
                assert "findFundamentalMat" in str(e)
                with pytest.raises(Exception):
                    o.process_new_image_pair(L, R, oc, repeat=rep)
                break
            b = o.process_new_image_pair(L, R, oc, repeat=rep)
            compare_step(a, b, n_oct, ids=bool(kw.get("vo_use_matches_ids")), where=f"{cfg} seed {seed} frame {i}",
                         delta=P.use_previous_pose_as_initial and bool(b["valid"] or b["error_code"] != 0))
            compared += 1
    finally:
        r.close()
    return compared


@pytest.mark.parametrize("cfg,nf,seed,n_frames", [("c1", 100, 1, 10), ("c1", 100, 100, 8), ("c1", 100, 105, 8),
                                                  ("c2", 500, 2, 4), ("c3", 1000, 3, 3)])
def test_pipeline_equals_reference(cfg, nf, seed, n_frames):
    """processNewImagePair on BASELINE configs 1-3 (FAST+ORB, DescBF, DescBF + RANSAC, Gauss-Newton), real OpenCV calls."""
    assert run_both(cfg, nf, seed, n_frames) >= 3


@pytest.mark.parametrize("cfg,nf,seed,n_frames,n_oct", [("c1", 100, 4, 6, 2), ("c2", 500, 5, 4, 3), ("c3", 1000, 6, 3, 2)])
def test_octaves_equal_reference(cfg, nf, seed, n_frames, n_oct):
    """params_rectify.nOctaves > 1: pyramid, per-octave quota, per-octave stages 2-4, flattened stage 5 (x 2^octave)."""
    assert run_both(cfg, nf, seed, n_frames, n_oct=n_oct) == n_frames


def test_port_backend_equals_reference():
    """the same with the C++ port of FAST / ORB::compute / BFMatcher behind the reference's calls (what -m gpu tests use)"""
    assert run_both("c1", 100, 10, 6, backend="port") == 6


@pytest.mark.parametrize("kw", [dict(enable_robust_1to1_match=True), dict(match_method=1), dict(ifm_method=1),
                                dict(match_method=1, ifm_method=1, enable_robust_1to1_match=True),
                                dict(non_maximal_suppression=False), dict(min_distance=6), dict(max_y_diff=0.0),
                                dict(use_robust_kernel=False), dict(initial_max_iters=3, max_iters=5),
                                dict(residual_threshold=2.0), dict(use_previous_pose_as_initial=False),
                                dict(bad_tracking_th=500), dict(orb_max_distance=25.0), dict(initial_FAST_threshold=30)],
                         ids=lambda d: ",".join(f"{k}={v}" for k, v in d.items()))
def test_parameter_variants_equal_reference(kw):
    """each parameter the hot path reads, flipped away from its default (libstereo-odometry.h:266-508)"""
    assert run_both("c1", 100, 20, 5, **kw) >= 4


def test_match_ids_keyframes_and_reset_equal_reference():
    """vo_use_matches_ids, setThisFrameAsKF, resetIds, setMaxMatchID (libstereo-odometry.h:675-726, S3:172, S4:288-305, P:254-267)"""
    def hook(i, r, o):
        if i == 2:
            r.set_this_frame_as_kf(); o.set_this_frame_as_kf()
        if i == 4:
            r.reset_ids(); o.reset_ids()
        if i == 6:
            r.set_max_match_id(5000); o.set_max_match_id(5000)
    assert run_both("c1", 100, 30, 9, vo_use_matches_ids=True, hook=hook) == 9
    assert run_both("c1", 100, 31, 6, n_oct=2, vo_use_matches_ids=True, hook=hook) == 6


def test_threshold_setters_between_frames_equal_reference():
    """setFASTThreshold / setORBThreshold between frames keep the sequence state (libstereo-odometry.h:529-541): the
    change applies to the next call and tracking continues."""
    def hook(i, r, o):
        if i == 3:
            r.set_fast_threshold(12); o.m_current_fast_th = 12
            r.set_orb_threshold(55); o.m_current_orb_th = 55
            assert r.thresholds() == (12, 55)
        if i == 5:
            r.set_fast_threshold(1000); o.m_current_fast_th = 30        # clamped to fast_max_th (30)
            r.set_orb_threshold(1); o.m_current_orb_th = 30             # clamped to orb_min_th (30)
            assert r.thresholds() == (30, 30)
    assert run_both("c1", 100, 40, 8, hook=hook) == 8


def test_repeat_and_bad_tracking_state_machine_equals_reference():
    """request.repeat and the voecBadTracking retention rule (P:86-100, P:326-330, P:380-381)"""
    frames, cam = synth.make_sequence("c1", seed=50, n_frames=6)
    blank = np.zeros_like(frames[0][0])
    oc = ocam(cam)
    P = orc.Params(orb_nfeats=100)
    r = ref.RefEstimator(P)
    o = orc.OracleEstimator(P)
    seq = [(frames[0], False), (frames[1], False), ((blank, blank), False), (frames[2], False), (frames[2], True), (frames[3], False)]
    for i, ((L, R), rep) in enumerate(seq):
        a = r.process_new_image_pair(L, R, oc, repeat=rep)
        b = o.process_new_image_pair(L, R, oc, repeat=rep)
        compare_step(a, b, 1, where=f"step {i}")
        assert a["state"]["m_error"] == o.m_error and a["state"]["it_counter"] == o.m_it_counter
    r.close()


def test_get_change_in_pose_kats_equal_reference():
    """getChangeInPose (common.cpp:355-413 -> stage5_optimize) on the SURVEY A.7 known-answer inputs, exact and noisy."""
    from katdata import kat_inputs
    for noise in (False, True):
        a, cam, N = kat_inputs(noise, 1.2, 0.9)
        kp = [np.zeros(N, orc.KEYPOINT) for _ in range(4)]
        for k, v in zip(kp, a):
            k["x"], k["y"], k["response"] = v[:, 0], v[:, 1], 1.0
        m = np.zeros(N, orc.DMATCH)
        m["queryIdx"] = m["trainIdx"] = np.arange(N)
        tp = np.stack([np.arange(N), np.arange(N)], 1)
        r = ref.RefEstimator(orc.Params())
        got = r.get_change_in_pose(tp, m, m, kp[0], kp[1], kp[2], kp[3], cam, 640, 480)
        ores, oresid, oinl = orc.stage5(*a, np.ones(N, np.float32), np.arange(N), 480, 640, 3, cam, orc.LSParams.defaults(), np.zeros(6))
        assert (got["valid"], got["error_code"], got["num_it"], got["num_it_final"]) == (bool(ores.valid), ores.error_code, ores.num_it, ores.num_it_final)
        assert np.array_equal(got["outliers"], oinl)
        np.testing.assert_allclose(got["pose"]["delta"], np.array(ores.delta_pose[:]), atol=POSE_TOL)
        np.testing.assert_allclose(got["pose"]["mat"], np.array(ores.out_pose_mat[:]).reshape(4, 4), atol=POSE_TOL)
        np.testing.assert_allclose(got["out_residual"], oresid, rtol=1e-9)
        if not noise:                                     # the known answer itself: the inverse of the simulated motion
            np.testing.assert_allclose(got["pose"]["delta"], [0.01, -0.02, 0.005, 0.05, -0.01, 0.10], atol=2e-4)
        r.close()


def test_non_max_sup_tie_order_equals_reference():
    """m_non_max_sup (stage2_detect.cpp:296-370) called directly: std::sort with KeypointResponseSorter on many equal
    responses (the permutation is libstdc++'s), grid cells of 1 and 3 pixels, quota below and above the candidate count."""
    rng = np.random.default_rng(3)
    r = ref.RefEstimator(orc.Params())
    try:
        for case in range(12):
            n = int(rng.integers(50, 3000))
            H, W = 240, 320
            kps = np.zeros(n, orc.KEYPOINT)
            kps["x"] = rng.integers(31, W - 31, n)
            kps["y"] = rng.integers(31, H - 31, n)
            kps["response"] = rng.integers(20, 20 + (4 if case % 2 else 60), n)        # heavy ties in the odd cases
            kps["size"], kps["angle"], kps["class_id"] = 7, -1, -1
            desc = rng.integers(0, 256, (n, 32)).astype(np.uint8)
            for md in (3, 6):
                r.L.svo_ref_set_param(r.h, b"min_distance", float(md))
                for quota in (n // 3, n + 10):
                    k1, d1 = r.non_max_sup(quota, kps, desc, H, W)
                    k2, d2, _ = orc.non_max_sup(quota, kps, desc, md, H, W)
                    assert np.array_equal(k1, k2) and np.array_equal(d1, d2), (case, md, quota)
            k1, d1, idx1 = r.update_indexes(kps, desc, H, W)
            k2, d2, _, idx2 = orc.update_indexes(kps, desc, H)
            assert np.array_equal(k1, k2) and np.array_equal(d1, d2) and np.array_equal(idx1, idx2), case
    finally:
        r.close()


def test_aux_non_max_sup_equals_reference():
    """the auxiliary m_non_max_sup (stage2_detect.cpp:225-283) stage 5 runs on the previous left keypoints (S5:465-474)"""
    rng = np.random.default_rng(4)
    r = ref.RefEstimator(orc.Params())
    try:
        for case in range(8):
            n = int(rng.integers(20, 800))
            kps = np.zeros(n, orc.KEYPOINT)
            kps["x"] = rng.integers(0, 60, n) * (2.0 if case % 2 else 1.0)             # rescaled octave points collide
            kps["y"] = rng.integers(0, 40, n) * (2.0 if case % 2 else 1.0)
            kps["response"] = rng.integers(20, 30, n)
            s1 = r.non_max_sup_aux(kps, 240, 320, n)
            xy = np.stack([kps["x"], kps["y"]], 1).astype(np.float32)
            assert np.array_equal(s1, orc.aux_nms_mask(xy, kps["response"].astype(np.float32), 240, 320, 3)), case
    finally:
        r.close()


def test_pinhole_projection_and_jacobian_equal_reference():
    """m_pinhole_stereo_projection (stage5_optimization.cpp:35-257) incl. the small-angle branch and the typo at :162"""
    rng = np.random.default_rng(5)
    r = ref.RefEstimator(orc.Params())
    cam = orc.make_camera(500, 500, 320, 240, 0.12)
    try:
        lm = np.stack([rng.uniform(-3, 3, 64), rng.uniform(-2, 2, 64), rng.uniform(2, 12, 64)], 1)
        for pose in ([0.0] * 6, [1e-7, -2e-7, 1e-7, 0.01, 0, 0.02], [0.01, -0.02, 0.005, 0.05, -0.01, 0.1], [0.3, 0.2, -0.4, 1, 2, 3]):
            px, J = r.pinhole_projection(lm, cam, pose)
            px2, J2 = orc.pinhole_projection(lm, cam, np.array(pose, np.float64))
            assert np.array_equal(px, px2)
            np.testing.assert_allclose(J, J2, rtol=1e-13, atol=1e-13)
    finally:
        r.close()


def test_dmorb_branch_equals_reference():
    """detect_method = dmORB (stage2_detect.cpp:458-497): cv::ORB::detectAndCompute with 1.5 x orb_nfeats, then the glue"""
    pytest.importorskip("cv2")
    assert run_both("c1", 100, 60, 4, detect_method=0) == 4


@pytest.mark.parametrize("kw", [dict(match_method=1), dict(ifm_method=1), dict(match_method=1, ifm_method=1)])
def test_dmorb_with_the_band_limited_matchers_equals_reference(kw):
    """dmORB's float keypoints through smDescRbR / ifmDescWin: the float comparisons of those branches"""
    pytest.importorskip("cv2")
    assert run_both("c1", 100, 61, 4, detect_method=0, **kw) == 4


def test_reference_dump_files(tmp_path):
    """vo_save_files: the reference's own writers (P:179-204,278-287, S4:80-82,173-234,308-327, S5:477-492,703-713)
    produce the per-frame files the host shim reproduces; their contents agree with the state read back."""
    frames, cam = synth.make_sequence("c1", seed=70, n_frames=3)
    oc = ocam(cam)
    r = ref.RefEstimator(orc.Params(orb_nfeats=100))
    r.set_save_files(str(tmp_path))
    try:
        for L, R in frames:
            out = r.process_new_image_pair(L, R, oc)
    finally:
        r.close()
    names = sorted(os.listdir(tmp_path))
    for want in ("left_feats_0002.txt", "right_feats_0002.txt", "matches_0002.txt", "inter_frame_matching_0002.txt",
                 "l-l_0002.txt", "r-r_0002.txt", "out_residual_0002.txt", "out_residual_final_0002.txt", "optimization.txt"):
        assert want in names, names
    ifm = np.loadtxt(tmp_path / "inter_frame_matching_0002.txt").reshape(-1, 11)
    assert np.array_equal(ifm[:, 9:11].astype(np.int64), out["tracked"][0])
    m = np.loadtxt(tmp_path / "matches_0002.txt").reshape(-1, 4)
    assert np.array_equal(m[:, 1].astype(np.int32), out["frame"].matches["queryIdx"])
