"""CPU tests of the multi-octave oracle (params_rectify.nOctaves > 1, the reference's ctor default of 3,
stage1_rectify.cpp:28,80-83): pyramid restatements, per-octave quotas, and that the octave-aware driver reduces to the
single-octave one for nOctaves = 1."""
import numpy as np
import pytest

from oracle import oracle as orc
from stereo_vo_b200 import synth

cv2 = pytest.importorskip("cv2")


def ocam(cam):
    return orc.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])


@pytest.mark.parametrize("shape", [(240, 320), (375, 1242), (187, 621), (65, 67), (64, 64), (3, 5)])
def test_pyr_down5_is_cv2_pyrdown(shape):
    rng = np.random.default_rng(shape[0] * 31 + shape[1])
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    img[: shape[0] // 2, : shape[1] // 3] = 255          # saturated block: rounding at the top of the range
    h2, w2 = shape[0] >> 1, shape[1] >> 1
    ref = cv2.pyrDown(img, dstsize=(w2, h2))
    assert np.array_equal(orc.pyr_down5(img), ref)


def test_pyr_half_avg_known_answers():
    img = np.array([[0, 1, 255, 255, 10, 11], [0, 0, 255, 254, 12, 13], [9, 9, 9, 9, 9, 9]], np.uint8)
    # rows averaged first with round-half-up, then columns with round-half-up; the odd last row is dropped
    assert orc.pyr_half_avg(img).tolist() == [[1, 255, 12]]
    rng = np.random.default_rng(1)
    big = rng.integers(0, 256, (48, 64), dtype=np.uint8)
    half = orc.pyr_half_avg(big)
    assert half.shape == (24, 32)
    exact = big.reshape(24, 2, 32, 2).astype(np.float64).mean(axis=(1, 3))
    assert np.all(half >= np.floor(exact)) and np.all(half <= np.floor(exact) + 1)   # the double rounding never rounds down


def test_build_pyramid_modes():
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, (240, 320), dtype=np.uint8)
    pyr = orc.build_pyramid(img, 3)
    assert [p.shape for p in pyr] == [(240, 320), (120, 160), (60, 80)]
    assert np.array_equal(pyr[1], orc.pyr_half_avg(img))                 # 320 % 16 == 0: the SSE2 path
    odd = rng.integers(0, 256, (375, 1242), dtype=np.uint8)
    pyr = orc.build_pyramid(odd, 2)
    assert np.array_equal(pyr[1], cv2.pyrDown(odd, dstsize=(621, 187)))  # 1242 % 16 != 0: cvPyrDown
    assert np.array_equal(orc.build_pyramid(img, 2, orc.PYR_PYRDOWN)[1], cv2.pyrDown(img))


def test_kps_to_detect():
    assert orc.kps_to_detect(500, 1) == [1000]
    assert orc.kps_to_detect(500, 3) == [428, 214, 107]
    assert orc.kps_to_detect(100, 2) == [133, 67]                       # 66.5 rounds half away from zero
    assert orc.kps_to_detect(1000, 4) == [533, 267, 133, 67]


@pytest.mark.parametrize("ids", [False, True])
def test_one_octave_equals_the_single_octave_driver(ids):
    frames, cam = synth.make_sequence("c1", seed=5, n_frames=4)
    p = orc.Params(orb_nfeats=100, vo_use_matches_ids=ids)
    a = orc.OracleEstimator(p, backend="port", ransac="off")
    b = orc.OracleEstimatorOctaves(orc.Params(orb_nfeats=100, vo_use_matches_ids=ids), backend="port", ransac="off")
    for i, (L, R) in enumerate(frames):
        if ids and i == 2:
            a.reset_ids(); b.reset_ids()
        ra = a.process_new_image_pair(L, R, ocam(cam))
        rb = b.process_new_image_pair(L, R, ocam(cam))
        assert ra["error_code"] == rb["error_code"] and ra["valid"] == rb["valid"]
        assert np.array_equal(ra["frame"].kpsL, rb["frames"][0].kpsL)
        assert np.array_equal(ra["frame"].matches, rb["frames"][0].matches)
        assert np.array_equal(ra["tracked"], rb["tracked"][0])
        assert np.array_equal(ra["outliers"], rb["outliers"])
        assert np.array_equal(ra["out_residual"], rb["out_residual"])
        if ids:
            assert np.array_equal(ra["frame"].ids, rb["frames"][0].ids)
        if ra["pose"] is not None:
            assert np.array_equal(ra["pose"]["delta"], rb["pose"]["delta"])


def test_three_octaves_track_and_estimate_the_motion():
    frames, cam = synth.make_sequence("c2", seed=1, n_frames=3)
    est = orc.OracleEstimatorOctaves(orc.Params(orb_nfeats=500, nOctaves=3, vo_use_matches_ids=True), backend="port",
                                     ransac="off")
    r0 = est.process_new_image_pair(*frames[0], ocam(cam))
    assert r0["error_code"] == orc.VOEC["voecFirstIteration"]
    assert [d[0] for d in r0["detected_feats"]] == [428, 214, 107]
    # first-frame IDs run through the octaves in order
    ids = np.concatenate([f.ids for f in r0["frames"]])
    assert np.array_equal(ids, np.arange(sum(r0["stereo_matches"])))
    r1 = est.process_new_image_pair(*frames[1], ocam(cam))
    assert r1["valid"] and all(len(t) > 20 for t in r1["tracked"])
    assert r1["tracked_feats_from_last_frame"] == sum(len(t) for t in r1["tracked"])
    assert len(r1["out_residual"]) == r1["tracked_feats_from_last_frame"]
    # ground truth: 0.05 m along x and 0.2 degrees of yaw per frame
    assert abs(r1["pose"]["t"][0] - 0.05) < 5e-3 and abs(abs(r1["pose"]["rotvec"][1]) - np.deg2rad(0.2)) < 5e-4
    # the coarser octaves enter stage 5 at the scale of octave 0
    g = r1["gn_dbg"]
    n0 = len(r1["tracked"][0])
    assert g["l1l"][n0:, 0].max() > 320 and np.all(g["l1l"][n0:n0 + len(r1["tracked"][1])] % 2 == 0)


def test_oracle_port_reproduces_the_opencv_generated_octave_goldens(golden_octaves):
    """tests/golden/golden_octaves.npz was generated with the real OpenCV calls on every pyramid level
    (make_golden_octaves.py); the all-C++ restatement must reproduce it bit for bit."""
    g = golden_octaves
    frames, cam = synth.make_sequence("c1", seed=3, n_frames=3)
    est = orc.OracleEstimatorOctaves(orc.Params(orb_nfeats=100, nOctaves=2, vo_use_matches_ids=True), backend="port", ransac="off")
    for i, (L, R) in enumerate(frames):
        o = est.process_new_image_pair(L, R, ocam(cam))
        for k, f in enumerate(o["frames"]):
            assert np.array_equal(np.stack([f.kpsL["x"], f.kpsL["y"], f.kpsL["response"]], 1).astype(np.int32), g[f"f{i}_o{k}_kpsL"])
            assert np.array_equal(f.descL, g[f"f{i}_o{k}_descL"]) and np.array_equal(f.descR, g[f"f{i}_o{k}_descR"])
            assert np.array_equal(np.stack([f.matches["queryIdx"], f.matches["trainIdx"], f.matches["distance"].astype(np.int32)], 1),
                                  g[f"f{i}_o{k}_matches"])
            assert np.array_equal(f.ids, g[f"f{i}_o{k}_ids"])
            if i > 0:
                assert np.array_equal(o["tracked"][k], g[f"f{i}_o{k}_tracked"])
        assert [o["error_code"], int(o["valid"]), o["num_it"], o["num_it_final"]] == g[f"f{i}_status"].tolist()
        if i > 0:
            assert np.array_equal(o["pose"]["delta"], g[f"f{i}_delta"]) and np.array_equal(o["outliers"], g[f"f{i}_outliers"])
    f3, _ = synth.make_sequence("c3", seed=2, n_frames=1)
    assert np.array_equal(orc.build_pyramid(f3[0][0], 2)[1], g["c3_pyrdown_L"])
    assert np.array_equal(orc.build_pyramid(f3[0][1], 2)[1], g["c3_pyrdown_R"])
