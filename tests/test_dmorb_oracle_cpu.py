"""Groundwork for SURVEY section 8(f) rank 4 (detect_method = dmORB, stage2_detect.cpp:458-496): the oracle runs that
branch with the real OpenCV call the reference makes (cv::ORB::detectAndCompute, oriented, 8 levels, Harris score) and
the reference's own glue downstream.  No device implementation exists yet; this pins what one has to reproduce."""
import numpy as np
import pytest

from oracle import oracle as orc
from stereo_vo_b200 import synth

cv2 = pytest.importorskip("cv2")


def test_dmorb_branch_runs_end_to_end_with_the_real_opencv():
    frames, cam = synth.make_sequence("c2", seed=0, n_frames=3)
    oc = orc.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])
    est = orc.OracleEstimator(orc.Params(orb_nfeats=500, detect_method=0, max_y_diff=2.0), backend="cv2", ransac="cv2")
    outs = [est.process_new_image_pair(L, R, oc) for (L, R) in frames]
    f = outs[0]["frame"]
    assert outs[0]["error_code"] == orc.VOEC["voecFirstIteration"]
    assert 400 < len(f.kpsL) <= 750                       # 1.5 x orb_nfeats extracted, thinned by m_non_max_sup (S2:462-465, 586-597)
    assert set(f.kpsL["octave"]) == set(range(8))         # all 8 levels in one list: nOctaves is forced to 1 (S1:80)
    assert np.all(np.diff(f.kpsL["y"]) >= 0)              # m_update_indexes: sorted by row (S2:65-130)
    assert (f.kpsL["angle"] >= 0).all()                   # oriented keypoints (FAST keypoints carry angle -1)
    for o in outs[1:]:
        assert o["valid"] and len(o["tracked"]) > 40
        assert abs(o["pose"]["t"][0] - 0.05) < 6e-3       # ground truth: 0.05 m along x per frame


def test_dmorb_needs_the_opencv_backend():
    est = orc.OracleEstimator(orc.Params(detect_method=0), backend="port", ransac="off")
    with pytest.raises(AssertionError):
        est.detect(np.zeros((240, 320), np.uint8))
