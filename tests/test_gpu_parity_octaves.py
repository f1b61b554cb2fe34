"""GPU parity of the multi-octave path (params_rectify.nOctaves > 1; the reference's ctor default is 3,
stage1_rectify.cpp:28,80-83) against the octave-aware oracle: pyramid planes, per-octave keypoints / descriptors /
stereo matches / match IDs / tracked pairs bit-exact, the flattened stage-5 problem (x 2^octave rescale,
stage5_optimization.cpp:419-461) within the pose tolerance north_star states (1e-4 rad, 1e-3 m)."""
import numpy as np
import pytest

from oracle import oracle as orc
from stereo_vo_b200 import capi, synth

pytestmark = pytest.mark.gpu

POSE_TOL_RAD = 1e-4
POSE_TOL_M = 1e-3


def xys(k):
    return np.stack([k["x"], k["y"], k["response"]], 1).astype(np.int32)


def mtuple(m):
    return np.stack([m["queryIdx"], m["trainIdx"], m["distance"].astype(np.int32)], 1)


def ocam(cam):
    return orc.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])


def gcam(cam):
    return capi.make_camera(cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"])


def device_masks(ctx):
    """The oracle takes its two findFundamentalMat masks from the device kernel (whose bit-exactness against OpenCV is
    the subject of tests/test_gpu_parity.py): these tests are about the octave plumbing around it."""
    def fn(p1, p2):
        if len(p1) < 7:
            return None
        return ctx.fundmat_ransac(p1, p2)[1]
    return fn


def check_frame(ctx, o, r, slot, i, ids):
    n_oct = ctx.n_octaves
    stats = ctx.octave_stats(slot)
    for k in range(n_oct):
        f = o["frames"][k]
        assert np.array_equal(xys(ctx.keypoints(slot, 0, 0, octave=k)), xys(f.kpsL)), f"frame {i} octave {k}: left keypoints"
        assert np.array_equal(xys(ctx.keypoints(slot, 0, 1, octave=k)), xys(f.kpsR)), f"frame {i} octave {k}: right keypoints"
        assert np.array_equal(ctx.descriptors(slot, 0, 0, octave=k), f.descL), f"frame {i} octave {k}: left descriptors"
        assert np.array_equal(ctx.descriptors(slot, 0, 1, octave=k), f.descR), f"frame {i} octave {k}: right descriptors"
        assert np.array_equal(mtuple(ctx.matches(slot, 0, octave=k)), mtuple(f.matches)), f"frame {i} octave {k}: stereo matches"
        assert tuple(stats[k][:3]) == (len(f.kpsL), len(f.kpsR), len(f.matches))
        if ids:
            assert np.array_equal(ctx.match_ids(slot, 0, octave=k), f.ids), f"frame {i} octave {k}: match IDs"
        if "track_dbg" in o:
            assert np.array_equal(ctx.tracked_pairs(slot, octave=k), o["tracked"][k]), f"frame {i} octave {k}: tracked pairs"
            assert stats[k][3] == len(o["tracked"][k])
    f0 = o["frames"][0]
    assert (r.detected_left, r.detected_right, r.stereo_matches) == (len(f0.kpsL), len(f0.kpsR), len(f0.matches))
    if "track_dbg" in o:
        assert r.tracked_feats_from_last_frame == o["tracked_feats_from_last_frame"]
        assert r.tracked_feats_from_last_KF == o["tracked_feats_from_last_KF"]
    assert (r.error_code, bool(r.valid), r.num_it, r.num_it_final) == (o["error_code"], o["valid"], o["num_it"], o["num_it_final"]), \
        f"frame {i}: status {list(r.delta_pose)} vs {None if o['pose'] is None else o['pose']['delta']}"
    if o["pose"] is not None:
        np.testing.assert_allclose(r.delta_pose[:3], o["pose"]["delta"][:3], atol=POSE_TOL_RAD)
        np.testing.assert_allclose(r.delta_pose[3:], o["pose"]["delta"][3:], atol=POSE_TOL_M)
        np.testing.assert_allclose(r.out_pose_rotvec[:], o["pose"]["rotvec"], atol=POSE_TOL_RAD)
        np.testing.assert_allclose(r.out_pose_t[:], o["pose"]["t"], atol=POSE_TOL_M)
        assert np.array_equal(ctx.outliers(slot), o["outliers"]), f"frame {i}: inlier list"
        np.testing.assert_allclose(ctx.residuals(slot), o["out_residual"], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("config,n_oct,pyr_mode", [("c1", 2, 0), ("c1", 2, 2), ("c2", 3, 0), ("c3", 2, 0), ("c3", 2, 1)])
def test_pyramid_planes_bit_exact(config, n_oct, pyr_mode):
    cfg = synth.CONFIGS[config]
    frames, cam = synth.make_sequence(config, seed=2, n_frames=1)
    ctx = capi.Context(width=cfg["W"], height=cfg["H"], orb_nfeats=cfg["orb_nfeats"], max_seqs=2, max_corners=32768,
                       n_octaves=n_oct, pyr_mode=pyr_mode)
    try:
        L, R = frames[0]
        ctx.process_batch([L, R], [R, L], [gcam(cam)] * 2)          # two items: the second one's planes sit behind the first's
        pyrL = orc.build_pyramid(L, n_oct, pyr_mode)
        pyrR = orc.build_pyramid(R, n_oct, pyr_mode)
        for k in range(1, n_oct):
            assert ctx.octave_info[k][:2] == (cfg["W"] >> k, cfg["H"] >> k)
            assert np.array_equal(ctx.pyramid_plane(0, k, 0), pyrL[k]), f"octave {k} left"
            assert np.array_equal(ctx.pyramid_plane(0, k, 1), pyrR[k]), f"octave {k} right"
            assert np.array_equal(ctx.pyramid_plane(1, k, 0), pyrR[k]) and np.array_equal(ctx.pyramid_plane(1, k, 1), pyrL[k])
    finally:
        ctx.close()


@pytest.mark.parametrize("config,n_oct,ids,n_frames", [("c2", 3, True, 4), ("c1", 2, False, 3), ("c3", 2, True, 2)])
def test_pipeline_octaves_vs_oracle(config, n_oct, ids, n_frames):
    cfg = synth.CONFIGS[config]
    frames, cam = synth.make_sequence(config, seed=4, n_frames=n_frames)
    ctx = capi.Context(width=cfg["W"], height=cfg["H"], orb_nfeats=cfg["orb_nfeats"], max_seqs=2, max_corners=32768,
                       n_octaves=n_oct, vo_use_matches_ids=int(ids))
    try:
        assert [o[2] for o in ctx.octave_info] == orc.kps_to_detect(cfg["orb_nfeats"], n_oct)
        est = orc.OracleEstimatorOctaves(orc.Params(orb_nfeats=cfg["orb_nfeats"], nOctaves=n_oct, vo_use_matches_ids=ids),
                                         backend="port", ransac=device_masks(ctx))
        last = None
        for i, (L, R) in enumerate(frames):
            if ids and i == 2:                     # resetIds(): renumbering across all octaves (P:254-267)
                est.reset_ids()
                ctx.reset_ids(1)
            o = est.process_new_image_pair(L, R, ocam(cam))
            r = ctx.process_batch([L], [R], [gcam(cam)], slots=[1])[0]
            check_frame(ctx, o, r, 1, i, ids)
            if ids and i > 0:
                for k in range(n_oct):             # previous frame's IDs (renumbered after the reset)
                    assert np.array_equal(ctx.match_ids(1, 1, octave=k), est.prev[k].ids)
            last = (o, r)
        o, r = last
        assert r.valid and all(len(t) > 10 for t in o["tracked"])
        assert abs(r.out_pose_t[0] - 0.05) < 5e-3
        assert len(ctx.residuals(1)) == sum(len(t) for t in o["tracked"])
    finally:
        ctx.close()


def test_pipeline_octaves_with_the_band_limited_matchers():
    """smDescRbR + ifmDescWin (S3:208-408, S4:458-735 loop over the octaves like the brute-force branches)."""
    frames, cam = synth.make_sequence("c2", seed=6, n_frames=3)
    ctx = capi.Context(width=640, height=480, orb_nfeats=500, max_seqs=1, n_octaves=3, vo_use_matches_ids=1,
                       match_method=1, if_match_method=1)
    try:
        est = orc.OracleEstimatorOctaves(orc.Params(orb_nfeats=500, nOctaves=3, vo_use_matches_ids=True, match_method=1,
                                                    ifm_method=1), backend="port", ransac=device_masks(ctx))
        for i, (L, R) in enumerate(frames):
            o = est.process_new_image_pair(L, R, ocam(cam))
            r = ctx.process_batch([L], [R], [gcam(cam)])[0]
            check_frame(ctx, o, r, 0, i, True)
        assert r.valid and all(len(t) > 10 for t in o["tracked"])
    finally:
        ctx.close()


def test_bad_tracking_gate_counts_all_octaves():
    """P:326-330 compares the TOTAL over the octaves with bad_tracking_th; an unrelated second frame trips it, the
    frame after that is matched against the kept previous frame (P:86-89)."""
    a, cam = synth.make_sequence("c1", seed=7, n_frames=3)
    b, _ = synth.make_sequence("c1", seed=8, n_frames=1)
    frames = [a[0], b[0], a[1], a[2]]
    ctx = capi.Context(width=320, height=240, orb_nfeats=100, max_seqs=1, n_octaves=2)
    try:
        est = orc.OracleEstimatorOctaves(orc.Params(orb_nfeats=100, nOctaves=2), backend="port", ransac=device_masks(ctx))
        codes = []
        for i, (L, R) in enumerate(frames):
            o = est.process_new_image_pair(L, R, ocam(cam))
            r = ctx.process_batch([L], [R], [gcam(cam)])[0]
            check_frame(ctx, o, r, 0, i, False)
            codes.append(r.error_code)
        assert codes[0] == capi.VOEC["voecFirstIteration"] and codes[1] == capi.VOEC["voecBadTracking"]
        assert codes[2] == capi.VOEC["voecNone"]
    finally:
        ctx.close()


def test_octave_batch_in_concurrent_groups_equals_slots_alone():
    """64 slots (two groups on two streams) over 4 distinct sequences: every slot gets what the sequence gets alone."""
    seqs = [synth.make_sequence("c1", seed=40 + s, n_frames=3) for s in range(4)]
    key = lambda r: (tuple(r.delta_pose), r.num_it, r.num_it_final, r.stereo_matches, r.tracked_feats_from_last_frame,
                     r.error_code, r.valid)
    ctx = capi.Context(width=320, height=240, orb_nfeats=100, max_seqs=64, n_octaves=2)
    try:
        alone = []
        for s, (frames, cam) in enumerate(seqs):
            ctx.reset_slot(0)
            alone.append([(key(ctx.process_batch([L], [R], [gcam(cam)], slots=[0])[0]), ctx.octave_stats(0).tolist())
                          for (L, R) in frames])
        for s in range(64):
            ctx.reset_slot(s)
        order = list(range(63, -1, -1))            # slot list in reverse: item i is slot 63 - i
        for f in range(3):
            Ls = [seqs[i % 4][0][f][0] for i in range(64)]
            Rs = [seqs[i % 4][0][f][1] for i in range(64)]
            rs = ctx.process_batch(Ls, Rs, [gcam(seqs[i % 4][1]) for i in range(64)], slots=order)
            for i, r in enumerate(rs):
                assert key(r) == alone[i % 4][f][0], f"frame {f} item {i}"
                assert ctx.octave_stats(order[i]).tolist() == alone[i % 4][f][1]
        assert any(a[0][6] for a in alone[0][1:])   # the sequences do produce valid poses
    finally:
        ctx.close()


def test_get_change_in_pose_on_a_multi_octave_context():
    """getChangeInPose (C:355-413) is single-octave by construction (`const size_t octave = 0`); on a context with
    octaves it must give what a one-octave context gives."""
    frames, cam = synth.make_sequence("c1", seed=9, n_frames=2)
    est = orc.OracleEstimator(orc.Params(orb_nfeats=100), backend="port", ransac="off")
    est.process_new_image_pair(*frames[0], ocam(cam))
    o = est.process_new_image_pair(*frames[1], ocam(cam))
    prev, cur = est.prev, est.cur
    got = []
    for n_oct in (1, 2):
        ctx = capi.Context(width=320, height=240, orb_nfeats=100, max_seqs=1, n_octaves=n_oct)
        try:
            r = ctx.get_change_in_pose(o["tracked"], prev.matches, cur.matches, prev.kpsL, prev.kpsR, cur.kpsL, cur.kpsR, gcam(cam))
            got.append((tuple(r.delta_pose), r.num_it, r.num_it_final, r.valid, tuple(ctx.outliers(0)), tuple(ctx.residuals(0))))
        finally:
            ctx.close()
    assert got[0] == got[1] and got[0][3] == 1
    np.testing.assert_allclose(got[0][0][3:], o["pose"]["delta"][3:], atol=POSE_TOL_M)


def test_host_shim_demo_with_two_octaves(tmp_path):
    """The C++ host shim with params_rectify.nOctaves = 2 (demo loop of demo-main.cpp:210-253): composed camera poses
    equal the octave-aware oracle's, and the per-octave result vectors have one entry per octave."""
    import os
    import struct
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    demo = os.path.join(root, "stereo_vo_b200", "host", "svo_demo")
    assert os.path.exists(demo), "host shim not built (run __graft_entry__.build())"
    frames, cam = synth.make_sequence("c1", seed=9, n_frames=4)
    seq = tmp_path / "seq.bin"
    with open(seq, "wb") as f:
        f.write(struct.pack("<3i5d", 320, 240, len(frames), cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"]))
        for (L, R) in frames:
            f.write(L.tobytes())
            f.write(R.tobytes())
    out = tmp_path / "camera_pose.txt"
    p = subprocess.run([demo, str(seq), str(out), "100", "2"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    got = np.loadtxt(out).reshape(-1, 6)
    ctx = capi.Context(width=320, height=240, orb_nfeats=100, max_seqs=1, n_octaves=2)
    try:
        est = orc.OracleEstimatorOctaves(orc.Params(orb_nfeats=100, nOctaves=2), backend="port", ransac=device_masks(ctx))
        pose = np.eye(4)
        exp = []
        for (L, R) in frames:
            o = est.process_new_image_pair(L, R, ocam(cam))
            if o["valid"]:
                pose = pose @ o["pose"]["mat"]
            ypr = [np.arctan2(pose[1, 0], pose[0, 0]), np.arctan2(-pose[2, 0], np.hypot(pose[0, 0], pose[1, 0])),
                   np.arctan2(pose[2, 1], pose[2, 2])]
            exp.append([pose[0, 3], pose[1, 3], pose[2, 3]] + ypr)
    finally:
        ctx.close()
    np.testing.assert_allclose(got, np.array(exp), atol=1.5e-3)     # the file holds %.3f values
    assert abs(got[-1, 0] - 0.15) < 0.02


def test_host_shim_vo_save_files_dumps(tmp_path):
    """params_general.vo_save_files (P:179-203, 278-287; S4:81-82, 308-328; S5:703-713): the per-frame text dumps of the
    reference, written by the shim from the device state, two octaves; the descriptor .yml must parse with
    cv::FileStorage (cv2) and hold the descriptors the C-ABI readers return."""
    import os
    import struct
    import subprocess
    cv2 = pytest.importorskip("cv2")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    demo = os.path.join(root, "stereo_vo_b200", "host", "svo_demo")
    frames, cam = synth.make_sequence("c1", seed=6, n_frames=3)
    seq = tmp_path / "seq.bin"
    with open(seq, "wb") as f:
        f.write(struct.pack("<3i5d", 320, 240, len(frames), cam["fx"], cam["fy"], cam["cx"], cam["cy"], cam["baseline"]))
        for (L, R) in frames:
            f.write(L.tobytes())
            f.write(R.tobytes())
    outdir = tmp_path / "dumps"
    text = open(os.path.join(root, "stereo_vo_b200", "host", "hot_path.ini")).read()
    text = text.replace("orb_nfeats = 500", "orb_nfeats = 100").replace("nOctaves = 1", "nOctaves = 2")
    text = text.replace("vo_save_files = false", f"vo_save_files = true\nvo_out_dir = {outdir}")
    ini = tmp_path / "cfg.ini"
    ini.write_text(text)
    p = subprocess.run([demo, str(seq), str(tmp_path / "camera_pose.txt"), str(ini)], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    ctx = capi.Context(width=320, height=240, orb_nfeats=100, max_seqs=1, n_octaves=2)
    try:
        for i, (L, R) in enumerate(frames):
            r = ctx.process_batch([L], [R], [gcam(cam)])[0]
            kl = [ctx.keypoints(0, 0, 0, octave=k) for k in range(2)]
            feats = np.loadtxt(outdir / f"left_feats_{i:04d}.txt").reshape(-1, 4)
            exp = np.concatenate([np.stack([np.full(len(k), o), k["x"], k["y"], k["response"]], 1) for o, k in enumerate(kl)])
            assert np.array_equal(feats, exp)
            assert len(np.loadtxt(outdir / f"right_feats_{i:04d}.txt").reshape(-1, 4)) == ctx.octave_stats(0)[:, 1].sum()
            for k in range(2):
                for side, name in ((0, "left"), (1, "right")):
                    fs = cv2.FileStorage(str(outdir / f"{name}_desc_o{k:02d}_{i:04d}.yml"), cv2.FILE_STORAGE_READ)
                    assert np.array_equal(fs.getNode(f"{name}DescMat").mat(), ctx.descriptors(0, 0, side, octave=k))
            m = np.loadtxt(outdir / f"matches_{i:04d}.txt").reshape(-1, 4)
            expm = np.concatenate([np.stack([np.full(len(mm), o), mm["queryIdx"], mm["trainIdx"], mm["distance"]], 1)
                                   for o, mm in enumerate(ctx.matches(0, 0, octave=k) for k in range(2))])
            assert np.array_equal(m, expm)
            if i == 0:
                assert not (outdir / "inter_frame_matching_0000.txt").exists()
                continue
            ifm = np.loadtxt(outdir / f"inter_frame_matching_{i:04d}.txt").reshape(-1, 11)
            tps = [ctx.tracked_pairs(0, octave=k) for k in range(2)]
            assert len(ifm) == r.tracked_feats_from_last_frame == sum(len(t) for t in tps)
            assert np.array_equal(ifm[:, [0, 9, 10]], np.concatenate([np.column_stack([np.full(len(t), o), t]) for o, t in enumerate(tps)]))
            cl = kl[0][ctx.matches(0, 0, octave=0)["queryIdx"][tps[0][:, 1]]]
            assert np.array_equal(ifm[:len(tps[0]), 5], cl["x"]) and np.array_equal(ifm[:len(tps[0]), 6], cl["y"])
            res = np.loadtxt(outdir / f"out_residual_{i:04d}.txt").reshape(-1)
            np.testing.assert_allclose(res, ctx.residuals(0), atol=5.1e-4)          # %.3f
            assert np.array_equal(res, np.loadtxt(outdir / f"out_residual_final_{i:04d}.txt").reshape(-1))
    finally:
        ctx.close()


def test_build_pyramid_device_alone():
    """svo_build_pyramid_device (the pyramid of stage1_rectify.cpp:80-83 on images already in HBM) == the oracle's
    pyramid; a one-octave context refuses the call."""
    import torch
    frames, cam = synth.make_sequence("c2", seed=3, n_frames=1)
    L, R = frames[0]
    ctx = capi.Context(width=640, height=480, orb_nfeats=500, max_seqs=3, n_octaves=3)
    try:
        dL = torch.from_numpy(np.stack([L, R, L])).cuda()
        dR = torch.from_numpy(np.stack([R, L, R])).cuda()
        torch.cuda.synchronize()                   # the library works on its own stream
        ctx.build_pyramid_device(dL.data_ptr(), dR.data_ptr(), 3, 640, 640 * 480)
        pl, pr = orc.build_pyramid(L, 3), orc.build_pyramid(R, 3)
        for k in (1, 2):
            assert np.array_equal(ctx.pyramid_plane(0, k, 0), pl[k]) and np.array_equal(ctx.pyramid_plane(0, k, 1), pr[k])
            assert np.array_equal(ctx.pyramid_plane(1, k, 0), pr[k]) and np.array_equal(ctx.pyramid_plane(2, k, 1), pr[k])
    finally:
        ctx.close()
    one = capi.Context(width=640, height=480, orb_nfeats=500, max_seqs=1)
    try:
        with pytest.raises(capi.SvoError) as e:
            one.build_pyramid_device(dL.data_ptr(), dR.data_ptr(), 1, 640, 640 * 480)
        assert e.value.code == capi.SVO_ERR_STATE
    finally:
        one.close()


def test_octaves_against_opencv_generated_goldens(golden_octaves):
    """The device against tests/golden/golden_octaves.npz (real OpenCV FAST / ORB::compute / BFMatcher on every pyramid
    level, F-matrix filter off): per-octave keypoints, descriptors, stereo matches, match IDs, tracked pairs, status,
    pose; and the cvPyrDown level of a config-3 image against cv2.pyrDown's own output."""
    g = golden_octaves
    frames, cam = synth.make_sequence("c1", seed=3, n_frames=3)
    ctx = capi.Context(width=320, height=240, orb_nfeats=100, max_seqs=1, n_octaves=2, vo_use_matches_ids=1, fundmat_filter=0)
    try:
        for i, (L, R) in enumerate(frames):
            r = ctx.process_batch([L], [R], [gcam(cam)])[0]
            for k in range(2):
                assert np.array_equal(xys(ctx.keypoints(0, 0, 0, octave=k)), g[f"f{i}_o{k}_kpsL"])
                assert np.array_equal(xys(ctx.keypoints(0, 0, 1, octave=k)), g[f"f{i}_o{k}_kpsR"])
                assert np.array_equal(ctx.descriptors(0, 0, 0, octave=k), g[f"f{i}_o{k}_descL"])
                assert np.array_equal(ctx.descriptors(0, 0, 1, octave=k), g[f"f{i}_o{k}_descR"])
                assert np.array_equal(mtuple(ctx.matches(0, 0, octave=k)), g[f"f{i}_o{k}_matches"])
                assert np.array_equal(ctx.match_ids(0, 0, octave=k), g[f"f{i}_o{k}_ids"])
                if i > 0:
                    assert np.array_equal(ctx.tracked_pairs(0, octave=k), g[f"f{i}_o{k}_tracked"])
            assert [r.error_code, int(r.valid), r.num_it, r.num_it_final] == g[f"f{i}_status"].tolist()
            if i > 0:
                np.testing.assert_allclose(r.delta_pose[:3], g[f"f{i}_delta"][:3], atol=POSE_TOL_RAD)
                np.testing.assert_allclose(r.delta_pose[3:], g[f"f{i}_delta"][3:], atol=POSE_TOL_M)
                assert np.array_equal(ctx.outliers(0), g[f"f{i}_outliers"])
                np.testing.assert_allclose(ctx.residuals(0), g[f"f{i}_residual"], rtol=1e-6, atol=1e-9)
    finally:
        ctx.close()
    f3, cam3 = synth.make_sequence("c3", seed=2, n_frames=1)
    ctx = capi.Context(width=1242, height=375, orb_nfeats=1000, max_seqs=1, max_corners=32768, n_octaves=2)
    try:
        ctx.process_batch([f3[0][0]], [f3[0][1]], [gcam(cam3)])
        assert np.array_equal(ctx.pyramid_plane(0, 1, 0), g["c3_pyrdown_L"]) and np.array_equal(ctx.pyramid_plane(0, 1, 1), g["c3_pyrdown_R"])
    finally:
        ctx.close()
