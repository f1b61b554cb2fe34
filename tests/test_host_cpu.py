"""CPU-side tests (-m "not gpu"): the C-ABI library loads and exports every symbol of include/svo_b200.h,
fails loudly without a device (no CPU fallback), and the host/device shared headers of the product
(introsort emulation, F-matrix RANSAC core) agree with libstdc++ / OpenCV when compiled for the host.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from stereo_vo_b200 import capi
    capi.build()
    return capi


def test_library_exports_every_header_symbol(built_lib):
    capi = built_lib
    hdr = open(os.path.join(ROOT, "include", "svo_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(svo_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 25
    L = C.CDLL(capi.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"{name} is declared in svo_b200.h but not exported"
    assert sorted(capi.EXPORTS) == declared
    assert L.svo_abi_version() == 3


def test_struct_layouts_match_header(built_lib):
    capi = built_lib
    assert capi.KEYPOINT.itemsize == 28 and capi.DMATCH.itemsize == 16     # cv::KeyPoint / cv::DMatch
    assert C.sizeof(capi.Camera) == 72
    assert C.sizeof(capi.Result) == 8 * (16 + 3 + 3 + 3 + 6) + 4 * 12
    cfg = capi.default_config()
    assert (cfg.orb_nfeats, cfg.fast_threshold, cfg.min_distance, cfg.orb_threshold) == (500, 20, 3, 40)
    assert (cfg.max_iters, cfg.initial_max_iters, cfg.max_incr_cost, cfg.bad_tracking_th) == (100, 10, 3, 5)
    assert cfg.kernel_param == 3.0 and cfg.residual_threshold == 10.0 and cfg.min_mod_out_vector == 1e-3
    assert (cfg.match_method, cfg.if_match_method, cfg.ifm_win_w, cfg.ifm_win_h, cfg.orb_max_distance) == (0, 0, 20, 20, 40.0)
    assert (cfg.n_octaves, cfg.pyr_mode) == (1, 0)


def test_octave_configuration_is_validated_before_any_device_work(built_lib):
    """Argument errors of svo_create (the reference throws via ASSERT_) do not need a device."""
    capi = built_lib
    for kw, text in ((dict(n_octaves=0), "n_octaves"), (dict(n_octaves=5), "n_octaves"), (dict(pyr_mode=3), "pyr_mode"),
                     (dict(width=320, height=240, n_octaves=3), "64 x 64"), (dict(width=1024, height=1024, orb_nfeats=1, n_octaves=4), "orb_nfeats")):
        with pytest.raises(capi.SvoError) as e:
            capi.Context(**kw)
        assert e.value.code == capi.SVO_ERR_ARG and text in str(e.value), kw


def test_no_cpu_fallback_without_device(built_lib):
    capi = built_lib
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.SvoError) as e:
        capi.Context()
    assert e.value.code == capi.SVO_ERR_CUDA and "no CPU fallback" in str(e.value)


def test_ransac_is_built_without_fma_contraction(built_lib):
    """The F-matrix RANSAC kernel must evaluate the same IEEE operations as the host build of fundmat_core.cuh that is
    pinned against cv2: ransac.cu is compiled with --fmad=false, everything else with --fmad=true."""
    mk = open(os.path.join(ROOT, "stereo_vo_b200", "csrc", "Makefile")).read()
    assert re.search(r"^ransac\.o: NVFLAGS := .*--fmad=false", mk, re.M)
    log = os.path.join(ROOT, "stereo_vo_b200", "csrc", "ransac.ptxas.log")
    assert os.path.exists(log)                    # written by the build; the object exists
    import shutil
    if shutil.which("cuobjdump"):
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", "fundmat_kernel", os.path.join(ROOT, "stereo_vo_b200", "csrc", "ransac.o")],
                              capture_output=True, text=True).stdout
        if "DMUL" in sass:                        # explicit products stay products (DFMA only inside division / sqrt / libm)
            assert sass.count("DMUL") > 200


def test_product_does_not_reference_oracle():
    """The shipped package must never import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "stereo_vo_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), f
                assert "svo_oracle" not in txt and "libsvo_oracle" not in txt, f


def test_introsort_emulation_equals_libstdcxx_std_sort(tmp_path):
    exe = tmp_path / "test_introsort"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "cpp", "test_introsort.cpp")])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "failures=0" in out.stdout


def _two_views(rng, n, outl, integer=True):
    X = np.stack([rng.uniform(-3, 3, n), rng.uniform(-2, 2, n), rng.uniform(3, 12, n)], 1)
    f, c = 500.0, np.array([320.0, 240.0])
    w = rng.normal(0, 0.02, 3)
    t = rng.normal(0, 0.1, 3)
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    R = np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K
    Y = X @ R.T + t
    p1 = f * X[:, :2] / X[:, 2:] + c
    p2 = f * Y[:, :2] / Y[:, 2:] + c + rng.normal(0, 0.3, (n, 2))
    k = int(outl * n)
    p2[:k] += rng.uniform(-50, 50, (k, 2))
    if integer:
        p1, p2 = np.rint(p1), np.rint(p2)
    return p1.astype(np.float32), p2.astype(np.float32)


@pytest.fixture(scope="module")
def fundmat_host(tmp_path_factory):
    so = tmp_path_factory.mktemp("fm") / "libfundmat_host.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", str(so),
                           os.path.join(ROOT, "tests", "cpp", "fundmat_host.cpp")])
    lib = C.CDLL(str(so))

    def run(p1, p2):
        n = len(p1)
        mask = np.zeros(max(n, 1), np.uint8)
        r = lib.fundmat_host(p1.ctypes.data_as(C.c_void_p), p2.ctypes.data_as(C.c_void_p), n,
                             mask.ctypes.data_as(C.c_void_p), None)
        return r, mask[:n]
    return run


def test_fundmat_core_equals_opencv_ransac(fundmat_host):
    """stereo_vo_b200/csrc/fundmat_core.cuh (host build) against cv2.findFundamentalMat(FM_RANSAC,1.0,0.99):
    RNG stream, subset rejection, 7-point solver, adaptive iteration count, inlier mask."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    bad = total = 0
    for n in (14, 15, 16, 20, 50, 100, 300, 700):
        for outl in (0.0, 0.1, 0.3, 0.6):
            for integer in (True, False):
                for _ in range(2):
                    p1, p2 = _two_views(rng, n, outl, integer)
                    F, m = cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, 1.0, 0.99)
                    m = np.zeros(n, np.uint8) if m is None else m.ravel()
                    r, mine = fundmat_host(p1, p2)
                    total += 1
                    bad += int(not np.array_equal(m, mine))
    assert bad == 0, f"{bad}/{total} masks differ"


def test_fundmat_core_golden_masks(fundmat_host, golden):
    for i in (1, 2):
        for side in "LR":
            r, m = fundmat_host(golden[f"c1_f{i}_p1{side}"], golden[f"c1_f{i}_p2{side}"])
            assert np.array_equal(m, golden[f"c1_f{i}_inl{side}"])


def test_fundmat_core_small_inputs(fundmat_host):
    rng = np.random.default_rng(1)
    p1, p2 = _two_views(rng, 7, 0.0)
    assert fundmat_host(p1, p2)[0] == 7 and fundmat_host(p1[:6], p2[:6])[0] == -1
    p1, p2 = _two_views(rng, 11, 0.2)
    assert fundmat_host(p1, p2)[0] == 7     # LMedS on 8..13 points keeps exactly its 7 sample points


def test_host_shim_builds_and_fails_loudly_without_gpu(built_lib, tmp_path):
    import struct
    import torch
    demo = os.path.join(ROOT, "stereo_vo_b200", "host", "svo_demo")
    assert os.path.exists(demo) and os.path.exists(os.path.join(ROOT, "stereo_vo_b200", "host", "libstereo_odometry_b200.so"))
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    seq = tmp_path / "seq.bin"
    with open(seq, "wb") as f:
        f.write(struct.pack("<3i5d", 320, 240, 1, 250.0, 250.0, 160.0, 120.0, 0.12))
        f.write(bytes(320 * 240 * 2))
    p = subprocess.run([demo, str(seq), str(tmp_path / "pose.txt"), "100"], capture_output=True, text=True)
    assert p.returncode == 1 and "no CUDA device" in p.stderr


def test_host_shim_ini_loader(built_lib, tmp_path):
    """loadParamsFromConfigFile / loadParamsFromConfigFileName of the host shim (libstereo-odometry.h:554-672)."""
    exe = tmp_path / "test_config"
    host = os.path.join(ROOT, "stereo_vo_b200", "host")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "cpp", "test_config.cpp"),
                           "-L" + host, "-lstereo_odometry_b200", "-L" + os.path.join(ROOT, "stereo_vo_b200"), "-lsvo_b200",
                           "-Wl,-rpath," + host, "-Wl,-rpath," + os.path.join(ROOT, "stereo_vo_b200")])
    p = subprocess.run([str(exe), os.path.join(host, "hot_path.ini")], capture_output=True, text=True)
    assert p.returncode == 0 and "failures=0" in p.stdout, p.stdout + p.stderr
