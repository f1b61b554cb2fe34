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
    assert L.svo_abi_version() == 4


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


@pytest.fixture(scope="module")
def fundmat_lib(tmp_path_factory):
    so = tmp_path_factory.mktemp("fm7") / "libfundmat_host7.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-mfma", "-o", str(so),
                           os.path.join(ROOT, "tests", "cpp", "fundmat_host.cpp")])
    lib = C.CDLL(str(so))
    for f in (lib.cr_pow13, lib.cr_cos, lib.cr_acos):
        f.restype = C.c_double
        f.argtypes = [C.c_double]
    return lib


def test_fm7_svd_is_opencv_jacobi_svd_bit_for_bit(fundmat_lib):
    """fm7_svd_rows (csrc/fundmat_core.cuh) = the SVDecomp(A, W, U, Vt, MODIFY_A + FULL_UV) call of run7Point: all nine
    right singular vectors, including the two null-space rows built from OpenCV's fixed RNG stream, equal cv2's bits."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for t in range(300):
        A = rng.normal(size=(7, 9))
        if t % 3 == 0:
            A[:, 8] = 1.0
        if t % 50 == 7:
            A[3] = A[2]                                  # rank deficient: a zero singular value among the first seven
        _, _, vt = cv2.SVDecomp(A.copy(), flags=cv2.SVD_FULL_UV)
        At = np.zeros((9, 9))
        At[:7] = A
        fundmat_lib.fundmat_host_svd_rows(At.ctypes.data_as(C.c_void_p))
        assert np.array_equal(At, vt), t


def test_solve_cubic_is_opencv_solve_cubic_bit_for_bit(fundmat_lib):
    """solve_cubic = cv::solveCubic of OpenCV 4.13 (coefficient rescaling, expanded discriminant), incl. degenerate degrees."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for t in range(6000):
        c = rng.normal(size=4) * 10.0 ** rng.integers(-3, 4)
        if t % 7 == 0:
            c[0] = 0.0
        if t % 49 == 0:
            c[1] = 0.0
        if t % 11 == 0:
            c[0] *= 1e-18
        if t % 13 == 0:                                   # double root: d == 0 exactly for small integers
            a, b = rng.integers(-5, 6, 2)
            c = np.array([1.0, -(2 * a + b), a * a + 2 * a * b, -a * a * b], np.float64)
        n, roots = cv2.solveCubic(c)
        r = np.zeros(3)
        k = fundmat_lib.fundmat_host_cubic(c.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p))
        assert k == n, (t, c, n, k)
        assert np.array_equal(roots.ravel()[:max(n, 0)], r[:max(n, 0)]), (t, c)


def test_run7point_is_opencv_run7point_bit_for_bit(fundmat_lib):
    """cv2.findFundamentalMat(7 points, FM_7POINT) returns run7Point's stacked models: same count, same bits."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for t in range(3000):
        p1 = rng.uniform(0, 640, (7, 2)).astype(np.float32)
        p2 = (p1 + rng.normal(0, 6, (7, 2))).astype(np.float32)
        if t % 2 == 0:
            p1, p2 = np.round(p1), np.round(p2)
        if t % 10 == 0:                                   # sideways motion of coplanar points: the degenerate samples
            p2 = (p1 + np.array([float(rng.integers(1, 9)), 0.0], np.float32)).astype(np.float32)
        try:
            F, _ = cv2.findFundamentalMat(p1, p2, cv2.FM_7POINT)
        except cv2.error:
            continue                                     # cv2's own rowRange assertion (DESIGN.md limit (d))
        Fm = np.zeros(27)
        k = fundmat_lib.fundmat_host_7point(p1.ctypes.data_as(C.c_void_p), p2.ctypes.data_as(C.c_void_p), Fm.ctypes.data_as(C.c_void_p))
        if F is None:
            assert k <= 0, t
            continue
        assert F.shape[0] == 3 * k and np.array_equal(F.ravel(), Fm[:9 * k]), t


def test_cr_math_is_correctly_rounded_and_agrees_with_glibc(fundmat_lib):
    """csrc/cr_math.cuh (what the device evaluates instead of libdevice's 1-2 ulp pow / acos / cos): equal to the
    correctly rounded value (mpmath, 250 bits) on every sample, and to glibc on all but its own rare misroundings."""
    mpmath = pytest.importorskip("mpmath")
    import math
    mpmath.mp.prec = 250
    rng = np.random.default_rng(1)
    y = mpmath.mpf(1 / 3)
    n, differ = 6000, 0
    for i in range(n):
        x = float(rng.uniform(1e-6, 3)) if i % 2 else float(10 ** rng.uniform(-12, 6))
        m = fundmat_lib.cr_pow13(x)
        assert m == float(mpmath.power(mpmath.mpf(x), y)), x
        differ += m != math.pow(x, 1 / 3)
        a = float(rng.uniform(0, 5.3)) if i % 3 else float(rng.uniform(-100, 100))
        m = fundmat_lib.cr_cos(a)
        assert m == float(mpmath.cos(mpmath.mpf(a))), a
        differ += m != math.cos(a)
        v = float(rng.uniform(-1, 1)) if i % 3 else float((1 - 10 ** rng.uniform(-12, 0)) * rng.choice([-1, 1]))
        m = fundmat_lib.cr_acos(v)
        assert m == float(mpmath.acos(mpmath.mpf(v))), v
        differ += m != math.acos(v)
    assert differ <= 3 * n * 0.004, differ               # glibc: < 1 ulp, correctly rounded in ~99.9 % of the calls


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
