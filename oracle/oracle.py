"""ctypes wrapper + pipeline driver for the CPU oracle (TEST INFRASTRUCTURE -- not product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It restates `rso::CStereoOdometryEstimator::processNewImagePair`
(libstereo-odometry/src/process_new_image_pair.cpp:41-385) in its FAST+ORB / smDescBF /
ifmDescBF configuration on the CPU:

  backend="cv2"   third-party calls go to the real OpenCV in this image (cv2 4.13.0), exactly the
                  calls the reference makes (stage2_detect.cpp:510-512, stage3_match_left_right.cpp:88-94,
                  stage4_match_consecutive.cpp:137-142,202,237); the reference's own glue and the
                  Gauss-Newton optimiser run in the C++ restatement oracle/svo_oracle.cpp.
  backend="port"  everything (FAST, blur+BRIEF, Hamming matcher too) runs in oracle/svo_oracle.cpp.

`ransac` selects how the two cv::findFundamentalMat masks of stage4_match_consecutive.cpp:202,237
are obtained: "cv2" (real OpenCV), "off" (masks absent => filter skipped, as the reference does when
fewer than 8 inliers are found, :243-259) or a callable (p1, p2) -> mask.

Parity status: PARITY UNPINNED BY THE REFERENCE (it ships no vectors for this path and cannot be built in this
image); pinned against the real OpenCV (cv2) goldens, libstdc++'s std::sort and SURVEY.md A.7's known answers instead.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsvo_oracle.so")
_SRC = os.path.join(_HERE, "svo_oracle.cpp")

KEYPOINT = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
DMATCH = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])

VOEC = {"voecNone": 0, "voecBadCondNumber": 1, "voecIncrFuncCostStg1": 2, "voecIncrFuncCostStg2": 3,
        "voecFirstIteration": 4, "voecBadTracking": 5}


class Camera(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("fxl", "fyl", "cxl", "cyl", "fxr", "fyr", "cxr", "cyr", "baseline")]


class LSParams(C.Structure):
    """TLeastSquaresParams (libstereo-odometry.h:313-350), defaults of common.cpp:69-82."""
    _fields_ = [("use_robust_kernel", C.c_int32), ("kernel_param", C.c_double), ("max_iters", C.c_int32),
                ("initial_max_iters", C.c_int32), ("min_mod_out_vector", C.c_double),
                ("max_incr_cost", C.c_int32), ("residual_threshold", C.c_double)]

    @staticmethod
    def defaults():
        return LSParams(1, 3.0, 100, 10, 1e-3, 3, 10.0)


class GNResult(C.Structure):
    _fields_ = [("valid", C.c_int32), ("error_code", C.c_int32), ("num_it", C.c_int32), ("num_it_final", C.c_int32),
                ("delta_pose", C.c_double * 6), ("out_pose_rotvec", C.c_double * 3), ("out_pose_t", C.c_double * 3),
                ("out_pose_mat", C.c_double * 16), ("out_pose_ypr", C.c_double * 3),
                ("n_outliers_field", C.c_int32), ("n_residuals", C.c_int32), ("reached_end", C.c_int32),
                ("bad_cond", C.c_int32), ("last_cost", C.c_double)]


def build(force=False):
    """Compile oracle/svo_oracle.cpp with g++ (libstdc++'s std::sort is part of the oracle)."""
    if (not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC)
            and os.path.getmtime(_SO) >= os.path.getmtime(os.path.join(_HERE, "orb_pattern.inc"))):
        return _SO
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", _SO, _SRC]
    subprocess.check_call(cmd, cwd=_HERE)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_fast9.restype = C.c_int
        _lib.orc_orb_compute.restype = C.c_int
        _lib.orc_non_max_sup.restype = C.c_int
        _lib.orc_bf_match.restype = C.c_int
        _lib.orc_lr_filter.restype = C.c_int
        _lib.orc_track_filter.restype = C.c_int
        _lib.orc_projected_coords.restype = C.c_int
        _lib.orc_match_rbr.restype = C.c_int
        _lib.orc_track_win.restype = C.c_int
        _lib.orc_track_consistency.restype = C.c_int
        _lib.orc_eval_rgn.restype = C.c_int
        assert _lib.orc_sizeof_keypoint() == KEYPOINT.itemsize
        assert _lib.orc_sizeof_dmatch() == DMATCH.itemsize
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _img(a):
    a = np.ascontiguousarray(a, np.uint8)
    assert a.ndim == 2
    return a


# ------------------------------------------------------------------------------------------------
# thin wrappers (one per oracle entry point)
# ------------------------------------------------------------------------------------------------
def fast9(img, th):
    img = _img(img)
    h, w = img.shape
    cap = max(16, (w * h) // 4 + 16)
    out = np.zeros(cap, KEYPOINT)
    n = lib().orc_fast9(_p(img), w, h, w, int(th), _p(out), cap)
    return out[:n].copy()


def blur7(img):
    img = _img(img)
    h, w = img.shape
    out = np.empty_like(img)
    lib().orc_blur7(_p(img), w, h, w, _p(out), w)
    return out


def orb_compute(img, kps):
    img = _img(img)
    h, w = img.shape
    k = np.ascontiguousarray(kps, KEYPOINT).copy()
    desc = np.zeros((len(k), 32), np.uint8)
    n = lib().orc_orb_compute(_p(img), w, h, w, _p(k), len(k), _p(desc))
    return k[:n].copy(), desc[:n].copy()


def non_max_sup(num_out, kps, desc, min_distance, H, W, sort_mode=0):
    kps = np.ascontiguousarray(kps, KEYPOINT)
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    n = len(kps)
    ok = np.zeros(max(n, 1), KEYPOINT)
    od = np.zeros((max(n, 1), 32), np.uint8)
    sel = np.zeros(max(n, 1), np.int32)
    m = lib().orc_non_max_sup(int(num_out), _p(kps), n, _p(desc), int(min_distance), int(H), int(W),
                              _p(ok), _p(od), _p(sel), int(sort_mode))
    return ok[:m].copy(), od[:m].copy(), sel[:m].copy()


def update_indexes(kps, desc, H, sort_mode=0):
    k = np.ascontiguousarray(kps, KEYPOINT).copy()
    d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32).copy()
    perm = np.zeros(max(len(k), 1), np.int32)
    row_index = np.zeros(H, np.int64)
    lib().orc_update_indexes(_p(k), _p(d), len(k), int(H), _p(perm), _p(row_index), int(sort_mode))
    return k, d, perm[:len(k)].copy(), row_index


def bf_match(q, t):
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
    t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    out = np.zeros(max(len(q), 1), DMATCH)
    n = lib().orc_bf_match(_p(q), len(q), _p(t), len(t), _p(out))
    return out[:n].copy()


def lr_filter(matches, kpsL, kpsR, robust_1to1, max_y_diff, orb_th, width):
    m = np.ascontiguousarray(matches, DMATCH).copy()
    kpsL = np.ascontiguousarray(kpsL, KEYPOINT)
    kpsR = np.ascontiguousarray(kpsR, KEYPOINT)
    n = lib().orc_lr_filter(_p(m), len(m), _p(kpsL), _p(kpsR), len(kpsR), int(bool(robust_1to1)),
                            C.c_double(max_y_diff), int(orb_th), int(width))
    return m[:n].copy()


def lr_row_index(matches, kpsL, H):
    m = np.ascontiguousarray(matches, DMATCH)
    kpsL = np.ascontiguousarray(kpsL, KEYPOINT)
    ri = np.zeros(H + 1, np.int64)
    lib().orc_lr_row_index(_p(m), len(m), _p(kpsL), len(kpsL), int(H), _p(ri))
    return ri


def track_filter(matL, matR, cur_n, orb_th):
    l = np.ascontiguousarray(matL, DMATCH).copy()
    r = np.ascontiguousarray(matR, DMATCH).copy()
    assert len(l) == len(r)
    n = lib().orc_track_filter(_p(l), _p(r), len(l), int(cur_n), int(orb_th))
    return l[:n].copy(), r[:n].copy()


def match_rbr(kpsL, descL, idxL, kpsR, descR, idxR, H, W, max_y_diff, orb_max_distance, robust_1to1):
    """stage 3, smDescRbR branch (stage3_match_left_right.cpp:185-419)."""
    kl = np.ascontiguousarray(kpsL, KEYPOINT)
    kr = np.ascontiguousarray(kpsR, KEYPOINT)
    dl = np.ascontiguousarray(descL, np.uint8).reshape(-1, 32)
    dr = np.ascontiguousarray(descR, np.uint8).reshape(-1, 32)
    il = np.ascontiguousarray(idxL, np.int64)
    ir = np.ascontiguousarray(idxR, np.int64)
    assert len(il) == H and len(ir) == H
    out = np.zeros(max(len(kl), 1), DMATCH)
    n = lib().orc_match_rbr(_p(kl), _p(dl), len(kl), _p(il), _p(kr), _p(dr), len(kr), _p(ir), int(H), int(W),
                            C.c_double(max_y_diff), C.c_double(orb_max_distance), int(bool(robust_1to1)), _p(out))
    return out[:n].copy()


def track_win(prev, cur, win_w, win_h):
    """stage 4, ifmDescWin branch (stage4_match_consecutive.cpp:435-640) up to the candidate list: (n, 2) array of
    (previous match index, current match index) in current-match order."""
    pm = np.ascontiguousarray(prev.matches, DMATCH)
    cm = np.ascontiguousarray(cur.matches, DMATCH)
    pri = lr_row_index(pm, prev.kpsL, prev.H)
    cri = lr_row_index(cm, cur.kpsL, cur.H)
    op = np.zeros(max(len(cm), 1), np.int32)
    oc = np.zeros(max(len(cm), 1), np.int32)
    a = [np.ascontiguousarray(x, KEYPOINT) for x in (prev.kpsL, prev.kpsR, cur.kpsL, cur.kpsR)]
    pd = np.ascontiguousarray(prev.descL, np.uint8).reshape(-1, 32)
    cd = np.ascontiguousarray(cur.descL, np.uint8).reshape(-1, 32)
    n = lib().orc_track_win(_p(pm), len(pm), _p(a[0]), _p(a[1]), _p(pd), _p(pri), _p(cm), len(cm), _p(a[2]), _p(a[3]),
                            _p(cd), _p(cri), int(cur.H), int(cur.W), int(win_w), int(win_h), _p(op), _p(oc))
    return np.stack([op[:n], oc[:n]], 1).astype(np.int64)


def projected_coords(pre_matches, pre_left, pre_right, other_matches_tracked, cam, delta_pose):
    """getProjectedCoords (common.cpp:415-470) for delta_pose = [w t] of the inverse of change_pose."""
    pm = np.ascontiguousarray(pre_matches, DMATCH)
    kl = np.ascontiguousarray(pre_left, KEYPOINT)
    kr = np.ascontiguousarray(pre_right, KEYPOINT)
    fl = np.ascontiguousarray(other_matches_tracked, np.int32)
    dp = np.ascontiguousarray(delta_pose, np.float64)
    out = np.zeros((max(len(pm), 1), 4), np.float32)
    n = lib().orc_projected_coords(_p(pm), len(pm), _p(kl), _p(kr), _p(fl), C.byref(cam), _p(dp), _p(out))
    return out[:n].copy()


def aux_nms_mask(xy, response, H, W, min_distance, num_out=None, sort_mode=0):
    """auxiliary m_non_max_sup (stage2_detect.cpp:225-283) alone: survivors mask"""
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    resp = np.ascontiguousarray(response, np.float32)
    out = np.zeros(max(len(xy), 1), np.uint8)
    lib().orc_aux_non_max_sup(_p(xy), _p(resp), len(xy), int(H), int(W), int(len(xy) if num_out is None else num_out),
                              int(min_distance), int(sort_mode), _p(out))
    return out[:len(xy)].astype(bool)


def pinhole_projection(lmks, cam, delta_pose):
    """m_pinhole_stereo_projection (stage5_optimization.cpp:35-257) alone: (n x 4 float32 pixels, n x 4 x 6 Jacobians)"""
    lm = np.ascontiguousarray(lmks, np.float64).reshape(-1, 3)
    dp = np.ascontiguousarray(delta_pose, np.float64)
    px = np.zeros((len(lm), 4), np.float32)
    J = np.zeros((len(lm), 4, 6))
    lib().orc_pinhole_projection(_p(lm), len(lm), C.byref(cam), _p(dp), _p(px), _p(J))
    return px, J


def track_consistency(matL, matR, inlL, inlR):
    l = np.ascontiguousarray(matL, DMATCH)
    r = np.ascontiguousarray(matR, DMATCH)
    il = None if inlL is None else np.ascontiguousarray(inlL, np.uint8)
    ir = None if inlR is None else np.ascontiguousarray(inlR, np.uint8)
    tp = np.zeros((max(len(l), 1), 2), np.int64)
    n = lib().orc_track_consistency(_p(l), _p(r), len(l), _p(il), _p(ir), _p(tp))
    return tp[:n].copy()


def eval_rgn(l1l, l1r, l2l, l2r, cam, ls, delta_pose):
    a = [np.ascontiguousarray(v, np.float32).reshape(-1, 2) for v in (l1l, l1r, l2l, l2r)]
    n = len(a[0])
    dp = np.ascontiguousarray(delta_pose, np.float64)
    new = np.zeros(6)
    grad = np.zeros(6)
    res = np.zeros(n)
    cost = C.c_double(0)
    rc = lib().orc_eval_rgn(_p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), n, C.byref(cam), C.byref(ls), _p(dp),
                            _p(new), _p(grad), _p(res), C.byref(cost))
    return rc, new, grad, res, cost.value


def stage5(l1l, l1r, l2l, l2r, resp1l, cur_match_idx, H, W, min_distance, cam, ls, initial_pose, sort_mode=0,
           survivors=None):
    a = [np.ascontiguousarray(v, np.float32).reshape(-1, 2) for v in (l1l, l1r, l2l, l2r)]
    n = len(a[0])
    resp = np.ascontiguousarray(resp1l, np.float32)
    cmi = np.ascontiguousarray(cur_match_idx, np.int64)
    ip = np.ascontiguousarray(initial_pose, np.float64)
    res = GNResult()
    out_res = np.full(max(n, 1), np.finfo(np.float64).max)
    inl = np.zeros(max(n, 1), np.int64)
    sv = None if survivors is None else np.ascontiguousarray(survivors, np.uint8)
    lib().orc_stage5(_p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(resp), _p(cmi), n, int(H), int(W), int(min_distance),
                     C.byref(cam), C.byref(ls), _p(ip), int(sort_mode), _p(sv), C.byref(res), _p(out_res), _p(inl))
    return res, out_res[:res.n_residuals].copy(), inl[:res.n_outliers_field].copy()


# ------------------------------------------------------------------------------------------------
# cv2 back-end for the third-party calls (the calls the reference makes)
# ------------------------------------------------------------------------------------------------
def _cv2():
    import cv2
    return cv2


def cv2_kps_to_array(kps):
    out = np.zeros(len(kps), KEYPOINT)
    for i, k in enumerate(kps):
        out[i] = (k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave, k.class_id)
    return out


def cv2_fast_orb(img, th):
    """stage2_detect.cpp:510-512 with the real OpenCV."""
    cv2 = _cv2()
    kps = cv2.FastFeatureDetector_create(int(th)).detect(img)
    kps, desc = cv2.ORB_create().compute(img, kps)
    if desc is None:
        desc = np.zeros((0, 32), np.uint8)
    return cv2_kps_to_array(kps), np.ascontiguousarray(desc)


def cv2_orb_detect_and_compute(img, n_feats_to_extract, nlevels, fast_th):
    """stage2_detect.cpp:484-496 (dmORB, OpenCV >= 3): oriented multi-scale ORB in one call.  There is no C++ restatement
    of this branch yet: the real OpenCV is the only back-end."""
    cv2 = _cv2()
    orb = cv2.ORB_create(int(n_feats_to_extract), 1.2, int(nlevels), 31, 0, 2, cv2.ORB_HARRIS_SCORE, 31, int(fast_th))
    kps, desc = orb.detectAndCompute(img, None)
    if desc is None:
        desc = np.zeros((0, 32), np.uint8)
    return cv2_kps_to_array(kps), np.ascontiguousarray(desc)


def cv2_bf_match(q, t):
    """stage3_match_left_right.cpp:88-94 with the real OpenCV."""
    cv2 = _cv2()
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
    t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    if len(q) == 0 or len(t) == 0:
        return np.zeros(0, DMATCH)
    ms = cv2.BFMatcher(cv2.NORM_HAMMING, False).match(q, t)
    out = np.zeros(len(ms), DMATCH)
    for i, m in enumerate(ms):
        out[i] = (m.queryIdx, m.trainIdx, m.imgIdx, m.distance)
    return out


def cv2_fmat_mask(p1, p2):
    """stage4_match_consecutive.cpp:202,237: findFundamentalMat(p1,p2,FM_RANSAC,1.0,0.99,mask).
    Returns None when OpenCV returns no mask (fewer than 7 points)."""
    cv2 = _cv2()
    p1 = np.ascontiguousarray(p1, np.float32).reshape(-1, 2)
    p2 = np.ascontiguousarray(p2, np.float32).reshape(-1, 2)
    if len(p1) < 7:
        return None
    F, mask = cv2.findFundamentalMat(p1, p2, cv2.FM_RANSAC, 1.0, 0.99)
    if mask is None:
        return None
    return np.ascontiguousarray(mask.ravel().astype(np.uint8))


_fm_host = None


def host_fmat_mask(p1, p2):
    """The host build of stereo_vo_b200/csrc/fundmat_core.cuh (tests/cpp/fundmat_host.cpp): bit-identical to
    cv2.findFundamentalMat on every input cv2 answers (tests/test_host_cpu.py, tools/ransac_host_vs_cv2.py).  Used ONLY
    where cv2 4.13 itself raises (its solveCubic returns -1 for a sample whose cubic vanishes and runKernel then asks
    for a negative row range, matrix.cpp:764): this continues the way OpenCV's loop does for any other sample without a
    model (nmodels <= 0 -> next iteration)."""
    global _fm_host
    if _fm_host is None:
        so = os.path.join(_HERE, "libfundmat_host.so")
        src = os.path.join(os.path.dirname(_HERE), "tests", "cpp", "fundmat_host.cpp")
        hdr = os.path.join(os.path.dirname(_HERE), "stereo_vo_b200", "csrc", "fundmat_core.cuh")
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-o", so, src])
        _fm_host = C.CDLL(so)
    p1 = np.ascontiguousarray(p1, np.float32).reshape(-1, 2)
    p2 = np.ascontiguousarray(p2, np.float32).reshape(-1, 2)
    n = len(p1)
    if n < 7:
        return None
    mask = np.zeros(n, np.uint8)
    _fm_host.fundmat_host(_p(p1), _p(p2), n, _p(mask), None)
    return mask


def cv2_fmat_mask_tolerant(p1, p2):
    """cv2.findFundamentalMat; where cv2 4.13 raises its internal assertion, the host build of the same arithmetic."""
    try:
        return cv2_fmat_mask(p1, p2)
    except _cv2().error:
        cv2_fmat_mask_tolerant.raised += 1
        return host_fmat_mask(p1, p2)


cv2_fmat_mask_tolerant.raised = 0


# ------------------------------------------------------------------------------------------------
# Parameters (same names as the reference's param structs)
# ------------------------------------------------------------------------------------------------
class Params:
    """The subset of the reference's parameters that the hot path reads; names as in
    libstereo-odometry.h:266-508, defaults as in SURVEY.md Appendix C (the INI that selects the path)."""

    def __init__(self, **kw):
        # TDetectParams (H:383-447)
        self.detect_method = 1                # TDMethod: 0 = dmORB (cv2 back-end only), 1 = dmFAST_ORB (H:389)
        self.orb_nlevels = 8
        self.orb_nfeats = 500
        self.initial_FAST_threshold = 20
        self.non_maximal_suppression = True
        self.min_distance = 3
        # TRectifyParams
        self.nOctaves = 1
        # TLeftRightMatchParams (H:449-499)
        self.max_y_diff = 1.0
        self.orb_max_distance = 40.0
        self.enable_robust_1to1_match = False
        self.match_method = 0                 # TSMMethod: 0 = smDescBF, 1 = smDescRbR (H:454)
        # TInterFrameMatchingParams (H:285-310); the window sizes have no default in the reference (empty ctor, C:84)
        self.ifm_method = 0                   # TIFMMethod: 0 = ifmDescBF, 1 = ifmDescWin (H:291)
        self.ifm_win_w = 20
        self.ifm_win_h = 20
        # TLeastSquaresParams (H:313-350)
        self.use_robust_kernel = True
        self.kernel_param = 3.0
        self.max_iters = 100
        self.initial_max_iters = 10
        self.min_mod_out_vector = 1e-3
        self.max_incr_cost = 3
        self.residual_threshold = 10.0
        self.bad_tracking_th = 5
        self.use_previous_pose_as_initial = True
        self.use_custom_initial_pose = False
        # TGeneralParams
        self.vo_use_matches_ids = False
        for k, v in kw.items():
            if not hasattr(self, k):
                raise AttributeError(k)
            setattr(self, k, v)

    def ls(self):
        return LSParams(int(self.use_robust_kernel), float(self.kernel_param), int(self.max_iters),
                        int(self.initial_max_iters), float(self.min_mod_out_vector), int(self.max_incr_cost),
                        float(self.residual_threshold))


def make_camera(fx, fy, cx, cy, baseline):
    return Camera(fx, fy, cx, cy, fx, fy, cx, cy, baseline)


class FrameData:
    """TImagePairData (libstereo-odometry.h:756-811) for nOctaves == 1."""
    __slots__ = ("kpsL", "descL", "kpsR", "descR", "matches", "ids", "H", "W", "idxL", "idxR")


class OracleEstimator:
    """CPU restatement of CStereoOdometryEstimator on the FAST+ORB / DescBF / DescBF path."""

    def __init__(self, params=None, backend="cv2", ransac="cv2", sort_mode=0):
        self.p = params or Params()
        self.backend = backend
        self.ransac = ransac
        self.sort_mode = sort_mode
        # ctor state, common.cpp:28-50; thresholds after an INI load: H:661-662
        self.m_current_fast_th = int(self.p.initial_FAST_threshold)
        self.m_current_orb_th = int(self.p.orb_max_distance)
        self.m_error = VOEC["voecNone"]
        self.m_last_computed_pose = np.zeros(6)
        self.m_last_match_ID = 0
        self.m_last_kf_max_id = 0
        self.m_reset = False
        self.m_it_counter = 0
        self.prev = None
        self.cur = None
        self.trace = {}

    # -- match-ID / key-frame control (libstereo-odometry.h:675-726) --------------------------------
    def set_this_frame_as_kf(self):                  # H:675-683
        assert self.cur is not None and self.cur.ids is not None and len(self.cur.ids) > 0
        self.m_last_kf_max_id = int(self.cur.ids.max())

    def reset_ids(self):                             # H:684
        self.m_reset = True

    def set_ids(self, ids):                          # H:685-694
        f = self.cur if self.cur is not None else self.prev
        f.ids = np.array(ids, np.int64)

    def set_max_match_id(self, id_):                 # H:726
        self.m_last_match_ID = int(id_)

    # -- stage 2 (stage2_detect.cpp:385-671, FAST_ORB branch) --------------------------------------
    def detect(self, img):
        H, W = img.shape
        th = self.m_current_fast_th
        if self.p.detect_method == 0:              # dmORB, S2:458-496: more features when the NMS thins them out afterwards
            assert self.backend == "cv2", "dmORB has no C++ restatement: backend must be cv2"
            n_extract = int(1.5 * self.p.orb_nfeats) if self.p.non_maximal_suppression else int(self.p.orb_nfeats)
            kps, desc = cv2_orb_detect_and_compute(img, n_extract, self.p.orb_nlevels, th)
        elif self.backend == "cv2":
            kps, desc = cv2_fast_orb(img, th)
        else:
            kps = fast9(img, th)
            kps, desc = orb_compute(img, kps)
        raw = (kps.copy(), desc.copy())
        nOct = 1 if self.p.detect_method == 0 else int(self.p.nOctaves)      # S1:80: ORB builds its own pyramid
        quota = int(self.p.orb_nfeats * (2 * nOct) / (2 ** nOct - 1))  # S2:405
        if self.p.non_maximal_suppression:
            kps, desc, sel = non_max_sup(quota, kps, desc, self.p.min_distance, H, W, self.sort_mode)
        kps, desc, perm, row_index = update_indexes(kps, desc, H, self.sort_mode)
        self._last_row_index = row_index           # pyr_feats_index (only the row-by-row matcher reads it)
        return kps, desc, raw

    # -- stage 3 (stage3_match_left_right.cpp:83-178) ------------------------------------------------
    def match_lr(self, f):
        if self.p.match_method == 1:               # smDescRbR, S3:185-419 (no BF post-filter on this branch)
            m = match_rbr(f.kpsL, f.descL, f.idxL, f.kpsR, f.descR, f.idxR, f.H, f.W, self.p.max_y_diff,
                          self.p.orb_max_distance, self.p.enable_robust_1to1_match)
            return m, m.copy()
        m = (cv2_bf_match if self.backend == "cv2" else bf_match)(f.descL, f.descR)
        raw = m.copy()
        m = lr_filter(m, f.kpsL, f.kpsR, self.p.enable_robust_1to1_match, self.p.max_y_diff,
                      self.m_current_orb_th, f.W)
        return m, raw

    # -- stage 4 (stage4_match_consecutive.cpp:88-329) -----------------------------------------------
    def track(self, prev, cur):
        if self.p.ifm_method == 1:                 # ifmDescWin, S4:435-738
            return self.track_window(prev, cur)
        bf = cv2_bf_match if self.backend == "cv2" else bf_match
        preL = prev.descL[prev.matches["queryIdx"]]
        preR = prev.descR[prev.matches["trainIdx"]]
        curL = cur.descL[cur.matches["queryIdx"]]
        curR = cur.descR[cur.matches["trainIdx"]]
        matL = bf(preL, curL)
        matR = bf(preR, curR)
        raw = (matL.copy(), matR.copy())
        if len(matL) != len(matR):  # only when one of the train sets is empty
            matL = matL[:0]
            matR = matR[:0]
        matL, matR = track_filter(matL, matR, len(cur.matches), self.m_current_orb_th)
        # point sets for the two findFundamentalMat calls, S4:181-232
        pl = prev.matches[matL["queryIdx"]]
        cl = cur.matches[matL["trainIdx"]]
        p1L = np.stack([prev.kpsL["x"][pl["queryIdx"]], prev.kpsL["y"][pl["queryIdx"]]], 1).astype(np.float32)
        p2L = np.stack([cur.kpsL["x"][cl["queryIdx"]], cur.kpsL["y"][cl["queryIdx"]]], 1).astype(np.float32)
        pr = prev.matches[matR["queryIdx"]]
        cr = cur.matches[matR["trainIdx"]]
        p1R = np.stack([prev.kpsR["x"][pr["trainIdx"]], prev.kpsR["y"][pr["trainIdx"]]], 1).astype(np.float32)
        p2R = np.stack([cur.kpsR["x"][cr["trainIdx"]], cur.kpsR["y"][cr["trainIdx"]]], 1).astype(np.float32)
        if self.ransac == "off":
            inlL = inlR = None
        else:
            fn = cv2_fmat_mask if self.ransac == "cv2" else (cv2_fmat_mask_tolerant if self.ransac == "cv2_tolerant" else self.ransac)
            inlL = fn(p1L, p2L)
            inlR = fn(p1R, p2R)
        tracked = track_consistency(matL, matR, inlL, inlR)
        return tracked, dict(matL_raw=raw[0], matR_raw=raw[1], matL=matL, matR=matR, inlL=inlL, inlR=inlR,
                             p1L=p1L, p2L=p2L, p1R=p1R, p2R=p2R)

    def track_window(self, prev, cur):
        cand = track_win(prev, cur, self.p.ifm_win_w, self.p.ifm_win_h)
        pm = prev.matches[cand[:, 0]]
        cm = cur.matches[cand[:, 1]]
        def xy(k, idx):
            return np.stack([k["x"][idx], k["y"][idx]], 1).astype(np.float32)
        p1L, p1R = xy(prev.kpsL, pm["queryIdx"]), xy(prev.kpsR, pm["trainIdx"])
        p2L, p2R = xy(cur.kpsL, cm["queryIdx"]), xy(cur.kpsR, cm["trainIdx"])
        keep = np.ones(len(cand), bool)
        inlL = inlR = None
        if self.ransac != "off" and len(cand):
            fn = cv2_fmat_mask if self.ransac == "cv2" else (cv2_fmat_mask_tolerant if self.ransac == "cv2_tolerant" else self.ransac)
            inlL = fn(p1L, p2L)                    # S4:693-713: the right filter only runs when the left one found >= 8
            if inlL is not None and int(np.count_nonzero(inlL)) >= 8:
                inlR = fn(p1R, p2R)
                if inlR is not None and int(np.count_nonzero(inlR)) >= 8:
                    keep = (np.asarray(inlL).ravel() != 0) & (np.asarray(inlR).ravel() != 0)
        tracked = cand[keep]
        return tracked, dict(cand=cand, inlL=inlL, inlR=inlR, p1L=p1L, p2L=p2L, p1R=p1R, p2R=p2R)

    # -- stage 5 (stage5_optimization.cpp:392-736) via the tracked pairs ------------------------------
    def optimize(self, prev, cur, tracked, cam, initial_estimation=None):
        pm = prev.matches[tracked[:, 0]]
        cm = cur.matches[tracked[:, 1]]
        def xy(k, idx):
            return np.stack([k["x"][idx], k["y"][idx]], 1).astype(np.float32)
        l1l = xy(prev.kpsL, pm["queryIdx"]); l1r = xy(prev.kpsR, pm["trainIdx"])
        l2l = xy(cur.kpsL, cm["queryIdx"]); l2r = xy(cur.kpsR, cm["trainIdx"])
        resp = prev.kpsL["response"][pm["queryIdx"]].astype(np.float32)
        if self.p.use_custom_initial_pose:
            ini = np.asarray(initial_estimation if initial_estimation is not None else np.zeros(6), np.float64)
        elif self.p.use_previous_pose_as_initial:
            ini = self.m_last_computed_pose.copy()
        else:
            ini = np.zeros(6)
        res, residual, inl = stage5(l1l, l1r, l2l, l2r, resp, tracked[:, 1], prev.H, prev.W, self.p.min_distance,
                                    cam, self.p.ls(), ini, self.sort_mode)
        if res.bad_cond:
            self.m_error = VOEC["voecBadCondNumber"]
        if res.reached_end and (not self.p.use_custom_initial_pose) and self.p.use_previous_pose_as_initial:
            self.m_last_computed_pose = np.array(res.delta_pose[:])  # S5:720-721
        return res, residual, inl, dict(l1l=l1l, l1r=l1r, l2l=l2l, l2r=l2r, resp=resp, ini=ini)

    # -- processNewImagePair (process_new_image_pair.cpp:41-385) --------------------------------------
    def process_new_image_pair(self, imgL, imgR, cam, repeat=False):
        imgL = _img(imgL)
        imgR = _img(imgR)
        out = dict(error_code=VOEC["voecNone"], valid=False, num_it=0, num_it_final=0, outliers=np.zeros(0, np.int64),
                   out_residual=np.zeros(0), tracked=np.zeros((0, 2), np.int64), pose=None)
        if (not repeat) and self.m_error not in (VOEC["voecBadTracking"], VOEC["voecBadCondNumber"]):
            self.prev = self.cur  # P:86-89
        self.m_error = VOEC["voecNone"]
        f = FrameData()
        f.H, f.W = imgL.shape
        f.kpsL, f.descL, rawL = self.detect(imgL)   # P:166
        f.idxL = self._last_row_index
        f.kpsR, f.descR, rawR = self.detect(imgR)   # P:167
        f.idxR = self._last_row_index
        out["detected_feats"] = (len(f.kpsL), len(f.kpsR))
        first = self.prev is None
        if self.m_reset and self.prev is not None:    # P:254-267 (the reference crashes without a previous frame)
            n_ids = len(self.prev.ids) if self.prev.ids is not None else 0
            if n_ids:
                self.prev.ids = np.arange(n_ids, dtype=np.int64)
            self.m_last_match_ID = n_ids
            self.m_last_kf_max_id = self.m_last_match_ID - 1
            self.m_reset = False
        f.matches, raw_m = self.match_lr(f)           # P:269
        f.ids = None
        if self.p.vo_use_matches_ids and first:       # S3:67,172-173
            f.ids = np.arange(self.m_last_match_ID, self.m_last_match_ID + len(f.matches), dtype=np.int64)
            self.m_last_match_ID += len(f.matches)
        out["stereo_matches"] = len(f.matches)
        self.cur = f
        out["frame"] = f
        out["raw"] = dict(L=rawL, R=rawR, lr_matches=raw_m)
        if self.prev is not None:
            tracked, tr = self.track(self.prev, f)    # P:314
            out["tracked"] = tracked
            out["track_dbg"] = tr
            if self.p.vo_use_matches_ids:             # S4:268-305
                ids = np.zeros(len(f.matches), np.int64)
                got = np.zeros(len(f.matches), bool)
                for a, b in tracked:
                    ids[b] = self.prev.ids[a]
                    got[b] = True
                for k in range(len(ids)):
                    if not got[k]:
                        ids[k] = self.m_last_match_ID
                        self.m_last_match_ID += 1
                f.ids = ids
            # S4:743-752 (members); the RESULT fields are assigned at S5:724-725 only, when stage 5 runs to its end
            n_frame = len(tracked)
            n_kf = int((f.ids <= self.m_last_kf_max_id).sum()) if self.p.vo_use_matches_ids else 0
            out["m_num_tracked_pairs_from_last_frame"], out["m_num_tracked_pairs_from_last_kf"] = n_frame, n_kf
            out["tracked_feats_from_last_frame"] = out["tracked_feats_from_last_KF"] = 0
            if len(tracked) < self.p.bad_tracking_th:  # P:326-330
                self.m_error = out["error_code"] = VOEC["voecBadTracking"]
            if self.m_error != VOEC["voecBadTracking"]:
                res, residual, inl, dbg = self.optimize(self.prev, f, tracked, cam)  # P:338
                out["gn"] = res
                out["gn_dbg"] = dbg
                out["error_code"] = res.error_code
                out["valid"] = bool(res.valid)
                out["num_it"] = res.num_it
                out["num_it_final"] = res.num_it_final
                out["out_residual"] = residual
                out["outliers"] = inl
                if res.reached_end:
                    out["tracked_feats_from_last_frame"], out["tracked_feats_from_last_KF"] = n_frame, n_kf
                    out["pose"] = dict(rotvec=np.array(res.out_pose_rotvec[:]), t=np.array(res.out_pose_t[:]),
                                       mat=np.array(res.out_pose_mat[:]).reshape(4, 4),
                                       ypr=np.array(res.out_pose_ypr[:]), delta=np.array(res.delta_pose[:]))
        else:
            out["error_code"] = VOEC["voecFirstIteration"]  # P:348-352
            out["valid"] = False
        if not repeat:
            self.m_it_counter += 1
        return out


# ------------------------------------------------------------------------------------------------
# Multi-octave path (params_rectify.nOctaves > 1; the reference's ctor default is 3, stage1_rectify.cpp:28)
# ------------------------------------------------------------------------------------------------
PYR_MRPT_AUTO, PYR_AVG2X2, PYR_PYRDOWN = 0, 1, 2


def pyr_half_avg(img):
    """MRPT 1.x CImage::scaleHalfSmooth, SSE2 path (image_SSE2_scale_half_smooth_1c8u): two rows are averaged with
    _mm_avg_epu8 (rounding up), then horizontal neighbours with _mm_avg_epu16.  MRPT's source is not in this image:
    restated from its published implementation, UNPINNED."""
    img = _img(img)
    h2, w2 = img.shape[0] >> 1, img.shape[1] >> 1
    a = img[0:2 * h2:2, :2 * w2].astype(np.uint16)
    b = img[1:2 * h2:2, :2 * w2].astype(np.uint16)
    v = (a + b + 1) >> 1
    return ((v[:, 0::2] + v[:, 1::2] + 1) >> 1).astype(np.uint8)


def _reflect101(i, n):
    if n == 1:
        return 0
    while i < 0 or i >= n:
        i = -i if i < 0 else 2 * (n - 1) - i
    return i


def pyr_down5(img):
    """cv::pyrDown / cvPyrDown (MRPT's fallback in scaleHalfSmooth when the SSE2 conditions do not hold) into a
    (w>>1) x (h>>1) image: separable [1 4 6 4 1], integer accumulation, (sum + 128) >> 8, BORDER_REFLECT_101.
    Pinned against cv2.pyrDown in tests/test_octaves_cpu.py."""
    img = _img(img)
    h, w = img.shape
    h2, w2 = h >> 1, w >> 1
    k = (1, 4, 6, 4, 1)
    cols = [np.array([_reflect101(2 * x + d, w) for x in range(w2)]) for d in range(-2, 3)]
    rows = [np.array([_reflect101(2 * y + d, h) for y in range(h2)]) for d in range(-2, 3)]
    src = img.astype(np.int32)
    hor = sum(kk * src[:, c] for kk, c in zip(k, cols))          # (h, w2)
    ver = sum(kk * hor[r, :] for kk, r in zip(k, rows))          # (h2, w2)
    return ((ver + 128) >> 8).astype(np.uint8)


def pyr_half(img, mode=PYR_MRPT_AUTO):
    if mode == PYR_AVG2X2 or (mode == PYR_MRPT_AUTO and img.shape[1] % 16 == 0):
        return pyr_half_avg(img)
    return pyr_down5(img)


def build_pyramid(img, n_octaves, mode=PYR_MRPT_AUTO):
    """CImagePyramid::buildPyramidFast(img, nOctaves) (stage1_rectify.cpp:82-83; smooth_halves defaults to true)."""
    out = [_img(img)]
    for _ in range(1, n_octaves):
        out.append(pyr_half(out[-1], mode))
    return out


def kps_to_detect(orb_nfeats, n_octaves):
    """stage2_detect.cpp:404-407."""
    q0 = int(orb_nfeats * (2 * n_octaves) / (2.0 ** n_octaves - 1))
    return [q0] + [int(np.floor(q0 / 2.0 ** k + 0.5)) for k in range(1, n_octaves)]


class OracleEstimatorOctaves(OracleEstimator):
    """processNewImagePair with params_rectify.nOctaves >= 1 on the FAST+ORB path with either stereo matcher and either
    tracker: stage 2-4 run per octave (stage2_detect.cpp:423-624, stage3_match_left_right.cpp:87-177 / :208-408,
    stage4_match_consecutive.cpp:91-328 / :458-735),
    the match-ID generator and the tracking counters run through the octaves in order, stage 5 flattens all octaves with
    the x 2^octave rescale (stage5_optimization.cpp:419-461)."""

    def __init__(self, params=None, backend="cv2", ransac="cv2", sort_mode=0, pyr_mode=PYR_MRPT_AUTO):
        super().__init__(params, backend, ransac, sort_mode)
        self.pyr_mode = pyr_mode

    def set_this_frame_as_kf(self):                  # H:675-683: octave 0 only ("const size_t octave = 0")
        assert self.cur is not None and self.cur[0].ids is not None and len(self.cur[0].ids) > 0
        self.m_last_kf_max_id = int(self.cur[0].ids.max())

    def set_ids(self, ids):                          # H:685-694: octave 0 only
        f = self.cur if self.cur is not None else self.prev
        f[0].ids = np.array(ids, np.int64)

    def detect_octave(self, img, quota):
        H, W = img.shape
        th = self.m_current_fast_th
        if self.backend == "cv2":
            kps, desc = cv2_fast_orb(img, th)
        else:
            kps = fast9(img, th)
            kps, desc = orb_compute(img, kps)
        if self.p.non_maximal_suppression:
            kps, desc, _ = non_max_sup(quota, kps, desc, self.p.min_distance, H, W, self.sort_mode)
        kps, desc, _, row_index = update_indexes(kps, desc, H, self.sort_mode)
        return kps, desc, row_index

    def process_new_image_pair(self, imgL, imgR, cam, repeat=False):
        nOct = int(self.p.nOctaves)
        out = dict(error_code=VOEC["voecNone"], valid=False, num_it=0, num_it_final=0, outliers=np.zeros(0, np.int64),
                   out_residual=np.zeros(0), tracked=[np.zeros((0, 2), np.int64)] * nOct, pose=None)
        if (not repeat) and self.m_error not in (VOEC["voecBadTracking"], VOEC["voecBadCondNumber"]):
            self.prev = self.cur
        self.m_error = VOEC["voecNone"]
        pyrL = build_pyramid(imgL, nOct, self.pyr_mode)
        pyrR = build_pyramid(imgR, nOct, self.pyr_mode)
        quotas = kps_to_detect(self.p.orb_nfeats, nOct)
        frames = []
        for k in range(nOct):
            f = FrameData()
            f.H, f.W = pyrL[k].shape
            f.kpsL, f.descL, f.idxL = self.detect_octave(pyrL[k], quotas[k])
            f.kpsR, f.descR, f.idxR = self.detect_octave(pyrR[k], quotas[k])
            f.ids = None
            frames.append(f)
        out["detected_feats"] = [(len(f.kpsL), len(f.kpsR)) for f in frames]
        first = self.prev is None
        if self.m_reset and self.prev is not None:    # P:254-267
            self.m_last_match_ID = 0
            for pf in self.prev:
                n_ids = len(pf.ids) if pf.ids is not None else 0
                if n_ids:
                    pf.ids = np.arange(self.m_last_match_ID, self.m_last_match_ID + n_ids, dtype=np.int64)
                self.m_last_match_ID += n_ids
            self.m_reset = False
            self.m_last_kf_max_id = self.m_last_match_ID - 1
        for f in frames:                              # S3:87-177
            f.matches, _ = self.match_lr(f)
            if self.p.vo_use_matches_ids and first:
                f.ids = np.arange(self.m_last_match_ID, self.m_last_match_ID + len(f.matches), dtype=np.int64)
                self.m_last_match_ID += len(f.matches)
        out["stereo_matches"] = [len(f.matches) for f in frames]
        self.cur = frames
        out["frames"] = frames
        out["pyramid"] = (pyrL, pyrR)
        if self.prev is None:
            out["error_code"] = VOEC["voecFirstIteration"]
            if not repeat:
                self.m_it_counter += 1
            return out
        tracked_all, dbg_all = [], []
        for k in range(nOct):                         # S4:91-328
            pf, f = self.prev[k], frames[k]
            tracked, tr = self.track(pf, f)
            tracked_all.append(tracked)
            dbg_all.append(tr)
            if self.p.vo_use_matches_ids:
                ids = np.zeros(len(f.matches), np.int64)
                got = np.zeros(len(f.matches), bool)
                for a, b in tracked:
                    ids[b] = pf.ids[a]
                    got[b] = True
                for m in range(len(ids)):
                    if not got[m]:
                        ids[m] = self.m_last_match_ID
                        self.m_last_match_ID += 1
                f.ids = ids
        out["tracked"] = tracked_all
        out["track_dbg"] = dbg_all
        total = sum(len(t) for t in tracked_all)      # S4:743-752 (members; the result fields: S5:724-725)
        n_kf = (sum(int((f.ids <= self.m_last_kf_max_id).sum()) for f in frames) if self.p.vo_use_matches_ids else 0)
        out["m_num_tracked_pairs_from_last_frame"], out["m_num_tracked_pairs_from_last_kf"] = total, n_kf
        out["tracked_feats_from_last_frame"] = out["tracked_feats_from_last_KF"] = 0
        if total < self.p.bad_tracking_th:            # P:326-330
            self.m_error = out["error_code"] = VOEC["voecBadTracking"]
            if not repeat:
                self.m_it_counter += 1
            return out
        # S5:419-461: all octaves into four flat lists at the scale of octave 0
        def xy(kp, idx, s):
            return (np.stack([kp["x"][idx], kp["y"][idx]], 1).astype(np.float32) * np.float32(s)).astype(np.float32)
        L = dict(l1l=[], l1r=[], l2l=[], l2r=[], resp=[], cmi=[])
        for k in range(nOct):
            s = float(2 ** k) if nOct > 1 else 1.0
            pf, f, tp = self.prev[k], frames[k], tracked_all[k]
            pm = pf.matches[tp[:, 0]]
            cm = f.matches[tp[:, 1]]
            L["l1l"].append(xy(pf.kpsL, pm["queryIdx"], s)); L["l1r"].append(xy(pf.kpsR, pm["trainIdx"], s))
            L["l2l"].append(xy(f.kpsL, cm["queryIdx"], s)); L["l2r"].append(xy(f.kpsR, cm["trainIdx"], s))
            L["resp"].append(pf.kpsL["response"][pm["queryIdx"]].astype(np.float32))
            L["cmi"].append(tp[:, 1].astype(np.int64))
        flat = {k: np.concatenate(v) for k, v in L.items()}
        if self.p.use_previous_pose_as_initial:
            ini = self.m_last_computed_pose.copy()
        else:
            ini = np.zeros(6)
        H0, W0 = self.prev[0].H, self.prev[0].W
        res, residual, inl = stage5(flat["l1l"], flat["l1r"], flat["l2l"], flat["l2r"], flat["resp"], flat["cmi"], H0, W0,
                                    self.p.min_distance, cam, self.p.ls(), ini, self.sort_mode)
        if res.bad_cond:
            self.m_error = VOEC["voecBadCondNumber"]
        if res.reached_end and self.p.use_previous_pose_as_initial:
            self.m_last_computed_pose = np.array(res.delta_pose[:])
        out.update(gn=res, gn_dbg=dict(flat, ini=ini), error_code=res.error_code, valid=bool(res.valid), num_it=res.num_it,
                   num_it_final=res.num_it_final, out_residual=residual, outliers=inl)
        if res.reached_end:
            out["tracked_feats_from_last_frame"], out["tracked_feats_from_last_KF"] = total, n_kf
            out["pose"] = dict(rotvec=np.array(res.out_pose_rotvec[:]), t=np.array(res.out_pose_t[:]),
                               mat=np.array(res.out_pose_mat[:]).reshape(4, 4), ypr=np.array(res.out_pose_ypr[:]),
                               delta=np.array(res.delta_pose[:]))
        if not repeat:
            self.m_it_counter += 1
        return out
