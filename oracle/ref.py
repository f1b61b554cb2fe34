"""TEST INFRASTRUCTURE -- the reference's own code as a checker.

oracle/_ref/libsvo_ref.so is the reference's UNMODIFIED stage sources (common.cpp, process_new_image_pair.cpp,
stage1_rectify.cpp .. stage5_optimization.cpp under /root/reference/libstereo-odometry/src) compiled against the stand-in
third-party headers of oracle/ref_stubs/ (recipe: oracle/Makefile.ref).  The OpenCV calls the reference makes are
forwarded to the real OpenCV of this image (cv2) through C callbacks installed here, or to the C++ port of
oracle/svo_oracle.cpp (backend="port").  RefEstimator mirrors oracle.OracleEstimator's interface so that tests can
assert  restated oracle == compiled reference text  on identical inputs (tests/test_ref_cpu.py).

Only tests/, __graft_entry__ and bench.py's CPU legs may import this module; nothing under stereo_vo_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libsvo_ref.so")
REF_SRC = "/root/reference/libstereo-odometry"

KP = orc.KEYPOINT
DM = orc.DMATCH


def available():
    return os.path.exists(LIB) or os.path.isdir(REF_SRC)


def build(force=False):
    """Compiles the reference where /root/reference exists (this container); elsewhere the prebuilt library is used."""
    if os.path.isdir(REF_SRC) and (force or not os.path.exists(LIB) or any(
            os.path.getmtime(os.path.join(HERE, f)) > os.path.getmtime(LIB)
            for f in ("ref_glue.cpp", "Makefile.ref", os.path.join("ref_stubs", "svo_ref_stubs.h")))):
        subprocess.check_call(["make", "-s", "-f", "Makefile.ref"] + (["-B"] if force else []), cwd=HERE)
    if not os.path.exists(LIB):
        raise RuntimeError("oracle/_ref/libsvo_ref.so is missing and /root/reference is not here to build it from")
    return LIB


class _Backend(C.Structure):
    _fields_ = [("fast", C.c_void_p), ("orb_compute", C.c_void_p), ("orb_detect_and_compute", C.c_void_p),
                ("bf_match", C.c_void_p), ("find_fundamental", C.c_void_p), ("pyr_down", C.c_void_p)]


_FAST = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int)
_ORBC = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p)
_ORBD = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                    C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p)
_BFM = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p)
_FFM = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p)
_PYR = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int)

_lib = None
_state = dict(backend="cv2", ransac="cv2", keep=None, raised=0, last_fast=None)


def _arr(ptr, dtype, n):
    if n == 0:
        return np.zeros(0, dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n)


def _image(ptr, w, h, stride):
    a = _arr(ptr, np.uint8, stride * h).reshape(h, stride)[:, :w]
    return np.ascontiguousarray(a)


def _fast_kps_to_array(cvk):
    """cv2.KeyPoint list of FastFeatureDetector -> structured array without a Python loop per field: FAST reports size 7,
    angle -1, octave 0, class_id -1 for every corner (checked on the first one)"""
    import cv2
    n = len(cvk)
    out = np.zeros(n, KP)
    if n == 0:
        return out
    k0 = cvk[0]
    if (k0.size, k0.angle, k0.octave, k0.class_id) != (7.0, -1.0, 0, -1):
        return orc.cv2_kps_to_array(cvk)
    pts = cv2.KeyPoint_convert(cvk)
    out["x"], out["y"] = pts[:, 0], pts[:, 1]
    out["response"] = np.fromiter((k.response for k in cvk), np.float32, n)
    out["size"], out["angle"], out["class_id"] = 7.0, -1.0, -1
    return out


def _cb_fast(img, w, h, stride, th, out, cap):
    try:
        im = _image(img, w, h, stride)
        if _state["backend"] == "cv2":
            import cv2
            cvk = cv2.FastFeatureDetector_create(int(th)).detect(im)
            kps = _fast_kps_to_array(cvk)
            _state["last_fast"] = (cvk, kps)            # stage2_detect.cpp:511-512 hands the same list to ORB::compute next
        else:
            kps = orc.fast9(im, th)
        n = len(kps)
        if n <= cap:
            _arr(out, KP, n)[:] = kps
        return n
    except Exception:                                   # noqa: BLE001
        return -1


def _cb_orb_compute(img, w, h, stride, kps_io, n_in, desc_out):
    try:
        im = _image(img, w, h, stride)
        kin = _arr(kps_io, KP, n_in).copy()
        if _state["backend"] == "cv2":
            import cv2
            last = _state.get("last_fast")
            if last is not None and len(last[1]) == n_in and np.array_equal(last[1], kin):
                cvk = last[0]                            # the very list FAST returned: no need to rebuild 10 k Python objects
            else:
                cvk = [cv2.KeyPoint(float(k["x"]), float(k["y"]), float(k["size"]), float(k["angle"]), float(k["response"]),
                                    int(k["octave"]), int(k["class_id"])) for k in kin]
            cvk, desc = cv2.ORB_create().compute(im, cvk)
            kps = _fast_kps_to_array(cvk)
            desc = np.zeros((0, 32), np.uint8) if desc is None else desc
        else:
            kps, desc = orc.orb_compute(im, kin)
        n = len(kps)
        _arr(kps_io, KP, n)[:] = kps
        _arr(desc_out, np.uint8, n * 32)[:] = np.ascontiguousarray(desc, np.uint8).ravel()
        return n
    except Exception:                                   # noqa: BLE001
        return -1


def _cb_orb_detect(img, w, h, stride, nfeatures, scale, nlevels, edge, first, wta, score, patch, fast_th, out, cap, desc_out):
    try:
        import cv2
        im = _image(img, w, h, stride)
        orb = cv2.ORB_create(int(nfeatures), float(scale), int(nlevels), int(edge), int(first), int(wta), int(score), int(patch), int(fast_th))
        cvk, desc = orb.detectAndCompute(im, None)
        kps = orc.cv2_kps_to_array(cvk)
        n = len(kps)
        if n <= cap:
            _arr(out, KP, n)[:] = kps
            if n:
                _arr(desc_out, np.uint8, n * 32)[:] = np.ascontiguousarray(desc, np.uint8).ravel()
        return n
    except Exception:                                   # noqa: BLE001
        return -1


def _cb_bf_match(q, nq, t, nt, out):
    try:
        qa = _arr(q, np.uint8, nq * 32).reshape(nq, 32)
        ta = _arr(t, np.uint8, nt * 32).reshape(nt, 32)
        m = (orc.cv2_bf_match if _state["backend"] == "cv2" else orc.bf_match)(qa, ta)
        _arr(out, DM, len(m))[:] = m
        return len(m)
    except Exception:                                   # noqa: BLE001
        return -1


def _cb_find_fundamental(p1, p2, n, method, thr, conf, mask):
    try:
        a = _arr(p1, np.float32, 2 * n).reshape(n, 2).copy()
        b = _arr(p2, np.float32, 2 * n).reshape(n, 2).copy()
        assert (method, thr, conf) == (8, 1.0, 0.99)                 # FM_RANSAC, 1.0, 0.99: S4:56,202,237,684,696
        r = _state["ransac"]
        if r == "off":                                               # every point an inlier: the filter keeps everything
            m = np.ones(n, np.uint8) if n >= 8 else None
        elif r == "cv2":
            m = orc.cv2_fmat_mask(a, b)
        elif r == "cv2_tolerant":
            m = orc.cv2_fmat_mask_tolerant(a, b)
        else:
            m = r(a, b)
        if m is None:
            return 0
        _arr(mask, np.uint8, n)[:] = np.asarray(m, np.uint8).ravel()
        return n
    except Exception:                                   # noqa: BLE001  (cv2 4.13 raises on some inputs: DESIGN.md section 4)
        _state["raised"] += 1
        return -1


def _cb_pyr_down(img, w, h, stride, out, ow, oh):
    try:
        im = _image(img, w, h, stride)
        if _state["backend"] == "cv2":
            import cv2
            d = cv2.pyrDown(im, dstsize=(int(ow), int(oh)))
        else:
            d = orc.pyr_down5(im)
        _arr(out, np.uint8, ow * oh)[:] = np.ascontiguousarray(d, np.uint8).ravel()
        return 0
    except Exception:                                   # noqa: BLE001
        return -1


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.svo_ref_create.restype = C.c_void_p
        L.svo_ref_last_error.restype = C.c_char_p
        for name in ("svo_ref_destroy", "svo_ref_reset_thresholds", "svo_ref_reset_ids"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = None
        L.svo_ref_set_param.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.svo_ref_set_out_dir.argtypes = [C.c_void_p, C.c_char_p]
        cbs = (_FAST(_cb_fast), _ORBC(_cb_orb_compute), _ORBD(_cb_orb_detect), _BFM(_cb_bf_match),
               _FFM(_cb_find_fundamental), _PYR(_cb_pyr_down))
        _state["keep"] = cbs
        b = _Backend(*[C.cast(c, C.c_void_p) for c in cbs])
        L.svo_ref_set_backend(C.byref(b))
        _lib = L
    return _lib


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def _camv(cam):
    return np.array([cam.fxl, cam.fyl, cam.cxl, cam.cyl, cam.fxr, cam.fyr, cam.cxr, cam.cyr, cam.baseline], np.float64)


class RefEstimator:
    """rso::CStereoOdometryEstimator itself (compiled from the reference's text), driven like oracle.OracleEstimator."""

    PARAM_MAP = dict(ifm_method="ifm_method")

    def __init__(self, params=None, backend="cv2", ransac="cv2"):
        self.p = params or orc.Params()
        self.backend, self.ransac = backend, ransac
        self.L = lib()
        self.h = C.c_void_p(self.L.svo_ref_create())
        for k, v in vars(self.p).items():
            rc = self.L.svo_ref_set_param(self.h, k.encode(), float(v))
            assert rc == 0, self.L.svo_ref_last_error()
        self.L.svo_ref_reset_thresholds(self.h)          # what loadParamsFromConfigFile does last (H:661-662)

    def close(self):
        if self.h:
            self.L.svo_ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:                               # noqa: BLE001
            pass

    def _enter(self):
        _state["backend"], _state["ransac"] = self.backend, self.ransac

    # -- controls (libstereo-odometry.h:529-541, 675-726) ------------------------------------------------
    def set_fast_threshold(self, v):
        self.L.svo_ref_set_fast_threshold(self.h, int(v))

    def set_orb_threshold(self, v):
        self.L.svo_ref_set_orb_threshold(self.h, int(v))

    def thresholds(self):
        return self.L.svo_ref_get_fast_threshold(self.h), self.L.svo_ref_get_orb_threshold(self.h)

    def set_this_frame_as_kf(self):
        assert self.L.svo_ref_set_this_frame_as_kf(self.h) == 0, self.L.svo_ref_last_error()

    def reset_ids(self):
        self.L.svo_ref_reset_ids(self.h)

    def set_max_match_id(self, i):
        self.L.svo_ref_set_max_match_id(self.h, C.c_int64(int(i)))

    def set_ids(self, ids):
        a = np.ascontiguousarray(ids, np.int64)
        self.L.svo_ref_set_ids(self.h, _vp(a), len(a))

    def set_save_files(self, out_dir):
        os.makedirs(out_dir, exist_ok=True)
        self.L.svo_ref_set_param(self.h, b"vo_save_files", 1.0)
        self.L.svo_ref_set_out_dir(self.h, out_dir.encode())

    # -- readers -----------------------------------------------------------------------------------------
    def _get(self, fn, dtype, *args, width=1):
        n = fn(self.h, *args, None, 0)
        if n <= 0:
            return np.zeros((0, width) if width > 1 else 0, dtype)
        out = np.zeros(n * width, dtype)
        fn(self.h, *args, _vp(out), n)
        return out.reshape(n, width) if width > 1 else out

    def frame(self, which=0, octave=0):
        f = orc.FrameData()
        L = self.L
        f.kpsL = self._get(L.svo_ref_get_keypoints, KP, which, 0, octave)
        f.kpsR = self._get(L.svo_ref_get_keypoints, KP, which, 1, octave)
        f.descL = self._get(L.svo_ref_get_descriptors, np.uint8, which, 0, octave, width=32).reshape(-1, 32)
        f.descR = self._get(L.svo_ref_get_descriptors, np.uint8, which, 1, octave, width=32).reshape(-1, 32)
        f.matches = self._get(L.svo_ref_get_matches, DM, which, octave)
        ids = self._get(L.svo_ref_get_ids, np.int64, which, octave)
        f.ids = ids if len(ids) else None
        f.idxL = self._get(L.svo_ref_get_feats_index, np.int64, which, 0, octave)
        f.idxR = self._get(L.svo_ref_get_feats_index, np.int64, which, 1, octave)
        w, h = C.c_int(0), C.c_int(0)
        L.svo_ref_get_pyramid(self.h, which, 0, octave, None, C.byref(w), C.byref(h))
        f.W, f.H = w.value, h.value
        return f

    def row_index(self, which=0, octave=0):
        return self._get(self.L.svo_ref_get_row_index, np.int64, which, octave)

    def pyramid(self, which, side, octave):
        w, h = C.c_int(0), C.c_int(0)
        if self.L.svo_ref_get_pyramid(self.h, which, side, octave, None, C.byref(w), C.byref(h)) != 0:
            return None
        out = np.zeros((h.value, w.value), np.uint8)
        self.L.svo_ref_get_pyramid(self.h, which, side, octave, _vp(out), C.byref(w), C.byref(h))
        return out

    def tracked(self, octave=0):
        return self._get(self.L.svo_ref_get_tracked, np.int64, octave, width=2).reshape(-1, 2)

    @staticmethod
    def _result(r):
        out = dict(error_code=int(r[0]), valid=bool(r[1]), num_it=int(r[2]), num_it_final=int(r[3]),
                   tracked_feats_from_last_frame=int(r[4]), tracked_feats_from_last_KF=int(r[5]))
        mat = np.eye(4)
        mat[:3, :3] = r[8:17].reshape(3, 3)
        mat[:3, 3] = r[17:20]
        out["pose"] = dict(mat=mat, t=r[17:20].copy(), ypr=r[20:23].copy(), delta=r[23:29].copy())
        out["state"] = dict(m_error=int(r[29]), fast_th=int(r[30]), orb_th=int(r[31]), last_match_id=int(r[32]),
                            last_kf_max_id=int(r[33]), it_counter=int(r[34]))
        return out

    # -- processNewImagePair -------------------------------------------------------------------------------
    def process_new_image_pair(self, imgL, imgR, cam, repeat=False, light=False):
        """light=True: only the TStereoOdometryResult fields (what a caller of the class gets back); the default also reads
        the frame data, the per-octave counters and the tracked pairs for the parity tests"""
        self._enter()
        imgL = np.ascontiguousarray(imgL, np.uint8)
        imgR = np.ascontiguousarray(imgR, np.uint8)
        h, w = imgL.shape
        r = np.zeros(40)
        cv = _camv(cam)
        rc = self.L.svo_ref_process(self.h, _vp(imgL), _vp(imgR), w, h, w, _vp(cv), int(bool(repeat)), _vp(r))
        if rc != 0:
            raise RuntimeError("the reference threw: " + self.L.svo_ref_last_error().decode(errors="replace"))
        out = self._result(r)
        out["outliers"] = self._get(self.L.svo_ref_get_outliers, np.int64)
        out["out_residual"] = self._get(self.L.svo_ref_get_residuals, np.float64)
        if light:
            return out
        n_oct = self.L.svo_ref_n_octaves(self.h, 0)
        out["frames"] = [self.frame(0, k) for k in range(n_oct)]
        out["frame"] = out["frames"][0]
        cnt = np.zeros(3 * max(n_oct, 1), np.int64)
        self.L.svo_ref_get_counts(self.h, _vp(cnt), n_oct)
        out["detected_feats"] = [(int(cnt[3 * k]), int(cnt[3 * k + 1])) for k in range(n_oct)]
        out["stereo_matches"] = [int(cnt[3 * k + 2]) for k in range(n_oct)]
        out["tracked"] = [self.tracked(k) for k in range(n_oct)]
        return out

    # -- getChangeInPose (common.cpp:355-413) ----------------------------------------------------------------
    def get_change_in_pose(self, tracked, pre_m, cur_m, pl, pr, cl, cr, cam, w, h, ini=None):
        self._enter()
        tp = np.ascontiguousarray(tracked, np.int64).reshape(-1, 2)
        arrs = [np.ascontiguousarray(a, DM) for a in (pre_m, cur_m)] + [np.ascontiguousarray(a, KP) for a in (pl, pr, cl, cr)]
        ini = np.zeros(6) if ini is None else np.ascontiguousarray(ini, np.float64)
        r = np.zeros(40)
        cv = _camv(cam)
        rc = self.L.svo_ref_get_change_in_pose(self.h, _vp(tp), len(tp), _vp(arrs[0]), len(arrs[0]), _vp(arrs[1]), len(arrs[1]),
                                               _vp(arrs[2]), len(arrs[2]), _vp(arrs[3]), len(arrs[3]), _vp(arrs[4]), len(arrs[4]),
                                               _vp(arrs[5]), len(arrs[5]), _vp(cv), int(w), int(h), _vp(ini), _vp(r))
        if rc != 0:
            raise RuntimeError("the reference threw: " + self.L.svo_ref_last_error().decode(errors="replace"))
        out = self._result(r)
        out["outliers"] = self._get(self.L.svo_ref_get_outliers, np.int64)
        out["out_residual"] = self._get(self.L.svo_ref_get_residuals, np.float64)
        return out

    # -- private members, called directly ----------------------------------------------------------------------
    def non_max_sup(self, num_out, kps, desc, H, W):
        kps = np.ascontiguousarray(kps, KP)
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        ok = np.zeros(len(kps), KP)
        od = np.zeros((len(kps), 32), np.uint8)
        n = self.L.svo_ref_non_max_sup(self.h, int(num_out), _vp(kps), len(kps), _vp(desc), int(H), int(W), _vp(ok), _vp(od))
        return ok[:n].copy(), od[:n].copy()

    def non_max_sup_aux(self, kps, H, W, num_out):
        kps = np.ascontiguousarray(kps, KP)
        s = np.zeros(max(len(kps), 1), np.uint8)
        self.L.svo_ref_non_max_sup_aux(self.h, _vp(kps), len(kps), int(H), int(W), int(num_out), _vp(s))
        return s[:len(kps)].astype(bool)

    def update_indexes(self, kps, desc, H, W):
        kps = np.ascontiguousarray(kps, KP).copy()
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32).copy()
        idx = np.zeros(H, np.int64)
        self.L.svo_ref_update_indexes(self.h, _vp(kps), len(kps), _vp(desc), int(H), int(W), _vp(idx))
        return kps, desc, idx

    def pinhole_projection(self, lmks, cam, pose6):
        lm = np.ascontiguousarray(lmks, np.float64).reshape(-1, 3)
        px = np.zeros((len(lm), 4), np.float32)
        J = np.zeros((len(lm), 4, 6))
        cv = _camv(cam)
        p6 = np.ascontiguousarray(pose6, np.float64)
        self.L.svo_ref_pinhole_projection(self.h, _vp(lm), len(lm), _vp(cv), _vp(p6), _vp(px), _vp(J))
        return px, J
