"""ctypes binding of the C-ABI in include/svo_b200.h (libsvo_b200.so).

This is the Python host side of the drop-in boundary.  It contains no arithmetic of the hot path:
everything is computed by the hand-written sm_100a kernels inside libsvo_b200.so, and every call
fails loudly (RuntimeError) when the library or a CUDA device is missing -- there is no CPU fallback.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsvo_b200.so")
CSRC = os.path.join(_HERE, "csrc")

KEYPOINT = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])   # cv::KeyPoint
DMATCH = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])  # cv::DMatch

SVO_OK, SVO_ERR_ARG, SVO_ERR_CUDA, SVO_ERR_CAPACITY, SVO_ERR_STATE = 0, -1, -2, -3, -4

VOEC = {"voecNone": 0, "voecBadCondNumber": 1, "voecIncrFuncCostStg1": 2, "voecIncrFuncCostStg2": 3,
        "voecFirstIteration": 4, "voecBadTracking": 5}

# every symbol declared in include/svo_b200.h (checked by tests/test_capi_symbols.py)
EXPORTS = [
    "svo_default_config", "svo_create", "svo_destroy", "svo_last_error", "svo_reset_slot",
    "svo_process_batch", "svo_submit_batch", "svo_wait_batch", "svo_process_batch_device", "svo_fetch_results", "svo_synchronize", "svo_process_pair",
    "svo_get_keypoints", "svo_get_descriptors", "svo_get_matches", "svo_get_match_ids", "svo_get_tracked_pairs",
    "svo_get_residuals", "svo_get_outliers", "svo_get_track_candidates", "svo_get_plane",
    "svo_detect", "svo_match_hamming", "svo_match_hamming_device", "svo_match_lr", "svo_get_change_in_pose",
    "svo_fundmat_ransac", "svo_fundmat_7point", "svo_set_thresholds", "svo_get_thresholds", "svo_host_alloc", "svo_host_free", "svo_set_result_detail", "svo_wait_batch_detail", "svo_get_track_candidates_oct", "svo_get_stage5_survivors", "svo_get_ransac_stats", "svo_abi_version", "svo_kernel_launches", "svo_stream",
    "svo_enable_stage_timing", "svo_get_stage_times", "svo_measure_pipe_peaks", "svo_set_concurrency",
    "svo_set_input_format", "svo_set_rectify_maps", "svo_prepare_rectify", "svo_prepare_rectify_device",
    "svo_get_projected_coords", "svo_set_this_frame_as_kf", "svo_reset_ids", "svo_set_ids", "svo_set_max_match_id",
    "svo_get_keypoints_oct", "svo_get_descriptors_oct", "svo_get_matches_oct", "svo_get_match_ids_oct",
    "svo_get_tracked_pairs_oct", "svo_get_octave_info", "svo_get_octave_stats", "svo_get_pyramid_plane",
    "svo_track", "svo_optimize", "svo_build_pyramid_device", "svo_get_orb_raw_keypoints", "svo_get_orb_level",
]


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("max_seqs", C.c_int32), ("max_corners", C.c_int32),
                ("orb_nfeats", C.c_int32), ("fast_threshold", C.c_int32), ("non_maximal_suppression", C.c_int32),
                ("min_distance", C.c_int32),
                ("max_y_diff", C.c_double), ("orb_threshold", C.c_int32), ("enable_robust_1to1_match", C.c_int32),
                ("fundmat_filter", C.c_int32),
                ("use_robust_kernel", C.c_int32), ("kernel_param", C.c_double), ("max_iters", C.c_int32),
                ("initial_max_iters", C.c_int32), ("min_mod_out_vector", C.c_double), ("max_incr_cost", C.c_int32),
                ("residual_threshold", C.c_double), ("bad_tracking_th", C.c_int32),
                ("use_previous_pose_as_initial", C.c_int32), ("vo_use_matches_ids", C.c_int32), ("sort_mode", C.c_int32),
                ("match_method", C.c_int32), ("if_match_method", C.c_int32), ("ifm_win_w", C.c_int32),
                ("ifm_win_h", C.c_int32), ("orb_max_distance", C.c_double),
                ("n_octaves", C.c_int32), ("pyr_mode", C.c_int32), ("detect_method", C.c_int32), ("orb_nlevels", C.c_int32)]


class Camera(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("fxl", "fyl", "cxl", "cyl", "fxr", "fyr", "cxr", "cyr", "baseline")]


class Result(C.Structure):
    _fields_ = [("out_pose", C.c_double * 16), ("out_pose_rotvec", C.c_double * 3), ("out_pose_t", C.c_double * 3),
                ("out_pose_ypr", C.c_double * 3), ("delta_pose", C.c_double * 6),
                ("num_it", C.c_int32), ("num_it_final", C.c_int32), ("valid", C.c_int32), ("error_code", C.c_int32),
                ("tracked_feats_from_last_KF", C.c_int32), ("tracked_feats_from_last_frame", C.c_int32),
                ("detected_left", C.c_int32), ("detected_right", C.c_int32), ("stereo_matches", C.c_int32),
                ("n_outliers", C.c_int32), ("n_residuals", C.c_int32), ("status", C.c_int32)]


def build(verbose=False):
    """Compile every CUDA source for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libsvo_b200.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)
    host = subprocess.run(["make", "-C", os.path.join(_HERE, "host")], capture_output=True, text=True)
    if host.returncode != 0:
        raise RuntimeError("building the host shim failed:\n" + host.stdout + host.stderr)
    return LIB_PATH


_lib = None


def lib():
    """Load libsvo_b200.so; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(stereo_vo_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.svo_last_error.restype = C.c_char_p
        L.svo_last_error.argtypes = [C.c_void_p]
        L.svo_kernel_launches.restype = C.c_int64
        L.svo_kernel_launches.argtypes = [C.c_void_p]
        L.svo_stream.restype = C.c_void_p
        L.svo_stream.argtypes = [C.c_void_p]
        L.svo_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
        L.svo_destroy.argtypes = [C.c_void_p]
        L.svo_destroy.restype = None
        for name in EXPORTS:
            getattr(L, name)
        _lib = L
    return _lib


def default_config(**kw):
    cfg = Config()
    lib().svo_default_config(C.byref(cfg))
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


def make_camera(fx, fy, cx, cy, baseline):
    return Camera(fx, fy, cx, cy, fx, fy, cx, cy, baseline)


def _vp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class SvoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"svo_b200 error {code}: {msg}")
        self.code = code


class Context:
    """Thin object wrapper of svo_ctx.  One per GPU; `max_seqs` independent sequence slots."""

    def __init__(self, cfg=None, **kw):
        self.cfg = cfg if cfg is not None else default_config(**kw)
        self._h = C.c_void_p()
        rc = lib().svo_create(C.byref(self.cfg), C.byref(self._h))
        if rc != SVO_OK:
            raise SvoError(rc, lib().svo_last_error(None).decode())
        self.n_octaves = int(self.cfg.n_octaves)
        self.octave_info = [self._octave_info(k) for k in range(self.n_octaves)]   # (W, H, quota, kp_cap) per octave
        self.kp_cap = sum(o[3] for o in self.octave_info)      # room for any per-octave or flattened list

    def close(self):
        if getattr(self, "_h", None):
            lib().svo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise SvoError(rc, lib().svo_last_error(self._h).decode())
        return rc

    # ---- pipeline ---------------------------------------------------------------------------------
    def process_batch(self, left, right, cams, slots=None, repeat=None):
        """left/right: uint8 arrays (n, H, W) or lists of (H, W) arrays in host memory."""
        L = [np.ascontiguousarray(a, np.uint8) for a in left]
        R = [np.ascontiguousarray(a, np.uint8) for a in right]
        n = len(L)
        H, W = L[0].shape[:2]
        assert (H, W) == (self.cfg.height, self.cfg.width)
        lp = (C.c_void_p * n)(*[a.ctypes.data for a in L])
        rp = (C.c_void_p * n)(*[a.ctypes.data for a in R])
        cam_arr = (Camera * n)(*cams)
        sl = None if slots is None else (C.c_int32 * n)(*slots)
        rep = None if repeat is None else (C.c_uint8 * n)(*[1 if r else 0 for r in repeat])
        res = (Result * n)()
        self._check(lib().svo_process_batch(self._h, n, sl, lp, rp, C.c_size_t(L[0].strides[0]), cam_arr, rep, res))
        return list(res)

    def submit_batch(self, left, right, cams, slots=None, repeat=None):
        """Asynchronous half of process_batch; keeps the host arrays alive until wait_batch."""
        L = [np.ascontiguousarray(a, np.uint8) for a in left]
        R = [np.ascontiguousarray(a, np.uint8) for a in right]
        n = len(L)
        lp = (C.c_void_p * n)(*[a.ctypes.data for a in L])
        rp = (C.c_void_p * n)(*[a.ctypes.data for a in R])
        cam_arr = (Camera * n)(*cams)
        sl = None if slots is None else (C.c_int32 * n)(*slots)
        rep = None if repeat is None else (C.c_uint8 * n)(*[1 if r else 0 for r in repeat])
        self._check(lib().svo_submit_batch(self._h, n, sl, lp, rp, C.c_size_t(L[0].strides[0]), cam_arr, rep))
        self._inflight = getattr(self, "_inflight", []) + [(L, R, n)]

    def wait_batch(self):
        L, R, n = self._inflight.pop(0)
        res = (Result * n)()
        self._check(lib().svo_wait_batch(self._h, res))
        return list(res)

    def process_pair(self, left, right, cam, repeat=False):
        left = np.ascontiguousarray(left, np.uint8)
        right = np.ascontiguousarray(right, np.uint8)
        res = Result()
        self._check(lib().svo_process_pair(self._h, _vp(left), _vp(right), C.c_size_t(left.strides[0]), C.byref(cam),
                                           int(bool(repeat)), C.byref(res)))
        return res

    def process_batch_device(self, d_left_ptr, d_right_ptr, n, stride, image_pitch, cams, slots=None, repeat=None):
        cam_arr = (Camera * n)(*cams)
        sl = None if slots is None else (C.c_int32 * n)(*slots)
        rep = None if repeat is None else (C.c_uint8 * n)(*[1 if r else 0 for r in repeat])
        self._check(lib().svo_process_batch_device(self._h, n, sl, C.c_void_p(d_left_ptr), C.c_void_p(d_right_ptr),
                                                   C.c_size_t(stride), C.c_size_t(image_pitch), cam_arr, rep))

    def fetch_results(self, n, slots=None):
        sl = None if slots is None else (C.c_int32 * n)(*slots)
        res = (Result * n)()
        self._check(lib().svo_fetch_results(self._h, n, sl, res))
        return list(res)

    def synchronize(self):
        self._check(lib().svo_synchronize(self._h))

    def reset_slot(self, slot):
        self._check(lib().svo_reset_slot(self._h, slot))

    def reset_all(self):
        self._check(lib().svo_reset_slot(self._h, -1))

    # ---- readers (octave = index into the image pyramid, params_rectify.nOctaves) ------------------
    def _octave_info(self, octave):
        out = (C.c_int32 * 4)()
        self._check(lib().svo_get_octave_info(self._h, octave, out))
        return tuple(out)

    def octave_stats(self, slot=0):
        """(n_octaves, 4) array: detected left, detected right, stereo matches, tracked pairs of the last step."""
        out = np.zeros((self.n_octaves, 4), np.int32)
        self._check(lib().svo_get_octave_stats(self._h, slot, _vp(out), self.n_octaves))
        return out

    def pyramid_plane(self, item=0, octave=1, side=0):
        w, h = self.octave_info[octave][:2]
        out = np.zeros((h, w), np.uint8)
        self._check(lib().svo_get_pyramid_plane(self._h, item, octave, side, _vp(out)))
        return out

    def keypoints(self, slot=0, which=0, side=0, octave=0):
        out = np.zeros(self.kp_cap, KEYPOINT)
        n = self._check(lib().svo_get_keypoints_oct(self._h, slot, octave, which, side, _vp(out), self.kp_cap))
        return out[:n].copy()

    def descriptors(self, slot=0, which=0, side=0, octave=0):
        out = np.zeros((self.kp_cap, 32), np.uint8)
        n = self._check(lib().svo_get_descriptors_oct(self._h, slot, octave, which, side, _vp(out), self.kp_cap))
        return out[:n].copy()

    def matches(self, slot=0, which=0, octave=0):
        out = np.zeros(self.kp_cap, DMATCH)
        n = self._check(lib().svo_get_matches_oct(self._h, slot, octave, which, _vp(out), self.kp_cap))
        return out[:n].copy()

    def match_ids(self, slot=0, which=0, octave=0):
        out = np.zeros(self.kp_cap, np.int64)
        n = self._check(lib().svo_get_match_ids_oct(self._h, slot, octave, which, _vp(out), self.kp_cap))
        return out[:n].copy()

    def tracked_pairs(self, slot=0, octave=0):
        out = np.zeros((self.kp_cap, 2), np.int64)
        n = self._check(lib().svo_get_tracked_pairs_oct(self._h, slot, octave, _vp(out), self.kp_cap))
        return out[:n].copy()

    def residuals(self, slot=0):
        out = np.zeros(self.kp_cap, np.float64)
        n = self._check(lib().svo_get_residuals(self._h, slot, _vp(out), self.kp_cap))
        return out[:n].copy()

    def outliers(self, slot=0):
        out = np.zeros(self.kp_cap, np.int64)
        n = self._check(lib().svo_get_outliers(self._h, slot, _vp(out), self.kp_cap))
        return out[:n].copy()

    def track_candidates(self, slot=0):
        q = np.zeros(self.kp_cap, np.int32)
        tl = np.zeros(self.kp_cap, np.int32)
        tr = np.zeros(self.kp_cap, np.int32)
        il = np.zeros(self.kp_cap, np.uint8)
        ir = np.zeros(self.kp_cap, np.uint8)
        n = self._check(lib().svo_get_track_candidates(self._h, slot, _vp(q), _vp(tl), _vp(tr), _vp(il), _vp(ir), self.kp_cap))
        return q[:n].copy(), tl[:n].copy(), tr[:n].copy(), il[:n].copy(), ir[:n].copy()

    def orb_raw_keypoints(self, slot=0, side=0):
        """cv::ORB::detectAndCompute's keypoints of the last step (dmORB), before m_non_max_sup / m_update_indexes"""
        cap = 4 * self.cfg.orb_nfeats + 256
        out = np.zeros(cap, KEYPOINT)
        n = self._check(lib().svo_get_orb_raw_keypoints(self._h, slot, side, _vp(out), cap))
        return out[:n].copy()

    def orb_level(self, slot=0, side=0, level=1, plane=0):
        w, h = C.c_int32(0), C.c_int32(0)
        self._check(lib().svo_get_orb_level(self._h, slot, side, level, plane, None, C.byref(w), C.byref(h)))
        out = np.zeros((h.value, w.value), np.uint8)
        self._check(lib().svo_get_orb_level(self._h, slot, side, level, plane, _vp(out), C.byref(w), C.byref(h)))
        return out

    def plane(self, slot=0, side=0, plane=0):
        out = np.zeros((self.cfg.height, self.cfg.width), np.uint8)
        self._check(lib().svo_get_plane(self._h, slot, side, plane, _vp(out)))
        return out

    # ---- per-stage entries --------------------------------------------------------------------------
    def detect(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        k = np.zeros(self.kp_cap, KEYPOINT)
        d = np.zeros((self.kp_cap, 32), np.uint8)
        n = self._check(lib().svo_detect(self._h, _vp(img), C.c_size_t(img.shape[1]), _vp(k), _vp(d), self.kp_cap))
        return k[:n].copy(), d[:n].copy()

    def match_hamming(self, q, t):
        q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
        t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
        out = np.zeros(max(len(q), 1), DMATCH)
        n = self._check(lib().svo_match_hamming(self._h, _vp(q), len(q), _vp(t), len(t), _vp(out)))
        return out[:n].copy()

    def match_hamming_device(self, d_q, nq, d_t, nt, reps, d_out):
        self._check(lib().svo_match_hamming_device(self._h, C.c_void_p(d_q), nq, C.c_void_p(d_t), nt, reps, C.c_void_p(d_out)))

    def match_lr(self, kpsL, descL, kpsR, descR):
        kpsL = np.ascontiguousarray(kpsL, KEYPOINT)
        kpsR = np.ascontiguousarray(kpsR, KEYPOINT)
        descL = np.ascontiguousarray(descL, np.uint8).reshape(-1, 32)
        descR = np.ascontiguousarray(descR, np.uint8).reshape(-1, 32)
        out = np.zeros(self.kp_cap, DMATCH)
        n = self._check(lib().svo_match_lr(self._h, _vp(kpsL), _vp(descL), len(kpsL), _vp(kpsR), _vp(descR), len(kpsR),
                                           _vp(out), self.kp_cap))
        return out[:n].copy()

    def track(self, pre, cur):
        """stage 4 alone: pre / cur = (kpsL, descL, kpsR, descR, matches) of the previous / current frame."""
        args = []
        for kl, dl, kr, dr, m in (pre, cur):
            kl = np.ascontiguousarray(kl, KEYPOINT)
            kr = np.ascontiguousarray(kr, KEYPOINT)
            dl = np.ascontiguousarray(dl, np.uint8).reshape(-1, 32)
            dr = np.ascontiguousarray(dr, np.uint8).reshape(-1, 32)
            m = np.ascontiguousarray(m, DMATCH)
            args += [_vp(kl), _vp(dl), len(kl), _vp(kr), _vp(dr), len(kr), _vp(m), len(m)]
            self._keep = getattr(self, "_keep", []) + [kl, dl, kr, dr, m]
        out = np.zeros((self.kp_cap, 2), np.int64)
        n = self._check(lib().svo_track(self._h, *args, _vp(out), self.kp_cap))
        self._keep = []
        return out[:n].copy()

    def get_change_in_pose(self, tracked_pairs, pre_matches, cur_matches, pre_left, pre_right, cur_left, cur_right,
                           cam, ini_estimation=None, slot=0):
        tp = np.ascontiguousarray(tracked_pairs, np.int64).reshape(-1, 2)
        pm = np.ascontiguousarray(pre_matches, DMATCH)
        cm = np.ascontiguousarray(cur_matches, DMATCH)
        kl = [np.ascontiguousarray(k, KEYPOINT) for k in (pre_left, pre_right, cur_left, cur_right)]
        ini = None if ini_estimation is None else np.ascontiguousarray(ini_estimation, np.float64)
        res = Result()
        self._check(lib().svo_get_change_in_pose(self._h, slot, _vp(tp), len(tp), _vp(pm), len(pm), _vp(cm), len(cm),
                                                 _vp(kl[0]), len(kl[0]), _vp(kl[1]), len(kl[1]), _vp(kl[2]), len(kl[2]),
                                                 _vp(kl[3]), len(kl[3]), C.byref(cam), int(ini is not None), _vp(ini),
                                                 C.byref(res)))
        return res

    def get_projected_coords(self, pre_matches, pre_left, pre_right, other_matches_tracked, cam, delta_pose):
        pm = np.ascontiguousarray(pre_matches, DMATCH)
        kl = np.ascontiguousarray(pre_left, KEYPOINT)
        kr = np.ascontiguousarray(pre_right, KEYPOINT)
        fl = np.ascontiguousarray(other_matches_tracked, np.int32)
        dp = np.ascontiguousarray(delta_pose, np.float64)
        assert len(fl) == len(pm) and len(dp) == 6
        out = np.zeros((max(len(pm), 1), 4), np.float32)
        n = self._check(lib().svo_get_projected_coords(self._h, _vp(pm), len(pm), _vp(kl), len(kl), _vp(kr), len(kr), _vp(fl),
                                                       C.byref(cam), _vp(dp), _vp(out), len(out)))
        return out[:n].copy()

    def fundmat_ransac(self, p1, p2):
        p1 = np.ascontiguousarray(p1, np.float32).reshape(-1, 2)
        p2 = np.ascontiguousarray(p2, np.float32).reshape(-1, 2)
        mask = np.zeros(max(len(p1), 1), np.uint8)
        n = self._check(lib().svo_fundmat_ransac(self._h, _vp(p1), _vp(p2), len(p1), _vp(mask)))
        return n, mask[:len(p1)].copy()

    def fundmat_7point(self, p1, p2):
        """run7Point on n sets of 7 correspondences: (n_models[n], F[n, 3, 3, 3])."""
        p1 = np.ascontiguousarray(p1, np.float32).reshape(-1, 7, 2)
        p2 = np.ascontiguousarray(p2, np.float32).reshape(-1, 7, 2)
        n = len(p1)
        F = np.zeros((max(n, 1), 27), np.float64)
        nm = np.zeros(max(n, 1), np.int32)
        self._check(lib().svo_fundmat_7point(self._h, _vp(p1), _vp(p2), n, _vp(F), _vp(nm)))
        return nm[:n].copy(), F[:n].reshape(n, 3, 3, 3).copy()

    def ransac_stats(self, slot=0):
        """{points, inliers, iterations} of the left-left and right-right findFundamentalMat calls of the last step"""
        out = np.zeros(6, np.int32)
        self._check(lib().svo_get_ransac_stats(self._h, int(slot), _vp(out)))
        return out.reshape(2, 3)

    # ---- introspection --------------------------------------------------------------------------------
    def kernel_launches(self):
        return int(lib().svo_kernel_launches(self._h))

    def stream(self):
        return lib().svo_stream(self._h)

    # ---- dynamic thresholds (libstereo-odometry.h:529-541) -----------------------------------------------
    def set_thresholds(self, fast_th=-1, orb_th=-1, slot=-1):
        """m_current_fast_th / m_current_orb_th of one slot (-1: every slot); a negative value keeps that threshold."""
        self._check(lib().svo_set_thresholds(self._h, int(slot), int(fast_th), int(orb_th)))

    def thresholds(self, slot=0):
        f, o = C.c_int32(0), C.c_int32(0)
        self._check(lib().svo_get_thresholds(self._h, int(slot), C.byref(f), C.byref(o)))
        return f.value, o.value

    # ---- match IDs / key frames (libstereo-odometry.h:675-726) ---------------------------------------------
    def set_this_frame_as_kf(self, slot=0):
        self._check(lib().svo_set_this_frame_as_kf(self._h, slot))

    def reset_ids(self, slot=0):
        self._check(lib().svo_reset_ids(self._h, slot))

    def set_ids(self, ids, slot=0):
        ids = np.ascontiguousarray(ids, np.int64)
        self._check(lib().svo_set_ids(self._h, slot, _vp(ids), len(ids)))

    def set_max_match_id(self, id_, slot=0):
        self._check(lib().svo_set_max_match_id(self._h, slot, C.c_int64(id_)))

    # ---- stage 1 ----------------------------------------------------------------------------------------
    def set_input_format(self, channels):
        self._check(lib().svo_set_input_format(self._h, int(channels)))

    def set_rectify_maps(self, map_xy_left=None, map_a_left=None, map_xy_right=None, map_a_right=None):
        if map_xy_left is None:
            self._check(lib().svo_set_rectify_maps(self._h, None, None, None, None))
            return
        H, W = self.cfg.height, self.cfg.width
        mx = [np.ascontiguousarray(m, np.int16).reshape(H, W, 2) for m in (map_xy_left, map_xy_right)]
        ma = [np.ascontiguousarray(m, np.uint16).reshape(H, W) for m in (map_a_left, map_a_right)]
        self._check(lib().svo_set_rectify_maps(self._h, _vp(mx[0]), _vp(ma[0]), _vp(mx[1]), _vp(ma[1])))

    def prepare_rectify(self, left, right):
        left = np.ascontiguousarray(left, np.uint8)
        right = np.ascontiguousarray(right, np.uint8)
        ch = 1 if left.ndim == 2 else left.shape[2]
        H, W = self.cfg.height, self.cfg.width
        assert left.shape[:2] == (H, W) and right.shape == left.shape
        oL = np.zeros((H, W), np.uint8)
        oR = np.zeros((H, W), np.uint8)
        self._check(lib().svo_prepare_rectify(self._h, _vp(left), _vp(right), C.c_size_t(W * ch), ch, _vp(oL), _vp(oR)))
        return oL, oR

    def prepare_rectify_device(self, d_left, d_right, n, stride, image_pitch, channels, d_out_left, d_out_right):
        self._check(lib().svo_prepare_rectify_device(self._h, n, C.c_void_p(d_left), C.c_void_p(d_right), C.c_size_t(stride),
                                                     C.c_size_t(image_pitch), channels, C.c_void_p(d_out_left),
                                                     C.c_void_p(d_out_right)))

    def build_pyramid_device(self, d_left, d_right, n, stride, image_pitch):
        self._check(lib().svo_build_pyramid_device(self._h, n, C.c_void_p(d_left), C.c_void_p(d_right), C.c_size_t(stride),
                                                   C.c_size_t(image_pitch)))

    def set_concurrency(self, chunks):
        self._check(lib().svo_set_concurrency(self._h, int(chunks)))

    def enable_stage_timing(self, on=True):
        self._check(lib().svo_enable_stage_timing(self._h, int(on)))

    def pipe_peaks(self):
        p, l = C.c_double(0), C.c_double(0)
        self._check(lib().svo_measure_pipe_peaks(self._h, C.byref(p), C.byref(l)))
        return {"popc_per_s": p.value, "lop3_per_s": l.value}

    def stage_times(self):
        names = (C.c_char_p * 32)()
        ms = (C.c_float * 32)()
        n = self._check(lib().svo_get_stage_times(self._h, names, ms, 32))
        return {names[i].decode(): float(ms[i]) for i in range(n)}
