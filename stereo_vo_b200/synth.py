"""Seeded synthetic stereo sequences (numpy only) for tests and bench.py.

Scene: three fronto-parallel textured planes (blurred uniform noise, which gives ~30 FAST(20)
corners per 1000 px so the keypoint cap of the reference's stage 2 is always active) seen by a
rectified pinhole stereo rig that moves with constant velocity.  Plane depths are chosen so
the frame-0 disparities are integers.  Rendering is an exact per-plane homography with
bilinear sampling; the generator runs on the host, outside every timed region.

Configurations follow BASELINE.json / SURVEY.md section 8(d):
  c1: 320x240,  orb_nfeats=100  (cap 200),  fx=250, baseline 0.12
  c2: 640x480,  orb_nfeats=500  (cap 1000), fx=500, baseline 0.12
  c3: 1242x375, orb_nfeats=1000 (cap 2000), fx=718, baseline 0.54
"""
import numpy as np

CONFIGS = {
    "c1": dict(W=320, H=240, fx=250.0, baseline=0.12, orb_nfeats=100),
    "c2": dict(W=640, H=480, fx=500.0, baseline=0.12, orb_nfeats=500),
    "c3": dict(W=1242, H=375, fx=718.0, baseline=0.54, orb_nfeats=1000),
}


def _gauss_blur5(img, sigma=1.2):
    x = np.arange(-2, 3, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    k /= k.sum()
    p = np.pad(img.astype(np.float64), 2, mode="reflect")
    H, W = img.shape
    r = sum(k[i] * p[:, i:i + W] for i in range(5))
    c = sum(k[i] * r[i:i + H, :] for i in range(5))
    return np.clip(np.rint(c), 0, 255).astype(np.uint8)


def make_texture(rng, h, w):
    return _gauss_blur5(rng.integers(0, 256, (h, w), dtype=np.uint8))


class Scene:
    def __init__(self, seed, W, H, fx, baseline, depths_disp=(15, 8, 4)):
        rng = np.random.default_rng(seed)
        self.W, self.H, self.fx, self.fy = W, H, float(fx), float(fx)
        self.cx, self.cy = W / 2.0, H / 2.0
        self.baseline = float(baseline)
        # depth of plane k such that the frame-0 disparity fx*B/Z is the integer depths_disp[k]
        self.depths = [self.fx * self.baseline / d for d in depths_disp]
        th, tw = 2 * H, 2 * W
        self.tex = [make_texture(rng, th, tw) for _ in self.depths]
        # plane 0 (nearest) covers the left 40 % of texture space, plane 1 the right 35 %, plane 2 everything
        self.extent = [(-np.inf, 0.4 * W, -np.inf, np.inf), (0.65 * W, np.inf, -np.inf, np.inf),
                       (-np.inf, np.inf, -np.inf, np.inf)]

    def camera(self):
        return dict(fx=self.fx, fy=self.fy, cx=self.cx, cy=self.cy, baseline=self.baseline)

    def render(self, R, t, right=False):
        """Image of the camera whose pose maps world -> camera as X_c = R X_w + t."""
        W, H = self.W, self.H
        u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
        d = np.stack([(u - self.cx) / self.fx, (v - self.cy) / self.fy, np.ones_like(u)], -1)  # camera rays
        tt = np.asarray(t, np.float64).copy()
        if right:
            tt = tt - np.array([self.baseline, 0, 0])  # right camera: X_r = X_l - (B,0,0)
        dw = d @ R                      # R^T d  (rows)
        ow = -(R.T @ tt)                # camera centre in world
        out = np.zeros((H, W), np.uint8)
        done = np.zeros((H, W), bool)
        for Z, tex, (x0, x1, y0, y1) in zip(self.depths, self.tex, self.extent):
            s = (Z - ow[2]) / dw[..., 2]
            Xw = ow[0] + s * dw[..., 0]
            Yw = ow[1] + s * dw[..., 1]
            tx = Xw * (self.fx / Z) + self.cx
            ty = Yw * (self.fy / Z) + self.cy
            ok = (~done) & (s > 0) & (tx >= x0) & (tx < x1) & (ty >= y0) & (ty < y1)
            th, tw = tex.shape
            txo = tx + tw / 4.0
            tyo = ty + th / 4.0
            x_0 = np.floor(txo)
            y_0 = np.floor(tyo)
            fx_ = txo - x_0
            fy_ = tyo - y_0
            xi = x_0.astype(np.int64) % tw
            yi = y_0.astype(np.int64) % th
            xi1 = (xi + 1) % tw
            yi1 = (yi + 1) % th
            val = (tex[yi, xi] * (1 - fx_) * (1 - fy_) + tex[yi, xi1] * fx_ * (1 - fy_)
                   + tex[yi1, xi] * (1 - fx_) * fy_ + tex[yi1, xi1] * fx_ * fy_)
            out[ok] = np.clip(np.rint(val[ok]), 0, 255).astype(np.uint8)
            done |= ok
        return out


def yaw_rot(a):
    """Rotation about the camera's y (down) axis."""
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def make_sequence(config="c2", seed=0, n_frames=4, dx=0.05, dyaw_deg=0.2, dy=0.0):
    """Returns (frames, cam) with frames = list of (left, right) uint8 images."""
    cfg = CONFIGS[config]
    sc = Scene(seed, cfg["W"], cfg["H"], cfg["fx"], cfg["baseline"])
    frames = []
    for f in range(n_frames):
        # camera moves by +dx along x and yaws by dyaw per frame: world->camera is the inverse
        Rwc = yaw_rot(np.deg2rad(dyaw_deg) * f)          # camera orientation in world
        cw = np.array([dx * f, dy * f, 0.0])             # camera centre in world
        R = Rwc.T
        t = -R @ cw
        frames.append((sc.render(R, t, False), sc.render(R, t, True)))
    return frames, sc.camera()


def make_sequence_two_motions(config="c2", seed=0, n_frames=4, frac=0.38):
    """A sequence with an independently moving part: the right `frac` of every image (left and right camera alike) shows the
    same scene under a second rigid motion.  Features inside it are detected, stereo-matched and tracked like any other, but
    they contradict the dominant motion: about `frac` of the tracks are outliers for the two F-matrix RANSAC filters and for
    the robust Gauss-Newton, so the adaptive RANSAC loop has to run many iterations instead of stopping after the first
    model (the plain sequences are ~99 % inliers)."""
    a, cam = make_sequence(config, seed=seed, n_frames=n_frames)
    # the second motion goes DOWN as well: its tracks are off the epipolar lines of the (horizontal) dominant motion
    b, _ = make_sequence(config, seed=seed, n_frames=n_frames, dx=-0.02, dyaw_deg=-0.1, dy=0.05)
    W = a[0][0].shape[1]
    x0 = int(round(W * (1.0 - frac)))
    out = []
    for (La, Ra), (Lb, Rb) in zip(a, b):
        L, R = La.copy(), Ra.copy()
        L[:, x0:] = Lb[:, x0:]
        R[:, x0:] = Rb[:, x0:]
        out.append((L, R))
    return out, cam
