"""Inputs of the Gauss-Newton known-answer tests."""
import numpy as np

from oracle import oracle as orc


def kat_inputs(noise, xr=3.0, yr=2.0):
    rng = np.random.default_rng(12345)
    N = 200
    X = rng.uniform(-xr, xr, N)
    Y = rng.uniform(-yr, yr, N)
    Z = rng.uniform(2, 10, N)
    fx = fy = 500.0
    cx, cy, B = 320.0, 240.0, 0.12
    w = np.array([0.01, -0.02, 0.005])
    t = np.array([0.05, -0.01, 0.10])
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    R = np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K
    P = np.stack([X, Y, Z], 1)
    Pc = P @ R.T + t

    def proj(P):
        l = np.stack([fx * P[:, 0] / P[:, 2] + cx, fy * P[:, 1] / P[:, 2] + cy], 1)
        r = np.stack([fx * (P[:, 0] - B) / P[:, 2] + cx, fy * P[:, 1] / P[:, 2] + cy], 1)
        return l, r
    pl, pr = proj(P)
    cl, cr = proj(Pc)
    if noise:
        pl = pl + rng.normal(0, 0.3, (N, 2))
        pr = pr + rng.normal(0, 0.3, (N, 2))
        cl = cl + rng.normal(0, 0.3, (N, 2))
        cr = cr + rng.normal(0, 0.3, (N, 2))
        cl[:20] += rng.uniform(-40, 40, (20, 2))
    arrs = [v.astype(np.float32) for v in (pl, pr, cl, cr)]
    return arrs, orc.make_camera(fx, fy, cx, cy, B), N
