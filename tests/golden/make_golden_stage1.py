#!/usr/bin/env python
"""Generate tests/golden/golden_stage1.npz from the REAL OpenCV (cv2) in this image: the two third-party calls
behind the reference's stage 1 (stage1_rectify.cpp:51-52 grayscale, :68-74 rectification).

    python tests/golden/make_golden_stage1.py
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def maps_for(W, H, side):
    """rectification maps as mrpt::vision::CStereoRectifyMap::setFromCamParams builds them:
    cv::initUndistortRectifyMap(K, D, R, P, size, CV_16SC2)."""
    K = np.array([[0.82 * W, 0, 0.5 * W + 3 * side], [0, 0.82 * W, 0.5 * H - 2 * side], [0, 0, 1.0]])
    D = np.array([-0.28 + 0.03 * side, 0.07, 0.0008, -0.0005, 0.0])
    R = cv2.Rodrigues(np.array([0.01, -0.02 + 0.015 * side, 0.005]))[0]
    P = np.array([[0.78 * W, 0, 0.5 * W, 0], [0, 0.78 * W, 0.5 * H, 0], [0, 0, 1.0, 0]])
    return cv2.initUndistortRectifyMap(K, D, R, P, (W, H), cv2.CV_16SC2)


def main():
    rng = np.random.default_rng(7)
    W, H = 192, 128
    out = {"cv2_version": np.array(cv2.__version__)}
    for side, name in enumerate(("L", "R")):
        bgr = cv2.GaussianBlur(rng.integers(0, 256, (H, W, 3), dtype=np.uint8), (5, 5), 1.2)
        bgr[:8, :8] = rng.integers(0, 256, (8, 8, 3), dtype=np.uint8)         # unsmoothed corner: extreme values
        gray = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
        mxy, ma = maps_for(W, H, side)
        rect = cv2.remap(gray, mxy, ma, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        out["bgr_" + name] = bgr
        out["gray_" + name] = gray
        out["map_xy_" + name] = mxy
        out["map_a_" + name] = ma
        out["rect_" + name] = rect
    np.savez_compressed(os.path.join(HERE, "golden_stage1.npz"), **out)
    print("wrote golden_stage1.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
