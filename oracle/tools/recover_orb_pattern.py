#!/usr/bin/env python
"""Recover the 256 BRIEF test pairs that cv2.ORB::compute applies (TEST INFRASTRUCTURE).

The reference calls `cv::ORB::create()->compute(img, kps, desc)` at
libstereo-odometry/src/stage2_detect.cpp:512.  The learned 31x31 test pattern lives
inside OpenCV (not in the reference tree, not vendored), so we recover it from the
OpenCV binary in this image (cv2 4.13.0) as a black box:

  * a constant image gives an all-zero descriptor (bit = blur[a] < blur[b] is false);
  * one bright impulse at offset o from a keypoint raises the blurred value in the 7x7
    square around o, so bit i flips to 1 exactly where o is "closer" to b_i than to a_i;
  * one dark impulse does the same for a_i.

Every keypoint of a sparse grid gets its own impulse offset, so two `compute` calls
(bright, dark) give the full flip maps.  Candidates for (a_i, b_i) are taken from the
bounding boxes of the flip maps and disambiguated by simulating the flip maps with the
exact float32 blur (oracle/pyoracle.blur7_exact).  The result is validated against
cv2 on random images (0 differing bits required) before anything is written.

Outputs (generated artefacts, committed):
  oracle/orb_pattern_31.txt               256 lines "ax ay bx by"
  oracle/orb_pattern.inc                  C initialiser used by the C++ oracle
  stereo_vo_b200/csrc/orb_pattern.inc     identical table for the CUDA kernels

Run:  python oracle/tools/recover_orb_pattern.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import blur7_exact  # noqa: E402

R = 19          # probe offsets in [-R, R]^2
PITCH = 64      # grid pitch between probe keypoints
BASE = 100


def _probe(impulse_value, angle):
    offs = [(ox, oy) for oy in range(-R, R + 1) for ox in range(-R, R + 1)]
    side = int(np.ceil(np.sqrt(len(offs))))
    size = PITCH * (side + 1)
    img = np.full((size, size), BASE, np.uint8)
    kps = []
    for k, (ox, oy) in enumerate(offs):
        cx = PITCH * (1 + k % side)
        cy = PITCH * (1 + k // side)
        img[cy + oy, cx + ox] = impulse_value
        kps.append(cv2.KeyPoint(float(cx), float(cy), 7.0, float(angle), 10.0, 0, -1))
    kps2, desc = cv2.ORB_create().compute(img, kps)
    assert len(kps2) == len(kps), "probe keypoints were dropped by the border filter"
    bits = np.unpackbits(desc, axis=1, bitorder="little")  # (n_offs, 256), LSB-first per byte
    m = np.zeros((256, 2 * R + 1, 2 * R + 1), bool)       # [bit, oy+R, ox+R]
    for k, (ox, oy) in enumerate(offs):
        m[:, oy + R, ox + R] = bits[k] != 0
    return m


def _impulse_response(impulse_value):
    """Blurred value at displacement d (|d|<=3) from a single impulse on BASE."""
    im = np.full((15, 15), BASE, np.uint8)
    im[7, 7] = impulse_value
    return blur7_exact(im)  # value at (7+dy, 7+dx) = response at displacement (dx,dy) from impulse


def _predict(resp, a, b):
    """flip[oy,ox] = blur(p+a) < blur(p+b) when the impulse sits at p+o."""
    m = np.zeros((2 * R + 1, 2 * R + 1), bool)
    for oy in range(-R, R + 1):
        for ox in range(-R, R + 1):
            def val(pt):
                dx, dy = pt[0] - ox, pt[1] - oy
                if abs(dx) > 3 or abs(dy) > 3:
                    return BASE
                return int(resp[7 + dy, 7 + dx])
            m[oy + R, ox + R] = val(a) < val(b)
    return m


def _bbox_candidates(mask):
    ys, xs = np.nonzero(mask)
    assert len(xs) > 0, "a bit never flipped: enlarge R"
    xs = xs - R
    ys = ys - R
    return [(x, y) for x in range(xs.max() - 3, xs.min() + 3 + 1)
            for y in range(ys.max() - 3, ys.min() + 3 + 1)]


def recover(angle):
    bright = _probe(255, angle)
    dark = _probe(0, angle)
    resp_b = _impulse_response(255)
    resp_d = _impulse_response(0)
    pattern = np.zeros((256, 4), np.int8)
    for i in range(256):
        cb = _bbox_candidates(bright[i])   # candidates for b_i
        ca = _bbox_candidates(dark[i])     # candidates for a_i
        good = []
        for a in ca:
            for b in cb:
                if a == b:
                    continue
                if (np.array_equal(_predict(resp_b, a, b), bright[i])
                        and np.array_equal(_predict(resp_d, a, b), dark[i])):
                    good.append((a, b))
        assert len(good) == 1, f"bit {i}: {len(good)} consistent candidates {good[:4]}"
        (ax, ay), (bx, by) = good[0]
        pattern[i] = (ax, ay, bx, by)
    return pattern


def validate(pattern, n_images=6):
    from oracle.pyoracle import brief256
    rng = np.random.default_rng(2024)
    bad = 0
    total = 0
    for t in range(n_images):
        img = rng.integers(0, 256, (240, 320), dtype=np.uint8)
        img = cv2.GaussianBlur(img, (5, 5), 1.2)
        kps = cv2.FastFeatureDetector_create(20).detect(img)
        kps, desc = cv2.ORB_create().compute(img, kps)
        xy = np.array([[int(k.pt[0]), int(k.pt[1])] for k in kps], np.int32)
        mine = brief256(blur7_exact(img), xy, pattern)
        bad += int(np.unpackbits(mine ^ desc).sum())
        total += desc.size * 8
    return bad, total


def main():
    p0 = recover(0.0)
    pm1 = recover(-1.0)
    assert np.array_equal(p0, pm1), "pattern at angle=-1 differs from angle=0"
    bad, total = validate(p0)
    print(f"validation vs cv2.ORB.compute: {bad} / {total} bits differ")
    assert bad == 0
    txt = os.path.join(ROOT, "oracle", "orb_pattern_31.txt")
    with open(txt, "w") as f:
        for r in p0:
            f.write("%d %d %d %d\n" % tuple(int(v) for v in r))
    body = ("/* GENERATED by oracle/tools/recover_orb_pattern.py from cv2 %s -- do not edit.\n"
            " * 256 BRIEF test pairs {ax, ay, bx, by}: bit i = blur[y+ay][x+ax] < blur[y+by][x+bx].\n"
            " * Replaces the table inside cv::ORB::compute called at\n"
            " * libstereo-odometry/src/stage2_detect.cpp:512 (angle = -1 deg rounds to these offsets). */\n"
            % cv2.__version__)
    body += ",\n".join("{%d,%d,%d,%d}" % tuple(int(v) for v in r) for r in p0) + "\n"
    for rel in ("oracle/orb_pattern.inc", "stereo_vo_b200/csrc/orb_pattern.inc"):
        with open(os.path.join(ROOT, rel), "w") as f:
            f.write(body)
    print("first pairs:", p0[:4].tolist(), "last:", p0[-1].tolist())
    print("max |coord| =", int(np.abs(p0).max()))


if __name__ == "__main__":
    main()
