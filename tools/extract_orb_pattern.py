"""Recover the 512 sampling points of ORB's 31 x 31 rBRIEF test pattern from cv2 itself.

OpenCV (the reference's third-party dependency: opencv-python 4.13, modules/features2d/src/orb.cpp) keeps the
learned pattern as a private table; it is not in /root/reference and not exposed through the Python API.  The
descriptor stage the reference calls (feature_detectors.py:25-26 -> cv2.ORB.detectAndCompute) is a pure function of
(image, keypoints), so the table can be read back by probing `cv2.ORB.compute` with step images and a single
angle-0 keypoint: bit m of the descriptor is I(p[2m]) < I(p[2m+1]) on the 7 x 7 Gaussian-blurred image, and a blurred
step edge at offset c is a monotone ramp over a known window, so the set of c for which the bit is on gives both
coordinates along the step direction.  Three directions (x, y, x + y) and two polarities resolve every pair.

Writes slam_experiments_b200/orb_pattern31.npy ([512, 2] int8, (x, y) per point) and prints a checksum.  The parity
tests then compare whole descriptors against cv2 on textured images, which is what actually pins the table.
"""
import os
import sys

import cv2
import numpy as np

N, C0 = 160, 80          # probe image size, keypoint position
CR = 44                  # step offsets -CR .. CR


def describe(img):
    orb = cv2.ORB.create(nfeatures=10)
    kp = [cv2.KeyPoint(float(C0), float(C0), 31.0, 0.0, 1.0, 0)]
    kps, d = orb.compute(img, kp)
    assert len(kps) == 1
    return np.unpackbits(d[0], bitorder="little")       # bit m = byte m // 8, bit m % 8


def on_sets(coord_fn):
    ys, xs = np.mgrid[0:N, 0:N]
    t = coord_fn(xs - C0, ys - C0)
    rising = np.zeros((2 * CR + 1, 256), bool)
    falling = np.zeros((2 * CR + 1, 256), bool)
    for i, c in enumerate(range(-CR, CR + 1)):
        rising[i] = describe(np.where(t >= c, 255, 0).astype(np.uint8)) == 1
        falling[i] = describe(np.where(t < c, 255, 0).astype(np.uint8)) == 1
    return rising, falling


def decode(rising, falling, lo_margin, hi_margin):
    """Per pair: (a, b) coordinates along the direction, or None where a == b (the bit never fires)."""
    a = np.full(256, 1000)
    b = np.full(256, 1000)
    cs = np.arange(-CR, CR + 1)
    for m in range(256):
        r, f = cs[rising[:, m]], cs[falling[:, m]]
        assert not (len(r) and len(f)), m
        if len(r):                                  # a < b: on for c in [a - lo, b + hi]
            assert np.array_equal(r, np.arange(r[0], r[-1] + 1)), m
            a[m], b[m] = r[0] + lo_margin, r[-1] - hi_margin
        elif len(f):                                # a > b: on for c in [b - lo, a + hi]
            assert np.array_equal(f, np.arange(f[0], f[-1] + 1)), m
            b[m], a[m] = f[0] + lo_margin, f[-1] - hi_margin
    return a, b


def write_inc(pts):
    """The same table as a C initialiser for csrc/hm_orb.cu."""
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "slam_experiments_b200", "csrc", "hm_orb_pattern.inc")
    with open(inc, "w") as f:
        f.write("// 512 sampling points (x, y) of ORB's 31 x 31 rBRIEF pattern, as recovered from cv2 by tools/extract_orb_pattern.py\n")
        for i in range(0, 512, 8):
            f.write("    " + " ".join(f"{{{int(x)}, {int(y)}}}," for x, y in pts[i:i + 8]) + "\n")


def main():
    rx, fx = on_sets(lambda x, y: x)
    ry, fy = on_sets(lambda x, y: y)
    rs, fs = on_sets(lambda x, y: x + y)
    xa, xb = decode(rx, fx, 2, 3)                   # 7-tap blur: ramp over t in [-3, 2]
    ya, yb = decode(ry, fy, 2, 3)
    sa, sb = decode(rs, fs, 5, 6)                   # two 7-tap passes along the diagonal: t in [-6, 5]
    pts = np.zeros((512, 2), np.int64)
    for m in range(256):
        x0, x1, y0, y1 = xa[m], xb[m], ya[m], yb[m]
        if x0 == 1000:                              # equal x: from the diagonal and y
            assert y0 != 1000 and sa[m] != 1000, m
            x0, x1 = sa[m] - y0, sb[m] - y1
            assert x0 == x1, m
        if y0 == 1000:
            assert sa[m] != 1000, m
            y0, y1 = sa[m] - x0, sb[m] - x1
            assert y0 == y1, m
        if sa[m] != 1000:
            assert sa[m] == x0 + y0 and sb[m] == x1 + y1, (m, sa[m], sb[m], x0, y0, x1, y1)
        pts[2 * m] = (x0, y0)
        pts[2 * m + 1] = (x1, y1)
    assert np.abs(pts).max() <= 15, np.abs(pts).max()
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "slam_experiments_b200", "orb_pattern31.npy")
    np.save(out, pts.astype(np.int8))
    write_inc(pts)
    print(out, "first pairs:", pts[:8].tolist(), "checksum", int((pts * np.arange(1, 1025).reshape(512, 2)).sum()))


if __name__ == "__main__":
    sys.exit(main())
