"""ORACLE (test infrastructure, not product code): CPU restatement of the DESCRIPTOR stage of OpenCV's ORB --
the step in front of the matcher (/root/reference/feature_detectors.py:25-26 ->
cv2.ORB.detectAndCompute; called from /root/reference/frontend.py:245-249).

The algorithm lives in a third-party dependency that is not part of /root/reference: opencv-python-headless 4.13.0
(modules/features2d/src/orb.cpp ORB_Impl::detectAndCompute / computeOrbDescriptors, imgproc resize INTER_LINEAR_EXACT,
imgproc GaussianBlur on a sub-matrix).  This file restates, in numpy integer arithmetic, what that code computes for
GIVEN keypoints (position, angle, octave as cv2's own detector produced them):

  1. gray image (BGR input: the 15-bit fixed-point luma cv2.cvtColor uses),
  2. the scale pyramid: level sizes from the float32 scale 1.2^level, each level resized from the PREVIOUS one with
     the bit-exact bilinear resampler (8.8 fixed-point coefficients, 16.16 accumulation, round to nearest),
  3. a 32-pixel BORDER_REFLECT_101 frame around every level, then a 7 x 7 sigma-2 Gaussian blur of the level's
     interior.  The level is a sub-matrix of one big pyramid buffer and the border mode is not ISOLATED, so cv2 does
     NOT take the fixed-point kernel (18 34 48 56 48 34 18) / 256 a stand-alone cv2.GaussianBlur call on a whole
     uint8 image uses; it runs its separable FLOAT filter (sepFilter2D, float32 kernel): a row pass uint8 -> float32
     and a column pass float32 -> uint8 (round half to even).  Float32 sums depend on the order and on fused
     multiply-adds, so the order is part of the specification (established against cv2.sepFilter2D of the
     opencv-python-headless 4.13.0 wheel, AVX2 dispatch, 0 differing pixels in 13.5 M; the parity test repeats it):
       row:    s = g0 p0, then s = fma(g_k, p_k, s) for k = 1 .. 6 in the 32-pixel vector body (x < 32 floor(w / 32));
               s = s + g_k p_k (two roundings) in the scalar tail,
       column: c = g3 r0, then c = fma(g_{3+k}, r_{+k} + r_{-k}, c) for k = 1 .. 3 (x < 4 floor(w / 4));
               c = c + g_{3+k} (r_{+k} + r_{-k}) in the tail.
     The frame keeps its un-blurred values,
  4. rBRIEF: 256 intensity comparisons at the learned 31 x 31 pattern rotated by the keypoint angle
     (x' = round(x cos - y sin), y' = round(x sin + y cos), float32, round-half-even), bit k of byte i from pair
     16 i + 2 k.

Parity pinned: tests/test_orb.py compares every function here against cv2 LIVE (cv2.resize, cv2.ORB.compute,
cv2.ORB.detectAndCompute) on synthetic textured images, all octaves, and tests/golden/orb_describe.npz holds the
recorded output of the unmodified reference's OrbFeatureDetector on one such image.  Only tests/, smoke() and bench.py's
cpu_baseline leg may import this file.
"""
from __future__ import annotations

import math
import os

import numpy as np

BORDER = 32                      # max(edgeThreshold 31, ceil(15 sqrt 2), HARRIS_BLOCK / 2) + 1
# cv2.getGaussianKernel(7, 2, CV_32F): exp(-x^2 / 8) normalised in double, cast to float32 (bit patterns)
GAUSS7 = np.array([0x3d8fafb1, 0x3e06387e, 0x3e434a39, 0x3e5d4ae0, 0x3e434a39, 0x3e06387e, 0x3d8fafb1], np.uint32).view(np.float32)
_LD = np.longdouble              # 64-bit mantissa: a float32 product plus a float32 addend is exact, so one rounding = fma


def _fma(a, b, c):
    return (a.astype(_LD) * b.astype(_LD) + c.astype(_LD)).astype(np.float32)


def _mul(a, b):
    return (a * b).astype(np.float32)
_PATTERN = None


def pattern31() -> np.ndarray:
    """[512, 2] (x, y): recovered from cv2 by tools/extract_orb_pattern.py."""
    global _PATTERN
    if _PATTERN is None:
        here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        _PATTERN = np.load(os.path.join(here, "slam_experiments_b200", "orb_pattern31.npy")).astype(np.int64)
    return _PATTERN


def to_gray(img: np.ndarray) -> np.ndarray:
    """cv2.cvtColor(BGR2GRAY) for uint8: (B 3735 + G 19235 + R 9798 + 2^14) >> 15."""
    if img.ndim == 2:
        return np.ascontiguousarray(img)
    b, g, r = (img[..., i].astype(np.int64) for i in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


def level_scale(level: int) -> np.float32:
    return np.float32(math.pow(float(np.float32(1.2)), float(level)))


def level_size(rows: int, cols: int, level: int) -> tuple[int, int]:
    inv = np.float32(1.0) / level_scale(level)
    # cvRound(int * float): float32 product, round half to even
    return int(np.rint(np.float32(rows) * inv)), int(np.rint(np.float32(cols) * inv))


def _linear_coeffs(src: int, dst: int):
    """offset and 8.8 coefficients of resize INTER_LINEAR_EXACT along one axis."""
    inv_scale = float(dst) / float(src)
    scale = 1.0 / inv_scale
    ofs = np.zeros(dst, np.int64)
    c0 = np.zeros(dst, np.int64)
    c1 = np.zeros(dst, np.int64)
    for d in range(dst):
        fval = scale * (d + 0.5) - 0.5
        ival = math.floor(fval)
        if ival >= 0 and src > 1:
            if ival < src - 1:
                ofs[d] = ival
                c1[d] = int(np.rint((fval - ival) * 256.0))
                c0[d] = 256 - c1[d]
            else:                         # right / bottom edge: replicate the last sample
                ofs[d] = src - 1
                c0[d], c1[d] = 256, 0
        else:                             # left / top edge: replicate the first sample
            ofs[d] = 0
            c0[d], c1[d] = 256, 0
    return ofs, c0, c1


def resize_linear_exact(src: np.ndarray, rows: int, cols: int) -> np.ndarray:
    xo, x0, x1 = _linear_coeffs(src.shape[1], cols)
    yo, y0, y1 = _linear_coeffs(src.shape[0], rows)
    s = src.astype(np.int64)
    xn = np.minimum(xo + 1, src.shape[1] - 1)
    yn = np.minimum(yo + 1, src.shape[0] - 1)
    h = s[:, xo] * x0[None, :] + s[:, xn] * x1[None, :]                    # 8.8
    v = h[yo, :] * y0[:, None] + h[yn, :] * y1[:, None]                    # 16.16
    return ((v + (1 << 15)) >> 16).astype(np.uint8)


def reflect101_frame(img: np.ndarray, border: int = BORDER) -> np.ndarray:
    return np.pad(img, border, mode="reflect")


def blur_float7(src_with_halo: np.ndarray, h: int, w: int) -> np.ndarray:
    """The 7 x 7 float filter on an image that carries a 3-pixel halo: [h + 6, w + 6] uint8 -> [h, w] uint8."""
    f = src_with_halo.astype(np.float32)
    taps = [f[:, k:k + w] for k in range(7)]
    K = [np.float32(g) for g in GAUSS7]
    rows = _mul(K[0], taps[0])
    for k in range(1, 7):
        rows = _fma(np.full_like(rows, K[k]), taps[k], rows)
    t0 = (w // 32) * 32
    if t0 < w:                                       # scalar tail of the row filter: multiply, then add
        tail = _mul(K[0], taps[0][:, t0:])
        for k in range(1, 7):
            tail = tail + _mul(K[k], taps[k][:, t0:])
        rows[:, t0:] = tail
    R = [rows[k:k + h, :] for k in range(7)]
    col = _mul(K[3], R[3])
    for k in range(1, 4):
        col = _fma(np.full_like(col, K[3 + k]), R[3 + k] + R[3 - k], col)
    c0 = (w // 4) * 4
    if c0 < w:
        tail = _mul(K[3], R[3][:, c0:])
        for k in range(1, 4):
            tail = tail + _mul(K[3 + k], R[3 + k][:, c0:] + R[3 - k][:, c0:])
        col[:, c0:] = tail
    return np.clip(np.rint(col), 0, 255).astype(np.uint8)


def blur_interior(framed: np.ndarray, border: int = BORDER) -> np.ndarray:
    """Blur of the interior of a framed level, reading the (un-blurred) frame; the frame is kept."""
    h, w = framed.shape[0] - 2 * border, framed.shape[1] - 2 * border
    out = framed.copy()
    out[border:border + h, border:border + w] = blur_float7(framed[border - 3:border + h + 3, border - 3:border + w + 3], h, w)
    return out


def build_pyramid(gray: np.ndarray, n_levels: int) -> list[np.ndarray]:
    """Framed, blurred levels 0 .. n_levels - 1 (what computeOrbDescriptors samples)."""
    levels, prev = [], gray
    for lv in range(n_levels):
        if lv > 0:
            r, c = level_size(gray.shape[0], gray.shape[1], lv)
            prev = resize_linear_exact(prev, r, c)
        levels.append(blur_interior(reflect101_frame(prev)))
    return levels


def describe(img: np.ndarray, pts: np.ndarray, angles: np.ndarray, octaves: np.ndarray) -> np.ndarray:
    """rBRIEF descriptors [n, 32] uint8 for keypoints (x, y) float32 in level-0 coordinates, angle in degrees."""
    gray = to_gray(img)
    n = len(pts)
    out = np.zeros((n, 32), np.uint8)
    if n == 0:
        return out
    levels = build_pyramid(gray, int(octaves.max()) + 1)
    pat = pattern31().astype(np.float32)
    px, py = pat[:, 0], pat[:, 1]
    for j in range(n):
        lv = int(octaves[j])
        inv = np.float32(1.0) / level_scale(lv)
        cx = int(np.rint(np.float32(pts[j, 0]) * inv)) + BORDER
        cy = int(np.rint(np.float32(pts[j, 1]) * inv)) + BORDER
        ang = np.float32(angles[j]) * np.float32(math.pi / np.float32(180.0))
        a, b = np.float32(math.cos(float(ang))), np.float32(math.sin(float(ang)))
        ix = np.rint(px * a - py * b).astype(np.int64)        # float32 products and difference, half-to-even
        iy = np.rint(px * b + py * a).astype(np.int64)
        v = levels[lv][cy + iy, cx + ix].astype(np.int64)
        bits = (v[0::2] < v[1::2]).astype(np.uint8)
        out[j] = np.packbits(bits, bitorder="little")
    return out
