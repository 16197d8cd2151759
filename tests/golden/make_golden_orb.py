"""Real-ORB fixtures beyond the bundled pair's 2000 features (build container only: needs /root/reference).

    python tests/golden/make_golden_orb.py

1. ``orb5k_pair.npz`` -- `1.png` / `2.png` upscaled 4x (2560 x 1920, SURVEY.md E9: the bundled 640 x 480 images
   cannot supply more than ~4.6 k features), ORB through the reference's own ``OrbFeatureDetector``
   (`/root/reference/feature_detectors.py:18-26`) with n_features=5000, matched by the reference's
   ``BruteForceFeatureMatcher`` (`feature_matchers.py:32-44`) and the cv2 calls of the north-star pipeline; same
   record layout as make_golden.py.
2. ``c2_sequence_orb2000.npz`` -- the C2 input of SURVEY.md 8(d): 100 frames (`/root/reference/euroc.py:40`) of
   752 x 480 grayscale made by resizing `1.png` and warping it along a smooth seeded homography trajectory plus
   Gaussian noise (sigma 2), ORB n_features=2000 per frame.  Stores descriptors [100, 2000, 32] uint8, keypoint
   positions [100, 2000, 2] float32, sizes / angles / octaves (inputs of a descriptor-extraction kernel), and the
   reference matcher's output for three of the 99 consecutive (last -> current) problems.
"""
from __future__ import annotations

import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, arr, record, save  # noqa: E402  (imports the reference's modules unmodified)
from feature_detectors import OrbFeatureDetector  # noqa: E402
from feature_matchers import BruteForceFeatureMatcher  # noqa: E402

SEED = 752480


def orb5k():
    det = OrbFeatureDetector(n_features=5000)
    descs = []
    for name in ("1.png", "2.png"):
        img = cv2.imread(os.path.join(REF, name), flags=cv2.IMREAD_COLOR)
        big = cv2.resize(img, (img.shape[1] * 4, img.shape[0] * 4), interpolation=cv2.INTER_CUBIC)
        _, d = det.detect_and_compute(big, None)
        descs.append(np.ascontiguousarray(d))
    # frontend.py:185-187: train = last frame (1.png), query = current frame (2.png)
    save("orb5k_pair.npz", record(descs[1], descs[0]))


def c2_sequence(n_frames=100, rows=2000):
    rng = np.random.default_rng(SEED)
    img = cv2.imread(os.path.join(REF, "1.png"), flags=cv2.IMREAD_GRAYSCALE)
    base = cv2.resize(img, (752, 480), interpolation=cv2.INTER_LINEAR)
    det = OrbFeatureDetector(n_features=rows)
    # smooth trajectory: slow drift + rotation + mild perspective, a few pixels per frame
    phase = rng.uniform(0, 2 * np.pi, 4)
    desc = np.zeros((n_frames, rows, 32), np.uint8)
    pts = np.zeros((n_frames, rows, 2), np.float32)
    meta = np.zeros((n_frames, rows, 3), np.float32)          # size, angle, octave
    counts = np.zeros(n_frames, np.int32)
    for i in range(n_frames):
        s = i / (n_frames - 1)
        ang = np.deg2rad(6.0 * np.sin(2 * np.pi * s + phase[0]))
        tx = 40.0 * np.sin(2 * np.pi * s * 0.7 + phase[1]) + 30.0 * s
        ty = 25.0 * np.sin(2 * np.pi * s * 0.9 + phase[2])
        sc = 1.0 + 0.06 * np.sin(2 * np.pi * s * 0.5 + phase[3])
        c, sn = np.cos(ang) * sc, np.sin(ang) * sc
        cx, cy = 376.0, 240.0
        H = np.array([[c, -sn, cx - c * cx + sn * cy + tx],
                      [sn, c, cy - sn * cx - c * cy + ty],
                      [1e-5 * np.sin(2 * np.pi * s), 8e-6 * np.cos(2 * np.pi * s), 1.0]])
        frame = cv2.warpPerspective(base, H, (752, 480), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT101)
        noisy = np.clip(frame.astype(np.float32) + rng.normal(0.0, 2.0, frame.shape), 0, 255).astype(np.uint8)
        kps, d = det.detect_and_compute(noisy, None)
        n = min(rows, len(kps))
        counts[i] = n
        desc[i, :n] = d[:n]
        pts[i, :n] = np.array([k.pt for k in kps[:n]], np.float32)
        meta[i, :n] = np.array([(k.size, k.angle, k.octave) for k in kps[:n]], np.float32)
    out = {"descriptors": desc, "points": pts, "meta": meta, "counts": counts}
    ref = BruteForceFeatureMatcher(norm_type=cv2.NORM_HAMMING)
    for i in (0, 37, 98):                                      # frontend.py:181-187: match(desc_last, desc_cur)
        out[f"ref_match_{i}"] = arr(ref.match(desc[i, :counts[i]], desc[i + 1, :counts[i + 1]]))
    save("c2_sequence_orb2000.npz", out)
    print("features per frame: min", counts.min(), "max", counts.max())


if __name__ == "__main__":
    orb5k()
    c2_sequence()
