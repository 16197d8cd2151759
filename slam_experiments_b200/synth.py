"""Seeded synthetic descriptor workloads, shared by tests and bench.py.

Distributions: (U) i.i.d. uniform bytes -- top-1 distances ~80-105 and ~20% of rows with a
top1 == top2 tie, which stresses tie-breaking; (M) "matchable": 60% of the queries are a train
row with each bit flipped w.p. 0.1, the rest uniform -- makes the ratio test and the mutual
check non-trivial (real ORB bit density is 0.54).
"""
from __future__ import annotations

import numpy as np

DESC_BYTES = 32


def uniform(n: int, seed: int) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, 256, (n, DESC_BYTES), dtype=np.uint8)


def matchable_queries(train: np.ndarray, nq: int, seed: int, flip_p: float = 0.1, frac: float = 0.6) -> np.ndarray:
    rng = np.random.default_rng(seed)
    nt = train.shape[0]
    src = rng.integers(0, nt, nq)
    q = train[src].copy()
    noisy = rng.random(nq) < frac
    k = int(noisy.sum())
    if k:
        flips = np.packbits(rng.random((k, 256)) < flip_p, axis=1)
        q[noisy] ^= flips
    if nq - k:
        q[~noisy] = rng.integers(0, 256, (nq - k, DESC_BYTES), dtype=np.uint8)
    return q


def sweep(n: int, dist: str = "U"):
    """C3: N x N, seed 1000 + log2(N)."""
    seed = 1000 + int(np.log2(max(n, 1)))
    t = uniform(n, seed)
    q = uniform(n, seed + 100) if dist == "U" else matchable_queries(t, n, seed + 100)
    return q, t


def keyframe_database(n_keyframes: int = 4096, rows: int = 2000, nq: int = 2000, seed: int = 4096):
    """C4: ``n_keyframes`` x ``rows`` train descriptors (one array [n_keyframes*rows, 32]) and a
    matchable query whose true matches are scattered over the keyframes."""
    rng = np.random.default_rng(seed)
    train = rng.integers(0, 256, (n_keyframes * rows, DESC_BYTES), dtype=np.uint8)
    q = matchable_queries(train, nq, seed + 1)
    return q, train


def frame_sequence(n_frames: int = 100, rows: int = 2000, seed: int = 752480, flip_p: float = 0.05) -> np.ndarray:
    """C2-shaped: consecutive frames share most features (each frame = previous frame with a few
    bits flipped, 30% of the rows replaced and the order shuffled).  [n_frames, rows, 32]."""
    rng = np.random.default_rng(seed)
    frames = np.empty((n_frames, rows, DESC_BYTES), dtype=np.uint8)
    frames[0] = rng.integers(0, 256, (rows, DESC_BYTES), dtype=np.uint8)
    for i in range(1, n_frames):
        f = frames[i - 1].copy()
        f ^= np.packbits(rng.random((rows, 256)) < flip_p, axis=1)
        new = rng.random(rows) < 0.3
        f[new] = rng.integers(0, 256, (int(new.sum()), DESC_BYTES), dtype=np.uint8)
        frames[i] = f[rng.permutation(rows)]
    return frames


def local_window(n_frames: int = 32, rows: int = 10000, seed: int = 3200):
    """C5: 32 train frames x 10,000 rows and one 10,000-row matchable query per frame."""
    rng = np.random.default_rng(seed)
    train = rng.integers(0, 256, (n_frames, rows, DESC_BYTES), dtype=np.uint8)
    query = np.stack([matchable_queries(train[i], rows, seed + 1 + i) for i in range(n_frames)])
    return query, train


def euroc_shaped_sequence() -> np.ndarray:
    """C2 on the SURVEY.md 8(d) input: 100 frames of 752 x 480 (`/root/reference/1.png` resized, warped along a smooth
    seeded homography trajectory, Gaussian noise sigma 2), 2000 REAL ORB descriptors per frame from the reference's
    own detector.  The images are not available on the GPU box, so the descriptors are a committed fixture
    (tests/golden/c2_sequence_orb2000.npz, written by tests/golden/make_golden_orb.py).  [100, 2000, 32] uint8."""
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                        "c2_sequence_orb2000.npz")
    with np.load(path) as z:
        return np.ascontiguousarray(z["descriptors"])


def textured_image(h: int = 480, w: int = 640, seed: int = 7, channels: int = 1) -> np.ndarray:
    """A seeded corner-rich test image for the ORB descriptor stage (8f rank 3): smooth random texture (low-resolution
    noise, bilinearly upsampled) with filled rectangles of random gray levels on top.  cv2's ORB finds several thousand
    keypoints on all 8 pyramid levels.  uint8 ``[h, w]`` or ``[h, w, 3]`` (BGR planes with different textures)."""
    rng = np.random.default_rng(seed)
    planes = []
    for _ in range(channels):
        gh, gw = h // 6 + 2, w // 6 + 2
        g = rng.integers(0, 256, (gh, gw)).astype(np.float64)
        ys, xs = np.arange(h) / 6.0, np.arange(w) / 6.0
        y0, x0 = ys.astype(int), xs.astype(int)
        fy, fx = (ys - y0)[:, None], (xs - x0)[None, :]
        img = (g[y0][:, x0] * (1 - fy) * (1 - fx) + g[y0][:, x0 + 1] * (1 - fy) * fx +
               g[y0 + 1][:, x0] * fy * (1 - fx) + g[y0 + 1][:, x0 + 1] * fy * fx)
        img = np.rint(img).astype(np.uint8)
        for _ in range(h * w // 5000):
            y, x = int(rng.integers(0, h - 8)), int(rng.integers(0, w - 8))
            img[y:y + int(rng.integers(6, 40)), x:x + int(rng.integers(6, 40))] = int(rng.integers(0, 256))
        planes.append(img)
    return planes[0] if channels == 1 else np.ascontiguousarray(np.stack(planes, axis=2))
