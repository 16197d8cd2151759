"""Generate the golden fixtures for the Hamming matching path.

Run in the build container (needs `/root/reference`, which does NOT exist on
the GPU box):

    python tests/golden/make_golden.py

It imports the reference's own modules unmodified --
`/root/reference/feature_matchers.py` (``BruteForceFeatureMatcher``, lines
32-44) and `/root/reference/feature_detectors.py` (``OrbFeatureDetector``,
lines 18-26) -- runs them on the bundled `1.png`/`2.png` and on seeded
synthetic descriptors, and stores inputs + outputs as small ``.npz`` files.
The cv2.BFMatcher calls that the north-star pipeline adds (knnMatch k=2,
crossCheck, the train-collection API) are recorded
next to them.  cv2 version used is stored in each file.
"""
from __future__ import annotations

import os
import sys

import cv2
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)

from feature_matchers import BruteForceFeatureMatcher  # noqa: E402  (the reference, as-is)
from feature_detectors import OrbFeatureDetector  # noqa: E402


def arr(matches):
    return np.array([(m.queryIdx, m.trainIdx, m.imgIdx, int(m.distance)) for m in matches],
                    dtype=np.int32).reshape(-1, 4)


def knn_arr(rows, k):
    """[Nq, k, 4] with -1 padding for short rows."""
    out = np.full((len(rows), k, 4), -1, dtype=np.int32)
    for i, r in enumerate(rows):
        for j, m in enumerate(r):
            out[i, j] = (m.queryIdx, m.trainIdx, m.imgIdx, int(m.distance))
    return out


def record(query, train, ratios=(0.7, 0.75, 0.8), thresholds=(30.0, 64.0)):
    """All reference / cv2 outputs for one (query, train) problem."""
    ref = BruteForceFeatureMatcher(norm_type=cv2.NORM_HAMMING)
    out = {"query": query, "train": train}
    # the reference's own call (feature_matchers.py:36-44): match(source=train, query)
    out["ref_match"] = arr(ref.match(train, query))
    for th in thresholds:
        out[f"ref_match_thr{int(th)}"] = arr(ref.match(train, query, dist_threshold=th))
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    out["knn1"] = knn_arr(bf.knnMatch(query, train, k=1), 1)
    rows = bf.knnMatch(query, train, k=2)
    out["knn2"] = knn_arr(rows, 2)
    cc = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(query, train)
    out["cross"] = arr(cc)
    ok = {(m.queryIdx, m.trainIdx) for m in cc}
    for r in ratios:
        good = [x[0] for x in rows if len(x) == 2 and x[0].distance < r * x[1].distance]
        out[f"ratio{int(round(r * 100))}"] = arr(good)
        out[f"pipe{int(round(r * 100))}"] = arr([m for m in good if (m.queryIdx, m.trainIdx) in ok])
    return out


def save(name, d):
    d = dict(d)
    d["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(HERE, name), **d)
    print(name, {k: getattr(v, "shape", None) for k, v in d.items() if k != "cv2_version"})


def main():
    # ---- C1: the bundled pair, ORB exactly as the reference builds it -------------
    img1 = cv2.imread(os.path.join(REF, "1.png"), flags=cv2.IMREAD_COLOR)   # main.py:36
    img2 = cv2.imread(os.path.join(REF, "2.png"), flags=cv2.IMREAD_COLOR)   # main.py:37
    for n in (200, 500, 2000):                                              # main.py:35 uses 200
        det = OrbFeatureDetector(n_features=n)
        _, d1 = det.detect_and_compute(img1, None)
        _, d2 = det.detect_and_compute(img2, None)
        # frontend.py:185-187: train = last frame (1.png), query = current frame (2.png)
        save(f"c1_orb{n}.npz", record(np.ascontiguousarray(d2), np.ascontiguousarray(d1)))

    # ---- seeded synthetic: uniform, tie-heavy, matchable, tiny/ragged shapes ------
    rng = np.random.default_rng(20261018)
    save("rand_777x1234.npz", record(rng.integers(0, 256, (777, 32), dtype=np.uint8),
                                     rng.integers(0, 256, (1234, 32), dtype=np.uint8)))
    # tie-heavy: one byte varying over 4 values
    def tie(n):
        a = np.zeros((n, 32), dtype=np.uint8)
        a[:, 7] = rng.integers(0, 4, n, dtype=np.uint8)
        return a
    save("ties_300x257.npz", record(tie(300), tie(257)))
    save("alleq_65x130.npz", record(np.full((65, 32), 0xA5, np.uint8), np.full((130, 32), 0xA5, np.uint8)))
    # matchable (distribution M of SURVEY.md 8d)
    t = rng.integers(0, 256, (1500, 32), dtype=np.uint8)
    perm = rng.permutation(1500)[:1000]
    q = t[perm].copy()
    flip = np.packbits(rng.random((1000, 256)) < 0.1, axis=1)
    noisy = rng.random(1000) < 0.6
    q[noisy] ^= flip[noisy]
    q[~noisy] = rng.integers(0, 256, (int((~noisy).sum()), 32), dtype=np.uint8)
    save("matchable_1000x1500.npz", record(q, t))
    for nq, nt in ((1, 1), (3, 1), (1, 2), (2, 3), (31, 33), (129, 127), (5, 300)):
        save(f"small_{nq}x{nt}.npz", record(rng.integers(0, 256, (nq, 32), dtype=np.uint8),
                                            rng.integers(0, 256, (nt, 32), dtype=np.uint8)))

    # ---- train-collection API (the keyframe-database config C4) ------
    # every image keeps >= k rows: with an image shorter than k, cv2 4.13 returns
    # uninitialised memory (garbage trainIdx / distance 0.0) -- outside its defined domain
    sizes = (50, 70, 2, 30, 3)
    trains = [rng.integers(0, 256, (s, 32), dtype=np.uint8) for s in sizes]
    trains[3][:5] = trains[0][:5]            # duplicate rows across images -> cross-image ties
    qc = rng.integers(0, 256, (40, 32), dtype=np.uint8)
    qc[:5] = trains[0][:5]
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    bf.add([x for x in trains])
    d = {"query": qc, "sizes": np.array(sizes, dtype=np.int32),
         "train_cat": np.concatenate(trains, axis=0),
         "knn2": knn_arr(bf.knnMatch(qc, k=2), 2), "match": arr(bf.match(qc))}
    save("collection_40.npz", d)


if __name__ == "__main__":
    main()
