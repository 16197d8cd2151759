"""CPU oracle for the Hamming brute-force matching path (TEST INFRASTRUCTURE ONLY).

This module is the *checker*: a numpy restatement of what the reference's hot
path computes.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  The product
(``slam_experiments_b200``) never does.

Where the arithmetic lives
--------------------------
The reference (`/root/reference/feature_matchers.py:32-44`) is six lines of
Python around ``cv2.BFMatcher``; the arithmetic is in the third-party
dependency **opencv-python, pinned 4.9.0.80** (`/root/reference/poetry.lock:1798-1799`),
whose sources are not vendored under `/root/reference`.  The published
algorithm restated here is OpenCV's ``BFMatcher::knnMatchImpl``
(modules/features2d/src/matchers.cpp) on top of ``cv::batchDistance``
(modules/core/src/batch_distance.cpp): for every query row, the integer
Hamming distance (XOR + popcount over the 32 descriptor bytes) to every train
row, then the K smallest by a *stable* ascending order, i.e. ties go to the
lowest train index.  Distances are returned as float32 holding exact integers.

Parity pin
----------
The reference has no tests, golden vectors or fixtures for this path
(SURVEY.md section 4), so the pin is outputs of the reference itself:
``tests/golden/make_golden.py`` imports `/root/reference/feature_matchers.py`
unmodified, runs it (and the ``cv2.BFMatcher`` calls it wraps) on the bundled
1.png/2.png ORB descriptors and on seeded random / tie-heavy inputs, and
commits the results under ``tests/golden/``.  ``tests/test_oracle.py`` checks
every function below against those fixtures, and additionally against
``cv2.BFMatcher`` live whenever cv2 is importable.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

NO_MATCH_KEY = np.uint64(0xFFFFFFFFFFFFFFFF)
DESC_BYTES = 32


def _as_desc(a: np.ndarray) -> np.ndarray:
    a = np.asarray(a)
    if a.ndim != 2 or a.dtype != np.uint8:
        raise ValueError("descriptors must be a 2-D uint8 array")
    return np.ascontiguousarray(a)


def hamming_matrix(query: np.ndarray, train: np.ndarray) -> np.ndarray:
    """Integer Hamming distance of every (query row, train row) pair.

    Follows cv::batchDistance with NORM_HAMMING (batch_distance.cpp,
    ``normHamming``): popcount of the XOR of the two rows.  Called through
    `/root/reference/feature_matchers.py:39`.
    """
    q = _as_desc(query)
    t = _as_desc(train)
    if q.shape[1] != t.shape[1]:
        raise ValueError("descriptor width mismatch")
    out = np.empty((q.shape[0], t.shape[0]), dtype=np.int32)
    step = max(1, (1 << 24) // max(1, t.shape[0] * q.shape[1]))
    for s in range(0, q.shape[0], step):
        x = q[s:s + step, None, :] ^ t[None, :, :]
        out[s:s + step] = np.bitwise_count(x).sum(axis=2, dtype=np.int32)
    return out


def knn(query: np.ndarray, train: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Stable k-nearest train rows per query row.

    Restates BFMatcher::knnMatchImpl (matchers.cpp) as reached from
    `/root/reference/feature_matchers.py:39`: ``k = min(k, Nt)``; results in
    ascending distance, ties to the lowest trainIdx.
    Returns ``(idx[Nq,k'], dist[Nq,k'])`` as int32.
    """
    d = hamming_matrix(query, train)
    nq, nt = d.shape
    kk = min(k, nt)
    if kk == 0:
        return np.empty((nq, 0), np.int32), np.empty((nq, 0), np.int32)
    key = (d.astype(np.int64) << 32) | np.arange(nt, dtype=np.int64)[None, :]
    if kk < nt:
        key = np.partition(key, kk - 1, axis=1)[:, :kk]
    key = np.sort(key, axis=1)[:, :kk]
    return (key & 0xFFFFFFFF).astype(np.int32), (key >> 32).astype(np.int32)


def knn2_keys(query: np.ndarray, train: np.ndarray, train_base: int = 0) -> np.ndarray:
    """Top-2 as packed ``uint64`` keys ``(dist << 32) | (train_base + trainIdx)``.

    This is the device output format of ``hm_knn2`` (include/hm_matcher.h);
    a missing neighbour (Nt < 2) is ``NO_MATCH_KEY``.  ``min`` over keys is the
    cv2 tie-break (lowest trainIdx).
    """
    idx, dist = knn(query, train, 2)
    nq = idx.shape[0]
    out = np.full((nq, 2), NO_MATCH_KEY, dtype=np.uint64)
    kk = idx.shape[1]
    if kk:
        out[:, :kk] = (dist.astype(np.uint64) << np.uint64(32)) | (
            idx.astype(np.uint64) + np.uint64(train_base))
    return out


def merge_top2_keys(keys: np.ndarray) -> np.ndarray:
    """Merge ``[G, Nq, 2]`` per-shard keys into the global ``[Nq, 2]`` top-2.

    Top-k over a union is the top-k of the per-set top-k's.
    """
    k = np.asarray(keys, dtype=np.uint64)
    g, nq, _ = k.shape
    flat = np.sort(k.transpose(1, 0, 2).reshape(nq, 2 * g), axis=1)
    return np.ascontiguousarray(flat[:, :2])


def match(query: np.ndarray, train: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """``cv2.BFMatcher(NORM_HAMMING).match(query, train)`` as ``(q, t, d)`` arrays.

    One match per query row ordered by queryIdx; empty when either side is
    empty.  `/root/reference/feature_matchers.py:39`.
    """
    q = np.asarray(query)
    if q.size == 0 or np.asarray(train).shape[0] == 0:
        e = np.empty(0, np.int32)
        return e, e.copy(), e.copy()
    idx, dist = knn(query, train, 1)
    return np.arange(idx.shape[0], dtype=np.int32), idx[:, 0].copy(), dist[:, 0].copy()


def cross_check_match(query: np.ndarray, train: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """``cv2.BFMatcher(NORM_HAMMING, crossCheck=True).match(query, train)``.

    Strict mutual nearest neighbour with lowest-index ties both ways
    (SURVEY.md E4): keep (q, t) iff t is q's stable argmin over train and q is
    t's stable argmin over queries.
    """
    q = np.asarray(query)
    if q.size == 0 or np.asarray(train).shape[0] == 0:
        e = np.empty(0, np.int32)
        return e, e.copy(), e.copy()
    d = hamming_matrix(query, train)
    fwd = np.argmin(d, axis=1)        # np.argmin returns the first minimum
    bwd = np.argmin(d, axis=0)
    qi = np.arange(d.shape[0])
    keep = bwd[fwd] == qi
    return (qi[keep].astype(np.int32), fwd[keep].astype(np.int32),
            d[qi[keep], fwd[keep]].astype(np.int32))


def ratio_lut(ratio: float) -> np.ndarray:
    """Integer form of Lowe's test: ``d1 < ratio * d2``  <=>  ``d1 < lut[d2]``.

    ``ratio * d2`` is evaluated in float64 exactly as the Python expression
    ``row[0].distance < ratio * row[1].distance`` does (distances are exact
    integers in float32), and for integer d1, ``d1 < x  <=>  d1 < ceil(x)``.
    """
    return np.array([math.ceil(ratio * float(d2)) for d2 in range(257)], dtype=np.int32)


def ratio_test(idx: np.ndarray, dist: np.ndarray, ratio: float) -> np.ndarray:
    """Boolean keep-mask of Lowe's ratio test over k=2 results.

    Rows with fewer than two neighbours are dropped.
    """
    if idx.shape[1] < 2:
        return np.zeros(idx.shape[0], dtype=bool)
    return dist[:, 0].astype(np.float64) < ratio * dist[:, 1].astype(np.float64)


def reference_match(source: np.ndarray, query: np.ndarray,
                    dist_threshold: Optional[float] = None):
    """Restatement of ``BruteForceFeatureMatcher.match``.

    `/root/reference/feature_matchers.py:36-44`: argument flip (source is the
    train set), then the optional strict ``distance < max(2*min_dist,
    dist_threshold)`` filter.  Returns ``(q, t, d)`` arrays.
    """
    q, t, d = match(query, source)
    if dist_threshold and len(q) != 0:
        min_dist = float(d.min())
        keep = d.astype(np.float64) < max(2 * min_dist, dist_threshold)
        return q[keep], t[keep], d[keep]
    return q, t, d


def pipeline(query: np.ndarray, train: np.ndarray, ratio: Optional[float] = 0.75,
             cross_check: bool = True):
    """North-star pipeline: knnMatch k=2 -> Lowe ratio ->
    mutual cross-check; output ordered by queryIdx, as ``(q, t, d)`` arrays."""
    qn = np.asarray(query)
    if qn.size == 0 or np.asarray(train).shape[0] == 0:
        e = np.empty(0, np.int32)
        return e, e.copy(), e.copy()
    idx, dist = knn(query, train, 2)
    nq = idx.shape[0]
    keep = np.ones(nq, dtype=bool)
    if ratio is not None:
        keep &= ratio_test(idx, dist, ratio)
    if cross_check:
        cq, ct, _ = cross_check_match(query, train)
        mutual = np.zeros(nq, dtype=bool)
        mutual[cq] = idx[cq, 0] == ct
        keep &= mutual
    qi = np.nonzero(keep)[0].astype(np.int32)
    return qi, idx[qi, 0].copy(), dist[qi, 0].copy()


def collection_knn(query: np.ndarray, trains: Sequence[np.ndarray], k: int):
    """cv2's train-collection API: ``bf.add(trains); bf.knnMatch(query, k)``.

    Global stable top-k over the row concatenation, reported as
    ``(imgIdx, trainIdx, dist)`` each ``[Nq, k']``; cross-image ties go to the
    lower imgIdx (matchers.cpp IMGIDX_SHIFT=18 packing).
    """
    sizes = np.array([np.asarray(t).shape[0] for t in trains], dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(sizes)])
    nonempty = [np.asarray(t) for t in trains if np.asarray(t).shape[0]]
    cat = np.concatenate(nonempty, axis=0) if nonempty else np.empty((0, DESC_BYTES), np.uint8)
    gidx, dist = knn(query, cat, k)
    img = (np.searchsorted(starts, gidx, side="right") - 1).astype(np.int32)
    local = (gidx - starts[img]).astype(np.int32)
    return img, local, dist


def keys_to_arrays(keys: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Split packed keys into ``(idx, dist, valid)``."""
    k = np.asarray(keys, dtype=np.uint64)
    valid = k != NO_MATCH_KEY
    return ((k & np.uint64(0xFFFFFFFF)).astype(np.int64), (k >> np.uint64(32)).astype(np.int32), valid)
