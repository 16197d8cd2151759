"""The reference path run through the real engine, cv2.BFMatcher.

TEST INFRASTRUCTURE ONLY - see oracle/hamming_oracle.py for the policy.

`/root/reference/feature_matchers.py:32-44` cannot travel to the GPU box
(`/root/reference` does not exist there) but the engine it calls,
``cv2.BFMatcher``, is importable in the image.  ``ReferenceMatcher`` restates the
reference class's six lines verbatim in behaviour so the reference arm of
``bench.py`` and the parity tests drive the same cv2 calls the reference makes.
"""
from __future__ import annotations

from operator import attrgetter
from typing import Optional, Sequence

import numpy as np

try:  # cv2 is present in this image; guarded so the numpy/C oracle works without it
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def available() -> bool:
    return cv2 is not None


class ReferenceMatcher:
    """Behavioural twin of the reference's ``BruteForceFeatureMatcher``
    (`/root/reference/feature_matchers.py:32-44`)."""

    def __init__(self, norm_type: int):
        self.bf = cv2.BFMatcher(normType=norm_type)

    def match(self, source_descriptors, query_descriptors, dist_threshold: Optional[float] = None):
        matches = self.bf.match(query_descriptors, source_descriptors)
        if dist_threshold and len(matches) != 0:
            min_dist = min(matches, key=attrgetter("distance")).distance
            return [m for m in matches if m.distance < max(2 * min_dist, dist_threshold)]
        return matches


def knn_match(query, train, k=2):
    return cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(query, train, k=k)


def knn2_keys(query, train) -> np.ndarray:
    """cv2 knnMatch(k=2) as packed keys (see hamming_oracle.knn2_keys)."""
    rows = knn_match(query, train, 2)
    out = np.full((len(rows), 2), 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
    for i, row in enumerate(rows):
        for j, m in enumerate(row):
            out[i, j] = (int(m.distance) << 32) | m.trainIdx
    return out


def pipeline(query, train, ratio: Optional[float] = 0.75, cross_check: bool = True):
    """cv2 composition of SURVEY.md 8a row P: knnMatch k=2 -> ratio -> crossCheck."""
    rows = knn_match(query, train, 2)
    good = [r[0] for r in rows
            if len(r) >= 1 and (ratio is None or (len(r) == 2 and r[0].distance < ratio * r[1].distance))]
    if cross_check:
        cc = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(query, train)
        ok = {(m.queryIdx, m.trainIdx) for m in cc}
        good = [m for m in good if (m.queryIdx, m.trainIdx) in ok]
    return good


def dmatches_to_arrays(matches: Sequence) -> tuple:
    q = np.array([m.queryIdx for m in matches], dtype=np.int32)
    t = np.array([m.trainIdx for m in matches], dtype=np.int32)
    d = np.array([int(m.distance) for m in matches], dtype=np.int32)
    img = np.array([m.imgIdx for m in matches], dtype=np.int32)
    return q, t, d, img
