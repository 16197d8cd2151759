"""CPU restatement of the reference code either side of the matcher call.

TEST INFRASTRUCTURE ONLY - see oracle/hamming_oracle.py for the policy.

The reference's own modules cannot be imported here (`utils.py` imports jaxlie, which is absent), so the
two functions below restate, on plain arrays, exactly what its loops compute:

* ``matched_point_lists``  - `/root/reference/utils.py:13-19` (``pose_estimation_2d2d``) and `:41-47`
  (``triangulation``, before its camera normalisation): for every match, the position of the train-side
  feature of the source (last) frame and of the query-side feature of the query (current) frame;
* ``detection_mask``       - `/root/reference/utils.py:58-74` (``get_featured_detection_mask``) without the
  optional re-projection branch: a constant mask with one filled, inclusive, clipped square per feature,
  drawn with the same ``cv2.rectangle`` call when cv2 is importable, else with numpy slicing.

``Feature.position`` is ``np.array(keypoint.pt, dtype=np.int32)`` (`primitives.py:108-110`), i.e. truncation
toward zero of the sub-pixel keypoint.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def matched_point_lists(matches: Sequence, source_positions: np.ndarray, query_positions: np.ndarray
                        ) -> Tuple[np.ndarray, np.ndarray]:
    """`utils.py:13-19`: the two lists built by the loop over ``matches`` (objects with ``trainIdx`` /
    ``queryIdx``), as ``(M, 2)`` arrays."""
    source_pts, query_pts = [], []
    for m in matches:
        source_pts.append(source_positions[m.trainIdx])
        query_pts.append(query_positions[m.queryIdx])
    return (np.array(source_pts, dtype=np.int32).reshape(-1, 2), np.array(query_pts, dtype=np.int32).reshape(-1, 2))


def detection_mask(shape, positions: np.ndarray, radius: int, inner: bool = True) -> np.ndarray:
    """`utils.py:66-74`: ``np.full(shape, 0 if inner else 255)`` then, per feature,
    ``cv2.rectangle(mask, pt - [r, r], pt + [r, r], 255 if inner else 0, cv2.FILLED)``."""
    mask = np.full(shape, fill_value=0 if inner else 255, dtype=np.uint8)
    value = 255 if inner else 0
    shift = np.array([radius, radius])
    for pt in np.asarray(positions).reshape(-1, 2):
        if cv2 is not None:
            mask = cv2.rectangle(mask, pt - shift, pt + shift, value, cv2.FILLED)
        else:  # inclusive corners, clipped to the image
            x0, y0 = max(int(pt[0]) - radius, 0), max(int(pt[1]) - radius, 0)
            x1, y1 = min(int(pt[0]) + radius, shape[1] - 1), min(int(pt[1]) + radius, shape[0] - 1)
            if x0 <= x1 and y0 <= y1:
                mask[y0:y1 + 1, x0:x1 + 1] = value
    return mask
