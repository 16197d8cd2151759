"""Drop-in for the reference's feature-matcher plugin, running on B200.

Mirrors `/root/reference/feature_matchers.py`:

* ``FeatureMatcher`` -- the ABC of lines 9-29 (``match(source, query)``);
* ``BruteForceFeatureMatcher`` -- lines 32-44: same constructor
  (``norm_type``), same ``match(source_descriptors, query_descriptors,
  dist_threshold=None)`` with the argument flip (source is the *train* set,
  line 39) and the strict ``distance < max(2*min_dist, dist_threshold)`` filter
  (lines 41-43); it returns cv2's tuple when unfiltered and a list when
  filtered, exactly like the reference;
* ``BFMatcher`` -- the object behind ``.bf``, shaped like ``cv2.BFMatcher``:
  ``match`` / ``knnMatch`` (two-matrix and train-collection forms), ``add``,
  ``clear``, ``empty``, ``getTrainDescriptors``; constructor keyword
  ``crossCheck``.

All distance work runs in the hand-written CUDA kernels behind
``include/hm_matcher.h``; PyTorch only carries the bytes to the device and
back.  Only NORM_HAMMING, 32-byte descriptors, k in {1, 2} and ``mask=None``
are in scope -- anything else raises (cv2-compatible ``cv2.error`` when cv2 is
importable) instead of silently falling back to the CPU.

Beyond the reference, keyword-only knobs give the north-star pipeline
(SURVEY.md 8a row P): ``ratio`` (Lowe test on k=2) and ``cross_check`` (mutual
nearest neighbour), and ``*_tensors`` twins return arrays instead of DMatch
objects, whose construction dominates end-to-end time at >= 2k matches.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as nat


try:  # cv2 supplies the DMatch type so results drop into cv2.drawMatches etc.
    import cv2 as _cv2
except Exception:  # pragma: no cover - cv2 is present in the target image
    _cv2 = None

NORM_HAMMING = 6          # cv2.NORM_HAMMING
_IMGIDX_SHIFT = 18        # cv2 packs (imgIdx, trainIdx) with 18 bits of local row (matchers.cpp)
_MAX_ROWS_PER_IMAGE = 1 << _IMGIDX_SHIFT
_MAX_IMAGES = 8192


if _cv2 is not None:
    DMatch = _cv2.DMatch

    class MatcherError(_cv2.error):
        """Raised where cv2.BFMatcher would raise cv2.error(-215)."""
else:  # pragma: no cover
    class DMatch:  # minimal stand-in with cv2.DMatch's fields
        __slots__ = ("queryIdx", "trainIdx", "imgIdx", "distance")

        def __init__(self, queryIdx=-1, trainIdx=-1, distance=float("inf")):
            self.queryIdx, self.trainIdx, self.imgIdx, self.distance = queryIdx, trainIdx, -1, distance

    class MatcherError(ValueError):
        pass


def _load_fast_builder():
    """C helper that writes cv2.DMatch fields directly (csrc/hmfast.c); verified once against the
    type's own constructor, otherwise unused."""
    if _cv2 is None:
        return None
    try:
        from . import _hmfast
        probe = _hmfast.dmatch_build(DMatch, np.array([7, 1], np.int32), np.array([9, 2], np.int32),
                                     np.array([12.0, 256.0], np.float32), None, 3, 0)
        ok = [(m.queryIdx, m.trainIdx, m.imgIdx, m.distance) for m in probe] == [(7, 9, 3, 12.0), (1, 2, 3, 256.0)]
        rows = _hmfast.dmatch_build(DMatch, np.array([7, 1], np.int32), np.array([9, 2], np.int32),
                                    np.array([12.0, 256.0], np.float32), np.array([4, 5], np.int32), 0, 2)
        ok = ok and isinstance(probe[0], DMatch) and [m.imgIdx for m in rows[0]] == [4, 5]
        pre = _hmfast.dmatch_alloc(DMatch, 2, 2)
        _hmfast.dmatch_fill(pre, np.array([7, 1], np.int32), np.array([9, 2], np.int32),
                            np.array([12.0, 256.0], np.float32), np.array([4, 5], np.int32), 0, 2)
        ok = ok and [(m.queryIdx, m.trainIdx, m.imgIdx, m.distance) for m in pre[0]] == [(7, 9, 4, 12.0), (1, 2, 5, 256.0)]
        return _hmfast.dmatch_build if ok else None
    except Exception:
        return None


_fast_build = _load_fast_builder()


def _build_dmatches(q, t, d, img=0, rows: int = 0):
    """Bulk DMatch construction from arrays.

    ``rows == 0`` -> list of DMatch; ``rows == k`` -> tuple of tuples of k (knnMatch shape).
    Fast path: csrc/hmfast.c (~40 ns/object).  Fallback: 3-arg ctor + one attribute store
    (~0.4 us/object; the 4-arg ctor costs 2.4 us).
    """
    q = np.ascontiguousarray(q, dtype=np.int32)
    t = np.ascontiguousarray(t, dtype=np.int32)
    d = np.ascontiguousarray(d, dtype=np.float32)
    img_arr = None if isinstance(img, (int, np.integer)) else np.ascontiguousarray(img, dtype=np.int32)
    if _fast_build is not None:
        return _fast_build(DMatch, q, t, d, img_arr, 0 if img_arr is not None else int(img), rows)
    out = list(map(DMatch, q.tolist(), t.tolist(), d.tolist()))
    if img_arr is None:
        for m in out:
            m.imgIdx = int(img)
    else:
        for m, i in zip(out, img_arr.tolist()):
            m.imgIdx = i
    if rows:
        return tuple(tuple(out[i:i + rows]) for i in range(0, len(out), rows))
    return out


def _prealloc_dmatches(n: int, rows: int = 0):
    """Containers + DMatch objects for ``n`` results, fields still zero: object allocation is the expensive part of the
    result list, and for knnMatch its size (``nq * k``) is known before the kernels finish -- so it runs while they do.
    Returns None when the C helper is unavailable (the caller then builds the list afterwards)."""
    if _fast_build is None or n <= 0:
        return None
    try:
        from . import _hmfast
        return _hmfast.dmatch_alloc(DMatch, int(n), int(rows))
    except Exception:
        return None


def _fill_dmatches(container, q, t, d, img=0, rows: int = 0):
    from . import _hmfast
    q = np.ascontiguousarray(q, dtype=np.int32)
    t = np.ascontiguousarray(t, dtype=np.int32)
    d = np.ascontiguousarray(d, dtype=np.float32)
    img_arr = None if isinstance(img, (int, np.integer)) else np.ascontiguousarray(img, dtype=np.int32)
    _hmfast.dmatch_fill(container, q, t, d, img_arr, 0 if img_arr is not None else int(img), rows)
    return container


class _Staging:
    """Pinned host buffers reused across calls (H2D / D2H without pageable-memory syncs).

    An H2D out of a pinned buffer is asynchronous: callers such as ``BFMatcher.knn_keys_device`` return
    unsynchronised device tensors, so the next call may arrive while the copy is still queued behind running
    kernels.  Every buffer therefore carries the event of its last H2D, and the next host write waits for it
    (the C path, ``hm_frame_put``, guards its staging buffers the same way)."""

    def __init__(self):
        self._pin = {}
        self._h2d_done = {}

    def pinned(self, key: str, nbytes: int) -> torch.Tensor:
        buf = self._pin.get(key)
        if buf is None or buf.numel() < nbytes:
            self.wait(key)                           # the old buffer may still feed a queued copy
            buf = torch.empty(max(nbytes, 4096), dtype=torch.uint8).pin_memory()
            self._pin[key] = buf
        return buf

    def wait(self, key: str) -> None:
        ev = self._h2d_done.get(key)
        if ev is not None:
            ev.synchronize()

    def to_device(self, key: str, arr: np.ndarray, device: torch.device) -> torch.Tensor:
        n = arr.shape[0]
        nbytes = n * nat.DESC_BYTES
        if nbytes == 0:
            return torch.empty((0, nat.DESC_BYTES), dtype=torch.uint8, device=device)
        host = self.pinned(key, nbytes)[:nbytes].view(n, nat.DESC_BYTES)
        self.wait(key)                               # the previous H2D has left this buffer
        host.numpy()[...] = arr                      # one memcpy, handles strided views
        out = host.to(device, non_blocking=True)
        ev = self._h2d_done.get(key)
        if ev is None:
            ev = self._h2d_done[key] = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        return out

    def to_host(self, key: str, t: torch.Tensor) -> np.ndarray:
        """D2H through pinned memory; synchronises the current stream."""
        nbytes = t.numel() * t.element_size()
        if nbytes == 0:
            return np.empty(tuple(t.shape), dtype=_np_dtype(t.dtype))
        flat = t.contiguous().view(torch.uint8).view(-1)
        host = self.pinned(key, nbytes)[:nbytes]
        host.copy_(flat, non_blocking=True)
        torch.cuda.current_stream(t.device).synchronize()
        return host.numpy().view(_np_dtype(t.dtype)).reshape(tuple(t.shape)).copy()


def _np_dtype(dt: torch.dtype):
    return {torch.int64: np.int64, torch.int32: np.int32, torch.uint8: np.uint8}[dt]


def _is_empty_query(a) -> bool:
    """cv2 returns () for any empty query, whatever its dtype/shape (np.array([]) from
    Frame.get_descriptors() with no features, `/root/reference/primitives.py:200-205`)."""
    if isinstance(a, torch.Tensor):
        return a.numel() == 0
    return np.asarray(a).size == 0


class BFMatcher:
    """cv2.BFMatcher-shaped brute-force Hamming matcher on B200.

    ``BFMatcher(normType=cv2.NORM_HAMMING, crossCheck=False)`` as in
    `/root/reference/feature_matchers.py:34`.
    """

    def __init__(self, normType: int = NORM_HAMMING, crossCheck: bool = False, *, device=None,
                 variant: str = "auto"):
        if normType != NORM_HAMMING:
            raise MatcherError(f"only NORM_HAMMING ({NORM_HAMMING}) is supported on the B200 path, got {normType}")
        nat.variant_id(variant)
        self.normType = normType
        self.crossCheck = bool(crossCheck)
        self.variant = variant
        self._device = device
        self._staging = _Staging()
        self._train: List[np.ndarray] = []          # collection API (host copies, like cv2's Mat list)
        self._train_dev: Optional[torch.Tensor] = None
        self._train_starts: Optional[np.ndarray] = None
        self._host_ctx: Optional[nat.HostContext] = None

    def _ctx(self) -> "nat.HostContext":
        """Lazily created hm_context: the numpy-in / numpy-out fast path (one C call per match)."""
        if self._host_ctx is None:
            with torch.cuda.device(self._dev()):
                self._host_ctx = nat.HostContext()
        return self._host_ctx

    def _on_device(self):
        """Context manager that makes the matcher's device current -- a no-op object when it already is (the
        torch context manager alone costs ~10 us, a fifth of a 200 x 200 match)."""
        return nat.on_device(self._dev())

    # ---- input handling -------------------------------------------------------------------
    def _dev(self) -> torch.device:
        d = self.__dict__.get("_dev_cached")
        if d is None:
            d = self._dev_cached = nat.require_cuda(self._device)
        return d

    def _validate(self, a, name: str):
        """cv2's type checks (batch_distance.cpp:274,282): uint8, 2-D, equal width; here width == 32."""
        if isinstance(a, torch.Tensor):
            if a.dtype != torch.uint8 or a.dim() != 2 or a.shape[1] != nat.DESC_BYTES:
                raise MatcherError(f"{name}: expected uint8 [N, {nat.DESC_BYTES}] descriptors, got "
                                   f"{a.dtype} {tuple(a.shape)}")
            return a
        a = np.asarray(a)
        if a.dtype != np.uint8 or a.ndim != 2:
            raise MatcherError(f"{name}: NORM_HAMMING needs 2-D uint8 descriptors, got {a.dtype} {a.shape} "
                               "(cv2: batch_distance.cpp:282)")
        if a.shape[1] != nat.DESC_BYTES:
            raise MatcherError(f"{name}: only {nat.DESC_BYTES}-byte (256-bit ORB) descriptors are supported on "
                               f"the B200 path, got width {a.shape[1]}")
        return a

    def _upload(self, key: str, a) -> torch.Tensor:
        dev = self._dev()
        if isinstance(a, torch.Tensor):
            t = a if a.is_cuda else a.to(dev, non_blocking=True)
            if t.stride(1) != 1 or t.stride(0) % 16 or t.data_ptr() % 16:
                t = t.contiguous()
            return t
        return self._staging.to_device(key, a, dev)

    # ---- tensor-returning twins -----------------------------------------------------------------
    def knn_keys_device(self, queryDescriptors, trainDescriptors) -> torch.Tensor:
        """Packed top-2 keys on the device, ``[Nq, 2]`` int64 (uint64 bit pattern)."""
        q = self._upload("q", self._validate(queryDescriptors, "queryDescriptors"))
        t = self._upload("t", self._validate(trainDescriptors, "trainDescriptors"))
        return nat.knn2_keys(q, t, variant=self.variant)

    def knn_tensors(self, queryDescriptors, trainDescriptors, k: int = 2) -> Tuple[np.ndarray, np.ndarray]:
        """``(trainIdx[Nq, k'], distance[Nq, k'])`` with ``k' = min(k, Nt)``; no DMatch objects."""
        if k not in (1, 2):
            raise MatcherError("only k in {1, 2} is supported on the B200 path")
        if isinstance(queryDescriptors, np.ndarray) and isinstance(trainDescriptors, np.ndarray):
            q = self._validate(queryDescriptors, "queryDescriptors")
            t = self._validate(trainDescriptors, "trainDescriptors")
            with self._on_device():
                keys = self._ctx().knn2_keys(q, t, self.variant)
        else:
            keys = self._staging.to_host("keys", self.knn_keys_device(queryDescriptors, trainDescriptors))
        idx, dist, valid = nat.split_keys(keys)
        kk = int(valid[0].sum()) if len(valid) else 0
        kk = min(kk, k)
        return idx[:, :kk].astype(np.int32), dist[:, :kk]

    def match_tensors(self, queryDescriptors, trainDescriptors, *, ratio: Optional[float] = None,
                      cross_check: Optional[bool] = None, dist_threshold: Optional[float] = None
                      ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Fused pipeline ``(queryIdx, trainIdx, distance)`` ordered by queryIdx."""
        cross = self.crossCheck if cross_check is None else bool(cross_check)
        if _is_empty_query(queryDescriptors):
            e = np.empty(0, np.int32)
            return e, e.copy(), e.copy()
        q = self._validate(queryDescriptors, "queryDescriptors")
        t = self._validate(trainDescriptors, "trainDescriptors")
        if q.shape[0] == 0 or t.shape[0] == 0:
            e = np.empty(0, np.int32)
            return e, e.copy(), e.copy()
        if isinstance(q, np.ndarray) and isinstance(t, np.ndarray):
            with self._on_device():
                return self._ctx().match(q, t, ratio=ratio, cross_check=cross, dist_threshold=dist_threshold,
                                         variant=self.variant)
        qd, td = self._upload("q", q), self._upload("t", t)
        oq, ot, od, cnt = nat.match_fused(qd.unsqueeze(0), td.unsqueeze(0), ratio=ratio, cross_check=cross,
                                          dist_threshold=dist_threshold, variant=self.variant)
        packed = torch.cat([cnt.view(1), oq.view(-1), ot.view(-1), od.view(-1)])
        host = self._staging.to_host("m", packed)
        n, nq = int(host[0]), q.shape[0]
        return host[1:1 + n].copy(), host[1 + nq:1 + nq + n].copy(), host[1 + 2 * nq:1 + 2 * nq + n].copy()

    # ---- cv2.BFMatcher API ----------------------------------------------------------------------
    def match(self, queryDescriptors, trainDescriptors=None, mask=None) -> tuple:
        """``cv2.BFMatcher.match``: best train row per query row, as a tuple of DMatch."""
        if mask is not None:
            raise MatcherError("mask is not supported on the B200 path")
        if trainDescriptors is None:
            rows = self._collection_knn(queryDescriptors, 1)
            return tuple(r[0] for r in rows if r)
        if _is_empty_query(queryDescriptors):
            return ()
        q, t, d = self.match_tensors(queryDescriptors, trainDescriptors)
        return tuple(_build_dmatches(q, t, d, 0))

    def knnMatch(self, queryDescriptors, trainDescriptors=None, k: int = None, mask=None,
                 compactResult: bool = False) -> tuple:
        """``cv2.BFMatcher.knnMatch``: per query row a tuple of up to k DMatch."""
        if isinstance(trainDescriptors, int) and k is None:      # collection form: knnMatch(query, k)
            trainDescriptors, k = None, trainDescriptors
        if k is None:
            raise TypeError("knnMatch() missing required argument 'k'")
        if mask is not None:
            raise MatcherError("mask is not supported on the B200 path")
        if self.crossCheck and k != 1:
            raise MatcherError("crossCheck=True requires k == 1 (cv2: batch_distance.cpp:303)")
        if trainDescriptors is None:
            return self._collection_knn(queryDescriptors, k)
        if _is_empty_query(queryDescriptors):
            return ()
        if self.crossCheck:
            q, t, d = self.match_tensors(queryDescriptors, trainDescriptors)
            nq = np.asarray(queryDescriptors).shape[0] if not isinstance(queryDescriptors, torch.Tensor) \
                else queryDescriptors.shape[0]
            rows: List[tuple] = [()] * nq
            for m in _build_dmatches(q, t, d, 0):
                rows[m.queryIdx] = (m,)
            return tuple(rows)
        idx, dist = self.knn_tensors(queryDescriptors, trainDescriptors, k)
        nq, kk = idx.shape
        if kk == 0:
            return tuple(() for _ in range(nq))
        qi = np.repeat(np.arange(nq, dtype=np.int32), kk)
        return _build_dmatches(qi, idx.reshape(-1), dist.reshape(-1), 0, rows=kk)

    # ---- train collection (keyframe database on one GPU; SURVEY.md call stack C) ---------------
    def add(self, descriptors: Sequence) -> None:
        new = []
        for dsc in descriptors:
            a = dsc.cpu().numpy() if isinstance(dsc, torch.Tensor) else np.asarray(dsc)
            a = np.ascontiguousarray(self._validate(a, "descriptors"))
            if a.shape[0] >= _MAX_ROWS_PER_IMAGE:
                raise MatcherError("too many rows in one train image (cv2: matchers.cpp:860)")
            new.append(a)
        if len(self._train) + len(new) >= _MAX_IMAGES:
            raise MatcherError("too many train images (cv2: matchers.cpp:856)")
        self._train.extend(new)
        self._train_dev = None

    def clear(self) -> None:
        self._train = []
        self._train_dev = None
        self._train_starts = None

    def empty(self) -> bool:
        return len(self._train) == 0

    def getTrainDescriptors(self) -> tuple:
        return tuple(self._train)

    def train(self) -> None:
        """Upload the collection once; later queries only move the query descriptors."""
        if self._train_dev is None:
            sizes = np.array([a.shape[0] for a in self._train], dtype=np.int64)
            self._train_starts = np.concatenate([[0], np.cumsum(sizes)])
            cat = np.concatenate(self._train, axis=0) if len(self._train) else np.empty((0, 32), np.uint8)
            self._train_dev = torch.from_numpy(cat).to(self._dev())

    def _collection_knn(self, queryDescriptors, k: int) -> tuple:
        if k not in (1, 2):
            raise MatcherError("only k in {1, 2} is supported on the B200 path")
        if _is_empty_query(queryDescriptors):
            return ()
        self.train()
        q = self._upload("q", self._validate(queryDescriptors, "queryDescriptors"))
        keys = self._staging.to_host("keys", nat.knn2_keys(q, self._train_dev, variant=self.variant))
        gidx, dist, valid = nat.split_keys(keys)
        nq = gidx.shape[0]
        kk = min(k, int(valid[0].sum()) if nq else 0)
        if kk == 0:
            return tuple(() for _ in range(nq))
        gidx, dist = gidx[:, :kk], dist[:, :kk]
        img, local = nat.locate_rows(self._train_starts, gidx)
        qi = np.repeat(np.arange(nq, dtype=np.int32), kk)
        return _build_dmatches(qi, local.reshape(-1), dist.reshape(-1), img.reshape(-1), rows=kk)


class FeatureMatcher(ABC):
    """Plugin contract of `/root/reference/feature_matchers.py:9-29`."""

    @abstractmethod
    def match(self, source_descriptors: np.ndarray, query_descriptors: np.ndarray) -> Sequence:
        raise NotImplementedError

    @classmethod
    def draw_matches(cls, source_img, source_keypoints, query_img, query_keypoints, matches) -> None:
        """Source img is on the right and Query img is on the left (reference lines 16-29; GUI helper)."""
        if _cv2 is None:  # pragma: no cover
            raise RuntimeError("draw_matches needs cv2")
        matches_img = _cv2.drawMatches(query_img, query_keypoints, source_img, source_keypoints, matches, None)
        _cv2.imshow("Matches", matches_img)


class BruteForceFeatureMatcher(FeatureMatcher):
    """Drop-in for `/root/reference/feature_matchers.py:32-44`, constructed the same way
    (``BruteForceFeatureMatcher(norm_type=cv2.NORM_HAMMING)``, `slam.py:24`) and injected
    into ``Frontend`` unchanged (`frontend.py:55-69`)."""

    def __init__(self, norm_type: int, *, ratio: Optional[float] = None, cross_check: bool = False,
                 device=None, variant: str = "auto"):
        self.bf = BFMatcher(normType=norm_type, device=device, variant=variant)
        self.ratio = ratio
        self.cross_check = bool(cross_check)

    def match_tensors(self, source_descriptors, query_descriptors, dist_threshold: Optional[float] = None):
        """Array twin of ``match``: ``(queryIdx, trainIdx, distance)``; same argument order."""
        if _is_empty_query(query_descriptors):
            e = np.empty(0, np.int32)
            return e, e.copy(), e.copy()
        return self.bf.match_tensors(query_descriptors, source_descriptors, ratio=self.ratio,
                                     cross_check=self.cross_check,
                                     dist_threshold=dist_threshold if dist_threshold else None)

    def match(self, source_descriptors: np.ndarray, query_descriptors: np.ndarray,
              dist_threshold: Optional[float] = None) -> Sequence:
        # reference line 39: bf.match(query_descriptors, source_descriptors) -- source is the train set
        q, t, d = self.match_tensors(source_descriptors, query_descriptors, dist_threshold)
        matches = _build_dmatches(q, t, d, 0)
        # reference lines 41-44: a list when the distance filter ran, cv2's tuple otherwise
        if dist_threshold and len(matches) != 0:
            return matches
        return tuple(matches)
