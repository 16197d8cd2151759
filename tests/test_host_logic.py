"""Host-side logic of the drop-in classes that needs no GPU: cv2-compatible validation,
empty-input short circuits, key decoding, ratio LUT, shard planning, loud failure without CUDA."""
import numpy as np
import pytest
import torch

import slam_experiments_b200 as sx
from slam_experiments_b200 import _native as nat
from slam_experiments_b200.feature_matchers import _build_dmatches
from oracle import hamming_oracle as ho

cv2 = pytest.importorskip("cv2")


def test_constructor_signature_matches_reference():
    m = sx.BruteForceFeatureMatcher(norm_type=cv2.NORM_HAMMING)       # slam.py:24
    assert isinstance(m, sx.FeatureMatcher) and hasattr(m, "bf")
    assert m.bf.empty() and m.bf.getTrainDescriptors() == ()
    with pytest.raises(cv2.error):
        sx.BruteForceFeatureMatcher(norm_type=cv2.NORM_L2)
    with pytest.raises(TypeError):
        sx.FeatureMatcher()                                             # abstract, like the reference ABC


def test_empty_query_short_circuits_like_cv2():
    m = sx.BruteForceFeatureMatcher(cv2.NORM_HAMMING)
    t = np.zeros((5, 32), np.uint8)
    assert m.match(t, np.array([])) == ()                               # Frame with no features
    assert m.match(t, np.empty((0, 32), np.uint8)) == ()
    assert m.bf.match(np.array([]), t) == ()
    assert m.bf.knnMatch(np.array([]), t, k=2) == ()
    assert m.bf.knnMatch(np.empty((0, 32), np.uint8), np.array([]), k=2) == ()


def test_validation_raises_cv2_error():
    bf = sx.BFMatcher(cv2.NORM_HAMMING)
    q = np.zeros((3, 32), np.uint8)
    for bad in (np.zeros((3, 32), np.float32), np.zeros((3, 32), np.int8), np.array([]),
                np.zeros((3, 64), np.uint8), np.zeros((3, 16), np.uint8)):
        with pytest.raises(cv2.error):
            bf.match(q, bad)
    with pytest.raises(cv2.error):
        bf.match(q, q, mask=np.ones((3, 3), np.uint8))
    with pytest.raises(cv2.error):
        sx.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).knnMatch(q, q, k=2)   # batch_distance.cpp:303
    with pytest.raises(cv2.error):
        bf.knnMatch(q, q, k=3)
    with pytest.raises(TypeError):
        bf.knnMatch(q, q)
    with pytest.raises(cv2.error):
        bf.add([np.zeros((1 << 18, 32), np.uint8)])                           # matchers.cpp:860
    bf.add([np.zeros((4, 32), np.uint8), np.ones((2, 32), np.uint8)])
    assert not bf.empty() and [a.shape for a in bf.getTrainDescriptors()] == [(4, 32), (2, 32)]
    bf.clear()
    assert bf.empty()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_fails_loudly_without_gpu():
    m = sx.BruteForceFeatureMatcher(cv2.NORM_HAMMING)
    q = np.zeros((3, 32), np.uint8)
    with pytest.raises(sx.NativeError, match="no CPU fallback"):
        m.match(q, q)
    with pytest.raises(sx.NativeError):
        m.bf.knnMatch(q, q, k=2)


def test_product_never_imports_the_oracle():
    import os, re
    from conftest import ROOT
    pkg = os.path.join(ROOT, "slam_experiments_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "libhamming_oracle" not in src, f


def test_ratio_lut_matches_oracle_and_float_compare():
    for r in (0.6, 0.7, 0.75, 0.8, 0.9):
        assert np.array_equal(nat.ratio_lut(r).astype(np.int32), ho.ratio_lut(r))


def test_split_keys_roundtrip():
    keys = np.array([[(7 << 32) | 5, nat.NO_MATCH], [(256 << 32) | 0xFFFFFFFE, (0 << 32) | 1]], dtype=np.uint64)
    idx, dist, valid = nat.split_keys(keys.view(np.int64))
    assert idx.tolist() == [[5, 0xFFFFFFFF], [0xFFFFFFFE, 1]]
    assert dist[0, 0] == 7 and dist[1, 0] == 256 and dist[1, 1] == 0
    assert valid.tolist() == [[True, False], [True, True]]


def test_dmatch_fields_like_cv2():
    ms = _build_dmatches([0, 1], [4, 3], [78.0, 12.0], 0)
    assert [(m.queryIdx, m.trainIdx, m.imgIdx, m.distance) for m in ms] == [(0, 4, 0, 78.0), (1, 3, 0, 12.0)]
    assert isinstance(ms[0], cv2.DMatch)
    ms = _build_dmatches([0, 1], [4, 3], [78.0, 12.0], [2, 5])
    assert [m.imgIdx for m in ms] == [2, 5]


def test_shard_ranges_tile_the_collection():
    rng = np.random.default_rng(3)
    for world in (1, 2, 3, 4, 8):
        for sizes in ([2000] * 4096, rng.integers(0, 50, 37).tolist(), [5], [0, 0, 7, 0], [1] * 3):
            r = sx.shard_ranges(sizes, world)
            assert len(r) == world and r[0][0] == 0 and r[-1][1] == len(sizes)
            starts = np.concatenate([[0], np.cumsum(sizes)])
            for a, b in zip(r[:-1], r[1:]):
                assert a[1] == b[0] and a[3] == b[2]
            for kf_lo, kf_hi, row_lo, row_hi in r:
                assert kf_lo <= kf_hi and row_lo == starts[kf_lo] and row_hi == starts[kf_hi]
    r = sx.shard_ranges([2000] * 4096, 8)
    assert all(hi - lo == 512 for lo, hi, _, _ in r)


def test_locate_rows_equals_searchsorted():
    """(imgIdx, trainIdx) decoding of global rows: the division fast path for equal-sized keyframes and the
    binary search for ragged collections give cv2's collection indices."""
    from slam_experiments_b200 import _native as nat
    rng = np.random.default_rng(0)
    for sizes in ([2000] * 64, [5, 7, 0, 3], [4, 4, 4, 5], [1], [3, 3], [0, 0, 9], [7] * 3 + [0]):
        starts = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        g = rng.integers(0, starts[-1], (50, 2)).astype(np.int64)
        img, loc = nat.locate_rows(starts, g)
        e = np.searchsorted(starts, g, side="right") - 1
        assert img.dtype == np.int32 and np.array_equal(img, e) and np.array_equal(loc, g - starts[e]), sizes[:4]


def test_dmatch_builders_agree(monkeypatch):
    """cv2.DMatch objects from the C helper (direct field writes through an assumed object layout, probed at import),
    from its two-phase twin (allocate while the kernels run, fill afterwards) and from the pure-Python fallback
    (the type's own constructor) are identical, for the flat (match) and the row (knnMatch) shapes."""
    cv2 = pytest.importorskip("cv2")
    from slam_experiments_b200 import feature_matchers as fm
    rng = np.random.default_rng(4)
    n = 600
    q = np.repeat(np.arange(n // 2, dtype=np.int32), 2)
    t = rng.integers(0, 1 << 18, n).astype(np.int32)
    d = rng.integers(0, 257, n).astype(np.float32)
    img = rng.integers(0, 8191, n).astype(np.int32)

    def rows_of(out, rows):
        flat = [m for r in out for m in r] if rows else list(out)
        assert all(type(m) is cv2.DMatch for m in flat)
        return [(m.queryIdx, m.trainIdx, m.imgIdx, m.distance) for m in flat], type(out), (type(out[0]) if rows else None)

    for rows in (0, 2):
        for im in (0, 3, img):
            fast = fm._build_dmatches(q, t, d, im, rows=rows)
            with monkeypatch.context() as mp:
                mp.setattr(fm, "_fast_build", None)
                slow = fm._build_dmatches(q, t, d, im, rows=rows)
            assert rows_of(fast, rows) == rows_of(slow, rows)
            if fm._fast_build is not None:
                pre = fm._prealloc_dmatches(n, rows)
                assert pre is not None
                assert rows_of(fm._fill_dmatches(pre, q, t, d, im, rows=rows), rows) == rows_of(slow, rows)
    assert fm._fast_build is not None, "the C helper should load in this image (falls back silently elsewhere)"
