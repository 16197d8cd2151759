"""Pin the oracle (numpy + C restatements) to the reference's own outputs.

The fixtures under tests/golden/ were produced by running
/root/reference/feature_matchers.py unmodified (tests/golden/make_golden.py).
cv2.BFMatcher is additionally consulted live when it is importable.
"""
import os

import numpy as np
import pytest

from conftest import load_golden, pair_goldens, golden_files
from oracle import hamming_oracle as ho
from oracle import c_oracle as co
from oracle import cv2_ref

RATIOS = (70, 75, 80)


def _ids(paths):
    return [os.path.basename(p)[:-4] for p in paths]


def _knn_expected(g, k):
    a = g[f"knn{k}"]
    return a[:, :, 1], a[:, :, 3]


@pytest.mark.parametrize("path", pair_goldens(), ids=_ids(pair_goldens()))
def test_numpy_oracle_vs_reference_outputs(path):
    g = load_golden(path)
    q, t = g["query"], g["train"]
    # the reference's own call (feature_matchers.py:39)
    mq, mt, md = ho.reference_match(t, q)
    assert np.array_equal(np.stack([mq, mt, md], 1), g["ref_match"][:, [0, 1, 3]])
    assert (g["ref_match"][:, 2] == 0).all()          # imgIdx is 0 for two-matrix calls
    for th in (30, 64):                               # feature_matchers.py:41-43
        fq, ft, fd = ho.reference_match(t, q, dist_threshold=float(th))
        assert np.array_equal(np.stack([fq, ft, fd], 1), g[f"ref_match_thr{th}"][:, [0, 1, 3]])
    for k in (1, 2):
        idx, dist = ho.knn(q, t, k)
        eidx, edist = _knn_expected(g, k)
        kk = min(k, t.shape[0])
        assert np.array_equal(idx, eidx[:, :kk]) and np.array_equal(dist, edist[:, :kk])
        assert (eidx[:, kk:] == -1).all()             # k = min(k, Nt): short rows
    cq, ct, cd = ho.cross_check_match(q, t)
    assert np.array_equal(np.stack([cq, ct, cd], 1), g["cross"][:, [0, 1, 3]])
    idx, dist = ho.knn(q, t, 2)
    for r in RATIOS:
        keep = ho.ratio_test(idx, dist, r / 100.0)
        got = np.stack([np.nonzero(keep)[0], idx[keep, 0], dist[keep, 0]], 1)
        assert np.array_equal(got, g[f"ratio{r}"][:, [0, 1, 3]])
        pq, pt, pd = ho.pipeline(q, t, ratio=r / 100.0, cross_check=True)
        assert np.array_equal(np.stack([pq, pt, pd], 1), g[f"pipe{r}"][:, [0, 1, 3]])


@pytest.mark.parametrize("path", pair_goldens(), ids=_ids(pair_goldens()))
def test_c_oracle_vs_reference_outputs(path):
    g = load_golden(path)
    q, t = g["query"], g["train"]
    keys = co.knn2_keys(q, t)
    assert np.array_equal(keys, ho.knn2_keys(q, t))
    eidx, edist = _knn_expected(g, 2)
    idx, dist, valid = ho.keys_to_arrays(keys)
    assert np.array_equal(np.where(valid, idx, -1), eidx)
    assert np.array_equal(np.where(valid, dist, -1), edist)
    for r in RATIOS:
        pq, pt, pd = co.pipeline(q, t, ratio=r / 100.0, cross_check=True)
        assert np.array_equal(np.stack([pq, pt, pd], 1), g[f"pipe{r}"][:, [0, 1, 3]])
    col = co.colmin_keys(q, t)
    d = ho.hamming_matrix(q, t)
    assert np.array_equal(col & np.uint64(0xFFFFFFFF), np.argmin(d, axis=0).astype(np.uint64))


def test_collection_api_golden():
    g = load_golden(golden_files("collection_40.npz")[0])
    sizes = g["sizes"]
    starts = np.concatenate([[0], np.cumsum(sizes)])
    trains = [g["train_cat"][starts[i]:starts[i + 1]] for i in range(len(sizes))]
    img, loc, dist = ho.collection_knn(g["query"], trains, 2)
    exp = g["knn2"]
    assert np.array_equal(loc, exp[:, :, 1])
    assert np.array_equal(img, exp[:, :, 2])
    assert np.array_equal(dist, exp[:, :, 3])
    # cross-image duplicate rows: the lower imgIdx wins
    assert (img[:5, 0] == 0).all() and (img[:5, 1] == 3).all() and (dist[:5] == 0).all()


def test_ratio_lut_equals_float_compare():
    for ratio in (0.5, 0.6, 0.7, 0.75, 0.8, 0.9, 1.0, 0.123456789):
        lut = ho.ratio_lut(ratio)
        d1 = np.arange(257)[:, None].astype(np.float64)
        d2 = np.arange(257)[None, :].astype(np.float64)
        assert np.array_equal(d1 < ratio * d2, np.arange(257)[:, None] < lut[None, :])


def test_merge_top2_equals_global():
    rng = np.random.default_rng(5)
    q = rng.integers(0, 256, (64, 32), dtype=np.uint8)
    t = rng.integers(0, 4, (900, 32), dtype=np.uint8)       # tie-heavy
    full = ho.knn2_keys(q, t)
    for cuts in ((0, 900), (0, 1, 900), (0, 300, 301, 777, 900), (0, 450, 900)):
        parts = [ho.knn2_keys(q, t[a:b], train_base=a) for a, b in zip(cuts[:-1], cuts[1:])]
        assert np.array_equal(ho.merge_top2_keys(np.stack(parts)), full)


def test_edge_shapes():
    e32 = np.empty((0, 32), np.uint8)
    q = np.arange(96, dtype=np.uint8).reshape(3, 32)
    assert ho.match(e32, q)[0].size == 0
    assert ho.match(q, e32)[0].size == 0
    assert ho.match(np.array([]), q)[0].size == 0          # Frame.get_descriptors() with no features
    idx, dist = ho.knn(q, e32, 2)
    assert idx.shape == (3, 0)
    idx, dist = ho.knn(q, q[:1], 2)
    assert idx.shape == (3, 1)                             # k = min(k, Nt)
    keys = ho.knn2_keys(q, q[:1])
    assert (keys[:, 1] == ho.NO_MATCH_KEY).all()
    assert ho.pipeline(q, q[:1])[0].size == 0              # ratio test drops rows shorter than 2


@pytest.mark.skipif(not cv2_ref.available(), reason="cv2 not importable")
def test_oracles_vs_cv2_live():
    import cv2
    rng = np.random.default_rng(11)
    for nq, nt, hi in ((257, 513, 256), (100, 1000, 2), (64, 64, 256), (300, 2, 256)):
        q = rng.integers(0, hi, (nq, 32), dtype=np.uint8)
        t = rng.integers(0, hi, (nt, 32), dtype=np.uint8)
        ref = cv2_ref.knn2_keys(q, t)
        assert np.array_equal(ho.knn2_keys(q, t), ref)
        assert np.array_equal(co.knn2_keys(q, t), ref)
        for ratio in (0.7, 0.9):
            pq, pt, pd, _ = cv2_ref.dmatches_to_arrays(cv2_ref.pipeline(q, t, ratio, True))
            oq, ot, od = ho.pipeline(q, t, ratio, True)
            assert np.array_equal(pq, oq) and np.array_equal(pt, ot) and np.array_equal(pd, od)
        m = cv2_ref.ReferenceMatcher(cv2.NORM_HAMMING).match(t, q, dist_threshold=40.0)
        mq, mt, md, _ = cv2_ref.dmatches_to_arrays(m)
        oq, ot, od = ho.reference_match(t, q, 40.0)
        assert np.array_equal(mq, oq) and np.array_equal(mt, ot) and np.array_equal(md, od)


def test_glue_oracle_mask_numpy_equals_cv2_rectangles():
    """The detection-mask restatement (utils.py:58-74): its numpy branch equals its cv2.rectangle branch,
    including squares that leave the image."""
    cv2 = pytest.importorskip("cv2")
    from oracle import glue_oracle as go
    rng = np.random.default_rng(4)
    for shape, n, r in (((48, 64), 30, 5), ((7, 9), 8, 0), ((20, 20), 12, 30)):
        pos = np.stack([rng.integers(-10, shape[1] + 10, n), rng.integers(-10, shape[0] + 10, n)], 1).astype(np.int32)
        for inner in (True, False):
            with_cv2 = go.detection_mask(shape, pos, r, inner)
            saved, go.cv2 = go.cv2, None
            try:
                plain = go.detection_mask(shape, pos, r, inner)
            finally:
                go.cv2 = saved
            assert np.array_equal(with_cv2, plain)


def test_candidate_pass_is_the_mutual_check():
    """The algorithm of hm_match_fused's kind::mxf4 path, restated on the oracle: ratio test first, the best train rows of
    its survivors become CANDIDATES, and the swapped (train -> query) top-1 runs over the candidate rows only.  A match
    (q, t) is mutual iff q is the best query of t, and only candidates are ever asked -- so the result equals the full
    cross-check pipeline, on random, matchable and tie-heavy inputs, with and without the ratio test."""
    rng = np.random.default_rng(42)
    cases = [(rng.integers(0, 256, (90, 32), dtype=np.uint8), rng.integers(0, 256, (140, 32), dtype=np.uint8)),
             (rng.integers(0, 3, (120, 32), dtype=np.uint8), rng.integers(0, 3, (77, 32), dtype=np.uint8))]
    t = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    q = t[rng.integers(0, 300, 200)].copy()
    q[:120] ^= np.packbits(rng.random((120, 256)) < 0.1, axis=1)
    cases.append((q, t))
    for q, t in cases:
        for ratio in (None, 0.75, 0.9):
            idx, dist = ho.knn(q, t, 2)
            keep = np.ones(len(q), bool) if ratio is None else ho.ratio_test(idx, dist, ratio)
            cand = np.unique(idx[keep, 0])                                   # candidate train rows
            if len(cand):
                bidx, _ = ho.knn(t[cand], q, 1)                              # best query of each candidate only
                best_query = dict(zip(cand.tolist(), bidx[:, 0].tolist()))
                keep &= np.array([best_query.get(int(idx[r, 0]), -1) == r for r in range(len(q))])
            got = np.nonzero(keep)[0]
            eq, et, ed = ho.pipeline(q, t, ratio, True)
            assert np.array_equal(got, eq) and np.array_equal(idx[got, 0], et) and np.array_equal(dist[got, 0], ed)


def test_group_level_top2_is_exact():
    """The tensor-core scan's data structure, restated in numpy: keep the two 8-column groups with the largest maxima
    (strict '>' in ascending order = earliest group on ties) and only THEIR dots; the exact (distance, index) top-2 of
    those 16 columns is the row's top-2 with cv2's tie rule.  Random, tie-heavy and planted-duplicate rows, ragged ends."""
    rng = np.random.default_rng(7)
    for nt, low in ((8, 256), (13, 256), (64, 256), (1000, 256), (1000, 3), (517, 2)):
        t = rng.integers(0, low, (nt, 32), dtype=np.uint8)
        q = rng.integers(0, low, (40, 32), dtype=np.uint8)
        q[:10] = t[rng.integers(0, nt, 10)]                               # exact duplicates: distance 0
        if nt > 20:
            t[nt - 1] = q[3]; t[7] = q[3]; t[8] = q[3]                     # ties across a group boundary and at the ragged end
        D = ho.hamming_matrix(q, t).astype(np.int64)
        dots = 256 - 2 * D
        pad = (-nt) % 8
        dots_p = np.concatenate([dots, np.full((len(q), pad), -10**6)], axis=1)     # masked tail (mask_tail in the kernel)
        for r in range(len(q)):
            v1 = v2 = -10**9
            g1 = g2 = -1
            for g in range(dots_p.shape[1] // 8):
                x = dots_p[r, 8 * g:8 * g + 8].max()
                if x > v2:
                    if x > v1:
                        v2, g2, v1, g1 = v1, g1, x, g
                    else:
                        v2, g2 = x, g
            cols = [c for g in (g1, g2) if g >= 0 for c in range(8 * g, 8 * g + 8) if c < nt]
            keys = sorted((int(D[r, c]), c) for c in cols)[:2]
            exact = sorted((int(D[r, c]), c) for c in range(nt))[:2]
            assert keys == exact, (nt, low, r)
