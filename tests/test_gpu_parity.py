"""Parity of the CUDA path (through the C ABI) against the oracle and the golden fixtures.

Bit-exact everywhere: this is integer work.  Every test runs BOTH kernel variants
((a) POPC, (b) tcgen05 int8) unless the variant is irrelevant.
"""
import os

import numpy as np
import pytest
import torch

import slam_experiments_b200 as sx
from slam_experiments_b200 import _native as nat
from slam_experiments_b200 import synth
from conftest import load_golden, pair_goldens, golden_files
from oracle import hamming_oracle as ho
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")
VARIANTS = ("popc", "i8", "f4")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def gpu_keys(q, t, variant, base=0):
    return nat.knn2_keys(dev(q), dev(t), train_base=base, variant=variant).cpu().numpy().view(np.uint64)


def _ids(paths):
    return [os.path.basename(p)[:-4] for p in paths]


def _rows(matches):
    return [(m.queryIdx, m.trainIdx, m.imgIdx, int(m.distance)) for m in matches]


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("path", pair_goldens(), ids=_ids(pair_goldens()))
def test_golden_fixtures_through_dropin(path, variant):
    """The reference's own outputs (tests/golden) reproduced by the drop-in classes."""
    g = load_golden(path)
    q, t = g["query"], g["train"]
    m = sx.BruteForceFeatureMatcher(norm_type=cv2.NORM_HAMMING, variant=variant)
    out = m.match(t, q)                                            # feature_matchers.py:36-39
    assert isinstance(out, tuple) and _rows(out) == [tuple(r) for r in g["ref_match"].tolist()]
    for th in (30, 64):                                            # feature_matchers.py:41-43
        out = m.match(t, q, dist_threshold=float(th))
        assert isinstance(out, list) and _rows(out) == [tuple(r) for r in g[f"ref_match_thr{th}"].tolist()]
    for k in (1, 2):
        rows = m.bf.knnMatch(q, t, k=k)
        exp = g[f"knn{k}"]
        assert len(rows) == exp.shape[0]
        for r, e in zip(rows, exp):
            assert _rows(r) == [tuple(x) for x in e.tolist() if x[0] >= 0]
    cc = sx.BFMatcher(cv2.NORM_HAMMING, crossCheck=True, variant=variant)
    assert _rows(cc.match(q, t)) == [tuple(r) for r in g["cross"].tolist()]
    rows = cc.knnMatch(q, t, k=1)
    assert [len(r) for r in rows] == [int(i in set(g["cross"][:, 0].tolist())) for i in range(q.shape[0])]
    for r in (70, 75, 80):
        pm = sx.BruteForceFeatureMatcher(cv2.NORM_HAMMING, ratio=r / 100.0, variant=variant)
        assert _rows(pm.match(t, q)) == [tuple(x) for x in g[f"ratio{r}"].tolist()]
        pm = sx.BruteForceFeatureMatcher(cv2.NORM_HAMMING, ratio=r / 100.0, cross_check=True, variant=variant)
        assert _rows(pm.match(t, q)) == [tuple(x) for x in g[f"pipe{r}"].tolist()]


@pytest.mark.parametrize("variant", VARIANTS)
def test_collection_api_golden(variant):
    g = load_golden(golden_files("collection_40.npz")[0])
    starts = np.concatenate([[0], np.cumsum(g["sizes"])])
    trains = [g["train_cat"][starts[i]:starts[i + 1]] for i in range(len(g["sizes"]))]
    bf = sx.BFMatcher(cv2.NORM_HAMMING, variant=variant)
    bf.add(trains)
    rows = bf.knnMatch(g["query"], k=2)
    assert [_rows(r) for r in rows] == [[tuple(x) for x in e.tolist()] for e in g["knn2"]]
    assert _rows(bf.match(g["query"])) == [tuple(x) for x in g["match"].tolist()]
    db = sx.ShardedKeyframeDatabase(g["sizes"], trains, variant=variant)
    assert [_rows(r) for r in db.knnMatch(g["query"], 2)] == [[tuple(x) for x in e.tolist()] for e in g["knn2"]]


SHAPES = [(1, 1), (1, 2), (2, 1), (3, 3), (31, 33), (32, 32), (33, 31), (127, 129), (128, 128), (129, 127),
          (255, 257), (256, 256), (257, 255), (200, 200), (1000, 1000), (5, 3000), (3000, 5), (2000, 2000),
          (640, 4096), (4100, 1030)]


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("nq,nt", SHAPES)
def test_knn2_keys_vs_oracle_shapes(nq, nt, variant):
    rng = np.random.default_rng(nq * 100003 + nt)
    for hi in (256, 3):                        # uniform, and tie-heavy (every byte in {0,1,2})
        q = rng.integers(0, hi, (nq, 32), dtype=np.uint8)
        t = rng.integers(0, hi, (nt, 32), dtype=np.uint8)
        assert np.array_equal(gpu_keys(q, t, variant), co.knn2_keys(q, t))
    assert np.array_equal(gpu_keys(q, t, variant, base=12345), co.knn2_keys(q, t, train_base=12345))


@pytest.mark.parametrize("variant", VARIANTS)
def test_adversarial_inputs(variant):
    rng = np.random.default_rng(99)
    nt = 1500
    # all-equal: every pair ties -> indices 0 and 1 for every query
    q = np.full((130, 32), 0x5A, np.uint8)
    t = np.full((nt, 32), 0x5A, np.uint8)
    k = gpu_keys(q, t, variant)
    assert (k[:, 0] == 0).all() and (k[:, 1] == 1).all()
    # monotonically improving distances: the top-2 update path fires on every train row
    t = np.zeros((256, 32), np.uint8)
    bits = np.zeros((256, 256), np.uint8)
    for j in range(256):
        bits[j, : 256 - j] = 1                 # row j has 256-j ones -> distance to zero query shrinks with j
    t = np.packbits(bits, axis=1, bitorder="little")
    q = np.zeros((64, 32), np.uint8)
    assert np.array_equal(gpu_keys(q, t, variant), co.knn2_keys(q, t))
    # extremes: distance 0 and 256
    q = rng.integers(0, 256, (70, 32), dtype=np.uint8)
    t = np.concatenate([~q[:35], q[35:]])
    assert np.array_equal(gpu_keys(q, t, variant), co.knn2_keys(q, t))
    # strided (non-contiguous) inputs, as cv2 accepts
    wide = rng.integers(0, 256, (300, 64), dtype=np.uint8)
    qs = torch.from_numpy(wide).cuda()[:, :32]
    ts = torch.from_numpy(wide).cuda()[::2, 32:]
    got = nat.knn2_keys(qs, ts, variant=variant).cpu().numpy().view(np.uint64)
    assert np.array_equal(got, co.knn2_keys(wide[:, :32], wide[::2, 32:]))
    bf = sx.BFMatcher(cv2.NORM_HAMMING, variant=variant)
    ref = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(wide[:, :32], wide[::2, 32:], k=2)
    assert [_rows(r) for r in bf.knnMatch(wide[:, :32], wide[::2, 32:], k=2)] == [_rows(r) for r in ref]


@pytest.mark.parametrize("variant", VARIANTS)
def test_empty_and_short_train(variant):
    bf = sx.BFMatcher(cv2.NORM_HAMMING, variant=variant)
    q = np.arange(96, dtype=np.uint8).reshape(3, 32)
    e = np.empty((0, 32), np.uint8)
    assert bf.match(q, e) == () and bf.knnMatch(q, e, k=2) == ((), (), ())
    rows = bf.knnMatch(q, q[:1], k=2)
    assert [len(r) for r in rows] == [1, 1, 1]                                 # k = min(k, Nt)
    assert sx.BruteForceFeatureMatcher(cv2.NORM_HAMMING, ratio=0.75, variant=variant).match(q[:1], q) == ()


@pytest.mark.parametrize("variant", VARIANTS)
def test_fused_pipeline_and_batched_vs_oracle(variant):
    frames = synth.frame_sequence(6, 700, seed=5)
    fd = torch.from_numpy(frames).cuda()
    # overlapping windows: query = frame i+1, train = frame i (frontend.py:185-187), zero-copy
    keys = nat.knn2_keys_batched(fd[1:], fd[:-1], variant=variant).cpu().numpy().view(np.uint64)
    for i in range(5):
        assert np.array_equal(keys[i], co.knn2_keys(frames[i + 1], frames[i]))
    for ratio, cross, thr in ((0.8, True, None), (0.75, False, None), (None, True, None), (None, False, 40.0),
                              (0.9, True, 64.5)):
        oq, ot, od, cnt = nat.match_fused(fd[1:], fd[:-1], ratio=ratio, cross_check=cross, dist_threshold=thr,
                                          variant=variant)
        oq, ot, od, cnt = oq.cpu().numpy(), ot.cpu().numpy(), od.cpu().numpy(), cnt.cpu().numpy()
        for i in range(5):
            eq, et, ed = co.pipeline(frames[i + 1], frames[i], ratio=ratio, cross_check=cross)
            if thr:
                allq, allt, alld = ho.match(frames[i + 1], frames[i])
                lim = max(2 * float(alld.min()), thr)
                keep = ed.astype(np.float64) < lim
                eq, et, ed = eq[keep], et[keep], ed[keep]
            n = int(cnt[i])
            assert n == len(eq)
            assert np.array_equal(oq[i, :n], eq) and np.array_equal(ot[i, :n], et) and np.array_equal(od[i, :n], ed)


def test_merge_top2_kernel_and_shard_split():
    rng = np.random.default_rng(21)
    q = rng.integers(0, 3, (333, 32), dtype=np.uint8)
    t = rng.integers(0, 3, (5000, 32), dtype=np.uint8)
    full = co.knn2_keys(q, t)
    for cuts in ((0, 5000), (0, 1, 5000), (0, 1234, 1235, 4000, 5000), (0, 2500, 5000)):
        parts = [nat.knn2_keys(dev(q), dev(t[a:b]), train_base=a, variant=("popc", "i8", "f4")[i % 3])
                 for i, (a, b) in enumerate(zip(cuts[:-1], cuts[1:]))]
        merged = nat.merge_top2(torch.stack(parts)).cpu().numpy().view(np.uint64)
        assert np.array_equal(merged, full)


@pytest.mark.parametrize("variant", ("i8", "f4"))
def test_prepared_operands_resident_database(variant):
    q, t = synth.keyframe_database(16, 500, 300, seed=1)
    qd, td = dev(q), dev(t)
    tp = nat.prepare(td, variant=variant)
    assert tp.numel() == nat.prepared_bytes(t.shape[0], variant)
    got = nat.knn2_keys_prepared(nat.prepare(qd, variant=variant), q.shape[0], tp, t.shape[0], train_base=77,
                                 variant=variant)
    assert np.array_equal(got.cpu().numpy().view(np.uint64), co.knn2_keys(q, t, train_base=77))
    # the prepared image really is the +/-1 expansion
    if variant == "i8":
        img = tp.cpu().numpy().view(np.int8)
        assert set(np.unique(img[: 128 * 256]).tolist()) <= {-1, 1}
    else:   # e2m1 nibbles: +1.0 = 0x2, -1.0 = 0xA
        img = tp.cpu().numpy()[: 128 * 128]
        assert set(np.unique(img & 0xF).tolist()) <= {0x2, 0xA} and set(np.unique(img >> 4).tolist()) <= {0x2, 0xA}


@pytest.mark.parametrize("variant", VARIANTS)
def test_incremental_keyframe_database(variant):
    """Resident database grown keyframe by keyframe (backend.py:31-37) == cv2 collection over all of them."""
    rng = np.random.default_rng(8)
    sizes = [300, 90, 1, 515, 0, 128, 127, 700]
    kfs = [rng.integers(0, 3, (s, 32), dtype=np.uint8) for s in sizes]
    q = rng.integers(0, 3, (150, 32), dtype=np.uint8)
    db = sx.ShardedKeyframeDatabase(sizes[:1], kfs[:1], variant=variant)
    for i in range(1, len(sizes)):
        assert db.append_keyframe(kfs[i]) == i
        img, loc, d = db.knn_tensors(q, 2)
        eimg, eloc, ed = ho.collection_knn(q, kfs[:i + 1], 2)
        assert np.array_equal(img, eimg) and np.array_equal(loc, eloc) and np.array_equal(d, ed), i


def test_host_context_c_abi_only():
    ctx = nat.HostContext()
    rng = np.random.default_rng(2)
    for nq, nt, v in ((200, 200, "auto"), (1000, 1300, "popc"), (1000, 1300, "i8"), (1000, 1300, "f4"), (3, 1, "auto")):
        q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
        t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
        assert np.array_equal(ctx.knn2_keys(q, t, v), co.knn2_keys(q, t))
    ctx.close()


@pytest.mark.parametrize("variant", VARIANTS)
def test_mid_size_vs_c_oracle(variant):
    """8k x 8k uniform (C3 point) and a 2000 x 64k database slice (C4-shaped), bit-exact."""
    q, t = synth.sweep(8192, "U")
    assert np.array_equal(gpu_keys(q, t, variant), co.knn2_keys(q, t))
    q, t = synth.keyframe_database(32, 2000, 2000, seed=4096)
    assert np.array_equal(gpu_keys(q, t, variant), co.knn2_keys(q, t))


def test_full_size_properties():
    """BASELINE-size checks through size-independent properties (the oracle cannot run these).

    64k x 64k sweep point: (1) all three variants agree bit-for-bit; (2) planted exact duplicates are
    found at distance 0 with the lowest index first; (3) permuting train rows permutes trainIdx
    wherever the top-2 distances are tie-free; (4) a random sample of rows equals the oracle."""
    n = 65536
    q, t = synth.sweep(n, "M")
    rng = np.random.default_rng(0)
    planted = rng.choice(n, 64, replace=False)
    t[planted] = q[planted]                    # query i == train i for the planted rows
    dup = rng.choice(np.setdiff1d(np.arange(n), planted), 64, replace=False)
    t[dup] = q[planted]                        # second copy elsewhere
    qd, td = dev(q), dev(t)
    k_i8 = nat.knn2_keys(qd, td, variant="i8").cpu().numpy().view(np.uint64)
    k_pc = nat.knn2_keys(qd, td, variant="popc").cpu().numpy().view(np.uint64)
    assert np.array_equal(k_i8, k_pc)
    k_f4 = nat.knn2_keys(qd, td, variant="f4").cpu().numpy().view(np.uint64)
    assert np.array_equal(k_f4, k_pc)
    idx, dist, _ = nat.split_keys(k_i8)
    assert (dist[planted] == 0).all()
    assert np.array_equal(idx[planted, 0], np.minimum(planted, dup))
    assert np.array_equal(idx[planted, 1], np.maximum(planted, dup))
    sample = rng.choice(n, 256, replace=False)
    assert np.array_equal(k_i8[sample], co.knn2_keys(q[sample], t))
    perm = rng.permutation(n)
    k_perm = nat.knn2_keys(qd, dev(t[perm]), variant="i8").cpu().numpy().view(np.uint64)
    pidx, pdist, _ = nat.split_keys(k_perm)
    assert np.array_equal(pdist, dist)         # distances are permutation invariant
    tie_free = dist[:, 0] != dist[:, 1]
    assert np.array_equal(perm[pidx[tie_free, 0]], idx[tie_free, 0])


@pytest.mark.parametrize("variant", ("i8", "f4"))
def test_split_launch_shared_row_thresholds(variant):
    """Few query blocks against a long train set: the train dimension is split over many CTAs per query
    row, which publish / read the row's second-best distance (floor kernels).  Ties across splits must still
    resolve to the lowest trainIdx, bit-exact."""
    rng = np.random.default_rng(77)
    nq, nt = 300, 40000
    # tie-heavy: one varying byte -> a handful of distinct distances, every row ties with thousands of others
    t = np.zeros((nt, 32), np.uint8)
    t[:, 5] = rng.integers(0, 4, nt)
    q = np.zeros((nq, 32), np.uint8)
    q[:, 5] = rng.integers(0, 4, nq)
    assert np.array_equal(gpu_keys(q, t, variant), co.knn2_keys(q, t))
    # all train rows equal: every pair ties, the answer is rows 0 and 1 for everybody
    t[:] = 0x5A
    assert np.array_equal(gpu_keys(q, t, variant), co.knn2_keys(q, t))
    # uniform rows with exact copies of every query planted late, early and in the middle of the train set
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    where = rng.choice(nt, (3, nq), replace=False)
    for w in where:
        t[w] = q
    keys = gpu_keys(q, t, variant)
    assert np.array_equal(keys, co.knn2_keys(q, t))
    idx, dist, _ = nat.split_keys(keys)
    assert (dist == 0).all() and np.array_equal(idx, np.sort(where, axis=0)[:2].T)
    # distances all above 128 (negative dot products): complement-like rows only
    t2 = (~q[rng.integers(0, nq, nt)]) ^ rng.integers(0, 2, (nt, 32), dtype=np.uint8)
    assert np.array_equal(gpu_keys(q, t2, variant), co.knn2_keys(q, t2))


@pytest.mark.parametrize("variant", ("i8", "f4"))
def test_batched_split_launch(variant):
    """Several independent problems in one launch, each split over many CTAs per query row: the shared row
    thresholds are indexed per (problem, row), and ragged sizes leave padding rows / columns in every tile."""
    rng = np.random.default_rng(21)
    b, nq, nt = 3, 333, 20011
    q = rng.integers(0, 256, (b, nq, 32), dtype=np.uint8)
    t = rng.integers(0, 4, (b, nt, 32), dtype=np.uint8)                 # low-entropy train rows: many ties
    t[1] = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    q[2, :50] = t[2, rng.choice(nt, 50, replace=False)]                 # exact matches in problem 2 only
    keys = nat.knn2_keys_batched(dev(q), dev(t), variant=variant).cpu().numpy().view(np.uint64)
    for i in range(b):
        assert np.array_equal(keys[i], co.knn2_keys(q[i], t[i])), i


def test_host_calls_replay_cuda_graphs_with_fresh_data():
    """The host-buffer entry points capture a CUDA graph of their stream sequence the second time a shape is
    seen and replay it afterwards: every replay must see that call's data, flags and thresholds."""
    rng = np.random.default_rng(99)
    ctx = nat.HostContext()
    for nq, nt in ((200, 200), (333, 1500)):
        for it in range(5):                                   # direct, capture, replay, replay, replay
            q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
            t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
            q[: nq // 2] = t[rng.choice(nt, nq // 2, replace=False)] ^ rng.integers(0, 2, (nq // 2, 32), dtype=np.uint8)
            for kwargs in ({}, {"ratio": 0.75, "cross_check": True}, {"ratio": 0.8}, {"dist_threshold": 30.0}):
                gq, gt, gd = ctx.match(q, t, **kwargs)
                if "dist_threshold" in kwargs:
                    eq, et, ed = ho.reference_match(t, q, 30.0)
                else:
                    eq, et, ed = co.pipeline(q, t, kwargs.get("ratio"), kwargs.get("cross_check", False))
                assert np.array_equal(gq, eq) and np.array_equal(gt, et) and np.array_equal(gd, ed), (nq, nt, it, kwargs)
            # resident frames: same slots, new contents every iteration
            pos_q = rng.integers(0, 700, (nq, 2)).astype(np.int32)
            pos_t = rng.integers(0, 700, (nt, 2)).astype(np.int32)
            ctx.frame_put(0, t, pos_t)
            ctx.frame_put(1, q, pos_q)
            gq, gt, gd, pq, pt = ctx.frame_match(0, 1, nq, ratio=0.75, cross_check=True, want_points=True)
            eq, et, ed = co.pipeline(q, t, 0.75, True)
            assert np.array_equal(gq, eq) and np.array_equal(gt, et) and np.array_equal(gd, ed)
            assert np.array_equal(pq, pos_q[eq]) and np.array_equal(pt, pos_t[et])
    ctx.close()


# =====================================================================================================
# BASELINE.json full-size configs, every row against the oracle
# =====================================================================================================
def _c4_database(kind):
    """2000 queries x 4096 keyframes x 2000 rows (C4).  At N=1 every CTA of the tensor-core kernels sees
    1730 tiles, i.e. runs the late refresh cadence of the shared row thresholds (tiles >= 512)."""
    if kind == "uniform":                      # the bench's own input
        return synth.keyframe_database(4096, 2000, 2000, seed=4096)
    rng = np.random.default_rng(4097)
    nt, nq = 4096 * 2000, 2000
    if kind == "dups":
        # 64 pool rows, each planted ~1280 times all over the database: the best distance of every query is
        # tied across hundreds of CTAs and across early and late tiles -> the two LOWEST copies must win
        t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
        pool = rng.integers(0, 256, (64, 32), dtype=np.uint8)
        where = rng.choice(nt, 64 * 1280, replace=False)
        t[where] = pool[rng.integers(0, 64, where.size)]
        q = pool[rng.integers(0, 64, nq)].copy()
        flips = rng.integers(0, 4, nq)
        for i in range(nq):
            for b in rng.choice(256, flips[i], replace=False):
                q[i, b >> 3] ^= 1 << (b & 7)
        return q, t
    # "lowentropy": every byte in {0, 1, 2}: a handful of distinct distances, ties everywhere, the exact
    # insertion path of the scan runs for most chunks of every tile
    return rng.integers(0, 3, (nq, 32), dtype=np.uint8), rng.integers(0, 3, (nt, 32), dtype=np.uint8)


@pytest.fixture(scope="module", params=("uniform", "dups", "lowentropy"))
def c4_case(request):
    q, t = _c4_database(request.param)
    return request.param, q, t, co.knn2_keys(q, t)


@pytest.mark.parametrize("variant", ("f4", "i8"))
def test_full_c4_all_rows_vs_oracle(c4_case, variant):
    """Full C4 at N=1, all 2000 rows bit-exact, through the resident (prepared) database path bench.py times."""
    kind, q, t, expect = c4_case
    if kind == "lowentropy" and variant == "i8":
        pytest.skip("covered by f4; keeps the suite short")
    td = dev(t)
    tp = nat.prepare(td, variant=variant)
    got = nat.knn2_keys_prepared(nat.prepare(dev(q), variant=variant), q.shape[0], tp, t.shape[0], 0, variant=variant)
    got = got.cpu().numpy().view(np.uint64)
    launch = nat.describe_launch(2000, t.shape[0], 1, variant)
    assert int(launch.split("tiles_per_cta=")[1]) > 512 and "_floor" in launch, launch   # the late-cadence branch runs
    bad = np.flatnonzero((got != expect).any(axis=1))
    assert bad.size == 0, f"{kind}/{variant}: {bad.size} rows differ, first {bad[:5]}"
    if kind == "dups":
        idx, dist, _ = nat.split_keys(got)
        assert (dist[:, 0] <= 3).all() and (idx[:, 0] < idx[:, 1]).all()


def test_full_c4_popc_sample_and_dropin(c4_case):
    """The POPC core on the full database (every 8th row, it is 25x slower) and the cv2-shaped collection API on top
    of the same keys: (imgIdx, trainIdx) decoding of global rows."""
    kind, q, t, expect = c4_case
    if kind != "uniform":
        pytest.skip("one database is enough for the slow core")
    got = gpu_keys(q[::8], t, "popc")
    assert np.array_equal(got, expect[::8])
    db = sx.ShardedKeyframeDatabase([2000] * 4096, [t[i * 2000:(i + 1) * 2000] for i in range(4096)])
    img, loc, d = db.knn_tensors(q, 2)
    gidx, gdist, _ = nat.split_keys(expect)
    assert np.array_equal(img, gidx // 2000) and np.array_equal(loc, gidx % 2000) and np.array_equal(d, gdist)


def test_full_c5_window_pipeline_vs_oracle():
    """Full C5: 32 problems of 10k x 10k, knnMatch k=2 -> ratio 0.75 -> mutual cross-check; q, t and distance
    of every surviving match of every problem, and the forward keys."""
    qs, ts = synth.local_window(32, 10000)
    qd, td = dev(qs), dev(ts)
    for variant in ("f4", "i8"):
        oq, ot, od, cnt = nat.match_fused(qd, td, ratio=0.75, cross_check=True, variant=variant)
        oq, ot, od, cnt = oq.cpu().numpy(), ot.cpu().numpy(), od.cpu().numpy(), cnt.cpu().numpy()
        for i in range(32):
            eq, et, ed = co.pipeline(qs[i], ts[i], 0.75, True)
            n = int(cnt[i])
            assert n == len(eq), (variant, i)
            assert np.array_equal(oq[i, :n], eq) and np.array_equal(ot[i, :n], et) and np.array_equal(od[i, :n], ed), (variant, i)
    keys = nat.knn2_keys_batched(qd, td, variant="f4").cpu().numpy().view(np.uint64)
    for i in (0, 13, 31):
        assert np.array_equal(keys[i], co.knn2_keys(qs[i], ts[i]))


def test_c2_real_orb_sequence():
    """C2 on the SURVEY 8(d) input: 100 warped 752 x 480 frames with real ORB descriptors (bit density 0.54,
    correlated rows).  All 99 consecutive problems, batched and one by one through the drop-in, against the
    oracle; three of them against the recorded output of the reference's matcher."""
    g = load_golden(golden_files("c2_sequence_orb2000.npz")[0])
    frames = g["descriptors"]
    assert frames.shape == (100, 2000, 32) and (g["counts"] == 2000).all()
    fd = dev(frames)
    for variant in VARIANTS:
        keys = nat.knn2_keys_batched(fd[1:], fd[:-1], variant=variant).cpu().numpy().view(np.uint64)
        for i in range(99):
            assert np.array_equal(keys[i], co.knn2_keys(frames[i + 1], frames[i])), (variant, i)
    m = sx.BruteForceFeatureMatcher(norm_type=cv2.NORM_HAMMING)
    for i in (0, 37, 98):
        out = m.match(frames[i], frames[i + 1])                   # frontend.py:181-187
        assert _rows(out) == [tuple(r) for r in g[f"ref_match_{i}"].tolist()]
    pm = sx.BruteForceFeatureMatcher(norm_type=cv2.NORM_HAMMING, ratio=0.75, cross_check=True)
    for i in (5, 60):
        eq, et, ed = co.pipeline(frames[i + 1], frames[i], 0.75, True)
        assert [(x.queryIdx, x.trainIdx, int(x.distance)) for x in pm.match(frames[i], frames[i + 1])] == \
            list(zip(eq.tolist(), et.tolist(), ed.tolist()))


def test_back_to_back_uploads_do_not_alias_staging():
    """Two unsynchronised device-level calls in a row: the second call's host copy must not overwrite the
    pinned staging buffer while the first call's H2D is still queued."""
    bf = sx.BFMatcher(cv2.NORM_HAMMING)
    rng = np.random.default_rng(3)
    t = rng.integers(0, 256, (60000, 32), dtype=np.uint8)
    qa = rng.integers(0, 256, (3000, 32), dtype=np.uint8)
    qb = rng.integers(0, 256, (3000, 32), dtype=np.uint8)
    ka = bf.knn_keys_device(qa, t)
    kb = bf.knn_keys_device(qb, t)
    assert np.array_equal(ka.cpu().numpy().view(np.uint64), co.knn2_keys(qa, t))
    assert np.array_equal(kb.cpu().numpy().view(np.uint64), co.knn2_keys(qb, t))


@pytest.mark.parametrize("variant", ("i8", "f4"))
def test_resident_database_packed_query(variant):
    """hm_knn2_resident: packed query rows against a prepared database.  The f4 core expands the query inside the
    k-NN kernel (every CTA writes its own A operand image): ragged query counts, padding rows inside the last
    256-row block, strided query views, a train_base."""
    rng = np.random.default_rng(31)
    t = rng.integers(0, 256, (70001, 32), dtype=np.uint8)
    tp = nat.prepare(dev(t), variant=variant)
    for nq in (1, 127, 256, 257, 700, 2000, 4097):
        q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
        q[: min(nq, 40)] = t[rng.choice(t.shape[0], min(nq, 40), replace=False)]         # exact matches: distance 0
        got = nat.knn2_keys_resident(dev(q), tp, t.shape[0], train_base=5, variant=variant)
        assert np.array_equal(got.cpu().numpy().view(np.uint64), co.knn2_keys(q, t, train_base=5)), nq
    wide = rng.integers(0, 256, (900, 64), dtype=np.uint8)
    got = nat.knn2_keys_resident(torch.from_numpy(wide).cuda()[:, 32:], tp, t.shape[0], variant=variant)
    assert np.array_equal(got.cpu().numpy().view(np.uint64), co.knn2_keys(wide[:, 32:], t))
    # tie-heavy low-entropy rows through the same path
    t2 = rng.integers(0, 3, (30000, 32), dtype=np.uint8)
    q2 = rng.integers(0, 3, (513, 32), dtype=np.uint8)
    got = nat.knn2_keys_resident(dev(q2), nat.prepare(dev(t2), variant=variant), t2.shape[0], variant=variant)
    assert np.array_equal(got.cpu().numpy().view(np.uint64), co.knn2_keys(q2, t2))


@pytest.mark.parametrize("variant", ("i8", "f4"))
@pytest.mark.parametrize("nt", (200, 1000, 40001, 300000))
def test_top2_at_group_tile_and_split_boundaries(variant, nt):
    """The tensor-core scan keeps its top-2 at the granularity of 8-column groups and takes the exact top-2 from the
    saved dots of two groups: plant the best and second-best train rows of every query at positions that stress it --
    both in one group, in adjacent groups, across the 64-column halves of a tile, across tiles and splits, in the last
    (ragged) rows, as exact duplicates (distance ties broken by index) and as near copies."""
    rng = np.random.default_rng(nt)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    pairs = [(0, 1), (6, 7), (7, 8), (63, 64), (127, 128), (120, 135), (nt - 2, nt - 1), (nt - 9, nt - 8), (5, nt - 1),
             (nt // 2, nt // 2 + 1), (nt // 2 - 1, nt // 2 + 64), (3, 3)]
    q = rng.integers(0, 256, (len(pairs) * 3, 32), dtype=np.uint8)
    for i, (a, b) in enumerate(pairs):
        base = q[3 * i].copy()
        q[3 * i + 1], q[3 * i + 2] = base, base
        t[a] = base                                   # distance 0
        if b != a:
            t[b] = base                               # an exact duplicate later: same distance, higher index
        near = base.copy()
        near[0] ^= 0x01
        if a + 3 < nt and a + 3 != b:
            t[a + 3] = near                           # distance 1 for the first two queries of the triple
        q[3 * i + 2, 1] ^= 0x03                       # third query: distance 2 to both copies, 3 to the near copy
    assert np.array_equal(gpu_keys(q, t, variant), co.knn2_keys(q, t))
    big_q = np.concatenate([q, synth.matchable_queries(t, 700, 5)])
    assert np.array_equal(gpu_keys(big_q, t, variant), co.knn2_keys(big_q, t))


def test_full_c4_repeated_launches_are_stable(c4_case):
    """Thirty full-size launches in a row give the same keys as the oracle every time: intermittent hazards (an
    experiment that kept the A operand in tensor memory lost a neighbour in a few of thirty launches) do not show in a
    single run."""
    kind, q, t, expect = c4_case
    if kind != "uniform":
        pytest.skip("one database is enough")
    tp = nat.prepare(dev(t), variant="f4")
    qp = nat.prepare(dev(q), variant="f4")
    qd = dev(q)
    for it in range(30):
        got = nat.knn2_keys_prepared(qp, q.shape[0], tp, t.shape[0], 0, variant="f4") if it % 2 else \
            nat.knn2_keys_resident(qd, tp, t.shape[0], variant="f4")
        bad = np.flatnonzero((got.cpu().numpy().view(np.uint64) != expect).any(axis=1))
        assert bad.size == 0, f"launch {it}: {bad.size} rows differ, first {bad[:5]}"
