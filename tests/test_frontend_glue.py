"""The callers either side of the matcher: descriptor packing, the
`_match_features` call site and the `queryIdx`/`trainIdx` consumers, on reference-shaped objects."""
import numpy as np
import pytest
import torch

import slam_experiments_b200 as sx
from conftest import load_golden, golden_files
from oracle import glue_oracle as go, hamming_oracle as ho

cv2 = pytest.importorskip("cv2")


class Feature:                      # /root/reference/primitives.py:92-112, the fields the path touches
    def __init__(self, kp, descriptor):
        self.keypoint, self.descriptor, self.map_point, self.is_outlier = kp, descriptor, None, False

    @property
    def position(self):
        return np.array(self.keypoint.pt, dtype=np.int32)


class Frame:                        # /root/reference/primitives.py:160-211
    def __init__(self, descriptors, seed=0):
        rng = np.random.default_rng(seed)
        self.features = [Feature(cv2.KeyPoint(float(x), float(y), 31.0), d)
                         for d, (x, y) in zip(descriptors, rng.uniform(0, 600, (len(descriptors), 2)))]

    def get_descriptors(self):      # primitives.py:200-205
        return np.array([f.descriptor for f in self.features])


class RecordingMatcher(sx.FeatureMatcher):
    def __init__(self):
        self.calls = []

    def match(self, source_descriptors, query_descriptors, dist_threshold=None):
        self.calls.append((source_descriptors, query_descriptors, dist_threshold))
        q, t, d = ho.reference_match(source_descriptors, query_descriptors, dist_threshold)
        return tuple(cv2.DMatch(int(a), int(b), 0, float(c)) for a, b, c in zip(q, t, d))


def test_get_descriptors_matches_reference_semantics():
    g = load_golden(golden_files("c1_orb200.npz")[0])
    f = Frame(g["train"])
    d = sx.get_descriptors(f.features)
    assert d.dtype == np.uint8 and d.shape == (200, 32) and np.array_equal(d, g["train"])
    e = sx.get_descriptors([])
    assert e.shape == (0,) and e.dtype == np.float64          # the reference's empty-frame quirk


def test_match_features_argument_order_and_consumers():
    g = load_golden(golden_files("c1_orb200.npz")[0])
    last, cur = Frame(g["train"], 1), Frame(g["query"], 2)
    m = RecordingMatcher()
    matches = sx.match_features(m, last, cur)                 # frontend.py:185-187
    src, qry, thr = m.calls[0]
    assert np.array_equal(src, g["train"]) and np.array_equal(qry, g["query"]) and thr is None
    assert [(x.queryIdx, x.trainIdx, int(x.distance)) for x in matches] == [tuple(r) for r in g["ref_match"][:, [0, 1, 3]].tolist()]
    # consumers: frontend.py:174-177
    for i, f in enumerate(last.features):
        f.map_point = ("mp", i) if i % 3 == 0 else None
    n = sx.propagate_map_points(matches, last.features, cur.features)
    exp = 0
    for x in matches:
        if x.trainIdx % 3 == 0:
            assert cur.features[x.queryIdx].map_point is not None
            exp += 1
    assert n == exp
    q, t = g["ref_match"][:, 0], g["ref_match"][:, 1]
    for f in cur.features:
        f.map_point = None
    assert sx.propagate_map_points((q, t), last.features, cur.features) == exp
    # utils.py:13-19: packed point arrays == the reference's list-building loops
    sp, qp = sx.matched_point_arrays(q, t, sx.keypoint_array(last.features), sx.keypoint_array(cur.features))
    ref_s = np.array([last.features[x.trainIdx].position for x in matches])
    ref_q = np.array([cur.features[x.queryIdx].position for x in matches])
    assert np.array_equal(sp, ref_s) and np.array_equal(qp, ref_q)
    assert sx.keypoint_array([]).shape == (0, 2)


@pytest.mark.gpu
def test_dropin_through_the_call_site_and_resident_store():
    """a5 on the GPU: `_match_features` with the B200 matcher injected == the reference's own output,
    and the device-resident store gives the same matches without re-uploading the train frame."""
    for name in ("c1_orb200.npz", "c1_orb2000.npz"):
        g = load_golden(golden_files(name)[0])
        last, cur = Frame(g["train"], 1), Frame(g["query"], 2)
        matcher = sx.BruteForceFeatureMatcher(norm_type=cv2.NORM_HAMMING)
        out = sx.match_features(matcher, last, cur)
        rows = [(m.queryIdx, m.trainIdx, m.imgIdx, int(m.distance)) for m in out]
        assert isinstance(out, tuple) and rows == [tuple(r) for r in g["ref_match"].tolist()]
        out = sx.match_features(matcher, last, cur, 30.0)
        assert isinstance(out, list) and len(out) == len(g["ref_match_thr30"])
        empty = Frame([])
        assert sx.match_features(matcher, last, empty) == ()          # current frame lost all features
        with pytest.raises(cv2.error):
            sx.match_features(matcher, empty, cur)                    # cv2 raises for an np.array([]) train set

        store = sx.FrameDescriptorStore(capacity=3)
        store.put("last", last.get_descriptors())
        store.put("cur", cur.get_descriptors())
        rows = [(m.queryIdx, m.trainIdx, m.imgIdx, int(m.distance)) for m in store.match("last", "cur")]
        assert rows == [tuple(r) for r in g["ref_match"].tolist()]
        thr = store.match("last", "cur", 30.0)
        assert isinstance(thr, list) and [(m.queryIdx, m.trainIdx) for m in thr] == [tuple(r) for r in g["ref_match_thr30"][:, :2].tolist()]
        store.put("none", np.array([]))
        assert store.match("last", "none") == () and store.match("none", "cur") == ()
        for i in range(4):
            store.put(i, g["query"])
        assert len(store) == 3 and "last" not in store
        pipe = sx.FrameDescriptorStore(ratio=0.75, cross_check=True)
        pipe.put(0, g["train"]); pipe.put(1, g["query"])
        q, t, d = pipe.match_tensors(0, 1)
        assert np.array_equal(np.stack([q, t, d], 1), g["pipe75"][:, [0, 1, 3]])


@pytest.mark.gpu
def test_matched_points_gathered_on_device():
    """8f rank 2: the point lists of `utils.py:13-19` / `:41-47` (a Python loop over DMatch objects in the
    reference) come back as two packed arrays gathered on the device, for the plain matcher and the pipeline."""
    rng = np.random.default_rng(3)
    for name in ("c1_orb200.npz", "c1_orb2000.npz"):
        g = load_golden(golden_files(name)[0])
        last, cur = Frame(g["train"], 1), Frame(g["query"], 2)
        pos_last, pos_cur = sx.keypoint_array(last.features), sx.keypoint_array(cur.features)
        for kwargs, dist_thr in (({}, None), ({}, 30.0), ({"ratio": 0.75, "cross_check": True}, None)):
            store = sx.FrameDescriptorStore(**kwargs)
            store.put("last", last.get_descriptors(), pos_last)
            store.put("cur", cur.get_descriptors(), pos_cur)
            sp, qp = store.matched_points("last", "cur", dist_thr)
            matches = store.match("last", "cur", dist_thr)                       # same matches as DMatch objects
            ref_s, ref_q = go.matched_point_lists(matches, pos_last, pos_cur)          # utils.py:13-19
            assert sp.dtype == np.int32 and np.array_equal(sp, ref_s) and np.array_equal(qp, ref_q)
        store.put("none", np.array([]), np.empty((0, 2)))
        sp, qp = store.matched_points("last", "none")
        assert sp.shape == (0, 2) and qp.shape == (0, 2)
        store.put("nopos", g["query"])
        with pytest.raises(cv2.error):
            store.matched_points("last", "nopos")
    # raw kernel, batched, random indices
    b, nq, nt = 3, 500, 700
    q_idx = torch.from_numpy(rng.integers(0, nq, (b, nq)).astype(np.int32)).cuda()
    t_idx = torch.from_numpy(rng.integers(0, nt, (b, nq)).astype(np.int32)).cuda()
    cnt = torch.tensor([nq, 0, 123], dtype=torch.int32).cuda()
    qp = torch.from_numpy(rng.integers(-5, 2000, (b, nq, 2)).astype(np.int32)).cuda()
    tp = torch.from_numpy(rng.integers(-5, 2000, (b, nt, 2)).astype(np.int32)).cuda()
    oq, ot = sx._native.gather_points(q_idx, t_idx, cnt, qp, tp)
    for i, n in enumerate(cnt.tolist()):
        assert torch.equal(oq[i, :n], qp[i][q_idx[i, :n].long()]) and torch.equal(ot[i, :n], tp[i][t_idx[i, :n].long()])


@pytest.mark.gpu
def test_detection_mask_equals_cv2_rectangles():
    """8f rank 4: `utils.get_featured_detection_mask` (`utils.py:58-74`), one cv2.rectangle per feature in the
    reference, rasterised on the device -- identical masks, borders and out-of-image points included."""
    rng = np.random.default_rng(11)

    reference = go.detection_mask                          # utils.py:66-74 on Feature.position arrays

    for shape, n, radius in (((480, 752), 2000, 10), ((480, 640), 200, 10), ((33, 47), 40, 3), ((5, 9), 6, 0), ((64, 64), 0, 4)):
        pos = np.stack([rng.integers(-15, shape[1] + 15, n), rng.integers(-15, shape[0] + 15, n)], 1).astype(np.int32)
        if n:
            pos[0] = (0, 0); pos[-1] = (shape[1] - 1, shape[0] - 1)          # corners
        for inner in (True, False):
            got = sx.get_featured_detection_mask(shape, pos, radius, inner)
            assert got.dtype == np.uint8 and got.shape == shape
            assert np.array_equal(got, reference(shape, pos, radius, inner)), (shape, n, radius, inner)
    # through Feature objects, as the frontend calls it (frontend.py:236-243)
    g = load_golden(golden_files("c1_orb200.npz")[0])
    fr = Frame(g["train"], 5)
    pos = np.array([f.position for f in fr.features])
    assert np.array_equal(sx.get_featured_detection_mask((480, 640), fr.features, 10, True), reference((480, 640), pos, 10, True))


@pytest.mark.gpu
def test_store_accepts_cuda_tensors():
    """Frames that already live on the device (CUDA tensors) are matched through the device-pointer entry
    points; same results as the numpy / frame-slot path."""
    g = load_golden(golden_files("c1_orb500.npz")[0])
    last, cur = Frame(g["train"], 1), Frame(g["query"], 2)
    pos_last, pos_cur = sx.keypoint_array(last.features), sx.keypoint_array(cur.features)
    a, b = sx.FrameDescriptorStore(ratio=0.8), sx.FrameDescriptorStore(ratio=0.8)
    a.put(0, g["train"], pos_last); a.put(1, g["query"], pos_cur)
    b.put(0, torch.from_numpy(g["train"]).cuda(), pos_last); b.put(1, torch.from_numpy(g["query"]).cuda(), pos_cur)
    for x, y in zip(a.match_tensors(0, 1), b.match_tensors(0, 1)):
        assert np.array_equal(x, y)
    for x, y in zip(a.matched_points(0, 1), b.matched_points(0, 1)):
        assert np.array_equal(x, y)
    assert [(m.queryIdx, m.trainIdx) for m in a.match(0, 1)] == [(m.queryIdx, m.trainIdx) for m in b.match(0, 1)]


@pytest.mark.gpu
def test_keyframe_window_batched_match_equals_oracle():
    """C5-shaped local window: resident keyframes, one batched pipeline call, one query for all / one query each."""
    from oracle import c_oracle as co
    from slam_experiments_b200 import synth
    rng = np.random.default_rng(17)
    kfs = rng.integers(0, 256, (5, 700, 32), dtype=np.uint8)
    win = sx.KeyframeWindow(capacity=6, ratio=0.8, cross_check=True, variant="f4")
    for i in (0, 1, 2, 4, 5):
        win.put(i, kfs[min(i, 4) if i != 5 else 3])
    order = [0, 1, 2, 4, 3]
    q1 = synth.matchable_queries(kfs[2], 650, 3)
    out = win.match_tensors(q1)
    assert len(out) == 5
    for (q, t, d), k in zip(out, order):
        eq, et, ed = co.pipeline(q1, kfs[k], 0.8, True)
        assert np.array_equal(q, eq) and np.array_equal(t, et) and np.array_equal(d, ed)
    qs = np.stack([synth.matchable_queries(kfs[k], 300, 40 + k) for k in order])
    for (q, t, d), k, qq in zip(win.match_tensors(qs), order, qs):
        eq, et, ed = co.pipeline(qq, kfs[k], 0.8, True)
        assert np.array_equal(q, eq) and np.array_equal(t, et) and np.array_equal(d, ed)
    dm = win.match(q1)
    assert [m.imgIdx for m in dm[2][:1]] == [2] and [len(x) for x in dm] == [len(o[0]) for o in out]
    with pytest.raises(sx.MatcherError):
        win.put(3, kfs[0][:10])
