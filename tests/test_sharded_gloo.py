"""N>1 path of the sharded keyframe database on CPU: world_size-2/3 gloo process groups.

The device kernels cannot run here, so the local k-NN and the merge are supplied by the
oracle through the ops hook; what is under test is the product's host logic: shard
planning, global row ids, the all-gather exchange, the merge order and the
(imgIdx, trainIdx) decoding -- against cv2's collection API on the whole database.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from slam_experiments_b200.keyframe_db import ShardedKeyframeDatabase, shard_ranges
from oracle import hamming_oracle as ho


class OracleOps:
    """CPU stand-in for NativeOps (tests only)."""

    def upload(self, a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint8))

    def make_shard(self, train):
        return {"bits": train, "prepared": None, "nt": train.shape[0]}

    def append_rows(self, shard, rows):
        cat = np.concatenate([shard["bits"].numpy(), rows]) if shard["nt"] else rows
        return self.make_shard(torch.from_numpy(np.ascontiguousarray(cat)))

    def local_knn2(self, query, shard, train_base):
        keys = ho.knn2_keys(query.numpy(), shard["bits"].numpy(), train_base=train_base)
        return torch.from_numpy(keys.view(np.int64))

    def all_gather(self, keys, group):
        world = dist.get_world_size(group)
        out = [torch.empty_like(keys) for _ in range(world)]
        dist.all_gather(out, keys.contiguous(), group=group)
        return torch.stack(out)

    def merge(self, gathered):
        return torch.from_numpy(ho.merge_top2_keys(gathered.numpy().view(np.uint64)).view(np.int64))

    def to_host(self, keys):
        return keys.numpy()


def make_db(seed=7, nkf=13):
    rng = np.random.default_rng(seed)
    sizes = rng.integers(2, 40, nkf)
    kfs = [rng.integers(0, 4, (int(s), 32), dtype=np.uint8) for s in sizes]     # tie-heavy
    kfs[-1][:2] = kfs[1][:2]                                                     # cross-shard duplicates
    q = rng.integers(0, 4, (25, 32), dtype=np.uint8)
    q[:3] = kfs[1][:3]
    return sizes, kfs, q


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sizes, kfs, q = make_db()
        lo, hi, _, _ = shard_ranges(sizes, world)[rank]
        db = ShardedKeyframeDatabase(sizes, kfs[lo:hi], rank=rank, world_size=world, ops=OracleOps())
        img, loc, d = db.knn_tensors(q, 2)
        ret[rank] = (img, loc, d)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_matches_whole_database(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    sizes, kfs, q = make_db()
    img, loc, d = ho.collection_knn(q, kfs, 2)
    for r in range(world):
        gi, gl, gd = ret[r]
        assert np.array_equal(gi, img) and np.array_equal(gl, loc) and np.array_equal(gd, d)
    cv2 = pytest.importorskip("cv2")
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    bf.add(kfs)
    rows = bf.knnMatch(q, k=2)
    exp = np.array([[(m.imgIdx, m.trainIdx, int(m.distance)) for m in r] for r in rows])
    assert np.array_equal(exp[:, :, 0], img) and np.array_equal(exp[:, :, 1], loc) and np.array_equal(exp[:, :, 2], d)


def test_append_keyframes_incrementally():
    """Map.insert_keyframe-style growth: the database after N appends == the database built at once."""
    sizes, kfs, q = make_db(seed=3, nkf=6)
    db = ShardedKeyframeDatabase(sizes[:2], kfs[:2], ops=OracleOps())
    for i in range(2, 6):
        assert db.append_keyframe(kfs[i]) == i
    assert db.append_keyframe(np.array([])) == 6                   # a keyframe without features
    img, loc, d = db.knn_tensors(q, 2)
    eimg, eloc, ed = ho.collection_knn(q, kfs + [np.empty((0, 32), np.uint8)], 2)
    assert np.array_equal(img, eimg) and np.array_equal(loc, eloc) and np.array_equal(d, ed)
    cv2 = pytest.importorskip("cv2")
    with pytest.raises(cv2.error):
        db.append_keyframe(np.zeros((3, 16), np.uint8))


def test_single_rank_database_equals_collection_oracle():
    sizes, kfs, q = make_db(seed=11, nkf=5)
    db = ShardedKeyframeDatabase(sizes, kfs, ops=OracleOps())
    img, loc, d = db.knn_tensors(q, 2)
    eimg, eloc, ed = ho.collection_knn(q, kfs, 2)
    assert np.array_equal(img, eimg) and np.array_equal(loc, eloc) and np.array_equal(d, ed)
    rows = db.knnMatch(q, 1)
    assert all(len(r) == 1 for r in rows) and rows[0][0].imgIdx == eimg[0, 0]
