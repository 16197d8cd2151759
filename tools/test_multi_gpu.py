"""Multi-GPU parity of the sharded keyframe database (run under torchrun, one rank per GPU; tests/test_multi_gpu.py
spawns it when >= 2 GPUs are visible): fused exchange (peer stores over symmetric memory, inside the k-NN kernel or
as hm_exchange_merge) == NCCL all-gather + merge == oracle, for homogeneous shards, for MIXED shards (some ranks with a
prepared tensor-core image, some without, one empty) and for a database grown with append_keyframe."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import slam_experiments_b200 as sx
from oracle import c_oracle


def check(db, q, cat, rank, what):
    """device-level call (torch tensors) and host-level call (numpy in, the C host context where it applies)"""
    from slam_experiments_b200 import _native as nat
    expect = c_oracle.knn2_keys(q, cat)
    keys = db.knn2_keys_device(torch.from_numpy(q).cuda()).cpu().numpy().view(np.uint64)
    good = np.array_equal(keys, expect)
    img, loc, d = db.knn_tensors(q, 2)
    gidx, gdist, _ = nat.split_keys(expect)
    eimg, eloc = nat.locate_rows(db.starts, gidx)
    good &= np.array_equal(img, eimg) and np.array_equal(loc, eloc) and np.array_equal(d, gdist)
    if not good:
        print(f"rank {rank} {what} nq {q.shape[0]}: MISMATCH", flush=True)
    return good


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rng = np.random.default_rng(5)
    ok = True

    # ---- homogeneous shards, every variant and both exchange paths -------------------------------------
    nkf = 37
    sizes = rng.integers(50, 3000, nkf).tolist()
    kfs = [rng.integers(0, 3, (s, 32), dtype=np.uint8) for s in sizes]            # tie-heavy
    kfs[30][:7] = kfs[2][:7]                                                      # cross-shard duplicates
    cat = np.concatenate(kfs)
    lo, hi, _, _ = sx.shard_ranges(sizes, world)[rank]
    for mode, variant in (("fused", "f4"), ("fused", "i8"), ("fused", "popc"), ("nccl", "f4"), ("nccl", "auto")):
        db = sx.ShardedKeyframeDatabase(sizes, kfs[lo:hi], rank=rank, world_size=world, group=dist.group.WORLD,
                                        exchange=mode, variant=variant)
        if rank == 0:
            print(f"homogeneous: mode {mode} variant {variant}: exchange_mode={db.exchange_mode}", flush=True)
        for nq in (700, 1, 2000, 333, 4096, 129, 700):
            q = rng.integers(0, 3, (nq, 32), dtype=np.uint8)
            q[: min(nq, 7)] = kfs[2][: min(nq, 7)]
            ok &= check(db, q, cat, rank, f"{mode}/{variant}")
        rows = db.knnMatch(q, 2)
        ok &= rows[0][0].imgIdx == 2 and rows[0][0].trainIdx == 0 and rows[0][1].imgIdx == 30

    # ---- mixed shards under variant="auto": two big keyframes land on rank 0 (>= 65536 rows: prepared image, exchange
    # inside the k-NN kernel), the small ones behind them on the other ranks (packed bits, POPC + hm_exchange_merge),
    # and with world > 3 some ranks are empty.  Both kernels must post and wait on the same flags.
    sizes = [70000, 60000] + [900] * 3
    kfs = [rng.integers(0, 256, (s, 32), dtype=np.uint8) for s in sizes]
    cat = np.concatenate(kfs)
    lo, hi, rlo, rhi = sx.shard_ranges(sizes, world)[rank]
    db = sx.ShardedKeyframeDatabase(sizes, kfs[lo:hi], rank=rank, world_size=world, group=dist.group.WORLD,
                                    exchange="fused", variant="auto")
    kinds = [None] * world
    dist.all_gather_object(kinds, (rhi - rlo, db.shard["prepared"] is not None))
    if rank == 0:
        print(f"mixed: (rows, prepared) per rank = {kinds}", flush=True)
    ok &= len({k[1] for k in kinds}) == 2          # the test is only meaningful if the ranks really differ
    for nq in (2000, 1, 257, 4096, 2000):
        q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
        q[: min(nq, 50)] = cat[rng.choice(cat.shape[0], min(nq, 50), replace=False)]
        ok &= check(db, q, cat, rank, "mixed/auto")

    # ---- growth: small database, the LAST rank's shard crosses the tensor-core threshold through append_keyframe ----
    sizes = [1000] * (2 * world)
    kfs = [rng.integers(0, 256, (s, 32), dtype=np.uint8) for s in sizes]
    lo, hi, _, _ = sx.shard_ranges(sizes, world)[rank]
    db = sx.ShardedKeyframeDatabase(sizes, kfs[lo:hi], rank=rank, world_size=world, group=dist.group.WORLD,
                                    exchange="fused", variant="auto")
    q = rng.integers(0, 256, (2000, 32), dtype=np.uint8)
    ok &= check(db, q, np.concatenate(kfs), rank, "growth/before")
    for s in (40000, 30000):
        extra = rng.integers(0, 256, (s, 32), dtype=np.uint8)
        kfs.append(extra)
        db.append_keyframe(extra)
        ok &= check(db, q, np.concatenate(kfs), rank, f"growth/+{s}")
    kinds = [None] * world
    dist.all_gather_object(kinds, db.shard["prepared"] is not None)
    ok &= kinds[-1] and not kinds[0]
    if rank == 0:
        print(f"growth: prepared per rank = {kinds}", flush=True)

    t = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_PARITY", "OK" if int(t) == 1 else "FAILED", f"world={world}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t) == 1 else 1)


if __name__ == "__main__":
    main()
