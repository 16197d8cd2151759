"""Multi-GPU parity of the sharded keyframe database (run under torchrun, one rank per GPU):
fused exchange (hm_exchange_merge over symmetric memory) == NCCL all-gather + merge == oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import slam_experiments_b200 as sx
from slam_experiments_b200 import synth
from oracle import c_oracle


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rng = np.random.default_rng(5)
    nkf = 37
    sizes = rng.integers(50, 3000, nkf).tolist()
    kfs = [rng.integers(0, 3, (s, 32), dtype=np.uint8) for s in sizes]            # tie-heavy
    kfs[30][:7] = kfs[2][:7]                                                      # cross-shard duplicates
    cat = np.concatenate(kfs)
    lo, hi, _, _ = sx.shard_ranges(sizes, world)[rank]
    ok = True
    for mode, variant in (("fused", "f4"), ("fused", "i8"), ("fused", "popc"), ("nccl", "f4"), ("nccl", "auto")):
        db = sx.ShardedKeyframeDatabase(sizes, kfs[lo:hi], rank=rank, world_size=world, group=dist.group.WORLD,
                                        exchange=mode, variant=variant)
        if rank == 0:
            print(f"mode {mode} variant {variant}: exchange_mode={db.exchange_mode}", flush=True)
        for it, nq in enumerate((700, 1, 2000, 333, 4096, 129, 700)):
            q = rng.integers(0, 3, (nq, 32), dtype=np.uint8)
            q[: min(nq, 7)] = kfs[2][: min(nq, 7)]
            keys = db.knn2_keys_device(torch.from_numpy(q).cuda()).cpu().numpy().view(np.uint64)
            exp = c_oracle.knn2_keys(q, cat)
            good = np.array_equal(keys, exp)
            ok &= good
            if not good:
                print(f"rank {rank} mode {mode} variant {variant} call {it} nq {nq}: MISMATCH", flush=True)
        rows = db.knnMatch(q, 2)
        ok &= rows[0][0].imgIdx == 2 and rows[0][0].trainIdx == 0 and rows[0][1].imgIdx == 30
    t = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_PARITY", "OK" if int(t) == 1 else "FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t) == 1 else 1)


if __name__ == "__main__":
    main()
