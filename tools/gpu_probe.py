"""Device-side timing of the three kernel variants over a few shapes (development aid)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from slam_experiments_b200 import _native as nat, synth


def time_it(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], ts[0]


def kernel_us(fn, iters=10):
    """Device time of the dominant kernel alone (event pair recorded by the library around its launch)."""
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        nat.profile_events(a, b)
        fn()
    torch.cuda.synchronize()
    nat.profile_events(None, None)
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2] * 1e3


def main():
    shapes = [(200, 200), (2000, 2000), (4096, 4096), (8192, 8192), (16384, 16384), (32768, 32768), (65536, 65536),
              (2000, 1024000)]
    print("sm_count", nat.sm_count())
    for nq, nt in shapes:
        q = torch.from_numpy(synth.uniform(nq, 1)).cuda()
        t = torch.from_numpy(synth.uniform(nt, 2)).cuda()
        row = {"nq": nq, "nt": nt}
        for v in ("popc", "i8", "f4"):
            if v == "popc" and nq * nt > 3e9 * 2:
                continue
            med, best = time_it(lambda: nat.knn2_keys(q, t, variant=v), iters=5 if nq * nt > 1e9 else 20)
            kus = kernel_us(lambda: nat.knn2_keys(q, t, variant=v), iters=5 if nq * nt > 1e9 else 10)
            row[v] = {"ms": round(med, 4), "best_ms": round(best, 4), "gpairs": round(nq * nt / med / 1e6, 1),
                      "kernel_us": round(kus, 1), "kernel_gpairs": round(nq * nt / kus / 1e3, 1)}
        if nt >= 65536:
            for v in ("i8", "f4"):
                tp = nat.prepare(t, variant=v)
                def f():
                    nat.knn2_keys_prepared(nat.prepare(q, variant=v), nq, tp, nt, variant=v)
                med, best = time_it(f, iters=5)
                row[v + "_prepared"] = {"ms": round(med, 4), "gpairs": round(nq * nt / med / 1e6, 1)}
                med, best = time_it(lambda: nat.prepare(t, variant=v), iters=5)
                row[v + "_prepare_t_ms"] = round(med, 4)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
