"""Development aid: the 8-GPU shard shape (2000 x 1,024,000) timed (a) on one resident 131 MB window re-read by every
call (what a rank of the sharded database does) and (b) rotating over eight windows of a 1.05 GB image, so that no call
finds its train tiles in L2.  Separates cache-residency effects from the kernel's own per-CTA costs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from slam_experiments_b200 import _native as nat, synth
v = "f4"
nq, win = 2000, 1024000
q = torch.from_numpy(synth.uniform(nq, 1)).cuda()
t = torch.from_numpy(synth.uniform(8 * win, 2)).cuda()
tp = nat.prepare(t, variant=v)
qp = nat.prepare(q, variant=v)
row_bytes = nat.PREPARED_ROW_BYTES[nat.tensor_variant(v)]
wins = [tp[k * win * row_bytes:] for k in range(8)]
def run(order, n=24):
    for k in order[:4]:
        nat.knn2_keys_prepared(qp, nq, wins[k], win, variant=v)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        nat.knn2_keys_prepared(qp, nq, wins[order[i % len(order)]], win, variant=v)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, order in (("same window", [0]), ("rotating 8 windows", list(range(8))), ("rotating 2 windows", [0, 4])):
    ms = run(order)
    print(f"{name}: {ms:.4f} ms, {ms * 1e-3 * 1.965e9 / (8 * 8000 / 148):.0f} cycles per tile")
