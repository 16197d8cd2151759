"""Where a sharded-database step spends its GPU time (development aid): event stamps between the calls."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from slam_experiments_b200 import _native as nat, synth

nq, nt = 2000, int(sys.argv[1]) if len(sys.argv) > 1 else 1024000
q = torch.from_numpy(synth.uniform(nq, 1)).cuda()
t = torch.from_numpy(synth.uniform(nt, 2)).cuda()
tp = nat.prepare(t)
steps = 200
def ev(): return torch.cuda.Event(enable_timing=True)
E = [[ev() for _ in range(4)] for _ in range(steps)]
K = [(ev(), ev()) for _ in range(steps)]
for _ in range(5):
    nat.knn2_keys_prepared(nat.prepare(q), nq, tp, nt)
torch.cuda.synchronize()
tot0, tot1 = ev(), ev()
tot0.record()
for i in range(steps):
    E[i][0].record()
    qp = nat.prepare(q)
    E[i][1].record()
    nat.profile_events(*K[i])
    keys = nat.knn2_keys_prepared(qp, nq, tp, nt)
    E[i][2].record()
tot1.record()
torch.cuda.synchronize()
nat.profile_events(None, None)
seg = lambda a, b: float(np.mean([E[i][a].elapsed_time(E[i][b]) for i in range(steps)])) * 1e3
print(f"nt={nt}: step {tot0.elapsed_time(tot1) / steps * 1e3:.1f} us | prepare(q) call {seg(0, 1):.1f} us | knn2_prepared call {seg(1, 2):.1f} us "
      f"| main kernel alone {np.mean([a.elapsed_time(b) for a, b in K]) * 1e3:.1f} us | gap between steps "
      f"{np.mean([E[i][2].elapsed_time(E[i + 1][0]) for i in range(steps - 1)]) * 1e3:.1f} us")
# same loop without any event recording inside (pure step time)
tot0.record()
for i in range(steps):
    nat.knn2_keys_prepared(nat.prepare(q), nq, tp, nt)
tot1.record()
torch.cuda.synchronize()
print(f"step without inner events: {tot0.elapsed_time(tot1) / steps * 1e3:.1f} us")
import time
t0 = time.perf_counter()
for i in range(steps):
    nat.knn2_keys_prepared(nat.prepare(q), nq, tp, nt)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"host enqueue time per step: {(t1 - t0) / steps * 1e6:.1f} us")
