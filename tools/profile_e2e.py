"""Development aid: where the host time of the C4 end-to-end call goes (cProfile of knnMatch / knn_tensors)."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import slam_experiments_b200 as sx
from slam_experiments_b200 import synth

nkf, rows, nq = 4096, 2000, 2000
q, t = synth.keyframe_database(nkf, rows, nq, seed=4096)
sizes = [rows] * nkf
db = sx.ShardedKeyframeDatabase(sizes, [t[i * rows:(i + 1) * rows] for i in range(nkf)])
for fn_name in ("knn_tensors", "knnMatch"):
    fn = getattr(db, fn_name)
    for _ in range(5):
        fn(q, 2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        fn(q, 2)
    dt = (time.perf_counter() - t0) / 50
    print(f"{fn_name}: {dt * 1e3:.3f} ms per call")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(50):
        fn(q, 2)
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14)
    print(s.getvalue()[:3500])
