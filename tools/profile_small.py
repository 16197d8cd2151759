"""Development aid: host-side profile of one small match() call (C1: 200 x 200)."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cv2
import slam_experiments_b200 as sx
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "c1_orb200.npz"))
q, t = np.ascontiguousarray(g["query"]), np.ascontiguousarray(g["train"])
m = sx.BruteForceFeatureMatcher(cv2.NORM_HAMMING)
for name, fn in (("match", lambda: m.match(t, q)), ("match_tensors", lambda: m.match_tensors(t, q))):
    for _ in range(20):
        fn()
    t0 = time.perf_counter()
    for _ in range(2000):
        fn()
    print(f"{name}: {(time.perf_counter() - t0) / 2000 * 1e6:.1f} us per call")
pr = cProfile.Profile(); pr.enable()
for _ in range(2000):
    m.match_tensors(t, q)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(12); print(s.getvalue()[:2600])
ref = cv2.BFMatcher(cv2.NORM_HAMMING)
t0 = time.perf_counter()
for _ in range(2000):
    ref.match(q, t)
print(f"cv2: {(time.perf_counter() - t0) / 2000 * 1e6:.1f} us per call")
