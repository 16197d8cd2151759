"""Small end-to-end case for compute-sanitizer (one tool per gpurun call, see B200_PROFILING.md)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from slam_experiments_b200 import _native as nat, synth
from oracle import c_oracle

ok = True
for nq, nt in ((300, 700), (129, 1030), (513, 257)):
    q, t = synth.uniform(nq, nq), synth.uniform(nt, nt + 1)
    qd, td = torch.from_numpy(q).cuda(), torch.from_numpy(t).cuda()
    exp = c_oracle.knn2_keys(q, t)
    for v in ("popc", "i8", "f4"):
        got = nat.knn2_keys(qd, td, variant=v).cpu().numpy().view(np.uint64)
        ok &= bool(np.array_equal(got, exp))
    oq, ot, od, cnt = nat.match_fused(qd[None], td[None], ratio=0.8, cross_check=True, variant="i8")
    eq, et, ed = c_oracle.pipeline(q, t, 0.8, True)
    n = int(cnt[0])
    ok &= n == len(eq) and bool(np.array_equal(oq[0, :n].cpu().numpy(), eq))
    # kind::mxf4 pipeline: ratio test + candidate selection in the k-NN kernel, candidate pass (device-side row count)
    oq, ot, od, cnt = nat.match_fused(qd[None], td[None], ratio=0.8, cross_check=True, variant="f4")
    n = int(cnt[0])
    ok &= n == len(eq) and bool(np.array_equal(oq[0, :n].cpu().numpy(), eq)) and bool(np.array_equal(ot[0, :n].cpu().numpy(), et))
# ORB descriptor stage
import slam_experiments_b200 as sx
from oracle import orb_oracle
from slam_experiments_b200.feature_detectors import keypoint_arrays
img = synth.textured_image(200, 260, 3, 3)
kps, desc = sx.OrbFeatureDetector(n_features=200).detect_and_compute(img, None)
ok &= bool(np.array_equal(desc, orb_oracle.describe(img, *keypoint_arrays(kps))))
torch.cuda.synchronize()
print("SANITIZE_CASE", "OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
