"""Development aid: dump the per-tile pipeline time stamps of CTA 0 of the tensor-core kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the trace stamps are compiled in only with -DHM_TC_TRACE=1: HM_BUILD_TRACE=1 python -m slam_experiments_b200.build
os.environ["HM_MATCHER_SO"] = os.environ.get("HM_TRACE_SO") or os.path.join(ROOT, "slam_experiments_b200", "libhm_matcher_trace.so")
if not os.path.exists(os.environ["HM_MATCHER_SO"]):
    import subprocess
    subprocess.check_call([sys.executable, "-m", "slam_experiments_b200.build"], cwd=ROOT,
                          env=dict(os.environ, HM_BUILD_TRACE="1"))
os.environ["HM_I8_TRACE"] = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace_i8.txt"
import torch
from slam_experiments_b200 import _native as nat, synth
nq, nt = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (65536, 65536)
q = torch.from_numpy(synth.uniform(nq, 1)).cuda()
t = torch.from_numpy(synth.uniform(nt, 2)).cuda()
os.environ.pop("HM_I8_TRACE")
for _ in range(2):
    nat.knn2_keys(q, t, variant=os.environ.get("HM_TRACE_VARIANT", "i8"))
torch.cuda.synchronize()
os.environ["HM_I8_TRACE"] = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace_i8.txt"
nat.knn2_keys(q, t, variant=os.environ.get("HM_TRACE_VARIANT", "i8"))
torch.cuda.synchronize()
print(open(os.environ["HM_I8_TRACE"]).read())
