"""Multi-GPU parity as a pytest: spawns tools/test_multi_gpu.py under torchrun on all visible GPUs (>= 2) and keeps its
log under gpurun_out/ (copied to profiles/ by hand).  The host logic of the sharded path is covered on CPU by
tests/test_sharded_gloo.py; this is the device side: fused exchange == NCCL == oracle, mixed and growing shards."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_database_parity_on_all_gpus():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2|4|8)")
    n = 8 if n >= 8 else 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "test_multi_gpu.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, f"multi_gpu_parity_n{n}.log"), "w") as f:
        f.write(r.stdout + "\n--- stderr ---\n" + r.stderr[-4000:])
    assert r.returncode == 0 and f"MULTI_GPU_PARITY OK world={n}" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
