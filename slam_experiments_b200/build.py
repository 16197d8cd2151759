"""Build the C-ABI CUDA library in-tree: slam_experiments_b200/libhm_matcher.so.

nvcc cross-compiles sm_100a without a GPU.  The built .so is git-ignored but
travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libhm_matcher.so")
SOURCES = ["hm_api.cu", "hm_popc.cu", "hm_tc.cu", "hm_epilogue.cu", "hm_orb.cu", "hm_small.cu"]
HEADERS = ["hm_common.cuh", "hm_tcgen05.cuh", "hm_orb_pattern.inc"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def is_stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(ROOT, "include", "hm_matcher.h")]
    return any(os.path.getmtime(d) > t for d in deps)


FAST_SRC = os.path.join(CSRC, "hmfast.c")
FAST_SO = os.path.join(HERE, "_hmfast.so")


def build_hmfast(force: bool = False) -> str:
    """Host-side helper (CPython C API, gcc): bulk cv2.DMatch construction.  Optional: the Python
    side falls back to calling the DMatch type when it is missing."""
    import sysconfig
    if not force and os.path.exists(FAST_SO) and os.path.getmtime(FAST_SO) >= os.path.getmtime(FAST_SRC):
        return FAST_SO
    inc = sysconfig.get_paths()["include"]
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-I", inc, "-o", FAST_SO, FAST_SRC])
    return FAST_SO


def build(force: bool = False, verbose: bool = False) -> str:
    try:
        build_hmfast(force)
    except Exception as e:  # pragma: no cover - optional accelerator
        print(f"warning: _hmfast not built ({e}); DMatch construction falls back to Python", file=sys.stderr)
    trace = os.environ.get("HM_BUILD_TRACE") == "1"
    if trace:          # separate file: load it with HM_MATCHER_SO=<path> (tools/trace_i8.py does)
        so = os.path.join(HERE, "libhm_matcher_trace.so")
        force = True
    else:
        so = SO
    if os.environ.get("HM_BUILD_VARIANT"):
        force = True
    if not force and not is_stale():
        return SO
    cmd = [
        nvcc_path(), "-O3", "-std=c++17", "-lineinfo",
        "-gencode", "arch=compute_100a,code=sm_100a",
        "--cudart", "static", "-shared", "-Xcompiler", "-fPIC,-fvisibility=hidden",
        "-I", os.path.join(ROOT, "include"), "-I", CSRC,
        "-o", so,
    ] + [os.path.join(CSRC, f) for f in SOURCES]
    if trace:      # pipeline-trace build of the tensor-core kernels
        cmd.insert(1, "-DHM_TC_TRACE=" + os.environ.get("HM_BUILD_TRACE_STAMPS", "1"))
        if os.environ.get("HM_BUILD_DEFINE"):          # e.g. HM_BUILD_DEFINE=HM_SPIN_FULL -> libhm_matcher_HM_SPIN_FULL.so
            cmd.insert(1, "-D" + os.environ["HM_BUILD_DEFINE"])
            so = os.path.join(HERE, f"libhm_matcher_{os.environ['HM_BUILD_DEFINE'].split('=')[0]}.so")
            cmd[cmd.index("-o") + 1] = so
        if os.environ.get("HM_BUILD_EXPERIMENT"):
            cmd.insert(1, "-DHM_TC_EXPERIMENT=" + os.environ["HM_BUILD_EXPERIMENT"])
            so = os.path.join(HERE, f"libhm_matcher_exp{os.environ['HM_BUILD_EXPERIMENT']}.so")
            cmd[cmd.index("-o") + 1] = so
    variant = os.environ.get("HM_BUILD_VARIANT")
    if variant and not trace:   # experiment build: HM_BUILD_VARIANT=name HM_BUILD_FLAGS="-DX=1 -DY=2" -> libhm_matcher_<name>.so
        so = os.path.join(HERE, f"libhm_matcher_{variant}.so")
        cmd[cmd.index("-o") + 1] = so
        for f in os.environ.get("HM_BUILD_FLAGS", "").split():
            cmd.insert(1, f)
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return so


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(SO)
