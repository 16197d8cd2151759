"""The C-ABI library loads and exports every symbol include/hm_matcher.h declares.
No compute calls here (no GPU in CI): compute entry points must fail loudly instead."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from slam_experiments_b200 import _native as nat
from slam_experiments_b200 import build as hm_build

HEADER = os.path.join(ROOT, "include", "hm_matcher.h")


@pytest.fixture(scope="module")
def lib():
    hm_build.build()
    return nat.lib()


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"HM_API\s+[\w\s\*]+?\b(hm_\w+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(nat.EXPORTS)


def test_library_exports_every_declared_symbol(lib):
    raw = ctypes.CDLL(nat.SO_PATH)
    for name in declared_symbols():
        assert hasattr(raw, name), f"{name} declared in hm_matcher.h but not exported"


def test_version_and_sizing(lib):
    assert lib.hm_version() == 4
    I8, F4 = nat.VARIANT_I8, nat.VARIANT_F4
    assert lib.hm_prepared_bytes(0, I8) == 0
    assert lib.hm_prepared_bytes(1, I8) == 256 * 256        # padded to whole 256-row tiles
    assert lib.hm_prepared_bytes(257, I8) == 512 * 256
    assert lib.hm_prepared_bytes(257, F4) == 512 * 128      # e2m1: half the bytes
    assert lib.hm_prepared_bytes(257, nat.VARIANT_POPC) == 0   # no prepared form
    tc = lib.hm_default_tensor_variant()
    assert tc in (I8, F4)
    assert lib.hm_prepared_bytes(257, 0) == lib.hm_prepared_bytes(257, tc)
    for v in (0, 1, 2, 3):
        assert lib.hm_workspace_bytes(2000, 2000, 1, v) > 0
    for v in (I8, F4):
        assert lib.hm_workspace_bytes(2000, 8192000, 1, v) >= lib.hm_prepared_bytes(8192000, v)
        assert lib.hm_prepared_workspace_bytes(2000, 8192000, v) > 0
        assert lib.hm_resident_workspace_bytes(2000, 8192000, v) >= lib.hm_prepared_workspace_bytes(2000, 8192000, v)
    # the i8 core expands the query into the workspace, the f4 core inside the k-NN kernel
    assert lib.hm_resident_workspace_bytes(2000, 8192000, I8) >= lib.hm_resident_workspace_bytes(2000, 8192000, F4) + lib.hm_prepared_bytes(2000, I8) - 4096
    assert lib.hm_select_variant(200, 200, 1) == nat.VARIANT_POPC
    assert lib.hm_select_variant(16384, 16384, 1) == tc


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu(lib):
    """No CPU fallback: every compute entry point reports HM_ERR_NO_DEVICE."""
    buf = np.zeros((4, 32), dtype=np.uint8)
    out = np.zeros((4, 2), dtype=np.uint64)
    rc = lib.hm_knn2(buf.ctypes.data, 4, 32, buf.ctypes.data, 4, 32, 0, out.ctypes.data, 0, None, 0, None)
    assert rc == -5 and b"no CUDA device" in lib.hm_last_error()
    assert lib.hm_merge_top2(out.ctypes.data, 1, 2, out.ctypes.data, None) == -5
    assert lib.hm_prepare(buf.ctypes.data, 4, 32, buf.ctypes.data, 2, None) == -5
    h = ctypes.c_void_p()
    assert lib.hm_context_create(ctypes.byref(h)) == -5
    assert lib.hm_device_sm_count() == -5


def test_describe_launch_without_a_device(lib):
    """Launch planning is host arithmetic: it answers without a GPU (148 SMs assumed)."""
    d = nat.describe_launch
    assert d(200, 200).startswith("hm_popc_knn2_kernel")
    c4 = d(2000, 8192000, 1, "f4")
    assert c4.startswith("hm_f4_knn2_floor_kernel grid=(8,") and "cluster=2" in c4      # 8 query blocks, train split
    assert d(2000, 8192000, 1, "i8").startswith("hm_i8_knn2_floor_kernel")
    assert d(2000, 2000, 99).startswith(("hm_f4_knn2_kernel grid=(8,1,99)", "hm_i8_knn2_kernel grid=(8,1,99)"))
    with pytest.raises(nat.NativeError):
        d(0, 5)
