"""ctypes loader for oracle/libhamming_oracle.so (the C restatement).

TEST INFRASTRUCTURE ONLY - see oracle/hamming_oracle.py for the policy.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhamming_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "hamming_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        L.oracle_num_threads.restype = ctypes.c_int
        for fn in (L.oracle_knn2_keys, L.oracle_colmin_keys):
            fn.restype = None
        L.oracle_knn2_keys.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                       ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                       ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int]
        L.oracle_colmin_keys.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                         ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                         ctypes.c_void_p, ctypes.c_int]
        _lib = L
    return _lib


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def _prep(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 2 and a.shape[1] == 32
    return a


def knn2_keys(query, train, train_base: int = 0, threads: int = 0) -> np.ndarray:
    q, t = _prep(query), _prep(train)
    out = np.empty((q.shape[0], 2), dtype=np.uint64)
    lib().oracle_knn2_keys(q.ctypes.data, q.shape[0], 32, t.ctypes.data, t.shape[0], 32,
                           train_base, out.ctypes.data, threads)
    return out


def colmin_keys(query, train, threads: int = 0) -> np.ndarray:
    q, t = _prep(query), _prep(train)
    out = np.empty(t.shape[0], dtype=np.uint64)
    lib().oracle_colmin_keys(q.ctypes.data, q.shape[0], 32, t.ctypes.data, t.shape[0], 32,
                             out.ctypes.data, threads)
    return out


def pipeline(query, train, ratio=0.75, cross_check=True, threads: int = 0):
    """knn2 -> Lowe ratio (integer LUT) -> mutual check, as (q, t, d) arrays."""
    from .hamming_oracle import ratio_lut, NO_MATCH_KEY
    keys = knn2_keys(query, train, threads=threads)
    nq = keys.shape[0]
    d1 = (keys[:, 0] >> np.uint64(32)).astype(np.int64)
    t1 = (keys[:, 0] & np.uint64(0xFFFFFFFF)).astype(np.int64)
    keep = keys[:, 0] != NO_MATCH_KEY
    if ratio is not None:
        has2 = keys[:, 1] != NO_MATCH_KEY
        d2 = np.where(has2, (keys[:, 1] >> np.uint64(32)).astype(np.int64), 0)
        keep &= has2 & (d1 < ratio_lut(ratio)[np.clip(d2, 0, 256)])
    if cross_check and np.asarray(train).shape[0]:
        col = colmin_keys(query, train, threads=threads)
        bq = (col & np.uint64(0xFFFFFFFF)).astype(np.int64)
        keep &= bq[np.clip(t1, 0, len(bq) - 1)] == np.arange(nq)
    qi = np.nonzero(keep)[0]
    return qi.astype(np.int32), t1[qi].astype(np.int32), d1[qi].astype(np.int32)
