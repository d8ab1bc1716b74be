"""ctypes loader for oracle/sim.c (CPU ORACLE -- test infrastructure only, not the product)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_sim.so")
_SRC = os.path.join(_HERE, "sim.c")
_lib = None


def build(force: bool = False) -> str:
    """gcc -O2 -ffp-contract=off (no FMA: the cumulative sums must be plain sequential adds)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-o", _SO, _SRC, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        dp, ip, bp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
        _lib.oracle_simulate_injected.restype = C.c_int
        _lib.oracle_simulate_injected.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, ip, ip, dp, dp, dp, C.c_int64, dp, bp, ip, bp]
        _lib.oracle_simulate_splitmix.restype = C.c_int
        _lib.oracle_simulate_splitmix.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, ip, ip, dp, dp, dp, C.c_int, C.c_int64,
                                                  C.c_uint64, C.c_int, C.c_int, bp, ip, bp]
        _lib.oracle_splitmix_first_word.restype = C.c_uint64
        _lib.oracle_splitmix_first_word.argtypes = [C.c_uint64]
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _prep(parent, leaf_row, ws, ds, p):
    parent = np.ascontiguousarray(parent, dtype=np.int32)
    leaf_row = np.ascontiguousarray(leaf_row, dtype=np.int32)
    ws = np.ascontiguousarray(ws, dtype=np.float64)
    ds = np.ascontiguousarray(ds, dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.float64)
    n, c, k, k2 = p.shape
    assert k == k2 and ds.shape == (c, k) and ws.shape == (c,) and parent.shape == (n,) and leaf_row.shape == (n,)
    return parent, leaf_row, ws, ds, p, n, c, k


def simulate_injected(is_mixture, parent, leaf_row, ws, ds, p, u, keep_internal=False):
    """(comps[S] int32, leaves[T][S] uint8, internal[N][S] uint8 | None) for the uniform matrix u."""
    parent, leaf_row, ws, ds, p, n, c, k = _prep(parent, leaf_row, ws, ds, p)
    u = np.ascontiguousarray(u, dtype=np.float64)
    s = u.shape[1]
    assert u.shape[0] == n + (2 if is_mixture else 1)
    t = int((leaf_row >= 0).sum())
    leaves = np.zeros((t, s), dtype=np.uint8)
    comps = np.zeros(s, dtype=np.int32)
    internal = np.zeros((n, s), dtype=np.uint8) if keep_internal else None
    rc = lib().oracle_simulate_injected(n, k, c, int(bool(is_mixture)), _p(parent, C.c_int32), _p(leaf_row, C.c_int32),
                                        _p(ws, C.c_double), _p(ds, C.c_double), _p(p, C.c_double), s, _p(u, C.c_double),
                                        _p(leaves, C.c_uint8), _p(comps, C.c_int32), _p(internal, C.c_uint8))
    if rc != 0:
        raise RuntimeError("categorical: bad weights! (rc=%d)" % rc)
    return comps, leaves, internal


def simulate_splitmix(is_mixture, parent, leaf_row, ws, ds, p, s, seed, nchunks=0, nthreads=1, keep_internal=False,
                      precomputed_cum=False):
    """Free-running walk on the reference's SplitMix stream; nchunks=0 -> serial API, >=1 -> *Par API."""
    parent, leaf_row, ws, ds, p, n, c, k = _prep(parent, leaf_row, ws, ds, p)
    t = int((leaf_row >= 0).sum())
    leaves = np.zeros((t, s), dtype=np.uint8)
    comps = np.zeros(s, dtype=np.int32)
    internal = np.zeros((n, s), dtype=np.uint8) if keep_internal else None
    rc = lib().oracle_simulate_splitmix(n, k, c, int(bool(is_mixture)), _p(parent, C.c_int32), _p(leaf_row, C.c_int32),
                                        _p(ws, C.c_double), _p(ds, C.c_double), _p(p, C.c_double), int(precomputed_cum), s,
                                        C.c_uint64(seed & (2 ** 64 - 1)), nchunks, nthreads, _p(leaves, C.c_uint8),
                                        _p(comps, C.c_int32), _p(internal, C.c_uint8))
    if rc != 0:
        raise RuntimeError("categorical: bad weights! (rc=%d)" % rc)
    return comps, leaves, internal
