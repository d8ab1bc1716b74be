"""ctypes loader for oracle/freerun_spec.c -- executable spec of the product's free-running draw (test infrastructure)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfreerun_spec.so")
_SRC = os.path.join(_HERE, "freerun_spec.c")
_lib = None


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", _SO, _SRC])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.spec_freerun.restype = C.c_int
        _lib.spec_freerun_pair.restype = C.c_int
        _lib.spec_freerun_quad.restype = C.c_int
        _lib.spec_quad_ties.restype = C.c_int64
        _lib.spec_freerun_duo.restype = C.c_int
        _lib.spec_duo_ties.restype = C.c_int64
    return _lib


def philox4x32(ctr, key, rounds=10):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().spec_philox4x32(rounds, c, k, o)
    return [int(x) for x in o]


def alias_bits(k):
    a = 1
    while (1 << a) < k:
        a += 1
    return max(a, 2)


def freerun_pair(is_mixture, parent, leaf_row, ws, ds, pair, seed, site0, nsites, rounds=10, keep_internal=False):
    """Pair mode (K <= 4): pair = (node_a[G], node_b[G], pe16[G][C][K][16], pqlo[G][C][K][16]) in walk order."""
    parent = np.ascontiguousarray(parent, dtype=np.int32)
    leaf_row = np.ascontiguousarray(leaf_row, dtype=np.int32)
    node_a = np.ascontiguousarray(pair[0], dtype=np.int32)
    node_b = np.ascontiguousarray(pair[1], dtype=np.int32)
    pe16 = np.ascontiguousarray(pair[2], dtype=np.uint16)
    pqlo = np.ascontiguousarray(pair[3], dtype=np.uint32)
    g, c, k, cols = pe16.shape
    assert cols == 16 and pqlo.shape == pe16.shape and node_a.shape == (g,) and node_b.shape == (g,)
    n = len(parent)
    cumw = np.cumsum(np.asarray(ws, dtype=np.float64)) if is_mixture else np.ones(1)
    cumpi = np.ascontiguousarray(np.cumsum(np.asarray(ds, dtype=np.float64).reshape(c, k), axis=1))
    t = int((leaf_row >= 0).sum())
    leaves = np.zeros((t, max(nsites, 1)), dtype=np.uint8)
    comps = np.zeros(max(nsites, 1), dtype=np.int32)
    internal = np.zeros((n, max(nsites, 1)), dtype=np.uint8) if keep_internal else None
    vp = C.c_void_p
    rc = lib().spec_freerun_pair(n, k, c, int(bool(is_mixture)), rounds, vp(parent.ctypes.data), vp(leaf_row.ctypes.data),
                                 vp(cumw.ctypes.data), vp(cumpi.ctypes.data), g, vp(node_a.ctypes.data), vp(node_b.ctypes.data),
                                 vp(pe16.ctypes.data), vp(pqlo.ctypes.data), C.c_uint64(seed & (2 ** 64 - 1)), C.c_int64(site0),
                                 C.c_int64(nsites), vp(leaves.ctypes.data), C.c_int64(max(nsites, 1)), vp(comps.ctypes.data),
                                 vp(internal.ctypes.data) if keep_internal else None, C.c_int64(max(nsites, 1)))
    if rc:
        raise RuntimeError("spec_freerun_pair rc=%d" % rc)
    return leaves[:, :nsites], comps[:nsites], (internal[:, :nsites] if keep_internal else None)


def freerun(is_mixture, parent, leaf_row, ws, ds, e16, qlo, seed, site0, nsites, rounds=10, keep_internal=False):
    parent = np.ascontiguousarray(parent, dtype=np.int32)
    leaf_row = np.ascontiguousarray(leaf_row, dtype=np.int32)
    e16 = np.ascontiguousarray(e16, dtype=np.uint16)
    qlo = np.ascontiguousarray(qlo, dtype=np.uint32)
    n, c, k, kp = e16.shape
    cumw = np.cumsum(np.asarray(ws, dtype=np.float64)) if is_mixture else np.ones(1)
    cumpi = np.ascontiguousarray(np.cumsum(np.asarray(ds, dtype=np.float64).reshape(c, k), axis=1))
    t = int((leaf_row >= 0).sum())
    leaves = np.zeros((t, max(nsites, 1)), dtype=np.uint8)
    comps = np.zeros(max(nsites, 1), dtype=np.int32)
    internal = np.zeros((n, max(nsites, 1)), dtype=np.uint8) if keep_internal else None
    vp = C.c_void_p
    rc = lib().spec_freerun(n, k, c, int(bool(is_mixture)), alias_bits(k), rounds, vp(parent.ctypes.data), vp(leaf_row.ctypes.data),
                            vp(cumw.ctypes.data), vp(cumpi.ctypes.data), vp(e16.ctypes.data), vp(qlo.ctypes.data),
                            C.c_uint64(seed & (2 ** 64 - 1)), C.c_int64(site0), C.c_int64(nsites), vp(leaves.ctypes.data),
                            C.c_int64(max(nsites, 1)), vp(comps.ctypes.data), vp(internal.ctypes.data) if keep_internal else None,
                            C.c_int64(max(nsites, 1)))
    if rc:
        raise RuntimeError("spec_freerun rc=%d" % rc)
    return leaves[:, :nsites], comps[:nsites], (internal[:, :nsites] if keep_internal else None)


QUAD_REC_INTS = 52


def freerun_quad(is_mixture, parent, leaf_row, ws, ds, quad, seed, site0, nsites, rounds=10):
    """Quad mode (K <= 4, leaves only): quad = (recs[G][52], qe32[G][C][4][256], qres32[G][C][4][256]) in walk order."""
    leaf_row = np.ascontiguousarray(leaf_row, dtype=np.int32)
    recs = np.ascontiguousarray(quad[0], dtype=np.int32)
    qe32 = np.ascontiguousarray(quad[1], dtype=np.uint32)
    qres32 = np.ascontiguousarray(quad[2], dtype=np.uint32)
    g, c, four, cols = qe32.shape
    assert four == 4 and cols == 256 and qres32.shape == qe32.shape and recs.shape == (g, QUAD_REC_INTS)
    n = len(leaf_row)
    k = np.asarray(ds).reshape(c, -1).shape[1]
    cumw = np.cumsum(np.asarray(ws, dtype=np.float64)) if is_mixture else np.ones(1)
    cumpi = np.ascontiguousarray(np.cumsum(np.asarray(ds, dtype=np.float64).reshape(c, k), axis=1))
    t = int((leaf_row >= 0).sum())
    leaves = np.zeros((t, max(nsites, 1)), dtype=np.uint8)
    comps = np.zeros(max(nsites, 1), dtype=np.int32)
    vp = C.c_void_p
    rc = lib().spec_freerun_quad(n, k, c, int(bool(is_mixture)), rounds, vp(leaf_row.ctypes.data), vp(cumw.ctypes.data),
                                 vp(cumpi.ctypes.data), g, vp(recs.ctypes.data), vp(qe32.ctypes.data), vp(qres32.ctypes.data),
                                 C.c_uint64(seed & (2 ** 64 - 1)), C.c_int64(site0), C.c_int64(nsites), vp(leaves.ctypes.data),
                                 C.c_int64(max(nsites, 1)), vp(comps.ctypes.data))
    if rc:
        raise RuntimeError("spec_freerun_quad rc=%d" % rc)
    return leaves[:, :nsites], comps[:nsites], None


def quad_ties():
    """Number of escape-level draws (t == 0) the last freerun_quad call took."""
    return int(lib().spec_quad_ties())


def freerun_duo(is_mixture, parent, leaf_row, ws, ds, duo, seed, site0, nsites, rounds=10):
    """Duo mode (5 <= K <= 22, leaves only): duo = (recs[G][52], de32[G][C][K][512], dres32[G][C][K][512]) in walk order."""
    leaf_row = np.ascontiguousarray(leaf_row, dtype=np.int32)
    recs = np.ascontiguousarray(duo[0], dtype=np.int32)
    de32 = np.ascontiguousarray(duo[1], dtype=np.uint32)
    dres32 = np.ascontiguousarray(duo[2], dtype=np.uint32)
    g, c, k, cols = de32.shape
    assert cols == 512 and dres32.shape == de32.shape and recs.shape == (g, QUAD_REC_INTS)
    n = len(leaf_row)
    assert np.asarray(ds).reshape(c, -1).shape[1] == k
    cumw = np.cumsum(np.asarray(ws, dtype=np.float64)) if is_mixture else np.ones(1)
    cumpi = np.ascontiguousarray(np.cumsum(np.asarray(ds, dtype=np.float64).reshape(c, k), axis=1))
    t = int((leaf_row >= 0).sum())
    leaves = np.zeros((t, max(nsites, 1)), dtype=np.uint8)
    comps = np.zeros(max(nsites, 1), dtype=np.int32)
    vp = C.c_void_p
    rc = lib().spec_freerun_duo(n, k, c, int(bool(is_mixture)), rounds, vp(leaf_row.ctypes.data), vp(cumw.ctypes.data),
                                vp(cumpi.ctypes.data), g, vp(recs.ctypes.data), vp(de32.ctypes.data), vp(dres32.ctypes.data),
                                C.c_uint64(seed & (2 ** 64 - 1)), C.c_int64(site0), C.c_int64(nsites), vp(leaves.ctypes.data),
                                C.c_int64(max(nsites, 1)), vp(comps.ctypes.data))
    if rc:
        raise RuntimeError("spec_freerun_duo rc=%d" % rc)
    return leaves[:, :nsites], comps[:nsites], None


def duo_ties():
    """Number of escape-level draws (t == 0) the last freerun_duo call took."""
    return int(lib().spec_duo_ties())
