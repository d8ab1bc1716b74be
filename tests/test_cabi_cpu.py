"""CPU-side checks of the boundary: the C-ABI library loads, exports every declared symbol, validates its
inputs the way the reference does (`error` -> status + message), and never silently falls back to the CPU."""
import os
import re

import numpy as np
import pytest

import elynx_b200 as eb
from elynx_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    from elynx_b200 import build

    build.build()


def declared_symbols(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(elx_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = L.lib()
    syms = declared_symbols("elynx_b200.h")
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), s
        assert s in L.SIGNATURES, "no ctypes signature for " + s
    assert lib.elx_version() >= 100


class T3:
    parent = np.array([-1, 0, 0], dtype=np.int32)
    brlen = np.array([0.0, 1.0, 1.0])
    leaf_row = np.array([-1, 0, 1], dtype=np.int32)


JC = (np.full(4, 0.25), np.ones((4, 4)) - np.eye(4))


def expect_error(match, fn):
    with pytest.raises(eb.ElxError, match=match):
        fn()


def test_validation_matches_reference_errors():
    d, e = JC

    class Neg(T3):
        brlen = np.array([0.0, -1.0, 1.0])

    expect_error("Time is negative", lambda: eb.Simulator(None, d, e, Neg, is_mixture=False))  # MarkovProcess.hs:37
    expect_error("not square", lambda: eb.Simulator(None, d, np.ones((4, 3)), T3, is_mixture=False))  # MarkovProcess.hs:35

    class NotPre(T3):
        parent = np.array([-1, 2, 0], dtype=np.int32)

    expect_error("preorder", lambda: eb.Simulator(None, d, e, NotPre, is_mixture=False))

    class BadLeaf(T3):
        leaf_row = np.array([-1, 0, 0], dtype=np.int32)

    expect_error("permutation", lambda: eb.Simulator(None, d, e, BadLeaf, is_mixture=False))
    expect_error("Number of weights", lambda: eb.Simulator(np.ones(3), np.stack([d, d]), np.stack([e, e]), T3))  # MixtureModel.hs:86
    expect_error("bad weights", lambda: eb.Simulator(np.zeros(2), np.stack([d, d]), np.stack([e, e]), T3))
    big = np.ones((40, 40)) - np.eye(40)
    big[0, 1] = 2.0  # non-reversible models take the Pade path, which stops at 32 states
    expect_error("at most 32 states", lambda: eb.Simulator(None, np.full(40, 1 / 40), big, T3, is_mixture=False))


def test_no_cpu_fallback_without_a_device():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    d, e = JC
    with pytest.raises(eb.ElxError) as ei:
        eb.simulateAndFlattenPar(10, d, e, T3, 0)
    assert ei.value.code == L.ELX_E_CUDA and "no CPU fallback" in str(ei.value)


def _walk_info(parent, leaf_row):
    import ctypes as C

    parent = np.ascontiguousarray(parent, dtype=np.int32)
    leaf_row = np.ascontiguousarray(leaf_row, dtype=np.int32)
    brlen = np.zeros(len(parent))
    t = L.elx_tree(len(parent), parent.ctypes.data_as(C.POINTER(C.c_int32)), brlen.ctypes.data_as(C.POINTER(C.c_double)),
                   leaf_row.ctypes.data_as(C.POINTER(C.c_int32)))
    nl, ns = C.c_int32(), C.c_int32()
    order = np.empty(len(parent), dtype=np.int32)
    L.check(L.lib().elx_tree_walk_info(C.byref(t), C.byref(nl), C.byref(ns), C.c_void_p(order.ctypes.data)))
    return nl.value, ns.value, order


def test_walk_program_slot_bound():
    """Heavy-child-last walk: parents before children, every node once, and at most floor(log2 N)+1 live slots for ANY
    tree -- ladders of 10^4 taxa included (no HBM scratch needed for deep trees)."""
    import math

    from elynx_b200 import model as pm

    def ladder(n, left):  # caterpillar with the internal child first (left) or last (right), in preorder
        m = n - 1  # internal nodes
        if left:
            parent = [-1] + list(range(m - 1)) + [m - 1, m - 1] + list(range(m - 2, -1, -1))
            leaf_row = [-1] * m + list(range(n))
        else:
            parent, leaf_row = [-1], [-1]
            cur = 0
            for i in range(m - 1):
                parent += [cur, cur]
                leaf_row += [i, -1]
                cur = len(parent) - 1
            parent += [cur, cur]
            leaf_row += [m - 1, m]
        return np.array(parent), np.array(leaf_row)

    cases = [ladder(10000, True), ladder(10000, False)]
    for n, seed in ((1000, 1), (10000, 3), (3, 0), (2, 0)):
        t = pm.birth_death_tree(n, 1.0, 0.5, seed=seed)
        cases.append((t.parent, t.leaf_row))
    cases.append((np.array([-1]), np.array([0])))  # single node
    for parent, leaf_row in cases:
        n = len(parent)
        nl, ns, order = _walk_info(parent, leaf_row)
        assert nl == int((np.asarray(leaf_row) >= 0).sum())
        assert sorted(order.tolist()) == list(range(n))
        pos = np.empty(n, dtype=np.int64)
        pos[order] = np.arange(n)
        assert all(pos[parent[i]] < pos[i] for i in range(1, n))
        assert ns <= int(math.floor(math.log2(n))) + 1 if n > 1 else ns == 1, (n, ns)
