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
    asym = e.copy()
    asym[0, 1] = 2.0
    expect_error("symmetric", lambda: eb.Simulator(None, d, asym, T3, is_mixture=False))


def test_no_cpu_fallback_without_a_device():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    d, e = JC
    with pytest.raises(eb.ElxError) as ei:
        eb.simulateAndFlattenPar(10, d, e, T3, 0)
    assert ei.value.code == L.ELX_E_CUDA and "no CPU fallback" in str(ei.value)
