"""TEST INFRASTRUCTURE (never imported by the product): restatement of the reference's birth-death point process and of the
tree it is turned into -- what `tlynx simulate -n N birthdeath -l lambda -m mu` does for the sub-critical process without
a given height:

  elynx-tree/src/ELynx/Tree/Distribution/TimeOfOrigin.hs   cumulative :52-59, density :66-75, quantile :79-92
  elynx-tree/src/ELynx/Tree/Distribution/BirthDeath.hs     cumulative :54-62, density :69-78, quantile :82-95
  elynx-tree/src/ELynx/Tree/Simulate/PointProcess.hs       simulateRandom :119-148, simulateOrigin :151-177,
                                                           sortPP / flattenIndices :201-219, toReconstructedTree :271-317

parity unpinned by the reference (its test-suite holds no known answer for the point process, elynx-tree/test/**): the
restatement is pinned (a) by the algebraic identities quantile(cumulative(x)) = x and d/dx cumulative = density, checked in
tests/test_birthdeath_cpu.py, and (b) the product's generator is held to it bit for bit on the same uniform stream.

The reference draws its uniforms from the caller's StdGen through `genContVar` = `quantile d <$> uniformDouble01M`; the
product's generator draws them from its own SplitMix64 stream (`elx_model.cpp` birth_death): SplitMix64 below restates
THAT stream so that both sides consume the same uniforms -- like the injected-uniforms mode of the simulator."""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

MASK = (1 << 64) - 1


class SplitMix64:
    """The product's uniform stream (Steele/Lea/Flood SplitMix64, 53-bit mantissa + 0.5 ulp: doubles in (0,1))."""

    def __init__(self, seed: int):
        self.s = seed & MASK

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
        return z ^ (z >> 31)

    def uniform(self) -> float:
        return (float(self.next() >> 11) + 0.5) * 2.0 ** -53


# ---- TimeOfOrigin.hs ---------------------------------------------------------------------------------------------
def origin_cumulative(n: int, l: float, m: float, x: float) -> float:
    if x <= 0:
        return 0.0
    d = l - m
    te = l * (1.0 - math.exp(-d * x)) / (l - m * math.exp(-d * x))
    return te ** n


def origin_density(n: int, l: float, m: float, x: float) -> float:
    if x < 0:
        return 0.0
    d = l - m
    ex = math.exp(-d * x)
    return n * l ** n * d ** 2 * (1.0 - ex) ** (n - 1.0) * ex / ((l - m * ex) ** (n + 1.0))


def origin_quantile(n: int, l: float, m: float, p: float) -> float:
    if not (0 <= p <= 1):
        raise ValueError("PointProcess.quantile: p must be in [0,1] range. Got: %r." % p)
    d = l - m
    pn = math.pow(p, 1.0 / n)
    return -1.0 / d * math.log(l * (1.0 - pn) / (l - pn * m))


# ---- BirthDeath.hs -----------------------------------------------------------------------------------------------
def bd_cumulative(t: float, l: float, m: float, x: float) -> float:
    if x <= 0:
        return 0.0
    if x > t:
        return 1.0
    d = l - m
    t1 = (1.0 - math.exp(-d * x)) / (l - m * math.exp(-d * x))
    t2 = (l - m * math.exp(-d * t)) / (1.0 - math.exp(-d * t))
    return t1 * t2


def bd_density(t: float, l: float, m: float, x: float) -> float:
    if x < 0 or x > t:
        return 0.0
    d = l - m
    t1 = math.exp(-d * x) / ((l - m * math.exp(-d * x)) ** 2)
    t2 = (l - m * math.exp(-d * t)) / (1.0 - math.exp(-d * t))
    return d ** 2 * t1 * t2


def bd_quantile(t: float, l: float, m: float, p: float) -> float:
    if not (0 <= p <= 1):
        raise ValueError("BirthDeath.quantile: p must be in range [0,1] but got %r." % p)
    d = l - m
    t2 = (l - m * math.exp(-d * t)) / (1.0 - math.exp(-d * t))
    return (-1.0 / d) * math.log((1.0 - p * l / t2) / (1.0 - p * m / t2))


# ---- PointProcess.hs ---------------------------------------------------------------------------------------------
def simulate_random(n: int, l: float, m: float, gen: SplitMix64) -> Tuple[List[float], float]:
    """simulateRandom (:119-148) for the sub-critical process: (merge values vs[n-1], origin)."""
    if n < 1:
        raise ValueError("Number of samples needs to be one or larger.")
    if l < 0.0:
        raise ValueError("Birth rate needs to be positive.")
    if m > l:
        raise ValueError("simulateRandom: Please specify height if mu > lambda.")
    t = origin_quantile(n, l, m, gen.uniform())
    return [bd_quantile(t, l, m, gen.uniform()) for _ in range(n - 1)], t  # simulateOrigin (:151-177)


def _flatten_indices(idx: Sequence[int]) -> List[int]:
    """flattenIndices / fAcc (:208-219): every index minus the number of EARLIER indices below it (Fenwick tree instead of
    the reference's quadratic filter)."""
    n = max(idx) + 2 if idx else 1
    tree = [0] * (n + 1)
    out = []
    for i in idx:
        c, j = 0, i  # earlier indices < i  = prefix count over [0, i)
        while j > 0:
            c += tree[j]
            j -= j & -j
        out.append(i - c)
        j = i + 1
        while j <= n:
            tree[j] += 1
            j += j & -j
    return out


def to_reconstructed_tree(vs: Sequence[float], origin: float):
    """toReconstructedTree (:271-317): leaves 0..n-1 left to right; the sorted merge values join neighbouring subtrees.
    Returns the preorder flattening (parent, brlen, leaf_row, names) the product's FlatTree uses."""
    if len(vs) <= 1:
        raise ValueError("Too few values.")
    order = sorted(range(len(vs)), key=lambda i: (vs[i], i))  # sortListWithIndices: stable ascending
    vs_sorted = [vs[i] for i in order]
    is_sorted = _flatten_indices(order)
    trs = [[0.0, ("leaf", p)] for p in range(len(vs) + 1)]  # [stem length, node]
    hs = [0.0] * (len(vs) + 1)
    for i, v in zip(is_sorted, vs_sorted):
        hl, hr = hs[i], hs[i + 1]
        tl, tr = trs[i], trs[i + 1]
        tl[0] += v - hl
        tr[0] += v - hr
        trs[i:i + 2] = [[0.0, ("node", tl, tr)]]
        hs[i:i + 2] = [hl + (v - hl)]
    root = trs[0]
    root[0] += origin - vs_sorted[-1]
    parent, brlen, leaf_row, names = [], [], [], []
    stack = [(root, -1)]
    while stack:
        (stem, node), par = stack.pop()
        me = len(parent)
        parent.append(par)
        brlen.append(stem)
        if node[0] == "leaf":
            leaf_row.append(node[1])
            names.append(str(node[1]))
        else:
            leaf_row.append(-1)
            names.append("0")
            stack.append((node[2], me))
            stack.append((node[1], me))
    return np.array(parent, dtype=np.int32), np.array(brlen), np.array(leaf_row, dtype=np.int32), names


def birth_death_tree(n: int, l: float, m: float, seed: int):
    """simulateReconstructedTree n Random l m (:245-259) on the product's uniform stream."""
    vs, t = simulate_random(n, l, m, SplitMix64(seed))
    return to_reconstructed_tree(vs, t)


def node_heights(parent, brlen, leaf_row):
    """heights above the leaves of the internal nodes (= the merge values) and the origin, from a flat ultrametric tree"""
    n = len(parent)
    depth = np.zeros(n)
    for i in range(n):
        depth[i] = brlen[i] + (depth[parent[i]] if parent[i] >= 0 else 0.0)
    total = depth[np.asarray(leaf_row) >= 0].max()
    inner = np.asarray(leaf_row) < 0
    return total - depth[inner], total
