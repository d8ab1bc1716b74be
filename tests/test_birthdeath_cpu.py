"""f3: the birth-death tree generator (host C++, elx_model.cpp) against the oracle restatement of the
reference's point process (oracle/birthdeath.py <- PointProcess.hs, BirthDeath.hs, TimeOfOrigin.hs).

  * the restatement is pinned by the reference formulas' own identities (quantile inverts cumulative, density is its
    derivative, cumulative(origin) = 1);
  * the product's trees equal the restatement's on the same uniform stream (topology exactly, branch lengths to the last bit or two) (structure by the reference's
    sort-and-join algorithm, the product builds a Cartesian tree);
  * distributional: the merge heights of generated trees follow BirthDeath.cumulative given the origin, and the origins
    over many seeds follow TimeOfOrigin.cumulative (Kolmogorov-Smirnov against the reference's CDFs)."""
import numpy as np
import pytest

import oracle.birthdeath as bd
from elynx_b200 import model as pm


@pytest.mark.parametrize("n,l,m", [(10, 1.0, 0.5), (1000, 1.0, 0.5), (50, 2.0, 0.3), (5, 1.0, 0.9)])
def test_restated_distributions_are_consistent(n, l, m):
    for p in (1e-9, 0.01, 0.3, 0.5, 0.9, 0.999999):
        t = bd.origin_quantile(n, l, m, p)
        assert abs(bd.origin_cumulative(n, l, m, t) - p) < 1e-9 * max(p, 1e-3)
        x = bd.bd_quantile(t, l, m, p)
        assert 0 <= x <= t
        assert abs(bd.bd_cumulative(t, l, m, x) - p) < 1e-9
    t = bd.origin_quantile(n, l, m, 0.5)
    assert abs(bd.bd_cumulative(t, l, m, t) - 1.0) < 1e-12 and bd.bd_cumulative(t, l, m, 0.0) == 0.0
    for x in np.linspace(0.05 * t, 0.95 * t, 7):
        h = 1e-6 * t
        num = (bd.bd_cumulative(t, l, m, x + h) - bd.bd_cumulative(t, l, m, x - h)) / (2 * h)
        assert abs(num - bd.bd_density(t, l, m, x)) < 1e-6 * max(1.0, num)
        num = (bd.origin_cumulative(n, l, m, x + h) - bd.origin_cumulative(n, l, m, x - h)) / (2 * h)
        assert abs(num - bd.origin_density(n, l, m, x)) < 1e-5 * max(1.0, num)


@pytest.mark.parametrize("n,l,m,seed", [(2, 1.0, 0.5, 0), (3, 1.0, 0.5, 1), (17, 1.0, 0.5, 2), (1000, 1.0, 0.5, 1), (500, 1.0, 0.5, 2),
                                       (10000, 1.0, 0.5, 3), (300, 2.0, 0.1, 99)])
def test_generator_equals_restatement(n, l, m, seed):
    if n == 2:
        # the reference rejects a single merge value (toReconstructedTree "Too few values."); the product accepts n = 2
        with pytest.raises(ValueError):
            bd.birth_death_tree(n, l, m, seed)
        assert pm.birth_death_tree(n, l, m, seed=seed).n_leaves == 2
        return
    parent, brlen, leaf_row, names = bd.birth_death_tree(n, l, m, seed)
    t = pm.birth_death_tree(n, l, m, seed=seed)
    assert np.array_equal(t.parent, parent) and np.array_equal(t.leaf_row, leaf_row)
    # same uniforms, same quantile formulas; the reference accumulates subtree heights as hl + (v - hl), which is v up to
    # one rounding, the product subtracts the merge values directly: a few branches differ in the last bit
    assert np.max(np.abs(t.brlen - brlen)) <= 4e-16 * max(1.0, float(brlen.max()))
    assert (t.brlen != brlen).mean() < 0.02
    assert t.names == names


def test_merge_heights_follow_the_reference_cdf():
    from scipy import stats

    n, l, m = 4000, 1.0, 0.5
    pvals = []
    for seed in range(6):
        t = pm.birth_death_tree(n, l, m, seed=100 + seed)
        hs, origin = bd.node_heights(t.parent, t.brlen, t.leaf_row)
        assert len(hs) == n - 1 and abs(origin - t.height()) < 1e-9
        # ultrametric: every leaf at the same depth
        depth = np.zeros(t.n_nodes)
        for i in range(t.n_nodes):
            depth[i] = t.brlen[i] + (depth[t.parent[i]] if t.parent[i] >= 0 else 0.0)
        leaves = depth[t.leaf_row >= 0]
        assert np.max(np.abs(leaves - leaves[0])) < 1e-9 * max(1.0, origin)
        cdf = np.vectorize(lambda x: bd.bd_cumulative(origin, l, m, x))
        pvals.append(stats.kstest(hs, cdf).pvalue)
    assert min(pvals) > 1e-3, pvals


def test_origins_follow_the_reference_cdf():
    from scipy import stats

    n, l, m = 50, 1.0, 0.5
    origins = np.array([pm.birth_death_tree(n, l, m, seed=s).height() for s in range(1500)])
    cdf = np.vectorize(lambda x: bd.origin_cumulative(n, l, m, x))
    assert stats.kstest(origins, cdf).pvalue > 1e-3


def test_reference_error_cases():
    from elynx_b200 import ElxError

    with pytest.raises(ElxError, match="mu > lambda"):
        pm.birth_death_tree(10, 1.0, 2.0)
    with pytest.raises(ValueError, match="mu > lambda"):
        bd.simulate_random(10, 1.0, 2.0, bd.SplitMix64(0))
