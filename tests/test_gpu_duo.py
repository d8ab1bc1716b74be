"""GPU parity tests of duo mode (elynx_b200/csrc/elx_duo.cu): amino-acid models, leaves-only calls -- one draw per cap of at
most two frontier nodes on component-bucketed sites.  Bit-exact against the executable spec (oracle/freerun_spec.c
spec_freerun_duo) on the device's own tables; the tables against brute-force sums; the output distribution against exact
Felsenstein-pruning pattern probabilities."""
import numpy as np
import pytest

import oracle.freerun_spec as spec
import oracle.model as om
import oracle.pruning as pruning
from helpers import check_duo_tables, model_case, random_tree, spec_freerun, tree_from_file

pytestmark = pytest.mark.gpu

import elynx_b200 as eb  # noqa: E402

SB = 1 << 20  # sites per superblock of the component bucketing (models with up to 16 components)


def make_sim(model, tree, **kw):
    is_mix, ws, ds, es, _ = model
    return eb.Simulator(ws if is_mix else None, ds, es, tree, is_mixture=is_mix, **kw)


@pytest.mark.parametrize("name", ["lg", "lg_g4", "c10"])
@pytest.mark.parametrize("force_generic", [False, True])
def test_duo_matches_spec_across_superblocks(name, force_generic):
    """Site ranges that start inside a superblock, end inside the next one, or lie beyond 2^32: ranks restart at every
    superblock boundary and count the sites in front of the range."""
    model = model_case(name)
    tree = tree_from_file("Newick.tree")
    with make_sim(model, tree, force_generic=force_generic) as sim:
        assert sim.duo_groups > 0
        assert sim.last_kernel == ""
        for seed, site0, s in ((1, 0, 3001), (1, SB - 700, 1500), (9, 3 * SB + 12345, 2222), (9, (1 << 34) + SB - 5, 40), (2, 17, 1)):
            leaves, comps, _ = sim.simulate(seed, s, site0=site0, want_comps=True)
            l0, c0, _ = spec_freerun(sim, model, tree, seed, site0, s)
            assert np.array_equal(comps, c0), (seed, site0)
            assert np.array_equal(leaves, l0), (seed, site0)
        assert sim.last_kernel.startswith("k2_duo_generic" if force_generic else "k2_duo_tiled")


def test_duo_partition_invariance():
    """Any partition of a site range over calls gives the same alignment (ranks are relative to the superblock start)."""
    model = model_case("lg_g4")
    tree = tree_from_file("UltraMetric.tree")
    with make_sim(model, tree) as sim:
        site0, s = SB - 5000, 12000
        whole, cw, _ = sim.simulate(5, s, site0=site0, want_comps=True)
        cuts = [0, 1, 4999, 5000, 5001, 9000, s]
        parts = [sim.simulate(5, b - a, site0=site0 + a, want_comps=True) for a, b in zip(cuts[:-1], cuts[1:])]
        assert np.array_equal(whole, np.concatenate([p[0] for p in parts], axis=1))
        assert np.array_equal(cw, np.concatenate([p[1] for p in parts]))


@pytest.mark.parametrize("shape", ["multifurcating", "caterpillar", "unary", "single", "cherry"])
def test_duo_odd_trees(shape):
    rng = np.random.default_rng(12)
    model = model_case("wag_g4")
    if shape == "multifurcating":
        tree = random_tree(37, rng, multifurcate=True, zero_frac=0.2)
    elif shape == "caterpillar":
        s_ = "t0:0.1"
        for i in range(1, 70):
            s_ = "(%s,t%d:0.05):0.02" % (s_, i)
        tree = om.parse_newick(s_ + ";")
    elif shape == "unary":
        tree = om.parse_newick("(((a:0.1):0.2,(b:0.3,c:0.1,d:0.2,e:0.4,f:0.01):0.1):0.3);")
    elif shape == "single":
        tree = om.parse_newick("a:0.3;")
    else:
        tree = om.parse_newick("(a:0.1,b:0.2):0.5;")
    for fg in (True, False):
        with make_sim(model, tree, force_generic=fg) as sim:
            assert (sim.duo_groups > 0) == (tree.n_nodes > 1)
            if sim.duo_groups:
                check_duo_tables(sim.ptables(), sim.duo_tables(), tree.parent, tree.leaf_row)
            for s, site0 in ((4099, 0), (1000, 77)):
                leaves, comps, _ = sim.simulate(5, s, site0=site0, want_comps=True)
                l1, c1, _ = spec_freerun(sim, model, tree, 5, site0, s)
                assert np.array_equal(leaves, l1) and np.array_equal(comps, c1)


def test_duo_refinement_path_is_exercised():
    """Ties (draw == threshold in the 15 high bits) come at 2^-15 per draw: run enough draws that the spec sees dozens."""
    model = model_case("lg_g4")
    tree = om.parse_newick("(" + ",".join("t%d:0.013" % i for i in range(32)) + ");")
    s = 150_000  # 16 records x 1.5e5 sites = 2.4e6 draws: 73 ties expected
    for fg in (False, True):
        with make_sim(model, tree, force_generic=fg) as sim:
            leaves, comps, _ = sim.simulate(3, s, want_comps=True)
            l0, c0, _ = spec_freerun(sim, model, tree, 3, 0, s)
            assert spec.duo_ties() >= 20
            assert np.array_equal(leaves, l0) and np.array_equal(comps, c0)


def test_duo_chi_square_vs_exact_pattern_probabilities():
    """Free-running duo output on a 4-taxon tree with both inner nodes summed out: the 160 000 site-pattern counts against
    exact pruning probabilities (LG + Gamma(4): the component draw, the bucketing and the joint rows all enter)."""
    from scipy import stats

    model = model_case("lg_g4")
    is_mix, ws, ds, es, _ = model
    tree = om.parse_newick("((a:0.3,b:0.1):0.2,(c:0.4,d:0.15):0.1);")
    n = 6_000_000
    with make_sim(model, tree) as sim:
        recs = sim.duo_tables()[0]
        assert any(int(r[14]) > sum(1 for x in r[1:5] if x >= 0) for r in recs), "no cap sums a node out"
        leaves, _, _ = sim.simulate(11, n, site0=SB - 1_000_000)
        probs = pruning.pattern_probabilities(is_mix, ws, ds, sim.ptables(), tree.parent, tree.leaf_row)
    assert len(recs) == 2
    idx = ((leaves[0].astype(np.int64) * 20 + leaves[1]) * 20 + leaves[2]) * 20 + leaves[3]
    counts = np.bincount(idx, minlength=160000).astype(np.float64)
    probs = np.asarray(probs).reshape(-1)
    assert abs(probs.sum() - 1) < 1e-12
    # pool the rare patterns (expected < 5) into one cell
    exp = probs * n
    big = exp >= 5
    o = np.concatenate([counts[big], [counts[~big].sum()]])
    e = np.concatenate([exp[big], [exp[~big].sum()]])
    pv = stats.chisquare(o, e).pvalue
    assert pv > 1e-4, pv


def _multi_superblock_check():
    model = model_case("lg_g4")
    tree = om.parse_newick("((a:0.3,b:0.1):0.2,((c:0.4,d:0.15):0.1,e:0.2):0.05,(f:0.1,g:0.2):0.3);")
    site0, s = SB - 4097, 7 * SB + 5000
    with make_sim(model, tree) as sim:
        whole, cw, _ = sim.simulate(21, s, site0=site0, want_comps=True)
        cuts = [site0] + [k * SB for k in range(1, 9)] + [site0 + s]
        for a, b in zip(cuts[:-1], cuts[1:]):
            part, cp, _ = sim.simulate(21, b - a, site0=a, want_comps=True)
            assert np.array_equal(whole[:, a - site0:b - site0], part), (a, b)
            assert np.array_equal(cw[a - site0:b - site0], cp)


@pytest.mark.parametrize("pipeline", [None, "0", "1"])
def test_duo_many_superblocks_call_equals_superblock_calls(pipeline):
    """A call spanning many superblocks must equal the concatenation of one-superblock calls.  Calls of at least 8 superblocks are
    cut into batches whose un-permute passes run on a second stream under the next batch's simulation (None: that default, the
    whole call here spans 9); ELX_DUO_PIPELINE=0 / 1 switch the pipeline off / on for every call (the switch is read once per
    process: those runs happen in a child process)."""
    if pipeline is None:
        return _multi_superblock_check()
    import os
    import subprocess
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, ELX_DUO_PIPELINE=pipeline, PYTHONPATH=os.pathsep.join([os.path.dirname(here), here, os.environ.get("PYTHONPATH", "")]))
    r = subprocess.run([sys.executable, "-c", "import test_gpu_duo as t; t._multi_superblock_check(); print('pipeline ok')"],
                       env=env, cwd=here, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "pipeline ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("k,duo", [(5, True), (7, True), (22, True), (23, False)])
def test_duo_state_counts_other_than_20(k, duo):
    """The ABI takes any number of states: duo mode covers 5..22 (K^2 <= 512 columns; the outcome decode o -> (o mod K, o div K)
    is a magic multiply that depends on K), 23 and more stay node by node."""
    rng = np.random.default_rng(k)
    c = 3
    ds = rng.dirichlet(np.ones(k) * 3, size=c)
    es = rng.uniform(0.2, 2.0, size=(c, k, k))
    es = (es + es.transpose(0, 2, 1)) / 2
    for i in range(c):
        np.fill_diagonal(es[i], 0.0)
    ws = np.array([0.5, 0.2, 0.3])
    model = (True, ws, ds, es, None)
    tree = tree_from_file("UltraMetric.tree")
    for fg in (False, True):
        with make_sim(model, tree, force_generic=fg) as sim:
            assert (sim.duo_groups > 0) == duo
            if duo and not fg:
                check_duo_tables(sim.ptables(), sim.duo_tables(), tree.parent, tree.leaf_row, max_records=4)
            leaves, comps, _ = sim.simulate(3, 5000, site0=SB - 2500, want_comps=True)
            l0, c0, _ = spec_freerun(sim, model, tree, 3, SB - 2500, 5000)
            assert np.array_equal(leaves, l0) and np.array_equal(comps, c0)
            assert leaves.max() == k - 1
