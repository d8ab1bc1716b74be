"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Bit-exact for states/indices; P(t) within 1e-12 (max-norm relative) of the hmatrix-expm restatement."""
import numpy as np
import pytest

import oracle.freerun_spec as spec
import oracle.model as om
import oracle.pruning as pruning
import oracle.sim as osim
from helpers import alias_masses, check_duo_tables, check_pair_tables, check_quad_tables, model_case, random_tree, spec_freerun, tree_from_file, uniforms

pytestmark = pytest.mark.gpu

import elynx_b200 as eb  # noqa: E402


def make_sim(model, tree, **kw):
    is_mix, ws, ds, es, _ = model
    return eb.Simulator(ws if is_mix else None, ds, es, tree, is_mixture=is_mix, **kw)


# ---------------------------------------------------------------------------------------------
# K1: P(t) tables
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,tree_file", [("jc", "Newick.tree"), ("hky_g4", "Newick.tree"), ("lg_g4", "Newick.tree"),
                                            ("wag_g4", "UltraMetric.tree"), ("c10_g4", "UltraMetric.tree"), ("edm_lg", "Newick.tree")])
def test_ptables_match_hmatrix_expm(name, tree_file):
    model = model_case(name)
    tree = tree_from_file(tree_file)
    ref = om.prob_tables(model[2], model[3], tree.brlen)
    with make_sim(model, tree) as sim:
        p = sim.ptables()
        cum = sim.ptables(cumulative=True)
    assert p.shape == ref.shape
    # north_star tolerance: 1e-12 relative -- held ELEMENT-WISE by the default K1 path (the Pade kernel, hmatrix expm's
    # own algorithm, MarkovProcess.hs:33-41)
    rel_maxnorm = np.max(np.abs(p - ref), axis=(2, 3)) / np.max(np.abs(ref), axis=(2, 3))
    assert rel_maxnorm.max() < 1e-12, rel_maxnorm.max()
    elementwise = np.max(np.abs(p - ref) / np.maximum(ref, 1e-300))
    assert elementwise < 1e-12, elementwise
    assert np.array_equal(p[tree.brlen == 0], np.broadcast_to(np.eye(p.shape[-1]), p[tree.brlen == 0].shape))
    # the cumulative tables are the sequential left-to-right sums of the rows (scanl1 (+)), bit for bit
    assert np.array_equal(cum, np.cumsum(p, axis=-1))
    assert np.all(p >= 0) and np.allclose(p.sum(-1), 1.0, atol=1e-13)


def test_ptables_short_and_long_branches():
    model = model_case("lg")
    tree = om.parse_newick("((a:1e-6,b:1e-4):1e-2,(c:0.1,d:1.0):5.0,e:20.0,f:0.0);")
    ref = om.prob_tables(model[2], model[3], tree.brlen)
    with make_sim(model, tree) as sim:
        p = sim.ptables()
    rel_maxnorm = np.max(np.abs(p - ref), axis=(2, 3)) / np.max(np.abs(ref), axis=(2, 3))
    assert rel_maxnorm.max() < 1e-12
    assert np.max(np.abs(p - ref) / np.maximum(ref, 1e-300)) < 1e-12  # element-wise, also at t = 1e-6 and t = 20


@pytest.mark.parametrize("name", ["hky_g4", "lg_g4"])
def test_ptables_eigen_form(name):
    """ELX_FLAG_EIGEN: P = I + A diag(expm1(lambda t)) B from the host eigendecomposition -- 1e-12 in the max norm,
    ~1e-10 element-wise on vanishing entries."""
    model = model_case(name)
    tree = om.parse_newick("((a:1e-6,b:1e-4):1e-2,(c:0.1,d:1.0):5.0,e:20.0,f:0.0);")
    ref = om.prob_tables(model[2], model[3], tree.brlen)
    with make_sim(model, tree, eigen=True) as sim:
        p = sim.ptables()
        cum = sim.ptables(cumulative=True)
    rel_maxnorm = np.max(np.abs(p - ref), axis=(2, 3)) / np.max(np.abs(ref), axis=(2, 3))
    assert rel_maxnorm.max() < 1e-12
    assert np.max(np.abs(p - ref) / np.maximum(ref, 1e-300)) < 1e-9
    assert np.array_equal(cum, np.cumsum(p, axis=-1))


# ---------------------------------------------------------------------------------------------
# K2 injected-uniforms mode: bit-exact against the reference algorithm
# ---------------------------------------------------------------------------------------------
CASES = [("jc", "Newick.tree", 10000), ("hky", "Newick.tree", 3001), ("hky_g4", "Newick.tree", 4099),
         ("gtr4_g4", "UltraMetric.tree", 2500), ("lg_g4", "Newick.tree", 1500), ("c10_g4", "UltraMetric.tree", 1203),
         ("edm_lg", "Newick.tree", 1000), ("mixture_dna", "Multifurcating.tree", None)]


@pytest.mark.parametrize("name,tree_file,s", CASES)
def test_injected_bit_exact(name, tree_file, s):
    rng = np.random.default_rng(abs(hash((name, tree_file))) % (2 ** 32))
    model = model_case(name)
    is_mix, ws, ds, es, _ = model
    if tree_file == "Multifurcating.tree":
        # that fixture has no branch lengths: give it some (toLengthTree would reject it otherwise)
        tree = random_tree(37, rng, multifurcate=True, zero_frac=0.2)
        s = 2000
    else:
        tree = tree_from_file(tree_file)
    u = uniforms(rng, tree.n_nodes + (2 if is_mix else 1), s)
    with make_sim(model, tree) as sim:
        p_gpu = sim.ptables()
        leaves, comps, internal = sim.simulate_injected(u, want_comps=True, keep_internal=True)
        # (1) same tables on both sides -> byte-identical alignments
        c0, l0, i0 = osim.simulate_injected(is_mix, tree.parent, tree.leaf_row, ws, ds, p_gpu, u, keep_internal=True)
        assert np.array_equal(leaves, l0)
        assert np.array_equal(comps, c0)
        assert np.array_equal(internal, i0)
        # (2) the oracle's own (Pade) tables uploaded to the device -> identical to the pure-oracle run
        p_ref = om.prob_tables(ds, es, tree.brlen)
        sim.set_ptables(p_ref)
        leaves2, comps2, internal2 = sim.simulate_injected(u, want_comps=True, keep_internal=True)
        c1, l1, i1 = osim.simulate_injected(is_mix, tree.parent, tree.leaf_row, ws, ds, p_ref, u, keep_internal=True)
        assert np.array_equal(leaves2, l1) and np.array_equal(comps2, c1) and np.array_equal(internal2, i1)


def test_injected_edge_uniforms():
    """u = 1.0, u = 2^-65, and u landing exactly on cumulative thresholds."""
    model = model_case("hky_g4")
    is_mix, ws, ds, es, _ = model
    tree = om.parse_newick("((a:0.1,b:0.0):0.05,(c:2.0,d:1e-9):0.3);")
    with make_sim(model, tree) as sim:
        p = sim.ptables()
        cum = np.cumsum(p, axis=-1)
        rows = tree.n_nodes + 2
        specials = [1.0, 2.0 ** -65, 0.5, 0.25, 0.75, np.nextafter(1.0, 0)]
        for n in range(tree.n_nodes):
            for c in range(4):
                for i in range(4):
                    specials += list(cum[n, c, i] / cum[n, c, i, -1])
        specials = np.array([x for x in specials if 0 < x <= 1.0])
        rng = np.random.default_rng(1)
        s = 4096
        u = rng.choice(specials, size=(rows, s))
        leaves, comps, internal = sim.simulate_injected(u, want_comps=True, keep_internal=True)
        c0, l0, i0 = osim.simulate_injected(is_mix, tree.parent, tree.leaf_row, ws, ds, p, u, keep_internal=True)
        assert np.array_equal(leaves, l0) and np.array_equal(comps, c0) and np.array_equal(internal, i0)


def test_injected_single_node_and_empty():
    model = model_case("jc")
    tree = om.parse_newick("x;")
    with make_sim(model, tree) as sim:
        u = np.array([[0.5799093793674204], [0.9711]])  # KAT-1's stream
        leaves, _, internal = sim.simulate_injected(u, keep_internal=True)
        assert leaves.tolist() == [[2]] and internal.tolist() == [[2]]
        leaves, _, _ = sim.simulate_injected(np.zeros((2, 0)))
        assert leaves.shape == (1, 0)
        assert sim.simulate(0, 0)[0].shape == (1, 0)


def test_injected_bad_weights_is_an_error():
    """mwc-random raises "bad weights" when no cumulative entry reaches p; u > 1 provokes it."""
    model = model_case("jc")
    tree = om.parse_newick("(a:0.1,b:0.1);")
    with make_sim(model, tree) as sim:
        u = np.full((tree.n_nodes + 1, 8), 0.5)
        u[2, 3] = 1.5
        with pytest.raises(eb.ElxError, match="bad weights"):
            sim.simulate_injected(u)
        sim.simulate_injected(np.full((tree.n_nodes + 1, 8), 0.5))  # handle still usable


# ---------------------------------------------------------------------------------------------
# K2 free-running mode
# ---------------------------------------------------------------------------------------------
def check_alias_tables(p, e16, qlo, k):
    """The alias tables encode integer masses m_i with sum 2^48 and m_i / 2^48 == p_i / sum(p) within 2^-45."""
    mass = alias_masses(e16, qlo, spec.alias_bits(k))
    assert np.all(mass.sum(-1) == 1 << 48)
    assert np.all(mass[..., k:] == 0)
    pn = p / p.sum(-1, keepdims=True)
    assert np.max(np.abs(mass[..., :k] / float(1 << 48) - pn)) < 2.0 ** -44


@pytest.mark.parametrize("name,tree_file,s", [("jc", "Newick.tree", 10000), ("hky_g4", "Newick.tree", 5000),
                                             ("lg_g4", "UltraMetric.tree", 3000), ("c10", "UltraMetric.tree", 2000),
                                             ("mixture_dna", "Newick.tree", 2000), ("c60", "UltraMetric.tree", 9000),
                                             ("c60", "Newick.tree", 1000)])
@pytest.mark.parametrize("force_generic", [True, False])
@pytest.mark.parametrize("draws", ["default", "pair", "single"])
def test_freerun_matches_spec(name, tree_file, s, force_generic, draws):
    """Every free-running kernel (generic / tiled; quad mode / pair mode / node by node) against the executable spec, on
    the device's own tables, which are themselves checked against the P rows in exact integer arithmetic."""
    model = model_case(name)
    is_mix, ws, ds, es, _ = model
    k = ds.shape[1]
    if draws == "pair" and k > 4:
        pytest.skip("pair_draws only changes models with K <= 4")
    tree = tree_from_file(tree_file)
    with make_sim(model, tree, force_generic=force_generic, single_draws=draws == "single", pair_draws=draws == "pair") as sim:
        e16, qlo = sim.alias_tables()
        check_alias_tables(sim.ptables(), e16, qlo, k)
        assert (sim.pair_groups > 0) == (k <= 4 and draws != "single")
        assert (sim.quad_groups > 0) == (k <= 4 and draws == "default")
        assert (sim.duo_groups > 0) == (5 <= k <= 22 and draws == "default")
        if sim.duo_groups:
            check_duo_tables(sim.ptables(), sim.duo_tables(), tree.parent, tree.leaf_row)
        if sim.pair_groups:
            check_pair_tables(sim.ptables(), sim.pair_tables(), tree.parent)
        if sim.quad_groups:
            check_quad_tables(sim.ptables(), sim.quad_tables(), tree.parent, tree.leaf_row)
        for seed, site0 in ((0, 0), (0xDEADBEEFCAFEF00D, 0), (7, 8 * 1000003), (7, 13)):
            # calls that keep the internal states never sum nodes out: pair mode / node by node
            leaves, comps, internal = sim.simulate(seed, s, site0=site0, want_comps=True, keep_internal=True)
            l0, c0, i0 = spec_freerun(sim, model, tree, seed, site0, s, keep_internal=True)
            assert np.array_equal(internal, i0), (seed, site0)
            assert np.array_equal(leaves, l0) and np.array_equal(comps, c0)
            # leaves only: quad mode when the handle has it (a different stream), else the same stream as above
            leaves2, comps2, _ = sim.simulate(seed, s, site0=site0, want_comps=True)
            l1, c1, _ = spec_freerun(sim, model, tree, seed, site0, s)
            assert np.array_equal(leaves2, l1) and np.array_equal(comps2, c1)
            if not sim.quad_groups and not sim.duo_groups:
                assert np.array_equal(l1, l0)


@pytest.mark.parametrize("name,kw", [("hky_g4", {"pair_draws": True}), ("hky_g4", {"single_draws": True}), ("lg_g4", {"single_draws": True})])
def test_one_stream_for_both_call_shapes(name, kw):
    """include/elynx_b200.h: a handle created with ELX_FLAG_PAIR_DRAWS / ELX_FLAG_SINGLE_DRAWS draws the same stream whether or
    not the internal states are kept (the default handle sums nodes out in leaves-only calls: a different stream)."""
    model = model_case(name)
    tree = tree_from_file("Newick.tree")
    with make_sim(model, tree, **kw) as sim:
        a, ca, internal = sim.simulate(17, 3000, site0=123, want_comps=True, keep_internal=True)
        b, cb, _ = sim.simulate(17, 3000, site0=123, want_comps=True)
        assert np.array_equal(a, b) and np.array_equal(ca, cb)
        assert np.array_equal(internal[tree.leaf_row >= 0][np.argsort(tree.leaf_row[tree.leaf_row >= 0])], a)
    with make_sim(model, tree) as sim:
        a, _, _ = sim.simulate(17, 3000, site0=123, keep_internal=True)
        b, _, _ = sim.simulate(17, 3000, site0=123)
        assert not np.array_equal(a, b)


@pytest.mark.parametrize("shape", ["multifurcating", "caterpillar", "unary", "single", "cherry"])
def test_pair_walk_odd_trees(shape):
    """Pair mode on trees that are not binary: groups of one, several groups per node, a node that is root and leaf."""
    rng = np.random.default_rng(11)
    model = model_case("hky_g4")
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
        for pair_draws in (True, False):
            with make_sim(model, tree, force_generic=fg, pair_draws=pair_draws) as sim:
                check_pair_tables(sim.ptables(), sim.pair_tables(), tree.parent)
                assert (sim.quad_groups > 0) == (not pair_draws and tree.n_nodes > 1)
                if sim.quad_groups:
                    check_quad_tables(sim.ptables(), sim.quad_tables(), tree.parent, tree.leaf_row)
                for s, site0 in ((4099, 0), (1000, 77)):
                    leaves, comps, internal = sim.simulate(5, s, site0=site0, want_comps=True, keep_internal=True)
                    l0, c0, i0 = spec_freerun(sim, model, tree, 5, site0, s, keep_internal=True)
                    assert np.array_equal(internal, i0) and np.array_equal(leaves, l0) and np.array_equal(comps, c0)
                    l1, c1, _ = spec_freerun(sim, model, tree, 5, site0, s)
                    leaves2, comps2, _ = sim.simulate(5, s, site0=site0, want_comps=True)
                    assert np.array_equal(leaves2, l1) and np.array_equal(comps2, c1)


def test_freerun_refinement_path_is_exercised():
    """Tables with many exact ties in the high bits: force r == q_hi often by using a tree of identical
    short branches and a long run; compares against the spec which always takes the same path."""
    model = model_case("hky")
    is_mix, ws, ds, es, _ = model
    tree = om.parse_newick("(" + ",".join("t%d:0.013" % i for i in range(64)) + ");")
    for kw in ({"pair_draws": True}, {"single_draws": True}):
        with make_sim(model, tree, **kw) as sim:
            s = 200000
            leaves, _, _ = sim.simulate(3, s)
            l0, _, _ = spec_freerun(sim, model, tree, 3, 0, s)
            assert np.array_equal(leaves, l0)


def test_quad_refinement_path_is_exercised():
    """Quad mode ties (draw == threshold in the 16 high bits) come at 2^-16 per draw: run enough draws that the spec sees
    dozens, on both quad kernels."""
    model = model_case("hky_g4")
    tree = om.parse_newick("(" + ",".join("t%d:0.013" % i for i in range(64)) + ");")
    s = 300_000  # 16 records x 3e5 sites = 4.8e6 draws: 73 ties expected
    with make_sim(model, tree) as sim:
        assert sim.quad_groups == 16
        l0, _, _ = spec_freerun(sim, model, tree, 3, 0, s)
        assert spec.quad_ties() > 20
        assert np.array_equal(sim.simulate(3, s)[0], l0)
    with make_sim(model, tree, force_generic=True) as sim:
        assert np.array_equal(sim.simulate(3, s)[0], l0)


@pytest.mark.parametrize("k", [2, 3])
def test_small_alphabets_quad_and_pair(k):
    """K = 2 / 3 models run in quad and pair mode too (states >= K are zero-mass outcomes of the 2-bit fields): spec
    parity on a random tree, and chi-square against exact pruning probabilities on a 4-taxon tree."""
    from scipy import stats

    rng = np.random.default_rng(40 + k)
    pi = rng.dirichlet(np.ones(k) * 3)
    ex = rng.uniform(0.3, 2.0, size=(k, k))
    ex = (ex + ex.T) / 2
    ex[np.arange(k), np.arange(k)] = 0
    ds, es = pi[None], ex[None]
    model = (False, np.ones(1), ds)
    tree = random_tree(23, rng, multifurcate=True, zero_frac=0.1)
    for kw in ({}, {"force_generic": True}, {"pair_draws": True}):
        with eb.Simulator(None, ds, es, tree, is_mixture=False, **kw) as sim:
            assert (sim.quad_groups > 0) == ("pair_draws" not in kw) and sim.pair_groups > 0
            if sim.quad_groups:
                check_quad_tables(sim.ptables(), sim.quad_tables(), tree.parent, tree.leaf_row)
            leaves = sim.simulate(9, 5000, site0=3)[0]
            l0, _, _ = spec_freerun(sim, model, tree, 9, 3, 5000)
            assert np.array_equal(leaves, l0) and leaves.max() < k
    t4 = om.parse_newick("((a:0.1,b:0.3):0.05,(c:0.02,d:0.7):0.2):0.15;")
    s = 1_000_000
    with eb.Simulator(None, ds, es, t4, is_mixture=False) as sim:
        leaves = sim.simulate(77, s)[0]
        probs = pruning.pattern_probabilities(False, np.ones(1), ds, sim.ptables(), t4.parent, t4.leaf_row)
    idx = np.zeros(s, dtype=np.int64)
    for t in range(4):
        idx = idx * k + leaves[t]
    obs = np.bincount(idx, minlength=k ** 4)
    exp = probs * s
    keep = exp >= 5
    chi2 = float(((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum())
    assert stats.chi2.sf(chi2, int(keep.sum()) - 1) > 1e-4


def test_quad_large_tree_many_components_and_far_site_offsets():
    """Quad mode away from the small fixtures: a 3000-taxon birth-death tree (~1100 records, 8+ slots), GTR + Gamma(8)
    (32 KB of rows per record: a two-stage ring), site ranges that start beyond 2^32 and 2^40 (high counter words) at
    odd offsets, tiled and generic kernels against the executable spec."""
    from elynx_b200 import model as pm

    tree = pm.birth_death_tree(3000, 1.0, 0.5, seed=11)
    model = pm.get_phylo_model("GTR4[1,2,3,4,5,6]{0.2,0.3,0.4,0.1}", gamma=(8, 0.7))
    mt = (True, model.weights, model.pi)
    for fg in (False, True):
        with eb.Simulator(model.weights, model.pi, model.exch, tree, is_mixture=True, force_generic=fg) as sim:
            assert sim.quad_groups > 1000 and sim.n_components == 8
            if not fg:
                check_quad_tables(sim.ptables(), sim.quad_tables(), tree.parent, tree.leaf_row, max_records=12)
            for site0, s in ((0, 4099), ((1 << 32) + 5, 3001), ((1 << 40) - 7, 2050)):
                leaves, comps, _ = sim.simulate(17, s, site0=site0, want_comps=True)
                l0, c0, _ = spec_freerun(sim, mt, tree, 17, site0, s)
                assert np.array_equal(leaves, l0) and np.array_equal(comps, c0), (fg, site0)


def test_quad_tables_of_the_benchmark_configuration_are_exact():
    """BASELINE config 2 (HKY + Gamma(4) on the 1000-taxon birth-death tree bench.py uses): EVERY record's alias rows carry
    the integer masses of the cap's joint distribution (brute-force sums over the summed-out nodes), and the walk covers
    the tree -- with spec parity this makes the benchmarked stream an exact sampler of the reference's leaf distribution."""
    from elynx_b200 import model as pm

    tree = pm.birth_death_tree(1000, 1.0, 0.5, seed=1)
    model = pm.get_phylo_model("HKY[6.0]{0.2,0.3,0.3,0.2}", gamma=(4, 0.5))
    with eb.Simulator(model.weights, model.pi, model.exch, tree, is_mixture=True) as sim:
        quad = sim.quad_tables()
        assert 330 <= sim.quad_groups <= 400
        check_quad_tables(sim.ptables(), quad, tree.parent, tree.leaf_row, max_records=10 ** 9)
        frontier = (quad[0][:, 1:5] >= 0).sum(axis=1)
        assert frontier.max() == 4 and frontier.sum() >= 1000  # leaves + kept internal nodes


def test_quad_and_pair_mode_agree_statistically():
    """Two exact samplers of the same leaf distribution (quad mode sums internal nodes out, pair mode draws every node):
    two-sample chi-square on the joint states of random leaf pairs and triples of a 150-taxon tree."""
    from scipy import stats

    from elynx_b200 import model as pm

    tree = pm.birth_death_tree(150, 1.0, 0.5, seed=21)
    model = pm.get_phylo_model("GTR4[1,2,3,4,5,6]{0.2,0.3,0.4,0.1}", gamma=(4, 0.5))
    s = 1_500_000
    with eb.Simulator(model.weights, model.pi, model.exch, tree, is_mixture=True) as sim:
        assert sim.quad_groups > 0
        q = sim.simulate(5, s)[0]
    with eb.Simulator(model.weights, model.pi, model.exch, tree, is_mixture=True, pair_draws=True) as sim:
        assert sim.quad_groups == 0
        p = sim.simulate(6, s)[0]
    rng = np.random.default_rng(0)
    pvals = []
    for _ in range(40):
        rows = rng.choice(150, size=3, replace=False)
        hq = np.bincount((q[rows[0]].astype(np.int64) * 4 + q[rows[1]]) * 4 + q[rows[2]], minlength=64)
        hp = np.bincount((p[rows[0]].astype(np.int64) * 4 + p[rows[1]]) * 4 + p[rows[2]], minlength=64)
        keep = (hq + hp) >= 10
        pvals.append(stats.chi2_contingency(np.stack([hq[keep], hp[keep]]))[1])
    pvals = np.array(pvals)
    assert pvals.min() > 1e-5 and (pvals < 0.01).sum() <= 3, pvals


def test_freerun_partition_invariance():
    """Output is a pure function of (seed, global site, node): any split of the site range gives the same bytes."""
    model = model_case("hky_g4")
    tree = tree_from_file("Newick.tree")
    with make_sim(model, tree) as sim:
        whole, cw, _ = sim.simulate(11, 5000, want_comps=True)
        cuts = [0, 1, 17, 1024, 1031, 4096, 5000]
        parts = [sim.simulate(11, b - a, site0=a, want_comps=True) for a, b in zip(cuts[:-1], cuts[1:])]
        assert np.array_equal(whole, np.concatenate([p[0] for p in parts], axis=1))
        assert np.array_equal(cw, np.concatenate([p[1] for p in parts]))
    with make_sim(model, tree, force_generic=True) as sim2:
        assert np.array_equal(whole, sim2.simulate(11, 5000)[0])


@pytest.mark.parametrize("keep_internal", [False, True])
def test_host_buffer_call_packed_transport_is_byte_identical(keep_internal, monkeypatch):
    """elx_simulate with host buffers: alignments over <= 4 states cross PCIe packed to 2 bits per site (k_pack2) and are
    expanded by host threads; the bytes must equal the plain 1-byte transport for multi-batch, ragged site ranges, with the
    component vector and the internal states, and the packed call must move a quarter of the bytes."""
    model = model_case("hky_g4")
    tree = tree_from_file("Newick.tree")
    nsites = 2 * (1 << 21) + 12345  # three device batches of 2^21 sites, the last one ragged (not a multiple of 4)
    with make_sim(model, tree) as sim:
        monkeypatch.setenv("ELX_PACKED_D2H", "0")
        b0 = sim.d2h_bytes
        plain = sim.simulate(5, nsites, site0=77, want_comps=True, keep_internal=keep_internal)
        b1 = sim.d2h_bytes
        rows = tree.n_leaves + (tree.n_nodes if keep_internal else 0)
        assert b1 - b0 == nsites * (rows + 4)
        for threads in ("1", "5"):
            monkeypatch.setenv("ELX_PACKED_D2H", "1")
            monkeypatch.setenv("ELX_HOST_THREADS", threads)
            b1 = sim.d2h_bytes
            packed = sim.simulate(5, nsites, site0=77, want_comps=True, keep_internal=keep_internal)
            moved = sim.d2h_bytes - b1
            assert np.array_equal(plain[0], packed[0]) and np.array_equal(plain[1], packed[1])
            if keep_internal:
                assert np.array_equal(plain[2], packed[2])
            assert moved < 0.26 * nsites * rows + 4 * nsites + (1 << 16)
        monkeypatch.delenv("ELX_HOST_THREADS")
        monkeypatch.delenv("ELX_PACKED_D2H")
        if keep_internal:  # (calls that keep the internal states draw sibling pairs: another stream than leaves-only calls)
            return
        small = sim.simulate(5, 1000, site0=77)[0]   # below the size threshold: plain transport
        assert np.array_equal(small, plain[0][:, :1000])
        auto = sim.simulate(5, nsites, site0=77)[0]   # above it: packed by default
        assert np.array_equal(auto, plain[0])
        # a caller's pitch and an unaligned destination
        import ctypes as C
        from elynx_b200 import _lib as L
        pitch = nsites + 13
        raw = np.full(tree.n_leaves * pitch + 64, 0xEE, dtype=np.uint8)
        L.check(L.lib().elx_simulate(sim.handle, C.c_uint64(5), 77, nsites, C.c_void_p(raw.ctypes.data + 3), pitch, None, None, 0), sim.handle)
        view = raw[3:3 + tree.n_leaves * pitch].reshape(tree.n_leaves, pitch)
        assert np.array_equal(view[:, :nsites], plain[0])
        assert (view[:-1, nsites:] == 0xEE).all() and (raw[:3] == 0xEE).all()
        # FASTA through host buffers: 2 bits per character over PCIe, `>name` lines and characters written on the host
        names = tree.leaf_names()
        alphabet = om.ALPHABET_CHARS[model[4]]
        for packed_env in ("0", "1"):
            monkeypatch.setenv("ELX_PACKED_D2H", packed_env)
            fasta, fc = sim.simulate_fasta(5, nsites, names, alphabet, want_comps=True)
            assert fasta == om.sequences_to_fasta(names, sim.simulate(5, nsites)[0], model[4]), packed_env
            assert np.array_equal(fc, sim.simulate(5, nsites, want_comps=True)[1])
        monkeypatch.delenv("ELX_PACKED_D2H")
        # a page-locked destination
        import torch
        host = torch.full((tree.n_leaves, nsites), 0xEE, dtype=torch.uint8).pin_memory()
        L.check(L.lib().elx_simulate(sim.handle, C.c_uint64(5), 77, nsites, C.c_void_p(host.data_ptr()), nsites, None, None, 0), sim.handle)
        assert np.array_equal(host.numpy(), plain[0])


@pytest.mark.parametrize("name", ["hky_g4", "lg_g4"])
def test_packed_transport_forced_on_small_and_ragged_calls(name, monkeypatch):
    """ELX_PACKED_D2H=1 takes the packed path whatever the size: 1 site, sizes around the 32-site packing group and the
    65 536-site work unit, with components and internal states; FASTA too."""
    model = model_case(name)
    tree = tree_from_file("Newick.tree")
    names = tree.leaf_names()
    alphabet = om.ALPHABET_CHARS[model[4]]
    with make_sim(model, tree) as sim:
        for nsites in (1, 31, 32, 33, 4097, 65536, 65537, 131073):
            monkeypatch.setenv("ELX_PACKED_D2H", "0")
            plain = sim.simulate(3, nsites, site0=9, want_comps=True, keep_internal=True)
            fasta0 = sim.simulate_fasta(3, nsites, names, alphabet)[0]
            monkeypatch.setenv("ELX_PACKED_D2H", "1")
            packed = sim.simulate(3, nsites, site0=9, want_comps=True, keep_internal=True)
            for a, b in zip(plain, packed):
                assert np.array_equal(a, b), nsites
            assert sim.simulate_fasta(3, nsites, names, alphabet)[0] == fasta0, nsites


@pytest.mark.parametrize("keep_internal", [False, True])
def test_host_buffer_call_5bit_transport_protein(keep_internal, monkeypatch):
    """Amino-acid alignments cross PCIe at 5 bits per site (k_pack5 / unpack5_rows): byte-identical to the 1-byte transport,
    for elx_simulate (with components and internal states) and for the FASTA image."""
    model = model_case("lg_g4")
    tree = tree_from_file("Newick.tree")
    nsites = (1 << 20) + 777
    names = tree.leaf_names()
    alphabet = om.ALPHABET_CHARS[model[4]]
    with make_sim(model, tree) as sim:
        monkeypatch.setenv("ELX_PACKED_D2H", "0")
        plain = sim.simulate(9, nsites, site0=5, want_comps=True, keep_internal=keep_internal)
        fasta0 = sim.simulate_fasta(9, nsites, names, alphabet)[0] if not keep_internal else None
        rows = tree.n_leaves + (tree.n_nodes if keep_internal else 0)
        for threads in ("1", "6"):
            monkeypatch.setenv("ELX_PACKED_D2H", "1")
            monkeypatch.setenv("ELX_HOST_THREADS", threads)
            b1 = sim.d2h_bytes
            packed = sim.simulate(9, nsites, site0=5, want_comps=True, keep_internal=keep_internal)
            moved = sim.d2h_bytes - b1
            assert np.array_equal(plain[0], packed[0]) and np.array_equal(plain[1], packed[1])
            if keep_internal:
                assert np.array_equal(plain[2], packed[2])
            assert moved < 0.63 * nsites * rows + 4 * nsites + (1 << 16)
        if not keep_internal:
            assert plain[0].max() == 19
            assert sim.simulate_fasta(9, nsites, names, alphabet)[0] == fasta0
            assert fasta0 == om.sequences_to_fasta(names, sim.simulate(9, nsites)[0], model[4])


@pytest.mark.parametrize("name", ["jc", "hky_g4", "mixture_dna"])
def test_freerun_chi_square_vs_exact_pattern_probabilities(name):
    """Free-running output vs exact Felsenstein-pruning pattern probabilities on a 4-taxon tree (256 patterns)."""
    from scipy import stats

    model = model_case(name)
    is_mix, ws, ds, es, _ = model
    tree = om.parse_newick("((a:0.1,b:0.3):0.05,(c:0.02,d:0.7):0.2):0.15;")
    s = 2_000_000
    with make_sim(model, tree) as sim:
        leaves, _, _ = sim.simulate(2024, s)
        probs = pruning.pattern_probabilities(is_mix, ws, ds, sim.ptables(), tree.parent, tree.leaf_row)
    assert abs(probs.sum() - 1) < 1e-12
    idx = np.zeros(s, dtype=np.int64)
    for t in range(4):
        idx = idx * 4 + leaves[t]
    obs = np.bincount(idx, minlength=256)
    exp = probs * s
    keep = exp >= 5
    chi2 = float(((obs[keep] - exp[keep]) ** 2 / exp[keep]).sum())
    pval = stats.chi2.sf(chi2, int(keep.sum()) - 1)
    assert pval > 1e-4, (chi2, pval)


def test_freerun_chi_square_protein_marginals_and_pairs():
    """LG+G4: per-taxon state frequencies vs pi and the pair pattern table of two leaves vs exact probabilities."""
    from scipy import stats

    model = model_case("lg_g4")
    is_mix, ws, ds, es, _ = model
    tree = om.parse_newick("(a:0.2,b:0.4);")
    s = 1_000_000
    with make_sim(model, tree) as sim:
        leaves, comps, _ = sim.simulate(5, s, want_comps=True)
        probs = pruning.pattern_probabilities(is_mix, ws, ds, sim.ptables(), tree.parent, tree.leaf_row)
    obs = np.bincount(leaves[0].astype(np.int64) * 20 + leaves[1], minlength=400)
    exp = probs * s
    chi2 = float(((obs - exp) ** 2 / exp).sum())
    assert stats.chi2.sf(chi2, 399) > 1e-4, chi2
    cc = np.bincount(comps, minlength=4)
    assert stats.chisquare(cc).pvalue > 1e-4  # gamma categories are equiprobable (weights all 1.0)


def test_freerun_two_sample_vs_oracle_stream():
    """Free-running GPU output vs the oracle's free-running output on the reference's SplitMix stream
    (two-sample chi-square on site patterns of a 3-taxon tree)."""
    from scipy import stats

    model = model_case("gtr4_g4")
    is_mix, ws, ds, es, _ = model
    tree = om.parse_newick("((a:0.3,b:0.1):0.2,c:0.5);")
    s = 500_000
    with make_sim(model, tree) as sim:
        p = sim.ptables()
        g = sim.simulate(99, s)[0]
    o = osim.simulate_splitmix(is_mix, tree.parent, tree.leaf_row, ws, ds, p, s, 99, nchunks=4, nthreads=4)[1]

    def hist(l):
        return np.bincount((l[0].astype(np.int64) * 4 + l[1]) * 4 + l[2], minlength=64)

    table = np.stack([hist(g), hist(o)])
    assert stats.chi2_contingency(table)[1] > 1e-4


# ---------------------------------------------------------------------------------------------
# public API mirror, K3 FASTA, histograms
# ---------------------------------------------------------------------------------------------
def test_reference_api_surface():
    tree = tree_from_file("UltraMetric.tree")
    sm = om.jc(om.DO_NORMALIZE)
    leaves = eb.simulateAndFlattenPar(100, sm.stationary_distribution, sm.exchangeability_matrix, tree, 1)
    assert leaves.shape == (10, 100) and leaves.max() < 4
    internal = eb.simulate(10, sm.stationary_distribution, sm.exchangeability_matrix, tree, 1)
    assert internal.shape == (tree.n_nodes, 10)  # MarkovProcessAlongTreeSpec.hs:49-52 (length 10 at every node)
    _, ws, ds, es, _ = model_case("c10")
    comps, leaves = eb.simulateAndFlattenMixtureModelPar(64, ws, ds, es, tree, 2)
    assert comps.shape == (64,) and comps.max() < 10 and leaves.shape == (10, 64)
    assert eb.simulateMixtureModel(5, ws, ds, es, tree, 2).shape == (tree.n_nodes, 5)


@pytest.mark.parametrize("name,s", [("hky_g4", 1000), ("lg", 777), ("jc", 0), ("jc", 1)])
def test_fasta_matches_reference_layout(name, s):
    model = model_case(name)
    tree = tree_from_file("Newick.tree")
    names = tree.leaf_names()
    alphabet = om.ALPHABET_CHARS[model[4]]
    with make_sim(model, tree) as sim:
        fasta, _ = sim.simulate_fasta(42, s, names, alphabet)
        leaves, _, _ = sim.simulate(42, s)
    assert fasta == om.sequences_to_fasta(names, leaves, model[4])


def test_histograms_device():
    torch = pytest.importorskip("torch")
    model = model_case("hky")
    tree = om.parse_newick("((a:0.1,b:0.3):0.05,(c:0.02,d:0.7):0.2);")
    s = 100003
    with make_sim(model, tree) as sim:
        leaves, _, _ = sim.simulate(1, s)
        d = torch.from_numpy(leaves).cuda()
        sc = torch.zeros((4, 4), dtype=torch.int64, device="cuda")
        pc = torch.zeros(256, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        sim.histogram_device(d.data_ptr(), s, s, sc.data_ptr(), pc.data_ptr())
        torch.cuda.current_stream().synchronize()
        import ctypes
        torch.cuda.synchronize()
        sc, pc = sc.cpu().numpy(), pc.cpu().numpy()
    for t in range(4):
        assert np.array_equal(sc[t], np.bincount(leaves[t], minlength=4))
    idx = ((leaves[0].astype(np.int64) * 4 + leaves[1]) * 4 + leaves[2]) * 4 + leaves[3]
    assert np.array_equal(pc, np.bincount(idx, minlength=256))


# ---------------------------------------------------------------------------------------------
# BASELINE-size runs checked through size-independent properties
# ---------------------------------------------------------------------------------------------
def _device_run(sim, seed, site0, n, t, torch):
    pitch = (n + 127) // 128 * 128
    buf = torch.empty((t, pitch), dtype=torch.uint8, device="cuda")
    sim.simulate_device(seed, site0, n, buf.data_ptr(), pitch)
    return buf, pitch


@pytest.mark.parametrize("spec_str,gamma,taxa,sites", [("HKY[6.0]{0.2,0.3,0.3,0.2}", (4, 0.5), 1000, 10_000_000),
                                                       ("LG", (4, 0.5), 1000, 2_000_000)])
def test_full_size_stationarity_and_partition_checksums(spec_str, gamma, taxa, sites):
    """BASELINE configs 2/3 at (near) full size on the device: (1) the process is stationary, so every taxon's state
    frequencies must match pi (chi-square per taxon, device histograms); (2) checksum of checksums: the state-count
    histogram of the whole run equals the sum over an uneven 3-way split of the site range (pure function of
    (seed, site, node))."""
    torch = pytest.importorskip("torch")
    from scipy import stats
    from elynx_b200 import model as pm

    tree = pm.birth_death_tree(taxa, 1.0, 0.5, seed=1)
    model = pm.get_phylo_model(spec_str, gamma=gamma)
    k = model.n_states
    with eb.Simulator(model.weights, model.pi, model.exch, tree, is_mixture=True) as sim:
        stream = torch.cuda.ExternalStream(sim.stream)
        buf, pitch = _device_run(sim, 5, 0, sites, taxa, torch)
        whole = torch.zeros((taxa, k), dtype=torch.int64, device="cuda")
        sim.histogram_device(buf.data_ptr(), pitch, sites, whole.data_ptr())
        stream.synchronize()
        whole = whole.cpu().numpy()
        del buf
        parts = np.zeros_like(whole)
        cuts = [0, 16 * 1001, sites // 3 + 5, sites]
        for a, b in zip(cuts[:-1], cuts[1:]):
            pbuf, ppitch = _device_run(sim, 5, a, b - a, taxa, torch)
            h = torch.zeros((taxa, k), dtype=torch.int64, device="cuda")
            sim.histogram_device(pbuf.data_ptr(), ppitch, b - a, h.data_ptr())
            stream.synchronize()
            parts += h.cpu().numpy()
            del pbuf
    assert np.array_equal(whole, parts)
    assert np.all(whole.sum(1) == sites)
    pi = model.pi[0]  # gamma categories share pi
    pvals = np.array([stats.chisquare(whole[t], pi * sites).pvalue for t in range(taxa)])
    # 1000 independent-ish tests: no catastrophic p-value, and the p-values are not piled up near zero
    assert pvals.min() > 1e-7, pvals.min()
    assert (pvals < 0.01).mean() < 0.05


def test_cli_slynx_test_matrix(tmp_path, ref_data):
    """scripts/slynx-test:25-50: the model spellings the reference's smoke script runs must parse and run."""
    import os

    from elynx_b200 import cli

    tree = os.path.join(ref_data, "Newick.tree")
    edm = os.path.join(ref_data, "EDMDistsPhylobayes.txt")
    prof = os.path.join(ref_data, "HSSPMany.siteprofiles")
    cases = [["-s", "HKY[6.0]{0.2,0.3,0.3,0.2}"], ["-s", "GTR4[1,2,3,4,5,6]{0.2,0.3,0.4,0.1}"], ["-e", edm, "-m", "EDM(LG-Custom)"],
             ["-e", edm, "-m", "EDM(LG-Custom)", "-w", "[0.9,0.05,0.03,0.02]"], ["-m", "C20", "-n"],
             ["-g", "(4,0.2)", "-s", "HKY[6.0]{0.2,0.3,0.3,0.2}"], ["-m", "C10", "-n", "-g", "(4,0.2)"],
             ["-m", "MIXTURE(JC,HKY[6.0]{0.25,0.25,0.25,0.25})", "-w", "[0.4,0.6]"], ["-m", "EDM(Poisson-Custom)", "-p", prof]]
    for i, c in enumerate(cases):
        base = str(tmp_path / ("out%d" % i))
        assert cli.main(["-v", "Quiet", "-o", base, "simulate", "-t", tree] + c + ["-l", "1000", "-S", "[0]"]) == 0
        lines = open(base + ".fasta", "rb").read().split(b"\n")
        assert len(lines) == 81 and lines[-1] == b"" and lines[0] == b">Aeropyrum0"
        assert all(len(l) == 1000 for l in lines[1::2][:40])
        alphabet = b"ACGT" if ("HKY" in " ".join(c) or "GTR4" in " ".join(c) or "JC" in " ".join(c)) else b"ACDEFGHIKLMNPQRSTVWY"
        assert set(b"".join(lines[1::2])) <= set(alphabet)
    sd = open(str(tmp_path / "out2") + ".sitedists").read().split("\n")
    assert len(sd) == 1000 and sd[0].startswith("1 ") and len(sd[0].split()) == 21


def test_fasta_straight_to_pinned_file_image(tmp_path):
    """K3 packs into a cudaHostRegister'ed mmap of the output file; emulating 3 ranks one after the other on one GPU
    must give the bytes of the single-call FASTA path (site ranges are disjoint, counters use global site indices)."""
    pytest.importorskip("torch")
    from elynx_b200.distributed import simulate_fasta_to_file

    model = model_case("lg_g4")
    tree = tree_from_file("Newick.tree")
    names = tree.leaf_names()
    s = 10007
    with make_sim(model, tree) as sim:
        ref, _ = sim.simulate_fasta(9, s, names, om.ALPHABET_CHARS[model[4]])
        path = str(tmp_path / "aln.fasta")
        for rank in range(3):
            simulate_fasta_to_file(sim, 9, s, names, om.ALPHABET_CHARS[model[4]], path, rank=rank, world=3, batch_sites=2048)
        assert open(path, "rb").read() == ref
        one = str(tmp_path / "one.fasta")
        simulate_fasta_to_file(sim, 9, s, names, om.ALPHABET_CHARS[model[4]], one)
        assert open(one, "rb").read() == ref


# ---------------------------------------------------------------------------------------------
# general (Pade) K1 path: non-reversible models and zero stationary frequencies
# ---------------------------------------------------------------------------------------------
def test_pade_path_matches_hmatrix_expm_closely():
    """ELX_FLAG_PADE runs hmatrix expm's own algorithm on the device: agreement with the restatement ~1e-14."""
    model = model_case("lg_g4")
    tree = om.parse_newick("((a:1e-6,b:1e-4):1e-2,(c:0.1,d:1.0):5.0,e:20.0,f:0.0);")
    ref = om.prob_tables(model[2], model[3], tree.brlen)
    with make_sim(model, tree, pade=True) as sim:
        p = sim.ptables()
        cum = sim.ptables(cumulative=True)
    assert np.max(np.abs(p - ref) / np.maximum(ref, 1e-300)) < 1e-12  # element-wise, also on the 1e-6 branch
    assert np.max(np.abs(p - ref)) < 5e-15
    assert np.array_equal(cum, np.cumsum(p, axis=-1))


@pytest.mark.parametrize("kind", ["asymmetric", "zero_pi", "zero_pi_protein"])
def test_non_reversible_models(kind):
    rng = np.random.default_rng(3)
    if kind == "asymmetric":
        ds = np.array([[0.1, 0.2, 0.3, 0.4]])
        es = rng.uniform(0.2, 2.0, size=(1, 4, 4))
        es[0][np.arange(4), np.arange(4)] = 0
    elif kind == "zero_pi":
        ds = np.array([[0.5, 0.5, 0.0, 0.0]])
        es = (np.ones((4, 4)) - np.eye(4))[None]
    else:
        pi = om.lg_stat_dist().copy()
        pi[[3, 7]] = 0.0
        ds = (pi / pi.sum())[None]
        es = om.lg_exch()[None]
    tree = tree_from_file("UltraMetric.tree")
    ref = om.prob_tables(ds, es, tree.brlen)
    sim = eb.Simulator(None, ds, es, tree, is_mixture=False)
    with sim:
        p = sim.ptables()
        assert np.max(np.abs(p - ref)) < 1e-13
        u = uniforms(rng, tree.n_nodes + 1, 3000)
        leaves, _, internal = sim.simulate_injected(u, keep_internal=True)
        _, l0, i0 = osim.simulate_injected(False, tree.parent, tree.leaf_row, np.ones(1), ds, p, u, keep_internal=True)
        assert np.array_equal(leaves, l0) and np.array_equal(internal, i0)
        fr = sim.simulate(4, 3000)[0]
        l1, _, _ = spec_freerun(sim, (False, np.ones(1), ds), tree, 4, 0, 3000)
        assert np.array_equal(fr, l1)
        if kind != "asymmetric":
            dead = np.nonzero(ds[0] == 0)[0]
            assert not np.isin(fr, dead).any()  # states with zero frequency are never entered


@pytest.mark.parametrize("name,tree_file", [("hky_g4", "Newick.tree"), ("lg_g4", "UltraMetric.tree"), ("c60", "UltraMetric.tree")])
def test_philox7_option_matches_spec(name, tree_file):
    """philox_rounds = 7 (the minimum Crush-resistant round count; +15 % throughput) is a library option; both the generic
    and the tiled kernels must follow the same spec with 7 rounds."""
    model = model_case(name)
    is_mix, ws, ds, es, _ = model
    tree = tree_from_file(tree_file)
    s = 6000
    outs = []
    for fg in (True, False):
        with make_sim(model, tree, force_generic=fg, philox_rounds=7) as sim:
            leaves, comps, _ = sim.simulate(21, s, site0=32, want_comps=True)
            outs.append(leaves)
            l0, c0, _ = spec_freerun(sim, model, tree, 21, 32, s, rounds=7)
            l10, _, _ = spec_freerun(sim, model, tree, 21, 32, s, rounds=10)
    assert np.array_equal(outs[0], l0) and np.array_equal(outs[1], l0) and np.array_equal(comps, c0)
    assert not np.array_equal(l10, l0)
