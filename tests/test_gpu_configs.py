"""GPU parity tests on the BASELINE.json shapes round 1 left untested on the device ("Next round" item 1):

  cfg5  10 000-taxon birth-death tree (seed 3), WAG + Gamma(4)
  cfg4  500-taxon birth-death tree (seed 2): (i) -m C60 -n, (ii) -m "EDM(LG-Custom)" -n with the 60 C60 profiles in a
        Phylobayes file (EDMModelPhylobayes.hs:33-64), and C60 + Gamma(4) (C = 240)
  caps  quad mode on >= 3 caps against exact Felsenstein-pruning pattern probabilities (8 taxa, 4^8 patterns); the protein
        sampler on 3 taxa (20^3 patterns)

Every kernel family the handle can pick (tiled / generic) is held to the executable spec of the product's draws on the
handle's own tables (bit-exact), and the output distribution to ground truth the sampler never saw (chi-square)."""
import numpy as np
import pytest

import oracle.model as om
import oracle.pruning as pruning
from helpers import spec_freerun

pytestmark = pytest.mark.gpu

import elynx_b200 as eb  # noqa: E402
from elynx_b200 import model as pm  # noqa: E402


def _sim(model, tree, **kw):
    return eb.Simulator(model.weights if model.is_mixture else None, model.pi, model.exch, tree, is_mixture=model.is_mixture, **kw)


def _mt(model):
    return (model.is_mixture, model.weights, model.pi)


def _device_hist(sim, seed, site0, n, t, k, want_comps=False):
    """state counts [T][K] (and the component vector) of a device-resident run"""
    torch = pytest.importorskip("torch")
    pitch = (n + 127) // 128 * 128
    buf = torch.empty((t, pitch), dtype=torch.uint8, device="cuda")
    comps = torch.empty(n, dtype=torch.int32, device="cuda") if want_comps else None
    sc = torch.zeros((t, k), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    sim.simulate_device(seed, site0, n, buf.data_ptr(), pitch, d_comps=comps.data_ptr() if want_comps else 0)
    sim.histogram_device(buf.data_ptr(), pitch, n, sc.data_ptr())
    sim.synchronize()
    return sc.cpu().numpy(), (comps.cpu().numpy() if want_comps else None)


def _stationarity(counts, pi_mix, n):
    """per-taxon chi-square of the state counts against the stationary mixture frequencies"""
    from scipy import stats

    exp = pi_mix * n
    keep = exp >= 5
    chi2 = ((counts[:, keep] - exp[keep]) ** 2 / exp[keep]).sum(axis=1)
    return stats.chi2.sf(chi2, int(keep.sum()) - 1)


# ---------------------------------------------------------------------------------------------
# cfg5: 10 000 taxa, WAG + Gamma(4)
# ---------------------------------------------------------------------------------------------
def test_cfg5_shape_tiled_generic_and_spec_agree():
    tree = pm.birth_death_tree(10000, 1.0, 0.5, seed=3)
    model = pm.get_phylo_model("WAG", gamma=(4, 0.5))
    assert tree.n_leaves == 10000 and tree.n_nodes == 19999 and model.n_components == 4
    s = 2051
    outs = {}
    for fg in (False, True):
        with _sim(model, tree, force_generic=fg) as sim:
            for site0 in (0, (1 << 32) + 5):
                leaves, comps, _ = sim.simulate(31, s, site0=site0, want_comps=True)
                outs[(fg, site0)] = leaves
                l0, c0, _ = spec_freerun(sim, _mt(model), tree, 31, site0, s)
                assert np.array_equal(leaves, l0) and np.array_equal(comps, c0), (fg, site0)
            if not fg:
                # the P(t) tables of this tree, element-wise against the hmatrix-expm restatement (sampled branches)
                p = sim.ptables()
                idx = np.linspace(0, tree.n_nodes - 1, 64).astype(int)
                ref = om.prob_tables(model.pi, model.exch, tree.brlen[idx])
                assert np.max(np.abs(p[idx] - ref)) < 5e-14
                assert np.max(np.abs(p[idx] - ref) / np.maximum(ref, 1e-300)) < 1e-12
    for site0 in (0, (1 << 32) + 5):
        assert np.array_equal(outs[(False, site0)], outs[(True, site0)])


@pytest.mark.parametrize("taxa,seed,mstr,gamma", [(1000, 1, "HKY[6.0]{0.2,0.3,0.3,0.2}", (4, 0.5)), (1000, 1, "LG", (4, 0.5)),
                                                  (500, 2, "C60", None)])
def test_ptables_elementwise_on_baseline_trees(taxa, seed, mstr, gamma):
    """probMatrix parity (MarkovProcess.hs:33-41) ELEMENT-WISE <= 1e-12 on the trees of BASELINE configs 2, 3 and 4 (sampled
    branches, every component) -- the default K1 path."""
    tree = pm.birth_death_tree(taxa, 1.0, 0.5, seed=seed)
    model = pm.get_phylo_model(mstr, gamma=gamma) if mstr != "C60" else pm.get_phylo_model(mixture_model="C60", global_normalization=True)
    with _sim(model, tree) as sim:
        p = sim.ptables()
    idx = np.unique(np.concatenate([np.linspace(0, tree.n_nodes - 1, 48).astype(int), np.argsort(tree.brlen)[:9], np.argsort(tree.brlen)[-8:]]))
    ref = om.prob_tables(model.pi, model.exch, tree.brlen[idx])
    assert np.max(np.abs(p[idx] - ref) / np.maximum(ref, 1e-300)) < 1e-12


def test_cfg5_shape_stationarity_from_device_histograms():
    """The process is stationary: every one of the 10 000 taxa must show WAG's frequencies (device histograms of a
    device-resident run; chi-square per taxon)."""
    tree = pm.birth_death_tree(10000, 1.0, 0.5, seed=3)
    model = pm.get_phylo_model("WAG", gamma=(4, 0.5))
    n = 100_000
    with _sim(model, tree) as sim:
        counts, comps = _device_hist(sim, 77, (1 << 33) + 16, n, tree.n_leaves, 20, want_comps=True)
    assert np.all(counts.sum(1) == n)
    pv = _stationarity(counts, model.pi[0], n)
    assert pv.min() > 1e-8, pv.min()
    assert (pv < 0.01).mean() < 0.03
    from scipy import stats
    assert stats.chisquare(np.bincount(comps, minlength=4)).pvalue > 1e-5


# ---------------------------------------------------------------------------------------------
# cfg4: 500 taxa, C60 / EDM(LG-Custom) with the C60 profiles / C60 + Gamma(4)
# ---------------------------------------------------------------------------------------------
def _c60_as_phylobayes():
    """The 60 C60 profiles (weights + stationary distributions) written in the Phylobayes EDM format."""
    c60 = om.cxx(60)
    lines = ["20 A C D E F G H I K L M N P Q R S T V W Y", "60"]
    for w, sm in zip(c60.weights, c60.models):
        lines.append(" ".join(repr(float(x)) for x in [w] + list(sm.stationary_distribution)))
    return "\n".join(lines) + "\n"


def _cfg4_models():
    edm = pm.parse_edm_phylobayes(_c60_as_phylobayes())
    assert len(edm) == 60
    return {
        "c60": pm.get_phylo_model(mixture_model="C60", global_normalization=True),
        "edm_lg": pm.get_phylo_model(mixture_model="EDM(LG-Custom)", global_normalization=True, edm_components=edm),
    }


@pytest.mark.parametrize("name", ["c60", "edm_lg"])
def test_cfg4_shape_kernels_match_spec_and_mixture_statistics(name):
    from scipy import stats

    tree = pm.birth_death_tree(500, 1.0, 0.5, seed=2)
    model = _cfg4_models()[name]
    assert model.n_components == 60 and model.n_states == 20 and tree.n_leaves == 500
    # host model against the oracle's construction of the same model
    edm_o = om.parse_edm_phylobayes(_c60_as_phylobayes())
    ref = om.get_phylo_model(None, "C60", global_norm=True) if name == "c60" else om.get_phylo_model(None, "EDM(LG-Custom)", global_norm=True, edm_cs=edm_o)
    _, ws, ds, es = om.flatten_model(ref)
    assert np.allclose(model.weights, ws, rtol=1e-13) and np.allclose(model.pi, ds, rtol=1e-12) and np.allclose(model.exch, es, rtol=1e-12)
    s = 3001
    outs = []
    for fg in (False, True):
        with _sim(model, tree, force_generic=fg) as sim:
            leaves, comps, _ = sim.simulate(5, s, site0=(1 << 32) - 7, want_comps=True)
            l0, c0, _ = spec_freerun(sim, _mt(model), tree, 5, (1 << 32) - 7, s)
            assert np.array_equal(leaves, l0) and np.array_equal(comps, c0), fg
            outs.append(leaves)
    assert np.array_equal(outs[0], outs[1])
    n = 400_000
    with _sim(model, tree) as sim:
        counts, comps = _device_hist(sim, 9, 48, n, tree.n_leaves, 20, want_comps=True)
    # component histogram against the weights, stationarity against sum_c w_c pi_c
    w = model.weights / model.weights.sum()
    assert stats.chisquare(np.bincount(comps, minlength=60), w * n).pvalue > 1e-5
    pv = _stationarity(counts, w @ model.pi, n)
    assert pv.min() > 1e-7 and (pv < 0.01).mean() < 0.05, (pv.min(), (pv < 0.01).mean())


def test_c60_gamma4_generic_fallback():
    """C60 + Gamma(4): 240 components, 307 KB of rows per branch -- no shared-memory layout holds that, the handle must fall
    back to tables through L1/L2 and still follow the spec; category-major component order (expand, MixtureModel)."""
    from scipy import stats

    tree = pm.birth_death_tree(60, 1.0, 0.5, seed=4)
    model = pm.get_phylo_model(mixture_model="C60", global_normalization=True, gamma=(4, 0.5))
    assert model.n_components == 240
    s = 4099
    with _sim(model, tree) as sim:
        leaves, comps, _ = sim.simulate(3, s, site0=11, want_comps=True)
        l0, c0, _ = spec_freerun(sim, _mt(model), tree, 3, 11, s)
        assert np.array_equal(leaves, l0) and np.array_equal(comps, c0)
        n = 300_000
        counts, comps = _device_hist(sim, 4, 0, n, tree.n_leaves, 20, want_comps=True)
    w = model.weights / model.weights.sum()
    assert stats.chisquare(np.bincount(comps, minlength=240), w * n).pvalue > 1e-5
    pv = _stationarity(counts, w @ model.pi, n)
    assert pv.min() > 1e-6


# ---------------------------------------------------------------------------------------------
# cap composition against independent ground truth
# ---------------------------------------------------------------------------------------------
def _pattern_chi2(sim, tree, model, seed, n, k):
    """device pattern histogram of n free-running sites against exact pruning probabilities"""
    torch = pytest.importorskip("torch")
    from scipy import stats

    t = tree.n_leaves
    pitch = (n + 127) // 128 * 128
    buf = torch.empty((t, pitch), dtype=torch.uint8, device="cuda")
    sc = torch.zeros((t, k), dtype=torch.int64, device="cuda")
    pc = torch.zeros(k ** t, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    sim.simulate_device(seed, 0, n, buf.data_ptr(), pitch)
    sim.histogram_device(buf.data_ptr(), pitch, n, sc.data_ptr(), pc.data_ptr())
    sim.synchronize()
    obs = pc.cpu().numpy()
    assert obs.sum() == n
    probs = pruning.pattern_probabilities(model.is_mixture, model.weights, model.pi, sim.ptables(), tree.parent, tree.leaf_row)
    assert abs(probs.sum() - 1) < 1e-10
    exp = probs * n
    keep = exp >= 5
    # the rare patterns are pooled into one cell
    o = np.append(obs[keep], obs[~keep].sum())
    e = np.append(exp[keep], exp[~keep].sum())
    if e[-1] < 5:
        o, e = o[:-1], e[:-1]
    chi2 = float(((o - e) ** 2 / e).sum())
    return stats.chi2.sf(chi2, len(e) - 1), int(keep.sum())


@pytest.mark.parametrize("seed", [1, 2])
def test_quad_multi_cap_chi_square_vs_exact_pruning(seed):
    """8 taxa, >= 3 caps: the composition of caps (frontier nodes kept in slots, inner nodes summed out) against exact
    pattern probabilities over all 4^8 = 65 536 site patterns -- ground truth that shares nothing with the cap tables."""
    tree = pm.birth_death_tree(8, 1.0, 0.5, seed=10 + seed)
    tree = tree.scaled(0.6 / max(tree.height(), 1e-9))  # moderate divergence: many patterns with expected counts >= 5
    model = pm.get_phylo_model("GTR4[1,2,3,4,5,6]{0.2,0.3,0.4,0.1}", gamma=(4, 0.7))
    with _sim(model, tree) as sim:
        assert sim.quad_groups >= 3, sim.quad_groups
        pval, cells = _pattern_chi2(sim, tree, model, 1000 + seed, 40_000_000, 4)
    assert cells > 5000
    assert pval > 1e-4, pval


def test_quad_caterpillar_many_caps_chi_square():
    """A caterpillar forces caps that chain through kept internal nodes (every cap hangs below the previous one's frontier)."""
    s_ = "t0:0.05"
    for i in range(1, 8):
        s_ = "(%s,t%d:%g):%g" % (s_, i, 0.03 * i, 0.02)
    tree = pm.parse_newick(s_ + ";")
    model = pm.get_phylo_model("HKY[6.0]{0.2,0.3,0.3,0.2}", gamma=(4, 0.5))
    with _sim(model, tree) as sim:
        assert sim.quad_groups >= 3
        pval, cells = _pattern_chi2(sim, tree, model, 5, 30_000_000, 4)
    assert cells > 1000 and pval > 1e-4, (cells, pval)


@pytest.mark.parametrize("spec_str,gamma", [("LG", None), ("LG", (4, 0.5)), ("WAG", (4, 1.0))])
def test_protein_three_taxa_chi_square_vs_exact_pruning(spec_str, gamma):
    """The protein sampler on a 3-taxon tree with an internal node: all 20^3 = 8000 patterns against pruning."""
    tree = pm.parse_newick("((a:0.3,b:0.1):0.2,c:0.5);")
    model = pm.get_phylo_model(spec_str, gamma=gamma)
    with _sim(model, tree) as sim:
        pval, cells = _pattern_chi2(sim, tree, model, 12, 20_000_000, 20)
    assert cells > 3000 and pval > 1e-4, (cells, pval)


def test_protein_four_taxa_mixture_chi_square_vs_exact_pruning():
    """C10 profile mixture on 4 taxa (20^4 = 160 000 patterns, rare ones pooled)."""
    tree = pm.parse_newick("((a:0.1,b:0.3):0.05,(c:0.02,d:0.7):0.2):0.15;")
    model = pm.get_phylo_model(mixture_model="C10", global_normalization=True)
    with _sim(model, tree) as sim:
        pval, cells = _pattern_chi2(sim, tree, model, 3, 30_000_000, 20)
    assert cells > 5000 and pval > 1e-4, (cells, pval)
