"""The product's C++ host-side model algebra / parsers / tree helpers (include/elynx_b200_model.h) against the
oracle's numpy restatement and the reference's known-answer tests.  CPU only -- nothing here needs a GPU."""
import os

import numpy as np
import pytest

import oracle.model as om
from elynx_b200 import ElxError
from elynx_b200 import model as pm


@pytest.fixture(scope="module", autouse=True)
def built():
    from elynx_b200 import build

    build.build()


def same_model(a: pm.PhyloModel, b, tol=1e-13):
    is_mix, ws, ds, es = om.flatten_model(b)
    assert a.is_mixture == is_mix and a.alphabet == b.alphabet
    assert a.pi.shape == ds.shape and a.exch.shape == es.shape
    assert np.allclose(a.weights, ws, rtol=0, atol=tol)
    assert np.allclose(a.pi, ds, rtol=0, atol=tol)
    assert np.allclose(a.exch, es, rtol=tol, atol=tol)


SLYNX_TEST_MATRIX = [  # scripts/slynx-test:25-50 (the CLI acceptance matrix), following the code where the script is stale
    dict(s="HKY[6.0]{0.2,0.3,0.3,0.2}"),
    dict(s="GTR4[1,2,3,4,5,6]{0.2,0.3,0.4,0.1}"),
    dict(m="EDM(LG-Custom)", edm=True),
    dict(m="EDM(LG-Custom)", edm=True, w=[0.9, 0.05, 0.03, 0.02]),
    dict(m="C20", n=True),
    dict(s="HKY[6.0]{0.2,0.3,0.3,0.2}", g=(4, 0.2)),
    dict(m="C10", n=True, g=(4, 0.2)),
    dict(m="MIXTURE(JC,HKY[6.0]{0.25,0.25,0.25,0.25})", w=[0.4, 0.6]),
    dict(m="EDM(Poisson-Custom)", prof=True),
    dict(s="JC"), dict(s="LG", g=(4, 0.5)), dict(s="WAG", g=(4, 0.5)), dict(m="C60", n=True),
    dict(m="EDM(LG-Custom)", edm=True, n=True), dict(s="F81{0.1,0.2,0.3,0.4}"), dict(s="Poisson"),
    dict(m="MIXTURE(LG,WAG,Poisson)", w=[1, 2, 3], n=True),
]


@pytest.mark.parametrize("case", SLYNX_TEST_MATRIX, ids=lambda c: str(c))
def test_model_grammar_matches_oracle(case, ref_data):
    edm_txt = open(os.path.join(ref_data, "EDMDistsPhylobayes.txt")).read()
    prof_txt = open(os.path.join(ref_data, "HSSPMany.siteprofiles")).read()
    comps_o = comps_p = None
    if case.get("edm"):
        comps_o, comps_p = om.parse_edm_phylobayes(edm_txt), pm.parse_edm_phylobayes(edm_txt)
    if case.get("prof"):
        comps_o, comps_p = om.parse_siteprofiles_phylobayes(prof_txt)[:50], pm.parse_siteprofiles_phylobayes(prof_txt)[:50]
    ref = om.get_phylo_model(case.get("s"), case.get("m"), global_norm=case.get("n", False), mws=case.get("w"), edm_cs=comps_o)
    if case.get("g"):
        ref = om.expand(case["g"][0], case["g"][1], ref)
    got = pm.get_phylo_model(case.get("s"), case.get("m"), global_normalization=case.get("n", False), weights=case.get("w"),
                             edm_components=comps_p, gamma=case.get("g"))
    same_model(got, ref)


def test_kats(kats, ref_data):
    lg = pm.get_phylo_model("LG")
    pi_py = np.array(kats["lg_pi_alpha_python_unnormalised"])
    assert np.allclose(lg.pi[0], pi_py / pi_py.sum(), atol=1e-12)  # KAT-5
    poisson_lgpi = pm.get_phylo_model("Poisson-Custom{" + ",".join(repr(float(x)) for x in om.lg_stat_dist()) + "}")
    assert np.allclose(poisson_lgpi.pi[0], om.lg_stat_dist(), atol=1e-15)
    cs = pm.parse_edm_phylobayes(open(os.path.join(ref_data, "EDMDistsPhylobayes.txt")).read())  # KAT-7
    assert len(cs) == 4
    for (w, p), k in zip(cs, kats["edm_phylobayes_components"]):
        assert w == k["weight"] and np.array_equal(p, k["pi"])
    prof = pm.parse_siteprofiles_phylobayes(open(os.path.join(ref_data, "HSSPMany.siteprofiles")).read())  # KAT-8
    assert len(prof) == 701 and all(w == 1.0 for w, _ in prof) and np.array_equal(prof[0][1], kats["siteprofiles_first"])
    hky = pm.get_phylo_model("HKY[6.0]{0.2,0.3,0.3,0.2}")  # KAT-4: pi is stationary for Q
    q = om.from_exchangeability_matrix(hky.exch[0], hky.pi[0])
    assert np.allclose(hky.pi[0] @ q, 0, atol=1e-15) and abs(om.total_rate(q) - 1) < 1e-12


def test_gamma_means():
    for n, a in ((4, 0.5), (4, 0.2), (8, 1.3), (3, 5.0), (1, 0.7), (16, 0.05)):
        assert np.allclose(pm.gamma_means(n, a), om.gamma_means(n, a), rtol=0, atol=1e-11), (n, a)
    with pytest.raises(ElxError, match="zero or negative"):
        pm.gamma_means(0, 1.0)


@pytest.mark.parametrize("bad,msg", [
    (dict(), "No model was given"),
    (dict(substitution_model="JC", mixture_model="C10"), "use only one"),
    (dict(substitution_model="JC", global_normalization=True), "Global normalization not possible"),
    (dict(substitution_model="JC", weights=[1.0]), "Weights given"),
    (dict(mixture_model="C10"), "Local normalization impossible"),
    (dict(mixture_model="C15", global_normalization=True), "cannot create CXX model"),
    (dict(mixture_model="C10", global_normalization=True, weights=[1.0]), "Number of weights does not match"),
    (dict(mixture_model="MIXTURE(JC,LG)", weights=[1, 1]), "alphabets"),
    (dict(mixture_model="MIXTURE(JC)"), "No weights provided"),
    (dict(substitution_model="HKY[6.0]{0.2,0.3,0.3,0.3}"), "should be 1.0"),
    (dict(substitution_model="HKY[6.0]{0.5,0.5}"), "Length of stationary distribution"),
    (dict(substitution_model="FOO"), "Cannot assemble"),
    (dict(substitution_model="GTR4[1,2,3]{0.25,0.25,0.25,0.25}"), "number of exchangeabilities"),
])
def test_reference_error_cases(bad, msg):
    with pytest.raises(ElxError, match=msg):
        pm.get_phylo_model(**bad)


def test_newick_matches_oracle(ref_data):
    for name in ("Newick.tree", "UltraMetric.tree"):
        txt = open(os.path.join(ref_data, name)).read()
        a, b = pm.parse_newick(txt), om.parse_newick(txt)
        assert np.array_equal(a.parent, b.parent) and np.array_equal(a.brlen, b.brlen) and np.array_equal(a.leaf_row, b.leaf_row)
        assert a.names == b.names and a.leaf_names() == b.leaf_names()
    # whitespace is legal only where the reference grammar skips it: before names, around ':', after ','
    t = pm.parse_newick(" ( 'a b' : 1, (c:2.5e-1[90], d:0.5)x:1e0)root;\n")
    o = om.parse_newick(" ( 'a b' : 1, (c:2.5e-1[90], d:0.5)x:1e0)root;\n")
    assert o.names == t.names and o.brlen.tolist() == t.brlen.tolist()
    assert t.names == ["root", "a b", "x", "c", "d"] and t.brlen.tolist() == [0.0, 1.0, 1.0, 0.25, 0.5]
    for bad, msg in (("((a:1,b):1,c:1);", "Length unavailable"), ("(a:1,b:-1);", "negative"), ("(a:1,b:1)", "expected ';'"),
                     ("(a:1,b:1;", "expected"), ("(a:1 ,b:1);", "expected")):
        with pytest.raises(ElxError, match=msg):
            pm.parse_newick(bad)


def test_birth_death_tree_shape():
    """tlynx simulate -n 1000 birthdeath -l 1 -m 0.5: SURVEY appendix B.2 ranges for the restated point process."""
    tree, nwk = pm.birth_death_tree(1000, 1.0, 0.5, seed=1, want_newick=True)
    assert tree.n_leaves == 1000 and tree.n_nodes == 1999
    assert tree.leaf_names() == [str(i) for i in range(1000)]
    # ultrametric: every leaf at the same height = time of origin
    h = np.zeros(tree.n_nodes)
    for i in range(tree.n_nodes):
        h[i] = tree.brlen[i] + (h[tree.parent[i]] if tree.parent[i] >= 0 else 0)
    hl = h[tree.leaf_row >= 0]
    assert np.allclose(hl, hl[0], rtol=1e-12) and 8 < hl[0] < 20
    assert 0.5 < tree.brlen.mean() < 0.9 and np.all(tree.brlen >= 0)
    # the Newick writer round-trips through both parsers (Export/Newick.hs format, doubleDec numbers)
    back = om.parse_newick(nwk)
    assert np.array_equal(back.parent, tree.parent) and np.array_equal(back.brlen, tree.brlen)
    assert pm.birth_death_tree(1000, 1.0, 0.5, seed=1).brlen.tolist() == tree.brlen.tolist()  # deterministic in the seed
    assert pm.birth_death_tree(50, 1.0, 1.0, seed=3).n_nodes == 99  # critical process
