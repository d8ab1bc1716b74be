"""Pin the CPU oracle (oracle/) against every known-answer test the reference holds for this path
(KAT-1..KAT-8) and against independent ground truth (scipy/mpmath expm,
closed-form gamma means).  CPU only."""
import os

import numpy as np
import pytest

import oracle.model as om
import oracle.sim as osim


def test_kat1_splitmix_jc_single_node(kats):
    # MarkovProcessAlongTreeSpec.hs:45-48: simulate 1 pi_JC E_JC (Node 0 []) (mkStdGen 0) -> Node 2 []
    sm = om.jc(om.DO_NORMALIZE)
    parent = np.array([-1], dtype=np.int32)
    leaf_row = np.array([0], dtype=np.int32)
    p = om.prob_tables(sm.stationary_distribution[None], sm.exchangeability_matrix[None], np.array([0.0]))
    _, leaves, internal = osim.simulate_splitmix(False, parent, leaf_row, np.ones(1), sm.stationary_distribution[None], p, 1, 0,
                                                 nchunks=0, keep_internal=True)
    assert leaves[0, 0] == kats["kat1_jc_single_node_mkstdgen0_state"] == internal[0, 0]
    # python restatement of the same stream agrees with the C one
    g = om.SplitMix.mk_std_gen(0)
    assert g.next_word64() == osim.lib().oracle_splitmix_first_word(0) == 0x9474F0EB06D79FD8


def test_kat1b_shape_ten_sites():
    # MarkovProcessAlongTreeSpec.hs:49-52
    sm = om.jc(om.DO_NORMALIZE)
    tree = om.parse_newick("(a:1.0,b:1.0):0.0;")
    p = om.prob_tables(sm.stationary_distribution[None], sm.exchangeability_matrix[None], tree.brlen)
    _, leaves, internal = osim.simulate_splitmix(False, tree.parent, tree.leaf_row, np.ones(1), sm.stationary_distribution[None],
                                                 p, 10, 0, nchunks=0, keep_internal=True)
    assert internal.shape == (3, 10) and leaves.shape == (2, 10)
    assert leaves.max() < 4


def test_kat23_exch_from_list(kats):
    es = kats["exch_list_1_10"]
    assert np.array_equal(om.exch_from_list_lower(5, es).ravel(), kats["exch_from_list_lower_5"])
    assert np.array_equal(om.exch_from_list_upper(5, es).ravel(), kats["exch_from_list_upper_5"])
    with pytest.raises(ValueError):
        om.exch_from_list_lower(5, es[:-1])


def test_kat4_hky_stationary(kats):
    sm = om.hky(om.DO_NORMALIZE, kats["hky_kappa"], kats["hky_pi"])
    sd = om.get_stationary_distribution(sm.rate_matrix())
    assert np.allclose(sd, kats["hky_pi"], rtol=0, atol=1e-12)
    assert abs(sm.total_rate() - 1.0) < 1e-12


def test_kat56_lg(kats):
    pi_py = np.array(kats["lg_pi_alpha_python_unnormalised"])
    pi_py = pi_py / pi_py.sum()
    sm = om.lg(om.DO_NORMALIZE)
    assert np.allclose(sm.stationary_distribution, pi_py, rtol=0, atol=1e-12)
    # AminoAcidSpec.hs:501-505 compares the exchangeabilities of `lg DoNormalize` to the python matrix at 1e-4
    assert np.allclose(sm.exchangeability_matrix, np.array(kats["lg_exch_alpha_python"]), rtol=0, atol=1e-4) or \
        np.allclose(om.lg_exch(), np.array(kats["lg_exch_alpha_python"]), rtol=0, atol=1e-4)
    assert np.allclose(om.get_stationary_distribution(sm.rate_matrix()), sm.stationary_distribution, atol=1e-4)
    u = np.full(20, 0.05)
    assert np.allclose(om.get_stationary_distribution(om.lg_custom(None, om.DO_NORMALIZE, u).rate_matrix()), u, atol=1e-12)
    assert np.allclose(om.get_stationary_distribution(om.poisson(om.DO_NORMALIZE).rate_matrix()), u, atol=1e-12)
    assert np.allclose(om.get_stationary_distribution(om.poisson_custom(None, om.DO_NORMALIZE, pi_py).rate_matrix()), pi_py, atol=1e-12)


def test_kat7_edm(kats, ref_data):
    cs = om.parse_edm_phylobayes(open(os.path.join(ref_data, "EDMDistsPhylobayes.txt")).read())
    assert len(cs) == 4
    for (w, pi), k in zip(cs, kats["edm_phylobayes_components"]):
        assert w == k["weight"] and np.array_equal(pi, k["pi"])


def test_kat8_siteprofiles(kats, ref_data):
    cs = om.parse_siteprofiles_phylobayes(open(os.path.join(ref_data, "HSSPMany.siteprofiles")).read())
    assert len(cs) == kats["siteprofiles_count"]
    assert all(w == 1.0 for w, _ in cs)
    assert all(abs(p.sum() - 1.0) < 1e-5 for _, p in cs)
    assert np.array_equal(cs[0][1], kats["siteprofiles_first"])


def test_newick_fixture(ref_data):
    t = om.parse_newick(open(os.path.join(ref_data, "Newick.tree")).read())
    assert t.n_leaves == 40 and t.n_nodes == 78
    assert abs(t.brlen.sum() - 15.60) < 0.01 and t.brlen[0] == 0.0
    assert t.leaf_names()[0] == "Aeropyrum0" and t.leaf_names()[-1] == "Sulfolobus"
    t2 = om.parse_newick(open(os.path.join(ref_data, "UltraMetric.tree")).read())
    assert t2.n_leaves == 10 and t2.leaf_names() == [str(i) for i in range(10)]
    assert t2.brlen[0] == pytest.approx(9.936372265738624e-4)
    with pytest.raises(ValueError):
        om.parse_newick("((a:1,b):1,c:1);")  # missing inner length -> toLengthTree error


def test_expm_vs_ground_truth():
    import scipy.linalg as sl
    for sm in (om.lg(om.DO_NORMALIZE), om.hky(om.DO_NORMALIZE, 6.0, [0.2, 0.3, 0.3, 0.2])):
        q = sm.rate_matrix()
        for t in (1e-4, 1e-2, 0.1, 1.0, 5.0, 20.0):
            p = om.prob_matrix(q, t)
            ref = sl.expm(t * q)
            assert np.max(np.abs(p - ref) / ref) < 2e-13
            assert np.allclose(p.sum(1), 1.0, atol=1e-13)
    assert np.array_equal(om.prob_matrix(q, 0.0), np.eye(4))
    with pytest.raises(ValueError):
        om.prob_matrix(q, -1.0)


def test_expm_vs_mpmath_50_digits():
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    q = om.lg(om.DO_NORMALIZE).rate_matrix()
    for t in (1e-4, 0.1, 5.0):
        ref = np.array(mp.expm(mp.matrix((t * q).tolist())).tolist(), dtype=np.float64)
        p = om.prob_matrix(q, t)
        assert np.max(np.abs(p - ref) / ref) < 1e-13


def test_gamma_means():
    m = om.gamma_means(4, 0.5)
    assert np.allclose(m, [0.033388, 0.251916, 0.820268, 2.894428], atol=1e-6)
    assert abs(m.mean() - 1.0) < 1e-14
    for n, a in ((4, 0.2), (4, 0.5), (8, 1.3), (3, 5.0)):
        assert np.allclose(om.gamma_means(n, a), om.gamma_means_by_integration(n, a), atol=1e-8)  # reference eps


def test_mixture_wiring():
    sm = om.get_phylo_model("HKY[6.0]{0.2,0.3,0.3,0.2}", None)
    mm = om.expand(4, 0.5, sm)
    assert mm.weights == [1.0] * 4 and len(mm.models) == 4
    r = om.gamma_means(4, 0.5)
    for k in range(4):
        assert np.allclose(mm.models[k].exchangeability_matrix, r[k] * sm.exchangeability_matrix)
    c10 = om.get_phylo_model(None, "C10", global_norm=True)
    tot = sum(w * m.total_rate() for w, m in zip(c10.weights, c10.models)) / sum(c10.weights)
    assert abs(tot - 1.0) < 1e-12
    with pytest.raises(ValueError):
        om.get_phylo_model(None, "C10")  # local normalisation impossible (PhyloModel.hs:180)
    g = om.expand(4, 0.2, c10)
    assert len(g.models) == 40 and g.weights[:10] == c10.weights and g.weights[10:20] == c10.weights  # category-major
    mix = om.get_phylo_model(None, "MIXTURE(JC,HKY[6.0]{0.25,0.25,0.25,0.25})", mws=[0.4, 0.6])
    assert [m.name for m in mix.models] == ["JC", "HKY"] and all(abs(m.total_rate() - 1) < 1e-12 for m in mix.models)
    with pytest.raises(ValueError):
        om.get_phylo_model("HKY[6.0]{0.2,0.3,0.3,0.3}", None)  # pi literal must sum to 1 within 1e-12


def test_c_and_python_walks_agree(ref_data):
    rng = np.random.default_rng(5)
    tree = om.parse_newick(open(os.path.join(ref_data, "UltraMetric.tree")).read())
    mm = om.expand(4, 0.5, om.get_phylo_model("GTR4[1,2,3,4,5,6]{0.2,0.3,0.4,0.1}", None))
    is_mix, ws, ds, es = om.flatten_model(mm)
    p = om.prob_tables(ds, es, tree.brlen)
    u = 1.0 - rng.random((tree.n_nodes + 2, 200))
    c1, l1, s1 = om.simulate_injected_py(is_mix, ws, ds, p, tree, u)
    c2, l2, s2 = osim.simulate_injected(is_mix, tree.parent, tree.leaf_row, ws, ds, p, u, keep_internal=True)
    assert np.array_equal(c1, c2) and np.array_equal(l1, l2) and np.array_equal(s1, s2)


def test_get_chunks_and_par_stream():
    assert om.get_chunks(3, 10) == [4, 3, 3] and om.get_chunks(4, 2) == [1, 1, 0, 0]
    sm = om.jc(om.DO_NORMALIZE)
    tree = om.parse_newick("((a:0.1,b:0.2):0.3,c:0.4);")
    p = om.prob_tables(sm.stationary_distribution[None], sm.exchangeability_matrix[None], tree.brlen)
    a = osim.simulate_splitmix(False, tree.parent, tree.leaf_row, np.ones(1), sm.stationary_distribution[None], p, 1000, 7,
                               nchunks=4, nthreads=4)[1]
    b = osim.simulate_splitmix(False, tree.parent, tree.leaf_row, np.ones(1), sm.stationary_distribution[None], p, 1000, 7,
                               nchunks=4, nthreads=1)[1]
    assert np.array_equal(a, b)  # threading does not change the chunked stream
    cum = np.cumsum(p, axis=-1)
    c = osim.simulate_splitmix(False, tree.parent, tree.leaf_row, np.ones(1), sm.stationary_distribution[None], cum, 1000, 7,
                               nchunks=4, nthreads=2, precomputed_cum=True)[1]
    assert np.array_equal(a, c)


def test_fasta_and_double_dec():
    out = om.sequences_to_fasta(["x", "yy"], np.array([[0, 1, 2, 3], [3, 3, 0, 0]]), om.DNA)
    assert out == b">x\nACGT\n>yy\nTTAA\n"
    assert om._double_dec(0.1) == "0.1" and om._double_dec(6.10600573606971e-2) == "6.10600573606971e-2"
    assert om._double_dec(1.0) == "1.0" and om._double_dec(12345678.0) == "1.2345678e7" and om._double_dec(0.05) == "5.0e-2"
