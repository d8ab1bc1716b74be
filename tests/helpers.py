"""Shared builders for the tests (oracle-side model/tree construction)."""
import os

import numpy as np

import oracle.model as om

REF_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_data")


def tree_from_file(name):
    return om.parse_newick(open(os.path.join(REF_DATA, name)).read())


def random_tree(n_leaves, rng, mean_len=0.3, multifurcate=False, zero_frac=0.0):
    """Random rooted tree as a Newick string -> FlatTree (exercises the parser and odd shapes)."""
    nodes = ["t%d:%r" % (i, float(rng.exponential(mean_len))) for i in range(n_leaves)]
    while len(nodes) > 1:
        k = 2 if not multifurcate or len(nodes) < 3 else int(rng.integers(2, min(4, len(nodes)) + 1))
        idx = sorted(rng.choice(len(nodes), size=k, replace=False).tolist(), reverse=True)
        kids = [nodes.pop(i) for i in idx]
        bl = 0.0 if rng.random() < zero_frac else float(rng.exponential(mean_len))
        nodes.append("(" + ",".join(kids) + "):%r" % bl)
    s = nodes[0]
    s = s[: s.rindex(":")] + ";"  # no root stem
    return om.parse_newick(s)


def model_case(name):
    """(is_mixture, ws, ds, es, alphabet) for a named model configuration."""
    edm = om.parse_edm_phylobayes(open(os.path.join(REF_DATA, "EDMDistsPhylobayes.txt")).read())
    if name == "jc":
        pm = om.get_phylo_model("JC", None)
    elif name == "hky":
        pm = om.get_phylo_model("HKY[6.0]{0.2,0.3,0.3,0.2}", None)
    elif name == "hky_g4":
        pm = om.expand(4, 0.5, om.get_phylo_model("HKY[6.0]{0.2,0.3,0.3,0.2}", None))
    elif name == "gtr4_g4":
        pm = om.expand(4, 0.5, om.get_phylo_model("GTR4[1,2,3,4,5,6]{0.2,0.3,0.4,0.1}", None))
    elif name == "lg":
        pm = om.get_phylo_model("LG", None)
    elif name == "lg_g4":
        pm = om.expand(4, 0.5, om.get_phylo_model("LG", None))
    elif name == "wag_g4":
        pm = om.expand(4, 0.5, om.get_phylo_model("WAG", None))
    elif name == "c10":
        pm = om.get_phylo_model(None, "C10", global_norm=True)
    elif name == "c10_g4":
        pm = om.expand(4, 0.2, om.get_phylo_model(None, "C10", global_norm=True))
    elif name == "c60":
        pm = om.get_phylo_model(None, "C60", global_norm=True)
    elif name == "edm_lg":
        pm = om.get_phylo_model(None, "EDM(LG-Custom)", edm_cs=edm)
    elif name == "mixture_dna":
        pm = om.get_phylo_model(None, "MIXTURE(JC,HKY[6.0]{0.25,0.25,0.25,0.25})", mws=[0.4, 0.6])
    else:
        raise KeyError(name)
    is_mix, ws, ds, es = om.flatten_model(pm)
    return is_mix, ws, ds, es, pm.alphabet


def uniforms(rng, rows, s):
    """doubles in (0,1] like uniformDoublePositive01M: w64/2^64 + 2^-65"""
    w = rng.integers(0, 2 ** 64, size=(rows, s), dtype=np.uint64)
    return w.astype(np.float64) / 18446744073709551616.0 + 2.0 ** -65


def spec_freerun(sim, model, tree, seed, site0, s, **kw):
    """The executable spec of the product's free-running draw for this handle, run on the handle's own tables:
    pair mode (K <= 4: sibling pairs drawn jointly) or node by node."""
    import oracle.freerun_spec as spec

    is_mix, ws, ds = model[0], model[1], model[2]
    pair = sim.pair_tables()
    if pair is not None:
        return spec.freerun_pair(is_mix, tree.parent, tree.leaf_row, ws, ds, pair, seed, site0, s, **kw)
    e16, qlo = sim.alias_tables()
    return spec.freerun(is_mix, tree.parent, tree.leaf_row, ws, ds, e16, qlo, seed, site0, s, **kw)


def alias_masses(e, qlo, a):
    """Integer outcome masses (sum 2^48 per row) encoded by alias rows e = [q_hi : 16-a][alias : a], qlo."""
    kp = 1 << a
    cap = 1 << (48 - a)
    e = e.astype(np.int64)
    al = e & (kp - 1)
    q = ((e >> a) << 32) | qlo.astype(np.int64)
    q = np.where(al == np.arange(kp), cap, q)  # alias == self means "always accept"
    mass = np.zeros(e.shape, dtype=np.int64)
    for j in range(kp):
        mass[..., j] += q[..., j]
        np.put_along_axis(mass, al[..., j:j + 1], np.take_along_axis(mass, al[..., j:j + 1], -1) + (cap - q[..., j:j + 1]), -1)
    return mass


def check_pair_tables(p, pair, parent):
    """Every non-root node is in exactly one record, members of a record are siblings, records are in walk order, and
    the joint rows carry integer masses (sum 2^48) equal to P_A[i][sa] * P_B[i][sb] (normalised) within 2^-44."""
    node_a, node_b, pe16, pqlo = pair
    n, c, k, _ = p.shape
    seen = np.zeros(n, dtype=bool)
    for g, (a, b) in enumerate(zip(node_a, node_b)):
        assert 0 <= a < n and not seen[a] and (parent[a] < 0 or seen[parent[a]])
        seen[a] = True
        if b >= 0:
            assert not seen[b] and parent[b] == parent[a]
            seen[b] = True
    assert seen.all()
    assert node_a[0] == 0 and node_b[0] == -1
    mass = alias_masses(pe16, pqlo, 4)
    assert np.all(mass.sum(-1) == 1 << 48)
    pn = p / p.sum(-1, keepdims=True)
    pa = pn[node_a]  # [G][C][K][K]
    pb = np.where((node_b >= 0)[:, None, None, None], pn[np.maximum(node_b, 0)], (np.arange(k) == 0).astype(float))
    joint = np.zeros(pe16.shape)
    for sa in range(k):
        for sb in range(k):
            joint[..., sa | (sb << 2)] = pa[..., sa] * pb[..., sb]
    assert np.max(np.abs(mass / float(1 << 48) - joint)) < 2.0 ** -44
