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
