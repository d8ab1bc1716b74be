"""Host-side mirrors of the model/tree construction the reference's caller performs before it reaches the
simulator (slynx/src/SLynx/Simulate/{PhyloModel,Simulate}.hs; elynx-markov MarkovProcess/*; elynx-tree Newick
and PointProcess).  Thin wrappers over the C++ implementation behind include/elynx_b200_model.h."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L

DNA, PROTEIN = "DNA", "Protein"
# elynx-seq Alphabet.hs:145-146,192-193: state index -> character = Set.elemAt over the sorted alphabet
ALPHABET_CHARS = {DNA: "ACGT", PROTEIN: "ACDEFGHIKLMNPQRSTVWY"}


def _vp(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


@dataclass
class PhyloModel:
    """PhyloModel = SubstitutionModel | MixtureModel (elynx-markov PhyloModel.hs:30-31), flattened to the
    arrays simulateAlignment hands the simulator (slynx Simulate.hs:104-116)."""

    is_mixture: bool
    alphabet: str
    name: str
    weights: np.ndarray  # [C]
    pi: np.ndarray  # [C][K]
    exch: np.ndarray  # [C][K][K]
    component_names: List[str]

    @property
    def n_states(self):
        return self.pi.shape[1]

    @property
    def n_components(self):
        return self.pi.shape[0]

    @property
    def alphabet_chars(self):
        return ALPHABET_CHARS[self.alphabet]


def _export(h) -> PhyloModel:
    lib = L.lib()
    k, c, mix, alph = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    L.check(lib.elx_phylo_model_info(h, C.byref(k), C.byref(c), C.byref(mix), C.byref(alph)))
    ws = np.empty(c.value)
    pi = np.empty((c.value, k.value))
    ex = np.empty((c.value, k.value, k.value))
    L.check(lib.elx_phylo_model_arrays(h, _vp(ws), _vp(pi), _vp(ex)))
    names = [lib.elx_phylo_model_name(h, i).decode() for i in range(c.value)]
    return PhyloModel(bool(mix.value), PROTEIN if alph.value == 1 else DNA, lib.elx_phylo_model_name(h, -1).decode(), ws, pi, ex, names)


def get_phylo_model(substitution_model: Optional[str] = None, mixture_model: Optional[str] = None, global_normalization: bool = False,
                    weights: Optional[Sequence[float]] = None, edm_components: Optional[Sequence[Tuple[float, Sequence[float]]]] = None,
                    gamma: Optional[Tuple[int, float]] = None) -> PhyloModel:
    """getPhyloModel (slynx PhyloModel.hs:211-231) followed, when `gamma=(n, alpha)` is given, by
    `expand n alpha` (GammaRateHeterogeneity.hs:47-51) -- the order simulateCmd uses (Simulate.hs:284-297)."""
    lib = L.lib()
    ws = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    ew = ep = None
    n_edm = k_edm = 0
    if edm_components is not None:
        n_edm = len(edm_components)
        ew = np.ascontiguousarray([c[0] for c in edm_components], dtype=np.float64)
        ep = np.ascontiguousarray([c[1] for c in edm_components], dtype=np.float64).reshape(n_edm, -1) if n_edm else np.zeros((0, 0))
        k_edm = ep.shape[1] if n_edm else 0
    h = C.c_void_p()
    L.check(lib.elx_phylo_model_parse(None if substitution_model is None else substitution_model.encode(),
                                      None if mixture_model is None else mixture_model.encode(), int(global_normalization),
                                      _vp(ws), 0 if ws is None else len(ws), _vp(ew), _vp(ep) if ep is not None else None, n_edm, k_edm,
                                      C.byref(h)))
    try:
        if gamma is not None:
            L.check(lib.elx_phylo_model_expand_gamma(h, int(gamma[0]), float(gamma[1])))
        return _export(h)
    finally:
        lib.elx_phylo_model_free(h)


def gamma_means(n: int, alpha: float) -> np.ndarray:
    """getMeans (GammaRateHeterogeneity.hs:88-107)"""
    if n <= 0:
        raise L.ElxError(L.ELX_E_INVALID, "getMeans: Number of rate categories is zero or negative.")
    out = np.empty(n)
    L.check(L.lib().elx_gamma_means(n, alpha, _vp(out)))
    return out


def _parse_components(fn, text: str):
    data = text.encode()
    n, k = C.c_int(), C.c_int()
    w, p = C.POINTER(C.c_double)(), C.POINTER(C.c_double)()
    L.check(fn(data, len(data), C.byref(n), C.byref(k), C.byref(w), C.byref(p)))
    try:
        ws = np.ctypeslib.as_array(w, shape=(max(n.value, 1),))[: n.value].copy()
        ps = np.ctypeslib.as_array(p, shape=(max(n.value * k.value, 1),))[: n.value * k.value].copy().reshape(n.value, k.value)
    finally:
        L.lib().elx_free(w)
        L.lib().elx_free(p)
    return [(float(ws[i]), ps[i]) for i in range(n.value)]


def parse_edm_phylobayes(text: str):
    """phylobayes (EDMModelPhylobayes.hs:33-40) -> [(weight, pi)]"""
    return _parse_components(L.lib().elx_parse_edm_phylobayes, text)


def parse_siteprofiles_phylobayes(text: str):
    """siteprofiles (SiteprofilesPhylobayes.hs:33-44) -> [(1.0, pi)]"""
    return _parse_components(L.lib().elx_parse_siteprofiles_phylobayes, text)


@dataclass
class FlatTree:
    """Preorder flattening of the reference's `Tree Length Name` (see include/elynx_b200.h)."""

    parent: np.ndarray
    brlen: np.ndarray
    leaf_row: np.ndarray
    names: List[str]

    @property
    def n_nodes(self):
        return len(self.parent)

    @property
    def n_leaves(self):
        return int((self.leaf_row >= 0).sum())

    def leaf_names(self) -> List[str]:
        """`leaves` order: DFS left to right (Rooted.hs:244-248)"""
        out = [""] * self.n_leaves
        for i, r in enumerate(self.leaf_row):
            if r >= 0:
                out[r] = self.names[i]
        return out

    def scaled(self, factor: float) -> "FlatTree":
        return FlatTree(self.parent.copy(), self.brlen * factor, self.leaf_row.copy(), list(self.names))

    def height(self) -> float:
        h = np.zeros(self.n_nodes)
        for i in range(self.n_nodes):
            h[i] = self.brlen[i] + (h[self.parent[i]] if self.parent[i] >= 0 else 0.0)
        return float(h[self.leaf_row >= 0].max())


def _export_tree(h, want_newick=False):
    lib = L.lib()
    n, t = C.c_int(), C.c_int()
    L.check(lib.elx_flat_tree_size(h, C.byref(n), C.byref(t)))
    parent = np.empty(n.value, dtype=np.int32)
    brlen = np.empty(n.value, dtype=np.float64)
    leaf_row = np.empty(n.value, dtype=np.int32)
    L.check(lib.elx_flat_tree_arrays(h, _vp(parent), _vp(brlen), _vp(leaf_row)))
    names = [lib.elx_flat_tree_name(h, i).decode() for i in range(n.value)]
    tree = FlatTree(parent, brlen, leaf_row, names)
    if not want_newick:
        return tree
    s = C.c_void_p()
    L.check(lib.elx_flat_tree_to_newick(h, C.byref(s)))
    try:
        text = C.string_at(s).decode()
    finally:
        lib.elx_free(s)
    return tree, text


def parse_newick(text: str) -> FlatTree:
    """newick Standard + toLengthTree (elynx-tree Import/Newick.hs:90-96, Phylogeny.hs:374-381)"""
    data = text.encode()
    h = C.c_void_p()
    L.check(L.lib().elx_newick_parse(data, len(data), C.byref(h)))
    try:
        return _export_tree(h)
    finally:
        L.lib().elx_flat_tree_free(h)


def birth_death_tree(n_leaves: int, birth_rate: float = 1.0, death_rate: float = 0.5, seed: int = 0, want_newick: bool = False):
    """simulateReconstructedTree n Random lambda mu (elynx-tree PointProcess.hs:245-259) -- what
    `tlynx simulate -n N birthdeath -l lambda -m mu` produces."""
    h = C.c_void_p()
    L.check(L.lib().elx_birth_death_tree(n_leaves, birth_rate, death_rate, C.c_uint64(seed), C.byref(h)))
    try:
        return _export_tree(h, want_newick)
    finally:
        L.lib().elx_flat_tree_free(h)
