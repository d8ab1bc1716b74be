"""CPU ORACLE (test infrastructure only) -- numpy restatement of the reference's model algebra.

This file is NOT part of the product.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it.  The product (elynx_b200/) never does.

Every function cites the reference lines (relative to /root/reference) it restates.  The
reference is Haskell and cannot be compiled in this image (no GHC); parity is pinned against
the reference's own known-answer tests (tests/golden/reference_kats.json, made by
tools/gen_model_data.py) -- see tests/test_oracle_kats.py.  Three third-party pieces are
restated from their published algorithms because they are not vendored under /root/reference:

* hmatrix `expm` (Golub & Van Loan Alg. 11.3.1, scaled Pade, q = 7) -- `expm_hmatrix`.
  PARITY UNPINNED: no reference test evaluates `probMatrix`; checked against scipy/mpmath instead.
* mwc-random `categorical` (scanl1 (+); p = last * u; first index with cv >= p) -- `categorical`.
  Pinned only by KAT-1 (MarkovProcessAlongTreeSpec.hs:45-48) together with SplitMix below.
* random's StdGen = SplitMix + uniformDoublePositive01M -- `SplitMix`.
* statistics/integration gamma quantiles and tanh-sinh integral -- `gamma_means` uses the
  closed form (Yang 1994); PARITY UNPINNED, reference tolerance is its own eps = 1e-8.
"""
from __future__ import annotations

import json
import math
import os
import re
from dataclasses import dataclass, replace
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(_HERE, "data", "model_data.json")) as _f:
    _DATA = json.load(_f)

DNA = "DNA"
PROTEIN = "Protein"
# elynx-seq/src/ELynx/Alphabet/Alphabet.hs:145-146,192-193: state index -> char = Set.elemAt over sorted chars
ALPHABET_CHARS = {DNA: "ACGT", PROTEIN: "ACDEFGHIKLMNPQRSTVWY"}

DO_NORMALIZE = "DoNormalize"
DO_NOT_NORMALIZE = "DoNotNormalize"

# --------------------------------------------------------------------------------------
# RateMatrix.hs
# --------------------------------------------------------------------------------------
EPS_RELAXED = 1e-5  # RateMatrix.hs:53-54
EPS = 1e-12  # RateMatrix.hs:105-106


def is_valid(d) -> bool:
    """RateMatrix.hs:57-58"""
    return EPS_RELAXED > abs(np.sum(np.abs(d)) - 1.0)


def normalize_sd(d):
    """RateMatrix.hs:61-62"""
    d = np.asarray(d, dtype=np.float64)
    return d / np.sum(np.abs(d))


def _set_diag_zero(m):
    """RateMatrix.hs:64-66"""
    m = np.array(m, dtype=np.float64)
    return m - np.diag(np.diag(m))


def set_diagonal(m):
    """RateMatrix.hs:87-91: rows sum to zero; row sums are norm_1 (abs sums) of the off-diagonals."""
    z = _set_diag_zero(m)
    return z - np.diag(np.sum(np.abs(z), axis=1))


def from_exchangeability_matrix(em, d):
    """RateMatrix.hs:101-103: Q = setDiagonal (E <> diag pi)"""
    em = np.asarray(em, dtype=np.float64)
    d = np.asarray(d, dtype=np.float64)
    return set_diagonal(em @ np.diag(d))


def get_stationary_distribution(m):
    """RateMatrix.hs:121-130: eigenvector of Q^T for the eigenvalue of smallest magnitude."""
    vals, vecs = np.linalg.eig(np.asarray(m, dtype=np.float64).T)
    i = int(np.argmin(np.abs(vals)))
    if not EPS > abs(vals[i]):
        raise ValueError("getStationaryDistribution: Could not retrieve stationary distribution.")
    v = np.real(vecs[:, i])
    return v / np.sum(v)


def total_rate_with(d, m):
    """RateMatrix.hs:69-70: norm_1 (d <# offdiag m)"""
    return float(np.sum(np.abs(np.asarray(d) @ _set_diag_zero(m))))


def total_rate(m):
    """RateMatrix.hs:73-74"""
    return total_rate_with(get_stationary_distribution(m), m)


def _choose2(i: int) -> int:
    return i * (i - 1) // 2


def exch_from_list_lower(n: int, es: Sequence[float]):
    """RateMatrix.hs:150-154,176-187,219-220: k = C(i,2) + j for i > j (PAML order)."""
    if len(es) != _choose2(n):
        raise ValueError("exchFromListlower: the number of exchangeabilities does not match the matrix size")
    m = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            if i > j:
                m[i, j] = es[_choose2(i) + j]
            elif i < j:
                m[i, j] = es[_choose2(j) + i]
    return m


def exch_from_list_upper(n: int, es: Sequence[float]):
    """RateMatrix.hs:168-172,191-200,224-225: k = i*(n-2) - C(i,2) + j - 1 for i < j."""
    if len(es) != _choose2(n):
        raise ValueError("exchFromListlower: the number of exchangeabilities does not match the matrix size")
    m = np.zeros((n, n))

    def k(i, j):
        return i * (n - 2) - _choose2(i) + j - 1

    for i in range(n):
        for j in range(n):
            if i < j:
                m[i, j] = es[k(i, j)]
            elif i > j:
                m[i, j] = es[k(j, i)]
    return m


# --------------------------------------------------------------------------------------
# SubstitutionModel.hs
# --------------------------------------------------------------------------------------
@dataclass
class SubstitutionModel:
    """SubstitutionModel.hs:65-77"""

    alphabet: str
    name: str
    params: List[float]
    stationary_distribution: np.ndarray
    exchangeability_matrix: np.ndarray

    def rate_matrix(self):
        """SubstitutionModel.hs:80-84"""
        return from_exchangeability_matrix(self.exchangeability_matrix, self.stationary_distribution)

    def total_rate(self):
        """SubstitutionModel.hs:87-88 (recomputes pi by eig -- the reference's quirk)."""
        return total_rate(self.rate_matrix())


def sm_scale(r: float, sm: SubstitutionModel) -> SubstitutionModel:
    """SubstitutionModel.hs:122-125"""
    return replace(sm, exchangeability_matrix=r * sm.exchangeability_matrix)


def sm_normalize(sm: SubstitutionModel) -> SubstitutionModel:
    """SubstitutionModel.hs:129-130"""
    return sm_scale(1.0 / sm.total_rate(), sm)


def substitution_model(alphabet, name, params, nz, d, e) -> SubstitutionModel:
    """SubstitutionModel.hs:101-119"""
    d = np.asarray(d, dtype=np.float64)
    if not is_valid(d):
        raise ValueError("substitionModel: Stationary distribution does not sum to 1.0: %r" % (d,))
    sm = SubstitutionModel(alphabet, name, list(params), d / np.sum(d), np.array(e, dtype=np.float64))
    return sm_normalize(sm) if nz == DO_NORMALIZE else sm


# --------------------------------------------------------------------------------------
# Nucleotide.hs
# --------------------------------------------------------------------------------------
def _jc_exch():
    """Nucleotide.hs:59-78"""
    return np.ones((4, 4)) - np.eye(4)


def jc(nz):
    """Nucleotide.hs:84-85"""
    return substitution_model(DNA, "JC", [], nz, np.full(4, 0.25), _jc_exch())


def f81(nz, d):
    """Nucleotide.hs:88-89"""
    return substitution_model(DNA, "F81", [], nz, d, _jc_exch())


def hky(nz, k, d):
    """Nucleotide.hs:91-98; state order A,C,G,T; kappa at (A,G),(C,T)."""
    e = np.array([[0, 1, k, 1], [1, 0, 1, k], [k, 1, 0, 1], [1, k, 1, 0]], dtype=np.float64)
    return substitution_model(DNA, "HKY", [k], nz, d, e)


def gtr4(nz, es, d):
    """Nucleotide.hs:101-104"""
    return substitution_model(DNA, "GTR", list(es), nz, d, exch_from_list_upper(4, es))


# --------------------------------------------------------------------------------------
# AminoAcid.hs
# --------------------------------------------------------------------------------------
AA_ALPHA_ORDER = "ACDEFGHIKLMNPQRSTVWY"  # AminoAcid.hs:46-71
AA_PAML_ORDER = "ARNDCQEGHILKMFPSTWYV"  # AminoAcid.hs:74-99
_ALPHA_TO_PAML = [AA_PAML_ORDER.index(c) for c in AA_ALPHA_ORDER]  # :101-107
_PAML_TO_ALPHA = [AA_ALPHA_ORDER.index(c) for c in AA_PAML_ORDER]  # :109-115


def paml_to_alpha_vec(v):
    """AminoAcid.hs:118-119"""
    v = np.asarray(v, dtype=np.float64)
    return v[_ALPHA_TO_PAML]


def alpha_to_paml_vec(v):
    """AminoAcid.hs:122-123"""
    v = np.asarray(v, dtype=np.float64)
    return v[_PAML_TO_ALPHA]


def paml_to_alpha_mat(m):
    """AminoAcid.hs:126-130"""
    m = np.asarray(m, dtype=np.float64)
    return m[np.ix_(_ALPHA_TO_PAML, _ALPHA_TO_PAML)]


def _norm_sum(v):
    v = np.asarray(v, dtype=np.float64)
    return v / np.sum(v)


def lg_exch():
    """AminoAcid.hs:329-330"""
    return paml_to_alpha_mat(exch_from_list_lower(20, _DATA["lg_exch_paml_lower"]))


def lg_stat_dist():
    """AminoAcid.hs:339-367"""
    return paml_to_alpha_vec(_norm_sum(_DATA["lg_pi_paml"]))


def wag_exch():
    """AminoAcid.hs:575-576"""
    return paml_to_alpha_mat(exch_from_list_lower(20, _DATA["wag_exch_paml_lower"]))


def wag_stat_dist():
    """AminoAcid.hs:579-607"""
    return paml_to_alpha_vec(_norm_sum(_DATA["wag_pi_paml"]))


def lg(nz):
    """AminoAcid.hs:370-371"""
    return substitution_model(PROTEIN, "LG", [], nz, lg_stat_dist(), lg_exch())


def lg_custom(name, nz, d):
    """AminoAcid.hs:374-377"""
    return substitution_model(PROTEIN, name or "LG-Custom", [], nz, d, lg_exch())


def wag(nz):
    """AminoAcid.hs:610-611"""
    return substitution_model(PROTEIN, "WAG", [], nz, wag_stat_dist(), wag_exch())


def wag_custom(name, nz, d):
    """AminoAcid.hs:614-617"""
    return substitution_model(PROTEIN, name or "WAG-Custom", [], nz, d, wag_exch())


def _poisson_exch():
    """AminoAcid.hs:623-627"""
    return np.ones((20, 20)) - np.eye(20)


def poisson(nz):
    """AminoAcid.hs:633-634"""
    return substitution_model(PROTEIN, "Poisson", [], nz, np.full(20, 1.0 / 20), _poisson_exch())


def poisson_custom(name, nz, d):
    """AminoAcid.hs:637-640"""
    return substitution_model(PROTEIN, name or "Poisson-Custom", [], nz, d, _poisson_exch())


def gtr20(nz, es, d):
    """AminoAcid.hs:643-646"""
    return substitution_model(PROTEIN, "GTR", list(es), nz, d, exch_from_list_upper(20, es))


# --------------------------------------------------------------------------------------
# MixtureModel.hs
# --------------------------------------------------------------------------------------
@dataclass
class MixtureModel:
    """MixtureModel.hs:44-57 (components = parallel lists of weight and substitution model)"""

    name: str
    alphabet: str
    weights: List[float]
    models: List[SubstitutionModel]


def _normalize_globally(ws, sms):
    """MixtureModel.hs:67-72"""
    cks = [sm.total_rate() for sm in sms]
    c = sum(w * ck for w, ck in zip(ws, cks)) / sum(ws)
    return [sm_scale(1.0 / c, sm) for sm in sms]


def mm_from_substitution_models(name, nz, ws, sms) -> MixtureModel:
    """MixtureModel.hs:78-97"""
    ws = [float(w) for w in ws]
    sms = list(sms)
    if len(ws) == 0:
        raise ValueError("fromSubstitutionModels: No weights given.")
    if len(ws) != len(sms):
        raise ValueError("fromSubstitutionModels: Number of weights and substitution models does not match.")
    if any(sm.alphabet != sms[0].alphabet for sm in sms):
        raise ValueError("fromSubstitutionModels: alphabets of substitution models are not equal.")
    if nz == DO_NORMALIZE:
        sms = _normalize_globally(ws, sms)
    return MixtureModel(name, sms[0].alphabet, ws, sms)


def mm_concatenate(name, mms: Sequence[MixtureModel]) -> MixtureModel:
    """MixtureModel.hs:100-105"""
    ws = [w for mm in mms for w in mm.weights]
    sms = [sm for mm in mms for sm in mm.models]
    return mm_from_substitution_models(name, DO_NOT_NORMALIZE, ws, sms)


def mm_scale(s: float, mm: MixtureModel) -> MixtureModel:
    """MixtureModel.hs:107-115"""
    return replace(mm, models=[sm_scale(s, sm) for sm in mm.models])


# --------------------------------------------------------------------------------------
# GammaRateHeterogeneity.hs
# --------------------------------------------------------------------------------------
def gamma_means(n: int, alpha: float):
    """GammaRateHeterogeneity.hs:88-107 -- mean rate of each of the n equal-probability categories of
    Gamma(shape alpha, scale 1/alpha).  The reference integrates n*x*f(x) numerically (tanh-sinh,
    abs tol 1e-8, :109-123); this is the closed form r_i = n [P(alpha+1, alpha q_{i+1}) - P(alpha+1, alpha q_i)]
    (regularised lower incomplete gamma), equal to the reference within its own tolerance."""
    from scipy.special import gammainc, gammaincinv

    if n <= 0:
        raise ValueError("getMeans: Number of rate categories is zero or negative.")
    qs = [0.0] + [float(gammaincinv(alpha, i / n)) / alpha for i in range(1, n)] + [math.inf]
    cdf = [0.0 if q == 0.0 else (1.0 if math.isinf(q) else float(gammainc(alpha + 1.0, alpha * q))) for q in qs]
    return np.array([n * (cdf[i + 1] - cdf[i]) for i in range(n)])


def gamma_means_by_integration(n: int, alpha: float):
    """Same quantity by numerical integration of n*x*density (the reference's route), for cross-checks."""
    from scipy import integrate
    from scipy.stats import gamma as G

    g = G(alpha, scale=1.0 / alpha)
    qs = [g.ppf(i / n) for i in range(n)] + [math.inf]
    return np.array([integrate.quad(lambda x: n * x * g.pdf(x), qs[i], qs[i + 1], epsabs=1e-13, epsrel=1e-13, limit=500)[0]
                     for i in range(n)])


def expand_substitution_model(n, alpha, sm: SubstitutionModel) -> MixtureModel:
    """GammaRateHeterogeneity.hs:61-74: n components `scale r_k sm`, weights all 1.0, DoNotNormalize."""
    means = gamma_means(n, alpha)
    sms = [replace(sm_scale(float(r), sm), name=sm.name + "; gamma rate category %d" % (i + 1)) for i, r in enumerate(means)]
    return mm_from_substitution_models(sm.name + " with discrete gamma rate heterogeneity", DO_NOT_NORMALIZE, [1.0] * n, sms)


def expand_mixture_model(n, alpha, mm: MixtureModel) -> MixtureModel:
    """GammaRateHeterogeneity.hs:76-83: category-major concatenation of scaled copies."""
    means = gamma_means(n, alpha)
    return mm_concatenate(mm.name + " with discrete gamma rate heterogeneity", [mm_scale(float(r), mm) for r in means])


def expand(n, alpha, pm):
    """GammaRateHeterogeneity.hs:47-51"""
    if isinstance(pm, SubstitutionModel):
        return expand_substitution_model(n, alpha, pm)
    return expand_mixture_model(n, alpha, pm)


# --------------------------------------------------------------------------------------
# CXXModels.hs
# --------------------------------------------------------------------------------------
def cxx(n: int, mws: Optional[Sequence[float]] = None) -> MixtureModel:
    """CXXModels.hs:31-45,115-148: Poisson exchangeabilities, normalizeSD pi, global normalisation."""
    if n not in (10, 20, 30, 40, 50, 60):
        raise ValueError("cxx: cannot create CXX model with %d components." % n)
    ds = _DATA["c%d_pi" % n]
    ws = _DATA["c%d_w" % n] if mws is None else list(mws)
    if len(ws) != n:
        raise ValueError("Number of weights does not match C%d model." % n)
    sms = [poisson_custom("C%d; component %d" % (n, i + 1), DO_NOT_NORMALIZE, normalize_sd(d)) for i, d in enumerate(ds)]
    return mm_from_substitution_models("C%d" % n, DO_NORMALIZE, ws, sms)


# --------------------------------------------------------------------------------------
# Import/MarkovProcess/*.hs
# --------------------------------------------------------------------------------------
_HSPACE = " \t"


def parse_edm_phylobayes(text: str) -> List[Tuple[float, np.ndarray]]:
    """EDMModelPhylobayes.hs:33-64"""
    m = re.match(r"(\d+)[ \t]*(A C D E F G H I K L M N P Q R S T V W Y|A C G T)\s*", text)
    if not m:
        raise ValueError("phylobayes: headerLine")
    n = int(m.group(1))
    rest = text[m.end():]
    m = re.match(r"(\d+)\s*", rest)
    if not m:
        raise ValueError("phylobayes: kComponentsLine")
    k = int(m.group(1))
    lines = [ln for ln in rest[m.end():].splitlines() if ln.strip()]
    if len(lines) != k:
        raise ValueError("phylobayes: expected %d data lines, found %d" % (k, len(lines)))
    out = []
    for ln in lines:
        xs = [float(x) for x in ln.split()]
        if len(xs) - 1 != n:
            raise ValueError("Did not find correct number of entries.")
        out.append((xs[0], np.array(xs[1:])))
    return out


def parse_siteprofiles_phylobayes(text: str) -> List[Tuple[float, np.ndarray]]:
    """SiteprofilesPhylobayes.hs:33-62: header line ignored; `siteNo pi...`; weight 1.0 each."""
    lines = text.splitlines()[1:]
    out = []
    for ln in lines:
        if not ln.strip():
            continue
        xs = ln.split()
        int(xs[0])
        out.append((1.0, np.array([float(x) for x in xs[1:]])))
    if len({len(p) for _, p in out}) != 1:
        raise ValueError("The site profiles have a different number of entries.")
    return out


# --------------------------------------------------------------------------------------
# slynx/src/SLynx/Simulate/PhyloModel.hs -- model string grammar
# --------------------------------------------------------------------------------------
_NAME_STOP = "[]{}(),"


def _parse_name(s, i):
    j = i
    while j < len(s) and s[j] not in _NAME_STOP:
        j += 1
    if j == i:
        raise ValueError("model string: expected a name at %d in %r" % (i, s))
    return s[i:j], j


def _parse_list(s, i, open_c, close_c):
    if i >= len(s) or s[i] != open_c:
        return None, i
    j = s.index(close_c, i)
    xs = [float(x) for x in s[i + 1:j].split(",")]
    return xs, j + 1


def assemble_substitution_model(nz, n, mps, mf) -> SubstitutionModel:
    """PhyloModel.hs:108-143"""
    def need(d, k):
        if len(d) != k:
            raise ValueError("Length of stationary distribution is %d but should be %d." % (len(d), k))
        return d

    P, D = mps is not None, mf is not None
    if n == "JC" and not P and not D:
        return jc(nz)
    if n == "F81" and not P and D:
        return f81(nz, need(mf, 4))
    if n == "HKY" and P and len(mps) == 1 and D:
        return hky(nz, mps[0], need(mf, 4))
    if n == "GTR4" and P and D:
        return gtr4(nz, mps, need(mf, 4))
    if n == "Poisson" and not P and not D:
        return poisson(nz)
    if n == "Poisson-Custom" and not P and D:
        return poisson_custom(None, nz, need(mf, 20))
    if n == "LG" and not P and not D:
        return lg(nz)
    if n == "LG-Custom" and not P and D:
        return lg_custom(None, nz, need(mf, 20))
    if n == "WAG" and not P and not D:
        return wag(nz)
    if n == "WAG-Custom" and not P and D:
        return wag_custom(None, nz, need(mf, 20))
    if n == "GTR20" and P and D:
        return gtr20(nz, mps, need(mf, 20))
    raise ValueError("Cannot assemble substitution model. Name: %r Parameters: %r Stationary distribution: %r" % (n, mps, mf))


def _parse_sm(nz, s, i):
    """PhyloModel.hs:73-92,145-153; a literal {pi} must sum to 1 within 1e-12 (:86)."""
    n, i = _parse_name(s, i)
    mps, i = _parse_list(s, i, "[", "]")
    mf, i = _parse_list(s, i, "{", "}")
    if mf is not None and not abs(sum(abs(x) for x in mf) - 1.0) < 1e-12:
        raise ValueError("Sum of stationary distribution is %r but should be 1.0." % sum(mf))
    return assemble_substitution_model(nz, n, mps, mf), i


def get_phylo_model(ms: Optional[str], mm: Optional[str], global_norm: bool = False,
                    mws: Optional[Sequence[float]] = None, edm_cs=None):
    """PhyloModel.hs:211-231 (+ :155-202). Returns SubstitutionModel or MixtureModel."""
    if ms is None and mm is None:
        raise ValueError("No model was given. See help.")
    if ms is not None and mm is not None:
        raise ValueError("Both, substitution and mixture model string given; use only one.")
    if ms is not None:
        if mws is not None:
            raise ValueError("Weights given; but cannot be used with substitution model.")
        if edm_cs is not None:
            raise ValueError("Empirical distribution mixture model components given; but cannot be used with substitution model.")
        if global_norm:
            raise ValueError("Global normalization not possible for substitution models.")
        sm, i = _parse_sm(DO_NORMALIZE, ms, 0)
        if i != len(ms):
            raise ValueError("trailing input in model string %r" % ms)
        return sm
    sub_nz, mm_nz = (DO_NOT_NORMALIZE, DO_NORMALIZE) if global_norm else (DO_NORMALIZE, DO_NOT_NORMALIZE)
    if edm_cs is not None:  # :155-177
        m = re.fullmatch(r"EDM\((.*)\)", mm)
        if not m:
            raise ValueError("edmModel: cannot parse %r" % mm)
        inner = m.group(1)
        n, i = _parse_name(inner, 0)
        mps, i = _parse_list(inner, i, "[", "]")
        if len(edm_cs) == 0:
            raise ValueError("edmModel: no EDM components given.")
        sms = [assemble_substitution_model(sub_nz, n, mps, c[1]) for c in edm_cs]
        ws = [c[0] for c in edm_cs] if mws is None else list(mws)
        if len(sms) != len(ws):
            raise ValueError("edmModel: number of substitution models and weights differs.")
        return mm_from_substitution_models("EDM%d" % len(edm_cs), mm_nz, ws, sms)
    m = re.fullmatch(r"C(\d+)", mm)
    if m:  # :179-184
        if not global_norm:
            raise ValueError("Local normalization impossible with CXX models.")
        return cxx(int(m.group(1)), mws)
    if mws is None:
        raise ValueError("No weights provided.")
    m = re.fullmatch(r"MIXTURE\((.*)\)", mm)  # :186-197
    if not m:
        raise ValueError("mixtureModel: cannot parse %r" % mm)
    inner = m.group(1)
    sms, i = [], 0
    while True:
        sm, i = _parse_sm(sub_nz, inner, i)
        sms.append(sm)
        if i < len(inner) and inner[i] == ",":
            i += 1
            continue
        break
    if i != len(inner):
        raise ValueError("trailing input in mixture string %r" % mm)
    return mm_from_substitution_models("MIXTURE", mm_nz, mws, sms)


def flatten_model(pm):
    """What simulateAlignment hands the simulator (slynx Simulate.hs:100-116):
    (is_mixture, weights[C], pi[C][K], exch[C][K][K])."""
    if isinstance(pm, SubstitutionModel):
        return False, np.ones(1), pm.stationary_distribution[None, :].copy(), pm.exchangeability_matrix[None].copy()
    ws = np.array(pm.weights, dtype=np.float64)
    ds = np.stack([sm.stationary_distribution for sm in pm.models])
    es = np.stack([sm.exchangeability_matrix for sm in pm.models])
    return True, ws, ds, es


# --------------------------------------------------------------------------------------
# elynx-tree: Newick (Standard) -> preorder arrays
# --------------------------------------------------------------------------------------
@dataclass
class FlatTree:
    """Preorder flattening of `Data.Tree.Tree Double` (root first, children left to right)."""

    parent: np.ndarray  # int32 [N], -1 for the root
    brlen: np.ndarray  # float64 [N]; root stem 0 if absent (Phylogeny.hs:374-381)
    leaf_row: np.ndarray  # int32 [N]; DFS left-to-right leaf index (Rooted.hs:244-248) or -1
    names: List[str]  # [N]

    @property
    def n_nodes(self):
        return len(self.parent)

    @property
    def n_leaves(self):
        return int((self.leaf_row >= 0).sum())

    def leaf_names(self):
        return [self.names[i] for i in range(self.n_nodes) if self.leaf_row[i] >= 0]


_NAME_BAD = " :;()[],"


def parse_newick(text: str) -> FlatTree:
    """elynx-tree/src/ELynx/Tree/Import/Newick.hs:90-201 (Standard flavour: `name:length[support]`),
    followed by toLengthTree (Phylogeny.hs:374-381): a missing root length becomes 0, any other
    missing length is an error; negative lengths are rejected (Length.hs `toLength`)."""
    s = text
    pos = 0
    parent, brlen, leaf_row, names = [], [], [], []
    n_leaf = 0

    def skip_ws():
        nonlocal pos
        while pos < len(s) and s[pos].isspace():
            pos += 1

    def parse_name():
        nonlocal pos
        skip_ws()
        if pos < len(s) and s[pos] == "'":
            j = s.index("'", pos + 1)
            nm = s[pos + 1:j]
            pos = j + 1
            return nm
        j = pos
        while j < len(s) and s[j] not in _NAME_BAD:
            j += 1
        nm = s[pos:j]
        pos = j
        return nm

    num = re.compile(r"[-+]?(\d+\.?\d*|\.\d+)([eE][-+]?\d+)?")

    def parse_phylo():
        nonlocal pos
        save = pos
        length = None
        skip_ws()
        if pos < len(s) and s[pos] == ":":
            pos += 1
            skip_ws()
            m = num.match(s, pos)
            if not m:
                raise ValueError("newick: branchLengthSimple; double at %d" % pos)
            length = float(m.group(0))
            if length < 0:
                raise ValueError("newick: negative branch length")
            pos = m.end()
        else:
            pos = save
        if pos < len(s) and s[pos] == "[":
            j = s.index("]", pos)
            float(s[pos + 1:j])
            pos = j + 1
        return length

    def parse_tree(par):
        nonlocal pos, n_leaf
        me = len(parent)
        parent.append(par)
        brlen.append(None)
        leaf_row.append(-1)
        names.append("")
        if pos < len(s) and s[pos] == "(":
            pos += 1
            while True:
                parse_tree(me)
                if pos < len(s) and s[pos] == ",":
                    pos += 1
                    skip_ws()
                    continue
                break
            if pos >= len(s) or s[pos] != ")":
                raise ValueError("newick: expected ')' at %d" % pos)
            pos += 1
        else:
            leaf_row[me] = n_leaf
            n_leaf += 1
        names[me] = parse_name()
        brlen[me] = parse_phylo()

    skip_ws()
    parse_tree(-1)
    if pos >= len(s) or s[pos] != ";":
        raise ValueError("newick: expected ';' at %d" % pos)
    if brlen[0] is None:
        brlen[0] = 0.0
    if any(b is None for b in brlen):
        raise ValueError("toLengthTree: Length unavailable for some branches.")
    return FlatTree(np.array(parent, dtype=np.int32), np.array(brlen, dtype=np.float64),
                    np.array(leaf_row, dtype=np.int32), names)


# --------------------------------------------------------------------------------------
# Simulate/MarkovProcess.hs -- probMatrix (hmatrix expm) and categorical (mwc-random)
# --------------------------------------------------------------------------------------
def _golub_q() -> int:
    """hmatrix Internal/Algorithms.hs `geps eps`: first k with 2^(3-2k) k!^2 / ((2k)! (2k+1)!) < eps."""
    eps = np.finfo(np.float64).eps
    k = 1
    while True:
        g = 2.0 ** (3 - 2 * k) * math.factorial(k) ** 2 / (math.factorial(2 * k) * math.factorial(2 * k + 1))
        if g < eps:
            return k
        k += 1


_Q_PADE = _golub_q()


def expm_hmatrix(m):
    """hmatrix `expm` = expGolub (Golub & Van Loan Alg. 11.3.1): j = max 0 (floor (log2 ||m||_inf)),
    a = m / 2^j, diagonal Pade of order q (= 7 for doubles), f = D^-1 N, square j times.
    Works on a stack of matrices [..., K, K]."""
    m = np.asarray(m, dtype=np.float64)
    nrm = np.max(np.sum(np.abs(m), axis=-1), axis=-1)
    with np.errstate(divide="ignore"):
        j = np.maximum(0, np.floor(np.log2(nrm))).astype(np.int64)
    j = np.where(nrm > 0, j, 0)
    a = m / (2.0 ** j)[..., None, None]
    q = _Q_PADE
    eye = np.broadcast_to(np.eye(m.shape[-1]), m.shape).copy()
    c = 1.0
    x = eye.copy()
    n = eye.copy()
    d = eye.copy()
    for k in range(1, q + 1):
        c = c * (q - k + 1) / ((2 * q - k + 1) * k)
        x = a @ x
        n = n + c * x
        d = d + ((-1) ** k * c) * x
    f = np.linalg.solve(d, n)
    jmax = int(j.max()) if j.size else 0
    for it in range(jmax):
        sq = f @ f
        f = np.where((j > it)[..., None, None], sq, f)
    return f


def prob_matrix(q, t: float):
    """Simulate/MarkovProcess.hs:33-41"""
    q = np.asarray(q, dtype=np.float64)
    if q.shape[0] != q.shape[1]:
        raise ValueError("probMatrix: Matrix is not square.")
    if t == 0:
        return np.eye(q.shape[0])
    if t < 0:
        raise ValueError("probMatrix: Time is negative.")
    return expm_hmatrix(t * q)


def prob_tables(ds, es, brlen):
    """toProbTree / toProbTreeMixtureModel (MarkovProcessAlongTree.hs:54-55,157-161):
    P[n][c] = probMatrix (fromExchangeabilityMatrix es[c] ds[c]) brlen[n].  Returns [N][C][K][K]."""
    ds = np.asarray(ds, dtype=np.float64)
    es = np.asarray(es, dtype=np.float64)
    brlen = np.asarray(brlen, dtype=np.float64)
    if np.any(brlen < 0):
        raise ValueError("probMatrix: Time is negative.")
    qs = np.stack([from_exchangeability_matrix(e, d) for e, d in zip(es, ds)])  # [C][K][K]
    tq = brlen[:, None, None, None] * qs[None]
    p = expm_hmatrix(tq)
    k = qs.shape[-1]
    p[brlen == 0] = np.eye(k)
    return p


def categorical(w, u: float) -> int:
    """mwc-random `categorical`: cv = scanl1 (+) w; p = last cv * u; first i with cv[i] >= p."""
    cv = np.cumsum(np.asarray(w, dtype=np.float64))
    p = cv[-1] * u
    idx = np.nonzero(cv >= p)[0]
    if len(idx) == 0:
        raise ValueError("categorical: bad weights!")
    return int(idx[0])


class SplitMix:
    """random's StdGen (splitmix SMGen) + uniformDoublePositive01M; see SURVEY Appendix A.2/A.3."""

    M64 = (1 << 64) - 1

    def __init__(self, seed: int, gamma: int):
        self.seed, self.gamma = seed & self.M64, gamma & self.M64

    @staticmethod
    def _mix64(z):
        M = SplitMix.M64
        z &= M
        z ^= z >> 33
        z = (z * 0xFF51AFD7ED558CCD) & M
        z ^= z >> 33
        z = (z * 0xC4CEB9FE1A85EC53) & M
        z ^= z >> 33
        return z

    @staticmethod
    def _mix64v13(z):
        M = SplitMix.M64
        z &= M
        z ^= z >> 30
        z = (z * 0xBF58476D1CE4E5B9) & M
        z ^= z >> 27
        z = (z * 0x94D049BB133111EB) & M
        z ^= z >> 31
        return z

    @staticmethod
    def _mix_gamma(z):
        z = SplitMix._mix64v13(z) | 1
        if bin(z ^ (z >> 1)).count("1") >= 24:
            return z
        return z ^ 0xAAAAAAAAAAAAAAAA

    @classmethod
    def mk_std_gen(cls, s: int):
        return cls(cls._mix64(s), cls._mix_gamma((s + 0x9E3779B97F4A7C15) & cls.M64))

    def next_word64(self) -> int:
        self.seed = (self.seed + self.gamma) & self.M64
        return self._mix64(self.seed)

    def split(self):
        """returns the new generator; self becomes the other half"""
        s1 = (self.seed + self.gamma) & self.M64
        s2 = (s1 + self.gamma) & self.M64
        other = SplitMix(self._mix64(s1), self._mix_gamma(s2))
        self.seed = s2
        return other

    def uniform_double_positive01(self) -> float:
        w = self.next_word64()
        return float(w) / 18446744073709551616.0 + 2.0 ** -65


# --------------------------------------------------------------------------------------
# Simulate/MarkovProcessAlongTree.hs -- small-case pure numpy walk (big cases: oracle/sim.c)
# --------------------------------------------------------------------------------------
def simulate_injected_py(is_mixture: bool, ws, ds, p, tree: FlatTree, u):
    """Literal restatement of simulateAndFlatten' / simulateAndFlattenMixtureModel'
    (MarkovProcessAlongTree.hs:57-63,88-98,163-173,195-208) consuming an injected uniform matrix
    u[(1|2)+N][S] in the order of SURVEY Appendix A.1.  p = [N][C][K][K] probability matrices.
    Returns (comps[S], leaf_states[T][S], all_states[N][S])."""
    u = np.asarray(u, dtype=np.float64)
    n_nodes = tree.n_nodes
    s = u.shape[1]
    assert u.shape[0] == n_nodes + (2 if is_mixture else 1)
    row = 0
    if is_mixture:
        comps = np.array([categorical(ws, u[0, i]) for i in range(s)], dtype=np.int32)
        row = 1
    else:
        comps = np.zeros(s, dtype=np.int32)
    root = np.array([categorical(ds[comps[i]], u[row, i]) for i in range(s)], dtype=np.int64)
    row += 1
    states = np.zeros((n_nodes, s), dtype=np.uint8)
    for n in range(n_nodes):
        par = root if tree.parent[n] < 0 else states[tree.parent[n]]
        for i in range(s):
            states[n, i] = categorical(p[n, comps[i], par[i]], u[row + n, i])
    leaves = np.zeros((tree.n_leaves, s), dtype=np.uint8)
    for n in range(n_nodes):
        if tree.leaf_row[n] >= 0:
            leaves[tree.leaf_row[n]] = states[n]
    return comps, leaves, states


def get_chunks(c: int, n: int) -> List[int]:
    """MarkovProcessAlongTree.hs:210-215"""
    q, r = divmod(n, c)
    return [q + 1] * r + [q] * (c - r)


def sequences_to_fasta(names: Sequence[str], leaves, alphabet: str) -> bytes:
    """elynx-seq/src/ELynx/Sequence/Export/Fasta.hs:24-36 with empty description;
    state -> char per slynx Simulate.hs:117-124."""
    lut = np.frombuffer(ALPHABET_CHARS[alphabet].encode(), dtype=np.uint8)
    out = []
    for nm, row in zip(names, leaves):
        out.append(b">" + nm.encode() + b"\n" + lut[np.asarray(row)].tobytes() + b"\n")
    return b"".join(out)


def site_dists_lines(comps, ds) -> bytes:
    """slynx Simulate.hs:79-89: per site `i pi_paml...` (1-based), joined by newlines, PAML order."""
    def dbl(x):
        return _double_dec(float(x))
    lines = []
    for i, c in enumerate(comps):
        lines.append(str(i + 1) + " " + " ".join(dbl(x) for x in alpha_to_paml_vec(ds[c])))
    return "\n".join(lines).encode()


def _double_dec(x: float) -> str:
    """ByteString.Builder.doubleDec = Haskell `show` for Double: shortest round-trip digits,
    scientific outside [0.1, 10^7)."""
    if x == 0:
        return "0.0"
    r = repr(abs(x))
    sign = "-" if x < 0 else ""
    if "e" in r or "E" in r:
        mant, ex = r.lower().split("e")
        ex = int(ex)
    else:
        mant, ex = r, 0
    if "." in mant:
        ip, fp = mant.split(".")
    else:
        ip, fp = mant, ""
    digits = (ip + fp).lstrip("0")
    # decimal exponent of the first significant digit
    lead = len(ip.lstrip("0"))
    if lead == 0:
        stripped = fp.lstrip("0")
        e10 = -(len(fp) - len(stripped)) - 1 + ex
    else:
        e10 = lead - 1 + ex
    digits = digits.rstrip("0") or "0"
    if 0.1 <= abs(x) < 1e7:
        if e10 >= 0:
            ip2 = digits[: e10 + 1].ljust(e10 + 1, "0")
            fp2 = digits[e10 + 1:] or "0"
        else:
            ip2 = "0"
            fp2 = "0" * (-e10 - 1) + digits
        return sign + ip2 + "." + fp2
    return sign + digits[0] + "." + (digits[1:] or "0") + "e" + str(e10)
