"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy / pure Python) of the alignment statistics behind `slynx examine`.

Follows, line by line where floating point order matters:
  elynx-seq/src/ELynx/Alphabet/DistributionDiversity.hs:37-110  (xLogX, entropy, kEffEntropy, homoplasy, kEffHomoplasy,
                                                                  countCharacters, saveDivision, frequencyCharacters)
  elynx-seq/src/ELynx/Sequence/Alignment.hs:209-300             (filterColsConstant, filterColsConstantSoft,
                                                                  filterColsOnlyStd, filterColsNoGaps, toFrequencyData,
                                                                  distribution, kEffEntropy, kEffHomoplasy, countGaps)
  elynx-seq/src/ELynx/Sequence/Distance.hs:24-33                (hamming)
  elynx-seq/src/ELynx/Sequence/Divergence.hs:38-64              (divergence)
  slynx/src/SLynx/Examine/Examine.hs:52-160                     (examineAlignment: which statistic over which columns)

An alignment is a [T][S] uint8 matrix of codes: 0..K-1 = the standard characters in sorted order, K = '-', K+1 = '.'
(the two gap characters of the DNAX / ProteinX alphabets, Alphabet.hs:158-160, 229-231; `toStd` of a gap is [] so gaps
never enter a frequency).  IUPAC alphabets (DNAI, ProteinS, ProteinI) are out of scope.
Pinned by the reference's own known answers: elynx-seq/test/ELynx/Alphabet/DistributionDiversitySpec.hs:21-57."""
from __future__ import annotations

import math

import numpy as np

EPS = 1e-12  # DistributionDiversity.hs:37-38


def x_log_x(x: float) -> float:
    if x < 0.0:
        raise ValueError("Argument lower than zero.")
    if EPS > x:
        return 0.0
    return x * math.log(x)


def entropy(v) -> float:
    return -sum(x_log_x(float(x)) for x in v)  # V.sum (V.map xLogX v): left to right


def k_eff_entropy(v) -> float:
    e = entropy(v)
    return 1.0 if e < EPS else math.exp(e)


def homoplasy(v) -> float:
    return sum(float(x) * float(x) for x in v)


def k_eff_homoplasy(v) -> float:
    h = homoplasy(v)
    return math.inf if h == 0.0 else 1.0 / h


def frequency_characters(col, k: int):
    """frequencyCharacters: counts of the standard characters / their sum (0 when the column has none)."""
    counts = [int((col == i).sum()) for i in range(k)]
    s = sum(counts)
    return [0.0 if s == 0 else c / s for c in counts]


def examine(a: np.ndarray, k: int, divergence: bool = False) -> dict:
    """Statistics of examineAlignment (Examine.hs:52-160) for a code matrix a[T][S]."""
    a = np.asarray(a, dtype=np.uint8)
    t, s = a.shape
    is_gap = a >= k
    cols_const = [bool(np.all(a[:, j] == a[0, j])) for j in range(s)]
    cols_nogap = [not bool(is_gap[:, j].any()) for j in range(s)]

    def const_soft(col, gap):
        std = col[~gap]
        if std.size == 0:
            return False
        return bool(np.all(std == std[0]))

    cols_soft = [const_soft(a[:, j], is_gap[:, j]) for j in range(s)]
    freqs = [frequency_characters(a[:, j], k) for j in range(s)]
    dist = [0.0] * k
    if s:
        acc = list(freqs[0])
        for f in freqs[1:]:  # foldl1 (V.zipWith (+)) over the columns
            acc = [x + y for x, y in zip(acc, f)]
        dist = [x / s for x in acc]

    def mean(xs):
        return sum(xs) / len(xs) if xs else math.nan

    ke = [k_eff_entropy(f) for f in freqs]
    kh = [k_eff_homoplasy(f) for f in freqs]
    ke_ng = [x for x, g in zip(ke, cols_nogap) if g]
    kh_ng = [x for x, g in zip(kh, cols_nogap) if g]
    n_gaps = int(is_gap.sum())
    out = {
        "n_sites": s, "n_sequences": t,
        "n_constant": sum(cols_const), "n_constant_soft": sum(cols_soft), "n_no_gaps": sum(cols_nogap),
        "n_only_std": sum(cols_nogap), "n_std_chars": t * s - n_gaps, "n_gaps": n_gaps,
        "distribution": dist, "keff_entropy_per_site": ke,
        "keff_entropy_mean": mean(ke), "keff_entropy_mean_no_gaps": mean(ke_ng),
        "keff_homoplasy_mean": mean(kh), "keff_homoplasy_mean_no_gaps": mean(kh_ng),
    }
    # pairwise Hamming distances per site over all ordered pairs x /= y (Examine.hs:121-128); Statistics.Sample.meanVariance
    # is the maximum-likelihood variance (divide by n)
    ham = np.zeros((t, t), dtype=np.int64)
    for i in range(t):
        ham[i] = (a != a[i]).sum(axis=1)
    out["hamming"] = ham
    pair = [ham[i, j] / s for i in range(t) for j in range(t) if i != j] if s else []
    if pair:
        m = sum(pair) / len(pair)
        out.update(hamming_mean=m, hamming_var=sum((x - m) ** 2 for x in pair) / len(pair), hamming_min=min(pair), hamming_max=max(pair))
    else:
        out.update(hamming_mean=math.nan, hamming_var=math.nan, hamming_min=math.nan, hamming_max=math.nan)
    if divergence:
        div = {}
        for i in range(t):
            for j in range(i + 1, t):
                both = (a[i] < k) & (a[j] < k)  # only sites where both characters are standard (Divergence.hs:50-56)
                m = np.zeros((k, k), dtype=np.int64)
                np.add.at(m, (a[i][both], a[j][both]), 1)
                div[(i, j)] = m
        out["divergence"] = div
    return out
