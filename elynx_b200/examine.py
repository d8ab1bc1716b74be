"""Host-side mirror of `slynx examine` (slynx/src/SLynx/Examine/Examine.hs) on top of elx_examine[_device].

An alignment is a [T][S] uint8 code matrix: 0..K-1 = the standard characters in sorted order, K = '-', K+1 = '.'
(alphabets DNA / DNAX / Protein / ProteinX of elynx-seq Alphabet.hs:145-232; IUPAC alphabets are out of scope).
`examine_alignment` returns the statistics, `format_report` lays them out like examineAlignment (Examine.hs:52-113)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Sequence

import numpy as np

from . import _lib as L

ALPHABETS = {"DNA": "ACGT", "DNAX": "ACGT", "Protein": "ACDEFGHIKLMNPQRSTVWY", "ProteinX": "ACDEFGHIKLMNPQRSTVWY"}
GAPS = "-."


def encode(seqs: Sequence[bytes], alphabet: str) -> np.ndarray:
    """FASTA characters -> codes; raises like the reference's parser for characters outside the alphabet."""
    std = ALPHABETS[alphabet]
    lut = np.full(256, 255, dtype=np.uint8)
    for i, ch in enumerate(std):
        lut[ord(ch)] = i
    if alphabet.endswith("X"):
        for i, ch in enumerate(GAPS):
            lut[ord(ch)] = len(std) + i
    if len({len(s) for s in seqs}) > 1:
        raise L.ElxError(L.ELX_E_INVALID, "Sequences do not have equal lengths.")
    a = np.stack([lut[np.frombuffer(s, dtype=np.uint8)] for s in seqs]) if seqs else np.zeros((0, 0), dtype=np.uint8)
    if (a == 255).any():
        raise L.ElxError(L.ELX_E_INVALID, "alignment holds characters outside the %s alphabet" % alphabet)
    return a


def _result(r: L.elx_examine_result, k: int) -> Dict:
    d = {f: getattr(r, f) for f, _ in r._fields_ if f not in ("distribution", "reserved")}
    d["distribution"] = [r.distribution[i] for i in range(k)]
    return d


def examine_alignment(states: np.ndarray, n_states: int, per_site: bool = False, hamming_matrix: bool = False,
                      divergence: bool = False) -> Dict:
    """Statistics of a host code matrix [T][S] (copied to the current CUDA device and examined there)."""
    a = np.ascontiguousarray(states, dtype=np.uint8)
    if a.ndim != 2:
        raise L.ElxError(L.ELX_E_INVALID, "examine: expected a [sequences][sites] matrix")
    t, s = a.shape
    res = L.elx_examine_result()
    ke = np.empty(s, dtype=np.float64) if per_site else None
    eq = np.zeros((t, t), dtype=np.int64) if hamming_matrix else None
    div = np.zeros((t * (t - 1) // 2, n_states, n_states), dtype=np.int64) if divergence else None
    ptr = lambda x: None if x is None else C.c_void_p(x.ctypes.data)  # noqa: E731
    L.check(L.lib().elx_examine(ptr(a) if a.size else None, max(s, 1) if a.size == 0 else s, t, s, n_states, C.byref(res), ptr(ke), ptr(eq), ptr(div)))
    out = _result(res, n_states)
    if per_site:
        out["keff_entropy_per_site"] = ke
    if hamming_matrix:
        ham = s - eq
        ham = np.triu(ham, 1)
        out["hamming"] = ham + ham.T
    if divergence:
        out["divergence"] = div
    return out


def examine_device(d_states: int, pitch: int, n_sequences: int, n_sites: int, n_states: int, stream: int = 0,
                   d_keff_per_site: int = 0, d_equal_counts: int = 0, d_divergence: int = 0) -> Dict:
    """The same on a device-resident matrix (e.g. the buffer elx_simulate_device just filled); raw device addresses."""
    res = L.elx_examine_result()
    L.check(L.lib().elx_examine_device(C.c_void_p(d_states), pitch, n_sequences, n_sites, n_states, C.byref(res),
                                       C.c_void_p(d_keff_per_site or None), C.c_void_p(d_equal_counts or None),
                                       C.c_void_p(d_divergence or None), C.c_void_p(stream or None)))
    return _result(res, n_states)


def _p_row(name: str, val: str) -> str:
    """pRow (Examine.hs:46-50): alignLeft 50 name <> alignRight 10 value (both trim when longer)."""
    return name[:50].ljust(50) + val[:10].rjust(10)


def format_report(r: Dict, alphabet: str) -> str:
    """examineAlignment (Examine.hs:52-113), line for line (the optional per-site block is appended by the CLI)."""
    std = ALPHABETS[alphabet]
    n_tot = r["n_sites"] * r["n_sequences"]
    n_gaps, n_iupac, n_unknown = r["n_gaps"], 0, 0
    pct = lambda x: 100.0 * x / n_tot if n_tot else float("nan")  # noqa: E731
    f3 = lambda x: "%.3f" % x  # noqa: E731
    lines = [
        "Sequences have equal length (multi sequence alignment, or single sequence).",
        "Number of columns in alignment:",
        _p_row("  Total:", str(r["n_sites"])),
        _p_row("  Constant:", str(r["n_constant"])),
        _p_row("  Constant (including gaps or unknowns):", str(r["n_constant_soft"])),
        _p_row("  Without gaps:", str(r["n_no_gaps"])),
        _p_row("  With standard characters only:", str(r["n_only_std"])),
        "",
        _p_row("Total number of characters:", str(n_tot)),
        _p_row("Standard (i.e., not extended IUPAC) characters:", str(n_tot - n_iupac - n_gaps - n_unknown)),
        _p_row("Extended IUPAC characters:", str(n_iupac)),
        _p_row("Gaps:", str(n_gaps)),
        _p_row("Unknowns:", str(n_unknown)),
        "",
        _p_row("Percentage of standard characters:", "%2.2f" % (100.0 - pct(n_iupac) - pct(n_gaps) - pct(n_unknown))),
        _p_row("Percentage of extended IUPAC characters:", "%2.2f" % pct(n_iupac)),
        _p_row("Percentage of gaps:", "%2.2f" % pct(n_gaps)),
        _p_row("Percentage of unknowns:", "%2.2f" % pct(n_unknown)),
        "",
        "Distribution of characters:",
        "".join(ch + "     " for ch in std),
        " ".join(f3(x) for x in r["distribution"]),
        "",
        "Pairwise hamming distances (per site):",
        _p_row("  Mean:", f3(r["hamming_mean"])),
        _p_row("  Standard deviation:", f3(r["hamming_var"] ** 0.5)),
        _p_row("  Minimum:", f3(r["hamming_min"])),
        _p_row("  Maximum:", f3(r["hamming_max"])),
        "",
        "Mean effective number of states (measured using entropy):",
        _p_row("Across whole alignment:", f3(r["keff_entropy_mean"])),
        _p_row("Across columns without gaps:", f3(r["keff_entropy_mean_no_gaps"])),
        _p_row("Across columns without extended IUPAC characters:", f3(r["keff_entropy_mean_no_gaps"])),
        "",
        "Mean effective number of states (measured using homoplasy):",
        _p_row("Across whole alignment:", f3(r["keff_homoplasy_mean"])),
        _p_row("Across columns without gaps:", f3(r["keff_homoplasy_mean_no_gaps"])),
        _p_row("Across columns without extended IUPAC characters:", f3(r["keff_homoplasy_mean_no_gaps"])),
        "",
        "Divergence matrix:",
    ]
    return "\n".join(lines) + "\n"
