"""`slynx examine` core: the CPU restatement against the reference's known answers, and the device
kernels against the restatement through the C ABI."""
import ctypes as C
import math

import numpy as np
import pytest

import oracle.examine as ox

import elynx_b200 as eb
from elynx_b200 import _lib as L
from elynx_b200 import examine as ex


# ---------------------------------------------------------------------------------------------
# oracle vs the reference's own known answers (elynx-seq/test/ELynx/Alphabet/DistributionDiversitySpec.hs:21-57)
# ---------------------------------------------------------------------------------------------
def test_oracle_distribution_diversity_kats():
    a1, a2, a3 = [0.0] * 20, [0, 0, 0, 1, 0], [0.3, 0.4, 0.7]
    assert ox.entropy(a1) == 0.0 and ox.entropy(a2) == 0.0
    assert abs(ox.entropy(a3) - 0.9773805948045555) < 1e-12          # nearlyEq = 1e-12 (ELynx.Tools.Equality)
    assert ox.k_eff_entropy(a1) == 1.0 and ox.k_eff_entropy(a2) == 1.0
    assert abs(ox.k_eff_entropy(a3) - 2.6574860842252765) < 1e-12
    assert ox.homoplasy(a1) == 0.0 and ox.homoplasy(a2) == 1.0 and abs(ox.homoplasy(a3) - 0.74) < 1e-12
    assert math.isinf(ox.k_eff_homoplasy(a1)) and ox.k_eff_homoplasy(a2) == 1.0
    assert abs(ox.k_eff_homoplasy(a3) - 1.3513513513513513) < 1e-12


def test_oracle_examine_hand_checked():
    """AlignmentSpec.hs:28-36's little alignment (AAA / GAA / TAA) plus a gap row, counted by hand."""
    a = ex.encode([b"AAA", b"GAA", b"TAA", b"-A."], "DNAX")
    r = ox.examine(a, 4, divergence=True)
    assert (r["n_sites"], r["n_sequences"]) == (3, 4)
    assert r["n_constant"] == 1 and r["n_constant_soft"] == 2 and r["n_no_gaps"] == 1 and r["n_gaps"] == 2
    # column 0: A,G,T,- -> f = (1/3, 0, 1/3, 1/3); columns 1, 2: all A
    assert np.allclose(r["distribution"], [(1 / 3 + 2) / 3, 0, 1 / 9, 1 / 9])
    assert abs(r["keff_entropy_per_site"][0] - 3.0) < 1e-12 and r["keff_entropy_per_site"][1] == 1.0
    assert r["hamming"][0, 1] == 1 and r["hamming"][0, 3] == 2 and r["hamming"][1, 2] == 1
    assert r["divergence"][(0, 1)][0, 2] == 1 and r["divergence"][(0, 1)][0, 0] == 2 and r["divergence"][(0, 3)].sum() == 1


def test_report_layout():
    a = ex.encode([b"AAA", b"GAA", b"TAA"], "DNA")
    r = ox.examine(a, 4)
    text = ex.format_report(r, "DNA")
    lines = text.split("\n")
    assert lines[2] == "  Total:" + " " * 42 + " " * 9 + "3"                       # pRow: alignLeft 50 <> alignRight 10
    assert lines[3].startswith("  Constant:") and lines[3].endswith("2") and len(lines[3]) == 60
    assert lines[20] == "A     C     G     T     " and lines[21] == "0.778 0.000 0.111 0.111"
    assert lines[-2] == "Divergence matrix:"


def test_examine_has_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    a = np.zeros((2, 8), dtype=np.uint8)
    with pytest.raises(eb.ElxError, match="no CUDA device"):
        ex.examine_alignment(a, 4)


# ---------------------------------------------------------------------------------------------
# device kernels vs the restatement
# ---------------------------------------------------------------------------------------------
def _close(x, y, tol=1e-12):
    if isinstance(x, float) and (math.isnan(x) or math.isnan(y)):
        return math.isnan(x) and math.isnan(y)
    if isinstance(x, float) and (math.isinf(x) or math.isinf(y)):
        return x == y
    return abs(x - y) <= tol * max(1.0, abs(x), abs(y))


def _check(a, k):
    r0 = ox.examine(a, k, divergence=True)
    r = ex.examine_alignment(a, k, per_site=True, hamming_matrix=True, divergence=True)
    for key in ("n_sites", "n_sequences", "n_constant", "n_constant_soft", "n_no_gaps", "n_only_std", "n_std_chars", "n_gaps"):
        assert r[key] == r0[key], key
    for key in ("keff_entropy_mean", "keff_entropy_mean_no_gaps", "keff_homoplasy_mean", "keff_homoplasy_mean_no_gaps",
                "hamming_mean", "hamming_var", "hamming_min", "hamming_max"):
        assert _close(r[key], r0[key]), (key, r[key], r0[key])
    assert np.allclose(r["distribution"], r0["distribution"], rtol=0, atol=1e-13)
    assert np.allclose(r["keff_entropy_per_site"], r0["keff_entropy_per_site"], rtol=1e-13, atol=0)
    assert np.array_equal(r["hamming"], r0["hamming"])
    t = a.shape[0]
    p = 0
    for i in range(t):
        for j in range(i + 1, t):
            assert np.array_equal(r["divergence"][p], r0["divergence"][(i, j)]), (i, j)
            p += 1


@pytest.mark.gpu
@pytest.mark.parametrize("k,t,s,gap_frac", [(4, 1, 17, 0.0), (4, 2, 1, 0.0), (4, 5, 1023, 0.1), (4, 40, 4099, 0.0), (4, 70, 2500, 0.3),
                                            (20, 3, 64, 0.0), (20, 33, 3001, 0.05), (20, 65, 1024, 1.0), (2, 9, 777, 0.5)])
def test_examine_matches_restatement(k, t, s, gap_frac):
    rng = np.random.default_rng(k * 1000 + t)
    p = rng.dirichlet(np.ones(k) * 0.7)
    a = rng.choice(k, size=(t, s), p=p).astype(np.uint8)
    a[:, : s // 7] = a[0, : s // 7]                      # constant columns
    gaps = rng.random((t, s)) < gap_frac
    a[gaps] = rng.choice([k, k + 1], size=int(gaps.sum())).astype(np.uint8)
    _check(a, k)


@pytest.mark.gpu
def test_examine_edge_cases():
    assert ex.examine_alignment(np.zeros((3, 0), dtype=np.uint8), 4)["n_sites"] == 0
    bad = np.full((2, 10), 7, dtype=np.uint8)
    with pytest.raises(eb.ElxError, match="codes above"):
        ex.examine_alignment(bad, 4)
    with pytest.raises(eb.ElxError):
        ex.examine_alignment(np.zeros((2, 10), dtype=np.uint8), 1)


@pytest.mark.gpu
def test_examine_simulated_alignment_on_device():
    """simulate -> examine without leaving the device, at a size the CPU restatement cannot reach: the character
    distribution must be the model's stationary distribution, and partial results must add up (two halves of the
    site range: counts add, means combine)."""
    torch = pytest.importorskip("torch")
    from elynx_b200 import model as pm

    tree = pm.birth_death_tree(200, 1.0, 0.5, seed=4)
    model = pm.get_phylo_model("HKY[6.0]{0.2,0.3,0.3,0.2}", gamma=(4, 0.5))
    s, t = 2_000_000, 200
    with eb.Simulator(model.weights, model.pi, model.exch, tree, is_mixture=True) as sim:
        pitch = (s + 127) // 128 * 128
        buf = torch.empty((t, pitch), dtype=torch.uint8, device="cuda")
        sim.simulate_device(3, 0, s, buf.data_ptr(), pitch)
        sim.synchronize()
        whole = ex.examine_device(buf.data_ptr(), pitch, t, s, 4)
        h = s // 2 + 64
        left = ex.examine_device(buf.data_ptr(), pitch, t, h, 4)
        right = ex.examine_device(buf.data_ptr() + h, pitch, t, s - h, 4)
    assert whole["n_std_chars"] == t * s and whole["n_gaps"] == 0 and whole["n_no_gaps"] == s
    assert np.allclose(whole["distribution"], model.pi[0], atol=2e-3)
    assert whole["n_constant"] == left["n_constant"] + right["n_constant"]
    assert abs(whole["keff_entropy_mean"] * s - left["keff_entropy_mean"] * h - right["keff_entropy_mean"] * (s - h)) < 1e-6 * s
    assert abs(whole["hamming_mean"] * s - left["hamming_mean"] * h - right["hamming_mean"] * (s - h)) < 1e-9 * s
    assert 0.0 < whole["hamming_min"] <= whole["hamming_mean"] <= whole["hamming_max"] < 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("alphabet,fn", [("DNA", "Nucleotide.fasta"), ("Protein", "AminoAcid.fasta")])
def test_cli_examine_on_reference_fixtures(tmp_path, ref_data, alphabet, fn):
    """scripts/slynx-test:13-17 runs `slynx examine` on these two files: the driver must produce the report of
    examineAlignment with the numbers of the restatement, and the divergence file."""
    import os

    from elynx_b200 import cli

    path = os.path.join(ref_data, fn)
    base = str(tmp_path / "ex")
    assert cli.main(["-o", base, "examine", "-a", alphabet, path, "--per-site", "--divergence"]) == 0
    out = open(base + ".out").read()
    seqs = cli.read_fasta(open(path, "rb").read())
    a = ex.encode([c for _, _, c in seqs], alphabet)
    k = len(ex.ALPHABETS[alphabet])
    r0 = ox.examine(a, k, divergence=True)
    report = ex.format_report(r0, alphabet)
    assert report in out                                  # same text as the restatement's numbers (3-decimal fields)
    assert out.startswith("For each sequence, the 60 first bases are shown.\nList contains %d sequences.\n" % len(seqs))
    assert "Effective number of used states per site (measured using entropy):\n[" in out
    div = open(base + ".div").read().split("\n")
    assert div[0] == "> %s, %s" % (seqs[0][0].decode(), seqs[1][0].decode())
    first = np.array([[int(x) for x in row.split()] for row in div[2:2 + k]])
    assert np.array_equal(first, r0["divergence"][(0, 1)])
