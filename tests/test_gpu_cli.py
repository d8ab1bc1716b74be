"""`slynx simulate` driver end to end on the GPU: the full output set of the reference (Simulate.hs:100-151, ELynx.hs:51-75),
--force semantics, the streamed .fasta.gz writer and one handle driving several GPUs."""
import gzip
import json
import os

import numpy as np
import pytest

import oracle.model as om
from helpers import model_case, tree_from_file

pytestmark = pytest.mark.gpu

import elynx_b200 as eb  # noqa: E402
from elynx_b200 import cli  # noqa: E402


def test_cli_full_output_set(tmp_path, ref_data):
    tree = os.path.join(ref_data, "Newick.tree")
    edm = os.path.join(ref_data, "EDMDistsPhylobayes.txt")
    base = str(tmp_path / "run")
    argv = ["-o", base, "simulate", "-t", tree, "-e", edm, "-m", "EDM(LG-Custom)", "-g", "(4,0.5)", "-l", "500", "-S", "[3]"]
    assert cli.main(argv) == 0
    for sfx in (".fasta", ".sitedists", ".model.gz", ".log", ".elynx"):
        assert os.path.exists(base + sfx), sfx
    lines = open(base + ".fasta", "rb").read().split(b"\n")
    assert len(lines) == 81 and all(len(l) == 500 for l in lines[1::2][:40])
    log = open(base + ".log").read()
    for needle in ("Command name: simulate", "Seed: fixed to 3.", "Read tree from file", "Number of leaves: 40",
                   "Empiricial distribution mixture model with 4 components.", "Simulate alignment.", "Length: 500.",
                   "Write simulated multi sequence alignment to file '%s.fasta'." % base, "Write ELynx reproduction file.",
                   "Wall clock run time:"):
        assert needle in log, needle
    model_text = gzip.open(base + ".model.gz", "rb").read().decode()
    assert model_text.startswith("MixtureModel (MixtureModel {name = ") and model_text.count("Component {weight = ") == 16
    rec = json.load(open(base + ".elynx"))
    assert rec["files"] == [tree, edm, base + ".model.gz", base + ".fasta"] and len(rec["checkSums"]) == 4
    assert rec["reproducible"]["local"]["argsSeed"] == {"tag": "Fixed", "contents": 3}
    # a second run refuses to overwrite (InputOutput.hs:52-65) unless -f is given; the alignment is reproducible
    first = open(base + ".fasta", "rb").read()
    with pytest.raises(SystemExit) as e:
        cli.main(argv)
    assert "File exists" in str(e.value)
    assert cli.main(["-f"] + argv) == 0
    assert open(base + ".fasta", "rb").read() == first
    # --no-elynx-file: neither the reproduction file nor the machine-readable model
    base2 = str(tmp_path / "run2")
    assert cli.main(["--no-elynx-file", "-v", "Quiet", "-o", base2] + argv[2:]) == 0
    assert os.path.exists(base2 + ".fasta") and not os.path.exists(base2 + ".elynx") and not os.path.exists(base2 + ".model.gz")
    assert not os.path.exists(base2 + ".log")
    assert open(base2 + ".fasta", "rb").read() == first


@pytest.mark.parametrize("name,s", [("hky_g4", 300_000), ("lg_g4", 150_000), ("jc", 7)])
def test_fasta_gz_stream_equals_fasta(tmp_path, name, s):
    """elx_simulate_fasta_gz: gunzip of the streamed multi-member file == the bytes of elx_simulate_fasta (several members,
    several row blocks; component indices alike)."""
    model = model_case(name)
    is_mix, ws, ds, es, alpha = model
    tree = tree_from_file("Newick.tree")
    names = tree.leaf_names()
    chars = om.ALPHABET_CHARS[alpha]
    path = str(tmp_path / "a.fasta.gz")
    with eb.Simulator(ws if is_mix else None, ds, es, tree, is_mixture=is_mix) as sim:
        ref, c0 = sim.simulate_fasta(5, s, names, chars, want_comps=True)
        c1 = sim.simulate_fasta_gz(5, s, names, chars, path, level=1, threads=4, want_comps=True)
    raw = open(path, "rb").read()
    assert gzip.decompress(raw) == ref
    assert np.array_equal(c0, c1)
    n_members = raw.count(b"\x1f\x8b\x08\x00\x00\x00\x00\x00")
    assert n_members >= max(1, len(ref) // (4 << 20))


def test_cli_fasta_gz_and_site_components(tmp_path, ref_data):
    tree = os.path.join(ref_data, "Newick.tree")
    base = str(tmp_path / "g")
    assert cli.main(["-v", "Quiet", "-o", base, "simulate", "-t", tree, "-m", "C10", "-n", "-l", "2000", "-S", "[1]", "--fasta-gz"]) == 0
    assert cli.main(["-v", "Quiet", "-o", base + "p", "simulate", "-t", tree, "-m", "C10", "-n", "-l", "2000", "-S", "[1]"]) == 0
    assert gzip.open(base + ".fasta.gz", "rb").read() == open(base + "p.fasta", "rb").read()
    assert open(base + ".sitedists", "rb").read() == open(base + "p.sitedists", "rb").read()
    # elx_site_components: the `cs` of simulateAndFlattenMixtureModelPar alone
    model = model_case("c10")
    is_mix, ws, ds, es, _ = model
    t = tree_from_file("Newick.tree")
    with eb.Simulator(ws, ds, es, t, is_mixture=True) as sim:
        _, comps, _ = sim.simulate(1, 3000, site0=1234567, want_comps=True)
        assert np.array_equal(sim.site_components(1, 3000, site0=1234567), comps)


@pytest.mark.parametrize("name,site0,s", [("hky_g4", 0, (1 << 22) + 333), ("hky_g4", 96, 5000), ("lg_g4", 64, (1 << 20) + 4097), ("c10", 5, 3001)])
def test_simulate_packed_is_the_alignment(name, site0, s):
    """elx_simulate_packed: the 2- / 5-bit rows expand (elx_unpack2_rows / elx_unpack5_rows) to exactly what elx_simulate returns,
    several batches, ragged ends, unaligned site0; d2h bytes are the packed size."""
    import ctypes as C

    from elynx_b200 import _lib as L

    model = model_case(name)
    is_mix, ws, ds, es, _ = model
    tree = tree_from_file("UltraMetric.tree")
    with eb.Simulator(ws if is_mix else None, ds, es, tree, is_mixture=is_mix) as sim:
        ref, c0, _ = sim.simulate(4, s, site0=site0, want_comps=True)
        b0 = sim.d2h_bytes
        packed, bits, c1 = sim.simulate_packed(4, s, site0=site0, want_comps=True)
        moved = sim.d2h_bytes - b0
    assert bits == (2 if ds.shape[1] <= 4 else 5)
    assert packed.shape == (tree.n_leaves, (s + 31) // 32 * (8 if bits == 2 else 20))
    out = np.empty((tree.n_leaves, s), dtype=np.uint8)
    fn = L.lib().elx_unpack2_rows if bits == 2 else L.lib().elx_unpack5_rows
    L.check(fn(C.c_void_p(packed.ctypes.data), packed.shape[1], C.c_void_p(out.ctypes.data), s, tree.n_leaves, s, None, 2))
    assert np.array_equal(out, ref) and np.array_equal(c0, c1)
    assert moved <= packed.size + 4 * s + (1 << 12)
