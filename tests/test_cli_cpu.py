"""Host-side pieces of the `slynx simulate` driver (elynx_b200/cli.py): the ELynx logger layout, --force semantics, the
machine-readable model file and the reproduction file (elynx-tools Logger.hs, InputOutput.hs:52-65, Reproduction.hs:131-145)."""
import gzip
import hashlib
import io
import json
import os
import re
import types

import numpy as np
import pytest

from elynx_b200 import cli
from elynx_b200 import model as pm


def test_check_file_needs_force(tmp_path):
    p = str(tmp_path / "x.fasta")
    cli.check_file(p, False)
    open(p, "w").write("old")
    with pytest.raises(SystemExit) as e:
        cli.check_file(p, False)
    assert "File exists" in str(e.value) and "--force" in str(e.value)
    cli.check_file(p, True)


def test_runlog_layout(tmp_path):
    base = str(tmp_path / "run")
    buf = io.StringIO()
    log = cli.RunLog("Info", base, False, stream=buf)
    log.header("simulate", cli.SIMULATE_DSC, ["-o", base, "simulate", "-l", "10"])
    log.info("Seed: fixed to 0.")
    log.warn("careful")
    log.footer()
    log.close()
    text = open(base + ".log").read()
    assert text == buf.getvalue()
    lines = text.split("\n")
    assert lines[0] == "   " + "-" * 78
    assert "   Command name: simulate" in lines and "   Simulate multi sequence alignments." in lines
    assert any(re.fullmatch(r"   Starting time: [A-Z][a-z]+ \d+, \d{4}, at \d\d:\d\d [ap]m, UTC\.", l) for l in lines)
    assert any(l.startswith("   Command line: slynx-b200 -o ") for l in lines)
    assert "W: careful" in lines
    assert any(re.fullmatch(r"   Wall clock run time: \d\d:\d\d:\d\d\.", l) for l in lines)
    # an existing log is only replaced with --force; Quiet writes no log file at all
    with pytest.raises(SystemExit):
        cli.RunLog("Info", base, False, stream=buf)
    cli.RunLog("Info", base, True, stream=buf).close()
    q = cli.RunLog("Quiet", str(tmp_path / "quiet"), False)
    q.info("x")
    q.close()
    assert not os.path.exists(str(tmp_path / "quiet.log"))


@pytest.mark.parametrize("kind", ["hky", "c10_g4"])
def test_show_model_carries_the_model(kind):
    model = pm.get_phylo_model("HKY[6.0]{0.2,0.3,0.3,0.2}") if kind == "hky" else \
        pm.get_phylo_model(mixture_model="C10", global_normalization=True, gamma=(4, 0.2))
    text = cli.show_model(model)
    assert text.startswith("SubstitutionModel (SubstitutionModel {" if kind == "hky" else "MixtureModel (MixtureModel {name = ")
    assert text.count("stationaryDistribution = [") == model.n_components
    k = model.n_states
    vecs = re.findall(r"stationaryDistribution = \[([^\]]*)\]", text)
    got = np.array([[float(x) for x in v.split(",")] for v in vecs])
    assert np.array_equal(got, model.pi)  # `show` of a Double round-trips
    mats = re.findall(r"\(%d><%d\)\n \[ (.*?) \]" % (k, k), text, flags=re.S)
    assert len(mats) == model.n_components
    m0 = np.array([[float(x) for x in row.split(",")] for row in mats[0].split("\n , ")])
    assert np.array_equal(m0, model.exch[0])
    if model.is_mixture:
        ws = [float(x) for x in re.findall(r"Component \{weight = ([^,]*),", text)]
        assert np.array_equal(ws, model.weights)


def test_reproduction_file(tmp_path):
    base = str(tmp_path / "r")
    tree = str(tmp_path / "t.tree")
    open(tree, "w").write("(a:0.1,b:0.2);")
    open(base + ".fasta", "w").write(">a\nAC\n>b\nAG\n")
    with gzip.open(base + ".model.gz", "wb") as f:
        f.write(b"SubstitutionModel (...)\n")
    args = types.SimpleNamespace(tree_file=tree, substitution_model="JC", mixture_model=None, global_normalization=False, edm_file=None,
                                 siteprofile_files=None, mixture_model_weights=None, gamma_rate_heterogeneity="(4,0.5)", length=2, seed="[7]",
                                 verbosity="Info", force=True)
    argv = ["-f", "-o", base, "simulate", "-t", tree, "-s", "JC", "-g", "(4,0.5)", "-l", "2", "-S", "[7]"]
    cli.write_reproduction(base, argv, args, 7, [tree])
    rec = json.load(open(base + ".elynx"))
    assert rec["progName"] == "slynx-b200" and rec["argsStr"] == argv and rec["rVersion"] == cli.ELYNX_VERSION
    assert rec["files"] == [tree, base + ".model.gz", base + ".fasta"]
    assert rec["checkSums"] == [hashlib.sha256(open(f, "rb").read()).hexdigest() for f in rec["files"]]
    loc, glob = rec["reproducible"]["local"], rec["reproducible"]["global"]
    assert loc["argsSeed"] == {"tag": "Fixed", "contents": 7} and loc["argsGammaParams"] == [4, 0.5] and loc["argsLength"] == 2
    assert loc["argsMixtureModelGlobalNormalization"] == "LocalNormalization" and loc["argsEDMFile"] is None
    assert glob == {"verbosity": "Info", "outFileBaseName": base, "executionMode": "Overwrite", "writeElynxFile": True}
    assert re.fullmatch(r"[0-9a-f]{64}", rec["rHash"])
    # the hash covers the check sums: changing an output changes it
    open(base + ".fasta", "a").write("x")
    cli.write_reproduction(base, argv, args, 7, [tree])
    assert json.load(open(base + ".elynx"))["rHash"] != rec["rHash"]


def test_site_dists_numbering():
    model = pm.get_phylo_model(mixture_model="C10", global_normalization=True)
    a = cli.site_dists(np.array([3, 1, 0]), model).split(b"\n")
    b = cli.site_dists(np.array([1, 0]), model, first=1).split(b"\n")
    assert a[1:] == b and a[0].startswith(b"1 ") and b[0].startswith(b"2 ") and len(a[0].split()) == 21
