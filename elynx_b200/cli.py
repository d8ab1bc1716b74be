"""`python -m elynx_b200.cli simulate|examine ...` -- thin native driver with the `slynx simulate` / `slynx examine` surface.

Mirrors slynx/src/SLynx/Simulate/Options.hs:65-204 (arguments) and Simulate.hs:239-306 (simulateCmd): reads a Newick
tree, assembles the phylogenetic model from the model string (+ EDM / site-profile files, weights, gamma rate
heterogeneity), simulates the alignment on the GPU and writes `<base>.fasta` (and `<base>.sitedists` for mixture
models, PAML order, Simulate.hs:79-89).  Out of scope: `.elynx` reproduction files,
`.model.gz`, the ELynx logger; a GPU counter-based stream replaces SplitMix, so the bytes differ from the Haskell
binary for the same seed."""
from __future__ import annotations

import argparse
import gzip
import re
import sys
import time

import numpy as np

from . import model as pm
from .simulate import Simulator

AA_PAML_ORDER = "ARNDCQEGHILKMFPSTWYV"  # AminoAcid.hs:74-99


def _parse_weights(s):
    """`-w "[0.9,0.05,...]"` (Options.hs mixture-model-weights: a Haskell list literal)"""
    s = s.strip()
    if not (s.startswith("[") and s.endswith("]")):
        raise SystemExit("mixture weights must look like [w1,w2,...]")
    return [float(x) for x in s[1:-1].split(",")]


def _parse_gamma(s):
    """`-g "(4,0.5)"` (Options.hs gamma-rate-heterogeneity: (number of categories, alpha))"""
    m = re.fullmatch(r"\(\s*(\d+)\s*,\s*([0-9.eE+-]+)\s*\)", s.strip())
    if not m:
        raise SystemExit("gamma rate heterogeneity must look like (n,alpha)")
    return int(m.group(1)), float(m.group(2))


def _parse_seed(s):
    """`-S [0]` / `-S 0`: the reference takes an Int since random-1.2"""
    return int(s.strip().strip("[]").split(",")[0])


def double_dec(x: float) -> str:
    """ByteString.Builder.doubleDec = Haskell `show`: shortest round-trip digits, scientific outside [0.1, 10^7)."""
    if x == 0:
        return "0.0"
    r = repr(abs(float(x)))
    sign = "-" if x < 0 else ""
    mant, ex = (r.lower().split("e") + ["0"])[:2]
    ex = int(ex)
    ip, fp = (mant.split(".") + [""])[:2]
    digits = (ip + fp).lstrip("0")
    lead = len(ip.lstrip("0"))
    e10 = (lead - 1 + ex) if lead else (-(len(fp) - len(fp.lstrip("0"))) - 1 + ex)
    digits = digits.rstrip("0") or "0"
    if 0.1 <= abs(x) < 1e7:
        if e10 >= 0:
            return sign + digits[: e10 + 1].ljust(e10 + 1, "0") + "." + (digits[e10 + 1:] or "0")
        return sign + "0." + "0" * (-e10 - 1) + digits
    return sign + digits[0] + "." + (digits[1:] or "0") + "e" + str(e10)


def site_dists(comps, model) -> bytes:
    """writeSiteDists (Simulate.hs:79-89): `i pi...` per site, 1-based, component pi in PAML order."""
    alpha = "ACDEFGHIKLMNPQRSTVWY"
    if model.alphabet == pm.PROTEIN:
        perm = [alpha.index(c) for c in AA_PAML_ORDER]
    else:
        raise SystemExit("site distributions are written in PAML order and only make sense for protein models")
    rows = [" ".join(double_dec(v) for v in model.pi[c][perm]) for c in range(model.n_components)]
    return "\n".join("%d %s" % (i + 1, rows[c]) for i, c in enumerate(comps)).encode()


def write_gz(path: str, data: bytes, threads: int = 0) -> None:
    """writeGZFile (elynx-tools InputOutput.hs:80-83) through the parallel multi-member writer elx_gzip_write."""
    import ctypes as C

    from . import _lib as L

    buf = (C.c_char * len(data)).from_buffer_copy(data) if data else None
    L.check(L.lib().elx_gzip_write(path.encode(), C.cast(buf, C.c_void_p) if buf is not None else None, len(data), threads, 6, 0))


def simulate_cmd(args) -> int:
    t0 = time.time()
    log = (lambda *a: None) if args.verbosity == "Quiet" else (lambda *a: print(*a, file=sys.stderr))
    log("Read tree from file '%s'." % args.tree_file)
    tree = pm.parse_newick(open(args.tree_file).read())
    log("Number of leaves: %d" % tree.n_leaves)
    if args.edm_file and args.siteprofile_files:
        raise SystemExit("Got both: --edm-file and --siteprofile-files.")
    comps = None
    if args.edm_file:
        comps = pm.parse_edm_phylobayes(open(args.edm_file).read())
    if args.siteprofile_files:
        comps = sum((pm.parse_siteprofiles_phylobayes(open(f).read()) for f in args.siteprofile_files), [])
    if comps is not None:
        log("Empiricial distribution mixture model with %d components." % len(comps))
    model = pm.get_phylo_model(args.substitution_model, args.mixture_model, args.global_normalization,
                               _parse_weights(args.mixture_model_weights) if args.mixture_model_weights else None, comps,
                               _parse_gamma(args.gamma_rate_heterogeneity) if args.gamma_rate_heterogeneity else None)
    log("%s model '%s': %d states, %d component(s)." % (model.alphabet, model.name, model.n_states, model.n_components))
    log("Simulate alignment.")
    log("Length: %d." % args.length)
    seed = _parse_seed(args.seed) if args.seed is not None else int(time.time_ns() & (2 ** 63 - 1))
    import os

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        # torchrun: one process per GPU, every rank packs its site range into its own pinned buffer (K3 slice mode) and
        # writes its row segments of <base>.fasta; no data-path collective (elynx_b200/distributed.py)
        import torch
        import torch.distributed as dist

        from .distributed import simulate_fasta_to_file

        if not args.output_file_basename:
            raise SystemExit("multi-GPU runs need -o/--output-file-basename")
        if args.seed is None:
            raise SystemExit("multi-GPU runs need -S/--seed (every rank must use the same seed)")
        rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        with Simulator(model.weights if model.is_mixture else None, model.pi, model.exch, tree, is_mixture=model.is_mixture,
                       device=local) as sim:
            simulate_fasta_to_file(sim, seed, args.length, tree.leaf_names(), model.alphabet_chars,
                                   args.output_file_basename + ".fasta", rank=rank, world=world, barrier=dist.barrier)
        dist.destroy_process_group()
        if rank == 0:
            log("Wrote simulated multi sequence alignment to '%s.fasta' with %d GPUs." % (args.output_file_basename, world))
        return 0
    with Simulator(model.weights if model.is_mixture else None, model.pi, model.exch, tree, is_mixture=model.is_mixture,
                   device=args.device) as sim:
        fasta, cs = sim.simulate_fasta(seed, args.length, tree.leaf_names(), model.alphabet_chars, want_comps=model.is_mixture)
    if args.output_file_basename:
        fn = args.output_file_basename + ".fasta"
        if fn.endswith(".gz"):
            write_gz(fn, fasta)
        else:
            open(fn, "wb").write(fasta)
        log("Wrote simulated multi sequence alignment to '%s'." % fn)
        if model.is_mixture and model.alphabet == pm.PROTEIN:
            open(args.output_file_basename + ".sitedists", "wb").write(site_dists(cs, model))
    else:
        sys.stdout.buffer.write(fasta)
    log("Done in %.2f s." % (time.time() - t0))
    return 0


def read_fasta(text: bytes):
    """[(name, description, characters)] -- elynx-seq Import/Fasta.hs:33-70 (`>name description`, sequence lines until the
    next header); character validation against the alphabet happens in examine.encode."""
    seqs, name, desc, chunks = [], None, b"", []
    for line in text.splitlines():
        if line.startswith(b">"):
            if name is not None:
                seqs.append((name, desc, b"".join(chunks)))
            parts = line[1:].split(None, 1)
            if not parts:
                raise SystemExit("fasta: empty sequence header")
            name, desc, chunks = parts[0], (parts[1] if len(parts) > 1 else b""), []
        elif line.strip():
            if name is None:
                raise SystemExit("fasta: sequence data before the first header")
            chunks.append(line.strip())
    if name is not None:
        seqs.append((name, desc, b"".join(chunks)))
    if not seqs:
        raise SystemExit("fasta: no sequences")
    return seqs


def summarize_sequences(seqs, alphabet: str) -> str:
    """Seq.summarizeSequences (elynx-seq Sequence.hs:105-170, Defaults.hs:19-34): a table of the first 200 sequences."""
    name_w, field_w, summary_len, summary_n = 23, 13, 60, 200
    al = lambda n, x: x[:n].ljust(n)  # noqa: E731
    ar = lambda n, x: x[:n].rjust(n)  # noqa: E731
    lines = []
    if len(seqs) > summary_n:
        lines.append("%d out of %d sequences are shown." % (summary_n, len(seqs)))
    lines += ["For each sequence, the %d first bases are shown." % summary_len, "List contains %d sequences." % len(seqs), "",
              " ".join([al(name_w, "Name"), ar(field_w, "Code"), ar(field_w, "Length"), ar(field_w, "Gaps [%]"), "Sequence"])]
    for name, _, chars in seqs[:summary_n]:
        n_gaps = sum(chars.count(g) for g in (b"-", b"."))
        pg = 100.0 * n_gaps / len(chars) if chars else float("nan")
        shown = chars[:summary_len] + b"..." if len(chars) >= summary_len else chars
        lines.append(" ".join([al(name_w, name.decode()), ar(field_w, alphabet), ar(field_w, str(len(chars))), ar(field_w, "%2.2f" % pg),
                               shown.decode()]))
    return "\n".join(lines) + "\n"


def examine_cmd(args) -> int:
    """examineCmd (slynx/src/SLynx/Examine/Examine.hs:187-196): `<base>.out` (or stdout) and, with --divergence, `<base>.div`."""
    from . import examine as ex

    if args.alphabet not in ex.ALPHABETS:
        raise SystemExit("alphabet %s is not supported (IUPAC alphabets are out of scope); use one of %s" % (args.alphabet, ", ".join(ex.ALPHABETS)))
    raw = (gzip.open if args.input_file.endswith(".gz") else open)(args.input_file, "rb").read()
    seqs = read_fasta(raw)
    text = summarize_sequences(seqs, args.alphabet)
    k = len(ex.ALPHABETS[args.alphabet])
    if len({len(c) for _, _, c in seqs}) == 1:  # M.fromSequences succeeds only for equal lengths (Alignment.hs:87-95)
        a = ex.encode([c for _, _, c in seqs], args.alphabet)
        r = ex.examine_alignment(a, k, per_site=args.per_site, divergence=args.divergence)
        per_site = r.get("keff_entropy_per_site") if args.per_site else None
        report = ex.format_report(r, args.alphabet)
        if per_site is not None:
            report += "Effective number of used states per site (measured using entropy):\n"
            report += "[" + ",".join(double_dec(float(x)) for x in per_site) + "]\n"
        text += "\n" + report
        if args.divergence:
            if not args.output_file_basename:
                raise SystemExit("--divergence needs -o/--output-file-basename")
            with open(args.output_file_basename + ".div", "w") as f:  # writeDivergenceMatrix (Examine.hs:170-177)
                p = 0
                for i in range(len(seqs)):
                    for j in range(i + 1, len(seqs)):
                        f.write("> %s, %s\n" % (seqs[i][0].decode(), seqs[j][0].decode()))
                        m = r["divergence"][p]
                        w = max(len(str(int(v))) for v in m.flat)
                        f.write("%dx%d\n" % (k, k))                      # hmatrix dispf 0: dimensions, then the rows
                        for row in m:
                            f.write("  ".join(str(int(v)).rjust(w) for v in row) + "\n")
                        p += 1
    if args.output_file_basename:
        open(args.output_file_basename + ".out", "w").write(text)
    else:
        sys.stdout.write(text)
    return 0


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="slynx-b200", description="B200-native `slynx simulate` (alignment simulation on the GPU)")
    ap.add_argument("-v", "--verbosity", default="Info", choices=["Quiet", "Warn", "Info", "Debug"])
    ap.add_argument("-o", "--output-file-basename")
    ap.add_argument("-f", "--force", action="store_true", help="accepted for compatibility (outputs are always overwritten)")
    ap.add_argument("--no-elynx-file", action="store_true", help="accepted for compatibility (no .elynx file is written)")
    sub = ap.add_subparsers(dest="cmd", required=True)
    s = sub.add_parser("simulate", help="Simulate multi sequence alignments (slynx/src/SLynx/Simulate/Options.hs)")
    s.add_argument("-t", "--tree-file", required=True)
    s.add_argument("-s", "--substitution-model")
    s.add_argument("-m", "--mixture-model")
    s.add_argument("-e", "--edm-file")
    s.add_argument("-p", "--siteprofile-files", nargs="+")
    s.add_argument("-w", "--mixture-model-weights")
    s.add_argument("-n", "--global-normalization", action="store_true")
    s.add_argument("-g", "--gamma-rate-heterogeneity")
    s.add_argument("-l", "--length", type=int, required=True)
    s.add_argument("-S", "--seed")
    s.add_argument("--device", type=int, default=-1)
    e = sub.add_parser("examine", help="Examine sequences / alignments (slynx/src/SLynx/Examine/Options.hs)")
    e.add_argument("-a", "--alphabet", required=True, help="DNA, DNAX, Protein or ProteinX")
    e.add_argument("input_file", metavar="INPUT-FILE")
    e.add_argument("--per-site", action="store_true", help="Report per site summary statistics")
    e.add_argument("--divergence", action="store_true", help="Compute pairwise divergence matrices")
    args = ap.parse_args(argv)
    return examine_cmd(args) if args.cmd == "examine" else simulate_cmd(args)


if __name__ == "__main__":
    sys.exit(main())
