"""`python -m elynx_b200.cli simulate|examine ...` -- thin native driver with the `slynx simulate` / `slynx examine` surface.

Mirrors slynx/src/SLynx/Simulate/Options.hs:65-204 (arguments) and Simulate.hs:239-306 (simulateCmd): reads a Newick
tree, assembles the phylogenetic model from the model string (+ EDM / site-profile files, weights, gamma rate
heterogeneity), simulates the alignment on the GPU and writes `<base>.fasta` (and `<base>.sitedists` for mixture
models, PAML order, Simulate.hs:79-89), `<base>.model.gz` (Simulate.hs:137-151), `<base>.log` (the ELynx logger, elynx-tools
Logger.hs) and the `<base>.elynx` reproduction file (Reproduction.hs:131-145); existing outputs need -f/--force
(InputOutput.hs:52-65).  A GPU counter-based stream replaces SplitMix, so the alignment differs from the Haskell binary's for
the same seed; `elynx validate/redo` are out of scope."""
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


def site_dists(comps, model, first: int = 0) -> bytes:
    """writeSiteDists (Simulate.hs:79-89): `i pi...` per site, 1-based, component pi in PAML order."""
    alpha = "ACDEFGHIKLMNPQRSTVWY"
    if model.alphabet == pm.PROTEIN:
        perm = [alpha.index(c) for c in AA_PAML_ORDER]
    else:
        raise SystemExit("site distributions are written in PAML order and only make sense for protein models")
    rows = [" ".join(double_dec(v) for v in model.pi[c][perm]) for c in range(model.n_components)]
    return "\n".join("%d %s" % (first + i + 1, rows[c]) for i, c in enumerate(comps)).encode()


def write_gz(path: str, data: bytes, threads: int = 0) -> None:
    """writeGZFile (elynx-tools InputOutput.hs:80-83) through the parallel multi-member writer elx_gzip_write."""
    import ctypes as C

    from . import _lib as L

    buf = (C.c_char * len(data)).from_buffer_copy(data) if data else None
    L.check(L.lib().elx_gzip_write(path.encode(), C.cast(buf, C.c_void_p) if buf is not None else None, len(data), threads, 6, 0))


ELYNX_VERSION = "0.9.0.0"
HLINE = "-" * 78


def check_file(path: str, force: bool) -> None:
    """checkFile (elynx-tools InputOutput.hs:52-65): existing results are only overwritten with -f/--force."""
    import os

    if os.path.exists(path) and not force:
        raise SystemExit("File exists: %s.\nPlease use --force to overwrite results of a previous analysis." % path)


def render_time(t: float) -> str:
    """renderTime (Logger.hs:186): "%B %-e, %Y, at %H:%M %P, %Z." in UTC."""
    tm = time.gmtime(t)
    return time.strftime("%B ", tm) + str(tm.tm_mday) + time.strftime(", %Y, at %H:%M ", tm) + ("am" if tm.tm_hour < 12 else "pm") + ", UTC."


class RunLog:
    """The ELynx logger (elynx-tools Logger.hs:95-212, Environment.hs:57-70): info lines carry a three-space prefix and go to
    standard output and, with -o, to `<base>.log`; `-v Quiet` silences both and writes no log file."""

    def __init__(self, verbosity: str, base, force: bool, stream=None):
        self.level = ["Quiet", "Warn", "Info", "Debug"].index(verbosity)
        self.t0 = time.time()
        self.handles = []
        self.file = None
        if self.level > 0:
            self.handles.append(stream or sys.stdout)
            if base:
                check_file(base + ".log", force)
                self.file = open(base + ".log", "w")
                self.handles.append(self.file)

    def _out(self, prefix: str, msg: str) -> None:
        for h in self.handles:
            h.write("".join(prefix + line + "\n" for line in msg.split("\n")))
            h.flush()

    def info(self, msg: str) -> None:
        if self.level >= 2:
            self._out("   ", msg)

    def warn(self, msg: str) -> None:
        if self.level >= 1:
            self._out("W: ", msg)

    def header(self, cmd_name: str, cmd_dsc, argv) -> None:
        """logInfoHeader (Logger.hs:188-202)"""
        self.info(HLINE)
        self.info("\n".join(["ELynx Suite version %s (elynx_b200: B200-native simulator)." % ELYNX_VERSION, "Developed by Dominik Schrempf.",
                             "CUDA path: elynx_b200/libelynx_b200.so."]))
        self.info("\n".join([HLINE, "Command name: " + cmd_name] + list(cmd_dsc) +
                            ["Starting time: " + render_time(self.t0), "Command line: slynx-b200 " + " ".join(argv)]))
        self.info(HLINE)

    def footer(self) -> None:
        """logInfoFooter (Logger.hs:205-211)"""
        te = time.time()
        ts = int(round(te - self.t0))
        self.info("Wall clock run time: %02d:%02d:%02d." % (ts // 3600, ts % 3600 // 60, ts % 60))
        self.info("End time: " + render_time(te))

    def close(self) -> None:
        if self.file:
            self.file.close()
            self.file = None


def show_double(x: float) -> str:
    return double_dec(float(x))


def show_model(model) -> str:
    """`show` of the PhyloModel (Simulate.hs:137-151 writes it to <base>.model.gz): the record structure of
    MarkovProcess/{PhyloModel,MixtureModel,SubstitutionModel}.hs with hmatrix's renderings of vectors and matrices.  The numbers
    are the model's (content parity); per-component names and parameter lists are not carried by the flat model, so the byte
    stream differs from GHC's."""
    k = model.n_states

    def vec(v):
        return "[" + ",".join(show_double(x) for x in v) + "]"

    def mat(m):
        cells = [[show_double(x) for x in row] for row in m]
        w = max(len(c) for row in cells for c in row)
        rows = [", ".join(c.rjust(w) for c in row) for row in cells]
        return "(%d><%d)\n [ " % (k, k) + "\n , ".join(rows) + " ]"

    def sm(c, name):
        return ("SubstitutionModel {alphabet = %s, name = %s, params = [], stationaryDistribution = %s, exchangeabilityMatrix = %s}"
                % (model.alphabet, _hs_str(name), vec(model.pi[c]), mat(model.exch[c])))

    if not model.is_mixture:
        return "SubstitutionModel (%s)" % sm(0, model.name)
    comps = ",".join("Component {weight = %s, substModel = %s}" % (show_double(model.weights[c]), sm(c, "%s; component %d" % (model.name, c)))
                     for c in range(model.n_components))
    return "MixtureModel (MixtureModel {name = %s, alphabet = %s, components = [%s]})" % (_hs_str(model.name), model.alphabet, comps)


def _hs_str(x: str) -> str:
    return '"' + x.replace("\\", "\\\\").replace('"', '\\"') + '"'


def sha256_file(path: str) -> str:
    import hashlib

    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
    return h.hexdigest()


SIMULATE_DSC = ["Simulate multi sequence alignments."]
SIMULATE_OUT_SUFFIXES = [".model.gz", ".fasta"]  # Options.hs:83


def write_reproduction(base: str, argv, args, seed: int, in_files) -> None:
    """writeReproduction (elynx-tools Reproduction.hs:131-145): `<base>.elynx`, the JSON of the Reproduction record -- program
    name, arguments, version, input files + output files (base ++ outSuffixes) with their SHA256 sums, the parsed arguments
    (aeson's generic encoding), and the hash over all of it (getReproductionHash, :77-97)."""
    import hashlib
    import json

    files = list(in_files) + [base + sfx for sfx in SIMULATE_OUT_SUFFIXES]
    sums = [sha256_file(f) for f in files]
    local = {
        "argsTreeFile": args.tree_file,
        "argsSubstitutionModelString": args.substitution_model,
        "argsMixtureModelString": args.mixture_model,
        "argsMixtureModelGlobalNormalization": "GlobalNormalization" if args.global_normalization else "LocalNormalization",
        "argsEDMFile": args.edm_file,
        "argsSiteprofilesFiles": args.siteprofile_files,
        "argsMixtureWeights": _parse_weights(args.mixture_model_weights) if args.mixture_model_weights else None,
        "argsGammaParams": list(_parse_gamma(args.gamma_rate_heterogeneity)) if args.gamma_rate_heterogeneity else None,
        "argsLength": args.length,
        "argsSeed": {"tag": "Fixed" if args.seed is not None else "RandomSet", "contents": seed},
    }
    glob = {"verbosity": args.verbosity, "outFileBaseName": base, "executionMode": "Overwrite" if args.force else "Fail",
            "writeElynxFile": True}
    lines = ["slynx-b200"] + list(argv) + [ELYNX_VERSION] + files + sums + list(in_files) + SIMULATE_OUT_SUFFIXES + ["simulate"] + SIMULATE_DSC
    rhash = hashlib.sha256("".join(x + "\n" for x in lines).encode()).hexdigest()
    rec = {"progName": "slynx-b200", "argsStr": list(argv), "rVersion": ELYNX_VERSION, "rHash": rhash, "files": files, "checkSums": sums,
           "reproducible": {"global": glob, "local": local}}
    with open(base + ".elynx", "w") as f:
        json.dump(rec, f)


def simulate_cmd(args, argv=None) -> int:
    """simulateCmd (Simulate.hs:239-306) inside eLynxWrapper (ELynx.hs:51-101): header, seed line, the worker, the
    reproduction file, footer."""
    import os

    argv = list(sys.argv[1:] if argv is None else argv)
    base, force = args.output_file_basename, args.force
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    # logs go to standard output like the reference's; when the alignment itself goes there (no -o) they move to stderr
    log = RunLog(args.verbosity if rank == 0 else "Quiet", base, force, stream=None if base else sys.stderr)
    try:
        log.header("simulate", SIMULATE_DSC, argv)
        seed = _parse_seed(args.seed) if args.seed is not None else int(time.time_ns() & (2 ** 63 - 1))
        log.info("Seed: fixed to %d." % seed if args.seed is not None else "Seed: random; set to %d." % seed)
        rc = _simulate_worker(args, log, seed, world, rank)
        if rc == 0 and rank == 0:
            if args.no_elynx_file:
                log.info("No elynx file option --- skip writing ELynx file for reproducible runs.")
            elif not base:
                log.info("No output file given --- skip writing ELynx file for reproducible runs.")
            else:
                log.info("Write ELynx reproduction file.")
                in_files = [args.tree_file] + ([args.edm_file] if args.edm_file else []) + list(args.siteprofile_files or [])
                if os.path.exists(base + ".fasta"):
                    write_reproduction(base, argv, args, seed, in_files)
                else:
                    log.warn("'%s.fasta' was not written (--fasta-gz); the reproduction file needs it: skipped." % base)
        log.footer()
        return rc
    finally:
        log.close()


def _simulate_worker(args, log, seed: int, world: int, rank: int) -> int:
    import os

    base, force = args.output_file_basename, args.force
    log.info("Read tree from file '%s'." % args.tree_file)
    tree = pm.parse_newick(open(args.tree_file).read())
    log.info("Number of leaves: %d" % tree.n_leaves)
    if args.edm_file and args.siteprofile_files:
        raise SystemExit("Got both: --edm-file and --siteprofile-files.")
    comps = None
    if args.edm_file:
        log.info("Read EDM file.")
        comps = pm.parse_edm_phylobayes(open(args.edm_file).read())
    if args.siteprofile_files:
        log.info("Read siteprofiles from %d file(s)." % len(args.siteprofile_files))
        comps = sum((pm.parse_siteprofiles_phylobayes(open(f).read()) for f in args.siteprofile_files), [])
    if comps is not None:
        log.info("Empiricial distribution mixture model with %d components." % len(comps))
    if args.mixture_model_weights:
        log.info("Read mixture weights from command line.")
    model = pm.get_phylo_model(args.substitution_model, args.mixture_model, args.global_normalization,
                               _parse_weights(args.mixture_model_weights) if args.mixture_model_weights else None, comps,
                               _parse_gamma(args.gamma_rate_heterogeneity) if args.gamma_rate_heterogeneity else None)
    if args.gamma_rate_heterogeneity:
        n, alpha = _parse_gamma(args.gamma_rate_heterogeneity)
        log.info("Discrete gamma rate heterogeneity: %d categories, shape parameter %s." % (n, show_double(alpha)))
    log.info("%s model '%s': %d states, %d component(s)." % (model.alphabet, model.name, model.n_states, model.n_components))
    # reportModel (Simulate.hs:137-151)
    if rank == 0:
        if args.no_elynx_file:
            log.info("No elynx file required; omit writing machine-readable phylogenetic model.")
        elif not base:
            log.info("No output file provided; omit writing machine-readable phylogenetic model.")
        else:
            fn = base + ".model.gz"
            log.info("Write model definition (machine readable) to file '%s'." % fn)
            check_file(fn, force)
            with gzip.open(fn, "wb") as f:
                f.write((show_model(model) + "\n").encode())
    log.info("Simulate alignment.")
    log.info("Length: %d." % args.length)
    want_dists = model.is_mixture and model.alphabet == pm.PROTEIN
    names, chars = tree.leaf_names(), model.alphabet_chars
    if world > 1:
        # torchrun: one process per GPU on ONE node (the ranks share the output file: a common file system is assumed), every
        # rank packs its site range into its own pinned buffer (K3 slice mode) and writes its row segments of <base>.fasta; no
        # data-path collective (elynx_b200/distributed.py).  --device is ignored: rank r drives GPU LOCAL_RANK.
        import torch
        import torch.distributed as dist

        from .distributed import simulate_fasta_to_file, site_range

        if not base:
            raise SystemExit("multi-GPU runs need -o/--output-file-basename")
        if args.seed is None:
            raise SystemExit("multi-GPU runs need -S/--seed (every rank must use the same seed)")
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        fn = base + ".fasta"
        if rank == 0:
            check_file(fn, force)
            log.info("Write simulated multi sequence alignment to file '%s'." % fn)
        with Simulator(model.weights if model.is_mixture else None, model.pi, model.exch, tree, is_mixture=model.is_mixture,
                       device=local) as sim:
            simulate_fasta_to_file(sim, seed, args.length, names, chars, fn, rank=rank, world=world, barrier=dist.barrier)
            if not os.path.exists(fn):
                raise SystemExit("rank %d cannot see '%s': multi-GPU runs need all ranks on one file system" % (rank, fn))
            if want_dists:  # every rank writes the lines of its site range, rank 0 joins them (writeSiteDists, Simulate.hs:79-89)
                s0, n = site_range(rank, world, args.length, align=16)
                cs = sim.site_components(seed, n, site0=s0) if n > 0 else np.zeros(0, dtype=np.int32)
                text = site_dists(cs, model, first=s0)
                open("%s.sitedists.part%d" % (base, rank), "wb").write(text + (b"\n" if n > 0 and s0 + n < args.length else b""))
                dist.barrier()
                if rank == 0:
                    check_file(base + ".sitedists", force)
                    with open(base + ".sitedists", "wb") as f:
                        for r in range(world):
                            part = "%s.sitedists.part%d" % (base, r)
                            f.write(open(part, "rb").read())
                            os.remove(part)
        dist.destroy_process_group()
        log.info("Done with %d GPUs." % world)
        return 0
    with Simulator(model.weights if model.is_mixture else None, model.pi, model.exch, tree, is_mixture=model.is_mixture,
                   device=args.device, n_devices=args.gpus) as sim:
        if sim.n_devices > 1:
            log.info("One handle drives %d GPUs (contiguous site ranges, getChunks rule)." % sim.n_devices)
        if base and args.fasta_gz:
            fn = base + ".fasta.gz"
            log.info("Write simulated multi sequence alignment to file '%s'." % fn)
            check_file(fn, force)
            cs = sim.simulate_fasta_gz(seed, args.length, names, chars, fn, want_comps=want_dists)
            fasta = None
        else:
            fasta, cs = sim.simulate_fasta(seed, args.length, names, chars, want_comps=want_dists)
    if want_dists and base:
        check_file(base + ".sitedists", force)
        open(base + ".sitedists", "wb").write(site_dists(cs, model))
    log.info("")
    if fasta is not None:
        if base:
            fn = base + ".fasta"
            log.info("Write simulated multi sequence alignment to file '%s'." % fn)
            check_file(fn, force)
            open(fn, "wb").write(fasta)
        else:
            log.info("Write simulated multi sequence alignment to standard output.")
            sys.stdout.buffer.write(fasta)
            sys.stdout.buffer.flush()
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
    ap.add_argument("-f", "--force", action="store_true", help="Ignore previous analysis and overwrite existing output files.")
    ap.add_argument("--no-elynx-file", action="store_true", help="Do not write data required to reproduce an analysis.")
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
    s.add_argument("--device", type=int, default=-1, help="first CUDA device (extension)")
    s.add_argument("--gpus", type=int, default=1, help="GPUs one handle drives inside the call, -1 = all visible (extension)")
    s.add_argument("--fasta-gz", action="store_true", help="write <base>.fasta.gz through the streaming gzip writer (extension)")
    e = sub.add_parser("examine", help="Examine sequences / alignments (slynx/src/SLynx/Examine/Options.hs)")
    e.add_argument("-a", "--alphabet", required=True, help="DNA, DNAX, Protein or ProteinX")
    e.add_argument("input_file", metavar="INPUT-FILE")
    e.add_argument("--per-site", action="store_true", help="Report per site summary statistics")
    e.add_argument("--divergence", action="store_true", help="Compute pairwise divergence matrices")
    args = ap.parse_args(argv)
    return examine_cmd(args) if args.cmd == "examine" else simulate_cmd(args, argv)


if __name__ == "__main__":
    sys.exit(main())
