#!/usr/bin/env python3
"""Benchmark of the alignment-simulation hot path (BASELINE.json metric: alignment cells/s and HBM GB/s vs peak).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config cfg2|cfg1|cfg3|cfg4|cfg5|...]

A "step" is one pass of the hot path over one batch of synthetic input: the simulation of `sites` alignment
columns for every taxon of the tree (T x S uint8 leaf states written once to HBM, layout [taxon][site]).
At N GPUs every rank simulates its own contiguous site range of an (N x S)-column alignment (weak scaling;
no data-path collective: sites are independent and RNG counters use the global site index).

Workload of the headline line (config.workload): BASELINE.json configs[1] -- HKY + Gamma(4) DNA, 1000-taxon birth-death
tree (tlynx defaults lambda=1, mu=0.5, tree seed 1), 10M sites on one B200.  The same JSON line carries a `configs` block
with every other BASELINE configuration (cfg1 wall time, cfg2 with GTR4, cfg2 on the height-normalised tree, cfg3, cfg4 (i)
C60 and (ii) EDM(LG-Custom) with the C60 profiles, cfg5): cells/s, fraction of the HBM roofline, kernel name.

JSON keys follow the driver contract; in addition
  roofline      the simulation kernel against the measured HBM bandwidth (1 algorithmic byte per cell)
  cpu_baseline  the oracle port of the reference algorithm timed on the box's host cores (rank 0, N=1)
  e2e           the same metric through the reference-facing C-ABI call on HOST buffers (D2H inside the timed region): at N
                GPUs ONE elx_simulate call on ONE handle that drives all N devices (elx_opts.n_devices); the caller's buffer
                is a plain pageable allocation made fresh for every step, as a Haskell caller's mallocForeignPtrBytes is

The reference arm (--impl reference) builds its inputs with oracle/ only: its process never maps the product's library."""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HKY = "HKY[6.0]{0.2,0.3,0.3,0.2}"
GTR = "GTR4[1,2,3,4,5,6]{0.2,0.3,0.4,0.1}"
CONFIGS = {
    # name: workload description, taxa (None: the bundled Newick.tree), tree seed, model, sites per GPU
    "cfg1": dict(desc="slynx simulate JC DNA, 10k sites, bundled 40-taxon Newick.tree", taxa=None, tseed=None, sm="JC", sites=10_000),
    "cfg2": dict(desc=HKY + " + Gamma(4,0.5) DNA, 1000-taxon birth-death tree (lambda=1, mu=0.5, seed 1), 10M sites",
                 taxa=1000, tseed=1, sm=HKY, gamma=(4, 0.5), sites=10_000_000),
    "cfg2gtr": dict(desc=GTR + " + Gamma(4,0.5) DNA, 1000-taxon birth-death tree (seed 1), 10M sites",
                    taxa=1000, tseed=1, sm=GTR, gamma=(4, 0.5), sites=10_000_000),
    "cfg2h": dict(desc=HKY + " + Gamma(4,0.5) DNA, the cfg2 tree scaled to height 1.0 (SURVEY 8d), 10M sites",
                  taxa=1000, tseed=1, sm=HKY, gamma=(4, 0.5), sites=10_000_000, height=1.0),
    "cfg3": dict(desc="LG + Gamma(4,0.5) protein, 1000-taxon birth-death tree (seed 1), 12.5M sites per GPU (10^8 sites over 8 GPUs)",
                 taxa=1000, tseed=1, sm="LG", gamma=(4, 0.5), sites=12_500_000),
    "cfg4": dict(desc="C60 (-m C60 -n) profile mixture, 500-taxon birth-death tree (seed 2), 10M sites", taxa=500, tseed=2,
                 mm="C60", global_norm=True, sites=10_000_000),
    "cfg4edm": dict(desc="EDM(LG-Custom) -n with the 60 C60 profiles as a Phylobayes EDM file, 500-taxon birth-death tree (seed 2), 10M sites",
                    taxa=500, tseed=2, mm="EDM(LG-Custom)", global_norm=True, edm="c60", sites=10_000_000),
    "cfg5": dict(desc="WAG + Gamma(4,0.5) protein, 10000-taxon birth-death tree (seed 3), 1.25M sites per GPU (10^7 sites over 8 GPUs)",
                 taxa=10000, tseed=3, sm="WAG", gamma=(4, 0.5), sites=1_250_000),
}
EXTRA_CONFIGS = ["cfg2gtr", "cfg2h", "cfg3", "cfg4", "cfg4edm", "cfg5"]  # the `configs` block (cfg1 is a wall-time entry)


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def mark(self):
        return len(self.lines)

    def summary(self, first=0):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[first:]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}

    def stop(self, first=0):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        return self.summary(first)


# ----------------------------------------------------------------------------------------------------
# inputs.  The product arm builds them with the product's host code (C++ behind include/elynx_b200_model.h), the reference
# arm with the oracle's restatements only (oracle.model, oracle.birthdeath) -- same trees, same models (tests hold the two
# constructions together: tests/test_host_model_cpu.py, tests/test_birthdeath_cpu.py).
# ----------------------------------------------------------------------------------------------------
class _Obj:
    pass


def _edm_text(weights, profiles):
    """the Phylobayes EDM file format (EDMModelPhylobayes.hs:33-64)"""
    lines = ["20 A C D E F G H I K L M N P Q R S T V W Y", str(len(weights))]
    for w, d in zip(weights, profiles):
        lines.append(" ".join(repr(float(x)) for x in [w] + list(d)))
    return "\n".join(lines) + "\n"


def build_inputs(cfg_name):
    """(desc, tree, model, sites) with the product's host-side model / tree code"""
    from elynx_b200 import model as pm

    cfg = CONFIGS[cfg_name]
    if cfg["taxa"] is None:
        tree = pm.parse_newick(open(os.path.join(ROOT, "tests", "golden", "ref_data", "Newick.tree")).read())
    else:
        tree = pm.birth_death_tree(cfg["taxa"], 1.0, 0.5, seed=cfg["tseed"])
    if cfg.get("height"):
        tree = tree.scaled(cfg["height"] / tree.height())
    edm = None
    if cfg.get("edm") == "c60":  # the 60 C60 profiles + weights, written as a Phylobayes file and parsed back
        c60 = pm.get_phylo_model(mixture_model="C60", global_normalization=True)
        edm = pm.parse_edm_phylobayes(_edm_text(c60.weights, c60.pi))
    model = pm.get_phylo_model(substitution_model=cfg.get("sm"), mixture_model=cfg.get("mm"), global_normalization=cfg.get("global_norm", False),
                               edm_components=edm, gamma=cfg.get("gamma"))
    return cfg["desc"], tree, model, cfg["sites"]


def build_inputs_oracle(cfg_name):
    """the same inputs from oracle/ alone (the reference arm must not load the product)"""
    import oracle.birthdeath as obd
    import oracle.model as om

    cfg = CONFIGS[cfg_name]
    tree = _Obj()
    if cfg["taxa"] is None:
        t = om.parse_newick(open(os.path.join(ROOT, "tests", "golden", "ref_data", "Newick.tree")).read())
        tree.parent, tree.brlen, tree.leaf_row = np.asarray(t.parent, dtype=np.int32), np.asarray(t.brlen, dtype=np.float64), np.asarray(t.leaf_row, dtype=np.int32)
    else:
        tree.parent, tree.brlen, tree.leaf_row, _ = obd.birth_death_tree(cfg["taxa"], 1.0, 0.5, cfg["tseed"])
    if cfg.get("height"):
        hs, origin = obd.node_heights(tree.parent, tree.brlen, tree.leaf_row)
        tree.brlen = tree.brlen * (cfg["height"] / origin)
    tree.n_nodes = len(tree.parent)
    tree.n_leaves = int((tree.leaf_row >= 0).sum())
    edm = None
    if cfg.get("edm") == "c60":
        c60 = om.cxx(60)
        edm = om.parse_edm_phylobayes(_edm_text(c60.weights, [sm.stationary_distribution for sm in c60.models]))
    pmod = om.get_phylo_model(cfg.get("sm"), cfg.get("mm"), global_norm=cfg.get("global_norm", False), edm_cs=edm)
    if cfg.get("gamma"):
        pmod = om.expand(cfg["gamma"][0], cfg["gamma"][1], pmod)
    model = _Obj()
    model.is_mixture, model.weights, model.pi, model.exch = om.flatten_model(pmod)
    model.n_components, model.n_states = model.pi.shape
    return cfg["desc"], tree, model, cfg["sites"]


def config_block(desc, tree, model, sites, world):
    """identical in both arms (the driver compares them)"""
    t = tree.n_leaves
    return {"workload": desc, "taxa": t, "nodes": tree.n_nodes, "sites_per_gpu": sites, "states": model.n_states,
            "components": model.n_components, "parallelism": "site-range sharding x%d, no data-path collective" % world,
            "l2": "output %d MB per step >> 126 MB L2, no flush needed" % (t * sites // 2 ** 20)}


# ----------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference algorithm on the host cores
# ----------------------------------------------------------------------------------------------------
def cpu_reference_rate(tree, model, target_seconds, cores, steps=1, warmup=0):
    """cells/s of the oracle (C restatement of simulateAndFlatten[MixtureModel]Par, chunked over `cores`
    threads by the getChunks rule) on a bounded sample of the same workload."""
    import oracle.model as om
    import oracle.sim as osim

    osim.build()
    p = om.prob_tables(model.pi, model.exch, tree.brlen)
    cum = np.ascontiguousarray(np.cumsum(p, axis=-1))
    t = tree.n_leaves

    def run(s):
        t0 = time.perf_counter()
        osim.simulate_splitmix(model.is_mixture, tree.parent, tree.leaf_row, model.weights, model.pi, cum, s, 0, nchunks=cores,
                               nthreads=cores, precomputed_cum=True)
        return time.perf_counter() - t0

    probe = max(cores * 64, 2048)
    dt = run(probe)
    while dt < 2.0 and probe < 16_000_000:  # grow the probe until per-chunk overheads no longer dominate the estimate
        probe *= 4
        dt = run(probe)
    rate = t * probe / dt
    s = int(max(probe, min(rate * target_seconds / t, 50_000_000)))
    for _ in range(warmup):
        run(s)
    times = [run(s) for _ in range(max(steps, 1))]
    best = float(np.mean(times))
    return t * s / best, s, best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    desc, tree, model, sites = build_inputs_oracle(args.config)
    cores = os.cpu_count() or 1
    value, s, dt = cpu_reference_rate(tree, model, target_seconds=10.0, cores=cores, steps=args.steps, warmup=min(args.warmup, 1))
    sample = "%d taxa x %d sites per step (bounded sample of the %d-site workload), tables prebuilt, %d threads" % (tree.n_leaves, s, sites, cores)
    line = {
        "impl": "reference", "metric": "alignment_cells_per_second", "value": value, "unit": "cells/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": config_block(desc, tree, model, sites, args.gpus),
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "C restatement of the reference algorithm (oracle/sim.c), not the Haskell binary (no GHC in the image); "
                                 "inputs built with oracle/ only, the product library is not loaded"},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    assert not any("libelynx_b200" in ln for ln in open("/proc/self/maps")), "the reference arm must not map the product library"
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------------
# product arm
# ----------------------------------------------------------------------------------------------------
def run_product(args):
    import torch
    import torch.distributed as dist

    import elynx_b200 as eb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        host_group = dist.new_group(backend="gloo")  # host-side rendezvous of the e2e section: blocks in a socket, no spinning
    dev = torch.device("cuda", local_rank)
    peak, peak_src = measured_peak_gbs()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def make_sim(model, tree, **kw):
        return eb.Simulator(model.weights if model.is_mixture else None, model.pi, model.exch, tree, is_mixture=model.is_mixture,
                            force_generic=args.force_generic, philox_rounds=args.philox_rounds, single_draws=args.single_draws,
                            pair_draws=args.pair_draws, **kw)

    sampler = ClockSampler(local_rank)
    sampler.start()
    seed = 0
    names = [args.config] + ([] if args.no_configs or args.sites else [c for c in EXTRA_CONFIGS if c != args.config])
    inputs = {n: build_inputs(n) for n in names}
    need = max(inputs[n][1].n_leaves * (((args.sites or inputs[n][3]) + 127) // 128 * 128) for n in names)
    d_buf = torch.empty(need, dtype=torch.uint8, device="cuda")  # one output buffer for every configuration

    def measure(name, steps, warmup, post_roll):
        """device-resident metric of one configuration: K steps timed with CUDA events on the handle's stream, max over ranks"""
        desc, tree, model, sites = inputs[name]
        if args.sites:
            sites = args.sites
        T = tree.n_leaves
        pitch = (sites + 127) // 128 * 128
        site0 = rank * sites  # contiguous site ranges of one (world x sites)-column alignment
        torch.cuda.synchronize()
        t_setup = time.perf_counter()
        sim = make_sim(model, tree, device=local_rank)  # elx_create: eigensystems, H2D of model + tree, K1 (P(t), alias tables)
        sim.synchronize()
        setup_ms = (time.perf_counter() - t_setup) * 1e3
        stream = torch.cuda.ExternalStream(sim.stream, device=dev)

        def step():
            sim.simulate_device(seed, site0, sites, d_buf.data_ptr(), pitch)

        for _ in range(warmup):
            step()
        barrier()
        mark = sampler.mark()
        launches0 = sim.launch_count
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        t0 = time.perf_counter()
        for a, b in evs:
            a.record(stream)
            step()
            b.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        launches = sim.launch_count - launches0
        sim.synchronize()  # raises if any kernel flagged an error (e.g. "categorical: bad weights!")
        # the timed region is shorter than nvidia-smi's sampling period: keep the GPU under the identical load for a short
        # post-roll so that the clock / throttle-reason record has samples taken under this very kernel
        t_roll = time.perf_counter()
        while time.perf_counter() - t_roll < post_roll:
            step()
            torch.cuda.synchronize()
        clocks = sampler.summary(mark)
        kernel_ms = [a.elapsed_time(b) for a, b in evs]
        total_ms = max_over_ranks(evs[0][0].elapsed_time(evs[-1][1]))
        ms_per_step = total_ms / steps
        kms = float(np.mean(kernel_ms))
        achieved = T * sites / (kms * 1e-3) / 1e9
        res = dict(desc=desc, tree=tree, model=model, sites=sites, T=T, value=world * T * sites / (ms_per_step * 1e-3), ms_per_step=ms_per_step,
                   kernel_ms=kms, achieved=achieved, frac=achieved / peak, kernel=sim.last_kernel, launches=int(launches), setup_ms=setup_ms,
                   wall_ms=wall / steps * 1e3, clocks=clocks, quad_groups=sim.quad_groups, pair_groups=sim.pair_groups, duo_groups=sim.duo_groups)
        sim.close()
        return res

    def draws_text(r):
        if r["quad_groups"]:
            return "one draw per cap of <= 4 frontier nodes (inner nodes summed out), %d draws per site" % r["quad_groups"]
        if r["duo_groups"]:
            return ("duo mode: one draw per cap of <= 2 frontier nodes (inner nodes summed out), %d draws per site, on component-bucketed "
                    "sites; a step = bucketing passes + %s + the un-permute pass back to site order, all inside the timed region"
                    % (r["duo_groups"], r["kernel"].split("<")[0]))
        return "joint sibling pairs" if r["pair_groups"] else "node by node"

    head = measure(args.config, args.steps, args.warmup, 0.6)
    desc, tree, model, sites, T = head["desc"], head["tree"], head["model"], head["sites"], head["T"]
    site0 = rank * sites

    # ---- the other BASELINE configurations (same protocol, 3 timed steps each)
    configs = {}
    for name in names[1:]:
        try:
            r = measure(name, 3, 3, 0.3)
            configs[name] = {"workload": r["desc"], "taxa": r["T"], "sites_per_gpu": r["sites"], "components": r["model"].n_components,
                             "states": r["model"].n_states, "value": r["value"], "unit": "cells/s", "ms_per_step": r["ms_per_step"],
                             "kernel": r["kernel"], "gpu_launches_per_step": r["launches"] / 3, "draws": draws_text(r),
                             "setup_ms_once_per_handle": round(r["setup_ms"], 2),
                             "roofline": {"bound": "hbm", "achieved": r["achieved"], "peak": peak, "unit": "GB/s", "frac": r["frac"],
                                          "kernel_ms": r["kernel_ms"], "algorithmic_bytes_per_launch": r["T"] * r["sites"]},
                             "clocks": r["clocks"]}
        except Exception as e:  # a configuration that fails must not take the headline down with it
            configs[name] = {"error": "%s: %s" % (type(e).__name__, e)}
    if "cfg1" not in names and not args.no_configs and not args.sites:
        # cfg1 is latency-bound (0.4 MB of output): wall time of the whole reference-facing call, handle creation included
        from elynx_b200 import model as pm
        d1, t1, m1, s1 = build_inputs("cfg1")
        ts = []
        for _ in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with make_sim(m1, t1, device=local_rank) as s_:
                s_.simulate_fasta(seed, s1, t1.leaf_names(), m1.alphabet_chars)
            ts.append((time.perf_counter() - t0) * 1e3)
        configs["cfg1"] = {"workload": d1, "taxa": t1.n_leaves, "sites": s1, "wall_ms": float(np.median(ts)), "wall_ms_all": [round(x, 2) for x in ts],
                           "includes": "elx_create + elx_simulate_fasta into a host buffer + elx_destroy (median of 5)",
                           "value": t1.n_leaves * s1 / (float(np.median(ts)) * 1e-3), "unit": "cells/s"}

    # ---- end to end through the reference-facing call on host buffers
    e2e = None
    if not args.no_e2e:
        import ctypes as C
        from elynx_b200 import _lib as L

        e2e_sites = min(sites, args.e2e_sites)
        # ONE call on ONE multi-device handle simulates the whole (world x sites) alignment; the host buffer is kept under ~24 GB
        total_sites = max(64, min(e2e_sites * world, int(24e9 / T) // 64 * 64))
        k_e2e = max(1, min(args.steps, 5))
        if rank == 0 and world > 1 and "ELX_HOST_THREADS" not in os.environ:
            os.environ["ELX_HOST_THREADS"] = str(max(1, (os.cpu_count() or 4) - 1))  # this process alone uses the box's cores here

        def host_barrier():
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier(group=host_group)

        def e2e_variant(alloc, label, packed=False):
            """rank 0 drives all `world` GPUs from inside elx_simulate; the other ranks idle at the barriers"""
            step_ms, d2h = [], 0
            ppitch = (total_sites + 31) // 32 * (8 if model.n_states <= 4 else 20)
            for it in range(k_e2e + 1):  # first iteration untimed
                host_barrier()
                if rank == 0:
                    t1 = time.perf_counter()
                    host = alloc((T, ppitch if packed else total_sites))
                    s2 = make_sim(model, tree, device=0, n_devices=world)
                    if packed:
                        L.check(L.lib().elx_simulate_packed(s2.handle, C.c_uint64(seed), 0, total_sites, C.c_void_p(host.ctypes.data), ppitch, None), s2.handle)
                    else:
                        L.check(L.lib().elx_simulate(s2.handle, C.c_uint64(seed), 0, total_sites, C.c_void_p(host.ctypes.data), total_sites, None, None, 0),
                                s2.handle)
                    d2h = s2.d2h_bytes
                    s2.close()
                    dt = (time.perf_counter() - t1) * 1e3
                    if it:
                        step_ms.append(dt)
                    del host
                host_barrier()
            if rank != 0:
                return None
            return {"label": label, "ms_steps": [round(x, 2) for x in step_ms], "ms_mean": float(np.mean(step_ms)), "ms_median": float(np.median(step_ms)),
                    "value_mean": T * total_sites / (float(np.mean(step_ms)) * 1e-3), "value_median": T * total_sites / (float(np.median(step_ms)) * 1e-3),
                    "d2h_bytes_per_step": int(d2h)}

        pinned_keep = []

        def alloc_pinned(shape):
            if not pinned_keep:
                pinned_keep.append(torch.empty(shape, dtype=torch.uint8, pin_memory=True))
            return pinned_keep[0].numpy()

        def alloc_fresh(shape):
            # what malloc / mallocForeignPtrBytes hand a caller for a buffer this size: a fresh anonymous mapping, untouched pages,
            # no huge-page advice from the allocator (np.empty would add its own madvise(MADV_HUGEPAGE))
            import mmap
            m = mmap.mmap(-1, int(np.prod(shape)), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)  # (Python's default is MAP_SHARED: shmem)
            return np.frombuffer(m, dtype=np.uint8).reshape(shape)

        def alloc_library(shape):
            # the library's own result allocator (elx_result_alloc / elx_result_free, as INTEGRATION.md's Haskell stub uses them):
            # 2 MB aligned, huge-page advised, recycled on free -- faulted in by the first (untimed) step only
            import weakref
            n = int(np.prod(shape))
            L.lib().elx_result_alloc.restype = C.c_void_p
            ptr = L.lib().elx_result_alloc(C.c_int64(n))
            if not ptr:
                raise MemoryError("elx_result_alloc(%d)" % n)
            arr = np.ctypeslib.as_array((C.c_uint8 * n).from_address(ptr)).reshape(shape)
            weakref.finalize(arr, L.lib().elx_result_free, C.c_void_p(ptr))
            return arr

        pageable = e2e_variant(alloc_fresh, "fresh anonymous mapping per step (untouched pageable memory: first-touch faults inside the timed region)")
        pinned = e2e_variant(alloc_pinned, "page-locked, pre-faulted buffer reused across steps (round 1's measurement)") if not args.no_e2e_pinned else None
        pinned_keep.clear()
        library_owned = e2e_variant(alloc_library, "elx_result_alloc + elx_result_free per step (the library's recycled, huge-page advised result buffers)")
        if rank == 0:
            L.lib().elx_result_pool_trim()
        packed_form = e2e_variant(alloc_fresh,
                                  "elx_simulate_packed into a fresh anonymous mapping: the caller keeps the %d-bit form (no expansion on the host)"
                                  % (2 if model.n_states <= 4 else 5), packed=True) if model.n_states <= 32 else None
        if rank == 0:
            h2d = int(model.weights.nbytes + model.pi.nbytes + model.exch.nbytes + tree.parent.nbytes + tree.brlen.nbytes + tree.leaf_row.nbytes) * world
            e2e = {"value": pageable["value_mean"], "unit": "cells/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": pageable["d2h_bytes_per_step"],
                   "statistic": "mean of %d steps (median alongside)" % k_e2e, "value_median": pageable["value_median"],
                   "ms_per_step": pageable["ms_mean"], "ms_steps": pageable["ms_steps"], "sites_total": total_sites, "n_devices_in_handle": world,
                   "result_bytes_per_step": int(T * total_sites), "caller_buffer": pageable["label"],
                   "transport": ("%d bits per cell over PCIe (k_pack%d), expanded into the caller's buffer by host threads"
                                 % ((2, 2) if model.n_states <= 4 else (5, 5)) if pageable["d2h_bytes_per_step"] < T * total_sites else "1 byte per cell over PCIe"),
                   "includes": "elx_create (eigensystems, H2D of model+tree, K1 tables; per device) + simulation + D2H + the caller's [T][S] uint8 host "
                               "buffer completely written + elx_destroy; ONE call on ONE handle driving %d device(s)" % world,
                   "caller_buffer_advice": "madvise(MADV_HUGEPAGE) on the caller's buffer by the library (first-touch faults of 2 MB; "
                                           "tools/bench_prefault.py: 10 GB in 0.16 s instead of 0.50 s on this host class)",
                   "pinned_prefaulted": pinned, "library_result_buffer": library_owned, "packed_result_form": packed_form}

    traffic = None
    try:
        tj = {}
        for fn in ("r1_traffic.json", "r2_traffic.json"):
            p = os.path.join(ROOT, "profiles", fn)
            if os.path.exists(p):
                tj.update(json.load(open(p)))
        traffic = tj.get("%s:%d" % (args.config, sites), {}).get("traffic_bytes")
        for name, c in configs.items():
            if "roofline" in c:
                c["roofline"]["traffic"] = tj.get("%s:%d" % (name, c["sites_per_gpu"]), {}).get("traffic_bytes")
    except Exception:
        pass
    clocks = sampler.stop()
    clocks["window"] = "whole run; per-configuration windows (warm-up + timed steps + post-roll of identical steps) under configs.*.clocks"
    clocks["headline_window"] = head["clocks"]
    cfg = config_block(desc, tree, model, sites, world)
    line = {
        "metric": "alignment_cells_per_second", "value": head["value"], "unit": "cells/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": cfg,
        "setup_ms_once_per_handle": round(head["setup_ms"], 2),  # elx_create (K1 tables), outside the timed region; reported separately
        "kernel_info": {"kernel": head["kernel"], "path": "generic" if args.force_generic else "auto",
                        "draws": draws_text(head), "philox_rounds": args.philox_rounds or 10},
        "roofline": {"bound": "hbm", "achieved": head["achieved"], "peak": peak, "unit": "GB/s", "frac": head["frac"], "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": T * sites, "kernel_ms": head["kernel_ms"], "kernel": head["kernel"],
                     "note": "1 byte per alignment cell (uint8 leaf state written once); CUDA events on the handle's stream; "
                             "traffic = DRAM read+write bytes per launch from the committed ncu --set full capture of this workload "
                             "(profiles/r*_traffic.json), null when none was taken"},
        "clocks": clocks,
        "gpu_launches": head["launches"],
        "wall_ms_per_step": head["wall_ms"],
        "configs": configs,
    }
    if e2e:
        line["e2e"] = e2e
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        _, otree, omodel, _ = build_inputs_oracle(args.config)
        v, s, dt = cpu_reference_rate(otree, omodel, target_seconds=args.cpu_seconds, cores=cores)
        line["cpu_baseline"] = {"value": v, "unit": "cells/s", "cores": cores, "kind": "port",
                                "sample": "%d taxa x %d sites (bounded sample, %.1f s), tables prebuilt, getChunks over %d threads" % (T, s, dt, cores)}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--sites", type=int, default=0, help="override sites per GPU (parity/profiling runs only; drops the configs block)")
    ap.add_argument("--e2e-sites", type=int, default=10_000_000)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-pinned", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="headline configuration only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--force-generic", action="store_true")
    ap.add_argument("--single-draws", action="store_true", help="K <= 4 models: node-by-node draws instead of joint sibling pairs (A/B runs)")
    ap.add_argument("--pair-draws", action="store_true", help="K <= 4 models: joint sibling pairs instead of quad mode (A/B runs)")
    ap.add_argument("--philox-rounds", type=int, default=0, choices=[0, 7, 10], help="0 = library default (10)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "product" and not args.sites:
        args.warmup = 3  # timing rules: W >= 3
    if args.impl == "reference":
        return run_reference(args)
    return run_product(args)


if __name__ == "__main__":
    sys.exit(main())
