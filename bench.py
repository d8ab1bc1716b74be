#!/usr/bin/env python3
"""Benchmark of the alignment-simulation hot path (BASELINE.json metric: alignment cells/s and HBM GB/s vs peak).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config cfg2|cfg1|cfg3|cfg4|cfg5]

A "step" is one pass of the hot path over one batch of synthetic input: the simulation of `sites` alignment
columns for every taxon of the tree (T x S uint8 leaf states written once to HBM, layout [taxon][site]).
At N GPUs every rank simulates its own contiguous site range of an (N x S)-column alignment (weak scaling;
no data-path collective: sites are independent and RNG counters use the global site index).

Workload at N=1 (config.workload): BASELINE.json configs[1] -- HKY + Gamma(4) DNA, 1000-taxon birth-death tree
(tlynx defaults lambda=1, mu=0.5, tree seed 1), 10M sites on one B200.

JSON keys follow the driver contract; in addition
  roofline      the simulation kernel against the measured HBM bandwidth (1 algorithmic byte per cell)
  cpu_baseline  the oracle port of the reference algorithm timed on the box's host cores (rank 0, N=1)
  e2e           the same metric through the reference-facing call (host buffers, D2H inside the timed region)
"""
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

CONFIGS = {
    # name: (workload description, taxa, tree seed, model kwargs, sites per GPU)
    "cfg1": ("slynx simulate JC DNA, 10k sites, bundled 40-taxon Newick.tree", None, None, dict(substitution_model="JC"), 10_000),
    "cfg2": ("HKY[6.0]{0.2,0.3,0.3,0.2} + Gamma(4,0.5) DNA, 1000-taxon birth-death tree (lambda=1, mu=0.5, seed 1), 10M sites",
             1000, 1, dict(substitution_model="HKY[6.0]{0.2,0.3,0.3,0.2}", gamma=(4, 0.5)), 10_000_000),
    "cfg2gtr": ("GTR4[1,2,3,4,5,6]{0.2,0.3,0.4,0.1} + Gamma(4,0.5) DNA, 1000-taxon birth-death tree (seed 1), 10M sites",
                1000, 1, dict(substitution_model="GTR4[1,2,3,4,5,6]{0.2,0.3,0.4,0.1}", gamma=(4, 0.5)), 10_000_000),
    "cfg3": ("LG + Gamma(4,0.5) protein, 1000-taxon birth-death tree (seed 1), 12.5M sites per GPU", 1000, 1,
             dict(substitution_model="LG", gamma=(4, 0.5)), 12_500_000),
    "cfg4": ("C60 (-m C60 -n) profile mixture, 500-taxon birth-death tree (seed 2), 10M sites", 500, 2,
             dict(mixture_model="C60", global_normalization=True), 10_000_000),
    "cfg5": ("WAG + Gamma(4,0.5) protein, 10000-taxon birth-death tree (seed 3), 1.25M sites per GPU", 10000, 3,
             dict(substitution_model="WAG", gamma=(4, 0.5)), 1_250_000),
}


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

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
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


def build_inputs(cfg_name):
    from elynx_b200 import model as pm

    desc, taxa, tseed, mkw, sites = CONFIGS[cfg_name]
    if taxa is None:
        tree = pm.parse_newick(open(os.path.join(ROOT, "tests", "golden", "ref_data", "Newick.tree")).read())
    else:
        tree = pm.birth_death_tree(taxa, 1.0, 0.5, seed=tseed)
    model = pm.get_phylo_model(**mkw)
    return desc, tree, model, sites


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
    desc, tree, model, sites = build_inputs(args.config)
    cores = os.cpu_count() or 1
    value, s, dt = cpu_reference_rate(tree, model, target_seconds=10.0, cores=cores, steps=args.steps, warmup=min(args.warmup, 1))
    sample = "%d taxa x %d sites per step (bounded sample of the %d-site workload), tables prebuilt, %d threads" % (tree.n_leaves, s, sites, cores)
    line = {
        "impl": "reference", "metric": "alignment_cells_per_second", "value": value, "unit": "cells/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": desc, "taxa": tree.n_leaves, "nodes": tree.n_nodes, "sites_per_gpu": sites, "states": model.n_states,
                   "components": model.n_components},
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "C restatement of the reference algorithm (oracle/sim.c), not the Haskell binary (no GHC in the image)"},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    desc, tree, model, sites = build_inputs(args.config)
    if args.sites:
        sites = args.sites
    T = tree.n_leaves
    site0 = rank * sites  # contiguous site ranges of one (world x sites)-column alignment
    pitch = (sites + 127) // 128 * 128

    sim = eb.Simulator(model.weights if model.is_mixture else None, model.pi, model.exch, tree, is_mixture=model.is_mixture,
                       device=local_rank, force_generic=args.force_generic, philox_rounds=args.philox_rounds,
                       single_draws=args.single_draws, pair_draws=args.pair_draws)
    stream = torch.cuda.ExternalStream(sim.stream, device=torch.device("cuda", local_rank))
    d_leaves = torch.empty((T, pitch), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    seed = 0

    def step():
        sim.simulate_device(seed, site0, sites, d_leaves.data_ptr(), pitch)

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = sim.launch_count
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
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
    # the timed region (K x ~12 ms) is shorter than nvidia-smi's sampling period: keep the GPU under the identical load
    # for a short post-roll so that the clock / throttle-reason record has samples taken under this very kernel
    t_roll = time.perf_counter()
    while time.perf_counter() - t_roll < 0.6:
        step()
        torch.cuda.synchronize()
    clocks = sampler.stop()
    clocks["window"] = "warm-up + timed steps + 0.6 s post-roll of identical steps"
    kernel_ms = [a.elapsed_time(b) for a, b in evs]
    dev_total_ms = evs[0][0].elapsed_time(evs[-1][1])
    # max over ranks of the device time for the K steps
    tt = torch.tensor([dev_total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = float(tt.item())
    ms_per_step = total_ms / args.steps
    value = world * T * sites / (ms_per_step * 1e-3)

    # ---- end to end through the reference-facing call: host buffers, table build + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        e2e_sites = min(sites, args.e2e_sites)
        if world > 1:  # keep the pinned host buffers of all ranks of the box together under ~24 GB
            e2e_sites = max(16, min(e2e_sites, int(24e9 / (T * world)) // 16 * 16))
        if world > 1 and "ELX_HOST_THREADS" not in os.environ:  # the ranks of a box share its cores
            os.environ["ELX_HOST_THREADS"] = str(max(1, ((os.cpu_count() or 3) - 3) // world))
        host = torch.empty((T, e2e_sites), dtype=torch.uint8, pin_memory=True)
        hnp = host.numpy()
        import ctypes as C
        from elynx_b200 import _lib as L

        def e2e_step():
            s2 = eb.Simulator(model.weights if model.is_mixture else None, model.pi, model.exch, tree, is_mixture=model.is_mixture,
                              device=local_rank, force_generic=args.force_generic, philox_rounds=args.philox_rounds,
                              single_draws=args.single_draws, pair_draws=args.pair_draws)
            L.check(L.lib().elx_simulate(s2.handle, C.c_uint64(seed), site0, e2e_sites, C.c_void_p(hnp.ctypes.data), e2e_sites, None, None, 0),
                    s2.handle)
            n = s2.d2h_bytes
            s2.close()
            return n

        e2e_step()
        # every step is timed on its own (barrier on both sides, max over ranks) and the MEDIAN step is reported: the host
        # side of a shared box hiccups (one step in ~10 takes 2-4x as long, with either transport); all steps are listed
        k_e2e = max(1, min(args.steps, 5))
        step_ms = []
        for _ in range(k_e2e):
            barrier()
            t1 = time.perf_counter()
            d2h = e2e_step()
            barrier()
            tstep = torch.tensor([time.perf_counter() - t1], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tstep, op=dist.ReduceOp.MAX)
            step_ms.append(float(tstep.item()) * 1e3)
        te = torch.tensor([float(np.median(step_ms)) * 1e-3], dtype=torch.float64, device="cuda")
        h2d = int(model.weights.nbytes + model.pi.nbytes + model.exch.nbytes + tree.parent.nbytes + tree.brlen.nbytes + tree.leaf_row.nbytes)
        e2e = {"value": world * T * e2e_sites / float(te.item()), "unit": "cells/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": int(d2h), "sites_per_gpu": e2e_sites, "ms_per_step": float(te.item()) * 1e3,
               "result_bytes_per_step": int(T * e2e_sites), "steps": k_e2e, "ms_steps": [round(x, 2) for x in step_ms],
               "statistic": "median step",
               "transport": ("%d bits per cell over PCIe (k_pack%d), expanded into the caller's buffer by host threads"
                             % ((2, 2) if model.n_states <= 4 else (5, 5)) if d2h < T * e2e_sites else "1 byte per cell over PCIe"),
               "includes": "elx_create (eigensystems, H2D of model+tree, K1 tables) + simulation + D2H + the caller's [T][S] uint8 host "
                           "buffer completely written + destroy"}
        del host

    peak, peak_src = measured_peak_gbs()
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json"))).get("%s:%d" % (args.config, sites), {}).get("traffic_bytes")
    except Exception:
        pass
    kms = float(np.mean(kernel_ms))
    achieved = T * sites / (kms * 1e-3) / 1e9
    line = {
        "metric": "alignment_cells_per_second", "value": value, "unit": "cells/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": desc, "taxa": T, "nodes": tree.n_nodes, "sites_per_gpu": sites, "states": model.n_states,
                   "components": model.n_components, "parallelism": "site-range sharding x%d, no data-path collective" % world,
                   "l2": "output %d MB per step >> 126 MB L2, no flush needed" % (T * sites // 2 ** 20),
                   "kernel": "generic" if args.force_generic else "auto",
                   "draws": ("one draw per cap of <= 4 frontier nodes (inner nodes summed out), %d draws per site" % sim.quad_groups) if sim.quad_groups
                   else ("joint sibling pairs" if sim.pair_groups else "node by node"), "philox_rounds": args.philox_rounds or 10},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": T * sites, "kernel_ms": kms,
                     "note": "1 byte per alignment cell (uint8 leaf state written once); CUDA events on the handle's stream; "
                             "traffic = DRAM read+write bytes per launch from the committed ncu --set full capture of this workload "
                             "(profiles/r1_traffic.json), null when none was taken"},
        "clocks": clocks,
        "gpu_launches": int(launches),
        "wall_ms_per_step": wall / args.steps * 1e3,
    }
    if e2e:
        line["e2e"] = e2e
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, s, dt = cpu_reference_rate(tree, model, target_seconds=args.cpu_seconds, cores=cores)
        line["cpu_baseline"] = {"value": v, "unit": "cells/s", "cores": cores, "kind": "port",
                                "sample": "%d taxa x %d sites (bounded sample, %.1f s), tables prebuilt, getChunks over %d threads" % (T, s, dt, cores)}
    sim.close()
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
    ap.add_argument("--sites", type=int, default=0, help="override sites per GPU (parity/profiling runs only)")
    ap.add_argument("--e2e-sites", type=int, default=10_000_000)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-e2e", action="store_true")
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
