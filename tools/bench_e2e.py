"""A/B of the host-buffer result path (elx_create + elx_simulate into a pinned [T][S] buffer + destroy) on cfg2:

  python tools/bench_e2e.py [sites] [reps] [config]     # prints every repetition: the host side of the box is noisy"""
import ctypes as C
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import elynx_b200 as eb  # noqa: E402
from elynx_b200 import _lib as L  # noqa: E402


def main():
    sites = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    desc, tree, model, _ = bench.build_inputs(sys.argv[3] if len(sys.argv) > 3 else "cfg2")
    T = tree.n_leaves
    host = torch.empty((T, sites), dtype=torch.uint8, pin_memory=True)

    def step():
        with eb.Simulator(model.weights if model.is_mixture else None, model.pi, model.exch, tree, is_mixture=model.is_mixture) as s2:
            L.check(L.lib().elx_simulate(s2.handle, C.c_uint64(0), 0, sites, C.c_void_p(host.data_ptr()), sites, None, None, 0), s2.handle)

    modes = [("plain", {"ELX_PACKED_D2H": "0"})] + [("packed/%s threads" % t, {"ELX_PACKED_D2H": "1", "ELX_HOST_THREADS": t}) for t in ("8", "11", "13", "14")]
    for rnd in range(2):
        for name, env in modes:
            for k in ("ELX_PACKED_D2H", "ELX_HOST_THREADS"):
                os.environ.pop(k, None)
            os.environ.update(env)
            step()
            ts = []
            for _ in range(reps):
                t0 = time.perf_counter()
                step()
                ts.append((time.perf_counter() - t0) * 1e3)
            print("%-20s %s ms   best %.3g cells/s" % (name, " ".join("%6.1f" % t for t in ts), T * sites / (min(ts) * 1e-3)), flush=True)


if __name__ == "__main__":
    main()
