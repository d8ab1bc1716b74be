"""Where the time of a host-buffer call goes: elx_create / the call / elx_destroy, for the expanded result (elx_simulate) and
the packed one (elx_simulate_packed), into a fresh pageable buffer, a touched pageable buffer and a pinned one.

  python tools/bench_e2e_phases.py [sites] [reps] [config]"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import elynx_b200 as eb  # noqa: E402
from elynx_b200 import _lib as L  # noqa: E402


def main():
    sites = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    desc, tree, model, _ = bench.build_inputs(sys.argv[3] if len(sys.argv) > 3 else "cfg2")
    T = tree.n_leaves
    ppitch = (sites + 31) // 32 * (8 if model.n_states <= 4 else 20)
    pinned = torch.empty((T, sites), dtype=torch.uint8, pin_memory=True).numpy()
    touched = np.zeros((T, sites), dtype=np.uint8)
    lib = L.lib()

    def fresh(shape):
        import mmap
        return np.frombuffer(mmap.mmap(-1, int(np.prod(shape)), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS), dtype=np.uint8).reshape(shape)

    def run(buf, packed):
        t0 = time.perf_counter()
        s2 = eb.Simulator(model.weights if model.is_mixture else None, model.pi, model.exch, tree, is_mixture=model.is_mixture)
        t1 = time.perf_counter()
        if packed:
            L.check(lib.elx_simulate_packed(s2.handle, C.c_uint64(0), 0, sites, C.c_void_p(buf.ctypes.data), ppitch, None), s2.handle)
        else:
            L.check(lib.elx_simulate(s2.handle, C.c_uint64(0), 0, sites, C.c_void_p(buf.ctypes.data), sites, None, None, 0), s2.handle)
        t2 = time.perf_counter()
        s2.close()
        t3 = time.perf_counter()
        return (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3

    for label, mk, packed in (("expanded / fresh pageable", lambda: fresh((T, sites)), False),
                              ("expanded / touched pageable", lambda: touched, False),
                              ("expanded / pinned", lambda: pinned, False),
                              ("packed / fresh pageable", lambda: fresh((T, ppitch)), True),
                              ("packed / touched pageable", lambda: touched.reshape(-1)[:T * ppitch].reshape(T, ppitch), True),
                              ("packed / pinned", lambda: pinned.reshape(-1)[:T * ppitch].reshape(T, ppitch), True)):
        if os.environ.get("PHASES_ONLY") and os.environ["PHASES_ONLY"] not in label:
            continue
        run(mk(), packed)
        for _ in range(reps):
            t0 = time.perf_counter()
            buf = mk()
            ta = (time.perf_counter() - t0) * 1e3
            c, s, d = run(buf, packed)
            print("%-30s alloc %6.1f  create %6.1f  call %6.1f  destroy %6.1f  = %6.1f ms" % (label, ta, c, s, d, ta + c + s + d), flush=True)
            del buf


if __name__ == "__main__":
    main()
