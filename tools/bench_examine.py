"""Times the `slynx examine` kernels on a simulated alignment that stays on the device (run under ncu for per-kernel times):

  python tools/bench_examine.py [taxa] [sites]

Prints cells/s of the whole elx_examine_device call (column pass + pairwise Hamming + the small D2H of the results)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import elynx_b200 as eb  # noqa: E402
from elynx_b200 import examine as ex  # noqa: E402
from elynx_b200 import model as pm  # noqa: E402


def main():
    taxa = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    sites = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
    tree = pm.birth_death_tree(taxa, 1.0, 0.5, seed=1)
    model = pm.get_phylo_model("HKY[6.0]{0.2,0.3,0.3,0.2}", gamma=(4, 0.5))
    pitch = (sites + 127) // 128 * 128
    buf = torch.empty((taxa, pitch), dtype=torch.uint8, device="cuda")
    with eb.Simulator(model.weights, model.pi, model.exch, tree, is_mixture=True) as sim:
        sim.simulate_device(1, 0, sites, buf.data_ptr(), pitch)
        sim.synchronize()
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = ex.examine_device(buf.data_ptr(), pitch, taxa, sites, 4)
        dt = time.perf_counter() - t0
        print("examine %d x %d: %.1f ms, %.3g cells/s, %.3g pair-sites/s; constant columns %d, kEff(entropy) %.3f, Hamming mean %.3f"
              % (taxa, sites, dt * 1e3, taxa * sites / dt, taxa * (taxa - 1) / 2 * sites / dt, r["n_constant"], r["keff_entropy_mean"],
                 r["hamming_mean"]))


if __name__ == "__main__":
    main()
