"""Times the validation histogram kernel (per-taxon state counts) on a simulated alignment that stays on the device:

  python tools/bench_hist.py [taxa] [sites] [model]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import elynx_b200 as eb  # noqa: E402
from elynx_b200 import model as pm  # noqa: E402


def main():
    taxa = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    sites = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
    spec = sys.argv[3] if len(sys.argv) > 3 else "HKY[6.0]{0.2,0.3,0.3,0.2}"
    tree = pm.birth_death_tree(taxa, 1.0, 0.5, seed=1)
    model = pm.get_phylo_model(spec, gamma=(4, 0.5))
    k = len(model.alphabet_chars)
    pitch = (sites + 127) // 128 * 128
    buf = torch.empty((taxa, pitch), dtype=torch.uint8, device="cuda")
    counts = torch.zeros((taxa, k), dtype=torch.int64, device="cuda")
    with eb.Simulator(model.weights, model.pi, model.exch, tree, is_mixture=True) as sim:
        stream = torch.cuda.ExternalStream(sim.stream)
        sim.simulate_device(1, 0, sites, buf.data_ptr(), pitch)
        for it in range(4):
            counts.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            sim.histogram_device(buf.data_ptr(), pitch, sites, counts.data_ptr())
            stream.synchronize()
            dt = time.perf_counter() - t0
            print("state counts %d x %d (K = %d): %.2f ms, %.2f TB/s read" % (taxa, sites, k, dt * 1e3, taxa * sites / dt / 1e12))
        got = counts.cpu().numpy()
        ref = np.stack([torch.bincount(buf[t, :sites].to(torch.int64), minlength=k).cpu().numpy() for t in range(0, taxa, max(1, taxa // 8))])
        assert np.array_equal(got[:: max(1, taxa // 8)], ref), "state counts differ from torch.bincount"
        print("checked against torch.bincount on %d rows" % ref.shape[0])


if __name__ == "__main__":
    main()
