"""Times the FASTA packer K3 on a simulated alignment that stays on the device (run under ncu for the kernel time):

  python tools/bench_fasta.py [taxa] [sites]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import elynx_b200 as eb  # noqa: E402
from elynx_b200 import model as pm  # noqa: E402


def main():
    taxa = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    sites = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
    tree = pm.birth_death_tree(taxa, 1.0, 0.5, seed=1)
    model = pm.get_phylo_model("HKY[6.0]{0.2,0.3,0.3,0.2}", gamma=(4, 0.5))
    names = tree.leaf_names()
    pitch = (sites + 127) // 128 * 128
    buf = torch.empty((taxa, pitch), dtype=torch.uint8, device="cuda")
    with eb.Simulator(model.weights, model.pi, model.exch, tree, is_mixture=True) as sim:
        size = sim.fasta_size(names, sites)
        out = torch.empty(size, dtype=torch.uint8, device="cuda")
        stream = torch.cuda.ExternalStream(sim.stream)
        sim.simulate_device(1, 0, sites, buf.data_ptr(), pitch)
        for it in range(4):
            stream.synchronize()
            t0 = time.perf_counter()
            sim.fasta_pack_device(buf.data_ptr(), pitch, 0, sites, sites, names, model.alphabet_chars, out.data_ptr())
            stream.synchronize()
            dt = time.perf_counter() - t0
            print("fasta pack %d x %d: %.2f ms, %.2f TB/s (read + write)" % (taxa, sites, dt * 1e3, 2 * taxa * sites / dt / 1e12))
        head = bytes(out[:40].cpu().numpy())
        print(head)


if __name__ == "__main__":
    main()
