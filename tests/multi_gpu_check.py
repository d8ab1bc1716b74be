"""Multi-GPU acceptance check, run under torchrun on a box with >= 2 GPUs (not collected by pytest):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py

Each rank simulates its contiguous site range on its own GPU; checks
  (1) the per-rank device histograms summed with ONE NCCL all-reduce equal the histogram of a single-GPU run (rank 0),
  (2) the sharded FASTA file (K3 slice mode into pinned host memory, one process per GPU) is byte-identical to the
      single-GPU FASTA."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import elynx_b200 as eb  # noqa: E402
from elynx_b200 import model as pm  # noqa: E402
from elynx_b200.distributed import all_reduce_histograms, simulate_fasta_to_file, site_range  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tree = pm.parse_newick("((a:0.1,b:0.3):0.05,(c:0.02,d:0.7):0.2);")
    model = pm.get_phylo_model("HKY[6.0]{0.2,0.3,0.3,0.2}", gamma=(4, 0.5))
    total, seed = 3_000_017, 123
    with eb.Simulator(model.weights, model.pi, model.exch, tree, is_mixture=True, device=local) as sim:
        stream = torch.cuda.ExternalStream(sim.stream)
        s0, n = site_range(rank, world, total, align=16)
        pitch = (n + 127) // 128 * 128
        buf = torch.empty((4, pitch), dtype=torch.uint8, device="cuda")
        sc = torch.zeros((4, 4), dtype=torch.int64, device="cuda")
        pc = torch.zeros(256, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        sim.simulate_device(seed, s0, n, buf.data_ptr(), pitch)
        sim.histogram_device(buf.data_ptr(), pitch, n, sc.data_ptr(), pc.data_ptr())
        stream.synchronize()
        all_reduce_histograms(sc, pc)  # NCCL, 2 KiB
        if rank == 0:
            full = torch.empty((4, (total + 127) // 128 * 128), dtype=torch.uint8, device="cuda")
            sc1 = torch.zeros((4, 4), dtype=torch.int64, device="cuda")
            pc1 = torch.zeros(256, dtype=torch.int64, device="cuda")
            torch.cuda.synchronize()
            sim.simulate_device(seed, 0, total, full.data_ptr(), full.shape[1])
            sim.histogram_device(full.data_ptr(), full.shape[1], total, sc1.data_ptr(), pc1.data_ptr())
            stream.synchronize()
            assert torch.equal(sc, sc1) and torch.equal(pc, pc1), "NCCL-reduced histograms differ from the single-GPU run"
            assert int(pc.sum()) == total
        names = tree.leaf_names()
        path = os.path.join(tempfile.gettempdir(), "elx_multi_gpu_check.fasta")
        simulate_fasta_to_file(sim, seed, 200_003, names, model.alphabet_chars, path, rank=rank, world=world, batch_sites=65536,
                               barrier=dist.barrier)
        if rank == 0:
            ref, _ = sim.simulate_fasta(seed, 200_003, names, model.alphabet_chars)
            assert open(path, "rb").read() == ref, "sharded FASTA differs from the single-GPU FASTA"
            print("multi_gpu_check ok: world=%d, histograms reduced over NCCL, sharded FASTA byte-identical" % world)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
