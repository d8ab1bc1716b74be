"""Multi-GPU plumbing: one process per GPU (torchrun), contiguous site ranges per rank, no data-path collective.

Sites are independent (MarkovProcessAlongTree.hs:94-98, 202-208) and the reference already shards them over cores by
the `getChunks` rule (:210-215).  Here every rank owns one contiguous site range of the same alignment and draws with
counters keyed on the GLOBAL site index, so the N-rank result is byte-identical to the 1-rank result.  The only
collective is the sum of the small validation histograms (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def get_chunks(c: int, n: int) -> List[int]:
    """getChunks c n (MarkovProcessAlongTree.hs:210-215): r chunks of n'+1 sites followed by c-r chunks of n' sites."""
    if c < 1:
        raise ValueError("get_chunks: need at least one chunk")
    q, r = divmod(n, c)
    return [q + 1] * r + [q] * (c - r)


def site_range(rank: int, world: int, n_sites: int, align: int = 1) -> Tuple[int, int]:
    """(site0, nsites) of `rank` under the getChunks rule.  With align > 1 the chunk boundaries are rounded to
    multiples of `align` (the kernels store 16 sites per 128-bit word; aligned ranges keep every store vectorised).
    The union of all ranges is exactly [0, n_sites) for any alignment."""
    if not 0 <= rank < world:
        raise ValueError("site_range: rank out of range")
    sizes = get_chunks(world, n_sites)
    bounds = np.concatenate([[0], np.cumsum(sizes)])
    if align > 1:
        inner = (bounds[1:-1] + align - 1) // align * align
        bounds = np.concatenate([[0], np.minimum(inner, n_sites), [n_sites]])
    return int(bounds[rank]), int(bounds[rank + 1] - bounds[rank])


def all_reduce_histograms(state_counts, pattern_counts=None):
    """Sum per-rank validation histograms in place (torch tensors on the device of the process group's backend)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(state_counts, op=dist.ReduceOp.SUM)
        if pattern_counts is not None:
            dist.all_reduce(pattern_counts, op=dist.ReduceOp.SUM)
    return state_counts, pattern_counts


def simulate_sharded(sim, seed: int, n_sites: int, rank: int, world: int, want_comps: bool = False):
    """This rank's slice of the alignment: (site0, leaves[T][nsites_rank], comps | None)."""
    site0, n = site_range(rank, world, n_sites, align=16)
    leaves, comps, _ = sim.simulate(seed, n, site0=site0, want_comps=want_comps)
    return site0, leaves, comps


def simulate_fasta_to_file(sim, seed: int, total_sites: int, names, alphabet: str, path: str, rank: int = 0, world: int = 1,
                           batch_sites: int = 1 << 22, barrier=None):
    """Write the FASTA file of a `total_sites`-column alignment with `world` GPUs (one process each).

    Rank r owns the columns site_range(r, world, total_sites, align=16) of every sequence line.  Kernel K3 (slice
    mode) of every GPU packs the characters of its columns STRAIGHT into that rank's pinned host buffer over PCIe
    (no device-side file image, no host-side interleaving, no collective; the pack call only enqueues -- its alphabet travels
    in the kernel's parameter block -- so batch b+1 is simulated and packed while the host writes batch b); the host then writes
    each row segment at its byte offset in the file (one pwrite per row and batch, straight from the pinned buffer).  Rank 0 creates the file at its final size and writes
    the `>name` lines and the newlines.  `barrier` orders file creation before the other ranks open it
    (torch.distributed.barrier under torchrun)."""
    import os

    import torch

    size = sim.fasta_size(names, total_sites)
    t = sim.n_leaves
    row_off = [sim.fasta_row_offset(names, total_sites, i) for i in range(t)]
    if rank == 0:
        with open(path, "wb") as f:
            f.truncate(size)
            for i in range(t):
                f.seek(row_off[i] - len(names[i].encode()) - 2)
                f.write(b">" + names[i].encode() + b"\n")
                f.seek(row_off[i] + total_sites)
                f.write(b"\n")
    if barrier is not None:
        barrier()
    site0, n = site_range(rank, world, total_sites, align=16)
    if n > 0:
        nb = max(16, min(batch_sites, n))
        pitch = (nb + 127) // 128 * 128
        d_leaves = torch.empty((t, pitch), dtype=torch.uint8, device="cuda")
        host = [torch.empty((t, pitch), dtype=torch.uint8, pin_memory=True) for _ in range(2)]  # double-buffered pinned slices
        stream = torch.cuda.ExternalStream(sim.stream)
        torch.cuda.synchronize()
        fd = os.open(path, os.O_RDWR)
        try:
            pending = None
            done = 0
            b = 0
            while done < n or pending is not None:
                cur = None
                if done < n:
                    k = min(nb, n - done)
                    hb = host[b & 1]
                    sim.simulate_device(seed, site0 + done, k, d_leaves.data_ptr(), pitch)
                    sim.fasta_pack_slice_device(d_leaves.data_ptr(), pitch, k, alphabet, hb.data_ptr(), pitch)
                    ev = torch.cuda.Event()
                    ev.record(stream)
                    cur = (ev, hb, site0 + done, k)
                    done += k
                    b += 1
                if pending is not None:  # write batch b-1 while the GPU works on batch b
                    ev, hb, s0, k = pending
                    ev.synchronize()
                    arr = hb.numpy()
                    for i in range(t):
                        os.pwrite(fd, memoryview(arr[i, :k]), row_off[i] + s0)  # (a view of the pinned row: no copy)
                pending = cur
        finally:
            os.close(fd)
    if barrier is not None:
        barrier()
    return site0, n
