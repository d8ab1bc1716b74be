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
