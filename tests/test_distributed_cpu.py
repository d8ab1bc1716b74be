"""N > 1 host logic on CPU: site-range partition (getChunks rule) and the histogram reduction over a world_size-2
gloo group.  The GPU-side property (any partition of the site range gives identical bytes) is covered by
tests/test_gpu_parity.py::test_freerun_partition_invariance."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from elynx_b200.distributed import get_chunks, site_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_get_chunks_matches_reference_rule():
    assert get_chunks(3, 10) == [4, 3, 3]  # MarkovProcessAlongTree.hs:210-215
    assert get_chunks(4, 2) == [1, 1, 0, 0]
    assert get_chunks(8, 100_000_000) == [12_500_000] * 8  # BASELINE config 3 at 8 GPUs


def test_site_ranges_partition_exactly():
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 15, 16, 17, 1000, 10_000_019):
            for align in (1, 16, 128):
                pos = 0
                for r in range(world):
                    s0, k = site_range(r, world, n, align)
                    assert s0 == pos and k >= 0
                    if align > 1 and 0 < r and s0 < n:
                        assert s0 % align == 0
                    pos += k
                assert pos == n


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from elynx_b200.distributed import site_range, all_reduce_histograms
    dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 100003
    rng = np.random.default_rng(0)
    aln = rng.integers(0, 4, size=(3, n), dtype=np.uint8)       # the same "alignment" on every rank
    s0, k = site_range(rank, world, n, align=16)
    mine = aln[:, s0:s0 + k]
    sc = torch.from_numpy(np.stack([np.bincount(mine[t], minlength=4) for t in range(3)]).astype(np.int64))
    idx = (mine[0].astype(np.int64) * 4 + mine[1]) * 4 + mine[2]
    pc = torch.from_numpy(np.bincount(idx, minlength=64).astype(np.int64))
    all_reduce_histograms(sc, pc)
    full = (aln[0].astype(np.int64) * 4 + aln[1]) * 4 + aln[2]
    assert np.array_equal(pc.numpy(), np.bincount(full, minlength=64))
    assert np.array_equal(sc.numpy(), np.stack([np.bincount(aln[t], minlength=4) for t in range(3)]))
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_histogram_allreduce_gloo_world2(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert "rank %d ok" % r in o
