"""The multi-GPU acceptance check (tests/multi_gpu_check.py) under pytest: launched with torchrun on min(device_count, 4)
GPUs of the box, skipped below 2.  Covers the NCCL reduction of the validation histograms, the sharded FASTA writer and the
multi-device C ABI against single-GPU output on real devices."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    torch = pytest.importorskip("torch")
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def test_multi_gpu_check_under_torchrun():
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs on the box (found %d)" % n)
    world = min(n, 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "multi_gpu_check ok: world=%d" % world in r.stdout


@pytest.mark.parametrize("name", ["hky_g4", "lg_g4"])
def test_multi_device_handle_matches_single_device(name, monkeypatch):
    """elx_opts.n_devices: ONE handle drives every GPU of the box from inside elx_simulate / elx_simulate_fasta (one host
    thread + stream per device, site ranges by the getChunks rule); the bytes must equal the single-device result -- with
    the packed and the plain transport, with components and internal states, at a site offset."""
    import numpy as np

    import elynx_b200 as eb
    import oracle.model as om
    from helpers import model_case, tree_from_file

    n = _n_gpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs on the box (found %d)" % n)
    is_mix, ws, ds, es, alphabet = model_case(name)
    tree = tree_from_file("Newick.tree")
    names = tree.leaf_names()
    nsites = 3 * (1 << 20) + 4321
    with eb.Simulator(ws, ds, es, tree, is_mixture=is_mix, device=0) as one, \
            eb.Simulator(ws, ds, es, tree, is_mixture=is_mix, device=0, n_devices=-1) as many:
        assert one.n_devices == 1 and many.n_devices == n
        for packed in ("1", "0"):
            monkeypatch.setenv("ELX_PACKED_D2H", packed)
            a = one.simulate(5, nsites, site0=(1 << 33) + 7, want_comps=True, keep_internal=packed == "0")
            b0 = many.d2h_bytes
            b = many.simulate(5, nsites, site0=(1 << 33) + 7, want_comps=True, keep_internal=packed == "0")
            assert many.d2h_bytes > b0
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
            if a[2] is not None:
                assert np.array_equal(a[2], b[2])
            fa, ca = one.simulate_fasta(5, nsites, names, om.ALPHABET_CHARS[alphabet], want_comps=True)
            fb, cb = many.simulate_fasta(5, nsites, names, om.ALPHABET_CHARS[alphabet], want_comps=True)
            assert fa == fb and np.array_equal(ca, cb)
        small = many.simulate(5, 1000)[0]  # below the sharding threshold: the first device alone, same bytes
        assert np.array_equal(small, one.simulate(5, 1000)[0])


def test_n_devices_beyond_the_box_is_rejected():
    import elynx_b200 as eb
    from helpers import model_case, tree_from_file

    n = _n_gpus()
    if n < 1:
        pytest.skip("needs a GPU")
    is_mix, ws, ds, es, _ = model_case("jc")
    with pytest.raises(eb.ElxError, match="visible device"):
        eb.Simulator(None, ds, es, tree_from_file("Newick.tree"), is_mixture=False, n_devices=n + 1)
