"""CPU-side checks of the boundary: the C-ABI library loads, exports every declared symbol, validates its
inputs the way the reference does (`error` -> status + message), and never silently falls back to the CPU."""
import os
import re

import numpy as np
import pytest

import elynx_b200 as eb
from elynx_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    from elynx_b200 import build

    build.build()


def declared_symbols(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(elx_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = L.lib()
    syms = declared_symbols("elynx_b200.h")
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), s
        assert s in L.SIGNATURES, "no ctypes signature for " + s
    assert lib.elx_version() >= 100
    from elynx_b200 import _model_sigs

    msyms = declared_symbols("elynx_b200_model.h")
    assert len(msyms) >= 15
    for s in msyms:
        assert hasattr(lib, s), s
        assert s in _model_sigs.SIGNATURES, "no ctypes signature for " + s


class T3:
    parent = np.array([-1, 0, 0], dtype=np.int32)
    brlen = np.array([0.0, 1.0, 1.0])
    leaf_row = np.array([-1, 0, 1], dtype=np.int32)


JC = (np.full(4, 0.25), np.ones((4, 4)) - np.eye(4))


def expect_error(match, fn):
    with pytest.raises(eb.ElxError, match=match):
        fn()


def test_validation_matches_reference_errors():
    d, e = JC

    class Neg(T3):
        brlen = np.array([0.0, -1.0, 1.0])

    expect_error("Time is negative", lambda: eb.Simulator(None, d, e, Neg, is_mixture=False))  # MarkovProcess.hs:37
    expect_error("not square", lambda: eb.Simulator(None, d, np.ones((4, 3)), T3, is_mixture=False))  # MarkovProcess.hs:35

    class NotPre(T3):
        parent = np.array([-1, 2, 0], dtype=np.int32)

    expect_error("preorder", lambda: eb.Simulator(None, d, e, NotPre, is_mixture=False))

    class BadLeaf(T3):
        leaf_row = np.array([-1, 0, 0], dtype=np.int32)

    expect_error("permutation", lambda: eb.Simulator(None, d, e, BadLeaf, is_mixture=False))
    expect_error("Number of weights", lambda: eb.Simulator(np.ones(3), np.stack([d, d]), np.stack([e, e]), T3))  # MixtureModel.hs:86
    expect_error("bad weights", lambda: eb.Simulator(np.zeros(2), np.stack([d, d]), np.stack([e, e]), T3))
    big = np.ones((40, 40)) - np.eye(40)
    big[0, 1] = 2.0  # non-reversible models take the Pade path, which stops at 32 states
    expect_error("at most 32 states", lambda: eb.Simulator(None, np.full(40, 1 / 40), big, T3, is_mixture=False))


def test_no_cpu_fallback_without_a_device():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    d, e = JC
    with pytest.raises(eb.ElxError) as ei:
        eb.simulateAndFlattenPar(10, d, e, T3, 0)
    assert ei.value.code == L.ELX_E_CUDA and "no CPU fallback" in str(ei.value)


def _walk_info(parent, leaf_row):
    import ctypes as C

    parent = np.ascontiguousarray(parent, dtype=np.int32)
    leaf_row = np.ascontiguousarray(leaf_row, dtype=np.int32)
    brlen = np.zeros(len(parent))
    t = L.elx_tree(len(parent), parent.ctypes.data_as(C.POINTER(C.c_int32)), brlen.ctypes.data_as(C.POINTER(C.c_double)),
                   leaf_row.ctypes.data_as(C.POINTER(C.c_int32)))
    nl, ns = C.c_int32(), C.c_int32()
    order = np.empty(len(parent), dtype=np.int32)
    L.check(L.lib().elx_tree_walk_info(C.byref(t), C.byref(nl), C.byref(ns), C.c_void_p(order.ctypes.data)))
    return nl.value, ns.value, order


def test_walk_program_slot_bound():
    """Heavy-child-last walk: parents before children, every node once, and at most floor(log2 N)+1 live slots for ANY
    tree -- ladders of 10^4 taxa included (no HBM scratch needed for deep trees)."""
    import math

    from elynx_b200 import model as pm

    def ladder(n, left):  # caterpillar with the internal child first (left) or last (right), in preorder
        m = n - 1  # internal nodes
        if left:
            parent = [-1] + list(range(m - 1)) + [m - 1, m - 1] + list(range(m - 2, -1, -1))
            leaf_row = [-1] * m + list(range(n))
        else:
            parent, leaf_row = [-1], [-1]
            cur = 0
            for i in range(m - 1):
                parent += [cur, cur]
                leaf_row += [i, -1]
                cur = len(parent) - 1
            parent += [cur, cur]
            leaf_row += [m - 1, m]
        return np.array(parent), np.array(leaf_row)

    cases = [ladder(10000, True), ladder(10000, False)]
    for n, seed in ((1000, 1), (10000, 3), (3, 0), (2, 0)):
        t = pm.birth_death_tree(n, 1.0, 0.5, seed=seed)
        cases.append((t.parent, t.leaf_row))
    cases.append((np.array([-1]), np.array([0])))  # single node
    for parent, leaf_row in cases:
        n = len(parent)
        nl, ns, order = _walk_info(parent, leaf_row)
        assert nl == int((np.asarray(leaf_row) >= 0).sum())
        assert sorted(order.tolist()) == list(range(n))
        pos = np.empty(n, dtype=np.int64)
        pos[order] = np.arange(n)
        assert all(pos[parent[i]] < pos[i] for i in range(1, n))
        assert ns <= int(math.floor(math.log2(n))) + 1 if n > 1 else ns == 1, (n, ns)


def test_pair_walk_program():
    """Pair walk (K <= 4 models): every node in exactly one record, members are siblings, a record's parent is drawn
    earlier, src = -1 only right after the record that drew the parent as member A, slots are written before they are
    read and never clobbered while live, and the live-slot count stays logarithmic for binary trees."""
    import ctypes as C
    import math

    from elynx_b200 import model as pm

    rng = np.random.default_rng(5)
    cases = []
    for n, seed in ((1000, 1), (10000, 3), (3, 0), (2, 0), (257, 9)):
        t = pm.birth_death_tree(n, 1.0, 0.5, seed=seed)
        cases.append((np.asarray(t.parent), np.asarray(t.leaf_row), True))
    cases.append((np.array([-1]), np.array([0]), True))
    # random multifurcating / unary trees in preorder
    for _ in range(20):
        parent, kids = [-1], {0: 0}
        n = int(rng.integers(2, 200))
        stack = [0]
        while len(parent) < n:
            p = stack[-1]
            if rng.random() < 0.35 and len(stack) > 1:
                stack.pop()
                continue
            parent.append(p)
            kids[p] = kids.get(p, 0) + 1
            stack.append(len(parent) - 1)
        parent = np.array(parent)
        is_leaf = np.ones(len(parent), dtype=bool)
        is_leaf[parent[1:]] = False
        leaf_row = np.where(is_leaf, np.cumsum(is_leaf) - 1, -1)
        cases.append((parent, leaf_row, False))
    for parent, leaf_row, binary in cases:
        n = len(parent)
        parent = np.ascontiguousarray(parent, dtype=np.int32)
        leaf_row = np.ascontiguousarray(leaf_row, dtype=np.int32)
        brlen = np.zeros(n)
        t = L.elx_tree(n, parent.ctypes.data_as(C.POINTER(C.c_int32)), brlen.ctypes.data_as(C.POINTER(C.c_double)),
                       leaf_row.ctypes.data_as(C.POINTER(C.c_int32)))
        ng, ns = C.c_int32(), C.c_int32()
        recs = np.full((n, 7), -99, dtype=np.int32)
        L.check(L.lib().elx_tree_pair_walk_info(C.byref(t), C.byref(ng), C.byref(ns), C.c_void_p(recs.ctypes.data)))
        recs = recs[:ng.value]
        nchild = np.bincount(parent[1:], minlength=n)
        drawn = np.zeros(n, dtype=bool)
        slot_owner = {}  # slot -> node whose states it holds
        reads_left = {}  # node -> records that still have to read its slot
        prev_a = None
        for g, (a, b, src, dst_a, dst_b, leaf_a, leaf_b) in enumerate(recs.tolist()):
            assert not drawn[a] and (b < 0 or (not drawn[b] and parent[b] == parent[a]))
            par = parent[a]
            if g == 0:
                assert a == 0 and b == -1 and src == -1
            elif src == -1:
                assert par == prev_a, "registers hold the previous record's member A"
            else:
                assert slot_owner.get(src) == par, "slot must still hold the parent's states"
            if par >= 0:
                assert drawn[par]
                reads_left[par] -= 1
                if reads_left[par] == 0:
                    for s_, o in list(slot_owner.items()):
                        if o == par:
                            del slot_owner[s_]
            for node, dst, leaf in ((a, dst_a, leaf_a), (b, dst_b, leaf_b)):
                if node < 0:
                    assert dst == -1 and leaf == -1
                    continue
                drawn[node] = True
                assert leaf == (leaf_row[node] if nchild[node] == 0 else -1)
                groups = (nchild[node] + 1) // 2
                reads_left[node] = groups
                if dst >= 0:
                    assert nchild[node] > 0 and dst not in slot_owner, "would clobber a live slot"
                    slot_owner[dst] = node
                    assert dst < ns.value
            if nchild[a] > 0:
                assert dst_a >= 0 or (nchild[a] + 1) // 2 == 1
            if b >= 0 and nchild[b] > 0:
                assert dst_b >= 0
            prev_a = a if nchild[a] > 0 else None
        assert drawn.all() and not slot_owner
        if binary and n > 1:
            assert ng.value == (n - 1) // 2 + 1
            assert ns.value <= int(math.floor(math.log2(n))) + 1, (n, ns.value)


def test_quad_walk_program():
    """Quad walk (K <= 4 models, leaves only): the tree is cut into caps of at most 4 frontier nodes; helpers.check_quad_walk
    verifies the structure (every non-root node in exactly one cap, summed-out nodes keep all their children in the cap,
    roots are known when read, slots never clobbered while live).  Binary trees need between T/4 and T/2 records."""
    import ctypes as C

    from helpers import check_quad_walk
    from elynx_b200 import model as pm

    rng = np.random.default_rng(6)
    cases = []
    for n, seed in ((1000, 1), (10000, 3), (3, 0), (2, 0), (257, 9)):
        t = pm.birth_death_tree(n, 1.0, 0.5, seed=seed)
        cases.append((np.asarray(t.parent), np.asarray(t.leaf_row), True))
    cases.append((np.array([-1]), np.array([0]), True))
    # ladder (caterpillar) with 3000 leaves, a star with 1000 leaves and a unary chain of 40 nodes above a cherry
    parent, leaf_row, cur = [-1], [-1], 0
    for i in range(2999):
        parent += [cur, cur]
        leaf_row += [i, -1 if i < 2998 else 2999]
        cur = len(parent) - 1
    cases.append((np.array(parent), np.array(leaf_row), True))
    cases.append((np.array([-1] + [0] * 1000), np.array([-1] + list(range(1000))), False))
    cases.append((np.array([-1] + [0] * 30001), np.array([-1] + list(range(30001))), False))  # a wide star: the compiler must stay linear
    cases.append((np.array([-1] + list(range(40)) + [40, 40]), np.array([-1] * 41 + [0, 1]), False))
    for _ in range(30):
        parent = [-1]
        n = int(rng.integers(2, 300))
        stack = [0]
        while len(parent) < n:
            p = stack[-1]
            if rng.random() < 0.35 and len(stack) > 1:
                stack.pop()
                continue
            parent.append(p)
            stack.append(len(parent) - 1)
        parent = np.array(parent)
        is_leaf = np.ones(len(parent), dtype=bool)
        is_leaf[parent[1:]] = False
        cases.append((parent, np.where(is_leaf, np.cumsum(is_leaf) - 1, -1), False))
    for parent, leaf_row, binary in cases:
        n = len(parent)
        parent = np.ascontiguousarray(parent, dtype=np.int32)
        leaf_row = np.ascontiguousarray(leaf_row, dtype=np.int32)
        brlen = np.zeros(n)
        t = L.elx_tree(n, parent.ctypes.data_as(C.POINTER(C.c_int32)), brlen.ctypes.data_as(C.POINTER(C.c_double)),
                       leaf_row.ctypes.data_as(C.POINTER(C.c_int32)))
        ng, ns = C.c_int32(), C.c_int32()
        L.check(L.lib().elx_tree_quad_walk_info(C.byref(t), C.byref(ng), C.byref(ns), None))
        recs = np.full((max(ng.value, 1), L.ELX_QUAD_REC_INTS), -99, dtype=np.int32)
        L.check(L.lib().elx_tree_quad_walk_info(C.byref(t), C.byref(ng), C.byref(ns), C.c_void_p(recs.ctypes.data)))
        recs = recs[:ng.value]
        check_quad_walk(recs, parent, leaf_row, ns.value)
        nleaves = int((leaf_row >= 0).sum())
        if n > 1:
            assert (nleaves + 3) // 4 <= ng.value
            assert ns.value <= 24
        if binary and n > 3:
            assert ng.value <= (nleaves + 1) // 2, (nleaves, ng.value)


def test_parallel_gzip_writer(tmp_path):
    """elx_gzip_write (writeGZFile, elynx-tools InputOutput.hs:80-83): multi-member gzip that every gzip reader returns
    byte for byte; independent of the thread count; empty input gives one empty member."""
    import ctypes as C
    import gzip
    import subprocess

    rng = np.random.default_rng(3)
    data = (b">seq%d\n" % 7) + bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=3_000_001).tobytes()) + b"\n"
    outs = []
    for threads, block in ((1, 1 << 20), (4, 1 << 20), (3, 777_777)):
        path = str(tmp_path / ("t%d_%d.gz" % (threads, block)))
        buf = (C.c_char * len(data)).from_buffer_copy(data)
        L.check(L.lib().elx_gzip_write(path.encode(), C.cast(buf, C.c_void_p), len(data), threads, 6, block))
        raw = open(path, "rb").read()
        assert gzip.decompress(raw) == data
        assert raw.count(b"\x1f\x8b\x08") >= (len(data) + block - 1) // block
        assert subprocess.run(["gzip", "-dc", path], capture_output=True, check=True).stdout == data
        outs.append(raw)
    assert outs[0] == outs[1]                      # block size fixes the bytes, the thread count does not
    empty = str(tmp_path / "empty.gz")
    L.check(L.lib().elx_gzip_write(empty.encode(), None, 0, 2, 6, 0))
    assert gzip.decompress(open(empty, "rb").read()) == b""
    from elynx_b200 import cli
    p2 = str(tmp_path / "cli.gz")
    cli.write_gz(p2, data)
    assert gzip.decompress(open(p2, "rb").read()) == data
    with pytest.raises(L.ElxError):
        L.check(L.lib().elx_gzip_write(str(tmp_path / "nodir" / "x.gz").encode(), None, 0, 1, 6, 0))


def test_unpack2_rows_all_alignments():
    """elx_unpack2_rows -- the host half of the 2-bit transport of elx_simulate: every destination alignment, ragged lengths,
    state codes and the 4-character table, any thread count; bytes outside [0, nsites) of a row stay untouched."""
    import ctypes as C

    rng = np.random.default_rng(11)
    for rows, n, off, threads, lut in ((1, 0, 0, 1, None), (3, 1, 0, 1, None), (5, 7, 3, 2, b"ACGT"), (4, 257, 4, 3, None),
                                       (7, 1000, 28, 4, b"ACGT"), (6, 4099, 0, 2, None), (3, 70001, 13, 8, b"TGCA"),
                                       (33, 513, 32, 0, None)):
        states = rng.integers(0, 4, size=(rows, n), dtype=np.uint8)
        ppitch = (n + 3) // 4 + 5
        packed = rng.integers(0, 256, size=(rows, ppitch), dtype=np.uint8)  # garbage in the padding
        padded = np.zeros((rows, (n + 3) // 4 * 4), dtype=np.uint8)
        padded[:, :n] = states
        padded[:, n:] = rng.integers(0, 4, size=(rows, padded.shape[1] - n), dtype=np.uint8)  # pad sites of the last byte
        q = padded.reshape(rows, -1, 4)
        packed[:, :(n + 3) // 4] = q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4) | (q[:, :, 3] << 6)
        dpitch = n + off + 9
        raw = np.full(rows * dpitch + 64, 0xEE, dtype=np.uint8)
        base = (-raw.ctypes.data) % 32 + off  # destination of row 0: `off` bytes past a 32-byte boundary
        L.check(L.lib().elx_unpack2_rows(C.c_void_p(packed.ctypes.data), ppitch, C.c_void_p(raw.ctypes.data + base), dpitch, rows, n, lut, threads))
        want = np.full_like(raw, 0xEE)
        ref = states if lut is None else np.frombuffer(lut, dtype=np.uint8)[states]
        for r in range(rows):
            want[base + r * dpitch: base + r * dpitch + n] = ref[r]
        assert np.array_equal(raw, want), (rows, n, off, threads)
    with pytest.raises(L.ElxError):
        L.check(L.lib().elx_unpack2_rows(C.c_void_p(packed.ctypes.data), 1, C.c_void_p(raw.ctypes.data), 100, 2, 100, None, 1))


def test_unpack5_rows_all_alignments():
    """elx_unpack5_rows -- the 5-bit transport of amino-acid alignments (8 sites per 5 bytes): ragged lengths, every
    destination alignment, state codes and characters, bytes around a row untouched, no read past the packed row."""
    import ctypes as C

    rng = np.random.default_rng(12)
    aa = b"ACDEFGHIKLMNPQRSTVWY"
    for rows, n, off, threads, lut in ((1, 0, 0, 1, None), (2, 1, 0, 1, None), (3, 9, 5, 2, aa), (4, 63, 0, 2, None), (4, 64, 32, 1, aa),
                                       (5, 65, 1, 3, None), (3, 1000, 28, 4, aa), (2, 70001, 13, 8, None), (17, 513, 0, 0, aa)):
        k = 20 if lut else 32
        states = rng.integers(0, k, size=(rows, n), dtype=np.uint8)
        nbytes = (5 * n + 7) // 8
        bits = np.zeros((rows, nbytes * 8), dtype=np.uint8)
        for b in range(5):
            bits[:, b:5 * n:5] = (states >> b) & 1
        exact = np.packbits(bits, axis=1, bitorder="little")
        # rows packed back to back in an exactly sized allocation: any read past a row's bytes would leave the buffer
        packed = np.ascontiguousarray(exact)
        dpitch = n + off + 9
        raw = np.full(rows * dpitch + 64, 0xEE, dtype=np.uint8)
        base = (-raw.ctypes.data) % 32 + off
        L.check(L.lib().elx_unpack5_rows(C.c_void_p(packed.ctypes.data), nbytes, C.c_void_p(raw.ctypes.data + base), dpitch, rows, n, lut, threads))
        want = np.full_like(raw, 0xEE)
        ref = states if lut is None else np.frombuffer(lut, dtype=np.uint8)[states]
        for r in range(rows):
            want[base + r * dpitch: base + r * dpitch + n] = ref[r]
        assert np.array_equal(raw, want), (rows, n, off, threads)
    with pytest.raises(L.ElxError):
        L.check(L.lib().elx_unpack5_rows(C.c_void_p(packed.ctypes.data), 1, C.c_void_p(raw.ctypes.data), 100, 2, 100, None, 1))


def test_shard_range_follows_get_chunks():
    """elx_shard_range = getChunks (MarkovProcessAlongTree.hs:210-215) with aligned inner boundaries: tiles [0, n) exactly and
    agrees with the Python restatement used by the torchrun path (elynx_b200.distributed.site_range) and with the oracle."""
    import ctypes as C

    import oracle.model as om
    from elynx_b200 import _lib as L
    from elynx_b200.distributed import site_range

    lib = L.lib()
    for n in (0, 1, 63, 64, 65, 1000, 4096, 10_000_000, 10_000_019, (1 << 40) + 5):
        for shards in (1, 2, 3, 4, 7, 8):
            for align in (1, 16, 64):
                pos = 0
                for k in range(shards):
                    f, c = C.c_int64(), C.c_int64()
                    assert lib.elx_shard_range(n, shards, k, align, C.byref(f), C.byref(c)) == 0
                    assert (f.value, c.value) == site_range(k, shards, n, align=align)
                    assert f.value == pos and c.value >= 0
                    pos += c.value
                assert pos == n
            if n <= 10_000_019:
                sizes = []
                for k in range(shards):
                    f, c = C.c_int64(), C.c_int64()
                    lib.elx_shard_range(n, shards, k, 1, C.byref(f), C.byref(c))
                    sizes.append(c.value)
                assert sizes == om.get_chunks(shards, n)
    f, c = C.c_int64(), C.c_int64()
    assert lib.elx_shard_range(10, 0, 0, 1, C.byref(f), C.byref(c)) == L.ELX_E_INVALID
    assert lib.elx_shard_range(10, 2, 2, 1, C.byref(f), C.byref(c)) == L.ELX_E_INVALID


def test_result_pool_recycles_buffers():
    """elx_result_alloc / elx_result_free: 2 MB aligned host buffers, recycled for the next allocation of a similar size
    (the stand-in for the fresh allocation simulateAndFlatten*Par return their result in)."""
    import ctypes as C

    lib = L.lib()
    lib.elx_result_pool_trim()
    n = 5 << 20
    p = lib.elx_result_alloc(n)
    assert p and p % (2 << 20) == 0
    buf = (C.c_uint8 * n).from_address(p)
    buf[0] = 7
    buf[n - 1] = 9
    lib.elx_result_free(p)
    q = lib.elx_result_alloc(n - 4096)  # similar size: the same mapping comes back, contents and all
    assert q == p
    assert (C.c_uint8 * n).from_address(q)[0] == 7
    r = lib.elx_result_alloc(64 << 20)  # much larger: a new mapping
    assert r and r != q
    lib.elx_result_free(q)
    lib.elx_result_free(r)
    lib.elx_result_free(12345)  # not ours: ignored
    lib.elx_result_pool_trim()
    assert lib.elx_result_alloc(0) is None
