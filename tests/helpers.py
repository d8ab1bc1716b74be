"""Shared builders for the tests (oracle-side model/tree construction)."""
import os

import numpy as np

import oracle.model as om

REF_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_data")


def tree_from_file(name):
    return om.parse_newick(open(os.path.join(REF_DATA, name)).read())


def random_tree(n_leaves, rng, mean_len=0.3, multifurcate=False, zero_frac=0.0):
    """Random rooted tree as a Newick string -> FlatTree (exercises the parser and odd shapes)."""
    nodes = ["t%d:%r" % (i, float(rng.exponential(mean_len))) for i in range(n_leaves)]
    while len(nodes) > 1:
        k = 2 if not multifurcate or len(nodes) < 3 else int(rng.integers(2, min(4, len(nodes)) + 1))
        idx = sorted(rng.choice(len(nodes), size=k, replace=False).tolist(), reverse=True)
        kids = [nodes.pop(i) for i in idx]
        bl = 0.0 if rng.random() < zero_frac else float(rng.exponential(mean_len))
        nodes.append("(" + ",".join(kids) + "):%r" % bl)
    s = nodes[0]
    s = s[: s.rindex(":")] + ";"  # no root stem
    return om.parse_newick(s)


def model_case(name):
    """(is_mixture, ws, ds, es, alphabet) for a named model configuration."""
    edm = om.parse_edm_phylobayes(open(os.path.join(REF_DATA, "EDMDistsPhylobayes.txt")).read())
    if name == "jc":
        pm = om.get_phylo_model("JC", None)
    elif name == "hky":
        pm = om.get_phylo_model("HKY[6.0]{0.2,0.3,0.3,0.2}", None)
    elif name == "hky_g4":
        pm = om.expand(4, 0.5, om.get_phylo_model("HKY[6.0]{0.2,0.3,0.3,0.2}", None))
    elif name == "gtr4_g4":
        pm = om.expand(4, 0.5, om.get_phylo_model("GTR4[1,2,3,4,5,6]{0.2,0.3,0.4,0.1}", None))
    elif name == "lg":
        pm = om.get_phylo_model("LG", None)
    elif name == "lg_g4":
        pm = om.expand(4, 0.5, om.get_phylo_model("LG", None))
    elif name == "wag_g4":
        pm = om.expand(4, 0.5, om.get_phylo_model("WAG", None))
    elif name == "c10":
        pm = om.get_phylo_model(None, "C10", global_norm=True)
    elif name == "c10_g4":
        pm = om.expand(4, 0.2, om.get_phylo_model(None, "C10", global_norm=True))
    elif name == "c60":
        pm = om.get_phylo_model(None, "C60", global_norm=True)
    elif name == "edm_lg":
        pm = om.get_phylo_model(None, "EDM(LG-Custom)", edm_cs=edm)
    elif name == "mixture_dna":
        pm = om.get_phylo_model(None, "MIXTURE(JC,HKY[6.0]{0.25,0.25,0.25,0.25})", mws=[0.4, 0.6])
    else:
        raise KeyError(name)
    is_mix, ws, ds, es = om.flatten_model(pm)
    return is_mix, ws, ds, es, pm.alphabet


def uniforms(rng, rows, s):
    """doubles in (0,1] like uniformDoublePositive01M: w64/2^64 + 2^-65"""
    w = rng.integers(0, 2 ** 64, size=(rows, s), dtype=np.uint64)
    return w.astype(np.float64) / 18446744073709551616.0 + 2.0 ** -65


def spec_freerun(sim, model, tree, seed, site0, s, **kw):
    """The executable spec of the product's free-running draw for this handle, run on the handle's own tables:
    quad mode (K <= 4, leaves only: one draw per cap), duo mode (5 <= K <= 22, leaves only: caps of <= 2 frontier nodes), pair mode (K <= 4: sibling pairs drawn jointly) or node by node."""
    import oracle.freerun_spec as spec

    is_mix, ws, ds = model[0], model[1], model[2]
    quad = sim.quad_tables()
    if quad is not None and not kw.get("keep_internal"):
        kw.pop("keep_internal", None)
        return spec.freerun_quad(is_mix, tree.parent, tree.leaf_row, ws, ds, quad, seed, site0, s, **kw)
    duo = sim.duo_tables()
    if duo is not None and not kw.get("keep_internal"):
        kw.pop("keep_internal", None)
        return spec.freerun_duo(is_mix, tree.parent, tree.leaf_row, ws, ds, duo, seed, site0, s, **kw)
    pair = sim.pair_tables()
    if pair is not None:
        return spec.freerun_pair(is_mix, tree.parent, tree.leaf_row, ws, ds, pair, seed, site0, s, **kw)
    e16, qlo = sim.alias_tables()
    return spec.freerun(is_mix, tree.parent, tree.leaf_row, ws, ds, e16, qlo, seed, site0, s, **kw)


def alias_masses(e, qlo, a):
    """Integer outcome masses (sum 2^48 per row) encoded by alias rows e = [q_hi : 16-a][alias : a], qlo."""
    kp = 1 << a
    cap = 1 << (48 - a)
    e = e.astype(np.int64)
    al = e & (kp - 1)
    q = ((e >> a) << 32) | qlo.astype(np.int64)
    q = np.where(al == np.arange(kp), cap, q)  # alias == self means "always accept"
    mass = np.zeros(e.shape, dtype=np.int64)
    for j in range(kp):
        mass[..., j] += q[..., j]
        np.put_along_axis(mass, al[..., j:j + 1], np.take_along_axis(mass, al[..., j:j + 1], -1) + (cap - q[..., j:j + 1]), -1)
    return mass


def check_pair_tables(p, pair, parent):
    """Every non-root node is in exactly one record, members of a record are siblings, records are in walk order, and
    the joint rows carry integer masses (sum 2^48) equal to P_A[i][sa] * P_B[i][sb] (normalised) within 2^-44."""
    node_a, node_b, pe16, pqlo = pair
    n, c, k, _ = p.shape
    seen = np.zeros(n, dtype=bool)
    for g, (a, b) in enumerate(zip(node_a, node_b)):
        assert 0 <= a < n and not seen[a] and (parent[a] < 0 or seen[parent[a]])
        seen[a] = True
        if b >= 0:
            assert not seen[b] and parent[b] == parent[a]
            seen[b] = True
    assert seen.all()
    assert node_a[0] == 0 and node_b[0] == -1
    mass = alias_masses(pe16, pqlo, 4)
    assert np.all(mass.sum(-1) == 1 << 48)
    pn = p / p.sum(-1, keepdims=True)
    pa = pn[node_a]  # [G][C][K][K]
    pb = np.where((node_b >= 0)[:, None, None, None], pn[np.maximum(node_b, 0)], (np.arange(k) == 0).astype(float))
    joint = np.zeros(pe16.shape)
    for sa in range(k):
        for sb in range(k):
            joint[..., sa | (sb << 2)] = pa[..., sa] * pb[..., sb]
    assert np.max(np.abs(mass / float(1 << 48) - joint)) < 2.0 ** -44


def alias2_masses(e32, res32, tb, w):
    """Integer outcome masses (sum 2^w per row) encoded by two-level alias rows (elx_ptx.cuh): main entries e32 =
    [q : tb][alias : 24 - tb][0 : 8] on cells of 2^-24 (t = 1 .. 2^tb - 1; the kernels accept [t][j] < [q][alias]), escape entries
    res32 = [thr : 8 + tb][alias : 24 - tb] on the t == 0 draws (probability 2^-tb, column capacity 2^(w - 24))."""
    ab = 24 - tb
    kp = 1 << ab
    capm, capr, unit = (1 << tb) - 1, 1 << (w - 24), 1 << (w - 24)
    e = e32.astype(np.int64)
    assert np.all((e & 255) == 0)
    al = (e >> 8) & (kp - 1)
    q = e >> (32 - tb)
    cols = np.arange(kp)
    # t in 1..capm accepted: t < q, and t == q when j < alias
    acc = np.minimum(np.maximum(q - 1, 0), capm) + ((q >= 1) & (q <= capm) & (cols < al))
    assert np.all(acc <= capm)
    r = res32.astype(np.int64)
    ral, thr = r & (kp - 1), r >> ab
    assert np.all(thr < capr)
    shape = e.shape
    acc, al, thr, ral = (x.reshape(-1, kp) for x in (acc, al, thr, ral))
    mass = acc * unit + thr
    rows = np.arange(acc.shape[0])
    for j in range(kp):
        np.add.at(mass, (rows, al[:, j]), (capm - acc[:, j]) * unit)
        np.add.at(mass, (rows, ral[:, j]), capr - thr[:, j])
    return mass.reshape(shape)


def quad_masses(qe32, qres32):
    """Integer outcome masses (sum 2^48 per row) of quad rows: 16-bit t, 256 columns."""
    return alias2_masses(qe32, qres32, 16, 48)


def check_quad_walk(recs, parent, leaf_row, n_slots=None):
    """Structural validity of a quad walk: every leaf is a frontier node of exactly one record, every record hangs below a
    node whose state is known (the root or an earlier frontier node), cap nodes form a connected piece below it with
    parents listed first, inner nodes have ALL their children inside the cap, and slots are never read stale."""
    n = len(parent)
    kids = [[] for _ in range(n)]
    for v in range(1, n):
        kids[parent[v]].append(v)
    known = {0: 0}      # node -> slot holding its state
    live = {0: 0}       # slot -> node
    drawn = set()
    covered = set()     # nodes that appear in some cap (frontier or summed out)
    for r in recs:
        root, out, leaf, dst, src, ncap = int(r[0]), r[1:5], r[5:9], r[9:13], int(r[13]), int(r[14])
        assert root in known and known[root] == src and live.get(src) == root, "record reads a slot that does not hold its root"
        cap = [(int(r[16 + 3 * i]), int(r[17 + 3 * i]), int(r[18 + 3 * i])) for i in range(ncap)]
        nodes = [c[0] for c in cap]
        assert len(set(nodes)) == ncap
        fr_seen = {}
        for i, (node, par, fr) in enumerate(cap):
            assert par < i
            assert parent[node] == (root if par < 0 else cap[par][0])
            assert node not in covered
            covered.add(node)
            if fr >= 0:
                assert out[fr] == node and fr not in fr_seen
                fr_seen[fr] = node
                assert leaf[fr] == leaf_row[node]
                assert (dst[fr] >= 0) == bool(kids[node])
            else:
                assert kids[node], "a leaf cannot be summed out"
                in_cap = [c[0] for c in cap if c[1] == i]
                assert sorted(in_cap) == sorted(kids[node]), "a summed-out node must have all its children in the cap"
        assert sorted(fr_seen) == list(range(len(fr_seen))) and all(out[k] == -1 for k in range(len(fr_seen), 4))
        # slots: src is read first, then the frontier slots are written; a slot may only be overwritten once every cap
        # below the node it holds has been emitted (this record included)
        for k, node in fr_seen.items():
            drawn.add(node)
            if dst[k] >= 0:
                old = live.get(int(dst[k]))
                if old is not None:
                    assert all(c in covered for c in kids[old]), "slot overwritten while its node still has caps to draw"
                    del known[old]
                live[int(dst[k])] = node
                known[node] = int(dst[k])
    assert covered == set(range(1, n)), "every non-root node is in exactly one cap"
    assert all(v in drawn for v in range(n) if leaf_row[v] >= 0 and v != 0)
    for v in range(n):
        if kids[v] and (v == 0 or v in drawn):
            assert all(c in covered for c in kids[v])
    if n_slots is not None:
        assert max([0] + [int(x) for r in recs for x in list(r[9:13]) + [r[13]]]) < max(n_slots, 1)


def quad_joint_reference(p, rec, k):
    """P(frontier states | root state) [C][K][256] of one quad record by brute-force summation over the inner nodes."""
    c = p.shape[1]
    ncap = int(rec[14])
    cap = [(int(rec[16 + 3 * i]), int(rec[17 + 3 * i]), int(rec[18 + 3 * i])) for i in range(ncap)]
    inner = [i for i, x in enumerate(cap) if x[2] < 0]
    out = np.zeros((c, k, 256))
    nfr = sum(1 for x in cap if x[2] >= 0)
    import itertools
    for o in range(4 ** nfr):
        fs = [(o >> (2 * f)) & 3 for f in range(nfr)]
        if any(s >= k for s in fs):
            continue
        for assign in itertools.product(range(k), repeat=len(inner)):
            st = {}
            for i, a in zip(inner, assign):
                st[i] = a
            for i, x in enumerate(cap):
                if x[2] >= 0:
                    st[i] = fs[x[2]]
            for root_state in range(k):
                pr = np.ones(c)
                for i, (node, par, fr) in enumerate(cap):
                    ps = root_state if par < 0 else st[par]
                    pr = pr * p[node, :, ps, st[i]]
                out[:, root_state, o] += pr
    return out


def check_quad_tables(p, quad, parent, leaf_row, max_records=40):
    """The walk is valid and the alias rows carry integer masses (sum 2^48) equal to the caps' joint probabilities."""
    recs, qe32, qres32 = quad
    check_quad_walk(recs, parent, leaf_row)
    k = p.shape[2]
    mass = quad_masses(qe32, qres32)
    assert np.all(mass[:, :, :k].sum(-1) == 1 << 48)
    pn = p / p.sum(-1, keepdims=True)
    step = max(1, len(recs) // max_records)
    for g in range(0, len(recs), step):
        ref = quad_joint_reference(pn, recs[g], k)
        ref = ref / ref.sum(-1, keepdims=True)
        got = mass[g, :, :k] / float(1 << 48)
        assert np.max(np.abs(got - ref)) < 2.0 ** -40, (g, np.max(np.abs(got - ref)))


def duo_masses(de32, dres32):
    """Integer outcome masses (sum 2^47 per row) of duo rows: 15-bit t, 512 columns."""
    return alias2_masses(de32, dres32, 15, 47)


def duo_joint_reference(p, rec):
    """P(frontier states | root state) [C][K][512] of one duo record (o = s0 + K s1) by matrix products over the cap:
    inner nodes are summed out children first."""
    n, c, k, _ = p.shape
    ncap = int(rec[14])
    cap = [(int(rec[16 + 3 * i]), int(rec[17 + 3 * i]), int(rec[18 + 3 * i])) for i in range(ncap)]
    nfr = sum(1 for x in cap if x[2] >= 0)
    out = np.zeros((c, k, 512))
    for s1 in range(k if nfr > 1 else 1):
        for s0 in range(k):
            fs = (s0, s1)
            like = [np.ones((c, k)) for _ in range(ncap)]  # L[i][comp][state of cap node i]
            top = np.ones((c, k))
            for i in range(ncap - 1, -1, -1):
                node, par, fr = cap[i]
                if fr >= 0:
                    msg = p[node][:, :, fs[fr]]  # [c][parent state]
                else:
                    msg = np.einsum("cst,ct->cs", p[node], like[i])
                if par >= 0:
                    like[par] = like[par] * msg
                else:
                    top = top * msg
            out[:, :, s0 + k * s1] = top
    return out


def check_duo_tables(p, duo, parent, leaf_row, max_records=12):
    """The walk is valid (caps of at most two frontier nodes) and the alias rows carry integer masses (sum 2^48) equal to the
    caps' joint probabilities."""
    recs, de32, dres32 = duo
    check_quad_walk(recs, parent, leaf_row)
    assert np.all(recs[:, 3:5] == -1), "a duo cap has at most two frontier nodes"
    k = p.shape[2]
    pn = p / p.sum(-1, keepdims=True)
    step = max(1, len(recs) // max_records)
    for g in range(0, len(recs), step):
        mass = duo_masses(de32[g], dres32[g])
        assert np.all(mass.sum(-1) == 1 << 47)
        ref = duo_joint_reference(pn, recs[g])
        ref = ref / ref.sum(-1, keepdims=True)
        got = mass / float(1 << 47)
        assert np.max(np.abs(got - ref)) < 2.0 ** -40, (g, np.max(np.abs(got - ref)))
