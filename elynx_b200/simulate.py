"""Host-side mirror of ELynx.Simulate.MarkovProcessAlongTree (elynx-markov/src/ELynx/Simulate/MarkovProcessAlongTree.hs).

Same function names (snake_case + the reference's camelCase aliases), same argument meaning and order,
same error behaviour (`error` -> ElxError).  The tree argument is the preorder flattening of the
reference's `Tree Double` (see include/elynx_b200.h): any object with .parent, .brlen, .leaf_row arrays.
The generator argument of the reference (`IOGenM g`) becomes a 64-bit seed: the wrapper would draw one
Word64 from the caller's generator."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib as L


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _as_tree_arrays(tree):
    parent = np.ascontiguousarray(tree.parent, dtype=np.int32)
    brlen = np.ascontiguousarray(tree.brlen, dtype=np.float64)
    leaf_row = np.ascontiguousarray(tree.leaf_row, dtype=np.int32)
    if not (parent.shape == brlen.shape == leaf_row.shape) or parent.ndim != 1:
        raise L.ElxError(L.ELX_E_INVALID, "tree: parent/brlen/leaf_row must be 1-d arrays of equal length")
    return parent, brlen, leaf_row


class Simulator:
    """One handle = one (model, tree) pair resident on one GPU: eigensystems, P(t) tables (kernel K1),
    alias tables and the compiled tree walk."""

    def __init__(self, weights, ds, es, tree, is_mixture: bool = True, device: int = -1, philox_rounds: int = 0,
                 force_generic: bool = False, pade: bool = False, single_draws: bool = False, pair_draws: bool = False,
                 n_devices: int = 1, eigen: bool = False, flags: int = 0):
        ds = np.ascontiguousarray(ds, dtype=np.float64)
        es = np.ascontiguousarray(es, dtype=np.float64)
        if ds.ndim == 1:
            ds = ds[None]
        if es.ndim == 2:
            es = es[None]
        if es.ndim != 3 or es.shape[1] != es.shape[2]:
            raise L.ElxError(L.ELX_E_INVALID, "probMatrix: Matrix is not square.")
        c, k = ds.shape
        if es.shape[0] != c or es.shape[1] != k:
            raise L.ElxError(L.ELX_E_INVALID, "model: stationary distributions and exchangeability matrices do not match")
        ws = np.ones(c) if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        if ws.shape != (c,):
            raise L.ElxError(L.ELX_E_INVALID, "fromSubstitutionModels: Number of weights and substitution models does not match.")
        parent, brlen, leaf_row = _as_tree_arrays(tree)
        self._keep = (ws, ds, es, parent, brlen, leaf_row)
        m = L.elx_model(k, c, int(bool(is_mixture)), ws.ctypes.data_as(C.POINTER(C.c_double)),
                        ds.ctypes.data_as(C.POINTER(C.c_double)), es.ctypes.data_as(C.POINTER(C.c_double)))
        t = L.elx_tree(len(parent), parent.ctypes.data_as(C.POINTER(C.c_int32)), brlen.ctypes.data_as(C.POINTER(C.c_double)),
                       leaf_row.ctypes.data_as(C.POINTER(C.c_int32)))
        o = L.elx_opts(device, philox_rounds, (L.ELX_FLAG_FORCE_GENERIC if force_generic else 0) | (L.ELX_FLAG_PADE if pade else 0) |
                       (L.ELX_FLAG_SINGLE_DRAWS if single_draws else 0) | (L.ELX_FLAG_PAIR_DRAWS if pair_draws else 0) | (L.ELX_FLAG_EIGEN if eigen else 0) | int(flags), n_devices)
        h = C.c_void_p()
        self._h = None
        L.check(L.lib().elx_create(C.byref(m), C.byref(t), C.byref(o), C.byref(h)))
        self._h = h
        self.n_states, self.n_components, self.is_mixture = k, c, bool(is_mixture)
        self.n_devices = int(L.lib().elx_n_devices(h))  # host-buffer calls shard the site range over this many GPUs
        self.n_nodes = int(L.lib().elx_n_nodes(h))
        self.n_leaves = int(L.lib().elx_n_leaves(h))
        self.alias_columns = int(L.lib().elx_alias_columns(h))
        self.pair_groups = int(L.lib().elx_pair_groups(h))  # > 0: K <= 4, sibling pairs are drawn jointly (elx_pair.cu)
        # > 0: K <= 4, leaves-only calls draw one cap of <= 4 frontier nodes at a time, inner nodes summed out (elx_quad.cu)
        self.quad_groups = int(L.lib().elx_quad_groups(h))
        # > 0: 5 <= K <= 22, leaves-only calls draw one cap of <= 2 frontier nodes at a time on component-bucketed sites (elx_duo.cu)
        self.duo_groups = int(L.lib().elx_duo_groups(h))

    # -- lifetime --------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            L.lib().elx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self):
        return self._h

    @property
    def stream(self) -> int:
        """cudaStream_t all kernels of this handle run on (as an integer)."""
        return int(L.lib().elx_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(L.lib().elx_launch_count(self._h))

    @property
    def last_kernel(self) -> str:
        """name of the simulation kernel the last device call launched (elx_last_kernel)"""
        return (L.lib().elx_last_kernel(self._h) or b"").decode()

    @property
    def d2h_bytes(self) -> int:
        """Bytes the host-buffer calls have copied device -> host so far (2 bits per cell for alphabets of <= 4 states)."""
        return int(L.lib().elx_d2h_bytes(self._h))

    # -- parity hooks ------------------------------------------------------------------------
    def ptables(self, cumulative: bool = False) -> np.ndarray:
        """[N][C][K][K] transition probabilities (probMatrix, MarkovProcess.hs:33-41) or their row-cumulative sums."""
        out = np.empty((self.n_nodes, self.n_components, self.n_states, self.n_states), dtype=np.float64)
        L.check(L.lib().elx_ptables_get(self._h, _ptr(out), int(cumulative)), self._h)
        return out

    def set_ptables(self, p):
        p = np.ascontiguousarray(p, dtype=np.float64)
        if p.shape != (self.n_nodes, self.n_components, self.n_states, self.n_states):
            raise L.ElxError(L.ELX_E_INVALID, "set_ptables: expected shape [N][C][K][K]")
        L.check(L.lib().elx_ptables_set(self._h, _ptr(p)), self._h)

    def alias_tables(self) -> Tuple[np.ndarray, np.ndarray]:
        shape = (self.n_nodes, self.n_components, self.n_states, self.alias_columns)
        e16 = np.empty(shape, dtype=np.uint16)
        qlo = np.empty(shape, dtype=np.uint32)
        L.check(L.lib().elx_alias_get(self._h, _ptr(e16), _ptr(qlo)), self._h)
        return e16, qlo

    def pair_tables(self):
        """(node_a[G], node_b[G], pe16[G][C][K][16], pqlo[G][C][K][16]) of the pair-mode records, or None."""
        g = self.pair_groups
        if g == 0:
            return None
        node_a = np.empty(g, dtype=np.int32)
        node_b = np.empty(g, dtype=np.int32)
        shape = (g, self.n_components, self.n_states, 16)
        pe16 = np.empty(shape, dtype=np.uint16)
        pqlo = np.empty(shape, dtype=np.uint32)
        L.check(L.lib().elx_pair_get(self._h, _ptr(node_a), _ptr(node_b), _ptr(pe16), _ptr(pqlo)), self._h)
        return node_a, node_b, pe16, pqlo

    def quad_tables(self):
        """(recs[G][ELX_QUAD_REC_INTS], qe32[G][C][4][256], qres32[G][C][4][256]) of the quad-mode records, or None."""
        g = self.quad_groups
        if g == 0:
            return None
        recs = np.empty((g, L.ELX_QUAD_REC_INTS), dtype=np.int32)
        shape = (g, self.n_components, 4, 256)
        qe32 = np.empty(shape, dtype=np.uint32)
        qres32 = np.empty(shape, dtype=np.uint32)
        L.check(L.lib().elx_quad_get(self._h, _ptr(recs), _ptr(qe32), _ptr(qres32)), self._h)
        return recs, qe32, qres32

    # -- simulation ----------------------------------------------------------------------------
    def duo_tables(self):
        """(recs[G][ELX_QUAD_REC_INTS], de32[G][C][K][512], dres32[G][C][K][512]) of the duo-mode records, or None."""
        g = self.duo_groups
        if g == 0:
            return None
        recs = np.empty((g, L.ELX_QUAD_REC_INTS), dtype=np.int32)
        de32 = np.empty((g, self.n_components, self.n_states, 512), dtype=np.uint32)
        dres32 = np.empty_like(de32)
        L.check(L.lib().elx_duo_get(self._h, _ptr(recs), _ptr(de32), _ptr(dres32)), self._h)
        return recs, de32, dres32

    def simulate(self, seed: int, nsites: int, site0: int = 0, want_comps: bool = False, keep_internal: bool = False):
        """(leaves[T][nsites] uint8, comps[nsites] int32 | None, internal[N][nsites] uint8 | None) in host memory."""
        nsites = int(nsites)
        if nsites < 0:
            raise L.ElxError(L.ELX_E_INVALID, "simulate: negative number of sites")
        leaves = np.empty((self.n_leaves, nsites), dtype=np.uint8)
        comps = np.empty(nsites, dtype=np.int32) if want_comps else None
        internal = np.empty((self.n_nodes, nsites), dtype=np.uint8) if keep_internal else None
        L.check(L.lib().elx_simulate(self._h, C.c_uint64(seed & (2 ** 64 - 1)), site0, nsites, _ptr(leaves), max(nsites, 1),
                                     _ptr(comps), _ptr(internal), max(nsites, 1)), self._h)
        return leaves, comps, internal

    def simulate_packed(self, seed: int, nsites: int, site0: int = 0, want_comps: bool = False):
        """The alignment in the packed form it crosses PCIe in: (packed [T][elx_packed_pitch] uint8, bits per site, comps)."""
        lib = L.lib()
        pitch = int(lib.elx_packed_pitch(self._h, nsites))
        packed = np.empty((self.n_leaves, max(pitch, 1)), dtype=np.uint8)
        comps = np.empty(nsites, dtype=np.int32) if want_comps else None
        L.check(lib.elx_simulate_packed(self._h, C.c_uint64(seed & (2 ** 64 - 1)), site0, nsites, _ptr(packed), max(pitch, 1), _ptr(comps)), self._h)
        return packed, int(lib.elx_packed_bits(self._h)), comps

    def site_components(self, seed: int, nsites: int, site0: int = 0) -> np.ndarray:
        """Component index of every site (the `cs` of simulateAndFlattenMixtureModelPar) without simulating the alignment."""
        comps = np.empty(nsites, dtype=np.int32)
        L.check(L.lib().elx_site_components(self._h, C.c_uint64(seed & (2 ** 64 - 1)), site0, nsites, _ptr(comps)), self._h)
        return comps

    def simulate_injected(self, u, want_comps: bool = False, keep_internal: bool = False):
        u = np.ascontiguousarray(u, dtype=np.float64)
        rows = self.n_nodes + (2 if self.is_mixture else 1)
        if u.ndim != 2 or u.shape[0] != rows:
            raise L.ElxError(L.ELX_E_INVALID, "simulate_injected: U must have %d rows" % rows)
        nsites = u.shape[1]
        leaves = np.empty((self.n_leaves, nsites), dtype=np.uint8)
        comps = np.empty(nsites, dtype=np.int32) if want_comps else None
        internal = np.empty((self.n_nodes, nsites), dtype=np.uint8) if keep_internal else None
        L.check(L.lib().elx_simulate_injected(self._h, _ptr(u), max(nsites, 1), nsites, _ptr(leaves), max(nsites, 1), _ptr(comps),
                                              _ptr(internal), max(nsites, 1)), self._h)
        return leaves, comps, internal

    def simulate_device(self, seed: int, site0: int, nsites: int, d_leaves: int, leaves_pitch: int, d_comps: int = 0,
                        d_internal: int = 0, internal_pitch: int = 0):
        """Enqueue on self.stream; pointers are raw device addresses (e.g. torch.Tensor.data_ptr())."""
        L.check(L.lib().elx_simulate_device(self._h, C.c_uint64(seed & (2 ** 64 - 1)), site0, nsites, C.c_void_p(d_leaves),
                                            leaves_pitch, C.c_void_p(d_comps or None), C.c_void_p(d_internal or None),
                                            internal_pitch), self._h)

    def synchronize(self):
        """Wait for the enqueued device work and raise what the kernels flagged (elx_synchronize)."""
        L.check(L.lib().elx_synchronize(self._h), self._h)

    # -- FASTA -----------------------------------------------------------------------------------
    @staticmethod
    def _names(names: Sequence[str]):
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        return arr

    def fasta_size(self, names: Sequence[str], total_sites: int) -> int:
        return int(L.lib().elx_fasta_size(self._h, self._names(names), total_sites))

    def simulate_fasta(self, seed: int, total_sites: int, names: Sequence[str], alphabet: str, want_comps: bool = False):
        if len(names) != self.n_leaves:
            raise L.ElxError(L.ELX_E_INVALID, "fasta: need one name per leaf")
        size = self.fasta_size(names, total_sites)
        out = np.empty(size, dtype=np.uint8)
        comps = np.empty(total_sites, dtype=np.int32) if want_comps else None
        L.check(L.lib().elx_simulate_fasta(self._h, C.c_uint64(seed & (2 ** 64 - 1)), total_sites, self._names(names),
                                           alphabet.encode(), _ptr(out), size, _ptr(comps)), self._h)
        return out.tobytes(), comps

    def simulate_fasta_gz(self, seed: int, total_sites: int, names: Sequence[str], alphabet: str, path: str, level: int = 6,
                          threads: int = 0, want_comps: bool = False):
        """`<base>.fasta.gz` streamed from the device (writeGZFile, InputOutput.hs:80-83): D2H, deflate and the file write overlap."""
        if len(names) != self.n_leaves:
            raise L.ElxError(L.ELX_E_INVALID, "fasta: need one name per leaf")
        comps = np.empty(total_sites, dtype=np.int32) if want_comps else None
        L.check(L.lib().elx_simulate_fasta_gz(self._h, C.c_uint64(seed & (2 ** 64 - 1)), total_sites, self._names(names),
                                              alphabet.encode(), path.encode(), level, threads, _ptr(comps)), self._h)
        return comps

    def fasta_pack_device(self, d_leaves: int, leaves_pitch: int, site0: int, nsites: int, total_sites: int,
                          names: Sequence[str], alphabet: str, d_out: int):
        L.check(L.lib().elx_fasta_pack_device(self._h, C.c_void_p(d_leaves), leaves_pitch, site0, nsites, total_sites,
                                              self._names(names), alphabet.encode(), C.c_void_p(d_out)), self._h)

    def fasta_pack_slice_device(self, d_leaves: int, leaves_pitch: int, nsites: int, alphabet: str, d_out: int, out_pitch: int):
        """K3 slice mode: [T][out_pitch] characters into any device-accessible buffer (e.g. pinned host memory)."""
        L.check(L.lib().elx_fasta_pack_slice_device(self._h, C.c_void_p(d_leaves), leaves_pitch, nsites, alphabet.encode(),
                                                    C.c_void_p(d_out), out_pitch), self._h)

    def fasta_row_offset(self, names: Sequence[str], total_sites: int, t: int) -> int:
        return int(L.lib().elx_fasta_row_offset(self._h, self._names(names), total_sites, t))

    def histogram_device(self, d_leaves: int, leaves_pitch: int, nsites: int, d_state_counts: int = 0, d_pattern_counts: int = 0):
        L.check(L.lib().elx_histogram_device(self._h, C.c_void_p(d_leaves), leaves_pitch, nsites,
                                             C.c_void_p(d_state_counts or None), C.c_void_p(d_pattern_counts or None)), self._h)


# --------------------------------------------------------------------------------------------------
# The reference's function surface (MarkovProcessAlongTree.hs:18-28)
# --------------------------------------------------------------------------------------------------
def simulate_and_flatten_par(n: int, d, e, tree, seed: int, device: int = -1) -> np.ndarray:
    """simulateAndFlattenPar n d e t g (:101-125): leaf states [T][n], leaves in DFS left-to-right order."""
    with Simulator(None, d, e, tree, is_mixture=False, device=device) as sim:
        return sim.simulate(seed, n)[0]


def simulate_and_flatten_mixture_model_par(n: int, ws, ds, es, tree, seed: int, device: int = -1):
    """simulateAndFlattenMixtureModelPar n ws ds es t g (:238-262): (component indices [n], leaf states [T][n])."""
    with Simulator(ws, ds, es, tree, is_mixture=True, device=device) as sim:
        leaves, comps, _ = sim.simulate(seed, n, want_comps=True)
        return comps, leaves


def simulate(n: int, d, e, tree, seed: int, device: int = -1) -> np.ndarray:
    """simulate n d e t g (:130-142): states of every node [N][n] in preorder (the reference's `Tree [State]`)."""
    with Simulator(None, d, e, tree, is_mixture=False, device=device) as sim:
        return sim.simulate(seed, n, keep_internal=True)[2]


def simulate_mixture_model(n: int, ws, ds, es, tree, seed: int, device: int = -1) -> np.ndarray:
    """simulateMixtureModel n ws ds es t g (:267-280)."""
    with Simulator(ws, ds, es, tree, is_mixture=True, device=device) as sim:
        return sim.simulate(seed, n, keep_internal=True)[2]


simulateAndFlattenPar = simulate_and_flatten_par
simulateAndFlattenMixtureModelPar = simulate_and_flatten_mixture_model_par
simulateMixtureModel = simulate_mixture_model
