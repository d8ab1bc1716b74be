/*
 * elynx_b200 -- C ABI of the B200-native alignment simulator.
 *
 * This is the drop-in boundary for ONE path of dschrempf/elynx: the per-site continuous-time
 * Markov process run along a phylogeny,
 *
 *   elynx-markov/src/ELynx/Simulate/MarkovProcessAlongTree.hs
 *     simulateAndFlattenPar              :101-125   (single substitution model)
 *     simulateAndFlattenMixtureModelPar  :238-262   (mixture model, returns component indices)
 *     simulate / simulateMixtureModel    :130-155, :267-294  (same walk, keeps internal states)
 *   elynx-markov/src/ELynx/Simulate/MarkovProcess.hs
 *     probMatrix :33-41, jump :48-49
 *
 * called from slynx/src/SLynx/Simulate/Simulate.hs:100-127 (simulateAlignment).  The reference has
 * no FFI today; INTEGRATION.md shows the `foreign import ccall` stub a maintainer would add.
 *
 * Conventions
 *   - plain C types only; every array is caller-owned; the library owns its handle + device memory.
 *   - every function returns 0 on success and a negative elx_status otherwise;
 *     elx_last_error(h) (or elx_last_error(NULL) for a failed elx_create) gives the message.
 *     The reference's convention is `error` (abort): MarkovProcess.hs:35-37, RateMatrix.hs:125.
 *   - states are indices into the sorted alphabet (0..K-1) stored as uint8
 *     (reference: `type State = Int`, MarkovProcess.hs:29).
 *   - the tree arrives flattened in PREORDER (root first, children left to right), exactly the order
 *     in which simulateAndFlatten' (:88-98) consumes random numbers:
 *       parent[n] < n, parent[0] = -1; brlen[n] = length of the branch above n (root stem, 0 if
 *       absent: elynx-tree Phylogeny.hs:374-381); leaf_row[n] = index of the leaf in DFS left-to-right
 *       order (= `leaves`, Rooted.hs:244-248) or -1 for internal nodes.
 *   - model inputs are used AS GIVEN, like the reference simulator does: pi[c] and exch[c] are
 *     whatever SubstitutionModel/MixtureModel construction produced; Q_c = setDiagonal(exch_c * diag(pi_c))
 *     (RateMatrix.hs:87-103).  Weights need not be normalised (mwc-random `categorical` rescales).
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with ELX_E_CUDA.
 */
#ifndef ELYNX_B200_H
#define ELYNX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct elx_handle elx_handle;

typedef enum {
  ELX_OK = 0,
  ELX_E_INVALID = -1,     /* bad argument (what the reference rejects with `error`) */
  ELX_E_CUDA = -2,        /* CUDA runtime failure / no device */
  ELX_E_UNSUPPORTED = -3, /* e.g. a tree needing more than 32 live state slots, a non-reversible model with K > 32 */
  ELX_E_NOMEM = -4,
  ELX_E_BAD_WEIGHTS = -5  /* mwc-random "categorical: bad weights!" */
} elx_status;

/* (weights, stationary distributions, exchangeability matrices) -- the argument triple of
 * simulateAndFlattenMixtureModelPar (:238-246); n_components = 1 and is_mixture = 0 gives the
 * simulateAndFlattenPar (:101-108) flavour, which draws no component. */
typedef struct {
  int32_t n_states;      /* K: 4 (DNA) or 20 (Protein); any 2..64 accepted */
  int32_t n_components;  /* C >= 1 */
  int32_t is_mixture;    /* 0: single-Q API (C must be 1); 1: mixture API (component draw per site) */
  const double *weights; /* [C]      (ignored when is_mixture == 0) */
  const double *pi;      /* [C][K]   stationary distributions */
  const double *exch;    /* [C][K][K] exchangeability matrices, row-major (hmatrix Matrix R layout) */
} elx_model;

/* `Data.Tree.Tree Double` of branch lengths (slynx Simulate.hs:101), flattened in preorder. */
typedef struct {
  int32_t n_nodes;         /* N */
  const int32_t *parent;   /* [N] */
  const double *brlen;     /* [N], >= 0 (negative -> ELX_E_INVALID, cf. MarkovProcess.hs:37) */
  const int32_t *leaf_row; /* [N] */
} elx_tree;

typedef struct {
  int32_t device;        /* CUDA device ordinal; -1 = current device */
  int32_t philox_rounds; /* 0 = default (10) */
  int32_t flags;         /* ELX_FLAG_* */
  int32_t n_devices;     /* 0 or 1: one device.  N > 1: the handle drives devices device .. device+N-1 (device = -1: from 0);
                            -1: every visible device from `device` on.  The reference's *Par calls parallelise inside the call
                            (MarkovProcessAlongTree.hs:109-125, :221-262); so do elx_simulate / elx_simulate_fasta /
                            elx_simulate_packed of such a handle: contiguous site ranges by the getChunks rule (:210-215), one
                            host thread + stream per device, every device writing its columns of the caller's buffers.  The
                            *_device entry points and the parity hooks address the first device only. */
} elx_opts;

#define ELX_FLAG_FORCE_GENERIC 1 /* always use the generic global-table kernel (cross-check of the tiled one) */
#define ELX_FLAG_SINGLE_DRAWS 4  /* models with K <= 4: draw node by node (one alias row per node) instead of the default joint
                                    draw of sibling pairs; a different (equally exact) free-running stream */
#define ELX_FLAG_PAIR_DRAWS 8    /* models with K <= 4: draw sibling pairs (pair mode) instead of the default quad mode (one draw per cap
                                    of up to 4 frontier nodes, inner nodes summed out); a different, equally exact stream */
#define ELX_FLAG_PADE 2          /* build P(t) by the scaled Pade-7 + squaring kernel (hmatrix `expm`'s own algorithm).  This is
                                    the DEFAULT for models with at most 32 states (element-wise parity <= 1e-12 with the
                                    reference's probMatrix on every branch length); the flag forces it */
#define ELX_FLAG_EIGEN 16        /* build P(t) in the eigen form I + A diag(expm1(lambda t)) B instead (reversible components
                                    only; max-norm parity <= 1e-12, element-wise ~1e-10 on vanishing entries); models with more
                                    than 32 states always take it */

int elx_version(void);
/* The library keeps freed device buffers (tables, staging) in a per-device cache for the next handle / call (up to
 * ELX_DEVICE_CACHE_MB, default 8192; 0 = off); this returns them to the driver. */
void elx_trim_device_cache(void);
/* Result buffers owned by the library.  simulateAndFlatten*Par (MarkovProcessAlongTree.hs:101-125, :238-262) RETURN their
 * result, so the stand-in decides where it lives: a fresh 10 GB allocation costs more than the simulation and its transport
 * together (the kernel zeroes every page at its first write).  elx_result_alloc returns 2 MB aligned, huge-page advised host
 * memory; elx_result_free recycles it for the next elx_result_alloc of a similar size (up to ELX_RESULT_POOL_MB, default
 * 32768, of idle buffers are kept; elx_result_pool_trim unmaps them), so that a process that simulates repeatedly pays the
 * page faults once.  Any other host pointer works for every call as well. */
void *elx_result_alloc(int64_t bytes);
void elx_result_free(void *p);
void elx_result_pool_trim(void);
const char *elx_last_error(const elx_handle *h);

/* Builds the per-component eigensystems on the host, uploads the tree, and runs kernel K1
 * (batched fp64 P(t) = I + A diag(expm1(lambda t)) B for every branch x component, stored as per-row
 * cumulative tables + alias tables).  Replaces toProbTree / toProbTreeMixtureModel
 * (MarkovProcessAlongTree.hs:54-55, :157-161). */
int elx_create(const elx_model *model, const elx_tree *tree, const elx_opts *opts, elx_handle **out);
void elx_destroy(elx_handle *h);

/* Host-only: validates the preorder arrays and compiles the tree walk (heavy child last); reports the number of leaves
 * and of live state slots the walk needs (<= floor(log2 N) + 1 for any tree).  walk_nodes (optional, [N]) receives the
 * visiting order as preorder node indices. */
int elx_tree_walk_info(const elx_tree *tree, int32_t *n_leaves, int32_t *n_slots, int32_t *walk_nodes);

/* Host-only: the pair walk (see elx_pair_groups below) of a tree: number of records, live state slots, and the records
 * themselves, 7 int32 each: node_a, node_b (-1: group of one), src slot (-1: registers), dst slot of a / of b (-1: none),
 * leaf row of a / of b (-1: internal).  recs (optional) must hold 7 * n_nodes values (there are at most n_nodes records). */
int elx_tree_pair_walk_info(const elx_tree *tree, int32_t *n_groups, int32_t *n_slots, int32_t *recs);

/* number of devices the handle drives (elx_opts.n_devices resolved) */
int32_t elx_n_devices(const elx_handle *h);
/* Host-only: the site range of shard `shard` of `n_shards` over [0, nsites) -- getChunks c n (MarkovProcessAlongTree.hs:210-215:
 * r = n mod c chunks of n div c + 1 sites, then chunks of n div c) with the inner boundaries rounded up to multiples of `align`
 * (the multi-device calls use 64).  The shards tile [0, nsites) exactly for every alignment. */
int elx_shard_range(int64_t nsites, int32_t n_shards, int32_t shard, int32_t align, int64_t *first, int64_t *count);
int32_t elx_n_leaves(const elx_handle *h);
int32_t elx_n_nodes(const elx_handle *h);
int32_t elx_n_states(const elx_handle *h);
int32_t elx_n_components(const elx_handle *h);
/* the CUDA stream (cudaStream_t) every kernel of this handle is launched on */
void *elx_stream(const elx_handle *h);

/* Parity hook for probMatrix (MarkovProcess.hs:33-41): copy the [N][C][K][K] tables to the host,
 * cumulative != 0 -> per-row cumulative sums (scanl1 (+)), else the probabilities themselves. */
int elx_ptables_get(elx_handle *h, double *out, int cumulative);
/* Parity hook: replace K1's tables by caller-supplied probability matrices [N][C][K][K]
 * (e.g. the oracle's Pade tables) and rebuild the cumulative + alias tables from them on the device. */
int elx_ptables_set(elx_handle *h, const double *p);
/* Parity hook for the free-running sampler: alias tables derived from the P rows.
 * e16 [N][C][K][KP] (uint16: [threshold high bits : 16-A][alias : A], A = log2 KP), qlo [N][C][K][KP] (uint32 threshold low bits),
 * KP = elx_alias_columns(h).  See DESIGN.md "free-running draw". */
int32_t elx_alias_columns(const elx_handle *h);
int elx_alias_get(elx_handle *h, uint16_t *e16, uint32_t *qlo);

/* Pair mode (models with at most 4 states, unless ELX_FLAG_SINGLE_DRAWS): siblings are independent given their parent, so
 * the free-running kernels draw the children of a node two at a time from the joint row P_A[i][sa] * P_B[i][sb]
 * (16 outcomes, o = sa | sb << 2) -- one draw per alignment cell instead of two.  elx_pair_groups: number of records G
 * (0 = the handle draws node by node).  elx_pair_get: the records in walk order (node_a[G], node_b[G], -1 = a group of
 * one) and their alias rows pe16 / pqlo [G][C][K][16] ([threshold high bits : 12][alias : 4] / 32 low threshold bits). */
int32_t elx_pair_groups(const elx_handle *h);
int elx_pair_get(elx_handle *h, int32_t *node_a, int32_t *node_b, uint16_t *pe16, uint32_t *pqlo);

/* Quad mode (the default for models with at most 4 states when only the leaves are requested; ELX_FLAG_PAIR_DRAWS /
 * ELX_FLAG_SINGLE_DRAWS turn it off): simulateAndFlatten (:71-98) keeps the leaves only, so internal nodes whose state
 * nobody needs are summed out instead of drawn.  The tree is cut into "caps" -- pieces below a node r of known state with at
 * most 4 frontier nodes (leaves, or internal nodes kept because further caps hang below them) -- and every cap is ONE draw
 * from the 256-outcome joint row P(frontier states | state of r), o = sum_k state(out[k]) << 2k.  The joint distribution of
 * the leaves is exactly the reference's.  Calls that keep internal states (`internal` != NULL) draw pairs instead.
 * A record is ELX_QUAD_REC_INTS int32: [0] r, [1..4] out[k] (-1: absent), [5..8] leaf row of out[k] (-1: internal),
 * [9..12] slot receiving out[k] (-1: none), [13] slot holding r, [14] number of cap nodes n, [15] 0, then n triples
 * [16+3i..] = (node, index of its parent in this list or -1 for r, frontier position or -1 for a summed-out node), parents
 * first.  elx_quad_groups: number of records (0: no quad tables); elx_quad_get: the records in walk order and their alias
 * rows qe32 / qres32 [G][C][4][256] (uint32).  A draw is 24 bits, x24 = [t : 16][column j : 8] (three Philox calls per 16 sites).
 * Two-level rows, exact for the integer masses round(p 2^48) without any draw / threshold tie: t != 0 -> main entry qe32[..][j] =
 * [q : 16][alias : 8][0 : 8], outcome = x24 < (entry >> 8) ? j : alias (cells of 2^-24); t == 0 (2^-16) -> 32 fresh bits
 * u = [tr : 24][jr : 8] against the escape entry qres32[..][jr] = [thr : 24][alias : 8] (the residual masses): tr < thr ? jr : alias.
 * elx_tree_quad_walk_info: host-only, the records of a tree (recs may be NULL; it must hold ELX_QUAD_REC_INTS * n_nodes). */
#define ELX_QUAD_REC_INTS 52
int32_t elx_quad_groups(const elx_handle *h);
int elx_quad_get(elx_handle *h, int32_t *recs, uint32_t *qe32, uint32_t *qres32);
int elx_tree_quad_walk_info(const elx_tree *tree, int32_t *n_records, int32_t *n_slots, int32_t *recs);
/* the same cut with at most TWO frontier nodes per cap: the records of duo mode (below) */
int elx_tree_duo_walk_info(const elx_tree *tree, int32_t *n_records, int32_t *n_slots, int32_t *recs);

/* Duo mode (the default for models with 5..22 states -- amino acids -- when only the leaves are requested; ELX_FLAG_SINGLE_DRAWS
 * turns it off): the same summing-out with caps of at most TWO frontier nodes, one draw from the K*K-outcome joint row
 * P(frontier states | state of r) in 512 alias columns, o = state(out[0]) + K * state(out[1]).  A record's rows are 40 KB per
 * component, so every call first buckets its sites by component (stable inside superblocks of 2^20 .. 2^24 sites); the rank of a site
 * among the sites of its component and superblock selects its 16-site Philox group (see elx_duo.cu for the specification).
 * elx_duo_groups: number of records (0: the handle draws node by node); elx_duo_get: the records (ELX_QUAD_REC_INTS layout,
 * frontier positions 0..1) and their two-level alias rows de32 / dres32 [G][C][K][512] (uint32): x24 = [t : 15][column j : 9];
 * t != 0 -> de32[..][j] = [q : 15][alias : 9][0 : 8], outcome = x24 < (entry >> 8) ? j : alias; t == 0 (2^-15) -> u =
 * [tr : 23][jr : 9] against dres32[..][jr] = [thr : 23][alias : 9]; integer masses round(p 2^47). */
int32_t elx_duo_groups(const elx_handle *h);
int elx_duo_get(elx_handle *h, int32_t *recs, uint32_t *de32, uint32_t *dres32);

/* ---- simulateAndFlatten[MixtureModel]Par ------------------------------------------------------
 * Simulates sites [site0, site0 + nsites) of the alignment defined by (handle, seed, draw mode); the result is a
 * pure function of (seed, global site index, node), so any partition of the site range over calls,
 * GPUs or processes yields the same alignment.
 * Draw mode: leaves-only calls (internal == NULL) sum internal nodes out where they can -- quad mode for K <= 4, duo mode
 * for 5 <= K <= 22 -- while calls that keep the internal states draw every node (pair mode for K <= 4, node by node
 * otherwise): a different, equally exact stream, so the SAME handle and seed give different (identically distributed)
 * leaf rows with and without `internal`.  (In the reference `simulate` and `simulateAndFlatten` consume the generator
 * alike.)  A handle created with ELX_FLAG_PAIR_DRAWS (K <= 4) or ELX_FLAG_SINGLE_DRAWS (any K) uses ONE stream for both
 * call shapes: leaves(internal != NULL) == leaves(internal == NULL).
 *   leaves   [T][leaves_pitch]   uint8, row = leaf_row, column = site - site0       (required)
 *   comps    [nsites]            int32 component index per site, or NULL             (:257, `cs`)
 *   internal [N][internal_pitch] uint8 states of every node in preorder, or NULL     (`simulate`, :130-155)
 * The *_device variants take device pointers and only enqueue work on elx_stream(h);
 * the host variants copy results back and synchronise. */
int elx_simulate(elx_handle *h, uint64_t seed, int64_t site0, int64_t nsites, uint8_t *leaves, int64_t leaves_pitch,
                 int32_t *comps, uint8_t *internal, int64_t internal_pitch);
int elx_simulate_device(elx_handle *h, uint64_t seed, int64_t site0, int64_t nsites, uint8_t *d_leaves,
                        int64_t leaves_pitch, int32_t *d_comps, uint8_t *d_internal, int64_t internal_pitch);

/* The same alignment in the form it crosses PCIe in, for callers that keep alignments packed (statistics, compressors, a
 * .fasta.gz writer): leaves-only, no expansion on the host, so the call runs at PCIe speed instead of the host's store
 * bandwidth.  elx_packed_bits: 2 (K <= 4: site 4i+k of a row in bits 2k..2k+1 of byte i) or 5 (K <= 32: site i in bits
 * 5i..5i+4 of the row's little-endian bit stream, 8 sites = 5 bytes); rows are padded to whole 32-site groups:
 * elx_packed_pitch(nsites) bytes.  elx_unpack2_rows / elx_unpack5_rows (below) expand such rows.  K > 32: ELX_E_UNSUPPORTED. */
int32_t elx_packed_bits(const elx_handle *h);
int64_t elx_packed_pitch(const elx_handle *h, int64_t nsites);
int elx_simulate_packed(elx_handle *h, uint64_t seed, int64_t site0, int64_t nsites, uint8_t *packed, int64_t packed_pitch,
                        int32_t *comps);

/* The component index of every site of [site0, site0 + nsites) alone -- the `cs` of simulateAndFlattenMixtureModelPar (:257)
 * without the alignment (writeSiteDists of a sharded run, slynx Simulate.hs:79-89).  Same values as `comps` above. */
int elx_site_components(elx_handle *h, uint64_t seed, int64_t site0, int64_t nsites, int32_t *comps);

/* Waits for the work enqueued by the *_device entry points on elx_stream(h) and reports what the kernels flagged
 * ("categorical: bad weights!", ELX_E_BAD_WEIGHTS); the host variants do this themselves. */
int elx_synchronize(elx_handle *h);

/* Injected-uniforms mode: consumes the uniform matrix U[(is_mixture ? 2 : 1) + N][u_pitch] (doubles in
 * (0,1]) in the reference's draw order -- [component row (mixtures)] [root-state row] [one row per node in
 * preorder] -- with mwc-random `categorical` semantics (first i with cum[i] >= cum[K-1]*u, fp64).
 * Must reproduce the reference algorithm bit for bit given the same tables. */
int elx_simulate_injected(elx_handle *h, const double *U, int64_t u_pitch, int64_t nsites, uint8_t *leaves,
                          int64_t leaves_pitch, int32_t *comps, uint8_t *internal, int64_t internal_pitch);
int elx_simulate_injected_device(elx_handle *h, const double *d_U, int64_t u_pitch, int64_t nsites, uint8_t *d_leaves,
                                 int64_t leaves_pitch, int32_t *d_comps, uint8_t *d_internal, int64_t internal_pitch);

/* ---- FASTA emission (elynx-seq Export/Fasta.hs:24-36; slynx Simulate.hs:117-127) ---------------
 * Layout: for each leaf in leaf_row order `>` name `\n` S characters `\n`.
 * elx_fasta_size: total bytes for an alignment of total_sites columns.
 * elx_fasta_pack_device: K3 -- maps states -> characters of `alphabet` (K chars) and writes the columns
 *   [site0, site0+nsites) of every sequence line (plus the header lines when site0 == 0 and the trailing
 *   newlines when site0 + nsites == total_sites) into d_out, a buffer holding the WHOLE file image.
 * elx_simulate_fasta: simulate + pack + copy to the host buffer `out` (size >= elx_fasta_size). */
int64_t elx_fasta_size(const elx_handle *h, const char *const *names, int64_t total_sites);
int elx_fasta_pack_device(elx_handle *h, const uint8_t *d_leaves, int64_t leaves_pitch, int64_t site0, int64_t nsites,
                          int64_t total_sites, const char *const *names, const char *alphabet, uint8_t *d_out);
int elx_simulate_fasta(elx_handle *h, uint64_t seed, int64_t total_sites, const char *const *names, const char *alphabet,
                       uint8_t *out, int64_t out_capacity, int32_t *comps);
/* The same alignment written as a gzip file (what writeGZFile, elynx-tools InputOutput.hs:80-83, does for a path ending in
 * .gz): simulated into device memory, then streamed in blocks of sequence lines -- K3 writes the characters of a block into
 * pinned host memory while host threads deflate the previous block into gzip members (4 MiB of text each) that are written
 * in file order.  gunzip / the reference's `decompress` return exactly the bytes elx_simulate_fasta produces.
 * level: zlib 0..9 (else 6); n_threads <= 0: all host cores.  Fails with ELX_E_NOMEM when the alignment does not fit the device. */
int elx_simulate_fasta_gz(elx_handle *h, uint64_t seed, int64_t total_sites, const char *const *names, const char *alphabet,
                          const char *path, int32_t level, int32_t n_threads, int32_t *comps);
/* K3 in slice mode: the characters of columns [0, nsites) of every sequence as a plain [T][out_pitch] matrix (no header
 * lines, no newlines).  d_out may be any device-accessible pointer, in particular pinned host memory: with several GPUs
 * every rank packs its site range straight into its own pinned buffer and the host writes each row segment into the
 * file at elx_fasta_row_offset(t) + site0. */
int elx_fasta_pack_slice_device(elx_handle *h, const uint8_t *d_leaves, int64_t leaves_pitch, int64_t nsites,
                                const char *alphabet, uint8_t *d_out, int64_t out_pitch);
/* byte offset of the first sequence character of leaf row t in the FASTA image of a total_sites-column alignment */
int64_t elx_fasta_row_offset(const elx_handle *h, const char *const *names, int64_t total_sites, int32_t t);

/* ---- validation histograms (new; the reference has none) ---------------------------------------
 * d_state_counts [T][K] int64 += per-taxon state counts; d_pattern_counts [K^T] int64 += site-pattern counts
 * (pattern index = sum_t state_t * K^(T-1-t); only when K^T <= 2^24; pass NULL to skip).  Device buffers are
 * accumulated into, so per-GPU results can be summed with one NCCL all-reduce. */
int elx_histogram_device(elx_handle *h, const uint8_t *d_leaves, int64_t leaves_pitch, int64_t nsites,
                         int64_t *d_state_counts, int64_t *d_pattern_counts);

/* ---- alignment statistics: the core of `slynx examine` on the device ------------------
 * Replaces, for an alignment held as a code matrix, examineAlignment (slynx/src/SLynx/Examine/Examine.hs:52-160) and what it
 * calls: filterColsConstant / ConstantSoft / NoGaps / OnlyStd, toFrequencyData, distribution, kEffEntropy, kEffHomoplasy,
 * countGaps (elynx-seq Alignment.hs:209-300), entropy / homoplasy (DistributionDiversity.hs:37-110), hamming
 * (Distance.hs:24-33) and divergence (Divergence.hs:38-64).
 *   states [T][pitch] uint8 codes: 0..K-1 standard characters in sorted order (what elx_simulate writes), K = '-', K+1 = '.'
 *   (the gap characters of DNAX / ProteinX: they never enter a frequency).  IUPAC alphabets are out of scope.
 *   keff_entropy_per_site [n_sites] doubles or NULL (`slynx examine --per-site`).
 *   equal_counts [T][T] int64 or NULL: upper triangle (i < j) = number of sites where sequences i and j carry the same
 *   character; the Hamming distance is n_sites - equal_counts[i][j].
 *   divergence [T(T-1)/2][K][K] int64 or NULL: count matrices of the pairs (0,1), (0,2), ..., (1,2), ... (`--divergence`).
 * The _device variant takes device pointers, enqueues on `stream` (a cudaStream_t, may be NULL) and synchronises it before it
 * returns; elx_examine takes host pointers.  Errors: elx_last_error(NULL). */
typedef struct {
  int64_t n_sites, n_sequences;
  int32_t n_states, reserved;
  int64_t n_constant;      /* columns whose characters are all equal */
  int64_t n_constant_soft; /* ... counting only standard characters, at least one of them */
  int64_t n_no_gaps;       /* columns without gaps */
  int64_t n_only_std;      /* columns with standard characters only (= n_no_gaps without IUPAC codes) */
  int64_t n_std_chars, n_gaps;
  double distribution[64]; /* mean over the columns of the per-column frequencies of the standard characters */
  double keff_entropy_mean, keff_entropy_mean_no_gaps;     /* mean effective number of states, exp(entropy) per column */
  double keff_homoplasy_mean, keff_homoplasy_mean_no_gaps; /* ... 1 / sum f^2 per column */
  double hamming_mean, hamming_var, hamming_min, hamming_max; /* pairwise, per site; variance = maximum likelihood */
} elx_examine_result;

int elx_examine_device(const uint8_t *d_states, int64_t pitch, int32_t n_sequences, int64_t n_sites, int32_t n_states,
                       elx_examine_result *out, double *d_keff_entropy_per_site, int64_t *d_equal_counts, int64_t *d_divergence,
                       void *stream);
int elx_examine(const uint8_t *states, int64_t pitch, int32_t n_sequences, int64_t n_sites, int32_t n_states,
                elx_examine_result *out, double *keff_entropy_per_site, int64_t *equal_counts, int64_t *divergence);

/* Number of kernels this handle has launched so far (bench.py's gpu_launches). */
int64_t elx_launch_count(const elx_handle *h);
/* Name (with its template arguments) of the simulation kernel the last elx_simulate*_device call of this handle launched,
 * e.g. "k2_quad_tiled<10,256,2>"; "" before the first call.  bench.py reports it next to every roofline figure. */
const char *elx_last_kernel(const elx_handle *h);
/* Bytes the host-buffer calls of this handle (elx_simulate, elx_simulate_injected) have copied device -> host so far.
 * Alignments over <= 4 states cross PCIe packed to 2 bits per site, over <= 32 states to 5 bits, and are expanded into the
 * caller's buffer by host threads (ELX_HOST_THREADS, default: all cores but one -- divided by LOCAL_WORLD_SIZE under
 * torchrun --, at most 32; ELX_PACKED_D2H=0 turns the packed transport off). */
int64_t elx_d2h_bytes(const elx_handle *h);
/* The host half of that transport, exported for tests and for callers that keep packed alignments: expands a 2-bit packed
 * matrix (site 4i+k of a row in bits 2k..2k+1 of byte i) to one byte per site -- state codes 0..3, or lut4[state] when a
 * 4-character table is given -- with n_threads host threads (<= 0: all cores). */
int elx_unpack2_rows(const uint8_t *packed, int64_t packed_pitch, uint8_t *dst, int64_t dst_pitch, int64_t rows, int64_t nsites,
                     const char *lut4, int32_t n_threads);
/* Alphabets of 5..32 states (amino acids) travel at 5 bits per site: site i of a row in bits 5i..5i+4 of the row's
 * little-endian bit stream (8 sites = 5 bytes); lut32: up to 32 characters, or NULL for state codes. */
int elx_unpack5_rows(const uint8_t *packed, int64_t packed_pitch, uint8_t *dst, int64_t dst_pitch, int64_t rows, int64_t nsites,
                     const char *lut32, int32_t n_threads);

#ifdef __cplusplus
}
#endif
#endif /* ELYNX_B200_H */
