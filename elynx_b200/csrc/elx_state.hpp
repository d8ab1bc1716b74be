// The opaque handle behind include/elynx_b200.h and the argument block shared by the simulation kernels.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "elx_internal.hpp"

struct elx_handle {
  int N = 0, K = 0, C = 0, T = 0, is_mixture = 0, A = 0, KP = 0, device = 0, n_slots = 0, rounds = 10, flags = 0;
  cudaStream_t stream = nullptr;
  double *d_brlen = nullptr, *d_lambda = nullptr, *d_A = nullptr, *d_B = nullptr, *d_Q = nullptr;
  int use_pade = 0;
  double *d_P = nullptr, *d_cum = nullptr, *d_cumw = nullptr, *d_cumpi = nullptr;
  uint16_t *d_e16 = nullptr;
  uint32_t *d_qlo = nullptr;
  elx::WalkRec *d_walk = nullptr;
  int *d_error_flag = nullptr;
  // FASTA frame (names, offsets, alphabet LUT)
  uint8_t *d_names = nullptr, *d_lut = nullptr;
  int64_t *d_name_off = nullptr, *d_seq_off = nullptr;
  int64_t fasta_total_sites = -1;
  std::string fasta_names_key;
  // tiled ("fast") kernel tables, owned by elx_fast.cu
  void *fast = nullptr;
  // pair mode (K <= 4): joint two-children draws, owned by elx_pair.cu
  void *pair = nullptr;
  // quad mode (K <= 4): one draw per cap of up to 4 frontier nodes, inner nodes summed out; owned by elx_quad.cu
  void *quad = nullptr;
  // duo mode (5 <= K <= 22): one draw per cap of up to 2 frontier nodes on component-bucketed sites; owned by elx_duo.cu
  void *duo = nullptr;
  int pair_ready = 0, fast_ready = 0;  // pair rows / node-by-node stage images built from the current P tables (lazy)
  // multi-device handles (elx_opts.n_devices > 1): the handle the caller holds is the primary, peers[] are full handles on
  // the other devices (own stream, own tables); the host-buffer calls shard the site range over all of them
  std::vector<double> weights;  // host copy of the component weights (all 1 for the single-model API)
  std::vector<elx_handle *> peers;
  int is_peer = 0;
  int64_t launches = 0;
  int64_t d2h_bytes = 0;  // bytes the host-buffer calls moved device -> host so far
  std::string last_error;
  std::string last_kernel;  // name of the simulation kernel the last elx_simulate*_device call launched (elx_last_kernel)
};

namespace elx {

// Device memory of the library goes through a small per-device cache (elx_core.cu): cudaMalloc / cudaFree of the gigabyte-size
// staging buffers and tables cost tens of milliseconds per host-buffer call (measured: 40-60 ms of a 105 ms elx_simulate_packed
// call), and handles are typically created, used and destroyed in a loop.  dev_free keeps cudaFree's implicit device
// synchronisation (a block is never handed out again while work that uses it may be in flight); blocks beyond the cap
// (ELX_DEVICE_CACHE_MB, default 8192 per device; 0 = off) go back to the driver, and everything does when an allocation fails.
cudaError_t dev_malloc(void **p, size_t bytes);
cudaError_t dev_free(void *p);
void dev_cache_trim();  // return every cached block of every device to the driver

// Device form of a PairRec (16 bytes, streamed through shared memory next to the group's alias rows).
struct PairRecDev {
  int32_t nodeA;
  int16_t src;
  int8_t dstA, dstB;
  int32_t leafA, leafB;
};
static_assert(sizeof(PairRecDev) == 16, "PairRecDev must stay 16 bytes");

struct SimArgs {
  int N, K, C, is_mixture, n_slots, A, rounds;
  const WalkRec *walk;
  const double *cum;    // [N][C][K][K]
  const double *cumw;   // [C]
  const double *cumpi;  // [C][K]
  const uint16_t *e16;  // [N][C][K][KP]
  const uint32_t *qlo;  // [N][C][K][KP]
  // pair mode: G records in walk order, joint alias rows [G][C][K][16]
  int G;
  const PairRecDev *pairs;
  const int32_t *pair_nodeB;
  const uint16_t *pe16;
  const uint32_t *pqlo;
  uint8_t *leaves;      // [T][leaves_pitch]
  int64_t leaves_pitch;
  int32_t *comps;       // [nsites] or null
  uint8_t *internal;    // [N][internal_pitch] or null
  int64_t internal_pitch;
  int64_t site0, nsites;
  uint64_t seed;
  const double *U;      // injected mode only
  int64_t u_pitch;
  int *error_flag;
};

int fail(elx_handle *h, int code, const std::string &msg);
// elx_io.cpp: expand rows [row0, row1) of a 2-bit packed matrix (4 sites per byte) to one byte per site
void unpack2_rows(const uint8_t *packed, int64_t packed_pitch, uint8_t *dst, int64_t dst_pitch, int64_t row0, int64_t row1,
                  int64_t nsites, const uint8_t *lut4);
// the same for the 5-bit form (8 sites per 5 bytes; lut32: 32 characters or nullptr)
void unpack5_rows(const uint8_t *packed, int64_t packed_pitch, uint8_t *dst, int64_t dst_pitch, int64_t row0, int64_t row1,
                  int64_t nsites, const uint8_t *lut32);

// elx_io.cpp: ordered multi-member gzip stream (worker threads deflate members of block_bytes, written in order)
struct GzStream;
GzStream *gz_open(const char *path, int n_threads, int level, int64_t block_bytes, std::string &err);
bool gz_append(GzStream *g, const uint8_t *data, size_t n);
bool gz_close(GzStream *g, std::string &err);  // flushes, joins the workers, closes the file and frees g

// elx_fast.cu: tiled shared-memory kernels for the shapes that fit (DNA / protein with few components)
int fast_tables_build(elx_handle *h);
void fast_tables_free(elx_handle *h);
int fast_simulate(elx_handle *h, const SimArgs &a, int *used);

// elx_pair.cu: pair mode (models with K <= 4 unless ELX_FLAG_SINGLE_DRAWS): the children of a node are drawn two at a time
int pair_create(elx_handle *h, const int32_t *parent, const int32_t *leaf_row, std::string &msg);  // walk + buffers
int pair_tables_build(elx_handle *h);                                                             // joint alias rows from d_P
void pair_free(elx_handle *h);
int pair_groups(const elx_handle *h);
int pair_get(elx_handle *h, int32_t *nodeA, int32_t *nodeB, uint16_t *pe16, uint32_t *pqlo);
int pair_simulate(elx_handle *h, SimArgs &a, int *used);

// elx_quad.cu: quad mode (models with K <= 4 unless ELX_FLAG_SINGLE_DRAWS / ELX_FLAG_PAIR_DRAWS); leaves h->quad null when
// the tree or model does not fit (the handle then stays in pair mode)
int quad_create(elx_handle *h, const int32_t *parent, const int32_t *leaf_row);
int quad_tables_build(elx_handle *h);
void quad_free(elx_handle *h);
int quad_groups(const elx_handle *h);
void quad_rec_to_ints(const QuadRec &r, int32_t *o);  // ELX_QUAD_REC_INTS values, the layout include/elynx_b200.h documents
int quad_get(elx_handle *h, int32_t *recs, uint32_t *qe32, uint32_t *qres32);
int quad_simulate(elx_handle *h, SimArgs &a, int *used);

// elx_duo.cu: duo mode (models with 5..22 states unless ELX_FLAG_SINGLE_DRAWS; leaves-only calls); leaves h->duo null when
// the tree or model does not fit (the handle then draws node by node)
int duo_create(elx_handle *h, const int32_t *parent, const int32_t *leaf_row);
int duo_tables_build(elx_handle *h);
void duo_free(elx_handle *h);
int duo_groups(const elx_handle *h);
int duo_get(elx_handle *h, int32_t *recs, uint32_t *de32, uint32_t *dres32);
int duo_simulate(elx_handle *h, SimArgs &a, int *used);
int64_t duo_batch_sites(const elx_handle *h);  // sites per batch of a host-buffer call (whole superblocks), 0: no duo tables

// elx_core.cu
int ensure_pair_tables(elx_handle *h);  // build the pair rows / the node-by-node stage images on first use
int ensure_fast_tables(elx_handle *h);
void launch_k1_alias(elx_handle *h, int64_t rows, int K, int A, const double *P, uint16_t *e16, uint32_t *qlo);

}  // namespace elx

// every cudaMalloc / cudaFree of the library's translation units is the cached form
#define cudaMalloc(p, n) elx::dev_malloc((void **)(p), (n))
#define cudaFree(p) elx::dev_free((void *)(p))
