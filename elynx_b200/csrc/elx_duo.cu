// Duo mode of the free-running simulation (K2) for models with 5..22 states (amino acids) on sm_100a.
//
// Same idea as quad mode (elx_quad.cu): simulateAndFlatten keeps the leaves only (MarkovProcessAlongTree.hs:71-98, :178-208),
// so internal nodes whose state nobody needs are SUMMED OUT instead of drawn.  With 20 states a cap can have at most TWO
// frontier nodes (20 x 20 = 400 outcomes <= 512 alias columns; three would need 8000): a random binary tree of T leaves
// is cut into about 0.81 T such caps (compile_quad_walk with max_frontier = 2) -- 1.24 alignment cells per draw instead of
// the 0.5 of the node-by-node walk.
//
// A record's alias rows are K x 512 x 4 B = 40 KB PER COMPONENT, so one CTA must work on sites of ONE component
// (SURVEY hard part 2): every call buckets its sites by component first.  Sites are cut into superblocks of 2^SB sites (SB = 20
// for up to 16 components, 22 for up to 64, else 24); inside
// a superblock the sites of a component keep their order (a stable counting sort); the rank of a site among the sites
// of its component and superblock decides which 16-site group it shares its Philox calls with.  Everything is a pure
// function of (seed, global site), so any partition of a site range over calls and GPUs gives the same alignment.
//
// Draw specification (oracle/freerun_spec.c spec_freerun_duo executes it on the CPU; kernels must agree bit for bit):
//   site s: component c and root state from the shared component / root draws (elx_ptx.cuh draw_comp_root);
//   B = s >> SB;  rank = #{s' in [B << SB, s) : c(s') = c};  q = rank >> 4;  d = rank & 15
//   record g = (out[0..1] | r), state i of r:  48 bytes Z[0..47] = little-endian words of
//     philox(q | c << 20, B lo, out[0], (B hi << 8) | 4 * call), call = 0, 1, 2
//     x24 = draw d of the 12 words as in quad mode (elx_ptx.cuh ELX_DRAW24) = [t : 15][j : 9]
//     t != 0:  e = de32[g][c][i][j] = [q : 15][alias : 9][0 : 8];  o = x24 < (e >> 8) ? j : alias   (one 24-bit compare)
//     t == 0 (escape, 2^-15 per draw):  u = philox(s lo, s hi, out[0], TAG_REFINE).x = [tr : 23][jr : 9];
//              er = dres32[g][c][i][jr] = [thr : 23][alias : 9];  o = tr < thr ? jr : alias
//     state(out[0]) = o mod K,  state(out[1]) = o div K
// Rows carry integer masses round(p / sum * 2^47), split into a main level of 2^-24 cells and an exact residual (two-level
// alias rows, elx_ptx.cuh): no draw ever ties with a threshold.
//
// Per call: k_duo_comps (component + root state of every site from the superblock start, tile histograms) -> k_duo_scan
// (ranks of the tiles) -> k_duo_layout (segments = (superblock, component) runs of 16-site groups, CTA tiles) ->
// k_duo_scatter (site -> bucket slot) -> k2_duo_tiled / k2_duo_generic (states in bucket order, [T][slots]) ->
// k_duo_unpermute (back to site order, [T][S]).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/elynx_b200.h"
#include "elx_device.cuh"
#include "elx_ptx.cuh"
#include "elx_state.hpp"

namespace elx {

constexpr int DUO_COLS = 512;
constexpr int DUO_MAX_K = 22;       // K * K <= 512
constexpr int DUO_MAX_C = 256;      // components travel as bytes through the bucketing passes
constexpr int DUO_MAX_BATCH = 8;    // main / un-permute kernel pairs of one call
// log2 sites per superblock: the sites of ONE component inside a superblock should fill several CTA tiles (8192 slots), and a
// call is pipelined superblock-wise (the un-permute pass of one batch of superblocks runs under the simulation of the next)
__host__ __device__ inline int duo_sb_log2(int C) { return C <= 16 ? 20 : C <= 64 ? 22 : 24; }
constexpr int DUO_TILE = 2048;      // sites per block of the bucketing passes (divides a superblock)
constexpr int DUO_MAX_SLOTS = 24;
constexpr int DUO_MAX_NSTAGE = 4;

struct DuoRecDev {  // 32 bytes
  uint32_t node0;
  int8_t src, dst[2], pad[1];
  int32_t leaf[2];
  int32_t pad2[4];
};
static_assert(sizeof(DuoRecDev) == 32, "DuoRecDev must stay 32 bytes");

struct DuoCapDev {
  int32_t n;
  int32_t node[QUAD_MAX_CAP];
  int8_t par[QUAD_MAX_CAP], fr[QUAD_MAX_CAP];
  int32_t pad;
};

// One CTA of the simulation kernels: up to `threads` 16-site groups of one (superblock, component) segment.
struct DuoTile {
  uint32_t comp, q_first, ngroups, sb_lo, sb_hi, pad;
  uint64_t slot_off;  // first bucket slot of the tile
};
static_assert(sizeof(DuoTile) == 32, "DuoTile must stay 32 bytes");

// (superblock j, component c) -> the run of groups this call simulates and where its slots start
struct DuoSeg {
  uint32_t q0, ngroups;
  uint64_t slot_off;
};

struct DuoTables {
  QuadProgram prog;
  int G = 0;
  DuoRecDev *d_recs = nullptr;
  DuoCapDev *d_caps = nullptr;
  uint8_t *d_ftab = nullptr;    // [G][C] x (32 B header + K x 512 entries)
  uint32_t *d_dres32 = nullptr;  // [G][C][K][512]
  int tiled = 0, threads = 0, minb = 1, nstage = 0, stage_bytes = 0;  // threads: the biggest CTA class that fits
  size_t smem = 0;
  int nstage512 = 0;    // > 0: the 512-thread class fits as well (preferred whenever the tile fits it: 2 % faster)
  size_t smem512 = 0;
  // per-call scratch (grow-only)
  uint8_t *d_comp8 = nullptr, *d_root8 = nullptr, *d_rootb = nullptr, *d_tmp = nullptr;
  uint32_t *d_pos32 = nullptr, *d_site32 = nullptr, *d_hist = nullptr, *d_base = nullptr, *d_total = nullptr, *d_r0 = nullptr;
  DuoSeg *d_seg = nullptr;
  DuoTile *d_tiles = nullptr;
  int *d_ntiles = nullptr, *d_tile_begin = nullptr;
  int sb = 22;
  // pipelined calls: the simulation kernels of the batches run on sm[] (concurrently: the next batch's CTAs fill the SMs the
  // last wave of the previous one leaves idle), the un-permute passes on s2 under the next batch's simulation
  cudaStream_t s2 = nullptr, sm[DUO_MAX_BATCH] = {};
  cudaEvent_t ev_main[DUO_MAX_BATCH] = {}, ev_done = nullptr, ev_bucket = nullptr;
  size_t cap_sites = 0, cap_slots = 0, cap_tmp = 0, cap_tiles = 0, cap_hist = 0, cap_seg = 0;
};

struct DuoArgs {
  SimArgs s;
  int G;
  const DuoRecDev *recs;
  const uint8_t *ftab;
  const uint32_t *dres32;
  int stage_bytes, nstage, slot_stride;
  const DuoTile *tiles;
  const int *tile_begin;  // [nbatch + 1]
  int batch;
  const uint8_t *rootb;
  const uint32_t *site32;
  uint8_t *tmp;
  int64_t tmp_pitch;
  uint64_t base0;          // global site of bucketing index u = 0 (start of the first superblock)
  uint32_t mhi, negk256;   // decode of an outcome o = s0 + K s1: s1 = umulhi(o << 8, mhi), s0 << 8 = (o << 8) + s1 * negk256
  PhiloxRoundKeys rk;
};

// ---------------------------------------------------------------------------------------------
// K1 (duo rows): joint[g][c][i][o] = P(frontier states = o | state of the cap's root = i), o = s0 + K s1, the cap's inner
// nodes summed out by one pruning pass (children before parents).  One thread per (g, c, i, o).
// ---------------------------------------------------------------------------------------------
__global__ void k1_duo_joint(int G, int C, int K, const DuoCapDev *__restrict__ caps, const double *__restrict__ P,
                             double *__restrict__ joint) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)G * C * K * DUO_COLS) return;
  const int o = (int)(idx & (DUO_COLS - 1));
  const int64_t row = idx >> 9;
  const int i = (int)(row % K), c = (int)((row / K) % C), g = (int)(row / K / C);
  const int s0 = o % K, s1 = o / K;
  const DuoCapDev cap = caps[g];
  int nfr = 0;
  for (int n = 0; n < cap.n; ++n)
    if (cap.fr[n] >= 0) ++nfr;
  if (s1 >= K || (nfr < 2 && s1 != 0) || nfr < 1) { joint[idx] = 0.0; return; }  // absent frontier positions are fixed at state 0
  double L[QUAD_MAX_CAP][DUO_MAX_K];
  for (int n = 0; n < cap.n; ++n)
    for (int s = 0; s < K; ++s) L[n][s] = 1.0;
  double top = 1.0;
  for (int n = cap.n - 1; n >= 0; --n) {
    const double *pn = P + ((int64_t)cap.node[n] * C + c) * K * K;
    const int p = cap.par[n], fr = cap.fr[n];
    const int st = fr == 0 ? s0 : s1;
    for (int s = (p >= 0 ? 0 : i); s < (p >= 0 ? K : i + 1); ++s) {
      double msg;
      if (fr >= 0) {
        msg = pn[s * K + st];
        if (!(msg > 0.0)) msg = 0.0;
      } else {
        msg = 0.0;
        for (int t = 0; t < K; ++t) {
          double v = pn[s * K + t];
          if (!(v > 0.0)) v = 0.0;
          msg = __dadd_rn(msg, __dmul_rn(v, L[n][t]));
        }
      }
      if (p >= 0) L[p][s] = __dmul_rn(L[p][s], msg);
      else top = __dmul_rn(top, msg);
    }
  }
  joint[idx] = top;
}

// Two-level alias rows of 512 columns (elx_ptx.cuh) written straight into the stage images: integer masses
// round(p_o / sum * 2^47) (residual to the largest); main entry = [q : 15][alias : 9][0 : 8], escape entry dres32 =
// [thr : 23][alias : 9].  One thread per row (g, c, i).
__global__ void k1_alias_duo(int64_t rows, int K, int stage_bytes, const double *__restrict__ joint, uint8_t *__restrict__ ftab,
                             uint32_t *__restrict__ dres32) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  constexpr int KP = DUO_COLS, TB = 15, AB = 9, W = 47, SH = W - 24;
  int64_t m[KP], q[KP], R[KP];
  int16_t alias[KP], small[KP], large[KP];
  const double *p = joint + r * KP;
  double sum = 0.0;
  for (int j = 0; j < KP; ++j) sum += p[j] > 0.0 ? p[j] : 0.0;
  int64_t tot = 0;
  int big = 0;
  for (int j = 0; j < KP; ++j) {
    const double v = p[j] > 0.0 ? p[j] : 0.0;
    m[j] = sum > 0.0 ? (int64_t)llrint(v / sum * (double)((int64_t)1 << W)) : (j == 0 ? ((int64_t)1 << W) : 0);
    tot += m[j];
    if (m[j] > m[big]) big = j;
  }
  m[big] += ((int64_t)1 << W) - tot;
  int64_t msum = 0;
  for (int j = 0; j < KP; ++j) {
    const int64_t P = m[j], M = P >> SH;
    R[j] = P - (M << SH);
    m[j] = M;
    msum += M;
  }
  const int64_t excess = msum - (((int64_t)1 << TB) - 1) * KP;  // 1..KP, and the largest outcome holds at least 2^24 / KP cells
  m[big] -= excess;
  R[big] += excess << SH;
  uint32_t *e32 = reinterpret_cast<uint32_t *>(ftab + (r / K) * (int64_t)stage_bytes + 32) + (r % K) * KP;
  vose_exact<KP>(m, ((int64_t)1 << TB) - 1, q, alias, small, large);
  for (int j = 0; j < KP; ++j) e32[j] = alias2_main_entry<TB, AB>(j, q[j], alias[j]);
  vose_exact<KP>(R, (int64_t)1 << SH, q, alias, small, large);
  for (int j = 0; j < KP; ++j) dres32[r * KP + j] = alias2_escape_entry<AB>(j, q[j], alias[j]);
}

// Stage header (32 bytes in front of the rows of (record g, component c)): node0, slot holding the root's states; out[k] >= 0:
// leaf row; 0x80000000 | slot: internal frontier node; 0xffffffff: absent.  (Slot NUMBERS: the slot stride is the launch's.)
__global__ void k_build_duo_hdr(int G, int C, int stage_bytes, int threads, const DuoRecDev *__restrict__ recs, uint8_t *__restrict__ ftab) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)G * C) return;
  const DuoRecDev r = recs[i / C];
  (void)threads;
  uint32_t o[2];
  for (int k = 0; k < 2; ++k)
    o[k] = r.leaf[k] >= 0 ? (uint32_t)r.leaf[k] : (r.dst[k] >= 0 ? (0x80000000u | (uint32_t)r.dst[k]) : 0xffffffffu);
  uint4 *h = reinterpret_cast<uint4 *>(ftab + i * (int64_t)stage_bytes);
  h[0] = make_uint4(r.node0, (uint32_t)r.src, __umulhi(PHILOX_M1, r.node0), PHILOX_M1 * r.node0);  // .z, .w: the record's share of the first Philox round (philox3_split)
  h[1] = make_uint4(o[0], o[1], 0xffffffffu, 0xffffffffu);
}

// ---------------------------------------------------------------------------------------------
// Bucketing passes.  u = s - base0 indexes the sites from the start of the first superblock of the call; sites before
// the call's range are only counted (their ranks shift the ranks of the sites behind them).
// ---------------------------------------------------------------------------------------------
struct DuoBucketArgs {
  SimArgs s;
  uint64_t base0;
  int64_t M;   // sites from base0 to the end of the range
  int64_t u0;  // first site of the range (site0 - base0)
  int NSB;     // superblocks touched
  int sb;      // log2 sites per superblock
  int nbatch;  // batches of consecutive superblocks
  int batch_sb[DUO_MAX_BATCH + 1];  // first superblock (0-based in the call) of every batch
  int *tile_begin;                  // [nbatch + 1] first tile of every batch (device)
  uint8_t *comp8, *root8, *rootb;
  uint32_t *pos32, *site32, *hist, *base, *total, *r0;
  DuoSeg *seg;
  DuoTile *tiles;
  int *ntiles;
  int tile_groups, max_tiles;
};

template <int ROUNDS>
__global__ void __launch_bounds__(256) k_duo_comps(DuoBucketArgs b) {
  const SimArgs &a = b.s;
  __shared__ uint32_t hist[DUO_MAX_C];
  __shared__ uint32_t before[DUO_MAX_C];  // sites of the tile in front of the call's range
  hist[threadIdx.x] = 0;
  before[threadIdx.x] = 0;
  __syncthreads();
  const int64_t t0 = (int64_t)blockIdx.x * DUO_TILE;
  for (int k = 0; k < DUO_TILE / 256; ++k) {
    const int64_t u = t0 + k * 256 + threadIdx.x;
    if (u < b.M) {
      const uint64_t s = b.base0 + (uint64_t)u;
      const uint32_t cr = draw_comp_root<ROUNDS>(a.cumw, a.cumpi, a.C, a.K, a.N, a.is_mixture, a.error_flag, s, (uint32_t)a.seed,
                                                 (uint32_t)(a.seed >> 32));
      const uint32_t c = cr & 0xffffu;
      b.comp8[u] = (uint8_t)c;
      b.root8[u] = (uint8_t)(cr >> 16);
      atomicAdd(&hist[c], 1u);
      if (u < b.u0) atomicAdd(&before[c], 1u);
      else if (a.comps) a.comps[u - b.u0] = (int32_t)c;
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < a.C) {
    b.hist[(int64_t)blockIdx.x * a.C + threadIdx.x] = hist[threadIdx.x];
    if (before[threadIdx.x]) atomicAdd(&b.r0[threadIdx.x], before[threadIdx.x]);  // only tiles of the first superblock have any
  }
}

// base[tile][c] = sites of component c in the earlier tiles of the same superblock; total[j][c] = all of superblock j.
// One warp per (superblock, component): 32 tiles per step, a shuffle scan inside the step.
__global__ void k_duo_scan(DuoBucketArgs b) {
  const int C = b.s.C, lane = threadIdx.x & 31;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (wid >= (int64_t)b.NSB * C) return;
  const int j = (int)(wid / C), c = (int)(wid % C);
  const int64_t tiles_per_sb = ((int64_t)1 << b.sb) / DUO_TILE;
  const int64_t ntile = (b.M + DUO_TILE - 1) / DUO_TILE;
  const int64_t t_first = (int64_t)j * tiles_per_sb, t_last = min(t_first + tiles_per_sb, ntile);
  uint32_t carry = 0;
  for (int64_t t = t_first; t < t_last; t += 32) {
    const int64_t tt = t + lane;
    const uint32_t n = tt < t_last ? b.hist[tt * C + c] : 0u;
    uint32_t x = n;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (tt < t_last) b.base[tt * C + c] = carry + x - n;
    carry += __shfl_sync(0xffffffffu, x, 31);
  }
  if (lane == 0) b.total[(int64_t)j * C + c] = carry;
}

// Segments batch by batch, component-major inside a batch (concurrent CTAs then stream the same component's rows), and the
// CTA tiles; tile_begin[b] = first tile of batch b.
__global__ void k_duo_layout(DuoBucketArgs b) {
  // one block: thread 0 lays the segments out (a few hundred at most), then all threads write the tiles of their segments
  const int C = b.s.C, nseg = b.NSB * C;
  extern __shared__ int s_first[];  // [nseg] first tile of the segment with order index e (batch-major, component, superblock)
  if (threadIdx.x == 0) {
    uint64_t slot = 0;
    int nt = 0, e = 0;
    for (int bt = 0; bt < b.nbatch; ++bt) {
      b.tile_begin[bt] = nt < b.max_tiles ? nt : b.max_tiles;
      for (int c = 0; c < C; ++c)
        for (int j = b.batch_sb[bt]; j < b.batch_sb[bt + 1]; ++j, ++e) {
          const uint32_t r0 = j == 0 ? b.r0[c] : 0u, r1 = b.total[(int64_t)j * C + c];
          DuoSeg sg;
          sg.q0 = r0 >> 4;
          sg.ngroups = r1 > r0 ? ((r1 + 15u) >> 4) - sg.q0 : 0u;
          sg.slot_off = slot;
          b.seg[(int64_t)j * C + c] = sg;
          s_first[e] = nt;
          nt += (int)((sg.ngroups + (uint32_t)b.tile_groups - 1) / (uint32_t)b.tile_groups);
          slot += (uint64_t)sg.ngroups * 16u;
        }
    }
    b.tile_begin[b.nbatch] = nt < b.max_tiles ? nt : b.max_tiles;
    *b.ntiles = nt < b.max_tiles ? nt : b.max_tiles;
    if (nt > b.max_tiles) atomicExch(b.s.error_flag, -21);
  }
  __syncthreads();
  const uint64_t sb0 = b.base0 >> b.sb;
  for (int e = threadIdx.x; e < nseg; e += blockDim.x) {
    // order index -> (batch, component, superblock)
    int bt = 0, rem = e;
    while (bt + 1 < b.nbatch && rem >= (b.batch_sb[bt + 1] - b.batch_sb[bt]) * C) { rem -= (b.batch_sb[bt + 1] - b.batch_sb[bt]) * C; ++bt; }
    const int nsb = b.batch_sb[bt + 1] - b.batch_sb[bt], c = rem / nsb, j = b.batch_sb[bt] + rem % nsb;
    const DuoSeg sg = b.seg[(int64_t)j * C + c];
    int t = s_first[e];
    for (uint32_t g = 0; g < sg.ngroups; g += (uint32_t)b.tile_groups, ++t) {
      if (t >= b.max_tiles) break;
      DuoTile tile;
      tile.comp = (uint32_t)c; tile.q_first = sg.q0 + g; tile.ngroups = min((uint32_t)b.tile_groups, sg.ngroups - g);
      tile.sb_lo = (uint32_t)(sb0 + j); tile.sb_hi = (uint32_t)((sb0 + j) >> 32); tile.pad = 0;
      tile.slot_off = sg.slot_off + (uint64_t)g * 16u;
      b.tiles[t] = tile;
    }
  }
}

// Stable rank of every site inside its (superblock, component) -> bucket slot; writes the slot's root state (x 8: the row
// offset form the kernels keep states in) and site, and pos32[site] for the way back.
__global__ void __launch_bounds__(256) k_duo_scatter(DuoBucketArgs b) {
  const int C = b.s.C;
  __shared__ uint32_t run[DUO_MAX_C];
  __shared__ uint16_t wcnt[8][DUO_MAX_C];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int64_t t0 = (int64_t)blockIdx.x * DUO_TILE;
  const int j = (int)(t0 >> b.sb);
  run[tid] = tid < C ? b.base[(int64_t)blockIdx.x * C + tid] : 0u;
  for (int k = 0; k < DUO_TILE / 256; ++k) {
    for (int x = 0; x < 8; ++x) wcnt[x][tid] = 0;
    __syncthreads();
    const int64_t u = t0 + k * 256 + tid;
    const bool valid = u < b.M;
    const uint32_t c = valid ? b.comp8[u] : 0xffffu;
    const uint32_t mask = __match_any_sync(0xffffffffu, c);
    const uint32_t lrank = __popc(mask & ((1u << lane) - 1u));
    if (valid && lrank == 0) wcnt[w][c] = (uint16_t)__popc(mask);
    __syncthreads();
    if (valid) {
      uint32_t r = run[c] + lrank;
      for (int x = 0; x < w; ++x) r += wcnt[x][c];
      const DuoSeg sg = b.seg[(int64_t)j * C + c];
      if (r >= sg.q0 * 16u && r - sg.q0 * 16u < sg.ngroups * 16u) {  // (a segment without sites in the range has no slots)
        const uint64_t pos = sg.slot_off + (r - sg.q0 * 16u);
        b.rootb[pos] = (uint8_t)(b.root8[u] * 8u);
        b.site32[pos] = (uint32_t)u;
        if (u >= b.u0) b.pos32[u - b.u0] = (uint32_t)pos;
      }
    }
    __syncthreads();
    if (tid < C) {
      uint32_t add = 0;
      for (int x = 0; x < 8; ++x) add += wcnt[x][tid];
      run[tid] += add;
    }
  }
}

// Back to site order: out[t][s - site0] = tmp[t][pos32[s - site0]].  A warp reads 32 consecutive sites of one row (runs of
// consecutive bucket slots per component) and writes 32 consecutive bytes.
__global__ void __launch_bounds__(256) k_duo_unpermute(const uint8_t *__restrict__ tmp, int64_t tmp_pitch, const uint32_t *__restrict__ pos32,
                                                      int64_t nsites, int T, int rows_per_block, uint8_t *__restrict__ out, int64_t out_pitch) {
  const int64_t col = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (col >= nsites) return;
  const uint32_t pos = pos32[col];
  const int t0 = blockIdx.y * rows_per_block, t1 = min(T, t0 + rows_per_block);
  const uint8_t *src = tmp + (int64_t)t0 * tmp_pitch + pos;
  uint8_t *dst = out + (int64_t)t0 * out_pitch + col;
#pragma unroll 4
  for (int t = t0; t < t1; ++t) {
    *dst = *src;
    src += tmp_pitch;
    dst += out_pitch;
  }
}

// The same through shared memory (the hot version).  A block owns UNP_TS consecutive sites (aligned in bucketing index
// space, so that a block never straddles a superblock): the sites of one component are then ONE run of consecutive bucket
// slots.  Per row the block copies the runs (<= 1536 aligned words) into shared memory with coalesced 32-bit loads, then every
// thread picks the bytes of its 16 sites and writes them with one 128-bit store: 1 byte read + 1 byte written per cell.
constexpr int UNP_TS = 4096;
constexpr int UNP_WORDS = UNP_TS / 4 + 2 * DUO_MAX_C;  // every run wastes at most two words (its unaligned ends)
constexpr int UNP_R = 4;                               // rows per step
__global__ void __launch_bounds__(256, 3) k_duo_unpermute_smem(const uint8_t *__restrict__ tmp, int64_t tmp_pitch, const uint32_t *__restrict__ pos32,
                                                           const uint8_t *__restrict__ comp8, int64_t u0, int64_t ua, int64_t ub, int C, int T,
                                                           int rows_per_block, uint8_t *__restrict__ out, int64_t out_pitch) {
  __shared__ uint32_t s_min[DUO_MAX_C], s_cnt[DUO_MAX_C], s_woff[DUO_MAX_C + 1];
  __shared__ __align__(16) uint32_t s_buf[UNP_R][UNP_WORDS];
  const int tid = threadIdx.x;
  const int64_t tile0 = (ua / UNP_TS + blockIdx.x) * (int64_t)UNP_TS;
  s_min[tid] = 0xffffffffu;
  s_cnt[tid] = 0;
  __syncthreads();
  uint32_t cc[16], pp[16];
  const int64_t ubase = tile0 + tid * 16;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int64_t u = ubase + i;
    cc[i] = 0xffffffffu;
    pp[i] = 0;
    if (u >= ua && u < ub) {
      cc[i] = comp8[u];
      pp[i] = pos32[u - u0];
      atomicMin(&s_min[cc[i]], pp[i]);
      atomicAdd(&s_cnt[cc[i]], 1u);
    }
  }
  __syncthreads();
  if (tid < 32) {  // exclusive scan of the runs' word counts over the components (one warp, C <= 256)
    uint32_t carry = 0;
    for (int c0 = 0; c0 < C; c0 += 32) {
      const int c = c0 + tid;
      uint32_t n = 0;
      if (c < C && s_cnt[c]) n = (s_cnt[c] + (s_min[c] & 3u) + 3u) >> 2;
      uint32_t x = n;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (tid >= o) x += y;
      }
      if (c < C) s_woff[c] = carry + x - n;
      carry += __shfl_sync(0xffffffffu, x, 31);
    }
    if (tid == 0) s_woff[C] = carry;
  }
  __syncthreads();
  const uint32_t W = s_woff[C];
  uint16_t sidx[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) sidx[i] = cc[i] == 0xffffffffu ? (uint16_t)0 : (uint16_t)(s_woff[cc[i]] * 4u + (pp[i] - (s_min[cc[i]] & ~3u)));
  constexpr int NW = UNP_WORDS / 256;
  uint32_t srcoff[NW];
#pragma unroll
  for (int k = 0; k < NW; ++k) {
    const uint32_t w = (uint32_t)tid + 256u * k;
    srcoff[k] = 0;
    if (w < W) {
      int lo = 0, hi = C - 1;  // last component whose run starts at or before word w (empty runs share their successor's offset)
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_woff[mid] <= w) lo = mid; else hi = mid - 1;
      }
      srcoff[k] = (s_min[lo] & ~3u) + 4u * (w - s_woff[lo]);
    }
  }
  const int64_t col = ubase - u0;
  const bool any = ubase + 16 > ua && ubase < ub;
  const bool whole = ubase >= ua && ubase + 16 <= ub;
  const bool vec = whole && ((out_pitch & 15) == 0) && (((reinterpret_cast<uintptr_t>(out) + (uintptr_t)col) & 15) == 0);
  const int t0 = blockIdx.y * rows_per_block, t1 = min(T, t0 + rows_per_block);
  const uint8_t *src = tmp + (int64_t)t0 * tmp_pitch;
  uint8_t *dst = out + (int64_t)t0 * out_pitch + col;
  // UNP_R rows per step; the loads of the next step are in flight (in registers) while this one is gathered and stored:
  // a row-at-a-time loop is bound by the latency of its loads (2.4 TB/s measured), not by HBM
  uint32_t pre[UNP_R][NW];
  auto prefetch = [&](int t) {
#pragma unroll
    for (int r = 0; r < UNP_R; ++r)
#pragma unroll
      for (int k = 0; k < NW; ++k) {
        const uint32_t w = (uint32_t)tid + 256u * k;
        pre[r][k] = 0;
        if (w < W && t + r < t1) pre[r][k] = __ldcs(reinterpret_cast<const uint32_t *>(src + (int64_t)r * tmp_pitch + srcoff[k]));
      }
  };
  prefetch(t0);
  for (int t = t0; t < t1; t += UNP_R) {
#pragma unroll
    for (int r = 0; r < UNP_R; ++r)
#pragma unroll
      for (int k = 0; k < NW; ++k) {
        const uint32_t w = (uint32_t)tid + 256u * k;
        if (w < W) s_buf[r][w] = pre[r][k];
      }
    __syncthreads();
    src += (int64_t)UNP_R * tmp_pitch;
    if (t + UNP_R < t1) prefetch(t + UNP_R);
#pragma unroll
    for (int r = 0; r < UNP_R; ++r) {
      if (t + r < t1) {
        const uint8_t *bb = reinterpret_cast<const uint8_t *>(s_buf[r]);
        if (vec) {
          uint32_t v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            v[q] = (uint32_t)bb[sidx[4 * q]] | ((uint32_t)bb[sidx[4 * q + 1]] << 8) | ((uint32_t)bb[sidx[4 * q + 2]] << 16) |
                   ((uint32_t)bb[sidx[4 * q + 3]] << 24);
          stg_cs_v4(dst, make_uint4(v[0], v[1], v[2], v[3]));
        } else if (any) {
#pragma unroll 1
          for (int i = 0; i < 16; ++i)
            if (ubase + i >= ua && ubase + i < ub) dst[i] = bb[sidx[i]];
        }
      }
      dst += out_pitch;
    }
    __syncthreads();
  }
}

// The hot version: asynchronous copies, no load passes through a register, so enough bytes are in flight to cover the HBM
// latency (the register version above stops at 3.1 TB/s).  The runs of a row are brought into shared memory as 16-byte aligned
// supersets, UNP_R rows per stage, two stages in flight:
//   few components (TMA = true): one cp.async.bulk per run and row, issued by one warp, behind full mbarriers;
//   many components (runs of a few dozen bytes, where the TMA's per-copy cost dominates): 16-byte cp.async.cg per thread
//   with commit groups.
// The gather maps lane l to 4 consecutive sites of each quarter of the block, so that the byte reads of a warp fall into
// neighbouring words (few bank conflicts) and its 32-bit stores are contiguous.  THREADS = 128 (2048 sites per block, 17 KB of
// shared memory, 8 K registers) fits onto an SM NEXT TO a CTA of the simulation kernel: that is how the pass hides under the
// next batch's simulation.  Row buffer = TS + 32 C bytes: C <= UNP_ASYNC_MAX_C.
constexpr int UNP_ASYNC_MAX_C = 128;
__device__ __forceinline__ uint32_t unp_lds_u8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
constexpr int UNP_MAX_CHUNKS = 3;  // 16-byte chunks per thread and row of the cp.async form: (TS + 32 C) / 16 / THREADS
template <int THREADS, bool TMA>
__global__ void __launch_bounds__(THREADS) k_duo_unpermute_async(const uint8_t *__restrict__ tmp, int64_t tmp_pitch, const uint32_t *__restrict__ pos32,
                                                                const uint8_t *__restrict__ comp8, int64_t u0, int64_t ua, int64_t ub, int C, int T,
                                                                int rows_per_block, int rowb, uint8_t *__restrict__ out, int64_t out_pitch, int lean) {
  constexpr int TS = THREADS * 16, QS = THREADS * 4;
  extern __shared__ __align__(128) uint8_t dsm[];
  __shared__ uint32_t s_min[UNP_ASYNC_MAX_C], s_cnt[UNP_ASYNC_MAX_C], s_off[UNP_ASYNC_MAX_C + 1];
  __shared__ __align__(8) uint64_t s_bar[2];
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t tile0 = (ua / TS + blockIdx.x) * (int64_t)TS;
  for (int c = tid; c < UNP_ASYNC_MAX_C; c += THREADS) { s_min[c] = 0xffffffffu; s_cnt[c] = 0; }
  if (TMA && tid == 0) {
    mbar_init(smem_u32(&s_bar[0]), 1);
    mbar_init(smem_u32(&s_bar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  __syncthreads();
  // site of (quarter k, byte i): tile0 + QS k + 4 tid + i
  uint32_t cc[16], pp[16];
#pragma unroll
  for (int x = 0; x < 16; ++x) {
    const int64_t u = tile0 + QS * (x >> 2) + 4 * tid + (x & 3);
    cc[x] = 0xffffffffu;
    pp[x] = 0;
    if (u >= ua && u < ub) {
      cc[x] = comp8[u];
      pp[x] = pos32[u - u0];
      atomicMin(&s_min[cc[x]], pp[x]);
      atomicAdd(&s_cnt[cc[x]], 1u);
    }
  }
  __syncthreads();
  if (tid < 32) {  // byte offsets of the runs' aligned supersets in a row buffer
    uint32_t carry = 0;
    for (int c0 = 0; c0 < C; c0 += 32) {
      const int c = c0 + tid;
      uint32_t n = 0;
      if (c < C && s_cnt[c]) n = ((s_min[c] & 15u) + s_cnt[c] + 15u) & ~15u;
      uint32_t x = n;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (tid >= o) x += y;
      }
      if (c < C) s_off[c] = carry + x - n;
      carry += __shfl_sync(0xffffffffu, x, 31);
    }
    if (tid == 0) s_off[C] = carry;
  }
  __syncthreads();
  const uint32_t row_bytes = s_off[C];
  uint16_t sidx[16];
#pragma unroll
  for (int x = 0; x < 16; ++x) sidx[x] = cc[x] == 0xffffffffu ? (uint16_t)0 : (uint16_t)(s_off[cc[x]] + (pp[x] - (s_min[cc[x]] & ~15u)));
  const int t0 = blockIdx.y * rows_per_block, t1 = min(T, t0 + rows_per_block);
  const int ngroups = (t1 - t0 + UNP_R - 1) / UNP_R;
  const uint32_t stage_bytes = (uint32_t)(UNP_R * rowb);
  const uint32_t dsm_u32 = smem_u32(dsm);
  // cp.async form: this thread's 16-byte chunks of a row (chunk x of the row buffer <- tmp row + csrc[x])
  uint32_t csrc[UNP_MAX_CHUNKS];
  if (!TMA) {
#pragma unroll
    for (int k = 0; k < UNP_MAX_CHUNKS; ++k) {
      const uint32_t x = (uint32_t)tid + (uint32_t)THREADS * k;
      csrc[k] = 0xffffffffu;
      if (x * 16u < row_bytes) {
        int lo = 0, hi = C - 1;  // last component whose run starts at or before chunk x (empty runs share their successor's offset)
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (s_off[mid] <= x * 16u) lo = mid; else hi = mid - 1;
        }
        csrc[k] = (s_min[lo] & ~15u) + (x * 16u - s_off[lo]);
      }
    }
  }
  auto issue = [&](int g) {  // the rows of group g into stage g & 1
    const int s = g & 1, ta = t0 + g * UNP_R, nr = min(UNP_R, t1 - ta);
    if (TMA) {  // (warp 0 only)
      const uint32_t bar = smem_u32(&s_bar[s]);
      if (lane == 0) mbar_expect_tx(bar, row_bytes * (uint32_t)nr);
      __syncwarp();
      for (int c = lane; c < C; c += 32) {
        if (!s_cnt[c]) continue;
        const uint32_t a16 = s_min[c] & ~15u, nb = ((s_min[c] & 15u) + s_cnt[c] + 15u) & ~15u;
        for (int r = 0; r < nr; ++r)
          tma_bulk_g2s(dsm_u32 + (uint32_t)s * stage_bytes + (uint32_t)r * (uint32_t)rowb + s_off[c], tmp + (int64_t)(ta + r) * tmp_pitch + a16, nb, bar);
      }
    } else {  // (every thread)
#pragma unroll
      for (int k = 0; k < UNP_MAX_CHUNKS; ++k) {
        if (csrc[k] == 0xffffffffu) continue;
        const uint32_t dst = dsm_u32 + (uint32_t)s * stage_bytes + ((uint32_t)tid + (uint32_t)THREADS * k) * 16u;
        for (int r = 0; r < nr; ++r)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)r * (uint32_t)rowb),
                       "l"(tmp + (int64_t)(ta + r) * tmp_pitch + csrc[k]) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  };
  if (row_bytes && (!TMA || tid < 32)) {
    issue(0);
    if (ngroups > 1) issue(1);
    else if (!TMA) asm volatile("cp.async.commit_group;" ::: "memory");  // (keeps the group count uniform)
  }
  const int64_t colq = tile0 + 4 * tid - u0;  // column of this thread's first site of quarter 0
  bool vec[4], anyq[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t ub0 = tile0 + QS * k + 4 * tid;
    anyq[k] = ub0 + 4 > ua && ub0 < ub;
    vec[k] = ub0 >= ua && ub0 + 4 <= ub && ((out_pitch & 3) == 0) && (((reinterpret_cast<uintptr_t>(out) + (uintptr_t)(colq + QS * k)) & 3) == 0);
  }
  uint8_t *dst = out + (int64_t)t0 * out_pitch + colq;
  const bool all_vec = __syncthreads_and(lean && vec[0] && vec[1] && vec[2] && vec[3]) != 0;
  uint32_t sa[16];
#pragma unroll
  for (int x = 0; x < 16; ++x) sa[x] = dsm_u32 + (uint32_t)sidx[x];
  for (int g = 0; g < ngroups; ++g) {
    const int s = g & 1, ta = t0 + g * UNP_R, nr = min(UNP_R, t1 - ta);
    if (row_bytes) {
      if (TMA) {
        mbar_wait(smem_u32(&s_bar[s]), (uint32_t)((g >> 1) & 1));
      } else {
        asm volatile("cp.async.wait_group 1;" ::: "memory");  // all but the newest group (the next stage) have landed
        __syncthreads();
      }
    }
    const uint8_t *sb = dsm + (size_t)s * stage_bytes;
    if (all_vec) {
      // the common case, free of per-quarter tests: 16 byte loads at precomputed shared-memory addresses plus a uniform row
      // offset, packed on the FMA pipe, four contiguous 32-bit streaming stores (2.1 instructions per cell; the general path
      // below -- first / last block of a range, unaligned rows -- issues 8)
      const uint32_t so = (uint32_t)s * stage_bytes;
#pragma unroll
      for (int r = 0; r < UNP_R; ++r) {
        if (r < nr) {
          const uint32_t ro = so + (uint32_t)r * (uint32_t)rowb;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t b0 = unp_lds_u8(sa[4 * k] + ro), b1 = unp_lds_u8(sa[4 * k + 1] + ro), b2 = unp_lds_u8(sa[4 * k + 2] + ro),
                           b3 = unp_lds_u8(sa[4 * k + 3] + ro);
            __stcs(reinterpret_cast<uint32_t *>(dst + QS * k), b3 * 0x1000000u + (b2 * 0x10000u + (b1 * 0x100u + b0)));
          }
        }
        dst += out_pitch;
      }
    } else
#pragma unroll
    for (int r = 0; r < UNP_R; ++r) {
      if (r < nr) {
        const uint8_t *bb = sb + (size_t)r * rowb;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (vec[k]) {
            const uint32_t v = (uint32_t)bb[sidx[4 * k]] | ((uint32_t)bb[sidx[4 * k + 1]] << 8) | ((uint32_t)bb[sidx[4 * k + 2]] << 16) |
                               ((uint32_t)bb[sidx[4 * k + 3]] << 24);
            __stcs(reinterpret_cast<uint32_t *>(dst + QS * k), v);
          } else if (anyq[k]) {
#pragma unroll 1
            for (int i = 0; i < 4; ++i) {
              const int64_t u = tile0 + QS * k + 4 * tid + i;
              if (u >= ua && u < ub) dst[QS * k + i] = bb[sidx[4 * k + i]];
            }
          }
        }
      }
      dst += out_pitch;
    }
    __syncthreads();  // every thread is done with stage s
    if (row_bytes && (!TMA || tid < 32)) {
      if (g + 2 < ngroups) issue(g + 2);
      else if (!TMA) asm volatile("cp.async.commit_group;" ::: "memory");
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K2 duo mode, generic: tables through L1/L2, one thread per bucket slot (the cross-check of the tiled kernel).
// ---------------------------------------------------------------------------------------------
template <int ROUNDS>
__device__ __forceinline__ uint32_t duo_draw24(uint32_t q, uint32_t comp, uint32_t sb_lo, uint32_t sb_hi, int d, uint32_t node0, uint2 key) {
  const int w0 = 3 * (d >> 2), wa = (d & 3) < 3 ? w0 + (d & 3) : w0, wb = (d & 3) < 3 ? wa : w0 + 2;
  const uint32_t cx = q | (comp << 20);
  uint32_t X[12];
  for (int call = wa >> 2; call <= wb >> 2; ++call) {  // the one or two Philox calls that hold the draw's words
    const uint4 x = philox4x32<ROUNDS>(make_uint4(cx, sb_lo, node0, (sb_hi << 8) | (4u * (uint32_t)call)), key);
    X[4 * call] = x.x; X[4 * call + 1] = x.y; X[4 * call + 2] = x.z; X[4 * call + 3] = x.w;
  }
  return draw24_from_words(X, d);
}

template <int ROUNDS>
__global__ void __launch_bounds__(256) k2_duo_generic(DuoArgs fa) {
  const SimArgs &a = fa.s;
  const int tile_first = fa.tile_begin[fa.batch];
  if ((int)blockIdx.x >= fa.tile_begin[fa.batch + 1] - tile_first) return;
  const DuoTile tile = fa.tiles[tile_first + blockIdx.x];
  const uint32_t idx = blockIdx.y * 256u + threadIdx.x;  // slot of the tile
  if ((idx >> 4) >= tile.ngroups) return;
  const uint32_t q = tile.q_first + (idx >> 4);
  const int d = (int)(idx & 15u);
  const uint64_t pos = tile.slot_off + idx;
  const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
  const int K = a.K, c = (int)tile.comp;
  uint8_t slot[DUO_MAX_SLOTS];
  slot[0] = (uint8_t)(fa.rootb[pos] >> 3);
  for (int g = 0; g < fa.G; ++g) {
    const DuoRecDev r = fa.recs[g];
    const uint32_t x24 = duo_draw24<ROUNDS>(q, tile.comp, tile.sb_lo, tile.sb_hi, d, r.node0, key);
    const uint32_t j = x24 & 511u, t = x24 >> 9;
    const uint32_t row = slot[r.src];
    const uint32_t e = __ldg(reinterpret_cast<const uint32_t *>(fa.ftab + ((int64_t)g * a.C + c) * fa.stage_bytes + 32) + row * DUO_COLS + j);
    uint32_t o;
    if (t != 0u) {
      o = x24 < (e >> 8) ? j : ((e >> 8) & 511u);
    } else {  // escape level
      const uint64_t s = fa.base0 + fa.site32[pos];
      const uint4 z = philox4x32<ROUNDS>(make_uint4((uint32_t)s, (uint32_t)(s >> 32), r.node0, TAG_REFINE), key);
      const uint32_t er = __ldg(fa.dres32 + ((((int64_t)g * a.C + c) * K + row) << 9) + (z.x & 511u));
      o = (z.x >> 9) < (er >> 9) ? (z.x & 511u) : (er & 511u);
    }
    const uint32_t st[2] = {o % (uint32_t)K, o / (uint32_t)K};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (r.leaf[k] >= 0) fa.tmp[(int64_t)r.leaf[k] * fa.tmp_pitch + pos] = (uint8_t)st[k];
      else if (r.dst[k] >= 0) slot[r.dst[k]] = (uint8_t)st[k];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K2 duo mode, tiled (the hot kernel).  One CTA = one tile of one component: THREADS consumer threads x 16 bucket slots
// each + one producer warp streaming the records' (header + K x 2 KB rows) through a shared-memory ring with TMA bulk
// copies (full / empty mbarriers, per-warp release, no CTA barrier in the loop -- the protocol of k2_quad_tiled).  A slot
// holds one byte per site: state x 8, so that the byte moved to bits 8..15 is the row's byte offset (state x 2048).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void stg_v4(void *p, uint4 v) {
  asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint32_t duo_lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
template <uint32_t SEL>
__device__ __forceinline__ uint32_t duo_prmt(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "n"(SEL));
  return d;
}
__device__ __forceinline__ uint32_t duo_mad(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

template <int ROUNDS, bool FAST>
__device__ __forceinline__ void duo_records(const DuoArgs &fa, const DuoTile &tile, const uint32_t full_bar0, const uint32_t empty_bar0,
                                            const uint32_t stage_smem, const uint32_t slot_base, uint8_t *tmp_base, const bool active,
                                            const uint32_t q, const uint64_t pos0, const int lane) {
  const SimArgs &a = fa.s;
  const uint32_t cx = q | (tile.comp << 20), cy = tile.sb_lo, cw = tile.sb_hi << 8;
  const uint32_t stride = (uint32_t)fa.slot_stride;
  const int nstage = fa.nstage;
  const uint32_t mhi = fa.mhi, negk = fa.negk256;
  const uint32_t pitch32 = (uint32_t)fa.tmp_pitch;
  int buf = 0;
  uint32_t phase = 0;
  const PhiloxThreadPart tp = philox_thread_part<ROUNDS>(cx, cw, cw | 4u, cw | 8u, fa.rk);
  for (int g = 0; g < fa.G; ++g) {
    mbar_wait(full_bar0 + 8u * (uint32_t)buf, phase);
    const uint32_t sbuf = stage_smem + (uint32_t)(buf * fa.stage_bytes);
    const uint4 h0 = lds_v4(sbuf);
    const uint4 h1 = lds_v4(sbuf + 16u);
    const uint32_t node0 = h0.x;
    const uint32_t tbl = sbuf + 32u;
#ifdef ELX_DEBUG_BOUNDS
    if (h0.y >= (uint32_t)a.n_slots || node0 >= (uint32_t)a.N) atomicExch(a.error_flag, -8);
    if (h0.z != __umulhi(PHILOX_M1, node0) || h0.w != PHILOX_M1 * node0) atomicExch(a.error_flag, -9);
#endif
    const uint4 rv = lds_v4(slot_base + h0.y * stride);
    const uint32_t Rq[4] = {rv.x, rv.y, rv.z, rv.w};
    uint32_t X[12];
    philox3_split<ROUNDS>(tp, h0.z, h0.w, cy, fa.rk, X);  // = philox4x32(cx, cy, node0, cw | 4 * call), call = 0, 1, 2
    uint32_t P0[4], P1[4];
    uint32_t m = 0xffffffffu;
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) {
      uint32_t v0[4], s1[4], t[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int d = 4 * qq + i;
        const uint32_t w = ELX_DRAW24(X, d);  // [t : 15][j : 9][x : 8]
        // row offset = state x 2048: the slot byte (state x 8) moved to bits 8..15
        const uint32_t ro = i == 0 ? duo_prmt<0x4404u>(Rq[qq], 0u) : i == 1 ? duo_prmt<0x4414u>(Rq[qq], 0u)
                          : i == 2 ? duo_prmt<0x4424u>(Rq[qq], 0u) : duo_prmt<0x4434u>(Rq[qq], 0u);
        const uint32_t off = ((w >> 6) & 0x7fcu) | ro;
#ifdef ELX_DEBUG_BOUNDS
        if (off + 4u > (uint32_t)(fa.stage_bytes - 32)) atomicExch(a.error_flag, -7);
#endif
        const uint32_t e = duo_lds_u32(tbl + off);
        const uint32_t c = min(w, e);
        t[i] = w;
        const uint32_t mo = c & 0x1ff00u;        // o << 8
        s1[i] = __umulhi(mo, mhi);               // o div K
        v0[i] = duo_mad(s1[i], negk, mo);        // (o mod K) << 8
      }
      m = min(min(m, t[0]), min(t[1], min(t[2], t[3])));
      P0[qq] = __byte_perm(__byte_perm(v0[0], v0[1], 0x0051u), __byte_perm(v0[2], v0[3], 0x0051u), 0x5410u);
      P1[qq] = duo_mad(s1[3], 0x1000000u, duo_mad(s1[2], 0x10000u, duo_mad(s1[1], 0x100u, s1[0])));
    }
    if (m < 0x20000u && (FAST || active)) {
      // escape level (2^-15 per draw): some draw has t == 0; its outcome comes from the row's residual masses with 32 fresh
      // bits.  Unrolled with static register names: no calls, no dynamically indexed arrays (elx_fast.cu).
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        const uint32_t w = ELX_DRAW24(X, d);
        if (w < 0x20000u) {
          const uint32_t row8 = (Rq[d >> 2] >> (8 * (d & 3))) & 0xffu;
          const uint64_t sg = fa.base0 + (uint64_t)__ldg(fa.site32 + pos0 + d);
          const uint4 z = philox4x32_rk<ROUNDS>(make_uint4((uint32_t)sg, (uint32_t)(sg >> 32), node0, TAG_REFINE), fa.rk);
          const uint32_t jr = z.x & 0x1ffu;
          const uint32_t er = __ldg(fa.dres32 + (((((int64_t)g * a.C + tile.comp) * a.K + (row8 >> 3)) << 9) | jr));
          const uint32_t o = (z.x >> 9) < (er >> 9) ? jr : (er & 0x1ffu);
          const uint32_t s1 = o / (uint32_t)a.K, s0 = o - s1 * (uint32_t)a.K;
          P0[d >> 2] = (P0[d >> 2] & ~(0xffu << (8 * (d & 3)))) | (s0 << (8 * (d & 3)));
          P1[d >> 2] = (P1[d >> 2] & ~(0xffu << (8 * (d & 3)))) | (s1 << (8 * (d & 3)));
        }
      }
    }
    // this warp is done with the stage buffer (the header is in registers)
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar0 + 8u * (uint32_t)buf);
    if (++buf == nstage) { buf = 0; phase ^= 1u; }

#define ELX_DUO_OUT(PW, code)                                                                                   \
  if ((int32_t)(code) >= 0) {                                                                                   \
    if (FAST || active) stg_v4(tmp_base + (uint64_t)(uint32_t)(code) * (uint64_t)pitch32, make_uint4(PW[0], PW[1], PW[2], PW[3])); /* one IMAD.WIDE: tmp_pitch < 2^32 (duo_simulate) */ \
  } else if ((code) != 0xffffffffu) {                                                                           \
    sts_v4(slot_base + ((code) & 0x7fffffffu) * stride, make_uint4(PW[0] << 3, PW[1] << 3, PW[2] << 3, PW[3] << 3));                   \
  }
    ELX_DUO_OUT(P0, h1.x)
    ELX_DUO_OUT(P1, h1.y)
#undef ELX_DUO_OUT
  }
}

template <int ROUNDS, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT + 32, MINB) k2_duo_tiled(DuoArgs fa) {
  const int THREADS = (int)blockDim.x - 32;
  extern __shared__ __align__(128) uint8_t smem[];
  const int tile_first = fa.tile_begin[fa.batch];
  if ((int)blockIdx.x >= fa.tile_begin[fa.batch + 1] - tile_first) return;
  const int tid = threadIdx.x;
  const int nstage = fa.nstage;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem);
  uint64_t *empty_bar = full_bar + DUO_MAX_NSTAGE;
  const uint32_t stage_smem = smem_u32(smem + 128);
  const uint32_t slots_smem = stage_smem + (uint32_t)(nstage * fa.stage_bytes);
  const DuoTile tile = fa.tiles[tile_first + blockIdx.x];

  // warps without a single group of the tile leave at once (the tile that ends a segment is mostly empty, and big mixtures
  // have many short segments): the ring is released by the warps that stay
  const int live_warps = max(1, min(THREADS / 32, ((int)tile.ngroups + 31) >> 5));
  if (tid == 0) {
    for (int i = 0; i < nstage; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), live_warps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  __syncthreads();

  const int tid0 = __shfl_sync(0xffffffffu, tid, 0);  // (through a shuffle: provably warp-uniform for ptxas)
  if (tid0 < THREADS && (tid0 >> 5) >= live_warps) return;
  if (tid0 >= THREADS) {
    if (tid == THREADS) {  // producer: one lane keeps the ring full with the rows of this tile's component
      const uint8_t *src = fa.ftab + (size_t)tile.comp * fa.stage_bytes;
      const size_t step = (size_t)fa.s.C * fa.stage_bytes;
      int buf = 0;
      uint32_t phase = 1;  // the first pass over the ring finds every buffer free
      for (int g = 0; g < fa.G; ++g) {
        if (g >= nstage) mbar_wait_backoff(smem_u32(&empty_bar[buf]), phase, 256);
        mbar_expect_tx(smem_u32(&full_bar[buf]), (uint32_t)fa.stage_bytes);
        tma_bulk_g2s(stage_smem + (uint32_t)(buf * fa.stage_bytes), src + (size_t)g * step, (uint32_t)fa.stage_bytes,
                     smem_u32(&full_bar[buf]));
        if (++buf == nstage) { buf = 0; phase ^= 1u; }
      }
    }
    return;
  }

  const bool active = (uint32_t)tid < tile.ngroups;
  const uint32_t q = tile.q_first + (uint32_t)tid;
  const uint64_t pos0 = tile.slot_off + (uint64_t)tid * 16u;
  const uint32_t slot_base = slots_smem + (uint32_t)tid * 16u;
  uint4 rv = make_uint4(0u, 0u, 0u, 0u);
  if (active) rv = __ldg(reinterpret_cast<const uint4 *>(fa.rootb + pos0));
  sts_v4(slot_base, rv);  // slot 0 = the root states (x 8)
  uint8_t *tmp_base = fa.tmp + pos0;
  asm volatile("" : "+l"(tmp_base));
  const bool fast = __all_sync(0xffffffffu, active);
  if (fast)
    duo_records<ROUNDS, true>(fa, tile, smem_u32(full_bar), smem_u32(empty_bar), stage_smem, slot_base, tmp_base, true, q, pos0, tid & 31);
  else
    duo_records<ROUNDS, false>(fa, tile, smem_u32(full_bar), smem_u32(empty_bar), stage_smem, slot_base, tmp_base, active, q, pos0, tid & 31);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
#define DUO_CUDA(h, expr)                                                                                       \
  do {                                                                                                          \
    cudaError_t _e = (expr);                                                                                    \
    if (_e != cudaSuccess) return fail(h, ELX_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));      \
  } while (0)

int64_t duo_batch_sites(const elx_handle *h) {
  const DuoTables *dt = static_cast<const DuoTables *>(h->duo);
  if (!dt) return 0;
  return (int64_t)1 << std::min(dt->sb, 22);  // (2^24-site superblocks are walked in quarters: device buffers stay bounded)
}

int duo_groups(const elx_handle *h) {
  const DuoTables *dt = static_cast<const DuoTables *>(h->duo);
  return dt ? dt->G : 0;
}

// CTA size x CTAs per SM; the slot stack must leave room for a ring of at least two stages.  One 512-thread CTA per SM is
// preferred: a record's rows (40 KB) are then read from L2 once per 8192 draws.
static bool duo_geometry(int K, int n_slots, DuoTables *dt) {
  const size_t stage = 32 + (size_t)K * DUO_COLS * 4;
  int want_threads = 0, want_stages = 0, want_minb = 0;
  if (const char *e = getenv("ELX_DUO_THREADS")) want_threads = atoi(e);  // tuning knobs
  if (const char *e = getenv("ELX_DUO_NSTAGE")) want_stages = atoi(e);
  if (const char *e = getenv("ELX_DUO_MINB")) want_minb = atoi(e);
  const int cand[4][2] = {{640, 1}, {512, 1}, {256, 2}, {256, 1}};
  for (int ci = 0; ci < 4; ++ci) {
    const int th = cand[ci][0], minb = cand[ci][1];
    if (want_threads && th != want_threads) continue;
    if (want_minb && minb != want_minb) continue;
    const size_t per_cta = (size_t)226 * 1024 / minb;
    const size_t slots = (size_t)(n_slots > 0 ? n_slots : 1) * th * 16;
    if (128 + (th == 640 ? 3 : 2) * stage + slots > per_cta) continue;  // (the 640 class only with a full ring)
    int nstage = (int)((per_cta - 128 - slots) / stage);
    const int cap = want_stages ? want_stages : 3;
    if (nstage > cap) nstage = cap;
    if (nstage > DUO_MAX_NSTAGE) nstage = DUO_MAX_NSTAGE;
    if (nstage < 2) continue;
    dt->tiled = 1; dt->threads = th; dt->minb = minb; dt->nstage = nstage; dt->stage_bytes = (int)stage;
    dt->smem = 128 + (size_t)nstage * stage + slots;
    if (th == 640) {  // the 512 class next to it
      const size_t slots512 = (size_t)(n_slots > 0 ? n_slots : 1) * 512 * 16;
      int ns = (int)((per_cta - 128 - slots512) / stage);
      if (ns > cap) ns = cap;
      dt->nstage512 = ns; dt->smem512 = 128 + (size_t)ns * stage + slots512;
    }
    return true;
  }
  dt->tiled = 0; dt->threads = 512; dt->minb = 1; dt->nstage = 0; dt->stage_bytes = (int)stage; dt->smem = 0;
  return false;
}

int duo_create(elx_handle *h, const int32_t *parent, const int32_t *leaf_row) {
  if (h->K < 5 || h->K > DUO_MAX_K || h->C > DUO_MAX_C) return 0;
  DuoTables *dt = new DuoTables();
  compile_quad_walk(h->N, parent, leaf_row, dt->prog, 2);
  const int G = dt->G = (int)dt->prog.recs.size();
  if (G == 0 || dt->prog.n_slots > DUO_MAX_SLOTS) { delete dt; return 0; }
  duo_geometry(h->K, dt->prog.n_slots, dt);
  h->duo = dt;
  std::vector<DuoRecDev> dev((size_t)G);
  std::vector<DuoCapDev> caps((size_t)G);
  for (int g = 0; g < G; ++g) {
    const QuadRec &r = dt->prog.recs[g];
    DuoRecDev &d = dev[g];
    memset(&d, 0, sizeof d);
    d.node0 = (uint32_t)r.out[0];
    d.src = (int8_t)r.src;
    for (int k = 0; k < 2; ++k) { d.dst[k] = r.dst[k]; d.leaf[k] = r.leaf[k]; }
    DuoCapDev &c = caps[g];
    memset(&c, 0, sizeof c);
    c.n = r.n_cap < QUAD_MAX_CAP ? r.n_cap : QUAD_MAX_CAP;
    for (int i = 0; i < c.n; ++i) { c.node[i] = r.cap[i].node; c.par[i] = r.cap[i].par; c.fr[i] = r.cap[i].fr; }
  }
  const size_t rows = (size_t)G * h->C * h->K;
  DUO_CUDA(h, cudaMalloc((void **)&dt->d_recs, sizeof(DuoRecDev) * (size_t)G));
  DUO_CUDA(h, cudaMalloc((void **)&dt->d_caps, sizeof(DuoCapDev) * (size_t)G));
  DUO_CUDA(h, cudaMalloc((void **)&dt->d_ftab, (size_t)G * h->C * dt->stage_bytes));
  DUO_CUDA(h, cudaMalloc((void **)&dt->d_dres32, sizeof(uint32_t) * rows * DUO_COLS));
  DUO_CUDA(h, cudaMalloc((void **)&dt->d_ntiles, sizeof(int)));
  DUO_CUDA(h, cudaMalloc((void **)&dt->d_tile_begin, sizeof(int) * (DUO_MAX_BATCH + 1)));
  dt->sb = duo_sb_log2(h->C);
  {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    DUO_CUDA(h, cudaStreamCreateWithPriority(&dt->s2, cudaStreamNonBlocking, hi));
    for (cudaEvent_t &e : dt->ev_main) DUO_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    DUO_CUDA(h, cudaEventCreateWithFlags(&dt->ev_done, cudaEventDisableTiming));
    DUO_CUDA(h, cudaEventCreateWithFlags(&dt->ev_bucket, cudaEventDisableTiming));
    // earlier batches get the higher priority: the block scheduler then finishes them in order (measured without: four
    // concurrent grids share the SMs evenly and all finish at the end), yet the next batch fills the SMs a tail leaves idle
    for (int i = 0; i < DUO_MAX_BATCH; ++i) DUO_CUDA(h, cudaStreamCreateWithPriority(&dt->sm[i], cudaStreamNonBlocking, std::min(lo, hi + 1 + i)));
  }
  DUO_CUDA(h, cudaMemcpy(dt->d_recs, dev.data(), sizeof(DuoRecDev) * (size_t)G, cudaMemcpyHostToDevice));
  DUO_CUDA(h, cudaMemcpy(dt->d_caps, caps.data(), sizeof(DuoCapDev) * (size_t)G, cudaMemcpyHostToDevice));
  return 0;
}

// joint rows -> alias rows inside the stage images; the fp64 joint rows are a temporary of this function, built in slices
// of records so that big mixtures on big trees do not need gigabytes of it
int duo_tables_build(elx_handle *h) {
  DuoTables *dt = static_cast<DuoTables *>(h->duo);
  if (!dt) return 0;
  const int64_t rows_per_rec = (int64_t)h->C * h->K;
  int gstep = (int)std::max<int64_t>(1, ((int64_t)256 << 20) / (rows_per_rec * DUO_COLS * (int64_t)sizeof(double)));
  if (gstep > dt->G) gstep = dt->G;
  double *d_joint = nullptr;
  DUO_CUDA(h, cudaMalloc((void **)&d_joint, sizeof(double) * (size_t)gstep * rows_per_rec * DUO_COLS));
  for (int g0 = 0; g0 < dt->G; g0 += gstep) {
    const int gn = std::min(gstep, dt->G - g0);
    const int64_t ents = (int64_t)gn * rows_per_rec * DUO_COLS;
    k1_duo_joint<<<(unsigned)((ents + 255) / 256), 256, 0, h->stream>>>(gn, h->C, h->K, dt->d_caps + g0, h->d_P, d_joint);
    h->launches++;
    const int64_t rows = (int64_t)gn * rows_per_rec;
    k1_alias_duo<<<(unsigned)((rows + 63) / 64), 64, 0, h->stream>>>(rows, h->K, dt->stage_bytes, d_joint,
                                                                    dt->d_ftab + (size_t)g0 * h->C * dt->stage_bytes,
                                                                    dt->d_dres32 + (size_t)g0 * rows_per_rec * DUO_COLS);
    h->launches++;
  }
  const int64_t n = (int64_t)dt->G * h->C;
  k_build_duo_hdr<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(dt->G, h->C, dt->stage_bytes, dt->threads, dt->d_recs, dt->d_ftab);
  h->launches++;
  DUO_CUDA(h, cudaGetLastError());
  DUO_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaFree(d_joint);
  return 0;
}

void duo_free(elx_handle *h) {
  DuoTables *dt = static_cast<DuoTables *>(h->duo);
  if (!dt) return;
  cudaFree(dt->d_recs); cudaFree(dt->d_caps); cudaFree(dt->d_ftab); cudaFree(dt->d_dres32); cudaFree(dt->d_ntiles);
  cudaFree(dt->d_comp8); cudaFree(dt->d_root8); cudaFree(dt->d_rootb); cudaFree(dt->d_tmp); cudaFree(dt->d_pos32);
  cudaFree(dt->d_site32); cudaFree(dt->d_hist); cudaFree(dt->d_base); cudaFree(dt->d_total); cudaFree(dt->d_r0);
  cudaFree(dt->d_seg); cudaFree(dt->d_tiles); cudaFree(dt->d_tile_begin);
  if (dt->s2) { cudaStreamSynchronize(dt->s2); cudaStreamDestroy(dt->s2); }
  for (cudaEvent_t e : dt->ev_main) if (e) cudaEventDestroy(e);
  if (dt->ev_done) cudaEventDestroy(dt->ev_done);
  if (dt->ev_bucket) cudaEventDestroy(dt->ev_bucket);
  for (cudaStream_t st : dt->sm) if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
  delete dt;
  h->duo = nullptr;
}

int duo_get(elx_handle *h, int32_t *recs, uint32_t *de32, uint32_t *dres32) {
  DuoTables *dt = static_cast<DuoTables *>(h->duo);
  if (!dt) return fail(h, ELX_E_INVALID, "elx_duo_get: this handle has no duo tables");
  if (recs)
    for (int g = 0; g < dt->G; ++g) quad_rec_to_ints(dt->prog.recs[g], recs + (size_t)g * ELX_QUAD_REC_INTS);
  const size_t row_bytes = (size_t)h->K * DUO_COLS * sizeof(uint32_t);
  if (de32)
    DUO_CUDA(h, cudaMemcpy2DAsync(de32, row_bytes, dt->d_ftab + 32, (size_t)dt->stage_bytes, row_bytes, (size_t)dt->G * h->C,
                                  cudaMemcpyDeviceToHost, h->stream));
  if (dres32)
    DUO_CUDA(h, cudaMemcpyAsync(dres32, dt->d_dres32, (size_t)dt->G * h->C * row_bytes, cudaMemcpyDeviceToHost, h->stream));
  DUO_CUDA(h, cudaStreamSynchronize(h->stream));
  return 0;
}

template <typename T>
static cudaError_t duo_grow(T **p, size_t *cap, size_t need) {
  if (need <= *cap && *p) return cudaSuccess;
  cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  const size_t n = need + need / 8 + 256;
  cudaError_t e = cudaMalloc((void **)p, n * sizeof(T));
  if (e == cudaSuccess) *cap = n;
  return e;
}

template <int ROUNDS, int MAXT, int MINB>
static cudaError_t launch_duo_tiled(int threads, const DuoArgs &fa, int max_tiles, size_t smem, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(k2_duo_tiled<ROUNDS, MAXT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k2_duo_tiled<ROUNDS, MAXT, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  k2_duo_tiled<ROUNDS, MAXT, MINB><<<(unsigned)max_tiles, threads + 32, smem, stream>>>(fa);
  return cudaGetLastError();
}

template <int ROUNDS>
static cudaError_t launch_duo_tiled_t(int maxt, int minb, int threads, const DuoArgs &fa, int max_tiles, size_t smem, cudaStream_t stream) {
  if (maxt == 640) return launch_duo_tiled<ROUNDS, 640, 1>(threads, fa, max_tiles, smem, stream);
  if (maxt == 512) return launch_duo_tiled<ROUNDS, 512, 1>(threads, fa, max_tiles, smem, stream);
  if (minb == 2) return launch_duo_tiled<ROUNDS, 256, 2>(threads, fa, max_tiles, smem, stream);
  return launch_duo_tiled<ROUNDS, 256, 1>(threads, fa, max_tiles, smem, stream);
}

// Groups (threads) per CTA tile of this call: any multiple of 32 up to the kernel's class works (the slot stride and the
// shared-memory budget are those of the class).  A tile count just above a multiple of the SM count costs a whole extra
// wave (cfg5: 1.25 M sites in tiles of 512 groups = 156 tiles on 148 SMs), so the tile is sized from the EXPECTED segment
// lengths -- sites of the superblock x component weight -- to minimise  waves x max(tile, 448 threads).
static int duo_tile_groups(const elx_handle *h, const DuoTables *dt, int64_t u0, int64_t M) {
  if (const char *e = getenv("ELX_DUO_TILE")) {  // tuning knob
    const int t = atoi(e);
    if (t >= 32 && t <= dt->threads && t % 32 == 0) return t;
  }
  int ndev_sms = 148;
  cudaDeviceGetAttribute(&ndev_sms, cudaDevAttrMultiProcessorCount, h->device);
  const int conc = ndev_sms * (dt->threads == 256 ? dt->minb : 1);
  double wsum = 0;
  for (double w : h->weights) wsum += w;
  const int NSB = (int)((M + ((int64_t)1 << dt->sb) - 1) >> dt->sb);
  int best = dt->threads;
  double best_cost = 1e300;
  for (int tg = dt->threads; tg >= (dt->threads >= 512 ? 320 : 128); tg -= 32) {
    int64_t nt = 0;
    for (int j = 0; j < NSB; ++j) {
      const int64_t lo = std::max<int64_t>((int64_t)j << dt->sb, u0), hi = std::min<int64_t>(((int64_t)j + 1) << dt->sb, M);
      for (int c = 0; c < h->C; ++c) {
        const double n = (double)(hi - lo) * (wsum > 0 ? h->weights[(size_t)c] / wsum : 1.0 / h->C);
        if (n <= 0) continue;
        const double groups = (n + 4.0 * std::sqrt(n)) / 16.0 + 1.0;
        nt += (int64_t)std::ceil(groups / tg);
      }
    }
    // a CTA's time goes with the busiest sub-partition: ceil(warps / 4); fewer than 4 warps per sub-partition do not saturate it
    const int wps = (tg / 32 + 3) / 4;
    const double eff = dt->threads == 256 ? 1.0 : (wps >= 4 ? 1.0 : wps == 3 ? 0.87 : 0.65);
    const double cost = (double)((nt + conc - 1) / conc) * wps / eff;
    if (cost < best_cost) { best_cost = cost; best = tg; }
  }
  return best;
}

int duo_simulate(elx_handle *h, SimArgs &a, int *used) {
  *used = 0;
  DuoTables *dt = static_cast<DuoTables *>(h->duo);
  if (!dt || a.internal) return 0;  // summed-out nodes have no states: `simulate` (keep internal states) draws node by node
  const int C = h->C;
  const uint64_t base0 = ((uint64_t)a.site0 >> dt->sb) << dt->sb;
  const int64_t u0 = (int64_t)((uint64_t)a.site0 - base0), M = u0 + a.nsites;
  const int NSB = (int)((M + ((int64_t)1 << dt->sb) - 1) >> dt->sb);
  const int64_t ntile_b = (M + DUO_TILE - 1) / DUO_TILE;
  const size_t slots_bound = (size_t)a.nsites + 32u * (size_t)NSB * C + 16;
  if (slots_bound + 256 >= ((size_t)1 << 32) || M >= ((int64_t)1 << 32))  // (tmp_pitch fits 32 bits: the kernels form row offsets with one IMAD.WIDE)
    return fail(h, ELX_E_UNSUPPORTED, "duo mode: more than 2^32 sites in one device call; split the site range");
  if ((size_t)NSB * C * sizeof(int) > 40 * 1024)
    return fail(h, ELX_E_UNSUPPORTED, "duo mode: too many (superblock, component) segments in one device call; split the site range");
  const bool tiled = dt->tiled && !(h->flags & ELX_FLAG_FORCE_GENERIC);
  int tile_groups = duo_tile_groups(h, dt, u0, M);
  const int max_tiles = (int)((slots_bound / 16 + 511) / 512) * 2 + NSB * C + 1;  // (tiles are never smaller than 320 groups unless ELX_DUO_TILE says so)
  const int64_t tmp_pitch = (int64_t)((slots_bound + 16 + 127) / 128 * 128);  // + 16: the un-permute pass reads whole words
  size_t cs = dt->cap_sites;
  DUO_CUDA(h, duo_grow(&dt->d_comp8, &cs, (size_t)M));
  cs = dt->cap_sites;
  DUO_CUDA(h, duo_grow(&dt->d_root8, &cs, (size_t)M));
  DUO_CUDA(h, duo_grow(&dt->d_pos32, &dt->cap_sites, (size_t)M));
  size_t cl = dt->cap_slots;
  DUO_CUDA(h, duo_grow(&dt->d_rootb, &cl, slots_bound));
  DUO_CUDA(h, duo_grow(&dt->d_site32, &dt->cap_slots, slots_bound));
  size_t ch = dt->cap_hist;
  DUO_CUDA(h, duo_grow(&dt->d_hist, &ch, (size_t)ntile_b * C));
  DUO_CUDA(h, duo_grow(&dt->d_base, &dt->cap_hist, (size_t)ntile_b * C));
  size_t cg = dt->cap_seg;
  DUO_CUDA(h, duo_grow(&dt->d_total, &cg, (size_t)NSB * C));
  DUO_CUDA(h, duo_grow(&dt->d_seg, &dt->cap_seg, (size_t)NSB * C));
  if (!dt->d_r0) DUO_CUDA(h, cudaMalloc((void **)&dt->d_r0, sizeof(uint32_t) * DUO_MAX_C));
  DUO_CUDA(h, duo_grow(&dt->d_tiles, &dt->cap_tiles, (size_t)max_tiles));
  DUO_CUDA(h, duo_grow(&dt->d_tmp, &dt->cap_tmp, (size_t)tmp_pitch * h->T));

  // Pipeline: batches of consecutive superblocks, the un-permute pass of a batch (HBM-bound) on a second stream under the
  // simulation of the next one (integer-bound).  The two kernels do run side by side (128-thread un-permute blocks next to a
  // 544-thread simulation CTA), but each slows the other down (cfg3 time line: the simulation kernels end after 13.6 instead
  // of 10.2 ms, the passes take 6.5 instead of 4.7 ms in total), so the gain is a few per cent, and only on long calls.
  int conc = 148;
  cudaDeviceGetAttribute(&conc, cudaDevAttrMultiProcessorCount, h->device);
  conc *= dt->threads == 256 ? dt->minb : 1;
  // Round 2, end: with the leaner un-permute blocks cfg3 (12 superblocks, 4 batches) ends 4.5 % sooner (15.21 -> 14.52 ms), cfg4
  // (3 superblocks) is unchanged and cfg5 (1.2 superblocks: the one-wave tile sizing is lost) 38 % later: calls of at least 8
  // superblocks are pipelined, ELX_DUO_PIPELINE=0 / 1 forces either way.
  static const int pipe_env = getenv("ELX_DUO_PIPELINE") ? atoi(getenv("ELX_DUO_PIPELINE")) : -1;
  const int want_batches = (pipe_env > 0 || (pipe_env < 0 && NSB >= 8)) ? 0 : 1;
  if (NSB >= 2 && want_batches != 1 && tile_groups > 544) tile_groups = 512;  // (registers for a co-resident un-permute block)
  const double sb_tiles = (double)((int64_t)1 << dt->sb) / 16.0 / tile_groups;  // CTA tiles of one full superblock
  const int L = std::max(1, (int)std::ceil(0.75 * conc / sb_tiles));
  std::vector<int> batch_sb{0};
  if (want_batches != 1) {
    int rem = NSB;
    while ((int)batch_sb.size() < DUO_MAX_BATCH - 1 && rem >= 2 * L && rem >= 2) {
      const int take = std::max(L, (rem + 1) / 2);
      batch_sb.push_back(batch_sb.back() + take);
      rem -= take;
    }
  }
  if (batch_sb.back() != NSB) batch_sb.push_back(NSB);
  const int nbatch = (int)batch_sb.size() - 1;

  DuoBucketArgs b;
  memset(&b, 0, sizeof b);
  b.s = a; b.base0 = base0; b.M = M; b.u0 = u0; b.NSB = NSB; b.sb = dt->sb; b.nbatch = nbatch;
  for (int i = 0; i <= nbatch; ++i) b.batch_sb[i] = batch_sb[(size_t)i];
  b.tile_begin = dt->d_tile_begin;
  b.comp8 = dt->d_comp8; b.root8 = dt->d_root8; b.rootb = dt->d_rootb; b.pos32 = dt->d_pos32; b.site32 = dt->d_site32;
  b.hist = dt->d_hist; b.base = dt->d_base; b.total = dt->d_total; b.r0 = dt->d_r0; b.seg = dt->d_seg; b.tiles = dt->d_tiles;
  b.ntiles = dt->d_ntiles; b.tile_groups = tile_groups; b.max_tiles = max_tiles;
  DUO_CUDA(h, cudaMemsetAsync(dt->d_r0, 0, sizeof(uint32_t) * DUO_MAX_C, h->stream));
  DUO_CUDA(h, cudaMemsetAsync(dt->d_rootb, 0, slots_bound, h->stream));
  DUO_CUDA(h, cudaMemsetAsync(dt->d_site32, 0, slots_bound * sizeof(uint32_t), h->stream));
  if (h->rounds == 7) k_duo_comps<7><<<(unsigned)ntile_b, 256, 0, h->stream>>>(b);
  else k_duo_comps<10><<<(unsigned)ntile_b, 256, 0, h->stream>>>(b);
  k_duo_scan<<<(unsigned)(((int64_t)NSB * C * 32 + 255) / 256), 256, 0, h->stream>>>(b);
  k_duo_layout<<<1, 256, sizeof(int) * (size_t)NSB * C, h->stream>>>(b);
  k_duo_scatter<<<(unsigned)ntile_b, 256, 0, h->stream>>>(b);
  h->launches += 4;
  DUO_CUDA(h, cudaGetLastError());
  if (nbatch > 1) DUO_CUDA(h, cudaEventRecord(dt->ev_bucket, h->stream));

  DuoArgs fa;
  memset(&fa, 0, sizeof fa);
  fa.s = a;
  fa.s.n_slots = dt->prog.n_slots;
  fa.G = dt->G; fa.recs = dt->d_recs; fa.ftab = dt->d_ftab; fa.dres32 = dt->d_dres32;
  // the CTA class of this call: 512 threads whenever the tile fits (measured 2 % faster than the 640 class at the same tile)
  const bool cls512 = dt->threads == 640 && dt->nstage512 >= 2 && tile_groups <= 512;
  const int cls = cls512 ? 512 : dt->threads;
  const size_t cls_smem = cls512 ? dt->smem512 : dt->smem;
  fa.stage_bytes = dt->stage_bytes; fa.nstage = cls512 ? dt->nstage512 : dt->nstage; fa.slot_stride = cls * 16;
  fa.tiles = dt->d_tiles; fa.tile_begin = dt->d_tile_begin; fa.rootb = dt->d_rootb; fa.site32 = dt->d_site32;
  fa.tmp = dt->d_tmp; fa.tmp_pitch = tmp_pitch; fa.base0 = base0;
  fa.mhi = (uint32_t)((65536 / h->K + 1) << 8);
  fa.negk256 = (uint32_t)(-(h->K << 8));
  fa.rk = philox_round_keys(a.seed);
  static const int unperm_mode = getenv("ELX_DUO_UNPERM") ? atoi(getenv("ELX_DUO_UNPERM")) : 3;  // tuning knob: 1 = plain gather, 2 = word copies, 3 = TMA
  // the un-permute blocks' lean gather path (3.2 instead of 8.1 instructions per cell), measured on ONE box (box-to-box variance
  // is as large as the effect): cfg3 pipelined 14.53 against 14.96 ms, cfg4 7.04 against 7.18 ms, cfg5 16.77 against 17.33 ms.
  // ELX_DUO_UNPERM_LEAN=0 selects the general path for every block (A/B runs).
  static const int lean_env = getenv("ELX_DUO_UNPERM_LEAN") ? atoi(getenv("ELX_DUO_UNPERM_LEAN")) : -1;
  static const bool timeline = getenv("ELX_DUO_TIMELINE") != nullptr;  // debug: per-kernel start / end times of a call on stderr
  std::vector<cudaEvent_t> tev;
  auto mark = [&](cudaStream_t st) { if (timeline) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); tev.push_back(e); } };
  mark(h->stream);
  for (int bt = 0; bt < nbatch; ++bt) {
    // sites of this batch in bucketing index space, and an upper bound of its CTA tiles
    const int64_t ua = std::max<int64_t>(u0, (int64_t)batch_sb[(size_t)bt] << dt->sb), ub = std::min<int64_t>(M, (int64_t)batch_sb[(size_t)bt + 1] << dt->sb);
    int grid_tiles = 0;
    for (int j = batch_sb[(size_t)bt]; j < batch_sb[(size_t)bt + 1]; ++j) {
      const int64_t nj = std::min<int64_t>(M, ((int64_t)j + 1) << dt->sb) - std::max<int64_t>(u0, (int64_t)j << dt->sb);
      grid_tiles += (int)((nj / 16 + 2 * C) / tile_groups) + C + 1;
    }
    if (grid_tiles > max_tiles) grid_tiles = max_tiles;
    fa.batch = bt;
    cudaStream_t ms = nbatch > 1 ? dt->sm[bt] : h->stream;
    if (nbatch > 1) DUO_CUDA(h, cudaStreamWaitEvent(ms, dt->ev_bucket, 0));
    mark(ms);
    cudaError_t e;
    if (tiled) {
      e = h->rounds == 7 ? launch_duo_tiled_t<7>(cls, dt->minb, tile_groups, fa, grid_tiles, cls_smem, ms)
                         : launch_duo_tiled_t<10>(cls, dt->minb, tile_groups, fa, grid_tiles, cls_smem, ms);
    } else {
      const dim3 grid((unsigned)grid_tiles, (unsigned)(tile_groups * 16 / 256));
      if (h->rounds == 7) k2_duo_generic<7><<<grid, 256, 0, ms>>>(fa);
      else k2_duo_generic<10><<<grid, 256, 0, ms>>>(fa);
      e = cudaGetLastError();
    }
    if (e != cudaSuccess) return fail(h, ELX_E_CUDA, std::string("duo kernel launch: ") + cudaGetErrorString(e));
    h->launches++;
    mark(ms);
    cudaStream_t us = h->stream;
    if (nbatch > 1) {
      us = dt->s2;
      DUO_CUDA(h, cudaEventRecord(dt->ev_main[bt], ms));
      DUO_CUDA(h, cudaStreamWaitEvent(us, dt->ev_main[bt], 0));
    }
    mark(us);
    // enough blocks to fill the machine a few times over, rows split only when there are few site tiles
    auto rows_per_block = [&](int64_t nt) {
      int rpb = h->T;
      while (rpb > 64 && nt * ((h->T + rpb - 1) / rpb) < 148 * 16) rpb = (rpb + 1) / 2;
      return (rpb + UNP_R - 1) / UNP_R * UNP_R;
    };
    if (unperm_mode == 1) {
      const int rpb = 64;
      const dim3 ugrid((unsigned)((ub - ua + 255) / 256), (unsigned)((h->T + rpb - 1) / rpb));
      k_duo_unpermute<<<ugrid, 256, 0, us>>>(dt->d_tmp, tmp_pitch, dt->d_pos32 + (ua - u0), ub - ua, h->T, rpb, a.leaves + (ua - u0), a.leaves_pitch);
    } else if (unperm_mode == 3 && C <= 16) {
      constexpr int TS = 128 * 16;
      const int64_t nt = (ub + TS - 1) / TS - ua / TS;
      const int rpb = rows_per_block(nt), rowb = (TS + 32 * C + 127) / 128 * 128;
      const size_t dsmem = (size_t)2 * UNP_R * rowb;
      DUO_CUDA(h, cudaFuncSetAttribute(k_duo_unpermute_async<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem));
      // the same shared-memory carve-out as the simulation kernel: an SM does not run two carve-outs side by side
      DUO_CUDA(h, cudaFuncSetAttribute(k_duo_unpermute_async<128, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      const dim3 ugrid((unsigned)nt, (unsigned)((h->T + rpb - 1) / rpb));
      k_duo_unpermute_async<128, true><<<ugrid, 128, dsmem, us>>>(dt->d_tmp, tmp_pitch, dt->d_pos32, dt->d_comp8, u0, ua, ub, C, h->T, rpb, rowb,
                                                                a.leaves, a.leaves_pitch, lean_env >= 0 ? lean_env : 1);
    } else if (unperm_mode == 3 && C <= UNP_ASYNC_MAX_C) {
      constexpr int TS = 256 * 16;
      const int64_t nt = (ub + TS - 1) / TS - ua / TS;
      const int rpb = rows_per_block(nt), rowb = (TS + 32 * C + 127) / 128 * 128;
      const size_t dsmem = (size_t)2 * UNP_R * rowb;
      DUO_CUDA(h, cudaFuncSetAttribute(k_duo_unpermute_async<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem));
      const dim3 ugrid((unsigned)nt, (unsigned)((h->T + rpb - 1) / rpb));
      k_duo_unpermute_async<256, false><<<ugrid, 256, dsmem, us>>>(dt->d_tmp, tmp_pitch, dt->d_pos32, dt->d_comp8, u0, ua, ub, C, h->T, rpb, rowb,
                                                                 a.leaves, a.leaves_pitch, lean_env >= 0 ? lean_env : 1);
    } else {
      const int64_t nt = (ub + UNP_TS - 1) / UNP_TS - ua / UNP_TS;
      const int rpb = rows_per_block(nt);
      const dim3 ugrid((unsigned)nt, (unsigned)((h->T + rpb - 1) / rpb));
      k_duo_unpermute_smem<<<ugrid, 256, 0, us>>>(dt->d_tmp, tmp_pitch, dt->d_pos32, dt->d_comp8, u0, ua, ub, C, h->T, rpb, a.leaves, a.leaves_pitch);
    }
    h->launches++;
    mark(us);
    DUO_CUDA(h, cudaGetLastError());
  }
  if (timeline) {
    cudaDeviceSynchronize();
    fprintf(stderr, "[elx duo] %d batch(es), tile %d groups:", nbatch, tile_groups);
    for (size_t i = 1; i < tev.size(); ++i) {
      float ms = 0;
      cudaEventElapsedTime(&ms, tev[0], tev[i]);
      fprintf(stderr, "%s%.2f", (i - 1) % 4 == 0 ? " | sim " : (i - 1) % 4 == 2 ? " unp " : "-", ms);
    }
    fprintf(stderr, " ms\n");
    for (cudaEvent_t e : tev) cudaEventDestroy(e);
  }
  if (nbatch > 1) {  // the caller's stream sees the whole alignment
    DUO_CUDA(h, cudaEventRecord(dt->ev_done, dt->s2));
    DUO_CUDA(h, cudaStreamWaitEvent(h->stream, dt->ev_done, 0));
  }
  h->last_kernel = tiled ? "k2_duo_tiled<" + std::to_string(h->rounds) + "," + std::to_string(cls) + "," + std::to_string(cls == 256 ? dt->minb : 1) + ">"
                         : "k2_duo_generic<" + std::to_string(h->rounds) + ">";
  *used = 1;
  return 0;
}

}  // namespace elx
