// Quad mode of the free-running simulation (K2) for models with at most 4 states (DNA) on sm_100a.
//
// The reference draws one state per (site, node) from the row of the parent's state (`jump`, Simulate/MarkovProcess.hs:48-49,
// walk MarkovProcessAlongTree.hs:88-98) and simulateAndFlatten keeps the leaves only (:71-83).  The joint distribution of
// the leaves is all that matters, so internal nodes whose state nobody needs can be SUMMED OUT instead of drawn: the
// host cuts the tree into "caps" (elx_internal.hpp, compile_quad_walk) -- pieces below a node r of known state with at
// most 4 frontier nodes (leaves, or internal nodes kept in a slot because further caps hang below them) -- and every cap
// is ONE draw from the 256-outcome joint row  P(frontier states | state of r)  (Felsenstein-style sums over the cap's
// inner nodes, k1_quad_joint).  A random binary tree with T leaves needs about 0.36 T such draws, i.e. 2.8 alignment
// cells per draw instead of 1 (pair mode) or 0.5 (node by node), and every per-draw cost of the hot loop shrinks alike.
//
// Draw specification (oracle/freerun_spec.c spec_freerun_quad executes it on the CPU; kernels must agree bit for bit):
//   record g = (out[0..3] | r), site s, component c, state i of r.  A draw is 24 bits: the 16 sites of a group s >> 4 share
//   three Philox calls, 12 words X[0..11] = philox(s>>4 lo, s>>4 hi, out[0], 4 * call), call = 0, 1, 2;  d = s & 15:
//     x24 = X[3(d>>2) + (d&3)] >> 8 for d&3 < 3, else the low bytes of X[3(d>>2)], X[3(d>>2)+1], X[3(d>>2)+2];  x24 = [t : 16][j : 8]
//     t != 0:  e = qe32[g][c][i][j] = [q : 16][alias : 8][0 : 8];  o = x24 < (e >> 8) ? j : alias   (one 24-bit compare)
//     t == 0 (escape, 2^-16 per draw):  u = philox(s lo, s hi, out[0], TAG_REFINE).x = [tr : 24][jr : 8];
//              er = qres32[g][c][i][jr] = [thr : 24][alias : 8];  o = tr < thr ? jr : alias
//     state(out[k]) = (o >> 2k) & 3
// The joint rows carry integer masses round(p / sum * 2^48), split into a main level of 2^-24 cells and an exact residual
// (two-level alias rows, elx_ptx.cuh): no draw ever ties with a threshold, and the escape test needs the draw alone.  24-bit
// draws instead of 32-bit ones cut the Philox work -- 60 % of the hot loop -- by a quarter.
//
// Kernels: k1_quad_joint, k1_alias_quad (tables), k2_quad_generic (tables through L1/L2; the cross-check),
// k2_quad_tiled (the hot kernel: one record + its 4 KB x C joint rows per stage, streamed through shared memory by a
// producer warp with TMA bulk copies; 16 sites per thread; full/empty mbarriers, no CTA barrier in the loop).
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

constexpr int QUAD_MAX_SLOTS = 24;
constexpr int QUAD_MAX_NSTAGE = 8;
constexpr int QUAD_COLS = 256;

// Device form of a record (32 bytes, streamed through shared memory in front of the record's joint rows).
struct QuadRecDev {
  uint32_t node0;
  int8_t src, dst[4], nfr, pad[2];
  int32_t pad2;
  int32_t leaf[4];  // second 16-byte half
};
static_assert(sizeof(QuadRecDev) == 32, "QuadRecDev must stay 32 bytes");

struct QuadCapDev {
  int32_t n;
  int32_t node[QUAD_MAX_CAP];
  int8_t par[QUAD_MAX_CAP], fr[QUAD_MAX_CAP];
  int32_t pad;
};

struct QuadTables {
  QuadProgram prog;
  int G = 0;
  QuadRecDev *d_recs = nullptr;
  QuadCapDev *d_caps = nullptr;
  double *d_joint = nullptr;    // [G][C][4][256]
  uint32_t *d_qe32 = nullptr;   // [G][C][4][256]
  uint32_t *d_qres32 = nullptr;  // [G][C][4][256]
  uint8_t *d_ftab = nullptr;    // [G] x (32 B record + tb bytes of rows)
  int tiled = 0, threads = 0, minb = 1, nstage = 0, tb = 0, stage_bytes = 0;  // threads: the CTA class (MAXT)
  size_t smem = 0;
};

struct QuadArgs {
  SimArgs s;
  int G;
  const QuadRecDev *recs;
  const uint32_t *qe32;
  const uint32_t *qres32;
  const uint8_t *ftab;
  int tb, stage_bytes, nstage, slot_stride;
  uint32_t entry_bytes;  // 4 (see entry_addr)
  PhiloxRoundKeys rk;  // of the call's seed
};

// ---------------------------------------------------------------------------------------------
// K1 (quad rows): joint[g][c][i][o] = P(frontier states = o | state of the cap's root = i), the cap's inner nodes summed
// out by one pruning pass (children before parents).  One thread per (g, c, i, o); rows i >= K stay zero.
// ---------------------------------------------------------------------------------------------
__global__ void k1_quad_joint(int G, int C, int K, const QuadCapDev *__restrict__ caps, const double *__restrict__ P,
                              double *__restrict__ joint) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)G * C * 4 * QUAD_COLS) return;
  const int o = (int)(idx & 255), i = (int)((idx >> 8) & 3), c = (int)((idx >> 10) % C), g = (int)((idx >> 10) / C);
  if (i >= K) { joint[idx] = 0.0; return; }
  const QuadCapDev cap = caps[g];
  double L[QUAD_MAX_CAP][4];
  int used = 0;
  for (int n = 0; n < cap.n; ++n) {
    const int fr = cap.fr[n];
    if (fr >= 0) used |= 3 << (2 * fr);
    for (int s = 0; s < 4; ++s) L[n][s] = fr < 0 ? 1.0 : (((o >> (2 * fr)) & 3) == s ? 1.0 : 0.0);
  }
  if (o & ~used) { joint[idx] = 0.0; return; }  // absent frontier positions are fixed at state 0
  double top = 1.0;
  for (int n = cap.n - 1; n >= 0; --n) {
    const double *pn = P + ((int64_t)cap.node[n] * C + c) * K * K;
    const int p = cap.par[n];
    if (p >= 0) {
      for (int s = 0; s < K; ++s) {
        double msg = 0.0;
        for (int t = 0; t < K; ++t) {
          double v = pn[s * K + t];
          if (!(v > 0.0)) v = 0.0;
          msg = __dadd_rn(msg, __dmul_rn(v, L[n][t]));
        }
        L[p][s] = __dmul_rn(L[p][s], msg);
      }
    } else {
      double msg = 0.0;
      for (int t = 0; t < K; ++t) {
        double v = pn[i * K + t];
        if (!(v > 0.0)) v = 0.0;
        msg = __dadd_rn(msg, __dmul_rn(v, L[n][t]));
      }
      top = __dmul_rn(top, msg);
    }
  }
  joint[idx] = top;
}

// Two-level alias rows of 256 columns (elx_ptx.cuh alias2_row_warp): integer masses round(p_o / sum * 2^48) (residual to the
// largest); main entry = [q : 16][alias : 8][0 : 8], escape entry qres32 = [thr : 24][alias : 8].  One WARP per row (one
// thread per row with 5.5 KB of local arrays took 0.36 ms for cfg2's 5808 rows).
constexpr int ALIAS_WARPS = 4;
__global__ void __launch_bounds__(ALIAS_WARPS * 32) k1_alias_quad(int64_t rows, const double *__restrict__ joint, uint32_t *__restrict__ e32,
                                                              uint32_t *__restrict__ qres32) {
  constexpr int KP = QUAD_COLS;
  extern __shared__ __align__(16) uint8_t sm_alias[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * ALIAS_WARPS + w;
  if (r >= rows) return;
  uint8_t *base = sm_alias + (size_t)w * (KP * (8 + 8 + 8 + 6));
  double *v = reinterpret_cast<double *>(base);
  int64_t *m = reinterpret_cast<int64_t *>(base + KP * 8), *q = reinterpret_cast<int64_t *>(base + KP * 16);
  int16_t *alias = reinterpret_cast<int16_t *>(base + KP * 24), *small = alias + KP, *large = alias + 2 * KP;
  alias2_row_warp<KP, 16, 48>(joint + r * KP, v, m, q, alias, small, large, lane, e32 + r * KP, qres32 + r * KP);
}

// ---------------------------------------------------------------------------------------------
// K2 quad mode, generic: any C, tables through L1/L2, one thread per site (it recomputes the three Philox calls of its
// 16-site group for every record: the cross-check of the tiled kernel, not a fast path).
// ---------------------------------------------------------------------------------------------
template <int ROUNDS>
__device__ __forceinline__ uint32_t quad_draw24(uint64_t s, uint32_t node0, uint2 key) {
  const uint64_t grp = s >> 4;
  const int d = (int)(s & 15), w0 = 3 * (d >> 2), wa = (d & 3) < 3 ? w0 + (d & 3) : w0, wb = (d & 3) < 3 ? wa : w0 + 2;
  uint32_t X[12];
  for (int call = wa >> 2; call <= wb >> 2; ++call) {  // the one or two Philox calls that hold the draw's words
    const uint4 x = philox4x32<ROUNDS>(make_uint4((uint32_t)grp, (uint32_t)(grp >> 32), node0, 4u * (uint32_t)call), key);
    X[4 * call] = x.x; X[4 * call + 1] = x.y; X[4 * call + 2] = x.z; X[4 * call + 3] = x.w;
  }
  return draw24_from_words(X, d);
}

template <int ROUNDS>
__global__ void k2_quad_generic(QuadArgs fa) {
  const SimArgs &a = fa.s;
  const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= a.nsites) return;
  const uint64_t s = (uint64_t)(a.site0 + col);
  const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
  uint8_t slot[QUAD_MAX_SLOTS];
  const uint32_t cr = draw_comp_root<ROUNDS>(a.cumw, a.cumpi, a.C, a.K, a.N, a.is_mixture, a.error_flag, s, key.x, key.y);
  const uint32_t c = cr & 0xffffu, root = cr >> 16;
  if (a.comps) a.comps[col] = (int32_t)c;
  slot[0] = (uint8_t)root;
  for (int g = 0; g < fa.G; ++g) {
    const uint4 h0 = __ldg(reinterpret_cast<const uint4 *>(fa.recs + g));
    const uint4 h1 = __ldg(reinterpret_cast<const uint4 *>(fa.recs + g) + 1);
    const int src = (int)(int8_t)(h0.y & 0xffu);
    const int dst[4] = {(int)(int8_t)((h0.y >> 8) & 0xffu), (int)(int8_t)((h0.y >> 16) & 0xffu), (int)(int8_t)(h0.y >> 24),
                        (int)(int8_t)(h0.z & 0xffu)};
    const int leaf[4] = {(int)h1.x, (int)h1.y, (int)h1.z, (int)h1.w};
    const uint32_t x24 = quad_draw24<ROUNDS>(s, h0.x, key);
    const uint32_t j = x24 & 0xffu, t = x24 >> 8;
    const int64_t idx = ((((int64_t)g * a.C + c) * 4 + slot[src]) << 8) | j;
    uint32_t o;
    if (t != 0u) {
      const uint32_t e = __ldg(fa.qe32 + idx);
      o = x24 < (e >> 8) ? j : ((e >> 8) & 0xffu);
    } else {  // escape level
      const uint4 z = philox4x32<ROUNDS>(make_uint4((uint32_t)s, (uint32_t)(s >> 32), h0.x, TAG_REFINE), key);
      const uint32_t er = __ldg(fa.qres32 + (idx & ~(int64_t)255) + (z.x & 0xffu));
      o = (z.x >> 8) < (er >> 8) ? (z.x & 0xffu) : (er & 0xffu);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint8_t v = (uint8_t)((o >> (2 * k)) & 3u);
      if (leaf[k] >= 0) a.leaves[(int64_t)leaf[k] * a.leaves_pitch + col] = v;
      else if (dst[k] >= 0) slot[dst[k]] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K2 quad mode, tiled (the hot kernel).  THREADS consumer threads x 16 consecutive sites each (four Philox calls per
// record) + one producer warp.  A slot holds one byte per site: comp * 4 + state = the row of the record's table, so the
// entry index of a draw is  row << 8 | j  -- one PRMT to move the row byte, one LOP3 to merge j, one IMAD for the address.
// The outcome byte (4 x 2-bit states) falls out of a 32-bit min of draw and entry; a leaf row is (W >> 2k) & 0x03030303
// of the packed outcome bytes, written as 512 contiguous bytes per warp with one 128-bit streaming store per lane.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
// prmt.b32 in its default mode: selector nibble n picks byte (n & 7) of {a, b}; bit 3 of the nibble replicates that byte's
// sign bit over the whole byte instead (__byte_perm masks the selector with 0x7777 and loses this)
template <uint32_t SEL>
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "n"(SEL));
  return d;
}
// base + idx * entry_bytes as an integer multiply-add: the FMA pipe has room, the ALU pipe has not.  entry_bytes (= 4)
// arrives as a launch constant because ptxas turns a multiply by a literal 4 into an ALU-pipe LEA.
__device__ __forceinline__ uint32_t entry_addr(uint32_t idx, uint32_t entry_bytes, uint32_t base) {
  uint32_t d;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(idx), "r"(entry_bytes), "r"(base));
  return d;
}
// Stage header of the tiled kernel (32 bytes in front of the record's rows), decoded on the host so that the loop
// spends no instruction on it: out[k] >= 0: leaf row; out[k] = 0x80000000 | byte offset of the slot: internal frontier
// node; out[k] = 0xffffffff: absent.
struct QuadStageHdr {
  uint32_t node0;
  uint32_t src_off;  // byte offset of the slot holding the root's states (slot * slot stride)
  uint32_t nhi, nlo;  // hi / lo of PHILOX_M1 * node0: the record's share of the first Philox round (philox3_split)
  uint32_t out[4];
};
static_assert(sizeof(QuadStageHdr) == 32, "QuadStageHdr must stay 32 bytes");

// The record loop of one thread.  FAST: every 16-site group of this thread is stored with one aligned 128-bit store
// (whole groups inside the site range, 16-byte aligned rows) -- the loop is compiled twice so that the common case
// carries no store-mode tests.
template <int ROUNDS, bool FAST>
__device__ __forceinline__ void quad_records(const QuadArgs &fa, const uint32_t full_bar0, const uint32_t empty_bar0,
                                             const uint32_t stage_smem, const uint32_t slot_base, const uint32_t CB0,
                                             const uint32_t CB1, const uint32_t CB2, const uint32_t CB3, const int64_t sbase,
                                             const int64_t col0, uint8_t *leaf_base, const int store_mode, const int lane) {
  const SimArgs &a = fa.s;
  const uint64_t G0 = (uint64_t)sbase >> 4;  // the thread's 16-site group
  const uint32_t G0lo = (uint32_t)G0, Ghi = (uint32_t)(G0 >> 32);
  constexpr uint32_t SM = 0x03030303u;
  const uint32_t slot_base_m = slot_base - 0x80000000u;  // + (0x80000000 | offset) = slot_base + offset
  const int nstage = fa.nstage;
  const uint32_t pitch32 = (uint32_t)a.leaves_pitch;
  int buf = 0;
  uint32_t phase = 0;
  const PhiloxThreadPart tp = philox_thread_part<ROUNDS>(G0lo, 0u, 4u, 8u, fa.rk);
  for (int g = 0; g < fa.G; ++g) {
    mbar_wait(full_bar0 + 8u * (uint32_t)buf, phase);
    const uint32_t sbuf = stage_smem + (uint32_t)(buf * fa.stage_bytes);
    const uint4 h0 = lds_v4(sbuf);
    const uint4 h1 = lds_v4(sbuf + 16u);
    const uint32_t node0 = h0.x;
    const uint32_t tbl = sbuf + 32u;
#ifdef ELX_DEBUG_BOUNDS
    if (h0.y >= (uint32_t)a.n_slots * (uint32_t)fa.slot_stride || node0 >= (uint32_t)a.N) atomicExch(a.error_flag, -8);
    if (h0.z != __umulhi(PHILOX_M1, node0) || h0.w != PHILOX_M1 * node0) atomicExch(a.error_flag, -9);
#endif
    const uint4 rv = lds_v4(slot_base + h0.y);
    const uint32_t Rq[4] = {rv.x, rv.y, rv.z, rv.w};
    // 48 random bytes = 16 draws of 24 bits
    uint32_t X[12];
    philox3_split<ROUNDS>(tp, h0.z, h0.w, Ghi, fa.rk, X);  // = philox4x32(G0lo, Ghi, node0, 4 * call), call = 0, 1, 2
    uint32_t W[4];
    uint32_t m = 0xffffffffu;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t c[4], t[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int d = 4 * q + i;
        // draw d in the upper three bytes (the low byte is don't-care): [t : 16][j : 8][x : 8]
        const uint32_t w = ELX_DRAW24(X, d);
        // entry index = row << 8 | j in ONE byte permute: byte 0 = j (byte 1 of the draw), byte 1 = row byte of site d,
        // bytes 2,3 = that byte's sign replicated = 0 (rows are < 128)
        const uint32_t idx = i == 0 ? prmt<0xcc41u>(w, Rq[q]) : i == 1 ? prmt<0xdd51u>(w, Rq[q])
                           : i == 2 ? prmt<0xee61u>(w, Rq[q]) : prmt<0xff71u>(w, Rq[q]);
#ifdef ELX_DEBUG_BOUNDS
        if (idx * 4u + 4u > (uint32_t)fa.tb) atomicExch(a.error_flag, -7);
#endif
        const uint32_t e = lds_u32(entry_addr(idx, fa.entry_bytes, tbl));
        c[i] = min(w, e);
        t[i] = w;
      }
      m = min(min(m, t[0]), min(t[1], min(t[2], t[3])));
      W[q] = __byte_perm(__byte_perm(c[0], c[1], 0x0051u), __byte_perm(c[2], c[3], 0x0051u), 0x5410u);
    }
    if (m < 0x10000u) {
      // escape level (2^-16 per draw): some draw has t == 0; its outcome comes from the row's residual masses with 32 fresh
      // bits.  Unrolled with static register names: no calls, no dynamically indexed arrays (elx_fast.cu).
#pragma unroll
      for (int d = 0; d < 16; ++d) {
        const uint32_t w = ELX_DRAW24(X, d);
        if (w < 0x10000u) {
          const uint32_t row = (Rq[d >> 2] >> (8 * (d & 3))) & 0xffu;
          const uint64_t sg = (uint64_t)sbase + (uint64_t)d;
          const uint4 z = philox4x32_rk<ROUNDS>(make_uint4((uint32_t)sg, (uint32_t)(sg >> 32), node0, TAG_REFINE), fa.rk);
          const uint32_t jr = z.x & 0xffu;
          const uint32_t er = __ldg(fa.qres32 + ((((int64_t)g * a.C * 4 + row) << 8) | jr));
          const uint32_t o = (z.x >> 8) < (er >> 8) ? jr : (er & 0xffu);
          W[d >> 2] = (W[d >> 2] & ~(0xffu << (8 * (d & 3)))) | (o << (8 * (d & 3)));
        }
      }
    }
    // this warp is done with the stage buffer (the header is in registers)
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar0 + 8u * (uint32_t)buf);
    if (++buf == nstage) { buf = 0; phase ^= 1u; }

#define ELX_QUAD_OUT(k, code)                                                                                                  \
  if ((int32_t)(code) >= 0) {                                                                                                  \
    const uint4 v = make_uint4((W[0] >> (2 * k)) & SM, (W[1] >> (2 * k)) & SM, (W[2] >> (2 * k)) & SM, (W[3] >> (2 * k)) & SM); \
    if (FAST) stg_cs_v4(leaf_base + (uint64_t)(uint32_t)(code) * (uint64_t)pitch32, v); /* one IMAD.WIDE (FAST: pitch < 2^32) */ \
    else if (store_mode) store_leaf(a, leaf_base, (int)(code), col0, store_mode, v);                                           \
  } else if ((code) != 0xffffffffu) {                                                                                          \
    sts_v4(slot_base_m + (code),                                                                                               \
           make_uint4(((W[0] >> (2 * k)) & SM) | CB0, ((W[1] >> (2 * k)) & SM) | CB1, ((W[2] >> (2 * k)) & SM) | CB2,          \
                      ((W[3] >> (2 * k)) & SM) | CB3));                                                                        \
  }
    ELX_QUAD_OUT(0, h1.x)
    ELX_QUAD_OUT(1, h1.y)
    ELX_QUAD_OUT(2, h1.z)
    ELX_QUAD_OUT(3, h1.w)
#undef ELX_QUAD_OUT
  }
}

// MAXT: the largest CTA this instantiation is launched with (register budget); the actual number of consumer threads is
// blockDim.x - 32, any multiple of 32 (the launch picks MAXT; smaller CTAs are a tuning knob).  Slots are MAXT * 16 bytes
// apart whatever the CTA size, so that the stage headers (slot byte offsets) do not depend on it.
template <int ROUNDS, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT + 32, MINB) k2_quad_tiled(QuadArgs fa) {
  const int THREADS = (int)blockDim.x - 32;
  extern __shared__ __align__(128) uint8_t smem[];
  const SimArgs &a = fa.s;
  const int tid = threadIdx.x;
  const int nstage = fa.nstage;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem);
  uint64_t *empty_bar = full_bar + QUAD_MAX_NSTAGE;
  const uint32_t stage_smem = smem_u32(smem + 128);
  const uint32_t slots_smem = stage_smem + (uint32_t)(nstage * fa.stage_bytes);

  if (tid == 0) {
    for (int i = 0; i < nstage; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), THREADS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  __syncthreads();

  if (__shfl_sync(0xffffffffu, tid, 0) >= THREADS) {  // (through a shuffle: provably warp-uniform for ptxas)
    // producer warp: one lane keeps the ring full
    if (tid == THREADS) {
      int buf = 0;
      uint32_t phase = 1;  // the first pass over the ring finds every buffer free
      for (int g = 0; g < fa.G; ++g) {
        // the ring gives this warp three records of slack: poll the empty barrier with a back-off.  A tight try_wait loop
        // was 12 % of all issued instructions, all on this warp's sub-partition (top scheduling priority: highest warp
        // id), and the two consumer warps it starved there set the pace of the whole CTA through the ring.
        if (g >= nstage) mbar_wait_backoff(smem_u32(&empty_bar[buf]), phase, 256);
        mbar_expect_tx(smem_u32(&full_bar[buf]), (uint32_t)fa.stage_bytes);
        tma_bulk_g2s(stage_smem + (uint32_t)(buf * fa.stage_bytes), fa.ftab + (size_t)g * fa.stage_bytes, (uint32_t)fa.stage_bytes,
                     smem_u32(&full_bar[buf]));
        if (++buf == nstage) { buf = 0; phase ^= 1u; }
      }
    }
    return;
  }

  const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
  // 16 sites per thread, aligned on global site index multiples of 16; the lanes of a warp own consecutive groups
  const int64_t g16 = (a.site0 >> 4) + (int64_t)blockIdx.x * THREADS + tid;
  const int64_t sbase = g16 << 4;
  const int64_t col0 = sbase - a.site0;
  const bool any = sbase < a.site0 + a.nsites;
  const bool full = (col0 >= 0) && (col0 + 16 <= a.nsites);
  uint32_t CB[4] = {0, 0, 0, 0}, R[4] = {0, 0, 0, 0};  // per site one byte: comp * 4 and comp * 4 + state
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint32_t cr = draw_comp_root<ROUNDS>(a.cumw, a.cumpi, a.C, a.K, a.N, a.is_mixture, a.error_flag, (uint64_t)sbase + i,
                                               key.x, key.y);
    const uint32_t c = cr & 0xffffu, root = cr >> 16;
    CB[i >> 2] |= (c * 4u) << (8 * (i & 3));
    R[i >> 2] |= (c * 4u + root) << (8 * (i & 3));
    const int64_t col = col0 + i;
    if (a.comps && col >= 0 && col < a.nsites) a.comps[col] = (int32_t)c;
  }
  const uint32_t slot_base = slots_smem + (uint32_t)tid * 16u;
  sts_v4(slot_base, make_uint4(R[0], R[1], R[2], R[3]));  // slot 0 = the root states
  const int store_mode = !any ? 0
                              : (full && ((a.leaves_pitch & 15) == 0) && a.leaves_pitch < ((int64_t)1 << 32) && ((reinterpret_cast<uintptr_t>(a.leaves) & 15) == 0) &&
                                 ((col0 & 15) == 0)) ? 1 : 2;
  uint8_t *leaf_base = a.leaves + col0;
  asm volatile("" : "+l"(leaf_base));  // keep the sum in registers: recomputed per store it costs two ALU adds
  // the fast loop needs the whole WARP in store mode 1 (the loops synchronise per warp)
  const bool fast = __all_sync(0xffffffffu, store_mode == 1);
  if (fast)
    quad_records<ROUNDS, true>(fa, smem_u32(full_bar), smem_u32(empty_bar), stage_smem, slot_base, CB[0], CB[1], CB[2], CB[3],
                                        sbase, col0, leaf_base, 1, tid & 31);
  else
    quad_records<ROUNDS, false>(fa, smem_u32(full_bar), smem_u32(empty_bar), stage_smem, slot_base, CB[0], CB[1], CB[2], CB[3],
                                         sbase, col0, leaf_base, store_mode, tid & 31);
}

// stage image of record g: [QuadStageHdr][C x 4 x 256 entries]
__global__ void k_build_quad_ftab(int G, int stage_bytes, int threads, const QuadRecDev *__restrict__ recs, const uint32_t *__restrict__ qe32,
                                  uint8_t *__restrict__ ftab) {
  const int chunks = stage_bytes / 16;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)G * chunks) return;
  const int g = (int)(i / chunks), ch = (int)(i % chunks);
  uint4 v;
  if (ch < 2) {
    const QuadRecDev r = recs[g];
    const uint32_t stride = (uint32_t)threads * 16u;
    if (ch == 0) {
      v = make_uint4(r.node0, (uint32_t)r.src * stride, __umulhi(PHILOX_M1, r.node0), PHILOX_M1 * r.node0);
    } else {
      uint32_t o[4];
      for (int k = 0; k < 4; ++k)
        o[k] = r.leaf[k] >= 0 ? (uint32_t)r.leaf[k] : (r.dst[k] >= 0 ? (0x80000000u | ((uint32_t)r.dst[k] * stride)) : 0xffffffffu);
      v = make_uint4(o[0], o[1], o[2], o[3]);
    }
  } else {
    v = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint8_t *>(qe32) + (size_t)g * (stage_bytes - 32))[ch - 2];
  }
  reinterpret_cast<uint4 *>(ftab + (size_t)g * stage_bytes)[ch] = v;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
#define QUAD_CUDA(h, expr)                                                                                      \
  do {                                                                                                          \
    cudaError_t _e = (expr);                                                                                    \
    if (_e != cudaSuccess) return fail(h, ELX_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));      \
  } while (0)

int quad_groups(const elx_handle *h) {
  const QuadTables *qt = static_cast<const QuadTables *>(h->quad);
  return qt ? qt->G : 0;
}

void quad_rec_to_ints(const QuadRec &r, int32_t *o) {
  o[0] = r.root;
  for (int k = 0; k < 4; ++k) { o[1 + k] = r.out[k]; o[5 + k] = r.leaf[k]; o[9 + k] = r.dst[k]; }
  o[13] = r.src; o[14] = r.n_cap; o[15] = 0;
  for (int i = 0; i < QUAD_MAX_CAP; ++i) {
    const bool in = i < r.n_cap;
    o[16 + 3 * i] = in ? r.cap[i].node : -1;
    o[17 + 3 * i] = in ? r.cap[i].par : -1;
    o[18 + 3 * i] = in ? r.cap[i].fr : -1;
  }
}

// Geometry of the tiled kernel: CTA size x CTAs per SM, preferred order first (measured on cfg2: 2 x 256 threads beats
// 1 x 512 by 5 %); the slot stack must leave room for a ring of at least two stages.
static bool quad_geometry(int C, int n_slots, QuadTables *qt) {
  const int tb = C * 4 * QUAD_COLS * 4;
  const size_t stage = 32 + (size_t)tb;
  int want_threads = 0, want_stages = 0, want_minb = 0;
  if (const char *e = getenv("ELX_QUAD_THREADS")) want_threads = atoi(e);  // tuning knobs
  if (const char *e = getenv("ELX_QUAD_NSTAGE")) want_stages = atoi(e);
  if (const char *e = getenv("ELX_QUAD_MINB")) want_minb = atoi(e);        // 1 or 2 CTAs per SM
  const int cand[3][2] = {{256, 2}, {512, 1}, {256, 1}};
  for (int ci = 0; ci < 3; ++ci) {
    const int th = cand[ci][0], minb = cand[ci][1];
    if (want_threads && th != want_threads) continue;
    if (want_minb && minb != want_minb) continue;
    const size_t per_cta = (size_t)226 * 1024 / minb;
    const size_t slots = (size_t)(n_slots > 0 ? n_slots : 1) * th * 16;
    if (128 + 2 * stage + slots > per_cta) continue;
    int nstage = (int)((per_cta - 128 - slots) / stage);
    const int cap = want_stages ? want_stages : 4;
    if (nstage > cap) nstage = cap;
    if (nstage > QUAD_MAX_NSTAGE) nstage = QUAD_MAX_NSTAGE;
    if (nstage < 2) continue;
    qt->tiled = 1; qt->threads = th; qt->minb = minb; qt->nstage = nstage; qt->tb = tb; qt->stage_bytes = (int)stage;
    qt->smem = 128 + (size_t)nstage * stage + slots;
    return true;
  }
  return false;
}

int quad_create(elx_handle *h, const int32_t *parent, const int32_t *leaf_row) {
  QuadTables *qt = new QuadTables();
  compile_quad_walk(h->N, parent, leaf_row, qt->prog);
  const int G = qt->G = (int)qt->prog.recs.size();
  // quad mode needs the tiled geometry; otherwise the handle stays in pair mode
  if (G == 0 || qt->prog.n_slots > QUAD_MAX_SLOTS || (int64_t)h->C * 4 > 127) { delete qt; return 0; }
  // without a shared-memory geometry (very many components) the tiled pair kernel beats the generic quad kernel
  if (!quad_geometry(h->C, qt->prog.n_slots, qt)) { delete qt; return 0; }
  h->quad = qt;
  std::vector<QuadRecDev> dev((size_t)G);
  std::vector<QuadCapDev> caps((size_t)G);
  for (int g = 0; g < G; ++g) {
    const QuadRec &r = qt->prog.recs[g];
    QuadRecDev &d = dev[g];
    memset(&d, 0, sizeof d);
    d.node0 = (uint32_t)r.out[0];
    d.src = (int8_t)r.src;
    d.nfr = 0;
    for (int k = 0; k < 4; ++k) { d.dst[k] = r.dst[k]; d.leaf[k] = r.leaf[k]; if (r.out[k] >= 0) d.nfr++; }
    QuadCapDev &c = caps[g];
    memset(&c, 0, sizeof c);
    c.n = r.n_cap < QUAD_MAX_CAP ? r.n_cap : QUAD_MAX_CAP;
    for (int i = 0; i < c.n && i < QUAD_MAX_CAP; ++i) { c.node[i] = r.cap[i].node; c.par[i] = r.cap[i].par; c.fr[i] = r.cap[i].fr; }
  }
  const size_t ents = (size_t)G * h->C * 4 * QUAD_COLS;
  QUAD_CUDA(h, cudaMalloc((void **)&qt->d_recs, sizeof(QuadRecDev) * (size_t)G));
  QUAD_CUDA(h, cudaMalloc((void **)&qt->d_caps, sizeof(QuadCapDev) * (size_t)G));
  QUAD_CUDA(h, cudaMalloc((void **)&qt->d_joint, sizeof(double) * ents));
  QUAD_CUDA(h, cudaMalloc((void **)&qt->d_qe32, sizeof(uint32_t) * ents));
  QUAD_CUDA(h, cudaMalloc((void **)&qt->d_qres32, sizeof(uint32_t) * ents));
  QUAD_CUDA(h, cudaMemcpy(qt->d_recs, dev.data(), sizeof(QuadRecDev) * (size_t)G, cudaMemcpyHostToDevice));
  QUAD_CUDA(h, cudaMemcpy(qt->d_caps, caps.data(), sizeof(QuadCapDev) * (size_t)G, cudaMemcpyHostToDevice));
  if (qt->tiled) QUAD_CUDA(h, cudaMalloc((void **)&qt->d_ftab, (size_t)G * qt->stage_bytes));
  return 0;
}

int quad_tables_build(elx_handle *h) {
  QuadTables *qt = static_cast<QuadTables *>(h->quad);
  if (!qt) return 0;
  const int64_t ents = (int64_t)qt->G * h->C * 4 * QUAD_COLS;
  k1_quad_joint<<<(unsigned)((ents + 255) / 256), 256, 0, h->stream>>>(qt->G, h->C, h->K, qt->d_caps, h->d_P, qt->d_joint);
  h->launches++;
  QUAD_CUDA(h, cudaGetLastError());
  const int64_t rows = ents / QUAD_COLS;
  k1_alias_quad<<<(unsigned)((rows + ALIAS_WARPS - 1) / ALIAS_WARPS), ALIAS_WARPS * 32, (size_t)ALIAS_WARPS * QUAD_COLS * 30, h->stream>>>(rows, qt->d_joint, qt->d_qe32, qt->d_qres32);
  h->launches++;
  QUAD_CUDA(h, cudaGetLastError());
  if (qt->tiled) {
    const int64_t n = (int64_t)qt->G * (qt->stage_bytes / 16);
    k_build_quad_ftab<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(qt->G, qt->stage_bytes, qt->threads, qt->d_recs, qt->d_qe32, qt->d_ftab);
    h->launches++;
    QUAD_CUDA(h, cudaGetLastError());
  }
  return 0;
}

void quad_free(elx_handle *h) {
  QuadTables *qt = static_cast<QuadTables *>(h->quad);
  if (!qt) return;
  cudaFree(qt->d_recs); cudaFree(qt->d_caps); cudaFree(qt->d_joint); cudaFree(qt->d_qe32); cudaFree(qt->d_qres32); cudaFree(qt->d_ftab);
  delete qt;
  h->quad = nullptr;
}

int quad_get(elx_handle *h, int32_t *recs, uint32_t *qe32, uint32_t *qres32) {
  QuadTables *qt = static_cast<QuadTables *>(h->quad);
  if (!qt) return fail(h, ELX_E_INVALID, "elx_quad_get: this handle has no quad tables");
  if (recs)
    for (int g = 0; g < qt->G; ++g) quad_rec_to_ints(qt->prog.recs[g], recs + (size_t)g * ELX_QUAD_REC_INTS);
  const size_t n = (size_t)qt->G * h->C * 4 * QUAD_COLS;
  if (qe32) QUAD_CUDA(h, cudaMemcpyAsync(qe32, qt->d_qe32, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  if (qres32) QUAD_CUDA(h, cudaMemcpyAsync(qres32, qt->d_qres32, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  QUAD_CUDA(h, cudaStreamSynchronize(h->stream));
  return 0;
}

template <int ROUNDS, int MAXT, int MINB>
static cudaError_t launch_quad_tiled(int threads, const QuadArgs &fa, int64_t groups, size_t smem, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(k2_quad_tiled<ROUNDS, MAXT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const unsigned blocks = (unsigned)((groups + threads - 1) / threads);
  k2_quad_tiled<ROUNDS, MAXT, MINB><<<blocks, threads + 32, smem, stream>>>(fa);
  return cudaGetLastError();
}

template <int ROUNDS>
static cudaError_t launch_quad_tiled_t(int maxt, int minb, int threads, const QuadArgs &fa, int64_t groups, size_t smem,
                                       cudaStream_t stream) {
  if (maxt == 512) return launch_quad_tiled<ROUNDS, 512, 1>(threads, fa, groups, smem, stream);
  if (minb == 2) return launch_quad_tiled<ROUNDS, 256, 2>(threads, fa, groups, smem, stream);
  return launch_quad_tiled<ROUNDS, 256, 1>(threads, fa, groups, smem, stream);
}

int quad_simulate(elx_handle *h, SimArgs &a, int *used) {
  *used = 0;
  QuadTables *qt = static_cast<QuadTables *>(h->quad);
  if (!qt || a.internal) return 0;  // summed-out nodes have no states: `simulate` (keep internal states) draws pairs
  QuadArgs fa;
  fa.s = a;
  fa.s.n_slots = qt->prog.n_slots;
  fa.G = qt->G; fa.recs = qt->d_recs; fa.qe32 = qt->d_qe32; fa.qres32 = qt->d_qres32; fa.ftab = qt->d_ftab;
  fa.tb = qt->tb; fa.stage_bytes = qt->stage_bytes; fa.nstage = qt->nstage; fa.slot_stride = qt->threads * 16;
  fa.rk = philox_round_keys(a.seed);
  fa.entry_bytes = 4;
  cudaError_t e;
  if (qt->tiled && !(h->flags & ELX_FLAG_FORCE_GENERIC)) {
    const int64_t g0 = a.site0 >> 4, g1 = (a.site0 + a.nsites + 15) >> 4;
    // Measured on cfg2 (B200): SM throughput saturates at 16 consumer warps (2 x 256 or 1 x 512 threads) and drops with
    // smaller CTAs faster than a better-filled last wave gains, so every launch uses the largest CTA of its class.
    int threads = qt->threads;
    if (const char *ev = getenv("ELX_QUAD_CTA")) {  // tuning knob: a smaller CTA
      const int t = atoi(ev);
      if (t >= 32 && t <= qt->threads && t % 32 == 0) threads = t;
    }
    e = h->rounds == 7 ? launch_quad_tiled_t<7>(qt->threads, qt->minb, threads, fa, g1 - g0, qt->smem, h->stream)
                       : launch_quad_tiled_t<10>(qt->threads, qt->minb, threads, fa, g1 - g0, qt->smem, h->stream);
  } else {
    const unsigned blocks = (unsigned)((a.nsites + 127) / 128);
    if (h->rounds == 7) k2_quad_generic<7><<<blocks, 128, 0, h->stream>>>(fa);
    else k2_quad_generic<10><<<blocks, 128, 0, h->stream>>>(fa);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) return fail(h, ELX_E_CUDA, std::string("quad kernel launch: ") + cudaGetErrorString(e));
  h->launches++;
  h->last_kernel = (qt->tiled && !(h->flags & ELX_FLAG_FORCE_GENERIC))
                       ? "k2_quad_tiled<" + std::to_string(h->rounds) + "," + std::to_string(qt->threads) + "," + std::to_string(qt->threads == 512 ? 1 : qt->minb) + ">"
                       : "k2_quad_generic<" + std::to_string(h->rounds) + ">";
  *used = 1;
  return 0;
}

}  // namespace elx
