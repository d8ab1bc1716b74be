// Building blocks shared by the tiled kernels (elx_fast.cu, elx_pair.cu): mbarrier + TMA bulk copies, shared-memory
// vector accesses, the streaming 128-bit row store and the per-site component / root-state draws of the prologue.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "elx_device.cuh"
#include "elx_state.hpp"

namespace elx {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity, unsigned ns) {
  for (;;) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    if (ok) return;
    __nanosleep(ns);
  }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void stg_cs_v4(void *p, uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}


// prmt.b32 in its default mode: selector nibble n picks byte (n & 7) of {a, b}; bit 3 of the nibble replicates that byte's
// sign bit over the whole byte instead (__byte_perm masks the selector with 0x7777 and loses this)
template <uint32_t SEL>
__device__ __forceinline__ uint32_t prmt_sel(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "n"(SEL));
  return d;
}
// Draw d (0..15) of the 12 random words X[0..11] of a 16-site group (quad and duo mode), in the UPPER three bytes of the result
// (the low byte is don't-care): draws 4k, 4k+1, 4k+2 are the words X[3k], X[3k+1], X[3k+2] as they are, draw 4k+3 collects the
// low bytes those three leave over -- two byte permutes per four draws (contiguous 24-bit fields cost three).
#define ELX_DRAW24(X, d)                                                                                                     \
  ((((d) & 3) < 3) ? X[3 * ((d) >> 2) + (((d) & 3) < 3 ? ((d) & 3) : 0)]                                                     \
                   : elx::prmt_sel<0x4210u>(elx::prmt_sel<0x0400u>(X[3 * ((d) >> 2)], X[3 * ((d) >> 2) + 1]), X[3 * ((d) >> 2) + 2]))
// the same for one site of a group from scratch (generic kernels): word index and Philox call of draw d
__host__ __device__ __forceinline__ uint32_t draw24_from_words(const uint32_t *X, int d) {
  const int w0 = 3 * (d >> 2), i = d & 3;
  if (i < 3) return X[w0 + i] >> 8;
  return (X[w0] & 0xffu) | ((X[w0 + 1] & 0xffu) << 8) | ((X[w0 + 2] & 0xffu) << 16);
}

// ---- two-level alias rows (quad / duo mode) ------------------------------------------------------------------------
// A row of KP outcomes with integer masses P_o = round(p_o / sum * 2^W) (residual to the largest; sum = 2^W) is sampled EXACTLY
// by a draw x = [t : TB][j : log2 KP] without any tie between draw and threshold:
//   t != 0 (main level): the 2^TB - 1 values of t times the KP columns are cells of probability 2^-24 each (TB + log2 KP = 24);
//     masses M_o = floor(P_o / 2^(W-24)), reduced at the largest outcome so that sum M = (2^TB - 1) KP; Vose pairing with column
//     capacity 2^TB - 1 gives column j an acceptance count A_j and an alias.  The kernels compare  [t][j] < [q][alias]  as ONE
//     integer (outcome = that field of min(draw, entry)), which accepts t < q, and t == q when j < alias: the stored q is
//     A_j + 1 - [j < alias]  (0 when A_j = 0; a full column is its own alias with q = 2^TB - 1);
//   t == 0 (escape level, probability 2^-TB): the residual masses R_o = P_o - M_o 2^(W-24) >= 0 (sum = KP 2^(W-24) = 2^32) in a
//     second alias row of column capacity 2^(W-24): entry [thr : W-24][alias : log2 KP], sampled with 32 fresh bits
//     u = [tr : W-24][jr : log2 KP]:  tr < thr -> jr, else alias.
// Total: M_o 2^-24 + R_o 2^-W = P_o 2^-W.  (Round 1 compared t with a 16-bit threshold prefix and settled the ties t == q_hi
// with 24 more bits: the tie detector cost one XOR per draw in the hot loops; the escape test needs the draw alone.)
// vose_exact: single-thread Vose pairing over integer masses m[0..KP) with sum = CAP * KP; q[j] = acceptance units of column j
// (q[j] == CAP and alias[j] == j: a full column).  m is consumed.
template <int KP>
__device__ __forceinline__ void vose_exact(int64_t *m, const int64_t CAP, int64_t *q, int16_t *alias, int16_t *small, int16_t *large) {
  int ns = 0, nl = 0;
  for (int j = KP - 1; j >= 0; --j) {
    alias[j] = (int16_t)j;
    q[j] = CAP;
    if (m[j] < CAP) small[ns++] = (int16_t)j; else large[nl++] = (int16_t)j;
  }
  while (ns > 0 && nl > 0) {
    const int s_ = small[--ns], l_ = large[--nl];
    q[s_] = m[s_];
    alias[s_] = (int16_t)l_;
    m[l_] -= CAP - m[s_];
    if (m[l_] < CAP) small[ns++] = (int16_t)l_; else large[nl++] = (int16_t)l_;
  }
}
// main-level entry of column j: [q : TB][alias : AB][0 : 32 - TB - AB]
template <int TB, int AB>
__device__ __forceinline__ uint32_t alias2_main_entry(int j, int64_t A, int al) {
  constexpr int64_t CAPM = ((int64_t)1 << TB) - 1;
  int64_t qq;
  if (al == j || A >= CAPM) { qq = CAPM; al = j; }  // a full column: every t gives j
  else if (A <= 0) qq = 0;
  else qq = A + 1 - (j < al ? 1 : 0);
  return (uint32_t)((qq << (32 - TB)) | ((int64_t)al << (32 - TB - AB)));
}
// escape-level entry of column j: [thr : 32 - AB][alias : AB]
template <int AB>
__device__ __forceinline__ uint32_t alias2_escape_entry(int j, int64_t thr, int al) {
  constexpr int64_t CAPR = (int64_t)1 << (32 - AB);
  if (al == j || thr >= CAPR) { thr = CAPR - 1; al = j; }
  return (uint32_t)((thr << AB) | (int64_t)al);
}
// One two-level row by ONE WARP (quad rows: KP = 256, TB = 16, W = 48).  The order-sensitive parts -- the sum, the pairings --
// run on lane 0 over shared-memory arrays exactly as a single thread would; loads, masses, the arg-max and the stores are spread
// over the lanes.  sm: 3 KP int64 (v doubles as the residual masses) + 3 KP int16 per warp.
template <int KP, int TB, int W>
__device__ __forceinline__ void alias2_row_warp(const double *__restrict__ p, double *v, int64_t *m, int64_t *q, int16_t *alias,
                                                int16_t *small, int16_t *large, int lane, uint32_t *__restrict__ e32,
                                                uint32_t *__restrict__ res32) {
  constexpr int AB = 24 - TB;   // log2 KP
  constexpr int SH = W - 24;    // log2 of the main level's mass unit
  static_assert((1 << AB) == KP, "TB + log2 KP must be 24");
  for (int j = lane; j < KP; j += 32) { const double x = p[j]; v[j] = x > 0.0 ? x : 0.0; }
  __syncwarp();
  double sum = 0.0;
  if (lane == 0)
    for (int j = 0; j < KP; ++j) sum += v[j];
  sum = __shfl_sync(0xffffffffu, sum, 0);
  int64_t tot = 0, best = -1;
  int big = 0;
  for (int j = lane; j < KP; j += 32) {
    const int64_t mj = sum > 0.0 ? (int64_t)llrint(v[j] / sum * (double)((int64_t)1 << W)) : (j == 0 ? ((int64_t)1 << W) : 0);
    m[j] = mj;
    tot += mj;
    if (mj > best) { best = mj; big = j; }  // (ascending j per lane: the first maximum)
  }
  for (int o = 16; o > 0; o >>= 1) {
    tot += __shfl_xor_sync(0xffffffffu, tot, o);
    const int64_t ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oj = __shfl_xor_sync(0xffffffffu, big, o);
    if (ob > best || (ob == best && oj < big)) { best = ob; big = oj; }
  }
  __syncwarp();
  if (lane == 0) m[big] += ((int64_t)1 << W) - tot;
  __syncwarp();
  // split into the two levels
  int64_t *R = reinterpret_cast<int64_t *>(v);
  int64_t msum = 0;
  for (int j = lane; j < KP; j += 32) {
    const int64_t P = m[j], M = P >> SH;
    R[j] = P - (M << SH);
    m[j] = M;
    msum += M;
  }
  for (int o = 16; o > 0; o >>= 1) msum += __shfl_xor_sync(0xffffffffu, msum, o);
  __syncwarp();
  if (lane == 0) {
    const int64_t excess = msum - (((int64_t)1 << TB) - 1) * KP;  // 1..KP, and the largest outcome holds at least 2^24 / KP cells
    m[big] -= excess;
    R[big] += excess << SH;
    vose_exact<KP>(m, ((int64_t)1 << TB) - 1, q, alias, small, large);
  }
  __syncwarp();
  for (int j = lane; j < KP; j += 32) e32[j] = alias2_main_entry<TB, AB>(j, q[j], alias[j]);
  __syncwarp();
  if (lane == 0) vose_exact<KP>(R, (int64_t)1 << SH, q, alias, small, large);
  __syncwarp();
  for (int j = lane; j < KP; j += 32) res32[j] = alias2_escape_entry<AB>(j, q[j], alias[j]);
  __syncwarp();
}

// mwc-random `categorical`: the first i with cv[i] >= cv[K-1] * u.  cv is a running sum of non-negative weights, hence
// non-decreasing: for long vectors (C60: 60 components) a binary search finds the same index in 6 steps instead of 30.
__device__ __forceinline__ int cat_cum(const double *__restrict__ cv, int K, double u) {
  double p = __dmul_rn(cv[K - 1], u);
  if (K > 16) {
    if (!(cv[K - 1] >= p)) return -1;
    int lo = 0, hi = K - 1;  // invariant: cv[hi] >= p; everything below lo is < p
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cv[mid] >= p) hi = mid; else lo = mid + 1;
    }
    return lo;
  }
  for (int i = 0; i < K; ++i)
    if (cv[i] >= p) return i;
  return -1;
}

// Byte stores of a partial 16-site group (first / last group of a site range, unaligned buffers).  Rolled on purpose:
// unrolled, ptxas hoists the 16 64-bit column-bound predicates above the mode branch and every node pays for them.
__device__ __forceinline__ void store_leaf_bytes(uint8_t *dstp, int64_t col0, int64_t nsites, uint4 o) {
#pragma unroll 1
  for (int i = 0; i < 16; ++i) {
    const int64_t col = col0 + i;
    const uint32_t wv = i < 8 ? (i < 4 ? o.x : o.y) : (i < 12 ? o.z : o.w);
    if (col >= 0 && col < nsites) dstp[i] = (uint8_t)(wv >> ((i & 3) * 8));
  }
}

__device__ __forceinline__ void store_leaf(const SimArgs &a, uint8_t *leaf_base, int leaf_row, int64_t col0, int store_mode,
                                           uint4 o) {
  uint8_t *dstp = leaf_base + (int64_t)leaf_row * a.leaves_pitch;
  if (store_mode == 1) {
    stg_cs_v4(dstp, o);  // 32 lanes x 16 B = 512 contiguous bytes of the [taxon][site] row, streaming (evict-first)
  } else {
    store_leaf_bytes(dstp, col0, a.nsites, o);
  }
}

// Component and root-state draws of one site (once per site; exact fp64 inverse-CDF on 53-bit uniforms, identical
// to the generic kernel).  Out of line: it runs 16 times per thread and round in the prologue and must not inflate
// the register budget of the node loop.  Returns comp | root << 16.
template <int ROUNDS>
__device__ __noinline__ uint32_t draw_comp_root(const double *__restrict__ cumw, const double *__restrict__ cumpi, int C, int K,
                                                int N, int is_mixture, int *error_flag, uint64_t s, uint32_t k0, uint32_t k1) {
  // scalars by value: taking the address of the kernel parameter block would move it to local memory and turn every
  // later read of a launch constant into a per-thread (non-uniform) load
  const uint2 key = make_uint2(k0, k1);
  int c = 0;
  if (is_mixture) {
    uint4 x = philox4x32<ROUNDS>(make_uint4((uint32_t)s, (uint32_t)(s >> 32), 0u, TAG_COMP), key);
    c = cat_cum(cumw, C, u53_positive(x.x, x.y));
    if (c < 0) { atomicExch(error_flag, N + 1); c = 0; }
  }
  uint4 x = philox4x32<ROUNDS>(make_uint4((uint32_t)s, (uint32_t)(s >> 32), 0u, TAG_ROOT), key);
  int root = cat_cum(cumpi + c * K, K, u53_positive(x.x, x.y));
  if (root < 0) { atomicExch(error_flag, N + 2); root = 0; }
  return (uint32_t)c | ((uint32_t)root << 16);
}

}  // namespace elx
