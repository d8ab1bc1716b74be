// Pair mode of the free-running simulation (K2) for models with at most 4 states (DNA) on sm_100a.
//
// The reference draws one state per (site, node) from the row of the parent's state (`jump`, Simulate/MarkovProcess.hs:48-49,
// walk MarkovProcessAlongTree.hs:88-98).  Siblings are conditionally independent given their parent, so the two children
// of a node can be drawn TOGETHER from the joint row  P_A[i][sa] * P_B[i][sb]  (16 outcomes, o = sa | sb << 2) with one
// alias lookup: one 16-bit draw, one table entry, one packed min yields two node states.  A binary tree with T leaves has
// T-1 such groups (+ the root), i.e. ONE draw per alignment cell instead of two, which halves every per-draw cost
// (Philox, address arithmetic, LDS, min) of the hot kernel.
//
// Draw specification (oracle/freerun_spec.c spec_freerun_pair executes it on the CPU; kernels must agree bit for bit):
//   record g = (nodeA, nodeB | parent p), site s, component c, parent state i:
//     x16 = field (s & 7) of philox(s>>3 lo, s>>3 hi, nodeA, TAG_NODE);  x16 = [r : 12][j : 4]
//     e = pe16[g][c][i][j] = [q_hi : 12][alias : 4];  r < q_hi -> o = j;  r > q_hi -> o = alias;
//     tie -> philox(s lo, s hi, nodeA, TAG_REFINE).x < pqlo[g][c][i][j] ? j : alias
//     state(nodeA) = o & 3, state(nodeB) = o >> 2
// The joint rows carry integer masses round(pA*pB / sum * 2^48) (k1_alias on the 16-outcome row, column capacity 2^44).
//
// Kernels: k1_joint (outer products of the P rows), k2_pair_generic (tables through L1/L2, keeps internal states),
// k2_pair_tiled (the hot kernel: records + joint rows streamed through shared memory by TMA bulk copies).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/elynx_b200.h"
#include "elx_device.cuh"
#include "elx_ptx.cuh"
#include "elx_state.hpp"

namespace elx {

constexpr int PAIR_MAX_SLOTS = 32;
constexpr int PAIR_MAX_NSTAGE = 3;

struct PairTables {
  PairProgram prog;
  std::vector<int32_t> nodeA, nodeB;
  int G = 0;
  PairRecDev *d_recs = nullptr;
  int32_t *d_nodeA = nullptr, *d_nodeB = nullptr;
  double *d_joint = nullptr;  // [G][C][K][16] outer products (input of k1_alias)
  uint16_t *d_pe16 = nullptr;
  uint32_t *d_pqlo = nullptr;
  // tiled kernel
  int tiled = 0;
  uint8_t *d_ftab = nullptr;
  int tb = 0, nps = 0, nstage = 0, stage_bytes = 0, n_stages = 0;
  size_t smem = 0;
};

struct PairArgs {
  SimArgs s;
  const uint8_t *ftab;
  int tb, stage_bytes, n_stages, nps, nstage;
};

// ---------------------------------------------------------------------------------------------
// K1 (pair rows): joint[g][c][i][sa | sb << 2] = P[nodeA][c][i][sa] * P[nodeB][c][i][sb]; a group of one puts all of
// B's mass on sb = 0.  One thread per (g, c, i) row.
// ---------------------------------------------------------------------------------------------
__global__ void k1_joint(int G, int C, int K, const int32_t *__restrict__ nodeA, const int32_t *__restrict__ nodeB,
                         const double *__restrict__ P, double *__restrict__ joint) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= (int64_t)G * C * K) return;
  const int i = (int)(r % K), c = (int)((r / K) % C), g = (int)(r / ((int64_t)K * C));
  const int nA = nodeA[g], nB = nodeB[g];
  const double *pa = P + (((int64_t)nA * C + c) * K + i) * K;
  const double *pb = nB >= 0 ? P + (((int64_t)nB * C + c) * K + i) * K : nullptr;
  double *out = joint + r * 16;
  for (int o = 0; o < 16; ++o) {
    const int sa = o & 3, sb = o >> 2;
    double va = sa < K ? pa[sa] : 0.0;
    double vb = pb ? (sb < K ? pb[sb] : 0.0) : (sb == 0 ? 1.0 : 0.0);
    if (!(va > 0.0)) va = 0.0;
    if (!(vb > 0.0)) vb = 0.0;
    out[o] = __dmul_rn(va, vb);
  }
}

// ---------------------------------------------------------------------------------------------
// K2 pair mode, generic: any C, tables through L1/L2, one thread per 8 consecutive sites (one Philox call per record).
// Serves `internal` output (simulate / simulateMixtureModel keep every node's states) and the cross-check.
// ---------------------------------------------------------------------------------------------
template <int ROUNDS>
__global__ void k2_pair_generic(SimArgs a) {
  const int64_t g = (a.site0 >> 3) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t sbase = g << 3;
  if (sbase >= a.site0 + a.nsites) return;
  const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
  uint64_t slot[PAIR_MAX_SLOTS];
  int comp[8];
  uint64_t regs = 0;
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    const uint64_t s = (uint64_t)sbase + d;
    int c = 0;
    if (a.is_mixture) {
      uint4 x = philox4x32<ROUNDS>(make_uint4((uint32_t)s, (uint32_t)(s >> 32), 0u, TAG_COMP), key);
      c = cat_cum(a.cumw, a.C, u53_positive(x.x, x.y));
      if (c < 0) { atomicExch(a.error_flag, a.N + 1); c = 0; }
    }
    comp[d] = c;
    uint4 x = philox4x32<ROUNDS>(make_uint4((uint32_t)s, (uint32_t)(s >> 32), 0u, TAG_ROOT), key);
    int root = cat_cum(a.cumpi + (int64_t)c * a.K, a.K, u53_positive(x.x, x.y));
    if (root < 0) { atomicExch(a.error_flag, a.N + 2); root = 0; }
    regs |= (uint64_t)root << (8 * d);
    const int64_t col = (int64_t)s - a.site0;
    if (a.comps && col >= 0 && col < a.nsites) a.comps[col] = c;
  }
  auto put_row = [&](uint8_t *base, int64_t pitch, int row, uint64_t v) {
    uint8_t *dst = base + (int64_t)row * pitch + (sbase - a.site0);
    for (int d = 0; d < 8; ++d) {
      const int64_t col = sbase + d - a.site0;
      if (col >= 0 && col < a.nsites) dst[d] = (uint8_t)(v >> (8 * d));
    }
  };
  for (int k = 0; k < a.G; ++k) {
    const PairRecDev rec = a.pairs[k];
    const uint64_t par = rec.src >= 0 ? slot[rec.src] : regs;
    const uint4 x = philox4x32<ROUNDS>(make_uint4((uint32_t)g, (uint32_t)((uint64_t)g >> 32), (uint32_t)rec.nodeA, TAG_NODE), key);
    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
    uint64_t outA = 0, outB = 0;
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      const uint32_t x16 = (w[d >> 1] >> ((d & 1) * 16)) & 0xffffu;
      const uint32_t j = x16 & 15u, r = x16 >> 4;
      const int64_t idx = ((((int64_t)k * a.C + comp[d]) * a.K) + (int64_t)((par >> (8 * d)) & 0xff)) * 16 + j;
      const uint32_t e = __ldg(a.pe16 + idx), qh = e >> 4, al = e & 15u;
      uint32_t o;
      if (r < qh) o = j;
      else if (r > qh) o = al;
      else {
        const uint64_t s = (uint64_t)sbase + d;
        const uint4 z = philox4x32<ROUNDS>(make_uint4((uint32_t)s, (uint32_t)(s >> 32), (uint32_t)rec.nodeA, TAG_REFINE), key);
        o = z.x < __ldg(a.pqlo + idx) ? j : al;
      }
      outA |= (uint64_t)(o & 3u) << (8 * d);
      outB |= (uint64_t)(o >> 2) << (8 * d);
    }
    regs = outA;
    if (rec.dstA >= 0) slot[rec.dstA] = outA;
    if (rec.dstB >= 0) slot[rec.dstB] = outB;
    if (rec.leafA >= 0) put_row(a.leaves, a.leaves_pitch, rec.leafA, outA);
    if (rec.leafB >= 0) put_row(a.leaves, a.leaves_pitch, rec.leafB, outB);
    if (a.internal) {
      put_row(a.internal, a.internal_pitch, rec.nodeA, outA);
      const int nB = a.pair_nodeB[k];
      if (nB >= 0) put_row(a.internal, a.internal_pitch, nB, outB);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K2 pair mode, tiled (the hot kernel).  One thread = 16 consecutive sites (two Philox calls per record); a warp writes
// 512 contiguous bytes of a leaf row with one 128-bit streaming store per lane.  Per-site state of the parent lives in
// 16-bit lanes holding the BYTE OFFSET of the joint row: (comp*4 + state) * 32; the entry address is that offset OR'ed
// with 2*j (both lanes of a register in one LOP3).  Live sibling states wait in a shared-memory slot stack as state
// bytes [slot][thread] x 16 B (conflict-free 128-bit accesses); member A of a record stays in registers.
// Records + joint rows arrive in stages of `nps` records by TMA bulk copies (full mbarriers, one CTA barrier retires a
// stage buffer and thread 0 refills it).
// ---------------------------------------------------------------------------------------------
template <int ROUNDS>
__global__ void __launch_bounds__(128, 4) k2_pair_tiled(PairArgs fa) {
  constexpr int THREADS = 128;
  extern __shared__ __align__(128) uint8_t smem[];
  const SimArgs &a = fa.s;
  const int tid = threadIdx.x;
  const int nstage = fa.nstage, nps = fa.nps;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem);
  const uint32_t stage_smem = smem_u32(smem + 128);
  const uint32_t slots_smem = stage_smem + (uint32_t)(nstage * fa.stage_bytes);
  constexpr uint32_t slot_stride = THREADS * 16u;

  if (tid == 0) {
    for (int i = 0; i < nstage; ++i) mbar_init(smem_u32(&full_bar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  __syncthreads();
  if (tid == 0) {
    for (int i = 0; i < nstage && i < fa.n_stages; ++i) {
      mbar_expect_tx(smem_u32(&full_bar[i]), (uint32_t)fa.stage_bytes);
      tma_bulk_g2s(stage_smem + (uint32_t)(i * fa.stage_bytes), fa.ftab + (size_t)i * fa.stage_bytes, (uint32_t)fa.stage_bytes,
                   smem_u32(&full_bar[i]));
    }
  }

  const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
  // 16 sites per thread, aligned on global site index multiples of 16; the lanes of a warp own consecutive groups
  const int64_t g16 = (a.site0 >> 4) + (int64_t)blockIdx.x * THREADS + tid;
  const int64_t sbase = g16 << 4;
  const int64_t col0 = sbase - a.site0;
  const bool any = sbase < a.site0 + a.nsites;
  const bool full = (col0 >= 0) && (col0 + 16 <= a.nsites);
  uint32_t cbase[8], par[8];  // 16-bit lanes: comp*128 and (comp*4 + state)*32 = byte offset of the joint row
#pragma unroll
  for (int q = 0; q < 8; ++q) { cbase[q] = 0; par[q] = 0; }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint32_t cr = draw_comp_root<ROUNDS>(a.cumw, a.cumpi, a.C, a.K, a.N, a.is_mixture, a.error_flag, (uint64_t)sbase + i,
                                               key.x, key.y);
    const uint32_t c = cr & 0xffffu, root = cr >> 16;
    cbase[i >> 1] |= (c * 128u) << (16 * (i & 1));
    par[i >> 1] |= ((c * 4u + root) * 32u) << (16 * (i & 1));
    const int64_t col = col0 + i;
    if (a.comps && col >= 0 && col < a.nsites) a.comps[col] = (int32_t)c;
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) asm volatile("" : "+r"(cbase[q]));
  uint32_t slot_base = slots_smem + (uint32_t)tid * 16u;
  asm volatile("" : "+r"(slot_base));
  int store_mode = !any ? 0
                        : (full && ((a.leaves_pitch & 15) == 0) && ((reinterpret_cast<uintptr_t>(a.leaves) & 15) == 0) &&
                           ((col0 & 15) == 0)) ? 1 : 2;
  asm volatile("" : "+r"(store_mode));
  uint8_t *leaf_base = a.leaves + col0;
  asm volatile("" : "+l"(leaf_base));
  const uint64_t G0 = (uint64_t)sbase >> 3;
  const uint32_t G0lo = (uint32_t)G0, Ghi = (uint32_t)(G0 >> 32), G1lo = G0lo | 1u;

  int buf = 0;
  uint32_t phase = 0;
  for (int st = 0; st < fa.n_stages; ++st) {
    mbar_wait(smem_u32(&full_bar[buf]), phase);
    const uint32_t sbuf = stage_smem + (uint32_t)(buf * fa.stage_bytes);
    const int nn = min(nps, a.G - st * nps);
    for (int k = 0; k < nn; ++k) {
      const uint4 rec = lds_v4(sbuf + (uint32_t)k * 16u);
      const uint32_t nodeA = rec.x;
      const int src = (int)(int16_t)(rec.y & 0xffffu);
      const int dstA = (int)(int8_t)((rec.y >> 16) & 0xffu);
      const int dstB = (int)(int8_t)(rec.y >> 24);
      const int leafA = (int)rec.z, leafB = (int)rec.w;
      const uint32_t tbl = sbuf + (uint32_t)(nps * 16) + (uint32_t)k * (uint32_t)fa.tb;
#ifdef ELX_DEBUG_BOUNDS
      if (src >= a.n_slots || dstA >= a.n_slots || dstB >= a.n_slots || nodeA >= (uint32_t)a.N ||
          k * fa.tb + fa.tb + nps * 16 > fa.stage_bytes)
        atomicExch(a.error_flag, -8);
#endif
      if (src >= 0) {  // the parent's states were pushed as state bytes: expand to row-offset lanes
        const uint4 v = lds_v4(slot_base + (uint32_t)src * slot_stride);
        const uint32_t b[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          par[2 * q] = __byte_perm(b[q], 0u, 0x4140u) * 32u + cbase[2 * q];
          par[2 * q + 1] = __byte_perm(b[q], 0u, 0x4342u) * 32u + cbase[2 * q + 1];
        }
      }
      const uint4 x0 = philox4x32<ROUNDS>(make_uint4(G0lo, Ghi, nodeA, TAG_NODE), key);
      const uint4 x1 = philox4x32<ROUNDS>(make_uint4(G1lo, Ghi, nodeA, TAG_NODE), key);
      const uint32_t w[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      uint32_t c2[8], t[8];
#pragma unroll
      for (int pr = 0; pr < 8; ++pr) {  // sites 2pr, 2pr+1 share Philox word pr
        // both entry addresses in one LOP3: row offset | 2*j per 16-bit lane (an OR, not an ADD: see elx_fast.cu)
        const uint32_t ad = ((w[pr] << 1) & 0x001e001eu) | par[pr];
        const uint32_t aA = ad & 0xffffu, aB = ad >> 16;
#ifdef ELX_DEBUG_BOUNDS
        if (aA + 2 > (uint32_t)fa.tb || aB + 2 > (uint32_t)fa.tb || (aA & 1) || (aB & 1)) atomicExch(a.error_flag, -7);
#endif
        const uint32_t eA = lds_u16(tbl + aA);
        const uint32_t eB = lds_u16(tbl + aB);
        const uint32_t ee = __byte_perm(eA, eB, 0x5410u);
        c2[pr] = __vminu2(w[pr], ee);
        t[pr] = w[pr] ^ ee;
      }
      // tie detection: some 16-bit lane of t has its 12 threshold bits all zero
      const uint32_t m01 = __vimin3_u16x2(t[0], t[1], t[2]);
      const uint32_t m23 = __vimin3_u16x2(t[3], t[4], t[5]);
      const uint32_t m = __vimin3_u16x2(m01, m23, __vminu2(t[6], t[7]));
      if (((m & 0xffffu) < 16u) || ((m >> 16) < 16u)) {
        // exact path (2^-12 per draw): resolve the tied draws with 32 more random bits.  No calls, no arrays (elx_fast.cu).
        uint32_t mask = 0;
#pragma unroll
        for (int pr = 0; pr < 8; ++pr)
          if (((t[pr] & 0xffffu) < 16u) || ((t[pr] >> 16) < 16u)) mask |= 1u << pr;
        const int64_t grec = (int64_t)st * nps + k;
#pragma unroll 1
        while (mask) {
          const int pr = __ffs(mask) - 1;
          mask &= mask - 1;
          uint32_t tt = 0, ww = 0, cc = 0, oo = 0;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (q == pr) { tt = t[q]; ww = w[q]; cc = c2[q]; oo = par[q]; }
#pragma unroll 1
          for (int hh = 0; hh < 2; ++hh) {
            const uint32_t sh = (uint32_t)hh * 16u;
            if (((tt >> sh) & 0xffffu) < 16u) {
              const uint32_t x16 = (ww >> sh) & 0xffffu, off = (oo >> sh) & 0xffffu, j = x16 & 15u;
              const uint32_t e = lds_u16(tbl + off + 2u * j);
              const uint64_t sg = (uint64_t)sbase + 2 * pr + hh;
              const uint4 z = philox4x32<ROUNDS>(make_uint4((uint32_t)sg, (uint32_t)(sg >> 32), nodeA, TAG_REFINE), key);
              const int64_t idx = (grec * (a.C * 4) + (off >> 5)) * 16 + j;
              const uint32_t o = z.x < __ldg(a.pqlo + idx) ? j : (e & 15u);
              cc = (cc & ~(0xffffu << sh)) | (o << sh);
            }
          }
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (q == pr) c2[q] = cc;
        }
      }
      // outcome o = sa | sb << 2 in the low 4 bits of every 16-bit lane
      uint32_t by[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) by[q] = __byte_perm(c2[2 * q], c2[2 * q + 1], 0x6420u);
      constexpr uint32_t SM = 0x03030303u;
      if (leafA >= 0) {
        if (store_mode) store_leaf(a, leaf_base, leafA, col0, store_mode, make_uint4(by[0] & SM, by[1] & SM, by[2] & SM, by[3] & SM));
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) par[q] = ((c2[q] << 5) & 0x00600060u) | cbase[q];
        if (dstA >= 0) sts_v4(slot_base + (uint32_t)dstA * slot_stride, make_uint4(by[0] & SM, by[1] & SM, by[2] & SM, by[3] & SM));
      }
      if (leafB >= 0) {
        if (store_mode)
          store_leaf(a, leaf_base, leafB, col0, store_mode,
                     make_uint4((by[0] >> 2) & SM, (by[1] >> 2) & SM, (by[2] >> 2) & SM, (by[3] >> 2) & SM));
      } else if (dstB >= 0) {
        sts_v4(slot_base + (uint32_t)dstB * slot_stride,
               make_uint4((by[0] >> 2) & SM, (by[1] >> 2) & SM, (by[2] >> 2) & SM, (by[3] >> 2) & SM));
      }
    }
    __syncthreads();  // every warp is done with this stage buffer
    if (tid == 0 && st + nstage < fa.n_stages) {
      fence_async_smem();
      mbar_expect_tx(smem_u32(&full_bar[buf]), (uint32_t)fa.stage_bytes);
      tma_bulk_g2s(sbuf, fa.ftab + (size_t)(st + nstage) * fa.stage_bytes, (uint32_t)fa.stage_bytes, smem_u32(&full_bar[buf]));
    }
    if (++buf == nstage) { buf = 0; phase ^= 1u; }
  }
}

// stage = [nps x PairRecDev (16 B)] [nps x tb bytes of joint rows]; pe16 [G][C][4][16] u16 is already the row layout
__global__ void k_build_pair_ftab(int G, int nps, int tb, int stage_bytes, const PairRecDev *__restrict__ recs,
                                  const uint16_t *__restrict__ pe16, uint8_t *__restrict__ ftab) {
  const int chunks = (16 + tb) / 16;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)G * chunks) return;
  const int k = (int)(i / chunks), ch = (int)(i % chunks);
  const int st = k / nps, kk = k % nps;
  uint8_t *stage = ftab + (size_t)st * stage_bytes;
  if (ch == 0) {
    *reinterpret_cast<uint4 *>(stage + kk * 16) = *reinterpret_cast<const uint4 *>(&recs[k]);
  } else {
    const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint8_t *>(pe16) + (size_t)k * tb) + (ch - 1);
    *reinterpret_cast<uint4 *>(stage + nps * 16 + (size_t)kk * tb + (ch - 1) * 16) = *src;
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
#define PAIR_CUDA(h, expr)                                                                                      \
  do {                                                                                                          \
    cudaError_t _e = (expr);                                                                                    \
    if (_e != cudaSuccess) return fail(h, ELX_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));      \
  } while (0)

int pair_groups(const elx_handle *h) {
  const PairTables *pt = static_cast<const PairTables *>(h->pair);
  return pt ? pt->G : 0;
}

int pair_create(elx_handle *h, const int32_t *parent, const int32_t *leaf_row, std::string &msg) {
  PairTables *pt = new PairTables();
  compile_pair_walk(h->N, parent, leaf_row, pt->prog);
  if (pt->prog.n_slots > PAIR_MAX_SLOTS) {
    delete pt;
    msg = "tree: the pair walk needs more than 32 live state slots";
    return ELX_E_UNSUPPORTED;
  }
  h->pair = pt;
  const int G = pt->G = (int)pt->prog.recs.size();
  std::vector<PairRecDev> dev((size_t)G);
  pt->nodeA.resize(G);
  pt->nodeB.resize(G);
  for (int g = 0; g < G; ++g) {
    const PairRec &r = pt->prog.recs[g];
    dev[g].nodeA = r.nodeA; dev[g].src = r.src; dev[g].dstA = r.dstA; dev[g].dstB = r.dstB;
    dev[g].leafA = r.leafA; dev[g].leafB = r.leafB;
    pt->nodeA[g] = r.nodeA; pt->nodeB[g] = r.nodeB;
  }
  const size_t rows = (size_t)G * h->C * h->K;
  PAIR_CUDA(h, cudaMalloc((void **)&pt->d_recs, sizeof(PairRecDev) * (size_t)G));
  PAIR_CUDA(h, cudaMalloc((void **)&pt->d_nodeA, sizeof(int32_t) * (size_t)G));
  PAIR_CUDA(h, cudaMalloc((void **)&pt->d_nodeB, sizeof(int32_t) * (size_t)G));
  PAIR_CUDA(h, cudaMalloc((void **)&pt->d_joint, sizeof(double) * rows * 16));
  PAIR_CUDA(h, cudaMalloc((void **)&pt->d_pe16, sizeof(uint16_t) * rows * 16));
  PAIR_CUDA(h, cudaMalloc((void **)&pt->d_pqlo, sizeof(uint32_t) * rows * 16));
  PAIR_CUDA(h, cudaMemcpy(pt->d_recs, dev.data(), sizeof(PairRecDev) * (size_t)G, cudaMemcpyHostToDevice));
  PAIR_CUDA(h, cudaMemcpy(pt->d_nodeA, pt->nodeA.data(), sizeof(int32_t) * (size_t)G, cudaMemcpyHostToDevice));
  PAIR_CUDA(h, cudaMemcpy(pt->d_nodeB, pt->nodeB.data(), sizeof(int32_t) * (size_t)G, cudaMemcpyHostToDevice));

  // tiled kernel geometry: K == 4 (row = comp*4 + state), row offsets in 16-bit lanes, and at least two stages of one
  // record next to the slot stack
  pt->tiled = 0;
  if (h->K == 4 && (int64_t)h->C * 128 <= 65535) {
    const int tb = h->C * 128;
    const size_t per_rec = 16 + (size_t)tb;
    const size_t slots = (size_t)(pt->prog.n_slots > 0 ? pt->prog.n_slots : 1) * 128 * 16;
    int nstage = tb <= 2048 ? PAIR_MAX_NSTAGE : 2;
    size_t budget = 48 * 1024;
    if (const char *e = getenv("ELX_PAIR_SMEM_KB")) budget = (size_t)atoi(e) * 1024;  // tuning knob
    int nps = budget > slots + 128 ? (int)((budget - slots - 128) / nstage / per_rec) : 0;
    if (nps > 64) nps = 64;
    if (nps < 1) { nps = 1; nstage = 2; }
    const size_t smem = 128 + (size_t)nstage * nps * per_rec + slots;
    if (smem <= 200 * 1024) {
      pt->tiled = 1;
      pt->tb = tb; pt->nps = nps; pt->nstage = nstage; pt->smem = smem;
      pt->stage_bytes = nps * (int)per_rec;
      pt->n_stages = (G + nps - 1) / nps;
      const size_t bytes = (size_t)pt->n_stages * pt->stage_bytes;
      PAIR_CUDA(h, cudaMalloc((void **)&pt->d_ftab, bytes));
      PAIR_CUDA(h, cudaMemset(pt->d_ftab, 0, bytes));
    }
  }
  return 0;
}

int pair_tables_build(elx_handle *h) {
  PairTables *pt = static_cast<PairTables *>(h->pair);
  if (!pt) return 0;
  const int64_t rows = (int64_t)pt->G * h->C * h->K;
  k1_joint<<<(unsigned)((rows + 127) / 128), 128, 0, h->stream>>>(pt->G, h->C, h->K, pt->d_nodeA, pt->d_nodeB, h->d_P, pt->d_joint);
  h->launches++;
  PAIR_CUDA(h, cudaGetLastError());
  launch_k1_alias(h, rows, 16, 4, pt->d_joint, pt->d_pe16, pt->d_pqlo);
  if (pt->tiled) {
    const int chunks = (16 + pt->tb) / 16;
    const int64_t n = (int64_t)pt->G * chunks;
    k_build_pair_ftab<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(pt->G, pt->nps, pt->tb, pt->stage_bytes, pt->d_recs, pt->d_pe16,
                                                                           pt->d_ftab);
    h->launches++;
  }
  PAIR_CUDA(h, cudaGetLastError());
  return 0;
}

void pair_free(elx_handle *h) {
  PairTables *pt = static_cast<PairTables *>(h->pair);
  if (!pt) return;
  cudaFree(pt->d_recs); cudaFree(pt->d_nodeA); cudaFree(pt->d_nodeB); cudaFree(pt->d_joint); cudaFree(pt->d_pe16);
  cudaFree(pt->d_pqlo); cudaFree(pt->d_ftab);
  delete pt;
  h->pair = nullptr;
}

int pair_get(elx_handle *h, int32_t *nodeA, int32_t *nodeB, uint16_t *pe16, uint32_t *pqlo) {
  PairTables *pt = static_cast<PairTables *>(h->pair);
  if (!pt) return fail(h, ELX_E_INVALID, "elx_pair_get: this handle draws node by node (no pair tables)");
  for (int g = 0; g < pt->G; ++g) {
    if (nodeA) nodeA[g] = pt->nodeA[g];
    if (nodeB) nodeB[g] = pt->nodeB[g];
  }
  const size_t n = (size_t)pt->G * h->C * h->K * 16;
  if (pe16) PAIR_CUDA(h, cudaMemcpyAsync(pe16, pt->d_pe16, n * sizeof(uint16_t), cudaMemcpyDeviceToHost, h->stream));
  if (pqlo) PAIR_CUDA(h, cudaMemcpyAsync(pqlo, pt->d_pqlo, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  PAIR_CUDA(h, cudaStreamSynchronize(h->stream));
  return 0;
}

template <int ROUNDS>
static cudaError_t launch_pair_tiled(const PairArgs &fa, int64_t groups, size_t smem, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(k2_pair_tiled<ROUNDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const unsigned blocks = (unsigned)((groups + 127) / 128);
  k2_pair_tiled<ROUNDS><<<blocks, 128, smem, stream>>>(fa);
  return cudaGetLastError();
}

int pair_simulate(elx_handle *h, SimArgs &a, int *used) {
  *used = 0;
  PairTables *pt = static_cast<PairTables *>(h->pair);
  if (!pt) return 0;
  a.G = pt->G; a.pairs = pt->d_recs; a.pair_nodeB = pt->d_nodeB; a.pe16 = pt->d_pe16; a.pqlo = pt->d_pqlo;
  a.n_slots = pt->prog.n_slots;
  cudaError_t e;
  if (pt->tiled && !a.internal && !(h->flags & ELX_FLAG_FORCE_GENERIC)) {
    PairArgs fa;
    fa.s = a;
    fa.ftab = pt->d_ftab; fa.tb = pt->tb; fa.stage_bytes = pt->stage_bytes; fa.n_stages = pt->n_stages; fa.nps = pt->nps;
    fa.nstage = pt->nstage;
    const int64_t g0 = a.site0 >> 4, g1 = (a.site0 + a.nsites + 15) >> 4;
    e = h->rounds == 7 ? launch_pair_tiled<7>(fa, g1 - g0, pt->smem, h->stream) : launch_pair_tiled<10>(fa, g1 - g0, pt->smem, h->stream);
  } else {
    const int64_t g0 = a.site0 >> 3, g1 = (a.site0 + a.nsites + 7) >> 3;
    const unsigned blocks = (unsigned)((g1 - g0 + 127) / 128);
    if (h->rounds == 7) k2_pair_generic<7><<<blocks, 128, 0, h->stream>>>(a);
    else k2_pair_generic<10><<<blocks, 128, 0, h->stream>>>(a);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) return fail(h, ELX_E_CUDA, std::string("pair kernel launch: ") + cudaGetErrorString(e));
  h->launches++;
  h->last_kernel = std::string((pt->tiled && !a.internal && !(h->flags & ELX_FLAG_FORCE_GENERIC)) ? "k2_pair_tiled<" : "k2_pair_generic<") +
                   std::to_string(h->rounds) + ">";
  *used = 1;
  return 0;
}

}  // namespace elx
