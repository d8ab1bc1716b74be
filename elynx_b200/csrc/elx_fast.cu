// Tiled shared-memory simulation kernels (K2, free-running mode) for sm_100a.
//
// Same draw specification as k2_freerun_generic (elx_core.cu) -- the two must agree bit for bit -- but laid out for
// the machine:
//   * one thread owns 16 consecutive sites (two Philox4x32 calls per node = 16 x 16-bit draws) and writes each leaf
//     row with one 128-bit streaming store (a warp covers 512 contiguous bytes of the [taxon][site] row);
//   * the per-branch alias tables and the walk records are streamed through shared memory in stages of `nps` nodes
//     by TMA bulk copies (cp.async.bulk + mbarrier, up to 3 stages in flight; one CTA barrier per stage retires a buffer);
//   * live internal-node states sit in a shared-memory slot stack [slot][thread] (128-bit, conflict free); the
//     walk order (heavy child last) bounds the number of slots by floor(log2 N)+1;
//   * a draw is: PRMT (parent offset) + LOP3 (address) + LDS.U16 + half a PRMT/VIMNMX.U16x2/LOP3 -- the alias
//     decision is a packed 16-bit min: child bits = low bits of min(random, entry); ties in the high bits (2^-14)
//     fall to an exact 32-bit refinement path.
//
// DNA kernel: K = 4, KP = 4, C <= 8.  Per-site state byte = comp*32 + state*8 = byte offset of the table row.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>

#include <cuda_runtime.h>

#include "../../include/elynx_b200.h"
#include "elx_device.cuh"
#include "elx_state.hpp"

namespace elx {

constexpr int MAX_NSTAGE = 3;       // TMA stages in flight
constexpr int FAST_THREADS = 128;    // threads per CTA (4 warps); thread 0 also issues the TMA refills
constexpr int FAST_CTA = FAST_THREADS;

struct FastTables {
  int kind = 0;  // 0 = none (generic kernel), 1 = DNA layout, 2 = protein layout
  uint8_t *d_ftab = nullptr;
  int tb = 0;           // table bytes per node
  int nps = 0;          // nodes per TMA stage
  int nstage = 0;       // stages in flight (2 or 3)
  int stage_bytes = 0;  // nps * (16 + tb)
  int n_stages = 0;
  size_t smem = 0;
};

struct FastArgs {
  SimArgs s;
  const uint8_t *ftab;
  int tb, stage_bytes, n_stages, nps, nstage;
};

// ---- PTX helpers --------------------------------------------------------------------------------
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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok)
               : "r"(bar), "r"(parity)
               : "memory");
  return ok != 0;
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

__device__ __forceinline__ int cat_cum(const double *__restrict__ cv, int K, double u) {
  double p = __dmul_rn(cv[K - 1], u);
  for (int i = 0; i < K; ++i)
    if (cv[i] >= p) return i;
  return -1;
}

// ---------------------------------------------------------------------------------------------------
// Tiled kernel.  AA = false: DNA layout (K = 4, KP = 4, C <= 8; per-site byte = comp*32 + state*8, 4 packed regs).
//                AA = true : protein layout (K <= 32 > 4, KP = 32, rows of 64 B; per-site 16-bit lane =
//                            (comp*K + state)*64, 8 packed regs; C*K*64 <= 65535).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_leaf(const SimArgs &a, uint8_t *leaf_base, int leaf_row, int64_t col0, int store_mode,
                                           uint4 o) {
  uint8_t *dstp = leaf_base + (int64_t)leaf_row * a.leaves_pitch;
  if (store_mode == 1) {
    stg_cs_v4(dstp, o);  // 32 lanes x 16 B = 512 contiguous bytes of the [taxon][site] row, streaming (evict-first)
  } else {
    const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
    for (int i = 0; i < 16; ++i) {
      int64_t col = col0 + i;
      if (col >= 0 && col < a.nsites) dstp[i] = (uint8_t)(ow[i >> 2] >> ((i & 3) * 8));
    }
  }
}

// w >> 15 computed as the high word of w * 2^17: a multiply runs on the FMA pipe, a shift on the (binding) ALU pipe.
__device__ __forceinline__ uint32_t hi15(uint32_t w) {
  uint32_t r;
  asm("mul.hi.u32 %0, %1, 131072;" : "=r"(r) : "r"(w));
  return r;
}

template <bool AA>
struct Lay {
  static constexpr int NREG = AA ? 8 : 4;            // packed parent registers per thread (16 sites)
  static constexpr uint32_t JMASK = AA ? 31u : 3u;   // column bits of a 16-bit draw x = [r : 16-A][j : A]
  static constexpr uint32_t TIE = AA ? 32u : 4u;      // lane of (x ^ e) below this <=> threshold bits equal
  static constexpr int A = AA ? 5 : 2;
};

// Component and root-state draws of one site (once per site; exact fp64 inverse-CDF on 53-bit uniforms, identical
// to the generic kernel).  Out of line: it runs 16 times per thread in the prologue and must not inflate the
// register budget of the node loop.  Returns comp | root << 16.
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

#ifndef ELX_MIN_CTAS
#define ELX_MIN_CTAS 4
#endif
template <int ROUNDS, bool AA>
__global__ void __launch_bounds__(FAST_CTA, ELX_MIN_CTAS) k2_tiled(FastArgs fa) {
  using L = Lay<AA>;
  constexpr int NREG = L::NREG;
  extern __shared__ __align__(128) uint8_t smem[];
  const SimArgs &a = fa.s;
  const int tid = threadIdx.x;
  const int nstage = fa.nstage, nps = fa.nps;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem);  // [nstage] mbarriers: the stage's TMA bytes have landed
  const uint32_t stage_smem = smem_u32(smem + 128);
  const uint32_t slots_smem = stage_smem + (uint32_t)(nstage * fa.stage_bytes);
  const uint32_t my_slot0 = slots_smem + (uint32_t)tid * 16u;  // slot stack [slot][thread] x 16 B
  const uint32_t slot_stride = FAST_THREADS * 16u;

  if (tid == 0) {
    for (int i = 0; i < nstage; ++i) mbar_init(smem_u32(&full_bar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  __syncthreads();

  // Fill the ring: stages 0..nstage-1.  Refills happen at the end of each stage, after the CTA barrier that
  // retires the buffer, in straight-line code.  (Measured: per-warp release with an in-band producer loop was no
  // faster, and any data-dependent loop under `tid == 0` makes ptxas drop uniform-register addressing for the whole
  // node loop, which costs one extra ALU-pipe IADD3 per draw.)
  if (tid == 0) {
    for (int i = 0; i < nstage && i < fa.n_stages; ++i) {
      mbar_expect_tx(smem_u32(&full_bar[i]), (uint32_t)fa.stage_bytes);
      tma_bulk_g2s(stage_smem + (uint32_t)(i * fa.stage_bytes), fa.ftab + (size_t)i * fa.stage_bytes, (uint32_t)fa.stage_bytes,
                   smem_u32(&full_bar[i]));
    }
  }

  // 16 sites per thread, aligned on global site index multiples of 16
  const int64_t g16 = (a.site0 >> 4) + (int64_t)blockIdx.x * FAST_THREADS + tid;
  const int64_t sbase = g16 << 4;
  const int64_t col0 = sbase - a.site0;
  const bool any = sbase < a.site0 + a.nsites;
  const bool full = (col0 >= 0) && (col0 + 16 <= a.nsites);
  const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
  const uint64_t G0 = (uint64_t)sbase >> 3;
  const uint32_t G0lo = (uint32_t)G0, Ghi = (uint32_t)(G0 >> 32), G1lo = G0lo | 1u;

  // component and root-state draws (once per site; exact fp64 inverse-CDF, as in the generic kernel)
  uint32_t cbase[NREG], par[NREG];
#pragma unroll
  for (int q = 0; q < NREG; ++q) { cbase[q] = 0; par[q] = 0; }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint32_t cr = draw_comp_root<ROUNDS>(a.cumw, a.cumpi, a.C, a.K, a.N, a.is_mixture, a.error_flag, (uint64_t)sbase + i, key.x, key.y);
    const int c = (int)(cr & 0xffffu), root = (int)(cr >> 16);
    if (AA) {
      cbase[i >> 1] |= (uint32_t)(c * a.K * 64) << (16 * (i & 1));
      par[i >> 1] |= (uint32_t)((c * a.K + root) * 64) << (16 * (i & 1));
    } else {
      cbase[i >> 2] |= (uint32_t)(c * 32) << (8 * (i & 3));
      par[i >> 2] |= (uint32_t)(c * 32 + root * 8) << (8 * (i & 3));
    }
    int64_t col = col0 + i;
    if (a.comps && col >= 0 && col < a.nsites) a.comps[col] = c;
  }
  // Pin the loop invariants into single registers: without this ptxas keeps the per-site pieces of cbase[] apart and
  // re-ORs them (and re-derives the slot address and the store predicates) on every node.
#pragma unroll
  for (int q = 0; q < NREG; ++q) asm volatile("" : "+r"(cbase[q]));
  uint32_t slot_base = my_slot0;
  asm volatile("" : "+r"(slot_base));
  // slot 0 = root states.  DNA slots hold the packed offset bytes (= par[]); protein slots hold state bytes
  // (the leaf output format) and are expanded to 16-bit offset lanes on reload.
  if (AA) {
    uint32_t b[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t lo = (par[2 * q] - cbase[2 * q]) >> 6, hi = (par[2 * q + 1] - cbase[2 * q + 1]) >> 6;  // state lanes
      b[q] = __byte_perm(lo, hi, 0x6420u);
    }
    sts_v4(my_slot0, make_uint4(b[0], b[1], b[2], b[3]));
  } else {
    sts_v4(my_slot0, make_uint4(par[0], par[1], par[2], par[3]));
  }
  int reg_slot = 0;  // slot whose contents are mirrored in par[]

  // store mode: 0 = nothing to store (thread past the site range), 1 = 128-bit row stores, 2 = byte stores (ragged / unaligned)
  int store_mode = !any ? 0
                        : (full && ((a.leaves_pitch & 15) == 0) && ((reinterpret_cast<uintptr_t>(a.leaves) & 15) == 0) && ((col0 & 15) == 0)) ? 1 : 2;
  asm volatile("" : "+r"(store_mode));
  uint8_t *leaf_base = a.leaves + col0;
  asm volatile("" : "+l"(leaf_base));

  // ring position kept incrementally (a runtime % / would run on the vector pipes and make every shared-memory
  // address below non-uniform)
  int buf = 0;
  uint32_t phase = 0;
  for (int st = 0; st < fa.n_stages; ++st) {
    mbar_wait(smem_u32(&full_bar[buf]), phase);
    const uint32_t sbuf = stage_smem + (uint32_t)(buf * fa.stage_bytes);
    const int nn = min(nps, a.N - st * nps);
    for (int k = 0; k < nn; ++k) {
      const uint4 rec = lds_v4(sbuf + (uint32_t)k * 16u);
      const uint32_t node = rec.x;
      const int src = (int)(int16_t)(rec.y & 0xffffu);
      const int dst = (int)(int16_t)(rec.y >> 16);
      const int leaf_row = (int)rec.z;
      if (src != reg_slot) {
        const uint4 v = lds_v4(slot_base + (uint32_t)src * slot_stride);
        if (AA) {
          const uint32_t b[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {  // state bytes -> 16-bit lanes -> row offsets (state*64 + comp*K*64)
            par[2 * q] = __byte_perm(b[q], 0u, 0x4140u) * 64u + cbase[2 * q];
            par[2 * q + 1] = __byte_perm(b[q], 0u, 0x4342u) * 64u + cbase[2 * q + 1];
          }
        } else {
          par[0] = v.x; par[1] = v.y; par[2] = v.z; par[3] = v.w;
        }
        reg_slot = src;
      }
      const uint4 x0 = philox4x32<ROUNDS>(make_uint4(G0lo, Ghi, node, TAG_NODE), key);
      const uint4 x1 = philox4x32<ROUNDS>(make_uint4(G1lo, Ghi, node, TAG_NODE), key);
      const uint32_t w[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      const uint32_t tbl = sbuf + (uint32_t)(nps * 16) + (uint32_t)k * (uint32_t)fa.tb;
      uint32_t c2[8], t[8];
#pragma unroll
      for (int pr = 0; pr < 8; ++pr) {  // pair pr = sites 2pr, 2pr+1 share Philox word pr
        uint32_t offA, offB;
        if (AA) {
          offA = par[pr] & 0xffffu;
          offB = par[pr] >> 16;
        } else {
          offA = __byte_perm(par[pr >> 1], 0u, 0x4440u | ((pr & 1) * 2));
          offB = __byte_perm(par[pr >> 1], 0u, 0x4440u | ((pr & 1) * 2 + 1));
        }
        // entry address = row offset | 2*j.  Written as an OR: with an ADD ptxas folds the (warp-uniform) table base into a
        // 3-input IADD3 per draw instead of using [R + UR] addressing.
        uint32_t aA = ((w[pr] & L::JMASK) << 1) | offA;
        uint32_t aB = (((w[pr] >> 16) & L::JMASK) << 1) | offB;
        uint32_t eA = lds_u16(tbl + aA);
        uint32_t eB = lds_u16(tbl + aB);
        uint32_t ee = __byte_perm(eA, eB, 0x5410u);
        c2[pr] = __vminu2(w[pr], ee);
        t[pr] = w[pr] ^ ee;
      }
      // tie detection: some 16-bit lane of t has all its high (threshold) bits zero
      uint32_t m01 = __vimin3_u16x2(t[0], t[1], t[2]);
      uint32_t m23 = __vimin3_u16x2(t[3], t[4], t[5]);
      uint32_t m = __vimin3_u16x2(m01, m23, __vminu2(t[6], t[7]));
      if (((m & 0xffffu) < L::TIE) || ((m >> 16) < L::TIE)) {
        // exact path (rare): a draw tied in its threshold bits; resolve it with 32 more random bits.
        // No calls and no arrays here: either would make ptxas spill / re-materialise state on the FAST path.
        // The tied pairs are picked one by one out of the registers with a switch.
        uint32_t mask = 0;
#pragma unroll
        for (int pr = 0; pr < 8; ++pr)
          if (((t[pr] & 0xffffu) < L::TIE) || ((t[pr] >> 16) < L::TIE)) mask |= 1u << pr;
        constexpr uint32_t KPm1 = (1u << L::A) - 1u;
#pragma unroll 1
        while (mask) {
          const int pr = __ffs(mask) - 1;
          mask &= mask - 1;
          uint32_t tt = 0, ww = 0, cc = 0, oo = 0;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (q == pr) {
              tt = t[q]; ww = w[q]; cc = c2[q];
              oo = AA ? par[q] : (__byte_perm(par[q >> 1], 0u, 0x4440u | ((q & 1) * 2)) |
                                  (__byte_perm(par[q >> 1], 0u, 0x4440u | ((q & 1) * 2 + 1)) << 16));
            }
#pragma unroll 1
          for (int hh = 0; hh < 2; ++hh) {
            const uint32_t sh = (uint32_t)hh * 16u;
            if (((tt >> sh) & 0xffffu) < L::TIE) {
              const uint32_t x16 = (ww >> sh) & 0xffffu, off = (oo >> sh) & 0xffffu, j = x16 & KPm1;
              const uint32_t e = lds_u16(tbl + off + 2u * j);
              const uint64_t s = (uint64_t)sbase + 2 * pr + hh;
              const uint4 z = philox4x32<ROUNDS>(make_uint4((uint32_t)s, (uint32_t)(s >> 32), node, TAG_REFINE), key);
              // off / rowbytes = comp*K + state ; e16/qlo index = ((node*C + comp)*K + state)*KP + j
              const uint32_t row = AA ? (off >> 6) : (off >> 3);
              const int64_t idx = ((int64_t)node * (a.C * a.K) + row) * (KPm1 + 1) + j;
              const uint32_t child = z.x < __ldg(a.qlo + idx) ? j : (e & KPm1);
              cc = (cc & ~(0xffffu << sh)) | (child << sh);
            }
          }
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (q == pr) c2[q] = cc;
        }
      }
      if (AA) {
        constexpr uint32_t SM = 0x1f1f1f1fu;
        uint4 o;  // state bytes of the 16 sites: leaf output format = slot format
        o.x = __byte_perm(c2[0], c2[1], 0x6420u) & SM;
        o.y = __byte_perm(c2[2], c2[3], 0x6420u) & SM;
        o.z = __byte_perm(c2[4], c2[5], 0x6420u) & SM;
        o.w = __byte_perm(c2[6], c2[7], 0x6420u) & SM;
        if (dst >= 0) {
#pragma unroll
          for (int q = 0; q < 8; ++q) par[q] = (c2[q] & 0x001f001fu) * 64u + cbase[q];
          sts_v4(slot_base + (uint32_t)dst * slot_stride, o);
          reg_slot = dst;
        }
        if (leaf_row >= 0 && store_mode) store_leaf(a, leaf_base, leaf_row, col0, store_mode, o);
      } else {
        constexpr uint32_t SM = 0x03030303u;
        uint32_t raw[4];  // state bytes of the 16 sites (leaf output format)
#pragma unroll
        for (int q = 0; q < 4; ++q) raw[q] = __byte_perm(c2[2 * q], c2[2 * q + 1], 0x6420u) & SM;
        if (dst >= 0) {
#pragma unroll
          for (int q = 0; q < 4; ++q) par[q] = raw[q] * 8u + cbase[q];
          sts_v4(slot_base + (uint32_t)dst * slot_stride, make_uint4(par[0], par[1], par[2], par[3]));
          reg_slot = dst;
        }
        if (leaf_row >= 0 && store_mode) store_leaf(a, leaf_base, leaf_row, col0, store_mode, make_uint4(raw[0], raw[1], raw[2], raw[3]));
      }
    }
    // this warp is done with the stage buffer: hand it back to the producer
    __syncthreads();  // every warp is done with this stage buffer
    if (tid == 0 && st + nstage < fa.n_stages) {
      fence_async_smem();
      mbar_expect_tx(smem_u32(&full_bar[buf]), (uint32_t)fa.stage_bytes);
      tma_bulk_g2s(sbuf, fa.ftab + (size_t)(st + nstage) * fa.stage_bytes, (uint32_t)fa.stage_bytes, smem_u32(&full_bar[buf]));
    }
    if (++buf == nstage) { buf = 0; phase ^= 1u; }
  }
}

// Gather the per-node alias tables into walk order, interleaved with the walk records, one stage per nps nodes:
// stage = [nps x WalkRec (16 B)] [nps x tb table bytes].  The e16 layout [node][comp][state][column] u16 already is the
// row layout both kernels address (DNA: comp*32 + state*8 + j*2; protein: (comp*K + state)*64 + j*2).
__global__ void k_build_ftab(int N, int nps, int tb, int stage_bytes, const WalkRec *__restrict__ walk,
                             const uint16_t *__restrict__ e16, uint8_t *__restrict__ ftab) {
  const int chunks_per_node = (16 + tb) / 16;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)N * chunks_per_node) return;
  int k = (int)(i / chunks_per_node), ch = (int)(i % chunks_per_node);
  int st = k / nps, kk = k % nps;
  uint8_t *stage = ftab + (size_t)st * stage_bytes;
  if (ch == 0) {
    *reinterpret_cast<uint4 *>(stage + kk * 16) = *reinterpret_cast<const uint4 *>(&walk[k]);
  } else {
    int node = walk[k].node;
    const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint8_t *>(e16) + (size_t)node * tb) + (ch - 1);
    *reinterpret_cast<uint4 *>(stage + nps * 16 + (size_t)kk * tb + (ch - 1) * 16) = *src;
  }
}

int fast_tables_build(elx_handle *h) {
  FastTables *ft = static_cast<FastTables *>(h->fast);
  if (!ft) {
    ft = new FastTables();
    h->fast = ft;
  }
  int kind = 0, tb = 0, slot_bytes = 0;
  if (h->K == 4 && h->KP == 4 && h->C <= 8) {
    kind = 1; tb = h->C * 32; slot_bytes = 16;
  } else if (h->KP == 32 && (int64_t)h->C * h->K * 64 <= 65535) {
    kind = 2; tb = h->C * h->K * 64; slot_bytes = 16;
  }
  if (kind) {
    // stage geometry: keep a CTA under ~44 KB of shared memory so that 5 CTAs (20 warps) fit an SM; small tables
    // (DNA) stream 3 stages of up to 64 nodes, bigger ones 2 stages of a few nodes; tables that do not fit that
    // budget fall back to 2 stages x 1 node with fewer CTAs per SM, and to the generic kernel beyond 200 KB.
    const size_t slots = (size_t)h->n_slots * FAST_THREADS * slot_bytes;
    const size_t per_node = 16 + (size_t)tb;
    int nstage = tb <= 1024 ? MAX_NSTAGE : 2;
    size_t budget = 44 * 1024;
    if (const char *e = getenv("ELX_SMEM_BUDGET_KB")) budget = (size_t)atoi(e) * 1024;  // tuning knob
    int nps = budget > slots + 128 ? (int)((budget - slots - 128) / nstage / per_node) : 0;
    if (nps > 64) nps = 64;
    if (nps < 1) { nps = 1; nstage = 2; }
    size_t smem = 128 + (size_t)nstage * nps * per_node + slots;
    if (smem > 200 * 1024) kind = 0;
    ft->nps = nps; ft->nstage = nstage; ft->smem = smem;
  }
  ft->kind = kind;
  if (kind) {
    ft->tb = tb;
    ft->stage_bytes = ft->nps * (16 + tb);
    ft->n_stages = (h->N + ft->nps - 1) / ft->nps;
    size_t bytes = (size_t)ft->n_stages * ft->stage_bytes;
    if (!ft->d_ftab) {
      cudaError_t e = cudaMalloc((void **)&ft->d_ftab, bytes);
      if (e != cudaSuccess) return fail(h, ELX_E_CUDA, std::string("cudaMalloc(ftab): ") + cudaGetErrorString(e));
      cudaMemsetAsync(ft->d_ftab, 0, bytes, h->stream);
    }
    int chunks = (16 + tb) / 16;
    int64_t n = (int64_t)h->N * chunks;
    k_build_ftab<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->N, ft->nps, tb, ft->stage_bytes, h->d_walk, h->d_e16, ft->d_ftab);
    h->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, ELX_E_CUDA, std::string("k_build_ftab: ") + cudaGetErrorString(e));
  }
  return 0;
}

void fast_tables_free(elx_handle *h) {
  FastTables *ft = static_cast<FastTables *>(h->fast);
  if (!ft) return;
  cudaFree(ft->d_ftab);
  delete ft;
  h->fast = nullptr;
}

template <int ROUNDS, bool AA>
static cudaError_t launch_tiled(const FastArgs &fa, unsigned blocks, size_t smem, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(k2_tiled<ROUNDS, AA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k2_tiled<ROUNDS, AA><<<blocks, FAST_CTA, smem, stream>>>(fa);
  return cudaGetLastError();
}

int fast_simulate(elx_handle *h, const SimArgs &a, int *used) {
  *used = 0;
  FastTables *ft = static_cast<FastTables *>(h->fast);
  if (!ft || ft->kind == 0 || a.internal) return 0;
  FastArgs fa;
  fa.s = a;
  fa.ftab = ft->d_ftab;
  fa.tb = ft->tb;
  fa.stage_bytes = ft->stage_bytes;
  fa.n_stages = ft->n_stages;
  fa.nps = ft->nps;
  fa.nstage = ft->nstage;
  int64_t g0 = a.site0 >> 4, g1 = (a.site0 + a.nsites + 15) >> 4;
  unsigned blocks = (unsigned)((g1 - g0 + FAST_THREADS - 1) / FAST_THREADS);
  cudaError_t e;
  if (ft->kind == 1)
    e = h->rounds == 7 ? launch_tiled<7, false>(fa, blocks, ft->smem, h->stream) : launch_tiled<10, false>(fa, blocks, ft->smem, h->stream);
  else
    e = h->rounds == 7 ? launch_tiled<7, true>(fa, blocks, ft->smem, h->stream) : launch_tiled<10, true>(fa, blocks, ft->smem, h->stream);
  if (e != cudaSuccess) return fail(h, ELX_E_CUDA, std::string("k2_tiled launch: ") + cudaGetErrorString(e));
  h->launches++;
  *used = 1;
  return 0;
}

}  // namespace elx
