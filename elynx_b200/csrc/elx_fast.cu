// Tiled shared-memory simulation kernels (K2, free-running mode) for sm_100a.
//
// Same draw specification as k2_freerun_generic (elx_core.cu) -- the two must agree bit for bit -- but laid out for
// the machine:
//   * one thread owns 16 consecutive sites (two Philox4x32 calls per node = 16 x 16-bit draws) and writes each leaf
//     row with one 128-bit streaming store (a warp covers 512 contiguous bytes of the [taxon][site] row);
//   * the per-branch alias tables and the walk records are streamed through shared memory in stages of NPS nodes
//     by TMA bulk copies (cp.async.bulk + mbarrier, 3 stages in flight);
//   * live internal-node states sit in a shared-memory slot stack [slot][thread] (128-bit, conflict free); the
//     walk order (heavy child last) bounds the number of slots by floor(log2 N)+1;
//   * a draw is: PRMT (parent offset) + LOP3 (address) + LDS.U16 + half a PRMT/VIMNMX.U16x2/LOP3 -- the alias
//     decision is a packed 16-bit min: child bits = low bits of min(random, entry); ties in the high bits (2^-13)
//     fall to an exact 32-bit refinement path.
//
// DNA kernel: K = 4, KP = 4, C <= 8.  Per-site state byte = comp*32 + state*8 = byte offset of the table row.
#include <cstdint>
#include <cstdio>
#include <string>

#include <cuda_runtime.h>

#include "../../include/elynx_b200.h"
#include "elx_device.cuh"
#include "elx_state.hpp"

namespace elx {

constexpr int NPS = 64;      // nodes per TMA stage
constexpr int NSTAGE = 3;    // stages in flight
constexpr int FAST_THREADS = 128;

struct FastTables {
  int kind = 0;  // 0 = none, 1 = DNA tiled
  uint8_t *d_ftab = nullptr;
  int tb = 0;           // table bytes per node
  int stage_bytes = 0;  // NPS * (16 + tb)
  int n_stages = 0;
};

struct FastArgs {
  SimArgs s;
  const uint8_t *ftab;
  int tb, stage_bytes, n_stages;
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

// One pair of draws (two sites sharing one Philox word).  P = packed parent offsets of the quad, dA/dB = byte index.
// Returns the packed min (child bits at 1..2 of each half) and the tie detector t = w ^ ee.
template <int DA, int DB>
__device__ __forceinline__ void draw_pair(uint32_t tbl, uint32_t P, uint32_t w, uint32_t &c2, uint32_t &t) {
  uint32_t offA = __byte_perm(P, 0u, 0x4440u | DA);
  uint32_t offB = __byte_perm(P, 0u, 0x4440u | DB);
  uint32_t aA = offA | (w & 6u);
  uint32_t aB = offB | ((w >> 16) & 6u);
  uint32_t eA = lds_u16(tbl + aA);
  uint32_t eB = lds_u16(tbl + aB);
  uint32_t ee = __byte_perm(eA, eB, 0x5410u);
  c2 = __vminu2(w, ee);
  t = w ^ ee;
}

// Exact resolution of a tie in the 13 high bits (DNA layout): 32 more random bits against the low threshold word.
template <int ROUNDS>
__device__ __noinline__ uint32_t refine_draw(const SimArgs &a, uint32_t tbl, uint32_t off, uint32_t x16, uint64_t site,
                                             uint32_t node, uint2 key) {
  const uint32_t j = (x16 >> 1) & 3u;
  const uint32_t e = lds_u16(tbl + (off | (j << 1)));
  const uint4 z = philox4x32<ROUNDS>(make_uint4((uint32_t)site, (uint32_t)(site >> 32), node, TAG_REFINE), key);
  const int64_t idx = (((int64_t)node * a.C + (off >> 5)) * 4 + ((off >> 3) & 3u)) * 4 + j;
  return z.x < __ldg(a.qlo + idx) ? j : ((e >> 1) & 3u);
}

// ---------------------------------------------------------------------------------------------------
// DNA tiled kernel
// ---------------------------------------------------------------------------------------------------
template <int ROUNDS>
__global__ void __launch_bounds__(FAST_THREADS) k2_dna_tiled(FastArgs fa) {
  extern __shared__ __align__(128) uint8_t smem[];
  const SimArgs &a = fa.s;
  const int tid = threadIdx.x;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
  uint8_t *stage0 = smem + 128;
  const uint32_t stage_smem = smem_u32(stage0);
  const uint32_t slots_smem = stage_smem + (uint32_t)(NSTAGE * fa.stage_bytes);
  const uint32_t my_slot0 = slots_smem + (uint32_t)tid * 16u;
  const uint32_t slot_stride = FAST_THREADS * 16u;

  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  __syncthreads();
  if (tid == 0) {
    for (int i = 0; i < NSTAGE && i < fa.n_stages; ++i) {
      mbar_expect_tx(smem_u32(&bars[i]), (uint32_t)fa.stage_bytes);
      tma_bulk_g2s(stage_smem + (uint32_t)(i * fa.stage_bytes), fa.ftab + (size_t)i * fa.stage_bytes, (uint32_t)fa.stage_bytes,
                   smem_u32(&bars[i]));
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
  uint32_t cmask[4], par[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t cm = 0, pr = 0;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      uint64_t s = (uint64_t)sbase + q * 4 + d;
      int c = 0;
      if (a.is_mixture) {
        uint4 x = philox4x32<ROUNDS>(make_uint4((uint32_t)s, (uint32_t)(s >> 32), 0u, TAG_COMP), key);
        c = cat_cum(a.cumw, a.C, u53_positive(x.x, x.y));
        if (c < 0) { atomicExch(a.error_flag, a.N + 1); c = 0; }
      }
      uint4 x = philox4x32<ROUNDS>(make_uint4((uint32_t)s, (uint32_t)(s >> 32), 0u, TAG_ROOT), key);
      int root = cat_cum(a.cumpi + c * 4, 4, u53_positive(x.x, x.y));
      if (root < 0) { atomicExch(a.error_flag, a.N + 2); root = 0; }
      cm |= (uint32_t)(c * 32) << (8 * d);
      pr |= (uint32_t)(c * 32 + root * 8) << (8 * d);
      int64_t col = col0 + q * 4 + d;
      if (a.comps && col >= 0 && col < a.nsites) a.comps[col] = c;
    }
    cmask[q] = cm;
    par[q] = pr;
  }
  sts_v4(my_slot0, make_uint4(par[0], par[1], par[2], par[3]));
  int reg_slot = 0;  // slot whose contents are mirrored in par[]

  const bool vec_store = full && ((a.leaves_pitch & 15) == 0) && ((reinterpret_cast<uintptr_t>(a.leaves) & 15) == 0) &&
                         ((col0 & 15) == 0);

  for (int st = 0; st < fa.n_stages; ++st) {
    const int buf = st % NSTAGE;
    mbar_wait(smem_u32(&bars[buf]), (uint32_t)((st / NSTAGE) & 1));
    const uint32_t sbuf = stage_smem + (uint32_t)(buf * fa.stage_bytes);
    const int nn = min(NPS, a.N - st * NPS);
    for (int k = 0; k < nn; ++k) {
      const uint4 rec = lds_v4(sbuf + (uint32_t)k * 16u);
      const uint32_t node = rec.x;
      const int src = (int)(int16_t)(rec.y & 0xffffu);
      const int dst = (int)(int16_t)(rec.y >> 16);
      const int leaf_row = (int)rec.z;
      if (src != reg_slot) {
        uint4 v = lds_v4(my_slot0 + (uint32_t)src * slot_stride);
        par[0] = v.x; par[1] = v.y; par[2] = v.z; par[3] = v.w;
        reg_slot = src;
      }
      const uint4 x0 = philox4x32<ROUNDS>(make_uint4(G0lo, Ghi, node, TAG_NODE), key);
      const uint4 x1 = philox4x32<ROUNDS>(make_uint4(G1lo, Ghi, node, TAG_NODE), key);
      const uint32_t tbl = sbuf + (uint32_t)(NPS * 16) + (uint32_t)k * (uint32_t)fa.tb;
      uint32_t c2[8], t[8];
      draw_pair<0, 1>(tbl, par[0], x0.x, c2[0], t[0]);
      draw_pair<2, 3>(tbl, par[0], x0.y, c2[1], t[1]);
      draw_pair<0, 1>(tbl, par[1], x0.z, c2[2], t[2]);
      draw_pair<2, 3>(tbl, par[1], x0.w, c2[3], t[3]);
      draw_pair<0, 1>(tbl, par[2], x1.x, c2[4], t[4]);
      draw_pair<2, 3>(tbl, par[2], x1.y, c2[5], t[5]);
      draw_pair<0, 1>(tbl, par[3], x1.z, c2[6], t[6]);
      draw_pair<2, 3>(tbl, par[3], x1.w, c2[7], t[7]);
      uint32_t raw[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) raw[q] = __byte_perm(c2[2 * q], c2[2 * q + 1], 0x6420u);
      // tie detection: some 16-bit lane of t has all 13 high bits zero  <=>  lane < 8
      uint32_t m01 = __vimin3_u16x2(t[0], t[1], t[2]);
      uint32_t m23 = __vimin3_u16x2(t[3], t[4], t[5]);
      uint32_t m = __vimin3_u16x2(m01, m23, __vminu2(t[6], t[7]));
      if (((m & 0xffffu) < 8u) || ((m >> 16) < 8u)) {
        // exact path (rare: 16 * 2^-13 per thread-node): redo the tied draws with the 32-bit refinement.
        // Fully unrolled with static indices so that par[] / raw[] stay in registers on the fast path.
        const uint32_t w[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t tl = (t[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
          if (tl < 8u) {
            const uint32_t x16 = (w[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
            const uint32_t off = (par[i >> 2] >> ((i & 3) * 8)) & 0xffu;
            const uint32_t child = refine_draw<ROUNDS>(a, tbl, off, x16, (uint64_t)sbase + i, node, key);
            const uint32_t sh = (i & 3) * 8;
            raw[i >> 2] = (raw[i >> 2] & ~(0xffu << sh)) | ((child << 1) << sh);
          }
        }
      }
      if (dst >= 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) par[q] = ((raw[q] & 0x06060606u) << 2) | cmask[q];
        sts_v4(my_slot0 + (uint32_t)dst * slot_stride, make_uint4(par[0], par[1], par[2], par[3]));
        reg_slot = dst;
      }
      if (leaf_row >= 0 && any) {
        uint4 o;
        o.x = (raw[0] >> 1) & 0x03030303u;
        o.y = (raw[1] >> 1) & 0x03030303u;
        o.z = (raw[2] >> 1) & 0x03030303u;
        o.w = (raw[3] >> 1) & 0x03030303u;
        uint8_t *dstp = a.leaves + (int64_t)leaf_row * a.leaves_pitch + col0;
        if (vec_store) {
          stg_cs_v4(dstp, o);
        } else {
          const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
          for (int i = 0; i < 16; ++i) {
            int64_t col = col0 + i;
            if (col >= 0 && col < a.nsites) dstp[i] = (uint8_t)(ow[i >> 2] >> ((i & 3) * 8));
          }
        }
      }
    }
    __syncthreads();  // every warp is done with this stage buffer
    if (tid == 0 && st + NSTAGE < fa.n_stages) {
      fence_async_smem();
      mbar_expect_tx(smem_u32(&bars[buf]), (uint32_t)fa.stage_bytes);
      tma_bulk_g2s(sbuf, fa.ftab + (size_t)(st + NSTAGE) * fa.stage_bytes, (uint32_t)fa.stage_bytes, smem_u32(&bars[buf]));
    }
  }
}

// Gather the per-node alias tables into walk order, interleaved with the walk records, one stage per NPS nodes:
// stage = [NPS x WalkRec (16 B)] [NPS x tb table bytes].
__global__ void k_build_ftab(int N, int tb, int stage_bytes, const WalkRec *__restrict__ walk, const uint16_t *__restrict__ e16,
                             uint8_t *__restrict__ ftab) {
  const int chunks_per_node = (16 + tb) / 16;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)N * chunks_per_node) return;
  int k = (int)(i / chunks_per_node), ch = (int)(i % chunks_per_node);
  int st = k / NPS, kk = k % NPS;
  uint8_t *stage = ftab + (size_t)st * stage_bytes;
  if (ch == 0) {
    *reinterpret_cast<uint4 *>(stage + kk * 16) = *reinterpret_cast<const uint4 *>(&walk[k]);
  } else {
    int node = walk[k].node;
    const uint4 *src = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint8_t *>(e16) + (size_t)node * tb) + (ch - 1);
    *reinterpret_cast<uint4 *>(stage + NPS * 16 + (size_t)kk * tb + (ch - 1) * 16) = *src;
  }
}

int fast_tables_build(elx_handle *h) {
  FastTables *ft = static_cast<FastTables *>(h->fast);
  if (!ft) {
    ft = new FastTables();
    h->fast = ft;
  }
  if (h->K == 4 && h->C <= 8 && h->KP == 4) {
    ft->kind = 1;
    ft->tb = h->C * 32;
    ft->stage_bytes = NPS * (16 + ft->tb);
    ft->n_stages = (h->N + NPS - 1) / NPS;
    size_t bytes = (size_t)ft->n_stages * ft->stage_bytes;
    if (!ft->d_ftab) {
      cudaError_t e = cudaMalloc((void **)&ft->d_ftab, bytes);
      if (e != cudaSuccess) return fail(h, ELX_E_CUDA, std::string("cudaMalloc(ftab): ") + cudaGetErrorString(e));
      cudaMemsetAsync(ft->d_ftab, 0, bytes, h->stream);
    }
    int chunks = (16 + ft->tb) / 16;
    int64_t n = (int64_t)h->N * chunks;
    k_build_ftab<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->N, ft->tb, ft->stage_bytes, h->d_walk, h->d_e16, ft->d_ftab);
    h->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, ELX_E_CUDA, std::string("k_build_ftab: ") + cudaGetErrorString(e));
  } else {
    ft->kind = 0;
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
  size_t smem = 128 + (size_t)NSTAGE * ft->stage_bytes + (size_t)h->n_slots * FAST_THREADS * 16;
  if (smem > 200 * 1024) return 0;  // fall back to the generic kernel
  int64_t g0 = a.site0 >> 4, g1 = (a.site0 + a.nsites + 15) >> 4;
  int64_t groups = g1 - g0;
  unsigned blocks = (unsigned)((groups + FAST_THREADS - 1) / FAST_THREADS);
  cudaError_t e;
  if (h->rounds == 7) {
    e = cudaFuncSetAttribute(k2_dna_tiled<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) k2_dna_tiled<7><<<blocks, FAST_THREADS, smem, h->stream>>>(fa);
  } else {
    e = cudaFuncSetAttribute(k2_dna_tiled<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) k2_dna_tiled<10><<<blocks, FAST_THREADS, smem, h->stream>>>(fa);
  }
  if (e != cudaSuccess) return fail(h, ELX_E_CUDA, std::string("k2_dna_tiled attribute: ") + cudaGetErrorString(e));
  h->launches++;
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(h, ELX_E_CUDA, std::string("k2_dna_tiled launch: ") + cudaGetErrorString(e));
  *used = 1;
  return 0;
}

}  // namespace elx
