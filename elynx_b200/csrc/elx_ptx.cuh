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


__device__ __forceinline__ int cat_cum(const double *__restrict__ cv, int K, double u) {
  double p = __dmul_rn(cv[K - 1], u);
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
