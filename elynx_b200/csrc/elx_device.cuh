// Device-side building blocks shared by the kernels: Philox4x32, the free-running draw spec,
// and the table layout.  See DESIGN.md "free-running draw" for the specification these implement.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace elx {

// Philox counter tags (ctr.w)
constexpr uint32_t TAG_NODE = 0;    // ctr = (site>>3 lo, site>>3 hi, node, TAG_NODE): 8 x 16-bit draws, field = site & 7
constexpr uint32_t TAG_REFINE = 1;  // ctr = (site lo, site hi, node, TAG_REFINE): word 0 = 32 refinement bits
constexpr uint32_t TAG_COMP = 2;    // ctr = (site lo, site hi, 0, TAG_COMP): words 0,1 -> 53-bit uniform
constexpr uint32_t TAG_ROOT = 3;    // ctr = (site lo, site hi, 0, TAG_ROOT)

constexpr uint32_t PHILOX_M0 = 0xD2511F53u, PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u, PHILOX_W1 = 0xBB67AE85u;

template <int ROUNDS>
__host__ __device__ __forceinline__ uint4 philox4x32(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(PHILOX_M0, c.x), lo0 = PHILOX_M0 * c.x;
    uint32_t hi1 = __umulhi(PHILOX_M1, c.z), lo1 = PHILOX_M1 * c.z;
#else
    uint64_t p0 = (uint64_t)PHILOX_M0 * c.x, p1 = (uint64_t)PHILOX_M1 * c.z;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += PHILOX_W0;
    k.y += PHILOX_W1;
  }
  return c;
}

// The same function with the ten round keys (k + r * W) precomputed by the host: passed inside a kernel parameter block
// they are constant-bank operands of the XORs instead of twenty live registers per thread.
struct PhiloxRoundKeys {
  uint32_t k[20];
};
inline PhiloxRoundKeys philox_round_keys(uint64_t seed) {
  PhiloxRoundKeys rk;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) { rk.k[2 * r] = k0; rk.k[2 * r + 1] = k1; k0 += PHILOX_W0; k1 += PHILOX_W1; }
  return rk;
}
template <int ROUNDS>
__device__ __forceinline__ uint4 philox4x32_rk(uint4 c, const PhiloxRoundKeys &rk) {
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const uint32_t hi0 = __umulhi(PHILOX_M0, c.x), lo0 = PHILOX_M0 * c.x;
    const uint32_t hi1 = __umulhi(PHILOX_M1, c.z), lo1 = PHILOX_M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ rk.k[2 * r], lo1, hi0 ^ c.w ^ rk.k[2 * r + 1], lo0);
  }
  return c;
}

// Three calls that differ only in ctr.w (cw[0..2]) with ctr.x fixed per thread and ctr.z (the record's node id) the same
// for every thread -- the shape of the quad / duo record loops.  The first two rounds factor into a per-thread part
// (products of ctr.x, computed once in front of the record loop) and a per-record part whose one product M1 * ctr.z
// arrives precomputed in the stage header: 16 + 1/3 wide multiplies per call instead of 18 + 1/3.  Same bits as
// philox4x32_rk (algebra only); ROUNDS >= 2.
__device__ __forceinline__ void mul_wide(uint32_t m, uint32_t x, uint32_t &hi, uint32_t &lo) {
#ifdef ELX_PHILOX_AUTO_MUL
  hi = __umulhi(m, x);
  lo = m * x;
#else
  uint64_t p;
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(m), "r"(x));  // one IMAD.WIDE (ptxas otherwise splits some of them into IMAD.HI + IMAD)
  hi = (uint32_t)(p >> 32);
  lo = (uint32_t)p;
#endif
}
struct PhiloxThreadPart {
  uint32_t A[3], B[3], w1;
};
template <int ROUNDS>
__device__ __forceinline__ PhiloxThreadPart philox_thread_part(uint32_t cx, uint32_t cw0, uint32_t cw1, uint32_t cw2, const PhiloxRoundKeys &rk) {
  static_assert(ROUNDS >= 2, "the split form needs two rounds");
  PhiloxThreadPart t;
  const uint32_t hi0 = __umulhi(PHILOX_M0, cx);
  t.w1 = PHILOX_M0 * cx;
  const uint32_t cw[3] = {cw0, cw1, cw2};
#pragma unroll
  for (int cl = 0; cl < 3; ++cl) {
    const uint32_t z1 = hi0 ^ cw[cl] ^ rk.k[1];
    t.A[cl] = __umulhi(PHILOX_M1, z1);
    t.B[cl] = PHILOX_M1 * z1;
  }
  return t;
}
// X[0..11] = the twelve words of the three calls; (nhi, nlo) = hi / lo of M1 * ctr.z, cy = ctr.y
template <int ROUNDS>
__device__ __forceinline__ void philox3_split(const PhiloxThreadPart &t, uint32_t nhi, uint32_t nlo, uint32_t cy, const PhiloxRoundKeys &rk,
                                              uint32_t *X) {
  const uint32_t a1 = nhi ^ cy ^ rk.k[0];
  uint32_t H, L;
  mul_wide(PHILOX_M0, a1, H, L);
  H ^= t.w1 ^ rk.k[3];
  const uint32_t u = nlo ^ rk.k[2];
#pragma unroll
  for (int cl = 0; cl < 3; ++cl) {
    uint4 c = make_uint4(t.A[cl] ^ u, t.B[cl], H, L);
#pragma unroll
    for (int r = 2; r < ROUNDS; ++r) {
      uint32_t hi0, lo0, hi1, lo1;
      mul_wide(PHILOX_M0, c.x, hi0, lo0);
      mul_wide(PHILOX_M1, c.z, hi1, lo1);
      c = make_uint4(hi1 ^ c.y ^ rk.k[2 * r], lo1, hi0 ^ c.w ^ rk.k[2 * r + 1], lo0);
    }
    X[4 * cl] = c.x; X[4 * cl + 1] = c.y; X[4 * cl + 2] = c.z; X[4 * cl + 3] = c.w;
  }
}

// 53-bit uniform in (0,1] from two Philox words (used for the component and root-state draws).
__host__ __device__ __forceinline__ double u53_positive(uint32_t lo, uint32_t hi) {
  uint64_t w = ((uint64_t)hi << 32) | lo;
  return (double)((w >> 11) + 1) * 0x1p-53;
}

// Alias-table geometry.  KP = number of columns (power of two >= K), A = log2 KP.
// 16-bit draw x = [r : 16-A bits][j : A bits]; table entry e = [q_hi : 16-A][alias : A].
// Full acceptance threshold q = q_hi * 2^32 + q_lo in [0, 2^(48-A)); accept column j iff r*2^32 + x_lo < q.
__host__ __device__ __forceinline__ int alias_bits(int K) {
  int a = 1;
  while ((1 << a) < K) ++a;
  return a < 2 ? 2 : a;
}

}  // namespace elx
