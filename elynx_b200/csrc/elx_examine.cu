// `slynx examine` core on the device: statistics of an alignment held as a [T][pitch] uint8 code matrix in HBM
// (slynx/src/SLynx/Examine/Examine.hs:52-160, elynx-seq/src/ELynx/Sequence/Alignment.hs:209-300,
// elynx-seq/src/ELynx/Alphabet/DistributionDiversity.hs:37-110, Sequence/Distance.hs:24-33, Sequence/Divergence.hs:38-64).
//
// Codes: 0..K-1 = standard characters in sorted order (what the simulator writes), K = '-', K+1 = '.' (the gap characters
// of DNAX / ProteinX; toStd of a gap is [], so gaps never enter a frequency).  IUPAC alphabets are out of scope.
//
// Kernels
//   k_examine_columns  one pass over the matrix (1 byte read per cell: HBM-read bound).  A thread owns 4 consecutive sites
//                      and walks the T rows with 32-bit loads; per-site character counts live in shared memory as
//                      [code][site-in-thread][thread] uint32 (conflict-free).  Per site: constant / soft-constant / gap-free
//                      flags, frequencies, entropy -> kEff, homoplasy -> kEff; sums are reduced per warp and added to a
//                      small accumulator block with atomics.
//   k_examine_hamming  numbers of EQUAL characters of all sequence pairs: 32 x 32 pair tiles, both row tiles of a 512-site
//                      chunk staged in shared memory (padded rows: conflict-free), packed byte compares (4 sites / word).
//   k_examine_divergence  K x K count matrix of one sequence pair per CTA (shared-memory histogram), on request.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/elynx_b200.h"
#include "elx_state.hpp"

namespace elx {

constexpr int EX_THREADS = 256;
constexpr int EX_SPT = 4;                        // sites per thread
constexpr int EX_BLOCK_SITES = EX_THREADS * EX_SPT;
constexpr int EX_MAX_K = 62;                     // + 2 gap codes = 64 counter rows
constexpr double EX_EPS = 1e-12;                 // DistributionDiversity.hs:37-38

struct ExamineAcc {
  unsigned long long n_constant, n_constant_soft, n_no_gaps, n_std, n_gap, n_bad;
  double keff_e_all, keff_e_nogap, keff_h_all, keff_h_nogap;
  unsigned long long n_h_inf_all;  // columns without any standard character: kEffHomoplasy = 1/0 = Infinity
  double dist[64];
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned warp_sum_u(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// tab[c] = log c and tab[T + 1 + c] = c log c (0 for c = 0), c = 0..T: the entropy of a column with counts c_k (sum s) is
// log s - (sum_k c_k log c_k) / s, so the per-site epilogue costs K table look-ups instead of K logarithms (xLogX's
// "0 below 1e-12" rule, DistributionDiversity.hs:41-45, can only fire for c_k = 0: frequencies are >= 1 / T).
__global__ void k_examine_tables(int T, double *__restrict__ tab) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > T) return;
  const double l = c ? log((double)c) : 0.0;
  tab[c] = l;
  tab[T + 1 + c] = (double)c * l;
}

__global__ void __launch_bounds__(EX_THREADS, 4) k_examine_columns(const uint8_t *__restrict__ a, int64_t pitch, int T, int64_t S, int K,
                                                                 ExamineAcc *__restrict__ acc, double *__restrict__ keff_site, const double *__restrict__ tab) {
  extern __shared__ uint32_t cnt[];  // [(K + 3)][EX_SPT][EX_THREADS]: K characters, two gap codes, one row for invalid codes
  const int tid = threadIdx.x, R = K + 2;
  const bool aligned = ((pitch & 3) == 0) && ((reinterpret_cast<uintptr_t>(a) & 3) == 0);
  unsigned n_const = 0, n_soft = 0, n_nogap = 0, n_hinf = 0;
  unsigned long long n_std = 0, n_gap = 0, n_bad = 0;
  double ke_all = 0, ke_ng = 0, kh_all = 0, kh_ng = 0;
  constexpr int UNROLL = 8;  // rows in flight per thread: the loop is load-latency bound otherwise
  for (int64_t base = (int64_t)blockIdx.x * EX_BLOCK_SITES; base < S; base += (int64_t)gridDim.x * EX_BLOCK_SITES) {
    const int64_t s0 = base + (int64_t)tid * EX_SPT;
    for (int r = 0; r < (R + 1) * EX_SPT; ++r) cnt[r * EX_THREADS + tid] = 0;
    const bool whole = aligned && s0 + EX_SPT <= S;
    auto load_row = [&](int t) -> uint32_t {
      const uint8_t *p = a + (int64_t)t * pitch + s0;
      if (whole) return __ldg(reinterpret_cast<const uint32_t *>(p));
      uint32_t w = 0;
#pragma unroll
      for (int d = 0; d < EX_SPT; ++d)
        if (s0 + d < S) w |= (uint32_t)__ldg(p + d) << (8 * d);
      return w;
    };
    const uint32_t first = load_row(0);  // per site byte lanes: code of row 0
    uint32_t differs = 0;                // nonzero byte lanes where some row differs from it
    const uint32_t rmax4 = (uint32_t)R * 0x01010101u, rchk4 = (uint32_t)(128 - R) * 0x01010101u;
    const uint32_t cnt_s = (uint32_t)__cvta_generic_to_shared(cnt) + (uint32_t)tid * 4u;
    // NR rows at once: loads first (NR requests in flight), validity check, then the counters
    const uint8_t *col = a + s0;
    auto rows = [&](auto NRc, auto WHOLEc, int t0) {
      constexpr int NR = decltype(NRc)::value;
      constexpr bool WHOLE = decltype(WHOLEc)::value;
      uint32_t w[NR];
      const uint8_t *p = col + (int64_t)t0 * pitch;
#pragma unroll
      for (int u = 0; u < NR; ++u) {  // a running pointer: one 64-bit add per row instead of a 64-bit multiply-add
        w[u] = WHOLE ? __ldg(reinterpret_cast<const uint32_t *>(p)) : load_row(t0 + u);
        p += pitch;
      }
      // a byte is a valid code iff it is < R: bit 7 of b | (b + 128 - R) is clear (no carries between valid bytes)
      uint32_t chk = 0;
#pragma unroll
      for (int u = 0; u < NR; ++u) chk |= w[u] | (w[u] + rchk4);
      if (chk & 0x80808080u) {
#pragma unroll
        for (int u = 0; u < NR; ++u) w[u] = __vminu4(w[u], rmax4);  // invalid codes land in the extra row R (and are reported)
      }
#pragma unroll
      for (int u = 0; u < NR; ++u) {
        differs |= w[u] ^ first;
#pragma unroll
        for (int d = 0; d < EX_SPT; ++d) {
          // one byte permute for the code, one multiply-add for the address, one shared-memory reduction (fire and forget:
          // every thread owns its counters, but a load / add / store chain would be serialised by possible aliasing)
          const uint32_t code = __byte_perm(w[u], 0u, 0x4440u + (uint32_t)d);
          const uint32_t addr = code * (uint32_t)(EX_SPT * EX_THREADS * 4) + (cnt_s + (uint32_t)(d * EX_THREADS * 4));
          asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
        }
      }
    };
    int t0 = 0;
    if (whole)
      for (; t0 + UNROLL <= T; t0 += UNROLL) rows(std::integral_constant<int, UNROLL>{}, std::true_type{}, t0);
    else
      for (; t0 + UNROLL <= T; t0 += UNROLL) rows(std::integral_constant<int, UNROLL>{}, std::false_type{}, t0);
    for (; t0 < T; ++t0) rows(std::integral_constant<int, 1>{}, std::false_type{}, t0);
#pragma unroll
    for (int d = 0; d < EX_SPT; ++d)
      if (s0 + d < S) n_bad += cnt[(R * EX_SPT + d) * EX_THREADS + tid];
    double dist_part[EX_SPT];
#pragma unroll
    for (int d = 0; d < EX_SPT; ++d) {
      const bool live = s0 + d < S;
      unsigned sum = 0, kinds = 0;
      for (int k = 0; k < K; ++k) {
        const unsigned c = cnt[(k * EX_SPT + d) * EX_THREADS + tid];
        sum += c;
        kinds += c ? 1u : 0u;
      }
      const unsigned gaps = cnt[(K * EX_SPT + d) * EX_THREADS + tid] + cnt[((K + 1) * EX_SPT + d) * EX_THREADS + tid];
      // entropy (DistributionDiversity.hs:41-57) = log s - (sum c log c) / s and homoplasy (:67-68) = (sum c^2) / s^2 of the
      // frequencies c / s (saveDivision, :92-94: all zero when the column has no standard character)
      double clc = 0.0;
      unsigned long long c2 = 0;
      for (int k = 0; k < K; ++k) {
        const unsigned c = cnt[(k * EX_SPT + d) * EX_THREADS + tid];
        clc += __ldg(tab + T + 1 + c);
        c2 += (unsigned long long)c * c;
      }
      const double e = sum ? __ldg(tab + sum) - clc / (double)sum : 0.0;
      const double h = sum ? (double)c2 / ((double)sum * (double)sum) : 0.0;
      const double keff_e = e < EX_EPS ? 1.0 : exp(e);          // kEffEntropy (:61-62)
      if (live) {
        const bool nogap = gaps == 0;
        n_const += ((differs >> (8 * d)) & 0xffu) == 0;                   // filterColsConstant (Alignment.hs:217-218)
        n_soft += kinds == 1;                                            // filterColsConstantSoft (:222-228)
        n_nogap += nogap;                                                // filterColsNoGaps (:247-248)
        n_std += sum; n_gap += gaps;
        ke_all += keff_e;
        if (nogap) ke_ng += keff_e;
        if (h > 0.0) { kh_all += 1.0 / h; if (nogap) kh_ng += 1.0 / h; } else ++n_hinf;  // kEffHomoplasy (:72-73)
        if (keff_site) keff_site[s0 + d] = keff_e;
      }
      dist_part[d] = live && sum ? 1.0 / (double)sum : 0.0;
    }
    // distribution (Alignment.hs:275-283): sum over sites of the per-site frequencies, one warp reduction per character
    for (int k = 0; k < K; ++k) {
      double v = 0.0;
#pragma unroll
      for (int d = 0; d < EX_SPT; ++d) v += (double)cnt[(k * EX_SPT + d) * EX_THREADS + tid] * dist_part[d];
      v = warp_sum(v);
      if ((tid & 31) == 0 && v != 0.0) atomicAdd(&acc->dist[k], v);
    }
  }
  ke_all = warp_sum(ke_all); ke_ng = warp_sum(ke_ng); kh_all = warp_sum(kh_all); kh_ng = warp_sum(kh_ng);
  n_const = warp_sum_u(n_const); n_soft = warp_sum_u(n_soft); n_nogap = warp_sum_u(n_nogap); n_hinf = warp_sum_u(n_hinf);
  atomicAdd(&acc->n_std, n_std);
  atomicAdd(&acc->n_gap, n_gap);
  if (n_bad) atomicAdd(&acc->n_bad, n_bad);
  if ((tid & 31) == 0) {
    atomicAdd(&acc->n_constant, (unsigned long long)n_const);
    atomicAdd(&acc->n_constant_soft, (unsigned long long)n_soft);
    atomicAdd(&acc->n_no_gaps, (unsigned long long)n_nogap);
    atomicAdd(&acc->n_h_inf_all, (unsigned long long)n_hinf);
    atomicAdd(&acc->keff_e_all, ke_all);
    atomicAdd(&acc->keff_e_nogap, ke_ng);
    atomicAdd(&acc->keff_h_all, kh_all);
    atomicAdd(&acc->keff_h_nogap, kh_ng);
  }
}

// K == 4 and no gaps (what the DNA simulator writes): the same statistics without shared memory.  A thread owns 8
// consecutive sites (one 64-bit load per row, 16 rows in flight); the counts of the four codes live in the BYTE LANES of
// registers -- one shift and four 3-input logic operations turn a word into four one-hot words, four adds accumulate
// them (2.6 instructions per cell with the checks); byte lanes are folded into 16-bit lanes every 240 rows (T <= 65535).
// A byte >= 4 anywhere (gap or invalid code: one OR per word to find out) raises `fallback` and the host reruns the
// generic kernel.  Per-site arithmetic is the generic kernel's, operation for operation.
constexpr int D4_SPT = 8, D4_BLOCK_SITES = EX_THREADS * D4_SPT;
__global__ void __launch_bounds__(EX_THREADS, 2) k_examine_columns_k4(const uint8_t *__restrict__ a, int64_t pitch, int T, int64_t S,
                                                                    ExamineAcc *__restrict__ acc, double *__restrict__ keff_site,
                                                                    int *__restrict__ fallback, const double *__restrict__ tab) {
  constexpr int K = 4, UNROLL = 16;
  constexpr uint32_t M = 0x01010101u, L16 = 0x00ff00ffu;
  const int tid = threadIdx.x;
  const bool aligned = ((pitch & 7) == 0) && ((reinterpret_cast<uintptr_t>(a) & 7) == 0);
  unsigned n_const = 0;
  unsigned long long n_std = 0;
  double ke_all = 0, kh_all = 0;
  uint32_t hi = 0;
  for (int64_t base = (int64_t)blockIdx.x * D4_BLOCK_SITES; base < S; base += (int64_t)gridDim.x * D4_BLOCK_SITES) {
    const int64_t s0 = base + (int64_t)tid * D4_SPT;
    const bool whole = aligned && s0 + D4_SPT <= S;
    const uint8_t *col = a + s0;
    auto load_slow = [&](int t) -> uint2 {
      const uint8_t *p = col + (int64_t)t * pitch;
      uint32_t w[2] = {0, 0};
#pragma unroll
      for (int d = 0; d < D4_SPT; ++d)
        if (s0 + d < S) w[d >> 2] |= (uint32_t)__ldg(p + d) << (8 * (d & 3));
      return make_uint2(w[0], w[1]);
    };
    const uint2 first = whole ? __ldg(reinterpret_cast<const uint2 *>(col)) : load_slow(0);
    uint32_t c8[4][2] = {}, c16[4][2][2] = {}, differs[2] = {0, 0};
    int pending = 0;  // rows accumulated in the byte lanes
    auto fold = [&]() {
#pragma unroll
      for (int wi = 0; wi < 2; ++wi)
#pragma unroll
        for (int k = 0; k < 4; ++k) { c16[k][wi][0] += c8[k][wi] & L16; c16[k][wi][1] += (c8[k][wi] >> 8) & L16; c8[k][wi] = 0; }
      pending = 0;
    };
    auto rows = [&](auto NRc, auto WHOLEc, int t0) {
      constexpr int NR = decltype(NRc)::value;
      constexpr bool WHOLE = decltype(WHOLEc)::value;
      uint2 w[NR];
      const uint8_t *p = col + (int64_t)t0 * pitch;
#pragma unroll
      for (int u = 0; u < NR; ++u) {
        w[u] = WHOLE ? __ldg(reinterpret_cast<const uint2 *>(p)) : load_slow(t0 + u);
        p += pitch;
      }
#pragma unroll
      for (int u = 0; u < NR; ++u) {
        const uint32_t x[2] = {w[u].x, w[u].y};
        const uint32_t f[2] = {first.x, first.y};
#pragma unroll
        for (int wi = 0; wi < 2; ++wi) {
          const uint32_t s1 = x[wi] >> 1;
          c8[3][wi] += x[wi] & s1 & M;
          c8[2][wi] += s1 & ~x[wi] & M;
          c8[1][wi] += x[wi] & ~s1 & M;
          c8[0][wi] += ~(x[wi] | s1) & M;
          differs[wi] |= x[wi] ^ f[wi];
          hi |= x[wi];
        }
      }
      pending += NR;
      if (pending > 239) fold();
    };
    int t0 = 0;
    if (whole)
      for (; t0 + UNROLL <= T; t0 += UNROLL) rows(std::integral_constant<int, UNROLL>{}, std::true_type{}, t0);
    else
      for (; t0 + UNROLL <= T; t0 += UNROLL) rows(std::integral_constant<int, UNROLL>{}, std::false_type{}, t0);
    for (; t0 < T; ++t0) rows(std::integral_constant<int, 1>{}, std::false_type{}, t0);
    fold();
    double dist_k[K] = {0, 0, 0, 0};
#pragma unroll
    for (int d = 0; d < D4_SPT; ++d) {
      if (s0 + d >= S) continue;
      const int wi = d >> 2, b = d & 3, sh = 16 * (b >> 1), hl = b & 1;
      unsigned cnt[K];
#pragma unroll
      for (int k = 0; k < K; ++k) cnt[k] = (c16[k][wi][hl] >> sh) & 0xffffu;
      unsigned sum = 0;
#pragma unroll
      for (int k = 0; k < K; ++k) sum += cnt[k];
      double clc = 0.0;  // as in k_examine_columns
      unsigned long long c2 = 0;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        clc += __ldg(tab + T + 1 + cnt[k]);
        c2 += (unsigned long long)cnt[k] * cnt[k];
      }
      const double e = __ldg(tab + sum) - clc / (double)sum;
      const double h = (double)c2 / ((double)sum * (double)sum);
      const double keff_e = e < EX_EPS ? 1.0 : exp(e);
      n_const += ((differs[wi] >> (8 * b)) & 0xffu) == 0;  // without gaps: constant <=> soft-constant
      n_std += sum;
      ke_all += keff_e;
      kh_all += 1.0 / h;                                   // sum = T >= 1: h > 0
      if (keff_site) keff_site[s0 + d] = keff_e;
      const double inv = 1.0 / (double)sum;
#pragma unroll
      for (int k = 0; k < K; ++k) dist_k[k] += (double)cnt[k] * inv;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const double v = warp_sum(dist_k[k]);
      if ((tid & 31) == 0 && v != 0.0) atomicAdd(&acc->dist[k], v);
    }
  }
  if (hi & 0xfcfcfcfcu) atomicExch(fallback, 1);
  ke_all = warp_sum(ke_all); kh_all = warp_sum(kh_all);
  n_const = warp_sum_u(n_const);
  atomicAdd(&acc->n_std, n_std);
  if ((tid & 31) == 0) {
    atomicAdd(&acc->n_constant, (unsigned long long)n_const);
    atomicAdd(&acc->n_constant_soft, (unsigned long long)n_const);
    atomicAdd(&acc->keff_e_all, ke_all);
    atomicAdd(&acc->keff_e_nogap, ke_all);
    atomicAdd(&acc->keff_h_all, kh_all);
    atomicAdd(&acc->keff_h_nogap, kh_all);
  }
}

// eq[i][j] += number of sites where sequences i and j carry the same character (any code, gaps included: Distance.hs:24-33
// compares characters).  Tile = 32 x 32 pairs; grid = (tile pairs, site slices); thread (r, c) owns pairs (r + 8 m, c).
constexpr int HM_TILE = 32, HM_CHUNK = 512, HM_STRIDE = HM_CHUNK + 4;  // padded rows: word stride 129 = 1 mod 32
__global__ void __launch_bounds__(256) k_examine_hamming(const uint8_t *__restrict__ a, int64_t pitch, int T, int64_t S, int n_tiles,
                                                         int n_slices, unsigned long long *__restrict__ eq) {
  __shared__ __align__(16) uint8_t rows_i[HM_TILE * HM_STRIDE];
  __shared__ __align__(16) uint8_t rows_j[HM_TILE * HM_STRIDE];
  const int ti = blockIdx.x / n_tiles, tj = blockIdx.x % n_tiles;
  if (tj < ti) return;
  const int tid = threadIdx.x, r = tid >> 5, c = tid & 31;
  unsigned long long acc[4] = {0, 0, 0, 0};
  const int64_t n_chunks = (S + HM_CHUNK - 1) / HM_CHUNK;
  for (int64_t ch = blockIdx.y; ch < n_chunks; ch += n_slices) {
    const int64_t s0 = ch * HM_CHUNK;
    __syncthreads();
    for (int i = tid; i < HM_TILE * HM_CHUNK; i += 256) {
      const int row = i / HM_CHUNK, col = i % HM_CHUNK;
      const int64_t s = s0 + col;
      const int gi = ti * HM_TILE + row, gj = tj * HM_TILE + row;
      // out-of-range cells get codes that never compare equal
      rows_i[row * HM_STRIDE + col] = (gi < T && s < S) ? __ldg(a + (int64_t)gi * pitch + s) : (uint8_t)0x7e;
      rows_j[row * HM_STRIDE + col] = (gj < T && s < S) ? __ldg(a + (int64_t)gj * pitch + s) : (uint8_t)0x7d;
    }
    __syncthreads();
    unsigned part[4] = {0, 0, 0, 0};
    const uint32_t *yj = reinterpret_cast<const uint32_t *>(rows_j + c * HM_STRIDE);
#pragma unroll 4
    for (int w = 0; w < HM_CHUNK / 4; ++w) {
      const uint32_t y = yj[w];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const uint32_t x = reinterpret_cast<const uint32_t *>(rows_i + (r + 8 * m) * HM_STRIDE)[w];
        // bytes are < 128 (codes <= 63, padding 0x7d / 0x7e), so z + 0x7f sets bit 7 of exactly the non-zero bytes and
        // never carries: XOR, add, mask, POPC -- 4 instructions per 4 sites (a __vcmpeq4 costs 6)
        const uint32_t z = x ^ y;
        part[m] += __popc((z + 0x7f7f7f7fu) & 0x80808080u);
      }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) acc[m] += (unsigned)HM_CHUNK - part[m];
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int gi = ti * HM_TILE + r + 8 * m, gj = tj * HM_TILE + c;
    if (gi < T && gj < T && acc[m]) atomicAdd(&eq[(int64_t)gi * T + gj], acc[m]);
  }
}

// Gap-free alignments over at most 4 characters (what the DNA simulator writes): 2 bits per site, 16 sites per word, and
// a pair of words is compared with XOR / OR-shift / mask / POPC -- 5 instructions per 16 sites instead of 7 per 4.
__global__ void k_examine_pack2(const uint8_t *__restrict__ a, int64_t pitch, int T, int64_t S, int64_t W, uint32_t *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)T * W) return;
  const int t = (int)(i / W);
  const int64_t w = i % W, s0 = w * 16;
  const uint8_t *p = a + (int64_t)t * pitch + s0;
  uint32_t v = 0;
  if (s0 + 16 <= S && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
    const uint4 q = __ldg(reinterpret_cast<const uint4 *>(p));
    const uint32_t x[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t b = x[k] & 0x03030303u;  // four 2-bit codes in byte lanes -> 8 contiguous bits
      v |= ((b | (b >> 6) | (b >> 12) | (b >> 18)) & 0xffu) << (8 * k);
    }
  } else {
    for (int d = 0; d < 16; ++d)
      if (s0 + d < S) v |= (uint32_t)(__ldg(p + d) & 3u) << (2 * d);  // cells beyond S stay 0 in every row: they compare equal
  }
  out[i] = v;
}

constexpr int H2_CHUNK = 256, H2_STRIDE = H2_CHUNK + 1;  // words per row and chunk; padded: conflict-free
__global__ void __launch_bounds__(256) k_examine_hamming2(const uint32_t *__restrict__ a, int T, int64_t W, int n_tiles, int n_slices,
                                                          unsigned long long *__restrict__ eq) {
  extern __shared__ uint32_t h2_rows[];  // [2][HM_TILE][H2_STRIDE]
  uint32_t *rows_i = h2_rows, *rows_j = h2_rows + HM_TILE * H2_STRIDE;
  const int ti = blockIdx.x / n_tiles, tj = blockIdx.x % n_tiles;
  if (tj < ti) return;
  const int tid = threadIdx.x, r = tid >> 5, c = tid & 31;
  unsigned long long acc[4] = {0, 0, 0, 0};
  const int64_t n_chunks = (W + H2_CHUNK - 1) / H2_CHUNK;
  for (int64_t ch = blockIdx.y; ch < n_chunks; ch += n_slices) {
    const int64_t w0 = ch * H2_CHUNK;
    __syncthreads();
    for (int i = tid; i < HM_TILE * H2_CHUNK; i += 256) {
      const int row = i / H2_CHUNK, col = i % H2_CHUNK;
      const int64_t w = w0 + col;
      const int gi = ti * HM_TILE + row, gj = tj * HM_TILE + row;
      // words beyond W (or rows beyond T) are 0 in the i tile and all-ones in the j tile: 16 mismatches, no equal site
      rows_i[row * H2_STRIDE + col] = (gi < T && w < W) ? __ldg(a + (int64_t)gi * W + w) : 0u;
      rows_j[row * H2_STRIDE + col] = (gj < T && w < W) ? __ldg(a + (int64_t)gj * W + w) : 0xffffffffu;
    }
    __syncthreads();
    unsigned mis[4] = {0, 0, 0, 0};
    const uint32_t *yj = rows_j + c * H2_STRIDE;
#pragma unroll 4
    for (int w = 0; w < H2_CHUNK; ++w) {
      const uint32_t y = yj[w];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const uint32_t z = rows_i[(r + 8 * m) * H2_STRIDE + w] ^ y;
        mis[m] += __popc((z | (z >> 1)) & 0x55555555u);
      }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) acc[m] += (unsigned)(16 * H2_CHUNK) - mis[m];
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int gi = ti * HM_TILE + r + 8 * m, gj = tj * HM_TILE + c;
    if (gi < T && gj < T && acc[m]) atomicAdd(&eq[(int64_t)gi * T + gj], acc[m]);
  }
}

// the padding cells of the last word compare equal in every pair: take them out again
__global__ void k_examine_sub_pad(unsigned long long *__restrict__ eq, int T, unsigned long long pad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (int64_t)T * T) eq[i] -= eq[i] >= pad ? pad : eq[i];
}

// divergence (Divergence.hs:38-64): out[pair][x][y] = number of sites with standard characters x in sequence i and y in j,
// pairs (i < j) in the order of `tuples` (Examine.hs:165-167): (0,1), (0,2), ..., (1,2), ...
__global__ void __launch_bounds__(256) k_examine_divergence(const uint8_t *__restrict__ a, int64_t pitch, int T, int64_t S, int K,
                                                            unsigned long long *__restrict__ out) {
  extern __shared__ unsigned hist[];  // [K][K]
  // pair index -> (i, j)
  int64_t p = blockIdx.x;
  int i = 0;
  while (p >= T - 1 - i) { p -= T - 1 - i; ++i; }
  const int j = i + 1 + (int)p;
  for (int x = threadIdx.x; x < K * K; x += blockDim.x) hist[x] = 0;
  __syncthreads();
  const uint8_t *ri = a + (int64_t)i * pitch, *rj = a + (int64_t)j * pitch;
  unsigned long long *o = out + (int64_t)blockIdx.x * K * K;
  for (int64_t s0 = 0; s0 < S; s0 += (int64_t)1 << 30) {  // flush the 32-bit bins before they can overflow
    const int64_t s1 = s0 + ((int64_t)1 << 30) < S ? s0 + ((int64_t)1 << 30) : S;
    for (int64_t s = s0 + threadIdx.x; s < s1; s += blockDim.x) {
      const unsigned x = __ldg(ri + s), y = __ldg(rj + s);
      if (x < (unsigned)K && y < (unsigned)K) atomicAdd(&hist[x * K + y], 1u);
    }
    __syncthreads();
    for (int x = threadIdx.x; x < K * K; x += blockDim.x) { o[x] += hist[x]; hist[x] = 0; }
    __syncthreads();
  }
}

}  // namespace elx

using namespace elx;

extern "C" {

int elx_examine_device(const uint8_t *d_states, int64_t pitch, int32_t n_sequences, int64_t n_sites, int32_t n_states,
                       elx_examine_result *out, double *d_keff_entropy_per_site, int64_t *d_equal_counts, int64_t *d_divergence,
                       void *stream_) {
  if (!out) return fail(nullptr, ELX_E_INVALID, "elx_examine: result is NULL");
  memset(out, 0, sizeof *out);
  const int T = n_sequences, K = n_states;
  const int64_t S = n_sites;
  if (T < 1 || S < 0 || K < 2 || K > EX_MAX_K) return fail(nullptr, ELX_E_INVALID, "elx_examine: need n_sequences >= 1, n_sites >= 0, 2 <= n_states <= 62");
  if (!d_states && S > 0) return fail(nullptr, ELX_E_INVALID, "elx_examine: states buffer is NULL");
  if (pitch < S) return fail(nullptr, ELX_E_INVALID, "elx_examine: pitch < n_sites");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, ELX_E_CUDA, "no CUDA device available (elynx_b200 has no CPU fallback)");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
#define EX_CUDA(expr)                                                                                              \
  do {                                                                                                             \
    cudaError_t _e = (expr);                                                                                       \
    if (_e != cudaSuccess) { cudaFree(d_acc); cudaFree(d_tab); return fail(nullptr, ELX_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } \
  } while (0)
  ExamineAcc *d_acc = nullptr;
  double *d_tab = nullptr;  // log c, c log c (k_examine_tables)
  EX_CUDA(cudaMalloc((void **)&d_acc, sizeof(ExamineAcc)));
  EX_CUDA(cudaMemsetAsync(d_acc, 0, sizeof(ExamineAcc), stream));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  out->n_sites = S; out->n_sequences = T; out->n_states = K;
  ExamineAcc h;
  memset(&h, 0, sizeof h);
  if (S > 0) {
    EX_CUDA(cudaMalloc((void **)&d_tab, sizeof(double) * 2 * ((size_t)T + 1)));
    k_examine_tables<<<(unsigned)((T + 256) / 256), 256, 0, stream>>>(T, d_tab);
    const size_t smem = (size_t)(K + 3) * EX_SPT * EX_THREADS * sizeof(uint32_t);
    EX_CUDA(cudaFuncSetAttribute(k_examine_columns, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t nb = (S + EX_BLOCK_SITES - 1) / EX_BLOCK_SITES;
    int per_sm = (int)((size_t)200 * 1024 / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    const unsigned grid = (unsigned)(nb < (int64_t)sms * per_sm ? nb : (int64_t)sms * per_sm);
    bool generic = true;
    if (K == 4 && T <= 65535) {
      // gap-free DNA: the register-only kernel; a gap or an invalid code makes it raise a flag instead of a result
      int *d_flag = reinterpret_cast<int *>(&d_acc->n_bad);  // scratch: cleared again before the generic kernel runs
      const int64_t nb4 = (S + D4_BLOCK_SITES - 1) / D4_BLOCK_SITES;
      const unsigned grid4 = (unsigned)(nb4 < (int64_t)sms * 2 ? nb4 : (int64_t)sms * 2);
      k_examine_columns_k4<<<grid4, EX_THREADS, 0, stream>>>(d_states, pitch, T, S, d_acc, d_keff_entropy_per_site, d_flag, d_tab);
      EX_CUDA(cudaGetLastError());
      ExamineAcc probe4;
      EX_CUDA(cudaMemcpyAsync(&probe4, d_acc, sizeof probe4, cudaMemcpyDeviceToHost, stream));
      EX_CUDA(cudaStreamSynchronize(stream));
      generic = probe4.n_bad != 0;
      if (generic) EX_CUDA(cudaMemsetAsync(d_acc, 0, sizeof(ExamineAcc), stream));
      else {
        const unsigned long long ns = (unsigned long long)S;
        EX_CUDA(cudaMemcpyAsync(&d_acc->n_no_gaps, &ns, sizeof ns, cudaMemcpyHostToDevice, stream));  // every column is gap-free
      }
    }
    if (generic) {
      k_examine_columns<<<grid, EX_THREADS, smem, stream>>>(d_states, pitch, T, S, K, d_acc, d_keff_entropy_per_site, d_tab);
      EX_CUDA(cudaGetLastError());
    }
  }
  std::vector<unsigned long long> eq;
  unsigned long long *d_eq = reinterpret_cast<unsigned long long *>(d_equal_counts), *d_eq_own = nullptr;
  if (T >= 2 && S > 0) {
    if (!d_eq) {
      EX_CUDA(cudaMalloc((void **)&d_eq_own, sizeof(unsigned long long) * (size_t)T * T));
      d_eq = d_eq_own;
    }
    cudaError_t e = cudaMemsetAsync(d_eq, 0, sizeof(unsigned long long) * (size_t)T * T, stream);
    const int n_tiles = (T + HM_TILE - 1) / HM_TILE;
    const int64_t n_chunks = (S + HM_CHUNK - 1) / HM_CHUNK;
    const int64_t tile_pairs = (int64_t)n_tiles * (n_tiles + 1) / 2;
    int64_t slices = ((int64_t)sms * 8 + tile_pairs - 1) / tile_pairs;  // enough CTAs to fill the GPU, few atomics
    if (slices > n_chunks) slices = n_chunks;
    if (slices < 1) slices = 1;
    if (slices > 65535) slices = 65535;
    // gap-free alignments over <= 4 characters take the 2-bit path; whether there are gaps is known after the column pass
    bool two_bit = false;
    if (e == cudaSuccess && K <= 4) {
      ExamineAcc probe;
      e = cudaMemcpyAsync(&probe, d_acc, sizeof probe, cudaMemcpyDeviceToHost, stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
      two_bit = e == cudaSuccess && probe.n_gap == 0 && probe.n_bad == 0;
    }
    uint32_t *d_packed = nullptr;
    if (e == cudaSuccess && two_bit) {
      const int64_t W = (S + 15) / 16;
      e = cudaMalloc((void **)&d_packed, sizeof(uint32_t) * (size_t)T * W);
      if (e == cudaSuccess) {
        const int64_t n = (int64_t)T * W;
        k_examine_pack2<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_states, pitch, T, S, W, d_packed);
        const int64_t chunks2 = (W + H2_CHUNK - 1) / H2_CHUNK;
        int64_t sl = slices < chunks2 ? slices : chunks2;
        if (sl < 1) sl = 1;
        const size_t smem2 = sizeof(uint32_t) * 2 * HM_TILE * H2_STRIDE;
        e = cudaFuncSetAttribute(k_examine_hamming2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e == cudaSuccess) {
          k_examine_hamming2<<<dim3((unsigned)(n_tiles * n_tiles), (unsigned)sl), 256, smem2, stream>>>(d_packed, T, W, n_tiles, (int)sl, d_eq);
          // every chunk counts 16 * H2_CHUNK cells per pair: the cells past S (zero padding of the last word, and the
          // words past W of the last chunk, which never compare equal) -- only the former are equal
          const unsigned long long pad = (unsigned long long)(16 * W - S);
          if (pad) k_examine_sub_pad<<<(unsigned)(((int64_t)T * T + 255) / 256), 256, 0, stream>>>(d_eq, T, pad);
          e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        cudaFree(d_packed);
      }
    } else if (e == cudaSuccess) {
      k_examine_hamming<<<dim3((unsigned)(n_tiles * n_tiles), (unsigned)slices), 256, 0, stream>>>(d_states, pitch, T, S, n_tiles, (int)slices, d_eq);
      e = cudaGetLastError();
    }
    eq.resize((size_t)T * T);
    if (e == cudaSuccess) e = cudaMemcpyAsync(eq.data(), d_eq, sizeof(unsigned long long) * (size_t)T * T, cudaMemcpyDeviceToHost, stream);
    if (e != cudaSuccess) { cudaFree(d_eq_own); cudaFree(d_acc); cudaFree(d_tab); return fail(nullptr, ELX_E_CUDA, std::string("elx_examine (hamming): ") + cudaGetErrorString(e)); }
  }
  if (d_divergence && T >= 2) {
    const int64_t pairs = (int64_t)T * (T - 1) / 2;
    if (pairs > 0x7fffffff) { cudaFree(d_eq_own); cudaFree(d_acc); cudaFree(d_tab); return fail(nullptr, ELX_E_UNSUPPORTED, "elx_examine: too many sequence pairs for divergence matrices"); }
    cudaError_t e = cudaMemsetAsync(d_divergence, 0, sizeof(int64_t) * (size_t)pairs * K * K, stream);
    if (e == cudaSuccess && S > 0) {
      k_examine_divergence<<<(unsigned)pairs, 256, sizeof(unsigned) * K * K, stream>>>(d_states, pitch, T, S, K,
                                                                                         reinterpret_cast<unsigned long long *>(d_divergence));
      e = cudaGetLastError();
    }
    if (e != cudaSuccess) { cudaFree(d_eq_own); cudaFree(d_acc); cudaFree(d_tab); return fail(nullptr, ELX_E_CUDA, std::string("elx_examine (divergence): ") + cudaGetErrorString(e)); }
  }
  cudaError_t e = cudaMemcpyAsync(&h, d_acc, sizeof h, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  cudaFree(d_eq_own);
  cudaFree(d_acc);
  cudaFree(d_tab);
  if (e != cudaSuccess) return fail(nullptr, ELX_E_CUDA, std::string("elx_examine: ") + cudaGetErrorString(e));
#undef EX_CUDA
  if (h.n_bad) return fail(nullptr, ELX_E_INVALID, "elx_examine: the matrix holds codes above n_states + 1");
  out->n_constant = (int64_t)h.n_constant; out->n_constant_soft = (int64_t)h.n_constant_soft;
  out->n_no_gaps = (int64_t)h.n_no_gaps; out->n_only_std = (int64_t)h.n_no_gaps;
  out->n_std_chars = (int64_t)h.n_std; out->n_gaps = (int64_t)h.n_gap;
  const double nan = std::nan(""), inf = INFINITY;
  for (int k = 0; k < K; ++k) out->distribution[k] = S ? h.dist[k] / (double)S : 0.0;
  out->keff_entropy_mean = S ? h.keff_e_all / (double)S : nan;
  out->keff_entropy_mean_no_gaps = h.n_no_gaps ? h.keff_e_nogap / (double)h.n_no_gaps : nan;
  out->keff_homoplasy_mean = S ? (h.n_h_inf_all ? inf : h.keff_h_all / (double)S) : nan;
  out->keff_homoplasy_mean_no_gaps = h.n_no_gaps ? h.keff_h_nogap / (double)h.n_no_gaps : nan;
  // pairwise Hamming distances per site: mean / maximum-likelihood variance / min / max over all pairs (Examine.hs:121-131)
  out->hamming_mean = out->hamming_var = out->hamming_min = out->hamming_max = nan;
  if (T >= 2 && S > 0) {
    double sum = 0, mn = inf, mx = -inf;
    const double np = (double)T * (T - 1) / 2;
    for (int i = 0; i < T; ++i)
      for (int j = i + 1; j < T; ++j) {
        const double d = (double)(S - (int64_t)eq[(size_t)i * T + j]) / (double)S;
        sum += d; mn = d < mn ? d : mn; mx = d > mx ? d : mx;
      }
    const double mean = sum / np;
    double var = 0;
    for (int i = 0; i < T; ++i)
      for (int j = i + 1; j < T; ++j) {
        const double d = (double)(S - (int64_t)eq[(size_t)i * T + j]) / (double)S - mean;
        var += d * d;
      }
    out->hamming_mean = mean; out->hamming_var = var / np; out->hamming_min = mn; out->hamming_max = mx;
  }
  return 0;
}

int elx_examine(const uint8_t *states, int64_t pitch, int32_t n_sequences, int64_t n_sites, int32_t n_states,
                elx_examine_result *out, double *keff_entropy_per_site, int64_t *equal_counts, int64_t *divergence) {
  if (!out) return fail(nullptr, ELX_E_INVALID, "elx_examine: result is NULL");
  if (n_sequences < 1 || n_sites < 0 || pitch < n_sites || (!states && n_sites > 0))
    return fail(nullptr, ELX_E_INVALID, "elx_examine: bad matrix description");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, ELX_E_CUDA, "no CUDA device available (elynx_b200 has no CPU fallback)");
  const int T = n_sequences, K = n_states;
  const int64_t S = n_sites, dp = (S + 127) / 128 * 128;
  uint8_t *d_a = nullptr;
  double *d_ke = nullptr;
  int64_t *d_eq = nullptr, *d_div = nullptr;
  const int64_t pairs = (int64_t)T * (T - 1) / 2;
  auto done = [&](int rc) { cudaFree(d_a); cudaFree(d_ke); cudaFree(d_eq); cudaFree(d_div); return rc; };
  auto bad = [&](cudaError_t e) { return done(fail(nullptr, ELX_E_CUDA, std::string("elx_examine: ") + cudaGetErrorString(e))); };
  cudaError_t e = cudaMalloc((void **)&d_a, (size_t)(dp > 0 ? dp : 128) * T);
  if (e != cudaSuccess) return bad(e);
  if (S > 0) {
    e = cudaMemcpy2D(d_a, (size_t)dp, states, (size_t)pitch, (size_t)S, (size_t)T, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return bad(e);
  }
  if (keff_entropy_per_site && S > 0 && (e = cudaMalloc((void **)&d_ke, sizeof(double) * (size_t)S)) != cudaSuccess) return bad(e);
  if (equal_counts && (e = cudaMalloc((void **)&d_eq, sizeof(int64_t) * (size_t)T * T)) != cudaSuccess) return bad(e);
  if (divergence && pairs > 0 && K >= 2 && K <= EX_MAX_K &&
      (e = cudaMalloc((void **)&d_div, sizeof(int64_t) * (size_t)pairs * K * K)) != cudaSuccess)
    return bad(e);
  int rc = elx_examine_device(d_a, dp, n_sequences, n_sites, n_states, out, d_ke, d_eq, d_div, nullptr);
  if (rc) return done(rc);
  if (d_ke && (e = cudaMemcpy(keff_entropy_per_site, d_ke, sizeof(double) * (size_t)S, cudaMemcpyDeviceToHost)) != cudaSuccess) return bad(e);
  if (d_eq) {
    if (T >= 2 && S > 0) e = cudaMemcpy(equal_counts, d_eq, sizeof(int64_t) * (size_t)T * T, cudaMemcpyDeviceToHost);
    else memset(equal_counts, 0, sizeof(int64_t) * (size_t)T * T);
    if (e != cudaSuccess) return bad(e);
  }
  if (d_div && (e = cudaMemcpy(divergence, d_div, sizeof(int64_t) * (size_t)pairs * K * K, cudaMemcpyDeviceToHost)) != cudaSuccess) return bad(e);
  return done(0);
}

}  // extern "C"
