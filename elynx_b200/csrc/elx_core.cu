// elynx_b200 core: handle, kernels K1 (P(t) tables), K2 (tree simulation: injected + generic free-running),
// K3 (FASTA packer), validation histograms, and the C ABI of include/elynx_b200.h.
//
// Reference behaviour restated here (paths relative to /root/reference):
//   probMatrix                 elynx-markov/src/ELynx/Simulate/MarkovProcess.hs:33-41
//   jump = categorical (p ! i) elynx-markov/src/ELynx/Simulate/MarkovProcess.hs:48-49
//   walk / draw order          elynx-markov/src/ELynx/Simulate/MarkovProcessAlongTree.hs:57-63,88-98,163-173,195-208
//   FASTA layout               elynx-seq/src/ELynx/Sequence/Export/Fasta.hs:24-36
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#if defined(__linux__)
#include <sys/mman.h>
#endif

#include <cuda_runtime.h>

#include "../../include/elynx_b200.h"
#include "elx_device.cuh"
#include "elx_internal.hpp"
#include "elx_ptx.cuh"
#include "elx_state.hpp"

namespace elx {

static thread_local std::string g_last_error;

// ---- device memory cache (see elx_state.hpp) -----------------------------------------------------------------------
#undef cudaMalloc
#undef cudaFree
namespace {
struct DevBlock { void *p; size_t bytes; };
struct DevCache {
  std::mutex mu;
  std::vector<DevBlock> free_blocks[16];            // per device
  size_t cached_bytes[16] = {};
  std::vector<std::pair<void *, std::pair<int, size_t>>> live;  // ptr -> (device, bytes) of blocks handed out
};
DevCache &dev_cache() { static DevCache c; return c; }
size_t dev_cache_cap() {
  static const size_t cap = []() { const char *e = getenv("ELX_DEVICE_CACHE_MB"); return (size_t)(e ? std::max(0, atoi(e)) : 8192) << 20; }();
  return cap;
}
}  // namespace

cudaError_t dev_malloc(void **p, size_t bytes) {
  if (!p) return cudaErrorInvalidValue;
  *p = nullptr;
  if (bytes == 0) bytes = 1;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  DevCache &c = dev_cache();
  if (dev >= 0 && dev < 16 && dev_cache_cap() > 0) {
    std::lock_guard<std::mutex> lk(c.mu);
    auto &fl = c.free_blocks[dev];
    size_t best = fl.size();
    for (size_t i = 0; i < fl.size(); ++i)  // the tightest block that fits without wasting more than a quarter
      if (fl[i].bytes >= bytes && fl[i].bytes <= bytes + bytes / 4 + 4096 && (best == fl.size() || fl[i].bytes < fl[best].bytes)) best = i;
    if (best != fl.size()) {
      *p = fl[best].p;
      c.live.push_back({*p, {dev, fl[best].bytes}});
      c.cached_bytes[dev] -= fl[best].bytes;
      fl.erase(fl.begin() + (long)best);
      return cudaSuccess;
    }
  }
  e = cudaMalloc(p, bytes);
  if (e == cudaErrorMemoryAllocation) {  // give the cache back and try again
    cudaGetLastError();
    dev_cache_trim();
    e = cudaMalloc(p, bytes);
  }
  if (e == cudaSuccess && dev >= 0 && dev < 16) {
    std::lock_guard<std::mutex> lk(c.mu);
    c.live.push_back({*p, {dev, bytes}});
  }
  return e;
}

cudaError_t dev_free(void *p) {
  if (!p) return cudaSuccess;
  DevCache &c = dev_cache();
  int dev = -1;
  size_t bytes = 0;
  {
    std::lock_guard<std::mutex> lk(c.mu);
    for (size_t i = 0; i < c.live.size(); ++i)
      if (c.live[i].first == p) { dev = c.live[i].second.first; bytes = c.live[i].second.second; c.live.erase(c.live.begin() + (long)i); break; }
  }
  if (dev < 0) return cudaFree(p);  // not one of ours
  int cur = dev;
  cudaGetDevice(&cur);
  if (cur != dev) cudaSetDevice(dev);
  cudaDeviceSynchronize();  // what cudaFree does implicitly: nothing in flight uses the block when it is handed out again
  bool keep = false;
  {
    std::lock_guard<std::mutex> lk(c.mu);
    if (c.cached_bytes[dev] + bytes <= dev_cache_cap()) {
      c.free_blocks[dev].push_back({p, bytes});
      c.cached_bytes[dev] += bytes;
      keep = true;
    }
  }
  cudaError_t e = keep ? cudaSuccess : cudaFree(p);
  if (cur != dev) cudaSetDevice(cur);
  return e;
}

void dev_cache_trim() {
  DevCache &c = dev_cache();
  int cur = 0;
  cudaGetDevice(&cur);
  for (int d = 0; d < 16; ++d) {
    std::vector<DevBlock> blocks;
    {
      std::lock_guard<std::mutex> lk(c.mu);
      blocks.swap(c.free_blocks[d]);
      c.cached_bytes[d] = 0;
    }
    if (blocks.empty()) continue;
    cudaSetDevice(d);
    for (const DevBlock &b : blocks) cudaFree(b.p);
  }
  cudaSetDevice(cur);
}
#define cudaMalloc(p, n) elx::dev_malloc((void **)(p), (n))
#define cudaFree(p) elx::dev_free((void *)(p))

int fail(elx_handle *h, int code, const std::string &msg) {
  if (h) h->last_error = msg;
  g_last_error = msg;
  return code;
}

#define CUDA_TRY(h, expr)                                                                         \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return fail(h, ELX_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));             \
  } while (0)

// ---------------------------------------------------------------------------------------------
// K1: batched P(t) = I + A diag(expm1(lambda t)) B for every (node, component), one thread per row.
// Writes the probabilities and the per-row cumulative sums (sequential left-to-right fp64 adds, the
// order of mwc-random's scanl1 (+)).  t == 0 gives the identity exactly (MarkovProcess.hs:36).
// Algorithmic bytes: N*C*K*K*16 written once.  Tiny and batched -> no tensor cores.
// ---------------------------------------------------------------------------------------------
__global__ void k1_ptables(int N, int C, int K, const double *__restrict__ brlen, const double *__restrict__ lambda,
                           const double *__restrict__ A, const double *__restrict__ B, double *__restrict__ P,
                           double *__restrict__ cum) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t rows = (int64_t)N * C * K;
  if (r >= rows) return;
  int i = (int)(r % K);
  int c = (int)((r / K) % C);
  int n = (int)(r / ((int64_t)K * C));
  double t = brlen[n];
  double *prow = P + r * K;
  double *crow = cum + r * K;
  double ak[64];
  if (t != 0.0) {
    const double *Ai = A + ((int64_t)c * K + i) * K;
    const double *lam = lambda + (int64_t)c * K;
    for (int k = 0; k < K; ++k) ak[k] = Ai[k] * expm1(lam[k] * t);
  }
  double acc = 0.0;
  for (int j = 0; j < K; ++j) {
    double p;
    if (t == 0.0) {
      p = (i == j) ? 1.0 : 0.0;
    } else {
      const double *Bc = B + (int64_t)c * K * K + j;
      double s = 0.0;
      for (int k = 0; k < K; ++k) s = fma(ak[k], Bc[(int64_t)k * K], s);
      p = (i == j ? 1.0 : 0.0) + s;
      if (!(p > 0.0)) p = 0.0;  // rounding noise on vanishing entries (also maps NaN to 0)
    }
    prow[j] = p;
    acc = (j == 0) ? p : __dadd_rn(acc, p);
    crow[j] = acc;
  }
}

// Rebuild the cumulative tables from caller-supplied probabilities (elx_ptables_set).
__global__ void k1_cum_from_p(int64_t rows, int K, const double *__restrict__ P, double *__restrict__ cum) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  double acc = 0.0;
  for (int j = 0; j < K; ++j) {
    double p = P[r * K + j];
    acc = (j == 0) ? p : __dadd_rn(acc, p);
    cum[r * K + j] = acc;
  }
}

// K1 (general path): P = expm(t Q) by scaled diagonal Pade of order 7 with squaring -- the algorithm hmatrix `expm`
// runs (Golub & Van Loan 11.3.1: j = max 0 (floor (log2 ||tQ||_inf)), A = tQ/2^j, F = D^-1 N, square j times), one CTA
// per (branch, component) matrix, everything in shared memory.  Used when a component is not reversible (asymmetric
// exchangeabilities or a zero stationary frequency: no symmetric eigensystem) or when ELX_FLAG_PADE asks for it.
__global__ void k1_pade(int N, int C, int K, const double *__restrict__ brlen, const double *__restrict__ Q, double *__restrict__ P) {
  extern __shared__ double sm[];
  const int KK = K * K;
  double *A = sm, *X = sm + KK, *Nn = sm + 2 * KK, *D = sm + 3 * KK, *T = sm + 4 * KK;
  __shared__ double red[32];
  __shared__ int s_piv;
  const int m = blockIdx.x, n = m / C, c = m % C;
  const double t = brlen[n];
  double *out = P + (int64_t)m * KK;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (t == 0.0) {
    for (int e = tid; e < KK; e += nt) out[e] = (e / K == e % K) ? 1.0 : 0.0;
    return;
  }
  const double *Qc = Q + (int64_t)c * KK;
  for (int e = tid; e < KK; e += nt) A[e] = t * Qc[e];
  __syncthreads();
  // infinity norm = max row sum of |a_ij|
  double rmax = 0.0;
  for (int i = tid; i < K; i += nt) {
    double r = 0.0;
    for (int j = 0; j < K; ++j) r += fabs(A[i * K + j]);
    rmax = fmax(rmax, r);
  }
  for (int o = 16; o > 0; o >>= 1) rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
  if ((tid & 31) == 0) red[tid >> 5] = rmax;
  __syncthreads();
  double nrm = 0.0;
  for (int w = 0; w < (nt + 31) / 32; ++w) nrm = fmax(nrm, red[w]);
  int jsq = nrm > 0.0 ? (int)floor(log2(nrm)) : 0;
  if (jsq < 0) jsq = 0;
  const double scale = ldexp(1.0, -jsq);
  for (int e = tid; e < KK; e += nt) {
    A[e] *= scale;
    double id = (e / K == e % K) ? 1.0 : 0.0;
    X[e] = id; Nn[e] = id; D[e] = id;
  }
  __syncthreads();
  const int q = 7;
  double cf = 1.0;
  for (int k = 1; k <= q; ++k) {
    cf = cf * (double)(q - k + 1) / (double)((2 * q - k + 1) * k);
    for (int e = tid; e < KK; e += nt) {  // T = A X
      const int i = e / K, j = e % K;
      double acc = 0.0;
      for (int l = 0; l < K; ++l) acc = fma(A[i * K + l], X[l * K + j], acc);
      T[e] = acc;
    }
    __syncthreads();
    const double sgn = (k & 1) ? -1.0 : 1.0;
    for (int e = tid; e < KK; e += nt) {
      const double x = T[e];
      X[e] = x;
      Nn[e] += cf * x;
      D[e] += sgn * cf * x;
    }
    __syncthreads();
  }
  // solve D F = N by Gaussian elimination with partial pivoting (in place: D -> U, N -> F)
  for (int p = 0; p < K; ++p) {
    if (tid == 0) {
      int best = p;
      double bv = fabs(D[p * K + p]);
      for (int i = p + 1; i < K; ++i) { double v = fabs(D[i * K + p]); if (v > bv) { bv = v; best = i; } }
      s_piv = best;
    }
    __syncthreads();
    const int piv = s_piv;
    if (piv != p)
      for (int j = tid; j < 2 * K; j += nt) {
        double *M = j < K ? D : Nn;
        const int jj = j < K ? j : j - K;
        double a = M[p * K + jj]; M[p * K + jj] = M[piv * K + jj]; M[piv * K + jj] = a;
      }
    __syncthreads();
    const double dinv = 1.0 / D[p * K + p];
    for (int i = p + 1 + tid; i < K; i += nt) T[i] = D[i * K + p] * dinv;  // multipliers
    __syncthreads();
    for (int e = tid; e < (K - p - 1) * 2 * K; e += nt) {
      const int i = p + 1 + e / (2 * K), j = e % (2 * K);
      double *M = j < K ? D : Nn;
      const int jj = j < K ? j : j - K;
      if (j < K && jj <= p) continue;
      M[i * K + jj] -= T[i] * M[p * K + jj];
    }
    __syncthreads();
  }
  for (int p = K - 1; p >= 0; --p) {  // back substitution, column-parallel
    for (int j = tid; j < K; j += nt) {
      double acc = Nn[p * K + j];
      for (int l = p + 1; l < K; ++l) acc -= D[p * K + l] * Nn[l * K + j];
      Nn[p * K + j] = acc / D[p * K + p];
    }
    __syncthreads();
  }
  double *F = Nn, *G = X;
  for (int it = 0; it < jsq; ++it) {
    for (int e = tid; e < KK; e += nt) {
      const int i = e / K, j = e % K;
      double acc = 0.0;
      for (int l = 0; l < K; ++l) acc = fma(F[i * K + l], F[l * K + j], acc);
      G[e] = acc;
    }
    __syncthreads();
    double *tmp = F; F = G; G = tmp;
  }
  for (int e = tid; e < KK; e += nt) {
    double p = F[e];
    out[e] = p > 0.0 ? p : 0.0;  // rounding noise on vanishing entries
  }
}

// Alias tables for the free-running sampler, built in exact integer arithmetic (one thread per row).
// Row masses m_i = round(p_i / sum p * 2^48) (residual added to the largest), column capacity 2^(48-A).
__global__ void k1_alias(int64_t rows, int K, int A, const double *__restrict__ P, uint16_t *__restrict__ e16,
                         uint32_t *__restrict__ qlo) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int KP = 1 << A;
  const int QB = 48 - A;
  const int64_t CAP = (int64_t)1 << QB;
  int64_t m[64];
  int alias[64];
  int64_t q[64];
  int small[64], large[64];
  double sum = 0.0;
  for (int j = 0; j < K; ++j) {
    double p = P[r * K + j];
    if (!(p > 0.0)) p = 0.0;
    sum += p;
  }
  int64_t tot = 0;
  int big = 0;
  for (int j = 0; j < KP; ++j) {
    double p = 0.0;
    if (j < K) { p = P[r * K + j]; if (!(p > 0.0)) p = 0.0; }
    int64_t v = sum > 0.0 ? (int64_t)llrint(p / sum * 0x1p48) : (j == 0 ? ((int64_t)1 << 48) : 0);
    m[j] = v;
    tot += v;
    if (v > m[big]) big = j;
  }
  m[big] += ((int64_t)1 << 48) - tot;
  int ns = 0, nl = 0;
  for (int j = KP - 1; j >= 0; --j) {
    alias[j] = j;
    q[j] = CAP;
    if (m[j] < CAP) small[ns++] = j; else large[nl++] = j;
  }
  while (ns > 0 && nl > 0) {
    int s = small[--ns];
    int l = large[--nl];
    q[s] = m[s];
    alias[s] = l;
    m[l] -= CAP - m[s];
    if (m[l] < CAP) small[ns++] = l; else large[nl++] = l;
  }
  // whatever is left holds exactly CAP: always accept (alias = self)
  for (int j = 0; j < KP; ++j) {
    int64_t qq = q[j];
    int al = alias[j];
    if (qq >= CAP) { qq = CAP - 1; al = j; }
    e16[r * KP + j] = (uint16_t)(((qq >> 32) << A) | (int64_t)al);
    qlo[r * KP + j] = (uint32_t)(qq & 0xffffffffu);
  }
}

void launch_k1_alias(elx_handle *h, int64_t rows, int K, int A, const double *P, uint16_t *e16, uint32_t *qlo) {
  k1_alias<<<(unsigned)((rows + 127) / 128), 128, 0, h->stream>>>(rows, K, A, P, e16, qlo);
  h->launches++;
}

// ---------------------------------------------------------------------------------------------
// K2 (injected-uniforms mode): one thread per site, reference semantics bit for bit.
// categorical: cv = row-cumulative table; p = cv[K-1] * u (one rounded fp64 multiply); first i with cv[i] >= p.
// Test mode only (reads 8 bytes per draw) -- never benchmarked.
// ---------------------------------------------------------------------------------------------
constexpr int MAX_SLOTS = 32;


__device__ __forceinline__ int categorical_cum(const double *__restrict__ cv, int K, double u) {
  double p = __dmul_rn(cv[K - 1], u);
  for (int i = 0; i < K; ++i)
    if (cv[i] >= p) return i;
  return -1;
}

__global__ void k2_injected(SimArgs a) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.nsites) return;
  uint8_t slot[MAX_SLOTS];
  int64_t row = 0;
  int comp = 0;
  if (a.is_mixture) {
    comp = categorical_cum(a.cumw, a.C, a.U[s]);
    if (comp < 0) { atomicExch(a.error_flag, a.N + 1); comp = 0; }
    row = 1;
  }
  if (a.comps) a.comps[s] = comp;
  int root = categorical_cum(a.cumpi + (int64_t)comp * a.K, a.K, a.U[row * a.u_pitch + s]);
  if (root < 0) { atomicExch(a.error_flag, a.N + 2); root = 0; }
  row += 1;
  slot[0] = (uint8_t)root;
  const int64_t KK = (int64_t)a.K * a.K;
  for (int k = 0; k < a.N; ++k) {
    WalkRec rec = a.walk[k];
    int par = slot[rec.src];
    const double *cv = a.cum + ((int64_t)rec.node * a.C + comp) * KK + (int64_t)par * a.K;
    int child = categorical_cum(cv, a.K, a.U[(row + rec.node) * a.u_pitch + s]);
    if (child < 0) { atomicExch(a.error_flag, 1 + rec.node); child = 0; }
    if (rec.dst >= 0) slot[rec.dst] = (uint8_t)child;
    if (rec.leaf_row >= 0) a.leaves[(int64_t)rec.leaf_row * a.leaves_pitch + s] = (uint8_t)child;
    if (a.internal) a.internal[(int64_t)rec.node * a.internal_pitch + s] = (uint8_t)child;
  }
}

// ---------------------------------------------------------------------------------------------
// K2 (free-running, generic): any K <= 64, any C; tables read through L1/L2.  One thread per group of
// 8 consecutive sites (= one Philox4x32 call per node).  The tiled kernels in elx_fast.cu implement
// the same draw specification for the shapes that fit shared memory; this one is the fallback for big
// mixtures and the cross-check.
// ---------------------------------------------------------------------------------------------
template <int ROUNDS>
__device__ __forceinline__ int draw_child(const SimArgs &a, int node, int comp, int par, uint32_t x16, uint64_t site,
                                          uint2 key) {
  const int KP = 1 << a.A;
  int j = x16 & (KP - 1);
  uint32_t r = x16 >> a.A;
  int64_t idx = (((int64_t)node * a.C + comp) * a.K + par) * KP + j;
  uint32_t e = __ldg(a.e16 + idx);
  uint32_t qh = e >> a.A;
  int al = e & (KP - 1);
  if (r < qh) return j;
  if (r > qh) return al;
  uint4 c = philox4x32<ROUNDS>(make_uint4((uint32_t)site, (uint32_t)(site >> 32), (uint32_t)node, TAG_REFINE), key);
  return c.x < __ldg(a.qlo + idx) ? j : al;
}

template <int ROUNDS>
__global__ void k2_freerun_generic(SimArgs a) {
  // groups are aligned on global site index multiples of 8
  int64_t g0 = a.site0 >> 3;
  int64_t g = g0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t sbase = g << 3;
  if (sbase >= a.site0 + a.nsites) return;
  const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
  uint64_t slot[MAX_SLOTS];
  int comp[8];
  uint64_t st = 0;
#pragma unroll
  for (int d = 0; d < 8; ++d) {
    uint64_t s = (uint64_t)sbase + d;
    int c = 0;
    if (a.is_mixture) {
      uint4 x = philox4x32<ROUNDS>(make_uint4((uint32_t)s, (uint32_t)(s >> 32), 0u, TAG_COMP), key);
      c = categorical_cum(a.cumw, a.C, u53_positive(x.x, x.y));
      if (c < 0) { atomicExch(a.error_flag, a.N + 1); c = 0; }
    }
    comp[d] = c;
    uint4 x = philox4x32<ROUNDS>(make_uint4((uint32_t)s, (uint32_t)(s >> 32), 0u, TAG_ROOT), key);
    int root = categorical_cum(a.cumpi + (int64_t)c * a.K, a.K, u53_positive(x.x, x.y));
    if (root < 0) { atomicExch(a.error_flag, a.N + 2); root = 0; }
    st |= (uint64_t)root << (8 * d);
    int64_t col = (int64_t)s - a.site0;
    if (a.comps && col >= 0 && col < a.nsites) a.comps[col] = c;
  }
  slot[0] = st;
  const bool full = (sbase >= a.site0) && (sbase + 8 <= a.site0 + a.nsites);
  const bool vec_ok = full && (((sbase - a.site0) & 7) == 0);
  for (int k = 0; k < a.N; ++k) {
    WalkRec rec = a.walk[k];
    uint64_t par = slot[rec.src];
    uint4 x = philox4x32<ROUNDS>(make_uint4((uint32_t)g, (uint32_t)((uint64_t)g >> 32), (uint32_t)rec.node, TAG_NODE), key);
    uint32_t w[4] = {x.x, x.y, x.z, x.w};
    uint64_t out = 0;
#pragma unroll
    for (int d = 0; d < 8; ++d) {
      uint32_t x16 = (w[d >> 1] >> ((d & 1) * 16)) & 0xffffu;
      int child = draw_child<ROUNDS>(a, rec.node, comp[d], (int)((par >> (8 * d)) & 0xff), x16, (uint64_t)sbase + d, key);
      out |= (uint64_t)child << (8 * d);
    }
    if (rec.dst >= 0) slot[rec.dst] = out;
    if (rec.leaf_row >= 0) {
      uint8_t *dst = a.leaves + (int64_t)rec.leaf_row * a.leaves_pitch + (sbase - a.site0);
      if (vec_ok && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0)) {
        *reinterpret_cast<uint64_t *>(dst) = out;
      } else {
        for (int d = 0; d < 8; ++d) {
          int64_t col = sbase + d - a.site0;
          if (col >= 0 && col < a.nsites) dst[d] = (uint8_t)(out >> (8 * d));
        }
      }
    }
    if (a.internal) {
      uint8_t *dst = a.internal + (int64_t)rec.node * a.internal_pitch + (sbase - a.site0);
      for (int d = 0; d < 8; ++d) {
        int64_t col = sbase + d - a.site0;
        if (col >= 0 && col < a.nsites) dst[d] = (uint8_t)(out >> (8 * d));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K3: FASTA packer.  states -> alphabet characters, written into the whole-file image.
// Algorithmic bytes: 2 per cell (1 read + 1 written).
// ---------------------------------------------------------------------------------------------
// Four state bytes -> four characters.  With at most 8 characters the alphabet sits in two registers and ONE byte permute
// is the look-up (its selector = the four states packed into nibbles: three logic/shift pairs); bigger alphabets go
// through the shared-memory table byte by byte.
__device__ __forceinline__ uint32_t k3_map4(uint32_t v, bool small, uint32_t lut_lo, uint32_t lut_hi, const uint8_t *slut) {
  if (small) {
    const uint32_t x = v & 0x07070707u;
    const uint32_t y = x | (x >> 4);                            // byte 0: s0 | s1 << 4, byte 2: s2 | s3 << 4
    const uint32_t sel = (y & 0xffu) | ((y >> 8) & 0xff00u);  // nibbles s0, s1, s2, s3
    return __byte_perm(lut_lo, lut_hi, sel);
  }
  return (uint32_t)slut[v & 63] | ((uint32_t)slut[(v >> 8) & 63] << 8) | ((uint32_t)slut[(v >> 16) & 63] << 16) |
         ((uint32_t)slut[(v >> 24) & 63] << 24);
}

__global__ void k3_fasta_body(const uint8_t *__restrict__ leaves, int64_t pitch, int64_t site0, int64_t nsites,
                              const int64_t *__restrict__ seq_off, const uint8_t *__restrict__ lut, int K,
                              uint8_t *__restrict__ out) {
  __shared__ __align__(8) uint8_t slut[64];
  if (threadIdx.x < 64) slut[threadIdx.x] = threadIdx.x < K ? lut[threadIdx.x] : (uint8_t)'?';
  __syncthreads();
  const bool small = K <= 8;
  const uint32_t lut_lo = reinterpret_cast<const uint32_t *>(slut)[0], lut_hi = reinterpret_cast<const uint32_t *>(slut)[1];
  int t = blockIdx.y;
  const uint8_t *src = leaves + (int64_t)t * pitch;
  uint8_t *dst = out + seq_off[t] + site0;
  // each thread owns 16 destination-aligned bytes where possible
  int64_t mis = (16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15;  // bytes until dst is 16-aligned
  int64_t chunk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;    // chunk 0 = head, chunks >= 1 aligned
  int64_t lo = chunk == 0 ? 0 : mis + (chunk - 1) * 16;
  int64_t hi = chunk == 0 ? mis : lo + 16;
  if (hi > nsites) hi = nsites;
  if (lo >= hi) return;
  // the source window [lo, lo + 16) is misaligned by m bytes -- the same m for every chunk of the row: two aligned
  // 128-bit loads and four funnel shifts instead of sixteen byte loads
  const int m = (int)(reinterpret_cast<uintptr_t>(src + lo) & 15);
  if (hi - lo == 16 && lo - m >= 0 && lo - m + 32 <= pitch) {
    const uint4 A = __ldg(reinterpret_cast<const uint4 *>(src + lo - m));
    const uint4 B = m ? __ldg(reinterpret_cast<const uint4 *>(src + lo - m + 16)) : A;
    const uint32_t sh = 8u * (uint32_t)(m & 3);
    uint32_t in[4];
    switch (m >> 2) {  // uniform over the CTA
      case 0: in[0] = __funnelshift_r(A.x, A.y, sh); in[1] = __funnelshift_r(A.y, A.z, sh); in[2] = __funnelshift_r(A.z, A.w, sh); in[3] = __funnelshift_r(A.w, B.x, sh); break;
      case 1: in[0] = __funnelshift_r(A.y, A.z, sh); in[1] = __funnelshift_r(A.z, A.w, sh); in[2] = __funnelshift_r(A.w, B.x, sh); in[3] = __funnelshift_r(B.x, B.y, sh); break;
      case 2: in[0] = __funnelshift_r(A.z, A.w, sh); in[1] = __funnelshift_r(A.w, B.x, sh); in[2] = __funnelshift_r(B.x, B.y, sh); in[3] = __funnelshift_r(B.y, B.z, sh); break;
      default: in[0] = __funnelshift_r(A.w, B.x, sh); in[1] = __funnelshift_r(B.x, B.y, sh); in[2] = __funnelshift_r(B.y, B.z, sh); in[3] = __funnelshift_r(B.z, B.w, sh); break;
    }
    uint4 o;
    o.x = k3_map4(in[0], small, lut_lo, lut_hi, slut); o.y = k3_map4(in[1], small, lut_lo, lut_hi, slut);
    o.z = k3_map4(in[2], small, lut_lo, lut_hi, slut); o.w = k3_map4(in[3], small, lut_lo, lut_hi, slut);
    __stcs(reinterpret_cast<uint4 *>(dst + lo), o);
  } else {
    for (int64_t i = lo; i < hi; ++i) dst[i] = slut[src[i] & 63];
  }
}

// slice mode: [T][out_pitch] characters, no frame.  The alphabet travels in the kernel's parameter block: the call only
// enqueues (no table upload, no host synchronisation), so a caller's batches overlap.
struct FastaLut {
  uint8_t c[64];
};
__global__ void k3_fasta_slice(const uint8_t *__restrict__ leaves, int64_t pitch, int64_t nsites, const FastaLut lut,
                               int K, uint8_t *__restrict__ out, int64_t out_pitch) {
  __shared__ __align__(8) uint8_t slut[64];
  if (threadIdx.x < 64) slut[threadIdx.x] = threadIdx.x < K ? lut.c[threadIdx.x] : (uint8_t)'?';
  __syncthreads();
  const bool small = K <= 8;
  const uint32_t lut_lo = reinterpret_cast<const uint32_t *>(slut)[0], lut_hi = reinterpret_cast<const uint32_t *>(slut)[1];
  int t = blockIdx.y;
  const uint8_t *src = leaves + (int64_t)t * pitch;
  uint8_t *dst = out + (int64_t)t * out_pitch;
  int64_t lo = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16, hi = lo + 16;
  if (hi > nsites) hi = nsites;
  if (lo >= hi) return;
  if (hi - lo == 16 && ((reinterpret_cast<uintptr_t>(dst + lo) & 15) == 0) && ((reinterpret_cast<uintptr_t>(src + lo) & 15) == 0)) {
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + lo));
    uint4 o;
    o.x = k3_map4(v.x, small, lut_lo, lut_hi, slut); o.y = k3_map4(v.y, small, lut_lo, lut_hi, slut);
    o.z = k3_map4(v.z, small, lut_lo, lut_hi, slut); o.w = k3_map4(v.w, small, lut_lo, lut_hi, slut);
    *reinterpret_cast<uint4 *>(dst + lo) = o;
  } else {
    for (int64_t i = lo; i < hi; ++i) dst[i] = slut[src[i] & 63];
  }
}

__global__ void k3_fasta_frame(int T, const int64_t *__restrict__ seq_off, const int64_t *__restrict__ name_off,
                               const uint8_t *__restrict__ names, int64_t total_sites, int write_head, int write_tail,
                               uint8_t *__restrict__ out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  if (write_head) {
    int64_t len = name_off[t + 1] - name_off[t];
    uint8_t *p = out + seq_off[t] - len - 2;
    p[0] = '>';
    for (int64_t i = 0; i < len; ++i) p[1 + i] = names[name_off[t] + i];
    p[1 + len] = '\n';
  }
  if (write_tail) out[seq_off[t] + total_sites] = '\n';
}

// ---------------------------------------------------------------------------------------------
// Validation histograms (accumulated into caller buffers; summed across GPUs with one NCCL all-reduce).
// ---------------------------------------------------------------------------------------------
// Per-taxon state counts.  Every thread owns one counter per state in shared memory ([state][thread]: conflict-free,
// bumped with fire-and-forget reductions) and reads 16 sites per 128-bit load; states >= K land in an overflow row that
// is never reported.  1 algorithmic byte read per cell.
__global__ void k_state_counts(const uint8_t *__restrict__ leaves, int64_t pitch, int64_t nsites, int K,
                               unsigned long long *__restrict__ counts) {
  extern __shared__ uint32_t hc[];  // [K + 1][blockDim.x]
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int k = 0; k <= K; ++k) hc[k * nt + tid] = 0;
  const int t = blockIdx.y;
  const uint8_t *src = leaves + (int64_t)t * pitch;
  const bool aligned = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(hc) + (uint32_t)tid * 4u, stride = (uint32_t)nt * 4u;
  auto bump = [&](uint32_t code) {
    code = min(code, (uint32_t)K);
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(base + code * stride) : "memory");
  };
  for (int64_t i = ((int64_t)blockIdx.x * nt + tid) * 16; i < nsites; i += (int64_t)gridDim.x * nt * 16) {
    if (aligned && i + 16 <= nsites) {
      const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src + i));
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        bump(w[q] & 0xffu); bump((w[q] >> 8) & 0xffu); bump((w[q] >> 16) & 0xffu); bump(w[q] >> 24);
      }
    } else {
      for (int64_t j = i; j < nsites && j < i + 16; ++j) bump(src[j]);
    }
  }
  __syncthreads();
  // one warp per state: sum the threads' counters
  for (int k = tid >> 5; k < K; k += nt >> 5) {
    unsigned long long sum = 0;
    for (int j = tid & 31; j < nt; j += 32) sum += hc[k * nt + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((tid & 31) == 0 && sum) atomicAdd(&counts[(int64_t)t * K + k], sum);
  }
}

__global__ void k_pattern_counts(const uint8_t *__restrict__ leaves, int64_t pitch, int64_t nsites, int T, int K,
                                 unsigned long long *__restrict__ counts) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nsites; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t idx = 0;
    for (int t = 0; t < T; ++t) idx = idx * K + leaves[(int64_t)t * pitch + i];
    atomicAdd(&counts[idx], 1ull);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// the component draw of every site alone (getComponentsAndRootStates, MarkovProcessAlongTree.hs:163-173, first half)
template <int ROUNDS>
__global__ void k_site_components(SimArgs a) {
  const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= a.nsites) return;
  const uint32_t cr = draw_comp_root<ROUNDS>(a.cumw, a.cumpi, a.C, a.K, a.N, a.is_mixture, a.error_flag, (uint64_t)(a.site0 + col),
                                             (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
  a.comps[col] = (int32_t)(cr & 0xffffu);
}

int ensure_pair_tables(elx_handle *h) {
  if (!h->pair || h->pair_ready) return 0;
  const int rc = pair_tables_build(h);
  if (!rc) h->pair_ready = 1;
  return rc;
}

int ensure_fast_tables(elx_handle *h) {
  if (h->fast_ready) return 0;
  const int rc = fast_tables_build(h);
  if (!rc) h->fast_ready = 1;
  return rc;
}

static int build_tables_from_p(elx_handle *h, bool have_p) {
  const int64_t rows = (int64_t)h->N * h->C * h->K;
  const int threads = 128;
  const unsigned blocks = (unsigned)((rows + threads - 1) / threads);
  if (!have_p && h->use_pade) {
    const size_t smem = sizeof(double) * 5 * h->K * h->K;
    cudaFuncSetAttribute(k1_pade, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k1_pade<<<(unsigned)(h->N * h->C), 128, smem, h->stream>>>(h->N, h->C, h->K, h->d_brlen, h->d_Q, h->d_P);
    h->launches++;
    k1_cum_from_p<<<blocks, threads, 0, h->stream>>>(rows, h->K, h->d_P, h->d_cum);
  } else if (!have_p) {
    k1_ptables<<<blocks, threads, 0, h->stream>>>(h->N, h->C, h->K, h->d_brlen, h->d_lambda, h->d_A, h->d_B, h->d_P, h->d_cum);
  } else {
    k1_cum_from_p<<<blocks, threads, 0, h->stream>>>(rows, h->K, h->d_P, h->d_cum);
  }
  h->launches++;
  k1_alias<<<blocks, threads, 0, h->stream>>>(rows, h->K, h->A, h->d_P, h->d_e16, h->d_qlo);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  // Only the table family the handle's default call shape needs is built now: quad rows (K <= 4, leaves-only calls), else
  // pair rows, else the node-by-node stage images.  The others follow on the first call that needs them (a K <= 4 call
  // that keeps the internal states -> pair rows; elx_pair_get), see ensure_pair_tables / ensure_fast_tables.
  h->pair_ready = 0;
  h->fast_ready = 0;
  int rc = 0;
  if (h->quad) rc = quad_tables_build(h);
  else if (h->duo) rc = duo_tables_build(h);
  else if (h->pair) rc = ensure_pair_tables(h);
  else rc = ensure_fast_tables(h);
  if (rc) return rc;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return 0;
}

template <typename T>
static cudaError_t dalloc(T **p, size_t n) {
  return cudaMalloc((void **)p, n * sizeof(T) > 0 ? n * sizeof(T) : 1);
}

static int check_sim_args(elx_handle *h, int64_t site0, int64_t nsites, const void *leaves, int64_t lp, const void *internal,
                          int64_t ip) {
  if (!h) return fail(nullptr, ELX_E_INVALID, "null handle");
  if (nsites < 0 || site0 < 0) return fail(h, ELX_E_INVALID, "simulate: negative site range");
  if (!leaves && nsites > 0) return fail(h, ELX_E_INVALID, "simulate: leaves buffer is NULL");
  if (lp < nsites) return fail(h, ELX_E_INVALID, "simulate: leaves_pitch < nsites");
  if (internal && ip < nsites) return fail(h, ELX_E_INVALID, "simulate: internal_pitch < nsites");
  return 0;
}

static int check_error_flag(elx_handle *h) {
  int flag = 0;
  CUDA_TRY(h, cudaMemcpyAsync(&flag, h->d_error_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  if (flag != 0) {
    int zero = 0;
    cudaMemcpyAsync(h->d_error_flag, &zero, sizeof(int), cudaMemcpyHostToDevice, h->stream);
    std::string what = flag == h->N + 1 ? "the component draw" : flag == h->N + 2 ? "the root-state draw"
                       : flag >= 1 && flag <= h->N ? "the draw at node " + std::to_string(flag - 1) + " (preorder index)"
                                                   : "kernel self-check " + std::to_string(flag);
    return fail(h, ELX_E_BAD_WEIGHTS, "categorical: bad weights! (" + what + ")");
  }
  return 0;
}

static SimArgs base_args(elx_handle *h) {
  SimArgs a;
  memset(&a, 0, sizeof a);
  a.N = h->N; a.K = h->K; a.C = h->C; a.is_mixture = h->is_mixture; a.n_slots = h->n_slots; a.A = h->A;
  a.rounds = h->rounds;
  a.walk = h->d_walk; a.cum = h->d_cum; a.cumw = h->d_cumw; a.cumpi = h->d_cumpi; a.e16 = h->d_e16; a.qlo = h->d_qlo;
  a.error_flag = h->d_error_flag;
  return a;
}

}  // namespace elx

using namespace elx;

extern "C" {

int elx_version(void) { return 100; }

void elx_trim_device_cache(void) { dev_cache_trim(); }

const char *elx_last_error(const elx_handle *h) { return h ? h->last_error.c_str() : g_last_error.c_str(); }

// One handle on ONE device: eigensystems, H2D of model + tree, K1, walk programs.  device_override >= 0 replaces
// opts->device (the peers of a multi-device handle).
static int create_on_device(const elx_model *model, const elx_tree *tree, const elx_opts *opts, int device_override, elx_handle **out) {
  if (!out) return fail(nullptr, ELX_E_INVALID, "elx_create: out is NULL");
  *out = nullptr;
  if (!model || !tree) return fail(nullptr, ELX_E_INVALID, "elx_create: model/tree is NULL");
  const int K = model->n_states, C = model->n_components, N = tree->n_nodes;
  if (K < 2 || K > 64) return fail(nullptr, ELX_E_INVALID, "model: n_states must be in 2..64");
  if (C < 1) return fail(nullptr, ELX_E_INVALID, "fromSubstitutionModels: No weights given.");
  if (!model->is_mixture && C != 1) return fail(nullptr, ELX_E_INVALID, "model: the single-model API takes exactly one component");
  if (!model->pi || !model->exch || (model->is_mixture && !model->weights))
    return fail(nullptr, ELX_E_INVALID, "model: NULL array");
  if (N < 1 || !tree->parent || !tree->brlen || !tree->leaf_row) return fail(nullptr, ELX_E_INVALID, "tree: empty or NULL arrays");
  for (int n = 0; n < N; ++n) {
    if (tree->brlen[n] < 0) return fail(nullptr, ELX_E_INVALID, "probMatrix: Time is negative.");
    if (!std::isfinite(tree->brlen[n])) return fail(nullptr, ELX_E_INVALID, "tree: branch length is not finite");
  }
  std::vector<double> w(C, 1.0);
  if (model->is_mixture) {
    double tot = 0;
    for (int c = 0; c < C; ++c) {
      w[c] = model->weights[c];
      if (!(w[c] >= 0) || !std::isfinite(w[c])) return fail(nullptr, ELX_E_INVALID, "model: weights must be finite and >= 0");
      tot += w[c];
    }
    if (!(tot > 0)) return fail(nullptr, ELX_E_BAD_WEIGHTS, "categorical: bad weights!");
  }
  for (int64_t i = 0; i < (int64_t)C * K; ++i)
    if (!(model->pi[i] >= 0) || !std::isfinite(model->pi[i])) return fail(nullptr, ELX_E_INVALID, "model: pi must be finite and >= 0");

  std::string msg;
  WalkProgram walk;
  int T = 0;
  if (!compile_walk(N, tree->parent, tree->leaf_row, walk, T, msg)) return fail(nullptr, ELX_E_INVALID, msg);
  if (walk.n_slots > MAX_SLOTS) return fail(nullptr, ELX_E_UNSUPPORTED, "tree: needs more than 32 live state slots");

  // Eigen path (reversible components) or, when any component has no symmetric eigensystem (asymmetric
  // exchangeabilities, a zero stationary frequency) or ELX_FLAG_PADE is set, the general Pade path on
  // Q = setDiagonal(E diag pi) (RateMatrix.hs:87-103).
  std::vector<double> lambda((size_t)C * K), A((size_t)C * K * K), B((size_t)C * K * K), Qh;
  // Default: the Pade kernel (hmatrix expm's own algorithm: element-wise parity <= 1e-12 with the reference's tables on
  // every branch length); ELX_FLAG_EIGEN selects the eigen form, which models with more than 32 states always take.
  const int oflags = opts ? opts->flags : 0;
  bool use_pade = (K <= 32 && !(oflags & ELX_FLAG_EIGEN)) || (oflags & ELX_FLAG_PADE);
  for (int c = 0; c < C && !use_pade; ++c)
    if (!eigen_reversible(K, model->pi + (size_t)c * K, model->exch + (size_t)c * K * K, &lambda[(size_t)c * K],
                          &A[(size_t)c * K * K], &B[(size_t)c * K * K], msg))
      use_pade = true;
  if (use_pade) {
    if (K > 32) return fail(nullptr, ELX_E_UNSUPPORTED, "model: the Pade path supports at most 32 states (" + msg + ")");
    Qh.assign((size_t)C * K * K, 0.0);
    for (int c = 0; c < C; ++c)
      for (int i = 0; i < K; ++i) {
        double rs = 0;
        for (int j = 0; j < K; ++j) {
          if (j == i) continue;
          double v = model->exch[((size_t)c * K + i) * K + j] * model->pi[(size_t)c * K + j];
          if (!std::isfinite(v)) return fail(nullptr, ELX_E_INVALID, "model: exchangeabilities must be finite");
          Qh[((size_t)c * K + i) * K + j] = v;
          rs += std::fabs(v);
        }
        Qh[((size_t)c * K + i) * K + i] = -rs;
      }
  }

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, ELX_E_CUDA, "no CUDA device available (elynx_b200 has no CPU fallback)");
  int dev = device_override >= 0 ? device_override : (opts && opts->device >= 0 ? opts->device : -1);
  if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) return fail(nullptr, ELX_E_CUDA, "cudaGetDevice failed"); }
  if (dev >= ndev) return fail(nullptr, ELX_E_INVALID, "opts: device ordinal out of range");

  elx_handle *h = new elx_handle();
  h->N = N; h->K = K; h->C = C; h->T = T; h->is_mixture = model->is_mixture ? 1 : 0;
  h->A = alias_bits(K); h->KP = 1 << h->A; h->device = dev; h->n_slots = walk.n_slots;
  h->rounds = (opts && opts->philox_rounds == 7) ? 7 : 10;
  h->flags = opts ? opts->flags : 0;
  h->use_pade = use_pade ? 1 : 0;
  h->weights = w;
  if (opts && opts->philox_rounds != 0 && opts->philox_rounds != 7 && opts->philox_rounds != 10) {
    delete h;
    return fail(nullptr, ELX_E_INVALID, "opts: philox_rounds must be 0, 7 or 10");
  }
  auto bail = [&](int code, const std::string &m) { std::string mm = m; elx_destroy(h); return fail(nullptr, code, mm); };
#define CT(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return bail(ELX_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } while (0)
  CT(cudaSetDevice(dev));
  CT(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  const size_t rows = (size_t)N * C * K;
  CT(dalloc(&h->d_brlen, (size_t)N));
  CT(dalloc(&h->d_lambda, (size_t)C * K));
  CT(dalloc(&h->d_A, (size_t)C * K * K));
  CT(dalloc(&h->d_B, (size_t)C * K * K));
  if (use_pade) {
    CT(dalloc(&h->d_Q, (size_t)C * K * K));
    CT(cudaMemcpy(h->d_Q, Qh.data(), sizeof(double) * Qh.size(), cudaMemcpyHostToDevice));
  }
  CT(dalloc(&h->d_P, rows * K));
  CT(dalloc(&h->d_cum, rows * K));
  CT(dalloc(&h->d_e16, rows * h->KP));
  CT(dalloc(&h->d_qlo, rows * h->KP));
  CT(dalloc(&h->d_cumw, (size_t)C));
  CT(dalloc(&h->d_cumpi, (size_t)C * K));
  CT(dalloc(&h->d_walk, (size_t)N));
  CT(dalloc(&h->d_error_flag, 1));
  CT(cudaMemset(h->d_error_flag, 0, sizeof(int)));
  // cumulative weights / stationary distributions: scanl1 (+) exactly as categorical does
  std::vector<double> cumw(C), cumpi((size_t)C * K);
  double acc = 0;
  for (int c = 0; c < C; ++c) { acc = c ? acc + w[c] : w[0]; cumw[c] = acc; }
  for (int c = 0; c < C; ++c) {
    acc = 0;
    for (int i = 0; i < K; ++i) { acc = i ? acc + model->pi[(size_t)c * K + i] : model->pi[(size_t)c * K]; cumpi[(size_t)c * K + i] = acc; }
    if (!(acc > 0)) return bail(ELX_E_BAD_WEIGHTS, "categorical: bad weights!");
  }
  CT(cudaMemcpy(h->d_brlen, tree->brlen, sizeof(double) * N, cudaMemcpyHostToDevice));
  CT(cudaMemcpy(h->d_lambda, lambda.data(), sizeof(double) * lambda.size(), cudaMemcpyHostToDevice));
  CT(cudaMemcpy(h->d_A, A.data(), sizeof(double) * A.size(), cudaMemcpyHostToDevice));
  CT(cudaMemcpy(h->d_B, B.data(), sizeof(double) * B.size(), cudaMemcpyHostToDevice));
  CT(cudaMemcpy(h->d_cumw, cumw.data(), sizeof(double) * C, cudaMemcpyHostToDevice));
  CT(cudaMemcpy(h->d_cumpi, cumpi.data(), sizeof(double) * cumpi.size(), cudaMemcpyHostToDevice));
  CT(cudaMemcpy(h->d_walk, walk.recs.data(), sizeof(WalkRec) * N, cudaMemcpyHostToDevice));
#undef CT
  // Pair mode: with K <= 4 the free-running kernels draw the children of a node two at a time (elx_pair.cu).
  if (K <= 4 && !(h->flags & ELX_FLAG_SINGLE_DRAWS)) {
    std::string pmsg;
    int prc = pair_create(h, tree->parent, tree->leaf_row, pmsg);
    if (prc) { std::string m = pmsg.empty() ? h->last_error : pmsg; elx_destroy(h); return fail(nullptr, prc, m); }
  }
  // Quad mode on top of it (leaves-only calls): one draw per cap of up to 4 frontier nodes (elx_quad.cu).
  if (K <= 4 && !(h->flags & (ELX_FLAG_SINGLE_DRAWS | ELX_FLAG_PAIR_DRAWS))) {
    int qrc = quad_create(h, tree->parent, tree->leaf_row);
    if (qrc) { std::string m = h->last_error; elx_destroy(h); return fail(nullptr, qrc, m); }
  }
  // Duo mode (amino acids, leaves-only calls): one draw per cap of up to 2 frontier nodes on component-bucketed sites
  // (elx_duo.cu); calls that keep the internal states draw node by node.
  if (K > 4 && !(h->flags & ELX_FLAG_SINGLE_DRAWS)) {
    int drc = duo_create(h, tree->parent, tree->leaf_row);
    if (drc) { std::string m = h->last_error; elx_destroy(h); return fail(nullptr, drc, m); }
  }
  int rc = build_tables_from_p(h, false);
  if (rc) { std::string m = h->last_error; elx_destroy(h); return fail(nullptr, rc, m); }
  *out = h;
  return 0;
}


// Multi-device handles (opts->n_devices): the reference's *Par calls parallelise INSIDE the call
// (MarkovProcessAlongTree.hs:109-125, :221-262: one chunk per capability); here one handle drives n_devices GPUs -- a full
// copy of the tables per device, one host thread + stream per device in the host-buffer calls.
int elx_create(const elx_model *model, const elx_tree *tree, const elx_opts *opts, elx_handle **out) {
  if (!out) return fail(nullptr, ELX_E_INVALID, "elx_create: out is NULL");
  *out = nullptr;
  int want = opts ? opts->n_devices : 0;
  if (want == 0 || want == 1) return create_on_device(model, tree, opts, -1, out);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, ELX_E_CUDA, "no CUDA device available (elynx_b200 has no CPU fallback)");
  int dev0 = opts->device >= 0 ? opts->device : 0;
  if (want < 0) want = ndev - dev0;  // all visible devices from `device` on
  if (want < 1 || dev0 + want > ndev)
    return fail(nullptr, ELX_E_INVALID, "opts: n_devices = " + std::to_string(opts->n_devices) + " from device " + std::to_string(dev0) +
                                            " exceeds the " + std::to_string(ndev) + " visible device(s)");
  std::vector<elx_handle *> hs((size_t)want, nullptr);
  std::vector<int> rcs((size_t)want, 0);
  std::vector<std::string> msgs((size_t)want);
  std::vector<std::thread> th;
  for (int d = 0; d < want; ++d)
    th.emplace_back([&, d]() {
      rcs[d] = create_on_device(model, tree, opts, dev0 + d, &hs[d]);
      if (rcs[d]) msgs[d] = g_last_error;  // thread-local: carry it over to the caller's thread
    });
  for (auto &t : th) t.join();
  for (int d = 0; d < want; ++d)
    if (rcs[d]) {
      const int rc = rcs[d];
      const std::string m = msgs[d];
      for (elx_handle *x : hs) elx_destroy(x);
      return fail(nullptr, rc, m);
    }
  for (int d = 1; d < want; ++d) { hs[d]->is_peer = 1; hs[0]->peers.push_back(hs[d]); }
  *out = hs[0];
  return 0;
}

int32_t elx_n_devices(const elx_handle *h) { return h ? 1 + (int32_t)h->peers.size() : 0; }

int elx_shard_range(int64_t nsites, int32_t n_shards, int32_t shard, int32_t align, int64_t *first, int64_t *count) {
  if (nsites < 0 || n_shards < 1 || shard < 0 || shard >= n_shards || align < 1 || !first || !count)
    return fail(nullptr, ELX_E_INVALID, "elx_shard_range: bad argument");
  // getChunks (MarkovProcessAlongTree.hs:210-215): r chunks of q+1 sites, then n_shards - r chunks of q sites; inner
  // boundaries rounded up to multiples of `align`
  const int64_t q = nsites / n_shards, r = nsites % n_shards;
  auto bound = [&](int k) -> int64_t {
    if (k <= 0) return 0;
    if (k >= n_shards) return nsites;
    int64_t b = (int64_t)k * q + std::min<int64_t>(k, r);
    b = (b + align - 1) / align * align;
    return std::min(b, nsites);
  };
  *first = bound(shard);
  *count = bound(shard + 1) - bound(shard);
  return 0;
}

int elx_tree_walk_info(const elx_tree *tree, int32_t *n_leaves, int32_t *n_slots, int32_t *walk_nodes) {
  if (!tree || tree->n_nodes < 1 || !tree->parent || !tree->leaf_row) return fail(nullptr, ELX_E_INVALID, "tree: empty or NULL arrays");
  WalkProgram walk;
  int T = 0;
  std::string msg;
  if (!compile_walk(tree->n_nodes, tree->parent, tree->leaf_row, walk, T, msg)) return fail(nullptr, ELX_E_INVALID, msg);
  if (n_leaves) *n_leaves = T;
  if (n_slots) *n_slots = walk.n_slots;
  if (walk_nodes)
    for (int i = 0; i < tree->n_nodes; ++i) walk_nodes[i] = walk.recs[i].node;
  return 0;
}

int elx_tree_pair_walk_info(const elx_tree *tree, int32_t *n_groups, int32_t *n_slots, int32_t *recs) {
  if (!tree || tree->n_nodes < 1 || !tree->parent || !tree->leaf_row) return fail(nullptr, ELX_E_INVALID, "tree: empty or NULL arrays");
  WalkProgram walk;
  int T = 0;
  std::string msg;
  if (!compile_walk(tree->n_nodes, tree->parent, tree->leaf_row, walk, T, msg)) return fail(nullptr, ELX_E_INVALID, msg);
  PairProgram pp;
  compile_pair_walk(tree->n_nodes, tree->parent, tree->leaf_row, pp);
  if (n_groups) *n_groups = (int32_t)pp.recs.size();
  if (n_slots) *n_slots = pp.n_slots;
  if (recs)
    for (size_t g = 0; g < pp.recs.size(); ++g) {
      const PairRec &r = pp.recs[g];
      int32_t *o = recs + 7 * g;
      o[0] = r.nodeA; o[1] = r.nodeB; o[2] = r.src; o[3] = r.dstA; o[4] = r.dstB; o[5] = r.leafA; o[6] = r.leafB;
    }
  return 0;
}

int elx_tree_quad_walk_info(const elx_tree *tree, int32_t *n_records, int32_t *n_slots, int32_t *recs) {
  if (!tree || tree->n_nodes < 1 || !tree->parent || !tree->leaf_row) return fail(nullptr, ELX_E_INVALID, "tree: empty or NULL arrays");
  WalkProgram walk;
  int T = 0;
  std::string msg;
  if (!compile_walk(tree->n_nodes, tree->parent, tree->leaf_row, walk, T, msg)) return fail(nullptr, ELX_E_INVALID, msg);
  QuadProgram qp;
  compile_quad_walk(tree->n_nodes, tree->parent, tree->leaf_row, qp);
  if (n_records) *n_records = (int32_t)qp.recs.size();
  if (n_slots) *n_slots = qp.n_slots;
  if (recs)
    for (size_t g = 0; g < qp.recs.size(); ++g) quad_rec_to_ints(qp.recs[g], recs + g * ELX_QUAD_REC_INTS);
  return 0;
}

int elx_tree_duo_walk_info(const elx_tree *tree, int32_t *n_records, int32_t *n_slots, int32_t *recs) {
  if (!tree || tree->n_nodes < 1 || !tree->parent || !tree->leaf_row) return fail(nullptr, ELX_E_INVALID, "tree: empty or NULL arrays");
  WalkProgram walk;
  int T = 0;
  std::string msg;
  if (!compile_walk(tree->n_nodes, tree->parent, tree->leaf_row, walk, T, msg)) return fail(nullptr, ELX_E_INVALID, msg);
  QuadProgram qp;
  compile_quad_walk(tree->n_nodes, tree->parent, tree->leaf_row, qp, 2);
  if (n_records) *n_records = (int32_t)qp.recs.size();
  if (n_slots) *n_slots = qp.n_slots;
  if (recs)
    for (size_t g = 0; g < qp.recs.size(); ++g) quad_rec_to_ints(qp.recs[g], recs + g * ELX_QUAD_REC_INTS);
  return 0;
}

void elx_destroy(elx_handle *h) {
  if (!h) return;
  for (elx_handle *p : h->peers) elx_destroy(p);
  h->peers.clear();
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  fast_tables_free(h);
  pair_free(h);
  quad_free(h);
  duo_free(h);
  cudaFree(h->d_Q); cudaFree(h->d_brlen); cudaFree(h->d_lambda); cudaFree(h->d_A); cudaFree(h->d_B); cudaFree(h->d_P); cudaFree(h->d_cum);
  cudaFree(h->d_e16); cudaFree(h->d_qlo); cudaFree(h->d_cumw); cudaFree(h->d_cumpi); cudaFree(h->d_walk);
  cudaFree(h->d_error_flag); cudaFree(h->d_names); cudaFree(h->d_name_off); cudaFree(h->d_seq_off);
  cudaFree(h->d_lut);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int32_t elx_n_leaves(const elx_handle *h) { return h ? h->T : 0; }
int32_t elx_n_nodes(const elx_handle *h) { return h ? h->N : 0; }
int32_t elx_n_states(const elx_handle *h) { return h ? h->K : 0; }
int32_t elx_n_components(const elx_handle *h) { return h ? h->C : 0; }
int32_t elx_alias_columns(const elx_handle *h) { return h ? h->KP : 0; }
void *elx_stream(const elx_handle *h) { return h ? (void *)h->stream : nullptr; }
int64_t elx_launch_count(const elx_handle *h) {
  if (!h) return 0;
  int64_t n = h->launches;
  for (const elx_handle *p : h->peers) n += p->launches;
  return n;
}
int64_t elx_d2h_bytes(const elx_handle *h) {
  if (!h) return 0;
  int64_t n = h->d2h_bytes;
  for (const elx_handle *p : h->peers) n += p->d2h_bytes;
  return n;
}
const char *elx_last_kernel(const elx_handle *h) { return h ? h->last_kernel.c_str() : ""; }

int elx_ptables_get(elx_handle *h, double *out, int cumulative) {
  if (!h || !out) return fail(h, ELX_E_INVALID, "elx_ptables_get: NULL argument");
  CUDA_TRY(h, cudaSetDevice(h->device));
  size_t n = (size_t)h->N * h->C * h->K * h->K;
  CUDA_TRY(h, cudaMemcpyAsync(out, cumulative ? h->d_cum : h->d_P, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int elx_ptables_set(elx_handle *h, const double *p) {
  if (!h || !p) return fail(h, ELX_E_INVALID, "elx_ptables_set: NULL argument");
  CUDA_TRY(h, cudaSetDevice(h->device));
  size_t n = (size_t)h->N * h->C * h->K * h->K;
  CUDA_TRY(h, cudaMemcpyAsync(h->d_P, p, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  return build_tables_from_p(h, true);
}

int elx_alias_get(elx_handle *h, uint16_t *e16, uint32_t *qlo) {
  if (!h || !e16 || !qlo) return fail(h, ELX_E_INVALID, "elx_alias_get: NULL argument");
  CUDA_TRY(h, cudaSetDevice(h->device));
  size_t n = (size_t)h->N * h->C * h->K * h->KP;
  CUDA_TRY(h, cudaMemcpyAsync(e16, h->d_e16, n * sizeof(uint16_t), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(qlo, h->d_qlo, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int32_t elx_pair_groups(const elx_handle *h) { return h ? pair_groups(h) : 0; }

int elx_pair_get(elx_handle *h, int32_t *node_a, int32_t *node_b, uint16_t *pe16, uint32_t *pqlo) {
  if (!h) return fail(nullptr, ELX_E_INVALID, "elx_pair_get: null handle");
  CUDA_TRY(h, cudaSetDevice(h->device));
  const int rc = ensure_pair_tables(h);
  if (rc) return rc;
  return pair_get(h, node_a, node_b, pe16, pqlo);
}

int32_t elx_quad_groups(const elx_handle *h) { return h ? quad_groups(h) : 0; }

int elx_quad_get(elx_handle *h, int32_t *recs, uint32_t *qe32, uint32_t *qres32) {
  if (!h) return fail(nullptr, ELX_E_INVALID, "elx_quad_get: null handle");
  CUDA_TRY(h, cudaSetDevice(h->device));
  return quad_get(h, recs, qe32, qres32);
}

int32_t elx_duo_groups(const elx_handle *h) { return h ? duo_groups(h) : 0; }

int elx_duo_get(elx_handle *h, int32_t *recs, uint32_t *de32, uint32_t *dres32) {
  if (!h) return fail(nullptr, ELX_E_INVALID, "elx_duo_get: null handle");
  CUDA_TRY(h, cudaSetDevice(h->device));
  return duo_get(h, recs, de32, dres32);
}

int elx_simulate_device(elx_handle *h, uint64_t seed, int64_t site0, int64_t nsites, uint8_t *d_leaves, int64_t leaves_pitch,
                        int32_t *d_comps, uint8_t *d_internal, int64_t internal_pitch) {
  int rc = check_sim_args(h, site0, nsites, d_leaves, leaves_pitch, d_internal, internal_pitch);
  if (rc) return rc;
  if (nsites == 0) return 0;
  CUDA_TRY(h, cudaSetDevice(h->device));
  SimArgs a = base_args(h);
  a.leaves = d_leaves; a.leaves_pitch = leaves_pitch; a.comps = d_comps; a.internal = d_internal;
  a.internal_pitch = internal_pitch; a.site0 = site0; a.nsites = nsites; a.seed = seed;
  if (h->quad) {
    int used = 0;
    rc = quad_simulate(h, a, &used);
    if (rc) return rc;
    if (used) return 0;
  }
  if (h->duo) {
    int used = 0;
    rc = duo_simulate(h, a, &used);
    if (rc) return rc;
    if (used) return 0;
  }
  if (h->pair) {
    int used = 0;
    rc = ensure_pair_tables(h);
    if (rc) return rc;
    rc = pair_simulate(h, a, &used);
    if (rc) return rc;
    if (used) return 0;
  }
  if (!(h->flags & ELX_FLAG_FORCE_GENERIC)) {
    int used = 0;
    rc = ensure_fast_tables(h);
    if (rc) return rc;
    rc = fast_simulate(h, a, &used);
    if (rc) return rc;
    if (used) return 0;
  }
  int64_t g0 = site0 >> 3, g1 = (site0 + nsites + 7) >> 3;
  int64_t groups = g1 - g0;
  const int threads = 128;
  unsigned blocks = (unsigned)((groups + threads - 1) / threads);
  if (h->rounds == 7) k2_freerun_generic<7><<<blocks, threads, 0, h->stream>>>(a);
  else k2_freerun_generic<10><<<blocks, threads, 0, h->stream>>>(a);
  h->launches++;
  h->last_kernel = "k2_freerun_generic<" + std::to_string(h->rounds) + ">";
  CUDA_TRY(h, cudaGetLastError());
  return 0;
}

int elx_site_components(elx_handle *h, uint64_t seed, int64_t site0, int64_t nsites, int32_t *comps) {
  if (!h) return fail(nullptr, ELX_E_INVALID, "null handle");
  if (nsites < 0 || site0 < 0 || (!comps && nsites > 0)) return fail(h, ELX_E_INVALID, "site_components: bad argument");
  if (nsites == 0) return 0;
  CUDA_TRY(h, cudaSetDevice(h->device));
  int32_t *d = nullptr;
  CUDA_TRY(h, cudaMalloc((void **)&d, sizeof(int32_t) * (size_t)nsites));
  SimArgs a = base_args(h);
  a.comps = d; a.site0 = site0; a.nsites = nsites; a.seed = seed;
  const unsigned blocks = (unsigned)((nsites + 127) / 128);
  if (h->rounds == 7) k_site_components<7><<<blocks, 128, 0, h->stream>>>(a);
  else k_site_components<10><<<blocks, 128, 0, h->stream>>>(a);
  h->launches++;
  cudaError_t e = cudaMemcpyAsync(comps, d, sizeof(int32_t) * (size_t)nsites, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d);
  if (e != cudaSuccess) return fail(h, ELX_E_CUDA, std::string("site_components: ") + cudaGetErrorString(e));
  return check_error_flag(h);
}

int elx_synchronize(elx_handle *h) {
  if (!h) return fail(nullptr, ELX_E_INVALID, "null handle");
  CUDA_TRY(h, cudaSetDevice(h->device));
  return check_error_flag(h);
}

int elx_simulate_injected_device(elx_handle *h, const double *d_U, int64_t u_pitch, int64_t nsites, uint8_t *d_leaves,
                                 int64_t leaves_pitch, int32_t *d_comps, uint8_t *d_internal, int64_t internal_pitch) {
  int rc = check_sim_args(h, 0, nsites, d_leaves, leaves_pitch, d_internal, internal_pitch);
  if (rc) return rc;
  if (!d_U && nsites > 0) return fail(h, ELX_E_INVALID, "simulate_injected: U is NULL");
  if (u_pitch < nsites) return fail(h, ELX_E_INVALID, "simulate_injected: u_pitch < nsites");
  if (nsites == 0) return 0;
  CUDA_TRY(h, cudaSetDevice(h->device));
  SimArgs a = base_args(h);
  a.leaves = d_leaves; a.leaves_pitch = leaves_pitch; a.comps = d_comps; a.internal = d_internal;
  a.internal_pitch = internal_pitch; a.site0 = 0; a.nsites = nsites; a.U = d_U; a.u_pitch = u_pitch;
  const int threads = 128;
  unsigned blocks = (unsigned)((nsites + threads - 1) / threads);
  k2_injected<<<blocks, threads, 0, h->stream>>>(a);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return 0;
}

// ---- 2-bit transport for the host-buffer calls (K <= 4) ---------------------------------------------------------
// The host-buffer result path is PCIe-bound (1 byte per cell).  States of a <= 4-letter alphabet need 2 bits: the batch is
// packed on the device (k_pack2), crosses PCIe at a quarter of the size into pinned staging buffers owned by the library,
// and host threads expand it into the caller's buffer (elx_io.cpp:unpack2_rows) while the next chunks are in flight.
__global__ void k_pack2(const uint8_t *__restrict__ src, int64_t spitch, uint8_t *__restrict__ dst, int64_t dpitch, int64_t n16,
                        int64_t total) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / n16, g = idx - r * n16;
    const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(src + r * spitch + g * 16));
    auto p = [](uint32_t w) { w &= 0x03030303u; return (w | (w >> 6) | (w >> 12) | (w >> 18)) & 0xffu; };
    *reinterpret_cast<uint32_t *>(dst + r * dpitch + g * 4) = p(v.x) | (p(v.y) << 8) | (p(v.z) << 16) | (p(v.w) << 24);
  }
}

// 5-bit form for alphabets of 5..32 states: 32 sites (two 128-bit loads) -> 20 bytes (five 32-bit stores) per thread;
// site i of a row sits in bits 5i..5i+4 of the row's little-endian bit stream.
__global__ void k_pack5(const uint8_t *__restrict__ src, int64_t spitch, uint8_t *__restrict__ dst, int64_t dpitch, int64_t n32,
                        int64_t total) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / n32, g = idx - r * n32;
    const uint4 *p = reinterpret_cast<const uint4 *>(src + r * spitch + g * 32);
    const uint4 a = __ldcs(p), b = __ldcs(p + 1);
    auto p4 = [](uint32_t w) {  // 4 sites -> 20 bits
      w &= 0x1f1f1f1fu;
      return (w & 0x1fu) | ((w >> 3) & 0x3e0u) | ((w >> 6) & 0x7c00u) | ((w >> 9) & 0xf8000u);
    };
    const uint64_t q0 = (uint64_t)p4(a.x) | ((uint64_t)p4(a.y) << 20), q1 = (uint64_t)p4(a.z) | ((uint64_t)p4(a.w) << 20),
                   q2 = (uint64_t)p4(b.x) | ((uint64_t)p4(b.y) << 20), q3 = (uint64_t)p4(b.z) | ((uint64_t)p4(b.w) << 20);
    uint32_t *o = reinterpret_cast<uint32_t *>(dst + r * dpitch + g * 20);
    o[0] = (uint32_t)q0;
    o[1] = (uint32_t)(q0 >> 32) | (uint32_t)(q1 << 8);
    o[2] = (uint32_t)(q1 >> 24) | (uint32_t)(q2 << 16);
    o[3] = (uint32_t)(q2 >> 16) | (uint32_t)(q3 << 24);
    o[4] = (uint32_t)(q3 >> 8);
  }
}

namespace {
constexpr size_t PIN_BYTES = (size_t)128 << 20;  // one staging set, cut into stages (ELX_STAGE_MB, default 8 MB each)
std::mutex g_pin_mu;
std::vector<uint8_t *> g_pin_free;  // staging sets of PIN_BYTES, kept across handles (pinning is slow)

uint8_t *pin_acquire() {
  {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    if (!g_pin_free.empty()) {
      uint8_t *p = g_pin_free.back();
      g_pin_free.pop_back();
      return p;
    }
  }
  uint8_t *p = nullptr;
  if (cudaHostAlloc((void **)&p, PIN_BYTES + 64, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void pin_release(uint8_t *p) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> lk(g_pin_mu);
    if (g_pin_free.size() < 8) {  // one set per device of a multi-device call
      g_pin_free.push_back(p);
      return;
    }
  }
  cudaFreeHost(p);
}

struct PackedChunk {
  int stage = 0;
  const uint8_t *src = nullptr;  // pinned staging rows
  uint8_t *dst = nullptr;        // caller's buffer + first site of the batch; row r of the chunk starts at dst + row_off[r]
  const int64_t *row_off = nullptr;
  const uint8_t *lut = nullptr;
  int raw = 0;                   // 1: the caller takes the packed form itself (dst is a packed matrix: bytes are copied)
  int bits = 2;
  int64_t rows = 0, ppitch = 0, nb = 0, ncol = 0, units = 0;
  std::atomic<int> issued{0};  // its copy and event are enqueued
  std::atomic<int> ready{0};   // the chunk sits in its staging buffer (set by the waiter thread)
  std::atomic<int64_t> next{0}, done{0};
};
struct TransportSync {
  std::mutex mu;
  std::condition_variable cv;       // workers: a chunk has arrived (or abort)
  std::condition_variable cv_done;  // the dispatcher: a chunk is unpacked, its stage is free
  std::condition_variable cv_issue; // the waiter thread: the next chunk's copy is enqueued
  std::atomic<int> abort_flag{0};
};
static int64_t env_i64(const char *name, int64_t dflt) {
  const char *e = getenv(name);
  return e ? atoll(e) : dflt;
}
// Sites per work unit of the unpack threads: a unit is one row's piece of a batch, whole (2 MB of the caller's buffer for the
// default batch of 2^21 sites).  Round 1 cut the pieces into 64 KB units handed out in row order; on a FRESH result buffer the
// thirteen threads then all faulted on the same huge page at once -- each of them allocates and zeroes a 2 MB page and all but
// one throw theirs away (cfg2, 10 GB into a fresh mapping: 312 ms; units handed out row-interleaved 270 ms; whole pieces of
// 2^20 sites 226 ms; of 2^21 sites 217 ms; a reused buffer is indifferent: 95-98 ms either way).  Streaming stores stay ahead
// of plain ones on fresh pages too (226 against 257 ms).  ELX_UNPACK_COLS / ELX_UNPACK_ORDER=1 / ELX_HOST_BATCH_LOG2 for A/B runs.
static const int64_t UNPACK_COLS = std::max<int64_t>(4096, env_i64("ELX_UNPACK_COLS", (int64_t)1 << 21)) & ~(int64_t)63;
static const int UNPACK_ORDER = (int)env_i64("ELX_UNPACK_ORDER", 0);
}  // namespace

// Unpack threads of a host-buffer call: all cores but one (the calling thread that feeds the streams and the waiter sleep
// most of the time; round 1 kept three cores free, which measures 5 % slower into a fresh buffer and the same into a reused one); ELX_HOST_THREADS when several processes share a box.
static thread_local int g_thread_budget = 0;  // > 0: unpack threads of the host-buffer call running on this thread (multi-device calls divide the cores)
static int transport_threads() {
  if (g_thread_budget > 0) return g_thread_budget;
  int nt = (int)std::thread::hardware_concurrency() - 1;  // (round 2: 15 of 16 against 13: fresh buffer 218 -> 207 ms, reused buffer unchanged)
  if (const char *e = getenv("ELX_HOST_THREADS")) {
    nt = atoi(e);
  } else if (const char *w = getenv("LOCAL_WORLD_SIZE")) {  // torchrun: the ranks of this box share its cores
    const int ranks = atoi(w);
    if (ranks > 1) nt /= ranks;
  }
  return std::max(1, std::min(nt, 32));
}

// Packed or plain?  ELX_PACKED_D2H forces either.  Otherwise: from 8 MB of result on (below, the thread start-up is not
// worth it); the 2-bit form always (ahead even with 3 threads per GPU on an 8-GPU box: 1.00e11 against 8.7e10 cells/s);
// the 5-bit form only with >= 8 unpack threads -- it moves 5/8 of the bytes over PCIe but the host then reads them and
// stores the result, 2.25 bytes of host memory traffic per cell against 1 for the DMA alone, and the GPUs of a box share
// the host's memory system (8 GPUs, 3 threads each, cfg3: 358 ms packed against ~280 ms plain; 1 GPU, 13 threads: 146
// against 192 ms).
static bool want_packed_transport(const elx_handle *h, int64_t cells) {
  if (h->K > 32) return false;
  if (const char *e = getenv("ELX_PACKED_D2H")) return atoi(e) != 0;
  if (cells < ((int64_t)8 << 20)) return false;
  return h->K <= 4 || transport_threads() >= 8;
}

// leaves / internal: base pointers; row r of a matrix starts at base + off[r].  lut: nullptr = state codes, else the
// alphabet (FASTA sequence lines; 4 characters for the 2-bit form, 32 for the 5-bit form).
static int simulate_host_packed(elx_handle *h, uint64_t seed, int64_t site0, int64_t nsites, uint8_t *leaves,
                                const int64_t *leaves_off, int32_t *comps, uint8_t *internal, const int64_t *internal_off,
                                const uint8_t *lut, bool raw = false) {
  CUDA_TRY(h, cudaSetDevice(h->device));
  const int bits = h->K <= 4 ? 2 : 5;  // a packing group is 32 sites: 8 or 20 bytes
  auto packed_pitch = [bits](int64_t n) { return (n + 31) / 32 * (bits == 2 ? 8 : 20); };
  // Duo handles bucket every device call by component inside superblocks (elx_duo.cu): their batches are whole superblocks
  // (a batch that starts inside one has to draw the components of the sites in front of it again), the first one shorter
  // when site0 is not aligned.
  // 2^21 sites per batch (a row's piece of a batch is then a whole huge page of the caller's buffer), fewer for very tall
  // matrices: the two device staging sets stay under 16 GB
  int64_t batch_log2 = std::min<int64_t>(24, std::max<int64_t>(16, env_i64("ELX_HOST_BATCH_LOG2", 21)));
  while (batch_log2 > 16 && (((int64_t)2 * (h->T + (internal ? h->N : 0))) << batch_log2) > ((int64_t)16 << 30)) --batch_log2;
  const int64_t BATCH = duo_batch_sites(h) > 0 ? duo_batch_sites(h) : (int64_t)1 << batch_log2;
  int64_t boff = duo_batch_sites(h) > 0 ? site0 % BATCH : 0;
  if (raw && (boff & 31)) boff = 0;  // packed rows of the batches must join on byte boundaries (32-site packing groups)
  const int NB = 2;
  const int64_t nbatch = (nsites + boff + BATCH - 1) / BATCH;
  const int64_t bmax = nsites < BATCH ? nsites : BATCH;
  auto batch_begin = [&](int64_t b) { return std::max<int64_t>(0, b * BATCH - boff); };
  auto batch_len = [&](int64_t b) { return std::min<int64_t>(nsites, (b + 1) * BATCH - boff) - batch_begin(b); };
  const int64_t bpitch = (bmax + 31) & ~(int64_t)31, ppitch = packed_pitch(bmax);  // ppitch: of a full batch; a ragged last batch packs tighter
  size_t stage_mb = 8;
  if (const char *e = getenv("ELX_STAGE_MB")) stage_mb = (size_t)std::min(128, std::max(1, atoi(e)));
  const size_t PIN_STAGE_BYTES = std::max<size_t>(stage_mb << 20, (size_t)ppitch);  // a chunk is at least one packed row
  const int PIN_STAGES = (int)(PIN_BYTES / PIN_STAGE_BYTES);
  if (PIN_STAGES < 2) return fail(h, ELX_E_INVALID, "packed transport: a packed row does not fit the staging buffers");
  const int64_t rows_per_chunk = std::max<int64_t>(1, (int64_t)PIN_STAGE_BYTES / ppitch);
  const int nbuf = nbatch > 1 ? NB : 1;
  uint8_t *d_leaves[NB] = {nullptr, nullptr}, *d_internal[NB] = {nullptr, nullptr}, *d_pl[NB] = {nullptr, nullptr},
          *d_pi[NB] = {nullptr, nullptr};
  int32_t *d_comps[NB] = {nullptr, nullptr};
  cudaEvent_t sim_done[NB] = {nullptr, nullptr}, copy_done[NB] = {nullptr, nullptr};
  cudaStream_t copy_stream = nullptr;
  uint8_t *pin = nullptr;
  TransportSync sync;
  std::vector<std::thread> workers;
  // every chunk of the call is known up front: (batch, matrix, row range)
  struct Plan { int64_t batch; int which; int64_t r0, r1; };
  std::vector<Plan> plan;
  for (int64_t b = 0; b < nbatch; ++b)
    for (int which = 0; which < (internal ? 2 : 1); ++which) {
      const int64_t R = which ? h->N : h->T;
      for (int64_t r0 = 0; r0 < R; r0 += rows_per_chunk) plan.push_back({b, which, r0, std::min(R, r0 + rows_per_chunk)});
    }
  std::vector<PackedChunk> chunks(plan.size());
  std::vector<cudaEvent_t> stage_ev;
  std::thread waiter;
  auto cleanup = [&]() {
    {
      std::lock_guard<std::mutex> lk(sync.mu);
      sync.abort_flag.store(1);
    }
    sync.cv.notify_all();
    sync.cv_issue.notify_all();
    if (waiter.joinable()) waiter.join();
    for (auto &t : workers) t.join();
    workers.clear();
    for (cudaEvent_t e : stage_ev) cudaEventDestroy(e);
    stage_ev.clear();
    if (copy_stream) cudaStreamSynchronize(copy_stream);
    cudaStreamSynchronize(h->stream);
    for (int i = 0; i < NB; ++i) {
      cudaFree(d_leaves[i]); cudaFree(d_internal[i]); cudaFree(d_comps[i]); cudaFree(d_pl[i]); cudaFree(d_pi[i]);
      if (sim_done[i]) cudaEventDestroy(sim_done[i]);
      if (copy_done[i]) cudaEventDestroy(copy_done[i]);
    }
    if (copy_stream) cudaStreamDestroy(copy_stream);
    pin_release(pin);
  };
#define CT(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return fail(h, ELX_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } } while (0)
  pin = pin_acquire();
  if (!pin) { cleanup(); return fail(h, ELX_E_CUDA, "cannot allocate the pinned staging buffers"); }
  CT(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < nbuf; ++i) {
    CT(cudaMalloc((void **)&d_leaves[i], (size_t)h->T * bpitch));
    CT(cudaMalloc((void **)&d_pl[i], (size_t)h->T * ppitch));
    if (internal) {
      CT(cudaMalloc((void **)&d_internal[i], (size_t)h->N * bpitch));
      CT(cudaMalloc((void **)&d_pi[i], (size_t)h->N * ppitch));
    }
    if (comps) CT(cudaMalloc((void **)&d_comps[i], sizeof(int32_t) * (size_t)bpitch));
    CT(cudaEventCreateWithFlags(&sim_done[i], cudaEventDisableTiming));
    CT(cudaEventCreateWithFlags(&copy_done[i], cudaEventDisableTiming));
  }

  // One waiter thread turns "copy of chunk c finished" (an event per stage) into a wake-up of the workers.  A host
  // callback on the copy stream did the same at ~0.25 ms per chunk: the stream stalls until the callback has returned.
  for (int i = 0; i < PIN_STAGES; ++i) {
    cudaEvent_t e = nullptr;
    CT(cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync));  // the waiter sleeps, the cores unpack
    stage_ev.push_back(e);
  }
  waiter = std::thread([&chunks, &sync, &stage_ev, dev = h->device]() {
    cudaSetDevice(dev);
    for (size_t c = 0; c < chunks.size(); ++c) {
      PackedChunk &ch = chunks[c];
      {
        std::unique_lock<std::mutex> lk(sync.mu);
        sync.cv_issue.wait(lk, [&]() { return ch.issued.load(std::memory_order_acquire) || sync.abort_flag.load(); });
      }
      if (sync.abort_flag.load()) return;
      if (cudaEventSynchronize(stage_ev[ch.stage]) != cudaSuccess) {
        {
          std::lock_guard<std::mutex> lk(sync.mu);
          sync.abort_flag.store(2);
        }
        sync.cv.notify_all();
        sync.cv_done.notify_all();
        return;
      }
      {
        std::lock_guard<std::mutex> lk(sync.mu);
        ch.ready.store(1, std::memory_order_release);
      }
      sync.cv.notify_all();
    }
  });
  // waits are blocking: spinning
  // workers were measured to starve the dispatcher (16 spinning threads on 16 cores: 160 ms for cfg2, 8: 122 ms)
  const int nt = transport_threads();
  const bool debug = getenv("ELX_TRANSPORT_DEBUG") != nullptr;  // prints where the call's time went
  const auto t_start = std::chrono::steady_clock::now();
  auto since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
  std::atomic<int64_t> wait_us{0}, first_ready_us{-1};
  for (int w = 0; w < nt; ++w)
    workers.emplace_back([&chunks, &sync, &wait_us, &first_ready_us, t_start, debug]() {
      for (size_t c = 0; c < chunks.size(); ++c) {
        PackedChunk &ch = chunks[c];
        if (!ch.ready.load(std::memory_order_acquire)) {
          const auto t0 = std::chrono::steady_clock::now();
          std::unique_lock<std::mutex> lk(sync.mu);
          sync.cv.wait(lk, [&]() { return ch.ready.load(std::memory_order_acquire) || sync.abort_flag.load(); });
          if (debug) {
            wait_us.fetch_add(std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count());
            int64_t expect = -1;
            if (c == 0) first_ready_us.compare_exchange_strong(expect, std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t_start).count());
          }
        }
        if (sync.abort_flag.load()) return;
        for (;;) {
          const int64_t u = ch.next.fetch_add(1, std::memory_order_relaxed);
          if (u >= ch.units) break;
          int64_t r, k;
          if (UNPACK_ORDER) { k = u / ch.rows; r = u - k * ch.rows; }
          else { r = u / ch.ncol; k = u - r * ch.ncol; }
          const int64_t c0 = k * UNPACK_COLS, c1 = std::min(ch.nb, c0 + UNPACK_COLS);
          if (ch.raw) {
            const int64_t p0 = ch.bits == 2 ? c0 / 4 : c0 / 8 * 5;
            const int64_t p1 = c1 == ch.nb ? (ch.nb + 31) / 32 * (ch.bits == 2 ? 8 : 20) : (ch.bits == 2 ? c1 / 4 : c1 / 8 * 5);
            memcpy(ch.dst + ch.row_off[r] + p0, ch.src + r * ch.ppitch + p0, (size_t)(p1 - p0));
          } else if (ch.bits == 2) elx::unpack2_rows(ch.src + r * ch.ppitch + c0 / 4, ch.ppitch, ch.dst + ch.row_off[r] + c0, 0, 0, 1, c1 - c0, ch.lut);
          else elx::unpack5_rows(ch.src + r * ch.ppitch + c0 / 8 * 5, ch.ppitch, ch.dst + ch.row_off[r] + c0, 0, 0, 1, c1 - c0, ch.lut);
          if (ch.done.fetch_add(1, std::memory_order_acq_rel) + 1 == ch.units) {  // the stage can be refilled
            { std::lock_guard<std::mutex> lk(sync.mu); }
            sync.cv_done.notify_one();
          }
        }
      }
    });

  auto issue_batch = [&](int64_t b) -> int {
    const int i = (int)(b % nbuf);
    const int64_t b0 = batch_begin(b), nb = batch_len(b);
    if (b >= nbuf) CT(cudaStreamWaitEvent(h->stream, copy_done[i], 0));  // staging set i has left the device
    int rc = elx_simulate_device(h, seed, site0 + b0, nb, d_leaves[i], bpitch, d_comps[i], d_internal[i], bpitch);
    if (rc) { cleanup(); return rc; }
    const int64_t pp = packed_pitch(nb);
    auto pack = [&](const uint8_t *src, uint8_t *dst, int64_t rows) {
      if (bits == 2) {
        const int64_t n16 = (nb + 31) / 32 * 2;
        k_pack2<<<(unsigned)std::min<int64_t>((rows * n16 + 255) / 256, 148 * 16), 256, 0, h->stream>>>(src, bpitch, dst, pp, n16, rows * n16);
      } else {
        const int64_t n32 = (nb + 31) / 32;
        k_pack5<<<(unsigned)std::min<int64_t>((rows * n32 + 255) / 256, 148 * 16), 256, 0, h->stream>>>(src, bpitch, dst, pp, n32, rows * n32);
      }
      h->launches++;
    };
    pack(d_leaves[i], d_pl[i], h->T);
    if (internal) pack(d_internal[i], d_pi[i], h->N);
    CT(cudaGetLastError());
    CT(cudaEventRecord(sim_done[i], h->stream));
    return 0;
  };

  int rc = issue_batch(0);
  if (rc) return rc;
  size_t c = 0;
  for (int64_t b = 0; b < nbatch; ++b) {
    const int i = (int)(b % nbuf);
    const int64_t b0 = batch_begin(b), nb = batch_len(b);
    if (b + 1 < nbatch && (rc = issue_batch(b + 1))) return rc;
    CT(cudaStreamWaitEvent(copy_stream, sim_done[i], 0));
    if (comps) CT(cudaMemcpyAsync(comps + b0, d_comps[i], sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, copy_stream));
    const int64_t pp = packed_pitch(nb);  // of this batch
    for (; c < plan.size() && plan[c].batch == b; ++c) {
      const Plan &pl = plan[c];
      PackedChunk &ch = chunks[c];
      ch.stage = (int)(c % PIN_STAGES);
      if (c >= (size_t)PIN_STAGES) {  // the stage is free once chunk c - PIN_STAGES is unpacked
        PackedChunk &old = chunks[c - PIN_STAGES];
        std::unique_lock<std::mutex> lk(sync.mu);
        sync.cv_done.wait(lk, [&]() { return old.done.load(std::memory_order_acquire) >= old.units || sync.abort_flag.load(); });
        if (sync.abort_flag.load()) {
          lk.unlock();
          cleanup();
          return fail(h, ELX_E_CUDA, "packed transport: waiting for a device-to-host copy failed");
        }
      }
      uint8_t *stage = pin + (size_t)ch.stage * PIN_STAGE_BYTES;
      const size_t bytes = (size_t)(pl.r1 - pl.r0) * pp;
      CT(cudaMemcpyAsync(stage, (pl.which ? d_pi[i] : d_pl[i]) + pl.r0 * pp, bytes, cudaMemcpyDeviceToHost, copy_stream));
      h->d2h_bytes += (int64_t)bytes;
      ch.src = stage;
      ch.raw = raw ? 1 : 0;
      ch.dst = (pl.which ? internal : leaves) + (raw ? (bits == 2 ? b0 / 4 : b0 / 8 * 5) : b0);
      ch.row_off = (pl.which ? internal_off : leaves_off) + pl.r0;
      ch.lut = lut;
      ch.bits = bits;
      ch.rows = pl.r1 - pl.r0; ch.ppitch = pp; ch.nb = nb;
      ch.ncol = (nb + UNPACK_COLS - 1) / UNPACK_COLS;
      ch.units = ch.rows * ch.ncol;
      CT(cudaEventRecord(stage_ev[ch.stage], copy_stream));
      {
        std::lock_guard<std::mutex> lk(sync.mu);
        ch.issued.store(1, std::memory_order_release);
      }
      sync.cv_issue.notify_one();
    }
    CT(cudaEventRecord(copy_done[i], copy_stream));
    if (comps) h->d2h_bytes += (int64_t)sizeof(int32_t) * nb;
  }
  const double t_issued = since(t_start);
  CT(cudaStreamSynchronize(copy_stream));
  const double t_copied = since(t_start);
  waiter.join();
  for (auto &t : workers) t.join();
  workers.clear();
  if (sync.abort_flag.load()) {  // the waiter gave up on a copy: parts of the caller's buffer were never written
    cleanup();
    return fail(h, ELX_E_CUDA, "packed transport: waiting for a device-to-host copy failed");
  }
  for (const PackedChunk &ch : chunks)
    if (ch.done.load(std::memory_order_acquire) < ch.units) {
      cleanup();
      return fail(h, ELX_E_CUDA, "packed transport: a chunk was not unpacked completely");
    }
  if (debug)
    fprintf(stderr, "[elx transport] %zu chunks, %d threads: first chunk ready %.1f ms, all issued %.1f ms, all copied %.1f ms, all unpacked %.1f ms; "
            "workers waited %.1f ms each on average\n", chunks.size(), nt, first_ready_us.load() / 1e3, t_issued, t_copied, since(t_start),
            wait_us.load() / 1e3 / nt);
#undef CT
  cleanup();
  return check_error_flag(h);
}

// host-buffer variants: device staging + copies, then synchronise
static int simulate_host_common(elx_handle *h, bool injected, uint64_t seed, int64_t site0, const double *U, int64_t u_pitch,
                                int64_t nsites, uint8_t *leaves, int64_t lp, int32_t *comps, uint8_t *internal, int64_t ip) {
  int rc = check_sim_args(h, site0, nsites, leaves, lp, internal, ip);
  if (rc) return rc;
  if (nsites == 0) return 0;
  {
    if (!injected && want_packed_transport(h, (int64_t)(h->T + (internal ? h->N : 0)) * nsites)) {
      std::vector<int64_t> lo(h->T), io(internal ? h->N : 0);
      for (int t = 0; t < h->T; ++t) lo[t] = t * lp;
      for (size_t n = 0; n < io.size(); ++n) io[n] = (int64_t)n * ip;
      return simulate_host_packed(h, seed, site0, nsites, leaves, lo.data(), comps, internal, io.data(), nullptr);
    }
  }
  CUDA_TRY(h, cudaSetDevice(h->device));
  // Site batches through two device staging sets: the kernels of batch b+1 (compute stream) overlap the device-to-host
  // copies of batch b (copy stream).  The result path is PCIe-bound: 1 byte per cell has to cross to the host.
  const int64_t BATCH = 1 << 20;
  const int NB = 2;
  uint8_t *d_leaves[NB] = {nullptr, nullptr}, *d_internal[NB] = {nullptr, nullptr};
  int32_t *d_comps[NB] = {nullptr, nullptr};
  double *d_U[NB] = {nullptr, nullptr};
  cudaEvent_t sim_done[NB] = {nullptr, nullptr}, copy_done[NB] = {nullptr, nullptr};
  cudaStream_t copy_stream = nullptr;
  int64_t bmax = nsites < BATCH ? nsites : BATCH;
  int64_t bpitch = (bmax + 15) & ~(int64_t)15;
  int urows = h->N + (h->is_mixture ? 2 : 1);
  const int nbuf = nsites > BATCH ? NB : 1;
  auto cleanup = [&]() {
    if (copy_stream) cudaStreamSynchronize(copy_stream);
    cudaStreamSynchronize(h->stream);
    for (int i = 0; i < NB; ++i) {
      cudaFree(d_leaves[i]); cudaFree(d_internal[i]); cudaFree(d_comps[i]); cudaFree(d_U[i]);
      if (sim_done[i]) cudaEventDestroy(sim_done[i]);
      if (copy_done[i]) cudaEventDestroy(copy_done[i]);
    }
    if (copy_stream) cudaStreamDestroy(copy_stream);
  };
#define CT(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return fail(h, ELX_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } } while (0)
  CT(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < nbuf; ++i) {
    CT(cudaMalloc((void **)&d_leaves[i], (size_t)h->T * bpitch));
    if (internal) CT(cudaMalloc((void **)&d_internal[i], (size_t)h->N * bpitch));
    if (comps) CT(cudaMalloc((void **)&d_comps[i], sizeof(int32_t) * (size_t)bpitch));
    if (injected) CT(cudaMalloc((void **)&d_U[i], sizeof(double) * (size_t)urows * bpitch));
    CT(cudaEventCreateWithFlags(&sim_done[i], cudaEventDisableTiming));
    CT(cudaEventCreateWithFlags(&copy_done[i], cudaEventDisableTiming));
  }
  int64_t batch = 0;
  for (int64_t b0 = 0; b0 < nsites; b0 += BATCH, ++batch) {
    const int i = (int)(batch % nbuf);
    int64_t nb = nsites - b0 < BATCH ? nsites - b0 : BATCH;
    if (batch >= nbuf) CT(cudaStreamWaitEvent(h->stream, copy_done[i], 0));  // staging set i is free again
    if (injected) {
      CT(cudaMemcpy2DAsync(d_U[i], sizeof(double) * bpitch, U + b0, sizeof(double) * u_pitch, sizeof(double) * nb, urows,
                           cudaMemcpyHostToDevice, h->stream));
      rc = elx_simulate_injected_device(h, d_U[i], bpitch, nb, d_leaves[i], bpitch, d_comps[i], d_internal[i], bpitch);
    } else {
      rc = elx_simulate_device(h, seed, site0 + b0, nb, d_leaves[i], bpitch, d_comps[i], d_internal[i], bpitch);
    }
    if (rc) { cleanup(); return rc; }
    CT(cudaEventRecord(sim_done[i], h->stream));
    CT(cudaStreamWaitEvent(copy_stream, sim_done[i], 0));
    CT(cudaMemcpy2DAsync(leaves + b0, lp, d_leaves[i], bpitch, nb, h->T, cudaMemcpyDeviceToHost, copy_stream));
    if (internal) CT(cudaMemcpy2DAsync(internal + b0, ip, d_internal[i], bpitch, nb, h->N, cudaMemcpyDeviceToHost, copy_stream));
    if (comps) CT(cudaMemcpyAsync(comps + b0, d_comps[i], sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, copy_stream));
    CT(cudaEventRecord(copy_done[i], copy_stream));
    h->d2h_bytes += nb * (int64_t)(h->T + (internal ? h->N : 0)) + (comps ? (int64_t)sizeof(int32_t) * nb : 0);
  }
  CT(cudaStreamSynchronize(copy_stream));
#undef CT
  cleanup();
  return check_error_flag(h);
}

// Multi-device handles: the site range is cut by the getChunks rule (elx_shard_range, boundaries on multiples of 64 sites),
// one host thread per device runs the single-device path of its peer handle straight into the caller's buffers (column
// offset = the shard's first site); the unpack threads are divided among the devices.  Sites are independent and every
// draw is keyed on the global site index, so the bytes equal the single-device result.
static int run_sharded(elx_handle *h, int64_t nsites, const std::function<int(elx_handle *, int64_t, int64_t)> &shard_call) {
  const int nd = 1 + (int)h->peers.size();
  const int budget = std::max(1, transport_threads() / nd);
  std::vector<int> rcs((size_t)nd, 0);
  std::vector<std::thread> th;
  for (int d = 0; d < nd; ++d) {
    int64_t f = 0, n = 0;
    elx_shard_range(nsites, nd, d, 64, &f, &n);
    if (n == 0) continue;
    elx_handle *hd = d == 0 ? h : h->peers[(size_t)d - 1];
    th.emplace_back([&rcs, &shard_call, d, hd, f, n, budget]() {
      g_thread_budget = budget;
      rcs[(size_t)d] = shard_call(hd, f, n);
    });
  }
  for (auto &t : th) t.join();
  for (int d = 0; d < nd; ++d)
    if (rcs[(size_t)d]) {
      elx_handle *hd = d == 0 ? h : h->peers[(size_t)d - 1];
      if (hd != h) return fail(h, rcs[(size_t)d], "device " + std::to_string(hd->device) + ": " + hd->last_error);
      return rcs[(size_t)d];
    }
  return 0;
}

// A caller's result buffer is usually fresh memory (mallocForeignPtrBytes, np.empty): its pages are faulted in by the
// threads that write the result, 4 KB at a time -- 0.5 s for 10 GB on a 16-core B200 host, three times the transport itself.
// With transparent huge pages in `madvise` mode the same first touch takes 0.16 s (tools/bench_prefault.py), so big
// result buffers are advised accordingly (a hint about FUTURE faults of the caller's pages, nothing is touched or moved;
// ELX_NO_MADVISE=1 turns it off).
static void advise_result_buffer(const void *p, int64_t bytes) {
#if defined(__linux__) && defined(MADV_HUGEPAGE)
  static const bool off = getenv("ELX_NO_MADVISE") != nullptr;
  if (off || !p || bytes < ((int64_t)64 << 20)) return;
  const uintptr_t a = (reinterpret_cast<uintptr_t>(p) + 0x1fffff) & ~(uintptr_t)0x1fffff;
  const uintptr_t e = (reinterpret_cast<uintptr_t>(p) + (uintptr_t)bytes) & ~(uintptr_t)0x1fffff;
  if (e > a) madvise(reinterpret_cast<void *>(a), e - a, MADV_HUGEPAGE);
#else
  (void)p; (void)bytes;
#endif
}

int32_t elx_packed_bits(const elx_handle *h) { return !h ? 0 : h->K <= 4 ? 2 : h->K <= 32 ? 5 : 8; }

int64_t elx_packed_pitch(const elx_handle *h, int64_t nsites) {
  if (!h || nsites < 0) return -1;
  return h->K <= 4 ? (nsites + 31) / 32 * 8 : h->K <= 32 ? (nsites + 31) / 32 * 20 : nsites;
}

int elx_simulate_packed(elx_handle *h, uint64_t seed, int64_t site0, int64_t nsites, uint8_t *packed, int64_t packed_pitch,
                        int32_t *comps) {
  if (!h) return fail(nullptr, ELX_E_INVALID, "null handle");
  if (nsites < 0 || site0 < 0) return fail(h, ELX_E_INVALID, "simulate_packed: negative site range");
  if (h->K > 32) return fail(h, ELX_E_UNSUPPORTED, "simulate_packed: more than 32 states have no packed form (use elx_simulate)");
  if (!packed && nsites > 0) return fail(h, ELX_E_INVALID, "simulate_packed: result buffer is NULL");
  if (packed_pitch < elx_packed_pitch(h, nsites)) return fail(h, ELX_E_INVALID, "simulate_packed: packed_pitch < elx_packed_pitch(nsites)");
  if (nsites == 0) return 0;
  advise_result_buffer(packed, (int64_t)h->T * packed_pitch);
  std::vector<int64_t> lo((size_t)h->T);
  for (int t = 0; t < h->T; ++t) lo[(size_t)t] = t * packed_pitch;
  const int bits = h->K <= 4 ? 2 : 5;
  if (!h->peers.empty() && nsites >= 4096)  // shard boundaries are multiples of 64 sites: whole packing groups
    return run_sharded(h, nsites, [&](elx_handle *hd, int64_t f, int64_t n) {
      return simulate_host_packed(hd, seed, site0 + f, n, packed + (bits == 2 ? f / 4 : f / 8 * 5), lo.data(), comps ? comps + f : nullptr,
                                  nullptr, nullptr, nullptr, true);
    });
  return simulate_host_packed(h, seed, site0, nsites, packed, lo.data(), comps, nullptr, nullptr, nullptr, true);
}

int elx_simulate(elx_handle *h, uint64_t seed, int64_t site0, int64_t nsites, uint8_t *leaves, int64_t leaves_pitch,
                 int32_t *comps, uint8_t *internal, int64_t internal_pitch) {
  if (h && leaves && nsites > 0) advise_result_buffer(leaves, (int64_t)h->T * leaves_pitch);
  if (h && internal && nsites > 0) advise_result_buffer(internal, (int64_t)h->N * internal_pitch);
  if (h && !h->peers.empty() && nsites >= 4096) {
    int rc = check_sim_args(h, site0, nsites, leaves, leaves_pitch, internal, internal_pitch);
    if (rc) return rc;
    return run_sharded(h, nsites, [&](elx_handle *hd, int64_t f, int64_t n) {
      return simulate_host_common(hd, false, seed, site0 + f, nullptr, 0, n, leaves + f, leaves_pitch, comps ? comps + f : nullptr,
                                  internal ? internal + f : nullptr, internal_pitch);
    });
  }
  return simulate_host_common(h, false, seed, site0, nullptr, 0, nsites, leaves, leaves_pitch, comps, internal, internal_pitch);
}

int elx_simulate_injected(elx_handle *h, const double *U, int64_t u_pitch, int64_t nsites, uint8_t *leaves, int64_t leaves_pitch,
                          int32_t *comps, uint8_t *internal, int64_t internal_pitch) {
  if (!U && nsites > 0) return fail(h, ELX_E_INVALID, "simulate_injected: U is NULL");
  if (u_pitch < nsites) return fail(h, ELX_E_INVALID, "simulate_injected: u_pitch < nsites");
  return simulate_host_common(h, true, 0, 0, U, u_pitch, nsites, leaves, leaves_pitch, comps, internal, internal_pitch);
}

// ---- FASTA ------------------------------------------------------------------------------------
static void leaf_names_in_row_order(const elx_handle *h, const char *const *names, std::vector<std::string> &out) {
  out.assign(h->T, std::string());
  for (int t = 0; t < h->T; ++t) out[t] = names && names[t] ? names[t] : "";
}

int64_t elx_fasta_size(const elx_handle *h, const char *const *names, int64_t total_sites) {
  if (!h || total_sites < 0) return -1;
  int64_t tot = 0;
  for (int t = 0; t < h->T; ++t) tot += 1 + (int64_t)(names && names[t] ? strlen(names[t]) : 0) + 1 + total_sites + 1;
  return tot;
}

static int fasta_upload_frame(elx_handle *h, const char *const *names, const char *alphabet, int64_t total_sites) {
  std::vector<std::string> nm;
  leaf_names_in_row_order(h, names, nm);
  std::vector<int64_t> name_off(h->T + 1, 0), seq_off(h->T, 0);
  std::string all;
  int64_t pos = 0;
  for (int t = 0; t < h->T; ++t) {
    name_off[t] = (int64_t)all.size();
    all += nm[t];
    seq_off[t] = pos + 1 + (int64_t)nm[t].size() + 1;
    pos = seq_off[t] + total_sites + 1;
  }
  name_off[h->T] = (int64_t)all.size();
  if (strlen(alphabet) < (size_t)h->K) return fail(h, ELX_E_INVALID, "fasta: alphabet shorter than the number of states");
  cudaFree(h->d_names); cudaFree(h->d_name_off); cudaFree(h->d_seq_off); cudaFree(h->d_lut);
  h->d_names = nullptr; h->d_name_off = nullptr; h->d_seq_off = nullptr; h->d_lut = nullptr;
  CUDA_TRY(h, cudaMalloc((void **)&h->d_names, all.size() + 1));
  CUDA_TRY(h, cudaMalloc((void **)&h->d_name_off, sizeof(int64_t) * (h->T + 1)));
  CUDA_TRY(h, cudaMalloc((void **)&h->d_seq_off, sizeof(int64_t) * h->T));
  CUDA_TRY(h, cudaMalloc((void **)&h->d_lut, 64));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_names, all.data(), all.size(), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_name_off, name_off.data(), sizeof(int64_t) * (h->T + 1), cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(h->d_seq_off, seq_off.data(), sizeof(int64_t) * h->T, cudaMemcpyHostToDevice, h->stream));
  uint8_t lut[64];
  memset(lut, '?', sizeof lut);
  memcpy(lut, alphabet, (size_t)h->K);
  CUDA_TRY(h, cudaMemcpyAsync(h->d_lut, lut, 64, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));  // the host vectors go out of scope
  h->fasta_total_sites = total_sites;
  h->fasta_names_key = all + "\x01" + alphabet;
  for (int t = 0; t < h->T; ++t) h->fasta_names_key += "\x02" + std::to_string(name_off[t]);
  return 0;
}

int elx_fasta_pack_device(elx_handle *h, const uint8_t *d_leaves, int64_t leaves_pitch, int64_t site0, int64_t nsites,
                          int64_t total_sites, const char *const *names, const char *alphabet, uint8_t *d_out) {
  if (!h || !d_out || !alphabet) return fail(h, ELX_E_INVALID, "fasta: NULL argument");
  if (site0 < 0 || nsites < 0 || site0 + nsites > total_sites) return fail(h, ELX_E_INVALID, "fasta: site range outside the alignment");
  if (nsites > 0 && (!d_leaves || leaves_pitch < nsites)) return fail(h, ELX_E_INVALID, "fasta: bad leaves buffer");
  CUDA_TRY(h, cudaSetDevice(h->device));
  // (re)upload the frame description when the names / alphabet / length changed
  std::string key;
  {
    std::vector<std::string> nm;
    leaf_names_in_row_order(h, names, nm);
    std::string all;
    std::vector<size_t> offs;
    for (auto &s : nm) { offs.push_back(all.size()); all += s; }
    key = all + "\x01" + alphabet;
    for (size_t o : offs) key += "\x02" + std::to_string(o);
  }
  if (!h->d_seq_off || h->fasta_total_sites != total_sites || key != h->fasta_names_key) {
    int rc = fasta_upload_frame(h, names, alphabet, total_sites);
    if (rc) return rc;
  }
  if (nsites > 0) {
    const int threads = 256;
    int64_t chunks = (nsites + 15) / 16 + 1;
    dim3 grid((unsigned)((chunks + threads - 1) / threads), (unsigned)h->T);
    k3_fasta_body<<<grid, threads, 0, h->stream>>>(d_leaves, leaves_pitch, site0, nsites, h->d_seq_off, h->d_lut, h->K, d_out);
    h->launches++;
  }
  int head = site0 == 0, tail = site0 + nsites == total_sites;
  if (head || tail) {
    k3_fasta_frame<<<(h->T + 127) / 128, 128, 0, h->stream>>>(h->T, h->d_seq_off, h->d_name_off, h->d_names, total_sites, head,
                                                                tail, d_out);
    h->launches++;
  }
  CUDA_TRY(h, cudaGetLastError());
  return 0;
}

int64_t elx_fasta_row_offset(const elx_handle *h, const char *const *names, int64_t total_sites, int32_t t) {
  if (!h || total_sites < 0 || t < 0 || t >= h->T) return -1;
  int64_t pos = 0;
  for (int i = 0; i <= t; ++i) {
    int64_t len = (int64_t)(names && names[i] ? strlen(names[i]) : 0);
    if (i == t) return pos + 1 + len + 1;
    pos += 1 + len + 1 + total_sites + 1;
  }
  return -1;
}

int elx_fasta_pack_slice_device(elx_handle *h, const uint8_t *d_leaves, int64_t leaves_pitch, int64_t nsites, const char *alphabet,
                                uint8_t *d_out, int64_t out_pitch) {
  if (!h || !d_out || !alphabet || !d_leaves) return fail(h, ELX_E_INVALID, "fasta slice: NULL argument");
  if (nsites < 0 || leaves_pitch < nsites || out_pitch < nsites) return fail(h, ELX_E_INVALID, "fasta slice: bad site count / pitch");
  if (strlen(alphabet) < (size_t)h->K) return fail(h, ELX_E_INVALID, "fasta: alphabet shorter than the number of states");
  if (nsites == 0) return 0;
  CUDA_TRY(h, cudaSetDevice(h->device));
  FastaLut lut;
  memset(lut.c, '?', sizeof lut.c);
  memcpy(lut.c, alphabet, (size_t)h->K);
  dim3 grid((unsigned)((nsites + 16 * 256 - 1) / (16 * 256)), (unsigned)h->T);
  k3_fasta_slice<<<grid, 256, 0, h->stream>>>(d_leaves, leaves_pitch, nsites, lut, h->K, d_out, out_pitch);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return 0;
}

int elx_simulate_fasta(elx_handle *h, uint64_t seed, int64_t total_sites, const char *const *names, const char *alphabet,
                       uint8_t *out, int64_t out_capacity, int32_t *comps) {
  if (!h || !out || !alphabet) return fail(h, ELX_E_INVALID, "simulate_fasta: NULL argument");
  int64_t need = elx_fasta_size(h, names, total_sites);
  if (need < 0 || out_capacity < need) return fail(h, ELX_E_INVALID, "simulate_fasta: output buffer too small");
  advise_result_buffer(out, need);
  {
    // sequence lines cross PCIe at 2 / 5 bits per character and are expanded to characters by host threads, the
    // `>name` lines are written here (same bytes as K3 writes on the device)
    if (total_sites > 0 && want_packed_transport(h, (int64_t)h->T * total_sites)) {
      if (strlen(alphabet) < (size_t)h->K) return fail(h, ELX_E_INVALID, "fasta: alphabet shorter than the number of states");
      std::vector<int64_t> seq_off(h->T);
      int64_t pos = 0;
      for (int t = 0; t < h->T; ++t) {
        const char *nm = names && names[t] ? names[t] : "";
        const size_t len = strlen(nm);
        out[pos] = '>';
        memcpy(out + pos + 1, nm, len);
        out[pos + 1 + len] = '\n';
        seq_off[t] = pos + 1 + (int64_t)len + 1;
        pos = seq_off[t] + total_sites + 1;
        out[pos - 1] = '\n';
      }
      uint8_t lut[32];
      memset(lut, '?', sizeof lut);
      memcpy(lut, alphabet, (size_t)h->K);
      if (!h->peers.empty() && total_sites >= 4096)  // every device fills its columns of every sequence line
        return run_sharded(h, total_sites, [&](elx_handle *hd, int64_t f, int64_t n) {
          return simulate_host_packed(hd, seed, f, n, out + f, seq_off.data(), comps ? comps + f : nullptr, nullptr, nullptr, lut);
        });
      return simulate_host_packed(h, seed, 0, total_sites, out, seq_off.data(), comps, nullptr, nullptr, lut);
    }
  }
  CUDA_TRY(h, cudaSetDevice(h->device));
  uint8_t *d_out = nullptr, *d_leaves = nullptr;
  int32_t *d_comps = nullptr;
  const int64_t BATCH = 1 << 22;
  int64_t bmax = total_sites < BATCH ? total_sites : BATCH;
  int64_t bpitch = ((bmax > 0 ? bmax : 1) + 15) & ~(int64_t)15;
  auto cleanup = [&]() { cudaFree(d_out); cudaFree(d_leaves); cudaFree(d_comps); };
#define CT(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); return fail(h, ELX_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } } while (0)
  CT(cudaMalloc((void **)&d_out, (size_t)need > 0 ? (size_t)need : 1));
  CT(cudaMalloc((void **)&d_leaves, (size_t)h->T * bpitch));
  if (comps) CT(cudaMalloc((void **)&d_comps, sizeof(int32_t) * (size_t)bpitch));
  int rc = 0;
  if (total_sites == 0) rc = elx_fasta_pack_device(h, nullptr, 0, 0, 0, 0, names, alphabet, d_out);
  for (int64_t b0 = 0; b0 < total_sites && rc == 0; b0 += BATCH) {
    int64_t nb = total_sites - b0 < BATCH ? total_sites - b0 : BATCH;
    rc = elx_simulate_device(h, seed, b0, nb, d_leaves, bpitch, d_comps, nullptr, 0);
    if (rc == 0) rc = elx_fasta_pack_device(h, d_leaves, bpitch, b0, nb, total_sites, names, alphabet, d_out);
    if (rc == 0 && comps) CT(cudaMemcpyAsync(comps + b0, d_comps, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, h->stream));
  }
  if (rc) { cleanup(); return rc; }
  CT(cudaMemcpyAsync(out, d_out, (size_t)need, cudaMemcpyDeviceToHost, h->stream));
  CT(cudaStreamSynchronize(h->stream));
#undef CT
  cleanup();
  return check_error_flag(h);
}

// writeGZFile for the alignment (elynx-tools InputOutput.hs:80-83 after Export/Fasta.hs:24-36), streamed: the alignment is
// simulated into device memory, then blocks of whole sequence lines (or column chunks of one very long line) are mapped to
// characters by K3 straight into one of two pinned host buffers; while the GPU fills the next buffer the host appends the
// finished block to an ordered multi-member gzip stream whose members are deflated by worker threads and written in file
// order.  D2H, deflate and the file write overlap; host memory stays bounded (2 x 64 MB + the members in flight).
int elx_simulate_fasta_gz(elx_handle *h, uint64_t seed, int64_t total_sites, const char *const *names, const char *alphabet,
                          const char *path, int32_t level, int32_t n_threads, int32_t *comps) {
  if (!h || !alphabet || !path) return fail(h, ELX_E_INVALID, "simulate_fasta_gz: NULL argument");
  if (total_sites < 0) return fail(h, ELX_E_INVALID, "simulate_fasta_gz: negative site count");
  if (strlen(alphabet) < (size_t)h->K) return fail(h, ELX_E_INVALID, "fasta: alphabet shorter than the number of states");
  CUDA_TRY(h, cudaSetDevice(h->device));
  const int T = h->T;
  const int64_t S = total_sites, pitch = ((S > 0 ? S : 1) + 127) & ~(int64_t)127;
  uint8_t *d_leaves = nullptr, *pin = nullptr;
  int32_t *d_comps = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  GzStream *gz = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_leaves); cudaFree(d_comps);
    if (pin) pin_release(pin);
    for (cudaEvent_t e : ev) if (e) cudaEventDestroy(e);
  };
#define CT(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); if (gz) { std::string x; gz_close(gz, x); } return fail(h, _e == cudaErrorMemoryAllocation ? ELX_E_NOMEM : ELX_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } } while (0)
  CT(cudaMalloc((void **)&d_leaves, (size_t)T * pitch));
  if (comps) CT(cudaMalloc((void **)&d_comps, sizeof(int32_t) * (size_t)(S > 0 ? S : 1)));
  const int64_t BATCH = 1 << 22;
  for (int64_t b0 = 0; b0 < S; b0 += BATCH) {
    const int64_t nb = S - b0 < BATCH ? S - b0 : BATCH;
    const int rc = elx_simulate_device(h, seed, b0, nb, d_leaves + b0, pitch, d_comps ? d_comps + b0 : nullptr, nullptr, 0);
    if (rc) { cleanup(); return rc; }
  }
  if (comps && S > 0) CT(cudaMemcpyAsync(comps, d_comps, sizeof(int32_t) * (size_t)S, cudaMemcpyDeviceToHost, h->stream));
  pin = pin_acquire();
  if (!pin) { cleanup(); return fail(h, ELX_E_NOMEM, "simulate_fasta_gz: cannot pin host staging memory"); }
  for (cudaEvent_t &e : ev) CT(cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync));
  std::string gerr;
  gz = gz_open(path, n_threads, level, 0, gerr);
  if (!gz) { cleanup(); return fail(h, ELX_E_INVALID, "simulate_fasta_gz: " + gerr); }
  const int64_t half = (int64_t)(PIN_BYTES / 2);
  // blocks in file order: (first row, number of rows, first column, number of columns)
  struct Blk { int r0, nr; int64_t c0, nc; };
  std::vector<Blk> blks;
  if (S <= half) {
    const int rows = (int)std::max<int64_t>(1, std::min<int64_t>(T, half / std::max<int64_t>(S, 1)));
    for (int r = 0; r < T; r += rows) blks.push_back({r, std::min(rows, T - r), 0, S});
  } else {
    for (int r = 0; r < T; ++r)
      for (int64_t c = 0; c < S; c += half) blks.push_back({r, 1, c, std::min(half, S - c)});
  }
  FastaLut lut;
  memset(lut.c, '?', sizeof lut.c);
  memcpy(lut.c, alphabet, (size_t)h->K);
  auto issue = [&](size_t i) -> cudaError_t {
    const Blk &b = blks[i];
    if (b.nc > 0) {
      dim3 grid((unsigned)((b.nc + 16 * 256 - 1) / (16 * 256)), (unsigned)b.nr);
      k3_fasta_slice<<<grid, 256, 0, h->stream>>>(d_leaves + (size_t)b.r0 * pitch + b.c0, pitch, b.nc, lut, h->K, pin + (i & 1) * half, b.nc);
      h->launches++;
      h->d2h_bytes += (int64_t)b.nr * b.nc;
    }
    return cudaEventRecord(ev[i & 1], h->stream);
  };
  bool ok = true;
  if (!blks.empty()) CT(issue(0));
  for (size_t i = 0; i < blks.size() && ok; ++i) {
    if (i + 1 < blks.size()) CT(issue(i + 1));  // the GPU fills the other buffer while this one is compressed
    CT(cudaEventSynchronize(ev[i & 1]));
    const Blk &b = blks[i];
    const uint8_t *src = pin + (i & 1) * half;
    for (int r = 0; r < b.nr && ok; ++r) {
      if (b.c0 == 0) {
        const char *nm = names && names[b.r0 + r] ? names[b.r0 + r] : "";
        ok = gz_append(gz, (const uint8_t *)">", 1) && gz_append(gz, (const uint8_t *)nm, strlen(nm)) && gz_append(gz, (const uint8_t *)"\n", 1);
      }
      ok = ok && gz_append(gz, src + (size_t)r * b.nc, (size_t)b.nc);
      if (b.c0 + b.nc == S) ok = ok && gz_append(gz, (const uint8_t *)"\n", 1);
    }
  }
#undef CT
  const bool closed = gz_close(gz, gerr);
  gz = nullptr;
  cudaStreamSynchronize(h->stream);
  cleanup();
  if (!ok || !closed) return fail(h, ELX_E_INVALID, "simulate_fasta_gz: " + (gerr.empty() ? std::string("write failed") : gerr));
  return check_error_flag(h);
}

int elx_histogram_device(elx_handle *h, const uint8_t *d_leaves, int64_t leaves_pitch, int64_t nsites, int64_t *d_state_counts,
                         int64_t *d_pattern_counts) {
  if (!h || !d_leaves) return fail(h, ELX_E_INVALID, "histogram: NULL argument");
  if (nsites < 0 || leaves_pitch < nsites) return fail(h, ELX_E_INVALID, "histogram: bad site count / pitch");
  if (nsites == 0) return 0;
  CUDA_TRY(h, cudaSetDevice(h->device));
  if (d_state_counts) {
    int64_t bx = (nsites + 256 * 128 - 1) / (256 * 128);
    if (bx > 1024) bx = 1024;
    dim3 grid((unsigned)bx, (unsigned)h->T);
    const size_t smem = sizeof(uint32_t) * 256 * (size_t)(h->K + 1);
    if (smem > 48 * 1024) CUDA_TRY(h, cudaFuncSetAttribute(k_state_counts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_state_counts<<<grid, 256, smem, h->stream>>>(d_leaves, leaves_pitch, nsites, h->K, (unsigned long long *)d_state_counts);
    h->launches++;
  }
  if (d_pattern_counts) {
    double bins = pow((double)h->K, (double)h->T);
    if (bins > (double)(1 << 24)) return fail(h, ELX_E_INVALID, "histogram: K^T exceeds 2^24 pattern bins");
    int64_t bx = (nsites + 255) / 256;
    if (bx > 148 * 16) bx = 148 * 16;
    k_pattern_counts<<<(unsigned)bx, 256, 0, h->stream>>>(d_leaves, leaves_pitch, nsites, h->T, h->K,
                                                           (unsigned long long *)d_pattern_counts);
    h->launches++;
  }
  CUDA_TRY(h, cudaGetLastError());
  return 0;
}

}  // extern "C"
