// Microbenchmark: how do LOP3 (ALU pipe) and IMAD.WIDE.U32 (FMA-heavy pipe) share an SM sub-partition on sm_100a?
// Every thread runs 8 independent dependency chains; one loop iteration issues A three-input LOP3s and W wide multiplies
// per chain set.  Prints warp-instructions per clock per SM for a few mixes.  nvcc -arch=sm_100a -O3 -o mb tools/microbench_pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int A, int W>
__global__ void __launch_bounds__(256) k(uint32_t *out, int iters, uint32_t k0, uint32_t k1) {
  uint32_t x[8], y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 2654435761u + i; y[i] = blockIdx.x + i * 40503u; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int w = 0; w < W; ++w) {
        uint32_t lo, hi;
        asm volatile("{.reg .u64 t; mul.wide.u32 t, %2, 0xD2511F53; mov.b64 {%0,%1}, t;}" : "=r"(lo), "=r"(hi) : "r"(x[i]));
        x[i] = hi; y[i] ^= lo;
      }
#pragma unroll
      for (int a = 0; a < A; ++a) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(a & 1 ? k0 : k1));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= x[i] ^ y[i];
  if (s == 0x12345678u) out[0] = s;
}

// MODE 0: mul.lo (IMAD), 1: mul.hi (IMAD.HI), 2: mul.lo + mul.hi as two instructions, 3: mad.lo with a register addend,
// 4: shf.l (ALU shift) for reference, 5: prmt
template <int A, int MODE>
__global__ void __launch_bounds__(256) k2(uint32_t *out, int iters, uint32_t k0, uint32_t k1) {
  uint32_t x[8], y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 2654435761u + i; y[i] = blockIdx.x + i * 40503u; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t lo = 0, hi = 0;
      if (MODE == 0) { asm volatile("mul.lo.u32 %0, %1, 0xD2511F53;" : "=r"(lo) : "r"(x[i])); x[i] = lo; }
      if (MODE == 1) { asm volatile("mul.hi.u32 %0, %1, 0xD2511F53;" : "=r"(hi) : "r"(x[i])); x[i] = hi; }
      if (MODE == 2) {
        asm volatile("mul.lo.u32 %0, %1, 0xD2511F53;" : "=r"(lo) : "r"(x[i]));
        asm volatile("mul.hi.u32 %0, %1, 0xD2511F53;" : "=r"(hi) : "r"(x[i]));
        x[i] = hi; y[i] ^= lo;
      }
      if (MODE == 3) { asm volatile("mad.lo.u32 %0, %1, 0xD2511F53, %2;" : "=r"(lo) : "r"(x[i]), "r"(y[i])); x[i] = lo; }
      if (MODE == 4) { asm volatile("shf.l.wrap.b32 %0, %1, %2, 7;" : "=r"(lo) : "r"(x[i]), "r"(y[i])); x[i] = lo; }
      if (MODE == 5) { asm volatile("prmt.b32 %0, %1, %2, 0x2143;" : "=r"(lo) : "r"(x[i]), "r"(y[i])); x[i] = lo; }
#pragma unroll
      for (int a = 0; a < A; ++a) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(a & 1 ? k0 : k1));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= x[i] ^ y[i];
  if (s == 0x12345678u) out[0] = s;
}

template <int A, int MODE>
static void run2(const char *name, int warps_per_sm) {
  uint32_t *d;
  cudaMalloc(&d, 4);
  const int iters = 2048, ctas = 148 * warps_per_sm / 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k2<A, MODE><<<ctas, 256>>>(d, 64, 1, 2);
  cudaEventRecord(e0);
  k2<A, MODE><<<ctas, 256>>>(d, iters, 1, 2);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double cycles = ms * 1e-3 * clk_khz * 1e3;
  printf("%-34s warps/SM %2d  cycles per (%d lop3 + op) per SMSP: %.2f\n", name, warps_per_sm, A + (MODE == 2 ? 1 : 0),
         cycles / ((double)iters * 8 * warps_per_sm / 4));
  cudaFree(d);
}

template <int A, int W>
static void run(const char *name, int warps_per_sm) {
  uint32_t *d;
  cudaMalloc(&d, 4);
  const int iters = 2048, ctas = 148 * warps_per_sm / 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<A, W><<<ctas, 256>>>(d, 64, 1, 2);
  cudaEventRecord(e0);
  k<A, W><<<ctas, 256>>>(d, iters, 1, 2);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double cycles = ms * 1e-3 * clk_khz * 1e3;
  const double winst = (double)iters * 8 * (A + W + (W ? 1.0 * W : 0)) * warps_per_sm;  // per SM; the y ^= lo is one more LOP per wide mul
  printf("%-28s warps/SM %2d  %.3f ms  per SMSP per clk: lop3 %.3f  imad.wide %.3f  (all %.3f)\n", name, warps_per_sm, ms,
         (double)iters * 8 * (A + W) * warps_per_sm / 4 / cycles, (double)iters * 8 * W * warps_per_sm / 4 / cycles, winst / 4 / cycles);
  cudaFree(d);
}

int main() {
  for (int wps : {8, 16, 32}) {
    run<4, 0>("lop3 only", wps);
    run<0, 2>("imad.wide (+1 lop each)", wps);
    run<2, 2>("2 lop3 + 2 wide (+2 lop)", wps);
    run<1, 1>("1 lop3 + 1 wide (+1 lop)", wps);
    run<4, 1>("4 lop3 + 1 wide (+1 lop)", wps);
    run<6, 1>("6 lop3 + 1 wide (+1 lop)", wps);
  }
  for (int wps : {16, 32}) {
    run2<4, 0>("4 lop3 + mul.lo", wps);
    run2<4, 1>("4 lop3 + mul.hi", wps);
    run2<4, 2>("5 lop3 + mul.lo + mul.hi", wps);
    run2<4, 3>("4 lop3 + mad.lo reg", wps);
    run2<4, 4>("4 lop3 + shf", wps);
    run2<4, 5>("4 lop3 + prmt", wps);
    run2<0, 0>("mul.lo only", wps);
    run2<0, 1>("mul.hi only", wps);
  }
  return 0;
}
