// Tiled shared-memory simulation kernels (K2, free-running mode) for sm_100a.
//
// Same draw specification as k2_freerun_generic (elx_core.cu) -- the two must agree bit for bit -- but laid out for
// the machine:
//   * one thread owns R x 16 sites (two Philox4x32 calls per node and round = 16 x 16-bit draws) and writes each leaf
//     row with one 128-bit streaming store per round (a warp covers 512 contiguous bytes of the [taxon][site] row);
//   * the per-branch alias tables and the walk records are streamed through shared memory in stages of `nps` nodes
//     by TMA bulk copies (cp.async.bulk + mbarrier, up to 3 stages in flight; one CTA barrier per stage retires a buffer);
//   * live internal-node states sit in a shared-memory slot stack [slot][round][thread] (128-bit, conflict free); the
//     walk order (heavy child last) bounds the number of slots by floor(log2 N)+1;
//   * a draw is: PRMT/LOP (parent row) + LOP3 (entry address, [R+UR]) + LDS.U16 + half a SHF/PRMT/VIMNMX.U16x2/LOP3 --
//     the alias decision is a packed 16-bit min: child bits = low bits of min(random, entry); ties in the threshold
//     bits (2^-14 DNA, 2^-11 protein) fall to an exact 32-bit refinement path.
//
// Layouts (template parameter LAY):
//   0  DNA      K = 4,  KP = 4,  C <= 8: per-site byte = comp*32 + state*8 = byte offset of the table row (4 packed regs)
//   1  protein  K <= 32, KP = 32, C*K*64 <= 65535: per-site 16-bit lane = (comp*K + state)*64 = row byte offset (8 regs)
//   2  big mix  K <= 32, KP = 32, C*K*32 <= 65535: per-site 16-bit lane = (comp*K + state)*32 = first entry index of the
//               row (bits 0..4 free for the column, so both draws of a Philox word get their index in one logic op);
//               the rows of ONE
//               branch fill a whole stage (e.g. C60: 76.8 KB), so a CTA is 512 threads (8192 sites) to amortise the
//               table stream, 2 stages, 1 CTA per SM.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>

#include <cuda_runtime.h>

#include "../../include/elynx_b200.h"
#include "elx_device.cuh"
#include "elx_ptx.cuh"
#include "elx_state.hpp"

namespace elx {

constexpr int MAX_NSTAGE = 3;  // TMA stages in flight

struct FastTables {
  int kind = 0;  // 0 = none (generic kernel), 1 + LAY otherwise
  uint8_t *d_ftab = nullptr;
  int tb = 0;           // table bytes per node
  int nps = 0;          // nodes per TMA stage
  int nstage = 0;       // stages in flight (2 or 3)
  int stage_bytes = 0;  // nps * (16 + tb)
  int n_stages = 0;
  int threads = 128, rounds = 1;
  size_t smem = 0;
};

struct FastArgs {
  SimArgs s;
  const uint8_t *ftab;
  int tb, stage_bytes, n_stages, nps, nstage;
  uint32_t two;  // = 2, as a launch constant: `w * two` is an IMAD on the FMA pipe, `w << 1` a shift on the busier ALU pipe
};

template <int LAY>
struct Lay {
  static constexpr bool AA = LAY != 0;
  static constexpr int NREG = AA ? 8 : 4;            // packed parent registers per thread and round (16 sites)
  static constexpr uint32_t JMASK = AA ? 31u : 3u;   // column bits of a 16-bit draw x = [r : 16-A][j : A]
  static constexpr uint32_t TIE = AA ? 32u : 4u;      // lane of (x ^ e) below this <=> threshold bits equal
  static constexpr int A = AA ? 5 : 2;
};

// PW: a dedicated producer warp (thread THREADS of a THREADS + 32 CTA) keeps the ring full and the consumer warps
// release a stage buffer one by one (`empty` mbarriers) -- no CTA barrier in the node loop, as in k2_quad_tiled.
// !PW: thread 0 refills a buffer after a CTA barrier per stage.
template <int ROUNDS, int LAY, int THREADS, int R, bool PW>
__global__ void __launch_bounds__(THREADS + (PW ? 32 : 0), THREADS == 128 ? 4 : 1) k2_tiled(FastArgs fa) {
  using L = Lay<LAY>;
  constexpr int NREG = L::NREG;
  constexpr bool AA = L::AA;
  extern __shared__ __align__(128) uint8_t smem[];
  const SimArgs &a = fa.s;
  const int tid = threadIdx.x;
  const int nstage = fa.nstage, nps = fa.nps;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem);  // [nstage] mbarriers: the stage's TMA bytes have landed
  const uint32_t stage_smem = smem_u32(smem + 128);
  const uint32_t slots_smem = stage_smem + (uint32_t)(nstage * fa.stage_bytes);
  constexpr uint32_t slot_stride = THREADS * R * 16u;  // slot stack [slot][round][thread] x 16 B

  uint64_t *empty_bar = full_bar + MAX_NSTAGE;              // [nstage] (PW): every consumer warp is done with the buffer
  if (tid == 0) {
    for (int i = 0; i < nstage; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      if (PW) mbar_init(smem_u32(&empty_bar[i]), THREADS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  __syncthreads();
  // (the warp index through a shuffle: ptxas then knows the branch is warp-uniform and keeps uniform-register
  // addressing in the consumer loop)
  if (PW && __shfl_sync(0xffffffffu, tid >> 5, 0) >= THREADS / 32) {
    if (tid == THREADS) {
      int pbuf = 0;
      uint32_t pphase = 1;  // the first pass over the ring finds every buffer free
      for (int st = 0; st < fa.n_stages; ++st) {
        if (st >= nstage) mbar_wait_backoff(smem_u32(&empty_bar[pbuf]), pphase, 128);
        mbar_expect_tx(smem_u32(&full_bar[pbuf]), (uint32_t)fa.stage_bytes);
        tma_bulk_g2s(stage_smem + (uint32_t)(pbuf * fa.stage_bytes), fa.ftab + (size_t)st * fa.stage_bytes, (uint32_t)fa.stage_bytes,
                     smem_u32(&full_bar[pbuf]));
        if (++pbuf == nstage) { pbuf = 0; pphase ^= 1u; }
      }
    }
    return;
  }
  // Fill the ring: stages 0..nstage-1.  Refills happen at the end of each stage, after the CTA barrier that
  // retires the buffer, in straight-line code.  (Measured: per-warp release with an in-band producer loop was no
  // faster, and any data-dependent loop under `tid == 0` makes ptxas drop uniform-register addressing for the whole
  // node loop, which costs one extra ALU-pipe IADD3 per draw.)
  if (!PW && tid == 0) {
    for (int i = 0; i < nstage && i < fa.n_stages; ++i) {
      mbar_expect_tx(smem_u32(&full_bar[i]), (uint32_t)fa.stage_bytes);
      tma_bulk_g2s(stage_smem + (uint32_t)(i * fa.stage_bytes), fa.ftab + (size_t)i * fa.stage_bytes, (uint32_t)fa.stage_bytes,
                   smem_u32(&full_bar[i]));
    }
  }

  const uint2 key = make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32));
  // per-round thread state (16 sites per round); plain arrays with static indices only
  uint32_t cbase[R][NREG];  // component part of the packed row values (loop invariant)
  uint32_t par[R][NREG];    // packed row values of the parent states (mirror of slot `reg_slot`)
  uint32_t slot_base[R];    // shared-memory address of this thread's 16 bytes in slot 0
  uint8_t *leaf_base[R];    // leaves + column of the first site
  int64_t sbase[R], col0[R];
  int store_mode[R];        // 0 = nothing to store (past the site range), 1 = 128-bit row stores, 2 = byte stores
#pragma unroll
  for (int r = 0; r < R; ++r) {
    // 16 sites per thread and round, aligned on global site index multiples of 16; within a round the lanes of a
    // warp own consecutive 16-site groups, so every row store of a warp is one contiguous 512-byte segment
    const int64_t g16 = (a.site0 >> 4) + ((int64_t)blockIdx.x * R + r) * THREADS + tid;
    sbase[r] = g16 << 4;
    col0[r] = sbase[r] - a.site0;
    const bool any = sbase[r] < a.site0 + a.nsites;
    const bool full = (col0[r] >= 0) && (col0[r] + 16 <= a.nsites);
#pragma unroll
    for (int q = 0; q < NREG; ++q) { cbase[r][q] = 0; par[r][q] = 0; }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const uint32_t cr = draw_comp_root<ROUNDS>(a.cumw, a.cumpi, a.C, a.K, a.N, a.is_mixture, a.error_flag,
                                                 (uint64_t)sbase[r] + i, key.x, key.y);
      const int c = (int)(cr & 0xffffu), root = (int)(cr >> 16);
      if (LAY == 2) {  // lanes hold the first ENTRY index of the row, row * 32 (< 65536: the table fits shared memory)
        cbase[r][i >> 1] |= (uint32_t)(c * a.K * 32) << (16 * (i & 1));
        par[r][i >> 1] |= (uint32_t)((c * a.K + root) * 32) << (16 * (i & 1));
      } else if (LAY == 1) {
        cbase[r][i >> 1] |= (uint32_t)(c * a.K * 64) << (16 * (i & 1));
        par[r][i >> 1] |= (uint32_t)((c * a.K + root) * 64) << (16 * (i & 1));
      } else {
        cbase[r][i >> 2] |= (uint32_t)(c * 32) << (8 * (i & 3));
        par[r][i >> 2] |= (uint32_t)(c * 32 + root * 8) << (8 * (i & 3));
      }
      int64_t col = col0[r] + i;
      if (a.comps && col >= 0 && col < a.nsites) a.comps[col] = c;
    }
    // Pin the loop invariants into single registers: without this ptxas keeps the per-site pieces of cbase[] apart and
    // re-ORs them (and re-derives the slot address and the store predicates) on every node.
#pragma unroll
    for (int q = 0; q < NREG; ++q) asm volatile("" : "+r"(cbase[r][q]));
    slot_base[r] = slots_smem + (uint32_t)(r * THREADS + tid) * 16u;
    asm volatile("" : "+r"(slot_base[r]));
    // slot 0 = root states.  DNA slots hold the packed offset bytes (= par[]); protein slots hold state bytes
    // (the leaf output format) and are expanded to 16-bit lanes on reload.
    if (AA) {
      uint32_t b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t lo = par[r][2 * q] - cbase[r][2 * q], hi = par[r][2 * q + 1] - cbase[r][2 * q + 1];  // state lanes
        if (LAY == 1) { lo >>= 6; hi >>= 6; }
        if (LAY == 2) { lo >>= 5; hi >>= 5; }
        b[q] = __byte_perm(lo, hi, 0x6420u);
      }
      sts_v4(slot_base[r], make_uint4(b[0], b[1], b[2], b[3]));
    } else {
      sts_v4(slot_base[r], make_uint4(par[r][0], par[r][1], par[r][2], par[r][3]));
    }
    store_mode[r] = !any ? 0
                         : (full && ((a.leaves_pitch & 15) == 0) && ((reinterpret_cast<uintptr_t>(a.leaves) & 15) == 0) &&
                            ((col0[r] & 15) == 0)) ? 1 : 2;
    asm volatile("" : "+r"(store_mode[r]));
    leaf_base[r] = a.leaves + col0[r];
    asm volatile("" : "+l"(leaf_base[r]));
  }
  int reg_slot = 0;  // slot whose contents are mirrored in par[] (same walk for every round)

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
      const uint32_t tbl = sbuf + (uint32_t)(nps * 16) + (uint32_t)k * (uint32_t)fa.tb;
      const bool reload = src != reg_slot;
#ifdef ELX_DEBUG_BOUNDS
      if (src < 0 || src >= a.n_slots || dst >= a.n_slots || node >= (uint32_t)a.N || k * fa.tb + fa.tb + nps * 16 > fa.stage_bytes)
        atomicExch(a.error_flag, -8);
#endif
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (reload) {
          const uint4 v = lds_v4(slot_base[r] + (uint32_t)src * slot_stride);
          if (AA) {
            const uint32_t b[4] = {v.x, v.y, v.z, v.w};
            constexpr uint32_t scale = LAY == 1 ? 64u : 32u;
#pragma unroll
            for (int q = 0; q < 4; ++q) {  // state bytes -> 16-bit lanes -> packed row values
              par[r][2 * q] = __byte_perm(b[q], 0u, 0x4140u) * scale + cbase[r][2 * q];
              par[r][2 * q + 1] = __byte_perm(b[q], 0u, 0x4342u) * scale + cbase[r][2 * q + 1];
            }
          } else {
            par[r][0] = v.x; par[r][1] = v.y; par[r][2] = v.z; par[r][3] = v.w;
          }
        }
        const uint64_t G0 = (uint64_t)sbase[r] >> 3;
        const uint32_t G0lo = (uint32_t)G0, Ghi = (uint32_t)(G0 >> 32), G1lo = G0lo | 1u;
        const uint4 x0 = philox4x32<ROUNDS>(make_uint4(G0lo, Ghi, node, TAG_NODE), key);
        const uint4 x1 = philox4x32<ROUNDS>(make_uint4(G1lo, Ghi, node, TAG_NODE), key);
        const uint32_t w[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        uint32_t c2[8], t[8];
#pragma unroll
        for (int pr = 0; pr < 8; ++pr) {  // pair pr = sites 2pr, 2pr+1 share Philox word pr
          uint32_t offA, offB;
          if (AA) {
            offA = par[r][pr] & 0xffffu;
            offB = par[r][pr] >> 16;
            if (LAY == 2) { offA <<= 1; offB <<= 1; }  // entry index -> byte offset (replaced below)
          } else {
            offA = __byte_perm(par[r][pr >> 1], 0u, 0x4440u | ((pr & 1) * 2));
            offB = __byte_perm(par[r][pr >> 1], 0u, 0x4440u | ((pr & 1) * 2 + 1));
          }
          // entry address = row offset | 2*j.  Written as an OR: with an ADD ptxas folds the (warp-uniform) table base
          // into a 3-input IADD3 per draw instead of using [R + UR] addressing.
          uint32_t aA = ((w[pr] & L::JMASK) << 1) | offA;
          uint32_t aB = (((w[pr] >> 16) & L::JMASK) << 1) | offB;
          if (LAY == 1) {
            // both entry addresses in one 3-input logic op on the packed 16-bit lanes: the row offsets are multiples of
            // 64 (bits 1..5 free), bit 15 of the low draw leaks into bit 0 of the high lane and is masked
            const uint32_t a2 = ((w[pr] * fa.two) & 0x003e003eu) | par[r][pr];
            aA = a2 & 0xffffu;
            aB = a2 >> 16;
          }
          if (LAY == 2) {
            // the lanes are entry indices with bits 0..4 free: index | j in one logic op for both draws, then two
            // multiply-adds (FMA pipe) turn them into byte addresses
            const uint32_t i2 = (w[pr] & 0x001f001fu) | par[r][pr];
            aA = (i2 & 0xffffu) * fa.two;
            aB = (i2 >> 16) * fa.two;
          }
#ifdef ELX_DEBUG_BOUNDS  // compute-sanitizer is closed on the GPU pool: self-check every shared-memory table access
          if (aA + 2 > (uint32_t)fa.tb || aB + 2 > (uint32_t)fa.tb || (aA & 1) || (aB & 1)) atomicExch(a.error_flag, -7);
#endif
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
                oo = AA ? par[r][q] : (__byte_perm(par[r][q >> 1], 0u, 0x4440u | ((q & 1) * 2)) |
                                       (__byte_perm(par[r][q >> 1], 0u, 0x4440u | ((q & 1) * 2 + 1)) << 16));
              }
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
              const uint32_t sh = (uint32_t)hh * 16u;
              if (((tt >> sh) & 0xffffu) < L::TIE) {
                const uint32_t x16 = (ww >> sh) & 0xffffu, lane = (oo >> sh) & 0xffffu, j = x16 & KPm1;
                // lane -> row byte offset and row index (= comp*K + state); e16/qlo index = (node*C*K + row)*KP + j
                const uint32_t off = LAY == 2 ? lane << 1 : lane;
                const uint32_t row = LAY == 2 ? lane >> 5 : (LAY == 1 ? lane >> 6 : lane >> 3);
                const uint32_t e = lds_u16(tbl + off + 2u * j);
                const uint64_t sg = (uint64_t)sbase[r] + 2 * pr + hh;
                const uint4 z = philox4x32<ROUNDS>(make_uint4((uint32_t)sg, (uint32_t)(sg >> 32), node, TAG_REFINE), key);
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
            constexpr uint32_t scale = LAY == 1 ? 64u : 32u;
#pragma unroll
            for (int q = 0; q < 8; ++q) par[r][q] = (c2[q] & 0x001f001fu) * scale + cbase[r][q];
            sts_v4(slot_base[r] + (uint32_t)dst * slot_stride, o);
          }
          if (leaf_row >= 0 && store_mode[r]) store_leaf(a, leaf_base[r], leaf_row, col0[r], store_mode[r], o);
        } else {
          constexpr uint32_t SM = 0x03030303u;
          uint32_t raw[4];  // state bytes of the 16 sites (leaf output format)
#pragma unroll
          for (int q = 0; q < 4; ++q) raw[q] = __byte_perm(c2[2 * q], c2[2 * q + 1], 0x6420u) & SM;
          if (dst >= 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) par[r][q] = raw[q] * 8u + cbase[r][q];
            sts_v4(slot_base[r] + (uint32_t)dst * slot_stride, make_uint4(par[r][0], par[r][1], par[r][2], par[r][3]));
          }
          if (leaf_row >= 0 && store_mode[r])
            store_leaf(a, leaf_base[r], leaf_row, col0[r], store_mode[r], make_uint4(raw[0], raw[1], raw[2], raw[3]));
        }
      }
      if (reload) reg_slot = src;
      if (dst >= 0) reg_slot = dst;
    }
    if (PW) {
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(smem_u32(&empty_bar[buf]));
    } else {
      __syncthreads();  // every warp is done with this stage buffer
    }
    if (!PW && tid == 0 && st + nstage < fa.n_stages) {
      fence_async_smem();
      mbar_expect_tx(smem_u32(&full_bar[buf]), (uint32_t)fa.stage_bytes);
      tma_bulk_g2s(sbuf, fa.ftab + (size_t)(st + nstage) * fa.stage_bytes, (uint32_t)fa.stage_bytes, smem_u32(&full_bar[buf]));
    }
    if (++buf == nstage) { buf = 0; phase ^= 1u; }
  }
}

// Gather the per-node alias tables into walk order, interleaved with the walk records, one stage per nps nodes:
// stage = [nps x WalkRec (16 B)] [nps x tb table bytes].  The e16 layout [node][comp][state][column] u16 already is the
// row layout the kernels address (DNA: comp*32 + state*8 + j*2; protein: (comp*K + state)*64 + j*2).
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
  int kind = 0, tb = 0;
  if (h->K == 4 && h->KP == 4 && h->C <= 8) {
    kind = 1; tb = h->C * 32;
  } else if (h->KP == 32 && (int64_t)h->C * h->K * 64 <= 65535) {
    kind = 2; tb = h->C * h->K * 64;
  } else if (h->KP == 32 && (int64_t)h->C * h->K * 32 <= 65535) {
    kind = 3; tb = h->C * h->K * 64;
  }
  if (kind) {
    // stage geometry: keep a CTA under ~44 KB of shared memory so that 5 CTAs (20 warps) fit an SM; small tables
    // (DNA) stream 3 stages of up to 64 nodes, bigger ones 2 stages of a few nodes; tables that do not fit that
    // budget fall back to 2 stages x 1 node with fewer CTAs per SM.
    const size_t per_node = 16 + (size_t)tb;
    int threads = 128, rounds = 1, nstage = 0, nps = 0;
    size_t smem = 0;
    if (kind != 3) {
      const size_t slots = (size_t)h->n_slots * threads * 16;
      nstage = tb <= 1024 ? MAX_NSTAGE : 2;
      size_t budget = 44 * 1024;
      if (const char *e = getenv("ELX_SMEM_BUDGET_KB")) budget = (size_t)atoi(e) * 1024;  // tuning knob
      nps = budget > slots + 128 ? (int)((budget - slots - 128) / nstage / per_node) : 0;
      if (nps > 64) nps = 64;
      if (nps < 1) { nps = 1; nstage = 2; }
      smem = 128 + (size_t)nstage * nps * per_node + slots;
      if (smem > 200 * 1024) kind = 0;
    } else {
      // big mixture: one branch per stage, 2 stages, one 256-thread CTA per SM; 2 rounds (8192 sites per CTA) when the
      // slot stack still fits, which halves the table stream per cell
      // big mixture: one branch per stage, 2 stages, one CTA per SM: 512 threads (8192 sites per CTA amortise the table
      // stream; measured 10 % faster than 256 threads x 2 rounds), 256 threads when the slot stack of a deep tree
      // does not fit next to the two stage buffers
      nstage = 2; nps = 1; rounds = 1;
      for (threads = 512; threads >= 256; threads /= 2) {
        smem = 128 + (size_t)nstage * per_node + (size_t)h->n_slots * threads * 16;
        if (smem <= 220 * 1024) break;
      }
      if (threads < 256) rounds = 0;
      if (rounds < 1) kind = 0;
    }
    ft->nps = nps; ft->nstage = nstage; ft->smem = smem; ft->threads = threads; ft->rounds = rounds;
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

template <int ROUNDS, int LAY, int THREADS, int R, bool PW>
static cudaError_t launch_tiled_pw(const FastArgs &fa, int64_t groups, size_t smem, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(k2_tiled<ROUNDS, LAY, THREADS, R, PW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  unsigned blocks = (unsigned)((groups + (int64_t)THREADS * R - 1) / ((int64_t)THREADS * R));
  k2_tiled<ROUNDS, LAY, THREADS, R, PW><<<blocks, THREADS + (PW ? 32 : 0), smem, stream>>>(fa);
  return cudaGetLastError();
}

template <int ROUNDS, int LAY, int THREADS, int R>
static cudaError_t launch_tiled(const FastArgs &fa, int64_t groups, size_t smem, cudaStream_t stream) {
  // Measured (B200, 10 rounds): the producer warp is +37 % for the big-mixture layout (cfg4, C60: 2.90 -> 3.98e11 cells/s;
  // one 77 KB branch per stage, the CTA barrier was a third of all stall cycles) and -6 ... -11 % for the small-table
  // layouts (cfg3 / cfg5 / node-by-node DNA: 5 CTAs of 128 threads per SM, where a fifth warp per CTA costs more than the
  // barrier; at a register budget that keeps 5 CTAs of 160 threads resident, 72 registers, still -5 % / -2.5 %), so only
  // LAY 2 uses it.  ELX_FAST_PRODUCER=0 switches it off for A/B runs.
  if (LAY == 2) {
    static const bool pw = [] { const char *e = getenv("ELX_FAST_PRODUCER"); return e ? atoi(e) != 0 : true; }();
    if (pw) return launch_tiled_pw<ROUNDS, LAY, THREADS, R, LAY == 2>(fa, groups, smem, stream);
  }
  return launch_tiled_pw<ROUNDS, LAY, THREADS, R, false>(fa, groups, smem, stream);
}

template <int ROUNDS>
static cudaError_t launch_kind(const FastTables *ft, const FastArgs &fa, int64_t groups, cudaStream_t stream) {
  if (ft->kind == 1) return launch_tiled<ROUNDS, 0, 128, 1>(fa, groups, ft->smem, stream);
  if (ft->kind == 2) return launch_tiled<ROUNDS, 1, 128, 1>(fa, groups, ft->smem, stream);
  if (ft->threads == 512) return launch_tiled<ROUNDS, 2, 512, 1>(fa, groups, ft->smem, stream);
  return launch_tiled<ROUNDS, 2, 256, 1>(fa, groups, ft->smem, stream);
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
  fa.two = 2u;
  int64_t g0 = a.site0 >> 4, g1 = (a.site0 + a.nsites + 15) >> 4;
  cudaError_t e = h->rounds == 7 ? launch_kind<7>(ft, fa, g1 - g0, h->stream) : launch_kind<10>(ft, fa, g1 - g0, h->stream);
  if (e != cudaSuccess) return fail(h, ELX_E_CUDA, std::string("k2_tiled launch: ") + cudaGetErrorString(e));
  h->launches++;
  h->last_kernel = "k2_tiled<" + std::to_string(h->rounds) + "," + std::to_string(ft->kind - 1) + "," + std::to_string(ft->threads) + "," +
                   std::to_string(ft->rounds) + (ft->kind == 3 ? ",PW>" : ">");
  *used = 1;
  return 0;
}

}  // namespace elx
