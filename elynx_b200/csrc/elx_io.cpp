// Host-side output helper: parallel gzip writer for big alignments.
//
// The reference writes `<base>.fasta.gz` with `BL.writeFile f (compress r)` (elynx-tools/src/ELynx/Tools/InputOutput.hs:80-83):
// one deflate stream, one thread -- minutes for a 10 GB alignment the GPU simulates in milliseconds.  Here the buffer is
// cut into blocks that worker threads deflate independently, each into its own gzip MEMBER; members are written in order.
// A multi-member file is a valid gzip file: gunzip, Python's gzip module and the reference's reader (`decompress`,
// InputOutput.hs:73-76; zlib >= 0.6 concatenates members) all return the original bytes.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <map>
#include <sys/mman.h>
#include <zlib.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/elynx_b200.h"
#include "../../include/elynx_b200_model.h"
#include "elx_state.hpp"

namespace {

bool deflate_member(const uint8_t *src, size_t n, int level, std::string &out, std::string &err) {
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (deflateInit2(&zs, level, Z_DEFLATED, 15 + 16 /* gzip wrapper */, 8, Z_DEFAULT_STRATEGY) != Z_OK) {
    err = "deflateInit2 failed";
    return false;
  }
  out.resize(deflateBound(&zs, (uLong)n) + 32);
  zs.next_in = const_cast<Bytef *>(src);
  zs.avail_in = (uInt)n;
  zs.next_out = reinterpret_cast<Bytef *>(&out[0]);
  zs.avail_out = (uInt)out.size();
  const int rc = deflate(&zs, Z_FINISH);
  const size_t produced = out.size() - zs.avail_out;
  deflateEnd(&zs);
  if (rc != Z_STREAM_END) {
    err = "deflate did not finish";
    return false;
  }
  out.resize(produced);
  return true;
}

// ---- 2-bit transport decoding -------------------------------------------------------------------------------------
// Alignments over <= 4 states cross PCIe packed four sites to a byte (site 4i+k in bits 2k..2k+1, elx_core.cu:k_pack2);
// the host expands them to one byte per site -- state codes, or alphabet characters when a 4-entry table is given.
void unpack2_scalar(const uint8_t *src, uint8_t *dst, int64_t n, const uint8_t *lut4) {
  for (int64_t i = 0; i < n; ++i) {
    const uint8_t st = (uint8_t)((src[i >> 2] >> (2 * (i & 3))) & 3);
    dst[i] = lut4 ? lut4[st] : st;
  }
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) void unpack2_avx2(const uint8_t *src, uint8_t *dst, int64_t n, const uint8_t *lut4) {
  int64_t i = 0;
  // scalar head (in steps of 4 sites) until dst is 32-byte aligned, if it ever is
  const uintptr_t mis = (uintptr_t)dst & 31;
  const bool can_align = (mis & 3) == 0;
  if (can_align && mis) {
    int64_t head = (int64_t)(32 - mis);
    if (head > n) head = n;
    unpack2_scalar(src, dst, head, lut4);
    if (head & 3) return;  // n exhausted inside the head
    i = head;
  }
  const __m256i rep = _mm256_setr_epi8(0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7);
  const __m256i m0 = _mm256_set1_epi32(0x00000003), m1 = _mm256_set1_epi32(0x00000300), m2 = _mm256_set1_epi32(0x00030000),
                m3 = _mm256_set1_epi32(0x03000000);
  __m256i lutv = _mm256_setzero_si256();
  if (lut4) {
    uint32_t w;
    memcpy(&w, lut4, 4);
    lutv = _mm256_set1_epi32((int)w);  // bytes 0..3 of each lane; indices are 0..3
  }
  auto expand = [&](const uint8_t *p) __attribute__((target("avx2"))) -> __m256i {
    const __m128i q = _mm_loadl_epi64(reinterpret_cast<const __m128i *>(p));  // 8 packed bytes = 32 sites
    const __m256i x = _mm256_shuffle_epi8(_mm256_broadcastsi128_si256(q), rep);
    __m256i r = _mm256_and_si256(x, m0);
    r = _mm256_or_si256(r, _mm256_and_si256(_mm256_srli_epi32(x, 2), m1));
    r = _mm256_or_si256(r, _mm256_and_si256(_mm256_srli_epi32(x, 4), m2));
    r = _mm256_or_si256(r, _mm256_and_si256(_mm256_srli_epi32(x, 6), m3));
    return lut4 ? _mm256_shuffle_epi8(lutv, r) : r;
  };
  static const bool no_nt = getenv("ELX_UNPACK_NO_NT") != nullptr;
  if (can_align && no_nt) {
    for (; i + 64 <= n; i += 64) {
      const __m256i a = expand(src + (i >> 2)), b = expand(src + (i >> 2) + 8);
      _mm256_store_si256(reinterpret_cast<__m256i *>(dst + i), a);
      _mm256_store_si256(reinterpret_cast<__m256i *>(dst + i + 32), b);
    }
  } else if (can_align) {
    for (; i + 64 <= n; i += 64) {  // one cache line of output per iteration, streaming stores (no read for ownership)
      const __m256i a = expand(src + (i >> 2)), b = expand(src + (i >> 2) + 8);
      _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), a);
      _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i + 32), b);
    }
    _mm_sfence();
  } else {
    for (; i + 32 <= n; i += 32) _mm256_storeu_si256(reinterpret_cast<__m256i *>(dst + i), expand(src + (i >> 2)));
  }
  unpack2_scalar(src + (i >> 2), dst + i, n - i, lut4);
}
#endif

// ---- 5-bit transport decoding (alphabets of <= 32 states: amino acids) --------------------------------------------
// site i of a row sits in bits 5i..5i+4 of the row's little-endian bit stream (elx_core.cu:k_pack5): 8 sites = 5 bytes.
void unpack5_scalar(const uint8_t *src, uint8_t *dst, int64_t i0, int64_t n, const uint8_t *lut) {
  for (int64_t i = i0; i < n; ++i) {
    const int64_t bit = 5 * i;
    const uint32_t lo = src[bit >> 3], hi = (bit & 7) > 3 ? src[(bit >> 3) + 1] : 0u;  // a field spans at most 2 bytes
    const uint8_t st = (uint8_t)(((lo | (hi << 8)) >> (bit & 7)) & 31);
    dst[i] = lut ? lut[st] : st;
  }
}

#if defined(__x86_64__)
__attribute__((target("avx2,bmi2"))) void unpack5_bmi2(const uint8_t *src, uint8_t *dst, int64_t n, const uint8_t *lut) {
  const uint64_t M = 0x1f1f1f1f1f1f1f1fULL;
  __m256i lut_lo = _mm256_setzero_si256(), lut_hi = _mm256_setzero_si256();
  if (lut) {
    lut_lo = _mm256_broadcastsi128_si256(_mm_loadu_si128(reinterpret_cast<const __m128i *>(lut)));
    lut_hi = _mm256_broadcastsi128_si256(_mm_loadu_si128(reinterpret_cast<const __m128i *>(lut + 16)));
  }
  const bool aligned = ((uintptr_t)dst & 31) == 0;
  int64_t i = 0;
  // 32 sites = 20 packed bytes per iteration; the 8-byte loads read 3 bytes past each 5-byte group, so the last 32-site
  // group of the row is left to the exact scalar tail
  for (; i + 64 <= n; i += 32) {
    const uint8_t *p = src + (i >> 3) * 5;
    uint64_t v0, v1, v2, v3;
    memcpy(&v0, p, 8); memcpy(&v1, p + 5, 8); memcpy(&v2, p + 10, 8); memcpy(&v3, p + 15, 8);
    __m256i r = _mm256_set_epi64x((long long)_pdep_u64(v3, M), (long long)_pdep_u64(v2, M), (long long)_pdep_u64(v1, M), (long long)_pdep_u64(v0, M));
    if (lut) {
      const __m256i lo = _mm256_shuffle_epi8(lut_lo, r), hi = _mm256_shuffle_epi8(lut_hi, r);  // index bits 0..3, bit 7 clear
      r = _mm256_blendv_epi8(lo, hi, _mm256_cmpgt_epi8(r, _mm256_set1_epi8(15)));
    }
    if (aligned) _mm256_stream_si256(reinterpret_cast<__m256i *>(dst + i), r);
    else _mm256_storeu_si256(reinterpret_cast<__m256i *>(dst + i), r);
  }
  if (aligned) _mm_sfence();
  unpack5_scalar(src, dst, i, n, lut);
}

// AVX-512 VBMI: 64 sites per iteration -- one byte permute brings every 5-byte group into its own 64-bit lane, one
// multishift extracts the eight 5-bit fields of a lane, one more permute is the 32-entry alphabet look-up.
__attribute__((target("avx512f,avx512bw,avx512vbmi"))) void unpack5_vbmi(const uint8_t *src, uint8_t *dst, int64_t n, const uint8_t *lut) {
  alignas(64) uint8_t idx_b[64], ctl_b[64], lut_b[64];
  for (int q = 0; q < 8; ++q)
    for (int b = 0; b < 8; ++b) {
      idx_b[8 * q + b] = (uint8_t)std::min(5 * q + b, 63);
      ctl_b[8 * q + b] = (uint8_t)(5 * b);
    }
  memset(lut_b, '?', sizeof lut_b);
  if (lut) memcpy(lut_b, lut, 32);
  const __m512i idx = _mm512_load_si512(idx_b), ctl = _mm512_load_si512(ctl_b), lutv = _mm512_load_si512(lut_b);
  const __m512i m5 = _mm512_set1_epi8(31);
  const __mmask64 k40 = (__mmask64)((1ULL << 40) - 1);
  // scalar head (whole 8-site groups) up to a 64-byte boundary of dst, when there is one on a group boundary
  const bool aligned = ((uintptr_t)dst & 7) == 0;
  int64_t i = 0;
  if (aligned && ((uintptr_t)dst & 63)) {
    const int64_t head = std::min<int64_t>(n, 64 - (int64_t)((uintptr_t)dst & 63));
    unpack5_scalar(src, dst, 0, head, lut);
    i = head;
  }
  for (; i + 64 <= n; i += 64) {
    const __m512i v = _mm512_maskz_loadu_epi8(k40, src + (i >> 3) * 5);  // exactly the 40 bytes of these 64 sites
    __m512i r = _mm512_and_si512(_mm512_multishift_epi64_epi8(ctl, _mm512_permutexvar_epi8(idx, v)), m5);
    if (lut) r = _mm512_permutexvar_epi8(r, lutv);
    if (aligned) _mm512_stream_si512(reinterpret_cast<__m512i *>(dst + i), r);
    else _mm512_storeu_si512(dst + i, r);
  }
  if (aligned) _mm_sfence();
  unpack5_scalar(src, dst, i, n, lut);
}
#endif

}  // namespace

namespace elx {
void unpack2_rows(const uint8_t *packed, int64_t packed_pitch, uint8_t *dst, int64_t dst_pitch, int64_t row0, int64_t row1,
                  int64_t nsites, const uint8_t *lut4) {
#if defined(__x86_64__)
  static const bool have_avx2 = __builtin_cpu_supports("avx2");
  if (have_avx2) {
    for (int64_t r = row0; r < row1; ++r) unpack2_avx2(packed + r * packed_pitch, dst + r * dst_pitch, nsites, lut4);
    return;
  }
#endif
  for (int64_t r = row0; r < row1; ++r) unpack2_scalar(packed + r * packed_pitch, dst + r * dst_pitch, nsites, lut4);
}

void unpack5_rows(const uint8_t *packed, int64_t packed_pitch, uint8_t *dst, int64_t dst_pitch, int64_t row0, int64_t row1,
                  int64_t nsites, const uint8_t *lut32) {
#if defined(__x86_64__)
  static const bool vbmi = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vbmi") &&
                           !getenv("ELX_NO_AVX512");
  if (vbmi) {
    for (int64_t r = row0; r < row1; ++r) unpack5_vbmi(packed + r * packed_pitch, dst + r * dst_pitch, nsites, lut32);
    return;
  }
  static const bool fast = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2");
  if (fast) {
    for (int64_t r = row0; r < row1; ++r) unpack5_bmi2(packed + r * packed_pitch, dst + r * dst_pitch, nsites, lut32);
    return;
  }
#endif
  for (int64_t r = row0; r < row1; ++r) unpack5_scalar(packed + r * packed_pitch, dst + r * dst_pitch, 0, nsites, lut32);
}

// ---- ordered multi-member gzip stream ------------------------------------------------------------------------------
// append() cuts the byte stream into members of block_bytes; worker threads deflate them; the appending thread writes the
// finished members in order whenever more than `max_inflight` are pending (back-pressure) and at close.  The file is the
// same sequence of members elx_gzip_write produces for the same bytes and block size.
struct GzJob {
  std::string in, out, err;
  std::atomic<int> done{0};
  bool ok = true;
};
struct GzStream {
  FILE *f = nullptr;
  int level = 6;
  size_t block = (size_t)4 << 20;
  std::vector<std::thread> pool;
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  std::deque<GzJob *> queue;     // submitted, not yet taken by a worker
  std::deque<GzJob *> inflight;  // submitted, not yet written (file order)
  std::string cur;
  bool closing = false, failed = false, any = false;
  std::string error;
  size_t max_inflight = 8;
};

static void gz_worker(GzStream *g) {
  for (;;) {
    GzJob *j = nullptr;
    {
      std::unique_lock<std::mutex> lk(g->mu);
      g->cv_work.wait(lk, [&]() { return !g->queue.empty() || g->closing; });
      if (g->queue.empty()) return;
      j = g->queue.front();
      g->queue.pop_front();
    }
    j->ok = deflate_member(reinterpret_cast<const uint8_t *>(j->in.data()), j->in.size(), g->level, j->out, j->err);
    std::string().swap(j->in);
    {
      std::lock_guard<std::mutex> lk(g->mu);
      j->done.store(1, std::memory_order_release);
    }
    g->cv_done.notify_all();
  }
}

static void gz_drain(GzStream *g, size_t keep) {  // write the oldest members until at most `keep` are pending
  for (;;) {
    GzJob *j = nullptr;
    {
      std::unique_lock<std::mutex> lk(g->mu);
      if (g->inflight.size() <= keep) return;
      j = g->inflight.front();
      g->cv_done.wait(lk, [&]() { return j->done.load(std::memory_order_acquire) != 0; });
      g->inflight.pop_front();
    }
    if (!j->ok) { g->failed = true; if (g->error.empty()) g->error = j->err; }
    else if (!g->failed && fwrite(j->out.data(), 1, j->out.size(), g->f) != j->out.size()) { g->failed = true; g->error = "short write"; }
    delete j;
  }
}

static void gz_submit(GzStream *g) {
  GzJob *j = new GzJob();
  j->in.swap(g->cur);
  g->any = true;
  {
    std::lock_guard<std::mutex> lk(g->mu);
    g->queue.push_back(j);
    g->inflight.push_back(j);
  }
  g->cv_work.notify_one();
  gz_drain(g, g->max_inflight);
}

GzStream *gz_open(const char *path, int n_threads, int level, int64_t block_bytes, std::string &err) {
  FILE *f = fopen(path, "wb");
  if (!f) { err = std::string("cannot open ") + path; return nullptr; }
  GzStream *g = new GzStream();
  g->f = f;
  g->level = (level < 0 || level > 9) ? 6 : level;
  if (block_bytes > 0) g->block = (size_t)std::min<int64_t>(block_bytes, (int64_t)1 << 30);
  int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  g->max_inflight = (size_t)nt * 4;
  for (int t = 0; t < nt; ++t) g->pool.emplace_back(gz_worker, g);
  return g;
}

bool gz_append(GzStream *g, const uint8_t *data, size_t n) {
  while (n > 0 && !g->failed) {
    const size_t room = g->block - g->cur.size(), take = n < room ? n : room;
    g->cur.append(reinterpret_cast<const char *>(data), take);
    data += take;
    n -= take;
    if (g->cur.size() == g->block) gz_submit(g);
  }
  return !g->failed;
}

bool gz_close(GzStream *g, std::string &err) {
  if (!g->cur.empty() || !g->any) gz_submit(g);  // an empty stream is one empty member
  gz_drain(g, 0);
  {
    std::lock_guard<std::mutex> lk(g->mu);
    g->closing = true;
  }
  g->cv_work.notify_all();
  for (auto &t : g->pool) t.join();
  if (fclose(g->f) != 0 && !g->failed) { g->failed = true; g->error = "close failed"; }
  const bool ok = !g->failed;
  err = g->error;
  delete g;
  return ok;
}
}  // namespace elx

extern "C" {

int elx_unpack5_rows(const uint8_t *packed, int64_t packed_pitch, uint8_t *dst, int64_t dst_pitch, int64_t rows, int64_t nsites,
                     const char *lut32, int32_t n_threads) {
  if (rows < 0 || nsites < 0 || ((!packed || !dst) && rows > 0 && nsites > 0)) return elx::fail(nullptr, ELX_E_INVALID, "elx_unpack5_rows: bad argument");
  if (packed_pitch < (5 * nsites + 7) / 8 || dst_pitch < nsites) return elx::fail(nullptr, ELX_E_INVALID, "elx_unpack5_rows: pitch too small");
  uint8_t lut[32];
  if (lut32) {
    memset(lut, '?', sizeof lut);
    memcpy(lut, lut32, std::min<size_t>(32, strlen(lut32)));
  }
  const uint8_t *lp = lut32 ? lut : nullptr;
  int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt > rows) nt = (int)(rows > 0 ? rows : 1);
  std::vector<std::thread> pool;
  for (int w = 1; w < nt; ++w)
    pool.emplace_back([=]() { elx::unpack5_rows(packed, packed_pitch, dst, dst_pitch, rows * w / nt, rows * (w + 1) / nt, nsites, lp); });
  elx::unpack5_rows(packed, packed_pitch, dst, dst_pitch, 0, rows / nt, nsites, lp);
  for (auto &t : pool) t.join();
  return 0;
}

int elx_unpack2_rows(const uint8_t *packed, int64_t packed_pitch, uint8_t *dst, int64_t dst_pitch, int64_t rows, int64_t nsites,
                     const char *lut4, int32_t n_threads) {
  if (rows < 0 || nsites < 0 || ((!packed || !dst) && rows > 0 && nsites > 0)) return elx::fail(nullptr, ELX_E_INVALID, "elx_unpack2_rows: bad argument");
  if (packed_pitch < (nsites + 3) / 4 || dst_pitch < nsites) return elx::fail(nullptr, ELX_E_INVALID, "elx_unpack2_rows: pitch too small");
  if (lut4 && strlen(lut4) < 4) return elx::fail(nullptr, ELX_E_INVALID, "elx_unpack2_rows: the character table needs 4 entries");
  int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt > rows) nt = (int)(rows > 0 ? rows : 1);
  std::vector<std::thread> pool;
  for (int w = 1; w < nt; ++w)
    pool.emplace_back([=]() { elx::unpack2_rows(packed, packed_pitch, dst, dst_pitch, rows * w / nt, rows * (w + 1) / nt, nsites, (const uint8_t *)lut4); });
  elx::unpack2_rows(packed, packed_pitch, dst, dst_pitch, 0, rows / nt, nsites, (const uint8_t *)lut4);
  for (auto &t : pool) t.join();
  return 0;
}

int elx_gzip_write(const char *path, const uint8_t *data, int64_t size, int32_t n_threads, int32_t level, int64_t block_bytes) {
  if (!path || size < 0 || (!data && size > 0)) return elx::fail(nullptr, ELX_E_INVALID, "elx_gzip_write: bad argument");
  if (level < 0 || level > 9) level = 6;
  if (block_bytes <= 0) block_bytes = (int64_t)4 << 20;
  if (block_bytes > ((int64_t)1 << 30)) block_bytes = (int64_t)1 << 30;  // zlib counts in 32 bits
  int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  const int64_t n_blocks = size == 0 ? 1 : (size + block_bytes - 1) / block_bytes;  // an empty input is one empty member
  if (nt > n_blocks) nt = (int)n_blocks;
  FILE *f = fopen(path, "wb");
  if (!f) return elx::fail(nullptr, ELX_E_INVALID, std::string("elx_gzip_write: cannot open ") + path);
  // blocks are deflated in waves of `nt * 4` so that the compressed copies in flight stay small next to the input
  const int64_t wave = (int64_t)nt * 4;
  std::vector<std::string> outs((size_t)wave);
  std::string first_error;
  std::atomic<bool> failed{false};
  for (int64_t w0 = 0; w0 < n_blocks && !failed; w0 += wave) {
    const int64_t w1 = w0 + wave < n_blocks ? w0 + wave : n_blocks;
    std::atomic<int64_t> next{w0};
    std::vector<std::string> errs((size_t)nt);
    auto work = [&](int tidx) {
      for (;;) {
        const int64_t b = next.fetch_add(1);
        if (b >= w1 || failed) return;
        const int64_t off = b * block_bytes, len = off + block_bytes <= size ? block_bytes : size - off;
        if (!deflate_member(data + off, (size_t)(len > 0 ? len : 0), level, outs[(size_t)(b - w0)], errs[(size_t)tidx])) failed = true;
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto &th : pool) th.join();
    for (const std::string &e : errs)
      if (!e.empty() && first_error.empty()) first_error = e;
    if (failed) break;
    for (int64_t b = w0; b < w1; ++b) {
      const std::string &o = outs[(size_t)(b - w0)];
      if (fwrite(o.data(), 1, o.size(), f) != o.size()) { failed = true; first_error = "short write"; break; }
    }
  }
  if (fclose(f) != 0 && !failed) { failed = true; first_error = "close failed"; }
  if (failed) return elx::fail(nullptr, ELX_E_INVALID, std::string("elx_gzip_write: ") + (first_error.empty() ? "failed" : first_error));
  return 0;
}

// ---- result buffers owned by the library ---------------------------------------------------------------------------
// simulateAndFlatten*Par RETURN their result, so whoever stands in for them decides where it lives.  A fresh allocation of
// 10 GB costs more than the simulation and its transport together (every page is zeroed by the kernel when it is first
// written: +110 ms with 2 MB pages, +400 ms with 4 KB pages); buffers from this pool are huge-page advised, recycled on
// free and therefore faulted in once per process, not once per call.
namespace {
std::mutex g_res_mu;
std::map<void *, size_t> g_res_live;                  // handed out: pointer -> mapped bytes
std::multimap<size_t, void *> g_res_free;             // mapped bytes -> pointer
size_t g_res_free_bytes = 0;
size_t res_pool_cap() {
  static const size_t cap = [] {
    const char *e = getenv("ELX_RESULT_POOL_MB");
    return (size_t)(e ? atoll(e) : 32768) << 20;
  }();
  return cap;
}
}  // namespace

void *elx_result_alloc(int64_t bytes) {
  if (bytes <= 0) return nullptr;
  const size_t need = ((size_t)bytes + 0x1fffff) & ~(size_t)0x1fffff;
  {
    std::lock_guard<std::mutex> lk(g_res_mu);
    auto it = g_res_free.lower_bound(need);
    if (it != g_res_free.end() && it->first <= need + need / 4) {  // best fit within +25 %
      void *p = it->second;
      g_res_live[p] = it->first;
      g_res_free_bytes -= it->first;
      g_res_free.erase(it);
      return p;
    }
  }
  // 2 MB aligned, so that every huge page of the mapping can be one
  const size_t span = need + 0x200000;
  uint8_t *raw = static_cast<uint8_t *>(mmap(nullptr, span, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0));
  if (raw == MAP_FAILED) return nullptr;
  uint8_t *p = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 0x1fffff) & ~(uintptr_t)0x1fffff);
  if (p > raw) munmap(raw, (size_t)(p - raw));
  if (p + need < raw + span) munmap(p + need, (size_t)(raw + span - (p + need)));
#ifdef MADV_HUGEPAGE
  madvise(p, need, MADV_HUGEPAGE);
#endif
  std::lock_guard<std::mutex> lk(g_res_mu);
  g_res_live[p] = need;
  return p;
}

void elx_result_free(void *p) {
  if (!p) return;
  size_t n = 0;
  {
    std::lock_guard<std::mutex> lk(g_res_mu);
    auto it = g_res_live.find(p);
    if (it == g_res_live.end()) return;  // not ours
    n = it->second;
    g_res_live.erase(it);
    if (g_res_free_bytes + n <= res_pool_cap()) {
      g_res_free.emplace(n, p);
      g_res_free_bytes += n;
      return;
    }
  }
  munmap(p, n);
}

void elx_result_pool_trim(void) {
  std::multimap<size_t, void *> drop;
  {
    std::lock_guard<std::mutex> lk(g_res_mu);
    drop.swap(g_res_free);
    g_res_free_bytes = 0;
  }
  for (auto &kv : drop) munmap(kv.second, kv.first);
}

}  // extern "C"
