// Host-side output helper: parallel gzip writer for big alignments.
//
// The reference writes `<base>.fasta.gz` with `BL.writeFile f (compress r)` (elynx-tools/src/ELynx/Tools/InputOutput.hs:80-83):
// one deflate stream, one thread -- minutes for a 10 GB alignment the GPU simulates in milliseconds.  Here the buffer is
// cut into blocks that worker threads deflate independently, each into its own gzip MEMBER; members are written in order.
// A multi-member file is a valid gzip file: gunzip, Python's gzip module and the reference's reader (`decompress`,
// InputOutput.hs:73-76; zlib >= 0.6 concatenates members) all return the original bytes.
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <zlib.h>

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

}  // namespace

extern "C" {

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

}  // extern "C"
