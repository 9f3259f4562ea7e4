// Checks csrc/inflate.cpp (the BGZF member decoder) against zlib: streams compressed by zlib at every level and strategy from
// data of several kinds (BAM-like records, runs, random bytes, text, empty, one byte, 64 KB blocks), decoded by fast_inflate and
// compared byte for byte; then truncated and corrupted streams, which must be refused or decoded to something, never crash or
// write outside the output (guard bytes on both sides).   inflate_check [seconds]   exit code 0 = all good.
#include <zlib.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../breseq_b200/csrc/inflate.h"

using namespace brq;

static std::vector<uint8_t> deflate_raw(const std::vector<uint8_t>& in, int level, int strategy) {
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  deflateInit2(&zs, level, Z_DEFLATED, -15, 8, strategy);
  std::vector<uint8_t> out(deflateBound(&zs, in.size()) + 64);
  zs.next_in = const_cast<Bytef*>(in.data()); zs.avail_in = (uInt)in.size();
  zs.next_out = out.data(); zs.avail_out = (uInt)out.size();
  deflate(&zs, Z_FINISH);
  out.resize(zs.total_out);
  deflateEnd(&zs);
  return out;
}

int main(int argc, char** argv) {
  const double seconds = argc > 1 ? atof(argv[1]) : 3.0;
  std::mt19937_64 rng(12345);
  const auto t0 = std::chrono::steady_clock::now();
  size_t n_ok = 0, n_refused = 0, n_fallback = 0;
  for (int round = 0;; ++round) {
    if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > seconds && round >= 40) break;
    const int kind = round % 7;
    size_t n = kind == 4 ? 0 : kind == 5 ? 1 : kind == 6 ? 65280 : (size_t)(rng() % 70000);
    std::vector<uint8_t> data(n);
    if (kind == 0) for (size_t i = 0; i < n; ++i) data[i] = (uint8_t)rng();                                     // random: stored blocks
    else if (kind == 1) for (size_t i = 0; i < n; ++i) data[i] = (uint8_t)("ACGT"[rng() & 3]);                   // four symbols
    else if (kind == 2) for (size_t i = 0; i < n; ++i) data[i] = (uint8_t)(i / (1 + rng() % 300) % 7 + 40);      // runs: long matches, short distances
    else for (size_t i = 0; i < n; ++i) data[i] = (uint8_t)((i % 37 < 20) ? (i * 7 + (i >> 8)) : (rng() & 31)); // record-like mix
    for (int level : {1, 6, 9}) for (int strategy : {Z_DEFAULT_STRATEGY, Z_FIXED, Z_HUFFMAN_ONLY, Z_RLE}) {
      const std::vector<uint8_t> z = deflate_raw(data, level, strategy);
      std::vector<uint8_t> out(n + 64, 0xA5);
      const bool ok = fast_inflate(z.data(), z.size(), out.data() + 32, n);
      for (int g = 0; g < 32; ++g) if (out[g] != 0xA5 || out[32 + n + g] != 0xA5) { fprintf(stderr, "guard bytes overwritten\n"); return 1; }
      if (!ok) { ++n_fallback; fprintf(stderr, "refused a valid stream: kind %d n %zu level %d strategy %d\n", kind, n, level, strategy); return 1; }
      if (n && memcmp(out.data() + 32, data.data(), n) != 0) { fprintf(stderr, "wrong bytes: kind %d n %zu level %d strategy %d\n", kind, n, level, strategy); return 1; }
      ++n_ok;
      // damaged streams: truncated, one byte flipped, wrong output size
      if (z.size() > 2) {
        std::vector<uint8_t> bad(z.begin(), z.begin() + (size_t)(rng() % z.size()));
        std::vector<uint8_t> o2(n + 64, 0xA5);
        if (!fast_inflate(bad.data(), bad.size(), o2.data() + 32, n)) ++n_refused;
        for (int g = 0; g < 32; ++g) if (o2[g] != 0xA5 || o2[32 + n + g] != 0xA5) { fprintf(stderr, "guard bytes overwritten (truncated)\n"); return 1; }
        bad = z; bad[(size_t)(rng() % bad.size())] ^= (uint8_t)(1u << (rng() & 7));
        std::fill(o2.begin(), o2.end(), 0xA5);
        if (!fast_inflate(bad.data(), bad.size(), o2.data() + 32, n)) ++n_refused;
        for (int g = 0; g < 32; ++g) if (o2[g] != 0xA5 || o2[32 + n + g] != 0xA5) { fprintf(stderr, "guard bytes overwritten (flipped)\n"); return 1; }
        if (n > 0) {
          std::fill(o2.begin(), o2.end(), 0xA5);
          if (fast_inflate(z.data(), z.size(), o2.data() + 32, n - 1)) { fprintf(stderr, "accepted a stream longer than its output\n"); return 1; }
          for (int g = 0; g < 32; ++g) if (o2[g] != 0xA5 || o2[32 + n - 1 + g] != 0xA5) { fprintf(stderr, "guard bytes overwritten (short output)\n"); return 1; }
        }
      }
    }
  }
  printf("inflate_check: %zu streams equal, %zu damaged streams refused\n", n_ok, n_refused);
  return 0;
}
