#include "bam_io.h"
#include "inflate.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <mutex>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <thread>

namespace brq {

namespace {

// the file as it lies in the page cache: no copy, the inflate threads read it in place
struct MappedFile {
  const uint8_t* p = nullptr;
  size_t n = 0;
  explicit MappedFile(const std::string& path) {
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) throw std::runtime_error("cannot open " + path);
    struct stat sb;
    if (fstat(fd, &sb) != 0) { close(fd); throw std::runtime_error("cannot stat " + path); }
    n = (size_t)sb.st_size;
    if (n) {
      void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
      if (m == MAP_FAILED) { close(fd); throw std::runtime_error("cannot map " + path); }
      madvise(m, n, MADV_SEQUENTIAL);
      p = static_cast<const uint8_t*>(m);
    }
    close(fd);
  }
  ~MappedFile() { if (p) munmap(const_cast<uint8_t*>(p), n); }
  MappedFile(const MappedFile&) = delete;
  MappedFile& operator=(const MappedFile&) = delete;
  size_t size() const { return n; }
  const uint8_t& operator[](size_t i) const { return p[i]; }
};

struct BgzfBlock { size_t cdata, clen, uoff; uint32_t isize; };

// Pass 1 over the compressed image: block boundaries and output offsets, so the members can be
// inflated independently.
std::vector<BgzfBlock> index_blocks(const MappedFile& in, size_t& total) {
  std::vector<BgzfBlock> blocks;
  size_t p = 0;
  total = 0;
  while (p + 18 <= in.size()) {
    if (in[p] != 31 || in[p + 1] != 139 || !(in[p + 3] & 4)) throw std::runtime_error("not a BGZF member");
    uint32_t xlen = in[p + 10] | (in[p + 11] << 8);
    size_t x = p + 12, xend = x + xlen;
    int bsize = -1;
    while (x + 4 <= xend) {
      uint32_t slen = in[x + 2] | (in[x + 3] << 8);
      if (in[x] == 'B' && in[x + 1] == 'C' && slen == 2) bsize = in[x + 4] | (in[x + 5] << 8);
      x += 4 + slen;
    }
    if (bsize < 0) throw std::runtime_error("BGZF member without BC subfield");
    size_t len = (size_t)bsize + 1;
    if (p + len > in.size()) throw std::runtime_error("truncated BGZF member");
    BgzfBlock b;
    b.cdata = xend;
    b.clen = len - (xend - p) - 8;
    memcpy(&b.isize, &in[p + len - 4], 4);
    b.uoff = total;
    total += b.isize;
    blocks.push_back(b);
    p += len;
  }
  if (p != in.size()) throw std::runtime_error("trailing bytes after the last BGZF member");
  return blocks;
}

void inflate_block(const MappedFile& in, const BgzfBlock& b, uint8_t* out) {
  if (!b.isize) return;
  // our own decoder first (inflate.cpp: two to three times zlib's speed on BAM members); zlib for anything it refuses
  static const bool zlib_only = getenv("BRQ_ZLIB_INFLATE") != nullptr;
  if (!zlib_only && fast_inflate(&in[b.cdata], b.clen, out + b.uoff, b.isize)) return;
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (inflateInit2(&zs, -15) != Z_OK) throw std::runtime_error("inflateInit2 failed");
  zs.next_in = const_cast<Bytef*>(&in[b.cdata]);
  zs.avail_in = (uInt)b.clen;
  zs.next_out = out + b.uoff;
  zs.avail_out = b.isize;
  int r = inflate(&zs, Z_FINISH);
  inflateEnd(&zs);
  if (r != Z_STREAM_END) throw std::runtime_error("corrupt BGZF payload");
}

template <typename T> T rd(const uint8_t* p) { T v; memcpy(&v, p, sizeof(T)); return v; }

// Aux walk for the handful of tags the path reads (alignment.cpp:52-57, 73-76; alignment.h:389-410).
size_t aux_value_size(uint8_t type) {
  switch (type) { case 'A': case 'c': case 'C': return 1; case 's': case 'S': return 2;
                  case 'i': case 'I': case 'f': return 4; case 'd': return 8; default: return 0; }
}
int64_t aux_int(uint8_t type, const uint8_t* v) {
  switch (type) { case 'c': return rd<int8_t>(v); case 'C': return rd<uint8_t>(v); case 's': return rd<int16_t>(v);
                  case 'S': return rd<uint16_t>(v); case 'i': return rd<int32_t>(v); case 'I': return rd<uint32_t>(v);
                  default: return 0; }
}

void parse_rg_lines(const std::string& text, ReadGroups& rg) {
  size_t p = 0;
  while (p < text.size()) {
    size_t e = text.find('\n', p);
    if (e == std::string::npos) e = text.size();
    if (e - p >= 3 && text.compare(p, 3, "@RG") == 0) {
      std::string id, lb;
      bool has_id = false;
      size_t f = p + 3;
      while (f < e) {
        size_t g = text.find('\t', f + 1);
        if (g == std::string::npos || g > e) g = e;
        if (text[f] == '\t' && g - f >= 4 && text[f + 3] == ':') {
          std::string key = text.substr(f + 1, 2), val = text.substr(f + 4, g - f - 4);
          if (key == "ID") { id = val; has_id = true; }
          else if (key == "LB") lb = val;
        }
        f = g;
      }
      if (has_id) { rg.ids.push_back(id); rg.libraries.push_back(lb); }  // an @RG without ID is unreferable
    }
    p = e + 1;
  }
}

}  // namespace

void read_bam(const std::string& path, BamHeader& hdr, ReadBatch& reads, int threads, const ReadSpans* span_of) {
  static const bool phase_times = getenv("BRQ_STAGE_TIMES") != nullptr;  // wall time of every phase on stderr
  auto phase_t0 = std::chrono::steady_clock::now();
  auto phase_done = [&](const char* what) {
    if (!phase_times) return;
    const auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "read_bam: %-19s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t - phase_t0).count());
    phase_t0 = t;
  };
  MappedFile in(path);
  phase_done("map file");
  size_t total = 0;
  std::vector<BgzfBlock> blocks = index_blocks(in, total);
  phase_done("index members");
  RawVec<uint8_t> u;   // (uninitialised: the inflate threads touch every page first)
  u.resize(total);
  if (threads < 1) threads = 1;

  // ONE PIPELINE: the worker threads inflate the members (in file order, sixteen at a time); this thread follows the
  // inflated prefix, parses the header, finds every record (a serial chain: each record's length leads to the next) and
  // hands finished runs of records back to the workers, which decode them when no member is left to inflate.  The
  // record walk and the decode hide behind the inflate instead of following it.
  constexpr size_t CHUNK = 16, RUN = 8192;
  const size_t n_chunks = (blocks.size() + CHUNK - 1) / CHUNK;
  std::vector<std::atomic<uint8_t>> chunk_done(n_chunks + 1);
  for (auto& f : chunk_done) f.store(0, std::memory_order_relaxed);
  std::vector<size_t> chunk_end(n_chunks + 1, 0);   // uncompressed offset where every chunk ends
  for (size_t k = 0; k < n_chunks; ++k) { const BgzfBlock& last = blocks[std::min(blocks.size(), (k + 1) * CHUNK) - 1]; chunk_end[k] = last.uoff + last.isize; }
  std::atomic<size_t> next_chunk(0), runs_published(0), next_run(0);
  std::atomic<bool> failed(false), walk_finished(false);
  std::mutex err_mu;
  std::string err;
  auto fail = [&](const std::string& what) { std::lock_guard<std::mutex> g(err_mu); if (err.empty()) err = what; failed = true; };

  // destination arrays at their upper bounds (untouched pages cost nothing; trimmed when the walk is done): a record is at
  // least 36 bytes, a base at least one and a half, a CIGAR operation four
  const size_t n0 = reads.size();
  const uint64_t n_cig0 = reads.cigars.size(), n_seq0 = reads.bases.size();
  const size_t max_reads = total / 36 + 1, max_seq = total * 2 / 3 + 16, max_cig = total / 4 + 1;
  auto grow = [&](auto& v, size_t n) { v.resize(n); };
  grow(reads.tid, n0 + max_reads); grow(reads.pos, n0 + max_reads); grow(reads.flag, n0 + max_reads); grow(reads.mapq, n0 + max_reads);
  grow(reads.n_cigar, n0 + max_reads); grow(reads.cigar_off, n0 + max_reads); grow(reads.l_seq, n0 + max_reads); grow(reads.seq_off, n0 + max_reads);
  grow(reads.x1, n0 + max_reads); grow(reads.xl, n0 + max_reads); grow(reads.xr, n0 + max_reads); grow(reads.as, n0 + max_reads); grow(reads.rg, n0 + max_reads);
  grow(reads.cigars, n_cig0 + max_cig); grow(reads.bases, n_seq0 + max_seq); grow(reads.quals, n_seq0 + max_seq);
  RawVec<size_t> rec_at;     // offset of every record's fixed part
  rec_at.resize(max_reads);
  const ReadGroups* rgp = &hdr.read_groups;   // (filled by the walker before the first run is published)

  auto decode = [&](size_t r) {
    const ReadGroups& rg = *rgp;
    const uint8_t* x = &u[rec_at[r]];
    const int32_t block = rd<int32_t>(x - 4);
    const size_t i = n0 + r;
    const uint8_t l_name = x[8];
    const uint16_t n_cigar = rd<uint16_t>(x + 12);
    const int32_t l_seq = rd<int32_t>(x + 16);
    const uint8_t* q = x + 32 + l_name;
    reads.tid[i] = rd<int32_t>(x); reads.pos[i] = rd<int32_t>(x + 4); reads.flag[i] = rd<uint16_t>(x + 14); reads.mapq[i] = x[9];
    const uint64_t cig_at = reads.cigar_off[i], seq_at = reads.seq_off[i];   // (written by the walker)
    for (int k = 0; k < n_cigar; ++k) reads.cigars[cig_at + (size_t)k] = rd<uint32_t>(q + 4 * k);
    q += 4 * (size_t)n_cigar;
    uint8_t* bases = reads.bases.data() + seq_at;
    for (int b = 0; b < l_seq; ++b) bases[b] = (q[b >> 1] >> ((~b & 1) << 2)) & 0xf;
    q += ((size_t)l_seq + 1) >> 1;
    memcpy(reads.quals.data() + seq_at, q, (size_t)l_seq);
    q += l_seq;
    uint32_t x1 = 1; int32_t xl = -1, xr = -1, as = 0; uint8_t rgi = 0;
    bool seen_x1 = false, seen_xl = false, seen_xr = false, seen_rg = false;  // bam_aux_get returns the FIRST match
    const uint8_t* end = x + block;
    while (q + 3 <= end) {
      uint8_t t0 = q[0], t1 = q[1], type = q[2];
      const uint8_t* v = q + 3;
      size_t sz = aux_value_size(type);
      if (sz) {
        if (v + sz > end) throw std::runtime_error("truncated aux field in BAM record");
        if (t0 == 'X' && t1 == '1' && !seen_x1) { x1 = (uint32_t)aux_int(type, v); seen_x1 = true; }
        else if (t0 == 'X' && t1 == 'L' && !seen_xl) { xl = (int32_t)aux_int(type, v); seen_xl = true; }
        else if (t0 == 'X' && t1 == 'R' && !seen_xr) { xr = (int32_t)aux_int(type, v); seen_xr = true; }
        else if (t0 == 'A' && t1 == 'S') as = (int32_t)aux_int(type, v);
        q = v + sz;
      } else if (type == 'Z' || type == 'H') {
        const uint8_t* z = v;
        while (z < end && *z) ++z;
        if (t0 == 'R' && t1 == 'G' && type == 'Z' && !seen_rg) {
          seen_rg = true;
          if (rg.ids.size() > 1) {  // read_group_index_map::index (alignment.h:545-560)
            std::string id((const char*)v, (size_t)(z - v));
            for (size_t g = 0; g < rg.ids.size(); ++g) if (rg.ids[g] == id) { rgi = (uint8_t)g; break; }
          }
        }
        q = z + 1;
      } else if (type == 'B') {
        if (v + 5 > end) throw std::runtime_error("truncated aux array in BAM record");
        uint8_t st = v[0];
        uint32_t n = rd<uint32_t>(v + 1);
        const size_t es = aux_value_size(st);
        if (!es || (size_t)(end - (v + 5)) / es < n) throw std::runtime_error("truncated aux array in BAM record");
        q = v + 5 + (size_t)n * es;
      } else {
        throw std::runtime_error("unknown aux type in BAM record");
      }
    }
    reads.x1[i] = x1; reads.xl[i] = xl; reads.xr[i] = xr; reads.as[i] = as; reads.rg[i] = rgi;
  };

  std::atomic<size_t> n_found(0);   // records the walker has found so far (the last run may be short)
  auto work = [&]() {
    try {
      for (;;) {
        if (failed) return;
        const size_t k = next_chunk.load() < n_chunks ? next_chunk.fetch_add(1) : n_chunks;
        if (k < n_chunks) {
          for (size_t j = k * CHUNK; j < std::min(blocks.size(), (k + 1) * CHUNK); ++j) inflate_block(in, blocks[j], u.data());
          chunk_done[k].store(1, std::memory_order_release);
          continue;
        }
        const size_t have = runs_published.load(std::memory_order_acquire);
        size_t r = next_run.load();
        if (r < have) {
          if (!next_run.compare_exchange_strong(r, r + 1)) continue;
          const size_t lo = r * RUN, hi = std::min(n_found.load(std::memory_order_acquire), lo + RUN);
          for (size_t i = lo; i < hi; ++i) decode(i);
          continue;
        }
        if (walk_finished.load(std::memory_order_acquire) && next_run.load() >= runs_published.load(std::memory_order_acquire)) return;
        std::this_thread::yield();
      }
    } catch (const std::exception& e) { fail(e.what()); }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; ++t) pool.emplace_back(work);

  // ---- the walker (this thread).  need(p): every byte below p is inflated (it inflates, too, while it has to wait)
  size_t ready = 0, ready_chunk = 0;
  auto need = [&](size_t p) {
    if (p > total) p = total;
    while (ready < p) {
      if (failed) throw std::runtime_error("corrupt BGZF payload in " + path);
      if (ready_chunk < n_chunks && chunk_done[ready_chunk].load(std::memory_order_acquire)) { ready = chunk_end[ready_chunk++]; continue; }
      const size_t k = next_chunk.load() < n_chunks ? next_chunk.fetch_add(1) : n_chunks;
      if (k < n_chunks) {
        try { for (size_t j = k * CHUNK; j < std::min(blocks.size(), (k + 1) * CHUNK); ++j) inflate_block(in, blocks[j], u.data()); }
        catch (const std::exception& e) { fail(e.what()); throw; }
        chunk_done[k].store(1, std::memory_order_release);
      } else {
        std::this_thread::yield();
      }
    }
  };
  size_t n_new = 0;
  uint64_t n_cig_total = n_cig0, n_seq_total = n_seq0;
  try {
    need(12);
    if (total < 12 || memcmp(u.data(), "BAM\1", 4) != 0) throw std::runtime_error(path + " is not a BAM file");
    size_t p = 4;
    const int32_t l_text = rd<int32_t>(&u[p]); p += 4;
    if (l_text < 0 || p + (size_t)l_text + 4 > total) throw std::runtime_error("truncated BAM header");
    need(p + (size_t)l_text + 4);
    hdr.text.assign((const char*)&u[p], (size_t)l_text);
    hdr.text = hdr.text.c_str();  // cut at the first NUL, as a C string consumer would see it
    p += (size_t)l_text;
    const int32_t n_ref = rd<int32_t>(&u[p]); p += 4;
    if (n_ref < 0) throw std::runtime_error("truncated BAM header");
    for (int i = 0; i < n_ref; ++i) {
      need(p + 4);
      if (p + 4 > total) throw std::runtime_error("truncated BAM header");
      const int32_t l_name = rd<int32_t>(&u[p]); p += 4;
      if (l_name < 1 || p + (size_t)l_name + 4 > total) throw std::runtime_error("truncated BAM header");
      need(p + (size_t)l_name + 4);
      hdr.target_names.emplace_back((const char*)&u[p], strnlen((const char*)&u[p], (size_t)l_name));
      p += (size_t)l_name;
      hdr.target_lens.push_back((uint32_t)rd<int32_t>(&u[p])); p += 4;
    }
    parse_rg_lines(hdr.text, hdr.read_groups);
    // a shard's spans (see bam_io.h): which records to keep, and where a sorted file can be left
    std::vector<std::pair<int32_t, int32_t>> spans;
    int32_t last_tid = -1, last_end = 0;
    bool sorted_file = false;
    if (span_of) {
      spans = (*span_of)(hdr);
      spans.resize(hdr.target_names.size(), std::make_pair(0, 0));
      for (size_t t = 0; t < spans.size(); ++t) if (spans[t].first < spans[t].second) { last_tid = (int32_t)t; last_end = spans[t].second; }
      const size_t hd = hdr.text.find("@HD");
      sorted_file = hd != std::string::npos && hdr.text.find("SO:coordinate", hd) < hdr.text.find('\n', hd);
    }
    bool stopped = false;
    // records: one serial walk finds every record and gives it its place in the arrays (prefix sums of the CIGAR and
    // sequence lengths); a run of RUN records is published to the decoders once its last byte is inflated
    while (p + 4 <= total) {
      need(p + 36);
      const int32_t block = rd<int32_t>(&u[p]);
      if (block < 32 || p + 4 + (size_t)block > total) throw std::runtime_error("truncated BAM record");
      const uint8_t* x = &u[p + 4];
      const uint8_t l_name = x[8];
      const uint16_t n_cigar = rd<uint16_t>(x + 12);
      const int32_t l_seq = rd<int32_t>(x + 16);
      if (l_seq < 0 || 32 + (size_t)l_name + 4 * (size_t)n_cigar + (((size_t)l_seq + 1) >> 1) + (size_t)l_seq > (size_t)block)
        throw std::runtime_error("truncated BAM record");
      if (span_of) {
        const int32_t tid = rd<int32_t>(x), pos = rd<int32_t>(x + 4);
        // (targets come in header order in a sorted file, unplaced reads last)
        if (sorted_file && (tid < 0 || tid > last_tid || (tid == last_tid && pos >= last_end))) { stopped = true; break; }
        bool keep = tid >= 0 && (size_t)tid < spans.size() && spans[(size_t)tid].first < spans[(size_t)tid].second && pos < spans[(size_t)tid].second;
        if (keep) {
          need(p + 4 + 32 + (size_t)l_name + 4 * (size_t)n_cigar);
          int64_t end = pos;
          const uint8_t* cg = x + 32 + l_name;
          for (int k = 0; k < n_cigar; ++k) { const uint32_t c = rd<uint32_t>(cg + 4 * k), op = c & 15u; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) end += c >> 4; }
          if (end == pos) end = pos + 1;
          keep = end > spans[(size_t)tid].first;
        }
        if (!keep) { p += 4 + (size_t)block; continue; }
      }
      if (n_new >= max_reads) throw std::runtime_error("truncated BAM record");
      rec_at[n_new] = p + 4;
      reads.n_cigar[n0 + n_new] = n_cigar; reads.cigar_off[n0 + n_new] = n_cig_total;
      reads.l_seq[n0 + n_new] = (uint32_t)l_seq; reads.seq_off[n0 + n_new] = n_seq_total;
      n_cig_total += n_cigar; n_seq_total += (uint64_t)l_seq;
      p += 4 + (size_t)block;
      ++n_new;
      if (n_new % RUN == 0) { need(p); n_found.store(n_new, std::memory_order_release); runs_published.store(n_new / RUN, std::memory_order_release); }
    }
    if (stopped) { next_chunk.store(n_chunks); need(std::min(p, total)); }   // the members behind the shard are not inflated
    else need(total);
    n_found.store(n_new, std::memory_order_release);
    runs_published.store((n_new + RUN - 1) / RUN, std::memory_order_release);
  } catch (const std::exception& e) { fail(e.what()); }
  walk_finished.store(true, std::memory_order_release);
  work();   // the walker decodes, too, now
  for (auto& t : pool) t.join();
  if (failed) throw std::runtime_error(err.empty() ? "corrupt BGZF payload in " + path : err);
  const size_t n_all = n0 + n_new;
  reads.tid.resize(n_all); reads.pos.resize(n_all); reads.flag.resize(n_all); reads.mapq.resize(n_all);
  reads.n_cigar.resize(n_all); reads.cigar_off.resize(n_all); reads.l_seq.resize(n_all); reads.seq_off.resize(n_all);
  reads.x1.resize(n_all); reads.xl.resize(n_all); reads.xr.resize(n_all); reads.as.resize(n_all); reads.rg.resize(n_all);
  reads.cigars.resize(n_cig_total); reads.bases.resize(n_seq_total); reads.quals.resize(n_seq_total);
  phase_done("inflate, find and decode records");
}

// ---------------------------------------------------------------------------------- writers

namespace {

uint16_t reg2bin(int64_t beg, int64_t end) {
  --end;
  if (beg >> 14 == end >> 14) return (uint16_t)(((1 << 15) - 1) / 7 + (beg >> 14));
  if (beg >> 17 == end >> 17) return (uint16_t)(((1 << 12) - 1) / 7 + (beg >> 17));
  if (beg >> 20 == end >> 20) return (uint16_t)(((1 << 9) - 1) / 7 + (beg >> 20));
  if (beg >> 23 == end >> 23) return (uint16_t)(((1 << 6) - 1) / 7 + (beg >> 23));
  if (beg >> 26 == end >> 26) return (uint16_t)(((1 << 3) - 1) / 7 + (beg >> 26));
  return 0;
}

}  // namespace

// One BAM record of `r` appended to `rec` (block_size prefix included).
static void append_record(std::vector<uint8_t>& rec, const BamHeader& hdr, const ReadBatch& r, size_t i) {
  const size_t start = rec.size();
  std::string name = i < r.names.size() ? r.names[i] : ("r" + std::to_string(i));
  auto put = [&](const void* p, size_t n) { rec.insert(rec.end(), (const uint8_t*)p, (const uint8_t*)p + n); };
  auto put32 = [&](int32_t v) { put(&v, 4); };
  auto put16 = [&](uint16_t v) { put(&v, 2); };
  put32(0);  // block_size, patched below
  const uint32_t* cig = &r.cigars[r.cigar_off[i]];
  int64_t rlen = 0;
  for (uint32_t k = 0; k < r.n_cigar[i]; ++k) {
    uint32_t op = cig[k] & 0xf;
    if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += cig[k] >> 4;
  }
  put32(r.tid[i]); put32(r.pos[i]);
  rec.push_back((uint8_t)(name.size() + 1)); rec.push_back(r.mapq[i]);
  put16(reg2bin(r.pos[i], r.pos[i] + (rlen ? rlen : 1)));
  put16((uint16_t)r.n_cigar[i]); put16(r.flag[i]);
  put32((int32_t)r.l_seq[i]); put32(-1); put32(-1); put32(0);
  put(name.c_str(), name.size() + 1);
  put(cig, 4 * (size_t)r.n_cigar[i]);
  const uint8_t* b = &r.bases[r.seq_off[i]];
  for (uint32_t j = 0; j < r.l_seq[i]; j += 2) {
    uint8_t hi = b[j], lo = (j + 1 < r.l_seq[i]) ? b[j + 1] : 0;
    rec.push_back((uint8_t)((hi << 4) | lo));
  }
  put(&r.quals[r.seq_off[i]], r.l_seq[i]);
  auto tag_int = [&](const char* t, int64_t v) {
    rec.push_back((uint8_t)t[0]); rec.push_back((uint8_t)t[1]);
    if (v >= 0 && v < 256) { rec.push_back('C'); rec.push_back((uint8_t)v); }
    else { rec.push_back('i'); int32_t x = (int32_t)v; put(&x, 4); }
  };
  tag_int("AS", r.as[i]);
  if (r.x1[i] != 1 || (i % 3) != 0) tag_int("X1", r.x1[i]);  // leave the tag off some unique reads: absent == 1
  if (!hdr.read_groups.ids.empty()) {
    const std::string& id = hdr.read_groups.ids[r.rg[i]];
    rec.push_back('R'); rec.push_back('G'); rec.push_back('Z');
    put(id.c_str(), id.size() + 1);
  }
  if (r.xl[i] >= 0) tag_int("XL", r.xl[i]);
  if (r.xr[i] >= 0) tag_int("XR", r.xr[i]);
  const int32_t block = (int32_t)(rec.size() - start - 4);
  memcpy(&rec[start], &block, 4);
}

// The file is the same bytes whatever the thread count: records are serialised over contiguous read ranges, the
// uncompressed image is cut into members of 0xff00 bytes, and the members are deflated side by side.
void write_bam(const std::string& path, const BamHeader& hdr, const ReadBatch& r, int level, int threads) {
  if (threads < 1) threads = (int)std::max(1u, std::thread::hardware_concurrency());
  std::vector<uint8_t> head;
  {
    auto put = [&](const void* p, size_t n) { head.insert(head.end(), (const uint8_t*)p, (const uint8_t*)p + n); };
    auto put32 = [&](int32_t v) { put(&v, 4); };
    put("BAM\1", 4);
    put32((int32_t)hdr.text.size());
    put(hdr.text.data(), hdr.text.size());
    put32((int32_t)hdr.target_names.size());
    for (size_t i = 0; i < hdr.target_names.size(); ++i) {
      put32((int32_t)hdr.target_names[i].size() + 1);
      put(hdr.target_names[i].c_str(), hdr.target_names[i].size() + 1);
      put32((int32_t)hdr.target_lens[i]);
    }
  }
  const size_t n = r.size(), n_parts = (size_t)threads * 4;
  std::vector<std::vector<uint8_t>> parts(n_parts);
  auto run = [&](size_t n_jobs, auto&& body) {
    std::atomic<size_t> next(0);
    auto work = [&]() { for (;;) { const size_t k = next.fetch_add(1); if (k >= n_jobs) break; body(k); } };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
  };
  run(n_parts, [&](size_t k) {
    std::vector<uint8_t>& out = parts[k];
    const size_t lo = n * k / n_parts, hi = n * (k + 1) / n_parts;
    for (size_t i = lo; i < hi; ++i) append_record(out, hdr, r, i);
  });
  std::vector<size_t> part_at(n_parts + 1, head.size());
  for (size_t k = 0; k < n_parts; ++k) part_at[k + 1] = part_at[k] + parts[k].size();
  const size_t total = part_at[n_parts];
  std::vector<uint8_t> u(total);
  memcpy(u.data(), head.data(), head.size());
  run(n_parts, [&](size_t k) { if (!parts[k].empty()) memcpy(&u[part_at[k]], parts[k].data(), parts[k].size()); std::vector<uint8_t>().swap(parts[k]); });
  const size_t member = 0xff00, n_members = (total + member - 1) / member;
  std::vector<std::vector<uint8_t>> packed(n_members + 1);
  auto deflate_member = [&](const uint8_t* src, size_t len, std::vector<uint8_t>& out) {
    out.resize(0x10000);
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
    zs.next_in = const_cast<Bytef*>(src);
    zs.avail_in = (uInt)len;
    zs.next_out = out.data() + 18;
    zs.avail_out = (uInt)(out.size() - 18 - 8);
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { deflateEnd(&zs); throw std::runtime_error("deflate overflow"); }
    const size_t clen = zs.total_out;
    deflateEnd(&zs);
    static const uint8_t h12[12] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0};
    memcpy(out.data(), h12, 12);
    out[12] = 'B'; out[13] = 'C'; out[14] = 2; out[15] = 0;
    const uint16_t bsize = (uint16_t)(clen + 18 + 8 - 1);
    memcpy(out.data() + 16, &bsize, 2);
    const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), src, (uInt)len), isize = (uint32_t)len;
    memcpy(out.data() + 18 + clen, &crc, 4);
    memcpy(out.data() + 18 + clen + 4, &isize, 4);
    out.resize(18 + clen + 8);
  };
  run(n_members + 1, [&](size_t k) {
    if (k == n_members) deflate_member(u.data(), 0, packed[k]);  // empty member = BGZF EOF marker
    else deflate_member(&u[k * member], std::min(member, total - k * member), packed[k]);
  });
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) throw std::runtime_error("cannot create " + path);
  for (const auto& m : packed) if (fwrite(m.data(), 1, m.size(), f) != m.size()) { fclose(f); throw std::runtime_error("short write on " + path); }
  fclose(f);
}

void read_fasta(const std::string& path, RefSet& ref) {
  FILE* f = fopen(path.c_str(), "r");
  if (!f) throw std::runtime_error("cannot open " + path);
  char* line = nullptr;
  size_t cap = 0;
  ssize_t n;
  while ((n = getline(&line, &cap, f)) >= 0) {
    while (n > 0 && (line[n - 1] == '\n' || line[n - 1] == '\r')) line[--n] = 0;
    if (line[0] == '>') {
      std::string name(line + 1);
      size_t sp = name.find_first_of(" \t");
      if (sp != std::string::npos) name.resize(sp);
      ref.names.push_back(name);
      ref.seqs.emplace_back();
    } else if (!ref.seqs.empty()) {
      ref.seqs.back().append(line, (size_t)n);
    }
  }
  free(line);
  fclose(f);
}

void write_fasta(const std::string& path, const RefSet& ref, bool with_fai) {
  FILE* f = fopen(path.c_str(), "w");
  if (!f) throw std::runtime_error("cannot create " + path);
  FILE* fai = with_fai ? fopen((path + ".fai").c_str(), "w") : nullptr;
  long off = 0;
  const size_t width = 60;
  for (size_t i = 0; i < ref.names.size(); ++i) {
    off += fprintf(f, ">%s\n", ref.names[i].c_str());
    if (fai) fprintf(fai, "%s\t%zu\t%ld\t%zu\t%zu\n", ref.names[i].c_str(), ref.seqs[i].size(), off, width, width + 1);
    for (size_t p = 0; p < ref.seqs[i].size(); p += width) {
      size_t k = std::min(width, ref.seqs[i].size() - p);
      fwrite(ref.seqs[i].data() + p, 1, k, f);
      fputc('\n', f);
      off += (long)k + 1;
    }
  }
  fclose(f);
  if (fai) fclose(fai);
}

}  // namespace brq
