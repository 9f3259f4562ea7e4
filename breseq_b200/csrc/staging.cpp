#include "staging.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>
#include <thread>

namespace brq {

void make_read_file_partition(const ReadGroups& rg, const std::vector<ReadFileSetInfo>& sets,
                              std::vector<uint32_t>& base, std::vector<uint32_t>& count) {
  base.clear(); count.clear();
  if (sets.empty()) return;
  std::vector<uint32_t> set_base, set_count;
  uint32_t flat = 0;
  for (const ReadFileSetInfo& s : sets) { set_base.push_back(flat); set_count.push_back(s.n_files); flat += s.n_files; }
  for (size_t g = 0; g < rg.ids.size(); ++g) {
    size_t match = 0; bool found = false;
    if (!rg.libraries[g].empty())
      for (size_t s = 0; s < sets.size(); ++s) if (sets[s].base_name == rg.libraries[g]) { match = s; found = true; break; }
    if (!found) { match = g; found = g < set_base.size(); }
    if (found && match < set_base.size()) { base.push_back(set_base[match]); count.push_back(set_count[match]); }
    else { base.push_back(0); count.push_back(1); }
  }
}

namespace {

void* default_alloc(size_t bytes, bool* pinned) {
  *pinned = false;
  void* p = nullptr;
  if (posix_memalign(&p, 256, bytes ? bytes : 256) != 0) throw std::bad_alloc();
  return p;
}
void default_release(void* p, bool) { free(p); }

inline bool op_ref(uint32_t op) { return op == 0 || op == 2 || op == 3 || op == 7 || op == 8; }
inline bool op_match(uint32_t op) { return op == 0 || op == 7 || op == 8; }

struct ReadInfo {
  uint32_t L;          // l_seq
  int32_t end;         // exclusive reference end
  int32_t qs0, qe0;    // first/last non-soft-clipped query index (alignment.cpp:248-288)
  int32_t qb_end0;     // query_bounds_0 end (alignment.cpp:104-218, min_qual == 0)
  int32_t qb_start0;   // query_bounds_0 start: soft clips among the leading S / H / N operations
  uint8_t read_set;
  bool rev;
};

// Visit the pileup entries of one read restricted to columns [lo, hi): f(c, q, is_del, indel).
// Same per-column values htslib's pileup reports (qpos on a deleted column = first query base
// after the deletion; indel only on the last column before an I / D run, runs merged, P skipped).
template <class F>
inline void walk_read(const uint32_t* cig, uint32_t n_cig, int32_t pos, int32_t lo, int32_t hi, F&& f) {
  int32_t x = pos, y = 0;
  for (uint32_t k = 0; k < n_cig; ++k) {
    uint32_t op = cig[k] & 0xf;
    int32_t l = (int32_t)(cig[k] >> 4);
    if (op_ref(op)) {
      if (x >= hi) return;
      int32_t xe = x + l;
      if (xe > lo) {
        // indel reported at the last column of this op
        int indel = 0;
        if (k + 1 < n_cig) {
          uint32_t op2 = cig[k + 1] & 0xf;
          int32_t l2 = (int32_t)(cig[k + 1] >> 4);
          if (op2 == 2 && op != 2) {
            indel = -l2;
            for (uint32_t j = k + 2; j < n_cig && (cig[j] & 0xf) == 2; ++j) indel -= (int32_t)(cig[j] >> 4);
          } else if (op2 == 1) {
            indel = l2;
            for (uint32_t j = k + 2; j < n_cig; ++j) {
              uint32_t o = cig[j] & 0xf;
              if (o == 1) indel += (int32_t)(cig[j] >> 4);
              else if (o != 6) break;
            }
          } else if (op2 == 6 && k + 2 < n_cig) {
            int32_t l3 = 0;
            for (uint32_t j = k + 2; j < n_cig; ++j) {
              uint32_t o = cig[j] & 0xf;
              if (o == 1) l3 += (int32_t)(cig[j] >> 4);
              else if (op_ref(o)) break;
            }
            if (l3 > 0) indel = l3;
          }
        }
        int32_t c0 = std::max(x, lo), c1 = std::min(xe, hi);
        if (op_match(op)) {
          for (int32_t c = c0; c < c1; ++c) f(c, y + (c - x), false, (c == xe - 1) ? indel : 0);
        } else {
          for (int32_t c = c0; c < c1; ++c) f(c, y, true, (c == xe - 1) ? indel : 0);
        }
      }
      if (op_match(op)) y += l;
      x = xe;
    } else if (op == 1 || op == 4) {
      y += l;
    }
  }
}

struct Item { uint32_t v; int32_t lo, hi; size_t first_read, last_read; };
inline int32_t tlen_of(const std::string& s) { return (int32_t)s.size(); }

}  // namespace


// Targets in visit order (std::set<string> iteration == sorted strings; pileup_base.cpp:364-385), clipped to this
// process's contiguous coordinate shard: fills out.segments / out.n_base and the reference sequence of every segment.
void plan_segments(const BamHeader& hdr, const RefSet& ref, const StageConfig& cfg, PileupStream& out, std::vector<const std::string*>& refseq) {
  const size_t n_targets = hdr.target_names.size();
  std::vector<std::string> ids = cfg.call_seq_ids.empty() ? hdr.target_names : cfg.call_seq_ids;
  std::sort(ids.begin(), ids.end());
  ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
  refseq.clear();
  std::vector<Segment> full;
  uint64_t total_cols = 0;
  for (const std::string& id : ids) {
    size_t tid = std::find(hdr.target_names.begin(), hdr.target_names.end(), id) - hdr.target_names.begin();
    if (tid == n_targets) throw std::runtime_error("Could not find seq_id: " + id);
    size_t r = std::find(ref.names.begin(), ref.names.end(), id) - ref.names.begin();
    if (r == ref.names.size() || ref.seqs[r].size() != hdr.target_lens[tid])
      throw std::runtime_error("reference sequence missing or of the wrong length: " + id);
    full.push_back({(int32_t)tid, 0, (int32_t)hdr.target_lens[tid], total_cols});
    total_cols += hdr.target_lens[tid];
    refseq.push_back(&ref.seqs[r]);
    // coverage groups of the whole run, not of this shard: the ranks of a sharded run sum their coverage histograms
    const uint32_t g = cfg.coverage_group_of_tid.empty() ? (uint32_t)tid : cfg.coverage_group_of_tid[tid];
    if (g > 255) throw std::runtime_error("more than 256 coverage groups are not supported");
    if (g + 1 > out.n_groups) out.n_groups = g + 1;
  }
  // this process's shard: [g_lo, g_hi) of the concatenated visit-order columns.  Explicit bounds (StageConfig::shard_lo /
  // shard_hi, set by a caller that balances the shards by record count) win over the even split by columns.
  const uint64_t n_sh = std::max<uint32_t>(1, cfg.shard_count), rk = std::min<uint64_t>(cfg.shard_rank, n_sh - 1);
  uint64_t g_lo = total_cols * rk / n_sh, g_hi = total_cols * (rk + 1) / n_sh;
  if (cfg.shard_hi > cfg.shard_lo || cfg.shard_explicit) { g_lo = std::min<uint64_t>(cfg.shard_lo, total_cols); g_hi = std::min<uint64_t>(cfg.shard_hi, total_cols); }
  out.visit_targets = full;
  std::vector<const std::string*> kept;
  for (size_t v = 0; v < full.size(); ++v) {
    const uint64_t a = full[v].slot0, b = a + (uint64_t)full[v].hi;
    const uint64_t lo = std::max(a, g_lo), hi = std::min(b, g_hi);
    if (lo >= hi) continue;
    out.segments.push_back({full[v].tid, (int32_t)(lo - a), (int32_t)(hi - a), out.n_base});
    out.n_base += hi - lo;
    kept.push_back(refseq[v]);
  }
  refseq.swap(kept);
}

std::vector<UserRa> read_user_evidence_gd(const std::string& path) {
  FILE* f = fopen(path.c_str(), "r");
  if (!f) throw std::runtime_error("cannot open " + path);
  std::vector<UserRa> list;
  char* line = nullptr;
  size_t cap = 0;
  ssize_t n;
  while ((n = getline(&line, &cap, f)) >= 0) {
    while (n > 0 && (line[n - 1] == '\n' || line[n - 1] == '\r')) line[--n] = 0;
    std::vector<std::string> fld;
    for (char* p = line;;) { char* t = strchr(p, '\t'); fld.emplace_back(p, t ? (size_t)(t - p) : strlen(p)); if (!t) break; p = t + 1; }
    if (fld.size() < 8 || fld[0] != "RA") continue;   // type, id, parents, then the specification: seq_id position insert_position ref_base new_base
    UserRa e;
    e.seq_id = fld[3]; e.position = (uint32_t)strtoul(fld[4].c_str(), nullptr, 10); e.insert_position = (uint32_t)strtoul(fld[5].c_str(), nullptr, 10);
    e.ref_base = fld[6]; e.new_base = fld[7];
    list.push_back(e);
  }
  free(line);
  fclose(f);
  // cDiffEntry::compare for two RA rows: seq_id (as a string), position, then the specification fields in order
  std::stable_sort(list.begin(), list.end(), [](const UserRa& a, const UserRa& b) {
    if (a.seq_id != b.seq_id) return a.seq_id < b.seq_id;
    if (a.position != b.position) return a.position < b.position;
    if (a.insert_position != b.insert_position) return a.insert_position < b.insert_position;
    if (a.ref_base != b.ref_base) return a.ref_base < b.ref_base;
    return a.new_base < b.new_base;
  });
  return list;
}

std::vector<UserColumn> plan_user_evidence(const std::vector<UserRa>& list, const BamHeader& hdr, const std::vector<Segment>& visit_full,
                                           const std::vector<double>& skip_cutoff,
                                           const std::function<uint32_t(size_t, uint32_t, uint32_t)>& levels) {
  std::vector<UserColumn> out;
  size_t front = 0;
  for (size_t v = 0; v < visit_full.size() && front < list.size(); ++v) {
    const int32_t tid = visit_full[v].tid;
    if (!skip_cutoff.empty() && skip_cutoff[(size_t)tid] < 0.0) continue;   // the callback returns before it looks at the list (identify_mutations.cpp:1319-1336)
    const std::string& name = hdr.target_names[(size_t)tid];
    const uint32_t len = hdr.target_lens[(size_t)tid];
    uint32_t done = 0;   // columns of this target already passed
    while (front < list.size()) {
      const uint32_t p = list[front].position;
      if (p <= done || p > len) break;   // no later column of this target has the front entry's position
      UserColumn c;
      c.tid = tid; c.pos1 = p; c.slot = ~0ull;
      for (size_t u = front; u < list.size() && list[u].position == p; ++u) c.force_max = list[u].insert_position;
      const uint32_t top = levels(v, p, c.force_max);
      for (uint32_t k = 0; k <= top; ++k)
        while (front < list.size() && list[front].seq_id == name && list[front].position == p && list[front].insert_position == k) {
          c.consumed.emplace_back(k, (uint32_t)front);
          ++front;
        }
      out.push_back(c);
      done = p;
    }
  }
  return out;
}

// Table geometry of the device stream words (ScoreGeometry) from per-read statistics: mq[m] = bases of unique reads with
// MAPQ m, qc[q] = bases of unique reads with quality q.
ScoreGeometry choose_geometry(const uint64_t* mq, const uint64_t* qc, const StageConfig& cfg, uint32_t max_read_set_seen) {
  ScoreGeometry g;
  g.cutoff = cfg.base_quality_cutoff;
  g.n_st = (max_read_set_seen + 1) * 2;
  g.hot_mapq = 0;
  for (uint32_t m = 1; m < 256; ++m) if (mq[m] > mq[g.hot_mapq]) g.hot_mapq = m;
  // The per-slot class histogram of the tally kernel holds (read set, strand, quality) classes, at most 62
  // four-byte words per lane, and its contraction with the likelihood table walks every word: the table covers
  // the narrowest window of quality values (a multiple of four) that holds 99.5 % of the records; the few
  // records outside it take the side list.
  uint32_t q_first = 128, q_last = 0;
  uint64_t q_mass = 0;
  for (uint32_t q = g.cutoff; q < 128; ++q) if (qc[q]) { q_first = std::min(q_first, q); q_last = q; q_mass += qc[q]; }
  if (q_first > q_last) { q_first = g.cutoff; q_last = g.cutoff; }
  const uint32_t span = q_last - q_first + 1, nq_cap = (248u / g.n_st) & ~3u;
  uint32_t nq = 0, best_lo = q_first;
  for (uint32_t len = 4; nq == 0; len += 4) {
    uint64_t best = 0;
    uint32_t lo_best = q_first;
    for (uint32_t lo = q_first; lo == q_first || lo + len <= q_last + 1; ++lo) {
      uint64_t mass = 0;
      for (uint32_t q = lo; q < lo + len && q < 128; ++q) mass += qc[q];
      if (mass > best) { best = mass; lo_best = lo; }
    }
    if (len >= span || len + 4 > nq_cap || (double)best >= 0.995 * (double)q_mass) { nq = len; best_lo = lo_best; }
  }
  if (nq > nq_cap) nq = nq_cap;  // more than 62 read files x strands: no window fits, everything takes the side list
  g.q_lo = best_lo; g.n_q = nq;
  if (cfg.use_read_pos || cfg.use_base_repeat) { g.n_q = 0; g.side_stride = 2; }  // classes too many for a shared table: all cold
  return g;
}

void free_stream(PileupStream& s, const StageConfig& cfg) {
  auto rel = cfg.release ? cfg.release : default_release;
  for (void* p : {(void*)s.slot_ref, (void*)s.score_off, (void*)s.hist_off, (void*)s.slot_group, (void*)s.score_rec, (void*)s.hist_rec, (void*)s.side_rec, (void*)s.side_off, (void*)s.round_slot, (void*)s.score_cnt, (void*)s.round_off, (void*)s.round_side, (void*)s.hist16, (void*)s.hist_exc, (void*)s.score16, (void*)s.score_exc, (void*)s.score_exc_off})
    if (p) rel(p, s.pinned && !(p == (void*)s.score_rec && s.score_rec_plain) && !(p == (void*)s.hist_rec && s.hist_rec_plain));
  s = PileupStream();
}

void stage(const BamHeader& hdr, const RefSet& ref, const ReadBatch& R, const StageConfig& cfg, PileupStream& out) {
  auto alloc = cfg.alloc ? cfg.alloc : default_alloc;
  // BRQ_STAGE_TIMES=1: wall time of every phase on stderr (staging is upstream of the measured path)
  static const bool phase_times = getenv("BRQ_STAGE_TIMES") != nullptr;
  auto phase_t0 = std::chrono::steady_clock::now();
  auto phase_done = [&](const char* what) {
    if (!phase_times) return;
    const auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "stage: %-22s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t - phase_t0).count());
    phase_t0 = t;
  };
  const size_t n_reads = R.size();
  const size_t n_targets = hdr.target_names.size();
  out = PileupStream();

  std::vector<const std::string*> refseq;
  plan_segments(hdr, ref, cfg, out, refseq);
  const size_t n_visit = out.segments.size();

  // ---- read ranges per target; the BAM must be coordinate sorted (htslib's pileup aborts otherwise)
  std::vector<size_t> t_first(n_targets, 0), t_last(n_targets, 0);
  {
    int32_t last_tid = -1, last_pos = -1;
    for (size_t i = 0; i < n_reads; ++i) {
      int32_t t = R.tid[i];
      if (t < 0) { last_tid = INT32_MAX; continue; }   // unplaced reads sort last
      if (t < last_tid || (t == last_tid && R.pos[i] < last_pos)) throw std::runtime_error("BAM is not coordinate sorted");
      if (t != last_tid) { t_first[(size_t)t] = i; }
      t_last[(size_t)t] = i + 1;
      last_tid = t; last_pos = R.pos[i];
    }
  }

  phase_done("setup");
  // ---- per-read derived values
  std::vector<uint32_t> part_base, part_count;
  make_read_file_partition(hdr.read_groups, cfg.read_file_sets, part_base, part_count);
  std::vector<ReadInfo> info(n_reads);
  std::vector<int32_t> max_span(n_targets, 1);
  for (size_t i = 0; i < n_reads; ++i) {
    ReadInfo& ri = info[i];
    const uint32_t* cig = R.cigars.data() + R.cigar_off[i];
    uint32_t nc = R.n_cigar[i];
    ri.L = R.l_seq[i];
    ri.rev = (R.flag[i] & 16) != 0;
    int32_t rlen = 0, qlen = 0;
    for (uint32_t k = 0; k < nc; ++k) {
      uint32_t op = cig[k] & 0xf; int32_t l = (int32_t)(cig[k] >> 4);
      if (op_ref(op)) rlen += l;
      if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) qlen += l;
    }
    ri.end = R.pos[i] + (rlen ? rlen : 1);
    int32_t qs1 = 1;
    for (uint32_t k = 0; k < nc && (cig[k] & 0xf) == 4; ++k) qs1 += (int32_t)(cig[k] >> 4);
    int32_t qe1 = qlen;
    for (uint32_t k = nc; k-- > 1 && (cig[k] & 0xf) == 4;) qe1 -= (int32_t)(cig[k] >> 4);
    ri.qs0 = qs1 - 1; ri.qe0 = qe1 - 1;
    int32_t be1 = qlen;
    for (uint32_t k = nc; k-- > 1;) {
      uint32_t op = cig[k] & 0xf;
      if (op != 4 && op != 5 && op != 3) break;
      if (op == 4) be1 -= (int32_t)(cig[k] >> 4);
    }
    ri.qb_end0 = be1 - 1;
    int32_t bs1 = 1;
    for (uint32_t k = 0; k < nc; ++k) {
      uint32_t op = cig[k] & 0xf;
      if (op != 4 && op != 5 && op != 3) break;
      if (op == 4) bs1 += (int32_t)(cig[k] >> 4);
    }
    ri.qb_start0 = bs1 - 1;
    uint32_t g = R.rg[i];
    ri.read_set = 0;
    if (!part_base.empty() && g < part_base.size())
      ri.read_set = (uint8_t)(part_base[g] + (((R.flag[i] & 128) && part_count[g] > 1) ? 1 : 0));
    if (ri.read_set >= 32) throw std::runtime_error("more than 32 read files are not supported by the packed record");
    if (ri.read_set > out.max_read_set_seen) out.max_read_set_seen = ri.read_set;
    if (R.tid[i] >= 0 && rlen > max_span[(size_t)R.tid[i]]) max_span[(size_t)R.tid[i]] = rlen;
  }
  auto in_pileup = [&](size_t i) {  // bam_plp_push keeps mapped reads with a tid; a read without a walkable CIGAR cannot be resolved
    return R.tid[i] >= 0 && !(R.flag[i] & 4) && R.n_cigar[i] > 0;
  };

  // ---- work items: column ranges of visited targets
  std::vector<Item> items;
  for (size_t v = 0; v < n_visit; ++v) {
    const Segment& sg = out.segments[v];
    size_t tid = (size_t)sg.tid;
    int32_t len = (int32_t)hdr.target_lens[tid];
    size_t nr = t_last[tid] - t_first[tid];
    int32_t chunk = 8192;
    if (nr) { double depth = (double)nr * 100.0 / std::max(1, len); if (depth > 400) chunk = 2048; }
    size_t cursor = t_first[tid];
    for (int32_t lo = sg.lo; lo < sg.hi; lo += chunk) {
      Item it; it.v = (uint32_t)v; it.lo = lo; it.hi = std::min(sg.hi, lo + chunk);
      // reads are sorted by pos: first candidate has pos > lo - max_span
      while (cursor < t_last[tid] && R.pos[cursor] + max_span[tid] <= lo) ++cursor;
      it.first_read = cursor;
      size_t e = cursor;
      while (e < t_last[tid] && R.pos[e] < it.hi) ++e;
      it.last_read = e;
      items.push_back(it);
    }
  }
  const int n_threads = std::max(1, cfg.threads);
  std::string error;
  std::mutex error_mu;
  auto run_items = [&](auto&& body) {
    std::atomic<size_t> next(0);
    auto work = [&]() {
      try {
        for (;;) { size_t i = next.fetch_add(1); if (i >= items.size()) break; body(i); }
      } catch (const std::exception& e) { std::lock_guard<std::mutex> g(error_mu); if (error.empty()) error = e.what(); next = items.size(); }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    if (!error.empty()) throw std::runtime_error(error);
  };
  // [0, n) cut into parts handed out to the threads: body(first, last)
  auto parallel_ranges = [&](uint64_t n, auto&& body) {
    const size_t n_parts = (size_t)n_threads * 8;
    std::atomic<size_t> next(0);
    auto work = [&]() { for (;;) { const size_t k = next.fetch_add(1); if (k >= n_parts) break; body(n * k / n_parts, n * (k + 1) / n_parts); } };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
  };

  phase_done("per-read values");
  // ---- pass A1: which insert sub-columns exist.  Level k+1 exists iff a UNIQUE read has an
  // insertion longer than k after the column and a non-N base at level k
  // (identify_mutations.cpp:1577 precedes :1598).
  struct InsSupport { uint64_t slot; uint64_t mask; };
  std::vector<std::vector<InsSupport>> support(items.size());
  run_items([&](size_t ii) {
    const Item& it = items[ii];
    std::vector<InsSupport>& sup = support[ii];
    for (size_t i = it.first_read; i < it.last_read; ++i) {
      if (!in_pileup(i) || info[i].end <= it.lo || R.x1[i] != 1) continue;
      const uint32_t* cig = R.cigars.data() + R.cigar_off[i];
      bool any = false;
      for (uint32_t k = 0; k < R.n_cigar[i]; ++k) if ((cig[k] & 0xf) == 1) { any = true; break; }
      if (!any) continue;
      const uint8_t* seq = R.bases.data() + R.seq_off[i];
      walk_read(cig, R.n_cigar[i], R.pos[i], it.lo, it.hi, [&](int32_t c, int32_t q, bool is_del, int indel) {
        if (is_del || indel <= 0) return;
        if (indel > 63) throw std::runtime_error("insertions longer than 63 bases are not supported");
        uint64_t mask = 0;
        for (int k = 0; k < indel; ++k) if (seq[q + k] != 15) mask |= 1ull << k;
        sup.push_back({out.segments[it.v].slot0 + (uint64_t)(c - out.segments[it.v].lo), mask});
      });
    }
  });
  std::vector<uint32_t> sub_first;   // per base slot: first sub-slot index or ~0u
  std::vector<uint8_t> sub_k;        // per base slot: number of sub-columns
  {
    std::vector<InsSupport> all;
    for (auto& v : support) all.insert(all.end(), v.begin(), v.end());
    support.clear();
    std::sort(all.begin(), all.end(), [](const InsSupport& a, const InsSupport& b) { return a.slot < b.slot; });
    sub_first.assign(out.n_base, 0xFFFFFFFFu);
    sub_k.assign(out.n_base, 0);
    std::map<uint64_t, uint64_t> mask_of;   // base slot -> which insert levels the reads support
    for (size_t a = 0; a < all.size();) {
      size_t b = a; uint64_t mask = 0;
      while (b < all.size() && all[b].slot == all[a].slot) mask |= all[b++].mask;
      if (mask) mask_of[all[a].slot] = mask;
      a = b;
    }
    // user evidence: the columns where the pileup meets the list, the levels it forces there (identify_mutations.cpp:1346-1355)
    std::map<uint64_t, uint32_t> forced;
    if (!cfg.user_evidence.empty()) {
      out.user_list = cfg.user_evidence;
      auto slot_of = [&](size_t v, uint32_t pos1) -> uint64_t {  // base slot of a column of visited target v, ~0 outside this shard
        for (const Segment& sg : out.segments)
          if (sg.tid == out.visit_targets[v].tid && (int32_t)pos1 - 1 >= sg.lo && (int32_t)pos1 - 1 < sg.hi) return sg.slot0 + (uint64_t)((int32_t)pos1 - 1 - sg.lo);
        return ~0ull;
      };
      out.user_columns = plan_user_evidence(cfg.user_evidence, hdr, out.visit_targets, cfg.user_skip_cutoff, [&](size_t v, uint32_t pos1, uint32_t force_max) {
        const uint64_t slot = slot_of(v, pos1);
        if (slot == ~0ull) return force_max;   // another shard's column: its read support is not known here
        const auto m = mask_of.find(slot);
        return insert_levels(m == mask_of.end() ? 0 : m->second, force_max);
      });
      for (UserColumn& c : out.user_columns) {
        size_t v = 0;
        while (v < out.visit_targets.size() && out.visit_targets[v].tid != c.tid) ++v;
        c.slot = slot_of(v, c.pos1);
        if (c.slot != ~0ull && c.force_max) forced[c.slot] = std::max(forced[c.slot], c.force_max);
      }
    }
    for (const auto& f : forced) if (!mask_of.count(f.first)) mask_of[f.first] = 0;
    for (const auto& m : mask_of) {
      const auto f = forced.find(m.first);
      const uint32_t K = insert_levels(m.second, f == forced.end() ? 0 : f->second);
      if (K) {
        sub_first[m.first] = (uint32_t)out.ins_parent.size();
        sub_k[m.first] = (uint8_t)K;
        for (uint32_t k = 1; k <= K; ++k) { out.ins_parent.push_back(m.first); out.ins_count.push_back(k); }
      }
    }
    out.n_ins = out.ins_parent.size();
  }
  const uint64_t n_slots = out.n_slots();

  phase_done("pass A1");
  // ---- per-slot reference bases and coverage groups; offset arrays
  bool pinned = false, p2 = false;
  out.slot_ref = (uint8_t*)alloc(n_slots, &pinned);
  out.slot_group = (uint8_t*)alloc(out.n_base ? out.n_base : 1, &p2);
  out.score_off = (uint64_t*)alloc((n_slots + 1) * 8, &p2);
  out.hist_off = (uint64_t*)alloc((out.n_base + 1) * 8, &p2);
  out.pinned = pinned;
  for (size_t v = 0; v < n_visit; ++v) {
    const Segment& sg = out.segments[v];
    size_t tid = (size_t)sg.tid;
    uint32_t g = cfg.coverage_group_of_tid.empty() ? (uint32_t)tid : cfg.coverage_group_of_tid[tid];
    if (g > 255) throw std::runtime_error("more than 256 coverage groups are not supported");
    const std::string& s = *refseq[v];
    for (int32_t p = sg.lo; p < sg.hi; ++p) {
      uint8_t b = char_to_index(s[(size_t)p]);
      if (b > 5 || b == 4) throw std::runtime_error(std::string("Unrecognized base char in reference: ") + s[(size_t)p]);
      out.slot_ref[sg.slot0 + (uint64_t)(p - sg.lo)] = b;
      out.slot_group[sg.slot0 + (uint64_t)(p - sg.lo)] = (uint8_t)g;
    }
  }
  for (uint64_t j = 0; j < out.n_ins; ++j) out.slot_ref[out.n_base + j] = kBaseGap;

  // ---- table geometry of the device stream words (ScoreGeometry).  The dominant MAPQ and the quality
  // window are picked from per-read statistics (every base of every unique read), which follow the
  // scoring records closely and cost no column walk; the exact per-record histograms are gathered in pass B.
  if (cfg.want_score) {
    std::vector<uint64_t> mq(256, 0), qc(128, 0);
    std::mutex merge_mu;
    parallel_ranges(R.size(), [&](uint64_t lo, uint64_t hi) {  // (sums: the order of the parts does not matter)
      uint64_t mq_l[256] = {0}, qc_l[128] = {0};
      for (size_t i = lo; i < hi; ++i) {
        if (!in_pileup(i) || R.x1[i] != 1) continue;
        mq_l[R.mapq[i]] += info[i].L;
        const uint8_t* qual = R.quals.data() + R.seq_off[i];
        for (uint32_t k = 0; k < info[i].L; ++k) ++qc_l[qual[k] & 127];
      }
      std::lock_guard<std::mutex> g(merge_mu);
      for (int m = 0; m < 256; ++m) mq[m] += mq_l[m];
      for (int q = 0; q < 128; ++q) qc[q] += qc_l[q];
    });
    out.geo = choose_geometry(mq.data(), qc.data(), cfg, out.max_read_set_seen);
  }
  const ScoreGeometry geo = out.geo;
  const uint32_t n_hot = geo.n_hot();

  // Classic words of one pileup entry: the column itself (k = 0) and its insert sub-columns
  // (identify_mutations.cpp:1561-1657, error_count.cpp:1049-1105).  sink(slot, classic word, X1, ext).
  auto score_words = [&](size_t i, const ReadInfo& ri, int32_t q, bool is_del, int indel, uint64_t slot, auto&& sink) {
    const uint8_t* seq = R.bases.data() + R.seq_off[i];
    const uint8_t* qual = R.quals.data() + R.seq_off[i];
    const bool unique = R.x1[i] == 1;
    const uint32_t rev = ri.rev ? 1 : 0;
    const int32_t L = (int32_t)ri.L;
    const uint32_t mapq = R.mapq[i];
    const int ind = is_del ? -1 : std::max(indel, 0);
    const uint32_t K = sub_k[slot];
    const uint32_t q1 = (uint32_t)q + 1;
    for (uint32_t k = 0; k <= K; ++k) {
      const bool past_base = !(ind >= (int)k);
      uint8_t obs = past_base ? (uint8_t)kBaseGap : nibble_to_index(seq[q + (int32_t)k]);
      if (obs == kBaseN) continue;  // not even coverage
      uint32_t rec = obs;
      if (!rev) rec |= SR_TOP_BIT;
      bool trimmed = false;  // alignment.h:389-410 (unsigned comparisons as there)
      if (R.xl[i] >= 0 || R.xl[i] < -1) { if (q1 <= (uint32_t)R.xl[i]) trimmed = true; }
      if (R.xr[i] >= 0 || R.xr[i] < -1) {
        if ((uint32_t)L - q1 + 1 <= (uint32_t)R.xr[i]) trimmed = true;
        if (past_base && ((uint32_t)L - q1 == (uint32_t)R.xr[i])) trimmed = true;
      }
      if (trimmed) rec |= SR_TRIM_BIT;
      uint32_t ext = 0;  // read_pos and base_repeat of the quality position, when staged
      if (unique) {
        rec |= SR_UNIQUE_BIT;
        int32_t qp = q;
        bool ok = true;
        if (ind == -1) {
          qp += 1 - (int32_t)rev;
          if (qp >= L) throw std::runtime_error("deletion with no following read base (reference would assert)");
          if (seq[qp] == 15) ok = false;
        } else if (k > 0) {
          qp += std::min((int)k, ind) + 1 - (int32_t)rev;
          if (qp > ri.qb_end0) ok = false;
          else if (seq[qp] == 15) ok = false;
        }
        if (ok) {
          uint32_t qv = qual[qp];
          if (qv > 127) throw std::runtime_error("base quality above 127 cannot be packed");
          rec |= SR_OK_BIT | (qv << SR_QUAL_SHIFT);
          if (geo.side_stride == 2) {
            if (qp > 65535) throw std::runtime_error("read position above 65535 cannot be packed");
            ext = (uint32_t)qp;
            if (cfg.use_base_repeat) {  // alignment.cpp:371-390
              const uint8_t b = seq[qp];
              int32_t x = qp; uint32_t rp = 0;
              if (!rev) { while (x < ri.qe0) { ++x; if (seq[x] != b) break; ++rp; } }
              else { while (x > 0) { --x; if (seq[x] != b) break; ++rp; } }
              ext |= std::min<uint32_t>(rp, 255u) << 16;
            }
          }
        }
        rec |= mapq << SR_MAPQ_SHIFT;
        rec |= (uint32_t)ri.read_set << SR_SET_SHIFT;
      } else {
        rec |= std::min<uint32_t>(R.x1[i], SR_RED_MASK) << SR_RED_SHIFT;
      }
      sink(k == 0 ? slot : out.n_base + sub_first[slot] + k - 1, rec, R.x1[i], ext);
    }
  };
  // Device stream word (and side-list entry, if any) of a classic word in a slot with reference base `ref`.
  struct DevWord { uint32_t dev, side; bool has_side; };
  auto encode = [&](uint32_t rec, uint32_t x1, uint32_t ref) {
    DevWord w{0, 0, false};
    const bool is_top = (rec & SR_TOP_BIT) != 0;
    const uint32_t top = is_top ? DR_TOP_BIT : 0u;
    if (!(rec & SR_UNIQUE_BIT)) {
      w.dev = DR_REDUNDANT | top | std::min<uint32_t>(x1, DR_X1_MASK) << DR_X1_SHIFT | geo.special_counter(SC_TRASH) |
              (rec & 7u) << DR_RED_OBS_SHIFT | ((rec & SR_TRIM_BIT) ? DR_RED_TRIM_BIT : 0u);
      if (x1 >= DR_X1_MASK) { w.has_side = true; w.side = SIDE_BIG | x1; }
      return w;
    }
    const uint32_t qv = (rec >> SR_QUAL_SHIFT) & 127u, obs = rec & 7u;
    if ((rec & SR_TRIM_BIT) || !(rec & SR_OK_BIT) || qv < geo.cutoff) {
      w.dev = DR_IDLE | top | geo.special_counter(is_top ? SC_IDLE_TOP : SC_IDLE_BOT);
      return w;
    }
    const bool match = obs == ref;
    // HOT: a record of the shared table's classes that MATCHES the reference base; one that does not is COLD (side list)
    if (match && n_hot && ((rec >> SR_MAPQ_SHIFT) & 255u) == geo.hot_mapq && obs < 5 && qv >= geo.q_lo && qv - geo.q_lo < geo.n_q) {
      const uint32_t sq = ((rec >> 10) & 63u) * geo.n_q + (qv - geo.q_lo);
      w.dev = sq << DR_SQ_SHIFT | obs << DR_OBS_SHIFT | top | DR_MATCH_BIT | ScoreGeometry::counter_of(sq);
    } else {
      w.dev = DR_COLD | top | geo.special_counter(is_top ? SC_COLD_TOP : SC_COLD_BOT);
      w.has_side = true; w.side = rec | (match ? SR_MATCH_BIT : 0u);
    }
    return w;
  };

  phase_done("geometry");
  // ---- pass A2: record counts per slot
  std::vector<uint32_t> score_cnt(cfg.want_score ? n_slots : 0, 0), hist_cnt(cfg.want_hist ? out.n_base : 0, 0);
  std::vector<uint32_t> red_cnt(cfg.want_score ? n_slots : 0, 0);  // redundant records per slot: they lead the slot's run
  std::vector<uint32_t> side_cnt(cfg.want_score ? n_slots : 0, 0), side_red_cnt(cfg.want_score ? n_slots : 0, 0);
  std::vector<uint8_t> col_red(cfg.want_hist ? out.n_base : 0, 0);
  std::vector<uint8_t> col_qstart(cfg.want_hist && cfg.preprocess_stage ? out.n_base : 0, 0);  // bit 0 / 1: a top / bottom strand read starts here
  auto junction_read_end_min = [&](uint32_t L) -> uint32_t {  // Settings::required_junction_read_end_min_coordinate, settings.h:345-354
    const int32_t max_len = (int32_t)floor((double)((int32_t)L - (int32_t)cfg.unmatched_end_minimum_read_length) * cfg.unmatched_end_length_factor);
    return max_len <= 0 ? L : L - (uint32_t)max_len;
  };
  run_items([&](size_t ii) {
    const Item& it = items[ii];
    const uint64_t s0 = out.segments[it.v].slot0 - (uint64_t)out.segments[it.v].lo;  // slot = s0 + column
    for (size_t i = it.first_read; i < it.last_read; ++i) {
      if (!in_pileup(i) || info[i].end <= it.lo) continue;
      const bool unique = R.x1[i] == 1;
      const uint32_t L = info[i].L;
      walk_read(R.cigars.data() + R.cigar_off[i], R.n_cigar[i], R.pos[i], it.lo, it.hi, [&](int32_t c, int32_t q, bool is_del, int indel) {
        const uint64_t slot = s0 + (uint64_t)c;
        if ((uint32_t)q >= L && !is_del) throw std::runtime_error("CIGAR longer than the read sequence");
        if (cfg.want_hist && !is_del) { if (unique) ++hist_cnt[slot]; else col_red[slot] = 1; }
        if (!col_qstart.empty() && !is_del && unique && q == 0) {  // error_count.cpp:157-166; query_stranded_end_1: alignment.cpp:228-239
          const uint32_t stranded_end_1 = info[i].rev ? L - (uint32_t)(info[i].qb_start0 + 1) + 1 : (uint32_t)info[i].qb_end0 + 1;
          if (stranded_end_1 >= junction_read_end_min(L)) col_qstart[slot] |= info[i].rev ? 2 : 1;
        }
        if (cfg.want_score)
          score_words(i, info[i], q, is_del, indel, slot, [&](uint64_t s, uint32_t rec, uint32_t x1, uint32_t) {
            ++score_cnt[s];
            if (!unique) ++red_cnt[s];
            const DevWord w = encode(rec, x1, out.slot_ref[s]);
            if (w.has_side) { ++side_cnt[s]; if (!unique) ++side_red_cnt[s]; }
          });
      });
    }
  });

  phase_done("pass A2 (counts)");
  if (!col_qstart.empty()) {  // error_count.cpp:191-194: columns without a redundant read, both strands
    out.read_start_counts.assign(n_targets * 2, 0);
    for (const Segment& sg : out.segments)
      for (int32_t c = sg.lo; c < sg.hi; ++c) {
        const uint64_t slot = sg.slot0 + (uint64_t)(c - sg.lo);
        if (col_red[slot]) continue;
        ++out.read_start_counts[(size_t)sg.tid * 2 + (col_qstart[slot] & 1)];
        ++out.read_start_counts[(size_t)sg.tid * 2 + ((col_qstart[slot] >> 1) & 1)];
      }
  }

  // ---- offsets
  {
    bool p3 = false;
    out.side_off = (uint32_t*)alloc((n_slots + 1) * 4, &p3);
    uint64_t sacc = 0;
    // every slot's range starts on an even entry and is padded to an even count with SIDE_PAD (the tally kernel reads
    // two entries per request)
    for (uint64_t s = 0; s < n_slots; ++s) { out.side_off[s] = (uint32_t)sacc; if (cfg.want_score) sacc += (side_cnt[s] + 1u) & ~1u; }
    if (sacc >= (1ull << 29)) throw std::runtime_error("more than 2^29 side-list entries in one staged stream");
    out.side_off[n_slots] = (uint32_t)sacc; out.n_side = sacc;
    out.side_rec = (uint32_t*)alloc(sacc * 4 * out.geo.side_stride + 16, &p3);
    std::fill(out.side_rec, out.side_rec + sacc * out.geo.side_stride, SIDE_PAD);
    // Rounds of the tally kernel: 32 slots that share their reference base (its contraction multiplies the 32 class
    // histograms with ONE base's likelihood table) and have about the same depth (its 32 lanes walk their runs in
    // lock step).  Inside every block of ROUND_BLOCK consecutive slots the slots are grouped by base (A, C, G, T,
    // other) and ordered by depth; a group is padded to whole rounds with ROUND_NO_SLOT.
    out.score_cnt = (uint32_t*)alloc((n_slots + 1) * 4, &p3);
    uint64_t n_true = 0;
    for (uint64_t s = 0; s < n_slots; ++s) { out.score_cnt[s] = cfg.want_score ? score_cnt[s] : 0; n_true += out.score_cnt[s]; }
    out.n_score = n_true;
    {
      if (n_slots >= ROUND_NO_SLOT) throw std::runtime_error("more than 2^32 - 2 slots in one staged stream");
      std::vector<uint32_t> order;
      order.reserve(n_slots + n_slots / 16 + 160);
      std::vector<uint32_t> group[5];
      auto vecs = [&](uint32_t s) { return (out.score_cnt[s] + 7u) >> 3; };
      for (uint64_t b0 = 0; b0 < n_slots; b0 += ROUND_BLOCK) {
        const uint64_t b1 = std::min<uint64_t>(n_slots, b0 + ROUND_BLOCK);
        for (auto& g : group) g.clear();
        for (uint64_t s = b0; s < b1; ++s) group[out.slot_ref[s] < 4 ? out.slot_ref[s] : 4].push_back((uint32_t)s);
        for (auto& g : group) {
          std::stable_sort(g.begin(), g.end(), [&](uint32_t a, uint32_t b) { return vecs(a) < vecs(b); });
          order.insert(order.end(), g.begin(), g.end());
          while (order.size() & 31) order.push_back(ROUND_NO_SLOT);
        }
      }
      out.n_rounds = order.size() / 32;
      out.round_slot = (uint32_t*)alloc(order.size() * 4 + 16, &p3);
      std::copy(order.begin(), order.end(), out.round_slot);
      // The record stream is round-major and lane-interleaved (brq_types.h): round r holds as many 1 KB round vectors
      // as its deepest slot has 256-bit vectors; shallower lanes and idle lanes are filled with pad words.
      out.round_off = (uint64_t*)alloc((out.n_rounds + 1) * 8, &p3);
      uint64_t acc = 0;
      for (uint64_t r = 0; r < out.n_rounds; ++r) {
        out.round_off[r] = acc;
        uint32_t deepest = 0;
        for (uint32_t l = 0; l < 32; ++l) {
          const uint32_t sl = out.round_slot[r * 32 + l];
          if (sl == ROUND_NO_SLOT) continue;
          out.score_off[sl] = acc + l * 4u;
          deepest = std::max(deepest, vecs(sl));
        }
        acc += (uint64_t)deepest * ROUND_VECTOR_WORDS;
      }
      out.round_off[out.n_rounds] = acc;
      // what a lane needs of its slot besides the records, round-major so that the warp reads it coalesced
      out.round_side = (uint32_t*)alloc(order.size() * 8 + 16, &p3);
      for (size_t i = 0; i < order.size(); ++i) {
        const uint32_t sl = order[i];
        out.round_side[2 * i] = sl == ROUND_NO_SLOT ? (5u << 29) : (out.side_off[sl] | (uint32_t)out.slot_ref[sl] << 29);
        out.round_side[2 * i + 1] = sl == ROUND_NO_SLOT ? 0u : out.side_off[sl] + (cfg.want_score ? side_cnt[sl] : 0u);
      }
      out.score_off[n_slots] = acc;
      out.n_score_padded = acc;
    }
    uint64_t acc = 0;
    for (uint64_t c = 0; c < out.n_base; ++c) {
      out.hist_off[c] = acc | ((cfg.want_hist && col_red[c]) ? HIST_OFF_REDUNDANT_BIT : 0);
      if (cfg.want_hist) acc += hist_cnt[c];
    }
    out.hist_off[out.n_base] = acc; out.n_hist = acc;
    if (cfg.want_hist) for (uint64_t c = 0; c < out.n_base; ++c) if (hist_cnt[c] > out.max_hist_depth) out.max_hist_depth = hist_cnt[c];
  }
  phase_done("offsets and rounds");
  // the positional forms stay in plain memory when only their transfer / compact forms cross PCIe (pinning gigabytes is slow)
  const bool build_score16 = cfg.want_score && cfg.compact_score && out.n_rounds && out.n_rounds * 32 < (1ull << 32) - 1;
  const bool build_hist16 = cfg.want_hist && cfg.compact_hist && !(cfg.use_base_repeat || cfg.use_read_pos || out.max_read_set_seen > 7);
  out.score_rec_plain = build_score16; out.hist_rec_plain = build_hist16; out.hist_compact = build_hist16;
  out.score_rec = (uint32_t*)(build_score16 ? default_alloc(out.n_score_padded * 4, &p2) : alloc(out.n_score_padded * 4, &p2));
  parallel_ranges(out.n_score_padded, [&](uint64_t lo, uint64_t hi) { std::fill(out.score_rec + lo, out.score_rec + hi, geo.pad_word()); });  // pad word: the trash counter, no other bit (threads: the first touch of 2 GB)
  out.hist_bytes = (cfg.use_base_repeat || cfg.use_read_pos || out.max_read_set_seen > 7) ? 8 : 4;
  out.hist_rec = build_hist16 ? default_alloc(out.n_hist * out.hist_bytes, &p2) : alloc(out.n_hist * out.hist_bytes, &p2);

  // ---- pass B: fill.  Within every slot the redundant records come first and the unique ones
  // follow, each part in arrival (BAM) order: redundant records never score, so the scoring records
  // keep the reference's order, and the order-dependent sum of 1/X1 only has to walk the slot's head.
  std::vector<uint32_t>& score_cur = score_cnt;  // reused as per-slot cursors: unique records start after the redundant ones
  std::vector<uint32_t>& red_cur = red_cnt;
  std::vector<uint32_t>& hist_cur = hist_cnt;
  for (size_t s = 0; s < score_cur.size(); ++s) { score_cur[s] = red_cnt[s]; red_cur[s] = 0; }
  // side-list cursors: a slot's SIDE_BIG entries (redundant records) precede its cold entries, like the records themselves
  std::vector<uint32_t>& side_cur = side_cnt;
  std::vector<uint32_t>& side_red_cur = side_red_cnt;
  for (size_t s = 0; s < side_cur.size(); ++s) { side_cur[s] = side_red_cnt[s]; side_red_cur[s] = 0; }
  std::fill(hist_cur.begin(), hist_cur.end(), 0);
  std::vector<uint64_t> qual_counts((size_t)items.size() * 128, 0);
  std::vector<uint32_t> mapq_masks((size_t)items.size() * 8, 0);
  std::vector<uint64_t> mapq_counts((size_t)items.size() * 256, 0);
  std::vector<uint32_t> max_quals(items.size(), 0), max_hquals(items.size(), 0), max_rposs(items.size(), 0), max_srposs(items.size(), 0);
  run_items([&](size_t ii) {
    const Item& it = items[ii];
    const uint64_t s0 = out.segments[it.v].slot0 - (uint64_t)out.segments[it.v].lo;  // slot = s0 + column
    const std::string& rs = *refseq[it.v];
    auto ref_index = [&](int32_t p) -> uint32_t {  // forward-strand reference base; the byte past the end is the NUL terminator
      if (p >= tlen_of(rs)) return kBaseNul;
      uint8_t b = char_to_index(rs[(size_t)p]);
      if (b > 5 || b == 4) throw std::runtime_error(std::string("Unrecognized base char in reference: ") + rs[(size_t)p]);
      return b;
    };
    // per-item statistics are gathered on the stack and stored once: neighbouring items run on different threads, and
    // their slices of the shared vectors share cache lines (the 32-byte MAPQ masks did for every scoring record)
    uint32_t mq_mask[8] = {0};
    uint64_t mq_count[256] = {0}, q_count[128] = {0};
    uint32_t max_q = 0, max_hq = 0, max_rp = 0, max_srp = 0;
    for (size_t i = it.first_read; i < it.last_read; ++i) {
      if (!in_pileup(i) || info[i].end <= it.lo) continue;
      const ReadInfo& ri = info[i];
      const uint8_t* seq = R.bases.data() + R.seq_off[i];
      const uint8_t* qual = R.quals.data() + R.seq_off[i];
      const bool unique = R.x1[i] == 1;
      const uint32_t rev = ri.rev ? 1 : 0;
      const int32_t L = (int32_t)ri.L;
      if (R.x1[i] == 0) throw std::runtime_error("X1:i:0 is not a valid redundancy");
      const uint32_t mapq = R.mapq[i];
      walk_read(R.cigars.data() + R.cigar_off[i], R.n_cigar[i], R.pos[i], it.lo, it.hi, [&](int32_t c, int32_t q, bool is_del, int indel) {
        const uint64_t slot = s0 + (uint64_t)c;
        // ---------------- error_count record (error_count.cpp:125-199, 854-986)
        if (cfg.want_hist && !is_del && unique) {
          uint64_t rec = 0;
          const uint32_t qa = qual[q];
          if (qa > 127) throw std::runtime_error("base quality above 127 cannot be packed");
          const uint32_t obsA = nibble_to_index(seq[q]), refA = out.slot_ref[slot];
          auto strand = [&](uint32_t b) { return rev ? 3 - b : b; };  // complement on the read strand (A,C,G,T only)
          if (obsA < 4 && refA < 4) {  // error_count.cpp:870-905
            rec |= (uint64_t)strand(refA) << HR_REFA | (uint64_t)strand(obsA) << HR_OBSA | (uint64_t)qa << HR_QUALA | 1ull << HR_VALIDA;
            if (qa > max_hq) max_hq = qa;
          }
          rec |= (uint64_t)(ri.read_set & 7u) << HR_SET | (uint64_t)(ri.read_set >> 3) << HR_SET_HI;
          if (q > 65535) throw std::runtime_error("read position above 65535 cannot be packed");
          rec |= (uint64_t)q << HR_RPOS;
          if ((uint32_t)q > max_rp) max_rp = (uint32_t)q;
          auto base_repeat = [&](int32_t qp) -> uint64_t {  // alignment.cpp:371-390
            uint8_t b = seq[qp]; uint32_t rep = 0;
            if (!rev) { while (qp < ri.qe0) { ++qp; if (seq[qp] != b) break; ++rep; } }
            else { while (qp > 0) { --qp; if (seq[qp] != b) break; ++rep; } }
            return rep;
          };
          if (cfg.use_base_repeat) rec |= std::min<uint64_t>(base_repeat(q), 255) << HR_REPA;
          uint32_t cls = 0; int32_t m = -1; uint32_t refb = kBaseNul;
          if (indel == 0) {
            if (q < ri.qe0) {
              cls = 1; m = q + 1 - (int32_t)rev;
              int32_t mr = c + 1 - (int32_t)rev;
              refb = ref_index(mr);
            }
          } else if (indel == -1) {
            cls = 2; m = q + 1 - (int32_t)rev;
            refb = ref_index(c + 1);
          } else if (indel == 1) {
            m = q + 1;
            if (m >= L) throw std::runtime_error("Attempt to retrieve quality score for nonexistent base for '.N' state.");
            if (m <= ri.qe0 && m >= ri.qs0) cls = 3;
          }
          if (cls) {
            if (m < 0 || m >= L) throw std::runtime_error("Attempt to retrieve quality score for nonexistent base.");
            if (qual[m] > 127) throw std::runtime_error("base quality above 127 cannot be packed");
            const uint32_t obsB = nibble_to_index(seq[m]);
            uint32_t fB = kBaseGap, oB = kBaseGap;
            bool valid = obsB != kBaseN;                         // error_count.cpp:909-982
            if (cls == 1) valid = valid && refb != kBaseN;       // ('.', '.'): both only tested for N
            else if (cls == 2) { valid = valid && refb < 4; fB = valid ? strand(refb) : kBaseGap; }
            else oB = valid ? strand(obsB) : kBaseGap;
            if (valid) {
              rec |= (uint64_t)fB << HR_REFB | (uint64_t)oB << HR_OBSB | (uint64_t)qual[m] << HR_QUALB | 1ull << HR_VALIDB;
              if (qual[m] > max_hq) max_hq = qual[m];
              if (cfg.use_base_repeat) rec |= std::min<uint64_t>(base_repeat(m), 31) << HR_REPB;
            }
          }
          {  // the fast kind: A valid with ref == obs, B ('.', '.') or absent
            const bool a_match = (rec >> HR_VALIDA & 1) && ((rec >> HR_REFA & 7) == (rec >> HR_OBSA & 7));
            const bool b_valid = rec >> HR_VALIDB & 1;
            const bool b_dots = b_valid && (rec >> HR_REFB & 7) == kBaseGap && (rec >> HR_OBSB & 7) == kBaseGap;
            if (a_match && (b_dots || !b_valid)) {
              rec |= 1ull << HR_FAST;
              if (!b_valid) rec |= 127ull << HR_QUALB;
            }
          }
          const uint64_t at = (out.hist_off[slot] & ~HIST_OFF_REDUNDANT_BIT) + hist_cur[slot]++;
          if (out.hist_bytes == 8) static_cast<uint64_t*>(out.hist_rec)[at] = rec;
          else static_cast<uint32_t*>(out.hist_rec)[at] = (uint32_t)rec;
        }
        // ---------------- identify_mutations records
        if (!cfg.want_score) return;
        score_words(i, ri, q, is_del, indel, slot, [&](uint64_t s, uint32_t rec, uint32_t x1, uint32_t ext) {
          const DevWord w = encode(rec, x1, out.slot_ref[s]);
          out.score_rec[score_index(out.score_off[s], unique ? score_cur[s]++ : red_cur[s]++)] = w.dev;
          if (w.has_side) {
            const size_t e = (size_t)(out.side_off[s] + (unique ? side_cur[s]++ : side_red_cur[s]++)) * geo.side_stride;
            out.side_rec[e] = w.side;
            if (geo.side_stride == 2) { out.side_rec[e + 1] = ext; if ((w.side & SIDE_BIG) == 0 && (ext & 0xFFFFu) > max_srp) max_srp = ext & 0xFFFFu; }
          }
          const uint32_t kind = w.dev >> DR_KIND_SHIFT;
          if (kind == 0 || kind == 2) {  // a scoring record: exact statistics for the likelihood tables
            const uint32_t qv = (rec >> SR_QUAL_SHIFT) & 127u;
            mq_mask[mapq >> 5] |= 1u << (mapq & 31); ++mq_count[mapq]; ++q_count[qv]; if (qv > max_q) max_q = qv;
          }
        });
      });
    }
    max_quals[ii] = max_q; max_hquals[ii] = max_hq; max_rposs[ii] = max_rp; max_srposs[ii] = max_srp;
    std::copy(mq_mask, mq_mask + 8, &mapq_masks[ii * 8]);
    std::copy(mq_count, mq_count + 256, &mapq_counts[ii * 256]);
    std::copy(q_count, q_count + 128, &qual_counts[ii * 128]);
  });
  if (cfg.want_score) out.mapq_seen[geo.hot_mapq >> 5] |= 1u << (geo.hot_mapq & 31);  // the shared table is always built
  for (size_t ii = 0; ii < items.size(); ++ii) {
    for (int w = 0; w < 8; ++w) out.mapq_seen[w] |= mapq_masks[ii * 8 + (size_t)w];
    for (int m = 0; m < 256; ++m) out.mapq_count[m] += mapq_counts[ii * 256 + (size_t)m];
    for (int q = 0; q < 128; ++q) out.qual_count[q] += qual_counts[ii * 128 + (size_t)q];
    out.max_qual_seen = std::max(out.max_qual_seen, max_quals[ii]);
    out.max_hist_qual = std::max(out.max_hist_qual, max_hquals[ii]);
    out.max_hist_rpos = std::max(out.max_hist_rpos, max_rposs[ii]);
    out.max_score_rpos = std::max(out.max_score_rpos, max_srposs[ii]);
  }

  phase_done("pass B (fill)");
  // ---- transfer form of the scoring stream (brq_types.h): low halves + the words they do not determine
  if (build_score16) {
    const ScoreRecon rc = score_recon_of(geo);
    const uint64_t n_lanes = out.n_rounds * 32;
    bool p5 = false;
    out.score16 = (uint16_t*)alloc(out.n_score_padded * 2 + 32, &p5);
    out.score_exc_off = (uint32_t*)alloc((n_lanes + 1) * 4, &p5);
    auto parts = [&](auto&& body) { parallel_ranges(out.n_rounds, body); };
    // pass 1: low halves, exception flags, exceptions per lane (words visited in memory order)
    parts([&](uint64_t r0, uint64_t r1) {
      for (uint64_t r = r0; r < r1; ++r) {
        const uint64_t beg = out.round_off[r], end = out.round_off[r + 1];
        uint32_t ref[32], n_exc[32];
        for (uint32_t l = 0; l < 32; ++l) {
          const uint32_t sl = out.round_slot[r * 32 + l];
          ref[l] = sl == ROUND_NO_SLOT ? 5u : out.slot_ref[sl]; n_exc[l] = 0;
        }
        for (uint64_t p = beg; p < end; ++p) {  // word p belongs to lane (p - beg) / 4 % 32
          const uint32_t l = (uint32_t)((p - beg) >> 2) & 31u, w = out.score_rec[p], lo = w & 0x7FFFu;
          const bool same = !(w & S16_EXCEPTION) && score_word_from16(lo, ref[l], rc) == w;
          out.score16[p] = (uint16_t)(same ? lo : (lo | S16_EXCEPTION));
          n_exc[l] += !same;
        }
        for (uint32_t l = 0; l < 32; ++l) out.score_exc_off[r * 32 + l + 1] = n_exc[l];
      }
    });
    out.score_exc_off[0] = 0;
    uint64_t acc = 0;
    for (uint64_t i = 1; i <= n_lanes; ++i) { acc += out.score_exc_off[i]; out.score_exc_off[i] = (uint32_t)acc; }
    if (acc >= (1ull << 32)) throw std::runtime_error("more than 2^32 exception words in the scoring stream's transfer form");
    out.n_score_exc = acc;
    out.score_exc = (uint32_t*)alloc(acc * 4 + 32, &p5);
    // pass 2: the exception words, per lane in record order (memory order visits a lane's records in order)
    parts([&](uint64_t r0, uint64_t r1) {
      for (uint64_t r = r0; r < r1; ++r) {
        if (out.score_exc_off[r * 32 + 32] == out.score_exc_off[r * 32]) continue;
        const uint64_t beg = out.round_off[r], end = out.round_off[r + 1];
        uint32_t* dst[32];
        for (uint32_t l = 0; l < 32; ++l) dst[l] = out.score_exc + out.score_exc_off[r * 32 + l];
        for (uint64_t p = beg; p < end; ++p)
          if (out.score16[p] & S16_EXCEPTION) *dst[(uint32_t)((p - beg) >> 2) & 31u]++ = out.score_rec[p];
      }
    });
  }

  phase_done("score transfer form");
  // ---- compact histogram streams (brq_types.h): fast records in 16 bits, the others unchanged
  if (build_hist16) {
    const uint32_t* h = static_cast<const uint32_t*>(out.hist_rec);
    const size_t n = out.n_hist, n_parts = (size_t)std::max(1, n_threads) * 4;
    std::vector<uint64_t> c16(n_parts + 1, 0), cex(n_parts + 1, 0);
    auto parts = [&](auto&& body) {
      std::atomic<size_t> next(0);
      auto work = [&]() { for (;;) { const size_t k = next.fetch_add(1); if (k >= n_parts) break; body(k, n * k / n_parts, n * (k + 1) / n_parts); } };
      std::vector<std::thread> pool;
      for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
      work();
      for (auto& t : pool) t.join();
    };
    parts([&](size_t k, size_t lo, size_t hi) {
      uint64_t a = 0;
      for (size_t i = lo; i < hi; ++i) a += hist16_pack(h[i]) < 0x10000u;
      c16[k + 1] = a; cex[k + 1] = (hi - lo) - a;
    });
    for (size_t k = 0; k < n_parts; ++k) { c16[k + 1] += c16[k]; cex[k + 1] += cex[k]; }
    out.n_hist16 = c16[n_parts]; out.n_hist_exc = cex[n_parts];
    bool p4 = false;
    out.hist16 = (uint16_t*)alloc(out.n_hist16 * 2 + 32, &p4);
    out.hist_exc = (uint32_t*)alloc(out.n_hist_exc * 4 + 32, &p4);
    parts([&](size_t k, size_t lo, size_t hi) {
      uint16_t* d16 = out.hist16 + c16[k];
      uint32_t* dex = out.hist_exc + cex[k];
      for (size_t i = lo; i < hi; ++i) {
        const uint32_t r = hist16_pack(h[i]);
        if (r < 0x10000u) *d16++ = (uint16_t)r; else *dex++ = h[i];
      }
    });
  }
  phase_done("compact histogram");
}

}  // namespace brq
