#include "synth.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <thread>

namespace brq {

namespace {

inline uint64_t splitmix(uint64_t& s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
inline uint64_t mix3(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t s = a * 0x9E3779B97F4A7C15ull ^ (b + 0x7F4A7C15ull) * 0xBF58476D1CE4E5B9ull ^ (c + 0x1CE4E5B9ull) * 0x94D049BB133111EBull;
  splitmix(s);
  return splitmix(s);
}
struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() { return splitmix(s); }
  uint32_t below(uint32_t n) { return (uint32_t)((next() >> 32) * (uint64_t)n >> 32); }
  bool ppm(uint32_t p) { return below(1000000u) < p; }
  double normal() {  // Irwin-Hall(12) - 6
    uint64_t acc = 0;
    for (int i = 0; i < 12; ++i) acc += next() >> 48;
    return (double)acc / 65536.0 - 6.0;
  }
};

struct TmpRead {
  int32_t tid, pos;
  uint16_t flag;
  uint8_t mapq, rg;
  uint32_t x1;
  int32_t xl, xr, as;
  uint64_t frag;
  uint8_t mate;
  std::vector<uint8_t> bases, quals;
  std::vector<uint32_t> cigar;
};

struct Gap { int32_t tid, beg, end; };

void add_op(std::vector<uint32_t>& cig, uint32_t op, uint32_t len) {
  if (!cig.empty() && (cig.back() & 0xf) == op) cig.back() += len << 4;
  else cig.push_back(len << 4 | op);
}

struct Model {
  const SynthConfig* cfg;
  const RefSet* ref;
  std::vector<std::vector<uint32_t>> var_by_tid;  // indices into variants, sorted by pos0
  const std::vector<SynthVariant>* variants;
  std::vector<Gap> gaps;
  uint32_t perr_ppm[64];
};

bool make_read(const Model& m, int32_t tid, int32_t start, bool reversed, uint32_t L, uint64_t frag_gid,
               Rng& rng, TmpRead& out) {
  const SynthConfig& c = *m.cfg;
  const std::string& seq = m.ref->seqs[(size_t)tid];
  const int32_t reflen = (int32_t)seq.size();
  uint32_t sl = 0, sr = 0;
  if (rng.ppm(c.softclip_ppm)) {
    if (rng.below(2)) sl = 1 + rng.below(5); else sr = 1 + rng.below(5);
    if (L < 20) sl = sr = 0;
  }
  out.bases.assign(L, 0);
  out.quals.assign(L, 0);
  out.cigar.clear();
  auto qual_at = [&](uint32_t j) {
    uint32_t jj = reversed ? (L - 1 - j) : j;
    double mean = c.q_start + (c.q_end - c.q_start) * (L > 1 ? (double)jj / (double)(L - 1) : 0.0);
    int q = (int)floor(mean + c.q_sd * rng.normal() + 0.5);
    return (uint8_t)std::min(c.q_max, std::max(c.q_min, q));
  };
  auto emit = [&](uint32_t j, uint8_t base) {
    uint8_t q = qual_at(j);
    if (rng.ppm(c.n_base_ppm)) { out.bases[j] = 15; out.quals[j] = (uint8_t)c.q_min; return; }
    if (rng.below(1000000u) < m.perr_ppm[q]) base = (uint8_t)((base + 1 + rng.below(3)) & 3);
    out.bases[j] = (uint8_t)(1u << base);
    out.quals[j] = q;
  };
  uint32_t j = 0;
  for (; j < sl; ++j) emit(j, (uint8_t)rng.below(4));
  if (sl) add_op(out.cigar, 4, sl);
  const uint32_t aligned_end = L - sr;
  const std::vector<uint32_t>& vars = m.var_by_tid[(size_t)tid];
  size_t vi = std::lower_bound(vars.begin(), vars.end(), start, [&](uint32_t v, int32_t p) { return (*m.variants)[v].pos0 < p; }) - vars.begin();
  int32_t p = start;
  int last = -1;  // last emitted op: 0 M, 1 I, 2 D
  while (j < aligned_end) {
    if (p >= reflen) return false;
    while (vi < vars.size() && (*m.variants)[vars[vi]].pos0 < p) ++vi;
    const SynthVariant* v = (vi < vars.size() && (*m.variants)[vars[vi]].pos0 == p) ? &(*m.variants)[vars[vi]] : nullptr;
    bool carrier = v && (mix3(c.seed ^ 0x5eedull, vars.empty() ? 0 : vars[vi], frag_gid) % 1000000ull) < v->freq_ppm;
    if (v && carrier && v->kind == 1 && last == 0 && j + 2 <= aligned_end && p + (int32_t)v->len < reflen) {  // the reference reads the 2nd base after a deletion
      add_op(out.cigar, 2, v->len); p += v->len; last = 2; continue;
    }
    if (last == 0 && j + 2 <= aligned_end && rng.ppm(c.indel_error_ppm) && p + 1 < reflen) { add_op(out.cigar, 2, 1); p += 1; last = 2; continue; }
    uint8_t base = char_to_index(seq[(size_t)p]);
    if (base > 3) base = (uint8_t)rng.below(4);
    if (v && carrier && v->kind == 0) base = v->alt[0];
    emit(j, base);
    add_op(out.cigar, 0, 1); ++j; ++p; last = 0;
    if (v && carrier && v->kind == 2 && j + v->len < aligned_end) {
      for (uint32_t k = 0; k < v->len; ++k) emit(j + k, v->alt[k]);
      add_op(out.cigar, 1, v->len); j += v->len; last = 1;
    } else if (rng.ppm(c.indel_error_ppm) && j + 1 < aligned_end) {
      emit(j, (uint8_t)rng.below(4));
      add_op(out.cigar, 1, 1); ++j; last = 1;
    }
  }
  for (; j < L; ++j) emit(j, (uint8_t)rng.below(4));
  if (sr) add_op(out.cigar, 4, sr);
  for (const Gap& g : m.gaps) if (g.tid == tid && start < g.end && p > g.beg) return false;
  out.tid = tid; out.pos = start;
  out.mapq = rng.ppm(c.low_mapq_ppm) ? (uint8_t)rng.below(42) : 42;
  out.x1 = rng.ppm(c.redundant_ppm) ? 2 + rng.below(4) : 1;
  out.xl = out.xr = -1;
  if (rng.ppm(c.trim_ppm)) { out.xl = (int32_t)(1 + rng.below(5)); out.xr = (int32_t)rng.below(6); }
  else if (rng.below(2)) { out.xl = 0; out.xr = 0; }
  out.as = (int32_t)L;
  out.flag = reversed ? 16 : 0;
  return true;
}


// planted variants and sample gaps: a pure function of the configuration and the reference
void plant(const SynthConfig& cfg, const RefSet& ref, const std::vector<uint64_t>& cum, std::vector<SynthVariant>& variants, std::vector<Gap>& gaps) {
  const uint64_t total = ref.total_length();
  auto locate = [&](uint64_t g, int32_t& tid, int32_t& pos) {
    size_t t = std::upper_bound(cum.begin(), cum.end(), g) - cum.begin() - 1;
    tid = (int32_t)t; pos = (int32_t)(g - cum[t]);
  };
  variants.clear();
  gaps.clear();
    Rng rng(mix3(cfg.seed, 0xA11E1E, 0));
    uint32_t n_var = cfg.n_polymorphic + cfg.n_fixed;
    std::vector<uint64_t> gpos;
    for (uint32_t i = 0; i < n_var; ++i) {
      for (int tries = 0; tries < 1000; ++tries) {
        uint64_t g = rng.next() % total;
        bool ok = true;
        for (uint64_t o : gpos) if ((g > o ? g - o : o - g) < 12) { ok = false; break; }
        int32_t tid, pos; locate(g, tid, pos);
        if (pos < 5 || pos + 8 >= (int32_t)ref.seqs[(size_t)tid].size()) ok = false;
        if (ok) { gpos.push_back(g); break; }
      }
    }
    for (size_t i = 0; i < gpos.size(); ++i) {
      SynthVariant v;
      locate(gpos[i], v.tid, v.pos0);
      uint32_t k = rng.below(10);
      v.kind = k < 6 ? 0 : (k < 8 ? 1 : 2);
      v.len = v.kind == 0 ? 1 : (uint8_t)(1 + (rng.below(4) == 0 ? 1 + rng.below(2) : 0));
      uint8_t refb = char_to_index(ref.seqs[(size_t)v.tid][(size_t)v.pos0]);
      for (int a = 0; a < 3; ++a) v.alt[a] = (uint8_t)rng.below(4);
      if (v.kind == 0) v.alt[0] = (uint8_t)(((refb < 4 ? refb : 0) + 1 + rng.below(3)) & 3);
      v.freq_ppm = i < cfg.n_polymorphic ? cfg.min_freq_ppm + rng.below(cfg.max_freq_ppm - cfg.min_freq_ppm + 1) : 1000000u;
      variants.push_back(v);
    }
    for (uint32_t i = 0; i < cfg.n_gaps; ++i) {
      Gap g;
      uint32_t len = cfg.gap_min + rng.below(cfg.gap_max - cfg.gap_min + 1);
      int32_t pos; locate(rng.next() % total, g.tid, pos);
      int32_t tl = (int32_t)ref.seqs[(size_t)g.tid].size();
      if ((int32_t)len + 200 >= tl) continue;
      g.beg = std::min(std::max(pos, 100), tl - (int32_t)len - 100);
      g.end = g.beg + (int32_t)len;
      gaps.push_back(g);
    }
}

}  // namespace

void synth_reference(uint64_t seed, const std::vector<uint32_t>& lens, const std::string& prefix, RefSet& ref) {
  int width = 1;
  for (size_t n = lens.size(); n >= 10; n /= 10) ++width;
  for (size_t i = 0; i < lens.size(); ++i) {
    std::string name = prefix;
    if (lens.size() != 1) { const std::string digits = std::to_string(i + 1); name += std::string(digits.size() < (size_t)width ? (size_t)width - digits.size() : 0, '0') + digits; }
    ref.names.push_back(name);
    std::string s(lens[i], 'A');
    Rng rng(mix3(seed, 0x5EF, i));
    for (uint32_t p = 0; p < lens[i];) {
      uint64_t r = rng.next();
      for (int k = 0; k < 32 && p < lens[i]; ++k, ++p, r >>= 2) s[p] = "ACGT"[r & 3];
    }
    ref.seqs.push_back(std::move(s));
  }
}

void synth_reads(const SynthConfig& cfg, const RefSet& ref, BamHeader& hdr, ReadBatch& reads,
                 std::vector<SynthVariant>& variants) {
  Model m;
  m.cfg = &cfg; m.ref = &ref; m.variants = &variants;
  for (int q = 0; q < 64; ++q) m.perr_ppm[q] = (uint32_t)(pow(10.0, -q / 10.0) * 1e6 + 0.5);
  const size_t n_t = ref.seqs.size();
  uint64_t total = ref.total_length();
  std::vector<uint64_t> cum(n_t + 1, 0);
  for (size_t t = 0; t < n_t; ++t) cum[t + 1] = cum[t] + ref.seqs[t].size();
  auto locate = [&](uint64_t g, int32_t& tid, int32_t& pos) {
    size_t t = std::upper_bound(cum.begin(), cum.end(), g) - cum.begin() - 1;
    tid = (int32_t)t; pos = (int32_t)(g - cum[t]);
  };

  plant(cfg, ref, cum, variants, m.gaps);
  m.var_by_tid.assign(n_t, {});
  {
    std::vector<uint32_t> order(variants.size());
    for (uint32_t i = 0; i < order.size(); ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
      return variants[a].tid != variants[b].tid ? variants[a].tid < variants[b].tid : variants[a].pos0 < variants[b].pos0; });
    for (uint32_t i : order) m.var_by_tid[(size_t)variants[i].tid].push_back(i);
  }

  // header
  hdr.text = "@HD\tVN:1.6\tSO:coordinate\n";
  hdr.target_names = ref.names;
  hdr.target_lens.clear();
  for (size_t t = 0; t < n_t; ++t) {
    hdr.target_lens.push_back((uint32_t)ref.seqs[t].size());
    hdr.text += "@SQ\tSN:" + ref.names[t] + "\tLN:" + std::to_string(ref.seqs[t].size()) + "\n";
  }
  hdr.read_groups = ReadGroups();
  for (const SynthReadSet& s : cfg.sets) {
    hdr.text += "@RG\tID:" + s.name + "\tLB:" + s.name + "\tSM:" + s.name + "\n";
    hdr.read_groups.ids.push_back(s.name);
    hdr.read_groups.libraries.push_back(s.name);
  }

  // fragments, generated in parallel over contiguous fragment ranges
  int threads = std::max(1, cfg.threads);
  const bool windowed = cfg.window_hi > cfg.window_lo;
  uint64_t reach = 0;  // how far before the window a fragment that still overlaps it can start
  for (const SynthReadSet& s : cfg.sets) reach = std::max<uint64_t>(reach, (uint64_t)(s.paired ? s.frag_mean + 10 * s.frag_sd : 0) + 2 * s.read_len + 64);
  const uint64_t win_lo = cfg.window_lo > reach ? cfg.window_lo - reach : 0, win_hi = cfg.window_hi;
  std::vector<std::vector<TmpRead>> parts((size_t)threads * cfg.sets.size());
  for (size_t si = 0; si < cfg.sets.size(); ++si) {
    const SynthReadSet& s = cfg.sets[si];
    uint64_t n_frag = (uint64_t)(s.coverage * (double)total / ((double)s.read_len * (s.paired ? 2 : 1)) + 0.5);
    auto work = [&, si, n_frag](int t) {
      std::vector<TmpRead>& out = parts[si * (size_t)threads + (size_t)t];
      uint64_t lo = n_frag * (uint64_t)t / (uint64_t)threads, hi = n_frag * (uint64_t)(t + 1) / (uint64_t)threads;
      out.reserve((size_t)((hi - lo) * (s.paired ? 2 : 1)));
      TmpRead r;
      for (uint64_t f = lo; f < hi; ++f) {
        uint64_t gid = ((uint64_t)si << 40) | f;
        Rng rng(mix3(cfg.seed, 0xF4A6 + si, f));
        const uint64_t g = rng.next() % total;
        if (windowed && (g < win_lo || g >= win_hi)) continue;
        int32_t tid, pos; locate(g, tid, pos);
        int32_t tl = (int32_t)ref.seqs[(size_t)tid].size();
        bool flip = rng.below(2);
        if (!s.paired) {
          if (pos + (int32_t)s.read_len + 8 > tl) continue;
          if (!make_read(m, tid, pos, flip, s.read_len, gid, rng, r)) continue;
          r.rg = (uint8_t)si; r.frag = gid; r.mate = 0;
          out.push_back(r);
        } else {
          int32_t F = (int32_t)floor(s.frag_mean + s.frag_sd * rng.normal() + 0.5);
          if (F < (int32_t)s.read_len) F = (int32_t)s.read_len;
          if (pos + F + 8 > tl) continue;
          // left mate forward, right mate reverse; `flip` decides which of them is read 1
          TmpRead a, b;
          bool ok_a = make_read(m, tid, pos, false, s.read_len, gid, rng, a);
          bool ok_b = make_read(m, tid, pos + F - (int32_t)s.read_len, true, s.read_len, gid, rng, b);
          if (ok_a) { a.flag |= 1 | 2 | 32 | (flip ? 128 : 64); a.rg = (uint8_t)si; a.frag = gid; a.mate = 0; out.push_back(a); }
          if (ok_b) { b.flag |= 1 | 2 | (flip ? 64 : 128); b.rg = (uint8_t)si; b.frag = gid; b.mate = 1; out.push_back(b); }
        }
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& t : pool) t.join();
  }

  // coordinate sort with a total, thread-count independent tie order
  std::vector<const TmpRead*> all;
  for (auto& v : parts) for (auto& r : v) all.push_back(&r);
  std::sort(all.begin(), all.end(), [](const TmpRead* a, const TmpRead* b) {
    if (a->tid != b->tid) return a->tid < b->tid;
    if (a->pos != b->pos) return a->pos < b->pos;
    if (a->frag != b->frag) return a->frag < b->frag;
    return a->mate < b->mate;
  });
  reads = ReadBatch();
  size_t n = all.size(), nb = 0, nc = 0;
  for (const TmpRead* r : all) { nb += r->bases.size(); nc += r->cigar.size(); }
  reads.bases.reserve(nb); reads.quals.reserve(nb); reads.cigars.reserve(nc);
  for (auto* v : {&reads.tid, &reads.pos, &reads.xl, &reads.xr, &reads.as}) v->reserve(n);
  for (const TmpRead* r : all) {
    reads.tid.push_back(r->tid); reads.pos.push_back(r->pos); reads.flag.push_back(r->flag); reads.mapq.push_back(r->mapq);
    reads.rg.push_back(r->rg); reads.x1.push_back(r->x1); reads.xl.push_back(r->xl); reads.xr.push_back(r->xr);
    reads.as.push_back(r->as); reads.l_seq.push_back((uint32_t)r->bases.size()); reads.seq_off.push_back(reads.bases.size());
    reads.n_cigar.push_back((uint32_t)r->cigar.size()); reads.cigar_off.push_back(reads.cigars.size());
    reads.bases.insert(reads.bases.end(), r->bases.begin(), r->bases.end());
    reads.quals.insert(reads.quals.end(), r->quals.begin(), r->quals.end());
    reads.cigars.insert(reads.cigars.end(), r->cigar.begin(), r->cigar.end());
  }
}

std::vector<uint64_t> synth_shard_bounds(const SynthConfig& cfg, const RefSet& ref, uint32_t n_shards) {
  const uint64_t total = ref.total_length();
  std::vector<uint64_t> bounds(n_shards + 1, 0);
  bounds[n_shards] = total;
  if (n_shards <= 1 || !total) return bounds;
  const size_t n_t = ref.seqs.size();
  std::vector<uint64_t> cum(n_t + 1, 0);
  for (size_t t = 0; t < n_t; ++t) cum[t + 1] = cum[t] + ref.seqs[t].size();
  std::vector<SynthVariant> variants;
  std::vector<Gap> gaps;
  plant(cfg, ref, cum, variants, gaps);
  // Aligned bases by the bin their read starts in, from the first draws of every fragment's generator (position, strand,
  // fragment length: what synth_reads draws before it makes a read), with its rejections at contig ends and sample gaps.
  const uint64_t bin = 64, n_bins = (total + bin - 1) / bin;
  std::vector<uint64_t> w(n_bins + 1, 0);
  const int threads = std::max(1, cfg.threads);
  for (size_t si = 0; si < cfg.sets.size(); ++si) {
    const SynthReadSet& s = cfg.sets[si];
    const uint64_t n_frag = (uint64_t)(s.coverage * (double)total / ((double)s.read_len * (s.paired ? 2 : 1)) + 0.5);
    std::vector<std::vector<uint64_t>> part((size_t)threads, std::vector<uint64_t>(n_bins + 1, 0));
    auto work = [&](int t) {
      std::vector<uint64_t>& mine = part[(size_t)t];
      auto add_read = [&](int32_t tid, int32_t pos) {
        for (const Gap& g : gaps) if (g.tid == tid && pos < g.end && pos + (int32_t)s.read_len > g.beg) return;
        mine[(cum[(size_t)tid] + (uint64_t)pos) / bin] += s.read_len;
      };
      for (uint64_t f = n_frag * (uint64_t)t / (uint64_t)threads; f < n_frag * (uint64_t)(t + 1) / (uint64_t)threads; ++f) {
        Rng rng(mix3(cfg.seed, 0xF4A6 + si, f));
        const uint64_t g = rng.next() % total;
        const size_t ti = std::upper_bound(cum.begin(), cum.end(), g) - cum.begin() - 1;
        const int32_t tid = (int32_t)ti, pos = (int32_t)(g - cum[ti]), tl = (int32_t)ref.seqs[ti].size();
        rng.below(2);
        if (!s.paired) {
          if (pos + (int32_t)s.read_len + 8 > tl) continue;
          add_read(tid, pos);
        } else {
          int32_t F = (int32_t)floor(s.frag_mean + s.frag_sd * rng.normal() + 0.5);
          if (F < (int32_t)s.read_len) F = (int32_t)s.read_len;
          if (pos + F + 8 > tl) continue;
          add_read(tid, pos);
          add_read(tid, pos + F - (int32_t)s.read_len);
        }
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& t : pool) t.join();
    for (auto& v : part) for (uint64_t b = 0; b < n_bins; ++b) w[b] += v[b];
  }
  // a read of L bases starting in bin b lays its bases over the next L columns: spread the weights before cutting
  uint32_t max_len = 1;
  for (const SynthReadSet& s : cfg.sets) max_len = std::max(max_len, s.read_len);
  const uint64_t spread = std::max<uint64_t>(1, max_len / bin);
  std::vector<double> d(n_bins + spread + 1, 0.0);
  for (uint64_t b = 0; b < n_bins; ++b) for (uint64_t k = 0; k < spread; ++k) d[b + k] += (double)w[b] / (double)spread;
  double all = 0.0;
  for (uint64_t b = 0; b < n_bins; ++b) all += d[b];
  double acc = 0.0;
  uint64_t b = 0;
  for (uint32_t k = 1; k < n_shards; ++k) {
    const double want = all * (double)k / (double)n_shards;
    while (b < n_bins && acc + d[b] <= want) acc += d[b++];
    bounds[k] = std::min(total, b * bin);
  }
  for (uint32_t k = 1; k <= n_shards; ++k) if (bounds[k] < bounds[k - 1]) bounds[k] = bounds[k - 1];
  return bounds;
}

}  // namespace brq
