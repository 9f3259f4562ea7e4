#include "finalize.h"
#include "stats_math.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <limits>
#include <map>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <thread>

namespace brq {

// ============================================================================== covariates
static const char* kCovNames[COV_COUNT] = {"read_set", "ref_base", "prev_ref_base", "obs_base", "quality", "read_pos", "base_repeat"};

CovSpec parse_covariates(const std::string& s) {
  CovSpec c;
  size_t p = 0;
  while (p <= s.size()) {
    size_t e = s.find(',', p);
    if (e == std::string::npos) e = s.size();
    std::string item = s.substr(p, e - p), key = item, val;
    size_t eq = item.find('=');
    if (eq != std::string::npos) { key = item.substr(0, eq); val = item.substr(eq + 1); }
    auto num = [&]() { return (uint32_t)atoi(val.c_str()); };
    if (key == "ref_base") { c.used[COV_REF_BASE] = true; c.maxv[COV_REF_BASE] = 5; }
    else if (key == "obs_base") { c.used[COV_OBS_BASE] = true; c.maxv[COV_OBS_BASE] = 5; }
    else if (key == "quality") { c.used[COV_QUALITY] = true; c.maxv[COV_QUALITY] = num(); }
    else if (key == "read_set") { c.used[COV_READ_SET] = true; c.maxv[COV_READ_SET] = num(); }
    else if (key == "read_pos") { c.used[COV_READ_POS] = true; c.maxv[COV_READ_POS] = num(); }
    else if (key == "base_repeat") { c.used[COV_BASE_REPEAT] = true; c.maxv[COV_BASE_REPEAT] = num(); c.clamp[COV_BASE_REPEAT] = true; }
    else if (key == "ref_pos") c.per_position = true;
    else if (key == "prev_base") throw std::runtime_error("covariate prev_base is declared by the reference but never populated; not supported");
    else if (!key.empty()) throw std::runtime_error("Unrecognized covariate: " + key);
    p = e + 1;
  }
  uint32_t cur = 1;
  for (int i = 0; i < COV_COUNT; ++i) if (c.used[i]) {
    if (c.maxv[i] == 0) throw std::runtime_error(std::string("covariate needs a maximum: ") + kCovNames[i]);
    c.offset[i] = cur; cur *= c.maxv[i];
  }
  c.n_bins = cur;
  return c;
}

std::string CovSpec::text() const {
  std::string s;
  if (per_position) s += "ref_pos";
  for (int i = 0; i < COV_COUNT; ++i) {
    if (!used[i]) continue;
    if (!s.empty()) s += ",";
    s += kCovNames[i];
    if (i != COV_REF_BASE && i != COV_OBS_BASE) s += "=" + std::to_string(maxv[i]);
  }
  return s;
}

CovLayout to_layout(const CovSpec& c) {
  CovLayout l;
  l.off_set = c.offset[COV_READ_SET]; l.off_ref = c.offset[COV_REF_BASE]; l.off_obs = c.offset[COV_OBS_BASE];
  l.off_qual = c.offset[COV_QUALITY]; l.off_rpos = c.offset[COV_READ_POS]; l.off_rep = c.offset[COV_BASE_REPEAT];
  const uint32_t inf = 0xFFFFFFFFu;
  l.max_set = c.used[COV_READ_SET] ? c.maxv[COV_READ_SET] : inf;
  l.max_qual = c.used[COV_QUALITY] ? c.maxv[COV_QUALITY] : inf;
  l.max_rpos = c.used[COV_READ_POS] ? c.maxv[COV_READ_POS] : inf;
  l.max_rep = c.used[COV_BASE_REPEAT] ? c.maxv[COV_BASE_REPEAT] : inf;
  l.n_bins = c.n_bins;
  l.obs_used = c.used[COV_OBS_BASE];
  return l;
}

// ============================================================================== formatting
std::string format_default(double v) {
  char buf[64];
  snprintf(buf, sizeof buf, "%.6g", v);
  return buf;
}

std::string format_double(double v, uint32_t precision, bool scientific) {
  if (std::isnan(v)) return "NA";
  char buf[512];
  snprintf(buf, sizeof buf, scientific ? "%.*e" : "%.*f", (int)precision, v);
  std::string s(buf);
  if (scientific && s.size() >= 3 && s[s.size() - 3] == '0') s.erase(s.size() - 3, 1);  // 3-digit exponents
  return s;
}

static void write_table_rows(std::ostream& out, const CovSpec& c, size_t n, const std::function<std::string(size_t)>& value) {
  for (size_t idx = 0; idx < n; ++idx) {
    for (int i = 0; i < COV_COUNT; ++i) {
      if (!c.used[i]) continue;
      uint32_t j = ((uint32_t)idx / c.offset[i]) % c.maxv[i];
      if (i == COV_REF_BASE || i == COV_OBS_BASE) out << index_to_char((uint8_t)j) << '\t';
      else out << j << '\t';
    }
    out << value(idx) << '\n';
  }
}

void write_error_rates(const std::string& path, const CovSpec& c, const std::vector<double>& t) {
  std::ofstream out(path.c_str());
  if (!out) throw std::runtime_error("cannot create " + path);
  out << c.text() << '\n';
  for (int i = 0; i < COV_COUNT; ++i) if (c.used[i]) out << kCovNames[i] << '\t';
  out << "log10_probability\n";
  write_table_rows(out, c, t.size(), [&](size_t i) { return format_default(t[i]); });
}

void write_count_table(const std::string& path, const CovSpec& c, const std::vector<uint64_t>& counts) {
  std::ofstream out(path.c_str());
  if (!out) throw std::runtime_error("cannot create " + path);
  out << c.text() << '\n';
  for (int i = 0; i < COV_COUNT; ++i) if (c.used[i]) out << kCovNames[i] << '\t';
  out << "count\n";
  write_table_rows(out, c, counts.size(), [&](size_t i) { return std::to_string(counts[i]); });
}

// The count table of a covariate string that names ref_pos (error_count.cpp:105-111, 193-198, 803-846): the table is too big to
// keep per position, so the reference prints every position's non-empty bins as the pileup passes it and clears the table.
// Here: the positional histogram records of every column (hist_rec / hist_off of a stream on the host), both observations of
// each (kernels.cu: hist_record), counted into the column's bins and printed in bin order.
void write_count_table_per_position(const std::string& path, const CovSpec& c, const PileupStream& st) {
  std::ofstream out(path.c_str());
  if (!out) throw std::runtime_error("cannot create " + path);
  out << c.text() << '\n';
  out << "ref_pos\t";
  for (int i = 0; i < COV_COUNT; ++i) if (c.used[i]) out << kCovNames[i] << '\t';
  out << "count\n";
  if (!st.hist_rec || !st.hist_off) throw std::runtime_error("the per-position count table needs the histogram records on the host");
  const CovLayout lay = to_layout(c);
  const bool wide = st.hist_bytes == 8;
  if (!wide && (lay.off_rpos || lay.off_rep)) throw std::runtime_error("the stream was staged without read_pos / base_repeat (brq_stage_options.use_read_pos, use_base_repeat)");
  std::vector<uint32_t> bins(c.n_bins, 0), touched;
  const uint8_t* rec = static_cast<const uint8_t*>(st.hist_rec);
  auto check = [&](int cov, uint32_t v) {
    if (c.used[cov] && !c.clamp[cov] && v >= c.maxv[cov])
      throw std::runtime_error(std::string("Covariate '") + kCovNames[cov] + "' with value '" + std::to_string(v) + "' exceeded enforced maximum value of '" +
                               std::to_string(c.maxv[cov] - 1) + "'.");
  };
  for (const Segment& sg : st.segments) {
    for (int32_t col = sg.lo; col < sg.hi; ++col) {
      const uint64_t slot = sg.slot0 + (uint64_t)(col - sg.lo);
      const uint64_t r0 = st.hist_off[slot] & ~HIST_OFF_REDUNDANT_BIT, r1 = st.hist_off[slot + 1] & ~HIST_OFF_REDUNDANT_BIT;
      touched.clear();
      for (uint64_t r = r0; r < r1; ++r) {
        uint32_t lo, hi = 0;
        memcpy(&lo, rec + r * st.hist_bytes, 4);
        if (wide) memcpy(&hi, rec + r * st.hist_bytes + 4, 4);
        const uint32_t set = ((lo >> HR_SET) & 7u) | (wide ? (hi >> (HR_SET_HI - 32)) << 3 : 0u), rpos = wide ? (hi & 0xFFFFu) : 0u;
        check(COV_READ_SET, set);
        uint32_t base = set * lay.off_set;
        if (lay.off_rpos) { check(COV_READ_POS, rpos); base += rpos * lay.off_rpos; }
        auto observe = [&](uint32_t ref, uint32_t obs, uint32_t qual, uint32_t rep) {
          check(COV_QUALITY, qual);
          uint32_t idx = base + ref * lay.off_ref + obs * lay.off_obs + qual * lay.off_qual;
          if (lay.off_rep) idx += std::min(rep, lay.max_rep - 1) * lay.off_rep;
          if (bins[idx]++ == 0) touched.push_back(idx);
        };
        if (lo & (1u << HR_VALIDA)) observe(lo & 7u, (lo >> HR_OBSA) & 7u, (lo >> HR_QUALA) & 127u, wide ? (hi >> (HR_REPA - 32)) & 255u : 0u);
        if (lo & (1u << HR_VALIDB)) observe((lo >> HR_REFB) & 7u, (lo >> HR_OBSB) & 7u, (lo >> HR_QUALB) & 127u, wide ? (hi >> (HR_REPB - 32)) & 31u : 0u);
      }
      std::sort(touched.begin(), touched.end());
      for (uint32_t idx : touched) {
        out << (col + 1) << '\t';
        for (int i = 0; i < COV_COUNT; ++i) {
          if (!c.used[i]) continue;
          const uint32_t j = (idx / c.offset[i]) % c.maxv[i];
          if (i == COV_REF_BASE || i == COV_OBS_BASE) out << index_to_char((uint8_t)j) << '\t';
          else out << j << '\t';
        }
        out << bins[idx] << '\n';
        bins[idx] = 0;
      }
    }
  }
}

void read_error_rates(const std::string& path, CovSpec& c, std::vector<double>& t) {
  std::ifstream in(path.c_str());
  if (!in) throw std::runtime_error("cannot open " + path);
  std::string line;
  std::getline(in, line);
  c = parse_covariates(line);
  std::getline(in, line);  // column names: order is fixed, ignored
  t.assign(c.n_bins, 0.0);
  for (uint32_t i = 0; i < c.n_bins; ++i) {
    std::getline(in, line);
    size_t tab = line.rfind('\t');
    t[i] = strtod(line.c_str() + (tab == std::string::npos ? 0 : tab + 1), nullptr);
  }
}

struct WorkerPool::Impl {
  std::mutex m;
  std::condition_variable wake, done;
  std::vector<std::thread> threads;
  const std::function<void(size_t, size_t)>* job = nullptr;
  uint64_t generation = 0;
  size_t pending = 0;
  bool stop = false;
  std::exception_ptr failure;
};

WorkerPool::WorkerPool(size_t n_threads) : impl_(new Impl), n_(n_threads) {
  for (size_t t = 0; t < n_; ++t)
    impl_->threads.emplace_back([this, t] {
      uint64_t seen = 0;
      for (;;) {
        const std::function<void(size_t, size_t)>* job;
        {
          std::unique_lock<std::mutex> lk(impl_->m);
          impl_->wake.wait(lk, [&] { return impl_->stop || impl_->generation != seen; });
          if (impl_->stop) return;
          seen = impl_->generation; job = impl_->job;
        }
        try { (*job)(t + 1, n_ + 1); }
        catch (...) { std::lock_guard<std::mutex> g(impl_->m); if (!impl_->failure) impl_->failure = std::current_exception(); }
        { std::lock_guard<std::mutex> g(impl_->m); if (--impl_->pending == 0) impl_->done.notify_one(); }
      }
    });
}

WorkerPool::~WorkerPool() {
  { std::lock_guard<std::mutex> g(impl_->m); impl_->stop = true; }
  impl_->wake.notify_all();
  for (auto& t : impl_->threads) t.join();
  delete impl_;
}

void WorkerPool::run(const std::function<void(size_t, size_t)>& job) {
  { std::lock_guard<std::mutex> g(impl_->m); impl_->job = &job; impl_->pending = n_; impl_->failure = nullptr; ++impl_->generation; }
  impl_->wake.notify_all();
  job(0, n_ + 1);
  std::unique_lock<std::mutex> lk(impl_->m);
  impl_->done.wait(lk, [&] { return impl_->pending == 0; });
  if (impl_->failure) std::rethrow_exception(impl_->failure);
}

void canonicalise_table(const std::vector<double>& log10_prob, std::vector<double>& text, std::vector<double>& prob, WorkerPool* pool) {
  const size_t n = log10_prob.size();
  text.resize(n);
  prob.resize(n);
  auto run = [&](size_t lo, size_t hi) {
    char buf[64];
    for (size_t i = lo; i < hi; ++i) {
      snprintf(buf, sizeof buf, "%.6g", log10_prob[i]);  // format_default
      text[i] = strtod(buf, nullptr);
      prob[i] = pow(10, text[i]);
    }
  };
  // every entry is independent
  if (!pool || n < 512) { run(0, n); return; }
  pool->run([&](size_t part, size_t parts) { run(n * part / parts, n * (part + 1) / parts); });
}

void canonicalise_table(const std::vector<double>& log10_prob, std::vector<double>& text, std::vector<double>& prob) {
  canonicalise_table(log10_prob, text, prob, nullptr);
}

// error_count.cpp:697-785, index arithmetic restated literally (accumulate obs-major, read out b1*5+b2).
void write_base_qual_tables(const std::string& pattern, const CovSpec& c, const std::vector<uint64_t>& counts,
                            const std::vector<std::string>& readfiles) {
  const uint32_t nS = c.maxv[COV_READ_SET], nO = c.maxv[COV_OBS_BASE], nR = c.maxv[COV_REF_BASE], nQ = c.maxv[COV_QUALITY];
  const uint32_t per_set = nO * nR * nQ;
  std::vector<double> t((size_t)nS * per_set, 0.0);
  for (uint32_t idx = 0; idx < counts.size(); ++idx) {
    uint32_t at = 0;
    for (int i = 0; i < COV_COUNT; ++i) {
      if (!c.used[i]) continue;
      uint32_t j = (idx / c.offset[i]) % c.maxv[i];
      if (i == COV_READ_SET) at += per_set * j;
      else if (i == COV_REF_BASE) at += j;
      else if (i == COV_OBS_BASE) at += nR * j;
      else if (i == COV_QUALITY) at += nO * nR * j;
    }
    t[at] += (double)counts[idx];
  }
  double running = 0;
  for (uint32_t r = 0; r < nS; ++r) {
    for (uint32_t k = 0; k < per_set; ++k) {
      uint32_t i = r * per_set + k;
      running += t[i];
      if (i % nR == nR - 1) {
        for (uint32_t j = i - (nR - 1); j <= i; ++j) {
          if (running > 0) t[j] /= running;
          if (t[j] == 0) t[j] = std::numeric_limits<double>::quiet_NaN();
        }
        running = 0;
      }
    }
    if (r >= readfiles.size()) throw std::runtime_error("fewer read file names than read_set values");
    std::string fn = pattern;
    size_t h = fn.find('#');
    if (h != std::string::npos) fn.replace(h, 1, readfiles[r]);
    std::ofstream out(fn.c_str());
    if (!out) throw std::runtime_error("cannot create " + fn);
    out << "quality";
    for (uint32_t b1 = 0; b1 < nR; ++b1) for (uint32_t b2 = 0; b2 < nO; ++b2) out << '\t' << index_to_char((uint8_t)b1) << index_to_char((uint8_t)b2);
    out << '\n';
    for (uint32_t q = 0; q < nQ; ++q) {
      out << q;
      for (uint32_t b1 = 0; b1 < nR; ++b1) for (uint32_t b2 = 0; b2 < nO; ++b2) {
        double v = t[(size_t)r * per_set + q * nO * nR + b1 * nR + b2];
        out << '\t' << (std::isnan(v) ? std::string("NA") : format_default(v));
      }
      out << '\n';
    }
  }
}

void write_coverage_distributions(const std::string& dir, const std::vector<uint64_t>& cov, uint64_t stride, uint64_t n_groups) {
  for (uint64_t g = 0; g < n_groups; ++g) {
    std::string fn = dir + "/" + std::to_string(g) + ".unique_only_coverage_distribution.tab";
    std::ofstream out(fn.c_str());
    if (!out) throw std::runtime_error("cannot create " + fn);
    out << "coverage\tn\n";
    uint64_t last = 0;  // the reference's vector grows to the largest depth seen (error_count.cpp:182-185)
    for (uint64_t j = 0; j < stride; ++j) if (cov[g * stride + j]) last = j;
    for (uint64_t j = 1; j <= last; ++j) out << j << '\t' << cov[g * stride + j] << '\n';
  }
}

// ============================================================================== class table
// Geometry of every likelihood table of pass 2 (no transcendental math): MAPQ slots, the dominant
// MAPQ, the shared-memory table of the fit kernel and the quality window / copy count of the tally
// kernel's table.  Cheap, so it runs on every table installation; the values are filled in either on
// the device (build_tables_kernel, tables.cu) or on the host (build_class_lut and friends below).
void score_geometry(const CovSpec& c, const uint32_t mapq_seen[8], const ScoreGeometry& sg, ScoreParams& p, TableGeometry& g) {
  const bool wide = c.used[COV_READ_POS] || c.used[COV_BASE_REPEAT];
  if (wide && sg.side_stride != 2)
    throw std::runtime_error("the stream was staged without read_pos / base_repeat (brq_stage_options.use_read_pos, use_base_repeat)");
  if (!c.used[COV_OBS_BASE] || !c.used[COV_REF_BASE] || !c.used[COV_QUALITY])
    throw std::runtime_error("scoring needs ref_base, obs_base and quality covariates");
  const uint32_t n_set = c.used[COV_READ_SET] ? c.maxv[COV_READ_SET] : 1, Q = c.maxv[COV_QUALITY];
  memset(p.mapq_slot, 255, sizeof p.mapq_slot);
  g.mapqs.clear();
  for (uint32_t m = 0; m < 256; ++m) if (mapq_seen[m >> 5] >> (m & 31) & 1) { p.mapq_slot[m] = (uint8_t)g.mapqs.size(); g.mapqs.push_back(m); }
  if (g.mapqs.empty()) { p.mapq_slot[0] = 0; g.mapqs.push_back(0); }
  p.n_mapq_slots = (uint32_t)g.mapqs.size();
  p.max_qual = Q;
  p.max_set = c.used[COV_READ_SET] ? n_set : 32;
  g.off_set = c.used[COV_READ_SET] ? c.offset[COV_READ_SET] : 0; g.off_ref = c.offset[COV_REF_BASE];
  g.off_obs = c.offset[COV_OBS_BASE]; g.off_qual = c.offset[COV_QUALITY];
  g.off_rpos = c.used[COV_READ_POS] ? c.offset[COV_READ_POS] : 0; g.off_rep = c.used[COV_BASE_REPEAT] ? c.offset[COV_BASE_REPEAT] : 0;
  p.n_rpos = c.used[COV_READ_POS] ? c.maxv[COV_READ_POS] : 1; p.n_rep = c.used[COV_BASE_REPEAT] ? c.maxv[COV_BASE_REPEAT] : 1;
  const size_t W = (size_t)p.n_rpos * p.n_rep;  // classes per (set, strand, MAPQ, quality, obs) with read_pos / base_repeat
  if (n_set * 2 < sg.n_st)
    throw std::runtime_error("Covariate 'read_set' with value '" + std::to_string(sg.n_st / 2 - 1) +
                             "' exceeded enforced maximum value of '" + std::to_string(n_set - 1) + "'.");
  g.n_st = sg.n_st;  // the stream's words index the shared table with the stream's own set count
  g.n_lut = (size_t)g.n_st * g.mapqs.size() * Q * W * 5;
  if (g.n_lut >= (1ull << 31)) throw std::runtime_error("too many record classes (read sets x MAPQ values x quality x read_pos x base_repeat)");
  // the dominant MAPQ and the tally kernel's shared-memory window were fixed when the stream was staged
  if (p.mapq_slot[sg.hot_mapq] == 255) throw std::runtime_error("the stream's dominant MAPQ is missing from its MAPQ set");
  p.hot_mapq = sg.hot_mapq;
  const size_t n_hot = (size_t)g.n_st * Q * 5;
  p.n_hot = (!wide && n_hot * 48 <= 96 * 1024) ? (uint32_t)n_hot : 0;  // fit kernel: three CTAs per SM must each hold a copy
  g.n_hotR = n_hot;
  // MAPQ range of the global table of the tally kernel
  p.mq_min = g.mapqs.front(); p.n_mq = g.mapqs.back() - g.mapqs.front() + 1;
  g.n_cold = (size_t)g.n_st * p.n_mq * Q * W * 5;
  p.t_qlo = sg.q_lo; p.t_nq = sg.n_q; p.t_nsq = sg.n_sq(); p.t_nw = sg.n_words();
  p.t_stride = p.t_nsq * 64u;
  g.n_tally_cells = (size_t)5 * p.t_stride / 16;  // five observation planes
}

// Host copy of the per-class terms, with the libm calls the reference makes (identify_mutations.cpp:3359-3384).
// Only the host re-evaluation of flagged slots reads it (write_evidence), and only for the classes those slots hold:
// an entry is computed the first time it is asked for.
void ClassLut::reset(const CovSpec& c, const std::vector<double>& prob_, const ScoreParams& p_, const TableGeometry& g_) {
  spec = &c; prob = &prob_; p = &p_; g = &g_;
  terms.assign(g_.n_lut, ClassTerms());
  done.assign(g_.n_lut, 0);
}

const ClassTerms* ClassLut::get(size_t li) {
  ClassTerms& t = terms[li];
  if (done[li]) return &t;
  const uint32_t Q = p->max_qual;
  const size_t W = (size_t)p->n_rpos * p->n_rep;
  const uint32_t obs = (uint32_t)(li % 5), rr = (uint32_t)((li / 5) % W), q = (uint32_t)((li / (5 * W)) % Q);
  const size_t ms = (li / (5 * W * Q)) % g->mapqs.size();
  const uint32_t st = (uint32_t)(li / (5 * W * Q * g->mapqs.size()));
  const uint32_t rpos = rr / p->n_rep, rpt = rr % p->n_rep;
  auto comp = [](uint32_t b) { return b < 4 ? 3 - b : 4u; };
  const uint32_t set = st >> 1, top = st & 1;
  const double incorrect = pow(10, -(double)g->mapqs[ms] / 10);
  const double correct = 1 - incorrect;
  const double uniform = 1.0 / 5.0;
  const uint32_t o = top ? obs : comp(obs);
  double mx = -std::numeric_limits<double>::max();
  for (uint32_t b = 0; b < 5; ++b) {
    const uint32_t rf = top ? b : comp(b);
    const uint32_t idx = set * g->off_set + rf * g->off_ref + o * g->off_obs + q * g->off_qual + rpos * g->off_rpos + rpt * g->off_rep;
    double pr = correct * (*prob)[idx] + incorrect * uniform;
    if (pr < 0.0) pr = 0.0;
    t.L[b] = log10(pr);
    mx = std::max(mx, t.L[b]);
  }
  for (uint32_t b = 0; b < 5; ++b) t.r[b] = pow(10, t.L[b] - mx);
  t.M = mx;
  t.r2 = 0.0;
  done[li] = 1;
  return &t;
}

// ============================================================================== optional per-position outputs
// identify_mutations.cpp:1693-1733: one line per (column, insert_count) in visit order:
//   position insert_count ref_base consensus_score  then per base "X (bottom/top)" of the scoring records and
//   "rX (bottom/top)" of the untrimmed redundant records.  Numbers go through the default ostream formatting.
void write_per_position_file(const std::string& path, const BamHeader& hdr, const PileupStream& st, const std::vector<ColumnOut>& cols,
                             uint32_t base_quality_cutoff, const std::vector<double>& deletion_propagation_cutoff) {
  std::ofstream out(path.c_str());
  if (!out) throw std::runtime_error("cannot create " + path);
  auto line = [&](uint64_t slot, uint32_t position, uint32_t insert_count) {
    uint32_t uniq[5][2] = {{0}}, red[5][2] = {{0}};
    for_each_classic(st, slot, [&](uint32_t r, uint32_t, uint32_t) {
      const uint32_t obs = r & 7u, top = (r & SR_TOP_BIT) ? 1u : 0u;
      if (obs > 4) return;
      if (!(r & SR_UNIQUE_BIT)) { if (!(r & SR_TRIM_BIT)) ++red[obs][top]; return; }
      if ((r & SR_TRIM_BIT) || !(r & SR_OK_BIT) || ((r >> SR_QUAL_SHIFT) & 127u) < base_quality_cutoff) return;
      ++uniq[obs][top];
    });
    out << position << ' ' << insert_count << ' ' << index_to_char(st.slot_ref[slot]) << ' ' << format_default(cols[slot].consensus_score);
    for (int b = 0; b < 5; ++b) out << ' ' << index_to_char((uint8_t)b) << " (" << uniq[b][0] << '/' << uniq[b][1] << ')';
    for (int b = 0; b < 5; ++b) out << " r" << index_to_char((uint8_t)b) << " (" << red[b][0] << '/' << red[b][1] << ')';
    out << '\n';
  };
  size_t ins_cursor = 0;
  for (const Segment& sg : st.segments) {
    if (deletion_propagation_cutoff[(size_t)sg.tid] < 0.0) continue;  // the callback returns before this point (:1319-1336)
    for (int32_t c = sg.lo; c < sg.hi; ++c) {
      const uint64_t slot = sg.slot0 + (uint64_t)(c - sg.lo);
      line(slot, (uint32_t)c + 1, 0);
      while (ins_cursor < st.n_ins && st.ins_parent[ins_cursor] < slot) ++ins_cursor;
      for (; ins_cursor < st.n_ins && st.ins_parent[ins_cursor] == slot; ++ins_cursor) line(st.n_base + ins_cursor, (uint32_t)c + 1, st.ins_count[ins_cursor]);
    }
  }
}

// identify_mutations.cpp:2028-2052, 2173-2204: <seq>.coverage.tsv (--predict-copy-number), one file per visited target:
//   position ref_base unique_cov redundant_cov total_cov, the sums over both strands, total as their plain sum.
void write_coverage_tsv(const std::string& pattern, const BamHeader& hdr, const RefSet& ref, const PileupStream& st, const std::vector<ColumnOut>& cols,
                        const std::vector<std::vector<CoverageColumn>>& by_group) {
  for (const Segment& sg : st.segments) {
    std::string fn = pattern;
    const std::string& name = hdr.target_names[(size_t)sg.tid];
    const size_t at = fn.find('@');
    if (at != std::string::npos) fn.replace(at, 1, name);
    std::ofstream out(fn.c_str(), sg.lo == 0 ? std::ios::out : std::ios::app);
    if (!out) throw std::runtime_error("cannot create " + fn);
    if (sg.lo == 0) {
      out << "position\tref_base\tunique_cov\tredundant_cov\ttotal_cov";
      for (size_t g = 0; g < by_group.size(); ++g) out << "\tRG-" << g << "_unique_cov\tRG-" << g << "_redundant_cov\tRG-" << g << "_total_cov";
      out << '\n';
    }
    size_t ri = 0;
    while (ri < ref.names.size() && ref.names[ri] != name) ++ri;
    for (int32_t c = sg.lo; c < sg.hi; ++c) {
      const ColumnOut& co = cols[sg.slot0 + (uint64_t)(c - sg.lo)];
      const double unique = (double)co.unique[0] + (double)co.unique[1], redundant = co.redundant[0] + co.redundant[1];
      // the reference prints the FASTA character of the column (reference_base_char_1), not the folded index
      const char rc = ri < ref.seqs.size() ? ref.seqs[ri][(size_t)c] : index_to_char(st.slot_ref[sg.slot0 + (uint64_t)(c - sg.lo)]);
      out << (c + 1) << '\t' << rc << '\t' << format_default(unique) << '\t' << format_default(redundant) << '\t' << format_default(unique + redundant);
      for (const std::vector<CoverageColumn>& grp : by_group) {   // position_coverage::sum(): bottom strand + top strand
        const CoverageColumn& k = grp[sg.slot0 + (uint64_t)(c - sg.lo)];
        const double gu = (double)k.unique[1] + (double)k.unique[0], gr = k.redundant[1] + k.redundant[0];
        out << '\t' << format_default(gu) << '\t' << format_default(gr) << '\t' << format_default(gu + gr);
      }
      out << '\n';
    }
  }
}

// ============================================================================== statistics
namespace {

double log_binomial(double n, double k) { return log_gamma(n + 1) - log_gamma(k + 1) - log_gamma(n - k + 1); }

// Two-sided Fisher exact test (stats.cpp:2144-2171).
double fisher_2x2(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  const uint32_t r1 = a + b, r2 = c + d, c1 = a + c, c2 = b + d, n = r1 + r2;
  const uint32_t lo = r1 > c2 ? r1 - c2 : 0, hi = std::min(r1, c1);
  const double denom = log_binomial(n, r1);
  const double observed = log_binomial(c1, a) + log_binomial(c2, r1 - a) - denom;
  const double slack = log(1.0 + 1e-7);
  double total = 0.0;
  for (uint32_t x = lo; x <= hi; ++x) {
    const double lp = log_binomial(c1, x) + log_binomial(c2, r1 - x) - denom;
    if (lp <= observed + slack) total += exp(lp);
  }
  return std::min(total, 1.0);
}

// One-sided two-sample KS test on integer-valued qualities, alternative "less"
// (stats.cpp:2191-2288): x = minor-allele qualities, y = major-allele qualities, given as counts
// per quality value.
double ks_less(const std::vector<uint32_t>& x_by_q, const std::vector<uint32_t>& y_by_q) {
  uint32_t nx = 0, ny = 0;
  for (uint32_t v : x_by_q) nx += v;
  for (uint32_t v : y_by_q) ny += v;
  std::vector<char> boundary((size_t)nx + ny + 1, 1);
  uint32_t cx = 0, cy = 0, best_x = 0, best_y = 0;
  double min_z = 0.0;
  bool have = false;
  size_t consumed = 0;
  for (size_t q = 0; q < x_by_q.size(); ++q) {
    const uint32_t g = x_by_q[q] + y_by_q[q];
    if (!g) continue;
    for (size_t k = consumed + 1; k < consumed + g; ++k) boundary[k] = 0;  // interior of a tie run
    consumed += g;
    cx += x_by_q[q]; cy += y_by_q[q];
    const double z = (double)cx / nx - (double)cy / ny;
    if (!have || z < min_z) { min_z = z; have = true; best_x = cx; best_y = cy; }
  }
  const double statistic = -min_z;
  if ((double)nx * ny < 10000) {
    const uint32_t m = ny, n = nx;
    const int64_t threshold = (int64_t)best_y * n - (int64_t)best_x * m;
    std::vector<double> row((size_t)n + 1, 0.0);  // lattice-path count, one row at a time
    for (uint32_t i = 0; i <= m; ++i) {
      for (uint32_t j = 0; j <= n; ++j) {
        double v;
        if (i == 0 && j == 0) v = 1.0;
        else v = (i > 0 ? row[j] : 0.0) + (j > 0 ? row[j - 1] : 0.0);
        const int64_t level = (int64_t)i * n - (int64_t)j * m;
        if (boundary[i + j] && level >= threshold) v = 0.0;
        row[j] = v;
      }
    }
    const double total_paths = exp(log_binomial(m + n, m));
    const double p = 1.0 - row[n] / total_paths;
    return std::min(std::max(p, 0.0), 1.0);
  }
  const double n_eff = ((double)nx * ny) / (nx + ny);
  return exp(-2.0 * statistic * statistic * n_eff);
}

}  // namespace

double fisher_strand_p_value(uint32_t minor_top, uint32_t minor_bottom, uint32_t major_top, uint32_t major_bottom) {
  return fisher_2x2(minor_top, minor_bottom, major_top, major_bottom);
}

namespace {

// ============================================================================== per-slot re-evaluation
// Everything below walks the slot's records in arrival order and accumulates exactly as the
// reference's per-read loops do, so the numbers printed into RA rows carry the reference's
// rounding.  Terms come from the same class table the device uses.
struct SlotEval {
  std::vector<const ClassTerms*> reads;  // scoring records, arrival order
  std::vector<uint8_t> obs, qual;
  uint32_t count[6][2];                  // [base][0 bottom, 1 top]
  double ll[5];
  uint8_t best = 5, major = 5, minor = 5, variant = 5;
  double consensus = std::numeric_limits<double>::quiet_NaN(), variant_score = std::numeric_limits<double>::quiet_NaN();
  double f[5] = {0, 0, 0, 0, 0}, log10_likelihood = 0.0;
  bool base_predicted = false, emit = false;
  uint32_t n() const { return (uint32_t)reads.size(); }
};

struct Fit { double f[5] = {0, 0, 0, 0, 0}; double ll = 0.0; };

Fit fit_ordered(const SlotEval& s, uint32_t allowed, double tol) {  // identify_mutations.cpp:3240-3318
  Fit m;
  const uint32_t n = s.n();
  if (!n) return m;
  double init_total = 0.0;
  for (int b = 0; b < 5; ++b) if (allowed >> b & 1) m.f[b] = 0.5;
  for (uint32_t i = 0; i < n; ++i) if (s.obs[i] < 5 && (allowed >> s.obs[i] & 1)) m.f[s.obs[i]] += 1.0;
  for (int b = 0; b < 5; ++b) init_total += m.f[b];
  for (int b = 0; b < 5; ++b) m.f[b] /= init_total;
  for (uint32_t it = 1; it <= 50; ++it) {
    double w[5] = {0, 0, 0, 0, 0}, ll = 0.0;
    for (uint32_t i = 0; i < n; ++i) {
      const ClassTerms& t = *s.reads[i];
      double sum = 0.0, mx = t.L[0];
      for (int b = 0; b < 5; ++b) { if (allowed >> b & 1) sum += m.f[b] * t.r[b]; mx = std::max(mx, t.L[b]); }
      if (sum > 0.0) {
        ll += log10(sum) + mx;
        for (int b = 0; b < 5; ++b) if (allowed >> b & 1) w[b] += m.f[b] * t.r[b] / sum;
      } else {
        for (int b = 0; b < 5; ++b) if (allowed >> b & 1) w[b] += m.f[b];
      }
    }
    double max_delta = 0.0;
    for (int b = 0; b < 5; ++b) {
      if (!(allowed >> b & 1)) continue;
      const double f_new = w[b] / (double)n;
      max_delta = std::max(max_delta, fabs(f_new - m.f[b]));
      m.f[b] = f_new;
    }
    m.ll = ll;
    if (max_delta < tol) break;
  }
  return m;
}

double profile_ll(const SlotEval& s, int variant, double f_fixed, double tol) {  // identify_mutations.cpp:3094-3143
  const uint32_t n = s.n();
  double f[5], others_total = 0.0;
  for (int b = 0; b < 5; ++b) if (b != variant) others_total += s.f[b];
  for (int b = 0; b < 5; ++b)
    f[b] = b == variant ? f_fixed : (others_total > 0.0 ? (1.0 - f_fixed) * s.f[b] / others_total : (1.0 - f_fixed) / 4.0);
  double ll = 0.0;
  for (int it = 0; it < 50; ++it) {
    double w[5] = {0, 0, 0, 0, 0};
    ll = 0.0;
    for (uint32_t i = 0; i < n; ++i) {
      const ClassTerms& t = *s.reads[i];
      double sum = 0.0, mx = t.L[0];
      for (int b = 0; b < 5; ++b) { sum += f[b] * t.r[b]; mx = std::max(mx, t.L[b]); }
      if (sum > 0.0) { ll += log10(sum) + mx; for (int b = 0; b < 5; ++b) w[b] += f[b] * t.r[b] / sum; }
      else for (int b = 0; b < 5; ++b) w[b] += f[b];
    }
    double others = 0.0;
    for (int b = 0; b < 5; ++b) if (b != variant) others += w[b];
    double max_delta = 0.0;
    for (int b = 0; b < 5; ++b) {
      if (b == variant) continue;
      const double f_new = others > 0.0 ? (1.0 - f_fixed) * w[b] / others : (1.0 - f_fixed) / 4.0;
      max_delta = std::max(max_delta, fabs(f_new - f[b]));
      f[b] = f_new;
    }
    if (max_delta < tol) break;
  }
  return ll;
}

void evaluate_slot(SlotEval& s, uint8_t ref, const EvidenceParams& ep) {
  const uint32_t n = s.n();
  for (int b = 0; b < 5; ++b) s.ll[b] = 0.0;
  for (uint32_t i = 0; i < n; ++i) for (int b = 0; b < 5; ++b) s.ll[b] += s.reads[i]->L[b];
  if (n) {  // pure_genotype_call, identify_mutations.cpp:3398-3433
    int best = 0;
    for (int b = 1; b < 5; ++b) if (s.ll[b] > s.ll[best]) best = b;
    double off = -std::numeric_limits<double>::max();
    for (int b = 0; b < 5; ++b) if (b != best) off = std::max(off, s.ll[b]);
    double tot = 0;
    for (int b = 0; b < 5; ++b) if (b != best) tot += pow(10, s.ll[b] - off);
    double lt = log10(tot);
    lt += off;
    s.best = (uint8_t)best;
    s.consensus = (s.ll[best] - lt) - ep.log10_ref_length;
  }
  s.base_predicted = s.consensus >= ep.mutation_cutoff;
  const bool passed_consensus = (s.best != ref) && !std::isnan(s.consensus) && s.consensus > 0;
  Fit full = fit_ordered(s, 0x1F, ep.precision_decimal);
  for (int b = 0; b < 5; ++b) s.f[b] = full.f[b];
  s.log10_likelihood = full.ll;
  bool passed_poly = false;
  if (n) {
    const double thr = 0.5 / (double)n;
    int mj = 0;
    for (int b = 1; b < 5; ++b) if (s.f[b] > s.f[mj]) mj = b;
    s.major = s.f[mj] > 0.0 ? (uint8_t)mj : 5;
    auto next = [&](uint8_t exclude) {
      uint8_t pick = 5;
      for (uint8_t b = 0; b < 5; ++b) {
        if (b == exclude || s.f[b] < thr) continue;
        if (pick == 5 || s.f[b] > s.f[pick]) pick = b;
      }
      return pick;
    };
    s.minor = next(s.major);
    s.variant = next(ref);
    if (s.variant != 5) {
      Fit null_fit = fit_ordered(s, 0x1F & ~(1u << s.variant), ep.precision_decimal);
      s.variant_score = (full.ll - null_fit.ll) - ep.log10_ref_length;
      if (s.variant_score >= ep.polymorphism_cutoff) passed_poly = true;
    }
  }
  s.emit = passed_consensus || passed_poly;
}

}  // namespace

// ============================================================================== shard serialisation
namespace {
struct Writer {
  std::string out;
  void u32(uint32_t v) { out.append(reinterpret_cast<const char*>(&v), 4); }
  void u64(uint64_t v) { out.append(reinterpret_cast<const char*>(&v), 8); }
  void str(const std::string& s) { u32((uint32_t)s.size()); out.append(s); }
};
struct Reader {
  const char* p; const char* end;
  void need(size_t n) { if ((size_t)(end - p) < n) throw std::runtime_error("truncated evidence shard"); }
  uint32_t u32() { need(4); uint32_t v; memcpy(&v, p, 4); p += 4; return v; }
  uint64_t u64() { need(8); uint64_t v; memcpy(&v, p, 8); p += 8; return v; }
  std::string str() { const uint32_t n = u32(); need(n); std::string s(p, n); p += n; return s; }
};
constexpr uint32_t SHARD_MAGIC = 0x31515242u;  // "BRQ1"
}  // namespace

std::string serialize_shard(const EvidenceShard& sh) {
  Writer w;
  w.u32(SHARD_MAGIC);
  w.u32((uint32_t)sh.target_names.size());
  for (size_t i = 0; i < sh.target_names.size(); ++i) { w.str(sh.target_names[i]); w.u32(sh.target_lens[i]); }
  w.u32((uint32_t)sh.segments.size());
  for (const auto& sg : sh.segments) { w.u32((uint32_t)sg.tid); w.u32((uint32_t)sg.lo); w.u32((uint32_t)sg.hi); }
  w.u64(sh.rechecked); w.u64(sh.overturned);
  w.u64(sh.events.size());
  for (const EvidenceEvent& e : sh.events) {
    w.u32(e.tid); w.u32(e.pos1); w.u32(e.unique); w.u32(e.packed); w.u32((uint32_t)e.rows.size());
    for (const GdRow& r : e.rows) {
      w.u32((uint32_t)r.type); w.str(r.seq_id); w.u64(r.a); w.u64(r.b); w.u64(r.c); w.u64(r.d); w.str(r.ref_base); w.str(r.new_base);
      w.u32((uint32_t)r.kv.size());
      for (const auto& kv : r.kv) { w.str(kv.first); w.str(kv.second); }
    }
  }
  return w.out;
}

EvidenceShard parse_shard(const void* data, size_t bytes) {
  Reader r{static_cast<const char*>(data), static_cast<const char*>(data) + bytes};
  if (r.u32() != SHARD_MAGIC) throw std::runtime_error("not an evidence shard");
  EvidenceShard sh;
  const uint32_t nt = r.u32();
  for (uint32_t i = 0; i < nt; ++i) { sh.target_names.push_back(r.str()); sh.target_lens.push_back(r.u32()); }
  const uint32_t ns = r.u32();
  for (uint32_t i = 0; i < ns; ++i) { EvidenceShard::Seg sg; sg.tid = (int32_t)r.u32(); sg.lo = (int32_t)r.u32(); sg.hi = (int32_t)r.u32(); sh.segments.push_back(sg); }
  sh.rechecked = r.u64(); sh.overturned = r.u64();
  const uint64_t ne = r.u64();
  for (uint64_t i = 0; i < ne; ++i) {
    EvidenceEvent e;
    e.tid = r.u32(); e.pos1 = r.u32(); e.unique = r.u32(); e.packed = r.u32();
    const uint32_t nr = r.u32();
    for (uint32_t k = 0; k < nr; ++k) {
      GdRow g;
      g.type = (int)r.u32(); g.seq_id = r.str(); g.a = r.u64(); g.b = r.u64(); g.c = r.u64(); g.d = r.u64(); g.ref_base = r.str(); g.new_base = r.str();
      const uint32_t nk = r.u32();
      for (uint32_t j = 0; j < nk; ++j) { std::string key = r.str(); g.kv[key] = r.str(); }
      e.rows.push_back(std::move(g));
    }
    sh.events.push_back(std::move(e));
  }
  return sh;
}

// ---- one shard's share of the evidence: the event columns with their target coordinates and the RA rows to emit at them
EvidenceShard collect_evidence(const BamHeader& hdr, const PileupStream& st, const std::vector<WalkEvent>& events_in,
                               const std::vector<uint32_t>& flagged_in, const std::vector<ColumnOut>& flagged_cols,
                               const ScoreParams& sp, ClassLut& lut, const EvidenceParams& ep, const FlaggedRecords* fr) {
  EvidenceCounts counts;
  // flagged slots in ascending order (the kernels append them in no particular order), each with its full result
  std::vector<uint32_t> order(flagged_in.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (uint32_t)i;
  std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return flagged_in[a] < flagged_in[b]; });
  std::vector<uint32_t> flagged(order.size());
  for (size_t i = 0; i < order.size(); ++i) flagged[i] = flagged_in[order[i]];

  // the columns the interval walk has to look at, in ascending order (the device appends them in no particular order)
  std::vector<WalkEvent> events(events_in);
  std::sort(events.begin(), events.end(), [](const WalkEvent& a, const WalkEvent& b) { return a.slot < b.slot; });

  // ---- re-evaluate flagged slots in arrival order
  struct Reval { bool base_predicted; bool emit; GdRow row; std::vector<GdRow> user_rows; };
  std::vector<Reval> reval(flagged.size());  // aligned with `flagged`
  // first the records of every flagged slot (the class table fills in on demand: not thread-safe), then the fits, the
  // profile-likelihood bounds and the bias tests of the slots side by side on a few threads: they are independent
  std::vector<SlotEval> evals(flagged.size());
  for (size_t fi = 0; fi < flagged.size(); ++fi) {
    const uint32_t slot = flagged[fi];
    SlotEval& s = evals[fi];
    memset(s.count, 0, sizeof s.count);
    auto take = [&](uint32_t r, uint32_t, uint32_t ext) {
      const uint32_t q = (r >> SR_QUAL_SHIFT) & 127;
      if (!(r & SR_UNIQUE_BIT) || (r & SR_TRIM_BIT) || !(r & SR_OK_BIT) || q < ep.base_quality_cutoff) return;
      const uint32_t obs = r & 7, top = (r & SR_TOP_BIT) ? 1 : 0, mapq = (r >> SR_MAPQ_SHIFT) & 255, set = (r >> SR_SET_SHIFT) & 31;
      const size_t rr = class_rr(ext, sp);
      const size_t li = ((((((size_t)set * 2 + top) * sp.n_mapq_slots + sp.mapq_slot[mapq]) * sp.max_qual + q) * sp.n_rpos * sp.n_rep) + rr) * 5 + obs;
      s.reads.push_back(lut.get(li));
      s.obs.push_back((uint8_t)obs); s.qual.push_back((uint8_t)q);
      ++s.count[obs][top];
    };
    if (fr) {  // the stream is in HBM only: the slot's records were gathered there (entry order[fi] of the unsorted list)
      const size_t e = order[fi];
      const uint32_t* w = fr->words + fr->word_off[e];
      for_each_classic_words(st.geo, fr->word_off[e + 1] - fr->word_off[e], [&](uint64_t j) { return w[j]; },
                             fr->side + (size_t)fr->side_off[e] * st.geo.side_stride, take);
    } else {
      for_each_classic(st, slot, take);
    }
  }
  // user evidence: which list entries meet which slot (the column itself for insert level 0, else its sub-column slot)
  std::map<uint64_t, std::vector<uint32_t>> user_at;
  for (const UserColumn& uc : st.user_columns) {
    if (uc.slot == ~0ull) continue;
    for (const auto& lv : uc.consumed) {
      uint64_t slot = uc.slot;
      if (lv.first > 0) {
        const size_t j0 = std::lower_bound(st.ins_parent.begin(), st.ins_parent.end(), uc.slot) - st.ins_parent.begin();
        slot = ~0ull;
        for (size_t j = j0; j < st.ins_parent.size() && st.ins_parent[j] == uc.slot; ++j) if (st.ins_count[j] == lv.first) slot = st.n_base + j;
        if (slot == ~0ull) continue;
      }
      user_at[slot].push_back(lv.second);
    }
  }
  std::vector<uint8_t> overturned(flagged.size(), 0);
  auto evaluate_one = [&](size_t fi) {
    const uint32_t slot = flagged[fi];
    SlotEval& s = evals[fi];
    const uint8_t ref = fr ? fr->ref[order[fi]] : st.slot_ref[slot];
    evaluate_slot(s, ref, ep);
    const ColumnOut& co = flagged_cols[order[fi]];
    const bool dev_emit = (co.bits & CO_EMIT) != 0, dev_pred = (co.bits & CO_BASE_PREDICTED) != 0;
    if (dev_pred != s.base_predicted || (dev_emit && !s.emit)) overturned[fi] = 1;
    Reval rv;
    rv.base_predicted = s.base_predicted;
    rv.emit = s.emit;
    if (s.emit) {  // identify_mutations.cpp:1836-1910
      GdRow& row = rv.row;
      row.type = 0;
      uint64_t parent = slot < st.n_base ? slot : st.ins_parent[slot - st.n_base];
      row.b = slot < st.n_base ? 0 : st.ins_count[slot - st.n_base];
      size_t sg = 0;
      while (sg + 1 < st.segments.size() && st.segments[sg + 1].slot0 <= parent) ++sg;
      row.seq_id = hdr.target_names[(size_t)st.segments[sg].tid];
      row.a = (uint64_t)st.segments[sg].lo + (parent - st.segments[sg].slot0) + 1;
      row.ref_base = std::string(1, index_to_char(ref));
      row.new_base = std::string(1, index_to_char(s.variant));
      const uint32_t n = s.n();
      auto reported = [&](uint8_t b) { return (n == 0 || b >= 5) ? 0.0 : (s.f[b] < 0.5 / (double)n ? 0.0 : s.f[b]); };
      row.kv["score"] = format_double(s.variant_score, 1, false);
      row.kv["major_base"] = std::string(1, index_to_char(s.major));
      row.kv["minor_base"] = std::string(1, index_to_char(s.minor));
      row.kv["major_frequency"] = format_double(reported(s.major), ep.precision_places, true);
      row.kv["frequency"] = format_double(reported(s.variant), ep.precision_places, true);
      std::string spectrum;
      for (uint8_t b = 0; b < 5; ++b) {
        const double freq = reported(b);
        if (freq <= 0.0) continue;
        if (!spectrum.empty()) spectrum += ",";
        spectrum += std::string(1, index_to_char(b)) + ":" + format_double(freq, ep.precision_places, true);
      }
      row.kv["allele_frequencies"] = spectrum;
      // profile-likelihood bounds, identify_mutations.cpp:3175-3217
      double lower = 0.0, upper = 1.0;
      if (n > 0 && s.variant < 5) {
        const double drop = 0.587566, tol = ep.precision_decimal;
        const double f_hat = s.f[s.variant];
        const double target = profile_ll(s, s.variant, f_hat, tol) - drop;
        if (!(profile_ll(s, s.variant, 0.0, tol) >= target)) {
          double lo = 0.0, hi = f_hat;
          for (int i = 0; i < 40 && (hi - lo) > tol; ++i) {
            const double mid = 0.5 * (lo + hi);
            if (profile_ll(s, s.variant, mid, tol) >= target) hi = mid; else lo = mid;
          }
          lower = hi;
        }
        if (!(profile_ll(s, s.variant, 1.0, tol) >= target)) {
          double lo = f_hat, hi = 1.0;
          for (int i = 0; i < 40 && (hi - lo) > tol; ++i) {
            const double mid = 0.5 * (lo + hi);
            if (profile_ll(s, s.variant, mid, tol) >= target) lo = mid; else hi = mid;
          }
          upper = lo;
        }
      }
      row.kv["frequency_lower"] = format_double(lower, ep.precision_places, true);
      row.kv["frequency_upper"] = format_double(upper, ep.precision_places, true);
      // bias statistics, identify_mutations.cpp:3009-3033
      std::vector<uint32_t> major_q(128, 0), minor_q(128, 0);
      uint32_t n_major = 0, n_minor = 0;
      for (uint32_t i = 0; i < n; ++i) {
        if (s.obs[i] == s.major) { ++major_q[s.qual[i]]; ++n_major; }
        if (s.obs[i] == s.minor) { ++minor_q[s.qual[i]]; ++n_minor; }
      }
      double ks = 1.0;
      if (n_major && n_minor) ks = ks_less(minor_q, major_q);
      const double fisher = fisher_2x2(s.count[s.minor][1], s.count[s.minor][0], s.count[s.major][1], s.count[s.major][0]);
      row.kv["ks_quality_p_value"] = format_double(ks, 5, true);
      row.kv["fisher_strand_p_value"] = format_double(fisher, 5, true);
      auto cov = [&](uint8_t b) { return std::to_string(s.count[b][1]) + "/" + std::to_string(s.count[b][0]); };
      row.kv["ref_cov"] = cov(ref);
      row.kv["new_cov"] = cov(s.variant);
      row.kv["major_cov"] = cov(s.major);
      row.kv["minor_cov"] = cov(s.minor);
      uint32_t tot_top = 0, tot_bot = 0;
      for (int b = 0; b < 5; ++b) { tot_top += s.count[b][1]; tot_bot += s.count[b][0]; }
      row.kv["total_cov"] = std::to_string(tot_top) + "/" + std::to_string(tot_bot);
    }
    const auto ua = user_at.find(slot);
    if (ua != user_at.end()) {  // identify_mutations.cpp:1914-2019
      const uint32_t n = s.n();
      auto reported = [&](uint8_t b) { return (n == 0 || b >= 5) ? 0.0 : (s.f[b] < 0.5 / (double)n ? 0.0 : s.f[b]); };
      const uint64_t parent = slot < st.n_base ? slot : st.ins_parent[slot - st.n_base];
      const uint32_t level = slot < st.n_base ? 0u : st.ins_count[slot - st.n_base];
      (void)parent;
      for (uint32_t ei : ua->second) {
        const UserRa& e = st.user_list[ei];
        // the row the data already produced (same specification: cDiffEntry::compare): it only gains the mark
        if (s.emit && rv.row.seq_id == e.seq_id && rv.row.a == e.position && rv.row.b == e.insert_position && rv.row.ref_base == e.ref_base &&
            rv.row.new_base == e.new_base) { rv.row.kv["user_defined"] = "1"; continue; }
        GdRow u;
        u.type = 0; u.seq_id = e.seq_id; u.a = e.position; u.b = level; u.ref_base = e.ref_base; u.new_base = e.new_base;
        u.kv["user_defined"] = "1";
        const uint8_t uref = e.ref_base.size() == 1 ? char_to_index(e.ref_base[0]) : 255, uvar = e.new_base.size() == 1 ? char_to_index(e.new_base[0]) : 255;
        if (uref > 5 || uvar > 5) throw std::runtime_error("Unrecognized base char in user evidence: " + e.ref_base + " " + e.new_base);
        double score = std::numeric_limits<double>::quiet_NaN(), f_var = 0.0;
        if (uvar < 5) {
          if (n) { const Fit null_fit = fit_ordered(s, 0x1F & ~(1u << uvar), ep.precision_decimal); score = (s.log10_likelihood - null_fit.ll) - ep.log10_ref_length; }
          f_var = reported(uvar);
        }
        const double f_ref = reported(uref);
        const bool var_major = f_var > f_ref;
        const uint8_t mj = var_major ? uvar : uref, mn = var_major ? uref : uvar;
        u.kv["major_base"] = var_major ? e.new_base : e.ref_base;
        u.kv["minor_base"] = var_major ? e.ref_base : e.new_base;
        u.kv["major_frequency"] = format_double(var_major ? f_var : f_ref, ep.precision_places, true);
        u.kv["frequency"] = format_double(f_var, ep.precision_places, true);
        std::string spectrum;
        for (uint8_t b = 0; b < 5; ++b) {
          const double freq = reported(b);
          if (freq <= 0.0) continue;
          if (!spectrum.empty()) spectrum += ",";
          spectrum += std::string(1, index_to_char(b)) + ":" + format_double(freq, ep.precision_places, true);
        }
        u.kv["allele_frequencies"] = spectrum;
        double lower = 0.0, upper = 1.0;
        if (n > 0 && uvar < 5) {
          const double drop = 0.587566, tol = ep.precision_decimal, f_hat = s.f[uvar];
          const double target = profile_ll(s, uvar, f_hat, tol) - drop;
          if (!(profile_ll(s, uvar, 0.0, tol) >= target)) {
            double lo = 0.0, hi = f_hat;
            for (int i = 0; i < 40 && (hi - lo) > tol; ++i) { const double mid = 0.5 * (lo + hi); if (profile_ll(s, uvar, mid, tol) >= target) hi = mid; else lo = mid; }
            lower = hi;
          }
          if (!(profile_ll(s, uvar, 1.0, tol) >= target)) {
            double lo = f_hat, hi = 1.0;
            for (int i = 0; i < 40 && (hi - lo) > tol; ++i) { const double mid = 0.5 * (lo + hi); if (profile_ll(s, uvar, mid, tol) >= target) lo = mid; else hi = mid; }
            upper = lo;
          }
        }
        u.kv["frequency_lower"] = format_double(lower, ep.precision_places, true);
        u.kv["frequency_upper"] = format_double(upper, ep.precision_places, true);
        u.kv["prediction"] = ep.polymorphism_prediction ? "polymorphism" : (f_var > 0.5 ? "consensus" : "polymorphism");
        u.kv["score"] = format_double(score, 1, false);
        auto cov = [&](uint8_t b) { return std::to_string(s.count[b][1]) + "/" + std::to_string(s.count[b][0]); };
        u.kv["ref_cov"] = cov(uref); u.kv["new_cov"] = cov(uvar); u.kv["major_cov"] = cov(mj); u.kv["minor_cov"] = cov(mn);
        uint32_t tot_top = 0, tot_bot = 0;
        for (int b = 0; b < 5; ++b) { tot_top += s.count[b][1]; tot_bot += s.count[b][0]; }
        u.kv["total_cov"] = std::to_string(tot_top) + "/" + std::to_string(tot_bot);
        rv.user_rows.push_back(std::move(u));
      }
    }
    reval[fi] = std::move(rv);
  };
  {
    const size_t n_threads = std::min<size_t>(std::min<size_t>(16, std::max(1u, std::thread::hardware_concurrency())), (flagged.size() + 3) / 4);
    std::atomic<size_t> next{0};
    std::exception_ptr failure;
    std::mutex failure_lock;
    auto worker = [&] {
      try { for (size_t fi; (fi = next.fetch_add(1)) < flagged.size();) evaluate_one(fi); }
      catch (...) { std::lock_guard<std::mutex> g(failure_lock); if (!failure) failure = std::current_exception(); }
    };
    std::vector<std::thread> th;
    for (size_t t = 1; t < n_threads; ++t) th.emplace_back(worker);
    worker();
    for (auto& x : th) x.join();
    if (failure) std::rethrow_exception(failure);
  }
  counts.rechecked = flagged.size();
  for (uint8_t o : overturned) counts.overturned += o;


  EvidenceShard sh;
  sh.target_names = hdr.target_names; sh.target_lens = hdr.target_lens;
  sh.rechecked = counts.rechecked; sh.overturned = counts.overturned;
  for (const Segment& sg : st.segments) sh.segments.push_back({sg.tid, sg.lo, sg.hi});
  size_t ins_cursor = 0;  // ins slots are ordered by (parent, insert_count)
  // base slots are visited in ascending order and so are the insert sub-column slots: two cursors over the flagged list
  size_t ev_cur = 0;
  size_t base_cur = 0, ins_cur = std::lower_bound(flagged.begin(), flagged.end(), (uint32_t)std::min<uint64_t>(st.n_base, 0xFFFFFFFFull)) - flagged.begin();
  auto find_reval = [&](size_t& cur, uint64_t slot) -> const Reval* {
    while (cur < flagged.size() && flagged[cur] < slot) ++cur;
    return (cur < flagged.size() && flagged[cur] == slot) ? &reval[cur] : nullptr;
  };
  for (const Segment& sg : st.segments) {
    while (ev_cur < events.size() && events[ev_cur].slot < sg.slot0) ++ev_cur;
    for (; ev_cur < events.size() && events[ev_cur].slot < sg.slot0 + (uint64_t)(sg.hi - sg.lo); ++ev_cur) {
      const uint64_t slot = events[ev_cur].slot;
      EvidenceEvent e;
      e.tid = (uint32_t)sg.tid; e.pos1 = (uint32_t)(sg.lo + (int32_t)(slot - sg.slot0)) + 1;
      e.unique = events[ev_cur].w.unique; e.packed = events[ev_cur].w.packed;
      const Reval* rv = find_reval(base_cur, slot);
      if (rv) e.packed = (e.packed & ~1u) | (rv->base_predicted ? 1u : 0u);  // the host's verdict replaces the kernel's
      if (rv && rv->emit) e.rows.push_back(rv->row);
      if (rv) for (const GdRow& u : rv->user_rows) e.rows.push_back(u);
      while (ins_cursor < st.n_ins && st.ins_parent[ins_cursor] < slot) ++ins_cursor;
      for (; ins_cursor < st.n_ins && st.ins_parent[ins_cursor] == slot; ++ins_cursor) {
        const Reval* iv = find_reval(ins_cur, st.n_base + ins_cursor);
        if (iv && iv->emit) e.rows.push_back(iv->row);
        if (iv) for (const GdRow& u : iv->user_rows) e.rows.push_back(u);
      }
      sh.events.push_back(std::move(e));
    }
  }
  return sh;
}

// ---- the interval walk over the event columns of all shards of a run, and the GenomeDiff file
EvidenceCounts walk_evidence(const std::vector<const EvidenceShard*>& shards, const EvidenceParams& ep, const std::string& gd_path) {
  if (shards.empty()) throw std::runtime_error("no evidence shards");
  EvidenceCounts counts;
  const double nan = std::numeric_limits<double>::quiet_NaN();
  const std::vector<std::string>& names = shards[0]->target_names;
  const std::vector<uint32_t>& lens = shards[0]->target_lens;
  for (const EvidenceShard* sp : shards) {
    if (sp->target_names != names || sp->target_lens != lens) throw std::runtime_error("evidence shards of different references");
    counts.rechecked += sp->rechecked; counts.overturned += sp->overturned;
  }
  if (ep.deletion_propagation_cutoff.size() != names.size() || ep.deletion_seed_cutoff.size() != names.size())
    throw std::runtime_error("Number of targets in BAM file [" + std::to_string(names.size()) + "] does not match number in cutoff table [" +
                             std::to_string(ep.deletion_propagation_cutoff.size()) + "].");
  // per target: its pieces (shard, segment) in coordinate order; targets in visit order (alphabetical, pileup_base.cpp:364-385)
  struct Piece { int32_t lo, hi; const EvidenceShard* sh; };
  std::vector<std::vector<Piece>> pieces(names.size());
  for (const EvidenceShard* sp : shards) for (const auto& sg : sp->segments) if (sg.hi > sg.lo) pieces[(size_t)sg.tid].push_back({sg.lo, sg.hi, sp});
  std::vector<size_t> visit;
  for (size_t t = 0; t < names.size(); ++t) if (!pieces[t].empty()) visit.push_back(t);
  std::sort(visit.begin(), visit.end(), [&](size_t a, size_t b) { return names[a] < names[b]; });
  // cursor into every shard's event list (a shard's events are in visit order too)
  std::map<const EvidenceShard*, size_t> cursor;
  for (const EvidenceShard* sp : shards) cursor[sp] = 0;

  std::vector<GdRow> rows;
  uint64_t next_id = 0;
  auto add = [&](GdRow r) { r.id = ++next_id; rows.push_back(std::move(r)); };
  const uint32_t UNDEF = 0xFFFFFFFFu;
  struct Cov { double unique, redundant; int total; };
  for (size_t tid : visit) {
    std::vector<Piece>& pc = pieces[tid];
    std::sort(pc.begin(), pc.end(), [](const Piece& a, const Piece& b) { return a.lo < b.lo; });
    for (size_t i = 1; i < pc.size(); ++i) if (pc[i].lo != pc[i - 1].hi) throw std::runtime_error("evidence shards do not tile target " + names[tid]);
    const std::string& name = names[tid];
    const uint32_t tlen = lens[tid];
    const double prop = ep.deletion_propagation_cutoff[tid], seed = ep.deletion_seed_cutoff[tid];
    uint32_t del_start = UNDEF, del_end = UNDEF, red_start = UNDEF, red_end = UNDEF, unknown_start = UNDEF;
    bool reaches_seed = false, red_zero = false;
    Cov last = {nan, nan, 0}, left_out = {nan, nan, 0}, left_in = {nan, nan, 0};
    auto deletion_step = [&](uint32_t position, const Cov& cv) {  // identify_mutations.cpp:2262-2344
      if (position == 1) last = {nan, nan, 0};
      if (cv.unique <= prop && del_start == UNDEF) { del_start = position; left_out = last; left_in = cv; }
      if (!std::isnan(cv.unique) && cv.total <= seed) reaches_seed = true;
      if (del_start != UNDEF && (std::isnan(cv.unique) || cv.unique > prop)) {
        if (reaches_seed) {
          del_end = position - 1;
          if (red_end == UNDEF) red_end = del_end;
          if (red_start == UNDEF) red_start = del_start;
          GdRow r;
          r.type = 1; r.seq_id = name; r.a = del_start; r.b = del_end; r.c = red_start - del_start; r.d = del_end - red_end;
          r.kv["left_outside_cov"] = format_double(left_out.unique, 0, false);
          r.kv["left_inside_cov"] = format_double(left_in.unique, 0, false);
          r.kv["right_inside_cov"] = format_double(last.unique, 0, false);
          r.kv["right_outside_cov"] = format_double(cv.unique, 0, false);
          add(r);
          ++counts.mc;
        }
        reaches_seed = false; red_zero = false;
        del_start = del_end = red_start = red_end = UNDEF;
      }
      if (del_start != UNDEF) {
        if (cv.redundant == 0) { red_zero = true; red_end = UNDEF; }
        else if (cv.redundant > 0) { if (!red_zero) red_start = position; else if (red_end == UNDEF) red_end = position; }
      }
      last = cv;
    };
    auto unknown_step = [&](uint32_t position, bool predicted) {  // identify_mutations.cpp:2972-3007
      if (!predicted) { if (unknown_start == UNDEF) unknown_start = position; }
      else if (unknown_start != UNDEF) {
        GdRow r;
        r.type = 2; r.seq_id = name; r.a = unknown_start; r.b = position - 1;
        add(r);
        ++counts.un;
        unknown_start = UNDEF;
      }
    };
    // Only the event columns are walked: between two of them every column has unique coverage above the propagation
    // cutoff and a predicted base, no interval is open, and all such a column does to the state is overwrite `last`,
    // which the column before the next event overwrites again (it is an event itself: the neighbour of a non-boring one).
    for (const Piece& piece : pc) {
      size_t& cur = cursor[piece.sh];
      const std::vector<EvidenceEvent>& ev = piece.sh->events;
      for (; cur < ev.size() && ev[cur].tid == tid && (int32_t)ev[cur].pos1 <= piece.hi; ++cur) {
        if (prop < 0.0) continue;
        const EvidenceEvent& e = ev[cur];
        Cov cv;
        cv.unique = (double)e.unique;
        cv.redundant = (e.packed & 2u) ? 1.0 : 0.0;  // only its sign is looked at
        cv.total = (int)(e.packed >> 2);
        if (!ep.skip_missing_coverage_prediction) deletion_step(e.pos1, cv);
        unknown_step(e.pos1, (e.packed & 1u) != 0);
        for (const GdRow& r : e.rows) { add(r); ++counts.ra; }
      }
    }
    if ((uint32_t)pc.back().hi == tlen) {  // at_target_end, identify_mutations.cpp:2117-2164
      if (prop >= 0.0) {
        if (!ep.skip_missing_coverage_prediction) deletion_step(tlen + 1, Cov{nan, nan, 0});
        unknown_step(tlen + 1, true);
      } else if (!ep.skip_missing_coverage_prediction) {
        GdRow r;
        r.type = 1; r.seq_id = name; r.a = 1; r.b = tlen; r.c = 0; r.d = 0;
        r.kv["left_outside_cov"] = "NA";
        r.kv["left_inside_cov"] = format_double(0.0, 0, false);
        r.kv["right_inside_cov"] = format_double(0.0, 0, false);
        r.kv["right_outside_cov"] = "NA";
        add(r);
        ++counts.mc;
      }
    }
  }

  // ---- GenomeDiff text (genome_diff.cpp:685-760; sort keys genome_diff_entry.cpp:280-324, 566-700)
  std::stable_sort(rows.begin(), rows.end(), [](const GdRow& x, const GdRow& y) {
    if (x.type != y.type) return x.type < y.type;  // RA (3) < MC (4) < UN (7)
    if (x.seq_id != y.seq_id) return x.seq_id < y.seq_id;
    if (x.a != y.a) return x.a < y.a;
    if (x.b != y.b) return x.b < y.b;
    if (x.type == 0) { if (x.ref_base != y.ref_base) return x.ref_base < y.ref_base; if (x.new_base != y.new_base) return x.new_base < y.new_base; }
    else { if (x.c != y.c) return x.c < y.c; if (x.d != y.d) return x.d < y.d; }
    return x.id < y.id;
  });
  std::ofstream os(gd_path.c_str());
  if (!os) throw std::runtime_error("cannot create " + gd_path);
  os << "#=GENOME_DIFF\t1.0\n";
  static const char* type_name[3] = {"RA", "MC", "UN"};
  for (const GdRow& r : rows) {
    os << type_name[r.type] << '\t' << r.id << "\t.\t" << r.seq_id;
    if (r.type == 0) os << '\t' << r.a << '\t' << r.b << '\t' << r.ref_base << '\t' << r.new_base;
    else if (r.type == 1) os << '\t' << r.a << '\t' << r.b << '\t' << r.c << '\t' << r.d;
    else os << '\t' << r.a << '\t' << r.b;
    for (const auto& kv : r.kv) if (!kv.second.empty()) os << '\t' << kv.first << '=' << kv.second;
    os << '\n';
  }
  return counts;
}

EvidenceCounts write_evidence(const std::string& gd_path, const BamHeader& hdr, const PileupStream& st,
                              const std::vector<WalkEvent>& events, const std::vector<uint32_t>& flagged, const std::vector<ColumnOut>& flagged_cols,
                              const ScoreParams& sp, ClassLut& lut, const EvidenceParams& ep, const FlaggedRecords* fr) {
  const EvidenceShard sh = collect_evidence(hdr, st, events, flagged, flagged_cols, sp, lut, ep, fr);
  return walk_evidence({&sh}, ep, gd_path);
}

}  // namespace brq
