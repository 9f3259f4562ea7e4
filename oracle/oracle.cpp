// TEST INFRASTRUCTURE -- see oracle.h.  Restates, in reference order and reference arithmetic:
//   error_count.cpp:50-68, 125-199, 239-275, 477-690, 697-785, 854-1105
//   identify_mutations.cpp:48-88, 1309-1391, 1557-1910, 2028-2164, 2262-2344, 2972-3433
//   pileup_base.cpp:141-210, 308-385 ; alignment.cpp:44-57, 104-288 ; alignment.h:354-410, 533-591
//   common.h:202-316, 845-867 ; genome_diff.cpp:685-760 ; genome_diff_entry.cpp:566-700, 1323-1369
//   stats.cpp:534-650 (lngamma), 2074-2078, 2144-2171 (Fisher), 2191-2288 (KS)
// The DP/MP/PD/SC detectors co-resident in identify_mutations.cpp:1399-1555 are out of scope.
#include "oracle.h"

#include "htslib/faidx.h"
#include "htslib/sam.h"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <deque>
#include <list>
#include <sstream>

using namespace std;

namespace oracle {

#define ORACLE_ASSERT(c, m) do { if (!(c)) { cerr << "oracle ASSERT: " << (m) << endl; exit(1); } } while (0)

// ---------------------------------------------------------------- bases (common.h:202-316)
static const char base_char_list[] = {'A', 'C', 'G', 'T', '.'};
static const uint8_t base_list_size = 5;
static const uint8_t base_list_N_index = 5;

static inline bool bam_is_N(uint8_t b) { return b == 0xf; }          // _base_bam_is_N
static inline bool char_is_N(char c) { return c == 'N'; }            // _base_char_is_N
static uint8_t complement_base_bam(uint8_t b) {
  switch (b) { case 1: return 8; case 2: return 4; case 4: return 2; case 8: return 1; case 0xf: return 0xf; case '.': return '.'; }
  ORACLE_ASSERT(false, "Unrecognized BAM base"); return ' ';
}
static char complement_base_char(char c) {
  switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; case '.': return '.'; case 'N': return 'N'; }
  ORACLE_ASSERT(false, "Unrecognized base char"); return ' ';
}
static uint8_t complement_base_index(uint8_t b) { switch (b) { case 0: return 3; case 1: return 2; case 2: return 1; case 3: return 0; default: return 4; } }
static char basebam2char(uint8_t b) { switch (b) { case 1: return 'A'; case 2: return 'C'; case 4: return 'G'; case 8: return 'T'; case '.': return '.'; default: return 'N'; } }
static char baseindex2char(uint8_t b) { ORACLE_ASSERT(b < 6, "Unrecognized base index"); return "ACGT.N"[b]; }
static uint8_t basebam2index(uint8_t b) {
  switch (b) { case 1: return 0; case 2: return 1; case 4: return 2; case 8: return 3; case '.': return 4; }
  ORACLE_ASSERT(false, "BAM base not allowed"); return 0;
}
static uint8_t basechar2index(char c) {
  switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; case '.': return 4; case 'N': return 5; }
  ORACLE_ASSERT(false, string("Unrecognized base char: ") + c); return ' ';
}

// common.h:845-867 (the "-0." replacement there edits a temporary and has no effect)
static string to_string_double(double t, uint32_t precision = 1, bool use_scientific = false) {
  if (std::isnan(t)) return "NA";
  ostringstream interpreter;
  interpreter << (use_scientific ? scientific : fixed) << setprecision((int)precision) << t;
  string s = interpreter.str();
  if (use_scientific && s.size() >= 3 && s[s.size() - 3] == '0') s.erase(s.size() - 3, 1);
  return s;
}

// ---------------------------------------------------------------- alignment accessors
struct Aln {  // one pileup entry (alignment.h:43-416)
  const bam_pileup1_t* p;
  const bam1_t* a;
  explicit Aln(const bam_pileup1_t* pp) : p(pp), a(pp->b) {}
  bool is_del() const { return p->is_del; }
  int indel() const { return p->indel; }
  int on_base_indel() const { int i = indel(); if (i < 0) i = 0; if (is_del()) i = -1; return i; }
  uint32_t qpos0() const { return (uint32_t)p->qpos; }
  uint32_t qpos1() const { return (uint32_t)p->qpos + 1; }
  bool reversed() const { return bam_is_rev(a); }
  int strand() const { return reversed() ? -1 : +1; }
  uint32_t read_length() const { return (uint32_t)a->core.l_qseq; }
  uint8_t base_bam_0(uint32_t pos) const { ORACLE_ASSERT(pos < read_length(), "read_base_bam_0 out of range"); return bam_seqi(bam_get_seq(a), pos); }
  uint8_t on_base_bam(int32_t insert_count) const {
    uint32_t pos0 = qpos0() + (uint32_t)insert_count;
    if (on_base_indel() < insert_count) return '.';
    return base_bam_0(pos0);
  }
  uint8_t qual_0(uint32_t pos) const { ORACLE_ASSERT(pos < read_length(), "read_base_quality_0 out of range"); return bam_get_qual(a)[pos]; }
  uint32_t redundancy() const { uint8_t* x = bam_aux_get(a, "X1"); return x ? (uint32_t)bam_aux2i(x) : 1; }
  const char* read_group_id() const { uint8_t* x = bam_aux_get(a, "RG"); return x ? bam_aux2Z(x) : NULL; }
  uint32_t query_start_1() const {  // alignment.cpp:248-266
    uint32_t* cigar = bam_get_cigar(a);
    int32_t pos = 1;
    for (uint32_t j = 0; j < a->core.n_cigar; j++) {
      if ((cigar[j] & BAM_CIGAR_MASK) != BAM_CSOFT_CLIP) break;
      pos += (int32_t)(cigar[j] >> BAM_CIGAR_SHIFT);
    }
    return (uint32_t)pos;
  }
  uint32_t query_end_1() const {    // alignment.cpp:269-288
    uint32_t* cigar = bam_get_cigar(a);
    int32_t pos = (int32_t)bam_cigar2qlen((int)a->core.n_cigar, cigar);
    for (uint32_t j = a->core.n_cigar - 1; j > 0; j--) {
      if ((cigar[j] & BAM_CIGAR_MASK) != BAM_CSOFT_CLIP) break;
      pos -= (int32_t)(cigar[j] >> BAM_CIGAR_SHIFT);
    }
    return (uint32_t)pos;
  }
  void query_bounds_0(uint32_t& start, uint32_t& end) const {  // alignment.cpp:104-218 with min_qual == 0
    uint32_t* cigar = bam_get_cigar(a);
    uint32_t s = 1, e = (uint32_t)bam_cigar2qlen((int)a->core.n_cigar, cigar);
    for (uint32_t i = 0; i < a->core.n_cigar; i++) {
      uint32_t op = cigar[i] & BAM_CIGAR_MASK, len = cigar[i] >> BAM_CIGAR_SHIFT;
      if (op != BAM_CSOFT_CLIP && op != BAM_CHARD_CLIP && op != BAM_CREF_SKIP) break;
      if (op == BAM_CSOFT_CLIP) s += len;
    }
    for (uint32_t i = a->core.n_cigar - 1; i > 0; --i) {
      uint32_t op = cigar[i] & BAM_CIGAR_MASK, len = cigar[i] >> BAM_CIGAR_SHIFT;
      if (op != BAM_CSOFT_CLIP && op != BAM_CHARD_CLIP && op != BAM_CREF_SKIP) break;
      if (op == BAM_CSOFT_CLIP) e -= len;
    }
    start = s - 1; end = e - 1;
  }
  bool is_trimmed(bool past_base) const {  // alignment.h:389-410
    uint8_t* auxl = bam_aux_get(a, "XL");
    if (auxl) { if (qpos1() <= (uint32_t)bam_aux2i(auxl)) return true; }
    uint8_t* auxr = bam_aux_get(a, "XR");
    if (auxr) {
      if (read_length() - qpos1() + 1 <= (uint32_t)bam_aux2i(auxr)) return true;
      if (past_base && (read_length() - qpos1() == (uint32_t)bam_aux2i(auxr))) return true;
    }
    return false;
  }
  uint32_t base_repeat_0(uint32_t q) const {  // alignment.cpp:371-390
    uint8_t b = base_bam_0(q);
    uint32_t rep = 0;
    if (!reversed()) { while (q < query_end_1() - 1) { q++; if (b != base_bam_0(q)) break; rep++; } }
    else { while (q > 0) { q--; if (b != base_bam_0(q)) break; rep++; } }
    return rep;
  }
};

struct ReadGroupMap {  // alignment.h:533-565, alignment.cpp:545-563
  vector<string> ids, libraries;
  void build(bam_hdr_t* hdr) {
    int n = sam_hdr_count_lines(hdr, "RG");
    if (n <= 0) return;
    kstring_t ks = KS_INITIALIZE;
    for (int i = 0; i < n; i++) {
      const char* id = sam_hdr_line_name(hdr, "RG", i);
      if (id == NULL) continue;
      ids.push_back(id);
      ks.l = 0;
      libraries.push_back((sam_hdr_find_tag_pos(hdr, "RG", i, "LB", &ks) == 0) ? string(ks.s) : string());
    }
    ks_free(&ks);
  }
  uint32_t index(const Aln& a) const {
    if (ids.size() <= 1) return 0;
    const char* rg = a.read_group_id();
    if (rg == NULL) return 0;
    for (size_t i = 0; i < ids.size(); i++) if (ids[i] == rg) return (uint32_t)i;
    return 0;
  }
};

struct ReadFilePartition {  // alignment.h:576-591, alignment.cpp:565-605
  vector<uint32_t> base, count;
  bool empty() const { return base.empty(); }
  uint32_t index(uint32_t group, bool is_read2) const {
    if (group >= base.size()) return 0;
    return base[group] + ((is_read2 && count[group] > 1) ? 1 : 0);
  }
  static ReadFilePartition make(const ReadGroupMap& rg, const vector<ReadFileSet>& sets) {
    ReadFilePartition p;
    if (sets.empty()) return p;
    vector<uint32_t> set_base, set_count;
    uint32_t flat = 0;
    for (size_t s = 0; s < sets.size(); s++) { set_base.push_back(flat); set_count.push_back(sets[s].n_files); flat += sets[s].n_files; }
    for (size_t g = 0; g < rg.ids.size(); g++) {
      size_t match = 0; bool found = false;
      const string& lb = rg.libraries[g];
      if (!lb.empty()) for (size_t s = 0; s < sets.size(); s++) if (sets[s].base_name == lb) { match = s; found = true; break; }
      if (!found) { match = g; found = (g < set_base.size()); }
      if (found && match < set_base.size()) { p.base.push_back(set_base[match]); p.count.push_back(set_count[match]); }
      else { p.base.push_back(0); p.count.push_back(1); }
    }
    return p;
  }
};

// ---------------------------------------------------------------- error table (error_count.cpp:358-1123)
enum { k_read_set, k_ref_base, k_prev_base, k_obs_base, k_quality, k_read_pos, k_base_repeat, k_num_covariates };
static const char* covariate_names[] = {"read_set", "ref_base", "prev_ref_base", "obs_base", "quality", "read_pos", "base_repeat"};
typedef uint32_t cv_t[k_num_covariates];

struct ErrorTable {
  bool used[k_num_covariates], enforce_max[k_num_covariates], per_position = false;
  uint32_t maxv[k_num_covariates], offset[k_num_covariates];
  vector<double> count, log10_prob, prob;
  const ReadGroupMap* rg_map = NULL;
  ReadFilePartition partition;

  ErrorTable() { for (int i = 0; i < k_num_covariates; i++) { used[i] = enforce_max[i] = false; maxv[i] = offset[i] = 0; } }

  void read_covariates(const string& colnames) {  // :522-594
    for (int i = 0; i < k_num_covariates; i++) { used[i] = false; enforce_max[i] = false; }
    stringstream ss(colnames);
    string item;
    while (getline(ss, item, ',')) {
      string key = item, val;
      size_t eq = item.find('=');
      if (eq != string::npos) { key = item.substr(0, eq); val = item.substr(eq + 1); }
      if (key == "ref_base") { used[k_ref_base] = true; maxv[k_ref_base] = 5; }
      else if (key == "prev_base") { used[k_prev_base] = true; maxv[k_prev_base] = 5; }
      else if (key == "obs_base") { used[k_obs_base] = true; maxv[k_obs_base] = 5; }
      else if (key == "quality") { used[k_quality] = true; maxv[k_quality] = (uint32_t)atoi(val.c_str()); }
      else if (key == "read_set") { used[k_read_set] = true; maxv[k_read_set] = (uint32_t)atoi(val.c_str()); }
      else if (key == "ref_pos") { per_position = true; }
      else if (key == "read_pos") { used[k_read_pos] = true; maxv[k_read_pos] = (uint32_t)atoi(val.c_str()); }
      else if (key == "base_repeat") { used[k_base_repeat] = true; maxv[k_base_repeat] = (uint32_t)atoi(val.c_str()); enforce_max[k_base_repeat] = true; }
      else cerr << "Unrecognized covariate: " << key << endl;
    }
    uint32_t cur = 1;
    for (int i = 0; i < k_num_covariates; i++) if (used[i]) { offset[i] = cur; cur *= maxv[i]; }
  }
  void allocate() { int n = 1; for (int i = 0; i < k_num_covariates; i++) if (used[i]) n *= (int)maxv[i]; count.assign((size_t)n, 0.0); }

  uint32_t to_index(const cv_t cv) const {  // :477-497
    uint32_t idx = 0;
    for (int i = 0; i < k_num_covariates; i++) {
      if (!used[i]) continue;
      uint32_t val = cv[i];
      if (val >= maxv[i]) {
        ORACLE_ASSERT(enforce_max[i], string("Covariate '") + covariate_names[i] + "' exceeded enforced maximum value.");
        val = maxv[i] - 1;
      }
      idx += val * offset[i];
    }
    return idx;
  }
  string print_covariates() const {  // :601-621
    string s;
    if (per_position) s += "ref_pos";
    for (int i = 0; i < k_num_covariates; i++) {
      if (!used[i]) continue;
      if (!s.empty()) s += ",";
      s += covariate_names[i];
      if (i != k_ref_base && i != k_obs_base) s += "=" + std::to_string(maxv[i]);
    }
    return s;
  }
  uint32_t read_file_index(const Aln& a) const {  // error_count.h:210-214
    if (rg_map == NULL || partition.empty()) return 0;
    return partition.index(rg_map->index(a), (a.a->core.flag & BAM_FREAD2) != 0);
  }

  void count_alignment_position(const Aln& i, uint32_t ref_pos0, const char* ref_seq) {  // :854-986
    uint32_t reversed = i.reversed() ? 1 : 0;
    int32_t q_pos_0 = (int32_t)i.qpos0();
    int32_t q_start_0 = (int32_t)i.query_start_1() - 1, q_end_0 = (int32_t)i.query_end_1() - 1;
    int32_t q_length = (int32_t)i.read_length();
    const uint8_t* qseq = bam_get_seq(i.a);
    const uint8_t* qscore = bam_get_qual(i.a);
    cv_t cv;
    memset(cv, 0, sizeof cv);
    cv[k_read_set] = read_file_index(i);
    cv[k_read_pos] = (uint32_t)q_pos_0;
    {
      uint8_t obs = bam_seqi(qseq, q_pos_0);
      char ref = ref_seq[ref_pos0];
      if (!bam_is_N(obs) && !char_is_N(ref)) {
        if (reversed) { obs = complement_base_bam(obs); ref = complement_base_char(ref); }
        cv[k_quality] = qscore[q_pos_0];
        cv[k_obs_base] = basebam2index(obs);
        cv[k_ref_base] = basechar2index(ref);
        if (used[k_base_repeat]) cv[k_base_repeat] = i.base_repeat_0((uint32_t)q_pos_0);
        count[to_index(cv)]++;
      }
    }
    if (i.indel() == 0) {
      if (q_pos_0 < q_end_0) {
        int32_t mqpos = q_pos_0 + 1 - (int32_t)reversed;
        ORACLE_ASSERT(mqpos >= 0 && mqpos < q_length, "nonexistent base for '..' state");
        uint8_t obs = bam_seqi(qseq, mqpos);
        char ref = ref_seq[ref_pos0 + 1 - reversed];
        if (!bam_is_N(obs) && !char_is_N(ref)) {
          cv[k_quality] = qscore[mqpos];
          cv[k_obs_base] = 4; cv[k_ref_base] = 4;
          if (used[k_base_repeat]) cv[k_base_repeat] = i.base_repeat_0((uint32_t)mqpos);
          count[to_index(cv)]++;
        }
      }
    } else if (i.indel() == -1) {
      int32_t mqpos = q_pos_0 + 1 - (int32_t)reversed;
      ORACLE_ASSERT(mqpos >= 0 && mqpos < q_length, "nonexistent base for 'N.' state");
      uint8_t obs = bam_seqi(qseq, mqpos);
      char ref = ref_seq[ref_pos0 + 1];
      if (!bam_is_N(obs) && !char_is_N(ref)) {
        if (reversed) ref = complement_base_char(ref);
        cv[k_quality] = qscore[mqpos];
        cv[k_obs_base] = 4; cv[k_ref_base] = basechar2index(ref);
        if (used[k_base_repeat]) cv[k_base_repeat] = i.base_repeat_0((uint32_t)mqpos);
        count[to_index(cv)]++;
      }
    } else if (i.indel() == +1) {
      int32_t mqpos = q_pos_0 + 1;
      ORACLE_ASSERT(mqpos >= 0 && mqpos < q_length, "nonexistent base for '.N' state");
      if (mqpos <= q_end_0 && mqpos >= q_start_0) {
        uint8_t obs = bam_seqi(qseq, mqpos);
        if (!bam_is_N(obs)) {
          if (reversed) obs = complement_base_bam(obs);
          cv[k_quality] = qscore[mqpos];
          cv[k_obs_base] = basebam2index(obs); cv[k_ref_base] = 4;
          if (used[k_base_repeat]) cv[k_base_repeat] = i.base_repeat_0((uint32_t)mqpos);
          count[to_index(cv)]++;
        }
      }
    }
  }

  void counts_to_log10_prob() {  // :1005-1026 with the marginalising constructor :424-458
    log10_prob.assign(count.size(), 0.0);
    uint32_t s_offset[k_num_covariates], s_max[k_num_covariates];
    bool s_used[k_num_covariates];
    uint32_t cur = 1;
    for (int i = 0; i < k_num_covariates; i++) {
      s_used[i] = used[i] && i != k_obs_base;
      s_max[i] = s_used[i] ? maxv[i] : 0;
      s_offset[i] = s_used[i] ? cur : 0;
      if (s_used[i]) cur *= s_max[i];
    }
    vector<double> sums(cur, 0.0);
    auto sum_index = [&](uint32_t idx) {
      uint32_t j = 0;
      for (int i = 0; i < k_num_covariates; i++) if (s_used[i]) j += ((idx / offset[i]) % maxv[i]) * s_offset[i];
      return j;
    };
    for (uint32_t i = 0; i < count.size(); i++) sums[sum_index(i)] += count[i];
    const uint32_t smoothing_factor = 1;
    for (uint32_t i = 0; i < count.size(); i++)
      log10_prob[i] = log10((double)(count[i] + smoothing_factor)) - log10((double)sums[sum_index(i)] + smoothing_factor * maxv[k_obs_base]);
  }

  void write_rows(ostream& out, const vector<double>& v) const {
    for (uint32_t idx = 0; idx < v.size(); idx++) {
      for (int i = 0; i < k_num_covariates; i++) {
        if (!used[i]) continue;
        uint32_t j = (idx / offset[i]) % maxv[i];
        if (i == k_ref_base || i == k_obs_base) out << baseindex2char((uint8_t)j) << '\t';
        else out << j << '\t';
      }
      out << v[idx] << endl;
    }
  }
  void write_log10_prob_table(const string& fn) const {  // :660-690
    ofstream out(fn.c_str());
    out << print_covariates() << endl;
    for (int i = 0; i < k_num_covariates; i++) if (used[i]) out << covariate_names[i] << '\t';
    out << "log10_probability" << endl;
    write_rows(out, log10_prob);
  }
  void write_count_table(const string& fn) const {  // :791-846
    ofstream out(fn.c_str());
    out << print_covariates() << endl;
    for (int i = 0; i < k_num_covariates; i++) if (used[i]) out << covariate_names[i] << '\t';
    out << "count" << endl;
    out << setprecision(17);
    write_rows(out, count);
  }
  void read_log10_prob_table(const string& fn) {  // :629-654
    ifstream in(fn.c_str());
    ORACLE_ASSERT(in.good(), "cannot open error table " + fn);
    string s;
    getline(in, s);
    read_covariates(s);
    allocate();
    log10_prob.assign(count.size(), 0.0);
    getline(in, s);
    for (uint32_t i = 0; i < log10_prob.size(); i++) {
      getline(in, s);
      size_t t = s.rfind('\t');
      log10_prob[i] = strtod(s.c_str() + (t == string::npos ? 0 : t + 1), NULL);
    }
  }
  void log10_prob_to_prob() {  // :1032-1040
    prob.assign(log10_prob.size(), 0.0);
    for (uint32_t i = 0; i < log10_prob.size(); i++) prob[i] = pow(10, log10_prob[i]);
  }

  void write_base_qual_only_prob_table(const string& fn_pattern, const vector<string>& readfiles) const {  // :697-785
    const uint32_t nS = maxv[k_read_set], nO = maxv[k_obs_base], nR = maxv[k_ref_base], nQ = maxv[k_quality];
    vector<double> t((size_t)nS * nO * nR * nQ, 0.0);
    for (uint32_t idx = 0; idx < count.size(); idx++) {
      uint32_t add = 0;
      for (int i = 0; i < k_num_covariates; i++) {
        if (!used[i]) continue;
        uint32_t j = (idx / offset[i]) % maxv[i];
        if (i == k_read_set) add += nO * nR * nQ * j;
        else if (i == k_ref_base) add += j;
        else if (i == k_obs_base) add += nR * j;
        else if (i == k_quality) add += nO * nR * j;
      }
      t[add] += count[idx];
    }
    double running_total = 0;
    for (uint32_t r = 0; r < nS; r++) {
      for (uint32_t k = 0; k < nO * nR * nQ; k++) {
        uint32_t i = r * nO * nR * nQ + k;
        running_total += t[i];
        if (i % nR == nR - 1) {
          for (uint32_t j = i - (nR - 1); j <= i; j++) {
            if (running_total > 0) t[j] /= running_total;
            if (t[j] == 0) t[j] = NAN;
          }
          running_total = 0;
        }
      }
      string fn = fn_pattern;
      size_t h = fn.find('#');
      if (h != string::npos) fn.replace(h, 1, readfiles[r]);
      ofstream out(fn.c_str());
      out << "quality";
      for (uint32_t b1 = 0; b1 < nR; b1++) for (uint32_t b2 = 0; b2 < nO; b2++) out << "\t" << baseindex2char((uint8_t)b1) << baseindex2char((uint8_t)b2);
      out << endl;
      for (uint32_t q = 0; q < nQ; q++) {
        out << q;
        for (uint32_t b1 = 0; b1 < nR; b1++) for (uint32_t b2 = 0; b2 < nO; b2++) {
          uint32_t idx = r * nO * nR * nQ + q * nO * nR + b1 * nR + b2;
          out << "\t";
          if (std::isnan(t[idx])) out << "NA"; else out << t[idx];
        }
        out << endl;
      }
    }
  }

  // :1049-1105 ; ref_base is NOT filled in
  bool alignment_position_to_covariates(const Aln& a, int32_t insert_count, cv_t cv) const {
    int indel = a.on_base_indel();
    uint8_t read_base_bam = a.on_base_bam(insert_count);
    if (bam_is_N(read_base_bam)) return false;
    uint32_t q_start_0, q_end_0;
    a.query_bounds_0(q_start_0, q_end_0);
    uint32_t q_pos_0 = a.qpos0();
    if (indel == -1) {
      q_pos_0 += 1 - (a.reversed() ? 1 : 0);
      if (bam_is_N(a.base_bam_0(q_pos_0))) return false;
    } else if (insert_count > 0) {
      int32_t max_offset = insert_count;
      if (indel < max_offset) max_offset = indel;
      q_pos_0 += (uint32_t)(max_offset + 1 - (a.reversed() ? 1 : 0));
      if (q_pos_0 > q_end_0) return false;
      if (bam_is_N(a.base_bam_0(q_pos_0))) return false;
    }
    cv[k_obs_base] = basebam2index(read_base_bam);
    cv[k_quality] = a.qual_0(q_pos_0);
    cv[k_read_set] = read_file_index(a);
    cv[k_read_pos] = q_pos_0;
    if (used[k_base_repeat]) cv[k_base_repeat] = a.base_repeat_0(q_pos_0);
    return true;
  }
  double get_prob(const cv_t cv) const { uint32_t i = to_index(cv); ORACLE_ASSERT(i < prob.size(), "prob index"); return prob[i]; }
};

// ---------------------------------------------------------------- pileup driver (pileup_base.cpp)
struct PileupDriver {
  htsFile* bam = NULL; hts_idx_t* idx = NULL; bam_hdr_t* hdr = NULL; faidx_t* fai = NULL;
  vector<char*> refs; vector<int> ref_lens;
  ReadGroupMap read_groups;
  uint32_t last_position_1 = 0;
  uint64_t n_records = 0;

  PileupDriver(const string& bam_fn, const string& fasta_fn) {  // :61-88
    bam = hts_open(bam_fn.c_str(), "rb");
    ORACLE_ASSERT(bam, "Could not load bam file: " + bam_fn);
    idx = sam_index_load(bam, bam_fn.c_str());
    hdr = sam_hdr_read(bam);
    read_groups.build(hdr);
    fai = fai_load(fasta_fn.c_str());
    ORACLE_ASSERT(fai, "Could not load fasta: " + fasta_fn);
    for (int i = 0; i < hdr->n_targets; ++i) {
      int len = 0;
      char* s = fai_fetch(fai, hdr->target_name[i], &len);
      ORACLE_ASSERT(s && len > 0, "missing reference sequence");
      ORACLE_ASSERT((uint32_t)len == hdr->target_len[i], "reference length mismatch");
      refs.push_back(s); ref_lens.push_back(len);
    }
  }
  virtual ~PileupDriver() {
    for (char* s : refs) free(s);
    fai_destroy(fai); sam_hdr_destroy(hdr); hts_idx_destroy(idx); hts_close(bam);
  }
  uint32_t num_targets() const { return (uint32_t)hdr->n_targets; }
  uint32_t target_length(uint32_t tid) const { return hdr->target_len[tid]; }
  const char* target_name(uint32_t tid) const { return hdr->target_name[tid]; }

  virtual void pileup_callback(uint32_t tid, uint32_t pos1, int n, const bam_pileup1_t* pile) = 0;
  virtual void at_target_start(uint32_t) {}
  virtual void at_target_end(uint32_t) {}

  struct IterData { samFile* fp; hts_itr_t* iter; };
  static int iter_read(void* data, bam1_t* b) { IterData* d = (IterData*)data; return sam_itr_next(d->fp, d->iter, b); }

  void do_pileup_target(uint32_t target_id) {  // :308-359 for region "<name>:1-<len>"
    uint32_t end_pos_1 = target_length(target_id);
    last_position_1 = 0;
    at_target_start(target_id);
    hts_itr_t* iter = sam_itr_queryi(idx, (int)target_id, 0, (hts_pos_t)end_pos_1);
    ORACLE_ASSERT(iter, "Could not create iterator");
    IterData ird = {bam, iter};
    bam_plp_t plp = bam_plp_init(iter_read, &ird);
    bam_plp_set_maxcnt(plp, 1000000000);
    int tid2, pos2, n;
    const bam_pileup1_t* pile;
    while ((pile = bam_plp_auto(plp, &tid2, &pos2, &n)) != NULL) {
      uint32_t this_pos_1 = (uint32_t)pos2 + 1;  // first_level_pileup_callback :141-210 (same tid by construction)
      for (uint32_t p = last_position_1 + 1; p <= this_pos_1 - 1; p++) { pileup_callback(target_id, p, 0, NULL); last_position_1 = p; }
      n_records += (uint64_t)n;
      pileup_callback(target_id, this_pos_1, n, pile);
      last_position_1 = this_pos_1;
    }
    bam_plp_destroy(plp);
    hts_itr_destroy(iter);
    for (uint32_t p = last_position_1 + 1; p <= end_pos_1; p++) { pileup_callback(target_id, p, 0, NULL); last_position_1 = p; }
    at_target_end(target_id);
  }
  void do_pileup(const set<string>& seq_ids) {  // :364-385
    for (set<string>::const_iterator it = seq_ids.begin(); it != seq_ids.end(); it++) {
      uint32_t tid = 0; bool found = false;
      while (!found && tid < num_targets()) { if (*it == target_name(tid)) found = true; else tid++; }
      ORACLE_ASSERT(found, "Could not find seq_id: " + *it);
      do_pileup_target(tid);
    }
  }
};

// ---------------------------------------------------------------- pass 1
struct ErrorCountPileup : PileupDriver {
  const Settings& settings;
  bool do_errors;
  ErrorTable table;
  vector<vector<uint32_t> > unique_only_coverage;
  uint32_t on_group = 0;

  bool preprocess_stage = false;
  uint64_t read_found_starting_at_pos[2] = {0, 0};
  std::map<string, double> no_pos_hash_per_position_pr;

  ErrorCountPileup(const Settings& s, const string& bam, const string& fasta, bool errors, const string& covariates)
      : PileupDriver(bam, fasta), settings(s), do_errors(errors) {
    table.read_covariates(covariates);
    table.allocate();
    unique_only_coverage.resize(s.n_coverage_groups);
    table.rg_map = &read_groups;
    table.partition = ReadFilePartition::make(read_groups, s.read_file_sets);
  }
  void at_target_start(uint32_t tid) {  // error_count.cpp:201-214
    on_group = settings.seq_id_to_coverage_group.find(target_name(tid))->second;
    if (preprocess_stage) { read_found_starting_at_pos[0] = 0; read_found_starting_at_pos[1] = 0; }
  }
  void at_target_end(uint32_t tid) {  // error_count.cpp:217-229
    if (!preprocess_stage) return;
    const double total = (double)(read_found_starting_at_pos[0] + read_found_starting_at_pos[1]);
    no_pos_hash_per_position_pr[target_name(tid)] = total != 0 ? (double)read_found_starting_at_pos[0] / total : 1.0;
  }
  uint32_t required_junction_read_end_min_coordinate(uint32_t read_length) const {  // settings.h:345-354
    const int32_t max_len = (int32_t)floor((double)((int32_t)read_length - (int32_t)settings.unmatched_end_minimum_read_length) *
                                           settings.unmatched_end_length_factor);
    return max_len <= 0 ? read_length : read_length - (uint32_t)max_len;
  }
  void pileup_callback(uint32_t tid, uint32_t pos1, int n, const bam_pileup1_t* pile) {  // error_count.cpp:125-199
    vector<uint32_t>& cov = unique_only_coverage[on_group];
    size_t unique_coverage = 0;
    bool has_redundant_reads = false;
    int has_query_start[2] = {0, 0};
    for (int k = 0; k < n; ++k) {
      Aln i(&pile[k]);
      if (i.is_del()) continue;
      if (i.redundancy() > 1) { has_redundant_reads = true; continue; }
      ++unique_coverage;
      if (preprocess_stage && i.qpos1() == 1) {  // :157-166; query_stranded_end_1: alignment.cpp:228-239
        uint32_t qs0, qe0;
        i.query_bounds_0(qs0, qe0);
        const uint32_t stranded_end_1 = i.reversed() ? i.read_length() - (qs0 + 1) + 1 : qe0 + 1;
        if (stranded_end_1 >= required_junction_read_end_min_coordinate(i.read_length())) has_query_start[i.reversed() ? 1 : 0] = 1;
      }
      if (!do_errors) continue;
      table.count_alignment_position(i, pos1 - 1, refs[tid]);
    }
    if (!has_redundant_reads) {
      if (unique_coverage >= cov.size()) cov.resize(unique_coverage + 1, 0);
      ++cov[unique_coverage];
      if (preprocess_stage) { read_found_starting_at_pos[has_query_start[0]]++; read_found_starting_at_pos[has_query_start[1]]++; }
    }
  }
};

void error_count(const Settings& settings, const string& bam, const string& fasta, const string& output_dir,
                 const vector<string>& readfiles, bool do_coverage, bool do_errors, const string& covariates,
                 const string& counts_dump_file, bool preprocess_stage, std::map<string, double>* no_pos_hash_per_position_pr) {
  ErrorCountPileup ecp(settings, bam, fasta, do_errors, covariates);
  ecp.preprocess_stage = preprocess_stage;
  ecp.do_pileup(settings.call_mutations_seq_ids);
  if (no_pos_hash_per_position_pr) *no_pos_hash_per_position_pr = ecp.no_pos_hash_per_position_pr;
  if (do_coverage) {  // print_coverage :239-253
    for (size_t i = 0; i < ecp.unique_only_coverage.size(); ++i) {
      string fn = settings.unique_only_coverage_distribution_file_name;
      size_t at = fn.find('@');
      if (at != string::npos) fn.replace(at, 1, std::to_string(i));
      ofstream out((output_dir + "/" + fn).c_str());
      out << "coverage\tn" << endl;
      for (size_t j = 1; j < ecp.unique_only_coverage[i].size(); ++j) out << j << "\t" << ecp.unique_only_coverage[i][j] << endl;
    }
  }
  if (do_errors) {  // print_error :258-275
    if (!counts_dump_file.empty()) ecp.table.write_count_table(counts_dump_file);
    ecp.table.counts_to_log10_prob();
    ecp.table.write_log10_prob_table(settings.error_rates_file_name);
    ecp.table.write_base_qual_only_prob_table(output_dir + "/" + settings.base_qual_error_prob_file_name, readfiles);
  }
}

// ---------------------------------------------------------------- stats (stats.cpp)
static double lngamma(double x) {  // Cephes lgam as restated at stats.cpp:534-650, positive arguments
  const double ls2pi = 0.91893853320467274178;
  ORACLE_ASSERT(x >= 0, "lngamma domain");
  if (x < 13) {
    double z = 1, p = 0, u = x;
    while (u >= 3) { p = p - 1; u = x + p; z = z * u; }
    while (u < 2) { z = z / u; p = p + 1; u = x + p; }
    if (z < 0) z = -z;
    if (u == 2) return log(z);
    p = p - 2;
    x = x + p;
    static const double B[] = {-1378.25152569120859100, -38801.6315134637840924, -331612.992738871184744,
                               -1162370.97492762307383, -1721737.00820839662146, -853555.664245765465627};
    static const double C[] = {1, -351.815701436523470549, -17064.2106651881159223, -220528.590553854454839,
                               -1139334.44367982507207, -2532523.07177582951285, -2018891.41433532773231};
    double b = B[0];
    for (int i = 1; i < 6; i++) b = B[i] + x * b;
    double c = C[0];
    for (int i = 1; i < 7; i++) c = C[i] + x * c;
    return log(z) + x * b / c;
  }
  double q = (x - 0.5) * log(x) - x + ls2pi;
  if (x > 100000000) return q;
  double p = 1 / (x * x);
  if (x >= 1000.0) {
    q = q + ((7.9365079365079365079365 * 0.0001 * p - 2.7777777777777777777778 * 0.001) * p + 0.0833333333333333333333) / x;
  } else {
    double a = 8.11614167470508450300 * 0.0001;
    a = -5.95061904284301438324 * 0.0001 + p * a;
    a = 7.93650340457716943945 * 0.0001 + p * a;
    a = -2.77777777730099687205 * 0.001 + p * a;
    a = 8.33333333333331927722 * 0.01 + p * a;
    q = q + a / x;
  }
  return q;
}
static double log_choose(double n, double k) { return lngamma(n + 1) - lngamma(k + 1) - lngamma(n - k + 1); }

static double fisher_exact_test_2x2(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {  // stats.cpp:2144-2171
  uint32_t row1 = a + b, row2 = c + d, col1 = a + c, col2 = b + d, n = row1 + row2;
  uint32_t a_min = (row1 > col2) ? (row1 - col2) : 0, a_max = min(row1, col1);
  double log_denom = log_choose(n, row1);
  double log_p_observed = log_choose(col1, a) + log_choose(col2, row1 - a) - log_denom;
  const double log_rel_err = log(1.0 + 1e-7);
  double total_pr = 0.0;
  for (uint32_t a_i = a_min; a_i <= a_max; a_i++) {
    double log_p_i = log_choose(col1, a_i) + log_choose(col2, row1 - a_i) - log_denom;
    if (log_p_i <= log_p_observed + log_rel_err) total_pr += exp(log_p_i);
  }
  return min(total_pr, 1.0);
}

static double ks_test_two_sample_less(const vector<double>& x, const vector<double>& y) {  // stats.cpp:2191-2288
  uint32_t n_x = (uint32_t)x.size(), n_y = (uint32_t)y.size();
  vector<pair<double, bool> > combined;
  for (size_t i = 0; i < x.size(); i++) combined.push_back(make_pair(x[i], true));
  for (size_t i = 0; i < y.size(); i++) combined.push_back(make_pair(y[i], false));
  sort(combined.begin(), combined.end());
  uint32_t count_x = 0, count_y = 0, best_x = 0, best_y = 0;
  double min_z = 0.0;
  bool have_min = false;
  vector<bool> boundary(x.size() + y.size() + 1, true);
  size_t i = 0;
  while (i < combined.size()) {
    double v = combined[i].first;
    size_t group_start = i;
    while (i < combined.size() && combined[i].first == v) { if (combined[i].second) count_x++; else count_y++; i++; }
    for (size_t k = group_start + 1; k < i; k++) boundary[k] = false;
    double z = (double)count_x / n_x - (double)count_y / n_y;
    if (!have_min || z < min_z) { min_z = z; have_min = true; best_x = count_x; best_y = count_y; }
  }
  double statistic = -min_z;
  if (((double)n_x * n_y) < 10000) {
    uint32_t m = n_y, n = n_x;
    int64_t threshold = (int64_t)best_y * n - (int64_t)best_x * m;
    vector<vector<double> > u(m + 1, vector<double>(n + 1, 0.0));
    for (uint32_t a = 0; a <= m; a++) for (uint32_t b = 0; b <= n; b++) {
      double value;
      if (a == 0 && b == 0) value = 1.0;
      else value = ((a > 0) ? u[a - 1][b] : 0.0) + ((b > 0) ? u[a][b - 1] : 0.0);
      int64_t level = (int64_t)a * n - (int64_t)b * m;
      if (boundary[a + b] && level >= threshold) value = 0.0;
      u[a][b] = value;
    }
    double total_paths = exp(log_choose(m + n, m));
    double p = 1.0 - u[m][n] / total_paths;
    return min(max(p, 0.0), 1.0);
  }
  double n_eff = ((double)n_x * n_y) / (n_x + n_y);
  return exp(-2.0 * statistic * statistic * n_eff);
}

// ---------------------------------------------------------------- GenomeDiff subset
struct GdEntry {
  string type;               // "RA", "MC", "UN"
  uint32_t id = 0;
  vector<string> spec;       // positional fields after type/id/evidence
  map<string, string> kv;    // remaining key=value fields (std::map => alphabetical)
};
static int type_rank(const string& t) { return t == "RA" ? 3 : t == "MC" ? 4 : 7; }  // genome_diff_entry.cpp:280-308
static bool gd_less(const GdEntry& a, const GdEntry& b) {                            // genome_diff_entry.cpp:566-700
  if (type_rank(a.type) != type_rank(b.type)) return type_rank(a.type) < type_rank(b.type);
  if (a.spec[0] != b.spec[0]) return a.spec[0] < b.spec[0];
  uint32_t pa = (uint32_t)strtoul(a.spec[1].c_str(), NULL, 10), pb = (uint32_t)strtoul(b.spec[1].c_str(), NULL, 10);
  if (pa != pb) return pa < pb;
  for (size_t k = 2; k < a.spec.size(); ++k) {
    bool numeric = (a.type == "RA") ? (k == 2) : true;  // insert_position / end / ranges are integers, bases are strings
    if (numeric) {
      long va = strtol(a.spec[k].c_str(), NULL, 10), vb = strtol(b.spec[k].c_str(), NULL, 10);
      if (va != vb) return va < vb;
    } else if (a.spec[k] != b.spec[k]) return a.spec[k] < b.spec[k];
  }
  return a.id < b.id;
}
static void write_gd(const string& fn, vector<GdEntry>& entries) {  // genome_diff.cpp:685-760
  ofstream os(fn.c_str());
  os << "#=GENOME_DIFF\t1.0" << endl;
  stable_sort(entries.begin(), entries.end(), gd_less);
  for (const GdEntry& e : entries) {
    os << e.type << "\t" << e.id << "\t.";
    for (const string& s : e.spec) os << "\t" << s;
    for (const auto& kv : e.kv) if (!kv.second.empty()) os << "\t" << kv.first << "=" << kv.second;
    os << endl;
  }
}

// ---------------------------------------------------------------- pass 2
struct PositionCoverage {  // identify_mutations.h:99-126
  double unique[3], redundant[3];
  int raw_redundant[3], total;
  PositionCoverage() { memset(this, 0, sizeof(*this)); }
  explicit PositionCoverage(double v) { memset(this, 0, sizeof(*this)); for (int i = 0; i < 3; i++) unique[i] = redundant[i] = v; }
  void sum() {
    unique[1] = unique[0] + unique[2];
    redundant[1] = redundant[0] + redundant[2];
    raw_redundant[1] = raw_redundant[0] + raw_redundant[2];
    total = (int)round(unique[1]) + (int)round(redundant[1]);
  }
};
struct PolyData {  // identify_mutations.h:131-157
  char base_char; uint8_t quality; int strand; int32_t mapping_quality; cv_t cv;
  double log10_pr[5], r[5], log10_pr_max;
};
struct AlleleModel {  // identify_mutations.h:170-194
  double f[5], log10_likelihood = 0.0, sum_w[5]; uint32_t n = 0, iterations = 0;
  AlleleModel() { for (int b = 0; b < 5; b++) f[b] = sum_w[b] = 0.0; }
  double reported_frequency(uint8_t b) const {  // :3042-3046
    if (n == 0 || b >= base_list_size) return 0.0;
    return (f[b] < 0.5 / (double)n) ? 0.0 : f[b];
  }
  uint8_t major_index() const {  // :3060-3068
    if (n == 0) return base_list_N_index;
    uint8_t best = 0;
    for (uint8_t b = 1; b < base_list_size; b++) if (f[b] > f[best]) best = b;
    return (f[best] > 0.0) ? best : base_list_N_index;
  }
  uint8_t next_index(uint8_t exclude) const {  // :3070-3086
    if (n == 0) return base_list_N_index;
    const double thr = 0.5 / (double)n;
    uint8_t best = base_list_N_index;
    for (uint8_t b = 0; b < base_list_size; b++) {
      if (b == exclude) continue;
      if (f[b] < thr) continue;
      if (best == base_list_N_index || f[b] > f[best]) best = b;
    }
    return best;
  }
  string spectrum_string(uint32_t places) const {  // :3048-3058
    string s;
    for (uint8_t b = 0; b < base_list_size; b++) {
      double fr = reported_frequency(b);
      if (fr <= 0.0) continue;
      if (!s.empty()) s += ",";
      s += string(1, baseindex2char(b)) + ":" + to_string_double(fr, places, true);
    }
    return s;
  }
};

static const double kProfileLikelihoodLog10Drop = 0.587566;  // :3173
static const uint32_t UNDEF = 0xFFFFFFFFu;

struct IdentifyMutationsPileup : PileupDriver {
  const Settings& settings;
  vector<double> seed_cutoffs, propagation_cutoffs;
  double consensus_cutoff, polymorphism_cutoff, precision_decimal;
  uint32_t precision_places;
  double log10_ref_length = 0;
  ErrorTable table;
  vector<GdEntry> gd;
  uint32_t id_counter = 0;
  FILE* dump = NULL;
  // MC / UN state (:2054-2059)
  double this_prop_cutoff = 0, this_seed_cutoff = 0;
  uint32_t last_del_start = UNDEF, last_del_end = UNDEF, last_del_red_start = UNDEF, last_del_red_end = UNDEF, last_start_unknown = UNDEF;
  bool reaches_seed = false, red_reached_zero = false;
  PositionCoverage last_cov, left_outside, left_inside;

  IdentifyMutationsPileup(const Settings& s, const string& bam, const string& fasta, const vector<double>& prop,
                          const vector<double>& seed, double mut_cut, double poly_cut, double prec, uint32_t places)
      : PileupDriver(bam, fasta), settings(s), seed_cutoffs(seed), propagation_cutoffs(prop), consensus_cutoff(mut_cut),
        polymorphism_cutoff(poly_cut), precision_decimal(prec), precision_places(places) {
    ORACLE_ASSERT(hdr->n_targets == (int32_t)prop.size() && hdr->n_targets == (int32_t)seed.size(), "cutoff table size");
    for (int i = 0; i < hdr->n_targets; ++i) log10_ref_length += (double)hdr->target_len[i];  // :848-852
    log10_ref_length = log10(log10_ref_length);
    table.read_log10_prob_table(s.error_rates_file_name);  // :865-866
    table.log10_prob_to_prob();
    table.rg_map = &read_groups;
    table.partition = ReadFilePartition::make(read_groups, s.read_file_sets);
    if (!s.user_evidence_genome_diff_file_name.empty()) load_user_ra_evidence_from_gd(s.user_evidence_genome_diff_file_name);  // :878-881
  }
  // :1013-1020: the RA rows of the user's GenomeDiff, stripped to their specification, in cGenomeDiff::sort() order
  // (cDiffEntry::compare, genome_diff_entry.cpp:566-690: seq_id as a string, position, then the specification fields)
  std::deque<GdEntry> user_evidence_ra_list;
  void load_user_ra_evidence_from_gd(const string& path) {
    FILE* f = fopen(path.c_str(), "r");
    ORACLE_ASSERT(f != nullptr, "cannot open the user evidence GenomeDiff");
    char* line = nullptr; size_t cap = 0; ssize_t n;
    vector<GdEntry> rows;
    while ((n = getline(&line, &cap, f)) >= 0) {
      while (n > 0 && (line[n - 1] == '\n' || line[n - 1] == '\r')) line[--n] = 0;
      vector<string> fld;
      for (char* p = line;;) { char* t = strchr(p, '\t'); fld.emplace_back(p, t ? (size_t)(t - p) : strlen(p)); if (!t) break; p = t + 1; }
      if (fld.size() < 8 || fld[0] != "RA") continue;
      GdEntry e; e.type = "RA"; e.spec = {fld[3], fld[4], fld[5], fld[6], fld[7]};
      rows.push_back(e);
    }
    free(line); fclose(f);
    std::stable_sort(rows.begin(), rows.end(), [](const GdEntry& a, const GdEntry& b) {
      if (a.spec[0] != b.spec[0]) return a.spec[0] < b.spec[0];
      const unsigned long pa = strtoul(a.spec[1].c_str(), nullptr, 10), pb = strtoul(b.spec[1].c_str(), nullptr, 10);
      if (pa != pb) return pa < pb;
      const unsigned long ia = strtoul(a.spec[2].c_str(), nullptr, 10), ib = strtoul(b.spec[2].c_str(), nullptr, 10);
      if (ia != ib) return ia < ib;
      if (a.spec[3] != b.spec[3]) return a.spec[3] < b.spec[3];
      return a.spec[4] < b.spec[4];
    });
    user_evidence_ra_list.assign(rows.begin(), rows.end());
  }
  void add(GdEntry e) { e.id = ++id_counter; gd.push_back(e); }

  void fill_read_base_likelihoods(PolyData& pd) const {  // :3359-3384
    const double incorrect = pow(10, -(double)pd.mapping_quality / 10);
    const double correct = 1 - incorrect;
    const double uniform = 1.0 / (double)base_list_size;
    cv_t cv;
    memcpy(cv, pd.cv, sizeof cv);
    const bool top = (pd.strand == 1);
    if (!top) cv[k_obs_base] = complement_base_index((uint8_t)cv[k_obs_base]);
    pd.log10_pr_max = -numeric_limits<double>::max();
    for (uint8_t b = 0; b < base_list_size; b++) {
      cv[k_ref_base] = top ? b : complement_base_index(b);
      double pr = correct * table.get_prob(cv) + incorrect * uniform;
      if (pr < 0.0) pr = 0.0;
      pd.log10_pr[b] = log10(pr);
      pd.log10_pr_max = max(pd.log10_pr_max, pd.log10_pr[b]);
    }
    for (uint8_t b = 0; b < base_list_size; b++) pd.r[b] = pow(10, pd.log10_pr[b] - pd.log10_pr_max);
  }

  AlleleModel fit_allele_frequencies(const vector<PolyData>& pdata, const bool allowed[5]) const {  // :3240-3318
    AlleleModel m;
    m.n = (uint32_t)pdata.size();
    if (m.n == 0) return m;
    double init_total = 0.0;
    uint32_t n_allowed = 0;
    for (uint8_t b = 0; b < 5; b++) { if (!allowed[b]) continue; n_allowed++; m.f[b] = 0.5; }
    if (n_allowed == 0) return m;
    for (size_t i = 0; i < pdata.size(); ++i) { uint8_t obs = basechar2index(pdata[i].base_char); if (obs < 5 && allowed[obs]) m.f[obs] += 1.0; }
    for (uint8_t b = 0; b < 5; b++) init_total += m.f[b];
    for (uint8_t b = 0; b < 5; b++) m.f[b] /= init_total;
    const uint32_t k_max_iterations = 50;
    for (m.iterations = 1; m.iterations <= k_max_iterations; m.iterations++) {
      double sum_w[5] = {0, 0, 0, 0, 0};
      double ll = 0.0;
      for (size_t i = 0; i < pdata.size(); ++i) {
        const PolyData& it = pdata[i];
        double s = 0.0;
        for (uint8_t b = 0; b < 5; b++) if (allowed[b]) s += m.f[b] * it.r[b];
        if (s > 0.0) {
          ll += log10(s) + it.log10_pr_max;
          for (uint8_t b = 0; b < 5; b++) { if (!allowed[b]) continue; sum_w[b] += m.f[b] * it.r[b] / s; }
        } else {
          for (uint8_t b = 0; b < 5; b++) { if (!allowed[b]) continue; sum_w[b] += m.f[b]; }
        }
      }
      double max_delta = 0.0;
      for (uint8_t b = 0; b < 5; b++) {
        if (!allowed[b]) continue;
        double f_new = sum_w[b] / (double)m.n;
        max_delta = max(max_delta, fabs(f_new - m.f[b]));
        m.f[b] = f_new;
      }
      for (uint8_t b = 0; b < 5; b++) m.sum_w[b] = sum_w[b];
      m.log10_likelihood = ll;
      if (max_delta < precision_decimal) break;
    }
    if (m.iterations > k_max_iterations) m.iterations = k_max_iterations;
    return m;
  }
  double variant_presence_score(const vector<PolyData>& pdata, const AlleleModel& full, uint8_t variant) const {  // :3329-3344
    if (full.n == 0 || variant >= 5) return numeric_limits<double>::quiet_NaN();
    bool without[5];
    for (uint8_t b = 0; b < 5; b++) without[b] = (b != variant);
    AlleleModel null_fit = fit_allele_frequencies(pdata, without);
    return (full.log10_likelihood - null_fit.log10_likelihood) - log10_ref_length;
  }
  double profile_log10_likelihood(const vector<PolyData>& pdata, const AlleleModel& full, uint8_t variant, double f_fixed) const {  // :3094-3143
    if (full.n == 0 || variant >= 5) return 0.0;
    double f[5], other_total = 0.0;
    for (uint8_t b = 0; b < 5; b++) if (b != variant) other_total += full.f[b];
    for (uint8_t b = 0; b < 5; b++) {
      if (b == variant) f[b] = f_fixed;
      else f[b] = (other_total > 0.0) ? (1.0 - f_fixed) * full.f[b] / other_total : (1.0 - f_fixed) / (double)(5 - 1);
    }
    double ll = 0.0;
    for (uint32_t iter = 0; iter < 50; iter++) {
      double sum_w[5] = {0, 0, 0, 0, 0};
      ll = 0.0;
      for (size_t i = 0; i < pdata.size(); ++i) {
        const PolyData& it = pdata[i];
        double s = 0.0;
        for (uint8_t b = 0; b < 5; b++) s += f[b] * it.r[b];
        if (s > 0.0) { ll += log10(s) + it.log10_pr_max; for (uint8_t b = 0; b < 5; b++) sum_w[b] += f[b] * it.r[b] / s; }
        else for (uint8_t b = 0; b < 5; b++) sum_w[b] += f[b];
      }
      double others = 0.0;
      for (uint8_t b = 0; b < 5; b++) if (b != variant) others += sum_w[b];
      double max_delta = 0.0;
      for (uint8_t b = 0; b < 5; b++) {
        if (b == variant) continue;
        double f_new = (others > 0.0) ? (1.0 - f_fixed) * sum_w[b] / others : (1.0 - f_fixed) / (double)(5 - 1);
        max_delta = max(max_delta, fabs(f_new - f[b]));
        f[b] = f_new;
      }
      if (max_delta < precision_decimal) break;
    }
    return ll;
  }
  void frequency_bounds(const vector<PolyData>& pdata, const AlleleModel& am, uint8_t variant, double& lower, double& upper) const {  // :3175-3217
    lower = 0.0; upper = 1.0;
    if (am.n > 0 && variant < 5) {
      const double f_hat = am.f[variant];
      const double pl_max = profile_log10_likelihood(pdata, am, variant, f_hat);
      const double target = pl_max - kProfileLikelihoodLog10Drop;
      if (profile_log10_likelihood(pdata, am, variant, 0.0) >= target) lower = 0.0;
      else {
        double lo = 0.0, hi = f_hat;
        for (uint32_t i = 0; i < 40 && (hi - lo) > precision_decimal; i++) {
          double mid = 0.5 * (lo + hi);
          if (profile_log10_likelihood(pdata, am, variant, mid) >= target) hi = mid; else lo = mid;
        }
        lower = hi;
      }
      if (profile_log10_likelihood(pdata, am, variant, 1.0) >= target) upper = 1.0;
      else {
        double lo = f_hat, hi = 1.0;
        for (uint32_t i = 0; i < 40 && (hi - lo) > precision_decimal; i++) {
          double mid = 0.5 * (lo + hi);
          if (profile_log10_likelihood(pdata, am, variant, mid) >= target) lo = mid; else hi = mid;
        }
        upper = lo;
      }
    }
  }

  void check_deletion_completion(uint32_t seq_id, uint32_t position, const PositionCoverage& cov, double) {  // :2262-2344
    if (position == 1) last_cov = PositionCoverage(numeric_limits<double>::quiet_NaN());
    if (cov.unique[1] <= this_prop_cutoff) {
      if (last_del_start == UNDEF) { last_del_start = position; left_outside = last_cov; left_inside = cov; }
    }
    if (!std::isnan(cov.unique[1]) && (cov.total <= this_seed_cutoff)) reaches_seed = true;
    if (last_del_start != UNDEF && (std::isnan(cov.unique[1]) || cov.unique[1] > this_prop_cutoff)) {
      if (reaches_seed) {
        last_del_end = position - 1;
        if (last_del_red_end == UNDEF) last_del_red_end = last_del_end;
        if (last_del_red_start == UNDEF) last_del_red_start = last_del_start;
        GdEntry del;
        del.type = "MC";
        del.spec = {target_name(seq_id), std::to_string(last_del_start), std::to_string(last_del_end),
                    std::to_string(last_del_red_start - last_del_start), std::to_string(last_del_end - last_del_red_end)};
        del.kv["left_outside_cov"] = to_string_double(left_outside.unique[1], 0);
        del.kv["left_inside_cov"] = to_string_double(left_inside.unique[1], 0);
        del.kv["right_inside_cov"] = to_string_double(last_cov.unique[1], 0);
        del.kv["right_outside_cov"] = to_string_double(cov.unique[1], 0);
        add(del);
      }
      reaches_seed = false; red_reached_zero = false;
      last_del_start = last_del_end = last_del_red_start = last_del_red_end = UNDEF;
    }
    if (last_del_start != UNDEF) {
      if (cov.redundant[1] == 0) { red_reached_zero = true; last_del_red_end = UNDEF; }
      else if (cov.redundant[1] > 0) {
        if (!red_reached_zero) last_del_red_start = position;
        else if (last_del_red_end == UNDEF) last_del_red_end = position;
      }
    }
    last_cov = cov;
  }
  void update_unknown_intervals(uint32_t position, uint32_t seq_id, bool base_predicted, bool) {  // :2972-3007
    if (!base_predicted) {
      if (last_start_unknown == UNDEF) last_start_unknown = position;
    } else if (last_start_unknown != UNDEF) {
      GdEntry un;
      un.type = "UN";
      un.spec = {target_name(seq_id), std::to_string(last_start_unknown), std::to_string(position - 1)};
      add(un);
      last_start_unknown = UNDEF;
    }
  }
  void at_target_start(uint32_t) {  // :2054-2059
    last_del_start = last_del_end = last_del_red_start = last_del_red_end = last_start_unknown = UNDEF;
  }
  void at_target_end(uint32_t tid) {  // :2117-2164
    if (!settings.skip_missing_coverage_prediction)
      check_deletion_completion(tid, target_length(tid) + 1, PositionCoverage(numeric_limits<double>::quiet_NaN()), numeric_limits<double>::quiet_NaN());
    update_unknown_intervals(target_length(tid) + 1, tid, true, false);
    if (!settings.skip_missing_coverage_prediction && propagation_cutoffs[tid] < 0.0) {
      GdEntry del;
      del.type = "MC";
      del.spec = {target_name(tid), "1", std::to_string(target_length(tid)), "0", "0"};
      del.kv["left_outside_cov"] = "NA";
      del.kv["left_inside_cov"] = to_string_double(0.0, 0);
      del.kv["right_inside_cov"] = to_string_double(0.0, 0);
      del.kv["right_outside_cov"] = "NA";
      add(del);
    }
  }

  void pileup_callback(uint32_t tid, uint32_t position, int n, const bam_pileup1_t* pile) {  // :1309-2022
    this_prop_cutoff = propagation_cutoffs[tid];
    this_seed_cutoff = seed_cutoffs[tid];
    if (this_prop_cutoff < 0.0) return;
    int32_t insert_count = -1;
    bool next_insert_count_exists = true;
    // :1346-1355: levels the user list forces here; the front run is matched by POSITION alone, the last entry's level wins
    int32_t force_insert_count_max = 0;
    for (auto u = user_evidence_ra_list.begin(); u != user_evidence_ra_list.end() && strtoul(u->spec[1].c_str(), nullptr, 10) == position; ++u)
      force_insert_count_max = (int32_t)strtol(u->spec[2].c_str(), nullptr, 10);
    while (next_insert_count_exists || insert_count < force_insert_count_max) {
      ++insert_count;
      next_insert_count_exists = false;
      char ref_base_char = '.';
      if (!insert_count) ref_base_char = refs[tid][position - 1];
      uint32_t pos_info[6][3];
      memset(pos_info, 0, sizeof pos_info);
      PositionCoverage cov;
      bool unique_only = true;
      vector<PolyData> pdata;
      double log10_pr_sum[5] = {0, 0, 0, 0, 0};
      for (int k = 0; k < n; ++k) {
        Aln i(&pile[k]);
        int indel = i.indel();
        if (indel < 0) indel = 0;
        if (i.is_del()) indel = -1;
        uint8_t read_base_bam = '.';
        bool past_base = true;
        if (indel >= insert_count) { read_base_bam = i.base_bam_0(i.qpos0() + (uint32_t)insert_count); past_base = false; }
        if (bam_is_N(read_base_bam)) continue;
        int32_t redundancy = (int32_t)i.redundancy();
        int strand = i.strand();
        bool trimmed = i.is_trimmed(past_base);
        if (redundancy == 1) {
          ++cov.unique[1 + strand];
          if (indel > insert_count) next_insert_count_exists = true;
        } else {
          unique_only = false;
          cov.redundant[1 + strand] += 1.0 / redundancy;
          ++cov.raw_redundant[1 + strand];
        }
        if (redundancy > 1) continue;
        if (trimmed) continue;
        PolyData pd;
        memset(pd.cv, 0, sizeof pd.cv);
        bool is_ok = table.alignment_position_to_covariates(i, insert_count, pd.cv);
        if (is_ok) {
          if (pd.cv[k_quality] < settings.base_quality_cutoff) continue;
          ++pos_info[pd.cv[k_obs_base]][1 + strand];
          pd.base_char = baseindex2char((uint8_t)pd.cv[k_obs_base]);
          pd.quality = (uint8_t)pd.cv[k_quality];
          pd.strand = strand;
          pd.mapping_quality = (int32_t)i.a->core.qual;
          fill_read_base_likelihoods(pd);
          pdata.push_back(pd);
          for (uint8_t j = 0; j < 5; j++) log10_pr_sum[j] += pd.log10_pr[j];
        }
      }
      cov.sum();

      // pure_genotype_call :3398-3433
      char best_base_char = 'N';
      double snp_score = numeric_limits<double>::quiet_NaN();
      if (!pdata.empty()) {
        uint8_t best = 0;
        for (uint8_t b = 1; b < 5; b++) if (log10_pr_sum[b] > log10_pr_sum[best]) best = b;
        best_base_char = baseindex2char(best);
        double off = -numeric_limits<double>::max();
        for (uint8_t b = 0; b < 5; b++) if (b != best) off = max(off, log10_pr_sum[b]);
        double total_err = 0;
        for (uint8_t b = 0; b < 5; b++) if (b != best) total_err += pow(10, log10_pr_sum[b] - off);
        double log10_total_err = log10(total_err);
        log10_total_err += off;
        snp_score = log10_pr_sum[best] - log10_total_err;
      }
      double consensus_score = snp_score - log10_ref_length;
      bool base_predicted = (consensus_score >= consensus_cutoff);
      int total_cov[3] = {0, 0, 0};
      for (size_t j = 0; j < 5; ++j) { total_cov[2] += (int)round((double)pos_info[j][2]); total_cov[0] += (int)round((double)pos_info[j][0]); }

      if (insert_count == 0 && !settings.skip_missing_coverage_prediction) check_deletion_completion(tid, position, cov, consensus_score);

      bool passed_poly = false;
      bool passed_consensus = (best_base_char != ref_base_char) && (!std::isnan(consensus_score) && consensus_score > 0);
      bool all_alleles[5] = {true, true, true, true, true};
      AlleleModel amodel = fit_allele_frequencies(pdata, all_alleles);
      const uint8_t ref_index = (ref_base_char == 'N') ? base_list_N_index : basechar2index(ref_base_char);
      const uint8_t major_index = amodel.major_index();
      const uint8_t minor_index = amodel.next_index(major_index);
      const uint8_t variant_index = amodel.next_index(ref_index);
      char major_char = baseindex2char(major_index), minor_char = baseindex2char(minor_index), variant_char = baseindex2char(variant_index);
      double variant_score = numeric_limits<double>::quiet_NaN();
      if (variant_index != base_list_N_index) {
        variant_score = variant_presence_score(pdata, amodel, variant_index);
        if (variant_score >= polymorphism_cutoff) passed_poly = true;
      }
      if (insert_count == 0) update_unknown_intervals(position, tid, base_predicted, unique_only);

      bool emitted = passed_consensus || passed_poly;
      if (emitted) {  // :1836-1910
        GdEntry mut;
        mut.type = "RA";
        mut.spec = {target_name(tid), std::to_string(position), std::to_string(insert_count), string(1, ref_base_char), string(1, variant_char)};
        mut.kv["score"] = to_string_double(variant_score, 1);
        mut.kv["major_base"] = string(1, major_char);
        mut.kv["minor_base"] = string(1, minor_char);
        mut.kv["major_frequency"] = to_string_double(amodel.reported_frequency(major_index), precision_places, true);
        mut.kv["frequency"] = to_string_double(amodel.reported_frequency(variant_index), precision_places, true);
        mut.kv["allele_frequencies"] = amodel.spectrum_string(precision_places);
        double lower, upper;
        frequency_bounds(pdata, amodel, variant_index, lower, upper);
        mut.kv["frequency_lower"] = to_string_double(lower, precision_places, true);
        mut.kv["frequency_upper"] = to_string_double(upper, precision_places, true);
        {  // annotate_polymorphism_statistics :3009-3033
          uint8_t mj = basechar2index(major_char), mn = basechar2index(minor_char);
          vector<double> major_quals, minor_quals;
          for (size_t i = 0; i < pdata.size(); ++i) {
            if (pdata[i].base_char == major_char) major_quals.push_back((double)pdata[i].quality);
            if (pdata[i].base_char == minor_char) minor_quals.push_back((double)pdata[i].quality);
          }
          double ks = 1.0;
          if (!major_quals.empty() && !minor_quals.empty()) ks = ks_test_two_sample_less(minor_quals, major_quals);
          double fisher = fisher_exact_test_2x2(pos_info[mn][2], pos_info[mn][0], pos_info[mj][2], pos_info[mj][0]);
          mut.kv["ks_quality_p_value"] = to_string_double(ks, 5, true);
          mut.kv["fisher_strand_p_value"] = to_string_double(fisher, 5, true);
        }
        auto covs = [&](char c) { uint8_t b = basechar2index(c); return std::to_string((int32_t)pos_info[b][2]) + "/" + std::to_string((int32_t)pos_info[b][0]); };
        mut.kv["ref_cov"] = covs(ref_base_char);
        mut.kv["new_cov"] = covs(variant_char);
        mut.kv["major_cov"] = covs(major_char);
        mut.kv["minor_cov"] = covs(minor_char);
        mut.kv["total_cov"] = std::to_string(total_cov[2]) + "/" + std::to_string(total_cov[0]);
        add(mut);
      }
      // :1914-2019: user evidence at this position and insert level that the data did not already report
      while (!user_evidence_ra_list.empty() && user_evidence_ra_list.front().spec[0] == target_name(tid) &&
             strtoul(user_evidence_ra_list.front().spec[1].c_str(), nullptr, 10) == position &&
             strtol(user_evidence_ra_list.front().spec[2].c_str(), nullptr, 10) == insert_count) {
        const GdEntry& user = user_evidence_ra_list.front();
        if (emitted && gd.back().type == "RA" && gd.back().spec == user.spec) {
          gd.back().kv["user_defined"] = "1";
        } else {
          GdEntry mut;
          mut.type = "RA";
          mut.spec = user.spec;
          mut.kv["user_defined"] = "1";
          const char user_ref = user.spec[3][0], user_new = user.spec[4][0];
          const uint8_t user_ref_index = basechar2index(user_ref), user_variant_index = basechar2index(user_new);
          double user_score = numeric_limits<double>::quiet_NaN(), user_variant_frequency = 0.0;
          if (user_variant_index < 5) {
            user_score = variant_presence_score(pdata, amodel, user_variant_index);
            user_variant_frequency = amodel.reported_frequency(user_variant_index);
          }
          const double user_ref_frequency = amodel.reported_frequency(user_ref_index);
          const bool variant_is_major = user_variant_frequency > user_ref_frequency;
          const char mj = variant_is_major ? user_new : user_ref, mn = variant_is_major ? user_ref : user_new;
          mut.kv["major_base"] = string(1, mj);
          mut.kv["minor_base"] = string(1, mn);
          mut.kv["major_frequency"] = to_string_double(variant_is_major ? user_variant_frequency : user_ref_frequency, precision_places, true);
          mut.kv["frequency"] = to_string_double(user_variant_frequency, precision_places, true);
          mut.kv["allele_frequencies"] = amodel.spectrum_string(precision_places);
          double lower, upper;
          frequency_bounds(pdata, amodel, user_variant_index, lower, upper);
          mut.kv["frequency_lower"] = to_string_double(lower, precision_places, true);
          mut.kv["frequency_upper"] = to_string_double(upper, precision_places, true);
          mut.kv["prediction"] = settings.polymorphism_prediction ? "polymorphism" : (user_variant_frequency > 0.5 ? "consensus" : "polymorphism");
          mut.kv["score"] = to_string_double(user_score, 1);
          auto covs = [&](char c) { uint8_t b = basechar2index(c); return std::to_string((int32_t)pos_info[b][2]) + "/" + std::to_string((int32_t)pos_info[b][0]); };
          mut.kv["ref_cov"] = covs(user_ref);
          mut.kv["new_cov"] = covs(user_new);
          mut.kv["major_cov"] = covs(mj);
          mut.kv["minor_cov"] = covs(mn);
          mut.kv["total_cov"] = std::to_string(total_cov[2]) + "/" + std::to_string(total_cov[0]);
          add(mut);
        }
        user_evidence_ra_list.pop_front();
      }
      if (dump) {
        ColumnDump d;
        memset(&d, 0, sizeof d);
        d.tid = tid; d.pos1 = position; d.insert_count = (uint32_t)insert_count; d.n = (uint32_t)pdata.size();
        for (int b = 0; b < 5; b++) { d.ll[b] = log10_pr_sum[b]; d.f[b] = amodel.f[b]; }
        d.consensus_score = consensus_score; d.variant_score = variant_score; d.log10_likelihood = amodel.log10_likelihood;
        d.unique[0] = cov.unique[0]; d.unique[1] = cov.unique[2]; d.redundant[0] = cov.redundant[0]; d.redundant[1] = cov.redundant[2];
        d.raw_redundant[0] = cov.raw_redundant[0]; d.raw_redundant[1] = cov.raw_redundant[2]; d.total = cov.total;
        d.best = basechar2index(best_base_char); d.major = major_index; d.minor = minor_index; d.variant = variant_index;
        d.ref = (ref_base_char == 'N') ? 5 : basechar2index(ref_base_char);
        d.base_predicted = base_predicted; d.unique_only = unique_only; d.emitted = emitted; d.iterations = amodel.iterations;
        fwrite(&d, sizeof d, 1, dump);
      }
    }
  }
};

void identify_mutations(const Settings& settings, const string& bam, const string& fasta, const string& gd_file,
                        const vector<double>& prop, const vector<double>& seed, double mutation_cutoff,
                        double polymorphism_cutoff, double precision_decimal, uint32_t precision_places,
                        const string& columns_dump_file, uint64_t* n_records_out) {
  IdentifyMutationsPileup imp(settings, bam, fasta, prop, seed, mutation_cutoff, polymorphism_cutoff, precision_decimal, precision_places);
  if (!columns_dump_file.empty()) imp.dump = fopen(columns_dump_file.c_str(), "wb");
  imp.do_pileup(settings.call_mutations_seq_ids);
  if (imp.dump) fclose(imp.dump);
  write_gd(gd_file, imp.gd);
  if (n_records_out) *n_records_out = imp.n_records;
}

}  // namespace oracle
