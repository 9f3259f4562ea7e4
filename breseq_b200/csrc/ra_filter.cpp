// See ra_filter.h.  A row is held the way the reference holds it -- every column and key=value field in one ordered string map
// (cDiffEntry is a map<string,string>, genome_diff_entry.h:150) -- because the filter's reads of absent keys (operator[]) and
// the writer's rule "fixed columns first, then whatever is left in key order, empty values skipped" (cDiffEntry::marshal,
// genome_diff_entry.cpp:1323-1369) are part of what the output looks like.
#include "ra_filter.h"

#include "coverage_fit.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <sstream>
#include <stdexcept>
#include <vector>

namespace brq {
namespace {

typedef std::map<std::string, std::string> Row;
const char* const RA_COLUMNS[] = {"seq_id", "position", "insert_position", "ref_base", "new_base"};   // line_specification[RA]

std::vector<std::string> split(const std::string& s, char sep) {   // common.h:648-679: an empty string has no parts
  std::vector<std::string> out;
  if (s.empty()) return out;
  size_t start = 0;
  for (;;) {
    size_t end = s.find(sep, start);
    out.push_back(s.substr(start, end == std::string::npos ? std::string::npos : end - start));
    if (end == std::string::npos) break;
    start = end + 1;
  }
  return out;
}

template <typename T>
T number(const std::string& s) {   // from_string<T>, common.h:892-899
  if (s.empty()) throw std::runtime_error("evidence row: a number is expected where a field is empty or absent");
  T t = T();
  std::istringstream iss(s);
  iss >> t;
  return t;
}

double number_or_na(const std::string& s) {   // double_from_string, common.h:939-950
  std::string u;
  for (char c : s) u.push_back((char)toupper((unsigned char)c));
  if (u == "NA" || u == "#NA" || u == "NAN") return std::numeric_limits<double>::quiet_NaN();
  if (u == "INF") return std::numeric_limits<double>::infinity();
  if (u == "-INF") return -std::numeric_limits<double>::infinity();
  return number<double>(s);
}

bool has(const Row& r, const char* k) { return r.count(k) > 0; }

void add_reject_reason(Row& r, const char* reason) {   // genome_diff_entry.cpp:1399-1416
  if (!has(r, "reject")) { r["reject"] = reason; return; }
  std::vector<std::string> now = split(r["reject"], ',');
  for (const std::string& s : now) if (s == reason) return;
  now.push_back(reason);
  std::string joined;
  for (size_t i = 0; i < now.size(); ++i) joined += (i ? "," : "") + now[i];
  r["reject"] = joined;
}

double strand_sum(Row& r, const char* k, double* top = nullptr, double* bot = nullptr) {
  std::vector<std::string> tb = split(r[k], '/');
  if (tb.size() < 2) throw std::runtime_error(std::string("evidence row: ") + k + " is not top/bottom");
  const double t = number<double>(tb[0]), b = number<double>(tb[1]);
  if (top) *top = t;
  if (bot) *bot = b;
  return t + b;
}

// ---- the reference sequence the way FastaSequence::get_sequence_1 answers (fasta.h:58-95): (0,0) is empty, anything else is
// pulled into the sequence
struct Reference {
  const RefSet& ref;
  const std::string& seq(const std::string& id) const {
    for (size_t i = 0; i < ref.names.size(); ++i) if (ref.names[i] == id) return ref.seqs[i];
    throw std::runtime_error("evidence row names a sequence the FASTA does not hold: " + id);
  }
  static std::string base(const std::string& s, int64_t p) {
    if (p == 0 || s.empty()) return std::string();
    if (p < 1) p = 1;
    if (p > (int64_t)s.size()) p = (int64_t)s.size();
    return std::string(1, s[(size_t)p - 1]);
  }
};

// identify_mutations.cpp:129-172
void polymorphism_bias(Row& r, const RaFilterOptions& o) {
  if (o.polymorphism_ks_quality_p_value_cutoff != 0 && has(r, "ks_quality_p_value") &&
      number<double>(r["ks_quality_p_value"]) < o.polymorphism_ks_quality_p_value_cutoff)
    add_reject_reason(r, "KS_BASE_QUALITY");
  if (o.polymorphism_fisher_strand_p_value_cutoff != 0 && has(r, "fisher_strand_p_value") &&
      number<double>(r["fisher_strand_p_value"]) < o.polymorphism_fisher_strand_p_value_cutoff)
    add_reject_reason(r, "FISHER_STRAND");
}

// identify_mutations.cpp:231-322: only an allele that differs from the reference base has to show on both strands
void polymorphism_coverage(Row& r, const RaFilterOptions& o) {
  if (o.polymorphism_minimum_variant_coverage_each_strand > 0) {
    const double need = o.polymorphism_minimum_variant_coverage_each_strand;
    bool passed = true;
    const std::string ref_base = has(r, "ref_base") ? r["ref_base"] : "";
    bool major_is_variant = has(r, "major_base") && r["major_base"] != ref_base;
    bool minor_is_variant = has(r, "minor_base") && r["minor_base"] != ref_base;
    if (!major_is_variant && !minor_is_variant) major_is_variant = minor_is_variant = true;
    double top, bot;
    if (major_is_variant && has(r, "major_cov")) { strand_sum(r, "major_cov", &top, &bot); passed = passed && top >= need && bot >= need; }
    if (minor_is_variant && has(r, "minor_cov")) { strand_sum(r, "minor_cov", &top, &bot); passed = passed && top >= need && bot >= need; }
    if (!passed) add_reject_reason(r, "VARIANT_STRAND_COVERAGE");
  }
  if (o.polymorphism_minimum_total_coverage_each_strand > 0) {
    double top, bot;
    strand_sum(r, "total_cov", &top, &bot);
    if (!(top >= o.polymorphism_minimum_total_coverage_each_strand && bot >= o.polymorphism_minimum_total_coverage_each_strand))
      add_reject_reason(r, "TOTAL_STRAND_COVERAGE");
  }
  if (o.polymorphism_minimum_variant_coverage > 0) {
    const bool major = strand_sum(r, "major_cov") >= o.polymorphism_minimum_variant_coverage;
    const bool minor = strand_sum(r, "minor_cov") >= o.polymorphism_minimum_variant_coverage;
    if (!(major && minor)) add_reject_reason(r, "VARIANT_COVERAGE");
  }
  if (o.polymorphism_minimum_total_coverage > 0 && strand_sum(r, "total_cov") < o.polymorphism_minimum_total_coverage)
    add_reject_reason(r, "TOTAL_COVERAGE");
}

// identify_mutations.cpp:324-367
void consensus_coverage(Row& r, const RaFilterOptions& o) {
  double top, bot;
  if (o.consensus_minimum_variant_coverage_each_strand > 0) {
    strand_sum(r, "major_cov", &top, &bot);
    if (top < o.consensus_minimum_variant_coverage_each_strand || bot < o.consensus_minimum_variant_coverage_each_strand)
      add_reject_reason(r, "VARIANT_STRAND_COVERAGE");
  }
  if (o.consensus_minimum_total_coverage_each_strand > 0) {
    strand_sum(r, "total_cov", &top, &bot);
    if (top < o.consensus_minimum_total_coverage_each_strand || bot < o.consensus_minimum_total_coverage_each_strand)
      add_reject_reason(r, "TOTAL_STRAND_COVERAGE");
  }
  if (o.consensus_minimum_variant_coverage > 0 && strand_sum(r, "major_cov") < o.consensus_minimum_variant_coverage)
    add_reject_reason(r, "VARIANT_COVERAGE");
  if (o.consensus_minimum_total_coverage > 0 && strand_sum(r, "total_cov") < o.consensus_minimum_total_coverage)
    add_reject_reason(r, "TOTAL_COVERAGE");
}

// identify_mutations.cpp:369-485: an indel that lengthens or shortens a run of its own base; a substitution that joins two runs
void indel_homopolymer(Row& r, const Reference& R, uint32_t indel_length, uint32_t surrounding_length, bool no_indel_polymorphisms) {
  const bool is_indel = r["ref_base"] == "." || r["new_base"] == ".";
  if (indel_length && is_indel) {
    const std::string& s = R.seq(r["seq_id"]);
    int32_t mut_pos = (int32_t)number<uint32_t>(r["position"]);
    const bool is_insertion = number<int32_t>(r["insert_position"]) > 0;
    const std::string mut_base = r["ref_base"] == "." ? r["new_base"] : r["ref_base"];
    int32_t run = 0;
    bool no_match = false;
    if (is_insertion && R.base(s, mut_pos) != mut_base) {   // not the base before: the base after, then
      ++mut_pos;
      if (mut_pos > (int32_t)s.size() || R.base(s, mut_pos) != mut_base) no_match = true;
    }
    if (!no_match) {
      int32_t first = mut_pos - 1;
      while (R.base(s, first) == mut_base && first > 0) --first;
      ++first;
      int32_t last = mut_pos + 1;
      while (R.base(s, last) == mut_base && last <= (int32_t)s.size()) ++last;
      --last;
      run = last - first + 1;
    }
    if (run >= (int32_t)indel_length) add_reject_reason(r, "INDEL_HOMOPOLYMER");
  }
  if (surrounding_length && r["ref_base"] != "." && r["new_base"] != ".") {
    const std::string& s = R.seq(r["seq_id"]);
    const int32_t mut_pos = number<int32_t>(r["position"]);
    const std::string mut_base = r["new_base"];
    int32_t first = mut_pos - 1;
    while (first >= 1 && R.base(s, first) == mut_base) --first;
    ++first;
    int32_t last = mut_pos + 1;
    while (last <= (int32_t)s.size() && R.base(s, last) == mut_base) ++last;
    --last;
    if (first < mut_pos && last > mut_pos && last - first + 1 >= (int32_t)surrounding_length) add_reject_reason(r, "SURROUNDING_HOMOPOLYMER");
  }
  if (no_indel_polymorphisms && is_indel) add_reject_reason(r, "POLYMORPHIC_INDEL");
}

double ra_score(const Row& r) {   // identify_mutations.cpp:180-192: evidence from before the two scores were merged
  if (has(r, "score")) return number_or_na(r.at("score"));
  double legacy = std::numeric_limits<double>::quiet_NaN();
  if (has(r, "consensus_score")) legacy = number_or_na(r.at("consensus_score"));
  if (has(r, "polymorphism_score")) {
    const double p = number_or_na(r.at("polymorphism_score"));
    if (std::isnan(legacy) || p > legacy) legacy = p;
  }
  return legacy;
}

// One row through the two questions (identify_mutations.cpp:522-685).  The modes differ in the bound the consensus question
// reads -- the lower one ("confidently the majority") in consensus mode, the upper one ("cannot rule out fixed") in polymorphism
// mode -- and in what happens to a row that answers neither.  Returns whether the row is deleted.
bool test_row(Row& r, const Reference& R, const RaFilterOptions& o, RaFilterCounts& n) {
  const double score = ra_score(r);
  double lower, upper;
  if (has(r, "frequency_lower") && has(r, "frequency_upper")) {
    lower = number<double>(r["frequency_lower"]);
    upper = number<double>(r["frequency_upper"]);
  } else {
    // evidence written before the bounds were recorded: Clopper-Pearson bounds with the raw read count as n
    // (identify_mutations.cpp:215-228); without usable depth, the point estimate
    lower = upper = number<double>(r["frequency"]);
    std::vector<std::string> tb = has(r, "total_cov") ? split(r["total_cov"], '/') : std::vector<std::string>();
    if (tb.size() >= 2) {
      const double depth = number<double>(tb[0]) + number<double>(tb[1]);
      if (depth > 0.0) {
        const double k = number<double>(r["frequency"]) * depth;
        lower = binomial_frequency_lower_bound(k, depth);
        upper = binomial_frequency_upper_bound(k, depth);
      }
    }
  }

  if (score < o.mutation_log10_e_value_cutoff) add_reject_reason(r, "SCORE_CUTOFF");
  if (o.consensus_frequency_cutoff > 0.0 && (o.polymorphism_prediction ? upper : lower) < o.consensus_frequency_cutoff)
    add_reject_reason(r, "FREQUENCY_CUTOFF");
  consensus_coverage(r, o);
  indel_homopolymer(r, R, o.consensus_reject_indel_homopolymer_length, o.consensus_reject_surrounding_homopolymer_length, false);
  if (!has(r, "reject")) {
    r["prediction"] = "consensus";
    ++n.consensus;
    return r["ref_base"] == r["major_base"];   // nothing but the reference base
  }
  r["consensus_reject"] = r["reject"];
  r.erase("reject");

  if (score < o.polymorphism_log10_e_value_cutoff) add_reject_reason(r, "SCORE_CUTOFF");
  if (o.polymorphism_frequency_cutoff > 0.0 && lower < o.polymorphism_frequency_cutoff) add_reject_reason(r, "FREQUENCY_CUTOFF");
  polymorphism_bias(r, o);
  polymorphism_coverage(r, o);
  indel_homopolymer(r, R, o.polymorphism_reject_indel_homopolymer_length, o.polymorphism_reject_surrounding_homopolymer_length,
                    o.polymorphism_no_indels);
  if (!has(r, "reject")) {
    r["prediction"] = "polymorphism";
    ++n.polymorphism;
    return false;
  }
  if (!o.polymorphism_prediction) {   // consensus mode drops the row
    r["polymorphism_reject"] = r["reject"];
    r.erase("reject");
    return true;
  }
  r["prediction"] = "polymorphism";   // polymorphism mode keeps it, marked rejected
  ++n.rejected_kept;
  return false;
}

}  // namespace

RaFilterOptions ra_filter_defaults(bool polymorphism_prediction) {
  RaFilterOptions o;   // the consensus-mode values are the member initialisers
  o.polymorphism_prediction = polymorphism_prediction;
  if (polymorphism_prediction) {
    o.consensus_frequency_cutoff = 0.95;
    o.polymorphism_log10_e_value_cutoff = 2;
    o.polymorphism_frequency_cutoff = 0.05;
  }
  return o;
}

void normalise_reference(RefSet& ref) {
  for (std::string& s : ref.seqs)
    for (char& c : s) {
      c = (char)toupper((unsigned char)c);
      if (!strchr("ATCGN", c) || c == 0) c = 'N';
    }
}

namespace {

void parse_ra(const std::vector<std::string>& col, const std::string& line, Row& r) {
  if (col.size() < 8) throw std::runtime_error("evidence row with fewer than eight columns: " + line);
  for (int i = 0; i < 5; ++i) r[RA_COLUMNS[i]] = col[3 + i];
  for (size_t i = 8; i < col.size(); ++i) {
    const size_t eq = col[i].find('=');
    if (eq == std::string::npos || eq == 0 || eq + 1 == col[i].size()) continue;   // cKeyValuePair::valid, common.h:1284
    r[col[i].substr(0, eq)] = col[i].substr(eq + 1);
  }
}

// cDiffEntry::marshal (genome_diff_entry.cpp:1323-1369): type, id, evidence, the type's columns, then what is left in key order
std::string marshal(const std::string& type, const std::string& id, const std::string& evidence, Row r, const char* const* columns, int n_columns) {
  std::string out = type + '\t' + id + '\t' + (evidence.empty() ? "." : evidence);
  for (int i = 0; i < n_columns; ++i) {
    auto it = r.find(columns[i]);
    if (it == r.end()) throw std::runtime_error("Did not find required field '" + std::string(columns[i]) + "' to write in entry id " + id + " of type '" + type + "'.");
    out += '\t';
    out += it->second;
    r.erase(it);
  }
  for (const auto& kv : r) {
    if (kv.first[0] == '_' || kv.second.empty()) continue;
    out += '\t' + kv.first + '=' + kv.second;
  }
  return out;
}

}  // namespace

RaFilterCounts test_ra_evidence(const std::string& gd_in, const RefSet& ref, const RaFilterOptions& opt, const std::string& gd_out) {
  std::ifstream in(gd_in);
  if (!in) throw std::runtime_error("cannot open " + gd_in);
  const Reference R{ref};
  RaFilterCounts n;
  std::string out, line;
  while (std::getline(in, line)) {
    if (line.compare(0, 3, "RA\t") != 0) { out += line; out += '\n'; continue; }
    std::vector<std::string> col = split(line, '\t');
    Row r;
    parse_ra(col, line, r);
    if (!has(r, "score") && !has(r, "consensus_score") && !has(r, "polymorphism_score"))
      throw std::runtime_error("Expected field 'score' in evidence item\n" + line);
    if (!has(r, "frequency")) throw std::runtime_error("Expected field 'frequency' in evidence item\n" + line);
    r["_id"] = col[1];   // for messages; keys with a leading underscore are never written
    ++n.rows;
    bool gone = test_row(r, R, opt, n);
    if (has(r, "user_defined")) gone = false;   // user evidence is classified, never dropped
    if (gone) { ++n.deleted; continue; }
    out += marshal(col[0], col[1], col[2] == "." ? "" : col[2], r, RA_COLUMNS, 5);
    out += '\n';
  }
  std::ofstream f(gd_out, std::ios::binary);
  if (!f) throw std::runtime_error("cannot create " + gd_out);
  f << out;
  if (!f.flush()) throw std::runtime_error("cannot write " + gd_out);
  return n;
}

// ---------------------------------------------------------------------------------------------------------------------------
// RA rows -> SNP / DEL / INS / SUB

namespace {

struct Evidence {           // one line of the input
  std::string type, id, line;
  Row row;                  // RA rows only
  int64_t mc_start = 0, mc_end = 0;
  std::string mc_seq;
  bool changed = false;
};

struct Mutation {
  std::string type, id;
  std::vector<std::string> evidence;
  Row row;
};

int32_t to_int(const std::string& s) { return number<int32_t>(s); }

int type_order(const std::string& t) {   // sort_order, genome_diff_entry.cpp:310-314
  return t == "DEL" ? 1 : t == "SNP" ? 2 : t == "INS" ? 3 : 4;
}

// cDiffEntry::compare (genome_diff_entry.cpp:566-694) between two mutations, then cGenomeDiff::diff_entry_ptr_sort's id rule
bool mutation_before(const Mutation& a, const Mutation& b) {
  const std::string &sa = a.row.at("seq_id"), &sb = b.row.at("seq_id");
  if (sa != sb) return sa < sb;
  const uint32_t pa = number<uint32_t>(a.row.at("position")), pb = number<uint32_t>(b.row.at("position"));
  if (pa != pb) return pa < pb;
  if (type_order(a.type) != type_order(b.type)) return type_order(a.type) < type_order(b.type);
  // same type from here on: the fields of its line specification (INS: the extended one with insert_position)
  static const char* const by_type[4][3] = {{"size", nullptr, nullptr}, {"new_seq", nullptr, nullptr}, {"insert_position", "new_seq", nullptr}, {"size", "new_seq", nullptr}};
  for (const char* const* f = by_type[type_order(a.type) - 1]; *f; ++f) {
    const bool ea = a.row.count(*f) > 0, eb = b.row.count(*f) > 0;
    if (!ea && !eb) continue;
    if (!eb) return false;
    if (!ea) return true;
    if (!strcmp(*f, "new_seq")) {
      if (a.row.at(*f) != b.row.at(*f)) return a.row.at(*f) < b.row.at(*f);
    } else {
      const int32_t va = to_int(a.row.at(*f)), vb = to_int(b.row.at(*f));
      if (va != vb) return va < vb;
    }
  }
  const uint32_t ia = number<uint32_t>(a.id), ib = number<uint32_t>(b.id);
  if (ia != ib) return ia < ib;
  return a.id < b.id;
}

}  // namespace

RaMutationCounts predict_ra_mutations(const std::string& gd_in, const RefSet& ref, bool polymorphism_prediction, bool targeted_sequencing,
                                      bool call_mutations_overlapping_missing_coverage, const std::string& gd_out) {
  std::ifstream in(gd_in);
  if (!in) throw std::runtime_error("cannot open " + gd_in);
  const Reference R{ref};
  RaMutationCounts n;
  std::vector<std::string> header;
  std::vector<Evidence> rows;
  std::map<std::string, bool> id_used;
  std::string line;
  while (std::getline(in, line)) {
    if (line.empty()) continue;
    if (line[0] == '#') { header.push_back(line); continue; }
    std::vector<std::string> col = split(line, '\t');
    if (col.size() < 3) throw std::runtime_error("not a GenomeDiff row: " + line);
    Evidence e;
    e.type = col[0];
    e.id = col[1];
    e.line = line;
    for (const char* m : {"SNP", "SUB", "DEL", "INS", "MOB", "AMP", "INV", "CON", "INT"})
      if (e.type == m) throw std::runtime_error("an evidence file is expected: " + gd_in + " already holds mutation rows (" + line.substr(0, 40) + " ...)");
    if (e.type == "RA") parse_ra(col, line, e.row);
    if (e.type == "MC") {
      if (col.size() < 6) throw std::runtime_error("MC row with fewer than six columns: " + line);
      e.mc_seq = col[3];
      e.mc_start = number<uint32_t>(col[4]);
      e.mc_end = number<uint32_t>(col[5]);
    }
    id_used[e.id] = true;
    rows.push_back(std::move(e));
  }

  // RA rows that lie in missing coverage are spurious reads inside a deletion (mutation_predictor.cpp:1969-2008); an inserted
  // column (position, k > 0) lies behind its position, so the last position of the MC does not hold it (cReferenceCoordinate)
  std::vector<Evidence*> ra;
  for (Evidence& e : rows) if (e.type == "RA") ra.push_back(&e);
  if (!targeted_sequencing && !call_mutations_overlapping_missing_coverage) {
    for (Evidence* e : ra) {
      if (has(e->row, "user_defined")) continue;
      const int64_t p = number<uint32_t>(e->row["position"]), k = number<uint32_t>(e->row["insert_position"]);
      for (const Evidence& mc : rows) {
        if (mc.type != "MC" || mc.mc_seq != e->row["seq_id"]) continue;
        const bool from = p > mc.mc_start || (p == mc.mc_start && k >= 0);
        const bool to = p < mc.mc_end || (p == mc.mc_end && k <= 0);
        if (from && to) { e->row["deleted"] = "1"; e->changed = true; ++n.ra_marked_deleted; break; }
      }
    }
  }

  // MutationPredictor::sort_by_pos (:101-108), a stable sort
  std::stable_sort(ra.begin(), ra.end(), [](Evidence* a, Evidence* b) {
    if (a->row["seq_id"] != b->row["seq_id"]) return a->row["seq_id"] < b->row["seq_id"];
    if (a->row["position"] != b->row["position"]) return to_int(a->row["position"]) < to_int(b->row["position"]);
    return to_int(a->row["insert_position"]) < to_int(b->row["insert_position"]);
  });

  // neighbours become one mutation -- unless they are polymorphisms (:2014-2109)
  std::vector<Mutation> groups;
  bool first_time = true;
  Mutation mut;
  for (Evidence* e : ra) {
    Row& item = e->row;
    const std::string& seq = R.seq(item["seq_id"]);
    const int32_t position = to_int(item["position"]);
    const std::string ref_base = (has(item, "insert_position") && to_int(item["insert_position"]) != 0) ? "." : R.base(seq, position);
    const std::string new_base = item["major_base"] == ref_base ? item["minor_base"] : item["major_base"];
    const bool is_consensus = item["prediction"] == "consensus";
    if (!has(item, "user_defined") && (has(item, "reject") || has(item, "deleted"))) continue;
    if (!polymorphism_prediction && !is_consensus) continue;   // mixed calls stay unassigned evidence in consensus mode
    bool same = false;
    if (!first_time) {
      if ((mut.row["end"] == item["position"] && to_int(mut.row["insert_end"]) + 1 == to_int(item["insert_position"])) ||
          (to_int(mut.row["end"]) + 1 == to_int(item["position"]) && item["insert_position"] == "0"))
        same = true;
      if (polymorphism_prediction && (!is_consensus || mut.row["frequency"] != "1" || mut.row["seq_id"] != item["seq_id"])) same = false;
    }
    if (!same) {
      if (!first_time) groups.push_back(mut);
      first_time = false;
      Mutation m;
      m.evidence.push_back(e->id);
      m.row["seq_id"] = item["seq_id"];
      m.row["position"] = m.row["start"] = m.row["end"] = item["position"];
      m.row["insert_start"] = m.row["insert_end"] = item["insert_position"];
      m.row["ref_seq"] = ref_base != "." ? ref_base : "";
      m.row["new_seq"] = new_base != "." ? new_base : "";
      if (polymorphism_prediction)   // cDiffEntry::mutation_frequency (genome_diff_entry.h:344-349): a consensus call counts as 1
        m.row["frequency"] = has(item, "frequency") ? (is_consensus ? "1" : item["frequency"]) : "1";
      mut = m;
    } else {
      mut.row["insert_end"] = item["insert_position"];
      mut.row["end"] = item["position"];
      mut.row["ref_seq"] += ref_base != "." ? ref_base : "";
      mut.row["new_seq"] += new_base != "." ? new_base : "";
      mut.evidence.push_back(e->id);
    }
  }
  if (!first_time) groups.push_back(mut);

  // the fields of each mutation type (:2115-2208)
  std::vector<Mutation> made;
  uint32_t id_counter = 0;
  for (Mutation m : groups) {
    Row& r = m.row;
    if (r["ref_seq"].empty()) {
      m.type = "INS";
      r.erase("ref_seq");
      if (polymorphism_prediction) {
        if (r["frequency"] != "1" && r["insert_start"] != r["insert_end"]) throw std::runtime_error("Polymorphism has incorrectly merged INS mutations.");
        r["insert_position"] = r["insert_start"];
      } else if (to_int(r["insert_start"]) != 1) {
        continue;   // inserted columns that do not begin behind the reference base are not called
      }
    } else if (r["new_seq"].empty()) {
      m.type = "DEL";
      r["size"] = std::to_string(to_int(r["end"]) - to_int(r["start"]) + 1);
      r.erase("new_seq");
      r.erase("ref_seq");
    } else if (r["ref_seq"].size() > 1 || r["new_seq"].size() > 1) {
      int32_t lowest = -1;   // the first of its rows that sits on a reference base
      for (Evidence* e : ra) {
        bool mine = false;
        for (const std::string& id : m.evidence) mine = mine || id == e->id;
        if (mine && (lowest < 0 || lowest > to_int(e->row["position"])) && e->row["ref_base"] != ".") lowest = to_int(e->row["position"]);
      }
      if (lowest > -1) r["position"] = std::to_string(lowest);
      m.type = "SUB";
      r["size"] = std::to_string(r["ref_seq"].size());
      r.erase("ref_seq");
    } else {
      m.type = "SNP";
      r.erase("ref_seq");
    }
    for (const char* k : {"start", "end", "insert_start", "insert_end"}) r.erase(k);
    uint32_t id = ++id_counter;   // cGenomeDiff::new_unique_id (genome_diff.cpp:768-777)
    while (id_used.count(std::to_string(id))) ++id;
    m.id = std::to_string(id);
    id_used[m.id] = true;
    (m.type == "SNP" ? n.snp : m.type == "DEL" ? n.del : m.type == "INS" ? n.ins : n.sub)++;
    made.push_back(std::move(m));
  }
  std::stable_sort(made.begin(), made.end(), mutation_before);

  std::string out;
  for (const std::string& h : header) { out += h; out += '\n'; }
  static const char* const SNP_COLUMNS[] = {"seq_id", "position", "new_seq"};
  static const char* const SUB_COLUMNS[] = {"seq_id", "position", "size", "new_seq"};
  static const char* const DEL_COLUMNS[] = {"seq_id", "position", "size"};
  for (const Mutation& m : made) {
    std::string evidence;
    for (size_t i = 0; i < m.evidence.size(); ++i) evidence += (i ? "," : "") + m.evidence[i];
    if (m.type == "SUB") out += marshal(m.type, m.id, evidence, m.row, SUB_COLUMNS, 4);
    else if (m.type == "DEL") out += marshal(m.type, m.id, evidence, m.row, DEL_COLUMNS, 3);
    else out += marshal(m.type, m.id, evidence, m.row, SNP_COLUMNS, 3);   // INS has SNP's columns
    out += '\n';
  }
  for (const Evidence& e : rows) {
    if (e.changed) {
      std::vector<std::string> col = split(e.line, '\t');
      out += marshal(e.type, e.id, col[2] == "." ? "" : col[2], e.row, RA_COLUMNS, 5);
    } else {
      out += e.line;
    }
    out += '\n';
  }
  std::ofstream f(gd_out, std::ios::binary);
  if (!f) throw std::runtime_error("cannot create " + gd_out);
  f << out;
  if (!f.flush()) throw std::runtime_error("cannot write " + gd_out);
  return n;
}

}  // namespace brq
