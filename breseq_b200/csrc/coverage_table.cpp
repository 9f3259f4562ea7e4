// BAM2COV's table writer (host): see coverage_table.h.
#include "coverage_table.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <stdexcept>

namespace brq {

namespace {
std::string number(double v) {   // an ostream's default formatting of a double
  char buf[64];
  snprintf(buf, sizeof buf, "%.6g", v);
  return buf;
}
}  // namespace

// BAM2COV's table (coverage_output::table, coverage_output.cpp:190-283): one row per handled position of `region`
// ("seq_id:start-end", commas in the numbers ignored, no end = one position; common.h:1666-1725), positions thinned to about
// `resolution` rows when it is not 0 (pileup_base.cpp:121-134: kept where (position + start) is a multiple of
// floor(size / resolution)), then the region's averages as commented lines.
void write_coverage_table(const std::string& path, const BamHeader& hdr, const RefSet& ref, const PileupStream& st,
                          const std::vector<CoverageColumn>& cols, const std::vector<std::vector<CoverageColumn>>& by_group,
                          const std::string& region, uint32_t resolution, bool total_only, bool csv, const double* reference_average) {
  const size_t colon = region.find(':');
  if (colon == std::string::npos || region.find(':', colon + 1) != std::string::npos)
    throw std::runtime_error("Expected exactly one colon in region string:" + region);
  const std::string name = region.substr(0, colon);
  std::string span = region.substr(colon + 1);
  for (size_t at; (at = span.find("\xe2\x80\x93")) != std::string::npos;) span.replace(at, 3, "-");   // an en dash reads as a hyphen
  span.erase(std::remove(span.begin(), span.end(), ','), span.end());
  const size_t dash = span.find('-');
  if (dash != std::string::npos && span.find('-', dash + 1) != std::string::npos)
    throw std::runtime_error("Expected no more than one hyphen in start-end portion of region string:" + region);
  auto position_of = [&](const std::string& t) {   // "pos" or "pos.insert": the table has no insert columns
    const std::string head = t.substr(0, t.find('.'));
    if (head.empty() || head.find_first_not_of("0123456789") != std::string::npos) throw std::runtime_error("bad position in region string:" + region);
    return (uint32_t)strtoul(head.c_str(), nullptr, 10);
  };
  const uint32_t start = position_of(span.substr(0, dash));
  const uint32_t end = dash == std::string::npos ? start : position_of(span.substr(dash + 1));
  size_t tid = 0;
  while (tid < hdr.target_names.size() && hdr.target_names[tid] != name) ++tid;
  if (tid == hdr.target_names.size()) throw std::runtime_error("Target seq id was not found for region [" + name + "]");
  if (start < 1 || end < start || end > hdr.target_lens[tid]) throw std::runtime_error("region outside its target: " + region);
  const Segment* seg = nullptr;
  for (const Segment& sg : st.segments) if ((size_t)sg.tid == tid && sg.lo <= (int32_t)start - 1 && sg.hi >= (int32_t)end) seg = &sg;
  if (!seg) throw std::runtime_error("region " + region + " is not inside the staged range of this context");
  size_t ri = 0;
  while (ri < ref.names.size() && ref.names[ri] != name) ++ri;

  uint32_t downsample = 1;
  if (resolution != 0) {
    downsample = (uint32_t)floor((double)(end - start + 1) / (double)resolution);
    if (downsample < 1) downsample = 1;
  }
  std::ofstream out(path.c_str());
  if (!out) throw std::runtime_error("cannot create " + path);
  const char* d = csv ? "," : "\t";
  out << "position" << d << "ref_base";
  const std::vector<std::string> names = total_only ? std::vector<std::string>{"unique_cov", "redundant_cov", "total_cov"}
                                                    : std::vector<std::string>{"unique_top_cov", "unique_bot_cov", "redundant_top_cov", "redundant_bot_cov",
                                                                               "raw_redundant_top_cov", "raw_redundant_bot_cov", "unique_top_begin", "unique_bot_begin"};
  for (const std::string& nm : names) out << d << nm;
  for (size_t g = 0; g < by_group.size(); ++g)   // the per-read-group repeats follow the aggregate columns (:236-244)
    for (const std::string& nm : names) out << d << "RG-" << g << "_" << nm;
  out << '\n';
  // Which positions get a row (pileup_base.cpp:141-200, 308-358 with coverage_output.cpp:318-330).  Up to the last column the
  // pileup engine reports (L: the last position of the region any read spans), the handled positions: inside the region and,
  // when thinning, with (position + start) a multiple of `downsample`.  Past L the reference's loop only advances its "last
  // position" at handled positions, and the callback fills everything since that last position with zero rows: every position
  // from L + 1 to the last handled one gets a row (from position 1 when no read touches the region at all).
  auto handled = [&](uint32_t pos) { return pos >= start && pos <= end && (pos + start) % downsample == 0; };
  auto column = [&](uint32_t pos) -> const CoverageColumn* {
    const int64_t c = (int64_t)pos - 1;
    return c >= seg->lo && c < seg->hi ? &cols[seg->slot0 + (uint64_t)(c - seg->lo)] : nullptr;
  };
  uint32_t L = 0;
  for (uint32_t pos = end; pos >= start; --pos) { if (column(pos)->covered) { L = pos; break; } }
  std::vector<uint32_t> rows;
  for (uint32_t pos = start; pos <= L; ++pos) if (handled(pos)) rows.push_back(pos);
  for (uint32_t pos = L + 1, last = L; pos <= end; ++pos) {
    if (!handled(pos)) continue;
    for (uint32_t i = last + 1; i <= pos; ++i) rows.push_back(i);
    last = pos;
  }
  const CoverageColumn none = {{0, 0}, {0, 0}, {0, 0}, 0, 0, {0.0, 0.0}};
  uint32_t n_positions = 0;
  struct Sum { double unique = 0, repeat = 0, all = 0; };
  Sum total;
  std::vector<Sum> group_total(by_group.size());
  auto cells = [&](const CoverageColumn& c, Sum& sum) {
    sum.unique += c.unique[0] + c.unique[1];
    sum.repeat += c.redundant[0] + c.redundant[1];
    sum.all += c.unique[0] + c.unique[1] + c.redundant[0] + c.redundant[1];
    if (total_only)
      out << d << (c.unique[0] + c.unique[1]) << d << number(c.redundant[0] + c.redundant[1]) << d
          << number(c.unique[0] + c.unique[1] + c.redundant[0] + c.redundant[1]);
    else
      out << d << c.unique[0] << d << c.unique[1] << d << number(c.redundant[0]) << d << number(c.redundant[1]) << d
          << c.raw_redundant[0] << d << c.raw_redundant[1] << d << c.begin[0] << d << c.begin[1];
  };
  for (uint32_t pos : rows) {
    const bool counted = pos <= L && column(pos) != nullptr;   // (past L nothing is covered; before the region's start nothing is counted)
    const uint64_t at = counted ? seg->slot0 + (uint64_t)((int64_t)pos - 1 - seg->lo) : 0;
    const char rc = ri < ref.seqs.size() ? ref.seqs[ri][(size_t)pos - 1] : 'N';
    ++n_positions;
    out << pos << d << rc;
    cells(counted ? cols[at] : none, total);
    for (size_t g = 0; g < by_group.size(); ++g) cells(counted ? by_group[g][at] : none, group_total[g]);
    out << '\n';
  }
  const double sum_unique = total.unique, sum_repeat = total.repeat, sum_all = total.all;
  if (reference_average) out << "#" << d << "reference_unique_average_cov" << d << number(*reference_average) << '\n';
  out << "#" << d << "region_unique_average_cov" << d << number(sum_unique / n_positions) << '\n';
  out << "#" << d << "region_repeat_average_cov" << d << number(sum_repeat / n_positions) << '\n';
  out << "#" << d << "region_average_cov" << d << number(sum_all / n_positions) << '\n';
  out << "#" << d << "number_of_positions" << d << n_positions << '\n';
  for (size_t g = 0; g < by_group.size(); ++g) {   // the groups' averages over the same positions (:270-279)
    const std::string pre = "RG-" + std::to_string(g) + "_";
    out << "#" << d << pre << "region_unique_average_cov" << d << number(group_total[g].unique / n_positions) << '\n';
    out << "#" << d << pre << "region_repeat_average_cov" << d << number(group_total[g].repeat / n_positions) << '\n';
    out << "#" << d << pre << "region_average_cov" << d << number(group_total[g].all / n_positions) << '\n';
  }
}

}  // namespace brq
