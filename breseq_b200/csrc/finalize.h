// Host-side finalisation of the pileup: covariate layout, the error-table file formats, the
// per-class likelihood table, and RA / MC / UN evidence emission.
#pragma once
#include "bam_io.h"
#include "brq_types.h"
#include "expand_core.h"
#include "kernels.h"

#include <functional>
#include <map>
#include <string>
#include <vector>

namespace brq {

// Covariates in the reference's enum order (error_count.h:69).
enum { COV_READ_SET, COV_REF_BASE, COV_PREV_BASE, COV_OBS_BASE, COV_QUALITY, COV_READ_POS, COV_BASE_REPEAT, COV_COUNT };

struct CovSpec {
  bool used[COV_COUNT] = {false, false, false, false, false, false, false};
  bool clamp[COV_COUNT] = {false, false, false, false, false, false, false};
  uint32_t maxv[COV_COUNT] = {0, 0, 0, 0, 0, 0, 0};
  uint32_t offset[COV_COUNT] = {0, 0, 0, 0, 0, 0, 0};
  bool per_position = false;
  uint32_t n_bins = 0;
  std::string text() const;  // print_covariates(), error_count.cpp:601-621
};

CovSpec parse_covariates(const std::string& s);  // read_covariates(), error_count.cpp:522-594
CovLayout to_layout(const CovSpec& spec);

// Stream a double the way `ostream << double` does by default (6 significant digits).
std::string format_default(double v);
// common.h:845-867 (fixed / scientific with a given precision, "NA" for NaN)
std::string format_double(double v, uint32_t precision, bool scientific);

void write_error_rates(const std::string& path, const CovSpec& spec, const std::vector<double>& log10_prob);
void read_error_rates(const std::string& path, CovSpec& spec, std::vector<double>& log10_prob);
// log10 table -> the values pass 2 actually uses: text round trip, then pow(10, x)
void canonicalise_table(const std::vector<double>& log10_prob, std::vector<double>& log10_text, std::vector<double>& prob);

// A few parked threads for the host's short data-parallel jobs (table canonicalisation, 0.4 ms of libc calls on one
// thread): waking them costs microseconds, creating them would cost as much as the job.
class WorkerPool {
 public:
  explicit WorkerPool(size_t n_threads);
  ~WorkerPool();
  size_t size() const { return n_ + 1; }   // the caller works too
  // job(part, n_parts) for part in [0, n_parts), n_parts = size(); returns when all parts are done
  void run(const std::function<void(size_t, size_t)>& job);
 private:
  struct Impl;
  Impl* impl_;
  size_t n_;
};
void canonicalise_table(const std::vector<double>& log10_prob, std::vector<double>& log10_text, std::vector<double>& prob, WorkerPool* pool);
void write_base_qual_tables(const std::string& pattern, const CovSpec& spec, const std::vector<uint64_t>& counts,
                            const std::vector<std::string>& readfiles);
void write_count_table(const std::string& path, const CovSpec& spec, const std::vector<uint64_t>& counts);
// a covariate string that names ref_pos: every position's non-empty bins (error_count.cpp:193-198, 803-846), counted on the host
// from the positional histogram records of a stream whose arrays are on the host
void write_count_table_per_position(const std::string& path, const CovSpec& spec, const PileupStream& st);
void write_coverage_distributions(const std::string& dir, const std::vector<uint64_t>& cov, uint64_t stride, uint64_t n_groups);

// Sizes and index maps of the likelihood tables (see score_geometry in finalize.cpp, build_tables_kernel in tables.cu).
struct TableGeometry {
  std::vector<uint32_t> mapqs;   // MAPQ value of each slot
  uint32_t n_st = 0;             // read sets x 2 strands
  uint32_t off_set = 0, off_ref = 0, off_obs = 0, off_qual = 0, off_rpos = 0, off_rep = 0;  // covariate table strides (0 = unused)
  size_t n_lut = 0, n_hotR = 0, n_cold = 0, n_tally_cells = 0;
};
void score_geometry(const CovSpec& spec, const uint32_t mapq_seen[8], const ScoreGeometry& stream_geometry, ScoreParams& p,
                    TableGeometry& g);
// Host copy of the per-class terms (read_set, strand, MAPQ present, quality, obs), libm arithmetic as the reference;
// entries are computed on first use (index (((set*2 + top) * n_mapq + mapq slot) * Q + quality) * 5 + obs).
struct ClassLut {
  const CovSpec* spec = nullptr; const std::vector<double>* prob = nullptr; const ScoreParams* p = nullptr; const TableGeometry* g = nullptr;
  std::vector<ClassTerms> terms;
  std::vector<uint8_t> done;
  bool ready() const { return !terms.empty(); }
  void clear() { terms.clear(); done.clear(); }
  void reset(const CovSpec& spec, const std::vector<double>& prob, const ScoreParams& p, const TableGeometry& g);
  const ClassTerms* get(size_t index);
};

struct EvidenceParams {
  double mutation_cutoff, polymorphism_cutoff, precision_decimal;
  uint32_t precision_places;
  uint32_t base_quality_cutoff;
  double log10_ref_length;
  bool skip_missing_coverage_prediction;
  bool polymorphism_prediction = false;   // Settings::polymorphism_prediction: only words the `prediction` field of user-evidence rows
  std::vector<double> deletion_propagation_cutoff, deletion_seed_cutoff;  // by BAM tid
};
struct EvidenceCounts { uint64_t ra = 0, mc = 0, un = 0, rechecked = 0, overturned = 0; };

// One row of ra_mc_evidence.gd before ids are assigned.
struct GdRow {
  int type = 0;  // 0 RA, 1 MC, 2 UN
  uint64_t id = 0;
  std::string seq_id;
  uint64_t a = 0, b = 0, c = 0, d = 0;  // RA: position, insert ; MC: start, end, start_range, end_range ; UN: start, end
  std::string ref_base, new_base;
  std::map<std::string, std::string> kv;
};
// A column the interval walk looks at, in target coordinates, with the host's verdict on base_predicted folded into
// `packed` (WalkOut) and the RA rows to emit at it (its own, then its insert sub-columns').
struct EvidenceEvent { uint32_t tid = 0, pos1 = 0, unique = 0, packed = 0; std::vector<GdRow> rows; };
// What one context (one coordinate shard of a run) contributes to ra_mc_evidence.gd.  The MC / UN intervals cross shard
// boundaries, so a sharded run walks the shards' events together (walk_evidence): the events are a few thousand per shard.
struct EvidenceShard {
  struct Seg { int32_t tid, lo, hi; };
  std::vector<std::string> target_names;
  std::vector<uint32_t> target_lens;
  std::vector<Seg> segments;           // in visit order
  std::vector<EvidenceEvent> events;   // in visit order, ascending position inside a target
  uint64_t rechecked = 0, overturned = 0;
};
// The records of the flagged slots when the stream lives in HBM only (device staging): per entry i of `flagged` (in the
// order given), its device words in record order, its side-list entries and its reference base.
struct FlaggedRecords {
  const uint64_t* word_off; const uint64_t* side_off;   // [n + 1] each; side_off in entries
  const uint32_t* words; const uint32_t* side;
  const uint8_t* ref;
};
// fr = nullptr: the records are read from st's host arrays
EvidenceShard collect_evidence(const BamHeader& hdr, const PileupStream& st, const std::vector<WalkEvent>& events,
                               const std::vector<uint32_t>& flagged, const std::vector<ColumnOut>& flagged_cols,
                               const ScoreParams& sp, ClassLut& lut, const EvidenceParams& ep, const FlaggedRecords* fr = nullptr);
EvidenceCounts walk_evidence(const std::vector<const EvidenceShard*>& shards, const EvidenceParams& ep, const std::string& gd_path);
// flat byte form of a shard, for the trip between ranks
std::string serialize_shard(const EvidenceShard& sh);
EvidenceShard parse_shard(const void* data, size_t bytes);

// collect_evidence + walk_evidence for an unsharded run
EvidenceCounts write_evidence(const std::string& gd_path, const BamHeader& hdr, const PileupStream& st,
                              const std::vector<WalkEvent>& events, const std::vector<uint32_t>& flagged, const std::vector<ColumnOut>& flagged_cols,
                              const ScoreParams& sp, ClassLut& lut, const EvidenceParams& ep, const FlaggedRecords* fr = nullptr);

// Optional outputs of pass 2 (identify_mutations.cpp:1693-1733 and :2028-2052, 2173-2204); both read the full per-slot results.
void write_per_position_file(const std::string& path, const BamHeader& hdr, const PileupStream& st, const std::vector<ColumnOut>& cols,
                             uint32_t base_quality_cutoff, const std::vector<double>& deletion_propagation_cutoff);
// by_group: empty, or per read group the coverage walk's columns with pass 2's notion of coverage (expand_core.h: coverage_lane)
void write_coverage_tsv(const std::string& pattern, const BamHeader& hdr, const RefSet& ref, const PileupStream& st, const std::vector<ColumnOut>& cols,
                        const std::vector<std::vector<CoverageColumn>>& by_group);

// Two-sided Fisher exact test on a 2x2 table of strand counts (stats.cpp:2144-2171)
double fisher_strand_p_value(uint32_t minor_top, uint32_t minor_bottom, uint32_t major_top, uint32_t major_bottom);

}  // namespace brq
