// Host staging: ReadBatch -> pinned, columnar, reference-position-sorted PileupStream.
//
// Replaces, for this path, the htslib pileup engine plus the per-read accessor calls the
// reference makes once per (read, column): /root/reference/src/breseq/pileup_base.cpp:141-210,
// 308-385 (column construction incl. zero-depth columns), alignment.cpp:44-57, 104-288, 371-390,
// alignment.h:354-410, error_count.cpp:854-986 and 1049-1105 (which base / quality / neighbour a
// record contributes).  Everything that depends only on the read and the column is decided here,
// once, and packed; the kernels do the counting, table math and per-column statistics.
#pragma once
#include "bam_io.h"
#include "brq_types.h"

#include <functional>

namespace brq {

struct ReadFileSetInfo { std::string base_name; uint32_t n_files; };

struct StageConfig {
  std::vector<std::string> call_seq_ids;          // empty = every target; visited in alphabetical order
  std::vector<uint32_t> coverage_group_of_tid;    // empty = one group per target
  std::vector<ReadFileSetInfo> read_file_sets;    // empty = everything is read file 0
  bool use_base_repeat = false;
  bool use_read_pos = false;           // the covariate string names read_pos: histogram records carry it (8-byte records)
  uint32_t base_quality_cutoff = 3;    // Settings::base_quality_cutoff: decides which records score (settings.cpp:1335)
  int threads = 8;
  bool want_hist = true, want_score = true;
  // the stage 03 (preprocess) call of error_count (breseq_cmdline.cpp:1969): also count, per target, the position-strand
  // combinations with / without a read start inside the junction read-end bound (error_count.cpp:157-166, 191-194, 217-229)
  bool preprocess_stage = false;
  uint32_t unmatched_end_minimum_read_length = 50;   // settings.cpp:1309
  double unmatched_end_length_factor = 0.1;          // 1 - require_match_fraction (settings.cpp:1308)
  bool compact_score = true;           // also build the transfer form of score_rec (brq_types.h: score16 + score_exc)
  bool compact_hist = true;            // also build the 16-bit histogram stream the device reads (brq_types.h)
  uint32_t shard_rank = 0, shard_count = 1;         // contiguous reference-coordinate shard staged by this call
  // ... or its explicit bounds [shard_lo, shard_hi) in the concatenated visit-order columns (a caller that balances the
  // shards by record count, SURVEY.md 8e, computes them with shard_bounds_by_records)
  uint64_t shard_lo = 0, shard_hi = 0;
  bool shard_explicit = false;
  std::vector<UserRa> user_evidence;               // Settings::user_evidence_genome_diff_file_name, read and sorted (read_user_evidence_gd)
  std::vector<double> user_skip_cutoff;             // optional, by BAM tid: targets with a negative deletion propagation cutoff are not visited
  int staging_mode = 0;                             // 0 = on the device when the context has one, 1 = host, 2 = device
  // buffer allocator (pinned when a device is present); both must be set together
  void* (*alloc)(size_t bytes, bool* pinned) = nullptr;
  void (*release)(void* p, bool pinned) = nullptr;
};

// Throws std::runtime_error where the reference would ASSERT (unsorted input, quality out of the
// packable range, a deletion with no following read base, ...).
void stage(const BamHeader& hdr, const RefSet& ref, const ReadBatch& reads, const StageConfig& cfg, PileupStream& out);
void free_stream(PileupStream& s, const StageConfig& cfg);

// The host-side plan both staging paths share: the visited targets clipped to the shard, and the table geometry.
void plan_segments(const BamHeader& hdr, const RefSet& ref, const StageConfig& cfg, PileupStream& out, std::vector<const std::string*>& refseq);
ScoreGeometry choose_geometry(const uint64_t* mq, const uint64_t* qc, const StageConfig& cfg, uint32_t max_read_set_seen);

// RA rows of a GenomeDiff file, stripped to their specification and sorted like cGenomeDiff::sort() (genome_diff_entry.cpp:566-690).
std::vector<UserRa> read_user_evidence_gd(const std::string& path);
// Walks the visited targets in visit order the way the pileup meets the user list (identify_mutations.cpp:1346-1355, 1914-2019):
// at a column whose position is that of the list's front entry, the insert levels up to the LAST front-run entry's are
// forced (the run is matched by position alone, as there), and the front entries naming this target, position and a
// processed level are consumed in order.  levels(visit index, pos1, force_max) = the highest insert level the column ends up
// with (read support and force together).  Entries the walk never reaches stay at the front and block the rest, as there.
std::vector<UserColumn> plan_user_evidence(const std::vector<UserRa>& list, const BamHeader& hdr, const std::vector<Segment>& visit_full,
                                           const std::vector<double>& skip_cutoff,
                                           const std::function<uint32_t(size_t, uint32_t, uint32_t)>& levels);
// highest insert level of a column: level k + 1 exists iff bit k of `mask` is set (a unique read has a longer insertion and a
// base there) or the user list forces it (k < force_max)
inline uint32_t insert_levels(uint64_t mask, uint32_t force_max) {
  uint32_t L = 0;
  while (L < 63 && ((mask >> L & 1) || L < force_max)) ++L;
  return L;
}

// Flat read-file index of each read group (alignment.cpp:565-605).
void make_read_file_partition(const ReadGroups& rg, const std::vector<ReadFileSetInfo>& sets,
                              std::vector<uint32_t>& base, std::vector<uint32_t>& count);

}  // namespace brq
