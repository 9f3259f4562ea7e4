// BAM / FASTA I/O of the host staging layer (replaces the htslib calls behind
// /root/reference/src/breseq/pileup_base.cpp:61-88, 308-359 for this path).
#pragma once
#include <functional>
#include <utility>
#include "brq_types.h"

namespace brq {

struct BamHeader {
  std::string text;
  std::vector<std::string> target_names;
  std::vector<uint32_t> target_lens;
  ReadGroups read_groups;  // @RG lines in header order (ID, LB)
};

// Inflate the BGZF members of `path` with `threads` workers and parse every record.
// Throws std::runtime_error on malformed input.
// A rank of a run sharded by reference range does not need the whole file: `span_of` (called once the header is parsed) returns,
// per BAM target, the columns [first, second) whose reads to keep (first >= second: none); records that do not overlap them are
// skipped, and on a coordinate-sorted file (@HD SO:coordinate) reading stops at the first record past the last span, so
// the members behind it are never inflated.
typedef std::function<std::vector<std::pair<int32_t, int32_t>>(const BamHeader&)> ReadSpans;
void read_bam(const std::string& path, BamHeader& hdr, ReadBatch& reads, int threads, const ReadSpans* span_of = nullptr);

// Serialise `reads` (already coordinate sorted) as BGZF-compressed BAM.
void write_bam(const std::string& path, const BamHeader& hdr, const ReadBatch& reads, int level = 1, int threads = 0);

void read_fasta(const std::string& path, RefSet& ref);
void write_fasta(const std::string& path, const RefSet& ref, bool with_fai = true);

}  // namespace brq
