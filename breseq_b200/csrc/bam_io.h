// BAM / FASTA I/O of the host staging layer (replaces the htslib calls behind
// /root/reference/src/breseq/pileup_base.cpp:61-88, 308-359 for this path).
#pragma once
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
void read_bam(const std::string& path, BamHeader& hdr, ReadBatch& reads, int threads);

// Serialise `reads` (already coordinate sorted) as BGZF-compressed BAM.
void write_bam(const std::string& path, const BamHeader& hdr, const ReadBatch& reads, int level = 1, int threads = 0);

void read_fasta(const std::string& path, RefSet& ref);
void write_fasta(const std::string& path, const RefSet& ref, bool with_fai = true);

}  // namespace brq
