// Synthetic aligned-read generator: reads placed at known coordinates with known CIGARs (no
// aligner), a Phred/error model mirroring `breseq SIMULATE-READS`
// (/root/reference/src/breseq/breseq_cmdline.cpp:1160-1171), planted variants, and breseq's aux
// tags (AS, X1, RG, XL, XR: /root/reference/src/breseq/alignment.cpp:911-964).
// Counter-based: every read is a pure function of (seed, read set, fragment id), so generation is
// thread-parallel and reproducible.
#pragma once
#include "bam_io.h"
#include "brq_types.h"

namespace brq {

struct SynthReadSet {
  std::string name;
  bool paired = false;
  uint32_t read_len = 35;
  double coverage = 100.0;  // mean aligned depth contributed by this set
  double frag_mean = 400, frag_sd = 40;
};

struct SynthVariant {
  int32_t tid; int32_t pos0;
  uint8_t kind;      // 0 SNP, 1 deletion, 2 insertion (after pos0)
  uint8_t len;       // indel length 1..3
  uint8_t alt[3];    // SNP alt / inserted bases (indices 0..3)
  uint32_t freq_ppm; // carrier probability per fragment
};

struct SynthConfig {
  uint64_t seed = 1;
  std::vector<SynthReadSet> sets;
  uint32_t n_polymorphic = 60;        // planted at 5-50 %
  uint32_t n_fixed = 10;              // planted at 100 %
  uint32_t min_freq_ppm = 50000, max_freq_ppm = 500000;
  uint32_t n_gaps = 2;                // sample deletions: no read may overlap them
  uint32_t gap_min = 300, gap_max = 1500;
  double q_start = 38, q_end = 28, q_sd = 4;
  int q_min = 2, q_max = 41;
  uint32_t indel_error_ppm = 10;      // each of insertion and deletion
  uint32_t n_base_ppm = 1000;
  uint32_t softclip_ppm = 10000;
  uint32_t low_mapq_ppm = 30000;
  uint32_t redundant_ppm = 20000;
  uint32_t trim_ppm = 50000;
  int threads = 8;
  // only the fragments that can overlap columns [window_lo, window_hi) of the concatenated contigs (window_hi > window_lo):
  // a rank of a run sharded by reference range generates its own share; every read is the same as in the full run
  uint64_t window_lo = 0, window_hi = 0;
};

// Uniform random ACGT reference (GC ~ 50 %); contig i is named "<prefix><i+1>" (zero padded so
// alphabetical order == index order) unless there is one contig.
void synth_reference(uint64_t seed, const std::vector<uint32_t>& contig_lens, const std::string& prefix, RefSet& ref);

// Generate the reads of `cfg` against `ref`; fills reads (coordinate sorted), hdr and variants.
void synth_reads(const SynthConfig& cfg, const RefSet& ref, BamHeader& hdr, ReadBatch& reads,
                 std::vector<SynthVariant>& variants);

// Cut points of `n_shards` contiguous coordinate shards with about the same number of aligned bases (balance
// by record count, not by columns): bounds[0] = 0 .. bounds[n_shards] = total length, in concatenated-contig columns.
// A pure function of the configuration (fragment start positions only): every rank computes the same cuts.
std::vector<uint64_t> synth_shard_bounds(const SynthConfig& cfg, const RefSet& ref, uint32_t n_shards);

}  // namespace brq
