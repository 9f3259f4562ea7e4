// TEST INFRASTRUCTURE -- CPU oracle for the read-alignment evidence pileup.
//
// A plain, single-threaded restatement of the reference's algorithm for the hot path
// (breseq 0.50.0, /root/reference/src/breseq), driven through the htslib-compatible shim in
// hts_shim/.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build, link or execute anything in this directory; the product (breseq_b200/) never does.
//
// PARITY STATUS: pinned above the htslib boundary, unpinned below it.
//  * Above: oracle/ref_build.sh compiles the reference's OWN, unmodified libbreseq sources against
//    hts_shim/ into oracle/_ref/ref_cli (driver: ref_driver.cpp).  tests/golden/ holds what that binary
//    wrote for the test datasets (error_rates.tab, base_qual_error_prob.*.tab, coverage distributions,
//    ra_mc_evidence.gd, the per-position debug file); tests/test_golden.py requires this restatement
//    to reproduce every one of those files byte for byte and to agree per column.
//  * Below: htslib itself is not in this image, the reference ships no BAM-level fixture (its one
//    BAM is a missing blob) and its end-to-end goldens need bowtie2.  BAM
//    decode and pileup-column construction (bam_plp_auto semantics) in hts_shim/ are a restatement
//    of htslib 1.23.1 behaviour shared by the reference build and this oracle: "parity unpinned" there.
#pragma once
#include <cstdint>
#include <map>
#include <set>
#include <string>
#include <vector>

namespace oracle {

struct ReadFileSet { std::string base_name; uint32_t n_files; };

struct Settings {
  std::set<std::string> call_mutations_seq_ids;            // settings.h:993 (std::set => alphabetical visit order)
  std::map<std::string, uint32_t> seq_id_to_coverage_group; // settings.h:996
  uint32_t n_coverage_groups = 0;
  std::vector<ReadFileSet> read_file_sets;                  // empty for the standalone ERROR_COUNT path
  uint32_t base_quality_cutoff = 3;                         // settings.cpp:1335
  bool skip_missing_coverage_prediction = false;            // settings.cpp:843
  bool polymorphism_prediction = false;                     // settings.cpp:857 (words the `prediction` field of user-evidence rows)
  std::string user_evidence_genome_diff_file_name;          // RA rows reported whatever the data says (identify_mutations.cpp:879)
  std::string error_rates_file_name;                        // 07_error_calibration/error_rates.tab
  std::string unique_only_coverage_distribution_file_name;  // contains '@' replaced by the group index
  std::string base_qual_error_prob_file_name;               // contains '#' replaced by the read file name
  uint64_t total_reference_sequence_length = 0;
  // stage 03 (preprocess) call of error_count: junction read-end bound of settings.h:345-354 (settings.cpp:1308-1309)
  uint32_t unmatched_end_minimum_read_length = 50;
  double unmatched_end_length_factor = 0.1;                 // 1 - require_match_fraction
};

// error_count.h:41-52
void error_count(const Settings& settings, const std::string& bam, const std::string& fasta,
                 const std::string& output_dir, const std::vector<std::string>& readfiles,
                 bool do_coverage, bool do_errors, const std::string& covariates,
                 const std::string& counts_dump_file /* "" = none: raw count table, idx-ordered text */,
                 bool preprocess_stage = false,
                 std::map<std::string, double>* no_pos_hash_per_position_pr = nullptr /* Summary::preprocess_error_count, per seq id */);

struct ColumnDump {  // one per (column, insert_count); written raw to --columns-out
  uint32_t tid, pos1, insert_count, n;
  double ll[5];
  double consensus_score, variant_score;
  double f[5];
  double log10_likelihood;
  double unique[2], redundant[2];  // [0] bottom strand, [1] top strand
  int32_t raw_redundant[2];
  int32_t total;
  uint8_t best, major, minor, variant, ref, base_predicted, unique_only, emitted;
  uint32_t iterations;
};

// identify_mutations.h:46-60
void identify_mutations(const Settings& settings, const std::string& bam, const std::string& fasta,
                        const std::string& gd_file, const std::vector<double>& deletion_propagation_cutoff,
                        const std::vector<double>& deletion_seed_cutoff, double mutation_cutoff,
                        double polymorphism_cutoff, double polymorphism_precision_decimal,
                        uint32_t polymorphism_precision_places, const std::string& columns_dump_file,
                        uint64_t* n_records_out);

}  // namespace oracle
