// The Output stage's filter over the RA rows of pass 2's evidence file: what breseq runs between reading ra_mc_evidence.gd and
// merging it into evidence.gd (breseq_cmdline.cpp:2609-2614 -> test_RA_evidence, identify_mutations.cpp:687-749).  Every RA row
// is asked two questions in turn -- is the variant the consensus, is it present at all -- each a score cutoff, a frequency
// cutoff on the recorded confidence bound, coverage minima, the strand / quality bias p-values and the homopolymer rules; rows
// that fail both are dropped (consensus mode) or kept with their reject reasons (polymorphism mode).  Host text work over a few
// dozen rows: no device involved.  SURVEY.md 8f-4.
#pragma once
#include <cstdint>
#include <string>

#include "brq_types.h"

namespace brq {

struct RaFilterOptions {   // the members of breseq::Settings the filter reads (settings.h:473-500)
  bool polymorphism_prediction = false;
  double mutation_log10_e_value_cutoff = 10;
  double consensus_frequency_cutoff = 0.50;
  uint32_t consensus_minimum_variant_coverage = 0, consensus_minimum_total_coverage = 0;
  uint32_t consensus_minimum_variant_coverage_each_strand = 0, consensus_minimum_total_coverage_each_strand = 0;
  uint32_t consensus_reject_indel_homopolymer_length = 0, consensus_reject_surrounding_homopolymer_length = 0;
  double polymorphism_log10_e_value_cutoff = 10;
  double polymorphism_frequency_cutoff = 0.10;
  uint32_t polymorphism_minimum_variant_coverage = 0, polymorphism_minimum_total_coverage = 0;
  uint32_t polymorphism_minimum_variant_coverage_each_strand = 2, polymorphism_minimum_total_coverage_each_strand = 0;
  uint32_t polymorphism_reject_indel_homopolymer_length = 0, polymorphism_reject_surrounding_homopolymer_length = 0;
  double polymorphism_fisher_strand_p_value_cutoff = 0.05;
  double polymorphism_ks_quality_p_value_cutoff = 0;
  bool polymorphism_no_indels = false;
};

// what settings.cpp:862-896 (polymorphism mode) and :918-948 (consensus mode) leave in those members without further options
RaFilterOptions ra_filter_defaults(bool polymorphism_prediction);

struct RaFilterCounts {
  uint32_t rows = 0, consensus = 0, polymorphism = 0, rejected_kept = 0, deleted = 0;
};

// gd_in -> gd_out: RA rows filtered and annotated (prediction=, consensus_reject=, polymorphism_reject=, reject=), every other
// line as it was.  `ref` with upper-case ACGTN sequences (reference_sequence.cpp:794-812), see normalise_reference().
RaFilterCounts test_ra_evidence(const std::string& gd_in, const RefSet& ref, const RaFilterOptions& opt, const std::string& gd_out);

void normalise_reference(RefSet& ref);

// The RA step of mutation prediction (MutationPredictor::predictRAtoSNPorDELorINSorSUB, mutation_predictor.cpp:1955-2211) on
// an evidence file test_ra_evidence() has been through: RA rows inside an MC row are marked deleted=1 (unless the run is
// targeted sequencing or calls mutations over missing coverage), the accepted rows -- consensus calls, in polymorphism mode also
// polymorphisms -- are walked in position order, neighbours joined, and every group becomes a SNP, DEL, INS or SUB row that
// names its RA rows as evidence.  gd_out: the '#' lines, the new mutation rows in the GenomeDiff order, then the evidence rows.
struct RaMutationCounts {
  uint32_t snp = 0, del = 0, ins = 0, sub = 0, ra_marked_deleted = 0;
};
RaMutationCounts predict_ra_mutations(const std::string& gd_in, const RefSet& ref, bool polymorphism_prediction, bool targeted_sequencing,
                                      bool call_mutations_overlapping_missing_coverage, const std::string& gd_out);

}  // namespace brq
