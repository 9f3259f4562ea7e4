// The breseq-side binding of libbrq.so: the BODIES of breseq::error_count() and breseq::identify_mutations() with their
// exact signatures (/root/reference/src/breseq/error_count.h:41-52, identify_mutations.h:46-60), calling the C ABI of
// include/brq.h.  A breseq maintainer drops this file into src/breseq/, removes the two function bodies from error_count.cpp
// and identify_mutations.cpp (or keeps them as error_count_cpu / identify_mutations_cpu, the fallback for the options that
// hang other detectors on the same pileup), and links -lbrq.  Nothing else of breseq changes: breseq_cmdline.cpp:1015, 1969,
// 2293 and 2471 compile against the same declarations.
//
// oracle/ref_build.sh compiles exactly this file against the reference's unmodified headers and objects into
// oracle/_ref/ref_cli_brq (the reference's own two bodies renamed by -Derror_count=error_count_cpu
// -Didentify_mutations=identify_mutations_cpu at compile time, nothing copied); tests/test_gpu_adapter.py runs it with real
// Settings / Summary objects on every golden dataset and diffs the files.
#include "error_count.h"
#include "identify_mutations.h"
#include "reference_sequence.h"
#include "settings.h"
#include "summary.h"

#include "brq.h"

#include <cstdlib>
#include <set>
#include <string>
#include <vector>

using namespace std;

namespace breseq {

// the reference's own bodies, kept as the fallback (renamed at compile time; see the header comment)
void error_count_cpu(const Settings& settings, Summary& summary, const string& bam, const string& fasta, const string& output_dir,
                     const vector<string>& readfiles, bool do_coverage, bool do_errors, bool preprocess, uint8_t min_qual_score,
                     const string& covariates);
void identify_mutations_cpu(const Settings& settings, const Summary& summary, const string& bam, const string& fasta, const string& gd_file,
                            const cReferenceSequences& ref_seq_info, const vector<double>& deletion_propagation_cutoff,
                            const vector<double>& deletion_seed_cutoffs, double mutation_cutoff, double polymorphism_cutoff,
                            double polymorphism_precision_decimal, uint32_t polymorphism_precision_places, bool print_per_position_file);

namespace {

// ONE context per process: breseq calls error_count() (stage 07) and identify_mutations() (stage 08) on the same reference.bam
// one after the other; the second call finds the reads decoded and, with the same options, the streams staged in HBM.
brq_ctx* the_context() {
  static brq_ctx* ctx = nullptr;
  if (!ctx) {
    const char* dev = getenv("BRQ_DEVICE");
    brq_config cfg = {dev ? atoi(dev) : 0 /* CUDA device */, 0 /* host threads: all cores */};
    ctx = brq_create(&cfg);
  }
  return ctx;
}

void brq_check(brq_ctx* ctx, int rc) {   // the reference's failure mode: message, backtrace, exit(1) (common.h:111-183)
  if (rc) ERROR(string("libbrq: ") + brq_last_error(ctx));
}

// BAM targets in header order (pileup_base::target_name(tid)); the cutoff vectors and the coverage groups are indexed by tid
vector<string> bam_targets(const string& bam) {
  htsFile* f = hts_open(bam.c_str(), "r");   // as pileup_base.cpp:66-70 does
  ASSERT(f, "Could not open BAM file: " + bam);
  sam_hdr_t* h = sam_hdr_read(f);
  ASSERT(h, "Could not read the header of BAM file: " + bam);
  vector<string> names;
  for (int i = 0; i < h->n_targets; ++i) names.push_back(h->target_name[i]);
  sam_hdr_destroy(h);
  hts_close(f);
  return names;
}

struct StageArgs {   // keeps what brq_stage_options points to alive
  vector<string> seq_id_strings;
  vector<const char*> seq_ids;
  vector<brq_read_file_set> sets;
  vector<uint32_t> groups;
  brq_stage_options opt;
};

void fill_stage_options(StageArgs& a, const Settings& settings, const string& bam, const string& covariates, bool preprocess) {
  const set<string> ids = settings.call_mutations_seq_id_set();                  // settings.h:993
  a.seq_id_strings.assign(ids.begin(), ids.end());
  for (const string& s : a.seq_id_strings) a.seq_ids.push_back(s.c_str());
  for (const cReadFileSet& rfs : settings.read_file_sets)                        // one set per @RG; a paired set owns two read files
    a.sets.push_back(brq_read_file_set{rfs.m_base_name.c_str(), (uint32_t)rfs.m_files.size()});
  for (const string& name : bam_targets(bam))                                    // settings.h:996, by BAM tid
    a.groups.push_back(settings.refseq_settings.m_seq_id_to_coverage_group_map.count(name) ? settings.seq_id_to_coverage_group(name) : 0u);
  brq_stage_options& o = a.opt;
  o = brq_stage_options();
  o.seq_ids = a.seq_ids.data(); o.n_seq_ids = (uint32_t)a.seq_ids.size();
  o.read_file_sets = a.sets.empty() ? NULL : a.sets.data(); o.n_read_file_sets = (uint32_t)a.sets.size();
  o.coverage_group_of_tid = a.groups.data(); o.n_targets = (uint32_t)a.groups.size();
  o.use_base_repeat = covariates.find("base_repeat") != string::npos;
  o.use_read_pos = covariates.find("read_pos") != string::npos;
  o.shard_rank = 0; o.shard_count = 1;
  o.base_quality_cutoff = settings.base_quality_cutoff;                          // settings.cpp:1335
  o.preprocess_stage = preprocess;
  o.unmatched_end_minimum_read_length = settings.unmatched_end_minimum_read_length;
  o.require_match_fraction = settings.require_match_fraction;
  o.user_evidence_gd = settings.user_evidence_genome_diff_file_name.empty() ? NULL : settings.user_evidence_genome_diff_file_name.c_str();
}

}  // namespace

void error_count(const Settings& settings, Summary& summary, const string& bam, const string& fasta, const string& output_dir,
                 const vector<string>& readfiles, bool do_coverage, bool do_errors, bool preprocess, uint8_t min_qual_score,
                 const string& covariates)
{
  (void)min_qual_score;   // accepted and unused in the reference, too (error_count.h:295)
  brq_ctx* ctx = the_context();
  StageArgs a;
  fill_stage_options(a, settings, bam, covariates, preprocess);
  vector<const char*> rf;
  for (const string& s : readfiles) rf.push_back(s.c_str());
  brq_check(ctx, brq_run_error_count(ctx, bam.c_str(), fasta.c_str(), output_dir.c_str(), settings.error_rates_file_name.c_str(),
                                     rf.data(), (uint32_t)rf.size(), do_coverage, do_errors, covariates.c_str(), &a.opt));
  if (preprocess) {   // the stage 03 call (breseq_cmdline.cpp:1969): Summary::preprocess_error_count, error_count.cpp:217-229
    const uint64_t* starts; uint32_t n_targets;   // per BAM tid: position-strand combinations without / with a read start
    brq_check(ctx, brq_preprocess_read_starts(ctx, &starts, &n_targets));
    const vector<string> names = bam_targets(bam);
    const set<string> ids = settings.call_mutations_seq_id_set();
    for (uint32_t tid = 0; tid < n_targets && tid < names.size(); ++tid) {
      if (!ids.count(names[tid])) continue;
      const double total = (double)(starts[2 * tid] + starts[2 * tid + 1]);
      summary.preprocess_error_count[names[tid]].no_pos_hash_per_position_pr = total != 0 ? (double)starts[2 * tid] / total : 1.0;
    }
  }
}

void identify_mutations(const Settings& settings, const Summary& summary, const string& bam, const string& fasta, const string& gd_file,
                        const cReferenceSequences& ref_seq_info, const vector<double>& deletion_propagation_cutoff,
                        const vector<double>& deletion_seed_cutoffs, double mutation_cutoff, double polymorphism_cutoff,
                        double polymorphism_precision_decimal, uint32_t polymorphism_precision_places, bool print_per_position_file)
{
  // options that hang other detectors on the same pileup stay on the reference's own body
  if (settings.predict_soft_clipping || settings.predict_missing_pairs || settings.predict_pair_distance) {
    identify_mutations_cpu(settings, summary, bam, fasta, gd_file, ref_seq_info, deletion_propagation_cutoff, deletion_seed_cutoffs,
                           mutation_cutoff, polymorphism_cutoff, polymorphism_precision_decimal, polymorphism_precision_places,
                           print_per_position_file);
    return;
  }
  brq_ctx* ctx = the_context();
  StageArgs a;
  fill_stage_options(a, settings, bam, "", false);   // (the error table names its covariates: the library reads them from the file)
  brq_score_params p = {mutation_cutoff, polymorphism_cutoff, polymorphism_precision_decimal, polymorphism_precision_places,
                        settings.base_quality_cutoff,
                        summary.sequence_conversion.total_reference_sequence_length,   // identify_mutations.cpp:777
                        settings.polymorphism_prediction ? BRQ_SCORE_POLYMORPHISM_PREDICTION : 0u, 0};
  ASSERT(deletion_propagation_cutoff.size() == deletion_seed_cutoffs.size(), "cutoff tables of different sizes");
  brq_check(ctx, brq_run_identify_mutations(ctx, bam.c_str(), fasta.c_str(), settings.error_rates_file_name.c_str(), gd_file.c_str(),
                                            deletion_propagation_cutoff.data(), deletion_seed_cutoffs.data(),
                                            (uint32_t)deletion_propagation_cutoff.size(), &p, settings.skip_missing_coverage_prediction, &a.opt));
  if (print_per_position_file)   // identify_mutations.cpp:1693-1733
    brq_check(ctx, brq_write_per_position_file(ctx, settings.mutation_identification_per_position_file_name.c_str(),
                                               deletion_propagation_cutoff.data(), (uint32_t)deletion_propagation_cutoff.size()));
  if (settings.predict_copy_number)   // <seq>.coverage.tsv, identify_mutations.cpp:2028-2052, 2173-2204 ('@' = the seq id)
    brq_check(ctx, brq_write_coverage_tsv(ctx, settings.complete_coverage_text_file_name.c_str()));
}

}  // namespace breseq
