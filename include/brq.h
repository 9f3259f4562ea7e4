/* brq.h -- C ABI of the B200-native read-alignment evidence pileup (libbrq.so).
 *
 * Drop-in boundary for two breseq 0.50.0 entry points and nothing else:
 *   breseq::error_count()          /root/reference/src/breseq/error_count.h:41-52
 *   breseq::identify_mutations()   /root/reference/src/breseq/identify_mutations.h:46-60
 * The reference has no FFI of its own for this path; these are the entry
 * points a maintainer binds from error_count.cpp / identify_mutations.cpp.
 * Plain C: pointers and sizes only, ctx-owned result buffers, 0 on success and non-zero on
 * failure with the message in brq_last_error().  The library never falls back to a CPU
 * implementation of the kernels: without a usable CUDA device every compute call fails.
 */
#ifndef BRQ_H
#define BRQ_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct brq_ctx brq_ctx;

typedef struct brq_config {
  int32_t device;   /* CUDA ordinal; -1 = host-only context (staging / file formats, no kernels) */
  int32_t threads;  /* host staging threads, 0 = hardware concurrency */
} brq_config;

brq_ctx* brq_create(const brq_config* cfg);
void brq_destroy(brq_ctx* ctx);
const char* brq_last_error(const brq_ctx* ctx);
const char* brq_version(void);

/* ---- host staging: BAM -> pinned columnar stream (replaces pileup_base.cpp:61-88, 239-385 and
 *      the per-record accessor calls of alignment.h/.cpp on this path) ------------------------ */
typedef struct brq_read_file_set {  /* one cReadFileSet == one @RG; a paired set has 2 files */
  const char* base_name;
  uint32_t n_files;
} brq_read_file_set;

typedef struct brq_stage_options {
  const char* const* seq_ids;              /* Settings::call_mutations_seq_id_set(); NULL = all targets */
  uint32_t n_seq_ids;
  const brq_read_file_set* read_file_sets; /* Settings::read_file_sets; NULL = standalone ERROR_COUNT (read_set 0) */
  uint32_t n_read_file_sets;
  const uint32_t* coverage_group_of_tid;   /* Settings::seq_id_to_coverage_group per BAM tid; NULL = one per target */
  uint32_t n_targets;
  uint32_t use_base_repeat;                /* the covariate string names base_repeat */
  uint32_t use_read_pos;                   /* the covariate string names read_pos (either one: 8-byte histogram records) */
  uint32_t shard_rank, shard_count;        /* contiguous reference-coordinate shard of this process; 0,1 = all */
  uint32_t base_quality_cutoff;            /* Settings::base_quality_cutoff (decides which records score; 0 is a value:
                                              every quality scores); BRQ_DEFAULT_BASE_QUALITY_CUTOFF = the default, 3 */
  /* error_count(..., preprocess_stage = true), the stage 03 call (breseq_cmdline.cpp:1969, error_count.cpp:157-166, 191-194,
   * 217-229): also count the position-strand combinations with / without a read start inside the junction read-end bound */
  uint32_t preprocess_stage;
  uint32_t unmatched_end_minimum_read_length;  /* Settings::unmatched_end_minimum_read_length; 0 = the default, 50 */
  double require_match_fraction;               /* Settings::require_match_fraction; 0 = the default, 0.9 */
  /* explicit bounds of this process's shard in the concatenated visit-order columns (shard_hi > shard_lo): set by a caller
   * that balances its shards by record count (brq_shard_bounds); wins over shard_rank / shard_count */
  uint64_t shard_lo, shard_hi;
  /* where the streams are built: 0 = on the device when the context has one (csrc/expand.cu: the reads cross PCIe, kernels
   * expand the CIGARs and classify the records in HBM), 1 = on the host (csrc/staging.cpp), 2 = on the device or fail */
  uint32_t staging;
  uint32_t reserved;
  /* Settings::user_evidence_genome_diff_file_name (identify_mutations.cpp:879, 1013-1020): a GenomeDiff file whose RA rows are
   * reported whatever the data says; NULL = none.  Forces the insert sub-columns the rows name (:1346-1355); the rows come out
   * of brq_write_evidence with user_defined=1 (:1914-2019) */
  const char* user_evidence_gd;
} brq_stage_options;
#define BRQ_DEFAULT_BASE_QUALITY_CUTOFF 0xFFFFFFFFu

int brq_stage_bam(brq_ctx* ctx, const char* bam, const char* fasta, const brq_stage_options* opt);

/* Synthetic aligned reads (the reference's SIMULATE-READS error model, breseq_cmdline.cpp:1160-1171). */
typedef struct brq_synth_read_set {
  const char* name;
  uint32_t paired, read_len;
  double coverage, frag_mean, frag_sd;
} brq_synth_read_set;

typedef struct brq_synth_spec {
  uint64_t seed;
  const uint32_t* contig_lens;     /* random ACGT contigs ... */
  uint32_t n_contigs;
  const char* contig_prefix;
  const char* fasta;               /* ... or an existing FASTA (contig_lens == NULL) */
  const brq_synth_read_set* sets;
  uint32_t n_sets;
  uint32_t n_polymorphic, n_fixed, n_gaps;
  uint32_t min_freq_ppm, max_freq_ppm;
  /* only the reads that can overlap columns [window_lo, window_hi) of the concatenated contigs (window_hi > window_lo): a
   * rank of a run sharded by reference range generates its own share; each read is the same as in the full run */
  uint64_t window_lo, window_hi;
} brq_synth_spec;

int brq_synth_write(brq_ctx* ctx, const brq_synth_spec* spec, const char* bam_out, const char* fasta_out);
int brq_stage_synthetic(brq_ctx* ctx, const brq_synth_spec* spec, const brq_stage_options* opt);
/* cut points (n_shards + 1 values, concatenated visit-order columns) of shards with about the same number of aligned bases
 * (balance by record count); a pure function of the spec, so every rank computes the same cuts */
int brq_synth_shard_bounds(brq_ctx* ctx, const brq_synth_spec* spec, uint32_t n_shards, uint64_t* bounds);
/* the same for a BAM, from its records' start positions (one pass over the file's records) */
int brq_bam_shard_bounds(brq_ctx* ctx, const char* bam, uint32_t n_shards, uint64_t* bounds);

typedef struct brq_stream_info {
  uint64_t n_base, n_ins, n_score_records, n_hist_records, n_reads;
  uint64_t n_score_padded;         /* words in score_rec: round-major, lane-interleaved, padded (csrc/brq_types.h) */
  uint64_t bytes_host;             /* bytes of the staged stream (what brq_upload copies) */
  uint32_t n_targets, pinned;
  uint32_t device_built, hist_compact;  /* the stream was built in HBM (brq_stage_options.staging: the views below are copies made by this call); the device reads the compact histogram streams */
  uint32_t hist_record_bytes, side_stride;  /* 4, or 8 with read_pos / base_repeat / more than 16 read files; words per side-list entry (2 with read_pos / base_repeat) */
  uint64_t n_side;                 /* side-list entries (scoring records outside the shared table, X1 >= 511) */
  uint32_t base_quality_cutoff, hot_mapq, table_q_lo, table_n_q, table_n_st, table_words;  /* geometry baked into score_rec */
  const uint32_t* score_rec;       /* host views, valid until the next staging call; word layout: csrc/brq_types.h */
  const uint32_t* side_rec;
  const uint32_t* side_off;
  const uint64_t* score_off;       /* slot s: first word; record j at score_off[s] + (j>>3)*256 + ((j>>2)&1)*128 + (j&3) */
  const void* hist_rec;
  const uint64_t* hist_off;
  const uint8_t* slot_ref;
  const uint64_t* ins_parent;
  const uint32_t* ins_count;
  const uint32_t* round_slot;      /* [n_rounds * 32] tally rounds: 32 slots of one reference base and similar depth; 0xFFFFFFFF = idle lane */
  uint64_t n_rounds;
  const uint32_t* score_cnt;       /* [slots] records of every slot */
  const uint64_t* round_off;       /* [n_rounds + 1] first word of every round in score_rec */
  /* compact form of hist_rec, what brq_upload copies when present: the `fast` records in 16 bits each, the others unchanged (csrc/brq_types.h) */
  const uint16_t* hist16;
  const uint32_t* hist_exc;
  uint64_t n_hist16, n_hist_exc;
  /* transfer form of score_rec, what brq_upload copies when present: the low half of every word, bit 15 set where the word
   * does not follow from it, and those words in full per (round, lane) in record order (csrc/brq_types.h) */
  const uint16_t* score16;         /* [n_score_padded] */
  const uint32_t* score_exc;       /* [n_score_exc] */
  const uint32_t* score_exc_off;   /* [n_rounds * 32 + 1] */
  uint64_t n_score_exc;
} brq_stream_info;

int brq_stream(brq_ctx* ctx, brq_stream_info* info);
/* the counts and geometry of brq_stream_info without the array views (nothing is copied from HBM) */
int brq_stream_summary(brq_ctx* ctx, brq_stream_info* info);
/* device staging with the reads kept on the host (brq_stage_synthetic / brq_stage_bam leave them there): page-lock them once
 * (later brq_restage calls then copy at PCIe speed) and stage again from the host copy (H2D + expansion; what `e2e` times) */
int brq_pin_reads(brq_ctx* ctx);
int brq_restage(brq_ctx* ctx);
int brq_upload(brq_ctx* ctx);      /* host stream -> HBM (asynchronous on the ctx stream) */
int brq_sync(brq_ctx* ctx);

/* ---- pass 1: error_count (error_count.cpp:125-199, 854-1026) --------------------------------- */
int brq_error_count(brq_ctx* ctx, const char* covariates, int do_coverage, int do_errors);
/* device views for the one collective of the path (sum-allreduce of both integer histograms) */
/* preprocess stage (brq_stage_options.preprocess_stage): per BAM tid, the position-strand combinations of the staged range
 * without [2 tid] and with [2 tid + 1] a read start; Summary::preprocess_error_count[seq_id].no_pos_hash_per_position_pr is
 * without / (without + with), 1.0 when both are zero (error_count.cpp:217-229).  Shards add up. */
int brq_preprocess_read_starts(brq_ctx* ctx, const uint64_t** counts, uint32_t* n_targets);
/* A run sharded by reference range sizes the coverage histogram of every rank alike, so that the ranks can sum them: the
 * deepest unique column of this context's range, and a floor for the histogram's depth axis (the maximum over the ranks). */
int brq_max_coverage_depth(brq_ctx* ctx, uint64_t* depth);
int brq_set_min_coverage_depth(brq_ctx* ctx, uint64_t depth);
/* (the coverage histogram follows the counts in one allocation: coverage == counts + n_bins, so one collective over
 * n_bins + n_coverage words sums both) */
int brq_hist_device(brq_ctx* ctx, void** counts_u64, uint64_t* n_bins, void** coverage_u64, uint64_t* n_coverage);
int brq_hist_download(brq_ctx* ctx, const uint64_t** counts, uint64_t* n_bins, const uint64_t** coverage,
                      uint64_t* coverage_stride, uint64_t* n_groups);
/* counts -> log10 table on the device, canonicalised through the error_rates.tab text form
 * (error_count.cpp:660-690 writes 6 significant digits; :629-654 reads them back), then the
 * per-class likelihood table of pass 2 is built and uploaded. */
int brq_derive_error_table(brq_ctx* ctx);
int brq_error_table(brq_ctx* ctx, const double** log10_prob, uint64_t* n_bins);
int brq_write_error_count_files(brq_ctx* ctx, const char* output_dir, const char* error_rates_file,
                                const char* const* readfiles, uint32_t n_readfiles, int do_coverage, int do_errors,
                                const char* counts_dump_file /* NULL = none */);
int brq_load_error_table(brq_ctx* ctx, const char* error_rates_file);  /* read_log10_prob_table + log10_prob_to_prob */

/* ---- pass 2: identify_mutations (identify_mutations.cpp:1309-2022, 3240-3433) ---------------- */
typedef struct brq_score_params {
  double mutation_cutoff, polymorphism_cutoff, polymorphism_precision_decimal;
  uint32_t polymorphism_precision_places;
  uint32_t base_quality_cutoff;            /* Settings::base_quality_cutoff */
  uint64_t total_reference_length;         /* 0 = sum of BAM target lengths (identify_mutations.cpp:848-852) */
  uint32_t flags;                          /* BRQ_SCORE_* */
  uint32_t reserved;
} brq_score_params;

/* The reference fits the allele frequencies on every column (identify_mutations.cpp:1797) but only RA rows show the
 * result.  By default the kernel first bounds each column's presence score from above and runs the fit only where an RA
 * row is possible; this flag forces the fit on every column with scoring records (diagnostics, parity runs). */
#define BRQ_SCORE_FIT_ALL_COLUMNS 1u
/* Settings::polymorphism_prediction: words the `prediction` field of user-evidence rows (identify_mutations.cpp:1992-1996) */
#define BRQ_SCORE_POLYMORPHISM_PREDICTION 2u
/* diagnostics: a slot whose presence score the kernels only BOUNDED (the bound is under the cutoff: no fit) reports that upper
 * bound in brq_column.variant_score instead of NaN; bit 24 of `bits` still says that no fit ran (parity tests check
 * bound >= the reference's score on every column) */
#define BRQ_SCORE_KEEP_BOUNDS 4u

typedef struct brq_column {  /* one per slot: base columns of the visited targets, then insert sub-columns */
  double ll[5];
  double consensus_score, variant_score;
  double redundant[2];                     /* [0] bottom strand, [1] top strand */
  uint32_t unique[2], raw_redundant[2];
  uint32_t n, bits;                        /* bits: kernels.h ColumnOut; bit 24 = the EM fit was evaluated for this slot */
} brq_column;

int brq_score_columns(brq_ctx* ctx, const brq_score_params* p);
int brq_columns_download(brq_ctx* ctx, const brq_column** columns, uint64_t* n_slots, const uint32_t** flagged,
                         uint32_t* n_flagged);
int brq_columns_device(brq_ctx* ctx, void** columns, uint64_t* n_slots);
/* host finalisation: re-evaluates flagged slots in arrival order, emits RA rows with bounds and
 * bias statistics, runs the MC / UN interval state machines and writes ra_mc_evidence.gd */
int brq_write_evidence(brq_ctx* ctx, const char* gd_file, const double* deletion_propagation_cutoff,
                       const double* deletion_seed_cutoff, uint32_t n_targets, int skip_missing_coverage_prediction,
                       uint64_t* n_ra, uint64_t* n_mc, uint64_t* n_un);

/* Optional outputs of pass 2; both copy the full per-slot results to the host.
 * brq_write_per_position_file: the debug file identify_mutations() writes with print_per_position_file = true
 *   (identify_mutations.cpp:1693-1733, Settings::mutation_identification_per_position_file_name); targets whose
 *   deletion_propagation_cutoff is negative are skipped like the reference skips them.
 * brq_write_coverage_tsv: <seq>.coverage.tsv of --predict-copy-number (identify_mutations.cpp:2028-2052, 2173-2204,
 *   Settings::complete_coverage_text_file_name); '@' in `pattern` is replaced by the target name.  A BAM with two or
 *   more read groups gets the three columns once more per group ("RG-<n>_", :858-862): one walk per group over the reads the
 *   last staging left in HBM (needs device staging then). */
/* A run sharded by reference range (brq_stage_options.shard_rank / shard_count, one context per GPU): the MC and UN
 * intervals cross shard boundaries, so every context exports its share of the evidence (the event columns of its range
 * and its RA rows: a few hundred KB, valid until the next call on the context), the shares are gathered on one rank
 * (any transport) and brq_write_evidence_merged walks them together and writes ra_mc_evidence.gd.  `ctx` of the merged
 * call only carries the error message: it needs no device and no staged stream. */
int brq_evidence_export(brq_ctx* ctx, const double* deletion_propagation_cutoff, uint32_t n_targets, const void** data, uint64_t* bytes);
int brq_write_evidence_merged(brq_ctx* ctx, const void* const* shards, const uint64_t* sizes, uint32_t n_shards, const char* gd_file,
                              const double* deletion_propagation_cutoff, const double* deletion_seed_cutoff, uint32_t n_targets,
                              int skip_missing_coverage_prediction, uint64_t* n_ra, uint64_t* n_mc, uint64_t* n_un);
/* the CUDA stream (cudaStream_t) all of the context's device work is ordered on: a caller that enqueues its allreduce of
 * the brq_hist_device() buffers on it (or makes its own stream wait on it) needs no host synchronisation */
int brq_cuda_stream(brq_ctx* ctx, void** stream);
/* bytes the context has copied device -> host since the last reset (histograms, error table, walk events, flagged slots) */
int brq_d2h_bytes(brq_ctx* ctx, uint64_t* bytes, int reset);
int brq_write_per_position_file(brq_ctx* ctx, const char* path, const double* deletion_propagation_cutoff, uint32_t n_targets);
int brq_write_coverage_tsv(brq_ctx* ctx, const char* pattern);
/* The per-position count table of a covariate string that names ref_pos (error_count.cpp:105-111, 193-198, 803-846:
 * `error_counts.tab`, every position's non-empty covariate bins in bin order; a debugging output, gigabytes for a genome).
 * brq_run_error_count / brq_write_error_count_files write it instead of the error rates when the covariates say so; this call
 * writes it from any staged stream (the histogram records are counted on the host: no device needed after staging). */
int brq_write_per_position_counts(brq_ctx* ctx, const char* covariates, const char* path);
/* BAM2COV's table (`breseq BAM2COV -t`, coverage_output::table + pileup_callback, coverage_output.cpp:190-283, 307-470) for
 * region "seq_id:start-end" of the staged BAM: per position the unique coverage by strand (reads with an aligned base there:
 * a deletion over the position does not count), the redundant coverage by strand as the sum of 1 / X1 and as a count, and the
 * unique reads that begin there; or with total_only the three sums.  resolution = 0 writes every position, otherwise about
 * that many (the reference's thinning rule); csv = comma instead of tab.  The region's averages follow as '#' lines.  A walk
 * over the reads of the last staging on the device (a few ms): needs device staging.  per_read_group repeats the columns and
 * the averages once per @RG of the header (prefix "RG-<n>_", at least RG-0; one more walk per group).  Not written: the
 * read-begin and GC side files (their options are commented out in breseq's own command line, breseq_cmdline.cpp:332, 487-489). */
int brq_write_coverage_table(brq_ctx* ctx, const char* region, const char* path, uint32_t resolution, int total_only, int csv,
                             int per_read_group);
/* BAM2COV -a (--show-average): the same table with "# reference_unique_average_cov <value>" in front of the region's averages,
 * the value being Summary::references.reference[seq_id].coverage_average of breseq's summary.json (coverage_output.cpp:197-212,
 * 259-262) -- the caller reads it there, or takes brq_fit_coverage_distribution's nbinom mean. */
int brq_write_coverage_table_with_average(brq_ctx* ctx, const char* region, const char* path, uint32_t resolution, int total_only, int csv,
                                          int per_read_group, double reference_unique_average_cov);

/* ---- the collective of a sharded run, fused into pass 1 (csrc/exchange.cu) ------------------------------------------------
 * Instead of summing the brq_hist_device() buffers with a collective library between brq_error_count and
 * brq_derive_error_table, the ranks (one process per GPU of one NVLink / NVSwitch box, up to 16) can let pass 1 do it: every
 * context exports a handle of its inbox (64 bytes, a CUDA IPC memory handle), the handles are exchanged once by any means
 * (MPI, torch.distributed, a file) and attached in rank order; from then on every brq_error_count ends in a kernel that adds
 * the rank's histograms into its peers' inboxes over NVLink and its own inbox into its histograms, so that the histograms of
 * every rank hold the run's totals when the call's stream work is done.  All ranks have to call brq_error_count the same
 * number of times with the same histogram shape (covariates, coverage groups, brq_set_min_coverage_depth); a rank whose peers
 * do not show up within five seconds reports it with the next synchronising call.  world = 1 detaches. */
int brq_hist_exchange_export(brq_ctx* ctx, void* handle64, uint64_t* capacity_words);
int brq_hist_exchange_attach(brq_ctx* ctx, const void* handles, uint32_t world, uint32_t rank);

/* ---- between the passes: the coverage fit ---------------------------------------------------------
 * CoverageDistribution::fit (coverage_distribution.cpp:115-400, 422-436) without its plot: the censored negative-binomial fit
 * of a coverage group's unique-only coverage histogram and the deletion-propagation cutoff analyze_unique_coverage_distribution
 * stores in Summary::unique_coverage (:510-606) and stage 08 hands to identify_mutations() as deletion_propagation_cutoff.
 * deletion_propagation_pr_cutoff is the caller's (the reference uses 0.05 / sqrt(length of the group's sequences), :548).
 * brq_fit_coverage_distribution takes the histogram of the context's last brq_error_count (no file round trip);
 * brq_fit_coverage_file reads a <group>.unique_only_coverage_distribution.tab (host only: needs no device). */
typedef struct brq_coverage_fit {
  double average, variance, relative_variance;          /* Summary::unique_coverage[seq_id].average ... */
  double nbinom_size_parameter, nbinom_mean_parameter;  /* 0 / 0 = the fit failed */
  double deletion_coverage_propagation_cutoff;          /* -1 = the reference sequence itself is missing */
  uint32_t censor_start, censor_end;                    /* the fitting window around the histogram's peak */
} brq_coverage_fit;
int brq_fit_coverage_distribution(brq_ctx* ctx, uint32_t coverage_group, double deletion_propagation_pr_cutoff, brq_coverage_fit* out);
int brq_fit_coverage_file(brq_ctx* ctx, const char* distribution_file, double deletion_propagation_pr_cutoff, brq_coverage_fit* out);

/* ---- after pass 2: the Output stage's filter over the RA rows ---------------------------------------
 * test_RA_evidence (identify_mutations.cpp:687-749; called at breseq_cmdline.cpp:2614 between reading ra_mc_evidence.gd and
 * merging it into evidence.gd): every RA row is asked whether its variant is the consensus (score >= mutation cutoff, the 95 %
 * bound of its frequency >= consensus_frequency_cutoff -- the lower bound in consensus mode, the upper one in polymorphism
 * mode -- coverage minima, homopolymer rules) and, failing that, whether it is present at all (polymorphism cutoffs, strand and
 * quality bias p-values, both-strand coverage of the variant allele).  The row gains prediction=consensus|polymorphism and
 * consensus_reject= / polymorphism_reject= / reject= with the reasons; a row that answers neither is dropped in consensus mode
 * and kept with reject= in polymorphism mode; a consensus row whose major base is the reference base is dropped; user_defined
 * rows are never dropped.  gd_in -> gd_out, every other line as it was (the reference holds the rows in memory: its own
 * writer would add a #=TITLE line).  The members are breseq::Settings' (settings.h:473-500); brq_ra_filter_defaults fills them
 * with what settings.cpp:862-896 / :918-948 leaves there for the mode.  Host only.  counts5 (may be NULL): RA rows read,
 * predicted consensus, predicted polymorphism, kept rejected, deleted. */
typedef struct brq_ra_filter_options {
  int32_t polymorphism_prediction;
  double mutation_log10_e_value_cutoff;
  double consensus_frequency_cutoff;
  uint32_t consensus_minimum_variant_coverage, consensus_minimum_total_coverage;
  uint32_t consensus_minimum_variant_coverage_each_strand, consensus_minimum_total_coverage_each_strand;
  uint32_t consensus_reject_indel_homopolymer_length, consensus_reject_surrounding_homopolymer_length;
  double polymorphism_log10_e_value_cutoff;
  double polymorphism_frequency_cutoff;
  uint32_t polymorphism_minimum_variant_coverage, polymorphism_minimum_total_coverage;
  uint32_t polymorphism_minimum_variant_coverage_each_strand, polymorphism_minimum_total_coverage_each_strand;
  uint32_t polymorphism_reject_indel_homopolymer_length, polymorphism_reject_surrounding_homopolymer_length;
  double polymorphism_fisher_strand_p_value_cutoff;
  double polymorphism_ks_quality_p_value_cutoff;
  int32_t polymorphism_no_indels;
} brq_ra_filter_options;
void brq_ra_filter_defaults(int polymorphism_prediction, brq_ra_filter_options* out);
/* The exact (Clopper-Pearson) one-sided bounds on k / n the filter falls back to for rows without frequency_lower /
 * frequency_upper, i.e. evidence of an older breseq: binomial_frequency_lower_bound / _upper_bound (stats.h:134-135,
 * stats.cpp:2394-2414; the inverse incomplete beta behind them restated from Cephes like the reference's copy). */
void brq_binomial_frequency_bounds(double k, double n, double alpha, double* lower, double* upper);
/* fisher_strand_p_value of an RA row: the two-sided Fisher exact test (stats.cpp:2144-2171) on the strand counts of the minor
 * and the major allele, as pass 2's finalisation evaluates it (identify_mutations.cpp:3009-3033).  Exposed for known-answer
 * tests against the rows of the reference's own test suite. */
double brq_fisher_strand_p_value(uint32_t minor_top, uint32_t minor_bottom, uint32_t major_top, uint32_t major_bottom);
int brq_test_ra_evidence(brq_ctx* ctx, const char* gd_in, const char* fasta, const brq_ra_filter_options* options,
                         const char* gd_out, uint32_t* counts5);

/* The RA step of mutation prediction on the filtered file: MutationPredictor::predictRAtoSNPorDELorINSorSUB
 * (mutation_predictor.cpp:1955-2211, called from MutationPredictor::predict, :2946).  RA rows that lie inside an MC row are
 * marked deleted=1 (not in targeted_sequencing runs; not with call_mutations_overlapping_missing_coverage; never user_defined
 * rows); the accepted rows (prediction=consensus; in polymorphism mode every row without reject=) are taken in position
 * order, neighbouring ones joined (never polymorphisms), and each group becomes a SNP, DEL, INS or SUB row with its RA rows as
 * evidence, ids drawn like cGenomeDiff::new_unique_id, frequency= in polymorphism mode; consensus mode does not call inserted
 * columns that do not start at insert position 1.  gd_out = the '#' lines, the mutation rows in GenomeDiff order, the evidence
 * rows.  The three switches are Settings::polymorphism_prediction, ::targeted_sequencing,
 * ::call_mutations_overlapping_missing_coverage.  An input that already holds mutation rows is refused.  Host only.
 * counts5 (may be NULL): SNP, DEL, INS, SUB rows made, RA rows marked deleted. */
int brq_predict_ra_mutations(brq_ctx* ctx, const char* gd_in, const char* fasta, int polymorphism_prediction, int targeted_sequencing,
                             int call_mutations_overlapping_missing_coverage, const char* gd_out, uint32_t* counts5);

/* ---- one-call adapters with the reference entry points' argument meaning --------------------- */
int brq_run_error_count(brq_ctx* ctx, const char* bam, const char* fasta, const char* output_dir,
                        const char* error_rates_file, const char* const* readfiles, uint32_t n_readfiles,
                        int do_coverage, int do_errors, const char* covariates, const brq_stage_options* opt);
int brq_run_identify_mutations(brq_ctx* ctx, const char* bam, const char* fasta, const char* error_rates_file,
                               const char* gd_file, const double* deletion_propagation_cutoff,
                               const double* deletion_seed_cutoff, uint32_t n_targets, const brq_score_params* p,
                               int skip_missing_coverage_prediction, const brq_stage_options* opt);

/* bookkeeping for bench.py: kernels launched so far, device milliseconds of the last call of each kernel */
int brq_launch_count(void);
/* CUDA events on the ctx stream (slots 0..3) so callers can time a region on the launching stream */
int brq_event_record(brq_ctx* ctx, int slot);
int brq_event_elapsed_ms(brq_ctx* ctx, int slot_a, int slot_b, float* ms);
int brq_kernel_ms(brq_ctx* ctx, float* hist_ms, float* coverage_ms, float* derive_ms, float* score_ms);
/* the scoring pass is two kernels: the streaming tally and the EM fit of the slots that need one */
int brq_score_phase_ms(brq_ctx* ctx, float* tally_ms, float* fit_ms);

#ifdef __cplusplus
}
#endif
#endif
