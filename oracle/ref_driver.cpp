// TEST INFRASTRUCTURE -- command-line driver around the REFERENCE'S OWN, UNMODIFIED sources.
//
// oracle/ref_build.sh compiles /root/reference/src/breseq/*.cpp where they lie (every libbreseq
// source) against the htslib-compatible shim in hts_shim/ and links them with this file into
// oracle/_ref/ref_cli.  Nothing here restates reference logic: the
// driver fills breseq::Settings / Summary the way breseq_cmdline.cpp does at its two call sites
//   stage 07  breseq_cmdline.cpp:2276-2328  -> breseq::error_count()          error_count.h:41-52
//   stage 08  breseq_cmdline.cpp:2454-2488  -> breseq::identify_mutations()   identify_mutations.h:46-60
// (and like the standalone `breseq ERROR_COUNT`, breseq_cmdline.cpp:972-1030) and calls them.
// Same command line and same JSON timing line as oracle_cli, so tests and bench.py can swap the two.
//
// Used to (a) pin the restatement in oracle.cpp against the reference's real arithmetic and file
// writers, and (b) serve as the `kind: "reference"` CPU baseline.  The htslib layer under it is
// still the shim (htslib itself is not in this image): parity at the BAM-decode/pileup boundary
// remains a restatement.
#include "coverage_distribution.h"
#include "coverage_output.h"
#include "error_count.h"
#include "identify_mutations.h"
#include "mutation_predictor.h"
#include "reference_sequence.h"
#include "settings.h"
#include "stats.h"
#include "summary.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <map>
#include <sstream>

using namespace std;
using namespace breseq;

static vector<string> split_list(const string& s, char sep) {
  vector<string> out;
  if (s.empty()) return out;
  stringstream ss(s);
  string item;
  while (getline(ss, item, sep)) out.push_back(item);
  return out;
}

int main(int argc, char** argv) {
  if (argc < 2) { cerr << "usage: ref_cli error_count|identify_mutations ..." << endl; return 2; }
  string cmd = argv[1];
  map<string, string> opt;
  for (int i = 2; i < argc; ++i) {
    string k = argv[i];
    if (k.rfind("--", 0) != 0) { cerr << "bad argument " << k << endl; return 2; }
    k = k.substr(2);
    if (k == "no-coverage" || k == "no-errors" || k == "skip-mc" || k == "polymorphism-prediction" || k == "preprocess") opt[k] = "1";
    else if (i + 1 < argc) opt[k] = argv[++i];
  }
  auto get = [&](const string& k, const string& d) { return opt.count(k) ? opt[k] : d; };
  const string out = get("out", ".");

  if (cmd == "fit_coverage") {
    // CoverageDistribution::fit (coverage_distribution.cpp:422-498) on a <group>.unique_only_coverage_distribution.tab with the
    // probability cutoff analyze_unique_coverage_distribution derives (:548: 0.05 / sqrt(sequence length)).  fit() also draws
    // its plot through `gnuplot`: the caller puts a do-nothing stand-in on PATH (tests/golden/make_golden.py).
    const double pr = atof(get("pr-cutoff", "0.01").c_str());
    CoverageDistributionFitResult r = CoverageDistribution::fit(get("distribution", ""), out + "/coverage_plot.svg", pr);
    char line[512];
    snprintf(line, sizeof line, "average\t%.17g\nvariance\t%.17g\nrelative_variance\t%.17g\nnb_fit_size\t%.17g\nnb_fit_mu\t%.17g\n"
             "deletion_coverage_propagation_cutoff\t%.17g\n", r.average, r.variance, r.relative_variance, r.nb_fit_size, r.nb_fit_mu,
             r.deletion_coverage_propagation_cutoff);
    cout << line;
    return 0;
  }

  if (cmd == "binomial_bounds") {
    // binomial_frequency_lower_bound / _upper_bound (stats.cpp:2394-2414) for "k n" pairs read from stdin, at --alpha
    const double alpha = atof(get("alpha", "0.05").c_str());
    double k, n;
    while (cin >> k >> n) printf("%.17g\t%.17g\n", binomial_frequency_lower_bound(k, n, alpha), binomial_frequency_upper_bound(k, n, alpha));
    return 0;
  }

  if (cmd == "coverage_table") {
    // `breseq BAM2COV -t`: coverage_output::table (coverage_output.cpp:190-283) --bam --fasta --region seq:start-end
    // --resolution N (0 = every position) [--total-only 1] [--per-read-group 1] [--format tsv|csv] --table FILE
    coverage_output co(get("bam", ""), get("fasta", ""));
    co.total_only(get("total-only", "0") == "1");
    co.output_format(get("format", "tsv"));
    co.per_read_group(get("per-read-group", "0") == "1");
    co.table(get("region", ""), get("table", out + "/coverage.tab"), (uint32_t)atoi(get("resolution", "0").c_str()));
    return 0;
  }

  Settings::set_global_paths();  // as breseq_cmdline.cpp:2885 does first thing in main()
  Summary summary;
  Settings settings(out);

  cReferenceSequences ref_seq_info;
  vector<string> reference_file_names;
  reference_file_names.push_back(get("fasta", ""));
  ref_seq_info.LoadFiles(reference_file_names);
  settings.normal_reference_file_names = reference_file_names;
  settings.init_reference_sets(ref_seq_info);
  summary.sequence_conversion.total_reference_sequence_length = ref_seq_info.get_total_length();

  // --read-sets name:2,name:1  -> cReadFileSets (settings.h:122-163); one @RG per set, LB = base name
  uint32_t id = 0;
  for (const string& s : split_list(get("read-sets", ""), ',')) {
    size_t c = s.rfind(':');
    cReadFileSet rfs;
    rfs.m_base_name = s.substr(0, c);
    uint32_t n_files = (uint32_t)atoi(s.substr(c + 1).c_str());
    for (uint32_t f = 0; f < n_files; ++f) {
      cReadFile rf;
      rf.m_base_name = n_files == 2 ? rfs.m_base_name + (f == 0 ? "_R1" : "_R2") : rfs.m_base_name;
      rf.m_original_file_name = rf.m_base_name + ".fastq";
      rf.m_paired_end_group = (uint32_t)settings.read_file_sets.size();
      rf.m_error_group = id;
      rf.m_id = id++;
      rfs.m_files.push_back(rf);
    }
    settings.read_file_sets.push_back(rfs);
  }

  settings.base_quality_cutoff = (uint32_t)atoi(get("base-quality-cutoff", "3").c_str());
  settings.skip_missing_coverage_prediction = opt.count("skip-mc") > 0;
  settings.user_evidence_genome_diff_file_name = get("user-evidence", "");  // Settings::user_evidence_genome_diff_file_name (--user-evidence-gd)
  settings.polymorphism_prediction = opt.count("polymorphism-prediction") > 0;
  settings.error_rates_file_name = get("error-rates", out + "/error_rates.tab");
  settings.unique_only_coverage_distribution_file_name = out + "/@.unique_only_coverage_distribution.tab";
  settings.error_rates_base_qual_error_prob_file_name = out + "/base_qual_error_prob.#.tab";
  settings.mutation_identification_per_position_file_name = get("per-position", out + "/per_position_file.tab");
  settings.dp_candidate_regions_file_name = out + "/dp_candidate_regions.csv";
  if (opt.count("coverage-tsv")) {  // <seq>.coverage.tsv (identify_mutations.cpp:2028-2052)
    settings.predict_copy_number = true;
    settings.complete_coverage_text_file_name = get("coverage-tsv", out + "/@.coverage.tsv");
  }
  // user-chosen subset of targets (Settings::call_mutations_seq_id_set())
  vector<string> seq_ids = split_list(get("seq-ids", ""), ',');
  if (!seq_ids.empty()) {
    settings.refseq_settings.m_call_mutations_seq_id_set.clear();
    for (const string& s : seq_ids) settings.refseq_settings.m_call_mutations_seq_id_set.insert(s);
  }

  auto t0 = chrono::steady_clock::now();
  if (cmd == "error_count") {
    error_count(settings, summary, get("bam", ""), get("fasta", ""), out, split_list(get("readfiles", ""), ','),
                !opt.count("no-coverage"), !opt.count("no-errors"), opt.count("preprocess") > 0, (uint8_t)settings.base_quality_cutoff,
                get("covariates", ""));
    if (opt.count("preprocess")) {  // the stage 03 call (breseq_cmdline.cpp:1969): what it leaves in the Summary
      FILE* f = fopen((out + "/preprocess_error_count.tab").c_str(), "w");
      if (!f) { cerr << "cannot write preprocess_error_count.tab" << endl; return 1; }
      for (const auto& kv : summary.preprocess_error_count) fprintf(f, "%s\t%.17g\n", kv.first.c_str(), kv.second.no_pos_hash_per_position_pr);
      fclose(f);
    }
  } else if (cmd == "identify_mutations") {
    vector<double> prop, seed;
    for (const string& s : split_list(get("del-prop", ""), ',')) prop.push_back(atof(s.c_str()));
    for (const string& s : split_list(get("del-seed", ""), ',')) seed.push_back(atof(s.c_str()));
    while (prop.size() < ref_seq_info.size()) prop.push_back(prop.empty() ? 0.0 : prop.back());
    while (seed.size() < ref_seq_info.size()) seed.push_back(seed.empty() ? 0.0 : seed.back());
    identify_mutations(settings, summary, get("bam", ""), get("fasta", ""), get("gd", out + "/ra_mc_evidence.gd"), ref_seq_info,
                       prop, seed, atof(get("mutation-cutoff", "10").c_str()), atof(get("polymorphism-cutoff", "2").c_str()),
                       atof(get("precision", "1e-6").c_str()), (uint32_t)atoi(get("places", "8").c_str()),
                       opt.count("per-position") > 0);
  } else if (cmd == "test_ra") {
    // the Output stage's filter over the RA rows (breseq_cmdline.cpp:2609-2614 -> test_RA_evidence, identify_mutations.cpp:687-749):
    // --gd-in FILE --gd-out FILE, thresholds as --<Settings member> VALUE (the members settings.cpp:862-948 sets per mode)
    struct { const char* name; double* d; uint32_t* u; bool* b; } knobs[] = {
      {"mutation_log10_e_value_cutoff", &settings.mutation_log10_e_value_cutoff, 0, 0},
      {"consensus_frequency_cutoff", &settings.consensus_frequency_cutoff, 0, 0},
      {"consensus_minimum_variant_coverage", 0, &settings.consensus_minimum_variant_coverage, 0},
      {"consensus_minimum_total_coverage", 0, &settings.consensus_minimum_total_coverage, 0},
      {"consensus_minimum_variant_coverage_each_strand", 0, &settings.consensus_minimum_variant_coverage_each_strand, 0},
      {"consensus_minimum_total_coverage_each_strand", 0, &settings.consensus_minimum_total_coverage_each_strand, 0},
      {"consensus_reject_indel_homopolymer_length", 0, &settings.consensus_reject_indel_homopolymer_length, 0},
      {"consensus_reject_surrounding_homopolymer_length", 0, &settings.consensus_reject_surrounding_homopolymer_length, 0},
      {"polymorphism_log10_e_value_cutoff", &settings.polymorphism_log10_e_value_cutoff, 0, 0},
      {"polymorphism_frequency_cutoff", &settings.polymorphism_frequency_cutoff, 0, 0},
      {"polymorphism_minimum_variant_coverage", 0, &settings.polymorphism_minimum_variant_coverage, 0},
      {"polymorphism_minimum_total_coverage", 0, &settings.polymorphism_minimum_total_coverage, 0},
      {"polymorphism_minimum_variant_coverage_each_strand", 0, &settings.polymorphism_minimum_variant_coverage_each_strand, 0},
      {"polymorphism_minimum_total_coverage_each_strand", 0, &settings.polymorphism_minimum_total_coverage_each_strand, 0},
      {"polymorphism_reject_indel_homopolymer_length", 0, &settings.polymorphism_reject_indel_homopolymer_length, 0},
      {"polymorphism_reject_surrounding_homopolymer_length", 0, &settings.polymorphism_reject_surrounding_homopolymer_length, 0},
      {"polymorphism_fisher_strand_p_value_cutoff", &settings.polymorphism_fisher_strand_p_value_cutoff, 0, 0},
      {"polymorphism_ks_quality_p_value_cutoff", &settings.polymorphism_ks_quality_p_value_cutoff, 0, 0},
      {"polymorphism_no_indels", 0, 0, &settings.polymorphism_no_indels},
    };
    for (auto& k : knobs) {
      if (!opt.count(k.name)) continue;
      const string v = opt[k.name];
      if (k.d) *k.d = atof(v.c_str());
      if (k.u) *k.u = (uint32_t)strtoul(v.c_str(), 0, 10);
      if (k.b) *k.b = atoi(v.c_str()) != 0;
    }
    cGenomeDiff gd;
    gd.read(get("gd-in", ""));
    test_RA_evidence(gd, ref_seq_info, settings);
    gd.write(get("gd-out", out + "/filtered.gd"));
  } else if (cmd == "predict_ra") {
    // the RA step of mutation prediction (MutationPredictor::predict, mutation_predictor.cpp:2946 ->
    // predictRAtoSNPorDELorINSorSUB, :1955-2211) on an evidence file test_ra has been through: --gd-in FILE --gd-out FILE
    // [--polymorphism-prediction] [--targeted-sequencing 1] [--call-mutations-overlapping-missing-coverage 1]
    settings.targeted_sequencing = get("targeted-sequencing", "0") == "1";
    settings.call_mutations_overlapping_missing_coverage = get("call-mutations-overlapping-missing-coverage", "0") == "1";
    cGenomeDiff gd;
    gd.read(get("gd-in", ""));
    MutationPredictor mp(ref_seq_info);
    diff_entry_list_t ra = gd.get_list(make_vector<gd_entry_type>(RA));
    diff_entry_list_t mc = gd.get_list(make_vector<gd_entry_type>(MC));
    mp.predictRAtoSNPorDELorINSorSUB(settings, summary, gd, ra, mc);
    gd.write(get("gd-out", out + "/predicted.gd"));
  } else {
    cerr << "unknown command " << cmd << endl;
    return 2;
  }
  double sec = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
  printf("{\"cmd\": \"%s\", \"seconds\": %.6f, \"records\": 0, \"impl\": \"reference\"}\n", cmd.c_str(), sec);
  return 0;
}
