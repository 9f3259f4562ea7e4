// TEST INFRASTRUCTURE -- command-line driver for the CPU oracle (see oracle.h).
//   oracle_cli error_count        --bam B --fasta F --out DIR --covariates S [--readfiles a,b] [--read-sets n:2,m:1]
//                                 [--seq-ids a,b] [--counts-dump FILE] [--error-rates FILE] [--no-coverage] [--no-errors] [--preprocess]
//   oracle_cli identify_mutations --bam B --fasta F --error-rates FILE --gd OUT.gd [--del-prop a,b] [--del-seed a,b]
//                                 [--mutation-cutoff 10] [--polymorphism-cutoff 2] [--precision 1e-6] [--places 8]
//                                 [--base-quality-cutoff 3] [--skip-mc] [--columns-out FILE] [--read-sets ...] [--seq-ids ...]
// Prints one JSON line with the wall time of the call (steady_clock around the entry point, the
// same bracket the reference's own ExecutionTime uses: settings.cpp:2046-2063).
#include "oracle.h"

#include "htslib/faidx.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>

using namespace std;

static vector<string> split(const string& s, char sep) {
  vector<string> out;
  if (s.empty()) return out;
  stringstream ss(s);
  string item;
  while (getline(ss, item, sep)) out.push_back(item);
  return out;
}

int main(int argc, char** argv) {
  if (argc < 2) { cerr << "usage: oracle_cli error_count|identify_mutations ..." << endl; return 2; }
  string cmd = argv[1];
  map<string, string> opt;
  for (int i = 2; i < argc; ++i) {
    string k = argv[i];
    if (k.rfind("--", 0) != 0) { cerr << "bad argument " << k << endl; return 2; }
    k = k.substr(2);
    if (k == "no-coverage" || k == "no-errors" || k == "skip-mc" || k == "preprocess") opt[k] = "1";
    else if (i + 1 < argc) opt[k] = argv[++i];
  }
  auto get = [&](const string& k, const string& d) { return opt.count(k) ? opt[k] : d; };

  oracle::Settings st;
  faidx_t* fai = fai_load(get("fasta", "").c_str());
  if (!fai) { cerr << "cannot open fasta" << endl; return 1; }
  vector<string> names;
  for (int i = 0; i < faidx_nseq(fai); ++i) {
    names.push_back(faidx_iseq(fai, i));
    st.total_reference_sequence_length += (uint64_t)faidx_seq_len(fai, faidx_iseq(fai, i));
  }
  fai_destroy(fai);
  vector<string> seq_ids = split(get("seq-ids", ""), ',');
  if (seq_ids.empty()) seq_ids = names;
  for (const string& s : seq_ids) st.call_mutations_seq_ids.insert(s);
  for (size_t i = 0; i < names.size(); ++i) st.seq_id_to_coverage_group[names[i]] = (uint32_t)i;  // one group per reference
  st.n_coverage_groups = (uint32_t)names.size();
  for (const string& s : split(get("read-sets", ""), ',')) {
    size_t c = s.rfind(':');
    oracle::ReadFileSet rs;
    rs.base_name = s.substr(0, c);
    rs.n_files = (uint32_t)atoi(s.substr(c + 1).c_str());
    st.read_file_sets.push_back(rs);
  }
  st.base_quality_cutoff = (uint32_t)atoi(get("base-quality-cutoff", "3").c_str());
  st.skip_missing_coverage_prediction = opt.count("skip-mc") > 0;
  st.polymorphism_prediction = opt.count("polymorphism-prediction") > 0;
  st.user_evidence_genome_diff_file_name = get("user-evidence", "");
  string out = get("out", ".");
  st.error_rates_file_name = get("error-rates", out + "/error_rates.tab");
  st.unique_only_coverage_distribution_file_name = "@.unique_only_coverage_distribution.tab";
  st.base_qual_error_prob_file_name = "base_qual_error_prob.#.tab";

  auto t0 = chrono::steady_clock::now();
  unsigned long long records = 0;
  if (cmd == "error_count") {
    // --preprocess: the stage 03 call (breseq_cmdline.cpp:1969); Summary::preprocess_error_count goes to <out>/preprocess_error_count.tab
    map<string, double> pr;
    oracle::error_count(st, get("bam", ""), get("fasta", ""), out, split(get("readfiles", ""), ','), !opt.count("no-coverage"),
                        !opt.count("no-errors"), get("covariates", ""), get("counts-dump", ""), opt.count("preprocess") > 0, &pr);
    if (opt.count("preprocess")) {
      FILE* f = fopen((out + "/preprocess_error_count.tab").c_str(), "w");
      if (!f) { cerr << "cannot write preprocess_error_count.tab" << endl; return 1; }
      for (const auto& kv : pr) fprintf(f, "%s\t%.17g\n", kv.first.c_str(), kv.second);
      fclose(f);
    }
  } else if (cmd == "identify_mutations") {
    vector<double> prop, seed;
    for (const string& s : split(get("del-prop", ""), ',')) prop.push_back(atof(s.c_str()));
    for (const string& s : split(get("del-seed", ""), ',')) seed.push_back(atof(s.c_str()));
    while (prop.size() < names.size()) prop.push_back(prop.empty() ? 0.0 : prop.back());
    while (seed.size() < names.size()) seed.push_back(seed.empty() ? 0.0 : seed.back());
    uint64_t n = 0;
    oracle::identify_mutations(st, get("bam", ""), get("fasta", ""), get("gd", out + "/ra_mc_evidence.gd"), prop, seed,
                               atof(get("mutation-cutoff", "10").c_str()), atof(get("polymorphism-cutoff", "2").c_str()),
                               atof(get("precision", "1e-6").c_str()), (uint32_t)atoi(get("places", "8").c_str()),
                               get("columns-out", ""), &n);
    records = n;
  } else {
    cerr << "unknown command " << cmd << endl;
    return 2;
  }
  double sec = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
  printf("{\"cmd\": \"%s\", \"seconds\": %.6f, \"records\": %llu}\n", cmd.c_str(), sec, records);
  return 0;
}
