// Negative-binomial fit of a unique-only coverage histogram and the deletion-propagation cutoff it implies: what breseq
// computes between error_count() and identify_mutations() (coverage_distribution.cpp:115-400 and :510-606), so that the
// MC cutoffs of pass 2 come from pass 1's histogram without the file round trip.
#pragma once
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

namespace brq {

struct CoverageFit {
  double average = 0, variance = 0, relative_variance = 0;   // of the histogram, index 0 ignored
  double nb_size = 0, nb_mu = 0;                             // 0 / 0 = no fit
  double deletion_coverage_propagation_cutoff = 0;           // -1 = the reference sequence itself is missing
  uint32_t censor_start = 0, censor_end = 0;                 // the fitting window around the peak
};

// n[i] = reference positions of unique coverage i (n[0] is not used), N = the highest index.  parallel_for(n_jobs, job(i)) runs
// the independent restarts side by side when given (any order: the best one is picked afterwards, in the reference's order).
CoverageFit fit_coverage_distribution(const std::vector<double>& n, uint32_t N, double deletion_propagation_pr_cutoff,
                                      const std::function<void(size_t, const std::function<void(size_t)>&)>* parallel_for = nullptr);

// the histogram of a <group>.unique_only_coverage_distribution.tab file (coverage_distribution.cpp:34-65)
void read_coverage_distribution(const std::string& path, std::vector<double>& n, uint32_t& N);

// P(X <= k) of a negative binomial in (size, mu) form and its quantile (stats.cpp:982-991, 2047-2072)
double nbinom_cdf(double k, double size, double mu);
uint32_t nbinom_quantile(double target_pr, double size, double mu);

// Exact (Clopper-Pearson) one-sided confidence bounds on k / n (stats.cpp:2394-2414, through the inverse incomplete beta)
double binomial_frequency_lower_bound(double k, double n, double alpha = 0.05);
double binomial_frequency_upper_bound(double k, double n, double alpha = 0.05);

}  // namespace brq
