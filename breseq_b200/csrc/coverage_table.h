// BAM2COV's per-position coverage table (coverage_output.cpp:190-283): the host writer over the columns the device walk counted
// (expand_core.h: coverage_lane).
#pragma once
#include "bam_io.h"
#include "brq_types.h"
#include "expand_core.h"

#include <string>
#include <vector>

namespace brq {

// by_group: empty, or one array per read group (the table then repeats its columns per group, prefixed "RG-<n>_").
// reference_average: BAM2COV -a, the fit average of the sequence from breseq's summary.json as one more '#' line in front of
// the region's own averages (coverage_output.cpp:259-262); null = not asked for.
void write_coverage_table(const std::string& path, const BamHeader& hdr, const RefSet& ref, const PileupStream& st,
                          const std::vector<CoverageColumn>& cols, const std::vector<std::vector<CoverageColumn>>& by_group,
                          const std::string& region, uint32_t resolution, bool total_only, bool csv, const double* reference_average = nullptr);

}  // namespace brq
