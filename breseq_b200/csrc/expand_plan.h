// Host-side plan of a device staging call: the visited segments with their read ranges, tiles and reference bytes, and the
// read-file partition.  Plain C++ (shared by expand.cu and tests/expand_check.cpp).
#pragma once
#include "expand_core.h"
#include "staging.h"

namespace brq {

struct ExpandPlan {
  std::vector<ExpandSeg> segs;
  std::vector<uint8_t> refbytes;      // per segment: the reference characters of its columns and one byte beyond (ExpandSeg::ref_off)
  std::vector<int32_t> seg_of_tid;    // by BAM tid, -1 = not visited
  std::vector<uint32_t> part;         // [0, n_part): flat read-file index of every read group; [n_part, 2 n_part): its file count
  uint32_t n_part = 0, tiles = 0;
};

// Fills st.segments / n_base / n_groups (plan_segments) and the plan.  `tid`: BAM tid of every read, in file order.
void make_expand_plan(const BamHeader& hdr, const RefSet& ref, const int32_t* tid, size_t n_reads, const StageConfig& cfg, PileupStream& st,
                      ExpandPlan& plan);

}  // namespace brq
