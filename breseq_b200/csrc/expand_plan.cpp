#include "expand_plan.h"

#include <cstring>
#include <stdexcept>

namespace brq {

void make_expand_plan(const BamHeader& hdr, const RefSet& ref, const int32_t* tid, size_t n_reads, const StageConfig& cfg, PileupStream& st,
                      ExpandPlan& plan) {
  const size_t n_targets = hdr.target_names.size();
  std::vector<const std::string*> refseq;
  plan_segments(hdr, ref, cfg, st, refseq);
  const size_t n_visit = st.segments.size();
  if (st.n_base >= 0xFFFFFFF0ull) throw std::runtime_error("more than 2^32 - 2 slots in one staged stream");
  if (n_reads >= 0xFFFFFFF0ull) throw std::runtime_error("more than 2^32 reads in one staging call");
  // read ranges per target (the BAM is coordinate sorted: the device checks it)
  std::vector<uint32_t> t_first(n_targets, 0), t_last(n_targets, 0);
  for (size_t i = 0; i < n_reads; ++i) {
    const int32_t t = tid[i];
    if (t < 0 || (size_t)t >= n_targets) continue;
    if (t_last[(size_t)t] == 0) t_first[(size_t)t] = (uint32_t)i;
    t_last[(size_t)t] = (uint32_t)i + 1;
  }
  plan.segs.assign(n_visit, ExpandSeg());
  plan.refbytes.assign((size_t)st.n_base + n_visit + 16, 0);
  plan.seg_of_tid.assign(n_targets ? n_targets : 1, -1);
  plan.tiles = 0;
  uint32_t ref_off = 0;
  for (size_t v = 0; v < n_visit; ++v) {
    const Segment& sg = st.segments[v];
    ExpandSeg& e = plan.segs[v];
    e.tid = sg.tid; e.lo = sg.lo; e.hi = sg.hi; e.tlen = (int32_t)hdr.target_lens[(size_t)sg.tid];
    e.slot0 = (uint32_t)sg.slot0; e.tile0 = plan.tiles; plan.tiles += (uint32_t)((sg.hi - sg.lo + 31) / 32);
    e.read_first = t_first[(size_t)sg.tid]; e.read_last = t_last[(size_t)sg.tid];
    e.ref_off = ref_off;
    const uint32_t g = cfg.coverage_group_of_tid.empty() ? (uint32_t)sg.tid : cfg.coverage_group_of_tid[(size_t)sg.tid];
    if (g > 255) throw std::runtime_error("more than 256 coverage groups are not supported");
    e.group = g;
    const std::string& rs = *refseq[v];
    const size_t take = std::min<size_t>((size_t)(sg.hi - sg.lo) + 1, rs.size() - (size_t)sg.lo);
    memcpy(&plan.refbytes[ref_off], rs.data() + sg.lo, take);
    ref_off += (uint32_t)(sg.hi - sg.lo) + 1;
    plan.seg_of_tid[(size_t)sg.tid] = (int32_t)v;
  }
  std::vector<uint32_t> part_base, part_count;
  make_read_file_partition(hdr.read_groups, cfg.read_file_sets, part_base, part_count);
  plan.n_part = (uint32_t)part_base.size();
  plan.part.assign(2 * (size_t)plan.n_part + 2, 0);
  for (uint32_t g = 0; g < plan.n_part; ++g) { plan.part[g] = part_base[g]; plan.part[plan.n_part + g] = part_count[g]; }
}

}  // namespace brq
