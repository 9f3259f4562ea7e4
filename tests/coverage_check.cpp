// CPU check of BAM2COV's per-column logic (csrc/expand_core.h: coverage_lane, the function csrc/expand.cu's
// coverage_tile_kernel wraps): the same host+device inline function is run here serially over a BAM, tile by tile and lane by
// lane, and the table is written by the product's own writer (csrc/coverage_table.cpp).  tests/test_coverage_table.py compares
// the file with what the reference build's coverage_output::table wrote.  Test infrastructure only: the product never runs
// this walk on the host.
//
//   coverage_check BAM FASTA REGION RESOLUTION TOTAL_ONLY CSV OUT [PER_READ_GROUP]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../breseq_b200/csrc/bam_io.h"
#include "../breseq_b200/csrc/coverage_table.h"
#include "../breseq_b200/csrc/expand_core.h"
#include "../breseq_b200/csrc/expand_plan.h"
#include "../breseq_b200/csrc/staging.h"

using namespace brq;

int main(int argc, char** argv) {
  if (argc < 8 || argc > 10) { fprintf(stderr, "usage: coverage_check BAM FASTA REGION RESOLUTION TOTAL_ONLY CSV OUT [PER_READ_GROUP [REFERENCE_AVERAGE]]\n"); return 2; }
  try {
    BamHeader hdr; ReadBatch R; RefSet ref;
    read_fasta(argv[2], ref);
    read_bam(argv[1], hdr, R, 2);
    StageConfig cfg;
    cfg.threads = 2;
    PileupStream D;
    ExpandPlan plan;
    make_expand_plan(hdr, ref, R.tid.data(), R.tid.size(), cfg, D, plan);
    RawReads raw;
    raw.tid = R.tid.data(); raw.pos = R.pos.data(); raw.flag = R.flag.data(); raw.mapq = R.mapq.data(); raw.rg = R.rg.data();
    raw.x1 = R.x1.data(); raw.xl = R.xl.data(); raw.xr = R.xr.data(); raw.l_seq = R.l_seq.data(); raw.seq_off = R.seq_off.data();
    raw.n_cigar = R.n_cigar.data(); raw.cigar_off = R.cigar_off.data(); raw.bases = R.bases.data(); raw.quals = R.quals.data();
    raw.cigars = R.cigars.data(); raw.n = R.size();
    std::vector<ReadMeta> meta(R.size() + 1);
    std::vector<int32_t> max_span(hdr.target_names.size() + 1, 1);
    std::vector<uint32_t> stats(XS_WORDS, 0);
    for (uint64_t i = 0; i < raw.n; ++i) meta[i] = prep_read(raw, i, plan.part.data(), plan.part.data() + plan.n_part, plan.n_part, max_span.data(), stats.data());
    ExpandArgs a;
    memset(static_cast<void*>(&a), 0, sizeof a);
    a.meta = meta.data(); a.pos = R.pos.data(); a.tid = R.tid.data(); a.bases = R.bases.data(); a.quals = R.quals.data(); a.cigars = R.cigars.data();
    a.n_reads = raw.n; a.seg = plan.segs.data(); a.n_seg = (uint32_t)plan.segs.size(); a.n_tiles = plan.tiles; a.max_span = max_span.data();
    a.seg_of_tid = plan.seg_of_tid.data(); a.ref = plan.refbytes.data(); a.n_base = (uint32_t)D.n_base;
    std::vector<CoverageColumn> cols(D.n_base + 1);
    for (uint32_t t = 0; t < plan.tiles; ++t) for (uint32_t l = 0; l < 32; ++l) coverage_lane(a, t, l, cols.data());
    std::vector<std::vector<CoverageColumn>> by_group;
    if (argc >= 9 && atoi(argv[8])) {
      by_group.resize(hdr.read_groups.ids.size() > 1 ? hdr.read_groups.ids.size() : 1);
      for (size_t g = 0; g < by_group.size(); ++g) {
        by_group[g].resize(D.n_base + 1);
        for (uint32_t t = 0; t < plan.tiles; ++t) for (uint32_t l = 0; l < 32; ++l) coverage_lane(a, t, l, by_group[g].data(), (uint32_t)g);
      }
    }
    const double reference_average = argc >= 10 ? atof(argv[9]) : 0;   // BAM2COV -a
    write_coverage_table(argv[7], hdr, ref, D, cols, by_group, argv[3], (uint32_t)atoi(argv[4]), atoi(argv[5]) != 0, atoi(argv[6]) != 0,
                         argc >= 10 ? &reference_average : nullptr);
  } catch (const std::exception& e) {
    fprintf(stderr, "coverage_check: %s\n", e.what());
    return 1;
  }
  return 0;
}
