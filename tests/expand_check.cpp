// CPU check of the device expander's per-read / per-(read, column) logic (csrc/expand_core.h) against the host staging
// layer (csrc/staging.cpp): the same host+device inline functions the kernels of csrc/expand.cu wrap are run here serially,
// tile by tile and lane by lane, and every array of the stream they build must equal stage()'s, bit for bit.
// Test infrastructure only: the product never runs these functions on the host.  What this cannot see (the scan, sort and
// partition kernels of expand.cu) is restated here with std:: algorithms and checked on the GPU by tests/test_gpu_expand.py.
//
//   expand_check [dataset ...]     exit code 0 = all equal
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../breseq_b200/csrc/expand_core.h"
#include "../breseq_b200/csrc/expand_plan.h"
#include "../breseq_b200/csrc/staging.h"
#include "../breseq_b200/csrc/synth.h"

using namespace brq;

namespace {

struct Case {
  std::string name;
  uint64_t seed;
  std::vector<uint32_t> contigs;
  std::vector<SynthReadSet> sets;
  uint32_t n_poly, n_fixed, n_gaps;
  bool read_pos, base_repeat, preprocess;
  uint32_t shard_rank, shard_count;
  bool want_hist = true, want_score = true;
  std::string bam, fasta;   // reads from a BAM on disk instead of the generator (expand_check --bam BAM FASTA READ_SET [paired])
};

SynthReadSet rs(const char* name, bool paired, uint32_t len, double cov, double fm = 400, double fs = 40) {
  SynthReadSet s; s.name = name; s.paired = paired; s.read_len = len; s.coverage = cov; s.frag_mean = fm; s.frag_sd = fs; return s;
}

template <class T>
bool same(const char* what, const T* a, const T* b, size_t n, const std::string& cs) {
  for (size_t i = 0; i < n; ++i)
    if (a[i] != b[i]) { fprintf(stderr, "%s: %s differs at %zu of %zu: host %llx device %llx\n", cs.c_str(), what, i, n, (unsigned long long)a[i], (unsigned long long)b[i]); return false; }
  return true;
}

bool run_case(const Case& c) {
  RefSet ref; BamHeader hdr; ReadBatch R; std::vector<SynthVariant> variants;
  if (!c.bam.empty()) {
    read_fasta(c.fasta, ref);
    read_bam(c.bam, hdr, R, 4);
  } else {
    synth_reference(c.seed, c.contigs, "ctg", ref);
    SynthConfig sc; sc.seed = c.seed; sc.sets = c.sets; sc.n_polymorphic = c.n_poly; sc.n_fixed = c.n_fixed; sc.n_gaps = c.n_gaps; sc.threads = 4;
    synth_reads(sc, ref, hdr, R, variants);
    // a few flagged reads the pileup keeps (secondary, QC fail, duplicate) and one it drops (unmapped with a position)
    for (size_t i = 0; i < R.size(); i += 97) R.flag[i] |= (i % 3 == 0 ? 256 : i % 3 == 1 ? 512 : 1024);
    if (R.size() > 50) R.flag[50] |= 4;
  }

  StageConfig cfg;
  cfg.threads = 4;
  for (const SynthReadSet& s : c.sets) cfg.read_file_sets.push_back({s.name, s.paired ? 2u : 1u});
  cfg.use_read_pos = c.read_pos; cfg.use_base_repeat = c.base_repeat; cfg.preprocess_stage = c.preprocess;
  cfg.shard_rank = c.shard_rank; cfg.shard_count = c.shard_count;
  cfg.want_hist = c.want_hist; cfg.want_score = c.want_score;
  cfg.compact_score = false;
  PileupStream H;
  stage(hdr, ref, R, cfg, H);

  // ------------------------------------------------------------------ the device sequence, serially
  PileupStream D;
  ExpandPlan plan;
  make_expand_plan(hdr, ref, R.tid.data(), R.tid.size(), cfg, D, plan);
  const uint32_t n_base = (uint32_t)D.n_base;
  const size_t n_targets = hdr.target_names.size();
  RawReads raw;
  raw.tid = R.tid.data(); raw.pos = R.pos.data(); raw.flag = R.flag.data(); raw.mapq = R.mapq.data(); raw.rg = R.rg.data();
  raw.x1 = R.x1.data(); raw.xl = R.xl.data(); raw.xr = R.xr.data(); raw.l_seq = R.l_seq.data(); raw.seq_off = R.seq_off.data();
  raw.n_cigar = R.n_cigar.data(); raw.cigar_off = R.cigar_off.data(); raw.bases = R.bases.data(); raw.quals = R.quals.data();
  raw.cigars = R.cigars.data(); raw.n = R.size();
  std::vector<ReadMeta> meta(R.size() + 1);
  std::vector<int32_t> max_span(n_targets + 1, 1);
  std::vector<uint32_t> stats(XS_WORDS, 0);
  std::vector<uint64_t> geo(384, 0);
  for (uint64_t i = 0; i < raw.n; ++i) {
    meta[i] = prep_read(raw, i, plan.part.data(), plan.part.data() + plan.n_part, plan.n_part, max_span.data(), stats.data());
    if (cfg.want_score && (meta[i].flags & RM_LIVE) && meta[i].x1 == 1) {
      geo[meta[i].mapq] += meta[i].l_seq;
      for (uint32_t k = 0; k < meta[i].l_seq; ++k) ++geo[256 + (R.quals[meta[i].seq_off + k] & 127)];
    }
  }
  ExpandArgs a;
  memset(static_cast<void*>(&a), 0, sizeof a);
  a.meta = meta.data(); a.pos = R.pos.data(); a.tid = R.tid.data(); a.bases = R.bases.data(); a.quals = R.quals.data(); a.cigars = R.cigars.data();
  a.n_reads = raw.n; a.seg = plan.segs.data(); a.n_seg = (uint32_t)plan.segs.size(); a.n_tiles = plan.tiles; a.max_span = max_span.data();
  a.seg_of_tid = plan.seg_of_tid.data(); a.ref = plan.refbytes.data(); a.n_base = n_base;
  a.want_hist = cfg.want_hist; a.want_score = cfg.want_score; a.use_read_pos = cfg.use_read_pos; a.use_base_repeat = cfg.use_base_repeat;
  a.preprocess = cfg.preprocess_stage; a.unmatched_end_minimum_read_length = cfg.unmatched_end_minimum_read_length;
  a.unmatched_end_length_factor = cfg.unmatched_end_length_factor; a.stats = stats.data();
  std::vector<uint64_t> ins_mask(n_base + 1, 0);
  std::vector<uint8_t> sub_k(n_base + 1, 0);
  std::vector<uint32_t> sub_first(n_base + 2, 0);
  a.ins_mask = ins_mask.data(); a.sub_k = sub_k.data(); a.sub_first = sub_first.data();
  for (uint64_t i = 0; i < raw.n; ++i) ins_support(a, i);
  uint32_t n_ins = 0;
  for (uint32_t s = 0; s < n_base; ++s) {
    uint32_t K = 0;
    while (K < 63 && (ins_mask[s] >> K & 1)) ++K;
    sub_k[s] = (uint8_t)K; sub_first[s] = n_ins; n_ins += K;
  }
  D.n_ins = n_ins;
  const uint32_t n_slots = n_base + n_ins;
  D.max_read_set_seen = stats[XS_MAX_SET];
  if (cfg.want_score) D.geo = choose_geometry(geo.data(), geo.data() + 256, cfg, D.max_read_set_seen);
  a.geo = D.geo;
  D.hist_bytes = (cfg.use_base_repeat || cfg.use_read_pos || D.max_read_set_seen > 7) ? 8 : 4;
  a.hist_bytes = D.hist_bytes;
  std::vector<uint8_t> slot_ref(n_slots + 1, kBaseGap), slot_group(n_base + 1, 0);
  for (const ExpandSeg& sg : plan.segs)
    for (int32_t p = sg.lo; p < sg.hi; ++p) {
      const uint8_t b = xchar_to_index(plan.refbytes[sg.ref_off + (uint32_t)(p - sg.lo)]);
      slot_ref[sg.slot0 + (uint32_t)(p - sg.lo)] = b; slot_group[sg.slot0 + (uint32_t)(p - sg.lo)] = (uint8_t)sg.group;
    }
  for (uint32_t s = 0; s < n_base; ++s) for (uint32_t k = 1; k <= sub_k[s]; ++k) { D.ins_parent.push_back(s); D.ins_count.push_back(k); }
  a.slot_ref = slot_ref.data();
  std::vector<uint32_t> score_cnt(n_slots + 1, 0), red_cnt(n_slots + 1, 0), side_cnt(n_slots + 1, 0), side_red_cnt(n_slots + 1, 0), hist_cnt(n_base + 1, 0);
  std::vector<uint8_t> col_red(n_base + 1, 0), col_qstart(n_base + 1, 0);
  a.score_cnt = score_cnt.data(); a.red_cnt = red_cnt.data(); a.side_cnt = side_cnt.data(); a.side_red_cnt = side_red_cnt.data();
  a.hist_cnt = hist_cnt.data(); a.col_red = col_red.data(); a.col_qstart = col_qstart.data();
  for (uint32_t t = 0; t < plan.tiles; ++t) for (uint32_t l = 0; l < 32; ++l) tile_lane<false>(a, t, l);
  // offsets
  std::vector<uint32_t> side_off(n_slots + 2, 0);
  std::vector<uint64_t> hist_off(n_base + 2, 0), score_off(n_slots + 2, 0);
  uint64_t acc = 0;
  for (uint32_t s = 0; s < n_slots; ++s) { side_off[s] = (uint32_t)acc; acc += (side_cnt[s] + 1u) & ~1u; }
  side_off[n_slots] = (uint32_t)acc; D.n_side = acc;
  acc = 0;
  for (uint32_t s = 0; s < n_base; ++s) { hist_off[s] = acc | (col_red[s] ? HIST_OFF_REDUNDANT_BIT : 0); acc += hist_cnt[s]; D.max_hist_depth = std::max<uint64_t>(D.max_hist_depth, hist_cnt[s]); }
  hist_off[n_base] = acc; D.n_hist = acc;
  for (uint32_t s = 0; s < n_slots; ++s) D.n_score += score_cnt[s];
  // rounds: per block of 4096 slots, keys (group, vectors, slot) sorted, groups padded
  std::vector<uint32_t> round_slot, round_vecs;
  for (uint32_t b0 = 0; b0 < n_slots; b0 += ROUND_BLOCK) {
    std::vector<uint32_t> keys;
    for (uint32_t s = b0; s < std::min(n_slots, b0 + ROUND_BLOCK); ++s)
      keys.push_back((slot_ref[s] < 4 ? slot_ref[s] : 4u) << 28 | ((score_cnt[s] + 7u) >> 3) << 12 | (s - b0));
    std::sort(keys.begin(), keys.end());
    for (uint32_t g = 0; g < 5; ++g) {
      for (uint32_t k : keys) if (k >> 28 == g) round_slot.push_back(b0 + (k & 0xFFF));
      while (round_slot.size() & 31) round_slot.push_back(ROUND_NO_SLOT);
    }
  }
  D.n_rounds = round_slot.size() / 32;
  std::vector<uint64_t> round_off(D.n_rounds + 1, 0);
  std::vector<uint32_t> round_side(round_slot.size() * 2 + 2, 0);
  acc = 0;
  for (uint64_t r = 0; r < D.n_rounds; ++r) {
    round_off[r] = acc;
    uint32_t deepest = 0;
    for (uint32_t l = 0; l < 32; ++l) {
      const uint32_t sl = round_slot[r * 32 + l];
      if (sl == ROUND_NO_SLOT) { round_side[2 * (r * 32 + l)] = 5u << 29; continue; }
      score_off[sl] = acc + l * 4u;
      deepest = std::max(deepest, (score_cnt[sl] + 7u) >> 3);
      round_side[2 * (r * 32 + l)] = side_off[sl] | (uint32_t)slot_ref[sl] << 29;
      round_side[2 * (r * 32 + l) + 1] = side_off[sl] + (cfg.want_score ? side_cnt[sl] : 0u);
    }
    acc += (uint64_t)deepest * ROUND_VECTOR_WORDS;
  }
  round_off[D.n_rounds] = acc; score_off[n_slots] = acc; D.n_score_padded = acc;
  // fill
  std::vector<uint32_t> score_rec(D.n_score_padded + 1, D.geo.pad_word()), side_rec(D.n_side * D.geo.side_stride + 1, SIDE_PAD);
  std::vector<uint8_t> hist_rec(D.n_hist * D.hist_bytes + 8, 0);
  std::vector<uint32_t> sub_cur((size_t)n_ins * 4 + 4, 0);
  for (uint32_t j = 0; j < n_ins; ++j) { sub_cur[4 * j] = red_cnt[n_base + j]; sub_cur[4 * j + 2] = side_red_cnt[n_base + j]; }
  a.score_off = score_off.data(); a.side_off = side_off.data(); a.hist_off = hist_off.data();
  a.score_rec = score_rec.data(); a.side_rec = side_rec.data(); a.hist_rec = hist_rec.data(); a.sub_cur = sub_cur.data();
  for (uint32_t t = 0; t < plan.tiles; ++t) for (uint32_t l = 0; l < 32; ++l) tile_lane<true>(a, t, l);
  if (stats[XS_ERR]) { fprintf(stderr, "%s: device error word %x\n", c.name.c_str(), stats[XS_ERR]); return false; }
  uint32_t mapq_seen[8];
  for (int w = 0; w < 8; ++w) mapq_seen[w] = stats[XS_MAPQ_SEEN + w];
  if (cfg.want_score) mapq_seen[D.geo.hot_mapq >> 5] |= 1u << (D.geo.hot_mapq & 31);
  std::vector<uint64_t> starts;
  if (cfg.preprocess_stage && cfg.want_hist) {
    starts.assign(n_targets * 2, 0);
    for (uint32_t s = 0; s < n_base; ++s) {
      if (col_red[s]) continue;
      size_t v = 0;
      while (v + 1 < plan.segs.size() && plan.segs[v + 1].slot0 <= s) ++v;
      ++starts[(size_t)plan.segs[v].tid * 2 + (col_qstart[s] & 1)];
      ++starts[(size_t)plan.segs[v].tid * 2 + ((col_qstart[s] >> 1) & 1)];
    }
  }

  // ------------------------------------------------------------------ compare
  const std::string& cs = c.name;
  bool ok = true;
#define EQ(x) if (H.x != D.x) { fprintf(stderr, "%s: %s host %llu device %llu\n", cs.c_str(), #x, (unsigned long long)H.x, (unsigned long long)D.x); ok = false; }
  EQ(n_base) EQ(n_ins) EQ(n_rounds) EQ(n_score) EQ(n_hist) EQ(n_score_padded) EQ(n_side) EQ(hist_bytes) EQ(max_hist_depth) EQ(n_groups)
  EQ(max_read_set_seen) EQ(geo.cutoff) EQ(geo.hot_mapq) EQ(geo.q_lo) EQ(geo.n_q) EQ(geo.n_st) EQ(geo.side_stride)
  if (!ok) return false;
  if (H.max_qual_seen != stats[XS_MAX_Q] || H.max_hist_qual != stats[XS_MAX_HQ] || H.max_hist_rpos != stats[XS_MAX_RP] || H.max_score_rpos != stats[XS_MAX_SRP]) {
    fprintf(stderr, "%s: maxima differ: host %u %u %u %u device %u %u %u %u\n", cs.c_str(), H.max_qual_seen, H.max_hist_qual, H.max_hist_rpos, H.max_score_rpos,
            stats[XS_MAX_Q], stats[XS_MAX_HQ], stats[XS_MAX_RP], stats[XS_MAX_SRP]);
    ok = false;
  }
  ok = ok && same("mapq_seen", H.mapq_seen, mapq_seen, 8, cs);
  ok = ok && same("ins_parent", H.ins_parent.data(), D.ins_parent.data(), n_ins, cs) && same("ins_count", H.ins_count.data(), D.ins_count.data(), n_ins, cs);
  ok = ok && same("slot_ref", H.slot_ref, slot_ref.data(), n_slots, cs) && same("slot_group", H.slot_group, slot_group.data(), n_base, cs);
  ok = ok && same("score_cnt", H.score_cnt, score_cnt.data(), n_slots, cs) && same("side_off", H.side_off, side_off.data(), n_slots + 1, cs);
  ok = ok && same("hist_off", H.hist_off, hist_off.data(), n_base + 1, cs) && same("round_slot", H.round_slot, round_slot.data(), round_slot.size(), cs);
  ok = ok && same("round_off", H.round_off, round_off.data(), D.n_rounds + 1, cs) && same("score_off", H.score_off, score_off.data(), n_slots + 1, cs);
  ok = ok && same("round_side", H.round_side, round_side.data(), round_slot.size() * 2, cs);
  ok = ok && same("score_rec", H.score_rec, score_rec.data(), D.n_score_padded, cs);
  ok = ok && same("side_rec", H.side_rec, side_rec.data(), D.n_side * D.geo.side_stride, cs);
  ok = ok && same("hist_rec", static_cast<const uint8_t*>(H.hist_rec), hist_rec.data(), D.n_hist * D.hist_bytes, cs);
  if (ok && H.read_start_counts != starts) { fprintf(stderr, "%s: read_start_counts differ\n", cs.c_str()); ok = false; }
  if (ok && H.hist16) {  // the stable partition the device's compaction kernels restate
    std::vector<uint16_t> h16; std::vector<uint32_t> exc;
    const uint32_t* h = reinterpret_cast<const uint32_t*>(hist_rec.data());
    for (uint64_t i = 0; i < D.n_hist; ++i) { const uint32_t r = hist16_pack(h[i]); if (r < 0x10000u) h16.push_back((uint16_t)r); else exc.push_back(h[i]); }
    ok = h16.size() == H.n_hist16 && exc.size() == H.n_hist_exc && same("hist16", H.hist16, h16.data(), h16.size(), cs) && same("hist_exc", H.hist_exc, exc.data(), exc.size(), cs);
  }
  printf("%-28s %s  (%u slots, %llu records, %u sub-column slots, %llu side entries, %llu rounds)\n", cs.c_str(), ok ? "equal" : "DIFFERENT", n_slots,
         (unsigned long long)D.n_score, n_ins, (unsigned long long)D.n_side, (unsigned long long)D.n_rounds);
  free_stream(H, cfg);
  return ok;
}

}  // namespace

int main(int argc, char** argv) {
  std::vector<Case> cases = {
      {"lambda", 1, {48502}, {rs("lambda_reads", false, 35, 107.0)}, 60, 10, 2, false, false, false, 0, 1},
      {"multi", 7, {6000, 3500, 5000}, {rs("pe", true, 150, 60.0), rs("se", false, 36, 30.0)}, 25, 8, 2, false, false, false, 0, 1},
      {"tiny", 11, {1200, 800}, {rs("tp", true, 100, 30.0, 250, 25), rs("ts", false, 36, 15.0)}, 8, 4, 1, false, false, false, 0, 1},
      {"deep", 5, {700}, {rs("amp", false, 100, 1400.0)}, 10, 3, 1, false, false, false, 0, 1},
      {"ltee (read_pos, repeat)", 13, {3000}, {rs("s36", false, 36, 35.0), rs("p50", true, 50, 50.0, 160, 15)}, 12, 4, 1, true, true, false, 0, 1},
      {"multi shard 1/3", 7, {6000, 3500, 5000}, {rs("pe", true, 150, 60.0), rs("se", false, 36, 30.0)}, 25, 8, 2, false, false, false, 1, 3},
      {"multi shard 2/3", 7, {6000, 3500, 5000}, {rs("pe", true, 150, 60.0), rs("se", false, 36, 30.0)}, 25, 8, 2, false, false, false, 2, 3},
      {"tiny preprocess", 11, {1200, 800}, {rs("tp", true, 100, 30.0, 250, 25), rs("ts", false, 36, 15.0)}, 8, 4, 1, false, false, true, 0, 1},
      {"five read files", 17, {2500}, {rs("a", true, 75, 20.0, 200, 20), rs("b", true, 75, 20.0, 200, 20), rs("c", false, 50, 20.0)}, 10, 3, 1, false, false, false, 0, 1},
  };
  {
    Case hist_only = cases[1]; hist_only.name = "multi, histogram only"; hist_only.want_score = false; cases.push_back(hist_only);
    Case score_only = cases[1]; score_only.name = "multi, scoring only"; score_only.want_hist = false; cases.push_back(score_only);
  }
  if ((argc == 5 || argc == 6) && std::string(argv[1]) == "--bam") {   // one read set (single-end, or "paired"), reads from a BAM on disk
    Case external = cases[0];
    external.name = "external BAM";
    external.sets = {rs(argv[4], argc == 6, 0, 0.0)};
    external.bam = argv[2];
    external.fasta = argv[3];
    return run_case(external) ? 0 : 1;
  }
  bool ok = true;
  for (const Case& c : cases) {
    bool wanted = argc < 2;
    for (int i = 1; i < argc; ++i) if (c.name.find(argv[i]) != std::string::npos) wanted = true;
    if (wanted) ok = run_case(c) && ok;
  }
  return ok ? 0 : 1;
}
