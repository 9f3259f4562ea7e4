// Device-side staging ("expander"): aligned reads in HBM -> the columnar, reference-position-sorted streams the kernels read
// (brq_types.h), built by kernels instead of by the host's staging.cpp.  Replaces, on the device, the pileup engine and the
// per-record accessor calls of the reference: /root/reference/src/breseq/pileup_base.cpp:239-359 (column construction),
// alignment.cpp:104-288, 371-390, alignment.h:354-410 (query bounds, trims, base repeats), error_count.cpp:854-986 and
// 1049-1105 (which base / quality / neighbour a record contributes), identify_mutations.cpp:1359-1391, 1557-1657 (insert
// sub-columns, what counts as coverage).
//
// This header holds the per-read and per-(read, column) logic as host+device inline functions: expand.cu wraps them in
// kernels, tests/expand_check.cpp runs the same functions serially on the CPU against staging.cpp (a build without a GPU can
// still check the semantics; the product never runs them on the host).  The streams they build are bit-identical to
// staging.cpp's: tests compare every array.
//
// Shape of the work: a warp owns a TILE of 32 consecutive columns, one column per lane, and walks the candidate reads of
// the tile in BAM order (the reads are coordinate sorted: a binary search bounds them).  Every lane sees the reads that
// cover its column in arrival order, so a record's rank inside its slot is a per-lane counter: no atomics, no sort, and
// the order-dependent parts of the reference (the 1/X1 sums, the EM's summation order) keep their order.
#pragma once
#include "brq_types.h"

#include <cmath>

#ifdef __CUDA_ARCH__
#define BRQ_AOR64(p, v) atomicOr(reinterpret_cast<unsigned long long*>(p), (unsigned long long)(v))
#define BRQ_AOR32(p, v) atomicOr((p), (v))
#define BRQ_AMAX32(p, v) atomicMax((p), (v))
#define BRQ_AMAXI32(p, v) atomicMax((p), (v))
#define BRQ_AADD64(p, v) atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)(v))
#else
#define BRQ_AOR64(p, v) (*(p) |= (v))
#define BRQ_AOR32(p, v) (*(p) |= (v))
#define BRQ_AMAX32(p, v) (*(p) = *(p) < (v) ? (v) : *(p))
#define BRQ_AMAXI32(p, v) (*(p) = *(p) < (v) ? (v) : *(p))
#define BRQ_AADD64(p, v) (*(p) += (v))
#endif

namespace brq {

// What the expander reports instead of throwing (the host turns the lowest set bit into staging.cpp's message).
enum : uint32_t {
  EXP_ERR_UNSORTED = 1u << 0, EXP_ERR_READSET32 = 1u << 1, EXP_ERR_REFCHAR = 1u << 2, EXP_ERR_INS63 = 1u << 3,
  EXP_ERR_CIGAR_LONG = 1u << 4, EXP_ERR_X1ZERO = 1u << 5, EXP_ERR_QUAL127 = 1u << 6, EXP_ERR_DEL_NO_BASE = 1u << 7,
  EXP_ERR_RPOS = 1u << 8, EXP_ERR_NOQUAL_DOTN = 1u << 9, EXP_ERR_NOQUAL = 1u << 10, EXP_ERR_DEPTH = 1u << 11,
};
// words of the expander's statistics block (device, zeroed per staging call)
enum : uint32_t { XS_ERR = 0, XS_MAX_Q = 1, XS_MAX_HQ = 2, XS_MAX_RP = 3, XS_MAX_SRP = 4, XS_MAX_SET = 5, XS_MAX_HIST_DEPTH = 6, XS_SPARE = 7,
                  XS_MAPQ_SEEN = 8 /* 8 words */, XS_WORDS = 16 };

// One aligned read as the tile kernels see it: 64 bytes, read with four 128-bit loads that are uniform across the warp.
struct alignas(16) ReadMeta {
  int32_t pos, end;          // reference span [pos, end), end = pos + max(reference length, 1)
  uint32_t l_seq, n_cigar;
  uint64_t seq_off;          // into bases / quals
  uint64_t cigar_off;        // into cigars
  uint32_t x1;               // X1:i redundancy, 1 when absent
  int32_t xl, xr;            // XL / XR:i trims, -1 when absent
  int32_t qs0, qe0;          // first / last non-soft-clipped query index (alignment.cpp:248-288)
  int32_t qb_end0, qb_start0;  // query_bounds_0 (alignment.cpp:104-218, min_qual == 0)
  uint8_t mapq, read_set, flags, rg;   // rg: the read group's index among the header's @RG lines (0 when there is none)
};
static_assert(sizeof(ReadMeta) == 64, "ReadMeta is read as four 128-bit words");
constexpr uint8_t RM_LIVE = 1, RM_REV = 2, RM_HAS_INS = 4, RM_SIMPLE = 8;  // SIMPLE: one M / = / X run between clips: no CIGAR walk per column
// flags the pileup engine never shows a callback.  htslib 1.x sam.c, bam_plp_push(): "Skip only unmapped reads here, any
// additional filtering must be done in iter->func" -- the BAM_DEF_MASK that bam_plp_init() stores in flag_mask is no longer
// applied at push (samtools mpileup filters SECONDARY / QCFAIL / DUP in its own read function; breseq's read functions,
// pileup_base.cpp:225-236 and :290-301, filter nothing).  So only BAM_FUNMAP (and tid < 0) drops a record here, in
// staging.cpp, in the oracle and in oracle/hts_shim; tests/test_pileup_semantics.py feeds flagged reads to all of them.
constexpr uint32_t PILEUP_FLAG_MASK = 4u;

// the raw per-read arrays (what the host uploads), all in BAM order
struct RawReads {
  const int32_t* tid; const int32_t* pos; const uint16_t* flag; const uint8_t* mapq; const uint8_t* rg;
  const uint32_t* x1; const int32_t* xl; const int32_t* xr; const uint32_t* l_seq; const uint64_t* seq_off;
  const uint32_t* n_cigar; const uint64_t* cigar_off;
  const uint8_t* bases; const uint8_t* quals; const uint32_t* cigars;
  uint64_t n;
};

// a visited target clipped to this context's shard: columns [lo, hi) of BAM target tid are base slots slot0 ..
struct ExpandSeg {
  int32_t tid, lo, hi, tlen;
  uint32_t slot0;       // first base slot
  uint32_t tile0;       // first tile (32 columns each; tiles never straddle segments)
  uint32_t read_first, read_last;  // the reads of the target: [read_first, read_last) in BAM order
  uint32_t ref_off;     // into the reference bytes: columns lo .. hi (one byte past the range: the next base, or 0 at the target's end)
  uint32_t group;       // coverage group of the target
};

struct ExpandArgs {
  // reads
  const ReadMeta* meta; const int32_t* pos; const int32_t* tid; const uint8_t* bases; const uint8_t* quals; const uint32_t* cigars;
  uint64_t n_reads;
  // geometry
  const ExpandSeg* seg; uint32_t n_seg; uint32_t n_tiles;
  const int32_t* max_span;   // by BAM tid: the longest reference span of a read of the target
  const int32_t* seg_of_tid; // by BAM tid: index into seg, -1 = not visited
  const uint8_t* ref;        // reference characters as stored in the FASTA (see ExpandSeg::ref_off)
  uint32_t n_base;
  // options
  uint32_t want_hist, want_score, use_read_pos, use_base_repeat, preprocess;
  uint32_t unmatched_end_minimum_read_length; double unmatched_end_length_factor;
  ScoreGeometry geo; uint32_t hist_bytes;
  // per base slot
  uint8_t* slot_ref; uint64_t* ins_mask; const uint8_t* sub_k; const uint32_t* sub_first;
  // record counts per slot (count pass), base slots then insert sub-column slots
  uint32_t* score_cnt; uint32_t* red_cnt; uint32_t* side_cnt; uint32_t* side_red_cnt;
  uint32_t* hist_cnt; uint8_t* col_red; uint8_t* col_qstart;
  // fill pass
  const uint64_t* score_off; const uint32_t* side_off; const uint64_t* hist_off;
  uint32_t* score_rec; uint32_t* side_rec; void* hist_rec;
  uint32_t* sub_cur;  // [n_ins][4] cursors of the sub-column slots: unique, redundant, side (unique), side (redundant)
  uint32_t* stats;    // XS_* words
};

BRQ_HD inline bool xop_ref(uint32_t op) { return op == 0 || op == 2 || op == 3 || op == 7 || op == 8; }
BRQ_HD inline bool xop_match(uint32_t op) { return op == 0 || op == 7 || op == 8; }
// BAM 4-bit code -> base index (1, 2, 4, 8 = A, C, G, T; anything else is N): sixteen 4-bit entries in one constant
BRQ_HD inline uint8_t xnibble_to_index(uint8_t bam4) { return (uint8_t)((0x5555555355525105ull >> ((bam4 & 15u) * 4u)) & 15u); }
// FASTA character -> base index; 255 = not a base the reference accepts (NUL, the byte past a target's end, reads as kBaseNul)
BRQ_HD inline uint8_t xchar_to_index(uint8_t c) {
  return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : c == 'N' ? 5 : c == 0 ? (uint8_t)kBaseNul : 255;
}

// The indel htslib's pileup reports at the LAST column of reference-consuming operation k: next operation I => + the
// lengths of the consecutive I (P skipped); next operation D (and this one not D) => - the consecutive D lengths;
// P then I => the I lengths up to the next reference-consuming operation.
BRQ_HD inline int lookahead_indel(const uint32_t* cig, uint32_t n_cig, uint32_t k) {
  if (k + 1 >= n_cig) return 0;
  const uint32_t op = cig[k] & 0xf, op2 = cig[k + 1] & 0xf;
  const int32_t l2 = (int32_t)(cig[k + 1] >> 4);
  int indel = 0;
  if (op2 == 2 && op != 2) {
    indel = -l2;
    for (uint32_t j = k + 2; j < n_cig && (cig[j] & 0xf) == 2; ++j) indel -= (int32_t)(cig[j] >> 4);
  } else if (op2 == 1) {
    indel = l2;
    for (uint32_t j = k + 2; j < n_cig; ++j) {
      const uint32_t o = cig[j] & 0xf;
      if (o == 1) indel += (int32_t)(cig[j] >> 4);
      else if (o != 6) break;
    }
  } else if (op2 == 6 && k + 2 < n_cig) {
    int32_t l3 = 0;
    for (uint32_t j = k + 2; j < n_cig; ++j) {
      const uint32_t o = cig[j] & 0xf;
      if (o == 1) l3 += (int32_t)(cig[j] >> 4);
      else if (xop_ref(o)) break;
    }
    if (l3 > 0) indel = l3;
  }
  return indel;
}

// What the pileup engine reports for one read at one column it spans (pos <= c < end): the query position (on a deleted
// column: the first query base after the deletion), is_del, and the indel of the column.
struct ColumnHit { int32_t q; int indel; bool has, is_del; };
BRQ_HD inline ColumnHit column_hit(const uint32_t* cig, uint32_t n_cig, int32_t pos, int32_t c) {
  ColumnHit h{0, 0, false, false};
  int32_t x = pos, y = 0;
  for (uint32_t k = 0; k < n_cig; ++k) {
    const uint32_t op = cig[k] & 0xf;
    const int32_t l = (int32_t)(cig[k] >> 4);
    if (xop_ref(op)) {
      const int32_t xe = x + l;
      if (c < xe) {
        h.has = true; h.is_del = !xop_match(op);
        h.q = h.is_del ? y : y + (c - x);
        h.indel = c == xe - 1 ? lookahead_indel(cig, n_cig, k) : 0;
        return h;
      }
      if (xop_match(op)) y += l;
      x = xe;
    } else if (op == 1 || op == 4) {
      y += l;
    }
  }
  return h;
}

// ---- per read: the derived values every (read, column) visit needs (staging.cpp "per-read derived values")
// part_base / part_count: flat read-file index of every read group (alignment.cpp:565-605), n_part entries (0 = none)
BRQ_HD inline ReadMeta prep_read(const RawReads& R, uint64_t i, const uint32_t* part_base, const uint32_t* part_count, uint32_t n_part,
                                 int32_t* max_span, uint32_t* stats) {
  ReadMeta m;
  const uint32_t* cig = R.cigars + R.cigar_off[i];
  const uint32_t nc = R.n_cigar[i];
  m.pos = R.pos[i]; m.l_seq = R.l_seq[i]; m.n_cigar = nc; m.seq_off = R.seq_off[i]; m.cigar_off = R.cigar_off[i];
  m.x1 = R.x1[i]; m.xl = R.xl[i]; m.xr = R.xr[i]; m.mapq = R.mapq[i]; m.rg = R.rg[i];
  int32_t rlen = 0, qlen = 0;
  bool has_ins = false;
  uint32_t n_match_ops = 0, n_other_ops = 0;   // other: anything but M / = / X and the clips S, H
  for (uint32_t k = 0; k < nc; ++k) {
    const uint32_t op = cig[k] & 0xf; const int32_t l = (int32_t)(cig[k] >> 4);
    if (xop_ref(op)) rlen += l;
    if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) qlen += l;
    if (op == 1) has_ins = true;
    if (xop_match(op)) ++n_match_ops; else if (op != 4 && op != 5) ++n_other_ops;
  }
  m.end = m.pos + (rlen ? rlen : 1);
  int32_t qs1 = 1;
  for (uint32_t k = 0; k < nc && (cig[k] & 0xf) == 4; ++k) qs1 += (int32_t)(cig[k] >> 4);
  int32_t qe1 = qlen;
  for (uint32_t k = nc; k-- > 1 && (cig[k] & 0xf) == 4;) qe1 -= (int32_t)(cig[k] >> 4);
  m.qs0 = qs1 - 1; m.qe0 = qe1 - 1;
  int32_t be1 = qlen;
  for (uint32_t k = nc; k-- > 1;) {
    const uint32_t op = cig[k] & 0xf;
    if (op != 4 && op != 5 && op != 3) break;
    if (op == 4) be1 -= (int32_t)(cig[k] >> 4);
  }
  m.qb_end0 = be1 - 1;
  int32_t bs1 = 1;
  for (uint32_t k = 0; k < nc; ++k) {
    const uint32_t op = cig[k] & 0xf;
    if (op != 4 && op != 5 && op != 3) break;
    if (op == 4) bs1 += (int32_t)(cig[k] >> 4);
  }
  m.qb_start0 = bs1 - 1;
  const uint32_t g = R.rg[i];
  uint32_t set = 0;
  if (n_part && g < n_part) set = part_base[g] + (((R.flag[i] & 128) && part_count[g] > 1) ? 1u : 0u);
  if (set >= 32) { BRQ_AOR32(&stats[XS_ERR], EXP_ERR_READSET32); set = 31; }
  m.read_set = (uint8_t)set;
  BRQ_AMAX32(&stats[XS_MAX_SET], set);
  const int32_t tid = R.tid[i];
  if (tid >= 0 && rlen > 1) BRQ_AMAXI32(&max_span[tid], rlen);
  const bool live = tid >= 0 && !(R.flag[i] & PILEUP_FLAG_MASK) && nc > 0;
  m.flags = (uint8_t)((live ? RM_LIVE : 0) | ((R.flag[i] & 16) ? RM_REV : 0) | (has_ins ? RM_HAS_INS : 0) |
                      (n_match_ops == 1 && n_other_ops == 0 ? RM_SIMPLE : 0));
  // the BAM must be coordinate sorted (htslib's pileup aborts otherwise); reads without a target sort last
  if (i > 0) {
    const int32_t pt = R.tid[i - 1];
    const uint32_t a = pt < 0 ? 0x7FFFFFFFu : (uint32_t)pt, b = tid < 0 ? 0x7FFFFFFFu : (uint32_t)tid;
    if (b < a || (b == a && tid >= 0 && R.pos[i] < R.pos[i - 1])) BRQ_AOR32(&stats[XS_ERR], EXP_ERR_UNSORTED);
  }
  return m;
}

// ---- per read with an insertion: which insert sub-columns it supports (staging.cpp pass A1).  Level k + 1 of a column
// exists iff a UNIQUE read has an insertion longer than k after it and a non-N base at level k
// (identify_mutations.cpp:1577 precedes :1598): bit k of the column's mask.
BRQ_HD inline void ins_support(const ExpandArgs& a, uint64_t i) {
  const ReadMeta& m = a.meta[i];
  if ((m.flags & (RM_LIVE | RM_HAS_INS)) != (RM_LIVE | RM_HAS_INS) || m.x1 != 1) return;
  const uint32_t* cig = a.cigars + m.cigar_off;
  const uint8_t* seq = a.bases + m.seq_off;
  int32_t x = m.pos, y = 0;
  const int32_t v = a.seg_of_tid[a.tid[i]];   // (a live read has a target)
  if (v < 0) return;
  const ExpandSeg* sg = &a.seg[v];
  for (uint32_t k = 0; k < m.n_cigar; ++k) {
    const uint32_t op = cig[k] & 0xf;
    const int32_t l = (int32_t)(cig[k] >> 4);
    if (xop_ref(op)) {
      const int32_t xe = x + l, c = xe - 1;
      if (l > 0 && xop_match(op) && c >= sg->lo && c < sg->hi) {
        const int indel = lookahead_indel(cig, m.n_cigar, k);
        if (indel > 0) {
          if (indel > 63) BRQ_AOR32(&a.stats[XS_ERR], EXP_ERR_INS63);
          else {
            const int32_t q = y + l - 1;
            uint64_t mask = 0;
            for (int j = 0; j < indel; ++j) if ((uint32_t)(q + j) < m.l_seq && seq[q + j] != 15) mask |= 1ull << j;
            BRQ_AOR64(&a.ins_mask[sg->slot0 + (uint32_t)(c - sg->lo)], mask);
          }
        }
      }
      if (xop_match(op)) y += l;
      x = xe;
    } else if (op == 1 || op == 4) {
      y += l;
    }
  }
}

// ---- the lane of a tile: one column and what it accumulates
struct LaneState {
  uint32_t slot, K, sub0, ref;     // base slot, its sub-columns (count, first sub-column slot), reference base index
  uint32_t ref_next;               // base index of the NEXT reference column (kBaseNul past the target's end; 255 = not a base)
  int32_t col;                      // column inside the target
  // count pass
  uint32_t n_score, n_red, n_side, n_side_red, n_hist, red_flag, qstart;
  // fill pass
  uint64_t score_off, hist_at;
  uint32_t cur_u, cur_r, side_u, side_r;   // next record of each kind (records: index inside the slot; side: absolute entry)
  // statistics
  uint32_t max_q, max_hq, max_rp, max_srp;
};

struct DevWordX { uint32_t dev, side; bool has_side; };
// device stream word (and side-list entry, if any) of a classic word in a slot with reference base `ref` (brq_types.h)
BRQ_HD inline DevWordX encode_word(const ScoreGeometry& geo, uint32_t rec, uint32_t x1, uint32_t ref) {
  DevWordX w{0, 0, false};
  const bool is_top = (rec & SR_TOP_BIT) != 0;
  const uint32_t top = is_top ? DR_TOP_BIT : 0u;
  const uint32_t n_sq = geo.n_st * geo.n_q;
  auto special = [&](uint32_t which) { const uint32_t idx = n_sq + which; return (idx >> 2) * 128u + (idx & 3u) * 8u; };
  if (!(rec & SR_UNIQUE_BIT)) {
    w.dev = DR_REDUNDANT | top | (x1 < DR_X1_MASK ? x1 : DR_X1_MASK) << DR_X1_SHIFT | special(SC_TRASH) | (rec & 7u) << DR_RED_OBS_SHIFT |
            ((rec & SR_TRIM_BIT) ? DR_RED_TRIM_BIT : 0u);
    if (x1 >= DR_X1_MASK) { w.has_side = true; w.side = SIDE_BIG | x1; }
    return w;
  }
  const uint32_t qv = (rec >> SR_QUAL_SHIFT) & 127u, obs = rec & 7u;
  if ((rec & SR_TRIM_BIT) || !(rec & SR_OK_BIT) || qv < geo.cutoff) {
    w.dev = DR_IDLE | top | special(is_top ? SC_IDLE_TOP : SC_IDLE_BOT);
    return w;
  }
  const bool match = obs == ref;
  // HOT: a record of the shared table's classes that MATCHES the reference base.  One that does not (a sequencing error or a
  // variant: one in a thousand) is COLD like a record of another MAPQ: its terms come from its side-list entry
  if (match && n_sq && ((rec >> SR_MAPQ_SHIFT) & 255u) == geo.hot_mapq && obs < 5 && qv >= geo.q_lo && qv - geo.q_lo < geo.n_q) {
    const uint32_t sq = ((rec >> 10) & 63u) * geo.n_q + (qv - geo.q_lo);
    w.dev = sq << DR_SQ_SHIFT | obs << DR_OBS_SHIFT | top | DR_MATCH_BIT | ((sq >> 2) * 128u + (sq & 3u) * 8u);
  } else {
    w.dev = DR_COLD | top | special(is_top ? SC_COLD_TOP : SC_COLD_BOT);
    w.has_side = true; w.side = rec | (match ? SR_MATCH_BIT : 0u);
  }
  return w;
}

BRQ_HD inline uint32_t base_repeat_of(const uint8_t* seq, int32_t qp, bool rev, int32_t qe0) {  // alignment.cpp:371-390
  const uint8_t b = seq[qp];
  uint32_t rep = 0;
  if (!rev) { while (qp < qe0) { ++qp; if (seq[qp] != b) break; ++rep; } }
  else { while (qp > 0) { --qp; if (seq[qp] != b) break; ++rep; } }
  return rep;
}

// One pileup entry: read `m` at the lane's column (hit h).  FILL = false counts the records the entry makes (staging.cpp pass A2),
// FILL = true writes them (pass B).  sg: the lane's segment.
template <bool FILL>
BRQ_HD inline void visit_entry(const ExpandArgs& a, const ExpandSeg& sg, const ReadMeta& m, const ColumnHit& h, LaneState& st) {
  const uint8_t* seq = a.bases + m.seq_off;
  const uint8_t* qual = a.quals + m.seq_off;
  const bool unique = m.x1 == 1;
  const bool revb = (m.flags & RM_REV) != 0;
  const uint32_t rev = revb ? 1u : 0u;
  const int32_t L = (int32_t)m.l_seq, q = h.q, c = st.col;
  const bool is_del = h.is_del;
  const int indel = h.indel;
  uint32_t* err = &a.stats[XS_ERR];
  if ((uint32_t)q >= m.l_seq && !is_del) { BRQ_AOR32(err, EXP_ERR_CIGAR_LONG); return; }
  if (m.x1 == 0) { BRQ_AOR32(err, EXP_ERR_X1ZERO); return; }

  // ---------------- error_count record (error_count.cpp:125-199, 854-986)
  if (a.want_hist && !is_del) {
    if (!unique) st.red_flag = 1;
    else if (!FILL) ++st.n_hist;
    else {
      uint64_t rec = 0;
      const uint32_t qa = qual[q];
      if (qa > 127) BRQ_AOR32(err, EXP_ERR_QUAL127);
      const uint32_t obsA = xnibble_to_index(seq[q]), refA = st.ref;
      auto strand = [&](uint32_t b) { return rev ? 3u - b : b; };  // complement on the read strand (A, C, G, T only)
      if (obsA < 4 && refA < 4) {
        rec |= (uint64_t)strand(refA) << HR_REFA | (uint64_t)strand(obsA) << HR_OBSA | (uint64_t)(qa & 127u) << HR_QUALA | 1ull << HR_VALIDA;
        if (qa > st.max_hq) st.max_hq = qa;
      }
      rec |= (uint64_t)(m.read_set & 7u) << HR_SET | (uint64_t)(m.read_set >> 3) << HR_SET_HI;
      if (q > 65535) BRQ_AOR32(err, EXP_ERR_RPOS);
      rec |= (uint64_t)(q & 0xFFFF) << HR_RPOS;
      if ((uint32_t)q > st.max_rp) st.max_rp = (uint32_t)q;
      if (a.use_base_repeat) { const uint32_t rp = base_repeat_of(seq, q, revb, m.qe0); rec |= (uint64_t)(rp < 255u ? rp : 255u) << HR_REPA; }
      uint32_t cls = 0; int32_t mq = -1; uint32_t refb = kBaseNul;
      auto ref_index = [&](int32_t p) -> uint32_t {  // forward-strand reference base of this column or the next (per-lane constants)
        if (p == c) return st.ref;
        if (st.ref_next == 255u) { BRQ_AOR32(err, EXP_ERR_REFCHAR); return kBaseN; }
        return st.ref_next;
      };
      bool dead = false;
      if (indel == 0) {
        if (q < m.qe0) { cls = 1; mq = q + 1 - (int32_t)rev; refb = ref_index(c + 1 - (int32_t)rev); }
      } else if (indel == -1) {
        cls = 2; mq = q + 1 - (int32_t)rev; refb = ref_index(c + 1);
      } else if (indel == 1) {
        mq = q + 1;
        if (mq >= L) { BRQ_AOR32(err, EXP_ERR_NOQUAL_DOTN); dead = true; }
        else if (mq <= m.qe0 && mq >= m.qs0) cls = 3;
      }
      if (cls && !dead) {
        if (mq < 0 || mq >= L) BRQ_AOR32(err, EXP_ERR_NOQUAL);
        else {
          const uint32_t qb = qual[mq];
          if (qb > 127) BRQ_AOR32(err, EXP_ERR_QUAL127);
          const uint32_t obsB = xnibble_to_index(seq[mq]);
          uint32_t fB = kBaseGap, oB = kBaseGap;
          bool valid = obsB != kBaseN;
          if (cls == 1) valid = valid && refb != kBaseN;
          else if (cls == 2) { valid = valid && refb < 4; fB = valid ? strand(refb) : (uint32_t)kBaseGap; }
          else oB = valid ? strand(obsB) : (uint32_t)kBaseGap;
          if (valid) {
            rec |= (uint64_t)fB << HR_REFB | (uint64_t)oB << HR_OBSB | (uint64_t)(qb & 127u) << HR_QUALB | 1ull << HR_VALIDB;
            if (qb > st.max_hq) st.max_hq = qb;
            if (a.use_base_repeat) { const uint32_t rp = base_repeat_of(seq, mq, revb, m.qe0); rec |= (uint64_t)(rp < 31u ? rp : 31u) << HR_REPB; }
          }
        }
      }
      {
        const bool a_match = (rec >> HR_VALIDA & 1) && ((rec >> HR_REFA & 7) == (rec >> HR_OBSA & 7));
        const bool b_valid = rec >> HR_VALIDB & 1;
        const bool b_dots = b_valid && (rec >> HR_REFB & 7) == kBaseGap && (rec >> HR_OBSB & 7) == kBaseGap;
        if (a_match && (b_dots || !b_valid)) {
          rec |= 1ull << HR_FAST;
          if (!b_valid) rec |= 127ull << HR_QUALB;
        }
      }
      if (a.hist_bytes == 8) static_cast<uint64_t*>(a.hist_rec)[st.hist_at++] = rec;
      else static_cast<uint32_t*>(a.hist_rec)[st.hist_at++] = (uint32_t)rec;
    }
  }
  // preprocess stage: a unique read starts here inside the junction read-end bound (error_count.cpp:157-166)
  if (!FILL && a.preprocess && a.want_hist && !is_del && unique && q == 0) {
    const uint32_t Lu = m.l_seq;
    const uint32_t stranded_end_1 = revb ? Lu - (uint32_t)(m.qb_start0 + 1) + 1 : (uint32_t)m.qb_end0 + 1;
    const int32_t max_len = (int32_t)floor((double)((int32_t)Lu - (int32_t)a.unmatched_end_minimum_read_length) * a.unmatched_end_length_factor);
    const uint32_t end_min = max_len <= 0 ? Lu : Lu - (uint32_t)max_len;
    if (stranded_end_1 >= end_min) st.qstart |= revb ? 2u : 1u;
  }

  // ---------------- identify_mutations records: the column itself (k = 0) and its insert sub-columns
  // (identify_mutations.cpp:1561-1657, error_count.cpp:1049-1105)
  if (!a.want_score) return;
  const int ind = is_del ? -1 : (indel > 0 ? indel : 0);
  const uint32_t q1 = (uint32_t)q + 1;
  for (uint32_t k = 0; k <= st.K; ++k) {
    const bool past_base = !(ind >= (int)k);
    const uint8_t obs = past_base ? (uint8_t)kBaseGap : xnibble_to_index(seq[q + (int32_t)k]);
    if (obs == kBaseN) continue;  // not even coverage
    uint32_t rec = obs;
    if (!rev) rec |= SR_TOP_BIT;
    bool trimmed = false;  // alignment.h:389-410 (unsigned comparisons as there)
    if (m.xl >= 0 || m.xl < -1) { if (q1 <= (uint32_t)m.xl) trimmed = true; }
    if (m.xr >= 0 || m.xr < -1) {
      if ((uint32_t)L - q1 + 1 <= (uint32_t)m.xr) trimmed = true;
      if (past_base && ((uint32_t)L - q1 == (uint32_t)m.xr)) trimmed = true;
    }
    if (trimmed) rec |= SR_TRIM_BIT;
    uint32_t ext = 0;
    if (unique) {
      rec |= SR_UNIQUE_BIT;
      int32_t qp = q;
      bool ok = true;
      if (ind == -1) {
        qp += 1 - (int32_t)rev;
        if (qp >= L) { BRQ_AOR32(err, EXP_ERR_DEL_NO_BASE); return; }
        if (seq[qp] == 15) ok = false;
      } else if (k > 0) {
        qp += ((int)k < ind ? (int)k : ind) + 1 - (int32_t)rev;
        if (qp > m.qb_end0) ok = false;
        else if (seq[qp] == 15) ok = false;
      }
      if (ok) {
        const uint32_t qv = qual[qp];
        if (qv > 127) { BRQ_AOR32(err, EXP_ERR_QUAL127); return; }
        rec |= SR_OK_BIT | (qv << SR_QUAL_SHIFT);
        if (a.geo.side_stride == 2) {
          if (qp > 65535) { BRQ_AOR32(err, EXP_ERR_RPOS); return; }
          ext = (uint32_t)qp;
          if (a.use_base_repeat) { const uint32_t rp = base_repeat_of(seq, qp, revb, m.qe0); ext |= (rp < 255u ? rp : 255u) << 16; }
        }
      }
      rec |= (uint32_t)m.mapq << SR_MAPQ_SHIFT;
      rec |= (uint32_t)m.read_set << SR_SET_SHIFT;
    } else {
      rec |= (m.x1 < SR_RED_MASK ? m.x1 : SR_RED_MASK) << SR_RED_SHIFT;
    }
    const uint32_t s = k == 0 ? st.slot : st.sub0 + k - 1;
    const DevWordX w = encode_word(a.geo, rec, m.x1, k == 0 ? st.ref : (uint32_t)kBaseGap);
    if (!FILL) {
      if (k == 0) {
        ++st.n_score;
        if (!unique) ++st.n_red;
        if (w.has_side) { ++st.n_side; if (!unique) ++st.n_side_red; }
      } else {  // a sub-column slot belongs to its parent's lane alone: plain read-modify-write
        ++a.score_cnt[s];
        if (!unique) ++a.red_cnt[s];
        if (w.has_side) { ++a.side_cnt[s]; if (!unique) ++a.side_red_cnt[s]; }
      }
    } else {
      uint32_t at, side_at = 0;
      uint64_t off;
      if (k == 0) {
        at = unique ? st.cur_u++ : st.cur_r++;
        off = st.score_off;
        if (w.has_side) side_at = unique ? st.side_u++ : st.side_r++;
      } else {
        uint32_t* cur = a.sub_cur + (size_t)(s - a.n_base) * 4;
        at = unique ? cur[0]++ : cur[1]++;
        off = a.score_off[s];
        if (w.has_side) side_at = a.side_off[s] + (unique ? cur[2]++ : cur[3]++);
      }
      a.score_rec[score_index(off, at)] = w.dev;
      if (w.has_side) {
        const size_t e = (size_t)side_at * a.geo.side_stride;
        a.side_rec[e] = w.side;
        if (a.geo.side_stride == 2) { a.side_rec[e + 1] = ext; if (!(w.side & SIDE_BIG) && (ext & 0xFFFFu) > st.max_srp) st.max_srp = ext & 0xFFFFu; }
      }
      const uint32_t kind = w.dev >> DR_KIND_SHIFT;
      if (kind == 0 || kind == 2) {  // a scoring record: what the likelihood tables must cover
        const uint32_t qv = (rec >> SR_QUAL_SHIFT) & 127u;
        if (qv > st.max_q) st.max_q = qv;
        if (m.mapq != a.geo.hot_mapq) BRQ_AOR32(&a.stats[XS_MAPQ_SEEN + (m.mapq >> 5)], 1u << (m.mapq & 31));
      }
    }
  }
}

// the candidate reads of a tile [c0, c1) of segment sg: the reads of the target with pos + max_span > c0 and pos < c1
BRQ_HD inline void tile_candidates(const ExpandArgs& a, const ExpandSeg& sg, int32_t c0, int32_t c1, uint32_t& first, uint32_t& last) {
  const int32_t span = a.max_span[sg.tid];
  const int32_t lo_pos = c0 - span;  // first candidate: pos > c0 - span
  uint32_t lo = sg.read_first, hi = sg.read_last;
  while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (a.pos[mid] > lo_pos) hi = mid; else lo = mid + 1; }
  first = lo;
  hi = sg.read_last;
  while (lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if (a.pos[mid] >= c1) hi = mid; else lo = mid + 1; }
  last = lo;
}

BRQ_HD inline const ExpandSeg& seg_of_tile(const ExpandArgs& a, uint32_t tile) {
  uint32_t lo = 0, hi = a.n_seg;
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.seg[mid].tile0 <= tile) lo = mid; else hi = mid; }
  return a.seg[lo];
}

// One lane of one tile: lane `l` of tile `tile`.  FILL = false: the count pass; true: the fill pass.
template <bool FILL>
BRQ_HD inline void tile_lane(const ExpandArgs& a, uint32_t tile, uint32_t l) {
  const ExpandSeg& sg = seg_of_tile(a, tile);
  const int32_t c0 = sg.lo + (int32_t)((tile - sg.tile0) * 32u), c1 = c0 + 32 < sg.hi ? c0 + 32 : sg.hi;
  const int32_t c = c0 + (int32_t)l;
  uint32_t first, last;
  tile_candidates(a, sg, c0, c1, first, last);
  const bool live_lane = c < c1;
  LaneState st;
  st.col = c;
  st.slot = sg.slot0 + (uint32_t)(c - sg.lo);
  st.K = 0; st.sub0 = 0; st.ref = 5; st.ref_next = kBaseNul;
  st.n_score = st.n_red = st.n_side = st.n_side_red = st.n_hist = st.red_flag = st.qstart = 0;
  st.max_q = st.max_hq = st.max_rp = st.max_srp = 0;
  st.score_off = st.hist_at = 0; st.cur_u = st.cur_r = st.side_u = st.side_r = 0;
  if (live_lane) {
    st.K = a.want_score ? a.sub_k[st.slot] : 0u;
    st.sub0 = a.n_base + a.sub_first[st.slot];
    st.ref = a.slot_ref[st.slot];
    {  // the next column's base: the byte past the target's end is the FASTA buffer's NUL terminator
      const uint8_t b = c + 1 >= sg.tlen ? (uint8_t)kBaseNul : xchar_to_index(a.ref[sg.ref_off + (uint32_t)(c + 1 - sg.lo)]);
      st.ref_next = (b > 5 && b != kBaseNul) || b == 4 ? 255u : b;
    }
    if (FILL) {
      if (a.want_score) {
        st.score_off = a.score_off[st.slot];
        st.cur_u = a.red_cnt[st.slot]; st.cur_r = 0;
        st.side_r = a.side_off[st.slot]; st.side_u = st.side_r + a.side_red_cnt[st.slot];
      }
      if (a.want_hist) st.hist_at = a.hist_off[st.slot] & ~HIST_OFF_REDUNDANT_BIT;
    }
  }
  // the next read's 64 bytes are requested while this one is visited (the loop is a chain of dependent loads otherwise)
  ReadMeta nxt = first < last ? a.meta[first] : ReadMeta();
  for (uint32_t i = first; i < last; ++i) {
    const ReadMeta m = nxt;
    if (i + 1 < last) nxt = a.meta[i + 1];
    if (!(m.flags & RM_LIVE) || m.end <= c0) continue;   // (uniform across the tile)
    if (!live_lane || c < m.pos || c >= m.end) continue;
    ColumnHit h;
    if (m.flags & RM_SIMPLE) { h.has = true; h.is_del = false; h.indel = 0; h.q = m.qs0 + (c - m.pos); }   // one aligned run after the leading clip
    else h = column_hit(a.cigars + m.cigar_off, m.n_cigar, m.pos, c);
    if (!h.has) continue;
    visit_entry<FILL>(a, sg, m, h, st);
  }
  if (!live_lane) return;
  if (!FILL) {
    if (a.want_score) { a.score_cnt[st.slot] = st.n_score; a.red_cnt[st.slot] = st.n_red; a.side_cnt[st.slot] = st.n_side; a.side_red_cnt[st.slot] = st.n_side_red; }
    if (a.want_hist) { a.hist_cnt[st.slot] = st.n_hist; a.col_red[st.slot] = (uint8_t)st.red_flag; if (a.preprocess) a.col_qstart[st.slot] = (uint8_t)st.qstart; }
  } else {
    if (st.max_q) BRQ_AMAX32(&a.stats[XS_MAX_Q], st.max_q);
    if (st.max_hq) BRQ_AMAX32(&a.stats[XS_MAX_HQ], st.max_hq);
    if (st.max_rp) BRQ_AMAX32(&a.stats[XS_MAX_RP], st.max_rp);
    if (st.max_srp) BRQ_AMAX32(&a.stats[XS_MAX_SRP], st.max_srp);
  }
}

// ---- BAM2COV: the per-position coverage table of coverage_output::pileup_callback (coverage_output.cpp:307-470), from the same
// tile walk.  Per column, by strand (index 1 = reversed): unique reads covering it with an aligned base (a deletion or a
// reference skip over the column does not count), the reads among them whose first base it is, redundant reads (X1 > 1) as a
// count and as the sum of 1 / X1 in BAM order.
// `covered`: the pileup engine reports the column at all (any read spans it, deletions included): what the table's tail rule needs.
struct CoverageColumn { uint32_t unique[2], raw_redundant[2], begin[2], covered, pad; double redundant[2]; };
static_assert(sizeof(CoverageColumn) == 48, "CoverageColumn layout");

// group: only the reads of that read group count (the table's per-read-group column sets), COVERAGE_ALL_GROUPS: every read
constexpr uint32_t COVERAGE_ALL_GROUPS = 0xFFFFFFFFu;
// include_deleted: a deletion or reference skip over the column counts as coverage, as in pass 2's tallies (<seq>.coverage.tsv)
BRQ_HD inline void coverage_lane(const ExpandArgs& a, uint32_t tile, uint32_t l, CoverageColumn* out, uint32_t group = COVERAGE_ALL_GROUPS,
                                 bool include_deleted = false) {
  const ExpandSeg& sg = seg_of_tile(a, tile);
  const int32_t c0 = sg.lo + (int32_t)((tile - sg.tile0) * 32u), c1 = c0 + 32 < sg.hi ? c0 + 32 : sg.hi;
  const int32_t c = c0 + (int32_t)l;
  uint32_t first, last;
  tile_candidates(a, sg, c0, c1, first, last);
  const bool live_lane = c < c1;
  CoverageColumn col = {{0, 0}, {0, 0}, {0, 0}, 0, 0, {0.0, 0.0}};
  ReadMeta nxt = first < last ? a.meta[first] : ReadMeta();
  for (uint32_t i = first; i < last; ++i) {
    const ReadMeta m = nxt;
    if (i + 1 < last) nxt = a.meta[i + 1];
    if (!(m.flags & RM_LIVE) || m.end <= c0) continue;
    if (!live_lane || c < m.pos || c >= m.end) continue;
    ColumnHit h;
    if (m.flags & RM_SIMPLE) { h.has = true; h.is_del = false; h.indel = 0; h.q = m.qs0 + (c - m.pos); }
    else h = column_hit(a.cigars + m.cigar_off, m.n_cigar, m.pos, c);
    if (!h.has) continue;
    col.covered = 1;
    if ((h.is_del && !include_deleted) || (group != COVERAGE_ALL_GROUPS && m.rg != group)) continue;   // :370-373, :380
    // pass 2 does not count a read whose base here is N, not even as coverage (identify_mutations.cpp:1575-1576); bam2cov does
    if (include_deleted && !h.is_del && (uint32_t)h.q < m.l_seq && xnibble_to_index(a.bases[m.seq_off + (uint32_t)h.q]) > 3) continue;
    const uint32_t rev = (m.flags & RM_REV) ? 1u : 0u;
    if (m.x1 == 1) {
      ++col.unique[rev];
      // the read's first base in its own orientation (:385-393): query position 1, or the last one of a reversed read
      if (!h.is_del && (rev ? h.q == (int32_t)m.l_seq - 1 : h.q == 0)) ++col.begin[rev];
    } else {
      ++col.raw_redundant[rev];
      col.redundant[rev] += 1.0 / (double)m.x1;
    }
  }
  if (live_lane) out[sg.slot0 + (uint32_t)(c - sg.lo)] = col;
}

}  // namespace brq
