// Device side of the pileup: launch wrappers for the sm_100a kernels in kernels.cu.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace brq {

// Mixed-radix layout of the covariate table (/root/reference/src/breseq/error_count.cpp:477-497,
// 585-593): idx = sum over USED covariates of value * offset, offsets assigned in enum order.
struct CovLayout {
  uint32_t off_set, off_ref, off_obs, off_qual, off_rpos, off_rep;  // 0 when unused
  uint32_t max_set, max_qual, max_rpos, max_rep;                    // UINT32_MAX when unused (no bound)
  uint32_t n_bins;
  uint32_t obs_used;                                                // obs_base is a covariate
};

enum : uint32_t {
  BRQ_ERR_QUALITY_RANGE = 1,    // quality >= table maximum (reference: fatal ASSERT, error_count.cpp:487)
  BRQ_ERR_READSET_RANGE = 2,
  BRQ_ERR_READPOS_RANGE = 4,
  BRQ_ERR_CLASS_OVERFLOW = 8,   // more distinct record classes in one column than the class table holds
  BRQ_ERR_DEPTH_RANGE = 16,     // unique depth beyond the coverage histogram
  BRQ_ERR_PEER_TIMEOUT = 32,    // a peer's share of the histograms did not arrive (exchange.cu)
};

// Per-slot result of the scoring kernel: 96 bytes.
struct ColumnOut {
  double ll[5];            // sum over scoring records of log10 P(obs | true base b)
  double consensus_score;  // pure-genotype log-odds minus log10(total reference length); NaN when n == 0
  double variant_score;    // presence score of the top non-reference allele; NaN when there is none
  double redundant[2];     // [0] bottom strand, [1] top strand; sequential sum of 1/X1 in arrival order
  uint32_t unique[2];
  uint32_t raw_redundant[2];
  uint32_t n;              // scoring records (unique, untrimmed, resolvable, quality >= cutoff)
  uint32_t bits;           // [2:0] best [5:3] major [8:6] minor [11:9] variant (5 = N)
                           // [12] base_predicted [13] unique_only [14] emit candidate [15] needs host re-check
                           // [23:16] EM iterations of the full fit [24] the EM fit was evaluated (else: no scoring
                           // record, or the presence bound proves the column cannot emit an RA row; major, minor
                           // and variant read 5 and variant_score NaN)
};
static_assert(sizeof(ColumnOut) == 96, "ColumnOut must stay 96 bytes");

// What the host's sequential MC / UN interval walk (identify_mutations.cpp:2262-2344, 2972-3007) reads of a column: 8 bytes
// instead of the 96-byte result.  packed = total << 2 | (redundant > 0) << 1 | base_predicted, with
// total = round(unique) + round(redundant) as the reference computes it.
struct WalkOut { uint32_t unique, packed; };
// A column the interval walk has to look at: the state machines only change state at columns whose unique coverage is at
// or below the propagation cutoff, that are not predicted, that are flagged for host re-evaluation, or that open / close a
// target; the walk needs those columns and their two neighbours (the flank coverages of an MC row), nothing else.
struct WalkEvent { uint32_t slot; WalkOut w; };

constexpr uint32_t CO_BASE_PREDICTED = 1u << 12, CO_UNIQUE_ONLY = 1u << 13, CO_EMIT = 1u << 14, CO_RECHECK = 1u << 15,
                   CO_FIT = 1u << 24;

struct ScoreParams {
  double log10_ref_length;
  double mutation_cutoff, polymorphism_cutoff, precision_decimal;
  uint32_t base_quality_cutoff;
  uint32_t n_mapq_slots;     // distinct MAPQ values present
  uint32_t max_qual;         // Q of the table (quality covariate maximum)
  uint32_t max_set;
  uint8_t mapq_slot[256];    // MAPQ -> slot, 255 = absent
  uint32_t hot_mapq;         // the dominant MAPQ value, whose class terms are staged in shared memory
  uint32_t n_hot;            // entries of the hot tables (max_set * 2 * max_qual * 5), 0 = disabled
  uint32_t fit_all;          // evaluate the EM fit on every column with scoring records (diagnostics / parity runs)
  uint32_t keep_bounds;      // diagnostics: a slot the bounds settled keeps its upper bound in variant_score (else NaN)
  // tally kernel: per-slot class histogram over sq = (set*2 + top) * t_nq + quality - t_qlo (t_nsq classes in
  // t_nsq / 4 words, then two words of special counters) and the shared-memory likelihood table of the dominant
  // MAPQ, [obs A,C,G,T,'.'][sq] x {L[0..4], ratio column, top strand ? 1 : 0, top strand ? 0 : 1} (64 bytes a class: the B operand
  // of the contraction), t_stride bytes between the five obs planes (A, C, G, T, .)
  uint32_t t_qlo, t_nq, t_nsq, t_nw, t_stride;
  uint32_t mq_min, n_mq;     // MAPQ range of the global table the other scoring records read
  uint32_t n_rpos, n_rep;    // read_pos / base_repeat values of the table (1 = the covariate is not used); with them every
                             // class index gains a factor n_rpos * n_rep between quality and obs (class_rr)
};

// read_pos * n_rep + min(base_repeat, n_rep - 1) of a record's extension word (brq_types.h): base_repeat clamps
// (error_count.cpp:489-496), read_pos is range-checked on the host against the stream's maximum
#ifdef __CUDACC__
__host__ __device__
#endif
inline uint32_t class_rr(uint32_t ext, const ScoreParams& p) {
  const uint32_t rpos = p.n_rpos > 1 ? (ext & 0xFFFFu) : 0u, rep = (ext >> 16) & 255u;
  return rpos * p.n_rep + (rep < p.n_rep ? rep : p.n_rep - 1u);
}

// Per-class likelihood terms, built on the host with the same libm calls the reference makes
// (identify_mutations.cpp:3359-3384) so the per-record terms are bit-identical.
// L[b] = log10 P(obs | true base b), M = max_b L[b], r[b] = 10^(L[b] - M);
// r2 = max_{b != obs} r[b], or +inf when obs is not the class's most likely true base.  96 bytes.
struct ClassTerms { double L[5]; double r2; double r[5]; double M; };

// Shared-memory forms of the class terms for the dominant MAPQ, indexed ((set*2+top)*Q+qual)*5+obs.
struct alignas(32) HotTerms { double L[5]; double M; double pad[2]; };    // M: the class's ratio column, max over b != obs of 10^(L[b] - L[obs]) (the presence bound); 64 bytes: one 256-bit and one 128-bit load
struct HotRatios { double r[5]; double M; };   // M = max_b L[b]

// The fused collective of pass 1 (exchange.cu): every rank's inbox (two histogram-shaped copies of `capacity` words) and its two
// arrival counters, as mapped into this process.
struct HistPeers {
  unsigned long long* inbox[16];
  uint32_t* arrived[16];
  uint32_t world, rank;
  uint64_t capacity;
};
void launch_hist_exchange(unsigned long long* local, uint64_t n, const HistPeers& peers, uint32_t copy, uint32_t* done, uint32_t* err,
                          double timeout_seconds, cudaStream_t s);
void launch_score_slots(const uint32_t* rec, const uint64_t* off, const uint32_t* cnt, const uint64_t* round_off,
                        const uint32_t* side, const uint32_t* side_off, const uint2* round_side,
                        const uint8_t* slot_ref, const uint32_t* round_slot, uint64_t n_rounds, uint64_t n_slots, uint64_t n_records,
                        const ClassTerms* lut, const double* tallyT, const HotTerms* coldT, const HotRatios* hotR, const ScoreParams& p,
                        ColumnOut* out, uint32_t* worklist, uint32_t* survivors, uint32_t* flagged, uint32_t* scalars, uint32_t flagged_cap,
                        uint32_t side_stride, cudaStream_t s, cudaEvent_t between);
// Compacts the walk records to the columns the host's interval walk needs (WalkEvent), in no particular order.
// walk: scratch of n_base records, filled here from the full results.  seg_first / seg_last / seg_prop: first slot, last slot and deletion propagation cutoff of every visited segment
// (cutoff < 0: the target is skipped); mark: scratch of n_base bytes; counter: one zeroed word.
void launch_walk_events(const ColumnOut* cols, WalkOut* walk, uint64_t n_base, const uint32_t* seg_first, const uint32_t* seg_last,
                        const double* seg_prop, uint32_t n_seg, const uint32_t* flagged, const uint32_t* n_flagged,
                        uint32_t flagged_cap, const uint64_t* ins_parent, uint8_t* mark, WalkEvent* events, uint32_t* counter,
                        cudaStream_t s);
// out[i] = cols[slots[i]]: the full results of the flagged slots, for the host re-evaluation
void launch_gather_columns(const ColumnOut* cols, const uint32_t* slots, uint32_t n, ColumnOut* out, cudaStream_t s);

// Fills every device likelihood table from the text-canonical probabilities (identify_mutations.cpp:3359-3384).
struct TableBuildArgs {
  const double* prob;          // [n_bins] pow(10, log10 value read back from error_rates.tab)
  const uint8_t* slot_mapq;    // [n_mapq_slots] MAPQ value of each slot
  uint32_t n_st, n_mapq_slots, Q, off_set, off_ref, off_obs, off_qual, off_rpos, off_rep, hot_slot;
  ClassTerms* lut; HotTerms* coldT; HotRatios* hotR; double* tallyT;
};
void launch_build_tables(const TableBuildArgs& a, const ScoreParams& p, cudaStream_t s);

void launch_hist(const void* rec, uint64_t n_rec, bool wide, const CovLayout& lay, unsigned long long* counts, cudaStream_t s);
// transfer form of the scoring stream (low halves + exception words) -> score_rec in HBM (brq_types.h)
struct ScoreGeometry;
void launch_expand_score(const uint16_t* s16, const uint64_t* round_off, const uint32_t* round_slot, const uint8_t* slot_ref, const uint32_t* exc,
                         const uint32_t* exc_off, uint64_t n_rounds, const ScoreGeometry& geo, uint32_t* score_rec, cudaStream_t s);
// log10 table -> probabilities through the six-significant-digit text round trip, on the device (canonical.h)
void launch_canonical_table(const double* log10_prob, uint32_t n_bins, double* prob, uint32_t* err, cudaStream_t s);
// the compact form: n16 fast records of 16 bits and n_exc 4-byte records without a 16-bit form (brq_types.h)
void launch_hist16(const void* rec16, uint64_t n16, const uint32_t* exc, uint64_t n_exc, const CovLayout& lay, unsigned long long* counts, cudaStream_t s);
void launch_coverage_hist(const uint64_t* hist_off, const uint8_t* group, uint64_t n_cols, uint32_t stride, uint32_t n_groups,
                          unsigned long long* cov_hist, uint32_t* err, cudaStream_t s);
void launch_derive_table(const unsigned long long* counts, const CovLayout& lay, double* log10_prob, cudaStream_t s);
int launch_count();  // kernels launched so far through these wrappers (bench bookkeeping)

}  // namespace brq
