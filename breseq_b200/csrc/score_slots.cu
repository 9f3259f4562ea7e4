// Per-slot scoring for sm_100a: a streaming tally kernel, an EM fit kernel, and two small kernels that compact what
// the host's finalisation reads.
//
//   tally_kernel  One thread per slot, a warp per ROUND of 32 slots (staging groups them: one reference base, nearly
//                 equal depth), persistent CTAs of 24 warps.  The kernel COUNTS FIRST AND MULTIPLIES ONCE: a scoring
//                 record that matches its slot's reference base (all but ~0.1 % of them) only increments a byte counter
//                 of its (read set, strand, quality) class in the lane's private histogram in shared memory; when the
//                 round's records are in, the counts are contracted with the likelihood table of the round's base:
//                     sums[slot][column] += counts[slot][class] x table[class][column]
//                 (identify_mutations.cpp:1392-1658, 3359-3384, 3398-3433).
//   counting      red.shared.add.u32 of 1 << 8*byte on the counter's word (ATOMS.ADD without return: no dependent
//                 chain).  Word w of lane l lives at w * 128 + l * 4 of the warp's block: any mix of classes is
//                 bank-conflict free, and the block is aligned to its size, so the address is
//                 (record & 0x1F80) | lane base and the shift is record & 31.  Records that are not class counts (idle,
//                 redundant, cold, HOT-but-mismatching, padding) increment special counters, so the loop is branch-free
//                 except for one test per vector for the rare mismatching HOT record, which reads its own table cell.
//                 Counters are bytes: a round deeper than 224 records is contracted every 28 vectors.
//   contraction   on the fp64 tensor pipe: mma.sync.m8n8k4.f64 (DMMA.8x8x4), A = the byte counters of eight slot
//                 lanes for one class word (four classes), B = the four table rows {L[0..4], M, top, bottom}; four
//                 8-slot tiles per class word share B.  The last two columns count the matching records by strand.
//                 Why the tensor pipe: a record-by-record kernel reads 48 bytes of table per record and a per-lane
//                 contraction reads 48 bytes per class, both from shared memory at 12 wavefronts per 32 lookups (a
//                 128-bit shared load takes four passes even when the lanes broadcast): the data pipe, not HBM, was
//                 the bound.  Only register reuse of the table escapes it, and that is what the MMA fragments are.
//   record ring   The stream is round-major and lane-interleaved (brq_types.h): a warp streams the round's 1 KB
//                 vectors through a four-stage ring in shared memory with cp.async (two coalesced 16-byte copies per
//                 lane and vector); wait_group counts them in order.  The next round's first vectors are requested
//                 before the contraction of the current one, the rest of it goes to L2 with one bulk prefetch per
//                 warp; slot numbers are read two rounds ahead, the slots' geometry one round ahead.
//   cold records  Scoring records outside the shared table (another MAPQ, a quality outside the window; every scoring
//                 record of a stream staged with read_pos / base_repeat) sit in the side list as classic words and read
//                 the global table of all MAPQ values.  '.' observations are ordinary classes (a fifth observation
//                 plane of the shared table): a read without an inserted base is a '.' observation of the insert
//                 sub-column, so nearly every record of such a slot is one, and whole rounds of them close the stream.
//   redundant records  lead each slot's run (staging.cpp): their order-dependent sum of 1/X1
//                 (identify_mutations.cpp:1605) is a short sequential walk of the slot's head: bit-exact.
//   presence bound  The reference fits the 5-allele EM on every column, but its result only surfaces
//                 in RA rows, i.e. when best != ref with a positive consensus score or when the
//                 presence score of the top non-reference allele reaches the polymorphism cutoff
//                 (identify_mutations.cpp:1789-1836).  The tally carries sum_i M_i (M_i = max_b L_i[b])
//                 next to the five sums and proves most columns cannot emit:
//                     L_full <= sum_i M_i                      (every s_i <= 1)
//                     L_null >= LL_null(EM start)              (EM never lowers the likelihood)
//                            >= n log10 g0[ref] + ll[ref],     g0[ref] >= (0.5 + c_ref) / (n + 2)
//                 so  score <= (sum M - ll[ref]) - n log10((0.5 + c_ref)/(n + 2)) - log10(ref length).
//                 Columns under the cutoff by a margin are final here; the others go to a work list.
//   fit_kernel    one warp per work-list slot: the 5-allele EM fit, the presence score of the top non-reference
//                 allele (second EM with it held out), emission flags (identify_mutations.cpp:1797-1821, 3240-3344).
//                 A slot of up to 256 entries packs its scoring records' ratios into registers once; an EM step is
//                 then ceil(n / 32) records per lane, a Newton reciprocal each, and a 5-value butterfly.
//   walk events   walk_mark / walk_compact kernels: the columns the host's MC / UN interval state machines have to see
//                 (coverage at or below the propagation cutoff, not predicted, flagged, target ends, and their
//                 neighbours), a few thousand of 4.6 M at C1; gather_columns: the full results of the flagged slots.
#include <cstdlib>
#include "kernels.h"
#include "brq_types.h"

namespace brq {

void note_launches(int n);

namespace {

constexpr int TALLY_MAX_TPB = 768;
constexpr uint32_t PREFETCH_VECTORS = 2;   // round vectors (KB) of a round kept in L2 ahead of the record ring (BRQ_TALLY_PREFETCH overrides; -1 = none)
constexpr int RING = 4;           // 16-byte stages of each lane's record ring (a power of two)
constexpr int FIT_TPB = 256;
constexpr int FIT_LANES = 32;     // lanes cooperating on one slot: the work list is short, so a slot's latency is what counts
constexpr uint32_t FIT_HASH = 1024;  // slots of a warp's class -> count table (deep slots are fitted by class)
constexpr int FIT_REG = 8;        // records per lane whose ratios stay in registers over the whole fit (slots up to 256 entries)
constexpr uint32_t CODE_NONE = 0xFFFFFFFFu, CODE_COLD = 0x80000000u;

struct f64x2 { double x, y; };
__device__ __forceinline__ f64x2 lds_f64x2(uint32_t shared_addr) {
  f64x2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(shared_addr));
  return v;
}
__device__ __forceinline__ f64x2 ldg_f64x2(const void* p) {
  f64x2 v;
  asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// 16 bytes global -> shared without passing through registers (LDGSTS); completion is tracked per thread
__device__ __forceinline__ void cp_async16(uint32_t shared_addr, const void* p) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(shared_addr), "l"(p) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ uint4 lds_u32x4(uint32_t shared_addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(shared_addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t shared_addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(shared_addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t shared_addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" :: "r"(shared_addr), "r"(v) : "memory");
}
// counter field of a device word: word * 128 + byte * 8 inside the lane's histogram (lane base `hist`, block-aligned)
__device__ __forceinline__ void red_count(uint32_t hist, uint32_t w) {
  const uint32_t addr = (w & DR_COUNTER_WORD_MASK) | hist, one = 1u << (w & 31u);
  asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(addr), "r"(one) : "memory");
}

// ask L2 for [p, p + bytes) ahead of use (16-byte aligned, a multiple of 16 bytes)
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool eligible(uint32_t r, uint32_t cutoff) {
  return (r & (SR_UNIQUE_BIT | SR_TRIM_BIT | SR_OK_BIT)) == (SR_UNIQUE_BIT | SR_OK_BIT) && ((r >> SR_QUAL_SHIFT) & 127) >= cutoff;
}
// ((set*2 + top) * Q + qual) * 5 + obs, with top and set adjacent in the record
__device__ __forceinline__ uint32_t hot_index(uint32_t r, uint32_t Q) {
  return (((r >> 10) & 63) * Q + ((r >> SR_QUAL_SHIFT) & 127)) * 5 + (r & 7);
}
__device__ __forceinline__ uint32_t cold_index(uint32_t r, uint32_t ext, const ScoreParams& p, const uint8_t* mapq_slot) {
  const uint32_t hi = (r >> 10) & 63, mapq = (r >> SR_MAPQ_SHIFT) & 255;
  return (((hi * p.n_mapq_slots + mapq_slot[mapq]) * p.max_qual + ((r >> SR_QUAL_SHIFT) & 127)) * (p.n_rpos * p.n_rep) + class_rr(ext, p)) * 5 + (r & 7);
}

struct Sums { double l0, l1, l2, l3, l4, m; };  // (m: the sum of the records' ratio column, see the presence bound)
__device__ __forceinline__ float up_float(double x) { return __double2float_ru(x); }

// 10^d for d in [-17, 0], relative error below 1e-14: 2^(n + f), |f| <= 1/2, 2^f by its Taylor polynomial of degree 13.
// (The sum it feeds is 1 + terms <= 1 whose log10 is added to numbers of magnitude 1 to 10^4: far inside the 1e-9 bar.)
__device__ __forceinline__ double exp10_neg(double d) {
  const double x = d * 3.3219280948873623, n = rint(x), f = (x - n) * 0.6931471805599453;  // f in natural-log units
  double q = 1.6059043836821613e-10;                                                       // 1/13!
  q = fma(q, f, 2.08767569878681e-09); q = fma(q, f, 2.505210838544172e-08); q = fma(q, f, 2.755731922398589e-07);
  q = fma(q, f, 2.755731922398589e-06); q = fma(q, f, 2.48015873015873e-05); q = fma(q, f, 1.984126984126984e-04);
  q = fma(q, f, 1.388888888888889e-03); q = fma(q, f, 8.333333333333333e-03); q = fma(q, f, 4.1666666666666664e-02);
  q = fma(q, f, 1.6666666666666666e-01); q = fma(q, f, 0.5); q = fma(q, f, 1.0); q = fma(q, f, 1.0);
  return __hiloint2double(__double2hiint(q) + (int)n * 1048576, __double2loint(q));          // * 2^n, n in [-57, 0]: no underflow
}

// a scoring record whose class is not in the shared table (classic word from the side list): {L[0..4], M} from the global table
__device__ __forceinline__ void cold_add(Sums& a, uint32_t r, uint32_t ext, const HotTerms* __restrict__ coldT, const ScoreParams& p) {
  const uint32_t st = (r >> 10) & 63u, mapq = (r >> SR_MAPQ_SHIFT) & 255u, qual = (r >> SR_QUAL_SHIFT) & 127u;
  const HotTerms* e = coldT + ((((size_t)(st * p.n_mq + (mapq - p.mq_min)) * p.max_qual + qual) * (p.n_rpos * p.n_rep) + class_rr(ext, p)) * 5u + (r & 7u));
  // 48 of the entry's 64 bytes in two requests (a scattered request costs the load unit per lane, not per byte)
  double l0, l1, l2, l3;
  asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(l0), "=d"(l1), "=d"(l2), "=d"(l3) : "l"(e));
  const f64x2 z = ldg_f64x2(reinterpret_cast<const char*>(e) + 32);
  a.l0 += l0; a.l1 += l1; a.l2 += l2; a.l3 += l3; a.l4 += z.x; a.m += z.y;
}

}  // namespace

// ------------------------------------------------------------------------------------------ tally
// D (8 slots x 8 columns) += A (8 slots x 4 classes) * B (4 classes x 8 columns), fp64 (DMMA.8x8x4).
// Lane l = 4 g + t holds A[g][t], B[t][g] and D[g][2t], D[g][2t + 1].
__device__ __forceinline__ void dmma_8x8x4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// byte (selector 0x4440 | index) of a word as a double
__device__ __forceinline__ double byte_to_double(uint32_t w, uint32_t sel) {
  return (double)__byte_perm(w, 0u, sel);
}
// eight records of one lane: its share of a round vector
struct Vec8 { uint4 a, b; };
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

// Shared memory of one CTA (dynamic, base rounded up to the histogram block size):
//   [n_warps x block]   per-warp class histograms: word w of lane l at w * 128 + l * 4 (bank l: conflict-free for any
//                       mix of classes), four byte counters per word; block = 4 KB (<= 32 words) or 8 KB, and the
//                       block is aligned to its size, so counter address = (record & 0x1FFF) | lane base.  The first
//                       2 KB double as the scratch through which the contraction's result tiles reach their lanes.
//   [n_warps x 4 KB]    record rings: 4 stages of one round vector (two planes of 32 lanes x 16 bytes)
//   [5 x t_stride]      likelihood table [obs][sq] x 8 doubles (ScoreParams), obs = A, C, G, T, '.'
__global__ void __launch_bounds__(TALLY_MAX_TPB, 1) tally_kernel(const uint32_t* __restrict__ rec, const uint64_t* __restrict__ round_off,
                                                                  const uint32_t* __restrict__ side, const uint2* __restrict__ round_side,
                                                                  const uint32_t* __restrict__ round_slot,
                                                                  uint64_t n_rounds, const double* __restrict__ tallyT,
                                                                  const HotTerms* __restrict__ coldT, ScoreParams p, ColumnOut* __restrict__ out,
                                                                  uint32_t* __restrict__ worklist, uint32_t* __restrict__ flagged,
                                                                  uint32_t* __restrict__ scalars, uint32_t flagged_cap, uint32_t hist_block, uint32_t side_stride, uint32_t pf_vec) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const uint32_t n_warps_cta = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t sm0 = ((uint32_t)__cvta_generic_to_shared(sm_raw) + hist_block - 1u) & ~(hist_block - 1u);
  const uint32_t wbase = sm0 + warp * hist_block;   // this warp's histogram block
  const uint32_t hist = wbase + lane * 4u;          // this lane's word 0
  const uint32_t ring = sm0 + n_warps_cta * hist_block + warp * (RING * 1024u) + lane * 16u;  // this lane's cell of stage 0, first plane
  const uint32_t tbl = sm0 + n_warps_cta * (hist_block + RING * 1024u);
  {
    const uint32_t n16 = 5u * (p.t_stride / 16u);  // 16-byte cells of the table (5 observation planes of t_stride bytes)
    const uint4* src = reinterpret_cast<const uint4*>(tallyT);
    for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) {
      const uint4 v = src[i];
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(tbl + i * 16u), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
    for (uint32_t w = 0; w < p.t_nw; ++w) sts_u32(hist + w * 128u, 0u);
  }
  __syncthreads();
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  const uint32_t n_cw = p.t_nsq >> 2;                   // class words; the two special words follow
  const uint32_t g8 = lane >> 2, t4 = lane & 3u;        // fragment coordinates of this lane in the contraction
  const uint32_t a_sel = 0x4440u | t4;                  // byte t4 of a histogram word, zero-extended
  const uint32_t a_base = wbase + g8 * 4u;              // word 0 of slot-lane g8 (m-tile mt adds 32 mt bytes)
  const uint32_t b_off = t4 * 64u + g8 * 8u;            // B[t4][g8] inside a class word's four table rows
  const uint32_t c_sts = wbase + g8 * 64u + t4 * 16u;   // scratch row of slot-lane g8 (m-tile mt adds 512 mt bytes)
  const uint32_t c_lds = wbase + lane * 64u;            // this lane's scratch row
  const uint64_t n_warps = (uint64_t)gridDim.x * n_warps_cta;
  // rounds go to the warps CTA by CTA (warp w of CTA b takes rounds w * gridDim.x + b, + n_warps, ...): when the rounds do not
  // divide by the warps, the warps with one round more are spread over all SMs instead of filling the first few
  uint64_t round = (uint64_t)warp * gridDim.x + blockIdx.x;

  // A round: its records (round_off: warp-uniform) and this lane's slot with what closes it.  Slot numbers are read
  // two rounds ahead and the rest one round ahead, so no round starts by waiting for a chain of dependent loads.
  struct Run { uint64_t beg; uint32_t slot, n_vec, ref, side0, side1; };
  auto load_slot = [&](uint64_t r) { return r < n_rounds ? __ldg(round_slot + (r << 5) + lane) : ROUND_NO_SLOT; };
  auto load_run = [&](uint64_t r, uint32_t slot) {
    Run x{0, slot, 0, 5, 0, 0};
    if (r < n_rounds) {
      const uint64_t o0 = __ldg(round_off + r), o1 = __ldg(round_off + r + 1);
      x.beg = o0; x.n_vec = (uint32_t)((o1 - o0) / ROUND_VECTOR_WORDS);
    }
    if (r < n_rounds) {  // coalesced: side-list range and reference base of this lane's slot
      const uint2 m = __ldg(round_side + (r << 5) + lane);
      x.ref = m.x >> 29; x.side0 = m.x & 0x1FFFFFFFu; x.side1 = m.y;
    }
    return x;
  };
  // The warp streams the round's 1 KB vectors through a four-stage ring in shared memory with cp.async (LDGSTS, two
  // fully coalesced 16-byte copies per lane and vector): four vectors are always on their way without holding
  // registers, and wait_group counts them in order.  A lane only ever reads its own cells.
  const uint4* vp = nullptr;  // this lane's 16 bytes of the round's first plane
  const bool pf_on = pf_vec - 1u < 0xFFFFu;
  auto fetch = [&](uint32_t stage, uint32_t i, bool on) {  // round vector i into a stage; always one commit
    if (on) {
      cp_async16(ring + stage * 1024u, vp + (size_t)i * (ROUND_VECTOR_WORDS / 4));
      cp_async16(ring + stage * 1024u + 512u, vp + (size_t)i * (ROUND_VECTOR_WORDS / 4) + ROUND_VECTOR_WORDS / 8);
    }
    cp_async_commit();
  };
  auto start_run = [&](const Run& x) {
    vp = reinterpret_cast<const uint4*>(rec + x.beg) + lane;
#pragma unroll
    for (int st = 0; st < RING; ++st) fetch((uint32_t)st, (uint32_t)st, (uint32_t)st < x.n_vec);
    // the rest of the round into L2 (one request per warp), and this lane's side-list entries
    // (at most pf_vec vectors of it: 3552 warps prefetching whole deep rounds would evict each other from the 126 MB L2;
    // the record loop keeps that distance ahead of the ring)
    if (lane == 0 && x.n_vec > (uint32_t)RING && pf_on)
      prefetch_l2_bulk(rec + x.beg + (uint64_t)RING * ROUND_VECTOR_WORDS, min(x.n_vec - (uint32_t)RING, pf_vec) * (ROUND_VECTOR_WORDS * 4u));
    if (x.side1 > x.side0) prefetch_l2(side + (size_t)x.side0 * side_stride);
  };
  // round vector i: wait for it, read this lane's eight records, and hand the stage to vector i + RING
  auto next_vec = [&](uint32_t stage, uint32_t i, uint32_t n_vec) {
    Vec8 v;
    const uint32_t cell = ring + stage * 1024u;
    cp_async_wait<RING - 1>();
    v.a = lds_u32x4(cell); v.b = lds_u32x4(cell + 512u);
    fetch(stage, i + (uint32_t)RING, i + (uint32_t)RING < n_vec);
    return v;
  };

  Run cur = load_run(round, load_slot(round));
  uint32_t slot_nxt = load_slot(round + n_warps);
  start_run(cur);

  for (; round < n_rounds; round += n_warps) {
    const Run nxt = load_run(round + n_warps, slot_nxt);  // first used when this round's records are in
    slot_nxt = load_slot(round + 2 * n_warps);
    const uint32_t my_slot = cur.slot, my_ref = cur.ref, n_vec = cur.n_vec;
    const bool live = my_slot != ROUND_NO_SLOT;

    Sums kept = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    double red_top = 0.0, red_bot = 0.0;
    uint32_t raw_top = 0, raw_bot = 0, n = 0, c_ref = 0, u_top = 0, u_bot = 0;
    // What the presence bound (closing arithmetic) needs of the slot's records besides their count: the sum of the ratio
    // column over all of them, and of the records that do not match the reference base (all of them are side-list entries),
    // by observed base g (g' = g - (g > ref)), their number (four byte counters, 255 = many) and the sum of L[g] - L[ref].
    float rho = 0.f, mis_rho = 0.f, mis_gain0 = 0.f, mis_gain1 = 0.f, mis_gain2 = 0.f, mis_gain3 = 0.f;
    uint32_t mis_cnt = 0;
    auto mismatch = [&](const Sums& t, uint32_t obs) {
      const double l_obs = obs == 0u ? t.l0 : obs == 1u ? t.l1 : obs == 2u ? t.l2 : obs == 3u ? t.l3 : t.l4;
      const double l_ref = my_ref == 0u ? t.l0 : my_ref == 1u ? t.l1 : my_ref == 2u ? t.l2 : my_ref == 3u ? t.l3 : t.l4;
      const float gain = up_float(l_obs - l_ref);
      const uint32_t g = min(obs - (obs > my_ref ? 1u : 0u), 3u);
      mis_gain0 += g == 0u ? gain : 0.f; mis_gain1 += g == 1u ? gain : 0.f; mis_gain2 += g == 2u ? gain : 0.f; mis_gain3 += g == 3u ? gain : 0.f;
      if (((mis_cnt >> (8u * g)) & 255u) != 255u) mis_cnt += 1u << (8u * g);
      mis_rho += up_float(t.m);
    };

    // the scoring records of this lane's slot whose class is not in the shared table (another MAPQ, a '.'
    // observation, a quality outside the window) sit in the side list as classic words: their terms come from the
    // global table.  SIDE_BIG entries (X1 of very redundant records) lead the slot's side range; the head walk reads
    // them.  Two entries per step, the next pair's words requested before this pair's table terms.
    uint32_t side_big = cur.side0;
    if (side_stride == 1u) {
      // ranges start on even entries and are padded to even counts (SIDE_PAD reads as "skip"): one request per pair
      uint32_t e = cur.side0;
      const uint2* side2 = reinterpret_cast<const uint2*>(side);
      uint2 w2 = e < cur.side1 ? __ldg(side2 + (e >> 1)) : make_uint2(SIDE_PAD, SIDE_PAD);
      uint32_t wa = w2.x, wb = w2.y;
      while (e < cur.side1) {
        e += 2u;
        w2 = e < cur.side1 ? __ldg(side2 + (e >> 1)) : make_uint2(SIDE_PAD, SIDE_PAD);
        const uint32_t na = w2.x, nb = w2.y;
        Sums ta = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, tb = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        if (!(wa & SIDE_BIG)) { cold_add(ta, wa, 0u, coldT, p); ++n; c_ref += (wa >> 27) & 1u; if (!(wa & SR_MATCH_BIT)) mismatch(ta, wa & 7u); }
        if (!(wb & SIDE_BIG)) { cold_add(tb, wb, 0u, coldT, p); ++n; c_ref += (wb >> 27) & 1u; if (!(wb & SR_MATCH_BIT)) mismatch(tb, wb & 7u); }
        kept.l0 += ta.l0; kept.l1 += ta.l1; kept.l2 += ta.l2; kept.l3 += ta.l3; kept.l4 += ta.l4; rho += up_float(ta.m);
        kept.l0 += tb.l0; kept.l1 += tb.l1; kept.l2 += tb.l2; kept.l3 += tb.l3; kept.l4 += tb.l4; rho += up_float(tb.m);
        wa = na; wb = nb;
      }
    } else {
      // read_pos / base_repeat streams: every scoring record is here, two words an entry (classic word, extension)
      for (uint32_t e = cur.side0; e < cur.side1; ++e) {
        const uint2 w = __ldg(reinterpret_cast<const uint2*>(side) + e);
        if (w.x & SIDE_BIG) continue;
        Sums t = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        cold_add(t, w.x, w.y, coldT, p);
        ++n; c_ref += (w.x >> 27) & 1u;
        if (!(w.x & SR_MATCH_BIT)) mismatch(t, w.x & 7u);
        kept.l0 += t.l0; kept.l1 += t.l1; kept.l2 += t.l2; kept.l3 += t.l3; kept.l4 += t.l4; rho += up_float(t.m);
      }
    }

    // redundant records lead the slot: an order-dependent double sum, taken in arrival order
    // (identify_mutations.cpp:1605); the first record of any other kind (or a pad word) ends the walk
    if (live && n_vec) {
      const uint32_t cnt = n_vec * 8u;
      cp_async_wait<RING - 1>();  // the round's first vector is in stage 0
      uint32_t j = 0, r = lds_u32(ring);
      while ((r >> DR_KIND_SHIFT) == 3u) {
        uint32_t red = (r >> DR_X1_SHIFT) & DR_X1_MASK;
        if (red == DR_X1_MASK) red = __ldg(side + (size_t)(side_big++) * side_stride) & ~SIDE_BIG;
        const double inv = 1.0 / (double)red;
        if (r & DR_TOP_BIT) { red_top += inv; ++raw_top; } else { red_bot += inv; ++raw_bot; }
        if (++j == cnt) break;
        r = j < 4u ? lds_u32(ring + j * 4u) : j < 8u ? lds_u32(ring + 512u + (j - 4u) * 4u) : __ldg(rec + score_index(cur.beg + lane * 4u, j));
      }
    }

    // this lane's share of a round vector: eight records
    auto tally8 = [&](const Vec8& v) {
      // every record increments one byte counter of this lane's histogram: a shared-memory reduction (no return value,
      // nothing to wait for) of 1 << 8 * byte on the counter's word
      red_count(hist, v.a.x); red_count(hist, v.a.y); red_count(hist, v.a.z); red_count(hist, v.a.w);
      red_count(hist, v.b.x); red_count(hist, v.b.y); red_count(hist, v.b.z); red_count(hist, v.b.w);
    };

    // the likelihood table of the round's reference base (rounds hold one base; insert sub-columns have '.', N columns no class counts)
    const uint32_t ref_round = __reduce_min_sync(0xFFFFFFFFu, my_ref);
    const uint32_t b_base = tbl + (ref_round < 5u ? ref_round : 0u) * p.t_stride + b_off;
    const uint32_t n_max = n_vec;  // the round's vectors: the same for every lane (shallower slots end in pad words)
    uint32_t i = 0;  // round vectors consumed so far
    bool more;
    do {
      // byte counters: at most 224 records between two contractions
      const uint32_t chunk_end = min(n_max, i + 28u);
      for (; i + 4u <= chunk_end; i += 4u) {  // i is a multiple of four here: the ring's stage numbers are constants
        // deep rounds: every eighth vector asks L2 for the next eight beyond the prefetch distance
        if ((i & 7u) == 0u && lane == 0 && pf_on && i + (uint32_t)RING + pf_vec < n_vec)
          prefetch_l2_bulk(rec + cur.beg + (uint64_t)(i + (uint32_t)RING + pf_vec) * ROUND_VECTOR_WORDS,
                           min(n_vec - (i + (uint32_t)RING + pf_vec), 8u) * (ROUND_VECTOR_WORDS * 4u));
#pragma unroll
        for (uint32_t s4 = 0; s4 < 4u; ++s4) tally8(next_vec(s4, i + s4, n_vec));
      }
      for (; i < chunk_end; ++i) tally8(next_vec(i & (uint32_t)(RING - 1), i, n_vec));
      more = i < n_max;
      if (!more) start_run(nxt);  // this round's records are all in: the next round's first vectors travel during the contraction

      // ---- contraction: sums[slot][column] += counts[slot][class] x table[class][column] on the fp64 tensor pipe.
      // Four 8-slot tiles; one k-step per class word (4 classes): A = the word's four byte counters of eight slot
      // lanes, B = the four table rows (columns L[0..4], M, top, bottom).
      __syncwarp();
      double d[4][2] = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};
      {
        uint32_t ap = a_base, bp = b_base;
        for (uint32_t kk = 0; kk < n_cw; ++kk, ap += 128u, bp += 256u) {
          const uint32_t w0 = lds_u32(ap), w1 = lds_u32(ap + 32u), w2 = lds_u32(ap + 64u), w3 = lds_u32(ap + 96u);
          double bv;
          asm volatile("ld.shared.f64 %0, [%1];" : "=d"(bv) : "r"(bp) : "memory");
          dmma_8x8x4(d[0][0], d[0][1], byte_to_double(w0, a_sel), bv);
          dmma_8x8x4(d[1][0], d[1][1], byte_to_double(w1, a_sel), bv);
          dmma_8x8x4(d[2][0], d[2][1], byte_to_double(w2, a_sel), bv);
          dmma_8x8x4(d[3][0], d[3][1], byte_to_double(w3, a_sel), bv);
        }
      }
      // special counters (idle / cold records by strand; redundant and pad words count into the trash byte)
      const uint32_t s0 = lds_u32(hist + n_cw * 128u);
      __syncwarp();  // every lane has read the counters: the block's head becomes the scratch of the result tiles
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(c_sts + (uint32_t)mt * 512u), "d"(d[mt][0]), "d"(d[mt][1]) : "memory");
      __syncwarp();
      {
        const f64x2 x = lds_f64x2(c_lds), y = lds_f64x2(c_lds + 16u), z = lds_f64x2(c_lds + 32u), c = lds_f64x2(c_lds + 48u);
        kept.l0 += x.x; kept.l1 += x.y; kept.l2 += y.x; kept.l3 += y.y; kept.l4 += z.x; rho += up_float(z.y);
        const uint32_t m_top = (uint32_t)c.x, m_bot = (uint32_t)c.y;  // matching HOT records by strand: exact small integers
        n += m_top + m_bot; c_ref += m_top + m_bot; u_top += m_top; u_bot += m_bot;
      }
      __syncwarp();
      for (uint32_t w = 0; w < p.t_nw; ++w) sts_u32(hist + w * 128u, 0u);
      if (p.t_nw < 16u) for (uint32_t w = p.t_nw; w < 16u; ++w) sts_u32(hist + w * 128u, 0u);  // the scratch spans 16 words
      {
        const uint32_t idle_t = s0 & 255u, idle_b = (s0 >> 8) & 255u, cold_t = (s0 >> 16) & 255u, cold_b = s0 >> 24;
        u_top += idle_t + cold_t; u_bot += idle_b + cold_b;
      }
    } while (more);

    cur = nxt;
    if (!live) continue;

    const uint32_t ref = my_ref;
    const double ll[5] = {kept.l0, kept.l1, kept.l2, kept.l3, kept.l4};
    double consensus = nan, kept_bound = nan;
    uint32_t best = 5;
    bool need_fit = false;
    const double slack = 1e-6;
    if (n > 0) {  // pure_genotype_call, identify_mutations.cpp:3398-3433
      best = 0;
      double lb = ll[0];
#pragma unroll
      for (int b = 1; b < 5; ++b) if (ll[b] > lb) { best = b; lb = ll[b]; }
      double offv = -1.7976931348623157e308;
#pragma unroll
      for (int b = 0; b < 5; ++b) if ((uint32_t)b != best) offv = fmax(offv, ll[b]);
      double tot = 0.0;
#pragma unroll
      for (int b = 0; b < 5; ++b) {
        if ((uint32_t)b == best) continue;
        const double dd = ll[b] - offv;
        // the runner-up itself contributes exactly 1; anything below 2^-54 cannot change that sum
        tot += (dd == 0.0) ? 1.0 : (dd < -17.0 ? 0.0 : exp10_neg(dd));
      }
      consensus = (lb - (log10(tot) + offv)) - p.log10_ref_length;
      // an RA row needs best != ref with a positive consensus score, or a presence score at the cutoff
      need_fit = p.fit_all != 0u || ref >= 5u || (best != ref && consensus > -slack);
      if (!need_fit) {
        // An upper bound of the presence score of EVERY candidate v != ref (variant_presence_score,
        // identify_mutations.cpp:3329-3344: LL of the five-allele fit minus LL of the fit without v, minus log10 of the
        // reference length), from sums the record loop has anyway.  Write a record's likelihoods relative to its own
        // observation, r'_b = 10^(L[b] - L[obs]), rho = max over b != obs of r'_b (the tables' ratio column), and group the
        // records by observed base g (n_g of them; the matching ones are group ref):
        //  * five-allele fit: s_i(f) <= f_g + rho_i; the mean inside the logarithm (Jensen); any f on the simplex:
        //      LL_5 <= sum_g n_g log10(n_g (1 + P) / n),  P = sum_g mean rho_g <= (sum over the matching records) / n_ref + sum over the others
        //    (the maximum of sum_g n_g log10(f_g + rho_g) under sum f = 1, signs of f free);
        //  * fit without v: an EM step never lowers the likelihood, so its LL is at least the one of its start
        //    g0_b = (0.5 + n_b) / (2 + n - n_v) >= (0.5 + n_b) / (n + 2) (:3251-3266), and a mixture is at least one of its
        //    terms: records of group w != v by allele w (r'_w = 1), records of group v by the reference base (r'_ref =
        //    10^-(L[v] - L[ref])).
        // The per-record normalisers cancel between the two.  Single precision, rounded up by more than its error; a NaN or an
        // infinity (a class whose observation has probability zero) sends the slot on.
        const uint32_t c0 = mis_cnt & 255u, c1 = (mis_cnt >> 8) & 255u, c2 = (mis_cnt >> 16) & 255u, c3 = mis_cnt >> 24;
        const bool many = c0 == 255u || c1 == 255u || c2 == 255u || c3 == 255u || n > (1u << 20);
        const float LG = 0.30103001f;
        const float fn2 = (float)n + 2.0f, f_ref = (float)c_ref;
        const float scale = (1.0f + rho / fmaxf(f_ref, 1.0f) + mis_rho) / (float)n;   // (rho: over all records, not less than over the matching ones)
        const float lg_ref = __log2f((f_ref + 0.5f) / fn2);
        float upper = c_ref ? f_ref * __log2f(f_ref * scale) : 0.f;
        float lower = f_ref * lg_ref;
        float best_v = 0.f;
        auto group = [&](uint32_t c, float gain) {
          if (!c) return;
          const float fc = (float)c, lg_g = __log2f((fc + 0.5f) / fn2);
          upper += fc * __log2f(fc * scale);
          lower += fc * lg_g;
          best_v = fmaxf(best_v, fc * (lg_g - lg_ref) * LG + gain);
        };
        group(c0, mis_gain0); group(c1, mis_gain1); group(c2, mis_gain2); group(c3, mis_gain3);
        const float bound_f = (upper - lower) * LG + best_v + (2e-3f + (float)n * 2e-6f);
        const double bound = (double)bound_f - p.log10_ref_length;
        need_fit = many || !(bound < p.polymorphism_cutoff - slack);
        if (p.keep_bounds) kept_bound = bound;
      }
    }
    const bool base_predicted = consensus >= p.mutation_cutoff;
    const bool recheck = n > 0 && fabs(consensus - p.mutation_cutoff) < slack;

    uint32_t bits = best | (5u << 3) | (5u << 6) | (5u << 9);
    if (base_predicted) bits |= CO_BASE_PREDICTED;
    if (raw_top + raw_bot == 0) bits |= CO_UNIQUE_ONLY;
    if (recheck) bits |= CO_RECHECK;

    ColumnOut o;
#pragma unroll
    for (int b = 0; b < 5; ++b) o.ll[b] = ll[b];
    o.consensus_score = consensus; o.variant_score = need_fit ? nan : kept_bound;
    o.redundant[0] = red_bot; o.redundant[1] = red_top;
    o.unique[0] = u_bot; o.unique[1] = u_top; o.raw_redundant[0] = raw_bot; o.raw_redundant[1] = raw_top;
    o.n = n; o.bits = bits;
    {  // the 96-byte result in three 256-bit stores
      const double* od = reinterpret_cast<const double*>(&o);
      double* dst = reinterpret_cast<double*>(out + my_slot);
#pragma unroll
      for (int j = 0; j < 3; ++j)
        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" :: "l"(dst + 4 * j), "d"(od[4 * j]), "d"(od[4 * j + 1]), "d"(od[4 * j + 2]), "d"(od[4 * j + 3]) : "memory");
    }
    if (need_fit) worklist[atomicAdd(&scalars[2], 1u)] = my_slot;
    else if (recheck) { const uint32_t kf = atomicAdd(&scalars[1], 1u); if (kf < flagged_cap) flagged[kf] = my_slot; }
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------ fit
namespace {

struct GroupCtx {
  const uint32_t* rec; uint64_t base, beg, end;   // index space [beg, end): the slot's n_main records (word score_index(base, k)), then its side-list entries
  const uint32_t* side; uint32_t side_beg, side_stride; uint64_t n_main;
  uint32_t hot_base; const ClassTerms* lut; const uint8_t* mapq_slot; const ScoreParams* p;
  const uint32_t* cache;  // table codes of the slot's first FIT_CACHE records
  uint32_t sub, mask;     // lane within the group, shuffle mask of the group
};

// 1 / x for x in (0, 5]: the hardware's 2^-23 approximation and two Newton steps (a couple of ulp; the EM stops at 1e-6)
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(fma(-x, r, 1.0), r, r);
  return fma(fma(-x, r, 1.0), r, r);
}
__device__ __forceinline__ double group_sum(double v, uint32_t mask) {
#pragma unroll
  for (int o = FIT_LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}
__device__ __forceinline__ uint32_t group_sum_u32(uint32_t v, uint32_t mask) {
#pragma unroll
  for (int o = FIT_LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

// Where a record's class terms live: a byte offset into the shared table, CODE_COLD | index into
// the global table, or CODE_NONE for a record that does not score.
__device__ __forceinline__ uint32_t code_of(const GroupCtx& g, uint2 rx) {
  const ScoreParams& p = *g.p;
  const uint32_t r = rx.x;
  if (!eligible(r, p.base_quality_cutoff)) return CODE_NONE;
  if (p.n_hot && ((r >> SR_MAPQ_SHIFT) & 255) == p.hot_mapq) return hot_index(r, p.max_qual) * 48u;
  return CODE_COLD | cold_index(r, rx.y, p, g.mapq_slot);
}
// Classic word at index i of the slot's records for the fit: a HOT device word is decoded, a side-list
// entry is taken as it is, everything else (IDLE, the COLD placeholder, REDUNDANT, padding) does not score.
__device__ __forceinline__ uint2 classic_at(const GroupCtx& g, uint64_t i) {  // {classic word, extension word}
  const uint64_t k = i - g.beg;
  if (k < g.n_main) {
    const uint32_t d = __ldg(g.rec + score_index(g.base, k));
    if ((d >> DR_KIND_SHIFT) != 0u) return make_uint2(0u, 0u);
    const ScoreParams& p = *g.p;
    const uint32_t sq = (d >> DR_SQ_SHIFT) & DR_SQ_MASK, obs = (d >> DR_OBS_SHIFT) & 7u, qual = p.t_qlo + sq % p.t_nq, st = sq / p.t_nq;
    return make_uint2(obs | qual << SR_QUAL_SHIFT | st << 10 | p.hot_mapq << SR_MAPQ_SHIFT | SR_UNIQUE_BIT | SR_OK_BIT, 0u);
  }
  const size_t e = (size_t)(g.side_beg + (uint32_t)(k - g.n_main)) * g.side_stride;
  const uint32_t w = __ldg(g.side + e);
  if (w & SIDE_BIG) return make_uint2(0u, 0u);
  return make_uint2(w, g.side_stride == 2u ? __ldg(g.side + e + 1) : 0u);
}
__device__ __forceinline__ uint32_t code_at(const GroupCtx& g, uint64_t i) {
  const uint64_t k = i - g.beg;
  (void)k;
  return code_of(g, classic_at(g, i));  // only the record-by-record path of a slot with too many classes gets here
}
// r[0..4] and M = max_b L[b]
__device__ __forceinline__ void load_ratios(const GroupCtx& g, uint32_t code, double* rr, double& M) {
  f64x2 a, b, c;
  if (!(code & CODE_COLD)) {
    const uint32_t e = g.hot_base + code;
    a = lds_f64x2(e); b = lds_f64x2(e + 16); c = lds_f64x2(e + 32);
  } else {
    const char* e = reinterpret_cast<const char*>(g.lut + (code & ~CODE_COLD)) + 48;
    a = ldg_f64x2(e); b = ldg_f64x2(e + 16); c = ldg_f64x2(e + 32);
  }
  rr[0] = a.x; rr[1] = a.y; rr[2] = b.x; rr[3] = b.y; rr[4] = c.x; M = c.y;
}

}  // namespace


// ------------------------------------------------------------------------------------------ screen
// Most work-list slots of a deep run are noise: a handful of stray mismatches whose presence score is far below the cutoff
// (1000x, polymorphism mode: one slot in six passes the tally's bound, one in eight thousand emits).  Two EM fits of a dozen
// iterations each settle nothing the two likelihood bounds below do not settle in ONE pass over the slot's classes:
//   L_full (the reference's iterate) <= max_f L(f) <= L(f~) + n log10 max_b D_b(f~),   D_b = (1/n) sum_i r_ib / s_i(f~)
//        (Jensen: L(f) - L(f~) = sum_i log(s_i / s~_i) <= n log((1/n) sum_i s_i / s~_i) = n log sum_b f_b D_b), any f~ inside
//        the simplex; f~ = the EM's own start, (0.5 + c_b) / (n + 2.5);
//   L_null(v) (the reference's iterate) >= LL_null(g0(v)),   g0(v) = the null EM's start (0.5 + c_b) / (n + 2 - c_v), b != v
//        (an EM step never lowers the likelihood, identify_mutations.cpp:3240-3318);
// so  score(v) <= sum_i [log10 s~_i - log10 s_null,i(v)] + n log10 max_b D_b - log10(reference length)  for every v != ref
// (the per-record maxima M_i cancel).  A slot whose bound is under the cutoff by a margin for all four candidates cannot emit
// an RA row through the polymorphism test; it keeps the tally's result (no fit: variant_score NaN, like every slot the
// tally's own bound settled).  The others go on to fit_kernel.  One warp per slot: the HOT records are counted into a dense
// [5][n_sq] array (their device words name the cell; MATCH.ANY elects one writer per cell), the few cold records are terms
// of their own; single-precision logarithms (the margin of 0.25 covers their rounding many times over).
constexpr int SCREEN_TPB = 256;
constexpr float SCREEN_MARGIN = 0.25f;
__global__ void __launch_bounds__(SCREEN_TPB) screen_kernel(const uint32_t* __restrict__ rec, const uint64_t* __restrict__ off, const uint32_t* __restrict__ cnt,
                                                             const uint32_t* __restrict__ side, const uint32_t* __restrict__ side_off,
                                                             const uint8_t* __restrict__ slot_ref, const uint32_t* __restrict__ worklist,
                                                             const ClassTerms* __restrict__ lut, const HotRatios* __restrict__ hotR,
                                                             ScoreParams p, ColumnOut* __restrict__ out, uint32_t* __restrict__ survivors,
                                                             uint32_t* __restrict__ scalars) {
  extern __shared__ __align__(16) unsigned char screen_sm[];
  __shared__ uint8_t mapq_slot[256];
  float* ratio = reinterpret_cast<float*>(screen_sm);                       // [n_hot][5], index ((st * Q + qual) * 5 + obs) like the fit's table
  const uint32_t n_cells = 5u * p.t_nsq;
  uint32_t* hist = reinterpret_cast<uint32_t*>(ratio + (size_t)p.n_hot * 5) + (threadIdx.x >> 5) * n_cells;
  const uint32_t n_work = scalars[2];
  if ((uint64_t)blockIdx.x * (SCREEN_TPB / 32) >= n_work) return;
  for (uint32_t i = threadIdx.x; i < p.n_hot * 5u; i += blockDim.x) ratio[i] = (float)hotR[i / 5u].r[i % 5u];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) mapq_slot[i] = p.mapq_slot[i];
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u;
  const float log10_2 = 0.30102999566f;
  for (;;) {
    uint32_t w = 0;
    if (lane == 0) w = atomicAdd(&scalars[5], 1u);
    w = __shfl_sync(0xFFFFFFFFu, w, 0);
    if (w >= n_work) break;
    const uint32_t slot = worklist[w];
    const uint32_t ref = slot_ref[slot], bits = out[slot].bits, best = bits & 7u;
    const double consensus = out[slot].consensus_score;
    // a consensus call against the reference emits whatever the presence score says; a slot without a reference base has
    // no reference allele to hold: both need the fit
    // (and a consensus score within rounding of its cutoff is re-checked on the host: the fit kernel flags it)
    bool keep = ref >= 5u || (best != ref && consensus > -1e-6) || (bits & CO_RECHECK);
    if (!keep) {
      const uint64_t base = off[slot];
      const uint32_t n_main = cnt[slot], s0 = side_off[slot], s1 = side_off[slot + 1];
      for (uint32_t h = lane; h < n_cells; h += 32) hist[h] = 0u;
      __syncwarp();
      for (uint32_t j0 = 0; j0 < n_main; j0 += 128u) {  // four requests in flight per lane
        uint32_t d[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const uint32_t j = j0 + (uint32_t)u * 32u + lane; d[u] = j < n_main ? __ldg(rec + score_index(base, j)) : DR_IDLE; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool hot = (d[u] >> DR_KIND_SHIFT) == 0u;
          const uint32_t idx = hot ? ((d[u] >> DR_OBS_SHIFT) & 7u) * p.t_nsq + ((d[u] >> DR_SQ_SHIFT) & DR_SQ_MASK) : 0xFFFFFFFFu;
          const uint32_t m = __match_any_sync(0xFFFFFFFFu, idx);
          if (hot && lane == (uint32_t)(__ffs(m) - 1)) hist[idx] += (uint32_t)__popc(m);
          __syncwarp();
        }
      }
      // observations per base: the dense cells, then the cold records (side list: classic words)
      uint32_t c[5] = {0, 0, 0, 0, 0};
      for (uint32_t h = lane; h < n_cells; h += 32) { const uint32_t v = hist[h], o = h / p.t_nsq; c[0] += o == 0 ? v : 0; c[1] += o == 1 ? v : 0; c[2] += o == 2 ? v : 0; c[3] += o == 3 ? v : 0; c[4] += o == 4 ? v : 0; }
      for (uint32_t e = s0 + lane; e < s1; e += 32) { const uint32_t sw = __ldg(side + e); if (!(sw & SIDE_BIG)) { const uint32_t o = sw & 7u; c[0] += o == 0; c[1] += o == 1; c[2] += o == 2; c[3] += o == 3; c[4] += o == 4; } }
#pragma unroll
      for (int b = 0; b < 5; ++b) c[b] = __reduce_add_sync(0xFFFFFFFFu, c[b]);
      const uint32_t n = c[0] + c[1] + c[2] + c[3] + c[4];
      if (n == 0) continue;   // (warp-uniform) nothing scores here: nothing to emit, the tally's result stands
      float wgt[5], inv_null[5];
#pragma unroll
      for (int b = 0; b < 5; ++b) { wgt[b] = 0.5f + (float)c[b]; inv_null[b] = 1.0f / ((float)n + 2.0f - (float)c[b]); }
      const float inv_tot = 1.0f / ((float)n + 2.5f);
      float D[5] = {0, 0, 0, 0, 0}, T[5] = {0, 0, 0, 0, 0};
      auto term = [&](const float* r, float count) {
        const float A = wgt[0] * r[0] + wgt[1] * r[1] + wgt[2] * r[2] + wgt[3] * r[3] + wgt[4] * r[4];
        const float s_full = A * inv_tot, k = count / s_full;
#pragma unroll
        for (int b = 0; b < 5; ++b) {
          D[b] += k * r[b];
          const float s_null = (A - wgt[b] * r[b]) * inv_null[b];   // zero or negative rounding: the logarithm says +inf / NaN and the slot goes on
          T[b] += count * __log2f(s_full / s_null);
        }
      };
      for (uint32_t h = lane; h < n_cells; h += 32) {
        const uint32_t v = hist[h];
        if (!v) continue;
        const uint32_t obs = h / p.t_nsq, sq = h % p.t_nsq;
        term(ratio + (((sq / p.t_nq) * p.max_qual + p.t_qlo + sq % p.t_nq) * 5u + obs) * 5u, (float)v);
      }
      for (uint32_t e = s0 + lane; e < s1; e += 32) {
        const uint32_t sw = __ldg(side + e);
        if (sw & SIDE_BIG) continue;
        const double* rd = lut[cold_index(sw, 0u, p, mapq_slot)].r;
        const float r[5] = {(float)rd[0], (float)rd[1], (float)rd[2], (float)rd[3], (float)rd[4]};
        term(r, 1.0f);
      }
      float maxD = 0.0f;
#pragma unroll
      for (int b = 0; b < 5; ++b) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { D[b] += __shfl_xor_sync(0xFFFFFFFFu, D[b], o); T[b] += __shfl_xor_sync(0xFFFFFFFFu, T[b], o); }
        maxD = fmaxf(maxD, D[b]);
      }
      const float slack = (float)n * __log2f(fmaxf(maxD / (float)n, 1.0f)) * log10_2;
      float bound = -3.0e38f;
      bool bad = false;
#pragma unroll
      for (int b = 0; b < 5; ++b) {
        if ((uint32_t)b == ref) continue;
        const float sc = T[b] * log10_2 + slack - (float)p.log10_ref_length;
        bad = bad || !(sc == sc) || sc > 3.0e38f;
        bound = fmaxf(bound, sc);
      }
      keep = bad || !(bound < (float)p.polymorphism_cutoff - SCREEN_MARGIN);
      if (!keep && p.keep_bounds && lane == 0) out[slot].variant_score = (double)bound;
    }
    if (keep && lane == 0) survivors[atomicAdd(&scalars[4], 1u)] = slot;
  }
}

__global__ void __launch_bounds__(FIT_TPB, 1) fit_kernel(const uint32_t* __restrict__ rec, const uint64_t* __restrict__ off, const uint32_t* __restrict__ cnt,
                                                          const uint32_t* __restrict__ side, const uint32_t* __restrict__ side_off,
                                                          const uint8_t* __restrict__ slot_ref, const uint32_t* __restrict__ worklist,
                                                          const uint32_t* __restrict__ n_work_ptr,
                                                          const ClassTerms* __restrict__ lut, const HotRatios* __restrict__ hotR,
                                                          ScoreParams p, ColumnOut* __restrict__ out, uint32_t* __restrict__ flagged,
                                                          uint32_t* __restrict__ scalars, uint32_t flagged_cap, uint32_t side_stride) {
  extern __shared__ __align__(16) double sm[];
  __shared__ uint8_t mapq_slot[256];
  // per warp, after the hot table: FIT_HASH class codes and FIT_HASH counts (the first 256 codes double as the
  // code cache of a shallow slot)
  uint32_t* warp_tab = reinterpret_cast<uint32_t*>(sm + (size_t)p.n_hot * 6) + (threadIdx.x / FIT_LANES) * (2u * FIT_HASH);
  const uint32_t n_work = *n_work_ptr;
  if ((uint64_t)blockIdx.x * (FIT_TPB / FIT_LANES) >= n_work) return;  // the work list is short: most CTAs have nothing to do
  {
    const double* src = reinterpret_cast<const double*>(hotR);
    for (uint32_t i = threadIdx.x; i < p.n_hot * 6; i += blockDim.x) sm[i] = src[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) mapq_slot[i] = p.mapq_slot[i];
  }
  __syncthreads();
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  const uint32_t lane = threadIdx.x & 31;
  GroupCtx g;
  g.rec = rec; g.hot_base = (uint32_t)__cvta_generic_to_shared(sm); g.lut = lut; g.mapq_slot = mapq_slot; g.p = &p;
  g.sub = lane % FIT_LANES;
  g.mask = FIT_LANES == 32 ? 0xFFFFFFFFu : (((1u << (FIT_LANES & 31)) - 1u) << (lane - g.sub));
  uint32_t* my_cache = warp_tab;
  uint32_t* my_count = warp_tab + FIT_HASH;
  g.cache = my_cache;
  for (;;) {
    uint32_t w = 0;
    if (g.sub == 0) w = atomicAdd(&scalars[3], 1u);  // slots are handed out one at a time: EM lengths vary widely
    w = __shfl_sync(g.mask, w, lane - g.sub);
    if (w >= n_work) break;
    const uint32_t slot = worklist[w];
    g.base = off[slot]; g.beg = 0; g.n_main = cnt[slot]; g.end = g.n_main;
    g.side = side; g.side_beg = side_off[slot]; g.side_stride = side_stride;
    g.end += side_off[slot + 1] - g.side_beg;
    // what closes the slot, requested before anything waits on the records
    const uint32_t ref = slot_ref[slot];
    const double consensus = out[slot].consensus_score;
    uint32_t bits = out[slot].bits;
    // A slot of ordinary depth (up to 256 entries) keeps each lane's records' ratios in registers: its records are
    // requested together (one round trip to DRAM instead of one per step), and the EM iterations, each a dependent
    // chain, run without a load.
    const bool in_regs = g.end - g.beg <= (uint64_t)(FIT_LANES * FIT_REG);
    uint32_t obs_count[5] = {0, 0, 0, 0, 0}, n = 0;
    double R[FIT_REG][5], Mr[FIT_REG], Cn[FIT_REG];  // ratios, max log-likelihood and weight (1, or the class count)
    bool ok[FIT_REG];
    bool by_class = false;
    uint32_t k_max = 0;  // records (or classes) per lane that hold anything (warp-uniform)
    if (in_regs) {
      uint2 rx[FIT_REG];
#pragma unroll
      for (int k = 0; k < FIT_REG; ++k) {
        const uint64_t i = g.beg + g.sub + (uint64_t)k * FIT_LANES;
        rx[k] = i < g.end ? classic_at(g, i) : make_uint2(0u, 0u);
      }
      // the scoring records are packed to the front (in index order) through the code cache, so that the loops below
      // run over ceil(n / 32) records per lane instead of over every entry
      uint32_t n_before = 0;
#pragma unroll
      for (int k = 0; k < FIT_REG; ++k) {
        const uint32_t code = code_of(g, rx[k]);
        const bool valid = code != CODE_NONE;
        const uint32_t m = __ballot_sync(g.mask, valid);
        if (valid) {
          my_cache[n_before + __popc(m & ((1u << g.sub) - 1u))] = code;
#pragma unroll
          for (int b = 0; b < 5; ++b) obs_count[b] += ((rx[k].x & 7) == (uint32_t)b);
          ++n;
        }
        n_before += __popc(m);
      }
      __syncwarp(g.mask);
      k_max = (n_before + FIT_LANES - 1) / FIT_LANES;
#pragma unroll
      for (int k = 0; k < FIT_REG; ++k) {
        Mr[k] = 0.0; Cn[k] = 1.0;
#pragma unroll
        for (int b = 0; b < 5; ++b) R[k][b] = 0.0;
        ok[k] = g.sub + (uint32_t)k * FIT_LANES < n_before;
        if (ok[k]) load_ratios(g, my_cache[g.sub + (uint32_t)k * FIT_LANES], R[k], Mr[k]);
      }
      __syncwarp(g.mask);
    } else {
      // A deep slot is fitted BY CLASS: its records fall into a few hundred (read set, strand, MAPQ, quality, obs)
      // classes, and a record's responsibilities depend on its class alone, so an EM step over classes weighted by
      // their counts is the same sum with a tenth of the terms.  The warp counts the classes in a hash table in
      // shared memory, packs the table, and keeps up to FIT_REG classes per lane in registers.
      // Counting without a hash when the stream has a shared table: a HOT record's device word names its cell
      // (observation, class) directly, so the warp counts into a dense [5][n_sq] array, lanes that hold the same cell
      // agreeing on one writer (MATCH.ANY: no atomics, no probe chains; thirty-two lanes adding to the same few cells with
      // shared-memory atomics serialise, which is what made this phase twenty times the EM it feeds).  The few records
      // outside the shared table (side list: another MAPQ, a quality outside the window) are classes of one record each.
      const uint32_t n_cells = 5u * p.t_nsq;
      bool overflow = false;
      uint32_t n_classes = 0;
      bool counted = false;
      if (p.n_hot != 0u && n_cells <= FIT_HASH && side_stride == 1u) {
        for (uint32_t h = g.sub; h < n_cells; h += FIT_LANES) my_count[h] = 0u;
        __syncwarp(g.mask);
        for (uint64_t j0 = 0; j0 < g.n_main; j0 += 4u * FIT_LANES) {  // four requests in flight per lane
          uint32_t d[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint64_t j = j0 + (uint64_t)u * FIT_LANES + g.sub;
            d[u] = j < g.n_main ? __ldg(g.rec + score_index(g.base, j)) : DR_IDLE;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const bool hot = (d[u] >> DR_KIND_SHIFT) == 0u;
            const uint32_t idx = hot ? ((d[u] >> DR_OBS_SHIFT) & 7u) * p.t_nsq + ((d[u] >> DR_SQ_SHIFT) & DR_SQ_MASK) : 0xFFFFFFFFu;
            const uint32_t m = __match_any_sync(g.mask, idx);
            if (hot && g.sub == (uint32_t)(__ffs(m) - 1)) my_count[idx] += (uint32_t)__popc(m);
            __syncwarp(g.mask);
          }
        }
        // the occupied cells, packed to the front as (table code, count)
        for (uint32_t h0 = 0; h0 < n_cells; h0 += FIT_LANES) {
          const uint32_t h = h0 + g.sub, cnt = h < n_cells ? my_count[h] : 0u;
          const uint32_t m = __ballot_sync(g.mask, cnt != 0u);
          __syncwarp(g.mask);
          if (cnt) {
            const uint32_t at = n_classes + __popc(m & ((1u << g.sub) - 1u));
            const uint32_t obs = h / p.t_nsq, sq = h % p.t_nsq;
            my_cache[at] = (((sq / p.t_nq) * p.max_qual + p.t_qlo + sq % p.t_nq) * 5u + obs) * 48u;
            my_count[at] = cnt;
          }
          n_classes += __popc(m);
          __syncwarp(g.mask);
        }
        // side-list entries (classic words of the cold records; SIDE_BIG / pad entries do not score)
        const uint32_t side_end = g.side_beg + (uint32_t)(g.end - g.n_main);
        for (uint32_t e0 = g.side_beg; e0 < side_end; e0 += FIT_LANES) {
          const uint32_t e = e0 + g.sub, w = e < side_end ? __ldg(g.side + e) : SIDE_PAD;
          const bool valid = !(w & SIDE_BIG);
          const uint32_t m = __ballot_sync(g.mask, valid);
          const uint32_t at = n_classes + __popc(m & ((1u << g.sub) - 1u));
          if (valid && at < FIT_HASH) { my_cache[at] = CODE_COLD | cold_index(w, 0u, p, mapq_slot); my_count[at] = 1u; }
          n_classes += __popc(m);
        }
        __syncwarp(g.mask);
        counted = n_classes <= (uint32_t)(FIT_LANES * FIT_REG);
      }
      if (!counted) {
      n_classes = 0;
      for (uint32_t h = g.sub; h < FIT_HASH; h += FIT_LANES) { my_cache[h] = CODE_NONE; my_count[h] = 0u; }
      __syncwarp(g.mask);
      for (uint64_t i = g.beg + g.sub; i < g.end; i += FIT_LANES) {
        const uint32_t code = code_of(g, classic_at(g, i));
        if (code == CODE_NONE) continue;
        uint32_t h = (code * 2654435761u) >> 22;
        uint32_t probes = 0;
        for (; probes < FIT_HASH; ++probes, h = (h + 1u) & (FIT_HASH - 1u)) {
          const uint32_t old = atomicCAS(&my_cache[h], CODE_NONE, code);
          if (old == CODE_NONE || old == code) { atomicAdd(&my_count[h], 1u); break; }
        }
        if (probes == FIT_HASH) overflow = true;
      }
      __syncwarp(g.mask);
      // pack the occupied slots to the front (in place: a slot's new position is never above its old one)
      for (uint32_t h0 = 0; h0 < FIT_HASH; h0 += FIT_LANES) {
        const uint32_t code = my_cache[h0 + g.sub], cnt = my_count[h0 + g.sub];
        const uint32_t m = __ballot_sync(g.mask, code != CODE_NONE);
        __syncwarp(g.mask);
        if (code != CODE_NONE) {
          const uint32_t at = n_classes + __popc(m & ((1u << g.sub) - 1u));
          my_cache[at] = code; my_count[at] = cnt;
        }
        n_classes += __popc(m);
        __syncwarp(g.mask);
      }
      }
      by_class = !__any_sync(g.mask, overflow) && n_classes <= (uint32_t)(FIT_LANES * FIT_REG);
      k_max = by_class ? (n_classes + FIT_LANES - 1) / FIT_LANES : 0u;
#pragma unroll
      for (int k = 0; k < FIT_REG; ++k) {
        Mr[k] = 0.0; Cn[k] = 1.0;
#pragma unroll
        for (int b = 0; b < 5; ++b) R[k][b] = 0.0;
        const uint32_t at = g.sub + (uint32_t)k * FIT_LANES;
        ok[k] = by_class && at < n_classes;
        if (ok[k]) {
          const uint32_t code = my_cache[at], cnt = my_count[at];
          load_ratios(g, code, R[k], Mr[k]);
          Cn[k] = (double)cnt;
          // the class index ends in the observed base (hot: index * 48, cold: CODE_COLD | index)
          const uint32_t obs = ((code & CODE_COLD) ? (code & ~CODE_COLD) : code / 48u) % 5u;
#pragma unroll
          for (int b = 0; b < 5; ++b) obs_count[b] += obs == (uint32_t)b ? cnt : 0u;
          n += cnt;
        }
      }
      __syncwarp(g.mask);
      if (!by_class) {  // more classes than the registers hold (read_pos covariates on a deep column): record by record
        for (uint64_t i = g.beg + g.sub; i < g.end; i += FIT_LANES) {
          const uint2 rx = classic_at(g, i);
          if (code_of(g, rx) == CODE_NONE) continue;
#pragma unroll
          for (int b = 0; b < 5; ++b) obs_count[b] += ((rx.x & 7) == (uint32_t)b);
          ++n;
        }
      }
    }
#pragma unroll
    for (int b = 0; b < 5; ++b) obs_count[b] = group_sum_u32(obs_count[b], g.mask);
    n = group_sum_u32(n, g.mask);
    if (n == 0) continue;
    const uint32_t best = bits & 7;
    bool recheck = (bits & CO_RECHECK) != 0;
    const double tol = p.precision_decimal, inv_n = 1.0 / (double)n, thr = 0.5 / (double)n;

    // Two EM fits share one copy of the code: all five alleles, then (when the first fit places a
    // non-reference allele at or above the half-read level) the same fit with that allele held out
    // (identify_mutations.cpp:3240-3344).  Records are strided over the group's lanes; responsibilities
    // are summed with a fixed butterfly, so every lane of the group holds the same frequencies.
    uint32_t allowed = 0x1F, major = 5, minor = 5, variant = 5, iterations = 0;
    double ll_fit[2] = {0.0, 0.0};
    for (int pass = 0; pass < 2; ++pass) {
      double f[5], f_prev[5], total = 0.0;
#pragma unroll
      for (int b = 0; b < 5; ++b) { f[b] = (allowed >> b & 1) ? 0.5 + (double)obs_count[b] : 0.0; total += f[b]; }
#pragma unroll
      for (int b = 0; b < 5; ++b) f[b] /= total;
      uint32_t it = 1;
      for (; it <= 50; ++it) {
        double w[5] = {0, 0, 0, 0, 0};
        if (in_regs || by_class) {
          // branch-free, the records' (classes') chains side by side: sums, then reciprocals, then responsibilities.
          // (A record whose sum is zero contributes f itself, like the reference; an absent one contributes nothing:
          // its ratios are zero and the select below drops it.)
#pragma unroll
          for (int k = 0; k < FIT_REG; ++k) {
            if ((uint32_t)k >= k_max) break;  // warp-uniform
            double a[5], sum = 0.0;
#pragma unroll
            for (int b = 0; b < 5; ++b) { a[b] = f[b] * R[k][b]; sum += a[b]; }
            const double inv = sum > 1e-280 ? fast_rcp(sum) : 1.0 / (sum > 0.0 ? sum : 1.0);  // the approximation flushes subnormals
#pragma unroll
            for (int b = 0; b < 5; ++b) {
              const double term = sum > 0.0 ? a[b] * inv : f[b];
              w[b] += ok[k] ? Cn[k] * term : 0.0;
            }
          }
        } else
        for (uint64_t i = g.beg + g.sub; i < g.end; i += FIT_LANES) {
          const uint32_t code = code_at(g, i);
          if (code == CODE_NONE) continue;
          double rr[5], M;
          load_ratios(g, code, rr, M);
          double a[5], sum = 0.0;
#pragma unroll
          for (int b = 0; b < 5; ++b) { a[b] = f[b] * rr[b]; sum += a[b]; }
          if (sum > 0.0) {
            const double inv = 1.0 / sum;
#pragma unroll
            for (int b = 0; b < 5; ++b) w[b] += a[b] * inv;
          } else {
#pragma unroll
            for (int b = 0; b < 5; ++b) w[b] += f[b];
          }
        }
        double max_delta = 0.0;
#pragma unroll
        for (int b = 0; b < 5; ++b) {
          f_prev[b] = f[b];
          if (allowed >> b & 1) {
            const double f_new = group_sum(w[b], g.mask) * inv_n;
            max_delta = fmax(max_delta, fabs(f_new - f[b]));
            f[b] = f_new;
          }
        }
        if (max_delta < tol) break;
      }
      // The committed likelihood belongs to the frequencies BEFORE the last update:
      // sum_i (log10 s_i + M_i).  The s_i (each in (0, 1]) are multiplied up and one log10 is taken
      // per ~200 decades, which is the same sum to within a few ulps of its terms.
      double log_sum = 0.0, prod = 1.0, m_sum = 0.0;
      if (in_regs || by_class) {
#pragma unroll
        for (int k = 0; k < FIT_REG; ++k) {
          if ((uint32_t)k >= k_max) break;  // warp-uniform
          if (!ok[k]) continue;
          double sum = 0.0;
#pragma unroll
          for (int b = 0; b < 5; ++b) sum += f_prev[b] * R[k][b];
          if (sum > 0.0) {
            if (by_class) { log_sum += Cn[k] * log10(sum); m_sum += Cn[k] * Mr[k]; }
            else {
              prod *= sum; m_sum += Mr[k];
              if (prod < 1e-200) { log_sum += log10(prod); prod = 1.0; }
            }
          }
        }
      } else
      for (uint64_t i = g.beg + g.sub; i < g.end; i += FIT_LANES) {
        const uint32_t code = code_at(g, i);
        if (code == CODE_NONE) continue;
        double rr[5], M;
        load_ratios(g, code, rr, M);
        double sum = 0.0;
#pragma unroll
        for (int b = 0; b < 5; ++b) sum += f_prev[b] * rr[b];
        if (sum > 0.0) {
          prod *= sum; m_sum += M;
          if (prod < 1e-200) { log_sum += log10(prod); prod = 1.0; }
        }
      }
      ll_fit[pass] = group_sum(log_sum + log10(prod) + m_sum, g.mask);
      if (pass == 1) break;
      iterations = it > 50 ? 50 : it;
      uint32_t mj = 0;
#pragma unroll
      for (int b = 1; b < 5; ++b) if (f[b] > f[mj]) mj = b;
      major = f[mj] > 0.0 ? mj : 5;
#pragma unroll
      for (int b = 0; b < 5; ++b) {
        if (fabs(f[b] - thr) <= 1e-9 * thr) recheck = true;  // may land on the other side on the host
        if (f[b] < thr) continue;
        if ((uint32_t)b != major && (minor == 5 || f[b] > f[minor])) minor = b;
        if ((uint32_t)b != ref && (variant == 5 || f[b] > f[variant])) variant = b;
      }
      if (variant == 5) break;
      allowed = 0x1F & ~(1u << variant);
    }
    const double variant_score = variant != 5 ? (ll_fit[0] - ll_fit[1]) - p.log10_ref_length : nan;
    const double slack = 1e-6;
    bool emit = false;
    if (best != ref && consensus > -slack) emit = true;
    if (variant != 5 && variant_score >= p.polymorphism_cutoff - slack) emit = true;
    bits = (bits & ~(0xFFFu | CO_EMIT | CO_RECHECK | (0xFFu << 16))) | best | (major << 3) | (minor << 6) | (variant << 9) | (iterations << 16) | CO_FIT;
    if (emit) bits |= CO_EMIT;
    if (recheck) bits |= CO_RECHECK;
    if (g.sub == 0) {
      out[slot].variant_score = variant_score;
      out[slot].bits = bits;
      if (emit || recheck) { const uint32_t k = atomicAdd(&scalars[1], 1u); if (k < flagged_cap) flagged[k] = slot; }
    }
    __syncwarp(g.mask);
  }
}

void launch_score_slots(const uint32_t* rec, const uint64_t* off, const uint32_t* cnt, const uint64_t* round_off,
                        const uint32_t* side, const uint32_t* side_off, const uint2* round_side,
                        const uint8_t* slot_ref, const uint32_t* round_slot, uint64_t n_rounds, uint64_t n_slots, uint64_t n_records,
                        const ClassTerms* lut, const double* tallyT, const HotTerms* coldT, const HotRatios* hotR, const ScoreParams& p,
                        ColumnOut* out, uint32_t* worklist, uint32_t* survivors, uint32_t* flagged, uint32_t* scalars, uint32_t flagged_cap,
                        uint32_t side_stride, cudaStream_t s, cudaEvent_t between) {
  if (!n_slots) return;
  const int kSMs = 148;
  const size_t smem_fit = (size_t)p.n_hot * 48 + (size_t)(FIT_TPB / FIT_LANES) * 2 * FIT_HASH * 4;
  // histogram block per warp: 4 KB holds 32 words per lane, 8 KB the maximum of 64; as many warps as 227 KB allow
  const uint32_t hist_block = p.t_nw <= 32 ? 4096u : 8192u;
  const size_t per_warp = hist_block + RING * 1024, fixed = (size_t)5 * p.t_stride + hist_block;
  int warps = TALLY_MAX_TPB / 32;
  while (warps > 1 && fixed + warps * per_warp > 227 * 1024) --warps;
  const size_t smem_tally = fixed + warps * per_warp;
  const int blocks = (int)std::min<uint64_t>((n_rounds + warps - 1) / warps, (uint64_t)kSMs);
  (void)n_records;
  static const uint32_t pf_vec = [] { const char* e = getenv("BRQ_TALLY_PREFETCH"); return e ? (uint32_t)atoi(e) : PREFETCH_VECTORS; }();  // -1: no L2 prefetch
  cudaFuncSetAttribute(tally_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tally);
  tally_kernel<<<blocks, warps * 32, smem_tally, s>>>(rec, round_off, side, round_side, round_slot, n_rounds, tallyT, coldT, p, out, worklist, flagged, scalars, flagged_cap, hist_block, side_stride, pf_vec);
  if (between) cudaEventRecord(between, s);
  // the screen (one pass of likelihood bounds per work-list slot) in front of the fit, where the stream has a shared table
  static const bool no_screen = getenv("BRQ_NO_SCREEN") != nullptr;
  const size_t smem_screen = (size_t)p.n_hot * 5 * 4 + (size_t)(SCREEN_TPB / 32) * 5 * p.t_nsq * 4 + 16;
  const bool screen = !no_screen && !p.fit_all && p.n_hot != 0 && side_stride == 1 && smem_screen <= 200 * 1024;
  if (screen) {
    cudaFuncSetAttribute(screen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_screen);
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (200 * 1024) / smem_screen));
    screen_kernel<<<kSMs * per_sm, SCREEN_TPB, smem_screen, s>>>(rec, off, cnt, side, side_off, slot_ref, worklist, lut, hotR, p, out, survivors, scalars);
    note_launches(1);
  }
  cudaFuncSetAttribute(fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fit);
  fit_kernel<<<kSMs * 3, FIT_TPB, smem_fit, s>>>(rec, off, cnt, side, side_off, slot_ref, screen ? survivors : worklist, screen ? scalars + 4 : scalars + 2,
                                                 lut, hotR, p, out, flagged, scalars, flagged_cap, side_stride);
  note_launches(2);
}

// ------------------------------------------------------------------------------------------ walk events
// the 8-byte walk record of a column, from the sums of its full result (the arithmetic of position_coverage::sum(),
// identify_mutations.h:115-120)
__device__ __forceinline__ WalkOut walk_of(const ColumnOut& co) {
  const uint32_t unique = co.unique[0] + co.unique[1];
  const double red = co.redundant[0] + co.redundant[1];
  const uint32_t total = unique + (uint32_t)(int)::round(red);
  return WalkOut{unique, total << 2 | (red > 0.0 ? 2u : 0u) | ((co.bits & CO_BASE_PREDICTED) ? 1u : 0u)};
}

__global__ void walk_mark_kernel(const ColumnOut* __restrict__ cols, WalkOut* __restrict__ walk, uint64_t n_base, const uint32_t* __restrict__ seg_first,
                                 const uint32_t* __restrict__ seg_last, const double* __restrict__ seg_prop, uint32_t n_seg,
                                 uint8_t* __restrict__ mark) {
  const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_base) return;
  uint32_t lo = 0, hi = n_seg;  // the segment that holds slot c
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (seg_first[mid] <= c) lo = mid; else hi = mid; }
  const double prop = seg_prop[lo];
  const WalkOut w = walk_of(cols[c]);
  walk[c] = w;
  const bool edge = c == seg_first[lo] || c == seg_last[lo];
  mark[c] = prop >= 0.0 && (edge || (double)w.unique <= prop || !(w.packed & 1u));
}

__global__ void walk_mark_flagged_kernel(const uint32_t* __restrict__ flagged, const uint32_t* __restrict__ n_flagged, uint32_t cap,
                                         uint64_t n_base, const uint64_t* __restrict__ ins_parent, uint8_t* __restrict__ mark) {
  const uint32_t n = min(*n_flagged, cap);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t slot = flagged[i];
    mark[slot < n_base ? (uint64_t)slot : ins_parent[slot - n_base]] = 1;  // an insert sub-column's RA row follows its parent column
  }
}

__global__ void walk_compact_kernel(const WalkOut* __restrict__ walk, const uint8_t* __restrict__ mark, uint64_t n_base,
                                    WalkEvent* __restrict__ events, uint32_t* __restrict__ counter) {
  const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_base) return;
  const bool keep = mark[c] || (c > 0 && mark[c - 1]) || (c + 1 < n_base && mark[c + 1]);
  if (keep) events[atomicAdd(counter, 1u)] = WalkEvent{(uint32_t)c, walk[c]};
}

void launch_walk_events(const ColumnOut* cols, WalkOut* walk, uint64_t n_base, const uint32_t* seg_first, const uint32_t* seg_last,
                        const double* seg_prop, uint32_t n_seg, const uint32_t* flagged, const uint32_t* n_flagged,
                        uint32_t flagged_cap, const uint64_t* ins_parent, uint8_t* mark, WalkEvent* events, uint32_t* counter,
                        cudaStream_t s) {
  if (!n_base || !n_seg) return;
  const uint32_t blocks = (uint32_t)((n_base + 255) / 256);
  walk_mark_kernel<<<blocks, 256, 0, s>>>(cols, walk, n_base, seg_first, seg_last, seg_prop, n_seg, mark);
  walk_mark_flagged_kernel<<<64, 256, 0, s>>>(flagged, n_flagged, flagged_cap, n_base, ins_parent, mark);
  walk_compact_kernel<<<blocks, 256, 0, s>>>(walk, mark, n_base, events, counter);
  note_launches(3);
}

__global__ void gather_columns_kernel(const ColumnOut* __restrict__ cols, const uint32_t* __restrict__ slots, uint32_t n, ColumnOut* __restrict__ out) {
  // six 16-byte pieces per result
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 6u) return;
  reinterpret_cast<uint4*>(out)[i] = reinterpret_cast<const uint4*>(cols + slots[i / 6u])[i % 6u];
}

void launch_gather_columns(const ColumnOut* cols, const uint32_t* slots, uint32_t n, ColumnOut* out, cudaStream_t s) {
  if (!n) return;
  gather_columns_kernel<<<(n * 6u + 255u) / 256u, 256, 0, s>>>(cols, slots, n, out);
  note_launches(1);
}

}  // namespace brq
