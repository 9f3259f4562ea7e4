// Per-slot scoring for sm_100a: a streaming tally kernel and an EM fit kernel.
//
//   tally_kernel  G lanes per slot (G = 4 for ordinary depth, 32 for deep columns).  A group walks its
//                 slot's records with 128-bit loads, G consecutive vectors per step, so a warp reads
//                 32/G runs of G*16 contiguous bytes instead of 32 scattered ones.  Coverage tallies
//                 and the five log-likelihood sums live in registers and are reduced over the group
//                 with a log2(G)-step butterfly (identify_mutations.cpp:1392-1658, 3398-3433).  Each
//                 group handles G consecutive slots one after the other and lane j keeps slot j's
//                 sums, so the per-slot closing arithmetic (consensus call, emission test, the 96-byte
//                 result) runs once per lane on 32 different slots.
//   likelihood table  The kernel is bound by shared-memory wavefronts: every scoring record reads its
//                 class's {L[0..4], M} (48 B, three 128-bit loads).  The table of the dominant MAPQ is
//                 therefore kept in three 16-byte planes with EIGHT interleaved copies; lane l reads
//                 copy l & 7, whose 16-byte cell lies in bank group l & 7, so the eight lanes served by
//                 one wavefront never collide whatever classes they ask for.  The copies have to fit
//                 in 227 KB: the table covers a window of quality values chosen from the stream's
//                 quality histogram (all of them when they fit).  The loop is branch-free: a record
//                 that does not score reads an all-zero cell.  Scoring records outside the shared
//                 table (another MAPQ, a '.' observation, a quality outside the window) wait in a
//                 three-entry register queue and read the full table from global memory when the
//                 slot is done; a lane that meets more of them rescans its share of the slot.
//   record ring   With one CTA of 16 warps per SM (the table fills shared memory), plain loads leave too few
//                 bytes in flight to cover DRAM latency.  Every lane therefore streams its 128-bit vectors
//                 through a private four-stage ring in shared memory with cp.async (LDGSTS): four vectors
//                 per lane, 32 KB per SM, are always on their way without holding registers, and the fetch
//                 cursor runs ahead across slot boundaries.  A lane only ever reads its own ring cells,
//                 so cp.async.wait_group is the only synchronisation.
//   redundant records  lead each slot's run (staging.cpp): their order-dependent sum of 1/X1
//                 (identify_mutations.cpp:1605) is a short sequential walk of the slot's head.
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
//   fit_kernel    one warp per work-list slot: the 5-allele EM fit, the presence score of the
//                 top non-reference allele (second EM with it held out), emission flags
//                 (identify_mutations.cpp:1797-1821, 3240-3344).
#include "kernels.h"
#include "brq_types.h"

namespace brq {

void note_launches(int n);

namespace {

constexpr int TALLY_TPB = 512;
constexpr int RING = (int)(TALLY_RING_BYTES / (TALLY_TPB * 16));  // 16-byte stages of each lane's record ring
static_assert(RING == 4, "the ring indexing below assumes four stages");
constexpr int FIT_TPB = 256;
constexpr int FIT_LANES = 32;     // lanes cooperating on one slot: the work list is short, so a slot's latency is what counts
constexpr int FIT_CACHE = 256;    // records per slot whose table code is cached in shared memory
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
__device__ __forceinline__ uint4 ldg_stream_u32x4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
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
__device__ __forceinline__ void sts_fill16(uint32_t shared_addr, uint32_t word) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" :: "r"(shared_addr), "r"(word) : "memory");
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
__device__ __forceinline__ uint32_t cold_index(uint32_t r, const ScoreParams& p, const uint8_t* mapq_slot) {
  const uint32_t hi = (r >> 10) & 63, mapq = (r >> SR_MAPQ_SHIFT) & 255;
  return ((hi * p.n_mapq_slots + mapq_slot[mapq]) * p.max_qual + ((r >> SR_QUAL_SHIFT) & 127)) * 5 + (r & 7);
}

template <int G>
__device__ __forceinline__ double group_add(double v, uint32_t mask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}
template <int G>
__device__ __forceinline__ uint32_t group_add_u32(uint32_t v, uint32_t mask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

struct Sums { double l0, l1, l2, l3, l4, m; };

// a scoring record whose class is not in the shared table (classic word from the side list): {L[0..4], M} from the global table
__device__ __forceinline__ void cold_add(Sums& a, uint32_t r, const HotTerms* __restrict__ coldT, const ScoreParams& p) {
  const uint32_t st = (r >> 10) & 63u, mapq = (r >> SR_MAPQ_SHIFT) & 255u, qual = (r >> SR_QUAL_SHIFT) & 127u;
  const char* e = reinterpret_cast<const char*>(coldT + (((st * p.n_mq + (mapq - p.mq_min)) * p.max_qual + qual) * 5u + (r & 7u)));
  const f64x2 x = ldg_f64x2(e), y = ldg_f64x2(e + 16), z = ldg_f64x2(e + 32);
  a.l0 += x.x; a.l1 += x.y; a.l2 += y.x; a.l3 += y.y; a.l4 += z.x; a.m += z.y;
}

}  // namespace

// ------------------------------------------------------------------------------------------ tally
template <int G>
__global__ void __launch_bounds__(TALLY_TPB, 1) tally_kernel(const uint32_t* __restrict__ rec, const uint64_t* __restrict__ off,
                                                              const uint32_t* __restrict__ side, const uint32_t* __restrict__ side_off,
                                                              const uint8_t* __restrict__ slot_ref, uint64_t n_slots,
                                                              const double* __restrict__ tallyT, const HotTerms* __restrict__ coldT,
                                                              ScoreParams p, ColumnOut* __restrict__ out, uint32_t* __restrict__ worklist,
                                                              uint32_t* __restrict__ flagged, uint32_t* __restrict__ scalars,
                                                              uint32_t flagged_cap) {
  // three planes ({L0,L1} {L2,L3} {L4,M}) of (t_nhot + 1) * t_copies 16-byte cells; the last cell of a plane is zero
  extern __shared__ __align__(16) double sm[];
  __shared__ double inv_red[64];
  {
    const uint32_t n_cells = 3u * (p.t_nhot + 1u) * p.t_copies;
    const double2* src = reinterpret_cast<const double2*>(tallyT);
    double2* dst = reinterpret_cast<double2*>(sm);
    for (uint32_t i = threadIdx.x; i < n_cells; i += blockDim.x) dst[i] = src[i];
    if (threadIdx.x < 64) inv_red[threadIdx.x] = 1.0 / (double)threadIdx.x;
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u, sub = lane & (uint32_t)(G - 1), g0 = lane - sub;
  const uint32_t cs = p.t_copies * 16u;                      // bytes between consecutive classes
  const uint32_t plane = (p.t_nhot + 1u) * cs;
  const uint32_t tbl = (uint32_t)__cvta_generic_to_shared(sm) + (lane & (p.t_copies - 1u)) * 16u;  // this lane's copy
  const uint32_t zero_addr = tbl + p.t_nhot * cs;
  const uint32_t ring = (uint32_t)__cvta_generic_to_shared(sm) + 3u * plane + threadIdx.x * 16u;  // this lane's cell of stage 0
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  const uint32_t gmask = G == 32 ? 0xFFFFFFFFu : (((1u << G) - 1u) << g0);
  const uint64_t n_rounds = (n_slots + 31) >> 5;  // a warp takes 32 consecutive slots per round
  const uint64_t n_warps = (uint64_t)gridDim.x * (TALLY_TPB / 32);
  for (uint64_t round = ((uint64_t)blockIdx.x * TALLY_TPB + threadIdx.x) >> 5; round < n_rounds; round += n_warps) {
    const uint64_t my_slot = (round << 5) + lane;  // the slot this lane closes; its group tallies slots g0 .. g0+G-1
    uint64_t my_beg = 0;
    uint32_t my_vec = 0, my_pad = 0, my_ref = 5;   // 128-bit vectors of the run, pad words in the last one
    uint32_t my_side0 = 0, my_side1 = 0;           // the slot's range of the side list
    if (my_slot < n_slots) {
      const uint64_t o0 = off[my_slot], o1 = off[my_slot + 1];
      my_beg = o0 & ~3ull; my_vec = (uint32_t)(((o1 & ~3ull) - my_beg) >> 2); my_pad = (uint32_t)o1 & 3u;
      my_ref = slot_ref[my_slot];
      my_side0 = side_off[my_slot]; my_side1 = side_off[my_slot + 1];
    }
    Sums kept = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    double red_top = 0.0, red_bot = 0.0;
    uint32_t tops = 0, n = 0, c_ref = 0, raw_top = 0, raw_bot = 0;
    // the records of this warp's NEXT round are pulled into L2 while this round is tallied: with 16
    // warps per SM the loads below cannot cover DRAM latency on their own
    uint64_t pf_lo = 0, pf_hi = 0;
    if (lane == 0 && round + n_warps < n_rounds) {
      const uint64_t s0 = (round + n_warps) << 5, s1 = s0 + 32 < n_slots ? s0 + 32 : n_slots;
      pf_lo = off[s0] & ~3ull; pf_hi = off[s1] & ~3ull;
    }

    // fetch cursor of this lane's vector stream: slot fk of the group, step f_it of f_nit, running RING vectors ahead
    int fk = -1;
    uint32_t f_it = 0, f_nit = 0, f_nvec = 0;
    const uint4* f_vp = nullptr;
    auto fetch_next_slot = [&]() {  // group-uniform
      f_it = 0; f_nit = 0;
      while (f_nit == 0 && ++fk < G) {
        f_nvec = __shfl_sync(gmask, my_vec, g0 + fk);
        f_vp = reinterpret_cast<const uint4*>(rec + __shfl_sync(gmask, my_beg, g0 + fk));
        f_nit = (f_nvec + (uint32_t)G - 1u) / (uint32_t)G;
      }
    };
    auto fetch = [&](uint32_t stage) {  // the next vector of the stream goes to ring[stage]; always one commit
      if (fk < G) {
        const uint32_t iv = f_it * (uint32_t)G + sub, dst = ring + stage * (uint32_t)(TALLY_TPB * 16);
        if (iv < f_nvec) cp_async16(dst, f_vp + iv); else sts_fill16(dst, p.t_nhot);  // past the run: pad words
        if (++f_it == f_nit) fetch_next_slot();
      }
      cp_async_commit();
    };
    fetch_next_slot();
#pragma unroll
    for (int st = 0; st < RING; ++st) fetch((uint32_t)st);
    uint32_t c_idx = 0;  // vectors consumed by this lane in this round

    // While the first vectors are on their way: the scoring records of this lane's own slot whose class
    // is not in the shared table (another MAPQ, a '.' observation, a quality outside the window) sit in
    // the side list as classic words; their terms come from the global table and start the slot's sums.
    for (uint32_t e = my_side0; e < my_side1; ++e) {
      const uint32_t w = __ldg(side + e);
      if (w & SIDE_BIG) continue;  // the X1 of a very redundant record, read by the walk below
      cold_add(kept, w, coldT, p);
      ++n;
      c_ref += (w >> 27) & 1u;
    }

#pragma unroll 1
    for (int k = 0; k < G; ++k) {
      const uint32_t ref = __shfl_sync(gmask, my_ref, g0 + k);
      const uint64_t beg = __shfl_sync(gmask, my_beg, g0 + k);
      const uint32_t n_vec = __shfl_sync(gmask, my_vec, g0 + k);
      const uint4* vp = reinterpret_cast<const uint4*>(rec + beg);
      double rt = 0.0, rb = 0.0;
      uint32_t t_rawt = 0, t_rawb = 0;
      Sums a = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      uint32_t t_tops = 0, t_n = 0, t_cref = 0;
      uint32_t acc_tm = 0, acc_h = 0;  // packed counters: tops in [21:12] and matches in [31:22]; hot records in [31:23]
      uint32_t side_cur = __shfl_sync(gmask, my_side0, g0 + k);
      const uint32_t n_it = (n_vec + (uint32_t)G - 1u) / (uint32_t)G;  // the same for every lane of the group
      // four device words of one 128-bit vector: the cell index addresses this lane's copy of the table
      // (every word that does not score, pad words included, carries the zero cell) and two masked adds
      // keep the counts
      auto tally4 = [&](const uint4& v) {
        const uint32_t r[4] = {v.x, v.y, v.z, v.w};
        uint32_t addr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          addr[j] = tbl + (r[j] & DR_CELL_MASK) * cs;
          acc_tm += r[j] & (DR_TOP_BIT | DR_MATCH_BIT);
          acc_h += r[j] & DR_HOT_BIT;
        }
        f64x2 x[4], y[4], z[4];  // {L0,L1}, {L2,L3}, {L4,M}: all twelve 128-bit loads in flight together
#pragma unroll
        for (int j = 0; j < 4; ++j) { x[j] = lds_f64x2(addr[j]); y[j] = lds_f64x2(addr[j] + plane); z[j] = lds_f64x2(addr[j] + 2u * plane); }
        // the zero cell adds +0.0 exactly; pairs first, so each sum waits on two dependent adds per vector, not four
        a.l0 += (x[0].x + x[1].x) + (x[2].x + x[3].x); a.l1 += (x[0].y + x[1].y) + (x[2].y + x[3].y);
        a.l2 += (y[0].x + y[1].x) + (y[2].x + y[3].x); a.l3 += (y[0].y + y[1].y) + (y[2].y + y[3].y);
        a.l4 += (z[0].x + z[1].x) + (z[2].x + z[3].x); a.m += (z[0].y + z[1].y) + (z[2].y + z[3].y);
      };
      auto flush_counts = [&]() {
        t_tops += (acc_tm >> 12) & 0x3FFu; t_cref += acc_tm >> 22; t_n += acc_h >> 23;
        acc_tm = 0; acc_h = 0;
      };
      for (uint32_t it = 0; it < n_it; ++it, ++c_idx) {
        const uint32_t stage = c_idx & (uint32_t)(RING - 1);
        cp_async_wait<RING - 1>();  // this lane's oldest vector has landed
        const uint4 v = lds_u32x4(ring + stage * (uint32_t)(TALLY_TPB * 16));
        fetch(stage);               // the cell is free again: request the vector RING steps ahead
        if (it == 0) {
          // redundant records lead the slot: an order-dependent double sum, taken in arrival order by
          // every lane of the group alike (identify_mutations.cpp:1605).  The first vector is in the
          // group's first lane; a pad word ends the walk like any non-redundant record does.
          const uint32_t hv[4] = {__shfl_sync(gmask, v.x, g0), __shfl_sync(gmask, v.y, g0), __shfl_sync(gmask, v.z, g0),
                                  __shfl_sync(gmask, v.w, g0)};
          const uint32_t cnt = n_vec * 4u;
          uint32_t i = 0, r = hv[0];
          while ((r >> DR_KIND_SHIFT) == 3u) {
            uint32_t red = (r >> DR_X1_SHIFT) & DR_X1_MASK;
            if (red == DR_X1_MASK) red = __ldg(side + side_cur++) & ~SIDE_BIG;
            const double inv = red < 64 ? inv_red[red] : 1.0 / (double)red;
            if (r & DR_TOP_BIT) { rt += inv; ++t_rawt; } else { rb += inv; ++t_rawb; }
            if (++i == cnt) break;
            r = i == 1 ? hv[1] : i == 2 ? hv[2] : i == 3 ? hv[3] : __ldg(rec + beg + i);
          }
        }
        tally4(v);
        if ((it & 63u) == 63u) flush_counts();  // 256 records per lane: the packed fields hold 511
      }
      flush_counts();
      a.l0 = group_add<G>(a.l0, gmask); a.l1 = group_add<G>(a.l1, gmask); a.l2 = group_add<G>(a.l2, gmask);
      a.l3 = group_add<G>(a.l3, gmask); a.l4 = group_add<G>(a.l4, gmask); a.m = group_add<G>(a.m, gmask);
      t_tops = group_add_u32<G>(t_tops, gmask); t_n = group_add_u32<G>(t_n, gmask); t_cref = group_add_u32<G>(t_cref, gmask);
      if (sub == (uint32_t)k) {
        kept.l0 += a.l0; kept.l1 += a.l1; kept.l2 += a.l2; kept.l3 += a.l3; kept.l4 += a.l4; kept.m += a.m;
        red_top = rt; red_bot = rb;
        tops = t_tops; n += t_n; c_ref += t_cref; raw_top = t_rawt; raw_bot = t_rawb;
      }
      if (k == 0 && pf_hi > pf_lo) prefetch_l2_bulk(rec + pf_lo, (uint32_t)((pf_hi - pf_lo) * 4u));
    }
    if (my_slot >= n_slots) continue;

    // every record of the slot is either unique or redundant; top-strand counts follow by subtraction
    const uint32_t cnt = my_vec * 4u - my_pad;
    const uint32_t u_all = cnt - raw_top - raw_bot, u_top = tops - raw_top;
    const uint32_t ref = my_ref;
    const double ll[5] = {kept.l0, kept.l1, kept.l2, kept.l3, kept.l4};
    double consensus = nan;
    uint32_t best = 5;
    bool need_fit = false;
    const double slack = 1e-6;
    if (n > 0) {  // pure_genotype_call, identify_mutations.cpp:3398-3433
      best = 0;
#pragma unroll
      for (int b = 1; b < 5; ++b) if (ll[b] > ll[best]) best = b;
      double offv = -1.7976931348623157e308;
#pragma unroll
      for (int b = 0; b < 5; ++b) if ((uint32_t)b != best) offv = fmax(offv, ll[b]);
      double tot = 0.0;
#pragma unroll
      for (int b = 0; b < 5; ++b) {
        if ((uint32_t)b == best) continue;
        const double d = ll[b] - offv;
        // the runner-up itself contributes exactly 1; anything below 2^-54 cannot change that sum
        tot += (d == 0.0) ? 1.0 : (d < -17.0 ? 0.0 : pow(10.0, d));
      }
      consensus = (ll[best] - (log10(tot) + offv)) - p.log10_ref_length;
      // an RA row needs best != ref with a positive consensus score, or a presence score at the cutoff
      need_fit = p.fit_all != 0u || ref >= 5u || (best != ref && consensus > -slack);
      if (!need_fit) {
        const double ll_ref = ref == 0 ? ll[0] : ref == 1 ? ll[1] : ref == 2 ? ll[2] : ref == 3 ? ll[3] : ll[4];
        const double bound = (kept.m - ll_ref) - (double)n * log10(((double)c_ref + 0.5) / ((double)n + 2.0)) - p.log10_ref_length;
        need_fit = !(bound < p.polymorphism_cutoff - slack);
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
    o.consensus_score = consensus; o.variant_score = nan;
    o.redundant[0] = red_bot; o.redundant[1] = red_top;
    o.unique[0] = u_all - u_top; o.unique[1] = u_top; o.raw_redundant[0] = raw_bot; o.raw_redundant[1] = raw_top;
    o.n = n; o.bits = bits;
    out[my_slot] = o;

    if (need_fit) worklist[atomicAdd(&scalars[2], 1u)] = (uint32_t)my_slot;
    else if (recheck) { const uint32_t kf = atomicAdd(&scalars[1], 1u); if (kf < flagged_cap) flagged[kf] = (uint32_t)my_slot; }
  }
}

// ------------------------------------------------------------------------------------------ fit
namespace {

struct GroupCtx {
  const uint32_t* rec; uint64_t beg, end;   // the slot's device words [beg, beg + n_main) followed, in index space, by its side-list entries
  const uint32_t* side; uint32_t side_beg; uint64_t n_main;
  uint32_t hot_base; const ClassTerms* lut; const uint8_t* mapq_slot; const ScoreParams* p;
  const uint32_t* cache;  // table codes of the slot's first FIT_CACHE records
  uint32_t sub, mask;     // lane within the group, shuffle mask of the group
};

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
__device__ __forceinline__ uint32_t code_of(const GroupCtx& g, uint32_t r) {
  const ScoreParams& p = *g.p;
  if (!eligible(r, p.base_quality_cutoff)) return CODE_NONE;
  if (p.n_hot && ((r >> SR_MAPQ_SHIFT) & 255) == p.hot_mapq) return hot_index(r, p.max_qual) * 48u;
  return CODE_COLD | cold_index(r, p, g.mapq_slot);
}
// Classic word at index i of the slot's records for the fit: a HOT device word is decoded, a side-list
// entry is taken as it is, everything else (IDLE, the COLD placeholder, REDUNDANT, padding) does not score.
__device__ __forceinline__ uint32_t classic_at(const GroupCtx& g, uint64_t i) {
  const uint64_t k = i - g.beg;
  if (k < g.n_main) {
    const uint32_t d = __ldg(g.rec + i);
    if ((d >> DR_KIND_SHIFT) != 0u || !(d & DR_HOT_BIT)) return 0u;
    const ScoreParams& p = *g.p;
    const uint32_t cell = d & DR_CELL_MASK, obs = cell & 3u, t = cell >> 2, qual = p.t_qlo + t % p.t_nq, st = t / p.t_nq;
    return obs | qual << SR_QUAL_SHIFT | st << 10 | p.hot_mapq << SR_MAPQ_SHIFT | SR_UNIQUE_BIT | SR_OK_BIT;
  }
  const uint32_t w = __ldg(g.side + g.side_beg + (uint32_t)(k - g.n_main));
  return (w & SIDE_BIG) ? 0u : w;
}
__device__ __forceinline__ uint32_t code_at(const GroupCtx& g, uint64_t i) {
  const uint64_t k = i - g.beg;
  return k < FIT_CACHE ? g.cache[k] : code_of(g, classic_at(g, i));
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

__global__ void __launch_bounds__(FIT_TPB, 3) fit_kernel(const uint32_t* __restrict__ rec, const uint64_t* __restrict__ off,
                                                          const uint32_t* __restrict__ side, const uint32_t* __restrict__ side_off,
                                                          const uint8_t* __restrict__ slot_ref, const uint32_t* __restrict__ worklist,
                                                          const ClassTerms* __restrict__ lut, const HotRatios* __restrict__ hotR,
                                                          ScoreParams p, ColumnOut* __restrict__ out, uint32_t* __restrict__ flagged,
                                                          uint32_t* __restrict__ scalars, uint32_t flagged_cap) {
  extern __shared__ __align__(16) double sm[];
  __shared__ uint8_t mapq_slot[256];
  __shared__ uint32_t cache[FIT_TPB / FIT_LANES][FIT_CACHE];
  const uint32_t n_work = scalars[2];
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
  uint32_t* my_cache = cache[threadIdx.x / FIT_LANES];
  g.cache = my_cache;
  for (;;) {
    uint32_t w = 0;
    if (g.sub == 0) w = atomicAdd(&scalars[3], 1u);  // slots are handed out one at a time: EM lengths vary widely
    w = __shfl_sync(g.mask, w, lane - g.sub);
    if (w >= n_work) break;
    const uint32_t slot = worklist[w];
    score_slot_range(off, slot, g.beg, g.end);
    g.n_main = g.end - g.beg;
    g.side = side; g.side_beg = side_off[slot];
    g.end += side_off[slot + 1] - g.side_beg;
    uint32_t obs_count[5] = {0, 0, 0, 0, 0}, n = 0;
    for (uint64_t i = g.beg + g.sub; i < g.end; i += FIT_LANES) {
      const uint32_t r = classic_at(g, i);
      const uint32_t code = code_of(g, r);
      if (i - g.beg < FIT_CACHE) my_cache[i - g.beg] = code;
      if (code == CODE_NONE) continue;
#pragma unroll
      for (int b = 0; b < 5; ++b) obs_count[b] += ((r & 7) == (uint32_t)b);
      ++n;
    }
    __syncwarp(g.mask);
#pragma unroll
    for (int b = 0; b < 5; ++b) obs_count[b] = group_sum_u32(obs_count[b], g.mask);
    n = group_sum_u32(n, g.mask);
    if (n == 0) continue;
    const uint32_t ref = slot_ref[slot];
    const double consensus = out[slot].consensus_score;
    uint32_t bits = out[slot].bits;
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

void launch_score_slots(const uint32_t* rec, const uint64_t* off, const uint32_t* side, const uint32_t* side_off,
                        const uint8_t* slot_ref, uint64_t n_slots, uint64_t n_records,
                        const ClassTerms* lut, const double* tallyT, const HotTerms* coldT, const HotRatios* hotR, const ScoreParams& p,
                        ColumnOut* out, uint32_t* worklist, uint32_t* flagged, uint32_t* scalars, uint32_t flagged_cap,
                        cudaStream_t s, cudaEvent_t between) {
  if (!n_slots) return;
  const int kSMs = 148;
  const size_t smem_tally = (size_t)3 * (p.t_nhot + 1) * p.t_copies * 16 + TALLY_RING_BYTES, smem_fit = (size_t)p.n_hot * 48;
  const uint64_t n_rounds = (n_slots + 31) / 32;
  const int blocks = (int)std::min<uint64_t>((n_rounds + TALLY_TPB / 32 - 1) / (TALLY_TPB / 32), (uint64_t)kSMs);
  // lanes per slot: 4 at ordinary depth, a whole warp once the mean column is deeper than 512 records
  if (n_records / n_slots < 512) {
    cudaFuncSetAttribute(tally_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tally);
    tally_kernel<4><<<blocks, TALLY_TPB, smem_tally, s>>>(rec, off, side, side_off, slot_ref, n_slots, tallyT, coldT, p, out, worklist, flagged, scalars, flagged_cap);
  } else {
    cudaFuncSetAttribute(tally_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tally);
    tally_kernel<32><<<blocks, TALLY_TPB, smem_tally, s>>>(rec, off, side, side_off, slot_ref, n_slots, tallyT, coldT, p, out, worklist, flagged, scalars, flagged_cap);
  }
  if (between) cudaEventRecord(between, s);
  cudaFuncSetAttribute(fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fit);
  fit_kernel<<<kSMs * 3, FIT_TPB, smem_fit, s>>>(rec, off, side, side_off, slot_ref, worklist, lut, hotR, p, out, flagged, scalars, flagged_cap);
  note_launches(2);
}

}  // namespace brq
