// Per-slot scoring for sm_100a: a streaming tally kernel and an EM fit kernel.
//
//   tally_kernel  one thread per slot (a reference column or an insert sub-column).  The thread
//                 walks its records in arrival order with 128-bit loads, so coverage tallies and
//                 the five log-likelihood sums live in registers and nothing is reduced across
//                 lanes (identify_mutations.cpp:1392-1658, 3398-3433).  The hot path is branch-free:
//                 a record that does not score reads an all-zero table entry.  A slot whose scoring
//                 records all show the reference base X, with X every record's best hypothesis and
//                 every other hypothesis bounded (see `pure`), is final here: the EM cannot lift any
//                 other allele to the half-read level.  All other non-empty slots go to a work list.
//   fit_kernel    eight lanes per work-list slot: the 5-allele EM fit, the presence score of the
//                 top non-reference allele (second EM with it held out), emission flags
//                 (identify_mutations.cpp:1797-1821, 3240-3344).
//
// Likelihood terms of the dominant MAPQ value are staged in shared memory (48 B per class, read as
// three 128-bit loads); records with any other MAPQ read the full table from global memory.
#include "kernels.h"
#include "brq_types.h"

namespace brq {

namespace {

constexpr int TALLY_TPB = 256;
constexpr int FIT_TPB = 256;
constexpr int FIT_LANES = 8;      // lanes cooperating on one slot
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

}  // namespace

// ------------------------------------------------------------------------------------------ tally
__global__ void __launch_bounds__(TALLY_TPB, 3) tally_kernel(const uint32_t* __restrict__ rec, const uint64_t* __restrict__ off,
                                                              const uint8_t* __restrict__ slot_ref, uint64_t n_slots,
                                                              const ClassTerms* __restrict__ lut, const HotTerms* __restrict__ hotL,
                                                              ScoreParams p, ColumnOut* __restrict__ out, uint32_t* __restrict__ worklist,
                                                              uint32_t* __restrict__ flagged, uint32_t* __restrict__ scalars,
                                                              uint32_t flagged_cap) {
  extern __shared__ __align__(16) double sm[];  // n_hot entries of 6 doubles, then one all-zero entry
  __shared__ uint8_t mapq_slot[256];
  __shared__ double inv_red[64];
  {
    const double* src = reinterpret_cast<const double*>(hotL);
    for (uint32_t i = threadIdx.x; i < p.n_hot * 6; i += blockDim.x) sm[i] = src[i];
    if (threadIdx.x < 6) sm[p.n_hot * 6 + threadIdx.x] = 0.0;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) mapq_slot[i] = p.mapq_slot[i];
    if (threadIdx.x < 64) inv_red[threadIdx.x] = 1.0 / (double)threadIdx.x;
  }
  __syncthreads();
  const uint32_t hot_base = (uint32_t)__cvta_generic_to_shared(sm);
  const uint32_t zero_addr = hot_base + p.n_hot * 48u;
  const uint32_t Q = p.max_qual, cutoff = p.base_quality_cutoff;
  // unique, untrimmed, resolvable -- and, for the shared table, the dominant MAPQ -- as masked compares
  const uint32_t flag_mask = SR_UNIQUE_BIT | SR_TRIM_BIT | SR_OK_BIT, flag_want = SR_UNIQUE_BIT | SR_OK_BIT;
  const uint32_t hot_mask = flag_mask | (255u << SR_MAPQ_SHIFT);
  // without a shared table no record can match: every scoring record takes the global path
  const uint32_t hot_want = p.n_hot ? (flag_want | (p.hot_mapq << SR_MAPQ_SHIFT)) : 0xFFFFFFFFu;

  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  const uint64_t stride = (uint64_t)gridDim.x * TALLY_TPB;
  for (uint64_t slot = (uint64_t)blockIdx.x * TALLY_TPB + threadIdx.x; slot < n_slots; slot += stride) {
    const uint64_t beg = off[slot], end = off[slot + 1];
    uint32_t tops = 0, raw_top = 0, raw_bot = 0, n = 0, obs_mask = 0;
    double red_top = 0.0, red_bot = 0.0, r2max = 0.0;
    double ll0 = 0.0, ll1 = 0.0, ll2 = 0.0, ll3 = 0.0, ll4 = 0.0;

    // 128-bit loads from the aligned vector that holds the slot's first record; all index math is
    // 32-bit and relative to the slot.  Elements outside the slot are zeroed: a zero record has no
    // flag set, scores nothing and reads the all-zero table entry.
    const uint32_t head = (uint32_t)beg & 3u, cnt = (uint32_t)(end - beg);
    const uint32_t n_vec = (head + cnt + 3u) >> 2;
    const uint4* vp = reinterpret_cast<const uint4*>(rec + (beg - head));
    uint4 cur = make_uint4(0, 0, 0, 0);
    if (n_vec) cur = __ldg(vp);
    for (uint32_t iv = 0; iv < n_vec; ++iv) {
      uint4 nxt = make_uint4(0, 0, 0, 0);
      if (iv + 1 < n_vec) nxt = __ldg(vp + iv + 1);  // requested before `cur` is consumed
      const uint32_t k0 = iv * 4u - head;              // slot-relative index of element 0 (wraps below zero)
      uint32_t r[4] = {cur.x, cur.y, cur.z, cur.w};
      uint32_t addr[4], n_flag_ok = 0, n_hot_here = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        r[j] = (k0 + (uint32_t)j < cnt) ? r[j] : 0u;
        const uint32_t qual = (r[j] >> SR_QUAL_SHIFT) & 127u;
        const bool qual_ok = qual >= cutoff;
        const bool flags_ok = (r[j] & flag_mask) == flag_want;
        const bool hotp = qual_ok && (r[j] & hot_mask) == hot_want;   // flags and MAPQ in one compare
        tops += (r[j] >> 10) & 1u;
        n_flag_ok += (flags_ok && qual_ok) ? 1u : 0u;
        n_hot_here += hotp ? 1u : 0u;
        const uint32_t e = ((((r[j] >> 10) & 63u) * Q + qual) * 5u + (r[j] & 7u)) * 48u;
        addr[j] = hotp ? hot_base + e : zero_addr;
        obs_mask |= hotp ? (1u << (r[j] & 7u)) : 0u;
      }
      n += n_hot_here;
      f64x2 a[4], b[4], c[4];  // L[0..1], L[2..3], {L[4], r2}: all twelve loads in flight together
#pragma unroll
      for (int j = 0; j < 4; ++j) { a[j] = lds_f64x2(addr[j]); b[j] = lds_f64x2(addr[j] + 16); c[j] = lds_f64x2(addr[j] + 32); }
#pragma unroll
      for (int j = 0; j < 4; ++j) {  // arrival order; the zero entry adds +0.0 exactly
        ll0 += a[j].x; ll1 += a[j].y; ll2 += b[j].x; ll3 += b[j].y; ll4 += c[j].x;
        r2max = fmax(r2max, c[j].y);
      }
      if (n_flag_ok != n_hot_here) {  // scoring records with another MAPQ: full table in global memory
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool cold = (r[j] & flag_mask) == flag_want && ((r[j] >> SR_QUAL_SHIFT) & 127u) >= cutoff && (r[j] & hot_mask) != hot_want;
          if (!cold) continue;
          const char* e = reinterpret_cast<const char*>(lut + cold_index(r[j], p, mapq_slot));
          const f64x2 x = ldg_f64x2(e), y = ldg_f64x2(e + 16), z = ldg_f64x2(e + 32);
          ll0 += x.x; ll1 += x.y; ll2 += y.x; ll3 += y.y; ll4 += z.x;
          r2max = fmax(r2max, z.y);
          obs_mask |= 1u << (r[j] & 7u);
          ++n;
        }
      }
      if (!(r[0] & r[1] & r[2] & r[3] & SR_UNIQUE_BIT)) {  // a redundant record (or a zeroed element) is present
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // order-dependent double sum, arrival order (identify_mutations.cpp:1605)
          if ((r[j] & SR_UNIQUE_BIT) || r[j] == 0u) continue;
          const uint32_t red = (r[j] >> SR_RED_SHIFT) & SR_RED_MASK;
          const double inv = red < 64 ? inv_red[red] : 1.0 / (double)red;
          if (r[j] & SR_TOP_BIT) { red_top += inv; ++raw_top; } else { red_bot += inv; ++raw_bot; }
        }
      }
      cur = nxt;
    }
    // every record of the slot is either unique or redundant; top-strand counts follow by subtraction
    const uint32_t u_all = cnt - raw_top - raw_bot, u_top = tops - raw_top;

    const uint32_t ref = slot_ref[slot];
    const double ll[5] = {ll0, ll1, ll2, ll3, ll4};
    double consensus = nan;
    uint32_t best = 5;
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
    }
    const double slack = 1e-6;
    const bool base_predicted = consensus >= p.mutation_cutoff;
    const bool recheck = n > 0 && fabs(consensus - p.mutation_cutoff) < slack;

    // `pure`: every scoring record shows the reference base X, X is each record's most likely true
    // base (r2 is +inf otherwise), and every other hypothesis b has r_i(b) <= f0[X] = (n+0.5)/(n+2.5).
    // Then s_i >= f[X] in every iteration, so f[b] only shrinks from 0.5/(n+2.5) < 0.5/n while f[X]
    // grows: the fit reports major = X and neither a minor nor a variant allele, and the reference
    // computes no presence score (identify_mutations.cpp:1806-1821).
    const bool pure = n > 0 && ref < 5 && obs_mask == (1u << ref) && r2max <= ((double)n + 0.5) / ((double)n + 2.5);
    uint32_t bits = best | ((pure ? ref : 5u) << 3) | (5u << 6) | (5u << 9);
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
    out[slot] = o;

    if (n > 0 && !pure) worklist[atomicAdd(&scalars[2], 1u)] = (uint32_t)slot;
    else if (recheck) { const uint32_t k = atomicAdd(&scalars[1], 1u); if (k < flagged_cap) flagged[k] = (uint32_t)slot; }
  }
}

// ------------------------------------------------------------------------------------------ fit
namespace {

struct GroupCtx {
  const uint32_t* rec; uint64_t beg, end;
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
__device__ __forceinline__ uint32_t code_at(const GroupCtx& g, uint64_t i) {
  const uint64_t k = i - g.beg;
  return k < FIT_CACHE ? g.cache[k] : code_of(g, __ldg(g.rec + i));
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
                                                          const uint8_t* __restrict__ slot_ref, const uint32_t* __restrict__ worklist,
                                                          const ClassTerms* __restrict__ lut, const HotRatios* __restrict__ hotR,
                                                          ScoreParams p, ColumnOut* __restrict__ out, uint32_t* __restrict__ flagged,
                                                          uint32_t* __restrict__ scalars, uint32_t flagged_cap) {
  extern __shared__ __align__(16) double sm[];
  __shared__ uint8_t mapq_slot[256];
  __shared__ uint32_t cache[FIT_TPB / FIT_LANES][FIT_CACHE];
  {
    const double* src = reinterpret_cast<const double*>(hotR);
    for (uint32_t i = threadIdx.x; i < p.n_hot * 6; i += blockDim.x) sm[i] = src[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) mapq_slot[i] = p.mapq_slot[i];
  }
  __syncthreads();
  const uint32_t n_work = scalars[2];
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  const uint32_t lane = threadIdx.x & 31;
  GroupCtx g;
  g.rec = rec; g.hot_base = (uint32_t)__cvta_generic_to_shared(sm); g.lut = lut; g.mapq_slot = mapq_slot; g.p = &p;
  g.sub = lane % FIT_LANES;
  g.mask = ((1u << FIT_LANES) - 1u) << (lane - g.sub);
  uint32_t* my_cache = cache[threadIdx.x / FIT_LANES];
  g.cache = my_cache;
  for (;;) {
    uint32_t w = 0;
    if (g.sub == 0) w = atomicAdd(&scalars[3], 1u);  // slots are handed out one at a time: EM lengths vary widely
    w = __shfl_sync(g.mask, w, lane - g.sub);
    if (w >= n_work) break;
    const uint32_t slot = worklist[w];
    g.beg = off[slot]; g.end = off[slot + 1];
    uint32_t obs_count[5] = {0, 0, 0, 0, 0}, n = 0;
    for (uint64_t i = g.beg + g.sub; i < g.end; i += FIT_LANES) {
      const uint32_t r = __ldg(rec + i);
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
    bits = (bits & ~(0xFFFu | CO_EMIT | CO_RECHECK | (0xFFu << 16))) | best | (major << 3) | (minor << 6) | (variant << 9) | (iterations << 16);
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

void launch_score_slots(const uint32_t* rec, const uint64_t* off, const uint8_t* slot_ref, uint64_t n_slots,
                        const ClassTerms* lut, const HotTerms* hotL, const HotRatios* hotR, const ScoreParams& p, ColumnOut* out,
                        uint32_t* worklist, uint32_t* flagged, uint32_t* scalars, uint32_t flagged_cap, cudaStream_t s,
                        cudaEvent_t between) {
  if (!n_slots) return;
  const int kSMs = 148;
  const size_t smem_tally = ((size_t)p.n_hot + 1) * 48, smem_fit = (size_t)p.n_hot * 48;
  cudaFuncSetAttribute(tally_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tally);
  cudaFuncSetAttribute(fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fit);
  int blocks = (int)std::min<uint64_t>((n_slots + TALLY_TPB - 1) / TALLY_TPB, (uint64_t)kSMs * 3);
  tally_kernel<<<blocks, TALLY_TPB, smem_tally, s>>>(rec, off, slot_ref, n_slots, lut, hotL, p, out, worklist, flagged, scalars, flagged_cap);
  if (between) cudaEventRecord(between, s);
  fit_kernel<<<kSMs * 3, FIT_TPB, smem_fit, s>>>(rec, off, slot_ref, worklist, lut, hotR, p, out, flagged, scalars, flagged_cap);
}

}  // namespace brq
