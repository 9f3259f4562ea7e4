// Per-slot scoring, thread-per-slot formulation (sm_100a).
//
// One thread owns one slot (a reference column or an insert sub-column) and walks its records in
// arrival order, so every per-slot quantity lives in registers, nothing is reduced across lanes,
// and floating-point sums accumulate in the same order as the reference's per-read loops
// (identify_mutations.cpp:1392-1658, 3240-3318).
//
//   tally_kernel  one HBM pass over the 4-byte records (128-bit loads): per-strand unique /
//                 redundant coverage, the five log-likelihood sums, the pure-genotype call.  Slots
//                 whose scoring records all show the reference base, with every record's likelihood
//                 ratio bounded so that no other allele can reach the half-read level, are final
//                 here: the EM fit cannot produce a variant for them (see `pure` below).  All other
//                 slots go to a work list.
//   fit_kernel    work-list slots only: the 5-allele EM, the presence score of the top
//                 non-reference allele (second EM with that allele held out), emission flags.
//
// The likelihood terms of the dominant MAPQ value are staged in shared memory (48 B per class);
// records with any other MAPQ read the full table from global memory.
#include "kernels.h"
#include "brq_types.h"

namespace brq {

namespace {

constexpr int TPB = 512;

__device__ __forceinline__ uint4 ld_stream_v4(const uint32_t* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ld_cached_v4(const uint32_t* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}

struct Decoded { uint32_t obs, qual, top, mapq, set; };
__device__ __forceinline__ Decoded decode(uint32_t r) {
  Decoded d;
  d.obs = r & 7; d.qual = (r >> SR_QUAL_SHIFT) & 127; d.top = (r >> 10) & 1; d.mapq = (r >> SR_MAPQ_SHIFT) & 255; d.set = (r >> SR_SET_SHIFT) & 31;
  return d;
}
__device__ __forceinline__ bool eligible(uint32_t r, uint32_t cutoff) {
  return (r & (SR_UNIQUE_BIT | SR_TRIM_BIT | SR_OK_BIT)) == (SR_UNIQUE_BIT | SR_OK_BIT) && ((r >> SR_QUAL_SHIFT) & 127) >= cutoff;
}

__device__ __forceinline__ void copy_to_smem(double* dst, const double* __restrict__ src, uint32_t n_doubles) {
  for (uint32_t i = threadIdx.x; i < n_doubles; i += blockDim.x) dst[i] = src[i];
}

}  // namespace

// ------------------------------------------------------------------------------------------ tally
__global__ void __launch_bounds__(TPB, 2) tally_kernel(const uint32_t* __restrict__ rec, const uint64_t* __restrict__ off,
                                                        const uint8_t* __restrict__ slot_ref, uint64_t n_slots,
                                                        const ClassTerms* __restrict__ lut, const HotTerms* __restrict__ hotL,
                                                        ScoreParams p, ColumnOut* __restrict__ out, uint32_t* __restrict__ worklist,
                                                        uint32_t* __restrict__ flagged, uint32_t* __restrict__ scalars,
                                                        uint32_t flagged_cap) {
  extern __shared__ __align__(16) double sm[];
  HotTerms* hot = reinterpret_cast<HotTerms*>(sm);
  copy_to_smem(sm, reinterpret_cast<const double*>(hotL), p.n_hot * 6);
  __shared__ uint8_t mapq_slot[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) mapq_slot[i] = p.mapq_slot[i];
  __syncthreads();

  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  const uint64_t stride = (uint64_t)gridDim.x * TPB;
  for (uint64_t slot = (uint64_t)blockIdx.x * TPB + threadIdx.x; slot < n_slots; slot += stride) {
    const uint64_t beg = off[slot], end = off[slot + 1];
    uint32_t u_top = 0, u_bot = 0, raw_top = 0, raw_bot = 0, n = 0, obs_mask = 0, err = 0;
    double red_top = 0.0, red_bot = 0.0, r2max = 0.0;
    double ll0 = 0.0, ll1 = 0.0, ll2 = 0.0, ll3 = 0.0, ll4 = 0.0;

    auto one = [&](uint32_t r) {
      if (r & SR_UNIQUE_BIT) {
        if (r & SR_TOP_BIT) ++u_top; else ++u_bot;
      } else {  // order-dependent double sum, arrival order (identify_mutations.cpp:1605)
        const double inv = 1.0 / (double)((r >> SR_RED_SHIFT) & 0xFFFFu);
        if (r & SR_TOP_BIT) { red_top += inv; ++raw_top; } else { red_bot += inv; ++raw_bot; }
        return;
      }
      if (!eligible(r, p.base_quality_cutoff)) return;
      const Decoded d = decode(r);
      if (d.qual >= p.max_qual || d.set >= p.max_set) { err |= BRQ_ERR_QUALITY_RANGE; return; }
      double r2;
      if (d.mapq == p.hot_mapq && p.n_hot) {
        const HotTerms& t = hot[((d.set * 2 + d.top) * p.max_qual + d.qual) * 5 + d.obs];
        ll0 += t.L[0]; ll1 += t.L[1]; ll2 += t.L[2]; ll3 += t.L[3]; ll4 += t.L[4];
        r2 = t.r2;
      } else {
        const ClassTerms& t = lut[((((d.set * 2 + d.top) * p.n_mapq_slots + mapq_slot[d.mapq]) * p.max_qual + d.qual) * 5 + d.obs)];
        ll0 += t.L[0]; ll1 += t.L[1]; ll2 += t.L[2]; ll3 += t.L[3]; ll4 += t.L[4];
        r2 = 0.0;
#pragma unroll
        for (int b = 0; b < 5; ++b) if ((uint32_t)b != d.obs) r2 = fmax(r2, t.r[b]);
        if (t.r[d.obs] != 1.0) r2 = inf;
      }
      r2max = fmax(r2max, r2);
      obs_mask |= 1u << d.obs;
      ++n;
    };

    // 128-bit loads from the 16-byte aligned vector that contains the slot's first record
    for (uint64_t v = beg & ~3ull; v < end; v += 4) {
      const uint4 q = ld_stream_v4(rec + v);
      if (v >= beg && v + 4 <= end) { one(q.x); one(q.y); one(q.z); one(q.w); }
      else {
        if (v >= beg && v < end) one(q.x);
        if (v + 1 >= beg && v + 1 < end) one(q.y);
        if (v + 2 >= beg && v + 2 < end) one(q.z);
        if (v + 3 >= beg && v + 3 < end) one(q.w);
      }
    }

    const uint32_t ref = slot_ref[slot];
    double ll[5] = {ll0, ll1, ll2, ll3, ll4};
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
        // below 2^-54 a term cannot change a sum that already holds the offset's own 1.0
        tot += (d == 0.0) ? 1.0 : (d < -17.0 ? 0.0 : pow(10.0, d));
      }
      consensus = (ll[best] - (log10(tot) + offv)) - p.log10_ref_length;
    }
    const double slack = 1e-6;
    const bool base_predicted = consensus >= p.mutation_cutoff;
    bool recheck = n > 0 && fabs(consensus - p.mutation_cutoff) < slack;

    // `pure`: every scoring record shows the reference base X, X is each record's most likely
    // true base, and every other hypothesis b has r_i(b) <= f0[X] = (n + 0.5) / (n + 2.5).
    // Then s_i >= f[X], so f[b] can only shrink from its start 0.5 / (n + 2.5) < 0.5 / n and f[X]
    // only grows: no allele but X ever reaches the half-read level, the fit reports major = X and
    // no minor / variant allele, and no presence score is computed (identify_mutations.cpp:1806-1821).
    const bool pure = n > 0 && ref < 5 && obs_mask == (1u << ref) && r2max <= ((double)n + 0.5) / ((double)n + 2.5);
    const bool needs_fit = n > 0 && !pure;
    uint32_t major = 5;
    if (pure) major = ref;
    uint32_t bits = best | (major << 3) | (5u << 6) | (5u << 9);
    if (base_predicted) bits |= CO_BASE_PREDICTED;
    if (raw_top + raw_bot == 0) bits |= CO_UNIQUE_ONLY;
    if (recheck) bits |= CO_RECHECK;

    ColumnOut o;
#pragma unroll
    for (int b = 0; b < 5; ++b) o.ll[b] = ll[b];
    o.consensus_score = consensus; o.variant_score = nan;
    o.redundant[0] = red_bot; o.redundant[1] = red_top;
    o.unique[0] = u_bot; o.unique[1] = u_top; o.raw_redundant[0] = raw_bot; o.raw_redundant[1] = raw_top;
    o.n = n; o.bits = bits;
    out[slot] = o;

    if (needs_fit) worklist[atomicAdd(&scalars[2], 1u)] = (uint32_t)slot;
    else if (recheck) { const uint32_t k = atomicAdd(&scalars[1], 1u); if (k < flagged_cap) flagged[k] = (uint32_t)slot; }
    if (err) atomicOr(&scalars[0], err);
  }
}

// ------------------------------------------------------------------------------------------ fit
namespace {

struct SlotRecords {
  const uint32_t* rec; uint64_t beg, end;
  const HotRatios* hot; const ClassTerms* lut; const uint8_t* mapq_slot; const ScoreParams* p;
};

// Visit the scoring records of a slot in arrival order: f(r[5], M, obs).
template <class F>
__device__ __forceinline__ void for_each_scoring(const SlotRecords& s, F&& f) {
  const ScoreParams& p = *s.p;
  for (uint64_t v = s.beg & ~3ull; v < s.end; v += 4) {
    const uint4 q = ld_cached_v4(s.rec + v);
    const uint32_t rr[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (v + j < s.beg || v + j >= s.end) continue;
      const uint32_t r = rr[j];
      if (!eligible(r, p.base_quality_cutoff)) continue;
      const Decoded d = decode(r);
      if (d.qual >= p.max_qual || d.set >= p.max_set) continue;
      if (d.mapq == p.hot_mapq && p.n_hot) {
        const HotRatios& t = s.hot[((d.set * 2 + d.top) * p.max_qual + d.qual) * 5 + d.obs];
        f(t.r, t.M, d.obs);
      } else {
        const ClassTerms& t = s.lut[((((d.set * 2 + d.top) * p.n_mapq_slots + s.mapq_slot[d.mapq]) * p.max_qual + d.qual) * 5 + d.obs)];
        double m = t.L[0];
#pragma unroll
        for (int b = 1; b < 5; ++b) m = fmax(m, t.L[b]);
        f(t.r, m, d.obs);
      }
    }
  }
}

struct Fit { double f[5]; double ll; uint32_t iterations; };

// identify_mutations.cpp:3240-3318, per-record, arrival order.
__device__ Fit em_fit(const SlotRecords& s, uint32_t n, const uint32_t obs_count[5], uint32_t allowed, double tol) {
  Fit m;
  double total = 0.0;
#pragma unroll
  for (int b = 0; b < 5; ++b) { m.f[b] = (allowed >> b & 1) ? 0.5 + (double)obs_count[b] : 0.0; total += m.f[b]; }
#pragma unroll
  for (int b = 0; b < 5; ++b) m.f[b] /= total;
  double f_prev[5];
  uint32_t it = 1;
  for (; it <= 50; ++it) {
    double w0 = 0, w1 = 0, w2 = 0, w3 = 0, w4 = 0;
    const double f0 = m.f[0], f1 = m.f[1], f2 = m.f[2], f3 = m.f[3], f4 = m.f[4];
    for_each_scoring(s, [&](const double* r, double, uint32_t) {
      const double a0 = f0 * r[0], a1 = f1 * r[1], a2 = f2 * r[2], a3 = f3 * r[3], a4 = f4 * r[4];
      const double sum = (((a0 + a1) + a2) + a3) + a4;
      if (sum > 0.0) {
        const double inv = 1.0 / sum;
        w0 += a0 * inv; w1 += a1 * inv; w2 += a2 * inv; w3 += a3 * inv; w4 += a4 * inv;
      } else { w0 += f0; w1 += f1; w2 += f2; w3 += f3; w4 += f4; }
    });
    const double w[5] = {w0, w1, w2, w3, w4};
    double max_delta = 0.0;
#pragma unroll
    for (int b = 0; b < 5; ++b) {
      f_prev[b] = m.f[b];
      if (allowed >> b & 1) {
        const double f_new = w[b] / (double)n;
        max_delta = fmax(max_delta, fabs(f_new - m.f[b]));
        m.f[b] = f_new;
      }
    }
    if (max_delta < tol) break;
  }
  m.iterations = it > 50 ? 50 : it;
  // the committed likelihood belongs to the frequencies BEFORE the last update
  double ll = 0.0;
  for_each_scoring(s, [&](const double* r, double M, uint32_t) {
    const double sum = (((f_prev[0] * r[0] + f_prev[1] * r[1]) + f_prev[2] * r[2]) + f_prev[3] * r[3]) + f_prev[4] * r[4];
    if (sum > 0.0) ll += log10(sum) + M;
  });
  m.ll = ll;
  return m;
}

}  // namespace

__global__ void __launch_bounds__(256, 2) fit_kernel(const uint32_t* __restrict__ rec, const uint64_t* __restrict__ off,
                                                      const uint8_t* __restrict__ slot_ref, const uint32_t* __restrict__ worklist,
                                                      const ClassTerms* __restrict__ lut, const HotRatios* __restrict__ hotR,
                                                      ScoreParams p, ColumnOut* __restrict__ out, uint32_t* __restrict__ flagged,
                                                      uint32_t* __restrict__ scalars, uint32_t flagged_cap) {
  extern __shared__ __align__(16) double sm[];
  copy_to_smem(sm, reinterpret_cast<const double*>(hotR), p.n_hot * 6);
  __shared__ uint8_t mapq_slot[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) mapq_slot[i] = p.mapq_slot[i];
  __syncthreads();
  const uint32_t n_work = scalars[2];
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n_work; w += gridDim.x * blockDim.x) {
    const uint32_t slot = worklist[w];
    SlotRecords s{rec, off[slot], off[slot + 1], reinterpret_cast<const HotRatios*>(sm), lut, mapq_slot, &p};
    uint32_t obs_count[5] = {0, 0, 0, 0, 0}, n = 0;
    for_each_scoring(s, [&](const double*, double, uint32_t obs) {
#pragma unroll
      for (int b = 0; b < 5; ++b) obs_count[b] += (obs == (uint32_t)b);
      ++n;
    });
    if (n == 0) continue;
    const uint32_t ref = slot_ref[slot];
    const double consensus = out[slot].consensus_score;
    uint32_t bits = out[slot].bits;
    const uint32_t best = bits & 7;
    bool recheck = (bits & CO_RECHECK) != 0;

    Fit full = em_fit(s, n, obs_count, 0x1F, p.precision_decimal);
    const double thr = 0.5 / (double)n;
    uint32_t major = 5, minor = 5, variant = 5, mj = 0;
#pragma unroll
    for (int b = 1; b < 5; ++b) if (full.f[b] > full.f[mj]) mj = b;
    major = full.f[mj] > 0.0 ? mj : 5;
#pragma unroll
    for (int b = 0; b < 5; ++b) {
      if (fabs(full.f[b] - thr) <= 1e-9 * thr) recheck = true;  // may land on the other side on the host
      if (full.f[b] < thr) continue;
      if ((uint32_t)b != major && (minor == 5 || full.f[b] > full.f[minor])) minor = b;
      if ((uint32_t)b != ref && (variant == 5 || full.f[b] > full.f[variant])) variant = b;
    }
    double variant_score = nan;
    if (variant != 5) {
      Fit null_fit = em_fit(s, n, obs_count, 0x1F & ~(1u << variant), p.precision_decimal);
      variant_score = (full.ll - null_fit.ll) - p.log10_ref_length;
    }
    const double slack = 1e-6;
    bool emit = false;
    if (best != ref && consensus > -slack) emit = true;
    if (variant != 5 && variant_score >= p.polymorphism_cutoff - slack) emit = true;
    bits = (bits & ~(0xFFFu | CO_EMIT | CO_RECHECK | (0xFFu << 16))) | best | (major << 3) | (minor << 6) | (variant << 9) | (full.iterations << 16);
    if (emit) bits |= CO_EMIT;
    if (recheck) bits |= CO_RECHECK;
    out[slot].variant_score = variant_score;
    out[slot].bits = bits;
    if (emit || recheck) { const uint32_t k = atomicAdd(&scalars[1], 1u); if (k < flagged_cap) flagged[k] = slot; }
  }
}

void launch_score_slots(const uint32_t* rec, const uint64_t* off, const uint8_t* slot_ref, uint64_t n_slots,
                        const ClassTerms* lut, const HotTerms* hotL, const HotRatios* hotR, const ScoreParams& p, ColumnOut* out,
                        uint32_t* worklist, uint32_t* flagged, uint32_t* scalars, uint32_t flagged_cap, cudaStream_t s,
                        cudaEvent_t between) {
  if (!n_slots) return;
  const int kSMs = 148;
  const size_t smem = (size_t)p.n_hot * 48;
  cudaFuncSetAttribute(tally_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int blocks = (int)std::min<uint64_t>((n_slots + TPB - 1) / TPB, (uint64_t)kSMs * 2);
  tally_kernel<<<blocks, TPB, smem, s>>>(rec, off, slot_ref, n_slots, lut, hotL, p, out, worklist, flagged, scalars, flagged_cap);
  if (between) cudaEventRecord(between, s);
  fit_kernel<<<kSMs * 2, 256, smem, s>>>(rec, off, slot_ref, worklist, lut, hotR, p, out, flagged, scalars, flagged_cap);
}

}  // namespace brq
