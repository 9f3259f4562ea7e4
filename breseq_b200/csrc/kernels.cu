// sm_100a kernels of the read-alignment evidence pileup.
//
//  hist_kernel           error_count covariate histogram  (error_count.cpp:125-199, 854-997)
//  coverage_hist_kernel  unique-only coverage histogram   (error_count.cpp:180-191)
//  derive_table_kernel   counts -> log10 probabilities    (error_count.cpp:1005-1026)
//  score_kernel          per-slot coverage tally, 5-way log-likelihood sums, pure-genotype call,
//                        5-allele EM fit and variant-presence score
//                        (identify_mutations.cpp:1591-1657, 3240-3344, 3398-3433)
//
// All of it is HBM-bound integer/byte work plus fp64 scalar math: no tensor cores.
#include "kernels.h"
#include "brq_types.h"

#include <atomic>
#include <cmath>

namespace brq {

static std::atomic<int> g_launches{0};
int launch_count() { return g_launches.load(); }
void note_launches(int n) { g_launches += n; }

__device__ __forceinline__ ulonglong2 ld_stream_u64x2(const ulonglong2* p) {
  ulonglong2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// ------------------------------------------------------------------------------------------
// error_count histogram
// ------------------------------------------------------------------------------------------
template <bool SMEM>
__device__ __forceinline__ void hist_add(uint32_t* sh, unsigned long long* counts, uint32_t idx) {
  if (SMEM) atomicAdd(&sh[idx], 1u);
  else atomicAdd(&counts[idx], 1ull);
}

template <bool SMEM>
__device__ __forceinline__ void hist_record(uint64_t r, const CovLayout& lay, uint32_t* sh, unsigned long long* counts,
                                            uint32_t& err) {
  const uint32_t lo = (uint32_t)r;
  const uint32_t obsA = lo & 7, refA = (lo >> HR_REFA) & 7, qa = (lo >> HR_QUALA) & 127, rev = (lo >> HR_REV) & 1;
  const uint32_t cls = (lo >> HR_CLASSB) & 3, obsB = (lo >> HR_OBSB) & 7, refB = (lo >> HR_REFB) & 7, qb = (lo >> HR_QUALB) & 127;
  const uint32_t set = (uint32_t)(r >> HR_SET) & 31, rpos = (uint32_t)(r >> HR_RPOS) & 0xFFFF;
  const uint32_t repA = (uint32_t)(r >> HR_REPA) & 255, repB = (uint32_t)(r >> HR_REPB) & 63;
  if (set >= lay.max_set) { err |= BRQ_ERR_READSET_RANGE; return; }
  if (rpos >= lay.max_rpos) { err |= BRQ_ERR_READPOS_RANGE; return; }
  const uint32_t base = set * lay.off_set + rpos * lay.off_rpos;
  // observation A: (ref base, observed base) on the read strand
  if (obsA < 4 && refA < 4) {
    if (qa >= lay.max_qual) err |= BRQ_ERR_QUALITY_RANGE;
    else {
      const uint32_t o = rev ? 3 - obsA : obsA, f = rev ? 3 - refA : refA;
      hist_add<SMEM>(sh, counts, base + f * lay.off_ref + o * lay.off_obs + qa * lay.off_qual + min(repA, lay.max_rep - 1) * lay.off_rep);
    }
  }
  // observation B: what follows this base in the read ('..', deletion, insertion)
  if (cls == 0 || obsB == kBaseN) return;
  uint32_t f = kBaseGap, o = kBaseGap;
  if (cls == 1) { if (refB == kBaseN) return; }
  else if (cls == 2) { if (refB >= 4) return; f = rev ? 3 - refB : refB; }
  else { o = rev ? 3 - obsB : obsB; }
  if (qb >= lay.max_qual) { err |= BRQ_ERR_QUALITY_RANGE; return; }
  hist_add<SMEM>(sh, counts, base + f * lay.off_ref + o * lay.off_obs + qb * lay.off_qual + min(repB, lay.max_rep - 1) * lay.off_rep);
}

template <bool SMEM>
__global__ void __launch_bounds__(256) hist_kernel(const uint64_t* __restrict__ rec, uint64_t n, CovLayout lay,
                                                    unsigned long long* __restrict__ counts, uint32_t* __restrict__ err_out) {
  extern __shared__ uint32_t sh[];
  if (SMEM) {
    for (uint32_t i = threadIdx.x; i < lay.n_bins; i += blockDim.x) sh[i] = 0;
    __syncthreads();
  }
  uint32_t err = 0;
  const ulonglong2* rec2 = reinterpret_cast<const ulonglong2*>(rec);
  const uint64_t n2 = n >> 1;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // four 128-bit loads in flight per thread
  for (; i + 3 * stride < n2; i += 4 * stride) {
    ulonglong2 a = ld_stream_u64x2(rec2 + i), b = ld_stream_u64x2(rec2 + i + stride);
    ulonglong2 c = ld_stream_u64x2(rec2 + i + 2 * stride), d = ld_stream_u64x2(rec2 + i + 3 * stride);
    hist_record<SMEM>(a.x, lay, sh, counts, err); hist_record<SMEM>(a.y, lay, sh, counts, err);
    hist_record<SMEM>(b.x, lay, sh, counts, err); hist_record<SMEM>(b.y, lay, sh, counts, err);
    hist_record<SMEM>(c.x, lay, sh, counts, err); hist_record<SMEM>(c.y, lay, sh, counts, err);
    hist_record<SMEM>(d.x, lay, sh, counts, err); hist_record<SMEM>(d.y, lay, sh, counts, err);
  }
  for (; i < n2; i += stride) {
    ulonglong2 a = ld_stream_u64x2(rec2 + i);
    hist_record<SMEM>(a.x, lay, sh, counts, err); hist_record<SMEM>(a.y, lay, sh, counts, err);
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) hist_record<SMEM>(rec[n - 1], lay, sh, counts, err);
  if (err) atomicOr(err_out, err);
  if (SMEM) {
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < lay.n_bins; b += blockDim.x) {
      uint32_t v = sh[b];
      if (v) atomicAdd(&counts[b], (unsigned long long)v);
    }
  }
}

void launch_hist(const uint64_t* rec, uint64_t n_rec, const CovLayout& lay, unsigned long long* counts, uint32_t* err,
                 cudaStream_t s) {
  const int kSMs = 148;
  const size_t smem = (size_t)lay.n_bins * 4;
  if (smem <= 200 * 1024) {
    cudaFuncSetAttribute(hist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = smem <= 24 * 1024 ? 8 : (smem <= 48 * 1024 ? 4 : (smem <= 100 * 1024 ? 2 : 1));
    hist_kernel<true><<<kSMs * per_sm, 256, smem, s>>>(rec, n_rec, lay, counts, err);
  } else {
    hist_kernel<false><<<kSMs * 8, 256, 0, s>>>(rec, n_rec, lay, counts, err);
  }
  ++g_launches;
}

// ------------------------------------------------------------------------------------------
// unique-only coverage histogram: one increment per column without redundant reads
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) coverage_hist_kernel(const uint64_t* __restrict__ hist_off, const uint8_t* __restrict__ group,
                                                             uint64_t n_cols, uint32_t stride, unsigned long long* __restrict__ cov,
                                                             uint32_t* __restrict__ err_out) {
  const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_cols; c += step) {
    const uint64_t a = hist_off[c], b = hist_off[c + 1];
    if (a & HIST_OFF_REDUNDANT_BIT) continue;
    const uint64_t depth = (b & ~HIST_OFF_REDUNDANT_BIT) - a;
    if (depth >= stride) { atomicOr(err_out, BRQ_ERR_DEPTH_RANGE); continue; }
    atomicAdd(&cov[(uint64_t)group[c] * stride + depth], 1ull);
  }
}

void launch_coverage_hist(const uint64_t* hist_off, const uint8_t* group, uint64_t n_cols, uint32_t stride,
                          unsigned long long* cov_hist, uint32_t* err, cudaStream_t s) {
  if (!n_cols) return;
  int blocks = (int)std::min<uint64_t>((n_cols + 255) / 256, 148 * 8);
  coverage_hist_kernel<<<blocks, 256, 0, s>>>(hist_off, group, n_cols, stride, cov_hist, err);
  ++g_launches;
}

// ------------------------------------------------------------------------------------------
// counts -> log10 probability
// ------------------------------------------------------------------------------------------
__global__ void derive_table_kernel(const unsigned long long* __restrict__ counts, CovLayout lay, double* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= lay.n_bins) return;
  const uint32_t j = (i / lay.off_obs) % 5;
  const uint32_t base = i - j * lay.off_obs;
  unsigned long long sum = 0;
#pragma unroll
  for (uint32_t k = 0; k < 5; ++k) sum += counts[base + k * lay.off_obs];
  // two log10 calls and a subtraction, the shape the reference uses
  out[i] = log10((double)counts[i] + 1.0) - log10((double)sum + 5.0);
}

void launch_derive_table(const unsigned long long* counts, const CovLayout& lay, double* log10_prob, cudaStream_t s) {
  derive_table_kernel<<<(lay.n_bins + 255) / 256, 256, 0, s>>>(counts, lay, log10_prob);
  ++g_launches;
}

// ------------------------------------------------------------------------------------------
// per-slot scoring: one warp per slot
// ------------------------------------------------------------------------------------------
constexpr int SC_WARPS = 8;          // warps per CTA
constexpr uint32_t SC_TABLE = 1024;  // class hash slots per warp
constexpr uint32_t SC_CAP = 960;     // distinct classes accepted per slot
constexpr uint32_t SC_EMPTY = 0xFFFFFFFFu;

struct WarpScratch {
  uint32_t keys[SC_TABLE];
  uint32_t counts[SC_TABLE];
  uint16_t list[SC_TABLE];
  uint32_t n_list;
  uint32_t pad[3];
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint32_t warp_sum_u32(uint32_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// key: the record's low 24 bits: [2:0] obs [9:3] qual [10] top [15:11] read_set [23:16] mapq
__device__ __forceinline__ uint32_t class_key(uint32_t rec) { return rec & 0xFFFFFFu; }
__device__ __forceinline__ uint32_t lut_index(uint32_t key, const ScoreParams& p, const uint8_t* mapq_slot) {
  const uint32_t obs = key & 7, qual = (key >> 3) & 127, top = (key >> 10) & 1, set = (key >> 11) & 31, mapq = (key >> 16) & 255;
  return ((((set * 2 + top) * p.n_mapq_slots + mapq_slot[mapq]) * p.max_qual + qual) * 5 + obs);
}

struct EmResult { double f[5]; double log10_likelihood; uint32_t iterations; };

// EM over record classes (identify_mutations.cpp:3240-3318).  Mathematically the per-read EM with
// equal reads merged: every read of a class contributes the same w_i(b), so the class adds
// count * w(b).  All lanes hold the same f[]; classes are strided over lanes.
__device__ EmResult em_fit(const WarpScratch& ws, const ClassTerms* __restrict__ lut, uint32_t n_cls,
                           uint32_t n, const uint32_t obs_count[5], uint32_t allowed_mask, double tolerance, int lane) {
  EmResult m;
  double init_total = 0.0;
#pragma unroll
  for (int b = 0; b < 5; ++b) {
    m.f[b] = (allowed_mask >> b & 1) ? 0.5 + (double)obs_count[b] : 0.0;
    init_total += m.f[b];
  }
#pragma unroll
  for (int b = 0; b < 5; ++b) m.f[b] /= init_total;
  double f_prev[5];
  uint32_t it = 1;
  for (; it <= 50; ++it) {
    double w[5] = {0, 0, 0, 0, 0};
    for (uint32_t c = lane; c < n_cls; c += 32) {
      const uint32_t h = ws.list[c];
      const ClassTerms& t = lut[ws.keys[h]];
      const double cnt = (double)ws.counts[h];
      double fr[5], s = 0.0;
#pragma unroll
      for (int b = 0; b < 5; ++b) { fr[b] = m.f[b] * t.r[b]; s += fr[b]; }
      if (s > 0.0) {
#pragma unroll
        for (int b = 0; b < 5; ++b) w[b] += cnt * (fr[b] / s);
      } else {
#pragma unroll
        for (int b = 0; b < 5; ++b) w[b] += cnt * m.f[b];
      }
    }
    double max_delta = 0.0;
#pragma unroll
    for (int b = 0; b < 5; ++b) {
      f_prev[b] = m.f[b];
      if (allowed_mask >> b & 1) {
        const double f_new = warp_sum(w[b]) / (double)n;
        max_delta = fmax(max_delta, fabs(f_new - m.f[b]));
        m.f[b] = f_new;
      }
    }
    if (max_delta < tolerance) break;
  }
  m.iterations = it > 50 ? 50 : it;
  // the committed likelihood is the one evaluated with the frequencies BEFORE the last update
  double ll = 0.0;
  for (uint32_t c = lane; c < n_cls; c += 32) {
    const uint32_t h = ws.list[c];
    const ClassTerms& t = lut[ws.keys[h]];
    const double cnt = (double)ws.counts[h];
    double s = 0.0, mx = t.L[0];
#pragma unroll
    for (int b = 0; b < 5; ++b) { s += f_prev[b] * t.r[b]; mx = fmax(mx, t.L[b]); }
    if (s > 0.0) ll += cnt * (log10(s) + mx);
  }
  m.log10_likelihood = warp_sum(ll);
  return m;
}

__global__ void __launch_bounds__(SC_WARPS * 32) score_kernel(const uint32_t* __restrict__ rec, const uint64_t* __restrict__ off,
                                                               const uint8_t* __restrict__ slot_ref, uint64_t n_slots,
                                                               const ClassTerms* __restrict__ lut, ScoreParams p,
                                                               ColumnOut* __restrict__ out, uint32_t* __restrict__ flagged,
                                                               uint32_t* __restrict__ n_flagged, uint32_t flagged_cap,
                                                               uint32_t* __restrict__ err_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WarpScratch* scratch = reinterpret_cast<WarpScratch*>(smem_raw);
  uint8_t* mapq_slot = smem_raw + sizeof(WarpScratch) * SC_WARPS;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) mapq_slot[i] = p.mapq_slot[i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpScratch& ws = scratch[warp];
  for (uint32_t i = lane; i < SC_TABLE; i += 32) { ws.keys[i] = SC_EMPTY; ws.counts[i] = 0; }
  if (lane == 0) ws.n_list = 0;
  __syncthreads();

  const uint64_t n_warps = (uint64_t)gridDim.x * SC_WARPS;
  for (uint64_t slot = (uint64_t)blockIdx.x * SC_WARPS + warp; slot < n_slots; slot += n_warps) {
    const uint64_t beg = off[slot], end = off[slot + 1];
    uint32_t uniq_top = 0, uniq_bot = 0, raw_top = 0, raw_bot = 0;
    double red_top = 0.0, red_bot = 0.0;
    uint32_t err = 0;

    for (uint64_t base = beg; base < end; base += 32) {
      const uint64_t i = base + lane;
      const bool have = i < end;
      const uint32_t r = have ? ld_stream_u32(rec + i) : 0u;
      const bool unique = have && (r & SR_UNIQUE_BIT), top = r & SR_TOP_BIT;
      uniq_top += __popc(__ballot_sync(0xffffffffu, unique && top));
      uniq_bot += __popc(__ballot_sync(0xffffffffu, unique && !top));
      uint32_t redm = __ballot_sync(0xffffffffu, have && !unique);
      while (redm) {  // order-dependent double sum: strictly in arrival order
        const int l = __ffs(redm) - 1;
        redm &= redm - 1;
        const uint32_t rr = __shfl_sync(0xffffffffu, r, l);
        const double inv = 1.0 / (double)((rr >> SR_RED_SHIFT) & SR_RED_MASK);
        if (rr & SR_TOP_BIT) { red_top += inv; ++raw_top; } else { red_bot += inv; ++raw_bot; }
      }
      const uint32_t qual = (r >> SR_QUAL_SHIFT) & 127;
      if (unique && !(r & SR_TRIM_BIT) && (r & SR_OK_BIT) && qual >= p.base_quality_cutoff) {
        if (qual >= p.max_qual || ((r >> SR_SET_SHIFT) & 31) >= p.max_set) err |= BRQ_ERR_QUALITY_RANGE;
        else {
          const uint32_t key = class_key(r);
          uint32_t h = (key * 2654435761u) >> 22;
          for (;;) {
            uint32_t k = *(volatile uint32_t*)&ws.keys[h];
            if (k == SC_EMPTY) {
              if (*(volatile uint32_t*)&ws.n_list >= SC_CAP) { err |= BRQ_ERR_CLASS_OVERFLOW; break; }
              k = atomicCAS(&ws.keys[h], SC_EMPTY, key);
              if (k == SC_EMPTY) { ws.list[atomicAdd(&ws.n_list, 1u)] = (uint16_t)h; k = key; }
            }
            if (k == key) { atomicAdd(&ws.counts[h], 1u); break; }
            h = (h + 1) & (SC_TABLE - 1);
          }
        }
      }
    }
    __syncwarp();
    const uint32_t n_cls = min(ws.n_list, SC_CAP);

    // per-class lookups: 5-way log-likelihood sums, per-base counts
    double ll[5] = {0, 0, 0, 0, 0};
    uint32_t obs_cnt[5] = {0, 0, 0, 0, 0};
    uint32_t n = 0;
    for (uint32_t c = lane; c < n_cls; c += 32) {
      const uint32_t h = ws.list[c];
      const uint32_t key = ws.keys[h], cnt = ws.counts[h];
      const uint32_t li = lut_index(key, p, mapq_slot);
      ws.keys[h] = li;  // from here on the slot holds the class's LUT index (obs == li % 5)
      const ClassTerms& t = lut[li];
#pragma unroll
      for (int b = 0; b < 5; ++b) { ll[b] += (double)cnt * t.L[b]; if ((key & 7) == (uint32_t)b) obs_cnt[b] += cnt; }
      n += cnt;
    }
#pragma unroll
    for (int b = 0; b < 5; ++b) { ll[b] = warp_sum(ll[b]); obs_cnt[b] = warp_sum_u32(obs_cnt[b]); }
    n = warp_sum_u32(n);
    __syncwarp();

    const uint32_t ref = slot_ref[slot];
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    double consensus = nan, variant_score = nan;
    uint32_t best = 5, major = 5, minor = 5, variant = 5, iters = 0;
    bool recheck = false;
    if (n > 0) {
      // pure_genotype_call
      best = 0;
#pragma unroll
      for (int b = 1; b < 5; ++b) if (ll[b] > ll[best]) best = b;
      double offv = -1.7976931348623157e308;
#pragma unroll
      for (int b = 0; b < 5; ++b) if ((uint32_t)b != best) offv = fmax(offv, ll[b]);
      double tot = 0.0;
#pragma unroll
      for (int b = 0; b < 5; ++b) if ((uint32_t)b != best) tot += pow(10.0, ll[b] - offv);
      consensus = (ll[best] - (log10(tot) + offv)) - p.log10_ref_length;

      // 5-allele fit, then the presence score of the top non-reference allele
      EmResult full = em_fit(ws, lut, n_cls, n, obs_cnt, 0x1F, p.precision_decimal, lane);
      iters = full.iterations;
      const double thr = 0.5 / (double)n;
      uint32_t mj = 0;
#pragma unroll
      for (int b = 1; b < 5; ++b) if (full.f[b] > full.f[mj]) mj = b;
      major = full.f[mj] > 0.0 ? mj : 5;
#pragma unroll
      for (int b = 0; b < 5; ++b) {
        if (full.f[b] < thr) continue;
        if ((uint32_t)b != major && (minor == 5 || full.f[b] > full.f[minor])) minor = b;
        if ((uint32_t)b != ref && (variant == 5 || full.f[b] > full.f[variant])) variant = b;
        // a frequency within rounding distance of the half-read threshold may flip on the host
        if (fabs(full.f[b] - thr) <= 1e-9 * thr) recheck = true;
      }
      if (variant != 5) {
        EmResult null_fit = em_fit(ws, lut, n_cls, n, obs_cnt, 0x1F & ~(1u << variant), p.precision_decimal, lane);
        variant_score = (full.log10_likelihood - null_fit.log10_likelihood) - p.log10_ref_length;
      }
    }

    // reset the class table for the next slot
    for (uint32_t c = lane; c < n_cls; c += 32) { const uint32_t h = ws.list[c]; ws.keys[h] = SC_EMPTY; ws.counts[h] = 0; }
    if (lane == 0) ws.n_list = 0;
    __syncwarp();

    // decisions.  Borderline values are flagged so the host can repeat them in reference
    // (arrival) order; clear-cut ones are final.
    const double slack = 1e-6;
    const bool base_predicted = consensus >= p.mutation_cutoff;
    if (n > 0 && fabs(consensus - p.mutation_cutoff) < slack) recheck = true;
    bool emit = false;
    if (n > 0) {
      if (best != ref && consensus > -slack) emit = true;
      if (variant != 5 && variant_score >= p.polymorphism_cutoff - slack) emit = true;
    }
    uint32_t bits = best | (major << 3) | (minor << 6) | (variant << 9) | (iters << 16);
    if (base_predicted) bits |= CO_BASE_PREDICTED;
    if (raw_top + raw_bot == 0) bits |= CO_UNIQUE_ONLY;
    if (emit) bits |= CO_EMIT;
    if (recheck) bits |= CO_RECHECK;

    // 96-byte result, written as 12 consecutive 8-byte words by lanes 0..11
    unsigned long long word = 0;
    switch (lane) {
      case 0: case 1: case 2: case 3: case 4: word = __double_as_longlong(ll[lane]); break;
      case 5: word = __double_as_longlong(consensus); break;
      case 6: word = __double_as_longlong(variant_score); break;
      case 7: word = __double_as_longlong(red_bot); break;
      case 8: word = __double_as_longlong(red_top); break;
      case 9: word = (unsigned long long)uniq_bot | ((unsigned long long)uniq_top << 32); break;
      case 10: word = (unsigned long long)raw_bot | ((unsigned long long)raw_top << 32); break;
      case 11: word = (unsigned long long)n | ((unsigned long long)bits << 32); break;
      default: break;
    }
    if (lane < 12) reinterpret_cast<unsigned long long*>(out + slot)[lane] = word;
    if (lane == 0 && (emit || recheck)) {
      const uint32_t k = atomicAdd(n_flagged, 1u);
      if (k < flagged_cap) flagged[k] = (uint32_t)slot;
    }
    err = __reduce_or_sync(0xffffffffu, err);
    if (err && lane == 0) atomicOr(err_out, err);
  }
}

void launch_score(const uint32_t* rec, const uint64_t* off, const uint8_t* slot_ref, uint64_t n_slots, const ClassTerms* lut,
                  const ScoreParams& p, ColumnOut* out, uint32_t* flagged, uint32_t* n_flagged, uint32_t flagged_cap,
                  uint32_t* err, cudaStream_t s) {
  if (!n_slots) return;
  const int kSMs = 148;
  int blocks = (int)std::min<uint64_t>((n_slots + SC_WARPS - 1) / SC_WARPS, (uint64_t)kSMs * 2);
  const size_t smem = sizeof(WarpScratch) * SC_WARPS + 256;
  cudaFuncSetAttribute(score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  score_kernel<<<blocks, SC_WARPS * 32, smem, s>>>(rec, off, slot_ref, n_slots, lut, p, out, flagged, n_flagged, flagged_cap, err);
  ++g_launches;
}

}  // namespace brq
