// sm_100a kernels of the read-alignment evidence pileup.
//
//  hist16_kernel, hist_kernel   error_count covariate histogram  (error_count.cpp:125-199, 854-997): 16-bit fast records +
//                               exceptions (the form the device reads at the default covariates), or 4- / 8-byte records
//  coverage_hist_kernel         unique-only coverage histogram   (error_count.cpp:180-191)
//  derive_table_kernel          counts -> log10 probabilities    (error_count.cpp:1005-1026)
//  canonical_table_kernel       log10 table -> probabilities through the six-digit text round trip (error_count.cpp:629-690)
//  expand_score_kernel          PCIe transfer form of the scoring stream -> score_rec (brq_types.h)
//  (per-slot scoring -- coverage tally, 5-way log-likelihood sums, consensus call, EM fit -- is in score_slots.cu)
//
// All of it is HBM-bound integer/byte work plus fp64 scalar math: no tensor cores.
#include "kernels.h"
#include "brq_types.h"
#include "canonical.h"

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
// Records arrive with both observations already resolved to table coordinates (brq_types.h), four bytes
// each unless the run needs read_pos / base_repeat, and the
// host has checked the stream's largest quality / read_set / read_pos against the table before the
// launch (the reference's fatal ASSERT, error_count.cpp:485-488), so a record costs two multiply-add
// chains and two shared-memory atomics.  FULL adds the read_pos / base_repeat covariates.
template <bool SMEM>
__device__ __forceinline__ void hist_add(uint32_t* sh, unsigned long long* counts, uint32_t idx) {
  if (SMEM) atomicAdd(&sh[idx], 1u);
  else atomicAdd(&counts[idx], 1ull);
}

// lo = the 4-byte record; hi = the high word of an 8-byte record (WIDE: read_pos / base_repeat covariates, > 16 read files)
template <bool SMEM, bool WIDE>
__device__ __forceinline__ void hist_record(uint32_t lo, uint32_t hi, const CovLayout& lay, uint32_t* sh, unsigned long long* counts) {
  uint32_t base = ((lo >> HR_SET) & 7u) * lay.off_set;
  if (WIDE) base += ((hi >> (HR_SET_HI - 32)) << 3) * lay.off_set + (hi & 0xFFFFu) * lay.off_rpos;
  if (lo & (1u << HR_VALIDA)) {
    uint32_t idx = base + (lo & 7u) * lay.off_ref + ((lo >> HR_OBSA) & 7u) * lay.off_obs + ((lo >> HR_QUALA) & 127u) * lay.off_qual;
    if (WIDE) idx += min((hi >> (HR_REPA - 32)) & 255u, lay.max_rep - 1) * lay.off_rep;
    hist_add<SMEM>(sh, counts, idx);
  }
  if (lo & (1u << HR_VALIDB)) {
    uint32_t idx = base + ((lo >> HR_REFB) & 7u) * lay.off_ref + ((lo >> HR_OBSB) & 7u) * lay.off_obs + ((lo >> HR_QUALB) & 127u) * lay.off_qual;
    if (WIDE) idx += min((hi >> (HR_REPB - 32)) & 31u, lay.max_rep - 1) * lay.off_rep;
    hist_add<SMEM>(sh, counts, idx);
  }
}

__device__ __forceinline__ uint4 ld_stream_u32x4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// JOINT: the shared-memory atomic unit, not HBM, bounds a kernel that issues two atomics per record (5 cycles per warp
// instruction and SM, scripts/micro/atoms_rate.cu).  All but ~0.2 % of the records are the same pair of observations: the
// aligned base matches the reference (ref == obs, quality qa) and the next base of the read is aligned too (('.', '.'),
// quality qb).  Those records take ONE atomic on a CTA-private joint histogram over (read set, base, qa, qb), whose two
// marginals are added to the covariate histogram when the CTA is done; every other record takes the two-atomic path.
// index of a fast record (brq_types.h: HR_FAST) in the joint histogram [sets][4 bases][Q][Q + 1]; an absent B reads quality
// 127 and lands in column Q
__device__ __forceinline__ uint32_t joint_index(uint32_t lo, uint32_t n_set, uint32_t Q) {
  const uint32_t set = n_set > 1 ? ((lo >> HR_SET) & 7u) : 0u, qb = min((lo >> HR_QUALB) & 127u, Q);
  return ((set * 4u + (lo & 3u)) * Q + ((lo >> HR_QUALA) & 127u)) * (Q + 1u) + qb;
}

// `rec` holds n records of 4 bytes (WIDE = false: four per 128-bit load) or 8 bytes (WIDE: two per load)
template <bool SMEM, bool WIDE, bool JOINT = false>
__global__ void __launch_bounds__(JOINT ? 1024 : 256) hist_kernel(const void* __restrict__ rec, uint64_t n, CovLayout lay,
                                                    unsigned long long* __restrict__ counts, uint32_t joint_sets = 0) {
  extern __shared__ uint32_t sh[];
  uint32_t* joint = sh + ((lay.n_bins + 3u) & ~3u);   // JOINT: [joint_sets][4][Q][Q + 1] after the covariate histogram
  const uint32_t n_joint = JOINT ? joint_sets * 4u * lay.max_qual * (lay.max_qual + 1u) : 0u;
  if (SMEM) {
    for (uint32_t i = threadIdx.x; i < lay.n_bins; i += blockDim.x) sh[i] = 0;
    for (uint32_t i = threadIdx.x; i < n_joint; i += blockDim.x) joint[i] = 0;
    __syncthreads();
  }
  const uint4* vec = reinterpret_cast<const uint4*>(rec);
  constexpr uint32_t PER = WIDE ? 2 : 4;  // records per vector
  const uint64_t n_vec = n / PER;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  auto one = [&](const uint4& v) {
    if (WIDE) { hist_record<SMEM, true>(v.x, v.y, lay, sh, counts); hist_record<SMEM, true>(v.z, v.w, lay, sh, counts); }
    else if (JOINT) {
      const uint32_t r[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if ((int32_t)r[j] < 0) atomicAdd(&joint[joint_index(r[j], joint_sets, lay.max_qual)], 1u);
        else hist_record<SMEM, false>(r[j], 0u, lay, sh, counts);
      }
    } else {
      hist_record<SMEM, false>(v.x, 0u, lay, sh, counts); hist_record<SMEM, false>(v.y, 0u, lay, sh, counts);
      hist_record<SMEM, false>(v.z, 0u, lay, sh, counts); hist_record<SMEM, false>(v.w, 0u, lay, sh, counts);
    }
  };
  // four 128-bit loads in flight per thread
  for (; i + 3 * stride < n_vec; i += 4 * stride) {
    const uint4 a = ld_stream_u32x4(vec + i), b = ld_stream_u32x4(vec + i + stride);
    const uint4 c = ld_stream_u32x4(vec + i + 2 * stride), d = ld_stream_u32x4(vec + i + 3 * stride);
    one(a); one(b); one(c); one(d);
  }
  for (; i < n_vec; i += stride) one(ld_stream_u32x4(vec + i));
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // the last partial vector
    const uint32_t* w = reinterpret_cast<const uint32_t*>(rec);
    for (uint64_t k = n_vec * PER; k < n; ++k) {
      if (WIDE) hist_record<SMEM, true>(w[2 * k], w[2 * k + 1], lay, sh, counts);
      else hist_record<SMEM, false>(w[k], 0u, lay, sh, counts);
    }
  }
  if (SMEM) {
    __syncthreads();
    if (JOINT) {  // the two marginals of the joint histogram
      const uint32_t Q = lay.max_qual;
      for (uint32_t i = threadIdx.x; i < n_joint; i += blockDim.x) {
        const uint32_t c = joint[i];
        if (!c) continue;
        const uint32_t qb = i % (Q + 1u), qa = (i / (Q + 1u)) % Q, base = (i / ((Q + 1u) * Q)) & 3u, set = i / (4u * Q * (Q + 1u));
        const uint32_t off = set * lay.off_set;
        atomicAdd(&sh[off + base * lay.off_ref + base * lay.off_obs + qa * lay.off_qual], c);
        if (qb < Q) atomicAdd(&sh[off + 4u * lay.off_ref + 4u * lay.off_obs + qb * lay.off_qual], c);
      }
      __syncthreads();
    }
    for (uint32_t b = threadIdx.x; b < lay.n_bins; b += blockDim.x) {
      uint32_t v = sh[b];
      if (v) atomicAdd(&counts[b], (unsigned long long)v);
    }
  }
}

// The compact form (brq_types.h): `rec16` holds n16 fast records of 16 bits (eight per 128-bit load), `exc` the n_exc
// records without a 16-bit form (4 bytes, generic path).  JOINT as above: one atomic per fast record on the CTA's joint
// histogram; otherwise a fast record is expanded to its 4-byte form and takes the generic two-atomic path.
__device__ __forceinline__ uint32_t joint_index16(uint32_t r, uint32_t n_set, uint32_t Q) {  // [sets][4 bases][Q][Q + 1]; r: 16 bits
  const uint32_t set = n_set > 1 ? r >> 14 : 0u;
  return ((set * 4u + (r & 3u)) * Q + ((r >> 2) & 63u)) * (Q + 1u) + min((r >> 8) & 63u, Q);
}
template <bool SMEM, bool JOINT>
__global__ void __launch_bounds__(JOINT ? 1024 : 256) hist16_kernel(const uint4* __restrict__ rec16, uint64_t n16, const uint32_t* __restrict__ exc, uint64_t n_exc,
                                                                      CovLayout lay, unsigned long long* __restrict__ counts, uint32_t joint_sets) {
  extern __shared__ uint32_t sh[];
  uint32_t* joint = sh + ((lay.n_bins + 3u) & ~3u);
  const uint32_t n_joint = JOINT ? joint_sets * 4u * lay.max_qual * (lay.max_qual + 1u) : 0u;
  if (SMEM) {
    for (uint32_t i = threadIdx.x; i < lay.n_bins; i += blockDim.x) sh[i] = 0;
    for (uint32_t i = threadIdx.x; i < n_joint; i += blockDim.x) joint[i] = 0;
    __syncthreads();
  }
  const uint32_t Q = lay.max_qual;
  auto fast = [&](uint32_t r) {  // r: one 16-bit record
    if (JOINT) atomicAdd(&joint[joint_index16(r, joint_sets, Q)], 1u);
    else hist_record<SMEM, false>(hist16_expand(r), 0u, lay, sh, counts);
  };
  auto one = [&](const uint4& v) {
    fast(v.x & 0xFFFFu); fast(v.x >> 16); fast(v.y & 0xFFFFu); fast(v.y >> 16);
    fast(v.z & 0xFFFFu); fast(v.z >> 16); fast(v.w & 0xFFFFu); fast(v.w >> 16);
  };
  const uint64_t n_vec = n16 / 8, stride = (uint64_t)gridDim.x * blockDim.x, tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t i = tid;
  for (; i + 3 * stride < n_vec; i += 4 * stride) {  // four 128-bit loads in flight per thread
    const uint4 a = ld_stream_u32x4(rec16 + i), b = ld_stream_u32x4(rec16 + i + stride);
    const uint4 c = ld_stream_u32x4(rec16 + i + 2 * stride), d = ld_stream_u32x4(rec16 + i + 3 * stride);
    one(a); one(b); one(c); one(d);
  }
  for (; i < n_vec; i += stride) one(ld_stream_u32x4(rec16 + i));
  if (tid == 0) {  // the last partial vector
    const uint16_t* h = reinterpret_cast<const uint16_t*>(rec16);
    for (uint64_t k = n_vec * 8; k < n16; ++k) fast(h[k]);
  }
  for (uint64_t k = tid; k < n_exc; k += stride) hist_record<SMEM, false>(ld_stream_u32(exc + k), 0u, lay, sh, counts);
  if (SMEM) {
    __syncthreads();
    if (JOINT) {  // the two marginals of the joint histogram
      for (uint32_t j = threadIdx.x; j < n_joint; j += blockDim.x) {
        const uint32_t c = joint[j];
        if (!c) continue;
        const uint32_t qb = j % (Q + 1u), qa = (j / (Q + 1u)) % Q, base = (j / ((Q + 1u) * Q)) & 3u, set = j / (4u * Q * (Q + 1u));
        const uint32_t off = set * lay.off_set;
        atomicAdd(&sh[off + base * lay.off_ref + base * lay.off_obs + qa * lay.off_qual], c);
        if (qb < Q) atomicAdd(&sh[off + 4u * lay.off_ref + 4u * lay.off_obs + qb * lay.off_qual], c);
      }
      __syncthreads();
    }
    for (uint32_t b = threadIdx.x; b < lay.n_bins; b += blockDim.x) {
      const uint32_t v = sh[b];
      if (v) atomicAdd(&counts[b], (unsigned long long)v);
    }
  }
}

void launch_hist16(const void* rec16, uint64_t n16, const uint32_t* exc, uint64_t n_exc, const CovLayout& lay, unsigned long long* counts, cudaStream_t s) {
  const int kSMs = 148;
  const size_t smem = (size_t)lay.n_bins * 4;
  const uint4* r = static_cast<const uint4*>(rec16);
  // the joint histogram needs exactly the four default covariates, every quality a fast record can hold inside the
  // table (<= 62 < max_qual is checked by the caller against the stream's maxima), and room beside the table
  const uint32_t joint_sets = lay.off_set ? lay.max_set : 1u;
  const size_t smem_joint = (((size_t)lay.n_bins + 3) & ~(size_t)3) * 4 + (size_t)joint_sets * 4 * lay.max_qual * (lay.max_qual + 1) * 4;
  if (lay.off_qual && lay.off_ref && lay.off_obs && !lay.off_rpos && !lay.off_rep && lay.max_qual <= 64 &&
      joint_sets <= 8 && smem_joint <= 110 * 1024) {
    const int per_sm = (int)std::min<size_t>(2, std::max<size_t>(1, (220 * 1024) / smem_joint));
    cudaFuncSetAttribute(hist16_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_joint);
    hist16_kernel<true, true><<<kSMs * per_sm, 1024, smem_joint, s>>>(r, n16, exc, n_exc, lay, counts, joint_sets);
  } else if (smem <= 200 * 1024) {
    const int per_sm = smem <= 24 * 1024 ? 8 : (smem <= 48 * 1024 ? 4 : (smem <= 100 * 1024 ? 2 : 1));
    cudaFuncSetAttribute(hist16_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    hist16_kernel<true, false><<<kSMs * per_sm, 256, smem, s>>>(r, n16, exc, n_exc, lay, counts, 0u);
  } else {
    hist16_kernel<false, false><<<kSMs * 8, 256, 0, s>>>(r, n16, exc, n_exc, lay, counts, 0u);
  }
  ++g_launches;
}

void launch_hist(const void* rec, uint64_t n_rec, bool wide, const CovLayout& lay, unsigned long long* counts, cudaStream_t s) {
  const int kSMs = 148;
  const size_t smem = (size_t)lay.n_bins * 4;
  // joint histogram of the dominant record kind: needs exactly the four default covariates and has to fit beside the table
  const uint32_t joint_sets = lay.off_set ? lay.max_set : 1u;
  const size_t smem_joint = (((size_t)lay.n_bins + 3) & ~(size_t)3) * 4 + (size_t)joint_sets * 4 * lay.max_qual * (lay.max_qual + 1) * 4;
  if (!wide && lay.off_qual && lay.off_ref && lay.off_obs && !lay.off_rpos && !lay.off_rep && lay.max_qual <= 64 &&
      joint_sets <= 8 && smem_joint <= 110 * 1024) {
    // CTAs of 1024 threads share one joint histogram: two of them keep the SM's 64 warps busy
    const int per_sm = (int)std::min<size_t>(2, std::max<size_t>(1, (220 * 1024) / smem_joint));
    cudaFuncSetAttribute(hist_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_joint);
    hist_kernel<true, false, true><<<kSMs * per_sm, 1024, smem_joint, s>>>(rec, n_rec, lay, counts, joint_sets);
    ++g_launches;
    return;
  }
  if (smem <= 200 * 1024) {
    int per_sm = smem <= 24 * 1024 ? 8 : (smem <= 48 * 1024 ? 4 : (smem <= 100 * 1024 ? 2 : 1));
    if (wide) {
      cudaFuncSetAttribute(hist_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      hist_kernel<true, true><<<kSMs * per_sm, 256, smem, s>>>(rec, n_rec, lay, counts);
    } else {
      cudaFuncSetAttribute(hist_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      hist_kernel<true, false><<<kSMs * per_sm, 256, smem, s>>>(rec, n_rec, lay, counts);
    }
  } else {
    if (wide) hist_kernel<false, true><<<kSMs * 8, 256, 0, s>>>(rec, n_rec, lay, counts);
    else hist_kernel<false, false><<<kSMs * 8, 256, 0, s>>>(rec, n_rec, lay, counts);
  }
  ++g_launches;
}

// ------------------------------------------------------------------------------------------
// unique-only coverage histogram: one increment per column without redundant reads
// ------------------------------------------------------------------------------------------
// Depths crowd into a few dozen bins, so the increments go to a CTA-private copy of the histogram in shared memory
// (when groups x (max depth + 1) fits) and reach the global bins once per CTA and non-zero bin.
constexpr uint32_t COV_SHARED_BINS = 8192;
__global__ void __launch_bounds__(256) coverage_hist_kernel(const uint64_t* __restrict__ hist_off, const uint8_t* __restrict__ group,
                                                             uint64_t n_cols, uint32_t stride, uint32_t n_bins,
                                                             unsigned long long* __restrict__ cov, uint32_t* __restrict__ err_out) {
  __shared__ uint32_t priv[COV_SHARED_BINS];
  const bool use_shared = n_bins <= COV_SHARED_BINS;
  if (use_shared) {
    for (uint32_t i = threadIdx.x; i < n_bins; i += blockDim.x) priv[i] = 0;
    __syncthreads();
  }
  const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_cols; c += step) {
    const uint64_t a = hist_off[c], b = hist_off[c + 1];
    if (a & HIST_OFF_REDUNDANT_BIT) continue;
    const uint64_t depth = (b & ~HIST_OFF_REDUNDANT_BIT) - a;
    if (depth >= stride) { atomicOr(err_out, BRQ_ERR_DEPTH_RANGE); continue; }
    const uint64_t bin = (uint64_t)group[c] * stride + depth;
    if (use_shared) atomicAdd(&priv[bin], 1u); else atomicAdd(&cov[bin], 1ull);
  }
  if (use_shared) {
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_bins; i += blockDim.x) if (priv[i]) atomicAdd(&cov[i], (unsigned long long)priv[i]);
  }
}

void launch_coverage_hist(const uint64_t* hist_off, const uint8_t* group, uint64_t n_cols, uint32_t stride, uint32_t n_groups,
                          unsigned long long* cov_hist, uint32_t* err, cudaStream_t s) {
  if (!n_cols) return;
  int blocks = (int)std::min<uint64_t>((n_cols + 255) / 256, 148 * 8);
  const uint64_t n_bins = (uint64_t)stride * n_groups;
  coverage_hist_kernel<<<blocks, 256, 0, s>>>(hist_off, group, n_cols, stride, (uint32_t)std::min<uint64_t>(n_bins, 0xFFFFFFFFull), cov_hist, err);
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

// ------------------------------------------------------------------------------------------
// transfer form of the scoring stream -> score_rec (brq_types.h): one warp per round, one lane per slot, the words
// written in the stream's own round-major order (two coalesced 128-bit stores per lane and round vector).  A flagged
// low half takes the lane's next exception word; every other word is decided by its low half and the slot's base.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) expand_score_kernel(const uint16_t* __restrict__ s16, const uint64_t* __restrict__ round_off,
                                                            const uint32_t* __restrict__ round_slot, const uint8_t* __restrict__ slot_ref,
                                                            const uint32_t* __restrict__ exc, const uint32_t* __restrict__ exc_off,
                                                            uint64_t n_rounds, ScoreRecon rc, uint32_t* __restrict__ out) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t n_warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
  for (uint64_t r = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rounds; r += n_warps) {
    const uint64_t beg = round_off[r], n_vec = (round_off[r + 1] - beg) / ROUND_VECTOR_WORDS;
    const uint32_t sl = round_slot[(r << 5) + lane], ref = sl == ROUND_NO_SLOT ? 5u : (uint32_t)slot_ref[sl];
    uint32_t cur = exc_off[(r << 5) + lane];
    const uint16_t* src = s16 + beg + lane * 4u;
    uint32_t* dst = out + beg + lane * 4u;
    auto word = [&](uint32_t v) { return (v & S16_EXCEPTION) ? __ldg(exc + cur++) : score_word_from16(v, ref, rc); };
    auto four = [&](uint2 v) { uint4 o; o.x = word(v.x & 0xFFFFu); o.y = word(v.x >> 16); o.z = word(v.y & 0xFFFFu); o.w = word(v.y >> 16); return o; };
    auto ld8 = [&](const uint16_t* p) { uint2 v; asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p)); return v; };
    uint64_t i = 0;
    for (; i + 4 <= n_vec; i += 4) {  // eight 64-bit loads in flight per lane
      uint2 a[4], b[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { a[k] = ld8(src + (i + k) * ROUND_VECTOR_WORDS); b[k] = ld8(src + (i + k) * ROUND_VECTOR_WORDS + ROUND_VECTOR_WORDS / 2); }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        *reinterpret_cast<uint4*>(dst + (i + k) * ROUND_VECTOR_WORDS) = four(a[k]);
        *reinterpret_cast<uint4*>(dst + (i + k) * ROUND_VECTOR_WORDS + ROUND_VECTOR_WORDS / 2) = four(b[k]);
      }
    }
    for (; i < n_vec; ++i) {
      const uint2 a = ld8(src + i * ROUND_VECTOR_WORDS), b = ld8(src + i * ROUND_VECTOR_WORDS + ROUND_VECTOR_WORDS / 2);
      *reinterpret_cast<uint4*>(dst + i * ROUND_VECTOR_WORDS) = four(a);
      *reinterpret_cast<uint4*>(dst + i * ROUND_VECTOR_WORDS + ROUND_VECTOR_WORDS / 2) = four(b);
    }
  }
}

void launch_expand_score(const uint16_t* s16, const uint64_t* round_off, const uint32_t* round_slot, const uint8_t* slot_ref, const uint32_t* exc,
                         const uint32_t* exc_off, uint64_t n_rounds, const ScoreGeometry& geo, uint32_t* score_rec, cudaStream_t s) {
  if (!n_rounds) return;
  const int blocks = (int)std::min<uint64_t>((n_rounds + 7) / 8, 148 * 8);
  expand_score_kernel<<<blocks, 256, 0, s>>>(s16, round_off, round_slot, slot_ref, exc, exc_off, n_rounds, score_recon_of(geo), score_rec);
  ++g_launches;
}

// The reference writes the table as text with six significant digits and scoring reads it back (error_count.cpp:629-690,
// 1032-1040): prob = 10^(the text's value).  canonical.h computes that value exactly, without the text, so a step has
// no host round trip between the histogram and the likelihood tables.  err is set when a value is outside its range.
__global__ void canonical_table_kernel(const double* __restrict__ log10_prob, uint32_t n, double* __restrict__ prob, uint32_t* __restrict__ err) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool ok = true;
  const double t = text_canonical_6g(log10_prob[i], &ok);
  if (!ok) atomicOr(err, 1u);
  prob[i] = pow(10.0, t);
}

void launch_canonical_table(const double* log10_prob, uint32_t n_bins, double* prob, uint32_t* err, cudaStream_t s) {
  canonical_table_kernel<<<(n_bins + 255) / 256, 256, 0, s>>>(log10_prob, n_bins, prob, err);
  ++g_launches;
}

void launch_derive_table(const unsigned long long* counts, const CovLayout& lay, double* log10_prob, cudaStream_t s) {
  derive_table_kernel<<<(lay.n_bins + 255) / 256, 256, 0, s>>>(counts, lay, log10_prob);
  ++g_launches;
}

}  // namespace brq
