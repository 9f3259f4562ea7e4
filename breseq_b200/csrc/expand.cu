// Device-side staging for sm_100a: the kernels around expand_core.h and the host sequence that runs them.
//
//   prep_reads_kernel      one thread per read: ReadMeta (reference span, query bounds, read file, flags), the longest span per
//                          target, the MAPQ statistics that pick the shared table's MAPQ, the sort check
//   qual_stats_kernel      a warp per read: quality histogram of the unique reads (picks the table's quality window)
//   ins_support_kernel     one thread per read with an insertion: which insert sub-columns exist (atomicOr of a 64-bit mask)
//   slot_ref_kernel        reference characters -> base indices, coverage groups
//   tile_kernel<false>     COUNT pass: a warp per tile of 32 columns, a lane per column, candidate reads in BAM order
//   order_*_kernel         the tally kernel's rounds: inside every block of 4096 slots, slots grouped by reference base and
//                          ordered by depth (a bitonic sort of 4096 keys in shared memory), groups padded to whole rounds
//   scan kernels           exclusive prefix sums (three passes: block sums, scan of the sums, write) for every offset array
//   tile_kernel<true>      FILL pass: the same walk, every record written at its final place in the round-major stream
//   hist16_*_kernel        stable partition of the histogram records into the 16-bit fast form and the exceptions
//
// All of it is integer / byte work bound by memory latency and the L2; nothing here is a tensor-core shape.
#include "expand.h"
#include "expand_plan.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <type_traits>
#include <vector>

namespace brq {

void note_launches(int n);
static int g_expand_launches = 0;
int expand_launch_count() { return g_expand_launches; }
static inline void launched(int n = 1) { g_expand_launches += n; note_launches(n); }

// ------------------------------------------------------------------------------------------ buffers
// the staging buffers belong to the process, not to a context: a fresh context on the same device finds them allocated
// (page-locking 64 MB costs tens of milliseconds, a whole C1 upload costs little more)
namespace {
struct RingSlots { void* slot[UploadRing::SLOTS] = {nullptr}; cudaEvent_t done[UploadRing::SLOTS] = {nullptr}; int next = 0; std::mutex mu; };
RingSlots& ring_of_current_device() {
  static std::mutex mu;
  static std::map<int, RingSlots*> rings;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> g(mu);
  RingSlots*& r = rings[dev];
  if (!r) r = new RingSlots;
  return *r;
}
}  // namespace

void UploadRing::copy(void* dst_device, const void* src_host, size_t bytes, cudaStream_t s) {
  if (!bytes) return;
  // brq_pin_reads registers an array in pieces, and the system may refuse some: every quarter-gigabyte span is looked at
  // on its own (copied directly when both of its ends are page-locked, through the ring otherwise)
  const size_t SPAN = (size_t)256 << 20;
  if (bytes > SPAN) {
    for (size_t at = 0; at < bytes; at += SPAN)
      copy(static_cast<char*>(dst_device) + at, static_cast<const char*>(src_host) + at, std::min(SPAN, bytes - at), s);
    return;
  }
  RingSlots& R = ring_of_current_device();
  std::lock_guard<std::mutex> ring_guard(R.mu);
  void** slot = R.slot;
  cudaEvent_t* done = R.done;
  int& next = R.next;
  auto page_locked = [](const void* p) {
    cudaPointerAttributes attr;
    const bool yes = cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    return yes;
  };
  if (bytes < ((size_t)1 << 20)) {  // too small to matter
    CUDA_OK(cudaMemcpyAsync(dst_device, src_host, bytes, cudaMemcpyHostToDevice, s));
    return;
  }
  if (page_locked(src_host) && page_locked(static_cast<const char*>(src_host) + bytes - 1)) {
    CUDA_OK(cudaMemcpyAsync(dst_device, src_host, bytes, cudaMemcpyHostToDevice, s));
    return;
  }
  for (size_t at = 0; at < bytes; at += SLOT_BYTES) {
    const size_t len = std::min(SLOT_BYTES, bytes - at);
    const int k = next;
    next = (next + 1) % SLOTS;
    if (!slot[k]) {
      CUDA_OK(cudaHostAlloc(&slot[k], SLOT_BYTES, cudaHostAllocDefault));
      CUDA_OK(cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming));
    } else {
      CUDA_OK(cudaEventSynchronize(done[k]));   // the copy that last used this slot has left it
    }
    const char* src = static_cast<const char*>(src_host) + at;
    char* stage = static_cast<char*>(slot[k]);
    if (parallel_for) {
      const size_t parts = 16, step = (len + parts - 1) / parts;
      parallel_for(parts, [&](size_t p) { const size_t a = p * step; if (a < len) memcpy(stage + a, src + a, std::min(step, len - a)); });
    } else {
      memcpy(stage, src, len);
    }
    CUDA_OK(cudaMemcpyAsync(static_cast<char*>(dst_device) + at, stage, len, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaEventRecord(done[k], s));
  }
}
void UploadRing::release() {}   // (the buffers are the process's: see ring_of_current_device)

void ReadsDev::upload(const ReadBatch& R, cudaStream_t s) {
  n = R.size();
  bytes = 0;
  auto up = [&](auto& buf, const auto& vec) {
    using T = typename std::remove_reference<decltype(*buf.p)>::type;
    buf.ensure(vec.size() + 16);
    ring.copy(buf.p, vec.data(), vec.size() * sizeof(T), s);
    bytes += vec.size() * sizeof(T);
  };
  up(tid, R.tid); up(pos, R.pos); up(flag, R.flag); up(mapq, R.mapq); up(rg, R.rg); up(x1, R.x1); up(xl, R.xl); up(xr, R.xr);
  up(l_seq, R.l_seq); up(seq_off, R.seq_off); up(n_cigar, R.n_cigar); up(cigar_off, R.cigar_off);
  up(bases, R.bases); up(quals, R.quals); up(cigars, R.cigars);
}
RawReads ReadsDev::view() const {
  RawReads r;
  r.tid = tid.p; r.pos = pos.p; r.flag = flag.p; r.mapq = mapq.p; r.rg = rg.p; r.x1 = x1.p; r.xl = xl.p; r.xr = xr.p;
  r.l_seq = l_seq.p; r.seq_off = seq_off.p; r.n_cigar = n_cigar.p; r.cigar_off = cigar_off.p;
  r.bases = bases.p; r.quals = quals.p; r.cigars = cigars.p; r.n = n;
  return r;
}
void ReadsDev::release() {
  tid.release(); pos.release(); xl.release(); xr.release(); flag.release(); mapq.release(); rg.release(); bases.release(); quals.release();
  x1.release(); l_seq.release(); n_cigar.release(); cigars.release(); seq_off.release(); cigar_off.release();
  ring.release();
  n = bytes = 0;
}
void StreamDev::release() {
  score_rec.release(); side_rec.release(); side_off.release(); round_slot.release(); score_cnt.release(); round_side.release();
  score_off.release(); hist_off.release(); round_off.release(); hist_rec.release(); slot_ref.release(); slot_group.release();
}
void ExpandScratch::release() {
  meta.release(); seg.release(); max_span.release(); seg_of_tid.release(); ref.release(); sub_k.release(); col_red.release(); col_qstart.release();
  part.release(); stats.release(); sub_first.release(); red_cnt.release(); side_cnt.release(); side_red_cnt.release(); hist_cnt.release();
  sub_cur.release(); block_entries.release(); block_base.release(); round_vecs.release(); scan_tmp32.release(); ins_mask.release();
  geo_stats.release(); scan_tmp.release(); ins_parent.release(); totals.release(); read_starts.release(); ins_count.release(); hist_pos.release();
  have_walk_args = false;
}

namespace {

// ------------------------------------------------------------------------------------------ per-read kernels
__global__ void __launch_bounds__(256) prep_reads_kernel(RawReads R, const uint32_t* __restrict__ part_base, const uint32_t* __restrict__ part_count,
                                                          uint32_t n_part, ReadMeta* __restrict__ meta, int32_t* max_span, uint32_t* stats,
                                                          unsigned long long* geo_stats, uint32_t want_geo) {
  __shared__ unsigned long long mq_sh[256];
  mq_sh[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < R.n) {
    const ReadMeta m = prep_read(R, i, part_base, part_count, n_part, max_span, stats);
    meta[i] = m;
    if (want_geo && (m.flags & RM_LIVE) && m.x1 == 1) atomicAdd(&mq_sh[m.mapq], (unsigned long long)m.l_seq);
  }
  __syncthreads();
  if (mq_sh[threadIdx.x]) atomicAdd(&geo_stats[threadIdx.x], mq_sh[threadIdx.x]);
}

__global__ void __launch_bounds__(256) qual_stats_kernel(const ReadMeta* __restrict__ meta, uint64_t n, const uint8_t* __restrict__ quals,
                                                          unsigned long long* geo_stats) {
  __shared__ uint32_t h[8][128];
  for (uint32_t j = threadIdx.x; j < 8 * 128; j += blockDim.x) (&h[0][0])[j] = 0;
  __syncthreads();
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint64_t n_warps = (uint64_t)gridDim.x * 8;
  for (uint64_t r = (uint64_t)blockIdx.x * 8 + warp; r < n; r += n_warps) {
    const uint8_t flags = meta[r].flags;
    if (!(flags & RM_LIVE) || meta[r].x1 != 1) continue;
    const uint8_t* q = quals + meta[r].seq_off;
    const uint32_t L = meta[r].l_seq;
    for (uint32_t k = lane; k < L; k += 32) atomicAdd(&h[warp][q[k] & 127u], 1u);
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    unsigned long long s = 0;
    for (int w = 0; w < 8; ++w) s += h[w][threadIdx.x];
    if (s) atomicAdd(&geo_stats[256 + threadIdx.x], s);
  }
}

__global__ void __launch_bounds__(256) ins_support_kernel(ExpandArgs a) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.n_reads) ins_support(a, i);
}

// ------------------------------------------------------------------------------------------ per-slot kernels
__global__ void __launch_bounds__(256) slot_ref_kernel(ExpandArgs a, uint8_t* __restrict__ slot_group) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.n_base) return;
  uint32_t lo = 0, hi = a.n_seg;
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.seg[mid].slot0 <= s) lo = mid; else hi = mid; }
  const ExpandSeg& sg = a.seg[lo];
  const uint8_t b = xchar_to_index(a.ref[sg.ref_off + (s - sg.slot0)]);
  if (b > 5 || b == 4) atomicOr(&a.stats[XS_ERR], EXP_ERR_REFCHAR);
  a.slot_ref[s] = b > 5 ? (uint8_t)kBaseN : b;
  slot_group[s] = (uint8_t)sg.group;
}

__global__ void __launch_bounds__(256) sub_k_kernel(const uint64_t* __restrict__ ins_mask, uint32_t n_base, uint8_t* __restrict__ sub_k) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_base) return;
  const uint64_t m = ins_mask[s];
  uint32_t K = m == ~0ull ? 64u : (uint32_t)__ffsll((long long)~m) - 1u;  // trailing ones
  sub_k[s] = (uint8_t)(K < 63u ? K : 63u);
}

__global__ void __launch_bounds__(256) fill_ins_kernel(const uint8_t* __restrict__ sub_k, const uint32_t* __restrict__ sub_first, uint32_t n_base,
                                                        uint64_t* __restrict__ ins_parent, uint32_t* __restrict__ ins_count, uint8_t* __restrict__ slot_ref) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_base) return;
  const uint32_t K = sub_k[s], f = sub_first[s];
  for (uint32_t k = 1; k <= K; ++k) { ins_parent[f + k - 1] = s; ins_count[f + k - 1] = k; slot_ref[n_base + f + k - 1] = kBaseGap; }
}

__global__ void __launch_bounds__(256) sub_cursor_kernel(const uint32_t* __restrict__ red_cnt, const uint32_t* __restrict__ side_red_cnt,
                                                          uint32_t n_base, uint32_t n_ins, uint32_t* __restrict__ sub_cur) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_ins) return;
  sub_cur[4 * j] = red_cnt[n_base + j]; sub_cur[4 * j + 1] = 0; sub_cur[4 * j + 2] = side_red_cnt[n_base + j]; sub_cur[4 * j + 3] = 0;
}

// preprocess stage (error_count.cpp:191-194): per target, the position-strand combinations without / with a read start,
// over the columns without a redundant read
__global__ void __launch_bounds__(256) read_starts_kernel(ExpandArgs a, unsigned long long* __restrict__ counts) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.n_base || a.col_red[s]) return;
  uint32_t lo = 0, hi = a.n_seg;
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.seg[mid].slot0 <= s) lo = mid; else hi = mid; }
  const uint32_t tid = (uint32_t)a.seg[lo].tid, qs = a.col_qstart[s];
  atomicAdd(&counts[tid * 2 + (qs & 1u)], 1ull);
  atomicAdd(&counts[tid * 2 + ((qs >> 1) & 1u)], 1ull);
}

// ------------------------------------------------------------------------------------------ the tile passes
// (the count pass fits 64 registers: four CTAs per SM; the fill pass spills there and runs faster with three)
__global__ void __launch_bounds__(256, 4) coverage_tile_kernel(ExpandArgs a, CoverageColumn* __restrict__ out, uint32_t group, bool include_deleted) {
  const uint32_t tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (tile >= a.n_tiles) return;
  coverage_lane(a, tile, threadIdx.x & 31u, out, group, include_deleted);
}

template <bool FILL>
__global__ void __launch_bounds__(256, FILL ? 3 : 4) tile_kernel(ExpandArgs a) {
  const uint32_t tile = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (tile >= a.n_tiles) return;
  tile_lane<FILL>(a, tile, threadIdx.x & 31u);
}

__global__ void __launch_bounds__(256) fill_words_kernel(uint32_t* __restrict__ p, uint64_t n, uint32_t v) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint4* p4 = reinterpret_cast<uint4*>(p);
  const uint64_t n4 = n / 4;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) p4[i] = make_uint4(v, v, v, v);
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) p[n4 * 4 + threadIdx.x] = v;
}

// ------------------------------------------------------------------------------------------ exclusive scans
// out[i] = sum of f(j) for j < i, three passes: sums of blocks of SCAN_BLOCK items, scan of the sums (one thread: a few
// thousand values), and the write pass.  out may have n + 1 entries (WRITE_TOTAL): out[n] = the total.
constexpr uint32_t SCAN_TPB = 256, SCAN_ITEMS = 16, SCAN_BLOCK = SCAN_TPB * SCAN_ITEMS;

struct FU8 { const uint8_t* p; __device__ uint64_t operator()(uint64_t i) const { return p[i]; } };
struct FU32 { const uint32_t* p; __device__ uint64_t operator()(uint64_t i) const { return p[i]; } };
struct FSideEven { const uint32_t* p; __device__ uint64_t operator()(uint64_t i) const { return (p[i] + 1u) & ~1u; } };
struct FVecs256 { const uint32_t* p; __device__ uint64_t operator()(uint64_t i) const { return (uint64_t)p[i] * ROUND_VECTOR_WORDS; } };

__device__ __forceinline__ uint64_t block_exclusive(uint64_t v, uint64_t& total) {  // exclusive scan of one value per thread
  __shared__ uint64_t warp_sum[SCAN_TPB / 32];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint64_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint64_t y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= (uint32_t)o) x += y; }
  if (lane == 31) warp_sum[warp] = x;
  __syncthreads();
  uint64_t before = 0, all = 0;
  for (uint32_t w = 0; w < SCAN_TPB / 32; ++w) { const uint64_t s = warp_sum[w]; if (w < warp) before += s; all += s; }
  __syncthreads();
  total = all;
  return before + x - v;
}

template <class F>
__global__ void __launch_bounds__(SCAN_TPB) scan_sums_kernel(F f, uint64_t n, unsigned long long* __restrict__ block_sums) {
  const uint64_t b0 = (uint64_t)blockIdx.x * SCAN_BLOCK;
  uint64_t s = 0;
  for (uint32_t k = 0; k < SCAN_ITEMS; ++k) { const uint64_t i = b0 + (uint64_t)k * SCAN_TPB + threadIdx.x; if (i < n) s += f(i); }
  uint64_t total;
  block_exclusive(s, total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}
__global__ void scan_blocks_kernel(unsigned long long* __restrict__ block_sums, uint32_t n_blocks, unsigned long long* __restrict__ total) {
  if (threadIdx.x || blockIdx.x) return;
  unsigned long long acc = 0;
  for (uint32_t b = 0; b < n_blocks; ++b) { const unsigned long long v = block_sums[b]; block_sums[b] = acc; acc += v; }
  *total = acc;
}
template <class F, class TOut, bool WRITE_TOTAL>
__global__ void __launch_bounds__(SCAN_TPB) scan_write_kernel(F f, uint64_t n, const unsigned long long* __restrict__ block_offs,
                                                               const unsigned long long* __restrict__ total, TOut* __restrict__ out) {
  // thread t owns items b0 + t * SCAN_ITEMS .. + SCAN_ITEMS - 1 (consecutive: the scan order is the array order)
  const uint64_t i0 = (uint64_t)blockIdx.x * SCAN_BLOCK + (uint64_t)threadIdx.x * SCAN_ITEMS;
  uint64_t v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (uint32_t k = 0; k < SCAN_ITEMS; ++k) { v[k] = i0 + k < n ? f(i0 + k) : 0; s += v[k]; }
  uint64_t tot;
  uint64_t acc = block_exclusive(s, tot) + block_offs[blockIdx.x];
#pragma unroll
  for (uint32_t k = 0; k < SCAN_ITEMS; ++k) { if (i0 + k < n) out[i0 + k] = (TOut)acc; acc += v[k]; }
  if (WRITE_TOTAL && blockIdx.x == 0 && threadIdx.x == 0) out[n] = (TOut)*total;
}

// runs the three passes; d_total (one word of scratch.totals) receives the total
template <class F, class TOut, bool WRITE_TOTAL>
void exclusive_scan(F f, uint64_t n, TOut* out, DevBuf<uint64_t>& tmp, unsigned long long* d_total, cudaStream_t s) {
  const uint32_t n_blocks = (uint32_t)((n + SCAN_BLOCK - 1) / SCAN_BLOCK);
  tmp.ensure(std::max<uint32_t>(n_blocks, 1) + 1);
  unsigned long long* sums = reinterpret_cast<unsigned long long*>(tmp.p);
  if (n_blocks) scan_sums_kernel<F><<<n_blocks, SCAN_TPB, 0, s>>>(f, n, sums);
  scan_blocks_kernel<<<1, 32, 0, s>>>(sums, n_blocks, d_total);
  if (out) {
    if (n_blocks) scan_write_kernel<F, TOut, WRITE_TOTAL><<<n_blocks, SCAN_TPB, 0, s>>>(f, n, sums, d_total, out);
    else if (WRITE_TOTAL) CUDA_OK(cudaMemsetAsync(out, 0, sizeof(TOut), s));
  }
  launched(out ? 3 : 2);
}

__global__ void __launch_bounds__(256) hist_off_flag_kernel(uint64_t* __restrict__ hist_off, const uint8_t* __restrict__ col_red, uint32_t n_base) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n_base && col_red[c]) hist_off[c] |= HIST_OFF_REDUNDANT_BIT;
}
__global__ void __launch_bounds__(256) max_u32_kernel(const uint32_t* __restrict__ v, uint32_t n, uint32_t* __restrict__ out) {
  uint32_t m = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, v[i]);
  m = __reduce_max_sync(0xFFFFFFFFu, m);
  if ((threadIdx.x & 31u) == 0 && m) atomicMax(out, m);
}

// ------------------------------------------------------------------------------------------ rounds of the tally kernel
// Inside every block of ROUND_BLOCK consecutive slots: groups by reference base (A, C, G, T, other), each ordered by
// (vectors, slot) and padded to whole rounds with ROUND_NO_SLOT (staging.cpp "offsets and rounds": a stable sort by vectors
// is the sort by (vectors, slot)).
__device__ __forceinline__ uint32_t group_of(uint8_t ref) { return ref < 4 ? ref : 4u; }

__global__ void __launch_bounds__(256) order_count_kernel(const uint8_t* __restrict__ slot_ref, uint32_t n_slots, uint32_t* __restrict__ block_entries) {
  __shared__ uint32_t cnt[5];
  if (threadIdx.x < 5) cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t b0 = blockIdx.x * ROUND_BLOCK;
  uint32_t mine[5] = {0, 0, 0, 0, 0};
  for (uint32_t i = threadIdx.x; i < ROUND_BLOCK && b0 + i < n_slots; i += blockDim.x) ++mine[group_of(slot_ref[b0 + i])];
#pragma unroll
  for (int g = 0; g < 5; ++g) { const uint32_t s = __reduce_add_sync(0xFFFFFFFFu, mine[g]); if ((threadIdx.x & 31u) == 0 && s) atomicAdd(&cnt[g], s); }
  __syncthreads();
  if (threadIdx.x == 0) { uint32_t e = 0; for (int g = 0; g < 5; ++g) e += (cnt[g] + 31u) & ~31u; block_entries[blockIdx.x] = e; }
}

__global__ void __launch_bounds__(1024) order_sort_kernel(const uint8_t* __restrict__ slot_ref, const uint32_t* __restrict__ score_cnt, uint32_t n_slots,
                                                           const uint32_t* __restrict__ block_base, uint32_t* __restrict__ round_slot,
                                                           uint32_t* __restrict__ round_vecs, uint32_t* stats) {
  __shared__ uint32_t key[ROUND_BLOCK];
  __shared__ uint32_t gstart[6], gpad[6];
  const uint32_t b0 = blockIdx.x * ROUND_BLOCK;
  for (uint32_t i = threadIdx.x; i < ROUND_BLOCK; i += blockDim.x) {
    uint32_t k = 0xFFFFFFFFu;
    if (b0 + i < n_slots) {
      const uint32_t v = (score_cnt[b0 + i] + 7u) >> 3;
      if (v > 0xFFFFu) atomicOr(&stats[XS_ERR], EXP_ERR_DEPTH);
      k = group_of(slot_ref[b0 + i]) << 28 | min(v, 0xFFFFu) << 12 | i;
    }
    key[i] = k;
  }
  __syncthreads();
  for (uint32_t size = 2; size <= ROUND_BLOCK; size <<= 1)
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      for (uint32_t t = threadIdx.x; t < ROUND_BLOCK / 2; t += blockDim.x) {
        const uint32_t lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
        const bool up = (lo & size) == 0;
        const uint32_t a = key[lo], b = key[hi];
        if ((a > b) == up) { key[lo] = b; key[hi] = a; }
      }
      __syncthreads();
    }
  if (threadIdx.x < 6) {  // first key of group g (keys of group g: [g << 28, (g + 1) << 28); unused entries sort last)
    const uint32_t want = threadIdx.x << 28;
    uint32_t lo = 0, hi = ROUND_BLOCK;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (key[mid] >= want) hi = mid; else lo = mid + 1; }
    gstart[threadIdx.x] = lo;
  }
  __syncthreads();
  if (threadIdx.x == 0) { uint32_t e = 0; for (int g = 0; g < 5; ++g) { gpad[g] = e; e += (gstart[g + 1] - gstart[g] + 31u) & ~31u; } gpad[5] = e; }
  __syncthreads();
  const uint32_t base = block_base[blockIdx.x], n_entries = gpad[5];
  for (uint32_t j = threadIdx.x; j < n_entries; j += blockDim.x) {
    uint32_t g = 0;
    while (j >= gpad[g + 1]) ++g;
    const uint32_t r = j - gpad[g], size = gstart[g + 1] - gstart[g];
    round_slot[base + j] = r < size ? b0 + (key[gstart[g] + r] & 0xFFFu) : ROUND_NO_SLOT;
    if ((j & 31u) == 0) {  // the round's deepest slot: its last real lane (ascending order)
      const uint32_t last = min(r + 31u, size - 1u);
      round_vecs[(base + j) >> 5] = (key[gstart[g] + last] >> 12) & 0xFFFFu;
    }
  }
}

// what a lane needs of its slot besides the records, round-major; the first word of every slot
__global__ void __launch_bounds__(256) assign_kernel(const uint32_t* __restrict__ round_slot, uint64_t n_entries, const uint64_t* __restrict__ round_off,
                                                      const uint32_t* __restrict__ side_off, const uint32_t* __restrict__ side_cnt,
                                                      const uint8_t* __restrict__ slot_ref, uint32_t want_score, uint64_t* __restrict__ score_off,
                                                      uint32_t* __restrict__ round_side) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_entries) return;
  const uint32_t sl = round_slot[i];
  if (sl == ROUND_NO_SLOT) { round_side[2 * i] = 5u << 29; round_side[2 * i + 1] = 0; return; }
  score_off[sl] = round_off[i >> 5] + (i & 31u) * 4u;
  round_side[2 * i] = side_off[sl] | (uint32_t)slot_ref[sl] << 29;
  round_side[2 * i + 1] = side_off[sl] + (want_score ? side_cnt[sl] : 0u);
}

// ------------------------------------------------------------------------------------------ compact histogram stream
// stable partition of the positional 4-byte records: the fast ones as 16 bits (brq_types.h: hist16_pack), the others unchanged
__device__ __forceinline__ uint32_t pack16_dev(uint32_t lo) {
  if (!(lo >> HR_FAST)) return 0x10000u;
  const uint32_t qa = (lo >> HR_QUALA) & 127u, qb = (lo >> HR_QUALB) & 127u, set = (lo >> HR_SET) & 7u;
  if (qa > 62u || (qb > 62u && qb != 127u) || set > 3u) return 0x10000u;
  const uint32_t r = (lo & 3u) | qa << 2 | (qb == 127u ? 63u : qb) << 8 | set << 14;
  return hist16_expand(r) == lo ? r : 0x10000u;
}
__global__ void __launch_bounds__(SCAN_TPB) hist16_count_kernel(const uint32_t* __restrict__ h, uint64_t n, uint32_t* __restrict__ block_counts) {
  const uint64_t b0 = (uint64_t)blockIdx.x * SCAN_BLOCK;
  uint64_t s = 0;
  for (uint32_t k = 0; k < SCAN_ITEMS; ++k) { const uint64_t i = b0 + (uint64_t)k * SCAN_TPB + threadIdx.x; if (i < n) s += pack16_dev(h[i]) < 0x10000u; }
  uint64_t total;
  block_exclusive(s, total);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = (uint32_t)total;
}
__global__ void __launch_bounds__(SCAN_TPB) hist16_scatter_kernel(const uint32_t* __restrict__ h, uint64_t n, const uint64_t* __restrict__ block_base16,
                                                                   uint16_t* __restrict__ out16, uint32_t* __restrict__ out_exc) {
  const uint64_t b0 = (uint64_t)blockIdx.x * SCAN_BLOCK, i0 = b0 + (uint64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t r[SCAN_ITEMS], w[SCAN_ITEMS];
  uint64_t s = 0;
#pragma unroll
  for (uint32_t k = 0; k < SCAN_ITEMS; ++k) { w[k] = i0 + k < n ? h[i0 + k] : 0u; r[k] = i0 + k < n ? pack16_dev(w[k]) : 0x10000u; s += r[k] < 0x10000u; }
  uint64_t tot;
  uint64_t fast_before = block_exclusive(s, tot) + block_base16[blockIdx.x];   // fast records before this thread's first item
  uint64_t exc_before = i0 - fast_before;
#pragma unroll
  for (uint32_t k = 0; k < SCAN_ITEMS; ++k) {
    if (i0 + k >= n) break;
    if (r[k] < 0x10000u) out16[fast_before++] = (uint16_t)r[k]; else out_exc[exc_before++] = w[k];
  }
}

// ------------------------------------------------------------------------------------------ flagged slots' records
struct SlotInfo { uint64_t off; uint32_t cnt, side0, side1, ref; };
__global__ void slot_info_kernel(const uint32_t* __restrict__ slots, uint32_t n, const uint64_t* __restrict__ score_off, const uint32_t* __restrict__ score_cnt,
                                 const uint32_t* __restrict__ side_off, const uint8_t* __restrict__ slot_ref, SlotInfo* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = slots[i];
  out[i] = SlotInfo{score_off[s], score_cnt[s], side_off[s], side_off[s + 1], slot_ref[s]};
}
__global__ void __launch_bounds__(256) gather_records_kernel(const SlotInfo* __restrict__ info, uint32_t n, const uint64_t* __restrict__ word_off,
                                                              const uint64_t* __restrict__ side_at, const uint32_t* __restrict__ score_rec,
                                                              const uint32_t* __restrict__ side_rec, uint32_t side_stride,
                                                              uint32_t* __restrict__ words, uint32_t* __restrict__ side) {
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += (gridDim.x * blockDim.x) >> 5) {
    const SlotInfo f = info[i];
    for (uint32_t j = lane; j < f.cnt; j += 32) words[word_off[i] + j] = score_rec[score_index(f.off, j)];
    const uint32_t ns = (f.side1 - f.side0) * side_stride;
    for (uint32_t j = lane; j < ns; j += 32) side[side_at[i] * side_stride + j] = side_rec[(size_t)f.side0 * side_stride + j];
  }
}

const char* expand_error_text(uint32_t err) {
  if (err & EXP_ERR_UNSORTED) return "BAM is not coordinate sorted";
  if (err & EXP_ERR_READSET32) return "more than 32 read files are not supported by the packed record";
  if (err & EXP_ERR_REFCHAR) return "Unrecognized base char in reference";
  if (err & EXP_ERR_INS63) return "insertions longer than 63 bases are not supported";
  if (err & EXP_ERR_CIGAR_LONG) return "CIGAR longer than the read sequence";
  if (err & EXP_ERR_X1ZERO) return "X1:i:0 is not a valid redundancy";
  if (err & EXP_ERR_QUAL127) return "base quality above 127 cannot be packed";
  if (err & EXP_ERR_DEL_NO_BASE) return "deletion with no following read base (reference would assert)";
  if (err & EXP_ERR_RPOS) return "read position above 65535 cannot be packed";
  if (err & EXP_ERR_NOQUAL_DOTN) return "Attempt to retrieve quality score for nonexistent base for '.N' state.";
  if (err & EXP_ERR_NOQUAL) return "Attempt to retrieve quality score for nonexistent base.";
  if (err & EXP_ERR_DEPTH) return "a column deeper than 524 280 records cannot be staged";
  return "device staging failed";
}

inline uint32_t blocks_for(uint64_t n, uint32_t tpb = 256) { return (uint32_t)std::max<uint64_t>(1, (n + tpb - 1) / tpb); }

}  // namespace

// ------------------------------------------------------------------------------------------ the sequence
void expand_on_device(const BamHeader& hdr, const RefSet& ref, const ReadBatch& host, const ReadsDev& reads, const StageConfig& cfg,
                      ExpandScratch& X, StreamDev& out, PileupStream& st, cudaStream_t s) {
  const size_t n_targets = hdr.target_names.size();
  X.have_walk_args = false;
  // BRQ_STAGE_TIMES=1: wall time of every phase on stderr (each ends in a synchronisation then)
  static const bool phase_times = getenv("BRQ_STAGE_TIMES") != nullptr;
  auto phase_t0 = std::chrono::steady_clock::now();
  auto phase_done = [&](const char* what) {
    if (!phase_times) return;
    cudaStreamSynchronize(s);
    const auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "expand: %-24s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t - phase_t0).count());
    phase_t0 = t;
  };
  st = PileupStream();
  st.device_built = true;
  ExpandPlan plan;
  make_expand_plan(hdr, ref, host.tid.data(), host.tid.size(), cfg, st, plan);
  const std::vector<ExpandSeg>& segs = plan.segs;
  const std::vector<uint8_t>& refbytes = plan.refbytes;
  const std::vector<int32_t>& seg_of_tid = plan.seg_of_tid;
  const std::vector<uint32_t>& part = plan.part;
  const uint32_t n_part = plan.n_part, tiles = plan.tiles;
  const size_t n_visit = st.segments.size();
  const uint32_t n_base = (uint32_t)st.n_base;
  const uint64_t n_reads = reads.n;
  if (n_reads != host.tid.size()) throw std::runtime_error("the reads in HBM are not the host's read batch");

  X.seg.ensure(n_visit + 1); X.ref.ensure(refbytes.size()); X.seg_of_tid.ensure(seg_of_tid.size()); X.max_span.ensure(n_targets + 1);
  X.part.ensure(part.size()); X.stats.ensure(XS_WORDS); X.geo_stats.ensure(384); X.totals.ensure(16); X.meta.ensure(n_reads + 1);
  if (n_visit) CUDA_OK(cudaMemcpyAsync(X.seg.p, segs.data(), n_visit * sizeof(ExpandSeg), cudaMemcpyHostToDevice, s));
  CUDA_OK(cudaMemcpyAsync(X.ref.p, refbytes.data(), refbytes.size(), cudaMemcpyHostToDevice, s));
  CUDA_OK(cudaMemcpyAsync(X.seg_of_tid.p, seg_of_tid.data(), seg_of_tid.size() * 4, cudaMemcpyHostToDevice, s));
  CUDA_OK(cudaMemcpyAsync(X.part.p, part.data(), part.size() * 4, cudaMemcpyHostToDevice, s));
  {
    std::vector<int32_t> ones(n_targets + 1, 1);
    CUDA_OK(cudaMemcpyAsync(X.max_span.p, ones.data(), ones.size() * 4, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaStreamSynchronize(s));  // (the host vectors above go out of scope)
  }
  CUDA_OK(cudaMemsetAsync(X.stats.p, 0, XS_WORDS * 4, s));
  CUDA_OK(cudaMemsetAsync(X.geo_stats.p, 0, 384 * 8, s));
  CUDA_OK(cudaMemsetAsync(X.totals.p, 0, 16 * 8, s));

  phase_done("plan and small uploads");
  // ---- per read
  const RawReads raw = reads.view();
  if (n_reads) {
    prep_reads_kernel<<<blocks_for(n_reads), 256, 0, s>>>(raw, X.part.p, X.part.p + n_part, n_part, X.meta.p, X.max_span.p, X.stats.p,
                                                          reinterpret_cast<unsigned long long*>(X.geo_stats.p), cfg.want_score ? 1u : 0u);
    launched();
    if (cfg.want_score) { qual_stats_kernel<<<148 * 8, 256, 0, s>>>(X.meta.p, n_reads, reads.quals.p, reinterpret_cast<unsigned long long*>(X.geo_stats.p)); launched(); }
  }

  ExpandArgs a;
  memset(static_cast<void*>(&a), 0, sizeof a);
  a.meta = X.meta.p; a.pos = reads.pos.p; a.bases = reads.bases.p; a.quals = reads.quals.p; a.cigars = reads.cigars.p; a.n_reads = n_reads;
  a.seg = X.seg.p; a.n_seg = (uint32_t)n_visit; a.n_tiles = tiles; a.max_span = X.max_span.p; a.seg_of_tid = X.seg_of_tid.p; a.tid = reads.tid.p;
  a.ref = X.ref.p; a.n_base = n_base;
  a.want_hist = cfg.want_hist; a.want_score = cfg.want_score; a.use_read_pos = cfg.use_read_pos; a.use_base_repeat = cfg.use_base_repeat;
  a.preprocess = cfg.preprocess_stage;
  a.unmatched_end_minimum_read_length = cfg.unmatched_end_minimum_read_length; a.unmatched_end_length_factor = cfg.unmatched_end_length_factor;
  a.stats = X.stats.p;

  // ---- base slots: reference bases, coverage groups, insert sub-columns
  X.ins_mask.ensure((size_t)n_base + 1); X.sub_k.ensure((size_t)n_base + 1); X.sub_first.ensure((size_t)n_base + 2);
  out.slot_group.ensure((size_t)n_base + 1);
  CUDA_OK(cudaMemsetAsync(X.ins_mask.p, 0, ((size_t)n_base + 1) * 8, s));
  a.ins_mask = X.ins_mask.p; a.sub_k = X.sub_k.p; a.sub_first = X.sub_first.p;
  unsigned long long* d_tot = reinterpret_cast<unsigned long long*>(X.totals.p);
  if (n_reads && n_visit) { ins_support_kernel<<<blocks_for(n_reads), 256, 0, s>>>(a); launched(); }
  sub_k_kernel<<<blocks_for(n_base), 256, 0, s>>>(X.ins_mask.p, n_base, X.sub_k.p);
  launched();
  if (!cfg.user_evidence.empty()) {
    // user evidence: the columns where the pileup meets the list and the insert levels it forces there (staging.cpp does the
    // same with its host masks): a handful of columns, each read back and patched on its own
    CUDA_OK(cudaStreamSynchronize(s));
    st.user_list = cfg.user_evidence;
    auto slot_of = [&](size_t v, uint32_t pos1) -> uint64_t {
      for (const Segment& sg : st.segments)
        if (sg.tid == st.visit_targets[v].tid && (int32_t)pos1 - 1 >= sg.lo && (int32_t)pos1 - 1 < sg.hi) return sg.slot0 + (uint64_t)((int32_t)pos1 - 1 - sg.lo);
      return ~0ull;
    };
    auto mask_at = [&](uint64_t slot) { uint64_t m = 0; CUDA_OK(cudaMemcpy(&m, X.ins_mask.p + slot, 8, cudaMemcpyDeviceToHost)); return m; };
    st.user_columns = plan_user_evidence(cfg.user_evidence, hdr, st.visit_targets, cfg.user_skip_cutoff, [&](size_t v, uint32_t pos1, uint32_t force_max) {
      const uint64_t slot = slot_of(v, pos1);
      return slot == ~0ull ? force_max : insert_levels(mask_at(slot), force_max);
    });
    for (UserColumn& c : st.user_columns) {
      size_t v = 0;
      while (v < st.visit_targets.size() && st.visit_targets[v].tid != c.tid) ++v;
      c.slot = slot_of(v, c.pos1);
      if (c.slot == ~0ull || !c.force_max) continue;
      uint8_t have = 0;
      CUDA_OK(cudaMemcpy(&have, X.sub_k.p + c.slot, 1, cudaMemcpyDeviceToHost));
      const uint8_t K = (uint8_t)std::max<uint32_t>(have, insert_levels(mask_at(c.slot), c.force_max));
      CUDA_OK(cudaMemcpy(X.sub_k.p + c.slot, &K, 1, cudaMemcpyHostToDevice));
    }
  }
  exclusive_scan<FU8, uint32_t, true>(FU8{X.sub_k.p}, n_base, X.sub_first.p, X.scan_tmp, d_tot + 0, s);

  phase_done("per-read kernels, sub-columns");
  // ---- first synchronisation: the geometry statistics and the number of sub-column slots
  uint64_t geo_stats[384], h_tot[16];
  uint32_t h_stats[XS_WORDS];
  CUDA_OK(cudaMemcpyAsync(geo_stats, X.geo_stats.p, sizeof geo_stats, cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaMemcpyAsync(h_tot, X.totals.p, sizeof h_tot, cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaMemcpyAsync(h_stats, X.stats.p, sizeof h_stats, cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
  if (h_stats[XS_ERR]) throw std::runtime_error(expand_error_text(h_stats[XS_ERR]));
  st.max_read_set_seen = h_stats[XS_MAX_SET];
  st.n_ins = h_tot[0];
  const uint64_t n_slots64 = st.n_base + st.n_ins;
  if (n_slots64 >= ROUND_NO_SLOT) throw std::runtime_error("more than 2^32 - 2 slots in one staged stream");
  const uint32_t n_slots = (uint32_t)n_slots64, n_ins = (uint32_t)st.n_ins;
  if (cfg.want_score) st.geo = choose_geometry(geo_stats, geo_stats + 256, cfg, st.max_read_set_seen);
  a.geo = st.geo;
  st.hist_bytes = (cfg.use_base_repeat || cfg.use_read_pos || st.max_read_set_seen > 7) ? 8 : 4;
  a.hist_bytes = st.hist_bytes;

  out.slot_ref.ensure((size_t)n_slots + 1);
  X.ins_parent.ensure((size_t)n_ins + 1); X.ins_count.ensure((size_t)n_ins + 1);
  a.slot_ref = out.slot_ref.p;
  slot_ref_kernel<<<blocks_for(n_base), 256, 0, s>>>(a, out.slot_group.p);
  fill_ins_kernel<<<blocks_for(n_base), 256, 0, s>>>(X.sub_k.p, X.sub_first.p, n_base, X.ins_parent.p, X.ins_count.p, out.slot_ref.p);
  launched(2);
  st.ins_parent.resize(n_ins); st.ins_count.resize(n_ins);
  if (n_ins) {
    CUDA_OK(cudaMemcpyAsync(st.ins_parent.data(), X.ins_parent.p, (size_t)n_ins * 8, cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaMemcpyAsync(st.ins_count.data(), X.ins_count.p, (size_t)n_ins * 4, cudaMemcpyDeviceToHost, s));
  }

  phase_done("slot_ref, ins arrays");
  // ---- count pass
  out.score_cnt.ensure((size_t)n_slots + 1);
  X.red_cnt.ensure((size_t)n_slots + 1); X.side_cnt.ensure((size_t)n_slots + 1); X.side_red_cnt.ensure((size_t)n_slots + 1);
  X.hist_cnt.ensure((size_t)n_base + 1); X.col_red.ensure((size_t)n_base + 1); X.col_qstart.ensure((size_t)n_base + 1);
  for (uint32_t* p : {out.score_cnt.p, X.red_cnt.p, X.side_cnt.p, X.side_red_cnt.p}) CUDA_OK(cudaMemsetAsync(p, 0, ((size_t)n_slots + 1) * 4, s));
  CUDA_OK(cudaMemsetAsync(X.hist_cnt.p, 0, ((size_t)n_base + 1) * 4, s));
  CUDA_OK(cudaMemsetAsync(X.col_red.p, 0, (size_t)n_base + 1, s));
  CUDA_OK(cudaMemsetAsync(X.col_qstart.p, 0, (size_t)n_base + 1, s));
  a.score_cnt = out.score_cnt.p; a.red_cnt = X.red_cnt.p; a.side_cnt = X.side_cnt.p; a.side_red_cnt = X.side_red_cnt.p;
  a.hist_cnt = X.hist_cnt.p; a.col_red = X.col_red.p; a.col_qstart = X.col_qstart.p;
  X.walk_args = a; X.have_walk_args = true;   // (what a later walk over the same reads needs: coverage_columns_on_device)
  if (tiles) { tile_kernel<false><<<blocks_for((uint64_t)tiles * 32), 256, 0, s>>>(a); launched(); }
  if (cfg.preprocess_stage && cfg.want_hist) {
    X.read_starts.ensure(n_targets * 2 + 2);
    CUDA_OK(cudaMemsetAsync(X.read_starts.p, 0, (n_targets * 2 + 2) * 8, s));
    read_starts_kernel<<<blocks_for(n_base), 256, 0, s>>>(a, reinterpret_cast<unsigned long long*>(X.read_starts.p));
    launched();
  }

  phase_done("count pass");
  // ---- offsets: side list, histogram records, rounds
  out.side_off.ensure((size_t)n_slots + 2); out.hist_off.ensure((size_t)n_base + 2); out.score_off.ensure((size_t)n_slots + 2);
  exclusive_scan<FSideEven, uint32_t, true>(FSideEven{X.side_cnt.p}, n_slots, out.side_off.p, X.scan_tmp, d_tot + 1, s);
  exclusive_scan<FU32, uint64_t, true>(FU32{X.hist_cnt.p}, n_base, out.hist_off.p, X.scan_tmp, d_tot + 2, s);
  hist_off_flag_kernel<<<blocks_for(n_base), 256, 0, s>>>(out.hist_off.p, X.col_red.p, n_base);
  max_u32_kernel<<<148, 256, 0, s>>>(X.hist_cnt.p, n_base, X.stats.p + XS_MAX_HIST_DEPTH);
  exclusive_scan<FU32, uint64_t, false>(FU32{out.score_cnt.p}, n_slots, (uint64_t*)nullptr, X.scan_tmp, d_tot + 3, s);  // records in all
  launched(2);
  const uint32_t n_oblocks = (n_slots + ROUND_BLOCK - 1) / ROUND_BLOCK;
  X.block_entries.ensure((size_t)n_oblocks + 1); X.block_base.ensure((size_t)n_oblocks + 2);
  if (n_oblocks) { order_count_kernel<<<n_oblocks, 256, 0, s>>>(out.slot_ref.p, n_slots, X.block_entries.p); launched(); }
  exclusive_scan<FU32, uint32_t, true>(FU32{X.block_entries.p}, n_oblocks, X.block_base.p, X.scan_tmp, d_tot + 4, s);
  // the padded order has at most n_slots + 5 * 31 entries per block
  const uint64_t max_entries = (uint64_t)n_slots + (uint64_t)n_oblocks * 160 + 32;
  out.round_slot.ensure(max_entries + 16); X.round_vecs.ensure(max_entries / 32 + 2); out.round_off.ensure(max_entries / 32 + 3);
  out.round_side.ensure(max_entries * 2 + 16);
  if (n_oblocks) { order_sort_kernel<<<n_oblocks, 1024, 0, s>>>(out.slot_ref.p, out.score_cnt.p, n_slots, X.block_base.p, out.round_slot.p, X.round_vecs.p, X.stats.p); launched(); }

  // ---- second synchronisation: the number of rounds (sizes the scan over them)
  CUDA_OK(cudaMemcpyAsync(h_tot, X.totals.p, sizeof h_tot, cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
  const uint64_t n_entries = h_tot[4];
  st.n_rounds = n_entries / 32;
  st.n_side = h_tot[1]; st.n_hist = h_tot[2]; st.n_score = h_tot[3];
  if (st.n_side >= (1ull << 29)) throw std::runtime_error("more than 2^29 side-list entries in one staged stream");
  exclusive_scan<FVecs256, uint64_t, true>(FVecs256{X.round_vecs.p}, st.n_rounds, out.round_off.p, X.scan_tmp, d_tot + 5, s);
  if (n_entries) {
    assign_kernel<<<blocks_for(n_entries), 256, 0, s>>>(out.round_slot.p, n_entries, out.round_off.p, out.side_off.p, X.side_cnt.p, out.slot_ref.p,
                                                        cfg.want_score ? 1u : 0u, out.score_off.p, out.round_side.p);
    launched();
  }
  CUDA_OK(cudaMemcpyAsync(h_tot, X.totals.p, sizeof h_tot, cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
  st.n_score_padded = h_tot[5];
  CUDA_OK(cudaMemcpyAsync(out.score_off.p + n_slots, X.totals.p + 5, 8, cudaMemcpyDeviceToDevice, s));

  phase_done("offsets and rounds");
  // ---- fill pass
  out.score_rec.ensure(st.n_score_padded + 64);
  out.side_rec.ensure(st.n_side * st.geo.side_stride + 4);
  const bool compact = cfg.want_hist && cfg.compact_hist && st.hist_bytes == 4;
  DevBuf<uint8_t>& hist_pos = compact ? X.hist_pos : out.hist_rec;
  hist_pos.ensure(st.n_hist * st.hist_bytes + 64);
  X.sub_cur.ensure((size_t)n_ins * 4 + 4);
  if (st.n_score_padded) { fill_words_kernel<<<148 * 8, 256, 0, s>>>(out.score_rec.p, st.n_score_padded, st.geo.pad_word()); launched(); }
  if (st.n_side) CUDA_OK(cudaMemsetAsync(out.side_rec.p, 0xFF, st.n_side * st.geo.side_stride * 4, s));
  if (n_ins) { sub_cursor_kernel<<<blocks_for(n_ins), 256, 0, s>>>(X.red_cnt.p, X.side_red_cnt.p, n_base, n_ins, X.sub_cur.p); launched(); }
  a.score_off = out.score_off.p; a.side_off = out.side_off.p; a.hist_off = out.hist_off.p;
  a.score_rec = out.score_rec.p; a.side_rec = out.side_rec.p; a.hist_rec = hist_pos.p; a.sub_cur = X.sub_cur.p;
  if (tiles) { tile_kernel<true><<<blocks_for((uint64_t)tiles * 32), 256, 0, s>>>(a); launched(); }

  phase_done("fill pass");
  // ---- compact histogram stream
  st.hist_compact = false;
  if (compact) {
    const uint32_t nb = (uint32_t)((st.n_hist + SCAN_BLOCK - 1) / SCAN_BLOCK);
    X.scan_tmp32.ensure((size_t)nb + 1);
    DevBuf<uint64_t>& base16 = X.ins_mask;   // free again: the sub-column masks are consumed
    base16.ensure((size_t)nb + 2);
    if (nb) { hist16_count_kernel<<<nb, SCAN_TPB, 0, s>>>(reinterpret_cast<const uint32_t*>(hist_pos.p), st.n_hist, X.scan_tmp32.p); launched(); }
    exclusive_scan<FU32, uint64_t, true>(FU32{X.scan_tmp32.p}, nb, base16.p, X.scan_tmp, d_tot + 6, s);
    CUDA_OK(cudaMemcpyAsync(h_tot, X.totals.p, sizeof h_tot, cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
    st.n_hist16 = h_tot[6]; st.n_hist_exc = st.n_hist - st.n_hist16;
    out.hist_exc_at = (st.n_hist16 * 2 + 31) & ~(size_t)15;
    out.hist_rec.ensure(out.hist_exc_at + st.n_hist_exc * 4 + 64);
    if (nb) {
      hist16_scatter_kernel<<<nb, SCAN_TPB, 0, s>>>(reinterpret_cast<const uint32_t*>(hist_pos.p), st.n_hist, base16.p,
                                                    reinterpret_cast<uint16_t*>(out.hist_rec.p), reinterpret_cast<uint32_t*>(out.hist_rec.p + out.hist_exc_at));
      launched();
    }
    st.hist_compact = true;
  }

  phase_done("compact histogram");
  // ---- statistics and errors
  CUDA_OK(cudaMemcpyAsync(h_stats, X.stats.p, sizeof h_stats, cudaMemcpyDeviceToHost, s));
  std::vector<uint64_t> starts;
  if (cfg.preprocess_stage && cfg.want_hist) {
    starts.resize(n_targets * 2);
    if (n_targets) CUDA_OK(cudaMemcpyAsync(starts.data(), X.read_starts.p, n_targets * 16, cudaMemcpyDeviceToHost, s));
  }
  CUDA_OK(cudaStreamSynchronize(s));
  if (h_stats[XS_ERR]) throw std::runtime_error(expand_error_text(h_stats[XS_ERR]));
  st.read_start_counts = starts;
  st.max_qual_seen = h_stats[XS_MAX_Q]; st.max_hist_qual = h_stats[XS_MAX_HQ]; st.max_hist_rpos = h_stats[XS_MAX_RP];
  st.max_score_rpos = h_stats[XS_MAX_SRP]; st.max_hist_depth = cfg.want_hist ? h_stats[XS_MAX_HIST_DEPTH] : 0;
  for (int w = 0; w < 8; ++w) st.mapq_seen[w] = h_stats[XS_MAPQ_SEEN + w];
  if (cfg.want_score) st.mapq_seen[st.geo.hot_mapq >> 5] |= 1u << (st.geo.hot_mapq & 31);
  st.bytes_uploaded = reads.bytes + refbytes.size() + n_visit * sizeof(ExpandSeg);
}

void gather_flagged_records(const StreamDev& ds, const PileupStream& st, const uint32_t* d_slots, uint32_t n, FlaggedRecordsHost& out,
                            cudaStream_t s, uint64_t* d2h_bytes) {
  out.word_off.assign((size_t)n + 1, 0); out.side_off.assign((size_t)n + 1, 0); out.words.clear(); out.side.clear(); out.ref.assign(n, 5);
  if (!n) return;
  DevBuf<SlotInfo> d_info;
  DevBuf<uint64_t> d_off;
  DevBuf<uint32_t> d_words, d_side;
  d_info.ensure(n);
  slot_info_kernel<<<blocks_for(n), 256, 0, s>>>(d_slots, n, ds.score_off.p, ds.score_cnt.p, ds.side_off.p, ds.slot_ref.p, d_info.p);
  std::vector<SlotInfo> info(n);
  CUDA_OK(cudaMemcpyAsync(info.data(), d_info.p, (size_t)n * sizeof(SlotInfo), cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
  for (uint32_t i = 0; i < n; ++i) {
    out.word_off[i + 1] = out.word_off[i] + info[i].cnt;
    out.side_off[i + 1] = out.side_off[i] + (info[i].side1 - info[i].side0);
    out.ref[i] = (uint8_t)info[i].ref;
  }
  const uint32_t ss = st.geo.side_stride;
  out.words.resize(out.word_off[n]); out.side.resize(out.side_off[n] * ss);
  d_off.ensure(2 * ((size_t)n + 1)); d_words.ensure(out.words.size() + 1); d_side.ensure(out.side.size() + 1);
  CUDA_OK(cudaMemcpyAsync(d_off.p, out.word_off.data(), ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, s));
  CUDA_OK(cudaMemcpyAsync(d_off.p + n + 1, out.side_off.data(), ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, s));
  gather_records_kernel<<<std::min<uint32_t>(blocks_for((uint64_t)n * 32), 148 * 8), 256, 0, s>>>(d_info.p, n, d_off.p, d_off.p + n + 1, ds.score_rec.p, ds.side_rec.p, ss,
                                                                                               d_words.p, d_side.p);
  launched(2);
  if (!out.words.empty()) CUDA_OK(cudaMemcpyAsync(out.words.data(), d_words.p, out.words.size() * 4, cudaMemcpyDeviceToHost, s));
  if (!out.side.empty()) CUDA_OK(cudaMemcpyAsync(out.side.data(), d_side.p, out.side.size() * 4, cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
  if (d2h_bytes) *d2h_bytes += (uint64_t)n * sizeof(SlotInfo) + out.words.size() * 4 + out.side.size() * 4;
  d_info.release(); d_off.release(); d_words.release(); d_side.release();
}

void coverage_columns_on_device(const ExpandScratch& X, uint64_t n_base, DevBuf<CoverageColumn>& out, uint32_t group, bool include_deleted, cudaStream_t s) {
  if (!X.have_walk_args) throw std::runtime_error("the coverage table needs reads staged on the device (brq_stage_options.staging = 0 or 2)");
  out.ensure(n_base + 1);
  const ExpandArgs& a = X.walk_args;
  if (a.n_tiles) { coverage_tile_kernel<<<blocks_for((uint64_t)a.n_tiles * 32), 256, 0, s>>>(a, out.p, group, include_deleted); launched(); }
}

}  // namespace brq
