// Pass 1's collective, fused into pass 1's stream work: the ranks of a run sharded by reference range sum their covariate and
// coverage histograms through each other's memory over NVLink, with no collective library call and no host in the loop.
//
// Every rank owns an INBOX in its own HBM (two copies of the histogram's shape, used in turn, and two arrival counters), which
// its peers have mapped (CUDA IPC).  After its histogram kernels a rank runs ONE kernel:
//   push     every non-zero bin of the local histogram is added to the inbox of every peer (system-scope reductions over
//            NVLink: a few thousand per peer), a system-wide fence, and the last CTA to finish bumps every peer's arrival counter;
//   combine  that same CTA waits until its own counter says all peers have pushed, adds its inbox to the local histogram
//            (which from then on holds the run's totals: table derivation, the files and the coverage fit read it as before),
//            and clears the inbox copy and the counter for their next use.
// Two inbox copies make the clearing safe without another round of messages: a peer can only push step k + 2 into copy k % 2
// after it has combined step k + 1, which needs this rank's push of step k + 1, which this rank's stream orders after its own
// combine (and clear) of step k.
#include "kernels.h"

#include <cuda_runtime.h>

namespace brq {

void note_launches(int n);

namespace {

__device__ __forceinline__ uint32_t load_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256) hist_exchange_kernel(unsigned long long* __restrict__ local, uint64_t n, HistPeers P, uint32_t copy,
                                                            uint32_t* __restrict__ done, uint32_t* __restrict__ err, long long timeout_cycles) {
  __shared__ bool last;
  // ---- push
  for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n; b += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long v = local[b];
    if (!v) continue;
    for (uint32_t r = 0; r < P.world; ++r)
      if (r != P.rank) atomicAdd_system(P.inbox[r] + (size_t)copy * P.capacity + b, v);
  }
  __threadfence_system();   // this thread's reductions are performed before the counter below can be seen to move
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence_system();
  if (threadIdx.x < P.world && threadIdx.x != P.rank) atomicAdd_system(P.arrived[threadIdx.x] + copy, 1u);
  // ---- combine
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    bool ok = true;
    while (load_acquire_sys(P.arrived[P.rank] + copy) < P.world - 1) {
      if (clock64() - t0 > timeout_cycles) { ok = false; break; }
      __nanosleep(200);
    }
    if (!ok) atomicOr(err, BRQ_ERR_PEER_TIMEOUT);
    *done = 0;
  }
  __syncthreads();
  unsigned long long* mine = P.inbox[P.rank] + (size_t)copy * P.capacity;
  for (uint64_t b = threadIdx.x; b < n; b += blockDim.x) {
    const unsigned long long v = __ldcv(mine + b);   // written by the peers: not through a stale cache line
    if (v) { local[b] += v; mine[b] = 0; }
  }
  __syncthreads();
  if (threadIdx.x == 0) P.arrived[P.rank][copy] = 0;
}

}  // namespace

void launch_hist_exchange(unsigned long long* local, uint64_t n, const HistPeers& peers, uint32_t copy, uint32_t* done, uint32_t* err,
                          double timeout_seconds, cudaStream_t s) {
  if (peers.world < 2 || !n) return;
  const long long cycles = (long long)(timeout_seconds * 1.9e9);
  hist_exchange_kernel<<<8, 256, 0, s>>>(local, n, peers, copy, done, err, cycles);
  note_launches(1);
}

}  // namespace brq
