// Micro-benchmark: issue rate of shared-memory reductions (ATOMS.ADD without return) against LDS+STS byte increments,
// one CTA per SM, 24 warps, every lane on its own bank.  Prints cycles per warp instruction per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k_atoms(uint32_t* out, int iters, int words) {
  extern __shared__ uint32_t sm[];
  for (int i = threadIdx.x; i < 24 * 1024; i += blockDim.x) sm[i] = 0;
  __syncthreads();
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + warp * 4096 + lane * 4;
  uint32_t x = threadIdx.x * 2654435761u + blockIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      x = x * 1664525u + 1013904223u;
      const uint32_t w = (x >> 16) % (uint32_t)words;
      asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(base + w * 128), "r"(1u << ((x >> 8) & 24)) : "memory");
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = (uint32_t)(t1 - t0);
  if (sm[threadIdx.x] == 0xdeadbeef) out[0] = 1;
}
__global__ void k_ldsts(uint32_t* out, int iters, int words) {
  extern __shared__ uint32_t sm[];
  for (int i = threadIdx.x; i < 24 * 1024; i += blockDim.x) sm[i] = 0;
  __syncthreads();
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + warp * 4096 + lane * 4;
  uint32_t x = threadIdx.x * 2654435761u + blockIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      x = x * 1664525u + 1013904223u;
      const uint32_t a = base + ((x >> 16) % (uint32_t)words) * 128 + ((x >> 11) & 3);
      uint32_t v;
      asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
      asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v + 1) : "memory");
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = (uint32_t)(t1 - t0);
  if (sm[threadIdx.x] == 0xdeadbeef) out[0] = 1;
}
int main() {
  uint32_t* d; cudaMalloc(&d, 148 * 4);
  const int iters = 2000, words = 26;
  cudaFuncSetAttribute(k_atoms, cudaFuncAttributeMaxDynamicSharedMemorySize, 24 * 4096);
  cudaFuncSetAttribute(k_ldsts, cudaFuncAttributeMaxDynamicSharedMemorySize, 24 * 4096);
  for (int warps : {4, 8, 16, 24}) {
    uint32_t h[148];
    k_atoms<<<148, warps * 32, 24 * 4096>>>(d, iters, words); cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    double a = (double)h[5] / ((double)iters * 8 * warps);
    k_ldsts<<<148, warps * 32, 24 * 4096>>>(d, iters, words); cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    double b = (double)h[5] / ((double)iters * 8 * warps);
    printf("warps %2d: ATOMS.ADD %.2f cycles per warp instruction per SM; LDS.U8+STS.U8 pair %.2f (%s)\n", warps, a, b, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
