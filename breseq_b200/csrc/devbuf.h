// Device buffers owned by a context: grown on demand, never shrunk, released with the context.
#pragma once
#include <cuda_runtime.h>

#include <stdexcept>
#include <string>

#define CUDA_OK(call)                                                                                  \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) throw std::runtime_error(std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

namespace brq {

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  void ensure(size_t want) {
    if (want <= n && p) return;
    release();
    CUDA_OK(cudaMalloc((void**)&p, (want ? want : 1) * sizeof(T)));
    n = want;
  }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

}  // namespace brq
