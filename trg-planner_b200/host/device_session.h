// Pinned-host / device buffer pairs used by the wavefront scheduler. Memory management only
// (CUDA runtime); every kernel is reached through the C ABI in include/trgb_kernels.h.
#pragma once
#include <cuda_runtime_api.h>

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>

namespace trg_b200 {

inline void cuda_check(cudaError_t e, const char* what) {
  if (e != cudaSuccess)
    throw std::runtime_error(std::string("trg_b200: CUDA failure in ") + what + ": " + cudaGetErrorString(e));
}

template <class T>
struct Mirror {
  T* h = nullptr;
  T* d = nullptr;
  size_t cap = 0;
  ~Mirror() { release(); }
  void release() {
    if (h) cudaFreeHost(h);
    if (d) cudaFree(d);
    h = nullptr; d = nullptr; cap = 0;
  }
  void ensure(size_t n) {
    if (n <= cap) return;
    size_t want = cap ? cap : 1024;
    while (want < n) want *= 2;
    release();
    cuda_check(cudaMallocHost(reinterpret_cast<void**>(&h), want * sizeof(T)), "cudaMallocHost");
    cuda_check(cudaMalloc(reinterpret_cast<void**>(&d), want * sizeof(T)), "cudaMalloc");
    cap = want;
  }
  void h2d(size_t n, cudaStream_t s) {
    if (n) cuda_check(cudaMemcpyAsync(d, h, n * sizeof(T), cudaMemcpyHostToDevice, s), "H2D");
  }
  void d2h(size_t n, cudaStream_t s) {
    if (n) cuda_check(cudaMemcpyAsync(h, d, n * sizeof(T), cudaMemcpyDeviceToHost, s), "D2H");
  }
  void zero_d(size_t n, cudaStream_t s) {
    if (n) cuda_check(cudaMemsetAsync(d, 0, n * sizeof(T), s), "memset");
  }
};

// All staging buffers of one TRG instance.
class DeviceSession {
 public:
  // sampling windows
  Mirror<float> node_xy;                 // 2 per node
  Mirror<int32_t> first_draw;            // per node, relative to the uploaded draw slice
  Mirror<float> draw_xy;                 // 2 per draw
  Mirror<unsigned long long> mask;       // words per node
  // sample / edge evaluation
  Mirror<float> p1;                      // 3 per item
  Mirror<float> p2;                      // 2 per item
  Mirror<float> z;                       // per item
  Mirror<uint8_t> stage;
  Mirror<float> weight;
  Mirror<float> dist;
  Mirror<uint8_t> tie;
  // deferred edge evaluations
  Mirror<float> dp1, dp2, dweight, ddist;
  Mirror<uint8_t> dstage;
  // generic queries
  Mirror<float> qxy;
  Mirror<uint8_t> qout;
  Mirror<int32_t> qcount;

  uint64_t bytes_h2d = 0, bytes_d2h = 0;
};

}  // namespace trg_b200
