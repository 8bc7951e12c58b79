// Pinned-host / device staging arenas used by the wavefront scheduler. Memory management only
// (CUDA runtime); every kernel is reached through the C ABI in include/trgb_kernels.h.
//
// One batch = one H2D copy of the whole input arena, a few kernel launches on sub-ranges of it,
// one D2H copy of the whole output arena, one stream synchronise.
#pragma once
#include <cuda_runtime_api.h>

#include "trgb_kernels.h"

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>

namespace trg_b200 {

inline void cuda_check(cudaError_t e, const char* what) {
  if (e != cudaSuccess)
    throw std::runtime_error(std::string("trg_b200: CUDA failure in ") + what + ": " + cudaGetErrorString(e));
}

// Bump allocator over a pinned host buffer and a device buffer of the same size.
class Arena {
 public:
  ~Arena() { release(); }
  void release() {
    if (h_) cudaFreeHost(h_);
    if (d_) cudaFree(d_);
    h_ = nullptr; d_ = nullptr; cap_ = 0; used_ = 0;
  }
  // Drop the contents and make sure `bytes` fit. Invalidates earlier offsets.
  void reset(size_t bytes) {
    used_ = 0;
    if (bytes <= cap_) return;
    size_t want = cap_ ? cap_ : (size_t)1 << 16;
    while (want < bytes) want *= 2;
    release();
    cuda_check(cudaMallocHost(reinterpret_cast<void**>(&h_), want), "cudaMallocHost(arena)");
    cuda_check(cudaMalloc(reinterpret_cast<void**>(&d_), want), "cudaMalloc(arena)");
    cap_ = want;
  }
  static size_t padded(size_t bytes) { return (bytes + 255) & ~(size_t)255; }
  size_t take(size_t bytes) {
    const size_t off = used_;
    used_ += padded(bytes);
    if (used_ > cap_) throw std::logic_error("trg_b200: arena overflow (reset() sized too small)");
    return off;
  }
  template <class T> T* h(size_t off) { return reinterpret_cast<T*>(h_ + off); }
  template <class T> T* d(size_t off) { return reinterpret_cast<T*>(d_ + off); }
  size_t used() const { return used_; }
  void h2d(cudaStream_t s) {
    if (used_) cuda_check(cudaMemcpyAsync(d_, h_, used_, cudaMemcpyHostToDevice, s), "H2D(arena)");
  }
  void d2h(cudaStream_t s) {
    if (used_) cuda_check(cudaMemcpyAsync(h_, d_, used_, cudaMemcpyDeviceToHost, s), "D2H(arena)");
  }
  void zero_d(size_t off, size_t bytes, cudaStream_t s) {
    if (bytes) cuda_check(cudaMemsetAsync(d_ + off, 0, bytes, s), "memset(arena)");
  }

 private:
  uint8_t* h_ = nullptr;
  uint8_t* d_ = nullptr;
  size_t cap_ = 0, used_ = 0;
};

// Device-resident copy of the sampling-draw offsets (expand_dist*cosf(angle), expand_dist*sinf(angle))
// for stream positions [base, base+count): uploaded once per generated block, indexed by the
// sampling-window kernel.
class DrawBuffer {
 public:
  ~DrawBuffer() { if (d_) { cudaDeviceSynchronize(); cudaFreeAsync(d_, 0); cudaStreamSynchronize(0); } }
  void clear() { count_ = 0; base_ = 0; }
  size_t base() const { return base_; }
  size_t end() const { return base_ + count_; }
  const float* dev() const { return d_; }
  void set_base(size_t b) { base_ = b; count_ = 0; }
  // append n (x,y) pairs that follow the current end
  void append(const float* xy, size_t n, cudaStream_t s) {
    if (count_ + n > cap_) {
      size_t want = cap_ ? cap_ : (size_t)1 << 21;
      while (want < count_ + n) want *= 2;
      // stream-ordered pool: growing a 100+ MB buffer must not stall on cudaMalloc / cudaFree
      float* nd = nullptr;
      cuda_check(cudaMallocAsync(reinterpret_cast<void**>(&nd), want * 2 * sizeof(float), s), "cudaMallocAsync(draws)");
      if (count_) cuda_check(cudaMemcpyAsync(nd, d_, count_ * 2 * sizeof(float), cudaMemcpyDeviceToDevice, s), "D2D(draws)");
      if (d_) cudaFreeAsync(d_, s);
      cuda_check(cudaStreamSynchronize(s), "sync(draws)");
      d_ = nd;
      cap_ = want;
    }
    // source is pageable host memory: the runtime stages it before returning
    cuda_check(cudaMemcpyAsync(d_ + count_ * 2, xy, n * 2 * sizeof(float), cudaMemcpyHostToDevice, s), "H2D(draws)");
    cuda_check(cudaStreamSynchronize(s), "sync(draws)");  // visible to every stream launched after this returns
    count_ += n;
  }

 private:
  float* d_ = nullptr;
  size_t cap_ = 0, count_ = 0, base_ = 0;
};

// All staging state of one TRG instance.
class DeviceSession {
 public:
  ~DeviceSession() {
    if (copy_) cudaStreamDestroy(copy_);
    if (nodes) trgb_nodes_destroy(nodes);
  }
  // device grid over the nodes of the graph being expanded (K5) and how much of node_seq it holds
  trgb_nodes* nodes = nullptr;
  const void* nodes_owner = nullptr;
  float nodes_box[5] = {0, 0, 0, 0, 0};
  size_t nodes_uploaded = 0;
  Arena in, out;    // batch staging (windows, speculative evaluation)
  Arena in2, out2;  // mid-commit flushes: must not disturb the results `out` still holds
  DrawBuffer draws;
  uint64_t batches = 0;
  cudaStream_t copyStream() {
    if (!copy_) cuda_check(cudaStreamCreateWithFlags(&copy_, cudaStreamNonBlocking), "cudaStreamCreate");
    return copy_;
  }

 private:
  cudaStream_t copy_ = nullptr;
};

}  // namespace trg_b200
