// The library's radix sorts and scans (cub::DeviceRadixSort / DeviceScan: CUDA toolkit library code),
// instantiated in ONE translation unit. voxel.cu (map ingestion) and expand.cu (edge grouping of the
// device BFS) call these wrappers: instantiating the same CUB kernels in two objects of one shared
// library registered them twice and the large-input (onesweep) path then died with SIGFPE.
#include <cub/cub.cuh>

#include "common.cuh"

namespace trgb {

namespace {
template <class K, class V>
int sort_pairs(const K* kin, K* kout, const V* vin, V* vout, int n, int end_bit, cudaStream_t st) {
  if (n <= 0) return TRGB_OK;
  size_t bytes = 0;
  TRGB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, kin, kout, vin, vout, n, 0, end_bit, st));
  void* tmp = nullptr;
  TRGB_CUDA(cudaMallocAsync(&tmp, std::max<size_t>(bytes, 16), st));
  cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, bytes, kin, kout, vin, vout, n, 0, end_bit, st);
  cudaFreeAsync(tmp, st);
  if (e != cudaSuccess) return cuda_fail(e, "cub::DeviceRadixSort::SortPairs", __FILE__, __LINE__);
  return TRGB_OK;
}
}  // namespace

int sort_pairs_u32_u32(const uint32_t* kin, uint32_t* kout, const uint32_t* vin, uint32_t* vout, int n, int end_bit,
                       cudaStream_t st) {
  return sort_pairs(kin, kout, vin, vout, n, end_bit, st);
}
int sort_pairs_u32_u64(const uint32_t* kin, uint32_t* kout, const unsigned long long* vin, unsigned long long* vout, int n,
                       int end_bit, cudaStream_t st) {
  return sort_pairs(kin, kout, vin, vout, n, end_bit, st);
}
int sort_pairs_u64_u32(const unsigned long long* kin, unsigned long long* kout, const uint32_t* vin, uint32_t* vout, int n,
                       int end_bit, cudaStream_t st) {
  return sort_pairs(kin, kout, vin, vout, n, end_bit, st);
}
int exclusive_sum_u32(const uint32_t* in, uint32_t* out, int n, cudaStream_t st) {
  if (n <= 0) return TRGB_OK;
  size_t bytes = 0;
  TRGB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, n, st));
  void* tmp = nullptr;
  TRGB_CUDA(cudaMallocAsync(&tmp, std::max<size_t>(bytes, 16), st));
  cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, n, st);
  cudaFreeAsync(tmp, st);
  if (e != cudaSuccess) return cuda_fail(e, "cub::DeviceScan::ExclusiveSum", __FILE__, __LINE__);
  return TRGB_OK;
}

}  // namespace trgb
