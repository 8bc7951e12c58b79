// K8 — voxel-grid centroid filter on the device: the optional down-sampling step of the map
// ingestion that precedes the hot path, TRGPlanner::loadPrebuiltMap -> pcl::VoxelGrid
// (src/planner/trg_planner.cpp:90-94; config/indoor.yaml:8-9 turns it on with a 0.2 m leaf).
// PCL is third-party and absent from this image; the algorithm below restates
// pcl/filters/impl/voxel_grid.hpp (PCL 1.10-1.12, recalled — PARITY UNPINNED at that boundary):
//   min_b = floor(min_p * inv_leaf), max_b = floor(max_p * inv_leaf), div_b = max_b - min_b + 1
//   idx(p) = (floor(p.x*inv) - min_b.x) + (floor(p.y*inv) - min_b.y)*div.x + (floor(p.z*inv) - min_b.z)*div.x*div.y
//   points sorted by idx, one output point per occupied voxel = centroid, output in ascending idx.
// PCL sums each voxel in float in the (unspecified, std::sort is not stable) order of its sorted
// index vector; here the sums are taken in double, which is within one float ulp of any such order.
// The sort is cub::DeviceRadixSort (CUDA toolkit library code: this is ingestion, not the hot path).
#include <cfloat>
#include <climits>
#include <cmath>

#include "common.cuh"

namespace trgb {

__global__ void __launch_bounds__(256) k_vox_minmax(const float* __restrict__ pts, int64_t n, int stride,
                                                    float* __restrict__ mm /* minx miny minz maxx maxy maxz */) {
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float* p = pts + i * stride;
    const float v[3] = {__ldg(p), __ldg(p + 1), __ldg(p + 2)};
    if (!isfinite(v[0]) || !isfinite(v[1]) || !isfinite(v[2])) continue;  // PCL skips non-finite points
#pragma unroll
    for (int k = 0; k < 3; ++k) { mn[k] = fminf(mn[k], v[k]); mx[k] = fmaxf(mx[k], v[k]); }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      mn[k] = fminf(mn[k], __shfl_xor_sync(FULL, mn[k], d));
      mx[k] = fmaxf(mx[k], __shfl_xor_sync(FULL, mx[k], d));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      // ordered-int trick for float atomics of any sign
      if (mn[k] >= 0.f) atomicMin(reinterpret_cast<int*>(mm + k), __float_as_int(mn[k]));
      else atomicMax(reinterpret_cast<unsigned*>(mm + k), __float_as_uint(mn[k]));
      if (mx[k] >= 0.f) atomicMax(reinterpret_cast<int*>(mm + 3 + k), __float_as_int(mx[k]));
      else atomicMin(reinterpret_cast<unsigned*>(mm + 3 + k), __float_as_uint(mx[k]));
    }
  }
}

__global__ void __launch_bounds__(256) k_vox_index(const float* __restrict__ pts, int64_t n, int stride, float inv,
                                                   int mbx, int mby, int mbz, int dx, int dxy,
                                                   uint32_t* __restrict__ key, uint32_t* __restrict__ val) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float* p = pts + i * stride;
    const float x = __ldg(p), y = __ldg(p + 1), z = __ldg(p + 2);
    uint32_t k = 0xffffffffu;  // non-finite points sort to the end and are dropped
    if (isfinite(x) && isfinite(y) && isfinite(z)) {
      const int ix = (int)(floorf(__fmul_rn(x, inv)) - (float)mbx);
      const int iy = (int)(floorf(__fmul_rn(y, inv)) - (float)mby);
      const int iz = (int)(floorf(__fmul_rn(z, inv)) - (float)mbz);
      k = (uint32_t)(ix + iy * dx + iz * dxy);
    }
    key[i] = k;
    val[i] = (uint32_t)i;
  }
}

// head[i] = 1 where a new voxel starts in the sorted key sequence
__global__ void __launch_bounds__(256) k_vox_heads(const uint32_t* __restrict__ key, int64_t n, uint32_t* __restrict__ head) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    head[i] = (key[i] != 0xffffffffu && (i == 0 || key[i] != key[i - 1])) ? 1u : 0u;
}

// one thread per occupied voxel: walk its run of the sorted sequence, centroid in double
__global__ void __launch_bounds__(256) k_vox_centroid(const float* __restrict__ pts, int stride,
                                                      const uint32_t* __restrict__ key, const uint32_t* __restrict__ val,
                                                      const uint32_t* __restrict__ head, const uint32_t* __restrict__ slot,
                                                      int64_t n, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (!head[i]) continue;
    const uint32_t k = key[i];
    double sx = 0, sy = 0, sz = 0;
    int64_t j = i;
    for (; j < n && key[j] == k; ++j) {
      const float* p = pts + (int64_t)val[j] * stride;
      sx += (double)__ldg(p); sy += (double)__ldg(p + 1); sz += (double)__ldg(p + 2);
    }
    const double cnt = (double)(j - i);
    float* o = out + 3 * (int64_t)slot[i];
    o[0] = (float)(sx / cnt); o[1] = (float)(sy / cnt); o[2] = (float)(sz / cnt);
  }
}

}  // namespace trgb

using namespace trgb;

// Device in, device out. *d_out (3 floats per point) is allocated from the stream-ordered pool and
// must be released with trgb_device_free. Returns TRGB_E_STATE (and passes the input through as a
// copy) when the voxel grid would overflow 32-bit indices — PCL's "leaf size is too small" case.
extern "C" int trgb_voxel_filter_dev(const float* d_pts, int64_t n, int stride_floats, float leaf, float** d_out,
                                     int64_t* n_out) {
  TRGB_ARG(d_pts && d_out && n_out && n > 0, "null pointer / empty cloud");
  TRGB_ARG(stride_floats >= 3 && leaf > 0.f, "bad stride / leaf");
  TRGB_ARG(n < (int64_t)0x7fffffff, "more than 2^31-1 points");
  tune_mempool_once();
  cudaStream_t st = 0;
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 16);
  float* d_mm = nullptr;
  TRGB_CUDA(cudaMallocAsync((void**)&d_mm, 6 * sizeof(float), st));
  const float init[6] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
  TRGB_CUDA(cudaMemcpyAsync(d_mm, init, sizeof(init), cudaMemcpyHostToDevice, st));
  {
    ProfScope ps("k_vox_minmax", st, (double)n);
    k_vox_minmax<<<grid, 256, 0, st>>>(d_pts, n, stride_floats, d_mm);
  }
  float mm[6];
  TRGB_CUDA(cudaMemcpyAsync(mm, d_mm, sizeof(mm), cudaMemcpyDeviceToHost, st));
  TRGB_CUDA(cudaStreamSynchronize(st));
  cudaFreeAsync(d_mm, st);
  const float inv = 1.0f / leaf;
  int mb[3], xb[3];
  int64_t div[3];
  for (int k = 0; k < 3; ++k) {
    mb[k] = (int)std::floor(mm[k] * inv);
    xb[k] = (int)std::floor(mm[3 + k] * inv);
    div[k] = (int64_t)xb[k] - mb[k] + 1;
  }
  if (!(mm[0] <= mm[3]) || div[0] * div[1] * div[2] > (int64_t)INT_MAX) {
    set_error("voxel_filter: leaf size too small for the cloud extent (index overflow); cloud passed through");
    float* out = nullptr;
    TRGB_CUDA(cudaMallocAsync((void**)&out, (size_t)n * 3 * sizeof(float), st));
    TRGB_CUDA(cudaMemcpy2DAsync(out, 3 * sizeof(float), d_pts, (size_t)stride_floats * sizeof(float), 3 * sizeof(float),
                                (size_t)n, cudaMemcpyDeviceToDevice, st));
    TRGB_CUDA(cudaStreamSynchronize(st));
    *d_out = out;
    *n_out = n;
    return TRGB_E_STATE;
  }
  uint32_t *key = nullptr, *val = nullptr, *key2 = nullptr, *val2 = nullptr, *head = nullptr, *slot = nullptr;
  TRGB_CUDA(cudaMallocAsync((void**)&key, n * sizeof(uint32_t), st));
  TRGB_CUDA(cudaMallocAsync((void**)&val, n * sizeof(uint32_t), st));
  TRGB_CUDA(cudaMallocAsync((void**)&key2, n * sizeof(uint32_t), st));
  TRGB_CUDA(cudaMallocAsync((void**)&val2, n * sizeof(uint32_t), st));
  TRGB_CUDA(cudaMallocAsync((void**)&head, n * sizeof(uint32_t), st));
  TRGB_CUDA(cudaMallocAsync((void**)&slot, n * sizeof(uint32_t), st));
  {
    ProfScope ps("k_vox_index", st, (double)n);
    k_vox_index<<<grid, 256, 0, st>>>(d_pts, n, stride_floats, inv, mb[0], mb[1], mb[2], (int)div[0], (int)(div[0] * div[1]),
                                      key, val);
  }
  { int rc = sort_pairs_u32_u32(key, key2, val, val2, (int)n, 32, st); if (rc) return rc; }
  k_vox_heads<<<grid, 256, 0, st>>>(key2, n, head);
  { int rc = exclusive_sum_u32(head, slot, (int)n, st); if (rc) return rc; }
  uint32_t last_slot = 0, last_head = 0;
  TRGB_CUDA(cudaMemcpyAsync(&last_slot, slot + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  TRGB_CUDA(cudaMemcpyAsync(&last_head, head + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  TRGB_CUDA(cudaStreamSynchronize(st));
  const int64_t nv = (int64_t)last_slot + last_head;
  float* out = nullptr;
  TRGB_CUDA(cudaMallocAsync((void**)&out, (size_t)std::max<int64_t>(nv, 1) * 3 * sizeof(float), st));
  {
    ProfScope ps("k_vox_centroid", st, (double)n);
    k_vox_centroid<<<grid, 256, 0, st>>>(d_pts, stride_floats, key2, val2, head, slot, n, out);
  }
  TRGB_CUDA(cudaGetLastError());
  cudaFreeAsync(key, st); cudaFreeAsync(val, st); cudaFreeAsync(key2, st); cudaFreeAsync(val2, st);
  cudaFreeAsync(head, st); cudaFreeAsync(slot, st);
  TRGB_CUDA(cudaStreamSynchronize(st));
  *d_out = out;
  *n_out = nv;
  return TRGB_OK;
}

extern "C" void trgb_device_free(void* p) {
  if (p) {
    cudaFreeAsync(p, 0);
    cudaStreamSynchronize(0);
  }
}

// Host in, host out: out_xyz must hold 3*n floats (the filter never grows the cloud).
extern "C" int trgb_voxel_filter(const float* xyz, int64_t n, int stride_floats, float leaf, float* out_xyz, int64_t* n_out) {
  TRGB_ARG(xyz && out_xyz && n_out && n > 0, "null pointer / empty cloud");
  tune_mempool_once();
  float* d_in = nullptr;
  const size_t bytes = (size_t)n * stride_floats * sizeof(float);
  TRGB_CUDA(cudaMallocAsync((void**)&d_in, bytes, 0));
  TRGB_CUDA(cudaMemcpyAsync(d_in, xyz, bytes, cudaMemcpyHostToDevice, 0));
  float* d_out = nullptr;
  int rc = trgb_voxel_filter_dev(d_in, n, stride_floats, leaf, &d_out, n_out);
  cudaFreeAsync(d_in, 0);
  if (d_out) {
    cudaError_t e = cudaMemcpy(out_xyz, d_out, (size_t)(*n_out) * 3 * sizeof(float), cudaMemcpyDeviceToHost);
    trgb_device_free(d_out);
    if (e != cudaSuccess) return cuda_fail(e, "D2H(voxel)", __FILE__, __LINE__);
  }
  return rc;
}
