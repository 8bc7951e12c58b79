// K1 — device map index build. Replaces the kd_insert2 loop of TRG::setGlobalMap /
// TRG::setLocalMap (trg.cpp:185-188, 203-206 -> kdtree.c:167-209): an O(N*depth) pointer-chasing
// insertion becomes a counting sort of the cloud into row-major grid cells.
//
// Passes over HBM (algorithmic 32 B/point = one float4 read + one sorted float4 write):
//   k_bbox      read xyz                              -> 4 floats
//   k_count     read xyz, atomicAdd cell histogram    (REDG, spread addresses)
//   scan        exclusive prefix over W*H+1 counters  (3 small kernels)
//   k_scatter   read xyz, write float4 at cell_start[cell] + atomic slot
//   k_sort_cell per cell: order its few points by original index (determinism: the slot
//               order of the atomics is arbitrary; the final layout is not)
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <vector>

#include <chrono>
#include <cstdio>
#include <cstdlib>

#include <mutex>

#include "common.cuh"

namespace trgb {

__device__ __forceinline__ float3 load_pt(const float* __restrict__ pts, int64_t i, int stride) {
  if (stride == 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(pts) + i);
    return make_float3(v.x, v.y, v.z);
  }
  const float* p = pts + i * stride;
  return make_float3(__ldg(p), __ldg(p + 1), __ldg(p + 2));
}

__device__ __forceinline__ void atomic_min_f(float* addr, float v) {
  // valid for any sign via the ordered-key trick on ints
  if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}

// bbox[0..3] = minx, miny, maxx, maxy (initialised to +FLT_MAX / -FLT_MAX by the host)
__global__ void __launch_bounds__(256) k_bbox(const float* __restrict__ pts, int64_t n, int stride,
                                              float* __restrict__ bbox) {
  float mnx = FLT_MAX, mny = FLT_MAX, mxx = -FLT_MAX, mxy = -FLT_MAX;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float3 p = load_pt(pts, i, stride);
    mnx = fminf(mnx, p.x); mny = fminf(mny, p.y);
    mxx = fmaxf(mxx, p.x); mxy = fmaxf(mxy, p.y);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    mnx = fminf(mnx, __shfl_xor_sync(FULL, mnx, d));
    mny = fminf(mny, __shfl_xor_sync(FULL, mny, d));
    mxx = fmaxf(mxx, __shfl_xor_sync(FULL, mxx, d));
    mxy = fmaxf(mxy, __shfl_xor_sync(FULL, mxy, d));
  }
  if ((threadIdx.x & 31) == 0) {
    atomic_min_f(bbox + 0, mnx);
    atomic_min_f(bbox + 1, mny);
    atomic_max_f(bbox + 2, mxx);
    atomic_max_f(bbox + 3, mxy);
  }
}

__global__ void __launch_bounds__(256) k_count(const float* __restrict__ pts, int64_t n, int stride,
                                               MapView m, uint32_t* __restrict__ counts) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float3 p = load_pt(pts, i, stride);
    const int cx = cell_coord(p.x, m.x0, m.inv_cell, m.W);
    const int cy = cell_coord(p.y, m.y0, m.inv_cell, m.H);
    atomicAdd(counts + (size_t)cy * m.W + cx, 1u);
  }
}

// ---- exclusive scan of uint32 counters: block sums -> scan of sums -> apply -------------
constexpr int kScanBlock = 256;
constexpr int kScanItems = 8;  // per thread
constexpr int kScanTile = kScanBlock * kScanItems;

__global__ void __launch_bounds__(kScanBlock) k_scan_reduce(const uint32_t* __restrict__ in, int64_t n,
                                                            uint32_t* __restrict__ block_sums) {
  __shared__ uint32_t warp_sums[kScanBlock / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const int64_t i = base + k * kScanBlock + threadIdx.x;
    if (i < n) s += in[i];
  }
  s = __reduce_add_sync(FULL, s);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < kScanBlock / 32; ++w) t += warp_sums[w];
    block_sums[blockIdx.x] = t;
  }
}

// single CTA: exclusive scan of block sums in place (nb up to a few hundred thousand)
__global__ void __launch_bounds__(1024) k_scan_sums(uint32_t* __restrict__ sums, int nb) {
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < nb; base += 1024) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < nb ? sums[i] : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(FULL, inc, d);
      if (lane >= d) inc += t;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      const uint32_t w = warp_tot[lane];
      uint32_t winc = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, winc, d);
        if (lane >= d) winc += t;
      }
      warp_tot[lane] = winc - w;  // exclusive offset of each warp inside the chunk
    }
    __syncthreads();
    const uint32_t excl = carry_s + warp_tot[wid] + (inc - v);
    if (i < nb) sums[i] = excl;
    __syncthreads();                               // everyone has read carry_s / warp_tot
    if (threadIdx.x == 1023) carry_s = excl + v;  // inclusive total through this chunk
    __syncthreads();
  }
}

// out[i] = exclusive prefix; also writes out[n] = total when i == n-1 (out has n+1 entries)
__global__ void __launch_bounds__(kScanBlock) k_scan_apply(const uint32_t* __restrict__ in, int64_t n,
                                                           const uint32_t* __restrict__ block_offs,
                                                           uint32_t* __restrict__ out) {
  __shared__ uint32_t warp_tot[kScanBlock / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t tsum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0u;
    tsum += v[k];
  }
  uint32_t inc = tsum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(FULL, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  uint32_t woff = 0;
  for (int w = 0; w < wid; ++w) woff += warp_tot[w];
  uint32_t run = block_offs[blockIdx.x] + woff + (inc - tsum);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < n) {
      out[base + k] = run;
      run += v[k];
      if (base + k == n - 1) out[n] = run;
    }
  }
}

__global__ void __launch_bounds__(256) k_scatter(const float* __restrict__ pts, int64_t n, int stride,
                                                 MapView m, const uint32_t* __restrict__ cell_start,
                                                 uint32_t* __restrict__ fill, float4* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float3 p = load_pt(pts, i, stride);
    const int cx = cell_coord(p.x, m.x0, m.inv_cell, m.W);
    const int cy = cell_coord(p.y, m.y0, m.inv_cell, m.H);
    const size_t c = (size_t)cy * m.W + cx;
    const uint32_t slot = cell_start[c] + atomicAdd(fill + c, 1u);
    out[slot] = make_float4(p.x, p.y, p.z, __int_as_float((int)i));
  }
}

// one thread per cell: insertion sort by original index (cells hold a handful of points). A cell of an
// un-voxelised or duplicate-laden cloud can hold thousands: beyond 48 points the thread switches to an in-place
// heap sort (O(k log k): 1e4 points cost 1e5 steps instead of the 1e8 of the insertion sort).
__device__ __forceinline__ void cell_sift_down(float4* a, uint32_t root, uint32_t n) {
  const float4 x = a[root];
  const int kx = __float_as_int(x.w);
  for (;;) {
    uint32_t child = 2 * root + 1;
    if (child >= n) break;
    if (child + 1 < n && __float_as_int(a[child + 1].w) > __float_as_int(a[child].w)) ++child;
    if (__float_as_int(a[child].w) <= kx) break;
    a[root] = a[child];
    root = child;
  }
  a[root] = x;
}

__global__ void __launch_bounds__(256) k_sort_cell(const uint32_t* __restrict__ cell_start,
                                                   int64_t ncells, float4* __restrict__ pts) {
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncells;
       c += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t s = cell_start[c], e = cell_start[c + 1];
    if (e - s > 48u) {
      float4* a = pts + s;
      const uint32_t n = e - s;
      for (uint32_t i = n / 2; i-- > 0;) cell_sift_down(a, i, n);
      for (uint32_t m = n - 1; m > 0; --m) {
        const float4 t = a[0]; a[0] = a[m]; a[m] = t;
        cell_sift_down(a, 0, m);
      }
      continue;
    }
    for (uint32_t i = s + 1; i < e; ++i) {
      const float4 key = pts[i];
      const int ki = __float_as_int(key.w);
      uint32_t j = i;
      while (j > s && __float_as_int(pts[j - 1].w) > ki) {
        pts[j] = pts[j - 1];
        --j;
      }
      if (j != i) pts[j] = key;
    }
  }
}

// Device memory comes from the stream-ordered pool with an unbounded release threshold: a TRG that
// is rebuilt again and again (bench steps, per-scan local maps) reuses the same blocks instead of
// paying cudaMalloc / cudaFree (which occasionally stall for a second on a freshly freed 100+ MB).
void tune_mempool_once() {
  static bool done = false;
  if (done) return;
  done = true;
  int dev = 0;
  cudaMemPool_t pool;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  cudaGetLastError();
}

// TRGB_MAP_TRACE=1: report (stderr) any phase of a map build / teardown that takes more than 20 ms
struct MapTrace {
  bool on;
  std::chrono::steady_clock::time_point t0;
  MapTrace() : on(std::getenv("TRGB_MAP_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
  void lap(const char* what) {
    if (!on) return;
    const auto t1 = std::chrono::steady_clock::now();
    const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    if (ms > 20.0) fprintf(stderr, "[trgb map trace] %s: %.1f ms\n", what, ms);
    t0 = t1;
  }
};

static int build_index(trgb_map* m, const float* d_in, int64_t n, int stride, float cell) {
  cudaStream_t st = m->stream;
  tune_mempool_once();
  MapTrace tr;
  const int sms = sm_count();
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sms * 16);

  float* d_bbox = nullptr;
  TRGB_CUDA(cudaMallocAsync((void**)&d_bbox, 4 * sizeof(float), st));
  const float init[4] = {FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX};
  TRGB_CUDA(cudaMemcpyAsync(d_bbox, init, sizeof(init), cudaMemcpyHostToDevice, st));
  {
    ProfScope ps("k_bbox", st, (double)n);
    k_bbox<<<grid, 256, 0, st>>>(d_in, n, stride, d_bbox);
  }
  float bbox[4];
  TRGB_CUDA(cudaMemcpyAsync(bbox, d_bbox, sizeof(bbox), cudaMemcpyDeviceToHost, st));
  tr.lap("bbox launch");
  TRGB_CUDA(cudaStreamSynchronize(st));
  tr.lap("bbox sync");
  cudaFreeAsync(d_bbox, st);
  if (!(bbox[0] <= bbox[2]) || !(bbox[1] <= bbox[3]) || !std::isfinite(bbox[0]) ||
      !std::isfinite(bbox[3])) {
    set_error("map_create: non-finite or empty bounding box");
    return TRGB_E_ARG;
  }
  MapView& v = m->view;
  v.cell = cell;
  v.inv_cell = 1.0f / cell;
  v.x0 = bbox[0];
  v.y0 = bbox[1];
  const double wx = ((double)bbox[2] - bbox[0]) / cell, wy = ((double)bbox[3] - bbox[1]) / cell;
  if (wx > 60000.0 || wy > 60000.0) {
    set_error("map_create: grid would exceed 60000 cells per side; use a larger cell_size");
    return TRGB_E_ARG;
  }
  v.W = (int)std::floor(wx) + 2;
  v.H = (int)std::floor(wy) + 2;
  v.n = n;
  const int64_t ncells = (int64_t)v.W * v.H;

  uint32_t *d_counts = nullptr, *d_fill = nullptr, *d_sums = nullptr;
  const int nb = (int)((ncells + kScanTile - 1) / kScanTile);
  TRGB_CUDA(cudaMallocAsync((void**)&d_counts, ncells * sizeof(uint32_t), st));
  TRGB_CUDA(cudaMallocAsync((void**)&d_fill, ncells * sizeof(uint32_t), st));
  TRGB_CUDA(cudaMallocAsync((void**)&d_sums, (size_t)std::max(nb, 1) * sizeof(uint32_t), st));
  TRGB_CUDA(cudaMallocAsync((void**)&m->d_cell_start, (ncells + 1) * sizeof(uint32_t), st));
  // kPtsPad zeroed points behind the last one: the grouped gather of the thread-per-query kernels
  // reads whole groups of four and masks what lies beyond a run (common.cuh)
  TRGB_CUDA(cudaMallocAsync((void**)&m->d_pts, ((size_t)n + kPtsPad) * sizeof(float4), st));
  TRGB_CUDA(cudaMemsetAsync(m->d_pts + n, 0, kPtsPad * sizeof(float4), st));
  m->device_bytes = (ncells + 1) * (int64_t)sizeof(uint32_t) + (n + kPtsPad) * (int64_t)sizeof(float4);
  tr.lap("allocations");
  TRGB_CUDA(cudaMemsetAsync(d_counts, 0, ncells * sizeof(uint32_t), st));
  TRGB_CUDA(cudaMemsetAsync(d_fill, 0, ncells * sizeof(uint32_t), st));
  {
    ProfScope ps("k_count", st, (double)n);
    k_count<<<grid, 256, 0, st>>>(d_in, n, stride, v, d_counts);
  }
  {
    ProfScope ps("k_scan", st, (double)ncells);
    k_scan_reduce<<<nb, kScanBlock, 0, st>>>(d_counts, ncells, d_sums);
    k_scan_sums<<<1, 1024, 0, st>>>(d_sums, nb);
    k_scan_apply<<<nb, kScanBlock, 0, st>>>(d_counts, ncells, d_sums, m->d_cell_start);
  }
  {
    ProfScope ps("k_scatter", st, (double)n);
    k_scatter<<<grid, 256, 0, st>>>(d_in, n, stride, v, m->d_cell_start, d_fill, m->d_pts);
  }
  {
    ProfScope ps("k_sort_cell", st, (double)n);
    const int g2 = (int)std::min<int64_t>((ncells + 255) / 256, (int64_t)sms * 16);
    k_sort_cell<<<g2, 256, 0, st>>>(m->d_cell_start, ncells, m->d_pts);
  }
  TRGB_CUDA(cudaGetLastError());
  cudaFreeAsync(d_counts, st);
  cudaFreeAsync(d_fill, st);
  cudaFreeAsync(d_sums, st);
  tr.lap("index launches");
  TRGB_CUDA(cudaStreamSynchronize(st));
  tr.lap("index sync");
  v.pts = m->d_pts;
  v.cell_start = m->d_cell_start;
  return TRGB_OK;
}

}  // namespace trgb

using namespace trgb;

extern "C" int trgb_map_create_dev(trgb_map** out, const float* dev_pts, int64_t n, int stride_floats,
                                   float cell_size) {
  TRGB_ARG(out != nullptr, "out is null");
  TRGB_ARG(dev_pts != nullptr && n > 0, "empty point cloud");
  TRGB_ARG(n < (int64_t)0x7fffffff, "more than 2^31-1 points per map handle");
  TRGB_ARG(stride_floats == 3 || stride_floats == 4, "stride_floats must be 3 or 4");
  TRGB_ARG(cell_size > 0.f && std::isfinite(cell_size), "cell_size must be > 0");
  trgb_map* m = new trgb_map();
  MapTrace tr;
  cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
  tr.lap("stream create");
  if (e != cudaSuccess) {
    delete m;
    return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__);
  }
  m->n = n;
  int rc = build_index(m, dev_pts, n, stride_floats, cell_size);
  if (rc != TRGB_OK) {
    trgb_map_destroy(m);
    return rc;
  }
  *out = m;
  return TRGB_OK;
}

// The device-side landing buffer of a host cloud is parked between calls (one per process, grow-only): a 50 M-point
// cloud is 600 MB, and taking / returning a block of that size from the pool on every map rebuild, next to the
// 800 MB index it is sorted into, sent every few allocations to the driver for fresh memory (140 - 570 ms stalls
// measured in a rebuild loop). A second thread that finds the buffer busy falls back to a pool allocation.
namespace {
struct LandingBuffer {
  std::mutex mx;
  float* p = nullptr;
  size_t bytes = 0;
  int device = -1;
} g_landing;
}  // namespace

extern "C" int trgb_map_create(trgb_map** out, const float* host_pts, int64_t n, int stride_floats,
                               float cell_size) {
  TRGB_ARG(host_pts != nullptr && n > 0, "empty point cloud");
  TRGB_ARG(stride_floats == 3 || stride_floats == 4, "stride_floats must be 3 or 4");
  tune_mempool_once();
  const size_t bytes = (size_t)n * stride_floats * sizeof(float);
  int dev = 0;
  TRGB_CUDA(cudaGetDevice(&dev));
  std::unique_lock<std::mutex> parked(g_landing.mx, std::try_to_lock);
  float* d_in = nullptr;
  bool own = false;
  if (parked.owns_lock()) {
    if (g_landing.device != dev || g_landing.bytes < bytes) {
      if (g_landing.p) {
        if (g_landing.device >= 0 && g_landing.device != dev) { cudaSetDevice(g_landing.device); cudaFree(g_landing.p); cudaSetDevice(dev); }
        else cudaFree(g_landing.p);
      }
      g_landing.p = nullptr; g_landing.bytes = 0; g_landing.device = dev;
      const size_t want = bytes + bytes / 8;
      cudaError_t ea = cudaMalloc((void**)&g_landing.p, want);
      if (ea != cudaSuccess) return cuda_fail(ea, "cudaMalloc (landing buffer of the cloud)", __FILE__, __LINE__);
      g_landing.bytes = want;
    }
    d_in = g_landing.p;
  } else {
    TRGB_CUDA(cudaMallocAsync((void**)&d_in, bytes, 0));
    own = true;
  }
  cudaError_t e = cudaMemcpyAsync(d_in, host_pts, bytes, cudaMemcpyHostToDevice, 0);
  if (e == cudaSuccess) e = cudaStreamSynchronize(0);
  if (e != cudaSuccess) {
    if (own) cudaFreeAsync(d_in, 0);
    return cuda_fail(e, "cudaMemcpy H2D (map points)", __FILE__, __LINE__);
  }
  int rc = trgb_map_create_dev(out, d_in, n, stride_floats, cell_size);
  // (the index build reads d_in on the new handle's stream: finish it before the buffer can be handed out again)
  if (rc == TRGB_OK && !own) rc = trgb_map_sync(*out);
  if (own) cudaFreeAsync(d_in, 0);
  return rc;
}

extern "C" void trgb_map_destroy(trgb_map* m) {
  if (!m) return;
  if (m->d_pts) cudaFreeAsync(m->d_pts, m->stream);
  if (m->d_cell_start) cudaFreeAsync(m->d_cell_start, m->stream);
  MapTrace tr;
  if (m->stream) cudaStreamSynchronize(m->stream);
  tr.lap("destroy: stream sync");
  if (m->d_stage) cudaFree(m->d_stage);
  if (m->h_stage) cudaFreeHost(m->h_stage);
  tr.lap("destroy: staging free");
  if (m->stream) cudaStreamDestroy(m->stream);
  tr.lap("destroy: stream destroy");
  delete m;
}

extern "C" int trgb_map_info(const trgb_map* m, TrgbMapInfo* info) {
  TRGB_ARG(m && info, "null handle");
  info->n_points = m->n;
  info->grid_w = m->view.W;
  info->grid_h = m->view.H;
  info->origin_x = m->view.x0;
  info->origin_y = m->view.y0;
  info->cell_size = m->view.cell;
  info->device_bytes = m->device_bytes;
  return TRGB_OK;
}

// the indexed points where they lie in HBM: n records (x, y, z, bit-cast original index), sorted by cell; complete
// once the handle's stream has been synchronised (trgb_map_sync)
extern "C" int trgb_map_points(const trgb_map* m, const float** d_xyzi, int64_t* n) {
  TRGB_ARG(m && d_xyzi && n, "null pointer");
  *d_xyzi = reinterpret_cast<const float*>(m->view.pts);
  *n = m->n;
  return TRGB_OK;
}

extern "C" void* trgb_map_stream(const trgb_map* m) { return m ? (void*)m->stream : nullptr; }

extern "C" int trgb_map_set_option(trgb_map* m, const char* key, int value) {
  TRGB_ARG(m && key, "null handle");
  if (std::string(key) == "force_warp_path") m->force_warp_path = value;
  else if (std::string(key) == "use_staging") m->use_staging = value;
  else TRGB_ARG(false, "unknown option");
  return TRGB_OK;
}

extern "C" int trgb_map_sync(const trgb_map* m) {
  TRGB_ARG(m, "null handle");
  TRGB_CUDA(cudaStreamSynchronize(m->stream));
  return TRGB_OK;
}
