// K5 — device-resident uniform grid over GRAPH NODES (append-only), the batched counterpart of the
// reference's kd_nearest2 on `node_tree` inside TRG::expandGraph (trg.cpp:408, kdtree.c:364-417).
// The wavefront scheduler asks, for every accepted sample of a batch, for the exact nearest node
// among the nodes that existed when the batch started; the host then only has to look at the few
// nodes created inside the batch. Distances use the same float arithmetic as kdtree.c
// (fl(fl(dx*dx)+fl(dy*dy)), strict `<`), so the minimum is the same number; exact ties are
// flagged and resolved on the host by the insertion-order tree.
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"

struct trgb_nodes {
  float x0 = 0, y0 = 0, cell = 1, inv = 1;
  int W = 0, H = 0;
  int64_t count = 0, cap = 0;
  int32_t* d_head = nullptr;  // W*H, -1 = empty
  float2* d_xy = nullptr;     // cap
  int32_t* d_next = nullptr;  // cap
};

namespace cg = cooperative_groups;

namespace trgb {

__device__ __forceinline__ int ncell(float v, float origin, float inv, int dim) {
  const float f = floorf(__fmul_rn(__fsub_rn(v, origin), inv));
  if (!(f > 0.0f)) return 0;
  if (f >= (float)dim) return dim - 1;
  return (int)f;
}

__global__ void __launch_bounds__(256) k_nodes_append(const float2* __restrict__ xy_new, int64_t n, int64_t base,
                                                      float x0, float y0, float inv, int W, int H,
                                                      float2* __restrict__ xy, int32_t* __restrict__ next,
                                                      int32_t* __restrict__ head) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float2 p = xy_new[i];
    const int id = (int)(base + i);
    xy[id] = p;
    const int c = ncell(p.y, y0, inv, H) * W + ncell(p.x, x0, inv, W);
    next[id] = atomicExch(head + c, id);
  }
}

struct NearAcc {
  float d2;
  int idx;
  int tie;
};
__device__ __forceinline__ void scan_cell(const int32_t* __restrict__ head, const float2* __restrict__ xy,
                                          const int32_t* __restrict__ next, int W, int H, int cx, int cy, float qx,
                                          float qy, NearAcc& b) {
  if (cx < 0 || cy < 0 || cx >= W || cy >= H) return;
  for (int e = __ldg(head + (size_t)cy * W + cx); e >= 0; e = __ldg(next + e)) {
    const float2 p = __ldg(xy + e);
    const float dx = __fsub_rn(p.x, qx), dy = __fsub_rn(p.y, qy);
    const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    if (d2 < b.d2) { b.d2 = d2; b.idx = e; b.tie = 0; }
    else if (d2 == b.d2 && e != b.idx) { b.tie = 1; }
  }
}

__global__ void __launch_bounds__(256) k_nodes_nearest(const float2* __restrict__ q, int64_t n, float x0, float y0,
                                                       float cell, float inv, int W, int H,
                                                       const int32_t* __restrict__ head, const float2* __restrict__ xy,
                                                       const int32_t* __restrict__ next, int64_t count,
                                                       int32_t* __restrict__ idx_out, float* __restrict__ d2_out,
                                                       uint8_t* __restrict__ tie_out) {
  const int maxr = max(W, H);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float2 p = __ldg(q + i);
    NearAcc b{INFINITY, -1, 0};
    if (count > 0) {
      const int qcx = ncell(p.x, x0, inv, W), qcy = ncell(p.y, y0, inv, H);
      const float fuzz = 4e-6f * (fabsf(p.x) + fabsf(p.y) + cell * (float)(W + H));
      for (int R = 1; R <= maxr; ++R) {
        if (R == 1) {
          for (int yy = qcy - 1; yy <= qcy + 1; ++yy)
            for (int xx = qcx - 1; xx <= qcx + 1; ++xx) scan_cell(head, xy, next, W, H, xx, yy, p.x, p.y, b);
        } else {
          for (int xx = qcx - R; xx <= qcx + R; ++xx) {
            scan_cell(head, xy, next, W, H, xx, qcy - R, p.x, p.y, b);
            scan_cell(head, xy, next, W, H, xx, qcy + R, p.x, p.y, b);
          }
          for (int yy = qcy - R + 1; yy <= qcy + R - 1; ++yy) {
            scan_cell(head, xy, next, W, H, qcx - R, yy, p.x, p.y, b);
            scan_cell(head, xy, next, W, H, qcx + R, yy, p.x, p.y, b);
          }
        }
        // cells outside the scanned block are at least R whole cells away from the query's cell
        const float g = (float)R * cell * 0.9999f - fuzz;
        if (b.idx >= 0 && g > 0.f && b.d2 <= g * g) break;
        if (qcx - R <= 0 && qcy - R <= 0 && qcx + R >= W - 1 && qcy + R >= H - 1) break;
      }
    }
    idx_out[i] = b.idx;
    d2_out[i] = b.d2;
    tie_out[i] = (uint8_t)b.tie;
  }
}

}  // namespace trgb

using namespace trgb;

extern "C" void trgb_nodes_destroy(trgb_nodes* g) {
  if (!g) return;
  cudaDeviceSynchronize();  // destruction is rare; work on any stream may still read the arrays
  cudaFree(g->d_head);
  if (g->d_xy) cudaFreeAsync(g->d_xy, 0);
  if (g->d_next) cudaFreeAsync(g->d_next, 0);
  cudaStreamSynchronize(0);
  delete g;
}

extern "C" int trgb_nodes_create(trgb_nodes** out, float x0, float y0, float x1, float y1, float cell) {
  TRGB_ARG(out, "out is null");
  TRGB_ARG(cell > 0.f && x1 > x0 && y1 > y0, "bad node grid extent");
  trgb::tune_mempool_once();
  trgb_nodes* g = new trgb_nodes();
  g->cell = cell;
  g->inv = 1.0f / cell;
  g->x0 = x0 - 2.f * cell;
  g->y0 = y0 - 2.f * cell;
  const double w = std::floor(((double)x1 - g->x0) / cell) + 4, h = std::floor(((double)y1 - g->y0) / cell) + 4;
  if (w * h > 2.0e9) { delete g; set_error("node grid too large"); return TRGB_E_ARG; }
  g->W = (int)w;
  g->H = (int)h;
  cudaError_t e = cudaMalloc((void**)&g->d_head, (size_t)g->W * g->H * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMemset(g->d_head, 0xff, (size_t)g->W * g->H * sizeof(int32_t));
  if (e != cudaSuccess) { trgb_nodes_destroy(g); return cuda_fail(e, "node grid alloc", __FILE__, __LINE__); }
  *out = g;
  return TRGB_OK;
}

extern "C" int trgb_nodes_reset(trgb_nodes* g, void* stream) {
  TRGB_ARG(g, "null handle");
  g->count = 0;
  TRGB_CUDA(cudaMemsetAsync(g->d_head, 0xff, (size_t)g->W * g->H * sizeof(int32_t), (cudaStream_t)stream));
  return TRGB_OK;
}

extern "C" int64_t trgb_nodes_count(const trgb_nodes* g) { return g ? g->count : 0; }

extern "C" int trgb_nodes_append_launch(trgb_nodes* g, const float* d_xy, int64_t n, void* stream) {
  TRGB_ARG(g && (n == 0 || d_xy), "null pointer");
  if (n <= 0) return TRGB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (g->count + n > g->cap) {
    int64_t want = g->cap ? g->cap : (1 << 20);
    while (want < g->count + n) want *= 2;
    // stream-ordered pool (see tune_mempool_once): growth never stalls on cudaMalloc / cudaFree
    float2* nxy = nullptr;
    int32_t* nnext = nullptr;
    TRGB_CUDA(cudaMallocAsync((void**)&nxy, (size_t)want * sizeof(float2), st));
    TRGB_CUDA(cudaMallocAsync((void**)&nnext, (size_t)want * sizeof(int32_t), st));
    if (g->count) {
      TRGB_CUDA(cudaMemcpyAsync(nxy, g->d_xy, (size_t)g->count * sizeof(float2), cudaMemcpyDeviceToDevice, st));
      TRGB_CUDA(cudaMemcpyAsync(nnext, g->d_next, (size_t)g->count * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    }
    if (g->d_xy) cudaFreeAsync(g->d_xy, st);
    if (g->d_next) cudaFreeAsync(g->d_next, st);
    g->d_xy = nxy; g->d_next = nnext; g->cap = want;
  }
  {
    ProfScope ps("k_nodes_append", st, (double)n);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 8));
    k_nodes_append<<<grid, 256, 0, st>>>(reinterpret_cast<const float2*>(d_xy), n, g->count, g->x0, g->y0, g->inv,
                                         g->W, g->H, g->d_xy, g->d_next, g->d_head);
  }
  TRGB_CUDA(cudaGetLastError());
  g->count += n;
  return TRGB_OK;
}

extern "C" int trgb_nodes_nearest_launch(const trgb_nodes* g, const float* d_xy, int64_t n, int32_t* d_idx,
                                         float* d_d2, uint8_t* d_tie, void* stream) {
  TRGB_ARG(g && (n == 0 || (d_xy && d_idx && d_d2 && d_tie)), "null pointer");
  if (n <= 0) return TRGB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope ps("k_nodes_nearest", st, (double)n);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 8));
  k_nodes_nearest<<<grid, 256, 0, st>>>(reinterpret_cast<const float2*>(d_xy), n, g->x0, g->y0, g->cell, g->inv, g->W,
                                        g->H, g->d_head, g->d_xy, g->d_next, g->count, d_idx, d_d2, d_tie);
  TRGB_CUDA(cudaGetLastError());
  return TRGB_OK;
}

// ------------------------------------------------------------------------------------------------------
// Insertion-order 2-D kd-tree of the graph nodes, built in parallel: the tree the reference grows with one
// kd_insert2 per node in TRG::cleanGraph / addNode (trg.cpp:249, 528-530; kdtree.c:167-194: go left iff
// pos[dir] < node.pos[dir], dir alternating x, y). Its SHAPE decides which of several in-range nodes
// TRG::setGoal picks (head of the result list = last node the traversal visits) and how exact distance ties
// of kd_nearest resolve, so it has to be the reference's tree, not a balanced one.
// Every node descends from the root at once; nodes that reach the same empty child slot in a round all share
// the same root path, and the one inserted first (lowest index) is the slot's occupant in the sequential
// tree: atomicMin decides, the losers step down into the winner's subtree next round. Rounds = tree depth.
// ------------------------------------------------------------------------------------------------------
namespace trgb {

__global__ void __launch_bounds__(256) k_kd_claim(int n, const int* __restrict__ at, const unsigned char* __restrict__ side,
                                                  const unsigned char* __restrict__ placed, int* __restrict__ claim) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (placed[i]) continue;
    atomicMin(claim + 2 * (size_t)at[i] + side[i], i);
  }
}

__global__ void __launch_bounds__(256) k_kd_place(int n, const float2* __restrict__ xy, int* __restrict__ at,
                                                  unsigned char* __restrict__ side, unsigned char* __restrict__ placed,
                                                  const int* __restrict__ claim, int* __restrict__ lo, int* __restrict__ hi,
                                                  int* __restrict__ parent, unsigned char* __restrict__ axis,
                                                  int* __restrict__ remaining) {
  int left = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (placed[i]) continue;
    const int p = at[i];
    const int sd = side[i];
    const int w = claim[2 * (size_t)p + sd];
    const unsigned char ax = axis[p] ^ 1;
    if (w == i) {
      placed[i] = 1;
      parent[i] = p;
      axis[i] = ax;
      if (sd) hi[p] = i; else lo[p] = i;
    } else {
      // the slot went to an earlier node: continue below it (its split axis is the parent's, flipped)
      const float2 me = xy[i], ww = xy[w];
      const bool low = ax ? (me.y < ww.y) : (me.x < ww.x);
      at[i] = w;
      side[i] = low ? 0 : 1;
      ++left;
    }
  }
  if (left) atomicAdd(remaining, left);
}

// all rounds in one cooperative launch: claim, grid barrier, place, grid barrier, until nobody is left
__global__ void __launch_bounds__(256) k_kd_build(int n, const float2* __restrict__ xy, int* __restrict__ at,
                                                  unsigned char* __restrict__ side, unsigned char* __restrict__ placed,
                                                  int* __restrict__ claim, int* __restrict__ lo, int* __restrict__ hi,
                                                  int* __restrict__ parent, unsigned char* __restrict__ axis,
                                                  int* __restrict__ counters /* [2] remaining per round parity, zeroed */) {
  cg::grid_group grid = cg::this_grid();
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, GT = gridDim.x * blockDim.x;
  for (int round = 0;; ++round) {
    for (int i = gt; i < n; i += GT)
      if (!placed[i]) atomicMin(claim + 2 * (size_t)at[i] + side[i], i);
    if (gt == 0) counters[round & 1] = 0;  // (its last readers were two rounds ago: everybody has passed two barriers since)
    grid.sync();
    int left = 0;
    for (int i = gt; i < n; i += GT) {
      if (placed[i]) continue;
      const int p = at[i];
      const int sd = side[i];
      const int w = ((volatile int*)claim)[2 * (size_t)p + sd];
      const unsigned char ax = axis[p] ^ 1;
      if (w == i) {
        placed[i] = 1;
        parent[i] = p;
        axis[i] = ax;
        if (sd) hi[p] = i; else lo[p] = i;
      } else {
        const float2 me = xy[i], ww = xy[w];
        const bool low = ax ? (me.y < ww.y) : (me.x < ww.x);
        at[i] = w;
        side[i] = low ? 0 : 1;
        ++left;
      }
    }
    left = __reduce_add_sync(0xffffffffu, left);
    if ((threadIdx.x & 31) == 0 && left) atomicAdd(counters + (round & 1), left);
    grid.sync();
    if (((volatile int*)counters)[round & 1] == 0) break;
  }
}

}  // namespace trgb

// xy: n (x, y) pairs in insertion order. Outputs (host, n entries each): children lo / hi (-1 = none), parent
// (-1 for the root) and split axis of every node — what n successive kd_insert2 calls would have built.
extern "C" int trgb_kdtree_build(const float* xy, int64_t n, int32_t* lo, int32_t* hi, int32_t* parent, uint8_t* axis) {
  TRGB_ARG(xy && lo && hi && parent && axis && n > 0 && n < (1ll << 31), "bad argument");
  trgb::tune_mempool_once();
  cudaStream_t st = nullptr;  // its own non-blocking stream: the legacy default stream would serialise with every other stream
  TRGB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  float2* d_xy = nullptr; int *d_at = nullptr, *d_claim = nullptr, *d_lo = nullptr, *d_hi = nullptr, *d_par = nullptr, *d_rem = nullptr;
  unsigned char *d_side = nullptr, *d_placed = nullptr, *d_axis = nullptr;
  const size_t N = (size_t)n;
  TRGB_CUDA(cudaMallocAsync((void**)&d_xy, N * sizeof(float2), st));
  TRGB_CUDA(cudaMallocAsync((void**)&d_at, N * sizeof(int), st));
  TRGB_CUDA(cudaMallocAsync((void**)&d_claim, 2 * N * sizeof(int), st));
  TRGB_CUDA(cudaMallocAsync((void**)&d_lo, N * sizeof(int), st));
  TRGB_CUDA(cudaMallocAsync((void**)&d_hi, N * sizeof(int), st));
  TRGB_CUDA(cudaMallocAsync((void**)&d_par, N * sizeof(int), st));
  TRGB_CUDA(cudaMallocAsync((void**)&d_rem, 2 * sizeof(int), st));
  TRGB_CUDA(cudaMallocAsync((void**)&d_side, N, st));
  TRGB_CUDA(cudaMallocAsync((void**)&d_placed, N, st));
  TRGB_CUDA(cudaMallocAsync((void**)&d_axis, N, st));
  TRGB_CUDA(cudaMemcpyAsync(d_xy, xy, N * sizeof(float2), cudaMemcpyHostToDevice, st));
  TRGB_CUDA(cudaMemsetAsync(d_claim, 0x7f, 2 * N * sizeof(int), st));
  TRGB_CUDA(cudaMemsetAsync(d_lo, 0xff, N * sizeof(int), st));
  TRGB_CUDA(cudaMemsetAsync(d_hi, 0xff, N * sizeof(int), st));
  TRGB_CUDA(cudaMemsetAsync(d_par, 0xff, N * sizeof(int), st));
  TRGB_CUDA(cudaMemsetAsync(d_placed, 0, N, st));
  TRGB_CUDA(cudaMemsetAsync(d_axis, 0, N, st));
  TRGB_CUDA(cudaMemsetAsync(d_at, 0, N * sizeof(int), st));
  TRGB_CUDA(cudaMemsetAsync(d_rem, 0, 2 * sizeof(int), st));
  // the root is node 0; everybody else starts at its low / high slot
  const unsigned char one = 1;
  TRGB_CUDA(cudaMemcpyAsync(d_placed, &one, 1, cudaMemcpyHostToDevice, st));
  {
    std::vector<unsigned char> side0(N);
    for (size_t i = 0; i < N; ++i) side0[i] = xy[2 * i] < xy[0] ? 0 : 1;  // axis 0 at the root
    TRGB_CUDA(cudaMemcpyAsync(d_side, side0.data(), N, cudaMemcpyHostToDevice, st));
    TRGB_CUDA(cudaStreamSynchronize(st));
  }
  if (n > 1) {
    int per_sm = 0;
    TRGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)k_kd_build, 256, 0));
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * std::max(1, std::min(per_sm, 4))));
    int nn = (int)n;
    void* args[] = {(void*)&nn, (void*)&d_xy, (void*)&d_at, (void*)&d_side, (void*)&d_placed, (void*)&d_claim, (void*)&d_lo,
                    (void*)&d_hi, (void*)&d_par, (void*)&d_axis, (void*)&d_rem};
    ProfScope ps("k_kd_build", st, (double)n);
    TRGB_CUDA(cudaLaunchCooperativeKernel((const void*)k_kd_build, dim3(grid), dim3(256), args, 0, st));
  }
  TRGB_CUDA(cudaMemcpyAsync(lo, d_lo, N * sizeof(int), cudaMemcpyDeviceToHost, st));
  TRGB_CUDA(cudaMemcpyAsync(hi, d_hi, N * sizeof(int), cudaMemcpyDeviceToHost, st));
  TRGB_CUDA(cudaMemcpyAsync(parent, d_par, N * sizeof(int), cudaMemcpyDeviceToHost, st));
  TRGB_CUDA(cudaMemcpyAsync(axis, d_axis, N, cudaMemcpyDeviceToHost, st));
  TRGB_CUDA(cudaStreamSynchronize(st));
  for (void* p : {(void*)d_xy, (void*)d_at, (void*)d_claim, (void*)d_lo, (void*)d_hi, (void*)d_par, (void*)d_rem, (void*)d_side,
                  (void*)d_placed, (void*)d_axis})
    cudaFreeAsync(p, st);
  cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  return TRGB_OK;
}
