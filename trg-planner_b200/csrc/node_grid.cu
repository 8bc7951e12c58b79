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

// One work item = a node still looking for its place: (node index, slot it competes for = 2 * parent + side).
// A round: [barrier] read the slot's winner; the winner is placed, every loser steps below the winner, takes
// part in the atomicMin of its next slot right away (all contenders of a slot arrive in the same round, so the
// minimum is complete at the next barrier) and appends itself to the next round's list. One barrier per level;
// the lists shrink with the number of nodes deeper than the level, and once a round fits one CTA the others
// leave and CTA 0 finishes with __syncthreads alone (BFS insertion order makes the tree deep - several hundred
// levels for 5e5 nodes - with a long thin tail).
struct KdCtl {
  unsigned int bar;      // monotonic arrival counter of the grid barrier
  int cnt[3];            // items in the list of round r at cnt[r % 3]
};

__device__ __forceinline__ void kd_barrier(unsigned int* bar, unsigned int target, bool solo) {
  __syncthreads();
  if (solo) return;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    while (*(volatile unsigned int*)bar < target) {}
    __threadfence();
  }
  __syncthreads();
}

// Near the root of a tree grown in BFS order almost every node competes for the same few slots, and same-address
// atomics serialise in L2 (5e5 of them ~ 1 ms per level). So a claim is reduced on the way: lanes of a warp with
// the same slot merge (match_any + reduce_min), warps merge in a small direct-mapped shared-memory table, and a
// CTA sends one atomicMin per distinct slot per round. The next round's list is filled a 1024-item chunk at a
// time with one global atomicAdd per chunk.
constexpr int kKdTable = 2048;

__global__ void __launch_bounds__(1024, 1) k_kd_build(int n, const float2* __restrict__ xy, int2* listA, int2* listB, int* claim,
                                                      int* __restrict__ lo, int* __restrict__ hi, int* __restrict__ parent,
                                                      unsigned char* __restrict__ axis, KdCtl* ctl) {
  __shared__ int s_key[kKdTable], s_val[kKdTable];
  __shared__ int s_wcnt[32], s_woff[32], s_base;
  const int tid = threadIdx.x, gt = blockIdx.x * blockDim.x + tid, GT = gridDim.x * blockDim.x;
  const unsigned lane = tid & 31, warp = tid >> 5;
  for (int h = tid; h < kKdTable; h += blockDim.x) { s_key[h] = -1; s_val[h] = 0x7fffffff; }
  __syncthreads();
  // merged claim of (slot, idx): through the CTA's table when the slot owns (or can take) its bucket
  auto claim_slot = [&](int slot, int idx, unsigned peers) {
    const int mn = __reduce_min_sync(peers, idx);
    if ((unsigned)(__ffs(peers) - 1) != lane) return;
    const int h = (int)(((unsigned)slot * 2654435761u) >> 21);  // 11 bits
    const int old = atomicCAS(&s_key[h], -1, slot);
    if (old == -1 || old == slot) atomicMin(&s_val[h], mn);
    else atomicMin(claim + slot, mn);
  };
  auto flush_table = [&]() {
    __syncthreads();
    for (int h = tid; h < kKdTable; h += blockDim.x) {
      const int key = s_key[h];
      if (key >= 0) { atomicMin(claim + key, s_val[h]); s_key[h] = -1; s_val[h] = 0x7fffffff; }
    }
  };
  // round "-1": everybody but the root competes for the root's low / high slot (axis 0 at the root)
  const float rx = xy[0].x;
  for (int i0 = gt - (int)lane; i0 < n; i0 += GT) {
    const int i = i0 + (int)lane;
    const unsigned act = __ballot_sync(0xffffffffu, i >= 1 && i < n);
    if (i >= 1 && i < n) {
      const int slot = xy[i].x < rx ? 0 : 1;
      listA[i - 1] = make_int2(i, slot);
      claim_slot(slot, i, __match_any_sync(act, slot));
    }
  }
  flush_table();
  if (gt == 0) { parent[0] = -1; axis[0] = 0; }
  bool solo = false;
  unsigned int epoch = 0;
  for (int round = 0;; ++round) {
    epoch += gridDim.x;
    kd_barrier(&ctl->bar, epoch, solo);
    const int m = __ldcg(&ctl->cnt[round % 3]);
    if (m == 0) break;
    if (!solo && m <= (int)blockDim.x) {
      if (blockIdx.x != 0) return;
      solo = true;
    }
    if (gt == 0) ctl->cnt[(round + 2) % 3] = 0;  // next round's appends go there; its last readers are two barriers back
    const int2* cur = (round & 1) ? listB : listA;
    int2* nxt = (round & 1) ? listA : listB;
    int* nxt_cnt = &ctl->cnt[(round + 1) % 3];
    const unsigned char ax = (unsigned char)((round + 1) & 1);  // nodes placed this round sit at depth round + 1
    const int c_first = solo ? 0 : (int)blockIdx.x * (int)blockDim.x, c_stride = solo ? (int)blockDim.x : GT;
    for (int c0 = c_first; c0 < m; c0 += c_stride) {  // CTA-uniform
      const int k = c0 + tid;
      bool lose = false;
      int2 it = make_int2(0, 0);
      int w = 0;
      if (k < m) {
        it = __ldcg(cur + k);
        w = __ldcg(claim + it.y);
        lose = w != it.x;
        if (!lose) {
          const int p = it.y >> 1;
          parent[it.x] = p;
          axis[it.x] = ax;
          if (it.y & 1) hi[p] = it.x; else lo[p] = it.x;
        }
      }
      const unsigned losers = __ballot_sync(0xffffffffu, lose);
      if (lane == 0) s_wcnt[warp] = __popc(losers);
      __syncthreads();
      if (warp == 0) {
        const int v = s_wcnt[lane];
        int inc = v;
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, inc, o);
          if ((int)lane >= o) inc += t;
        }
        s_woff[lane] = inc - v;
        if (lane == 31) s_base = inc ? atomicAdd(nxt_cnt, inc) : 0;
      }
      __syncthreads();
      if (lose) {
        const float2 me = xy[it.x], ww = xy[w];
        const bool low = ax ? (me.y < ww.y) : (me.x < ww.x);
        const int slot = 2 * w + (low ? 0 : 1);
        nxt[s_base + s_woff[warp] + __popc(losers & ((1u << lane) - 1u))] = make_int2(it.x, slot);
        claim_slot(slot, it.x, __match_any_sync(losers, slot));
      }
    }
    flush_table();
  }
}

}  // namespace trgb

// xy: n (x, y) pairs in insertion order. Outputs (host, n entries each): children lo / hi (-1 = none), parent
// (-1 for the root) and split axis of every node — what n successive kd_insert2 calls would have built.
extern "C" int trgb_kdtree_build(const float* xy, int64_t n, int32_t* lo, int32_t* hi, int32_t* parent, uint8_t* axis) {
  TRGB_ARG(xy && lo && hi && parent && axis && n > 0 && n < (1ll << 30), "bad argument");
  trgb::tune_mempool_once();
  cudaStream_t st = nullptr;  // its own non-blocking stream: the legacy default stream would serialise with every other stream
  TRGB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  float2* d_xy = nullptr; int2 *d_la = nullptr, *d_lb = nullptr; int *d_claim = nullptr, *d_lo = nullptr, *d_hi = nullptr, *d_par = nullptr;
  unsigned char* d_axis = nullptr; KdCtl* d_ctl = nullptr;
  const size_t N = (size_t)n;
  TRGB_CUDA(cudaMallocAsync((void**)&d_xy, N * sizeof(float2), st));
  TRGB_CUDA(cudaMallocAsync((void**)&d_la, N * sizeof(int2), st));
  TRGB_CUDA(cudaMallocAsync((void**)&d_lb, N * sizeof(int2), st));
  TRGB_CUDA(cudaMallocAsync((void**)&d_claim, 2 * N * sizeof(int), st));
  TRGB_CUDA(cudaMallocAsync((void**)&d_lo, N * sizeof(int), st));
  TRGB_CUDA(cudaMallocAsync((void**)&d_hi, N * sizeof(int), st));
  TRGB_CUDA(cudaMallocAsync((void**)&d_par, N * sizeof(int), st));
  TRGB_CUDA(cudaMallocAsync((void**)&d_axis, N, st));
  TRGB_CUDA(cudaMallocAsync((void**)&d_ctl, sizeof(KdCtl), st));
  TRGB_CUDA(cudaMemcpyAsync(d_xy, xy, N * sizeof(float2), cudaMemcpyHostToDevice, st));
  TRGB_CUDA(cudaMemsetAsync(d_claim, 0x7f, 2 * N * sizeof(int), st));
  TRGB_CUDA(cudaMemsetAsync(d_lo, 0xff, N * sizeof(int), st));
  TRGB_CUDA(cudaMemsetAsync(d_hi, 0xff, N * sizeof(int), st));
  const KdCtl ctl0{0u, {(int)n - 1, 0, 0}};
  TRGB_CUDA(cudaMemcpyAsync(d_ctl, &ctl0, sizeof(KdCtl), cudaMemcpyHostToDevice, st));
  {
    // every CTA must be resident for the grid barrier: one per SM, launched cooperatively
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 1023) / 1024, (int64_t)sm_count()));
    int nn = (int)n;
    void* args[] = {(void*)&nn, (void*)&d_xy, (void*)&d_la, (void*)&d_lb, (void*)&d_claim, (void*)&d_lo,
                    (void*)&d_hi, (void*)&d_par, (void*)&d_axis, (void*)&d_ctl};
    ProfScope ps("k_kd_build", st, (double)n);
    TRGB_CUDA(cudaLaunchCooperativeKernel((const void*)k_kd_build, dim3(grid), dim3(1024), args, 0, st));
  }
  TRGB_CUDA(cudaMemcpyAsync(lo, d_lo, N * sizeof(int), cudaMemcpyDeviceToHost, st));
  TRGB_CUDA(cudaMemcpyAsync(hi, d_hi, N * sizeof(int), cudaMemcpyDeviceToHost, st));
  TRGB_CUDA(cudaMemcpyAsync(parent, d_par, N * sizeof(int), cudaMemcpyDeviceToHost, st));
  TRGB_CUDA(cudaMemcpyAsync(axis, d_axis, N, cudaMemcpyDeviceToHost, st));
  TRGB_CUDA(cudaStreamSynchronize(st));
  for (void* p : {(void*)d_xy, (void*)d_la, (void*)d_lb, (void*)d_claim, (void*)d_lo, (void*)d_hi, (void*)d_par, (void*)d_axis, (void*)d_ctl})
    cudaFreeAsync(p, st);
  cudaStreamSynchronize(st);
  cudaStreamDestroy(st);
  return TRGB_OK;
}
