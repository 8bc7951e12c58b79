// K7 — batched risk-aware shortest path over the CSR traversal-risk graph.
// Replaces the per-query std::priority_queue A* of TRG::planSafePath (trg.cpp:618-688).
//
// One persistent CTA per concurrent query ("slot"); queries are pulled from a global counter.
// Per query the CTA runs a goal-directed near/far label-correcting search (delta-stepping on
// f = g + h with the reference's consistent heuristic h = 2-D distance to the goal,
// trg.cpp:675): nodes whose f lies below the current threshold are relaxed to a fixed point,
// then the threshold advances by delta. Labels are 64-bit (float g bits << 32 | parent id) and
// relaxed with one atomicMin, so the parent of a node is always the predecessor of its best
// label (ties -> lowest parent id; deterministic). Edge costs follow trg.cpp:674 exactly:
//   g' = fl(g + fl(fl(fl(sf*w) + 1) * dist))        (all float)
// HBM layout per slot: label u64[n] | 4 queues int32[n] | 2 bitmaps u32[ceil(n/32)].
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common.cuh"

struct trgb_graph {
  int32_t n = 0;
  int64_t e = 0;
  int64_t* d_row = nullptr;
  int32_t* d_col = nullptr;
  float* d_w = nullptr;
  float* d_dist = nullptr;
  float* d_cost = nullptr;  // per-edge (sf*w+1)*dist for the current safety factor
  float cost_sf = NAN;
  float mean_cost = 1.f;
  float2* d_pos = nullptr;
  int32_t* d_state = nullptr;
  // scratch
  int nslots = 0;
  unsigned long long* d_label = nullptr;
  int32_t* d_queue = nullptr;
  uint32_t* d_bits = nullptr;
  size_t label_bytes = 0, queue_bytes = 0, bits_bytes = 0;  // allocated sizes
  int device = 0;
  cudaStream_t stream = nullptr;
  int64_t relaxed_edges = 0, queries = 0;  // totals over the handle's life (k_sssp roofline: 20 B per relaxed edge)
};

namespace trgb {

#ifndef TRGB_SSSP_THREADS
#define TRGB_SSSP_THREADS 256
#endif
constexpr int kSsspThreads = TRGB_SSSP_THREADS;
constexpr unsigned long long kInfLabel = 0x7f800000ffffffffull;
constexpr int kGroup = 8;  // lanes cooperating on one node's edge list

__global__ void __launch_bounds__(256) k_edge_cost(const float* __restrict__ w, const float* __restrict__ dist,
                                                   int64_t e, float sf, float* __restrict__ cost) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < e; i += (int64_t)gridDim.x * blockDim.x)
    cost[i] = __fmul_rn(__fadd_rn(__fmul_rn(sf, w[i]), 1.0f), dist[i]);
}

__device__ __forceinline__ float label_g(unsigned long long l) { return __uint_as_float((unsigned)(l >> 32)); }
__device__ __forceinline__ int label_parent(unsigned long long l) { return (int)(unsigned)(l & 0xffffffffull); }
__device__ __forceinline__ unsigned long long make_label(float g, int parent) {
  return ((unsigned long long)__float_as_uint(g) << 32) | (unsigned)parent;
}
__device__ __forceinline__ float heur(const float2* __restrict__ pos, int v, float2 goal) {
  const float2 p = __ldg(pos + v);
  const float dx = __fsub_rn(goal.x, p.x), dy = __fsub_rn(goal.y, p.y);
  return __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
}
// test-and-set one bit; returns true when this call set it
__device__ __forceinline__ bool set_bit(uint32_t* bits, int v) {
  const uint32_t m = 1u << (v & 31);
  return (atomicOr(bits + (v >> 5), m) & m) == 0;
}
__device__ __forceinline__ void clear_bit(uint32_t* bits, int v) { atomicAnd(bits + (v >> 5), ~(1u << (v & 31))); }

struct SsspOut {
  uint8_t* found;
  float* cost;
  float* path_length;
  float* avg_risk;
  int64_t* path_off;   // per query offset into path_ids (claimed with an atomic)
  int32_t* path_len;   // per query length
  int32_t* path_ids;   // ids, start..goal
  long long capacity;
  unsigned long long* cursor;  // [0] = path write cursor, [1] = next query, [2] = edges relaxed (20 B each)
};

__global__ void __launch_bounds__(kSsspThreads) k_sssp(
    int n, const int64_t* __restrict__ row, const int32_t* __restrict__ col, const float* __restrict__ cost,
    const float* __restrict__ ew, const float* __restrict__ ed, const float2* __restrict__ pos,
    const int32_t* __restrict__ state, const int32_t* __restrict__ starts, const int32_t* __restrict__ goals,
    int64_t nq, float delta, unsigned long long* __restrict__ labels, int32_t* __restrict__ queues,
    uint32_t* __restrict__ bitmaps, SsspOut out) {
  const int words = (n + 31) >> 5;
  unsigned long long* label = labels + (size_t)blockIdx.x * n;
  int32_t* q0 = queues + (size_t)blockIdx.x * 4 * n;
  uint32_t* near_bits = bitmaps + (size_t)blockIdx.x * 2 * words;
  uint32_t* far_bits = near_bits + words;

  __shared__ int s_cnt[4];        // sizes: cur near, next near, cur far, next far
  __shared__ long long s_query;
  __shared__ float s_thr, s_best;
  __shared__ int s_done;
  __shared__ unsigned s_minf;
  __shared__ int s_plen;
  __shared__ long long s_poff;
  __shared__ float s_stage[2 * kSsspThreads];
  __shared__ unsigned long long s_relax;
  __shared__ unsigned long long s_dbg[5];  // near passes, threshold steps, far items scanned, nodes expanded, pops pruned

  const int tid = threadIdx.x;
  const int grp = tid / kGroup, gl = tid % kGroup;
  const int ngrp = kSsspThreads / kGroup;

  for (;;) {
    if (tid == 0) s_query = (long long)atomicAdd(out.cursor + 1, 1ull);
    __syncthreads();
    const long long qi = s_query;
    if (qi >= nq) return;
    const int start = starts[qi], goal = goals[qi];
    const float2 gpos = __ldg(pos + goal);

    for (int i = tid; i < n; i += kSsspThreads) label[i] = kInfLabel;
    for (int i = tid; i < 2 * words; i += kSsspThreads) near_bits[i] = 0u;
    __syncthreads();
    int32_t *qcur = q0, *qnxt = q0 + n, *qfar = q0 + 2 * (size_t)n, *qfar2 = q0 + 3 * (size_t)n;
    if (tid == 0) {
      s_relax = 0ull;
      for (int j = 0; j < 5; ++j) s_dbg[j] = 0ull;
      label[start] = make_label(0.f, start);
      qcur[0] = start;
      s_cnt[0] = 1; s_cnt[1] = 0; s_cnt[2] = 0; s_cnt[3] = 0;
      s_thr = __fadd_rn(heur(pos, start, gpos), delta);
      s_best = INFINITY;
      s_done = 0;
    }
    __syncthreads();

    while (!s_done) {
      // ---- relax the near pile to a fixed point under the current threshold.
      // near_bits marks membership of the NEXT queue only: the bits of the nodes about to be
      // processed are cleared before any relaxation of the pass starts, so a node improved
      // while (or after) it is being processed is simply queued again — no lost updates.
      while (s_cnt[0] > 0) {
        const int ncur = s_cnt[0];
        for (int k = tid; k < ncur; k += kSsspThreads) clear_bit(near_bits, qcur[k]);
        __syncthreads();
        const float thr = s_thr;
        const float best = label_g(__ldcg(label + goal));
        unsigned my_relax = 0;
        if (tid == 0) { s_dbg[0] += 1; }
        for (int k = grp; k < ncur; k += ngrp) {
          const int u = qcur[k];
          const float gu = label_g(__ldcg(label + u));
          // goal bound: with a consistent heuristic no path through u beats `best`
          if (__fadd_rn(gu, heur(pos, u, gpos)) > best * 1.000001f) { if (gl == 0) atomicAdd(&s_dbg[4], 1ull); continue; }
          if (gl == 0) atomicAdd(&s_dbg[3], 1ull);
          const int64_t e0 = __ldg(row + u), e1 = __ldg(row + u + 1);
          for (int64_t e = e0 + gl; e < e1; e += kGroup) {
            const int v = __ldg(col + e);
            if (__ldg(state + v) == -1) continue;  // trg.cpp:670 skip Invalid dst
            ++my_relax;
            const float ng = __fadd_rn(gu, __ldg(cost + e));
            const unsigned long long cand = make_label(ng, u);
            const unsigned long long old = atomicMin(label + v, cand);
            if (cand < old && ng < label_g(old)) {
              const float f = __fadd_rn(ng, heur(pos, v, gpos));
              if (f < thr) {
                if (set_bit(near_bits, v)) qnxt[atomicAdd(&s_cnt[1], 1)] = v;
              } else {
                if (set_bit(far_bits, v)) qfar[atomicAdd(&s_cnt[2], 1)] = v;
              }
            }
          }
        }
        my_relax = __reduce_add_sync(FULL, my_relax);
        if ((tid & 31) == 0 && my_relax) atomicAdd(&s_relax, (unsigned long long)my_relax);
        __syncthreads();
        if (tid == 0) { s_cnt[0] = s_cnt[1]; s_cnt[1] = 0; }
        int32_t* t = qcur; qcur = qnxt; qnxt = t;
        __syncthreads();
      }
      // ---- near pile empty: everything with f < thr is final
      if (tid == 0) {
        s_best = label_g(__ldcg(label + goal));
        if (s_best < s_thr || s_cnt[2] == 0) s_done = 1;
        s_minf = 0x7f800000u;
      }
      __syncthreads();
      if (s_done) break;
      // ---- advance the threshold past the smallest f waiting in the far pile
      const int nfar = s_cnt[2];
      const float best = s_best;
      if (tid == 0) { s_dbg[1] += 1; s_dbg[2] += (unsigned long long)nfar; }
      float myf = INFINITY;
      for (int k = tid; k < nfar; k += kSsspThreads) {
        const int v = qfar[k];
        myf = fminf(myf, __fadd_rn(label_g(__ldcg(label + v)), heur(pos, v, gpos)));
      }
      atomicMin(&s_minf, __float_as_uint(myf));  // non-negative floats order like uints
      __syncthreads();
      if (tid == 0) s_thr = __fadd_rn(fmaxf(s_thr, __uint_as_float(s_minf)), delta);
      __syncthreads();
      const float thr = s_thr;
      for (int k = tid; k < nfar; k += kSsspThreads) {
        const int v = qfar[k];
        const float f = __fadd_rn(label_g(__ldcg(label + v)), heur(pos, v, gpos));
        if (f > best * 1.000001f) { clear_bit(far_bits, v); continue; }  // can never matter
        if (f < thr) {
          clear_bit(far_bits, v);
          if (set_bit(near_bits, v)) qcur[atomicAdd(&s_cnt[0], 1)] = v;
        } else {
          qfar2[atomicAdd(&s_cnt[3], 1)] = v;  // stays far (bit remains set)
        }
      }
      __syncthreads();
      if (tid == 0) { s_cnt[2] = s_cnt[3]; s_cnt[3] = 0; }
      int32_t* t = qfar; qfar = qfar2; qfar2 = t;
      __syncthreads();
    }

    // ---- path extraction (goal -> start); sums accumulate in that order like trg.cpp:641-659
    int32_t* rev = q0;  // queues are free now
    float* step_d = reinterpret_cast<float*>(q0 + n);
    float* step_w = reinterpret_cast<float*>(q0 + 2 * (size_t)n);
    if (tid == 0) {
      int plen = 0;
      bool ok = s_best < INFINITY;
      if (ok) {
        int v = goal;
        while (true) {
          rev[plen++] = v;
          if (v == start || plen >= n) break;
          v = label_parent(__ldcg(label + v));
        }
        if (rev[plen - 1] != start) ok = false;
      }
      s_plen = ok ? plen : 0;
    }
    __syncthreads();
    const int plen = s_plen;
    for (int k = tid; k + 1 < plen; k += kSsspThreads) {  // edge rev[k] -> rev[k+1] (its parent)
      const int v = rev[k], u = rev[k + 1];
      float dd = 0.f, ww = 0.f;
      for (int64_t e = row[v]; e < row[v + 1]; ++e)
        if (col[e] == u) { dd = ed[e]; ww = ew[e]; break; }
      step_d[k] = dd;
      step_w[k] = ww;
    }
    __syncthreads();
    float sum_d = 0.f, sum_w = 0.f;
    for (int base = 0; base + 1 < plen; base += kSsspThreads) {
      const int k = base + tid;
      s_stage[tid] = (k + 1 < plen) ? step_d[k] : 0.f;
      s_stage[kSsspThreads + tid] = (k + 1 < plen) ? step_w[k] : 0.f;
      __syncthreads();
      if (tid == 0) {
        const int m = min(kSsspThreads, plen - 1 - base);
        for (int j = 0; j < m; ++j) {
          sum_d = __fadd_rn(sum_d, s_stage[j]);
          sum_w = __fadd_rn(sum_w, s_stage[kSsspThreads + j]);
        }
      }
      __syncthreads();
    }
    if (tid == 0) {
      const bool ok = plen > 0;
      out.found[qi] = ok ? 1 : 0;
      out.cost[qi] = ok ? s_best : 0.f;
      out.path_length[qi] = ok ? sum_d : 0.f;
      out.avg_risk[qi] = ok ? __fdiv_rn(sum_w, (float)plen) : 0.f;
      atomicAdd(out.cursor + 2, s_relax);
      for (int j = 0; j < 5; ++j) atomicAdd(out.cursor + 3 + j, s_dbg[j]);
      s_poff = (long long)atomicAdd(out.cursor, (unsigned long long)plen);
      out.path_off[qi] = s_poff;
      out.path_len[qi] = plen;
    }
    __syncthreads();
    const long long poff = s_poff;
    if (poff + plen <= out.capacity)
      for (int k = tid; k < plen; k += kSsspThreads) out.path_ids[poff + k] = rev[plen - 1 - k];
    __syncthreads();
  }
}

}  // namespace trgb

using namespace trgb;

// All device memory of a graph handle comes from the stream-ordered pool (release threshold
// raised to "never" by trgb::tune_mempool_once): a TRG that is rebuilt every few hundred
// milliseconds re-uses the multi-GB search scratch instead of paying cudaMalloc / cudaFree.
static void gfree(void* p, cudaStream_t st) {
  if (p) cudaFreeAsync(p, st);
}

// The search scratch (labels, queues, bitmaps: ~13 KB per node and slot, several GB for a
// half-million-node graph) outlives its graph handle: a TRG that is rebuilt every few hundred
// milliseconds destroys and re-creates the handle each time, and returning a multi-GB block to the
// pool only to carve it up for the next map's 160 MB left the pool fragmented - the following
// cudaMallocAsync calls then went to the driver for fresh memory (80 - 1100 ms stalls, measured).
// One set of buffers per process is parked here between handles.
namespace {
struct ScratchCache {
  void *label = nullptr, *queue = nullptr, *bits = nullptr;
  size_t label_bytes = 0, queue_bytes = 0, bits_bytes = 0;
  int device = -1;
};
ScratchCache g_scratch;
std::mutex g_scratch_mx;
}  // namespace

static void release_scratch(trgb_graph* g) {  // the handle's stream has been synchronised by the caller
  if (!g->d_label && !g->d_queue && !g->d_bits) return;
  std::lock_guard<std::mutex> lk(g_scratch_mx);
  const bool bigger = g->label_bytes >= g_scratch.label_bytes && g->queue_bytes >= g_scratch.queue_bytes &&
                      g->bits_bytes >= g_scratch.bits_bytes;
  if (g->d_label && g->d_queue && g->d_bits && (g_scratch.label == nullptr || bigger)) {
    if (g_scratch.label) {  // replace the parked (smaller) set
      cudaFreeAsync(g_scratch.label, g->stream); cudaFreeAsync(g_scratch.queue, g->stream); cudaFreeAsync(g_scratch.bits, g->stream);
    }
    g_scratch = ScratchCache{g->d_label, g->d_queue, g->d_bits, g->label_bytes, g->queue_bytes, g->bits_bytes, g->device};
  } else {
    gfree(g->d_label, g->stream); gfree(g->d_queue, g->stream); gfree(g->d_bits, g->stream);
  }
  g->d_label = nullptr; g->d_queue = nullptr; g->d_bits = nullptr;
  g->label_bytes = g->queue_bytes = g->bits_bytes = 0;
  g->nslots = 0;
}

extern "C" void trgb_graph_destroy(trgb_graph* g) {
  if (!g) return;
  cudaStream_t st = g->stream;
  gfree(g->d_row, st); gfree(g->d_col, st); gfree(g->d_w, st); gfree(g->d_dist, st); gfree(g->d_cost, st);
  gfree(g->d_pos, st); gfree(g->d_state, st);
  if (st) cudaStreamSynchronize(st);
  release_scratch(g);
  if (g->stream) {
    cudaStreamSynchronize(g->stream);
    cudaStreamDestroy(g->stream);
  }
  delete g;
}

extern "C" int trgb_graph_upload(trgb_graph** out, const TrgbGraphDesc* d) {
  TRGB_ARG(out && d, "null pointer");
  TRGB_ARG(d->n_nodes > 0 && d->row_ptr && d->pos_xyz && d->state, "empty graph");
  TRGB_ARG(d->n_edges == 0 || (d->col && d->weight && d->dist), "null edge arrays");
  trgb::tune_mempool_once();
  trgb_graph* g = new trgb_graph();
  {
    cudaError_t es = cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking);
    if (es != cudaSuccess) { delete g; return cuda_fail(es, "cudaStreamCreate", __FILE__, __LINE__); }
  }
  g->n = d->n_nodes;
  g->e = d->n_edges;
  cudaGetDevice(&g->device);
  const size_t n = g->n, e = g->e;
  std::vector<float2> pos(n);
  double mean = 0;
  for (size_t i = 0; i < n; ++i) pos[i] = make_float2(d->pos_xyz[3 * i], d->pos_xyz[3 * i + 1]);
  for (size_t i = 0; i < e; ++i) mean += d->dist[i];
  g->mean_cost = e ? (float)(mean / e) : 1.f;
#define UP(dst, src, bytes)                                                                     \
  do {                                                                                          \
    cudaError_t _e = cudaMallocAsync((void**)&(dst), (bytes) ? (bytes) : 1, g->stream);          \
    if (_e == cudaSuccess && (bytes)) _e = cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyHostToDevice, g->stream); \
    if (_e != cudaSuccess) { trgb_graph_destroy(g); return cuda_fail(_e, "graph upload", __FILE__, __LINE__); } \
  } while (0)
  UP(g->d_row, d->row_ptr, (n + 1) * sizeof(int64_t));
  UP(g->d_col, d->col, e * sizeof(int32_t));
  UP(g->d_w, d->weight, e * sizeof(float));
  UP(g->d_dist, d->dist, e * sizeof(float));
  UP(g->d_pos, pos.data(), n * sizeof(float2));
  UP(g->d_state, d->state, n * sizeof(int32_t));
#undef UP
  cudaError_t er = cudaMallocAsync((void**)&g->d_cost, (e ? e : 1) * sizeof(float), g->stream);
  if (er == cudaSuccess) er = cudaStreamSynchronize(g->stream);  // the host arrays may go away after return
  if (er != cudaSuccess) { trgb_graph_destroy(g); return cuda_fail(er, "graph alloc", __FILE__, __LINE__); }
  *out = g;
  return TRGB_OK;
}

static int ensure_slots(trgb_graph* g, int want) {
  if (g->nslots >= want) return TRGB_OK;
  const size_t n = g->n, words = (n + 31) / 32;
  const size_t need_l = (size_t)want * n * sizeof(unsigned long long);
  const size_t need_q = (size_t)want * 4 * n * sizeof(int32_t);
  const size_t need_b = (size_t)want * 2 * words * sizeof(uint32_t);
  if (g->d_label && g->label_bytes >= need_l && g->queue_bytes >= need_q && g->bits_bytes >= need_b) {
    g->nslots = want;
    return TRGB_OK;
  }
  cudaStreamSynchronize(g->stream);
  release_scratch(g);
  {
    std::lock_guard<std::mutex> lk(g_scratch_mx);
    if (g_scratch.label && g_scratch.device == g->device && g_scratch.label_bytes >= need_l &&
        g_scratch.queue_bytes >= need_q && g_scratch.bits_bytes >= need_b) {
      g->d_label = (unsigned long long*)g_scratch.label; g->d_queue = (int32_t*)g_scratch.queue; g->d_bits = (uint32_t*)g_scratch.bits;
      g->label_bytes = g_scratch.label_bytes; g->queue_bytes = g_scratch.queue_bytes; g->bits_bytes = g_scratch.bits_bytes;
      g_scratch = ScratchCache{};
      g->nslots = want;
      return TRGB_OK;
    }
  }
  // a little head room so that the next, slightly larger graph of a rebuild loop fits the parked set
  const size_t al = need_l + need_l / 16, aq = need_q + need_q / 16, ab = need_b + need_b / 16 + 256;
  TRGB_CUDA(cudaMallocAsync((void**)&g->d_label, al, g->stream));
  g->label_bytes = al;
  TRGB_CUDA(cudaMallocAsync((void**)&g->d_queue, aq, g->stream));
  g->queue_bytes = aq;
  TRGB_CUDA(cudaMallocAsync((void**)&g->d_bits, ab, g->stream));
  g->bits_bytes = ab;
  g->nslots = want;
  return TRGB_OK;
}

extern "C" int trgb_sssp_batch(trgb_graph* g, const int32_t* start_ids, const int32_t* goal_ids, int64_t nq,
                               float safety_factor, uint8_t* found, float* cost, float* path_length,
                               float* avg_risk, int64_t* path_offsets, int32_t* path_ids,
                               int64_t path_ids_capacity) {
  TRGB_ARG(g && start_ids && goal_ids && found && cost && path_length && avg_risk && path_offsets, "null pointer");
  TRGB_ARG(path_ids_capacity >= 0 && (path_ids || path_ids_capacity == 0), "bad path buffer");
  if (nq <= 0) { path_offsets[0] = 0; return TRGB_OK; }
  for (int64_t i = 0; i < nq; ++i)
    TRGB_ARG(start_ids[i] >= 0 && start_ids[i] < g->n && goal_ids[i] >= 0 && goal_ids[i] < g->n, "node id out of range");
  cudaStream_t st = g->stream;
  if (!(g->cost_sf == safety_factor)) {
    ProfScope ps("k_edge_cost", st, (double)g->e);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((g->e + 255) / 256, (int64_t)sm_count() * 8));
    k_edge_cost<<<grid, 256, 0, st>>>(g->d_w, g->d_dist, g->e, safety_factor, g->d_cost);
    g->cost_sf = safety_factor;
  }
  // slots: enough CTAs to fill the machine, bounded by the batch and by ~24 GB of scratch
  const size_t per_slot = (size_t)g->n * (8 + 16) + ((size_t)g->n / 4 + 8);
  int want = (int)std::min<int64_t>(nq, (int64_t)sm_count() * (2048 / kSsspThreads));
  want = (int)std::min<size_t>((size_t)want, std::max<size_t>(1, ((size_t)48 << 30) / per_slot));
  int rc = ensure_slots(g, want);
  if (rc) return rc;

  struct Dev { void* p = nullptr; cudaStream_t s = nullptr; ~Dev() { if (p) cudaFreeAsync(p, s); } };
  Dev d_s, d_g, d_found, d_cost, d_len, d_risk, d_off, d_plen, d_ids, d_cur;
#define DALLOC(b, bytes) do { (b).s = st; TRGB_CUDA(cudaMallocAsync(&(b).p, (bytes) ? (bytes) : 1, st)); } while (0)
  DALLOC(d_s, nq * sizeof(int32_t)); DALLOC(d_g, nq * sizeof(int32_t));
  DALLOC(d_found, nq); DALLOC(d_cost, nq * sizeof(float)); DALLOC(d_len, nq * sizeof(float));
  DALLOC(d_risk, nq * sizeof(float)); DALLOC(d_off, nq * sizeof(int64_t)); DALLOC(d_plen, nq * sizeof(int32_t));
  DALLOC(d_ids, (size_t)path_ids_capacity * sizeof(int32_t)); DALLOC(d_cur, 8 * sizeof(unsigned long long));
#undef DALLOC
  TRGB_CUDA(cudaMemcpyAsync(d_s.p, start_ids, nq * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  TRGB_CUDA(cudaMemcpyAsync(d_g.p, goal_ids, nq * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  TRGB_CUDA(cudaMemsetAsync(d_cur.p, 0, 8 * sizeof(unsigned long long), st));
  SsspOut o;
  o.found = (uint8_t*)d_found.p; o.cost = (float*)d_cost.p; o.path_length = (float*)d_len.p;
  o.avg_risk = (float*)d_risk.p; o.path_off = (int64_t*)d_off.p; o.path_len = (int32_t*)d_plen.p;
  o.path_ids = (int32_t*)d_ids.p; o.capacity = path_ids_capacity; o.cursor = (unsigned long long*)d_cur.p;
  // threshold step: 0.5..32 x the mean edge length measured; 1..2 is the flat optimum (profiles/README.md)
  const float delta = 1.5f * std::max(g->mean_cost, 1e-3f);
  {
    ProfScope ps("k_sssp", st, (double)nq);
    k_sssp<<<g->nslots < want ? g->nslots : want, kSsspThreads, 0, st>>>(
        g->n, g->d_row, g->d_col, g->d_cost, g->d_w, g->d_dist, g->d_pos, g->d_state, (const int32_t*)d_s.p,
        (const int32_t*)d_g.p, nq, delta, g->d_label, g->d_queue, g->d_bits, o);
  }
  TRGB_CUDA(cudaGetLastError());
  std::vector<int64_t> off(nq);
  std::vector<int32_t> plen(nq);
  unsigned long long cursor[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  TRGB_CUDA(cudaMemcpyAsync(found, d_found.p, nq, cudaMemcpyDeviceToHost, st));
  TRGB_CUDA(cudaMemcpyAsync(cost, d_cost.p, nq * sizeof(float), cudaMemcpyDeviceToHost, st));
  TRGB_CUDA(cudaMemcpyAsync(path_length, d_len.p, nq * sizeof(float), cudaMemcpyDeviceToHost, st));
  TRGB_CUDA(cudaMemcpyAsync(avg_risk, d_risk.p, nq * sizeof(float), cudaMemcpyDeviceToHost, st));
  TRGB_CUDA(cudaMemcpyAsync(off.data(), d_off.p, nq * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  TRGB_CUDA(cudaMemcpyAsync(plen.data(), d_plen.p, nq * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  TRGB_CUDA(cudaMemcpyAsync(cursor, d_cur.p, sizeof(cursor), cudaMemcpyDeviceToHost, st));
  TRGB_CUDA(cudaStreamSynchronize(st));
  const int64_t total = (int64_t)cursor[0];
  g->relaxed_edges += (int64_t)cursor[2];
  g->queries += nq;
  if (std::getenv("TRGB_SSSP_STATS"))
    fprintf(stderr, "[k_sssp] nq=%lld relaxed=%llu passes=%llu thr_steps=%llu far_scanned=%llu expanded=%llu pruned=%llu delta=%g\n", (long long)nq,
            cursor[2], cursor[3], cursor[4], cursor[5], cursor[6], cursor[7], (double)delta);
  if (total > path_ids_capacity) {
    path_offsets[nq] = total;
    set_error("sssp_batch: path_ids buffer too small; needed size returned in path_offsets[n]");
    return TRGB_E_NOMEM;
  }
  std::vector<int32_t> raw((size_t)total);
  if (total) TRGB_CUDA(cudaMemcpy(raw.data(), d_ids.p, (size_t)total * sizeof(int32_t), cudaMemcpyDeviceToHost));
  int64_t w = 0;
  for (int64_t i = 0; i < nq; ++i) {  // device claims path space in completion order; re-pack in query order
    path_offsets[i] = w;
    if (plen[i]) std::copy(raw.begin() + off[i], raw.begin() + off[i] + plen[i], path_ids + w);
    w += plen[i];
  }
  path_offsets[nq] = w;
  return TRGB_OK;
}

extern "C" int trgb_graph_stats(const trgb_graph* g, int64_t* relaxed_edges, int64_t* queries) {
  TRGB_ARG(g, "null handle");
  if (relaxed_edges) *relaxed_edges = g->relaxed_edges;
  if (queries) *queries = g->queries;
  return TRGB_OK;
}
