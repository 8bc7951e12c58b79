// K7 — batched risk-aware shortest path over the traversal-risk graph.
// Replaces the per-query std::priority_queue A* of TRG::planSafePath (trg.cpp:618-688).
//
// Search graph in HBM (built once per graph, on the device, from CSR arrays in the caller's numbering):
//   nodes are renumbered along a Morton curve over (x, y), so that the labels, node records and edge
//   lists a search front touches are neighbours in memory (the caller's ids are hash-map iteration
//   order, i.e. scattered); ids only exist at the boundary (ext2int on the way in, node.y on the way out)
//   node  int4[n+1]   {first edge, caller's id, x bits, y bits}      (entry n = sentinel: first edge = e)
//   edge  float4[e]   {dst (int bits), cost = fl(fl(fl(sf*w)+1)*dist), dst x, dst y}: everything one
//                     relaxation needs in one 16-byte load, no second trip for the heuristic of dst
//   w, dist float[e]  in the same edge order (path sums; cost recomputation for another safety factor)
//
// One persistent CTA per concurrent query ("slot"); queries are pulled from a global counter, longest
// (by straight-line distance) first. Per query the CTA runs a goal-directed near/far label-correcting
// search (delta-stepping on f = g + h with the reference's consistent heuristic h = 2-D distance to the
// goal, trg.cpp:675): nodes whose f lies below the current threshold are relaxed to a fixed point, then
// the threshold advances by delta. Labels are 64-bit (float g bits << 32 | caller's id of the parent) and
// relaxed with one atomicMin, so the parent of a node is always the predecessor of its best label
// (ties -> lowest parent id; deterministic and independent of the internal numbering). Edge costs follow
// trg.cpp:674 exactly:   g' = fl(g + fl(fl(fl(sf*w) + 1) * dist))        (all float)
// A search is a chain of dependent memory round trips per pass (queue -> label, node -> edges ->
// atomicMin -> queue bit), so the critical path of the longest query bounds the batch: wide CTAs (one
// 8-lane group per frontier node, all of a pass's nodes in flight at once), one __syncthreads per pass.
// HBM scratch per slot: label u64[n] | 4 queues int32[n] | 3 bitmaps u32[ceil(n/32)].
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common.cuh"

#ifndef TRGB_SSSP_THREADS
#define TRGB_SSSP_THREADS 512
#endif

struct trgb_graph {
  int32_t n_ext = 0;  // the caller's ids: 0 .. n_ext-1
  int32_t n = 0;      // nodes of the search graph (internal numbering)
  int64_t e = 0;
  int4* d_node = nullptr;
  float4* d_edge = nullptr;
  float* d_w = nullptr;
  float* d_dist = nullptr;
  int32_t* d_state = nullptr;    // internal numbering
  int32_t* d_ext2int = nullptr;  // n_ext entries, -1 = not part of the graph
  double* d_partial = nullptr;   // kPartials partial sums of dist (mean edge length -> threshold step)
  float cost_sf = NAN;
  float mean_cost = -1.f;        // < 0: not fetched yet
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

constexpr int kSsspThreads = TRGB_SSSP_THREADS;
constexpr unsigned long long kInfLabel = 0x7f800000ffffffffull;
constexpr int kGroup = 8;  // lanes cooperating on one node's edge list
constexpr int kPartials = 256;
// goal bound: a node is dropped when g + h exceeds the best goal label by more than the parity tolerance (1e-5
// relative): float sums over several hundred edges carry ~sqrt(hops) ulp of rounding, a few ulp of slack could
// prune a node of the true optimum on zero-risk, near-straight stretches where h is tight
constexpr float kGoalSlack = 1.00001f;

// ------------------------------------------------------------------------------------------------------
// search-graph construction
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int float_order(float f) {  // monotonic float -> int
  const int i = __float_as_int(f);
  return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float order_float(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }
__device__ __forceinline__ uint32_t spread16(uint32_t v) {
  v &= 0xffffu;
  v = (v | (v << 8)) & 0x00ff00ffu;
  v = (v | (v << 4)) & 0x0f0f0f0fu;
  v = (v | (v << 2)) & 0x33333333u;
  v = (v | (v << 1)) & 0x55555555u;
  return v;
}
__device__ __forceinline__ float2 src_xy(const float2* xy, const float* xyz, int i) {
  return xy ? xy[i] : make_float2(xyz[3 * (size_t)i], xyz[3 * (size_t)i + 1]);
}

__global__ void __launch_bounds__(256) k_g_bbox(int n_src, const float2* __restrict__ xy, const float* __restrict__ xyz,
                                                const int32_t* __restrict__ src2ext, int* __restrict__ box /* x0 y0 x1 y1 */) {
  int x0 = INT_MAX, y0 = INT_MAX, x1 = INT_MIN, y1 = INT_MIN;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_src; i += gridDim.x * blockDim.x) {
    if (src2ext && src2ext[i] < 0) continue;
    const float2 p = src_xy(xy, xyz, i);
    const int a = float_order(p.x), b = float_order(p.y);
    x0 = min(x0, a); x1 = max(x1, a); y0 = min(y0, b); y1 = max(y1, b);
  }
  x0 = __reduce_min_sync(FULL, x0); y0 = __reduce_min_sync(FULL, y0);
  x1 = __reduce_max_sync(FULL, x1); y1 = __reduce_max_sync(FULL, y1);
  if ((threadIdx.x & 31) == 0 && x0 <= x1) {
    atomicMin(box + 0, x0); atomicMin(box + 1, y0); atomicMax(box + 2, x1); atomicMax(box + 3, y1);
  }
}

// key = 31-bit Morton code of the node's cell (kept nodes) or 0xffffffff (dropped: they sort behind)
__global__ void __launch_bounds__(256) k_g_keys(int n_src, const float2* __restrict__ xy, const float* __restrict__ xyz,
                                                const int32_t* __restrict__ src2ext, const int* __restrict__ box,
                                                uint32_t* __restrict__ key, uint32_t* __restrict__ val) {
  const float x0 = order_float(box[0]), y0 = order_float(box[1]);
  const float ext = fmaxf(fmaxf(order_float(box[2]) - x0, order_float(box[3]) - y0), 1e-6f);
  const float s = 65535.0f / ext;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_src; i += gridDim.x * blockDim.x) {
    uint32_t k = 0xffffffffu;
    if (!src2ext || src2ext[i] >= 0) {
      const float2 p = src_xy(xy, xyz, i);
      const uint32_t qx = (uint32_t)fminf(fmaxf((p.x - x0) * s, 0.f), 65535.f);
      const uint32_t qy = (uint32_t)fminf(fmaxf((p.y - y0) * s, 0.f), 65535.f);
      k = (spread16(qx) | (spread16(qy) << 1)) >> 1;
    }
    key[i] = k;
    val[i] = (uint32_t)i;
  }
}

__global__ void __launch_bounds__(256) k_g_number(int n, const uint32_t* __restrict__ perm, const long long* __restrict__ row,
                                                  const int32_t* __restrict__ src2ext, int32_t* __restrict__ src2int,
                                                  int32_t* __restrict__ ext2int, uint32_t* __restrict__ deg) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int s = (int)perm[i];
    src2int[s] = i;
    ext2int[src2ext ? src2ext[s] : s] = i;
    deg[i] = (uint32_t)(row[s + 1] - row[s]);
  }
}

// one 8-lane group per node: node record, edge records (destination renumbered, its position inlined)
__global__ void __launch_bounds__(256) k_g_fill(int n, const uint32_t* __restrict__ perm, const uint32_t* __restrict__ begin,
                                                const long long* __restrict__ row, const int32_t* __restrict__ col,
                                                const float* __restrict__ w, const float* __restrict__ dist,
                                                const float2* __restrict__ xy, const float* __restrict__ xyz,
                                                const signed char* __restrict__ state8, const int32_t* __restrict__ state32,
                                                const int32_t* __restrict__ src2ext, const int32_t* __restrict__ src2int,
                                                int4* __restrict__ node, float4* __restrict__ edge, float* __restrict__ ow,
                                                float* __restrict__ od, int32_t* __restrict__ ostate) {
  const int gl = threadIdx.x % kGroup;
  const int groups = gridDim.x * (blockDim.x / kGroup);
  for (int i = blockIdx.x * (blockDim.x / kGroup) + threadIdx.x / kGroup; i <= n; i += groups) {
    if (i == n) {  // sentinel: one past the last edge
      if (gl == 0) node[n] = make_int4((int)begin[n], -1, 0, 0);
      continue;
    }
    const int s = (int)perm[i];
    const uint32_t b = begin[i];
    if (gl == 0) {
      const float2 p = src_xy(xy, xyz, s);
      node[i] = make_int4((int)b, src2ext ? src2ext[s] : s, __float_as_int(p.x), __float_as_int(p.y));
      ostate[i] = state8 ? (int32_t)state8[s] : state32[s];
    }
    const long long j0 = row[s], j1 = row[s + 1];
    for (long long j = j0 + gl; j < j1; j += kGroup) {
      const int ds = col[j];
      int di = src2int[ds];
      const float2 q = src_xy(xy, xyz, ds);
      const size_t o = (size_t)b + (size_t)(j - j0);
      float dd = dist[j];
      if (di < 0) { di = i; dd = INFINITY; }  // destination outside the graph: a self loop nobody can afford
      edge[o] = make_float4(__int_as_float(di), INFINITY, q.x, q.y);
      ow[o] = w[j];
      od[o] = dd;
    }
  }
}

// a CSR that does not describe n rows over e columns would be read out of bounds: flag it
__global__ void __launch_bounds__(256) k_g_validate(int n, long long e, const long long* __restrict__ row,
                                                    const int32_t* __restrict__ col, int* __restrict__ bad) {
  const long long gt = blockIdx.x * (long long)blockDim.x + threadIdx.x, GT = (long long)gridDim.x * blockDim.x;
  if (gt == 0 && (row[0] != 0 || row[n] != e)) atomicOr(bad, 1);
  for (long long i = gt; i < n; i += GT)
    if (row[i] > row[i + 1]) atomicOr(bad, 2);
  for (long long j = gt; j < e; j += GT)
    if (col[j] < 0 || col[j] >= n) atomicOr(bad, 4);
}

__global__ void __launch_bounds__(256) k_g_sum(const float* __restrict__ dist, const int4* __restrict__ sentinel,
                                               double* __restrict__ partial) {
  __shared__ double s[256];
  double acc = 0.0;
  const long long e = sentinel->x;  // edges actually written
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < e; i += (long long)gridDim.x * blockDim.x) {
    const float d = dist[i];
    if (d < INFINITY) acc += (double)d;
  }
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = s[0];
}

// cost of every edge for the current safety factor (trg.cpp:674); +inf into Invalid nodes (trg.cpp:670)
__global__ void __launch_bounds__(256) k_edge_cost(const float* __restrict__ w, const float* __restrict__ dist,
                                                   const int32_t* __restrict__ state, const int4* __restrict__ sentinel,
                                                   float sf, float4* __restrict__ edge) {
  const int64_t e = sentinel->x;  // edges actually written
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < e; i += (int64_t)gridDim.x * blockDim.x) {
    const int dst = __float_as_int(edge[i].x);
    float c = __fmul_rn(__fadd_rn(__fmul_rn(sf, w[i]), 1.0f), dist[i]);
    if (state[dst] == -1) c = INFINITY;
    edge[i].y = c;
  }
}

// ------------------------------------------------------------------------------------------------------
// search
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float label_g(unsigned long long l) { return __uint_as_float((unsigned)(l >> 32)); }
__device__ __forceinline__ int label_parent(unsigned long long l) { return (int)(unsigned)(l & 0xffffffffull); }
__device__ __forceinline__ unsigned long long make_label(float g, int parent) {
  return ((unsigned long long)__float_as_uint(g) << 32) | (unsigned)parent;
}
__device__ __forceinline__ float dist2d(float px, float py, float2 goal) {
  const float dx = __fsub_rn(goal.x, px), dy = __fsub_rn(goal.y, py);
  return __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
}
__device__ __forceinline__ float heur(const int4* __restrict__ node, int v, float2 goal) {
  const int4 r = __ldg(node + v);
  return dist2d(__int_as_float(r.z), __int_as_float(r.w), goal);
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
  int32_t* path_ids;   // caller's ids, start..goal
  long long capacity;
  // [0] path write cursor, [1] next query, [2] edges relaxed (20 B each),
  // [3] near passes, [4] threshold steps, [5] far entries rescanned, [6] nodes expanded, [7] pops pruned by the goal bound
  unsigned long long* cursor;
};

// queries in the order they are served: longest straight line first (the longest search bounds the batch)
__global__ void __launch_bounds__(256) k_sssp_order(const int4* __restrict__ node, const int32_t* __restrict__ ext2int,
                                                    const int32_t* __restrict__ starts, const int32_t* __restrict__ goals, int nq,
                                                    uint32_t* __restrict__ key, uint32_t* __restrict__ val) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += gridDim.x * blockDim.x) {
    const int s = ext2int[starts[i]], t = ext2int[goals[i]];
    uint32_t k = 0xffffffffu;
    if (s >= 0 && t >= 0) {
      const int4 a = node[s], b = node[t];
      k = ~__float_as_uint(dist2d(__int_as_float(a.z), __int_as_float(a.w), make_float2(__int_as_float(b.z), __int_as_float(b.w))));
    }
    key[i] = k;
    val[i] = (uint32_t)i;
  }
}

__global__ void __launch_bounds__(kSsspThreads, 1024 / kSsspThreads) k_sssp(
    int n, const int4* __restrict__ node, const float4* __restrict__ edge, const float* __restrict__ ew,
    const float* __restrict__ ed, const int32_t* __restrict__ ext2int, const int32_t* __restrict__ starts,
    const int32_t* __restrict__ goals, const uint32_t* __restrict__ order, int64_t nq, float delta,
    unsigned long long* __restrict__ labels, int32_t* __restrict__ queues, uint32_t* __restrict__ bitmaps, SsspOut out) {
  const int words = (n + 31) >> 5;
  unsigned long long* label = labels + (size_t)blockIdx.x * n;
  int32_t* q0 = queues + (size_t)blockIdx.x * 4 * n;
  uint32_t* bits0 = bitmaps + (size_t)blockIdx.x * 3 * words;

  __shared__ int s_near[3];       // size of the near queue of pass p at [p % 3]
  __shared__ int s_far[2];        // far pile: current, next
  __shared__ long long s_query;
  __shared__ float s_thr, s_best;
  __shared__ int s_done;
  __shared__ unsigned s_minf;
  __shared__ int s_plen;
  __shared__ long long s_poff;
  __shared__ float s_stage[2 * kSsspThreads];
  __shared__ unsigned long long s_relax;
  __shared__ unsigned long long s_dbg[5];  // near passes, threshold steps, far entries rescanned, nodes expanded, pops pruned

  const int tid = threadIdx.x;
  const int grp = tid / kGroup, gl = tid % kGroup;
  const int ngrp = kSsspThreads / kGroup;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_query = (long long)atomicAdd(out.cursor + 1, 1ull);
    __syncthreads();
    if (s_query >= nq) return;
    const long long qi = (long long)order[s_query];
    const int start = __ldg(ext2int + starts[qi]), goal = __ldg(ext2int + goals[qi]);
    if (start < 0 || goal < 0) {  // an id that is not part of the graph: no path
      if (tid == 0) {
        out.found[qi] = 0; out.cost[qi] = 0.f; out.path_length[qi] = 0.f; out.avg_risk[qi] = 0.f;
        out.path_off[qi] = 0; out.path_len[qi] = 0;
      }
      continue;
    }
    float2 gpos;
    {
      const int4 r = __ldg(node + goal);
      gpos = make_float2(__int_as_float(r.z), __int_as_float(r.w));
    }
    for (int i = tid; i < n; i += kSsspThreads) label[i] = kInfLabel;
    for (int i = tid; i < 3 * words; i += kSsspThreads) bits0[i] = 0u;
    int32_t *qcur = q0, *qnxt = q0 + n, *qfar = q0 + 2 * (size_t)n, *qfar2 = q0 + 3 * (size_t)n;
    uint32_t *cur_bits = bits0, *nxt_bits = bits0 + words, *far_bits = bits0 + 2 * (size_t)words;
    __syncthreads();
    if (tid == 0) {
      s_relax = 0ull;
      for (int j = 0; j < 5; ++j) s_dbg[j] = 0ull;
      label[start] = make_label(0.f, __ldg(node + start).y);
      qcur[0] = start;
      cur_bits[start >> 5] = 1u << (start & 31);
      s_near[0] = 1; s_near[1] = 0; s_near[2] = 0;
      s_far[0] = 0; s_far[1] = 0;
      s_thr = __fadd_rn(heur(node, start, gpos), delta);
      s_best = INFINITY;
      s_done = 0;
    }
    __syncthreads();
    int pass = 0;  // s_near[pass % 3] = entries of qcur; s_near[(pass + 1) % 3] == 0

    for (;;) {
      // ---- relax the near pile to a fixed point under the current threshold. A queued node owns a bit in
      // the bitmap of its queue; it drops the bit when it is taken up, and whoever improves it afterwards
      // (even during the same pass) queues it again in the other bitmap — no lost updates, and one
      // barrier per pass: the counter of the pass after next is zeroed while nobody looks at it.
      for (;;) {
        const int ncur = s_near[pass % 3];
        if (ncur == 0) break;
        int* nxt_cnt = &s_near[(pass + 1) % 3];
        if (tid == 0) { s_near[(pass + 2) % 3] = 0; s_dbg[0] += 1; }
        const float thr = s_thr;
        const float best = label_g(__ldcg(label + goal));
        unsigned my_relax = 0;
        for (int k = grp; k < ncur; k += ngrp) {
          const int u = qcur[k];
          if (gl == 0) clear_bit(cur_bits, u);
          const float gu = label_g(__ldcg(label + u));
          const int4 ru = __ldg(node + u);
          const int e1 = __ldg(&node[u + 1].x);
          // goal bound: with a consistent heuristic no path through u beats `best`
          if (__fadd_rn(gu, dist2d(__int_as_float(ru.z), __int_as_float(ru.w), gpos)) > best * kGoalSlack) {
            if (gl == 0) atomicAdd(&s_dbg[4], 1ull);
            continue;
          }
          if (gl == 0) atomicAdd(&s_dbg[3], 1ull);
          for (int e = ru.x + gl; e < e1; e += kGroup) {
            const float4 er = __ldg(edge + e);
            if (!(er.y < INFINITY)) continue;  // trg.cpp:670 skip Invalid dst
            const int v = __float_as_int(er.x);
            ++my_relax;
            const float ng = __fadd_rn(gu, er.y);
            const unsigned long long cand = make_label(ng, ru.y);
            const unsigned long long old = atomicMin(label + v, cand);
            if (cand < old && ng < label_g(old)) {
              const float f = __fadd_rn(ng, dist2d(er.z, er.w, gpos));
              if (f < thr) {
                if (set_bit(nxt_bits, v)) qnxt[atomicAdd(nxt_cnt, 1)] = v;
              } else {
                if (set_bit(far_bits, v)) qfar[atomicAdd(&s_far[0], 1)] = v;
              }
            }
          }
        }
        my_relax = __reduce_add_sync(FULL, my_relax);
        if ((tid & 31) == 0 && my_relax) atomicAdd(&s_relax, (unsigned long long)my_relax);
        __syncthreads();
        { int32_t* t = qcur; qcur = qnxt; qnxt = t; }
        { uint32_t* t = cur_bits; cur_bits = nxt_bits; nxt_bits = t; }
        ++pass;
      }
      // ---- near pile empty: everything with f < thr is final
      __syncthreads();
      if (tid == 0) {
        s_best = label_g(__ldcg(label + goal));
        if (s_best < s_thr || s_far[0] == 0) s_done = 1;
        s_minf = 0x7f800000u;
      }
      __syncthreads();
      if (s_done) break;
      // ---- advance the threshold past the smallest f waiting in the far pile
      const int nfar = s_far[0];
      const float best = s_best;
      if (tid == 0) { s_dbg[1] += 1; s_dbg[2] += (unsigned long long)nfar; }
      float myf = INFINITY;
      for (int k = tid; k < nfar; k += kSsspThreads) {
        const int v = qfar[k];
        myf = fminf(myf, __fadd_rn(label_g(__ldcg(label + v)), heur(node, v, gpos)));
      }
      atomicMin(&s_minf, __float_as_uint(myf));  // non-negative floats order like uints
      __syncthreads();
      if (tid == 0) s_thr = __fadd_rn(fmaxf(s_thr, __uint_as_float(s_minf)), delta);
      __syncthreads();
      const float thr = s_thr;
      int* cur_cnt = &s_near[pass % 3];  // zero here (the near loop ended on it)
      for (int k = tid; k < nfar; k += kSsspThreads) {
        const int v = qfar[k];
        const float f = __fadd_rn(label_g(__ldcg(label + v)), heur(node, v, gpos));
        if (f > best * kGoalSlack) { clear_bit(far_bits, v); continue; }  // can never matter
        if (f < thr) {
          clear_bit(far_bits, v);
          if (set_bit(cur_bits, v)) qcur[atomicAdd(cur_cnt, 1)] = v;
        } else {
          qfar2[atomicAdd(&s_far[1], 1)] = v;  // stays far (bit remains set)
        }
      }
      __syncthreads();
      if (tid == 0) { s_far[0] = s_far[1]; s_far[1] = 0; }
      { int32_t* t = qfar; qfar = qfar2; qfar2 = t; }
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
          v = __ldg(ext2int + label_parent(__ldcg(label + v)));
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
      const int e0 = node[v].x, e1 = node[v + 1].x;
      for (int e = e0; e < e1; ++e)
        if (__float_as_int(edge[e].x) == u) { dd = ed[e]; ww = ew[e]; break; }
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
      for (int k = tid; k < plen; k += kSsspThreads) out.path_ids[poff + k] = node[rev[plen - 1 - k]].y;
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

// The search scratch (labels, queues, bitmaps: ~24 bytes per node and slot, several GB for a
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
  gfree(g->d_node, st); gfree(g->d_edge, st); gfree(g->d_w, st); gfree(g->d_dist, st); gfree(g->d_state, st);
  gfree(g->d_ext2int, st); gfree(g->d_partial, st);
  if (st) cudaStreamSynchronize(st);
  release_scratch(g);
  if (g->stream) {
    cudaStreamSynchronize(g->stream);
    cudaStreamDestroy(g->stream);
  }
  delete g;
}

// Builds the search graph from CSR arrays resident on the device (source numbering 0 .. n_src-1; src2ext maps
// a source node to the caller's id or -1 = not part of the graph, nullptr = identity). All work is enqueued on
// `bs` (the stream that produced / owns the source arrays); the handle's own stream waits for it.
int trgb::graph_from_device(trgb_graph** out, const GraphSource& s, cudaStream_t bs) {
  TRGB_ARG(out && s.n_src > 0 && s.n_keep > 0 && s.n_keep <= s.n_src && s.n_ext >= s.n_keep && s.row && (s.xy || s.xyz) &&
               (s.state8 || s.state32),
           "empty graph");
  TRGB_ARG(s.e_src == 0 || (s.col && s.w && s.dist), "null edge arrays");
  TRGB_ARG(s.e_src < (1ll << 31), "more than 2^31 directed edges");
  trgb::tune_mempool_once();
  trgb_graph* g = new trgb_graph();
  {
    cudaError_t es = cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking);
    if (es != cudaSuccess) { delete g; return cuda_fail(es, "cudaStreamCreate", __FILE__, __LINE__); }
  }
  g->n_ext = s.n_ext;
  g->n = s.n_keep;
  g->e = s.e_src;
  cudaGetDevice(&g->device);
  const size_t n = (size_t)g->n, e = (size_t)std::max<int64_t>(1, g->e), ns = (size_t)s.n_src;
  uint32_t *key = nullptr, *key2 = nullptr, *val = nullptr, *perm = nullptr, *deg = nullptr, *begin = nullptr;
  int32_t* src2int = nullptr;
  int* box = nullptr;
  cudaError_t er = cudaSuccess;
  auto A = [&](void** p, size_t bytes) { if (er == cudaSuccess) er = cudaMallocAsync(p, bytes ? bytes : 1, bs); };
  A((void**)&g->d_node, (n + 1) * sizeof(int4)); A((void**)&g->d_edge, e * sizeof(float4));
  A((void**)&g->d_w, e * sizeof(float)); A((void**)&g->d_dist, e * sizeof(float));
  A((void**)&g->d_state, n * sizeof(int32_t)); A((void**)&g->d_ext2int, (size_t)g->n_ext * sizeof(int32_t));
  A((void**)&g->d_partial, kPartials * sizeof(double));
  A((void**)&key, ns * 4); A((void**)&key2, ns * 4); A((void**)&val, ns * 4); A((void**)&perm, ns * 4);
  A((void**)&deg, (n + 1) * 4); A((void**)&begin, (n + 1) * 4); A((void**)&src2int, ns * 4); A((void**)&box, 4 * sizeof(int));
  if (er != cudaSuccess) { trgb_graph_destroy(g); return cuda_fail(er, "graph alloc", __FILE__, __LINE__); }
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(((int64_t)ns + 255) / 256, (int64_t)sm_count() * 8));
  const int box0[4] = {INT_MAX, INT_MAX, INT_MIN, INT_MIN};
  cudaMemcpyAsync(box, box0, sizeof(box0), cudaMemcpyHostToDevice, bs);
  cudaMemsetAsync(src2int, 0xff, ns * 4, bs);
  cudaMemsetAsync(g->d_ext2int, 0xff, (size_t)g->n_ext * sizeof(int32_t), bs);
  cudaMemsetAsync(deg + n, 0, 4, bs);
  int rc = TRGB_OK;
  {
    ProfScope ps("k_graph_build", bs, (double)g->e);
    k_g_bbox<<<grid, 256, 0, bs>>>(s.n_src, s.xy, s.xyz, s.src2ext, box);
    k_g_keys<<<grid, 256, 0, bs>>>(s.n_src, s.xy, s.xyz, s.src2ext, box, key, val);
    rc = sort_pairs_u32_u32(key, key2, val, perm, s.n_src, 32, bs);
    if (!rc) {
      k_g_number<<<grid, 256, 0, bs>>>(g->n, perm, s.row, s.src2ext, src2int, g->d_ext2int, deg);
      rc = exclusive_sum_u32(deg, begin, g->n + 1, bs);
    }
    if (!rc) {
      const int fgrid = (int)std::max<int64_t>(1, std::min<int64_t>(((int64_t)n + 1 + 31) / 32, (int64_t)sm_count() * 16));
      k_g_fill<<<fgrid, 256, 0, bs>>>(g->n, perm, begin, s.row, s.col, s.w, s.dist, s.xy, s.xyz, s.state8, s.state32, s.src2ext,
                                      src2int, g->d_node, g->d_edge, g->d_w, g->d_dist, g->d_state);
      k_g_sum<<<kPartials, 256, 0, bs>>>(g->d_dist, g->d_node + g->n, g->d_partial);
    }
  }
  if (!rc) er = cudaGetLastError();
  for (void* p : {(void*)key, (void*)key2, (void*)val, (void*)perm, (void*)deg, (void*)begin, (void*)src2int, (void*)box})
    cudaFreeAsync(p, bs);
  cudaEvent_t ev = nullptr;
  if (!rc && er == cudaSuccess) er = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  if (!rc && er == cudaSuccess) er = cudaEventRecord(ev, bs);
  if (!rc && er == cudaSuccess) er = cudaStreamWaitEvent(g->stream, ev, 0);
  if (ev) cudaEventDestroy(ev);
  if (rc || er != cudaSuccess) {
    cudaStreamSynchronize(bs);
    trgb_graph_destroy(g);
    return rc ? rc : cuda_fail(er, "graph build", __FILE__, __LINE__);
  }
  *out = g;
  return TRGB_OK;
}

extern "C" int trgb_graph_upload(trgb_graph** out, const TrgbGraphDesc* d) {
  TRGB_ARG(out && d, "null pointer");
  TRGB_ARG(d->n_nodes > 0 && d->row_ptr && d->pos_xyz && d->state, "empty graph");
  TRGB_ARG(d->n_edges >= 0 && (d->n_edges == 0 || (d->col && d->weight && d->dist)), "null edge arrays");
  // a CSR that does not describe n_nodes rows over n_edges columns would be read out of bounds on the device
  TRGB_ARG(d->row_ptr[0] == 0 && d->row_ptr[d->n_nodes] == d->n_edges, "row_ptr does not span the edge arrays");
  for (int32_t i = 0; i < d->n_nodes; ++i) TRGB_ARG(d->row_ptr[i] <= d->row_ptr[i + 1], "row_ptr is not monotonic");
  for (int64_t j = 0; j < d->n_edges; ++j) TRGB_ARG(d->col[j] >= 0 && d->col[j] < d->n_nodes, "edge to a node id outside the graph");
  trgb::tune_mempool_once();
  cudaStream_t st = nullptr;
  TRGB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  const size_t n = (size_t)d->n_nodes, e = (size_t)d->n_edges;
  long long* row = nullptr; int32_t* col = nullptr; float *w = nullptr, *dist = nullptr, *xyz = nullptr; int32_t* state = nullptr;
  cudaError_t er = cudaSuccess;
  auto UP = [&](void** dst, const void* src, size_t bytes) {
    if (er == cudaSuccess) er = cudaMallocAsync(dst, bytes ? bytes : 1, st);
    if (er == cudaSuccess && bytes) er = cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, st);
  };
  UP((void**)&row, d->row_ptr, (n + 1) * sizeof(int64_t));
  UP((void**)&col, d->col, e * sizeof(int32_t));
  UP((void**)&w, d->weight, e * sizeof(float));
  UP((void**)&dist, d->dist, e * sizeof(float));
  UP((void**)&xyz, d->pos_xyz, 3 * n * sizeof(float));
  UP((void**)&state, d->state, n * sizeof(int32_t));
  int rc = TRGB_OK;
  if (er != cudaSuccess) {
    rc = cuda_fail(er, "graph upload", __FILE__, __LINE__);
  } else {
    GraphSource s;
    s.n_src = d->n_nodes; s.n_keep = d->n_nodes; s.n_ext = d->n_nodes; s.e_src = d->n_edges;
    s.row = row; s.col = col; s.w = w; s.dist = dist; s.xyz = xyz; s.state32 = state;
    rc = graph_from_device(out, s, st);
  }
  for (void* p : {(void*)row, (void*)col, (void*)w, (void*)dist, (void*)xyz, (void*)state})
    if (p) cudaFreeAsync(p, st);
  cudaStreamSynchronize(st);  // the host arrays may go away after return
  cudaStreamDestroy(st);
  return rc;
}

// Same as trgb_graph_upload with every array of the descriptor already resident on the device (e.g. a graph
// merged from several tiles by torch ops): nothing crosses PCIe. `stream` = the stream the arrays were produced on.
extern "C" int trgb_graph_upload_device(trgb_graph** out, const TrgbGraphDesc* d, void* stream) {
  TRGB_ARG(out && d, "null pointer");
  TRGB_ARG(d->n_nodes > 0 && d->row_ptr && d->pos_xyz && d->state, "empty graph");
  TRGB_ARG(d->n_edges >= 0 && (d->n_edges == 0 || (d->col && d->weight && d->dist)), "null edge arrays");
  trgb::tune_mempool_once();
  cudaStream_t st = (cudaStream_t)stream;
  int* d_bad = nullptr;
  TRGB_CUDA(cudaMallocAsync((void**)&d_bad, sizeof(int), st));
  TRGB_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
  k_g_validate<<<sm_count() * 4, 256, 0, st>>>(d->n_nodes, (long long)d->n_edges, (const long long*)d->row_ptr, d->col, d_bad);
  int bad = 0;
  TRGB_CUDA(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  TRGB_CUDA(cudaStreamSynchronize(st));
  cudaFreeAsync(d_bad, st);
  TRGB_ARG(!(bad & 1), "row_ptr does not span the edge arrays");
  TRGB_ARG(!(bad & 2), "row_ptr is not monotonic");
  TRGB_ARG(!(bad & 4), "edge to a node id outside the graph");
  GraphSource s;
  s.n_src = d->n_nodes; s.n_keep = d->n_nodes; s.n_ext = d->n_nodes; s.e_src = d->n_edges;
  s.row = (const long long*)d->row_ptr; s.col = d->col; s.w = d->weight; s.dist = d->dist; s.xyz = d->pos_xyz; s.state32 = d->state;
  const int rc = graph_from_device(out, s, st);
  if (!rc) TRGB_CUDA(cudaStreamSynchronize(st));  // the caller's arrays may go away after return
  return rc;
}

static int ensure_slots(trgb_graph* g, int want) {
  if (g->nslots >= want) return TRGB_OK;
  const size_t n = g->n, words = (n + 31) / 32;
  const size_t need_l = (size_t)want * n * sizeof(unsigned long long);
  const size_t need_q = (size_t)want * 4 * n * sizeof(int32_t);
  const size_t need_b = (size_t)want * 3 * words * sizeof(uint32_t);
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
  TRGB_ARG(nq < (1ll << 31), "too many queries in one batch");
  for (int64_t i = 0; i < nq; ++i)
    TRGB_ARG(start_ids[i] >= 0 && start_ids[i] < g->n_ext && goal_ids[i] >= 0 && goal_ids[i] < g->n_ext, "node id out of range");
  cudaStream_t st = g->stream;
  if (g->mean_cost < 0.f) {  // mean edge length, summed on the device when the graph was built
    double part[kPartials];
    TRGB_CUDA(cudaMemcpyAsync(part, g->d_partial, sizeof(part), cudaMemcpyDeviceToHost, st));
    TRGB_CUDA(cudaStreamSynchronize(st));
    double sum = 0;
    for (double p : part) sum += p;
    g->mean_cost = g->e ? (float)(sum / (double)g->e) : 1.f;
  }
  if (!(g->cost_sf == safety_factor)) {
    ProfScope ps("k_edge_cost", st, (double)g->e);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((g->e + 255) / 256, (int64_t)sm_count() * 8));
    k_edge_cost<<<grid, 256, 0, st>>>(g->d_w, g->d_dist, g->d_state, g->d_node + g->n, safety_factor, g->d_edge);
    g->cost_sf = safety_factor;
  }
  // slots: enough CTAs to fill the machine, bounded by the batch and by ~48 GB of scratch
  const size_t per_slot = (size_t)g->n * (8 + 16) + 3 * ((size_t)g->n / 8 + 8);
  int want = (int)std::min<int64_t>(nq, (int64_t)sm_count() * std::max(1, 1024 / kSsspThreads));
  want = (int)std::min<size_t>((size_t)want, std::max<size_t>(1, ((size_t)48 << 30) / per_slot));
  int rc = ensure_slots(g, want);
  if (rc) return rc;

  struct Dev { void* p = nullptr; cudaStream_t s = nullptr; ~Dev() { if (p) cudaFreeAsync(p, s); } };
  Dev d_s, d_g, d_found, d_cost, d_len, d_risk, d_off, d_plen, d_ids, d_cur, d_k1, d_k2, d_v1, d_ord;
#define DALLOC(b, bytes) do { (b).s = st; TRGB_CUDA(cudaMallocAsync(&(b).p, (bytes) ? (bytes) : 1, st)); } while (0)
  DALLOC(d_s, nq * sizeof(int32_t)); DALLOC(d_g, nq * sizeof(int32_t));
  DALLOC(d_found, nq); DALLOC(d_cost, nq * sizeof(float)); DALLOC(d_len, nq * sizeof(float));
  DALLOC(d_risk, nq * sizeof(float)); DALLOC(d_off, nq * sizeof(int64_t)); DALLOC(d_plen, nq * sizeof(int32_t));
  DALLOC(d_ids, (size_t)path_ids_capacity * sizeof(int32_t)); DALLOC(d_cur, 8 * sizeof(unsigned long long));
  DALLOC(d_k1, nq * 4); DALLOC(d_k2, nq * 4); DALLOC(d_v1, nq * 4); DALLOC(d_ord, nq * 4);
#undef DALLOC
  TRGB_CUDA(cudaMemcpyAsync(d_s.p, start_ids, nq * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  TRGB_CUDA(cudaMemcpyAsync(d_g.p, goal_ids, nq * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  TRGB_CUDA(cudaMemsetAsync(d_cur.p, 0, 8 * sizeof(unsigned long long), st));
  SsspOut o;
  o.found = (uint8_t*)d_found.p; o.cost = (float*)d_cost.p; o.path_length = (float*)d_len.p;
  o.avg_risk = (float*)d_risk.p; o.path_off = (int64_t*)d_off.p; o.path_len = (int32_t*)d_plen.p;
  o.path_ids = (int32_t*)d_ids.p; o.capacity = path_ids_capacity; o.cursor = (unsigned long long*)d_cur.p;
  // threshold step: 0.75..6 x the mean edge length measured at C2 (scripts/sssp_sweep.py): flat between 0.75 and 1.5
  float delta = 1.0f * std::max(g->mean_cost, 1e-3f);
  if (const char* ev = std::getenv("TRGB_SSSP_DELTA")) delta = (float)std::atof(ev) * std::max(g->mean_cost, 1e-3f);
  {
    ProfScope ps("k_sssp", st, (double)nq);
    k_sssp_order<<<(int)std::min<int64_t>((nq + 255) / 256, 1024), 256, 0, st>>>(
        g->d_node, g->d_ext2int, (const int32_t*)d_s.p, (const int32_t*)d_g.p, (int)nq, (uint32_t*)d_k1.p, (uint32_t*)d_v1.p);
    rc = sort_pairs_u32_u32((const uint32_t*)d_k1.p, (uint32_t*)d_k2.p, (const uint32_t*)d_v1.p, (uint32_t*)d_ord.p, (int)nq, 32, st);
    if (rc) return rc;
    k_sssp<<<g->nslots < want ? g->nslots : want, kSsspThreads, 0, st>>>(
        g->n, g->d_node, g->d_edge, g->d_w, g->d_dist, g->d_ext2int, (const int32_t*)d_s.p, (const int32_t*)d_g.p,
        (const uint32_t*)d_ord.p, nq, delta, g->d_label, g->d_queue, g->d_bits, o);
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
    fprintf(stderr, "[k_sssp] nq=%lld relaxed=%llu passes=%llu thr_steps=%llu far_scanned=%llu expanded=%llu pruned=%llu delta=%g\n",
            (long long)nq, cursor[2], cursor[3], cursor[4], cursor[5], cursor[6], cursor[7], (double)delta);
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
