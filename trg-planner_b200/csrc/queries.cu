// K2 / K3 / K4 and the sampling-window kernel: batched, pure functions of
// (query, static map index, params). One warp per query; the cells a query touches are
// contiguous float4 runs (one per grid row), loaded coalesced through the read-only path.
//
//   k_collision      TRG::isCollision                      trg.cpp:746-778 (+ kdtree.c:270-301)
//   k_range_count    kd_nearest_range2 result size         kdtree.c:479-501
//   k_sample_window  sampling loop of TRG::expandGraph     trg.cpp:387-403 (collision bits only)
//   k_nearest_z      kd_nearest2 on the map tree + z       trg.cpp:244-246 (+ kdtree.c:303-417)
//   k_edge_eval      geometric part of TRG::wireEdge       trg.cpp:276-363
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"

namespace trgb {

// ------------------------------------------------------------------------------------------
// K2
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_collision(MapView m, const float2* __restrict__ q,
                                                        int64_t n, float r, float hthr, float rthr,
                                                        int cap, uint8_t* __restrict__ out) {
  extern __shared__ float zsm[];
  float* zbuf = zsm + (threadIdx.x >> 5) * (cap + 256);  // + 256 bins for the histogram select
  const int lane = threadIdx.x & 31;
  const int64_t wstride = (int64_t)gridDim.x * kWarpsPerCta;
  for (int64_t i = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5); i < n; i += wstride) {
    const float2 p = __ldg(q + i);
    const bool c = warp_is_collision(m, p.x, p.y, r, hthr, rthr, zbuf, cap, nullptr);
    if (lane == 0) out[i] = c ? 1 : 0;
  }
}

__global__ void __launch_bounds__(kThreads) k_range_count(MapView m, const float2* __restrict__ q,
                                                          int64_t n, float r, int32_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const float r2 = __fmul_rn(r, r);
  const int64_t wstride = (int64_t)gridDim.x * kWarpsPerCta;
  for (int64_t i = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5); i < n; i += wstride) {
    const float2 p = __ldg(q + i);
    int cnt = 0;
    warp_for_each_candidate(m, p.x, p.y, inflate(r, p.x, p.y), [&](bool valid, const float4& c) {
      const float dx = __fsub_rn(c.x, p.x), dy = __fsub_rn(c.y, p.y);
      const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
      cnt += (valid && d2 <= r2);
    });
    cnt = __reduce_add_sync(FULL, cnt);
    if (lane == 0) out[i] = cnt;
  }
}

// one warp per (node, draw) pair; bit j of mask[node] = isCollision(node + draw[first+j])
__global__ void __launch_bounds__(kThreads) k_sample_window(
    MapView m, const float2* __restrict__ node_xy, const int32_t* __restrict__ first_draw,
    const float2* __restrict__ draw_xy, int64_t n_nodes, int window, float r, float hthr, float rthr,
    int cap, unsigned long long* __restrict__ mask) {
  extern __shared__ float zsm[];
  float* zbuf = zsm + (threadIdx.x >> 5) * (cap + 256);  // + 256 bins for the histogram select
  const int lane = threadIdx.x & 31;
  const int64_t items = n_nodes * window;
  const int64_t wstride = (int64_t)gridDim.x * kWarpsPerCta;
  for (int64_t it = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5); it < items; it += wstride) {
    const int64_t node = it / window;
    const int j = (int)(it - node * window);
    const float2 np = __ldg(node_xy + node);
    const float2 d = __ldg(draw_xy + (__ldg(first_draw + node) + j));
    // sample = node->pos_.head(2) + Vector2f(e*cos, e*sin)   (trg.cpp:396-397); d = (e*cos, e*sin)
    const float sx = __fadd_rn(np.x, d.x), sy = __fadd_rn(np.y, d.y);
    const bool c = warp_is_collision(m, sx, sy, r, hthr, rthr, zbuf, cap, nullptr);
    if (lane == 0 && c) atomicOr(mask + node * ((window + 63) >> 6) + (j >> 6), 1ull << (j & 63));
  }
}

// ------------------------------------------------------------------------------------------
// K3 — nearest map point in 2-D (strict `<` on float dist^2; ties -> lowest original index,
// flagged). Ring search over cell blocks until the best distance is provably final.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_nearest_z(MapView m, const float2* __restrict__ q,
                                                        int64_t n, float* __restrict__ z_out,
                                                        int64_t* __restrict__ idx_out,
                                                        uint8_t* __restrict__ tie_out,
                                                        const float* __restrict__ skip_d2, float skip_below) {
  const int lane = threadIdx.x & 31;
  const int64_t wstride = (int64_t)gridDim.x * kWarpsPerCta;
  const float extent = (float)m.W * m.cell + (float)m.H * m.cell;
  for (int64_t i = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5); i < n; i += wstride) {
    if (skip_d2 && __fsqrt_rn(__ldg(skip_d2 + i)) < skip_below) {  // speculation filter (warp-uniform)
      if (lane == 0) {
        if (z_out) z_out[i] = 0.f;
        if (idx_out) idx_out[i] = -1;
        if (tie_out) tie_out[i] = 0;
      }
      continue;
    }
    const float2 p = __ldg(q + i);
    const int qcx = cell_coord(p.x, m.x0, m.inv_cell, m.W);
    const int qcy = cell_coord(p.y, m.y0, m.inv_cell, m.H);
    const float fuzz = 2e-6f * (fabsf(p.x) + fabsf(p.y) + extent);
    float best = INFINITY, bz = 0.f;
    int bidx = 0x7fffffff;
    int tie = 0;
    for (int R = 1;; R *= 2) {
      const int bx0 = max(qcx - R, 0), bx1 = min(qcx + R, m.W - 1);
      const int by0 = max(qcy - R, 0), by1 = min(qcy + R, m.H - 1);
      best = INFINITY; bidx = 0x7fffffff; tie = 0;
      warp_for_each_in_cells(m, bx0, bx1, by0, by1, [&](bool valid, const float4& c) {
        if (!valid) return;
        const float dx = __fsub_rn(c.x, p.x), dy = __fsub_rn(c.y, p.y);
        const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        const int ci = __float_as_int(c.w);
        if (d2 < best) { best = d2; bidx = ci; bz = c.z; tie = 0; }
        else if (d2 == best) { tie = 1; if (ci < bidx) { bidx = ci; bz = c.z; } }
      });
      // warp arg-min (d2, idx) with tie tracking
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        const float ob = __shfl_xor_sync(FULL, best, d);
        const int oi = __shfl_xor_sync(FULL, bidx, d);
        const float oz = __shfl_xor_sync(FULL, bz, d);
        const int ot = __shfl_xor_sync(FULL, tie, d);
        if (ob < best) { best = ob; bidx = oi; bz = oz; tie = ot; }
        else if (ob == best && ob != INFINITY) {
          tie = (oi != bidx) ? 1 : (tie | ot);
          if (oi < bidx) { bidx = oi; bz = oz; }
        }
      }
      const bool whole = bx0 == 0 && by0 == 0 && bx1 == m.W - 1 && by1 == m.H - 1;
      if (whole) break;
      // distance from the query to the nearest side of the searched block that has cells beyond it
      float g = INFINITY;
      if (bx0 > 0) g = fminf(g, p.x - (m.x0 + (float)bx0 * m.cell));
      if (bx1 < m.W - 1) g = fminf(g, (m.x0 + (float)(bx1 + 1) * m.cell) - p.x);
      if (by0 > 0) g = fminf(g, p.y - (m.y0 + (float)by0 * m.cell));
      if (by1 < m.H - 1) g = fminf(g, (m.y0 + (float)(by1 + 1) * m.cell) - p.y);
      g = g * 0.9999f - fuzz;
      if (best < INFINITY && g > 0.f && best <= g * g) break;
    }
    if (lane == 0) {
      if (z_out) z_out[i] = bz;
      if (idx_out) idx_out[i] = best < INFINITY ? (int64_t)bidx : -1;
      if (tie_out) tie_out[i] = (uint8_t)tie;
    }
  }
}

// ------------------------------------------------------------------------------------------
// K4 — edge evaluation
// ------------------------------------------------------------------------------------------
// One-sided Jacobi SVD of a real 3x3 (row-major), the algorithm Eigen::JacobiSVD runs for a
// square matrix (trg.cpp:339): 2x2 real SVD steps over the pairs (1,0),(2,0),(2,1), left
// rotations accumulated into U, negative diagonals flipped, columns sorted descending.
struct Rot { float c, s; };
__device__ __forceinline__ Rot rot_mul(Rot a, Rot b) { return {a.c * b.c - a.s * b.s, a.c * b.s + a.s * b.c}; }
__device__ __forceinline__ Rot make_jacobi(float x, float y, float z) {
  const float deno = 2.f * fabsf(y);
  if (deno < 1.17549435e-38f) return {1.f, 0.f};
  const float tau = __fdiv_rn(x - z, deno);
  const float w = __fsqrt_rn(tau * tau + 1.f);
  const float t = tau > 0.f ? __fdiv_rn(1.f, tau + w) : __fdiv_rn(1.f, tau - w);
  const float sign_t = t > 0.f ? 1.f : -1.f;
  const float nn = __fdiv_rn(1.f, __fsqrt_rn(t * t + 1.f));
  Rot r;
  r.s = -sign_t * __fdiv_rn(y, fabsf(y)) * fabsf(t) * nn;
  r.c = nn;
  return r;
}
// One (p, q) step of the sweep with compile-time indices: W and U stay in registers (no local
// memory, no dynamic indexing). Same arithmetic, in the same order, as the loop form it replaces.
template <int P, int Q>
__device__ __forceinline__ void jacobi_pair(float (&W)[9], float (&U)[9], float& maxDiag, bool& finished) {
  const float precision = 2.f * 1.1920929e-07f;
  const float tiny = 1.17549435e-38f;
  const float thr = fmaxf(tiny, precision * maxDiag);
  if (!(fabsf(W[P * 3 + Q]) > thr || fabsf(W[Q * 3 + P]) > thr)) return;
  finished = false;
  float m0 = W[P * 3 + P], m1 = W[P * 3 + Q], m2 = W[Q * 3 + P], m3 = W[Q * 3 + Q];
  Rot rot1;
  const float t = m0 + m3, d = m2 - m1;
  if (fabsf(d) < tiny) { rot1.s = 0.f; rot1.c = 1.f; }
  else {
    const float u = __fdiv_rn(t, d);
    const float tmp = __fsqrt_rn(1.f + u * u);
    rot1.s = __fdiv_rn(1.f, tmp);
    rot1.c = __fdiv_rn(u, tmp);
  }
  if (!(rot1.c == 1.f && rot1.s == 0.f)) {  // rot_plane(mm+0, mm+2, n=2, rot1): rows (m0 m1) and (m2 m3)
    const float a0 = rot1.c * m0 + rot1.s * m2, b0 = -rot1.s * m0 + rot1.c * m2;
    const float a1 = rot1.c * m1 + rot1.s * m3, b1 = -rot1.s * m1 + rot1.c * m3;
    m0 = a0; m2 = b0; m1 = a1; m3 = b1;
  }
  const Rot jr = make_jacobi(m0, m1, m3);
  const Rot jl = rot_mul(rot1, Rot{jr.c, -jr.s});
  if (!(jl.c == 1.f && jl.s == 0.f)) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {  // rows P, Q of W  (applyOnTheLeft)
      const float x = W[P * 3 + i], y = W[Q * 3 + i];
      W[P * 3 + i] = jl.c * x + jl.s * y;
      W[Q * 3 + i] = -jl.s * x + jl.c * y;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {  // columns P, Q of U  (U.applyOnTheRight(p, q, j_left^T))
      const float x = U[i * 3 + P], y = U[i * 3 + Q];
      U[i * 3 + P] = jl.c * x + jl.s * y;
      U[i * 3 + Q] = -jl.s * x + jl.c * y;
    }
  }
  const Rot jrt{jr.c, -jr.s};
  if (!(jrt.c == 1.f && jrt.s == 0.f)) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {  // columns P, Q of W  (applyOnTheRight(p, q, j_right))
      const float x = W[i * 3 + P], y = W[i * 3 + Q];
      W[i * 3 + P] = jrt.c * x + jrt.s * y;
      W[i * 3 + Q] = -jrt.s * x + jrt.c * y;
    }
  }
  maxDiag = fmaxf(maxDiag, fmaxf(fabsf(W[P * 3 + P]), fabsf(W[Q * 3 + Q])));
}

template <int I, int J>
__device__ __forceinline__ void swap_cols(float (&U)[9], float (&sv)[3]) {
  const float ts = sv[I]; sv[I] = sv[J]; sv[J] = ts;
#pragma unroll
  for (int r = 0; r < 3; ++r) { const float tu = U[r * 3 + I]; U[r * 3 + I] = U[r * 3 + J]; U[r * 3 + J] = tu; }
}

__device__ __forceinline__ void jacobi_svd3(const float (&A)[9], float (&U)[9], float (&sv)[3]) {
  float scale = 0.f;
#pragma unroll
  for (int i = 0; i < 9; ++i) scale = fmaxf(scale, fabsf(A[i]));
#pragma unroll
  for (int i = 0; i < 9; ++i) U[i] = (i % 4 == 0) ? 1.f : 0.f;
  if (!isfinite(scale)) { sv[0] = sv[1] = sv[2] = NAN; return; }
  if (scale == 0.f) scale = 1.f;
  float W[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) W[i] = __fdiv_rn(A[i], scale);
  float maxDiag = fmaxf(fabsf(W[0]), fmaxf(fabsf(W[4]), fabsf(W[8])));
  bool finished = false;
#pragma unroll 1
  for (int guard = 0; !finished && guard < 1000; ++guard) {
    finished = true;
    jacobi_pair<1, 0>(W, U, maxDiag, finished);  // sweep order p = 1..2, q < p
    jacobi_pair<2, 0>(W, U, maxDiag, finished);
    jacobi_pair<2, 1>(W, U, maxDiag, finished);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float a = W[i * 3 + i];
    sv[i] = fabsf(a);
    if (a < 0.f) {
#pragma unroll
      for (int r = 0; r < 3; ++r) U[r * 3 + i] = -U[r * 3 + i];
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) sv[i] *= scale;
  // selection sort, descending, swapping U columns (first maximum wins ties; stop at a zero maximum)
  {
    int pos = 0;
    float best = sv[0];
    if (sv[1] > best) { best = sv[1]; pos = 1; }
    if (sv[2] > best) { best = sv[2]; pos = 2; }
    if (best == 0.f) return;
    if (pos == 1) swap_cols<0, 1>(U, sv);
    else if (pos == 2) swap_cols<0, 2>(U, sv);
  }
  {
    if (sv[2] > sv[1]) swap_cols<1, 2>(U, sv);   // (a zero maximum leaves equal zeros: no swap either way)
  }
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
  return v;
}

// covariance sums -> trg.cpp:337-363 (cov, JacobiSVD, Frobenius-normalised U, weight)
__device__ __forceinline__ float weight_from_sums(int npts, double sx, double sy, double sz, double sxx, double sxy,
                                                  double sxz, double syy, double syz, double szz) {
  // :337-338 cov = centered^T centered / (n-1)
  const double nn = (double)npts, dn = (double)(npts - 1);
  float cov[9];
  cov[0] = (float)((sxx - sx * sx / nn) / dn);
  cov[1] = cov[3] = (float)((sxy - sx * sy / nn) / dn);
  cov[2] = cov[6] = (float)((sxz - sx * sz / nn) / dn);
  cov[4] = (float)((syy - sy * sy / nn) / dn);
  cov[5] = cov[7] = (float)((syz - sy * sz / nn) / dn);
  cov[8] = (float)((szz - sz * sz / nn) / dn);
  float U[9], sv[3];
  jacobi_svd3(cov, U, sv);
  // :340 matrixU().normalized(): Frobenius norm of the 3x3
  float fro = 0.f;
  for (int k = 0; k < 9; ++k) fro += U[k] * U[k];
  float e20 = U[6], e21 = U[7];
  if (fro > 0.f) {
    const float nrm = __fsqrt_rn(fro);
    e20 = __fdiv_rn(e20, nrm);
    e21 = __fdiv_rn(e21, nrm);
  }
  const float hor = fabsf(e20), ver = fabsf(e21);  // :347-354
  const float ratio = 0.8f;
  float weight = __fadd_rn(__fmul_rn(ratio, hor), __fmul_rn(__fsub_rn(1.f, ratio), ver));  // :360
  if ((double)weight < 0.1) weight = 0.f;  // :361-363
  return weight;
}

struct EdgeGeom {
  float dist, dirx, diry;
};
// :276 dist = (node1 - node2).norm() ; :277 dir = (node2 - node1).normalized()
__device__ __forceinline__ EdgeGeom edge_geom(float p1x, float p1y, float p2x, float p2y) {
  EdgeGeom g;
  const float ax = __fsub_rn(p1x, p2x), ay = __fsub_rn(p1y, p2y);
  g.dist = __fsqrt_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)));
  const float ex = __fsub_rn(p2x, p1x), ey = __fsub_rn(p2y, p1y);
  const float sq = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
  g.dirx = ex;
  g.diry = ey;
  if (sq > 0.f) {
    const float s = __fsqrt_rn(sq);
    g.dirx = __fdiv_rn(ex, s);
    g.diry = __fdiv_rn(ey, s);
  }
  return g;
}

// geometric part of TRG::wireEdge for ONE edge executed by ONE WARP (any density / radius)
__device__ __forceinline__ void warp_edge_eval(const MapView& m, float p1x, float p1y, float p1z, float p2x, float p2y,
                                               float rs, float hthr, float cthr, float* zbuf, int cap, int* stage_o,
                                               float* weight_o, float* dist_o, int* npts_o) {
  const EdgeGeom g = edge_geom(p1x, p1y, p2x, p2y);
  const float dist = g.dist, dirx = g.dirx, diry = g.diry;
  int stage = TRGB_EDGE_OK;
  float weight = 0.f;
  int npts = 0;
  // :282-288 collision samples every robot_size/2 along the segment (float accumulator)
  const float ds = 0.5f * rs;
  for (float t = 0.f; t < dist; t = __fadd_rn(t, ds)) {
    const float sx = __fadd_rn(p1x, __fmul_rn(t, dirx)), sy = __fadd_rn(p1y, __fmul_rn(t, diry));
    if (warp_is_collision(m, sx, sy, rs, hthr, cthr, zbuf, cap, nullptr)) {
      stage = TRGB_EDGE_COLLISION;
      break;
    }
  }
  if (stage == TRGB_EDGE_OK) {
    // :291-297 ellipse with foci at the two nodes (circle when the nodes are close)
    const float c = 0.5f * dist;
    const float b = rs;
    float a = b;
    if (c >= b) a = __fsqrt_rn(__fadd_rn(__fmul_rn(c, c), __fmul_rn(b, b)));
    const bool circle = (a == b);
    const float cx = __fadd_rn(p1x, __fmul_rn(c, dirx)), cy = __fadd_rn(p1y, __fmul_rn(c, diry));
    const float a2 = __fmul_rn(a, a), b2 = __fmul_rn(b, b);
    const float rhs = __fmul_rn(__fmul_rn(a2, b), b);  // a*a*b*b, left to right
    const float ndiry = -diry;
    int nrange = 0;
    double sx = 0, sy = 0, sz = 0, sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0, szz = 0;
    warp_for_each_candidate(m, cx, cy, inflate(a, cx, cy), [&](bool valid, const float4& p) {
      const float qx = __fsub_rn(p.x, cx), qy = __fsub_rn(p.y, cy);
      const float d2 = __fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qy, qy));
      const bool in_range = valid && d2 <= a2;  // kd_nearest_range2(center, a)
      nrange += in_range;
      // :312-316 p2d = R * (pt - center), R = [dir.x -dir.y; dir.y dir.x]
      const float px = __fadd_rn(__fmul_rn(dirx, qx), __fmul_rn(ndiry, qy));
      const float py = __fadd_rn(__fmul_rn(diry, qx), __fmul_rn(dirx, qy));
      bool keep = in_range;
      if (!circle)  // :320
        keep = keep && (__fadd_rn(__fmul_rn(__fmul_rn(px, px), b2), __fmul_rn(__fmul_rn(py, py), a2)) < rhs);
      if (keep) {
        // covariance sums in double about the pivot z = p1.z: order-independent to ~1e-16,
        // i.e. at least as close to the reference's float result as any float ordering
        const double X = px, Y = py, Z = (double)p.z - (double)p1z;
        npts += 1;
        sx += X; sy += Y; sz += Z;
        sxx += X * X; sxy += X * Y; sxz += X * Z; syy += Y * Y; syz += Y * Z; szz += Z * Z;
      }
    });
    nrange = __reduce_add_sync(FULL, nrange);
    npts = __reduce_add_sync(FULL, npts);
    if (nrange == 0) stage = TRGB_EDGE_EMPTY;       // :305
    else if (npts < 3) stage = TRGB_EDGE_FEWPTS;    // :327
    else {
      sx = warp_sum_d(sx); sy = warp_sum_d(sy); sz = warp_sum_d(sz);
      sxx = warp_sum_d(sxx); sxy = warp_sum_d(sxy); sxz = warp_sum_d(sxz);
      syy = warp_sum_d(syy); syz = warp_sum_d(syz); szz = warp_sum_d(szz);
      weight = weight_from_sums(npts, sx, sy, sz, sxx, sxy, sxz, syy, syz, szz);
    }
  }
  *stage_o = stage; *weight_o = weight; *dist_o = dist; *npts_o = npts;
}

__global__ void __launch_bounds__(kThreads) k_edge_eval(MapView m, const float* __restrict__ p1_xyz,
                                                        const float2* __restrict__ p2_xy, int64_t n,
                                                        float rs, float hthr, float cthr, int cap,
                                                        uint8_t* __restrict__ stage_out,
                                                        float* __restrict__ w_out,
                                                        float* __restrict__ dist_out,
                                                        int32_t* __restrict__ npts_out) {
  extern __shared__ float zsm[];
  float* zbuf = zsm + (threadIdx.x >> 5) * (cap + 256);  // + 256 bins for the histogram select
  const int lane = threadIdx.x & 31;
  const int64_t wstride = (int64_t)gridDim.x * kWarpsPerCta;
  for (int64_t i = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5); i < n; i += wstride) {
    const float p1x = __ldg(p1_xyz + 3 * i), p1y = __ldg(p1_xyz + 3 * i + 1), p1z = __ldg(p1_xyz + 3 * i + 2);
    const float2 p2 = __ldg(p2_xy + i);
    int stage, npts;
    float weight, dist;
    warp_edge_eval(m, p1x, p1y, p1z, p2.x, p2.y, rs, hthr, cthr, zbuf, cap, &stage, &weight, &dist, &npts);
    if (lane == 0) {
      stage_out[i] = (uint8_t)stage;
      w_out[i] = weight;
      dist_out[i] = dist;
      if (npts_out) npts_out[i] = npts;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Thread-per-item kernels (the fast path for densities where a cylinder holds <= cap points).
// Items whose cylinder overflows the per-thread column fall back, inside the same launch, to the
// warp-cooperative routines above, so results never depend on which path ran.
// ------------------------------------------------------------------------------------------
// shared layout during the thread phase: column of thread t = zsm[k * kTqThreads + t];
// during the fallback phase warp w owns the contiguous slice zsm[w * 32 * cap ...)
#define TQ_FALLBACK_BEGIN(res)                                 \
  if (__syncthreads_or((res) == 2)) {                          \
    float* wbuf = zsm + (threadIdx.x >> 5) * 32 * cap;         \
    unsigned ov = __ballot_sync(FULL, (res) == 2);             \
    while (ov) {                                               \
      const int src = __ffs(ov) - 1;                           \
      ov &= ov - 1;
#define TQ_FALLBACK_END \
    }                   \
    __syncthreads();    \
  }

// item index -> (owner, slot): the emulated 64-bit division costs ~100 instructions per thread;
// every batch the builder launches fits 32 bits (uniform branch)
__device__ __forceinline__ int64_t idiv_items(int64_t it, int per, int64_t items) {
  if (items < (1ll << 31)) return (int64_t)((uint32_t)it / (uint32_t)per);
  return it / per;
}

__global__ void TQ_BOUNDS k_collision_tq(MapView m, const float2* __restrict__ q, int64_t n,
                                                             float r, float hthr, float rthr, int cap,
                                                             uint8_t* __restrict__ out) {
  extern __shared__ float zsm[];
  float* zcol = zsm + threadIdx.x;
  const int lane = threadIdx.x & 31;
  for (int64_t base = (int64_t)blockIdx.x * kTqThreads; base < n; base += (int64_t)gridDim.x * kTqThreads) {
    const int64_t i = base + threadIdx.x;
    float2 p = make_float2(0.f, 0.f);
    int res = 0;
    if (i < n) {
      p = __ldg(q + i);
      res = thread_is_collision(m, p.x, p.y, r, hthr, rthr, zcol, kTqThreads);
    }
    TQ_FALLBACK_BEGIN(res)
      const float qx = __shfl_sync(FULL, p.x, src), qy = __shfl_sync(FULL, p.y, src);
      const bool c = warp_is_collision(m, qx, qy, r, hthr, rthr, wbuf, 32 * cap - 256, nullptr);
      if (lane == src) res = c ? 1 : 0;
    TQ_FALLBACK_END
    if (i < n) out[i] = (uint8_t)res;
  }
}

// one thread per (node, draw); bit j of mask[node] = isCollision(node + draw[first+j])
__global__ void TQ_BOUNDS k_sample_window_tq(
    MapView m, const float2* __restrict__ node_xy, const int32_t* __restrict__ first_draw,
    const float2* __restrict__ draw_xy, int64_t n_nodes, int window, float r, float hthr, float rthr,
    int cap, unsigned long long* __restrict__ mask) {
  extern __shared__ float zsm[];
  float* zcol = zsm + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int64_t items = n_nodes * window;
  const int words = (window + 63) >> 6;
  for (int64_t base = (int64_t)blockIdx.x * kTqThreads; base < items; base += (int64_t)gridDim.x * kTqThreads) {
    const int64_t it = base + threadIdx.x;
    float sx = 0.f, sy = 0.f;
    int res = 0;
    int64_t node = 0;
    int j = 0;
    if (it < items) {
      node = idiv_items(it, window, items);
      j = (int)(it - node * window);
      const float2 np = __ldg(node_xy + node);
      const float2 d = __ldg(draw_xy + (__ldg(first_draw + node) + j));
      // sample = node->pos_.head(2) + Vector2f(e*cos, e*sin)   (trg.cpp:396-397); d = (e*cos, e*sin)
      sx = __fadd_rn(np.x, d.x);
      sy = __fadd_rn(np.y, d.y);
      res = thread_is_collision(m, sx, sy, r, hthr, rthr, zcol, kTqThreads);
    }
    TQ_FALLBACK_BEGIN(res)
      const float qx = __shfl_sync(FULL, sx, src), qy = __shfl_sync(FULL, sy, src);
      const bool c = warp_is_collision(m, qx, qy, r, hthr, rthr, wbuf, 32 * cap - 256, nullptr);
      if (lane == src) res = c ? 1 : 0;
    TQ_FALLBACK_END
    if (it < items && res) atomicOr(mask + node * words + (j >> 6), 1ull << (j & 63));
  }
}

// Sampling windows with the node's neighbourhood staged in shared memory: one CTA per node. All
// `window` samples of a node lie within max_offset of it, so the cells they can touch form one
// small block of the grid (<= kStRows x kStCols cells, <= kStPts points, a few KB): its cell table
// and its points are copied once, coalesced, into shared memory and every thread then walks its
// query's cell runs there instead of issuing per-thread global loads. Nodes whose neighbourhood
// does not fit (very dense spots) and threads whose query leaves the staged block use the global path.
constexpr int kStRows = 16, kStCols = 16, kStPts = 1024;

__device__ __forceinline__ int thread_is_collision_staged(const float4* __restrict__ spts, const uint32_t* __restrict__ scs,
                                                          const uint32_t* __restrict__ srow, int cols1, int bx0, int by0,
                                                          const MapView& m, float qx, float qy, float r, float hthr,
                                                          float rthr, float* zcol, int stride, int cx0, int cx1, int cy0,
                                                          int cy1) {
  const float r2 = __fmul_rn(r, r);
  int n = 0;
  for (int row = cy0; row <= cy1; ++row) {
    const int rr_ = row - by0;
    const uint32_t base = scs[rr_ * cols1];
    const uint32_t s = scs[rr_ * cols1 + (cx0 - bx0)] - base + srow[rr_];
    const uint32_t e = scs[rr_ * cols1 + (cx1 - bx0) + 1] - base + srow[rr_];
#pragma unroll 4
    for (uint32_t i = s; i < e; ++i) {
      const float4 p = spts[i];
      const float dx = __fsub_rn(p.x, qx), dy = __fsub_rn(p.y, qy);
      const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
      if (d2 <= r2) {
        if (n < kTqCap) zcol[n * stride] = p.z;
        ++n;
      }
    }
  }
  if (n == 0) return 1;
  if (n > kTqCap) return 2;
  int cnt;
  if (__all_sync(__activemask(), n <= 32)) cnt = median_outlier_count<32>(zcol, stride, n, hthr);
  else cnt = median_outlier_count<64>(zcol, stride, n, hthr);
  const float ratio = __fdiv_rn((float)cnt, (float)n);
  return ratio > rthr ? 1 : 0;
}

__global__ void __launch_bounds__(kTqThreads) k_sample_window_sm(
    MapView m, const float2* __restrict__ node_xy, const int32_t* __restrict__ first_draw,
    const float2* __restrict__ draw_xy, int64_t n_nodes, int window, float max_offset, float r, float hthr, float rthr,
    int cap, unsigned long long* __restrict__ mask) {
  extern __shared__ float zsm[];                       // kTqThreads * kTqCap floats (z columns / fallback buffers)
  __shared__ float4 spts[kStPts];
  __shared__ uint32_t scs[kStRows * (kStCols + 1)];
  __shared__ uint32_t srow[kStRows + 1];
  __shared__ int s_ok;
  float* zcol = zsm + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int words = (window + 63) >> 6;
  for (int64_t node = blockIdx.x; node < n_nodes; node += gridDim.x) {
    const float2 np = __ldg(node_xy + node);
    // block of cells any sample of this node can touch
    const float reach = max_offset + inflate(r, fabsf(np.x) + max_offset, fabsf(np.y) + max_offset);
    const int bx0 = cell_coord(np.x - reach, m.x0, m.inv_cell, m.W), bx1 = cell_coord(np.x + reach, m.x0, m.inv_cell, m.W);
    const int by0 = cell_coord(np.y - reach, m.y0, m.inv_cell, m.H), by1 = cell_coord(np.y + reach, m.y0, m.inv_cell, m.H);
    const int rows = by1 - by0 + 1, cols1 = bx1 - bx0 + 2;
    const bool fits = rows <= kStRows && cols1 <= kStCols + 1;
    if (fits) {
      for (int t = threadIdx.x; t < rows * cols1; t += kTqThreads) {
        const int rr_ = t / cols1, cc = t - rr_ * cols1;
        scs[rr_ * cols1 + cc] = __ldg(m.cell_start + (size_t)(by0 + rr_) * m.W + bx0 + cc);
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t acc = 0;
      if (fits) {
        for (int rr_ = 0; rr_ < rows; ++rr_) {
          srow[rr_] = acc;
          acc += scs[rr_ * cols1 + cols1 - 1] - scs[rr_ * cols1];
        }
        srow[rows] = acc;
      }
      s_ok = fits && acc <= (uint32_t)kStPts;
    }
    __syncthreads();
    const bool staged = s_ok != 0;
    if (staged) {
      const uint32_t total = srow[rows];
      for (uint32_t i = threadIdx.x; i < total; i += kTqThreads) {
        int rr_ = 0;
        while (i >= srow[rr_ + 1]) ++rr_;
        spts[i] = ld_pt(m.pts + scs[rr_ * cols1] + (i - srow[rr_]));
      }
    }
    __syncthreads();
    const int32_t fd = __ldg(first_draw + node);
    for (int jb = 0; jb < window; jb += kTqThreads) {   // block-uniform trip count
      const int j = jb + threadIdx.x;
      float sx = 0.f, sy = 0.f;
      int res = 0;
      if (j < window) {
        const float2 d = __ldg(draw_xy + (fd + j));
        sx = __fadd_rn(np.x, d.x);   // trg.cpp:396-397
        sy = __fadd_rn(np.y, d.y);
        const float rr = inflate(r, sx, sy);
        const int cx0 = cell_coord(sx - rr, m.x0, m.inv_cell, m.W), cx1 = cell_coord(sx + rr, m.x0, m.inv_cell, m.W);
        const int cy0 = cell_coord(sy - rr, m.y0, m.inv_cell, m.H), cy1 = cell_coord(sy + rr, m.y0, m.inv_cell, m.H);
        if (staged && cx0 >= bx0 && cx1 <= bx1 && cy0 >= by0 && cy1 <= by1)
          res = thread_is_collision_staged(spts, scs, srow, cols1, bx0, by0, m, sx, sy, r, hthr, rthr, zcol, kTqThreads, cx0,
                                           cx1, cy0, cy1);
        else
          res = thread_is_collision(m, sx, sy, r, hthr, rthr, zcol, kTqThreads);
      }
      TQ_FALLBACK_BEGIN(res)
        const float qx = __shfl_sync(FULL, sx, src), qy = __shfl_sync(FULL, sy, src);
        const bool c = warp_is_collision(m, qx, qy, r, hthr, rthr, wbuf, 32 * cap - 256, nullptr);
        if (lane == src) res = c ? 1 : 0;
      TQ_FALLBACK_END
      if (j < window && res) atomicOr(mask + node * words + (j >> 6), 1ull << (j & 63));
    }
    __syncthreads();   // spts / scs are rewritten for the next node
  }
}

// ---- K4 fast path, two launches on one stream ------------------------------------------------
// (a) k_edge_collide_tq: one thread per (edge, segment sample k). Sample k sits at the k-th value
//     of the reference's accumulating float counter `for (float i = 0; i < dist; i += ds)`
//     (trg.cpp:283). Any colliding sample marks the edge TRGB_EDGE_COLLISION in `stage`
//     (zeroed by the launcher). kmax = samples needed by the longest edge of the batch.
// (b) k_edge_pca<L>: L = 4 or 8 lanes per edge split the cell rows of the ellipse gather, combine
//     their covariance sums and run the Jacobi SVD; edges already marked colliding are skipped.
__global__ void TQ_BOUNDS k_edge_collide_tq(MapView m, const float* __restrict__ p1_xyz,
                                                                const float2* __restrict__ p2_xy, int64_t n, int kmax,
                                                                float rs, float hthr, float cthr, int cap,
                                                                uint8_t* __restrict__ stage,
                                                                const float* __restrict__ skip_d2, int64_t n_skip,
                                                                float skip_below) {
  extern __shared__ float zsm[];
  float* zcol = zsm + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int64_t items = n * kmax;
  for (int64_t base = (int64_t)blockIdx.x * kTqThreads; base < items; base += (int64_t)gridDim.x * kTqThreads) {
    const int64_t it = base + threadIdx.x;
    float sx = 0.f, sy = 0.f;
    int res = 0;
    int64_t e = 0;
    bool live = it < items;
    if (live) {
      e = idiv_items(it, kmax, items);
      if (e < n_skip && __fsqrt_rn(__ldg(skip_d2 + e)) < skip_below) live = false;  // speculation filter
    }
    if (live) {
      const int k = (int)(it - e * kmax);
      const float p1x = __ldg(p1_xyz + 3 * e), p1y = __ldg(p1_xyz + 3 * e + 1);
      const float2 p2 = __ldg(p2_xy + e);
      const EdgeGeom g = edge_geom(p1x, p1y, p2.x, p2.y);
      const float ds = 0.5f * rs;
      float t = 0.f;
      for (int j = 0; j < k; ++j) t = __fadd_rn(t, ds);
      // the thread of the last slot also walks any samples beyond kmax (long edges), so no
      // sample is ever skipped whatever kmax is
      while (t < g.dist) {
        sx = __fadd_rn(p1x, __fmul_rn(t, g.dirx));
        sy = __fadd_rn(p1y, __fmul_rn(t, g.diry));
        res = thread_is_collision(m, sx, sy, rs, hthr, cthr, zcol, kTqThreads);
        if (res != 0 || k != kmax - 1) break;
        t = __fadd_rn(t, ds);
      }
    }
    TQ_FALLBACK_BEGIN(res)
      const float qx = __shfl_sync(FULL, sx, src), qy = __shfl_sync(FULL, sy, src);
      const bool c = warp_is_collision(m, qx, qy, rs, hthr, cthr, wbuf, 32 * cap - 256, nullptr);
      if (lane == src) res = c ? 1 : 0;
    TQ_FALLBACK_END
    if (it < items && res == 1) stage[e] = (uint8_t)TRGB_EDGE_COLLISION;
  }
}

template <int kPcaLanes>
__global__ void __launch_bounds__(256) k_edge_pca(MapView m, const float* __restrict__ p1_xyz,
                                                  const float2* __restrict__ p2_xy, int64_t n, float rs,
                                                  uint8_t* __restrict__ stage_io, float* __restrict__ w_out,
                                                  float* __restrict__ dist_out, int32_t* __restrict__ npts_out,
                                                  const float* __restrict__ skip_d2, int64_t n_skip,
                                                  float skip_below) {
  __shared__ double s_sum[9][256 / kPcaLanes];
  __shared__ int s_cnt[3][256 / kPcaLanes];
  __shared__ float s_dist[256 / kPcaLanes];
  const int sub = threadIdx.x & (kPcaLanes - 1);
  const int64_t gstride = (int64_t)gridDim.x * (256 / kPcaLanes);
  for (int64_t base = (int64_t)blockIdx.x * (256 / kPcaLanes); base < n; base += gstride) {
    const int64_t i = base + (threadIdx.x / kPcaLanes);
    const bool live = i < n;
    float p1x = 0.f, p1y = 0.f, p1z = 0.f, p2x = 1.f, p2y = 0.f;
    int st = TRGB_EDGE_OK;
    if (live) {
      p1x = __ldg(p1_xyz + 3 * i); p1y = __ldg(p1_xyz + 3 * i + 1); p1z = __ldg(p1_xyz + 3 * i + 2);
      const float2 p2 = __ldg(p2_xy + i);
      p2x = p2.x; p2y = p2.y;
      st = stage_io[i];
      if (i < n_skip && __fsqrt_rn(__ldg(skip_d2 + i)) < skip_below) st = TRGB_EDGE_SKIPPED;
    }
    const EdgeGeom g = edge_geom(p1x, p1y, p2x, p2y);
    const float dist = g.dist, dirx = g.dirx, diry = g.diry;
    int nrange = 0, npts = 0;
    double sx = 0, sy = 0, sz = 0, sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0, szz = 0;
    if (live && st != TRGB_EDGE_COLLISION && st != TRGB_EDGE_SKIPPED) {
      // :291-297 ellipse with foci at the two nodes (circle when the nodes are close)
      const float c = 0.5f * dist;
      const float b = rs;
      float a = b;
      if (c >= b) a = __fsqrt_rn(__fadd_rn(__fmul_rn(c, c), __fmul_rn(b, b)));
      const bool circle = (a == b);
      const float cx = __fadd_rn(p1x, __fmul_rn(c, dirx)), cy = __fadd_rn(p1y, __fmul_rn(c, diry));
      const float a2 = __fmul_rn(a, a), b2 = __fmul_rn(b, b);
      const float rhs = __fmul_rn(__fmul_rn(a2, b), b);  // a*a*b*b, left to right
      const float ndiry = -diry;
      const float rr = inflate(a, cx, cy);
      const int cx0 = cell_coord(cx - rr, m.x0, m.inv_cell, m.W), cx1 = cell_coord(cx + rr, m.x0, m.inv_cell, m.W);
      const int cy0 = cell_coord(cy - rr, m.y0, m.inv_cell, m.H), cy1 = cell_coord(cy + rr, m.y0, m.inv_cell, m.H);
      for (int row = cy0 + sub; row <= cy1; row += kPcaLanes) {
        int ca, cb;
        row_chord(m, cx, cy, rr, row, cx0, cx1, &ca, &cb);  // chord of the radius-a disc in this row
        if (ca > cb) continue;
        const size_t rb = (size_t)row * (size_t)m.W;
        const uint32_t s = __ldg(m.cell_start + rb + ca), e = __ldg(m.cell_start + rb + cb + 1);
#pragma unroll 4
        for (uint32_t k = s; k < e; ++k) {
          const float4 p = ld_pt(m.pts + k);
          const float qx = __fsub_rn(p.x, cx), qy = __fsub_rn(p.y, cy);
          const float d2 = __fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qy, qy));
          if (!(d2 <= a2)) continue;  // kd_nearest_range2(center, a)
          ++nrange;
          // :312-316 p2d = R * (pt - center), R = [dir.x -dir.y; dir.y dir.x]
          const float px = __fadd_rn(__fmul_rn(dirx, qx), __fmul_rn(ndiry, qy));
          const float py = __fadd_rn(__fmul_rn(diry, qx), __fmul_rn(dirx, qy));
          if (!circle && !(__fadd_rn(__fmul_rn(__fmul_rn(px, px), b2), __fmul_rn(__fmul_rn(py, py), a2)) < rhs)) continue;
          // covariance sums in double about the pivot z = p1.z (see warp_edge_eval)
          const double X = px, Y = py, Z = (double)p.z - (double)p1z;
          ++npts;
          sx += X; sy += Y; sz += Z;
          sxx += X * X; sxy += X * Y; sxz += X * Z; syy += Y * Y; syz += Y * Z; szz += Z * Z;
        }
      }
    }
    // combine the kPcaLanes partial sums (all lanes of the warp take part; dead groups add zeros)
#pragma unroll
    for (int d = 1; d < kPcaLanes; d <<= 1) {
      nrange += __shfl_xor_sync(FULL, nrange, d);
      npts += __shfl_xor_sync(FULL, npts, d);
      sx += __shfl_xor_sync(FULL, sx, d); sy += __shfl_xor_sync(FULL, sy, d); sz += __shfl_xor_sync(FULL, sz, d);
      sxx += __shfl_xor_sync(FULL, sxx, d); sxy += __shfl_xor_sync(FULL, sxy, d); sxz += __shfl_xor_sync(FULL, sxz, d);
      syy += __shfl_xor_sync(FULL, syy, d); syz += __shfl_xor_sync(FULL, syz, d); szz += __shfl_xor_sync(FULL, szz, d);
    }
    // phase 2: the block's 64 edges are finished by its first 64 threads, one edge per thread, so the
    // Jacobi SVD runs in two full warps instead of one lane in four of all eight
    const int le = threadIdx.x / kPcaLanes;
    if (sub == 0) {
      s_sum[0][le] = sx; s_sum[1][le] = sy; s_sum[2][le] = sz; s_sum[3][le] = sxx; s_sum[4][le] = sxy;
      s_sum[5][le] = sxz; s_sum[6][le] = syy; s_sum[7][le] = syz; s_sum[8][le] = szz;
      s_cnt[0][le] = npts; s_cnt[1][le] = nrange; s_cnt[2][le] = live ? st : -1;
      s_dist[le] = dist;
    }
    __syncthreads();
    if (threadIdx.x < 256 / kPcaLanes) {
      const int t = threadIdx.x;
      const int64_t ei = base + t;
      int st2 = s_cnt[2][t];
      if (ei < n && st2 >= 0) {
        int np2 = s_cnt[0][t];
        const int nr2 = s_cnt[1][t];
        float weight = 0.f;
        if (st2 == TRGB_EDGE_COLLISION || st2 == TRGB_EDGE_SKIPPED) {
          np2 = 0;
        } else if (nr2 == 0) {
          st2 = TRGB_EDGE_EMPTY;   // :305
        } else if (np2 < 3) {
          st2 = TRGB_EDGE_FEWPTS;  // :327
        } else {
          weight = weight_from_sums(np2, s_sum[0][t], s_sum[1][t], s_sum[2][t], s_sum[3][t], s_sum[4][t], s_sum[5][t],
                                    s_sum[6][t], s_sum[7][t], s_sum[8][t]);
        }
        stage_io[ei] = (uint8_t)st2;
        w_out[ei] = weight;
        dist_out[ei] = s_dist[t];
        if (npts_out) npts_out[ei] = np2;
      }
    }
    __syncthreads();
  }
}

#ifndef PCA_VARIANT
#define PCA_VARIANT 0  // ellipse gather of k_edge_pca_t: 0 = row by row, 1 = flattened over rows, 2 = row by row in 256-bit groups
#endif
#ifndef PCA_MINBLK
#define PCA_MINBLK 8  // 63 registers: measured best at saturation (0.38 ms per 1e6 edges; 0.43 with ptxas' own 72)
#endif
constexpr int kPcaRows = 7;  // cell rows the flattened ellipse gather holds in registers

// (b') k_edge_pca_t: one THREAD per edge. Every lane walks the cell rows of its own ellipse — a contiguous run of
// points per row, streamed through L1 (neighbouring edges of a spatially coherent batch share the lines) — so no
// lane idles on a short row, nothing is combined across lanes and the Jacobi SVD runs in every lane at once.
__global__ void __launch_bounds__(128, PCA_MINBLK) k_edge_pca_t(MapView m, const float* __restrict__ p1_xyz,
                                                    const float2* __restrict__ p2_xy, int64_t n, float rs,
                                                    uint8_t* __restrict__ stage_io, float* __restrict__ w_out,
                                                    float* __restrict__ dist_out, int32_t* __restrict__ npts_out,
                                                    const float* __restrict__ skip_d2, int64_t n_skip,
                                                    float skip_below) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float p1x = __ldg(p1_xyz + 3 * i), p1y = __ldg(p1_xyz + 3 * i + 1), p1z = __ldg(p1_xyz + 3 * i + 2);
    const float2 p2 = __ldg(p2_xy + i);
    int st = stage_io[i];
    if (i < n_skip && __fsqrt_rn(__ldg(skip_d2 + i)) < skip_below) st = TRGB_EDGE_SKIPPED;
    const EdgeGeom g = edge_geom(p1x, p1y, p2.x, p2.y);
    const float dist = g.dist, dirx = g.dirx, diry = g.diry;
    int nrange = 0, npts = 0;
    double sx = 0, sy = 0, sz = 0, sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0, szz = 0;
    if (st != TRGB_EDGE_COLLISION && st != TRGB_EDGE_SKIPPED) {
      // :291-297 ellipse with foci at the two nodes (circle when the nodes are close)
      const float c = 0.5f * dist;
      const float b = rs;
      float a = b;
      if (c >= b) a = __fsqrt_rn(__fadd_rn(__fmul_rn(c, c), __fmul_rn(b, b)));
      const bool circle = (a == b);
      const float cx = __fadd_rn(p1x, __fmul_rn(c, dirx)), cy = __fadd_rn(p1y, __fmul_rn(c, diry));
      const float a2 = __fmul_rn(a, a), b2 = __fmul_rn(b, b);
      const float rhs = __fmul_rn(__fmul_rn(a2, b), b);  // a*a*b*b, left to right
      const float ndiry = -diry;
      const float rr = inflate(a, cx, cy);
      const int cx0 = cell_coord(cx - rr, m.x0, m.inv_cell, m.W), cx1 = cell_coord(cx + rr, m.x0, m.inv_cell, m.W);
      const int cy0 = cell_coord(cy - rr, m.y0, m.inv_cell, m.H), cy1 = cell_coord(cy + rr, m.y0, m.inv_cell, m.H);
#define PCA_POINT(P, OK)                                                                                           \
      {                                                                                                            \
        const float qx = __fsub_rn(P.x, cx), qy = __fsub_rn(P.y, cy);                                              \
        const float d2 = __fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qy, qy));                                          \
        if ((OK) && d2 <= a2) { /* kd_nearest_range2(center, a) */                                                 \
          ++nrange;                                                                                                \
          /* :312-316 p2d = R * (pt - center), R = [dir.x -dir.y; dir.y dir.x] */                                  \
          const float px = __fadd_rn(__fmul_rn(dirx, qx), __fmul_rn(ndiry, qy));                                   \
          const float py = __fadd_rn(__fmul_rn(diry, qx), __fmul_rn(dirx, qy));                                    \
          if (circle || __fadd_rn(__fmul_rn(__fmul_rn(px, px), b2), __fmul_rn(__fmul_rn(py, py), a2)) < rhs) {     \
            /* covariance sums in double about the pivot z = p1.z (see warp_edge_eval) */                          \
            const double X = px, Y = py, Z = (double)P.z - (double)p1z;                                            \
            ++npts;                                                                                                \
            sx += X; sy += Y; sz += Z;                                                                             \
            sxx += X * X; sxy += X * Y; sxz += X * Z; syy += Y * Y; syz += Y * Z; szz += Z * Z;                    \
          }                                                                                                        \
        }                                                                                                          \
      }
      if (PCA_VARIANT == 1 && cy1 - cy0 < kPcaRows) {
        // flattened gather (see thread_is_collision): the rows' runs are located first, then walked as one loop
        // over groups of four points fetched with two 256-bit loads - lanes differ in how their candidates split
        // over rows, hardly in the total
        uint32_t gb[kPcaRows], gs[kPcaRows], ge[kPcaRows], gc[kPcaRows];
        uint32_t totg = 0;
        const float slack = 1e-3f * m.cell + 4e-6f * (fabsf(cy) + fabsf(m.y0));
#pragma unroll
        for (int k = 0; k < kPcaRows; ++k) {
          const int row = cy0 + k;
          const float ylo = m.y0 + (float)row * m.cell - slack;
          const float yhi = m.y0 + (float)(row + 1) * m.cell + slack;
          const float dy = fmaxf(0.f, fmaxf(ylo - cy, cy - yhi));
          const float h2 = rr * rr - dy * dy;
          float half;
          asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(half) : "f"(fmaxf(h2, 0.f)));
          half = half * 1.0001f + 1e-6f;
          const int ca = max(cx0, min(m.W - 1, __float2int_rd(__fmul_rn(__fsub_rn(cx - half, m.x0), m.inv_cell))));
          const int cb = min(cx1, max(0, __float2int_rd(__fmul_rn(__fsub_rn(cx + half, m.x0), m.inv_cell))));
          const bool any = row <= cy1 && h2 > 0.f && ca <= cb;
          const uint32_t* cs = m.cell_start + (size_t)min(row, cy1) * (size_t)m.W;
          const uint32_t s0 = any ? __ldg(cs + ca) : 0u;
          const uint32_t e0 = any ? __ldg(cs + cb + 1) : 0u;
          const uint32_t sa = s0 & ~1u;  // groups start at even indices: 32-byte aligned
          gb[k] = sa - 4u * totg;
          gs[k] = s0;
          ge[k] = e0;
          totg += e0 > s0 ? (e0 - sa + 3u) >> 2 : 0u;
          gc[k] = totg;
        }
#pragma unroll 2
        for (uint32_t gi = 0; gi < totg; ++gi) {
          uint32_t bb = gb[0], lo = gs[0], en = ge[0];
#pragma unroll
          for (int k = 1; k < kPcaRows; ++k) {
            const bool in = gi >= gc[k - 1];
            bb = in ? gb[k] : bb;
            lo = in ? gs[k] : lo;
            en = in ? ge[k] : en;
          }
          const uint32_t i0 = bb + 4u * gi;  // >= lo - 1; only slot 0 can lie before the run, slots 1..3 behind it
          float4 q0, q1, q2, q3;
          ld_pt2(m.pts + i0, q0, q1);
          ld_pt2(m.pts + i0 + 2, q2, q3);
          PCA_POINT(q0, i0 >= lo) PCA_POINT(q1, i0 + 1 < en) PCA_POINT(q2, i0 + 2 < en) PCA_POINT(q3, i0 + 3 < en)
        }
      } else {
        for (int row = cy0; row <= cy1; ++row) {
          int ca, cb;
          row_chord(m, cx, cy, rr, row, cx0, cx1, &ca, &cb);  // chord of the radius-a disc in this row
          if (ca > cb) continue;
          const size_t rb = (size_t)row * (size_t)m.W;
          const uint32_t s0 = __ldg(m.cell_start + rb + ca), e0 = __ldg(m.cell_start + rb + cb + 1);
          if (PCA_VARIANT == 2) {
            // groups of four points from an even index: two 256-bit loads each (the point array is padded)
#pragma unroll 2
            for (uint32_t i0 = s0 & ~1u; i0 < e0; i0 += 4) {
              float4 q0, q1, q2, q3;
              ld_pt2(m.pts + i0, q0, q1);
              ld_pt2(m.pts + i0 + 2, q2, q3);
              PCA_POINT(q0, i0 >= s0) PCA_POINT(q1, i0 + 1 < e0) PCA_POINT(q2, i0 + 2 < e0) PCA_POINT(q3, i0 + 3 < e0)
            }
          } else {
#pragma unroll 4
            for (uint32_t k = s0; k < e0; ++k) {
              const float4 q0 = ld_pt(m.pts + k);
              PCA_POINT(q0, true)
            }
          }
        }
      }
#undef PCA_POINT
    }
    float weight = 0.f;
    if (st == TRGB_EDGE_COLLISION || st == TRGB_EDGE_SKIPPED) {
      npts = 0;
    } else if (nrange == 0) {
      st = TRGB_EDGE_EMPTY;   // :305
    } else if (npts < 3) {
      st = TRGB_EDGE_FEWPTS;  // :327
    } else {
      weight = weight_from_sums(npts, sx, sy, sz, sxx, sxy, sxz, syy, syz, szz);
    }
    stage_io[i] = (uint8_t)st;
    w_out[i] = weight;
    dist_out[i] = dist;
    if (npts_out) npts_out[i] = npts;
  }
}

// per-warp shared z-buffer capacity for a radius-r cylinder on this map
static double map_density(const trgb_map* m) {
  const double area = (double)m->view.W * m->view.H * (double)m->view.cell * m->view.cell;
  return (double)m->n / std::max(area, 1e-9);
}
static int pick_cap(const trgb_map* m, float r) {
  const double side = 2.0 * r + m->view.cell;
  double cand = 4.0 * map_density(m) * side * side + 64.0;
  int cap = 64;
  while (cap < cand && cap < 4096) cap <<= 1;
  return cap;
}

static int launch_cfg(const trgb_map* m, float r, int64_t n_items, int* grid, int* cap, size_t* smem,
                      const void* kernel) {
  *cap = pick_cap(m, r);
  *smem = (size_t)kWarpsPerCta * (*cap + 256) * sizeof(float);  // hits + 256 histogram bins per warp
  if (*smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(smem)", __FILE__, __LINE__);
  }
  const int per_sm = *smem > 64 * 1024 ? 1 : (*smem > 24 * 1024 ? 2 : 8);
  *grid = grid_for_warps(n_items, per_sm);
  return TRGB_OK;
}

// Thread-per-item configuration: usable when the expected cylinder population (with 60 % head
// room for clustering) fits a per-thread column of at most 96 floats. Returns false otherwise
// (large radii / dense maps -> warp-per-item kernels).
static bool tq_cfg(const trgb_map* m, float r, int64_t n_items, int* grid, int* cap, size_t* smem, const void* kernel) {
  if (m->force_warp_path) return false;
  const double k = map_density(m) * 3.14159265358979 * (double)r * r;
  if (k > 0.625 * kTqCap) return false;  // expected population must leave 60 % head room in the column
  *cap = kTqCap;
  *smem = (size_t)kTqThreads * kTqCap * sizeof(float);
  if (*smem > 48 * 1024 &&
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  const int64_t need = (n_items + kTqThreads - 1) / kTqThreads;
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(12, (200 * 1024) / std::max<size_t>(*smem, 1)));
  const int64_t cap_grid = (int64_t)sm_count() * per_sm;
  *grid = (int)std::max<int64_t>(1, std::min(need, cap_grid));
  return true;
}

}  // namespace trgb

using namespace trgb;

// ---- tier 2 -------------------------------------------------------------------------------
extern "C" int trgb_collision_launch(const trgb_map* m, const float* d_xy, int64_t n, float radius,
                                     float height_thr, float ratio_thr, uint8_t* d_out) {
  TRGB_ARG(m && d_xy && d_out, "null pointer");
  TRGB_ARG(radius > 0.f, "radius must be > 0");
  if (n <= 0) return TRGB_OK;
  int grid, cap; size_t smem;
  if (tq_cfg(m, radius, n, &grid, &cap, &smem, (const void*)k_collision_tq)) {
    ProfScope ps("k_collision", m->stream, (double)n);
    k_collision_tq<<<grid, kTqThreads, smem, m->stream>>>(m->view, reinterpret_cast<const float2*>(d_xy), n, radius,
                                                          height_thr, ratio_thr, cap, d_out);
    TRGB_CUDA(cudaGetLastError());
    return TRGB_OK;
  }
  int rc = launch_cfg(m, radius, n, &grid, &cap, &smem, (const void*)k_collision);
  if (rc) return rc;
  ProfScope ps("k_collision_warp", m->stream, (double)n);
  k_collision<<<grid, kThreads, smem, m->stream>>>(m->view, reinterpret_cast<const float2*>(d_xy), n, radius,
                                                   height_thr, ratio_thr, cap, d_out);
  TRGB_CUDA(cudaGetLastError());
  return TRGB_OK;
}

extern "C" int trgb_range_count_launch(const trgb_map* m, const float* d_xy, int64_t n, float radius,
                                       int32_t* d_out) {
  TRGB_ARG(m && d_xy && d_out, "null pointer");
  TRGB_ARG(radius > 0.f, "radius must be > 0");
  if (n <= 0) return TRGB_OK;
  ProfScope ps("k_range_count", m->stream, (double)n);
  k_range_count<<<grid_for_warps(n, 8), kThreads, 0, m->stream>>>(m->view, reinterpret_cast<const float2*>(d_xy),
                                                                   n, radius, d_out);
  TRGB_CUDA(cudaGetLastError());
  return TRGB_OK;
}

extern "C" int trgb_sample_window_launch(const trgb_map* m, const float* d_node_xy,
                                         const int32_t* d_first_draw, const float* d_draw_xy,
                                         int64_t n_nodes, int window, float radius, float height_thr,
                                         float ratio_thr, unsigned long long* d_mask) {
  return trgb_sample_window_launch2(m, d_node_xy, d_first_draw, d_draw_xy, n_nodes, window, 0.f, radius, height_thr,
                                    ratio_thr, d_mask);
}

extern "C" int trgb_sample_window_launch2(const trgb_map* m, const float* d_node_xy,
                                          const int32_t* d_first_draw, const float* d_draw_xy,
                                          int64_t n_nodes, int window, float max_offset, float radius, float height_thr,
                                          float ratio_thr, unsigned long long* d_mask) {
  TRGB_ARG(m && d_node_xy && d_first_draw && d_draw_xy && d_mask, "null pointer");
  TRGB_ARG(window >= 1 && window <= 256, "window must be in [1,256]");
  TRGB_ARG(radius > 0.f, "radius must be > 0");
  if (n_nodes <= 0) return TRGB_OK;
  int grid, cap; size_t smem;
  if (max_offset > 0.f && m->use_staging &&
      tq_cfg(m, radius, n_nodes * kTqThreads, &grid, &cap, &smem, (const void*)k_sample_window_sm)) {
    // staged variant: one CTA per node (grid-stride), neighbourhood in shared memory
    // (static 17.5 KB + dynamic 32 KB exceeds the 48 KB default: opt in explicitly)
    TRGB_CUDA(cudaFuncSetAttribute((const void*)k_sample_window_sm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_sm = 4;
    const int g2 = (int)std::max<int64_t>(1, std::min<int64_t>(n_nodes, (int64_t)sm_count() * per_sm));
    ProfScope ps("k_sample_window", m->stream, (double)n_nodes * window);
    k_sample_window_sm<<<g2, kTqThreads, smem, m->stream>>>(
        m->view, reinterpret_cast<const float2*>(d_node_xy), d_first_draw,
        reinterpret_cast<const float2*>(d_draw_xy), n_nodes, window, max_offset, radius, height_thr, ratio_thr, cap, d_mask);
    TRGB_CUDA(cudaGetLastError());
    return TRGB_OK;
  }
  if (tq_cfg(m, radius, n_nodes * window, &grid, &cap, &smem, (const void*)k_sample_window_tq)) {
    ProfScope ps("k_sample_window", m->stream, (double)n_nodes * window);
    k_sample_window_tq<<<grid, kTqThreads, smem, m->stream>>>(
        m->view, reinterpret_cast<const float2*>(d_node_xy), d_first_draw,
        reinterpret_cast<const float2*>(d_draw_xy), n_nodes, window, radius, height_thr, ratio_thr, cap, d_mask);
    TRGB_CUDA(cudaGetLastError());
    return TRGB_OK;
  }
  int rc = launch_cfg(m, radius, n_nodes * window, &grid, &cap, &smem, (const void*)k_sample_window);
  if (rc) return rc;
  ProfScope ps("k_sample_window_warp", m->stream, (double)n_nodes * window);
  k_sample_window<<<grid, kThreads, smem, m->stream>>>(
      m->view, reinterpret_cast<const float2*>(d_node_xy), d_first_draw,
      reinterpret_cast<const float2*>(d_draw_xy), n_nodes, window, radius, height_thr, ratio_thr, cap, d_mask);
  TRGB_CUDA(cudaGetLastError());
  return TRGB_OK;
}

// (the device BFS runs K3 beside K4 on a side stream: neither needs the other's output)
int trgb::nearest_z_launch_on(const trgb_map* m, const float* d_xy, int64_t n, float* d_z, int64_t* d_idx, uint8_t* d_tie,
                              const float* d_skip_d2, float skip_below, cudaStream_t st) {
  TRGB_ARG(m && d_xy, "null pointer");
  if (n <= 0) return TRGB_OK;
  ProfScope ps("k_nearest_z", st, (double)n);
  k_nearest_z<<<grid_for_warps(n, 8), kThreads, 0, st>>>(m->view, reinterpret_cast<const float2*>(d_xy), n, d_z, d_idx,
                                                         d_tie, d_skip_d2, skip_below);
  TRGB_CUDA(cudaGetLastError());
  return TRGB_OK;
}

extern "C" int trgb_nearest_z_launch_skip(const trgb_map* m, const float* d_xy, int64_t n, float* d_z,
                                          int64_t* d_idx, uint8_t* d_tie, const float* d_skip_d2, float skip_below) {
  TRGB_ARG(m, "null pointer");
  return nearest_z_launch_on(m, d_xy, n, d_z, d_idx, d_tie, d_skip_d2, skip_below, m->stream);
}

extern "C" int trgb_nearest_z_launch(const trgb_map* m, const float* d_xy, int64_t n, float* d_z,
                                     int64_t* d_idx, uint8_t* d_tie) {
  return trgb_nearest_z_launch_skip(m, d_xy, n, d_z, d_idx, d_tie, nullptr, 0.f);
}

extern "C" int trgb_edge_eval_launch(const trgb_map* m, const float* d_p1_xyz, const float* d_p2_xy,
                                     int64_t n, const TrgbEdgeParams* prm, uint8_t* d_stage,
                                     float* d_weight, float* d_dist, int32_t* d_npts) {
  return trgb_edge_eval_launch_skip(m, d_p1_xyz, d_p2_xy, n, prm, d_stage, d_weight, d_dist, d_npts, nullptr, 0, 0.f);
}

extern "C" int trgb_edge_eval_launch_skip(const trgb_map* m, const float* d_p1_xyz, const float* d_p2_xy,
                                          int64_t n, const TrgbEdgeParams* prm, uint8_t* d_stage,
                                          float* d_weight, float* d_dist, int32_t* d_npts,
                                          const float* d_skip_d2, int64_t n_skip, float skip_below) {
  if (!d_skip_d2) n_skip = 0;
  TRGB_ARG(m && d_p1_xyz && d_p2_xy && prm && d_stage && d_weight && d_dist, "null pointer");
  TRGB_ARG(prm->robot_size > 0.f, "robot_size must be > 0");
  if (n <= 0) return TRGB_OK;
  int grid, cap; size_t smem;
  // the ellipse gather reaches sqrt((1.25*expand)^2 + r^2) but keeps nothing in shared memory;
  // the shared buffer only serves the radius-robot_size collision samples
  const int kmax = prm->max_edge_samples > 0 ? prm->max_edge_samples : 8;
  if (tq_cfg(m, prm->robot_size, n * kmax, &grid, &cap, &smem, (const void*)k_edge_collide_tq)) {
    TRGB_CUDA(cudaMemsetAsync(d_stage, 0, (size_t)n, m->stream));
    {
      ProfScope ps("k_edge_collide", m->stream, (double)n);
      k_edge_collide_tq<<<grid, kTqThreads, smem, m->stream>>>(m->view, d_p1_xyz, reinterpret_cast<const float2*>(d_p2_xy),
                                                               n, kmax, prm->robot_size, prm->height_threshold,
                                                               prm->collision_threshold, cap, d_stage, d_skip_d2, n_skip,
                                                               skip_below);
    }
    {
      ProfScope ps("k_edge_pca", m->stream, (double)n);
      // lanes per edge: 8 for the small, latency-bound batches of the graph build (shorter gather
      // chain per lane), 4 at saturation (fewer idle lanes): 17 vs 22 ms per build, 1.75 vs 1.19 G edges/s
      static const int pca_mode = [] { const char* e = std::getenv("TRGB_PCA_MODE"); return e ? std::atoi(e) : 0; }();
      if (pca_mode == 1 || (pca_mode == 0 && n >= 200000)) {
        const int g3 = (int)std::max<int64_t>(1, std::min<int64_t>((n + 127) / 128, (int64_t)sm_count() * 16));
        k_edge_pca_t<<<g3, 128, 0, m->stream>>>(m->view, d_p1_xyz, reinterpret_cast<const float2*>(d_p2_xy), n,
                                                prm->robot_size, d_stage, d_weight, d_dist, d_npts, d_skip_d2, n_skip,
                                                skip_below);
        TRGB_CUDA(cudaGetLastError());
        return TRGB_OK;
      }
      const int lanes = n < 200000 ? 8 : 4;
      const int64_t need = (n + (256 / lanes) - 1) / (256 / lanes);
      const int g2 = (int)std::max<int64_t>(1, std::min<int64_t>(need, (int64_t)sm_count() * 8));
      if (lanes == 8)
        k_edge_pca<8><<<g2, 256, 0, m->stream>>>(m->view, d_p1_xyz, reinterpret_cast<const float2*>(d_p2_xy), n,
                                                 prm->robot_size, d_stage, d_weight, d_dist, d_npts, d_skip_d2, n_skip,
                                                 skip_below);
      else
        k_edge_pca<4><<<g2, 256, 0, m->stream>>>(m->view, d_p1_xyz, reinterpret_cast<const float2*>(d_p2_xy), n,
                                                 prm->robot_size, d_stage, d_weight, d_dist, d_npts, d_skip_d2, n_skip,
                                                 skip_below);
    }
    TRGB_CUDA(cudaGetLastError());
    return TRGB_OK;
  }
  int rc = launch_cfg(m, prm->robot_size, n, &grid, &cap, &smem, (const void*)k_edge_eval);
  if (rc) return rc;
  ProfScope ps("k_edge_eval_warp", m->stream, (double)n);
  k_edge_eval<<<grid, kThreads, smem, m->stream>>>(m->view, d_p1_xyz, reinterpret_cast<const float2*>(d_p2_xy), n,
                                                   prm->robot_size, prm->height_threshold,
                                                   prm->collision_threshold, cap, d_stage, d_weight, d_dist, d_npts);
  TRGB_CUDA(cudaGetLastError());
  return TRGB_OK;
}

// ---- tier 1: host buffers ------------------------------------------------------------------
namespace {
struct DevBuf {  // staging of the host-buffer tier: stream-ordered pool, no cudaMalloc / cudaFree stalls
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFreeAsync(p, 0); }
  int alloc(size_t bytes) { return cudaMallocAsync(&p, bytes ? bytes : 1, 0) == cudaSuccess && cudaStreamSynchronize(0) == cudaSuccess ? 0 : -1; }
  template <class T> T* as() { return static_cast<T*>(p); }
};
#define TRGB_ALLOC(buf, bytes) \
  if ((buf).alloc(bytes)) return cuda_fail(cudaErrorMemoryAllocation, "cudaMalloc(staging)", __FILE__, __LINE__)
}  // namespace

extern "C" int trgb_collision_batch(trgb_map* m, const float* xy, int64_t n, float radius, float height_thr,
                                    float ratio_thr, uint8_t* out) {
  TRGB_ARG(m && (n == 0 || (xy && out)), "null pointer");
  if (n <= 0) return TRGB_OK;
  DevBuf dq, dout;
  TRGB_ALLOC(dq, n * 2 * sizeof(float));
  TRGB_ALLOC(dout, n);
  TRGB_CUDA(cudaMemcpyAsync(dq.p, xy, n * 2 * sizeof(float), cudaMemcpyHostToDevice, m->stream));
  int rc = trgb_collision_launch(m, dq.as<float>(), n, radius, height_thr, ratio_thr, dout.as<uint8_t>());
  if (rc) return rc;
  TRGB_CUDA(cudaMemcpyAsync(out, dout.p, n, cudaMemcpyDeviceToHost, m->stream));
  TRGB_CUDA(cudaStreamSynchronize(m->stream));
  return TRGB_OK;
}

extern "C" int trgb_range_count_batch(trgb_map* m, const float* xy, int64_t n, float radius, int32_t* out) {
  TRGB_ARG(m && (n == 0 || (xy && out)), "null pointer");
  if (n <= 0) return TRGB_OK;
  DevBuf dq, dout;
  TRGB_ALLOC(dq, n * 2 * sizeof(float));
  TRGB_ALLOC(dout, n * sizeof(int32_t));
  TRGB_CUDA(cudaMemcpyAsync(dq.p, xy, n * 2 * sizeof(float), cudaMemcpyHostToDevice, m->stream));
  int rc = trgb_range_count_launch(m, dq.as<float>(), n, radius, dout.as<int32_t>());
  if (rc) return rc;
  TRGB_CUDA(cudaMemcpyAsync(out, dout.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, m->stream));
  TRGB_CUDA(cudaStreamSynchronize(m->stream));
  return TRGB_OK;
}

extern "C" int trgb_nearest_z_batch(trgb_map* m, const float* xy, int64_t n, float* z, int64_t* idx, uint8_t* tie) {
  TRGB_ARG(m && (n == 0 || xy), "null pointer");
  if (n <= 0) return TRGB_OK;
  DevBuf dq, dz, di, dt;
  TRGB_ALLOC(dq, n * 2 * sizeof(float));
  TRGB_ALLOC(dz, n * sizeof(float));
  TRGB_ALLOC(di, n * sizeof(int64_t));
  TRGB_ALLOC(dt, n);
  TRGB_CUDA(cudaMemcpyAsync(dq.p, xy, n * 2 * sizeof(float), cudaMemcpyHostToDevice, m->stream));
  int rc = trgb_nearest_z_launch(m, dq.as<float>(), n, dz.as<float>(), di.as<int64_t>(), dt.as<uint8_t>());
  if (rc) return rc;
  if (z) TRGB_CUDA(cudaMemcpyAsync(z, dz.p, n * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
  if (idx) TRGB_CUDA(cudaMemcpyAsync(idx, di.p, n * sizeof(int64_t), cudaMemcpyDeviceToHost, m->stream));
  if (tie) TRGB_CUDA(cudaMemcpyAsync(tie, dt.p, n, cudaMemcpyDeviceToHost, m->stream));
  TRGB_CUDA(cudaStreamSynchronize(m->stream));
  return TRGB_OK;
}

extern "C" int trgb_edge_eval_batch(trgb_map* m, const float* p1_xyz, const float* p2_xyz, int64_t n,
                                    const TrgbEdgeParams* prm, uint8_t* stage, float* weight, float* dist,
                                    int32_t* npts) {
  TRGB_ARG(m && prm && (n == 0 || (p1_xyz && p2_xyz && stage && weight && dist)), "null pointer");
  if (n <= 0) return TRGB_OK;
  std::vector<float> p2xy((size_t)n * 2);
  for (int64_t i = 0; i < n; ++i) { p2xy[2 * i] = p2_xyz[3 * i]; p2xy[2 * i + 1] = p2_xyz[3 * i + 1]; }
  DevBuf d1, d2, ds, dw, dd, dn;
  TRGB_ALLOC(d1, n * 3 * sizeof(float));
  TRGB_ALLOC(d2, n * 2 * sizeof(float));
  TRGB_ALLOC(ds, n);
  TRGB_ALLOC(dw, n * sizeof(float));
  TRGB_ALLOC(dd, n * sizeof(float));
  TRGB_ALLOC(dn, n * sizeof(int32_t));
  TRGB_CUDA(cudaMemcpyAsync(d1.p, p1_xyz, n * 3 * sizeof(float), cudaMemcpyHostToDevice, m->stream));
  TRGB_CUDA(cudaMemcpyAsync(d2.p, p2xy.data(), n * 2 * sizeof(float), cudaMemcpyHostToDevice, m->stream));
  int rc = trgb_edge_eval_launch(m, d1.as<float>(), d2.as<float>(), n, prm, ds.as<uint8_t>(), dw.as<float>(),
                                 dd.as<float>(), dn.as<int32_t>());
  if (rc) return rc;
  TRGB_CUDA(cudaMemcpyAsync(stage, ds.p, n, cudaMemcpyDeviceToHost, m->stream));
  TRGB_CUDA(cudaMemcpyAsync(weight, dw.p, n * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
  TRGB_CUDA(cudaMemcpyAsync(dist, dd.p, n * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
  if (npts) TRGB_CUDA(cudaMemcpyAsync(npts, dn.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, m->stream));
  TRGB_CUDA(cudaStreamSynchronize(m->stream));
  // host-side slope gate (trg.cpp:269-274): glibc atan2f, float overloads
  const float max_slope = atan2f(prm->height_threshold, prm->robot_size);
  for (int64_t i = 0; i < n; ++i) {
    const float dx = p1_xyz[3 * i] - p2_xyz[3 * i], dy = p1_xyz[3 * i + 1] - p2_xyz[3 * i + 1];
    const float nrm = sqrtf(dx * dx + dy * dy);
    const float slope = atan2f(fabsf(p1_xyz[3 * i + 2] - p2_xyz[3 * i + 2]), nrm);
    if (slope > max_slope) {
      stage[i] = TRGB_EDGE_SLOPE;
      weight[i] = 0.f;
      if (npts) npts[i] = 0;
    }
  }
  return TRGB_OK;
}
