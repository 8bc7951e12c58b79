// Shared device/host helpers of the sm_100a TRG kernels.
//
// Parity-critical arithmetic (set membership, segment sample positions, medians) is written
// with explicit round-to-nearest intrinsics (__fadd_rn/__fmul_rn/__fsub_rn/__fdiv_rn/
// __fsqrt_rn): they are never contracted into FMAs, so results are bit-identical to the
// reference's baseline-x86-64 float arithmetic (SURVEY.md hard part 2). The whole library is
// additionally compiled with --fmad=false.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <mutex>
#include <string>

#include "trgb_kernels.h"

namespace trgb {

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define TRGB_CUDA(expr)                                                        \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) return ::trgb::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define TRGB_ARG(cond, msg)                                   \
  do {                                                        \
    if (!(cond)) {                                            \
      ::trgb::set_error(std::string("bad argument: ") + msg); \
      return TRGB_E_ARG;                                      \
    }                                                         \
  } while (0)

// ------------------------------------------------------------------------------------------
// profiling (CUDA events on the launching stream) + launch counter
// ------------------------------------------------------------------------------------------
struct ProfScope {
  ProfScope(const char* name, cudaStream_t s, double bytes);
  ~ProfScope();
  const char* name_;
  cudaStream_t s_;
  double bytes_;
  cudaEvent_t start_ = nullptr, stop_ = nullptr;
};

// ------------------------------------------------------------------------------------------
// map view (HBM layout)
//   pts        float4[n]  (x, y, z, bit-cast original index), sorted by row-major cell id,
//                         ascending original index inside a cell (deterministic)
//   cell_start uint32[W*H+1]  exclusive prefix of points per cell; the cells of one grid row
//                         are consecutive, so a query's cell span in a row is ONE contiguous run
// ------------------------------------------------------------------------------------------
struct MapView {
  const float4* pts;
  const uint32_t* cell_start;
  float x0, y0, inv_cell, cell;
  int W, H;
  int64_t n;
};

#ifdef __CUDACC__

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int cell_coord(float v, float origin, float inv_cell, int dim) {
  // monotone non-decreasing in v: float sub, float mul, floor, clamp
  float f = floorf(__fmul_rn(__fsub_rn(v, origin), inv_cell));
  if (!(f > 0.0f)) return 0;  // also catches NaN
  if (f >= (float)dim) return dim - 1;
  return (int)f;
}

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

__device__ __forceinline__ float4 ld_pt(const float4* p) { return __ldg(p); }
// two consecutive points with one 256-bit load (sm_100: LDG.E.ENL2.256); p must be 32-byte aligned
__device__ __forceinline__ void ld_pt2(const float4* p, float4& a, float4& b) {
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
      : "l"(p));
}

// order-preserving float -> uint key (ascending)
__device__ __forceinline__ uint32_t fkey(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

// Inflated search half-width for a radius-r membership test `fl(dx^2+dy^2) <= fl(r^2)`:
// any point passing the float test has |dx| <= r*(1+4eps); the cell span is computed from
// q -+ rr with rr larger than r by more than the rounding of q -+ rr itself (which scales with
// |q|), and cell_coord is monotone, so no hit can be missed.
__device__ __forceinline__ float inflate(float r, float qx, float qy) {
  return r * 1.00001f + 1e-6f + 1e-6f * (fabsf(qx) + fabsf(qy));
}

// Cell columns [ca, cb] of grid row `row` that can hold a point within rr of (qx, qy). A point
// stored in this row has y in the row's band up to the rounding of cell_coord; the band is widened
// by `slack` (>> that rounding) before the chord half-width sqrt(rr^2 - dy^2) is taken, so the span
// is conservative. ca > cb means the row cannot contain a hit.
__device__ __forceinline__ void row_chord(const MapView& m, float qx, float qy, float rr, int row, int cx0, int cx1,
                                          int* ca, int* cb) {
  const float slack = 1e-3f * m.cell + 4e-6f * (fabsf(qy) + fabsf(m.y0));
  const float ylo = m.y0 + (float)row * m.cell - slack;
  const float yhi = m.y0 + (float)(row + 1) * m.cell + slack;
  const float dy = fmaxf(0.f, fmaxf(ylo - qy, qy - yhi));
  const float h2 = rr * rr - dy * dy;
  if (!(h2 > 0.f)) {
    *ca = 1; *cb = 0;
    return;
  }
  const float half = sqrtf(h2) * 1.0001f + 1e-6f;
  *ca = max(cx0, cell_coord(qx - half, m.x0, m.inv_cell, m.W));
  *cb = min(cx1, cell_coord(qx + half, m.x0, m.inv_cell, m.W));
}

// Warp-cooperative iteration over every point stored in the cell block [cx0..cx1] x [cy0..cy1]
// (when rr > 0: in each row only the chord of the radius-rr disc around (qx, qy)).
// f(valid, p) is called by ALL lanes each round (so it may use warp collectives); `valid` is
// false on padding lanes. Per grid row the block is one contiguous run of `pts`; lane r holds the
// run of row r, the rows' runs are concatenated by a warp scan and every lane finds the row of its
// element with a 5-step binary search over the scan (7 shuffles per 32 elements, any row count).
template <class F>
__device__ __forceinline__ void warp_for_each_in_cells(const MapView& m, int cx0, int cx1, int cy0,
                                                       int cy1, float qx, float qy, float rr, F&& f) {
  const int lane = threadIdx.x & 31;
  for (int rbase = cy0; rbase <= cy1; rbase += 32) {
    const int row = rbase + lane;
    uint32_t s = 0, e = 0;
    if (row <= cy1) {
      int ca = cx0, cb = cx1;
      if (rr > 0.f) row_chord(m, qx, qy, rr, row, cx0, cx1, &ca, &cb);
      if (ca <= cb) {
        const size_t b = (size_t)row * (size_t)m.W;
        s = __ldg(m.cell_start + b + ca);
        e = __ldg(m.cell_start + b + cb + 1);
      }
    }
    const uint32_t cnt = e - s;
    uint32_t inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(FULL, inc, d);
      if (lane >= d) inc += t;
    }
    const uint32_t total = __shfl_sync(FULL, inc, 31);
    const uint32_t exc = inc - cnt;
    for (uint32_t j0 = 0; j0 < total; j0 += 32) {
      const uint32_t j = j0 + lane;
      // first row whose inclusive prefix exceeds j (rows beyond the block have inc == total)
      int lo = 0, hi = 31;
#pragma unroll
      for (int step = 0; step < 5; ++step) {
        const int mid = (lo + hi) >> 1;
        const uint32_t v = __shfl_sync(FULL, inc, mid);
        if (j >= v) lo = mid + 1; else hi = mid;
      }
      const uint32_t sr = __shfl_sync(FULL, s, lo);
      const uint32_t er = __shfl_sync(FULL, exc, lo);
      const bool valid = j < total;
      float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) p = ld_pt(m.pts + sr + (j - er));
      f(valid, p);
    }
  }
}
template <class F>
__device__ __forceinline__ void warp_for_each_in_cells(const MapView& m, int cx0, int cx1, int cy0, int cy1, F&& f) {
  warp_for_each_in_cells(m, cx0, cx1, cy0, cy1, 0.f, 0.f, 0.f, f);
}

// ... over the cells that can hold a point within rr of (qx, qy)
template <class F>
__device__ __forceinline__ void warp_for_each_candidate(const MapView& m, float qx, float qy,
                                                        float rr, F&& f) {
  const int cx0 = cell_coord(qx - rr, m.x0, m.inv_cell, m.W);
  const int cx1 = cell_coord(qx + rr, m.x0, m.inv_cell, m.W);
  const int cy0 = cell_coord(qy - rr, m.y0, m.inv_cell, m.H);
  const int cy1 = cell_coord(qy + rr, m.y0, m.inv_cell, m.H);
  warp_for_each_in_cells(m, cx0, cx1, cy0, cy1, qx, qy, rr, f);
}

// rank-select on a warp-private buffer: value with 0-based ascending rank `k` among zbuf[0..n)
__device__ __forceinline__ float warp_select_smem(const float* zbuf, int n, int k) {
  const int lane = threadIdx.x & 31;
  if (n <= 32) {
    const float z = lane < n ? zbuf[lane] : 0.f;
    int less = 0, leq = 0;
    for (int i = 0; i < n; ++i) {
      const float zi = __shfl_sync(FULL, z, i);
      less += (zi < z);
      leq += (zi <= z);
    }
    const unsigned b = __ballot_sync(FULL, lane < n && less <= k && k < leq);
    return __shfl_sync(FULL, z, __ffs(b) - 1);
  }
  // byte-wise radix select on order-preserving keys, MSB first: 4 passes, each a 256-bin histogram
  // in shared memory (zbuf[n .. n+256) is used as the bin array: callers size the buffer for it)
  uint32_t* bins = reinterpret_cast<uint32_t*>(const_cast<float*>(zbuf)) + n;
  uint32_t prefix = 0, mask = 0;
  int kk = k;
#pragma unroll 1
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int b = lane; b < 256; b += 32) bins[b] = 0u;
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
      const uint32_t key = fkey(zbuf[i]);
      if ((key & mask) == prefix) atomicAdd(&bins[(key >> shift) & 255u], 1u);
    }
    __syncwarp();
    // lane l owns bins [8l, 8l+8): warp scan of the per-lane sums, then a short walk inside the lane
    uint32_t c[8];
    uint32_t tot = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) { c[b] = bins[8 * lane + b]; tot += c[b]; }
    uint32_t inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(FULL, inc, d);
      if (lane >= d) inc += t;
    }
    const uint32_t before = inc - tot;
    const bool mine = (uint32_t)kk >= before && (uint32_t)kk < inc;
    int digit = 0, kk_new = kk;
    if (mine) {
      uint32_t run = before;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        if ((uint32_t)kk >= run && (uint32_t)kk < run + c[b]) { digit = 8 * lane + b; kk_new = kk - (int)run; }
        run += c[b];
      }
    }
    const int src = __ffs(__ballot_sync(FULL, mine)) - 1;
    digit = __shfl_sync(FULL, digit, src);
    kk = __shfl_sync(FULL, kk_new, src);
    prefix |= (uint32_t)digit << shift;
    mask |= 255u << shift;
    __syncwarp();
  }
  return fkey_inv(prefix);
}

// TRG::isCollision (trg.cpp:746-778) for one query, executed by one warp.
// zbuf: warp-private shared buffer of `cap` floats. Returns a warp-uniform result;
// *nhits_out receives the number of points in the cylinder.
__device__ __forceinline__ bool warp_is_collision(const MapView& m, float qx, float qy, float r,
                                                  float hthr, float rthr, float* zbuf, int cap,
                                                  int* nhits_out) {
  const int lane = threadIdx.x & 31;
  const float r2 = __fmul_rn(r, r);
  const float rr = inflate(r, qx, qy);
  const unsigned lt = lanemask_lt();
  int nhits = 0;
  warp_for_each_candidate(m, qx, qy, rr, [&](bool valid, const float4& p) {
    const float dx = __fsub_rn(p.x, qx), dy = __fsub_rn(p.y, qy);
    const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    const bool hit = valid && (d2 <= r2);
    const unsigned b = __ballot_sync(FULL, hit);
    if (hit) {
      const int pos = nhits + __popc(b & lt);
      if (pos < cap) zbuf[pos] = p.z;
    }
    nhits += __popc(b);
  });
  if (nhits_out) *nhits_out = nhits;
  if (nhits == 0) return true;  // :749-752 empty cylinder => collision
  __syncwarp();
  const int mid = nhits / 2;  // :764 upper median
  int cnt = 0;
  if (nhits <= cap) {
    const float zmed = warp_select_smem(zbuf, nhits, mid);
    for (int i = lane; i < nhits; i += 32) cnt += (fabsf(__fsub_rn(zbuf[i], zmed)) > hthr);
    cnt = __reduce_add_sync(FULL, cnt);
  } else {
    // overflow path (cylinder holds more points than the shared buffer): radix select by
    // re-scanning the candidates from L2/HBM, 32 passes. Correct for any density.
    uint32_t prefix = 0, mask = 0;
    int kk = mid;
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t bm = 1u << bit;
      int c0 = 0;
      warp_for_each_candidate(m, qx, qy, rr, [&](bool valid, const float4& p) {
        const float dx = __fsub_rn(p.x, qx), dy = __fsub_rn(p.y, qy);
        const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        const uint32_t key = fkey(p.z);
        c0 += (valid && d2 <= r2 && (key & mask) == prefix && !(key & bm));
      });
      c0 = __reduce_add_sync(FULL, c0);
      if (kk >= c0) {
        kk -= c0;
        prefix |= bm;
      }
      mask |= bm;
    }
    const float zmed = fkey_inv(prefix);
    warp_for_each_candidate(m, qx, qy, rr, [&](bool valid, const float4& p) {
      const float dx = __fsub_rn(p.x, qx), dy = __fsub_rn(p.y, qy);
      const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
      cnt += (valid && d2 <= r2 && fabsf(__fsub_rn(p.z, zmed)) > hthr);
    });
    cnt = __reduce_add_sync(FULL, cnt);
  }
  __syncwarp();
  const float ratio = __fdiv_rn((float)cnt, (float)nhits);  // :773
  return ratio > rthr;
}

// ------------------------------------------------------------------------------------------
// Thread-per-query variant of TRG::isCollision (trg.cpp:746-778): one THREAD walks the cell
// runs of its query and appends the in-cylinder z values to a private shared-memory column
// zcol[k * stride] (bank-conflict free: stride is a multiple of 32, column index = lane). The
// column is then pulled into registers and ordered by a compile-time bitonic network — no
// data-dependent shared-memory chain, no divergence — and the upper median is picked at rank
// n/2. 32 queries advance per warp instruction; neighbouring threads should hold neighbouring
// queries so that the cell runs they read share L1 lines.
// Returns 0 / 1, or 2 when the cylinder holds more than kTqCap points (the caller falls back to
// the warp-cooperative routine, which needs no storage).
// ------------------------------------------------------------------------------------------
// Build-time knobs of the thread-per-query kernels (scripts/build_variants.sh sweeps them on the GPU;
// the defaults are the measured optimum, see profiles/README.md).
#ifndef TQ_CAP
#define TQ_CAP 64
#endif
#ifndef TQ_FLAT
#define TQ_FLAT 1
#endif
#ifndef TQ_LD256
#define TQ_LD256 1  // 256-bit point loads (two points each) in the grouped gather
#endif
#ifndef TQ_UNROLL
#define TQ_UNROLL 2  // groups of four loads in flight per thread
#endif
#ifndef TQ_MINBLK
#define TQ_MINBLK 6  // resident CTAs per SM asked of ptxas (80 registers; 0 = its own heuristic)
#endif
#if TQ_MINBLK > 0
#define TQ_BOUNDS __launch_bounds__(128, TQ_MINBLK)
#else
#define TQ_BOUNDS __launch_bounds__(128)
#endif
#define TQ_PRAGMA_(x) _Pragma(#x)
#define TQ_PRAGMA_UNROLL(n) TQ_PRAGMA_(unroll n)
constexpr int kTqCap = TQ_CAP;  // per-thread column length (floats)

template <int N>
__device__ __forceinline__ void bitonic_sort_regs(float (&a)[N]) {
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int i = 0; i < N; ++i) {
        const int l = i ^ j;
        if (l > i) {
          const float lo = fminf(a[i], a[l]), hi = fmaxf(a[i], a[l]);
          if ((i & k) == 0) { a[i] = lo; a[l] = hi; }
          else { a[i] = hi; a[l] = lo; }
        }
      }
    }
  }
}

template <int N>
__device__ __forceinline__ int median_outlier_count(const float* zcol, int stride, int n, float hthr) {
  // The n values are padded to N with -inf / +inf alternating from index n on: ceil((N-n)/2) pads
  // sort below the data and floor((N-n)/2) above it, which puts the upper median (rank n/2 of the
  // data, :764) at index N/2 of the sorted array whatever n is. Only that one output is read, so
  // the compiler prunes the last merge of the network to the 31 (N = 32) half compare-exchanges
  // that feed it - a selection network - and the outliers are counted over the column itself.
  const float pad_even = (n & 1) ? INFINITY : -INFINITY;  // pad at an even index k >= n
  const float pad_odd = -pad_even;
  float a[N];
#pragma unroll
  for (int k = 0; k < N; ++k) a[k] = (k < n) ? zcol[k * stride] : ((k & 1) ? pad_odd : pad_even);
  bitonic_sort_regs<N>(a);
  const float zmed = a[N / 2];
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < N; ++k) cnt += (k < n && fabsf(__fsub_rn(zcol[k * stride], zmed)) > hthr);
  return cnt;
}

// kTqRows: grid rows the flattened gather below holds in registers. A radius-r disc spans
// floor(2 rr / cell) + 1 or + 2 rows; the map is built with cell ~ 0.67 r, i.e. 4 - 5 rows.
#ifndef TQ_ROWS
#define TQ_ROWS 5
#endif
constexpr int kTqRows = TQ_ROWS;

__device__ __forceinline__ int thread_is_collision(const MapView& m, float qx, float qy, float r, float hthr,
                                                   float rthr, float* zcol, int stride) {
  const float r2 = __fmul_rn(r, r);
  const float rr = inflate(r, qx, qy);
  const int cx0 = cell_coord(qx - rr, m.x0, m.inv_cell, m.W);
  const int cx1 = cell_coord(qx + rr, m.x0, m.inv_cell, m.W);
  const int cy0 = cell_coord(qy - rr, m.y0, m.inv_cell, m.H);
  const int cy1 = cell_coord(qy + rr, m.y0, m.inv_cell, m.H);
  int n = 0;
  if (TQ_FLAT && cy1 - cy0 < kTqRows) {
    // Flattened gather: the rows' runs (only the chord of the inflated disc each row of cells can
    // contain: ~20 % fewer candidates) are located first - all cell_start loads in flight at once -
    // and then walked as ONE loop over groups of four consecutive points starting at an even index
    // (a group may cover one point before its run and up to three behind it - masked; the point
    // array is padded), fetched with two 256-bit loads: the kernel is bound by L1 tag look-ups, one
    // per distinct line a warp-wide load touches, and wide loads halve their number. Lanes of
    // a warp differ in how their candidates split over rows but hardly in the total, so the loop
    // runs with most lanes active, where a loop per row idles every lane whose run is shorter than
    // the longest one in the warp. Group g belongs to the row k whose cumulative group count first
    // exceeds g and starts at pts[gb[k] + 4 g]; one address per group and a
    // branch-free body (predicated stores at a running column offset).
    uint32_t gb[kTqRows], gs[kTqRows], ge[kTqRows], gc[kTqRows];
    uint32_t totg = 0;
    const float slack = 1e-3f * m.cell + 4e-6f * (fabsf(qy) + fabsf(m.y0));
#pragma unroll
    for (int k = 0; k < kTqRows; ++k) {
      const int row = cy0 + k;
      // row_chord() without its branches: same conservative band and chord (the approximate
      // square root is off by ~1e-7 relative, the chord is widened by 1e-4), clamped to the
      // query's cell block, and an empty span for rows past cy1 or outside the disc
      const float ylo = m.y0 + (float)row * m.cell - slack;
      const float yhi = m.y0 + (float)(row + 1) * m.cell + slack;
      const float dy = fmaxf(0.f, fmaxf(ylo - qy, qy - yhi));
      const float h2 = rr * rr - dy * dy;
      float half;
      asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(half) : "f"(fmaxf(h2, 0.f)));
      half = half * 1.0001f + 1e-6f;
      const int ca = max(cx0, min(m.W - 1, __float2int_rd(__fmul_rn(__fsub_rn(qx - half, m.x0), m.inv_cell))));
      const int cb = min(cx1, max(0, __float2int_rd(__fmul_rn(__fsub_rn(qx + half, m.x0), m.inv_cell))));
      const bool any = row <= cy1 && h2 > 0.f && ca <= cb;
      const uint32_t* cs = m.cell_start + (size_t)min(row, cy1) * (size_t)m.W;
      const uint32_t s = any ? __ldg(cs + ca) : 0u;
      const uint32_t e = any ? __ldg(cs + cb + 1) : 0u;
      const uint32_t sa = TQ_LD256 ? s & ~1u : s;  // groups start at even indices: 32-byte aligned
      gb[k] = sa - 4u * totg;
      gs[k] = s;
      ge[k] = e;
      totg += e > s ? (e - sa + 3u) >> 2 : 0u;
      gc[k] = totg;
    }
    int wo = 0;                                // column offset of the next store (floats)
    const int wlim = (kTqCap - 4) * stride;    // checked once per group: at most 4 stores past it
    bool ovf = false;
    TQ_PRAGMA_UNROLL(TQ_UNROLL)
    for (uint32_t g = 0; g < totg; ++g) {
      uint32_t b = gb[0], lo = gs[0], e = ge[0];
#pragma unroll
      for (int k = 1; k < kTqRows; ++k) {
        const bool in = g >= gc[k - 1];
        b = in ? gb[k] : b;
        lo = in ? gs[k] : lo;
        e = in ? ge[k] : e;
      }
      const uint32_t i0 = b + 4u * g;  // >= lo - 1; only slot 0 can lie before the run, slots 1..3 behind it
      float4 p0, p1, p2, p3;
      if (TQ_LD256) {
        ld_pt2(m.pts + i0, p0, p1);
        ld_pt2(m.pts + i0 + 2, p2, p3);
      } else {
        p0 = ld_pt(m.pts + i0), p1 = ld_pt(m.pts + i0 + 1), p2 = ld_pt(m.pts + i0 + 2), p3 = ld_pt(m.pts + i0 + 3);
      }
#define TQ_CAND(P, U)                                                              \
      {                                                                            \
        const float dx = __fsub_rn(P.x, qx), dy = __fsub_rn(P.y, qy);              \
        const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));          \
        if (d2 <= r2 && (U == 0 ? i0 >= lo : i0 + U < e)) {                        \
          zcol[wo] = P.z;                                                          \
          wo += stride;                                                            \
        }                                                                          \
      }
      TQ_CAND(p0, 0) TQ_CAND(p1, 1) TQ_CAND(p2, 2) TQ_CAND(p3, 3)
#undef TQ_CAND
      // (no early exit: a second way out of this loop keeps the warp from reconverging before the
      //  sorting network below, which then runs once per distinct trip count)
      ovf |= wo > wlim;
      wo = min(wo, wlim);
    }
    if (ovf) return 2;  // the column overflowed: the caller's fallback needs no storage
    n = wo / stride;
  } else {
    for (int row = cy0; row <= cy1; ++row) {
      int ca, cb;
      row_chord(m, qx, qy, rr, row, cx0, cx1, &ca, &cb);
      if (ca > cb) continue;
      const size_t b = (size_t)row * (size_t)m.W;
      const uint32_t s = __ldg(m.cell_start + b + ca);
      const uint32_t e = __ldg(m.cell_start + b + cb + 1);
#pragma unroll 4
      for (uint32_t i = s; i < e; ++i) {
        const float4 p = ld_pt(m.pts + i);
        const float dx = __fsub_rn(p.x, qx), dy = __fsub_rn(p.y, qy);
        const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        if (d2 <= r2) {
          if (n < kTqCap) zcol[n * stride] = p.z;
          ++n;
        }
      }
    }
  }
  if (n == 0) return 1;  // :749-752 empty cylinder => collision
  if (n > kTqCap) return 2;
  // the whole warp takes the short network when every lane fits it (keeps the branch uniform)
  int cnt;
  if (__all_sync(__activemask(), n <= 32)) cnt = median_outlier_count<32>(zcol, stride, n, hthr);
  else cnt = median_outlier_count<64>(zcol, stride, n, hthr);
  const float ratio = __fdiv_rn((float)cnt, (float)n);  // :773
  return ratio > rthr ? 1 : 0;
}

#endif  // __CUDACC__

// shared launch geometry: 8 warps per CTA, one query per warp at a time, grid sized to a
// multiple of the SM count (persistent-style grid-stride loop)
constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;
// thread-per-query kernels: 128 threads per CTA, one shared z column per thread
constexpr int kTqThreads = 128;
constexpr int kPtsPad = 4;  // float4 slots allocated (zeroed) behind the sorted points of a map
// radix sorts / scan, instantiated once (sort.cu)
int sort_pairs_u32_u32(const uint32_t* kin, uint32_t* kout, const uint32_t* vin, uint32_t* vout, int n, int end_bit, cudaStream_t st);
int sort_pairs_u32_u64(const uint32_t* kin, uint32_t* kout, const unsigned long long* vin, unsigned long long* vout, int n, int end_bit, cudaStream_t st);
int sort_pairs_u64_u32(const unsigned long long* kin, unsigned long long* kout, const uint32_t* vin, uint32_t* vout, int n, int end_bit, cudaStream_t st);
int exclusive_sum_u32(const uint32_t* in, uint32_t* out, int n, cudaStream_t st);
bool prof_enabled();                 // the CUDA-event profiler is recording (core.cu)
void count_launches(int64_t n);      // kernels launched through a replayed CUDA graph (core.cu)
// K7 search graph from CSR arrays resident on the device (sssp.cu). Source numbering 0 .. n_src-1; src2ext[i] =
// the caller's id of source node i or -1 (node dropped), nullptr = identity. Exactly one of xy / xyz and of
// state8 / state32 is set.
struct GraphSource {
  int32_t n_src = 0, n_keep = 0, n_ext = 0;
  int64_t e_src = 0;
  const long long* row = nullptr;
  const int32_t* col = nullptr;
  const float* w = nullptr;
  const float* dist = nullptr;
  const float2* xy = nullptr;
  const float* xyz = nullptr;
  const signed char* state8 = nullptr;
  const int32_t* state32 = nullptr;
  const int32_t* src2ext = nullptr;
};
int graph_from_device(trgb_graph** out, const GraphSource& src, cudaStream_t build_stream);

int nearest_z_launch_on(const trgb_map* m, const float* d_xy, int64_t n, float* d_z, int64_t* d_idx, uint8_t* d_tie,
                        const float* d_skip_d2, float skip_below, cudaStream_t st);  // K3 on a stream of the caller's (queries.cu)
void tune_mempool_once();  // raise the default mempool's release threshold (map_index.cu)
int sm_count();
int grid_for_warps(int64_t n_warps, int ctas_per_sm);

}  // namespace trgb

// opaque handle bodies
struct trgb_map {
  trgb::MapView view{};
  float4* d_pts = nullptr;
  uint32_t* d_cell_start = nullptr;
  int64_t n = 0;
  int64_t device_bytes = 0;
  cudaStream_t stream = nullptr;
  int zcap = 128;  // per-warp shared z buffer (floats) for the collision kernels
  int force_warp_path = 0;  // tests: route every launch through the warp-per-item kernels
  int use_staging = 0;      // shared-memory staged sampling-window kernel (measured slower than the L1 path: off)
  // pinned/device staging for the host-buffer tier
  void* h_stage = nullptr;
  size_t h_stage_bytes = 0;
  void* d_stage = nullptr;
  size_t d_stage_bytes = 0;
};
