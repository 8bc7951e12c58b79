// K9 — device-resident graph expansion: TRG::expandGraph (trg.cpp:372-454) as a BFS that lives on the
// GPU, decisions included. The host only feeds the sampling stream and polls a status block.
//
// The reference pops one node at a time: draw angles until `sample_num` collision-free samples
// (:384-403), then for each sample look up the nearest EXISTING node and either wire to it or insert
// a new node (:406-452). Every decision depends on the nodes inserted by earlier samples, so the
// loop looks sequential. It is not: a sample's outcome depends only on nodes closer to it than its
// parent (expand_dist away), and within one BFS frontier the chains of "my outcome depends on an
// earlier, still undecided sample next to me" are short (5 - 8 links). One STEP of this engine
// handles up to a few thousand queued pops at once:
//
//   (plan)         pops of this step, their guessed stream positions (running mean / variance): done by
//                  the commit kernel of the step before (k_exp_plan for the first step of a queued group)
//   k_exp_window   collision bit of every (pop, draw) in a window around the guess   [K2 device code]
//   k_exp_tables   per pop: draws consumed as a function of the start offset; composed per 32 pops
//   k_exp_top      the exact chain o_{i+1} = o_i + consumed_i(o_i) over the composed blocks
//   k_exp_emit     accepted samples of every pop whose chain position is now known + their nearest node
//                  among those that existed before the step                             [K5 grid]
//   k_exp_deps     per sample: the few earlier samples of the step whose fate can change its own
//   K3 + K4        height and parent edge of every sample that may become a node (existing kernels)
// (one captured CUDA graph per launch size replays these; k_exp_deps runs beside K3 / K4)
//   k_exp_commit   deterministic-reservation rounds: a sample is decided once no earlier undecided
//                  sample that could still become a node lies within its current nearest distance;
//                  decided samples are applied in (pop, sample) order: node ids by prefix sum, queue
//                  pushes, wire requests — exactly the reference's sequential outcome
//
// Edges to existing nodes do not feed back into the expansion when step 3 of expandGraph is off
// (trg.cpp:429, the mountain configuration): they are recorded as requests and evaluated in one
// saturated K4 launch at the end (trgb_expander_finalize), then grouped into adjacency lists in
// request order by two radix sorts.
//
// Exactness hazards are never guessed: an exact distance tie between two nodes, or a slope-gate
// decision within 3 ulp of the threshold (the reference uses glibc's atan2f), interrupts the engine
// at that pop; the host library handles that single pop with the reference's tie / libm rules and
// resumes (trgb_expander_apply_pop).
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace trgb {

constexpr int kExpMaxPops = 8192;     // pops per step (upper bound of TrgbExpandParams::max_pops)
constexpr int kExpMaxS = 32;          // sample_num upper bound of the device engine
constexpr int kExpSlowWords = 17;     // slow mode: one pop, 1088-bit window (>= 1001 + sample_num draws)
constexpr int kExpHash = 8192;        // buckets of the commit kernel's hash grid (shared memory)

enum { EXP_INT_NONE = 0, EXP_INT_TIE = 1, EXP_INT_SLOPE = 2, EXP_INT_CAPACITY = 3, EXP_INT_DRAWS = 4 };

// sample decision states
enum : int { ST_UNDECIDED = 0, ST_SKIP = 1, ST_WIRE = 2, ST_CREATE = 3, ST_VOID = 4 };

struct ExpCtl {
  // persistent
  int head, tail;          // BFS queue [head, tail)
  int n_nodes;
  int interrupt, interrupt_pop;  // interrupt_pop: queue index of the pop the host must handle
  int stuck;               // consecutive steps whose chain could not finish the first pop
  long long n_req;
  long long pos;           // absolute stream position of the next draw
  long long draws_base, draws_end;
  float mean, var;         // running draws / pop statistics
  // step scratch
  int m, W, words, D, slow;
  int n_done;              // pops of this step whose chain position is known
  int n_commit;
  int ipop;                // first pop of the step the host must handle (tie / slope), INT_MAX = none
  int blocks_done;         // CTAs of k_exp_tables that have finished (the last one walks the chain)
  int und[2][16];          // undecided samples per commit CTA, double-buffered by round parity
  unsigned long long cta_tot[16];  // packed (created, queued, requests) totals per commit CTA
  long long pos0, pos_done;
  // statistics
  long long window_tests, steps, steps_active, rounds, pops, z_ties, redo_pops, samples, created, pc_samples;
};

struct ExpView {
  ExpCtl* ctl;
  // graph nodes
  float2* node_xy; float* node_z; signed char* node_state; int* queue;
  int node_cap;
  // K5 grid over nodes
  int* ghead; int* gnext; float gx0, gy0, gcell, ginv; int GW, GH;
  // requests
  int* req_a; int* req_b; float* req_w; float* req_d; long long req_cap;
  // stream
  const float2* draws;
  // step scratch
  int* g_off; unsigned long long* mask; unsigned char* ctab; int* blk_end; unsigned char* blk_stop; int* blk_start;
  int* pop_off; unsigned short* pop_cons; unsigned char* pop_acc;
  float2* s_xy; float* s_p1; int* s_nn; float* s_d2; unsigned char* s_tie;
  float* s_z; unsigned char* s_ztie; unsigned char* s_stage; float* s_w; float* s_d;
  int* st; float* cur_d2; int* cur_nn; int* hnext; unsigned char* s_pc;
  int* dhead; int* dnext; int* dep_j; float* dep_d2; int* dep_n;  // per-step hash of potential creators + dependency lists
  // params
  float e, r, hthr, cthr, max_slope; int S, C, norm_words, new_state;
};

__device__ __forceinline__ int exp_ncell(float v, float origin, float inv, int dim) {
  const float f = floorf(__fmul_rn(__fsub_rn(v, origin), inv));
  if (!(f > 0.0f)) return 0;
  if (f >= (float)dim) return dim - 1;
  return (int)f;
}

// ------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------
// Plan of the next step, executed by one whole thread block (any size): pops taken, window geometry,
// guessed stream positions, zeroed masks.
__device__ void exp_plan_block(const ExpView& v, int c_step, const ExpCtl* current = nullptr) {
  ExpCtl* c = v.ctl;
  __shared__ int s_m, s_words, s_D, s_fit;
  __shared__ float s_mean, s_var;
  __shared__ long long s_room;
  __syncthreads();
  ExpCtl k;  // thread 0's copy: one read (or the caller's shared copy, `current`), one write-back at the very end
  if (threadIdx.x == 0) {
    k = current ? *current : *c;
    int m = 0;
    k.slow = 0;
    k.n_done = 0;
    k.n_commit = 0;
    k.ipop = 0x7fffffff;
    k.blocks_done = 0;
    int words = v.norm_words;
    if (!k.interrupt && k.head < k.tail) {
      const int avail = k.tail - k.head;
      if (k.stuck >= 1) {  // the first pop needs more draws than a normal window covers
        m = 1;
        words = kExpSlowWords;
        k.slow = 1;
      } else {
        const int D = 64 * words - 32;
        // longest chain whose drift (random walk, variance `var` per pop) stays inside the admissible
        // start offsets with ~2 sigma: 2 + 2 sqrt(L var) <= D / 2
        const float half = 0.5f * (float)D - 2.f;
        int L = (int)((half * half * 0.25f) / fmaxf(k.var, 0.05f));
        L = max(L, 64);
        m = min(min(avail, c_step), min(L, v.C));
      }
      k.W = 64 * words; k.words = words; k.D = k.slow ? 1 : 64 * words - 32;
      k.pos0 = k.pos;
    }
    s_m = m; s_words = words; s_D = k.D; s_fit = m;
    s_mean = k.mean; s_var = k.var;
    s_room = k.draws_end - k.pos;  // draws available from pos0 on
  }
  __syncthreads();
  // guesses (relative to pos0): exact for pop 0, extrapolated with the running mean minus a lead that
  // grows like the random walk for the others; never below the S draws every pop consumes. The step
  // keeps the pops whose whole window lies inside the draws pushed so far.
  const int m0 = s_m, W = 64 * s_words, D = s_D;
  const int smin = min(v.S, 1001);
  for (int i = threadIdx.x; i < m0; i += blockDim.x) {
    int g = 0;
    if (i > 0) {
      const float lead = fminf(2.f + 2.f * sqrtf((float)i * s_var), 0.5f * (float)D);
      g = max(i * smin, (int)floorf((float)i * s_mean - lead));
    }
    v.g_off[i] = g;
    if ((long long)g + W > s_room) atomicMin(&s_fit, i);  // guesses are non-decreasing in i
  }
  __syncthreads();
  const int m = s_fit;  // 0: the host must push draws first (it sees pos / draws_end in the status block)
  if (threadIdx.x == 0) {
    k.m = m;
    if (m > 0) k.steps_active++;
    *c = k;
  }
  const int n = m * s_words;
  for (int k = threadIdx.x; k < n; k += blockDim.x) v.mask[k] = 0ull;
}

// stand-alone plan: first step of a queued group, sized for that group's launches (the other steps are
// planned by the commit kernel before them; planning twice is harmless)
__global__ void __launch_bounds__(256) k_exp_plan(ExpView v, int c_step) {
  exp_plan_block(v, c_step);
}

// ------------------------------------------------------------------------------------------
// sampling windows: one thread per (pop, draw); bit j of mask[pop] = isCollision(node + draw[g + j])
// (trg.cpp:395-398). Same per-thread routine as k_sample_window_tq.
// ------------------------------------------------------------------------------------------
#define EXP_FALLBACK_BEGIN(res)                                \
  if (__syncthreads_or((res) == 2)) {                          \
    float* wbuf = zsm + (threadIdx.x >> 5) * 32 * cap;         \
    unsigned ov = __ballot_sync(FULL, (res) == 2);             \
    while (ov) {                                               \
      const int src = __ffs(ov) - 1;                           \
      ov &= ov - 1;
#define EXP_FALLBACK_END \
    }                    \
    __syncthreads();     \
  }

__global__ void TQ_BOUNDS k_exp_window(MapView m, ExpView v, int cap) {
  extern __shared__ float zsm[];
  float* zcol = zsm + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const ExpCtl* c = v.ctl;
  const int np = c->m;
  if (np == 0) return;
  const int W = c->W, words = c->words;
  const int items = np * W;
  const long long dbase = c->pos0 - c->draws_base;
  const int head = c->head;
  for (int base = blockIdx.x * kTqThreads; base < items; base += gridDim.x * kTqThreads) {
    const int it = base + threadIdx.x;
    float sx = 0.f, sy = 0.f;
    int res = 0, pop = 0, j = 0;
    if (it < items) {
      pop = it / W;
      j = it - pop * W;
      const float2 np2 = v.node_xy[v.queue[head + pop]];
      const float2 d = __ldg(v.draws + (dbase + v.g_off[pop] + j));
      sx = __fadd_rn(np2.x, d.x);  // trg.cpp:396-397
      sy = __fadd_rn(np2.y, d.y);
      res = thread_is_collision(m, sx, sy, v.r, v.hthr, v.cthr, zcol, kTqThreads);
    }
    EXP_FALLBACK_BEGIN(res)
      const float qx = __shfl_sync(FULL, sx, src), qy = __shfl_sync(FULL, sy, src);
      const bool cc = warp_is_collision(m, qx, qy, v.r, v.hthr, v.cthr, wbuf, 32 * cap - 256, nullptr);
      if (lane == src) res = cc ? 1 : 0;
    EXP_FALLBACK_END
    if (it < items && res) atomicOr(v.mask + (size_t)pop * words + (j >> 6), 1ull << (j & 63));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd((unsigned long long*)&v.ctl->window_tests, (unsigned long long)items);
}

// ---- bit helpers on a pop's window (collision bits; a free draw is a ZERO bit) -------------------------
// draws consumed by a pop that starts at bit `r` of its window: S collision-free draws (trg.cpp:387-403;
// the trial cap of :388 cannot trigger inside a normal window of <= 256 draws). 255 = the window ends first.
__device__ __forceinline__ int consumed_from(const unsigned long long (&mk)[4], int words, int r, int S) {
  int w = r >> 6;
  unsigned long long cur = ~mk[w] & (~0ull << (r & 63));
  int need = S;
  while (true) {
    const int pc = __popcll(cur);
    if (pc >= need) {
      for (int k = 1; k < need; ++k) cur &= cur - 1;
      return min(255, w * 64 + __ffsll((long long)cur) - r);  // index of the S-th free draw + 1 - r
    }
    need -= pc;
    if (++w >= words) return 255;
    cur = ~mk[w];
  }
}
// number of free draws below bit b
__device__ __forceinline__ int free_below(const unsigned long long (&mk)[4], int b) {
  const int w = b >> 6;
  int n = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < w) n += __popcll(~mk[k]);
  if (w < 4) n += __popcll(~mk[w] & ((1ull << (b & 63)) - 1ull));
  return n;
}

constexpr int kRow = 228;  // bytes per table row in shared memory
constexpr int kWarpSm = 32 * kRow + 256;  // per warp: 32 table rows + the free-draw positions of one pop
constexpr int kDepMax = 16;    // earlier potential creators kept per sample (more: full-scan fallback)
constexpr int kDepHash = 32768;  // buckets of the per-step hash over potential creators (global memory)

__device__ __forceinline__ unsigned dep_hash(int cx, int cy) {
  return ((unsigned)cx * 73856093u ^ (unsigned)cy * 19349663u) & (kDepHash - 1);
}

// ---- chain kernels --------------------------------------------------------------------------------------
// tables: one CTA (32 warps) per block of 32 pops, one pop per warp. All 32 lanes of a warp work on one
// pop at a time: positions of its free draws, then consumed(start offset) for every admissible offset
// (the S-th free draw at or after r is free draw number free_below(r) + S - 1). Afterwards thread r
// composes the block's map (start offset at its first pop -> end position, pops completed). The CTA
// that finishes last walks the chain over all blocks (exp_top_phase).
__device__ __forceinline__ void exp_tables_block(const ExpView& v, unsigned char* smem, int b) {
  const ExpCtl* c = v.ctl;
  const int m = c->m;
  const int words = c->words, D = c->D, S = v.S, W = c->W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* t = smem;                            // 32 rows of kRow bytes
  unsigned char* zp = smem + 32 * kRow + warp * 256;  // free-draw positions of the pop a warp works on
  const int first = b << 5;
  const int cnt = min(32, m - first);
  for (int p = warp; p < cnt; p += (int)(blockDim.x >> 5)) {
    unsigned long long mk[4] = {~0ull, ~0ull, ~0ull, ~0ull};
    for (int w = 0; w < words; ++w) mk[w] = v.mask[(size_t)(first + p) * words + w];
    for (int bpos = lane; bpos < W; bpos += 32)
      if (!((mk[bpos >> 6] >> (bpos & 63)) & 1ull)) zp[free_below(mk, bpos)] = (unsigned char)bpos;
    int nz = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) nz += __popcll(~mk[w]);
    __syncwarp();
    for (int r = lane; r < D; r += 32) {
      const int idx = free_below(mk, r) + S - 1;
      const int cc = idx < nz ? min(255, (int)zp[idx] + 1 - r) : 255;
      v.ctab[(size_t)(first + p) * 256 + r] = (unsigned char)cc;
      t[p * kRow + r] = (unsigned char)cc;
    }
    __syncwarp();
  }
  __shared__ int s_gl[32];
  if (threadIdx.x < 32) s_gl[threadIdx.x] = threadIdx.x < cnt ? v.g_off[first + threadIdx.x] : 0;
  __syncthreads();
  for (int r = threadIdx.x; r < D; r += blockDim.x) {
    int a = s_gl[0] + r;  // relative to pos0
    int stop = cnt;
    for (int p = 0; p < cnt; ++p) {
      const int rr = a - s_gl[p];
      if (rr < 0 || rr >= D) { stop = p; break; }
      const int cc = t[p * kRow + rr];
      if (cc == 255) { stop = p; break; }
      a += cc;
    }
    v.blk_end[(size_t)b * 256 + r] = a;
    v.blk_stop[(size_t)b * 256 + r] = (unsigned char)stop;
  }
}

// top (one CTA): the chain over the composed blocks, staged in shared memory and walked by one
// thread; or the single pop of slow mode
__device__ __forceinline__ void exp_top_phase(const ExpView& v, unsigned char* smem) {
  ExpCtl* c = v.ctl;
  const int m = c->m;
  const int S = v.S;
  if (c->slow) {
    if (threadIdx.x != 0) return;
    // sequential walk with the trial cap (trg.cpp:387-403): `if (trial_sample > 1000) break` is
    // checked before every draw
    const unsigned long long* mk = v.mask;
    int acc = 0, trials = 0, j = 0;
    const int W = c->W;
    bool ok = true;
    while (acc < S) {
      if (trials > 1000) break;
      if (j >= W) { ok = false; break; }
      const bool coll = (mk[j >> 6] >> (j & 63)) & 1ull;
      ++j;
      if (coll) ++trials; else ++acc;
    }
    if (!ok) {  // cannot happen: W >= 1001 + S
      c->n_done = 0;
      return;
    }
    v.pop_off[0] = 0;
    v.pop_cons[0] = (unsigned short)j;
    v.blk_start[0] = 0;
    c->n_done = 1;
    c->pos_done = c->pos0 + j;
    return;
  }
  const int D = c->D;
  const int nblk = (m + 31) >> 5;
  int* s_end = reinterpret_cast<int*>(smem);                                   // nblk * D
  unsigned char* s_stop = reinterpret_cast<unsigned char*>(s_end + nblk * D);  // nblk * D
  for (int k = threadIdx.x; k < nblk * D; k += blockDim.x) {
    const int b = k / D, r = k - b * D;
    s_end[k] = __ldcg(v.blk_end + (size_t)b * 256 + r);   // written by other CTAs of this launch: L2 is the coherence point
    s_stop[k] = __ldcg(v.blk_stop + (size_t)b * 256 + r);
  }
  __shared__ int s_g[kExpMaxPops / 32];
  for (int b = threadIdx.x; b < nblk; b += blockDim.x) s_g[b] = v.g_off[b << 5];
  __syncthreads();
  if (threadIdx.x != 0) return;
  int a = 0, done = 0;
  for (int b = 0; b < nblk; ++b) {
    const int first = b << 5;
    const int rr = a - s_g[b];
    if (rr < 0 || rr >= D) break;
    v.blk_start[b] = a;
    const int cnt = min(32, m - first);
    const int stop = s_stop[b * D + rr];
    a = s_end[b * D + rr];
    done = first + stop;
    if (stop < cnt) break;
  }
  c->n_done = done;
  c->pos_done = c->pos0 + a;
}

// nearest node (among those that existed before the step) of one sample; a sample that may become a node
// (nearest node at least robot_size away, trg.cpp:414) is entered into the step's hash
__device__ __forceinline__ void exp_nearest_one(const ExpView& v, int s, float hinv) {
  const float2 p = v.s_xy[s];
  float bd = INFINITY;
  int bi = -1, tie = 0;
  const int qcx = exp_ncell(p.x, v.gx0, v.ginv, v.GW), qcy = exp_ncell(p.y, v.gy0, v.ginv, v.GH);
  const float fuzz = 4e-6f * (fabsf(p.x) + fabsf(p.y) + v.gcell * (float)(v.GW + v.GH));
  const int maxr = max(v.GW, v.GH);
  auto scan = [&](int e) {
    for (; e >= 0; e = v.gnext[e]) {
      const float2 q = v.node_xy[e];
      const float dx = __fsub_rn(q.x, p.x), dy = __fsub_rn(q.y, p.y);
      const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
      if (d2 < bd) { bd = d2; bi = e; tie = 0; }
      else if (d2 == bd && e != bi) tie = 1;
    }
  };
  {  // the 3 x 3 block: all nine list heads are fetched before any list is walked
    int hd[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int xx = qcx + (k % 3) - 1, yy = qcy + (k / 3) - 1;
      hd[k] = (xx < 0 || yy < 0 || xx >= v.GW || yy >= v.GH) ? -1 : v.ghead[(size_t)yy * v.GW + xx];
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) scan(hd[k]);
  }
  for (int R = 1; R <= maxr; ++R) {
    if (R > 1) {
      for (int yy = qcy - R; yy <= qcy + R; ++yy) {
        if (yy < 0 || yy >= v.GH) continue;
        const bool edge_row = (yy == qcy - R || yy == qcy + R);
        for (int xx = qcx - R; xx <= qcx + R; ++xx) {
          if (xx < 0 || xx >= v.GW) continue;
          if (!edge_row && xx != qcx - R && xx != qcx + R) continue;  // ring only
          scan(v.ghead[(size_t)yy * v.GW + xx]);
        }
      }
    }
    // cells outside the scanned block are at least R whole cells away from the query's cell
    const float g = (float)R * v.gcell * 0.9999f - fuzz;
    if (bi >= 0 && g > 0.f && bd <= g * g) break;
    if (qcx - R <= 0 && qcy - R <= 0 && qcx + R >= v.GW - 1 && qcy + R >= v.GH - 1) break;
  }
  v.s_d2[s] = bd;
  v.s_nn[s] = bi;
  v.s_tie[s] = (unsigned char)tie;
  const bool pc = __fsqrt_rn(bd) >= v.r;
  v.s_pc[s] = pc ? 1 : 0;
  if (pc) {
    const unsigned h = dep_hash((int)floorf((p.x - v.gx0) * hinv), (int)floorf((p.y - v.gy0) * hinv));
    v.dnext[s] = atomicExch(v.dhead + h, s);
  }
}

// One CTA per block of 32 pops. Warp 0 walks the block again from its known start (lane p owns the window
// of pop p and answers for it) and writes the accepted samples; then every thread of the CTA takes
// samples of the block and looks up their nearest existing node. Slots without a sample get d2 = 0,
// which makes the K3 / K4 launches skip them.
__global__ void __launch_bounds__(256) k_exp_emit(ExpView v) {
  const ExpCtl* c = v.ctl;
  const int n_done = c->n_done;
  const int words = c->words, S = v.S;
  const int b = blockIdx.x;
  const int first = b << 5;
  const int cnt = max(0, min(32, n_done - first));
  __shared__ unsigned char s_acc[32];
  __shared__ unsigned char s_tab[32 * 224];
  __shared__ int s_start[32], s_cons[32], s_gl2[32];
  const int D = c->D;
  if (cnt > 0 && !c->slow) {
    for (int k = threadIdx.x; k < cnt * D; k += blockDim.x) {
      const int p = k / D, r = k - p * D;
      s_tab[p * 224 + r] = v.ctab[(size_t)(first + p) * 256 + r];
    }
    if (threadIdx.x < cnt) s_gl2[threadIdx.x] = v.g_off[first + threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
      int a = v.blk_start[b];
      for (int p = 0; p < cnt; ++p) {
        const int cc = s_tab[p * 224 + (a - s_gl2[p])];
        s_start[p] = a;
        s_cons[p] = cc;
        a += cc;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int my_start = 0, my_cons = 0;
    if (lane < cnt) {
      if (c->slow) { my_start = 0; my_cons = v.pop_cons[0]; }
      else { my_start = s_start[lane]; my_cons = s_cons[lane]; }
    }
    int acc = 0;
    if (lane < cnt) {
      const int i = first + lane;
      const int g = v.g_off[i];
      const unsigned long long* mk = v.mask + (size_t)i * words;
      const int node = v.queue[c->head + i];
      const float2 np2 = v.node_xy[node];
      const float nz = v.node_z[node];
      const long long dbase = c->pos0 - c->draws_base;
      v.pop_off[i] = my_start;
      v.pop_cons[i] = (unsigned short)my_cons;
      for (int j = my_start - g; j < my_start - g + my_cons; ++j) {
        const bool coll = (mk[j >> 6] >> (j & 63)) & 1ull;
        if (coll) continue;
        const float2 d = __ldg(v.draws + (dbase + g + j));
        const size_t slot = (size_t)i * S + acc;
        v.s_xy[slot] = make_float2(__fadd_rn(np2.x, d.x), __fadd_rn(np2.y, d.y));
        v.s_p1[3 * slot] = np2.x; v.s_p1[3 * slot + 1] = np2.y; v.s_p1[3 * slot + 2] = nz;
        ++acc;
      }
      v.pop_acc[i] = (unsigned char)acc;  // < S only when the trial cap ended the pop
    }
    s_acc[lane] = (unsigned char)acc;
  }
  __syncthreads();  // (block-scope: the sample writes above are visible to the whole CTA)
  const float hinv = 1.0f / (v.e * 1.001f + 1e-5f);
  for (int k = threadIdx.x; k < 32 * S; k += blockDim.x) {
    const int p = k / S, j = k - p * S;
    const int s = (first + p) * S + j;
    if (p >= cnt || j >= s_acc[p]) {
      v.s_d2[s] = 0.f;
      v.s_nn[s] = -1;
      v.s_tie[s] = 0;
      v.s_pc[s] = 0;
      continue;
    }
    exp_nearest_one(v, s, hinv);
  }
}

// Per sample, the earlier samples of the step that may become a node no farther away than the
// sample's nearest existing node — the only ones whose fate can change the sample's own (a node is
// relevant only if strictly nearer, or exactly as near: a tie the host resolves). Usually 0 - 3.
__global__ void __launch_bounds__(256) k_exp_deps(ExpView v) {
  const ExpCtl* c = v.ctl;
  const int S = v.S;
  const int ns = c->n_done * S;
  if (ns == 0) return;
  const float hinv = 1.0f / (v.e * 1.001f + 1e-5f);
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < ns; s += gridDim.x * blockDim.x) {
    const int i = s / S, j = s - i * S;
    if (j >= v.pop_acc[i]) continue;
    const float2 p = v.s_xy[s];
    const float bd = v.s_d2[s];
    const int cx = (int)floorf((p.x - v.gx0) * hinv), cy = (int)floorf((p.y - v.gy0) * hinv);
    int hd[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) hd[k] = v.dhead[dep_hash(cx + (k % 3) - 1, cy + (k / 3) - 1)];
    int n = 0;
#pragma unroll
    for (int k = 0; k < 9; ++k)
      for (int q = hd[k]; q >= 0; q = v.dnext[q]) {
        if (q >= s) continue;
        const float2 pq = v.s_xy[q];
        const float dx = __fsub_rn(pq.x, p.x), dy = __fsub_rn(pq.y, p.y);
        const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        if (!(d2 <= bd)) continue;
        // (two buckets of the 3 x 3 block can collide in the hash: the same entry must not be listed twice)
        bool dup = false;
        for (int u = 0; u < min(n, kDepMax); ++u) dup |= v.dep_j[(size_t)s * kDepMax + u] == q;
        if (dup) continue;
        if (n < kDepMax) { v.dep_j[(size_t)s * kDepMax + n] = q; v.dep_d2[(size_t)s * kDepMax + n] = d2; }
        ++n;
      }
    v.dep_n[s] = n;
  }
}

__global__ void __launch_bounds__(1024) k_exp_tables(ExpView v) {
  extern __shared__ unsigned char chsm[];
  ExpCtl* c = v.ctl;
  // (spare work: empty the hash the emit kernel fills)
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < kDepHash; k += gridDim.x * blockDim.x) v.dhead[k] = -1;
  const int m = c->m;
  if (m == 0) return;
  if (c->slow) {
    if (blockIdx.x == 0) exp_top_phase(v, chsm);
    return;
  }
  const int nblk = (m + 31) >> 5;
  if ((int)blockIdx.x >= nblk) return;
  exp_tables_block(v, chsm, blockIdx.x);
  // the CTA that finishes last walks the chain
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&c->blocks_done, 1) == nblk - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  exp_top_phase(v, chsm);
}

// ------------------------------------------------------------------------------------------
// commit: one CTA
// ------------------------------------------------------------------------------------------
// slope gate of wireEdge (trg.cpp:269-274) with the reference's float operands; the comparison of
// the float results of glibc's atan2f is reproduced only when it cannot depend on the last ulps:
// 0 pass, 1 reject, 2 within 3 float ulps of the threshold (host decides)
__device__ __forceinline__ int slope_gate(float x1, float y1, float z1, float x2, float y2, float z2, float max_slope) {
  const float dz = fabsf(__fsub_rn(z1, z2));
  const float ax = __fsub_rn(x1, x2), ay = __fsub_rn(y1, y2);
  const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)));
  const double s = atan2((double)dz, (double)d);
  const double ms = (double)max_slope;
  const double band = 3.0 * 5.9604644775390625e-08 * fmax(1.0, ms * 2.0);  // 3 ulp of a float in [0.5, 2)
  if (s > ms + band) return 1;
  if (s < ms - band) return 0;
  return 2;
}

constexpr int kCommitCtas = 8;  // thread-block cluster of the commit kernel

// Reservation rounds + application of the decisions in (pop, sample) order + plan of the next step, on a
// cluster of kCommitCtas CTAs: samples are dealt round-robin to the cluster's threads; a round costs an
// undecided sample a walk over its short dependency list (k_exp_deps) and the cluster two barriers.
__global__ void __launch_bounds__(1024, 1) k_exp_commit(ExpView v, int c_step_next) {
  cg::cluster_group cl = cg::this_cluster();
  const int NC = (int)cl.num_blocks(), rank = (int)cl.block_rank();
  ExpCtl* c = v.ctl;
  const int tid = threadIdx.x, T = blockDim.x;
  const int gtid = rank * T + tid, GT = NC * T;
  const int n_done = c->n_done;
  const int m_step = c->m;
  if (n_done == 0) {
    if (rank == 0) {
      if (tid == 0) {
        c->steps++;
        if (m_step > 0) {
          c->stuck++;  // not even the first pop fitted its window: slow mode next
          c->redo_pops += m_step;
        }
      }
      exp_plan_block(v, c_step_next);
    }
    return;
  }
  __shared__ int s_any;
  __shared__ unsigned long long s_part[64];  // warp totals, then exclusive warp bases
  __shared__ unsigned long long s_base;
  const int S = v.S;
  const float r = v.r;
  const float hinv = 1.0f / (v.e * 1.001f + 1e-5f);
  const int head = c->head;
  const int n_nodes0 = c->n_nodes;
  const long long req0 = c->n_req;
  const int tail0 = c->tail;
  const int ns = n_done * S;
  // ---- initial state ------------------------------------------------------------------------------
  for (int s = gtid; s < ns; s += GT) {
    const int i = s / S, j = s - i * S;
    int st = ST_UNDECIDED;
    unsigned char pc = 0;
    if (j >= v.pop_acc[i]) st = ST_VOID;
    else {
      if (v.s_tie[s]) atomicMin(&c->ipop, i);  // tie among existing nodes: the host resolves this pop
      if (v.s_pc[s]) {                         // trg.cpp:414 with the pre-step nearest node: may become a node
        // validity of the would-be node = its parent edge (trg.cpp:425, 447): K4 stage + slope gate
        int valid = 0;
        if (v.s_stage[s] == TRGB_EDGE_OK) {
          const float2 p = v.s_xy[s];
          const int g = slope_gate(v.s_p1[3 * s], v.s_p1[3 * s + 1], v.s_p1[3 * s + 2], p.x, p.y, v.s_z[s], v.max_slope);
          valid = (g == 0) ? 1 : (g == 1 ? 0 : 2);
        } else if (v.s_stage[s] == TRGB_EDGE_SKIPPED) {
          valid = 3;  // cannot happen for a potential creator
        }
        pc = (unsigned char)(1 | (valid << 1));  // bit 0 potential creator, bits 1-2: 0 invalid 1 valid 2/3 host decides
      }
    }
    v.st[s] = st;
    v.s_pc[s] = pc;
    v.cur_d2[s] = v.s_d2[s];
  }
  cl.sync();
  // ---- reservations, asynchronously ---------------------------------------------------------------------
  // No rounds, no barriers: every thread keeps sweeping its undecided samples. A sample is decided as soon
  // as no earlier sample that may still become a node lies within its current nearest distance; what it
  // reads of other samples (their state, their current nearest distance) only ever moves one way
  // (undecided -> decided once, distance downwards), so a stale read can only delay it. The earliest
  // undecided sample never waits, all CTAs of the cluster are resident: the sweep terminates.
  int rounds = 0;
  volatile int* vst = v.st;
  volatile float* vd2 = v.cur_d2;
  // A blocked sample polls only the earlier sample that blocked it (one round trip) until that one is decided or
  // can no longer become a node, and walks its whole list again only then: a chain of k dependent samples costs
  // k polls instead of k list walks. (Kept for the thread's first sample: a step rarely has more samples than
  // the cluster has threads.)
  int blocker0 = -1;
  for (int pending = 1; pending;) {
    pending = 0;
    ++rounds;
    for (int s = gtid; s < ns; s += GT) {
      if (vst[s] != ST_UNDECIDED) continue;
      if (s == gtid && blocker0 >= 0) {
        const int sb = vst[blocker0];
        const float db = vd2[blocker0];
        if (sb == ST_UNDECIDED && __fsqrt_rn(db) >= r) { ++pending; continue; }
        blocker0 = -1;
      }
      float bd = __ldg(v.s_d2 + s);
      int bn = __ldg(v.s_nn + s);
      int tie = 0;
      float mu = INFINITY;  // closest earlier undecided potential creator
      int mu_q = -1;
      const int nd = __ldg(v.dep_n + s);
      if (nd <= kDepMax) {
        // the list in chunks of four: indices, states and distances of a chunk are fetched together (independent
        // loads in flight at once) instead of one dependent round trip after the other
#pragma unroll
        for (int u0 = 0; u0 < kDepMax; u0 += 4) {
          if (u0 >= nd) break;
          int qs[4], sq[4];
          float dd[4], vq[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const bool in = u0 + u < nd;
            qs[u] = in ? __ldg(v.dep_j + (size_t)s * kDepMax + u0 + u) : -1;
            dd[u] = in ? __ldg(v.dep_d2 + (size_t)s * kDepMax + u0 + u) : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) sq[u] = qs[u] >= 0 ? vst[qs[u]] : ST_VOID;
#pragma unroll
          for (int u = 0; u < 4; ++u) vq[u] = sq[u] == ST_UNDECIDED ? vd2[qs[u]] : 0.f;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (sq[u] == ST_CREATE) {
              if (dd[u] < bd) { bd = dd[u]; bn = -2 - qs[u]; tie = 0; }
              else if (dd[u] == bd && bn != -2 - qs[u]) tie = 1;
            } else if (sq[u] == ST_UNDECIDED && __fsqrt_rn(vq[u]) >= r) {
              if (dd[u] < mu) { mu = dd[u]; mu_q = qs[u]; }
            }
          }
        }
      } else {
        // list overflow (a crowd of samples in one spot): walk the hash itself
        const float2 p = v.s_xy[s];
        const int cx = (int)floorf((p.x - v.gx0) * hinv), cy = (int)floorf((p.y - v.gy0) * hinv);
        for (int yy = cy - 1; yy <= cy + 1; ++yy)
          for (int xx = cx - 1; xx <= cx + 1; ++xx) {
            const unsigned h = dep_hash(xx, yy);
            bool seen = false;  // hash collision inside the 3 x 3 block: walk each bucket once
            for (int y2 = cy - 1; y2 <= yy; ++y2)
              for (int x2 = cx - 1; x2 <= cx + 1; ++x2)
                if ((y2 < yy || x2 < xx) && dep_hash(x2, y2) == h) seen = true;
            if (seen) continue;
            for (int q = v.dhead[h]; q >= 0; q = v.dnext[q]) {
              if (q >= s) continue;
              const int sq = vst[q];
              if (sq != ST_CREATE && sq != ST_UNDECIDED) continue;
              const float2 pq = v.s_xy[q];
              const float dx = __fsub_rn(pq.x, p.x), dy = __fsub_rn(pq.y, p.y);
              const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
              if (sq == ST_CREATE) {
                if (d2 < bd) { bd = d2; bn = -2 - q; tie = 0; }
                else if (d2 == bd && bn != -2 - q) tie = 1;
              } else if (__fsqrt_rn(vd2[q]) >= r) {
                mu = fminf(mu, d2);
              }
            }
          }
      }
      vd2[s] = bd;  // published either way: lets later samples see that this one can no longer become a node
      if (mu <= bd) {
        if (s == gtid) blocker0 = mu_q;  // (-1 after the hash walk: no single blocker recorded there)
        ++pending;
        continue;
      }
      // (a CREATE entry at a distance above mu cannot have been missed: entries are only skipped when decided otherwise)
      const int i = s / S;
      if (tie) atomicMin(&c->ipop, i);
      int nstate;  // state of the nearest node (NodeState: -1 invalid)
      if (bn >= 0) nstate = v.node_state[bn];
      else if (bn <= -2) nstate = (((v.s_pc[-2 - bn] >> 1) & 3) == 0) ? -1 : 0;
      else nstate = -1;  // no node at all (cannot happen: the root exists)
      int st;
      if (nstate == -1) st = ST_SKIP;                               // trg.cpp:411
      else if (__fsqrt_rn(bd) < r) st = ST_WIRE;                    // trg.cpp:414
      else {
        st = ST_CREATE;                                             // trg.cpp:421
        if (((v.s_pc[s] >> 1) & 3) >= 2) atomicMin(&c->ipop, i);    // slope gate too close to call
      }
      v.cur_nn[s] = bn;  // nearest node: >= 0 existing, <= -2 sample of this step (read after the barrier below)
      __threadfence();
      vst[s] = st;
    }
  }
  cl.sync();
  // ---- apply the decisions of the pops before the first one the host must handle --------------------
  const int n_commit = min(((volatile int*)&c->ipop)[0], n_done);
  const int nsc = n_commit * S;
  const int per = (nsc + GT - 1) / GT;  // every thread of the cluster applies a contiguous run
  const int a0 = min(nsc, gtid * per), a1 = min(nsc, a0 + per);
  // packed counters: created (bits 0-20), queued (21-41), requests (42-62)
  unsigned long long mine = 0;
  for (int s = a0; s < a1; ++s) {
    const int st = v.st[s];
    if (st == ST_CREATE) {
      const int valid = ((v.s_pc[s] >> 1) & 3) == 1;
      mine += 1ull + (valid ? (1ull << 21) + (1ull << 42) : 0ull);
    } else if (st == ST_WIRE) {
      mine += 1ull << 42;
    }
  }
  {  // exclusive scan of this CTA's 1024 partials: shuffles inside each warp, then over the 32 warp totals
    unsigned long long inc = mine;
    const int lane = tid & 31, wid = tid >> 5;
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long t = __shfl_up_sync(FULL, inc, d);
      if (lane >= d) inc += t;
    }
    if (lane == 31) s_part[wid] = inc;   // warp totals in the first 32 slots
    __syncthreads();
    if (wid == 0) {
      const unsigned long long wt = s_part[lane];
      unsigned long long winc = wt;
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(FULL, winc, d);
        if (lane >= d) winc += t;
      }
      s_part[32 + lane] = winc - wt;      // exclusive warp bases
      if (lane == 31) c->cta_tot[rank] = winc;
    }
    __syncthreads();
    mine = s_part[32 + wid] + inc - mine;  // exclusive prefix of this thread inside the CTA
  }
  cl.sync();
  if (tid == 0) {
    unsigned long long base = 0, tot = 0;
    for (int k = 0; k < NC; ++k) {
      const unsigned long long x = ((volatile unsigned long long*)c->cta_tot)[k];
      if (k < rank) base += x;
      tot += x;
    }
    s_base = base;
    const int n_new = (int)(tot & 0x1fffff), n_q = (int)((tot >> 21) & 0x1fffff);
    const long long n_rq = (long long)((tot >> 42) & 0x1fffff);
    // capacity: never write past the arrays (the packing bound of the caller makes this unreachable)
    s_any = (n_nodes0 + n_new > v.node_cap || req0 + n_rq > v.req_cap || tail0 + n_q > v.node_cap) ? 1 : 0;
  }
  __syncthreads();
  if (s_any) {  // (the same verdict in every CTA)
    if (rank == 0) {
      if (tid == 0) {
        c->steps++;
        c->interrupt = EXP_INT_CAPACITY;
        c->interrupt_pop = head;
        c->n_commit = 0;
      }
      exp_plan_block(v, c_step_next);  // (plans an idle step: the interrupt flag is set)
    }
    return;
  }
  const unsigned long long run0 = s_base + mine;
  {
    unsigned long long run = run0;
    for (int s = a0; s < a1; ++s) {
      const int st = v.st[s];
      const int i = s / S;
      const int parent = v.queue[head + i];
      const int nid = n_nodes0 + (int)(run & 0x1fffff);
      const int qi = tail0 + (int)((run >> 21) & 0x1fffff);
      const long long ri = req0 + (long long)((run >> 42) & 0x1fffff);
      if (st == ST_CREATE) {
        const int valid = ((v.s_pc[s] >> 1) & 3) == 1;
        const float2 p = v.s_xy[s];
        v.node_xy[nid] = p;
        v.node_z[nid] = v.s_z[s];
        v.node_state[nid] = (signed char)(valid ? v.new_state : -1);
        v.hnext[s] = nid;  // id of the node this sample became (read by later samples' requests)
        const int cell = exp_ncell(p.y, v.gy0, v.ginv, v.GH) * v.GW + exp_ncell(p.x, v.gx0, v.ginv, v.GW);
        v.gnext[nid] = atomicExch(v.ghead + cell, nid);
        if (v.s_ztie[s]) atomicAdd((unsigned long long*)&c->z_ties, 1ull);
        if (valid) {
          v.queue[qi] = nid;
          v.req_a[ri] = parent;
          v.req_b[ri] = nid | 0x80000000;  // parent edge: evaluated already
          v.req_w[ri] = v.s_w[s];
          v.req_d[ri] = v.s_d[s];
        }
        run += 1ull + (valid ? (1ull << 21) + (1ull << 42) : 0ull);
      } else if (st == ST_WIRE) {
        v.req_a[ri] = parent;
        v.req_b[ri] = v.cur_nn[s];  // >= 0 existing node, <= -2 sample index (patched below)
        run += 1ull << 42;
      }
    }
  }
  cl.sync();
  // requests that point at a node created in this step: sample index -> node id
  {
    unsigned long long run = run0;
    for (int s = a0; s < a1; ++s) {
      const int st = v.st[s];
      const long long ri = req0 + (long long)((run >> 42) & 0x1fffff);
      if (st == ST_CREATE) {
        const int valid = ((v.s_pc[s] >> 1) & 3) == 1;
        run += 1ull + (valid ? (1ull << 21) + (1ull << 42) : 0ull);
      } else if (st == ST_WIRE) {
        const int b = v.req_b[ri];
        if (b <= -2) v.req_b[ri] = v.hnext[-2 - b];
        run += 1ull << 42;
      }
    }
  }
  // (no cluster barrier here: what CTA 0 does below reads nothing the other CTAs are still writing)
  if (rank != 0) return;
  // ---- CTA 0: control block, statistics, plan of the next step -----------------------------------------
  __shared__ unsigned long long s_sum, s_sq;
  __shared__ ExpCtl s_ctl;
  if (tid == 0) { s_sum = 0; s_sq = 0; }
  __syncthreads();
  {  // draws / pop statistics of the committed pops
    unsigned a = 0, b = 0;  // (a pop consumes < 65536 draws; a thread owns few pops)
    for (int i = tid; i < n_commit; i += T) { const unsigned cc = v.pop_cons[i]; a += cc; b += cc * cc; }
    a = __reduce_add_sync(FULL, a);
    b = __reduce_add_sync(FULL, b);
    if ((tid & 31) == 0 && a) { atomicAdd(&s_sum, (unsigned long long)a); atomicAdd(&s_sq, (unsigned long long)b); }
  }
  __syncthreads();
  if (tid == 0) {
    unsigned long long tot = 0;
    for (int k = 0; k < NC; ++k) tot += ((volatile unsigned long long*)c->cta_tot)[k];
    const int n_new = (int)(tot & 0x1fffff), n_q = (int)((tot >> 21) & 0x1fffff);
    const long long n_rq = (long long)((tot >> 42) & 0x1fffff);
    __threadfence();
    ExpCtl k = *c;  // one read, one write-back: every field access below would otherwise be an L2 round trip
    k.steps++;
    k.n_nodes = n_nodes0 + n_new;
    k.tail = tail0 + n_q;
    k.n_req = req0 + n_rq;
    k.head = head + n_commit;
    k.n_commit = n_commit;
    // stream position after the last committed pop
    long long pos = k.pos0;
    if (n_commit > 0) pos = k.pos0 + v.pop_off[n_commit - 1] + v.pop_cons[n_commit - 1];
    k.pos = pos;
    k.pops += n_commit;
    k.created += n_new;
    k.samples += nsc;
    k.rounds += rounds;
    k.redo_pops += k.m - n_commit;
    k.stuck = 0;
    if (n_commit < n_done) {
      k.interrupt = EXP_INT_TIE;  // tie or slope: the host handles pop `head` with the reference's rules
      k.interrupt_pop = k.head;
    }
    // running draws / pop statistics (step estimate blended in)
    if (n_commit >= 8) {
      const double mu = (double)s_sum / n_commit;
      const double var = fmax(0.0, (double)s_sq / n_commit - mu * mu);
      const float w = n_commit >= 256 ? 0.5f : 0.2f;
      k.mean += w * ((float)mu - k.mean);
      k.var += w * ((float)var - k.var);
      if (k.var < 0.05f) k.var = 0.05f;
    }
    s_ctl = k;  // handed to the plan below in shared memory; written back to the device copy once, there
  }
  // ---- plan the next step --------------------------------------------------------------------------
  exp_plan_block(v, c_step_next, &s_ctl);
}

// a pop handled by the host (interrupt): append its nodes, queue entries and requests
__global__ void k_exp_apply(ExpView v, int n_new, const float* __restrict__ nodes /* x,y,z,state per node */,
                            int n_req, const int* __restrict__ ra, const int* __restrict__ rb,
                            const float* __restrict__ rw, const float* __restrict__ rd, long long new_pos) {
  ExpCtl* c = v.ctl;
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int nid = c->n_nodes;
  int tail = c->tail;
  for (int k = 0; k < n_new; ++k, ++nid) {
    const float x = nodes[4 * k], y = nodes[4 * k + 1];
    v.node_xy[nid] = make_float2(x, y);
    v.node_z[nid] = nodes[4 * k + 2];
    const int stt = (int)nodes[4 * k + 3];
    v.node_state[nid] = (signed char)stt;
    const int cell = exp_ncell(y, v.gy0, v.ginv, v.GH) * v.GW + exp_ncell(x, v.gx0, v.ginv, v.GW);
    v.gnext[nid] = v.ghead[cell];
    v.ghead[cell] = nid;
    if (stt != -1) v.queue[tail++] = nid;
  }
  long long ri = c->n_req;
  for (int k = 0; k < n_req; ++k, ++ri) {
    v.req_a[ri] = ra[k]; v.req_b[ri] = rb[k]; v.req_w[ri] = rw[k]; v.req_d[ri] = rd[k];
  }
  c->n_nodes = nid;
  c->tail = tail;
  c->n_req = ri;
  c->head += 1;
  c->pos = new_pos;
  c->pops += 1;
  c->interrupt = EXP_INT_NONE;
  c->stuck = 0;
}

// ------------------------------------------------------------------------------------------
// finalize: evaluate the wire requests, keep the first success of every unordered pair, group by node
// ------------------------------------------------------------------------------------------
__global__ void k_fin_prepare(ExpView v, long long n_req, float* __restrict__ p1, float2* __restrict__ p2,
                              float* __restrict__ skip) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n_req; k += (long long)gridDim.x * blockDim.x) {
    const int a = v.req_a[k], b = v.req_b[k];
    if (b < 0) {  // parent edge, evaluated during the expansion
      skip[k] = 0.f;
      p1[3 * k] = p1[3 * k + 1] = p1[3 * k + 2] = 0.f;
      p2[k] = make_float2(1.f, 0.f);
      continue;
    }
    const float2 pa = v.node_xy[a], pb = v.node_xy[b];
    p1[3 * k] = pa.x; p1[3 * k + 1] = pa.y; p1[3 * k + 2] = v.node_z[a];
    p2[k] = pb;
    skip[k] = INFINITY;
  }
}

// ok[k] = 1 when request k would create its edge if it were the first of its pair (trg.cpp:255, 269-329);
// 2 = slope gate too close to call (host decides)
__global__ void k_fin_gate(ExpView v, long long n_req, const unsigned char* __restrict__ stage, const float* __restrict__ w,
                           const float* __restrict__ d, unsigned char* __restrict__ ok, int* __restrict__ unc_list,
                           int* __restrict__ unc_count, int unc_cap) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n_req; k += (long long)gridDim.x * blockDim.x) {
    const int a = v.req_a[k], b = v.req_b[k];
    if (b < 0) { ok[k] = 1; continue; }
    if (a == b) { ok[k] = 0; continue; }  // trg.cpp:255
    int res = 0;
    if (stage[k] == TRGB_EDGE_OK) {
      const float2 pa = v.node_xy[a], pb = v.node_xy[b];
      const int g = slope_gate(pa.x, pa.y, v.node_z[a], pb.x, pb.y, v.node_z[b], v.max_slope);
      if (g == 0) res = 1;
      else if (g == 2) {
        res = 2;
        const int slot = atomicAdd(unc_count, 1);
        if (slot < unc_cap) unc_list[slot] = (int)k;
      }
    }
    if (res == 1) { v.req_w[k] = w[k]; v.req_d[k] = d[k]; }
    ok[k] = (unsigned char)res;
  }
}

__global__ void k_fin_keys(ExpView v, long long n_req, const unsigned char* __restrict__ ok,
                           unsigned long long* __restrict__ key, unsigned int* __restrict__ val, int* __restrict__ count) {
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n_req; k += (long long)gridDim.x * blockDim.x) {
    if (ok[k] != 1) continue;
    const int a = v.req_a[k], b = v.req_b[k] & 0x7fffffff;
    const unsigned lo = (unsigned)min(a, b), hi = (unsigned)max(a, b);
    const int slot = atomicAdd(count, 1);
    key[slot] = ((unsigned long long)lo << 32) | hi;
    val[slot] = (unsigned)k;
  }
}

// after the (pair, request index) sort: the first entry of every pair is the edge; emit both directions
__global__ void k_fin_edges(ExpView v, int n, const unsigned long long* __restrict__ key, const unsigned int* __restrict__ val,
                            unsigned long long* __restrict__ dkey, unsigned int* __restrict__ dval, int* __restrict__ count) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (i > 0 && key[i - 1] == key[i]) continue;  // a later request for the same pair: edge exists (trg.cpp:258-267)
    const unsigned k = val[i];
    const int a = v.req_a[k], b = v.req_b[k] & 0x7fffffff;
    const int slot = atomicAdd(count, 2);
    dkey[slot] = ((unsigned long long)(unsigned)a << 32) | k;      // node1->edges_.push_back(edge to node2)
    dval[slot] = (unsigned)b;
    dkey[slot + 1] = ((unsigned long long)(unsigned)b << 32) | k;  // node2->edges_.push_back(edge to node1)
    dval[slot + 1] = (unsigned)a;
  }
}

__global__ void k_fin_csr(ExpView v, int n_dir, int n_nodes, const unsigned long long* __restrict__ dkey,
                          const unsigned int* __restrict__ dval, long long* __restrict__ row_ptr, int* __restrict__ col,
                          float* __restrict__ w, float* __restrict__ d) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= n_dir; i += gridDim.x * blockDim.x) {
    const int src = i < n_dir ? (int)(dkey[i] >> 32) : n_nodes;
    const int prev = i > 0 ? (int)(dkey[i - 1] >> 32) : -1;
    for (int s = prev + 1; s <= src; ++s) row_ptr[s] = i;  // rows without edges get an empty range
    if (i < n_dir) {
      const unsigned k = (unsigned)(dkey[i] & 0xffffffffu);
      col[i] = (int)dval[i];
      w[i] = v.req_w[k];
      d[i] = v.req_d[k];
    }
  }
}

}  // namespace trgb

using namespace trgb;

struct trgb_expander {
  const trgb_map* map = nullptr;
  TrgbExpandParams prm{};
  ExpView v{};
  cudaStream_t st = nullptr;
  int cap = 0;               // z-column capacity of the window kernel
  size_t win_smem = 0;
  int win_grid = 0;
  size_t top_smem = 0;       // dynamic shared memory of k_exp_top
  cudaStream_t st2 = nullptr;  // side stream of the dependency kernel
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  struct StepGraph { int c_step; cudaGraphExec_t exec; };
  std::vector<StepGraph> graphs;
  const float2* graph_draws = nullptr;
  const trgb_map* graph_map = nullptr;
  int graph_kernels = 0;
  bool use_graphs = true;
  // owned device memory
  std::vector<void*> owned;
  float2* d_draws = nullptr;
  long long draws_cap = 0, draws_base = 0, draws_end = 0;
  ExpCtl* h_status[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  // finalize outputs
  long long* d_row = nullptr; int* d_col = nullptr; float* d_w = nullptr; float* d_d = nullptr;
  int n_dir = 0;
  int n_nodes_final = 0;
  // finalize arena: every temporary and output array of trgb_expander_finalize, carved from one grow-only block
  // (a 50 M-point build needs 2 GB of them; taking and returning them from the pool each build fragmented it:
  // 190 - 570 ms stalls in a rebuild loop)
  char* fin_arena = nullptr;
  size_t fin_bytes = 0;
  // pinned host staging of trgb_expander_download_view (grow-only: page-locking 70 MB costs tens of ms)
  void* h_pin = nullptr;
  size_t h_pin_bytes = 0;
};

static void drop_graphs(trgb_expander* e);

namespace {
template <class T>
int dalloc(trgb_expander* e, T** p, size_t n) {
  void* q = nullptr;
  cudaError_t err = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T));
  if (err != cudaSuccess) return cuda_fail(err, "cudaMalloc(expander)", __FILE__, __LINE__);
  e->owned.push_back(q);
  *p = static_cast<T*>(q);
  return TRGB_OK;
}
#define EXP_ALLOC(ptr, n) do { int _rc = dalloc(e, &(ptr), (n)); if (_rc) return _rc; } while (0)
}  // namespace

extern "C" void trgb_expander_destroy(trgb_expander* e) {
  if (!e) return;
  cudaDeviceSynchronize();
  drop_graphs(e);
  if (e->st2) cudaStreamDestroy(e->st2);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_join) cudaEventDestroy(e->ev_join);
  for (void* p : e->owned) cudaFree(p);
  if (e->d_draws) cudaFree(e->d_draws);
  for (int k = 0; k < 2; ++k) {
    if (e->h_status[k]) cudaFreeHost(e->h_status[k]);
    if (e->ev[k]) cudaEventDestroy(e->ev[k]);
  }
  if (e->h_pin) cudaFreeHost(e->h_pin);
  if (e->fin_arena) cudaFree(e->fin_arena);  // (d_row / d_col / d_w / d_d live inside it)
  cudaStreamSynchronize(0);
  delete e;
}

extern "C" int trgb_expander_create(trgb_expander** out, const trgb_map* map, const TrgbExpandParams* prm, float x0,
                                    float y0, float x1, float y1, int64_t node_capacity) {
  TRGB_ARG(out && map && prm, "null pointer");
  TRGB_ARG(prm->sample_num >= 1 && prm->sample_num <= kExpMaxS, "sample_num out of range for the device expander");
  TRGB_ARG(prm->max_pops >= 32 && prm->max_pops <= kExpMaxPops, "max_pops out of range");
  TRGB_ARG(prm->window_words >= 2 && prm->window_words <= 4, "window_words must be 2..4");
  TRGB_ARG(node_capacity >= 1024 && node_capacity < (1ll << 21) * 512, "node_capacity out of range");
  TRGB_ARG(x1 > x0 && y1 > y0, "bad extent");
  // the window kernel is the thread-per-query routine: the map must be sparse enough for its column
  const double area = (double)map->view.W * map->view.H * (double)map->view.cell * map->view.cell;
  const double k = (double)map->n / std::max(area, 1e-9) * 3.14159265358979 * (double)prm->robot_size * prm->robot_size;
  if (map->force_warp_path || k > 0.625 * kTqCap) {
    set_error("device expander: map too dense for the thread-per-query window kernel");
    return TRGB_E_STATE;
  }
  tune_mempool_once();
  trgb_expander* e = new trgb_expander();
  e->map = map;
  e->prm = *prm;
  e->st = map->stream;
  ExpView& v = e->v;
  v.e = prm->expand_dist; v.r = prm->robot_size; v.hthr = prm->height_threshold; v.cthr = prm->collision_threshold;
  v.max_slope = prm->max_slope; v.S = prm->sample_num; v.C = prm->max_pops; v.norm_words = prm->window_words;
  v.new_state = prm->new_state;
  v.node_cap = (int)node_capacity;
  v.req_cap = node_capacity * 10;
  v.gcell = 1.5f * prm->robot_size;
  v.ginv = 1.0f / v.gcell;
  v.gx0 = x0 - 2.f - 2.f * v.gcell;
  v.gy0 = y0 - 2.f - 2.f * v.gcell;
  const double gw = std::floor(((double)x1 + 2.0 - v.gx0) / v.gcell) + 4, gh = std::floor(((double)y1 + 2.0 - v.gy0) / v.gcell) + 4;
  if (gw * gh > 2.0e9) { delete e; set_error("node grid too large"); return TRGB_E_ARG; }
  v.GW = (int)gw; v.GH = (int)gh;
  const size_t C = (size_t)prm->max_pops, S = (size_t)prm->sample_num, CS = C * S;
  EXP_ALLOC(v.ctl, 1);
  EXP_ALLOC(v.node_xy, node_capacity); EXP_ALLOC(v.node_z, node_capacity); EXP_ALLOC(v.node_state, node_capacity);
  EXP_ALLOC(v.queue, node_capacity); EXP_ALLOC(v.gnext, node_capacity);
  EXP_ALLOC(v.ghead, (size_t)v.GW * v.GH);
  EXP_ALLOC(v.req_a, v.req_cap); EXP_ALLOC(v.req_b, v.req_cap); EXP_ALLOC(v.req_w, v.req_cap); EXP_ALLOC(v.req_d, v.req_cap);
  EXP_ALLOC(v.g_off, C);
  EXP_ALLOC(v.mask, std::max<size_t>(C * 4, kExpSlowWords));
  EXP_ALLOC(v.ctab, C * 256);
  const size_t nblk = (C + 31) / 32;
  EXP_ALLOC(v.blk_end, nblk * 256); EXP_ALLOC(v.blk_stop, nblk * 256); EXP_ALLOC(v.blk_start, nblk);
  EXP_ALLOC(v.pop_off, C); EXP_ALLOC(v.pop_cons, C); EXP_ALLOC(v.pop_acc, C);
  EXP_ALLOC(v.s_xy, CS); EXP_ALLOC(v.s_p1, 3 * CS); EXP_ALLOC(v.s_nn, CS); EXP_ALLOC(v.s_d2, CS); EXP_ALLOC(v.s_tie, CS);
  EXP_ALLOC(v.s_z, CS); EXP_ALLOC(v.s_ztie, CS); EXP_ALLOC(v.s_stage, CS); EXP_ALLOC(v.s_w, CS); EXP_ALLOC(v.s_d, CS);
  EXP_ALLOC(v.st, CS); EXP_ALLOC(v.cur_d2, CS); EXP_ALLOC(v.cur_nn, CS); EXP_ALLOC(v.hnext, CS);
  EXP_ALLOC(v.dhead, kDepHash); EXP_ALLOC(v.dnext, CS); EXP_ALLOC(v.dep_j, CS * kDepMax); EXP_ALLOC(v.dep_d2, CS * kDepMax); EXP_ALLOC(v.dep_n, CS); EXP_ALLOC(v.s_pc, CS);
  for (int k2 = 0; k2 < 2; ++k2) {
    TRGB_CUDA(cudaMallocHost((void**)&e->h_status[k2], sizeof(ExpCtl)));
    TRGB_CUDA(cudaEventCreateWithFlags(&e->ev[k2], cudaEventDisableTiming));
  }
  {
    const size_t top_smem = ((size_t)prm->max_pops + 31) / 32 * (64 * (size_t)prm->window_words - 32) * 5;
    if (top_smem > 200 * 1024) { trgb_expander_destroy(e); set_error("expander: max_pops x window too large for the chain kernel"); return TRGB_E_ARG; }
    e->top_smem = std::max<size_t>(top_smem, 1024);
    TRGB_CUDA(cudaFuncSetAttribute((const void*)k_exp_tables, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)std::max<size_t>(e->top_smem, (size_t)32 * kRow + 32 * 256)));

  }
  TRGB_CUDA(cudaStreamCreateWithFlags(&e->st2, cudaStreamNonBlocking));
  TRGB_CUDA(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
  TRGB_CUDA(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
  if (const char* g = std::getenv("TRGB_EXPAND_GRAPHS")) e->use_graphs = std::atoi(g) != 0;
  e->cap = kTqCap;
  e->win_smem = (size_t)kTqThreads * kTqCap * sizeof(float);
  TRGB_CUDA(cudaFuncSetAttribute((const void*)k_exp_window, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->win_smem));
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(12, (200 * 1024) / e->win_smem));
  e->win_grid = sm_count() * per_sm;
  *out = e;
  return TRGB_OK;
}

// a rebuilt map of the same extent: keep every buffer, swap the map (and its stream)
extern "C" int trgb_expander_rebind(trgb_expander* e, const trgb_map* map, float x0, float y0, float x1, float y1,
                                    int64_t node_capacity) {
  TRGB_ARG(e && map, "null pointer");
  const ExpView& v = e->v;
  const float gx0 = x0 - 2.f - 2.f * v.gcell, gy0 = y0 - 2.f - 2.f * v.gcell;
  const double gw = std::floor(((double)x1 + 2.0 - gx0) / v.gcell) + 4, gh = std::floor(((double)y1 + 2.0 - gy0) / v.gcell) + 4;
  const double area = (double)map->view.W * map->view.H * (double)map->view.cell * map->view.cell;
  const double k = (double)map->n / std::max(area, 1e-9) * 3.14159265358979 * (double)v.r * v.r;
  if (gx0 != v.gx0 || gy0 != v.gy0 || (int)gw != v.GW || (int)gh != v.GH || node_capacity > v.node_cap || map->force_warp_path ||
      k > 0.625 * kTqCap) {
    set_error("expander: map extent changed");
    return TRGB_E_STATE;
  }
  // (e->st belonged to the old map and may be gone: builds are synchronous, nothing of ours is in flight)
  drop_graphs(e);  // the captured launches hold the old map's view and stream
  e->map = map;
  e->st = map->stream;
  return TRGB_OK;
}

extern "C" int trgb_expander_begin(trgb_expander* e, float root_x, float root_y, float root_z, int64_t draw_pos) {
  TRGB_ARG(e, "null handle");
  ExpView& v = e->v;
  cudaStream_t st = e->st;
  TRGB_CUDA(cudaMemsetAsync(v.ghead, 0xff, (size_t)v.GW * v.GH * sizeof(int), st));
  ExpCtl c{};
  c.head = 0; c.tail = 1; c.n_nodes = 1; c.pos = draw_pos; c.draws_base = draw_pos; c.draws_end = draw_pos;
  c.mean = 1.15f * (float)v.S; c.var = 2.0f;
  e->draws_base = e->draws_end = draw_pos;
  TRGB_CUDA(cudaMemcpyAsync(v.ctl, &c, sizeof(c), cudaMemcpyHostToDevice, st));
  const float2 rxy = make_float2(root_x, root_y);
  const signed char rs = 0;
  const int zero = 0;
  TRGB_CUDA(cudaMemcpyAsync(v.node_xy, &rxy, sizeof(rxy), cudaMemcpyHostToDevice, st));
  TRGB_CUDA(cudaMemcpyAsync(v.node_z, &root_z, sizeof(float), cudaMemcpyHostToDevice, st));
  TRGB_CUDA(cudaMemcpyAsync(v.node_state, &rs, 1, cudaMemcpyHostToDevice, st));
  TRGB_CUDA(cudaMemcpyAsync(v.queue, &zero, sizeof(int), cudaMemcpyHostToDevice, st));
  // root into the node grid (host computes the cell with the same float arithmetic)
  auto cellof = [](float val, float origin, float inv, int dim) {
    const float f = floorf((val - origin) * inv);
    if (!(f > 0.0f)) return 0;
    if (f >= (float)dim) return dim - 1;
    return (int)f;
  };
  const int cell = cellof(root_y, v.gy0, v.ginv, v.GH) * v.GW + cellof(root_x, v.gx0, v.ginv, v.GW);
  const int minus1 = -1;
  TRGB_CUDA(cudaMemcpyAsync(v.ghead + cell, &zero, sizeof(int), cudaMemcpyHostToDevice, st));
  TRGB_CUDA(cudaMemcpyAsync(v.gnext, &minus1, sizeof(int), cudaMemcpyHostToDevice, st));
  TRGB_CUDA(cudaStreamSynchronize(st));
  return TRGB_OK;
}

// append n draws (x, y offsets) that follow the current end of the device copy of the stream
extern "C" int trgb_expander_push_draws(trgb_expander* e, const float* xy, int64_t n) {
  TRGB_ARG(e && (n == 0 || xy), "null pointer");
  if (n <= 0) return TRGB_OK;
  cudaStream_t st = e->st;
  const long long have = e->draws_end - e->draws_base;
  if (have + n > e->draws_cap) {
    long long want = e->draws_cap ? e->draws_cap : (1ll << 22);
    while (want < have + n) want *= 2;
    float2* nd = nullptr;
    TRGB_CUDA(cudaMalloc((void**)&nd, (size_t)want * sizeof(float2)));
    TRGB_CUDA(cudaStreamSynchronize(st));  // queued steps still read the old buffer
    if (have) TRGB_CUDA(cudaMemcpy(nd, e->d_draws, (size_t)have * sizeof(float2), cudaMemcpyDeviceToDevice));
    if (e->d_draws) cudaFree(e->d_draws);
    e->d_draws = nd;
    e->draws_cap = want;
    e->v.draws = nd;
  }
  TRGB_CUDA(cudaMemcpyAsync(e->d_draws + have, xy, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice, st));
  e->draws_end += n;
  // published to the device control block in stream order (steps queued later see it)
  TRGB_CUDA(cudaMemcpyAsync(&e->v.ctl->draws_end, &e->draws_end, sizeof(long long), cudaMemcpyHostToDevice, st));
  return TRGB_OK;
}

// the kernels of one step, in stream order on e->st (K3 and the dependency lists run on a side stream beside K4)
static int launch_step(trgb_expander* e, int c_step) {
  ExpView& v = e->v;
  cudaStream_t st = e->st;
  const trgb_map* m = e->map;
  const int S = v.S;
  const int64_t ns = (int64_t)c_step * S;
  const int nblk = c_step / 32;
  const int kmax = std::max(2, std::min(16, (int)std::ceil((v.e * 1.001f) / (0.5f * v.r))));
  TrgbEdgeParams ep{v.r, v.hthr, v.cthr, kmax};
  {
    ProfScope ps("k_exp_window", st, (double)c_step * 64.0 * v.norm_words);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(((int64_t)c_step * 64 * v.norm_words + kTqThreads - 1) / kTqThreads, e->win_grid));
    k_exp_window<<<grid, kTqThreads, e->win_smem, st>>>(m->view, v, e->cap);
  }
  {
    ProfScope ps("k_exp_tables", st, (double)c_step);
    const size_t smem = std::max<size_t>((size_t)32 * kRow + 32 * 256, (size_t)nblk * (64 * v.norm_words - 32) * 5);
    k_exp_tables<<<nblk, 1024, smem, st>>>(v);
  }
  {
    ProfScope ps("k_exp_emit", st, (double)ns);
    k_exp_emit<<<nblk, 256, 0, st>>>(v);
  }
  TRGB_CUDA(cudaEventRecord(e->ev_fork, st));
  TRGB_CUDA(cudaStreamWaitEvent(e->st2, e->ev_fork, 0));
  // side stream: heights of the would-be nodes (K3) and the dependency lists; main stream: their parent edges (K4,
  // which reads the parent's height, not the sample's). The commit needs all three.
  int rc = nearest_z_launch_on(m, reinterpret_cast<const float*>(v.s_xy), ns, v.s_z, nullptr, v.s_ztie, v.s_d2, v.r, e->st2);
  if (rc) return rc;
  {
    ProfScope ps("k_exp_deps", e->st2, (double)ns);
    k_exp_deps<<<(int)((ns + 255) / 256), 256, 0, e->st2>>>(v);
  }
  TRGB_CUDA(cudaEventRecord(e->ev_join, e->st2));
  rc = trgb_edge_eval_launch_skip(m, v.s_p1, reinterpret_cast<const float*>(v.s_xy), ns, &ep, v.s_stage, v.s_w, v.s_d, nullptr,
                                  v.s_d2, ns, v.r);
  if (rc) return rc;
  TRGB_CUDA(cudaStreamWaitEvent(st, e->ev_join, 0));
  {
    ProfScope ps("k_exp_commit", st, (double)ns);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kCommitCtas);
    cfg.blockDim = dim3(1024);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCommitCtas;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TRGB_CUDA(cudaLaunchKernelEx(&cfg, k_exp_commit, v, c_step));
  }
  TRGB_CUDA(cudaGetLastError());
  return TRGB_OK;
}

static void drop_graphs(trgb_expander* e) {
  for (auto& g : e->graphs) cudaGraphExecDestroy(g.exec);
  e->graphs.clear();
}

extern "C" int trgb_expander_enqueue(trgb_expander* e, int n_steps, int pops_hint) {
  TRGB_ARG(e && n_steps >= 0, "bad argument");
  ExpView& v = e->v;
  cudaStream_t st = e->st;
  // launch sizes come in powers of two, so that a handful of captured graphs cover a build
  int c_step = 64;
  while (c_step < pops_hint && c_step < v.C) c_step <<= 1;
  c_step = std::min(c_step, v.C) & ~31;
  if (n_steps > 0) {
    ProfScope ps("k_exp_plan", st, 1.0);
    k_exp_plan<<<1, 256, 0, st>>>(v, c_step);
  }
  if (prof_enabled() || !e->use_graphs) {
    for (int k = 0; k < n_steps; ++k) {
      const int rc = launch_step(e, c_step);
      if (rc) return rc;
    }
    return TRGB_OK;
  }
  // one captured CUDA graph per launch size; replayed once per step
  if (e->graph_draws != v.draws || e->graph_map != e->map) {
    drop_graphs(e);
    e->graph_draws = v.draws;
    e->graph_map = e->map;
  }
  cudaGraphExec_t exec = nullptr;
  for (auto& g : e->graphs)
    if (g.c_step == c_step) exec = g.exec;
  if (!exec) {
    cudaGraph_t graph = nullptr;
    const int64_t l0 = trgb_launch_count();
    TRGB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    const int rc = launch_step(e, c_step);
    cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) return cuda_fail(ce, "cudaStreamEndCapture", __FILE__, __LINE__);
    ce = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) return cuda_fail(ce, "cudaGraphInstantiate", __FILE__, __LINE__);
    e->graph_kernels = (int)(trgb_launch_count() - l0);
    count_launches(-e->graph_kernels);  // capturing launched nothing
    e->graphs.push_back({c_step, exec});
  }
  for (int k = 0; k < n_steps; ++k) TRGB_CUDA(cudaGraphLaunch(exec, st));
  count_launches((int64_t)n_steps * e->graph_kernels);
  return TRGB_OK;
}

static void fill_status(const ExpCtl& c, TrgbExpandStatus* s) {
  s->head = c.head; s->tail = c.tail; s->n_nodes = c.n_nodes; s->interrupt = c.interrupt; s->interrupt_pop = c.interrupt_pop;
  s->n_req = c.n_req; s->pos = c.pos; s->draws_end = c.draws_end; s->window_tests = c.window_tests; s->steps = c.steps;
  s->steps_active = c.steps_active; s->rounds = c.rounds; s->pops = c.pops; s->z_ties = c.z_ties; s->redo_pops = c.redo_pops;
  s->samples = c.samples; s->created = c.created; s->mean = c.mean; s->var = c.var;
}

extern "C" int trgb_expander_snapshot(trgb_expander* e, int slot) {
  TRGB_ARG(e && (slot == 0 || slot == 1), "bad argument");
  TRGB_CUDA(cudaMemcpyAsync(e->h_status[slot], e->v.ctl, sizeof(ExpCtl), cudaMemcpyDeviceToHost, e->st));
  TRGB_CUDA(cudaEventRecord(e->ev[slot], e->st));
  return TRGB_OK;
}

extern "C" int trgb_expander_wait(trgb_expander* e, int slot, TrgbExpandStatus* out) {
  TRGB_ARG(e && out && (slot == 0 || slot == 1), "bad argument");
  TRGB_CUDA(cudaEventSynchronize(e->ev[slot]));
  fill_status(*e->h_status[slot], out);
  return TRGB_OK;
}

// nodes [from, to): xyz (3 floats each) and state
extern "C" int trgb_expander_nodes(trgb_expander* e, int64_t from, int64_t to, float* xyz, int8_t* state) {
  TRGB_ARG(e && from >= 0 && to >= from, "bad range");
  const int64_t n = to - from;
  if (n == 0) return TRGB_OK;
  std::vector<float2> xy((size_t)n);
  std::vector<float> z((size_t)n);
  TRGB_CUDA(cudaMemcpyAsync(xy.data(), e->v.node_xy + from, (size_t)n * sizeof(float2), cudaMemcpyDeviceToHost, e->st));
  TRGB_CUDA(cudaMemcpyAsync(z.data(), e->v.node_z + from, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, e->st));
  if (state) TRGB_CUDA(cudaMemcpyAsync(state, e->v.node_state + from, (size_t)n, cudaMemcpyDeviceToHost, e->st));
  TRGB_CUDA(cudaStreamSynchronize(e->st));
  if (xyz)
    for (int64_t i = 0; i < n; ++i) { xyz[3 * i] = xy[i].x; xyz[3 * i + 1] = xy[i].y; xyz[3 * i + 2] = z[i]; }
  return TRGB_OK;
}

// the pop at the queue head, for the host to handle after an interrupt
extern "C" int trgb_expander_head_pop(trgb_expander* e, int32_t* node_id) {
  TRGB_ARG(e && node_id, "null pointer");
  ExpCtl c;
  TRGB_CUDA(cudaMemcpyAsync(&c, e->v.ctl, sizeof(c), cudaMemcpyDeviceToHost, e->st));
  TRGB_CUDA(cudaStreamSynchronize(e->st));
  TRGB_CUDA(cudaMemcpyAsync(node_id, e->v.queue + c.head, sizeof(int), cudaMemcpyDeviceToHost, e->st));
  TRGB_CUDA(cudaStreamSynchronize(e->st));
  return TRGB_OK;
}

extern "C" int trgb_expander_apply_pop(trgb_expander* e, int n_new, const float* nodes_xyzs, int n_req, const int32_t* req_a,
                                       const int32_t* req_b, const float* req_w, const float* req_d, int64_t new_pos) {
  TRGB_ARG(e && n_new >= 0 && n_req >= 0, "bad argument");
  cudaStream_t st = e->st;
  float* dn = nullptr; int* da = nullptr; int* db = nullptr; float* dw = nullptr; float* dd = nullptr;
  TRGB_CUDA(cudaMallocAsync((void**)&dn, std::max(1, n_new) * 4 * sizeof(float), st));
  TRGB_CUDA(cudaMallocAsync((void**)&da, std::max(1, n_req) * sizeof(int), st));
  TRGB_CUDA(cudaMallocAsync((void**)&db, std::max(1, n_req) * sizeof(int), st));
  TRGB_CUDA(cudaMallocAsync((void**)&dw, std::max(1, n_req) * sizeof(float), st));
  TRGB_CUDA(cudaMallocAsync((void**)&dd, std::max(1, n_req) * sizeof(float), st));
  if (n_new) TRGB_CUDA(cudaMemcpyAsync(dn, nodes_xyzs, (size_t)n_new * 4 * sizeof(float), cudaMemcpyHostToDevice, st));
  if (n_req) {
    TRGB_CUDA(cudaMemcpyAsync(da, req_a, (size_t)n_req * sizeof(int), cudaMemcpyHostToDevice, st));
    TRGB_CUDA(cudaMemcpyAsync(db, req_b, (size_t)n_req * sizeof(int), cudaMemcpyHostToDevice, st));
    TRGB_CUDA(cudaMemcpyAsync(dw, req_w, (size_t)n_req * sizeof(float), cudaMemcpyHostToDevice, st));
    TRGB_CUDA(cudaMemcpyAsync(dd, req_d, (size_t)n_req * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  k_exp_apply<<<1, 32, 0, st>>>(e->v, n_new, dn, n_req, da, db, dw, dd, (long long)new_pos);
  TRGB_CUDA(cudaGetLastError());
  cudaFreeAsync(dn, st); cudaFreeAsync(da, st); cudaFreeAsync(db, st); cudaFreeAsync(dw, st); cudaFreeAsync(dd, st);
  TRGB_CUDA(cudaStreamSynchronize(st));
  return TRGB_OK;
}

extern "C" int trgb_expander_finalize(trgb_expander* e, int64_t* n_nodes, int64_t* n_dir_edges) {
  TRGB_ARG(e && n_nodes && n_dir_edges, "null pointer");
  cudaStream_t st = e->st;
  ExpView& v = e->v;
  ExpCtl c;
  TRGB_CUDA(cudaMemcpyAsync(&c, v.ctl, sizeof(c), cudaMemcpyDeviceToHost, st));
  TRGB_CUDA(cudaStreamSynchronize(st));
  const long long nr = c.n_req;
  const int nn = c.n_nodes;
  e->n_nodes_final = nn;
  // one grow-only block for everything below (sizes are bounded by the request count: at most nr requests
  // succeed, each giving two directed entries)
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t NR = (size_t)std::max<long long>(nr, 1);
  const int unc_cap = 1 << 16;
  size_t need = 0;
  auto take = [&](size_t bytes) { const size_t o = need; need += up(bytes); return o; };
  const size_t o_row = take(((size_t)nn + 1) * sizeof(long long)), o_col = take(2 * NR * sizeof(int)), o_ow = take(2 * NR * sizeof(float)),
               o_od = take(2 * NR * sizeof(float)), o_p1 = take(NR * 3 * sizeof(float)), o_p2 = take(NR * sizeof(float2)),
               o_skip = take(NR * sizeof(float)), o_stage = take(NR), o_w = take(NR * sizeof(float)), o_d = take(NR * sizeof(float)),
               o_ok = take(NR), o_cnt = take(4 * sizeof(int)), o_unc = take((size_t)unc_cap * sizeof(int)),
               o_key = take(NR * 8), o_key2 = take(NR * 8), o_val = take(NR * 4), o_val2 = take(NR * 4),
               o_dkey = take(2 * NR * 8), o_dkey2 = take(2 * NR * 8), o_dval = take(2 * NR * 4), o_dval2 = take(2 * NR * 4);
  if (need > e->fin_bytes) {
    if (e->fin_arena) cudaFree(e->fin_arena);
    e->fin_arena = nullptr; e->fin_bytes = 0;
    e->d_row = nullptr; e->d_col = nullptr; e->d_w = nullptr; e->d_d = nullptr;
    const size_t want = need + need / 8;
    TRGB_CUDA(cudaMalloc((void**)&e->fin_arena, want));
    e->fin_bytes = want;
  }
  char* A = e->fin_arena;
  e->d_row = reinterpret_cast<long long*>(A + o_row);
  e->d_col = reinterpret_cast<int*>(A + o_col);
  e->d_w = reinterpret_cast<float*>(A + o_ow);
  e->d_d = reinterpret_cast<float*>(A + o_od);
  if (nr == 0) {
    TRGB_CUDA(cudaMemsetAsync(e->d_row, 0, ((size_t)nn + 1) * sizeof(long long), st));
    TRGB_CUDA(cudaStreamSynchronize(st));
    e->n_dir = 0;
    *n_nodes = nn; *n_dir_edges = 0;
    return TRGB_OK;
  }
  const int grid = sm_count() * 8;
  float* p1 = reinterpret_cast<float*>(A + o_p1); float2* p2 = reinterpret_cast<float2*>(A + o_p2);
  float* skip = reinterpret_cast<float*>(A + o_skip); unsigned char* stage = reinterpret_cast<unsigned char*>(A + o_stage);
  float* w = reinterpret_cast<float*>(A + o_w); float* d = reinterpret_cast<float*>(A + o_d);
  unsigned char* ok = reinterpret_cast<unsigned char*>(A + o_ok); int* cnt = reinterpret_cast<int*>(A + o_cnt);
  int* unc = reinterpret_cast<int*>(A + o_unc);
  TRGB_CUDA(cudaMemsetAsync(cnt, 0, 4 * sizeof(int), st));
  k_fin_prepare<<<grid, 256, 0, st>>>(v, nr, p1, p2, skip);
  // wire requests reach expand_dist + robot_size (a node created robot_size from the sample's parent circle)
  const int kmax = std::max(2, std::min(16, (int)std::ceil((v.e + v.r) / (0.5f * v.r))));
  TrgbEdgeParams ep{v.r, v.hthr, v.cthr, kmax};
  int rc = trgb_edge_eval_launch_skip(e->map, p1, reinterpret_cast<const float*>(p2), nr, &ep, stage, w, d, nullptr, skip, nr, 1.0f);
  if (rc) return rc;
  {
    ProfScope ps("k_fin_gate", st, (double)nr);
    k_fin_gate<<<grid, 256, 0, st>>>(v, nr, stage, w, d, ok, unc, cnt + 1, unc_cap);
  }
  int h_unc = 0;
  TRGB_CUDA(cudaMemcpyAsync(&h_unc, cnt + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  TRGB_CUDA(cudaStreamSynchronize(st));
  if (h_unc > unc_cap) { set_error("finalize: too many uncertain slope gates"); return TRGB_E_STATE; }
  if (h_unc > 0) {
    // glibc's atan2f decides, as in the reference (trg.cpp:269-274)
    std::vector<int> list((size_t)h_unc);
    TRGB_CUDA(cudaMemcpy(list.data(), unc, (size_t)h_unc * sizeof(int), cudaMemcpyDeviceToHost));
    const float max_slope = v.max_slope;
    for (int k : list) {
      int ab[2];
      TRGB_CUDA(cudaMemcpy(&ab[0], v.req_a + k, sizeof(int), cudaMemcpyDeviceToHost));
      TRGB_CUDA(cudaMemcpy(&ab[1], v.req_b + k, sizeof(int), cudaMemcpyDeviceToHost));
      float2 pa, pb; float za, zb;
      TRGB_CUDA(cudaMemcpy(&pa, v.node_xy + ab[0], sizeof(float2), cudaMemcpyDeviceToHost));
      TRGB_CUDA(cudaMemcpy(&pb, v.node_xy + ab[1], sizeof(float2), cudaMemcpyDeviceToHost));
      TRGB_CUDA(cudaMemcpy(&za, v.node_z + ab[0], sizeof(float), cudaMemcpyDeviceToHost));
      TRGB_CUDA(cudaMemcpy(&zb, v.node_z + ab[1], sizeof(float), cudaMemcpyDeviceToHost));
      const float dx = pa.x - pb.x, dy = pa.y - pb.y;
      const float slope = atan2f(fabsf(za - zb), sqrtf(dx * dx + dy * dy));
      const unsigned char res = slope > max_slope ? 0 : 1;
      TRGB_CUDA(cudaMemcpy(ok + k, &res, 1, cudaMemcpyHostToDevice));
      if (res) {
        TRGB_CUDA(cudaMemcpy(v.req_w + k, w + k, sizeof(float), cudaMemcpyDeviceToDevice));
        TRGB_CUDA(cudaMemcpy(v.req_d + k, d + k, sizeof(float), cudaMemcpyDeviceToDevice));
      }
    }
  }
  // (pair, request index) of every request that would succeed, in request order per pair
  unsigned long long *key = reinterpret_cast<unsigned long long*>(A + o_key), *key2 = reinterpret_cast<unsigned long long*>(A + o_key2);
  unsigned int *val = reinterpret_cast<unsigned int*>(A + o_val), *val2 = reinterpret_cast<unsigned int*>(A + o_val2);
  // compaction by atomics loses the request order inside a pair: sort by (pair, index) instead, in
  // two stable passes (index first — it is the value —, then pair)
  k_fin_keys<<<grid, 256, 0, st>>>(v, nr, ok, key, val, cnt);
  int n_ok = 0;
  TRGB_CUDA(cudaMemcpyAsync(&n_ok, cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
  TRGB_CUDA(cudaStreamSynchronize(st));
  {
    ProfScope ps("k_fin_sort", st, (double)n_ok);
    // pass 1: by request index (keys = val); pass 2: stable by pair
    int rc2 = sort_pairs_u32_u64(val, val2, key, key2, n_ok, 32, st);
    if (rc2) return rc2;
    rc2 = sort_pairs_u64_u32(key2, key, val2, val, n_ok, 64, st);
    if (rc2) return rc2;
  }
  // first of each pair -> two directed entries keyed (source node, request index)
  unsigned long long *dkey = reinterpret_cast<unsigned long long*>(A + o_dkey), *dkey2 = reinterpret_cast<unsigned long long*>(A + o_dkey2);
  unsigned int *dval = reinterpret_cast<unsigned int*>(A + o_dval), *dval2 = reinterpret_cast<unsigned int*>(A + o_dval2);
  k_fin_edges<<<grid, 256, 0, st>>>(v, n_ok, key, val, dkey, dval, cnt + 2);
  int n_dir = 0;
  TRGB_CUDA(cudaMemcpyAsync(&n_dir, cnt + 2, sizeof(int), cudaMemcpyDeviceToHost, st));
  TRGB_CUDA(cudaStreamSynchronize(st));
  if (n_dir > 0) {
    ProfScope ps("k_fin_sort", st, (double)n_dir);
    const int rc3 = sort_pairs_u64_u32(dkey, dkey2, dval, dval2, n_dir, 64, st);
    if (rc3) return rc3;
  }
  k_fin_csr<<<grid, 256, 0, st>>>(v, n_dir, nn, dkey2, dval2, e->d_row, e->d_col, e->d_w, e->d_d);
  TRGB_CUDA(cudaGetLastError());
  TRGB_CUDA(cudaStreamSynchronize(st));
  e->n_dir = n_dir;
  *n_nodes = nn;
  *n_dir_edges = n_dir;
  return TRGB_OK;
}

// The finalized graph in page-locked host memory owned by the engine (valid until the next finalize / destroy):
// one DMA per array at full PCIe rate into buffers that are locked once, instead of pageable copies into vectors
// the caller has just allocated (and faults in page by page).
extern "C" int trgb_expander_download_view(trgb_expander* e, TrgbExpandedGraph* out) {
  TRGB_ARG(e && e->d_row && out, "finalize first");
  const size_t nn = (size_t)e->n_nodes_final, nd = (size_t)e->n_dir;
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  const size_t o_xy = 0, o_z = o_xy + up(nn * sizeof(float2)), o_state = o_z + up(nn * sizeof(float)),
               o_row = o_state + up(nn), o_col = o_row + up((nn + 1) * sizeof(long long)), o_w = o_col + up(nd * sizeof(int)),
               o_d = o_w + up(nd * sizeof(float)), total = o_d + up(nd * sizeof(float));
  if (total > e->h_pin_bytes) {
    if (e->h_pin) cudaFreeHost(e->h_pin);
    e->h_pin = nullptr; e->h_pin_bytes = 0;
    const size_t want = total + total / 8;
    TRGB_CUDA(cudaHostAlloc(&e->h_pin, want, cudaHostAllocDefault));
    e->h_pin_bytes = want;
  }
  char* h = static_cast<char*>(e->h_pin);
  cudaStream_t st = e->st;
  TRGB_CUDA(cudaMemcpyAsync(h + o_xy, e->v.node_xy, nn * sizeof(float2), cudaMemcpyDeviceToHost, st));
  TRGB_CUDA(cudaMemcpyAsync(h + o_z, e->v.node_z, nn * sizeof(float), cudaMemcpyDeviceToHost, st));
  TRGB_CUDA(cudaMemcpyAsync(h + o_state, e->v.node_state, nn, cudaMemcpyDeviceToHost, st));
  TRGB_CUDA(cudaMemcpyAsync(h + o_row, e->d_row, (nn + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
  if (nd) {
    TRGB_CUDA(cudaMemcpyAsync(h + o_col, e->d_col, nd * sizeof(int), cudaMemcpyDeviceToHost, st));
    TRGB_CUDA(cudaMemcpyAsync(h + o_w, e->d_w, nd * sizeof(float), cudaMemcpyDeviceToHost, st));
    TRGB_CUDA(cudaMemcpyAsync(h + o_d, e->d_d, nd * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  TRGB_CUDA(cudaStreamSynchronize(st));
  out->n_nodes = (int64_t)nn;
  out->n_directed_edges = (int64_t)nd;
  out->xy = reinterpret_cast<const float*>(h + o_xy);
  out->z = reinterpret_cast<const float*>(h + o_z);
  out->state = reinterpret_cast<const int8_t*>(h + o_state);
  out->row_ptr = reinterpret_cast<const int64_t*>(h + o_row);
  out->col = reinterpret_cast<const int32_t*>(h + o_col);
  out->weight = reinterpret_cast<const float*>(h + o_w);
  out->dist = reinterpret_cast<const float*>(h + o_d);
  return TRGB_OK;
}

// K7 search graph straight from the finalized arrays (no trip through the host): old2new[i] = id of node i in
// the cleaned graph or -1 (dropped by cleanGraph), n_new = number of kept nodes = id range of the new graph.
extern "C" int trgb_expander_make_graph(trgb_expander* e, const int32_t* old2new, int32_t n_new, trgb_graph** out) {
  TRGB_ARG(e && e->d_row && old2new && out, "finalize first");
  const int nn = e->n_nodes_final;
  TRGB_ARG(n_new > 0 && n_new <= nn, "bad node count");
  int kept = 0;
  for (int i = 0; i < nn; ++i) {
    TRGB_ARG(old2new[i] >= -1 && old2new[i] < n_new, "old2new out of range");
    kept += old2new[i] >= 0;
  }
  TRGB_ARG(kept == n_new, "old2new does not keep n_new nodes");
  cudaStream_t st = e->st;
  int32_t* d_map = nullptr;
  TRGB_CUDA(cudaMallocAsync((void**)&d_map, (size_t)nn * sizeof(int32_t), st));
  TRGB_CUDA(cudaMemcpyAsync(d_map, old2new, (size_t)nn * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  GraphSource s;
  s.n_src = nn; s.n_keep = n_new; s.n_ext = n_new; s.e_src = e->n_dir;
  s.row = e->d_row; s.col = e->d_col; s.w = e->d_w; s.dist = e->d_d;
  s.xy = e->v.node_xy; s.state8 = e->v.node_state; s.src2ext = d_map;
  const int rc = graph_from_device(out, s, st);
  cudaFreeAsync(d_map, st);
  return rc;
}

extern "C" int trgb_expander_download(trgb_expander* e, float* xyz, int8_t* state, int64_t* row_ptr, int32_t* col, float* weight,
                                      float* dist) {
  TRGB_ARG(e && e->d_row, "finalize first");
  const int nn = e->n_nodes_final;
  int rc = trgb_expander_nodes(e, 0, nn, xyz, state);
  if (rc) return rc;
  cudaStream_t st = e->st;
  if (row_ptr) TRGB_CUDA(cudaMemcpyAsync(row_ptr, e->d_row, ((size_t)nn + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
  if (e->n_dir > 0) {
    if (col) TRGB_CUDA(cudaMemcpyAsync(col, e->d_col, (size_t)e->n_dir * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (weight) TRGB_CUDA(cudaMemcpyAsync(weight, e->d_w, (size_t)e->n_dir * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (dist) TRGB_CUDA(cudaMemcpyAsync(dist, e->d_d, (size_t)e->n_dir * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  TRGB_CUDA(cudaStreamSynchronize(st));
  return TRGB_OK;
}
