// Host-side index over graph nodes. Replaces the reference's `kdtree* node_tree`
// (trg.h:106) with two cooperating structures:
//
//   NodeGrid     uniform hash grid -> exact nearest / in-range sets in O(1) expected time,
//                using the same float arithmetic as kdtree.c (dist^2 = fl(fl(dx*dx)+fl(dy*dy)),
//                strict `<` for nearest, inclusive `<=` for range);
//   OrderTree2D  an insertion-order 2-D tree kept ONLY for the order-dependent semantics the
//                reference leaks through its kd-tree: which of several exactly
//                equidistant nodes kd_nearest2 returns, and the order in which
//                kd_nearest_range2 results are iterated (head of list = last node visited;
//                kdtree.c:270-301, 759-777). It is built lazily from the insertion sequence.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <future>
#include <thread>
#include <limits>
#include <mutex>
#include <vector>

namespace trg_b200 {

// Helper threads a burst of host work may use: the cores of the box shared between the ranks
// of a multi-GPU run (torchrun exports LOCAL_WORLD_SIZE), capped at 8.
inline int thread_budget() {
  static const int budget = [] {
    int hw = static_cast<int>(std::thread::hardware_concurrency());
    if (hw <= 0) hw = 4;
    int ranks = 1;
    if (const char* e = std::getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, std::atoi(e));
    return std::max(1, std::min(8, (hw - 2 * ranks) / ranks));  // two pipeline threads per rank are always busy
  }();
  return budget;
}

class OrderTree2D {
 public:
  void clear() {
    x_.clear(); y_.clear(); lo_.clear(); hi_.clear(); axis_.clear(); payload_.clear(); parent_.clear();
    have_box_ = false;
  }
  size_t size() const { return x_.size(); }
  void reserve(size_t n) {
    x_.reserve(n); y_.reserve(n); lo_.reserve(n); hi_.reserve(n); axis_.reserve(n); payload_.reserve(n);
  }

  // descend from the root: smaller coordinate on the split axis goes to the low child, ties and
  // larger go to the high child; the split axis alternates x, y with depth (kdtree.c:167-194)
  void insert(float x, float y, int payload) {
    const int self = static_cast<int>(x_.size());
    uint8_t axis = 0;
    parent_of_new_ = -1;
    if (self > 0) {
      int at = 0;
      while (true) {
        const bool low = axis_[at] ? (y < y_[at]) : (x < x_[at]);
        int& slot = low ? lo_[at] : hi_[at];
        if (slot < 0) {
          slot = self;
          axis = axis_[at] ^ 1;
          parent_of_new_ = at;
          break;
        }
        at = slot;
      }
    }
    x_.push_back(x); y_.push_back(y); lo_.push_back(-1); hi_.push_back(-1);
    axis_.push_back(axis); payload_.push_back(payload);
    parent_.push_back(parent_of_new_);
    if (!have_box_) {
      bmin_[0] = bmax_[0] = x; bmin_[1] = bmax_[1] = y; have_box_ = true;
    } else {
      bmin_[0] = std::fmin(bmin_[0], x); bmax_[0] = std::fmax(bmax_[0], x);
      bmin_[1] = std::fmin(bmin_[1], y); bmax_[1] = std::fmax(bmax_[1], y);
    }
  }

  // Builds exactly the tree that inserting (x[i], y[i], payload i) for i = 0..n-1 in order would
  // build, by recursive stable partition: the root of a subtree is the FIRST point (in insertion
  // order) that fell into it, the rest split into "< split" (low) and ">= split" (high) keeping
  // their order. Sequential memory passes instead of one pointer chase per level and point.
  void build_bulk(const float* x, const float* y, int n) {
    clear();
    if (n <= 0) return;
    x_.assign(x, x + n); y_.assign(y, y + n);
    lo_.assign(n, -1); hi_.assign(n, -1); axis_.assign(n, 0); payload_.resize(n);
    bmin_[0] = bmax_[0] = x[0]; bmin_[1] = bmax_[1] = y[0]; have_box_ = true;
    for (int i = 0; i < n; ++i) {
      payload_[i] = i;
      bmin_[0] = std::fmin(bmin_[0], x[i]); bmax_[0] = std::fmax(bmax_[0], x[i]);
      bmin_[1] = std::fmin(bmin_[1], y[i]); bmax_[1] = std::fmax(bmax_[1], y[i]);
    }
    std::vector<int> idx(n), tmp(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    build_range(x, y, idx.data(), tmp.data(), 0, n, 0, 0);
    parent_.assign(n, -1);
    for (int i = 0; i < n; ++i) {
      if (lo_[i] >= 0) parent_[lo_[i]] = i;
      if (hi_[i] >= 0) parent_[hi_[i]] = i;
    }
  }

  // Adopt a tree built elsewhere (trgb_kdtree_build: the same insertion-order tree, grown on the device).
  // xy: n interleaved (x, y) pairs in insertion order.
  void adopt(const float* xy, int n, std::vector<int>&& lo, std::vector<int>&& hi, std::vector<int>&& parent,
             std::vector<uint8_t>&& axis) {
    clear();
    if (n <= 0) return;
    lo_ = std::move(lo); hi_ = std::move(hi); parent_ = std::move(parent); axis_ = std::move(axis);
    x_.resize(n); y_.resize(n); payload_.resize(n);
    const int parts = std::max(1, std::min(thread_budget(), n / 65536));
    std::vector<float> box(4 * (size_t)parts);
    auto fill = [&](int part) {
      const int b = (int)((int64_t)n * part / parts), e = (int)((int64_t)n * (part + 1) / parts);
      float x0 = xy[2 * b], x1 = x0, y0 = xy[2 * b + 1], y1 = y0;
      for (int i = b; i < e; ++i) {
        const float x = xy[2 * i], y = xy[2 * i + 1];
        x_[i] = x; y_[i] = y; payload_[i] = i;
        x0 = std::fmin(x0, x); x1 = std::fmax(x1, x); y0 = std::fmin(y0, y); y1 = std::fmax(y1, y);
      }
      box[4 * part] = x0; box[4 * part + 1] = x1; box[4 * part + 2] = y0; box[4 * part + 3] = y1;
    };
    std::vector<std::future<void>> helpers;
    for (int k = 1; k < parts; ++k) helpers.push_back(std::async(std::launch::async, fill, k));
    fill(0);
    for (auto& h : helpers) h.get();
    bmin_[0] = box[0]; bmax_[0] = box[1]; bmin_[1] = box[2]; bmax_[1] = box[3];
    for (int k = 1; k < parts; ++k) {
      bmin_[0] = std::fmin(bmin_[0], box[4 * k]); bmax_[0] = std::fmax(bmax_[0], box[4 * k + 1]);
      bmin_[1] = std::fmin(bmin_[1], box[4 * k + 2]); bmax_[1] = std::fmax(bmax_[1], box[4 * k + 3]);
    }
    have_box_ = true;
  }

  // Order the given tree nodes (indices = insertion order) the way kd_nearest_range2's result
  // iterator yields them for a query at (qx, qy) whose range contains them all: the traversal
  // (kdtree.c:270-301) is pre-order, query-side child first, and results are PREPENDED, so the
  // iteration runs from the last visited to the first. Only the root paths of the candidates are
  // walked (depth ~ 2.5 log2 n each) instead of traversing the tree around the query.
  void order_like_range(std::vector<int>& cand, float qx, float qy) const {
    if (cand.size() < 2) return;
    auto path_of = [&](int v, std::vector<int>& out) {
      out.clear();
      for (int a = v; a >= 0; a = parent_[a]) out.push_back(a);  // v ... root
    };
    std::vector<std::vector<int>> paths(cand.size());
    for (size_t i = 0; i < cand.size(); ++i) path_of(cand[i], paths[i]);
    // visited_before(a, b): a is reached before b by the traversal
    auto visited_before = [&](size_t ia, size_t ib) {
      const auto& pa = paths[ia];
      const auto& pb = paths[ib];
      size_t ka = pa.size(), kb = pb.size();   // walk down from the root while the paths agree
      while (ka > 0 && kb > 0 && pa[ka - 1] == pb[kb - 1]) { --ka; --kb; }
      if (ka == 0) return true;    // a is an ancestor of b (or equal): pre-order visits it first
      if (kb == 0) return false;   // b is an ancestor of a
      const int l = pa[ka];        // lowest common ancestor (last agreeing node)
      const float dx = axis_[l] ? (qy - y_[l]) : (qx - x_[l]);
      const int near = dx <= 0.0f ? lo_[l] : hi_[l];
      return pa[ka - 1] == near;   // a sits in the subtree the traversal enters first
    };
    std::vector<size_t> ord(cand.size());
    for (size_t i = 0; i < ord.size(); ++i) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return a != b && visited_before(b, a); });  // reverse visit order
    std::vector<int> out(cand.size());
    for (size_t i = 0; i < ord.size(); ++i) out[i] = cand[ord[i]];
    cand.swap(out);
  }

 private:
  // subtree over idx[b, e) (insertion order preserved); large halves run on helper threads —
  // the two halves touch disjoint slices of idx / tmp and disjoint tree nodes
  void build_range(const float* x, const float* y, int* idx, int* tmp, int b, int e, uint8_t axis, int depth) {
    struct Job { int b, e; uint8_t axis; int depth; };
    std::vector<Job> jobs;
    std::vector<std::future<void>> helpers;
    jobs.push_back({b, e, axis, depth});
    while (!jobs.empty()) {
      const Job j = jobs.back();
      jobs.pop_back();
      const int root = idx[j.b];
      axis_[root] = j.axis;
      const float split = j.axis ? y[root] : x[root];
      const float* key = j.axis ? y : x;
      int nl = 0, nh = 0;
      for (int k = j.b + 1; k < j.e; ++k) {
        const int p = idx[k];
        if (key[p] < split) idx[j.b + 1 + nl++] = p;
        else tmp[j.b + nh++] = p;
      }
      for (int k = 0; k < nh; ++k) idx[j.b + 1 + nl + k] = tmp[j.b + k];
      const uint8_t nax = static_cast<uint8_t>(j.axis ^ 1);
      if (nl) lo_[root] = idx[j.b + 1];
      if (nh) hi_[root] = idx[j.b + 1 + nl];
      const int lb = j.b + 1, le = j.b + 1 + nl, hb = le, he = j.e, nd = j.depth + 1;
      if (nh) {
        if ((1 << j.depth) < thread_budget() && nh > 8192 && nl > 8192)
          helpers.push_back(std::async(std::launch::async, [=] { build_range(x, y, idx, tmp, hb, he, nax, nd); }));
        else
          jobs.push_back({hb, he, nax, nd});
      }
      if (nl) jobs.push_back({lb, le, nax, nd});
    }
    for (auto& h : helpers) h.get();
  }

 public:
  // In-range payloads in the order the reference's result iterator yields them: the traversal
  // is pre-order, query side first, far side only if |delta| < r; results are prepended, so
  // the iteration order is the reverse of the visit order.
  void range(float qx, float qy, float r, std::vector<int>& out) const {
    out.clear();
    if (x_.empty()) return;
    const float r2 = r * r;
    walk_.clear();
    walk_.push_back({0, 0.f, false});
    while (!walk_.empty()) {
      Step s = walk_.back();
      walk_.pop_back();
      if (s.second_half) {
        if (std::fabs(s.delta) < r) {
          const int far = s.delta <= 0.0f ? hi_[s.node] : lo_[s.node];
          if (far >= 0) walk_.push_back({far, 0.f, false});
        }
        continue;
      }
      const int n = s.node;
      const float dx = x_[n] - qx, dy = y_[n] - qy;
      float d2 = 0.f;
      d2 += dx * dx;
      d2 += dy * dy;
      if (d2 <= r2) out.push_back(payload_[n]);
      const float delta = axis_[n] ? (qy - y_[n]) : (qx - x_[n]);
      walk_.push_back({n, delta, true});
      const int near = delta <= 0.0f ? lo_[n] : hi_[n];
      if (near >= 0) walk_.push_back({near, 0.f, false});
    }
    for (size_t a = 0, b = out.size(); a + 1 < b; ++a, --b) std::swap(out[a], out[b - 1]);
  }

  // kd_nearest semantics (kdtree.c:303-417): best starts at the root; visit nearer subtree,
  // then the node (strict <), then the farther subtree if its bounding box can still win.
  int nearest(float qx, float qy) const {
    if (x_.empty()) return -1;
    float mn[2] = {bmin_[0], bmin_[1]}, mx[2] = {bmax_[0], bmax_[1]};
    const float q[2] = {qx, qy};
    int best = 0;
    float best_d2 = 0.f;
    best_d2 += (x_[0] - qx) * (x_[0] - qx);
    best_d2 += (y_[0] - qy) * (y_[0] - qy);
    frames_.clear();
    frames_.push_back({0, 0, 0.f, false});
    while (!frames_.empty()) {
      Frame& f = frames_.back();
      const int n = f.node;
      const int ax = axis_[n];
      const float split = ax ? y_[n] : x_[n];
      if (f.stage == 0) {
        f.low_side = (q[ax] - split) <= 0.f;
        f.stage = 1;
        const int near = f.low_side ? lo_[n] : hi_[n];
        if (near >= 0) {
          float* edge = f.low_side ? &mx[ax] : &mn[ax];
          f.saved = *edge;
          *edge = split;
          frames_.push_back({near, 0, 0.f, false});
        }
        continue;
      }
      if (f.stage == 1) {
        const int near = f.low_side ? lo_[n] : hi_[n];
        if (near >= 0) *(f.low_side ? &mx[ax] : &mn[ax]) = f.saved;
        float d2 = 0.f;
        d2 += (x_[n] - qx) * (x_[n] - qx);
        d2 += (y_[n] - qy) * (y_[n] - qy);
        if (d2 < best_d2) {
          best = n;
          best_d2 = d2;
        }
        f.stage = 2;
        const int far = f.low_side ? hi_[n] : lo_[n];
        if (far >= 0) {
          float* edge = f.low_side ? &mn[ax] : &mx[ax];
          f.saved = *edge;
          *edge = split;
          float bd = 0.f;
          for (int k = 0; k < 2; ++k) {
            if (q[k] < mn[k]) bd += (mn[k] - q[k]) * (mn[k] - q[k]);
            else if (q[k] > mx[k]) bd += (mx[k] - q[k]) * (mx[k] - q[k]);
          }
          if (bd < best_d2) {
            frames_.push_back({far, 0, 0.f, false});
            continue;
          }
          *edge = f.saved;  // pruned: undo the slice right away
          f.stage = 3;
        } else {
          f.stage = 3;
        }
        continue;
      }
      if (f.stage == 2) {  // back from the far subtree
        *(f.low_side ? &mn[ax] : &mx[ax]) = f.saved;
      }
      frames_.pop_back();
    }
    return payload_[best];
  }

  // raw arrays (device upload for batched goal/start snapping)
  const std::vector<int>& low() const { return lo_; }
  const std::vector<int>& high() const { return hi_; }
  const std::vector<uint8_t>& axis() const { return axis_; }
  const std::vector<int>& payload() const { return payload_; }
  const std::vector<int>& parent() const { return parent_; }

 private:
  struct Step { int node; float delta; bool second_half; };
  struct Frame { int node; int stage; float saved; bool low_side; };
  std::vector<float> x_, y_;
  std::vector<int> lo_, hi_;
  std::vector<uint8_t> axis_;
  std::vector<int> payload_;
  std::vector<int> parent_;
  int parent_of_new_ = -1;
  float bmin_[2] = {0, 0}, bmax_[2] = {0, 0};
  bool have_box_ = false;
  mutable std::vector<Step> walk_;
  mutable std::vector<Frame> frames_;
};

// Which of several points at the IDENTICAL float distance from (qx, qy) kd_nearest returns, without
// materialising the insertion-order tree. kd_nearest_i (kdtree.c:303-362) visits, at every tree
// node, the subtree on the query's side, then the node, then the other subtree, and replaces its
// result only on a strictly smaller distance — so the first tied point visited wins, and the root
// (the initial result) always wins. The root path of a point is recovered by one pass over the
// earlier insertions: the first inserted point that falls into a subtree's region is its root.
// x, y: coordinates in insertion order (at least up to the largest candidate); cand: ascending
// insertion indices of the tied points. Returns the winning index.
inline int first_visited_of(const float* x, const float* y, const std::vector<int>& cand, float qx, float qy) {
  if (cand.size() == 1 || cand[0] == 0) return cand[0];
  struct Step { int node; bool low; };  // ancestor and the side taken below it
  auto root_path = [&](int target) {
    std::vector<Step> path;
    float lo[2] = {-std::numeric_limits<float>::infinity(), -std::numeric_limits<float>::infinity()};
    float hi[2] = {std::numeric_limits<float>::infinity(), std::numeric_limits<float>::infinity()};
    const float t[2] = {x[target], y[target]};
    int axis = 0;
    for (int i = 0; i < target; ++i) {
      const float p[2] = {x[i], y[i]};
      // region of the current subtree: low side is `< split`, high side is `>= split` (kdtree.c:190-193)
      if (!(p[0] >= lo[0] && p[0] < hi[0] && p[1] >= lo[1] && p[1] < hi[1])) continue;
      const bool low = t[axis] < p[axis];
      path.push_back({i, low});
      if (low) hi[axis] = p[axis]; else lo[axis] = p[axis];
      axis ^= 1;
    }
    return path;
  };
  std::vector<std::vector<Step>> paths;
  for (int c : cand) paths.push_back(root_path(c));
  const float q[2] = {qx, qy};
  auto near_low = [&](int node, int depth) { const int ax = depth & 1; return (q[ax] - (ax ? y[node] : x[node])) <= 0.f; };
  // true when candidate a is visited before candidate b
  auto before = [&](size_t a, size_t b) {
    const auto& pa = paths[a];
    const auto& pb = paths[b];
    size_t j = 0;
    while (j < pa.size() && j < pb.size() && pa[j].node == pb[j].node && pa[j].low == pb[j].low) ++j;
    if (j == pa.size())  // a is an ancestor of b: b comes first iff it sits in a's near subtree
      return !(pb[j].low == near_low(cand[a], static_cast<int>(j)));
    if (j == pb.size())  // b is an ancestor of a
      return pa[j].low == near_low(cand[b], static_cast<int>(j));
    // they part below the common ancestor pa[j].node (same node in both paths, different sides)
    return pa[j].low == near_low(pa[j].node, static_cast<int>(j));
  };
  size_t best = 0;
  for (size_t k = 1; k < cand.size(); ++k)
    if (before(k, best)) best = k;
  return cand[best];
}

// Uniform grid over node positions. One 64-byte, cache-line-aligned bucket per cell holding up to
// 5 entries inline (x, y, entry id): a nearest query touches 9 independent cache lines instead of
// chasing linked lists. TRG nodes are pairwise >= robot_size apart (a node is only created when its
// nearest neighbour is at least that far, trg.cpp:414-421), so with cell = 1.5 * robot_size a cell
// rarely holds more than 4; anything beyond spills into per-cell overflow chains.
class NodeGrid {
 public:
  void configure(float x0, float y0, float x1, float y1, float cell) {
    cell_ = cell;
    inv_ = 1.0f / cell;
    x0_ = x0 - 2.f * cell;
    y0_ = y0 - 2.f * cell;
    w_ = static_cast<int>(std::floor((x1 - x0_) * inv_)) + 4;
    h_ = static_cast<int>(std::floor((y1 - y0_) * inv_)) + 4;
    bw_ = (w_ + kTile - 1) / kTile;
    const int bh = (h_ + kTile - 1) / kTile;
    cells_.assign(static_cast<size_t>(bw_) * bh * kTile * kTile, Cell{});
    over_.clear();
    count_ = 0;
    stale_ = false;
  }
  // prefetch the buckets a nearest() query at (qx, qy) will read first
  void prefetch(float qx, float qy) const {
    if (cells_.empty()) return;
    const int qcx = cx(qx), qcy = cy(qy);
    for (int yy = qcy - 1; yy <= qcy + 1; ++yy)
      for (int xx = qcx - 1; xx <= qcx + 1; ++xx)
        if (xx >= 0 && yy >= 0 && xx < w_ && yy < h_) __builtin_prefetch(&cells_[cidx(xx, yy)], 0, 1);
  }
  bool configured() const { return w_ > 0; }
  // Emptying is deferred: queries on an empty grid return at once, the next insert() wipes the buckets, and
  // rebuild() wipes them band by band on its helper threads (31 MB for a 300 m map: 3 ms if done serially).
  void clear() {
    if (count_ == 0 && over_.empty() && !stale_) return;
    stale_ = true;
    count_ = 0;
  }
  size_t size() const { return count_; }

  // The grid over n points at once (entry i = point i, as n insert() calls would number them): helper threads
  // own contiguous bands of bucket tiles, wipe them and insert the points that fall inside.
  void rebuild(const float* xy, int n, int threads) {
    over_.clear();
    const size_t tiles = cells_.size() / (kTile * kTile);
    threads = static_cast<int>(std::max<size_t>(1, std::min<size_t>(static_cast<size_t>(std::max(1, threads)), tiles)));
    std::mutex over_mx;
    auto band = [&](int t) {
      const size_t c0 = tiles * t / threads * (kTile * kTile), c1 = tiles * (t + 1) / threads * (kTile * kTile);
      std::fill(cells_.begin() + c0, cells_.begin() + c1, Cell{});
      for (int i = 0; i < n; ++i) {
        const float x = xy[2 * i], y = xy[2 * i + 1];
        const size_t ci = cidx(cx(x), cy(y));
        if (ci < c0 || ci >= c1) continue;
        Cell& c = cells_[ci];
        if (c.n < kInline) {
          c.x[c.n] = x; c.y[c.n] = y; c.id[c.n] = i;
          ++c.n;
        } else {
          std::lock_guard<std::mutex> lk(over_mx);
          over_.push_back({x, y, i, c.over});
          c.over = static_cast<int>(over_.size()) - 1;
        }
      }
    };
    std::vector<std::future<void>> helpers;
    for (int t = 1; t < threads; ++t) helpers.push_back(std::async(std::launch::async, band, t));
    band(0);
    for (auto& h : helpers) h.get();
    stale_ = false;
    count_ = static_cast<size_t>(n);
  }

  // entries are numbered in insertion order (0, 1, 2, ...)
  int insert(float x, float y) {
    if (stale_) {
      std::fill(cells_.begin(), cells_.end(), Cell{});
      over_.clear();
      stale_ = false;
    }
    const int id = static_cast<int>(count_++);
    Cell& c = cells_[cidx(cx(x), cy(y))];
    if (c.n < kInline) {
      c.x[c.n] = x; c.y[c.n] = y; c.id[c.n] = id;
      ++c.n;
    } else {
      over_.push_back({x, y, id, c.over});
      c.over = static_cast<int>(over_.size()) - 1;
    }
    return id;
  }

  struct Nearest { int entry; float d2; bool tie; };
  // exact global nearest (float dist^2, strict <). `tie` = some other entry has the identical
  // dist^2 — the caller then asks the OrderTree2D which one the reference would return.
  Nearest nearest(float qx, float qy) const {
    Nearest best{-1, std::numeric_limits<float>::infinity(), false};
    if (count_ == 0) return best;
    const int qcx = cx(qx), qcy = cy(qy);
    const int maxr = std::max(w_, h_);
    const float fuzz = 4e-6f * (std::fabs(qx) + std::fabs(qy) + cell_ * static_cast<float>(w_ + h_));
    for (int R = 1; R <= maxr; ++R) {
      // first pass scans the 3x3 block, later passes the ring at Chebyshev distance R
      if (R == 1) {
        for (int yy = qcy - 1; yy <= qcy + 1; ++yy)
          for (int xx = qcx - 1; xx <= qcx + 1; ++xx) scan_cell(xx, yy, qx, qy, best);
      } else {
        scan_ring(qcx, qcy, R, qx, qy, best);
      }
      const float g = static_cast<float>(R) * cell_ * 0.9999f - fuzz;
      // cells outside the scanned block are at least R whole cells away from the query's cell
      if (best.entry >= 0 && g > 0.f && best.d2 <= g * g) break;
      if (qcx - R <= 0 && qcy - R <= 0 && qcx + R >= w_ - 1 && qcy + R >= h_ - 1) break;
    }
    return best;
  }

  // all entries with fl(dx^2+dy^2) <= fl(r^2), unordered
  template <class F>
  void for_each_in_range(float qx, float qy, float r, F&& f) const {
    if (count_ == 0) return;
    const float r2 = r * r;
    const float rr = r * 1.0001f + 1e-5f + 4e-6f * (std::fabs(qx) + std::fabs(qy));
    const int cx0 = cx(qx - rr), cx1 = cx(qx + rr), cy0 = cy(qy - rr), cy1 = cy(qy + rr);
    for (int yy = cy0; yy <= cy1; ++yy)
      for (int xx = cx0; xx <= cx1; ++xx)
        visit(cells_[cidx(xx, yy)], [&](float x, float y, int e) {
          const float dx = x - qx, dy = y - qy;
          float d2 = 0.f;
          d2 += dx * dx;
          d2 += dy * dy;
          if (d2 <= r2) f(e);
        });
  }
  // entries with fl(dx^2+dy^2) <= d2max, searching the cells within `reach` of the query
  template <class F>
  void for_each_within_d2(float qx, float qy, float d2max, float reach, F&& f) const {
    if (count_ == 0) return;
    const float rr = reach * 1.0001f + 1e-5f + 4e-6f * (std::fabs(qx) + std::fabs(qy));
    const int cx0 = cx(qx - rr), cx1 = cx(qx + rr), cy0 = cy(qy - rr), cy1 = cy(qy + rr);
    for (int yy = cy0; yy <= cy1; ++yy)
      for (int xx = cx0; xx <= cx1; ++xx)
        visit(cells_[cidx(xx, yy)], [&](float x, float y, int e) {
          const float dx = x - qx, dy = y - qy;
          float d2 = 0.f;
          d2 += dx * dx;
          d2 += dy * dy;
          if (d2 <= d2max) f(e);
        });
  }
  int count_in_range(float qx, float qy, float r) const {
    int c = 0;
    for_each_in_range(qx, qy, r, [&](int) { ++c; });
    return c;
  }

 private:
  static constexpr int kInline = 4;
  // buckets are stored in 8x8 tiles (64 buckets = 4 KB = one page): a 3x3 neighbourhood lives in
  // at most four pages instead of three rows tens of KB apart
  static constexpr int kTile = 8;
  size_t cidx(int xx, int yy) const {
    return (static_cast<size_t>(yy / kTile) * bw_ + static_cast<size_t>(xx / kTile)) * (kTile * kTile) +
           static_cast<size_t>((yy % kTile) * kTile + (xx % kTile));
  }
  struct alignas(64) Cell {
    uint32_t n = 0;
    int32_t over = -1;  // head of the overflow chain (index into over_)
    float x[kInline] = {0, 0, 0, 0};
    float y[kInline] = {0, 0, 0, 0};
    int32_t id[kInline] = {0, 0, 0, 0};
  };
  struct Over { float x, y; int32_t id; int32_t next; };

  template <class F>
  void visit(const Cell& c, F&& f) const {
    for (uint32_t k = 0; k < c.n; ++k) f(c.x[k], c.y[k], c.id[k]);
    for (int o = c.over; o >= 0; o = over_[o].next) f(over_[o].x, over_[o].y, over_[o].id);
  }
  int cx(float x) const {
    const float f = std::floor((x - x0_) * inv_);
    if (!(f > 0.f)) return 0;
    if (f >= static_cast<float>(w_)) return w_ - 1;
    return static_cast<int>(f);
  }
  int cy(float y) const {
    const float f = std::floor((y - y0_) * inv_);
    if (!(f > 0.f)) return 0;
    if (f >= static_cast<float>(h_)) return h_ - 1;
    return static_cast<int>(f);
  }
  void scan_cell(int xx, int yy, float qx, float qy, Nearest& best) const {
    if (xx < 0 || yy < 0 || xx >= w_ || yy >= h_) return;
    visit(cells_[cidx(xx, yy)], [&](float x, float y, int e) {
      const float dx = x - qx, dy = y - qy;
      float d2 = 0.f;
      d2 += dx * dx;
      d2 += dy * dy;
      if (d2 < best.d2) {
        best.d2 = d2;
        best.entry = e;
        best.tie = false;
      } else if (d2 == best.d2 && e != best.entry) {
        best.tie = true;
      }
    });
  }
  void scan_ring(int qcx, int qcy, int ring, float qx, float qy, Nearest& best) const {
    for (int xx = qcx - ring; xx <= qcx + ring; ++xx) {
      scan_cell(xx, qcy - ring, qx, qy, best);
      scan_cell(xx, qcy + ring, qx, qy, best);
    }
    for (int yy = qcy - ring + 1; yy <= qcy + ring - 1; ++yy) {
      scan_cell(qcx - ring, yy, qx, qy, best);
      scan_cell(qcx + ring, yy, qx, qy, best);
    }
  }

  float cell_ = 1.f, inv_ = 1.f, x0_ = 0.f, y0_ = 0.f;
  int w_ = 0, h_ = 0, bw_ = 0;
  size_t count_ = 0;
  bool stale_ = false;  // buckets still hold entries of before the last clear()
  std::vector<Cell> cells_;
  std::vector<Over> over_;
};

// Nodes created while the current batch is being committed: they are not yet in the device node
// grid that produced the batch's nearest-node candidates, so the commit merges those candidates
// with an exact nearest search over this small, cache-resident hash grid.
class ChunkTable {
 public:
  void configure(float x0, float y0, float cell) { x0_ = x0; y0_ = y0; cell_ = cell; inv_ = 1.0f / cell; }
  bool empty() const { return ent_.empty(); }
  void clear() {
    for (uint32_t h : touched_) { head_[h] = -1; occ_[h >> 6] = 0; }
    touched_.clear();
    ent_.clear();
  }
  // forget nodes with seq < min_seq (they are on the device grid by now); entries are in seq order
  void prune(int min_seq) {
    size_t first = 0;
    while (first < ent_.size() && ent_[first].seq < min_seq) ++first;
    if (first == 0) return;
    std::vector<E> keep(ent_.begin() + first, ent_.end());
    clear();
    for (const E& e : keep) insert(e.x, e.y, e.seq);
  }
  void insert(float x, float y, int seq) {
    const int cx = cc(x, x0_), cy = cc(y, y0_);
    const uint32_t h = hash(cx, cy);
    if (head_[h] < 0) touched_.push_back(h);
    occ_[h >> 6] |= 1ull << (h & 63);
    ent_.push_back({x, y, seq, head_[h], cx, cy});
    head_[h] = static_cast<int32_t>(ent_.size()) - 1;
  }
  // improve (d2, seq, tie) with any stored node that is strictly nearer; equal distance -> tie
  void refine(float qx, float qy, float& d2, int& seq, bool& tie) const {
    if (ent_.empty()) return;
    if (!(d2 < std::numeric_limits<float>::infinity())) {
      for (const E& e : ent_) consider(e, qx, qy, d2, seq, tie);
      return;
    }
    const float fuzz = 1e-5f + 4e-6f * (std::fabs(qx) + std::fabs(qy));
    {
      // the device candidate bounds the search: only cells the disc of radius sqrt(d2) touches
      // (inflated so that a node at exactly the same distance is seen and flagged as a tie) -
      // usually one or two cells instead of nine
      const float reach = std::sqrt(d2) * 1.00001f + fuzz;
      const int cx0 = cc(qx - reach, x0_), cx1 = cc(qx + reach, x0_);
      const int cy0 = cc(qy - reach, y0_), cy1 = cc(qy + reach, y0_);
      if ((cx1 - cx0 + 1) * (cy1 - cy0 + 1) <= 9) {
        for (int yy = cy0; yy <= cy1; ++yy)
          for (int xx = cx0; xx <= cx1; ++xx) scan(xx, yy, qx, qy, d2, seq, tie);
        return;
      }
    }
    const int qcx = cc(qx, x0_), qcy = cc(qy, y0_);
    for (int R = 1;; ++R) {
      if (R == 1) {
        for (int yy = qcy - 1; yy <= qcy + 1; ++yy)
          for (int xx = qcx - 1; xx <= qcx + 1; ++xx) scan(xx, yy, qx, qy, d2, seq, tie);
      } else {
        for (int xx = qcx - R; xx <= qcx + R; ++xx) { scan(xx, qcy - R, qx, qy, d2, seq, tie); scan(xx, qcy + R, qx, qy, d2, seq, tie); }
        for (int yy = qcy - R + 1; yy <= qcy + R - 1; ++yy) { scan(qcx - R, yy, qx, qy, d2, seq, tie); scan(qcx + R, yy, qx, qy, d2, seq, tie); }
      }
      // everything outside the scanned block is at least R whole cells away
      const float g = static_cast<float>(R) * cell_ * 0.9999f - fuzz;
      if (g > 0.f && d2 <= g * g) break;
      if (R > 64) {  // pathological (nearest far away): finish exhaustively
        for (const E& e : ent_) consider(e, qx, qy, d2, seq, tie);
        break;
      }
    }
  }

 private:
  struct E { float x, y; int seq; int32_t next; int cx, cy; };
  static constexpr uint32_t kMask = (1u << 14) - 1;
  // toroidal 128 x 128 tile of cells: neighbouring cells are neighbouring buckets (the 2 x 2 block a
  // query usually probes sits in two cache lines) and there is nothing to multiply; cells 128 apart
  // share a bucket and are told apart by (cx, cy)
  static uint32_t hash(int cx, int cy) {
    return ((static_cast<uint32_t>(cy) & 127u) << 7) | (static_cast<uint32_t>(cx) & 127u);
  }
  int cc(float v, float o) const { return static_cast<int>(std::floor((v - o) * inv_)); }
  static void consider(const E& e, float qx, float qy, float& d2, int& seq, bool& tie) {
    const float dx = e.x - qx, dy = e.y - qy;
    float v = 0.f;
    v += dx * dx;
    v += dy * dy;
    if (v < d2) { d2 = v; seq = e.seq; tie = false; }
    else if (v == d2 && e.seq != seq) tie = true;
  }
  void scan(int cx, int cy, float qx, float qy, float& d2, int& seq, bool& tie) const {
    const uint32_t h = hash(cx, cy);
    if (!((occ_[h >> 6] >> (h & 63)) & 1ull)) return;  // 2 KB bitmap: almost every probe ends here
    for (int32_t i = head_[h]; i >= 0; i = ent_[i].next)
      if (ent_[i].cx == cx && ent_[i].cy == cy) consider(ent_[i], qx, qy, d2, seq, tie);
  }
  float x0_ = 0.f, y0_ = 0.f, cell_ = 1.f, inv_ = 1.f;
  std::vector<int32_t> head_ = std::vector<int32_t>(kMask + 1, -1);
  std::vector<uint64_t> occ_ = std::vector<uint64_t>((kMask + 1) / 64, 0);
  std::vector<uint32_t> touched_;
  std::vector<E> ent_;
};

}  // namespace trg_b200
