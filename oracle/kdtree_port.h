// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product path.
//
// Restatement of the reference's 2-D, float, insertion-order kd-tree
// (cpp/trg_planner/core/trg_planner/src/kdtree/kdtree.c) on flat arrays.
// Only the semantics the TRG path observes are restated:
//   insert            kdtree.c:167-209  (left iff pos[dir] <  node.pos[dir]; dir alternates)
//   range query       kdtree.c:270-301  (inclusive `<=` on float dx*dx+dy*dy; pre-order visit,
//                                        query-side child first; other side iff |dx| < range)
//   result order      kdtree.c:759-777  (head insertion => iteration = REVERSE visit order)
//   single nearest    kdtree.c:303-417  (strict `<`; nearer subtree, node, farther subtree;
//                                        bounding-box pruning; initial best = root)
// It is validated against the verbatim reference kdtree.c (oracle/_ref) by
// tests/test_oracle_cpu.py (the "refkd" and "ref" builds of oracle/Makefile against the golden fixtures, plus
// restatement == reference on fresh maps) and can be swapped for it with -DORACLE_USE_REF_KDTREE.
#ifndef ORACLE_KDTREE_PORT_H_
#define ORACLE_KDTREE_PORT_H_

#include <cmath>
#include <cstdint>
#include <vector>

namespace kdport {

struct Tree2 {
  std::vector<float> px, py;
  std::vector<int32_t> left, right;
  std::vector<uint8_t> dir;
  std::vector<int64_t> data;  // payload (index into the caller's array)
  float rmin[2] = {0, 0}, rmax[2] = {0, 0};
  bool has_rect = false;

  void clear() {
    px.clear(); py.clear(); left.clear(); right.clear(); dir.clear(); data.clear();
    has_rect = false;
  }
  size_t size() const { return px.size(); }

  void insert(float x, float y, int64_t payload) {
    int32_t id = static_cast<int32_t>(px.size());
    const float pos[2] = {x, y};
    int d = 0;
    if (id != 0) {
      int32_t cur = 0;
      for (;;) {
        const float np = dir[cur] ? py[cur] : px[cur];
        int nd = (dir[cur] + 1) % 2;
        int32_t* child = (pos[dir[cur]] < np) ? &left[cur] : &right[cur];
        if (*child < 0) {
          *child = id;
          d = nd;
          break;
        }
        cur = *child;
      }
    }
    px.push_back(x); py.push_back(y); left.push_back(-1); right.push_back(-1);
    dir.push_back(static_cast<uint8_t>(d)); data.push_back(payload);
    if (!has_rect) {
      rmin[0] = rmax[0] = x; rmin[1] = rmax[1] = y; has_rect = true;
    } else {
      if (x < rmin[0]) rmin[0] = x;
      if (x > rmax[0]) rmax[0] = x;
      if (y < rmin[1]) rmin[1] = y;
      if (y > rmax[1]) rmax[1] = y;
    }
  }

  // Returns payloads in the order the reference's result-set iterator yields them
  // (reverse of the DFS visit order).
  void range(float x, float y, float range, std::vector<int64_t>* out) const {
    out->clear();
    if (px.empty()) return;
    const float r2 = range * range;
    // explicit stack emulating the recursion: a node is entered (far_pending=false), its
    // query-side child is explored, then (far_pending=true) the other side is considered.
    struct Item { int32_t node; bool far_pending; float dx; };
    std::vector<Item> stack;
    stack.push_back({0, false, 0.f});
    while (!stack.empty()) {
      Item it = stack.back();
      stack.pop_back();
      if (it.far_pending) {
        // returning from the near child of it.node: maybe descend into the far child
        if (std::fabs(it.dx) < range) {
          int32_t far = it.dx <= 0.0f ? right[it.node] : left[it.node];
          if (far >= 0) stack.push_back({far, false, 0.f});
        }
        continue;
      }
      int32_t n = it.node;
      float ddx = px[n] - x, ddy = py[n] - y;
      float dist_sq = 0;
      dist_sq += ddx * ddx;
      dist_sq += ddy * ddy;
      if (dist_sq <= r2) out->push_back(data[n]);
      float dx = (dir[n] ? y : x) - (dir[n] ? py[n] : px[n]);
      stack.push_back({n, true, dx});
      int32_t near = dx <= 0.0f ? left[n] : right[n];
      if (near >= 0) stack.push_back({near, false, 0.f});
    }
    // head insertion => reverse
    for (size_t i = 0, j = out->size(); i + 1 < j; ++i, --j) std::swap((*out)[i], (*out)[j - 1]);
  }

  // single nearest; returns payload, or -1 on an empty tree
  int64_t nearest(float x, float y) const {
    if (px.empty()) return -1;
    float rect_min[2] = {rmin[0], rmin[1]}, rect_max[2] = {rmax[0], rmax[1]};
    const float pos[2] = {x, y};
    int32_t result = 0;
    float best = 0;
    best += (px[0] - x) * (px[0] - x);
    best += (py[0] - y) * (py[0] - y);
    nearest_rec(0, pos, &result, &best, rect_min, rect_max);
    return data[result];
  }

 private:
  static float rect_dist_sq(const float* mn, const float* mx, const float* pos) {
    float r = 0;
    for (int i = 0; i < 2; ++i) {
      if (pos[i] < mn[i]) r += (mn[i] - pos[i]) * (mn[i] - pos[i]);
      else if (pos[i] > mx[i]) r += (mx[i] - pos[i]) * (mx[i] - pos[i]);
    }
    return r;
  }
  // Recursion depth equals tree depth; inputs are shuffled clouds / BFS-ordered graph
  // nodes (depth ~ O(100..1000)), well inside the default stack.
  void nearest_rec(int32_t n, const float* pos, int32_t* result, float* best,
                   float* mn, float* mx) const {
    const int d = dir[n];
    const float npos = d ? py[n] : px[n];
    float dummy = pos[d] - npos;
    int32_t nearer, farther;
    float *near_c, *far_c;
    if (dummy <= 0) {
      nearer = left[n]; farther = right[n]; near_c = mx + d; far_c = mn + d;
    } else {
      nearer = right[n]; farther = left[n]; near_c = mn + d; far_c = mx + d;
    }
    if (nearer >= 0) {
      dummy = *near_c;
      *near_c = npos;
      nearest_rec(nearer, pos, result, best, mn, mx);
      *near_c = dummy;
    }
    float dist_sq = 0;
    dist_sq += (px[n] - pos[0]) * (px[n] - pos[0]);
    dist_sq += (py[n] - pos[1]) * (py[n] - pos[1]);
    if (dist_sq < *best) {
      *result = n;
      *best = dist_sq;
    }
    if (farther >= 0) {
      dummy = *far_c;
      *far_c = npos;
      if (rect_dist_sq(mn, mx, pos) < *best) nearest_rec(farther, pos, result, best, mn, mx);
      *far_c = dummy;
    }
  }
};

}  // namespace kdport
#endif  // ORACLE_KDTREE_PORT_H_
