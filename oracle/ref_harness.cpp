// ORACLE — TEST INFRASTRUCTURE ONLY (see trg_oracle.h).
//
// C facade (the `orc_*` entry points of trg_oracle.h) over the reference's OWN, UNMODIFIED
//   /root/reference/cpp/trg_planner/core/trg_planner/src/graph/trg.cpp
//   /root/reference/cpp/trg_planner/core/trg_planner/src/kdtree/kdtree.c
// which oracle/Makefile compiles where they lie, against the stand-in headers of oracle/shim/
// (Eigen, PCL, OpenCV, yaml-cpp are absent from this image; nlohmann/json.hpp is the real one
// found under site-packages), into oracle/_ref/libtrg_ref.so. No reference source is copied:
// this file only *calls* the reference class through its public interface and reads its
// protected members through a subclass (the way SURVEY.md §0-2 suggests reseeding `gen_`).
//
// Every decision the reference takes (RNG stream, collision tests, node acceptance, kd-tree order,
// unordered_map order, A*) therefore runs in the reference's own code. The only stand-in
// arithmetic is the float summation order of the covariance (shim/Eigen/Core header note).
#include <dlfcn.h>
#include <setjmp.h>

#include <chrono>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "trg_planner/include/graph/trg.h"

#include "trg_oracle.h"

namespace {

using Clock = std::chrono::steady_clock;
inline double secs_since(Clock::time_point t0) {
  return std::chrono::duration<double>(Clock::now() - t0).count();
}

// Subclass: access to the protected state of the reference class (trg.h:100-146), nothing overridden.
class RefTRG : public TRG {
 public:
  explicit RefTRG(const OrcParams& p)
      : TRG(p.is_verbose != 0, p.expand_dist, p.robot_size, p.sample_num, p.height_threshold,
            p.collision_threshold, p.update_collision_threshold, p.safety_factor, p.goal_tolerance) {
    reseed(0u);
  }
  void reseed(uint32_t s) {  // the reference seeds from std::random_device (trg.cpp:20)
    gen_.seed(s);
    distr_.reset();
    gen0_ = gen_;
    draws_base_ = 0;
  }
  // generator calls since the last reseed (uniform_real_distribution<float> takes one per draw)
  int64_t draws() {
    while (!(gen0_ == gen_)) {
      gen0_.discard(1);
      ++draws_base_;
      if (draws_base_ > (int64_t(1) << 40)) return -1;
    }
    return draws_base_;
  }
  trgStruct& graphOf(const char* type) { return *trgMap_.at(std::string(type)); }
  bool goalKnown() const { return goal_.isKnown; }
  const Node* goalNode() const { return goal_.node; }
  float robotSize() const { return param_.robot_size; }

  std::map<std::string, double> secs_;

 private:
  std::mt19937 gen0_;
  int64_t draws_base_ = 0;
};

inline RefTRG* H(void* h) { return static_cast<RefTRG*>(h); }

PointCloudPtr make_cloud(const float* xyz, int64_t n) {
  PointCloudPtr c(new pcl::PointCloud<PtsDefault>());
  c->points.resize(static_cast<size_t>(n));
  for (int64_t i = 0; i < n; ++i) {
    c->points[i].x = xyz[3 * i];
    c->points[i].y = xyz[3 * i + 1];
    c->points[i].z = xyz[3 * i + 2];
  }
  c->width = static_cast<uint32_t>(n);
  c->height = 1;
  return c;
}

// TRG::initGraph calls exit(1) when no root can be placed (trg.cpp:49-52). The library is linked
// with -Bsymbolic-functions, so that call binds to the `exit` below: while a harness call is in
// flight it jumps back to the harness (which releases TRG::mtx.graph through the public
// unlockGraph() and reports -1) instead of ending the test process; otherwise it is libc's exit.
thread_local jmp_buf* g_exit_jmp = nullptr;

}  // namespace

extern "C" void exit(int code) {
  if (g_exit_jmp) longjmp(*g_exit_jmp, 1);
  using exit_fn = void (*)(int);
  exit_fn real = reinterpret_cast<exit_fn>(dlsym(RTLD_NEXT, "exit"));
  if (real) real(code);
  _exit(code);
}

// The library is compiled with -fvisibility=hidden: the reference's class is called `TRG` like the
// product's drop-in class, and both libraries live in one test process — only the orc_* facade may
// be visible, or the dynamic linker would resolve one library's TRG::TRG(...) to the other's.
#pragma GCC visibility push(default)
extern "C" {

void* orc_create(const OrcParams* p) { return new RefTRG(*p); }
void orc_destroy(void* h) { delete H(h); }
void orc_seed(void* h, uint32_t seed) { H(h)->reseed(seed); }

int orc_set_global_map(void* h, const float* xyz, int64_t n) {
  PointCloudPtr c = make_cloud(xyz, n);
  auto t0 = Clock::now();
  H(h)->setGlobalMap(c);
  H(h)->secs_["set_global_map"] = secs_since(t0);
  return 0;
}
int orc_set_local_map(void* h, float sx, float sy, const float* xyz, int64_t n) {
  PointCloudPtr c = make_cloud(xyz, n);
  auto t0 = Clock::now();
  H(h)->setLocalMap(Eigen::Vector2f(sx, sy), c);
  H(h)->secs_["set_local_map"] = secs_since(t0);
  return 0;
}
// -1 = the reference gave up placing a root ("Failed to generate root node" + exit(1), trg.cpp:49-52)
int orc_init_graph(void* h, int is_pre_map, float sx, float sy, float sz) {
  if (H(h)->graphOf("global").cloud_map->size() == 0) return -2;
  auto t0 = Clock::now();
  jmp_buf jb;
  if (setjmp(jb) != 0) {
    g_exit_jmp = nullptr;
    H(h)->unlockGraph();  // initGraph's lock_guard (trg.cpp:37) never unwound
    H(h)->secs_["init_graph"] = secs_since(t0);
    return -1;
  }
  g_exit_jmp = &jb;
  H(h)->initGraph(is_pre_map != 0, Eigen::Vector3f(sx, sy, sz));
  g_exit_jmp = nullptr;
  H(h)->secs_["init_graph"] = secs_since(t0);
  return 0;
}
int orc_update_graph(void* h) {
  auto t0 = Clock::now();
  H(h)->updateGraph();
  H(h)->secs_["update_graph"] = secs_since(t0);
  return 0;
}

int orc_graph_counts(void* h, const char* type, int64_t* n_nodes, int64_t* n_edges) {
  auto& g = H(h)->graphOf(type);
  int64_t e = 0;
  for (auto& kv : g.nodes) e += static_cast<int64_t>(kv.second->edges_.size());
  *n_nodes = static_cast<int64_t>(g.nodes.size());
  *n_edges = e;
  return 0;
}

int orc_graph_export(void* h, const char* type, int32_t* iter_ids, int32_t* ids_sorted,
                     float* pos_xyz, int32_t* state, int64_t* row_ptr, int32_t* col,
                     float* weight, float* dist) {
  auto& g = H(h)->graphOf(type);
  std::vector<int> ids;
  int64_t k = 0;
  for (auto& kv : g.nodes) {
    if (iter_ids) iter_ids[k] = kv.first;
    ids.push_back(kv.first);
    ++k;
  }
  std::sort(ids.begin(), ids.end());
  int64_t e = 0;
  for (size_t i = 0; i < ids.size(); ++i) {
    const TRG::Node* n = g.nodes.at(ids[i]);
    if (ids_sorted) ids_sorted[i] = ids[i];
    if (pos_xyz) {
      pos_xyz[3 * i] = n->pos_.x();
      pos_xyz[3 * i + 1] = n->pos_.y();
      pos_xyz[3 * i + 2] = n->pos_.z();
    }
    if (state) state[i] = static_cast<int32_t>(n->state_);
    if (row_ptr) row_ptr[i] = e;
    for (const TRG::Edge* ed : n->edges_) {
      if (col) col[e] = ed->dst_id_;
      if (weight) weight[e] = ed->weight_;
      if (dist) dist[e] = ed->dist_;
      ++e;
    }
  }
  if (row_ptr) row_ptr[ids.size()] = e;
  return 0;
}

// planSafePath returns positions only; node ids are recovered by exact position match (a node's
// pos_ is copied verbatim into the path, trg.cpp:653). n_expanded is not observable (-1).
int orc_plan(void* h, float sx, float sy, float gx, float gy, float gz, float* path_xyz,
             int32_t* node_ids, int max_pts, int* n_pts, float* direct_dist, float* path_length,
             float* avg_risk, int* goal_known, int64_t* n_expanded) {
  std::vector<Eigen::Vector3f> path;
  Eigen::Vector2f s(sx, sy);
  Eigen::Vector3f g3(gx, gy, gz);
  float dd = 0, pl = 0, ar = 0;
  if (H(h)->graphOf("global").nodes.empty()) return -2;
  auto t0 = Clock::now();
  bool ok = H(h)->planSafePath(s, g3, path, dd, pl, ar);
  H(h)->secs_["plan"] = secs_since(t0);
  if (direct_dist) *direct_dist = dd;
  if (path_length) *path_length = pl;
  if (avg_risk) *avg_risk = ar;
  if (goal_known) *goal_known = H(h)->goalKnown() ? 1 : 0;
  if (n_expanded) *n_expanded = -1;
  int n = static_cast<int>(path.size());
  if (n_pts) *n_pts = n;
  auto& g = H(h)->graphOf("global");
  for (int i = 0; i < n && i < max_pts; ++i) {
    if (path_xyz) {
      path_xyz[3 * i] = path[i].x();
      path_xyz[3 * i + 1] = path[i].y();
      path_xyz[3 * i + 2] = path[i].z();
    }
    if (node_ids) {
      // nearest node of the node tree at the point itself: distance 0 = the node (or an exact twin)
      kdres* res = kd_nearest2(g.node_tree, path[i].x(), path[i].y());
      const TRG::Node* nn = reinterpret_cast<TRG::Node*>(kd_res_item_data(res));
      kd_res_free(res);
      node_ids[i] = (nn->pos_.x() == path[i].x() && nn->pos_.y() == path[i].y()) ? nn->id_ : -1;
    }
  }
  return ok ? 1 : 0;
}

int orc_refine_path(void* h, const float* in_xyz, int n_in, float* out_xyz, int* n_out) {
  std::vector<Eigen::Vector3f> in, out;
  for (int i = 0; i < n_in; ++i) in.emplace_back(in_xyz[3 * i], in_xyz[3 * i + 1], in_xyz[3 * i + 2]);
  if (n_in < 1) {  // the reference's `in_path.size() - 1` wraps on an empty path (trg.cpp:696)
    *n_out = 0;
    return 0;
  }
  H(h)->refinePath(in, out);
  *n_out = static_cast<int>(out.size());
  for (size_t i = 0; i < out.size(); ++i) {
    out_xyz[3 * i] = out[i].x();
    out_xyz[3 * i + 1] = out[i].y();
    out_xyz[3 * i + 2] = out[i].z();
  }
  return 0;
}

int orc_check_reached(void* h, float x, float y) {
  Eigen::Vector2f p(x, y);
  return H(h)->checkReadched(p) ? 1 : 0;
}
int orc_check_replan(void* h, float x, float y, const float* path_xyz, int n_path) {
  std::vector<Eigen::Vector3f> path;
  for (int i = 0; i < n_path; ++i) path.emplace_back(path_xyz[3 * i], path_xyz[3 * i + 1], path_xyz[3 * i + 2]);
  Eigen::Vector2f p(x, y);
  return H(h)->checkReplan(p, path) ? 1 : 0;
}

int orc_is_collision_batch(void* h, const char* type, const float* xy, int64_t n, float threshold,
                           uint8_t* out) {
  for (int64_t i = 0; i < n; ++i) {
    Eigen::Vector2f p(xy[2 * i], xy[2 * i + 1]);
    out[i] = H(h)->isCollision(p, std::string(type), threshold) ? 1 : 0;
  }
  return 0;
}

int orc_range_count_batch(void* h, const char* type, const float* xy, int64_t n, float radius,
                          int32_t* out) {
  auto& g = H(h)->graphOf(type);
  for (int64_t i = 0; i < n; ++i) {
    kdres* res = kd_nearest_range2(g.map_tree, xy[2 * i], xy[2 * i + 1], radius);
    out[i] = kd_res_size(res);
    kd_res_free(res);
  }
  return 0;
}

// trg.cpp:244-246 on its own
int orc_nearest_z_batch(void* h, const char* type, const float* xy, int64_t n, float* z_out,
                        int64_t* idx_out, uint8_t* tie_out) {
  auto& g = H(h)->graphOf(type);
  if (g.cloud_map->size() == 0) return -2;
  const PtsDefault* base = g.cloud_map->points.data();
  for (int64_t i = 0; i < n; ++i) {
    float x = xy[2 * i], y = xy[2 * i + 1];
    kdres* res = kd_nearest2(g.map_tree, x, y);
    const PtsDefault* pt = reinterpret_cast<PtsDefault*>(kd_res_item_data(res));
    kd_res_free(res);
    if (z_out) z_out[i] = pt->z;
    if (idx_out) idx_out[i] = static_cast<int64_t>(pt - base);
    if (tie_out) {
      float dx = pt->x - x, dy = pt->y - y;
      float d2 = 0;
      d2 += dx * dx;
      d2 += dy * dy;
      float rad = sqrtf(d2) * 1.0001f + 1e-6f;
      kdres* rr = kd_nearest_range2(g.map_tree, x, y, rad);
      int cnt = 0;
      while (!kd_res_end(rr)) {
        const PtsDefault* q = reinterpret_cast<PtsDefault*>(kd_res_item_data(rr));
        float ex = q->x - x, ey = q->y - y;
        float e2 = 0;
        e2 += ex * ex;
        e2 += ey * ey;
        if (e2 == d2) cnt++;
        kd_res_next(rr);
      }
      kd_res_free(rr);
      tie_out[i] = cnt > 1 ? 1 : 0;
    }
  }
  return 0;
}

// TRG::wireEdge (trg.cpp:254-370) on two throw-away nodes without edges: the duplicate checks
// (:255-267) pass trivially, so the call is the pure geometric part. The reference does not say
// WHERE it returned: stage is 0 (edge created) or 255 (returned early); weight64 / npts are not
// observable (NaN / -1).
int orc_edge_eval_batch(void* h, const char* type, const float* p1, const float* p2, int64_t n,
                        uint8_t* stage, float* weight, double* weight64, float* dist,
                        int32_t* npts) {
  for (int64_t i = 0; i < n; ++i) {
    Eigen::Vector2f a(p1[3 * i], p1[3 * i + 1]), b(p2[3 * i], p2[3 * i + 1]);
    TRG::Node n1(0, a, p1[3 * i + 2], TRG::NodeState::Valid);
    TRG::Node n2(1, b, p2[3 * i + 2], TRG::NodeState::Valid);
    H(h)->wireEdge(&n1, &n2, std::string(type));
    bool ok = !n1.edges_.empty();
    if (stage) stage[i] = ok ? 0 : 255;
    if (weight) weight[i] = ok ? n1.edges_[0]->weight_ : 0.f;
    if (dist) dist[i] = ok ? n1.edges_[0]->dist_ : 0.f;
    if (weight64) weight64[i] = std::numeric_limits<double>::quiet_NaN();
    if (npts) npts[i] = -1;
    for (auto* e : n1.edges_) delete e;
    for (auto* e : n2.edges_) delete e;
  }
  return 0;
}

int orc_is_frontier_batch(void* h, const float* xy, int64_t n, uint8_t* out) {
  for (int64_t i = 0; i < n; ++i) {
    Eigen::Vector2f p(xy[2 * i], xy[2 * i + 1]);
    out[i] = H(h)->isFrontier(p) ? 1 : 0;
  }
  return 0;
}

double orc_last_seconds(void* h, const char* what) {
  auto it = H(h)->secs_.find(what);
  return it == H(h)->secs_.end() ? -1.0 : it->second;
}
int64_t orc_stat(void* h, const char* what) {
  if (std::string(what) == "rng_draws") return H(h)->draws();
  return -1;  // the other counters need hooks inside the reference's functions
}

// TRG::saveGraph / loadPrebuiltGraph (trg.cpp:130-177 / 66-128). The library is built with
// TRG_DIR="/" so that the reference's `TRG_DIR + "/../../" + filepath` resolves an absolute path.
int orc_save_graph(void* h, const char* path) {
  H(h)->saveGraph(std::string(path));
  return 0;
}
int orc_load_graph(void* h, const char* path) {
  H(h)->loadPrebuiltGraph(std::string(path));
  return 0;
}

// 1 = this library is the reference's own trg.cpp (oracle/_ref/libtrg_ref.so)
int orc_is_reference_build(void) { return 1; }

}  // extern "C"
#pragma GCC visibility pop
