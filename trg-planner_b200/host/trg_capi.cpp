// C facade of the host TRG class (include/trg_b200.h). Exceptions stop here.
#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <vector>
#include <cstring>
#include <limits>
#include <string>

#include "map_io.h"
#include "trg.h"
#include "trg_b200.h"
#include "trgb_kernels.h"

namespace {
thread_local std::string g_err;
inline TRG* T(void* h) { return static_cast<TRG*>(h); }

// subclass only to reach the protected param_ (front ends of the reference do the same)
struct TRGX : TRG {
  using TRG::TRG;
  float robotSize() const { return param_.robot_size; }
  float heightThr() const { return param_.height_threshold; }
  float collThr() const { return param_.collision_threshold; }
};
inline TRGX* X(void* h) { return static_cast<TRGX*>(h); }

// lockGraph() / unlockGraph() as a scope: an exception between the two must not leave TRG::mtx.graph locked
struct GraphLock {
  explicit GraphLock(TRG* t) : t_(t) { t_->lockGraph(); }
  ~GraphLock() { t_->unlockGraph(); }
  TRG* t_;
};
inline void check_type(const char* type) {
  if (!type) throw std::runtime_error("trg_b200: graph type is null");
  const std::string s(type);
  if (s != "global" && s != "local" && s != "prebuilt") throw std::runtime_error("trg_b200: unknown graph type '" + s + "'");
}

template <class F>
int guard(F&& f) {
  try {
    return f();
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  } catch (...) {
    g_err = "unknown exception";
    return -1;
  }
}
}  // namespace

extern "C" {

const char* trg_last_error(void) { return g_err.c_str(); }

void* trg_create(const TrgParams* p) {
  try {
    if (!p) { g_err = "null params"; return nullptr; }
    return new TRGX(p->is_verbose != 0, p->expand_dist, p->robot_size, p->sample_num, p->height_threshold,
                    p->collision_threshold, p->update_collision_threshold, p->safety_factor, p->goal_tolerance);
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void trg_destroy(void* h) { delete X(h); }
void trg_seed(void* h, uint32_t seed) { T(h)->reseed(seed); }

int trg_set_global_map(void* h, const float* xyz, int64_t n) {
  return guard([&] { T(h)->setGlobalMapRaw(xyz, n, 3, false); return 0; });
}
int trg_set_global_map_dev(void* h, const float* dev_xyz, int64_t n, int stride) {
  return guard([&] { T(h)->setGlobalMapRaw(dev_xyz, n, stride, true); return 0; });
}
int trg_set_local_map(void* h, float sx, float sy, const float* xyz, int64_t n) {
  return guard([&] { T(h)->setLocalMapRaw(Eigen::Vector2f(sx, sy), xyz, n, 3); return 0; });
}
int trg_init_graph(void* h, int is_pre_map, float sx, float sy, float sz) {
  return guard([&] { T(h)->initGraph(is_pre_map != 0, Eigen::Vector3f(sx, sy, sz)); return 0; });
}
int trg_update_graph(void* h) {
  return guard([&] { T(h)->updateGraph(); return 0; });
}

int trg_graph_counts(void* h, const char* type, int64_t* n_nodes, int64_t* n_edges) {
  return guard([&] {
    check_type(type);
    if (!n_nodes || !n_edges) throw std::runtime_error("trg_b200: null output pointer");
    GraphLock lock(T(h));
    const auto& g = T(h)->getGraphRef(type);
    int64_t e = 0;
    for (auto& kv : g) e += (int64_t)kv.second->edges_.size();
    *n_nodes = (int64_t)g.size();
    *n_edges = e;
    return 0;
  });
}

int trg_graph_export(void* h, const char* type, int32_t* iter_ids, int32_t* ids_sorted, float* pos_xyz,
                     int32_t* state, int64_t* row_ptr, int32_t* col, float* weight, float* dist) {
  return guard([&] {
    check_type(type);
    GraphLock lock(T(h));
    if (std::string(type) == "global" && T(h)->exportBuiltGraph(iter_ids, ids_sorted, pos_xyz, state, row_ptr, col, weight, dist))
      return 0;  // straight from the pools of the last device build
    const auto& g = T(h)->getGraphRef(type);
    std::vector<std::pair<int, TRG::Node*>> ids;
    ids.reserve(g.size());
    int64_t k = 0;
    bool dense = true;  // keys are exactly 0 .. n-1 (every cleaned graph): rows are placed by key instead of sorted
    const int64_t n = (int64_t)g.size();
    for (auto& kv : g) {
      if (iter_ids) iter_ids[k] = kv.first;
      ids.emplace_back(kv.first, kv.second);
      dense = dense && kv.first >= 0 && kv.first < n;
      ++k;
    }
    if (dense) {
      std::vector<std::pair<int, TRG::Node*>> by_key((size_t)n, {-1, nullptr});
      for (auto& p : ids) {
        if (by_key[(size_t)p.first].second) { dense = false; break; }
        by_key[(size_t)p.first] = p;
      }
      if (dense) ids.swap(by_key);
    }
    if (!dense) std::sort(ids.begin(), ids.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    const bool want_edges = row_ptr || col || weight || dist;
    // rows are independent once their first edge slot is known: offsets first, then helper threads fill
    std::vector<int64_t> first(ids.size() + 1, 0);
    if (want_edges)
      for (size_t i = 0; i < ids.size(); ++i) first[i + 1] = first[i] + (int64_t)ids[i].second->edges_.size();
    auto fill = [&](size_t b0, size_t e0) {
      for (size_t i = b0; i < e0; ++i) {
        TRG::Node* nd = ids[i].second;
        if (ids_sorted) ids_sorted[i] = ids[i].first;
        if (pos_xyz) { pos_xyz[3 * i] = nd->pos_.x(); pos_xyz[3 * i + 1] = nd->pos_.y(); pos_xyz[3 * i + 2] = nd->pos_.z(); }
        if (state) state[i] = (int32_t)nd->state_;
        if (row_ptr) row_ptr[i] = first[i];
        if (!want_edges) continue;
        int64_t e = first[i];
        for (auto* ed : nd->edges_) {
          if (col) col[e] = ed->dst_id_;
          if (weight) weight[e] = ed->weight_;
          if (dist) dist[e] = ed->dist_;
          ++e;
        }
      }
    };
    const int threads = ids.size() >= 65536 ? trg_b200::thread_budget() : 1;
    if (threads <= 1) {
      fill(0, ids.size());
    } else {
      std::vector<std::future<void>> jobs;
      for (int t = 0; t < threads; ++t)
        jobs.push_back(std::async(std::launch::async, fill, ids.size() * t / threads, ids.size() * (t + 1) / threads));
      for (auto& j : jobs) j.get();
    }
    const int64_t e = first[ids.size()];
    if (row_ptr) row_ptr[ids.size()] = e;
    return 0;
  });
}

int trg_save_graph(void* h, const char* path) {
  return guard([&] { T(h)->saveGraph(path); return 0; });
}
int trg_load_graph(void* h, const char* path) {
  return guard([&] { T(h)->loadPrebuiltGraph(path); return 0; });
}

int trg_plan(void* h, float sx, float sy, float gx, float gy, float gz, float* path_xyz, int32_t* node_ids,
             int max_pts, int* n_pts, float* direct_dist, float* path_length, float* avg_risk, int* goal_known,
             int64_t* n_expanded) {
  return guard([&] {
    std::vector<Eigen::Vector3f> path;
    Eigen::Vector2f s(sx, sy);
    Eigen::Vector3f g(gx, gy, gz);
    float dd = 0, pl = 0, ar = 0;
    bool ok = T(h)->planSafePath(s, g, path, dd, pl, ar);
    if (direct_dist) *direct_dist = dd;
    if (path_length) *path_length = pl;
    if (avg_risk) *avg_risk = ar;
    if (goal_known) *goal_known = T(h)->goalKnown() ? 1 : 0;
    if (n_expanded) *n_expanded = 0;
    const auto& ids = T(h)->lastPathIds();
    int n = (int)path.size();
    if (n_pts) *n_pts = n;
    for (int i = 0; i < n && i < max_pts; ++i) {
      if (path_xyz) { path_xyz[3 * i] = path[i].x(); path_xyz[3 * i + 1] = path[i].y(); path_xyz[3 * i + 2] = path[i].z(); }
      if (node_ids) node_ids[i] = ids[i];
    }
    return ok ? 1 : 0;
  });
}

int trg_plan_batch(void* h, const float* queries, int64_t n, uint8_t* found, float* cost, float* path_length,
                   float* avg_risk, float* direct_dist, uint8_t* goal_known, int64_t* path_offsets,
                   int32_t* path_ids, int64_t cap) {
  return guard([&] {
    TRG::PathBatch b;
    T(h)->planSafePathBatch(queries, n, b);
    for (int64_t i = 0; i < n; ++i) {
      if (found) found[i] = b.found[i];
      if (cost) cost[i] = b.cost[i];
      if (path_length) path_length[i] = b.path_length[i];
      if (avg_risk) avg_risk[i] = b.avg_risk[i];
      if (direct_dist) direct_dist[i] = b.direct_dist[i];
      if (goal_known) goal_known[i] = b.goal_known[i];
    }
    if (path_offsets) std::memcpy(path_offsets, b.offsets.data(), (size_t)(n + 1) * sizeof(int64_t));
    if ((int64_t)b.node_ids.size() > cap) { g_err = "path_ids buffer too small"; return -3; }
    if (path_ids && !b.node_ids.empty()) std::memcpy(path_ids, b.node_ids.data(), b.node_ids.size() * sizeof(int32_t));
    return 0;
  });
}

int trg_refine_path(void* h, const float* in_xyz, int n_in, float* out_xyz, int* n_out) {
  return guard([&] {
    std::vector<Eigen::Vector3f> in(n_in), out;
    for (int i = 0; i < n_in; ++i) in[i] = Eigen::Vector3f(in_xyz[3 * i], in_xyz[3 * i + 1], in_xyz[3 * i + 2]);
    T(h)->refinePath(in, out);
    *n_out = (int)out.size();
    for (size_t i = 0; i < out.size(); ++i) { out_xyz[3 * i] = out[i].x(); out_xyz[3 * i + 1] = out[i].y(); out_xyz[3 * i + 2] = out[i].z(); }
    return 0;
  });
}

int trg_check_reached(void* h, float x, float y) {
  return guard([&] { Eigen::Vector2f p(x, y); return T(h)->checkReadched(p) ? 1 : 0; });
}
int trg_check_replan(void* h, float x, float y, const float* path_xyz, int n_path) {
  return guard([&] {
    Eigen::Vector2f p(x, y);
    std::vector<Eigen::Vector3f> path(n_path);
    for (int i = 0; i < n_path; ++i) path[i] = Eigen::Vector3f(path_xyz[3 * i], path_xyz[3 * i + 1], path_xyz[3 * i + 2]);
    return T(h)->checkReplan(p, path) ? 1 : 0;
  });
}

int trg_load_params_yaml(const char* config_path, TrgParams* out, int* is_prebuilt_map, char* prebuilt_map_path,
                         int path_cap, int* is_voxelize, float* voxel_size, int* is_update) {
  return guard([&] {
    const trg_b200::PlannerParams p = trg_b200::load_params_yaml(config_path);
    if (out) {
      // the nine arguments TRGPlanner::init passes to TRG::TRG (trg_planner.cpp:23-31)
      *out = TrgParams{p.isVerbose ? 1 : 0, p.expandDist, p.robotSize, p.sampleNum, p.heightThreshold,
                       p.collisionThreshold, p.updateCollisionThreshold, p.safetyFactor, p.goal_tolerance};
    }
    if (is_prebuilt_map) *is_prebuilt_map = p.isPreMap ? 1 : 0;
    if (prebuilt_map_path && path_cap > 0) {
      std::strncpy(prebuilt_map_path, p.preMapPath.c_str(), (size_t)path_cap - 1);
      prebuilt_map_path[path_cap - 1] = 0;
    }
    if (is_voxelize) *is_voxelize = p.isVoxelize ? 1 : 0;
    if (voxel_size) *voxel_size = p.VoxelSize;
    if (is_update) *is_update = p.isUpdate ? 1 : 0;
    return 0;
  });
}

int trg_load_pcd(const char* path, float* xyz, int64_t cap_points, int64_t* n) {
  return guard([&] {
    const std::vector<float> v = trg_b200::load_pcd_xyz(path);
    const int64_t np = (int64_t)v.size() / 3;
    if (n) *n = np;
    if (xyz) {
      if (np > cap_points) throw std::runtime_error("trg_b200: PCD buffer too small");
      std::memcpy(xyz, v.data(), v.size() * sizeof(float));
    }
    return 0;
  });
}

int trg_save_pcd(const char* path, const float* xyz, int64_t n, int binary) {
  return guard([&] { trg_b200::save_pcd_xyz(path, xyz, n, binary != 0); return 0; });
}

int trg_load_prebuilt_map(void* h, const char* pcd_path, int is_voxelize, float voxel_size, int64_t* n_raw, int64_t* n_map) {
  return guard([&] {
    std::vector<float> raw = trg_b200::load_pcd_xyz(pcd_path);
    int64_t n = (int64_t)raw.size() / 3;
    if (n_raw) *n_raw = n;
    if (n == 0) throw std::runtime_error("trg_b200: empty prebuilt map");
    if (is_voxelize) {
      std::vector<float> out(raw.size());
      int64_t nv = 0;
      const int rc = trgb_voxel_filter(raw.data(), n, 3, voxel_size, out.data(), &nv);
      if (rc != TRGB_OK && rc != TRGB_E_STATE) throw std::runtime_error(trgb_last_error());
      out.resize((size_t)nv * 3);
      raw.swap(out);
      n = nv;
    }
    if (n_map) *n_map = n;
    T(h)->setGlobalMapRaw(raw.data(), n, 3, false);
    return 0;
  });
}

int trg_is_collision_batch(void* h, const char* type, const float* xy, int64_t n, float threshold, uint8_t* out) {
  return guard([&] { T(h)->isCollisionBatch(xy, n, type, threshold, out); return 0; });
}

int trg_range_count_batch(void* h, const char* type, const float* xy, int64_t n, float radius, int32_t* out) {
  return guard([&] {
    trgb_map* m = T(h)->deviceMap(type);
    if (!m) throw std::runtime_error("trg_b200: no map loaded");
    if (trgb_range_count_batch(m, xy, n, radius, out) != TRGB_OK) throw std::runtime_error(trgb_last_error());
    return 0;
  });
}

int trg_nearest_z_batch(void* h, const char* type, const float* xy, int64_t n, float* z, int64_t* idx, uint8_t* tie) {
  return guard([&] {
    trgb_map* m = T(h)->deviceMap(type);
    if (!m) throw std::runtime_error("trg_b200: no map loaded");
    if (trgb_nearest_z_batch(m, xy, n, z, idx, tie) != TRGB_OK) throw std::runtime_error(trgb_last_error());
    return 0;
  });
}

int trg_edge_eval_batch(void* h, const char* type, const float* p1, const float* p2, int64_t n, uint8_t* stage,
                        float* weight, double* weight64, float* dist, int32_t* npts) {
  return guard([&] {
    trgb_map* m = T(h)->deviceMap(type);
    if (!m) throw std::runtime_error("trg_b200: no map loaded");
    TrgbEdgeParams prm{X(h)->robotSize(), X(h)->heightThr(), X(h)->collThr(), 0};
    if (trgb_edge_eval_batch(m, p1, p2, n, &prm, stage, weight, dist, npts) != TRGB_OK)
      throw std::runtime_error(trgb_last_error());
    if (weight64) for (int64_t i = 0; i < n; ++i) weight64[i] = std::numeric_limits<double>::quiet_NaN();
    return 0;
  });
}

int trg_is_frontier_batch(void* h, const float* xy, int64_t n, uint8_t* out) {
  return guard([&] {
    for (int64_t i = 0; i < n; ++i) {
      Eigen::Vector2f p(xy[2 * i], xy[2 * i + 1]);
      out[i] = T(h)->isFrontier(p) ? 1 : 0;
    }
    return 0;
  });
}

void* trg_device_map(void* h, const char* type) {
  try {
    return T(h)->deviceMap(type);
  } catch (...) {
    return nullptr;
  }
}

double trg_last_seconds(void* h, const char* what) { return T(h)->lastSeconds(what); }
int64_t trg_stat(void* h, const char* what) { return T(h)->stat(what); }
int trg_set_tuning(void* h, const char* key, double value) {
  return guard([&] {
    std::string k(key);
    if (k == "chunk_nodes") T(h)->tuning_.chunk_nodes = (int)value;
    else if (k == "window") T(h)->tuning_.window = (int)value;
    else if (k == "lookahead") T(h)->tuning_.lookahead = (int)value;
    else if (k == "map_cell_scale") T(h)->tuning_.map_cell_scale = (float)value;
    else if (k == "table_cell_scale") T(h)->tuning_.table_cell_scale = (float)value;
    else if (k == "parallel_min_nodes") T(h)->tuning_.parallel_min_nodes = (int)value;
    else if (k == "overlap") T(h)->tuning_.overlap = value != 0;
    else if (k == "split_commit") T(h)->tuning_.split_commit = value != 0;
    else if (k == "device_expand") T(h)->tuning_.device_expand = value != 0;
    else if (k == "expand_steps") T(h)->tuning_.expand_steps = std::max(1, (int)value);
    else if (k == "expand_window_words") T(h)->tuning_.expand_window_words = std::min(4, std::max(2, (int)value));
    else if (k == "expand_max_pops") T(h)->tuning_.expand_max_pops = std::min(8192, std::max(32, (int)value));
    else throw std::runtime_error("unknown tuning key " + k);
    return 0;
  });
}

}  // extern "C"
