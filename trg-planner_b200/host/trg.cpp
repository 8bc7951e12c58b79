// TRG — B200-native implementation of the reference graph core
//   cpp/trg_planner/core/trg_planner/src/graph/trg.cpp   (cited below as trg.cpp:LINE).
//
// Same public behaviour, different machinery: every map query (the reference's kd_nearest_range2 /
// kd_nearest2 on `map_tree`) is a batched kernel launch through include/trgb_kernels.h; graph
// expansion is a wavefront scheduler (class Expander) that evaluates the pure-function work of
// many queue pops at once and then commits decisions in the reference's exact sequential order.
// Nothing here touches oracle/; without a CUDA device every map-dependent call throws.
#include "trg.h"

#include <math.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <limits>
#include <atomic>
#include <exception>
#include <functional>
#include <future>
#include <stdexcept>
#include <thread>

#include "device_session.h"
#include "trgb_kernels.h"

#ifdef TRG_FINE_TIMERS
#include <x86intrin.h>
#define FT_DECL(name) uint64_t ft_##name = 0
#define FT_BEGIN() uint64_t _ft0 = __rdtsc()
#define FT_LAP(name) do { const uint64_t _t = __rdtsc(); ft_##name += _t - _ft0; _ft0 = _t; } while (0)
#else
#define FT_DECL(name)
#define FT_BEGIN()
#define FT_LAP(name)
#endif

namespace {
using Clock = std::chrono::steady_clock;
inline double since(Clock::time_point t0) { return std::chrono::duration<double>(Clock::now() - t0).count(); }

[[noreturn]] void fail(const std::string& what) {
  throw std::runtime_error("trg_b200: " + what + ": " + trgb_last_error());
}
inline void K(int rc, const char* what) {
  if (rc != TRGB_OK) fail(what);
}
inline float norm2(float dx, float dy) { return sqrtf(dx * dx + dy * dy); }  // Vector2f::norm()
}  // namespace

// ================================================================================================
// construction / reset
// ================================================================================================
TRG::TRG(bool isVerbose, float expand_dist, float robot_size, int sample_num, float height_threshold,
         float collision_threshold, float update_collision_threshold, float safety_factor,
         float goal_tolerance)
    : gen_(rd_()), distr_(0.0, 1.0), dev_(new trg_b200::DeviceSession()) {  // trg.cpp:11-34
  param_.isVerbose                  = isVerbose;
  param_.expand_dist                = expand_dist;
  param_.robot_size                 = robot_size;
  param_.sample_num                 = sample_num;
  param_.height_threshold           = height_threshold;
  param_.collision_threshold        = collision_threshold;
  param_.update_collision_threshold = update_collision_threshold;
  param_.safety_factor              = safety_factor;
  param_.goal_tolerance             = goal_tolerance;
  this->resetGraph("global");
  this->resetGraph("local");
  this->resetMap("global");
  this->resetMap("local");
}

TRG::~TRG() {
  if (graph_disposal_.valid()) graph_disposal_.get();
  joinDrawPrefetch();
  destroyExpander();
  for (auto& kv : trgMap_) {
    if (kv.second->map_index) trgb_map_destroy(kv.second->map_index);
    kv.second->map_index = nullptr;
  }
  if (dev_graph_) trgb_graph_destroy(dev_graph_);
}

void TRG::resetGraph(std::string type) {  // trg.cpp:732-737
  trgStruct& graph = *trgMap_.at(type);
  graph.nodes.clear();
  nodeIndexReset(graph);
  graph.node_id = 0;
  if (type == "global") invalidateDeviceGraph();
}

void TRG::resetMap(std::string type) {  // trg.cpp:739-744
  trgStruct& graph = *trgMap_.at(type);
  if (graph.map_index && graph.map_index == expander_map_) expander_map_ = nullptr;
  if (graph.map_index) trgb_map_destroy(graph.map_index);
  graph.map_index  = nullptr;
  graph.map_points = 0;
  graph.cloud_map.reset(new pcl::PointCloud<PtsDefault>());
}

void TRG::reseed(uint32_t seed) {
  joinDrawPrefetch();  // a block generated ahead belongs to the old stream
  gen_.seed(seed);
  distr_.reset();
  draw_u_.clear();
  draw_xy_.clear();
  draw_base_ = draw_next_ = 0;
  dev_->draws.set_base(0);
}

// Nodes and edges live in pools (the reference leaks every `new Node` / `new Edge`). A full rebuild
// (initGraph / loadPrebuiltGraph) rewinds the pools and recycles the objects, adjacency vectors
// included, so repeated builds neither grow memory nor re-allocate 0.5 M small vectors.
TRG::Node* TRG::newNode(int id, Eigen::Vector2f& p, float z, NodeState s) {
  if (node_used_ < node_pool_.size()) {
    Node& n = node_pool_[node_used_++];
    n.id_ = id;
    n.pos_ = Eigen::Vector3f(p.x(), p.y(), z);
    n.state_ = s;
    n.edges_.clear();
    return &n;
  }
  node_pool_.emplace_back(id, p, z, s);
  node_pool_.back().edges_.reserve(8);  // typical degree 7: one allocation instead of four
  ++node_used_;
  return &node_pool_.back();
}
TRG::Edge* TRG::newEdge(int dst, float w, float d) {
  if (edge_used_ < edge_pool_.size()) {
    Edge& e = edge_pool_[edge_used_++];
    e.dst_id_ = dst; e.weight_ = w; e.dist_ = d;
    return &e;
  }
  edge_pool_.emplace_back(dst, w, d);
  ++edge_used_;
  return &edge_pool_.back();
}
void TRG::rewindPools() {
  node_used_ = 0;
  edge_used_ = 0;
  goal_.node = nullptr;  // pointed into the recycled pool
  last_path_ids_.clear();
}

// ================================================================================================
// sampling stream
// ================================================================================================
// One block of the sampling stream: u_k = distr_(gen_), then the offsets of trg.cpp:395-397
//   float angle = distr_(gen_) * 2 * M_PI;  Vector2f(expand_dist * cos(angle), expand_dist * sin(angle))
// (float overloads, glibc). Runs on a helper thread one block ahead of consumption: the stream is
// consumed strictly in order, so generating it early changes nothing.
void TRG::generateDrawBlock(size_t n, std::vector<float>& u, std::vector<float>& xy) {
  u.resize(n);
  xy.resize(2 * n);
  const float e = param_.expand_dist;
  for (size_t k = 0; k < n; ++k) u[k] = distr_(gen_);  // the generator is sequential ...
  auto trig = [&u, &xy, e](size_t b, size_t en) {      // ... glibc cosf / sinf are not: slices on helper threads
    for (size_t k = b; k < en; ++k) {
      const float angle = u[k] * 2 * M_PI;
      xy[2 * k]     = e * cosf(angle);
      xy[2 * k + 1] = e * sinf(angle);
    }
  };
  const int threads = n >= ((size_t)1 << 16) ? std::min(6, trg_b200::thread_budget()) : 1;
  if (threads <= 1) {
    trig(0, n);
    return;
  }
  std::vector<std::future<void>> jobs;
  for (int t = 1; t < threads; ++t) jobs.push_back(std::async(std::launch::async, trig, n * t / threads, n * (t + 1) / threads));
  trig(0, n / threads);
  for (auto& j : jobs) j.get();
}

void TRG::joinDrawPrefetch() {
  if (draw_prefetch_.valid()) draw_prefetch_.get();
}

void TRG::ensureDraws(size_t upto) {
  auto t0 = Clock::now();
  while (upto > draw_base_ + draw_u_.size()) {
    const size_t have = draw_base_ + draw_u_.size();
    const size_t block = have < ((size_t)1 << 15) ? ((size_t)1 << 12) : ((size_t)1 << 18);
    if (draw_prefetch_.valid()) {
      draw_prefetch_.get();  // block generated ahead by the helper
    } else {
      generateDrawBlock(block, pre_u_, pre_xy_);
    }
    draw_u_.insert(draw_u_.end(), pre_u_.begin(), pre_u_.end());
    draw_xy_.insert(draw_xy_.end(), pre_xy_.begin(), pre_xy_.end());
    // start the next block right away once the stream is clearly in heavy use
    if (draw_u_.size() >= ((size_t)1 << 15))
      draw_prefetch_ = std::async(std::launch::async, [this] { generateDrawBlock((size_t)1 << 18, pre_u_, pre_xy_); });
  }
  us_draws_ += 1e6 * since(t0);  // (plain member: runs on the helper thread)
}

// make the device copy cover exactly the host buffer [draw_base_, draw_base_ + size)
void TRG::syncDraws() {
  trg_b200::DrawBuffer& db = dev_->draws;
  const size_t host_end = draw_base_ + draw_u_.size();
  if (db.base() != draw_base_ || db.end() > host_end) db.set_base(draw_base_);
  if (db.end() < host_end)
    db.append(draw_xy_.data() + 2 * (db.end() - draw_base_), host_end - db.end(), dev_->copyStream());
}

float TRG::nextUniform() {
  ensureDraws(draw_next_ + 1);
  return draw_u_[draw_next_++ - draw_base_];
}

void TRG::compactDraws() {
  const size_t used = draw_next_ - draw_base_;
  if (used < ((size_t)1 << 20)) return;
  draw_u_.erase(draw_u_.begin(), draw_u_.begin() + used);
  draw_xy_.erase(draw_xy_.begin(), draw_xy_.begin() + 2 * used);
  draw_base_ = draw_next_;
  dev_->draws.set_base(draw_base_);  // dropped: syncDraws() re-uploads the remainder
}

// ================================================================================================
// maps
// ================================================================================================
void TRG::buildMapIndex(trgStruct& g, const float* xyz, int64_t n, int stride, bool device) {
  if (g.map_index && g.map_index == expander_map_) expander_map_ = nullptr;  // the engine is re-bound at the next build
  if (g.map_index) trgb_map_destroy(g.map_index);
  g.map_index  = nullptr;
  g.map_points = 0;
  if (n <= 0) return;
  const float cell = tuning_.map_cell_scale * param_.robot_size;
  if (device) K(trgb_map_create_dev(&g.map_index, xyz, n, stride, cell), "trgb_map_create_dev");
  else K(trgb_map_create(&g.map_index, xyz, n, stride, cell), "trgb_map_create");
  g.map_points = n;
  TrgbMapInfo mi;
  K(trgb_map_info(g.map_index, &mi), "trgb_map_info");
  g.bbox[0] = mi.origin_x;
  g.bbox[1] = mi.origin_y;
  g.bbox[2] = mi.origin_x + mi.grid_w * mi.cell_size;
  g.bbox[3] = mi.origin_y + mi.grid_h * mi.cell_size;
}

trgb_map* TRG::requireMap(trgStruct& g, const char* who) {
  if (!g.map_index) throw std::runtime_error(std::string("trg_b200: ") + who + ": no " + g.type + " map loaded");
  return g.map_index;
}

void TRG::setGlobalMap(PointCloudPtr& map) {  // trg.cpp:179-193
  auto t0 = Clock::now();
  trgStruct& g = *trgMap_["global"];
  this->resetMap("global");
  *g.cloud_map = *map;
  const int64_t n = (int64_t)g.cloud_map->size();
  if (n > 0) buildMapIndex(g, reinterpret_cast<const float*>(g.cloud_map->points.data()), n, 4, false);
  secs_["set_global_map"] = since(t0);
}

void TRG::setGlobalMapRaw(const float* xyz, int64_t n, int stride_floats, bool device) {
  auto t0 = Clock::now();
  trgStruct& g = *trgMap_["global"];
  this->resetMap("global");
  buildMapIndex(g, xyz, n, stride_floats, device);
  secs_["set_global_map"] = since(t0);
}

void TRG::setLocalMap(Eigen::Vector2f start2d, PointCloudPtr& map) {  // trg.cpp:195-209
  std::lock_guard<std::mutex> lock(mtx.graph);
  auto t0 = Clock::now();
  trgStruct& l = *trgMap_["local"];
  this->resetMap("local");
  l.root_pos   = start2d;
  *l.cloud_map = *map;
  const int64_t n = (int64_t)l.cloud_map->size();
  if (n > 0) buildMapIndex(l, reinterpret_cast<const float*>(l.cloud_map->points.data()), n, 4, false);
  this->setLocalGraph(false);
  secs_["set_local_map"] = since(t0);
}

void TRG::setLocalMapRaw(Eigen::Vector2f start2d, const float* xyz, int64_t n, int stride_floats) {
  std::lock_guard<std::mutex> lock(mtx.graph);
  auto t0 = Clock::now();
  trgStruct& l = *trgMap_["local"];
  this->resetMap("local");
  l.root_pos = start2d;
  buildMapIndex(l, xyz, n, stride_floats, false);
  this->setLocalGraph(false);
  secs_["set_local_map"] = since(t0);
}

void TRG::setLocalGraph(bool useMutex) {  // trg.cpp:211-231
  if (useMutex) {
    std::lock_guard<std::mutex> lock(mtx.graph);  // (sic) the reference's guard dies here too
  }
  trgStruct& g = *trgMap_["global"];
  trgStruct& l = *trgMap_["local"];
  this->resetGraph("local");
  if (g.nodes.empty()) return;
  // The reference range-counts the local map around EVERY global node, in the map's iteration order (:215-229).
  // Only nodes within robot_size / 2 of the local cloud's bounding box can count a point, and right after
  // cleanGraph the iteration order is node_seq's: candidates come from one pass over the flat position array
  // instead of a walk over a million hash nodes, and only they go to the device.
  const float reach = (float)(param_.robot_size * 0.5) * 1.01f + 1e-3f;
  const bool boxed = l.map_index && l.bbox[2] >= l.bbox[0] && l.bbox[3] >= l.bbox[1];
  const float bx0 = l.bbox[0] - reach, by0 = l.bbox[1] - reach, bx1 = l.bbox[2] + reach, by1 = l.bbox[3] + reach;
  std::vector<Node*> order;
  std::vector<float> xy;
  if (g.seq_in_iter_order && g.node_seq.size() == g.nodes.size()) {
    const size_t n = g.node_seq.size();
    for (size_t i = 0; i < n; ++i) {
      const float x = g.seq_xy[2 * i], y = g.seq_xy[2 * i + 1];
      if (boxed && !(x >= bx0 && x <= bx1 && y >= by0 && y <= by1)) continue;
      order.push_back(g.node_seq[i]);
      xy.push_back(x);
      xy.push_back(y);
    }
  } else {
    for (auto& node : g.nodes) {
      const float x = node.second->pos_.x(), y = node.second->pos_.y();
      if (boxed && !(x >= bx0 && x <= bx1 && y >= by0 && y <= by1)) continue;
      order.push_back(node.second);
      xy.push_back(x);
      xy.push_back(y);
    }
  }
  std::vector<int32_t> cnt(order.size(), 0);
  if (l.map_index && !order.empty())
    K(trgb_range_count_batch(l.map_index, xy.data(), (int64_t)order.size(), (float)(param_.robot_size * 0.5), cnt.data()),
      "trgb_range_count_batch");
  for (size_t k = 0; k < order.size(); ++k) {
    if (cnt[k] == 0) continue;
    l.nodes[order[k]->id_] = order[k];   // (key == id_ for every node of the global map)
    nodeIndexInsert(l, order[k]);
  }
}

// ================================================================================================
// node index (replaces kdtree* node_tree)
// ================================================================================================
void TRG::nodeIndexReset(trgStruct& g) {
  g.node_seq.clear();
  g.seq_xy.clear();
  g.node_grid.clear();
  g.node_tree.clear();
  g.grid_built = 0;
  g.tree_built = 0;
  g.iter_rank.clear();
  g.seq_in_iter_order = true;  // (empty)
  if (dev_ && dev_->nodes_owner == &g) dev_->nodes_owner = nullptr;  // device copy is stale
}

void TRG::ensureGrid(trgStruct& g) {
  if (g.node_grid.configured()) return;
  // nodes live within robot_size of a map point; pad generously (positions are clamped anyway)
  const trgStruct& m = global_trg_.map_index ? global_trg_ : g;
  float x0 = m.bbox[0], y0 = m.bbox[1], x1 = m.bbox[2], y1 = m.bbox[3];
  if (!(x1 > x0) || !(y1 > y0)) { x0 = y0 = -64.f; x1 = y1 = 64.f; }
  const float cell = 1.5f * param_.robot_size;
  g.node_grid.configure(x0 - 2.f, y0 - 2.f, x1 + 2.f, y1 + 2.f, cell);
}

void TRG::nodeIndexInsert(trgStruct& g, Node* n) {
  g.node_seq.push_back(n);  // the host grid / order tree catch up lazily (ensureGridBuilt / ensureTree)
  g.seq_xy.push_back(n->pos_.x());
  g.seq_xy.push_back(n->pos_.y());
  g.seq_in_iter_order = false;  // a hash map does not iterate in insertion order; cleanGraph re-establishes it
  if (!g.iter_rank.empty()) g.iter_rank.clear();
}

void TRG::ensureGridBuilt(trgStruct& g) {
  ensureGrid(g);
  if (g.grid_built == 0 && g.node_seq.size() >= 65536) {  // a whole graph at once (after cleanGraph): banded, on helper threads
    g.node_grid.rebuild(g.seq_xy.data(), (int)g.node_seq.size(), trg_b200::thread_budget());
    g.grid_built = g.node_seq.size();
    return;
  }
  for (; g.grid_built < g.node_seq.size(); ++g.grid_built)
    g.node_grid.insert(g.seq_xy[2 * g.grid_built], g.seq_xy[2 * g.grid_built + 1]);
}

void TRG::ensureTree(trgStruct& g) {
  const size_t n = g.node_seq.size();
  const size_t pending = n - g.tree_built;
  // Catching up by single insertions costs one cache-missing descent per node; past a few tens of
  // thousands of pending nodes (exact-distance ties during a big build are ~1e-7 per query, so the
  // tree is usually far behind when it is needed) a parallel bulk rebuild of the identical tree wins.
  if (n >= 256 && (g.tree_built == 0 || pending > 20000)) {
    auto ta = Clock::now();
    if (n >= 20000 && trgb_device_count() > 0) {
      // large graph: the same tree, grown on the device one level per round (a few ms instead of ~50)
      std::vector<int> lo(n), hi(n), par(n);
      std::vector<uint8_t> ax(n);
      auto tb = Clock::now();
      K(trgb_kdtree_build(g.seq_xy.data(), (int64_t)n, lo.data(), hi.data(), par.data(), ax.data()), "trgb_kdtree_build");
      auto tc = Clock::now();
      g.node_tree.adopt(g.seq_xy.data(), (int)n, std::move(lo), std::move(hi), std::move(par), std::move(ax));
      us_tree_parts_[0] += (int64_t)(1e6 * std::chrono::duration<double>(tb - ta).count());
      us_tree_parts_[1] += (int64_t)(1e6 * std::chrono::duration<double>(tc - tb).count());
      us_tree_parts_[2] += (int64_t)(1e6 * since(tc));
    } else {
      std::vector<float> xs(n), ys(n);
      for (size_t i = 0; i < n; ++i) { xs[i] = g.seq_xy[2 * i]; ys[i] = g.seq_xy[2 * i + 1]; }
      g.node_tree.build_bulk(xs.data(), ys.data(), (int)n);
    }
    g.tree_built = n;
    return;
  }
  for (; g.tree_built < n; ++g.tree_built)
    g.node_tree.insert(g.seq_xy[2 * g.tree_built], g.seq_xy[2 * g.tree_built + 1], (int)g.tree_built);
}

// kd_nearest2 on node_tree (kdtree.c:364-417): exact float argmin; exact ties by tree visit order
TRG::Node* TRG::nearestNode(trgStruct& g, float x, float y) {
  if (g.node_seq.empty()) return nullptr;
  ensureGridBuilt(g);
  auto nn = g.node_grid.nearest(x, y);
  if (nn.entry >= 0 && !nn.tie) return g.node_seq[nn.entry];
  ++n_node_ties_;
  if (g.tree_built == g.node_seq.size()) return g.node_seq[g.node_tree.nearest(x, y)];  // tree is current: ask it
  return resolveNearestTie(g, x, y, nn.d2);
}

// Which of several nodes at the IDENTICAL float distance kd_nearest returns, without materialising
// the insertion-order tree (during a big build it is far behind and a tie shows up ~1e-7 per query).
// kd_nearest_i (kdtree.c:303-362) visits, at every tree node, the subtree on the query's side, then
// the node, then the other subtree, and replaces its result only on a strictly smaller distance —
// so the first tied node visited wins, and the root (the initial result, :393-395) always wins.
// The root path of a node is recovered by one pass over the earlier insertions: the first inserted
// point that falls into a subtree's region is that subtree's root.
TRG::Node* TRG::resolveNearestTie(trgStruct& g, float qx, float qy, float d2min) {
  std::vector<int> cand;
  g.node_grid.for_each_within_d2(qx, qy, d2min, sqrtf(d2min) * 1.001f + 1e-4f, [&](int e) { cand.push_back(e); });
  std::sort(cand.begin(), cand.end());
  if (cand.empty()) return g.node_seq[g.node_grid.nearest(qx, qy).entry];
  if (cand.size() == 1 || cand[0] == 0) return g.node_seq[cand[0]];
  const int last = cand.back();
  std::vector<float> xs((size_t)last + 1), ys((size_t)last + 1);
  for (int i = 0; i <= last; ++i) { xs[i] = g.node_seq[i]->pos_.x(); ys[i] = g.node_seq[i]->pos_.y(); }
  return g.node_seq[trg_b200::first_visited_of(xs.data(), ys.data(), cand, qx, qy)];
}

// kd_nearest_range2 on node_tree in the reference's result-iteration order: the in-range SET comes
// from the hash grid, the ORDER from the root paths of those few nodes in the insertion-order tree
// (OrderTree2D::order_like_range) instead of a traversal of the unbalanced tree around the query.
void TRG::rangeNodesOrdered(trgStruct& g, float x, float y, float r, std::vector<Node*>& out) {
  out.clear();
  if (g.node_seq.empty()) return;
  ensureTree(g);
  ensureGridBuilt(g);
  static thread_local std::vector<int> idx;
  idx.clear();
  g.node_grid.for_each_in_range(x, y, r, [&](int e) { idx.push_back(e); });
  g.node_tree.order_like_range(idx, x, y);
  for (int i : idx) out.push_back(g.node_seq[i]);
}

int TRG::countNodesInRange(trgStruct& g, float x, float y, float r) {
  if (g.node_seq.empty()) return 0;
  ensureGridBuilt(g);
  return g.node_grid.count_in_range(x, y, r);
}

// ================================================================================================
// single-shot primitives of the public API (each is one tiny batch on the device)
// ================================================================================================
bool TRG::isCollision(Eigen::Vector2f& pos, std::string type, float threshold) {  // trg.cpp:746-778
  trgStruct& g = *trgMap_.at(type);
  uint8_t out = 1;
  const float xy[2] = {pos.x(), pos.y()};
  K(trgb_collision_batch(requireMap(g, "isCollision"), xy, 1, param_.robot_size, param_.height_threshold, threshold, &out),
    "trgb_collision_batch");
  stat_["collision_calls"]++;
  return out != 0;
}

void TRG::isCollisionBatch(const float* xy, int64_t n, const std::string& type, float threshold, uint8_t* out) {
  trgStruct& g = *trgMap_.at(type);
  K(trgb_collision_batch(requireMap(g, "isCollisionBatch"), xy, n, param_.robot_size, param_.height_threshold, threshold, out),
    "trgb_collision_batch");
  stat_["collision_calls"] += n;
}

bool TRG::isFrontier(Eigen::Vector2f& pos) {  // trg.cpp:780-803
  trgStruct& g = *trgMap_["global"];
  trgStruct& l = *trgMap_["local"];
  Eigen::Vector2f dir = pos - l.root_pos;
  dir.normalize();
  Eigen::Vector2f check = pos + 2 * param_.robot_size * dir;
  if (countNodesInRange(g, check.x(), check.y(), param_.robot_size) > 0) return false;
  int32_t cnt = 0;
  const float xy[2] = {check.x(), check.y()};
  if (l.map_index)
    K(trgb_range_count_batch(l.map_index, xy, 1, (float)(0.5 * param_.robot_size), &cnt), "trgb_range_count_batch");
  return cnt == 0;
}

bool TRG::addNode(int node_id, Eigen::Vector2f& node_pos, NodeState state, std::string type) {  // trg.cpp:233-252
  trgStruct& graph = *trgMap_.at(type);
  if (node_id == 0) {
    if (this->isCollision(node_pos, graph.type, param_.collision_threshold)) return false;
  }
  float z = 0.f;
  uint8_t tie = 0;
  const float xy[2] = {node_pos.x(), node_pos.y()};
  K(trgb_nearest_z_batch(requireMap(graph, "addNode"), xy, 1, &z, nullptr, &tie), "trgb_nearest_z_batch");
  stat_["nearest_map"]++;
  if (tie) stat_["z_ties"]++;
  Node* node           = newNode(node_id, node_pos, z, state);
  graph.nodes[node_id] = node;
  nodeIndexInsert(graph, node);
  graph.node_id++;
  if (&graph == &global_trg_) invalidateDeviceGraph();
  return true;
}

namespace {
// trg.cpp:269-274 — slope gate, float overloads (SURVEY.md hard part 2):
//   atan2(fabs(dz), norm) > atan2(height_threshold, robot_size)
// atan2f is only evaluated when dz/norm is within 0.1 % of the threshold ratio; outside that band
// the comparison of the two angles is decided by the ratio alone (atan is strictly monotone and
// glibc's atan2f error is a few ulp, far below the band).
struct SlopeGate {
  float max_slope, lo, hi;
  SlopeGate(float height_thr, float robot_size) {
    max_slope = atan2f(height_thr, robot_size);
    const float t = height_thr / robot_size;
    lo = t * 0.999f;
    hi = t * 1.001f;
  }
  bool rejects(const Eigen::Vector3f& a, const Eigen::Vector3f& b) const {
    const float dz = fabsf(a.z() - b.z());
    const float d  = norm2(a.x() - b.x(), a.y() - b.y());
    if (d > 0.f && lo > 0.f) {
      if (dz < lo * d) return false;
      if (dz > hi * d) return true;
    }
    return atan2f(dz, d) > max_slope;
  }
};
inline bool slope_rejects(const Eigen::Vector3f& a, const Eigen::Vector3f& b, float height_thr, float robot_size) {
  const float max_slope = atan2f(height_thr, robot_size);
  const float slope     = atan2f(fabsf(a.z() - b.z()), norm2(a.x() - b.x(), a.y() - b.y()));
  return slope > max_slope;
}
inline bool has_edge_to(const TRG::Node* n, int id) {
  for (auto* e : n->edges_)
    if (e->dst_id_ == id) return true;
  return false;
}
}  // namespace

void TRG::wireEdge(Node* node1, Node* node2, std::string type) {  // trg.cpp:254-370
  if (node1->id_ == node2->id_) return;
  if (has_edge_to(node1, node2->id_) || has_edge_to(node2, node1->id_)) return;
  trgStruct& graph = *trgMap_.at(type);
  const float p1[3] = {node1->pos_.x(), node1->pos_.y(), node1->pos_.z()};
  const float p2[3] = {node2->pos_.x(), node2->pos_.y(), node2->pos_.z()};
  TrgbEdgeParams prm{param_.robot_size, param_.height_threshold, param_.collision_threshold, 0};
  uint8_t stage = 0;
  float w = 0.f, d = 0.f;
  K(trgb_edge_eval_batch(requireMap(graph, "wireEdge"), p1, p2, 1, &prm, &stage, &w, &d, nullptr), "trgb_edge_eval_batch");
  stat_["edge_evals"]++;
  if (stage != TRGB_EDGE_OK) return;
  node1->edges_.push_back(newEdge(node2->id_, w, d));
  node2->edges_.push_back(newEdge(node1->id_, w, d));
  if (&graph == &global_trg_) invalidateDeviceGraph();
}

// ================================================================================================
// Expander — wavefront scheduler for TRG::expandGraph (trg.cpp:372-454)
//
// The reference pops one node at a time: draw angles until sample_num collision-free samples, then
// for each sample look up the nearest existing node and wire / insert. Three facts make it
// batchable without changing a single decision:
//   (1) sampling never looks at the graph: the draw positions consumed by successive pops form a
//       chain o_{i+1} = o_i + consumed_i that depends only on (node position, map, stream);
//   (2) the z of a would-be new node and the geometric part of wireEdge(node, new) are pure
//       functions of (node, sample, map) -> evaluated speculatively for every accepted sample;
//   (3) wireEdge to an *existing* node only appends edges; nothing decided later in the same
//       expansion depends on its outcome except the emptiness of a brand-new node's edge list,
//       so those evaluations are deferred (placeholder edges keep the reference's edge order and
//       duplicate suppression) and resolved in the next batch.
// Per batch of pops: [windows kernel] -> host chain scan -> [nearest-z + edge kernels] -> host
// commit in (pop, sample) order.
// ================================================================================================
namespace trg_b200 {

// (ChunkTable - the host table of nodes created since the last hand-over - lives in node_index.h)

// Single-producer / single-consumer queue of edge-list operations. While the committing thread
// decides (nearest node, new node or not, queue push), a second thread applies what those decisions
// imply for the adjacency lists — duplicate checks, slope gate, placeholder edges — in the same
// order. Legal whenever no decision reads an edge list, i.e. without the step-3 neighbour wiring
// (there a new node is Invalid exactly when its parent edge failed, which the decider knows).
struct EdgeOp {
  TRG::Node* a;
  TRG::Node* b;
  float w, d;
  int kind;  // 0 = wireEdge(a, b) deferred, 1 = parent edge a <-> b with (w, d)
};

template <class Apply>
class EdgeWorker {
 public:
  explicit EdgeWorker(Apply apply) : apply_(std::move(apply)), ring_(kCap) {
    thread_ = std::thread([this] { loop(); });
  }
  ~EdgeWorker() {
    stop_.store(true, std::memory_order_release);
    if (thread_.joinable()) thread_.join();
  }
  void push(const EdgeOp& op) {
    const size_t t = tail_.load(std::memory_order_relaxed);
    while (t - head_.load(std::memory_order_acquire) >= kCap) pause();
    ring_[t & (kCap - 1)] = op;
    tail_.store(t + 1, std::memory_order_release);
  }
  // wait until every pushed operation has been applied; rethrows a failure of the worker
  void drain() {
    const size_t t = tail_.load(std::memory_order_relaxed);
    while (head_.load(std::memory_order_acquire) < t && !failed_.load(std::memory_order_acquire)) pause();
    if (failed_.load(std::memory_order_acquire)) std::rethrow_exception(error_);
  }

 private:
  static constexpr size_t kCap = (size_t)1 << 16;
  static void pause() {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  void loop() {
    try {
      size_t h = 0;
      while (true) {
        const size_t t = tail_.load(std::memory_order_acquire);
        if (h == t) {
          if (stop_.load(std::memory_order_acquire)) return;
          pause();
          continue;
        }
        for (; h < t; ++h) apply_(ring_[h & (kCap - 1)]);
        head_.store(h, std::memory_order_release);
      }
    } catch (...) {
      error_ = std::current_exception();
      failed_.store(true, std::memory_order_release);
    }
  }
  Apply apply_;
  std::vector<EdgeOp> ring_;
  std::atomic<size_t> head_{0}, tail_{0};
  std::atomic<bool> stop_{false}, failed_{false};
  std::exception_ptr error_;
  std::thread thread_;
};

class Expander {
 public:
  Expander(TRG& t, TRG::trgStruct& g)
      : t_(t), g_(g), P_(t.param_), gate_(t.param_.height_threshold, t.param_.robot_size) {
    map_ = t.requireMap(g, "expandGraph");
    st_  = (cudaStream_t)trgb_map_stream(map_);
    cuda_check(cudaGetDevice(&device_), "cudaGetDevice");
    // trg.cpp:429 — `float - float < double * float`
    step3_ = (P_.expand_dist - P_.robot_size < 0.25 * P_.expand_dist);
    mean_  = 1.15 * P_.sample_num;
    // device grid over the graph's nodes (K5), kept across expansions of the same graph
    DeviceSession& d = *t.dev_;
    const float cell = 1.5f * P_.robot_size;
    const float box[5] = {t.global_trg_.bbox[0], t.global_trg_.bbox[1], t.global_trg_.bbox[2], t.global_trg_.bbox[3], cell};
    if (!d.nodes || std::memcmp(box, d.nodes_box, sizeof(box)) != 0) {
      if (d.nodes) trgb_nodes_destroy(d.nodes);
      d.nodes = nullptr;
      K(trgb_nodes_create(&d.nodes, box[0] - 2.f, box[1] - 2.f, box[2] + 2.f, box[3] + 2.f, cell), "trgb_nodes_create");
      std::memcpy(d.nodes_box, box, sizeof(box));
      d.nodes_owner = nullptr;
    }
    if (d.nodes_owner != &g || d.nodes_uploaded > g.node_seq.size()) {
      K(trgb_nodes_reset(d.nodes, st_), "trgb_nodes_reset");
      d.nodes_owner = &g;
      d.nodes_uploaded = 0;
    }
    handed_nodes_ = d.nodes_uploaded;
    // most samples have a device candidate within robot_size, so ChunkTable::refine probes the 1 - 4
    // cells of this width that the candidate's disc touches (about one stored node each)
    table_.configure(box[0], box[1], t_.tuning_.table_cell_scale * P_.robot_size);
  }

  // expandGraph(root) for every root, in order (trg.cpp:372-454 / :483-487).
  //
  // Two host threads when possible: while this thread commits batch k, a helper runs the device
  // phases (sampling windows, chain scan, speculative evaluation) of batch k+1, whose pops are
  // already sitting in the BFS queue. That is legal because the sampling chain never looks at the
  // graph; what batch k creates in the meantime reaches the device one batch later and is covered
  // on the host by ChunkTable. Used for a single BFS without step-3 stalls; otherwise serial.
  void run(const std::vector<TRG::Node*>& roots) {
    t_.compactDraws();
    const size_t C = (size_t)std::max(1, t_.tuning_.chunk_nodes);
    const bool overlap = t_.tuning_.overlap && roots.size() == 1 && !step3_;
    auto apply = [this](const EdgeOp& op) { applyEdgeOp(op); };
    std::unique_ptr<EdgeWorker<decltype(apply)>> worker;
    if (t_.tuning_.split_commit && !step3_) {
      worker.reset(new EdgeWorker<decltype(apply)>(apply));
      push_op_ = [&worker](const EdgeOp& op) { worker->push(op); };
    } else {
      push_op_ = [this](const EdgeOp& op) { applyEdgeOp(op); };
    }
    size_t root_i = 0;
    int cur_ref = -1;
    int cur = 0;
    // first batch: nothing to overlap with
    if (!prepare(B_[cur], roots, root_i, cur_ref, C, t_.draw_next_)) return;
    feed(B_[cur]);
    while (true) {
      Batch& b = B_[cur];
      Batch& nb = B_[cur ^ 1];
      std::future<void> fut;
      bool launched = false;
      if (overlap && bfs_.size() - sent_ >= std::min<size_t>((size_t)std::max(1, t_.tuning_.lookahead), C)) {
        launched = prepare(nb, roots, root_i, cur_ref, C, b.chain_end);
        if (launched)
          fut = std::async(std::launch::async, [this, &nb] {
            cuda_check(cudaSetDevice(device_), "cudaSetDevice(helper)");  // the current device is per thread
            feed(nb);
          });
      }
      auto tc = Clock::now();
      try {
        commit(b, roots, root_i, cur_ref);
        if (worker) worker->drain();  // edge lists and pending_ are settled before the next hand-over
      } catch (...) {
        if (launched) fut.wait();
        throw;
      }
      us_commit_ += 1e6 * since(tc);
      if (launched) {
        auto tw = Clock::now();
        fut.get();
        us_wait_ += 1e6 * since(tw);
      } else {
        if (!prepare(nb, roots, root_i, cur_ref, C, t_.draw_next_)) break;
        feed(nb);
      }
      absorb(b);
      cur ^= 1;
    }
    absorb(B_[cur]);
    worker.reset();  // joins the edge thread
    flushDeferred();
    t_.stat_["pops"] += n_pops_;
    t_.stat_["nearest_node"] += n_nearest_;
    t_.stat_["z_ties"] += n_zties_;
    t_.stat_["stalls"] += n_stalls_;
    t_.stat_["us_commit"] += (int64_t)us_commit_;
    t_.stat_["us_wait"] += (int64_t)us_wait_;
#ifdef TRG_FINE_TIMERS
    t_.stat_["cyc_nearest"] += (int64_t)ft_nearest;
    t_.stat_["cyc_wire"] += (int64_t)ft_wire;
    t_.stat_["cyc_newnode"] += (int64_t)ft_newnode;
    t_.stat_["cyc_nn_a"] += (int64_t)ft_nn_a;
    t_.stat_["cyc_nn_b"] += (int64_t)ft_nn_b;
    t_.stat_["cyc_nn_c"] += (int64_t)ft_nn_c;
    t_.stat_["cyc_pre"] += (int64_t)ft_pre;
    t_.stat_["cyc_alloc"] += (int64_t)ft_alloc;
    t_.stat_["cyc_umap"] += (int64_t)ft_umap;
    t_.stat_["cyc_index"] += (int64_t)ft_index;
    t_.stat_["cyc_table"] += (int64_t)ft_table;
#endif
  }

 private:
  struct Pop {
    TRG::Node* node;
    float x, y, z;  // cached: the helper thread never dereferences graph objects
    int ref_id;
    bool is_root;
    size_t draw_start = 0;
    int consumed = 0;
    uint32_t acc_begin = 0, acc_count = 0;
  };
  struct Sample { float x, y; };
  struct Deferred {
    TRG::Node *a, *b;
    TRG::Edge *ea, *eb;
    float ax, ay, az, bx, by;
  };
  // one unit of device work; filled by prepare() (committing thread), feed() (either thread)
  struct Batch {
    std::vector<Pop> pops;
    std::vector<float> new_xy;       // nodes created since the previous hand-over -> device node grid
    std::vector<Deferred> deferred;  // edge evaluations deferred by earlier commits
    size_t chain_start = 0, chain_end = 0;
    size_t grid_count = 0;           // node_seq prefix the nearest-node candidates cover
    std::vector<Sample> acc;
    std::vector<float> z, w, d, nn_d2, d_w, d_d;
    std::vector<uint8_t> tie, stage, nn_tie, d_stage;
    std::vector<int32_t> nn_idx;
    // helper-side counters, merged by absorb()
    int64_t window_launches = 0, window_tests = 0, eval_launches = 0, edge_evals = 0, nearest_map = 0, batches = 0;
    double us_sample = 0, us_eval = 0;
  };

  // ---- batch formation (committing thread) --------------------------------------------------
  // The pop sequence is: the BFS queue of the current root, then — speculatively, and only when
  // the whole queue is inside this batch — the following roots (updateGraph).
  bool prepare(Batch& b, const std::vector<TRG::Node*>& roots, size_t root_i, int cur_ref, size_t C, size_t chain_start) {
    b.pops.clear();
    // leave about half of what is queued for the next batch, so that it can be fed while this one
    // commits (a BFS generation is only a few thousand pops); small remainders go out whole
    const size_t avail = bfs_.size() - sent_;
    const size_t look  = (size_t)std::max(1, t_.tuning_.lookahead);
    if (avail > look) C = std::min(C, std::max(look, avail / 2));
    for (size_t k = sent_; k < bfs_.size() && b.pops.size() < C; ++k) {
      TRG::Node* n = bfs_[k];
      b.pops.push_back({n, n->pos_.x(), n->pos_.y(), n->pos_.z(), cur_ref, false});
    }
    const size_t from_queue = b.pops.size();
    if (sent_ + from_queue == bfs_.size() && sent_ == head_) {
      // queue fully covered and nothing of it in flight: following roots may be speculated
      for (size_t r = root_i; r < roots.size() && b.pops.size() < C; ++r) {
        TRG::Node* n = roots[r];
        b.pops.push_back({n, n->pos_.x(), n->pos_.y(), n->pos_.z(), n->id_, true});
      }
    }
    if (b.pops.empty()) return false;
    sent_ += from_queue;
    b.chain_start = chain_start;
    b.new_xy.clear();
    for (size_t i = handed_nodes_; i < g_.node_seq.size(); ++i) {
      b.new_xy.push_back(g_.node_seq[i]->pos_.x());
      b.new_xy.push_back(g_.node_seq[i]->pos_.y());
    }
    handed_nodes_ = g_.node_seq.size();
    b.grid_count  = handed_nodes_;
    b.deferred.swap(pending_);
    pending_.clear();
    return true;
  }

  void feed(Batch& b) {
    auto ta = Clock::now();
    sampleBatch(b);
    auto tb = Clock::now();
    evalBatch(b);
    b.us_sample += 1e6 * std::chrono::duration<double>(tb - ta).count();
    b.us_eval += 1e6 * since(tb);
  }

  void absorb(Batch& b) {
    t_.stat_["window_launches"] += b.window_launches;
    t_.stat_["window_tests"] += b.window_tests;
    t_.stat_["eval_launches"] += b.eval_launches;
    t_.stat_["edge_evals"] += b.edge_evals;
    t_.stat_["nearest_map"] += b.nearest_map;
    t_.stat_["us_sample"] += (int64_t)b.us_sample;
    t_.stat_["us_eval"] += (int64_t)b.us_eval;
    t_.dev_->batches += (uint64_t)b.batches;
    b.window_launches = b.window_tests = b.eval_launches = b.edge_evals = b.nearest_map = b.batches = 0;
    b.us_sample = b.us_eval = 0;
  }

  // ---- phase A: sampling windows (trg.cpp:384-403) ------------------------------------------
  void sampleBatch(Batch& b) {
    const int S = P_.sample_num;
    const size_t m = b.pops.size();
    b.acc.clear();
    size_t done = 0;
    size_t pos = b.chain_start;  // stream position reached by the chain
    int part_acc = 0, part_trials = 0;
    bool part_open = false;
    int W = std::min(256, std::max(8, t_.tuning_.window));
    guess_.resize(m);
    int stuck = 0;
    while (done < m) {
      const int words = (W + 63) >> 6;
      // window start guesses: exact for the first open node, extrapolated with the running mean
      // draws/pop minus a lead that grows like a random walk for the others
      const size_t n_live = m - done;
      size_t hi = 0;
      for (size_t i = done; i < m; ++i) {
        const size_t k = i - done;
        size_t gpos = pos;
        if (k > 0) {
          // every pop consumes at least min(S, 1001) draws (trg.cpp:389-392)
          const size_t smin = (size_t)std::min(S, 1001);
          const size_t lb   = pos + (size_t)std::max(0, (int)smin - part_acc) + (k - 1) * smin;
          double lead = 2.0 + 2.0 * std::sqrt((double)k * var_);
          lead = std::min(lead, std::max(2.0, 0.5 * ((double)W - 2.0 * mean_)));
          const double rem0 = (double)(S - part_acc) * (mean_ / (double)S);
          const double ex   = (double)pos + rem0 + (double)(k - 1) * mean_ - lead;
          gpos = (size_t)std::max((double)lb, std::floor(ex));
        }
        guess_[i] = gpos;
        hi = std::max(hi, gpos + (size_t)W);
      }
      t_.ensureDraws(hi);
      t_.syncDraws();
      DrawBuffer& db = t_.dev_->draws;
      Arena& in  = t_.dev_->in;
      Arena& out = t_.dev_->out;
      in.reset(Arena::padded(n_live * 2 * sizeof(float)) + Arena::padded(n_live * sizeof(int32_t)));
      out.reset(Arena::padded(n_live * words * sizeof(unsigned long long)));
      const size_t o_xy = in.take(n_live * 2 * sizeof(float));
      const size_t o_fd = in.take(n_live * sizeof(int32_t));
      const size_t o_mk = out.take(n_live * words * sizeof(unsigned long long));
      float* xy   = in.h<float>(o_xy);
      int32_t* fd = in.h<int32_t>(o_fd);
      for (size_t i = done; i < m; ++i) {
        xy[2 * (i - done)]     = b.pops[i].x;
        xy[2 * (i - done) + 1] = b.pops[i].y;
        fd[i - done]           = (int32_t)(guess_[i] - db.base());
      }
      in.h2d(st_);
      out.zero_d(o_mk, n_live * words * sizeof(unsigned long long), st_);
      // every draw offset is (e*cosf, e*sinf): each component is at most e in magnitude (+ rounding)
      K(trgb_sample_window_launch2(map_, in.d<float>(o_xy), in.d<int32_t>(o_fd), db.dev(), (int64_t)n_live, W,
                                   P_.expand_dist * 1.0001f, P_.robot_size, P_.height_threshold, P_.collision_threshold,
                                   out.d<unsigned long long>(o_mk)),
        "trgb_sample_window_launch");
      out.d2h(st_);
      cuda_check(cudaStreamSynchronize(st_), "sync(windows)");
      b.batches++;
      b.window_launches++;
      b.window_tests += (int64_t)n_live * W;
      const unsigned long long* mk = out.h<unsigned long long>(o_mk);
      // host chain scan
      const size_t done_before = done;
      const size_t pos_before  = pos;
      for (size_t i = done; i < m; ++i) {
        Pop& p = b.pops[i];
        if (!part_open) {
          p.draw_start = pos;
          p.acc_begin  = (uint32_t)b.acc.size();
          part_acc = 0;
          part_trials = 0;
          part_open = true;
        }
        if (pos < guess_[i]) break;  // the window starts past the chain position: re-plan from here
        const unsigned long long* w = mk + (i - done_before) * words;
        bool miss = false;
        while (part_acc < S) {
          if (part_trials > 1000) break;  // trg.cpp:390-392 (checked before every draw)
          const size_t rel = pos - guess_[i];
          if (rel >= (size_t)W) { miss = true; break; }
          const bool coll = (w[rel >> 6] >> (rel & 63)) & 1ull;
          const size_t d  = pos - t_.draw_base_;
          ++pos;
          if (coll) { ++part_trials; continue; }
          b.acc.push_back({p.x + t_.draw_xy_[2 * d], p.y + t_.draw_xy_[2 * d + 1]});  // trg.cpp:396-397
          ++part_acc;
        }
        if (miss) break;
        p.consumed  = (int)(pos - p.draw_start);
        p.acc_count = (uint32_t)b.acc.size() - p.acc_begin;
        part_open   = false;
        ++done;
        const double c = (double)p.consumed;
        mean_ += 0.02 * (c - mean_);
        var_  += 0.02 * ((c - mean_) * (c - mean_) - var_);
        if (var_ < 0.25) var_ = 0.25;
      }
      if (done == done_before && pos == pos_before) {
        if (++stuck > 64) throw std::logic_error("trg_b200: sampling windows make no progress");
      } else {
        stuck = 0;
      }
      if (done == done_before) W = std::min(256, W * 2);  // a single pop needs a longer window
    }
    b.chain_end = pos;
  }

  // ---- phase B: nearest node, z and parent edge per accepted sample, plus deferred edges ----
  void evalBatch(Batch& b) {
    const size_t ns = b.acc.size();
    const size_t nd = b.deferred.size();
    const size_t n_new = b.new_xy.size() / 2;
    DeviceSession& dv = *t_.dev_;
    const size_t ne = ns + nd;
    b.z.resize(ns); b.tie.resize(ns); b.stage.resize(ns); b.w.resize(ns); b.d.resize(ns);
    b.nn_idx.resize(ns); b.nn_d2.resize(ns); b.nn_tie.resize(ns);
    b.d_stage.resize(nd); b.d_w.resize(nd); b.d_d.resize(nd);
    if (ne + n_new == 0) return;
    Arena& in  = dv.in;
    Arena& out = dv.out;
    in.reset(Arena::padded(ne * 3 * sizeof(float)) + Arena::padded(ne * 2 * sizeof(float)) + Arena::padded(n_new * 2 * sizeof(float)));
    out.reset(Arena::padded(ns * sizeof(float)) + Arena::padded(ns) + Arena::padded(ne) + 2 * Arena::padded(ne * sizeof(float)) +
              2 * Arena::padded(ns * sizeof(float)) + Arena::padded(ns));
    const size_t o_p1 = in.take(ne * 3 * sizeof(float));
    const size_t o_p2 = in.take(ne * 2 * sizeof(float));
    const size_t o_nn = in.take(n_new * 2 * sizeof(float));
    const size_t o_z  = out.take(ns * sizeof(float));
    const size_t o_t  = out.take(ns);
    const size_t o_s  = out.take(ne);
    const size_t o_w  = out.take(ne * sizeof(float));
    const size_t o_d  = out.take(ne * sizeof(float));
    const size_t o_ni = out.take(ns * sizeof(int32_t));
    const size_t o_nd = out.take(ns * sizeof(float));
    const size_t o_nt = out.take(ns);
    if (n_new) std::memcpy(in.h<float>(o_nn), b.new_xy.data(), n_new * 2 * sizeof(float));
    float* p1 = in.h<float>(o_p1);
    float* p2 = in.h<float>(o_p2);
    size_t k = 0;
    for (const Pop& p : b.pops) {
      for (uint32_t j = 0; j < p.acc_count; ++j, ++k) {
        p1[3 * k] = p.x; p1[3 * k + 1] = p.y; p1[3 * k + 2] = p.z;
        p2[2 * k] = b.acc[p.acc_begin + j].x; p2[2 * k + 1] = b.acc[p.acc_begin + j].y;
      }
    }
    for (const Deferred& d : b.deferred) {
      p1[3 * k] = d.ax; p1[3 * k + 1] = d.ay; p1[3 * k + 2] = d.az;
      p2[2 * k] = d.bx; p2[2 * k + 1] = d.by;
      ++k;
    }
    in.h2d(st_);
    K(trgb_nodes_append_launch(dv.nodes, in.d<float>(o_nn), (int64_t)n_new, st_), "trgb_nodes_append_launch");
    dv.nodes_uploaded += n_new;
    if (ns) {
      K(trgb_nodes_nearest_launch(dv.nodes, in.d<float>(o_p2), (int64_t)ns, out.d<int32_t>(o_ni), out.d<float>(o_nd),
                                  out.d<uint8_t>(o_nt), st_),
        "trgb_nodes_nearest_launch");
      // speculation filter: a sample whose nearest node (already on the device) is closer than
      // robot_size can only be wired to an existing node (trg.cpp:414-417) — no height, no parent edge
      K(trgb_nearest_z_launch_skip(map_, in.d<float>(o_p2), (int64_t)ns, out.d<float>(o_z), nullptr, out.d<uint8_t>(o_t),
                                   out.d<float>(o_nd), P_.robot_size),
        "trgb_nearest_z_launch");
    }
    if (ne) {
      // threads per edge in the segment-collision kernel: deferred edges reach expand_dist + robot_size
      const int kmax = (int)std::ceil((P_.expand_dist + P_.robot_size) / (0.5f * P_.robot_size));
      TrgbEdgeParams prm{P_.robot_size, P_.height_threshold, P_.collision_threshold, std::max(2, std::min(kmax, 16))};
      K(trgb_edge_eval_launch_skip(map_, in.d<float>(o_p1), in.d<float>(o_p2), (int64_t)ne, &prm, out.d<uint8_t>(o_s),
                                   out.d<float>(o_w), out.d<float>(o_d), nullptr, ns ? out.d<float>(o_nd) : nullptr,
                                   (int64_t)ns, P_.robot_size),
        "trgb_edge_eval_launch");
    }
    out.d2h(st_);
    cuda_check(cudaStreamSynchronize(st_), "sync(eval)");
    b.batches++;
    b.eval_launches++;
    b.nearest_map += (int64_t)ns;
    b.edge_evals += (int64_t)ne;
    // results leave the staging arena: the next batch may be fed while this one is committed
    if (ns) {
      std::memcpy(b.z.data(), out.h<float>(o_z), ns * sizeof(float));
      std::memcpy(b.tie.data(), out.h<uint8_t>(o_t), ns);
      std::memcpy(b.stage.data(), out.h<uint8_t>(o_s), ns);
      std::memcpy(b.w.data(), out.h<float>(o_w), ns * sizeof(float));
      std::memcpy(b.d.data(), out.h<float>(o_d), ns * sizeof(float));
      std::memcpy(b.nn_idx.data(), out.h<int32_t>(o_ni), ns * sizeof(int32_t));
      std::memcpy(b.nn_d2.data(), out.h<float>(o_nd), ns * sizeof(float));
      std::memcpy(b.nn_tie.data(), out.h<uint8_t>(o_nt), ns);
    }
    if (nd) {
      std::memcpy(b.d_stage.data(), out.h<uint8_t>(o_s) + ns, nd);
      std::memcpy(b.d_w.data(), out.h<float>(o_w) + ns, nd * sizeof(float));
      std::memcpy(b.d_d.data(), out.h<float>(o_d) + ns, nd * sizeof(float));
    }
  }

  // Resolve deferred evaluations in call order. If a resolved edge a->b exists by now, the entry
  // was the opposite-orientation retry of an evaluation that succeeded: a duplicate, dropped.
  void resolveDeferred(std::vector<Deferred>& list, const uint8_t* stage, const float* w, const float* d) {
    auto drop = [](TRG::Node* n, TRG::Edge* x) {
      auto it = std::find(n->edges_.begin(), n->edges_.end(), x);
      if (it != n->edges_.end()) n->edges_.erase(it);
    };
    for (size_t i = 0; i < list.size(); ++i) {
      Deferred& e = list[i];
      bool ok = stage[i] == TRGB_EDGE_OK;
      if (ok) {
        for (TRG::Edge* x : e.a->edges_)
          if (x->dst_id_ == e.b->id_ && x->dist_ >= 0.f) { ok = false; break; }
      }
      if (ok) {
        e.ea->weight_ = e.eb->weight_ = w[i];
        e.ea->dist_ = e.eb->dist_ = d[i];
      } else {
        drop(e.a, e.ea);
        drop(e.b, e.eb);
      }
    }
    list.clear();
  }

  // synchronous resolution of everything deferred and not yet handed over (end of expansion /
  // validity stall); earlier hand-overs are always resolved before this is reached
  void flushDeferred() {
    const size_t nd = pending_.size();
    if (!nd) return;
    Arena& in  = t_.dev_->in2;
    Arena& out = t_.dev_->out2;
    in.reset(Arena::padded(nd * 3 * sizeof(float)) + Arena::padded(nd * 2 * sizeof(float)));
    out.reset(Arena::padded(nd) + 2 * Arena::padded(nd * sizeof(float)));
    const size_t o_p1 = in.take(nd * 3 * sizeof(float));
    const size_t o_p2 = in.take(nd * 2 * sizeof(float));
    const size_t o_s  = out.take(nd);
    const size_t o_w  = out.take(nd * sizeof(float));
    const size_t o_d  = out.take(nd * sizeof(float));
    float* p1 = in.h<float>(o_p1);
    float* p2 = in.h<float>(o_p2);
    for (size_t k = 0; k < nd; ++k) {
      const Deferred& d = pending_[k];
      p1[3 * k] = d.ax; p1[3 * k + 1] = d.ay; p1[3 * k + 2] = d.az;
      p2[2 * k] = d.bx; p2[2 * k + 1] = d.by;
    }
    in.h2d(st_);
    TrgbEdgeParams prm{P_.robot_size, P_.height_threshold, P_.collision_threshold, 0};
    K(trgb_edge_eval_launch(map_, in.d<float>(o_p1), in.d<float>(o_p2), (int64_t)nd, &prm, out.d<uint8_t>(o_s),
                            out.d<float>(o_w), out.d<float>(o_d), nullptr),
      "trgb_edge_eval_launch");
    out.d2h(st_);
    cuda_check(cudaStreamSynchronize(st_), "sync(flush)");
    t_.dev_->batches++;
    t_.stat_["flush_launches"]++;
    t_.stat_["edge_evals"] += (int64_t)nd;
    resolveDeferred(pending_, out.h<uint8_t>(o_s), out.h<float>(o_w), out.h<float>(o_d));
  }

  // wireEdge(a, b) against an existing node (trg.cpp:254-370): duplicate check + slope gate now,
  // geometry in a later batch. Placeholder edges (dist_ = -1; weight_ = -1 on the origin side, -2 on
  // the far side) keep the reference's edge order. Evaluation is a pure function of the ORIENTED
  // pair, so while (b, a) is pending a call (a, b) is kept as a conditional retry: the reference
  // would run it if (b, a) failed (see resolveDeferred).
  void wireDeferred(TRG::Node* a, TRG::Node* b) {
    if (a->id_ == b->id_) return;
    bool opposite_pending = false;
    for (TRG::Edge* e : a->edges_) {
      if (e->dst_id_ != b->id_) continue;
      if (e->dist_ >= 0.f) return;     // resolved edge exists
      if (e->weight_ == -1.f) return;  // same orientation already pending: same outcome
      opposite_pending = true;
    }
    if (!opposite_pending) {
      for (TRG::Edge* e : b->edges_)
        if (e->dst_id_ == a->id_ && e->dist_ >= 0.f) return;
    }
    if (gate_.rejects(a->pos_, b->pos_)) return;
    TRG::Edge* ea = t_.newEdge(b->id_, -1.f, -1.f);
    TRG::Edge* eb = t_.newEdge(a->id_, -2.f, -1.f);
    a->edges_.push_back(ea);
    b->edges_.push_back(eb);
    pending_.push_back({a, b, ea, eb, a->pos_.x(), a->pos_.y(), a->pos_.z(), b->pos_.x(), b->pos_.y()});
  }

  void applyEdgeOp(const EdgeOp& op) {
    if (op.kind == 0) {
      wireDeferred(op.a, op.b);
    } else {
      op.a->edges_.push_back(t_.newEdge(op.b->id_, op.w, op.d));
      op.b->edges_.push_back(t_.newEdge(op.a->id_, op.w, op.d));
    }
  }

  // ---- phase C: commit a batch in the reference's order (trg.cpp:406-452) -------------------
  void commit(Batch& b, const std::vector<TRG::Node*>& roots, size_t& root_i, int& cur_ref) {
    resolveDeferred(b.deferred, b.d_stage.data(), b.d_w.data(), b.d_d.data());
    table_.prune(static_cast<int>(b.grid_count));
    for (size_t i = 0; i < b.pops.size(); ++i) {
      Pop& p = b.pops[i];
      if (p.is_root) {
        if (head_ != bfs_.size()) break;  // the previous root's BFS is still running: speculation is stale
        if (p.draw_start != t_.draw_next_) break;
        cur_ref = p.ref_id;
        ++root_i;
      } else {
        if (p.draw_start != t_.draw_next_) throw std::logic_error("trg_b200: sampling chain out of step");
        ++head_;
      }
      commitPop(b, p);
    }
    (void)roots;
  }

  void commitPop(Batch& b, const Pop& p) {
    TRG::Node* node = p.node;
    t_.draw_next_ = p.draw_start + (size_t)p.consumed;
    ++n_pops_;
    const TRG::NodeState new_state = (p.ref_id == 0) ? TRG::NodeState::Valid : TRG::NodeState::Frontier;
    const size_t s0 = p.acc_begin;
    for (uint32_t j = 0; j < p.acc_count; ++j) {
      const size_t si = s0 + j;
      const Sample& s = b.acc[si];
      ++n_nearest_;
      FT_BEGIN();
      // kd_nearest2(node_tree, sample) (trg.cpp:408): device candidate (nodes that existed when the
      // batch was handed over) merged with the nodes created since; exact ties -> reference tree order
      TRG::Node* ex;
      {
        float d2 = b.nn_d2[si];
        int seq = b.nn_idx[si];
        bool tie = b.nn_tie[si] != 0;
        if (seq < 0) d2 = std::numeric_limits<float>::infinity();
        FT_LAP(nn_a);
        table_.refine(s.x, s.y, d2, seq, tie);
        FT_LAP(nn_b);
        ex = (tie || seq < 0) ? t_.nearestNode(g_, s.x, s.y) : g_.node_seq[seq];
        FT_LAP(nn_c);
      }
      if (ex->state_ == TRG::NodeState::Invalid) continue;
      FT_LAP(pre);
      if (norm2(ex->pos_.x() - s.x, ex->pos_.y() - s.y) < P_.robot_size) {
        push_op_({node, ex, 0.f, 0.f, 0});  // wireEdge(node, existing_node)
        FT_LAP(wire);
        continue;
      }
      // 2. new node (addNode never fails for id != 0)
      if (b.stage[si] == TRGB_EDGE_SKIPPED) throw std::logic_error("trg_b200: speculation filter skipped a sample that became a node");
      if (b.tie[si]) ++n_zties_;
      Eigen::Vector2f pos2(s.x, s.y);
      FT_LAP(pre);
      TRG::Node* nn = t_.newNode(g_.node_id, pos2, b.z[si], new_state);
      FT_LAP(alloc);
      g_.nodes[g_.node_id] = nn;
      FT_LAP(umap);
      t_.nodeIndexInsert(g_, nn);
      FT_LAP(index);
      table_.insert(s.x, s.y, static_cast<int>(g_.node_seq.size()) - 1);
      FT_LAP(table);
      g_.node_id++;
      // 2.1 wireEdge(node, new): slope gate on the host, geometry from the speculative batch
      bool parent_ok = false;
      if (b.stage[si] == TRGB_EDGE_OK && !gate_.rejects(node->pos_, nn->pos_)) {
        push_op_({node, nn, b.w[si], b.d[si], 1});
        parent_ok = true;
      }
      // 3. wire to the neighbours within expand_dist, in kd result order
      if (step3_) {
        t_.rangeNodesOrdered(g_, nn->pos_.x(), nn->pos_.y(), P_.expand_dist, cand_);
        for (TRG::Node* en : cand_) {
          if (en->state_ == TRG::NodeState::Invalid) continue;
          wireDeferred(nn, en);
        }
        if (!parent_ok && !nn->edges_.empty()) {
          // the node survives only if one of its pending edges does: resolve them now
          ++n_stalls_;
          flushDeferred();
        }
      }
      // 4. `new_node->edges_.size() < 1` (trg.cpp:447): without step 3 the only possible edge is the
      // parent edge; with it the lists are current (operations are applied inline in that mode)
      if (step3_ ? nn->edges_.size() < 1 : !parent_ok) {
        nn->state_ = TRG::NodeState::Invalid;
        continue;
      }
      bfs_.push_back(nn);
      FT_LAP(newnode);
    }
  }
  FT_DECL(nearest); FT_DECL(wire); FT_DECL(newnode); FT_DECL(nn_a); FT_DECL(nn_b); FT_DECL(nn_c);
  FT_DECL(pre); FT_DECL(alloc); FT_DECL(umap); FT_DECL(index); FT_DECL(table);

  ChunkTable table_;
  TRG& t_;
  TRG::trgStruct& g_;
  const decltype(TRG::param_)& P_;
  SlopeGate gate_;
  trgb_map* map_ = nullptr;
  cudaStream_t st_ = nullptr;
  int device_ = 0;
  int64_t n_pops_ = 0, n_nearest_ = 0, n_zties_ = 0, n_stalls_ = 0;
  double us_commit_ = 0, us_wait_ = 0;
  bool step3_ = false;
  double mean_ = 8.0, var_ = 2.0;  // helper-side: running draws/pop statistics
  Batch B_[2];
  std::function<void(const EdgeOp&)> push_op_;
  std::vector<TRG::Node*> bfs_;    // every node ever queued, in queue order
  size_t head_ = 0;                // next queue entry to commit
  size_t sent_ = 0;                // next queue entry not yet placed in a batch
  size_t handed_nodes_ = 0;        // node_seq prefix already handed to the device node grid
  std::vector<size_t> guess_;
  std::vector<Deferred> pending_;  // deferred since the last hand-over
  std::vector<TRG::Node*> cand_;
};

}  // namespace trg_b200

void TRG::runExpansion(const std::vector<Node*>& roots, trgStruct& g) {
  if (roots.empty()) return;
  // the device copy of the draw stream is keyed to the global map's stream; make sure it exists
  trg_b200::Expander ex(*this, g);
  ex.run(roots);
  if (&g == &global_trg_) invalidateDeviceGraph();
}

void TRG::expandGraph(int ref_id, std::string type) {  // trg.cpp:372-454
  trgStruct& graph = *trgMap_.at(type);
  std::vector<Node*> roots{graph.nodes.at(ref_id)};
  runExpansion(roots, graph);
}

// ================================================================================================
// build / update / clean
// ================================================================================================
void TRG::initGraph(bool /*isPreMap*/, Eigen::Vector3f start3d) {  // trg.cpp:36-64
  std::lock_guard<std::mutex> lock(mtx.graph);
  auto t0 = Clock::now();
  trgStruct& graph = *trgMap_["global"];
  // A large previous graph is disposed of on a helper thread (freeing half a million hash nodes takes
  // ~20 ms). clear() would keep the bucket array, which decides the iteration order of the refilled
  // map (hence the renumbering of cleanGraph): an empty map rehashed to the same bucket count is in
  // the identical state (same count, same policy threshold).
  if (graph_disposal_.valid()) graph_disposal_.get();
  if (graph.nodes.size() > (size_t)tuning_.parallel_min_nodes) {
    auto* old = new std::unordered_map<int, Node*>();
    old->swap(graph.nodes);
    graph.nodes.rehash(old->bucket_count());
    graph_disposal_ = std::async(std::launch::async, [old] { delete old; });
  }
  this->resetGraph(graph.type);
  this->resetGraph("local");  // local nodes are pointers into the global graph
  rewindPools();
  requireMap(graph, "initGraph");

  graph.root_pos           = start3d.head(2);
  Eigen::Vector2f root_pos = graph.root_pos;
  root_pos.x()             = root_pos.x() + param_.expand_dist;
  int cnt                  = 0;
  while (!this->addNode(graph.node_id, root_pos, NodeState::Valid, graph.type)) {
    if (cnt > 100) throw std::runtime_error("trg_b200: Failed to generate root node");  // reference: exit(1)
    // g++ evaluates the two constructor arguments right-to-left: y takes the first draw
    const float ry = param_.expand_dist * nextUniform();
    const float rx = param_.expand_dist * nextUniform();
    root_pos = root_pos + Eigen::Vector2f(rx, ry);
    cnt++;
  }
  if (!buildGraphOnDevice(graph)) {  // device-resident BFS + cleanGraph in one go when applicable
    this->expandGraph(graph.node_id - 1, graph.type);
    auto t1 = Clock::now();
    this->cleanGraph(false);
    stat_["us_clean"] += (int64_t)(1e6 * since(t1));
  }
  secs_["init_graph"] = since(t0);
}

void TRG::cleanGraph(bool updateLocal) {  // trg.cpp:491-535
  trgStruct&                     g = *trgMap_["global"];
  std::unordered_map<int, Node*> new_nodes;
  int                            new_id = 0;
  // old id -> node / new id tables instead of the reference's old2new map and per-edge map lookups:
  // same mapping, direct indexing. Every node of the map is in node_seq with id_ == its key
  // (trg.cpp:248-250, 502, 526; loadPrebuiltGraph :96), so the tables fill sequentially.
  int max_id = -1;
  for (Node* n : g.node_seq) max_id = std::max(max_id, n->id_);
  std::vector<Node*> by_id((size_t)(max_id + 1), nullptr);
  std::vector<int>   old2new((size_t)(max_id + 1), -1);
  for (Node* n : g.node_seq) by_id[n->id_] = n;
  // survivors get their new ids in the map's iteration order (:497-504). The new map is keyed
  // 0, 1, 2, ... in that order, so its insertions do not depend on which nodes survive: on large
  // graphs a second thread builds it (same sequence of operator[] calls => same buckets, order and
  // rehash history) while this one is still walking the old map and, later, rewriting edges.
  std::vector<Node*> kept(g.nodes.size(), nullptr);
  const bool piped = g.nodes.size() > (size_t)tuning_.parallel_min_nodes && trg_b200::thread_budget() > 1;
  std::atomic<size_t> published{0};
  std::atomic<bool>   walked{false};
  std::future<void>   builder;
  if (piped) {
    builder = std::async(std::launch::async, [&] {
      size_t k = 0;
      for (;;) {
        const size_t avail = published.load(std::memory_order_acquire);
        for (; k < avail; ++k) new_nodes[(int)k] = kept[k];
        if (walked.load(std::memory_order_acquire) && k == published.load(std::memory_order_acquire)) break;
        if (k == avail) std::this_thread::yield();
      }
    });
  }
  for (auto& node : g.nodes) {
    if (node.second->state_ == NodeState::Invalid || node.second->edges_.size() < 1) continue;
    if (!piped) new_nodes[new_id] = node.second;
    old2new[node.first] = new_id;
    kept[new_id] = node.second;
    new_id++;
    if (piped && (new_id & 1023) == 0) published.store((size_t)new_id, std::memory_order_release);
  }
  published.store((size_t)new_id, std::memory_order_release);
  walked.store(true, std::memory_order_release);
  kept.resize((size_t)new_id);
  // The reference collects the ids of Invalid edge targets in a vector and std::find()s every edge
  // against it (:515); the same predicate is "target node is Invalid", evaluated directly here.
  // Edges are rewritten in place (same order, same values as the reference's fresh copies); nodes
  // are independent of each other, so the rewrite runs on a few threads.
  auto rewrite = [&](size_t b, size_t e) {
    for (size_t k = b; k < e; ++k) {
      Node*  nd    = kept[k];
      auto&  edges = nd->edges_;
      size_t keep  = 0;
      for (Edge* edge : edges) {
        if (by_id[edge->dst_id_]->state_ == NodeState::Invalid) continue;
        edge->dst_id_ = old2new[edge->dst_id_] >= 0 ? old2new[edge->dst_id_] : 0;  // (:518 old2new[] default-inserts 0)
        edges[keep++] = edge;
      }
      edges.resize(keep);
    }
  };
  const size_t nk = kept.size();
  const size_t nthreads = nk > (size_t)tuning_.parallel_min_nodes ? (size_t)trg_b200::thread_budget() : 1;
  if (nthreads <= 1) {
    rewrite(0, nk);
  } else {
    std::vector<std::future<void>> jobs;
    for (size_t t = 0; t < nthreads; ++t)
      jobs.push_back(std::async(std::launch::async, rewrite, nk * t / nthreads, nk * (t + 1) / nthreads));
    for (auto& j : jobs) j.get();
  }
  for (size_t k = 0; k < nk; ++k) kept[k]->id_ = (int)k;
  if (piped) builder.get();
  // the old map's half a million hash nodes are freed on the side (clear() would keep only its
  // bucket array, which the assignment below replaces anyway)
  std::unordered_map<int, Node*> old_nodes;
  old_nodes.swap(g.nodes);
  std::future<void> disposal;
  if (piped) disposal = std::async(std::launch::async, [&old_nodes] { old_nodes.clear(); });
  this->resetGraph(g.type);
  // (the reference copy-assigns; a move leaves the same buckets, order and rehash state)
  g.nodes   = std::move(new_nodes);
  g.node_id = new_id;
  g.node_seq.reserve(g.nodes.size());
  for (auto& node : g.nodes) nodeIndexInsert(g, node.second);  // node_tree order = new map's iteration order (:528-530)
  g.seq_in_iter_order = true;
  if (piped) disposal.get();
  invalidateDeviceGraph();
  if (updateLocal) {
    auto tl = Clock::now();
    this->setLocalGraph(false);
    stat_["us_local_graph"] += (int64_t)(1e6 * since(tl));
  }
}

void TRG::updateGraph() {  // trg.cpp:456-489
  std::lock_guard<std::mutex> lock(mtx.graph);
  auto t0 = Clock::now();
  trgStruct& g = *trgMap_["global"];
  trgStruct& l = *trgMap_["local"];
  // the per-node tests are pure functions of (node, local map, node set at entry): batch them
  std::vector<Node*> order;
  std::vector<float> xy, chk;
  std::vector<uint8_t> far;
  for (auto& node : l.nodes) {
    Node* n = node.second;
    order.push_back(n);
    xy.push_back(n->pos_.x());
    xy.push_back(n->pos_.y());
    Eigen::Vector2f npos2d = n->pos_.head(2);
    far.push_back((npos2d - l.root_pos).norm() > 2.0 * param_.expand_dist ? 1 : 0);
    Eigen::Vector2f dir = npos2d - l.root_pos;  // isFrontier probe (trg.cpp:783-787)
    dir.normalize();
    Eigen::Vector2f check = npos2d + 2 * param_.robot_size * dir;
    chk.push_back(check.x());
    chk.push_back(check.y());
  }
  const int64_t n = (int64_t)order.size();
  std::vector<uint8_t> coll(n, 1);
  std::vector<int32_t> lcnt(n, 0);
  if (n > 0 && l.map_index) {
    K(trgb_collision_batch(l.map_index, xy.data(), n, param_.robot_size, param_.height_threshold,
                           param_.update_collision_threshold, coll.data()), "trgb_collision_batch");
    K(trgb_range_count_batch(l.map_index, chk.data(), n, (float)(0.5 * param_.robot_size), lcnt.data()),
      "trgb_range_count_batch");
  }
  std::vector<Node*> expand_queue;
  for (int64_t i = 0; i < n; ++i) {
    Node* node = order[i];
    if (far[i]) {
      if (coll[i] || node->edges_.size() < 1) {
        node->state_ = NodeState::Invalid;
        continue;
      }
    }
    const bool frontier = countNodesInRange(g, chk[2 * i], chk[2 * i + 1], param_.robot_size) == 0 && lcnt[i] == 0;
    if (frontier && node->state_ == NodeState::Frontier) {
      node->state_ = NodeState::Frontier;
      expand_queue.push_back(node);
      continue;
    }
    expand_queue.push_back(node);
    node->state_ = NodeState::Valid;
  }
  stat_["us_update_tests"] += (int64_t)(1e6 * since(t0));
  runExpansion(expand_queue, g);  // == expandGraph(node->id_, "global") for each, in order
  auto tc = Clock::now();
  this->cleanGraph(true);
  stat_["us_clean"] += (int64_t)(1e6 * since(tc));
  secs_["update_graph"] = since(t0);
}

// ================================================================================================
// path queries
// ================================================================================================
void TRG::invalidateDeviceGraph() {
  if (dev_graph_) trgb_graph_destroy(dev_graph_);
  dev_graph_ = nullptr;
  dev_graph_relaxed_ = 0;
  dev_graph_nodes_.clear();
  built_row_.clear();  // (every change of the global graph comes through here)
}

void TRG::ensureDeviceGraph() {
  if (dev_graph_) return;
  trgStruct& g = *trgMap_["global"];
  // rows by node id; every node of the map is in node_seq with id_ == its key (see cleanGraph)
  int max_id = -1;
  for (Node* nd : g.node_seq) max_id = std::max(max_id, nd->id_);
  const int n = max_id + 1;
  if (n <= 0) throw std::runtime_error("trg_b200: planSafePath on an empty graph");
  dev_graph_nodes_.assign(n, nullptr);
  for (Node* nd : g.node_seq) dev_graph_nodes_[nd->id_] = nd;
  std::vector<int64_t> row(n + 1, 0);
  for (int i = 0; i < n; ++i) row[i + 1] = row[i] + (dev_graph_nodes_[i] ? (int64_t)dev_graph_nodes_[i]->edges_.size() : 0);
  const int64_t e = row[n];
  std::vector<int32_t> col(e), state(n, -1);
  std::vector<float> w(e), d(e), pos(3 * (size_t)n, 0.f);
  auto fill = [&](int b, int en) {
    for (int i = b; i < en; ++i) {
      Node* nd = dev_graph_nodes_[i];
      if (!nd) continue;
      state[i] = (int32_t)nd->state_;
      pos[3 * i] = nd->pos_.x(); pos[3 * i + 1] = nd->pos_.y(); pos[3 * i + 2] = nd->pos_.z();
      int64_t k = row[i];
      for (Edge* ed : nd->edges_) {
        col[k] = ed->dst_id_; w[k] = ed->weight_; d[k] = ed->dist_;
        ++k;
      }
    }
  };
  const int nthreads = n > tuning_.parallel_min_nodes ? trg_b200::thread_budget() : 1;
  if (nthreads <= 1) {
    fill(0, n);
  } else {
    std::vector<std::future<void>> jobs;
    for (int t = 0; t < nthreads; ++t)
      jobs.push_back(std::async(std::launch::async, fill, (int)((int64_t)n * t / nthreads), (int)((int64_t)n * (t + 1) / nthreads)));
    for (auto& j : jobs) j.get();
  }
  TrgbGraphDesc desc{n, e, row.data(), col.data(), w.data(), d.data(), pos.data(), state.data()};
  K(trgb_graph_upload(&dev_graph_, &desc), "trgb_graph_upload");
}

// TRG::setGoal (trg.cpp:537-565) without the side effect: the goal node and whether it is "known".
// Read-only on the graph once the order tree and the hash grid are current (ensureTree /
// ensureGridBuilt), so a batch of queries snaps on several threads.
std::pair<TRG::Node*, bool> TRG::snapGoal(trgStruct& g, const Eigen::Vector3f& goal) {
  static thread_local std::vector<Node*> res;
  rangeNodesOrdered(g, goal.x(), goal.y(), param_.robot_size, res);
  if (!res.empty()) return {res[0], true};  // head of the kd result list = last node the traversal visited
  // Reference (:543-553): linear scan over the node map, strict `<` on dist = sqrt(dx^2+dy^2),
  // i.e. the first node IN MAP ITERATION ORDER among those whose rounded dist is minimal. The
  // grid finds the minimal dist^2; every node whose sqrtf equals the minimal dist is a
  // candidate (several dist^2 values can round to one dist), ranked by iteration order.
  Node* out = nullptr;
  auto nn = g.node_grid.nearest(goal.x(), goal.y());
  if (nn.entry >= 0) {
    const float min_dist = sqrtf(nn.d2);
    float d2max = nn.d2;
    for (int k = 0; k < 8; ++k) {
      const float up = std::nextafter(d2max, std::numeric_limits<float>::infinity());
      if (sqrtf(up) != min_dist) break;
      d2max = up;
    }
    int n_cand = 0;
    Node* only = nullptr;
    const float reach = min_dist * 1.001f + 1e-4f;
    g.node_grid.for_each_within_d2(goal.x(), goal.y(), d2max, reach, [&](int e) { ++n_cand; only = g.node_seq[e]; });
    if (n_cand == 1) {
      out = only;
    } else {
      // rare: rank the candidates by map iteration order (built once per graph state)
      std::lock_guard<std::mutex> lk(iter_rank_mx_);
      if (g.iter_rank.size() != g.nodes.size()) {
        g.iter_rank.clear();
        g.iter_rank.reserve(g.nodes.size());
        size_t k = 0;
        for (auto& node : g.nodes) g.iter_rank[node.second] = k++;
      }
      size_t best = std::numeric_limits<size_t>::max();
      g.node_grid.for_each_within_d2(goal.x(), goal.y(), d2max, reach, [&](int e) {
        Node* c = g.node_seq[e];
        const size_t rk = g.iter_rank.at(c);
        if (rk < best) { best = rk; out = c; }
      });
    }
  }
  return {out, false};
}

void TRG::setGoalUnlocked(Eigen::Vector3f& goal) {  // trg.cpp:537-565
  trgStruct& g = *trgMap_["global"];
  goal_.pose3d = goal;
  goal_.pose2d = goal.head(2);
  ensureTree(g);
  ensureGridBuilt(g);
  const auto r = snapGoal(g, goal);
  goal_.node    = r.first;
  goal_.isKnown = r.second;
}

void TRG::setGoal(Eigen::Vector3f& goal) { setGoalUnlocked(goal); }

bool TRG::checkReadched(Eigen::Vector2f& pos2d) {  // trg.cpp:567-574
  float dist = (goal_.pose2d - pos2d).norm();
  return dist < param_.goal_tolerance;
}

bool TRG::checkReplan(Eigen::Vector2f& pos2d, std::vector<Eigen::Vector3f>& path) {  // trg.cpp:576-601
  if (goal_.node == nullptr) return false;
  float dist2subgoal = norm2(goal_.node->pos_.x() - pos2d.x(), goal_.node->pos_.y() - pos2d.y());
  if (!goal_.isKnown && dist2subgoal < param_.goal_tolerance) return true;
  if (!goal_.isKnown && goal_.node->state_ != NodeState::Frontier) return true;
  trgStruct& g = *trgMap_["global"];
  for (auto& pt : path)
    if (countNodesInRange(g, pt.x(), pt.y(), param_.robot_size) == 0) return true;
  return false;
}

bool TRG::planSafePath(Eigen::Vector2f& start2d, Eigen::Vector3f& goal_pose, std::vector<Eigen::Vector3f>& out_path,
                       float& direct_dist, float& path_length, float& avg_risk) {  // trg.cpp:603-690
  std::lock_guard<std::mutex> lock(mtx.graph);
  auto t0 = Clock::now();
  last_path_ids_.clear();
  trgStruct& g = *trgMap_["global"];
  if (g.nodes.empty()) return false;
  this->setGoalUnlocked(goal_pose);
  Node* start_node = nearestNode(g, start2d.x(), start2d.y());
  direct_dist = norm2(goal_.node->pos_.x() - start_node->pos_.x(), goal_.node->pos_.y() - start_node->pos_.y());
  ensureDeviceGraph();
  const int32_t s = start_node->id_, t = goal_.node->id_;
  uint8_t found = 0;
  float cost = 0.f, plen = 0.f, risk = 0.f;
  int64_t offs[2] = {0, 0};
  std::vector<int32_t> ids(std::max<size_t>(1024, 64));
  int rc = trgb_sssp_batch(dev_graph_, &s, &t, 1, param_.safety_factor, &found, &cost, &plen, &risk, offs, ids.data(),
                           (int64_t)ids.size());
  if (rc == TRGB_E_NOMEM && offs[1] > (int64_t)ids.size()) {
    ids.resize((size_t)offs[1]);
    rc = trgb_sssp_batch(dev_graph_, &s, &t, 1, param_.safety_factor, &found, &cost, &plen, &risk, offs, ids.data(),
                         (int64_t)ids.size());
  }
  K(rc, "trgb_sssp_batch");
  secs_["plan"] = since(t0);
  if (!found) return false;
  for (int64_t k = offs[0]; k < offs[1]; ++k) {
    out_path.push_back(dev_graph_nodes_[ids[k]]->pos_);
    last_path_ids_.push_back(ids[k]);
  }
  path_length = plen;
  avg_risk    = risk;
  return true;
}

void TRG::planSafePathBatch(const float* queries, int64_t n, PathBatch& out) {
  std::lock_guard<std::mutex> lock(mtx.graph);
  auto t0 = Clock::now();
  trgStruct& g = *trgMap_["global"];
  out = PathBatch();
  out.found.assign(n, 0); out.goal_known.assign(n, 0);
  out.cost.assign(n, 0.f); out.path_length.assign(n, 0.f); out.avg_risk.assign(n, 0.f); out.direct_dist.assign(n, 0.f);
  out.offsets.assign(n + 1, 0);
  if (n <= 0 || g.nodes.empty()) return;
  {
    // per-graph preparation, three independent pieces side by side: CSR build + upload (this
    // thread), order tree for goal snapping, hash grid for start snapping
    std::atomic<int64_t> us_tree{0}, us_grid{0};
    int device = 0;
    cudaGetDevice(&device);
    auto f_tree = std::async(std::launch::async, [&, device] {
      cudaSetDevice(device);  // (per-thread state: the tree may be grown by trgb_kdtree_build)
      auto a = Clock::now();
      ensureTree(g);
      us_tree = (int64_t)(1e6 * since(a));
    });
    auto f_grid = std::async(std::launch::async, [&] { auto a = Clock::now(); ensureGridBuilt(g); us_grid = (int64_t)(1e6 * since(a)); });
    auto tc = Clock::now();
    try {
      ensureDeviceGraph();
    } catch (...) {
      f_tree.wait();
      f_grid.wait();
      throw;
    }
    stat_["us_prep_csr"] += (int64_t)(1e6 * since(tc));
    f_tree.get();
    f_grid.get();
    stat_["us_prep_tree"] += us_tree.load();
    stat_["us_prep_grid"] += us_grid.load();
  }
  secs_["plan_prep"] = since(t0);
  std::vector<int32_t> s(n), t(n);
  {
    // start / goal snapping (trg.cpp:611-616): independent per query, read-only on the graph -> helper threads
    std::vector<Node*> gn((size_t)n, nullptr);
    auto snap = [&](int64_t b, int64_t e) {
      for (int64_t i = b; i < e; ++i) {
        const float* q = queries + 5 * i;
        const auto r = snapGoal(g, Eigen::Vector3f(q[2], q[3], q[4]));
        Node* st = nearestNode(g, q[0], q[1]);
        gn[(size_t)i] = r.first;
        s[i] = st->id_;
        t[i] = r.first->id_;
        out.goal_known[i] = r.second ? 1 : 0;
        out.direct_dist[i] = norm2(r.first->pos_.x() - st->pos_.x(), r.first->pos_.y() - st->pos_.y());
      }
    };
    const int threads = n >= 64 ? trg_b200::thread_budget() : 1;
    if (threads <= 1) {
      snap(0, n);
    } else {
      std::vector<std::future<void>> jobs;
      for (int k = 0; k < threads; ++k) jobs.push_back(std::async(std::launch::async, snap, n * k / threads, n * (k + 1) / threads));
      for (auto& j : jobs) j.get();
    }
    // goal_ ends as the reference leaves it after the last query of the batch
    const float* q = queries + 5 * (n - 1);
    goal_.pose3d = Eigen::Vector3f(q[2], q[3], q[4]);
    goal_.pose2d = Eigen::Vector2f(q[2], q[3]);
    goal_.node = gn[(size_t)n - 1];
    goal_.isKnown = out.goal_known[n - 1] != 0;
  }
  secs_["plan_snap"] = since(t0);
  size_t cap = std::max<size_t>((size_t)1 << 20, (size_t)n * 1024);
  out.node_ids.resize(cap);
  int rc = trgb_sssp_batch(dev_graph_, s.data(), t.data(), n, param_.safety_factor, out.found.data(), out.cost.data(),
                           out.path_length.data(), out.avg_risk.data(), out.offsets.data(), out.node_ids.data(), (int64_t)cap);
  if (rc == TRGB_E_NOMEM && out.offsets[n] > (int64_t)cap) {
    cap = (size_t)out.offsets[n];
    out.node_ids.resize(cap);
    rc = trgb_sssp_batch(dev_graph_, s.data(), t.data(), n, param_.safety_factor, out.found.data(), out.cost.data(),
                         out.path_length.data(), out.avg_risk.data(), out.offsets.data(), out.node_ids.data(), (int64_t)cap);
  }
  K(rc, "trgb_sssp_batch");
  out.node_ids.resize((size_t)out.offsets[n]);
  {
    int64_t relaxed = 0, nq_total = 0;
    trgb_graph_stats(dev_graph_, &relaxed, &nq_total);
    stat_["sssp_relaxed_edges"] += relaxed - dev_graph_relaxed_;
    stat_["sssp_queries"] += n;
    dev_graph_relaxed_ = relaxed;
  }
  secs_["plan_batch"] = since(t0);
}

void TRG::refinePath(std::vector<Eigen::Vector3f>& in_path, std::vector<Eigen::Vector3f>& out_path) {  // trg.cpp:692-730
  // point_between == 1 in the reference: every interior point is emitted twice, then a 3-tap mean
  std::vector<Eigen::Vector3f> dense;
  for (int i = 0; i + 1 < (int)in_path.size(); ++i) {
    dense.push_back(in_path[i]);
    dense.push_back(in_path[i + 1]);
  }
  std::vector<Eigen::Vector3f> smooth;
  const int m = (int)dense.size();
  for (int i = 0; i < m; ++i) {
    if (i == m - 1) {
      smooth.push_back(dense[i]);
      break;
    }
    Eigen::Vector3f sum(0.f, 0.f, 0.f);
    int cnt = 0;
    for (int j = i - 1; j < i + 2; ++j) {
      if (j < 0 || j >= m) continue;
      sum += dense[j];
      cnt++;
    }
    smooth.push_back(sum / (float)cnt);
  }
  out_path = smooth;
}

// ================================================================================================
// accessors
// ================================================================================================
std::unordered_map<int, TRG::Node*> TRG::getGraph(std::string type) { return trgMap_.at(type)->nodes; }  // trg.cpp:805-808

std::unordered_map<int, TRG::Node*> TRG::getGraphCopy(std::string type) {  // trg.cpp:810-824
  std::lock_guard<std::mutex>         lock(mtx.graph);
  std::unordered_map<int, TRG::Node*> nodes_copy;
  trgStruct&                          graph = *trgMap_.at(type);
  for (auto& node : graph.nodes) {
    Eigen::Vector2f pos = node.second->pos_.head(2);
    Node* node_copy = new Node(node.second->id_, pos, node.second->pos_.z(), node.second->state_);
    for (auto& edge : node.second->edges_) node_copy->edges_.push_back(new Edge(edge->dst_id_, edge->weight_, edge->dist_));
    nodes_copy[node.first] = node_copy;
  }
  return nodes_copy;
}

void TRG::lockGraph() { mtx.graph.lock(); }
void TRG::unlockGraph() { mtx.graph.unlock(); }

double TRG::lastSeconds(const std::string& what) const {
  auto it = secs_.find(what);
  return it == secs_.end() ? -1.0 : it->second;
}
int64_t TRG::stat(const std::string& what) const {
  if (what == "rng_draws") return (int64_t)draw_next_;
  if (what == "us_draws") return (int64_t)us_draws_;
  if (what == "node_ties") return n_node_ties_;
  if (what == "us_tree_split") return us_tree_parts_[0];
  if (what == "us_tree_device") return us_tree_parts_[1];
  if (what == "us_tree_adopt") return us_tree_parts_[2];
  if (what == "batches") return (int64_t)dev_->batches;
  auto it = stat_.find(what);
  return it == stat_.end() ? 0 : it->second;
}
trgb_map* TRG::deviceMap(const std::string& type) { return trgMap_.at(type)->map_index; }
