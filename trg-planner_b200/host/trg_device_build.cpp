// TRG::initGraph fast path: expandGraph(0) runs as a device-resident BFS (K9, csrc/expand.cu) —
// sampling windows, chain scan, nearest node, height, parent edge AND the commit decisions on the
// GPU — followed by cleanGraph(false), materialised straight into the cleaned host containers the
// reference API exposes (std::unordered_map<int, Node*>, Node::edges_).
//
// What stays on the host, and why:
//  * the sampling stream: std::mt19937 + uniform_real_distribution<float> + glibc cosf / sinf
//    (trg.cpp:395-397) — generated ahead in blocks and pushed to the device;
//  * a pop the device cannot decide bit-exactly (two nodes at the identical float distance from a
//    sample: the reference's kd-tree visit order decides; a slope within 3 ulp of the gate: glibc's
//    atan2f decides): handleInterruptedPop() runs the reference's rules for that one pop;
//  * the two std::unordered_map iteration orders cleanGraph's renumbering goes through
//    (trg.cpp:497-504, 528-530): ids must be bit-exact, so the first map's order is replayed from
//    libstdc++'s own rehash policy and the second map is the real container.
#include <assert.h>
#include <math.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <future>
#include <limits>
#include <stdexcept>
#include <thread>

#include <cuda_runtime.h>

#include "device_session.h"
#include "map_order.h"
#include "trg.h"
#include "trgb_kernels.h"

namespace {
using Clock = std::chrono::steady_clock;
inline double since(Clock::time_point t0) { return std::chrono::duration<double>(Clock::now() - t0).count(); }
[[noreturn]] void fail(const std::string& what) { throw std::runtime_error("trg_b200: " + what + ": " + trgb_last_error()); }
inline void K(int rc, const char* what) {
  if (rc != TRGB_OK) fail(what);
}

using trg_b200::sequential_map_order;

template <class F>
void parallel_for(size_t n, int threads, F&& f) {
  if (threads <= 1 || n < 4096) {
    f(0, n);
    return;
  }
  std::vector<std::future<void>> jobs;
  for (int t = 0; t < threads; ++t)
    jobs.push_back(std::async(std::launch::async, [&f, n, t, threads] { f(n * t / threads, n * (t + 1) / threads); }));
  for (auto& j : jobs) j.get();
}
}  // namespace

void TRG::destroyExpander() {
  if (expander_) trgb_expander_destroy(expander_);
  expander_     = nullptr;
  expander_map_ = nullptr;
}

bool TRG::buildGraphOnDevice(trgStruct& g) {
  if (!tuning_.device_expand) return false;
  if (param_.expand_dist - param_.robot_size < 0.25 * param_.expand_dist) return false;  // step-3 wiring (trg.cpp:429)
  if (param_.sample_num < 1 || param_.sample_num > 32) return false;
  if (g.node_seq.size() != 1) return false;
  trgb_map* map = requireMap(g, "initGraph");
  auto t_begin = Clock::now();

  // ---- engine ------------------------------------------------------------------------------
  const double area = ((double)g.bbox[2] - g.bbox[0] + 4.0) * ((double)g.bbox[3] - g.bbox[1] + 4.0);
  // nodes are pairwise >= robot_size apart (trg.cpp:414-421): hexagonal packing bounds their number
  const double dens = 2.0 / (std::sqrt(3.0) * (double)param_.robot_size * param_.robot_size);
  const int64_t cap = (int64_t)std::min(area * dens * 1.05 + 65536.0, 6.0e8);
  if (expander_ && expander_map_ != map) {
    // a rebuilt map of the same extent keeps the engine's buffers
    if (trgb_expander_rebind(expander_, map, g.bbox[0], g.bbox[1], g.bbox[2], g.bbox[3], cap) == TRGB_OK) expander_map_ = map;
    else destroyExpander();
  }
  if (!expander_) {
    TrgbExpandParams ep{};
    ep.expand_dist = param_.expand_dist;
    ep.robot_size = param_.robot_size;
    ep.height_threshold = param_.height_threshold;
    ep.collision_threshold = param_.collision_threshold;
    ep.sample_num = param_.sample_num;
    ep.max_slope = atan2f(param_.height_threshold, param_.robot_size);  // trg.cpp:269, float overload
    ep.max_pops = tuning_.expand_max_pops;
    ep.window_words = tuning_.expand_window_words;
    ep.new_state = 0;  // ref_id == 0: new nodes are Valid (trg.cpp:420)
    const int rc = trgb_expander_create(&expander_, map, &ep, g.bbox[0], g.bbox[1], g.bbox[2], g.bbox[3], cap);
    if (rc == TRGB_E_STATE || rc == TRGB_E_ARG || rc == TRGB_E_NOMEM) {
      expander_ = nullptr;
      stat_["device_expand_unavailable"]++;
      return false;
    }
    K(rc, "trgb_expander_create");
    expander_map_ = map;
  }
  trgb_expander* e = expander_;
  Node* root = g.node_seq[0];
  compactDraws();
  const size_t pos_start = draw_next_;
  K(trgb_expander_begin(e, root->pos_.x(), root->pos_.y(), root->pos_.z(), (int64_t)pos_start), "trgb_expander_begin");

  // ---- draws: keep the device copy of the stream `ahead` draws in front of the chain ------------
  size_t pushed_end = pos_start;
  auto push_upto = [&](size_t upto) {
    if (upto <= pushed_end) return;
    ensureDraws(upto);
    const size_t have_end = draw_base_ + draw_u_.size();
    // push whole generated blocks: fewer, larger copies
    K(trgb_expander_push_draws(e, draw_xy_.data() + 2 * (pushed_end - draw_base_), (int64_t)(have_end - pushed_end)),
      "trgb_expander_push_draws");
    pushed_end = have_end;
  };
  const int S = param_.sample_num;
  const int group = std::max(1, tuning_.expand_steps);
  // two groups of steps are in flight; a step consumes ~ (S + collisions) draws per pop it takes
  size_t ahead = (size_t)1 << 17;
  push_upto(pos_start + ahead);

  // host mirror of the node list, filled lazily (interrupts only)
  std::vector<float> mx, my, mz;
  std::vector<int8_t> mstate;
  int64_t n_interrupts = 0;
  bool out_of_capacity = false;

  auto handle_interrupt = [&](const TrgbExpandStatus& st) {
    ++n_interrupts;
    if (st.interrupt == 3) { out_of_capacity = true; return; }
    const int64_t n = st.n_nodes;
    const int64_t have = (int64_t)mx.size();
    if (n > have) {
      std::vector<float> xyz((size_t)(n - have) * 3);
      mstate.resize((size_t)n);
      K(trgb_expander_nodes(e, have, n, xyz.data(), mstate.data() + have), "trgb_expander_nodes");
      mx.resize((size_t)n); my.resize((size_t)n); mz.resize((size_t)n);
      for (int64_t i = have; i < n; ++i) {
        mx[i] = xyz[3 * (i - have)]; my[i] = xyz[3 * (i - have) + 1]; mz[i] = xyz[3 * (i - have) + 2];
      }
    }
    int32_t pop_id = -1;
    K(trgb_expander_head_pop(e, &pop_id), "trgb_expander_head_pop");
    const float px = mx[pop_id], py = my[pop_id], pz = mz[pop_id];
    // sampling (trg.cpp:384-403) with the reference's loop, collision bits fetched 64 draws at a time
    size_t pos = (size_t)st.pos;
    std::vector<std::pair<float, float>> samples;
    int trial_sample = 0;
    std::vector<float> qxy;
    std::vector<uint8_t> coll;
    size_t chunk_pos = 0;
    size_t chunk_n = 0;
    while ((int)samples.size() < S) {
      if (trial_sample > 1000) break;
      if (pos >= chunk_pos + chunk_n) {
        chunk_pos = pos;
        chunk_n = 64;
        ensureDraws(chunk_pos + chunk_n);
        qxy.resize(2 * chunk_n);
        coll.resize(chunk_n);
        for (size_t k = 0; k < chunk_n; ++k) {
          const size_t d = chunk_pos + k - draw_base_;
          qxy[2 * k]     = px + draw_xy_[2 * d];
          qxy[2 * k + 1] = py + draw_xy_[2 * d + 1];
        }
        K(trgb_collision_batch(map, qxy.data(), (int64_t)chunk_n, param_.robot_size, param_.height_threshold,
                               param_.collision_threshold, coll.data()), "trgb_collision_batch");
      }
      const size_t k = pos - chunk_pos;
      ++pos;
      if (coll[k]) { ++trial_sample; continue; }
      samples.emplace_back(qxy[2 * k], qxy[2 * k + 1]);
    }
    // per sample: kd_nearest2 over every node so far, exact ties by the reference tree's visit order
    std::vector<float> new_nodes;  // x, y, z, state
    std::vector<int32_t> ra, rb;
    std::vector<float> rw, rd;
    for (auto& s : samples) {
      const float sx = s.first, sy = s.second;
      float best = std::numeric_limits<float>::infinity();
      std::vector<int> cand;
      const size_t N = mx.size();
      for (size_t i = 0; i < N; ++i) {
        const float dx = mx[i] - sx, dy = my[i] - sy;
        const float d2 = dx * dx + dy * dy;
        if (d2 < best) { best = d2; cand.clear(); cand.push_back((int)i); }
        else if (d2 == best) cand.push_back((int)i);
      }
      int ex = cand[0];
      if (cand.size() > 1) {
        ++n_node_ties_;
        ex = trg_b200::first_visited_of(mx.data(), my.data(), cand, sx, sy);
      }
      if (mstate[ex] == -1) continue;                                  // trg.cpp:411
      if (sqrtf(best) < param_.robot_size) {                           // trg.cpp:414
        ra.push_back(pop_id); rb.push_back(ex); rw.push_back(0.f); rd.push_back(0.f);
        continue;
      }
      float z = 0.f;
      uint8_t tie = 0;
      const float xy[2] = {sx, sy};
      K(trgb_nearest_z_batch(map, xy, 1, &z, nullptr, &tie), "trgb_nearest_z_batch");
      if (tie) stat_["z_ties"]++;
      const float p1[3] = {px, py, pz}, p2[3] = {sx, sy, z};
      TrgbEdgeParams prm{param_.robot_size, param_.height_threshold, param_.collision_threshold, 0};
      uint8_t stage = 0;
      float w = 0.f, d = 0.f;
      K(trgb_edge_eval_batch(map, p1, p2, 1, &prm, &stage, &w, &d, nullptr), "trgb_edge_eval_batch");  // incl. glibc slope gate
      const bool valid = stage == TRGB_EDGE_OK;
      const int nid = (int)mx.size();
      mx.push_back(sx); my.push_back(sy); mz.push_back(z); mstate.push_back(valid ? 0 : -1);
      new_nodes.push_back(sx); new_nodes.push_back(sy); new_nodes.push_back(z); new_nodes.push_back(valid ? 0.f : -1.f);
      if (valid) { ra.push_back(pop_id); rb.push_back((int32_t)((uint32_t)nid | 0x80000000u)); rw.push_back(w); rd.push_back(d); }
    }
    K(trgb_expander_apply_pop(e, (int)(new_nodes.size() / 4), new_nodes.data(), (int)ra.size(), ra.data(), rb.data(), rw.data(),
                              rd.data(), (int64_t)pos), "trgb_expander_apply_pop");
  };

  // ---- the renumbered node map (trg.cpp:502, 526) is a fresh std::unordered_map filled with the keys
  // 0 .. m-1 in order, m = number of surviving nodes. Every Valid node ever queued survives (it has its
  // parent edge), so the queue tail is a lower bound of m that only grows: a helper thread performs the
  // insertions (placeholder values) while the BFS runs — same sequence of operator[] calls, hence the
  // same buckets, order and rehash history as the reference's loop.
  std::unordered_map<int, Node*> new_nodes;
  std::atomic<int> map_target{0};
  std::atomic<bool> map_stop{false};
  std::thread map_builder([&new_nodes, &map_target, &map_stop] {
    int k = 0;
    for (;;) {
      const int t = map_target.load(std::memory_order_acquire);
      for (; k < t; ++k) new_nodes[k] = nullptr;
      if (map_stop.load(std::memory_order_acquire) && k >= map_target.load(std::memory_order_acquire)) break;
      if (k >= t) std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
  });
  struct JoinGuard {
    std::thread& t; std::atomic<bool>& stop;
    ~JoinGuard() { stop.store(true, std::memory_order_release); if (t.joinable()) t.join(); }
  } map_guard{map_builder, map_stop};

  // ---- the BFS: two groups of steps in flight, one status poll per group ----------------------
  int hint = 64;
  TrgbExpandStatus st{};
  int slot = 0;
  K(trgb_expander_enqueue(e, group, hint), "trgb_expander_enqueue");
  K(trgb_expander_snapshot(e, slot), "trgb_expander_snapshot");
  int64_t polls = 0;
  while (true) {
    // next group goes out before this one is awaited
    K(trgb_expander_enqueue(e, group, hint), "trgb_expander_enqueue");
    K(trgb_expander_snapshot(e, slot ^ 1), "trgb_expander_snapshot");
    K(trgb_expander_wait(e, slot, &st), "trgb_expander_wait");
    ++polls;
    slot ^= 1;
    if (st.interrupt) {
      K(trgb_expander_wait(e, slot, &st), "trgb_expander_wait");  // the group in flight is idle: drain it
      handle_interrupt(st);
      if (out_of_capacity) {  // cannot happen with the packing bound; the host-driven path takes over from the root
        stat_["device_expand_capacity"]++;
        return false;
      }
      K(trgb_expander_enqueue(e, group, hint), "trgb_expander_enqueue");
      K(trgb_expander_snapshot(e, slot), "trgb_expander_snapshot");
      continue;
    }
    if (st.head >= st.tail) {
      K(trgb_expander_wait(e, slot, &st), "trgb_expander_wait");  // drain the idle group
      if (st.interrupt || st.head < st.tail) throw std::logic_error("trg_b200: device expander restarted after draining");
      break;
    }
    const int q = st.tail - st.head;
    if (st.tail > 1) map_target.store(st.tail, std::memory_order_release);
    hint = std::min(tuning_.expand_max_pops, std::max(64, q + q / 4 + 64));
    ahead = std::min<size_t>((size_t)1 << 22, std::max<size_t>((size_t)1 << 17, (size_t)4 * group * (size_t)hint * (size_t)(S + 2)));
    push_upto((size_t)st.pos + ahead);
    if (polls > 4000000) throw std::logic_error("trg_b200: device expander makes no progress");
  }
  draw_next_ = (size_t)st.pos;
  const double t_bfs = since(t_begin);

  // ---- edges on the device, then everything back ------------------------------------------------
  auto t_fin = Clock::now();
  int64_t n_all = 0, n_dir = 0;
  K(trgb_expander_finalize(e, &n_all, &n_dir), "trgb_expander_finalize");
  const double t_finalize = since(t_fin);
  TrgbExpandedGraph dl{};
  K(trgb_expander_download_view(e, &dl), "trgb_expander_download_view");
  const float *xy = dl.xy, *zz = dl.z, *ew = dl.weight, *ed = dl.dist;
  const int8_t* state = dl.state;
  const int64_t* row = dl.row_ptr;
  const int32_t* col = dl.col;
  (void)n_dir;
  const double t_edges = since(t_fin);

  // ---- cleanGraph(false) (trg.cpp:491-535), straight into the cleaned containers ------------------
  auto t_mat = Clock::now();
  const size_t n = (size_t)n_all;
  // survivors in the iteration order of the map the reference filled with ids 0 .. n-1 (:497-504)
  const std::vector<int> order1 = sequential_map_order(n, g.nodes.bucket_count());
  std::vector<int> old2new(n, -1);
  std::vector<int> kept;
  kept.reserve(n);
  for (int id : order1) {
    if (state[id] == -1 || row[id + 1] == row[id]) continue;  // Invalid or no edges (:498)
    old2new[id] = (int)kept.size();
    kept.push_back(id);
  }
  const size_t m = kept.size();
  // the K7 search graph of the cleaned graph, built by the device from the arrays it still holds
  struct GraphGuard {
    trgb_graph* g = nullptr;
    ~GraphGuard() { if (g) trgb_graph_destroy(g); }
  } search_graph;
  if (m > 0) K(trgb_expander_make_graph(e, old2new.data(), (int32_t)m, &search_graph.g), "trgb_expander_make_graph");
  const int threads = m > (size_t)tuning_.parallel_min_nodes ? trg_b200::thread_budget() : 1;

  // The old containers go first, so that the index structures of the new graph (what the reference's cleanGraph
  // rebuilds as node_tree, :528-530) can grow on helper threads while the node objects are filled. The renumbered
  // map's iteration order (the insertion order of that tree) is again that of a sequentially filled fresh map.
  std::unordered_map<int, Node*> old_nodes;
  old_nodes.swap(g.nodes);
  this->resetGraph(g.type);
  this->resetGraph("local");
  const std::vector<int> order2 = sequential_map_order(m, 1);
  g.seq_xy.resize(2 * m);
  parallel_for(m, threads, [&](size_t b, size_t en) {
    for (size_t i = b; i < en; ++i) {
      const size_t old = (size_t)kept[(size_t)order2[i]];
      g.seq_xy[2 * i] = xy[2 * old];
      g.seq_xy[2 * i + 1] = xy[2 * old + 1];
    }
  });
  std::future<void> f_grid, f_tree;
  const bool eager_index = m >= 20000;
  if (eager_index) {
    ensureGrid(g);
    f_grid = std::async(std::launch::async, [&] { g.node_grid.rebuild(g.seq_xy.data(), (int)m, std::max(1, threads / 2)); });
    int device = 0;
    cudaGetDevice(&device);
    f_tree = std::async(std::launch::async, [&, device] {
      cudaSetDevice(device);  // the current device is per thread: a helper would otherwise build the tree on GPU 0
      std::vector<int> lo(m), hi(m), par(m);
      std::vector<uint8_t> ax(m);
      K(trgb_kdtree_build(g.seq_xy.data(), (int64_t)m, lo.data(), hi.data(), par.data(), ax.data()), "trgb_kdtree_build");
      g.node_tree.adopt(g.seq_xy.data(), (int)m, std::move(lo), std::move(hi), std::move(par), std::move(ax));
    });
  }
  struct Joiner {  // the helpers reference this frame: never leave it (exception included) before they are done
    std::future<void>&a, &b;
    ~Joiner() { if (a.valid()) a.wait(); if (b.valid()) b.wait(); }
  } joiner{f_grid, f_tree};

  // node / edge objects from the pools (recycled across builds), filled in parallel
  rewindPools();
  while (node_pool_.size() < m) {
    Eigen::Vector2f z2(0.f, 0.f);
    node_pool_.emplace_back(0, z2, 0.f, NodeState::Valid);
    node_pool_.back().edges_.reserve(8);
  }
  size_t e_total = 0;
  std::vector<size_t> e_off(m + 1, 0);
  for (size_t k = 0; k < m; ++k) {
    e_off[k] = e_total;
    e_total += (size_t)(row[kept[k] + 1] - row[kept[k]]);
  }
  e_off[m] = e_total;
  while (edge_pool_.size() < e_total) edge_pool_.emplace_back(0, 0.f, 0.f);
  node_used_ = m;
  edge_used_ = e_total;
  dev_graph_nodes_.resize(m);
  parallel_for(m, threads, [&](size_t b, size_t en) {
    for (size_t k = b; k < en; ++k) {
      const int old = kept[k];
      Node& nd = node_pool_[k];
      dev_graph_nodes_[k] = &nd;  // the search graph numbers nodes by pool slot
      nd.id_ = (int)k;
      nd.pos_ = Eigen::Vector3f(xy[2 * (size_t)old], xy[2 * (size_t)old + 1], zz[(size_t)old]);
      nd.state_ = static_cast<NodeState>((int)state[old]);
      nd.edges_.clear();
      size_t eo = e_off[k];
      for (int64_t j = row[old]; j < row[old + 1]; ++j) {
        // no edge leads to an Invalid node here (a node is Invalid from birth and never wired, :411), so
        // the del_edges filter of :505-517 keeps everything; ids are remapped (:518)
        Edge& ed2 = edge_pool_[eo++];
        ed2.dst_id_ = old2new[col[(size_t)j]];
        ed2.weight_ = ew[(size_t)j];
        ed2.dist_ = ed[(size_t)j];
        nd.edges_.push_back(&ed2);
      }
    }
  });
  // the renumbered map: the helper inserted most keys already; finish, then point every key at its node
  map_target.store((int)m, std::memory_order_release);
  map_stop.store(true, std::memory_order_release);
  map_builder.join();
  if (new_nodes.size() != m) throw std::logic_error("trg_b200: renumbered map out of step with the survivors");
  parallel_for(m, threads, [&](size_t b, size_t en) {
    for (size_t k = b; k < en; ++k) new_nodes.find((int)k)->second = &node_pool_[k];
  });
  g.nodes   = std::move(new_nodes);
  g.node_id = (int)m;
  g.node_seq.resize(m);
  parallel_for(m, threads, [&](size_t b, size_t en) {
    for (size_t i = b; i < en; ++i) g.node_seq[i] = &node_pool_[(size_t)order2[i]];
  });
#ifndef NDEBUG
  {
    size_t i = 0;
    for (auto& node : g.nodes) assert(node.second == g.node_seq[i++]);
  }
#endif
  if (f_grid.valid()) f_grid.get();
  if (f_tree.valid()) f_tree.get();
  if (eager_index) g.grid_built = g.tree_built = m;
  dev_graph_ = search_graph.g;  // (resetGraph above dropped the previous one)
  search_graph.g = nullptr;
  dev_graph_relaxed_ = 0;
  built_row_.assign(e_off.begin(), e_off.end());
  const double t_materialize = since(t_mat);

  stat_["pops"] += st.pops;
  stat_["window_tests"] += st.window_tests;
  stat_["window_launches"] += st.steps_active;
  stat_["eval_launches"] += st.steps_active;
  stat_["z_ties"] += st.z_ties;
  stat_["device_steps"] += st.steps;
  stat_["device_steps_active"] += st.steps_active;
  stat_["device_rounds"] += st.rounds;
  stat_["device_redo_pops"] += st.redo_pops;
  stat_["device_interrupts"] += n_interrupts;
  stat_["device_polls"] += polls;
  stat_["device_builds"]++;
  stat_["nearest_node"] += st.samples;
  stat_["edge_evals"] += st.n_req;
  stat_["us_device_bfs"] += (int64_t)(1e6 * t_bfs);
  stat_["us_device_edges"] += (int64_t)(1e6 * t_edges);
  stat_["us_device_finalize"] += (int64_t)(1e6 * t_finalize);
  stat_["us_materialize"] += (int64_t)(1e6 * t_materialize);
  return true;
}

bool TRG::exportBuiltGraph(int32_t* iter_ids, int32_t* ids_sorted, float* pos_xyz, int32_t* state, int64_t* row_ptr,
                           int32_t* col, float* weight, float* dist) {
  trgStruct& g = *trgMap_["global"];
  const size_t m = g.nodes.size();
  if (built_row_.size() != m + 1 || g.node_seq.size() != m || !g.seq_in_iter_order || node_used_ != m) return false;
  const int threads = m > (size_t)tuning_.parallel_min_nodes ? trg_b200::thread_budget() : 1;
  parallel_for(m, threads, [&](size_t b, size_t en) {
    for (size_t k = b; k < en; ++k) {
      const Node& nd = node_pool_[k];
      if (iter_ids) iter_ids[k] = g.node_seq[k]->id_;
      if (ids_sorted) ids_sorted[k] = (int32_t)k;
      if (pos_xyz) { pos_xyz[3 * k] = nd.pos_.x(); pos_xyz[3 * k + 1] = nd.pos_.y(); pos_xyz[3 * k + 2] = nd.pos_.z(); }
      if (state) state[k] = (int32_t)nd.state_;
      if (row_ptr) row_ptr[k] = built_row_[k];
      if (!col && !weight && !dist) continue;
      int64_t e = built_row_[k];
      for (const Edge* ed : nd.edges_) {
        if (col) col[e] = ed->dst_id_;
        if (weight) weight[e] = ed->weight_;
        if (dist) dist[e] = ed->dist_;
        ++e;
      }
    }
  });
  if (row_ptr) row_ptr[m] = built_row_[m];
  return true;
}
