// ORACLE — TEST INFRASTRUCTURE ONLY (see trg_oracle.h).
//
// Function-by-function CPU restatement of the reference TRG graph core,
//   cpp/trg_planner/core/trg_planner/src/graph/trg.cpp   (cited as trg.cpp:LINE)
//   cpp/trg_planner/core/trg_planner/include/graph/trg.h (cited as trg.h:LINE)
// on minimal value types. The same libstdc++ containers the reference uses are kept
// (std::unordered_map<int,Node*>, std::deque, std::priority_queue, std::sort,
// std::mt19937 + std::uniform_real_distribution<float>) so iteration order, RNG
// mapping and sort behaviour are the real thing, not a restatement. Eigen value
// arithmetic is restated in eigen_restate.h (PARITY UNPINNED there); the kd-tree is
// either kdtree_port.h (default) or the verbatim reference kdtree.c
// (-DORACLE_USE_REF_KDTREE, built by oracle/Makefile into oracle/_ref/).
//
// Build flags mirror the reference (cpp/trg_planner/CMakeLists.txt:11-15: Release,
// C++17, no -march) plus -ffp-contract=off so no FMA contraction can occur.
#include "trg_oracle.h"

#include <math.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <deque>
#include <functional>
#include <limits>
#include <map>
#include <queue>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>

#include "eigen_restate.h"

#ifdef ORACLE_USE_REF_KDTREE
#include "trg_planner/include/kdtree/kdtree.h"
#else
#include "kdtree_port.h"
#endif

namespace {

struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct PointXYZ { float x, y, z, pad; };  // pcl::PointXYZ is 16 bytes

// ---- kd-tree adapter: payload = integer index --------------------------------
#ifdef ORACLE_USE_REF_KDTREE
struct Kd {
  kdtree* t = kd_create(2);
  ~Kd() { kd_free(t); }
  void clear() { kd_clear(t); }
  void insert(float x, float y, int64_t payload) {
    kd_insert2(t, x, y, reinterpret_cast<void*>(static_cast<intptr_t>(payload + 1)));
  }
  void range(float x, float y, float r, std::vector<int64_t>* out) const {
    out->clear();
    kdres* res = kd_nearest_range2(t, x, y, r);
    while (!kd_res_end(res)) {
      out->push_back(static_cast<int64_t>(reinterpret_cast<intptr_t>(kd_res_item_data(res))) - 1);
      kd_res_next(res);
    }
    kd_res_free(res);
  }
  int64_t nearest(float x, float y) const {
    kdres* res = kd_nearest2(t, x, y);
    if (!res) return -1;
    int64_t p = static_cast<int64_t>(reinterpret_cast<intptr_t>(kd_res_item_data(res))) - 1;
    kd_res_free(res);
    return p;
  }
};
#else
struct Kd {
  kdport::Tree2 t;
  void clear() { t.clear(); }
  void insert(float x, float y, int64_t payload) { t.insert(x, y, payload); }
  void range(float x, float y, float r, std::vector<int64_t>* out) const { t.range(x, y, r, out); }
  int64_t nearest(float x, float y) const { return t.nearest(x, y); }
};
#endif

using Clock = std::chrono::steady_clock;
inline double secs_since(Clock::time_point t0) {
  return std::chrono::duration<double>(Clock::now() - t0).count();
}

// ---- TRG (trg.h:18-147) -------------------------------------------------------
class TRGOracle {
 public:
  struct Edge {  // trg.h:20-25
    Edge(int d, float w, float l) : dst_id_(d), weight_(w), dist_(l) {}
    int dst_id_;
    float weight_;
    float dist_;
  };
  enum NodeState { Valid = 0, Invalid = -1, Frontier = 1 };  // trg.h:27-31
  struct Node {                                               // trg.h:33-40
    Node(int id, V2 p, float z, NodeState s) : id_(id), pos_{p.x, p.y, z}, state_(s) {}
    int id_;
    V3 pos_;
    NodeState state_;
    std::vector<Edge*> edges_;
  };
  struct OptimizeNode {  // trg.h:42-48
    OptimizeNode(int i, float f, float g) : id_(i), parent_(nullptr), f_(f), g_(g) {}
    int id_;
    OptimizeNode* parent_;
    float f_;
    float g_;
  };
  struct Graph {  // trg.h:101-112 trgStruct
    std::unordered_map<int, Node*> nodes;
    Kd node_tree;  // payload: index into node_by_payload
    std::vector<Node*> node_payload;
    int node_id = 0;
    V2 root_pos{0, 0};
    Kd map_tree;  // payload: index into cloud
    std::vector<PointXYZ> cloud;
  };
  struct Param {  // trg.h:132-142
    bool isVerbose;
    float expand_dist, robot_size;
    int sample_num;
    float height_threshold, collision_threshold, update_collision_threshold, safety_factor,
        goal_tolerance;
  } param_;

  Graph global_, local_;
  struct { V3 pose3d; V2 pose2d; Node* node = nullptr; bool isKnown = false; } goal_;  // trg.h:121-126
  std::mt19937 gen_;                             // trg.h:129 (reference: gen_(rd_()))
  std::uniform_real_distribution<float> distr_;  // trg.h:130, trg.cpp:20 distr_(0.0, 1.0)

  std::map<std::string, double> secs_;
  struct Stats { int64_t rng_draws = 0, collision_calls = 0, edge_evals = 0, nearest_map = 0, nearest_node = 0; } stat_;
  std::deque<Node> node_arena_;  // owns nodes (the reference leaks raw new)
  std::deque<Edge> edge_arena_;

  explicit TRGOracle(const OrcParams& p) : gen_(0u), distr_(0.0, 1.0) {  // trg.cpp:11-34
    param_ = {p.is_verbose != 0, p.expand_dist, p.robot_size, p.sample_num, p.height_threshold,
              p.collision_threshold, p.update_collision_threshold, p.safety_factor,
              p.goal_tolerance};
    resetGraph(global_);
    resetGraph(local_);
    resetMap(global_);
    resetMap(local_);
  }

  Graph& graphOf(const std::string& type) { return type == "local" ? local_ : global_; }

  Node* newNode(int id, V2 p, float z, NodeState s) {
    node_arena_.emplace_back(id, p, z, s);
    return &node_arena_.back();
  }
  Edge* newEdge(int d, float w, float l) {
    edge_arena_.emplace_back(d, w, l);
    return &edge_arena_.back();
  }
  void nodeTreeInsert(Graph& g, Node* n) {
    g.node_payload.push_back(n);
    g.node_tree.insert(n->pos_.x, n->pos_.y, static_cast<int64_t>(g.node_payload.size() - 1));
  }

  // trg.cpp:732-737
  void resetGraph(Graph& g) {
    g.nodes.clear();
    g.node_tree.clear();
    g.node_payload.clear();
    g.node_id = 0;
  }
  // trg.cpp:739-744
  void resetMap(Graph& g) {
    g.map_tree.clear();
    g.cloud.clear();
  }

  // trg.cpp:179-193
  void setGlobalMap(const float* xyz, int64_t n) {
    auto t0 = Clock::now();
    resetMap(global_);
    global_.cloud.resize(n);
    for (int64_t i = 0; i < n; ++i) global_.cloud[i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 1.0f};
    for (int64_t i = 0; i < n; ++i) {
      const PointXYZ& pt = global_.cloud[i];
      global_.map_tree.insert(pt.x, pt.y, i);
    }
    secs_["set_global_map"] = secs_since(t0);
  }

  // trg.cpp:195-209
  void setLocalMap(V2 start2d, const float* xyz, int64_t n) {
    auto t0 = Clock::now();
    resetMap(local_);
    local_.root_pos = start2d;
    local_.cloud.resize(n);
    for (int64_t i = 0; i < n; ++i) local_.cloud[i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 1.0f};
    for (int64_t i = 0; i < n; ++i) {
      const PointXYZ& pt = local_.cloud[i];
      local_.map_tree.insert(pt.x, pt.y, i);
    }
    setLocalGraph();
    secs_["set_local_map"] = secs_since(t0);
  }

  // trg.cpp:211-231
  void setLocalGraph() {
    resetGraph(local_);
    std::vector<int64_t> res;
    for (auto& node : global_.nodes) {
      local_.map_tree.range(node.second->pos_.x, node.second->pos_.y,
                            static_cast<float>(param_.robot_size * 0.5), &res);
      if (res.empty()) continue;
      local_.nodes[node.first] = node.second;
      nodeTreeInsert(local_, node.second);
    }
  }

  // trg.cpp:746-778
  bool isCollision(V2 pos, Graph& graph, float threshold) {
    stat_.collision_calls++;
    std::vector<int64_t>& res = scratch_;
    graph.map_tree.range(pos.x, pos.y, param_.robot_size, &res);
    if (res.empty()) return true;
    std::vector<const PointXYZ*>& pts = scratch_pts_;
    pts.clear();
    for (int64_t i : res) pts.push_back(&graph.cloud[i]);
    std::sort(pts.begin(), pts.end(), [](const PointXYZ* a, const PointXYZ* b) { return a->z < b->z; });
    float z_med = pts[pts.size() / 2]->z;
    int total = static_cast<int>(pts.size());
    int cnt = 0;
    for (auto& pt : pts) {
      if (fabsf(pt->z - z_med) > param_.height_threshold) cnt++;
    }
    float ratio = static_cast<float>(cnt) / total;
    if (ratio > threshold) return true;
    return false;
  }

  // trg.cpp:233-252
  bool addNode(int node_id, V2 node_pos, NodeState state, Graph& graph) {
    if (node_id == 0) {
      if (isCollision(node_pos, graph, param_.collision_threshold)) return false;
    }
    stat_.nearest_map++;
    int64_t pi = graph.map_tree.nearest(node_pos.x, node_pos.y);
    float z = graph.cloud[pi].z;
    Node* node = newNode(node_id, node_pos, z, state);
    graph.nodes[node_id] = node;
    nodeTreeInsert(graph, node);
    graph.node_id++;
    return true;
  }

  struct EdgeEval {
    int stage;
    float weight;
    double weight64;
    float dist;
    int npts;
  };

  // trg.cpp:269-363 — everything in wireEdge between the duplicate check and the push
  EdgeEval evalEdge(const V3& p1, const V3& p2, Graph& graph) {
    stat_.edge_evals++;
    EdgeEval out{ORC_EDGE_OK, 0.f, 0.0, 0.f, 0};
    // :269-274 slope gate (float overloads of atan2/fabs, SURVEY.md hard part 2)
    float max_slope = atan2f(param_.height_threshold, param_.robot_size);
    float slope = atan2f(fabsf(p1.z - p2.z), erst::v2_norm(p1.x - p2.x, p1.y - p2.y));
    // :276-278
    float dist = erst::v2_norm(p1.x - p2.x, p1.y - p2.y);
    out.dist = dist;
    if (slope > max_slope) {
      out.stage = ORC_EDGE_SLOPE;
      return out;
    }
    float dirx, diry;
    erst::v2_normalized(p2.x - p1.x, p2.y - p1.y, &dirx, &diry);
    // 0.5 * dist * dir : (double)0.5*dist converted to float, then float * Vector2f
    float half = static_cast<float>(0.5 * dist);
    V2 center{p1.x + half * dirx, p1.y + half * diry};
    // :282-288 segment samples
    float ds = static_cast<float>(param_.robot_size * 0.5);
    for (float i = 0; i < dist; i += ds) {
      V2 pos{p1.x + i * dirx, p1.y + i * diry};
      if (isCollision(pos, graph, param_.collision_threshold)) {
        out.stage = ORC_EDGE_COLLISION;
        return out;
      }
    }
    // :291-297
    float c = static_cast<float>(0.5 * dist);
    float b = param_.robot_size;
    float a = b;
    if (c >= b) a = sqrtf(c * c + b * b);
    bool isCircle = (a == b);
    // :302-326 rotation R = [dir.x -dir.y; dir.y dir.x]
    std::vector<int64_t>& res = scratch_;
    graph.map_tree.range(center.x, center.y, a, &res);
    if (res.empty()) {
      out.stage = ORC_EDGE_EMPTY;
      return out;
    }
    std::vector<float>& rows = scratch_rows_;
    rows.clear();
    for (int64_t idx : res) {
      const PointXYZ& pt = graph.cloud[idx];
      float qx = pt.x - center.x, qy = pt.y - center.y;
      float px = dirx * qx + (-diry) * qy;
      float py = diry * qx + dirx * qy;
      if (isCircle) {
        rows.push_back(px); rows.push_back(py); rows.push_back(pt.z);
      } else if ((px * px) * (b * b) + (py * py) * (a * a) < a * a * b * b) {
        rows.push_back(px); rows.push_back(py); rows.push_back(pt.z);
      }
    }
    out.npts = static_cast<int>(rows.size() / 3);
    if (rows.size() / 3 < 3) {
      out.stage = ORC_EDGE_FEWPTS;
      return out;
    }
    // :332-363
    out.weight = erst::edge_weight_from_rows<float>(rows);
    out.weight64 = erst::edge_weight_from_rows<double>(rows);
    return out;
  }

  // trg.cpp:254-370
  void wireEdge(Node* node1, Node* node2, Graph& graph) {
    if (node1->id_ == node2->id_) return;
    for (auto& edge : node1->edges_)
      if (edge->dst_id_ == node2->id_) return;
    for (auto& edge : node2->edges_)
      if (edge->dst_id_ == node1->id_) return;
    EdgeEval ev = evalEdge(node1->pos_, node2->pos_, graph);
    if (ev.stage != ORC_EDGE_OK) return;
    node1->edges_.push_back(newEdge(node2->id_, ev.weight, ev.dist));
    node2->edges_.push_back(newEdge(node1->id_, ev.weight, ev.dist));
  }

  // trg.cpp:372-454
  void expandGraph(int ref_id, Graph& graph) {
    Node* ref_node = graph.nodes.at(ref_id);
    std::deque<Node*> expand_queue;
    expand_queue.push_back(ref_node);
    std::vector<int64_t> res2;
    while (!expand_queue.empty()) {
      Node* node = expand_queue.front();
      expand_queue.pop_front();
      // :384-403 sampling
      std::vector<V2> samples;
      int max_trial_sample = 1000;
      int trial_sample = 0;
      while (static_cast<int>(samples.size()) < param_.sample_num) {
        if (trial_sample > max_trial_sample) break;
        float expand_dist = param_.expand_dist;
        stat_.rng_draws++;
        float angle = distr_(gen_) * 2 * M_PI;
        V2 sample{node->pos_.x + expand_dist * cosf(angle), node->pos_.y + expand_dist * sinf(angle)};
        if (isCollision(sample, graph, param_.collision_threshold)) {
          trial_sample++;
          continue;
        }
        samples.push_back(sample);
      }
      // :406-452
      for (auto& sample : samples) {
        stat_.nearest_node++;
        Node* existing_node = graph.node_payload[graph.node_tree.nearest(sample.x, sample.y)];
        if (existing_node->state_ == Invalid) continue;
        if (erst::v2_norm(existing_node->pos_.x - sample.x, existing_node->pos_.y - sample.y) <
            param_.robot_size) {
          wireEdge(node, existing_node, graph);
          continue;
        }
        NodeState new_state = (ref_id == 0) ? Valid : Frontier;
        if (!addNode(graph.node_id, sample, new_state, graph)) continue;
        Node* new_node = graph.nodes.at(graph.node_id - 1);
        wireEdge(node, new_node, graph);
        // :429 float - float < double(0.25) * float
        if (param_.expand_dist - param_.robot_size < 0.25 * param_.expand_dist) {
          graph.node_tree.range(new_node->pos_.x, new_node->pos_.y, param_.expand_dist, &res2);
          std::vector<int64_t> cand = res2;  // wireEdge reuses scratch buffers
          for (int64_t pi : cand) {
            Node* en = graph.node_payload[pi];
            if (en->state_ == Invalid) continue;
            wireEdge(new_node, en, graph);
          }
        }
        if (new_node->edges_.size() < 1) {
          new_node->state_ = Invalid;
          continue;
        }
        expand_queue.push_back(new_node);
      }
    }
  }

  // trg.cpp:36-64 ; returns false where the reference calls exit(1)
  bool initGraph(V3 start3d) {
    auto t0 = Clock::now();
    Graph& graph = global_;
    resetGraph(graph);
    graph.root_pos = {start3d.x, start3d.y};
    V2 root_pos = graph.root_pos;
    root_pos.x = root_pos.x + param_.expand_dist;
    int cnt = 0;
    while (!addNode(graph.node_id, root_pos, Valid, graph)) {
      if (cnt > 100) return false;
      // trg.cpp:53-54 builds Vector2f(e*distr_(gen_), e*distr_(gen_)); argument evaluation
      // order is unspecified in C++ — g++ (the reference's compiler) goes right-to-left
      // (probed with g++ 13 at -O0/-O2), so y takes the first draw.
      float ry = param_.expand_dist * distr_(gen_);
      float rx = param_.expand_dist * distr_(gen_);
      stat_.rng_draws += 2;
      root_pos = {root_pos.x + rx, root_pos.y + ry};
      cnt++;
    }
    expandGraph(graph.node_id - 1, graph);
    cleanGraph(false);
    secs_["init_graph"] = secs_since(t0);
    return true;
  }

  // trg.cpp:491-535
  void cleanGraph(bool updateLocal) {
    Graph& g = global_;
    std::unordered_map<int, int> old2new;
    std::unordered_map<int, Node*> new_nodes;
    std::vector<int> del_edges;
    int new_id = 0;
    for (auto& node : g.nodes) {
      if (node.second->state_ == Invalid || node.second->edges_.size() < 1) continue;
      node.second->id_ = new_id;
      new_nodes[new_id] = node.second;
      old2new[node.first] = new_id;
      new_id++;
      for (auto& edge : node.second->edges_) {
        if (g.nodes[edge->dst_id_]->state_ == Invalid) del_edges.push_back(edge->dst_id_);
      }
    }
    // :515 is a linear std::find over del_edges per edge; a sorted copy gives the same answer
    std::vector<int> del_sorted = del_edges;
    std::sort(del_sorted.begin(), del_sorted.end());
    for (auto& node : new_nodes) {
      std::vector<Edge*> new_edges;
      for (auto& edge : node.second->edges_) {
        if (std::binary_search(del_sorted.begin(), del_sorted.end(), edge->dst_id_)) continue;
        new_edges.push_back(newEdge(old2new[edge->dst_id_], edge->weight_, edge->dist_));
      }
      node.second->edges_.clear();
      node.second->edges_ = new_edges;
    }
    resetGraph(g);
    g.nodes = new_nodes;
    g.node_id = new_id;
    for (auto& node : g.nodes) nodeTreeInsert(g, node.second);
    if (updateLocal) setLocalGraph();
  }

  // trg.cpp:780-803
  bool isFrontier(V2 pos) {
    float dx = pos.x - local_.root_pos.x, dy = pos.y - local_.root_pos.y;
    // Vector2f::normalize(): z = squaredNorm(); if (z > 0) *this /= sqrt(z)
    float nx, ny;
    erst::v2_normalized(dx, dy, &nx, &ny);
    float k = 2 * param_.robot_size;
    V2 check{pos.x + k * nx, pos.y + k * ny};
    std::vector<int64_t> r;
    global_.node_tree.range(check.x, check.y, param_.robot_size, &r);
    if (!r.empty()) return false;
    local_.map_tree.range(check.x, check.y, static_cast<float>(0.5 * param_.robot_size), &r);
    if (r.empty()) return true;
    return false;
  }

  // trg.cpp:456-489
  void updateGraph() {
    auto t0 = Clock::now();
    std::deque<Node*> expand_queue;
    for (auto& node : local_.nodes) {
      V2 npos2d{node.second->pos_.x, node.second->pos_.y};
      if (erst::v2_norm(npos2d.x - local_.root_pos.x, npos2d.y - local_.root_pos.y) >
          2.0 * param_.expand_dist) {
        if (isCollision(npos2d, local_, param_.update_collision_threshold) ||
            node.second->edges_.size() < 1) {
          node.second->state_ = Invalid;
          continue;
        }
      }
      if (isFrontier(npos2d) && node.second->state_ == Frontier) {
        node.second->state_ = Frontier;
        expand_queue.push_back(node.second);
        continue;
      }
      expand_queue.push_back(node.second);
      node.second->state_ = Valid;
    }
    while (!expand_queue.empty()) {
      Node* node = expand_queue.front();
      expand_queue.pop_front();
      expandGraph(node->id_, global_);
    }
    cleanGraph(true);
    secs_["update_graph"] = secs_since(t0);
  }

  // trg.cpp:537-565
  void setGoal(V3 goal) {
    goal_.pose3d = goal;
    goal_.pose2d = {goal.x, goal.y};
    std::vector<int64_t> res;
    global_.node_tree.range(goal.x, goal.y, param_.robot_size, &res);
    if (res.empty()) {
      float min_dist = std::numeric_limits<float>::max();
      for (auto& node : global_.nodes) {
        float dist = erst::v2_norm(node.second->pos_.x - goal.x, node.second->pos_.y - goal.y);
        if (dist < min_dist) {
          min_dist = dist;
          goal_.node = node.second;
        }
      }
      goal_.isKnown = false;
    } else {
      goal_.node = global_.node_payload[res[0]];  // head of the result list
      goal_.isKnown = true;
    }
  }

  // trg.cpp:603-690
  bool planSafePath(V2 start2d, V3 goal_pose, std::vector<V3>& out_path, std::vector<int>& out_ids,
                    float& direct_dist, float& path_length, float& avg_risk, int64_t& n_expanded) {
    auto t0 = Clock::now();
    n_expanded = 0;
    setGoal(goal_pose);
    Graph& g = global_;
    if (g.nodes.empty() || goal_.node == nullptr) return false;
    Node* start_node = g.node_payload[g.node_tree.nearest(start2d.x, start2d.y)];

    std::deque<OptimizeNode> arena;  // the reference leaks `new OptimizeNode`
    auto cmp = [](OptimizeNode* a, OptimizeNode* b) { return a->f_ > b->f_; };
    std::priority_queue<OptimizeNode*, std::vector<OptimizeNode*>,
                        std::function<bool(OptimizeNode*, OptimizeNode*)>>
        open_list(cmp);
    std::vector<OptimizeNode*> open_check(g.nodes.size(), nullptr);

    direct_dist = erst::v2_norm(goal_.node->pos_.x - start_node->pos_.x,
                                goal_.node->pos_.y - start_node->pos_.y);
    double g_cost = 0.0;
    double f_cost = g_cost + direct_dist;
    arena.emplace_back(start_node->id_, static_cast<float>(f_cost), static_cast<float>(g_cost));
    OptimizeNode* st = &arena.back();
    st->parent_ = nullptr;
    open_list.push(st);
    open_check[st->id_] = st;
    std::vector<OptimizeNode*> close_list(g.nodes.size(), nullptr);

    bool found = false;
    while (!open_list.empty()) {
      OptimizeNode* opti_node = open_list.top();
      open_list.pop();
      open_check[opti_node->id_] = nullptr;
      if (opti_node->id_ == goal_.node->id_) {
        OptimizeNode* node = opti_node;
        float sum_dist = 0.0;
        float sum_weight = 0.0;
        float avg_weight = 0.0;
        while (node != nullptr) {
          Node* n = g.nodes.at(node->id_);
          for (auto& edge : n->edges_) {
            if (node->parent_ != nullptr && edge->dst_id_ == node->parent_->id_) {
              sum_dist += edge->dist_;
              sum_weight += edge->weight_;
              break;
            }
          }
          out_path.push_back(n->pos_);
          out_ids.push_back(n->id_);
          node = node->parent_;
        }
        avg_weight = sum_weight / out_path.size();
        std::reverse(out_path.begin(), out_path.end());
        std::reverse(out_ids.begin(), out_ids.end());
        path_length = sum_dist;
        avg_risk = avg_weight;
        found = true;
        break;
      }
      Node* curr_node = g.nodes.at(opti_node->id_);
      close_list[curr_node->id_] = opti_node;
      n_expanded++;
      for (auto e : curr_node->edges_) {
        Node* dst_node = g.nodes.at(e->dst_id_);
        if (close_list[dst_node->id_] != nullptr || dst_node->state_ == Invalid) continue;
        double next_g_cost = opti_node->g_ + (param_.safety_factor * e->weight_ + 1) * e->dist_;
        double next_f_cost = next_g_cost + erst::v2_norm(goal_.node->pos_.x - dst_node->pos_.x,
                                                         goal_.node->pos_.y - dst_node->pos_.y);
        arena.emplace_back(dst_node->id_, static_cast<float>(next_f_cost),
                           static_cast<float>(next_g_cost));
        OptimizeNode* dn = &arena.back();
        dn->parent_ = opti_node;
        if (open_check[dst_node->id_] == nullptr) {
          open_list.push(dn);
          open_check[dst_node->id_] = dn;
        } else if (dn->g_ < open_check[dst_node->id_]->g_) {
          open_list.push(dn);
          open_check[dst_node->id_] = dn;
        }
      }
    }
    secs_["plan"] = secs_since(t0);
    return found;
  }

  // trg.cpp:567-574
  bool checkReadched(V2 pos2d) {
    float dist = erst::v2_norm(goal_.pose2d.x - pos2d.x, goal_.pose2d.y - pos2d.y);
    if (dist < param_.goal_tolerance) return true;
    return false;
  }

  // trg.cpp:576-601
  bool checkReplan(V2 pos2d, const std::vector<V3>& path) {
    if (goal_.node == nullptr) return false;
    float dist2subgoal = erst::v2_norm(goal_.node->pos_.x - pos2d.x, goal_.node->pos_.y - pos2d.y);
    if (!goal_.isKnown && dist2subgoal < param_.goal_tolerance) return true;
    if (!goal_.isKnown && goal_.node->state_ != Frontier) return true;
    std::vector<int64_t> r;
    for (auto& pt : path) {
      global_.node_tree.range(pt.x, pt.y, param_.robot_size, &r);
      if (r.empty()) return true;
    }
    return false;
  }

  // trg.cpp:692-730 (point_between = 1)
  void refinePath(const std::vector<V3>& in_path, std::vector<V3>& out_path) {
    std::deque<V3> dense_path;
    for (int i = 0; i + 1 < static_cast<int>(in_path.size()); ++i) {
      dense_path.push_back(in_path[i]);
      dense_path.push_back(in_path[i + 1]);
    }
    std::deque<V3> smooth_path;
    for (int i = 0; i < static_cast<int>(dense_path.size()); ++i) {
      if (i == static_cast<int>(dense_path.size()) - 1) {
        smooth_path.push_back(dense_path[i]);
        break;
      }
      V3 sum{0.f, 0.f, 0.f};
      int cnt = 0;
      for (int j = i - 1; j < i + 2; ++j) {
        if (j < 0 || j >= static_cast<int>(dense_path.size())) continue;
        sum = {sum.x + dense_path[j].x, sum.y + dense_path[j].y, sum.z + dense_path[j].z};
        cnt++;
      }
      float fc = static_cast<float>(cnt);
      smooth_path.push_back({sum.x / fc, sum.y / fc, sum.z / fc});
    }
    out_path.assign(smooth_path.begin(), smooth_path.end());
  }

  std::vector<int64_t> scratch_;
  std::vector<const PointXYZ*> scratch_pts_;
  std::vector<float> scratch_rows_;
};

inline TRGOracle* H(void* h) { return static_cast<TRGOracle*>(h); }

}  // namespace

extern "C" {

void* orc_create(const OrcParams* p) { return new TRGOracle(*p); }
void orc_destroy(void* h) { delete H(h); }
void orc_seed(void* h, uint32_t seed) {
  H(h)->gen_.seed(seed);
  H(h)->distr_.reset();
}

int orc_set_global_map(void* h, const float* xyz, int64_t n) {
  H(h)->setGlobalMap(xyz, n);
  return 0;
}
int orc_set_local_map(void* h, float sx, float sy, const float* xyz, int64_t n) {
  H(h)->setLocalMap({sx, sy}, xyz, n);
  return 0;
}
int orc_init_graph(void* h, int, float sx, float sy, float sz) {
  if (H(h)->global_.cloud.empty()) return -2;
  return H(h)->initGraph({sx, sy, sz}) ? 0 : -1;
}
int orc_update_graph(void* h) {
  H(h)->updateGraph();
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
    auto* n = g.nodes.at(ids[i]);
    if (ids_sorted) ids_sorted[i] = ids[i];
    if (pos_xyz) {
      pos_xyz[3 * i] = n->pos_.x; pos_xyz[3 * i + 1] = n->pos_.y; pos_xyz[3 * i + 2] = n->pos_.z;
    }
    if (state) state[i] = static_cast<int32_t>(n->state_);
    if (row_ptr) row_ptr[i] = e;
    for (auto* ed : n->edges_) {
      if (col) col[e] = ed->dst_id_;
      if (weight) weight[e] = ed->weight_;
      if (dist) dist[e] = ed->dist_;
      ++e;
    }
  }
  if (row_ptr) row_ptr[ids.size()] = e;
  return 0;
}

int orc_plan(void* h, float sx, float sy, float gx, float gy, float gz, float* path_xyz,
             int32_t* node_ids, int max_pts, int* n_pts, float* direct_dist, float* path_length,
             float* avg_risk, int* goal_known, int64_t* n_expanded) {
  std::vector<V3> path;
  std::vector<int> ids;
  float dd = 0, pl = 0, ar = 0;
  int64_t ne = 0;
  bool ok = H(h)->planSafePath({sx, sy}, {gx, gy, gz}, path, ids, dd, pl, ar, ne);
  if (direct_dist) *direct_dist = dd;
  if (path_length) *path_length = pl;
  if (avg_risk) *avg_risk = ar;
  if (goal_known) *goal_known = H(h)->goal_.isKnown ? 1 : 0;
  if (n_expanded) *n_expanded = ne;
  int n = static_cast<int>(path.size());
  if (n_pts) *n_pts = n;
  for (int i = 0; i < n && i < max_pts; ++i) {
    if (path_xyz) {
      path_xyz[3 * i] = path[i].x; path_xyz[3 * i + 1] = path[i].y; path_xyz[3 * i + 2] = path[i].z;
    }
    if (node_ids) node_ids[i] = ids[i];
  }
  return ok ? 1 : 0;
}

int orc_refine_path(void* h, const float* in_xyz, int n_in, float* out_xyz, int* n_out) {
  std::vector<V3> in(n_in), out;
  for (int i = 0; i < n_in; ++i) in[i] = {in_xyz[3 * i], in_xyz[3 * i + 1], in_xyz[3 * i + 2]};
  H(h)->refinePath(in, out);
  *n_out = static_cast<int>(out.size());
  for (size_t i = 0; i < out.size(); ++i) {
    out_xyz[3 * i] = out[i].x; out_xyz[3 * i + 1] = out[i].y; out_xyz[3 * i + 2] = out[i].z;
  }
  return 0;
}

int orc_check_reached(void* h, float x, float y) { return H(h)->checkReadched({x, y}) ? 1 : 0; }
int orc_check_replan(void* h, float x, float y, const float* path_xyz, int n_path) {
  std::vector<V3> path(n_path);
  for (int i = 0; i < n_path; ++i) path[i] = {path_xyz[3 * i], path_xyz[3 * i + 1], path_xyz[3 * i + 2]};
  return H(h)->checkReplan({x, y}, path) ? 1 : 0;
}

int orc_is_collision_batch(void* h, const char* type, const float* xy, int64_t n, float threshold,
                           uint8_t* out) {
  auto& g = H(h)->graphOf(type);
  for (int64_t i = 0; i < n; ++i) out[i] = H(h)->isCollision({xy[2 * i], xy[2 * i + 1]}, g, threshold) ? 1 : 0;
  return 0;
}

int orc_range_count_batch(void* h, const char* type, const float* xy, int64_t n, float radius,
                          int32_t* out) {
  auto& g = H(h)->graphOf(type);
  std::vector<int64_t> r;
  for (int64_t i = 0; i < n; ++i) {
    g.map_tree.range(xy[2 * i], xy[2 * i + 1], radius, &r);
    out[i] = static_cast<int32_t>(r.size());
  }
  return 0;
}

int orc_nearest_z_batch(void* h, const char* type, const float* xy, int64_t n, float* z_out,
                        int64_t* idx_out, uint8_t* tie_out) {
  auto& g = H(h)->graphOf(type);
  if (g.cloud.empty()) return -2;
  std::vector<int64_t> r;
  for (int64_t i = 0; i < n; ++i) {
    float x = xy[2 * i], y = xy[2 * i + 1];
    int64_t pi = g.map_tree.nearest(x, y);
    if (z_out) z_out[i] = g.cloud[pi].z;
    if (idx_out) idx_out[i] = pi;
    if (tie_out) {
      // a tie = another point at exactly the same float dist_sq (kd visit order decides)
      float dx = g.cloud[pi].x - x, dy = g.cloud[pi].y - y;
      float d2 = 0; d2 += dx * dx; d2 += dy * dy;
      float rad = sqrtf(d2) * 1.0001f + 1e-6f;
      g.map_tree.range(x, y, rad, &r);
      int cnt = 0;
      for (int64_t j : r) {
        float ex = g.cloud[j].x - x, ey = g.cloud[j].y - y;
        float e2 = 0; e2 += ex * ex; e2 += ey * ey;
        if (e2 == d2) cnt++;
      }
      tie_out[i] = cnt > 1 ? 1 : 0;
    }
  }
  return 0;
}

int orc_edge_eval_batch(void* h, const char* type, const float* p1, const float* p2, int64_t n,
                        uint8_t* stage, float* weight, double* weight64, float* dist,
                        int32_t* npts) {
  auto& g = H(h)->graphOf(type);
  for (int64_t i = 0; i < n; ++i) {
    auto ev = H(h)->evalEdge({p1[3 * i], p1[3 * i + 1], p1[3 * i + 2]},
                             {p2[3 * i], p2[3 * i + 1], p2[3 * i + 2]}, g);
    if (stage) stage[i] = static_cast<uint8_t>(ev.stage);
    if (weight) weight[i] = ev.weight;
    if (weight64) weight64[i] = ev.weight64;
    if (dist) dist[i] = ev.dist;
    if (npts) npts[i] = ev.npts;
  }
  return 0;
}

int orc_is_frontier_batch(void* h, const float* xy, int64_t n, uint8_t* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = H(h)->isFrontier({xy[2 * i], xy[2 * i + 1]}) ? 1 : 0;
  return 0;
}

double orc_last_seconds(void* h, const char* what) {
  auto it = H(h)->secs_.find(what);
  return it == H(h)->secs_.end() ? -1.0 : it->second;
}
int64_t orc_stat(void* h, const char* what) {
  const auto& s = H(h)->stat_;
  std::string w(what);
  if (w == "rng_draws") return s.rng_draws;
  if (w == "collision_calls") return s.collision_calls;
  if (w == "edge_evals") return s.edge_evals;
  if (w == "nearest_map") return s.nearest_map;
  if (w == "nearest_node") return s.nearest_node;
  return -1;
}

}  // extern "C"
