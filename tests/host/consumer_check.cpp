// A C++ consumer of trg-planner_b200/host/trg.h written the way the reference's own consumers use class TRG:
//   TRGPlanner::init / graph thread / planning thread  (src/planner/trg_planner.cpp:23-47, 187-226, 263-292)
//   the ROS nodes' graph visualisation                  (pipelines/ros2/src/ros2_node.cpp:265-341)
//   the pybind module                                   (python/trg_planner/pybind/trg_planner_pybind.cpp:19-25)
// It must compile against the header unchanged (CPU test) and run on a GPU box (GPU test): the numbers it
// prints are compared with the same build driven through the C facade.
#include <cmath>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#include "trg.h"

using PtsDefault = pcl::PointXYZ;

int main(int argc, char** argv) {
  const int side = argc > 1 ? std::atoi(argv[1]) : 160;
  const std::string graph_file = argc > 2 ? argv[2] : "/tmp/consumer_check_graph.json";
  // config/mountain.yaml
  std::shared_ptr<TRG> trg_ = std::make_shared<TRG>(false, 0.6f, 0.3f, 7, 0.16f, 0.1f, 0.5f, 3.0f, 0.8f);
  if (trg_ == nullptr) return 2;
  trg_->reseed(42);  // [+] the reference seeds from std::random_device

  // a rolling heightfield with a wall, 0.1 m lattice (TRGPlanner::loadPrebuiltMap fills the same container)
  pcl::PointCloud<PtsDefault>::Ptr preMapPtr(new pcl::PointCloud<PtsDefault>());
  preMapPtr->clear();
  for (int j = 0; j < side; ++j)
    for (int i = 0; i < side; ++i) {
      const float x = 0.1f * i, y = 0.1f * j;
      float z = 0.15f * std::sin(0.7f * x) * std::cos(0.5f * y);
      if (i > side / 2 && i < side / 2 + 6 && j > side / 4) z += 1.0f;
      preMapPtr->push_back(PtsDefault(x, y, z));
    }
  if (argc > 3) {  // the cloud as built here (libm's float sin / cos), for the facade-driven twin of this run
    std::FILE* f = std::fopen(argv[3], "wb");
    if (!f) return 4;
    for (auto& p : preMapPtr->points) {
      const float xyz[3] = {p.x, p.y, p.z};
      std::fwrite(xyz, sizeof(float), 3, f);
    }
    std::fclose(f);
  }
  trg_->setGlobalMap(preMapPtr);
  Eigen::Vector3f pose3d(2.0f, 2.0f, 0.0f);
  trg_->initGraph(false, pose3d);

  // ros2_node.cpp:265-341: walk nodes and edges under the graph lock
  trg_->lockGraph();
  std::unordered_map<int, TRG::Node*> nodes = trg_->getGraph("global");
  size_t n_edges = 0, n_frontier = 0;
  double sum_w = 0.0;
  for (auto& node : nodes) {
    if (node.second->state_ == TRG::NodeState::Frontier) ++n_frontier;
    for (TRG::Edge* e : node.second->edges_) {
      if (nodes.find(e->dst_id_) == nodes.end()) { trg_->unlockGraph(); std::printf("dangling edge\n"); return 3; }
      sum_w += e->weight_;
      ++n_edges;
    }
  }
  trg_->unlockGraph();

  // planning thread: planSafePath -> refinePath -> checkReadched / checkReplan
  Eigen::Vector2f pose2d(2.0f, 2.0f);
  Eigen::Vector3f goal(0.1f * side - 2.0f, 0.1f * side - 2.5f, 0.0f);
  std::vector<Eigen::Vector3f> raw, smooth;
  float direct_dist = 0.f, raw_path_length = 0.f, avg_risk = 0.f;
  const bool found = trg_->planSafePath(pose2d, goal, raw, direct_dist, raw_path_length, avg_risk);
  if (found) trg_->refinePath(raw, smooth);
  const bool reached = trg_->checkReadched(pose2d);
  const bool replan = found ? trg_->checkReplan(pose2d, raw) : false;

  // graph thread: a scan around the robot, then updateGraph (trg_planner.cpp:187-190)
  pcl::PointCloud<PtsDefault>::Ptr obsPtr(new pcl::PointCloud<PtsDefault>());
  for (auto& p : preMapPtr->points)
    if (std::fabs(p.x - 4.0f) < 3.0f && std::fabs(p.y - 4.0f) < 3.0f) obsPtr->push_back(p);
  Eigen::Vector2f scan_at(4.0f, 4.0f);
  trg_->setLocalMap(scan_at, obsPtr);
  trg_->updateGraph();
  const size_t n_after_update = trg_->getGraph("global").size();

  // save / reset / load (trg_planner.cpp:205-226) and the pybind accessor
  trg_->saveGraph(graph_file);
  trg_->resetGraph("global");
  trg_->resetGraph("local");
  const size_t n_after_reset = trg_->getGraph("global").size();
  trg_->loadPrebuiltGraph(graph_file);
  std::unordered_map<int, TRG::Node*> copy = trg_->getGraphCopy("global");
  const size_t n_loaded = copy.size();
  for (auto& kv : copy) {
    for (TRG::Edge* e : kv.second->edges_) delete e;
    delete kv.second;
  }
  std::vector<Eigen::Vector3f> raw2;
  float d2 = 0.f, l2 = 0.f, r2 = 0.f;
  const bool found2 = trg_->planSafePath(pose2d, goal, raw2, d2, l2, r2);

  std::printf("{\"nodes\": %zu, \"edges\": %zu, \"frontier\": %zu, \"sum_w\": %.9g, \"found\": %d, \"path_pts\": %zu, "
              "\"smooth_pts\": %zu, \"direct_dist\": %.9g, \"path_length\": %.9g, \"avg_risk\": %.9g, \"reached\": %d, "
              "\"replan\": %d, \"nodes_after_update\": %zu, \"nodes_after_reset\": %zu, \"nodes_loaded\": %zu, "
              "\"found_after_load\": %d, \"path_pts_after_load\": %zu}\n",
              nodes.size(), n_edges, n_frontier, sum_w, (int)found, raw.size(), smooth.size(), (double)direct_dist,
              (double)raw_path_length, (double)avg_risk, (int)reached, (int)replan, n_after_update, n_after_reset, n_loaded,
              (int)found2, raw2.size());
  return 0;
}
