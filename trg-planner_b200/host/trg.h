// TRG — traversal risk graph, B200-native drop-in for the reference class
//   cpp/trg_planner/core/trg_planner/include/graph/trg.h:18-147.
//
// The public interface (nested types with their field names, the 9-argument constructor and
// the 24 public methods) is source-compatible with the reference, so TRGPlanner
// (planner.cpp:23-31,47,187-190,198,205,212-213,222,263-292), the pybind module
// (trg_planner_pybind.cpp:19-78) and the ROS nodes compile against it unchanged. What changed
// is everything behind it: the two kd-trees per graph (trg.h:106,110) are gone — the map lives
// in HBM as a cell index (include/trgb_kernels.h) and graph construction runs as a
// wavefront scheduler that batches the pure-function work (collision tests, nearest-z,
// edge PCA) of one BFS generation into a few kernel launches while committing decisions in
// the reference's exact sequential order.
//
// Additions to the reference API are marked [+]. There is no CPU fallback: methods that need
// the device throw std::runtime_error when no usable GPU / kernel library is present.
#ifndef TRG_PLANNER_B200_HOST_TRG_H_
#define TRG_PLANNER_B200_HOST_TRG_H_

#include <atomic>
#include <cstdint>
#include <deque>
#include <future>
#include <memory>
#include <mutex>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>

#include "node_index.h"
#include "trg_types.h"

#define EPS 1e-6

struct trgb_map;
struct trgb_graph;
struct trgb_expander;

namespace trg_b200 {
class DeviceSession;
class Expander;
std::string json_number(float f);  // a float as nlohmann::json::dump writes it (trg_io.cpp; saveGraph)
}  // namespace trg_b200

class TRG {
 public:
  struct Edge {
    Edge(int dst_id, float weight, float dist) : dst_id_(dst_id), weight_(weight), dist_(dist) {}
    int   dst_id_;
    float weight_;
    float dist_;
  };

  enum struct NodeState {
    Valid    = 0,
    Invalid  = -1,
    Frontier = 1,
  };

  struct Node {
    Node(int id, Eigen::Vector2f& pos2d, float z, NodeState state)
        : id_(id), pos_(Eigen::Vector3f(pos2d.x(), pos2d.y(), z)), state_(state) {}
    int                id_;
    Eigen::Vector3f    pos_;
    NodeState          state_;
    std::vector<Edge*> edges_;
  };

  struct OptimizeNode {
    OptimizeNode(int i, float f, float g) : id_(i), f_(f), g_(g) {}
    int           id_;
    OptimizeNode* parent_;
    float         f_;
    float         g_;
  };

 public:
  TRG(bool  isVerbose,
      float expand_dist,
      float robot_size,
      int   sample_num,
      float height_threshold,
      float collision_threshold,
      float update_collision_threshold,
      float safety_factor,
      float goal_tolerance);
  virtual ~TRG();

  void initGraph(bool isPreMap, Eigen::Vector3f start3d = Eigen::Vector3f::Zero());
  void loadPrebuiltGraph(const std::string& filepath);
  void saveGraph(const std::string& filepath);

  void setGlobalMap(PointCloudPtr& map);
  void setLocalMap(Eigen::Vector2f start2d, PointCloudPtr& map);
  void setLocalGraph(bool useMutex = false);

  bool addNode(int node_id, Eigen::Vector2f& node_pos, NodeState state, std::string type);
  void wireEdge(Node* node1, Node* node2, std::string type);

  void expandGraph(int node_id, std::string type);
  void cleanGraph(bool updateLocal = true);
  void updateGraph();

  void setGoal(Eigen::Vector3f& goal);
  bool checkReadched(Eigen::Vector2f& pos2d);
  bool checkReplan(Eigen::Vector2f& pos2d, std::vector<Eigen::Vector3f>& path);

  bool planSafePath(Eigen::Vector2f&              start2d,
                    Eigen::Vector3f&              goal_pose,
                    std::vector<Eigen::Vector3f>& out_path,
                    float&                        direct_dist,
                    float&                        path_length,
                    float&                        avg_risk);
  void refinePath(std::vector<Eigen::Vector3f>& in_path, std::vector<Eigen::Vector3f>& out_path);

  void resetGraph(std::string type);
  void resetMap(std::string type);

  bool isCollision(Eigen::Vector2f& pos2d, std::string type, float threshold = 0.1);
  bool isFrontier(Eigen::Vector2f& pos2d);

  std::unordered_map<int, Node*> getGraph(std::string type = "global");
  std::unordered_map<int, Node*> getGraphCopy(std::string type = "global");
  // [+] the same map by reference (no half-million-entry copy); the caller holds lockGraph()
  const std::unordered_map<int, Node*>& getGraphRef(const std::string& type = "global") const { return trgMap_.at(type)->nodes; }
  void                           lockGraph();
  void                           unlockGraph();

  // ---- [+] additions --------------------------------------------------------------------
  // Reseed the sampling stream (the reference seeds gen_ from std::random_device, trg.cpp:20,
  // and is therefore not reproducible). Also drops the look-ahead buffer of generated draws.
  void reseed(uint32_t seed);
  // Raw-pointer map ingest: n records of `stride_floats` floats (3 = xyz, 4 = pcl::PointXYZ).
  // `device` != 0 means `xyz` already lives in HBM (bench "inputs resident" path).
  void setGlobalMapRaw(const float* xyz, int64_t n, int stride_floats, bool device = false);
  void setLocalMapRaw(Eigen::Vector2f start2d, const float* xyz, int64_t n, int stride_floats);
  // Batched planSafePath: queries rows (sx, sy, gx, gy, gz); one GPU launch sequence for all.
  struct PathBatch {
    std::vector<uint8_t> found, goal_known;
    std::vector<float>   cost, path_length, avg_risk, direct_dist;
    std::vector<int64_t> offsets;   // n+1
    std::vector<int32_t> node_ids;  // concatenated start..goal id sequences
  };
  void planSafePathBatch(const float* queries, int64_t n, PathBatch& out);
  // Pure batched evaluations (kernel-level parity / benchmarking)
  void isCollisionBatch(const float* xy, int64_t n, const std::string& type, float threshold,
                        uint8_t* out);
  // wall-clock seconds of the last call at the reference's timer sites (planner.cpp:185-270)
  // [+] The global graph as flat arrays without walking the node map, when it still is what the device build
  // materialised (ids == pool slots, edges of a node consecutive in the edge pool): rows by id. Returns false when
  // the graph has changed since (the caller then walks the map). Any output may be null. The caller holds lockGraph().
  bool exportBuiltGraph(int32_t* iter_ids, int32_t* ids_sorted, float* pos_xyz, int32_t* state, int64_t* row_ptr,
                        int32_t* col, float* weight, float* dist);
  double lastSeconds(const std::string& what) const;
  int64_t stat(const std::string& what) const;
  trgb_map* deviceMap(const std::string& type);
  const std::vector<int32_t>& lastPathIds() const { return last_path_ids_; }
  bool goalKnown() const { return goal_.isKnown; }

 protected:
  struct trgStruct {
    explicit trgStruct(std::string type) : type(type), node_id(0) {}
    std::string type;

    std::unordered_map<int, Node*> nodes;
    int                            node_id;
    Eigen::Vector2f                root_pos = Eigen::Vector2f::Zero();
    PointCloudPtr                  cloud_map = nullptr;

    // replaces `kdtree* node_tree` (trg.h:106): insertion sequence + grid + lazy order tree
    std::vector<Node*>      node_seq;
    std::vector<float>      seq_xy;      // (x, y) of node_seq, flat: the index structures are built without chasing Node pointers
    trg_b200::NodeGrid      node_grid;
    trg_b200::OrderTree2D   node_tree;
    size_t                  grid_built = 0;  // prefix of node_seq already in node_grid (filled lazily)
    size_t                  tree_built = 0;  // prefix of node_seq already in node_tree
    bool                    seq_in_iter_order = true;  // node_seq is also the iteration order of `nodes` (true right after cleanGraph)
    std::unordered_map<const Node*, size_t> iter_rank;  // lazily: position in `nodes` iteration order
    // replaces `kdtree* map_tree` (trg.h:110): device cell index
    trgb_map*               map_index = nullptr;
    int64_t                 map_points = 0;
    float                   bbox[4] = {0, 0, 0, 0};
  };
  trgStruct prebuilt_trg_ = trgStruct("prebuilt");
  trgStruct global_trg_   = trgStruct("global");
  trgStruct local_trg_    = trgStruct("local");

  std::unordered_map<std::string, trgStruct*> trgMap_ = {{"prebuilt", &prebuilt_trg_},
                                                         {"global", &global_trg_},
                                                         {"local", &local_trg_}};

  struct goalStruct {
    Eigen::Vector3f pose3d;
    Eigen::Vector2f pose2d;
    Node*           node    = nullptr;
    bool            isKnown = false;
  } goal_;

  std::random_device                    rd_;
  std::mt19937                          gen_;
  std::uniform_real_distribution<float> distr_;

  struct Param {
    bool  isVerbose                  = true;
    float expand_dist                = 0.5;
    float robot_size                 = 0.5;
    int   sample_num                 = 20;
    float height_threshold           = 0.5;
    float collision_threshold        = 0.5;
    float update_collision_threshold = 0.5;
    float safety_factor              = 3.0;
    float goal_tolerance             = 0.2;
  } param_;

  struct Mutex {
    std::mutex graph;
  } mtx;

 private:
  friend class trg_b200::Expander;
  // Buffered view of (gen_, distr_): position k of the stream, generated on demand in blocks and
  // consumed strictly in order. The sampling-window kernel reads the derived offsets
  // (expand_dist*cosf(angle_k), expand_dist*sinf(angle_k)) from a device-resident copy.
  float nextUniform();
  void  ensureDraws(size_t upto);
  void  compactDraws();
  void  generateDrawBlock(size_t n, std::vector<float>& u, std::vector<float>& xy);
  void  joinDrawPrefetch();
  std::future<void> draw_prefetch_;   // helper generating the next block into pre_u_/pre_xy_
  std::vector<float> pre_u_, pre_xy_;
  void  syncDraws();
  std::vector<float> draw_u_;   // u_k for k >= draw_base_
  std::vector<float> draw_xy_;  // 2 per draw
  size_t draw_base_ = 0;        // stream index of draw_u_[0]
  size_t draw_next_ = 0;        // next unconsumed stream index

  void nodeIndexInsert(trgStruct& g, Node* n);
  void nodeIndexReset(trgStruct& g);
  void ensureGrid(trgStruct& g);
  void ensureGridBuilt(trgStruct& g);
  void ensureTree(trgStruct& g);
  std::atomic<int64_t> us_tree_parts_[3] = {};  // ensureTree: host split / device build / adopt
  Node* nearestNode(trgStruct& g, float x, float y);
  Node* resolveNearestTie(trgStruct& g, float qx, float qy, float d2min);
  void rangeNodesOrdered(trgStruct& g, float x, float y, float r, std::vector<Node*>& out);
  int  countNodesInRange(trgStruct& g, float x, float y, float r);
  void buildMapIndex(trgStruct& g, const float* xyz, int64_t n, int stride, bool device);
  trgb_map* requireMap(trgStruct& g, const char* who);
  void invalidateDeviceGraph();
  void ensureDeviceGraph();
  Node* newNode(int id, Eigen::Vector2f& p, float z, NodeState s);
  Edge* newEdge(int dst, float w, float d);
  void  rewindPools();
  size_t node_used_ = 0, edge_used_ = 0;  // pool cursors
  void setGoalUnlocked(Eigen::Vector3f& goal);
  std::pair<Node*, bool> snapGoal(trgStruct& g, const Eigen::Vector3f& goal);
  std::mutex iter_rank_mx_;
  void runExpansion(const std::vector<Node*>& roots, trgStruct& g);
  // initGraph fast path (trg_device_build.cpp): expandGraph(0) as a device-resident BFS (K9) followed
  // by cleanGraph(false), materialised straight into the cleaned containers. false = not applicable
  // (step-3 wiring on, map too dense, capacity): the caller runs the host-driven path instead.
  bool buildGraphOnDevice(trgStruct& g);
  void destroyExpander();
  std::future<void> graph_disposal_;  // helper freeing the previous build's node map
  trgb_expander*  expander_     = nullptr;
  const trgb_map* expander_map_ = nullptr;

  std::unique_ptr<trg_b200::DeviceSession> dev_;
  // K7 search graph of the global graph: built on the device by the device build, else uploaded by ensureDeviceGraph
  trgb_graph* dev_graph_ = nullptr;
  int64_t dev_graph_relaxed_ = 0;  // edges relaxed by the handle so far (already booked in stat_)
  std::vector<Node*> dev_graph_nodes_;  // row -> node of the uploaded CSR
  std::vector<int64_t> built_row_;      // first edge-pool slot of every node of the last device build (+ total); empty = graph changed since
  std::deque<Node> node_pool_;
  std::deque<Edge> edge_pool_;
  std::unordered_map<std::string, double>  secs_;
  std::unordered_map<std::string, int64_t> stat_;
  std::vector<int32_t> last_path_ids_;
  double  us_draws_ = 0;
  std::atomic<int64_t> n_node_ties_{0};

 public:
  // [+] tuning knobs of the wavefront scheduler (defaults are fine; exposed for benchmarks)
  struct Tuning {
    int chunk_nodes = 4096;   // pops evaluated per batch
    int window = 128;         // sampling-window draws per node (<= 256)
    int lookahead = 768;      // queued pops needed before the next batch is fed ahead of the current commit
    float map_cell_scale = 0.67f;  // map index cell = map_cell_scale * robot_size (0.34..1.0 measured: 0.67-1.0 best)
    float table_cell_scale = 2.0f; // cell of the host table of just-created nodes = table_cell_scale * robot_size
    int parallel_min_nodes = 50000;  // graphs larger than this use helper threads in cleanGraph / CSR export
    bool overlap = true;      // run the device phases of batch k+1 on a helper thread while batch k commits
    bool split_commit = false;  // apply edge-list operations on a second thread while the first decides (measured slower on a 10 M-point map: cross-core traffic on the adjacency lists; kept for experiments)
    bool device_expand = true;  // initGraph: run the whole BFS, decisions included, on the device (K9) when applicable
    int  expand_steps = 8;      // device BFS: steps queued per status poll (two polls in flight)
    int  expand_window_words = 4;  // device BFS: sampling window per pop = 64 * words draws (longer exact chains per step)
    int  expand_max_pops = 4096;   // device BFS: pops per step at most
  } tuning_;
};

#endif  // TRG_PLANNER_B200_HOST_TRG_H_
