// ORACLE — TEST INFRASTRUCTURE ONLY.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library, and only as the checker / the timed CPU baseline.
// The product (trg-planner_b200/) never links, loads or calls anything in oracle/.
//
// C facade of the CPU oracle of the reference TRG hot path
// (cpp/trg_planner/core/trg_planner/src/graph/trg.cpp + src/kdtree/kdtree.c). Two implementations
// export it: trg_oracle.cpp (a restatement) and ref_harness.cpp (the reference's own UNMODIFIED
// trg.cpp + kdtree.c compiled against oracle/shim/ into oracle/_ref/libtrg_ref.so). The restatement
// is pinned to the reference build bit for bit (tests/test_oracle_cpu.py) and tests/golden/ are
// outputs of the reference build. What stays PARITY UNPINNED is Eigen alone (absent from the image
// and from /root/reference): JacobiSVD and the float summation order of the 3x3 covariance are
// restated in eigen_restate.h / shim/Eigen/Core for BOTH builds.
//
// The same entry points, with prefix `trg_` instead of `orc_`, are exported by the
// product's host library so tests drive both through one binding.
#ifndef ORACLE_TRG_ORACLE_H_
#define ORACLE_TRG_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

// TRG::TRG ctor arguments, trg.h:51-59 (same order)
typedef struct OrcParams {
  int   is_verbose;
  float expand_dist;
  float robot_size;
  int   sample_num;
  float height_threshold;
  float collision_threshold;
  float update_collision_threshold;
  float safety_factor;
  float goal_tolerance;
} OrcParams;

// stage at which TRG::wireEdge (trg.cpp:254-370) returned, for the pure geometric part
enum {
  ORC_EDGE_OK = 0,         // reached :365, edge created
  ORC_EDGE_SLOPE = 1,      // :272
  ORC_EDGE_COLLISION = 2,  // :285
  ORC_EDGE_EMPTY = 3,      // :305
  ORC_EDGE_FEWPTS = 4      // :327
};

void* orc_create(const OrcParams* p);
void  orc_destroy(void* h);
// reseed TRG::gen_ (trg.h:129; reference seeds from random_device, trg.cpp:20)
void  orc_seed(void* h, uint32_t seed);

// xyz: n packed (x,y,z) float triples
int orc_set_global_map(void* h, const float* xyz, int64_t n);                     // trg.cpp:179
int orc_set_local_map(void* h, float sx, float sy, const float* xyz, int64_t n);  // trg.cpp:195
int orc_init_graph(void* h, int is_pre_map, float sx, float sy, float sz);        // trg.cpp:36 (-1 instead of exit(1))
int orc_update_graph(void* h);                                                    // trg.cpp:456

// type: "global" | "local"
int orc_graph_counts(void* h, const char* type, int64_t* n_nodes, int64_t* n_edges);
// iter_ids: ids in std::unordered_map iteration order (what saveGraph would emit);
// ids_sorted ascending; pos/state/row_ptr follow ids_sorted; col = dst id in edges_ order.
int orc_graph_export(void* h, const char* type, int32_t* iter_ids, int32_t* ids_sorted,
                     float* pos_xyz, int32_t* state, int64_t* row_ptr, int32_t* col,
                     float* weight, float* dist);

// TRG::planSafePath trg.cpp:603. returns 1 found, 0 not found, <0 error.
// path_xyz / node_ids: caller buffers of max_pts entries (start..goal order).
int orc_plan(void* h, float sx, float sy, float gx, float gy, float gz,
             float* path_xyz, int32_t* node_ids, int max_pts, int* n_pts,
             float* direct_dist, float* path_length, float* avg_risk,
             int* goal_known, int64_t* n_expanded);
// TRG::refinePath trg.cpp:692
int orc_refine_path(void* h, const float* in_xyz, int n_in, float* out_xyz, int* n_out);
int orc_check_reached(void* h, float x, float y);                                         // trg.cpp:567
int orc_check_replan(void* h, float x, float y, const float* path_xyz, int n_path);       // trg.cpp:576

// ---- pure functions of (query, static map, params): kernel-level parity ----
int orc_is_collision_batch(void* h, const char* type, const float* xy, int64_t n,
                           float threshold, uint8_t* out);                         // trg.cpp:746
int orc_range_count_batch(void* h, const char* type, const float* xy, int64_t n,
                          float radius, int32_t* out);                             // kdtree.c:479
int orc_nearest_z_batch(void* h, const char* type, const float* xy, int64_t n,
                        float* z_out, int64_t* idx_out, uint8_t* tie_out);         // trg.cpp:244-246
// geometric part of wireEdge (no duplicate check): stage, weight (float pipeline),
// weight64 (double pipeline on the same point set), dist, number of PCA points
int orc_edge_eval_batch(void* h, const char* type, const float* p1_xyz, const float* p2_xyz,
                        int64_t n, uint8_t* stage, float* weight, double* weight64,
                        float* dist, int32_t* npts);
int orc_is_frontier_batch(void* h, const float* xy, int64_t n, uint8_t* out);      // trg.cpp:780

// wall-clock seconds of the last call at the reference's own timer sites
double orc_last_seconds(void* h, const char* what);  // "set_global_map" | "init_graph" | "update_graph" | "plan" | "set_local_map"
int64_t orc_stat(void* h, const char* what);          // "rng_draws" | "collision_calls" | "edge_evals" | "nearest_map" | "nearest_node"

#ifdef __cplusplus
}
#endif
#endif  // ORACLE_TRG_ORACLE_H_
