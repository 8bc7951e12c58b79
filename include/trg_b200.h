/* trg_b200.h — C facade of the B200-native TRG host library (libtrg_b200.so).
 *
 * This is the drop-in boundary for front ends that are not C++: each entry point maps 1:1 to a
 * public method of the reference class `TRG`
 *   cpp/trg_planner/core/trg_planner/include/graph/trg.h:50-98
 * (C++ callers — TRGPlanner, the pybind module, the ROS nodes — include
 * trg-planner_b200/host/trg.h instead, which is source-compatible with that header).
 * Plain pointers and sizes; int status: >= 0 ok, < 0 error (text via trg_last_error()).
 * The CPU oracle exports the same functions with the prefix `orc_` (oracle/trg_oracle.h) so the
 * parity tests drive both through one binding. There is no CPU fallback in this library.
 */
#ifndef TRG_B200_H_
#define TRG_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* TRG::TRG constructor arguments, trg.h:51-59 (same order) */
typedef struct TrgParams {
  int   is_verbose;
  float expand_dist;
  float robot_size;
  int   sample_num;
  float height_threshold;
  float collision_threshold;
  float update_collision_threshold;
  float safety_factor;
  float goal_tolerance;
} TrgParams;

const char* trg_last_error(void);

void* trg_create(const TrgParams* p);                                   /* TRG::TRG        trg.cpp:11  */
void  trg_destroy(void* h);
void  trg_seed(void* h, uint32_t seed);                                 /* reseeds TRG::gen_ (trg.cpp:20 uses random_device) */

int trg_set_global_map(void* h, const float* xyz, int64_t n);           /* TRG::setGlobalMap trg.cpp:179 */
int trg_set_global_map_dev(void* h, const float* dev_xyz, int64_t n, int stride_floats); /* cloud already in HBM */
int trg_set_local_map(void* h, float sx, float sy, const float* xyz, int64_t n);         /* TRG::setLocalMap trg.cpp:195 */
int trg_init_graph(void* h, int is_pre_map, float sx, float sy, float sz);               /* TRG::initGraph   trg.cpp:36  */
int trg_update_graph(void* h);                                                           /* TRG::updateGraph trg.cpp:456 */

int trg_graph_counts(void* h, const char* type, int64_t* n_nodes, int64_t* n_edges);
/* iter_ids: ids in std::unordered_map iteration order (what saveGraph emits, trg.cpp:141);
 * ids_sorted ascending; pos/state/row_ptr follow ids_sorted; col = dst id in edges_ order. */
int trg_graph_export(void* h, const char* type, int32_t* iter_ids, int32_t* ids_sorted, float* pos_xyz,
                     int32_t* state, int64_t* row_ptr, int32_t* col, float* weight, float* dist);
int trg_save_graph(void* h, const char* path);                          /* TRG::saveGraph         trg.cpp:130 */
int trg_load_graph(void* h, const char* path);                          /* TRG::loadPrebuiltGraph trg.cpp:66  */

/* TRG::planSafePath trg.cpp:603. returns 1 found, 0 not found, <0 error. */
int trg_plan(void* h, float sx, float sy, float gx, float gy, float gz, float* path_xyz, int32_t* node_ids,
             int max_pts, int* n_pts, float* direct_dist, float* path_length, float* avg_risk, int* goal_known,
             int64_t* n_expanded);
/* batched planSafePath: queries = n rows (sx, sy, gx, gy, gz) */
int trg_plan_batch(void* h, const float* queries, int64_t n, uint8_t* found, float* cost, float* path_length,
                   float* avg_risk, float* direct_dist, uint8_t* goal_known, int64_t* path_offsets,
                   int32_t* path_ids, int64_t path_ids_capacity);
int trg_refine_path(void* h, const float* in_xyz, int n_in, float* out_xyz, int* n_out);   /* TRG::refinePath trg.cpp:692 */
int trg_check_reached(void* h, float x, float y);                                          /* TRG::checkReadched trg.cpp:567 */
int trg_check_replan(void* h, float x, float y, const float* path_xyz, int n_path);        /* TRG::checkReplan   trg.cpp:576 */

/* ---- the step before the path: configuration and map ingestion ------------------------------
 * TRGPlanner::setParams (src/planner/trg_planner.cpp:103-129): reads the config yaml files; `out` receives the
 * nine TRG constructor arguments, the map settings come back through the other pointers (any may
 * be NULL; strings are copied into caller buffers of `path_cap` bytes). */
int trg_load_params_yaml(const char* config_path, TrgParams* out, int* is_prebuilt_map, char* prebuilt_map_path,
                         int path_cap, int* is_voxelize, float* voxel_size, int* is_update);
/* pcl::io::loadPCDFile (trg_planner.cpp:85): number of points in *n; xyz may be NULL to query the size */
int trg_load_pcd(const char* path, float* xyz, int64_t cap_points, int64_t* n);
int trg_save_pcd(const char* path, const float* xyz, int64_t n, int binary);
/* TRGPlanner::loadPrebuiltMap (trg_planner.cpp:76-101): PCD -> optional voxel filter (device) ->
 * TRG::setGlobalMap. n_raw / n_map receive the point counts before / after the filter. */
int trg_load_prebuilt_map(void* h, const char* pcd_path, int is_voxelize, float voxel_size, int64_t* n_raw, int64_t* n_map);

/* pure batched evaluations */
int trg_is_collision_batch(void* h, const char* type, const float* xy, int64_t n, float threshold, uint8_t* out); /* trg.cpp:746 */
int trg_range_count_batch(void* h, const char* type, const float* xy, int64_t n, float radius, int32_t* out);
int trg_nearest_z_batch(void* h, const char* type, const float* xy, int64_t n, float* z_out, int64_t* idx_out,
                        uint8_t* tie_out);
int trg_edge_eval_batch(void* h, const char* type, const float* p1_xyz, const float* p2_xyz, int64_t n,
                        uint8_t* stage, float* weight, double* weight64_unused, float* dist, int32_t* npts);
int trg_is_frontier_batch(void* h, const float* xy, int64_t n, uint8_t* out);               /* trg.cpp:780 */

/* the device map index behind a TRG ("global" | "local") as a trgb_map* for the tier-2 launches of
 * include/trgb_kernels.h (NULL when no map is loaded); owned by the TRG */
void* trg_device_map(void* h, const char* type);

double  trg_last_seconds(void* h, const char* what);
int64_t trg_stat(void* h, const char* what);
int     trg_set_tuning(void* h, const char* key, double value);  /* "chunk_nodes" | "window" | "lookahead" | "map_cell_scale" | "table_cell_scale" | "overlap" | "split_commit" */

#ifdef __cplusplus
}
#endif
#endif /* TRG_B200_H_ */
