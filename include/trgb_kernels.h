/* trgb_kernels.h — thin C ABI over the hand-written sm_100a CUDA kernels of the TRG hot path.
 *
 * This is the seam that replaces the reference's per-call C kd-tree API
 *   cpp/trg_planner/core/trg_planner/include/kdtree/kdtree.h:30-115   (kd_create / kd_insert2 /
 *   kd_nearest2 / kd_nearest_range2 / kd_res_*), which trg.cpp drives one point at a time,
 * with batched, pure functions of (queries, static map, params). Plain pointers and sizes,
 * `int` status (0 ok, <0 error; text via trgb_last_error()), no exceptions, no torch types.
 * Caller-owned host buffers, library-owned device memory, one CUDA stream per map handle;
 * a handle may be used by one thread at a time (the reference serialises on TRG::mtx.graph).
 *
 * Two tiers:
 *   *_batch   host buffers in/out, synchronous (what a cgo/ctypes/pybind stub binds);
 *   *_launch  device pointers, asynchronous on the handle's stream (what the C++ host
 *             library trg-planner_b200/host uses inside TRG::expandGraph, and what bench.py
 *             uses for the "inputs already resident in HBM" number).
 * There is no CPU fallback: every entry point fails with TRGB_E_CUDA when no device is usable.
 */
#ifndef TRGB_KERNELS_H_
#define TRGB_KERNELS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { TRGB_OK = 0, TRGB_E_ARG = -1, TRGB_E_CUDA = -2, TRGB_E_NOMEM = -3, TRGB_E_STATE = -4 };

/* stage at which TRG::wireEdge's geometric part stops (trg.cpp:269-329). The slope gate
 * (:269-274, glibc atan2f) is applied on the host side of the boundary. */
enum { TRGB_EDGE_OK = 0, TRGB_EDGE_SLOPE = 1, TRGB_EDGE_COLLISION = 2, TRGB_EDGE_EMPTY = 3, TRGB_EDGE_FEWPTS = 4,
       TRGB_EDGE_SKIPPED = 5 /* not evaluated: see trgb_edge_eval_launch_skip */ };

typedef struct trgb_map trgb_map;     /* device-resident cell index over one point cloud  */
typedef struct trgb_graph trgb_graph; /* device-resident CSR graph + node kd-tree arrays  */

typedef struct TrgbMapInfo {
  int64_t n_points;
  int32_t grid_w, grid_h;
  float   origin_x, origin_y, cell_size;
  int64_t device_bytes;
} TrgbMapInfo;

/* TRG::Param subset the edge kernel needs (trg.h:132-142) */
typedef struct TrgbEdgeParams {
  float robot_size;
  float height_threshold;
  float collision_threshold;
  /* threads per edge in the segment-collision kernel = expected upper bound on the samples of an
   * edge, ceil(dist / (robot_size/2)). 0 = library default (8). Purely a performance hint: the
   * last thread of an edge walks any samples beyond it, so results never depend on the value. */
  int32_t max_edge_samples;
} TrgbEdgeParams;

const char* trgb_last_error(void);
int trgb_device_count(void);
int trgb_set_device(int device);

/* ---- K1: map index build — replaces the kd_insert2 loop of TRG::setGlobalMap / setLocalMap
 *      (trg.cpp:185-188, 203-206; kdtree.c:167-209). Points: n records of `stride_floats`
 *      floats (3 = packed xyz, 4 = pcl::PointXYZ). cell_size <= 0 picks robot_size/2 semantics
 *      left to the caller; must be > 0. */
int  trgb_map_create(trgb_map** out, const float* host_pts, int64_t n, int stride_floats, float cell_size);
int  trgb_map_create_dev(trgb_map** out, const float* dev_pts, int64_t n, int stride_floats, float cell_size);
void trgb_map_destroy(trgb_map* m);
int  trgb_map_info(const trgb_map* m, TrgbMapInfo* info);
/* the indexed cloud in HBM: n records of 4 floats (x, y, z, bit-cast original index), sorted by cell (valid after trgb_map_sync) */
int  trgb_map_points(const trgb_map* m, const float** d_xyzi, int64_t* n);
void* trgb_map_stream(const trgb_map* m); /* cudaStream_t */
int  trgb_map_sync(const trgb_map* m);
/* options: "force_warp_path" (0/1) routes every query launch through the warp-per-item kernels
 * instead of the thread-per-item fast path; "use_staging" (0/1) selects the shared-memory staged
 * sampling-window kernel (off by default: measured 14 % slower than the L1-cached per-thread loads,
 * profiles/README.md). All variants give identical results (tests). */
int  trgb_map_set_option(trgb_map* m, const char* key, int value);

/* ---- tier 1: host-buffer batches (synchronous) ---------------------------------------- */
/* K2 — TRG::isCollision (trg.cpp:746-778) for n query points xy[2n]; out[i] in {0,1} */
int trgb_collision_batch(trgb_map* m, const float* xy, int64_t n, float radius, float height_thr,
                         float ratio_thr, uint8_t* out);
/* kd_nearest_range2 result size (kdtree.c:479-501) */
int trgb_range_count_batch(trgb_map* m, const float* xy, int64_t n, float radius, int32_t* out);
/* K3 — kd_nearest2 on the map tree + payload z (trg.cpp:244-246). tie[i]=1 when another
 * point has the identical float dist^2 (the reference resolves by kd-tree visit order; this
 * library by lowest original point index). */
int trgb_nearest_z_batch(trgb_map* m, const float* xy, int64_t n, float* z, int64_t* idx, uint8_t* tie);
/* K4 — geometric part of TRG::wireEdge (trg.cpp:269-363) incl. the host-side slope gate */
int trgb_edge_eval_batch(trgb_map* m, const float* p1_xyz, const float* p2_xyz, int64_t n,
                         const TrgbEdgeParams* prm, uint8_t* stage, float* weight, float* dist,
                         int32_t* npts);

/* ---- tier 2: device-pointer launches (asynchronous on trgb_map_stream(m)) -------------- */
int trgb_collision_launch(const trgb_map* m, const float* d_xy, int64_t n, float radius,
                          float height_thr, float ratio_thr, uint8_t* d_out);
int trgb_range_count_launch(const trgb_map* m, const float* d_xy, int64_t n, float radius, int32_t* d_out);
/* sampling windows of TRG::expandGraph (trg.cpp:387-403): for node i and j < window, test
 * isCollision(node_xy[i] + draw_xy[first_draw[i] + j]); bit j of d_mask[i] = collision.
 * Each node owns (window+63)/64 consecutive 64-bit words of d_mask, which the caller zeroes;
 * window <= 256. */
int trgb_sample_window_launch(const trgb_map* m, const float* d_node_xy, const int32_t* d_first_draw,
                              const float* d_draw_xy, int64_t n_nodes, int window, float radius,
                              float height_thr, float ratio_thr, unsigned long long* d_mask);
/* same, with the promise that every |draw_xy| offset is at most max_offset (> 0): lets the kernel stage
 * each node's neighbourhood cells in shared memory once and run all its window tests from there */
int trgb_sample_window_launch2(const trgb_map* m, const float* d_node_xy, const int32_t* d_first_draw,
                               const float* d_draw_xy, int64_t n_nodes, int window, float max_offset, float radius,
                               float height_thr, float ratio_thr, unsigned long long* d_mask);
int trgb_nearest_z_launch(const trgb_map* m, const float* d_xy, int64_t n, float* d_z, int64_t* d_idx,
                          uint8_t* d_tie);
/* p1: (x,y,z) start node; p2: (x,y) end point (its z only enters the host-side slope gate).
 * d_stage gets OK/COLLISION/EMPTY/FEWPTS. */
int trgb_edge_eval_launch(const trgb_map* m, const float* d_p1_xyz, const float* d_p2_xy, int64_t n,
                          const TrgbEdgeParams* prm, uint8_t* d_stage, float* d_weight, float* d_dist,
                          int32_t* d_npts);

/* Speculation filter for the wavefront scheduler: items i < n_skip whose sqrt(d_skip_d2[i]) < skip_below
 * are not evaluated (stage = TRGB_EDGE_SKIPPED, z = 0). d_skip_d2 is the squared distance from the
 * sample to its nearest graph node (trgb_nodes_nearest_launch on the same stream): such a sample can
 * only be wired to an existing node (trg.cpp:414-417), never become a new one, so neither its height
 * nor its parent edge will be needed. */
int trgb_edge_eval_launch_skip(const trgb_map* m, const float* d_p1_xyz, const float* d_p2_xy, int64_t n,
                               const TrgbEdgeParams* prm, uint8_t* d_stage, float* d_weight, float* d_dist,
                               int32_t* d_npts, const float* d_skip_d2, int64_t n_skip, float skip_below);
int trgb_nearest_z_launch_skip(const trgb_map* m, const float* d_xy, int64_t n, float* d_z, int64_t* d_idx,
                               uint8_t* d_tie, const float* d_skip_d2, float skip_below);

/* ---- K5: device grid over graph nodes (append-only) — batched kd_nearest2 on the node tree
 *      (trg.cpp:408; kdtree.c:364-417). Indices are append order. All launches are asynchronous on
 *      the given cudaStream_t. Exact float argmin; tie[i]=1 when two nodes share the minimal dist^2
 *      (the caller resolves those with the reference's tree order). idx = -1 on an empty grid. */
typedef struct trgb_nodes trgb_nodes;
int     trgb_nodes_create(trgb_nodes** out, float x0, float y0, float x1, float y1, float cell);
void    trgb_nodes_destroy(trgb_nodes* g);
int     trgb_nodes_reset(trgb_nodes* g, void* stream);
int64_t trgb_nodes_count(const trgb_nodes* g);
int     trgb_nodes_append_launch(trgb_nodes* g, const float* d_xy, int64_t n, void* stream);
int     trgb_nodes_nearest_launch(const trgb_nodes* g, const float* d_xy, int64_t n, int32_t* d_idx, float* d_d2,
                                  uint8_t* d_tie, void* stream);

/* The insertion-order 2-D kd-tree n successive kd_insert2 calls build over the graph nodes (trg.cpp:249,
 * 528-530; kdtree.c:167-194), grown in parallel on the device (one round per tree level). xy: n (x, y) pairs
 * in insertion order; outputs per node: children (-1 = none), parent (-1 = root), split axis. The tree's
 * shape is what TRG::setGoal's choice among several in-range nodes and kd_nearest's tie order depend on. */
int trgb_kdtree_build(const float* xy, int64_t n, int32_t* lo, int32_t* hi, int32_t* parent, uint8_t* axis);

/* ---- K9: device-resident graph expansion — TRG::expandGraph (trg.cpp:372-454) as a BFS that runs on the
 *      GPU including its decisions (trg-planner_b200/csrc/expand.cu). Usable when step 3 of expandGraph
 *      (neighbour wiring, trg.cpp:429) is off, i.e. expand_dist - robot_size >= 0.25 * expand_dist, and
 *      the map is sparse enough for the thread-per-query kernels (TRGB_E_STATE otherwise: the caller
 *      keeps its host-driven path). The host feeds the sampling stream (mt19937 + glibc cosf / sinf live
 *      on the host side of the boundary) and polls a status block; an exact distance tie between nodes or
 *      a slope gate within 3 ulp of its threshold interrupts the engine at that pop, which the caller
 *      then handles itself (reference tie / libm rules) and hands back with trgb_expander_apply_pop. */
typedef struct trgb_expander trgb_expander;
typedef struct TrgbExpandParams {
  float   expand_dist, robot_size, height_threshold, collision_threshold;
  int32_t sample_num;
  float   max_slope;     /* atan2f(height_threshold, robot_size) as the HOST libm computes it (trg.cpp:269) */
  int32_t max_pops;      /* queue pops per step, 32..8192 */
  int32_t window_words;  /* sampling window per pop = 64 * window_words draws, 2..4 */
  int32_t new_state;     /* NodeState of created nodes: 0 Valid (ref_id == 0, trg.cpp:420) or 1 Frontier */
} TrgbExpandParams;
typedef struct TrgbExpandStatus {
  int32_t head, tail, n_nodes, interrupt, interrupt_pop;
  int64_t n_req, pos, draws_end;
  int64_t window_tests, steps, steps_active, rounds, pops, z_ties, redo_pops, samples, created;
  float   mean, var;
} TrgbExpandStatus;
int  trgb_expander_create(trgb_expander** out, const trgb_map* map, const TrgbExpandParams* prm, float x0, float y0,
                          float x1, float y1, int64_t node_capacity);
void trgb_expander_destroy(trgb_expander* e);
/* point an existing engine at a rebuilt map of the same extent (keeps its buffers); TRGB_E_STATE if it does not fit */
int  trgb_expander_rebind(trgb_expander* e, const trgb_map* map, float x0, float y0, float x1, float y1,
                          int64_t node_capacity);
/* reset: the root (node 0, Valid) is the only node and the only queue entry; draws start at stream position draw_pos */
int  trgb_expander_begin(trgb_expander* e, float root_x, float root_y, float root_z, int64_t draw_pos);
/* n more (expand_dist*cosf(angle), expand_dist*sinf(angle)) pairs following the ones pushed so far */
int  trgb_expander_push_draws(trgb_expander* e, const float* xy, int64_t n);
/* queue n_steps steps (asynchronous); pops_hint sizes the launches (a step takes at most that many pops) */
int  trgb_expander_enqueue(trgb_expander* e, int n_steps, int pops_hint);
/* asynchronous copy of the status block into slot 0/1 + wait for it */
int  trgb_expander_snapshot(trgb_expander* e, int slot);
int  trgb_expander_wait(trgb_expander* e, int slot, TrgbExpandStatus* out);
int  trgb_expander_nodes(trgb_expander* e, int64_t from, int64_t to, float* xyz, int8_t* state);
int  trgb_expander_head_pop(trgb_expander* e, int32_t* node_id);
/* result of a pop the caller handled: nodes (x, y, z, state) in creation order, requests in call order
 * (req_b with bit 31 set = parent edge carrying its weight / dist; otherwise wireEdge(a, b) to evaluate) */
int  trgb_expander_apply_pop(trgb_expander* e, int n_new, const float* nodes_xyzs, int n_req, const int32_t* req_a,
                             const int32_t* req_b, const float* req_w, const float* req_d, int64_t new_pos);
/* evaluate the recorded wireEdge requests (one saturated K4 launch), keep the first success of every pair,
 * group by node in request order = the reference's edges_ order. Ids = creation order (root 0). */
int  trgb_expander_finalize(trgb_expander* e, int64_t* n_nodes, int64_t* n_directed_edges);
int  trgb_expander_download(trgb_expander* e, float* xyz, int8_t* state, int64_t* row_ptr, int32_t* col, float* weight,
                            float* dist);
/* Same data as trgb_expander_download, in page-locked memory owned by the engine (valid until the next
 * trgb_expander_finalize / trgb_expander_destroy): xy = 2 floats per node, z, state per node; CSR over engine ids. */
typedef struct TrgbExpandedGraph {
  int64_t n_nodes, n_directed_edges;
  const float* xy;
  const float* z;
  const int8_t* state;
  const int64_t* row_ptr;
  const int32_t* col;
  const float* weight;
  const float* dist;
} TrgbExpandedGraph;
int  trgb_expander_download_view(trgb_expander* e, TrgbExpandedGraph* out);
/* The K7 search graph of the cleaned graph, built on the device from the finalized arrays (nothing crosses
 * PCIe but old2new): old2new[i] = id of engine node i after TRG::cleanGraph (trg.cpp:491-535) or -1 = dropped;
 * n_new = number of kept nodes (new ids are 0 .. n_new-1). The handle is used like one from trgb_graph_upload. */
int trgb_expander_make_graph(trgb_expander* e, const int32_t* old2new, int32_t n_new, trgb_graph** out);

/* ---- K8: voxel-grid centroid filter — the optional down-sampling of the map ingestion,
 *      TRGPlanner::loadPrebuiltMap -> pcl::VoxelGrid (src/planner/trg_planner.cpp:90-94). One output
 *      point per occupied leaf = centroid, in ascending leaf index (PCL's order). Returns TRGB_E_STATE
 *      and passes the cloud through when the leaf grid would overflow 32-bit indices (PCL does the
 *      same with a warning). out_xyz of the host variant must hold 3*n floats. */
int  trgb_voxel_filter(const float* xyz, int64_t n, int stride_floats, float leaf, float* out_xyz, int64_t* n_out);
int  trgb_voxel_filter_dev(const float* d_pts, int64_t n, int stride_floats, float leaf, float** d_out, int64_t* n_out);
void trgb_device_free(void* p);

/* ---- K7: graph upload + batched risk-aware shortest path (TRG::planSafePath, trg.cpp:618-688).
 *      CSR rows = node id 0..n-1, columns in `edges_` order. Start/goal snapping
 *      (TRG::setGoal trg.cpp:537-565, kd_nearest2 :615) is order-dependent host logic and stays
 *      in the C++ host library; this layer takes node ids. */
typedef struct TrgbGraphDesc {
  int32_t n_nodes;
  int64_t n_edges;
  const int64_t* row_ptr;   /* n_nodes+1 */
  const int32_t* col;       /* n_edges */
  const float*   weight;    /* n_edges */
  const float*   dist;      /* n_edges */
  const float*   pos_xyz;   /* 3*n_nodes */
  const int32_t* state;     /* n_nodes: 0 valid, -1 invalid, 1 frontier */
} TrgbGraphDesc;

int  trgb_graph_upload(trgb_graph** out, const TrgbGraphDesc* g);
/* the same with every array of the descriptor resident on the device (stream = where they were produced) */
int  trgb_graph_upload_device(trgb_graph** out, const TrgbGraphDesc* g_on_device, void* stream);
void trgb_graph_destroy(trgb_graph* g);
/* n (start, goal) node-id pairs. Outputs per query: found; cost = float-accumulated
 * g(goal) = sum (sf*w+1)*dist in the reference's evaluation order (trg.cpp:674);
 * path_length = sum dist and avg_risk = sum w / #path nodes accumulated goal->start like
 * trg.cpp:641-659; node id sequences start..goal concatenated in path_ids, path_offsets[n+1].
 * Returns TRGB_E_NOMEM (and the needed size in path_offsets[n]) when path_ids_capacity is
 * too small. */
int trgb_sssp_batch(trgb_graph* g, const int32_t* start_ids, const int32_t* goal_ids, int64_t n,
                    float safety_factor, uint8_t* found, float* cost, float* path_length,
                    float* avg_risk, int64_t* path_offsets, int32_t* path_ids,
                    int64_t path_ids_capacity);

/* edges relaxed (atomicMin on a label) and queries answered over the handle's life: the K7 roofline counts
 * 20 algorithmic bytes per relaxed edge (12 B edge + 4 B label read + 4 B label write) */
int trgb_graph_stats(const trgb_graph* g, int64_t* relaxed_edges, int64_t* queries);

/* ---- profiling: CUDA-event timing of every kernel launched by this library ------------- */
int trgb_prof_enable(int on);
int trgb_prof_reset(void);
/* fills up to cap entries; returns the number of distinct kernels seen. `units` = work items the
 * launches processed in total: map points for the K1 kernels, queries / window tests / edges /
 * path queries for the others (bench.py multiplies by the per-unit algorithmic bytes of DESIGN.md). */
typedef struct TrgbProfEntry { char name[48]; int64_t launches; double total_ms; double units; } TrgbProfEntry;
int trgb_prof_collect(TrgbProfEntry* out, int cap);
int64_t trgb_launch_count(void); /* kernels launched by this library since load / reset */

#ifdef __cplusplus
}
#endif
#endif /* TRGB_KERNELS_H_ */
