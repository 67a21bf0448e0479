// b2g_runtime.h — host-side management of a batch: allocation, (de)serialisation between the
// C-ABI snapshot records (include/b2gpu.h) and the blocked world-minor device layout, and the
// per-step launch sequence.  Compiled by nvcc for the product library (libb2gpu.so); the same
// text is compiled by g++ with -DB2G_HOSTSIM into the test-only host simulator (tests/hostsim),
// where a "launch" is a plain loop over thread ids.  The product never links the simulator.
#pragma once
#include <string>
#include <vector>

#include "b2g_step.h"
#include "b2g_levels.h"
#include "b2g_query.h"
#include "b2g_island_layout.h"

namespace b2g {

// stages of one step, in launch order (profiling / ncu names)
enum {
  STAGE_PRE = 0, STAGE_COLLIDE, STAGE_ISLAND, STAGE_INTEGRATE, STAGE_SOLVER_INIT, STAGE_VELOCITY, STAGE_POST_VELOCITY,
  STAGE_POSITION, STAGE_FINALIZE, STAGE_SLEEP, STAGE_SYNC_FIXTURES, STAGE_TREE_PAIRS, STAGE_BODY_END, STAGE_OTHER, STAGE_COUNT
};
struct ProfSpan {
  int stage;
  void *e0, *e1;  // cudaEvent_t
};
struct Ctx {
  int device = 0;
  void* stream = nullptr;  // cudaStream_t
  bool own_stream = false;
  long long launches = 0;
  bool profiling = false;  // record a CUDA event pair around every launch
  std::vector<void*> ev_free;
  std::vector<ProfSpan> ev_pending;
  double stage_ms[STAGE_COUNT] = {0};
  long long stage_launches[STAGE_COUNT] = {0};
};
int ctx_collect_profile(Ctx* ctx);

// One world in compact SoA form (host): the unit of upload/download.
struct WorldImage {
  std::vector<int> ws, b_flags, n_moved, move_buf, c_flags, b_chead;
  std::vector<float4> b_xf, b_pos, b_pos0, b_vel, b_mass, b_force, b_misc, n_aabb, p_aabb, c_mat, c_m0, c_m1, c_m2, j_s0, j_s1;
  std::vector<int4> n_link, c_fix, c_m3;
  std::vector<int2> c_next;
};

struct Topology {  // shared by every world of a batch (host copies kept for validation/download)
  std::vector<b2gpu_body_rec> bodies;  // static fields only are authoritative (type, fixture_head, fixture_count)
  std::vector<b2gpu_fixture_rec> fixtures;
  std::vector<b2gpu_shape_rec> shapes;
  std::vector<b2gpu_proxy_rec> proxies;
  std::vector<int4> proxy_s;
  std::vector<int> sync_order, sync_rank, node_proxy;
  std::vector<b2gpu_joint_rec> joints;  // static fields authoritative (type, bodies, COLLIDE_CONNECTED, anchors, param 0-2 / 0-4)
  std::vector<int> jadj_off, jadj;      // per body: joint edges in the reference's list order (newest first)
};

struct StreamGroup {
  int wb_first = 0, wb_count = 0;
  void* stream = nullptr;   // cudaStream_t
  void* ev_init = nullptr;  // cudaEvent_t: solver set-up of the first step issued (stagger point)
  void* ev_done = nullptr;  // cudaEvent_t: all work of the call issued on this stream
};

struct StepGraph {  // instantiated CUDA graph of one run_steps call signature
  float dt = 0.0f;
  int vi = 0, pi = 0, steps = 0;
  const void* forces = nullptr;
  void* state = nullptr;
  bool compact = false;
  void* exec = nullptr;  // cudaGraphExec_t
  long long launches = 0;
};

struct BatchHost {
  Ctx* ctx = nullptr;
  Batch B = {};     // device pointers + dims
  Topology topo;
  std::vector<void*> allocs;
  int* b_wake = nullptr;
  int* b_chead = nullptr;
  int2* c_next = nullptr;
  int* stack = nullptr;
  float* state_dev = nullptr;    // [n_worlds][NB][8] gather buffer
  float* forces_dev = nullptr;   // [n_worlds][NB][3]
  float* vel_scratch = nullptr;  // [n_worlds][2] staging of batch_set_linear_velocity
  int* dyn_idx = nullptr;        // [n_dyn] body indices of the prototype's dynamic bodies (compact I/O)
  int n_dyn = 0;
  int* status_dev = nullptr;     // reduced WS_STATUS of the batch (StatusK)
  int* status_host = nullptr;    // pinned copy of it
  int last_fetch_status = 0;     // WS_STATUS of the world last fetched by image_fetch
  void* stage_dev = nullptr;     // staging for single-world upload/download
  size_t stage_bytes = 0;
  bool smem_island = false;      // shared-memory island DFS in use (b2g_island_smem.cuh)
  IslandSmemLayout island_layout;
  std::string timeline_path;        // diagnostic: where batch_destroy writes the Gauss-Seidel CTA timeline
  bool smem_solver = false;      // straight-line shared-memory Gauss-Seidel kernels in use (b2g_solver_smem.cuh)
  std::vector<StreamGroup> groups;  // independent pipelines over windows of world blocks (created on first use)
  void* ev_entry = nullptr;         // cudaEvent_t: fork point on the context stream
  std::vector<StepGraph> graphs;    // CUDA graphs per call signature (see run_steps)
  bool use_graphs = true;
  bool stagger_groups = true;       // group g starts behind group g-1's solver set-up (B2GPU_STAGGER=0: all start together)
  int stream_groups = 0;            // 0 = automatic (4 for batches of >= 16 world blocks, 8 for >= 64, 1 for worlds under 96 bodies), 1 = single stream
  bool stepped = false;          // at least one dt > 0 step ran: island arrays are meaningful
  bool pre_step_needed = true;   // some world may carry m_new_contacts / a non-empty move buffer
  long long total_bytes = 0;
  StepParams last_sp{};
  // large-world mode (b2g_large.h): one world, flat stages + scans + sorts, host-driven control flow
  bool large = false;
  int lw_level_min = LW_LEVEL_MIN_DEFAULT;  // islands with at least this many contacts (no joints) are swept level by level by a CTA; <= 0: never
  bool lw_exact_tree = false;    // large-world mode that keeps the replica tree (sequential re-insertion, reference contact order)
  Large L = {};
  int* lw_host = nullptr;        // pinned readback buffer (world scalars, scan totals)
  int lw_cc = 0;                 // contact count after the last large-mode step (host copy)
  bool lw_cc_valid = false;
  void* lw_tmp = nullptr;        // scan / sort temporary storage
  size_t lw_tmp_bytes = 0;
  int lw_edge_bits = 0, lw_body_bits = 0;
  long long lw_keys = 0;         // capacity of the key buffers
  void* query_buf = nullptr;     // device scratch of the world-query calls (grown on demand)
  size_t query_bytes = 0;
};

const char* last_error();
void set_error(const std::string& s);

int batch_create(Ctx* ctx, const b2gpu_snapshot* proto, int n_worlds, const b2gpu_caps* caps, int lane_block, BatchHost** out);
void batch_destroy(BatchHost* b);
int batch_upload_world(BatchHost* b, int world, const b2gpu_snapshot* in);
int batch_reset(BatchHost* b, const b2gpu_snapshot* in);
int batch_status(BatchHost* b, int* out);
int batch_apply_device_forces(BatchHost* b);
int batch_refresh_device_state(BatchHost* b);
int batch_post_solve_events(BatchHost* b, int world, b2gpu_post_solve_event* out, int capacity);
int batch_last_download_status(BatchHost* b);
int batch_snapshot_sizes(BatchHost* b, int world, b2gpu_snapshot_sizes* out);
int batch_download_world(BatchHost* b, int world, b2gpu_snapshot* out);
int batch_step(BatchHost* b, float dt, int vi, int pi, int steps);
int batch_step_host(BatchHost* b, const float* host_forces, float* host_state_out, float dt, int vi, int pi, int steps);
int batch_get_stats(BatchHost* b, int first, int count, b2gpu_step_stats* out);
int batch_step_host_dynamic(BatchHost* b, const float* host_forces, float* host_state_out, float dt, int vi, int pi, int steps);
int batch_dynamic_bodies(BatchHost* b, int* out, int capacity);
int batch_get_body_state(BatchHost* b, float* host_out, int first, int count);
int batch_set_forces(BatchHost* b, const float* host, int first, int count);
int batch_set_linear_velocity(BatchHost* b, int body, const float* host_vxvy, int first, int count);
int batch_set_gravity(BatchHost* b, const float* host_gxgy, int first, int count);
int batch_set_joint_control(BatchHost* b, int joint, int control, const float* host_values, int first, int count);
int batch_ray_cast_closest(BatchHost* b, const float* host_rays, int rays_per_world, b2gpu_ray_hit* host_out);
int batch_query_aabb(BatchHost* b, const float* host_boxes, int n, int max_hits, int* host_counts, int* host_hits);
int batch_query_aabb_per_world(BatchHost* b, const float* host_boxes, int boxes_per_world, int max_hits, int* host_counts, int* host_hits);
int ctx_sync(Ctx* ctx);
int debug_sincos(Ctx* ctx, const float* host_in, float* host_sin, float* host_cos, int n);
long long batch_algorithmic_bytes(BatchHost* b);

}  // namespace b2g

struct b2gpu_ctx {
  b2g::Ctx c;
};
