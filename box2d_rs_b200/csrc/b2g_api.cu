// b2g_api.cu — extern "C" entry points of include/b2gpu.h (context + batched worlds).
// No exception or C++ type crosses this boundary; every failure is an error code plus
// b2gpu_last_error().  Without a CUDA device every call that needs one returns
// B2GPU_E_NO_DEVICE: there is no CPU fallback in the product library.
#include <string.h>

#include <new>

#include "b2g_runtime.h"
#if !defined(B2G_HOSTSIM)
#include <cuda_runtime.h>
#endif

using namespace b2g;

struct b2gpu_batch {
  BatchHost* h;
  b2gpu_ctx* ctx;
};

#define GUARD_BEGIN try {
#define GUARD_END                                   \
  }                                                 \
  catch (const std::bad_alloc&) {                   \
    set_error("out of host memory");                \
    return B2GPU_E_INVALID;                         \
  }                                                 \
  catch (...) {                                     \
    set_error("unexpected C++ exception");          \
    return B2GPU_E_INVALID;                         \
  }

extern "C" {

int b2gpu_abi_version(void) { return B2GPU_ABI_VERSION; }
const char* b2gpu_last_error(void) { return last_error(); }

int b2gpu_device_count(void) {
#if defined(B2G_HOSTSIM)
  return 1;
#else
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    return B2GPU_E_NO_DEVICE;
  }
  return n;
#endif
}

int b2gpu_init(int device, void* stream, b2gpu_ctx** out) {
  GUARD_BEGIN
  if (!out) { set_error("b2gpu_init: out is NULL"); return B2GPU_E_INVALID; }
  *out = nullptr;
#if !defined(B2G_HOSTSIM)
  int n = b2gpu_device_count();
  if (n <= 0) { if (n == 0) set_error("no CUDA device visible (there is no CPU fallback)"); return B2GPU_E_NO_DEVICE; }
  if (device < 0 || device >= n) { set_error("b2gpu_init: bad device index"); return B2GPU_E_INVALID; }
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { set_error(std::string("cudaSetDevice: ") + cudaGetErrorString(e)); return B2GPU_E_CUDA; }
#endif
  b2gpu_ctx* c = new b2gpu_ctx();
  c->c.device = device;
  c->c.stream = stream;
  c->c.own_stream = false;
#if !defined(B2G_HOSTSIM)
  if (!stream) {
    cudaStream_t s;
    e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; set_error(std::string("cudaStreamCreate: ") + cudaGetErrorString(e)); return B2GPU_E_CUDA; }
    c->c.stream = (void*)s;
    c->c.own_stream = true;
  }
#endif
  *out = c;
  return 0;
  GUARD_END
}
void b2gpu_shutdown(b2gpu_ctx* ctx) {
  if (!ctx) return;
#if !defined(B2G_HOSTSIM)
  if (ctx->c.own_stream && ctx->c.stream) cudaStreamDestroy((cudaStream_t)ctx->c.stream);
#endif
  delete ctx;
}
int b2gpu_sync(b2gpu_ctx* ctx) {
  if (!ctx) { set_error("b2gpu_sync: ctx is NULL"); return B2GPU_E_INVALID; }
  return ctx_sync(&ctx->c);
}
void* b2gpu_stream(b2gpu_ctx* ctx) { return ctx ? ctx->c.stream : nullptr; }
int64_t b2gpu_launch_count(b2gpu_ctx* ctx) { return ctx ? ctx->c.launches : 0; }

int b2gpu_batch_create(b2gpu_ctx* ctx, const b2gpu_snapshot* proto, int n_worlds, const b2gpu_caps* caps, b2gpu_batch** out) {
  GUARD_BEGIN
  if (!ctx || !out) { set_error("b2gpu_batch_create: bad argument"); return B2GPU_E_INVALID; }
  *out = nullptr;
  BatchHost* h = nullptr;
  int lane_block = caps ? caps->reserved[0] : 0;  // 0 = automatic (32 for >= 32 worlds, else 1)
  int rc = batch_create(&ctx->c, proto, n_worlds, caps, lane_block, &h);
  if (rc) return rc;
  b2gpu_batch* b = new b2gpu_batch();
  b->h = h;
  b->ctx = ctx;
  *out = b;
  return 0;
  GUARD_END
}
void b2gpu_batch_destroy(b2gpu_batch* b) {
  if (!b) return;
  batch_destroy(b->h);
  delete b;
}
int b2gpu_batch_world_count(b2gpu_batch* b) { return b ? b->h->B.n_worlds : B2GPU_E_INVALID; }
int b2gpu_batch_step(b2gpu_batch* b, float dt, int vi, int pi, int steps) {
  GUARD_BEGIN
  if (!b) { set_error("b2gpu_batch_step: batch is NULL"); return B2GPU_E_INVALID; }
  return batch_step(b->h, dt, vi, pi, steps);
  GUARD_END
}
int b2gpu_batch_query_aabb(b2gpu_batch* b, const float* aabbs, int boxes_per_world, int max_hits, int32_t* counts, int32_t* hits) {
  GUARD_BEGIN
  if (!b) { set_error("b2gpu_batch_query_aabb: batch is NULL"); return B2GPU_E_INVALID; }
  return batch_query_aabb_per_world(b->h, aabbs, boxes_per_world, max_hits, counts, hits);
  GUARD_END
}
int b2gpu_batch_ray_cast_closest(b2gpu_batch* b, const float* p1p2, int rays_per_world, b2gpu_ray_hit* out) {
  GUARD_BEGIN
  if (!b) { set_error("b2gpu_batch_ray_cast_closest: batch is NULL"); return B2GPU_E_INVALID; }
  return batch_ray_cast_closest(b->h, p1p2, rays_per_world, out);
  GUARD_END
}
int b2gpu_batch_upload_world(b2gpu_batch* b, int world, const b2gpu_snapshot* in) {
  GUARD_BEGIN
  if (!b) { set_error("batch is NULL"); return B2GPU_E_INVALID; }
  return batch_upload_world(b->h, world, in);
  GUARD_END
}
int b2gpu_batch_snapshot_sizes(b2gpu_batch* b, int world, b2gpu_snapshot_sizes* out) {
  GUARD_BEGIN
  if (!b) { set_error("batch is NULL"); return B2GPU_E_INVALID; }
  return batch_snapshot_sizes(b->h, world, out);
  GUARD_END
}
int b2gpu_batch_download_world(b2gpu_batch* b, int world, b2gpu_snapshot* out) {
  GUARD_BEGIN
  if (!b) { set_error("batch is NULL"); return B2GPU_E_INVALID; }
  int rc = batch_download_world(b->h, world, out);
  if (rc) return rc;
  return batch_last_download_status(b->h);  // buffers are filled; a failed world reports its device status
  GUARD_END
}
int b2gpu_batch_post_solve_events(b2gpu_batch* b, int world, b2gpu_post_solve_event* out, int capacity) {
  GUARD_BEGIN
  if (!b) { set_error("batch is NULL"); return B2GPU_E_INVALID; }
  return batch_post_solve_events(b->h, world, out, capacity);
  GUARD_END
}
int b2gpu_batch_reset(b2gpu_batch* b, const b2gpu_snapshot* in) {
  GUARD_BEGIN
  if (!b) { set_error("batch is NULL"); return B2GPU_E_INVALID; }
  return batch_reset(b->h, in);
  GUARD_END
}
int b2gpu_batch_status(b2gpu_batch* b) {
  GUARD_BEGIN
  if (!b) { set_error("batch is NULL"); return B2GPU_E_INVALID; }
  int st = 0;
  int rc = batch_status(b->h, &st);
  if (rc) return rc;
  if (st) set_error("a world of the batch failed on the device (see b2gpu_step_stats.status per world)");
  return st;
  GUARD_END
}
int b2gpu_batch_set_level_threshold(b2gpu_batch* b, int contacts) {
  if (!b) { set_error("batch is NULL"); return B2GPU_E_INVALID; }
  b->h->lw_level_min = contacts == 0 ? (int)LW_LEVEL_MIN_DEFAULT : contacts;
  return 0;
}
int b2gpu_batch_get_stats(b2gpu_batch* b, int first, int count, b2gpu_step_stats* out) {
  GUARD_BEGIN
  if (!b) { set_error("batch is NULL"); return B2GPU_E_INVALID; }
  return batch_get_stats(b->h, first, count, out);
  GUARD_END
}
int b2gpu_batch_set_forces(b2gpu_batch* b, const float* host, int first, int count) {
  GUARD_BEGIN
  if (!b) { set_error("batch is NULL"); return B2GPU_E_INVALID; }
  return batch_set_forces(b->h, host, first, count);
  GUARD_END
}
int b2gpu_batch_set_linear_velocity(b2gpu_batch* b, int body, const float* host_vxvy, int first, int count) {
  GUARD_BEGIN
  if (!b) { set_error("batch is NULL"); return B2GPU_E_INVALID; }
  return batch_set_linear_velocity(b->h, body, host_vxvy, first, count);
  GUARD_END
}
int b2gpu_batch_set_gravity(b2gpu_batch* b, const float* host_gxgy, int first, int count) {
  GUARD_BEGIN
  if (!b) { set_error("batch is NULL"); return B2GPU_E_INVALID; }
  return batch_set_gravity(b->h, host_gxgy, first, count);
  GUARD_END
}
int b2gpu_batch_set_joint_control(b2gpu_batch* b, int joint, int control, const float* host_values, int first, int count) {
  GUARD_BEGIN
  if (!b) { set_error("batch is NULL"); return B2GPU_E_INVALID; }
  return batch_set_joint_control(b->h, joint, control, host_values, first, count);
  GUARD_END
}
int b2gpu_batch_get_body_state(b2gpu_batch* b, float* host_out, int first, int count) {
  GUARD_BEGIN
  if (!b) { set_error("batch is NULL"); return B2GPU_E_INVALID; }
  return batch_get_body_state(b->h, host_out, first, count);
  GUARD_END
}

void* b2gpu_batch_body_state_device(b2gpu_batch* b, int64_t* bytes) {
  if (!b) return nullptr;
  if (bytes) *bytes = (int64_t)b->h->B.n_worlds * b->h->B.NB * 8 * 4;
  return b->h->state_dev;
}
void* b2gpu_batch_forces_device(b2gpu_batch* b, int64_t* bytes) {
  if (!b) return nullptr;
  if (bytes) *bytes = (int64_t)b->h->B.n_worlds * b->h->B.NB * 3 * 4;
  return b->h->forces_dev;
}
int b2gpu_batch_apply_device_forces(b2gpu_batch* b) {
  GUARD_BEGIN
  if (!b) { set_error("batch is NULL"); return B2GPU_E_INVALID; }
  return batch_apply_device_forces(b->h);
  GUARD_END
}
int b2gpu_batch_refresh_device_state(b2gpu_batch* b) {
  GUARD_BEGIN
  if (!b) { set_error("batch is NULL"); return B2GPU_E_INVALID; }
  return batch_refresh_device_state(b->h);
  GUARD_END
}
int b2gpu_batch_step_host(b2gpu_batch* b, const float* host_forces, float* host_state_out, float dt, int vi, int pi, int steps) {
  GUARD_BEGIN
  if (!b) { set_error("batch is NULL"); return B2GPU_E_INVALID; }
  return batch_step_host(b->h, host_forces, host_state_out, dt, vi, pi, steps);
  GUARD_END
}
int b2gpu_batch_step_host_dynamic(b2gpu_batch* b, const float* host_forces, float* host_state_out, float dt, int vi, int pi, int steps) {
  GUARD_BEGIN
  if (!b) { set_error("batch is NULL"); return B2GPU_E_INVALID; }
  return batch_step_host_dynamic(b->h, host_forces, host_state_out, dt, vi, pi, steps);
  GUARD_END
}
int b2gpu_batch_dynamic_bodies(b2gpu_batch* b, int32_t* out, int capacity) {
  GUARD_BEGIN
  if (!b) { set_error("batch is NULL"); return B2GPU_E_INVALID; }
  return batch_dynamic_bodies(b->h, out, capacity);
  GUARD_END
}
int64_t b2gpu_batch_algorithmic_bytes(b2gpu_batch* b) { return b ? batch_algorithmic_bytes(b->h) : -1; }
static const char* k_stage_names[STAGE_COUNT] = {"pre_step_pairs", "collide", "island", "integrate", "solver_init", "velocity",
                                                  "post_velocity", "position", "finalize", "sleep", "sync_fixtures",
                                                  "tree_pairs", "body_end", "other"};
int b2gpu_stage_count(void) { return STAGE_COUNT; }
const char* b2gpu_stage_name(int stage) { return stage >= 0 && stage < STAGE_COUNT ? k_stage_names[stage] : ""; }
int b2gpu_set_profiling(b2gpu_ctx* ctx, int on) {
  if (!ctx) { set_error("ctx is NULL"); return B2GPU_E_INVALID; }
  int rc = ctx_collect_profile(&ctx->c);
  ctx->c.profiling = on != 0;
  for (int i = 0; i < STAGE_COUNT; ++i) { ctx->c.stage_ms[i] = 0.0; ctx->c.stage_launches[i] = 0; }
  return rc;
}
int b2gpu_get_stage_times(b2gpu_ctx* ctx, double* ms_out, int64_t* launches_out, int n) {
  if (!ctx || !ms_out || n < 0) { set_error("get_stage_times: bad argument"); return B2GPU_E_INVALID; }
  int rc = ctx_collect_profile(&ctx->c);
  for (int i = 0; i < n && i < STAGE_COUNT; ++i) {
    ms_out[i] = ctx->c.stage_ms[i];
    if (launches_out) launches_out[i] = ctx->c.stage_launches[i];
  }
  return rc;
}
int b2gpu_debug_sincos(b2gpu_ctx* ctx, const float* host_in, float* host_sin, float* host_cos, int n) {
  GUARD_BEGIN
  if (!ctx) { set_error("ctx is NULL"); return B2GPU_E_INVALID; }
  return debug_sincos(&ctx->c, host_in, host_sin, host_cos, n);
  GUARD_END
}

}  // extern "C"
