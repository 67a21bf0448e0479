// b2g_runtime.cu — batch management and the step launch sequence (see b2g_runtime.h).
#include "b2g_runtime.h"
#include <cstdlib>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#if !defined(B2G_HOSTSIM)
#include <cuda_runtime.h>
#include <cub/cub.cuh>

#include "b2g_island_smem.cuh"
#include "b2g_solver_smem.cuh"
#endif

namespace b2g {

static thread_local std::string g_error;
const char* last_error() { return g_error.c_str(); }
void set_error(const std::string& s) { g_error = s; }

#define RC(x) do { int rc_ = (x); if (rc_ != 0) return rc_; } while (0)

static int status_error(int st);
static int large_alloc(BatchHost* bh);
static int lw_build_lbvh(BatchHost* bh, int stage);
static void large_free(BatchHost* bh);

// ------------------------------------------------------------------ device abstraction
#if defined(B2G_HOSTSIM)
static int dev_alloc(void** p, size_t bytes) { *p = calloc(1, bytes ? bytes : 1); return *p ? 0 : B2GPU_E_CUDA; }
static void dev_free(void* p) { free(p); }
static int dev_h2d(Ctx*, void* d, const void* h, size_t n) { memcpy(d, h, n); return 0; }
static int dev_d2h(Ctx*, void* h, const void* d, size_t n) { memcpy(h, d, n); return 0; }
static int dev_zero(Ctx*, void* d, size_t n) { memset(d, 0, n); return 0; }
int ctx_sync(Ctx*) { return 0; }
template <class K> static int launch(Ctx* ctx, const K& k, int n, int /*block*/, int stage = STAGE_OTHER) {
  for (int t = 0; t < n; ++t) k(t);
  ctx->launches++;
  ctx->stage_launches[stage] += 1;
  return 0;
}
template <class K> static int launch_occ(Ctx* ctx, const K& k, int n, int stage) { return launch(ctx, k, n, 128, stage); }
// CTA functors k(cta, thread, threads): one thread per CTA here, so a barrier is the end of a loop
template <class K> static int launch_cta(Ctx* ctx, const K& k, int n_cta, int /*threads*/, int stage, size_t /*smem_bytes*/ = 0) {
  for (int g = 0; g < n_cta; ++g) k(g, 0, 1);
  ctx->launches++;
  ctx->stage_launches[stage] += 1;
  return 0;
}
int ctx_collect_profile(Ctx*) { return 0; }
#else
static int cuda_fail(cudaError_t e, const char* what) {
  set_error(std::string(what) + ": " + cudaGetErrorString(e));
  return B2GPU_E_CUDA;
}
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return cuda_fail(e_, #x); } while (0)
static int dev_alloc(void** p, size_t bytes) {
  CU(cudaMalloc(p, bytes ? bytes : 16));
  // zero-fill on the default stream and finish it here: the context stream may be a non-blocking one
  CU(cudaMemset(*p, 0, bytes ? bytes : 16));
  CU(cudaStreamSynchronize(0));
  return 0;
}
static void dev_free(void* p) { cudaFree(p); }
static int dev_h2d(Ctx* c, void* d, const void* h, size_t n) {
  CU(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, (cudaStream_t)c->stream));
  CU(cudaStreamSynchronize((cudaStream_t)c->stream));
  return 0;
}
static int dev_d2h(Ctx* c, void* h, const void* d, size_t n) {
  CU(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, (cudaStream_t)c->stream));
  CU(cudaStreamSynchronize((cudaStream_t)c->stream));
  return 0;
}
static int dev_zero(Ctx* c, void* d, size_t n) {
  CU(cudaMemsetAsync(d, 0, n, (cudaStream_t)c->stream));
  return 0;
}
int ctx_sync(Ctx* c) {
  CU(cudaStreamSynchronize((cudaStream_t)c->stream));
  return 0;
}
template <class K> __global__ void stage_kernel(const K k, int n) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) k(t);
}
static int prof_event(Ctx* ctx, cudaEvent_t* ev) {
  if (!ctx->ev_free.empty()) { *ev = (cudaEvent_t)ctx->ev_free.back(); ctx->ev_free.pop_back(); }
  else CU(cudaEventCreate(ev));
  CU(cudaEventRecord(*ev, (cudaStream_t)ctx->stream));
  return 0;
}
struct LaunchScope {  // brackets one launch with profiling events and checks the launch status
  Ctx* ctx;
  int stage;
  cudaEvent_t e0 = nullptr;
  int begin() { return ctx->profiling ? prof_event(ctx, &e0) : 0; }
  int end() {
    if (ctx->profiling) {
      cudaEvent_t e1 = nullptr;
      RC(prof_event(ctx, &e1));
      ProfSpan sp = {stage, (void*)e0, (void*)e1};
      ctx->ev_pending.push_back(sp);
    }
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
    return 0;
  }
};
template <class K> static int launch(Ctx* ctx, const K& k, int n, int block, int stage = STAGE_OTHER) {
  if (n <= 0) return 0;
  int grid = (n + block - 1) / block;
  LaunchScope ls = {ctx, stage};
  RC(ls.begin());
  stage_kernel<K><<<grid, block, 0, (cudaStream_t)ctx->stream>>>(k, n);
  return ls.end();
}
// Latency-bound flat stages (narrowphase): cap registers so that 8 blocks of 128 threads fit an SM.
template <class K> __global__ void __launch_bounds__(128, 8) stage_kernel_occ(const K k, int n) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) k(t);
}
template <class K> static int launch_occ(Ctx* ctx, const K& k, int n, int stage) {
  if (n <= 0) return 0;
  LaunchScope ls = {ctx, stage};
  RC(ls.begin());
  stage_kernel_occ<K><<<(n + 127) / 128, 128, 0, (cudaStream_t)ctx->stream>>>(k, n);
  return ls.end();
}
// CTA functors k(cta, thread, threads) that synchronise their threads (b2g_levels.h)
template <class K> __global__ void cta_kernel(const K k) { k((int)blockIdx.x, (int)threadIdx.x, (int)blockDim.x); }
template <class K> static int launch_cta(Ctx* ctx, const K& k, int n_cta, int threads, int stage, size_t smem_bytes = 0) {
  if (n_cta <= 0) return 0;
  if (smem_bytes > 48 * 1024) {  // opt in to the large dynamic shared memory once per kernel
    static size_t allowed = 0;
    if (smem_bytes > allowed) {
      CU(cudaFuncSetAttribute(cta_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
      allowed = smem_bytes;
    }
  }
  LaunchScope ls = {ctx, stage};
  RC(ls.begin());
  cta_kernel<K><<<n_cta, threads, smem_bytes, (cudaStream_t)ctx->stream>>>(k);
  return ls.end();
}
int ctx_collect_profile(Ctx* ctx) {
  CU(cudaStreamSynchronize((cudaStream_t)ctx->stream));
  for (const ProfSpan& sp : ctx->ev_pending) {
    float ms = 0.0f;
    CU(cudaEventElapsedTime(&ms, (cudaEvent_t)sp.e0, (cudaEvent_t)sp.e1));
    ctx->stage_ms[sp.stage] += ms;
    ctx->stage_launches[sp.stage] += 1;
    ctx->ev_free.push_back(sp.e0);
    ctx->ev_free.push_back(sp.e1);
  }
  ctx->ev_pending.clear();
  return 0;
}
#endif


// ------------------------------------------------------------------ layout movers (4-byte words)
struct MoveK {
  int* dev;        // device array (blocked world-minor), elements of E words
  int* compact;    // compact one-world array [N][E]
  int N, E, mode;  // mode 0: compact -> world `w`; 1: world `w` -> compact; 2: compact -> every world
  int LB, lb_shift, n_wblocks, w;
  B2G_HD void operator()(int tid) const {
    const int k = tid % E, e = tid / E;
    if (mode == 2) {
      const int i = (e >> lb_shift) % N;
      dev[tid] = compact[i * E + k];
    } else {
      const int wb = w >> lb_shift, wl = w & (LB - 1);
      const int idx = (wb * N + e) * LB + wl;
      if (mode == 0) dev[idx * E + k] = compact[e * E + k];
      else compact[e * E + k] = dev[idx * E + k];
    }
  }
};

struct ArrRef {
  void* dev;
  void* host;
  int E, N;
};
static std::vector<ArrRef> array_table(BatchHost* bh, WorldImage& im) {
  Batch& B = bh->B;
  std::vector<ArrRef> t;
  auto add = [&](void* dev, void* host, int bytes, int N) { ArrRef r = {dev, host, bytes / 4, N}; t.push_back(r); };
  add(B.ws, im.ws.data(), 4, WS_COUNT);
  add(B.b_flags, im.b_flags.data(), 4, B.NB);
  add(B.b_xf, im.b_xf.data(), 16, B.NB);
  add(B.b_pos, im.b_pos.data(), 16, B.NB);
  add(B.b_pos0, im.b_pos0.data(), 16, B.NB);
  add(B.b_vel, im.b_vel.data(), 16, B.NB);
  add(B.b_mass, im.b_mass.data(), 16, B.NB);
  add(B.b_force, im.b_force.data(), 16, B.NB);
  add(B.b_misc, im.b_misc.data(), 16, B.NB);
  add(B.n_aabb, im.n_aabb.data(), 16, B.NN);
  add(B.n_link, im.n_link.data(), 16, B.NN);
  add(B.n_moved, im.n_moved.data(), 4, B.NN);
  add(B.p_aabb, im.p_aabb.data(), 16, B.NP);
  add(B.move_buf, im.move_buf.data(), 4, B.NMOVE);
  add(B.c_fix, im.c_fix.data(), 16, B.NC);
  add(B.c_flags, im.c_flags.data(), 4, B.NC);
  add(B.c_mat, im.c_mat.data(), 16, B.NC);
  add(B.c_m0, im.c_m0.data(), 16, B.NC);
  add(B.c_m1, im.c_m1.data(), 16, B.NC);
  add(B.c_m2, im.c_m2.data(), 16, B.NC);
  add(B.c_m3, im.c_m3.data(), 16, B.NC);
  add(bh->b_chead, im.b_chead.data(), 4, B.NB);
  add(bh->c_next, im.c_next.data(), 8, B.NC);
  if (B.NJ > 0) {
    add(B.j_s0, im.j_s0.data(), 16, B.NJ);
    add(B.j_s1, im.j_s1.data(), 16, B.NJ);
  }
  return t;
}
static void image_alloc(const Batch& B, WorldImage& im) {
  im.ws.assign(WS_COUNT, 0);
  im.b_flags.assign(B.NB, 0);
  float4 z4 = make_float4(0, 0, 0, 0);
  im.b_xf.assign(B.NB, z4); im.b_pos.assign(B.NB, z4); im.b_pos0.assign(B.NB, z4); im.b_vel.assign(B.NB, z4);
  im.b_mass.assign(B.NB, z4); im.b_force.assign(B.NB, z4); im.b_misc.assign(B.NB, z4);
  im.n_aabb.assign(B.NN, z4);
  im.n_link.assign(B.NN, make_int4(-1, -1, -1, -1));
  im.n_moved.assign(B.NN, 0);
  im.p_aabb.assign(B.NP, z4);
  im.move_buf.assign(B.NMOVE, -1);
  im.c_fix.assign(B.NC, make_int4(0, 0, 0, 0));
  im.c_flags.assign(B.NC, 0);
  im.c_mat.assign(B.NC, z4); im.c_m0.assign(B.NC, z4); im.c_m1.assign(B.NC, z4); im.c_m2.assign(B.NC, z4);
  im.c_m3.assign(B.NC, make_int4(0, 0, 0, 0));
  im.b_chead.assign(B.NB, -1);
  im.c_next.assign(B.NC, make_int2(-1, -1));
  im.j_s0.assign(std::max(B.NJ, 1), z4);
  im.j_s1.assign(std::max(B.NJ, 1), z4);
}

static int fbits(float f) { int i; memcpy(&i, &f, 4); return i; }
static float bitsf(int i) { float f; memcpy(&f, &i, 4); return f; }

// snapshot -> image
static int image_pack(const BatchHost* bh, const b2gpu_snapshot* s, WorldImage& im) {
  const Batch& B = bh->B;
  image_alloc(B, im);
  if (s->n.node_count > B.NN || s->n.contact_count > B.NC || s->n.move_count > B.NMOVE) {
    set_error("snapshot exceeds the batch capacities (tree nodes / contacts / move buffer)");
    return B2GPU_E_CAPACITY;
  }
  const b2gpu_world_rec& w = s->world;
  im.ws[WS_GRAVITY_X] = fbits(w.gravity_x);
  im.ws[WS_GRAVITY_Y] = fbits(w.gravity_y);
  im.ws[WS_INV_DT0] = fbits(w.inv_dt0);
  im.ws[WS_FLAGS] = (int)w.flags;
  im.ws[WS_TREE_ROOT] = w.tree_root;
  im.ws[WS_TREE_FREE] = w.tree_free_list;
  im.ws[WS_TREE_COUNT] = w.tree_node_count;
  im.ws[WS_TREE_CAP] = w.tree_node_capacity;
  im.ws[WS_TREE_INSERTIONS] = w.tree_insertion_count;
  im.ws[WS_PROXY_COUNT] = w.proxy_count;
  im.ws[WS_CONTACT_COUNT] = s->n.contact_count;
  im.ws[WS_MOVE_COUNT] = s->n.move_count;
  im.ws[WS_TOPO_DIRTY] = 1;
  for (int i = 0; i < B.NB; ++i) {
    const b2gpu_body_rec& b = s->bodies[i];
    im.b_flags[i] = (int)(b.flags & 0xffffu) | (b.type << 16);
    im.b_xf[i] = make_float4(b.xf_px, b.xf_py, b.xf_qs, b.xf_qc);
    im.b_pos[i] = make_float4(b.c_x, b.c_y, b.a, b.sleep_time);
    im.b_pos0[i] = make_float4(b.c0_x, b.c0_y, b.a0, 0.0f);
    im.b_vel[i] = make_float4(b.vx, b.vy, b.w, 0.0f);
    im.b_mass[i] = make_float4(b.inv_mass, b.inv_inertia, b.lc_x, b.lc_y);
    im.b_force[i] = make_float4(b.fx, b.fy, b.torque, b.gravity_scale);
    im.b_misc[i] = make_float4(b.mass, b.inertia, b.linear_damping, b.angular_damping);
  }
  for (int i = 0; i < s->n.node_count; ++i) {
    const b2gpu_tree_node_rec& n = s->nodes[i];
    im.n_aabb[i] = make_float4(n.aabb[0], n.aabb[1], n.aabb[2], n.aabb[3]);
    im.n_link[i] = make_int4(n.parent, n.child1, n.child2, n.height);
    im.n_moved[i] = n.moved;
  }
  for (int i = 0; i < B.NP; ++i) {
    const float* a = s->proxies[i].aabb;
    im.p_aabb[i] = make_float4(a[0], a[1], a[2], a[3]);
  }
  for (int i = 0; i < s->n.move_count; ++i) im.move_buf[i] = s->move_buffer[i];
  for (int i = 0; i < B.NJ; ++i) {
    const b2gpu_joint_rec& j = s->joints[i];
    // revolute: param 3 = max_motor_torque, 4 = motor_speed are per world (an RL action); distance joints keep
    // everything static (their slots of j_s1 are unused)
    if (j.type == B2GPU_JOINT_GEAR) {  // one accumulated impulse; impulse[1..6] are static data
      im.j_s0[i] = make_float4(j.impulse[0], 0.0f, 0.0f, 0.0f);
      im.j_s1[i] = make_float4(0.0f, 0.0f, 0.0f, bitsf(0));
      continue;
    }
    im.j_s0[i] = make_float4(j.impulse[0], j.impulse[1], j.impulse[2], j.impulse[3]);
    const bool motorised = j.type == B2GPU_JOINT_REVOLUTE || j.type == B2GPU_JOINT_PRISMATIC || j.type == B2GPU_JOINT_WHEEL ||
                           j.type == B2GPU_JOINT_MOUSE;  // mouse: param 3, 4 = the target
    im.j_s1[i] = make_float4(j.impulse[4], motorised ? j.param[4] : 0.0f, motorised ? j.param[3] : 0.0f,
                             bitsf((int)(j.flags & (B2GPU_JOINT_ENABLE_LIMIT | B2GPU_JOINT_ENABLE_MOTOR))));
  }
  for (int i = 0; i < s->n.contact_count; ++i) {
    const b2gpu_contact_rec& c = s->contacts[i];
    if (c.fixture_a < 0 || c.fixture_a >= B.NF || c.fixture_b < 0 || c.fixture_b >= B.NF) {
      set_error("contact references an unknown fixture");
      return B2GPU_E_INVALID;
    }
    im.c_fix[i] = make_int4(c.fixture_a, c.fixture_b, c.index_a, c.index_b);
    im.c_flags[i] = (int)(c.flags & 0xffu);
    im.c_mat[i] = make_float4(c.friction, c.restitution, c.restitution_threshold, c.tangent_speed);
    const b2gpu_manifold& m = c.manifold;
    im.c_m0[i] = make_float4(m.points[0].lp_x, m.points[0].lp_y, m.points[0].normal_impulse, m.points[0].tangent_impulse);
    im.c_m1[i] = make_float4(m.points[1].lp_x, m.points[1].lp_y, m.points[1].normal_impulse, m.points[1].tangent_impulse);
    im.c_m2[i] = make_float4(m.ln_x, m.ln_y, m.lp_x, m.lp_y);
    im.c_m3[i] = make_int4((int)m.points[0].id, (int)m.points[1].id, m.type, m.point_count);
    // push_front on both bodies' edge lists
    const int ba = bh->topo.fixtures[c.fixture_a].body, bb = bh->topo.fixtures[c.fixture_b].body;
    im.c_next[i] = make_int2(im.b_chead[ba], 0);
    im.b_chead[ba] = 2 * i;
    im.c_next[i].y = im.b_chead[bb];
    im.b_chead[bb] = 2 * i + 1;
  }
  return 0;
}

// image -> snapshot (caller-provided capacities in out->n)
static int image_unpack(const BatchHost* bh, const WorldImage& im, b2gpu_snapshot* out) {
  const Batch& B = bh->B;
  const Topology& T = bh->topo;
  b2gpu_snapshot_sizes need;
  need.body_count = B.NB; need.fixture_count = B.NF; need.shape_count = B.NS; need.proxy_count = B.NP;
  need.node_count = im.ws[WS_TREE_CAP]; need.contact_count = im.ws[WS_CONTACT_COUNT]; need.move_count = im.ws[WS_MOVE_COUNT];
  need.joint_count = B.NJ;
  if (B.NJ > 0 && (out->n.joint_count < B.NJ || !out->joints)) {
    set_error("download buffers smaller than snapshot_sizes (joints)");
    return B2GPU_E_INVALID;
  }
  if (out->n.body_count < need.body_count || out->n.fixture_count < need.fixture_count || out->n.shape_count < need.shape_count ||
      out->n.proxy_count < need.proxy_count || out->n.node_count < need.node_count ||
      out->n.contact_count < need.contact_count || out->n.move_count < need.move_count) {
    set_error("download buffers smaller than snapshot_sizes");
    return B2GPU_E_INVALID;
  }
  out->n = need;
  b2gpu_world_rec& w = out->world;
  memset(&w, 0, sizeof(w));
  w.gravity_x = bitsf(im.ws[WS_GRAVITY_X]); w.gravity_y = bitsf(im.ws[WS_GRAVITY_Y]);
  w.inv_dt0 = bitsf(im.ws[WS_INV_DT0]);
  w.flags = (uint32_t)im.ws[WS_FLAGS];
  w.tree_root = im.ws[WS_TREE_ROOT]; w.tree_free_list = im.ws[WS_TREE_FREE]; w.tree_node_count = im.ws[WS_TREE_COUNT];
  w.tree_node_capacity = im.ws[WS_TREE_CAP]; w.tree_insertion_count = im.ws[WS_TREE_INSERTIONS];
  w.proxy_count = im.ws[WS_PROXY_COUNT];
  for (int i = 0; i < B.NB; ++i) {
    b2gpu_body_rec& b = out->bodies[i];
    b = T.bodies[i];
    b.type = (im.b_flags[i] >> 16) & 0xff;
    b.flags = (uint32_t)(im.b_flags[i] & 0xffff);
    b.xf_px = im.b_xf[i].x; b.xf_py = im.b_xf[i].y; b.xf_qs = im.b_xf[i].z; b.xf_qc = im.b_xf[i].w;
    b.c_x = im.b_pos[i].x; b.c_y = im.b_pos[i].y; b.a = im.b_pos[i].z; b.sleep_time = im.b_pos[i].w;
    b.c0_x = im.b_pos0[i].x; b.c0_y = im.b_pos0[i].y; b.a0 = im.b_pos0[i].z;
    b.vx = im.b_vel[i].x; b.vy = im.b_vel[i].y; b.w = im.b_vel[i].z;
    b.inv_mass = im.b_mass[i].x; b.inv_inertia = im.b_mass[i].y; b.lc_x = im.b_mass[i].z; b.lc_y = im.b_mass[i].w;
    b.fx = im.b_force[i].x; b.fy = im.b_force[i].y; b.torque = im.b_force[i].z; b.gravity_scale = im.b_force[i].w;
    b.mass = im.b_misc[i].x; b.inertia = im.b_misc[i].y; b.linear_damping = im.b_misc[i].z; b.angular_damping = im.b_misc[i].w;
  }
  memcpy(out->fixtures, T.fixtures.data(), sizeof(b2gpu_fixture_rec) * B.NF);
  memcpy(out->shapes, T.shapes.data(), sizeof(b2gpu_shape_rec) * B.NS);
  for (int i = 0; i < B.NP; ++i) {
    out->proxies[i] = T.proxies[i];
    out->proxies[i].aabb[0] = im.p_aabb[i].x; out->proxies[i].aabb[1] = im.p_aabb[i].y;
    out->proxies[i].aabb[2] = im.p_aabb[i].z; out->proxies[i].aabb[3] = im.p_aabb[i].w;
  }
  for (int i = 0; i < need.node_count; ++i) {
    b2gpu_tree_node_rec& n = out->nodes[i];
    n.aabb[0] = im.n_aabb[i].x; n.aabb[1] = im.n_aabb[i].y; n.aabb[2] = im.n_aabb[i].z; n.aabb[3] = im.n_aabb[i].w;
    n.parent = im.n_link[i].x; n.child1 = im.n_link[i].y; n.child2 = im.n_link[i].z; n.height = im.n_link[i].w;
    n.proxy = (n.height == 0 && i < (int)T.node_proxy.size()) ? T.node_proxy[i] : -1;
    n.moved = im.n_moved[i];
  }
  for (int i = 0; i < need.contact_count; ++i) {
    b2gpu_contact_rec& c = out->contacts[i];
    memset(&c, 0, sizeof(c));
    c.fixture_a = im.c_fix[i].x; c.fixture_b = im.c_fix[i].y; c.index_a = im.c_fix[i].z; c.index_b = im.c_fix[i].w;
    c.flags = (uint32_t)(im.c_flags[i] & 0xff);
    c.friction = im.c_mat[i].x; c.restitution = im.c_mat[i].y; c.restitution_threshold = im.c_mat[i].z; c.tangent_speed = im.c_mat[i].w;
    b2gpu_manifold& m = c.manifold;
    m.points[0].lp_x = im.c_m0[i].x; m.points[0].lp_y = im.c_m0[i].y;
    m.points[0].normal_impulse = im.c_m0[i].z; m.points[0].tangent_impulse = im.c_m0[i].w;
    m.points[1].lp_x = im.c_m1[i].x; m.points[1].lp_y = im.c_m1[i].y;
    m.points[1].normal_impulse = im.c_m1[i].z; m.points[1].tangent_impulse = im.c_m1[i].w;
    m.points[0].id = (uint32_t)im.c_m3[i].x; m.points[1].id = (uint32_t)im.c_m3[i].y;
    m.ln_x = im.c_m2[i].x; m.ln_y = im.c_m2[i].y; m.lp_x = im.c_m2[i].z; m.lp_y = im.c_m2[i].w;
    m.type = im.c_m3[i].z; m.point_count = im.c_m3[i].w;
  }
  for (int i = 0; i < need.move_count; ++i) out->move_buffer[i] = im.move_buf[i];
  for (int i = 0; i < B.NJ; ++i) {
    b2gpu_joint_rec& j = out->joints[i];
    j = T.joints[i];
    const float4 s0 = im.j_s0[i], s1 = im.j_s1[i];
    if (j.type == B2GPU_JOINT_GEAR) { j.impulse[0] = s0.x; continue; }  // impulse[1..6] are static data of the record
    j.impulse[0] = s0.x; j.impulse[1] = s0.y; j.impulse[2] = s0.z; j.impulse[3] = s0.w; j.impulse[4] = s1.x;
    j.impulse[5] = j.impulse[6] = j.impulse[7] = 0.0f;
    if (j.type == B2GPU_JOINT_REVOLUTE || j.type == B2GPU_JOINT_PRISMATIC || j.type == B2GPU_JOINT_WHEEL || j.type == B2GPU_JOINT_MOUSE) {
      j.param[4] = s1.y;
      j.param[3] = s1.z;
      j.flags = (j.flags & B2GPU_JOINT_COLLIDE_CONNECTED) | ((uint32_t)fbits(s1.w) & (B2GPU_JOINT_ENABLE_LIMIT | B2GPU_JOINT_ENABLE_MOTOR));
    }
  }
  return 0;
}

static int topology_build(const b2gpu_snapshot* s, Topology& T) {
  const b2gpu_snapshot_sizes& n = s->n;
  T.bodies.assign(s->bodies, s->bodies + n.body_count);
  T.fixtures.assign(s->fixtures, s->fixtures + n.fixture_count);
  T.shapes.assign(s->shapes, s->shapes + n.shape_count);
  T.proxies.assign(s->proxies, s->proxies + n.proxy_count);
  T.proxy_s.resize(n.proxy_count);
  T.node_proxy.assign(std::max(n.node_count, 1), -1);
  for (int p = 0; p < n.proxy_count; ++p) {
    const b2gpu_proxy_rec& pr = T.proxies[p];
    if (pr.fixture < 0 || pr.fixture >= n.fixture_count || pr.proxy_id < 0 || pr.proxy_id >= n.node_count) {
      set_error("proxy record out of range");
      return B2GPU_E_INVALID;
    }
    T.proxy_s[p] = make_int4(pr.fixture, pr.child_index, pr.proxy_id, T.fixtures[pr.fixture].body);
    T.node_proxy[pr.proxy_id] = p;
  }
  for (int f = 0; f < n.fixture_count; ++f) {
    const b2gpu_fixture_rec& fx = T.fixtures[f];
    if (fx.body < 0 || fx.body >= n.body_count || fx.shape_first < 0 || fx.shape_first + fx.child_count > n.shape_count) {
      set_error("fixture record out of range");
      return B2GPU_E_INVALID;
    }
    if (fx.proxy_first >= 0 && fx.proxy_first + fx.child_count > n.proxy_count) {
      set_error("fixture proxies out of range");
      return B2GPU_E_INVALID;
    }
  }
  // joints: static table + per-body joint lists.  create_joint pushes edge A on body A's list and then edge B on
  // body B's (b2_world.rs(private):176-208); the lists are push_front lists, so a body iterates its edges by
  // descending joint creation order.
  T.joints.assign(s->joints, s->joints + (s->joints ? n.joint_count : 0));
  if ((int)T.joints.size() != n.joint_count) { set_error("joint table missing"); return B2GPU_E_INVALID; }
  T.jadj_off.assign(n.body_count + 1, 0);
  T.jadj.assign(2 * (size_t)n.joint_count, 0);
  for (const b2gpu_joint_rec& j : T.joints) {
    if (j.type != B2GPU_JOINT_REVOLUTE && j.type != B2GPU_JOINT_DISTANCE && j.type != B2GPU_JOINT_WELD && j.type != B2GPU_JOINT_PRISMATIC &&
        j.type != B2GPU_JOINT_WHEEL && j.type != B2GPU_JOINT_FRICTION && j.type != B2GPU_JOINT_MOTOR && j.type != B2GPU_JOINT_PULLEY &&
        j.type != B2GPU_JOINT_MOUSE && j.type != B2GPU_JOINT_GEAR) {
      set_error("unknown joint type");
      return B2GPU_E_UNSUPPORTED;
    }
    if (j.type == B2GPU_JOINT_GEAR) {
      int32_t bc, bd;
      memcpy(&bc, &j.impulse[5], 4); memcpy(&bd, &j.impulse[6], 4);
      if (bc < 0 || bc >= n.body_count || bd < 0 || bd >= n.body_count) { set_error("gear joint record: body C / D out of range"); return B2GPU_E_INVALID; }
    }
    if (j.body_a < 0 || j.body_a >= n.body_count || j.body_b < 0 || j.body_b >= n.body_count || j.body_a == j.body_b) {
      set_error("joint record out of range");
      return B2GPU_E_INVALID;
    }
    T.jadj_off[j.body_a + 1] += 1;
    T.jadj_off[j.body_b + 1] += 1;
  }
  for (int b = 0; b < n.body_count; ++b) T.jadj_off[b + 1] += T.jadj_off[b];
  {
    std::vector<int> fill(T.jadj_off.begin(), T.jadj_off.end() - 1);
    for (int j = n.joint_count - 1; j >= 0; --j) {  // newest joint first
      T.jadj[fill[T.joints[j].body_a]++] = 2 * j;
      T.jadj[fill[T.joints[j].body_b]++] = 2 * j + 1;
    }
  }
  // synchronize order: body list newest first, each body's fixtures newest first, children ascending
  T.sync_order.clear();
  for (int b = n.body_count - 1; b >= 0; --b)
    for (int f = T.bodies[b].fixture_head; f != -1; f = T.fixtures[f].next) {
      if (f < 0 || f >= n.fixture_count) { set_error("fixture list corrupt"); return B2GPU_E_INVALID; }
      if (T.fixtures[f].proxy_first < 0) continue;
      for (int c = 0; c < T.fixtures[f].child_count; ++c) T.sync_order.push_back(T.fixtures[f].proxy_first + c);
    }
  T.sync_rank.assign(std::max(n.proxy_count, 1), 0);
  for (size_t r = 0; r < T.sync_order.size(); ++r)
    if (T.sync_order[r] >= 0 && T.sync_order[r] < n.proxy_count) T.sync_rank[T.sync_order[r]] = (int)r;
  if ((int)T.sync_order.size() != n.proxy_count) {
    set_error("proxy table does not match the fixture lists");
    return B2GPU_E_INVALID;
  }
  return 0;
}

static bool topology_matches(const Topology& T, const b2gpu_snapshot* s) {
  const b2gpu_snapshot_sizes& n = s->n;
  if (n.body_count != (int)T.bodies.size() || n.fixture_count != (int)T.fixtures.size() ||
      n.shape_count != (int)T.shapes.size() || n.proxy_count != (int)T.proxies.size())
    return false;
  if (n.fixture_count && memcmp(s->fixtures, T.fixtures.data(), sizeof(b2gpu_fixture_rec) * n.fixture_count)) return false;
  if (n.shape_count && memcmp(s->shapes, T.shapes.data(), sizeof(b2gpu_shape_rec) * n.shape_count)) return false;
  for (int p = 0; p < n.proxy_count; ++p)
    if (s->proxies[p].fixture != T.proxies[p].fixture || s->proxies[p].child_index != T.proxies[p].child_index ||
        s->proxies[p].proxy_id != T.proxies[p].proxy_id)
      return false;
  for (int b = 0; b < n.body_count; ++b)
    if (s->bodies[b].type != T.bodies[b].type || s->bodies[b].fixture_head != T.bodies[b].fixture_head) return false;
  if (n.joint_count != (int)T.joints.size()) return false;
  for (int j = 0; j < n.joint_count; ++j) {
    const b2gpu_joint_rec &a = s->joints[j], &b = T.joints[j];
    if (a.type != b.type || a.body_a != b.body_a || a.body_b != b.body_b ||
        (a.flags & B2GPU_JOINT_COLLIDE_CONNECTED) != (b.flags & B2GPU_JOINT_COLLIDE_CONNECTED) ||
        memcmp(a.local_anchor_a, b.local_anchor_a, 8) || memcmp(a.local_anchor_b, b.local_anchor_b, 8))
      return false;
    const bool motorised = a.type == B2GPU_JOINT_REVOLUTE || a.type == B2GPU_JOINT_PRISMATIC || a.type == B2GPU_JOINT_WHEEL ||
                           a.type == B2GPU_JOINT_MOUSE;
    const int n_static = motorised ? 3 : 5;  // revolute / prismatic: motor torque (force) / speed are per world
    if (memcmp(a.param, b.param, 4 * n_static)) return false;
    if (a.type == B2GPU_JOINT_PRISMATIC && memcmp(a.param + 5, b.param + 5, 8)) return false;   // the local axis
    if (a.type == B2GPU_JOINT_WHEEL && memcmp(a.param + 5, b.param + 5, 12)) return false;      // the local axis, damping
    if (a.type == B2GPU_JOINT_PULLEY && memcmp(a.param + 5, b.param + 5, 12)) return false;     // length_b, ratio, constant
    if (a.type == B2GPU_JOINT_GEAR && (memcmp(a.param + 5, b.param + 5, 12) || memcmp(a.impulse + 1, b.impulse + 1, 24) ||
                                       ((a.flags ^ b.flags) & (B2GPU_JOINT_GEAR_PRISMATIC_1 | B2GPU_JOINT_GEAR_PRISMATIC_2)))) return false;
  }
  return true;
}

template <class T> static int alloc_arr(BatchHost* bh, T** p, long long count) {
  void* v = nullptr;
  RC(dev_alloc(&v, (size_t)count * sizeof(T)));
  bh->allocs.push_back(v);
  bh->total_bytes += count * (long long)sizeof(T);
  *p = (T*)v;
  return 0;
}

static int move_array(BatchHost* bh, const ArrRef& a, int mode, int world) {
  Batch& B = bh->B;
  const size_t bytes = (size_t)a.N * a.E * 4;
  if (bytes > bh->stage_bytes) { set_error("staging buffer too small"); return B2GPU_E_INVALID; }
  MoveK k;
  k.dev = (int*)a.dev; k.compact = (int*)bh->stage_dev; k.N = a.N; k.E = a.E; k.mode = mode;
  k.LB = B.LB; k.lb_shift = B.lb_shift; k.n_wblocks = B.n_wblocks; k.w = world;
  if (mode == 1) {
    RC(launch(bh->ctx, k, a.N * a.E, 256));
    RC(dev_d2h(bh->ctx, a.host, bh->stage_dev, bytes));
  } else {
    RC(dev_h2d(bh->ctx, bh->stage_dev, a.host, bytes));
    const long long n = mode == 2 ? (long long)B.n_wblocks * a.N * B.LB * a.E : (long long)a.N * a.E;
    if (n > 0x7fffffffLL) { set_error("array too large for 32-bit indexing"); return B2GPU_E_CAPACITY; }
    RC(launch(bh->ctx, k, (int)n, 256));
    RC(ctx_sync(bh->ctx));
  }
  return 0;
}

int batch_create(Ctx* ctx, const b2gpu_snapshot* proto, int n_worlds, const b2gpu_caps* caps, int lane_block, BatchHost** out) {
  if (!ctx || !proto || n_worlds < 1 || !out) { set_error("batch_create: bad argument"); return B2GPU_E_INVALID; }
  RC(b2gpu_snapshot_validate(proto));
  BatchHost* bh = new BatchHost();
  bh->ctx = ctx;
  int rc = topology_build(proto, bh->topo);
  if (rc) { delete bh; return rc; }
  Batch& B = bh->B;
  memset(&B, 0, sizeof(B));
  B.n_worlds = n_worlds;
  B.LB = lane_block > 0 ? lane_block : (n_worlds >= 32 ? 32 : 1);
  if (B.LB & (B.LB - 1)) { set_error("lane block must be a power of two"); delete bh; return B2GPU_E_INVALID; }
  B.lb_shift = 0;
  while ((1 << B.lb_shift) < B.LB) ++B.lb_shift;
  B.n_wblocks = (n_worlds + B.LB - 1) / B.LB;
  B.wb_first = 0;
  B.wb_count = B.n_wblocks;
  const b2gpu_snapshot_sizes& n = proto->n;
  B.NB = n.body_count; B.NF = n.fixture_count; B.NS = n.shape_count; B.NP = n.proxy_count;
  if (B.NB < 1 || B.NP < 0) { set_error("empty world"); delete bh; return B2GPU_E_INVALID; }
  B.NN = std::max(n.node_count, 16);
  // broadphase stress scenes (AddPair) reach ~30 contacts per body: a single world gets generous room,
  // batches (thousands of replicas) default to 10 per proxy; b2gpu_caps.max_contacts overrides both
  int want_contacts = std::max(n.contact_count * 2, B.NP * (n_worlds == 1 ? 40 : 10) + 64);
  if (caps && caps->max_contacts > 0) want_contacts = std::max(caps->max_contacts, n.contact_count);
  B.NC = want_contacts;
  B.NMOVE = std::max(2 * B.NP, n.move_count) + 16;
  B.NIB = B.NB + B.NC;
  B.NJ = n.joint_count;
  B.NMW = std::max((B.NP + 31) / 32, 1);
  if (B.NP < 1) B.NP = 0;
  const long long W = (long long)B.n_wblocks * B.LB;
  if (W * B.NC * VC_Q > 0x7fffffffLL || W * B.NIB > 0x7fffffffLL) {
    set_error("batch too large for 32-bit element indexing");
    delete bh;
    return B2GPU_E_CAPACITY;
  }
#define AL(ptr, count) do { rc = alloc_arr(bh, &ptr, (count)); if (rc) { batch_destroy(bh); return rc; } } while (0)
  // shared topology
  b2gpu_fixture_rec* d_fix; b2gpu_shape_rec* d_shape; int4* d_ps; int* d_so; int* d_np; int* d_sr;
  AL(d_fix, std::max(B.NF, 1)); AL(d_shape, std::max(B.NS, 1)); AL(d_ps, std::max(B.NP, 1)); AL(d_so, std::max(B.NP, 1));
  AL(d_np, (long long)bh->topo.node_proxy.size());
  AL(d_sr, std::max(B.NP, 1));
  B.fixtures = d_fix; B.shapes = d_shape; B.proxy_s = d_ps; B.sync_order = d_so; B.node_proxy = d_np; B.sync_rank = d_sr;
  b2gpu_joint_rec* d_joints; int* d_joff; int* d_jadj;
  AL(d_joints, std::max(B.NJ, 1)); AL(d_joff, B.NB + 1); AL(d_jadj, std::max(2 * B.NJ, 1));
  B.joints = d_joints; B.jadj_off = d_joff; B.jadj = d_jadj;
  const int NPa = std::max(B.NP, 1);
  AL(B.ws, W * WS_COUNT);
  AL(B.b_flags, W * B.NB); AL(B.b_xf, W * B.NB); AL(B.b_pos, W * B.NB); AL(B.b_pos0, W * B.NB); AL(B.b_vel, W * B.NB);
  AL(B.b_mass, W * B.NB); AL(B.b_force, W * B.NB); AL(B.b_misc, W * B.NB); AL(B.b_rot, W * B.NB);
  AL(B.n_aabb, W * B.NN); AL(B.n_link, W * B.NN); AL(B.n_moved, W * B.NN);
  AL(B.p_aabb, W * NPa); AL(B.p_fat, W * NPa); AL(B.p_move, W * B.NMW);
  AL(B.move_buf, W * B.NMOVE);
  AL(B.c_fix, W * B.NC); AL(B.c_flags, W * B.NC); AL(B.c_mat, W * B.NC);
  AL(B.c_m0, W * B.NC); AL(B.c_m1, W * B.NC); AL(B.c_m2, W * B.NC); AL(B.c_m3, W * B.NC);
  AL(B.isl_body, W * B.NIB); AL(B.isl_contact, W * B.NC); AL(B.isl_range, W * B.NB); AL(B.isl_flags, W * B.NB);
  AL(B.c_isl, W * B.NC); AL(B.vc, W * B.NC * VC_Q); AL(B.pc, W * B.NC * PC_Q);
  if (B.NJ > 0) {
    AL(B.j_s0, W * B.NJ); AL(B.j_s1, W * B.NJ); AL(B.isl_joint, W * B.NJ); AL(B.isl_jrange, W * B.NB); AL(B.j_flag, W * B.NJ);
    AL(B.j_tmp, W * B.NJ * JT_Q);
  }
  AL(bh->b_wake, W * B.NB); AL(bh->b_chead, W * B.NB); AL(bh->c_next, W * B.NC); AL(bh->stack, W * B.NB);
  AL(bh->state_dev, (long long)n_worlds * B.NB * 8);
  AL(bh->forces_dev, (long long)n_worlds * B.NB * 3);
  AL(bh->status_dev, 4);
  AL(bh->vel_scratch, (long long)n_worlds * 2);
  {
    std::vector<int> dyn;
    for (int b = 0; b < B.NB; ++b)
      if (bh->topo.bodies[b].type == B2GPU_DYNAMIC_BODY) dyn.push_back(b);
    bh->n_dyn = (int)dyn.size();
    AL(bh->dyn_idx, std::max(bh->n_dyn, 1));
    if (bh->n_dyn) { rc = dev_h2d(ctx, bh->dyn_idx, dyn.data(), dyn.size() * 4); if (rc) { batch_destroy(bh); return rc; } }
  }
#if defined(B2G_HOSTSIM)
  bh->status_host = (int*)calloc(4, 4);
#else
  { cudaError_t e_ = cudaMallocHost((void**)&bh->status_host, 16); if (e_ != cudaSuccess) { batch_destroy(bh); return cuda_fail(e_, "cudaMallocHost(status)"); } bh->status_host[0] = 0; }
#endif
  bh->stage_bytes = 16 * (size_t)std::max(std::max(std::max(B.NB, B.NJ), B.NN), std::max(std::max(B.NC, B.NMOVE), (int)WS_COUNT));
  {
    void* v = nullptr;
    rc = dev_alloc(&v, bh->stage_bytes);
    if (rc) { batch_destroy(bh); return rc; }
    bh->stage_dev = v;
    bh->allocs.push_back(v);
  }
#undef AL
  const Topology& T = bh->topo;
#define UP(dst, vec) do { if (!(vec).empty()) { rc = dev_h2d(ctx, (void*)(dst), (vec).data(), (vec).size() * sizeof((vec)[0])); if (rc) { batch_destroy(bh); return rc; } } } while (0)
  UP(d_fix, T.fixtures); UP(d_shape, T.shapes); UP(d_ps, T.proxy_s); UP(d_so, T.sync_order); UP(d_np, T.node_proxy); UP(d_sr, T.sync_rank);
  UP(d_joints, T.joints); UP(d_joff, T.jadj_off); UP(d_jadj, T.jadj);
#undef UP
  WorldImage im;
  rc = image_pack(bh, proto, im);
  if (rc) { batch_destroy(bh); return rc; }
  std::vector<ArrRef> tab = array_table(bh, im);
  for (const ArrRef& a : tab) {
    rc = move_array(bh, a, 2, 0);
    if (rc) { batch_destroy(bh); return rc; }
  }
  bh->pre_step_needed = true;
  if (caps && (caps->reserved[1] == 11 || caps->reserved[1] == 12)) {  // large-world mode (12: with the exact replica tree): one world, data-parallel ordered stages (b2g_large.h)
    if (n_worlds != 1 || B.LB != 1) { set_error("large-world mode needs a batch of exactly one world"); batch_destroy(bh); return B2GPU_E_INVALID; }
    rc = large_alloc(bh);
    if (rc) { batch_destroy(bh); return rc; }
    bh->large = true;
    bh->lw_exact_tree = caps->reserved[1] == 12;
  }
#if !defined(B2G_HOSTSIM)
  // shared-memory Gauss-Seidel stages: batches in 32-world memory blocks whose bodies fit one SM
  bh->smem_solver = false;
  if (B.LB == 32 && !(caps && caps->reserved[1] == 1)) {
    int max_optin = 0;
    CU(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    const size_t need = std::max(velocity_smem_bytes(B.NB), position_sl_smem_bytes(B.NB));
    if (need <= (size_t)max_optin) {
      CU(cudaFuncSetAttribute(velocity_sl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)velocity_smem_bytes(B.NB)));
      CU(cudaFuncSetAttribute(position_sl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)position_sl_smem_bytes(B.NB)));
      bh->smem_solver = true;
      if (const char* e = getenv("B2GPU_TIMELINE")) {  // diagnostic: per-CTA start/end of the Gauss-Seidel kernels
        const size_t cap = 600000;
        unsigned long long* t = nullptr;
        rc = alloc_arr(bh, &t, 2 + cap * 3);
        if (rc) { batch_destroy(bh); return rc; }
        const unsigned long long head[2] = {0ull, (unsigned long long)cap};
        rc = dev_h2d(ctx, t, head, sizeof(head));
        if (rc) { batch_destroy(bh); return rc; }
        bh->B.timeline = t;
        bh->timeline_path = e;
      }
    }
    if (const char* e = getenv("B2GPU_STREAM_GROUPS")) bh->stream_groups = atoi(e);  // tuning experiments
    if (const char* e = getenv("B2GPU_STAGGER")) bh->stagger_groups = atoi(e) != 0;
    if (caps && caps->reserved[1] == 5) bh->stream_groups = 1;
    if (caps && caps->reserved[1] == 6) bh->use_graphs = false;
    bh->island_layout = island_smem_layout(B.NB, B.NF, (size_t)max_optin - 1024);
    // worlds with joints build their islands with the generic traversal (joint edges: SerialAK::islands_global)
    if (B.NB < 32768 && B.NC < 65536 && bh->island_layout.ECAP >= 64 && B.NJ == 0) {
      CU(cudaFuncSetAttribute(island_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bh->island_layout.total));
      bh->smem_island = true;
    }
  }
#endif
  *out = bh;
  return 0;
}

void batch_destroy(BatchHost* bh) {
  if (!bh) return;
#if !defined(B2G_HOSTSIM)
  for (StreamGroup& sg : bh->groups) {
    cudaStreamSynchronize((cudaStream_t)sg.stream);
    cudaEventDestroy((cudaEvent_t)sg.ev_init);
    cudaEventDestroy((cudaEvent_t)sg.ev_done);
    cudaStreamDestroy((cudaStream_t)sg.stream);
  }
  if (bh->ev_entry) cudaEventDestroy((cudaEvent_t)bh->ev_entry);
  for (StepGraph& g : bh->graphs) cudaGraphExecDestroy((cudaGraphExec_t)g.exec);
#endif
#if !defined(B2G_HOSTSIM)
  if (bh->B.timeline && !bh->timeline_path.empty()) {
    unsigned long long head[2] = {0, 0};
    cudaDeviceSynchronize();
    cudaMemcpy(head, bh->B.timeline, sizeof(head), cudaMemcpyDeviceToHost);
    const size_t n = (size_t)std::min(head[0], head[1]);
    std::vector<unsigned long long> rows(n * 3);
    if (n) cudaMemcpy(rows.data(), bh->B.timeline + 2, n * 3 * 8, cudaMemcpyDeviceToHost);
    if (FILE* f = fopen(bh->timeline_path.c_str(), "wb")) { fwrite(rows.data(), 8, rows.size(), f); fclose(f); }
  }
#endif
  large_free(bh);
#if defined(B2G_HOSTSIM)
  free(bh->status_host);
#else
  if (bh->status_host) cudaFreeHost(bh->status_host);
#endif
  if (bh->query_buf) dev_free(bh->query_buf);
  for (void* p : bh->allocs) dev_free(p);
  delete bh;
}

int batch_upload_world(BatchHost* bh, int world, const b2gpu_snapshot* in) {
  if (!bh || !in || world < 0 || world >= bh->B.n_worlds) { set_error("upload_world: bad argument"); return B2GPU_E_INVALID; }
  RC(b2gpu_snapshot_validate(in));  // every index inside its table: a damaged snapshot must not reach the kernels
  if (!topology_matches(bh->topo, in)) {
    set_error("upload_world: snapshot topology (fixtures/shapes/proxies/body types) differs from the batch prototype");
    return B2GPU_E_INVALID;
  }
  WorldImage im;
  RC(image_pack(bh, in, im));
  std::vector<ArrRef> tab = array_table(bh, im);
  for (const ArrRef& a : tab) RC(move_array(bh, a, 0, world));
  bh->pre_step_needed = true;
  bh->lw_cc_valid = false;
  RC(dev_zero(bh->ctx, bh->status_dev, 4));  // recomputed from the worlds' own (sticky) status words on the next check
  return 0;
}

// Every world of the batch back to the state of `in` (an RL-style reset of all environments): the same broadcast
// batch_create performs for the prototype.
int batch_reset(BatchHost* bh, const b2gpu_snapshot* in) {
  if (!bh || !in) { set_error("batch_reset: bad argument"); return B2GPU_E_INVALID; }
  RC(b2gpu_snapshot_validate(in));
  if (!topology_matches(bh->topo, in)) {
    set_error("batch_reset: snapshot topology (fixtures/shapes/proxies/body types) differs from the batch prototype");
    return B2GPU_E_INVALID;
  }
  if (bh->large) { set_error("batch_reset: use upload_world for the single world of the large-world mode"); return B2GPU_E_INVALID; }
  WorldImage im;
  RC(image_pack(bh, in, im));
  RC(ctx_sync(bh->ctx));
  std::vector<ArrRef> tab = array_table(bh, im);
  for (const ArrRef& a : tab) RC(move_array(bh, a, 2, 0));
  bh->pre_step_needed = true;
  bh->stepped = false;
  RC(dev_zero(bh->ctx, bh->status_dev, 4));
  return 0;
}

static int image_fetch(BatchHost* bh, int world, WorldImage& im) {
  if (bh->smem_island && bh->stepped) {  // materialise the contacts' ISLAND bits (see ContactIslandFlagsK)
    const int n = bh->B.n_wblocks * bh->B.LB * bh->B.NC;
    { ContactIslandFlagsK k = {bh->B, 0}; RC(launch(bh->ctx, k, n, 128)); }
    { ContactIslandFlagsK k = {bh->B, 1}; RC(launch(bh->ctx, k, n, 128)); }
  }
  image_alloc(bh->B, im);
  std::vector<ArrRef> tab = array_table(bh, im);
  for (const ArrRef& a : tab) RC(move_array(bh, a, 1, world));
  bh->last_fetch_status = im.ws[WS_STATUS];
  if (bh->large && !bh->lw_exact_tree) {
    // large-world mode does not maintain the replica tree: the leaf boxes are current, the topology is the
    // one last uploaded.  Refit the internal boxes (children before parents) so the snapshot carries a valid
    // bounding hierarchy for whoever continues from it (host-side proxy creation, the exact mode).
    const int root = im.ws[WS_TREE_ROOT];
    if (root >= 0) {
      std::vector<int> order, st;
      st.push_back(root);
      while (!st.empty()) {
        const int i = st.back();
        st.pop_back();
        const int4 l = im.n_link[i];
        if (l.y == -1) continue;
        order.push_back(i);
        st.push_back(l.y);
        st.push_back(l.z);
      }
      for (size_t k = order.size(); k-- > 0;) {
        const int i = order[k];
        const float4 a = im.n_aabb[im.n_link[i].y], b = im.n_aabb[im.n_link[i].z];
        im.n_aabb[i] = make_float4(std::min(a.x, b.x), std::min(a.y, b.y), std::max(a.z, b.z), std::max(a.w, b.w));
      }
    }
  }
  return 0;
}

int batch_snapshot_sizes(BatchHost* bh, int world, b2gpu_snapshot_sizes* out) {
  if (!bh || !out || world < 0 || world >= bh->B.n_worlds) { set_error("snapshot_sizes: bad argument"); return B2GPU_E_INVALID; }
  WorldImage im;
  image_alloc(bh->B, im);
  ArrRef a = {bh->B.ws, im.ws.data(), 1, WS_COUNT};
  RC(move_array(bh, a, 1, world));
  out->body_count = bh->B.NB; out->fixture_count = bh->B.NF; out->shape_count = bh->B.NS; out->proxy_count = bh->B.NP;
  out->node_count = im.ws[WS_TREE_CAP]; out->contact_count = im.ws[WS_CONTACT_COUNT]; out->move_count = im.ws[WS_MOVE_COUNT];
  out->joint_count = bh->B.NJ;
  return 0;
}

int batch_download_world(BatchHost* bh, int world, b2gpu_snapshot* out) {
  if (!bh || !out || world < 0 || world >= bh->B.n_worlds) { set_error("download_world: bad argument"); return B2GPU_E_INVALID; }
  WorldImage im;
  RC(image_fetch(bh, world, im));
  return image_unpack(bh, im, out);
}
// Error code of the world last downloaded (its sticky WS_STATUS), 0 if it is healthy: the buffers of the download
// are filled either way, so a failed world can still be inspected.
int batch_last_download_status(BatchHost* bh) { return bh ? status_error(bh->last_fetch_status) : B2GPU_E_INVALID; }

// ------------------------------------------------------------------ step
// body state gather / force scatter: flat over bodies
struct StateGatherK {
  Batch B;
  float* out;  // [n_worlds][NB][8]
  B2G_HD void operator()(int tid) const {
    int w, b;
    if (!flat_decode(B, tid, B.NB, w, b)) return;
    WIdx x = widx(B, w);
    const int bi = x.at(B.NB, b);
    const float4 pos = B.b_pos[bi], vel = B.b_vel[bi], xf = B.b_xf[bi];
    float* o = out + ((size_t)w * B.NB + b) * 8;
    o[0] = pos.x; o[1] = pos.y; o[2] = pos.z; o[3] = vel.x; o[4] = vel.y; o[5] = vel.z; o[6] = xf.x; o[7] = xf.y;
  }
};
struct ForceScatterK {
  Batch B;
  const float* in;  // [count][NB][3] for worlds first..first+count
  int first, count;
  B2G_HD void operator()(int tid) const {
    int w, b;
    if (!flat_decode(B, tid, B.NB, w, b)) return;
    if (w < first || w >= first + count) return;
    WIdx x = widx(B, w);
    const int bi = x.at(B.NB, b);
    const int bf = B.b_flags[bi];
    // B2body::apply_force / apply_torque with wake=false: only awake dynamic bodies accumulate
    if (body_type(bf) != B2GPU_DYNAMIC_BODY || !(bf & B2GPU_BODY_AWAKE)) return;
    const float* f = in + ((size_t)(w - first) * B.NB + b) * 3;
    float4 fo = B.b_force[bi];
    fo.x += f[0]; fo.y += f[1]; fo.z += f[2];
    B.b_force[bi] = fo;
  }
};
// Compact I/O (RL loops): only the prototype's DYNAMIC bodies, in body order — forces [n_worlds][nd][3] in, state
// [n_worlds][nd][6] = (c.x, c.y, a, v.x, v.y, w) out (24 B per body: SURVEY §8e's 5040 B per Pyramid world).
struct StateGatherDynK {
  Batch B;
  float* out;
  const int* dyn;
  int nd;
  B2G_HD void operator()(int tid) const {
    int w, j;
    if (!flat_decode(B, tid, nd, w, j)) return;
    WIdx x = widx(B, w);
    const int bi = x.at(B.NB, dyn[j]);
    const float4 pos = B.b_pos[bi], vel = B.b_vel[bi];
    float* o = out + ((size_t)w * nd + j) * 6;
    o[0] = pos.x; o[1] = pos.y; o[2] = pos.z; o[3] = vel.x; o[4] = vel.y; o[5] = vel.z;
  }
};
struct ForceScatterDynK {
  Batch B;
  const float* in;  // [count][nd][3] for worlds first..first+count
  const int* dyn;
  int nd, first, count;
  B2G_HD void operator()(int tid) const {
    int w, j;
    if (!flat_decode(B, tid, nd, w, j)) return;
    if (w < first || w >= first + count) return;
    WIdx x = widx(B, w);
    const int bi = x.at(B.NB, dyn[j]);
    if (!(B.b_flags[bi] & B2GPU_BODY_AWAKE)) return;  // apply_force / apply_torque with wake = false
    const float* f = in + ((size_t)(w - first) * nd + j) * 3;
    float4 fo = B.b_force[bi];
    fo.x += f[0]; fo.y += f[1]; fo.z += f[2];
    B.b_force[bi] = fo;
  }
};
struct VelScatterK {
  Batch B;
  const float* in;  // [count][2]
  int body, first, count;
  B2G_HD void operator()(int i) const {
    if (i >= count) return;
    const int w = first + i;
    WIdx x = widx(B, w);
    const int bi = x.at(B.NB, body);
    const int bf = B.b_flags[bi];
    if (body_type(bf) == B2GPU_STATIC_BODY) return;  // B2body::set_linear_velocity
    const float vx = in[2 * i], vy = in[2 * i + 1];
    if (vx * vx + vy * vy > 0.0f) {
      B.b_flags[bi] = bf | B2GPU_BODY_AWAKE;
      B.b_pos[bi].w = 0.0f;
      if (!(bf & B2GPU_BODY_AWAKE)) ws_of(B, x)[WS_TOPO_DIRTY] = 1;
    }
    float4 v = B.b_vel[bi];
    v.x = vx; v.y = vy;
    B.b_vel[bi] = v;
  }
};

struct GravityScatterK {  // b2gpu_batch_set_gravity
  Batch B;
  const float* in;  // [count][2]
  int first, count;
  B2G_HD void operator()(int i) const {
    if (i >= count) return;
    WIdx x = widx(B, first + i);
    Ws ws = ws_of(B, x);
    ws[WS_GRAVITY_X] = f2i(in[2 * i]);
    ws[WS_GRAVITY_Y] = f2i(in[2 * i + 1]);
  }
};
// b2gpu_batch_set_joint_control: the revolute / mouse setters in every world of a range (include/b2gpu.h)
struct JointControlK {
  Batch B;
  const float* in;  // [count] or [count][2]
  int joint, control, first, count;
  B2G_HD void wake(const WIdx& x, int body) const {  // B2body::set_awake(true): static bodies stay as they are
    const int bi = x.at(B.NB, body);
    const int bf = B.b_flags[bi];
    if (body_type(bf) == B2GPU_STATIC_BODY) return;
    B.b_flags[bi] = bf | B2GPU_BODY_AWAKE;
    B.b_pos[bi].w = 0.0f;
    if (!(bf & B2GPU_BODY_AWAKE)) ws_of(B, x)[WS_TOPO_DIRTY] = 1;
  }
  B2G_HD void operator()(int i) const {
    if (i >= count) return;
    WIdx x = widx(B, first + i);
    const b2gpu_joint_rec& jr = B.joints[joint];
    const int ji = x.at(B.NJ, joint);
    float4 s1 = B.j_s1[ji];
    if (control == B2GPU_JOINT_CONTROL_TARGET) {
      const float tx = in[2 * i], ty = in[2 * i + 1];
      if (tx == s1.z && ty == s1.y) return;
      wake(x, jr.body_b);
      s1.z = tx; s1.y = ty;
    } else {
      const float v = in[i];
      float& slot = control == B2GPU_JOINT_CONTROL_MOTOR_SPEED ? s1.y : s1.z;
      if (v == slot) return;
      wake(x, jr.body_a);
      wake(x, jr.body_b);
      slot = v;
    }
    B.j_s1[ji] = s1;
  }
};

// Device-side failures (contact table / move buffer / island list full, unregistered shape pair, query stack
// overflow) are recorded per world in WS_STATUS and stay set; this reduces them to one word so that every call that
// already synchronises can return the first failure instead of 0 (the word is the most negative code of any world).
struct StatusK {
  Batch B;
  int* out;
  B2G_HD void operator()(int w) const {
    if (w >= B.n_worlds) return;
    WIdx x = widx(B, w);
    const int st = ws_of(B, x)[WS_STATUS];
    if (st < 0) B2G_ATOMIC_MIN(out, st);
  }
};
static int status_error(int st) {
  if (st == 0) return 0;
  set_error(st == B2GPU_E_CAPACITY ? "a world ran out of device capacity during a step (contact table, move buffer or island list full: "
                                     "raise b2gpu_caps.max_contacts); its state is no longer the reference's"
            : st == B2GPU_E_UNSUPPORTED ? "a world hit an unsupported case on the device (unregistered shape pair, e.g. edge against edge)"
                                        : "a world raised an internal device status");
  return st;
}

// All stages of `steps` consecutive steps for the world-block window of Bw, on ctx->stream.
// `init_done` (optional cudaEvent_t) is recorded after the solver set-up stage of the first step: the
// next stream group starts behind it, so the groups run staggered (one group's flat stages overlap the
// other groups' latency-bound Gauss-Seidel kernels instead of all groups doing the same stage at once).
static int step_window(BatchHost* bh, const Batch& Bw, const StepParams& sp, int steps, void* init_done) {
  Ctx* ctx = bh->ctx;
  const Batch& B = Bw;
  const int W = B.wb_count * B.LB;
  const int ordered_block = 32;
  const float dt = sp.dt;
  for (int s = 0; s < steps; ++s) {
    {
      TreePairsK k = {B, bh->b_chead, bh->c_next, sp, 1};
      RC(launch(ctx, k, W, ordered_block, STAGE_PRE));
    }
    {
      CollideK k = {B, bh->b_wake};
      RC(launch_occ(ctx, k, W * B.NC, STAGE_COLLIDE));
    }
    {
      SerialAK k = {B, bh->b_wake, bh->b_chead, bh->c_next, bh->stack, sp};
#if !defined(B2G_HOSTSIM)
      if (bh->smem_island) {
        LaunchScope ls = {ctx, STAGE_ISLAND};
        RC(ls.begin());
        island_smem_kernel<<<B.wb_count, 32, bh->island_layout.total, (cudaStream_t)ctx->stream>>>(k, bh->island_layout);
        RC(ls.end());
      } else
#endif
      RC(launch(ctx, k, W, ordered_block, STAGE_ISLAND));
    }
    if (dt > 0.0f) {
      { IntegrateK k = {B, sp}; RC(launch(ctx, k, W * B.NIB, 128, STAGE_INTEGRATE)); }
      { SolverInitK k = {B, sp}; RC(launch(ctx, k, W * B.NC, 128, STAGE_SOLVER_INIT)); }
#if !defined(B2G_HOSTSIM)
      if (init_done && s == 0) CU(cudaEventRecord((cudaEvent_t)init_done, (cudaStream_t)ctx->stream));
      if (bh->smem_solver) {
        LaunchScope ls = {ctx, STAGE_VELOCITY};
        RC(ls.begin());
        velocity_sl_kernel<<<B.wb_count, 32, velocity_smem_bytes(B.NB), (cudaStream_t)ctx->stream>>>(B, sp);
        RC(ls.end());
      } else
#endif
      { VelocityK k = {B, sp}; RC(launch(ctx, k, W * B.NB, 64, STAGE_VELOCITY)); }
      { PostVelocityK k = {B, sp}; RC(launch(ctx, k, W * B.NIB, 128, STAGE_POST_VELOCITY)); }
#if !defined(B2G_HOSTSIM)
      if (bh->smem_solver) {
        LaunchScope ls = {ctx, STAGE_POSITION};
        RC(ls.begin());
        position_sl_kernel<<<B.wb_count, 32, position_sl_smem_bytes(B.NB), (cudaStream_t)ctx->stream>>>(B, sp);
        RC(ls.end());
      } else
#endif
      { PositionK k = {B, sp}; RC(launch(ctx, k, W * B.NB, 64, STAGE_POSITION)); }
      { FinalizeK k = {B, sp}; RC(launch(ctx, k, W * B.NIB, 128, STAGE_FINALIZE)); }
      { SleepK k = {B}; RC(launch(ctx, k, W * B.NB, 128, STAGE_SLEEP)); }
      if (B.NP > 0) { SyncFixturesK k = {B}; RC(launch(ctx, k, W * B.NP, 128, STAGE_SYNC_FIXTURES)); }
    }
#if !defined(B2G_HOSTSIM)
    else if (init_done && s == 0) CU(cudaEventRecord((cudaEvent_t)init_done, (cudaStream_t)ctx->stream));
#endif
    {
      TreePairsK k = {B, bh->b_chead, bh->c_next, sp, 0};
      RC(launch(ctx, k, W, ordered_block, STAGE_TREE_PAIRS));
    }
    { BodyEndK k = {B}; RC(launch(ctx, k, W * B.NB, 128, STAGE_BODY_END)); }
  }
  return 0;
}

// ------------------------------------------------------------------ large-world mode (b2g_large.h)
// Host-driven: the world's scalar slots are read back at a few points of the step (a 4-byte-scale copy and
// a stream synchronisation each), so every launch is sized by the live counts and stages with nothing to do
// are skipped.  Scans and sorts are CUB device primitives (std:: algorithms in the host simulator).
#if defined(B2G_HOSTSIM)
static int lw_scan_int(BatchHost*, const int* in, int* out, int n, int) {
  int acc = 0;
  for (int i = 0; i < n; ++i) { const int v = in[i]; out[i] = acc; acc += v; }
  return 0;
}
static int lw_scan_u64(BatchHost*, const u64* in, u64* out, int n, int) {
  u64 acc = 0;
  for (int i = 0; i < n; ++i) { const u64 v = in[i]; out[i] = acc; acc += v; }
  return 0;
}
static int lw_sort_keys(BatchHost*, const u64* in, u64* out, int n, int, int) {
  std::copy(in, in + n, out);
  std::sort(out, out + n);
  return 0;
}
static int lw_sort_pairs32(BatchHost*, const unsigned* kin, unsigned* kout, const unsigned* vin, unsigned* vout, int n, int) {
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return kin[a] < kin[b]; });
  for (int i = 0; i < n; ++i) { kout[i] = kin[order[i]]; vout[i] = vin[order[i]]; }
  return 0;
}
static int lw_read(BatchHost* bh, const void* dev, int words) { memcpy(bh->lw_host, dev, (size_t)words * 4); return 0; }
#else
static int lw_sort_pairs32(BatchHost* bh, const unsigned* kin, unsigned* kout, const unsigned* vin, unsigned* vout, int n, int stage) {
  if (n <= 0) return 0;
  LaunchScope ls = {bh->ctx, stage};
  RC(ls.begin());
  size_t bytes = bh->lw_tmp_bytes;
  CU(cub::DeviceRadixSort::SortPairs(bh->lw_tmp, bytes, kin, kout, vin, vout, n, 0, 32, (cudaStream_t)bh->ctx->stream));
  return ls.end();
}
static int lw_scan_int(BatchHost* bh, const int* in, int* out, int n, int stage) {
  if (n <= 0) return 0;
  LaunchScope ls = {bh->ctx, stage};
  RC(ls.begin());
  size_t bytes = bh->lw_tmp_bytes;
  CU(cub::DeviceScan::ExclusiveSum(bh->lw_tmp, bytes, in, out, n, (cudaStream_t)bh->ctx->stream));
  return ls.end();
}
static int lw_scan_u64(BatchHost* bh, const u64* in, u64* out, int n, int stage) {
  if (n <= 0) return 0;
  LaunchScope ls = {bh->ctx, stage};
  RC(ls.begin());
  size_t bytes = bh->lw_tmp_bytes;
  CU(cub::DeviceScan::ExclusiveSum(bh->lw_tmp, bytes, in, out, n, (cudaStream_t)bh->ctx->stream));
  return ls.end();
}
static int lw_sort_keys(BatchHost* bh, const u64* in, u64* out, int n, int end_bit, int stage) {
  if (n <= 0) return 0;
  LaunchScope ls = {bh->ctx, stage};
  RC(ls.begin());
  size_t bytes = bh->lw_tmp_bytes;
  CU(cub::DeviceRadixSort::SortKeys(bh->lw_tmp, bytes, in, out, n, 0, end_bit, (cudaStream_t)bh->ctx->stream));
  return ls.end();
}
static int lw_read(BatchHost* bh, const void* dev, int words) {
  CU(cudaMemcpyAsync(bh->lw_host, dev, (size_t)words * 4, cudaMemcpyDeviceToHost, (cudaStream_t)bh->ctx->stream));
  CU(cudaStreamSynchronize((cudaStream_t)bh->ctx->stream));
  return 0;
}
#endif
static int ceil_log2(long long v) { int b = 0; while ((1LL << b) < v) ++b; return b; }

static int large_alloc(BatchHost* bh) {
  Batch& B = bh->B;
  Large& L = bh->L;
  if (B.NB >= (1 << 20) || B.NC >= (1 << 24) || B.NP < 1) {
    set_error("large-world mode: needs 1 <= proxies, bodies < 2^20, contact capacity < 2^24");
    return B2GPU_E_CAPACITY;
  }
  bh->lw_edge_bits = ceil_log2(2LL * B.NC);
  bh->lw_body_bits = ceil_log2(B.NB);
  bh->lw_keys = std::max<long long>(2LL * B.NC, B.NP) + 1;
  L.NCAND = 2 * B.NC + B.NP;
  int rc = 0;
#define AL(ptr, count) do { rc = alloc_arr(bh, &ptr, (count)); if (rc) return rc; } while (0)
  AL(L.keys, bh->lw_keys); AL(L.keys_alt, bh->lw_keys); AL(L.sort_in, 2LL * B.NP); AL(L.sort_out, 2LL * B.NP);
  AL(L.lb_box, B.NP); AL(L.lb_child, B.NP); AL(L.lb_parent, 2LL * B.NP); AL(L.lb_flag, B.NP); AL(L.lb_leaf, B.NP);
  const long long nq = std::max(B.NMOVE, B.NMW) + 1;
  AL(L.q_cnt, nq); AL(L.q_off, nq); AL(L.q_local, (long long)B.NMOVE * LW_QLOCAL);
  AL(L.cand, L.NCAND); AL(L.cand_flag, L.NCAND + 1LL); AL(L.cand_pos, L.NCAND + 1LL); AL(L.cand_fix, L.NCAND);
  AL(L.uf_parent, B.NB); AL(L.cnt_b, B.NB); AL(L.cnt_c, B.NB); AL(L.seed, B.NB); AL(L.isl_seed, B.NB);
  AL(L.cnt_j, B.NB); AL(L.pj_in, B.NB + 1LL); AL(L.pj_out, B.NB + 1LL);
  AL(L.pk_in, B.NB + 1LL); AL(L.pk_out, B.NB + 1LL);
  AL(L.keep_flag, B.NC + 1LL); AL(L.keep_pos, B.NC + 1LL);
  AL(L.first_idx, B.NN); AL(L.vc_idx, B.NC); AL(L.scratch4, 16);
  AL(L.wake_idx, B.NB + 1LL);
  AL(L.sleep_min, B.NB + 1LL); AL(L.lv_meta, 4); AL(L.lv_info, LW_MAXG); AL(L.lv_isl_giant, B.NB + 1LL); AL(L.lv_last, B.NB + 1LL); AL(L.lv_level, B.NC + 1LL);
  AL(L.lv_count, (long long)B.NC + B.NB + 2); AL(L.lv_start, (long long)B.NC + B.NB + 2); AL(L.lv_order, B.NC + 1LL); AL(L.lv_ix, B.NC + 1LL); AL(L.lv_vrec, (B.NC + 1LL) * LV_VQ); AL(L.lv_prec, (B.NC + 1LL) * LV_PQ);
  AL(L.state, B.NB); AL(L.adj, 2LL * B.NC); AL(L.eadj, 2LL * B.NC); AL(L.erow, B.NB); AL(L.row_start, B.NB); AL(L.row_end, B.NB);
#undef AL
#if defined(B2G_HOSTSIM)
  bh->lw_host = (int*)calloc(WS_COUNT + 16, 4);
#else
  CU(cudaMallocHost((void**)&bh->lw_host, (WS_COUNT + 16) * 4));
  size_t need = 0, b = 0;
  CU(cub::DeviceScan::ExclusiveSum(nullptr, b, (const int*)nullptr, (int*)nullptr, L.NCAND + 1));
  need = std::max(need, b);
  CU(cub::DeviceScan::ExclusiveSum(nullptr, b, (const u64*)nullptr, (u64*)nullptr, B.NB + 1));
  need = std::max(need, b);
  CU(cub::DeviceRadixSort::SortKeys(nullptr, b, (const u64*)nullptr, (u64*)nullptr, (int)bh->lw_keys, 0, 64));
  need = std::max(need, b);
  CU(cub::DeviceRadixSort::SortPairs(nullptr, b, (const unsigned*)nullptr, (unsigned*)nullptr, (const unsigned*)nullptr, (unsigned*)nullptr, B.NP, 0, 32));
  need = std::max(need, b);
  void* v = nullptr;
  RC(dev_alloc(&v, need + 256));
  bh->allocs.push_back(v);
  bh->lw_tmp = v;
  bh->lw_tmp_bytes = need + 256;
#endif
  return 0;
}
static void large_free(BatchHost* bh) {
  if (!bh->lw_host) return;
#if defined(B2G_HOSTSIM)
  free(bh->lw_host);
#else
  cudaFreeHost(bh->lw_host);
#endif
  bh->lw_host = nullptr;
}

// per-body contact edge lists from scratch (after the contact set changed)
static int lw_rebuild_lists(BatchHost* bh, int cc, int stage) {
  Ctx* ctx = bh->ctx;
  const Batch& B = bh->B;
  { LwEdgeKeysK k = {B, bh->L, cc, bh->lw_edge_bits}; RC(launch(ctx, k, std::max(2 * cc, B.NB), 256, stage)); }
  if (cc == 0) return 0;
  RC(lw_sort_keys(bh, bh->L.keys, bh->L.keys_alt, 2 * cc, bh->lw_edge_bits + bh->lw_body_bits, stage));
  { LwEdgeRowsK k = {B, bh->L, 2 * cc, bh->lw_edge_bits}; RC(launch(ctx, k, 2 * cc, 256, stage)); }
  return 0;
}

// LBVH over the fat boxes of all proxies: Morton keys, radix sort, Karras construction, bottom-up refit
static int lw_build_lbvh(BatchHost* bh, int stage) {
  Ctx* ctx = bh->ctx;
  const Batch& B = bh->B;
  const Large& L = bh->L;
  const int n = B.NP;
  { LwMortonK k = {B, L}; RC(launch(ctx, k, n, 256, stage)); }
  RC(lw_sort_pairs32(bh, L.sort_in, L.sort_out, L.sort_in + n, L.sort_out + n, n, stage));
  { LwKeyPackK k = {B, L}; RC(launch(ctx, k, n, 256, stage)); }
  { LwKarrasK k = {B, L, n}; RC(launch(ctx, k, n, 128, stage)); }
  { LwRefitK k = {B, L, n}; RC(launch(ctx, k, n, 128, stage)); }
  return 0;
}

// B2broadPhase::update_pairs + add_pair over the current move buffer (mc entries, cc contacts before)
static int lw_update_pairs(BatchHost* bh, int mc, int cc, int stage, int use_tree, int* created_out = nullptr) {
  if (created_out) *created_out = 0;
  Ctx* ctx = bh->ctx;
  const Batch& B = bh->B;
  const Large& L = bh->L;
  if (mc <= 0) return 0;
  const int n = B.NP;
  if (!use_tree) RC(lw_build_lbvh(bh, stage));
  if (use_tree) {
    { LwMoveFirstK k = {B, L, mc, 0}; RC(launch(ctx, k, mc, 256, stage)); }
    { LwMoveFirstK k = {B, L, mc, 1}; RC(launch(ctx, k, mc, 256, stage)); }
  }
  { LwQueryK k = {B, L, mc, n, 0, use_tree}; RC(launch(ctx, k, mc + 1, 64, stage)); }
  RC(lw_scan_int(bh, L.q_cnt, L.q_off, mc + 1, stage));
  RC(lw_read(bh, L.q_off + mc, 1));
  int n_cand = bh->lw_host[0], status = 0;
  if (n_cand > L.NCAND) { n_cand = L.NCAND; status = B2GPU_E_CAPACITY; }
  int created = 0;
  if (n_cand > 0) {
    { LwQueryK k = {B, L, mc, n, 1, use_tree}; RC(launch(ctx, k, mc, 64, stage)); }
    { LwAddPairK k = {B, L, n_cand}; RC(launch(ctx, k, n_cand + 1, 128, stage)); }
    RC(lw_scan_int(bh, L.cand_flag, L.cand_pos, n_cand + 1, stage));
    RC(lw_read(bh, L.cand_pos + n_cand, 1));
    created = bh->lw_host[0];
    if (cc + created > B.NC) { created = B.NC - cc; status = B2GPU_E_CAPACITY; }
    if (created > 0) { LwCreateK k = {B, L, n_cand, cc}; RC(launch(ctx, k, n_cand, 128, stage)); }
  }
  { LwClearMovedK k = {B, mc}; RC(launch(ctx, k, mc, 256, stage)); }
  { LwPairsFinishK k = {B, mc, n_cand, created, status}; RC(launch(ctx, k, 1, 32, stage)); }
  if (created > 0) RC(lw_rebuild_lists(bh, cc + created, stage));
  if (created_out) *created_out = created;
  return 0;
}

static int step_large(BatchHost* bh, const StepParams& sp, int steps) {
  Ctx* ctx = bh->ctx;
  const Batch& B = bh->B;
  const Large& L = bh->L;
  int* hw = bh->lw_host;
  const float dt = sp.dt;
  for (int s = 0; s < steps; ++s) {
    // ---- top of step: stats, find_new_contacts when m_new_contacts (b2_world.rs(private):912-915)
    { LwStatsResetK k = {B}; RC(launch(ctx, k, 1, 32, STAGE_PRE)); }
    // the contact count is known on the host from the previous step of this call sequence (count after the
    // destruction pass + contacts created): one read-back and stream synchronisation less per step
    const bool know = bh->lw_cc_valid && !(bh->pre_step_needed && s == 0);
    if (!know) RC(lw_read(bh, B.ws, WS_COUNT));
    else hw[WS_FLAGS] &= ~B2GPU_WORLD_NEW_CONTACTS;
    int cc = know ? bh->lw_cc : hw[WS_CONTACT_COUNT];
    if (bh->pre_step_needed && s == 0) RC(lw_rebuild_lists(bh, cc, STAGE_PRE));  // first step after an upload: contact rows
    if (hw[WS_FLAGS] & B2GPU_WORLD_NEW_CONTACTS) {
      const int mc = hw[WS_MOVE_COUNT];
      RC(lw_update_pairs(bh, mc, cc, STAGE_PRE, 1));
      { LwClearNewContactsK k = {B}; RC(launch(ctx, k, 1, 32, STAGE_PRE)); }
      RC(lw_read(bh, B.ws, WS_COUNT));
      cc = hw[WS_CONTACT_COUNT];
    }
    // ---- collide
    { CollideK k = {B, bh->b_wake}; RC(launch_occ(ctx, k, cc, STAGE_COLLIDE)); }
    RC(lw_read(bh, B.ws, WS_COUNT));
    if (hw[WS_EV_WAKE]) {
      { LwWakeK k = {B, L, bh->b_wake, cc, 0}; RC(launch(ctx, k, B.NB, 256, STAGE_ISLAND)); }
      { LwWakeK k = {B, L, bh->b_wake, cc, 1}; RC(launch(ctx, k, cc, 256, STAGE_ISLAND)); }
      for (int round = 0; round <= cc; ++round) {
        { LwWakeK k = {B, L, bh->b_wake, cc, 2}; RC(launch_occ(ctx, k, cc, STAGE_ISLAND)); }
        RC(lw_read(bh, L.wake_idx + B.NB, 1));
        if (!hw[0]) break;
        { LwWakeK k = {B, L, bh->b_wake, cc, 3}; RC(launch(ctx, k, 1, 32, STAGE_ISLAND)); }
      }
      { LwWakeK k = {B, L, bh->b_wake, cc, 3}; RC(launch(ctx, k, 1, 32, STAGE_ISLAND)); }
      RC(lw_read(bh, B.ws, WS_COUNT));
    }
    if (hw[WS_TOPO_DIRTY]) { LwWakeMergeK k = {B, bh->b_wake}; RC(launch(ctx, k, B.NB, 256, STAGE_ISLAND)); }
    bool dirty = hw[WS_TOPO_DIRTY] != 0;
    if (hw[WS_EV_DESTROY]) {
      const int new_cc = cc - hw[WS_EV_DESTROY];
      { LwDestroyFlagK k = {B, L, cc}; RC(launch(ctx, k, cc + 1, 256, STAGE_ISLAND)); }
      RC(lw_scan_int(bh, L.keep_flag, L.keep_pos, cc + 1, STAGE_ISLAND));
      { LwCompactK k = {B, L, cc, 0}; RC(launch(ctx, k, cc, 256, STAGE_ISLAND)); }
      { LwCompactK k = {B, L, new_cc, 1}; RC(launch(ctx, k, new_cc, 256, STAGE_ISLAND)); }
      { LwDestroyFinishK k = {B, sp, cc, new_cc}; RC(launch(ctx, k, 1, 32, STAGE_ISLAND)); }
      cc = new_cc;
      RC(lw_rebuild_lists(bh, cc, STAGE_ISLAND));
      dirty = true;
    }
    if (dt > 0.0f) {
      // ---- islands
      if (dirty) {
        const int nbc = std::max(std::max(B.NB, cc), B.NJ);
        { LwIslInitK k = {B, L, cc}; RC(launch(ctx, k, nbc, 256, STAGE_ISLAND)); }
        { LwUnionK k = {B, L, cc}; RC(launch(ctx, k, std::max(cc, B.NJ), 256, STAGE_ISLAND)); }
        { LwCountK k = {B, L, cc}; RC(launch(ctx, k, nbc, 256, STAGE_ISLAND)); }
        { LwSeedPackK k = {B, L}; RC(launch(ctx, k, B.NB + 1, 256, STAGE_ISLAND)); }
        RC(lw_scan_u64(bh, L.pk_in, L.pk_out, B.NB + 1, STAGE_ISLAND));
        if (B.NJ > 0) RC(lw_scan_int(bh, L.pj_in, L.pj_out, B.NB + 1, STAGE_ISLAND));
        { LwRangeK k = {B, L}; RC(launch(ctx, k, B.NB + 1, 256, STAGE_ISLAND)); }
        RC(lw_read(bh, B.ws, WS_COUNT));
        { LwAdjInfoK k = {B, L, 2 * cc}; RC(launch(ctx, k, 2 * cc + 1, 256, STAGE_ISLAND)); }
        RC(lw_scan_int(bh, L.cand_flag, L.cand_pos, 2 * cc + 1, STAGE_ISLAND));
        { LwAdjCompactK k = {B, L, 2 * cc}; RC(launch(ctx, k, std::max(2 * cc, B.NB), 256, STAGE_ISLAND)); }
        // giant islands (>= lw_level_min contacts, no joints): chosen here — their traversal has a prefetching warp, their
        // sweeps are level-scheduled (b2g_levels.h); every other island one thread
        {
          const bool lv = bh->lw_level_min > 0 && hw[WS_ISL_CONTACTS] >= bh->lw_level_min;
          { LwLevelResetK k = {L}; RC(launch(ctx, k, 1, 32, STAGE_ISLAND)); }
          { LwGiantSelectK k = {B, L, hw[WS_ISL_COUNT], lv ? bh->lw_level_min : 0}; RC(launch(ctx, k, hw[WS_ISL_COUNT], 256, STAGE_ISLAND)); }
          if (lv) { LwDfsGiantK k = {B, L, bh->stack}; RC(launch_cta(ctx, k, LW_MAXG, 64, STAGE_ISLAND)); }
        }
        { LwDfsK k = {B, L, bh->stack, hw[WS_ISL_COUNT]}; RC(launch(ctx, k, hw[WS_ISL_COUNT], 32, STAGE_ISLAND)); }
        { LwIslFlagsK k = {B, hw[WS_ISL_BODIES], hw[WS_ISL_CONTACTS]}; RC(launch(ctx, k, std::max(hw[WS_ISL_BODIES], hw[WS_ISL_CONTACTS]), 256, STAGE_ISLAND)); }
      } else {
        LwIslCachedK k = {B};
        RC(launch(ctx, k, 1, 32, STAGE_ISLAND));
      }
      const int ni = hw[WS_ISL_COUNT], nib = hw[WS_ISL_BODIES], nic = hw[WS_ISL_CONTACTS];
      { LwStaticRotK k = {B}; RC(launch(ctx, k, B.NB, 256, STAGE_INTEGRATE)); }
      { IntegrateK k = {B, sp}; RC(launch(ctx, k, nib, 128, STAGE_INTEGRATE)); }
      { SolverInitK k = {B, sp}; RC(launch(ctx, k, nic, 128, STAGE_SOLVER_INIT)); }
      // LwVelocity7K: records and bodies staged through a cp.async shared-memory ring (islands under 16 contacts take the
      // register form LwVelocity5K); LwPosition6K: two alternating register sets.  The other forms round 1 measured
      // (profiles/r01_large_world.md) were removed from the library.
      { LwVcIdxK k = {B, L, nic}; RC(launch(ctx, k, nic, 256, STAGE_SOLVER_INIT)); }
      // giant islands without joints (>= lw_level_min contacts): level schedule rebuilt with the islands, one CTA per island
      // sweeps level by level (b2g_levels.h); every other island one thread
      const bool levels = bh->lw_level_min > 0 && nic >= bh->lw_level_min;
      if (dirty && levels) { LwLevelBuildK k = {B, L}; RC(launch_cta(ctx, k, LW_MAXG, LW_LEVEL_BUILD_NT, STAGE_ISLAND)); }
      if (levels) {
        { LwLevelGatherK k = {B, L, nic}; RC(launch(ctx, k, nic, 256, STAGE_SOLVER_INIT)); }
        { LwLevelVelocityK k = {B, L, sp}; RC(launch_cta(ctx, k, LW_MAXG, LW_LEVEL_NT, STAGE_VELOCITY, LwLevelVelocityK::smem_bytes())); }
      }
      { LwVelocity7K k = {B, L, sp, ni}; RC(launch(ctx, k, ni, 32, STAGE_VELOCITY)); }
      { PostVelocityK k = {B, sp}; RC(launch(ctx, k, std::max(std::max(ni, nib), nic), 128, STAGE_POST_VELOCITY)); }
      if (levels) { LwLevelPositionK k = {B, L, sp}; RC(launch_cta(ctx, k, LW_MAXG, LW_LEVEL_NT, STAGE_POSITION, LwLevelPositionK::smem_bytes())); }
      { LwPosition6K k = {B, L, sp, ni}; RC(launch(ctx, k, ni, 32, STAGE_POSITION)); }
      { FinalizeK k = {B, sp}; RC(launch(ctx, k, nib, 128, STAGE_FINALIZE)); }
      for (int phase = 0; phase < 3; ++phase) { LwSleepK k = {B, L, ni, phase}; RC(launch(ctx, k, phase == 0 ? ni : 32 * ni, 256, STAGE_SLEEP)); }
      { SyncFixturesK k = {B}; RC(launch(ctx, k, B.NP, 128, STAGE_SYNC_FIXTURES)); }
      // ---- find_new_contacts
      RC(lw_read(bh, B.ws, WS_COUNT));
      int mc = hw[WS_MOVE_COUNT];
      if (hw[WS_EV_MOVED] && bh->lw_exact_tree) {
        { LwTreeMoveK k = {B}; RC(launch(ctx, k, 1, 32, STAGE_TREE_PAIRS)); }
        mc += hw[WS_EV_MOVED];
        if (mc > B.NMOVE) { set_error("move buffer capacity exceeded"); return B2GPU_E_CAPACITY; }
      } else if (hw[WS_EV_MOVED]) {
        { LwMoveCountK k = {B, L}; RC(launch(ctx, k, B.NMW + 1, 256, STAGE_TREE_PAIRS)); }
        RC(lw_scan_int(bh, L.q_cnt, L.q_off, B.NMW + 1, STAGE_TREE_PAIRS));
        { LwMoveEmitK k = {B, L, mc}; RC(launch(ctx, k, B.NMW, 128, STAGE_TREE_PAIRS)); }
        mc += hw[WS_EV_MOVED];
        if (mc > B.NMOVE) { set_error("move buffer capacity exceeded"); return B2GPU_E_CAPACITY; }
        { LwMoveFinishK k = {B, mc}; RC(launch(ctx, k, 1, 32, STAGE_TREE_PAIRS)); }
      }
      int created = 0;
      RC(lw_update_pairs(bh, mc, cc, STAGE_TREE_PAIRS, bh->lw_exact_tree ? 1 : 0, &created));
      cc += created;
    }
    bh->lw_cc = cc;
    bh->lw_cc_valid = true;
    { LwStepEndK k = {B, sp}; RC(launch(ctx, k, 1, 32, STAGE_TREE_PAIRS)); }
    { BodyEndK k = {B}; RC(launch(ctx, k, B.NB, 128, STAGE_BODY_END)); }
  }
  return 0;
}

// ------------------------------------------------------------------ world queries (b2g_query.h)
// Device scratch of the query calls: one buffer per batch, grown on demand and kept (a cudaMalloc / cudaFree pair
// per call cost more than the ray casts of a small world).
static int query_scratch(BatchHost* bh, size_t bytes, char** out) {
  if (bytes > bh->query_bytes) {
    if (bh->query_buf) {
      RC(ctx_sync(bh->ctx));
      dev_free(bh->query_buf);
      bh->query_buf = nullptr;
      bh->query_bytes = 0;
    }
    void* v = nullptr;
    const size_t want = bytes + bytes / 2 + 4096;
    RC(dev_alloc(&v, want));
    bh->query_buf = v;
    bh->query_bytes = want;
  }
  *out = (char*)bh->query_buf;
  return 0;
}
static size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }
static int query_prepare(BatchHost* bh, int& use_lbvh) {
  use_lbvh = bh->large && !bh->lw_exact_tree;  // the replica tree is current in every other mode
  if (use_lbvh && bh->B.NP > 0) RC(lw_build_lbvh(bh, STAGE_OTHER));
  return 0;
}
int batch_ray_cast_closest(BatchHost* bh, const float* host_rays, int rays_per_world, b2gpu_ray_hit* host_out) {
  if (!bh || !host_rays || !host_out || rays_per_world < 0) { set_error("ray_cast: bad argument"); return B2GPU_E_INVALID; }
  const Batch& B = bh->B;
  const long long total = (long long)B.n_worlds * rays_per_world;
  if (total == 0) return 0;
  if (total > 0x7fffffffLL) { set_error("ray_cast: too many rays"); return B2GPU_E_CAPACITY; }
  for (long long i = 0; i < total; ++i)
    if (host_rays[4 * i] == host_rays[4 * i + 2] && host_rays[4 * i + 1] == host_rays[4 * i + 3]) {
      set_error("ray_cast: p1 == p2 (the reference asserts length_squared > 0)");
      return B2GPU_E_INVALID;
    }
  char* base = nullptr;
  const size_t rays_bytes = align256((size_t)total * 16);
  RC(query_scratch(bh, rays_bytes + (size_t)total * sizeof(b2gpu_ray_hit), &base));
  float* d_rays = (float*)base;
  b2gpu_ray_hit* d_out = (b2gpu_ray_hit*)(base + rays_bytes);
  RC(dev_h2d(bh->ctx, d_rays, host_rays, (size_t)total * 16));
  int use_lbvh = 0;
  RC(query_prepare(bh, use_lbvh));
  { RayCastK k = {B, bh->L, d_rays, d_out, rays_per_world, use_lbvh, B.NP}; RC(launch(bh->ctx, k, (int)total, 64)); }
  RC(dev_d2h(bh->ctx, host_out, d_out, (size_t)total * sizeof(b2gpu_ray_hit)));
  return 0;
}
// per_world == 0: `n` boxes against the single world of `bh`; per_world > 0: per_world boxes for every world of the
// batch, boxes [n_worlds][per_world][4], n = n_worlds * per_world.
static int query_aabb_impl(BatchHost* bh, const float* host_boxes, int n, int per_world, int max_hits, int* host_counts, int* host_hits) {
  if (n == 0) return 0;
  if ((long long)n * (max_hits > 0 ? max_hits : 1) > 0x3fffffffLL) { set_error("query_aabb: too many boxes x hits"); return B2GPU_E_CAPACITY; }
  char* base = nullptr;
  const size_t boxes_bytes = align256((size_t)n * 16), counts_bytes = align256((size_t)n * 4);
  RC(query_scratch(bh, boxes_bytes + counts_bytes + (size_t)n * max_hits * 8 + 16, &base));
  float* d_boxes = (float*)base;
  int* d_counts = (int*)(base + boxes_bytes);
  int* d_hits = (int*)(base + boxes_bytes + counts_bytes);
  RC(dev_h2d(bh->ctx, d_boxes, host_boxes, (size_t)n * 16));
  int use_lbvh = 0;
  RC(query_prepare(bh, use_lbvh));
  { QueryAabbK k = {bh->B, bh->L, d_boxes, d_counts, d_hits, n, max_hits, use_lbvh, bh->B.NP, per_world}; RC(launch(bh->ctx, k, n, 64)); }
  RC(dev_d2h(bh->ctx, host_counts, d_counts, (size_t)n * 4));
  if (max_hits > 0) RC(dev_d2h(bh->ctx, host_hits, d_hits, (size_t)n * max_hits * 8));
  return 0;
}
int batch_query_aabb(BatchHost* bh, const float* host_boxes, int n, int max_hits, int* host_counts, int* host_hits) {
  if (!bh || !host_boxes || !host_counts || n < 0 || max_hits < 0 || (max_hits > 0 && !host_hits)) { set_error("query_aabb: bad argument"); return B2GPU_E_INVALID; }
  if (bh->B.n_worlds != 1) { set_error("query_aabb: one world at a time (use b2gpu_batch_query_aabb for a batch)"); return B2GPU_E_INVALID; }
  return query_aabb_impl(bh, host_boxes, n, 0, max_hits, host_counts, host_hits);
}
int batch_query_aabb_per_world(BatchHost* bh, const float* host_boxes, int boxes_per_world, int max_hits, int* host_counts, int* host_hits) {
  if (!bh || !host_boxes || !host_counts || boxes_per_world < 0 || max_hits < 0 || (max_hits > 0 && !host_hits)) { set_error("batch_query_aabb: bad argument"); return B2GPU_E_INVALID; }
  const long long total = (long long)bh->B.n_worlds * boxes_per_world;
  if (total > 0x7fffffffLL) { set_error("batch_query_aabb: too many boxes"); return B2GPU_E_CAPACITY; }
  if (boxes_per_world == 0) return 0;
  return query_aabb_impl(bh, host_boxes, (int)total, boxes_per_world, max_hits, host_counts, host_hits);
}

static StepParams make_params(float dt, int vi, int pi) {
  StepParams sp;
  sp.dt = dt;
  sp.inv_dt = dt > 0.0f ? 1.0f / dt : 0.0f;
  sp.velocity_iterations = vi;
  sp.position_iterations = pi;
  return sp;
}

#if !defined(B2G_HOSTSIM)
// Stream groups: the world blocks of a batch are split into a few windows, each an independent in-order
// pipeline on its own stream (worlds never interact).  Created on first use.
static int ensure_groups(BatchHost* bh) {
  if (!bh->groups.empty()) return 0;
  const Batch& B = bh->B;
  // automatic: 8 groups for >= 64 world blocks, 4 for >= 16 — but one stream for worlds of fewer than 96 bodies, whose step is
  // ~20 short launches that eight staggered groups only multiply (profiles/r02_stream_groups.md: car 0.32 -> 0.12 ms/step,
  // joints_mix 1.66 -> 0.64 at one step per call, and still ahead at 200 steps per call)
  const int ng_auto = B.NB < 96 ? 1 : (B.n_wblocks >= 64 ? 8 : 4);
  int ng = (B.LB == 32 && B.n_wblocks >= 16 && bh->stream_groups != 1) ? (bh->stream_groups > 1 ? bh->stream_groups : ng_auto) : 1;
  if (ng > B.n_wblocks) ng = B.n_wblocks;
  for (int g = 0; g < ng; ++g) {
    StreamGroup sg;
    sg.wb_first = (int)((long long)B.n_wblocks * g / ng);
    sg.wb_count = (int)((long long)B.n_wblocks * (g + 1) / ng) - sg.wb_first;
    cudaStream_t st;
    CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    sg.stream = (void*)st;
    cudaEvent_t e;
    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); sg.ev_init = (void*)e;
    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); sg.ev_done = (void*)e;
    bh->groups.push_back(sg);
  }
  cudaEvent_t e;
  CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  bh->ev_entry = (void*)e;
  return 0;
}
#endif

// Enqueues `steps` steps (and the optional host copies) on the context stream / the stream groups.
// Pure enqueue: no host synchronisation, so the same code runs under stream capture.
static int enqueue_steps(BatchHost* bh, const StepParams& sp, int steps, const float* host_forces, float* host_state_out, bool compact) {
  Batch& B = bh->B;
  Ctx* ctx = bh->ctx;
  (void)ctx;
  Batch all = B;
  all.wb_first = 0;
  all.wb_count = B.n_wblocks;
#if defined(B2G_HOSTSIM)
  if (host_forces) {
    if (compact) {
      memcpy(bh->forces_dev, host_forces, (size_t)B.n_worlds * bh->n_dyn * 3 * 4);
      ForceScatterDynK k = {all, bh->forces_dev, bh->dyn_idx, bh->n_dyn, 0, B.n_worlds};
      RC(launch(ctx, k, B.n_wblocks * B.LB * bh->n_dyn, 128));
    } else RC(batch_set_forces(bh, host_forces, 0, B.n_worlds));
  }
  if (bh->large) RC(step_large(bh, sp, steps));
  else RC(step_window(bh, all, sp, steps, nullptr));
  if (host_state_out && compact) {
    StateGatherDynK k = {all, bh->state_dev, bh->dyn_idx, bh->n_dyn};
    RC(launch(ctx, k, B.n_wblocks * B.LB * bh->n_dyn, 128));
    memcpy(host_state_out, bh->state_dev, (size_t)B.n_worlds * bh->n_dyn * 6 * 4);
    int st = 0;
    RC(batch_status(bh, &st));
    return status_error(st);
  }
  if (host_state_out) return batch_get_body_state(bh, host_state_out, 0, B.n_worlds);
  return 0;
#else
  const int fstride = compact ? bh->n_dyn * 3 : B.NB * 3, sstride = compact ? bh->n_dyn * 6 : B.NB * 8;
  const int io_n = compact ? bh->n_dyn : B.NB;
  cudaStream_t main_s = (cudaStream_t)ctx->stream;
  const bool grouped = bh->groups.size() > 1 && !ctx->profiling && steps > 0 && !bh->large;
  if (!grouped) {
    if (host_forces) {
      CU(cudaMemcpyAsync(bh->forces_dev, host_forces, (size_t)B.n_worlds * fstride * 4, cudaMemcpyHostToDevice, main_s));
      if (compact) { ForceScatterDynK k = {all, bh->forces_dev, bh->dyn_idx, bh->n_dyn, 0, B.n_worlds}; RC(launch(ctx, k, B.n_wblocks * B.LB * io_n, 128)); }
      else { ForceScatterK k = {all, bh->forces_dev, 0, B.n_worlds}; RC(launch(ctx, k, B.n_wblocks * B.LB * B.NB, 128)); }
    }
    if (bh->large) RC(step_large(bh, sp, steps));
    else RC(step_window(bh, all, sp, steps, nullptr));
    if (host_state_out) {
      if (compact) { StateGatherDynK k = {all, bh->state_dev, bh->dyn_idx, bh->n_dyn}; RC(launch(ctx, k, B.n_wblocks * B.LB * io_n, 128)); }
      else { StateGatherK k = {all, bh->state_dev}; RC(launch(ctx, k, B.n_wblocks * B.LB * B.NB, 128)); }
      CU(cudaMemcpyAsync(host_state_out, bh->state_dev, (size_t)B.n_worlds * sstride * 4, cudaMemcpyDeviceToHost, main_s));
      { StatusK k = {all, bh->status_dev}; RC(launch(ctx, k, B.n_worlds, 128)); }
      CU(cudaMemcpyAsync(bh->status_host, bh->status_dev, 4, cudaMemcpyDeviceToHost, main_s));
    }
    return 0;
  }
  CU(cudaEventRecord((cudaEvent_t)bh->ev_entry, main_s));
  int rc = 0;
  for (size_t g = 0; g < bh->groups.size() && !rc; ++g) {
    StreamGroup& sg = bh->groups[g];
    cudaStream_t gs = (cudaStream_t)sg.stream;
    Batch Bw = B;
    Bw.wb_first = sg.wb_first;
    Bw.wb_count = sg.wb_count;
    const int w0 = sg.wb_first * B.LB;
    const int wn = std::min(B.n_worlds, (sg.wb_first + sg.wb_count) * B.LB) - w0;
    CU(cudaStreamWaitEvent(gs, (cudaEvent_t)bh->ev_entry, 0));
    ctx->stream = (void*)gs;
    if (host_forces && wn > 0) {
      const size_t off = (size_t)w0 * fstride;
      cudaError_t e = cudaMemcpyAsync(bh->forces_dev + off, host_forces + off, (size_t)wn * fstride * 4, cudaMemcpyHostToDevice, gs);
      if (e != cudaSuccess) { ctx->stream = (void*)main_s; return cuda_fail(e, "cudaMemcpyAsync(forces)"); }
      if (compact) { ForceScatterDynK k = {Bw, bh->forces_dev + off, bh->dyn_idx, bh->n_dyn, w0, wn}; rc = launch(ctx, k, sg.wb_count * B.LB * io_n, 128); }
      else { ForceScatterK k = {Bw, bh->forces_dev + off, w0, wn}; rc = launch(ctx, k, sg.wb_count * B.LB * B.NB, 128); }
    }
    // stagger: start behind the previous group's solver set-up of its first step
    if (!rc && g > 0 && bh->stagger_groups) {
      cudaError_t e = cudaStreamWaitEvent(gs, (cudaEvent_t)bh->groups[g - 1].ev_init, 0);
      if (e != cudaSuccess) { ctx->stream = (void*)main_s; return cuda_fail(e, "cudaStreamWaitEvent"); }
    }
    if (!rc) rc = step_window(bh, Bw, sp, steps, sg.ev_init);
    if (!rc && host_state_out && wn > 0) {
      if (compact) { StateGatherDynK k = {Bw, bh->state_dev, bh->dyn_idx, bh->n_dyn}; rc = launch(ctx, k, sg.wb_count * B.LB * io_n, 128); }
      else { StateGatherK k = {Bw, bh->state_dev}; rc = launch(ctx, k, sg.wb_count * B.LB * B.NB, 128); }
      if (!rc) {
        const size_t off = (size_t)w0 * sstride;
        cudaError_t e = cudaMemcpyAsync(host_state_out + off, bh->state_dev + off, (size_t)wn * sstride * 4, cudaMemcpyDeviceToHost, gs);
        if (e != cudaSuccess) { ctx->stream = (void*)main_s; return cuda_fail(e, "cudaMemcpyAsync(state)"); }
      }
    }
    ctx->stream = (void*)main_s;
    if (!rc) {
      CU(cudaEventRecord((cudaEvent_t)sg.ev_done, gs));
      CU(cudaStreamWaitEvent(main_s, (cudaEvent_t)sg.ev_done, 0));
    }
  }
  if (!rc && host_state_out) {
    { StatusK k = {all, bh->status_dev}; RC(launch(ctx, k, B.n_worlds, 128)); }
    CU(cudaMemcpyAsync(bh->status_host, bh->status_dev, 4, cudaMemcpyDeviceToHost, main_s));
  }
  return rc;
#endif
}

#if !defined(B2G_HOSTSIM)
static bool host_pointer_capturable(const void* p) {  // memcpy nodes need page-locked host memory
  if (!p) return true;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}
#endif

// Runs `steps` steps; optional host buffers: forces are uploaded before the first step, the body state is
// downloaded after the last.  The enqueue sequence of a call signature (dt, iterations, steps, host
// pointers) is captured once into a CUDA graph and replayed afterwards: a step is 13 launches per stream
// group, and at a few milliseconds per step the launch overhead of the host would otherwise show.
static int run_steps(BatchHost* bh, float dt, int vi, int pi, int steps, const float* host_forces, float* host_state_out, bool compact = false) {
  if (!bh || steps < 0 || vi < 0 || pi < 0) { set_error("batch_step: bad argument"); return B2GPU_E_INVALID; }
  Ctx* ctx = bh->ctx;
  (void)ctx;
  const StepParams sp = make_params(dt, vi, pi);
  bh->last_sp = sp;
#if defined(B2G_HOSTSIM)
  const int rc_sim = enqueue_steps(bh, sp, steps, host_forces, host_state_out, compact);
  if (rc_sim && rc_sim != B2GPU_E_CAPACITY && rc_sim != B2GPU_E_UNSUPPORTED && rc_sim != B2GPU_E_INTERNAL) return rc_sim;
#else
  RC(ensure_groups(bh));
  cudaStream_t main_s = (cudaStream_t)ctx->stream;
  const bool use_graph = bh->use_graphs && !bh->large && !ctx->profiling && steps > 0 && host_pointer_capturable(host_forces) &&
                         host_pointer_capturable(host_state_out);
  if (!use_graph) {
    RC(enqueue_steps(bh, sp, steps, host_forces, host_state_out, compact));
  } else {
    StepGraph* hit = nullptr;
    for (StepGraph& sgph : bh->graphs)
      if (sgph.dt == dt && sgph.vi == vi && sgph.pi == pi && sgph.steps == steps && sgph.forces == (const void*)host_forces &&
          sgph.state == (void*)host_state_out && sgph.compact == compact)
        hit = &sgph;
    if (!hit) {
      const long long launches_before = ctx->launches;
      CU(cudaStreamBeginCapture(main_s, cudaStreamCaptureModeThreadLocal));
      int rc = enqueue_steps(bh, sp, steps, host_forces, host_state_out, compact);
      cudaGraph_t graph = nullptr;
      cudaError_t e = cudaStreamEndCapture(main_s, &graph);
      if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
      if (e != cudaSuccess) return cuda_fail(e, "cudaStreamEndCapture");
      cudaGraphExec_t exec = nullptr;
      e = cudaGraphInstantiate(&exec, graph, 0);
      cudaGraphDestroy(graph);
      if (e != cudaSuccess) return cuda_fail(e, "cudaGraphInstantiate");
      if (bh->graphs.size() >= 8) {  // small cache: drop the oldest signature
        cudaGraphExecDestroy((cudaGraphExec_t)bh->graphs.front().exec);
        bh->graphs.erase(bh->graphs.begin());
      }
      StepGraph ng;
      ng.dt = dt; ng.vi = vi; ng.pi = pi; ng.steps = steps; ng.forces = host_forces; ng.state = host_state_out; ng.compact = compact;
      ng.exec = (void*)exec;
      ng.launches = ctx->launches - launches_before;
      ctx->launches = launches_before;  // capture issued nothing; replays are counted below
      bh->graphs.push_back(ng);
      hit = &bh->graphs.back();
    }
    CU(cudaGraphLaunch((cudaGraphExec_t)hit->exec, main_s));
    ctx->launches += hit->launches;
  }
  if (host_state_out) CU(cudaStreamSynchronize(main_s));
#endif
  bh->pre_step_needed = false;
  if (steps > 0 && dt > 0.0f) bh->stepped = true;
#if !defined(B2G_HOSTSIM)
  if (host_state_out) return status_error(*bh->status_host);
  return 0;
#else
  return rc_sim;  // a device status raised by a world (the state was still written)
#endif
}

int batch_step(BatchHost* bh, float dt, int vi, int pi, int steps) { return run_steps(bh, dt, vi, pi, steps, nullptr, nullptr); }

// touching / awake counters on demand: one thread per world
struct StatsK {
  Batch B;
  B2G_HD void operator()(int w) const {
    if (w >= B.n_worlds) return;
    WIdx x = widx(B, w);
    Ws ws = ws_of(B, x);
    int touching = 0, awake = 0;
    const int cc = ws[WS_CONTACT_COUNT];
    for (int c = 0; c < cc; ++c) touching += (B.c_flags[x.at(B.NC, c)] & B2GPU_CONTACT_TOUCHING) ? 1 : 0;
    for (int b = 0; b < B.NB; ++b) awake += (B.b_flags[x.at(B.NB, b)] & B2GPU_BODY_AWAKE) ? 1 : 0;
    ws[WS_ST_TOUCHING] = touching;
    ws[WS_ST_AWAKE] = awake;
    ws[WS_ST_CONTACTS] = cc;
  }
};

int batch_get_stats(BatchHost* bh, int first, int count, b2gpu_step_stats* out) {
  if (!bh || !out || first < 0 || count < 0 || first + count > bh->B.n_worlds) { set_error("get_stats: bad argument"); return B2GPU_E_INVALID; }
  Batch& B = bh->B;
  const int W = B.n_wblocks * B.LB;
  if (bh->large) {  // one world of 10^5 contacts: count flat
    RC(lw_read(bh, B.ws, WS_COUNT));
    const int cc = bh->lw_host[WS_CONTACT_COUNT];
    { LwStatsK k = {B, cc, 0}; RC(launch(bh->ctx, k, 1, 32)); }
    { LwStatsK k = {B, cc, 1}; RC(launch(bh->ctx, k, std::max(cc, B.NB), 256)); }
  } else {
    StatsK k = {B};
    RC(launch(bh->ctx, k, W, 32));
  }
  std::vector<int> all((size_t)W * WS_COUNT);
  RC(dev_d2h(bh->ctx, all.data(), B.ws, all.size() * 4));
  for (int i = 0; i < count; ++i) {
    const int w = first + i;
    const int wb = w >> B.lb_shift, wl = w & (B.LB - 1);
    auto get = [&](int slot) { return all[((size_t)wb * WS_COUNT + slot) * B.LB + wl]; };
    b2gpu_step_stats& s = out[i];
    memset(&s, 0, sizeof(s));
    s.status = get(WS_STATUS);
    s.contacts = get(WS_ST_CONTACTS); s.touching = get(WS_ST_TOUCHING); s.destroyed = get(WS_ST_DESTROYED);
    s.islands = get(WS_ST_ISLANDS); s.island_bodies = get(WS_ST_ISL_BODIES); s.island_contacts = get(WS_ST_ISL_CONTACTS);
    s.moved = get(WS_ST_MOVED); s.pairs = get(WS_ST_PAIRS); s.created = get(WS_ST_CREATED); s.awake_bodies = get(WS_ST_AWAKE);
    s.solver_levels = get(WS_ST_LEVELS);
#if defined(B2G_LV_DEBUG)
    if (bh->large) { int dbg[4]; dev_d2h(bh->ctx, dbg, bh->L.lv_meta, 16); s.reserved[0] = dbg[1]; s.reserved[1] = dbg[2]; s.reserved[2] = dbg[3]; }
#endif
  }
  return 0;
}

// Most negative WS_STATUS of any world of the batch (0: none).  Synchronises.
int batch_status(BatchHost* bh, int* out) {
  if (!bh || !out) { set_error("batch_status: bad argument"); return B2GPU_E_INVALID; }
  Batch all = bh->B;
  all.wb_first = 0;
  all.wb_count = bh->B.n_wblocks;
  { StatusK k = {all, bh->status_dev}; RC(launch(bh->ctx, k, all.n_worlds, 128)); }
  RC(dev_d2h(bh->ctx, bh->status_host, bh->status_dev, 4));
  *out = *bh->status_host;
  return 0;
}
int batch_get_body_state(BatchHost* bh, float* host_out, int first, int count) {
  if (!bh || !host_out || first < 0 || count < 0 || first + count > bh->B.n_worlds) { set_error("get_body_state: bad argument"); return B2GPU_E_INVALID; }
  Batch& B = bh->B;
  { StateGatherK k = {B, bh->state_dev}; RC(launch(bh->ctx, k, B.n_wblocks * B.LB * B.NB, 128)); }
  RC(dev_d2h(bh->ctx, host_out, bh->state_dev + (size_t)first * B.NB * 8, (size_t)count * B.NB * 8 * 4));
  int st = 0;
  RC(batch_status(bh, &st));
  return status_error(st);
}
int batch_set_forces(BatchHost* bh, const float* host, int first, int count) {
  if (!bh || !host || first < 0 || count < 0 || first + count > bh->B.n_worlds) { set_error("set_forces: bad argument"); return B2GPU_E_INVALID; }
  Batch& B = bh->B;
  RC(dev_h2d(bh->ctx, bh->forces_dev, host, (size_t)count * B.NB * 3 * 4));
  { ForceScatterK k = {B, bh->forces_dev, first, count}; RC(launch(bh->ctx, k, B.n_wblocks * B.LB * B.NB, 128)); }
  return 0;
}
int batch_set_linear_velocity(BatchHost* bh, int body, const float* host_vxvy, int first, int count) {
  if (!bh || !host_vxvy || body < 0 || body >= bh->B.NB || first < 0 || count < 0 || first + count > bh->B.n_worlds) {
    set_error("set_linear_velocity: bad argument");
    return B2GPU_E_INVALID;
  }
  RC(dev_h2d(bh->ctx, bh->vel_scratch, host_vxvy, (size_t)count * 2 * 4));  // its own scratch: forces_dev belongs to the caller
  { VelScatterK k = {bh->B, bh->vel_scratch, body, first, count}; RC(launch(bh->ctx, k, count, 128)); }
  return 0;
}
int batch_set_gravity(BatchHost* bh, const float* host_gxgy, int first, int count) {
  if (!bh || !host_gxgy || first < 0 || count < 0 || first + count > bh->B.n_worlds) { set_error("set_gravity: bad argument"); return B2GPU_E_INVALID; }
  if (count == 0) return 0;
  RC(dev_h2d(bh->ctx, bh->vel_scratch, host_gxgy, (size_t)count * 2 * 4));
  Batch all = bh->B;
  all.wb_first = 0;
  all.wb_count = bh->B.n_wblocks;
  { GravityScatterK k = {all, bh->vel_scratch, first, count}; RC(launch(bh->ctx, k, count, 128)); }
  return 0;
}
int batch_set_joint_control(BatchHost* bh, int joint, int control, const float* host_values, int first, int count) {
  if (!bh || !host_values || joint < 0 || joint >= bh->B.NJ || first < 0 || count < 0 || first + count > bh->B.n_worlds) {
    set_error("set_joint_control: bad argument");
    return B2GPU_E_INVALID;
  }
  const int type = bh->topo.joints[joint].type;
  const bool motorised = type == B2GPU_JOINT_REVOLUTE || type == B2GPU_JOINT_PRISMATIC || type == B2GPU_JOINT_WHEEL;
  const bool ok = control == B2GPU_JOINT_CONTROL_TARGET ? type == B2GPU_JOINT_MOUSE
                  : (control == B2GPU_JOINT_CONTROL_MOTOR_SPEED || control == B2GPU_JOINT_CONTROL_MAX_MOTOR_TORQUE) && motorised;
  if (!ok) { set_error("set_joint_control: the joint is not of a type this control edits"); return B2GPU_E_INVALID; }
  if (count == 0) return 0;
  const int per = control == B2GPU_JOINT_CONTROL_TARGET ? 2 : 1;
  RC(dev_h2d(bh->ctx, bh->vel_scratch, host_values, (size_t)count * per * 4));  // [n_worlds][2] staging
  Batch all = bh->B;
  all.wb_first = 0;
  all.wb_count = bh->B.n_wblocks;
  { JointControlK k = {all, bh->vel_scratch, joint, control, first, count}; RC(launch(bh->ctx, k, count, 128)); }
  return 0;
}
// The device-pointer forms (zero-copy consumers, e.g. torch tensors over b2gpu_batch_forces_device /
// b2gpu_batch_body_state_device): scatter the caller-written force buffer into the worlds / refresh the state buffer.
// Asynchronous on the context stream.
int batch_apply_device_forces(BatchHost* bh) {
  if (!bh) { set_error("apply_device_forces: bad argument"); return B2GPU_E_INVALID; }
  Batch all = bh->B;
  all.wb_first = 0;
  all.wb_count = bh->B.n_wblocks;
  ForceScatterK k = {all, bh->forces_dev, 0, all.n_worlds};
  return launch(bh->ctx, k, all.n_wblocks * all.LB * all.NB, 128);
}
int batch_refresh_device_state(BatchHost* bh) {
  if (!bh) { set_error("refresh_device_state: bad argument"); return B2GPU_E_INVALID; }
  Batch all = bh->B;
  all.wb_first = 0;
  all.wb_count = bh->B.n_wblocks;
  StateGatherK k = {all, bh->state_dev};
  return launch(bh->ctx, k, all.n_wblocks * all.LB * all.NB, 128);
}

// One end-to-end call through HOST buffers: forces H2D, `steps` steps, body state D2H.  Synchronous:
// returns when the state is in `host_state_out`.
int batch_step_host(BatchHost* bh, const float* host_forces, float* host_state_out, float dt, int vi, int pi, int steps) {
  RC(run_steps(bh, dt, vi, pi, steps, host_forces, host_state_out));
  if (!host_state_out) RC(ctx_sync(bh->ctx));
  return 0;
}

int batch_step_host_dynamic(BatchHost* bh, const float* host_forces, float* host_state_out, float dt, int vi, int pi, int steps) {
  if (!bh || bh->n_dyn <= 0) { set_error("step_host_dynamic: the prototype has no dynamic body"); return B2GPU_E_INVALID; }
  RC(run_steps(bh, dt, vi, pi, steps, host_forces, host_state_out, true));
  if (!host_state_out) RC(ctx_sync(bh->ctx));
  return 0;
}
int batch_dynamic_bodies(BatchHost* bh, int* out, int capacity) {
  if (!bh || capacity < 0 || (capacity > 0 && !out)) { set_error("dynamic_bodies: bad argument"); return B2GPU_E_INVALID; }
  int n = 0;
  for (int b = 0; b < bh->B.NB; ++b)
    if (bh->topo.bodies[b].type == B2GPU_DYNAMIC_BODY) { if (n < capacity) out[n] = b; ++n; }
  return n;
}

// post_solve reports of the last step of one world: island contact slot k -> (fixtures, children, impulses)
struct PostSolveGatherK {
  Batch B;
  b2gpu_post_solve_event* out;
  int w;
  B2G_HD void operator()(int k) const {
    WIdx x = widx(B, w);
    Ws ws = ws_of(B, x);
    if (k >= ws[WS_ISL_CONTACTS]) return;
    const float4 q6 = B.vc[vc_at(B, x, k, 6)], q8 = B.vc[vc_at(B, x, k, 8)];
    const int4 fx = B.c_fix[x.at(B.NC, f2i(q8.w))];
    b2gpu_post_solve_event e;
    e.fixture_a = fx.x; e.fixture_b = fx.y; e.index_a = fx.z; e.index_b = fx.w;
    e.count = f2i(q8.z) & 0xff;
    e.normal_impulses[0] = e.count > 0 ? q6.x : 0.0f; e.tangent_impulses[0] = e.count > 0 ? q6.y : 0.0f;
    e.normal_impulses[1] = e.count > 1 ? q6.z : 0.0f; e.tangent_impulses[1] = e.count > 1 ? q6.w : 0.0f;
    e.reserved[0] = e.reserved[1] = e.reserved[2] = 0;
    out[k] = e;
  }
};
int batch_post_solve_events(BatchHost* bh, int world, b2gpu_post_solve_event* out, int capacity) {
  if (!bh || world < 0 || world >= bh->B.n_worlds || capacity < 0 || (capacity > 0 && !out)) { set_error("post_solve_events: bad argument"); return B2GPU_E_INVALID; }
  if (!bh->stepped) return 0;
  Batch all = bh->B;
  all.wb_first = 0;
  all.wb_count = bh->B.n_wblocks;
  WorldImage im;
  image_alloc(bh->B, im);
  ArrRef a = {bh->B.ws, im.ws.data(), 1, WS_COUNT};
  RC(move_array(bh, a, 1, world));
  if (!im.ws[WS_ISL_VALID]) return 0;  // the last step did not solve (dt = 0)
  const int n = im.ws[WS_ISL_CONTACTS];
  if (n == 0 || capacity == 0) return n;
  char* base = nullptr;
  RC(query_scratch(bh, (size_t)n * sizeof(b2gpu_post_solve_event), &base));
  { PostSolveGatherK k = {all, (b2gpu_post_solve_event*)base, world}; RC(launch(bh->ctx, k, n, 128)); }
  RC(dev_d2h(bh->ctx, out, base, (size_t)std::min(n, capacity) * sizeof(b2gpu_post_solve_event)));
  return n;
}

// sin/cos of an array of angles on the device (diagnostic: pins rot_from_angle against libm)
struct SinCosK {
  const float* in;
  float* s;
  float* c;
  B2G_HD void operator()(int i) const { sincos_ref(in[i], &s[i], &c[i]); }
};
int debug_sincos(Ctx* ctx, const float* host_in, float* host_sin, float* host_cos, int n) {
  if (!ctx || !host_in || !host_sin || !host_cos || n < 0) { set_error("debug_sincos: bad argument"); return B2GPU_E_INVALID; }
  void* d = nullptr;
  RC(dev_alloc(&d, (size_t)n * 12));
  float* din = (float*)d;
  int rc = dev_h2d(ctx, din, host_in, (size_t)n * 4);
  if (!rc) { SinCosK k = {din, din + n, din + 2 * (size_t)n}; rc = launch(ctx, k, n, 256); }
  if (!rc) rc = dev_d2h(ctx, host_sin, din + n, (size_t)n * 4);
  if (!rc) rc = dev_d2h(ctx, host_cos, din + 2 * (size_t)n, (size_t)n * 4);
  dev_free(d);
  return rc;
}

// compulsory HBM traffic of the last step, summed over all worlds
long long batch_algorithmic_bytes(BatchHost* bh) {
  if (!bh) return -1;
  Batch& B = bh->B;
  std::vector<b2gpu_step_stats> st(B.n_worlds);
  if (batch_get_stats(bh, 0, B.n_worlds, st.data())) return -1;
  long long total = 0;
  for (const b2gpu_step_stats& s : st)
    total += 116LL * s.awake_bodies + 36LL * B.NP + 16LL * s.moved + 272LL * s.contacts + 168LL * s.created;
  return total;
}

}  // namespace b2g
