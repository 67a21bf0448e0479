// b2g_common.h — device-side data layout of a batch of worlds.
//
// HBM layout ("blocked world-minor"): the worlds of a batch are grouped in blocks of LB worlds
// (LB = 32 for batches, 1 for a single large world).  Element i of a per-world array with capacity
// N lives at ((w / LB) * N + i) * LB + (w % LB).  Consequences:
//   * a warp whose lanes are 32 consecutive worlds touching the same element reads one contiguous
//     128 B (float) / 512 B (float4) segment — both for the flat (element, world) kernels and for
//     the world-per-lane sequential kernels (island order, Gauss-Seidel sweeps, tree updates);
//   * consecutive elements of one world block are contiguous, so the sequential kernels stream;
//   * with LB = 1 every array degenerates to a plain SoA array of one world.
// Topology that the reference only changes outside the step (fixtures, shapes, proxy <-> fixture
// mapping, body types) is stored once per batch and shared by all worlds.
#pragma once
#include "../../include/b2gpu.h"
#include "b2g_math.h"

#if !defined(__CUDACC__)
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
static inline float2 make_float2(float x, float y) { float2 r = {x, y}; return r; }
static inline int4 make_int4(int x, int y, int z, int w) { int4 r = {x, y, z, w}; return r; }
static inline int2 make_int2(int x, int y) { int2 r = {x, y}; return r; }
#endif

namespace b2g {

// per-world scalar slots (int; floats are bit-cast)
enum {
  WS_GRAVITY_X = 0, WS_GRAVITY_Y, WS_INV_DT0, WS_FLAGS,
  WS_TREE_ROOT, WS_TREE_FREE, WS_TREE_COUNT, WS_TREE_CAP, WS_TREE_INSERTIONS, WS_PROXY_COUNT,
  WS_CONTACT_COUNT, WS_MOVE_COUNT,
  WS_ISL_COUNT, WS_ISL_BODIES, WS_ISL_CONTACTS, WS_ISL_JOINTS,
  WS_EV_WAKE,      // collide woke a sleeping body (needs the ordered fix-up pass)
  WS_EV_DESTROY,   // contacts flagged for destruction this step
  WS_EV_MOVED,     // proxies that left their fat box this step
  WS_TOPO_DIRTY,   // island order must be rebuilt
  WS_ISL_VALID,    // island arrays describe the last dt > 0 step of this world
  WS_SCHED_ROUNDS, // rounds of the level schedule of the island contacts (-1: none, solve in list order)
  WS_STATUS,
  // stats of the last step (b2gpu_step_stats order from `contacts` on)
  WS_ST_CONTACTS, WS_ST_TOUCHING, WS_ST_DESTROYED, WS_ST_ISLANDS, WS_ST_ISL_BODIES, WS_ST_ISL_CONTACTS,
  WS_ST_MOVED, WS_ST_PAIRS, WS_ST_CREATED, WS_ST_AWAKE, WS_ST_LEVELS,
  WS_COUNT
};

// velocity-constraint record: VC_Q float4 per island contact
enum { VC_Q = 9, PC_Q = 6 };
// per-step joint scratch: JT_Q float4 per joint (see b2g_joint.h)
enum { JT_Q = 5 };

struct Batch {
  int n_worlds, LB, lb_shift, n_wblocks;
  int wb_first, wb_count;    // window of world blocks a launch works on (stream groups); default: all
  int NB, NF, NS, NP;        // bodies, fixtures, child shapes, proxies (exact, shared topology)
  int NN;                    // tree node pool (physical capacity)
  int NC, NMOVE;             // capacities: contacts, move buffer
  int NIB;                   // island body list capacity (NB + NC: static bodies repeat per island)
  int NMW;                   // words of the per-world moved-proxy bitmap ((NP + 31) / 32)
  int NJ;                    // joints (exact, shared topology; 0 for most scenes)
  // ---- shared topology
  const b2gpu_fixture_rec* fixtures;
  const b2gpu_shape_rec* shapes;
  const int4* proxy_s;       // {fixture, child index, tree node id, body}
  const int* sync_order;     // proxies in synchronize_fixtures order (bodies newest first, fixtures newest first)
  const int* sync_rank;      // inverse of sync_order
  const int* node_proxy;     // tree node id -> proxy index (-1 for internal / unused nodes)
  const b2gpu_joint_rec* joints;  // static part of the joint table: type, bodies, COLLIDE_CONNECTED, anchors, limits, lengths
  const int* jadj_off;       // [NB+1] per body: its joint edges (2*joint + side) in the order the reference's list iterates (newest first)
  const int* jadj;           // [2*NJ]
  // ---- per world (blocked world-minor)
  int* ws;                   // [WS_COUNT]
  int* b_flags;              // BodyFlags | type << 16
  float4* b_xf;              // p.x p.y q.s q.c
  float4* b_pos;             // c.x c.y a sleep_time
  float4* b_pos0;            // c0.x c0.y a0 -
  float4* b_vel;             // v.x v.y w -
  float4* b_mass;            // inv_mass inv_I lc.x lc.y
  float4* b_force;           // f.x f.y torque gravity_scale
  float4* b_misc;            // mass I linear_damping angular_damping
  float4* n_aabb;            // tree: fat AABB lo.x lo.y hi.x hi.y
  int4* n_link;              // parent(or next free) child1 child2 height
  int* n_moved;
  float4* p_aabb;            // proxy tight AABB
  int* move_buf;
  int4* c_fix;               // fixture_a fixture_b index_a index_b
  int* c_flags;
  float4* c_mat;             // friction restitution threshold tangent_speed
  float4* c_m0;              // point0: lp.x lp.y normal_impulse tangent_impulse
  float4* c_m1;              // point1
  float4* c_m2;              // local_normal.xy local_point.xy
  int4* c_m3;                // id0 id1 type point_count
  float4* j_s0;              // joint: impulse.x impulse.y motor_impulse lower_impulse   (distance: impulse - - lower)
  float4* j_s1;              // joint: upper_impulse motor_speed max_motor_torque (int) ENABLE_LIMIT | ENABLE_MOTOR bits
  // ---- per-step scratch
  float4* b_rot;             // sin(a) cos(a) of the running angle (position pass cache), spare, spare
  float4* p_fat;             // new fat AABB of a proxy that must be re-inserted
  int* p_move;               // [NMW] bitmap over synchronize ranks: proxies that left their fat box
  int* adj_off;              // [NB+1] CSR of eligible contacts per body (newest first)
  int* adj;                  // [2*NC]
  int* isl_body;             // [NIB] island body order
  int* isl_contact;          // [NC] island contact order
  int4* isl_range;           // [NB] per island: body_first, body_end, contact_first, contact_end
  int* isl_flags;            // [NB] per island: bit0 = position solved
  int* c_isl;                // [NC] island index of each island contact slot
  int* isl_joint;            // [NJ] island joint order
  int2* isl_jrange;          // [NB] per island: joint_first, joint_end
  int* j_flag;               // [NJ] m_island_flag of the island DFS
  float4* j_tmp;             // [NJ * JT_Q] per-step joint solver data (b2g_joint.h)
  float4* vc;                // [NC * VC_Q] velocity constraint records
  float4* pc;                // [NC * PC_Q] position constraint records
  unsigned long long* timeline;  // diagnostic (B2GPU_TIMELINE): [0] = entries used, [1] = capacity, then {kind<<32|cta, t0, t1} per CTA
};

struct WIdx {  // index helper of one thread's world
  int wb, wl, LB;
  B2G_HD int at(int N, int i) const { return (wb * N + i) * LB + wl; }
};
B2G_HD WIdx widx(const Batch& B, int w) {
  WIdx x;
  x.wb = w >> B.lb_shift;
  x.wl = w & (B.LB - 1);
  x.LB = B.LB;
  return x;
}
B2G_HD float i2f(int i) {
#if defined(__CUDA_ARCH__)
  return __int_as_float(i);
#else
  float f; memcpy(&f, &i, 4); return f;
#endif
}
B2G_HD int f2i(float f) { return (int)f2u(f); }

B2G_HD int body_type(int flags) { return (flags >> 16) & 0xff; }
B2G_HD int lowest_bit(unsigned v) {
#if defined(__CUDA_ARCH__)
  return __ffs((int)v) - 1;
#else
  return __builtin_ctz(v);
#endif
}

struct StepParams {
  float dt, inv_dt;
  int velocity_iterations, position_iterations;
};

}  // namespace b2g
