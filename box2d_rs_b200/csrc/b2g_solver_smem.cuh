// b2g_solver_smem.cuh — the two Gauss-Seidel stages for batches (LB == 32), sm_100a.
//
// The sweeps of b2_contact_solver_private.rs (:228 warm_start, :268 solve_velocity_constraints,
// :653 solve_position_constraints) are sequential by definition inside one world, and the exact
// dependency DAG of the reference order is almost a chain (Pyramid: 1771 levels for 3200 constraint
// solves).  The parallelism is the batch: one lane per world, one warp (32 worlds = one memory block)
// per CTA, so a CTA's working set is
//   * the 32 worlds' body velocities (positions) in shared memory, laid out [body][component][lane]:
//     every lane hits its own bank whatever body its world is touching (conflict-free gathers);
//   * the constraint stream of those 32 worlds, which SolverInitK wrote k-major per world block, so
//     "constraint k of all 32 worlds" is one contiguous 4.6 KB (velocity) / 3 KB (position) segment.
//     It is staged through a shared-memory ring with cp.async (LDGSTS, 16 B per lane per row), D-1
//     constraints ahead of the solve, so HBM/L2 latency never sits on the dependent chain.
// What remains on the critical path of a lane is the reference's own arithmetic chain (about 60
// dependent fp32 operations per two-point constraint), shared-memory reads of two bodies, and their
// write-back.  Results are bit-identical to the generic stages in b2g_step.h (same device functions).
#pragma once
#include "b2g_step.h"
#include <type_traits>

namespace b2g {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// Diagnostic timeline (env B2GPU_TIMELINE=<file>): every CTA of the two Gauss-Seidel kernels records its start
// and end on the global nanosecond timer, so the overlap of the stream groups can be reconstructed.
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void timeline_record(const Batch& B, int kind, unsigned long long t0) {
  if (B.timeline && threadIdx.x == 0) {
    const unsigned long long slot = atomicAdd(B.timeline, 1ull);
    if (slot < B.timeline[1]) {
      B.timeline[2 + slot * 3] = ((unsigned long long)kind << 32) | (unsigned)(blockIdx.x + (B.wb_first << 8));
      B.timeline[3 + slot * 3] = t0;
      B.timeline[4 + slot * 3] = global_ns();
    }
  }
}

constexpr int VEL_RING = 8;  // stages of the velocity constraint ring
constexpr int POS_RING = 8;

inline size_t velocity_smem_bytes(int NB) { return (size_t)(NB + 1) * 32 * 16 + (size_t)VEL_RING * VC_Q * 32 * 16 + VEL_RING * 8; }
inline size_t position_smem_bytes(int NB) { return (size_t)NB * 32 * (16 + 8) + (size_t)POS_RING * PC_Q * 32 * 16; }

// Body-state accessors of the joint rows (b2g_joint.h) over this file's [body][lane] shared-memory rows.
struct BodyStateVelSmem {  // velocity stage: velocities in shared memory, positions / rotations read-only in HBM
  const Batch& B;
  const WIdx& x;
  float4* vl;  // this lane's column of the velocity rows
  __device__ __forceinline__ float4 vel(int b) const { return vl[b * 32]; }
  __device__ __forceinline__ void set_vel(int b, float4 v) const { vl[b * 32] = v; }
  __device__ __forceinline__ float4 pos(int b) const { return B.b_pos[x.at(B.NB, b)]; }
  __device__ __forceinline__ Rot rot(int b) const { const float4 r = B.b_rot[x.at(B.NB, b)]; Rot q; q.s = r.x; q.c = r.y; return q; }
  __device__ __forceinline__ void set_pos(int, float4) const {}
  __device__ __forceinline__ void set_rot(int, Rot) const {}
};
struct BodyStatePosSmem {  // position stage: positions and cached rotations in shared memory
  float4* pl;
  float2* ql;
  __device__ __forceinline__ float4 vel(int) const { return make_float4(0.0f, 0.0f, 0.0f, 0.0f); }
  __device__ __forceinline__ void set_vel(int, float4) const {}
  __device__ __forceinline__ float4 pos(int b) const { return pl[b * 32]; }
  __device__ __forceinline__ void set_pos(int b, float4 p) const { pl[b * 32] = p; }
  __device__ __forceinline__ Rot rot(int b) const { const float2 r = ql[b * 32]; Rot q; q.s = r.x; q.c = r.y; return q; }
  __device__ __forceinline__ void set_rot(int b, Rot q) const { ql[b * 32] = make_float2(q.s, q.c); }
};

struct VcRegs {  // one velocity constraint of one world, in registers
  float4 q0, q1, q2, q3, q4, q5, q6, q7;
  int ba, bb, cnt;
};
__device__ __forceinline__ VcRegs vc_load(const float4* st) {
  VcRegs r;
  r.q0 = st[0 * 32]; r.q1 = st[1 * 32]; r.q2 = st[2 * 32]; r.q3 = st[3 * 32];
  r.q4 = st[4 * 32]; r.q5 = st[5 * 32]; r.q6 = st[6 * 32]; r.q7 = st[7 * 32];
  const float4 q8 = st[8 * 32];
  r.ba = __float_as_int(q8.x);
  r.bb = __float_as_int(q8.y);
  r.cnt = __float_as_int(q8.z) & 0xff;
  return r;
}

// Resident form of the velocity stage: the whole constraint stream of the block fits the ring; load it once
// and iterate in shared memory (small islands).
template <class JointInit, class JointSolve>
__device__ __forceinline__ void velocity_resident(float4* rl, float4* vl, const float4* src, float4* q6_out, int nc, int ncm,
                                                  bool warm, bool block, int sweeps, const JointInit& joint_init,
                                                  const JointSolve& joint_solve) {
  for (int k = 0; k < ncm; ++k) {
#pragma unroll
    for (int q = 0; q < VC_Q; ++q) cp_async16(rl + (k * VC_Q + q) * 32, src + (size_t)(k * VC_Q + q) * 32);
  }
  cp_async_commit();
  cp_async_wait<0>();
  for (int sweep = 0; sweep < sweeps; ++sweep) {
    // joints: init_velocity_constraints after the contact warm start, then before the contacts in every iteration
    if (sweep == 1) joint_init();
    if (sweep >= 1) joint_solve();
    if (sweep == 0 && !__any_sync(0xffffffffu, warm)) continue;
    for (int k = 0; k < ncm; ++k) {
      if (k >= nc || (sweep == 0 && !warm)) continue;
      float4* st = rl + (k * VC_Q) * 32;
      VcRegs c = vc_load(st);
      if (c.cnt == 0) continue;
      const float4 va = vl[c.ba * 32], vb = vl[c.bb * 32];
      VelState s;
      s.v_a = v2(va.x, va.y); s.w_a = va.z;
      s.v_b = v2(vb.x, vb.y); s.w_b = vb.z;
      if (sweep == 0) {
        warm_start_one(s, c.q0, c.q1, c.q2, c.q6, c.q7, c.cnt);
      } else {
        solve_velocity_one(s, c.q0, c.q1, c.q2, c.q3, c.q4, c.q5, c.q6, c.q7, c.cnt, block);
        st[6 * 32] = c.q6;
        if (sweep == sweeps - 1) q6_out[(size_t)k * VC_Q * 32] = c.q6;
      }
      vl[c.ba * 32] = make_float4(s.v_a.x, s.v_a.y, s.w_a, 0.0f);
      vl[c.bb * 32] = make_float4(s.v_b.x, s.v_b.y, s.w_b, 0.0f);
    }
  }
  if (sweeps == 1) joint_init();  // no velocity iterations: the joints' warm start still applies (b2_island_private.rs:198-201)
}

// ------------------------------------------------------------------------------------------
// Straight-line form of the same stage (the default).  The pipelined loop above spends 40 % of a visit in
// bookkeeping around the arithmetic (ncu source page, round 1: 92 of 280 instructions at 3 cycles each:
// convergence barriers of five data-dependent branches, the ring refill, register forwarding), none of which
// can overlap the dependent fp32 chain because every branch ends a scheduling region.  Here one visit is ONE
// basic block: the refill is unconditional (the stream wraps, so reading ahead past the end is harmless), an
// inactive visit (world with fewer constraints, empty manifold, no warm start) runs the same arithmetic on a
// scratch row instead of being predicated, selects are kept from becoming branches (sel_mask), the warm-start
// sweep has its own loop, and the choice between the two-point block solver and the general path is a warp
// vote taken one visit ahead.  ptxas can
// then fill the issue slots the chain leaves empty with the next visit's loads.
// ------------------------------------------------------------------------------------------
static_assert(VC_Q == 9, "cp_async_record copies nine rows");
__device__ __forceinline__ void cp_async_record(float4* smem_dst, const float4* gmem_src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile(
      "cp.async.cg.shared.global [%0], [%1], 16;\n"
      "cp.async.cg.shared.global [%0+512], [%1+512], 16;\n"
      "cp.async.cg.shared.global [%0+1024], [%1+1024], 16;\n"
      "cp.async.cg.shared.global [%0+1536], [%1+1536], 16;\n"
      "cp.async.cg.shared.global [%0+2048], [%1+2048], 16;\n"
      "cp.async.cg.shared.global [%0+2560], [%1+2560], 16;\n"
      "cp.async.cg.shared.global [%0+3072], [%1+3072], 16;\n"
      "cp.async.cg.shared.global [%0+3584], [%1+3584], 16;\n"
      "cp.async.cg.shared.global [%0+4096], [%1+4096], 16;\n"
      "cp.async.commit_group;\n" ::"r"(s),
      "l"(gmem_src)
      : "memory");
}

__global__ void __launch_bounds__(32) velocity_sl_kernel(const Batch B, const StepParams sp) {
  extern __shared__ float4 smem4[];
  const unsigned long long t_start = B.timeline ? global_ns() : 0ull;
  float4* ring = smem4;                        // [VEL_RING][VC_Q][32]
  float4* vel = smem4 + VEL_RING * VC_Q * 32;  // [NB][32]: v.x v.y w -
  const int lane = threadIdx.x;
  const int wb = blockIdx.x + B.wb_first;
  const int w = wb * 32 + lane;
  const bool live = w < B.n_worlds;
  WIdx x;
  x.wb = wb; x.wl = lane; x.LB = 32;
  Ws ws = ws_of(B, x);
  const int nc = live ? ws[WS_ISL_CONTACTS] : 0;
  const int wflags = live ? ws[WS_FLAGS] : 0;
  const bool warm = (wflags & B2GPU_WORLD_WARM_STARTING) != 0;
  const bool block = (wflags & B2GPU_WORLD_BLOCK_SOLVE) != 0;
  const int ncm = __reduce_max_sync(0xffffffffu, nc);
  // joints of this lane's world, all islands concatenated in island order: islands share no movable body, so one pass
  // over the concatenation equals the reference's per-island passes
  const int nj = (live && B.NJ > 0) ? ws[WS_ISL_JOINTS] : 0;
  const int njm = B.NJ > 0 ? __reduce_max_sync(0xffffffffu, nj) : 0;
  if (ncm == 0 && njm == 0) return;
  if (live)
    for (int b = 0; b < B.NB; ++b) vel[b * 32 + lane] = B.b_vel[x.at(B.NB, b)];
  const float4* src = B.vc + (size_t)wb * B.NC * VC_Q * 32 + lane;  // + (k * VC_Q + q) * 32
  float4* q6_out = B.vc + (size_t)wb * B.NC * VC_Q * 32 + 6 * 32 + lane;
  float4* vl = vel + lane;
  float4* rl = ring + lane;
  const BodyStateVelSmem jst = {B, x, vl};
  const float dt_ratio = live ? i2f(ws[WS_INV_DT0]) * sp.dt : 0.0f;
  auto joint_init = [&]() {
    for (int q = 0; q < nj; ++q) joint_init_velocity(B, x, jst, B.isl_joint[x.at(B.NJ, q)], warm, dt_ratio, sp.dt);
  };
  auto joint_solve = [&]() {
    for (int q = 0; q < nj; ++q) joint_solve_velocity(B, x, jst, B.isl_joint[x.at(B.NJ, q)], sp.dt, sp.inv_dt);
  };
  if (ncm <= VEL_RING) {
    velocity_resident(rl, vl, src, q6_out, nc, ncm, warm, block, 1 + sp.velocity_iterations, joint_init, joint_solve);
  } else {
    const int n_warm = __any_sync(0xffffffffu, warm) ? ncm : 0;  // positions of the warm-start sweep
    const int total = n_warm + sp.velocity_iterations * ncm;
    int fk = 0;
    const float4* fsrc = src;
    auto fetch = [&](int p) {  // next record of the (wrapping) stream -> stage p % RING
      cp_async_record(rl + ((p & (VEL_RING - 1)) * VC_Q) * 32, fsrc);
      const bool wrap = (fk + 1 == ncm);
      fk = wrap ? 0 : fk + 1;
      fsrc = wrap ? src : fsrc + VC_Q * 32;
    };
#pragma unroll
    for (int p = 0; p < VEL_RING; ++p) fetch(p);
    cp_async_wait<VEL_RING - 1>();
    // An inactive visit (world with fewer constraints, empty manifold, warm-start sweep of a world without
    // warm starting) runs the same arithmetic on row NB of the velocity array, a scratch row no body owns, so
    // the visit needs no predicate: its loads, stores and forwards only ever touch that row, and the impulses
    // it writes belong to a record nobody reads (k >= nc, or zero points).
    const int scratch = B.NB;
    VcRegs ca = vc_load(rl), cb;
    int k = 0, pos = 0;
    const bool act0 = nc > 0 && ca.cnt > 0 && (warm || n_warm == 0);
    bool fa = __all_sync(0xffffffffu, !act0 || (ca.cnt == 2 && block)), fb = false;
    if (!act0) { ca.ba = scratch; ca.bb = scratch; }
    float4 vaa = vl[ca.ba * 32], vab = vl[ca.bb * 32], vba = vaa, vbb = vab;
    cb = ca;
    auto half = [&](auto WARM, auto FAST, VcRegs& cur, float4& va, float4& vb, VcRegs& nxt, bool& nfast, float4& nva,
                    float4& nvb) {
      // -- position pos+1: rows from its ring stage and the velocities of its two bodies, into the other register set
      const int kc = k;
      k = (k + 1 == ncm) ? 0 : k + 1;
      cp_async_wait<VEL_RING - 2>();
      nxt = vc_load(rl + (((pos + 1) & (VEL_RING - 1)) * VC_Q) * 32);
      const bool nact = (k < nc) && nxt.cnt > 0 && (warm || pos + 1 >= n_warm);
      nfast = __all_sync(0xffffffffu, !nact || (nxt.cnt == 2 && block));
      nxt.ba = nact ? nxt.ba : scratch;
      nxt.bb = nact ? nxt.bb : scratch;
      nva = vl[nxt.ba * 32];
      nvb = vl[nxt.bb * 32];
      fetch(pos);  // this position's stage is free (cur is in registers): refill it with position pos + RING
      // -- position pos: the reference's arithmetic
      VelState s;
      s.v_a = v2(va.x, va.y); s.w_a = va.z;
      s.v_b = v2(vb.x, vb.y); s.w_b = vb.z;
      if (decltype(WARM)::value) {
        warm_start_one(s, cur.q0, cur.q1, cur.q2, cur.q6, cur.q7, decltype(FAST)::value ? 2 : cur.cnt);
      } else {
        float4 q6 = cur.q6;
        if (decltype(FAST)::value)
          solve_velocity_one(s, cur.q0, cur.q1, cur.q2, cur.q3, cur.q4, cur.q5, q6, cur.q7, 2, true);
        else
          solve_velocity_one(s, cur.q0, cur.q1, cur.q2, cur.q3, cur.q4, cur.q5, q6, cur.q7, cur.cnt, block);
        q6_out[(size_t)kc * VC_Q * 32] = q6;
      }
      // the fourth component is carried so that its register stays owned by this value (a scratch reuse would
      // wait for the shared-memory load that also writes it)
      va = make_float4(s.v_a.x, s.v_a.y, s.w_a, va.w);
      vb = make_float4(s.v_b.x, s.v_b.y, s.w_b, vb.w);
      vl[cur.ba * 32] = va;
      vl[cur.bb * 32] = vb;
      // -- forward the fresh velocities to the next constraint where it shares a body with this one (the loads
      //    above were issued before these stores)
      const int maa = sel_mask(nxt.ba == cur.ba), mab = sel_mask(nxt.ba == cur.bb);
      const int mba = sel_mask(nxt.bb == cur.ba), mbb = sel_mask(nxt.bb == cur.bb);
      nva.x = msel(maa, va.x, msel(mab, vb.x, nva.x));
      nva.y = msel(maa, va.y, msel(mab, vb.y, nva.y));
      nva.z = msel(maa, va.z, msel(mab, vb.z, nva.z));
      nvb.x = msel(mba, va.x, msel(mbb, vb.x, nvb.x));
      nvb.y = msel(mba, va.y, msel(mbb, vb.y, nvb.y));
      nvb.z = msel(mba, va.z, msel(mbb, vb.z, nvb.z));
      ++pos;
    };
    auto step_ab = [&](auto WARM) {
      if (fa) half(WARM, std::true_type{}, ca, vaa, vab, cb, fb, vba, vbb);
      else half(WARM, std::false_type{}, ca, vaa, vab, cb, fb, vba, vbb);
    };
    auto step_ba = [&](auto WARM) {
      if (fb) half(WARM, std::true_type{}, cb, vba, vbb, ca, fa, vaa, vab);
      else half(WARM, std::false_type{}, cb, vba, vbb, ca, fa, vaa, vab);
    };
    auto run_to = [&](auto WARM, const int end) {
      while (pos + 2 <= end) {
        step_ab(WARM);
        step_ba(WARM);
      }
      if (pos < end) {  // odd count: one more visit, then the register sets swap roles
        step_ab(WARM);
        ca = cb; fa = fb; vaa = vba; vab = vbb;
      }
    };
    run_to(std::true_type{}, n_warm);
    if (njm == 0) {
      run_to(std::false_type{}, total);
    } else {
      // joint rows between the contact sweeps: they change velocities in the shared-memory rows, so the velocities the
      // pipeline already holds for its current constraint are re-read afterwards
      joint_init();
      for (int it = 0; it < sp.velocity_iterations; ++it) {
        joint_solve();
        vaa = vl[ca.ba * 32];
        vab = vl[ca.bb * 32];
        run_to(std::false_type{}, n_warm + (it + 1) * ncm);
      }
    }
    cp_async_wait<0>();
  }
  __syncwarp();
  if (live)
    for (int b = 0; b < B.NB; ++b) B.b_vel[x.at(B.NB, b)] = vel[b * 32 + lane];
  timeline_record(B, 1, t_start);
}

// ------------------------------------------------------------------------------------------
// position iterations with per-island early exit.  Same CTA shape and the same software pipeline;
// bodies carry (c.x, c.y, a) and the cached rotation (sin a, cos a), see solve_position_one.
// ------------------------------------------------------------------------------------------
struct PcRegs {
  float4 p0, p1, p2, p3;
  float ra, rb;
  int ba, bb, cnt, type, isl;
};
__device__ __forceinline__ PcRegs pc_load(const float4* st) {
  PcRegs r;
  r.p0 = st[0 * 32]; r.p1 = st[1 * 32]; r.p2 = st[2 * 32]; r.p3 = st[3 * 32];
  const float4 p4 = st[4 * 32], p5 = st[5 * 32];
  r.ra = p4.x; r.rb = p4.y;
  r.ba = __float_as_int(p4.z); r.bb = __float_as_int(p4.w);
  const int packed = __float_as_int(p5.x);
  r.cnt = packed & 0xff; r.type = (packed >> 8) & 0xff;
  r.isl = __float_as_int(p5.y);
  return r;
}

// ------------------------------------------------------------------------------------------
// Straight-line position stage (the default for batches): one lane per world like velocity_sl_kernel, one
// visit = one basic block.  Face manifolds with two points (the vote taken one visit ahead says whether all
// active lanes have one) go through solve_position_face2: selects instead of the face-type branches, and an
// unconditional branch-free rotation refresh.  Island bookkeeping is arithmetic: the running minimum
// separation is closed into a per-island byte table in shared memory at the island's last constraint
// (address-selected store: a visit that closes nothing writes a scratch entry), solved islands and finished
// worlds run on the scratch body row.  A sweep starts with a fresh prologue, so the early exit of
// b2_island_private.rs:257-274 is evaluated once per sweep, outside the visits.
// ------------------------------------------------------------------------------------------
static_assert(PC_Q == 6, "cp_async_prec copies six rows");
__device__ __forceinline__ void cp_async_prec(float4* smem_dst, const float4* gmem_src) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile(
      "cp.async.cg.shared.global [%0], [%1], 16;\n"
      "cp.async.cg.shared.global [%0+512], [%1+512], 16;\n"
      "cp.async.cg.shared.global [%0+1024], [%1+1024], 16;\n"
      "cp.async.cg.shared.global [%0+1536], [%1+1536], 16;\n"
      "cp.async.cg.shared.global [%0+2048], [%1+2048], 16;\n"
      "cp.async.cg.shared.global [%0+2560], [%1+2560], 16;\n"
      "cp.async.commit_group;\n" ::"r"(s),
      "l"(gmem_src)
      : "memory");
}

inline size_t position_sl_smem_bytes(int NB) {
  return (size_t)POS_RING * PC_Q * 32 * 16 + (size_t)(NB + 1) * 32 * (16 + 8) + 2 * (size_t)(NB + 1) * 32;
}

__global__ void __launch_bounds__(32) position_sl_kernel(const Batch B, const StepParams sp) {
  extern __shared__ float4 smem4[];
  const unsigned long long t_start = B.timeline ? global_ns() : 0ull;
  float4* ring = smem4;                                          // [POS_RING][PC_Q][32]
  float4* pos = smem4 + POS_RING * PC_Q * 32;                    // [NB + 1][32]: c.x c.y a -; row NB is scratch
  float2* rot = (float2*)(pos + (size_t)(B.NB + 1) * 32);        // [NB + 1][32]: sin a, cos a
  unsigned char* tab = (unsigned char*)(rot + (size_t)(B.NB + 1) * 32);  // [NB + 1][32]: island solved; row NB is scratch
  unsigned char* tab_prev = tab + (size_t)(B.NB + 1) * 32;               // [NB + 1][32]: the table before the current sweep (joints)
  const int lane = threadIdx.x;
  const int wb = blockIdx.x + B.wb_first;
  const int w = wb * 32 + lane;
  const bool live = w < B.n_worlds;
  WIdx x;
  x.wb = wb; x.wl = lane; x.LB = 32;
  Ws ws = ws_of(B, x);
  const int nc = live ? ws[WS_ISL_CONTACTS] : 0;
  const int nisl = live ? ws[WS_ISL_COUNT] : 0;
  const int ncm = __reduce_max_sync(0xffffffffu, nc);
  const int nj = (live && B.NJ > 0) ? ws[WS_ISL_JOINTS] : 0;
  const int njm = B.NJ > 0 ? __reduce_max_sync(0xffffffffu, nj) : 0;
  if ((ncm == 0 && njm == 0) || sp.position_iterations <= 0) return;
  if (live) {
    for (int b = 0; b < B.NB; ++b) {
      const float4 p = B.b_pos[x.at(B.NB, b)];
      const float4 r = B.b_rot[x.at(B.NB, b)];
      pos[b * 32 + lane] = p;
      rot[b * 32 + lane] = make_float2(r.x, r.y);
    }
    for (int i = 0; i < nisl; ++i) tab[i * 32 + lane] = (unsigned char)(B.isl_flags[x.at(B.NB, i)] & 1);
  }
  pos[B.NB * 32 + lane] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  rot[B.NB * 32 + lane] = make_float2(0.0f, 1.0f);
  tab[B.NB * 32 + lane] = 0;
  const float4* src = B.pc + (size_t)wb * B.NC * PC_Q * 32 + lane;
  float4* pl = pos + lane;
  float2* ql = rot + lane;
  unsigned char* tl = tab + lane;
  float4* rl = ring + lane;
  const int scratch = B.NB;
  int fk = 0;
  const float4* fsrc = src;
  auto fetch = [&](int p) {  // next record of the (wrapping) stream -> stage p % RING
    cp_async_prec(rl + ((p & (POS_RING - 1)) * PC_Q) * 32, fsrc);
    const bool wrap = (fk + 1 == ncm);
    fk = wrap ? 0 : fk + 1;
    fsrc = wrap ? src : fsrc + PC_Q * 32;
  };
  if (ncm > 0) {
#pragma unroll
    for (int p = 0; p < POS_RING; ++p) fetch(p);
  }
  PcRegs ca, cb;
  bool acta = false, actb = false, fa = true, fb = true;
  float4 paa, pab, pba, pbb;
  float2 qaa, qab, qba, qbb;
  int k = 0, p = 0;
  bool done = !live || (nc == 0 && nj == 0), all_solved = true;
  const BodyStatePosSmem jst = {pl, ql};
  float min_separation = 0.0f;
  // activity of the constraint at stream index kk whose record is r: its island is still open in this sweep
  auto prepare = [&](PcRegs& r, int kk, bool& act, bool& fast, float4& pa, float4& pb, float2& qa, float2& qb) {
    const bool in = kk < nc;
    const int isl = in ? r.isl : scratch;
    act = in && !done && tl[isl * 32] == 0;
    fast = __all_sync(0xffffffffu, !act || (r.type != B2GPU_MANIFOLD_CIRCLES && r.cnt == 2));
    r.ba = act ? r.ba : scratch;
    r.bb = act ? r.bb : scratch;
    r.cnt = act ? r.cnt : 0;
    r.isl = isl;
    pa = pl[r.ba * 32]; pb = pl[r.bb * 32];
    qa = ql[r.ba * 32]; qb = ql[r.bb * 32];
  };
  auto half = [&](auto FAST, PcRegs& cur, const bool act, float4& pa, float4& pb, float2& qa, float2& qb, PcRegs& nxt,
                  bool& nact, bool& nfast, float4& npa, float4& npb, float2& nqa, float2& nqb) {
    // -- position p+1 into the other register set (at the end of a sweep this is discarded: the next sweep
    //    starts from its own prologue, after the island table and `done` are final)
    const int kc = k;
    k = (k + 1 == ncm) ? 0 : k + 1;
    cp_async_wait<POS_RING - 2>();
    nxt = pc_load(rl + (((p + 1) & (POS_RING - 1)) * PC_Q) * 32);
    prepare(nxt, k, nact, nfast, npa, npb, nqa, nqb);
    fetch(p);  // this position's stage is free (cur is in registers): refill it with position p + RING
    // -- position p: the reference's arithmetic
    PosState s;
    s.c_a = v2(pa.x, pa.y); s.a_a = pa.z; s.q_a.s = qa.x; s.q_a.c = qa.y;
    s.c_b = v2(pb.x, pb.y); s.a_b = pb.z; s.q_b.s = qb.x; s.q_b.c = qb.y;
    float ms;
    if (decltype(FAST)::value) {
      const PosState s0 = s;
      bool wide;
      ms = solve_position_face2(s, cur.p0, cur.p1, cur.p2, cur.p3, cur.type == B2GPU_MANIFOLD_FACE_A, cur.ra, cur.rb, min_separation, wide);
      if (__any_sync(0xffffffffu, wide)) {  // never in practice: an angle beyond +-120 rad
        s = s0;
        ms = solve_position_one(s, cur.p0, cur.p1, cur.p2, cur.p3, cur.type, cur.cnt, cur.ra, cur.rb, min_separation);
      }
    } else {
      ms = solve_position_one(s, cur.p0, cur.p1, cur.p2, cur.p3, cur.type, cur.cnt, cur.ra, cur.rb, min_separation);
    }
    pa = make_float4(s.c_a.x, s.c_a.y, s.a_a, pa.w);
    pb = make_float4(s.c_b.x, s.c_b.y, s.a_b, pb.w);
    qa = make_float2(s.q_a.s, s.q_a.c);
    qb = make_float2(s.q_b.s, s.q_b.c);
    pl[cur.ba * 32] = pa; ql[cur.ba * 32] = qa;
    pl[cur.bb * 32] = pb; ql[cur.bb * 32] = qb;
    // -- forward the fresh state to the next constraint where it shares a body with this one
    const int maa = sel_mask(nxt.ba == cur.ba), mab = sel_mask(nxt.ba == cur.bb);
    const int mba = sel_mask(nxt.bb == cur.ba), mbb = sel_mask(nxt.bb == cur.bb);
    npa.x = msel(maa, pa.x, msel(mab, pb.x, npa.x));
    npa.y = msel(maa, pa.y, msel(mab, pb.y, npa.y));
    npa.z = msel(maa, pa.z, msel(mab, pb.z, npa.z));
    nqa.x = msel(maa, qa.x, msel(mab, qb.x, nqa.x));
    nqa.y = msel(maa, qa.y, msel(mab, qb.y, nqa.y));
    npb.x = msel(mba, pa.x, msel(mbb, pb.x, npb.x));
    npb.y = msel(mba, pa.y, msel(mbb, pb.y, npb.y));
    npb.z = msel(mba, pa.z, msel(mbb, pb.z, npb.z));
    nqb.x = msel(mba, qa.x, msel(mbb, qb.x, nqb.x));
    nqb.y = msel(mba, qa.y, msel(mbb, qb.y, nqb.y));
    // -- island bookkeeping: close the island at its last constraint
    const bool last = act && ((kc + 1 == nc) || (nxt.isl != cur.isl));
    min_separation = act ? ms : min_separation;
    const bool solved = min_separation >= -3.0f * B2G_LINEAR_SLOP;
    tl[(last ? cur.isl : scratch) * 32] = (unsigned char)(solved ? 1 : 0);
    all_solved = all_solved && (!last || solved);
    min_separation = last ? 0.0f : min_separation;
    ++p;
  };
  auto step_ab = [&]() {
    if (fa) half(std::true_type{}, ca, acta, paa, pab, qaa, qab, cb, actb, fb, pba, pbb, qba, qbb);
    else half(std::false_type{}, ca, acta, paa, pab, qaa, qab, cb, actb, fb, pba, pbb, qba, qbb);
  };
  auto step_ba = [&]() {
    if (fb) half(std::true_type{}, cb, actb, pba, pbb, qba, qbb, ca, acta, fa, paa, pab, qaa, qab);
    else half(std::false_type{}, cb, actb, pba, pbb, qba, qbb, ca, acta, fa, paa, pab, qaa, qab);
  };
  for (int sweep = 0; sweep < sp.position_iterations; ++sweep) {
    if (njm > 0 && !done)
      for (int i = 0; i < nisl; ++i) tab_prev[i * 32 + lane] = tl[i * 32];
    if (ncm > 0) {
      // prologue of the sweep: position p (stream index 0) from its ring stage
      k = 0;
      cp_async_wait<POS_RING - 1>();
      ca = pc_load(rl + ((p & (POS_RING - 1)) * PC_Q) * 32);
      prepare(ca, 0, acta, fa, paa, pab, qaa, qab);
      int i = 0;
      for (; i + 2 <= ncm; i += 2) {
        step_ab();
        step_ba();
      }
      if (i < ncm) step_ab();
    }
    if (njm > 0 && !done) {
      // joint rows after the contact rows (b2_island_private.rs:262-266), island by island: an island that was open at the
      // start of this sweep solves all its joints; it is solved when its contacts passed (closed to 1 by this sweep, or it
      // has none) AND every joint is within tolerance
      all_solved = true;
      for (int i = 0; i < nisl; ++i) {
        if (tab_prev[i * 32 + lane] == 0) {
          const int4 rg = B.isl_range[x.at(B.NB, i)];
          const int2 jr = B.isl_jrange[x.at(B.NB, i)];
          bool ok = tl[i * 32] != 0 || rg.z == rg.w;
          for (int q = jr.x; q < jr.y; ++q) {
            const bool joint_okay = joint_solve_position(B, x, jst, B.isl_joint[x.at(B.NJ, q)]);
            ok = ok && joint_okay;
          }
          tl[i * 32] = (unsigned char)(ok ? 1 : 0);
        }
        all_solved = all_solved && tl[i * 32] != 0;
      }
    }
    // end of the sweep: a world is finished when every island passed the exit test
    if (!done && all_solved) done = true;
    all_solved = true;
    min_separation = 0.0f;
    if (__all_sync(0xffffffffu, done)) break;
  }
  cp_async_wait<0>();
  __syncwarp();
  if (live && (nc > 0 || nj > 0)) {
    // the rows carry the fourth component of b_pos (sleep time) through unchanged; of b_rot only the running
    // rotation (first two components) is written, so neither store needs a read
    for (int b = 0; b < B.NB; ++b) {
      const int bi = x.at(B.NB, b);
      B.b_pos[bi] = pos[b * 32 + lane];
      *reinterpret_cast<float2*>(&B.b_rot[bi]) = rot[b * 32 + lane];
    }
    for (int i = 0; i < nisl; ++i)
      if (tab[i * 32 + lane]) B.isl_flags[x.at(B.NB, i)] = 1;  // bit 0 is the only bit of the word
  }
  timeline_record(B, 2, t_start);
}
}  // namespace b2g
