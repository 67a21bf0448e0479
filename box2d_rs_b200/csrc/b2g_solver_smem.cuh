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

namespace b2g {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

constexpr int VEL_RING = 8;  // stages of the velocity constraint ring
constexpr int POS_RING = 8;

inline size_t velocity_smem_bytes(int NB) { return (size_t)NB * 3 * 32 * 4 + (size_t)VEL_RING * VC_Q * 32 * 16; }
inline size_t position_smem_bytes(int NB) { return (size_t)NB * 5 * 32 * 4 + (size_t)POS_RING * PC_Q * 32 * 16; }

// ------------------------------------------------------------------------------------------
// warm start + velocity iterations.  grid = world blocks, block = 32 lanes (one world each).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) velocity_smem_kernel(const Batch B, const StepParams sp) {
  extern __shared__ float4 smem4[];
  float4* ring = smem4;                                        // [VEL_RING][VC_Q][32]
  float* vel = (float*)(smem4 + VEL_RING * VC_Q * 32);         // [NB][3][32]
  const int lane = threadIdx.x;
  const int wb = blockIdx.x;
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
  if (ncm == 0) return;
  // stage body velocities (coalesced float4 reads, conflict-free scalar smem writes)
  if (live) {
    for (int b = 0; b < B.NB; ++b) {
      const float4 v = B.b_vel[x.at(B.NB, b)];
      vel[(b * 3 + 0) * 32 + lane] = v.x;
      vel[(b * 3 + 1) * 32 + lane] = v.y;
      vel[(b * 3 + 2) * 32 + lane] = v.z;
    }
  }
  const float4* src = B.vc + (size_t)wb * B.NC * VC_Q * 32 + lane;  // + (k * VC_Q + q) * 32
  const int sweeps = 1 + sp.velocity_iterations;                   // sweep 0 = warm start
  const int total = sweeps * ncm;
  const bool resident = ncm <= VEL_RING;  // the whole stream fits: load once, iterate in shared memory
  auto fetch = [&](int pos) {              // stage constraint (pos % ncm) of sweep (pos / ncm)
    if (pos < total) {
      const int k = pos % ncm;
      if (k < nc) {
        float4* dst = ring + (size_t)((resident ? k : pos % VEL_RING) * VC_Q) * 32 + lane;
        const float4* s = src + (size_t)k * VC_Q * 32;
#pragma unroll
        for (int q = 0; q < VC_Q; ++q) cp_async16(dst + q * 32, s + q * 32);
      }
    }
    cp_async_commit();
  };
  const int prologue = resident ? ncm : VEL_RING - 1;
  for (int p = 0; p < VEL_RING - 1; ++p) {
    if (resident) { if (p < ncm) fetch(p); else cp_async_commit(); }
    else fetch(p);
  }
  if (resident && ncm == VEL_RING) fetch(VEL_RING - 1);
  (void)prologue;
  int k = 0, sweep = 0;
  for (int pos = 0; pos < total; ++pos) {
    if (resident) cp_async_wait<0>(); else cp_async_wait<VEL_RING - 2>();
    if (!resident) fetch(pos + VEL_RING - 1);
    if (k < nc && (sweep > 0 || warm)) {
      float4* st = ring + (size_t)((resident ? k : pos % VEL_RING) * VC_Q) * 32 + lane;
      const float4 q8 = st[8 * 32];
      const int ba = __float_as_int(q8.x), bb = __float_as_int(q8.y), vc_points = __float_as_int(q8.z) & 0xff;
      if (vc_points > 0) {
        VelState s;
        s.v_a = v2(vel[(ba * 3 + 0) * 32 + lane], vel[(ba * 3 + 1) * 32 + lane]);
        s.w_a = vel[(ba * 3 + 2) * 32 + lane];
        s.v_b = v2(vel[(bb * 3 + 0) * 32 + lane], vel[(bb * 3 + 1) * 32 + lane]);
        s.w_b = vel[(bb * 3 + 2) * 32 + lane];
        const float4 q0 = st[0 * 32], q1 = st[1 * 32], q2 = st[2 * 32], q7 = st[7 * 32];
        float4 q6 = st[6 * 32];
        if (sweep == 0) {
          warm_start_one(s, q0, q1, q2, q6, q7, vc_points);
        } else {
          const float4 q3 = st[3 * 32], q4 = st[4 * 32], q5 = st[5 * 32];
          solve_velocity_one(s, q0, q1, q2, q3, q4, q5, q6, q7, vc_points, block);
          if (resident) st[6 * 32] = q6;
          if (!resident || sweep == sweeps - 1) B.vc[vc_at(B, x, k, 6)] = q6;
        }
        vel[(ba * 3 + 0) * 32 + lane] = s.v_a.x;
        vel[(ba * 3 + 1) * 32 + lane] = s.v_a.y;
        vel[(ba * 3 + 2) * 32 + lane] = s.w_a;
        vel[(bb * 3 + 0) * 32 + lane] = s.v_b.x;
        vel[(bb * 3 + 1) * 32 + lane] = s.v_b.y;
        vel[(bb * 3 + 2) * 32 + lane] = s.w_b;
      }
    }
    if (++k == ncm) { k = 0; ++sweep; }
  }
  cp_async_wait<0>();
  if (live) {
    for (int b = 0; b < B.NB; ++b)
      B.b_vel[x.at(B.NB, b)] = make_float4(vel[(b * 3 + 0) * 32 + lane], vel[(b * 3 + 1) * 32 + lane],
                                           vel[(b * 3 + 2) * 32 + lane], 0.0f);
  }
}

// ------------------------------------------------------------------------------------------
// position iterations with per-island early exit.  Same CTA shape; bodies carry (c.x, c.y, a, sin a, cos a).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) position_smem_kernel(const Batch B, const StepParams sp) {
  extern __shared__ float4 smem4[];
  float4* ring = smem4;                                        // [POS_RING][PC_Q][32]
  float* pos = (float*)(smem4 + POS_RING * PC_Q * 32);         // [NB][5][32]
  const int lane = threadIdx.x;
  const int wb = blockIdx.x;
  const int w = wb * 32 + lane;
  const bool live = w < B.n_worlds;
  WIdx x;
  x.wb = wb; x.wl = lane; x.LB = 32;
  Ws ws = ws_of(B, x);
  const int nc = live ? ws[WS_ISL_CONTACTS] : 0;
  const int ncm = __reduce_max_sync(0xffffffffu, nc);
  if (ncm == 0 || sp.position_iterations <= 0) return;
  if (live) {
    for (int b = 0; b < B.NB; ++b) {
      const float4 p = B.b_pos[x.at(B.NB, b)];
      const float4 r = B.b_rot[x.at(B.NB, b)];
      pos[(b * 5 + 0) * 32 + lane] = p.x;
      pos[(b * 5 + 1) * 32 + lane] = p.y;
      pos[(b * 5 + 2) * 32 + lane] = p.z;
      pos[(b * 5 + 3) * 32 + lane] = r.x;
      pos[(b * 5 + 4) * 32 + lane] = r.y;
    }
  }
  const float4* src = B.pc + (size_t)wb * B.NC * PC_Q * 32 + lane;
  const int total = sp.position_iterations * ncm;
  const bool resident = ncm <= POS_RING;
  auto fetch = [&](int p) {
    if (p < total) {
      const int k = p % ncm;
      if (k < nc) {
        float4* dst = ring + (size_t)((resident ? k : p % POS_RING) * PC_Q) * 32 + lane;
        const float4* s = src + (size_t)k * PC_Q * 32;
#pragma unroll
        for (int q = 0; q < PC_Q; ++q) cp_async16(dst + q * 32, s + q * 32);
      }
    }
    cp_async_commit();
  };
  for (int p = 0; p < POS_RING - 1; ++p) {
    if (resident) { if (p < ncm) fetch(p); else cp_async_commit(); }
    else fetch(p);
  }
  if (resident && ncm == POS_RING) fetch(POS_RING - 1);
  int k = 0;
  int cur = -1;
  bool skip = false, all_solved = true, done = !live || nc == 0;
  float min_separation = 0.0f;
  for (int p = 0; p < total; ++p) {
    if (resident) cp_async_wait<0>(); else cp_async_wait<POS_RING - 2>();
    if (!resident) fetch(p + POS_RING - 1);
    if (k < nc && !done) {
      const float4* st = ring + (size_t)((resident ? k : p % POS_RING) * PC_Q) * 32 + lane;
      const float4 p5 = st[5 * 32];
      const int isl = __float_as_int(p5.y);
      if (isl != cur) {
        if (cur >= 0 && !skip) {
          if (min_separation >= -3.0f * B2G_LINEAR_SLOP) B.isl_flags[x.at(B.NB, cur)] |= 1; else all_solved = false;
        }
        cur = isl;
        skip = (B.isl_flags[x.at(B.NB, cur)] & 1) != 0;
        min_separation = 0.0f;
      }
      if (!skip) {
        const float4 p4 = st[4 * 32];
        const int ba = __float_as_int(p4.z), bb = __float_as_int(p4.w), packed = __float_as_int(p5.x);
        PosState s;
        s.c_a = v2(pos[(ba * 5 + 0) * 32 + lane], pos[(ba * 5 + 1) * 32 + lane]);
        s.a_a = pos[(ba * 5 + 2) * 32 + lane];
        s.q_a.s = pos[(ba * 5 + 3) * 32 + lane];
        s.q_a.c = pos[(ba * 5 + 4) * 32 + lane];
        s.c_b = v2(pos[(bb * 5 + 0) * 32 + lane], pos[(bb * 5 + 1) * 32 + lane]);
        s.a_b = pos[(bb * 5 + 2) * 32 + lane];
        s.q_b.s = pos[(bb * 5 + 3) * 32 + lane];
        s.q_b.c = pos[(bb * 5 + 4) * 32 + lane];
        min_separation = solve_position_one(s, st[0 * 32], st[1 * 32], st[2 * 32], st[3 * 32], (packed >> 8) & 0xff,
                                            packed & 0xff, p4.x, p4.y, min_separation);
        pos[(ba * 5 + 0) * 32 + lane] = s.c_a.x;
        pos[(ba * 5 + 1) * 32 + lane] = s.c_a.y;
        pos[(ba * 5 + 2) * 32 + lane] = s.a_a;
        pos[(ba * 5 + 3) * 32 + lane] = s.q_a.s;
        pos[(ba * 5 + 4) * 32 + lane] = s.q_a.c;
        pos[(bb * 5 + 0) * 32 + lane] = s.c_b.x;
        pos[(bb * 5 + 1) * 32 + lane] = s.c_b.y;
        pos[(bb * 5 + 2) * 32 + lane] = s.a_b;
        pos[(bb * 5 + 3) * 32 + lane] = s.q_b.s;
        pos[(bb * 5 + 4) * 32 + lane] = s.q_b.c;
      }
    }
    if (++k == ncm) {  // end of a sweep: close the last island, test the early exit of this world
      k = 0;
      if (!done) {
        if (cur >= 0 && !skip) {
          if (min_separation >= -3.0f * B2G_LINEAR_SLOP) B.isl_flags[x.at(B.NB, cur)] |= 1; else all_solved = false;
        }
        if (all_solved) done = true;
        cur = -1;
        skip = false;
        all_solved = true;
        min_separation = 0.0f;
      }
      if (__all_sync(0xffffffffu, done)) break;
    }
  }
  cp_async_wait<0>();
  if (live && nc > 0) {
    for (int b = 0; b < B.NB; ++b) {
      const int bi = x.at(B.NB, b);
      float4 p = B.b_pos[bi];
      float4 r = B.b_rot[bi];
      p.x = pos[(b * 5 + 0) * 32 + lane];
      p.y = pos[(b * 5 + 1) * 32 + lane];
      p.z = pos[(b * 5 + 2) * 32 + lane];
      r.x = pos[(b * 5 + 3) * 32 + lane];
      r.y = pos[(b * 5 + 4) * 32 + lane];
      B.b_pos[bi] = p;
      B.b_rot[bi] = r;
    }
  }
}

}  // namespace b2g
