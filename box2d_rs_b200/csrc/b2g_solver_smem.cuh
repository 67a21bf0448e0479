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

// TMA 1-D bulk copy (cp.async.bulk, SASS UBLKCP) completing on an mbarrier: one elected lane moves a whole
// constraint record of the 32-world block (VC_Q rows x 512 B, contiguous in HBM) into a ring stage.
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   (unsigned)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned done = 0;
  const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"(addr), "r"(parity)
                 : "memory");
  }
}


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
__device__ __forceinline__ void velocity_resident(float4* rl, float4* vl, const float4* src, float4* q6_out, int nc, int ncm,
                                                  bool warm, bool block, int sweeps) {
  for (int k = 0; k < ncm; ++k) {
#pragma unroll
    for (int q = 0; q < VC_Q; ++q) cp_async16(rl + (k * VC_Q + q) * 32, src + (size_t)(k * VC_Q + q) * 32);
  }
  cp_async_commit();
  cp_async_wait<0>();
  for (int sweep = 0; sweep < sweeps; ++sweep) {
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
}

// ------------------------------------------------------------------------------------------
// warm start + velocity iterations.  grid = world blocks, block = 32 lanes (one world each).
//
// The loop is software-pipelined by hand: while constraint p is being solved (a ~60-deep chain of
// dependent fp32 operations), the rows of constraint p+1 are read from the ring into registers and
// the velocities of its two bodies are read from shared memory.  Those velocity reads can be stale
// for a body that constraint p is about to update — which, in the reference's DFS contact order, is
// the usual case — so after the solve the fresh values are forwarded from registers (two compares
// and selects instead of a store -> load round trip through shared memory on the dependent chain).
// ------------------------------------------------------------------------------------------
template <bool USE_TMA>
__global__ void __launch_bounds__(32) velocity_smem_kernel(const Batch B, const StepParams sp) {
  extern __shared__ float4 smem4[];
  float4* ring = smem4;                      // [VEL_RING][VC_Q][32]
  float4* vel = smem4 + VEL_RING * VC_Q * 32;  // [NB][32]: v.x v.y w -
  uint64_t* bars = (uint64_t*)(vel + (size_t)(B.NB + 1) * 32);  // [VEL_RING] one mbarrier per ring stage (TMA form)
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
  if (ncm == 0) return;
  if (live)
    for (int b = 0; b < B.NB; ++b) vel[b * 32 + lane] = B.b_vel[x.at(B.NB, b)];
  const float4* src = B.vc + (size_t)wb * B.NC * VC_Q * 32 + lane;  // + (k * VC_Q + q) * 32
  float4* q6_out = B.vc + (size_t)wb * B.NC * VC_Q * 32 + 6 * 32 + lane;
  const int sweeps = 1 + sp.velocity_iterations;  // sweep 0 = warm start
  const int total = sweeps * ncm;
  float4* vl = vel + lane;
  float4* rl = ring + lane;

  if (ncm <= VEL_RING) {
    velocity_resident(rl, vl, src, q6_out, nc, ncm, warm, block, sweeps);
  } else {
    // ---- streaming form: ring of VEL_RING stages, constraint p lives in stage p % VEL_RING
    int fk = 0;                 // next constraint index to fetch (wraps at ncm)
    int fpos = 0;               // its flattened position
    const float4* fsrc = src;
    const float4* fsrc_block = B.vc + (size_t)wb * B.NC * VC_Q * 32;  // the block's records (all 32 lanes)
    if (USE_TMA) {
      if (lane == 0)
        for (int st = 0; st < VEL_RING; ++st) mbar_init(&bars[st], 1);
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncwarp();
    }
    auto fetch = [&]() {
      if (fpos < total) {
        if (USE_TMA) {
          __syncwarp();  // every lane has copied this stage's previous record into registers
          if (lane == 0) {
            const int st = fpos & (VEL_RING - 1);
            // the impulses of this record were last written by ordinary stores of all 32 lanes: order them
            // before the async-proxy read
            asm volatile("fence.proxy.async.global;\n" ::: "memory");
            mbar_expect_tx(&bars[st], VC_Q * 32 * 16);
            bulk_copy_g2s(ring + (size_t)st * VC_Q * 32, fsrc_block + (size_t)fk * VC_Q * 32, VC_Q * 32 * 16, &bars[st]);
          }
        } else {
          float4* dst = rl + ((fpos & (VEL_RING - 1)) * VC_Q) * 32;
#pragma unroll
          for (int q = 0; q < VC_Q; ++q) cp_async16(dst + q * 32, fsrc + q * 32);
          fsrc += VC_Q * 32;
        }
        if (++fk == ncm) { fk = 0; fsrc = src; }
      }
      ++fpos;
      if (!USE_TMA) cp_async_commit();
    };
    auto landed = [&](int at_pos, int pending_ok) {  // the record of position at_pos is in its stage
      if (USE_TMA) {
        if (at_pos < total) mbar_wait(&bars[at_pos & (VEL_RING - 1)], (unsigned)(at_pos / VEL_RING) & 1u);
      } else if (pending_ok == VEL_RING - 1) {
        cp_async_wait<VEL_RING - 1>();
      } else {
        cp_async_wait<VEL_RING - 2>();
      }
    };
#pragma unroll
    for (int p = 0; p < VEL_RING; ++p) fetch();  // positions 0 .. RING-1 in flight
    landed(0, VEL_RING - 1);
    // Two register sets (A, B) alternate as "current" and "next": the loop body handles two positions
    // so that the hand-over between them is pure register renaming.
    VcRegs ca = vc_load(rl), cb;
    int k = 0, sweep = 0, pos = 0;
    bool acta = (k < nc) && warm && ca.cnt > 0, actb = false;
    if (!acta) { ca.ba = 0; ca.bb = 0; }
    float4 vaa = vl[ca.ba * 32], vab = vl[ca.bb * 32], vba, vbb;
    auto half_step = [&](VcRegs& cur, bool act, float4& va, float4& vb, VcRegs& nxt, bool& nact, float4& nva, float4& nvb) {
      // -- prefetch position pos+1 into registers (its stage landed: at most RING-2 younger groups pending)
      const int kc = k, sc = sweep;
      if (++k == ncm) { k = 0; ++sweep; }
      landed(pos + 1, VEL_RING - 2);
      nxt = vc_load(rl + (((pos + 1) & (VEL_RING - 1)) * VC_Q) * 32);
      nact = (pos + 1 < total) && (k < nc) && (sweep > 0 || warm) && nxt.cnt > 0;
      if (!nact) { nxt.ba = 0; nxt.bb = 0; }
      nva = vl[nxt.ba * 32];
      nvb = vl[nxt.bb * 32];
      fetch();  // the stage of position pos is free now (cur is in registers): refill it with position pos + RING
      if (act) {
        VelState s;
        s.v_a = v2(va.x, va.y); s.w_a = va.z;
        s.v_b = v2(vb.x, vb.y); s.w_b = vb.z;
        if (sc == 0) {
          warm_start_one(s, cur.q0, cur.q1, cur.q2, cur.q6, cur.q7, cur.cnt);
        } else {
          if (__all_sync(__activemask(), cur.cnt == 2 && block))
            solve_velocity_one(s, cur.q0, cur.q1, cur.q2, cur.q3, cur.q4, cur.q5, cur.q6, cur.q7, 2, true);
          else
            solve_velocity_one(s, cur.q0, cur.q1, cur.q2, cur.q3, cur.q4, cur.q5, cur.q6, cur.q7, cur.cnt, block);
          q6_out[(size_t)kc * VC_Q * 32] = cur.q6;
        }
        va = make_float4(s.v_a.x, s.v_a.y, s.w_a, 0.0f);
        vb = make_float4(s.v_b.x, s.v_b.y, s.w_b, 0.0f);
        vl[cur.ba * 32] = va;
        vl[cur.bb * 32] = vb;
        // -- forward the fresh velocities to the next constraint where it shares a body with this one
        if (nxt.ba == cur.ba) nva = va; else if (nxt.ba == cur.bb) nva = vb;
        if (nxt.bb == cur.ba) nvb = va; else if (nxt.bb == cur.bb) nvb = vb;
      }
      ++pos;
    };
    while (pos < total) {
      half_step(ca, acta, vaa, vab, cb, actb, vba, vbb);
      half_step(cb, actb, vba, vbb, ca, acta, vaa, vab);
    }
    cp_async_wait<0>();
  }
  __syncwarp();
  if (live)
    for (int b = 0; b < B.NB; ++b) B.b_vel[x.at(B.NB, b)] = vel[b * 32 + lane];
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
  if (ncm == 0) return;
  if (live)
    for (int b = 0; b < B.NB; ++b) vel[b * 32 + lane] = B.b_vel[x.at(B.NB, b)];
  const float4* src = B.vc + (size_t)wb * B.NC * VC_Q * 32 + lane;  // + (k * VC_Q + q) * 32
  float4* q6_out = B.vc + (size_t)wb * B.NC * VC_Q * 32 + 6 * 32 + lane;
  float4* vl = vel + lane;
  float4* rl = ring + lane;
  if (ncm <= VEL_RING) {
    velocity_resident(rl, vl, src, q6_out, nc, ncm, warm, block, 1 + sp.velocity_iterations);
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
    run_to(std::false_type{}, total);
    cp_async_wait<0>();
  }
  __syncwarp();
  if (live)
    for (int b = 0; b < B.NB; ++b) B.b_vel[x.at(B.NB, b)] = vel[b * 32 + lane];
  timeline_record(B, 1, t_start);
}

// ------------------------------------------------------------------------------------------
// Warp-specialised form of the straight-line kernel (experiment, solver='producer'; measured 8 % slower than
// velocity_sl_kernel: what the solving warp saves in copy instructions it pays back in barrier tests and in
// the shared-memory flags that replace the warp vote): a second warp of the CTA does nothing
// but keep the ring full, so the solving warp issues no copy instruction at all.  Hand-over per ring stage
// through two mbarriers: `full` (the 32 producer lanes' cp.async copies of a record have landed:
// cp.async.mbarrier.arrive.noinc) and `empty` (the 32 solving lanes have the record in registers).  The
// solving warp tests `full` two visits ahead with a non-blocking test_wait whose predicate is consumed at
// the end of the visit, so the barrier's latency is off the chain; the producer sleeps in try_wait.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_test(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok)
               : "r"(bar), "r"(parity)
               : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_spin(unsigned bar, unsigned parity) {
  unsigned done = 0;
  while (!done) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
  }
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async_record_signal(float4* smem_dst, const float4* gmem_src, unsigned bar) {
  const unsigned s = smem_u32(smem_dst);
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
      "cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%2];\n" ::"r"(s),
      "l"(gmem_src), "r"(bar)
      : "memory");
}

inline size_t velocity_ws_smem_bytes(int NB) { return velocity_smem_bytes(NB) + 2 * VEL_RING * 8 + 32; }

__global__ void __launch_bounds__(64) velocity_ws_kernel(const Batch B, const StepParams sp) {
  extern __shared__ float4 smem4[];
  const unsigned long long t_start = B.timeline ? global_ns() : 0ull;
  float4* ring = smem4;                        // [VEL_RING][VC_Q][32]
  float4* vel = smem4 + VEL_RING * VC_Q * 32;  // [NB + 1][32]: v.x v.y w -; row NB is scratch
  uint64_t* bars = (uint64_t*)(vel + (size_t)(B.NB + 1) * 32);  // full[RING], empty[RING]
  const int lane = threadIdx.x & 31;
  const int role = threadIdx.x >> 5;  // 0: solver, 1: producer
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
  if (ncm == 0) return;
  if (live)  // both warps share the load of the velocity rows
    for (int b = role; b < B.NB; b += 2) vel[b * 32 + lane] = B.b_vel[x.at(B.NB, b)];
  const float4* src = B.vc + (size_t)wb * B.NC * VC_Q * 32 + lane;  // + (k * VC_Q + q) * 32
  float4* q6_out = B.vc + (size_t)wb * B.NC * VC_Q * 32 + 6 * 32 + lane;
  float4* vl = vel + lane;
  float4* rl = ring + lane;
  const bool streaming = ncm > VEL_RING;
  const int n_warm = __any_sync(0xffffffffu, warm) ? ncm : 0;  // positions of the warm-start sweep
  const int total = n_warm + sp.velocity_iterations * ncm;
  const unsigned full0 = smem_u32(bars), empty0 = smem_u32(bars + VEL_RING);
  // "some lane needs the general path at position p" flags, slot p % 4: written (benign same-value race) one
  // visit ahead by the solving warp, which runs in lockstep, and cleared two visits before reuse.  A warp
  // vote would do, but inside the role branch it costs a convergence check that splits the visit's block.
  volatile int* gflag = (volatile int*)(bars + 2 * VEL_RING);
  if (streaming && threadIdx.x == 0) {
    for (int st = 0; st < VEL_RING; ++st) { mbar_init(&bars[st], 32); mbar_init(&bars[VEL_RING + st], 32); }
    for (int i = 0; i < 4; ++i) gflag[i] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (!streaming) {
    if (role == 0) velocity_resident(rl, vl, src, q6_out, nc, ncm, warm, block, 1 + sp.velocity_iterations);
  } else if (role == 1) {
    // ---- producer: position p -> stage p % RING, as soon as the solver has released the stage's previous
    //      record; two positions past the end so that the solver's look-ahead test always completes
    int fk = 0;
    const float4* fsrc = src;
    for (int p = 0; p < total + 2; ++p) {
      const int st = p & (VEL_RING - 1);
      if (p >= VEL_RING) mbar_spin(empty0 + st * 8, (unsigned)((p / VEL_RING) - 1) & 1u);
      cp_async_record_signal(rl + (st * VC_Q) * 32, fsrc, full0 + st * 8);
      const bool wrap = (fk + 1 == ncm);
      fk = wrap ? 0 : fk + 1;
      fsrc = wrap ? src : fsrc + VC_Q * 32;
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
  } else {
    // ---- solver (see velocity_sl_kernel for the structure of a visit)
    const int scratch = B.NB;
    mbar_spin(full0, 0u);
    mbar_spin(full0 + 8, 0u);
    VcRegs ca = vc_load(rl), cb;
    int k = 0, pos = 0;
    const bool act0 = nc > 0 && ca.cnt > 0 && (warm || n_warm == 0);
    if (act0 && !(ca.cnt == 2 && block)) gflag[0] = 1;
    __syncwarp();
    bool fa = gflag[0] == 0, fb = false;
    if (!act0) { ca.ba = scratch; ca.bb = scratch; }
    float4 vaa = vl[ca.ba * 32], vab = vl[ca.bb * 32], vba = vaa, vbb = vab;
    cb = ca;
    auto half = [&](auto WARM, auto FAST, VcRegs& cur, float4& va, float4& vb, VcRegs& nxt, bool& nfast, float4& nva,
                    float4& nvb) {
      // -- look-ahead: has the record of position pos+2 landed?  (consumed at the end of this visit)
      const bool ahead = mbar_test(full0 + ((pos + 2) & (VEL_RING - 1)) * 8, (unsigned)((pos + 2) / VEL_RING) & 1u);
      // -- position pos+1 (known to have landed): rows and body velocities into the other register set
      const int kc = k;
      k = (k + 1 == ncm) ? 0 : k + 1;
      nxt = vc_load(rl + (((pos + 1) & (VEL_RING - 1)) * VC_Q) * 32);
      const bool nact = (k < nc) && nxt.cnt > 0 && (warm || pos + 1 >= n_warm);
      gflag[(nact && !(nxt.cnt == 2 && block)) ? ((pos + 1) & 3) : 4] = 1;  // slot 4: nobody reads it
      gflag[(pos + 3) & 3] = 0;
      nfast = gflag[(pos + 1) & 3] == 0;
      nxt.ba = nact ? nxt.ba : scratch;
      nxt.bb = nact ? nxt.bb : scratch;
      nva = vl[nxt.ba * 32];
      nvb = vl[nxt.bb * 32];
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
      va = make_float4(s.v_a.x, s.v_a.y, s.w_a, va.w);
      vb = make_float4(s.v_b.x, s.v_b.y, s.w_b, vb.w);
      vl[cur.ba * 32] = va;
      vl[cur.bb * 32] = vb;
      const int maa = sel_mask(nxt.ba == cur.ba), mab = sel_mask(nxt.ba == cur.bb);
      const int mba = sel_mask(nxt.bb == cur.ba), mbb = sel_mask(nxt.bb == cur.bb);
      nva.x = msel(maa, va.x, msel(mab, vb.x, nva.x));
      nva.y = msel(maa, va.y, msel(mab, vb.y, nva.y));
      nva.z = msel(maa, va.z, msel(mab, vb.z, nva.z));
      nvb.x = msel(mba, va.x, msel(mbb, vb.x, nvb.x));
      nvb.y = msel(mba, va.y, msel(mbb, vb.y, nvb.y));
      nvb.z = msel(mba, va.z, msel(mbb, vb.z, nvb.z));
      // -- this position's stage is free (its rows were consumed above; the impulses written here are ordered
      //    before the producer's re-read of the record by the release/acquire pair on the barrier)
      mbar_arrive(empty0 + (pos & (VEL_RING - 1)) * 8);
      if (!ahead) mbar_spin(full0 + ((pos + 2) & (VEL_RING - 1)) * 8, (unsigned)((pos + 2) / VEL_RING) & 1u);
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
    run_to(std::false_type{}, total);
  }
  __syncthreads();
  if (live)
    for (int b = role; b < B.NB; b += 2) B.b_vel[x.at(B.NB, b)] = vel[b * 32 + lane];
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

__global__ void __launch_bounds__(32) position_smem_kernel(const Batch B, const StepParams sp) {
  extern __shared__ float4 smem4[];
  float4* ring = smem4;                                   // [POS_RING][PC_Q][32]
  float4* pos = smem4 + POS_RING * PC_Q * 32;             // [NB][32]: c.x c.y a -
  float2* rot = (float2*)(pos + (size_t)B.NB * 32);       // [NB][32]: sin a, cos a
  const int lane = threadIdx.x;
  const int wb = blockIdx.x + B.wb_first;
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
      pos[b * 32 + lane] = p;
      rot[b * 32 + lane] = make_float2(r.x, r.y);
    }
  }
  const float4* src = B.pc + (size_t)wb * B.NC * PC_Q * 32 + lane;
  const int total = sp.position_iterations * ncm;
  float4* pl = pos + lane;
  float2* ql = rot + lane;
  float4* rl = ring + lane;
  const bool resident = ncm <= POS_RING;
  int fk = 0, fpos = 0;
  const float4* fsrc = src;
  auto fetch = [&]() {
    if (fpos < total && !(resident && fpos >= ncm)) {
      float4* dst = rl + ((resident ? fk : (fpos & (POS_RING - 1))) * PC_Q) * 32;
#pragma unroll
      for (int q = 0; q < PC_Q; ++q) cp_async16(dst + q * 32, fsrc + q * 32);
      fsrc += PC_Q * 32;
      if (++fk == ncm) { fk = 0; fsrc = src; }
    }
    ++fpos;
    cp_async_commit();
  };
#pragma unroll
  for (int p = 0; p < POS_RING; ++p) fetch();
  if (resident) cp_async_wait<0>(); else cp_async_wait<POS_RING - 1>();
  PcRegs ca = pc_load(rl), cb;
  int k = 0, p = 0;
  bool acta = k < nc, actb = false;
  if (!acta) { ca.ba = 0; ca.bb = 0; }
  float4 paa = pl[ca.ba * 32], pab = pl[ca.bb * 32], pba, pbb;
  float2 qaa = ql[ca.ba * 32], qab = ql[ca.bb * 32], qba, qbb;
  int isl = -1;
  bool skip = false, all_solved = true, done = !live || nc == 0, stop = false;
  float min_separation = 0.0f;
  auto half_step = [&](PcRegs& cur, bool act, float4& pa, float4& pb, float2& qa, float2& qb, PcRegs& nxt, bool& nact,
                       float4& npa, float4& npb, float2& nqa, float2& nqb) {
    if (++k == ncm) k = 0;
    if (!resident) cp_async_wait<POS_RING - 2>();
    nxt = pc_load(rl + ((resident ? k : ((p + 1) & (POS_RING - 1))) * PC_Q) * 32);
    nact = (p + 1 < total) && (k < nc);
    if (!nact) { nxt.ba = 0; nxt.bb = 0; }
    npa = pl[nxt.ba * 32]; npb = pl[nxt.bb * 32];
    nqa = ql[nxt.ba * 32]; nqb = ql[nxt.bb * 32];
    if (!resident) fetch();
    if (act && !done) {
      if (cur.isl != isl) {  // island boundary: close the previous island, open the next
        if (isl >= 0 && !skip) {
          if (min_separation >= -3.0f * B2G_LINEAR_SLOP) B.isl_flags[x.at(B.NB, isl)] |= 1; else all_solved = false;
        }
        isl = cur.isl;
        skip = (B.isl_flags[x.at(B.NB, isl)] & 1) != 0;
        min_separation = 0.0f;
      }
      if (!skip) {
        PosState s;
        s.c_a = v2(pa.x, pa.y); s.a_a = pa.z; s.q_a.s = qa.x; s.q_a.c = qa.y;
        s.c_b = v2(pb.x, pb.y); s.a_b = pb.z; s.q_b.s = qb.x; s.q_b.c = qb.y;
        min_separation = solve_position_one(s, cur.p0, cur.p1, cur.p2, cur.p3, cur.type, cur.cnt, cur.ra, cur.rb, min_separation);
        pa = make_float4(s.c_a.x, s.c_a.y, s.a_a, 0.0f);
        pb = make_float4(s.c_b.x, s.c_b.y, s.a_b, 0.0f);
        qa = make_float2(s.q_a.s, s.q_a.c);
        qb = make_float2(s.q_b.s, s.q_b.c);
        pl[cur.ba * 32] = pa; ql[cur.ba * 32] = qa;
        pl[cur.bb * 32] = pb; ql[cur.bb * 32] = qb;
        if (nxt.ba == cur.ba) { npa = pa; nqa = qa; } else if (nxt.ba == cur.bb) { npa = pb; nqa = qb; }
        if (nxt.bb == cur.ba) { npb = pa; nqb = qa; } else if (nxt.bb == cur.bb) { npb = pb; nqb = qb; }
      }
    }
    if (k == 0) {  // end of a sweep: close the last island, test the early exit of this world
      if (!done) {
        if (isl >= 0 && !skip) {
          if (min_separation >= -3.0f * B2G_LINEAR_SLOP) B.isl_flags[x.at(B.NB, isl)] |= 1; else all_solved = false;
        }
        if (all_solved) done = true;
        isl = -1;
        skip = false;
        all_solved = true;
        min_separation = 0.0f;
      }
      if (__all_sync(0xffffffffu, done)) stop = true;
    }
    ++p;
  };
  while (p < total && !stop) {
    half_step(ca, acta, paa, pab, qaa, qab, cb, actb, pba, pbb, qba, qbb);
    if (p >= total || stop) break;
    half_step(cb, actb, pba, pbb, qba, qbb, ca, acta, paa, pab, qaa, qab);
  }
  cp_async_wait<0>();
  __syncwarp();
  if (live && nc > 0) {
    for (int b = 0; b < B.NB; ++b) {
      const int bi = x.at(B.NB, b);
      const float4 p = pos[b * 32 + lane];
      const float2 q = rot[b * 32 + lane];
      float4 op = B.b_pos[bi];
      float4 r = B.b_rot[bi];
      op.x = p.x; op.y = p.y; op.z = p.z;
      r.x = q.x; r.y = q.y;
      B.b_pos[bi] = op;
      B.b_rot[bi] = r;
    }
  }
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
  return (size_t)POS_RING * PC_Q * 32 * 16 + (size_t)(NB + 1) * 32 * (16 + 8) + (size_t)(NB + 1) * 32;
}

__global__ void __launch_bounds__(32) position_sl_kernel(const Batch B, const StepParams sp) {
  extern __shared__ float4 smem4[];
  const unsigned long long t_start = B.timeline ? global_ns() : 0ull;
  float4* ring = smem4;                                          // [POS_RING][PC_Q][32]
  float4* pos = smem4 + POS_RING * PC_Q * 32;                    // [NB + 1][32]: c.x c.y a -; row NB is scratch
  float2* rot = (float2*)(pos + (size_t)(B.NB + 1) * 32);        // [NB + 1][32]: sin a, cos a
  unsigned char* tab = (unsigned char*)(rot + (size_t)(B.NB + 1) * 32);  // [NB + 1][32]: island solved; row NB is scratch
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
  if (ncm == 0 || sp.position_iterations <= 0) return;
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
#pragma unroll
  for (int p = 0; p < POS_RING; ++p) fetch(p);
  PcRegs ca, cb;
  bool acta = false, actb = false, fa = true, fb = true;
  float4 paa, pab, pba, pbb;
  float2 qaa, qab, qba, qbb;
  int k = 0, p = 0;
  bool done = !live || nc == 0, all_solved = true;
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
    // end of the sweep: a world is finished when every island passed the exit test
    if (!done && all_solved) done = true;
    all_solved = true;
    min_separation = 0.0f;
    if (__all_sync(0xffffffffu, done)) break;
  }
  cp_async_wait<0>();
  __syncwarp();
  if (live && nc > 0) {
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

// ==========================================================================================
// Level-scheduled Gauss-Seidel: SCHED_G lanes cooperate on one world (used for the position stage; the
// velocity stage measured faster in the one-lane-per-world form above, whose register forwarding
// removes every shared-memory round trip from the chain — see profiles/).
//
// The island kernel packs each world's island contacts into rounds of at most SCHED_G constraints that
// share no movable body (b2g_island_smem.cuh).  Constraints of one round commute exactly, so the G
// lanes of a world solve them concurrently; rounds are separated by __syncwarp().  A CTA is one warp =
// 32 / SCHED_G worlds; the worlds of a 32-world memory block are spread over SCHED_G CTAs.  Per world
// the dependent chain shrinks from "constraints per sweep" to "rounds per sweep" (Pyramid: 400 -> 296).
// Body state lives in shared memory as [body][world] rows, each
// lane streams the records of its own constraints through a private cp.async ring.
// Worlds without a schedule (WS_SCHED_ROUNDS == -1) are solved by lane 0 in list order.
// ==========================================================================================
constexpr int ML_WPC = 32 / SCHED_G;  // worlds per CTA
constexpr int ML_RING = 4;

inline size_t position_ml_smem_bytes(int NB) { return (size_t)NB * ML_WPC * (16 + 8 + 4) + (size_t)ML_RING * PC_Q * 32 * 16 + 8 * 32 * 4; }

// [body][world] rows: the lanes of one slot g read one contiguous row segment each, conflict-free
__device__ __forceinline__ int ml_col(int body, int wq) { return body * ML_WPC + wq; }

// island contact handled by lane slot g of a world in round r (or -1)
__device__ __forceinline__ int ml_item(const int* sched_w, int rounds, int nc, int r, int g) {
  if (rounds < 0) return (g == 0 && r < nc) ? r : -1;  // no schedule: list order on lane 0
  return r < rounds ? sched_w[(size_t)(r * SCHED_G + g) * 32] : -1;
}

// Position iterations, level-scheduled.  Per-island state of a sweep lives in a shared-memory table:
// the most negative separation seen (as float bits: for negative floats larger bits = more negative, so
// the lanes merge their contributions with atomicMax), or ML_SOLVED once the island passed the
// reference's exit test min_separation >= -3 * linear_slop (b2_island_private.rs:257-274).
constexpr unsigned ML_SOLVED = 0xffffffffu;

__global__ void __launch_bounds__(32) position_ml_kernel(const Batch B, const StepParams sp) {
  extern __shared__ float4 smem4[];
  const unsigned long long t_start = B.timeline ? global_ns() : 0ull;
  float4* ring = smem4;                                   // [ML_RING][PC_Q][32]
  float4* pos = smem4 + ML_RING * PC_Q * 32;              // [NB][ML_WPC]: c.x c.y a -
  float2* rot = (float2*)(pos + (size_t)B.NB * ML_WPC);   // [NB][ML_WPC]: sin a, cos a
  unsigned* tab = (unsigned*)(rot + (size_t)B.NB * ML_WPC);  // [NB][ML_WPC] per island (see above)
  const int lane = threadIdx.x;
  const int g = lane / ML_WPC, wq = lane % ML_WPC;
  const int wb = blockIdx.x / SCHED_G + B.wb_first;
  const int wl = (blockIdx.x % SCHED_G) * ML_WPC + wq;
  const int w = wb * 32 + wl;
  const bool live = w < B.n_worlds;
  WIdx x;
  x.wb = wb; x.wl = wl; x.LB = 32;
  Ws ws = ws_of(B, x);
  const int nc = live ? ws[WS_ISL_CONTACTS] : 0;
  const int nisl = live ? ws[WS_ISL_COUNT] : 0;
  const int rounds_w = live ? ws[WS_SCHED_ROUNDS] : 0;
  int rlen = nc == 0 ? 0 : (rounds_w < 0 ? (nc < SCHED_MIN_ROUNDS ? SCHED_MIN_ROUNDS : nc) : rounds_w);
  const int rm = __reduce_max_sync(0xffffffffu, rlen);
  if (rm == 0 || sp.position_iterations <= 0) return;
  if (live) {
    for (int b = g; b < B.NB; b += SCHED_G) {
      const float4 p = B.b_pos[x.at(B.NB, b)];
      const float4 r = B.b_rot[x.at(B.NB, b)];
      pos[ml_col(b, wq)] = p;
      rot[ml_col(b, wq)] = make_float2(r.x, r.y);
    }
    for (int i = g; i < nisl; i += SCHED_G) tab[ml_col(i, wq)] = (B.isl_flags[x.at(B.NB, i)] & 1) ? ML_SOLVED : 0u;
  }
  __syncwarp();
  const int* sched_w = B.sched + (size_t)wb * B.NC * SCHED_G * 32 + wl;
  const float4* src = B.pc + (size_t)wb * B.NC * PC_Q * 32 + wl;
  float4* rl = ring + lane;
  const int total = sp.position_iterations * rm;
  // Schedule entries travel through a small shared-memory queue filled by 4-byte cp.async copies that
  // ride in the same commit groups as the record fetches, ML_AHEAD rounds before they are needed, so no
  // register ever waits on a global load of the schedule (the record stream evicts it from L2).
  constexpr int ML_AHEAD = 4, ML_Q = 8;
  int* iq = (int*)(tab + (size_t)B.NB * ML_WPC) + lane;  // [ML_Q][32]
  const bool have_sched = rounds_w >= 0;
  auto item_now = [&](int at_pos, int rr) -> int {  // entry of position at_pos (round rr) once it has landed
    if (at_pos >= total) return -1;
    if (!have_sched) return (g == 0 && rr < nc) ? rr : -1;
    return rr < rounds_w ? iq[(at_pos & (ML_Q - 1)) * 32] : -1;
  };
  auto item_request = [&](int at_pos, int rr) {       // start the copy of the entry of position at_pos
    if (have_sched && at_pos < total && rr < rounds_w)
      cp_async4(&iq[(at_pos & (ML_Q - 1)) * 32], sched_w + (size_t)(rr * SCHED_G + g) * 32);
  };
  auto fetch = [&](int at_pos, int k) {
    if (k >= 0) {
      float4* dst = rl + ((at_pos & (ML_RING - 1)) * PC_Q) * 32;
      const float4* s = src + (size_t)k * PC_Q * 32;
#pragma unroll
      for (int q = 0; q < PC_Q; ++q) cp_async16(dst + q * 32, s + q * 32);
    }
  };
  // prologue: entries of positions 0 .. RING-2+AHEAD, then the records of positions 0 .. RING-2
  int rq = 0;  // round of the next entry to request
  for (int i = 0; i < ML_RING - 1 + ML_AHEAD; ++i) { item_request(i, rq); if (++rq == rm) rq = 0; }
  cp_async_commit();
  cp_async_wait<0>();
  int rf = 0;  // round of the next record to fetch
  for (int i = 0; i < ML_RING - 1; ++i) { fetch(i, item_now(i, rf)); cp_async_commit(); if (++rf == rm) rf = 0; }
  int r = 0;
  bool done = !live || nc == 0;
  // software pipeline: the record of the next round is read into registers while this round is solved
  cp_async_wait<ML_RING - 2>();
  PcRegs cn;
  int kn = item_now(0, 0);
  if (kn >= 0) cn = pc_load(rl);
  for (int p = 0; p < total; ++p) {
    const int k = kn;
    PcRegs c = cn;
    const bool act = k >= 0 && !done;
    float4 pa, pb;
    float2 qa, qb;
    bool solve = false;
    if (act) {
      solve = tab[ml_col(c.isl, wq)] != ML_SOLVED;
      pa = pos[ml_col(c.ba, wq)]; pb = pos[ml_col(c.bb, wq)];
      qa = rot[ml_col(c.ba, wq)]; qb = rot[ml_col(c.bb, wq)];
    }
    // refill: record of position p + RING - 1 (its entry landed AHEAD rounds ago), entry of position p + RING - 1 + AHEAD
    fetch(p + ML_RING - 1, item_now(p + ML_RING - 1, rf));
    if (++rf == rm) rf = 0;
    item_request(p + ML_RING - 1 + ML_AHEAD, rq);
    if (++rq == rm) rq = 0;
    cp_async_commit();
    // next round's record into registers (group of position p + 1 is complete when <= RING-2 groups are pending)
    cp_async_wait<ML_RING - 2>();
    {
      int rn = r + 1;
      if (rn == rm) rn = 0;
      kn = item_now(p + 1, rn);
      if (kn >= 0) cn = pc_load(rl + (((p + 1) & (ML_RING - 1)) * PC_Q) * 32);
    }
    if (solve) {
      PosState s;
      s.c_a = v2(pa.x, pa.y); s.a_a = pa.z; s.q_a.s = qa.x; s.q_a.c = qa.y;
      s.c_b = v2(pb.x, pb.y); s.a_b = pb.z; s.q_b.s = qb.x; s.q_b.c = qb.y;
      const float ms = solve_position_one(s, c.p0, c.p1, c.p2, c.p3, c.type, c.cnt, c.ra, c.rb, 0.0f);
      if (ms < 0.0f) atomicMax(&tab[ml_col(c.isl, wq)], __float_as_uint(ms));
      if (c.p0.x != 0.0f || c.p0.y != 0.0f) {  // immovable bodies may be shared inside a round: never written
        pos[ml_col(c.ba, wq)] = make_float4(s.c_a.x, s.c_a.y, s.a_a, 0.0f);
        rot[ml_col(c.ba, wq)] = make_float2(s.q_a.s, s.q_a.c);
      }
      if (c.p0.z != 0.0f || c.p0.w != 0.0f) {
        pos[ml_col(c.bb, wq)] = make_float4(s.c_b.x, s.c_b.y, s.a_b, 0.0f);
        rot[ml_col(c.bb, wq)] = make_float2(s.q_b.s, s.q_b.c);
      }
    }
    __syncwarp();
    if (++r == rm) {  // end of a sweep: per island exit test, then the early exit of the whole world
      r = 0;
      bool open_left = false;
      if (!done) {
        for (int i = g; i < nisl; i += SCHED_G) {
          const unsigned m = tab[ml_col(i, wq)];
          if (m == ML_SOLVED) continue;
          const float min_separation = m == 0u ? 0.0f : __uint_as_float(m);
          if (min_separation >= -3.0f * B2G_LINEAR_SLOP) tab[ml_col(i, wq)] = ML_SOLVED;
          else { tab[ml_col(i, wq)] = 0u; open_left = true; }
        }
      }
      // a world is finished when none of its lanes still holds an unsolved island
      const unsigned open_mask = __ballot_sync(0xffffffffu, open_left);
      bool world_open = false;
#pragma unroll
      for (int gg = 0; gg < SCHED_G; ++gg) world_open = world_open || ((open_mask >> (gg * ML_WPC + wq)) & 1u);
      if (!world_open) done = true;
      __syncwarp();
      if (__all_sync(0xffffffffu, done)) break;
    }
  }
  cp_async_wait<0>();
  __syncwarp();
  if (live && nc > 0) {
    for (int b = g; b < B.NB; b += SCHED_G) {
      const int bi = x.at(B.NB, b);
      const float4 p = pos[ml_col(b, wq)];
      const float2 q = rot[ml_col(b, wq)];
      float4 op = B.b_pos[bi];
      float4 rr = B.b_rot[bi];
      op.x = p.x; op.y = p.y; op.z = p.z;
      rr.x = q.x; rr.y = q.y;
      B.b_pos[bi] = op;
      B.b_rot[bi] = rr;
    }
    for (int i = g; i < nisl; i += SCHED_G)
      if (tab[ml_col(i, wq)] == ML_SOLVED) B.isl_flags[x.at(B.NB, i)] |= 1;
  }
  timeline_record(B, 2, t_start);
}

// Warm start + velocity iterations, level-scheduled: the same round structure and pipeline as
// position_ml_kernel (schedule queue in shared memory, next round's record read into registers while
// the current round is solved).
inline size_t velocity_ml_smem_bytes(int NB) { return (size_t)NB * ML_WPC * 16 + (size_t)ML_RING * VC_Q * 32 * 16 + 8 * 32 * 4; }

__global__ void __launch_bounds__(32) velocity_ml_kernel(const Batch B, const StepParams sp) {
  extern __shared__ float4 smem4[];
  float4* ring = smem4;                          // [ML_RING][VC_Q][32] one private column per lane
  float4* vel = smem4 + ML_RING * VC_Q * 32;     // [NB][ML_WPC]
  const int lane = threadIdx.x;
  const int g = lane / ML_WPC, wq = lane % ML_WPC;
  const int wb = blockIdx.x / SCHED_G + B.wb_first;
  const int wl = (blockIdx.x % SCHED_G) * ML_WPC + wq;
  const int w = wb * 32 + wl;
  const bool live = w < B.n_worlds;
  WIdx x;
  x.wb = wb; x.wl = wl; x.LB = 32;
  Ws ws = ws_of(B, x);
  const int nc = live ? ws[WS_ISL_CONTACTS] : 0;
  const int rounds_w = live ? ws[WS_SCHED_ROUNDS] : 0;
  const int wflags = live ? ws[WS_FLAGS] : 0;
  const bool warm = (wflags & B2GPU_WORLD_WARM_STARTING) != 0;
  const bool block = (wflags & B2GPU_WORLD_BLOCK_SOLVE) != 0;
  int rlen = nc == 0 ? 0 : (rounds_w < 0 ? (nc < SCHED_MIN_ROUNDS ? SCHED_MIN_ROUNDS : nc) : rounds_w);
  const int rm = __reduce_max_sync(0xffffffffu, rlen);
  if (rm == 0) return;
  if (live)
    for (int b = g; b < B.NB; b += SCHED_G) vel[ml_col(b, wq)] = B.b_vel[x.at(B.NB, b)];
  __syncwarp();
  const int* sched_w = B.sched + (size_t)wb * B.NC * SCHED_G * 32 + wl;
  const float4* src = B.vc + (size_t)wb * B.NC * VC_Q * 32 + wl;
  float4* q6_out = B.vc + (size_t)wb * B.NC * VC_Q * 32 + 6 * 32 + wl;
  float4* rl = ring + lane;
  const int sweeps = 1 + sp.velocity_iterations;  // sweep 0 = warm start
  const int total = sweeps * rm;
  constexpr int ML_AHEAD = 4, ML_Q = 8;
  int* iq = (int*)(vel + (size_t)B.NB * ML_WPC) + lane;  // [ML_Q][32] schedule queue
  const bool have_sched = rounds_w >= 0;
  auto item_now = [&](int at_pos, int rr) -> int {
    if (at_pos >= total) return -1;
    if (!have_sched) return (g == 0 && rr < nc) ? rr : -1;
    return rr < rounds_w ? iq[(at_pos & (ML_Q - 1)) * 32] : -1;
  };
  auto item_request = [&](int at_pos, int rr) {
    if (have_sched && at_pos < total && rr < rounds_w)
      cp_async4(&iq[(at_pos & (ML_Q - 1)) * 32], sched_w + (size_t)(rr * SCHED_G + g) * 32);
  };
  auto fetch = [&](int at_pos, int k) {
    if (k >= 0) {
      float4* dst = rl + ((at_pos & (ML_RING - 1)) * VC_Q) * 32;
      const float4* s = src + (size_t)k * VC_Q * 32;
#pragma unroll
      for (int q = 0; q < VC_Q; ++q) cp_async16(dst + q * 32, s + q * 32);
    }
  };
  int rq = 0, rf = 0;
  for (int i = 0; i < ML_RING - 1 + ML_AHEAD; ++i) { item_request(i, rq); if (++rq == rm) rq = 0; }
  cp_async_commit();
  cp_async_wait<0>();
  for (int i = 0; i < ML_RING - 1; ++i) { fetch(i, item_now(i, rf)); cp_async_commit(); if (++rf == rm) rf = 0; }
  cp_async_wait<ML_RING - 2>();
  VcRegs cn;
  int kn = item_now(0, 0);
  if (kn >= 0) cn = vc_load(rl);
  int r = 0, sweep = 0;
  for (int pos = 0; pos < total; ++pos) {
    const int k = kn;
    VcRegs c = cn;
    const bool act = k >= 0 && (sweep > 0 || warm) && c.cnt > 0;
    float4 va, vb;
    if (act) {
      va = vel[ml_col(c.ba, wq)];
      vb = vel[ml_col(c.bb, wq)];
    }
    fetch(pos + ML_RING - 1, item_now(pos + ML_RING - 1, rf));
    if (++rf == rm) rf = 0;
    item_request(pos + ML_RING - 1 + ML_AHEAD, rq);
    if (++rq == rm) rq = 0;
    cp_async_commit();
    cp_async_wait<ML_RING - 2>();
    {
      int rn = r + 1;
      if (rn == rm) rn = 0;
      kn = item_now(pos + 1, rn);
      if (kn >= 0) cn = vc_load(rl + (((pos + 1) & (ML_RING - 1)) * VC_Q) * 32);
    }
    if (act) {
      VelState s;
      s.v_a = v2(va.x, va.y); s.w_a = va.z;
      s.v_b = v2(vb.x, vb.y); s.w_b = vb.z;
      if (sweep == 0) {
        warm_start_one(s, c.q0, c.q1, c.q2, c.q6, c.q7, c.cnt);
      } else {
        solve_velocity_one(s, c.q0, c.q1, c.q2, c.q3, c.q4, c.q5, c.q6, c.q7, c.cnt, block);
        q6_out[(size_t)k * VC_Q * 32] = c.q6;
      }
      // immovable bodies (zero inverse mass and inertia) can be shared by the constraints of a round: never written
      if (c.q7.x != 0.0f || c.q7.y != 0.0f) vel[ml_col(c.ba, wq)] = make_float4(s.v_a.x, s.v_a.y, s.w_a, 0.0f);
      if (c.q7.z != 0.0f || c.q7.w != 0.0f) vel[ml_col(c.bb, wq)] = make_float4(s.v_b.x, s.v_b.y, s.w_b, 0.0f);
    }
    __syncwarp();
    if (++r == rm) { r = 0; ++sweep; }
  }
  cp_async_wait<0>();
  __syncwarp();
  if (live)
    for (int b = g; b < B.NB; b += SCHED_G) B.b_vel[x.at(B.NB, b)] = vel[ml_col(b, wq)];
}


// ------------------------------------------------------------------------------------------
// Level-scheduled velocity stage, straight-line form (experiment, solver='levels2').  Two lanes per world
// solve the (at most two) constraints of a round of the island kernel's schedule concurrently; a round is one
// basic block per lane (scratch rows for empty slots and immovable bodies, unconditional refills with clamped
// addresses, vote one round ahead), body velocities travel between the two lanes through shared memory with
// one __syncwarp() per round.  Fewer rounds than constraints (Pyramid 296 vs 400), but every round pays the
// store -> barrier -> load round trip that register forwarding avoids in velocity_sl_kernel.
// ------------------------------------------------------------------------------------------
inline size_t velocity_ml2_smem_bytes(int NB) {
  return (size_t)(NB + SCHED_G) * ML_WPC * 16 + (size_t)ML_RING * VC_Q * 32 * 16 + 8 * 32 * 4;
}
__device__ __forceinline__ void st_global_v4_if(float4* p, float4 v, bool on) {
  asm volatile("{\n.reg .pred q;\nsetp.ne.s32 q, %5, 0;\n@q st.global.v4.f32 [%0], {%1, %2, %3, %4};\n}\n" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w), "r"((int)on)
               : "memory");
}

__global__ void __launch_bounds__(32) velocity_ml2_kernel(const Batch B, const StepParams sp) {
  extern __shared__ float4 smem4[];
  const unsigned long long t_start = B.timeline ? global_ns() : 0ull;
  float4* ring = smem4;                          // [ML_RING][VC_Q][32] one private column per lane
  float4* vel = smem4 + ML_RING * VC_Q * 32;     // [NB + SCHED_G][ML_WPC]; rows NB + g are lane slot g's scratch
  const int lane = threadIdx.x;
  const int g = lane / ML_WPC, wq = lane % ML_WPC;
  const int wb = blockIdx.x / SCHED_G + B.wb_first;
  const int wl = (blockIdx.x % SCHED_G) * ML_WPC + wq;
  const int w = wb * 32 + wl;
  const bool live = w < B.n_worlds;
  WIdx x;
  x.wb = wb; x.wl = wl; x.LB = 32;
  Ws ws = ws_of(B, x);
  const int nc = live ? ws[WS_ISL_CONTACTS] : 0;
  const int rounds_w = live ? ws[WS_SCHED_ROUNDS] : 0;
  const int wflags = live ? ws[WS_FLAGS] : 0;
  const bool warm = (wflags & B2GPU_WORLD_WARM_STARTING) != 0;
  const bool block = (wflags & B2GPU_WORLD_BLOCK_SOLVE) != 0;
  const bool have_sched = rounds_w >= 0;
  const int rlen = nc == 0 ? 0 : (have_sched ? rounds_w : nc);  // rounds of this world (list order: one per contact)
  const int rm = __reduce_max_sync(0xffffffffu, rlen);
  if (rm == 0) return;
  if (live)
    for (int b = g; b < B.NB; b += SCHED_G) vel[ml_col(b, wq)] = B.b_vel[x.at(B.NB, b)];
  const int scratch = B.NB + g;
  vel[ml_col(scratch, wq)] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  __syncwarp();
  const int* sched_w = B.sched + (size_t)wb * B.NC * SCHED_G * 32 + wl;
  const float4* src = B.vc + (size_t)wb * B.NC * VC_Q * 32 + wl;
  float4* q6_out = B.vc + (size_t)wb * B.NC * VC_Q * 32 + 6 * 32 + wl;
  float4* rl = ring + lane;
  float4* vcol = vel + wq;                       // + body * ML_WPC
  constexpr int AHEAD = 4, QN = 8;
  int* iq = (int*)(vel + (size_t)(B.NB + SCHED_G) * ML_WPC) + lane;  // [QN][32] schedule queue
  const int n_warm = __any_sync(0xffffffffu, warm) ? rm : 0;
  const int total = n_warm + sp.velocity_iterations * rm;
  // entry of the position whose round is rr (landed in the queue): island contact index or -1
  auto item_now = [&](int at_pos, int rr) -> int {
    const int q = iq[(at_pos & (QN - 1)) * 32];
    const int listed = (g == 0 && rr < nc) ? rr : -1;
    return have_sched ? (rr < rounds_w ? q : -1) : listed;
  };
  auto item_request = [&](int at_pos, int rr) {  // unconditional: a round past the end re-reads round 0
    const int r2 = (have_sched && rr < rounds_w) ? rr : 0;
    cp_async4(&iq[(at_pos & (QN - 1)) * 32], sched_w + (size_t)(r2 * SCHED_G + g) * 32);
  };
  auto fetch = [&](int at_pos, int k) {          // unconditional: an empty slot re-reads record 0
    cp_async_record(rl + ((at_pos & (ML_RING - 1)) * VC_Q) * 32, src + (size_t)(k < 0 ? 0 : k) * VC_Q * 32);
  };
  auto next_round = [&](int rr) { return rr + 1 == rm ? 0 : rr + 1; };
  int rq = 0, rf = 0;
  for (int i = 0; i < ML_RING - 1 + AHEAD; ++i) { item_request(i, rq); rq = next_round(rq); }
  cp_async_commit();
  cp_async_wait<0>();
  // cp.async groups alternate record / entry from here on (the record copy commits its own group)
  for (int i = 0; i < ML_RING - 1; ++i) { fetch(i, item_now(i, rf)); rf = next_round(rf); cp_async_commit(); }
  cp_async_wait<2 * (ML_RING - 1) - 2>();  // record 0
  int r = 0, pos = 0;
  int ka = item_now(0, 0), kb = -1;
  VcRegs ca = vc_load(rl), cb;
  bool acta = ka >= 0 && ca.cnt > 0 && (warm || n_warm == 0), actb = false;
  bool fa = __all_sync(0xffffffffu, !acta || (ca.cnt == 2 && block)), fb = true;
  ca.ba = acta ? ca.ba : scratch;
  ca.bb = acta ? ca.bb : scratch;
  cb = ca;
  auto round = [&](auto WARM, auto FAST, VcRegs& c, const int kc, const bool act, VcRegs& cn, int& kn, bool& nact, bool& nfast) {
    // -- this round's bodies (the barrier of the previous round made the partner lane's stores visible)
    const float4 va = vcol[c.ba * ML_WPC], vb = vcol[c.bb * ML_WPC];
    // -- refill: record of position pos + RING - 1 (its entry landed AHEAD rounds ago), entry of pos + RING - 1 + AHEAD
    fetch(pos + ML_RING - 1, item_now(pos + ML_RING - 1, rf));
    rf = next_round(rf);
    item_request(pos + ML_RING - 1 + AHEAD, rq);
    rq = next_round(rq);
    cp_async_commit();
    // -- next round's record into the other register set: 2 (RING - 1) - 1 groups are younger than it
    cp_async_wait<2 * (ML_RING - 1) - 1>();
    r = next_round(r);
    kn = item_now(pos + 1, r);
    cn = vc_load(rl + (((pos + 1) & (ML_RING - 1)) * VC_Q) * 32);
    nact = kn >= 0 && cn.cnt > 0 && (warm || pos + 1 >= n_warm);
    nfast = __all_sync(0xffffffffu, !nact || (cn.cnt == 2 && block));
    cn.ba = nact ? cn.ba : scratch;
    cn.bb = nact ? cn.bb : scratch;
    // -- the reference's arithmetic
    VelState s;
    s.v_a = v2(va.x, va.y); s.w_a = va.z;
    s.v_b = v2(vb.x, vb.y); s.w_b = vb.z;
    if (decltype(WARM)::value) {
      warm_start_one(s, c.q0, c.q1, c.q2, c.q6, c.q7, decltype(FAST)::value ? 2 : c.cnt);
    } else {
      float4 q6 = c.q6;
      if (decltype(FAST)::value)
        solve_velocity_one(s, c.q0, c.q1, c.q2, c.q3, c.q4, c.q5, q6, c.q7, 2, true);
      else
        solve_velocity_one(s, c.q0, c.q1, c.q2, c.q3, c.q4, c.q5, q6, c.q7, c.cnt, block);
      st_global_v4_if(q6_out + (size_t)(kc < 0 ? 0 : kc) * VC_Q * 32, q6, act);
    }
    // immovable bodies (zero inverse mass and inertia) can be shared by the constraints of a round: never written
    const int sa = (c.q7.x != 0.0f || c.q7.y != 0.0f) ? c.ba : scratch;
    const int sb = (c.q7.z != 0.0f || c.q7.w != 0.0f) ? c.bb : scratch;
    vcol[sa * ML_WPC] = make_float4(s.v_a.x, s.v_a.y, s.w_a, va.w);
    vcol[sb * ML_WPC] = make_float4(s.v_b.x, s.v_b.y, s.w_b, vb.w);
    __syncwarp();
    ++pos;
  };
  auto step_ab = [&](auto WARM) {
    if (fa) round(WARM, std::true_type{}, ca, ka, acta, cb, kb, actb, fb);
    else round(WARM, std::false_type{}, ca, ka, acta, cb, kb, actb, fb);
  };
  auto step_ba = [&](auto WARM) {
    if (fb) round(WARM, std::true_type{}, cb, kb, actb, ca, ka, acta, fa);
    else round(WARM, std::false_type{}, cb, kb, actb, ca, ka, acta, fa);
  };
  auto run_to = [&](auto WARM, const int end) {
    while (pos + 2 <= end) {
      step_ab(WARM);
      step_ba(WARM);
    }
    if (pos < end) {  // odd count: one more round, then the register sets swap roles
      step_ab(WARM);
      ca = cb; ka = kb; acta = actb; fa = fb;
    }
  };
  run_to(std::true_type{}, n_warm);
  run_to(std::false_type{}, total);
  cp_async_wait<0>();
  __syncwarp();
  if (live)
    for (int b = g; b < B.NB; b += SCHED_G) B.b_vel[x.at(B.NB, b)] = vel[ml_col(b, wq)];
  timeline_record(B, 1, t_start);
}

}  // namespace b2g
