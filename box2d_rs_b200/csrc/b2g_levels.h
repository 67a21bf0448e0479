// b2g_levels.h — level-scheduled Gauss-Seidel sweeps of a GIANT island (large-world mode).
//
// The reference sweeps an island's contact constraints in list order (b2_contact_solver_private.rs:228-266 warm start,
// :268-583 velocity, b2_island_private.rs:257-274 position).  Two constraints that share no MOVABLE body commute: each
// reads and writes only its own impulses and the state of its two bodies, and a body without inverse mass and inertia
// is never written.  So the sweep is a dependency DAG — constraint k waits for the previous constraint of each of its
// movable bodies — and any schedule that respects it produces the reference's bits.  profiles/r02_dag_depth.md measured
// the DAG of the configs' giant islands: one sweep of the settled 100k pile (295k constraints) is ~39k levels deep.
//
// One thread per island (LwVelocity7K / LwPosition6K) pays (visits) x (latency of a visit).  Here ONE CTA sweeps the
// island level by level: level(k) = max over its movable bodies of (level of the body's previous constraint + 1),
// computed once per island rebuild by a sequential scan (LwLevelBuildK) and counting-sorted; a sweep then costs
// (levels) x (latency of a level) with a __syncthreads between levels, the constraints of a level one per thread.  The
// body state stays in the plain global arrays (L1 / L2: the CTA is one SM, its own writes are visible to it after the
// barrier).  Islands with joints keep the one-thread form (joint rows interleave with the contact rows per iteration).
//
// The functors take (cta, thread, threads); the host simulator runs them with one thread per CTA, i.e. it executes the
// constraints in LEVEL order — a CPU test that the reordering keeps every bit.
#pragma once
#include "b2g_large.h"

namespace b2g {

enum { LW_MAXG = 16, LW_LEVEL_NT = 256, LW_LEVEL_BUILD_NT = 1024, LW_LEVEL_MIN_DEFAULT = 1024 };
// L.lv_meta: [0] giant islands chosen at the last island rebuild (may exceed LW_MAXG: the surplus keeps the one-thread form)

B2G_HD void lv_cta_sync() {
#if defined(__CUDA_ARCH__)
  __syncthreads();
#endif
}
B2G_HD int lv_fetch_add(int* p, int v) {
#if defined(__CUDA_ARCH__)
  return atomicAdd(p, v);
#else
  const int o = *p;
  *p = o + v;
  return o;
#endif
}
B2G_HD int lv_giants(const Large& L) { return imin(L.lv_meta[0], (int)LW_MAXG); }
// Which position of a level a thread takes: positions go round the WARPS first (position j -> warp j mod W, lane j / W), so
// that the few constraints of a typical level (the settled 100k pile averages 7.6) sit in different warps — each runs its
// visit without divergence (one- and two-point manifolds take different paths) and on its own scheduler.
B2G_HD int lv_slot(int tid, int nt) {
  const int nw = nt >> 5;
  return nw > 0 ? (tid & 31) * nw + (tid >> 5) : tid;
}
// lv_level entry of a constraint: its level, and per body whether the body's state is FRESH at that level — written by the
// level just before (or level 0, which follows the last level of the previous sweep), so it can only be read after the
// barrier.  A body that is not fresh was last written at least two levels earlier (or never: immovable) and may be
// requested one level ahead.
enum { LV_LEVEL_MASK = 0x1fffffff, LV_FRESH_A = 1 << 29, LV_FRESH_B = 1 << 30, LV_IX_FRESH_A = 1 << 8, LV_IX_FRESH_B = 1 << 9 };
B2G_HD int lv_pack(int level, bool fresh_a, bool fresh_b) { return level | (fresh_a ? (int)LV_FRESH_A : 0) | (fresh_b ? (int)LV_FRESH_B : 0); }

struct LwLevelResetK {  // one thread, before the selection
  Large L;
  B2G_HD void operator()(int) const { L.lv_meta[0] = 0; }
};
struct LwGiantSelectK {  // flat over islands: which islands take the level-scheduled form
  Batch B;
  Large L;
  int n_islands, level_min;
  B2G_HD void operator()(int isl) const {
    if (isl >= n_islands) return;
    const int4 rg = B.isl_range[isl];
    const int n = rg.w - rg.z;
    bool giant = level_min > 0 && n >= level_min && !lw_island_has_joints(B, isl);
    if (giant) {
      const int slot = lv_fetch_add(&L.lv_meta[0], 1);
      if (slot < LW_MAXG) L.lv_info[slot] = make_int4(isl, rg.z, n, 0);
      else giant = false;
    }
    L.lv_isl_giant[isl] = giant ? 1 : 0;
  }
};

// Levels of one giant island: CTA g, after SolverInitK / LwVcIdxK of the step that rebuilt the islands.
//   lv_level[first + k]   level of constraint k (island order) | LV_FRESH_A / LV_FRESH_B
//   lv_start[first + isl + l]  first position of level l in lv_order (island-relative), l = 0 .. depth (at depth: n);
//                              first + isl grows by more than an island's level count from one island to the next
//   lv_order[first + p]   constraint (absolute index) at position p of the level order
struct LwLevelBuildK {
  Batch B;
  Large L;
  B2G_HD void operator()(int g, int tid, int nt) const {
#if defined(__CUDA_ARCH__)
    __shared__ int part[LW_LEVEL_BUILD_NT];
#else
    int part[1];
#endif
    if (g >= lv_giants(L)) return;
    const int4 info = L.lv_info[g];
    const int isl = info.x, first = info.y, n = info.z, base = first + isl;
    const int4 rg = B.isl_range[isl];
    for (int i = rg.x + tid; i < rg.y; i += nt) L.lv_last[B.isl_body[i]] = 0;
    for (int i = tid; i <= n; i += nt) L.lv_count[base + i] = 0;
    lv_cta_sync();
    if (tid == 0) {
      // the scan is a chain through lv_last (a constraint's level needs its bodies' latest levels); the indices and masses of
      // four constraints are requested ahead of it
      int depth = 0;
      int k = 0;
      for (; k + 4 <= n; k += 4) {
        int4 ix[4];
        float4 q7[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 4; ++j) { ix[j] = L.vc_idx[first + k + j]; q7[j] = B.vc[(size_t)(first + k + j) * VC_Q + 7]; }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 4; ++j) {
          const bool mov_a = q7[j].x != 0.0f || q7[j].y != 0.0f, mov_b = q7[j].z != 0.0f || q7[j].w != 0.0f;
          const int la = mov_a ? L.lv_last[ix[j].x] : 0, lb = mov_b ? L.lv_last[ix[j].y] : 0;
          const int lvl = imax(la, lb);
          L.lv_level[first + k + j] = lv_pack(lvl, mov_a && (la == lvl), mov_b && (lb == lvl));
          if (mov_a) L.lv_last[ix[j].x] = lvl + 1;
          if (mov_b) L.lv_last[ix[j].y] = lvl + 1;
          depth = imax(depth, lvl + 1);
        }
      }
      for (; k < n; ++k) {
        const int4 ix = L.vc_idx[first + k];
        const float4 q7 = B.vc[(size_t)(first + k) * VC_Q + 7];
        const bool mov_a = q7.x != 0.0f || q7.y != 0.0f, mov_b = q7.z != 0.0f || q7.w != 0.0f;
        const int la = mov_a ? L.lv_last[ix.x] : 0, lb = mov_b ? L.lv_last[ix.y] : 0;
        const int lvl = imax(la, lb);
        L.lv_level[first + k] = lv_pack(lvl, mov_a && (la == lvl), mov_b && (lb == lvl));
        if (mov_a) L.lv_last[ix.x] = lvl + 1;
        if (mov_b) L.lv_last[ix.y] = lvl + 1;
        depth = imax(depth, lvl + 1);
      }
      L.lv_info[g].w = depth;
    }
    lv_cta_sync();
    const int depth = L.lv_info[g].w;
    for (int k = tid; k < n; k += nt) lv_fetch_add(&L.lv_count[base + (L.lv_level[first + k] & LV_LEVEL_MASK)], 1);
    lv_cta_sync();
    // exclusive scan of the level sizes: a contiguous chunk per thread, the chunk sums scanned by thread 0
    const int per = (depth + 1 + nt - 1) / nt;
    const int c0 = imin(tid * per, depth + 1), c1 = imin(c0 + per, depth + 1);
    int sum = 0;
    for (int l = c0; l < c1; ++l) sum += L.lv_count[base + l];
    part[tid] = sum;
    lv_cta_sync();
    if (tid == 0) {
      int run = 0;
      for (int t = 0; t < nt; ++t) { const int s = part[t]; part[t] = run; run += s; }
    }
    lv_cta_sync();
    int run = part[tid];
    for (int l = c0; l < c1; ++l) {
      const int c = L.lv_count[base + l];
      L.lv_start[base + l] = run;
      L.lv_count[base + l] = 0;  // becomes the fill cursor of the level
      run += c;
    }
    lv_cta_sync();
    for (int k = tid; k < n; k += nt) {
      const int lvl = L.lv_level[first + k] & LV_LEVEL_MASK;
      const int pos = L.lv_start[base + lvl] + lv_fetch_add(&L.lv_count[base + lvl], 1);
      L.lv_order[first + pos] = first + k;
    }
  }
};

// Indices of the giant islands' constraints in level order, every step after LwVcIdxK (the point counts in vc_idx follow the
// manifolds): a sweep reads (constraint, bodies) of position i with two independent loads instead of a chain of three.
struct LwLevelIdxK {  // flat over the island contact slots
  Batch B;
  Large L;
  int n;
  B2G_HD void operator()(int i) const {
    if (i >= n || !L.lv_isl_giant[B.c_isl[i]]) return;
    const int k = L.lv_order[i];
    int4 ix = L.vc_idx[k];
    const int lv = L.lv_level[k];
    ix.z |= ((lv & LV_FRESH_A) ? (int)LV_IX_FRESH_A : 0) | ((lv & LV_FRESH_B) ? (int)LV_IX_FRESH_B : 0);
    L.lv_ix[i] = ix;
  }
};

// A sweep costs (levels) x (latency of a level), so everything a level needs that does not depend on the level before it is
// requested ahead: the bounds of level t+3, the (constraint, bodies) indices of level t+2 and the constraint record of level
// t+1 are in flight while level t is solved; after the barrier only the two body loads stand before the arithmetic.  A
// record's impulses are rewritten by its own visit one pass (= depth levels) earlier: with depth >= 3 that store is at least
// two barriers old when the record is requested.  Shallower islands take the plain loop.
struct LwLevelVelocityK {
  Batch B;
  Large L;
  StepParams sp;
  B2G_HD void solve(int k, const int4 ix, const float4 q0, const float4 q1, const float4 q2, const float4 q3, const float4 q4,
                    const float4 q5, float4 q6, const float4 q7, bool warm_pass, bool block, const float4 ea, const float4 eb) const {
    const int points = ix.z & 0xff;
    if (points == 0) return;
    // a fresh body is read now, after the barrier; the others were requested a level ago (ea / eb)
    const float4 va = (ix.z & LV_IX_FRESH_A) ? B.b_vel[ix.x] : ea, vb = (ix.z & LV_IX_FRESH_B) ? B.b_vel[ix.y] : eb;
    VelState s;
    s.v_a = v2(va.x, va.y); s.w_a = va.z;
    s.v_b = v2(vb.x, vb.y); s.w_b = vb.z;
    if (warm_pass) {
      warm_start_one(s, q0, q1, q2, q6, q7, points);
    } else {
      // the common case — two points, block solver — as its own call: the constants fold its branches away
      if (points == 2 && block) solve_velocity_one(s, q0, q1, q2, q3, q4, q5, q6, q7, 2, true);
      else solve_velocity_one(s, q0, q1, q2, q3, q4, q5, q6, q7, points, block);
      B.vc[(size_t)k * VC_Q + 6] = q6;
    }
    // a body without inverse mass and inertia may sit in several constraints of a level (and in several islands): never written
    if (q7.x != 0.0f || q7.y != 0.0f) B.b_vel[ix.x] = make_float4(s.v_a.x, s.v_a.y, s.w_a, 0.0f);
    if (q7.z != 0.0f || q7.w != 0.0f) B.b_vel[ix.y] = make_float4(s.v_b.x, s.v_b.y, s.w_b, 0.0f);
  }
  B2G_HD void visit(int k, bool warm_pass, bool block) const {  // everything read now
    const LwVcRec r = lw_load_vc(B.vc, k);
    int4 ix = L.vc_idx[k];
    ix.z |= LV_IX_FRESH_A | LV_IX_FRESH_B;
    const float4 none = make_float4(0, 0, 0, 0);
    solve(k, ix, r.q0, r.q1, r.q2, r.q3, r.q4, r.q5, r.q6, r.q7, warm_pass, block, none, none);
  }
  B2G_HD void operator()(int g, int tid, int nt) const {
    if (g >= lv_giants(L)) return;
    const int4 info = L.lv_info[g];
    const int first = info.y, depth = info.w, base = first + info.x;
    const int slot = lv_slot(tid, nt);
    const bool warm = (B.ws[WS_FLAGS] & B2GPU_WORLD_WARM_STARTING) != 0;
    const bool block = (B.ws[WS_FLAGS] & B2GPU_WORLD_BLOCK_SOLVE) != 0;
    const int passes = (warm ? 1 : 0) + sp.velocity_iterations;
    if (tid == 0) lv_fetch_add(&B.ws[WS_ST_LEVELS], depth);  // b2gpu_step_stats.solver_levels: levels of one sweep, summed over the giant islands
    if (depth < 3) {
      for (int p = 0; p < passes; ++p) {
        int s = L.lv_start[base];
        for (int l = 0; l < depth; ++l) {
          const int e = L.lv_start[base + l + 1];
          for (int i = s + slot; i < e; i += nt) visit(L.lv_order[first + i], warm && p == 0, block);
          s = e;
          lv_cta_sync();
        }
      }
      return;
    }
    const long long total = (long long)passes * depth;
    // level t: bounds (s0, e0), this thread's first constraint k0 / ix0 with its record; t+1: (s1, e1), k1 / ix1; t+2: (s2, e2)
    int s0 = L.lv_start[base], e0 = L.lv_start[base + 1], s1 = e0, e1 = L.lv_start[base + 2], s2 = e1, e2 = L.lv_start[base + 3];
    int l3 = 3 < depth ? 3 : 0;  // level index of t+3
    int k0 = -1, k1 = -1;
    int4 ix0 = make_int4(0, 0, 0, 0), ix1 = ix0;
    if (s0 + slot < e0) { k0 = L.lv_order[first + s0 + slot]; ix0 = L.lv_ix[first + s0 + slot]; }
    if (s1 + slot < e1) { k1 = L.lv_order[first + s1 + slot]; ix1 = L.lv_ix[first + s1 + slot]; }
    float4 q0 = make_float4(0, 0, 0, 0), q1 = q0, q2 = q0, q3 = q0, q4 = q0, q5 = q0, q6 = q0, q7 = q0, ea = q0, eb = q0;
    if (k0 >= 0) {
      const float4* r = B.vc + (size_t)k0 * VC_Q;
      q0 = r[0]; q1 = r[1]; q2 = r[2]; q3 = r[3]; q4 = r[4]; q5 = r[5]; q6 = r[6]; q7 = r[7];
      ix0.z |= LV_IX_FRESH_A | LV_IX_FRESH_B;  // the first level of the stage: nothing was requested ahead
    }
    for (long long t = 0; t < total; ++t) {
      // requests for the levels ahead
      const int s3 = L.lv_start[base + l3], e3 = L.lv_start[base + l3 + 1];
      int k2 = -1;
      int4 ix2 = make_int4(0, 0, 0, 0);
      if (s2 + slot < e2) { k2 = L.lv_order[first + s2 + slot]; ix2 = L.lv_ix[first + s2 + slot]; }
      float4 n0 = make_float4(0, 0, 0, 0), n1 = n0, n2 = n0, n3 = n0, n4 = n0, n5 = n0, n6 = n0, n7 = n0;
      if (k1 >= 0) {
        const float4* r = B.vc + (size_t)k1 * VC_Q;
        n0 = r[0]; n1 = r[1]; n2 = r[2]; n3 = r[3]; n4 = r[4]; n5 = r[5]; n6 = r[6]; n7 = r[7];
      }
      float4 fa = make_float4(0, 0, 0, 0), fb = fa;  // the bodies of level t+1 that level t does not write
      if (k1 >= 0 && !(ix1.z & LV_IX_FRESH_A)) fa = B.b_vel[ix1.x];
      if (k1 >= 0 && !(ix1.z & LV_IX_FRESH_B)) fb = B.b_vel[ix1.y];
      // this level
      const bool warm_pass = warm && t < depth;
      if (k0 >= 0) solve(k0, ix0, q0, q1, q2, q3, q4, q5, q6, q7, warm_pass, block, ea, eb);
      for (int i = s0 + slot + nt; i < e0; i += nt) visit(L.lv_order[first + i], warm_pass, block);
      lv_cta_sync();
      k0 = k1; ix0 = ix1; q0 = n0; q1 = n1; q2 = n2; q3 = n3; q4 = n4; q5 = n5; q6 = n6; q7 = n7; ea = fa; eb = fb;
      k1 = k2; ix1 = ix2;
      s0 = s1; e0 = e1; s1 = s2; e1 = e2; s2 = s3; e2 = e3;
      if (++l3 == depth) l3 = 0;
    }
  }
};

// Position iterations of giant island g (the loop of LwPosition6K, level by level; the running minimum separation of a
// sweep is reduced over the CTA: a minimum does not depend on the order).  The position records do not change during the
// stage, so the same requests ahead are always safe.
struct LwLevelPositionK {
  Batch B;
  Large L;
  StepParams sp;
  B2G_HD float solve(const int4 ix, const float4 p0, const float4 p1, const float4 p2, const float4 p3, const float4 p4,
                     float min_separation, float4 pa, float4 ra, float4 pb, float4 rb) const {
    const int ba = ix.x, bb = ix.y, packed = ix.w;
    // a fresh body is read now, after the barrier; the others were requested a level ago
    if (ix.z & LV_IX_FRESH_A) { pa = B.b_pos[ba]; ra = B.b_rot[ba]; }
    if (ix.z & LV_IX_FRESH_B) { pb = B.b_pos[bb]; rb = B.b_rot[bb]; }
    PosState s;
    s.c_a = v2(pa.x, pa.y); s.a_a = pa.z; s.q_a.s = ra.x; s.q_a.c = ra.y;
    s.c_b = v2(pb.x, pb.y); s.a_b = pb.z; s.q_b.s = rb.x; s.q_b.c = rb.y;
    const int type = (packed >> 8) & 0xff, points = packed & 0xff;
    bool done = false;
    if (points == 2 && type != B2GPU_MANIFOLD_CIRCLES) {  // a face manifold with two points: the straight-line form of position_sl_kernel
      const PosState s0 = s;
      bool wide = false;
      const float ms = solve_position_face2(s, p0, p1, p2, p3, type == B2GPU_MANIFOLD_FACE_A, p4.x, p4.y, min_separation, wide);
      if (!wide) { min_separation = ms; done = true; }
      else s = s0;  // an angle beyond +-120 rad: the general form
    }
    if (!done) min_separation = solve_position_one(s, p0, p1, p2, p3, type, points, p4.x, p4.y, min_separation);
    if (p0.x != 0.0f || p0.y != 0.0f) {
      pa.x = s.c_a.x; pa.y = s.c_a.y; pa.z = s.a_a; ra.x = s.q_a.s; ra.y = s.q_a.c;
      B.b_pos[ba] = pa;
      B.b_rot[ba] = ra;
    }
    if (p0.z != 0.0f || p0.w != 0.0f) {
      pb.x = s.c_b.x; pb.y = s.c_b.y; pb.z = s.a_b; rb.x = s.q_b.s; rb.y = s.q_b.c;
      B.b_pos[bb] = pb;
      B.b_rot[bb] = rb;
    }
    return min_separation;
  }
  B2G_HD float visit(int k, float min_separation) const {  // everything read now
    const LwPcRec r = lw_load_pc(B.pc, k);
    int4 ix = L.vc_idx[k];
    ix.z |= LV_IX_FRESH_A | LV_IX_FRESH_B;
    const float4 none = make_float4(0, 0, 0, 0);
    return solve(ix, r.p0, r.p1, r.p2, r.p3, r.p4, min_separation, none, none, none, none);
  }
  // one sweep; returns this thread's minimum separation
  B2G_HD float sweep(int first, int depth, int base, int tid, int nt) const {
    float ms = 0.0f;
    const int slot = lv_slot(tid, nt);
    if (depth < 3) {
      int s = L.lv_start[base];
      for (int l = 0; l < depth; ++l) {
        const int e = L.lv_start[base + l + 1];
        for (int i = s + slot; i < e; i += nt) ms = visit(L.lv_order[first + i], ms);
        s = e;
        lv_cta_sync();
      }
      return ms;
    }
    int s0 = L.lv_start[base], e0 = L.lv_start[base + 1], s1 = e0, e1 = L.lv_start[base + 2], s2 = e1, e2 = L.lv_start[base + 3];
    int k0 = -1, k1 = -1;
    int4 ix0 = make_int4(0, 0, 0, 0), ix1 = ix0;
    if (s0 + slot < e0) { k0 = L.lv_order[first + s0 + slot]; ix0 = L.lv_ix[first + s0 + slot]; }
    if (s1 + slot < e1) { k1 = L.lv_order[first + s1 + slot]; ix1 = L.lv_ix[first + s1 + slot]; }
    float4 p0 = make_float4(0, 0, 0, 0), p1 = p0, p2 = p0, p3 = p0, p4 = p0, epa = p0, era = p0, epb = p0, erb = p0;
    if (k0 >= 0) {
      const float4* r = B.pc + (size_t)k0 * PC_Q;
      p0 = r[0]; p1 = r[1]; p2 = r[2]; p3 = r[3]; p4 = r[4];
      ix0.z |= LV_IX_FRESH_A | LV_IX_FRESH_B;  // the first level of the sweep: nothing was requested ahead
    }
    for (int l = 0; l < depth; ++l) {
      const int l3 = l + 3 < depth ? l + 3 : depth - 1;  // past the sweep's end: any valid level, the request is dropped
      const int s3 = L.lv_start[base + l3], e3 = L.lv_start[base + l3 + 1];
      int k2 = -1;
      int4 ix2 = make_int4(0, 0, 0, 0);
      if (l + 2 < depth && s2 + slot < e2) { k2 = L.lv_order[first + s2 + slot]; ix2 = L.lv_ix[first + s2 + slot]; }
      float4 n0 = make_float4(0, 0, 0, 0), n1 = n0, n2 = n0, n3 = n0, n4 = n0;
      if (k1 >= 0) {
        const float4* r = B.pc + (size_t)k1 * PC_Q;
        n0 = r[0]; n1 = r[1]; n2 = r[2]; n3 = r[3]; n4 = r[4];
      }
      float4 fpa = make_float4(0, 0, 0, 0), fra = fpa, fpb = fpa, frb = fpa;  // the bodies of level l+1 that level l does not write
      if (k1 >= 0 && !(ix1.z & LV_IX_FRESH_A)) { fpa = B.b_pos[ix1.x]; fra = B.b_rot[ix1.x]; }
      if (k1 >= 0 && !(ix1.z & LV_IX_FRESH_B)) { fpb = B.b_pos[ix1.y]; frb = B.b_rot[ix1.y]; }
      if (k0 >= 0) ms = solve(ix0, p0, p1, p2, p3, p4, ms, epa, era, epb, erb);
      for (int i = s0 + slot + nt; i < e0; i += nt) ms = visit(L.lv_order[first + i], ms);
      lv_cta_sync();
      k0 = k1; ix0 = ix1; p0 = n0; p1 = n1; p2 = n2; p3 = n3; p4 = n4; epa = fpa; era = fra; epb = fpb; erb = frb;
      k1 = k2; ix1 = ix2;
      s0 = s1; e0 = e1; s1 = s2; e1 = e2; s2 = s3; e2 = e3;
    }
    return ms;
  }
  B2G_HD void operator()(int g, int tid, int nt) const {
#if defined(__CUDA_ARCH__)
    __shared__ float red[LW_LEVEL_NT];
#else
    float red[1];
#endif
    if (g >= lv_giants(L)) return;
    const int4 info = L.lv_info[g];
    const int isl = info.x, first = info.y, depth = info.w, base = first + isl;
    for (int it = 0; it < sp.position_iterations; ++it) {
      const float ms = sweep(first, depth, base, tid, nt);
      red[tid] = ms;
      lv_cta_sync();
      for (int h = nt >> 1; h > 0; h >>= 1) {
        if (tid < h) red[tid] = fmin_sel(red[tid], red[tid + h]);
        lv_cta_sync();
      }
      const float m = red[0];
      lv_cta_sync();
      if (m >= -3.0f * B2G_LINEAR_SLOP) {  // b2_island_private.rs:257-274
        if (tid == 0) B.isl_flags[isl] |= 1;
        break;
      }
    }
  }
};

}  // namespace b2g
