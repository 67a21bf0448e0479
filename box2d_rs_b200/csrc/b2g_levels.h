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
// (levels) x (latency of a level) with a barrier between levels, the constraints of a level one per thread.  The body
// state stays in the plain global arrays (L1 / L2: the CTA is one SM, its own writes are visible to it after the
// barrier).  Islands with joints keep the one-thread form (joint rows interleave with the contact rows per iteration).
//
// The functors take (cta, thread, threads); the host simulator runs them with one thread per CTA, i.e. it executes the
// constraints in LEVEL order — a CPU test that the reordering keeps every bit.
#pragma once
#include "b2g_large.h"

namespace b2g {

enum { LW_MAXG = 16, LW_LEVEL_NT = 288, LW_LEVEL_BUILD_NT = 1024, LW_LEVEL_MIN_DEFAULT = 1024 };  // LW_LEVEL_NT: eight consumer warps + the producer warp
enum { LV_VQ = 8, LV_PQ = 5 };  // float4 per velocity / position record in the level-ordered copies
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

// The traversal of a giant island (LwDfsK's walk, 89 ms on the settled 100k pile: three dependent L2 misses per body — its
// row, the row's entries, the states of the bodies they name).  The walk itself is the reference's lexicographic DFS and
// stays one thread; a second warp reads along: every body the walker pushes is handed over through shared memory, and the
// helper loads that body's row, the entries the walker will read first, the states of their bodies and those bodies' own
// rows — plain loads whose values are thrown away, so that the walker finds the lines in this SM's L1 when it pops the body.  Nothing the helper
// does can change a result (it writes nothing but its queue position); a lost or overwritten hand-over is a missed prefetch.
struct LwDfsGiantK {
  Batch B;
  Large L;
  int* stack;
  struct Hook {
    int* q;     // [64] bodies pushed, a ring
    int* head;  // pushes so far
    int h;
    B2G_HD void pushed(int o) {
#if defined(__CUDA_ARCH__)
      *(volatile int*)(q + (h & 63)) = o;
      *(volatile int*)head = ++h;
#else
      (void)o;
#endif
    }
  };
  B2G_HD void operator()(int g, int tid, int nt) const {
#if defined(__CUDA_ARCH__)
    __shared__ int q[64];
    __shared__ int ctl[2];  // [0] pushes so far, [1] the walk is over
#else
    int q[1], ctl[2];
#endif
    if (g >= lv_giants(L)) return;
    const int isl = L.lv_info[g].x;
    if (tid == 0) { ctl[0] = 0; ctl[1] = 0; }
    lv_cta_sync();
    if (tid == 0) {
      Hook hook;
      hook.q = q; hook.head = &ctl[0]; hook.h = 0;
      lw_dfs_walk(B, L, stack, isl, hook);
#if defined(__CUDA_ARCH__)
      *(volatile int*)&ctl[1] = 1;
#endif
    }
#if defined(__CUDA_ARCH__)
    else if (tid >= 32) {
      const int lane = tid - 32, lanes = nt - 32;
      int tail = 0, sink = 0;
      for (;;) {
        const int head = *(volatile int*)&ctl[0];
        if (head == tail) {
          if (*(volatile int*)&ctl[1]) break;
          continue;
        }
        if (head - tail > 64) tail = head - 64;  // overrun: the oldest hand-overs are gone
        // eight lanes per handed-over body, one per entry of its row (the walker reads a row from its end): the entry, the
        // state of the body it names, and — one level further, for when that body is pushed and popped in turn — that
        // body's own row and its last entry
        for (int i = tail + (lane >> 3); i < head; i += (lanes >> 3)) {
          const int o = *(volatile int*)(q + (i & 63));
          const int2 row = L.erow[o];
          const int j = row.y - 1 - (lane & 7);
          if (j >= row.x) {
            const int2 ent = L.eadj[j];
            if (ent.y < 0 && !(ent.y & 0x40000000)) {
              const int other = ent.y & 0x3fffffff;
              sink += L.state[other];
              const int2 r2 = L.erow[other];
              if (r2.y > r2.x) sink += L.eadj[r2.y - 1].x;
            }
          }
        }
        tail = head;
      }
      if (sink == 0x7fffffff) L.lv_meta[3] = 1;  // keeps the loads alive
    }
#endif
  }
};

// Levels of one giant island: CTA g, after SolverInitK / LwVcIdxK of the step that rebuilt the islands.
//   lv_level[first + k]   level of constraint k (island order)
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
    // which bodies of a constraint move (inverse mass or inertia): flat, into lv_level, where the scan finds them as one
    // contiguous word per constraint instead of a record row per constraint
    for (int k = tid; k < n; k += nt) {
      const float4 q7 = B.vc[(size_t)(first + k) * VC_Q + 7];
      L.lv_level[first + k] = ((q7.x != 0.0f || q7.y != 0.0f) ? 1 : 0) | ((q7.z != 0.0f || q7.w != 0.0f) ? 2 : 0);
    }
#if defined(__CUDA_ARCH__)
    __shared__ int scan_at;  // constraint the scan has reached (published every eight)
    if (tid == 0) scan_at = 0;
#endif
    lv_cta_sync();
#if defined(__CUDA_ARCH__)
    if (tid >= 32 && tid < 64) {
      // a second warp reads along ahead of the scan: lv_last of the bodies of the next few hundred constraints, plain loads
      // whose values are thrown away — they fill this SM's L1 (the scan's own stores keep those lines current)
      int done = 0, sink = 0;
      for (;;) {
        const int at = *(volatile int*)&scan_at;
        if (at >= n) break;
        const int lo = max(done, at + 32), hi = min(n, at + 32 + 512);
        for (int c = lo + (tid - 32); c < hi; c += 32) {
          const int4 ix = L.vc_idx[first + c];
          sink += L.lv_last[ix.x] + L.lv_last[ix.y];
        }
        if (hi > done) done = hi;
      }
      if (sink == 0x7fffffff) L.lv_meta[3] = 1;  // keeps the loads alive
    }
#endif
    if (tid == 0) {
      // the scan is a chain through lv_last (a constraint's level needs its bodies' latest levels); the indices of eight
      // constraints are requested ahead of it
      int depth = 0;
      int k = 0;
      for (; k + 8 <= n; k += 8) {
#if defined(__CUDA_ARCH__)
        *(volatile int*)&scan_at = k;
#endif
        int4 ix[8];
        int mv[8];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 8; ++j) { ix[j] = L.vc_idx[first + k + j]; mv[j] = L.lv_level[first + k + j]; }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 8; ++j) {
          const bool mov_a = (mv[j] & 1) != 0, mov_b = (mv[j] & 2) != 0;
          const int la = mov_a ? L.lv_last[ix[j].x] : 0, lb = mov_b ? L.lv_last[ix[j].y] : 0;
          const int lvl = imax(la, lb);
          L.lv_level[first + k + j] = lvl;
          if (mov_a) L.lv_last[ix[j].x] = lvl + 1;
          if (mov_b) L.lv_last[ix[j].y] = lvl + 1;
          depth = imax(depth, lvl + 1);
        }
      }
      for (; k < n; ++k) {
        const int4 ix = L.vc_idx[first + k];
        const int mv = L.lv_level[first + k];
        const bool mov_a = (mv & 1) != 0, mov_b = (mv & 2) != 0;
        const int la = mov_a ? L.lv_last[ix.x] : 0, lb = mov_b ? L.lv_last[ix.y] : 0;
        const int lvl = imax(la, lb);
        L.lv_level[first + k] = lvl;
        if (mov_a) L.lv_last[ix.x] = lvl + 1;
        if (mov_b) L.lv_last[ix.y] = lvl + 1;
        depth = imax(depth, lvl + 1);
      }
      L.lv_info[g].w = depth;
#if defined(__CUDA_ARCH__)
      *(volatile int*)&scan_at = n;
#endif
    }
    lv_cta_sync();
    const int depth = L.lv_info[g].w;
    for (int k = tid; k < n; k += nt) lv_fetch_add(&L.lv_count[base + L.lv_level[first + k]], 1);
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
      const int lvl = L.lv_level[first + k];
      const int pos = L.lv_start[base + lvl] + lv_fetch_add(&L.lv_count[base + lvl], 1);
      L.lv_order[first + pos] = first + k;
    }
  }
};

// Level-ordered copies of what a sweep reads, every step after SolverInitK / LwVcIdxK: position i of the level order gets its
// constraint's indices, velocity record (q0..q7) and position record (p0..p4).  The sweeps then read
// CONTIGUOUS streams in the order they visit them — no indirection between a position and its data.
struct LwLevelGatherK {  // flat over the island contact slots
  Batch B;
  Large L;
  int n;
  B2G_HD void operator()(int i) const {
    if (i >= n || !L.lv_isl_giant[B.c_isl[i]]) return;
    const int k = L.lv_order[i];
    L.lv_ix[i] = L.vc_idx[k];
    const float4* v = B.vc + (size_t)k * VC_Q;
    float4* dv = L.lv_vrec + (size_t)i * LV_VQ;
    for (int q = 0; q < LV_VQ; ++q) dv[q] = v[q];
    const float4* c = B.pc + (size_t)k * PC_Q;
    float4* dc = L.lv_prec + (size_t)i * LV_PQ;
    for (int q = 0; q < LV_PQ; ++q) dc[q] = c[q];
  }
};

// ------------------------------------------------------------------------------------------
// The sweep engine.  A sweep costs (levels) x (latency of a level): a level must be nothing but "read two bodies, solve,
// write two bodies, barrier".  What was measured on the way (profiles/r02_levels.md): with every thread running its own
// prefetch pipeline the bookkeeping of the idle threads WAS the level time, and so was a producer warp that worked level by
// level inside the CTA barrier; a register load in flight shares a scoreboard with the body loads of the level and makes
// their first use wait for L2 (~550 cycles); cp.async.cg and prefetch.global.L1 do not fill L1, a plain load does
// (41 cycles afterwards), and a body written by this SM is an L1 hit for the next reader.  So:
//   * the CONSUMER warps sweep level by level with a barrier of their own (bar.sync 1): bounds of the next level requested a
//     level ahead, position -> thread round the warps (lv_slot), the constraint from the ring when it is there (plain loads
//     when not: the consumers never wait for the producer), its bodies by plain loads, arithmetic, stores, barrier;
//   * one PRODUCER warp runs DECOUPLED from the levels: in chunks it streams the level-ordered indices and records of the
//     positions ahead through a ring in shared memory (cp.async, contiguous source; the ring is indexed by the running
//     position of the sweeps, pass * n + position) and then touches the bodies of those constraints with plain loads, which
//     fills this SM's L1.  It follows the consumers' published position: a ring slot is reused only when the consumers are
//     past it, and a record is requested only when its previous visit — one pass, n positions, earlier: it rewrites the
//     record's impulses — is behind a consumer barrier:  position < consumers' position + min(ring, n).
// Hand-over: producer  cp.async -> wait_group 0 -> __threadfence_block -> ready = position;  consumers  ready ->
// __threadfence_block -> ring.  Host simulator: one thread plays both roles, the producer keeps the ring full.
// ------------------------------------------------------------------------------------------
enum { LV_RING = 256, LV_PRODUCER = 32, LV_CHUNK = 48 };
#if defined(__CUDA_ARCH__)
#define LW_CP4(dst, src) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory")
#else
#define LW_CP4(dst, src) (*(int*)(dst) = *(const int*)(src))
#endif
template <int Q> struct LvRing {
  enum { RS = Q + 1 };  // record stride in float4 (one of padding: wide levels read with a stride of several records)
  float4* rec;  // [LV_RING][RS]
  int4* ix;     // [LV_RING]
  int* k;       // [LV_RING]
  int* ctl;     // [0] positions below this are in the ring (producer), [1] position the consumers are at
  static B2G_HD size_t bytes() { return (size_t)LV_RING * (RS * 16 + 16 + 4) + 16; }
  B2G_HD void carve(float4* smem) {
    rec = smem;
    ix = (int4*)(rec + (size_t)LV_RING * RS);
    k = (int*)(ix + LV_RING);
    ctl = k + LV_RING;
  }
};
B2G_HD int lv_vload(const int* p) {
#if defined(__CUDA_ARCH__)
  return *(volatile const int*)p;
#else
  return *p;
#endif
}
B2G_HD void lv_vstore(int* p, int v) {
#if defined(__CUDA_ARCH__)
  *(volatile int*)p = v;
#else
  *p = v;
#endif
}
B2G_HD void lv_fence_block() {
#if defined(__CUDA_ARCH__)
  __threadfence_block();
#endif
}
B2G_HD void lv_consumer_sync(int nc) {
#if defined(__CUDA_ARCH__)
  asm volatile("bar.sync 1, %0;" ::"r"(nc) : "memory");
#endif
}

// POL: Q; stream(): level-ordered records of the island (Q float4 per position); touch(ix): plain loads of a constraint's
// bodies, returns something that depends on them; ring(t, rel, k, ix, rec): solve a constraint from the ring;
// direct(t, rel, k): solve it with plain loads only.
template <class POL> struct LvProducer {  // the producer's state (host simulator: advanced between the consumers' levels)
  int pf;        // positions requested (and landed) so far
  float touched, pending;
};
template <class POL>
B2G_HD void lv_produce(POL& pol, const Large& L, LvRing<POL::Q>& R, LvProducer<POL>& P, const float4* stream, int first, int n, int total_pos,
                       int cons, int lane, int lanes) {
  typedef LvRing<POL::Q> Ring;
  const int lim = imin(total_pos, cons + imin((int)LV_RING, n));
  if (lim <= P.pf) return;
  const int pieces = POL::Q + 2;
  for (int c = lane; c < (lim - P.pf) * pieces; c += lanes) {
    const int gp = P.pf + c / pieces, part = c % pieces, rel = gp % n, rs = gp & (LV_RING - 1);
    if (part < POL::Q) LW_CP16(R.rec + (size_t)rs * Ring::RS + part, stream + (size_t)rel * POL::Q + part);
    else if (part == POL::Q) LW_CP16(R.ix + rs, &L.lv_ix[first + rel]);
    else LW_CP4(R.k + rs, &L.lv_order[first + rel]);
  }
  LW_CP_COMMIT();
  P.touched += P.pending;  // the loads of the previous chunk's touch are looked at only now
  P.pending = 0.0f;
  LW_CP_WAIT0();
  lv_fence_block();
#if defined(__CUDA_ARCH__)
  __syncwarp();
#endif
  if (lane == 0) lv_vstore(&R.ctl[0], lim);
  // touch the bodies of the chunk (its indices are in the ring now): plain loads fill this SM's L1
  for (int gp = P.pf + lane; gp < lim; gp += lanes) P.pending += pol.touch(R.ix[gp & (LV_RING - 1)]);
  P.pf = lim;
}
template <class POL>
B2G_HD void lv_sweeps(POL& pol, const Large& L, float4* smem, int first, int n, int depth, int base, int passes, int tid, int nt) {
  typedef LvRing<POL::Q> Ring;
  Ring R;
  R.carve(smem);
#if defined(__CUDA_ARCH__)
  const int nc = nt - LV_PRODUCER;          // consumer threads
  const bool producer = tid >= nc;
  const int lane = tid - nc, lanes = LV_PRODUCER;
#else
  const int nc = 1;
  const int lane = 0, lanes = 1;
#endif
  const int total = passes * depth, total_pos = passes * n;
  const float4* stream = pol.stream();
  LvProducer<POL> P;
  P.pf = 0;
  P.touched = P.pending = 0.0f;
  if (tid == 0) { lv_vstore(&R.ctl[0], 0); lv_vstore(&R.ctl[1], 0); }
  lv_cta_sync();
#if defined(__CUDA_ARCH__)
  if (producer) {
    int spins = 0;
    while (P.pf < total_pos) {
      const int cons = lv_vload(&R.ctl[1]);
      if (cons >= total_pos) break;  // the consumers are through (they publish total_pos when they leave)
      lv_fence_block();
      const int lim = imin(total_pos, cons + imin((int)LV_RING, n));
      if (lim - P.pf < imin((int)LV_CHUNK, imax(1, n >> 1)) && lim < total_pos) {  // wait for room: a chunk at a time
        if (++spins > (1 << 24)) break;  // never the reason a kernel does not end: the consumers do not need the producer
        __nanosleep(100);
        continue;
      }
      lv_produce(pol, L, R, P, stream, first, n, total_pos, cons, lane, lanes);
    }
    if (P.touched + P.pending == 12345.678f) L.lv_meta[3] = 1;  // keeps the touch loads alive
    return;
  }
#endif
  // ---- consumers
  const int slot = lv_slot(tid, nc);
  int lc = 0, pbc = 0;  // level index and pass base of level t
  int s = L.lv_start[base], e = L.lv_start[base + 1];
  int rd = 0;
#if defined(B2G_LV_DEBUG) && defined(__CUDA_ARCH__)
  long long dbg_a = 0, dbg_b = 0, dbg_c = 0, dbg_t = clock64();
#endif
  for (int t = 0; t < total; ++t) {
    if ((t & 3) == 0 && tid == nc - 1) {  // every position below is behind a barrier: publish (release) the consumers' position, every 4th
                                          // level — by the thread of the LAST slot, which has a constraint only in the widest levels
                                          // (the fence costs ~200 cycles; thread 0 has a constraint in every level)
      lv_fence_block();
      lv_vstore(&R.ctl[1], pbc + s);
    }
#if !defined(__CUDA_ARCH__)
    lv_produce(pol, L, R, P, stream, first, n, total_pos, pbc + s, lane, lanes);
#endif
    const int rd_now = lv_vload(&R.ctl[0]);
    if (rd_now != rd) {  // the producer published a chunk: acquire it
      lv_fence_block();
      rd = rd_now;
    }
    // bounds of the next level, requested now
    int ln = lc + 1, pbn = pbc;
    if (ln == depth) { ln = 0; pbn += n; }
    const int sn = L.lv_start[base + ln], en = L.lv_start[base + ln + 1];
#if defined(B2G_LV_DEBUG) && defined(__CUDA_ARCH__)
    { const long long c = clock64(); dbg_a += c - dbg_t; dbg_t = c; }
#endif
    for (int rel = s + slot; rel < e; rel += nc) {
      const int gp = pbc + rel;
      if (gp < rd) {
        const int rs = gp & (LV_RING - 1);
        pol.ring(t, rel, R.k[rs], R.ix[rs], R.rec + (size_t)rs * Ring::RS);
      } else {
        pol.direct(t, rel, L.lv_order[first + rel]);
      }
    }
#if defined(B2G_LV_DEBUG) && defined(__CUDA_ARCH__)
    { const long long c = clock64(); dbg_b += c - dbg_t; dbg_t = c; }
#endif
    lv_consumer_sync(nc);
#if defined(B2G_LV_DEBUG) && defined(__CUDA_ARCH__)
    { const long long c = clock64(); dbg_c += c - dbg_t; dbg_t = c; }
#endif
    lc = ln; pbc = pbn; s = sn; e = en;
  }
#if defined(B2G_LV_DEBUG) && defined(__CUDA_ARCH__)
  if (tid == 0) {  // thread 0 takes the first position of every level: cycles / 1024 before the visit, in it, at the barrier
    atomicAdd(&L.lv_meta[1], (int)(dbg_a >> 10));
    atomicAdd(&L.lv_meta[2], (int)(dbg_b >> 10));
    atomicAdd(&L.lv_meta[3], (int)(dbg_c >> 10));
  }
#endif
  if (tid == nc - 1) lv_vstore(&R.ctl[1], total_pos);  // releases the producer, whatever it was waiting for
}
template <class POL>
B2G_HD void lv_sweeps_plain(POL& pol, const Large& L, int first, int depth, int base, int passes, int tid, int nt) {
  const int slot = lv_slot(tid, nt);
  for (int p = 0; p < passes; ++p) {
    int s = L.lv_start[base];
    for (int l = 0; l < depth; ++l) {
      const int e = L.lv_start[base + l + 1];
      for (int i = s + slot; i < e; i += nt) pol.direct(p * depth + l, i, L.lv_order[first + i]);
      s = e;
      lv_cta_sync();
    }
  }
}

// Warm start + velocity iterations of giant island g (the loops of LwVelocity7K, level by level).
struct LwLevelVelocityK {
  Batch B;
  Large L;
  StepParams sp;
  static size_t smem_bytes() { return LvRing<LV_VQ>::bytes(); }
  struct Pol {
    enum { Q = LV_VQ };
    const LwLevelVelocityK* K;
    int first, depth;
    bool warm, block;
    B2G_HD const float4* stream() const { return K->L.lv_vrec + (size_t)first * LV_VQ; }
    B2G_HD float touch(const int4 ix) const { return K->B.b_vel[ix.x].w + K->B.b_vel[ix.y].w; }
    B2G_HD void solve(int t, int rel, int k, const int4 ix, const float4 q0, const float4 q1, const float4 q2, const float4 q3,
                      const float4 q4, const float4 q5, float4 q6, const float4 q7) const {
      const Batch& B = K->B;
      const int points = ix.z & 0xff;
      if (points == 0) return;
      const float4 va = B.b_vel[ix.x], vb = B.b_vel[ix.y];  // L1: touched by the producer two levels ago, or written by this SM
      VelState s;
      s.v_a = v2(va.x, va.y); s.w_a = va.z;
      s.v_b = v2(vb.x, vb.y); s.w_b = vb.z;
      if (warm && t < depth) {
        warm_start_one(s, q0, q1, q2, q6, q7, points);
      } else {
        // the common case — two points, block solver — as its own call: the constants fold its branches away
        if (points == 2 && block) solve_velocity_one(s, q0, q1, q2, q3, q4, q5, q6, q7, 2, true);
        else solve_velocity_one(s, q0, q1, q2, q3, q4, q5, q6, q7, points, block);
        B.vc[(size_t)k * VC_Q + 6] = q6;                                 // what PostVelocityK stores into the manifold
        K->L.lv_vrec[(size_t)(first + rel) * LV_VQ + 6] = q6;            // what the next pass reads
      }
      // a body without inverse mass and inertia may sit in several constraints of a level (and in several islands): never written
      if (q7.x != 0.0f || q7.y != 0.0f) B.b_vel[ix.x] = make_float4(s.v_a.x, s.v_a.y, s.w_a, 0.0f);
      if (q7.z != 0.0f || q7.w != 0.0f) B.b_vel[ix.y] = make_float4(s.v_b.x, s.v_b.y, s.w_b, 0.0f);
    }
    B2G_HD void ring(int t, int rel, int k, const int4 ix, const float4* r) const {
      solve(t, rel, k, ix, r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]);
    }
    B2G_HD void direct(int t, int rel, int k) const {  // everything read now
      const float4* r = stream() + (size_t)rel * LV_VQ;
      solve(t, rel, k, K->L.vc_idx[k], r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]);
    }
  };
  B2G_HD void operator()(int g, int tid, int nt) const {
#if defined(__CUDA_ARCH__)
    extern __shared__ float4 lv_smem[];
    float4* smem = lv_smem;
#else
    float4 host_ring[(LvRing<LV_VQ>::RS + 2) * LV_RING + 8];
    float4* smem = host_ring;
#endif
    if (g >= lv_giants(L)) return;
    const int4 info = L.lv_info[g];
    const int first = info.y, n = info.z, depth = info.w, base = first + info.x;
    Pol pol;
    pol.K = this;
    pol.first = first;
    pol.depth = depth;
    pol.warm = (B.ws[WS_FLAGS] & B2GPU_WORLD_WARM_STARTING) != 0;
    pol.block = (B.ws[WS_FLAGS] & B2GPU_WORLD_BLOCK_SOLVE) != 0;
    const int passes = (pol.warm ? 1 : 0) + sp.velocity_iterations;
    if (tid == 0) lv_fetch_add(&B.ws[WS_ST_LEVELS], depth);  // b2gpu_step_stats.solver_levels: levels of one sweep, summed over the giant islands
    if (depth < 3 || (long long)passes * n > 0x3fffffffLL) lv_sweeps_plain(pol, L, first, depth, base, passes, tid, nt);
    else lv_sweeps(pol, L, smem, first, n, depth, base, passes, tid, nt);
  }
};

// Position iterations of giant island g (the loop of LwPosition6K, level by level; the running minimum separation of a
// sweep is reduced over the CTA: a minimum does not depend on the order).  The position records do not change during the
// stage; every sweep runs the engine for one pass.
struct LwLevelPositionK {
  Batch B;
  Large L;
  StepParams sp;
  static size_t smem_bytes() { return LvRing<LV_PQ>::bytes(); }
  struct Pol {
    enum { Q = LV_PQ };
    const LwLevelPositionK* K;
    int first;
    float ms;  // this thread's running minimum separation of the sweep
    B2G_HD const float4* stream() const { return K->L.lv_prec + (size_t)first * LV_PQ; }
    B2G_HD float touch(const int4 ix) const {
      const Batch& B = K->B;
      return B.b_pos[ix.x].w + B.b_rot[ix.x].w + B.b_pos[ix.y].w + B.b_rot[ix.y].w;
    }
    B2G_HD void solve(const int4 ix, const float4 p0, const float4 p1, const float4 p2, const float4 p3, const float4 p4) {
      const Batch& B = K->B;
      const int ba = ix.x, bb = ix.y, packed = ix.w;
      float4 pa = B.b_pos[ba], ra = B.b_rot[ba], pb = B.b_pos[bb], rb = B.b_rot[bb];  // L1: touched by the producer, or written by this SM
      PosState s;
      s.c_a = v2(pa.x, pa.y); s.a_a = pa.z; s.q_a.s = ra.x; s.q_a.c = ra.y;
      s.c_b = v2(pb.x, pb.y); s.a_b = pb.z; s.q_b.s = rb.x; s.q_b.c = rb.y;
      const int type = (packed >> 8) & 0xff, points = packed & 0xff;
      bool done = false;
      if (points == 2 && type != B2GPU_MANIFOLD_CIRCLES) {  // a face manifold with two points: the straight-line form of position_sl_kernel
        const PosState s0 = s;
        bool wide = false;
        const float m = solve_position_face2(s, p0, p1, p2, p3, type == B2GPU_MANIFOLD_FACE_A, p4.x, p4.y, ms, wide);
        if (!wide) { ms = m; done = true; }
        else s = s0;  // an angle beyond +-120 rad: the general form
      }
      if (!done) ms = solve_position_one(s, p0, p1, p2, p3, type, points, p4.x, p4.y, ms);
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
    }
    B2G_HD void ring(int, int, int, const int4 ix, const float4* r) { solve(ix, r[0], r[1], r[2], r[3], r[4]); }
    B2G_HD void direct(int, int rel, int k) {  // everything read now
      const float4* r = stream() + (size_t)rel * LV_PQ;
      solve(K->L.vc_idx[k], r[0], r[1], r[2], r[3], r[4]);
    }
  };
  B2G_HD void operator()(int g, int tid, int nt) const {
#if defined(__CUDA_ARCH__)
    extern __shared__ float4 lv_smem[];
    __shared__ float red[LW_LEVEL_NT];
    float4* smem = lv_smem;
#else
    float4 host_ring[(LvRing<LV_PQ>::RS + 2) * LV_RING + 8];
    float red[1];
    float4* smem = host_ring;
#endif
    if (g >= lv_giants(L)) return;
    const int4 info = L.lv_info[g];
    const int isl = info.x, first = info.y, n = info.z, depth = info.w, base = first + isl;
    Pol pol;
    pol.K = this;
    pol.first = first;
    for (int it = 0; it < sp.position_iterations; ++it) {
      pol.ms = 0.0f;
      if (depth < 3) lv_sweeps_plain(pol, L, first, depth, base, 1, tid, nt);
      else lv_sweeps(pol, L, smem, first, n, depth, base, 1, tid, nt);
      red[tid] = pol.ms;
      lv_cta_sync();
      for (int h = 256; h > 0; h >>= 1) {  // LW_LEVEL_NT <= 512
        if (tid < h && tid + h < nt) red[tid] = fmin_sel(red[tid], red[tid + h]);
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
