// b2g_step.h — the per-step hot path of B2world::step as data-parallel stages.
//
// Reference call tree (box2d-rs): src/private/dynamics/b2_world.rs:903-959 (step) ->
//   b2_contact_manager.rs(private):83-171 (collide) -> b2_contact.rs(private):103-218 (update)
//   b2_world.rs(private):356-531 (solve: island DFS, island.solve, synchronize_fixtures, find_new_contacts)
//   b2_island_private.rs:129-328, b2_contact_solver_private.rs:20-730
//   b2_fixture.rs(private):147-173, b2_dynamic_tree.rs(private):109-168, src/b2_broad_phase.rs:200-249
//   b2_contact_manager.rs(private):178-302 (add_pair), :24-78 (destroy)
//
// Two kinds of stages:
//   * flat stages: one thread per (element, world) — narrowphase, integration, constraint setup,
//     AABB synchronisation.  Thread id == memory index of the element's array (b2g_common.h), so
//     every access of a warp is one contiguous segment.
//   * ordered stages: one thread per world — the parts whose result depends on the reference's
//     iteration order (island DFS, the Gauss-Seidel sweeps, tree re-insertion, pair reporting,
//     contact creation).  Lanes of a warp are 32 worlds; parallelism comes from the batch.
// Every stage is a functor with a B2G_HD operator()(int tid) so the same code is launched as a
// CUDA kernel by b2g_runtime.cu and stepped on the host by the test-only simulator (tests/hostsim).
#pragma once
#include "b2g_common.h"
#include "b2g_narrow.h"
#include "b2g_distance.h"
#include "b2g_tree.h"
#include "b2g_joint.h"

#if defined(__CUDA_ARCH__)
#define B2G_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define B2G_ATOMIC_OR(p, v) atomicOr((p), (v))
#define B2G_ATOMIC_MAX(p, v) atomicMax((p), (v))
#define B2G_ATOMIC_MIN(p, v) atomicMin((p), (v))
#else
#define B2G_ATOMIC_ADD(p, v) (*(p) += (v))
#define B2G_ATOMIC_OR(p, v) (*(p) |= (v))
#define B2G_ATOMIC_MAX(p, v) (*(p) = (*(p) > (v) ? *(p) : (v)))
#define B2G_ATOMIC_MIN(p, v) (*(p) = (*(p) < (v) ? *(p) : (v)))
#endif

namespace b2g {

// ws counter += 1 from many threads.  In a single large world (LB = 1) every thread of a flat stage hits the SAME
// word (100k proxies leaving their fat boxes in one step: ncu showed SyncFixturesK three times slower than on a
// batch of nine times as many proxies), so lanes that target one address elect a leader that adds their count.
B2G_HD void counter_inc(int* p) {
#if defined(__CUDA_ARCH__)
  const unsigned active = __activemask();
  const unsigned peers = __match_any_sync(active, (unsigned long long)p);
  if ((threadIdx.x & 31) == (unsigned)(__ffs((int)peers) - 1)) atomicAdd(p, __popc(peers));
#else
  *p += 1;
#endif
}

// internal contact flag bits (never leave the device)
enum { CF_DESTROY = 0x100, CF_SKIPPED = 0x200, CF_WOKE = 0x400, CF_INTERNAL = 0xff00 };

struct Ws {
  int* p;
  int stride;
  B2G_HD int& operator[](int slot) const { return p[slot * stride]; }
};
B2G_HD Ws ws_of(const Batch& B, const WIdx& x) {
  Ws s;
  s.p = B.ws + x.at(WS_COUNT, 0);
  s.stride = x.LB;
  return s;
}

// Decode a flat thread id into (world, element) for an array of capacity N; false if out of range.
B2G_HD bool flat_decode(const Batch& B, int tid, int N, int& w, int& i) {
  int wl = tid & (B.LB - 1);
  int rest = tid >> B.lb_shift;
  i = rest % N;
  int wb = rest / N + B.wb_first;
  w = (wb << B.lb_shift) + wl;
  return wb < B.wb_first + B.wb_count && w < B.n_worlds;
}

B2G_HD Xf load_xf(const Batch& B, const WIdx& x, int b) {
  float4 t = B.b_xf[x.at(B.NB, b)];
  Xf xf;
  xf.p = v2(t.x, t.y);
  xf.q.s = t.z;
  xf.q.c = t.w;
  return xf;
}
B2G_HD Box load_box(const float4* a, int i) {
  float4 t = a[i];
  Box b;
  b.lo = v2(t.x, t.y);
  b.hi = v2(t.z, t.w);
  return b;
}
B2G_HD bool filter_should_collide(const b2gpu_fixture_rec& a, const b2gpu_fixture_rec& b) {  // b2_world_callbacks.rs(private):6-18
  if (a.group_index == b.group_index && a.group_index != 0) return a.group_index > 0;
  return (a.mask_bits & b.category_bits) != 0 && (a.category_bits & b.mask_bits) != 0;
}
B2G_HD bool body_should_collide(int flags_a, int flags_b) {  // b2_body.rs(private):391-398; the joint test (:400-413) is joints_prevent_collision
  return body_type(flags_a) == B2GPU_DYNAMIC_BODY || body_type(flags_b) == B2GPU_DYNAMIC_BODY;
}

// ------------------------------------------------------------------------------------------
// collide: one contact (b2_contact_manager_collide loop body + B2contact::update).
// `awake_a/awake_b` are the AWAKE bits the reference's sequential loop would see at this contact.
// Wake-ups are recorded in b_wake marks (merged after the pass) so the flat pass never races
// with the activity test of another contact.
// ------------------------------------------------------------------------------------------
// `wake_idx` (large-world mode, b2g_large.h): when given, the ordered pass is evaluated out of order — a body
// counts as woken "by then" when a contact with a larger index (visited earlier by the reference's newest-first
// loop) woke it, and wake-ups are recorded as the largest such index.
B2G_HD void collide_one(const Batch& B, const WIdx& x, const Ws& ws, int c, int* b_wake, bool ordered_pass, int* wake_idx = nullptr) {
  const int ci = x.at(B.NC, c);
  int flags = B.c_flags[ci];
  const int4 fx = B.c_fix[ci];
  const b2gpu_fixture_rec& fa = B.fixtures[fx.x];
  const b2gpu_fixture_rec& fb = B.fixtures[fx.y];
  const int ba = fa.body, bb = fb.body;
  const int bai = x.at(B.NB, ba), bbi = x.at(B.NB, bb);
  const int bfa = B.b_flags[bai], bfb = B.b_flags[bbi];
  if (!ordered_pass) {
    flags &= ~CF_INTERNAL;
    if (flags & B2GPU_CONTACT_FILTER) {
      if (!body_should_collide(bfb, bfa) || joints_prevent_collision(B, bb, ba) || !filter_should_collide(fa, fb)) {
        B.c_flags[ci] = flags | CF_DESTROY;
        counter_inc(&ws[WS_EV_DESTROY]);
        return;
      }
      flags &= ~B2GPU_CONTACT_FILTER;
    }
  }
  bool awake_a = (bfa & B2GPU_BODY_AWAKE) != 0, awake_b = (bfb & B2GPU_BODY_AWAKE) != 0;
  if (ordered_pass && wake_idx) {
    awake_a = awake_a || wake_idx[bai] > c;
    awake_b = awake_b || wake_idx[bbi] > c;
  } else if (ordered_pass) {
    awake_a = awake_a || b_wake[bai] != 0;
    awake_b = awake_b || b_wake[bbi] != 0;
  }
  const bool active_a = awake_a && body_type(bfa) != B2GPU_STATIC_BODY;
  const bool active_b = awake_b && body_type(bfb) != B2GPU_STATIC_BODY;
  if (!active_a && !active_b) {
    B.c_flags[ci] = flags | CF_SKIPPED;
    return;
  }
  flags &= ~CF_SKIPPED;
  const int node_a = B.proxy_s[fa.proxy_first + fx.z].z, node_b = B.proxy_s[fb.proxy_first + fx.w].z;
  if (!box_overlap(load_box(B.n_aabb, x.at(B.NN, node_a)), load_box(B.n_aabb, x.at(B.NN, node_b)))) {
    B.c_flags[ci] = flags | CF_DESTROY;
    counter_inc(&ws[WS_EV_DESTROY]);
    return;
  }
  // ---- B2contact::update
  if (!(flags & B2GPU_CONTACT_ENABLED)) ws[WS_TOPO_DIRTY] = 1;
  flags |= B2GPU_CONTACT_ENABLED;
  const bool was_touching = (flags & B2GPU_CONTACT_TOUCHING) != 0;
  bool touching = false;
  if (fa.is_sensor || fb.is_sensor) {
    // sensors don't generate manifolds; touching = GJK overlap of the two child shapes (b2_contact.rs(private):149-163)
    touching = test_overlap_shapes(&B.shapes[fa.shape_first + fx.z], &B.shapes[fb.shape_first + fx.w], load_xf(B, x, ba), load_xf(B, x, bb));
    int4 m3 = B.c_m3[ci];
    m3.w = 0;
    B.c_m3[ci] = m3;
  } else {
    const float4 o0 = B.c_m0[ci], o1 = B.c_m1[ci];
    const int4 o3 = B.c_m3[ci];
    Manifold m;
    manifold_clear(m);
    evaluate_contact(m, &B.shapes[fa.shape_first + fx.z], load_xf(B, x, ba), &B.shapes[fb.shape_first + fx.w], load_xf(B, x, bb));
    touching = m.count > 0;
    for (int i = 0; i < m.count; ++i) {
      m.ni[i] = 0.0f;
      m.ti[i] = 0.0f;
      if (o3.w > 0 && (uint32_t)o3.x == m.id[i]) { m.ni[i] = o0.z; m.ti[i] = o0.w; }
      else if (o3.w > 1 && (uint32_t)o3.y == m.id[i]) { m.ni[i] = o1.z; m.ti[i] = o1.w; }
    }
    B.c_m0[ci] = make_float4(m.pt[0].x, m.pt[0].y, m.ni[0], m.ti[0]);
    B.c_m1[ci] = make_float4(m.pt[1].x, m.pt[1].y, m.ni[1], m.ti[1]);
    B.c_m2[ci] = make_float4(m.ln.x, m.ln.y, m.lp.x, m.lp.y);
    B.c_m3[ci] = make_int4((int)m.id[0], (int)m.id[1], m.type, m.count);
    if (touching != was_touching) {
      // set_awake(true) on both bodies (src/b2_body.rs:783-801): static bodies ignore it
      flags |= CF_WOKE;
      if (body_type(bfa) != B2GPU_STATIC_BODY) {
        b_wake[bai] = 1;
        if (!awake_a) ws[WS_EV_WAKE] = 1;
        if (wake_idx) B2G_ATOMIC_MAX(&wake_idx[bai], c);
      }
      if (body_type(bfb) != B2GPU_STATIC_BODY) {
        b_wake[bbi] = 1;
        if (!awake_b) ws[WS_EV_WAKE] = 1;
        if (wake_idx) B2G_ATOMIC_MAX(&wake_idx[bbi], c);
      }
      ws[WS_TOPO_DIRTY] = 1;
    }
  }
  if (touching) flags |= B2GPU_CONTACT_TOUCHING; else flags &= ~B2GPU_CONTACT_TOUCHING;
  B.c_flags[ci] = flags;
}

struct CollideK {  // flat over contact slots
  Batch B;
  int* b_wake;
  B2G_HD void operator()(int tid) const {
    int w, c;
    if (!flat_decode(B, tid, B.NC, w, c)) return;
    WIdx x = widx(B, w);
    Ws ws = ws_of(B, x);
    if (c >= ws[WS_CONTACT_COUNT]) return;
    collide_one(B, x, ws, c, b_wake, false);
  }
};

// push_front of contact c on the edge lists of its two bodies (b2_contact_manager.rs(private):272-297)
B2G_HD void link_contact(const Batch& B, const WIdx& x, int* b_chead, int2* c_next, int c) {
  const int4 fx = B.c_fix[x.at(B.NC, c)];
  const int ba = B.fixtures[fx.x].body, bb = B.fixtures[fx.y].body;
  int2 nx;
  nx.x = b_chead[x.at(B.NB, ba)];
  b_chead[x.at(B.NB, ba)] = 2 * c;
  nx.y = b_chead[x.at(B.NB, bb)];
  b_chead[x.at(B.NB, bb)] = 2 * c + 1;
  c_next[x.at(B.NC, c)] = nx;
}
B2G_HD void rebuild_contact_lists(const Batch& B, const WIdx& x, int* b_chead, int2* c_next, int cc) {
  for (int b = 0; b < B.NB; ++b) b_chead[x.at(B.NB, b)] = -1;
  for (int c = 0; c < cc; ++c) link_contact(B, x, b_chead, c_next, c);
}
B2G_HD void set_awake_true(const Batch& B, const WIdx& x, int b) {
  const int bi = x.at(B.NB, b);
  const int f = B.b_flags[bi];
  if (body_type(f) == B2GPU_STATIC_BODY) return;
  B.b_flags[bi] = f | B2GPU_BODY_AWAKE;
  B.b_pos[bi].w = 0.0f;
}

// ------------------------------------------------------------------------------------------
// Ordered stage A (one thread per world): collide fix-up in list order, wake merge, contact
// destruction, island construction.
// ------------------------------------------------------------------------------------------
struct SerialAK {
  Batch B;
  int* b_wake;
  int* b_chead;   // per body: newest contact edge (2*c + side) or -1
  int2* c_next;   // per contact: next older edge of body A / body B
  int* stack;     // [NB] DFS stack
  StepParams sp;
  B2G_HD void operator()(int tid) const {
    const int w = tid + (B.wb_first << B.lb_shift);
    if (w >= B.n_worlds) return;
    WIdx x = widx(B, w);
    Ws ws = ws_of(B, x);
    if (prologue(x, ws)) islands_global(x, ws);
  }
  // Collide fix-up, wake merge, contact destruction.  Returns true when the island order must be rebuilt.
  B2G_HD bool prologue(const WIdx& x, const Ws& ws) const {
    int cc = ws[WS_CONTACT_COUNT];
    // (1) A sleeping body was woken inside collide: contacts later in list order (older) that were
    //     skipped as inactive may have become active.  Replay the list newest-first with the
    //     running awake state, like the reference's sequential loop.
    if (ws[WS_EV_WAKE]) {
      for (int b = 0; b < B.NB; ++b) b_wake[x.at(B.NB, b)] = 0;
      for (int c = cc - 1; c >= 0; --c) {
        const int ci = x.at(B.NC, c);
        const int flags = B.c_flags[ci];
        if (flags & CF_SKIPPED) {
          collide_one(B, x, ws, c, b_wake, true);
        } else if (flags & CF_WOKE) {
          const int4 fx = B.c_fix[ci];
          const int ba = B.fixtures[fx.x].body, bb = B.fixtures[fx.y].body;
          if (body_type(B.b_flags[x.at(B.NB, ba)]) != B2GPU_STATIC_BODY) b_wake[x.at(B.NB, ba)] = 1;
          if (body_type(B.b_flags[x.at(B.NB, bb)]) != B2GPU_STATIC_BODY) b_wake[x.at(B.NB, bb)] = 1;
        }
      }
      ws[WS_EV_WAKE] = 0;
    }
    // (2) merge wake marks: set_awake(true)
    if (ws[WS_TOPO_DIRTY]) {
      for (int b = 0; b < B.NB; ++b) {
        const int bi = x.at(B.NB, b);
        if (b_wake[bi]) {
          b_wake[bi] = 0;
          set_awake_true(B, x, b);
        }
      }
    }
    // (3) destroy flagged contacts after the loop (box2d-rs deviation, b2_contact_manager.rs(private):164-170),
    //     keeping the survivors in creation order.
    if (ws[WS_EV_DESTROY]) {
      if (!(sp.dt > 0.0f) && ws[WS_ISL_VALID]) {
        // collide-only step: the islands are not rebuilt, so the contacts' ISLAND bits (materialised lazily
        // from the island order, see ContactIslandFlagsK) must be baked in before the compaction below
        // shifts contact indices; afterwards the stale island order no longer describes the flags
        for (int c = 0; c < cc; ++c) B.c_flags[x.at(B.NC, c)] &= ~B2GPU_CONTACT_ISLAND;
        const int nic = ws[WS_ISL_CONTACTS];
        for (int k = 0; k < nic; ++k) B.c_flags[x.at(B.NC, B.isl_contact[x.at(B.NC, k)])] |= B2GPU_CONTACT_ISLAND;
        ws[WS_ISL_VALID] = 0;
      }
      int out = 0;
      for (int c = 0; c < cc; ++c) {
        const int ci = x.at(B.NC, c);
        const int flags = B.c_flags[ci];
        if (flags & CF_DESTROY) {
          const int4 fx = B.c_fix[ci];
          const b2gpu_fixture_rec& fa = B.fixtures[fx.x];
          const b2gpu_fixture_rec& fb = B.fixtures[fx.y];
          if (B.c_m3[ci].w > 0 && !fa.is_sensor && !fb.is_sensor) {  // b2_contact.rs(private):39-45
            set_awake_true(B, x, fa.body);
            set_awake_true(B, x, fb.body);
          }
          continue;
        }
        if (out != c) {
          const int oi = x.at(B.NC, out);
          B.c_fix[oi] = B.c_fix[ci];
          B.c_flags[oi] = flags;
          B.c_mat[oi] = B.c_mat[ci];
          B.c_m0[oi] = B.c_m0[ci];
          B.c_m1[oi] = B.c_m1[ci];
          B.c_m2[oi] = B.c_m2[ci];
          B.c_m3[oi] = B.c_m3[ci];
        }
        ++out;
      }
      ws[WS_ST_DESTROYED] += cc - out;
      cc = out;
      ws[WS_CONTACT_COUNT] = cc;
      ws[WS_EV_DESTROY] = 0;
      ws[WS_TOPO_DIRTY] = 1;
      rebuild_contact_lists(B, x, b_chead, c_next, cc);
    }
    if (!(sp.dt > 0.0f)) return false;  // collide only: islands (and their flags) stay as the last solve left them
    // Island cache: the DFS below is a pure function of the body list (types, AWAKE/ENABLED flags), the
    // per-body contact edge lists and the contacts' ENABLED/TOUCHING flags.  Every stage that changes one
    // of those raises WS_TOPO_DIRTY; when nothing changed since the previous step, last step's island
    // order (and the ISLAND flags it left behind) is exactly what the reference would rebuild.
    if (!ws[WS_TOPO_DIRTY]) {
      ws[WS_ST_ISLANDS] = ws[WS_ISL_COUNT];
      ws[WS_ST_ISL_BODIES] = ws[WS_ISL_BODIES];
      ws[WS_ST_ISL_CONTACTS] = ws[WS_ISL_CONTACTS];
      return false;
    }
    return true;
  }
  // (4) islands from global memory (any world size): b2_world.rs(private):376-507.
  B2G_HD void islands_global(const WIdx& x, const Ws& ws) const {
    const int cc = ws[WS_CONTACT_COUNT];
    bool dirty_next = false;
    // (4) islands: b2_world.rs(private):376-507.  Seeds newest body first, LIFO stack, each body's
    //     edge list newest first.
    for (int b = 0; b < B.NB; ++b) B.b_flags[x.at(B.NB, b)] &= ~B2GPU_BODY_ISLAND;
    for (int c = 0; c < cc; ++c) B.c_flags[x.at(B.NC, c)] &= ~B2GPU_CONTACT_ISLAND;
    for (int j = 0; j < B.NJ; ++j) B.j_flag[x.at(B.NJ, j)] = 0;
    int nisl = 0, nb = 0, nc = 0, nj = 0;
    for (int seed = B.NB - 1; seed >= 0; --seed) {
      const int sf = B.b_flags[x.at(B.NB, seed)];
      if (sf & B2GPU_BODY_ISLAND) continue;
      if (!(sf & B2GPU_BODY_AWAKE) || !(sf & B2GPU_BODY_ENABLED)) continue;
      if (body_type(sf) == B2GPU_STATIC_BODY) continue;
      const int body_first = nb, contact_first = nc, joint_first = nj;
      int sp_ = 0;
      stack[x.at(B.NB, sp_++)] = seed;
      B.b_flags[x.at(B.NB, seed)] = sf | B2GPU_BODY_ISLAND;
      while (sp_ > 0) {
        const int b = stack[x.at(B.NB, --sp_)];
        if (nb >= B.NIB) { ws[WS_STATUS] = B2GPU_E_CAPACITY; break; }
        B.isl_body[x.at(B.NIB, nb++)] = b;
        const int bi = x.at(B.NB, b);
        const int bf = B.b_flags[bi];
        if (body_type(bf) == B2GPU_STATIC_BODY) continue;
        if (!(bf & B2GPU_BODY_AWAKE)) dirty_next = true;  // a sleeper joined: it is a seed candidate next step
        B.b_flags[bi] = bf | B2GPU_BODY_AWAKE;
        for (int e = b_chead[bi]; e != -1;) {
          const int c = e >> 1, side = e & 1;
          const int ci = x.at(B.NC, c);
          const int2 nx = c_next[ci];
          e = side ? nx.y : nx.x;
          const int cf = B.c_flags[ci];
          if (cf & B2GPU_CONTACT_ISLAND) continue;
          if (!(cf & B2GPU_CONTACT_ENABLED) || !(cf & B2GPU_CONTACT_TOUCHING)) continue;
          const int4 fx = B.c_fix[ci];
          const b2gpu_fixture_rec& fa = B.fixtures[fx.x];
          const b2gpu_fixture_rec& fb = B.fixtures[fx.y];
          if (fa.is_sensor || fb.is_sensor) continue;
          B.isl_contact[x.at(B.NC, nc)] = c;
          B.c_isl[x.at(B.NC, nc)] = nisl;
          ++nc;
          B.c_flags[ci] = cf | B2GPU_CONTACT_ISLAND;
          const int other = side ? fa.body : fb.body;
          const int oi = x.at(B.NB, other);
          const int of = B.b_flags[oi];
          if (of & B2GPU_BODY_ISLAND) continue;
          stack[x.at(B.NB, sp_++)] = other;
          B.b_flags[oi] = of | B2GPU_BODY_ISLAND;
        }
        // joints connected to this body (b2_world.rs(private):461-483), newest edge first
        if (B.NJ > 0)
          for (int q = B.jadj_off[b]; q < B.jadj_off[b + 1]; ++q) {
            const int je = B.jadj[q], j = je >> 1;
            const int ji = x.at(B.NJ, j);
            if (B.j_flag[ji]) continue;
            const int other = (je & 1) ? B.joints[j].body_a : B.joints[j].body_b;
            const int oi = x.at(B.NB, other);
            const int of = B.b_flags[oi];
            if (!(of & B2GPU_BODY_ENABLED)) continue;  // don't simulate joints connected to disabled bodies
            B.isl_joint[x.at(B.NJ, nj++)] = j;
            B.j_flag[ji] = 1;
            if (of & B2GPU_BODY_ISLAND) continue;
            stack[x.at(B.NB, sp_++)] = other;
            B.b_flags[oi] = of | B2GPU_BODY_ISLAND;
          }
      }
      B.isl_range[x.at(B.NB, nisl)] = make_int4(body_first, nb, contact_first, nc);
      if (B.NJ > 0) B.isl_jrange[x.at(B.NB, nisl)] = make_int2(joint_first, nj);
      ++nisl;
      for (int k = body_first; k < nb; ++k) {  // static bodies may join other islands (:500-506)
        const int bi = x.at(B.NB, B.isl_body[x.at(B.NIB, k)]);
        const int bf = B.b_flags[bi];
        if (body_type(bf) == B2GPU_STATIC_BODY) B.b_flags[bi] = bf & ~B2GPU_BODY_ISLAND;
      }
    }
    ws[WS_ISL_COUNT] = nisl;
    ws[WS_ISL_BODIES] = nb;
    ws[WS_ISL_CONTACTS] = nc;
    ws[WS_ISL_JOINTS] = nj;
    ws[WS_ST_ISLANDS] = nisl;
    ws[WS_ST_ISL_BODIES] = nb;
    ws[WS_ST_ISL_CONTACTS] = nc;
    ws[WS_TOPO_DIRTY] = dirty_next ? 1 : 0;
    ws[WS_ISL_VALID] = 1;
    ws[WS_SCHED_ROUNDS] = -1;  // no level schedule from this path: the solver stages keep list order
  }
};

// ------------------------------------------------------------------------------------------
// island.solve, part 1 (b2_island_private.rs:136-170): flat over island body slots.
// ------------------------------------------------------------------------------------------
struct IntegrateK {
  Batch B;
  StepParams sp;
  B2G_HD void operator()(int tid) const {
    int w, k;
    if (!flat_decode(B, tid, B.NIB, w, k)) return;
    WIdx x = widx(B, w);
    Ws ws = ws_of(B, x);
    if (k >= ws[WS_ISL_BODIES]) return;
    const int b = B.isl_body[x.at(B.NIB, k)];
    const int bi = x.at(B.NB, b);
    const int bf = B.b_flags[bi];
    const float4 pos = B.b_pos[bi];
    B.b_pos0[bi] = make_float4(pos.x, pos.y, pos.z, 0.0f);  // c0 = c, a0 = a
    Rot q = rot_from_angle(pos.z);
    B.b_rot[bi] = make_float4(q.s, q.c, q.s, q.c);  // running rotation, rotation at (c0, a0)
    if (body_type(bf) == B2GPU_DYNAMIC_BODY) {
      const float h = sp.dt;
      float4 vel = B.b_vel[bi];
      const float4 ms = B.b_mass[bi], fo = B.b_force[bi], mi = B.b_misc[bi];
      const V2 g = v2(i2f(ws[WS_GRAVITY_X]), i2f(ws[WS_GRAVITY_Y]));
      V2 v = v2(vel.x, vel.y);
      float wv = vel.z;
      v = v + (h * ms.x) * ((fo.w * mi.x) * g + v2(fo.x, fo.y));
      wv = wv + h * ms.y * fo.z;
      v = (1.0f / (1.0f + h * mi.z)) * v;
      wv = wv * (1.0f / (1.0f + h * mi.w));
      B.b_vel[bi] = make_float4(v.x, v.y, wv, 0.0f);
    }
  }
};

// Constraint streams written by SolverInitK, one record per island contact slot k; records of one
// world block are contiguous (k-major), so the ordered stages stream them.
// velocity record, VC_Q float4:
//  0: rA0.xy rB0.xy      1: rA1.xy rB1.xy       2: normal.xy friction tangent_speed
//  3: nMass0 tMass0 bias0 nMass1     4: tMass1 bias1 K11 K12     5: K22 NM11 NM12 NM22
//  6: nImp0 tImp0 nImp1 tImp1 (mutable)   7: mA iA mB iB
//  8: (int) bodyA bodyB (vc_points | pc_points<<8 | manifold type<<16) contact
// position record, PC_Q float4:
//  0: mA iA mB iB   1: lcA.xy lcB.xy   2: local_point0.xy local_point1.xy   3: local_normal.xy local_point.xy
//  4: radiusA radiusB (int)bodyA (int)bodyB      5: (int) pc_points | type<<8, island, -, -
B2G_HD int vc_at(const Batch& B, const WIdx& x, int k, int q) { return x.at(B.NC * VC_Q, k * VC_Q + q); }
B2G_HD int pc_at(const Batch& B, const WIdx& x, int k, int q) { return x.at(B.NC * PC_Q, k * PC_Q + q); }

// B2contactSolver::new + initialize_velocity_constraints (b2_contact_solver_private.rs:20-226): flat.
struct SolverInitK {
  Batch B;
  StepParams sp;
  B2G_HD void operator()(int tid) const {
    int w, k;
    if (!flat_decode(B, tid, B.NC, w, k)) return;
    WIdx x = widx(B, w);
    Ws ws = ws_of(B, x);
    if (k >= ws[WS_ISL_CONTACTS]) return;
    const int c = B.isl_contact[x.at(B.NC, k)];
    const int ci = x.at(B.NC, c);
    const int4 fx = B.c_fix[ci];
    const b2gpu_fixture_rec& fa = B.fixtures[fx.x];
    const b2gpu_fixture_rec& fb = B.fixtures[fx.y];
    const float radius_a = B.shapes[fa.shape_first].radius, radius_b = B.shapes[fb.shape_first].radius;
    const int ba = fa.body, bb = fb.body;
    const int bai = x.at(B.NB, ba), bbi = x.at(B.NB, bb);
    const float4 mat = B.c_mat[ci];
    const float4 m0 = B.c_m0[ci], m1 = B.c_m1[ci], m2 = B.c_m2[ci];
    const int4 m3 = B.c_m3[ci];
    const float4 msa = B.b_mass[bai], msb = B.b_mass[bbi];
    const float m_a = msa.x, i_a = msa.y, m_b = msb.x, i_b = msb.y;
    const bool warm = (ws[WS_FLAGS] & B2GPU_WORLD_WARM_STARTING) != 0;
    const bool block = (ws[WS_FLAGS] & B2GPU_WORLD_BLOCK_SOLVE) != 0;
    const float dt_ratio = i2f(ws[WS_INV_DT0]) * sp.dt;
    const int point_count = m3.w;
    float ni[2] = {0.0f, 0.0f}, ti[2] = {0.0f, 0.0f};
    if (warm) {
      if (point_count > 0) { ni[0] = dt_ratio * m0.z; ti[0] = dt_ratio * m0.w; }
      if (point_count > 1) { ni[1] = dt_ratio * m1.z; ti[1] = dt_ratio * m1.w; }
    }
    // initialize_velocity_constraints
    const float4 pa = B.b_pos[bai], pb = B.b_pos[bbi];
    const float4 va = B.b_vel[bai], vb = B.b_vel[bbi];
    const float4 ra4 = B.b_rot[bai], rb4 = B.b_rot[bbi];
    const V2 c_a = v2(pa.x, pa.y), c_b = v2(pb.x, pb.y);
    const V2 v_a = v2(va.x, va.y), v_b = v2(vb.x, vb.y);
    const float w_a = va.z, w_b = vb.z;
    Xf xf_a, xf_b;
    xf_a.q.s = ra4.x; xf_a.q.c = ra4.y;
    xf_b.q.s = rb4.x; xf_b.q.c = rb4.y;
    xf_a.p = c_a - rot_mul(xf_a.q, v2(msa.z, msa.w));
    xf_b.p = c_b - rot_mul(xf_b.q, v2(msb.z, msb.w));
    Manifold m;
    m.pt[0] = v2(m0.x, m0.y); m.pt[1] = v2(m1.x, m1.y);
    m.ln = v2(m2.x, m2.y); m.lp = v2(m2.z, m2.w);
    m.type = m3.z; m.count = point_count;
    V2 normal, wp[2];
    world_manifold(normal, wp, m, xf_a, radius_a, xf_b, radius_b);
    V2 r_a[2], r_b[2];
    float nmass[2] = {0.0f, 0.0f}, tmass[2] = {0.0f, 0.0f}, bias[2] = {0.0f, 0.0f};
    r_a[0] = r_a[1] = r_b[0] = r_b[1] = v2(0.0f, 0.0f);
    for (int j = 0; j < point_count; ++j) {
      r_a[j] = wp[j] - c_a;
      r_b[j] = wp[j] - c_b;
      const float rn_a = cross(r_a[j], normal), rn_b = cross(r_b[j], normal);
      const float k_normal = m_a + m_b + i_a * rn_a * rn_a + i_b * rn_b * rn_b;
      nmass[j] = k_normal > 0.0f ? 1.0f / k_normal : 0.0f;
      const V2 tangent = cross_vs(normal, 1.0f);
      const float rt_a = cross(r_a[j], tangent), rt_b = cross(r_b[j], tangent);
      const float k_tangent = m_a + m_b + i_a * rt_a * rt_a + i_b * rt_b * rt_b;
      tmass[j] = k_tangent > 0.0f ? 1.0f / k_tangent : 0.0f;
      const float v_rel = dot(normal, v_b + cross_sv(w_b, r_b[j]) - v_a - cross_sv(w_a, r_a[j]));
      if (v_rel < -mat.z) bias[j] = -mat.y * v_rel;
    }
    int vc_points = point_count;
    float k11 = 0.0f, k12 = 0.0f, k22 = 0.0f, n11 = 0.0f, n12 = 0.0f, n22 = 0.0f;
    if (point_count == 2 && block) {
      const float rn1_a = cross(r_a[0], normal), rn1_b = cross(r_b[0], normal);
      const float rn2_a = cross(r_a[1], normal), rn2_b = cross(r_b[1], normal);
      const float q11 = m_a + m_b + i_a * rn1_a * rn1_a + i_b * rn1_b * rn1_b;
      const float q22 = m_a + m_b + i_a * rn2_a * rn2_a + i_b * rn2_b * rn2_b;
      const float q12 = m_a + m_b + i_a * rn1_a * rn2_a + i_b * rn1_b * rn2_b;
      const float k_max_condition_number = 1000.0f;
      if (q11 * q11 < k_max_condition_number * (q11 * q22 - q12 * q12)) {
        k11 = q11; k12 = q12; k22 = q22;
        // B2Mat22::get_inverse (src/b2_math.rs:261-274) of ex=(k11,k12) ey=(k12,k22)
        float det = k11 * k22 - k12 * k12;
        if (det != 0.0f) det = 1.0f / det;
        n11 = det * k22;
        n12 = -det * k12;
        n22 = det * k11;
      } else {
        vc_points = 1;
      }
    }
    B.vc[vc_at(B, x, k, 0)] = make_float4(r_a[0].x, r_a[0].y, r_b[0].x, r_b[0].y);
    B.vc[vc_at(B, x, k, 1)] = make_float4(r_a[1].x, r_a[1].y, r_b[1].x, r_b[1].y);
    B.vc[vc_at(B, x, k, 2)] = make_float4(normal.x, normal.y, mat.x, mat.w);
    B.vc[vc_at(B, x, k, 3)] = make_float4(nmass[0], tmass[0], bias[0], nmass[1]);
    B.vc[vc_at(B, x, k, 4)] = make_float4(tmass[1], bias[1], k11, k12);
    B.vc[vc_at(B, x, k, 5)] = make_float4(k22, n11, n12, n22);
    B.vc[vc_at(B, x, k, 6)] = make_float4(ni[0], ti[0], ni[1], ti[1]);
    B.vc[vc_at(B, x, k, 7)] = make_float4(m_a, i_a, m_b, i_b);
    B.vc[vc_at(B, x, k, 8)] = make_float4(i2f(ba), i2f(bb), i2f(vc_points | (point_count << 8) | (m3.z << 16)), i2f(c));
    B.pc[pc_at(B, x, k, 0)] = make_float4(m_a, i_a, m_b, i_b);
    B.pc[pc_at(B, x, k, 1)] = make_float4(msa.z, msa.w, msb.z, msb.w);
    B.pc[pc_at(B, x, k, 2)] = make_float4(m0.x, m0.y, m1.x, m1.y);
    B.pc[pc_at(B, x, k, 3)] = m2;
    B.pc[pc_at(B, x, k, 4)] = make_float4(radius_a, radius_b, i2f(ba), i2f(bb));
    B.pc[pc_at(B, x, k, 5)] = make_float4(i2f(point_count | (m3.z << 8)), i2f(B.c_isl[x.at(B.NC, k)]), 0.0f, 0.0f);
  }
};

// One velocity constraint in the reference's row order: friction rows of every point, then the
// normal rows (sequential for one point, 2x2 block LCP by enumeration for two).
// b2_contact_solver_private.rs:268-583.
struct VelState {
  V2 v_a, v_b;
  float w_a, w_b;
};
B2G_HD void solve_velocity_one(VelState& s, const float4 q0, const float4 q1, const float4 q2, const float4 q3,
                               const float4 q4, const float4 q5, float4& q6, const float4 q7, int vc_points, bool block) {
  const float m_a = q7.x, i_a = q7.y, m_b = q7.z, i_b = q7.w;
  const V2 normal = v2(q2.x, q2.y);
  const V2 tangent = cross_vs(normal, 1.0f);
  const float friction = q2.z, tangent_speed = q2.w;
  V2 v_a = s.v_a, v_b = s.v_b;
  float w_a = s.w_a, w_b = s.w_b;
  const V2 ra0 = v2(q0.x, q0.y), rb0 = v2(q0.z, q0.w), ra1 = v2(q1.x, q1.y), rb1 = v2(q1.z, q1.w);
  {  // friction, point 0
    const V2 dv = v_b + cross_sv(w_b, rb0) - v_a - cross_sv(w_a, ra0);
    const float vt = dot(dv, tangent) - tangent_speed;
    float lambda = q3.y * (-vt);
    const float max_friction = friction * q6.x;
    const float new_impulse = fclamp_sel(q6.y + lambda, -max_friction, max_friction);
    lambda = new_impulse - q6.y;
    q6.y = new_impulse;
    const V2 p = lambda * tangent;
    v_a = v_a - m_a * p;
    w_a -= i_a * cross(ra0, p);
    v_b = v_b + m_b * p;
    w_b += i_b * cross(rb0, p);
  }
  if (vc_points > 1) {  // friction, point 1
    const V2 dv = v_b + cross_sv(w_b, rb1) - v_a - cross_sv(w_a, ra1);
    const float vt = dot(dv, tangent) - tangent_speed;
    float lambda = q4.x * (-vt);
    const float max_friction = friction * q6.z;
    const float new_impulse = fclamp_sel(q6.w + lambda, -max_friction, max_friction);
    lambda = new_impulse - q6.w;
    q6.w = new_impulse;
    const V2 p = lambda * tangent;
    v_a = v_a - m_a * p;
    w_a -= i_a * cross(ra1, p);
    v_b = v_b + m_b * p;
    w_b += i_b * cross(rb1, p);
  }
  if (vc_points == 1 || !block) {
    {
      const V2 dv = v_b + cross_sv(w_b, rb0) - v_a - cross_sv(w_a, ra0);
      const float vn = dot(dv, normal);
      float lambda = -q3.x * (vn - q3.z);
      const float new_impulse = fmax_sel(q6.x + lambda, 0.0f);
      lambda = new_impulse - q6.x;
      q6.x = new_impulse;
      const V2 p = lambda * normal;
      v_a = v_a - m_a * p;
      w_a -= i_a * cross(ra0, p);
      v_b = v_b + m_b * p;
      w_b += i_b * cross(rb0, p);
    }
    if (vc_points > 1) {
      const V2 dv = v_b + cross_sv(w_b, rb1) - v_a - cross_sv(w_a, ra1);
      const float vn = dot(dv, normal);
      float lambda = -q3.w * (vn - q4.y);
      const float new_impulse = fmax_sel(q6.z + lambda, 0.0f);
      lambda = new_impulse - q6.z;
      q6.z = new_impulse;
      const V2 p = lambda * normal;
      v_a = v_a - m_a * p;
      w_a -= i_a * cross(ra1, p);
      v_b = v_b + m_b * p;
      w_b += i_b * cross(rb1, p);
    }
  } else {
    // block solver (:352-576): K = [k11 k12; k12 k22], normal_mass = K^-1.  The reference tries the four
    // LCP cases in order and takes the first that satisfies its conditions; the candidates are evaluated
    // branch-free here and selected in the same priority order (identical arithmetic per case).
    const float k11 = q4.z, k12 = q4.w, k22 = q5.x, n11 = q5.y, n12 = q5.z, n22 = q5.w;
    const V2 a = v2(q6.x, q6.z);
    const V2 dv1 = v_b + cross_sv(w_b, rb0) - v_a - cross_sv(w_a, ra0);
    const V2 dv2 = v_b + cross_sv(w_b, rb1) - v_a - cross_sv(w_a, ra1);
    const float vn1 = dot(dv1, normal);
    const float vn2 = dot(dv2, normal);
    V2 b = v2(vn1 - q3.z, vn2 - q4.y);
    b = b - v2(k11 * a.x + k12 * a.y, k12 * a.x + k22 * a.y);
    const V2 x1 = -v2(n11 * b.x + n12 * b.y, n12 * b.x + n22 * b.y);  // case 1: both points active
    const bool ok1 = x1.x >= 0.0f && x1.y >= 0.0f;
    const float x2 = -q3.x * b.x;                                       // case 2: point 1 active
    const bool ok2 = x2 >= 0.0f && (k12 * x2 + b.y) >= 0.0f;
    const float x3 = -q3.w * b.y;                                       // case 3: point 2 active
    const bool ok3 = x3 >= 0.0f && (k12 * x3 + b.x) >= 0.0f;
    const bool ok4 = b.x >= 0.0f && b.y >= 0.0f;                        // case 4: none active
    V2 xs;
    xs.x = ok1 ? x1.x : (ok2 ? x2 : 0.0f);
    xs.y = ok1 ? x1.y : (ok2 ? 0.0f : (ok3 ? x3 : 0.0f));
    {  // "no solution, give up" (:570) leaves everything untouched: select, do not branch
      const int any = sel_mask(ok1 || ok2 || ok3 || ok4);
      const V2 d = xs - a;
      const V2 p1 = d.x * normal, p2 = d.y * normal;
      const V2 nv_a = v_a - m_a * (p1 + p2);
      const float nw_a = w_a - i_a * (cross(ra0, p1) + cross(ra1, p2));
      const V2 nv_b = v_b + m_b * (p1 + p2);
      const float nw_b = w_b + i_b * (cross(rb0, p1) + cross(rb1, p2));
      v_a = v2(msel(any, nv_a.x, v_a.x), msel(any, nv_a.y, v_a.y));
      w_a = msel(any, nw_a, w_a);
      v_b = v2(msel(any, nv_b.x, v_b.x), msel(any, nv_b.y, v_b.y));
      w_b = msel(any, nw_b, w_b);
      q6.x = msel(any, xs.x, q6.x);
      q6.z = msel(any, xs.y, q6.z);
    }
  }
  s.v_a = v_a; s.v_b = v_b; s.w_a = w_a; s.w_b = w_b;
}

// warm_start of one constraint (:228-266)
B2G_HD void warm_start_one(VelState& s, const float4 q0, const float4 q1, const float4 q2, const float4 q6, const float4 q7,
                           int vc_points) {
  const float m_a = q7.x, i_a = q7.y, m_b = q7.z, i_b = q7.w;
  const V2 normal = v2(q2.x, q2.y);
  const V2 tangent = cross_vs(normal, 1.0f);
  {
    const V2 p = q6.x * normal + q6.y * tangent;
    s.w_a -= i_a * cross(v2(q0.x, q0.y), p);
    s.v_a = s.v_a - m_a * p;
    s.w_b += i_b * cross(v2(q0.z, q0.w), p);
    s.v_b = s.v_b + m_b * p;
  }
  if (vc_points > 1) {
    const V2 p = q6.z * normal + q6.w * tangent;
    s.w_a -= i_a * cross(v2(q1.x, q1.y), p);
    s.v_a = s.v_a - m_a * p;
    s.w_b += i_b * cross(v2(q1.z, q1.w), p);
    s.v_b = s.v_b + m_b * p;
  }
}

// Ordered stage B, generic global-memory form: warm start + velocity iterations, one thread per island
// (islands share no dynamic body, so they are independent; inside an island the reference order is kept).
struct VelocityK {
  Batch B;
  StepParams sp;
  // one pass over the island's contact constraints: the warm start (:228-266) or one velocity iteration (:268-583)
  B2G_HD void contact_sweep(const WIdx& x, const int4 rg, bool warm_pass, bool block) const {
    for (int k = rg.z; k < rg.w; ++k) {
      const float4 q8 = B.vc[vc_at(B, x, k, 8)];
      const int ba = f2i(q8.x), bb = f2i(q8.y), vc_points = f2i(q8.z) & 0xff;
      if (vc_points == 0) continue;
      const int bai = x.at(B.NB, ba), bbi = x.at(B.NB, bb);
      const float4 va = B.b_vel[bai], vb = B.b_vel[bbi];
      VelState s;
      s.v_a = v2(va.x, va.y); s.w_a = va.z;
      s.v_b = v2(vb.x, vb.y); s.w_b = vb.z;
      const float4 q0 = B.vc[vc_at(B, x, k, 0)], q1 = B.vc[vc_at(B, x, k, 1)], q2 = B.vc[vc_at(B, x, k, 2)];
      float4 q6 = B.vc[vc_at(B, x, k, 6)];
      const float4 q7 = B.vc[vc_at(B, x, k, 7)];
      if (warm_pass) {
        warm_start_one(s, q0, q1, q2, q6, q7, vc_points);
      } else {
        const float4 q3 = B.vc[vc_at(B, x, k, 3)], q4 = B.vc[vc_at(B, x, k, 4)], q5 = B.vc[vc_at(B, x, k, 5)];
        solve_velocity_one(s, q0, q1, q2, q3, q4, q5, q6, q7, vc_points, block);
        B.vc[vc_at(B, x, k, 6)] = q6;
      }
      // a static / kinematic body may sit in several islands: its velocity never changes, leave it alone
      if (q7.x != 0.0f || q7.y != 0.0f) B.b_vel[bai] = make_float4(s.v_a.x, s.v_a.y, s.w_a, 0.0f);
      if (q7.z != 0.0f || q7.w != 0.0f) B.b_vel[bbi] = make_float4(s.v_b.x, s.v_b.y, s.w_b, 0.0f);
    }
  }
  B2G_HD void operator()(int tid) const {
    int w, isl;
    if (!flat_decode(B, tid, B.NB, w, isl)) return;
    WIdx x = widx(B, w);
    Ws ws = ws_of(B, x);
    if (isl >= ws[WS_ISL_COUNT]) return;
    const int4 rg = B.isl_range[x.at(B.NB, isl)];
    int2 jr = make_int2(0, 0);
    if (B.NJ > 0) jr = B.isl_jrange[x.at(B.NB, isl)];
    if (rg.z == rg.w && jr.x == jr.y) return;
    const bool warm = (ws[WS_FLAGS] & B2GPU_WORLD_WARM_STARTING) != 0;
    const bool block = (ws[WS_FLAGS] & B2GPU_WORLD_BLOCK_SOLVE) != 0;
    // order of B2island::solve (b2_island_private.rs:193-215): contact warm start, every joint's
    // init_velocity_constraints (which applies the joint's own warm start), then per iteration joints before contacts
    if (warm) contact_sweep(x, rg, true, block);
    const float dt_ratio = i2f(ws[WS_INV_DT0]) * sp.dt;
    const BodyStateGlobal st = {B, x};
    for (int q = jr.x; q < jr.y; ++q) joint_init_velocity(B, x, st, B.isl_joint[x.at(B.NJ, q)], warm, dt_ratio, sp.dt);
    for (int it = 0; it < sp.velocity_iterations; ++it) {
      for (int q = jr.x; q < jr.y; ++q) joint_solve_velocity(B, x, st, B.isl_joint[x.at(B.NJ, q)], sp.dt, sp.inv_dt);
      contact_sweep(x, rg, false, block);
    }
  }
};

// store_impulses (:585-598) + position integration (b2_island_private.rs:222-252): flat.
struct PostVelocityK {
  Batch B;
  StepParams sp;
  B2G_HD void operator()(int tid) const {
    int w, k;
    if (!flat_decode(B, tid, B.NIB, w, k)) return;
    WIdx x = widx(B, w);
    Ws ws = ws_of(B, x);
    if (k < ws[WS_ISL_COUNT]) {
      const int4 rg = B.isl_range[x.at(B.NB, k)];
      bool no_joints = true;
      if (B.NJ > 0) { const int2 jr = B.isl_jrange[x.at(B.NB, k)]; no_joints = jr.x == jr.y; }
      B.isl_flags[x.at(B.NB, k)] = (rg.z == rg.w && no_joints && sp.position_iterations > 0) ? 1 : 0;
    }
    if (k < ws[WS_ISL_CONTACTS]) {
      const float4 q8 = B.vc[vc_at(B, x, k, 8)];
      const int vc_points = f2i(q8.z) & 0xff, c = f2i(q8.w);
      const float4 q6 = B.vc[vc_at(B, x, k, 6)];
      const int ci = x.at(B.NC, c);
      if (vc_points > 0) { B.c_m0[ci].z = q6.x; B.c_m0[ci].w = q6.y; }
      if (vc_points > 1) { B.c_m1[ci].z = q6.z; B.c_m1[ci].w = q6.w; }
    }
    if (k < ws[WS_ISL_BODIES]) {
      const int b = B.isl_body[x.at(B.NIB, k)];
      const int bi = x.at(B.NB, b);
      if (body_type(B.b_flags[bi]) == B2GPU_STATIC_BODY) return;  // immovable; may sit in several islands
      const float h = sp.dt;
      float4 pos = B.b_pos[bi];
      float4 vel = B.b_vel[bi];
      V2 v = v2(vel.x, vel.y);
      float wv = vel.z;
      const V2 translation = h * v;
      if (dot(translation, translation) > B2G_MAX_TRANSLATION_SQUARED) {
        const float ratio = B2G_MAX_TRANSLATION / length(translation);
        v = ratio * v;
      }
      const float rotation = h * wv;
      if (rotation * rotation > B2G_MAX_ROTATION_SQUARED) {
        const float ratio = B2G_MAX_ROTATION / fabsf(rotation);
        wv *= ratio;
      }
      pos.x += h * v.x;
      pos.y += h * v.y;
      pos.z += h * wv;
      B.b_pos[bi] = pos;
      B.b_vel[bi] = make_float4(v.x, v.y, wv, 0.0f);
      Rot q = rot_from_angle(pos.z);
      float4 r = B.b_rot[bi];
      r.x = q.s;
      r.y = q.c;
      B.b_rot[bi] = r;
    }
  }
};

// One position constraint (b2_contact_solver_private.rs:600-730).  The reference rebuilds both
// transforms (sin/cos) for every manifold point from the running angles; sin/cos are pure
// functions of the angle, so they are cached per body (b_rot) and recomputed only when a
// correction actually changed the angle — bit-identical, far fewer evaluations.
struct PosState {
  V2 c_a, c_b;
  float a_a, a_b;
  Rot q_a, q_b;
};
B2G_HD float solve_position_one(PosState& s, const float4 q7, const float4 q9, const float4 lps, const float4 m2,
                                int type, int pc_points, float radius_a, float radius_b, float min_separation) {
  const float m_a = q7.x, i_a = q7.y, m_b = q7.z, i_b = q7.w;
  const V2 lc_a = v2(q9.x, q9.y), lc_b = v2(q9.z, q9.w);
  const float4 m0 = make_float4(lps.x, lps.y, 0.0f, 0.0f), m1 = make_float4(lps.z, lps.w, 0.0f, 0.0f);
  for (int j = 0; j < pc_points; ++j) {
    Xf xf_a, xf_b;
    xf_a.q = s.q_a;
    xf_b.q = s.q_b;
    xf_a.p = s.c_a - rot_mul(xf_a.q, lc_a);
    xf_b.p = s.c_b - rot_mul(xf_b.q, lc_b);
    V2 normal, point;
    float separation;
    if (type == B2GPU_MANIFOLD_CIRCLES) {
      const V2 point_a = xf_mul(xf_a, v2(m2.z, m2.w));
      const V2 point_b = xf_mul(xf_b, v2(m0.x, m0.y));
      normal = point_b - point_a;
      normalize(normal);
      point = 0.5f * (point_a + point_b);
      separation = dot(point_b - point_a, normal) - radius_a - radius_b;
    } else if (type == B2GPU_MANIFOLD_FACE_A) {
      normal = rot_mul(xf_a.q, v2(m2.x, m2.y));
      const V2 plane_point = xf_mul(xf_a, v2(m2.z, m2.w));
      const V2 clip_point = xf_mul(xf_b, j == 0 ? v2(m0.x, m0.y) : v2(m1.x, m1.y));
      separation = dot(clip_point - plane_point, normal) - radius_a - radius_b;
      point = clip_point;
    } else {
      normal = rot_mul(xf_b.q, v2(m2.x, m2.y));
      const V2 plane_point = xf_mul(xf_b, v2(m2.z, m2.w));
      const V2 clip_point = xf_mul(xf_a, j == 0 ? v2(m0.x, m0.y) : v2(m1.x, m1.y));
      separation = dot(clip_point - plane_point, normal) - radius_a - radius_b;
      point = clip_point;
      normal = -normal;
    }
    const V2 r_a = point - s.c_a, r_b = point - s.c_b;
    min_separation = fmin_sel(min_separation, separation);
    const float cc = fclamp_sel(B2G_BAUMGARTE * (separation + B2G_LINEAR_SLOP), -B2G_MAX_LINEAR_CORRECTION, 0.0f);
    const float rn_a = cross(r_a, normal), rn_b = cross(r_b, normal);
    const float kk = m_a + m_b + i_a * rn_a * rn_a + i_b * rn_b * rn_b;
    const float impulse = kk > 0.0f ? -cc / kk : 0.0f;
    const V2 p = impulse * normal;
    s.c_a = s.c_a - m_a * p;
    const float na = s.a_a - i_a * cross(r_a, p);
    s.c_b = s.c_b + m_b * p;
    const float nb = s.a_b + i_b * cross(r_b, p);
    if (f2u(na) != f2u(s.a_a)) { s.a_a = na; s.q_a = rot_from_angle(na); }
    if (f2u(nb) != f2u(s.a_b)) { s.a_b = nb; s.q_b = rot_from_angle(nb); }
  }
  return min_separation;
}

// Straight-line form of solve_position_one for the common case, a face manifold with two points.  FACE_A and
// FACE_B differ only in which body carries the reference face, so the (rotation, origin) pair of the
// reference body and of the incident body are selected (sel_mask: a select that cannot become a branch) and
// the arithmetic is shared, operation for operation as in solve_position_one.  The rotations are refreshed
// unconditionally with the branch-free sincos_mid: b_rot always holds rot_from_angle of the running angle, so
// an unchanged angle reproduces the same sine and cosine bits.  An angle outside sincos_mid's domain sets
// `wide`: the result is then invalid and the caller redoes the constraint with solve_position_one.
B2G_HD float solve_position_face2(PosState& s, const float4 q7, const float4 q9, const float4 lps, const float4 m2,
                                  bool face_a, float radius_a, float radius_b, float min_separation, bool& wide) {
  const float m_a = q7.x, i_a = q7.y, m_b = q7.z, i_b = q7.w;
  const V2 lc_a = v2(q9.x, q9.y), lc_b = v2(q9.z, q9.w);
  const int fa = sel_mask(face_a);
  wide = false;
  const int sign = fa ? 0 : (int)0x80000000u;  // FACE_B reports the negated normal
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < 2; ++j) {
    const V2 mj = j == 0 ? v2(lps.x, lps.y) : v2(lps.z, lps.w);
    Xf xf_a, xf_b;
    xf_a.q = s.q_a;
    xf_b.q = s.q_b;
    xf_a.p = s.c_a - rot_mul(xf_a.q, lc_a);
    xf_b.p = s.c_b - rot_mul(xf_b.q, lc_b);
    Xf xr, xi;  // reference-face body, incident body
    xr.q.s = msel(fa, xf_a.q.s, xf_b.q.s); xr.q.c = msel(fa, xf_a.q.c, xf_b.q.c);
    xr.p.x = msel(fa, xf_a.p.x, xf_b.p.x); xr.p.y = msel(fa, xf_a.p.y, xf_b.p.y);
    xi.q.s = msel(fa, xf_b.q.s, xf_a.q.s); xi.q.c = msel(fa, xf_b.q.c, xf_a.q.c);
    xi.p.x = msel(fa, xf_b.p.x, xf_a.p.x); xi.p.y = msel(fa, xf_b.p.y, xf_a.p.y);
    V2 normal = rot_mul(xr.q, v2(m2.x, m2.y));
    const V2 plane_point = xf_mul(xr, v2(m2.z, m2.w));
    const V2 clip_point = xf_mul(xi, mj);
    const float separation = dot(clip_point - plane_point, normal) - radius_a - radius_b;
    const V2 point = clip_point;
    normal = v2(i2f(f2i(normal.x) ^ sign), i2f(f2i(normal.y) ^ sign));
    const V2 r_a = point - s.c_a, r_b = point - s.c_b;
    min_separation = fmin_sel(min_separation, separation);
    const float cc = fclamp_sel(B2G_BAUMGARTE * (separation + B2G_LINEAR_SLOP), -B2G_MAX_LINEAR_CORRECTION, 0.0f);
    const float rn_a = cross(r_a, normal), rn_b = cross(r_b, normal);
    const float kk = m_a + m_b + i_a * rn_a * rn_a + i_b * rn_b * rn_b;
    const float impulse = kk > 0.0f ? -cc / kk : 0.0f;
    const V2 p = impulse * normal;
    s.c_a = s.c_a - m_a * p;
    s.a_a = s.a_a - i_a * cross(r_a, p);
    s.c_b = s.c_b + m_b * p;
    s.a_b = s.a_b + i_b * cross(r_b, p);
    sincos_mid(s.a_a, &s.q_a.s, &s.q_a.c);
    sincos_mid(s.a_b, &s.q_b.s, &s.q_b.c);
    wide = wide || !(sincos_mid_domain(s.a_a) && sincos_mid_domain(s.a_b));
  }
  return min_separation;
}

// Ordered stage C, generic form: position iterations of one island per thread, with the reference's early
// exit (b2_island_private.rs:257-274).  Islands without contacts were marked solved by PostVelocityK.
struct PositionK {
  Batch B;
  StepParams sp;
  B2G_HD void operator()(int tid) const {
    int w, isl;
    if (!flat_decode(B, tid, B.NB, w, isl)) return;
    WIdx x = widx(B, w);
    Ws ws = ws_of(B, x);
    if (isl >= ws[WS_ISL_COUNT]) return;
    const int4 rg = B.isl_range[x.at(B.NB, isl)];
    int2 jr = make_int2(0, 0);
    if (B.NJ > 0) jr = B.isl_jrange[x.at(B.NB, isl)];
    if (rg.z == rg.w && jr.x == jr.y) return;
    for (int it = 0; it < sp.position_iterations; ++it) {
      float min_separation = 0.0f;
      for (int k = rg.z; k < rg.w; ++k) {
        const float4 p5 = B.pc[pc_at(B, x, k, 5)];
        const float4 p4 = B.pc[pc_at(B, x, k, 4)];
        const float4 p0 = B.pc[pc_at(B, x, k, 0)];
        const int ba = f2i(p4.z), bb = f2i(p4.w), packed = f2i(p5.x);
        const int bai = x.at(B.NB, ba), bbi = x.at(B.NB, bb);
        float4 pa = B.b_pos[bai], pb = B.b_pos[bbi];
        float4 ra = B.b_rot[bai], rb = B.b_rot[bbi];
        PosState s;
        s.c_a = v2(pa.x, pa.y); s.a_a = pa.z; s.q_a.s = ra.x; s.q_a.c = ra.y;
        s.c_b = v2(pb.x, pb.y); s.a_b = pb.z; s.q_b.s = rb.x; s.q_b.c = rb.y;
        min_separation = solve_position_one(s, p0, B.pc[pc_at(B, x, k, 1)], B.pc[pc_at(B, x, k, 2)],
                                            B.pc[pc_at(B, x, k, 3)], (packed >> 8) & 0xff, packed & 0xff, p4.x, p4.y, min_separation);
        if (p0.x != 0.0f || p0.y != 0.0f) {  // immovable bodies are shared between islands: never written
          pa.x = s.c_a.x; pa.y = s.c_a.y; pa.z = s.a_a;
          ra.x = s.q_a.s; ra.y = s.q_a.c;
          B.b_pos[bai] = pa; B.b_rot[bai] = ra;
        }
        if (p0.z != 0.0f || p0.w != 0.0f) {
          pb.x = s.c_b.x; pb.y = s.c_b.y; pb.z = s.a_b;
          rb.x = s.q_b.s; rb.y = s.q_b.c;
          B.b_pos[bbi] = pb; B.b_rot[bbi] = rb;
        }
      }
      bool joints_okay = true;  // b2_island_private.rs:262-266: every joint is solved, none short-circuits
      for (int q = jr.x; q < jr.y; ++q) {
        const bool joint_okay = joint_solve_position(B, x, BodyStateGlobal{B, x}, B.isl_joint[x.at(B.NJ, q)]);
        joints_okay = joints_okay && joint_okay;
      }
      if (min_separation >= -3.0f * B2G_LINEAR_SLOP && joints_okay) {
        B.isl_flags[x.at(B.NB, isl)] |= 1;
        break;
      }
    }
  }
};

// Copy-back + synchronize_transform (b2_island_private.rs:277-285, src/b2_body.rs:974-977) and the
// per-body part of the sleep bookkeeping (:291-318): flat over island body slots.
struct FinalizeK {
  Batch B;
  StepParams sp;
  B2G_HD void operator()(int tid) const {
    int w, k;
    if (!flat_decode(B, tid, B.NIB, w, k)) return;
    WIdx x = widx(B, w);
    Ws ws = ws_of(B, x);
    if (k >= ws[WS_ISL_BODIES]) return;
    const int b = B.isl_body[x.at(B.NIB, k)];
    const int bi = x.at(B.NB, b);
    const int bf = B.b_flags[bi];
    if (body_type(bf) == B2GPU_STATIC_BODY) return;
    float4 pos = B.b_pos[bi];
    const float4 r = B.b_rot[bi], ms = B.b_mass[bi];
    Rot q;
    q.s = r.x;
    q.c = r.y;
    const V2 p = v2(pos.x, pos.y) - rot_mul(q, v2(ms.z, ms.w));
    B.b_xf[bi] = make_float4(p.x, p.y, q.s, q.c);
    if (ws[WS_FLAGS] & B2GPU_WORLD_ALLOW_SLEEP) {
      const float4 vel = B.b_vel[bi];
      const float lin_tol_sqr = B2G_LINEAR_SLEEP_TOLERANCE * B2G_LINEAR_SLEEP_TOLERANCE;
      const float ang_tol_sqr = B2G_ANGULAR_SLEEP_TOLERANCE * B2G_ANGULAR_SLEEP_TOLERANCE;
      if (!(bf & B2GPU_BODY_AUTO_SLEEP) || vel.z * vel.z > ang_tol_sqr || dot(v2(vel.x, vel.y), v2(vel.x, vel.y)) > lin_tol_sqr)
        pos.w = 0.0f;
      else
        pos.w += sp.dt;
      B.b_pos[bi].w = pos.w;
    }
  }
};

// Island-wide sleep decision (:319-327): one thread per island.
struct SleepK {
  Batch B;
  B2G_HD void operator()(int tid) const {
    int w, isl;
    if (!flat_decode(B, tid, B.NB, w, isl)) return;
    WIdx x = widx(B, w);
    Ws ws = ws_of(B, x);
    if (!(ws[WS_FLAGS] & B2GPU_WORLD_ALLOW_SLEEP)) return;
    if (isl >= ws[WS_ISL_COUNT]) return;
    const int4 rg = B.isl_range[x.at(B.NB, isl)];
    float min_sleep_time = B2G_MAX_FLOAT;
    for (int k = rg.x; k < rg.y; ++k) {
      const int bi = x.at(B.NB, B.isl_body[x.at(B.NIB, k)]);
      if (body_type(B.b_flags[bi]) == B2GPU_STATIC_BODY) continue;
      min_sleep_time = fmin_sel(min_sleep_time, B.b_pos[bi].w);
    }
    if (min_sleep_time >= B2G_TIME_TO_SLEEP && (B.isl_flags[x.at(B.NB, isl)] & 1)) {
      for (int k = rg.x; k < rg.y; ++k) {  // set_awake(false), src/b2_body.rs:783-801
        const int bi = x.at(B.NB, B.isl_body[x.at(B.NIB, k)]);
        const int bf = B.b_flags[bi];
        if (body_type(bf) == B2GPU_STATIC_BODY) continue;
        B.b_flags[bi] = bf & ~B2GPU_BODY_AWAKE;
        B.b_pos[bi].w = 0.0f;
        B.b_vel[bi] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        float4 fo = B.b_force[bi];
        fo.x = 0.0f; fo.y = 0.0f; fo.z = 0.0f;
        B.b_force[bi] = fo;
      }
      ws[WS_TOPO_DIRTY] = 1;
    }
  }
};

// synchronize_fixtures (b2_body.rs(private):455-475, b2_fixture.rs(private):147-173) and the
// keep-or-reinsert test of move_proxy (b2_dynamic_tree.rs(private):109-168): flat over proxies.
struct SyncFixturesK {
  Batch B;
  B2G_HD void operator()(int tid) const {
    int w, p;
    if (!flat_decode(B, tid, B.NP, w, p)) return;
    WIdx x = widx(B, w);
    Ws ws = ws_of(B, x);
    const int4 ps = B.proxy_s[p];
    const int bi = x.at(B.NB, ps.w);
    const int bf = B.b_flags[bi];
    if (!(bf & B2GPU_BODY_ISLAND) || body_type(bf) == B2GPU_STATIC_BODY) return;
    const float4 t = B.b_xf[bi];
    Xf xf2;
    xf2.p = v2(t.x, t.y);
    xf2.q.s = t.z;
    xf2.q.c = t.w;
    Xf xf1 = xf2;
    if (bf & B2GPU_BODY_AWAKE) {
      const float4 p0 = B.b_pos0[bi], r = B.b_rot[bi], ms = B.b_mass[bi];
      xf1.q.s = r.z;
      xf1.q.c = r.w;
      xf1.p = v2(p0.x, p0.y) - rot_mul(xf1.q, v2(ms.z, ms.w));
    }
    const b2gpu_shape_rec* sh = &B.shapes[B.fixtures[ps.x].shape_first + ps.y];
    const Box a1 = shape_aabb(sh, xf1), a2 = shape_aabb(sh, xf2);
    const Box tight = box_union(a1, a2);
    const V2 displacement = box_center(a2) - box_center(a1);
    B.p_aabb[x.at(B.NP, p)] = make_float4(tight.lo.x, tight.lo.y, tight.hi.x, tight.hi.y);
    // move_proxy
    const V2 r = v2(B2G_AABB_EXTENSION, B2G_AABB_EXTENSION);
    Box fat;
    fat.lo = tight.lo - r;
    fat.hi = tight.hi + r;
    const V2 d = B2G_AABB_MULTIPLIER * displacement;
    if (d.x < 0.0f) fat.lo.x += d.x; else fat.hi.x += d.x;
    if (d.y < 0.0f) fat.lo.y += d.y; else fat.hi.y += d.y;
    const Box tree_box = load_box(B.n_aabb, x.at(B.NN, ps.z));
    if (box_contains(tree_box, tight)) {
      Box huge;
      huge.lo = fat.lo - 4.0f * r;
      huge.hi = fat.hi + 4.0f * r;
      if (box_contains(huge, tree_box)) return;
    }
    B.p_fat[x.at(B.NP, p)] = make_float4(fat.lo.x, fat.lo.y, fat.hi.x, fat.hi.y);
    // mark the proxy in the world's move bitmap at its rank in synchronize order, so the ordered stage
    // visits exactly the moved proxies, in the reference's order, without scanning all of them
    const int rank = B.sync_rank[p];
    B2G_ATOMIC_OR(&B.p_move[x.at(B.NMW, rank >> 5)], 1 << (rank & 31));
    counter_inc(&ws[WS_EV_MOVED]);
  }
};

B2G_HD Tree tree_of(const Batch& B, const WIdx& x, const Ws& ws) {
  Tree t;
  t.aabb = B.n_aabb + x.at(B.NN, 0);
  t.link = B.n_link + x.at(B.NN, 0);
  t.moved = B.n_moved + x.at(B.NN, 0);
  t.stride = x.LB;
  t.ws = ws.p;
  t.ws_stride = ws.stride;
  t.phys_cap = B.NN;
  t.status = &ws[WS_STATUS];
  return t;
}

B2G_HD bool type_pair_primary(int ta, int tb) {  // b2_contact_registers.rs:67-103
  return (ta == B2GPU_SHAPE_CIRCLE && tb == B2GPU_SHAPE_CIRCLE) || (ta == B2GPU_SHAPE_POLYGON && tb == B2GPU_SHAPE_CIRCLE) ||
         (ta == B2GPU_SHAPE_POLYGON && tb == B2GPU_SHAPE_POLYGON) || (ta == B2GPU_SHAPE_EDGE && tb == B2GPU_SHAPE_CIRCLE) ||
         (ta == B2GPU_SHAPE_EDGE && tb == B2GPU_SHAPE_POLYGON) || (ta == B2GPU_SHAPE_CHAIN && tb == B2GPU_SHAPE_CIRCLE) ||
         (ta == B2GPU_SHAPE_CHAIN && tb == B2GPU_SHAPE_POLYGON);
}

// B2contactManager::add_pair (b2_contact_manager.rs(private):178-302)
B2G_HD void add_pair(const Batch& B, const WIdx& x, const Ws& ws, int* b_chead, int2* c_next, int proxy_a, int proxy_b) {
  const int4 pa = B.proxy_s[proxy_a], pb = B.proxy_s[proxy_b];
  int fixture_a = pa.x, fixture_b = pb.x, index_a = pa.y, index_b = pb.y;
  const int body_a = pa.w, body_b = pb.w;
  if (body_a == body_b) return;
  for (int e = b_chead[x.at(B.NB, body_b)]; e != -1;) {
    const int c = e >> 1, side = e & 1;
    const int ci = x.at(B.NC, c);
    const int2 nx = c_next[ci];
    e = side ? nx.y : nx.x;
    const int4 fx = B.c_fix[ci];
    if (fx.x == fixture_a && fx.y == fixture_b && fx.z == index_a && fx.w == index_b) return;
    if (fx.x == fixture_b && fx.y == fixture_a && fx.z == index_b && fx.w == index_a) return;
  }
  if (!body_should_collide(B.b_flags[x.at(B.NB, body_b)], B.b_flags[x.at(B.NB, body_a)])) return;
  if (joints_prevent_collision(B, body_b, body_a)) return;
  const b2gpu_fixture_rec* fa = &B.fixtures[fixture_a];
  const b2gpu_fixture_rec* fb = &B.fixtures[fixture_b];
  if (!filter_should_collide(*fa, *fb)) return;
  if (!type_pair_primary(fa->shape_type, fb->shape_type)) {
    if (!type_pair_primary(fb->shape_type, fa->shape_type)) {
      ws[WS_STATUS] = B2GPU_E_UNSUPPORTED;  // the reference panics (unwrap on an unregistered pair)
      return;
    }
    int t = fixture_a; fixture_a = fixture_b; fixture_b = t;
    t = index_a; index_a = index_b; index_b = t;
    const b2gpu_fixture_rec* tf = fa; fa = fb; fb = tf;
  }
  const int c = ws[WS_CONTACT_COUNT];
  if (c >= B.NC) { ws[WS_STATUS] = B2GPU_E_CAPACITY; return; }
  const int ci = x.at(B.NC, c);
  B.c_fix[ci] = make_int4(fixture_a, fixture_b, index_a, index_b);
  B.c_flags[ci] = B2GPU_CONTACT_ENABLED;
  B.c_mat[ci] = make_float4(sqrtf(fa->friction * fb->friction),
                            fa->restitution > fb->restitution ? fa->restitution : fb->restitution,
                            fa->restitution_threshold < fb->restitution_threshold ? fa->restitution_threshold
                                                                                  : fb->restitution_threshold,
                            0.0f);
  B.c_m0[ci] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  B.c_m1[ci] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  B.c_m2[ci] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  B.c_m3[ci] = make_int4(0, 0, 0, 0);
  link_contact(B, x, b_chead, c_next, c);
  ws[WS_CONTACT_COUNT] = c + 1;
  ws[WS_ST_CREATED] += 1;
}

// B2broadPhase::update_pairs (src/b2_broad_phase.rs:200-249) with the contact manager as callback.
B2G_HD void update_pairs(const Batch& B, const WIdx& x, const Ws& ws, int* b_chead, int2* c_next) {
  const int mc = ws[WS_MOVE_COUNT];
  if (mc == 0) return;
  Tree t = tree_of(B, x, ws);
  int np = 0;
  for (int i = 0; i < mc; ++i) {
    const int q = B.move_buf[x.at(B.NMOVE, i)];
    if (q == -1) continue;
    const Box fat = t.A(q);
    int stack[B2G_QUERY_STACK];
    int sp_ = 0;
    stack[sp_++] = t.root();
    while (sp_ > 0) {
      const int id = stack[--sp_];
      if (id == -1) continue;
      if (!box_overlap(t.A(id), fat)) continue;
      const int4 l = t.L(id);
      if (l.y == -1) {
        // b2_broad_phase_query_callback (b2_broad_phase.rs(private):86-111).  The reference appends the
        // pair to m_pair_buffer and calls add_pair for the whole buffer afterwards; add_pair changes
        // neither the tree nor the moved flags, so calling it here, in the same order, is equivalent and
        // needs no pair buffer (whose size is unbounded: every moved proxy re-reports all its overlaps).
        if (id == q) continue;
        if (t.moved[id * t.stride] && id > q) continue;
        ++np;
        add_pair(B, x, ws, b_chead, c_next, B.node_proxy[imin(id, q)], B.node_proxy[imax(id, q)]);
      } else {
        if (sp_ + 2 > B2G_QUERY_STACK) { ws[WS_STATUS] = B2GPU_E_CAPACITY; continue; }
        stack[sp_++] = l.y;
        stack[sp_++] = l.z;
      }
    }
  }
  for (int i = 0; i < mc; ++i) {
    const int q = B.move_buf[x.at(B.NMOVE, i)];
    if (q == -1) continue;
    t.moved[q * t.stride] = 0;
  }
  ws[WS_ST_MOVED] += mc;
  ws[WS_ST_PAIRS] += np;
  ws[WS_MOVE_COUNT] = 0;
}

// Ordered stage D (one thread per world): tree re-insertion in synchronize order, pair finding,
// contact creation, end-of-step scalars (b2_world.rs(private):948-950).
struct TreePairsK {
  Batch B;
  int* b_chead;
  int2* c_next;
  StepParams sp;
  int pre_step;  // 1: the find_new_contacts call at the top of step (m_new_contacts), no tree moves
  B2G_HD void operator()(int tid) const {
    const int w = tid + (B.wb_first << B.lb_shift);
    if (w >= B.n_worlds) return;
    WIdx x = widx(B, w);
    Ws ws = ws_of(B, x);
    if (pre_step) {
      for (int s = WS_ST_CONTACTS; s <= WS_ST_LEVELS; ++s) ws[s] = 0;  // stats of this step start here
      if (ws[WS_FLAGS] & B2GPU_WORLD_NEW_CONTACTS) {
        update_pairs(B, x, ws, b_chead, c_next);
        ws[WS_FLAGS] &= ~B2GPU_WORLD_NEW_CONTACTS;
      }
      return;
    }
    if (sp.dt > 0.0f) {
      if (ws[WS_EV_MOVED]) {
        Tree t = tree_of(B, x, ws);
        int mc = ws[WS_MOVE_COUNT];
        for (int wi = 0; wi < B.NMW; ++wi) {
          unsigned bits = (unsigned)B.p_move[x.at(B.NMW, wi)];
          if (!bits) continue;
          B.p_move[x.at(B.NMW, wi)] = 0;
          while (bits) {
            const int bit = lowest_bit(bits);
            bits &= bits - 1;
            const int p = B.sync_order[wi * 32 + bit];
            const int node = B.proxy_s[p].z;
            t.remove_leaf(node);
            t.aabb[node * t.stride] = B.p_fat[x.at(B.NP, p)];
            t.insert_leaf(node);
            t.moved[node * t.stride] = 1;
            if (mc >= B.NMOVE) { ws[WS_STATUS] = B2GPU_E_CAPACITY; break; }
            B.move_buf[x.at(B.NMOVE, mc++)] = node;
          }
        }
        ws[WS_MOVE_COUNT] = mc;
        ws[WS_EV_MOVED] = 0;
      }
      update_pairs(B, x, ws, b_chead, c_next);
      ws[WS_INV_DT0] = f2i(sp.inv_dt);
    }
    ws[WS_ST_CONTACTS] = ws[WS_CONTACT_COUNT];
  }
};

// The ISLAND bit of contacts is only consumed by the island DFS itself, so the shared-memory DFS keeps it
// in shared memory; before a snapshot leaves the device this stage makes the bit match what the reference
// would hold after the same step: set exactly on the contacts of the current island order.
struct ContactIslandFlagsK {  // flat over contact slots
  Batch B;
  int phase;  // 0: clear every live contact's bit, 1: set it on island contacts
  B2G_HD void operator()(int tid) const {
    int w, c;
    if (!flat_decode(B, tid, B.NC, w, c)) return;
    WIdx x = widx(B, w);
    Ws ws = ws_of(B, x);
    if (!ws[WS_ISL_VALID]) return;
    if (phase == 0) {
      if (c < ws[WS_CONTACT_COUNT]) B.c_flags[x.at(B.NC, c)] &= ~B2GPU_CONTACT_ISLAND;
    } else if (c < ws[WS_ISL_CONTACTS]) {
      B.c_flags[x.at(B.NC, B.isl_contact[x.at(B.NC, c)])] |= B2GPU_CONTACT_ISLAND;
    }
  }
};

// clear_forces (b2_world.rs(private):961-967): flat over bodies.
struct BodyEndK {
  Batch B;
  B2G_HD void operator()(int tid) const {
    int w, b;
    if (!flat_decode(B, tid, B.NB, w, b)) return;
    WIdx x = widx(B, w);
    Ws ws = ws_of(B, x);
    if (ws[WS_FLAGS] & B2GPU_WORLD_CLEAR_FORCES) {
      const int bi = x.at(B.NB, b);
      float4 fo = B.b_force[bi];
      fo.x = 0.0f; fo.y = 0.0f; fo.z = 0.0f;
      B.b_force[bi] = fo;
    }
  }
};

}  // namespace b2g
