// b2g_large.h — data-parallel forms of the ordered stages for ONE large world (LB = 1, configs 2/4/5 of
// BASELINE.json: 10k–100k bodies).  b2g_step.h runs the order-dependent parts of a step as one thread per
// world, which is the right shape for thousands of small worlds and hopeless for one world of 100k bodies
// (558 ms of tree maintenance + 58 ms of island DFS per step, profiles/r01_single_world_*.json).  The stages
// here produce the same results from flat kernels plus scans and sorts:
//
//  * broadphase ("set-exact", SURVEY H1 option ii): move_proxy's keep-or-refatten decision stays the flat
//    SyncFixturesK; the moved proxies are compacted into the move buffer in synchronize order; an LBVH
//    (Morton order of the fat-box centres, Karras construction, bottom-up refit) is rebuilt over all proxies;
//    every moved proxy queries it (one thread per query, count -> scan -> emit) under the reference's pair
//    rule (b2_broad_phase.rs(private):86-111: skip self, skip a moved partner with the larger id);
//    add_pair's tests (b2_contact_manager.rs(private):178-302) run flat over the candidate pairs and the
//    survivors are appended by a prefix sum.  The pair SET and the created contact SET are the reference's;
//    the creation ORDER inside one update_pairs call is (move-buffer order, LBVH traversal order) instead of
//    (move-buffer order, reference-tree traversal order), because the reference's order is a function of its
//    incrementally balanced tree, whose maintenance is sequential.  The replica tree is not maintained in
//    this mode (its leaf boxes are; the internal nodes are refitted on download).
//  * contact destruction: flag -> scan -> stable compaction (same order of survivors as the reference list).
//  * per-body contact edge lists (push_front lists = each body's contacts by descending index): kept as contiguous
//    rows (CSR), rebuilt by one radix sort of (body, edge) keys whenever the contact set changed.
//  * islands (b2_world.rs(private):376-507): connected components by a lock-free union-find over the eligible
//    contacts give each island's member set, its seed (the newest awake body: the reference's seed loop runs
//    newest first) and its body / contact counts; a prefix sum in seed order lays out the island arrays; then
//    ONE THREAD PER ISLAND runs the reference's LIFO traversal from its seed, so every island's contact order —
//    the Gauss-Seidel order — is exactly the reference's.  Static bodies never propagate an island and are
//    not listed (they only needed their rotation cached, done flat by LwStaticRotK).
//
// Every functor is B2G_HD and is stepped by the test-only host simulator as well (tests/hostsim).
#pragma once
#include "b2g_step.h"

#if defined(__CUDA_ARCH__)
#define B2G_ATOMIC_MIN(p, v) atomicMin((p), (v))
#define B2G_ATOMIC_CAS(p, c, v) atomicCAS((p), (c), (v))
#define B2G_FENCE() __threadfence()
#define B2G_VLOAD(p) (*(volatile const int*)(p))
#else
#define B2G_ATOMIC_MIN(p, v) (*(p) = (*(p) < (v) ? *(p) : (v)))
#define B2G_FENCE()
#define B2G_VLOAD(p) (*(p))
static inline int b2g_host_cas(int* p, int c, int v) { int o = *p; if (o == c) *p = v; return o; }
#define B2G_ATOMIC_CAS(p, c, v) b2g_host_cas((p), (c), (v))
#endif

namespace b2g {

typedef unsigned long long u64;

struct Large {  // device scratch of the large-world mode
  // sort keys (LBVH Morton keys, edge-list keys)
  u64* keys;      // [NK]
  u64* keys_alt;  // [NK]
  unsigned* sort_in;   // [2 NP] Morton codes, then proxy indices (32-bit key / value pairs of the LBVH sort)
  unsigned* sort_out;  // [2 NP]
  // LBVH over the NP proxies: internal nodes 0..n-2 (0 = root), leaves by sorted position
  float4* lb_box;  // [NP] boxes of internal nodes
  int2* lb_child;  // [NP] children: >= 0 internal node, < 0 leaf at sorted position ~c
  int* lb_parent;  // [2 NP] parent of internal node i at [i], of the leaf at sorted position s at [NP + s]
  int* lb_flag;    // [NP] refit arrival counters
  int* lb_leaf;    // [NP] tree node id (= proxy id of the reference) of the leaf at sorted position s
  // pair finding
  int* q_cnt;      // [NMOVE + 1] candidates per moved proxy; moved-word popcounts
  int* q_off;      // [NMOVE + 1]
  int* q_local;    // [NMOVE][LW_QLOCAL] the first candidates of each query, kept by the counting pass
  int2* cand;      // [NCAND] candidate pairs (query node id, other node id)
  int* cand_flag;  // [NCAND + 1] 1 = add_pair creates a contact
  int* cand_pos;   // [NCAND + 1]
  int4* cand_fix;  // [NCAND] (fixture_a, fixture_b, index_a, index_b) after the register-order swap
  int NCAND;
  float4* scratch4; // [16] store target for bodies that must not be written (immovable: shared between islands)
  int4* vc_idx;    // [NC] per island contact slot: (body A, body B, velocity points, -), see LwVelocity4K
  int* first_idx;  // [NN] first move-buffer index of a tree node (host edits can buffer a proxy more than once)
  // level schedule of giant islands (b2g_levels.h)
  int* sleep_min;     // [NB] per island: bits of the minimum sleep time of its bodies (LwSleepK)
  int* lv_meta;       // [4] [0] = giant islands chosen at the last island rebuild
  int4* lv_info;      // [LW_MAXG] (island, first constraint, constraints, levels)
  int* lv_isl_giant;  // [NB] per island: 1 = swept by the level-scheduled kernels
  int* lv_last;       // [NB] per body: level after its latest constraint (build scratch)
  int* lv_level;      // [NC] level of the constraint at island-order position k
  int* lv_count;      // [NC + NB + 2] constraints per level (build scratch), at first + island + level
  int* lv_start;      // [NC + NB + 2] first position of a level in lv_order, at first + island + level
  int* lv_order;      // [NC] constraints in level order, per island at its contact range
  int4* lv_ix;        // [NC] vc_idx of the constraint at each position of the level order (refreshed every step)
  float4* lv_vrec;    // [NC * 8] velocity records q0..q7 in level order (refreshed every step; q6 follows the sweeps)
  float4* lv_prec;    // [NC * 5] position records p0..p4 in level order (refreshed every step)
  // islands
  int* uf_parent;  // [NB]
  int* cnt_b;      // [NB] per root: non-static bodies
  int* cnt_c;      // [NB] per root: eligible contacts
  int* seed;       // [NB] per root: newest awake, enabled, non-static member or -1
  u64* pk_in;      // [NB + 1] islands in seed order: 1 << 44 | bodies << 24 | contacts
  u64* pk_out;     // [NB + 1]
  int* isl_seed;   // [NB]
  int* cnt_j;      // [NB] per root: joints of the component (joints with an enabled other body)
  int* pj_in;      // [NB + 1] joint counts in seed order
  int* pj_out;     // [NB + 1]
  int* wake_idx;   // [NB + 1] collide wake-up cascade (LwWakeK); [NB] = another round needed
  int* state;      // [NB] island traversal: 0 unvisited, 1 on the stack, 2 listed
  // per-body contact rows (CSR): the edges 2 c + side of a body, ascending = oldest first
  int* adj;        // [2 NC] edge ids sorted by (body, edge)
  int2* eadj;      // [2 NC] the ELIGIBLE row entries only (enabled, touching, no sensor), refreshed when islands are
                   // rebuilt: (edge, other body | static << 30 | 1 << 31), same order
  int2* erow;      // [NB] begin / end of a body's eligible entries in eadj
  int* row_start;  // [NB]
  int* row_end;    // [NB]
  // destroy compaction
  int* keep_flag;  // [NC + 1]
  int* keep_pos;   // [NC + 1]
};

B2G_HD Box lw_box(const float4* a, int i) { return load_box(a, i); }

// ------------------------------------------------------------------------------------------
// union-find (lock-free hooking by atomicCAS, path halving; the larger index becomes the root)
// ------------------------------------------------------------------------------------------
B2G_HD int uf_find(int* parent, int x) {
  for (;;) {
    const int p = B2G_VLOAD(&parent[x]);
    if (p == x) return x;
    const int gp = B2G_VLOAD(&parent[p]);
    if (gp != p) parent[x] = gp;  // only ever points further up: safe under concurrent unions
    x = p;
  }
}
B2G_HD void uf_union(int* parent, int a, int b) {
  for (;;) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }
    if (B2G_ATOMIC_CAS(&parent[b], b, a) == b) return;
  }
}

// ------------------------------------------------------------------------------------------
// prologue: wake merge, contact destruction
// ------------------------------------------------------------------------------------------
struct LwStatsResetK {  // one thread: what TreePairsK(pre_step) does before the optional find_new_contacts
  Batch B;
  B2G_HD void operator()(int) const {
    for (int s = WS_ST_CONTACTS; s <= WS_ST_LEVELS; ++s) B.ws[s] = 0;
  }
};

struct LwClearNewContactsK {  // one thread
  Batch B;
  B2G_HD void operator()(int) const { B.ws[WS_FLAGS] &= ~B2GPU_WORLD_NEW_CONTACTS; }
};

// Wake-up cascade inside collide (only when the flat narrowphase woke a sleeping body).  The reference's loop
// visits contacts newest first; a contact it skipped as inactive in the flat pass must be evaluated after all
// iff one of its bodies is woken by a contact with a LARGER index.  wake_idx[b] = largest index of a contact
// whose touching state changed on body b.  Rounds of a flat kernel evaluate every skipped contact c with
// wake_idx[body] > c; an evaluation that changes touching raises wake_idx and asks for another round.  The
// fixpoint is the reference's result: by induction over descending c, "woken by then" only ever depends on
// contacts above c.  phase 0: reset (flat over bodies), 1: seed from the flat pass (flat over contacts),
// 2: one round (flat over contacts), 3: done (one thread).
struct LwWakeK {
  Batch B;
  Large L;
  int* b_wake;
  int cc, phase;
  B2G_HD void operator()(int t) const {
    if (phase == 0) {
      if (t < B.NB) L.wake_idx[t] = -1;
      if (t == 0) L.wake_idx[B.NB] = 0;  // "another round" flag
      return;
    }
    if (phase == 3) { B.ws[WS_EV_WAKE] = 0; L.wake_idx[B.NB] = 0; return; }
    if (t >= cc) return;
    const int flags = B.c_flags[t];
    if (phase == 1) {
      if (!(flags & CF_WOKE)) return;
      const int4 fx = B.c_fix[t];
      const int ba = B.fixtures[fx.x].body, bb = B.fixtures[fx.y].body;
      if (body_type(B.b_flags[ba]) != B2GPU_STATIC_BODY) B2G_ATOMIC_MAX(&L.wake_idx[ba], t);
      if (body_type(B.b_flags[bb]) != B2GPU_STATIC_BODY) B2G_ATOMIC_MAX(&L.wake_idx[bb], t);
      return;
    }
    if (!(flags & CF_SKIPPED)) return;
    const int4 fx = B.c_fix[t];
    const int ba = B.fixtures[fx.x].body, bb = B.fixtures[fx.y].body;
    const bool wa = body_type(B.b_flags[ba]) != B2GPU_STATIC_BODY && L.wake_idx[ba] > t;
    const bool wb = body_type(B.b_flags[bb]) != B2GPU_STATIC_BODY && L.wake_idx[bb] > t;
    if (!wa && !wb) return;
    WIdx x = widx(B, 0);
    Ws ws = ws_of(B, x);
    collide_one(B, x, ws, t, b_wake, true, L.wake_idx);
    if (B.c_flags[t] & CF_WOKE) L.wake_idx[B.NB] = 1;
  }
};

struct LwWakeMergeK {  // flat over bodies: set_awake(true) for every body collide marked
  Batch B;
  int* b_wake;
  B2G_HD void operator()(int b) const {
    if (b >= B.NB || !b_wake[b]) return;
    b_wake[b] = 0;
    const int f = B.b_flags[b];
    if (body_type(f) == B2GPU_STATIC_BODY) return;
    B.b_flags[b] = f | B2GPU_BODY_AWAKE;
    B.b_pos[b].w = 0.0f;
  }
};

struct LwDestroyFlagK {  // flat over cc + 1 contact slots
  Batch B;
  Large L;
  int cc;
  B2G_HD void operator()(int c) const {
    if (c > cc) return;
    if (c == cc) { L.keep_flag[c] = 0; return; }
    const int flags = B.c_flags[c];
    const bool gone = (flags & CF_DESTROY) != 0;
    L.keep_flag[c] = gone ? 0 : 1;
    if (!gone) return;
    const int4 fx = B.c_fix[c];
    const b2gpu_fixture_rec& fa = B.fixtures[fx.x];
    const b2gpu_fixture_rec& fb = B.fixtures[fx.y];
    if (B.c_m3[c].w > 0 && !fa.is_sensor && !fb.is_sensor) {  // b2_contact.rs(private):39-45
      const int bs[2] = {fa.body, fb.body};
      for (int s = 0; s < 2; ++s) {
        const int b = bs[s];
        if (body_type(B.b_flags[b]) == B2GPU_STATIC_BODY) continue;
        B2G_ATOMIC_OR(&B.b_flags[b], B2GPU_BODY_AWAKE);
        B.b_pos[b].w = 0.0f;
      }
    }
  }
};

// Stable compaction through the (not yet written) velocity-constraint stream as scratch:
// sections of NC float4 each: fix, mat, m0, m1, m2, m3, flags.
struct LwCompactK {
  Batch B;
  Large L;
  int n;      // phase 0: old contact count; phase 1: new contact count
  int phase;  // 0: survivors -> scratch at their new index; 1: scratch -> contact arrays
  B2G_HD void operator()(int c) const {
    if (c >= n) return;
    float4* t = B.vc;
    const size_t NC = (size_t)B.NC;
    if (phase == 0) {
      if (!L.keep_flag[c]) return;
      const size_t k = (size_t)L.keep_pos[c];
      const int4 fx = B.c_fix[c], m3 = B.c_m3[c];
      t[k] = make_float4(i2f(fx.x), i2f(fx.y), i2f(fx.z), i2f(fx.w));
      t[NC + k] = B.c_mat[c];
      t[2 * NC + k] = B.c_m0[c];
      t[3 * NC + k] = B.c_m1[c];
      t[4 * NC + k] = B.c_m2[c];
      t[5 * NC + k] = make_float4(i2f(m3.x), i2f(m3.y), i2f(m3.z), i2f(m3.w));
      ((int*)(t + 6 * NC))[k] = B.c_flags[c];
    } else {
      const size_t k = (size_t)c;
      const float4 fx = t[k], m3 = t[5 * NC + k];
      B.c_fix[c] = make_int4(f2i(fx.x), f2i(fx.y), f2i(fx.z), f2i(fx.w));
      B.c_mat[c] = t[NC + k];
      B.c_m0[c] = t[2 * NC + k];
      B.c_m1[c] = t[3 * NC + k];
      B.c_m2[c] = t[4 * NC + k];
      B.c_m3[c] = make_int4(f2i(m3.x), f2i(m3.y), f2i(m3.z), f2i(m3.w));
      B.c_flags[c] = ((const int*)(t + 6 * NC))[k];
    }
  }
};

struct LwDestroyFinishK {  // one thread
  Batch B;
  StepParams sp;
  int old_cc, new_cc;
  B2G_HD void operator()(int) const {
    int* ws = B.ws;
    ws[WS_ST_DESTROYED] += old_cc - new_cc;
    ws[WS_CONTACT_COUNT] = new_cc;
    ws[WS_EV_DESTROY] = 0;
    ws[WS_TOPO_DIRTY] = 1;
    if (!(sp.dt > 0.0f)) ws[WS_ISL_VALID] = 0;  // collide-only step: the island order no longer describes the contacts
  }
};

// ------------------------------------------------------------------------------------------
// per-body contact edge lists from a sort of (body, edge) keys
// ------------------------------------------------------------------------------------------
struct LwEdgeKeysK {  // flat over max(2 cc, NB): keys of the 2 cc edges; rows reset
  Batch B;
  Large L;
  int cc, edge_bits;
  B2G_HD void operator()(int t) const {
    if (t < B.NB) { L.row_start[t] = 0; L.row_end[t] = 0; }
    if (t >= 2 * cc) return;
    const int c = t >> 1;
    const int4 fx = B.c_fix[c];
    const int body = (t & 1) ? B.fixtures[fx.y].body : B.fixtures[fx.x].body;
    L.keys[t] = ((u64)(unsigned)body << edge_bits) | (u64)(unsigned)t;
  }
};
struct LwEdgeRowsK {  // flat over the 2 cc sorted keys: one contiguous row of edges per body
  Batch B;
  Large L;
  int n, edge_bits;
  B2G_HD void operator()(int i) const {
    if (i >= n) return;
    const u64 k = L.keys_alt[i];
    const u64 mask = ((u64)1 << edge_bits) - 1;
    const u64 body = k >> edge_bits;
    L.adj[i] = (int)(k & mask);
    if (i == 0 || (L.keys_alt[i - 1] >> edge_bits) != body) L.row_start[(int)body] = i;
    if (i == n - 1 || (L.keys_alt[i + 1] >> edge_bits) != body) L.row_end[(int)body] = i + 1;
  }
};

// ------------------------------------------------------------------------------------------
// islands
// ------------------------------------------------------------------------------------------
B2G_HD bool lw_contact_eligible(const Batch& B, int c, int& ba, int& bb) {
  const int cf = B.c_flags[c];
  if (!(cf & B2GPU_CONTACT_ENABLED) || !(cf & B2GPU_CONTACT_TOUCHING)) return false;
  const int4 fx = B.c_fix[c];
  const b2gpu_fixture_rec& fa = B.fixtures[fx.x];
  const b2gpu_fixture_rec& fb = B.fixtures[fx.y];
  if (fa.is_sensor || fb.is_sensor) return false;
  ba = fa.body;
  bb = fb.body;
  return true;
}

struct LwIslInitK {  // flat over max(NB, cc)
  Batch B;
  Large L;
  int cc;
  B2G_HD void operator()(int t) const {
    if (t < B.NB) {
      L.uf_parent[t] = t;
      L.cnt_b[t] = 0;
      L.cnt_c[t] = 0;
      if (B.NJ > 0) L.cnt_j[t] = 0;
      L.seed[t] = -1;
      L.state[t] = 0;
      B.b_flags[t] &= ~B2GPU_BODY_ISLAND;
    }
    if (t < cc) B.c_flags[t] &= ~B2GPU_CONTACT_ISLAND;
  }
};
struct LwUnionK {  // flat over contacts
  Batch B;
  Large L;
  int cc;
  B2G_HD void operator()(int c) const {
    if (c < B.NJ) {  // joints connect their two bodies like contacts do (b2_world.rs(private):461-483): static bodies do not
                     // propagate, and a joint to a disabled body is not simulated
      const b2gpu_joint_rec& jr = B.joints[c];
      const int fa = B.b_flags[jr.body_a], fb = B.b_flags[jr.body_b];
      if (body_type(fa) != B2GPU_STATIC_BODY && body_type(fb) != B2GPU_STATIC_BODY && (fa & B2GPU_BODY_ENABLED) &&
          (fb & B2GPU_BODY_ENABLED))
        uf_union(L.uf_parent, jr.body_a, jr.body_b);
    }
    if (c >= cc) return;
    int ba, bb;
    if (!lw_contact_eligible(B, c, ba, bb)) return;
    if (body_type(B.b_flags[ba]) == B2GPU_STATIC_BODY || body_type(B.b_flags[bb]) == B2GPU_STATIC_BODY) return;
    uf_union(L.uf_parent, ba, bb);
  }
};
struct LwCountK {  // flat over max(NB, cc): sizes and seed of every component, at its root
  Batch B;
  Large L;
  int cc;
  B2G_HD void operator()(int t) const {
    if (t < B.NB) {
      const int f = B.b_flags[t];
      if (body_type(f) != B2GPU_STATIC_BODY) {
        const int r = uf_find(L.uf_parent, t);
        B2G_ATOMIC_ADD(&L.cnt_b[r], 1);
        if ((f & B2GPU_BODY_AWAKE) && (f & B2GPU_BODY_ENABLED)) B2G_ATOMIC_MAX(&L.seed[r], t);
      }
    }
    if (t < cc) {
      int ba, bb;
      if (lw_contact_eligible(B, t, ba, bb)) {
        const int m = body_type(B.b_flags[ba]) != B2GPU_STATIC_BODY ? ba : bb;
        if (body_type(B.b_flags[m]) != B2GPU_STATIC_BODY) B2G_ATOMIC_ADD(&L.cnt_c[uf_find(L.uf_parent, m)], 1);
      }
    }
    if (t < B.NJ) {  // a joint belongs to the island of its movable body (both movable: the same island)
      const b2gpu_joint_rec& jr = B.joints[t];
      const int fa = B.b_flags[jr.body_a], fb = B.b_flags[jr.body_b];
      if ((fa & B2GPU_BODY_ENABLED) && (fb & B2GPU_BODY_ENABLED)) {
        const int m = body_type(fa) != B2GPU_STATIC_BODY ? jr.body_a : jr.body_b;
        if (body_type(B.b_flags[m]) != B2GPU_STATIC_BODY) B2G_ATOMIC_ADD(&L.cnt_j[uf_find(L.uf_parent, m)], 1);
      }
    }
  }
};
enum { LW_PK_ISL = 44, LW_PK_BODY = 24 };
struct LwSeedPackK {  // flat over NB + 1, newest body first: one entry per island at its seed
  Batch B;
  Large L;
  B2G_HD void operator()(int i) const {
    if (i > B.NB) return;
    u64 v = 0;
    if (i < B.NB) {
      const int b = B.NB - 1 - i;
      if (body_type(B.b_flags[b]) != B2GPU_STATIC_BODY) {
        const int r = uf_find(L.uf_parent, b);
        L.uf_parent[b] = r;
        if (L.seed[r] == b) {
          v = ((u64)1 << LW_PK_ISL) | ((u64)L.cnt_b[r] << LW_PK_BODY) | (u64)L.cnt_c[r];
          if (B.NJ > 0) L.pj_in[i] = L.cnt_j[r];
        } else if (B.NJ > 0) L.pj_in[i] = 0;
      } else if (B.NJ > 0) L.pj_in[i] = 0;
    } else if (B.NJ > 0) L.pj_in[i] = 0;
    L.pk_in[i] = v;
  }
};
struct LwRangeK {  // flat over NB + 1
  Batch B;
  Large L;
  B2G_HD void operator()(int i) const {
    if (i > B.NB) return;
    const u64 o = L.pk_out[i];
    const int isl = (int)(o >> LW_PK_ISL), bf = (int)((o >> LW_PK_BODY) & 0xfffff), cf = (int)(o & 0xffffff);
    if (i == B.NB) {
      int* ws = B.ws;
      ws[WS_ISL_COUNT] = isl; ws[WS_ISL_BODIES] = bf; ws[WS_ISL_CONTACTS] = cf;
      ws[WS_ST_ISLANDS] = isl; ws[WS_ST_ISL_BODIES] = bf; ws[WS_ST_ISL_CONTACTS] = cf;
      ws[WS_TOPO_DIRTY] = 0;  // the island traversal raises it again when a sleeper joined
      ws[WS_ISL_VALID] = 1;
      ws[WS_SCHED_ROUNDS] = -1;
      ws[WS_ISL_JOINTS] = B.NJ > 0 ? L.pj_out[i] : 0;
      return;
    }
    if (!(L.pk_in[i] >> LW_PK_ISL)) return;
    const int b = B.NB - 1 - i;
    const int r = L.uf_parent[b];
    B.isl_range[isl] = make_int4(bf, bf + L.cnt_b[r], cf, cf + L.cnt_c[r]);
    if (B.NJ > 0) B.isl_jrange[isl] = make_int2(L.pj_out[i], L.pj_out[i] + L.cnt_j[r]);
    L.isl_seed[isl] = b;
  }
};
struct LwAdjInfoK {  // flat over the 2 cc row entries (+1 tail): which edges the traversal may follow
  Batch B;
  Large L;
  int n;
  B2G_HD void operator()(int i) const {
    if (i > n) return;
    int ba, bb;
    L.cand_flag[i] = (i < n && lw_contact_eligible(B, L.adj[i] >> 1, ba, bb)) ? 1 : 0;
  }
};
struct LwAdjCompactK {  // flat over max(2 cc, NB): eligible entries packed in row order; per-body ranges
  Batch B;
  Large L;
  int n;
  B2G_HD void operator()(int i) const {
    if (i < B.NB) L.erow[i] = make_int2(L.cand_pos[L.row_start[i]], L.cand_pos[L.row_end[i]]);
    if (i >= n || !L.cand_flag[i]) return;
    const int e = L.adj[i];
    const int4 fx = B.c_fix[e >> 1];
    const int other = (e & 1) ? B.fixtures[fx.x].body : B.fixtures[fx.y].body;
    L.eadj[L.cand_pos[i]] = make_int2(e, other | (int)0x80000000u | (body_type(B.b_flags[other]) == B2GPU_STATIC_BODY ? 0x40000000 : 0));
  }
};
// One thread per island: the reference's traversal from the island's seed (LIFO stack, every unvisited
// neighbour pushed when a body is listed, each body's edges newest first).  Works on the rows of ELIGIBLE
// edges: a row is contiguous, so the loads of one listed body are independent of each other (a linked list would
// chain them), they are requested four edges at a time, and the many non-touching contacts of a crowded body
// (AddPair: ~60 per circle) cost nothing.  "Contact already in the island" needs no contact flag: an
// eligible contact was added when the first of its two movable bodies was listed, so it is skipped exactly when
// the other body is already listed (state 2).  Body / contact ISLAND flags are set flat afterwards.
struct LwDfsNoHook {
  B2G_HD void pushed(int) const {}
};
// hook.pushed(body): called for every body put on the stack (the giant-island form hands it to a prefetching warp)
template <class Hook>
B2G_HD void lw_dfs_walk(const Batch& B, const Large& L, int* stack, int isl, Hook& hook) {
  const int4 rg = B.isl_range[isl];
  int* st = stack + rg.x;
  int nb = rg.x, nc = rg.z, sp_ = 0;
  int nj = B.NJ > 0 ? B.isl_jrange[isl].x : 0;
  const int seed = L.isl_seed[isl];
  st[sp_++] = seed;
  L.state[seed] = 1;
  while (sp_ > 0) {
    const int b = st[--sp_];
    B.isl_body[nb++] = b;
    L.state[b] = 2;
    const int2 row = L.erow[b];
    const int r0 = row.x, r1 = row.y;
    for (int i = r1 - 1; i >= r0; i -= 4) {
      int e[4], info[4], sv[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int j = 0; j < 4; ++j) {
        const int2 ent = i - j >= r0 ? L.eadj[i - j] : make_int2(0, 0);
        e[j] = ent.x;
        info[j] = ent.y;
      }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int j = 0; j < 4; ++j) sv[j] = (info[j] < 0 && !(info[j] & 0x40000000)) ? L.state[info[j] & 0x3fffffff] : 0;
      int pushed[4] = {-1, -1, -1, -1};
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
      for (int j = 0; j < 4; ++j) {
        if (info[j] >= 0) continue;  // not eligible (or padding)
        const int other = info[j] & 0x3fffffff;
        const bool is_static = (info[j] & 0x40000000) != 0;
        if (!is_static && sv[j] == 2) continue;  // added when `other` was listed
        B.isl_contact[nc] = e[j] >> 1;
        B.c_isl[nc] = isl;
        ++nc;
        if (is_static || sv[j] != 0) continue;  // static bodies never propagate; not listed in this mode
        if (other == pushed[0] || other == pushed[1] || other == pushed[2]) continue;  // pushed by an earlier edge of this group
        pushed[j] = other;
        st[sp_++] = other;
        L.state[other] = 1;
        hook.pushed(other);
      }
    }
    // joints of this body, newest edge first (b2_world.rs(private):461-483).  "Already in the island" needs no joint
    // flag either: a joint is added when the first of its movable bodies is listed
    if (B.NJ > 0)
      for (int q = B.jadj_off[b]; q < B.jadj_off[b + 1]; ++q) {
        const int je = B.jadj[q], jn = je >> 1;
        const int other = (je & 1) ? B.joints[jn].body_a : B.joints[jn].body_b;
        const int of = B.b_flags[other];
        if (!(of & B2GPU_BODY_ENABLED)) continue;
        const bool is_static = body_type(of) == B2GPU_STATIC_BODY;
        const int so = is_static ? 0 : L.state[other];
        if (!is_static && so == 2) continue;
        B.isl_joint[nj++] = jn;
        if (is_static || so != 0) continue;
        st[sp_++] = other;
        L.state[other] = 1;
        hook.pushed(other);
      }
  }
  if (nb != rg.y || nc != rg.w || (B.NJ > 0 && nj != B.isl_jrange[isl].y)) B.ws[WS_STATUS] = B2GPU_E_INTERNAL;
}
struct LwDfsK {
  Batch B;
  Large L;
  int* stack;  // [NB]: island i uses the slots of its body range
  int n_islands;
  B2G_HD void operator()(int isl) const {
    if (isl >= n_islands || L.lv_isl_giant[isl]) return;  // a giant island is walked by LwDfsGiantK (b2g_levels.h)
    LwDfsNoHook hook;
    lw_dfs_walk(B, L, stack, isl, hook);
  }
};
struct LwIslFlagsK {  // flat over max(island bodies, island contacts): what the traversal leaves on bodies and contacts
  Batch B;
  int nib, nic;
  B2G_HD void operator()(int k) const {
    if (k < nib) {
      const int b = B.isl_body[k];
      const int bf = B.b_flags[b];
      if (!(bf & B2GPU_BODY_AWAKE)) B.ws[WS_TOPO_DIRTY] = 1;  // a sleeper joined: it is a seed candidate next step
      B.b_flags[b] = bf | B2GPU_BODY_ISLAND | B2GPU_BODY_AWAKE;
    }
    if (k < nic) B.c_flags[B.isl_contact[k]] |= B2GPU_CONTACT_ISLAND;
  }
};
struct LwIslCachedK {  // one thread: nothing the island order depends on changed
  Batch B;
  B2G_HD void operator()(int) const {
    int* ws = B.ws;
    ws[WS_ST_ISLANDS] = ws[WS_ISL_COUNT];
    ws[WS_ST_ISL_BODIES] = ws[WS_ISL_BODIES];
    ws[WS_ST_ISL_CONTACTS] = ws[WS_ISL_CONTACTS];
  }
};
struct LwStaticRotK {  // flat over bodies: what IntegrateK does for the static members of an island
  Batch B;
  B2G_HD void operator()(int b) const {
    if (b >= B.NB || body_type(B.b_flags[b]) != B2GPU_STATIC_BODY) return;
    const float4 pos = B.b_pos[b];
    B.b_pos0[b] = make_float4(pos.x, pos.y, pos.z, 0.0f);
    const Rot q = rot_from_angle(pos.z);
    B.b_rot[b] = make_float4(q.s, q.c, q.s, q.c);
  }
};

// ------------------------------------------------------------------------------------------
// broadphase: move buffer, LBVH, pair queries, add_pair
// ------------------------------------------------------------------------------------------
B2G_HD int popcount32(unsigned v) {
#if defined(__CUDA_ARCH__)
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}
struct LwMoveCountK {  // flat over NMW + 1 bitmap words
  Batch B;
  Large L;
  B2G_HD void operator()(int w) const {
    if (w > B.NMW) return;
    L.q_cnt[w] = w < B.NMW ? popcount32((unsigned)B.p_move[w]) : 0;
  }
};
struct LwMoveEmitK {  // flat over bitmap words: re-fattened proxies enter the move buffer in synchronize order
  Batch B;
  Large L;
  int base;  // move count before this step's moves
  B2G_HD void operator()(int w) const {
    if (w >= B.NMW) return;
    unsigned bits = (unsigned)B.p_move[w];
    if (!bits) return;
    B.p_move[w] = 0;
    int at = base + L.q_off[w];
    while (bits) {
      const int bit = lowest_bit(bits);
      bits &= bits - 1;
      const int p = B.sync_order[w * 32 + bit];
      const int node = B.proxy_s[p].z;
      B.n_aabb[node] = B.p_fat[p];
      B.n_moved[node] = 1;
      if (at < B.NMOVE) B.move_buf[at] = node;
      ++at;
    }
  }
};
struct LwMoveFinishK {  // one thread
  Batch B;
  int mc;
  B2G_HD void operator()(int) const {
    B.ws[WS_MOVE_COUNT] = mc;
    B.ws[WS_EV_MOVED] = 0;
  }
};

// Exact-order variant of the mode (b2gpu_world_set_large_mode(w, 2)): the replica of the reference's tree IS
// maintained — one thread re-inserts the moved proxies in synchronize order, exactly as TreePairsK does — and
// every query walks that tree, so contacts are created in the reference's order and free-running state stays
// bit-identical to the reference.  Everything else of the mode is unchanged.  The price is the sequential
// re-insertion: fine while few proxies move per step (AddPair: hundreds), slow when most of a 100k world moves.
struct LwTreeMoveK {  // one thread
  Batch B;
  B2G_HD void operator()(int) const {
    WIdx x = widx(B, 0);
    Ws ws = ws_of(B, x);
    if (!ws[WS_EV_MOVED]) return;
    Tree t = tree_of(B, x, ws);
    int mc = ws[WS_MOVE_COUNT];
    for (int wi = 0; wi < B.NMW; ++wi) {
      unsigned bits = (unsigned)B.p_move[wi];
      if (!bits) continue;
      B.p_move[wi] = 0;
      while (bits) {
        const int bit = lowest_bit(bits);
        bits &= bits - 1;
        const int p = B.sync_order[wi * 32 + bit];
        const int node = B.proxy_s[p].z;
        t.remove_leaf(node);
        t.aabb[node] = B.p_fat[p];
        t.insert_leaf(node);
        t.moved[node] = 1;
        if (mc >= B.NMOVE) { ws[WS_STATUS] = B2GPU_E_CAPACITY; break; }
        B.move_buf[mc++] = node;
      }
    }
    ws[WS_MOVE_COUNT] = mc;
    ws[WS_EV_MOVED] = 0;
  }
};

B2G_HD unsigned lw_spread16(unsigned v) {  // 16 bits -> even bit positions
  v &= 0xffffu;
  v = (v | (v << 8)) & 0x00ff00ffu;
  v = (v | (v << 4)) & 0x0f0f0f0fu;
  v = (v | (v << 2)) & 0x33333333u;
  v = (v | (v << 1)) & 0x55555555u;
  return v;
}
struct LwMortonK {  // flat over proxies: key = Morton code of the fat-box centre (1/32 m cells, +-1024 m) << 32 | proxy
  Batch B;
  Large L;
  B2G_HD void operator()(int p) const {
    if (p >= B.NP) return;
    const float4 a = B.n_aabb[B.proxy_s[p].z];
    const float cx = 0.5f * (a.x + a.z), cy = 0.5f * (a.y + a.w);
    float fx = cx * 32.0f + 32768.0f, fy = cy * 32.0f + 32768.0f;
    fx = fx > 0.0f ? fx : 0.0f;  // also maps NaN to cell 0
    fy = fy > 0.0f ? fy : 0.0f;
    const unsigned qx = fx < 65535.0f ? (unsigned)fx : 65535u, qy = fy < 65535.0f ? (unsigned)fy : 65535u;
    const unsigned code = lw_spread16(qx) | (lw_spread16(qy) << 1);
    // (code, proxy) as a 32-bit key / value pair: a stable 4-pass radix sort by code leaves equal codes in proxy order,
    // i.e. the order of the unique 64-bit keys code << 32 | proxy that LwKeyPackK rebuilds (an 8-pass sort before)
    L.sort_in[p] = code;
    L.sort_in[B.NP + p] = (unsigned)p;
  }
};
struct LwKeyPackK {  // flat over proxies: sorted (code, proxy) pairs -> unique 64-bit keys for the radix-tree construction
  Batch B;
  Large L;
  B2G_HD void operator()(int i) const {
    if (i >= B.NP) return;
    L.keys_alt[i] = ((u64)L.sort_out[i] << 32) | (u64)L.sort_out[B.NP + i];
  }
};
B2G_HD int lw_clz64(u64 v) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)v);
#else
  return v ? __builtin_clzll(v) : 64;
#endif
}
B2G_HD int lw_delta(const u64* k, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  return lw_clz64(k[i] ^ k[j]);  // keys are unique (proxy index in the low word)
}
struct LwKarrasK {  // flat over n: internal node i < n - 1 (Karras 2012), leaf table
  Batch B;
  Large L;
  int n;
  B2G_HD void operator()(int i) const {
    if (i >= n) return;
    const u64* k = L.keys_alt;
    L.lb_leaf[i] = B.proxy_s[(int)(k[i] & 0xffffffffu)].z;
    if (i >= n - 1) return;
    L.lb_flag[i] = 0;
    const int d = lw_delta(k, n, i, i + 1) - lw_delta(k, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = lw_delta(k, n, i, i - d);
    int lmax = 2;
    while (lw_delta(k, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
      if (lw_delta(k, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = lw_delta(k, n, i, j);
    int s = 0;
    for (int t = l;;) {
      t = (t + 1) >> 1;
      if (lw_delta(k, n, i, i + (s + t) * d) > dnode) s += t;
      if (t <= 1) break;
    }
    const int gamma = i + s * d + (d < 0 ? d : 0);
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    int2 ch;
    if (lo == gamma) { ch.x = ~gamma; L.lb_parent[B.NP + gamma] = i; } else { ch.x = gamma; L.lb_parent[gamma] = i; }
    if (hi == gamma + 1) { ch.y = ~(gamma + 1); L.lb_parent[B.NP + gamma + 1] = i; } else { ch.y = gamma + 1; L.lb_parent[gamma + 1] = i; }
    L.lb_child[i] = ch;
    if (i == 0) L.lb_parent[0] = -1;
  }
};
B2G_HD float4 lw_ldcg(const float4* p) {
#if defined(__CUDA_ARCH__)
  return __ldcg(p);
#else
  return *p;
#endif
}
struct LwRefitK {  // flat over leaves: the second thread to arrive at a node computes its box
  Batch B;
  Large L;
  int n;
  B2G_HD void operator()(int s) const {
    if (s >= n || n < 2) return;
    int node = L.lb_parent[B.NP + s];
    while (node != -1) {
      B2G_FENCE();
      if (B2G_ATOMIC_ADD(&L.lb_flag[node], 1) == 0) return;
      B2G_FENCE();
      const int2 ch = L.lb_child[node];
      // boxes of internal children were written by other threads of this launch: read them past L1 (a line
      // fetched earlier for a neighbouring node may hold a stale copy)
      const float4 a = ch.x >= 0 ? lw_ldcg(&L.lb_box[ch.x]) : B.n_aabb[L.lb_leaf[~ch.x]];
      const float4 b = ch.y >= 0 ? lw_ldcg(&L.lb_box[ch.y]) : B.n_aabb[L.lb_leaf[~ch.y]];
      L.lb_box[node] = make_float4(a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y, a.z > b.z ? a.z : b.z, a.w > b.w ? a.w : b.w);
      node = L.lb_parent[node];
    }
  }
};
B2G_HD bool lw_overlap(const float4 a, const Box& q) {  // b2_test_overlap(AABB), src/b2_collision.rs:355-368
  Box b;
  b.lo = v2(a.x, a.y);
  b.hi = v2(a.z, a.w);
  return box_overlap(b, q);
}
// A proxy edited twice on the host (created, then set_transform) sits twice in the move buffer; the reference
// queries it twice and add_pair rejects the second round of pairs because the first round's contacts exist by
// then.  Flat add_pair sees only the contacts that existed before the call, so the repeated queries are marked
// (their pairs still count as reported) and create nothing.  phase 0: reset, phase 1: first index per node.
struct LwMoveFirstK {
  Batch B;
  Large L;
  int mc, phase;
  B2G_HD void operator()(int i) const {
    if (i >= mc) return;
    const int q = B.move_buf[i];
    if (q == -1) return;
    if (phase == 0) L.first_idx[q] = 0x7fffffff;
    else B2G_ATOMIC_MIN(&L.first_idx[q], i);
  }
};
enum { LW_STACK = 128, LW_QLOCAL = 24 };
// One thread per move-buffer entry; emit = 0 counts, 1 writes the candidate pairs.  use_tree = 1 walks the
// uploaded replica of the reference's tree instead of the LBVH (child2 first, b2_dynamic_tree.rs:239-267): the
// find_new_contacts call at the top of a step (m_new_contacts) always follows an upload, whose tree is current,
// and its contacts enter an island in the same step — so there the reference's creation order is kept exactly.
struct LwQueryK {
  Batch B;
  Large L;
  int mc, n, emit, use_tree;
  B2G_HD void operator()(int i) const {
    if (i > mc) return;
    if (i == mc) { if (!emit) L.q_cnt[i] = 0; return; }
    const int q = B.move_buf[i];
    int count = 0;
    int2* out = emit ? L.cand + L.q_off[i] : nullptr;
    const int room = emit ? L.NCAND - L.q_off[i] : 0;
    int* local = L.q_local + (size_t)i * LW_QLOCAL;
    if (emit && q != -1) {
      // the counting pass kept the first LW_QLOCAL candidates: a query that fits needs no second walk
      const int cnt = L.q_off[i + 1] - L.q_off[i];
      if (cnt <= LW_QLOCAL) {
        const int qq = (use_tree && L.first_idx[q] != i) ? ~q : q;  // repeated move-buffer entry: report, never create
        for (int k = 0; k < cnt && k < room; ++k) out[k] = make_int2(qq, local[k]);
        return;
      }
    }
    if (q != -1 && n > 0) {
      const Box qb = lw_box(B.n_aabb, q);
      int stack[LW_STACK];
      int sp_ = 0;
      if (use_tree) {
        const int qq = L.first_idx[q] != i ? ~q : q;  // repeated entry: report, never create
        stack[sp_++] = B.ws[WS_TREE_ROOT];
        while (sp_ > 0) {
          const int id = stack[--sp_];
          if (id == -1) continue;
          if (!lw_overlap(B.n_aabb[id], qb)) continue;
          const int4 l = B.n_link[id];
          if (l.y == -1) {
            if (id == q) continue;
            if (B.n_moved[id] && id > q) continue;
            if (emit) { if (count < room) out[count] = make_int2(qq, id); }
            else if (count < LW_QLOCAL) local[count] = id;
            ++count;
          } else {
            if (sp_ + 2 > LW_STACK) { B.ws[WS_STATUS] = B2GPU_E_CAPACITY; continue; }
            stack[sp_++] = l.y;
            stack[sp_++] = l.z;
          }
        }
      } else {
        stack[sp_++] = n > 1 ? 0 : ~0;
        while (sp_ > 0) {
          const int c = stack[--sp_];
          if (c < 0) {
            // b2_broad_phase_query_callback (b2_broad_phase.rs(private):86-111)
            const int id = L.lb_leaf[~c];
            if (id == q) continue;
            if (!lw_overlap(B.n_aabb[id], qb)) continue;
            if (B.n_moved[id] && id > q) continue;
            if (emit) { if (count < room) out[count] = make_int2(q, id); }
            else if (count < LW_QLOCAL) local[count] = id;
            ++count;
          } else {
            if (!lw_overlap(L.lb_box[c], qb)) continue;
            const int2 ch = L.lb_child[c];
            if (sp_ + 2 > LW_STACK) { B.ws[WS_STATUS] = B2GPU_E_CAPACITY; continue; }
            stack[sp_++] = ch.y;
            stack[sp_++] = ch.x;
          }
        }
      }
    }
    if (!emit) L.q_cnt[i] = count;
  }
};
struct LwAddPairK {  // flat over candidates (+1 tail): the tests of add_pair, no side effects
  Batch B;
  Large L;
  int n;
  B2G_HD void operator()(int j) const {
    if (j > n) return;
    if (j == n) { L.cand_flag[j] = 0; return; }
    L.cand_flag[j] = 0;
    const int2 pr = L.cand[j];
    if (pr.x < 0) return;  // reported by a repeated move-buffer entry (see LwMoveFirstK)
    const int proxy_a = B.node_proxy[imin(pr.x, pr.y)], proxy_b = B.node_proxy[imax(pr.x, pr.y)];
    const int4 pa = B.proxy_s[proxy_a], pb = B.proxy_s[proxy_b];
    int fixture_a = pa.x, fixture_b = pb.x, index_a = pa.y, index_b = pb.y;
    const int body_a = pa.w, body_b = pb.w;
    if (body_a == body_b) return;
    // a contact between the two fixtures is in both bodies' contact rows: scan the row of the movable one when
    // the other is static (the ground's row holds every resting contact of the world)
    const int fb_ = B.b_flags[body_b], fa_ = B.b_flags[body_a];
    const int walk = (body_type(fb_) == B2GPU_STATIC_BODY && body_type(fa_) != B2GPU_STATIC_BODY) ? body_a : body_b;
    for (int i = L.row_start[walk], i1 = L.row_end[walk]; i < i1; ++i) {
      const int4 fx = B.c_fix[L.adj[i] >> 1];
      if (fx.x == fixture_a && fx.y == fixture_b && fx.z == index_a && fx.w == index_b) return;
      if (fx.x == fixture_b && fx.y == fixture_a && fx.z == index_b && fx.w == index_a) return;
    }
    if (!body_should_collide(fb_, fa_)) return;
    if (joints_prevent_collision(B, body_b, body_a)) return;
    const b2gpu_fixture_rec* fa = &B.fixtures[fixture_a];
    const b2gpu_fixture_rec* fb = &B.fixtures[fixture_b];
    if (!filter_should_collide(*fa, *fb)) return;
    if (!type_pair_primary(fa->shape_type, fb->shape_type)) {
      if (!type_pair_primary(fb->shape_type, fa->shape_type)) {
        B.ws[WS_STATUS] = B2GPU_E_UNSUPPORTED;  // the reference panics (unwrap on an unregistered pair)
        return;
      }
      int t = fixture_a; fixture_a = fixture_b; fixture_b = t;
      t = index_a; index_a = index_b; index_b = t;
    }
    L.cand_fix[j] = make_int4(fixture_a, fixture_b, index_a, index_b);
    L.cand_flag[j] = 1;
  }
};
struct LwCreateK {  // flat over candidates: B2contact::create at contact index cc0 + rank among the survivors
  Batch B;
  Large L;
  int n, cc0;
  B2G_HD void operator()(int j) const {
    if (j >= n || !L.cand_flag[j]) return;
    const int c = cc0 + L.cand_pos[j];
    if (c >= B.NC) return;
    const int4 fx = L.cand_fix[j];
    const b2gpu_fixture_rec* fa = &B.fixtures[fx.x];
    const b2gpu_fixture_rec* fb = &B.fixtures[fx.y];
    B.c_fix[c] = fx;
    B.c_flags[c] = B2GPU_CONTACT_ENABLED;
    B.c_mat[c] = make_float4(sqrtf(fa->friction * fb->friction),
                             fa->restitution > fb->restitution ? fa->restitution : fb->restitution,
                             fa->restitution_threshold < fb->restitution_threshold ? fa->restitution_threshold
                                                                                   : fb->restitution_threshold,
                             0.0f);
    B.c_m0[c] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    B.c_m1[c] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    B.c_m2[c] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    B.c_m3[c] = make_int4(0, 0, 0, 0);
  }
};
struct LwClearMovedK {  // flat over the move buffer
  Batch B;
  int mc;
  B2G_HD void operator()(int i) const {
    if (i >= mc) return;
    const int q = B.move_buf[i];
    if (q != -1) B.n_moved[q] = 0;
  }
};
struct LwPairsFinishK {  // one thread
  Batch B;
  int mc, n_cand, created, status;
  B2G_HD void operator()(int) const {
    int* ws = B.ws;
    ws[WS_CONTACT_COUNT] += created;
    ws[WS_ST_CREATED] += created;
    ws[WS_ST_MOVED] += mc;
    ws[WS_ST_PAIRS] += n_cand;
    ws[WS_MOVE_COUNT] = 0;
    if (status) ws[WS_STATUS] = status;
  }
};
// ------------------------------------------------------------------------------------------
// Gauss-Seidel sweeps of one island per thread, software-pipelined for global memory: the island's chain of
// constraints is sequential by definition, so the time of a sweep is (visits) x (latency of one visit).  The
// generic VelocityK / PositionK pay two dependent memory round trips per visit (record, then the two bodies
// the record names).  Here the record of visit k+2 and the bodies of visit k+1 are requested before visit k is
// solved, and a body shared with the visit just solved is forwarded from registers (the prefetched copy is
// stale by construction), so the arithmetic of visit k overlaps the loads of the next two.  Same functions,
// same order, same bits as VelocityK / PositionK.
// ------------------------------------------------------------------------------------------
struct LwVcRec { float4 q0, q1, q2, q3, q4, q5, q6, q7, q8; };
B2G_HD LwVcRec lw_load_vc(const float4* vc, int k) {
  const float4* r = vc + (size_t)k * VC_Q;
  LwVcRec o;
  o.q0 = r[0]; o.q1 = r[1]; o.q2 = r[2]; o.q3 = r[3]; o.q4 = r[4]; o.q5 = r[5]; o.q6 = r[6]; o.q7 = r[7]; o.q8 = r[8];
  return o;
}
struct LwVcIdxK {  // flat over island contacts: (body A, body B, velocity points, position points | type << 8)
  Batch B;
  Large L;
  int n;
  B2G_HD void operator()(int k) const {
    if (k >= n) return;
    const float4 q8 = B.vc[(size_t)k * VC_Q + 8];
    L.vc_idx[k] = make_int4(f2i(q8.x), f2i(q8.y), f2i(q8.z) & 0xff, f2i(B.pc[(size_t)k * PC_Q + 5].x));
  }
};
template <bool WARM>
B2G_HD void lw_velocity_run(const Batch& B, const Large& L, int first, int n, int sweeps, bool block) {
  const long long total = (long long)n * sweeps;
  int4 ix[4];
  float4 q0[4], q1[4], q2[4], q3[4], q4[4], q5[4], q6[4], q7[4], va[4], vb[4];
  ix[0] = L.vc_idx[first];
  ix[1] = L.vc_idx[first + 1];
  ix[2] = L.vc_idx[first + 2];
  ix[3] = L.vc_idx[first + 3];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int j = 0; j < 2; ++j) {
    const float4* r = B.vc + (size_t)(first + j) * VC_Q;
    q0[j] = r[0]; q1[j] = r[1]; q2[j] = r[2]; q6[j] = r[6]; q7[j] = r[7];
    if (!WARM) { q3[j] = r[3]; q4[j] = r[4]; q5[j] = r[5]; }
    va[j] = B.b_vel[ix[j].x];
    vb[j] = B.b_vel[ix[j].y];
  }
  int k = 0, k2 = 2, k4 = 4 % n;
  float4* scratch = L.scratch4;  // where the "results" of immovable bodies go
  // Forwarding at the point of USE: the results of a visit stay in its register set (va / vb), the bodies it
  // touched in hba / hbb; a visit patches its own inputs from the sets of the two previous visits before the set
  // of visit v-2 is reused for the requests of visit v+2.  (Forwarding INTO the sets still in flight — as
  // LwVelocity4K does — makes every select wait for the load it patches: ncu showed 37 % of the samples there;
  // keeping the history in separate registers instead cost ~25 moves per visit.)  Body -1 matches nothing.
  int hba[4] = {-1, -1, -1, -1}, hbb[4] = {-1, -1, -1, -1};
  long long v = 0;
  for (; v + 4 <= total; v += 4) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < 4; ++j) {
      const int j2 = (j + 2) & 3, j3 = (j + 3) & 3;
      // inputs of this visit: requested two visits ago, so possibly older than the last two visits' results
      const int ba = ix[j].x, bb = ix[j].y, vc_points = ix[j].z;
      float4 a = va[j], b = vb[j];
      a = ba == hba[j3] ? va[j3] : ba == hbb[j3] ? vb[j3] : ba == hba[j2] ? va[j2] : ba == hbb[j2] ? vb[j2] : a;
      b = bb == hba[j3] ? va[j3] : bb == hbb[j3] ? vb[j3] : bb == hba[j2] ? va[j2] : bb == hbb[j2] ? vb[j2] : b;
      hba[j] = ba;
      hbb[j] = bb;
      // requests: the indices of visit v+4 take this visit's slot (they are needed two visits from now, to request
      // the bodies of visit v+4: an index that is still in flight when its bodies are requested stalls the warp —
      // ncu, second build); the record and the bodies of visit v+2 take the set of visit v-2
      ix[j] = L.vc_idx[first + k4];
      {
        const float4* r = B.vc + (size_t)(first + k2) * VC_Q;
        q0[j2] = r[0]; q1[j2] = r[1]; q2[j2] = r[2]; q6[j2] = r[6]; q7[j2] = r[7];
        if (!WARM) { q3[j2] = r[3]; q4[j2] = r[4]; q5[j2] = r[5]; }
        va[j2] = B.b_vel[ix[j2].x];
        vb[j2] = B.b_vel[ix[j2].y];
      }
      VelState s;
      s.v_a = v2(a.x, a.y); s.w_a = a.z;
      s.v_b = v2(b.x, b.y); s.w_b = b.z;
      if (WARM) {
        warm_start_one(s, q0[j], q1[j], q2[j], q6[j], q7[j], vc_points);
      } else {
        solve_velocity_one(s, q0[j], q1[j], q2[j], q3[j], q4[j], q5[j], q6[j], q7[j], vc_points, block);
        B.vc[(size_t)(first + k) * VC_Q + 6] = q6[j];
      }
      // a static / kinematic body may sit in several islands: its velocity never changes — the result of the
      // arithmetic on it (inverse mass 0) is its old value, which is stored to a scratch slot instead
      const bool mov_a = q7[j].x != 0.0f || q7[j].y != 0.0f, mov_b = q7[j].z != 0.0f || q7[j].w != 0.0f;
      va[j] = mov_a ? make_float4(s.v_a.x, s.v_a.y, s.w_a, 0.0f) : a;
      vb[j] = mov_b ? make_float4(s.v_b.x, s.v_b.y, s.w_b, 0.0f) : b;
      *(mov_a ? &B.b_vel[ba] : scratch) = va[j];
      *(mov_b ? &B.b_vel[bb] : scratch + 1) = vb[j];
      if (++k == n) k = 0;
      if (++k2 == n) k2 = 0;
      if (++k4 == n) k4 = 0;
    }
  }
  for (; v < total; ++v) {  // at most three visits left: every store above went to memory, plain loads are current
    const int kk = first + k;
    const int4 ixx = L.vc_idx[kk];
    const float4 a = B.b_vel[ixx.x], b = B.b_vel[ixx.y];
    VelState s;
    s.v_a = v2(a.x, a.y); s.w_a = a.z;
    s.v_b = v2(b.x, b.y); s.w_b = b.z;
    LwVcRec r = lw_load_vc(B.vc, kk);
    if (WARM) {
      warm_start_one(s, r.q0, r.q1, r.q2, r.q6, r.q7, ixx.z);
    } else {
      solve_velocity_one(s, r.q0, r.q1, r.q2, r.q3, r.q4, r.q5, r.q6, r.q7, ixx.z, block);
      B.vc[(size_t)kk * VC_Q + 6] = r.q6;
    }
    if (r.q7.x != 0.0f || r.q7.y != 0.0f) B.b_vel[ixx.x] = make_float4(s.v_a.x, s.v_a.y, s.w_a, 0.0f);
    if (r.q7.z != 0.0f || r.q7.w != 0.0f) B.b_vel[ixx.y] = make_float4(s.v_b.x, s.v_b.y, s.w_b, 0.0f);
    if (++k == n) k = 0;
  }
}
// `sweeps` passes of one kind (warm start, or velocity iterations) over the contact constraints [first, first + n) of an
// island in the register-pipelined form; islands under four contacts take the plain loop.
template <bool WARM>
B2G_HD void lw_contact_sweeps_small(const Batch& B, const Large& L, int first, int n, int sweeps, bool block) {
  if (n <= 0 || sweeps <= 0) return;
  if (n < 4) {  // the pipeline assumes a constraint is not in flight twice
    for (int it = 0; it < sweeps; ++it) {
      for (int k = first; k < first + n; ++k) {
        const int4 ix = L.vc_idx[k];
        if (ix.z == 0) continue;
        const float4 va = B.b_vel[ix.x], vb = B.b_vel[ix.y];
        VelState s;
        s.v_a = v2(va.x, va.y); s.w_a = va.z;
        s.v_b = v2(vb.x, vb.y); s.w_b = vb.z;
        LwVcRec r = lw_load_vc(B.vc, k);
        if (WARM) {
          warm_start_one(s, r.q0, r.q1, r.q2, r.q6, r.q7, ix.z);
        } else {
          solve_velocity_one(s, r.q0, r.q1, r.q2, r.q3, r.q4, r.q5, r.q6, r.q7, ix.z, block);
          B.vc[(size_t)k * VC_Q + 6] = r.q6;
        }
        if (r.q7.x != 0.0f || r.q7.y != 0.0f) B.b_vel[ix.x] = make_float4(s.v_a.x, s.v_a.y, s.w_a, 0.0f);
        if (r.q7.z != 0.0f || r.q7.w != 0.0f) B.b_vel[ix.y] = make_float4(s.v_b.x, s.v_b.y, s.w_b, 0.0f);
      }
    }
    return;
  }
  lw_velocity_run<WARM>(B, L, first, n, sweeps, block);
}
// Joint rows of island `isl` (large-world mode: LB = 1, body state in the plain arrays)
B2G_HD void lw_joints_init(const Batch& B, int isl, bool warm, const StepParams& sp) {
  if (B.NJ == 0) return;
  const int2 jr = B.isl_jrange[isl];
  WIdx x;
  x.wb = 0; x.wl = 0; x.LB = 1;
  const BodyStateGlobal st = {B, x};
  const float dt_ratio = i2f(B.ws[WS_INV_DT0]) * sp.dt;
  for (int q = jr.x; q < jr.y; ++q) joint_init_velocity(B, x, st, B.isl_joint[q], warm, dt_ratio, sp.dt);
}
B2G_HD void lw_joints_velocity(const Batch& B, int isl, const StepParams& sp) {
  if (B.NJ == 0) return;
  const int2 jr = B.isl_jrange[isl];
  WIdx x;
  x.wb = 0; x.wl = 0; x.LB = 1;
  const BodyStateGlobal st = {B, x};
  for (int q = jr.x; q < jr.y; ++q) joint_solve_velocity(B, x, st, B.isl_joint[q], sp.dt, sp.inv_dt);
}
B2G_HD bool lw_joints_position(const Batch& B, int isl) {
  if (B.NJ == 0) return true;
  const int2 jr = B.isl_jrange[isl];
  WIdx x;
  x.wb = 0; x.wl = 0; x.LB = 1;
  const BodyStateGlobal st = {B, x};
  bool ok = true;
  for (int q = jr.x; q < jr.y; ++q) {
    const bool joint_okay = joint_solve_position(B, x, st, B.isl_joint[q]);
    ok = ok && joint_okay;
  }
  return ok;
}
B2G_HD bool lw_island_has_joints(const Batch& B, int isl) {
  if (B.NJ == 0) return false;
  const int2 jr = B.isl_jrange[isl];
  return jr.x != jr.y;
}

// ------------------------------------------------------------------------------------------
// Velocity sweeps through a cp.async shared-memory ring (experiment, B2GPU_LW_VELOCITY=7).  Prefetching into
// registers cannot take the global-memory round trip off the chain (six scoreboard slots alias the loads in
// flight: profiles/r01_large_world.md), so here — as in the batch kernels — the constraint records are copied by
// cp.async into a per-thread ring eight visits ahead and the two bodies two visits ahead (their indices come
// from the record already in the ring); a visit reads its record and bodies with shared-memory loads and patches
// the bodies from the results of the last three visits, which covers every store the asynchronous copy may or
// may not have seen.  Same functions, same order, same bits.  In the host simulator a copy is a plain assignment.
// ------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define LW_CP16(dst, src) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory")
#define LW_CP_COMMIT() asm volatile("cp.async.commit_group;\n" ::: "memory")
#define LW_CP_WAIT1() asm volatile("cp.async.wait_group 1;\n" ::: "memory")
#define LW_CP_WAIT0() asm volatile("cp.async.wait_group 0;\n" ::: "memory")
#else
#define LW_CP16(dst, src) (*(float4*)(dst) = *(const float4*)(src))
#define LW_CP_COMMIT()
#define LW_CP_WAIT1()
#define LW_CP_WAIT0()
#endif
enum { LW_RING = 8, LW_BRING = 4 };
struct LwHist { int a, b; float ax, ay, az, bx, by, bz; };  // bodies a visit touched and what it left in them
// One visit of the ring form.  `out` still holds the results of visit v-3 when the visit starts and receives this
// visit's results; h1 / h2 are visits v-1 / v-2.  The three history slots rotate through a loop unrolled by three,
// so nothing is moved (ncu, fourth capture: the shuffle of a three-level history was ≈30 of ≈70 moves per visit).
template <bool WARM>
B2G_HD void lw_ring_visit(const Batch& B, int first, int n, bool block, float4* ring, float4* bod, int stride, float4* scratch,
                          LwHist& out, const LwHist& h1, const LwHist& h2, long long v, int& k, int& kf) {
  const int slot = (int)(v & (LW_RING - 1)), bs = (int)(v & (LW_BRING - 1));
  LW_CP_WAIT1();
  const float4* rs = ring + (slot * VC_Q) * stride;
  const float4 q0 = rs[0], q1 = rs[stride], q2 = rs[2 * stride], q3 = rs[3 * stride], q4 = rs[4 * stride], q5 = rs[5 * stride];
  float4 q6 = rs[6 * stride];
  const float4 q7 = rs[7 * stride], q8 = rs[8 * stride];
  const float4 la = bod[(bs * 2) * stride], lb = bod[(bs * 2 + 1) * stride];
  const int ba = f2i(q8.x), bb = f2i(q8.y), vc_points = f2i(q8.z) & 0xff;
  float ax = la.x, ay = la.y, az = la.z, bx = lb.x, by = lb.y, bz = lb.z;
  if (ba == h1.a) { ax = h1.ax; ay = h1.ay; az = h1.az; } else if (ba == h1.b) { ax = h1.bx; ay = h1.by; az = h1.bz; }
  else if (ba == h2.a) { ax = h2.ax; ay = h2.ay; az = h2.az; } else if (ba == h2.b) { ax = h2.bx; ay = h2.by; az = h2.bz; }
  else if (ba == out.a) { ax = out.ax; ay = out.ay; az = out.az; } else if (ba == out.b) { ax = out.bx; ay = out.by; az = out.bz; }
  if (bb == h1.a) { bx = h1.ax; by = h1.ay; bz = h1.az; } else if (bb == h1.b) { bx = h1.bx; by = h1.by; bz = h1.bz; }
  else if (bb == h2.a) { bx = h2.ax; by = h2.ay; bz = h2.az; } else if (bb == h2.b) { bx = h2.bx; by = h2.by; bz = h2.bz; }
  else if (bb == out.a) { bx = out.ax; by = out.ay; bz = out.az; } else if (bb == out.b) { bx = out.bx; by = out.by; bz = out.bz; }
  {  // refill this slot with the record eight visits ahead; request the bodies of the visit two ahead
    const float4* r = B.vc + (size_t)(first + kf) * VC_Q;
    for (int q = 0; q < VC_Q; ++q) LW_CP16(ring + (slot * VC_Q + q) * stride, r + q);
    if (++kf == n) kf = 0;
    const int s2 = (int)((v + 2) & (LW_RING - 1)), b2 = (int)((v + 2) & (LW_BRING - 1));
    const float4 n8 = ring[(s2 * VC_Q + 8) * stride];
    LW_CP16(bod + (b2 * 2) * stride, &B.b_vel[f2i(n8.x)]);
    LW_CP16(bod + (b2 * 2 + 1) * stride, &B.b_vel[f2i(n8.y)]);
    LW_CP_COMMIT();
  }
  VelState s;
  s.v_a = v2(ax, ay); s.w_a = az;
  s.v_b = v2(bx, by); s.w_b = bz;
  if (WARM) {
    warm_start_one(s, q0, q1, q2, q6, q7, vc_points);
  } else {
    solve_velocity_one(s, q0, q1, q2, q3, q4, q5, q6, q7, vc_points, block);
    B.vc[(size_t)(first + k) * VC_Q + 6] = q6;
  }
  // a static / kinematic body may sit in several islands: its velocity never changes — the result of the arithmetic
  // on it (inverse mass 0) is its old value, which is stored to a scratch slot instead
  const bool mov_a = q7.x != 0.0f || q7.y != 0.0f, mov_b = q7.z != 0.0f || q7.w != 0.0f;
  out.a = ba; out.b = bb;
  out.ax = mov_a ? s.v_a.x : ax; out.ay = mov_a ? s.v_a.y : ay; out.az = mov_a ? s.w_a : az;
  out.bx = mov_b ? s.v_b.x : bx; out.by = mov_b ? s.v_b.y : by; out.bz = mov_b ? s.w_b : bz;
  *(mov_a ? &B.b_vel[ba] : scratch) = make_float4(out.ax, out.ay, out.az, 0.0f);
  *(mov_b ? &B.b_vel[bb] : scratch + 1) = make_float4(out.bx, out.by, out.bz, 0.0f);
  if (++k == n) k = 0;
}
template <bool WARM>
B2G_HD void lw_velocity_ring(const Batch& B, int first, int n, int sweeps, bool block, float4* ring, float4* bod, int stride,
                             float4* scratch) {
  const long long total = (long long)n * sweeps;
  int kf = 0;  // constraint whose record is fetched next
  for (int p = 0; p < LW_RING; ++p) {
    const float4* r = B.vc + (size_t)(first + kf) * VC_Q;
    for (int q = 0; q < VC_Q; ++q) LW_CP16(ring + (p * VC_Q + q) * stride, r + q);
    if (++kf == n) kf = 0;
  }
  LW_CP_COMMIT();
  LW_CP_WAIT0();
  for (int p = 0; p < 2; ++p) {
    const float4 q8 = ring[(p * VC_Q + 8) * stride];
    LW_CP16(bod + (p * 2) * stride, &B.b_vel[f2i(q8.x)]);
    LW_CP16(bod + (p * 2 + 1) * stride, &B.b_vel[f2i(q8.y)]);
  }
  LW_CP_COMMIT();
  LW_CP_WAIT0();
  LW_CP_COMMIT();  // an empty group, so that "all but the newest group" in a visit always means "two visits back"
  LwHist H0, H1, H2;
  H0.a = H0.b = H1.a = H1.b = H2.a = H2.b = -1;  // body -1 matches nothing
  H0.ax = H0.ay = H0.az = H0.bx = H0.by = H0.bz = 0.0f;
  H1 = H0;
  H2 = H0;
  H1.a = H1.b = H2.a = H2.b = -1;
  int k = 0;
  long long v = 0;
  for (; v + 3 <= total; v += 3) {  // before a group: H2 = visit v-1, H1 = v-2, H0 = v-3
    lw_ring_visit<WARM>(B, first, n, block, ring, bod, stride, scratch, H0, H2, H1, v, k, kf);
    lw_ring_visit<WARM>(B, first, n, block, ring, bod, stride, scratch, H1, H0, H2, v + 1, k, kf);
    lw_ring_visit<WARM>(B, first, n, block, ring, bod, stride, scratch, H2, H1, H0, v + 2, k, kf);
  }
  if (v < total) { lw_ring_visit<WARM>(B, first, n, block, ring, bod, stride, scratch, H0, H2, H1, v, k, kf); ++v; }
  if (v < total) lw_ring_visit<WARM>(B, first, n, block, ring, bod, stride, scratch, H1, H0, H2, v, k, kf);
  LW_CP_WAIT0();
}
struct LwVelocity7K {
  Batch B;
  Large L;
  StepParams sp;
  int n_islands;
  B2G_HD void operator()(int isl) const {
#if defined(__CUDA_ARCH__)
    __shared__ float4 ring_s[LW_RING * VC_Q * 32];
    __shared__ float4 bod_s[LW_BRING * 2 * 32];
    float4* ring = ring_s + (threadIdx.x & 31);
    float4* bod = bod_s + (threadIdx.x & 31);
    const int stride = 32;
#else
    float4 ring_s[LW_RING * VC_Q], bod_s[LW_BRING * 2];
    float4* ring = ring_s;
    float4* bod = bod_s;
    const int stride = 1;
#endif
    if (isl >= n_islands || L.lv_isl_giant[isl]) return;  // a giant island is swept by LwLevelVelocityK
    const int4 rg = B.isl_range[isl];
    const bool joints = lw_island_has_joints(B, isl);
    if (rg.z == rg.w && !joints) return;
    const bool warm = (B.ws[WS_FLAGS] & B2GPU_WORLD_WARM_STARTING) != 0;
    const bool block = (B.ws[WS_FLAGS] & B2GPU_WORLD_BLOCK_SOLVE) != 0;
    const int first = rg.z, n = rg.w - rg.z;
    const bool small = n < 2 * LW_RING;  // a record must not be in the ring while its impulses are rewritten: small islands take the register form
    if (!joints) {
      if (small) {
        if (warm) lw_contact_sweeps_small<true>(B, L, first, n, 1, block);
        lw_contact_sweeps_small<false>(B, L, first, n, sp.velocity_iterations, block);
      } else {
        if (warm) lw_velocity_ring<true>(B, first, n, 1, block, ring, bod, stride, L.scratch4 + 8);
        if (sp.velocity_iterations > 0) lw_velocity_ring<false>(B, first, n, sp.velocity_iterations, block, ring, bod, stride, L.scratch4 + 8);
      }
      return;
    }
    // an island with joints (b2_island_private.rs:193-215): contact warm start, every joint's init_velocity_constraints,
    // then per iteration the joint rows before the contact rows
    if (warm) {
      if (small) lw_contact_sweeps_small<true>(B, L, first, n, 1, block);
      else lw_velocity_ring<true>(B, first, n, 1, block, ring, bod, stride, L.scratch4 + 8);
    }
    lw_joints_init(B, isl, warm, sp);
    for (int it = 0; it < sp.velocity_iterations; ++it) {
      lw_joints_velocity(B, isl, sp);
      if (small) lw_contact_sweeps_small<false>(B, L, first, n, 1, block);
      else lw_velocity_ring<false>(B, first, n, 1, block, ring, bod, stride, L.scratch4 + 8);
    }
  }
};

struct LwPcRec { float4 p0, p1, p2, p3, p4, p5; };
B2G_HD LwPcRec lw_load_pc(const float4* pc, int k) {
  const float4* r = pc + (size_t)k * PC_Q;
  LwPcRec o;
  o.p0 = r[0]; o.p1 = r[1]; o.p2 = r[2]; o.p3 = r[3]; o.p4 = r[4]; o.p5 = r[5];
  return o;
}
struct LwPosSet { float4 p0, p1, p2, p3, p4, pa, pb, ra, rb; int4 ix; int hba, hbb; };
B2G_HD void lw_pos_request(const Batch& B, const Large& L, int kk, LwPosSet& t) {  // t.ix was loaded a visit earlier
  const float4* r = B.pc + (size_t)kk * PC_Q;
  t.p0 = r[0]; t.p1 = r[1]; t.p2 = r[2]; t.p3 = r[3]; t.p4 = r[4];
  t.pa = B.b_pos[t.ix.x]; t.ra = B.b_rot[t.ix.x];
  t.pb = B.b_pos[t.ix.y]; t.rb = B.b_rot[t.ix.y];
}
// One visit on set t.  The other set o holds the previous visit's results (o.pa .. o.rb of bodies o.hba / o.hbb):
// they are forwarded into this visit's inputs first; then o is reused for the request of the next visit
// (constraint `next_k`, whose indices o.ix were requested a visit ago), and t.ix is requested for constraint
// `next_ix` (this set's next use, two visits from now); then the arithmetic; the results stay in t.
B2G_HD float lw_pos_visit(const Batch& B, const Large& L, float4* scratch, LwPosSet& t, LwPosSet& o, int next_k, int next_ix,
                          float min_separation) {
  const int ba = t.ix.x, bb = t.ix.y, packed = t.ix.w;
  float4 pa = t.pa, ra = t.ra, pb = t.pb, rb = t.rb;
  if (ba == o.hba) { pa = o.pa; ra = o.ra; } else if (ba == o.hbb) { pa = o.pb; ra = o.rb; }
  if (bb == o.hba) { pb = o.pa; rb = o.ra; } else if (bb == o.hbb) { pb = o.pb; rb = o.rb; }
  const float4 p0 = t.p0, p1 = t.p1, p2 = t.p2, p3 = t.p3, p4 = t.p4;
  t.hba = ba;
  t.hbb = bb;
  t.ix = L.vc_idx[next_ix];
  if (next_k >= 0) lw_pos_request(B, L, next_k, o);
  PosState s;
  s.c_a = v2(pa.x, pa.y); s.a_a = pa.z; s.q_a.s = ra.x; s.q_a.c = ra.y;
  s.c_b = v2(pb.x, pb.y); s.a_b = pb.z; s.q_b.s = rb.x; s.q_b.c = rb.y;
  min_separation = solve_position_one(s, p0, p1, p2, p3, (packed >> 8) & 0xff, packed & 0xff, p4.x, p4.y, min_separation);
  const bool mov_a = p0.x != 0.0f || p0.y != 0.0f, mov_b = p0.z != 0.0f || p0.w != 0.0f;
  if (mov_a) { pa.x = s.c_a.x; pa.y = s.c_a.y; pa.z = s.a_a; ra.x = s.q_a.s; ra.y = s.q_a.c; }
  if (mov_b) { pb.x = s.c_b.x; pb.y = s.c_b.y; pb.z = s.a_b; rb.x = s.q_b.s; rb.y = s.q_b.c; }
  *(mov_a ? &B.b_pos[ba] : scratch) = pa;
  *(mov_a ? &B.b_rot[ba] : scratch + 1) = ra;
  *(mov_b ? &B.b_pos[bb] : scratch + 2) = pb;
  *(mov_b ? &B.b_rot[bb] : scratch + 3) = rb;
  t.pa = pa; t.ra = ra; t.pb = pb; t.rb = rb;
  return min_separation;
}
struct LwPosition6K {
  Batch B;
  Large L;
  StepParams sp;
  int n_islands;
  B2G_HD void operator()(int isl) const {
    if (isl >= n_islands || L.lv_isl_giant[isl]) return;  // a giant island is swept by LwLevelPositionK
    const int4 rg = B.isl_range[isl];
    const bool joints = lw_island_has_joints(B, isl);
    if (rg.z == rg.w && !joints) return;
    const int first = rg.z, n = rg.w - rg.z;
    float4* scratch = L.scratch4 + 4;  // where the "results" of immovable bodies go
    for (int it = 0; it < sp.position_iterations; ++it) {
      float min_separation = 0.0f;
      if (n > 0) {
        LwPosSet s0, s1;
        const int last = first + n - 1;
        s0.ix = L.vc_idx[first];
        s1.ix = L.vc_idx[first + 1 <= last ? first + 1 : last];
        s1.hba = -1; s1.hbb = -1;
        s1.pa = s1.ra = s1.pb = s1.rb = make_float4(0, 0, 0, 0);
        lw_pos_request(B, L, first, s0);
        int k = 0;
        for (; k + 2 <= n; k += 2) {
          const int c1 = first + k + 1, c2 = first + k + 2 <= last ? first + k + 2 : last, c3 = first + k + 3 <= last ? first + k + 3 : last;
          min_separation = lw_pos_visit(B, L, scratch, s0, s1, c1, c2, min_separation);
          min_separation = lw_pos_visit(B, L, scratch, s1, s0, c2, c3, min_separation);
        }
        if (k < n) min_separation = lw_pos_visit(B, L, scratch, s0, s1, -1, last, min_separation);
      }
      const bool joints_okay = !joints || lw_joints_position(B, isl);  // :262-266, after the contact rows
      if (min_separation >= -3.0f * B2G_LINEAR_SLOP && joints_okay) {  // b2_island_private.rs:257-274
        B.isl_flags[isl] |= 1;
        break;
      }
    }
  }
};

// SleepK for a large world (b2_island_private.rs:283-312): 32 threads per island instead of one — the settled 100k pile is
// ONE island of 100k bodies.  phase 0 (flat over islands): minimum = max float; phase 1 (flat over 32 x islands): minimum of
// the bodies' sleep times (non-negative floats order like their bit patterns: atomicMin on the bits); phase 2 (flat over
// 32 x islands): set_awake(false) for every body of an island that is solved and has rested long enough.
struct LwSleepK {
  Batch B;
  Large L;
  int n_islands, phase;
  B2G_HD void operator()(int t) const {
    const int* ws = B.ws;
    if (!(ws[WS_FLAGS] & B2GPU_WORLD_ALLOW_SLEEP)) return;
    if (phase == 0) {
      if (t < n_islands) L.sleep_min[t] = f2i(B2G_MAX_FLOAT);
      return;
    }
    const int isl = t >> 5, lane = t & 31;
    if (isl >= n_islands) return;
    const int4 rg = B.isl_range[isl];
    if (phase == 1) {
      float m = B2G_MAX_FLOAT;
      for (int k = rg.x + lane; k < rg.y; k += 32) {
        const int bi = B.isl_body[k];
        if (body_type(B.b_flags[bi]) == B2GPU_STATIC_BODY) continue;
        m = fmin_sel(m, B.b_pos[bi].w);
      }
      if (m < B2G_MAX_FLOAT) B2G_ATOMIC_MIN(&L.sleep_min[isl], f2i(m));
      return;
    }
    if (!(i2f(L.sleep_min[isl]) >= B2G_TIME_TO_SLEEP && (B.isl_flags[isl] & 1))) return;
    for (int k = rg.x + lane; k < rg.y; k += 32) {  // set_awake(false), src/b2_body.rs:783-801
      const int bi = B.isl_body[k];
      const int bf = B.b_flags[bi];
      if (body_type(bf) == B2GPU_STATIC_BODY) continue;
      B.b_flags[bi] = bf & ~B2GPU_BODY_AWAKE;
      B.b_pos[bi].w = 0.0f;
      B.b_vel[bi] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      float4 fo = B.b_force[bi];
      fo.x = 0.0f; fo.y = 0.0f; fo.z = 0.0f;
      B.b_force[bi] = fo;
    }
    if (lane == 0) B.ws[WS_TOPO_DIRTY] = 1;
  }
};

struct LwStatsK {  // touching / awake counters on demand: phase 0 one thread (reset), phase 1 flat over max(cc, NB)
  Batch B;
  int cc, phase;
  B2G_HD void operator()(int t) const {
    int* ws = B.ws;
    if (phase == 0) {
      ws[WS_ST_TOUCHING] = 0;
      ws[WS_ST_AWAKE] = 0;
      ws[WS_ST_CONTACTS] = cc;
      return;
    }
    if (t < cc && (B.c_flags[t] & B2GPU_CONTACT_TOUCHING)) counter_inc(&ws[WS_ST_TOUCHING]);
    if (t < B.NB && (B.b_flags[t] & B2GPU_BODY_AWAKE)) counter_inc(&ws[WS_ST_AWAKE]);
  }
};
struct LwStepEndK {  // one thread: the tail of TreePairsK
  Batch B;
  StepParams sp;
  B2G_HD void operator()(int) const {
    int* ws = B.ws;
    if (sp.dt > 0.0f) ws[WS_INV_DT0] = f2i(sp.inv_dt);
    ws[WS_ST_CONTACTS] = ws[WS_CONTACT_COUNT];
  }
};

}  // namespace b2g
