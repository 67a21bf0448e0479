// b2g_island_smem.cuh — island construction for batches (LB == 32) with the world graph in shared memory.
//
// Reference: src/private/dynamics/b2_world.rs:376-507 — seeds in body-list order (newest first), an
// explicit LIFO stack, each body's contact-edge list newest first.  The traversal is a lexicographic
// DFS, sequential inside a world; the lanes of a warp are 32 worlds.  From global memory every edge
// visit is a chain of dependent ~600-cycle loads (contact flags -> fixtures -> other body's flags);
// here each world's graph is first compacted into shared memory — only contacts that can join an
// island (enabled, touching, non-sensor), 16-bit indices, lane-minor layout so that staging reads are
// coalesced — and the DFS then runs at shared-memory latency.  Worlds whose eligible contacts exceed
// the shared-memory capacity fall back to SerialAK::islands_global.
#pragma once
#include "b2g_island_layout.h"
#include "b2g_step.h"

namespace b2g {

__global__ void __launch_bounds__(32) island_smem_kernel(const SerialAK K, const IslandSmemLayout L) {
  extern __shared__ unsigned char smem_raw[];
  const Batch& B = K.B;
  uint32_t* enext = (uint32_t*)(smem_raw + L.off_enext);   // [ECAP][32] next edge of body A | body B << 16
  uint32_t* ebody = (uint32_t*)(smem_raw + L.off_ebody);   // [ECAP][32] body A | body B << 16
  uint16_t* eorig = (uint16_t*)(smem_raw + L.off_eorig);   // [ECAP][32] contact index
  uint16_t* chead = (uint16_t*)(smem_raw + L.off_chead);   // [NB][32] newest eligible edge (2e + side) or 0xffff
  uint16_t* stack = (uint16_t*)(smem_raw + L.off_stack);   // [NB][32]
  uint8_t* eisl = (uint8_t*)(smem_raw + L.off_eisl);       // [ECAP][32] 1 = already in an island
  uint8_t* bflag = (uint8_t*)(smem_raw + L.off_bflag);     // [NB][32] bit0 ISLAND bit1 AWAKE bit2 ENABLED bit3 static
  const int lane = threadIdx.x;
  const int wb = blockIdx.x;
  const int w = wb * 32 + lane;
  const bool live = w < B.n_worlds;
  WIdx x;
  x.wb = wb; x.wl = lane; x.LB = 32;
  Ws ws = ws_of(B, x);
  bool need = false;
  if (live) need = K.prologue(x, ws);
  if (!__any_sync(0xffffffffu, need)) return;
  const int cc = need ? ws[WS_CONTACT_COUNT] : 0;
  const int NB = B.NB;
  // ---- stage: body flags, eligible contacts with per-body edge lists (ascending contact index +
  //      push_front == newest first)
  if (need) {
    for (int b = 0; b < NB; ++b) {
      const int f = B.b_flags[x.at(NB, b)];
      bflag[b * 32 + lane] = (uint8_t)(((f & B2GPU_BODY_AWAKE) ? 2 : 0) | ((f & B2GPU_BODY_ENABLED) ? 4 : 0) |
                                       (body_type(f) == B2GPU_STATIC_BODY ? 8 : 0));
      chead[b * 32 + lane] = 0xffffu;
    }
  }
  int ne = 0;
  bool overflow = false;
  if (need) {
#pragma unroll 4
    for (int c = 0; c < cc; ++c) {
      const int ci = x.at(B.NC, c);
      const int cf = B.c_flags[ci];
      const int4 fx = B.c_fix[ci];
      const b2gpu_fixture_rec& fa = B.fixtures[fx.x];
      const b2gpu_fixture_rec& fb = B.fixtures[fx.y];
      const bool eligible = (cf & B2GPU_CONTACT_ENABLED) && (cf & B2GPU_CONTACT_TOUCHING) && !fa.is_sensor && !fb.is_sensor;
      if (!eligible) continue;
      if (ne >= L.ECAP) { overflow = true; break; }
      const int ba = fa.body, bb = fb.body;
      const uint32_t na = chead[ba * 32 + lane];
      chead[ba * 32 + lane] = (uint16_t)(2 * ne);
      const uint32_t nb_ = chead[bb * 32 + lane];   // after the A push: a self pair is impossible (add_pair rejects it)
      chead[bb * 32 + lane] = (uint16_t)(2 * ne + 1);
      enext[ne * 32 + lane] = na | (nb_ << 16);
      ebody[ne * 32 + lane] = (uint32_t)ba | ((uint32_t)bb << 16);
      eorig[ne * 32 + lane] = (uint16_t)c;
      eisl[ne * 32 + lane] = 0;
      ++ne;
    }
  }
  if (overflow) {  // graph does not fit: this world takes the global-memory path
    K.islands_global(x, ws);
    need = false;
  }
  if (!need) return;
  // ---- DFS
  int nisl = 0, nbod = 0, ncon = 0;
  bool dirty_next = false;
  for (int seed = NB - 1; seed >= 0; --seed) {
    const uint32_t sf = bflag[seed * 32 + lane];
    if ((sf & 1) || !(sf & 2) || !(sf & 4) || (sf & 8)) continue;
    const int body_first = nbod, contact_first = ncon;
    int sp_ = 0;
    stack[(sp_++) * 32 + lane] = (uint16_t)seed;
    bflag[seed * 32 + lane] = (uint8_t)(sf | 1);
    while (sp_ > 0) {
      const int b = stack[(--sp_) * 32 + lane];
      if (nbod >= B.NIB) { ws[WS_STATUS] = B2GPU_E_CAPACITY; break; }
      B.isl_body[x.at(B.NIB, nbod++)] = b;
      const uint32_t bf = bflag[b * 32 + lane];
      if (bf & 8) continue;
      if (!(bf & 2)) dirty_next = true;
      bflag[b * 32 + lane] = (uint8_t)(bf | 2);
      for (uint32_t e = chead[b * 32 + lane]; e != 0xffffu;) {
        const int ei = (int)(e >> 1), side = (int)(e & 1);
        const uint32_t nx = enext[ei * 32 + lane];
        e = side ? (nx >> 16) : (nx & 0xffffu);
        if (eisl[ei * 32 + lane]) continue;
        eisl[ei * 32 + lane] = 1;
        B.isl_contact[x.at(B.NC, ncon)] = eorig[ei * 32 + lane];
        B.c_isl[x.at(B.NC, ncon)] = nisl;
        ++ncon;
        const uint32_t bod = ebody[ei * 32 + lane];
        const int other = side ? (int)(bod & 0xffffu) : (int)(bod >> 16);
        const uint32_t of = bflag[other * 32 + lane];
        if (of & 1) continue;
        stack[(sp_++) * 32 + lane] = (uint16_t)other;
        bflag[other * 32 + lane] = (uint8_t)(of | 1);
      }
    }
    B.isl_range[x.at(NB, nisl)] = make_int4(body_first, nbod, contact_first, ncon);
    ++nisl;
    for (int k = body_first; k < nbod; ++k) {  // static bodies may join other islands
      const int b = B.isl_body[x.at(B.NIB, k)];
      const uint32_t bf = bflag[b * 32 + lane];
      if (bf & 8) bflag[b * 32 + lane] = (uint8_t)(bf & ~1u);
    }
  }
  // ---- write the ISLAND / AWAKE flags back
  for (int b = 0; b < NB; ++b) {
    const int bi = x.at(NB, b);
    const uint32_t bf = bflag[b * 32 + lane];
    int f = B.b_flags[bi] & ~(B2GPU_BODY_ISLAND | B2GPU_BODY_AWAKE);
    f |= ((bf & 1) ? B2GPU_BODY_ISLAND : 0) | ((bf & 2) ? B2GPU_BODY_AWAKE : 0);
    B.b_flags[bi] = f;
  }
  {
    int e = 0;
#pragma unroll 4
    for (int c = 0; c < cc; ++c) {
      const int ci = x.at(B.NC, c);
      int cf = B.c_flags[ci] & ~B2GPU_CONTACT_ISLAND;
      if (e < ne && eorig[e * 32 + lane] == (uint16_t)c) {
        if (eisl[e * 32 + lane]) cf |= B2GPU_CONTACT_ISLAND;
        ++e;
      }
      B.c_flags[ci] = cf;
    }
  }
  ws[WS_ISL_COUNT] = nisl;
  ws[WS_ISL_BODIES] = nbod;
  ws[WS_ISL_CONTACTS] = ncon;
  ws[WS_ST_ISLANDS] = nisl;
  ws[WS_ST_ISL_BODIES] = nbod;
  ws[WS_ST_ISL_CONTACTS] = ncon;
  ws[WS_TOPO_DIRTY] = dirty_next ? 1 : 0;
}

}  // namespace b2g
