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
  uint16_t* smark = (uint16_t*)(smem_raw + L.off_smark);   // [NB][32] static bodies: 1 + island that holds them
  uint8_t* eisl = (uint8_t*)(smem_raw + L.off_eisl);       // [ECAP][32] 1 = already in an island
  uint8_t* bflag = (uint8_t*)(smem_raw + L.off_bflag);     // [NB][32] bit0 ISLAND bit1 AWAKE bit2 ENABLED bit3 static
  uint32_t* fxb = (uint32_t*)(smem_raw + L.off_fxb);       // [fxb_count] fixture -> body | sensor << 31
  uint16_t* korder = (uint16_t*)(smem_raw + L.off_korder); // [ECAP][32] island contact slot k -> eligible edge
  const int lane = threadIdx.x;
  const int wb = blockIdx.x + B.wb_first;
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
  for (int f = lane; f < L.fxb_count; f += 32)
    fxb[f] = (uint32_t)B.fixtures[f].body | (B.fixtures[f].is_sensor ? 0x80000000u : 0u);
  __syncwarp();
  auto fix_word = [&](int f) -> uint32_t {
    if (L.fxb_count) return fxb[f];
    return (uint32_t)B.fixtures[f].body | (B.fixtures[f].is_sensor ? 0x80000000u : 0u);
  };
  // ---- stage: body flags, eligible contacts with per-body edge lists (ascending contact index +
  //      push_front == newest first).  Global reads are issued eight at a time.
  if (need) {
    for (int b0 = 0; b0 < NB; b0 += 8) {
      int f[8];
      float4 ms[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        f[j] = b0 + j < NB ? B.b_flags[x.at(NB, b0 + j)] : 0;
        ms[j] = b0 + j < NB ? B.b_mass[x.at(NB, b0 + j)] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (b0 + j >= NB) break;
        // bit4: the solver can change this body's velocity/position (non-zero inverse mass or inertia)
        bflag[(b0 + j) * 32 + lane] = (uint8_t)(((f[j] & B2GPU_BODY_AWAKE) ? 2 : 0) | ((f[j] & B2GPU_BODY_ENABLED) ? 4 : 0) |
                                                (body_type(f[j]) == B2GPU_STATIC_BODY ? 8 : 0) |
                                                ((ms[j].x != 0.0f || ms[j].y != 0.0f) ? 16 : 0));
        chead[(b0 + j) * 32 + lane] = 0xffffu;
        smark[(b0 + j) * 32 + lane] = 0;
      }
    }
  }
  int ne = 0;
  bool overflow = false;
  if (need) {
    for (int c0 = 0; c0 < cc && !overflow; c0 += 8) {
      int cf[8];
      int4 fx[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ci = x.at(B.NC, c0 + j < cc ? c0 + j : cc - 1);
        cf[j] = B.c_flags[ci];
        fx[j] = B.c_fix[ci];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (c0 + j >= cc) break;
        if (!(cf[j] & B2GPU_CONTACT_ENABLED) || !(cf[j] & B2GPU_CONTACT_TOUCHING)) continue;
        const uint32_t wa = fix_word(fx[j].x), wbq = fix_word(fx[j].y);
        if ((wa | wbq) & 0x80000000u) continue;  // sensors never join an island
        if (ne >= L.ECAP) { overflow = true; break; }
        const int ba = (int)wa, bb = (int)wbq;
        const uint32_t na = chead[ba * 32 + lane];
        chead[ba * 32 + lane] = (uint16_t)(2 * ne);
        const uint32_t nb_ = chead[bb * 32 + lane];  // after the A push; a self pair is impossible (add_pair rejects it)
        chead[bb * 32 + lane] = (uint16_t)(2 * ne + 1);
        enext[ne * 32 + lane] = na | (nb_ << 16);
        ebody[ne * 32 + lane] = (uint32_t)ba | ((uint32_t)bb << 16);
        eorig[ne * 32 + lane] = (uint16_t)(c0 + j);
        eisl[ne * 32 + lane] = 0;
        ++ne;
      }
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
        korder[ncon * 32 + lane] = (uint16_t)ei;
        ++ncon;
        const uint32_t bod = ebody[ei * 32 + lane];
        const int other = side ? (int)(bod & 0xffffu) : (int)(bod >> 16);
        const uint32_t of = bflag[other * 32 + lane];
        if (of & 8) {  // static: joins every island that touches it, once (flag un-set after each island, :500-506)
          if (smark[other * 32 + lane] == (uint16_t)(nisl + 1)) continue;
          smark[other * 32 + lane] = (uint16_t)(nisl + 1);
        } else {
          if (of & 1) continue;
          bflag[other * 32 + lane] = (uint8_t)(of | 1);
        }
        stack[(sp_++) * 32 + lane] = (uint16_t)other;
      }
    }
    B.isl_range[x.at(NB, nisl)] = make_int4(body_first, nbod, contact_first, ncon);
    ++nisl;
  }
  // ---- write the ISLAND / AWAKE flags of the bodies back (contacts: see ContactIslandFlagsK)
  for (int b0 = 0; b0 < NB; b0 += 8) {
    int f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = b0 + j < NB ? B.b_flags[x.at(NB, b0 + j)] : 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (b0 + j >= NB) break;
      const uint32_t bf = bflag[(b0 + j) * 32 + lane];
      const int nf = (f[j] & ~(B2GPU_BODY_ISLAND | B2GPU_BODY_AWAKE)) | ((bf & 1) ? B2GPU_BODY_ISLAND : 0) |
                     ((bf & 2) ? B2GPU_BODY_AWAKE : 0);
      if (nf != f[j]) B.b_flags[x.at(NB, b0 + j)] = nf;
    }
  }
  ws[WS_SCHED_ROUNDS] = -1;  // no level schedule: the Gauss-Seidel stages keep list order (round-1's level-scheduled
                             // two-lane kernels measured slower and were removed, profiles/r01_ncu_summary.md)
  ws[WS_ISL_COUNT] = nisl;
  ws[WS_ISL_BODIES] = nbod;
  ws[WS_ISL_CONTACTS] = ncon;
  ws[WS_ST_ISLANDS] = nisl;
  ws[WS_ST_ISL_BODIES] = nbod;
  ws[WS_ST_ISL_CONTACTS] = ncon;
  ws[WS_TOPO_DIRTY] = dirty_next ? 1 : 0;
  ws[WS_ISL_VALID] = 1;
}

}  // namespace b2g
