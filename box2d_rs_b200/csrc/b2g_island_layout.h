// b2g_island_layout.h — shared-memory carve-up of the island DFS kernel (b2g_island_smem.cuh).
#pragma once
#include <stddef.h>

namespace b2g {

struct IslandSmemLayout {
  int NB, ECAP;
  size_t off_chead, off_stack, off_bflag, off_enext, off_ebody, off_eorig, off_eisl, off_smark, off_fxb, off_korder, total;
  int fxb_count;  // fixtures whose (body, sensor) word is cached in shared memory (0: read the global table)
};
inline IslandSmemLayout island_smem_layout(int NB, int NF, size_t budget) {
  IslandSmemLayout L;
  L.NB = NB;
  L.fxb_count = NF <= 4096 ? NF : 0;
  const size_t per_body = 32 * (2 + 2 + 2 + 1), per_edge = 32 * (4 + 4 + 2 + 1 + 2);
  const size_t fixed = (size_t)NB * per_body + (size_t)L.fxb_count * 4 + 512;
  long long ecap = budget > fixed ? (long long)((budget - fixed) / per_edge) : 0;
  ecap = (ecap / 4) * 4;
  if (ecap > 32760) ecap = 32760;
  L.ECAP = (int)ecap;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 15) / 16 * 16; return r; };
  L.off_enext = take((size_t)L.ECAP * 32 * 4);
  L.off_ebody = take((size_t)L.ECAP * 32 * 4);
  L.off_eorig = take((size_t)L.ECAP * 32 * 2);
  L.off_korder = take((size_t)L.ECAP * 32 * 2);
  L.off_chead = take((size_t)NB * 32 * 2);
  L.off_stack = take((size_t)NB * 32 * 2);
  L.off_smark = take((size_t)NB * 32 * 2);
  L.off_fxb = take((size_t)L.fxb_count * 4);
  L.off_eisl = take((size_t)L.ECAP * 32);
  L.off_bflag = take((size_t)NB * 32);
  L.total = o;
  return L;
}

}  // namespace b2g
