// evp_persist_plan.h -- host-side planning for KERNEL_PERSISTENT (evp_persist.cu): the tiling of a sub-domain over the SMs,
// the shared-memory layout of a tile, and the (slot, thread) -> cell tables.  Plain C++ (used by evp_abi.cu and, unchanged, by the
// host emulation of the kernel in tests/emu_persist.cpp).
#pragma once
#include <algorithm>
#include <string>
#include <vector>

#include "evp_internal.h"

namespace evp {

struct PersistTables {
  std::vector<unsigned> tslot, uslot;  // [4 shapes][PERSIST_SLOTS * nthreads] packed words, see persist_word
};

// table word of one (slot, thread): position in the tile (t = lj*pitch + li), its coordinates, and `pidx`, the index of the cell in
// the arrays only its own thread touches (stresses, static operands).  pidx is dense in the order the warps hold the cells, so a
// warp's 32 lanes always touch 32 consecutive elements there -- no shared-memory bank conflict, whatever the tile shape.
constexpr unsigned PERSIST_NONE = 0xffffffffu;
inline unsigned persist_word(int t, int li, int lj, int pidx) { return (unsigned)t | ((unsigned)li << 10) | ((unsigned)lj << 16) | ((unsigned)pidx << 22); }

// extent of a tile of shape s (bit 0: last tile column, bit 1: last tile row) in U points
inline void persist_shape_extent(const PersistPlan &pp, int nx, int ny, int shape, int &ebx, int &eby) {
  ebx = (shape & 1) ? nx - (pp.ntx - 1) * pp.bx : pp.bx;
  eby = (shape & 2) ? ny - (pp.nty - 1) * pp.by : pp.by;
}

// shared-memory layout for an instantiation that keeps kT static T arrays and kU static U arrays on chip
inline bool persist_layout(PersistPlan &pp, int kT, int kU, size_t smem_max) {
  pp.nT = (pp.bx + 1) * (pp.by + 1);
  pp.nU = pp.bx * pp.by;
  pp.nring = (pp.bx + 2) * (pp.by + 2);
  pp.kT = kT; pp.kU = kU;
  pp.off_uv = 0; pp.off_str = 2 * pp.nring; pp.off_sig = pp.off_str + 8 * pp.nT;
  pp.off_T = pp.off_sig + 12 * pp.nT; pp.off_U = pp.off_T + kT * pp.nT;
  pp.smem_bytes = (unsigned)((8 * (pp.off_U + kU * pp.nU) + 15) & ~15);
  return pp.smem_bytes <= smem_max;
}

// The (slot, thread) tables of the four tile shapes; false if a shape does not fit PERSIST_SLOTS cells per thread.
//   T cells: interior cells (they read only the tile's own velocities) fill slot 0 of every warp, then slot 1 from warp 0 up; the
//            tile-edge cells -- the only readers of the ring -- sit in slot 1 of the LAST ewT warps.  Warps left without a slot-1
//            cell ("light" warps, lw0 .. lw0+nlw-1) refresh the ring while the others relax their first cell.
//   U points: the tile-edge points (what other tiles read) in slot 0 of the FIRST ewU warps, published at once; interior points
//            fill slot 0 and then slot 1 of the other warps, and only then slot 1 of the first ewU warps.
// worst[0] / worst[1]: the largest number of T / U warp-tasks any scheduler (warp % 4) gets in one subcycle.
inline bool persist_build_tables(PersistPlan &pp, int nx, int ny, PersistTables &tb, int worst[2]) {
  const int nthreads = pp.nthreads, cap = PERSIST_SLOTS * nthreads, nwarps = nthreads / 32;
  const int tw = pp.bx + 1;
  static_assert(PERSIST_SLOTS == 2, "the fill order below is written for two slots");
  if (pp.bx + 1 > 63 || pp.by + 1 > 63 || pp.nT > 1024) return false;
  tb.tslot.assign((size_t)4 * cap, PERSIST_NONE);
  tb.uslot.assign((size_t)4 * cap, PERSIST_NONE);
  worst[0] = worst[1] = 0;
  for (int s = 0; s < 4; ++s) {
    pp.ewT[s] = pp.ewU[s] = pp.lw0[s] = pp.nlw[s] = 0;
    int ebx, eby;
    persist_shape_extent(pp, nx, ny, s, ebx, eby);
    if (ebx < 1 || eby < 1 || ebx > pp.bx || eby > pp.by) return false;
    struct Cell { int t, li, lj; };
    std::vector<Cell> tin, ted, uin, ued;
    for (int lj = 0; lj <= eby; ++lj)
      for (int li = 0; li <= ebx; ++li)
        ((li == 0 || lj == 0 || li == ebx || lj == eby) ? ted : tin).push_back({lj * tw + li, li, lj});
    for (int lj = 0; lj < eby; ++lj)
      for (int li = 0; li < ebx; ++li)
        ((li == 0 || lj == 0 || li == ebx - 1 || lj == eby - 1) ? ued : uin).push_back({lj * pp.bx + li, li, lj});
    const int ewT = ((int)ted.size() + 31) / 32, ewU = ((int)ued.size() + 31) / 32;
    if (ewT > nwarps || ewU > nwarps) return false;
    if ((int)tin.size() > cap - 32 * ewT) return false;
    pp.ewT[s] = ewT; pp.ewU[s] = ewU;
    unsigned *T = tb.tslot.data() + (size_t)s * cap, *U = tb.uslot.data() + (size_t)s * cap;
    // T: interior at positions 0.., edge in the last ewT warps of slot 1; pidx: interior 0..nin-1, edge nin..
    for (size_t q = 0; q < tin.size(); ++q) T[q] = persist_word(tin[q].t, tin[q].li, tin[q].lj, (int)q);
    for (size_t q = 0; q < ted.size(); ++q) T[cap - 32 * ewT + q] = persist_word(ted[q].t, ted[q].li, ted[q].lj, (int)(tin.size() + q));
    // light warps: no slot-1 cell and not an edge warp
    const int used1 = ((int)tin.size() > nthreads) ? ((int)tin.size() - nthreads + 31) / 32 : 0;  // warps of slot 1 holding interior cells
    pp.lw0[s] = used1;
    pp.nlw[s] = std::max(0, std::min(4, nwarps - ewT - used1));
    // U: edge at positions 0..32*ewU-1 (pidx 0..), interior: slot 0 of warps ewU.., slot 1 of warps ewU.., slot 1 of warps 0..ewU-1
    for (size_t q = 0; q < ued.size(); ++q) U[q] = persist_word(ued[q].t, ued[q].li, ued[q].lj, (int)q);
    std::vector<int> pos;
    for (int q = 32 * ewU; q < nthreads; ++q) pos.push_back(q);
    for (int q = nthreads + 32 * ewU; q < cap; ++q) pos.push_back(q);
    for (int q = nthreads; q < nthreads + 32 * ewU; ++q) pos.push_back(q);
    if (uin.size() > pos.size()) return false;
    for (size_t q = 0; q < uin.size(); ++q) U[pos[q]] = persist_word(uin[q].t, uin[q].li, uin[q].lj, (int)(ued.size() + q));
    for (int which = 0; which < 2; ++which) {
      const unsigned *tab = which ? U : T;
      int per[4] = {0, 0, 0, 0};
      for (int r = 0; r < PERSIST_SLOTS; ++r)
        for (int w = 0; w < nwarps; ++w) {
          bool any = false;
          for (int l = 0; l < 32; ++l) any = any || tab[r * nthreads + w * 32 + l] != PERSIST_NONE;
          if (any) ++per[w % 4];
        }
      worst[which] = std::max(worst[which], *std::max_element(per, per + 4));
    }
  }
  return true;
}

// Choose the tiling: at most one tile per SM; every shape fits the slots; the tile fits shared memory with one of the
// instantiations (all static arrays on chip, or 6 T + 3 U arrays).  Fewest T warp-tasks on the busiest scheduler wins (that is
// what a subcycle costs), then the smallest tile, then the widest rows.
inline bool persist_plan(int nx, int ny, int num_sms, int nthreads, size_t smem_max, PersistPlan &out, PersistTables &tb, std::string &why) {
  bool found = false;
  long best_cost = 0;
  for (int ntx = 1; ntx <= num_sms; ++ntx)
    for (int nty = 1; ntx * nty <= num_sms; ++nty) {
      PersistPlan pp{};
      pp.ntx = ntx; pp.nty = nty; pp.nthreads = nthreads;
      pp.bx = (nx + ntx - 1) / ntx; pp.by = (ny + nty - 1) / nty;
      if ((nx + pp.bx - 1) / pp.bx != ntx || (ny + pp.by - 1) / pp.by != nty) continue;  // no empty tiles
      if ((pp.bx + 1) * (pp.by + 1) > PERSIST_SLOTS * nthreads || pp.bx + 1 > 63 || pp.by + 1 > 63) continue;
      if (!persist_layout(pp, 10, 11, smem_max) && !persist_layout(pp, 6, 3, smem_max)) continue;
      PersistTables t;
      int worst[2];
      if (!persist_build_tables(pp, nx, ny, t, worst)) continue;
      const long cost = ((long)worst[0] * 64 + worst[1]) * 4096L * 4096L + (long)pp.nT * 4096L - pp.bx;
      if (!found || cost < best_cost) { found = true; best_cost = cost; out = pp; tb = std::move(t); }
    }
  if (!found) why = "no tiling with <= 1 tile per SM keeps the sub-domain's carried state on chip";
  return found;
}

}  // namespace evp
