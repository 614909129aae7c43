// evp_dom.cuh -- index and store helpers on the device sub-domain layout (struct Dom, evp_internal.h), shared by the B-grid
// kernels of evp_kernels.cu and evp_lane2.cuh.  Plain C++ apart from the CUDA function qualifiers, so that the kernels built
// on them can also be run thread by thread on the host (tests/emu_bgrid.cpp).
#pragma once
#include "evp_internal.h"

namespace evp {

// cell index inside a dom array; init refuses sub-domains of 2^31 cells or more, so 32 bits are enough and every
// access costs one IMAD.WIDE instead of a 64-bit add pair
__device__ __forceinline__ int at(const Dom &d, int i, int j) { return j * d.ld + i; }

// store a new velocity and, where the ghost ring aliases the rank's own interior (cyclic direction
// entirely local), the ghost copies too: the on-rank part of dyn_haloUpdate (ice_dyn_evp.F90:908-910)
__device__ __forceinline__ void store_uv(const Dom &d, double *__restrict__ U, double *__restrict__ V, int i, int j,
                                         double un, double vn) {
  U[at(d, i, j)] = un;
  V[at(d, i, j)] = vn;
  int ig = -1, jg = -1;
  if (d.wrap_ew) ig = (i == 1) ? d.nx + 1 : (i == d.nx ? 0 : -1);
  if (d.wrap_ns) jg = (j == 1) ? d.ny + 1 : (j == d.ny ? 0 : -1);
  if (ig >= 0) { U[at(d, ig, j)] = un; V[at(d, ig, j)] = vn; }
  if (jg >= 0) { U[at(d, i, jg)] = un; V[at(d, i, jg)] = vn; }
  if (ig >= 0 && jg >= 0) { U[at(d, ig, jg)] = un; V[at(d, ig, jg)] = vn; }
  // a 1-wide interior aliases both ghosts
  if (d.wrap_ew && d.nx == 1) { U[at(d, 0, j)] = un; V[at(d, 0, j)] = vn; }
  if (d.wrap_ns && d.ny == 1) { U[at(d, i, 0)] = un; V[at(d, i, 0)] = vn; }
}

// row index of the push CSR for a U point whose value some other sub-domain (or this one's own ghost ring, across a tripole
// fold) needs: the boundary points (i==1 | i==nx | j==1 | j==ny) and, on ranks below a tripole fold, row `fold_row` = ny-1
// (the ghost row ny+1 is fed from it, ice_boundary.F90:1689-1722).  Host twin: P2PState::setup (evp_halo.cu).  Shared by fused_kernel<..,P2P> and
// persist_kernel<..,P2P>.
__device__ __forceinline__ bool is_push_point(const Dom &d, int fold_row, int i, int j) {
  return i == 1 || i == d.nx || j == 1 || j == d.ny || j == fold_row;
}
__device__ __forceinline__ int edge_index(const Dom &d, int fold_row, int i, int j) {
  if (j == 1) return i - 1;
  if (j == d.ny) return d.nx + i - 1;
  if (j == fold_row) return 2 * d.nx + 2 * (d.ny - 2) + (i - 1);
  if (i == 1) return 2 * d.nx + (j - 2);
  return 2 * d.nx + (d.ny - 2) + (j - 2);
}

// operands of one U point.  uvel_init/vvel_init enter stepu only as revp * uvel_init (ice_dyn_shared.F90:957-958);
// in classic EVP revp = 0 and the product is a zero that can change the sum brlx*uold + 0 only when that sum is
// itself a zero, so the two arrays are read only then (or when revp != 0): same bits, 16 B per point less traffic.
__device__ __forceinline__ void load_uin(const Dom &d, const KParams &k, int cur, int c, double (&uin)[16]) {
  uin[0] = d.u[cur][c]; uin[1] = d.v[cur][c]; uin[2] = d.cdn[c]; uin[3] = d.aiu[c]; uin[4] = d.uocn[c]; uin[5] = d.vocn[c];
  uin[6] = d.waterx[c]; uin[7] = d.watery[c]; uin[8] = d.forcex[c]; uin[9] = d.forcey[c]; uin[10] = d.umassdti[c];
  uin[11] = d.fm[c]; uin[12] = d.uarear[c]; uin[13] = d.TbU[c];
  uin[14] = 0.0; uin[15] = 0.0;
  if (k.revp != 0.0 || uin[0] == 0.0 || uin[1] == 0.0) { uin[14] = d.uinit[c]; uin[15] = d.vinit[c]; }
}

}  // namespace evp
