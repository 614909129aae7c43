// emu_persist.cpp -- KERNEL_PERSISTENT (cice_b200/csrc/evp_persist.cu) run THREAD BY THREAD ON THE HOST, all CTAs at once.
//
// Test infrastructure only (tests/test_emu_persist.py); nothing in the product links this.  The kernel translation unit is included
// unchanged with EVP_HOST_EMU defined on top of tests/cuda_emu.h; the tiling, the shared-memory layout and the (slot, thread) tables
// come from the product's own planner (evp_persist_plan.h) for a pretended SM count and CTA size, so a handful of small tiles with
// 64- or 128-thread CTAs exercise what 144 tiles of 512 threads do on the GPU: interior-first / edge-last ordering, the ring refresh
// behind the neighbours' counters, early publication of the tile-edge velocities, narrower tiles in the last column / row,
// on-rank cyclic wrap, write-back of the carried state.
#include <algorithm>
#include <string>
#include <vector>

#include "cuda_emu.h"

#define EVP_HOST_EMU 1
#define EVP_NS exact
#include "evp_persist.cu"
#include "evp_persist_plan.h"

using namespace evp;
using namespace evp::exact;

template <int NT>
static int run_nt(const Dom &d, const KParams &k, PersistPlan &pp) {
  const int nctas = pp.ntx * pp.nty;
  static const P2PParams nop2p{};
  if (pp.kT == 10 && pp.kU == 11) emu::launch_concurrent(nctas, NT, pp.smem_bytes, [&] { persist_kernel<NT, 10, 11>(d, k, pp, nop2p); });
  else if (pp.kT == 6 && pp.kU == 3) emu::launch_concurrent(nctas, NT, pp.smem_bytes, [&] { persist_kernel<NT, 6, 3>(d, k, pp, nop2p); });
  else return 3;
  return 0;
}

// One block in the reference's layout (nghost = 1) IS a dom: ld = nx_block, interior 1..nx_block-2.
// force_k63: use the instantiation that keeps only 6 T + 3 U static arrays on chip even though everything would fit.
// info[0..7]: ntx, nty, bx, by, kT, kU, ewT[0], ewU[0]
extern "C" int emu_persist_run(int nthreads, int num_sms, int force_k63, int nxb, int nyb, int wrap_ew, int wrap_ns, const KParams *kp,
                               int ndte, const int32_t *maskT, const int32_t *maskU, double *sig /*[12][n]*/, double *u, double *v,
                               const double *geo /*[10][n]*/, const double *strength, const double *in /*[11][n]*/,
                               double *diag /*[4][n]*/, int *info) {
  const size_t n = (size_t)nxb * nyb;
  std::vector<unsigned char> mT(n), mU(n);
  for (size_t q = 0; q < n; ++q) { mT[q] = maskT[q] != 0; mU[q] = maskU[q] != 0; }
  std::vector<double> sig1(sig, sig + 12 * n), u1(u, u + n), v1(v, v + n), uinit(u, u + n), vinit(v, v + n);
  Dom d{};
  d.nx = nxb - 2; d.ny = nyb - 2; d.ld = nxb; d.nyd = nyb; d.wrap_ew = wrap_ew; d.wrap_ns = wrap_ns;
  d.u[0] = u; d.u[1] = u1.data(); d.v[0] = v; d.v[1] = v1.data();
  for (int q = 0; q < 12; ++q) { d.sig[0][q] = sig + q * n; d.sig[1][q] = sig1.data() + q * n; }
  d.strength = strength;
  d.dxT = geo; d.dyT = geo + n; d.dxhy = geo + 2 * n; d.dyhx = geo + 3 * n; d.cxp = geo + 4 * n; d.cyp = geo + 5 * n;
  d.cxm = geo + 6 * n; d.cym = geo + 7 * n; d.DminTarea = geo + 8 * n; d.uarear = geo + 9 * n;
  d.cdn = in; d.aiu = in + n; d.uocn = in + 2 * n; d.vocn = in + 3 * n; d.waterx = in + 4 * n; d.watery = in + 5 * n;
  d.forcex = in + 6 * n; d.forcey = in + 7 * n; d.umassdti = in + 8 * n; d.fm = in + 9 * n; d.TbU = in + 10 * n;
  d.uinit = uinit.data(); d.vinit = vinit.data();
  d.strintx = diag; d.strinty = diag + n; d.taubx = diag + 2 * n; d.tauby = diag + 3 * n;
  d.maskT = mT.data(); d.maskU = mU.data();

  PersistPlan pp{};
  PersistTables tb;
  std::string why;
  if (!persist_plan(d.nx, d.ny, num_sms, nthreads, 232448, pp, tb, why)) return 1;
  if (force_k63) {
    if (!persist_layout(pp, 6, 3, 232448)) return 1;
  }
  std::vector<unsigned> progress((size_t)PERSIST_CTR_STRIDE * pp.ntx * pp.nty, 0u);
  int err = 0;
  pp.tslot = tb.tslot.data(); pp.uslot = tb.uslot.data(); pp.progress = progress.data(); pp.err = &err;
  pp.ndte = ndte; pp.use_init = kp->revp != 0.0;
  info[0] = pp.ntx; info[1] = pp.nty; info[2] = pp.bx; info[3] = pp.by; info[4] = pp.kT; info[5] = pp.kU; info[6] = pp.ewT[0]; info[7] = pp.ewU[0];
  int rc;
  if (nthreads == 64) rc = run_nt<64>(d, *kp, pp);
  else if (nthreads == 128) rc = run_nt<128>(d, *kp, pp);
  else return 2;
  if (rc) return rc;
  if (err) return 4;
  if (ndte & 1) {  // the result sits in copy 1
    memcpy(sig, sig1.data(), 12 * n * sizeof(double));
    memcpy(u, u1.data(), n * sizeof(double));
    memcpy(v, v1.data(), n * sizeof(double));
  }
  return 0;
}
