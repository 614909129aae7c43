// emu_bgrid.cpp -- the B-grid CUDA kernels of cice_b200/csrc/evp_kernels.cu run THREAD BY THREAD ON THE HOST.
//
// Test infrastructure only (tests/test_emu_bgrid.py); nothing in the product links this.  The kernel translation unit is
// included unchanged with EVP_HOST_EMU defined (launchers compiled out, PTX helpers replaced by their plain C++ meaning:
// evp_ptx.cuh) on top of tests/cuda_emu.h, and compiled with g++ -ffp-contract=off like the `exact` namespace.
#include "cuda_emu.h"

#define EVP_HOST_EMU 1
#define EVP_USE_PDL 0
#define EVP_NS exact
#include "evp_kernels.cu"

using namespace evp;
using namespace evp::exact;

namespace {
struct Host {
  Dom d{};
  std::vector<unsigned char> mT, mU;
  std::vector<double> sig1, u1, v1, uinit, vinit, str;
};

template <int SPEC>
void fused_step(const Dom &d, const KParams &k, int cur, int flags) {
  static const P2PParams nop2p{};
  emu::launch({(d.nx + 30) / 31, (d.ny + 6) / 7, 1}, {32, 8, 1}, [&] { fused_kernel<32, 8, 2, false, SPEC>(d, k, cur, nop2p, 0, flags); });
}
}  // namespace

// kind 0: stress_kernel + stepu_kernel (in place)            sub: unused
// kind 1: fused_kernel<32,8,2,false,SPEC>                     sub: SPEC bits (1 speculative loads, 2 cp.async, 4 interleaved div/sqrt, 32 derived
//                                                                  geometry); 4 = L2-resident form, 3 = HBM-streaming form, 35 = streaming + derived
// kind 5: fused_kernel<32,8,2,true,4> without peers           sub: 0 = edge tiles first and counted, 1 = none counted
// One block in the reference's layout (nghost = 1) IS a dom: ld = nx_block, interior 1..nx_block-2.
extern "C" int emu_bgrid_run(int kind, int sub, int nxb, int nyb, int wrap_ew, int wrap_ns, const KParams *kp, int ndte,
                             const int32_t *maskT, const int32_t *maskU, double *sig /*[12][n]*/, double *u, double *v,
                             const double *geo /*[10][n]*/, const double *strength, const double *in /*[11][n]*/, double *diag /*[4][n]*/) {
  const size_t n = (size_t)nxb * nyb;
  Host h;
  h.mT.resize(n); h.mU.resize(n);
  for (size_t q = 0; q < n; ++q) { h.mT[q] = maskT[q] != 0; h.mU[q] = maskU[q] != 0; }
  h.sig1.assign(sig, sig + 12 * n); h.u1.assign(u, u + n); h.v1.assign(v, v + n); h.uinit.assign(u, u + n); h.vinit.assign(v, v + n);
  h.str.assign(8 * n, 0.0);
  Dom &d = h.d;
  d.nx = nxb - 2; d.ny = nyb - 2; d.ld = nxb; d.nyd = nyb; d.wrap_ew = wrap_ew; d.wrap_ns = wrap_ns;
  d.u[0] = u; d.u[1] = h.u1.data(); d.v[0] = v; d.v[1] = h.v1.data();
  for (int q = 0; q < 12; ++q) { d.sig[0][q] = sig + q * n; d.sig[1][q] = h.sig1.data() + q * n; }
  for (int q = 0; q < 8; ++q) d.str[q] = h.str.data() + q * n;
  d.strength = strength;
  d.dxT = geo; d.dyT = geo + n; d.dxhy = geo + 2 * n; d.dyhx = geo + 3 * n; d.cxp = geo + 4 * n; d.cyp = geo + 5 * n;
  d.cxm = geo + 6 * n; d.cym = geo + 7 * n; d.DminTarea = geo + 8 * n; d.uarear = geo + 9 * n;
  d.cdn = in; d.aiu = in + n; d.uocn = in + 2 * n; d.vocn = in + 3 * n; d.waterx = in + 4 * n; d.watery = in + 5 * n;
  d.forcex = in + 6 * n; d.forcey = in + 7 * n; d.umassdti = in + 8 * n; d.fm = in + 9 * n; d.TbU = in + 10 * n;
  d.uinit = h.uinit.data(); d.vinit = h.vinit.data();
  d.strintx = diag; d.strinty = diag + n; d.taubx = diag + 2 * n; d.tauby = diag + 3 * n;
  d.maskT = h.mT.data(); d.maskU = h.mU.data();
  const KParams &k = *kp;

  // kind 5: the in-kernel-halo instantiation on one rank without peers
  P2PParams pp{};
  std::vector<int> order, push_start(2 * d.nx + 2 * d.ny + 8, 0);
  unsigned long long done = 0, epoch = 1, flags_mem[64] = {};
  int err = 0;
  if (kind == 5) {
    const int ntx = (d.nx + 30) / 31, nty = (d.ny + 6) / 7;
    int n_edge = 0;
    for (int pass = 0; pass < 2; ++pass)
      for (int t = 0; t < ntx * nty; ++t) {
        const int bx = t % ntx, by = t / ntx;
        const bool edge = (bx == 0 || bx == ntx - 1 || by == 0 || by == nty - 1);
        if (edge == (pass == 0)) { order.push_back(((t / ntx) << 16) | (t % ntx)); n_edge += edge; }
      }
    pp.enabled = 1; pp.npeers = 0; pp.n_edge_tiles = (sub & 1) ? 0 : n_edge; pp.ntx = ntx; pp.nty = nty;
    pp.tile_order = order.data(); pp.push_start = push_start.data(); pp.push_peer = push_start.data(); pp.push_dst = push_start.data();
    pp.my_flags = flags_mem; pp.done_ctr = &done; pp.epoch_base = &epoch; pp.err = &err;
  }

  int cur = 0;
  for (int ks = 0; ks < ndte; ++ks) {
    const int flags = (ks == ndte - 1) ? 1 : 0;
    if (kind == 0) {
      emu::launch({(d.nx + 1 + 31) / 32, (d.ny + 1 + 7) / 8, 1}, {32, 8, 1}, [&] { stress_kernel(d, k, 0); });
      emu::launch({(d.nx + 31) / 32, (d.ny + 7) / 8, 1}, {32, 8, 1}, [&] { stepu_kernel(d, k, 0); });
      continue;  // in place
    } else if (kind == 1) {
      switch (sub) {
        case 1: fused_step<1>(d, k, cur, flags); break;
        case 2: fused_step<2>(d, k, cur, flags); break;
        case 3: fused_step<3>(d, k, cur, flags); break;
        case 4: fused_step<4>(d, k, cur, flags); break;
        case 5: fused_step<5>(d, k, cur, flags); break;
        case 7: fused_step<7>(d, k, cur, flags); break;
        case 35: fused_step<3 | 32>(d, k, cur, flags); break;   // streaming form + derived geometry (emu_set_metric first)
        case 36: fused_step<4 | 32>(d, k, cur, flags); break;   // resident form + derived geometry
        default: return 1;
      }
    } else if (kind == 5) {
      emu::launch({pp.ntx * pp.nty, 1, 1}, {32, 8, 1}, [&] { fused_kernel<32, 8, 2, true, 4>(d, k, cur, pp, ks, flags); });
      if (err) return 2;
    } else {
      return 1;
    }
    cur ^= 1;
  }
  if (cur == 1) {  // the result sits in copy 1
    memcpy(sig, h.sig1.data(), 12 * n * sizeof(double));
    memcpy(u, h.u1.data(), n * sizeof(double));
    memcpy(v, h.v1.data(), n * sizeof(double));
  }
  return 0;
}

// the two post-loop kernels on one block: deform_kernel (T cells 1..nx+1, 1..ny+1) and finish_kernel (U points 1..nx, 1..ny)
extern "C" int emu_deform_run(int nxb, int nyb, const int32_t *maskT, const double *u, const double *v, const double *geo /*[10][n]*/,
                              const double *dxU, const double *dyU, const double *tarear, double *out /*[5][n]: divu shear vort rdg_conv
                              rdg_shear*/, double e_factor) {
  const size_t n = (size_t)nxb * nyb;
  std::vector<unsigned char> mT(n);
  for (size_t q = 0; q < n; ++q) mT[q] = maskT[q] != 0;
  Dom d{};
  d.nx = nxb - 2; d.ny = nyb - 2; d.ld = nxb; d.nyd = nyb;
  d.dxT = geo; d.dyT = geo + n; d.cxp = geo + 4 * n; d.cyp = geo + 5 * n; d.cxm = geo + 6 * n; d.cym = geo + 7 * n;
  d.maskT = mT.data();
  emu::launch({(d.nx + 1 + 31) / 32, (d.ny + 1 + 7) / 8, 1}, {32, 8, 1},
              [&] { deform_kernel(d, u, v, dxU, dyU, tarear, out, out + n, out + 2 * n, out + 3 * n, out + 4 * n, e_factor); });
  return 0;
}
extern "C" int emu_finish_run(int nxb, int nyb, const int32_t *maskU, const double *u, const double *v, const double *in /*[11][n]*/,
                              double *strocnx, double *strocny, double rhow, double cosw, double sinw) {
  const size_t n = (size_t)nxb * nyb;
  std::vector<unsigned char> mU(n);
  for (size_t q = 0; q < n; ++q) mU[q] = maskU[q] != 0;
  Dom d{};
  d.nx = nxb - 2; d.ny = nyb - 2; d.ld = nxb; d.nyd = nyb;
  d.cdn = in; d.aiu = in + n; d.uocn = in + 2 * n; d.vocn = in + 3 * n; d.fm = in + 9 * n;
  d.maskU = mU.data();
  emu::launch({(d.nx + 31) / 32, (d.ny + 7) / 8, 1}, {32, 8, 1}, [&] { finish_kernel(d, u, v, strocnx, strocny, rhow, cosw, sinw); });
  return 0;
}

// derived geometry: publish the metric arrays to the kernels (what set_metric does on the device) and run the bitwise check
extern "C" int emu_set_metric(int nxb, int nyb, const double *geo /*[10][n]*/, const double *HTN, const double *HTE, double deltamin,
                              int skip_e, int skip_n) {
  const size_t n = (size_t)nxb * nyb;
  Dom d{};
  d.nx = nxb - 2; d.ny = nyb - 2; d.ld = nxb; d.nyd = nyb;
  d.dxT = geo; d.dyT = geo + n; d.dxhy = geo + 2 * n; d.dyhx = geo + 3 * n; d.cxp = geo + 4 * n; d.cyp = geo + 5 * n;
  d.cxm = geo + 6 * n; d.cym = geo + 7 * n; d.DminTarea = geo + 8 * n;
  int bad = 0;
  emu::launch({(d.nx + 1 + 31) / 32, (d.ny + 1 + 7) / 8, 1}, {32, 8, 1}, [&] { metric_verify_kernel(d, HTN, HTE, deltamin, skip_e, skip_n, &bad); });
  c_HTN = HTN; c_HTE = HTE; c_deltamin = deltamin;
  return bad;
}

// the one-kernel halo update of a single rank (tripole fold): p2p_fold_kernel without peers on the host-exported plan
extern "C" int emu_halo_local(double *U, double *V, const int *dst, const int *c1, const int *c2, const signed char *code, int n) {
  if (n > FOLD_MAX) return 1;
  static const P2PParams nop2p{};
  emu::launch({1, 1, 1}, {FOLD_THREADS, 1, 1}, [&] { p2p_fold_kernel(nop2p, U, V, dst, c1, c2, code, n, 0, 0); });
  return 0;
}
