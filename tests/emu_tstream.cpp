// emu_tstream.cpp -- the tile-streaming kernel of cice_b200/csrc/evp_tstream.cu run THREAD BY THREAD ON THE HOST.
//
// Test infrastructure only (tests/test_emu_tstream.py); nothing in the product links this.  The kernel translation unit is included
// unchanged with EVP_HOST_EMU defined: a tensor map is a plain array description, a TMA box load a synchronous copy with the
// hardware's zero fill outside the tensor, an mbarrier a counter of completed phases (evp_tma.cuh); the CTAs of the persistent grid
// run concurrently as groups of host threads (tests/cuda_emu.h: launch_concurrent).  Compiled with g++ -ffp-contract=off like the
// `exact` namespace.  What this checks: the cut into strips, segments and blocks, box coordinates and shared-memory offsets, the
// ownership rules, the carried row, stage reuse and barrier phases, wrap stores, ping-pong parity.
#include "cuda_emu.h"

#define EVP_HOST_EMU 1
#define EVP_USE_PDL 0
#define EVP_NS exact
#include "evp_tstream.cu"

using namespace evp;
using namespace evp::exact;

// One block in the reference's layout (nghost = 1) IS a dom: ld = nx_block, interior 1..nx_block-2.
// rows: 12 or 6.  nb_override > 0 replaces the planner's blocks-per-segment; nctas > 0 the grid size (<= items).
extern "C" int emu_tstream_run(int rows, int nb_override, int nctas, int nxb, int nyb, int wrap_ew, int wrap_ns, int fold_top, const KParams *kp,
                               int ndte, const int32_t *maskT, const int32_t *maskU, double *sig /*[12][n]*/, double *u, double *v,
                               const double *geo /*[10][n]*/, const double *HTN, const double *HTE, double deltamin, const double *strength,
                               const double *in /*[11][n]*/, double *diag /*[4][n]*/, int *plan_out /*[5]: nstrips nseg nb nitems ctas*/) {
  const size_t n = (size_t)nxb * nyb;
  std::vector<unsigned char> mT(n), mU(n);
  for (size_t q = 0; q < n; ++q) { mT[q] = maskT[q] != 0; mU[q] = maskU[q] != 0; }
  std::vector<double> sig1(sig, sig + 12 * n), u1(u, u + n), v1(v, v + n), uinit(u, u + n), vinit(v, v + n);
  Dom d{};
  d.nx = nxb - 2; d.ny = nyb - 2; d.ld = nxb; d.nyd = nyb; d.wrap_ew = wrap_ew; d.wrap_ns = wrap_ns; d.fold_top = fold_top;
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

  // the tensor maps, as tstream_plan builds them
  static TsMaps tmaps;
  TmaMap *maps = tmaps.m;
  auto f64 = [&](int idx, const double *base, int bx, int by) { maps[idx] = TmaMap{base, 8, nxb, nyb, (long long)nxb * 8, bx, by}; };
  auto u8 = [&](int idx, const unsigned char *base) { maps[idx] = TmaMap{base, 1, nxb, nyb, (long long)nxb, TS_MW, rows}; };
  for (int b = 0; b < 2; ++b) {
    f64(TS_MAP_U + b, d.u[b], TS_W, rows + 1);
    f64(TS_MAP_V + b, d.v[b], TS_W, rows + 1);
    for (int q = 0; q < 12; ++q) f64(TS_MAP_SIG + 12 * b + q, d.sig[b][q], TS_W, rows);
  }
  f64(TS_MAP_STRENGTH, d.strength, TS_W, rows);
  f64(TS_MAP_DXT, d.dxT, TS_W, rows);
  f64(TS_MAP_DYT, d.dyT, TS_W, rows);
  f64(TS_MAP_HTN, HTN, TS_W, rows + 1);
  f64(TS_MAP_HTE, HTE, TS_W, rows);
  const double *uop[12] = {d.cdn, d.aiu, d.uocn, d.vocn, d.waterx, d.watery, d.forcex, d.forcey, d.umassdti, d.fm, d.uarear, d.TbU};
  for (int q = 0; q < 12; ++q) f64(TS_MAP_UOP + q, uop[q], TS_W, rows);
  u8(TS_MAP_MASKT, d.maskT);
  u8(TS_MAP_MASKU, d.maskU);

  TsPlan ts{};
  int err = 0;
  tstream_cut(d.nx, d.ny, 4, rows, &ts);
  if (nb_override > 0) {
    const int H = rows * nb_override;
    ts.nb = nb_override; ts.nseg = (d.ny + H - 2) / (H - 1); ts.nitems = ts.nstrips * ts.nseg;
    if (ts.ctas > ts.nitems) ts.ctas = ts.nitems;
  }
  if (nctas > 0) ts.ctas = nctas < ts.nitems ? nctas : ts.nitems;
  ts.maps = &tmaps; ts.err = &err; ts.deltamin = deltamin;
  if (plan_out) { plan_out[0] = ts.nstrips; plan_out[1] = ts.nseg; plan_out[2] = ts.nb; plan_out[3] = ts.nitems; plan_out[4] = ts.ctas; }

  const KParams &k = *kp;
  int cur = 0;
  for (int ks = 0; ks < ndte; ++ks) {
    const int last = (ks == ndte - 1) ? 1 : 0;
    if (rows == 12)
      emu::launch_concurrent(ts.ctas, 32 * 12, TsL<12>::TOTAL + 128, [&] { tstream_kernel<12, 1>(d, k, ts, tmaps, cur, last); });
    else if (rows == 6)
      emu::launch_concurrent(ts.ctas, 32 * 6, TsL<6>::TOTAL + 128, [&] { tstream_kernel<6, 2>(d, k, ts, tmaps, cur, last); });
    else
      return 1;
    if (err) return 2;
    if (evp::emu_tma_misaligned.load()) return 3;   // a box load that the hardware rejects (start not on a 16-byte boundary)
    cur ^= 1;
  }
  if (cur == 1) {  // the result sits in copy 1
    memcpy(sig, sig1.data(), 12 * n * sizeof(double));
    memcpy(u, u1.data(), n * sizeof(double));
    memcpy(v, v1.data(), n * sizeof(double));
  }
  return 0;
}

// the planner alone: the cut for a sub-domain of nx x ny cells on num_sms SMs
extern "C" void emu_tstream_cut(int nx, int ny, int num_sms, int rows, int *out /*[5]*/) {
  TsPlan ts{};
  tstream_cut(nx, ny, num_sms, rows, &ts);
  out[0] = ts.nstrips; out[1] = ts.nseg; out[2] = ts.nb; out[3] = ts.nitems; out[4] = ts.ctas;
}
