// evp_kernels.cu -- CUDA kernels of the EVP subcycle for sm_100a.
//
// Compiled twice (see build.py): -DEVP_NS=exact -fmad=false  and  -DEVP_NS=fast.
//
// Kernels here work on the device sub-domain layout of evp_internal.h (struct Dom):
//   stress_kernel / stepu_kernel : the reference's two sweeps, one thread per cell, `str` through
//                                  global memory (the first correct path; KERNEL_SPLIT)
//   fused_kernel                 : one launch per subcycle; a CTA relaxes the stresses of a
//                                  32 x 8 patch of T cells, hands the 8 `str` terms to the momentum
//                                  step through shared memory and advances the 31 x 7 U points the
//                                  patch closes.  Reads only the `cur` copies of the carried state and
//                                  writes only the other copy, so overlapping patches never race
//                                  (KERNEL_FUSED).  Three forms: L2-resident, HBM-streaming, HBM-streaming
//                                  with derived geometry; each also with the in-kernel NVLink halo (P2P)
//   p2p_fold_kernel              : what the halo update does beyond copying a neighbour's value (tripole
//                                  fold, ice_boundary.F90:1550-1724), after the peers' stores have arrived
// Neither is GEMM shaped; both are bandwidth/latency bound fp64 stencils -> no tensor cores.
#include "evp_math.cuh"
#include "evp_dom.cuh"
#include "evp_ptx.cuh"

#ifndef EVP_USE_PDL
#define EVP_USE_PDL 1
#endif

#ifndef EVP_NS
#error "compile with -DEVP_NS=exact or -DEVP_NS=fast"
#endif

namespace evp {
namespace EVP_NS {

__device__ __forceinline__ void load_sigma(const Dom &d, int b, int c, Sigma &s) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    s.p[q] = d.sig[b][q][c];
    s.m[q] = d.sig[b][4 + q][c];
    s.s12[q] = d.sig[b][8 + q][c];
  }
}
__device__ __forceinline__ void store_sigma(const Dom &d, int b, int c, const Sigma &s) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    d.sig[b][q][c] = s.p[q];
    d.sig[b][4 + q][c] = s.m[q];
    d.sig[b][8 + q][c] = s.s12[q];
  }
}

__device__ __forceinline__ void stress_at(const Dom &d, const KParams &k, int cur, int i, int j, Sigma &sg,
                                          double (&str)[8]) {
  const int c = at(d, i, j), w = c - 1, s = c - d.ld, sw = s - 1;
  const double *__restrict__ U = d.u[cur];
  const double *__restrict__ V = d.v[cur];
  load_sigma(d, cur, c, sg);
  stress_point(U[c], V[c], U[w], V[w], U[s], V[s], U[sw], V[sw], d.dxT[c], d.dyT[c], d.dxhy[c], d.dyhx[c],
               d.cxp[c], d.cyp[c], d.cxm[c], d.cym[c], d.DminTarea[c], d.strength[c], k, sg, str);
}

// ---------------------------------------------------------------------------------------------
// KERNEL_SPLIT
// ---------------------------------------------------------------------------------------------
// T cells (1..nx+1, 1..ny+1): the reference's index list includes the N and E ghost cells
// (ice_dyn_shared.F90:740-749).  Cells off the ice get str = 0 (`str(:,:,:) = c0`, ice_dyn_evp.F90:1537).
__global__ void __launch_bounds__(256) stress_kernel(const __grid_constant__ Dom d, const __grid_constant__ KParams k,
                                                     int cur) {
  const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (i > d.nx + 1 || j > d.ny + 1) return;
  const int c = at(d, i, j);
  double str[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (d.maskT[c]) {
    Sigma sg;
    stress_at(d, k, cur, i, j, sg, str);
    store_sigma(d, cur, c, sg);
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) d.str[q][c] = str[q];
}

__global__ void __launch_bounds__(256) stepu_kernel(const __grid_constant__ Dom d, const __grid_constant__ KParams k,
                                                    int cur) {
  const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (i > d.nx || j > d.ny) return;
  const int c = at(d, i, j);
  if (!d.maskU[c]) return;
  const int e = c + 1, n = c + d.ld, ne = n + 1;
  const UOut o = stepu_point(d.u[cur][c], d.v[cur][c], d.cdn[c], d.aiu[c], d.uocn[c], d.vocn[c], d.waterx[c],
                             d.watery[c], d.forcex[c], d.forcey[c], d.umassdti[c], d.fm[c], d.uarear[c], d.TbU[c],
                             d.uinit[c], d.vinit[c], d.str[0][c], d.str[1][e], d.str[2][n], d.str[3][ne],
                             d.str[4][c], d.str[5][n], d.str[6][e], d.str[7][ne], k);
  store_uv(d, d.u[cur], d.v[cur], i, j, o.u, o.v);
  d.strintx[c] = o.strintx;
  d.strinty[c] = o.strinty;
  d.taubx[c] = o.taubx;
  d.tauby[c] = o.tauby;
}

// deformations (ice_dyn_shared.F90:1756-1860): the step right after the loop, from the resident velocities
__global__ void __launch_bounds__(256) deform_kernel(const __grid_constant__ Dom d, const double *__restrict__ U,
                                                     const double *__restrict__ V, const double *__restrict__ dxU,
                                                     const double *__restrict__ dyU, const double *__restrict__ tarear,
                                                     double *__restrict__ divu, double *__restrict__ shear,
                                                     double *__restrict__ vort, double *__restrict__ rdg_conv,
                                                     double *__restrict__ rdg_shear, double e_factor) {
  const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (i > d.nx + 1 || j > d.ny + 1) return;
  const int c = at(d, i, j), w = c - 1, s = c - d.ld, sw = s - 1;
  if (!d.maskT[c]) return;
  const double ucc = U[c], vcc = V[c], uee = U[w], vee = V[w], use_ = U[s], vse = V[s], une = U[sw], vne = V[sw];
  const double dxT = d.dxT[c], dyT = d.dyT[c], cxp = d.cxp[c], cyp = d.cyp[c], cxm = d.cxm[c], cym = d.cym[c];
  double div[4], ten[4], shr[4];
  div[NE] = cyp * ucc - dyT * uee + cxp * vcc - dxT * vse;
  div[NW] = cym * uee + dyT * ucc + cxp * vee - dxT * vne;
  div[SW] = cym * une + dyT * use_ + cxm * vne + dxT * vee;
  div[SE] = cyp * use_ - dyT * une + cxm * vse + dxT * vcc;
  ten[NE] = -cym * ucc - dyT * uee + cxm * vcc + dxT * vse;
  ten[NW] = -cyp * uee + dyT * ucc + cxm * vee + dxT * vne;
  ten[SW] = -cyp * une + dyT * use_ + cxp * vne - dxT * vee;
  ten[SE] = -cym * use_ - dyT * une + cxp * vse - dxT * vcc;
  shr[NE] = -cym * vcc - dyT * vee - cxm * ucc - dxT * use_;
  shr[NW] = -cyp * vee + dyT * vcc - cxm * uee - dxT * une;
  shr[SW] = -cyp * vne + dyT * vse - cxp * une + dxT * uee;
  shr[SE] = -cym * vse - dyT * vne - cxp * use_ + dxT * ucc;
  double Delta[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) Delta[q] = sqrt(div[q] * div[q] + e_factor * (ten[q] * ten[q] + shr[q] * shr[q]));
  const double ta = tarear[c];
  // reference order of the corner sums: ne + nw + se + sw
  const double dv = 0.25 * (div[NE] + div[NW] + div[SE] + div[SW]) * ta;
  const double tmp = 0.25 * (Delta[NE] + Delta[NW] + Delta[SE] + Delta[SW]) * ta;
  divu[c] = dv;
  rdg_conv[c] = -fmin(dv, 0.0);
  rdg_shear[c] = 0.5 * (tmp - fabs(dv));
  const double tsum = ten[NE] + ten[NW] + ten[SE] + ten[SW], ssum = shr[NE] + shr[NW] + shr[SE] + shr[SW];
  shear[c] = 0.25 * ta * sqrt(tsum * tsum + ssum * ssum);
  const double dvdxn = dyU[c] * vcc - dyU[w] * vee;
  const double dvdxs = dyU[s] * vse - dyU[sw] * vne;
  const double dudye = dxU[c] * ucc - dxU[s] * use_;
  const double dudyw = dxU[w] * uee - dxU[sw] * une;
  vort[c] = 0.5 * ta * (dvdxn + dvdxs - dudye - dudyw);
}

// dyn_finish (ice_dyn_shared.F90:1291-1365): the ice-ocean stress at the ice U points, from the resident velocities and the
// U-point inputs of the last upload; points off the U list keep what the arrays held
__global__ void __launch_bounds__(256) finish_kernel(const __grid_constant__ Dom d, const double *__restrict__ U,
                                                     const double *__restrict__ V, double *__restrict__ strocnx,
                                                     double *__restrict__ strocny, double rhow, double cosw, double sinw) {
  const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (i > d.nx || j > d.ny) return;
  const int c = at(d, i, j);
  if (!d.maskU[c]) return;
  const double du = d.uocn[c] - U[c], dv = d.vocn[c] - V[c];
  double vrel = rhow * d.cdn[c] * sqrt(du * du + dv * dv);
  vrel = vrel * d.aiu[c];
  const double sg = copysign(1.0, d.fm[c]);
  strocnx[c] = vrel * (du * cosw - dv * sinw * sg);
  strocny[c] = vrel * (dv * cosw + du * sinw * sg);
}

// The step preparation of evp() between the halo updates of the T-point inputs and the subcycle loop, one
// thread per U point of the sub-domain:
//   * grid_average_X2YS 'NE' (ice_grid.F90:4187-4205) of tmass, aice_init, cdn_ocn, uocn, vocn (, ss_tltx, ss_tlty): masked,
//     area-weighted four-point averages, every product and sum in the reference's order;
//   * grid_average_X2YF 'NE' (ice_grid.F90:4644-4655) of the wind stress;
//   * dyn_prep2 (ice_dyn_shared.F90:705-837) at the U point: new ice mask from the OLD one the device kept, velocity of new ice
//     points = ocean current, zero where masked out, Coriolis, water and forcing terms (geostrophic or coupled tilt).
// Writes the eleven U-point inputs of the loop, the new mask and the velocities (both ping-pong copies, wrap ghosts included).
__global__ void __launch_bounds__(256) prep_kernel(const __grid_constant__ Dom d, const __grid_constant__ PrepArgs a) {
  const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (i > d.nx || j > d.ny) return;
  const int c = at(d, i, j), e = c + 1, n = c + d.ld, ne = n + 1;
  const double mc = a.hm[c], me = a.hm[e], mn = a.hm[n], mne = a.hm[ne];
  const double tc = a.tarea[c], te = a.tarea[e], tn = a.tarea[n], tne = a.tarea[ne];
  const double wtmp = mc * tc + me * te + mn * tn + mne * tne;
  auto avgS = [&](const double *w) {
    double r = 0.0;
    if (wtmp != 0.0) r = (mc * w[c] * tc + me * w[e] * te + mn * w[n] * tn + mne * w[ne] * tne) / wtmp;
    return r;
  };
  auto avgF = [&](const double *w) { return 0.25 * (w[c] * tc + w[e] * te + w[n] * tn + w[ne] * tne) / a.uarea[c]; };
  const double umass = avgS(a.tmass), aiU = avgS(a.aice), cdnU = avgS(a.cdn), uo = avgS(a.uocn), vo = avgS(a.vocn);
  const double sax = avgF(a.sax), say = avgF(a.say);
  const bool old = a.maskU[c] != 0;
  const bool ice = a.umask[c] && aiU > a.area_min && umass > a.mass_min;
  double u = d.u[0][c], v = d.v[0][c];
  double waterx = 0.0, watery = 0.0, forcex = 0.0, forcey = 0.0, umassdti = 0.0, fm = a.fm[c];
  if (ice) {
    if (!old) { u = uo; v = vo; }   // new ice points start from the ocean surface current
    umassdti = umass / a.dt;
    fm = a.fcor[c] * umass;
    const double sg = copysign(1.0, fm);
    waterx = uo * a.cosw - vo * a.sinw * sg;
    watery = vo * a.cosw + uo * a.sinw * sg;
    double tx, ty;
    if (a.coupled_tilt) {
      tx = -a.gravit * umass * avgS(a.tltx);
      ty = -a.gravit * umass * avgS(a.tlty);
    } else {
      tx = -fm * vo;
      ty = fm * uo;
    }
    forcex = sax + tx;
    forcey = say + ty;
  } else {
    u = 0.0; v = 0.0;
    a.strintx[c] = 0.0; a.strinty[c] = 0.0;
  }
  a.maskU[c] = ice ? 1 : 0;
  a.cdnU[c] = cdnU; a.aiU[c] = aiU; a.uocnU[c] = uo; a.vocnU[c] = vo;
  a.waterx[c] = waterx; a.watery[c] = watery; a.forcex[c] = forcex; a.forcey[c] = forcey;
  a.umassdti[c] = umassdti; a.fm[c] = fm;
  a.TbU[c] = a.TbU_in ? a.TbU_in[c] : 0.0;
  a.taubx[c] = 0.0; a.tauby[c] = 0.0;   // dyn_prep2 clears them everywhere; the loop's last subcycle sets them on the ice
  store_uv(d, d.u[0], d.v[0], i, j, u, v);
  store_uv(d, d.u[1], d.v[1], i, j, u, v);
}

#ifndef EVP_HOST_EMU  // launchers: not part of the host emulation (tests/emu_bgrid.cpp)
cudaError_t launch_prep(const Dom &d, const PrepArgs &args, cudaStream_t s) {
  dim3 b(32, 8), g((d.nx + b.x - 1) / b.x, (d.ny + b.y - 1) / b.y);
  prep_kernel<<<g, b, 0, s>>>(d, args);
  return cudaGetLastError();
}
cudaError_t launch_finish(const Dom &d, int cur, double *strocnx, double *strocny, double rhow, double cosw, double sinw, cudaStream_t s) {
  dim3 b(32, 8), g((d.nx + b.x - 1) / b.x, (d.ny + b.y - 1) / b.y);
  finish_kernel<<<g, b, 0, s>>>(d, d.u[cur], d.v[cur], strocnx, strocny, rhow, cosw, sinw);
  return cudaGetLastError();
}
cudaError_t launch_deform(const Dom &d, int cur, const double *dxU, const double *dyU, const double *tarear, double *divu,
                          double *shear, double *vort, double *rdg_conv, double *rdg_shear, double e_factor, cudaStream_t s) {
  dim3 b(32, 8), g((d.nx + 1 + b.x - 1) / b.x, (d.ny + 1 + b.y - 1) / b.y);
  deform_kernel<<<g, b, 0, s>>>(d, d.u[cur], d.v[cur], dxU, dyU, tarear, divu, shear, vort, rdg_conv, rdg_shear, e_factor);
  return cudaGetLastError();
}

cudaError_t launch_stress(const Dom &d, const KParams &p, int cur, cudaStream_t s) {
  dim3 b(32, 8), g((d.nx + 1 + b.x - 1) / b.x, (d.ny + 1 + b.y - 1) / b.y);
  stress_kernel<<<g, b, 0, s>>>(d, p, cur);
  return cudaGetLastError();
}
cudaError_t launch_stepu(const Dom &d, const KParams &p, int cur, cudaStream_t s) {
  dim3 b(32, 8), g((d.nx + b.x - 1) / b.x, (d.ny + b.y - 1) / b.y);
  stepu_kernel<<<g, b, 0, s>>>(d, p, cur);
  return cudaGetLastError();
}

#endif  // EVP_HOST_EMU

// ---------------------------------------------------------------------------------------------
// KERNEL_FUSED
// ---------------------------------------------------------------------------------------------
// FBX x FBY threads relax an FBX x FBY patch of T cells and advance the (FBX-1) x (FBY-1) U points it closes.
// MINB = CTAs per SM the register allocation is bounded for.

// Derived geometry (SPEC bit 5, after evp_b200_set_metric).
// Seven of the ten static T-cell arrays are functions of the two metric arrays HTN, HTE and of dxT, dyT
// (ice_dyn_shared.F90:384-388, 401-441): on sub-domains that stream from HBM, reading HTN/HTE (their i-1 / j-1 neighbours come
// from the same cache lines) instead of dxhy, dyhx, cxp, cyp, cxm, cym, DminTarea moves 360 instead of 400 B per cell and
// subcycle -- the algorithmic figure of SURVEY 8d (measured on B200: 3600x2400 164.4 vs 170.4 ms per step).  The expressions are
// the reference's, operation for operation; whether they reproduce the host's arrays bit for bit on every cell the loop can touch
// is CHECKED on the device when the metric arrays are handed over (metric_verify_kernel); if a single cell differs the library
// keeps reading the arrays.
__constant__ const double *c_HTN, *c_HTE;
__constant__ double c_deltamin;
// (derive_geometry itself: evp_math.cuh)
// counts the T cells (1..nx+1-skip_e, 1..ny+1-skip_n) on which the derived values differ from the arrays in any bit
__global__ void __launch_bounds__(256) metric_verify_kernel(const __grid_constant__ Dom d, const double *__restrict__ HTN,
                                                            const double *__restrict__ HTE, double deltamin, int skip_e, int skip_n,
                                                            int *__restrict__ mismatches) {
  const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (i > d.nx + 1 - skip_e || j > d.ny + 1 - skip_n) return;
  const int c = at(d, i, j);
  double dxhy, dyhx, cxp, cyp, cxm, cym, dmin;
  derive_geometry(HTN[c], HTN[c - d.ld], HTE[c], HTE[c - 1], d.dxT[c], d.dyT[c], deltamin, dxhy, dyhx, cxp, cyp, cxm, cym, dmin);
  const bool same = __double_as_longlong(dxhy) == __double_as_longlong(d.dxhy[c]) && __double_as_longlong(dyhx) == __double_as_longlong(d.dyhx[c]) &&
                    __double_as_longlong(cxp) == __double_as_longlong(d.cxp[c]) && __double_as_longlong(cyp) == __double_as_longlong(d.cyp[c]) &&
                    __double_as_longlong(cxm) == __double_as_longlong(d.cxm[c]) && __double_as_longlong(cym) == __double_as_longlong(d.cym[c]) &&
                    __double_as_longlong(dmin) == __double_as_longlong(d.DminTarea[c]);
  if (!same) atomicAdd(mismatches, 1);
}
#ifndef EVP_HOST_EMU  // launchers: not part of the host emulation (tests/emu_bgrid.cpp)
cudaError_t set_metric(const double *HTN, const double *HTE, double deltamin) {
  cudaError_t e = cudaMemcpyToSymbol(c_HTN, &HTN, sizeof HTN);
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_HTE, &HTE, sizeof HTE);
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_deltamin, &deltamin, sizeof deltamin);
  return e;
}
cudaError_t launch_metric_verify(const Dom &d, const double *HTN, const double *HTE, double deltamin, int skip_e, int skip_n,
                                 int *mismatches, cudaStream_t s) {
  dim3 b(32, 8), g((d.nx + 1 + b.x - 1) / b.x, (d.ny + 1 + b.y - 1) / b.y);
  metric_verify_kernel<<<g, b, 0, s>>>(d, HTN, HTE, deltamin, skip_e, skip_n, mismatches);
  return cudaGetLastError();
}
#endif  // EVP_HOST_EMU

// The forms of fused_kernel (SPEC bits).  Same arithmetic, same ownership rules in all of them:
//  bit 0 (SPT)  every operand of the T cell is requested before the ice masks are known (addresses are always inside the dom),
//               static operands even BEFORE the programmatic grid dependency is resolved, i.e. while the previous subcycle's
//               kernel is still draining;
//  bit 1 (CPU)  the 12 momentum operands go global -> shared with cp.async and are picked up after the CTA barrier;
//  bit 2 (IL)   the IEEE divisions / square roots of the four corners as interleaved fast-path sequences (evp_math.cuh);
//  bit 5 (DER)  derived geometry (above).
// FORM_RESIDENT = IL: sub-domains whose arrays fit the 126 MB L2 are latency bound (gx1: 2.22 ms per step against 2.39 for SPT|CPU);
// FORM_STREAM = SPT|CPU: larger ones stream from HBM and want every load in flight early (3600x2400: 164 vs 182 ms per step).
constexpr int FORM_RESIDENT = 4, FORM_STREAM = 3, FORM_STREAM_DER = 3 | 32;
constexpr int NUOP = 12;
template <int FBX, int FBY, bool P2P, int SPEC>
__device__ __forceinline__ void fused_body(const Dom &d, const KParams &k, int cur, const P2PParams &pp, int last, int i, int j,
                                           bool inT, int c, double (&sstr)[8][FBY][FBX]) {
  constexpr bool SPT = (SPEC & 1) != 0, CPU = (SPEC & 2) != 0, IL = (SPEC & 4) != 0, DER = (SPEC & 32) != 0;
  __shared__ double sU[CPU ? NUOP : 1][FBY * FBX];
  const int tx = threadIdx.x, ty = threadIdx.y, t = ty * FBX + tx;
  const int nxt = cur ^ 1;
  const bool uspot = tx < FBX - 1 && ty < FBY - 1 && i <= d.nx && j <= d.ny;
  if (CPU) {
    if (uspot) {
      const double *src[NUOP] = {d.cdn, d.aiu, d.uocn, d.vocn, d.waterx, d.watery, d.forcex, d.forcey, d.umassdti, d.fm, d.uarear, d.TbU};
#pragma unroll
      for (int q = 0; q < NUOP; ++q) cp_async8(&sU[q][t], src[q] + c);
    }
    cp_async_commit();
  }
  const unsigned mT = ld_nc_u8(d.maskT + c), mU = ld_nc_u8(d.maskU + c);
  double dxT, dyT, dxhy, dyhx, cxp, cyp, cxm, cym, dmin, strength;
  if (SPT) {
    dxT = ld_nc_f64(d.dxT + c); dyT = ld_nc_f64(d.dyT + c);
    if (DER) {
      const double hn = ld_nc_f64(c_HTN + c), hs = ld_nc_f64(c_HTN + c - d.ld), he = ld_nc_f64(c_HTE + c), hw = ld_nc_f64(c_HTE + c - 1);
      derive_geometry(hn, hs, he, hw, dxT, dyT, c_deltamin, dxhy, dyhx, cxp, cyp, cxm, cym, dmin);
    } else {
      dxhy = ld_nc_f64(d.dxhy + c); dyhx = ld_nc_f64(d.dyhx + c);
      cxp = ld_nc_f64(d.cxp + c); cyp = ld_nc_f64(d.cyp + c); cxm = ld_nc_f64(d.cxm + c); cym = ld_nc_f64(d.cym + c);
      dmin = ld_nc_f64(d.DminTarea + c);
    }
    strength = ld_nc_f64(d.strength + c);
  }
#if EVP_USE_PDL
  cudaGridDependencySynchronize();
#endif
  const int w = c - 1, s = c - d.ld, sw = s - 1;
  const double *__restrict__ U = d.u[cur];
  const double *__restrict__ V = d.v[cur];
  double ucc, vcc;
  double str[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (SPT) {
    ucc = ld_f64(U + c); vcc = ld_f64(V + c);
    const double uee = ld_f64(U + w), vee = ld_f64(V + w);
    const double use_ = ld_f64(U + s), vse = ld_f64(V + s), une = ld_f64(U + sw), vne = ld_f64(V + sw);
    Sigma sg;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      sg.p[q] = ld_f64(d.sig[cur][q] + c);
      sg.m[q] = ld_f64(d.sig[cur][4 + q] + c);
      sg.s12[q] = ld_f64(d.sig[cur][8 + q] + c);
    }
    if (inT && mT) {
      stress_point<IL>(ucc, vcc, uee, vee, use_, vse, une, vne, dxT, dyT, dxhy, dyhx, cxp, cyp, cxm, cym, dmin, strength, k, sg, str);
      // each T cell is stored by exactly one CTA: the one that holds it off its E/N overlap edge
      const bool own = (tx < FBX - 1 || i == d.nx + 1) && (ty < FBY - 1 || j == d.ny + 1);
      if (own) store_sigma(d, nxt, c, sg);
    }
  } else {
    ucc = U[c]; vcc = V[c];
    if (inT && mT) {
      Sigma sg;
      load_sigma(d, cur, c, sg);
      if (DER) {
        const double dxT_ = d.dxT[c], dyT_ = d.dyT[c];
        derive_geometry(c_HTN[c], c_HTN[s], c_HTE[c], c_HTE[w], dxT_, dyT_, c_deltamin, dxhy, dyhx, cxp, cyp, cxm, cym, dmin);
        stress_point<IL>(ucc, vcc, U[w], V[w], U[s], V[s], U[sw], V[sw], dxT_, dyT_, dxhy, dyhx, cxp, cyp, cxm, cym, dmin, d.strength[c], k, sg, str);
      } else
      stress_point<IL>(ucc, vcc, U[w], V[w], U[s], V[s], U[sw], V[sw], d.dxT[c], d.dyT[c], d.dxhy[c], d.dyhx[c],
                       d.cxp[c], d.cyp[c], d.cxm[c], d.cym[c], d.DminTarea[c], d.strength[c], k, sg, str);
      const bool own = (tx < FBX - 1 || i == d.nx + 1) && (ty < FBY - 1 || j == d.ny + 1);
      if (own) store_sigma(d, nxt, c, sg);
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) sstr[q][ty][tx] = str[q];
  if (CPU) cp_async_wait_all();  // own copies only: each thread reads back what it requested itself
  __syncthreads();
  // An off-ice point of the row below a tripole fold is not advanced by the momentum step but REWRITTEN by every halo update
  // (0.5*(own - mirror), pole points negated: ice_boundary.F90:1641-1649), so unlike every other off-ice cell it changes during the
  // loop: its current value travels to the other ping-pong copy (and to the peers) like a computed one.
  const bool doU = uspot && mU, carry = uspot && !mU && d.fold_top && j == d.ny;
  double un = ucc, vn = vcc;
  if (doU) {
    double uo[NUOP];
    if (CPU) {
#pragma unroll
      for (int q = 0; q < NUOP; ++q) uo[q] = sU[q][t];
    } else {
      uo[0] = d.cdn[c]; uo[1] = d.aiu[c]; uo[2] = d.uocn[c]; uo[3] = d.vocn[c]; uo[4] = d.waterx[c]; uo[5] = d.watery[c];
      uo[6] = d.forcex[c]; uo[7] = d.forcey[c]; uo[8] = d.umassdti[c]; uo[9] = d.fm[c]; uo[10] = d.uarear[c]; uo[11] = d.TbU[c];
    }
    double ui = 0.0, vi = 0.0;
    if (k.revp != 0.0 || ucc == 0.0 || vcc == 0.0) { ui = d.uinit[c]; vi = d.vinit[c]; }  // see load_uin
    const UOut o = stepu_point<IL>(ucc, vcc, uo[0], uo[1], uo[2], uo[3], uo[4], uo[5], uo[6], uo[7], uo[8], uo[9], uo[10], uo[11], ui, vi,
                               sstr[0][ty][tx], sstr[1][ty][tx + 1], sstr[2][ty + 1][tx], sstr[3][ty + 1][tx + 1], sstr[4][ty][tx],
                               sstr[5][ty + 1][tx], sstr[6][ty][tx + 1], sstr[7][ty + 1][tx + 1], k);
    un = o.u; vn = o.v;
    if (last) {
      // the loop overwrites these every subcycle and nothing reads them in between (ice_dyn_shared.F90:948-965);
      // only the last subcycle's values survive, as in the reference's own 1-D solver (calc_diag_1d, ice_dyn_core1d.F90:607)
      d.strintx[c] = o.strintx;
      d.strinty[c] = o.strinty;
      d.taubx[c] = o.taubx;
      d.tauby[c] = o.tauby;
    }
  }
  if (doU || carry) {
    store_uv(d, d.u[nxt], d.v[nxt], i, j, un, vn);
    if (P2P && is_push_point(d, pp.fold_row, i, j)) {
      // this point is a ghost cell (or a fold source) of up to three other sub-domains: store it there over NVLink right away,
      // so the traffic is spread over the kernel and long acknowledged when the hand-over fence below is issued.  Bit 8 of the
      // peer word: the value crosses the tripole fold and arrives negated (ice_boundary.F90:1689-1722, field_type_vector).
      const int e = edge_index(d, pp.fold_row, i, j);
      for (int q = pp.push_start[e]; q < pp.push_start[e + 1]; ++q) {
        const int pw = pp.push_peer[q], pr = pw & 0xff;
        const int dst = pp.push_dst[q];
        const bool neg = (pw & 0x100) != 0;
        pp.peer_u[nxt][pr][dst] = neg ? neg_f64(un) : un;
        pp.peer_v[nxt][pr][dst] = neg ? neg_f64(vn) : vn;
      }
    }
  }
}

template <int FBX, int FBY, int MINB, bool P2P, int SPEC>
__global__ void __launch_bounds__(FBX *FBY, MINB) fused_kernel(const __grid_constant__ Dom d,
                                                                const __grid_constant__ KParams k, int cur,
                                                                const __grid_constant__ P2PParams pp, int ksub, int flags) {
  __shared__ double sstr[8][FBY][FBX];
  const int tx = threadIdx.x, ty = threadIdx.y;
  int tbx = blockIdx.x, tby = blockIdx.y;
  bool edge_tile = false;
  const int last = flags & 1;
#if EVP_USE_PDL
  // programmatic dependent launch, early trigger: once every CTA of this grid has started, the next subcycle's CTAs may
  // be scheduled onto SMs this grid no longer fills (its last, partial wave); they fetch masks and static operands and
  // then block in cudaGridDependencySynchronize() until this grid has completed and flushed
  if (flags & 2) cudaTriggerProgrammaticLaunchCompletion();
#endif
  if (P2P) {
    const int b = blockIdx.x;
    // edge tiles first; the table holds (tby << 16 | tbx) so that no runtime integer division (~80 instructions and a MUFU
    // round trip at the top of every CTA) is needed to decode it
    const int tile = pp.tile_order[b];
    tbx = tile & 0xffff;
    tby = tile >> 16;
    edge_tile = b < pp.n_edge_tiles;
  }
  const int i = 1 + tbx * (FBX - 1) + tx;  // T cell of this thread
  const int j = 1 + tby * (FBY - 1) + ty;
  unsigned long long base = 0;
  unsigned long long *tl = (P2P && pp.dbg) ? pp.dbg + 8 + 8 * (size_t)ksub : nullptr;  // per-kernel timeline (ns)
  if (tl && tx == 0 && ty == 0) atomicMin(tl + 0, gtime());
  if (P2P && edge_tile) {
    // the ghost ring of copy `cur` was written by the neighbour GPUs during their previous subcycle
    base = *pp.epoch_base;
    const int t = ty * FBX + tx;
    const long long tw0 = clock64();
    if (t < pp.npeers) wait_flag(pp.my_flags + pp.peer_rank[t], base + (unsigned long long)ksub, pp.err);
    if (t == 0 && pp.dbg) {
      const unsigned long long dt = (unsigned long long)(clock64() - tw0);
      atomicAdd(pp.dbg, dt);
      atomicAdd(pp.dbg + 1, 1ULL);
      atomicMax(pp.dbg + 2, dt);
      if (tl) atomicMax(tl + 5, gtime());  // last edge CTA released from its wait
    }
    __syncthreads();
  }
  const bool inT = (i <= d.nx + 1) && (j <= d.ny + 1);
  const int c = at(d, inT ? i : 1, inT ? j : 1);
  fused_body<FBX, FBY, P2P, SPEC>(d, k, cur, pp, last, i, j, inT, c, sstr);
  if (P2P && edge_tile) {
    // Hand-over.  Every edge CTA counts itself done with gpu-scope ordering (a system-scope fence per CTA costs
    // ~3 us per kernel).  The CTA that arrives last issues the ONE system-scope fence -- cumulative over the NVLink
    // stores of all edge CTAs it has synchronised with through the counter -- and then raises the peers' flags.
    const int t = ty * FBX + tx;
    __syncthreads();
    if (t == 0) {
      __threadfence();
      const unsigned long long old = atomicAdd(pp.done_ctr, 1ULL);
      if (old + 1 == (unsigned long long)pp.n_edge_tiles * (unsigned long long)(ksub + 1)) {
        if (tl) tl[2] = gtime();  // last edge CTA done computing
        __threadfence_system();
        if (tl) tl[3] = gtime();  // fenced
        for (int q = 0; q < pp.npeers; ++q)
          st_relaxed_sys(pp.peer_flag[q], base + (unsigned long long)ksub + 1ULL);
        if (tl) tl[4] = gtime();  // flags written
      }
    }
  }
  if (tl && tx == 0 && ty == 0) atomicMax(tl + 1, gtime());
}

// loop hand-shake: "I have entered loop `base`" (my buffers are ready to be written), and the closing wait
__global__ void p2p_start_kernel(const __grid_constant__ P2PParams pp) {
  if (threadIdx.x == 0) {
    const unsigned long long base = *pp.epoch_base;
    __threadfence_system();
    for (int q = 0; q < pp.npeers; ++q) st_release_sys(pp.peer_flag[q], base);
  }
}
__global__ void p2p_finish_kernel(const __grid_constant__ P2PParams pp, int ndte) {
  const unsigned long long base = *pp.epoch_base;
  if ((int)threadIdx.x < pp.npeers) wait_flag(pp.my_flags + pp.peer_rank[threadIdx.x], base + (unsigned long long)ndte, pp.err);
  __syncthreads();
  if (threadIdx.x == 0) *pp.epoch_base = base + (unsigned long long)ndte + 1ULL;
}

// ---------------------------------------------------------------------------------------------
// The part of the (uvel,vvel) halo update that is more than copying a neighbour's value: the tripole fold
// (ice_boundary.F90:1550-1724).  One CTA per sub-domain that has such entries, once per subcycle, after the subcycle kernel:
//   * waits until every peer's stores of subcycle `ksub` have arrived (the same epoch flags the next subcycle kernel's edge
//     CTAs wait for) -- raw top-row values from other ranks sit in the staging rows ny+2, ny+3 of this rank's own arrays;
//   * reads ALL its sources (every output is a function of pre-update values only, evp_halo.cu), synchronises, writes.
// codes as in evp_halo.cu: 0 copy, 1 negate, 2 0.5*(a-b), 3 -(0.5*(a-b)).  With no peers (one rank) it is the whole fold.
// ---------------------------------------------------------------------------------------------
constexpr int FOLD_THREADS = 1024, FOLD_PER_THREAD = 8, FOLD_MAX = FOLD_THREADS * FOLD_PER_THREAD;
__global__ void __launch_bounds__(FOLD_THREADS) p2p_fold_kernel(const __grid_constant__ P2PParams pp, double *U, double *V,
                                                                const int *__restrict__ dst, const int *__restrict__ c1,
                                                                const int *__restrict__ c2, const signed char *__restrict__ code,
                                                                int n, int ksub, int pdl) {
  // the index lists are static: fetched while the subcycle kernel in front of this one is still running (programmatic launch)
  int kd[FOLD_PER_THREAD], ka[FOLD_PER_THREAD], kb[FOLD_PER_THREAD], kop[FOLD_PER_THREAD];
#pragma unroll
  for (int q = 0; q < FOLD_PER_THREAD; ++q) {
    const int kk = threadIdx.x + q * FOLD_THREADS;
    kd[q] = ka[q] = kb[q] = kop[q] = 0;
    if (kk < n) { kd[q] = dst[kk]; ka[q] = c1[kk]; kb[q] = c2[kk]; kop[q] = code[kk]; }
  }
#if EVP_USE_PDL
  if (pdl) {
    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();
  }
#endif
  if (pp.npeers > 0) {
    const unsigned long long base = *pp.epoch_base;
    if ((int)threadIdx.x < pp.npeers)
      wait_flag(pp.my_flags + pp.peer_rank[threadIdx.x], base + (unsigned long long)ksub + 1ULL, pp.err);
    __syncthreads();
  }
  double u[FOLD_PER_THREAD], v[FOLD_PER_THREAD];
#pragma unroll
  for (int q = 0; q < FOLD_PER_THREAD; ++q) {
    const int kk = threadIdx.x + q * FOLD_THREADS;
    u[q] = 0.0; v[q] = 0.0;
    if (kk < n) {
      const int a = ka[q], op = kop[q];
      // written by other SMs / other GPUs during this launch sequence: read through L2, never a stale L1 line
      u[q] = ld_cg_f64(U + a); v[q] = ld_cg_f64(V + a);
      if (op == 1) {
        u[q] = neg_f64(u[q]); v[q] = neg_f64(v[q]);
      } else if (op == 2 || op == 3) {
        // xavg = 0.5*(x1 + isign*x2) with isign = -1, x1 the partner with the lower i (ice_boundary.F90:1641-1646).  The lower
        // partner receives xavg, the upper one isign*xavg: -(0.5*(x1-x2)) and 0.5*(x2-x1) differ in the sign of a zero result
        const int b = kb[q];
        u[q] = 0.5 * (u[q] - ld_cg_f64(U + b));
        v[q] = 0.5 * (v[q] - ld_cg_f64(V + b));
        if (op == 3) { u[q] = neg_f64(u[q]); v[q] = neg_f64(v[q]); }
      }
    }
  }
  __syncthreads();  // every source has been read
#pragma unroll
  for (int q = 0; q < FOLD_PER_THREAD; ++q) {
    const int kk = threadIdx.x + q * FOLD_THREADS;
    if (kk < n) { U[kd[q]] = u[q]; V[kd[q]] = v[q]; }
  }
}

#ifndef EVP_HOST_EMU  // launchers: not part of the host emulation (tests/emu_bgrid.cpp)
cudaError_t launch_fold(const P2PParams &pp, double *U, double *V, const int *dst, const int *c1, const int *c2,
                        const signed char *code, int n, int ksub, int pdl, cudaStream_t s) {
  if (n > FOLD_MAX) return cudaErrorInvalidValue;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(1); cfg.blockDim = dim3(FOLD_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, p2p_fold_kernel, pp, U, V, dst, c1, c2, code, n, ksub, pdl);
}
int fold_max_entries() { return FOLD_MAX; }
cudaError_t set_wait_timeout(unsigned long long ns) { return cudaMemcpyToSymbol(g_wait_timeout_ns, &ns, sizeof ns); }

template <int SPEC>
static cudaError_t launch_fused_t(const Dom &d, const KParams &p, int cur, cudaStream_t s, bool pdl, int last) {
  constexpr int FBX = 32, FBY = 8;
  dim3 b(FBX, FBY), g((d.nx + FBX - 2) / (FBX - 1), (d.ny + FBY - 2) / (FBY - 1));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = g; cfg.blockDim = b; cfg.dynamicSmemBytes = 0; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  static const P2PParams nop2p{};
  return cudaLaunchKernelEx(&cfg, fused_kernel<FBX, FBY, 2, false, SPEC>, d, p, cur, nop2p, 0, last);  // last: bit 0 = last subcycle, bit 1 = early PDL trigger
}

// form: 0 L2-resident, 1 HBM-streaming, 2 HBM-streaming with derived geometry (evp_abi.cu chooses from the sub-domain size)
cudaError_t launch_fused(const Dom &d, const KParams &p, int cur, cudaStream_t s, int form, bool pdl, int last) {
  switch (form) {
    case 1: return launch_fused_t<FORM_STREAM>(d, p, cur, s, pdl, last);
    case 2: return launch_fused_t<FORM_STREAM_DER>(d, p, cur, s, pdl, last);
    default: return launch_fused_t<FORM_RESIDENT>(d, p, cur, s, pdl, last);
  }
}

// ksub = -1: loop start hand-shake; ksub <= -2: closing wait after -2-ksub subcycles
cudaError_t launch_fused_p2p(const Dom &d, const KParams &p, const P2PParams &pp, int cur, int ksub, int last, int form, cudaStream_t s) {
  if (ksub == -1) {
    p2p_start_kernel<<<1, 32, 0, s>>>(pp);
    return cudaGetLastError();
  }
  if (ksub <= -2) {
    p2p_finish_kernel<<<1, 32, 0, s>>>(pp, -2 - ksub);
    return cudaGetLastError();
  }
  // `last`: bit 0 = last subcycle, bit 1 = early programmatic-launch trigger, bit 2 = launch with the PDL attribute
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(pp.ntx * pp.nty); cfg.blockDim = dim3(32, 8); cfg.dynamicSmemBytes = 0; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = (last & 4) ? 1 : 0;
  const int flags = last & 3;
  switch (form) {
    case 1: return cudaLaunchKernelEx(&cfg, fused_kernel<32, 8, 2, true, FORM_STREAM>, d, p, cur, pp, ksub, flags);
    case 2: return cudaLaunchKernelEx(&cfg, fused_kernel<32, 8, 2, true, FORM_STREAM_DER>, d, p, cur, pp, ksub, flags);
    default: return cudaLaunchKernelEx(&cfg, fused_kernel<32, 8, 2, true, FORM_RESIDENT>, d, p, cur, pp, ksub, flags);
  }
}
#endif  // EVP_HOST_EMU

}  // namespace EVP_NS
}  // namespace evp
