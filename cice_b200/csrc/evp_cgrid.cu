// evp_cgrid.cu -- C-grid EVP subcycle (grid_ice = 'C', configs[2]) for sm_100a.
//
// One subcycle of ice_dyn_evp.F90:938-1097 is five kernels, cut exactly where a value is needed at a
// neighbour cell (= where the reference has a halo point):
//   k1  strain_rates_U                                  shared.F90:2319-2430      -> shearU halo
//   k2  strain_rates_Tdt + stressC_T                    shared.F90:2250-2311, evp.F90:1758-1883
//                                                                                 -> zetax2T, etax2T, stresspT, stressmT halo
//   k3  grid_average_X2YS 'NE' (T->U) + stressC_U       grid.F90:4159-4211, evp.F90:1898-1970   -> stress12U halo
//   k4  div_stress_Ex/Ny + stepu_C / stepv_C            evp.F90:2195-2248, 2364-2416, shared.F90:1090-1283
//                                                                                 -> uvelE, vvelN halo
//   k5  grid_average_X2YA E->N, N->E, E->U, N->U, masks grid.F90:4388-4606, evp.F90:1073-1090
//                                                                                 -> uvelN, vvelE, uvel, vvel halo
// The halo points themselves are the on-rank cyclic wrap: the producing thread also stores the ghost
// copies (ring_store).  One GPU, non-tripole.  Compiled twice (exact: -fmad=false, fast) like the B-grid
// kernels; every expression keeps the reference's operator order.
#include "evp_math.cuh"  // div_fast: the IEEE division fast path as straight-line code

#ifndef EVP_NS
#error "compile with -DEVP_NS=exact or -DEVP_NS=fast"
#endif

namespace evp {
namespace EVP_NS {

// sub-domains are refused at 2^31 cells (evp_b200_init), so 32-bit cell indices are enough
#define AT(i, j) ((j)*d.ld + (i))

// store val at (i,j) and at the ghost cells that alias it under an on-rank cyclic wrap.
// sides: 3 = both ghost sides (a halo update), 1 = only the E/N ghost (cells the reference computes redundantly).
// zero_closed: additionally zero the ghost neighbours outside a non-cyclic edge (arrays the reference zero-fills
// over the whole block every subcycle and whose outer ghosts no halo update touches).
__device__ __forceinline__ void ring_store(const CDom &d, double *__restrict__ A, int i, int j, double val, int sides,
                                           bool zero_closed) {
  A[AT(i, j)] = val;
  if (!(i == 1 || i == d.nx || j == 1 || j == d.ny)) return;  // only the outermost interior ring has ghost aliases
  int ig = -1, jg = -1;
  if (d.wrap_ew) ig = (i == 1) ? d.nx + 1 : ((i == d.nx && (sides & 2)) ? 0 : -1);
  if (d.wrap_ns) jg = (j == 1) ? d.ny + 1 : ((j == d.ny && (sides & 2)) ? 0 : -1);
  if (ig >= 0) A[AT(ig, j)] = val;
  if (jg >= 0) A[AT(i, jg)] = val;
  if (ig >= 0 && jg >= 0) A[AT(ig, jg)] = val;
  if (zero_closed) {
    const int zi = d.wrap_ew ? -1 : (i == 1 ? 0 : (i == d.nx ? d.nx + 1 : -1));
    const int zj = d.wrap_ns ? -1 : (j == 1 ? 0 : (j == d.ny ? d.ny + 1 : -1));
    if (zi >= 0) A[AT(zi, j)] = 0.0;
    if (zj >= 0) A[AT(i, zj)] = 0.0;
    if (zi >= 0 && zj >= 0) A[AT(zi, zj)] = 0.0;
    if (zi >= 0 && jg >= 0) A[AT(zi, jg)] = 0.0;
    if (ig >= 0 && zj >= 0) A[AT(ig, zj)] = 0.0;
  }
}

__device__ __forceinline__ void visc_c(double strength, double dmin, double Delta, const KParams &k, double &zetax2,
                                       double &etax2, double &rep_prs) {
  // capping == 1: the second term is (1-1)*(finite, >= 0) = +0 and x + 0 == x bit for bit (strength >= 0, dmin > 0), so the
  // second division is skipped -- same shortcut as the B-grid stress_point
  const double tmp = (k.capping == 1.0) ? strength / fmax(Delta, dmin)
                                        : k.capping * (strength / fmax(Delta, dmin)) + (1.0 - k.capping) * (strength / (Delta + dmin));
  zetax2 = (1.0 + k.Ktens) * tmp;
  rep_prs = (1.0 - k.Ktens) * tmp * Delta;
  etax2 = k.epp2i * zetax2;
}

#define CELL_IJ(NXE, NYE)                                          \
  const int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;         \
  const int j = 1 + blockIdx.y * blockDim.y + threadIdx.y;         \
  if (i > (NXE) || j > (NYE)) return

// Arrays the loop writes are read through M<CG>(): inside the single-launch cooperative kernel (CG = true) another SM
// wrote them a phase ago, so the load must come from L2 (ld.global.cg), never from a stale L1 line; the five-kernel
// form (CG = false) uses ordinary loads.  Static geometry and the per-step inputs stay L1-cacheable in both.
template <bool CG>
__device__ __forceinline__ double M(const double *p) { return CG ? __ldcg(p) : *p; }

// ---- k1 ------------------------------------------------------------------------------------------------
// strain rates at one U point (shared.F90:2319-2430); zero where there is no ice (the reference zero-fills)
template <bool CG = false>
__device__ __forceinline__ void strain_U_at(const CDom &d, const KParams &k, int c, double &div, double &ten, double &shr, double &del) {
  div = 0.0; ten = 0.0; shr = 0.0; del = 0.0;
  if (d.maskU[c]) {
    const int e = c + 1, n = c + d.ld;
    const double npc = d.npm[c], npe = d.npm[e], epc = d.epm[c], epn = d.epm[n];
    const double uNip1j = M<CG>(d.uvelN + e) * npe + (npc - npe) * npc * d.ratiodxN[c] * M<CG>(d.uvelN + c);
    const double uNij = M<CG>(d.uvelN + c) * npc + (npe - npc) * npe * d.ratiodxNr[c] * M<CG>(d.uvelN + e);
    const double vEijp1 = M<CG>(d.vvelE + n) * epn + (epc - epn) * epc * d.ratiodyE[c] * M<CG>(d.vvelE + c);
    const double vEij = M<CG>(d.vvelE + c) * epc + (epn - epc) * epn * d.ratiodyEr[c] * M<CG>(d.vvelE + n);
    const double dyU = d.dyU[c], dxU = d.dxU[c], uU = M<CG>(d.uvel + c), vU = M<CG>(d.vvel + c);
    const double ddyN = d.dyN[e] - d.dyN[c], ddxE = d.dxE[n] - d.dxE[c];
    div = dyU * (uNip1j - uNij) + uU * ddyN + dxU * (vEijp1 - vEij) + vU * ddxE;
    ten = dyU * (uNip1j - uNij) - uU * ddyN - dxU * (vEijp1 - vEij) + vU * ddxE;
    const double uEijp1 = M<CG>(d.uvelE + n) * epn + (epc - epn) * epc * d.ratiodyE[c] * M<CG>(d.uvelE + c);
    const double uEij = M<CG>(d.uvelE + c) * epc + (epn - epc) * epn * d.ratiodyEr[c] * M<CG>(d.uvelE + n);
    const double vNip1j = M<CG>(d.vvelN + e) * npe + (npc - npe) * npc * d.ratiodxN[c] * M<CG>(d.vvelN + c);
    const double vNij = M<CG>(d.vvelN + c) * npc + (npe - npc) * npe * d.ratiodxNr[c] * M<CG>(d.vvelN + e);
    shr = dxU * (uEijp1 - uEij) - uU * ddxE + dyU * (vNip1j - vNij) - vU * ddyN;
    del = sqrt(div * div + k.e_factor * (ten * ten + shr * shr));
  }
}
template <bool CG>
__device__ __forceinline__ void p1_strain_U(const CDom &d, const KParams &k, int i, int j) {
  if (i > d.nx || j > d.ny) return;
  const int c = AT(i, j);
  double div, ten, shr, del;
  strain_U_at<CG>(d, k, c, div, ten, shr, del);
  d.divergU[c] = div;
  d.tensionU[c] = ten;
  d.deltaU[c] = del;
  ring_store(d, d.shearU, i, j, shr, 3, false);
}

__global__ void __launch_bounds__(256) k1_strain_U(const __grid_constant__ CDom d, const __grid_constant__ KParams k) {
  p1_strain_U<false>(d, k, 1 + blockIdx.x * blockDim.x + threadIdx.x, 1 + blockIdx.y * blockDim.y + threadIdx.y);
}

// ---- k2 ------------------------------------------------------------------------------------------------
// T cells 1..nx+1 x 1..ny+1 (N/E ghost cells are in the reference's list, shared.F90:740-749); a ghost row or
// column that merely aliases the interior under the on-rank wrap is filled by ring_store instead.
template <bool CG = false>
__device__ __forceinline__ void stress_T_at(const CDom &d, const KParams &k, int i, int j, int c, double shc, double shs, double shsw,
                                            double shw) {
  const int w = c - 1, s = c - d.ld, sw = s - 1;
  const double dyEc = d.dyE[c], dyEw = d.dyE[w], dxNc = d.dxN[c], dxNs = d.dxN[s];
  const double uEc = M<CG>(d.uvelE + c), uEw = M<CG>(d.uvelE + w), vNc = M<CG>(d.vvelN + c), vNs = M<CG>(d.vvelN + s);
  const double dxT = d.dxT[c], dyT = d.dyT[c];
  const double divT = dyEc * uEc - dyEw * uEw + dxNc * vNc - dxNs * vNs;
  // four independent quotients: fast paths first (interleaved by the scheduler), the rare out-of-range operand afterwards
  bool ok0, ok1, ok2, ok3;
  double q0 = div_fast(uEc, dyEc, ok0), q1 = div_fast(uEw, dyEw, ok1), q2 = div_fast(vNc, dxNc, ok2), q3 = div_fast(vNs, dxNs, ok3);
  if (!(ok0 && ok1 && ok2 && ok3)) {
    if (!ok0) q0 = uEc / dyEc;
    if (!ok1) q1 = uEw / dyEw;
    if (!ok2) q2 = vNc / dxNc;
    if (!ok3) q3 = vNs / dxNs;
  }
  const double tensionT = (dyT * dyT) * (q0 - q1) - (dxT * dxT) * (q2 - q3);
  const double uac = d.uarea[c], uas = d.uarea[s], uasw = d.uarea[sw], uaw = d.uarea[w];
  const double uareaavgr = d.uareaavgr[c];  // 1.0 / (uac + uas + uasw + uaw), static: divided once in evp_b200_init_cgrid
  const double shearTsqr = (shc * shc * uac + shs * shs * uas + shsw * shsw * uasw + shw * shw * uaw) * uareaavgr;
  const double shearT = (shc * uac + shs * uas + shsw * uasw + shw * uaw) * uareaavgr;
  const double DeltaT = sqrt(divT * divT + k.e_factor * (tensionT * tensionT + shearTsqr));
  double zetax2, etax2, rep;
  visc_c(d.strength[c], d.DminTarea[c], DeltaT, k, zetax2, etax2, rep);
  const double relax = 1.0 - k.arlx1i * k.revp;
  const double sp = (M<CG>(d.stresspT + c) * relax + k.arlx1i * (zetax2 * divT - rep)) * k.denom1;
  const double sm = (M<CG>(d.stressmT + c) * relax + k.arlx1i * etax2 * tensionT) * k.denom1;
  const double s12 = (M<CG>(d.stress12T + c) * relax + k.arlx1i * 0.5 * etax2 * shearT) * k.denom1;
  ring_store(d, d.zetax2T, i, j, zetax2, 3, false);
  ring_store(d, d.etax2T, i, j, etax2, 3, false);
  ring_store(d, d.stresspT, i, j, sp, 3, false);
  ring_store(d, d.stressmT, i, j, sm, 3, false);
  ring_store(d, d.stress12T, i, j, s12, 1, false);  // not halo-updated: only the redundantly computed N/E ghost copy
}
template <bool CG>
__device__ __forceinline__ void p2_stress_T(const CDom &d, const KParams &k, int i, int j) {
  if (i > (d.wrap_ew ? d.nx : d.nx + 1) || j > (d.wrap_ns ? d.ny : d.ny + 1)) return;
  const int c = AT(i, j);
  if (!d.maskT[c]) return;
  const int w = c - 1, s = c - d.ld, sw = s - 1;
  stress_T_at<CG>(d, k, i, j, c, M<CG>(d.shearU + c), M<CG>(d.shearU + s), M<CG>(d.shearU + sw), M<CG>(d.shearU + w));
}

__global__ void __launch_bounds__(256) k2_stress_T(const __grid_constant__ CDom d, const __grid_constant__ KParams k) {
  p2_stress_T<false>(d, k, 1 + blockIdx.x * blockDim.x + threadIdx.x, 1 + blockIdx.y * blockDim.y + threadIdx.y);
}

// ---- k3 ------------------------------------------------------------------------------------------------
// stress12U at one U point: returns whether the point is updated (ice), `avg` = the T->U average it stores either way
template <bool CG = false>
__device__ __forceinline__ bool stress_U_at(const CDom &d, const KParams &k, int c, const double *s12old, double &avg, double &s12) {
  const int e = c + 1, n = c + d.ld, ne = n + 1;
  const double *__restrict__ src = (k.visc_method == 1) ? d.strength : d.etax2T;
  const double mc = d.hm[c], me = d.hm[e], mn = d.hm[n], mne = d.hm[ne];
  const double wc = d.tarea[c], we = d.tarea[e], wn = d.tarea[n], wne = d.tarea[ne];
  const double wtmp = (mc * wc + me * we + mn * wn + mne * wne);
  avg = 0.0;
  if (wtmp != 0.0) avg = (mc * M<CG>(src + c) * wc + me * M<CG>(src + e) * we + mn * M<CG>(src + n) * wn + mne * M<CG>(src + ne) * wne) / wtmp;
  s12 = M<CG>(s12old + c);
  if (!d.maskU[c]) return false;
  const double relax = 1.0 - k.arlx1i * k.revp;
  double etax2U = avg;
  if (k.visc_method == 1) {
    const double DminUarea = k.deltaminEVP * d.uarea[c];
    double z, r;
    visc_c(avg, DminUarea, M<CG>(d.deltaU + c), k, z, etax2U, r);
  }
  s12 = (s12 * relax + k.arlx1i * 0.5 * etax2U * M<CG>(d.shearU + c)) * k.denom1;
  return true;
}
template <bool CG>
__device__ __forceinline__ void p3_stress_U(const CDom &d, const KParams &k, int i, int j) {
  if (i > d.nx || j > d.ny) return;
  const int c = AT(i, j);
  double avg, s12;
  const bool upd = stress_U_at<CG>(d, k, c, d.stress12U, avg, s12);
  if (k.visc_method == 1) d.strengthU[c] = avg; else d.etax2U[c] = avg;
  if (upd) ring_store(d, d.stress12U, i, j, s12, 3, false);
}

__global__ void __launch_bounds__(256) k3_stress_U(const __grid_constant__ CDom d, const __grid_constant__ KParams k) {
  p3_stress_U<false>(d, k, 1 + blockIdx.x * blockDim.x + threadIdx.x, 1 + blockIdx.y * blockDim.y + threadIdx.y);
}

// ---- k4 ------------------------------------------------------------------------------------------------
template <bool CG = false>
__device__ __forceinline__ void momentum_at(const CDom &d, const KParams &k, int i, int j, int c, double s12c, double s12s, double s12w) {
  if (d.maskE[c]) {
    const int e = c + 1, s = c - d.ld;
    const double dyE = d.dyE[c], dyTe = d.dyT[e], dyTc = d.dyT[c], dxUc = d.dxU[c], dxUs = d.dxU[s];
    const double strintx = d.rheofactE[c] * d.earear[c] *
                           (0.5 * dyE * (M<CG>(d.stresspT + e) - M<CG>(d.stresspT + c)) +
                            d.rhalf_dyE[c] * ((dyTe * dyTe) * M<CG>(d.stressmT + e) - (dyTc * dyTc) * M<CG>(d.stressmT + c)) +
                            d.r_dxE[c] * ((dxUc * dxUc) * s12c - (dxUs * dxUs) * s12s));
    d.strintxE[c] = strintx;
    const double uold = M<CG>(d.uvelE + c), vold = M<CG>(d.vvelE + c);
    const double du = d.uocnE[c] - uold, dv = d.vocnE[c] - vold;
    const double vrel = d.aiE[c] * k.rhow * d.cdnE[c] * sqrt(du * du + dv * dv);
    const double taux = vrel * d.waterxE[c];
    // no grounded ice: Tb is (+-)0 and (+-)0 / (finite positive) is that same zero (see stepu_point)
    const double Tb = d.TbE[c];
    const double Cb = (Tb == 0.0 && k.u0 > 0.0) ? Tb : Tb / (sqrt(uold * uold + vold * vold) + k.u0);
    const double m = d.emassdti[c], fm = d.fmE[c];
    const double cca = (k.brlx + k.revp) * m + vrel * k.cosw + Cb;
    const double ccb = fm + copysign(1.0, fm) * vrel * k.sinw;
    const double cc1 = strintx + d.forcexE[c] + taux + m * (k.brlx * uold + k.revp * d.uvelE_init[c]);
    const double un = (ccb * vold + cc1) / cca;
    ring_store(d, d.uvelE, i, j, un, 3, false);
    d.taubxE[c] = -un * Cb;
  }
  if (d.maskN[c]) {
    const int n = c + d.ld, w = c - 1;
    const double dxN = d.dxN[c], dxTn = d.dxT[n], dxTc = d.dxT[c], dyUc = d.dyU[c], dyUw = d.dyU[w];
    const double strinty = d.rheofactN[c] * d.narear[c] *
                           (0.5 * dxN * (M<CG>(d.stresspT + n) - M<CG>(d.stresspT + c)) -
                            d.rhalf_dxN[c] * ((dxTn * dxTn) * M<CG>(d.stressmT + n) - (dxTc * dxTc) * M<CG>(d.stressmT + c)) +
                            d.r_dyN[c] * ((dyUc * dyUc) * s12c - (dyUw * dyUw) * s12w));
    d.strintyN[c] = strinty;
    const double uold = M<CG>(d.uvelN + c), vold = M<CG>(d.vvelN + c);
    const double du = d.uocnN[c] - uold, dv = d.vocnN[c] - vold;
    const double vrel = d.aiN[c] * k.rhow * d.cdnN[c] * sqrt(du * du + dv * dv);
    const double tauy = vrel * d.wateryN[c];
    // no grounded ice: Tb is (+-)0 and (+-)0 / (finite positive) is that same zero (see stepu_point)
    const double Tb = d.TbN[c];
    const double Cb = (Tb == 0.0 && k.u0 > 0.0) ? Tb : Tb / (sqrt(uold * uold + vold * vold) + k.u0);
    const double m = d.nmassdti[c], fm = d.fmN[c];
    const double cca = (k.brlx + k.revp) * m + vrel * k.cosw + Cb;
    const double ccb = fm + copysign(1.0, fm) * vrel * k.sinw;
    const double cc2 = strinty + d.forceyN[c] + tauy + m * (k.brlx * vold + k.revp * d.vvelN_init[c]);
    const double vn = (-ccb * uold + cc2) / cca;
    ring_store(d, d.vvelN, i, j, vn, 3, false);
    d.taubyN[c] = -vn * Cb;
  }
}

// momentum_at with the two square roots and the two divisions of a point (E and N) taken together through sqrt_fast / div_fast, so
// that the four ~100-cycle chains interleave instead of running one after the other behind two separate mask branches.  Same
// expressions, same bits (the fast paths are the built-in operators' own sequences, evp_math.cuh); a masked-off half runs on
// harmless in-range dummies.  The form kB uses (measured 12.46 vs 12.52 ms per step at gx1).
__device__ __forceinline__ void momentum_il_at(const CDom &d, const KParams &k, int i, int j, int c, double s12c, double s12s, double s12w) {
  const bool mE = d.maskE[c] != 0, mN = d.maskN[c] != 0;
  if (!(mE || mN)) return;
  double strintx = 0.0, uoE = 0.0, voE = 0.0, xE = 1.0, strinty = 0.0, uoN = 0.0, voN = 0.0, xN = 1.0;
  if (mE) {
    const int e = c + 1, s = c - d.ld;
    const double dyE = d.dyE[c], dyTe = d.dyT[e], dyTc = d.dyT[c], dxUc = d.dxU[c], dxUs = d.dxU[s];
    strintx = d.rheofactE[c] * d.earear[c] *
              (0.5 * dyE * (d.stresspT[e] - d.stresspT[c]) +
               d.rhalf_dyE[c] * ((dyTe * dyTe) * d.stressmT[e] - (dyTc * dyTc) * d.stressmT[c]) +
               d.r_dxE[c] * ((dxUc * dxUc) * s12c - (dxUs * dxUs) * s12s));
    d.strintxE[c] = strintx;
    uoE = d.uvelE[c]; voE = d.vvelE[c];
    const double du = d.uocnE[c] - uoE, dv = d.vocnE[c] - voE;
    xE = du * du + dv * dv;
  }
  if (mN) {
    const int n = c + d.ld, w = c - 1;
    const double dxN = d.dxN[c], dxTn = d.dxT[n], dxTc = d.dxT[c], dyUc = d.dyU[c], dyUw = d.dyU[w];
    strinty = d.rheofactN[c] * d.narear[c] *
              (0.5 * dxN * (d.stresspT[n] - d.stresspT[c]) -
               d.rhalf_dxN[c] * ((dxTn * dxTn) * d.stressmT[n] - (dxTc * dxTc) * d.stressmT[c]) +
               d.r_dyN[c] * ((dyUc * dyUc) * s12c - (dyUw * dyUw) * s12w));
    d.strintyN[c] = strinty;
    uoN = d.uvelN[c]; voN = d.vvelN[c];
    const double du = d.uocnN[c] - uoN, dv = d.vocnN[c] - voN;
    xN = du * du + dv * dv;
  }
  bool okE, okN;
  double sE = sqrt_fast(xE, okE), sN = sqrt_fast(xN, okN);
  if (!(okE && okN)) {
    if (!okE) sE = sqrt_ieee(xE);
    if (!okN) sN = sqrt_ieee(xN);
  }
  double numE = 1.0, denE = 1.0, CbE = 0.0, numN = 1.0, denN = 1.0, CbN = 0.0;
  if (mE) {
    const double vrel = d.aiE[c] * k.rhow * d.cdnE[c] * sE;
    const double taux = vrel * d.waterxE[c];
    const double Tb = d.TbE[c];
    CbE = (Tb == 0.0 && k.u0 > 0.0) ? Tb : Tb / (sqrt(uoE * uoE + voE * voE) + k.u0);
    const double m = d.emassdti[c], fm = d.fmE[c];
    const double cca = (k.brlx + k.revp) * m + vrel * k.cosw + CbE;
    const double ccb = fm + copysign(1.0, fm) * vrel * k.sinw;
    const double cc1 = strintx + d.forcexE[c] + taux + m * (k.brlx * uoE + k.revp * d.uvelE_init[c]);
    numE = ccb * voE + cc1;
    denE = cca;
  }
  if (mN) {
    const double vrel = d.aiN[c] * k.rhow * d.cdnN[c] * sN;
    const double tauy = vrel * d.wateryN[c];
    const double Tb = d.TbN[c];
    CbN = (Tb == 0.0 && k.u0 > 0.0) ? Tb : Tb / (sqrt(uoN * uoN + voN * voN) + k.u0);
    const double m = d.nmassdti[c], fm = d.fmN[c];
    const double cca = (k.brlx + k.revp) * m + vrel * k.cosw + CbN;
    const double ccb = fm + copysign(1.0, fm) * vrel * k.sinw;
    const double cc2 = strinty + d.forceyN[c] + tauy + m * (k.brlx * voN + k.revp * d.vvelN_init[c]);
    numN = -ccb * uoN + cc2;
    denN = cca;
  }
  bool okdE, okdN;
  double un = div_fast(numE, denE, okdE), vn = div_fast(numN, denN, okdN);
  if (!(okdE && okdN)) {
    if (!okdE) un = div_ieee(numE, denE);
    if (!okdN) vn = div_ieee(numN, denN);
  }
  if (mE) {
    ring_store(d, d.uvelE, i, j, un, 3, false);
    d.taubxE[c] = -un * CbE;
  }
  if (mN) {
    ring_store(d, d.vvelN, i, j, vn, 3, false);
    d.taubyN[c] = -vn * CbN;
  }
}

template <bool CG>
__device__ __forceinline__ void p4_momentum(const CDom &d, const KParams &k, int i, int j) {
  if (i > d.nx || j > d.ny) return;
  const int c = AT(i, j);
  momentum_at<CG>(d, k, i, j, c, M<CG>(d.stress12U + c), M<CG>(d.stress12U + c - d.ld), M<CG>(d.stress12U + c - 1));
}

__global__ void __launch_bounds__(256) k4_momentum(const __grid_constant__ CDom d, const __grid_constant__ KParams k) {
  p4_momentum<false>(d, k, 1 + blockIdx.x * blockDim.x + threadIdx.x, 1 + blockIdx.y * blockDim.y + threadIdx.y);
}

// ---- k5 ------------------------------------------------------------------------------------------------
// grid_average_X2YA (ice_grid.F90:4388-4606): area-weighted average, zero where the weights vanish.  The numerator and
// the denominator are formed here, the four quotients of a point are then taken together (div_fast) in p5_interp.
template <bool CG>
__device__ __forceinline__ void avg4_terms(const double *w1, const double *__restrict__ wg, int a, int b, int c, int e, double &num,
                                           double &den) {
  const double wa = wg[a], wb = wg[b], wc = wg[c], we = wg[e];
  den = (wa + wb + wc + we);
  num = (M<CG>(w1 + a) * wa + M<CG>(w1 + b) * wb + M<CG>(w1 + c) * wc + M<CG>(w1 + e) * we);
}
template <bool CG>
__device__ __forceinline__ void avg2_terms(const double *w1, const double *__restrict__ wg, int a, int b, double &num, double &den) {
  const double wa = wg[a], wb = wg[b];
  den = (wa + wb);
  num = (M<CG>(w1 + a) * wa + M<CG>(w1 + b) * wb);
}
template <bool CG>
__device__ __forceinline__ void p5_interp(const CDom &d, int i, int j) {
  if (i > d.nx || j > d.ny) return;
  const int c = AT(i, j);
  const int e = c + 1, w = c - 1, n = c + d.ld, s = c - d.ld;
  double num[4], den[4], q[4];
  bool ok[4];
  avg4_terms<CG>(d.uvelE, d.earea, w, c, n - 1, n, num[0], den[0]);      // E2NA 'NW'
  avg4_terms<CG>(d.vvelN, d.narea, s, s + 1, c, e, num[1], den[1]);      // N2EA 'SE'
  avg2_terms<CG>(d.uvelE, d.earea, c, n, num[2], den[2]);                // E2UA 'N'
  avg2_terms<CG>(d.vvelN, d.narea, c, e, num[3], den[3]);                // N2UA 'E'
#pragma unroll
  for (int r = 0; r < 4; ++r) q[r] = div_fast(num[r], den[r], ok[r]);
  if (!(ok[0] && ok[1] && ok[2] && ok[3])) {
#pragma unroll
    for (int r = 0; r < 4; ++r) if (!ok[r]) q[r] = num[r] / den[r];
  }
  const double uN = ((den[0] != 0.0) ? q[0] : 0.0) * d.npm[c];
  const double vE = ((den[1] != 0.0) ? q[1] : 0.0) * d.epm[c];
  const double uU = ((den[2] != 0.0) ? q[2] : 0.0) * d.uvm[c];
  const double vU = ((den[3] != 0.0) ? q[3] : 0.0) * d.uvm[c];
  ring_store(d, d.uvelN, i, j, uN, 3, true);
  ring_store(d, d.vvelE, i, j, vE, 3, true);
  ring_store(d, d.uvel, i, j, uU, 3, true);
  ring_store(d, d.vvel, i, j, vU, 3, true);
}
__global__ void __launch_bounds__(256) k5_interp(const __grid_constant__ CDom d) {
  p5_interp<false>(d, 1 + blockIdx.x * blockDim.x + threadIdx.x, 1 + blockIdx.y * blockDim.y + threadIdx.y);
}

// ---- fused forms: three kernels per subcycle instead of five ------------------------------------------------
// kA = k1 + k2 and kB = k3 + k4 on overlapping 32 x 8 tiles, the way the B-grid fused kernel joins stress and stepu:
// thread (tx,ty) of a tile holds point (i,j) = (ti-1+tx, tj-1+ty); the first phase runs on all 32 x 8 points (the
// west column and south row belong to the neighbouring tiles and are recomputed, not stored), hands its result to the
// second phase through shared memory, and the second phase runs on the 31 x 7 points with tx,ty >= 1, which are the
// points the tile owns.  A ghost point that aliases the interior under an on-rank cyclic wrap is recomputed at the
// interior point it aliases; a ghost point outside a closed/open edge is never written by the loop and is read back.
// kB's first phase updates stress12U in place (new = f(old)), so the recomputed copies would race with the owner's
// store: stress12U is ping-ponged between two arrays (both start identical; cells off the ice are never written).
constexpr int GBX = 32;

// index of the interior point a ring point aliases (wrap), or -1 when the point is a constant ghost / outside
__device__ __forceinline__ bool alias_point(const CDom &d, int &i, int &j) {
  if (i < 1) { if (!d.wrap_ew) return false; i += d.nx; }
  if (j < 1) { if (!d.wrap_ns) return false; j += d.ny; }
  if (i > d.nx) { if (!d.wrap_ew) return false; i -= d.nx; }
  if (j > d.ny) { if (!d.wrap_ns) return false; j -= d.ny; }
  return i >= 1 && i <= d.nx && j >= 1 && j <= d.ny;
}

template <int GBY, int MINB>
__global__ void __launch_bounds__(GBX *GBY, MINB) kA_strainU_stressT(const __grid_constant__ CDom d, const __grid_constant__ KParams k) {
  __shared__ double sh[GBY][GBX];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i = blockIdx.x * (GBX - 1) + tx, j = blockIdx.y * (GBY - 1) + ty;  // ti = 1 + bx*(GBX-1), point ti-1+tx
  const int nxT = d.wrap_ew ? d.nx : d.nx + 1, nyT = d.wrap_ns ? d.ny : d.ny + 1;
  // phase 1: shearU (and divergU, tensionU, deltaU) at the U point (i,j)
  double shr = 0.0;
  if (i <= d.nx + 1 && j <= d.ny + 1) {
    int ia = i, ja = j;
    if (alias_point(d, ia, ja)) {
      double div, ten, del;
      strain_U_at(d, k, AT(ia, ja), div, ten, shr, del);
      if (tx >= 1 && ty >= 1 && ia == i && ja == j) {  // an interior point this tile owns
        const int c = AT(i, j);
        d.divergU[c] = div;
        d.tensionU[c] = ten;
        d.deltaU[c] = del;
        ring_store(d, d.shearU, i, j, shr, 3, false);
      }
    } else {
      shr = d.shearU[AT(i, j)];  // ghost outside a non-cyclic edge: the loop never writes it
    }
  }
  sh[ty][tx] = shr;
  __syncthreads();
  // phase 2: the T cell (i,j); its corners are the U points (i,j) (i-1,j) (i,j-1) (i-1,j-1)
  if (tx >= 1 && ty >= 1 && i <= nxT && j <= nyT) {
    const int c = AT(i, j);
    if (d.maskT[c]) stress_T_at(d, k, i, j, c, sh[ty][tx], sh[ty - 1][tx], sh[ty - 1][tx - 1], sh[ty][tx - 1]);
  }
}

template <int GBY, int MINB, bool ILM>
__global__ void __launch_bounds__(GBX *GBY, MINB) kB_stressU_momentum(const __grid_constant__ CDom d, const __grid_constant__ KParams k, int cur) {
  __shared__ double sh[GBY][GBX];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i = blockIdx.x * (GBX - 1) + tx, j = blockIdx.y * (GBY - 1) + ty;
  const double *__restrict__ s12old = cur ? d.stress12Ub : d.stress12U;
  double *__restrict__ s12new = cur ? d.stress12U : d.stress12Ub;
  // phase 1: stress12U (and etax2U / strengthU) at the U point (i,j)
  double s12 = 0.0;
  if (i <= d.nx + 1 && j <= d.ny + 1) {
    int ia = i, ja = j;
    if (alias_point(d, ia, ja)) {
      const int ca = AT(ia, ja);
      double avg;
      const bool upd = stress_U_at(d, k, ca, s12old, avg, s12);
      if (tx >= 1 && ty >= 1 && ia == i && ja == j) {
        if (k.visc_method == 1) d.strengthU[ca] = avg; else d.etax2U[ca] = avg;
        if (upd) ring_store(d, s12new, i, j, s12, 3, false);
      }
    } else {
      s12 = s12old[AT(i, j)];
    }
  }
  sh[ty][tx] = s12;
  __syncthreads();
  // phase 2: the E and N points (i,j)
  if (tx >= 1 && ty >= 1 && i <= d.nx && j <= d.ny) {
    if (ILM) momentum_il_at(d, k, i, j, AT(i, j), sh[ty][tx], sh[ty - 1][tx], sh[ty][tx - 1]);
    else momentum_at(d, k, i, j, AT(i, j), sh[ty][tx], sh[ty - 1][tx], sh[ty][tx - 1]);
  }
}

// quotients of static geometry that the reference re-divides every subcycle (ice_dyn_evp.F90:2230-2240, 2398-2408,
// ice_dyn_shared.F90:2296): an IEEE division of the same operands gives the same bits whenever it is done, so they are
// divided once here
__global__ void cgrid_static_quotients(const __grid_constant__ CDom d, double *rhalf_dyE, double *r_dxE, double *rhalf_dxN, double *r_dyN,
                                       double *uareaavgr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i > d.nx + 1 || j > d.ny + 1) return;
  const int c = AT(i, j);
  rhalf_dyE[c] = 0.5 / d.dyE[c];
  r_dxE[c] = 1.0 / d.dxE[c];
  rhalf_dxN[c] = 0.5 / d.dxN[c];
  r_dyN[c] = 1.0 / d.dyN[c];
  if (i >= 1 && j >= 1) uareaavgr[c] = 1.0 / (d.uarea[c] + d.uarea[c - d.ld] + d.uarea[c - d.ld - 1] + d.uarea[c - 1]);
}
#ifndef EVP_HOST_EMU  // launchers (and the cooperative kernel): not part of the host emulation (tests/emu_cgrid.cpp)
cudaError_t launch_cgrid_static(const CDom &d, double *rhalf_dyE, double *r_dxE, double *rhalf_dxN, double *r_dyN, double *uareaavgr,
                                cudaStream_t s) {
  dim3 b(32, 8), g((d.nx + 2 + 31) / 32, (d.ny + 2 + 7) / 8);
  cgrid_static_quotients<<<g, b, 0, s>>>(d, rhalf_dyE, r_dxE, rhalf_dxN, r_dyN, uareaavgr);
  return cudaGetLastError();
}
#endif  // EVP_HOST_EMU

#ifndef EVP_HOST_EMU  // launchers: not part of the host emulation (tests/emu_cgrid.cpp)
// 32 x 8 tiles, 4 CTAs per SM (62 registers); 32x{4,12,16} tiles measured within 4 % of it, chaining the three launches by
// programmatic dependent launch and a single cooperative launch with grid barriers both measured slower (profiles/)
cudaError_t launch_cgrid_subcycle_fused(const CDom &d, const KParams &p, int cur, cudaStream_t s, int *launches) {
  dim3 b(GBX, 8), g((d.nx + 1 + GBX - 2) / (GBX - 1), (d.ny + 1 + 8 - 2) / (8 - 1));
  dim3 b5(32, 8), g5((d.nx + 31) / 32, (d.ny + 7) / 8);
  kA_strainU_stressT<8, 4><<<g, b, 0, s>>>(d, p);
  kB_stressU_momentum<8, 4, true><<<g, b, 0, s>>>(d, p, cur);
  k5_interp<<<g5, b5, 0, s>>>(d);
  *launches += 3;
  return cudaGetLastError();
}
#endif  // EVP_HOST_EMU

// =============================================================================================================
// CD grid (grid_ice = 'CD'): one subcycle of ice_dyn_evp.F90:1125-1267 as four kernels, cut where a value
// is needed at a neighbour point:
//   kcd1  strain_rates_Tdtsd + stressCD_T                         shared.F90:2171-2243, evp.F90:1978-2080
//   kcd2  T->U averages (zetax2U, etax2U | strengthU), strain_rates_U, stressCD_U     grid.F90:4159-4211, evp.F90:2088-2178
//   kcd3  div_stress_Ex/Ey/Nx/Ny + stepuv_CD at E and at N        evp.F90:2195-2416, shared.F90:973-1085
//   kcd4  grid_average_X2YA E->U, N->U, times uvm                  grid.F90:4388-4606, evp.F90:1254-1259
// Halo points are the producing thread's wrap stores (ring_store), as on the C grid.  First correct form: one thread per
// point, plain IEEE division and square root, no fusion.  Bit-identical to the oracle and to the reference-source vectors.
// =============================================================================================================
__device__ __forceinline__ double t2u_S(const CDom &d, const double *__restrict__ src, int c) {   // grid_average_X2YS 'NE'
  const int e = c + 1, n = c + d.ld, ne = n + 1;
  const double mc = d.hm[c], me = d.hm[e], mn = d.hm[n], mne = d.hm[ne];
  const double wc = d.tarea[c], we = d.tarea[e], wn = d.tarea[n], wne = d.tarea[ne];
  const double wtmp = (mc * wc + me * we + mn * wn + mne * wne);
  if (wtmp == 0.0) return 0.0;
  return (mc * src[c] * wc + me * src[e] * we + mn * src[n] * wn + mne * src[ne] * wne) / wtmp;
}

__global__ void __launch_bounds__(256) kcd1_stress_T(const __grid_constant__ CDom d, const __grid_constant__ KParams k) {
  CELL_IJ(d.wrap_ew ? d.nx : d.nx + 1, d.wrap_ns ? d.ny : d.ny + 1);
  const int c = AT(i, j);
  if (!d.maskT[c]) return;
  const int w = c - 1, s = c - d.ld;
  const double dyEc = d.dyE[c], dyEw = d.dyE[w], dxNc = d.dxN[c], dxNs = d.dxN[s], dxT = d.dxT[c], dyT = d.dyT[c];
  const double uEc = d.uvelE[c], uEw = d.uvelE[w], vNc = d.vvelN[c], vNs = d.vvelN[s];
  const double divT = dyEc * uEc - dyEw * uEw + dxNc * vNc - dxNs * vNs;
  const double tensionT = (dyT * dyT) * (uEc / dyEc - uEw / dyEw) - (dxT * dxT) * (vNc / dxNc - vNs / dxNs);
  const double shearT = (dxT * dxT) * (d.uvelN[c] / dxNc - d.uvelN[s] / dxNs) + (dyT * dyT) * (d.vvelE[c] / dyEc - d.vvelE[w] / dyEw);
  const double DeltaT = sqrt(divT * divT + k.e_factor * (tensionT * tensionT + shearT * shearT));
  double zetax2, etax2, rep;
  visc_c(d.strength[c], d.DminTarea[c], DeltaT, k, zetax2, etax2, rep);
  const double relax = 1.0 - k.arlx1i * k.revp;
  const double sp = (d.stresspT[c] * relax + k.arlx1i * (zetax2 * divT - rep)) * k.denom1;
  const double sm = (d.stressmT[c] * relax + k.arlx1i * etax2 * tensionT) * k.denom1;
  const double s12 = (d.stress12T[c] * relax + k.arlx1i * 0.5 * etax2 * shearT) * k.denom1;
  ring_store(d, d.zetax2T, i, j, zetax2, 3, false);
  ring_store(d, d.etax2T, i, j, etax2, 3, false);
  ring_store(d, d.stresspT, i, j, sp, 3, false);
  ring_store(d, d.stressmT, i, j, sm, 3, false);
  ring_store(d, d.stress12T, i, j, s12, 3, false);
}

__global__ void __launch_bounds__(256) kcd2_stress_U(const __grid_constant__ CDom d, const __grid_constant__ KParams k) {
  CELL_IJ(d.nx, d.ny);
  const int c = AT(i, j);
  double zetax2U = 0.0, etax2U = 0.0, strengthU = 0.0;
  if (k.visc_method == 1) {
    strengthU = t2u_S(d, d.strength, c);
    d.strengthU[c] = strengthU;
  } else {
    zetax2U = t2u_S(d, d.zetax2T, c);
    etax2U = t2u_S(d, d.etax2T, c);
    d.zetax2U[c] = zetax2U;
    d.etax2U[c] = etax2U;
  }
  double div, ten, shr, del;
  strain_U_at(d, k, c, div, ten, shr, del);
  d.divergU[c] = div;
  d.tensionU[c] = ten;
  d.shearU[c] = shr;
  d.deltaU[c] = del;
  if (!d.maskU[c]) return;
  double lzetax2U, letax2U, lrep_prsU;
  if (k.visc_method == 1) {
    const double DminUarea = k.deltaminEVP * d.uarea[c];
    visc_c(strengthU, DminUarea, del, k, lzetax2U, letax2U, lrep_prsU);
  } else {
    lzetax2U = zetax2U;
    letax2U = etax2U;
    lrep_prsU = (1.0 - k.Ktens) / (1.0 + k.Ktens) * lzetax2U * del;
  }
  const double relax = 1.0 - k.arlx1i * k.revp;
  const double sp = (d.stresspU[c] * relax + k.arlx1i * (lzetax2U * div - lrep_prsU)) * k.denom1;
  const double sm = (d.stressmU[c] * relax + k.arlx1i * letax2U * ten) * k.denom1;
  const double s12 = (d.stress12U[c] * relax + k.arlx1i * 0.5 * letax2U * shr) * k.denom1;
  ring_store(d, d.stresspU, i, j, sp, 3, false);
  ring_store(d, d.stressmU, i, j, sm, 3, false);
  ring_store(d, d.stress12U, i, j, s12, 3, false);
}

// stepuv_CD at one point (shared.F90:1040-1078); returns the new velocity, stores the seabed stress components
__device__ __forceinline__ void stepuv_cd_at(const KParams &k, double uold, double vold, double ai, double Cw, double uocn, double vocn,
                                             double waterx, double watery, double forcex, double forcey, double massdti, double fm,
                                             double strintx, double strinty, double uinit, double vinit, double Tb, double &un,
                                             double &vn, double &taubx, double &tauby) {
  const double du = uocn - uold, dv = vocn - vold;
  const double vrel = ai * k.rhow * Cw * sqrt(du * du + dv * dv);
  const double taux = vrel * waterx;
  const double tauy = vrel * watery;
  // no grounded ice: Tb is (+-)0 and (+-)0 / (finite positive) is that same zero (see stepu_point)
  const double Cb = (Tb == 0.0 && k.u0 > 0.0) ? Tb : Tb / (sqrt(uold * uold + vold * vold) + k.u0);
  const double cca = (k.brlx + k.revp) * massdti + vrel * k.cosw + Cb;
  const double ccb = fm + copysign(1.0, fm) * vrel * k.sinw;
  const double ab2 = cca * cca + ccb * ccb;
  const double cc1 = strintx + forcex + taux + massdti * (k.brlx * uold + k.revp * uinit);
  const double cc2 = strinty + forcey + tauy + massdti * (k.brlx * vold + k.revp * vinit);
  un = (cca * cc1 + ccb * cc2) / ab2;
  vn = (cca * cc2 - ccb * cc1) / ab2;
  taubx = -un * Cb;
  tauby = -vn * Cb;
}

__global__ void __launch_bounds__(256) kcd3_momentum(const __grid_constant__ CDom d, const __grid_constant__ KParams k) {
  CELL_IJ(d.nx, d.ny);
  const int c = AT(i, j), e = c + 1, n = c + d.ld, s = c - d.ld, w = c - 1;
  if (d.maskE[c]) {
    const double dyE = d.dyE[c], dxE = d.dxE[c], dyTe = d.dyT[e], dyTc = d.dyT[c], dxUc = d.dxU[c], dxUs = d.dxU[s];
    const double fac = d.rheofactE[c] * d.earear[c];
    const double strintx = fac * (0.5 * dyE * (d.stresspT[e] - d.stresspT[c]) +
                                  d.rhalf_dyE[c] * ((dyTe * dyTe) * d.stressmT[e] - (dyTc * dyTc) * d.stressmT[c]) +
                                  d.r_dxE[c] * ((dxUc * dxUc) * d.stress12U[c] - (dxUs * dxUs) * d.stress12U[s]));
    const double strinty = fac * (0.5 * dxE * (d.stresspU[c] - d.stresspU[s]) -
                                  (0.5 / dxE) * ((dxUc * dxUc) * d.stressmU[c] - (dxUs * dxUs) * d.stressmU[s]) +
                                  (1.0 / dyE) * ((dyTe * dyTe) * d.stress12T[e] - (dyTc * dyTc) * d.stress12T[c]));
    d.strintxE[c] = strintx;
    d.strintyE[c] = strinty;
    double un, vn, tbx, tby;
    stepuv_cd_at(k, d.uvelE[c], d.vvelE[c], d.aiE[c], d.cdnE[c], d.uocnE[c], d.vocnE[c], d.waterxE[c], d.wateryE[c], d.forcexE[c],
                 d.forceyE[c], d.emassdti[c], d.fmE[c], strintx, strinty, d.uvelE_init[c], d.vvelE_init[c], d.TbE[c], un, vn, tbx, tby);
    ring_store(d, d.uvelE, i, j, un, 3, false);
    ring_store(d, d.vvelE, i, j, vn, 3, false);
    d.taubxE[c] = tbx;
    d.taubyE[c] = tby;
  }
  if (d.maskN[c]) {
    const double dxN = d.dxN[c], dyN = d.dyN[c], dxTn = d.dxT[n], dxTc = d.dxT[c], dyUc = d.dyU[c], dyUw = d.dyU[w];
    const double fac = d.rheofactN[c] * d.narear[c];
    const double strintx = fac * (0.5 * dyN * (d.stresspU[c] - d.stresspU[w]) +
                                  (0.5 / dyN) * ((dyUc * dyUc) * d.stressmU[c] - (dyUw * dyUw) * d.stressmU[w]) +
                                  (1.0 / dxN) * ((dxTn * dxTn) * d.stress12T[n] - (dxTc * dxTc) * d.stress12T[c]));
    const double strinty = fac * (0.5 * dxN * (d.stresspT[n] - d.stresspT[c]) -
                                  d.rhalf_dxN[c] * ((dxTn * dxTn) * d.stressmT[n] - (dxTc * dxTc) * d.stressmT[c]) +
                                  d.r_dyN[c] * ((dyUc * dyUc) * d.stress12U[c] - (dyUw * dyUw) * d.stress12U[w]));
    d.strintxN[c] = strintx;
    d.strintyN[c] = strinty;
    double un, vn, tbx, tby;
    stepuv_cd_at(k, d.uvelN[c], d.vvelN[c], d.aiN[c], d.cdnN[c], d.uocnN[c], d.vocnN[c], d.waterxN[c], d.wateryN[c], d.forcexN[c],
                 d.forceyN[c], d.nmassdti[c], d.fmN[c], strintx, strinty, d.uvelN_init[c], d.vvelN_init[c], d.TbN[c], un, vn, tbx, tby);
    ring_store(d, d.uvelN, i, j, un, 3, false);
    ring_store(d, d.vvelN, i, j, vn, 3, false);
    d.taubxN[c] = tbx;
    d.taubyN[c] = tby;
  }
}

__global__ void __launch_bounds__(256) kcd4_interp(const __grid_constant__ CDom d) {
  CELL_IJ(d.nx, d.ny);
  const int c = AT(i, j), e = c + 1, n = c + d.ld;
  double num, den;
  avg2_terms<false>(d.uvelE, d.earea, c, n, num, den);                // E2UA 'N'
  const double uU = ((den != 0.0) ? num / den : 0.0) * d.uvm[c];
  avg2_terms<false>(d.vvelN, d.narea, c, e, num, den);                // N2UA 'E'
  const double vU = ((den != 0.0) ? num / den : 0.0) * d.uvm[c];
  ring_store(d, d.uvel, i, j, uU, 3, true);
  ring_store(d, d.vvel, i, j, vU, 3, true);
}

#ifndef EVP_HOST_EMU  // launchers (and the cooperative kernel): not part of the host emulation (tests/emu_cgrid.cpp)
cudaError_t launch_cdgrid_subcycle(const CDom &d, const KParams &p, cudaStream_t s, int *launches) {
  dim3 b(32, 8), gU((d.nx + 31) / 32, (d.ny + 7) / 8), gT((d.nx + 1 + 31) / 32, (d.ny + 1 + 7) / 8);
  kcd1_stress_T<<<gT, b, 0, s>>>(d, p);
  kcd2_stress_U<<<gU, b, 0, s>>>(d, p);
  kcd3_momentum<<<gU, b, 0, s>>>(d, p);
  kcd4_interp<<<gU, b, 0, s>>>(d);
  *launches += 4;
  return cudaGetLastError();
}
#endif  // EVP_HOST_EMU

#ifndef EVP_HOST_EMU  // launchers: not part of the host emulation (tests/emu_cgrid.cpp)
cudaError_t launch_cgrid_subcycle(const CDom &d, const KParams &p, cudaStream_t s, int *launches) {
  dim3 b(32, 8), gU((d.nx + 31) / 32, (d.ny + 7) / 8), gT((d.nx + 1 + 31) / 32, (d.ny + 1 + 7) / 8);
  k1_strain_U<<<gU, b, 0, s>>>(d, p);
  k2_stress_T<<<gT, b, 0, s>>>(d, p);
  k3_stress_U<<<gU, b, 0, s>>>(d, p);
  k4_momentum<<<gU, b, 0, s>>>(d, p);
  k5_interp<<<gU, b, 0, s>>>(d);
  *launches += 5;
  return cudaGetLastError();
}
#endif  // EVP_HOST_EMU

}  // namespace EVP_NS
}  // namespace evp
