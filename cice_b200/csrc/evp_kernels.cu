// evp_kernels.cu -- CUDA kernels of the EVP subcycle for sm_100a.
//
// Compiled twice (see build.py): -DEVP_NS=exact -fmad=false  and  -DEVP_NS=fast.
//
// Kernels here work on the device sub-domain layout of evp_internal.h (struct Dom):
//   stress_kernel / stepu_kernel : the reference's two sweeps, one thread per cell, `str` through
//                                  global memory (the first correct path; KERNEL_SPLIT)
//   fused_kernel                 : one launch per subcycle; a CTA relaxes the stresses of a
//                                  BX x BY patch of T cells, hands the 8 `str` terms to the momentum
//                                  step through shared memory and advances the (BX-1) x (BY-1) U
//                                  points the patch closes.  Reads only the `cur` copies of the
//                                  carried state and writes only the other copy, so overlapping
//                                  patches never race (KERNEL_FUSED)
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

#ifndef EVP_HOST_EMU  // launchers: not part of the host emulation (tests/emu_bgrid.cpp)
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
// edge index of a boundary U point (i==1 | i==nx | j==1 | j==ny), the row index of the push CSR
__device__ __forceinline__ int edge_index(const Dom &d, int i, int j) {
  if (j == 1) return i - 1;
  if (j == d.ny) return d.nx + i - 1;
  if (i == 1) return 2 * d.nx + (j - 2);
  return 2 * d.nx + (d.ny - 2) + (j - 2);
}

// The speculative body of fused_kernel (SPEC = true, the default): same arithmetic, same ownership rules, but
//  * every operand of the T cell is requested before the ice masks are known (addresses are always inside the dom);
//  * static operands (geometry, strength, masks) and the momentum operands are requested BEFORE the programmatic
//    grid dependency is resolved, i.e. while the previous subcycle's kernel is still draining;
//  * the 12 momentum operands go global -> shared with cp.async and are picked up after the CTA barrier.
// the in-kernel-halo form's edge-first tile table in CONSTANT memory (SPEC bit 4, EVP_B200_P2P_CONST_TILES=1): the table
// lookup is the first thing a CTA does and every address depends on it; from global memory that is one more serialised L2
// round trip per CTA, from the constant cache it is a few cycles.  Round-2 candidate, not yet measured.
__constant__ int c_tile_order[P2P_CONST_TILES];
#ifndef EVP_HOST_EMU  // launchers: not part of the host emulation (tests/emu_bgrid.cpp)
cudaError_t set_p2p_tiles(const int *host_tiles, int n) {
  if (n > P2P_CONST_TILES) return cudaErrorInvalidValue;
  return cudaMemcpyToSymbol(c_tile_order, host_tiles, sizeof(int) * (size_t)n);
}
#endif  // EVP_HOST_EMU

// Derived geometry (SPEC bit 5, EVP_B200_FUSED_VARIANT=59/63 after evp_b200_set_metric; round-2 candidate, not yet measured).
// Seven of the ten static T-cell arrays are functions of the two metric arrays HTN, HTE and of dxT, dyT
// (ice_dyn_shared.F90:384-388, 401-441): on sub-domains that stream from HBM, reading HTN/HTE (their i-1 / j-1 neighbours come
// from the same cache lines) instead of dxhy, dyhx, cxp, cyp, cxm, cym, DminTarea moves 360 instead of 400 B per cell and
// subcycle -- the algorithmic figure of SURVEY 8d.  The expressions are the reference's, operation for operation; whether they
// reproduce the host's arrays bit for bit on every cell the loop can touch is CHECKED on the device when the metric arrays are
// handed over (metric_verify_kernel); if a single cell differs the library keeps reading the arrays.
__constant__ const double *c_HTN, *c_HTE;
__constant__ double c_deltamin;
__device__ __forceinline__ void derive_geometry(double hn, double hs, double he, double hw, double dxT, double dyT, double deltamin,
                                                double &dxhy, double &dyhx, double &cxp, double &cyp, double &cxm, double &cym,
                                                double &dmin) {
  dxhy = 0.5 * (he - hw);            // p5*(HTE(i,j) - HTE(i-1,j))
  dyhx = 0.5 * (hn - hs);            // p5*(HTN(i,j) - HTN(i,j-1))
  cyp = (1.5 * he - 0.5 * hw);       // c1p5*HTE(i,j) - p5*HTE(i-1,j)
  cxp = (1.5 * hn - 0.5 * hs);
  cym = -(1.5 * hw - 0.5 * he);
  cxm = -(1.5 * hs - 0.5 * hn);
  dmin = deltamin * (dxT * dyT);     // deltaminEVP*tarea, tarea = dxT*dyT (ice_grid.F90:681-715)
}
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

constexpr int NUOP = 12;
// SPEC bit 0: speculative T-cell operand loads; bit 1: momentum operands through cp.async
template <int FBX, int FBY, bool P2P, int SPEC>
__device__ __forceinline__ void fused_spec_body(const Dom &d, const KParams &k, int cur, const P2PParams &pp, int last, int i, int j,
                                                bool inT, int c, double (&sstr)[8][FBY][FBX]) {
  constexpr bool SPT = (SPEC & 1) != 0, CPU = (SPEC & 2) != 0, IL = (SPEC & 4) != 0;  // bit 2: interleaved div/sqrt
  constexpr bool PAIR = (SPEC & 8) != 0;                                                 // bit 3: pairwise named barriers
  constexpr bool DER = (SPEC & 32) != 0;                                                 // bit 5: derived geometry (see above)
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
  if (PAIR && FBX == 32) {
    // U row ty needs the str terms of T rows ty and ty+1 only, and a row is one warp: instead of a CTA-wide barrier, warp
    // ty+1 arrives on named barrier ty+1 once its terms are in shared memory and warp ty waits there (64 threads each,
    // every barrier used once per launch), so a warp is held up by one neighbour, not by the slowest of eight
    if (ty >= 1) bar_arrive64(ty);
    if (ty < FBY - 1) bar_sync64(ty + 1);
  } else {
    __syncthreads();
  }
  if (uspot && mU) {
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
    store_uv(d, d.u[nxt], d.v[nxt], i, j, o.u, o.v);
    if (last) {
      d.strintx[c] = o.strintx;
      d.strinty[c] = o.strinty;
      d.taubx[c] = o.taubx;
      d.tauby[c] = o.tauby;
    }
    if (P2P && (i == 1 || i == d.nx || j == 1 || j == d.ny)) {
      const int e = edge_index(d, i, j);
      for (int q = pp.push_start[e]; q < pp.push_start[e + 1]; ++q) {
        const int pr = pp.push_peer[q];
        const int dst = pp.push_dst[q];
        pp.peer_u[nxt][pr][dst] = o.u;
        pp.peer_v[nxt][pr][dst] = o.v;
      }
    }
  }
}

template <int FBX, int FBY, int MINB, bool HOIST, bool P2P = false, int SPEC = 0>
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
    const int tile = (SPEC & 16) ? c_tile_order[b] : pp.tile_order[b];
    tbx = tile & 0xffff;
    tby = tile >> 16;
    edge_tile = b < pp.n_edge_tiles;
  }
  const int i = 1 + tbx * (FBX - 1) + tx;  // T cell of this thread
  const int j = 1 + tby * (FBY - 1) + ty;
  const int nxt = cur ^ 1;
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
  if (SPEC) {
    fused_spec_body<FBX, FBY, P2P, SPEC>(d, k, cur, pp, last, i, j, inT, c, sstr);
  } else {
#if EVP_USE_PDL
  // programmatic dependent launch: everything above overlaps the previous subcycle's tail
  cudaGridDependencySynchronize();
#endif

  // the momentum step's operands are requested before the stress arithmetic so that their L2 latency
  // hides under it (the U point of this thread is its own T cell index)
  const bool doU = tx < FBX - 1 && ty < FBY - 1 && i <= d.nx && j <= d.ny && d.maskU[c];
  double uin[16];
  if (HOIST && doU) {
    load_uin(d, k, cur, c, uin);
  }

  double str[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (inT && d.maskT[c]) {
    Sigma sg;
    stress_at(d, k, cur, i, j, sg, str);
    // each T cell is stored by exactly one CTA: the one that holds it off its E/N overlap edge
    const bool own = (tx < FBX - 1 || i == d.nx + 1) && (ty < FBY - 1 || j == d.ny + 1);
    if (own) store_sigma(d, nxt, c, sg);
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) sstr[q][ty][tx] = str[q];
  __syncthreads();

  if (doU) {
    if (!HOIST) {
      load_uin(d, k, cur, c, uin);
    }
    const UOut o = stepu_point(uin[0], uin[1], uin[2], uin[3], uin[4], uin[5], uin[6], uin[7], uin[8], uin[9], uin[10], uin[11],
                               uin[12], uin[13], uin[14], uin[15], sstr[0][ty][tx], sstr[1][ty][tx + 1],
                               sstr[2][ty + 1][tx], sstr[3][ty + 1][tx + 1], sstr[4][ty][tx], sstr[5][ty + 1][tx],
                               sstr[6][ty][tx + 1], sstr[7][ty + 1][tx + 1], k);
    store_uv(d, d.u[nxt], d.v[nxt], i, j, o.u, o.v);
    if (last) {
      // the loop overwrites these every subcycle and nothing reads them in between (ice_dyn_shared.F90:948-965);
      // only the last subcycle's values survive, as in the reference's own 1-D solver (calc_diag_1d, ice_dyn_core1d.F90:607)
      d.strintx[c] = o.strintx;
      d.strinty[c] = o.strinty;
      d.taubx[c] = o.taubx;
      d.tauby[c] = o.tauby;
    }
    if (P2P && (i == 1 || i == d.nx || j == 1 || j == d.ny)) {
      // this point is a ghost cell of up to three neighbour GPUs: store it there over NVLink right away, so the
      // traffic is spread over the kernel and long acknowledged when the hand-over fence below is issued
      const int e = edge_index(d, i, j);
      for (int q = pp.push_start[e]; q < pp.push_start[e + 1]; ++q) {
        const int pr = pp.push_peer[q];
        const int dst = pp.push_dst[q];
        pp.peer_u[nxt][pr][dst] = o.u;
        pp.peer_v[nxt][pr][dst] = o.v;
      }
    }
  }
  }  // !SPEC
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

// ---------------------------------------------------------------------------------------------
// KERNEL_FUSED, strip form (the default without in-kernel NVLink halo).
// A CTA of 32 x 8 threads walks up a strip of 31 U columns in `m` chunks of 8 T rows.  The `str` terms of a chunk's
// top T row stay in shared memory (row 0) for the next chunk, so inside a strip no T row is relaxed twice: a CTA
// relaxes 8m T rows and advances 8m-1 U rows (fused_kernel: 8 and 7).  `m` is chosen on the host so that the whole
// grid is one wave of co-resident CTAs when the sub-domain is small (gx1: m = 2, 286 CTAs on 296 slots instead of
// 605 CTAs = 2.04 waves).  Ownership and ping-pong rules are those of fused_kernel.
// ---------------------------------------------------------------------------------------------
constexpr int SBX = 32, SBY = 8;
__global__ void __launch_bounds__(SBX *SBY, 2) strip_kernel(const __grid_constant__ Dom d, const __grid_constant__ KParams k,
                                                            int cur, int m, int last) {
  __shared__ double sstr[8][SBY + 1][SBX];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int nxt = cur ^ 1;
  const int i = 1 + blockIdx.x * (SBX - 1) + tx;
  const int j0 = 1 + blockIdx.y * (SBY * m - 1);
#if EVP_USE_PDL
  cudaGridDependencySynchronize();
#endif
  for (int ch = 0; ch < m; ++ch) {
    const int jT = j0 + SBY * ch + ty;
    if (jT - ty > d.ny + 1) break;  // uniform: the strip has left the sub-domain
    const bool inT = (i <= d.nx + 1) && (jT <= d.ny + 1);
    const int c = at(d, inT ? i : 1, inT ? jT : 1);
    double str[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (inT && d.maskT[c]) {
      Sigma sg;
      stress_at(d, k, cur, i, jT, sg, str);
      // the strip's last T row is relaxed again by the strip above (as its first row), which stores it
      const bool own = (tx < SBX - 1 || i == d.nx + 1) && (!(ch == m - 1 && ty == SBY - 1) || jT == d.ny + 1);
      if (own) store_sigma(d, nxt, c, sg);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) sstr[q][ty + 1][tx] = str[q];
    __syncthreads();
    // chunk 0 closes the U rows of its own first 7 T rows; later chunks close 8: the row carried in shared-memory
    // row 0 plus their own first 7
    const int jU = ch == 0 ? jT : jT - 1;
    const int r = ch == 0 ? ty + 1 : ty;
    const bool doU = (ch > 0 || ty < SBY - 1) && tx < SBX - 1 && i <= d.nx && jU <= d.ny;
    if (doU) {
      const int cu = at(d, i, jU);
      if (d.maskU[cu]) {
        double uin[16];
        load_uin(d, k, cur, cu, uin);
        const UOut o = stepu_point(uin[0], uin[1], uin[2], uin[3], uin[4], uin[5], uin[6], uin[7], uin[8], uin[9], uin[10], uin[11],
                                   uin[12], uin[13], uin[14], uin[15], sstr[0][r][tx], sstr[1][r][tx + 1], sstr[2][r + 1][tx],
                                   sstr[3][r + 1][tx + 1], sstr[4][r][tx], sstr[5][r + 1][tx], sstr[6][r][tx + 1],
                                   sstr[7][r + 1][tx + 1], k);
        store_uv(d, d.u[nxt], d.v[nxt], i, jU, o.u, o.v);
        if (last) {
          d.strintx[cu] = o.strintx;
          d.strinty[cu] = o.strinty;
          d.taubx[cu] = o.taubx;
          d.tauby[cu] = o.tauby;
        }
      }
    }
    if (ch + 1 < m) {
      __syncthreads();  // every read of this chunk's rows is done
      if (ty == SBY - 1) {
#pragma unroll
        for (int q = 0; q < 8; ++q) sstr[q][0][tx] = str[q];
      }
    }
  }
}

#ifndef EVP_HOST_EMU  // launchers: not part of the host emulation (tests/emu_bgrid.cpp)
cudaError_t launch_strip(const Dom &d, const KParams &p, int cur, int m, cudaStream_t s, bool pdl, int last) {
  dim3 b(SBX, SBY), g((d.nx + SBX - 2) / (SBX - 1), (d.ny + SBY * m - 2) / (SBY * m - 1));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = g; cfg.blockDim = b; cfg.dynamicSmemBytes = 0; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, strip_kernel, d, p, cur, m, last);
}

#endif  // EVP_HOST_EMU

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
// KERNEL_FUSED, corner-parallel: 1024 threads = a 32 x 8 patch of T cells x 4 corner lanes (stress_lane).
// Same patch geometry, ownership rule and ping-pong as fused_kernel; the momentum step runs on the first 256
// threads, one per U point.
// ---------------------------------------------------------------------------------------------
constexpr int F4X = 32, F4Y = 8;
__global__ void __launch_bounds__(F4X *F4Y * 4, 1) fused4_kernel(const __grid_constant__ Dom d, const __grid_constant__ KParams k, int cur) {
  __shared__ double sstr[8][F4Y][F4X];
  __shared__ double su[F4Y + 1][F4X + 1], sv[F4Y + 1][F4X + 1];
  const int t = threadIdx.x;
  const int nxt = cur ^ 1;
  const int i0 = 1 + blockIdx.x * (F4X - 1), j0 = 1 + blockIdx.y * (F4Y - 1);
  {
    const double *__restrict__ U = d.u[cur];
    const double *__restrict__ V = d.v[cur];
    for (int q = t; q < (F4X + 1) * (F4Y + 1); q += F4X * F4Y * 4) {
      const int a = q % (F4X + 1), b = q / (F4X + 1);
      const int gi = i0 - 1 + a, gj = j0 - 1 + b;
      const bool in = (gi <= d.nx + 1) && (gj <= d.ny + 1);
      su[b][a] = in ? U[at(d, gi, gj)] : 0.0;
      sv[b][a] = in ? V[at(d, gi, gj)] : 0.0;
    }
  }
  __syncthreads();
  {
    const int corner = t & 3, cell = t >> 2, cx = cell & (F4X - 1), cy = cell / F4X;
    const int i = i0 + cx, j = j0 + cy;
    const bool inT = (i <= d.nx + 1) && (j <= d.ny + 1);
    const int c = at(d, inT ? i : 1, inT ? j : 1);
    double str_u = 0.0, str_v = 0.0;
    if (inT && d.maskT[c]) {
      const int a = cx + 1, b = cy + 1;
      const int sa = a - ((corner == NW || corner == SW) ? 1 : 0), sb = b - ((corner >= SW) ? 1 : 0);
      const int xa = 2 * a - 1 - sa, yb = 2 * b - 1 - sb;
      double sp = d.sig[cur][corner][c], sm = d.sig[cur][4 + corner][c], s12 = d.sig[cur][8 + corner][c];
      const unsigned gmask = 0xFu << ((t & 31) & ~3);
      stress_lane(corner, su[sb][sa], sv[sb][sa], su[sb][xa], sv[sb][xa], su[yb][sa], sv[yb][sa], d.dxT[c], d.dyT[c], d.dxhy[c],
                  d.dyhx[c], d.cxp[c], d.cyp[c], d.cxm[c], d.cym[c], d.DminTarea[c], d.strength[c], k, gmask, sp, sm, s12, str_u, str_v);
      const bool own = (cx < F4X - 1 || i == d.nx + 1) && (cy < F4Y - 1 || j == d.ny + 1);
      if (own) {
        d.sig[nxt][corner][c] = sp;
        d.sig[nxt][4 + corner][c] = sm;
        d.sig[nxt][8 + corner][c] = s12;
      }
    }
    // NE: str1,str5  NW: str2,str7  SW: str4,str8  SE: str3,str6
    const int qu = (corner == NE) ? 0 : (corner == NW) ? 1 : (corner == SW) ? 3 : 2;
    const int qv = (corner == NE) ? 4 : (corner == NW) ? 6 : (corner == SW) ? 7 : 5;
    sstr[qu][cy][cx] = str_u;
    sstr[qv][cy][cx] = str_v;
  }
  __syncthreads();
  if (t < F4X * F4Y) {
    const int tx = t & (F4X - 1), ty = t / F4X;
    const int i = i0 + tx, j = j0 + ty;
    if (tx < F4X - 1 && ty < F4Y - 1 && i <= d.nx && j <= d.ny) {
      const int c = at(d, i, j);
      if (d.maskU[c]) {
        const UOut o = stepu_point(su[ty + 1][tx + 1], sv[ty + 1][tx + 1], d.cdn[c], d.aiu[c], d.uocn[c], d.vocn[c], d.waterx[c],
                                   d.watery[c], d.forcex[c], d.forcey[c], d.umassdti[c], d.fm[c], d.uarear[c], d.TbU[c],
                                   d.uinit[c], d.vinit[c], sstr[0][ty][tx], sstr[1][ty][tx + 1], sstr[2][ty + 1][tx],
                                   sstr[3][ty + 1][tx + 1], sstr[4][ty][tx], sstr[5][ty + 1][tx], sstr[6][ty][tx + 1],
                                   sstr[7][ty + 1][tx + 1], k);
        store_uv(d, d.u[nxt], d.v[nxt], i, j, o.u, o.v);
        d.strintx[c] = o.strintx;
        d.strinty[c] = o.strinty;
        d.taubx[c] = o.taubx;
        d.tauby[c] = o.tauby;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// KERNEL_QUEUE: the fused patches of ALL subcycles as one work queue in a single launch.
// Item w = (subcycle w / ntiles, patch w % ntiles).  A CTA claims the next item with an atomic counter, waits
// until the patch itself and its (up to 8, cyclically wrapped) neighbour patches have published the previous
// subcycle (per-patch progress counters, release/acquire through L2), then does exactly what fused_kernel does.
// The same wait also covers the write-after-read hazard of the ping-pong copies.  Items are claimed in order,
// so the oldest unfinished item always has its dependencies met: no deadlock as long as all CTAs are resident
// (cooperative launch).  No kernel boundaries, no partial last wave, subcycles overlap at the patch level.
// ---------------------------------------------------------------------------------------------
constexpr int QBX = 32, QBY = 8;
__global__ void __launch_bounds__(QBX *QBY, 2) queue_kernel(const __grid_constant__ Dom d, const __grid_constant__ KParams k,
                                                            int ndte, int ntx, int nty, unsigned *__restrict__ progress,
                                                            unsigned *__restrict__ counter) {
  __shared__ double sstr[8][QBY][QBX];
  __shared__ unsigned s_item;
  const int tx = threadIdx.x, ty = threadIdx.y, t = ty * QBX + tx;
  const unsigned ntiles = (unsigned)(ntx * nty), nitems = ntiles * (unsigned)ndte;
  for (;;) {
    if (t == 0) s_item = atomicAdd(counter, 1u);
    __syncthreads();
    const unsigned item = s_item;
    if (item >= nitems) break;
    const int ksub = (int)(item / ntiles), tile = (int)(item % ntiles);
    const int tbx = tile % ntx, tby = tile / ntx;
    const int cur = ksub & 1, nxt = cur ^ 1;
    if (ksub > 0 && t < 9) {
      int nx_ = tbx + (t % 3) - 1, ny_ = tby + (t / 3) - 1;
      if (d.wrap_ew) nx_ = (nx_ + ntx) % ntx;
      if (d.wrap_ns) ny_ = (ny_ + nty) % nty;
      if (nx_ >= 0 && nx_ < ntx && ny_ >= 0 && ny_ < nty) {
        const unsigned *f = progress + ny_ * ntx + nx_;
        while (ld_acquire_gpu(f) < (unsigned)ksub) {}
      }
    }
    __syncthreads();

    const int i = 1 + tbx * (QBX - 1) + tx, j = 1 + tby * (QBY - 1) + ty;
    const bool inT = (i <= d.nx + 1) && (j <= d.ny + 1);
    const int c = at(d, inT ? i : 1, inT ? j : 1);
    double str[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double uc = 0.0, vc = 0.0;
    if (inT && d.maskT[c]) {
      const int w = c - 1, s = c - d.ld, sw = s - 1;
      // state written by other CTAs inside this launch: read through L2 (ld.cg), never a stale L1 line
      const double *__restrict__ U = d.u[cur];
      const double *__restrict__ V = d.v[cur];
      uc = __ldcg(U + c); vc = __ldcg(V + c);
      Sigma sg;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        sg.p[q] = __ldcg(d.sig[cur][q] + c);
        sg.m[q] = __ldcg(d.sig[cur][4 + q] + c);
        sg.s12[q] = __ldcg(d.sig[cur][8 + q] + c);
      }
      stress_point(uc, vc, __ldcg(U + w), __ldcg(V + w), __ldcg(U + s), __ldcg(V + s), __ldcg(U + sw), __ldcg(V + sw), d.dxT[c],
                   d.dyT[c], d.dxhy[c], d.dyhx[c], d.cxp[c], d.cyp[c], d.cxm[c], d.cym[c], d.DminTarea[c], d.strength[c], k, sg, str);
      const bool own = (tx < QBX - 1 || i == d.nx + 1) && (ty < QBY - 1 || j == d.ny + 1);
      if (own) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          __stcg(d.sig[nxt][q] + c, sg.p[q]);
          __stcg(d.sig[nxt][4 + q] + c, sg.m[q]);
          __stcg(d.sig[nxt][8 + q] + c, sg.s12[q]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) sstr[q][ty][tx] = str[q];
    __syncthreads();
    if (tx < QBX - 1 && ty < QBY - 1 && i <= d.nx && j <= d.ny && d.maskU[c]) {
      if (!d.maskT[c]) { uc = __ldcg(d.u[cur] + c); vc = __ldcg(d.v[cur] + c); }
      const UOut o = stepu_point(uc, vc, d.cdn[c], d.aiu[c], d.uocn[c], d.vocn[c], d.waterx[c], d.watery[c], d.forcex[c],
                                 d.forcey[c], d.umassdti[c], d.fm[c], d.uarear[c], d.TbU[c], d.uinit[c], d.vinit[c],
                                 sstr[0][ty][tx], sstr[1][ty][tx + 1], sstr[2][ty + 1][tx], sstr[3][ty + 1][tx + 1],
                                 sstr[4][ty][tx], sstr[5][ty + 1][tx], sstr[6][ty][tx + 1], sstr[7][ty + 1][tx + 1], k);
      store_uv(d, d.u[nxt], d.v[nxt], i, j, o.u, o.v);
      if (ksub == ndte - 1) {
        d.strintx[c] = o.strintx;
        d.strinty[c] = o.strinty;
        d.taubx[c] = o.taubx;
        d.tauby[c] = o.tauby;
      }
    }
    __syncthreads();  // every store of the patch is issued; also protects sstr and s_item for the next item
    if (t == 0) {
      __threadfence();
      st_release_gpu(progress + tile, (unsigned)(ksub + 1));
    }
  }
}

#ifndef EVP_HOST_EMU  // launchers: not part of the host emulation (tests/emu_bgrid.cpp)
cudaError_t launch_queue(const Dom &d, const KParams &p, int ndte, unsigned *progress, unsigned *counter, int nctas, cudaStream_t s) {
  int ntx = (d.nx + QBX - 2) / (QBX - 1), nty = (d.ny + QBY - 2) / (QBY - 1);
  void *args[] = {(void *)&d, (void *)&p, (void *)&ndte, (void *)&ntx, (void *)&nty, (void *)&progress, (void *)&counter};
  return cudaLaunchCooperativeKernel((const void *)queue_kernel, dim3(nctas), dim3(QBX, QBY), args, 0, s);
}

#endif  // EVP_HOST_EMU

#include "evp_lane2.cuh"

#ifndef EVP_HOST_EMU  // launchers: not part of the host emulation (tests/emu_bgrid.cpp)
template <int PX, int PY, int MINB, bool IL, int MAP, bool SPT = false>
static cudaError_t launch_fused2_t(const Dom &d, const KParams &p, int cur, cudaStream_t s, bool pdl, int last) {
  dim3 b(2 * PX * PY), g((d.nx + PX - 2) / (PX - 1), (d.ny + PY - 2) / (PY - 1));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = g; cfg.blockDim = b; cfg.dynamicSmemBytes = 0; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, fused2_kernel<PX, PY, MINB, IL, MAP, SPT>, d, p, cur, last);
}

template <int FBX, int FBY, int MINB, bool HOIST = false, int SPEC = 0>
static cudaError_t launch_fused_t(const Dom &d, const KParams &p, int cur, cudaStream_t s, bool pdl, int last) {
  dim3 b(FBX, FBY), g((d.nx + FBX - 2) / (FBX - 1), (d.ny + FBY - 2) / (FBY - 1));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = g; cfg.blockDim = b; cfg.dynamicSmemBytes = 0; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  static const P2PParams nop2p{};
  return cudaLaunchKernelEx(&cfg, fused_kernel<FBX, FBY, MINB, HOIST, false, SPEC>, d, p, cur, nop2p, 0, last);  // last: bit 0 = last subcycle, bit 1 = early PDL trigger
}

cudaError_t launch_fused(const Dom &d, const KParams &p, int cur, cudaStream_t s, int variant, bool pdl, int last) {
  if (variant == 20) {
    dim3 g4((d.nx + F4X - 2) / (F4X - 1), (d.ny + F4Y - 2) / (F4Y - 1));
    fused4_kernel<<<g4, F4X * F4Y * 4, 0, s>>>(d, p, cur);
    return cudaGetLastError();
  }
  switch (variant) {
    case 1: return launch_fused_t<32, 8, 3>(d, p, cur, s, pdl, last);
    case 2: return launch_fused_t<32, 8, 4>(d, p, cur, s, pdl, last);
    case 3: return launch_fused_t<32, 4, 4>(d, p, cur, s, pdl, last);
    case 4: return launch_fused_t<32, 16, 1>(d, p, cur, s, pdl, last);
    case 5: return launch_fused_t<64, 4, 2>(d, p, cur, s, pdl, last);
    case 6: return launch_fused_t<32, 6, 3>(d, p, cur, s, pdl, last);
    case 7: return launch_fused_t<32, 8, 2, true>(d, p, cur, s, pdl, last);
    case 8: return launch_fused_t<33, 8, 2, false>(d, p, cur, s, pdl, last);
    case 9: return launch_fused_t<33, 8, 2, true>(d, p, cur, s, pdl, last);
    case 10: return launch_fused_t<32, 4, 4, true>(d, p, cur, s, pdl, last);
    case 11: return launch_fused_t<32, 9, 2>(d, p, cur, s, pdl, last);
    case 12: return launch_fused_t<32, 10, 2>(d, p, cur, s, pdl, last);
    case 13: return launch_fused_t<32, 7, 3>(d, p, cur, s, pdl, last);
    case 14: return launch_fused_t<32, 5, 4>(d, p, cur, s, pdl, last);
    case 15: return launch_fused_t<32, 11, 1>(d, p, cur, s, pdl, last);
    case 16: return launch_fused_t<32, 8, 2>(d, p, cur, s, pdl, last);  // the non-speculative form (round-1 default)
    case 17: return launch_fused_t<32, 8, 2, false, 1>(d, p, cur, s, pdl, last);  // speculative T loads only
    case 18: return launch_fused_t<32, 8, 2, false, 2>(d, p, cur, s, pdl, last);  // cp.async momentum operands only
    case 19: return launch_fused_t<32, 8, 2, false, 3>(d, p, cur, s, pdl, last);  // both
    case 21: return launch_fused_t<32, 8, 2, false, 5>(d, p, cur, s, pdl, last);  // 17 + interleaved div/sqrt
    case 22: return launch_fused_t<32, 8, 2, false, 7>(d, p, cur, s, pdl, last);  // 19 + interleaved div/sqrt
    case 23: return launch_fused_t<32, 8, 2, false, 4>(d, p, cur, s, pdl, last);  // 16 + interleaved div/sqrt
    case 31: return launch_fused_t<32, 8, 2, false, 12>(d, p, cur, s, pdl, last);  // 23 + pairwise named barriers
    case 32: return launch_fused_t<32, 8, 2, false, 11>(d, p, cur, s, pdl, last);  // 19 + pairwise named barriers
    case 24: return launch_fused_t<32, 8, 3, false, 4>(d, p, cur, s, pdl, last);  // 23 with other CTA shapes / residency
    case 25: return launch_fused_t<32, 6, 3, false, 4>(d, p, cur, s, pdl, last);
    case 26: return launch_fused_t<32, 5, 4, false, 4>(d, p, cur, s, pdl, last);
    case 27: return launch_fused_t<32, 4, 4, false, 4>(d, p, cur, s, pdl, last);
    case 28: return launch_fused_t<32, 16, 1, false, 4>(d, p, cur, s, pdl, last);
    case 29: return launch_fused_t<32, 12, 1, false, 4>(d, p, cur, s, pdl, last);
    case 59: return launch_fused_t<32, 8, 2, false, 3 | 32>(d, p, cur, s, pdl, last);  // 19 + derived geometry (after set_metric)
    case 63: return launch_fused_t<32, 8, 2, false, 4 | 32>(d, p, cur, s, pdl, last);  // 23 + derived geometry
    // two lanes per T cell (evp_lane2.cuh): <patch x, patch y, CTAs per SM, interleaved div/sqrt, lane mapping>
    case 40: return launch_fused2_t<32, 8, 2, true, 0>(d, p, cur, s, pdl, last);   // 512 threads, 64 registers
    case 41: return launch_fused2_t<32, 8, 1, true, 0>(d, p, cur, s, pdl, last);   // 512 threads, 128 registers
    case 42: return launch_fused2_t<16, 8, 3, true, 0>(d, p, cur, s, pdl, last);   // 256 threads, 80 registers
    case 43: return launch_fused2_t<16, 8, 4, true, 0>(d, p, cur, s, pdl, last);   // 256 threads, 64 registers
    case 44: return launch_fused2_t<32, 4, 3, true, 0>(d, p, cur, s, pdl, last);   // 256 threads, 80 registers
    case 45: return launch_fused2_t<16, 16, 2, true, 0>(d, p, cur, s, pdl, last);  // 512 threads, 64 registers
    case 46: return launch_fused2_t<32, 8, 2, false, 0>(d, p, cur, s, pdl, last);  // 40 with the built-in / and sqrt
    case 47: return launch_fused2_t<32, 8, 2, true, 0, true>(d, p, cur, s, pdl, last);   // 40 with speculative operand loads
    case 48: return launch_fused2_t<16, 8, 3, true, 0, true>(d, p, cur, s, pdl, last);   // 42 with speculative operand loads
    case 49: return launch_fused2_t<32, 8, 2, true, 1, true>(d, p, cur, s, pdl, last);   // 50 with speculative operand loads
    case 50: return launch_fused2_t<32, 8, 2, true, 1>(d, p, cur, s, pdl, last);   // warp-uniform roles, shared-memory swap
    case 54: return launch_fused2_t<32, 8, 2, true, 2>(d, p, cur, s, pdl, last);   // 50 with one specialised code path per role
    case 55: return launch_fused2_t<32, 4, 4, true, 2>(d, p, cur, s, pdl, last);
    case 56: return launch_fused2_t<32, 8, 2, true, 2, true>(d, p, cur, s, pdl, last);   // 54 with speculative operand loads
    case 51: return launch_fused2_t<32, 8, 1, true, 1>(d, p, cur, s, pdl, last);
    case 52: return launch_fused2_t<32, 4, 3, true, 1>(d, p, cur, s, pdl, last);
    case 53: return launch_fused2_t<32, 4, 4, true, 1>(d, p, cur, s, pdl, last);
    default: return launch_fused_t<32, 8, 2, false, 4>(d, p, cur, s, pdl, last);
  }
}

// ksub = -1: loop start hand-shake; ksub = -2 - ndte ... no: see launch_p2p_aux
cudaError_t launch_fused_p2p(const Dom &d, const KParams &p, const P2PParams &pp, int cur, int ksub, int last, int variant, cudaStream_t s) {
  if (ksub == -1) {
    p2p_start_kernel<<<1, 32, 0, s>>>(pp);
    return cudaGetLastError();
  }
  if (ksub <= -2) {  // closing wait after -2-ksub subcycles
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
  if ((variant & 0xff) == 40 || (variant & 0xff) == 50 || (variant & 0xff) == 54) {  // two lanes per T cell (evp_lane2.cuh), same 32 x 8 tile table
    cfg.blockDim = dim3(512);
    const int ctiles = (variant & 0x100) ? 1 : 0;
    if ((variant & 0xff) == 40) return cudaLaunchKernelEx(&cfg, fused2_p2p_kernel<32, 8, 2, true, 0>, d, p, cur, pp, ksub, flags, ctiles);
    if ((variant & 0xff) == 54) return cudaLaunchKernelEx(&cfg, fused2_p2p_kernel<32, 8, 2, true, 2>, d, p, cur, pp, ksub, flags, ctiles);
    return cudaLaunchKernelEx(&cfg, fused2_p2p_kernel<32, 8, 2, true, 1>, d, p, cur, pp, ksub, flags, ctiles);
  }
  if ((variant & 0xff) == 59) return cudaLaunchKernelEx(&cfg, fused_kernel<32, 8, 2, false, true, 3 | 32>, d, p, cur, pp, ksub, flags);
  if ((variant & 0xff) == 63) return cudaLaunchKernelEx(&cfg, fused_kernel<32, 8, 2, false, true, 4 | 32>, d, p, cur, pp, ksub, flags);
  if (variant & 0x100) {  // tile table in constant memory (set_p2p_tiles)
    if ((variant & 0xff) == 19) return cudaLaunchKernelEx(&cfg, fused_kernel<32, 8, 2, false, true, 3 | 16>, d, p, cur, pp, ksub, flags);
    return cudaLaunchKernelEx(&cfg, fused_kernel<32, 8, 2, false, true, 4 | 16>, d, p, cur, pp, ksub, flags);
  }
  if (variant == 19) return cudaLaunchKernelEx(&cfg, fused_kernel<32, 8, 2, false, true, 3>, d, p, cur, pp, ksub, flags);
  return cudaLaunchKernelEx(&cfg, fused_kernel<32, 8, 2, false, true, 4>, d, p, cur, pp, ksub, flags);
  return cudaGetLastError();
}
#endif  // EVP_HOST_EMU

}  // namespace EVP_NS
}  // namespace evp
