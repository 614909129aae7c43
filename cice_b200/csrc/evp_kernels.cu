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

#ifndef EVP_USE_PDL
#define EVP_USE_PDL 1
#endif

#ifndef EVP_NS
#error "compile with -DEVP_NS=exact or -DEVP_NS=fast"
#endif

namespace evp {
namespace EVP_NS {

__device__ __forceinline__ size_t at(const Dom &d, int i, int j) { return (size_t)j * d.ld + i; }

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

__device__ __forceinline__ void load_sigma(const Dom &d, int b, size_t c, Sigma &s) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    s.p[q] = d.sig[b][q][c];
    s.m[q] = d.sig[b][4 + q][c];
    s.s12[q] = d.sig[b][8 + q][c];
  }
}
__device__ __forceinline__ void store_sigma(const Dom &d, int b, size_t c, const Sigma &s) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    d.sig[b][q][c] = s.p[q];
    d.sig[b][4 + q][c] = s.m[q];
    d.sig[b][8 + q][c] = s.s12[q];
  }
}

__device__ __forceinline__ void stress_at(const Dom &d, const KParams &k, int cur, int i, int j, Sigma &sg,
                                          double (&str)[8]) {
  const size_t c = at(d, i, j), w = c - 1, s = c - d.ld, sw = s - 1;
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
  const size_t c = at(d, i, j);
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
  const size_t c = at(d, i, j);
  if (!d.maskU[c]) return;
  const size_t e = c + 1, n = c + d.ld, ne = n + 1;
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

// ---------------------------------------------------------------------------------------------
// KERNEL_FUSED
// ---------------------------------------------------------------------------------------------
// FBX x FBY threads relax an FBX x FBY patch of T cells and advance the (FBX-1) x (FBY-1) U points it closes.
// MINB = CTAs per SM the register allocation is bounded for.
template <int FBX, int FBY, int MINB>
__global__ void __launch_bounds__(FBX *FBY, MINB) fused_kernel(const __grid_constant__ Dom d,
                                                                const __grid_constant__ KParams k, int cur) {
  __shared__ double sstr[8][FBY][FBX];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i = 1 + blockIdx.x * (FBX - 1) + tx;  // T cell of this thread
  const int j = 1 + blockIdx.y * (FBY - 1) + ty;
  const int nxt = cur ^ 1;
  const bool inT = (i <= d.nx + 1) && (j <= d.ny + 1);
  const size_t c = at(d, inT ? i : 1, inT ? j : 1);
#if EVP_USE_PDL
  // programmatic dependent launch: everything above overlaps the previous subcycle's tail
  cudaGridDependencySynchronize();
#endif

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

  if (tx < FBX - 1 && ty < FBY - 1 && i <= d.nx && j <= d.ny && d.maskU[c]) {
    const UOut o = stepu_point(d.u[cur][c], d.v[cur][c], d.cdn[c], d.aiu[c], d.uocn[c], d.vocn[c], d.waterx[c],
                               d.watery[c], d.forcex[c], d.forcey[c], d.umassdti[c], d.fm[c], d.uarear[c],
                               d.TbU[c], d.uinit[c], d.vinit[c], sstr[0][ty][tx], sstr[1][ty][tx + 1],
                               sstr[2][ty + 1][tx], sstr[3][ty + 1][tx + 1], sstr[4][ty][tx], sstr[5][ty + 1][tx],
                               sstr[6][ty][tx + 1], sstr[7][ty + 1][tx + 1], k);
    store_uv(d, d.u[nxt], d.v[nxt], i, j, o.u, o.v);
    d.strintx[c] = o.strintx;
    d.strinty[c] = o.strinty;
    d.taubx[c] = o.taubx;
    d.tauby[c] = o.tauby;
  }
}

template <int FBX, int FBY, int MINB>
static cudaError_t launch_fused_t(const Dom &d, const KParams &p, int cur, cudaStream_t s, bool pdl) {
  dim3 b(FBX, FBY), g((d.nx + FBX - 2) / (FBX - 1), (d.ny + FBY - 2) / (FBY - 1));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = g; cfg.blockDim = b; cfg.dynamicSmemBytes = 0; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, fused_kernel<FBX, FBY, MINB>, d, p, cur);
}

cudaError_t launch_fused(const Dom &d, const KParams &p, int cur, cudaStream_t s, int variant, bool pdl) {
  switch (variant) {
    case 1: return launch_fused_t<32, 8, 3>(d, p, cur, s, pdl);
    case 2: return launch_fused_t<32, 8, 4>(d, p, cur, s, pdl);
    case 3: return launch_fused_t<32, 4, 4>(d, p, cur, s, pdl);
    case 4: return launch_fused_t<32, 16, 1>(d, p, cur, s, pdl);
    case 5: return launch_fused_t<64, 4, 2>(d, p, cur, s, pdl);
    case 6: return launch_fused_t<32, 6, 3>(d, p, cur, s, pdl);
    default: return launch_fused_t<32, 8, 2>(d, p, cur, s, pdl);
  }
}

}  // namespace EVP_NS
}  // namespace evp
