// host_math.cpp -- the DEVICE arithmetic (cice_b200/csrc/evp_math.cuh: stress_point, stepu_point) compiled for the HOST with g++
// -ffp-contract=off, so that the source the CUDA kernels are built from can be checked against the oracle on a machine without
// a GPU (tests/test_host_math.py).  Test infrastructure only; nothing in the product links this.
#include <math.h>
#include <stdint.h>

#include <string.h>

#include <cuda_runtime.h>   // host mode: __device__ / __forceinline__ expand to nothing / always_inline

// host stand-ins for the device intrinsics the header mentions (only the plain IEEE paths are instantiated here: the
// MUFU-seeded fast paths div_fast / sqrt_fast are device-only and stay dead code on the host)
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
static inline int __double2hiint(double x) { int64_t b; memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline int __double2loint(double x) { int64_t b; memcpy(&b, &x, 8); return (int)(b & 0xffffffff); }
static inline double __hiloint2double(int hi, int lo) { int64_t b = ((int64_t)hi << 32) | (uint32_t)lo; double x; memcpy(&x, &b, 8); return x; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }

#include "evp_math.cuh"

using namespace evp;

extern "C" int host_math_one_subcycle(int nxb, int nyb, int ilo, int ihi, int jlo, int jhi, const KParams *kp, const int32_t *maskT,
                                      const int32_t *maskU, double *sig /*[12][n]*/, double *u, double *v, const double *geo /*[10][n]*/,
                                      const double *strength, const double *in /*[11][n]: cdn aiu uocn vocn waterx watery forcex forcey
                                      umassdti fm TbU*/, double *diag /*[4][n]*/, double *str /*[8][n] scratch*/) {
  const KParams &k = *kp;
  const size_t n = (size_t)nxb * nyb;
  auto at = [&](int i, int j) { return (size_t)(j - 1) * nxb + (i - 1); };
  for (size_t q = 0; q < 8 * n; ++q) str[q] = 0.0;
  const double *dxT = geo, *dyT = geo + n, *dxhy = geo + 2 * n, *dyhx = geo + 3 * n, *cxp = geo + 4 * n, *cyp = geo + 5 * n,
               *cxm = geo + 6 * n, *cym = geo + 7 * n, *dmin = geo + 8 * n, *uarear = geo + 9 * n;
  for (int j = jlo; j <= jhi + 1; ++j)
    for (int i = ilo; i <= ihi + 1; ++i) {
      const size_t c = at(i, j), w = at(i - 1, j), s = at(i, j - 1), sw = at(i - 1, j - 1);
      if (!maskT[c]) continue;
      Sigma sg;
      for (int q = 0; q < 4; ++q) { sg.p[q] = sig[q * n + c]; sg.m[q] = sig[(4 + q) * n + c]; sg.s12[q] = sig[(8 + q) * n + c]; }
      double st[8];
      stress_point<false>(u[c], v[c], u[w], v[w], u[s], v[s], u[sw], v[sw], dxT[c], dyT[c], dxhy[c], dyhx[c], cxp[c], cyp[c], cxm[c],
                            cym[c], dmin[c], strength[c], k, sg, st);
      for (int q = 0; q < 4; ++q) { sig[q * n + c] = sg.p[q]; sig[(4 + q) * n + c] = sg.m[q]; sig[(8 + q) * n + c] = sg.s12[q]; }
      for (int q = 0; q < 8; ++q) str[q * n + c] = st[q];
    }
  const double *uinit = u, *vinit = v;  // one subcycle from the entry state: uvel_init == uvel where it is read (own point, before the store)
  for (int j = jlo; j <= jhi; ++j)
    for (int i = ilo; i <= ihi; ++i) {
      const size_t c = at(i, j), e = at(i + 1, j), nn = at(i, j + 1), ne = at(i + 1, j + 1);
      if (!maskU[c]) continue;
      const UOut o = stepu_point<false>(u[c], v[c], in[c], in[n + c], in[2 * n + c], in[3 * n + c], in[4 * n + c], in[5 * n + c], in[6 * n + c],
                                        in[7 * n + c], in[8 * n + c], in[9 * n + c], uarear[c], in[10 * n + c], uinit[c], vinit[c],
                                        str[c], str[n + e], str[2 * n + nn], str[3 * n + ne], str[4 * n + c], str[5 * n + nn],
                                        str[6 * n + e], str[7 * n + ne], k);
      u[c] = o.u;
      v[c] = o.v;
      diag[c] = o.strintx; diag[n + c] = o.strinty; diag[2 * n + c] = o.taubx; diag[3 * n + c] = o.tauby;
    }
  return 0;
}
