// emu_lane2.cpp -- the two-lanes-per-cell kernel of cice_b200/csrc/evp_lane2.cuh run THREAD BY THREAD ON THE HOST.
//
// Test infrastructure only (tests/test_emu_lane2.py); nothing in the product links this.  The kernel text is compiled
// unchanged with g++ -ffp-contract=off: every CUDA thread of a CTA is a host thread, __syncthreads is a pthread
// barrier, the named 64-thread barrier a second set of pthread barriers, __shfl_xor_sync a two-slot mailbox between the two
// lanes, __shared__ arrays are statics (CTAs run one after the other), the PDL calls are no-ops.  What this checks
// without a GPU: the kernel's index decoding, row/corner selection, ownership rule, stress swap, shared-memory hand-over
// and on-rank wrap stores, over several subcycles of the ping-pong, bit for bit against the oracle.  What it cannot
// check: the MUFU-seeded division / square root (IL = true is device-only; here IL = false) and anything about timing.
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include <cuda_runtime.h>  // host mode: the CUDA qualifiers expand to nothing

#undef __shared__
#define __shared__ static
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
#define EVP_USE_PDL 0
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#ifndef __grid_constant__
#define __grid_constant__
#endif

namespace emu {
constexpr int MAXT = 1024;
struct Idx { int x, y, z; };
thread_local Idx tid, bid;
thread_local unsigned shfl_seq;
struct Mail { std::atomic<unsigned> seq; double val[2]; };
Mail mail[MAXT];
pthread_barrier_t cta_barrier, named_barrier[16];
inline void syncthreads() { pthread_barrier_wait(&cta_barrier); }
inline void bar64(int id) { pthread_barrier_wait(&named_barrier[id]); }
inline double shfl_xor(double v, int lanemask) {
  const int me = tid.x, other = tid.x ^ lanemask;
  const unsigned n = ++shfl_seq;
  mail[me].val[n & 1] = v;
  mail[me].seq.store(n, std::memory_order_release);
  while (mail[other].seq.load(std::memory_order_acquire) < n) sched_yield();
  return mail[other].val[n & 1];
}
}  // namespace emu

#define threadIdx (emu::tid)
#define blockIdx (emu::bid)
#define __syncthreads() emu::syncthreads()
#define __shfl_xor_sync(mask, v, lanemask) emu::shfl_xor((v), (lanemask))
#define EVP_LANE2_BAR64(id) emu::bar64(id)

// device intrinsics evp_math.cuh mentions in its device-only fast paths (dead code here)
static inline int __double2hiint(double x) { int64_t b; memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline int __double2loint(double x) { int64_t b; memcpy(&b, &x, 8); return (int)(b & 0xffffffff); }
static inline double __hiloint2double(int hi, int lo) { int64_t b = ((int64_t)hi << 32) | (uint32_t)lo; double x; memcpy(&x, &b, 8); return x; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline double __shfl_sync(unsigned, double v, int, int) { return v; }  // stress_lane is not instantiated

#include "evp_math.cuh"
#include "evp_dom.cuh"

namespace evp {
namespace exact {
#include "evp_lane2.cuh"
}  // namespace exact
}  // namespace evp

using namespace evp;

template <int PX, int PY, int MAP>
static void run_grid(const Dom &d, const KParams &k, int cur, int flags) {
  const int nthreads = 2 * PX * PY;
  const int gx = (d.nx + PX - 2) / (PX - 1), gy = (d.ny + PY - 2) / (PY - 1);
  pthread_barrier_init(&emu::cta_barrier, nullptr, nthreads);
  for (auto &b : emu::named_barrier) pthread_barrier_init(&b, nullptr, 64);
  for (int t = 0; t < nthreads; ++t) emu::mail[t].seq.store(0);
  std::vector<std::thread> th;
  th.reserve(nthreads);
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([&, t] {
      emu::tid = {t, 0, 0};
      emu::shfl_seq = 0;
      for (int by = 0; by < gy; ++by)
        for (int bx = 0; bx < gx; ++bx) {
          emu::bid = {bx, by, 0};
          evp::exact::fused2_kernel<PX, PY, 1, false, MAP>(d, k, cur, flags);
          emu::syncthreads();  // the next CTA reuses the static "shared" arrays
        }
    });
  for (auto &x : th) x.join();
  pthread_barrier_destroy(&emu::cta_barrier);
  for (auto &b : emu::named_barrier) pthread_barrier_destroy(&b);
}

// One block in the reference's layout (nghost = 1) IS a dom: ld = nx_block, interior 1..nx_block-2.  `shape` selects the
// instantiation.  Runs ndte subcycles on two ping-pong copies and returns the final state in place.
extern "C" int emu_lane2_run(int shape, int nxb, int nyb, int wrap_ew, int wrap_ns, const KParams *kp, int ndte, const int32_t *maskT,
                             const int32_t *maskU, double *sig /*[12][n]*/, double *u, double *v, const double *geo /*[10][n]*/,
                             const double *strength, const double *in /*[11][n]*/, double *diag /*[4][n]*/) {
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
  for (int ks = 0; ks < ndte; ++ks) {
    const int cur = ks & 1, flags = (ks == ndte - 1) ? 1 : 0;
    switch (shape) {
      case 0: run_grid<32, 8, 0>(d, *kp, cur, flags); break;
      case 1: run_grid<16, 8, 0>(d, *kp, cur, flags); break;
      case 2: run_grid<32, 4, 0>(d, *kp, cur, flags); break;
      case 3: run_grid<16, 16, 0>(d, *kp, cur, flags); break;
      case 4: run_grid<32, 8, 1>(d, *kp, cur, flags); break;
      case 5: run_grid<32, 4, 1>(d, *kp, cur, flags); break;
      default: return 1;
    }
  }
  if (ndte & 1) {  // the result sits in copy 1
    memcpy(sig, sig1.data(), 12 * n * sizeof(double));
    memcpy(u, u1.data(), n * sizeof(double));
    memcpy(v, v1.data(), n * sizeof(double));
  }
  return 0;
}
