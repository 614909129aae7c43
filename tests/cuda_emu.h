// cuda_emu.h -- just enough of the CUDA execution model on host threads to run the repo's kernels THREAD BY THREAD on a CPU.
//
// Test infrastructure only (tests/emu_bgrid.cpp); nothing in the product includes this.  A kernel is compiled unchanged
// by g++ (-ffp-contract=off for the `exact` arithmetic): every CUDA thread of a CTA is a host thread, CTAs of a grid run one
// after the other (their __shared__ arrays are function statics), __syncthreads is a pthread barrier, named barriers are
// arrive/wait counters, warp shuffles are mailboxes, device atomics are GCC atomics, the PDL calls are compiled out.
// What this checks without a GPU: index decoding, ownership rules, shared-memory hand-over, lane exchanges, wrap stores,
// ping-pong parity -- the logic of the kernel text.  What it cannot check: timing, memory-model races between CTAs
// (CTAs are sequential here), and the hardware's MUFU seeds (evp_math.cuh substitutes host seeds under EVP_HOST_EMU).
#pragma once
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include <cuda_runtime.h>  // host mode: the CUDA function/memory-space qualifiers expand to nothing

#undef __shared__
#define __shared__ static
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#ifndef __grid_constant__
#define __grid_constant__
#endif

namespace emu {
constexpr int MAXT = 1024, RING = 64;
struct Idx { int x, y, z; };
thread_local Idx tid, bid, bdim, gdim;
thread_local int lin;            // linear thread id inside the CTA
thread_local unsigned shfl_seq;  // shuffles executed by this thread (every lane of a warp executes the same sequence)
struct Mail { std::atomic<unsigned> seq; double val[RING]; };
static Mail mail[MAXT];
static pthread_barrier_t cta_barrier;
struct NamedBar { std::atomic<int> count; std::atomic<unsigned> gen; };
static NamedBar named[16];
static int cta_order = 0;  // 0 forwards, 1 backwards, 2 alternating from both ends

// CTAs that run CONCURRENTLY (launch_concurrent, for kernels whose CTAs wait for each other): each CTA has its own barriers and
// its own dynamic shared memory; the sequential launch() above keeps using the process-wide ones
struct Cta {
  pthread_barrier_t bar;
  NamedBar named[16];
  NamedBar warp[MAXT / 32];
  std::vector<unsigned char> smem;
};
thread_local Cta *cta = nullptr;

inline void yield() { sched_yield(); }
inline void syncthreads() { pthread_barrier_wait(cta ? &cta->bar : &cta_barrier); }
inline void bar_arrive(int id, int n) {
  NamedBar &b = cta ? cta->named[id] : named[id];
  if (b.count.fetch_add(1) + 1 == n) { b.count.store(0); b.gen.fetch_add(1); }
}
inline void bar_wait_on(NamedBar &b, int n) {
  const unsigned g = b.gen.load();
  if (b.count.fetch_add(1) + 1 == n) { b.count.store(0); b.gen.fetch_add(1); }
  else while (b.gen.load() == g) yield();
}
inline void bar_sync(int id, int n) { bar_wait_on(cta ? cta->named[id] : named[id], n); }
inline void syncwarp() { if (cta) bar_wait_on(cta->warp[lin / 32], 32); }   // (sequential mode: kernels there use shuffles only)
inline unsigned char *dyn_smem() { return cta ? cta->smem.data() : nullptr; }
// publish v, then read the value lane `src` (linear id inside the CTA) published in the same shuffle.  A lane may run up to
// RING shuffles ahead of a reader before it overwrites a slot; CTA barriers bound the lead (kernels here: <= 12).
inline double shfl_from(double v, int src) {
  const unsigned n = ++shfl_seq;
  mail[lin].val[n % RING] = v;
  mail[lin].seq.store(n, std::memory_order_release);
  while (mail[src].seq.load(std::memory_order_acquire) < n) yield();
  return mail[src].val[n % RING];
}

// run `kernel()` for every thread of every CTA of the grid
template <class F>
inline void launch(Idx grid, Idx block, F kernel) {
  const int nthreads = block.x * block.y;
  pthread_barrier_init(&cta_barrier, nullptr, nthreads);
  for (auto &b : named) { b.count.store(0); b.gen.store(0); }
  for (int t = 0; t < nthreads; ++t) mail[t].seq.store(0);
  std::vector<std::thread> th;
  th.reserve(nthreads);
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([=] {
      lin = t;
      tid = {t % block.x, t / block.x, 0};
      bdim = block;
      gdim = grid;
      shfl_seq = 0;
      const int nctas = grid.x * grid.y;
      for (int l = 0; l < nctas; ++l) {
        // CTAs run one after the other; the ORDER is a test parameter: a kernel whose CTAs are independent within a launch (no CTA
        // reads what another one writes) gives the same bits forwards, backwards and interleaved
        int q = l;
        if (cta_order == 1) q = nctas - 1 - l;
        else if (cta_order == 2) q = (l % 2 == 0) ? l / 2 : nctas - 1 - l / 2;
        bid = {q % grid.x, q / grid.x, 0};
        kernel();
        syncthreads();  // the next CTA reuses the static "shared" arrays
      }
    });
  for (auto &x : th) x.join();
  pthread_barrier_destroy(&cta_barrier);
}

// every CTA of the grid at once: grid.x * block.x host threads (1-D), `smem_bytes` of dynamic shared memory per CTA.  For kernels
// with waits between CTAs (evp_persist.cu); keep grid and block small.
template <class F>
inline void launch_concurrent(int nctas, int nthreads, size_t smem_bytes, F kernel) {
  std::vector<Cta> ctas(nctas);
  for (auto &c : ctas) {
    pthread_barrier_init(&c.bar, nullptr, nthreads);
    for (auto &b : c.named) { b.count.store(0); b.gen.store(0); }
    for (auto &b : c.warp) { b.count.store(0); b.gen.store(0); }
    c.smem.assign(smem_bytes + 16, 0xff);   // poisoned: a read of something never written shows up as a NaN
  }
  std::vector<std::thread> th;
  th.reserve((size_t)nctas * nthreads);
  for (int q = 0; q < nctas; ++q)
    for (int t = 0; t < nthreads; ++t)
      th.emplace_back([&, q, t] {
        cta = &ctas[q];
        lin = t;
        tid = {t, 0, 0};
        bdim = {nthreads, 1, 1};
        gdim = {nctas, 1, 1};
        bid = {q, 0, 0};
        kernel();
      });
  for (auto &x : th) x.join();
  for (auto &c : ctas) pthread_barrier_destroy(&c.bar);
}
}  // namespace emu

#define threadIdx (emu::tid)
#define blockIdx (emu::bid)
#define blockDim (emu::bdim)
#define gridDim (emu::gdim)
#define __syncthreads() emu::syncthreads()
#define __shfl_xor_sync(mask, v, lanemask) emu::shfl_from((v), emu::lin ^ (lanemask))
#define __shfl_sync(mask, v, src, width) emu::shfl_from((v), (emu::lin & ~((width) - 1)) + (src))

// device intrinsics and atomics the kernels mention
static inline int __double2hiint(double x) { int64_t b; memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline int __double2loint(double x) { int64_t b; memcpy(&b, &x, 8); return (int)(b & 0xffffffff); }
static inline double __hiloint2double(int hi, int lo) { int64_t b = ((int64_t)hi << 32) | (uint32_t)lo; double x; memcpy(&x, &b, 8); return x; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline long long __double_as_longlong(double x) { long long b; memcpy(&b, &x, 8); return b; }
static inline double __longlong_as_double(long long b) { double x; memcpy(&x, &b, 8); return x; }
static inline long long clock64() { return 0; }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
template <class T> static inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicExch(T *p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <class T> static inline T atomicMin(T *p, T v) { T o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (v < o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
template <class T> static inline T atomicMax(T *p, T v) { T o = __atomic_load_n(p, __ATOMIC_SEQ_CST); while (v > o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {} return o; }
template <class T> static inline T __ldcg(const T *p) { return *(const volatile T *)p; }
template <class T> static inline void __stcg(T *p, T v) { *(volatile T *)p = v; }

extern "C" void emu_set_cta_order(int order) { emu::cta_order = order; }
