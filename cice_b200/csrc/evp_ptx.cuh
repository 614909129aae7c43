// evp_ptx.cuh -- the inline-PTX helpers of the B-grid kernels (evp_kernels.cu), in one place.
//
// With EVP_HOST_EMU defined (tests/emu_bgrid.cpp: the kernels run thread by thread on the host as a CPU-side check of their
// index logic) every helper has a plain C++ stand-in with the same meaning minus the timing: a volatile load is a load, an
// asynchronous copy is a copy, fences and acquire/release accesses are C++ atomics' business.  The product never defines it.
#pragma once
#include "evp_internal.h"

namespace evp {

#ifndef EVP_HOST_EMU

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// wait until flag >= want; bounded so that a lost peer cannot hang the GPU (sets *err instead).  The bound is wall-clock
// (%globaltimer, ns) and generous -- 60 s unless EVP_B200_P2P_TIMEOUT_S says otherwise (evp_abi.cu -> set_wait_timeout) -- because
// rank skew of seconds is ordinary: first-call graph instantiation or module load on one rank, MPS time slicing, a debugger.
static __device__ unsigned long long g_wait_timeout_ns = 60000000000ULL;
__device__ __forceinline__ void wait_flag(const unsigned long long *flag, unsigned long long want, int *err) {
  if (ld_acquire_sys(flag) >= want) return;
  const unsigned long long t0 = gtime();
  while (ld_acquire_sys(flag) < want) {
    if (gtime() - t0 > g_wait_timeout_ns) { atomicExch(err, 1); break; }
  }
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// loads the compiler may not sink into the branch that consumes them (asm volatile): the speculative form of the
// fused kernel issues every operand load of a cell at once, without waiting for the ice mask, so that a CTA pays
// one L2 round trip instead of four (mask U -> mask T -> stress operands -> momentum operands)
__device__ __forceinline__ double ld_f64(const double *p) {
  double v;
  asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_nc_f64_pinned(const double *p) {  // ... and not to be moved across memory operations either
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_nc_f64(const double *p) {  // never written while the loop runs
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
// data written by other SMs (or other GPUs) since this SM last saw the line: always served by L2
__device__ __forceinline__ double ld_cg_f64(const double *p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_nc_u8(const unsigned char *p) {
  unsigned v;
  asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
// 8-byte asynchronous global -> shared copy (LDGSTS): the momentum operands travel while the stresses are relaxed
__device__ __forceinline__ void cp_async8(double *smem, const double *g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// named barriers 1..15 over 64 threads (two warps): arrive without waiting / wait for both
__device__ __forceinline__ void bar_arrive64(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void bar_sync64(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned *p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// KERNEL_PERSISTENT (evp_persist.cu): a named barrier over n threads, the warp barrier, the dynamic shared memory of the CTA, and the
// bounded wait on a neighbour tile's progress counter (co-resident CTAs of one cooperative launch: the bound only turns a
// programming error into an error code instead of a hung GPU)
__device__ __forceinline__ void bar_sync_n(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive_n(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void syncwarp() { __syncwarp(); }
// cycle counter that stays where it is written (the memory clobber pins it against barriers and memory operations)
__device__ __forceinline__ long long clk() {
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
  return t;
}
__device__ __forceinline__ unsigned char *dyn_smem() {
  extern __shared__ __align__(16) unsigned char evp_dyn_smem[];
  return evp_dyn_smem;
}
// Tile progress counters (KERNEL_PERSISTENT).  Publishing is one release-add by one lane per warp; waiting polls with RELAXED loads
// (an acquire load invalidates the SM's L1 on every iteration; one warp per tile polls, counters sit on separate L2 lines) and turns
// into an acquire with one fence after the value has been seen.
__device__ __forceinline__ void publish_progress(unsigned *p) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned *p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void wait_progress(const unsigned *p, unsigned want, int *err) {
  if (ld_relaxed_gpu(p) < want) {
    const unsigned long long t0 = gtime();
    while (ld_relaxed_gpu(p) < want) {
      if (gtime() - t0 > 2000000000ULL) { atomicExch(err, 1); break; }
    }
  }
  fence_acq_rel_gpu();
}

#else  // EVP_HOST_EMU: see the header comment; emu:: is provided by tests/cuda_emu.h

inline unsigned long long gtime() { return 0; }
inline unsigned long long ld_acquire_sys(const unsigned long long *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void st_release_sys(unsigned long long *p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline void st_relaxed_sys(unsigned long long *p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
inline unsigned long long ld_relaxed_sys(const unsigned long long *p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
static unsigned long long g_wait_timeout_ns = ~0ULL;
inline void wait_flag(const unsigned long long *flag, unsigned long long want, int *err) {
  for (long spins = 0; ld_acquire_sys(flag) < want; ++spins) {
    if (spins > 200000000L) { *err = 1; break; }
    emu::yield();
  }
}
inline double ld_f64(const double *p) { return *(const volatile double *)p; }
inline double ld_nc_f64(const double *p) { return *p; }
inline double ld_nc_f64_pinned(const double *p) { return *p; }
inline double ld_cg_f64(const double *p) { return *(const volatile double *)p; }
inline unsigned ld_nc_u8(const unsigned char *p) { return *p; }
inline void cp_async8(double *smem, const double *g) { *smem = *g; }
inline void cp_async_commit() {}
inline void cp_async_wait_all() {}
inline void bar_arrive64(int id) { emu::bar_arrive(id, 64); }
inline void bar_sync64(int id) { emu::bar_sync(id, 64); }
inline unsigned ld_acquire_gpu(const unsigned *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void st_release_gpu(unsigned *p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }

inline void bar_sync_n(int id, int n) { emu::bar_sync(id, n); }
inline void bar_arrive_n(int id, int n) { emu::bar_arrive(id, n); }
inline void syncwarp() { emu::syncwarp(); }
inline long long clk() { return 0; }
inline unsigned char *dyn_smem() { return emu::dyn_smem(); }
inline void publish_progress(unsigned *p) { __atomic_fetch_add(p, 1u, __ATOMIC_RELEASE); }
inline void wait_progress(const unsigned *p, unsigned want, int *err) {
  for (long spins = 0; ld_acquire_gpu(p) < want; ++spins) {
    if (spins > 200000000L) { *err = 1; break; }
    emu::yield();
  }
}

#endif

// -x as a flip of the sign bit, whatever the compiler would make of `-x`: the tripole fold negates values that are often zeros
// (land, open water) and the reference's -0.0 / +0.0 must come out bit for bit
__device__ __forceinline__ double neg_f64(double x) { return __hiloint2double(__double2hiint(x) ^ (int)0x80000000, __double2loint(x)); }

}  // namespace evp
