// evp_tma.cuh -- the Tensor Memory Accelerator pieces the tile-streaming kernel (evp_tstream.cu) is built from: tensor maps,
// cp.async.bulk.tensor.2d box loads global -> shared that complete on an mbarrier, and the mbarrier itself.
//
// With EVP_HOST_EMU defined (tests/emu_tstream.cpp) a tensor map is a plain description of the array and a box load is a
// synchronous copy with the hardware's out-of-bounds rule (elements outside the tensor arrive as zeros); the mbarrier is a
// counter of completed phases.  The product never defines it.
#pragma once
#include "evp_internal.h"

#ifndef EVP_HOST_EMU
#include <cuda.h>
#endif

namespace evp {

#ifndef EVP_HOST_EMU

struct alignas(64) TmaMap { CUtensorMap m; };
typedef unsigned long long MBar;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
// first 128-byte boundary of the dynamic shared memory; pointer arithmetic on the array itself, so that the compiler keeps
// addressing it as shared memory (LDS/STS, not generic LD/ST)
__device__ __forceinline__ unsigned char *smem_align128(unsigned char *base) { return base + ((128u - (smem_u32(base) & 127u)) & 127u); }
__device__ __forceinline__ void mbar_init(MBar *bar, int arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
// makes the initialised barriers visible to the asynchronous proxy (the TMA unit) before the first box load names them
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// shared memory last touched by ordinary loads/stores is handed to the asynchronous proxy
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// one arrival that also announces `bytes` of box loads: the phase completes when they have all landed
__device__ __forceinline__ void mbar_arrive_expect_tx(MBar *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// box load: element (x, y) of the tensor is the first element of the box; rows of the box are dense in shared memory
__device__ __forceinline__ void tma_load_2d(void *dst, const TmaMap *map, int x, int y, MBar *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
               : "memory");
}
// wait for phase number `phase` (0, 1, 2 ... in the order the barrier completes them).  Bounded like every other in-kernel wait
// of this library: a programming error (byte count that never arrives) becomes an error code, not a hung GPU.
__device__ __forceinline__ void mbar_wait(MBar *bar, unsigned phase, int *err) {
  const unsigned addr = smem_u32(bar), parity = phase & 1u;
  unsigned done = 0;
  unsigned long long t0 = 0;
  for (unsigned spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (!done && (spins & 15u) == 15u) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000ULL) { atomicExch(err, 1); break; }
    }
  }
}

#else  // EVP_HOST_EMU

struct TmaMap {
  const void *base;
  int elem;                 // bytes per element
  int dim0, dim1;           // tensor extent in elements
  long long pitch;          // bytes between rows
  int box0, box1;           // box extent in elements
};
// arrivals and transaction bytes counted like the hardware does (several threads issue the loads of one phase, one of them arrives):
// phase p is complete when p+1 arrivals have been made and the bytes announced up to then have landed
struct MBar {
  std::atomic<unsigned> arrivals;
  std::atomic<long long> landed;      // bytes copied so far (all phases)
  std::atomic<long long> announced;   // bytes announced so far
  std::atomic<long long> due[4];      // `announced` after arrival number p (ring by p & 3)
};
static_assert(sizeof(MBar) <= 64, "two barriers share 128 bytes of the CTA's shared memory");

inline unsigned char *smem_align128(unsigned char *base) { return (unsigned char *)(((uintptr_t)base + 127) & ~(uintptr_t)127); }
inline void mbar_init(MBar *bar, int) {
  bar->arrivals.store(0); bar->landed.store(0); bar->announced.store(0);
  for (auto &d : bar->due) d.store(0);
}
inline void mbar_fence_init() {}
inline void fence_proxy_async() {}
inline void mbar_arrive_expect_tx(MBar *bar, unsigned bytes) {
  const long long a = bar->announced.fetch_add(bytes) + bytes;
  bar->due[bar->arrivals.load() & 3].store(a);
  bar->arrivals.fetch_add(1, std::memory_order_release);
}
// the hardware's rules that a plain copy would not notice: the box must start on a 16-byte boundary of its row (B200: anything else
// is an illegal instruction), its rows are a multiple of 16 bytes, the destination is 128-byte aligned.  (The row pitch is a multiple of 16
// bytes in the product -- Dom::ld is a multiple of 16 cells -- but not in the emulation, where one block of the caller IS the dom.)
static std::atomic<int> emu_tma_misaligned{0};
inline void tma_load_2d(void *dst, const TmaMap *m, int x, int y, MBar *bar) {
  if (((long long)x * m->elem) % 16 != 0 || (m->box0 * m->elem) % 16 != 0 || ((uintptr_t)dst & 127) != 0) emu_tma_misaligned.store(1);
  unsigned char *o = (unsigned char *)dst;
  for (int r = 0; r < m->box1; ++r)
    for (int c = 0; c < m->box0; ++c, o += m->elem) {
      const int gx = x + c, gy = y + r;
      if (gx >= 0 && gx < m->dim0 && gy >= 0 && gy < m->dim1) memcpy(o, (const unsigned char *)m->base + gy * m->pitch + (long long)gx * m->elem, m->elem);
      else memset(o, 0, m->elem);
    }
  bar->landed.fetch_add((long long)m->box0 * m->box1 * m->elem, std::memory_order_release);
}
inline void mbar_wait(MBar *bar, unsigned phase, int *err) {
  for (long spins = 0;; ++spins) {
    if (bar->arrivals.load(std::memory_order_acquire) > phase && bar->landed.load(std::memory_order_acquire) >= bar->due[phase & 3].load()) break;
    if (spins > 200000000L) { *err = 1; break; }
    emu::yield();
  }
}

#endif

struct TsMaps { TmaMap m[TS_NMAPS]; };   // every tensor map of KERNEL_TSTREAM, one kernel parameter

}  // namespace evp
