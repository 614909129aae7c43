// evp_halo_local.cuh -- the halo update of (uvel,vvel) when every source is on this rank (one GPU, tripole fold): halo_pack and
// halo_apply of evp_halo.cu as ONE kernel of one CTA (EVP_B200_HALO_FUSED=1; round-2 candidate, not yet measured).
//
// Every output of the update is a function of pre-update values only (evp_halo.cu), so a thread first reads the sources of all
// its entries into registers, the CTA synchronises, and only then the destinations are written -- the staging buffer and one
// kernel boundary per subcycle disappear, and the kernel joins the programmatic-dependent-launch chain of the subcycle kernels
// (tx1 on one GPU: 3 launches per subcycle, 11.8 us, against 9.2 us at gx1 with more cells).  Up to HALO_LOCAL_MAX entries.
// Plain C++ apart from the qualifiers, so the host emulation (tests/emu_bgrid.cpp) can run it against the plan executor.
#pragma once

namespace evp {

constexpr int HALO_LOCAL_THREADS = 1024, HALO_LOCAL_PER_THREAD = 8, HALO_LOCAL_MAX = HALO_LOCAL_THREADS * HALO_LOCAL_PER_THREAD;

// codes as in evp_halo.cu: 0 copy, 1 negate, 2 0.5*(a-b), 3 -(0.5*(a-b))
__global__ void __launch_bounds__(HALO_LOCAL_THREADS) halo_local_kernel(double *U, double *V, const int *__restrict__ dst,
                                                                       const int *__restrict__ c1, const int *__restrict__ c2,
                                                                       const signed char *__restrict__ code, int n, int pdl) {
#ifndef EVP_HOST_EMU
  if (pdl) {
    cudaTriggerProgrammaticLaunchCompletion();
    cudaGridDependencySynchronize();
  }
#endif
  double u[HALO_LOCAL_PER_THREAD], v[HALO_LOCAL_PER_THREAD];
#pragma unroll
  for (int q = 0; q < HALO_LOCAL_PER_THREAD; ++q) {
    const int k = threadIdx.x + q * HALO_LOCAL_THREADS;
    u[q] = 0.0; v[q] = 0.0;
    if (k < n) {
      const int a = c1[k], op = code[k];
      u[q] = U[a]; v[q] = V[a];
      if (op == 1) {
        u[q] = -u[q]; v[q] = -v[q];
      } else if (op == 2 || op == 3) {
        // the sign of a zero result matters: see halo_apply
        const int b = c2[k];
        u[q] = 0.5 * (u[q] - U[b]);
        v[q] = 0.5 * (v[q] - V[b]);
        if (op == 3) { u[q] = -u[q]; v[q] = -v[q]; }
      }
    }
  }
  __syncthreads();  // every source has been read
#pragma unroll
  for (int q = 0; q < HALO_LOCAL_PER_THREAD; ++q) {
    const int k = threadIdx.x + q * HALO_LOCAL_THREADS;
    if (k < n) { U[dst[k]] = u[q]; V[dst[k]] = v[q]; }
  }
}

}  // namespace evp
