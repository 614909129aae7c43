// fp64 pipe microbenchmark for sm_100a: dependent-chain latency and per-SMSP throughput of DADD/DMUL/DFMA, and the
// cost of IEEE division and square root.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP, int ILP>
__global__ void k(double *out, double a, double b, int n, long long *cyc) {
  double x[ILP];
  for (int q = 0; q < ILP; ++q) x[q] = a + q + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int q = 0; q < ILP; ++q) {
      if (OP == 0) x[q] = x[q] + b;
      if (OP == 1) x[q] = x[q] * b;
      if (OP == 2) x[q] = fma(x[q], b, a);
      if (OP == 3) x[q] = a / x[q];
      if (OP == 4) x[q] = sqrt(x[q] + b);
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int q = 0; q < ILP; ++q) s += x[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP, int ILP>
void run(const char *name, int threads, double b) {
  double *out; long long *cyc, h;
  cudaMalloc(&out, 8 * 2048); cudaMalloc(&cyc, 8);
  const int n = 2000;
  k<OP, ILP><<<1, threads>>>(out, 1.000001, b, n, cyc);
  k<OP, ILP><<<1, threads>>>(out, 1.000001, b, n, cyc);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-6s ILP=%d warps/SMSP=%.2f: %.2f cycles per op per warp-chain, %.3f warp-ops/clk/SMSP\n", name, ILP, threads / 128.0,
         (double)h / n, (double)n * ILP * (threads / 32) / 4.0 / (double)h);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int th : {32, 128, 256, 512, 1024}) {
    run<0, 1>("DADD", th, 1e-9); run<0, 4>("DADD", th, 1e-9);
    run<1, 1>("DMUL", th, 1.0000001); run<1, 4>("DMUL", th, 1.0000001);
    run<2, 1>("DFMA", th, 0.999999); run<2, 4>("DFMA", th, 0.999999);
    run<3, 1>("DDIV", th, 0); run<3, 4>("DDIV", th, 0);
    run<4, 1>("DSQRT", th, 1e-3); run<4, 4>("DSQRT", th, 1e-3);
  }
  return 0;
}
