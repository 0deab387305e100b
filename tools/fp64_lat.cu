// Dependent-issue latency of DMMA.8x8x4 and DFMA on one warp (cycles per step of a dependent chain)
// and the issue interval of independent DMMAs.  Motivates the loop order of K1 (posterior.cu): the
// asm statements of the MMA wrapper are volatile (issued in source order), so independent
// accumulator chains have to be interleaved in the source.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o fp64_lat fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void k_dmma(double *out, long long *cyc, int iters) {
  double c[CHAINS][2];
  for (int i = 0; i < CHAINS; i++) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
  const double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) dmma884(c[i][0], c[i][1], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < CHAINS; i++) s += c[i][0] + c[i][1];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void k_dfma(double *out, long long *cyc, int iters) {
  double x = threadIdx.x * 1e-3;
  const double a = 1.0 + 1e-9, b = 1e-9;
  const long long t0 = clock64();
#pragma unroll 16
  for (int it = 0; it < iters; it++) x = fma(x, a, b);
  const long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int CHAINS>
double run_dmma(double *out, long long *cyc, int iters) {
  k_dmma<CHAINS><<<1, 32>>>(out, cyc, iters);
  k_dmma<CHAINS><<<1, 32>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  return (double)h / iters;
}

int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 32 * sizeof(double)); cudaMalloc(&cyc, sizeof(long long));
  const int iters = 4096;
  const double d1 = run_dmma<1>(out, cyc, iters), d2 = run_dmma<2>(out, cyc, iters), d4 = run_dmma<4>(out, cyc, iters),
               d8 = run_dmma<8>(out, cyc, iters), d16 = run_dmma<16>(out, cyc, iters);
  k_dfma<<<1, 32>>>(out, cyc, iters); k_dfma<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("{\"dmma884_dependent_cycles\": %.1f, \"cycles_per_round_2_chains\": %.1f, \"4_chains\": %.1f, \"8_chains\": %.1f, "
         "\"16_chains\": %.1f, \"dmma_issue_interval_cycles\": %.1f, \"dfma_dependent_cycles\": %.1f}\n",
         d1, d2, d4, d8, d16, d16 / 16.0, (double)h / iters);
  return 0;
}
