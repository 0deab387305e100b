// Do DMMA (mma.sync m8n8k4 f64) and DFMA issue to the same FP64 pipe on B200?  Times a fixed DMMA
// stream per warp with NF extra independent DFMAs interleaved per iteration; if the pipes were
// separate the time would stay flat until the DFMA stream alone saturates.  Prints one JSON line.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_mix fp64_mix.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// per iteration: ND DMMAs (independent accumulators) + NF DFMAs (independent accumulators)
template <int ND, int NF>
__global__ void k_mix(double *out, int iters, double a0, double b0) {
  double c[ND > 0 ? ND : 1][2];
  double f[NF > 0 ? NF : 1];
#pragma unroll
  for (int i = 0; i < ND; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
#pragma unroll
  for (int i = 0; i < NF; i++) f[i] = threadIdx.x + i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; it++) {
    constexpr int R = (ND > 0 && NF > 0) ? (NF / ND) : 0;
#pragma unroll
    for (int i = 0; i < ND; i++) {
      dmma884(c[i][0], c[i][1], a, b);
#pragma unroll
      for (int j = 0; j < R; j++) f[i * R + j] = fma(f[i * R + j], a, b);
    }
    if (ND == 0) {
#pragma unroll
      for (int i = 0; i < NF; i++) f[i] = fma(f[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ND; i++) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < NF; i++) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// warp-specialised mix: even warps DMMA, odd warps DFMA
template <int ND, int NF>
__global__ void k_split(double *out, int iters, double a0, double b0, int dfma_on) {
  const int warp = threadIdx.x >> 5;
  double a = a0 + threadIdx.x * 1e-9, b = b0, s = 0;
  if ((warp & 1) == 0) {
    double c[ND][2];
#pragma unroll
    for (int i = 0; i < ND; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; it++)
#pragma unroll
      for (int i = 0; i < ND; i++) dmma884(c[i][0], c[i][1], a, b);
#pragma unroll
    for (int i = 0; i < ND; i++) s += c[i][0] + c[i][1];
  } else if (dfma_on) {
    double f[NF];
#pragma unroll
    for (int i = 0; i < NF; i++) f[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++)
#pragma unroll
      for (int i = 0; i < NF; i++) f[i] = fma(f[i], a, b);
#pragma unroll
    for (int i = 0; i < NF; i++) s += f[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch();
  CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaGetLastError());
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 4 * 1024));
  const int iters = 8192, threads = 256, blocks = sms * 2;
  const double wl = (double)blocks * (threads / 32) * iters;
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
#define MIX(ND, NF)                                                                              \
  {                                                                                              \
    double ms = time_ms([&] { k_mix<ND, NF><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5); \
    printf(", \"mix_d%d_f%d\": {\"ms\": %.3f, \"dmma_tf\": %.2f, \"dfma_tf\": %.2f}", ND, NF, ms,  \
           wl * ND * 512.0 / ms * 1e-9, wl * NF * 64.0 / ms * 1e-9);                             \
  }
  MIX(16, 0) MIX(0, 64) MIX(16, 16) MIX(16, 32) MIX(16, 64)
  {
    double m0 = time_ms([&] { k_split<16, 64><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9, 0); }, 5);
    double m1 = time_ms([&] { k_split<16, 64><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9, 1); }, 5);
    printf(", \"split_dmma_only_ms\": %.3f, \"split_both_ms\": %.3f", m0, m1);
  }
  printf("}\n");
  return 0;
}
