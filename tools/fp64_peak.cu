// FP64 peak micro-benchmark for B200 (sm_100a): DFMA vs DMMA (mma.sync m8n8k4 / m16n8k16 f64).
// Prints one JSON line.  Used once per pool to obtain the FP64 roofline denominator that
// MEASURED_PEAKS.json lacks.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
               "{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NACC>
__global__ void k_dmma884(double *out, int iters, double a0, double b0) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dmma16816(double *out, int iters, double a0, double b0) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = a0 + i * 1e-9;
#pragma unroll
  for (int i = 0; i < 4; i++) b[i] = b0 + i * 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma16816(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double *out, int iters, double a0, double b0) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; i++) c[i] = threadIdx.x + i;
  double a = a0, b = b0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i];
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
  double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 16 * 1024));
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
  const int iters = 4096;
  int wps[] = {4, 8, 16, 32};
  for (int wi = 0; wi < 4; wi++) {
    int warps = wps[wi];
    int threads = warps * 32 > 1024 ? 1024 : warps * 32;
    int blocks = sms * (warps * 32 / threads);
    {
      double ms = time_ms([&] { k_dmma884<16><<<blocks, threads>>>(out, iters, 1.0000001, 0.999999); }, 5);
      double fl = (double)blocks * (threads / 32) * iters * 16 * 512.0;
      printf(", \"dmma884_w%d_tflops\": %.2f", warps, fl / ms * 1e-9);
    }
    {
      double ms = time_ms([&] { k_dmma16816<8><<<blocks, threads>>>(out, iters / 4, 1.0000001, 0.999999); }, 5);
      double fl = (double)blocks * (threads / 32) * (iters / 4) * 8 * 4096.0;
      printf(", \"dmma16816_w%d_tflops\": %.2f", warps, fl / ms * 1e-9);
    }
    {
      double ms = time_ms([&] { k_dfma<16><<<blocks, threads>>>(out, iters * 4, 1.0000001, 1e-9); }, 5);
      double fl = (double)blocks * threads * (double)(iters * 4) * 16 * 2.0;
      printf(", \"dfma_w%d_tflops\": %.2f", warps, fl / ms * 1e-9);
    }
  }
  // sustained DMMA for ~2 s (power-capped clocks)
  {
    int threads = 512, blocks = sms * 2;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    int n = 0;
    for (; n < 200; n++) k_dmma884<16><<<blocks, threads>>>(out, iters * 8, 1.0000001, 0.999999);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    double fl = (double)n * blocks * (threads / 32) * (double)(iters * 8) * 16 * 512.0;
    printf(", \"dmma884_sustained_tflops\": %.2f, \"sustained_ms\": %.1f", fl / ms * 1e-9, ms);
  }
  printf("}\n");
  return 0;
}
