// FP64 peak micro-benchmark for B200 (sm_100a): DFMA vs DMMA (mma.sync m8n8k4 / m16n8k16 f64), plus
// vendor-library yardsticks on the same box ("the honest on-GPU bar"):
// cuBLAS DGEMM 8192^3, cuSOLVER DPOTRF + DPOTRI at n = 4096 (the dense part of one likelihood
// evaluation), cuBLAS DTRMM with a 4096 x 4096 triangle against 151 552 right-hand sides (the shape
// of k_trmm at C3).  TOOL ONLY: the product library links neither cuBLAS nor cuSOLVER.
// Prints one JSON line.  Used once per pool to obtain the FP64 roofline denominator that
// MEASURED_PEAKS.json lacks.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu -lcublas -lcusolver
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>

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

// Every launch is checked: a launch that fails (too many registers for the block size, ...) must
// abort the program, not be timed as an infinitely fast kernel.
template <typename F>
static double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); CK(cudaGetLastError());
  launch(); CK(cudaGetLastError());
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

#define CKB(x) do { cublasStatus_t s_ = (x); if (s_ != CUBLAS_STATUS_SUCCESS) { \
  fprintf(stderr, "cuBLAS error %d at %s:%d\n", (int)s_, __FILE__, __LINE__); exit(1);} } while (0)
#define CKS(x) do { cusolverStatus_t s_ = (x); if (s_ != CUSOLVER_STATUS_SUCCESS) { \
  fprintf(stderr, "cuSOLVER error %d at %s:%d\n", (int)s_, __FILE__, __LINE__); exit(1);} } while (0)

__global__ void k_fill_spd(double *A, int n) {   // diagonally dominant symmetric matrix
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)n * n) return;
  const int i = (int)(e / n), j = (int)(e % n);
  A[e] = i == j ? 2.0 : 1.0 / (1.0 + (double)(i > j ? i - j : j - i));
}
__global__ void k_fill(double *A, int64_t cnt, double v) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < cnt) A[e] = v + 1e-3 * (double)(e % 17);
}

static void vendor_yardsticks() {
  cublasHandle_t cb; CKB(cublasCreate(&cb));
  cusolverDnHandle_t cs; CKS(cusolverDnCreate(&cs));
  {  // DGEMM 8192^3
    const int n = 8192;
    double *A, *B, *C;
    CK(cudaMalloc(&A, sizeof(double) * n * n)); CK(cudaMalloc(&B, sizeof(double) * n * n)); CK(cudaMalloc(&C, sizeof(double) * n * n));
    k_fill<<<(n * n + 255) / 256, 256>>>(A, (int64_t)n * n, 0.5); k_fill<<<(n * n + 255) / 256, 256>>>(B, (int64_t)n * n, 0.25);
    CK(cudaGetLastError());
    const double one = 1.0, zero = 0.0;
    double ms = time_ms([&] { CKB(cublasDgemm(cb, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &one, A, n, B, n, &zero, C, n)); }, 5);
    printf(", \"cublas_dgemm_8192_tflops\": %.2f", 2.0 * n * (double)n * n / ms * 1e-9);
    CK(cudaFree(A)); CK(cudaFree(B)); CK(cudaFree(C));
  }
  {  // DPOTRF + DPOTRI, n = 4096, 8 matrices one after the other (cuSOLVER has no batched large-n variant)
    const int n = 4096, nb = 8;
    double *A; CK(cudaMalloc(&A, sizeof(double) * n * n * nb));
    int *info; CK(cudaMalloc(&info, sizeof(int)));
    int l1 = 0, l2 = 0;
    CKS(cusolverDnDpotrf_bufferSize(cs, CUBLAS_FILL_MODE_LOWER, n, A, n, &l1));
    CKS(cusolverDnDpotri_bufferSize(cs, CUBLAS_FILL_MODE_LOWER, n, A, n, &l2));
    const int lw = l1 > l2 ? l1 : l2;
    double *work; CK(cudaMalloc(&work, sizeof(double) * lw));
    auto fill = [&] { for (int b = 0; b < nb; b++) k_fill_spd<<<(n * n + 255) / 256, 256>>>(A + (size_t)b * n * n, n); };
    double t_f = 1e30, t_i = 1e30;
    cudaEvent_t e0, e1, e2; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    for (int rep = 0; rep < 4; rep++) {
      fill(); CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      for (int b = 0; b < nb; b++) CKS(cusolverDnDpotrf(cs, CUBLAS_FILL_MODE_LOWER, n, A + (size_t)b * n * n, n, work, lw, info));
      CK(cudaEventRecord(e1));
      for (int b = 0; b < nb; b++) CKS(cusolverDnDpotri(cs, CUBLAS_FILL_MODE_LOWER, n, A + (size_t)b * n * n, n, work, lw, info));
      CK(cudaEventRecord(e2)); CK(cudaEventSynchronize(e2));
      float a, c; CK(cudaEventElapsedTime(&a, e0, e1)); CK(cudaEventElapsedTime(&c, e1, e2));
      if (rep > 0) { if (a < t_f) t_f = a; if (c < t_i) t_i = c; }
    }
    int hinfo = -1; CK(cudaMemcpy(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost));
    const double n3 = (double)n * n * n;
    printf(", \"cusolver_dpotrf_4096_x8_ms\": %.3f, \"cusolver_dpotrf_tflops\": %.2f, \"cusolver_dpotri_4096_x8_ms\": %.3f, "
           "\"cusolver_dpotri_tflops\": %.2f, \"cusolver_potrf_potri_evals_per_s\": %.1f, \"cusolver_info\": %d",
           t_f, nb * n3 / 3.0 / t_f * 1e-9, t_i, nb * 2.0 * n3 / 3.0 / t_i * 1e-9, nb / ((t_f + t_i) * 1e-3), hinfo);
    CK(cudaFree(A)); CK(cudaFree(work)); CK(cudaFree(info));
  }
  {  // DTRMM: lower-triangular 4096 x 4096 times 4096 x 151 552 (out of place)
    const int n = 4096, m = 151552;
    double *A, *B, *C;
    CK(cudaMalloc(&A, sizeof(double) * n * n)); CK(cudaMalloc(&B, sizeof(double) * (size_t)n * m)); CK(cudaMalloc(&C, sizeof(double) * (size_t)n * m));
    k_fill<<<(n * n + 255) / 256, 256>>>(A, (int64_t)n * n, 0.5);
    k_fill<<<(unsigned)(((int64_t)n * m + 255) / 256), 256>>>(B, (int64_t)n * m, 0.25);
    CK(cudaGetLastError());
    const double one = 1.0;
    double ms = time_ms([&] { CKB(cublasDtrmm(cb, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, n, m,
                                              &one, A, n, B, n, C, n)); }, 3);
    printf(", \"cublas_dtrmm_4096x151552_ms\": %.3f, \"cublas_dtrmm_tflops\": %.2f", ms, (double)n * n * m / ms * 1e-9);
    CK(cudaFree(A)); CK(cudaFree(B)); CK(cudaFree(C));
  }
  cublasDestroy(cb); cusolverDnDestroy(cs);
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
    // at most 512 threads per block: these kernels hold 32-64 accumulator registers per thread, a
    // 1024-thread block does not launch (round 1 recorded such a failed launch as 62 861 TFLOP/s)
    int threads = warps * 32 > 512 ? 512 : warps * 32;
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
  vendor_yardsticks();
  printf("}\n");
  return 0;
}
