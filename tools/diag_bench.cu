// Times the 128x128 diagonal-block Cholesky + inverse kernel (the serial critical path of the blocked
// factorisation) on a batch of SPD blocks and checks L L^T = A, L^-1 L = I.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../mip-ego_b200/csrc -I../include -o diag_bench diag_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#define DIAG_PROFILE
#include "../mip-ego_b200/csrc/linalg.cu"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

static double *dA, *dL, *dI; static int *dst;
static std::vector<double> A;
static const int batch = 8, n = 128;

// MGP_DIAG_V1 is read once per process (linalg.cu: diag_version), so the two versions are timed by
// two runs of this binary: ./diag_bench and MGP_DIAG_V1=1 ./diag_bench
int main() {
  A.resize((size_t)batch * n * n);
  srand(1);
  for (int b = 0; b < batch; b++) {
    std::vector<double> G(n * n);
    for (auto &g : G) g = rand() / (double)RAND_MAX - 0.5;
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        double s = i == j ? 1.0 : 0.0;
        for (int k = 0; k < n; k++) s += 0.05 * G[i * n + k] * G[j * n + k];
        // the last block is badly scaled (pivots from 1e-8 to 1e8) to exercise the reciprocal square root
        if (b == batch - 1) s *= pow(10.0, (i % 17 - 8) * 0.5) * pow(10.0, (j % 17 - 8) * 0.5);
        A[((size_t)b * n + i) * n + j] = s;
      }
  }
  CK(cudaMalloc(&dA, A.size() * 8)); CK(cudaMalloc(&dL, A.size() * 8)); CK(cudaMalloc(&dI, A.size() * 8));
  CK(cudaMalloc(&dst, batch * sizeof(int))); CK(cudaMemset(dst, 0, batch * sizeof(int)));
  CK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemset(dL, 0xff, A.size() * 8)); CK(cudaMemset(dI, 0xff, A.size() * 8));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  for (int i = 0; i < 5; i++) CK(launch_potrf_diag(dA, dL, n, (int64_t)n * n, dI, (int64_t)n * n, 0, dst, batch, st));
  CK(cudaStreamSynchronize(st));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int reps = 200;
  CK(cudaEventRecord(e0, st));
  for (int i = 0; i < reps; i++) CK(launch_potrf_diag(dA, dL, n, (int64_t)n * n, dI, (int64_t)n * n, 0, dst, batch, st));
  CK(cudaEventRecord(e1, st)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<double> L(A.size()), I(A.size());
  std::vector<int> stat(batch);
  CK(cudaMemcpy(L.data(), dL, L.size() * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(I.data(), dI, I.size() * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(stat.data(), dst, batch * sizeof(int), cudaMemcpyDeviceToHost));
  double e_chol = 0, e_inv = 0, e_up = 0;
  int worst_i = -1, worst_j = -1, worst_b = -1;
  for (int b = 0; b < batch; b++)
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        if (j > i) {   // both outputs carry zeros above the diagonal
          e_up = fmax(e_up, fmax(fabs(L[((size_t)b * n + i) * n + j]), fabs(I[((size_t)b * n + i) * n + j])));
          if (!(L[((size_t)b * n + i) * n + j] == 0.0) || !(I[((size_t)b * n + i) * n + j] == 0.0)) e_up = fmax(e_up, 1.0);
          continue;
        }
        double s = 0, t = 0;
        for (int k = 0; k <= j; k++) s += L[((size_t)b * n + i) * n + k] * L[((size_t)b * n + j) * n + k];
        for (int k = j; k <= i; k++) t += I[((size_t)b * n + i) * n + k] * L[((size_t)b * n + k) * n + j];
        const double sc = sqrt(A[((size_t)b * n + i) * n + i] * A[((size_t)b * n + j) * n + j]);
        const double ec = fabs(s - A[((size_t)b * n + i) * n + j]) / sc;
        if (!(ec <= e_chol)) { e_chol = ec; worst_i = i; worst_j = j; worst_b = b; }
        // L^-1 L = I, relative to the row / column scaling of the badly scaled block
        const double si = sqrt(A[((size_t)b * n + j) * n + j] / A[((size_t)b * n + i) * n + i]);
        const double ei = fabs(t - (i == j ? 1.0 : 0.0)) * fmin(1.0, si);
        if (!(ei <= e_inv)) e_inv = ei;
      }
  long long clk[16];
  CK(cudaMemcpyFromSymbol(clk, g_diag_clk, sizeof(clk)));
  if (diag_version() == 1)
    printf("v1 cycles: load %lld | potrf32 %lld trtri32 %lld | panel(s=0) %lld update(s=0) %lld | steps1-3 %lld | storeL %lld | inverse %lld | storeInv %lld | total %lld\n",
           clk[1] - clk[0], clk[8] - clk[1], clk[2] - clk[8], clk[9] - clk[2], clk[3] - clk[9], clk[4] - clk[3], clk[5] - clk[4], clk[6] - clk[5], clk[7] - clk[6], clk[7] - clk[0]);
  else
    printf("v2 cycles: load %lld | factor + sub-panel (s=0) %lld | next diagonal sub-block (s=0) %lld | steps1-3 %lld | trtri32(3)+storeL %lld | inverse row 3 %lld | storeInv %lld | total %lld\n",
           clk[1] - clk[0], clk[8] - clk[1], clk[2] - clk[8], clk[4] - clk[2], clk[5] - clk[4], clk[6] - clk[5], clk[7] - clk[6], clk[7] - clk[0]);
  printf("status:");
  for (int b = 0; b < batch; b++) printf(" %d", stat[b]);
  printf("   worst LLt element: block %d (%d, %d)\n", worst_b, worst_i, worst_j);
  printf("{\"version\": %d, \"potrf_diag_us\": %.2f, \"batch\": %d, \"max_rel_err_LLt\": %.3e, \"max_err_LinvL\": %.3e, \"max_abs_upper\": %.3e}\n",
         diag_version(), ms * 1e3 / reps, batch, e_chol, e_inv, e_up);
  return 0;
}
