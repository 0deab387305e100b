// Checkpoint 1 of the INT8 (Ozaki-scheme) prototype: one tcgen05.mma.kind::i8 GEMM tile with TMEM
// accumulators and bulk-copy staged operands, checked against a CPU int32 product.
//   D[m][n] = sum_k A[m][k] * B[n][k],  A: 128 x K int8, B: N x K int8 (both K-major), D int32.
// Operands are PRE-PACKED in global memory in the shared-memory image the UMMA descriptors expect
// (no-swizzle "interleaved" K-major canonical layout: 8 rows x 16 bytes core matrices), so that a
// stage is one 1-D cp.async.bulk per operand -- no tensor maps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ozaki/umma_i8_test tools/ozaki/umma_i8_test.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int M = 128;
constexpr int BK = 64;            // K bytes per stage

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int cnt) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(cnt));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t phase) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(b)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 B (128 B contiguous); LBO = byte distance between the
// two 16-byte K chunks of one MMA, SBO = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;          // descriptor version (Blackwell)
  return d;                        // layout_type 0 = SWIZZLE_NONE, base_offset 0
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int N>
__global__ void __launch_bounds__(128) k_umma_test(const int8_t *Ap, const int8_t *Bp, int K, int32_t *D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *sA = smem;                       // [K/16][16 groups][8][16]  = M * K bytes
  uint8_t *sB = smem + M * BK * 2;          // two stages of A then two of B
  __shared__ uint64_t full[2], empty[2], done;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_init(&empty[0], 1); mbar_init(&empty[1], 1); mbar_init(&done, 1);
                  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(N < 32 ? 32 : N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base;
  const int nst = K / BK;
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  if (tid == 0) {
    // producer + MMA issuer in one thread (two-stage ring): enough for a correctness check
    for (int s = 0; s < nst; s++) {
      const int st = s & 1;
      if (s >= 2) mbar_wait(&empty[st], (uint32_t)(((s >> 1) - 1) & 1));   // the MMAs that read this stage are done
      mbar_expect_tx(&full[st], (M + N) * BK);
      bulk_g2s(sA + st * M * BK, Ap + (size_t)s * M * BK, M * BK, &full[st]);
      bulk_g2s(sB + st * N * BK, Bp + (size_t)s * N * BK, N * BK, &full[st]);
      mbar_wait(&full[st], (uint32_t)((s >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;");
      for (int k = 0; k < BK / 32; k++) {
        const uint64_t da = umma_desc(smem_u32(sA + st * M * BK) + k * 2 * (M * 16), M * 16, 128);
        const uint64_t db = umma_desc(smem_u32(sB + st * N * BK) + k * 2 * (N * 16), N * 16, 128);
        umma_i8(tmem, da, db, idesc, (s | k) ? 1u : 0u);
      }
      umma_commit(&empty[st]);   // arrives when the MMAs issued so far have completed
    }
    umma_commit(&done);
  }
  __syncthreads();
  mbar_wait(&done, 0u);
  asm volatile("tcgen05.fence::after_thread_sync;");
  // epilogue: warp w reads TMEM lanes 32w .. 32w+31 (rows), 8 columns at a time
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; j++) D[(size_t)tid * N + c0 + j] = (int32_t)v[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(N < 32 ? 32 : N));
}

// host packing: element (r, k) of an R x K K-major operand, stage s = k / BK:
//   stage base s * R * BK; inside: (kk / 16) * (R * 16) + (r / 8) * 128 + (r % 8) * 16 + kk % 16,  kk = k % BK
static void pack(const std::vector<int8_t> &src, int R, int K, std::vector<int8_t> &dst) {
  dst.assign((size_t)R * K, 0);
  for (int r = 0; r < R; r++)
    for (int k = 0; k < K; k++) {
      const int s = k / BK, kk = k % BK;
      dst[(size_t)s * R * BK + (size_t)(kk / 16) * (R * 16) + (r / 8) * 128 + (r % 8) * 16 + kk % 16] = src[(size_t)r * K + k];
    }
}

template <int N>
static int run(int K) {
  std::vector<int8_t> A((size_t)M * K), B((size_t)N * K), Ap, Bp;
  srand(1234 + N + K);
  for (auto &v : A) v = (int8_t)(rand() % 129 - 64);
  for (auto &v : B) v = (int8_t)(rand() % 129 - 64);
  pack(A, M, K, Ap); pack(B, N, K, Bp);
  int8_t *dA, *dB; int32_t *dD;
  CK(cudaMalloc(&dA, Ap.size())); CK(cudaMalloc(&dB, Bp.size())); CK(cudaMalloc(&dD, sizeof(int32_t) * M * N));
  CK(cudaMemcpy(dA, Ap.data(), Ap.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bp.data(), Bp.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, sizeof(int32_t) * M * N));
  const int smem = 2 * (M + N) * BK + 1024;
  CK(cudaFuncSetAttribute(k_umma_test<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_umma_test<N><<<1, 128, smem>>>(dA, dB, K, dD);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> D((size_t)M * N);
  CK(cudaMemcpy(D.data(), dD, sizeof(int32_t) * M * N, cudaMemcpyDeviceToHost));
  long bad = 0; int first = -1;
  for (int m = 0; m < M; m++)
    for (int n = 0; n < N; n++) {
      int32_t ref = 0;
      for (int k = 0; k < K; k++) ref += (int32_t)A[(size_t)m * K + k] * (int32_t)B[(size_t)n * K + k];
      if (ref != D[(size_t)m * N + n]) { if (first < 0) first = m * N + n; bad++; }
    }
  printf("{\"N\": %d, \"K\": %d, \"mismatches\": %ld, \"first\": %d, \"D0\": %d}\n", N, K, bad, first, D[0]);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return bad != 0;
}

int main() {
  int rc = 0;
  rc |= run<64>(64);
  rc |= run<64>(256);
  rc |= run<128>(512);
  rc |= run<256>(128);
  return rc;
}
