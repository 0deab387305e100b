// Checkpoint 1.5: what rate can the planned Ozaki tile sustain?  Persistent CTAs, warp-specialised:
// one producer lane streams S slices of the 128-candidate operand and S slices of the NROW-row L^-1
// operand per K = 64 stage from an L2-resident buffer (bulk copies), one MMA lane issues the
// NPAIR = S (S + 1) / 2 int8 MMAs per K = 32 step into S level accumulators in TMEM.  No epilogue.
// Prints cycles per stage; the FP64 DMMA path needs 128 * NROW * 64 / 64 cycles for the same work.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
constexpr int M = 128, BK = 64;
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int cnt) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(cnt)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t phase) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(b)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void umma_i8(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }

template <int NROW, int S, int NST, bool STACK = false>
__global__ void __launch_bounds__(128) k_rate(const int8_t *buf, size_t buf_bytes, int nstage, long long *clk) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int A_BYTES = S * M * BK, B_BYTES = S * NROW * BK, ST_BYTES = A_BYTES + B_BYTES;
  __shared__ uint64_t full[NST], empty[NST], done;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < NST; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base;
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NROW >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  const long long t0 = clock64();
  if (warp == 1 && (tid & 31) == 0) {          // producer
    size_t off = ((size_t)blockIdx.x * 7919 * 1024) % (buf_bytes - ST_BYTES);
    off &= ~(size_t)1023;
    for (int s = 0; s < nstage; s++) {
      const int st = s % NST;
      if (s >= NST) mbar_wait(&empty[st], (uint32_t)(((s / NST) - 1) & 1));
      mbar_expect_tx(&full[st], ST_BYTES);
      bulk_g2s(smem + st * ST_BYTES, buf + off, A_BYTES, &full[st]);
      bulk_g2s(smem + st * ST_BYTES + A_BYTES, buf + off + A_BYTES, B_BYTES, &full[st]);
      off += ST_BYTES;
      if (off + ST_BYTES > buf_bytes) off = 0;
    }
  } else if (warp == 2 && (tid & 31) == 0) {   // MMA issuer
    for (int s = 0; s < nstage; s++) {
      const int st = s % NST;
      mbar_wait(&full[st], (uint32_t)((s / NST) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;");
      const uint32_t a0 = smem_u32(smem + st * ST_BYTES), b0 = a0 + A_BYTES;
      if (STACK) {
        // slices of the row operand stacked along N: one MMA covers the levels p .. p + cnt - 1
        // (B image: 16-byte K chunk -> S slices x NROW rows contiguous, LBO = S * NROW * 16)
        for (int k = 0; k < BK / 32; k++)
          for (int p = 0; p < S; p++) {
            int q = 0;
            while (q < S - p) {
              int cnt = S - p - q;
              if (cnt * NROW > 256) cnt = 256 / NROW;
              const uint32_t id2 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((cnt * NROW) >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
              const uint64_t da = umma_desc(a0 + p * (M * BK) + k * 2 * (M * 16), M * 16, 128);
              const uint64_t db = umma_desc(b0 + k * 2 * (S * NROW * 16) + q * (NROW * 16), S * NROW * 16, 128);
              umma_i8(tmem + (uint32_t)((p + q) * NROW), da, db, id2, (s | k | p) ? 1u : 0u);
              q += cnt;
            }
          }
      } else
      for (int k = 0; k < BK / 32; k++)
        for (int p = 0; p < S; p++)
          for (int q = 0; p + q < S; q++) {
            const uint64_t da = umma_desc(a0 + p * (M * BK) + k * 2 * (M * 16), M * 16, 128);
            const uint64_t db = umma_desc(b0 + q * (NROW * BK) + k * 2 * (NROW * 16), NROW * 16, 128);
            umma_i8(tmem + (uint32_t)((p + q) * NROW), da, db, idesc, (s | k) ? 1u : ((p == 0 || q == 0) && (p + q >= 0) ? (uint32_t)(p > 0 && q > 0) : 1u));
          }
      umma_commit(&empty[st]);
    }
    umma_commit(&done);
    mbar_wait(&done, 0u);
  }
  __syncthreads();
  if (tid == 0) clk[blockIdx.x] = clock64() - t0;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <int NROW, int S, int NST, bool STACK = false>
static void run(int8_t *buf, size_t bytes, int grid) {
  const int nstage = 2000;
  constexpr int ST_BYTES = S * (M + NROW) * BK;
  const int smem = NST * ST_BYTES + 1024;
  long long *dclk; CK(cudaMalloc(&dclk, sizeof(long long) * grid));
  CK(cudaFuncSetAttribute(k_rate<NROW, S, NST, STACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_rate<NROW, S, NST, STACK><<<grid, 128, smem>>>(buf, bytes, 100, dclk); CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k_rate<NROW, S, NST, STACK><<<grid, 128, smem>>>(buf, bytes, nstage, dclk); CK(cudaGetLastError());
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  long long h[148]; CK(cudaMemcpy(h, dclk, sizeof(long long) * (grid < 148 ? grid : 148), cudaMemcpyDeviceToHost));
  const double npair = S * (S + 1) / 2.0;
  const double clk_stage = (double)h[0] / nstage;
  const double dmma_clk = 128.0 * NROW * 64 / 64.0;
  const double mac_rate = npair * 2 * 128.0 * NROW * 32 / clk_stage;     // int8 MACs per clk per SM
  printf("{\"stacked\": %d, \"rows\": %d, \"slices\": %d, \"stages\": %d, \"grid\": %d, \"buf_MB\": %.0f, \"ms\": %.3f, \"clk_per_stage\": %.0f, "
         "\"int8_mac_per_clk_per_sm\": %.0f, \"dmma_clk_same_work\": %.0f, \"speedup_vs_dmma_peak\": %.2f, \"l2_TBps\": %.2f}\n",
         (int)STACK, NROW, S, NST, grid, bytes / 1048576.0, ms, clk_stage, mac_rate, dmma_clk, dmma_clk / clk_stage,
         (double)grid * nstage * ST_BYTES / (ms * 1e-3) / 1e12);
  cudaFree(dclk);
}

int main() {
  int8_t *buf; const size_t big = (size_t)96 << 20, small = (size_t)4 << 20;
  CK(cudaMalloc(&buf, big)); CK(cudaMemset(buf, 1, big));
  run<64, 7, 2>(buf, small, 1);       // one SM, tiny buffer: the tensor / shared-memory limit
  run<64, 7, 2>(buf, big, 148);       // whole chip, L2-sized buffer
  run<64, 6, 2>(buf, big, 148);
  run<64, 8, 2>(buf, big, 148);
  run<128, 4, 2>(buf, big, 148);      // 4 levels x 128 rows (first pass of a two-pass scheme)
  run<32, 7, 3>(buf, big, 148);
  run<64, 8, 2, true>(buf, small, 1);
  run<64, 8, 2, true>(buf, big, 148);
  run<64, 7, 2, true>(buf, big, 148);
  run<64, 8, 1, true>(buf, big, 148);
  return 0;
}
