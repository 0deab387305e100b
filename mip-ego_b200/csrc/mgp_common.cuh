// Shared device/host helpers for libmgp_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define MGP_TILE 128  // row / column block of every blocked algorithm (n is padded to it)

// ---- FP64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col).  SASS: DMMA.8x8x4 ----------
// lane t holds A[t/4][t%4], B[t%4][t/4], C[t/4][2*(t%4) + {0,1}].
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---- mbarrier + bulk async copy (TMA engine; SASS: UBLKCP / SYNCS) --------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe of a phase
__device__ __forceinline__ bool mbar_test_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk copy, completion signalled on an mbarrier (bytes % 16 == 0)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- Ampere-style async copy (LDGSTS) for strided tiles -------------------------------------
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src_gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// ---- scalar math shared by the epilogues ------------------------------------------------------
#define MGP_SQRT3 1.7320508075688772
#define MGP_INV_SQRT2 0.70710678118654752440
#define MGP_INV_SQRT2PI 0.39894228040143267794

// standard normal cdf / pdf: scipy.stats.norm.cdf == ndtr (InfillCriteria.py:132)
__device__ __forceinline__ double norm_cdf(double x) { return 0.5 * erfc(-x * MGP_INV_SQRT2); }
__device__ __forceinline__ double norm_pdf(double x) { return exp(-0.5 * x * x) * MGP_INV_SQRT2PI; }

// correlation as a function of S = sum_i theta_i (x_i - x'_i)^2  (kernel.py:147-185, 277-317);
// CORR == 2 (absolute exponential, kernel.py:235-274): S = sum_i theta_i |x_i - x'_i|, r = exp(-S)
template <int CORR>
__device__ __forceinline__ double corr_from_S(double S) {
  S = fmax(S, 0.0);
  if (CORR != 1) return exp(-S);
  double s = MGP_SQRT3 * sqrt(S);
  return (1.0 + s) * exp(-s);
}

// ---- exp(-S), S >= 0, with a 64-entry table of 2^(j/64) ----------------------------------------
// -S = (64 k + j) ln2/64 + t with |t| <= ln2/128, exp(-S) = 2^k 2^(j/64) e^t and e^t - 1 by a degree-5
// polynomial (truncation t^6/720 < 3.5e-17): 10 FP64-pipe instructions (8 DFMA, DADD, DMUL) instead
// of ~22 for exp(); the range flush is an integer compare on the high word.  Measured relative
// error against long-double expl: tools/exp_check.cu (the table entries are correctly rounded).
// The K1 kernels are bound by the FP64 pipe, which DFMA shares with DMMA
// (profiles/r01_fp64_pipe_mix.json), so every instruction saved there is machine time.  `tab`
// points to the table in SHARED memory, stored as 64 entries x 16 replicas (entry j, replica
// lane & 15 at [16 j + (lane & 15)]) so that the data-dependent look-up of a half-warp never hits
// one bank twice (exp_tab_fill).
__constant__ double c_exp2_tab[64] = {
    1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284, 1.0442737824274138,
    1.0556451783605572, 1.0671404006768237, 1.0787607977571199, 1.0905077326652577, 1.102382583307841,
    1.1143867425958924, 1.1265216186082418, 1.1387886347566916, 1.1511892299529827, 1.1637248587775775,
    1.1763969916502812, 1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687,
    1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783, 1.2968395546510096,
    1.3109612115247644, 1.3252366431597413, 1.339667524053303, 1.3542555469368927, 1.3690024229745905,
    1.383909881963832, 1.3989796725383112, 1.4142135623730951, 1.42961333839197, 1.4451808069770467,
    1.460917794180647, 1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384,
    1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267, 1.6104903319492543,
    1.6280274218573478, 1.645755478153965, 1.6636765803267364, 1.681792830507429, 1.7001063537185235,
    1.718619298122478, 1.7373338352737062, 1.7562521603732995, 1.7753764925265212, 1.7947090750031072,
    1.8142521755003989, 1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656,
    1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951};

__device__ __forceinline__ double exp_neg_tab(double S, const double *tab) {
  const double x = -S;
  const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52: rounds to an integer
  double kf = fma(x, 92.33248261689366, MAGIC);            // 64 / ln 2
  const int N = __double2loint(kf);                        // 64 k + j (two's complement)
  kf -= MAGIC;
  double t = fma(kf, -0.01083042469326756, x);             // ln2/64, high part (32 significant bits)
  t = fma(kf, -2.9815858269852933e-12, t);                 // low part
  double p = fma(t, 1.0 / 120.0, 1.0 / 24.0);
  p = fma(p, t, 1.0 / 6.0);
  p = fma(p, t, 0.5);
  p = fma(p, t, 1.0);
  p *= t;
  const double tj = tab[((N & 63) << 4) | (threadIdx.x & 15)];   // replica = lane & 15: no bank conflicts
  const double y = fma(tj, p, tj);                         // in [1, 2)
  const double r = __hiloint2double(__double2hiint(y) + ((N >> 6) << 20), __double2loint(y));
  // below the normal range (S > 708, high word of 708.0 = 0x40862000; also NaN): flush.  An integer
  // compare on the high word keeps the test off the FP64 pipe.
  return __double2hiint(S) > 0x40862000 ? 0.0 : r;
}
#define MGP_EXP_TAB_DOUBLES 1024
__device__ __forceinline__ void exp_tab_fill(double *tab, int tid, int nthreads) {
  for (int i = tid; i < MGP_EXP_TAB_DOUBLES; i += nthreads) tab[i] = c_exp2_tab[i >> 4];
}
template <int CORR>
__device__ __forceinline__ double corr_from_S_tab(double S, const double *tab) {
  // S from the distance GEMM can come out a few ulp below zero for a candidate that sits on a
  // training point: exp_neg_tab then returns 1 + O(1e-15), which needs no clamp; the square root
  // of the Matern branch does
  if (CORR != 1) return exp_neg_tab(S, tab);
  const double s = MGP_SQRT3 * sqrt(fmax(S, 0.0));
  return (1.0 + s) * exp_neg_tab(s, tab);
}

// Function attributes (dynamic shared-memory opt-in) are per device: `first_use_on_device(flags)`
// is true once per (call site, current device).
#define MGP_MAX_DEVICES 64
static inline bool first_use_on_device(bool (&flags)[MGP_MAX_DEVICES]) {
  int dev = 0;
  cudaGetDevice(&dev);
  dev = dev < 0 ? 0 : (dev >= MGP_MAX_DEVICES ? MGP_MAX_DEVICES - 1 : dev);
  if (flags[dev]) return false;
  flags[dev] = true;
  return true;
}

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
