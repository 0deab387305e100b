// Posterior + acquisition kernels (sm_100a).
//
// Packed "slab" layout shared by the operands of K2c.  One slab = 16 tiles of 8x8 doubles
// (1024 doubles, 8 KB); a tile stores, for lane t of a warp, the two doubles that lane needs for
// two consecutive DMMA k-steps, so one conflict-free LDS.128 per lane feeds two MMAs:
//   Mp (A operand, L^-1 or R^-1): slab(I, j8) holds rows 128I..128I+127, columns 8j8..8j8+7;
//       tile i8 is the row-major 8x8 block   M[128I + 8 i8 + ri][8 j8 + kj]  at  ri*8 + kj.
//       Triangular packing keeps only j8 < 16(I+1): slab index 8 I (I+1) + j8.
//   rP (B operand, r = corr(X*, X)^T): slab(Cblk, j8) holds training rows 8j8..8j8+7 for the
//       128 candidates of block Cblk; tile c8 stores r[8 j8 + jj][128 Cblk + 8 c8 + cc] at cc*8 + jj.
// With this packing the MMA k index kk of step s maps to training index 8 j8 + 2 kk + s for both
// operands.  A pipeline stage (4 slabs of each operand = 32 k) is one contiguous 32 KB bulk copy
// per operand (cp.async.bulk -> UBLKCP) completing on an mbarrier.
#include "posterior.cuh"
#include "mgp_common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// per-fit operand preparation
// ---------------------------------------------------------------------------------------------
// Xsc: the training coordinates scaled by c_i = sqrt(theta_i) (theta_i for the absolute-exponential
// kernel), row stride xstride >= dpad, zero padded: the operand of the gradient kernel K5.
__global__ void k_prep_model(const double *Xc, const double *theta, int npad, int dpad, double *Bop,
                             double *an, double *Xsc, int xstride, int corr) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= npad) return;
  double s = 0.0;
  for (int i = 0; i < dpad; i++) {
    const double x = Xc[(int64_t)j * dpad + i], t = theta[i];
    Bop[(int64_t)j * dpad + i] = -2.0 * t * x;
    s = fma(t * x, x, s);
    Xsc[(int64_t)j * xstride + i] = (corr == MGP_CORR_ABSEXP ? t : sqrt(t)) * x;
  }
  for (int i = dpad; i < xstride; i++) Xsc[(int64_t)j * xstride + i] = 0.0;
  an[j] = s;
}

// ---------------------------------------------------------------------------------------------
// K1: cross-correlation tiles.  S[c][j] = bn_c + an_j + sum_i xs_ci * (-2 theta_i x_ji) as a
// DMMA GEMM over the (padded) feature dimension, then r = corr(S)  (kernel.py:147-185,277-317;
// the reference forms |x - x'| component-wise, gpr.py:453-455).  Also accumulates
// r . gamma (posterior mean, gpr.py:459) and r . a with a = L^-T Ft (Ft^T rt, gpr.py:467).
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int rgen_stride(int dpad) { return ((dpad + 11) / 16) * 16 + 4; }

__host__ __device__ constexpr int trend_tiles(int KS) { return (4 * KS + 1 + 7) / 8; }   // 8-col tiles for p <= 4 KS + 1
__host__ __device__ constexpr int trend_stride(int P8) { return ((8 * P8 + 15) / 16) * 16 + 2; }  // == 2 (mod 16)

// Bhat^T r on DMMA: the r values a lane holds after K1's distance GEMM (C fragment: candidate
// lr, training rows 2 lk and 2 lk + 1 of the tile) ARE the A fragments of two MMA k-steps whose
// k index is mapped to training row 2 lk + s; the B fragment is Bhat[2 lk + s][8 c8 + lr].
template <int P8>
__device__ __forceinline__ void trend_mma(double (&tacc)[P8][2], double r0, double r1, const double *sBh,
                                          int STB, int t8, int lr, int lk) {
#pragma unroll
  for (int c8 = 0; c8 < P8; c8++) {
    const double b0 = sBh[(8 * t8 + 2 * lk) * STB + 8 * c8 + lr];
    const double b1 = sBh[(8 * t8 + 2 * lk + 1) * STB + 8 * c8 + lr];
    dmma884(tacc[c8][0], tacc[c8][1], r0, b0);
    dmma884(tacc[c8][0], tacc[c8][1], r1, b1);
  }
}

// I8: instead of the packed FP64 r the kernel writes the eight 7-bit slices of r in the shared-memory
// image of K2c-i8's A operand (ozaki.cu; stage = 64 training rows, slice = 8 KB, byte
// (kk / 16) * 2048 + (c / 8) * 128 + (c % 8) * 16 + kk % 16): the lanes of a quad hold 2 x 2 adjacent
// training rows of one candidate for an even / odd pair of 8-row groups, i.e. the 16 bytes of one
// core-matrix row between them; two shuffles and a byte permute per slice give every lane one 32-bit
// word of it, and a warp stores one full 128-byte core matrix per slice.  Saves the FP64 round trip
// of r through HBM and the separate slicing pass (k_oz_slice_r).
__device__ __forceinline__ void oz_fixed56(double r, uint32_t &lo, uint32_t &hi) {
  const long long v = __double2ll_rn(r * 72057594037927936.0);   // rint(r 2^56), r in [0, 1]
  lo = (uint32_t)v;
  hi = (uint32_t)((unsigned long long)v >> 32);
}
template <int P>
__device__ __forceinline__ uint32_t oz_digit56(uint32_t lo, uint32_t hi) {   // digit P of 8: bits [7 (7 - P), 7 (8 - P))
  constexpr int sh = 7 * (7 - P);
  if (P == 0) return hi >> (sh - 32);                                        // keeps the overflow (r = 1 -> 128)
  return (sh >= 32 ? (hi >> (sh - 32)) : (sh == 0 ? lo : __funnelshift_r(lo, hi, sh))) & 127u;
}

template <int CORR, int KS, bool LIN, bool I8 = false>
__global__ void __launch_bounds__(256) k_rgen(RgenParams p) {
  // Persistent: work item = (candidate block cb of 128, training block jb of 128).  The operands
  // of the training block (Bop rows, an, gamma, a) are staged once per item in shared memory.
  constexpr int DPAD = 4 * KS, SB = rgen_stride(DPAD);   // SB % 16 == 4: conflict-free B fragments
  extern __shared__ __align__(16) double sm_rg[];
  constexpr int P8 = LIN ? trend_tiles(KS) : 1, STB = trend_stride(P8), PT = 8 * P8;
  double *sB = sm_rg, *sV = sm_rg + 128 * SB, *sTab = sV + 3 * 128, *sBh = sTab + MGP_EXP_TAB_DOUBLES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, lr = lane >> 2, lk = lane & 3;
  const int nItems = p.nCblk * p.JS;
  exp_tab_fill(sTab, tid, 256);
  for (int item = blockIdx.x; item < nItems; item += gridDim.x) {
    const int jb = item / p.nCblk, cb = item % p.nCblk;
    __syncthreads();
    for (int idx = tid; idx < 128 * DPAD; idx += 256) {
      const int r = idx / DPAD, i = idx % DPAD;
      sB[r * SB + i] = __ldg(p.Bop + (int64_t)(jb * 128 + r) * DPAD + i);
    }
    if (tid < 128) {
      sV[tid] = __ldg(p.an + jb * 128 + tid);
      sV[128 + tid] = __ldg(p.gamma + jb * 128 + tid);
      sV[256 + tid] = LIN ? 0.0 : __ldg(p.avec + jb * 128 + tid);
    }
    if (LIN) {
      for (int idx = tid; idx < 128 * PT; idx += 256) {
        const int r = idx / PT, c = idx % PT;
        sBh[r * STB + c] = __ldg(p.Bhat + (int64_t)(jb * 128 + r) * MGP_TP + c);
      }
    }
    double tacc[2][P8][2];
#pragma unroll
    for (int u = 0; u < 2; u++)
#pragma unroll
      for (int c8 = 0; c8 < P8; c8++) tacc[u][c8][0] = tacc[u][c8][1] = 0.0;
    double xa[2][KS], bn[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int64_t c = (int64_t)cb * 128 + (2 * warp + u) * 8 + lr;
      const bool valid = c < p.mc;
      double part = 0.0;
#pragma unroll
      for (int s = 0; s < KS; s++) {
        const int i = 4 * s + lk;
        double v = 0.0;
        if (valid && i < p.d) v = p.Xs[c * p.d + i] - p.center[i];
        xa[u][s] = v;
        part = fma(p.theta[i] * v, v, part);
      }
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      bn[u] = part;
    }
    __syncthreads();
    double macc[2] = {0.0, 0.0}, facc[2] = {0.0, 0.0};
    // output tile (cb, j8 = 16 jb + t8, c8 = 2 warp + u): one 64-bit base per item, 32-bit offsets inside
    double *const dst0 = I8 ? nullptr : p.rP + (((int64_t)cb * p.NJ8 + jb * 16) * 16 + 2 * warp) * 64 + lane * 2;
    // I8: stage (cb, ks = 2 jb + tp / 4); inside a slice: (tp % 4) * 2048 + (2 warp + u) * 128 + lr * 16 + 4 lk
    uint8_t *const i8dst = I8 ? p.RI8 + ((int64_t)cb * p.nks + 2 * jb) * 65536 + warp * 256 + lr * 16 + lk * 4 : nullptr;
#pragma unroll 1
    for (int tp = 0; tp < 8; tp++) {
      uint32_t flo[2][2][2], fhi[2][2][2];      // [e][u][value]: rint(r 2^56)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int t8 = 2 * tp + e;
        const int jrow = 8 * t8 + lr, jcol = 8 * t8 + 2 * lk;
        double bf[KS];
#pragma unroll
        for (int s = 0; s < KS; s++) bf[s] = sB[jrow * SB + 4 * s + lk];
        const double2 an2 = *reinterpret_cast<const double2 *>(sV + jcol);
        const double2 g2 = *reinterpret_cast<const double2 *>(sV + 128 + jcol);
        const double2 a2 = *reinterpret_cast<const double2 *>(sV + 256 + jcol);
#pragma unroll
        for (int u = 0; u < 2; u++) {
          double c0 = bn[u] + an2.x, c1 = bn[u] + an2.y;
#pragma unroll
          for (int s = 0; s < KS; s++) dmma884(c0, c1, xa[u][s], bf[s]);
          const double r0 = corr_from_S_tab<CORR>(c0, sTab), r1 = corr_from_S_tab<CORR>(c1, sTab);
          if (I8) {
            oz_fixed56(r0, flo[e][u][0], fhi[e][u][0]);
            oz_fixed56(r1, flo[e][u][1], fhi[e][u][1]);
          } else {
            *reinterpret_cast<double2 *>(dst0 + (t8 * 16 + u) * 64) = make_double2(r0, r1);
          }
          macc[u] = fma(g2.x, r0, macc[u]);
          macc[u] = fma(g2.y, r1, macc[u]);
          if (LIN) {
            trend_mma<P8>(tacc[u], r0, r1, sBh, STB, t8, lr, lk);
          } else {
            facc[u] = fma(a2.x, r0, facc[u]);
            facc[u] = fma(a2.y, r1, facc[u]);
          }
        }
      }
      if (I8) {
        // word j of a core-matrix row = training rows 4 j .. 4 j + 3 of the 16: e = j / 2, source lanes lk' = 2 (j % 2), + 1
        const int srcA = (lane & ~3) | ((lk & 1) << 1);
        const uint32_t sel = (lk & 2) ? 0x7632u : 0x5410u;
        uint8_t *const d0 = i8dst + (int64_t)(tp >> 2) * 65536 + (tp & 3) * 2048;
#pragma unroll
        for (int u = 0; u < 2; u++) {
#define OZ_EMIT(P)                                                                                               \
          {                                                                                                        \
            const uint32_t w = oz_digit56<P>(flo[0][u][0], fhi[0][u][0]) | (oz_digit56<P>(flo[0][u][1], fhi[0][u][1]) << 8) | \
                               (oz_digit56<P>(flo[1][u][0], fhi[1][u][0]) << 16) | (oz_digit56<P>(flo[1][u][1], fhi[1][u][1]) << 24); \
            const uint32_t xa_ = __shfl_sync(0xffffffffu, w, srcA), xb_ = __shfl_sync(0xffffffffu, w, srcA + 1);     \
            *reinterpret_cast<uint32_t *>(d0 + P * 8192 + u * 128) = __byte_perm(xa_, xb_, sel);                     \
          }
          OZ_EMIT(0) OZ_EMIT(1) OZ_EMIT(2) OZ_EMIT(3) OZ_EMIT(4) OZ_EMIT(5) OZ_EMIT(6) OZ_EMIT(7)
#undef OZ_EMIT
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
      macc[u] += __shfl_xor_sync(0xffffffffu, macc[u], 1);
      macc[u] += __shfl_xor_sync(0xffffffffu, macc[u], 2);
      facc[u] += __shfl_xor_sync(0xffffffffu, facc[u], 1);
      facc[u] += __shfl_xor_sync(0xffffffffu, facc[u], 2);
      const int64_t cand = (int64_t)cb * 128 + (2 * warp + u) * 8 + lr;
      if (lk == 0) {
        const int64_t o = (int64_t)jb * p.ldp + cand;
        p.meanpart[o] = macc[u];
        if (!LIN) p.ftpart[o] = facc[u];
      }
      if (LIN) {
#pragma unroll
        for (int c8 = 0; c8 < P8; c8++) {
          const int col = 8 * c8 + 2 * lk;
          p.tpart[((int64_t)jb * PT + col) * p.ldp + cand] = tacc[u][c8][0];
          p.tpart[((int64_t)jb * PT + col + 1) * p.ldp + cand] = tacc[u][c8][1];
        }
      }
    }
  }
}

// K1 for the absolute-exponential correlation r = exp(-sum_i theta_i |x_i - x'_i|)
// (kernel.py:235-274): the L1 distance has no GEMM form, so S is accumulated feature by feature on
// the FP64 pipe from shared-memory copies of the training block (transposed) and the candidate
// block.  Same work items, outputs and packed layout as k_rgen.
template <int DPAD>
__global__ void __launch_bounds__(256) k_rgen_l1(RgenParams p) {
  constexpr int XS = 132, CS = DPAD + 1;
  extern __shared__ __align__(16) double sm_rg[];
  double *sX = sm_rg;                 // [DPAD][XS]   training block, feature-major
  double *sC = sX + DPAD * XS;        // [128][CS]    candidate block
  double *sTh = sC + 128 * CS;        // [DPAD]
  double *sV = sTh + DPAD;            // gamma | avec
  double *sTab = sV + 2 * 128;        // 2^(j/64), 16 replicas
  exp_tab_fill(sTab, threadIdx.x, 256);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, lr = lane >> 2, lk = lane & 3;
  const int nItems = p.nCblk * p.JS;
  int last_cb = -1;
  for (int item = blockIdx.x; item < nItems; item += gridDim.x) {
    const int jb = item / p.nCblk, cb = item % p.nCblk;
    __syncthreads();
    for (int idx = tid; idx < 128 * DPAD; idx += 256) {
      const int r = idx / DPAD, i = idx % DPAD;
      sX[i * XS + r] = __ldg(p.Xc + (int64_t)(jb * 128 + r) * DPAD + i);
    }
    if (cb != last_cb) {
      for (int idx = tid; idx < 128 * DPAD; idx += 256) {
        const int r = idx / DPAD, i = idx % DPAD;
        const int64_t c = (int64_t)cb * 128 + r;
        sC[r * CS + i] = (c < p.mc && i < p.d) ? p.Xs[c * p.d + i] - p.center[i] : 0.0;
      }
      last_cb = cb;
    }
    for (int i = tid; i < DPAD; i += 256) sTh[i] = p.theta[i];
    if (tid < 128) {
      sV[tid] = __ldg(p.gamma + jb * 128 + tid);
      sV[128 + tid] = __ldg(p.avec + jb * 128 + tid);
    }
    __syncthreads();
    double macc[2] = {0.0, 0.0}, facc[2] = {0.0, 0.0};
    const double *c0p = sC + ((2 * warp) * 8 + lr) * CS, *c1p = sC + ((2 * warp + 1) * 8 + lr) * CS;
#pragma unroll 1
    for (int t8 = 0; t8 < 16; t8++) {
      const int jcol = 8 * t8 + 2 * lk;
      double s00 = 0.0, s01 = 0.0, s10 = 0.0, s11 = 0.0;
#pragma unroll 4
      for (int i = 0; i < DPAD; i++) {
        const double2 xj = *reinterpret_cast<const double2 *>(sX + i * XS + jcol);
        const double th = sTh[i], a = c0p[i], b = c1p[i];
        s00 = fma(th, fabs(a - xj.x), s00);
        s01 = fma(th, fabs(a - xj.y), s01);
        s10 = fma(th, fabs(b - xj.x), s10);
        s11 = fma(th, fabs(b - xj.y), s11);
      }
      const double2 g2 = *reinterpret_cast<const double2 *>(sV + jcol);
      const double2 a2 = *reinterpret_cast<const double2 *>(sV + 128 + jcol);
      const int j8 = jb * 16 + t8;
      const double r00 = exp_neg_tab(s00, sTab), r01 = exp_neg_tab(s01, sTab), r10 = exp_neg_tab(s10, sTab),
                   r11 = exp_neg_tab(s11, sTab);
      *reinterpret_cast<double2 *>(p.rP + (((int64_t)cb * p.NJ8 + j8) * 16 + 2 * warp) * 64 + lane * 2) =
          make_double2(r00, r01);
      *reinterpret_cast<double2 *>(p.rP + (((int64_t)cb * p.NJ8 + j8) * 16 + 2 * warp + 1) * 64 + lane * 2) =
          make_double2(r10, r11);
      macc[0] = fma(g2.x, r00, macc[0]); macc[0] = fma(g2.y, r01, macc[0]);
      macc[1] = fma(g2.x, r10, macc[1]); macc[1] = fma(g2.y, r11, macc[1]);
      facc[0] = fma(a2.x, r00, facc[0]); facc[0] = fma(a2.y, r01, facc[0]);
      facc[1] = fma(a2.x, r10, facc[1]); facc[1] = fma(a2.y, r11, facc[1]);
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
      macc[u] += __shfl_xor_sync(0xffffffffu, macc[u], 1);
      macc[u] += __shfl_xor_sync(0xffffffffu, macc[u], 2);
      facc[u] += __shfl_xor_sync(0xffffffffu, facc[u], 1);
      facc[u] += __shfl_xor_sync(0xffffffffu, facc[u], 2);
      if (lk == 0) {
        const int64_t o = (int64_t)jb * p.ldp + (int64_t)cb * 128 + (2 * warp + u) * 8 + lr;
        p.meanpart[o] = macc[u];
        p.ftpart[o] = facc[u];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K2c: T = L^-1 r on DMMA with the column sums of squares fused (|L^-1 r|^2, gpr.py:463,472);
// T never leaves registers.  FULL variant: W = R^-1 r written back in the rP layout (gradient
// path, replaces the d triangular solves of gpr.py:543).
// Work item = (row block I, candidate block cb); items are dealt round-robin to a persistent
// grid in order of decreasing cost.  One elected thread drives the bulk-copy pipeline.
// ---------------------------------------------------------------------------------------------
constexpr int TS = 3;             // pipeline stages
constexpr int STAGE_DBL = 4096;   // doubles per operand per stage (4 slabs, 32 KB)
constexpr int TRMM_SMEM = 2 * TS * STAGE_DBL * 8 + 2 * TS * 8 + 2 * 128 * 8;

// One 8-k slab of the 128x128 tile for one consumer warp: row tiles RT0..7 (interleaved: tile
// i8 = 2 rt + wr) times 4 column tiles.  RT0 is a template parameter so that skipping the zero
// tiles of the diagonal block is a real (warp-uniform) branch: predicated-off DMMAs would still
// pay their static issue stalls.
template <int RT0, int RT1 = 8>
__device__ __forceinline__ void slab_mma(double (&acc)[8][4][2], const double *as_slab,
                                         const double *bs_slab, int wr, int wc, int lane) {
  double2 b[4], a[8];
#pragma unroll
  for (int ct = 0; ct < 4; ct++)
    b[ct] = *reinterpret_cast<const double2 *>(bs_slab + (wc * 4 + ct) * 64 + lane * 2);
#pragma unroll
  for (int rt = RT0; rt < RT1; rt++)
    a[rt] = *reinterpret_cast<const double2 *>(as_slab + (2 * rt + wr) * 64 + lane * 2);
#pragma unroll
  for (int rt = RT0; rt < RT1; rt++)
#pragma unroll
    for (int ct = 0; ct < 4; ct++) dmma884(acc[rt][ct][0], acc[rt][ct][1], a[rt].x, b[ct].x);
#pragma unroll
  for (int rt = RT0; rt < RT1; rt++)
#pragma unroll
    for (int ct = 0; ct < 4; ct++) dmma884(acc[rt][ct][0], acc[rt][ct][1], a[rt].y, b[ct].y);
}

// A whole 32-k stage (four slabs, straight-line code) for a warp whose last row tiles are padding: at
// n = 500 the odd-parity warps of the last row block own 7 valid row tiles, and that row block has the
// longest items.
template <int RT1>
__device__ __forceinline__ void stage_rt1(double (&acc)[8][4][2], const double *as, const double *bs, int wr,
                                          int wc, int lane) {
#pragma unroll
  for (int js = 0; js < 4; js++) slab_mma<0, RT1>(acc, as + js * 1024, bs + js * 1024, wr, wc, lane);
}
__device__ __forceinline__ void stage_upto(int rt1, double (&acc)[8][4][2], const double *as, const double *bs,
                                           int wr, int wc, int lane) {
  switch (rt1) {
    case 1: stage_rt1<1>(acc, as, bs, wr, wc, lane); break;
    case 2: stage_rt1<2>(acc, as, bs, wr, wc, lane); break;
    case 3: stage_rt1<3>(acc, as, bs, wr, wc, lane); break;
    case 4: stage_rt1<4>(acc, as, bs, wr, wc, lane); break;
    case 5: stage_rt1<5>(acc, as, bs, wr, wc, lane); break;
    case 6: stage_rt1<6>(acc, as, bs, wr, wc, lane); break;
    case 7: stage_rt1<7>(acc, as, bs, wr, wc, lane); break;
    default: break;
  }
}

// One 32-k stage inside the diagonal 128x128 block for consumer row parity WR: slab js of stage DL skips
// the row tiles above the diagonal, rt0 = (4 DL + js - WR + 1) / 2.  With DL and WR as template parameters
// the four slabs are straight-line code, so the fragment loads of a slab are issued under the DMMAs of the
// one before (the loop over js with a switch per slab left the tensor pipe waiting for the loads at the head
// of every slab: 40 % of the stages at C2 lie in diagonal blocks).
template <int DL, int WR>
__device__ __forceinline__ void diag_stage(double (&acc)[8][4][2], const double *as, const double *bs, int wc,
                                           int lane) {
#pragma unroll
  for (int js = 0; js < 4; js++) {
    constexpr int base = 4 * DL - WR + 1;
    const int rt0c = (base + js) >> 1;   // compile-time after unrolling
    if (rt0c == 0) slab_mma<0>(acc, as + js * 1024, bs + js * 1024, WR, wc, lane);
    else if (rt0c == 1) slab_mma<1>(acc, as + js * 1024, bs + js * 1024, WR, wc, lane);
    else if (rt0c == 2) slab_mma<2>(acc, as + js * 1024, bs + js * 1024, WR, wc, lane);
    else if (rt0c == 3) slab_mma<3>(acc, as + js * 1024, bs + js * 1024, WR, wc, lane);
    else if (rt0c == 4) slab_mma<4>(acc, as + js * 1024, bs + js * 1024, WR, wc, lane);
    else if (rt0c == 5) slab_mma<5>(acc, as + js * 1024, bs + js * 1024, WR, wc, lane);
    else if (rt0c == 6) slab_mma<6>(acc, as + js * 1024, bs + js * 1024, WR, wc, lane);
    else if (rt0c == 7) slab_mma<7>(acc, as + js * 1024, bs + js * 1024, WR, wc, lane);
  }
}

// Item order: candidate blocks in groups of G = gridDim.x, inside a group the row blocks in decreasing
// cost, candidate block innermost.  Dealt round-robin, CTA w of a full group therefore walks ALL row
// blocks of ONE candidate block (g G + w) one after the other and finds its r tile, which it read from
// HBM for the longest row block, in L2 for the others: at C2 the DRAM reads of a launch fall from
// (1 + 2 + 3 + 4) / 4 = 2.5 x the r chunk to about 1 x, and every CTA still gets one item per cost class
// (the order "all candidate blocks of row block NI - 1, then NI - 2, ..." it replaces was equally balanced
// but re-read every r tile NI times from HBM).  The last group may be ragged.
__device__ __forceinline__ void trmm_item(int item, int NI, int nCblk, int G, int &I, int &cb) {
  const int per = G * NI;
  const int grp = item / per, rem = item - grp * per;
  const int gsz = min(G, nCblk - grp * G);
  I = NI - 1 - rem / gsz;
  cb = grp * G + rem % gsz;
}

template <bool FULL>
__global__ void __launch_bounds__(384, 1) k_trmm(TrmmParams p) {
  // Warp-specialised: warpgroups 0-1 (warps 0-7) are MMA consumers, warpgroup 2 is the producer.
  // Registers are partitioned per SM sub-partition (16K each), so the three warps that share a
  // sub-partition start at 168 registers; the producer group gives most of its share back
  // (setmaxnreg.dec) and the consumer groups grow to 232 (setmaxnreg.inc).
  extern __shared__ __align__(128) unsigned char smraw[];
  double *As = reinterpret_cast<double *>(smraw);
  double *Bs = As + TS * STAGE_DBL;
  uint64_t *full = reinterpret_cast<uint64_t *>(Bs + TS * STAGE_DBL);
  uint64_t *empty = full + TS;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nItems = p.NI * p.nCblk;

  if (tid == 0) {
    for (int s = 0; s < TS; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 8);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
    // ---- producer: one lane streams the operand slabs through the TS-stage ring (TMA bulk copies)
    if (warp == 8 && lane == 0) {
      int64_t x = 0;
      for (int item = blockIdx.x; item < nItems; item += gridDim.x) {
        int I, cb;
        trmm_item(item, p.NI, p.nCblk, (int)gridDim.x, I, cb);
        const int nkc = FULL ? p.NJ8 / 4 : 4 * (I + 1);
        for (int kc = 0; kc < nkc; kc++, x++) {
          const int stage = (int)(x % TS);
          if (x >= TS) mbar_wait(&empty[stage], (uint32_t)(((x / TS) - 1) & 1));
          const int64_t slabA =
              FULL ? ((int64_t)I * p.NJ8 + 4 * kc) : ((int64_t)8 * I * (I + 1) + 4 * kc);
          const double *srcA = p.Mp + slabA * 1024;
          const double *srcB = p.rP + ((int64_t)cb * p.NJ8 + 4 * kc) * 1024;
          double *dstA = As + stage * STAGE_DBL;
          const int dl = FULL ? -1 : kc - 4 * I;
          if (dl < 0) {
            mbar_expect_tx(&full[stage], 2 * STAGE_DBL * 8);
            bulk_g2s(dstA, srcA, STAGE_DBL * 8, &full[stage]);
          } else {
            // diagonal block: row tiles i8 < jl of slab jl are zero and never read -- skip them
            uint32_t bytes = STAGE_DBL * 8;
            for (int js = 0; js < 4; js++) bytes += (uint32_t)(16 - (4 * dl + js)) * 512u;
            mbar_expect_tx(&full[stage], bytes);
            for (int js = 0; js < 4; js++) {
              const int jl = 4 * dl + js;
              bulk_g2s(dstA + js * 1024 + jl * 64, srcA + js * 1024 + jl * 64,
                       (uint32_t)(16 - jl) * 512u, &full[stage]);
            }
          }
          bulk_g2s(Bs + stage * STAGE_DBL, srcB, STAGE_DBL * 8, &full[stage]);
        }
      }
    }
    return;
  }
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");

  // ---- 8 consumer warps: 2 (row parity) x 4 (column groups of 32) over a 128 x 128 tile ----
  const int wr = warp >> 2, wc = warp & 3, lr = lane >> 2, lk = lane & 3;
  int64_t x = 0;
  for (int item = blockIdx.x; item < nItems; item += gridDim.x) {
    int I, cb;
    trmm_item(item, p.NI, p.nCblk, (int)gridDim.x, I, cb);
    const int nkc = FULL ? p.NJ8 / 4 : 4 * (I + 1);
    double acc[8][4][2];
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int c = 0; c < 4; c++) acc[r][c][0] = acc[r][c][1] = 0.0;

    for (int kc = 0; kc < nkc; kc++, x++) {
      const int stage = (int)(x % TS);
      mbar_wait(&full[stage], (uint32_t)((x / TS) & 1));
      const double *as = As + stage * STAGE_DBL, *bs = Bs + stage * STAGE_DBL;
      const int dl = FULL ? -1 : kc - 4 * I;  // >= 0 inside the diagonal 128x128 block
      if (dl < 0) {
        // valid 8-row tiles of this warp in the row block (interleaved: tile i8 = 2 rt + wr)
        const int rt1 = (!FULL && I == p.NI - 1) ? min(8, max(0, (p.last_tiles - wr + 1) >> 1)) : 8;
        if (rt1 == 8) {
#pragma unroll
          for (int js = 0; js < 4; js++) slab_mma<0>(acc, as + js * 1024, bs + js * 1024, wr, wc, lane);
        } else {
          stage_upto(rt1, acc, as, bs, wr, wc, lane);
        }
      } else {
        // diagonal block: tiles above the diagonal (i8 < jl) are zero.  Row tiles are interleaved
        // between the two row-warps (i8 = 2 rt + wr), so both warps of every SM sub-partition
        // keep the same amount of work; rt0 = smallest rt with 2 rt + wr >= jl.
        switch (2 * dl + wr) {
          case 0: diag_stage<0, 0>(acc, as, bs, wc, lane); break;
          case 1: diag_stage<0, 1>(acc, as, bs, wc, lane); break;
          case 2: diag_stage<1, 0>(acc, as, bs, wc, lane); break;
          case 3: diag_stage<1, 1>(acc, as, bs, wc, lane); break;
          case 4: diag_stage<2, 0>(acc, as, bs, wc, lane); break;
          case 5: diag_stage<2, 1>(acc, as, bs, wc, lane); break;
          case 6: diag_stage<3, 0>(acc, as, bs, wc, lane); break;
          default: diag_stage<3, 1>(acc, as, bs, wc, lane); break;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
    }

    if (!FULL) {
      double qs[4][2];
#pragma unroll
      for (int ct = 0; ct < 4; ct++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          double s = 0.0;
#pragma unroll
          for (int rt = 0; rt < 8; rt++) s = fma(acc[rt][ct][e], acc[rt][ct][e], s);
          s += __shfl_xor_sync(0xffffffffu, s, 4);
          s += __shfl_xor_sync(0xffffffffu, s, 8);
          s += __shfl_xor_sync(0xffffffffu, s, 16);
          qs[ct][e] = s;
        }
      // Every warp writes the sums of ITS 64 rows (row parity wr) straight from registers: partial row
      // 2 I + wr of qpart, added in a fixed order by K3.  No barrier between the consumer warps at the end
      // of an item: a warp that is done moves on to the next item's stages while the other warp of its SM
      // sub-partition is still issuing DMMAs (the two consumer barriers per item of the earlier version,
      // which summed the two parities through shared memory, left the tensor pipe idle 8 % of the time).
      if (lr == 0) {
        double *qrow = p.qpart + (int64_t)(2 * I + wr) * p.ldq + (int64_t)cb * 128 + wc * 32 + 2 * lk;
#pragma unroll
        for (int ct = 0; ct < 4; ct++) *reinterpret_cast<double2 *>(qrow + ct * 8) = make_double2(qs[ct][0], qs[ct][1]);
      }
    } else {
      // W[j][c] for j in row block I: thread holds rows 8(2rt+wr)+lr, cols 8(wc*4+ct)+2lk+{0,1}.
      // rP-style layout wants, per (j8, c8) tile, element (cc, jj) at cc*8 + jj.
#pragma unroll
      for (int rt = 0; rt < 8; rt++)
#pragma unroll
        for (int ct = 0; ct < 4; ct++) {
          const int j8 = I * 16 + 2 * rt + wr, c8 = wc * 4 + ct;
          double *tile = p.wP + (((int64_t)cb * p.NJ8 + j8) * 16 + c8) * 64;
          tile[(2 * lk) * 8 + lr] = acc[rt][ct][0];
          tile[(2 * lk + 1) * 8 + lr] = acc[rt][ct][1];
        }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Small populations (<= 8 candidates per pass): the call pattern of the reference's own
// argmax_restart, one point per criterion call (optimizer/utils.py:38-42).  T = L^-1 r (W = R^-1 r)
// is then a matrix-vector product bound by reading the operand once; the 128-wide MMA tiles of
// K2c would waste 15/16 of their work on padding and keep only NI SMs busy.  A CTA owns 32 matrix
// rows (4 per warp), lanes stride over k; all candidates share one pass over the operand.
//   FULL == false: M = L^-1 (row-major, lower);  qpart[cta][c] = sum over the CTA's rows of T^2
//   FULL == true : M = R^-1 (row-major, symmetric, both triangles); W written in the packed layout
// r (and W) use the packed layout of K1 with candidate block 0, tile c8 = 0.
// ---------------------------------------------------------------------------------------------
template <bool FULL>
__global__ void __launch_bounds__(256) k_small_mv(const double *M, int npad, int n, const double *rP, int mc,
                                                  double *qpart, int64_t ldq, double *wP) {
  __shared__ double red[8][8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = blockIdx.x * 32 + warp * 4;
  double acc[4][8];
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int c = 0; c < 8; c++) acc[r][c] = 0.0;
  const int kmax = FULL ? n : min(n, row0 + 4);
  for (int k = lane; k < kmax; k += 32) {
    double rv[8];
    const double *rk = rP + (int64_t)(k >> 3) * 1024 + (k & 7);
#pragma unroll
    for (int c = 0; c < 8; c++) rv[c] = c < mc ? rk[c * 8] : 0.0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int row = row0 + r;
      const double m = (row < n && (FULL || k <= row)) ? M[(int64_t)row * npad + k] : 0.0;
#pragma unroll
      for (int c = 0; c < 8; c++) acc[r][c] = fma(m, rv[c], acc[r][c]);
    }
  }
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int c = 0; c < 8; c++) {
      double v = acc[r][c];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      acc[r][c] = v;
    }
  if (FULL) {
    if (lane < 8) {
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const int row = row0 + r;
        double v = 0.0;
#pragma unroll
        for (int c = 0; c < 8; c++) v = (lane == c) ? acc[r][c] : v;
        wP[(int64_t)(row >> 3) * 1024 + lane * 8 + (row & 7)] = v;
      }
    }
  } else {
    if (lane < 8) {
      double q = 0.0;
#pragma unroll
      for (int r = 0; r < 4; r++) {
        double v = 0.0;
#pragma unroll
        for (int c = 0; c < 8; c++) v = (lane == c) ? acc[r][c] : v;
        q = fma(v, v, q);
      }
      red[warp][lane] = q;
    }
    __syncthreads();
    if (tid < 8) {
      double q = 0.0;
      for (int w = 0; w < 8; w++) q += red[w][tid];   // fixed order
      qpart[(int64_t)blockIdx.x * ldq + tid] = q;
    }
  }
}

// upper triangle := lower triangle (row-major npad x npad), 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(1024) k_symmetrize(double *A, int npad) {
  __shared__ double t[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj > bi) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  t[ty][tx] = A[(int64_t)(bi * 32 + ty) * npad + bj * 32 + tx];
  __syncthreads();
  if (bi == bj) {
    if (tx > ty) A[(int64_t)(bi * 32 + ty) * npad + bj * 32 + tx] = t[tx][ty];
  } else {
    A[(int64_t)(bj * 32 + ty) * npad + bi * 32 + tx] = t[tx][ty];
  }
}

// ---------------------------------------------------------------------------------------------
// acquisition functions (InfillCriteria.py).  Returns the value and the coefficients A, B of
//   f_dx = A * yhat_dx + B * sd_dx      (yhat already sign-flipped when maximising)
// ---------------------------------------------------------------------------------------------
struct Acq {
  double val, A, B;
};
__device__ __forceinline__ Acq acq_eval(int kind, double par, double plugin, double yhat, double sd,
                                        double sigma2) {
  Acq o;
  o.val = 0.0; o.A = 0.0; o.B = 0.0;
  if (kind == MGP_ACQ_UCB) {  // InfillCriteria.py:91-111
    o.val = yhat + par * sd;
    o.A = 1.0;
    o.B = par;
  } else if (kind == MGP_ACQ_EI) {  // InfillCriteria.py:114-148
    if (sd / sqrt(sigma2) < 1e-6) return o;
    const double xcr_ = plugin - yhat, xcr = xcr_ / sd;
    const double cdf = norm_cdf(xcr), pdf = norm_pdf(xcr);
    o.val = xcr_ * cdf + sd * pdf;
    o.A = -cdf;
    o.B = pdf;
  } else if (kind == MGP_ACQ_PI) {  // EpsilonPI, InfillCriteria.py:167-187
    const double coef = yhat > 0.0 ? 1.0 - par : 1.0 + par;
    const double xcr = (plugin - coef * yhat) / sd;
    const double pdf = norm_pdf(xcr);
    o.val = norm_cdf(xcr);
    o.A = -coef * pdf / sd;
    o.B = -xcr * pdf / sd;
  } else {  // MGFI, InfillCriteria.py:213-252
    const double t = fmin(par, 22.36);
    if (fabs(sd) <= 1e-8) return o;  // np.isclose(sd, 0)
    const double bp = (plugin - (yhat - t * sd * sd)) / sd;
    const double expo = t * (plugin - yhat - 1.0) + t * t * sd * sd * 0.5;
    const double e = exp(expo), cdf = norm_cdf(bp), pdf = norm_pdf(bp);
    const double v = cdf * e;
    if (!isfinite(v) || !isfinite(e)) return o;  // overflow / invalid -> 0 (:223-236)
    o.val = v;
    o.A = e * (-pdf / sd - t * cdf);
    o.B = e * (pdf * (2.0 * t * sd - bp) / sd + cdf * t * t * sd);
  }
  return o;
}

// total order of the top-k selection: larger value first, ties -> smaller index first
__device__ __forceinline__ bool better(double v, int64_t i, double w, int64_t j) {
  return (v > w) || (v == w && i < j);
}

// K3: fixed-order reduction of the partial sums, MSE (gpr.py:465-474) and acquisition values.
// With blk_val != nullptr (top-1 requests) the block also reduces its 256 candidates to the best
// (value, global index) per parameter set, so that the selection kernel scans one entry per block
// instead of one per candidate.
__global__ void __launch_bounds__(256) k_epilogue(EpiParams p) {
  __shared__ double sbv[8];
  __shared__ int64_t sbi[8];
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = c < p.mc;
  if (!valid && !p.blk_val) return;
  double mean = 0.0, mse = 0.0;
  if (valid) {
    double q = 0.0;
    for (int I = 0; I < p.NI; I++) q += p.qpart[(int64_t)I * p.ld + c];
    double md = 0.0, ft = 0.0;
    for (int y = 0; y < p.JS; y++) {
      md += p.meanpart[(int64_t)y * p.ld + c];
      ft += p.ftpart[(int64_t)y * p.ld + c];
    }
    double u2 = 0.0;
    if (p.p == 1) {
      mean = p.beta + md;
      if (p.ordinary) {
        const double u = (ft - 1.0) / p.G;
        u2 = u * u;
      }
    } else {
      // f(x) = [1, x]; mean = beta . f + r . gamma; u = Bhat^T r - Cinv f  (= G^-T (Ft^T rt - f) up
      // to the signs of the rows of G, gpr.py:466-467)
      const double *x = p.Xs + c * p.d;
      double mt = p.betav[0];
      for (int i = 0; i < p.d; i++) mt = fma(p.betav[1 + i], x[i], mt);
      mean = mt + md;
      if (p.trend_est) {
        for (int k = 0; k < p.p; k++) {
          double w = 0.0;
          for (int y = 0; y < p.JS; y++) w += p.tpart[((int64_t)y * p.PT + k) * p.ld + c];
          double cf = p.Cinv[k * MGP_PMAX];
          for (int j = 1; j <= k; j++) cf = fma(p.Cinv[k * MGP_PMAX + j], x[j - 1], cf);
          const double u = w - cf;
          u2 = fma(u, u, u2);
        }
      }
    }
    mse = fabs(p.sigma2 * (1.0 - q + u2));
    if (p.out_mean) p.out_mean[c] = mean;
    if (p.out_mse) p.out_mse[c] = mse;
  }
  if (!p.out_val && !p.blk_val) return;
  const double sd = sqrt(mse);
  const double yhat = p.minimize ? mean : -mean;
  const double plug = p.minimize ? p.plugin : -p.plugin;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int s = 0; s < p.n_sets; s++) {
    double v = valid ? acq_eval(p.kind, p.params[s], plug, yhat, sd, p.sigma2).val : -INFINITY;
    if (p.out_val && valid) p.out_val[(int64_t)s * p.ld_val + c] = v;
    if (!p.blk_val) continue;
    // block arg-max (every thread of the block takes part; padding threads carry -inf)
    double bv = (v == v) ? v : -INFINITY;   // NaN -> -inf, like the selection kernel
    int64_t bi = valid ? p.idx_base + c : INT64_MAX;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    __syncthreads();
    if (lane == 0) { sbv[warp] = bv; sbi[warp] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; w++)
        if (better(sbv[w], sbi[w], bv, bi)) { bv = sbv[w]; bi = sbi[w]; }
      p.blk_val[(int64_t)s * p.ld_blk + blockIdx.x] = bv;
      p.blk_idx[(int64_t)s * p.ld_blk + blockIdx.x] = bi;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// running top-k (what argmax_restart keeps of a population, optimizer/utils.py:66-87)
// total order: larger value first, NaN last, ties -> smaller index first
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_topk_merge(const double *vals, const int64_t *idxs,
                                                     int64_t ld, int64_t mc, int64_t base, int k,
                                                     double *topv, int64_t *topi, int init) {
  __shared__ double sv[32];
  __shared__ int64_t si[32];
  __shared__ double nv[64];
  __shared__ int64_t ni[64];
  __shared__ double lastv;
  __shared__ int64_t lasti;
  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double *v = vals + (int64_t)s * ld;
  double *tv = topv + (int64_t)s * k;
  int64_t *ti = topi + (int64_t)s * k;
  const double NEG = -INFINITY;
  for (int r = 0; r < k; r++) {
    double bv = NEG;
    int64_t bi = INT64_MAX;
    const bool have_last = r > 0;
    const double lv = have_last ? lastv : 0.0;
    const int64_t li = have_last ? lasti : 0;
    auto consider = [&](double cv, int64_t ci) {
      if (!(cv == cv)) cv = NEG;  // NaN -> -inf
      if (have_last && !better(lv, li, cv, ci)) return;  // already selected (or before it)
      if (better(cv, ci, bv, bi)) { bv = cv; bi = ci; }
    };
    for (int64_t i = tid; i < mc; i += 1024) consider(v[i], idxs ? idxs[(int64_t)s * ld + i] : base + i);
    if (!init)
      for (int i = tid; i < k; i += 1024)
        if (ti[i] >= 0) consider(tv[i], ti[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      bv = sv[lane]; bi = si[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) {
        lastv = bv; lasti = bi;
        nv[r] = bv; ni[r] = (bi == INT64_MAX) ? -1 : bi;
      }
    }
    __syncthreads();
  }
  for (int i = tid; i < k; i += 1024) { tv[i] = nv[i]; ti[i] = ni[i]; }
}


// ---------------------------------------------------------------------------------------------
// K5: posterior gradient (GaussianProcess.gradient + corr_dx, gpr.py:511-617) for every candidate
// and the acquisition gradient (InfillCriteria.py:99-110,139-147,177-186,238-251).
// With w = R^-1 r and a = L^-T Ft:
//   mean_dx_i = sum_j gamma_j r_dx[j,i]
//   mse_dx_i  = 2 sigma2 sum_j (-w_j + kappa a_j) r_dx[j,i],  kappa = (Ft^T rt - 1)/(Ft^T Ft)
//   r_dx[j,i] = -2 theta_i D_i r_j (SE, gpr.py:592) | -3 theta_i D_i e^{-sqrt3 D} (Matern, :601-602)
//             | -theta_i sign(D_i) r_j (absolute exponential, :606-607)
// One warp per candidate; lanes stride over the training points; D = x* - x_j component-wise.
// A Matern candidate coinciding with a training point zeroes the whole Jacobian (gpr.py:588-615).
// ---------------------------------------------------------------------------------------------
// Work decomposition.  A thread owns ONE candidate (its accumulators never leave registers) and a
// SLICE of DPS <= 16 features; LPC = 1, 2 or 4 adjacent lanes share a candidate when dpad > 16, so
// that 3 x DPS doubles of state fit the register file for every d <= 64 (the previous
// warp-per-candidate kernel kept 5 x DP doubles per lane and spilled 0.5 - 6 KB per thread for
// DP >= 24).  The training rows stream through shared memory in tiles, pre-scaled by
// c_i = sqrt(theta_i) (theta_i for the absolute-exponential kernel), and are read as broadcasts;
// r and W = R^-1 r come straight from their packed tiles.  Only the Matern distance (and the trend
// projection of a general basis) needs the other slices: one or two shuffles per row.
// NG row groups per candidate split the rows of a tile and combine their sums through shared memory
// in a fixed order: NG = 1 for large populations (128 / LPC candidates per CTA, nothing to combine),
// NG = 8 for a few thousand candidates, NG = 128 / LPC (one CTA per candidate) below that and for the
// one-point-per-call pattern -- enough CTAs to fill the machine at every population size.
constexpr int GR_THREADS = 128;
template <int DPS, int LPC, int NGSEL>     // NGSEL: 0 -> NG = 1, 1 -> NG = 8, 2 -> NG = 128 / LPC
struct GradCfg {
  static constexpr int NG = NGSEL == 0 ? 1 : (NGSEL == 1 ? 8 : GR_THREADS / LPC);   // row groups
  static constexpr int NCAND = GR_THREADS / (LPC * NG);       // candidates per CTA
  static constexpr int JT = 32;                               // training rows per shared-memory tile (NG == 1)
  static constexpr int DPW = DPS * LPC;                       // padded feature width
  static constexpr int KS = DPS + 1;                          // trend coefficients per slice
  static constexpr int KW = KS * LPC;
  static constexpr int RED = 2 * DPS + 2;                     // per thread: am | as | q | coincide
  static int smem_doubles(bool gen) {
    const int tile = NG == 1 ? JT * DPW + 2 * JT + (gen ? JT * KW : 0) : NCAND * LPC * RED;   // NG > 1: the totals
    return tile + (gen ? 2 * NCAND * KW : 0) + (NG > 1 ? GR_THREADS * RED : 0);
  }
};

template <int CORR, int DPS, int LPC, int NGSEL>
__global__ void __launch_bounds__(GR_THREADS, 2) k_grad(GradParams p) {
  using C = GradCfg<DPS, LPC, NGSEL>;
  constexpr bool WIDE = C::NG > 1;
  extern __shared__ __align__(16) double gsm[];
  const bool gen = p.p > 1 && p.trend_est;
  constexpr int TILE = C::NG == 1 ? C::JT * C::DPW + 2 * C::JT : C::NCAND * LPC * C::RED;
  double *sX = gsm;                                 // NG == 1: [JT][DPW] scaled rows; NG > 1: the totals
  double *sG = sX + C::JT * C::DPW;                 // [JT] gamma                   (NG == 1)
  double *sA = sG + C::JT;                          // [JT] a = L^-T Ft             (NG == 1)
  double *sB = gsm + TILE;                          // [JT][KW] Bhat rows           (NG == 1, general trend)
  double *sU = sB + ((gen && C::NG == 1) ? C::JT * C::KW : 0);   // [NCAND][KW] u = Bhat^T r - Cinv f
  double *sK = sU + (gen ? C::NCAND * C::KW : 0);   // [NCAND][KW] kappa = C^-T u
  double *sR = sK + (gen ? C::NCAND * C::KW : 0);   // [THREADS][RED] row-group partial sums (WIDE)
  const int tid = threadIdx.x, s = tid % LPC, grp = tid / LPC;
  const int g = grp % C::NG, cl = grp / C::NG;     // row group, candidate within the CTA
  const int64_t c = (int64_t)blockIdx.x * C::NCAND + cl;
  const bool valid = c < p.mc;
  const int64_t cv = valid ? c : 0;                 // invalid threads shadow candidate 0, write nothing
  const int cb = (int)(cv >> 7), c8 = (int)((cv >> 3) & 15), crow = (int)(cv & 7);

  double md = 0.0, ft = 0.0;
  for (int y = 0; y < p.JS; y++) {
    md += p.meanpart[(int64_t)y * p.ldp + cv];
    if (p.p == 1) ft += p.ftpart[(int64_t)y * p.ldp + cv];
  }
  const double kappa = (p.p == 1 && p.ordinary) ? (ft - 1.0) / p.ftf : 0.0;
  // lanes that share this candidate (adjacent): the only ones a shuffle may name, since in WIDE mode
  // the row groups of a warp can leave the row loop at different times
  const unsigned gmask = LPC == 1 ? 0u : (((1u << LPC) - 1u) << ((tid & 31) & ~(LPC - 1)));
  // ---- general trend: u, kappa, |u|^2 per candidate (gpr.py:465-472, 548-551) ------------------
  double u2g = 0.0;
  if (gen) {
    const double *x = p.Xs + cv * p.d;
#pragma unroll
    for (int kk = 0; kk < C::KS; kk++) {
      const int k = s * C::KS + kk;
      double u = 0.0;
      if (k < p.p) {
        double w = 0.0;
        for (int y = 0; y < p.JS; y++) w += p.tpart[((int64_t)y * p.PT + k) * p.ldp + cv];
        double cf = p.Cinv[k * MGP_PMAX];
        for (int j = 1; j <= k; j++) cf = fma(p.Cinv[k * MGP_PMAX + j], x[j - 1], cf);
        u = w - cf;
      }
      if (!WIDE || g == 0) sU[cl * C::KW + k] = u;
    }
    __syncthreads();
    for (int k = 0; k < p.p; k++) u2g = fma(sU[cl * C::KW + k], sU[cl * C::KW + k], u2g);
    if (!WIDE || g == 0) {
#pragma unroll
      for (int kk = 0; kk < C::KS; kk++) {
        const int i = s * C::KS + kk;
        double t = 0.0;
        for (int k = i; k < p.p; k++) t = fma(p.Cinv[k * MGP_PMAX + i], sU[cl * C::KW + k], t);
        sK[cl * C::KW + i] = t;
      }
    }
  }
  // ---- this thread's slice of the (scaled, centred) candidate --------------------------------
  double xs[DPS], am[DPS], as[DPS];
#pragma unroll
  for (int i = 0; i < DPS; i++) {
    const int f = s * DPS + i;
    const double th = f < p.dpad ? p.theta[f] : 0.0;
    const double cf = CORR == 2 ? th : sqrt(th);
    xs[i] = f < p.d ? cf * (p.Xs[cv * p.d + f] - p.center[f]) : 0.0;
    am[i] = 0.0;
    as[i] = 0.0;
  }
  double q = 0.0;
  int coincide = 0;
  const double *rbase = p.rP + ((int64_t)cb * p.NJ8 * 16 + c8) * 64 + crow * 8;
  const double *wbase = p.wP + ((int64_t)cb * p.NJ8 * 16 + c8) * 64 + crow * 8;

  // one training row: r, w = its entries of r and R^-1 r; xj = this slice of the scaled row;
  // gam, aj = gamma_j, a_j; bj = this slice of Bhat_j (general trend)
  auto row = [&](double r, double w, const double *xj, double gam, double aj, const double *bj) {
    q = fma(r, w, q);
    double e = r;
    double df[DPS];
    double S = 0.0;
#pragma unroll
    for (int i = 0; i < DPS; i += 2) {
      const double2 x2 = *reinterpret_cast<const double2 *>(xj + i);
      df[i] = xs[i] - x2.x;
      df[i + 1] = xs[i + 1] - x2.y;
      if (CORR == 1) { S = fma(df[i], df[i], S); S = fma(df[i + 1], df[i + 1], S); }
      if (CORR == 2) {
        // absolute exponential: r_dx = -theta_i sign(D_i) r  (gpr.py:606-607; numpy sign(0) = 0)
        df[i] = df[i] > 0.0 ? 1.0 : (df[i] < 0.0 ? -1.0 : 0.0);
        df[i + 1] = df[i + 1] > 0.0 ? 1.0 : (df[i + 1] < 0.0 ? -1.0 : 0.0);
      }
    }
    double proj = 0.0;    // (Ahat kappa)_j: Bhat_j . u for a general trend, kappa a_j for the constant one
    if (gen) {
      const double *uj = sU + cl * C::KW + s * C::KS;
#pragma unroll
      for (int kk = 0; kk < C::KS; kk++) proj = fma(bj[kk], uj[kk], proj);
    } else if (p.p == 1) {
      proj = kappa * aj;
    }
    if (LPC > 1) {        // the other slices of this candidate sit in the adjacent lanes
#pragma unroll
      for (int o2 = 1; o2 < LPC; o2 <<= 1) {
        if (CORR == 1) S += __shfl_xor_sync(gmask, S, o2);
        if (gen) proj += __shfl_xor_sync(gmask, proj, o2);
      }
    }
    if (CORR == 1) {
      if (S == 0.0) coincide = 1;
      e = exp(-MGP_SQRT3 * sqrt(S));
    }
    const double cm = gam * e, cs = (proj - w) * e;
#pragma unroll
    for (int i = 0; i < DPS; i++) {
      am[i] = fma(cm, df[i], am[i]);
      as[i] = fma(cs, df[i], as[i]);
    }
  };
  if (C::NG == 1) {
    // large populations: the rows pass through shared memory in tiles and are read as broadcasts
    for (int j0 = 0; j0 < p.n; j0 += C::JT) {
      __syncthreads();                                // previous tile consumed
      const int jend = min(C::JT, p.n - j0);
      {
        const double2 *src = reinterpret_cast<const double2 *>(p.Xsc + (int64_t)j0 * C::DPW);
        double2 *dst = reinterpret_cast<double2 *>(sX);
        for (int idx = tid; idx < jend * C::DPW / 2; idx += GR_THREADS) dst[idx] = src[idx];
      }
      if (tid < jend) {
        sG[tid] = p.gamma[j0 + tid];
        sA[tid] = p.p == 1 ? p.avec[j0 + tid] : 0.0;
      }
      if (gen)
        for (int idx = tid; idx < jend * C::KW; idx += GR_THREADS) {
          const int jj = idx / C::KW, k = idx % C::KW;
          sB[idx] = p.Bhat[(int64_t)(j0 + jj) * MGP_TP + k];
        }
      __syncthreads();
      for (int jj = 0; jj < jend; jj++) {
        const int j = j0 + jj;
        const int64_t o = (int64_t)(j >> 3) * 1024 + (j & 7);
        row(rbase[o], wbase[o], sX + jj * C::DPW + s * DPS, sG[jj], sA[jj], sB + jj * C::KW + s * C::KS);
      }
    }
  } else {
    // few candidates: NG row groups per candidate read their rows straight from global memory (the
    // model is L2 resident), two rows in flight per thread; no tile, no barrier in the loop
    const double *xb = p.Xsc + s * DPS, *bb = p.Bhat + s * C::KS;
    int j = g;
    for (; j + C::NG < p.n; j += 2 * C::NG) {
      const int j1 = j + C::NG;
      const int64_t o0 = (int64_t)(j >> 3) * 1024 + (j & 7), o1 = (int64_t)(j1 >> 3) * 1024 + (j1 & 7);
      const double r0 = rbase[o0], w0 = wbase[o0], r1 = rbase[o1], w1 = wbase[o1];
      const double g0 = p.gamma[j], g1 = p.gamma[j1];
      const double a0 = p.p == 1 ? p.avec[j] : 0.0, a1 = p.p == 1 ? p.avec[j1] : 0.0;
      row(r0, w0, xb + (int64_t)j * C::DPW, g0, a0, bb + (int64_t)j * MGP_TP);
      row(r1, w1, xb + (int64_t)j1 * C::DPW, g1, a1, bb + (int64_t)j1 * MGP_TP);
    }
    if (j < p.n) {
      const int64_t o0 = (int64_t)(j >> 3) * 1024 + (j & 7);
      row(rbase[o0], wbase[o0], xb + (int64_t)j * C::DPW, p.gamma[j], p.p == 1 ? p.avec[j] : 0.0,
          bb + (int64_t)j * MGP_TP);
    }
  }
  if (WIDE) {
    // row-group sums -> shared memory; thread v then adds ONE of the NCAND x LPC x RED quantities over
    // the NG row groups in a fixed order (parallel over the quantities), row group 0 picks the totals up
    double *mine = sR + tid * C::RED;
#pragma unroll
    for (int i = 0; i < DPS; i++) { mine[2 * i] = am[i]; mine[2 * i + 1] = as[i]; }
    mine[2 * DPS] = q;
    mine[2 * DPS + 1] = (double)coincide;
    __syncthreads();
    double *tot = sX;                               // the row tiles are consumed: reuse their space
    for (int v = tid; v < C::NCAND * LPC * C::RED; v += GR_THREADS) {
      const int cs2 = v / C::RED, idx = v % C::RED;   // cs2 = candidate * LPC + slice
      const int c2 = cs2 / LPC, s2 = cs2 % LPC;
      double t = 0.0;
      for (int gg = 0; gg < C::NG; gg++) t += sR[((c2 * C::NG + gg) * LPC + s2) * C::RED + idx];
      tot[v] = t;
    }
    __syncthreads();
    if (g != 0) return;
    const double *o = tot + (cl * LPC + s) * C::RED;
#pragma unroll
    for (int i = 0; i < DPS; i++) { am[i] = o[2 * i]; as[i] = o[2 * i + 1]; }
    q = o[2 * DPS];
    coincide = o[2 * DPS + 1] > 0.0 ? 1 : 0;
  }
  if (!valid) return;
  const double fac = CORR == 0 ? -2.0 : (CORR == 1 ? -3.0 : -1.0);
  double mean = p.beta + md;
  double u2 = 0.0;
  if (p.p == 1) {
    if (p.ordinary) {
      const double u = (ft - 1.0) / p.G;
      u2 = u * u;
    }
  } else {
    const double *x = p.Xs + c * p.d;
    double mt = p.betav[0];
    for (int i = 0; i < p.d; i++) mt = fma(p.betav[1 + i], x[i], mt);
    mean = mt + md;
    u2 = u2g;
  }
  const double mse = fabs(p.sigma2 * (1.0 - q + u2));
  const double sd = sqrt(mse);
  if (s == 0) {
    if (p.out_mean) p.out_mean[c] = mean;
    if (p.out_mse) p.out_mse[c] = mse;
  }
  const double yhat = p.minimize ? mean : -mean;
  const double plug = p.minimize ? p.plugin : -p.plugin;
  // a Matern candidate on a training point zeroes the whole Jacobian (gpr.py:588-615)
  double gm[DPS], gs[DPS];
#pragma unroll
  for (int i = 0; i < DPS; i++) {
    const int f = s * DPS + i;
    const double th = f < p.dpad ? p.theta[f] : 0.0;
    const double cf = CORR == 2 ? th : sqrt(th);
    gm[i] = coincide ? 0.0 : fac * cf * am[i];
    gs[i] = coincide ? 0.0 : 2.0 * p.sigma2 * fac * cf * as[i];
    if (p.p > 1 && f < p.d) {
      // linear trend: f_dx = [0; I] (trend.py:106-109): beta^T f_dx = beta_{1+i} (gpr.py:539) and
      // u^T (Ft^T Ft)^-1 (-f_dx) = -kappa_{1+i} (gpr.py:548-551); both survive a zeroed r_dx
      gm[i] += p.betav[1 + f];
      if (gen) gs[i] -= 2.0 * p.sigma2 * sK[cl * C::KW + 1 + f];
    }
    if (f < p.d) {
      if (p.out_mean_dx) p.out_mean_dx[c * p.d + f] = gm[i];
      if (p.out_mse_dx) p.out_mse_dx[c * p.d + f] = gs[i];
    }
  }
  for (int st = 0; st < p.n_sets; st++) {
    const Acq a2 = acq_eval(p.kind, p.params[st], plug, yhat, sd, p.sigma2);
    if (p.out_dx) {
      const bool dead = a2.A == 0.0 && a2.B == 0.0;     // guarded points: zero gradient
#pragma unroll
      for (int i = 0; i < DPS; i++) {
        const int f = s * DPS + i;
        if (f < p.d) {
          const double ydx = p.minimize ? gm[i] : -gm[i];
          const double sd_dx = gs[i] / (2.0 * sd);
          p.out_dx[((int64_t)st * p.ld_val + c) * p.d + f] = dead ? 0.0 : a2.A * ydx + a2.B * sd_dx;
        }
      }
    }
    if (p.out_val && s == 0) p.out_val[(int64_t)st * p.ld_val + c] = a2.val;
  }
}

template <int CORR, int DPS, int LPC, int NGSEL>
static cudaError_t grad_launch(const GradParams &p, cudaStream_t st) {
  using C = GradCfg<DPS, LPC, NGSEL>;
  const bool gen = p.p > 1 && p.trend_est;
  const int smem = C::smem_doubles(gen) * (int)sizeof(double);
  static int smem_set[MGP_MAX_DEVICES] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  dev = dev < 0 || dev >= MGP_MAX_DEVICES ? 0 : dev;
  if (smem > smem_set[dev]) {
    cudaFuncSetAttribute(k_grad<CORR, DPS, LPC, NGSEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    smem_set[dev] = smem;
  }
  const unsigned grid = (unsigned)((p.mc + C::NCAND - 1) / C::NCAND);
  k_grad<CORR, DPS, LPC, NGSEL><<<grid, GR_THREADS, smem, st>>>(p);
  return cudaGetLastError();
}

template <int CORR>
static cudaError_t grad_dispatch(const GradParams &p, cudaStream_t st) {
  // enough CTAs for the machine at every population size (148 SMs, 2+ CTAs each).  The row-group
  // count fixes the order of the sums: values and gradients are bitwise independent of the
  // population size WITHIN each class (with LPC lanes per candidate: fewer than 4736 / LPC, fewer
  // than 37 888 / LPC, or more candidates per pass) -- the lock-step optimisers, whose batch shrinks
  // as trajectories finish, stay inside the first one.
#define GD(S, L)                                                                        \
  {                                                                                     \
    const int64_t w = p.mc * (L);                                                       \
    return w < 4736 ? grad_launch<CORR, S, L, 2>(p, st)                                 \
                    : (w < 37888 ? grad_launch<CORR, S, L, 1>(p, st) : grad_launch<CORR, S, L, 0>(p, st)); \
  }
  if (p.dpad <= 4) GD(4, 1);
  if (p.dpad <= 8) GD(8, 1);
  if (p.dpad <= 12) GD(12, 1);
  if (p.dpad <= 16) GD(16, 1);
  if (p.dpad <= 24) GD(12, 2);
  if (p.dpad <= 32) GD(16, 2);
  if (p.dpad <= 48) GD(12, 4);
  if (p.dpad <= 64) GD(16, 4);
#undef GD
  return cudaErrorInvalidValue;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
int grad_row_stride(int dpad) {   // DPS x LPC of the K5 instantiation that serves this dpad (grad_dispatch)
  return dpad <= 16 ? dpad : (dpad <= 24 ? 24 : (dpad <= 32 ? 32 : (dpad <= 48 ? 48 : 64)));
}
cudaError_t launch_prep_model(const double *Xc, const double *theta, int npad, int dpad,
                              double *Bop, double *an, double *Xsc, int corr, cudaStream_t st) {
  k_prep_model<<<(npad + 127) / 128, 128, 0, st>>>(Xc, theta, npad, dpad, Bop, an, Xsc, grad_row_stride(dpad), corr);
  return cudaGetLastError();
}

template <int CORR, int KS, bool LIN, bool I8 = false>
static cudaError_t rgen_launch(const RgenParams &p, int sm_count, cudaStream_t st) {
  static int occ = 0;
  static bool attr_done[MGP_MAX_DEVICES] = {};
  const int smem = (128 * rgen_stride(4 * KS) + 3 * 128 + MGP_EXP_TAB_DOUBLES +
                    (LIN ? 128 * trend_stride(trend_tiles(KS)) : 0)) * (int)sizeof(double);
  if (first_use_on_device(attr_done)) {
    cudaFuncSetAttribute(k_rgen<CORR, KS, LIN, I8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_rgen<CORR, KS, LIN, I8>, 256, smem) != cudaSuccess || occ < 1)
      occ = 1;
  }
  const int items = p.nCblk * p.JS;
  const int grid = items < sm_count * occ ? items : sm_count * occ;
  k_rgen<CORR, KS, LIN, I8><<<grid, 256, smem, st>>>(p);
  return cudaGetLastError();
}
// K1 can emit the int8 slices of K2c-i8 itself (8 slices, SE / Matern, constant trend, d <= 40)
bool rgen_can_slice(int corr, int dpad, bool general_trend, int slices) {
  return slices == 8 && !general_trend && dpad <= 40 && (corr == MGP_CORR_SE || corr == MGP_CORR_MATERN32);
}
int rgen_trend_cols(int dpad) { return 8 * ((dpad + 1 + 7) / 8); }
template <int CORR>
static cudaError_t rgen_dispatch(const RgenParams &p, int sm_count, cudaStream_t st) {
  switch (p.dpad / 4) {
#define RG(K) case K: return p.Bhat ? rgen_launch<CORR, K, true>(p, sm_count, st) : rgen_launch<CORR, K, false>(p, sm_count, st);
#define RGI(K) case K: return p.Bhat ? rgen_launch<CORR, K, true>(p, sm_count, st) : (p.RI8 ? rgen_launch<CORR, K, false, true>(p, sm_count, st) : rgen_launch<CORR, K, false>(p, sm_count, st));
    RGI(1) RGI(2) RGI(3) RGI(4) RGI(5) RGI(6) RGI(7) RGI(8) RGI(9) RGI(10) RG(11) RG(12) RG(13) RG(14) RG(15) RG(16)
#undef RG
#undef RGI
    default: return cudaErrorInvalidValue;
  }
}
// Bhat^T r for the absolute-exponential kernel with an estimated non-constant trend: K1' (k_rgen_l1)
// has no DMMA distance GEMM whose C fragments could feed the trend product, so the projection is a
// separate pass over the packed r it wrote: tpart[jb][k][c] = sum_{j in block jb} Bhat[j][k] r[c][j].
// One CTA per (candidate block, training block); thread = (candidate, half of the trend columns).
__global__ void __launch_bounds__(256) k_trend_proj(RgenParams p, int PT) {
  extern __shared__ __align__(16) double sBh[];          // [128][PT] rows of Bhat of this training block
  const int cb = blockIdx.x, jb = blockIdx.y, tid = threadIdx.x;
  for (int idx = tid; idx < 128 * PT; idx += 256) {
    const int j = idx / PT, k = idx % PT;
    sBh[idx] = p.Bhat[(int64_t)(jb * 128 + j) * MGP_TP + k];
  }
  __syncthreads();
  const int cl = tid & 127, half = tid >> 7;
  const int c8 = cl >> 3, crow = cl & 7;
  const int k0 = half * (PT / 2), k1 = k0 + PT / 2;
  const double *rb = p.rP + (((int64_t)cb * p.NJ8 + jb * 16) * 16 + c8) * 64 + crow * 8;
  for (int kk = k0; kk < k1; kk += 4) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int j = 0; j < 128; j++) {
      const double r = rb[(int64_t)(j >> 3) * 1024 + (j & 7)];
      const double *b = sBh + j * PT + kk;
      a0 = fma(b[0], r, a0); a1 = fma(b[1], r, a1); a2 = fma(b[2], r, a2); a3 = fma(b[3], r, a3);
    }
    const int64_t c = (int64_t)cb * 128 + cl;
    p.tpart[((int64_t)jb * PT + kk) * p.ldp + c] = a0;
    p.tpart[((int64_t)jb * PT + kk + 1) * p.ldp + c] = a1;
    p.tpart[((int64_t)jb * PT + kk + 2) * p.ldp + c] = a2;
    p.tpart[((int64_t)jb * PT + kk + 3) * p.ldp + c] = a3;
  }
}

template <int DPAD>
static cudaError_t rgen_l1_launch(const RgenParams &p, int sm_count, cudaStream_t st) {
  static int occ = 0;
  static bool attr_done[MGP_MAX_DEVICES] = {};
  const int smem = (DPAD * 132 + 128 * (DPAD + 1) + DPAD + 2 * 128 + MGP_EXP_TAB_DOUBLES) * (int)sizeof(double);
  if (first_use_on_device(attr_done)) {
    cudaFuncSetAttribute(k_rgen_l1<DPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_rgen_l1<DPAD>, 256, smem) != cudaSuccess || occ < 1)
      occ = 1;
  }
  const int items = p.nCblk * p.JS;
  const int grid = items < sm_count * occ ? items : sm_count * occ;
  k_rgen_l1<DPAD><<<grid, 256, smem, st>>>(p);
  return cudaGetLastError();
}
static cudaError_t rgen_l1_dispatch(const RgenParams &p, int sm_count, cudaStream_t st) {
  switch (p.dpad) {
#define RL(D) case D: return rgen_l1_launch<D>(p, sm_count, st);
    RL(4) RL(8) RL(12) RL(16) RL(20) RL(24) RL(28) RL(32) RL(36) RL(40) RL(44) RL(48) RL(52) RL(56) RL(60) RL(64)
#undef RL
    default: return cudaErrorInvalidValue;
  }
}
cudaError_t launch_rgen(const RgenParams &p, int corr, int sm_count, cudaStream_t st) {
  if (corr == MGP_CORR_SE) return rgen_dispatch<0>(p, sm_count, st);
  if (corr == MGP_CORR_MATERN32) return rgen_dispatch<1>(p, sm_count, st);
  if (corr == MGP_CORR_ABSEXP) {
    cudaError_t e = rgen_l1_dispatch(p, sm_count, st);
    if (e != cudaSuccess || !p.Bhat) return e;
    const int PT = rgen_trend_cols(p.dpad);            // multiple of 8
    const int smem = 128 * PT * (int)sizeof(double);
    static int smem_set[MGP_MAX_DEVICES] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    dev = dev < 0 || dev >= MGP_MAX_DEVICES ? 0 : dev;
    if (smem > smem_set[dev]) {
      cudaFuncSetAttribute(k_trend_proj, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      smem_set[dev] = smem;
    }
    k_trend_proj<<<dim3(p.nCblk, p.JS), 256, smem, st>>>(p, PT);
    return cudaGetLastError();
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_trmm(const TrmmParams &p, int sm_count, cudaStream_t st) {
  static bool attr_done[MGP_MAX_DEVICES] = {};
  if (first_use_on_device(attr_done)) {
    cudaFuncSetAttribute(k_trmm<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRMM_SMEM);
    cudaFuncSetAttribute(k_trmm<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRMM_SMEM);
  }
  const int items = p.NI * p.nCblk;
  if (items <= 0) return cudaSuccess;
  const int grid = items < sm_count ? items : sm_count;
  if (p.full) k_trmm<true><<<grid, 384, TRMM_SMEM, st>>>(p);
  else k_trmm<false><<<grid, 384, TRMM_SMEM, st>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_small_mv(const double *M, int npad, int n, const double *rP, int mc, int full,
                            double *qpart, int64_t ldq, double *wP, cudaStream_t st) {
  if (mc < 1 || mc > 8) return cudaErrorInvalidValue;
  if (full) k_small_mv<true><<<npad / 32, 256, 0, st>>>(M, npad, n, rP, mc, qpart, ldq, wP);
  else k_small_mv<false><<<npad / 32, 256, 0, st>>>(M, npad, n, rP, mc, qpart, ldq, wP);
  return cudaGetLastError();
}
cudaError_t launch_symmetrize(double *A, int npad, cudaStream_t st) {
  k_symmetrize<<<dim3(npad / 32, npad / 32), 1024, 0, st>>>(A, npad);
  return cudaGetLastError();
}

cudaError_t launch_epilogue(const EpiParams &p, cudaStream_t st) {
  if (p.mc <= 0) return cudaSuccess;
  k_epilogue<<<(unsigned)((p.mc + 255) / 256), 256, 0, st>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_topk_merge(const double *vals, const int64_t *idxs, int64_t ld, int64_t mc,
                              int64_t base, int n_sets, int k, double *topv, int64_t *topi, int init,
                              cudaStream_t st) {
  if (k > 64) return cudaErrorInvalidValue;
  k_topk_merge<<<n_sets, 1024, 0, st>>>(vals, idxs, ld, mc, base, k, topv, topi, init);
  return cudaGetLastError();
}

cudaError_t launch_grad(const GradParams &p, int corr, cudaStream_t st) {
  if (p.mc <= 0) return cudaSuccess;
  if (corr == MGP_CORR_SE) return grad_dispatch<0>(p, st);
  if (corr == MGP_CORR_MATERN32) return grad_dispatch<1>(p, st);
  if (corr == MGP_CORR_ABSEXP) return grad_dispatch<2>(p, st);
  return cudaErrorInvalidValue;
}
