// Dense FP64 building blocks on DMMA tensor cores (sm_100a): tiled GEMM with triangular
// k-ranges, 128x128 diagonal-block Cholesky / triangular inverse, blocked Cholesky, recursive
// triangular inverse, M^T M, triangular mat-vec, and the slab packing used by the posterior
// kernels.  These implement the reference's scipy.linalg calls on the hot path:
//   linalg.cholesky          gpr.py:681      -> chol_blocked
//   solve_triangular(L, .)   gpr.py:463,685,690,694 -> explicit L^-1 (trtri_blocked) + trmv / TRMM
//   cho_solve((L,True), I)   gpr.py:899      -> lauum_blocked (R^-1 = L^-T L^-1)
#include <map>
#include <mutex>
#include <utility>

#include "linalg.cuh"
#include "mgp_common.cuh"

namespace {

// ---- persistent, warp-specialised DMMA GEMM ---------------------------------------------------
// One CTA per SM: 8 consumer warps (2 x 4 over a 128 x 128 output tile, 64 accumulator registers
// per thread) and a producer warpgroup that streams 16-k stages of both operands into a ring of
// shared-memory buffers with cp.async, each thread's copies completing on the stage's mbarrier
// (cp.async.mbarrier.arrive.noinc).  There is no CTA-wide barrier in the loop: consumer warps
// wait on "full", release on "empty", and the producers run ahead across tile boundaries, so the
// epilogue (and prologue) of one tile overlaps the loads of the next.  Tiles are handed out
// dynamically (atomic counter, longest-k tiles first) through a two-slot mailbox in shared memory.
//
// Operand staging (strides chosen so that every LDS.128 below is bank-conflict free):
//   k-contiguous operand ([128][16], row stride LDK = 24): one LDS.128 per lane feeds two MMA
//     k-steps:  MMA step (p, s), lane k-index lk  <->  stage k = 8 p + 2 lk + s   (both operands);
//   k-major operand ([16][128], row stride LDM = 130): one LDS.128 at row k(p, s) yields two
//     neighbouring rows (columns) m, m + 1 which feed two MMAs; the tile's rows (columns) are
//     therefore processed in the permuted order 16 g + 2 lr + e  (g: pair group, e: even/odd MMA).
constexpr int GK = 16;         // k extent of one pipeline stage
constexpr int NST = 4;         // pipeline stages
constexpr int LDK = 24;
constexpr int LDM = 130;
constexpr int OP_DBL = 128 * LDK;   // doubles per operand per stage (>= 16 * LDM)
constexpr int GEMM_CONSUMERS = 256, GEMM_PRODUCERS = 128;
constexpr int GEMM_SMEM = NST * 2 * OP_DBL * 8 + (2 * NST + 4) * 8 + 16;

__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t *bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

struct TileCoord { int rowblk, colblk, z; };
__device__ __forceinline__ TileCoord decode_tile(const GemmDesc &g, int id, int tx, int ty, int nz) {
  TileCoord t;
  t.z = id % nz;
  const int r = id / nz;
  if (g.order == 3) { t.colblk = r / ty; t.rowblk = r % ty; }             // column-major, ascending
  else if (g.order == 2) { t.rowblk = ty - 1 - r / tx; t.colblk = r % tx; }  // rows descending
  else { t.rowblk = r / tx; t.colblk = r % tx; }                            // rows ascending
  return t;
}

// One 16-k pipeline stage of the 128 x 128 tile for one consumer warp, restricted to the row tiles
// [RT0, RT1) and column tiles [0, CT1) of the warp's 8 x 4 grid of 8 x 8 accumulator tiles.  The
// bounds are template parameters so that the zero part of a triangular operand's diagonal block is
// skipped by warp-uniform branches over straight-line code: predicated-off DMMAs would still pay
// their issue slots (profiles/r01_trmm_c2_v4_source_summary.txt).
// STAIR: diagonal output tile of a symmetric product (only the part on or below the diagonal is
// wanted): per row tile rt only the column tiles below ct_hi are computed,
//   1 / 2: k-contiguous operands, 8-granular:  ct <= rt + D,      D = 8 wr - 4 wc =  0 / -4
//   3 / 4: k-major operands, 16-granular:      ct/2 <= rt/2 + D', D' = 4 wr - 2 wc = 0 / -2
// (the warps with larger D compute everything, those with smaller D nothing: k_dgemm decides).
template <int STAIR>
__host__ __device__ constexpr int stair_ct_hi(int rt, int CT1) {
  int hi = CT1;
  if (STAIR == 1) hi = rt + 1;
  if (STAIR == 2) hi = rt - 3;
  if (STAIR == 3) hi = 2 * ((rt >> 1) + 1);
  if (STAIR == 4) hi = 2 * ((rt >> 1) - 1);
  return hi < 0 ? 0 : (hi > CT1 ? CT1 : hi);
}

template <bool TA, bool TB, int RT0, int RT1, int CT1, int STAIR = 0>
__device__ __forceinline__ void mma_stage(double (&acc)[8][4][2], const double *as, const double *bs,
                                          int rbase, int wc, int lr, int lk) {
#pragma unroll
  for (int p = 0; p < 2; p++) {
    double2 a[8], b[4];   // TA/TB == 0: [tile] = (step 0, step 1);  == 1: [2 s + group] = (even, odd)
    if (!TA) {
#pragma unroll
      for (int rt = RT0; rt < RT1; rt++)
        a[rt] = *reinterpret_cast<const double2 *>(as + (rbase + rt * 8 + lr) * LDK + 8 * p + 2 * lk);
    } else {
#pragma unroll
      for (int s2 = 0; s2 < 2; s2++)
#pragma unroll
        for (int rp = RT0 / 2; rp < (RT1 + 1) / 2; rp++)
          a[4 * s2 + rp] = *reinterpret_cast<const double2 *>(
              as + (8 * p + 2 * lk + s2) * LDM + rbase + 16 * rp + 2 * lr);
    }
    if (!TB) {
#pragma unroll
      for (int ct = 0; ct < CT1; ct++)
        b[ct] = *reinterpret_cast<const double2 *>(bs + (wc * 32 + ct * 8 + lr) * LDK + 8 * p + 2 * lk);
    } else {
#pragma unroll
      for (int s2 = 0; s2 < 2; s2++)
#pragma unroll
        for (int cp = 0; cp < (CT1 + 1) / 2; cp++)
          b[2 * s2 + cp] = *reinterpret_cast<const double2 *>(
              bs + (8 * p + 2 * lk + s2) * LDM + wc * 32 + 16 * cp + 2 * lr);
    }
#pragma unroll
    for (int s2 = 0; s2 < 2; s2++)
#pragma unroll
      for (int rt = RT0; rt < RT1; rt++) {
        const double av = TA ? ((rt & 1) ? a[4 * s2 + (rt >> 1)].y : a[4 * s2 + (rt >> 1)].x)
                             : (s2 ? a[rt].y : a[rt].x);
#pragma unroll
        for (int ct = 0; ct < CT1; ct++) {
          if (ct >= stair_ct_hi<STAIR>(rt, CT1)) continue;   // compile-time after unrolling
          const double bv = TB ? ((ct & 1) ? b[2 * s2 + (ct >> 1)].y : b[2 * s2 + (ct >> 1)].x)
                               : (s2 ? b[ct].y : b[ct].x);
          dmma884(acc[rt][ct][0], acc[rt][ct][1], av, bv);
        }
      }
  }
}

// MT: rows of the output tile.  MT = 64 (k-contiguous A, full k range only) halves the tile so that
// launches with fewer 128-row tiles than half the SMs -- the panel solves and panel updates on the
// serial chain of the blocked Cholesky -- spread over twice as many SMs.
template <bool TA, bool TB, int MT = 128>
__global__ void __launch_bounds__(GEMM_CONSUMERS + GEMM_PRODUCERS, 1) k_dgemm(GemmDesc g, int *sched_g) {
  static_assert(MT == 128 || ((MT == 64 || MT == 32) && !TA), "half / quarter tiles: k-contiguous A only");
  constexpr int RTN = MT / 16;   // 8-row accumulator tiles per consumer warp
  extern __shared__ __align__(128) unsigned char smraw[];
  double *As = reinterpret_cast<double *>(smraw);
  double *Bs = As + NST * OP_DBL;
  uint64_t *full = reinterpret_cast<uint64_t *>(Bs + NST * OP_DBL);
  uint64_t *empty = full + NST;
  uint64_t *sfull = empty + NST;      // tile mailbox: 2 slots
  uint64_t *sempty = sfull + 2;
  int *slot_id = reinterpret_cast<int *>(sempty + 2);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = g.N / 128, ty = g.M / MT, nz = g.n_inner * g.n_outer;
  const int total = tx * ty * nz;
  constexpr int NWARPS = (GEMM_CONSUMERS + GEMM_PRODUCERS) / 32;

  if (tid == 0) {
    for (int s = 0; s < NST; s++) {
      mbar_init(&full[s], GEMM_PRODUCERS);
      mbar_init(&empty[s], GEMM_CONSUMERS / 32);
    }
    for (int s = 0; s < 2; s++) {
      mbar_init(&sfull[s], 1);
      mbar_init(&sempty[s], NWARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const bool producer = warp >= GEMM_CONSUMERS / 32;
  const bool leader = tid == GEMM_CONSUMERS;   // lane 0 of the first producer warp

  // next tile id through the mailbox (-1: no more tiles); every warp of the CTA calls this once
  // per tile, the leader thread fetches from the global counter
  auto next_tile = [&](int q) -> int {
    const int slot = q & 1;
    if (leader) {
      if (q >= 2) mbar_wait(&sempty[slot], (uint32_t)(((q >> 1) - 1) & 1));
      int id;
      for (;;) {
        id = atomicAdd(&sched_g[0], 1);
        if (id >= total) { id = -1; break; }
        if (!g.lower_only) break;
        const TileCoord t = decode_tile(g, id, tx, ty, nz);
        if (t.colblk * 128 < (t.rowblk + 1) * MT) break;
      }
      slot_id[slot] = id;
      mbar_arrive(&sfull[slot]);
    }
    mbar_wait(&sfull[slot], (uint32_t)((q >> 1) & 1));
    const int id = slot_id[slot];
    __syncwarp();
    if (lane == 0) mbar_arrive(&sempty[slot]);
    return id;
  };
  auto k_range = [&](const TileCoord &t, int &klo, int &nk) {
    const int inner = t.z % g.n_inner;
    int khi = (inner == 0 && g.K0 > 0) ? g.K0 : g.K;
    klo = 0;
    if (g.kmode & 1) klo = t.colblk * 128;
    if (g.kmode & 4) klo = max(klo, t.rowblk * 128 - inner * g.koff_in);
    if (g.kmode & 2) khi = min(khi, (t.rowblk + 1) * 128);
    nk = max(0, (khi - klo) / GK);
  };

  if (producer) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
    const int ptid = tid - GEMM_CONSUMERS;
    int64_t x = 0;
    for (int q = 0;; q++) {
      const int id = next_tile(q);
      if (id < 0) break;
      const TileCoord t = decode_tile(g, id, tx, ty, nz);
      const int inner = t.z % g.n_inner, outer = t.z / g.n_inner;
      const int i0 = t.rowblk * MT, j0 = t.colblk * 128;
      int klo, nk;
      k_range(t, klo, nk);
      const double *A = g.A + inner * g.sA_in + outer * g.sA_out;
      const double *B = g.B + inner * g.sB_in + outer * g.sB_out;
      for (int kt = 0; kt < nk; kt++, x++) {
        const int stage = (int)(x % NST);
        if (x >= NST) mbar_wait(&empty[stage], (uint32_t)(((x / NST) - 1) & 1));
        const int k0 = klo + kt * GK;
        double *as = As + stage * OP_DBL, *bs = Bs + stage * OP_DBL;
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int c = ptid + GEMM_PRODUCERS * u;
          if (!TA) {
            const int row = c >> 3, kc = c & 7;
            if (u < MT / 16) cp_async16(as + row * LDK + kc * 2, A + (int64_t)(i0 + row) * g.lda + k0 + kc * 2);
          } else {
            const int kr = c >> 6, mc = c & 63;
            cp_async16(as + kr * LDM + mc * 2, A + (int64_t)(k0 + kr) * g.lda + i0 + mc * 2);
          }
          if (!TB) {
            const int row = c >> 3, kc = c & 7;
            cp_async16(bs + row * LDK + kc * 2, B + (int64_t)(j0 + row) * g.ldb + k0 + kc * 2);
          } else {
            const int kr = c >> 6, mc = c & 63;
            cp_async16(bs + kr * LDM + mc * 2, B + (int64_t)(k0 + kr) * g.ldb + j0 + mc * 2);
          }
        }
        cp_async_mbar_arrive(&full[stage]);
      }
    }
    // the last CTA to finish re-arms the scheduler words for the next launch on this stream
    if (leader) {
      __threadfence();
      if (atomicAdd(&sched_g[1], 1) == (int)gridDim.x - 1) {
        sched_g[0] = 0;
        sched_g[1] = 0;
        __threadfence();
      }
    }
    return;
  }

  // ---- consumers --------------------------------------------------------------------------------
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
  // warp -> (row half wr, column quarter wc): the two warps of an SM sub-partition (warp & 3) take
  // mirrored column quarters, so that skipping by column (below) leaves every sub-partition the
  // same amount of work
  const int wr = warp >> 2, wc = wr ? 3 - (warp & 3) : (warp & 3), lr = lane >> 2, lk = lane & 3;
  int64_t x = 0;
  for (int q = 0;; q++) {
    const int id = next_tile(q);
    if (id < 0) break;
    const TileCoord t = decode_tile(g, id, tx, ty, nz);
    const int inner = t.z % g.n_inner, outer = t.z / g.n_inner;
    const int i0 = t.rowblk * MT, j0 = t.colblk * 128;
    int klo, nk;
    k_range(t, klo, nk);
    const int rbase = wr * (MT / 2);
    const bool tri_lo = (g.kmode & 2) && (t.rowblk + 1) * 128 <= g.K && nk >= 8;
    const bool tri_hi = (g.kmode & 4) && klo == t.rowblk * 128 - inner * g.koff_in && nk >= 8;
    const bool tri_b = (g.kmode & 1) && klo == t.colblk * 128 && nk >= 8;
    // diagonal output tile of a symmetric product: 0 = compute everything, 1..4 = stair (mma_stage),
    // 5 = this warp's part lies entirely above the diagonal
    int stair = 0;
    if (MT == 128 && TA == TB && g.lower_only && t.rowblk == t.colblk) {
      const int D = TA ? 4 * wr - 2 * wc : 8 * wr - 4 * wc, unit = TA ? 2 : 4;
      stair = D >= unit ? 0 : (D == 0 ? (TA ? 3 : 1) : (D == -unit ? (TA ? 4 : 2) : 5));
    }
    double acc[8][4][2];
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int c = 0; c < 4; c++) acc[r][c][0] = acc[r][c][1] = 0.0;

    for (int kt = 0; kt < nk; kt++, x++) {
      const int stage = (int)(x % NST);
      mbar_wait(&full[stage], (uint32_t)((x / NST) & 1));
      const double *as = As + stage * OP_DBL, *bs = Bs + stage * OP_DBL;
      // Diagonal block of a triangular operand: 16-k stage s of that block touches only
      //   A lower triangular, k <= row (kmode 2, last block):   rows >= 16 s        -> row tiles >= rt0
      //   A^T of a lower triangle, k >= row (kmode 4, first):   rows <  16 (s + 1)  -> row tiles <  rt1
      //   B^T of a lower triangle, k >= col (kmode 1, first):   cols <  16 (s + 1)  -> col tiles <  ct1
      // (the skipped products are exact zeros: the diagonal tiles of L and L^-1 are stored with a
      // zero upper triangle).  One rule per stage; the column rule relies on the mirrored wc mapping.
      int rt0 = 0, rt1 = 8, ct1 = 4;
      if (!TA && tri_lo && kt >= nk - 8) rt0 = min(8, max(0, 2 * (kt - (nk - 8)) - 8 * wr));
      else if (TA && tri_hi && kt < 8) rt1 = 2 * min(4, max(0, kt - 4 * wr + 1));
      else if (TB && tri_b && kt < 8) ct1 = 2 * min(2, max(0, kt - 2 * wc + 1));
      if (MT == 128 && TA == TB && stair != 0 && rt0 == 0 && rt1 == 8 && ct1 == 4) {
        if (TA) {
          if (stair == 3) mma_stage<TA, TB, 0, 8, 4, 3>(acc, as, bs, rbase, wc, lr, lk);
          else if (stair == 4) mma_stage<TA, TB, 4, 8, 4, 4>(acc, as, bs, rbase, wc, lr, lk);
        } else {
          if (stair == 1) mma_stage<TA, TB, 0, 8, 4, 1>(acc, as, bs, rbase, wc, lr, lk);
          else if (stair == 2) mma_stage<TA, TB, 4, 8, 4, 2>(acc, as, bs, rbase, wc, lr, lk);
        }
      } else if (MT != 128 || (rt0 == 0 && rt1 == 8 && ct1 == 4)) {
        mma_stage<TA, TB, 0, RTN, 4>(acc, as, bs, rbase, wc, lr, lk);
      } else if (!TA && rt0 > 0) {
        switch (rt0) {
          case 1: mma_stage<TA, TB, 1, 8, 4>(acc, as, bs, rbase, wc, lr, lk); break;
          case 2: mma_stage<TA, TB, 2, 8, 4>(acc, as, bs, rbase, wc, lr, lk); break;
          case 3: mma_stage<TA, TB, 3, 8, 4>(acc, as, bs, rbase, wc, lr, lk); break;
          case 4: mma_stage<TA, TB, 4, 8, 4>(acc, as, bs, rbase, wc, lr, lk); break;
          case 5: mma_stage<TA, TB, 5, 8, 4>(acc, as, bs, rbase, wc, lr, lk); break;
          case 6: mma_stage<TA, TB, 6, 8, 4>(acc, as, bs, rbase, wc, lr, lk); break;
          case 7: mma_stage<TA, TB, 7, 8, 4>(acc, as, bs, rbase, wc, lr, lk); break;
          default: break;
        }
      } else if (TA && rt1 < 8) {
        switch (rt1) {
          case 2: mma_stage<TA, TB, 0, 2, 4>(acc, as, bs, rbase, wc, lr, lk); break;
          case 4: mma_stage<TA, TB, 0, 4, 4>(acc, as, bs, rbase, wc, lr, lk); break;
          case 6: mma_stage<TA, TB, 0, 6, 4>(acc, as, bs, rbase, wc, lr, lk); break;
          default: break;
        }
      } else if (TB && ct1 < 4) {
        if (ct1 == 2) mma_stage<TA, TB, 0, 8, 2>(acc, as, bs, rbase, wc, lr, lk);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
    }

    // ---- epilogue: C = alpha * acc + beta * C -------------------------------------------------
    double *C = g.C + inner * g.sC_in + outer * g.sC_out;
#pragma unroll
    for (int rt = 0; rt < RTN; rt++) {
      const int row = i0 + rbase + (TA ? 16 * (rt >> 1) + 2 * lr + (rt & 1) : rt * 8 + lr);
      double *crow = C + (int64_t)row * g.ldc + j0 + wc * 32;
      if (!TB) {
#pragma unroll
        for (int ct = 0; ct < 4; ct++) {
          double2 *ptr = reinterpret_cast<double2 *>(crow + ct * 8 + 2 * lk);
          double2 v = make_double2(g.alpha * acc[rt][ct][0], g.alpha * acc[rt][ct][1]);
          if (g.beta != 0.0) {
            const double2 o = *ptr;
            v.x += g.beta * o.x;
            v.y += g.beta * o.y;
          }
          *ptr = v;
        }
      } else {
#pragma unroll
        for (int cp = 0; cp < 2; cp++) {
          double2 *ptr = reinterpret_cast<double2 *>(crow + 16 * cp + 4 * lk);
          double2 v0 = make_double2(g.alpha * acc[rt][2 * cp][0], g.alpha * acc[rt][2 * cp + 1][0]);
          double2 v1 = make_double2(g.alpha * acc[rt][2 * cp][1], g.alpha * acc[rt][2 * cp + 1][1]);
          if (g.beta != 0.0) {
            const double2 o0 = ptr[0], o1 = ptr[1];
            v0.x += g.beta * o0.x; v0.y += g.beta * o0.y;
            v1.x += g.beta * o1.x; v1.y += g.beta * o1.y;
          }
          ptr[0] = v0;
          ptr[1] = v1;
        }
      }
    }
  }

}

// ---- 128x128 diagonal block: Cholesky + inverse, sub-blocked by 32 ---------------------------
// The serial critical path of the blocked Cholesky.  One CTA (512 threads) per matrix:
//   for each 32-wide sub-column: warp 0 factors the 32x32 diagonal sub-block in registers (lane =
//   row, the pivot column travels through 32 doubles of shared memory) and inverts it (lane =
//   column of the inverse, forward substitution); all warps then form the sub-panel
//   P = A D^-T and the symmetric update of the remaining rows.
// The inverse of the whole block follows from the four 32x32 inverses by block rows:
//   X_ij = -X_ii (sum_{k=j}^{i-1} L_ik X_kj).
// Row strides of 4 (mod 16) doubles: the 8x8x4 DMMA fragments of the small products (lane (lr, lk)
// reads element [lr][lk] or [lk][lr]) then fall on 16 distinct bank pairs per half-warp.  Per-lane
// row accesses (lane = row) would be 8-way conflicted at this stride; the one place that needs
// them, the register factorisation of a 32x32 sub-block, goes through a stride-33 scratch copy.
constexpr int DS = 132;   // smem row stride of the 128x128 block
constexpr int XS = 36;    // row stride of a 32x32 inverse
constexpr int PS = 33;    // row stride of the factorisation scratch (inside one XS-strided slot)
constexpr int DIAG_THREADS = 512;
constexpr int kDiagSmemDbl = 128 * DS + 4 * 32 * XS + 64 + 4 * 32;

__device__ __forceinline__ void warp_potrf32(double *D, double *scr, double *col, int lane, int pivot0,
                                             int *status) {
  // Straight-line code (fully unrolled, no branches): the critical path per pivot is
  // shuffle -> rsqrt -> multiply -> smem broadcast -> one FMA; the remaining updates of the step
  // overlap the next pivot's latency.  `col` is double-buffered (2 x 32 doubles).  The sub-block
  // travels through `scr` (32 x PS): row-wise copies with lane = column, register loads with lane = row.
  double a[32];
  // all loads first, then all stores: D and scr may alias as far as the compiler knows, and a
  // load-store-load chain would serialise on shared-memory latency
#pragma unroll
  for (int r = 0; r < 32; r++) a[r] = D[r * DS + lane];
#pragma unroll
  for (int r = 0; r < 32; r++) scr[r * PS + lane] = a[r];
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 32; k++) a[k] = scr[lane * PS + k];
  int bad = 0;
#pragma unroll
  for (int j = 0; j < 32; j++) {
    const double d = __shfl_sync(0xffffffffu, a[j], j);
    bad = (bad == 0 && !(d > 0.0)) ? pivot0 + j + 1 : bad;
    const double inv = rsqrt(d);
    const double l = (lane == j) ? d * inv : a[j] * inv;
    a[j] = l;
    double *cb = col + (j & 1) * 32;
    cb[lane] = l;
    __syncwarp();
#pragma unroll
    for (int k = j + 1; k < 32; k++) a[k] = fma(-l, cb[k], a[k]);
  }
  if (bad && lane == 0) atomicCAS(status, 0, bad);
#pragma unroll
  for (int k = 0; k < 32; k++) scr[lane * PS + k] = (k <= lane) ? a[k] : 0.0;
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 32; r++) a[r] = scr[r * PS + lane];
#pragma unroll
  for (int r = 0; r < 32; r++) D[r * DS + lane] = a[r];
  __syncwarp();
}

// X = D^-1 for a 32x32 lower-triangular D (stride DS); lane owns column `lane` of X.
__device__ __forceinline__ void warp_trtri32(const double *D, double *X, double *rdiag, int lane) {
  rdiag[lane] = 1.0 / D[lane * DS + lane];
  __syncwarp();
  double x[32];
#pragma unroll
  for (int i = 0; i < 32; i++) {
    double s[4] = {(i == lane) ? 1.0 : 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < i; k++) s[k & 3] = fma(-D[i * DS + k], x[k], s[k & 3]);
    x[i] = ((s[0] + s[1]) + (s[2] + s[3])) * rdiag[i];
  }
#pragma unroll
  for (int i = 0; i < 32; i++) X[i * XS + lane] = (lane <= i) ? x[i] : 0.0;
}

// S holds L (lower, 128x128, stride DS), Xd the four 32x32 diagonal inverses.  On return S holds
// L^-1 (lower part; the strictly upper part is unspecified).
// 4x4 register patch of a small shared-memory matrix product: acc[u][v] += sum_k A[u][k] B[k][v].
// A rows are (r0 + u) * lda + k.  BT: B given row-wise as (c0 + v) * ldb + k, else as k * ldb + c0 + v.
// Eight loads feed sixteen FMAs (the one-output-per-thread loops this replaces needed two per FMA).
template <bool BT>
__device__ __forceinline__ void patch4x4(double (&acc)[4][4], const double *A, int lda, const double *B,
                                         int ldb, int k0, int k1) {
#pragma unroll 4
  for (int k = k0; k < k1; k++) {
    double a[4], bq[4];
#pragma unroll
    for (int u = 0; u < 4; u++) a[u] = A[u * lda + k];
#pragma unroll
    for (int v = 0; v < 4; v++) bq[v] = BT ? B[v * ldb + k] : B[k * ldb + v];
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int v = 0; v < 4; v++) acc[u][v] = fma(a[u], bq[v], acc[u][v]);
  }
}

__device__ void diag_block_inverse(double *S, const double *Xd, int tid) {
  for (int idx = tid; idx < 4 * 32 * 32; idx += DIAG_THREADS) {
    const int s = idx >> 10, r = (idx >> 5) & 31, c = idx & 31;
    S[(32 * s + r) * DS + 32 * s + c] = Xd[s * 32 * XS + r * XS + c];
  }
  __syncthreads();
  for (int i = 1; i < 4; i++) {
    // block row i: 32 rows x 32 i columns = 8 x 8 i patches, one per thread
    const int npc = 8 * i;
    const bool on = tid < 8 * npc;
    const int pr = tid / npc, pc = tid - pr * npc;
    const int r0 = 32 * i + 4 * pr, c0 = 4 * pc, j = c0 >> 5;
    double acc[4][4];
    // T_ij = sum_{k=j}^{i-1} L_ik X_kj, from the untouched L blocks of block row i
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int v = 0; v < 4; v++) acc[u][v] = 0.0;
    if (on) patch4x4<false>(acc, S + r0 * DS, DS, S + c0, DS, 32 * j, 32 * i);
    __syncthreads();
    if (on) {
#pragma unroll
      for (int u = 0; u < 4; u++)
#pragma unroll
        for (int v = 0; v < 4; v++) S[(r0 + u) * DS + c0 + v] = acc[u][v];
    }
    __syncthreads();
    // X_ij = -X_ii T_ij   (X_ii lower triangular: k <= row)
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int v = 0; v < 4; v++) acc[u][v] = 0.0;
    if (on) patch4x4<false>(acc, S + r0 * DS + 32 * i, DS, S + 32 * i * DS + c0, DS, 0, 4 * pr + 4);
    __syncthreads();
    if (on) {
#pragma unroll
      for (int u = 0; u < 4; u++)
#pragma unroll
        for (int v = 0; v < 4; v++) S[(r0 + u) * DS + c0 + v] = -acc[u][v];
    }
    __syncthreads();
  }
}

#ifdef DIAG_PROFILE
__device__ long long g_diag_clk[16];
#define DIAG_T(i) do { __syncthreads(); if (threadIdx.x == 0 && blockIdx.x == 0) g_diag_clk[i] = clock64(); } while (0)
#else
#define DIAG_T(i)
#endif

// Barrier of the 15 warps that work beside the factoring warp (named barrier 1).
__device__ __forceinline__ void bar_side() { asm volatile("bar.sync 1, 480;\n" ::: "memory"); }
template <int NW>
__device__ __forceinline__ void bar_side_n() { asm volatile("bar.sync 1, %0;\n" ::"n"(NW * 32) : "memory"); }

// Block row i >= 1 of the inverse, in place: rows < i of S already hold the inverse (diagonal
// sub-blocks included), row i still holds L (already saved by the caller).  ALL: executed by the
// whole CTA (16 warps, __syncthreads) or by warps 1..15 only (bar_side).  The 32 x 32 i block row
// is cut into 8 x 8 DMMA tiles, dealt round-robin to the warps; the small products run on the
// tensor pipe because 4 x 4 register patches fed by scalar shared-memory loads are bound by
// shared-memory bandwidth (8 loads per 16 FMAs; one fragment pair feeds 256).  With the row stride
// DS = 4 (mod 16) every fragment load is conflict-free (at stride 129 they were 4-way conflicted
// and this version was slower than the patches: 63 against 56 us for the whole kernel).
template <bool ALL, int NSIDE = DIAG_THREADS / 32 - 1>
__device__ __forceinline__ void inverse_block_row(double *S, const double *Xd, int i, int tid) {
  auto sync = [] { if (ALL) __syncthreads(); else bar_side_n<NSIDE>(); };
  constexpr int NW = ALL ? DIAG_THREADS / 32 : NSIDE;
  constexpr int NT = NW * 32;
  const int t = ALL ? tid : tid - 32, w = t >> 5, lane = t & 31, lr = lane >> 2, lk = lane & 3;
  const int ntile = 16 * i;             // 4 tile rows x 4 i tile columns
  constexpr int MAXT = 4;               // ceil(48 / 15)
  double acc[MAXT][2];
  // T_ij = sum_{k=j}^{i-1} L_ik X_kj : tile (tr, tc), k from 32 (tc / 4) to 32 i
#pragma unroll
  for (int x = 0; x < MAXT; x++) {
    acc[x][0] = acc[x][1] = 0.0;
    const int tl = w + NW * x;
    if (tl < ntile) {
      const int tr = tl & 3, tc = tl >> 2;
      const double *ap = S + (32 * i + 8 * tr + lr) * DS + lk;
      const double *bp = S + lk * DS + 8 * tc + lr;
      for (int k = 32 * (tc >> 2); k < 32 * i; k += 4) dmma884(acc[x][0], acc[x][1], ap[k], bp[k * DS]);
    }
  }
  sync();
#pragma unroll
  for (int x = 0; x < MAXT; x++) {
    const int tl = w + NW * x;
    if (tl < ntile) {
      const int tr = tl & 3, tc = tl >> 2;
      double *cp = S + (32 * i + 8 * tr + lr) * DS + 8 * tc + 2 * lk;
      cp[0] = acc[x][0];
      cp[1] = acc[x][1];
    }
  }
  for (int idx = t; idx < 32 * 32; idx += NT) {   // X_ii replaces L_ii
    const int r = idx >> 5, c = idx & 31;
    S[(32 * i + r) * DS + 32 * i + c] = Xd[i * 32 * XS + r * XS + c];
  }
  sync();
  // X_ij = -X_ii T_ij   (X_ii lower triangular: k < 8 (tr + 1))
#pragma unroll
  for (int x = 0; x < MAXT; x++) {
    acc[x][0] = acc[x][1] = 0.0;
    const int tl = w + NW * x;
    if (tl < ntile) {
      const int tr = tl & 3, tc = tl >> 2;
      const double *ap = S + (32 * i + 8 * tr + lr) * DS + 32 * i + lk;
      const double *bp = S + (32 * i + lk) * DS + 8 * tc + lr;
      for (int k = 0; k < 8 * (tr + 1); k += 4) dmma884(acc[x][0], acc[x][1], ap[k], bp[k * DS]);
    }
  }
  sync();
#pragma unroll
  for (int x = 0; x < MAXT; x++) {
    const int tl = w + NW * x;
    if (tl < ntile) {
      const int tr = tl & 3, tc = tl >> 2;
      double *cp = S + (32 * i + 8 * tr + lr) * DS + 8 * tc + 2 * lk;
      cp[0] = -acc[x][0];
      cp[1] = -acc[x][1];
    }
  }
  sync();
}

// Rows [32 i, 32 i + 32) of the 128-wide lower-triangular block in S -> global (zeros above the
// diagonal), 16-byte stores.  NT threads, thread index t.
__device__ __forceinline__ void store_block_row(const double *S, double *G, int ldg, int i, int t, int NT) {
  for (int idx = t; idx < 32 * 64; idx += NT) {
    const int r = 32 * i + (idx >> 6), c = 2 * (idx & 63);
    double2 v;
    v.x = (c <= r) ? S[r * DS + c] : 0.0;
    v.y = (c + 1 <= r) ? S[r * DS + c + 1] : 0.0;
    *reinterpret_cast<double2 *>(G + (int64_t)r * ldg + c) = v;
  }
}

__global__ void __launch_bounds__(DIAG_THREADS) k_potrf_diag(double *A, double *Lout, int ld,
                                                              int64_t sA, double *linv,
                                                              int64_t sInv, int kb, int *status) {
  // While warp 0 factors and inverts the 32x32 diagonal sub-block s (the serial part), warps 1..15
  // save block row s-1 of L, turn it into block row s-1 of the inverse and save that too; only the
  // last block row is left for the end.
  extern __shared__ __align__(16) double S[];
  double *Xd = S + 128 * DS, *col = Xd + 4 * 32 * XS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.x;
  const double *Ab = A + b * sA + (int64_t)kb * 128 * ld + kb * 128;
  double *Lb = Lout + b * sA + (int64_t)kb * 128 * ld + kb * 128;
  double *Ib = linv + b * sInv + (int64_t)kb * 128 * 128;
  DIAG_T(0);
  {
    // lower triangle only, 16-byte loads, all of a thread's loads in flight together
    double2 v[16];
#pragma unroll
    for (int q = 0; q < 16; q++) {
      const int idx = tid + DIAG_THREADS * q, r = idx >> 6, c = 2 * (idx & 63);
      v[q] = (c <= r) ? *reinterpret_cast<const double2 *>(Ab + (int64_t)r * ld + c) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int q = 0; q < 16; q++) {
      const int idx = tid + DIAG_THREADS * q, r = idx >> 6, c = 2 * (idx & 63);
      S[r * DS + c] = v[q].x;
      S[r * DS + c + 1] = (c + 1 <= r) ? v[q].y : 0.0;
    }
  }
  __syncthreads();
  DIAG_T(1);
  for (int s = 0; s < 4; s++) {
    const int o = 32 * s, nb = 96 - o;   // nb rows below the sub-block
    if (warp == 0) {
      warp_potrf32(S + o * DS + o, Xd + s * 32 * XS, col, lane, kb * 128 + o, &status[b]);
      __syncwarp();
#ifdef DIAG_PROFILE
      if (s == 0 && lane == 0 && blockIdx.x == 0) g_diag_clk[8] = clock64();
#endif
      warp_trtri32(S + o * DS + o, Xd + s * 32 * XS, col + 64, lane);
    } else if (s > 0) {
      const int t = tid - 32, i = s - 1;
      store_block_row(S, Lb, ld, i, t, DIAG_THREADS - 32);
      bar_side();
      if (i == 0) {
        for (int idx = t; idx < 32 * 32; idx += DIAG_THREADS - 32)
          S[(idx >> 5) * DS + (idx & 31)] = Xd[(idx >> 5) * XS + (idx & 31)];
        bar_side();
      } else {
        inverse_block_row<false>(S, Xd, i, tid);
      }
      store_block_row(S, Ib, 128, i, t, DIAG_THREADS - 32);
    }
    __syncthreads();
    if (s == 0) DIAG_T(2);
    if (nb > 0) {
      // sub-panel P = A[below][o:o+32] D^-T : P[r][c] = sum_k A[r][o+k] X[c][k]  (X[c][k] = 0 for k > c).
      // Warp w owns the 8-row strip w (4 DMMA tiles): it reads and overwrites only its own rows.
      const double *X = Xd + s * 32 * XS;
      const int lr = lane >> 2, lk = lane & 3;
      if (warp < nb / 8) {
        double *row = S + (o + 32 + 8 * warp + lr) * DS + o;
        double acc[4][2];
#pragma unroll
        for (int tc = 0; tc < 4; tc++) acc[tc][0] = acc[tc][1] = 0.0;
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
          const double a = row[k + lk];
#pragma unroll
          for (int tc = 0; tc < 4; tc++)
            if (k < 8 * (tc + 1)) dmma884(acc[tc][0], acc[tc][1], a, X[(8 * tc + lr) * XS + k + lk]);
        }
        __syncwarp();
#pragma unroll
        for (int tc = 0; tc < 4; tc++) {
          row[8 * tc + 2 * lk] = acc[tc][0];
          row[8 * tc + 2 * lk + 1] = acc[tc][1];
        }
      }
      __syncthreads();
      if (s == 0) DIAG_T(9);
      // symmetric update of the remaining rows: A[r][c] -= sum_k P[r][k] P[c][k]  (c <= r): the lower
      // 8 x 8 tiles of the (nb / 8)^2 tile grid, dealt round-robin to the 16 warps (at most 5 each)
      {
        const int np8 = nb / 8, ntile = np8 * (np8 + 1) / 2;
        const double *P = S + (o + 32) * DS + o;
        double *Cp = S + (o + 32) * DS + o + 32;
        double acc[5][2];
        int trv[5], tcv[5];
#pragma unroll
        for (int x = 0; x < 5; x++) {
          acc[x][0] = acc[x][1] = 0.0;
          const int tl = warp + 16 * x;
          int tr = 0;
          while ((tr + 1) * (tr + 2) / 2 <= tl) tr++;
          trv[x] = tr;
          tcv[x] = tl - tr * (tr + 1) / 2;
        }
#pragma unroll
        for (int k = 0; k < 32; k += 4)
#pragma unroll
          for (int x = 0; x < 5; x++)
            if (warp + 16 * x < ntile)
              dmma884(acc[x][0], acc[x][1], P[(8 * trv[x] + lr) * DS + k + lk], P[(8 * tcv[x] + lr) * DS + k + lk]);
#pragma unroll
        for (int x = 0; x < 5; x++)
          if (warp + 16 * x < ntile) {
            const int r = 8 * trv[x] + lr, c = 8 * tcv[x] + 2 * lk;
            if (c <= r) Cp[r * DS + c] -= acc[x][0];
            if (c + 1 <= r) Cp[r * DS + c + 1] -= acc[x][1];
          }
      }
      __syncthreads();
      if (s == 0) DIAG_T(3);
    }
  }
  DIAG_T(4);
  store_block_row(S, Lb, ld, 3, tid, DIAG_THREADS);
  __syncthreads();
  DIAG_T(5);
  inverse_block_row<true>(S, Xd, 3, tid);
  DIAG_T(6);
  store_block_row(S, Ib, 128, 3, tid, DIAG_THREADS);
  DIAG_T(7);
}

// ---- version 2 of the diagonal-block kernel ---------------------------------------------------
// Same result, shorter serial chain.  What version 1 keeps on warp 0's critical path per 32-wide
// sub-column -- the inverse of the sub-block (needed for the sub-panel P = A D^-T), the sub-panel and
// the whole symmetric update -- is taken off it:
//   * warp 0 publishes every column of L_ss (column-major copy `Lc`, double-buffered over s) and
//     1 / l_jj, four columns at a time (release store of a counter); warps 13..15 solve the rows below
//     by forward SUBSTITUTION, one row per lane with the row in registers, consuming the columns as
//     they appear: the sub-panel is complete a few hundred cycles after the sub-block.  No inverse on
//     the path; the inverse of the sub-block is computed by a side warp (the same substitution applied
//     to the unit vectors) while the next sub-block is being factored;
//   * of the symmetric update only the ten 8x8 tiles of the NEXT diagonal sub-block are applied before
//     warp 0 starts factoring it; the rest of the update, the stores and the block row of the inverse
//     run on warps 1..12 meanwhile;
//   * inside the factorisation of a sub-block the pivot chain is  1/sqrt -> l -> (a - l^2) -> shuffle:
//     the next pivot comes from the pivot lane's own registers, the two columns right of the pivot are
//     updated from shuffled values, and the shared-memory round trip of the rest of the column update
//     is deferred by one step so that it runs under the next pivot's 1/sqrt;
//   * 1/sqrt(d) is MUFU.RSQ64H (rsqrt.approx.f64, ~2^-21) with one third-order correction, error
//     O(e^3), instead of the library routine.
constexpr int LCS = 34;      // column stride of Lc (even: the broadcast reads of a column are 16-byte loads)
constexpr int NSIDE2 = 12;   // warps 1..12: side work; warps 13..15: sub-panel rows
constexpr int kDiag2SmemDbl = 128 * DS + 4 * 32 * XS + 2 * 32 * LCS + 2 * 32 + 2;

__device__ __forceinline__ double rsqrt_pos(double d) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;\n" : "=d"(y0) : "d"(d));
  const double t = d * y0;
  const double e = fma(-t, y0, 1.0);          // 1 - d y0^2
  const double p = fma(0.375, e, 0.5);
  const double q = y0 * e;
  return fma(q, p, y0);                       // y0 (1 + e/2 + 3 e^2 / 8)
}

__device__ __forceinline__ void st_release_cta(int *p, int v) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;\n" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_cta(const int *p) {
  int v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}

// a[k] -= p * col[k] for k in [K0, 32): col is 16-byte aligned, broadcast reads as double2
template <int K0>
__device__ __forceinline__ void axpy_from(double (&a)[32], double p, const double *col) {
  constexpr int KE = K0 & ~1;
#pragma unroll
  for (int k = KE; k < 32; k += 2) {
    const double2 c = *reinterpret_cast<const double2 *>(col + k);
    if (k >= K0) a[k] = fma(-p, c.x, a[k]);
    a[k + 1] = fma(-p, c.y, a[k + 1]);
  }
}

template <int J>
__device__ __forceinline__ void potrf32_step(double (&a)[32], double &d, double &lprev, int &bad, double *Lc,
                                             double *rinv, int *ready, int ready0, int lane, int pivot0) {
  // deferred part of column J - 1 (published before the previous step's __syncwarp): independent of
  // the pivot chain below, the scheduler runs it under the latency of the reciprocal square root
  if constexpr (J >= 1) axpy_from<J + 2>(a, lprev, Lc + (J - 1) * LCS);
  bad = (bad == 0 && !(d > 0.0)) ? pivot0 + J + 1 : bad;
  const double inv = rsqrt_pos(d);
  const double l = (lane == J) ? d * inv : a[J] * inv;
  a[J] = l;
  Lc[J * LCS + lane] = (lane >= J) ? l : 0.0;
  if (lane == J) rinv[J] = inv;
  if constexpr (J < 31) {
    // the next pivot from the pivot lane's own registers
    const double dn = fma(-l, l, a[J + 1]);
    d = __shfl_sync(0xffffffffu, dn, J + 1);
    // the two columns the next step needs in full, from the owning lanes' registers
    const double l1 = __shfl_sync(0xffffffffu, l, J + 1);
    a[J + 1] = fma(-l, l1, a[J + 1]);
  }
  if constexpr (J < 30) {
    const double l2 = __shfl_sync(0xffffffffu, l, J + 2);
    a[J + 2] = fma(-l, l2, a[J + 2]);
  }
  lprev = l;
  __syncwarp();
  if constexpr ((J & 3) == 3) {
    if (lane == 0) st_release_cta(ready, ready0 + J + 1);   // columns .. J are visible to the solving warps
  }
  if constexpr (J < 31) potrf32_step<J + 1>(a, d, lprev, bad, Lc, rinv, ready, ready0, lane, pivot0);
}

// Factor the 32x32 sub-block D (stride DS, read only) : Lc[j * LCS + i] = L[i][j] (zeros above the
// diagonal), rinv[j] = 1 / L[j][j].  scr: 32 x PS scratch.  *ready counts the published columns
// (ready0 + 4, + 8, ... + 32).
__device__ __forceinline__ void warp_potrf32_v2(const double *D, double *scr, double *Lc, double *rinv,
                                                int *ready, int ready0, int lane, int pivot0, int *status) {
  double a[32];
#pragma unroll
  for (int r = 0; r < 32; r++) a[r] = D[r * DS + lane];
#pragma unroll
  for (int r = 0; r < 32; r++) scr[r * PS + lane] = a[r];
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 32; k++) a[k] = scr[lane * PS + k];
  int bad = 0;
  double d = __shfl_sync(0xffffffffu, a[0], 0), lprev = 0.0;
  potrf32_step<0>(a, d, lprev, bad, Lc, rinv, ready, ready0, lane, pivot0);
  if (bad && lane == 0) atomicCAS(status, 0, bad);
}

// Column-oriented forward substitution with L_ss: a <- the solution x of  x L_ss^T = a  (a row of the
// sub-panel), equally of  L_ss x = a  (a column of the inverse when a is a unit vector).
// WAIT: the columns are consumed while warp 0 is still producing them (four at a time).
template <int J, bool WAIT>
__device__ __forceinline__ void solve32_step(double (&a)[32], const double *Lc, const double *rinv,
                                             const int *ready, int ready0) {
  if constexpr (WAIT && (J & 3) == 0) {
    while (ld_acquire_cta(ready) < ready0 + J + 4) { }
  }
  const double p = a[J] * rinv[J];
  a[J] = p;
  axpy_from<J + 1>(a, p, Lc + J * LCS);
  if constexpr (J < 31) solve32_step<J + 1, WAIT>(a, Lc, rinv, ready, ready0);
}
// One row of the sub-panel P = A L_ss^-T, in place (row = 32 doubles at `row`, 16-byte aligned).
__device__ __forceinline__ void lane_solve32(double *row, const double *Lc, const double *rinv,
                                             const int *ready, int ready0) {
  double a[32];
#pragma unroll
  for (int k = 0; k < 32; k += 2) {
    const double2 v = *reinterpret_cast<const double2 *>(row + k);
    a[k] = v.x; a[k + 1] = v.y;
  }
  solve32_step<0, true>(a, Lc, rinv, ready, ready0);
#pragma unroll
  for (int k = 0; k < 32; k += 2) *reinterpret_cast<double2 *>(row + k) = make_double2(a[k], a[k + 1]);
}
// X = L_ss^-1: column `lane` of X solves L x = e_lane.
__device__ __forceinline__ void warp_trtri32_lc(const double *Lc, const double *rinv, double *X, int lane) {
  double a[32];
#pragma unroll
  for (int k = 0; k < 32; k++) a[k] = (k == lane) ? 1.0 : 0.0;
  solve32_step<0, false>(a, Lc, rinv, nullptr, 0);
#pragma unroll
  for (int i = 0; i < 32; i++) X[i * XS + lane] = (lane <= i) ? a[i] : 0.0;
}

// C[tr][tc] -= P[tr] P[tc]^T for one 8x8 tile of the rows below sub-block s (o = 32 s): P = rows
// o+32.. of column block s, C = the same rows of the columns right of it.  Two accumulator chains.
__device__ __forceinline__ void upd_tile(double *S, int o, int tr, int tc, int lane) {
  const int lr = lane >> 2, lk = lane & 3;
  const double *P = S + (o + 32) * DS + o;
  const double *ap = P + (8 * tr + lr) * DS + lk, *bp = P + (8 * tc + lr) * DS + lk;
  double e0 = 0.0, e1 = 0.0, f0 = 0.0, f1 = 0.0;
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    dmma884(e0, e1, ap[k], bp[k]);
    dmma884(f0, f1, ap[k + 4], bp[k + 4]);
  }
  double2 *cp = reinterpret_cast<double2 *>(S + (o + 32 + 8 * tr + lr) * DS + o + 32 + 8 * tc + 2 * lk);
  double2 c = *cp;
  c.x -= e0 + f0;
  c.y -= e1 + f1;
  *cp = c;
}

// Named barriers of version 2: 1 = the NSIDE2 side warps among themselves (bar_side_n), 2 = side warps
// (arrive) -> solving warps (sync): the rows below the next sub-block have received step i's update,
// 3 = warps 1..15 in the tail.
__device__ __forceinline__ void bar2_arrive() { asm volatile("bar.arrive 2, %0;\n" ::"n"((NSIDE2 + 3) * 32) : "memory"); }
__device__ __forceinline__ void bar2_sync() { asm volatile("bar.sync 2, %0;\n" ::"n"((NSIDE2 + 3) * 32) : "memory"); }
__device__ __forceinline__ void bar3_sync() { asm volatile("bar.sync 3, 480;\n" ::: "memory"); }

// What warps 1..12 do while warp 0 factors sub-block i + 1: the rest of step i's symmetric update
// (first the tile columns of block column i + 1, which the solving warps are waiting for), L block row
// i -> global, X_ii (warp 1), block row i of the inverse in place, -> global.
__device__ __forceinline__ void diag2_side(double *S, double *Xd, const double *lc, const double *ri,
                                           double *Lb, int ld, double *Ib, int i, int tid) {
  const int warp = tid >> 5, lane = tid & 31, t = tid - 32, o = 32 * i;
  const int np8 = (96 - o) / 8, m = np8 - 4, nLR = 4 * m;
  constexpr int NT = NSIDE2 * 32;
  // rows 4.. of tile columns 0..3 (block column i + 1 below its diagonal sub-block)
  for (int q = warp - 1; q < nLR; q += NSIDE2) upd_tile(S, o, 4 + (q >> 2), q & 3, lane);
  bar2_arrive();
  if (warp == 1) {
    warp_trtri32_lc(lc, ri, Xd + i * 32 * XS, lane);
  } else if (m > 0) {
    // the lower tiles of the (np8 - 4)^2 grid right of block column i + 1
    const int ntile = m * (m + 1) / 2;
    for (int u = warp - 2; u < ntile; u += NSIDE2 - 1) {
      int a = 0;
      while ((a + 1) * (a + 2) / 2 <= u) a++;
      upd_tile(S, o, 4 + a, 4 + u - a * (a + 1) / 2, lane);
    }
  }
  store_block_row(S, Lb, ld, i, t, NT);
  bar_side_n<NSIDE2>();
  if (i == 0) {
    for (int idx = t; idx < 32 * 32; idx += NT)
      S[(idx >> 5) * DS + (idx & 31)] = Xd[(idx >> 5) * XS + (idx & 31)];
    bar_side_n<NSIDE2>();
  } else {
    inverse_block_row<false, NSIDE2>(S, Xd, i, tid);
  }
  store_block_row(S, Ib, 128, i, t, NT);
}

__global__ void __launch_bounds__(DIAG_THREADS) k_potrf_diag2(double *A, double *Lout, int ld,
                                                               int64_t sA, double *linv,
                                                               int64_t sInv, int kb, int *status) {
  extern __shared__ __align__(16) double S[];
  double *Xd = S + 128 * DS, *Lc = Xd + 4 * 32 * XS, *rinv = Lc + 2 * 32 * LCS;
  int *ready = reinterpret_cast<int *>(rinv + 2 * 32);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.x;
  const double *Ab = A + b * sA + (int64_t)kb * 128 * ld + kb * 128;
  double *Lb = Lout + b * sA + (int64_t)kb * 128 * ld + kb * 128;
  double *Ib = linv + b * sInv + (int64_t)kb * 128 * 128;
  DIAG_T(0);
  if (tid == 0) *ready = 0;
  {
    double2 v[16];
#pragma unroll
    for (int q = 0; q < 16; q++) {
      const int idx = tid + DIAG_THREADS * q, r = idx >> 6, c = 2 * (idx & 63);
      v[q] = (c <= r) ? *reinterpret_cast<const double2 *>(Ab + (int64_t)r * ld + c) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int q = 0; q < 16; q++) {
      const int idx = tid + DIAG_THREADS * q, r = idx >> 6, c = 2 * (idx & 63);
      S[r * DS + c] = v[q].x;
      S[r * DS + c + 1] = (c + 1 <= r) ? v[q].y : 0.0;
    }
  }
  __syncthreads();
  DIAG_T(1);
  for (int s = 0; s < 4; s++) {
    const int o = 32 * s, nb = 96 - o;
    double *lc = Lc + (s & 1) * 32 * LCS, *ri = rinv + (s & 1) * 32;
    if (warp == 0) {
      warp_potrf32_v2(S + o * DS + o, Xd + s * 32 * XS, lc, ri, ready, 32 * s, lane, kb * 128 + o, &status[b]);
    } else if (warp <= NSIDE2) {
      if (s > 0)
        diag2_side(S, Xd, Lc + ((s - 1) & 1) * 32 * LCS, rinv + ((s - 1) & 1) * 32, Lb, ld, Ib, s - 1, tid);
    } else {
      // rows below sub-block s: final once step s - 1's update of block column s is in (barrier 2)
      if (s > 0) bar2_sync();
      if (warp - (NSIDE2 + 1) < nb / 32)
        lane_solve32(S + (o + 32 + 32 * (warp - (NSIDE2 + 1)) + lane) * DS + o, lc, ri, ready, 32 * s);
    }
    __syncthreads();   // L_ss published, sub-panel in place, the side work of step s - 1 complete
    if (s == 0) DIAG_T(8);
    if (s == 3) break;
    if (warp >= 1 && warp <= 10) {
      // the ten lower tiles of the next diagonal sub-block
      const int u = warp - 1;
      int a = 0;
      while ((a + 1) * (a + 2) / 2 <= u) a++;
      upd_tile(S, o, a, u - a * (a + 1) / 2, lane);
    } else if (warp > 10) {
      // L_ss row-major into S for the stores and the inverse (nothing reads the part above the diagonal)
      for (int idx = tid - 11 * 32; idx < 32 * 32; idx += 5 * 32)
        S[(o + (idx >> 5)) * DS + o + (idx & 31)] = lc[(idx & 31) * LCS + (idx >> 5)];
    }
    __syncthreads();   // next diagonal sub-block up to date
    if (s == 0) DIAG_T(2);
  }
  DIAG_T(4);
  {
    const double *lc = Lc + 32 * LCS, *ri = rinv + 32;
    if (warp == 0) {
      warp_trtri32_lc(lc, ri, Xd + 3 * 32 * XS, lane);
    } else {
      for (int idx = tid - 32; idx < 32 * 32; idx += DIAG_THREADS - 32)
        S[(96 + (idx >> 5)) * DS + 96 + (idx & 31)] = lc[(idx & 31) * LCS + (idx >> 5)];
      bar3_sync();
      store_block_row(S, Lb, ld, 3, tid - 32, DIAG_THREADS - 32);
    }
  }
  __syncthreads();
  DIAG_T(5);
  inverse_block_row<true>(S, Xd, 3, tid);
  DIAG_T(6);
  store_block_row(S, Ib, 128, 3, tid, DIAG_THREADS);
  DIAG_T(7);
}

__global__ void __launch_bounds__(DIAG_THREADS) k_trtri_diag(const double *L, int ld, int64_t sL,
                                                              double *linv, int64_t sInv) {
  extern __shared__ __align__(16) double S[];
  double *Xd = S + 128 * DS, *col = Xd + 4 * 32 * XS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, kb = blockIdx.x, b = blockIdx.y;
  const double *Lb = L + b * sL + (int64_t)kb * 128 * ld + kb * 128;
  for (int idx = tid; idx < 128 * 128; idx += DIAG_THREADS) {
    const int r = idx >> 7, c = idx & 127;
    S[r * DS + c] = (c <= r) ? Lb[(int64_t)r * ld + c] : 0.0;
  }
  __syncthreads();
  if (warp < 4) warp_trtri32(S + 32 * warp * DS + 32 * warp, Xd + warp * 32 * XS, col + 64 + 32 * warp, lane);
  __syncthreads();
  diag_block_inverse(S, Xd, tid);
  double *Ib = linv + b * sInv + (int64_t)kb * 128 * 128;
  for (int idx = tid; idx < 128 * 128; idx += DIAG_THREADS) {
    const int r = idx >> 7, c = idx & 127;
    Ib[idx] = (c <= r) ? S[r * DS + c] : 0.0;
  }
}

__global__ void __launch_bounds__(1024) k_scatter_diag(const double *linv, int64_t sInv, double *M, int ld,
                                                       int64_t sM, int kb0) {
  const int kb = blockIdx.x + kb0, b = blockIdx.y;
  const double2 *src = reinterpret_cast<const double2 *>(linv + b * sInv + (int64_t)kb * 128 * 128);
  double *dst = M + b * sM + (int64_t)kb * 128 * ld + kb * 128;
#pragma unroll
  for (int q = 0; q < 8; q++) {   // 8192 double2 per tile, 8 independent copies in flight per thread
    const int idx = threadIdx.x + 1024 * q;
    *reinterpret_cast<double2 *>(dst + (int64_t)(idx >> 6) * ld + 2 * (idx & 63)) = src[idx];
  }
}

// ---- triangular mat-vec -------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_trmv_n(const double *M, int npad, int64_t sM,
                                                const double *x, int64_t sx, double *y, int64_t sy,
                                                int n) {
  const int b = blockIdx.y;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= npad) return;
  const double *Mr = M + b * sM + (int64_t)row * npad;
  const double *xb = x ? x + b * sx : nullptr;
  double acc = 0.0;
  for (int k = lane; k <= row; k += 32) {
    const double xv = xb ? xb[k] : (k < n ? 1.0 : 0.0);
    acc = fma(Mr[k], xv, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) y[b * sy + row] = acc;
}

// Two products with one pass over M: y1 = M x1, y2 = M x2 (Yt and Ft of the likelihood share L^-1).
__global__ void __launch_bounds__(256) k_trmv_n2(const double *M, int npad, int64_t sM,
                                                 const double *x1, const double *x2, double *y1, double *y2,
                                                 int64_t sy, int n) {
  const int b = blockIdx.y;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= npad) return;
  const double *Mr = M + b * sM + (int64_t)row * npad;
  double a1 = 0.0, a2 = 0.0;
  for (int k = lane; k <= row; k += 32) {
    const double m = Mr[k];
    a1 = fma(m, x1[k], a1);
    a2 = fma(m, x2 ? x2[k] : (k < n ? 1.0 : 0.0), a2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
  }
  if (lane == 0) { y1[b * sy + row] = a1; y2[b * sy + row] = a2; }
}

// y[col] = sum_{k >= col} M[k][col] x[k]: a block owns 32 columns, its 32 warps stride over the
// rows (coalesced 256-byte row segments), partial sums are combined in a fixed order.
__global__ void __launch_bounds__(1024) k_trmv_t(const double *M, int npad, int64_t sM,
                                                 const double *x, int64_t sx, double *y, int64_t sy,
                                                 int n) {
  __shared__ double red[32][33];
  const int b = blockIdx.y, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;
  const double *Mb = M + b * sM;
  const double *xb = x ? x + b * sx : nullptr;
  double acc = 0.0;
  for (int k = blockIdx.x * 32 + w; k < npad; k += 32) {
    const double xv = xb ? xb[k] : (k < n ? 1.0 : 0.0);
    if (k >= col) acc = fma(Mb[(int64_t)k * npad + col], xv, acc);
  }
  red[w][lane] = acc;
  __syncthreads();
  if (w == 0) {
    double t = 0.0;
    for (int r = 0; r < 32; r++) t += red[r][lane];
    y[b * sy + col] = t;
  }
}

// ---- slab packing ---------------------------------------------------------------------------
// slab = 16 row tiles x (8x8 block, row-major) = 1024 doubles; see posterior.cu for the layout.
__global__ void __launch_bounds__(1024) k_pack_tri(const double *Minv, int n, int npad, double *Mp) {
  const int I = blockIdx.y, j8 = blockIdx.x;
  if (j8 >= 16 * (I + 1)) return;
  const int w = threadIdx.x, i8 = w >> 6, ri = (w >> 3) & 7, kj = w & 7;
  const int64_t slab = (int64_t)8 * I * (I + 1) + j8;
  const int row = 128 * I + 8 * i8 + ri, col = 8 * j8 + kj;
  // padded training slots (identity rows of the padded factor) must not contribute
  Mp[slab * 1024 + w] = (row < n && col < n) ? Minv[(int64_t)row * npad + col] : 0.0;
}
__global__ void __launch_bounds__(1024) k_pack_sym(const double *Rinv, int n, int npad, double *Rp) {
  const int I = blockIdx.y, j8 = blockIdx.x, NJ8 = npad / 8;
  const int w = threadIdx.x, i8 = w >> 6, ri = (w >> 3) & 7, kj = w & 7;
  const int row = 128 * I + 8 * i8 + ri, col = 8 * j8 + kj;
  const int64_t slab = (int64_t)I * NJ8 + j8;
  // only the part on or below the diagonal is valid (diagonal tiles included: the GEMM that forms
  // R^-1 skips what lies above it)
  const bool lower = col <= row;
  double v = lower ? Rinv[(int64_t)row * npad + col] : Rinv[(int64_t)col * npad + row];
  if (row >= n || col >= n) v = 0.0;
  Rp[slab * 1024 + w] = v;
}

}  // namespace

// ---- launchers ------------------------------------------------------------------------------
// Scheduler words {next tile, finished CTAs} per stream: kernels on one stream run back to back and
// re-arm the words themselves; concurrent streams (batch groups) must not share them.
static std::mutex g_sched_mu;
static std::map<std::pair<int, cudaStream_t>, int *> g_sched_table;
static int *sched_words(cudaStream_t st) {
  std::mutex &mu = g_sched_mu;
  auto &table = g_sched_table;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_pair(dev, st);
  auto it = table.find(key);
  if (it != table.end()) return it->second;
  int *p = nullptr;
  if (cudaMalloc(&p, 2 * sizeof(int)) != cudaSuccess) return nullptr;
  // Zero the words ON THE STREAM that will use them: cudaMemset on device memory is asynchronous
  // (legacy stream) and is not ordered against work on a non-blocking stream -- recycled memory
  // would otherwise hand the first kernel a stale tile counter.
  if (cudaMemsetAsync(p, 0, 2 * sizeof(int), st) != cudaSuccess) { cudaFree(p); return nullptr; }
  table[key] = p;
  return p;
}

void release_dgemm_stream(cudaStream_t st) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_sched_mu);
  auto it = g_sched_table.find(std::make_pair(dev, st));
  if (it == g_sched_table.end()) return;
  cudaFree(it->second);
  g_sched_table.erase(it);
}

// MGP_GEMM_QUARTER=0 switches the 32-row tiles off (A/B measurements)
static bool quarter_tiles() {
  static const bool on = [] { const char *e = getenv("MGP_GEMM_QUARTER"); return !(e && e[0] == '0'); }();
  return on;
}

cudaError_t launch_dgemm(const GemmDesc &g, cudaStream_t st) {
  static bool attr_done[MGP_MAX_DEVICES] = {};
  if (first_use_on_device(attr_done)) {
    cudaFuncSetAttribute(k_dgemm<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
    cudaFuncSetAttribute(k_dgemm<false, false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
    cudaFuncSetAttribute(k_dgemm<false, false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
    cudaFuncSetAttribute(k_dgemm<false, true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
    cudaFuncSetAttribute(k_dgemm<false, true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
    cudaFuncSetAttribute(k_dgemm<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
    cudaFuncSetAttribute(k_dgemm<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
    cudaFuncSetAttribute(k_dgemm<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
  }
  int dev = 0, sm_count = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  if (g.M <= 0 || g.N <= 0) return cudaSuccess;
  int *sched = sched_words(st);
  if (!sched) return cudaErrorMemoryAllocation;
  int total = (g.N / 128) * (g.M / 128) * g.n_inner * g.n_outer;
  int limit = sm_count;
  if (g.cta_cap > 0 && limit > g.cta_cap) limit = g.cta_cap;
  // see k_dgemm: MT = 64, or MT = 32 when even the half tiles leave half of the SMs idle (few small
  // matrices: a 128 x 128 x 128 tile is 17 us on one SM, and these launches sit on the serial chain)
  const bool plain = !g.ta && !g.tb && g.kmode == 0;
  // ... and the two products of the triangular inversion, L_C X_A (k-major triangular B: the k range depends
  // on the column block only) and -X_B T with a single 128-wide k block: 128-row tiles with k up to 128 kb
  // are what the last block rows of the inversion wait for (n = 500, 8 matrices: 52 + 18 us -> see
  // profiles/r02_nll500_launches_summary.txt); the small tiles take no triangular shortcuts (exact zeros)
  const bool tri_small = !g.ta && g.tb && !g.lower_only && (g.kmode == 1 || (g.kmode == 2 && g.K <= 128));
  const bool quarter = (plain || tri_small) && quarter_tiles() && 4 * total <= limit;
  const bool half = (plain || tri_small) && !quarter && 2 * total <= limit;
  if (half) total *= 2;
  if (quarter) total *= 4;
  const int grid = total < limit ? total : limit;
  const int threads = GEMM_CONSUMERS + GEMM_PRODUCERS;
  if (quarter && g.tb) k_dgemm<false, true, 32><<<grid, threads, GEMM_SMEM, st>>>(g, sched);
  else if (half && g.tb) k_dgemm<false, true, 64><<<grid, threads, GEMM_SMEM, st>>>(g, sched);
  else if (quarter) k_dgemm<false, false, 32><<<grid, threads, GEMM_SMEM, st>>>(g, sched);
  else if (half) k_dgemm<false, false, 64><<<grid, threads, GEMM_SMEM, st>>>(g, sched);
  else if (!g.ta && !g.tb) k_dgemm<false, false><<<grid, threads, GEMM_SMEM, st>>>(g, sched);
  else if (!g.ta && g.tb) k_dgemm<false, true><<<grid, threads, GEMM_SMEM, st>>>(g, sched);
  else if (g.ta && !g.tb) k_dgemm<true, false><<<grid, threads, GEMM_SMEM, st>>>(g, sched);
  else k_dgemm<true, true><<<grid, threads, GEMM_SMEM, st>>>(g, sched);
  return cudaGetLastError();
}

static const int kDiagSmem = kDiagSmemDbl * sizeof(double);
static const int kDiag2Smem = kDiag2SmemDbl * sizeof(double);
// MGP_DIAG_V1=1 selects the first version of the diagonal-block kernel (kept for A/B measurements)
static int diag_version() {
  static const int v = [] { const char *e = getenv("MGP_DIAG_V1"); return (e && e[0] == '1') ? 1 : 2; }();
  return v;
}

cudaError_t launch_potrf_diag(double *A, double *Lout, int ld, int64_t sA, double *linv,
                              int64_t sInv, int kb, int *status, int batch, cudaStream_t st) {
  static bool attr_done[MGP_MAX_DEVICES] = {};
  if (first_use_on_device(attr_done)) {
    cudaFuncSetAttribute(k_potrf_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, kDiagSmem);
    cudaFuncSetAttribute(k_potrf_diag2, cudaFuncAttributeMaxDynamicSharedMemorySize, kDiag2Smem);
    cudaFuncSetAttribute(k_trtri_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, kDiagSmem);
  }
  if (diag_version() == 1) k_potrf_diag<<<batch, DIAG_THREADS, kDiagSmem, st>>>(A, Lout, ld, sA, linv, sInv, kb, status);
  else k_potrf_diag2<<<batch, DIAG_THREADS, kDiag2Smem, st>>>(A, Lout, ld, sA, linv, sInv, kb, status);
  return cudaGetLastError();
}

cudaError_t launch_trtri_diag(const double *L, int ld, int64_t sL, double *linv, int64_t sInv,
                              int NI, int batch, cudaStream_t st) {
  static bool attr_done[MGP_MAX_DEVICES] = {};
  if (first_use_on_device(attr_done)) {
    cudaFuncSetAttribute(k_potrf_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, kDiagSmem);
    cudaFuncSetAttribute(k_potrf_diag2, cudaFuncAttributeMaxDynamicSharedMemorySize, kDiag2Smem);
    cudaFuncSetAttribute(k_trtri_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, kDiagSmem);
  }
  k_trtri_diag<<<dim3(NI, batch), DIAG_THREADS, kDiagSmem, st>>>(L, ld, sL, linv, sInv);
  return cudaGetLastError();
}

#define LA_TRY(x)                      \
  do {                                 \
    cudaError_t e__ = (x);             \
    if (e__ != cudaSuccess) return e__; \
  } while (0)

int chol_panels(int npad) { return (npad / 128 + 3) / 4; }

// Block row kb of L^-1 on the side stream, as soon as diagonal block kb is factored and inverted
// (everything it needs -- L[kb, 0:kb], the rows above of L^-1, inv(L_kk) -- is final then):
//   X[kb, 0:kb] = -inv(L_kk) (L[kb, 0:kb] X[0:kb, 0:kb])
// The two products overlap the panel solve / update / next diagonal block of the factorisation on
// `st`; after the last block row `st` waits for the side stream.
static cudaError_t rowinv_block_row(const RowInv &ri, const double *L, int npad, int64_t sMat,
                                    const double *linv, int64_t sInv, int kb, int batch, cudaStream_t st,
                                    int64_t *launches, int cta_cap) {
  const int NI = npad / 128;
  LA_TRY(cudaEventRecord(ri.ev_fork, st));
  LA_TRY(cudaStreamWaitEvent(ri.side, ri.ev_fork, 0));
  if (kb == 0 && ri.clean_upper)
    for (int b = 0; b < batch; b++)
      LA_TRY(cudaMemsetAsync(ri.Minv + b * sMat, 0, sizeof(double) * (size_t)npad * npad, ri.side));
  k_scatter_diag<<<dim3(1, batch), 1024, 0, ri.side>>>(linv, sInv, ri.Minv, npad, sMat, kb);
  LA_TRY(cudaGetLastError());
  (*launches)++;
  if (kb > 0) {
    const int64_t row = (int64_t)kb * 128 * npad;
    GemmDesc g = {};
    g.cta_cap = cta_cap;
    g.A = L + row; g.lda = npad; g.ta = 0;                 // L[kb, 0:kb]       [128][128 kb]
    g.B = ri.Minv; g.ldb = npad; g.tb = 1;                 // X[0:kb, 0:kb]     k-major, lower triangular
    g.C = ri.T + row; g.ldc = npad;
    g.M = 128; g.N = kb * 128; g.K = kb * 128; g.alpha = 1.0; g.beta = 0.0; g.kmode = 1; g.order = 3;
    g.n_inner = 1; g.n_outer = batch;
    g.sA_out = g.sB_out = g.sC_out = sMat;
    LA_TRY(launch_dgemm(g, ri.side));
    GemmDesc h = {};
    h.cta_cap = cta_cap;
    h.A = linv + (int64_t)kb * 128 * 128; h.lda = 128; h.ta = 0;   // inv(L_kk), lower triangular
    h.B = ri.T + row; h.ldb = npad; h.tb = 1;
    h.C = ri.Minv + row; h.ldc = npad;
    h.M = 128; h.N = kb * 128; h.K = 128; h.alpha = -1.0; h.beta = 0.0; h.kmode = 2; h.order = 2;
    h.n_inner = 1; h.n_outer = batch;
    h.sA_out = sInv; h.sB_out = h.sC_out = sMat;
    LA_TRY(launch_dgemm(h, ri.side));
    (*launches) += 2;
  }
  if (kb == NI - 1) {
    LA_TRY(cudaEventRecord(ri.ev_join, ri.side));
    LA_TRY(cudaStreamWaitEvent(st, ri.ev_join, 0));
  }
  return cudaSuccess;
}

cudaError_t chol_blocked(double *A, double *L, int npad, int64_t sA, double *linv, int64_t sInv,
                         int *status, int batch, cudaStream_t st, int64_t *launches,
                         bool clean_upper, int panel_lo, int panel_hi, int cta_cap, const RowInv *ri) {
  // Two-level right-looking blocked Cholesky.  Tiles are 128 wide; PB tiles form a panel.  Inside a
  // panel every tile column is factored (diagonal block, TRSM through the explicit inverse of the
  // diagonal block) and applied to the remaining columns OF THE PANEL only (k = 128); the matrix
  // right of the panel then receives one update with k = PB*128, which reads and writes every
  // trailing tile PB times less often than a tile-by-tile right-looking sweep.
  const int NI = npad / 128, PB = 4;
  // Nothing on the likelihood path reads the tiles above the block diagonal of L (or of L^-1), so
  // they are only zeroed when the factor is going to be exported (fitted state): for a batch of
  // likelihood evaluations that saves writing 2 x npad^2 doubles per evaluation.
  // panels [panel_lo, panel_hi) of PB tile columns each (callers that interleave the launches of
  // several batch groups walk the panels one at a time); the whole factorisation is [0, chol_panels)
  if (clean_upper && panel_lo == 0)
    for (int b = 0; b < batch; b++)
      LA_TRY(cudaMemsetAsync(L + b * sA, 0, sizeof(double) * (size_t)npad * npad, st));
  for (int k0 = panel_lo * PB; k0 < NI && k0 < panel_hi * PB; k0 += PB) {
    const int k1 = k0 + PB < NI ? k0 + PB : NI;
    for (int kb = k0; kb < k1; kb++) {
      LA_TRY(launch_potrf_diag(A, L, npad, sA, linv, sInv, kb, status, batch, st));
      (*launches)++;
      if (ri) LA_TRY(rowinv_block_row(*ri, L, npad, sA, linv, sInv, kb, batch, st, launches, cta_cap));
      const int below = NI - 1 - kb;
      if (below == 0) break;
      const int64_t off_panel = (int64_t)(kb + 1) * 128 * npad + (int64_t)kb * 128;
      GemmDesc g = {};
      g.cta_cap = cta_cap;
      // L[kb+1:, kb] = A[kb+1:, kb] * inv(L_kk)^T
      g.A = A + off_panel; g.lda = npad; g.ta = 0;
      g.B = linv + (int64_t)kb * 128 * 128; g.ldb = 128; g.tb = 0;
      g.C = L + off_panel; g.ldc = npad;
      g.M = below * 128; g.N = 128; g.K = 128;
      g.alpha = 1.0; g.beta = 0.0; g.kmode = 0; g.lower_only = 0;
      g.n_inner = 1; g.n_outer = batch;
      g.sA_out = sA; g.sB_out = sInv; g.sC_out = sA;
      LA_TRY(launch_dgemm(g, st));
      (*launches)++;
      const int inpanel = k1 - 1 - kb;
      if (inpanel > 0) {
        // A[kb+1:, kb+1:k1] -= L[kb+1:, kb] L[kb+1:k1, kb]^T  (lower tiles)
        GemmDesc u = {};
      u.cta_cap = cta_cap;
        u.A = L + off_panel; u.lda = npad; u.ta = 0;
        u.B = L + off_panel; u.ldb = npad; u.tb = 0;
        u.C = A + (int64_t)(kb + 1) * 128 * (npad + 1); u.ldc = npad;
        u.M = below * 128; u.N = inpanel * 128; u.K = 128;
        u.alpha = -1.0; u.beta = 1.0; u.kmode = 0; u.lower_only = 1;
        u.n_inner = 1; u.n_outer = batch;
        u.sA_out = sA; u.sB_out = sA; u.sC_out = sA;
        LA_TRY(launch_dgemm(u, st));
        (*launches)++;
      }
    }
    if (k1 < NI) {
      // A[k1:, k1:] -= L[k1:, k0:k1] L[k1:, k0:k1]^T  (lower tiles), k = (k1 - k0) * 128
      const int64_t off = (int64_t)k1 * 128 * npad + (int64_t)k0 * 128;
      GemmDesc u = {};
      u.cta_cap = cta_cap;
      u.A = L + off; u.lda = npad; u.ta = 0;
      u.B = L + off; u.ldb = npad; u.tb = 0;
      u.C = A + (int64_t)k1 * 128 * (npad + 1); u.ldc = npad;
      u.M = (NI - k1) * 128; u.N = (NI - k1) * 128; u.K = (k1 - k0) * 128;
      u.alpha = -1.0; u.beta = 1.0; u.kmode = 0; u.lower_only = 1;
      u.n_inner = 1; u.n_outer = batch;
      u.sA_out = sA; u.sB_out = sA; u.sC_out = sA;
      LA_TRY(launch_dgemm(u, st));
      (*launches)++;
    }
  }
  return cudaSuccess;
}

cudaError_t trtri_blocked(const double *L, double *Minv, double *T, int npad, int64_t sMat,
                          const double *linv, int64_t sInv, int batch, cudaStream_t st,
                          int64_t *launches, bool clean_upper, int cta_cap) {
  const int NI = npad / 128;
  if (clean_upper)
    for (int b = 0; b < batch; b++)
      LA_TRY(cudaMemsetAsync(Minv + b * sMat, 0, sizeof(double) * (size_t)npad * npad, st));
  k_scatter_diag<<<dim3(NI, batch), 1024, 0, st>>>(linv, sInv, Minv, npad, sMat, 0);
  LA_TRY(cudaGetLastError());
  (*launches)++;
  for (int sb = 1; sb < NI; sb *= 2) {
    const int s = sb * 128;
    const int full = NI / (2 * sb);
    const int rem = NI - full * 2 * sb;
    for (int pass = 0; pass < 2; pass++) {
      // pass 0: the `full` complete pairs; pass 1: a trailing pair whose B part is shorter
      int pairs, p0, brows;
      if (pass == 0) { pairs = full; p0 = 0; brows = s; }
      else { if (rem <= sb) break; pairs = 1; p0 = full; brows = (rem - sb) * 128; }
      if (pairs == 0) continue;
      const int64_t base = (int64_t)p0 * 2 * s * (npad + 1);
      const int64_t pstride = (int64_t)2 * s * (npad + 1);
      GemmDesc g = {};
      g.cta_cap = cta_cap;
      g.A = L + base + (int64_t)s * npad; g.lda = npad; g.ta = 0;   // C block  [brows][s]
      g.B = Minv + base; g.ldb = npad; g.tb = 1;                   // A^-1     [s][s] k-major
      g.C = T + base + (int64_t)s * npad; g.ldc = npad;
      g.M = brows; g.N = s; g.K = s; g.alpha = 1.0; g.beta = 0.0; g.kmode = 1; g.order = 3;
      g.n_inner = pairs; g.n_outer = batch;
      g.sA_in = g.sB_in = g.sC_in = pstride;
      g.sA_out = g.sB_out = g.sC_out = sMat;
      LA_TRY(launch_dgemm(g, st));
      GemmDesc h = {};
      h.cta_cap = cta_cap;
      h.A = Minv + base + (int64_t)s * (npad + 1); h.lda = npad; h.ta = 0;  // B^-1 [brows][brows]
      h.B = T + base + (int64_t)s * npad; h.ldb = npad; h.tb = 1;          // T    [brows][s]
      h.C = Minv + base + (int64_t)s * npad; h.ldc = npad;
      h.M = brows; h.N = s; h.K = brows; h.alpha = -1.0; h.beta = 0.0; h.kmode = 2; h.order = 2;
      h.n_inner = pairs; h.n_outer = batch;
      h.sA_in = h.sB_in = h.sC_in = pstride;
      h.sA_out = h.sB_out = h.sC_out = sMat;
      LA_TRY(launch_dgemm(h, st));
      (*launches) += 2;
    }
  }
  return cudaSuccess;
}

cudaError_t lauum_blocked(const double *Minv, double *Rinv, int npad, int64_t sMat, int batch,
                          cudaStream_t st, int64_t *launches, int cta_cap) {
  GemmDesc g = {};
  g.cta_cap = cta_cap;
  g.A = Minv; g.lda = npad; g.ta = 1;
  g.B = Minv; g.ldb = npad; g.tb = 1;
  g.C = Rinv; g.ldc = npad;
  g.M = g.N = g.K = npad; g.alpha = 1.0; g.beta = 0.0; g.kmode = 4; g.lower_only = 1;
  g.n_inner = 1; g.n_outer = batch;
  g.sA_out = g.sB_out = g.sC_out = sMat;
  (*launches)++;
  return launch_dgemm(g, st);
}

cudaError_t lauum_split(const double *Minv, double *Rinv_a, double *Rinv_b, int npad, int64_t sMat, int batch,
                        cudaStream_t st, int64_t *launches, int cta_cap) {
  const int NI = npad / 128, Kh = (NI / 2) * 128;
  GemmDesc g = {};
  g.cta_cap = cta_cap;
  g.A = Minv; g.lda = npad; g.ta = 1;
  g.B = Minv; g.ldb = npad; g.tb = 1;
  g.C = Rinv_a; g.ldc = npad;
  g.M = g.N = npad; g.K = npad - Kh; g.K0 = Kh; g.koff_in = Kh;
  g.alpha = 1.0; g.beta = 0.0; g.kmode = 4; g.lower_only = 1;
  g.n_inner = 2; g.n_outer = batch;
  g.sA_in = g.sB_in = (int64_t)Kh * npad;
  g.sC_in = (int64_t)(((intptr_t)Rinv_b - (intptr_t)Rinv_a) / (intptr_t)sizeof(double));
  g.sA_out = g.sB_out = g.sC_out = sMat;
  (*launches)++;
  return launch_dgemm(g, st);
}

cudaError_t launch_trmv(const double *M, int npad, int64_t sM, const double *x, int64_t sx,
                        double *y, int64_t sy, int n, int trans, int batch, cudaStream_t st) {
  if (!trans) k_trmv_n<<<dim3(npad / 8, batch), 256, 0, st>>>(M, npad, sM, x, sx, y, sy, n);
  else k_trmv_t<<<dim3(npad / 32, batch), 1024, 0, st>>>(M, npad, sM, x, sx, y, sy, n);
  return cudaGetLastError();
}

cudaError_t launch_trmv2(const double *M, int npad, int64_t sM, const double *x1, const double *x2,
                         double *y1, double *y2, int64_t sy, int n, int batch, cudaStream_t st) {
  k_trmv_n2<<<dim3(npad / 8, batch), 256, 0, st>>>(M, npad, sM, x1, x2, y1, y2, sy, n);
  return cudaGetLastError();
}

cudaError_t launch_pack_tri(const double *Minv, int n, int npad, double *Mp, cudaStream_t st) {
  const int NI = npad / 128;
  k_pack_tri<<<dim3(16 * NI, NI), 1024, 0, st>>>(Minv, n, npad, Mp);
  return cudaGetLastError();
}
cudaError_t launch_pack_sym(const double *Rinv, int n, int npad, double *Rp, cudaStream_t st) {
  const int NI = npad / 128;
  k_pack_sym<<<dim3(npad / 8, NI), 1024, 0, st>>>(Rinv, n, npad, Rp);
  return cudaGetLastError();
}
