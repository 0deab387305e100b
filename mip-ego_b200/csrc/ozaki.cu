// K2c-i8: T = L^-1 r with fused column sums of squares on the INT8 tensor pipe (tcgen05 / TMEM),
// an Ozaki-scheme evaluation of the FP64 product (opt-in: option "trmm_i8").
//
// Why: the exact-FP64 triangular product is n^2 flop per candidate on the DMMA pipe (37 TFLOP/s
// measured), which bounds C3 / C5 at the 96-97 % of that roofline K2c already reaches.  The INT8
// pipe of the same SM is ~120 x wider.  Both operands are split into S = 7 slices of 7 bits,
//   L^-1[i][j] = 2^ea_i sum_p a_p[i][j] 2^-7(p+1)   (balanced digits |a_p| <= 64, per-row exponent)
//   r[c][j]    =        sum_q b_q[c][j] 2^-7(q+1)   (unsigned digits 0..128; 0 < r <= 1 needs no exponent)
// every slice product a_p . b_q is an EXACT int32 GEMM (|sum| <= 8192 * 64 * 128 * 7 < 2^31), the
// 34 products with p + q <= 7 are accumulated by level s = p + q in eight TMEM accumulators, and the
// epilogue recombines them in FP64:  T = 2^ea_i sum_s 2^-7(s+2) G_s.  What is dropped (p + q >= 8)
// is below 2^-56 of the row scale per term; the slices themselves carry 2^-50 (both operands rounded
// in their last digit).  The whole GPU parity suite passes with this path switched on.
// The slices of r come from K1 itself where it can produce them (k_rgen<..., I8> in posterior.cu: 8
// slices, SE / Matern, constant trend, d <= 40: no FP64 r in HBM at all), else from k_oz_slice_r below.
//
// Kernel shape (one persistent CTA per SM, warp-specialised, 1-CTA MMA):
//   item = (64-row block I2 of L^-1, 128-candidate block cb);  D[c][i] in TMEM: lane = candidate,
//   column = level * 64 + row  (8 levels x 64 = all 512 columns);
//   warp 0 lane 0: producer -- per K = 64 stage ONE bulk copy of the 7 candidate slices (56 KB) and
//                  one of the 7 L^-1 slices (28 KB); both operands are pre-packed in global memory in
//                  the shared-memory image the UMMA descriptors expect (no-swizzle K-major core
//                  matrices), so no tensor map is needed; 2-stage mbarrier ring;
//   warp 1 lane 0: issues the tcgen05.mma.kind::i8 (M = 128, K = 32) of a stage: the slices of L^-1 are
//                  stacked along N in the shared-memory image, so ONE MMA with N = 256 multiplies a slice
//                  of r with four slices of L^-1 and accumulates into four consecutive levels (2 x 12
//                  MMAs per stage instead of 2 x 36 of N = 64: a third of the A-operand reads, which is
//                  what bounded the N = 64 version at 57 cycles per MMA against 32 at the INT8 peak);
//                  commits them to the ring / to the accumulator-full barrier;
//   warps 2-9:     epilogue -- a thread owns one candidate and 32 of the 64 rows: tcgen05.ld of the eight
//                  level sums, exact 64-bit integer recombination of levels 0..3 and 4..7, two FP64
//                  operations per row, square, sum: qpart[2 I2 + half][c].
// Measured mainloop rate on B200 (tools/ozaki/umma_i8_rate.cu): 3195 cycles per stage against the
// 8192 a DMMA tile needs at its peak: 2.56 x the FP64 tensor roofline.
#include <cuda_runtime.h>
#include <stdint.h>

#include "mgp_common.cuh"
#include "ozaki.cuh"

namespace {

constexpr int OZ_LV = 8;                   // levels p + q = 0 .. 7 kept (8 x 64 TMEM columns)
constexpr int OZ_M = 128;                  // candidates per tile (MMA M, TMEM lanes)
constexpr int OZ_N = 64;                   // L^-1 rows per tile (MMA N)
constexpr int OZ_BK = 64;                  // training columns per stage
constexpr int OZ_A_SLICE = OZ_M * OZ_BK;   // 8 KB
constexpr int OZ_B_SLICE = OZ_N * OZ_BK;   // 4 KB
constexpr int OZ_NST = 2;
constexpr int OZ_THREADS = 320;              // producer warp, MMA warp, 8 epilogue warps
constexpr int OZ_TMEM_COLS = 512;
// S = slices per operand: 8 (56 bits, 36 slice products per K step: the default, as accurate as the
// DMMA path) or 7 (49 bits, 34 products: 12 % faster, about 10 x the rounding of the DMMA path)
template <int S>
struct Oz {
  static constexpr int A_STAGE = S * OZ_A_SLICE;
  static constexpr int B_STAGE = S * OZ_B_SLICE;
  static constexpr int STAGE = A_STAGE + B_STAGE;       // 96 KB / 84 KB
  static constexpr int SMEM = OZ_NST * STAGE + 1024;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: a core matrix is 8 rows x 16 bytes (128 contiguous bytes); LBO = distance of
// the two 16-byte K chunks of one K = 32 MMA, SBO = distance of consecutive 8-row groups.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
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
// 16 consecutive columns of this warp's 32 lanes; the caller issues several and waits once
__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(addr));
}

// ---- per-fit: row scales of L^-1 and its int8 slices ------------------------------------------
// scale[i] = 2^ea_i with |L^-1[i][j]| / scale[i] < 1/2 for all j (1 for an all-zero / padding row)
__global__ void __launch_bounds__(256) k_oz_rowscale(const double *Minv, int n, int npad, double *scale) {
  __shared__ double red[8];
  const int i = blockIdx.x, tid = threadIdx.x;
  double m = 0.0;
  if (i < n)
    for (int j = tid; j <= i; j += 256) m = fmax(m, fabs(Minv[(int64_t)i * npad + j]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((tid & 31) == 0) red[tid >> 5] = m;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; w++) m = fmax(m, red[w]);
    int e = 0;
    if (m > 0.0) frexp(m, &e);            // m = f 2^e, f in [0.5, 1)
    scale[i] = m > 0.0 ? ldexp(1.0, e + 1) : 1.0;
  }
}

// L^-1 slices in the B-operand image: for the 64-row block I2 and the K = 64 stage ks <= I2 the S
// slices are STACKED ALONG N inside every 16-byte K chunk: byte (kk / 16) * S * 1024 + p * 1024 +
// (i / 8) * 128 + (i % 8) * 16 + kk % 16, so that one MMA with N = 64 cnt multiplies a slice of r with
// cnt consecutive slices of L^-1 and lands in cnt consecutive level accumulators.
// Stage index of (I2, ks): I2 (I2 + 1) / 2 + ks.  One CTA per stage; thread = (row i, 16-column chunk).
template <int OZ_S>
__global__ void __launch_bounds__(256) k_oz_pack_M(const double *Minv, const double *scale, int n, int npad, int8_t *MI8) {
  const int I2 = blockIdx.y, ks = blockIdx.x;
  if (ks > I2) return;
  const int tid = threadIdx.x, il = tid >> 2, kq = tid & 3;
  const int i = I2 * OZ_N + il, j0 = ks * OZ_BK + kq * 16;
  int8_t *dst = MI8 + ((int64_t)I2 * (I2 + 1) / 2 + ks) * Oz<OZ_S>::B_STAGE + kq * (OZ_S * 1024) + (il >> 3) * 128 + (il & 7) * 16;
  const double inv = 1.0 / scale[i];            // a power of two: the scaling is exact
  double x[16];
#pragma unroll
  for (int t = 0; t < 16; t++) {
    const int j = j0 + t;
    x[t] = (i < n && j < n && j <= i) ? Minv[(int64_t)i * npad + j] * inv : 0.0;      // |x| < 1/2, exact scaling
  }
#pragma unroll
  for (int p = 0; p < OZ_S; p++) {
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int t = 0; t < 16; t++) {
      const double y = x[t] * 128.0;                 // exact
      const double dg = rint(y);                     // |dg| <= 64
      x[t] = y - dg;                                 // exact, |x| <= 1/2
      w[t >> 2] |= ((uint32_t)(uint8_t)(int8_t)(int)dg) << (8 * (t & 3));
    }
    *reinterpret_cast<uint4 *>(dst + p * 1024) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ---- per chunk: r slices (A operand) from the packed FP64 r that K1 wrote -------------------------
// Stage (cb, ks): 7 slices x 8 KB contiguous; inside a slice byte (kk / 16) * 2048 + (c / 8) * 128 + (c % 8) * 16 + kk % 16.
// Unsigned digits of the fixed-point value rint(r 2^7S), r in (0, 1]: 0..127, the leading one 0..128.
template <int OZ_S>
__global__ void __launch_bounds__(256) k_oz_slice_r(const double *rP, int NJ8, int nks, uint8_t *RI8) {
  const int ks = blockIdx.x, cb = blockIdx.y, tid = threadIdx.x;
  uint8_t *stage = RI8 + ((int64_t)cb * nks + ks) * Oz<OZ_S>::A_STAGE;
#pragma unroll
  for (int rep = 0; rep < 2; rep++) {
    const int w = tid + 256 * rep, c = w >> 2, kq = w & 3;           // candidate, 16-column chunk
    const int j8 = ks * 8 + kq * 2;                                  // two 8-row groups of the packed r
    const double *src = rP + (((int64_t)cb * NJ8 + j8) * 16 + (c >> 3)) * 64 + (c & 7) * 8;
    double x[16];
#pragma unroll
    for (int t = 0; t < 8; t += 2) {
      const double2 a = *reinterpret_cast<const double2 *>(src + t);
      const double2 b = *reinterpret_cast<const double2 *>(src + 1024 + t);
      x[t] = a.x; x[t + 1] = a.y; x[8 + t] = b.x; x[8 + t + 1] = b.y;
    }
    uint8_t *dst = stage + kq * 2048 + (c >> 3) * 128 + (c & 7) * 16;
    // v = rint(r 2^7S) once (one multiply by a power of two + one conversion on the FP64 pipe), the
    // digits by shifts on the integer pipe: digit p = bits [7 (S-1-p), 7 (S-p)) of v, the leading digit
    // keeps the overflow (r = 1 -> 128).  The same fixed-point value as S - 1 floors and a final rint.
    uint32_t lo[16], hi[16];
#pragma unroll
    for (int t = 0; t < 16; t++) {
      const long long v = __double2ll_rn(x[t] * (double)(1ll << (7 * OZ_S)));
      lo[t] = (uint32_t)v; hi[t] = (uint32_t)((unsigned long long)v >> 32);
    }
#pragma unroll
    for (int p = 0; p < OZ_S; p++) {
      const int sh = 7 * (OZ_S - 1 - p);
      uint32_t wd[4] = {0, 0, 0, 0};
#pragma unroll
      for (int t = 0; t < 16; t++) {
        uint32_t dg = sh >= 32 ? (hi[t] >> (sh - 32)) : __funnelshift_r(lo[t], hi[t], sh);
        if (p > 0) dg &= 127u;
        wd[t >> 2] |= (dg & 0xffu) << (8 * (t & 3));
      }
      *reinterpret_cast<uint4 *>(dst + p * OZ_A_SLICE) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
    }
  }
}

// ---- the product -----------------------------------------------------------------------------------
// PAIR: launched as clusters of two CTAs that work on the SAME candidate block and on the row blocks
// 2j (rank 0) and 2j + 1 (rank 1).  The candidate slices of a stage (the larger operand, S x 8 KB) are
// then fetched from L2 once per pair: each CTA loads one half and multicasts it into both shared
// memories (cp.async.bulk ... .multicast::cluster; the completion bytes are signalled on the same
// barrier offset in both CTAs), which cuts the L2 -> SM traffic of a stage from 96 to 64 KB per CTA --
// Measured on B200 (C3 / C5 shapes): no gain -- 4.25e6 against 4.32e6 cand/s, the product is bound by
// the 1000 W power cap (SM clock 1.6 GHz under load), not by L2 bandwidth -- so the pair variant is
// an option (OzakiParams::pair, handle option "trmm_i8_pair"), off by default, kept under test.
// A stage buffer is rewritten by both producers, so its "empty" barrier counts the MMA commits of
// both CTAs (tcgen05.commit ... .multicast::cluster).  Both CTAs walk 2j + 2 stages; in the last one
// rank 0 has no L^-1 tile (above the diagonal) and only passes the candidate half on.
__device__ __forceinline__ void bulk_g2s_mc(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;\n" ::
          "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t *bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int OZ_S, bool PAIR>
__global__ void __launch_bounds__(OZ_THREADS, 1) k_trmm_i8(OzakiParams p) {
  constexpr int OZ_A_STAGE = Oz<OZ_S>::A_STAGE, OZ_B_STAGE = Oz<OZ_S>::B_STAGE, OZ_STAGE = Oz<OZ_S>::STAGE;
  extern __shared__ __align__(1024) uint8_t osm[];
  __shared__ uint64_t full[OZ_NST], empty[OZ_NST], acc_full, acc_empty;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = PAIR ? (int)(blockIdx.x & 1) : 0;               // cluster dims (2, 1, 1): rank = blockIdx.x % 2
  const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, nWorkers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int NR = PAIR ? p.NI2 / 2 : p.NI2;                           // row units (pairs of 64-row blocks, or blocks)
  const int nItems = NR * p.nCblk;
  if (tid == 0) {
    for (int s = 0; s < OZ_NST; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], PAIR ? 2 : 1); }
    mbar_init(&acc_full, 1);
    mbar_init(&acc_empty, 256);
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(OZ_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (PAIR) cluster_sync_all();      // the peer's barriers exist before anything is multicast into them
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base;
  // items: candidate blocks in groups of CG (their r slices stay L2 resident while every row unit
  // passes over them), row units in decreasing cost inside a group.  The list is ragged when nCblk is
  // not a multiple of CG: entries of the last group beyond its size decode to a negative row unit.
  const int CG = p.cgroup;
  auto decode = [&](int item, int &I2, int &cb, int &nst) {
    const int per = CG * NR;
    const int grp = item / per, rem = item - grp * per;
    const int gsz = min(CG, p.nCblk - grp * CG);
    const int ru = NR - 1 - rem / gsz;
    cb = grp * CG + rem % gsz;
    I2 = PAIR ? 2 * ru + rank : ru;
    nst = PAIR ? 2 * ru + 2 : ru + 1;          // stages the worker walks (rank 0 of a pair idles in the last one)
    return ru >= 0 && cb < p.nCblk;
  };

  if (warp == 0) {
    if (lane == 0) {
      int64_t x = 0;
      for (int item = worker; item < nItems; item += nWorkers) {
        int I2, cb, nst;
        if (!decode(item, I2, cb, nst)) continue;
        for (int ks = 0; ks < nst; ks++, x++) {
          const int st = (int)(x % OZ_NST);
          if (x >= OZ_NST) mbar_wait(&empty[st], (uint32_t)(((x / OZ_NST) - 1) & 1));
          const bool hasB = ks <= I2;
          mbar_expect_tx(&full[st], OZ_A_STAGE + (hasB ? OZ_B_STAGE : 0));
          const uint8_t *asrc = p.RI8 + ((int64_t)cb * p.nks + ks) * OZ_A_STAGE;
          if (PAIR) {
            constexpr int HALF = OZ_A_STAGE / 2;
            bulk_g2s_mc(osm + st * OZ_STAGE + rank * HALF, asrc + rank * HALF, HALF, &full[st], (uint16_t)3);
          } else {
            bulk_g2s(osm + st * OZ_STAGE, asrc, OZ_A_STAGE, &full[st]);
          }
          if (hasB)
            bulk_g2s(osm + st * OZ_STAGE + OZ_A_STAGE, p.MI8 + ((int64_t)I2 * (I2 + 1) / 2 + ks) * OZ_B_STAGE, OZ_B_STAGE,
                     &full[st]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D = S32, A = UINT8 (r digits), B = INT8 (L^-1 digits), both K-major, M = 128;
      // N = 64 x (number of stacked L^-1 slices, at most 4)
      const uint32_t idesc0 = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(OZ_M >> 4) << 24);
      int64_t x = 0;
      int it = 0;
      for (int item = worker; item < nItems; item += nWorkers) {
        int I2, cb, nst;
        if (!decode(item, I2, cb, nst)) continue;
        if (it > 0) mbar_wait(&acc_empty, (uint32_t)((it - 1) & 1));      // the epilogue has drained TMEM
        asm volatile("tcgen05.fence::after_thread_sync;");
        for (int ks = 0; ks < nst; ks++, x++) {
          const int st = (int)(x % OZ_NST);
          mbar_wait(&full[st], (uint32_t)((x / OZ_NST) & 1));
          asm volatile("tcgen05.fence::after_thread_sync;");
          const uint32_t a0 = smem_u32(osm + st * OZ_STAGE), b0 = a0 + OZ_A_STAGE;
          if (ks <= I2) {
#pragma unroll
            for (int k = 0; k < OZ_BK / 32; k++) {
              const bool first = ks == 0 && k == 0;   // the first product of a level in an item overwrites its accumulator
#pragma unroll
              for (int pa = 0; pa < OZ_S; pa++) {     // slice of r (A operand) against the slices pb = 0 .. nsl - 1 of L^-1
                const int nsl = OZ_S < OZ_LV - pa ? OZ_S : OZ_LV - pa;
                const uint64_t da = umma_desc(a0 + pa * OZ_A_SLICE + k * 2 * (OZ_M * 16), OZ_M * 16, 128);
                auto mma = [&](int q, int cnt, uint32_t accumulate) {
                  const uint64_t db = umma_desc(b0 + k * 2 * (OZ_S * 1024) + q * 1024, OZ_S * 1024, 128);
                  umma_i8(tmem + (uint32_t)((pa + q) * OZ_N), da, db, idesc0 | ((uint32_t)((cnt * OZ_N) >> 3) << 17), accumulate);
                };
#pragma unroll
                for (int q = 0; q < nsl; q += 4) {
                  const int cnt = nsl - q < 4 ? nsl - q : 4;
                  if (pa == 0) mma(q, cnt, first ? 0u : 1u);
                  else if (OZ_S < OZ_LV && pa == OZ_LV - OZ_S && q + cnt == nsl && first) {
                    // 7 slices: level 7 is first touched here (pa = 1, pb = 6) while levels 1..6 already hold pa = 0
                    if (cnt > 1) mma(q, cnt - 1, 1u);
                    mma(nsl - 1, 1, 0u);
                  } else mma(q, cnt, 1u);
                }
              }
            }
          }
          // frees the stage when these MMAs have read it (in a pair: in both CTAs, whose producers both write it)
          if (PAIR) umma_commit_mc(&empty[st], (uint16_t)3);
          else umma_commit(&empty[st]);
        }
        umma_commit(&acc_full);                     // all levels of this item are complete
        it++;
      }
    }
  } else {
    // epilogue warps 2..9: TMEM lane quarter = warp % 4 (the hardware's rule), column half = (warp - 2) / 4:
    // a thread owns one candidate and 32 of the 64 rows.  The eight level sums of a row are combined
    // EXACTLY in two 64-bit integers first (levels 0..3 and 4..7: |G_s| < 2^30, so |hi|, |lo| < 2^52 and
    // their conversion to FP64 is exact): 2 conversions and 2 FP64 operations per row instead of 8 + 8,
    // in chunks of 16 rows (8 x 16 TMEM words in registers); TMEM is released as soon as the last chunk
    // has been read.  The two halves write separate partial sums (K3 adds them in a fixed order).
    const int quarter = warp & 3, half = (warp - 2) >> 2, cl = quarter * 32 + lane;   // cl: candidate within the block
    const double w_hi = __longlong_as_double((int64_t)(1023 - 35) << 52);           // 2^-7(3+2): weight of level 3
    const double w_lo = __longlong_as_double((int64_t)(1023 - 63) << 52);           // 2^-7(7+2): weight of level 7
    int it = 0;
    for (int item = worker; item < nItems; item += nWorkers) {
      int I2, cb, nst;
      if (!decode(item, I2, cb, nst)) continue;
      mbar_wait(&acc_full, (uint32_t)(it & 1));
      asm volatile("tcgen05.fence::after_thread_sync;");
      const double *sc = p.scale + I2 * OZ_N + half * 32;
      double q = 0.0;
#pragma unroll
      for (int c = 0; c < 2; c++) {
        uint32_t v[OZ_LV][16];
#pragma unroll
        for (int s = 0; s < OZ_LV; s++)
          tmem_ld16(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s * OZ_N + half * 32 + c * 16), v[s]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c == 1) {
          asm volatile("tcgen05.fence::before_thread_sync;");
          mbar_arrive(&acc_empty);                    // TMEM may be overwritten
        }
#pragma unroll
        for (int i = 0; i < 16; i++) {
          int64_t hi = (int32_t)v[0][i], lo = (int32_t)v[4][i];
#pragma unroll
          for (int s = 1; s < 4; s++) {
            hi = hi * 128 + (int32_t)v[s][i];
            lo = lo * 128 + (int32_t)v[4 + s][i];
          }
          const double T = fma((double)hi, w_hi, (double)lo * w_lo) * sc[c * 16 + i];
          q = fma(T, T, q);
        }
      }
      p.qpart[(int64_t)(2 * I2 + half) * p.ldq + (int64_t)cb * OZ_M + cl] = q;
      it++;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (PAIR) cluster_sync_all();      // no CTA leaves while its peer may still multicast into it / arrive on its barriers
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(OZ_TMEM_COLS));
}

}  // namespace

static int oz_slices(int S) { return S == 7 ? 7 : 8; }
size_t ozaki_M_bytes(int npad, int S) {
  const int64_t NI2 = npad / OZ_N;
  return (size_t)(NI2 * (NI2 + 1) / 2) * oz_slices(S) * OZ_B_SLICE;
}
size_t ozaki_r_bytes(int npad, int64_t chunk, int S) {
  return (size_t)(chunk / OZ_M) * (npad / OZ_BK) * oz_slices(S) * OZ_A_SLICE;
}

cudaError_t launch_ozaki_pack_M(const double *Minv, int n, int npad, double *scale, int8_t *MI8, int S, cudaStream_t st) {
  k_oz_rowscale<<<npad, 256, 0, st>>>(Minv, n, npad, scale);
  const int NI2 = npad / OZ_N;
  if (oz_slices(S) == 7) k_oz_pack_M<7><<<dim3(NI2, NI2), 256, 0, st>>>(Minv, scale, n, npad, MI8);
  else k_oz_pack_M<8><<<dim3(NI2, NI2), 256, 0, st>>>(Minv, scale, n, npad, MI8);
  return cudaGetLastError();
}

cudaError_t launch_ozaki_slice_r(const double *rP, int npad, int nCblk, uint8_t *RI8, int S, cudaStream_t st) {
  const int nks = npad / OZ_BK;
  if (oz_slices(S) == 7) k_oz_slice_r<7><<<dim3(nks, nCblk), 256, 0, st>>>(rP, npad / 8, nks, RI8);
  else k_oz_slice_r<8><<<dim3(nks, nCblk), 256, 0, st>>>(rP, npad / 8, nks, RI8);
  return cudaGetLastError();
}

template <int S>
static cudaError_t trmm_i8_launch(const OzakiParams &p, int sm_count, cudaStream_t st) {
  static bool attr_done[MGP_MAX_DEVICES] = {};
  static int max_clusters[MGP_MAX_DEVICES] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(OZ_THREADS); cfg.dynamicSmemBytes = Oz<S>::SMEM; cfg.stream = st; cfg.attrs = at; cfg.numAttrs = 1;
  if (first_use_on_device(attr_done)) {
    cudaFuncSetAttribute(k_trmm_i8<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Oz<S>::SMEM);
    cudaFuncSetAttribute(k_trmm_i8<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Oz<S>::SMEM);
    // co-resident pairs (one CTA per SM; a GPC with an odd number of free SMs leaves one unused)
    cfg.gridDim = dim3(2 * (sm_count / 2));
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, k_trmm_i8<S, true>, &cfg) != cudaSuccess) { nc = 0; cudaGetLastError(); }
    if (dev >= 0 && dev < MGP_MAX_DEVICES) max_clusters[dev] = nc;
  }
  const int nc = (dev >= 0 && dev < MGP_MAX_DEVICES) ? max_clusters[dev] : 0;
  const int pair_items = (p.NI2 / 2) * p.nCblk;
  // pairs need an even number of row blocks (npad is a multiple of 128: always) and most of the machine
  if (p.pair && nc >= sm_count * 3 / 8 && p.NI2 % 2 == 0 && pair_items > 0) {
    cfg.gridDim = dim3(2 * (pair_items < nc ? pair_items : nc));
    return cudaLaunchKernelEx(&cfg, k_trmm_i8<S, true>, p);
  }
  const int items = p.NI2 * p.nCblk;
  if (items <= 0) return cudaSuccess;
  k_trmm_i8<S, false><<<items < sm_count ? items : sm_count, OZ_THREADS, Oz<S>::SMEM, st>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_trmm_i8(const OzakiParams &p, int S, int sm_count, cudaStream_t st) {
  return oz_slices(S) == 7 ? trmm_i8_launch<7>(p, sm_count, st) : trmm_i8_launch<8>(p, sm_count, st);
}
