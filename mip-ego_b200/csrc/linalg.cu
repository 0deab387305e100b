// Dense FP64 building blocks on DMMA tensor cores (sm_100a): tiled GEMM with triangular
// k-ranges, 128x128 diagonal-block Cholesky / triangular inverse, blocked Cholesky, recursive
// triangular inverse, M^T M, triangular mat-vec, and the slab packing used by the posterior
// kernels.  These implement the reference's scipy.linalg calls on the hot path:
//   linalg.cholesky          gpr.py:681      -> chol_blocked
//   solve_triangular(L, .)   gpr.py:463,685,690,694 -> explicit L^-1 (trtri_blocked) + trmv / TRMM
//   cho_solve((L,True), I)   gpr.py:899      -> lauum_blocked (R^-1 = L^-T L^-1)
#include "linalg.cuh"
#include "mgp_common.cuh"

namespace {

constexpr int GK = 16;        // k extent of one pipeline stage
constexpr int GSTAGES = 3;
constexpr int LDK = 20;       // row stride of a k-contiguous tile  [128][16] (conflict-free frags)
constexpr int LDM = 132;      // row stride of a k-major tile       [16][128]
constexpr int TILE_DBL = 2560;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) k_dgemm(GemmDesc g) {
  extern __shared__ __align__(16) double sm[];
  double *As = sm;
  double *Bs = sm + GSTAGES * TILE_DBL;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wr = warp >> 2, wc = warp & 3;
  const int rowblk = blockIdx.y, colblk = blockIdx.x;
  if (g.lower_only && colblk > rowblk) return;
  const int inner = blockIdx.z % g.n_inner, outer = blockIdx.z / g.n_inner;
  const double *A = g.A + inner * g.sA_in + outer * g.sA_out;
  const double *B = g.B + inner * g.sB_in + outer * g.sB_out;
  double *C = g.C + inner * g.sC_in + outer * g.sC_out;
  const int i0 = rowblk * 128, j0 = colblk * 128;
  int klo = 0, khi = g.K;
  if (g.kmode & 1) klo = colblk * 128;
  if (g.kmode & 4) klo = max(klo, rowblk * 128);
  if (g.kmode & 2) khi = min(khi, (rowblk + 1) * 128);
  const int nk = max(0, (khi - klo) / GK);

  auto load_tile = [&](int stage, int kt) {
    const int k0 = klo + kt * GK;
    double *as = As + stage * TILE_DBL, *bs = Bs + stage * TILE_DBL;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int c = tid + 256 * u;
      if (!TA) {
        const int row = c >> 3, kc = c & 7;
        cp_async16(as + row * LDK + kc * 2, A + (int64_t)(i0 + row) * g.lda + k0 + kc * 2);
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
  };

  double acc[8][4][2];
#pragma unroll
  for (int r = 0; r < 8; r++)
#pragma unroll
    for (int c = 0; c < 4; c++) acc[r][c][0] = acc[r][c][1] = 0.0;

#pragma unroll
  for (int s = 0; s < GSTAGES - 1; s++) {
    if (s < nk) load_tile(s, s);
    cp_async_commit();
  }
  const int lr = lane >> 2, lk = lane & 3;
  for (int kt = 0; kt < nk; kt++) {
    cp_async_wait<GSTAGES - 2>();
    __syncthreads();
    if (kt + GSTAGES - 1 < nk) load_tile((kt + GSTAGES - 1) % GSTAGES, kt + GSTAGES - 1);
    cp_async_commit();
    const double *as = As + (kt % GSTAGES) * TILE_DBL, *bs = Bs + (kt % GSTAGES) * TILE_DBL;
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
      double a[8], b[4];
#pragma unroll
      for (int r = 0; r < 8; r++)
        a[r] = TA ? as[(ks * 4 + lk) * LDM + wr * 64 + r * 8 + lr]
                  : as[(wr * 64 + r * 8 + lr) * LDK + ks * 4 + lk];
#pragma unroll
      for (int c = 0; c < 4; c++)
        b[c] = TB ? bs[(ks * 4 + lk) * LDM + wc * 32 + c * 8 + lr]
                  : bs[(wc * 32 + c * 8 + lr) * LDK + ks * 4 + lk];
#pragma unroll
      for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) dmma884(acc[r][c][0], acc[r][c][1], a[r], b[c]);
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int r = 0; r < 8; r++)
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const int row = i0 + wr * 64 + r * 8 + lr, col = j0 + wc * 32 + c * 8 + 2 * lk;
      double2 *ptr = reinterpret_cast<double2 *>(C + (int64_t)row * g.ldc + col);
      double2 v = make_double2(g.alpha * acc[r][c][0], g.alpha * acc[r][c][1]);
      if (g.beta != 0.0) {
        const double2 o = *ptr;
        v.x += g.beta * o.x;
        v.y += g.beta * o.y;
      }
      *ptr = v;
    }
}

// ---- 128x128 diagonal block: Cholesky (dpotf2-style, right looking) and in-place inverse ----
constexpr int DS = 129;  // smem row stride

// 1024 threads: 8 lanes per row (i = tid / 8, part = tid % 8), partial sums combined by shuffles.
__device__ void diag_inverse_inplace(double *S, int tid) {
  // X = L^-1 in place, columns from last to first:
  //   X_jj = 1/L_jj ;  X_ij = -(sum_{k=j+1..i} X_ik L_kj) / L_jj   (i > j)
  const int i = tid >> 3, part = tid & 7;
  for (int j = 127; j >= 0; j--) {
    const double inv = 1.0 / S[j * DS + j];
    double t = 0.0;
    if (i > j)
      for (int k = j + 1 + part; k <= i; k += 8) t = fma(S[i * DS + k], S[k * DS + j], t);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    t += __shfl_xor_sync(0xffffffffu, t, 4);
    __syncthreads();
    if (part == 0) {
      if (i > j) S[i * DS + j] = -t * inv;
      if (i == j) S[j * DS + j] = inv;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(1024) k_potrf_diag(double *A, double *Lout, int ld, int64_t sA,
                                                     double *linv, int64_t sInv, int kb,
                                                     int *status) {
  extern __shared__ __align__(16) double S[];
  const int tid = threadIdx.x, b = blockIdx.x;
  const double *Ab = A + b * sA + (int64_t)kb * 128 * ld + kb * 128;
  double *Lb = Lout + b * sA + (int64_t)kb * 128 * ld + kb * 128;
  for (int idx = tid; idx < 128 * 128; idx += 1024) {
    const int r = idx >> 7, c = idx & 127;
    S[r * DS + c] = (c <= r) ? Ab[(int64_t)r * ld + c] : 0.0;
  }
  __syncthreads();
  // dpotf2-style right-looking Cholesky; the trailing update of step j is spread over 8 column
  // phases per row (i = tid % 128, phase = tid / 128) and batched 4 deep to overlap smem latency
  const int i = tid & 127, ph = tid >> 7;
  for (int j = 0; j < 128; j++) {
    const double ajj = S[j * DS + j];
    if (!(ajj > 0.0)) {
      if (tid == 0) atomicCAS(&status[b], 0, kb * 128 + j + 1);
    }
    const double dj = sqrt(ajj);
    const double rinv = 1.0 / dj;
    __syncthreads();
    if (ph == 0) {
      if (i > j) S[i * DS + j] *= rinv;
      if (i == j) S[j * DS + j] = dj;
    }
    __syncthreads();
    if (i > j) {
      const double lij = S[i * DS + j];
      int k = j + 1 + ph;
      for (; k + 24 <= i; k += 32) {
        const double c0 = S[k * DS + j], c1 = S[(k + 8) * DS + j], c2 = S[(k + 16) * DS + j],
                     c3 = S[(k + 24) * DS + j];
        const double a0 = S[i * DS + k], a1 = S[i * DS + k + 8], a2 = S[i * DS + k + 16],
                     a3 = S[i * DS + k + 24];
        S[i * DS + k] = fma(-lij, c0, a0);
        S[i * DS + k + 8] = fma(-lij, c1, a1);
        S[i * DS + k + 16] = fma(-lij, c2, a2);
        S[i * DS + k + 24] = fma(-lij, c3, a3);
      }
      for (; k <= i; k += 8) S[i * DS + k] = fma(-lij, S[k * DS + j], S[i * DS + k]);
    }
    __syncthreads();
  }
  for (int idx = tid; idx < 128 * 128; idx += 1024) {
    const int r = idx >> 7, c = idx & 127;
    Lb[(int64_t)r * ld + c] = S[r * DS + c];
  }
  __syncthreads();
  diag_inverse_inplace(S, tid);
  double *Ib = linv + b * sInv + (int64_t)kb * 128 * 128;
  for (int idx = tid; idx < 128 * 128; idx += 1024) {
    const int r = idx >> 7, c = idx & 127;
    Ib[idx] = (c <= r) ? S[r * DS + c] : 0.0;
  }
}

__global__ void __launch_bounds__(1024) k_trtri_diag(const double *L, int ld, int64_t sL,
                                                     double *linv, int64_t sInv) {
  extern __shared__ __align__(16) double S[];
  const int tid = threadIdx.x, kb = blockIdx.x, b = blockIdx.y;
  const double *Lb = L + b * sL + (int64_t)kb * 128 * ld + kb * 128;
  for (int idx = tid; idx < 128 * 128; idx += 1024) {
    const int r = idx >> 7, c = idx & 127;
    S[r * DS + c] = (c <= r) ? Lb[(int64_t)r * ld + c] : 0.0;
  }
  __syncthreads();
  diag_inverse_inplace(S, tid);
  double *Ib = linv + b * sInv + (int64_t)kb * 128 * 128;
  for (int idx = tid; idx < 128 * 128; idx += 1024) {
    const int r = idx >> 7, c = idx & 127;
    Ib[idx] = (c <= r) ? S[r * DS + c] : 0.0;
  }
}

__global__ void k_scatter_diag(const double *linv, int64_t sInv, double *M, int ld, int64_t sM) {
  const int kb = blockIdx.x, b = blockIdx.y;
  const double *src = linv + b * sInv + (int64_t)kb * 128 * 128;
  double *dst = M + b * sM + (int64_t)kb * 128 * ld + kb * 128;
  for (int idx = threadIdx.x; idx < 128 * 128; idx += blockDim.x)
    dst[(int64_t)(idx >> 7) * ld + (idx & 127)] = src[idx];
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
  const bool lower = (col >> 7) <= I;
  double v = lower ? Rinv[(int64_t)row * npad + col] : Rinv[(int64_t)col * npad + row];
  if (row >= n || col >= n) v = 0.0;
  Rp[slab * 1024 + w] = v;
}

}  // namespace

// ---- launchers ------------------------------------------------------------------------------
cudaError_t launch_dgemm(const GemmDesc &g, cudaStream_t st) {
  static bool attr_done = false;
  const int smem = 2 * GSTAGES * TILE_DBL * sizeof(double);
  if (!attr_done) {
    cudaFuncSetAttribute(k_dgemm<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_dgemm<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_dgemm<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_dgemm<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr_done = true;
  }
  if (g.M <= 0 || g.N <= 0) return cudaSuccess;
  dim3 grid(g.N / 128, g.M / 128, g.n_inner * g.n_outer);
  if (!g.ta && !g.tb) k_dgemm<false, false><<<grid, 256, smem, st>>>(g);
  else if (!g.ta && g.tb) k_dgemm<false, true><<<grid, 256, smem, st>>>(g);
  else if (g.ta && !g.tb) k_dgemm<true, false><<<grid, 256, smem, st>>>(g);
  else k_dgemm<true, true><<<grid, 256, smem, st>>>(g);
  return cudaGetLastError();
}

static const int kDiagSmem = 128 * DS * sizeof(double);

cudaError_t launch_potrf_diag(double *A, double *Lout, int ld, int64_t sA, double *linv,
                              int64_t sInv, int kb, int *status, int batch, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(k_potrf_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, kDiagSmem);
    cudaFuncSetAttribute(k_trtri_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, kDiagSmem);
    attr_done = true;
  }
  k_potrf_diag<<<batch, 1024, kDiagSmem, st>>>(A, Lout, ld, sA, linv, sInv, kb, status);
  return cudaGetLastError();
}

cudaError_t launch_trtri_diag(const double *L, int ld, int64_t sL, double *linv, int64_t sInv,
                              int NI, int batch, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(k_potrf_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, kDiagSmem);
    cudaFuncSetAttribute(k_trtri_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, kDiagSmem);
    attr_done = true;
  }
  k_trtri_diag<<<dim3(NI, batch), 1024, kDiagSmem, st>>>(L, ld, sL, linv, sInv);
  return cudaGetLastError();
}

#define LA_TRY(x)                      \
  do {                                 \
    cudaError_t e__ = (x);             \
    if (e__ != cudaSuccess) return e__; \
  } while (0)

cudaError_t chol_blocked(double *A, double *L, int npad, int64_t sA, double *linv, int64_t sInv,
                         int *status, int batch, cudaStream_t st, int64_t *launches,
                         const LookAhead *la) {
  // Right-looking blocked Cholesky, block 128.  With look-ahead the trailing update of step kb is
  // split: the next block column is updated first, then the (serial, single-CTA) factorisation
  // of the next diagonal block runs on a side stream while the rest of the update proceeds.
  const int NI = npad / 128;
  for (int b = 0; b < batch; b++)
    LA_TRY(cudaMemsetAsync(L + b * sA, 0, sizeof(double) * (size_t)npad * npad, st));
  const bool ahead = la && la->side && NI > 2;
  LA_TRY(launch_potrf_diag(A, L, npad, sA, linv, sInv, 0, status, batch, st));
  (*launches)++;
  for (int kb = 0; kb < NI - 1; kb++) {
    const int below = NI - 1 - kb;
    const int64_t off_panel = (int64_t)(kb + 1) * 128 * npad + (int64_t)kb * 128;
    GemmDesc g = {};
    // panel: L[kb+1:, kb] = A[kb+1:, kb] * inv(L_kk)^T
    g.A = A + off_panel; g.lda = npad; g.ta = 0;
    g.B = linv + (int64_t)kb * 128 * 128; g.ldb = 128; g.tb = 0;
    g.C = L + off_panel; g.ldc = npad;
    g.M = below * 128; g.N = 128; g.K = 128;
    g.alpha = 1.0; g.beta = 0.0; g.kmode = 0; g.lower_only = 0;
    g.n_inner = 1; g.n_outer = batch;
    g.sA_out = sA; g.sB_out = sInv; g.sC_out = sA;
    LA_TRY(launch_dgemm(g, st));
    (*launches)++;
    // trailing update: A[i,j] -= L[i,kb] L[j,kb]^T on the lower tiles
    GemmDesc u = {};
    u.A = L + off_panel; u.lda = npad; u.ta = 0;
    u.B = L + off_panel; u.ldb = npad; u.tb = 0;
    u.C = A + (int64_t)(kb + 1) * 128 * (npad + 1); u.ldc = npad;
    u.M = below * 128; u.N = below * 128; u.K = 128;
    u.alpha = -1.0; u.beta = 1.0; u.kmode = 0; u.lower_only = 1;
    u.n_inner = 1; u.n_outer = batch;
    u.sA_out = sA; u.sB_out = sA; u.sC_out = sA;
    if (!ahead || below < 2) {
      LA_TRY(launch_dgemm(u, st));
      LA_TRY(launch_potrf_diag(A, L, npad, sA, linv, sInv, kb + 1, status, batch, st));
      (*launches) += 2;
    } else {
      GemmDesc c1 = u;   // (i) block column kb+1 (all rows below the diagonal of step kb)
      c1.N = 128; c1.lower_only = 0;
      LA_TRY(launch_dgemm(c1, st));
      LA_TRY(cudaEventRecord(la->ev_a, st));
      LA_TRY(cudaStreamWaitEvent(la->side, la->ev_a, 0));
      LA_TRY(launch_potrf_diag(A, L, npad, sA, linv, sInv, kb + 1, status, batch, la->side));
      LA_TRY(cudaEventRecord(la->ev_b, la->side));
      GemmDesc c2 = u;   // (ii) the remaining lower tiles: rows/cols from block kb+2
      c2.A = u.A + (int64_t)128 * npad; c2.B = u.B + (int64_t)128 * npad;
      c2.C = u.C + (int64_t)128 * (npad + 1);
      c2.M = (below - 1) * 128; c2.N = (below - 1) * 128;
      LA_TRY(launch_dgemm(c2, st));
      LA_TRY(cudaStreamWaitEvent(st, la->ev_b, 0));
      (*launches) += 3;
    }
  }
  return cudaSuccess;
}

cudaError_t trtri_blocked(const double *L, double *Minv, double *T, int npad, int64_t sMat,
                          const double *linv, int64_t sInv, int batch, cudaStream_t st,
                          int64_t *launches) {
  const int NI = npad / 128;
  for (int b = 0; b < batch; b++)
    LA_TRY(cudaMemsetAsync(Minv + b * sMat, 0, sizeof(double) * (size_t)npad * npad, st));
  k_scatter_diag<<<dim3(NI, batch), 256, 0, st>>>(linv, sInv, Minv, npad, sMat);
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
      g.A = L + base + (int64_t)s * npad; g.lda = npad; g.ta = 0;   // C block  [brows][s]
      g.B = Minv + base; g.ldb = npad; g.tb = 1;                   // A^-1     [s][s] k-major
      g.C = T + base + (int64_t)s * npad; g.ldc = npad;
      g.M = brows; g.N = s; g.K = s; g.alpha = 1.0; g.beta = 0.0; g.kmode = 1;
      g.n_inner = pairs; g.n_outer = batch;
      g.sA_in = g.sB_in = g.sC_in = pstride;
      g.sA_out = g.sB_out = g.sC_out = sMat;
      LA_TRY(launch_dgemm(g, st));
      GemmDesc h = {};
      h.A = Minv + base + (int64_t)s * (npad + 1); h.lda = npad; h.ta = 0;  // B^-1 [brows][brows]
      h.B = T + base + (int64_t)s * npad; h.ldb = npad; h.tb = 1;          // T    [brows][s]
      h.C = Minv + base + (int64_t)s * npad; h.ldc = npad;
      h.M = brows; h.N = s; h.K = brows; h.alpha = -1.0; h.beta = 0.0; h.kmode = 2;
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
                          cudaStream_t st, int64_t *launches) {
  GemmDesc g = {};
  g.A = Minv; g.lda = npad; g.ta = 1;
  g.B = Minv; g.ldb = npad; g.tb = 1;
  g.C = Rinv; g.ldc = npad;
  g.M = g.N = g.K = npad; g.alpha = 1.0; g.beta = 0.0; g.kmode = 4; g.lower_only = 1;
  g.n_inner = 1; g.n_outer = batch;
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
