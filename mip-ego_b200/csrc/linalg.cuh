// Host-callable launchers of the dense FP64 building blocks (linalg.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// C(i,j) = alpha * sum_k A(i,k) B(k,j) + beta * C(i,j) on DMMA tensor cores.
//   ta == 0: A stored [M][K] (k contiguous, lda);  ta == 1: A stored [K][M].
//   tb == 0: B stored [N][K] (k contiguous, ldb);  tb == 1: B stored [K][N].
// M, N multiples of 128, K multiple of 16.  Two batch levels: z = outer * n_inner + inner.
// kmode bits: 1 -> k starts at colblk*128; 2 -> k ends at (rowblk+1)*128; 4 -> k starts at
// rowblk*128 (triangular operands).  lower_only: skip tiles strictly above the block diagonal.
struct GemmDesc {
  const double *A, *B;
  double *C;
  int lda, ldb, ldc;
  int M, N, K;
  int ta, tb;
  double alpha, beta;
  int kmode, lower_only;
  int order;   // tile hand-out order (longest k first): 0/1 rows ascending, 2 rows descending, 3 columns ascending
  int n_inner, n_outer;
  int64_t sA_in, sB_in, sC_in, sA_out, sB_out, sC_out;
  // Split of the k range over the inner batch index (0 = off): inner i works on the operands shifted by
  // sA_in / sB_in = i * koff_in rows of k, so its triangular start (kmode 4) is rowblk*128 - i*koff_in, and
  // inner 0 stops at K0 instead of K.  Used to cut M^T M into two halves of k that run side by side.
  int koff_in, K0;
  // Upper bound on the CTAs of this launch (0 = one per SM): leaves SMs to the kernels of other
  // streams when several batch groups run concurrently.  Per launch, so that two handles (or two
  // host threads) never see each other's setting.
  int cta_cap;
};
cudaError_t launch_dgemm(const GemmDesc &g, cudaStream_t st);
// Frees the tile-scheduler words kept for `st` on the current device (called when a handle destroys
// its streams).
void release_dgemm_stream(cudaStream_t st);

// In-place Cholesky of the kb-th 128x128 diagonal block of each matrix of the batch
// (A row-major, ld, batch stride sA), writing L_kk into Lout (same geometry) and L_kk^-1 into
// linv (batch stride sInv; block kb at linv + kb*128*128).  status[b] is set to the 1-based
// global pivot index on failure (first failure wins).
cudaError_t launch_potrf_diag(double *A, double *Lout, int ld, int64_t sA, double *linv,
                              int64_t sInv, int kb, int *status, int batch, cudaStream_t st);
// Inverse of every 128x128 diagonal block of a given lower-triangular L (no factorisation).
cudaError_t launch_trtri_diag(const double *L, int ld, int64_t sL, double *linv, int64_t sInv,
                              int NI, int batch, cudaStream_t st);
// Full blocked Cholesky A -> L (lower, row-major npad x npad, A is destroyed) + diag inverses.
// Only the panels [panel_lo, panel_hi) of 4 tile columns are processed (all: 0 .. chol_panels(npad)).
int chol_panels(int npad);
// Optional companion of the factorisation: L^-1 block row by block row on a second stream while the
// factorisation proceeds (instead of trtri_blocked afterwards).  For few, small matrices the likelihood
// is bound by the serial chain of short launches; this takes the triangular inversion off that chain.
struct RowInv {
  double *Minv, *T;       // L^-1 (output) and scratch, same geometry as L
  cudaStream_t side;      // stream of the inversion's launches
  cudaEvent_t ev_fork, ev_join;
  bool clean_upper;       // zero Minv first (exported factors)
};
cudaError_t chol_blocked(double *A, double *L, int npad, int64_t sA, double *linv, int64_t sInv,
                         int *status, int batch, cudaStream_t st, int64_t *launches,
                         bool clean_upper, int panel_lo, int panel_hi, int cta_cap = 0,
                         const RowInv *ri = nullptr);
// Minv = L^-1 given L and the diagonal-block inverses; T is scratch (npad x npad per matrix).
cudaError_t trtri_blocked(const double *L, double *Minv, double *T, int npad, int64_t sMat,
                          const double *linv, int64_t sInv, int batch, cudaStream_t st,
                          int64_t *launches, bool clean_upper, int cta_cap = 0);
// Rinv(lower tiles) = Minv^T Minv.
cudaError_t lauum_blocked(const double *Minv, double *Rinv, int npad, int64_t sMat, int batch,
                          cudaStream_t st, int64_t *launches, int cta_cap = 0);

// The same product as two halves of the k range side by side in ONE launch: Rinv_a + Rinv_b = Minv^T Minv
// (lower tiles).  For few small matrices M^T M is bound by its longest tile (k = n on one SM); the halves
// have k <= n / 2 + 64.  The consumer (gradient contraction) adds the two.
cudaError_t lauum_split(const double *Minv, double *Rinv_a, double *Rinv_b, int npad, int64_t sMat, int batch,
                        cudaStream_t st, int64_t *launches, int cta_cap = 0);

// y = M x (trans == 0) or y = M^T x (trans == 1), M lower-triangular row-major npad x npad.
// x == nullptr means the all-ones vector restricted to the first n entries.
cudaError_t launch_trmv(const double *M, int npad, int64_t sM, const double *x, int64_t sx,
                        double *y, int64_t sy, int n, int trans, int batch, cudaStream_t st);

// y1 = M x1 and y2 = M x2 in one pass over M (x vectors shared by the batch; x2 == nullptr: ones).
cudaError_t launch_trmv2(const double *M, int npad, int64_t sM, const double *x1, const double *x2,
                         double *y1, double *y2, int64_t sy, int n, int batch, cudaStream_t st);

// Pack row-major L^-1 into the slab layout of the posterior kernel (lower block triangle).
cudaError_t launch_pack_tri(const double *Minv, int n, int npad, double *Mp, cudaStream_t st);
// Pack a symmetric matrix given by its lower tiles into the full slab layout (all blocks).
cudaError_t launch_pack_sym(const double *Rinv, int n, int npad, double *Rp, cudaStream_t st);
