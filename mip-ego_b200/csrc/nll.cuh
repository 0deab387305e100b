// Batched likelihood pipeline (nll.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mgp.h"

struct NllWork {
  int n, npad, dpad, B, n_par, mode, corr, ordinary, eval_grad;
  int gemm_cta_cap;       // CTA cap of the GEMM launches of this work item (0 = one per SM)
  // second stream of this work item (nullable) with a fork / join event pair: M^T M beside the vector
  // chain, and -- rowwise = 1 -- L^-1 block row by block row while the Cholesky runs (linalg.cuh: RowInv)
  // instead of the recursive inversion after the factorisation
  cudaStream_t side;
  cudaEvent_t ev_side_fork, ev_side_join;
  int rowwise;
  int have_M;             // 1 (with have_L): L^-1 is given in w.M as well; skip the triangular inversion
  int have_L;             // 1: L (B=1) is given in w.L; skip the correlation build and Cholesky
  double noise_var, beta0;
  const double *Xc;      // npad x dpad centred training inputs
  const double *y;       // npad, zero padded
  const double *params;  // device, B x n_par
  // workspaces, each B matrices of npad x npad
  double *R, *L, *M, *T;
  double *linv;          // B x NI x 128 x 128
  double *Yt, *Ft, *rho, *gamma;  // each B x npad
  double *avec;          // B x npad: L^-T Ft (restricted likelihood, ordinary kriging)
  // ---- general basis trend (p > 1 basis functions, e.g. linear_trend: p = d + 1) ----------------
  int p;                 // number of basis functions (1 = constant trend)
  const double *F;       // npad x MGP_TP basis matrix (rows >= n and columns >= p zero); p > 1, beta estimated
  const double *mu;      // npad: F beta for a GIVEN beta (simple kriging with a non-constant mean); nullable
  double *Ftm;           // B x npad x MGP_TP : L^-1 F
  double *Bhat;          // B x npad x MGP_TP : L^-T (Ft C^-T), C = chol(Ft^T Ft)  (REML gradient; fitted-state operand)
  double *tws;           // B x trend_ws_doubles(): beta | Ainv | C | scratch
  double logdetFF;       // log det(F^T F) of the raw basis (REML, gpr.py:736)
  int want_bhat;         // fitted-state run (mgp_finalize_fit / mgp_set_state): build Bhat even without a
                         // gradient and zero the unused upper tiles of L and L^-1 (they get exported)
  double *scal;          // B x nll_scal_doubles()
  double *gpart;         // B x ntiles x (dpad + 2)
  int *status;           // B
  double *out_llf;       // device, B (nullable)
  double *out_grad;      // device, B x n_par (nullable unless eval_grad)
};
size_t nll_scal_doubles();
#define MGP_TP 128          // padded number of basis columns (GEMM tile width)
#define MGP_PMAX 72         // max basis functions rounded up to 8 (d <= 64 -> p <= 65)
size_t trend_ws_doubles();  // per batch element: beta[PMAX] | Ainv[PMAX^2] | C[PMAX^2] | Cinv[PMAX^2]
enum { TWS_BETA = 0, TWS_AINV = MGP_PMAX, TWS_C = MGP_PMAX + MGP_PMAX * MGP_PMAX,
       TWS_CINV = MGP_PMAX + 2 * MGP_PMAX * MGP_PMAX };
// mu[j] = beta_0 + sum_i beta_{1+i} x_ji for the linear trend (x = Xc + center), 0 for j >= n
cudaError_t launch_trend_mu(const double *Xc, const double *center, int n, int npad, int d, int dpad,
                            const double *beta, double *mu, cudaStream_t st);
// F[j][0] = 1, F[j][1+i] = x_ji (j < n), zero elsewhere; npad x MGP_TP
cudaError_t launch_trend_basis(const double *Xc, const double *center, int n, int npad, int d, int dpad,
                               double *F, cudaStream_t st);
// A = Ft^T Ft, its Cholesky C, Ainv, beta = Ainv Ft^T Yt, rho = Yt - Ft beta (one refinement step),
// rr = rho^T rho, log det A.  One CTA per batch element; results in tws and scal (NLL_RR, NLL_LOGDETA).
cudaError_t launch_trend_solve(const double *Ftm, const double *Yt, double *rho, int n, int npad, int p,
                               double *tws, double *scal, int *status, int batch, cudaStream_t st);
// Qp = Ft C^-T (orthonormal columns), npad x MGP_TP
cudaError_t launch_trend_q(const double *Ftm, const double *tws, int n, int npad, int p, double *Qp,
                           int batch, cudaStream_t st);
// scal slots
enum { NLL_SIGMA2 = 0, NLL_LLF = 1, NLL_FTF = 2, NLL_FTY = 3, NLL_RR = 4, NLL_LOGDET = 5, NLL_S = 6,
       NLL_V = 7, NLL_BETA = 8, NLL_NV = 9, NLL_COEFA = 10, NLL_THSCALE = 11, NLL_LOGDETA = 12 };
// isotropic theta: Pfull[b] = [P[b][0] x d, P[b][1:]]  (n_short = 1 + extras values per row)
cudaError_t launch_iso_expand(const double *P, int B, int n_short, int d, double *Pfull, cudaStream_t st);
// ... and the gradient back: G[b][i] = Gfull[b][i], except the noise-share derivative of the
// concentrated noise_estim mode, which is Gfull[b][d].  The reference indexes its (n, n, d (+1))
// derivative tensor with range(n_par) (gpr.py:907-944, 760-783), so an isotropic theta receives the
// derivative w.r.t. the FIRST feature's theta and, in noisy mode with d > 1, sigma2 receives the
// derivative w.r.t. the second feature's theta: reproduced as is.
cudaError_t launch_iso_gather(const double *Gfull, int B, int n_short, int n_full, int d, int mode, double *G,
                              cudaStream_t st);
// Builds the tile rows >= bi0 (128 rows each) of the lower block triangle of R into w.R.
cudaError_t launch_build_R(const NllWork &w, int bi0, cudaStream_t st);
// number of hyper-parameters of a mode for d features (include/mgp.h)
int nll_n_par(int mode, int d);
// phase < 0 runs everything; otherwise one of nll_num_phases(w) phases (R build | one Cholesky panel
// each | the rest), so that a caller can interleave the launches of several batch groups.
int nll_num_phases(const NllWork &w);
cudaError_t nll_pipeline(const NllWork &w, cudaStream_t st, int64_t *launches, int phase = -1);
// flag[0] |= 1 if two of the n rows of Xc (npad x dpad) coincide (gpr.py:274-277).
cudaError_t launch_dup_check(const double *Xc, int n, int dpad, int *flag, cudaStream_t st);
