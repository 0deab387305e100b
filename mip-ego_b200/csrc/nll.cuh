// Batched likelihood pipeline (nll.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mgp.h"

struct NllWork {
  int n, npad, dpad, B, n_par, mode, corr, ordinary, eval_grad;
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
  double *scal;          // B x nll_scal_doubles()
  double *gpart;         // B x ntiles x (dpad + 2)
  int *status;           // B
  const struct LookAhead *la;  // optional (linalg.cuh)
  double *out_llf;       // device, B (nullable)
  double *out_grad;      // device, B x n_par (nullable unless eval_grad)
};
size_t nll_scal_doubles();
// scal slots
enum { NLL_SIGMA2 = 0, NLL_LLF = 1, NLL_FTF = 2, NLL_FTY = 3, NLL_RR = 4, NLL_LOGDET = 5, NLL_S = 6,
       NLL_V = 7, NLL_BETA = 8, NLL_NV = 9, NLL_COEFA = 10, NLL_THSCALE = 11 };
// number of hyper-parameters of a mode for d features (include/mgp.h)
int nll_n_par(int mode, int d);
cudaError_t nll_pipeline(const NllWork &w, cudaStream_t st, int64_t *launches);
// flag[0] |= 1 if two of the n rows of Xc (npad x dpad) coincide (gpr.py:274-277).
cudaError_t launch_dup_check(const double *Xc, int n, int dpad, int *flag, cudaStream_t st);
