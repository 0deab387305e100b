// Launchers of the posterior / acquisition kernels (posterior.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mgp_handle.h"
#include "nll.cuh"

// Per-fit operand preparation: Bop[j][i] = -2 theta_i xc_ji, an[j] = sum_i theta_i xc_ji^2.
// Xsc[j][i] = c_i xc_ji with c_i = sqrt(theta_i) (theta_i for the absolute-exponential kernel), row
// stride grad_row_stride(dpad), zero padded: the training operand of the gradient kernel K5.
int grad_row_stride(int dpad);
cudaError_t launch_prep_model(const double *Xc, const double *theta, int npad, int dpad,
                              double *Bop, double *an, double *Xsc, int corr, cudaStream_t st);

struct RgenParams {
  const double *Xs;   // chunk base: candidate rows, row-major with leading dimension d
  int64_t mc;         // valid candidates in the chunk
  int d, dpad, NJ8, nCblk, JS;   // JS = number of 128-row training blocks (= npad / 128)
  const double *center, *theta, *Bop, *an, *gamma, *avec;
  const double *Xc;   // centred training inputs (npad x dpad): used by the L1 (absolute-exponential) variant
  // general trend with estimated coefficients (p > 1): Bhat^T r per candidate instead of a . r
  const double *Bhat; // npad x MGP_TP: L^-T (Ft C^-T); nullptr for the constant trend
  double *tpart;      // [JS][PT][ldp] partial sums of Bhat^T r (PT = 8 * ceil((4 KS + 1) / 8))
  double *rP, *meanpart, *ftpart;
  int64_t ldp;        // mc rounded up to 128
  // K2c-i8: when set (and rgen_can_slice), K1 writes the eight int8 slices of r here INSTEAD of rP
  uint8_t *RI8;
  int nks;            // K = 64 stages per candidate block (npad / 64)
};
bool rgen_can_slice(int corr, int dpad, bool general_trend, int slices);
// number of trend columns K1 accumulates for a given dpad (multiple of 8, >= d + 1)
int rgen_trend_cols(int dpad);
// K1: r tile generator, persistent over (candidate block, training block) items.
cudaError_t launch_rgen(const RgenParams &p, int corr, int sm_count, cudaStream_t st);

struct TrmmParams {
  const double *Mp;   // packed operand (triangular L^-1 or full R^-1)
  const double *rP;   // packed r
  double *qpart;      // [2 NI][ldq] partial sums of squares: row 2 I + (row parity of the warp)   (full == 0)
  double *wP;         // packed W = R^-1 r, same layout as rP        (full == 1)
  int NI, nCblk, NJ8, full;
  int64_t ldq;
  int last_tiles;     // 8-row tiles of the last row block that contain training rows (1..16)
};
// K2c: T = L^-1 r with fused column sums of squares (or W = R^-1 r), persistent grid.
cudaError_t launch_trmm(const TrmmParams &p, int sm_count, cudaStream_t st);

// Small-population path (mc <= 8): T = L^-1 r or W = R^-1 r as a matrix-vector product over the
// row-major operand; qpart gets npad / 32 partial sums per candidate (full == 0).
cudaError_t launch_small_mv(const double *M, int npad, int n, const double *rP, int mc, int full,
                            double *qpart, int64_t ldq, double *wP, cudaStream_t st);
// Fill the upper triangle of a row-major symmetric matrix from its lower triangle.
cudaError_t launch_symmetrize(double *A, int npad, cudaStream_t st);

struct EpiParams {
  const double *qpart, *meanpart, *ftpart;
  int NI, JS;
  int64_t ld, mc;
  double beta, sigma2, G;
  int ordinary;
  double *out_mean, *out_mse;   // chunk-offset pointers or nullptr
  double *out_val;              // chunk-offset pointer or nullptr
  int64_t ld_val;               // distance between parameter sets in out_val
  int kind, n_sets, minimize;
  double plugin;
  double params[MGP_MAX_SETS];
  // non-constant trend (p > 1): f(x) = [1, x]; mean = beta . f + r . gamma;
  // estimated coefficients: u = Bhat^T r - Cinv f, MSE uses |u|^2 (gpr.py:465-472)
  int p, PT, d, trend_est;
  const double *Xs;       // chunk base of the candidates (raw coordinates, ld d)
  const double *betav;    // device, p coefficients
  const double *Cinv;     // device, p x p (row stride MGP_PMAX), lower triangular
  const double *tpart;    // [JS][PT][ld]
  // top-1 fast path: per-block best (value, global index) per set, [n_sets][ld_blk]; nullable
  double *blk_val;
  int64_t *blk_idx;
  int64_t ld_blk, idx_base;
};
// K3: mean / MSE / acquisition per candidate.
cudaError_t launch_epilogue(const EpiParams &p, cudaStream_t st);

// Running top-k: merges the k best of vals[s][0..mc) (global index = idxs[s][i], or base + i when
// idxs is null) into topv/topi [n_sets][k] (descending; ties -> lower index first).
// init != 0 resets them first.
cudaError_t launch_topk_merge(const double *vals, const int64_t *idxs, int64_t ld, int64_t mc,
                              int64_t base, int n_sets, int k, double *topv, int64_t *topi, int init,
                              cudaStream_t st);

struct GradParams {
  const double *Xs;    // chunk base (row-major, ld d)
  int64_t mc, ldp;
  int n, d, dpad, NJ8;
  const double *Xc, *center, *theta, *gamma, *avec;
  const double *Xsc;   // scaled training coordinates, row stride grad_row_stride(dpad)
  const double *rP, *wP;
  const double *meanpart, *ftpart;
  int JS;
  double beta, sigma2, G, ftf;
  int ordinary;
  double *out_mean, *out_mse;         // nullable (chunk-offset)
  double *out_mean_dx, *out_mse_dx;   // nullable, row-major mc x d (chunk-offset)
  double *out_val, *out_dx;           // nullable; out_dx is [n_sets][m][d]
  int64_t ld_val;                     // = m (elements between sets in out_val)
  int kind, n_sets, minimize;
  double plugin;
  double params[MGP_MAX_SETS];
  // non-constant trend (see EpiParams)
  int p, PT, trend_est;
  const double *betav, *Cinv, *tpart, *Bhat;
  int wide;   // one CTA per candidate (small-population path)
};
// K5: gradient contraction + acquisition-with-gradient epilogue.
cudaError_t launch_grad(const GradParams &p, int corr, cudaStream_t st);
