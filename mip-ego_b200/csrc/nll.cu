// Batched concentrated log-likelihood + gradient (GaussianProcess.log_likelihood_concentrated,
// gpr.py:798-946) for B hyper-parameter vectors at once.
//
// Per parameter vector b (all buffers are batch-strided):
//   R_b   = correlation_matrix(theta_b)                      gpr.py:658-668  (k_build_R)
//   L_b   = chol(R_b)                                        gpr.py:681      (chol_blocked)
//   M_b   = L_b^-1                                           replaces the solve_triangular calls
//   Yt=M y, Ft=M 1, rho, sigma2-hat, llf                     gpr.py:685-697, 829-835 / 855-874
//   gamma = M^T rho ; Rinv = M^T M                           gpr.py:897-899
//   grad_i = sum_{j<k} (gamma_j gamma_k / s - Rinv_jk) dR0_jk/dtheta_i     gpr.py:905-913, 932-944
// The n x n x d tensor of corr_grad_theta (gpr.py:622-656) is never formed: dR0/dtheta is
// regenerated tile by tile from X.
#include <math.h>

#include "linalg.cuh"
#include "mgp_common.cuh"
#include "nll.cuh"

namespace {

constexpr int XP = 132;  // padded row length of the transposed X tile in shared memory

// scal layout per batch element (doubles)
enum { SC_SIGMA2 = 0, SC_LLF, SC_FTF, SC_FTY, SC_RR, SC_LOGDET, SC_S, SC_V, SC_BETA, SC_NV, SC_COEFA,
       SC_THSCALE, SC_LOGDETA, SC_N };

// trend workspace layout per batch element
enum { TW_BETA = 0, TW_AINV = MGP_PMAX, TW_C = MGP_PMAX + MGP_PMAX * MGP_PMAX,
       TW_CINV = MGP_PMAX + 2 * MGP_PMAX * MGP_PMAX, TW_N = MGP_PMAX + 3 * MGP_PMAX * MGP_PMAX };

// mode = estimation mode | MGP_LIK_RESTRICTED.  Every mode factors
//   R = c R0 + (1 - c) I,   c = 1 | sigma2 / (sigma2 + noise_var) | alpha      (gpr.py:721-724, 838-839, 861-865)
struct ModePar { double c, sigma2, nv; };
__device__ __forceinline__ ModePar mode_params(int mode, const double *par, int n_par, double noise_var) {
  ModePar m;
  const int est = mode & 15;
  if (mode & MGP_LIK_RESTRICTED) {
    m.sigma2 = est == MGP_MODE_NOISE_ESTIM ? par[n_par - 2] : par[n_par - 1];
    m.nv = est == MGP_MODE_NOISELESS ? 0.0 : (est == MGP_MODE_NOISY ? noise_var : par[n_par - 1]);
    m.c = m.sigma2 / (m.sigma2 + m.nv);
  } else if (est == MGP_MODE_NOISY) {
    m.sigma2 = par[n_par - 1]; m.nv = noise_var; m.c = m.sigma2 / (m.sigma2 + m.nv);
  } else if (est == MGP_MODE_NOISE_ESTIM) {
    m.sigma2 = 0.0; m.nv = 0.0; m.c = par[n_par - 1];
  } else {
    m.sigma2 = 0.0; m.nv = 0.0; m.c = 1.0;
  }
  return m;
}
__host__ __device__ inline int mode_n_theta(int mode, int n_par) {
  const int est = mode & 15;
  int extra = 0;
  if (mode & MGP_LIK_RESTRICTED) extra = est == MGP_MODE_NOISE_ESTIM ? 2 : 1;
  else extra = est == MGP_MODE_NOISELESS ? 0 : 1;
  return n_par - extra;
}

// Feature scaling of the staged coordinates: with x' = c_i x the weighted distance needs one
// subtraction and one FMA per feature and pair instead of three FP64 instructions,
//   SE / Matern:  c_i = sqrt(theta_i),  sum_i theta_i (x_i - y_i)^2 = sum_i (x'_i - y'_i)^2
//   abs-exp:      c_i = theta_i,        sum_i theta_i |x_i - y_i|   = sum_i |x'_i - y'_i|
// (the likelihood kernels are bound by the FP64 pipe, which they share with the DMMAs of the other
// batch groups).  A theta_i of exactly zero cannot be folded into the coordinates -- the gradient
// with respect to it would be lost -- so such a vector takes the unscaled loop (CTA-uniform branch).
template <int CORR>
__device__ __forceinline__ double feature_scale(double theta) { return CORR == 2 ? theta : sqrt(theta); }

// R tile builder: lower tiles (bi >= bj); padded indices give the identity.
template <int CORR>
__global__ void __launch_bounds__(256) k_build_R(const double *Xc, int n, int npad, int dpad,
                                                 const double *params, int n_par, int mode,
                                                 double noise_var, double *R, int64_t sR, int bi0, int nsplit) {
  extern __shared__ __align__(16) double sm[];
  double *xjT = sm;               // [dpad][XP]  rows of tile bi
  double *xkT = sm + dpad * XP;   // [dpad][XP]  rows of tile bj
  double *th = xkT + dpad * XP;   // [dpad]      theta, then the feature scales
  double *tab = th + dpad;        // exp table (mgp_common.cuh)
  __shared__ int s_anyzero;
  // nsplit CTAs share a tile (16 / nsplit of its 16 row groups each): with few small matrices one CTA per
  // tile leaves SMs idle, and this kernel is the head of the fit's serial chain
  const int bj = blockIdx.x, bi = blockIdx.y / nsplit + bi0, part = blockIdx.y % nsplit, b = blockIdx.z;
  if (bj > bi) return;
  const int tid = threadIdx.x;
  const double *par = params + (int64_t)b * n_par;
  const int d = mode_n_theta(mode, n_par);
  if (tid == 0) s_anyzero = 0;
  exp_tab_fill(tab, tid, 256);
  __syncthreads();
  for (int i = tid; i < dpad; i += 256) {
    const double t = i < d ? par[i] : 0.0;
    th[i] = t;
    if (i < d && !(t > 0.0)) s_anyzero = 1;
  }
  __syncthreads();
  const bool scaled = s_anyzero == 0;
  for (int idx = tid; idx < 128 * dpad; idx += 256) {
    const int r = idx / dpad, i = idx % dpad;
    const double c = scaled ? feature_scale<CORR>(th[i]) : 1.0;
    xjT[i * XP + r] = c * Xc[(int64_t)(bi * 128 + r) * dpad + i];
    xkT[i * XP + r] = c * Xc[(int64_t)(bj * 128 + r) * dpad + i];
  }
  __syncthreads();
  const double scale = mode_params(mode, par, n_par, noise_var).c;
  double *Rb = R + b * sR;
  const int c = tid & 127;
  const int gk = bj * 128 + c;
  // a thread owns one column and walks its 64 rows four at a time: per feature one load of the
  // column's coordinate and two 16-byte broadcast loads of four row coordinates feed four pairs
  for (int q = part * (16 / nsplit); q < (part + 1) * (16 / nsplit); q++) {
    const int r = (tid >> 7) * 64 + 4 * q;
    double S[4] = {0.0, 0.0, 0.0, 0.0};
    if (scaled) {
#pragma unroll 4
      for (int i = 0; i < dpad; i++) {
        const double xk = xkT[i * XP + c];
        const double2 a = *reinterpret_cast<const double2 *>(xjT + i * XP + r);
        const double2 bq = *reinterpret_cast<const double2 *>(xjT + i * XP + r + 2);
        const double d0 = a.x - xk, d1 = a.y - xk, d2 = bq.x - xk, d3 = bq.y - xk;
        if (CORR == 2) {
          S[0] += fabs(d0); S[1] += fabs(d1); S[2] += fabs(d2); S[3] += fabs(d3);
        } else {
          S[0] = fma(d0, d0, S[0]); S[1] = fma(d1, d1, S[1]);
          S[2] = fma(d2, d2, S[2]); S[3] = fma(d3, d3, S[3]);
        }
      }
    } else {
      for (int i = 0; i < dpad; i++) {
        const double xk = xkT[i * XP + c], t = th[i];
        const double2 a = *reinterpret_cast<const double2 *>(xjT + i * XP + r);
        const double2 bq = *reinterpret_cast<const double2 *>(xjT + i * XP + r + 2);
        const double d0 = a.x - xk, d1 = a.y - xk, d2 = bq.x - xk, d3 = bq.y - xk;
        if (CORR == 2) {
          S[0] = fma(t, fabs(d0), S[0]); S[1] = fma(t, fabs(d1), S[1]);
          S[2] = fma(t, fabs(d2), S[2]); S[3] = fma(t, fabs(d3), S[3]);
        } else {
          S[0] = fma(t * d0, d0, S[0]); S[1] = fma(t * d1, d1, S[1]);
          S[2] = fma(t * d2, d2, S[2]); S[3] = fma(t * d3, d3, S[3]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int gj = bi * 128 + r + u;
      double v;
      if (gj == gk) v = 1.0;
      else if (gj >= n || gk >= n) v = 0.0;
      else v = scale * corr_from_S_tab<CORR>(S[u], tab);
      Rb[(int64_t)gj * npad + gk] = v;
    }
  }
}

__device__ double block_sum(double v, double *red) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < nw; w++) t += red[w];  // fixed order
  return t;
}

// One CTA per batch element: Ft/Yt inner products, rho, sigma2-hat, log det, llf.
// p == 1: Ft is the vector L^-1 1 (or L^-1 mu for a given non-constant mean, then beta0 = 1);
// p > 1: rho, rr, beta and log det(Ft^T Ft) were produced by k_trend_solve (scal slots).
__global__ void __launch_bounds__(256) k_nll_scalars(const double *L, int npad, int64_t sL, int n,
                                                     const double *Yt, const double *Ft,
                                                     double *rho, int64_t sv, const double *params,
                                                     int n_par, int mode, double noise_var,
                                                     int ordinary, double beta0, int p, double logdetFF,
                                                     double *scal, int *status, double *out_llf) {
  __shared__ double red[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  const double *yt = Yt + b * sv, *ft = Ft + b * sv;
  double *rh = rho + b * sv;
  double a = 0.0, c = 0.0, ld = 0.0;
  for (int j = tid; j < n; j += 256) {
    if (p == 1) {
      a = fma(ft[j], ft[j], a);
      c = fma(ft[j], yt[j], c);
    }
    ld += log(L[b * sL + (int64_t)j * (npad + 1)]);
  }
  double ftf = block_sum(a, red);
  const double fty = block_sum(c, red);
  const double logdet = block_sum(ld, red);
  double rr = 0.0;
  if (p == 1) {
    // ordinary: rho = Yt - Q Q^T Yt with Q = Ft/|Ft| (gpr.py:691-692); simple: Yt - L^-1 (beta 1)
    const double coef = ordinary ? fty / ftf : beta0;
    for (int j = tid; j < npad; j += 256) {
      const double r = j < n ? yt[j] - coef * ft[j] : 0.0;
      rh[j] = r;
      rr = fma(r, r, rr);
    }
    rr = block_sum(rr, red);
  } else {
    rr = scal[(int64_t)b * SC_N + SC_RR];
    ftf = exp(scal[(int64_t)b * SC_N + SC_LOGDETA]);   // det(Ft^T Ft) = prod diag(G)^2
  }
  const double logdetFtF = p == 1 ? log(ftf) : scal[(int64_t)b * SC_N + SC_LOGDETA];
  const double logdetF = p == 1 ? log((double)n) : logdetFF;
  if (tid == 0) {
    const ModePar mp = mode_params(mode, params + (int64_t)b * n_par, n_par, noise_var);
    const int est = mode & 15;
    const bool restricted = (mode & MGP_LIK_RESTRICTED) != 0;
    double sigma2, llf, s, v = 0.0, nv = 0.0, coefa = 0.0, thscale = 1.0;
    const double two_pi = 6.283185307179586;
    if (restricted) {
      // gpr.py:733-743.  F = 1: det(F^T F) = n; diag(G)^2 = Ft^T Ft.  The simple-kriging branch
      // carries the reference's minus sign on the log-determinant.
      sigma2 = mp.sigma2; nv = mp.nv; v = sigma2 + nv; s = v;
      if (ordinary) {
        llf = -0.5 * ((n - p) * log(two_pi * v) - logdetF + 2.0 * logdet + logdetFtF + rr / v);
        coefa = p == 1 ? v / ftf : v;   // v q q^T with q = L^-T Q  (gpr.py:754-755); p > 1: Bhat holds L^-T Q
      } else {
        llf = -0.5 * (n * log(two_pi * v) - 2.0 * logdet + rr / v);
      }
    } else if (est == MGP_MODE_NOISELESS) {
      const int k = ordinary ? p : 0;                                   // rank(Q Q^T) = p, gpr.py:829-832
      sigma2 = rr / (double)(n - k);                                    // gpr.py:833
      llf = -0.5 * (n * log(two_pi * sigma2) + 2.0 * logdet + n);       // gpr.py:834-835
      s = sigma2;
    } else if (est == MGP_MODE_NOISE_ESTIM) {
      const double tot = rr / (double)n;                                // gpr.py:848-851
      sigma2 = mp.c * tot; nv = (1.0 - mp.c) * tot;
      llf = -0.5 * (n * log(two_pi * tot) + 2.0 * logdet + n);
      s = tot; v = tot; thscale = mp.c;
    } else {
      sigma2 = mp.sigma2; nv = mp.nv;
      v = sigma2 + nv;
      llf = -0.5 * (n * log(two_pi * v) + 2.0 * logdet + rr / v);       // gpr.py:872-874
      s = v;
    }
    int st = status[b];
    if (st == 0 && !(llf <= 0.0)) st = -1;  // llf > 0 (or NaN) rejected, gpr.py:889-891, 745-748
    if (st != 0) llf = -INFINITY;
    status[b] = st;
    double *sc = scal + (int64_t)b * SC_N;
    sc[SC_SIGMA2] = sigma2; sc[SC_LLF] = llf; sc[SC_FTF] = ftf; sc[SC_FTY] = fty;
    sc[SC_RR] = rr; sc[SC_LOGDET] = logdet; sc[SC_S] = s; sc[SC_V] = v;
    if (p == 1) sc[SC_BETA] = ordinary ? fty / ftf : beta0;   // beta = G^-1 Q^T Yt  (gpr.py:673)
    sc[SC_NV] = nv; sc[SC_COEFA] = coefa; sc[SC_THSCALE] = thscale;
    if (out_llf) out_llf[b] = llf;
  }
}

// Gradient contraction over the lower pairs of one 128x128 tile, with
//   W_jk = gamma_j gamma_k / s + coefA a_j a_k - Rinv_jk     (a = L^-T Ft; coefA != 0 only for REML)
// acc[i] += W_jk dR0_jk/dtheta_i (j > k);  acc[DP] += 2 W_jk R0_jk (j > k);  acc[DP+1] += W_jj.
template <int CORR, int DP>
__global__ void __launch_bounds__(256, 2) k_nll_grad(const double *Xc, int n, int npad,
                                                  const double *params, int n_par, int mode,
                                                  const double *Rinv, const double *Rinv2, int64_t sR,
                                                  const double *gamma, const double *avec, int64_t sv,
                                                  const double *Bhat, int p,
                                                  const double *scal, double *gpart, int ntiles, int nsplit) {
  extern __shared__ __align__(16) double sm[];
  double *xjT = sm;
  double *xkT = sm + DP * XP;
  double *th = xkT + DP * XP;
  double *red = th + DP;  // [8][DP+2]
  double *tab = red + 8 * (DP + 2);   // exp table (mgp_common.cuh)
  __shared__ int s_anyzero;
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // tile index -> (bi, bj), bi >= bj.  The 16 row groups of a tile are summed in FOUR parts, each with its
  // own row of partial sums, and nsplit = 1, 2 or 4 CTAs share a tile (4 / nsplit parts each): with few
  // small matrices one CTA per tile leaves half of the SMs idle and the kernel sits on the serial chain
  // of the fit (n = 500, 8 matrices: 80 CTAs, 47 us).  The partial sums and their order do not depend on
  // nsplit, so a row of the batch gives bitwise the same gradient whatever the batch size.
  const int part0 = (blockIdx.x % nsplit) * (4 / nsplit);
  int t = blockIdx.x / nsplit, bi = 0;
  while ((bi + 1) * (bi + 2) / 2 <= t) bi++;
  const int bj = t - bi * (bi + 1) / 2;
  const double *par = params + (int64_t)b * n_par;
  const int d = mode_n_theta(mode, n_par);
  if (tid == 0) s_anyzero = 0;
  exp_tab_fill(tab, tid, 256);
  __syncthreads();
  for (int i = tid; i < DP; i += 256) {
    const double tv = i < d ? par[i] : 0.0;
    th[i] = tv;
    if (i < d && !(tv > 0.0)) s_anyzero = 1;
  }
  __syncthreads();
  // coordinates scaled per feature (see feature_scale): the accumulators then hold theta_i times
  // the wanted sums, undone when the tile partials are written
  const bool scaled = s_anyzero == 0;
  for (int idx = tid; idx < 128 * DP; idx += 256) {
    const int r = idx / DP, i = idx % DP;
    const double c = scaled ? feature_scale<CORR>(th[i]) : 1.0;
    xjT[i * XP + r] = c * Xc[(int64_t)(bi * 128 + r) * DP + i];
    xkT[i * XP + r] = c * Xc[(int64_t)(bj * 128 + r) * DP + i];
  }
  __syncthreads();
  const double s_div = scal[(int64_t)b * SC_N + SC_S];
  const double coefa = scal[(int64_t)b * SC_N + SC_COEFA];
  const double *Rb = Rinv + b * sR;
  const double *Rb2 = Rinv2 ? Rinv2 + b * sR : nullptr;   // R^-1 given as the sum of two halves (lauum_split)
  const double *gm = gamma + b * sv;
  const double *av = avec + b * sv;
  double acc[DP + 2];
  const int c = tid & 127, gk = bj * 128 + c;
  const double gk_v = gk < n ? gm[gk] : 0.0;
  const double ak_v = (coefa != 0.0 && gk < n) ? coefa * av[gk] : 0.0;
  // four rows per thread and two passes over the features: pass 1 accumulates S for the four
  // pairs, the pair weights W * fac follow, pass 2 regenerates the squared differences and adds
  // their weighted sum to the per-feature accumulators (no per-pair array of differences)
  for (int part = part0; part < part0 + 4 / nsplit; part++) {
#pragma unroll
  for (int i = 0; i <= DP + 1; i++) acc[i] = 0.0;
  for (int q = 4 * part; q < 4 * part + 4; q++) {
    const int r = (tid >> 7) * 64 + 4 * q;
    if (bi * 128 + r + 3 < gk || bi * 128 + r >= n || gk >= n) continue;   // nothing on or below the diagonal
    // issue the global loads of the pair weights first: they complete under pass 1
    double rinv4[4], gj4[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int gj = bi * 128 + r + u;
      const bool on = gj < n && gj >= gk;
      rinv4[u] = on ? __ldg(Rb + (int64_t)gj * npad + gk) : 0.0;
      if (Rb2 && on) rinv4[u] += __ldg(Rb2 + (int64_t)gj * npad + gk);
      gj4[u] = on ? __ldg(gm + gj) : 0.0;
    }
    double S[4] = {0.0, 0.0, 0.0, 0.0};
    if (scaled) {
#pragma unroll 4
      for (int i = 0; i < DP; i++) {
        const double xk = xkT[i * XP + c];
        const double2 a = *reinterpret_cast<const double2 *>(xjT + i * XP + r);
        const double2 bq = *reinterpret_cast<const double2 *>(xjT + i * XP + r + 2);
        const double d0 = a.x - xk, d1 = a.y - xk, d2 = bq.x - xk, d3 = bq.y - xk;
        if (CORR == 2) {
          S[0] += fabs(d0); S[1] += fabs(d1); S[2] += fabs(d2); S[3] += fabs(d3);
        } else {
          S[0] = fma(d0, d0, S[0]); S[1] = fma(d1, d1, S[1]);
          S[2] = fma(d2, d2, S[2]); S[3] = fma(d3, d3, S[3]);
        }
      }
    } else {
#pragma unroll 1
      for (int i = 0; i < DP; i++) {
        const double xk = xkT[i * XP + c], tv = th[i];
        const double2 a = *reinterpret_cast<const double2 *>(xjT + i * XP + r);
        const double2 bq = *reinterpret_cast<const double2 *>(xjT + i * XP + r + 2);
        const double d0 = a.x - xk, d1 = a.y - xk, d2 = bq.x - xk, d3 = bq.y - xk;
        if (CORR == 2) {
          S[0] = fma(tv, fabs(d0), S[0]); S[1] = fma(tv, fabs(d1), S[1]);
          S[2] = fma(tv, fabs(d2), S[2]); S[3] = fma(tv, fabs(d3), S[3]);
        } else {
          S[0] = fma(tv * d0, d0, S[0]); S[1] = fma(tv * d1, d1, S[1]);
          S[2] = fma(tv * d2, d2, S[2]); S[3] = fma(tv * d3, d3, S[3]);
        }
      }
    }
    double wf[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int gj = bi * 128 + r + u;
      wf[u] = 0.0;
      if (gj >= n || gj < gk) continue;
      double W = gj4[u] * gk_v / s_div - rinv4[u];
      if (coefa != 0.0) {
        if (p == 1) {
          W = fma(av[gj], ak_v, W);
        } else {   // v (L^-T Q)(L^-T Q)^T: a p-term dot product per pair
          const double *bj_ = Bhat + ((int64_t)b * npad + gj) * MGP_TP, *bk_ = Bhat + ((int64_t)b * npad + gk) * MGP_TP;
          double tq = 0.0;
          for (int k = 0; k < p; k++) tq = fma(bj_[k], bk_[k], tq);
          W = fma(coefa, tq, W);
        }
      }
      if (gj == gk) {
        acc[DP + 1] += W;  // R0_jj = 1: sigma2 / noise-variance derivatives only
        continue;
      }
      double R0, fac;
      if (CORR != 1) {
        R0 = exp_neg_tab(S[u], tab);
        fac = -R0;  // dR0/dtheta_i = -df_i^2 R0 (gpr.py:634) | -|df_i| R0 (gpr.py:648)
      } else {
        const double sq = MGP_SQRT3 * sqrt(S[u]), e = exp_neg_tab(sq, tab);
        R0 = (1.0 + sq) * e;
        fac = -1.5 * e;  // dR0/dtheta_i = -(3/2) df_i^2 e^{-sqrt3 D}   gpr.py:643
      }
      wf[u] = W * fac;
      acc[DP] = fma(2.0 * W, R0, acc[DP]);  // both (j,k) and (k,j)
    }
#pragma unroll
    for (int i = 0; i < DP; i++) {
      const double xk = xkT[i * XP + c];
      const double2 a = *reinterpret_cast<const double2 *>(xjT + i * XP + r);
      const double2 bq = *reinterpret_cast<const double2 *>(xjT + i * XP + r + 2);
      const double d0 = a.x - xk, d1 = a.y - xk, d2 = bq.x - xk, d3 = bq.y - xk;
      double tq;
      if (CORR == 2) {
        tq = wf[0] * fabs(d0);
        tq = fma(wf[1], fabs(d1), tq); tq = fma(wf[2], fabs(d2), tq); tq = fma(wf[3], fabs(d3), tq);
      } else {
        tq = (wf[0] * d0) * d0;
        tq = fma(wf[1] * d1, d1, tq); tq = fma(wf[2] * d2, d2, tq); tq = fma(wf[3] * d3, d3, tq);
      }
      acc[i] += tq;
    }
  }
#pragma unroll
  for (int i = 0; i <= DP + 1; i++) {
    double v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp * (DP + 2) + i] = v;
  }
  __syncthreads();
  for (int i = tid; i <= DP + 1; i += 256) {
    double v = 0.0;
    for (int w = 0; w < 8; w++) v += red[w * (DP + 2) + i];
    // scaled coordinates: the feature sums carry a factor theta_i (squared differences of
    // sqrt(theta_i) x, or absolute differences of theta_i x)
    if (scaled && i < DP) v = i < d ? v / th[i] : 0.0;
    gpart[((int64_t)b * ntiles * 4 + 4 * t + part) * (DP + 2) + i] = v;
  }
  __syncthreads();   // red is reused by the next part
  }
}

// Assemble d llf / d par from the tile partials (fixed order over tiles):
//   theta_i : thscale * sum_{j>k} W dR0/dtheta_i     (thscale = alpha for the estimated noise share, gpr.py:917)
//   concentrated noisy  sigma2 : (0.5 / v) sum_all W R0                    (gpr.py:938-944)
//   concentrated noise_estim alpha : sum_{j>k} W R0 = -1/2 (sum Rinv o (R0 - I) - ...)   (gpr.py:927-930)
//   restricted sigma2 : (0.5 / v) sum_all W R0;  noise_var : (0.5 / v) sum_j W_jj        (gpr.py:762-783)
__global__ void __launch_bounds__(1024) k_nll_grad_final(const double *gpart, int ntiles, int dp2, int n_par,
                                                         int mode, const double *scal, const int *status,
                                                         double *out_grad) {
  // one warp per parameter: lane l sums the tiles l, l + 32, ... in order, the 32 partial sums are
  // combined by a fixed shuffle tree -- the result does not depend on scheduling
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d = mode_n_theta(mode, n_par);
  const bool restricted = (mode & MGP_LIK_RESTRICTED) != 0;
  const int est = mode & 15;
  const bool live = status[b] == 0 || (restricted && status[b] == -1);
  auto col = [&](int src) {
    double t = 0.0;
    for (int q = lane; q < ntiles; q += 32) t += gpart[((int64_t)b * ntiles + q) * dp2 + src];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    return t;
  };
  for (int i = warp; i < n_par; i += 32) {
    double g = 0.0;
    if (live) {
      const double v = scal[(int64_t)b * SC_N + SC_V];
      if (i < d) {
        g = col(i) * scal[(int64_t)b * SC_N + SC_THSCALE];
      } else if (!restricted && est == MGP_MODE_NOISE_ESTIM) {
        g = 0.5 * col(dp2 - 2);
      } else if (i == d) {
        g = (col(dp2 - 2) + col(dp2 - 1)) * 0.5 / v;
      } else {
        g = col(dp2 - 1) * 0.5 / v;
      }
    }
    if (lane == 0) out_grad[(int64_t)b * n_par + i] = g;
  }
}

// ---- general basis trend -------------------------------------------------------------------------
__global__ void k_trend_mu(const double *Xc, const double *center, int n, int npad, int d, int dpad,
                           const double *beta, double *mu) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= npad) return;
  double s = 0.0;
  if (j < n) {
    s = beta[0];
    for (int i = 0; i < d; i++) s = fma(beta[1 + i], Xc[(int64_t)j * dpad + i] + center[i], s);
  }
  mu[j] = s;
}
__global__ void k_trend_basis(const double *Xc, const double *center, int n, int npad, int d, int dpad,
                              double *F) {
  const int j = blockIdx.x, c = threadIdx.x;   // one row per block, MGP_TP threads
  double v = 0.0;
  if (j < n) {
    if (c == 0) v = 1.0;
    else if (c <= d) v = Xc[(int64_t)j * dpad + c - 1] + center[c - 1];
  }
  F[(int64_t)j * MGP_TP + c] = v;
}

// G = [Ft | y]^T [Ft | y] restricted to what is needed: thread groups of S lanes share one (i, j)
// pair and split the rows; fixed reduction order (shuffles over the S lanes).
__device__ void trend_gram(const double *Ft, const double *yv, int n, int p, double *Asm /*[PMAX][PMAX+1]*/,
                           double *bsm /*[PMAX]*/, int tid, int nthreads) {
  const int npairs = p * (p + 1) / 2 + p;     // lower pairs of A, then b
  int S = 32;
  while (S > 1 && npairs * S > nthreads) S >>= 1;
  const int groups = nthreads / S;
  const int g = tid / S, sl = tid % S;
  for (int pr = g; pr < (npairs + groups - 1) / groups * groups; pr += groups) {
    int i = 0, j = 0;
    bool isb = false;
    if (pr < p * (p + 1) / 2) {
      while ((i + 1) * (i + 2) / 2 <= pr) i++;
      j = pr - i * (i + 1) / 2;
    } else if (pr < npairs) {
      isb = true;
      i = pr - p * (p + 1) / 2;
    }
    double acc = 0.0;
    if (pr < npairs) {
      for (int r = sl; r < n; r += S) {
        const double a = Ft[(int64_t)r * MGP_TP + i];
        const double c = isb ? yv[r] : Ft[(int64_t)r * MGP_TP + j];
        acc = fma(a, c, acc);
      }
    }
    for (int o = S >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (sl == 0 && pr < npairs) {
      if (isb) bsm[i] = acc;
      else { Asm[i * (MGP_PMAX + 1) + j] = acc; Asm[j * (MGP_PMAX + 1) + i] = acc; }
    }
  }
}

__global__ void __launch_bounds__(1024) k_trend_solve(const double *Ftm, const double *Yt, double *rho,
                                                      int n, int npad, int p, double *tws, double *scal,
                                                      int *status) {
  extern __shared__ double sm[];
  constexpr int LD = MGP_PMAX + 1;
  double *A = sm;                    // [PMAX][LD]  Gram matrix, overwritten by its Cholesky factor C
  double *Ci = A + MGP_PMAX * LD;    // [PMAX][LD]  C^-1
  double *bv = Ci + MGP_PMAX * LD;   // [PMAX]
  double *beta = bv + MGP_PMAX;      // [PMAX]
  double *red = beta + MGP_PMAX;     // [32]
  __shared__ int bad;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double *Ft = Ftm + (int64_t)b * npad * MGP_TP;
  const double *yt = Yt + (int64_t)b * npad;
  double *rh = rho + (int64_t)b * npad;
  double *tw = tws + (int64_t)b * TW_N;
  if (tid == 0) bad = 0;
  trend_gram(Ft, yt, n, p, A, bv, tid, 1024);
  __syncthreads();
  // Cholesky A = C C^T by one warp (column by column), then C^-1 by forward substitution
  if (warp == 0) {
    for (int j = 0; j < p; j++) {
      const double djj = A[j * LD + j];
      if (!(djj > 0.0) && lane == 0) bad = 1;
      const double inv = rsqrt(djj);
      __syncwarp();
      for (int i = j + lane; i < p; i += 32) A[i * LD + j] = (i == j) ? djj * inv : A[i * LD + j] * inv;
      __syncwarp();
      for (int i = j + 1 + lane; i < p; i += 32) {
        const double lij = A[i * LD + j];
        for (int k = j + 1; k <= i; k++) A[i * LD + k] = fma(-lij, A[k * LD + j], A[i * LD + k]);
      }
      __syncwarp();
    }
    // lane c owns column c (+32, +64) of C^-1
    for (int c = lane; c < p; c += 32) {
      for (int i = 0; i < p; i++) {
        double sacc = (i == c) ? 1.0 : 0.0;
        for (int k = c; k < i; k++) sacc = fma(-A[i * LD + k], Ci[k * LD + c], sacc);
        Ci[i * LD + c] = i < c ? 0.0 : sacc / A[i * LD + i];
      }
    }
  }
  __syncthreads();
  // Ainv = C^-T C^-1 ; beta = Ainv b
  for (int e = tid; e < p * p; e += 1024) {
    const int i = e / p, j = e % p;
    double acc = 0.0;
    for (int k = max(i, j); k < p; k++) acc = fma(Ci[k * LD + i], Ci[k * LD + j], acc);
    tw[TW_AINV + i * MGP_PMAX + j] = acc;
    tw[TW_C + i * MGP_PMAX + j] = j <= i ? A[i * LD + j] : 0.0;
    tw[TW_CINV + i * MGP_PMAX + j] = j <= i ? Ci[i * LD + j] : 0.0;
  }
  __syncthreads();
  if (tid < p) {
    double acc = 0.0;
    for (int j = 0; j < p; j++) acc = fma(tw[TW_AINV + tid * MGP_PMAX + j], bv[j], acc);
    beta[tid] = acc;
  }
  __syncthreads();
  // rho = Yt - Ft beta, then one refinement step: rho -= Ft Ainv (Ft^T rho)
  for (int pass = 0; pass < 2; pass++) {
    for (int r = tid; r < npad; r += 1024) {
      double v = 0.0;
      if (r < n) {
        v = pass == 0 ? yt[r] : rh[r];
        for (int k = 0; k < p; k++) v = fma(-Ft[(int64_t)r * MGP_TP + k], pass == 0 ? beta[k] : bv[k], v);
      }
      rh[r] = v;
    }
    __syncthreads();
    if (pass == 0) {
      // c = Ft^T rho (into bv), delta = Ainv c (back into bv), beta += delta
      {
        int S = 32;
        while (S > 1 && p * S > 1024) S >>= 1;
        const int g = tid / S, sl = tid % S;
        double acc = 0.0;
        if (g < p)
          for (int r = sl; r < n; r += S) acc = fma(Ft[(int64_t)r * MGP_TP + g], rh[r], acc);
        for (int o = S >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (sl == 0 && g < p) Ci[g] = acc;      // Ci is free now: reuse row 0 as scratch
      }
      __syncthreads();
      if (tid < p) {
        double acc = 0.0;
        for (int j = 0; j < p; j++) acc = fma(tw[TW_AINV + tid * MGP_PMAX + j], Ci[j], acc);
        bv[tid] = acc;
        beta[tid] += acc;
      }
      __syncthreads();
    }
  }
  double rr = 0.0;
  for (int r = tid; r < n; r += 1024) rr = fma(rh[r], rh[r], rr);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) rr += __shfl_xor_sync(0xffffffffu, rr, o);
  if (lane == 0) red[warp] = rr;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0, ld = 0.0;
    for (int w2 = 0; w2 < 32; w2++) t += red[w2];
    for (int j = 0; j < p; j++) ld += 2.0 * log(A[j * LD + j]);
    scal[(int64_t)b * SC_N + SC_RR] = t;
    scal[(int64_t)b * SC_N + SC_LOGDETA] = ld;
    scal[(int64_t)b * SC_N + SC_BETA] = beta[0];
    if (bad) atomicCAS(&status[b], 0, npad + 1);   // rank-deficient basis: treated like a failed Cholesky
  }
  if (tid < MGP_PMAX) tw[TW_BETA + tid] = tid < p ? beta[tid] : 0.0;
}

// Qp[r][c] = sum_{k <= c} Ft[r][k] Cinv[c][k]  (Ft C^-T); Cinv recomputed from C per CTA
__global__ void __launch_bounds__(256) k_trend_q(const double *Ftm, const double *tws, int n, int npad, int p,
                                                 double *Qp) {
  extern __shared__ double sm[];
  constexpr int LD = MGP_PMAX + 1;
  double *C = sm, *Ci = sm + MGP_PMAX * LD;
  const int b = blockIdx.y, tid = threadIdx.x;
  const double *tw = tws + (int64_t)b * TW_N;
  for (int e = tid; e < p * p; e += 256) C[(e / p) * LD + e % p] = tw[TW_C + (e / p) * MGP_PMAX + e % p];
  __syncthreads();
  for (int c = tid; c < p; c += 256)
    for (int i = 0; i < p; i++) {
      double sacc = (i == c) ? 1.0 : 0.0;
      for (int k = c; k < i; k++) sacc = fma(-C[i * LD + k], Ci[k * LD + c], sacc);
      Ci[i * LD + c] = i < c ? 0.0 : sacc / C[i * LD + i];
    }
  __syncthreads();
  const int r0 = blockIdx.x * 64;
  for (int e = tid; e < 64 * MGP_TP; e += 256) {
    const int r = r0 + e / MGP_TP, c = e % MGP_TP;
    if (r >= npad) break;
    double acc = 0.0;
    if (r < n && c < p) {
      const double *fr = Ftm + ((int64_t)b * npad + r) * MGP_TP;
      for (int k = 0; k <= c; k++) acc = fma(fr[k], Ci[c * LD + k], acc);
    }
    Qp[((int64_t)b * npad + r) * MGP_TP + c] = acc;
  }
}

__global__ void k_dup_check(const double *Xc, int n, int dpad, int *flag) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
  if (j >= n || j <= k) return;
  bool same = true;
  for (int i = 0; i < dpad; i++) same = same && (Xc[(int64_t)j * dpad + i] == Xc[(int64_t)k * dpad + i]);
  if (same) atomicOr(flag, 1);
}

__global__ void k_iso_expand(const double *P, int B, int n_short, int d, double *Pfull) {
  const int n_full = d + n_short - 1;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * n_full) return;
  const int b = e / n_full, i = e % n_full;
  Pfull[e] = P[(int64_t)b * n_short + (i < d ? 0 : i - d + 1)];
}
__global__ void k_iso_gather(const double *Gfull, int B, int n_short, int n_full, int d, int mode, double *G) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * n_short) return;
  const int b = e / n_short, i = e % n_short;
  const bool share = !(mode & MGP_LIK_RESTRICTED) && (mode & 15) == MGP_MODE_NOISE_ESTIM && i == n_short - 1;
  G[e] = Gfull[(int64_t)b * n_full + (share ? d : i)];
}

template <int CORR>
cudaError_t grad_dispatch(int dpad, dim3 grid, size_t smem, cudaStream_t st, const double *Xc, int n,
                          int npad, const double *params, int n_par, int mode, const double *Rinv,
                          const double *Rinv2, int64_t sR, const double *gamma, const double *avec, int64_t sv,
                          const double *Bhat, int p, const double *scal, double *gpart, int ntiles, int nsplit) {
  switch (dpad) {
#define NG(D)                                                                                   \
  case D: {                                                                                     \
    static bool attr[MGP_MAX_DEVICES] = {};                                                     \
    if (first_use_on_device(attr))                                                              \
      cudaFuncSetAttribute(k_nll_grad<CORR, D>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                           (int)smem);                                                          \
  }                                                                                             \
    k_nll_grad<CORR, D><<<grid, 256, smem, st>>>(Xc, n, npad, params, n_par, mode, Rinv, Rinv2, sR, \
                                                  gamma, avec, sv, Bhat, p, scal, gpart, ntiles, nsplit); \
    break;
    NG(4) NG(8) NG(12) NG(16) NG(20) NG(24) NG(28) NG(32) NG(36) NG(40) NG(44) NG(48) NG(52) NG(56) NG(60) NG(64)
#undef NG
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace

#define NL_TRY(x)                       \
  do {                                  \
    cudaError_t e__ = (x);              \
    if (e__ != cudaSuccess) return e__; \
  } while (0)

size_t nll_scal_doubles() { return SC_N; }
size_t trend_ws_doubles() { return TW_N; }

cudaError_t launch_trend_mu(const double *Xc, const double *center, int n, int npad, int d, int dpad,
                            const double *beta, double *mu, cudaStream_t st) {
  k_trend_mu<<<(npad + 127) / 128, 128, 0, st>>>(Xc, center, n, npad, d, dpad, beta, mu);
  return cudaGetLastError();
}
cudaError_t launch_trend_basis(const double *Xc, const double *center, int n, int npad, int d, int dpad,
                               double *F, cudaStream_t st) {
  k_trend_basis<<<npad, MGP_TP, 0, st>>>(Xc, center, n, npad, d, dpad, F);
  return cudaGetLastError();
}
cudaError_t launch_trend_solve(const double *Ftm, const double *Yt, double *rho, int n, int npad, int p,
                               double *tws, double *scal, int *status, int batch, cudaStream_t st) {
  const int smem = (2 * MGP_PMAX * (MGP_PMAX + 1) + 2 * MGP_PMAX + 32) * (int)sizeof(double);
  static bool done[MGP_MAX_DEVICES] = {};
  if (first_use_on_device(done)) cudaFuncSetAttribute(k_trend_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_trend_solve<<<batch, 1024, smem, st>>>(Ftm, Yt, rho, n, npad, p, tws, scal, status);
  return cudaGetLastError();
}
cudaError_t launch_trend_q(const double *Ftm, const double *tws, int n, int npad, int p, double *Qp,
                           int batch, cudaStream_t st) {
  const int smem = 2 * MGP_PMAX * (MGP_PMAX + 1) * (int)sizeof(double);
  static bool done[MGP_MAX_DEVICES] = {};
  if (first_use_on_device(done)) cudaFuncSetAttribute(k_trend_q, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_trend_q<<<dim3((npad + 63) / 64, batch), 256, smem, st>>>(Ftm, tws, n, npad, p, Qp);
  return cudaGetLastError();
}
cudaError_t launch_iso_expand(const double *P, int B, int n_short, int d, double *Pfull, cudaStream_t st) {
  const int tot = B * (d + n_short - 1);
  k_iso_expand<<<(tot + 255) / 256, 256, 0, st>>>(P, B, n_short, d, Pfull);
  return cudaGetLastError();
}
cudaError_t launch_iso_gather(const double *Gfull, int B, int n_short, int n_full, int d, int mode, double *G,
                              cudaStream_t st) {
  const int tot = B * n_short;
  k_iso_gather<<<(tot + 255) / 256, 256, 0, st>>>(Gfull, B, n_short, n_full, d, mode, G);
  return cudaGetLastError();
}
int nll_n_par(int mode, int d) { return d + (d - mode_n_theta(mode, d)); }

cudaError_t launch_dup_check(const double *Xc, int n, int dpad, int *flag, cudaStream_t st) {
  k_dup_check<<<dim3((n + 127) / 128, n), 128, 0, st>>>(Xc, n, dpad, flag);
  return cudaGetLastError();
}

cudaError_t launch_build_R(const NllWork &w, int bi0, cudaStream_t st) {
  const int NI = w.npad / 128;
  const int64_t sM = (int64_t)w.npad * w.npad;
  if (bi0 < 0 || bi0 >= NI) return cudaErrorInvalidValue;
  const size_t smem = (size_t)(2 * w.dpad * XP + w.dpad + MGP_EXP_TAB_DOUBLES) * sizeof(double);
  int dev = 0, sm_count = 148, nsplit = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  while (nsplit < 4 && (int64_t)NI * (NI - bi0 + 1) / 2 * w.B * nsplit < 2 * sm_count) nsplit *= 2;
  dim3 grid(NI, (NI - bi0) * nsplit, w.B);
  static size_t smem_set[MGP_MAX_DEVICES] = {};
  dev = dev < 0 || dev >= MGP_MAX_DEVICES ? 0 : dev;
  if (smem > smem_set[dev]) {   // the opt-in only ever needs to grow
    cudaFuncSetAttribute(k_build_R<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_build_R<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_build_R<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    smem_set[dev] = smem;
  }
  if (w.corr == MGP_CORR_SE)
    k_build_R<0><<<grid, 256, smem, st>>>(w.Xc, w.n, w.npad, w.dpad, w.params, w.n_par, w.mode, w.noise_var, w.R, sM, bi0, nsplit);
  else if (w.corr == MGP_CORR_MATERN32)
    k_build_R<1><<<grid, 256, smem, st>>>(w.Xc, w.n, w.npad, w.dpad, w.params, w.n_par, w.mode, w.noise_var, w.R, sM, bi0, nsplit);
  else if (w.corr == MGP_CORR_ABSEXP)
    k_build_R<2><<<grid, 256, smem, st>>>(w.Xc, w.n, w.npad, w.dpad, w.params, w.n_par, w.mode, w.noise_var, w.R, sM, bi0, nsplit);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

int nll_num_phases(const NllWork &w) { return w.have_L ? 2 : chol_panels(w.npad) + 2; }

// phase < 0: the whole pipeline.  Otherwise phase 0 builds R (or inverts the diagonal blocks of a
// given L), phases 1 .. P factor one Cholesky panel each and the last phase does the rest; a caller
// with several batch groups on concurrent streams issues phase by phase, group by group, so that
// every group has work queued within microseconds instead of after the ~100 launches of each
// preceding group (mgp_nll_batch_dev).
cudaError_t nll_pipeline(const NllWork &w, cudaStream_t st, int64_t *launches, int phase) {
  const int npad = w.npad, n = w.n, B = w.B, NI = npad / 128;
  const int64_t sM = (int64_t)npad * npad;
  const int last = nll_num_phases(w) - 1;
  if (phase <= 0) {
  NL_TRY(cudaMemsetAsync(w.status, 0, sizeof(int) * B, st));
  if (w.have_L && w.have_M) {
    // factor and inverse both given (rank-k append): nothing to prepare
  } else if (w.have_L) {
    NL_TRY(launch_trtri_diag(w.L, npad, sM, w.linv, (int64_t)NI * 128 * 128, NI, B, st));
    (*launches)++;
  } else {
    NL_TRY(launch_build_R(w, 0, st));
    (*launches)++;
  }
  }
  if (!w.have_L) {
    const int P = chol_panels(npad);
    const int lo = phase < 0 ? 0 : phase - 1, hi = phase < 0 ? P : phase;
    RowInv ri = {w.M, w.T, w.side, w.ev_side_fork, w.ev_side_join, w.want_bhat != 0};
    if (phase < 0 || (phase >= 1 && phase <= P))
      NL_TRY(chol_blocked(w.R, w.L, npad, sM, w.linv, (int64_t)NI * 128 * 128, w.status, B, st, launches,
                          w.want_bhat != 0, lo, hi, w.gemm_cta_cap, w.rowwise ? &ri : nullptr));
  }
  if (phase >= 0 && phase != last) return cudaSuccess;
  if (!(w.have_L && w.have_M) && !(w.rowwise && !w.have_L))
    NL_TRY(trtri_blocked(w.L, w.M, w.T, npad, sM, w.linv, (int64_t)NI * 128 * 128, B, st, launches,
                         w.want_bhat != 0, w.gemm_cta_cap));
  double *Yt = w.Yt, *Ft = w.Ft, *rho = w.rho, *gamma = w.gamma;
  const bool general = w.p > 1 && w.ordinary;   // p basis functions with estimated coefficients
  // M^T M (needed by the gradient only) depends on nothing below: beside the vector chain on the side stream
  const bool side_lauum = w.side && w.eval_grad && !general;
  const bool split_lauum = side_lauum && !w.have_L && NI >= 2 && NI <= 8;
  const double *rinv2 = split_lauum ? w.R : nullptr;
  if (side_lauum) {
    NL_TRY(cudaEventRecord(w.ev_side_fork, st));
    NL_TRY(cudaStreamWaitEvent(w.side, w.ev_side_fork, 0));
    // n <= 1024: bound by the longest tile (k = n on one SM) -> two halves of k side by side, the second
    // into the R workspace (dead since the factorisation); the gradient kernel adds them
    if (split_lauum) NL_TRY(lauum_split(w.M, w.T, w.R, npad, sM, B, w.side, launches, w.gemm_cta_cap));
    else NL_TRY(lauum_blocked(w.M, w.T, npad, sM, B, w.side, launches, w.gemm_cta_cap));
    NL_TRY(cudaEventRecord(w.ev_side_join, w.side));
  }
  if (!general) {
    // Yt = L^-1 y and Ft = L^-1 1, or L^-1 (F beta) for a given non-constant mean (then rho = Yt - 1 * Ft)
    NL_TRY(launch_trmv2(w.M, npad, sM, w.y, w.mu, Yt, Ft, npad, n, B, st));
  } else {
    NL_TRY(launch_trmv(w.M, npad, sM, w.y, 0, Yt, npad, n, 0, B, st));
    // Ft = L^-1 F as a GEMM (gpr.py:690), then the normal equations of the p x p trend problem
    GemmDesc g = {};
    g.cta_cap = w.gemm_cta_cap;
    g.A = w.M; g.lda = npad; g.ta = 0;
    g.B = w.F; g.ldb = MGP_TP; g.tb = 1;
    g.C = w.Ftm; g.ldc = MGP_TP;
    g.M = npad; g.N = MGP_TP; g.K = npad; g.alpha = 1.0; g.beta = 0.0; g.kmode = 2; g.order = 2;
    g.n_inner = 1; g.n_outer = B;
    g.sA_out = sM; g.sB_out = 0; g.sC_out = (int64_t)npad * MGP_TP;
    NL_TRY(launch_dgemm(g, st));
    NL_TRY(launch_trend_solve(w.Ftm, Yt, rho, n, npad, w.p, w.tws, w.scal, w.status, B, st));
    (*launches)++;
  }
  k_nll_scalars<<<B, 256, 0, st>>>(w.L, npad, sM, n, Yt, Ft, rho, npad, w.params, w.n_par, w.mode,
                                   w.noise_var, w.ordinary, w.mu ? 1.0 : w.beta0, general ? w.p : 1,
                                   w.logdetFF, w.scal, w.status, w.out_llf);
  NL_TRY(cudaGetLastError());
  NL_TRY(launch_trmv(w.M, npad, sM, rho, npad, gamma, npad, n, 1, B, st));
  (*launches) += general ? 4 : 3;
  const bool need_bhat = general && (w.want_bhat || (w.eval_grad && (w.mode & MGP_LIK_RESTRICTED)));
  if (need_bhat) {
    // Bhat = L^-T (Ft C^-T): REML's (L^-T Q)(L^-T Q)^T term and, for the fitted model, Ft^T L^-1 r
    NL_TRY(launch_trend_q(w.Ftm, w.tws, n, npad, w.p, w.T, B, st));   // T is free until LAUUM
    GemmDesc g = {};
    g.cta_cap = w.gemm_cta_cap;
    g.A = w.M; g.lda = npad; g.ta = 1;
    g.B = w.T; g.ldb = MGP_TP; g.tb = 1;
    g.C = w.Bhat; g.ldc = MGP_TP;
    g.M = npad; g.N = MGP_TP; g.K = npad; g.alpha = 1.0; g.beta = 0.0; g.kmode = 4; g.order = 1;
    g.n_inner = 1; g.n_outer = B;
    g.sA_out = sM; g.sB_out = (int64_t)npad * MGP_TP; g.sC_out = (int64_t)npad * MGP_TP;
    NL_TRY(launch_dgemm(g, st));
    (*launches) += 2;
  }
  if (w.eval_grad) {
    if ((w.mode & MGP_LIK_RESTRICTED) && w.ordinary && !general) {   // a = L^-T Ft for the q q^T term of REML
      NL_TRY(launch_trmv(w.M, npad, sM, Ft, npad, w.avec, npad, n, 1, B, st));
      (*launches)++;
    }
    if (side_lauum) NL_TRY(cudaStreamWaitEvent(st, w.ev_side_join, 0));
    else NL_TRY(lauum_blocked(w.M, w.T, npad, sM, B, st, launches, w.gemm_cta_cap));
    const int ntiles = NI * (NI + 1) / 2;
    const size_t smem = (size_t)(2 * w.dpad * XP + w.dpad + 8 * (w.dpad + 2) + MGP_EXP_TAB_DOUBLES) * sizeof(double);
    // CTAs per tile: up to 4 while the launch still fits the machine (2 CTAs per SM); gpart holds
    // ntiles x 4 rows per matrix whatever nsplit is
    int nsplit = 1, sm_count = 148;
    {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    }
    while (nsplit < 4 && (int64_t)ntiles * B * nsplit < 2 * sm_count) nsplit *= 2;
    dim3 grid(ntiles * nsplit, B);
    if (w.corr == MGP_CORR_SE)
      NL_TRY(grad_dispatch<0>(w.dpad, grid, smem, st, w.Xc, n, npad, w.params, w.n_par, w.mode, w.T, rinv2,
                              sM, gamma, w.avec, npad, w.Bhat, w.p, w.scal, w.gpart, ntiles, nsplit));
    else if (w.corr == MGP_CORR_MATERN32)
      NL_TRY(grad_dispatch<1>(w.dpad, grid, smem, st, w.Xc, n, npad, w.params, w.n_par, w.mode, w.T, rinv2,
                              sM, gamma, w.avec, npad, w.Bhat, w.p, w.scal, w.gpart, ntiles, nsplit));
    else
      NL_TRY(grad_dispatch<2>(w.dpad, grid, smem, st, w.Xc, n, npad, w.params, w.n_par, w.mode, w.T, rinv2,
                              sM, gamma, w.avec, npad, w.Bhat, w.p, w.scal, w.gpart, ntiles, nsplit));
    k_nll_grad_final<<<B, 1024, 0, st>>>(w.gpart, ntiles * 4, w.dpad + 2, w.n_par, w.mode, w.scal,
                                        w.status, w.out_grad);
    NL_TRY(cudaGetLastError());
    (*launches) += 2;
  }
  return cudaSuccess;
}
