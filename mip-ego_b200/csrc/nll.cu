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
       SC_THSCALE, SC_N };

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

// R tile builder: lower tiles (bi >= bj); padded indices give the identity.
template <int CORR>
__global__ void __launch_bounds__(256) k_build_R(const double *Xc, int n, int npad, int dpad,
                                                 const double *params, int n_par, int mode,
                                                 double noise_var, double *R, int64_t sR) {
  extern __shared__ double sm[];
  double *xjT = sm;               // [dpad][XP]  rows of tile bi
  double *xkT = sm + dpad * XP;   // [dpad][XP]  rows of tile bj
  double *th = xkT + dpad * XP;   // [dpad]
  const int bj = blockIdx.x, bi = blockIdx.y, b = blockIdx.z;
  if (bj > bi) return;
  const int tid = threadIdx.x;
  const double *par = params + (int64_t)b * n_par;
  const int d = mode_n_theta(mode, n_par);
  for (int idx = tid; idx < 128 * dpad; idx += 256) {
    const int r = idx / dpad, i = idx % dpad;
    xjT[i * XP + r] = Xc[(int64_t)(bi * 128 + r) * dpad + i];
    xkT[i * XP + r] = Xc[(int64_t)(bj * 128 + r) * dpad + i];
  }
  for (int i = tid; i < dpad; i += 256) th[i] = i < d ? par[i] : 0.0;
  __syncthreads();
  const double scale = mode_params(mode, par, n_par, noise_var).c;
  double *Rb = R + b * sR;
  const int c = tid & 127;
  const int gk = bj * 128 + c;
  for (int r = tid >> 7; r < 128; r += 2) {
    const int gj = bi * 128 + r;
    double v;
    if (gj == gk) v = 1.0;
    else if (gj >= n || gk >= n) v = 0.0;
    else {
      double S = 0.0;
      for (int i = 0; i < dpad; i++) {
        const double df = xjT[i * XP + r] - xkT[i * XP + c];
        S = CORR == 2 ? fma(th[i], fabs(df), S) : fma(th[i] * df, df, S);
      }
      v = scale * corr_from_S<CORR>(S);
    }
    Rb[(int64_t)gj * npad + gk] = v;
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
__global__ void __launch_bounds__(256) k_nll_scalars(const double *L, int npad, int64_t sL, int n,
                                                     const double *Yt, const double *Ft,
                                                     double *rho, int64_t sv, const double *params,
                                                     int n_par, int mode, double noise_var,
                                                     int ordinary, double beta0, double *scal,
                                                     int *status, double *out_llf) {
  __shared__ double red[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  const double *yt = Yt + b * sv, *ft = Ft + b * sv;
  double *rh = rho + b * sv;
  double a = 0.0, c = 0.0, ld = 0.0;
  for (int j = tid; j < n; j += 256) {
    a = fma(ft[j], ft[j], a);
    c = fma(ft[j], yt[j], c);
    ld += log(L[b * sL + (int64_t)j * (npad + 1)]);
  }
  const double ftf = block_sum(a, red);
  const double fty = block_sum(c, red);
  const double logdet = block_sum(ld, red);
  // ordinary: rho = Yt - Q Q^T Yt with Q = Ft/|Ft| (gpr.py:691-692); simple: Yt - L^-1 (beta 1)
  const double coef = ordinary ? fty / ftf : beta0;
  double rr = 0.0;
  for (int j = tid; j < npad; j += 256) {
    const double r = j < n ? yt[j] - coef * ft[j] : 0.0;
    rh[j] = r;
    rr = fma(r, r, rr);
  }
  rr = block_sum(rr, red);
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
        llf = -0.5 * ((n - 1) * log(two_pi * v) - log((double)n) + 2.0 * logdet + log(ftf) + rr / v);
        coefa = v / ftf;     // v q q^T with q = L^-T Ft / G  (gpr.py:754-755)
      } else {
        llf = -0.5 * (n * log(two_pi * v) - 2.0 * logdet + rr / v);
      }
    } else if (est == MGP_MODE_NOISELESS) {
      const int k = ordinary ? 1 : 0;
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
    sc[SC_BETA] = ordinary ? fty / ftf : beta0;   // beta = G^-1 Q^T Yt  (gpr.py:673)
    sc[SC_NV] = nv; sc[SC_COEFA] = coefa; sc[SC_THSCALE] = thscale;
    if (out_llf) out_llf[b] = llf;
  }
}

// Gradient contraction over the lower pairs of one 128x128 tile, with
//   W_jk = gamma_j gamma_k / s + coefA a_j a_k - Rinv_jk     (a = L^-T Ft; coefA != 0 only for REML)
// acc[i] += W_jk dR0_jk/dtheta_i (j > k);  acc[DP] += 2 W_jk R0_jk (j > k);  acc[DP+1] += W_jj.
template <int CORR, int DP>
__global__ void __launch_bounds__(256) k_nll_grad(const double *Xc, int n, int npad,
                                                  const double *params, int n_par, int mode,
                                                  const double *Rinv, int64_t sR,
                                                  const double *gamma, const double *avec, int64_t sv,
                                                  const double *scal, double *gpart, int ntiles) {
  extern __shared__ double sm[];
  double *xjT = sm;
  double *xkT = sm + DP * XP;
  double *th = xkT + DP * XP;
  double *red = th + DP;  // [8][DP+2]
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // tile index -> (bi, bj), bi >= bj
  int t = blockIdx.x, bi = 0;
  while ((bi + 1) * (bi + 2) / 2 <= t) bi++;
  const int bj = t - bi * (bi + 1) / 2;
  const double *par = params + (int64_t)b * n_par;
  const int d = mode_n_theta(mode, n_par);
  for (int idx = tid; idx < 128 * DP; idx += 256) {
    const int r = idx / DP, i = idx % DP;
    xjT[i * XP + r] = Xc[(int64_t)(bi * 128 + r) * DP + i];
    xkT[i * XP + r] = Xc[(int64_t)(bj * 128 + r) * DP + i];
  }
  for (int i = tid; i < DP; i += 256) th[i] = i < d ? par[i] : 0.0;
  __syncthreads();
  const double s_div = scal[(int64_t)b * SC_N + SC_S];
  const double coefa = scal[(int64_t)b * SC_N + SC_COEFA];
  const double *Rb = Rinv + b * sR;
  const double *gm = gamma + b * sv;
  const double *av = avec + b * sv;
  double acc[DP + 2];
#pragma unroll
  for (int i = 0; i <= DP + 1; i++) acc[i] = 0.0;
  const int c = tid & 127, gk = bj * 128 + c;
  const double gk_v = gk < n ? gm[gk] : 0.0;
  const double ak_v = (coefa != 0.0 && gk < n) ? coefa * av[gk] : 0.0;
  for (int r = tid >> 7; r < 128; r += 2) {
    const int gj = bi * 128 + r;
    if (gj >= n || gk >= n || gj < gk) continue;
    const double rinv = Rb[(int64_t)gj * npad + gk];
    double W = gm[gj] * gk_v / s_div - rinv;
    if (coefa != 0.0) W = fma(av[gj], ak_v, W);
    if (gj == gk) {
      acc[DP + 1] += W;  // R0_jj = 1: sigma2 / noise-variance derivatives only
      continue;
    }
    double df2[DP];
    double S = 0.0;
#pragma unroll
    for (int i = 0; i < DP; i++) {
      const double df = xjT[i * XP + r] - xkT[i * XP + c];
      df2[i] = CORR == 2 ? fabs(df) : df * df;   // abs-exp: dR0/dtheta_i = -|df_i| R0  (gpr.py:648)
      S = fma(th[i], df2[i], S);
    }
    double R0, fac;
    if (CORR != 1) {
      R0 = exp(-S);
      fac = -R0;  // dR0/dtheta_i = -df_i^2 R0            gpr.py:634
    } else {
      const double sq = MGP_SQRT3 * sqrt(S), e = exp(-sq);
      R0 = (1.0 + sq) * e;
      fac = -1.5 * e;  // dR0/dtheta_i = -(3/2) df_i^2 e^{-sqrt3 D}   gpr.py:643
    }
    const double wf = W * fac;
#pragma unroll
    for (int i = 0; i < DP; i++) acc[i] = fma(wf, df2[i], acc[i]);
    acc[DP] = fma(2.0 * W, R0, acc[DP]);  // both (j,k) and (k,j)
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
    gpart[((int64_t)b * ntiles + t) * (DP + 2) + i] = v;
  }
}

// Assemble d llf / d par from the tile partials (fixed order over tiles):
//   theta_i : thscale * sum_{j>k} W dR0/dtheta_i     (thscale = alpha for the estimated noise share, gpr.py:917)
//   concentrated noisy  sigma2 : (0.5 / v) sum_all W R0                    (gpr.py:938-944)
//   concentrated noise_estim alpha : sum_{j>k} W R0 = -1/2 (sum Rinv o (R0 - I) - ...)   (gpr.py:927-930)
//   restricted sigma2 : (0.5 / v) sum_all W R0;  noise_var : (0.5 / v) sum_j W_jj        (gpr.py:762-783)
__global__ void k_nll_grad_final(const double *gpart, int ntiles, int dp2, int n_par, int mode,
                                 const double *scal, const int *status, double *out_grad) {
  const int b = blockIdx.x, i = threadIdx.x;
  if (i >= n_par) return;
  const int d = mode_n_theta(mode, n_par);
  const bool restricted = (mode & MGP_LIK_RESTRICTED) != 0;
  const int est = mode & 15;
  double g = 0.0;
  if (status[b] == 0 || (restricted && status[b] == -1)) {
    auto col = [&](int src) {
      double t = 0.0;
      for (int q = 0; q < ntiles; q++) t += gpart[((int64_t)b * ntiles + q) * dp2 + src];
      return t;
    };
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
  out_grad[(int64_t)b * n_par + i] = g;
}

__global__ void k_dup_check(const double *Xc, int n, int dpad, int *flag) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
  if (j >= n || j <= k) return;
  bool same = true;
  for (int i = 0; i < dpad; i++) same = same && (Xc[(int64_t)j * dpad + i] == Xc[(int64_t)k * dpad + i]);
  if (same) atomicOr(flag, 1);
}

template <int CORR>
cudaError_t grad_dispatch(int dpad, dim3 grid, size_t smem, cudaStream_t st, const double *Xc, int n,
                          int npad, const double *params, int n_par, int mode, const double *Rinv,
                          int64_t sR, const double *gamma, const double *avec, int64_t sv,
                          const double *scal, double *gpart, int ntiles) {
  switch (dpad) {
#define NG(D)                                                                                   \
  case D:                                                                                       \
    cudaFuncSetAttribute(k_nll_grad<CORR, D>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                         (int)smem);                                                            \
    k_nll_grad<CORR, D><<<grid, 256, smem, st>>>(Xc, n, npad, params, n_par, mode, Rinv, sR,    \
                                                  gamma, avec, sv, scal, gpart, ntiles);        \
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
int nll_n_par(int mode, int d) { return d + (d - mode_n_theta(mode, d)); }

cudaError_t launch_dup_check(const double *Xc, int n, int dpad, int *flag, cudaStream_t st) {
  k_dup_check<<<dim3((n + 127) / 128, n), 128, 0, st>>>(Xc, n, dpad, flag);
  return cudaGetLastError();
}

cudaError_t nll_pipeline(const NllWork &w, cudaStream_t st, int64_t *launches) {
  const int npad = w.npad, n = w.n, B = w.B, NI = npad / 128;
  const int64_t sM = (int64_t)npad * npad;
  NL_TRY(cudaMemsetAsync(w.status, 0, sizeof(int) * B, st));
  if (w.have_L) {
    NL_TRY(launch_trtri_diag(w.L, npad, sM, w.linv, (int64_t)NI * 128 * 128, NI, B, st));
    (*launches)++;
  } else {
    const size_t smem = (size_t)(2 * w.dpad * XP + w.dpad) * sizeof(double);
    dim3 grid(NI, NI, B);
    if (w.corr == MGP_CORR_SE) {
      cudaFuncSetAttribute(k_build_R<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      k_build_R<0><<<grid, 256, smem, st>>>(w.Xc, n, npad, w.dpad, w.params, w.n_par, w.mode,
                                            w.noise_var, w.R, sM);
    } else if (w.corr == MGP_CORR_MATERN32) {
      cudaFuncSetAttribute(k_build_R<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      k_build_R<1><<<grid, 256, smem, st>>>(w.Xc, n, npad, w.dpad, w.params, w.n_par, w.mode,
                                            w.noise_var, w.R, sM);
    } else if (w.corr == MGP_CORR_ABSEXP) {
      cudaFuncSetAttribute(k_build_R<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      k_build_R<2><<<grid, 256, smem, st>>>(w.Xc, n, npad, w.dpad, w.params, w.n_par, w.mode,
                                            w.noise_var, w.R, sM);
    } else {
      return cudaErrorInvalidValue;
    }
    NL_TRY(cudaGetLastError());
    (*launches)++;
  }
  if (!w.have_L)
    NL_TRY(chol_blocked(w.R, w.L, npad, sM, w.linv, (int64_t)NI * 128 * 128, w.status, B, st, launches, w.la));
  NL_TRY(trtri_blocked(w.L, w.M, w.T, npad, sM, w.linv, (int64_t)NI * 128 * 128, B, st, launches));
  double *Yt = w.Yt, *Ft = w.Ft, *rho = w.rho, *gamma = w.gamma;
  NL_TRY(launch_trmv(w.M, npad, sM, w.y, 0, Yt, npad, n, 0, B, st));
  NL_TRY(launch_trmv(w.M, npad, sM, nullptr, 0, Ft, npad, n, 0, B, st));
  k_nll_scalars<<<B, 256, 0, st>>>(w.L, npad, sM, n, Yt, Ft, rho, npad, w.params, w.n_par, w.mode,
                                   w.noise_var, w.ordinary, w.beta0, w.scal, w.status, w.out_llf);
  NL_TRY(cudaGetLastError());
  NL_TRY(launch_trmv(w.M, npad, sM, rho, npad, gamma, npad, n, 1, B, st));
  (*launches) += 4;
  if (w.eval_grad) {
    if ((w.mode & MGP_LIK_RESTRICTED) && w.ordinary) {   // a = L^-T Ft for the q q^T term of REML
      NL_TRY(launch_trmv(w.M, npad, sM, Ft, npad, w.avec, npad, n, 1, B, st));
      (*launches)++;
    }
    NL_TRY(lauum_blocked(w.M, w.T, npad, sM, B, st, launches));
    const int ntiles = NI * (NI + 1) / 2;
    const size_t smem = (size_t)(2 * w.dpad * XP + w.dpad + 8 * (w.dpad + 2)) * sizeof(double);
    dim3 grid(ntiles, B);
    if (w.corr == MGP_CORR_SE)
      NL_TRY(grad_dispatch<0>(w.dpad, grid, smem, st, w.Xc, n, npad, w.params, w.n_par, w.mode, w.T,
                              sM, gamma, w.avec, npad, w.scal, w.gpart, ntiles));
    else if (w.corr == MGP_CORR_MATERN32)
      NL_TRY(grad_dispatch<1>(w.dpad, grid, smem, st, w.Xc, n, npad, w.params, w.n_par, w.mode, w.T,
                              sM, gamma, w.avec, npad, w.scal, w.gpart, ntiles));
    else
      NL_TRY(grad_dispatch<2>(w.dpad, grid, smem, st, w.Xc, n, npad, w.params, w.n_par, w.mode, w.T,
                              sM, gamma, w.avec, npad, w.scal, w.gpart, ntiles));
    k_nll_grad_final<<<B, 128, 0, st>>>(w.gpart, ntiles, w.dpad + 2, w.n_par, w.mode, w.scal,
                                        w.status, w.out_grad);
    NL_TRY(cudaGetLastError());
    (*launches) += 2;
  }
  return cudaSuccess;
}
