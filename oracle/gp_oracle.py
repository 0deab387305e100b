"""CPU oracle for the MIP-EGO Gaussian-process surrogate hot path.

TEST INFRASTRUCTURE ONLY.  This module is the *checker* for the CUDA path: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import
it.  The product (`mip-ego_b200/`) never does, and fails loudly without its CUDA library.

It restates, in vectorised float64 numpy/scipy, the algorithm of the reference
(wangronin/MIP-EGO, mipego 2.0.0).  Every function cites the reference file:line it follows
(paths relative to the reference root).  The arithmetic libraries are the reference's own
(numpy, scipy.linalg LAPACK wrappers, scipy.special.ndtr == scipy.stats.norm.cdf).

Parity pin: the reference's own tests hold no numeric vectors for this path (SURVEY.md
section 4), so the oracle is pinned against outputs of the *reference itself* run in the
build container: `oracle/gen_golden.py` imports the unmodified reference (through
`oracle/ref_shim.py`) and freezes its outputs under `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks this module against every one of them.

Batched semantics: the reference's EI / EpsilonPI / MGFI raise on more than one candidate
(InfillCriteria.py:122,171,220 -- truth value of an array); "batched" here means the
element-wise application of the reference's single-point rule.
"""
import math

import numpy as np
from scipy import linalg
from scipy.special import ndtr

SQRT3 = math.sqrt(3.0)
_NORM_PDF_C = math.sqrt(2.0 * math.pi)

CORR_TYPES = ("squared_exponential", "matern", "absolute_exponential")


# --------------------------------------------------------------------------------------
# correlation functions  (mipego/GaussianProcess/kernel.py)
# --------------------------------------------------------------------------------------
def corr(kind, theta, D):
    """r(theta, d) for component-wise distances D (P, d).

    squared_exponential: kernel.py:277-317   exp(-sum_i theta_i d_i^2)
    matern (nu=1.5):     kernel.py:147-185   (1+sqrt3 D) exp(-sqrt3 D), D=sqrt(sum theta_i d_i^2)
    absolute_exponential kernel.py:235-274   exp(-sum_i theta_i |d_i|)
    theta of size 1 is isotropic (kernel.py:312-313).
    """
    theta = np.asarray(theta, dtype=np.float64).ravel()
    D = np.asarray(D, dtype=np.float64)
    nf = D.shape[1]
    if theta.size == 1 and nf > 1:
        # isotropic: the reference multiplies the SUM over features by theta[0]
        # (kernel.py:312-313, 173-174, 267-268), which rounds differently from sum(theta_i d_i^2)
        if kind == "squared_exponential":
            return np.exp(-theta[0] * np.sum(D ** 2, axis=1))
        if kind == "matern":
            K = np.sqrt(theta[0] * np.sum(D ** 2, axis=1)) * SQRT3
            return (1.0 + K) * np.exp(-K)
        if kind == "absolute_exponential":
            return np.exp(-theta[0] * np.sum(np.abs(D), axis=1))
        raise ValueError(kind)
    if theta.size != nf:
        raise ValueError("Length of theta must be 1 or %s" % nf)
    if kind == "squared_exponential":
        return np.exp(-np.sum(theta.reshape(1, nf) * D ** 2, axis=1))
    if kind == "matern":
        dists = np.sqrt(np.sum(theta.reshape(1, nf) * D ** 2, axis=1))
        K = dists * SQRT3
        return (1.0 + K) * np.exp(-K)
    if kind == "absolute_exponential":
        return np.exp(-np.sum(theta.reshape(1, nf) * np.abs(D), axis=1))
    raise ValueError(kind)


def _full_theta(theta, d):
    theta = np.asarray(theta, dtype=np.float64).ravel()
    return np.repeat(theta, d) if theta.size == 1 else theta


def cross_corr(kind, theta, Xs, X):
    """(m, n) correlation between candidates Xs and training X (gpr.py:453-455)."""
    m, d = Xs.shape
    dx = np.abs(Xs[:, None, :] - X[None, :, :]).reshape(-1, d)
    return corr(kind, theta, dx).reshape(m, X.shape[0])


def correlation_matrix(kind, theta, X, block=None):
    """R = I + scatter(corr(theta, |x_j - x_k|))  (gpr.py:34-68, 658-668).

    The reference materialises the condensed (n(n-1)/2, d) distance array (10.7 GB at n = 8192,
    d = 40).  Above n = 2048 (or with `block` given) the same element-wise formulas are evaluated
    on row blocks of the full matrix instead: every entry is the same float64 expression of the
    same |x_j - x_k| row (reductions run along the feature axis only), so both constructions give
    bit-identical matrices (tests/test_oracle_golden.py checks it); exp(-0) = 1 on the diagonal."""
    n, d = X.shape
    if block is None and n > 2048:
        block = max(1, (1 << 24) // (n * d))
    if block:
        R = np.empty((n, n))
        for i0 in range(0, n, block):
            D = np.abs(X[i0:i0 + block, None, :] - X[None, :, :]).reshape(-1, d)
            R[i0:i0 + block] = corr(kind, theta, D).reshape(-1, n)
        R[np.arange(n), np.arange(n)] = 1.0
        return R
    iu = np.triu_indices(n, 1)
    D = np.abs(X[iu[0]] - X[iu[1]])
    r = corr(kind, theta, D)
    R = np.eye(n)
    R[iu[0], iu[1]] = r
    R[iu[1], iu[0]] = r
    return R


def has_duplicate_rows(X):
    """gpr.py:274-277: `np.min(np.sum(D, axis=1)) == 0` raises in _check_data."""
    n = X.shape[0]
    if n < 2:
        return False
    iu = np.triu_indices(n, 1)
    return bool(np.min(np.abs(X[iu[0]] - X[iu[1]]).sum(axis=1)) == 0.0)


def corr_grad_theta(kind, theta, X, R0):
    """dR/dtheta_i as an (n, n, d) tensor (gpr.py:622-656)."""
    diff = (X[:, None, :] - X[None, :, :]) ** 2.0
    if kind == "squared_exponential":
        return -diff * R0[..., None]
    if kind == "matern":
        D = np.sqrt(np.sum(_full_theta(theta, X.shape[1]) * diff, axis=-1))
        return -3.0 * np.exp(-SQRT3 * D)[..., None] * diff / 2.0
    if kind == "absolute_exponential":
        return -np.sqrt(diff) * R0[..., None]
    raise ValueError(kind)


# --------------------------------------------------------------------------------------
# auxiliary variables and likelihood  (gpr.py:676-697, 798-946)
# --------------------------------------------------------------------------------------
def trend_F(trend, X):
    """Basis functions of the prior mean at X: 'constant' -> 1 (trend.py:69-88),
    'linear' -> [1, x_1..x_d] (trend.py:90-110).  (m, p)."""
    X = np.atleast_2d(X)
    if trend == "constant":
        return np.ones((X.shape[0], 1))
    if trend == "linear":
        return np.c_[np.ones(X.shape[0]), X]
    raise ValueError(trend)


def trend_jacobian(trend, d):
    """f_dx (p, d), numerator layout: zeros (trend.py:86-88) | [0; I] (trend.py:106-109)."""
    if trend == "constant":
        return np.zeros((1, d))
    if trend == "linear":
        return np.r_[np.zeros((1, d)), np.eye(d)]
    raise ValueError(trend)


def _beta_vec(beta, p):
    """BasisExpansionTrend.set_beta (trend.py:27-34): a scalar is repeated p times."""
    b = np.asarray(beta, dtype=np.float64).ravel()
    if b.size == 1:
        b = np.repeat(b, p)
    if b.size != p:
        raise ValueError("Shapes of beta and F do not match.")
    return b.reshape(-1, 1)


def compute_aux_var(R, y, estimate_trend, beta, F=None):
    """L, Ft, Yt, Q, G, rho  (gpr.py:676-697).  F: (n, p) basis matrix (default: constant)."""
    n = R.shape[0]
    if F is None:
        F = np.ones((n, 1))
    L = linalg.cholesky(R, lower=True)
    Yt = linalg.solve_triangular(L, y.reshape(-1, 1), lower=True)
    if estimate_trend:
        Ft = linalg.solve_triangular(L, F, lower=True)
        Q, G = linalg.qr(Ft, mode="economic")
        rho = Yt - Q.dot(Q.T).dot(Yt)
    else:
        mu = F.dot(_beta_vec(beta, F.shape[1]))
        rho = Yt - linalg.solve_triangular(L, mu, lower=True)
        Ft, Q, G = None, None, None
    return L, Ft, Yt, Q, G, rho


def log_likelihood_concentrated(X, y, par, kind="squared_exponential", estimate_trend=True,
                                beta=0.0, mode="noiseless", noise_var=0.0, eval_grad=False,
                                env=None, trend="constant"):
    """Concentrated log-likelihood and its gradient w.r.t. the raw parameters.

    Follows gpr.py:798-946.  mode 'noiseless': par=theta (gpr.py:814-835);
    'noisy': par=[theta, sigma2], R=(sigma2 R0 + noise_var I)/(sigma2+noise_var) (gpr.py:855-874);
    'noise_estim': par=[theta, alpha] (gpr.py:837-853).
    Not-PD Cholesky -> (-inf, zeros) (gpr.py:820-826); llf > 0 -> (-inf, zeros) (gpr.py:889-891).
    The gradient is the reference's formula (gpr.py:896-944): it ignores the dependence of
    beta-hat on theta and uses sigma2-hat with the (n-k) divisor, and is
    d/dtheta, not d/dlog10(theta).
    k = matrix_rank(Q Q^T) (gpr.py:829-832) equals p=1 for a full-rank constant trend (Q11).
    Returns llf (float) or (llf, grad (n_par,1)).
    """
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    par = np.asarray(par, dtype=np.float64).ravel()
    n, d = X.shape
    n_par = par.size
    fail = (-np.inf, np.zeros((n_par, 1))) if eval_grad else -np.inf

    if mode == "noiseless":
        theta = par
        nv = 0.0
        R0 = correlation_matrix(kind, theta, X)
        R = R0
    elif mode == "noisy":
        theta, sigma2 = par[:-1], par[-1]
        nv = float(np.asarray(noise_var).ravel()[0])
        sigma2_total = sigma2 + nv
        R0 = correlation_matrix(kind, theta, X)
        R = (sigma2 * R0 + nv * np.eye(n)) / sigma2_total
    elif mode == "noise_estim":
        theta, alpha = par[:-1], par[-1]
        R0 = correlation_matrix(kind, theta, X)
        R = alpha * R0 + (1.0 - alpha) * np.eye(n)
    else:
        raise ValueError(mode)

    try:
        L, Ft, Yt, Q, G, rho = compute_aux_var(R, y, estimate_trend, beta, trend_F(trend, X))
    except linalg.LinAlgError:
        return fail

    logdet2 = 2.0 * np.log(np.diag(L)).sum()
    rr = float((rho ** 2.0).sum())
    if mode == "noiseless":
        k = Ft.shape[1] if estimate_trend else 0     # matrix_rank(Q Q^T) = p (gpr.py:829-832)
        sigma2 = rr / (n - k)
        llf = -0.5 * (n * math.log(2.0 * math.pi * sigma2) + logdet2 + n) if sigma2 > 0 \
            else np.inf
    elif mode == "noise_estim":
        sigma2_total = rr / n
        sigma2, nv = alpha * sigma2_total, (1.0 - alpha) * sigma2_total
        llf = -0.5 * (n * math.log(2.0 * math.pi * sigma2_total) + logdet2 + n)
    else:
        llf = -0.5 * (n * math.log(2.0 * math.pi * sigma2_total) + logdet2 + rr / sigma2_total)

    if env is not None:
        env.update(sigma2=sigma2, noise_var=nv, rho=rho, Yt=Yt, C=L, Ft=Ft, G=G, Q=Q)

    if llf > 0 or not np.isfinite(llf):
        return fail
    if not eval_grad:
        return float(llf)

    gamma = linalg.solve_triangular(L.T, rho, lower=False).reshape(-1, 1)
    Rinv = linalg.cho_solve((L, True), np.eye(n))
    iu = np.triu_indices(n, 1)
    Rinv_u = Rinv[iu]
    gg_u = gamma.dot(gamma.T)[iu]
    grad = np.zeros((n_par, 1))
    if mode == "noiseless":
        T = corr_grad_theta(kind, theta, X, R0)
        for i in range(n_par):
            Tu = T[:, :, i][iu]
            grad[i] = np.sum(gg_u * Tu) / sigma2 - np.sum(Rinv_u * Tu)
    elif mode == "noise_estim":
        T = alpha * corr_grad_theta(kind, theta, X, R0)
        for i in range(n_par - 1):
            Tu = T[:, :, i][iu]
            grad[i] = np.sum(gg_u * Tu) / sigma2_total - np.sum(Rinv_u * Tu)
        R_dv = R0 - np.eye(n)
        grad[n_par - 1] = -0.5 * (np.sum(Rinv * R_dv)
                                  - gamma.T.dot(R_dv.dot(gamma)).item() / sigma2_total)
    else:
        gamma_ = gamma / sigma2_total
        Cinv = Rinv / sigma2_total
        T = sigma2_total * corr_grad_theta(kind, theta, X, R0)
        T = np.concatenate([T, R0[..., None]], axis=2)
        for i in range(n_par):
            Cg = T[:, :, i]
            grad[i] = -0.5 * (np.sum(Cinv * Cg) - gamma_.T.dot(Cg).dot(gamma_).item())
    return float(llf), grad


def log_likelihood_restricted(X, y, par, kind="squared_exponential", estimate_trend=True,
                              beta=0.0, mode="noiseless", noise_var=0.0, eval_grad=False, env=None,
                              trend="constant"):
    """Restricted (REML) log-likelihood and gradient, gpr.py:699-796.

    par = [theta, sigma2] ('noiseless': noise 0; 'noisy': the given nugget) or
    [theta, sigma2, noise_var] ('noise_estim') (gpr.py:712-719);
    R = (sigma2 R0 + noise_var I) / (sigma2 + noise_var) (:721-724).
    Ordinary kriging (:733-739): -1/2 ((n-p) log(2 pi v) - log det(F^T F) + 2 sum log L_ii
    + log(prod diag(G))^2 + rho^T rho / v); simple kriging (:740-743) carries the reference's
    MINUS sign on the log-determinant term -- reproduced, not corrected.
    exp(llf) > 1 -> -inf (:745-748; the gradient is still returned).  Gradient (:750-783):
    g_i = -1/2 (sum C^-1 o dC_i - gamma_^T dC_i gamma_ - sum (qq^T) o dC_i), q = L^-T Q (ordinary
    kriging only), dC = v dR0/dtheta_i | R0 (sigma2) | I (noise_var).
    """
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    par = np.asarray(par, dtype=np.float64).ravel()
    n, d = X.shape
    n_par = par.size
    if mode == "noiseless":
        theta, sigma2, nv = par[:-1], par[-1], 0.0
    elif mode == "noisy":
        theta, sigma2, nv = par[:-1], par[-1], float(np.asarray(noise_var).ravel()[0])
    elif mode == "noise_estim":
        theta, sigma2, nv = par[:-2], par[-2], par[-1]
    else:
        raise ValueError(mode)
    R0 = correlation_matrix(kind, theta, X)
    v = sigma2 + nv
    R = (sigma2 * R0 + nv * np.eye(n)) / v
    F = trend_F(trend, X)
    try:
        L, Ft, Yt, Q, G, rho = compute_aux_var(R, y, estimate_trend, beta, F)
    except linalg.LinAlgError:
        return (-np.inf, np.zeros((n_par, 1))) if eval_grad else -np.inf
    logdet2 = 2.0 * np.log(np.diag(L)).sum()
    rr = float((rho ** 2.0).sum())
    if estimate_trend:
        p = Ft.shape[1]
        llf = -0.5 * ((n - p) * math.log(2.0 * math.pi * v) - math.log(np.linalg.det(F.T.dot(F)))
                      + logdet2 + math.log(float(np.diag(G).prod()) ** 2) + rr / v)
    else:
        llf = -0.5 * (n * math.log(2.0 * math.pi * v) - logdet2 + rr / v)
    if llf > 0 or not np.isfinite(llf):       # np.exp(llf) > 1
        llf = -np.inf
    if env is not None:
        env.update(sigma2=sigma2, noise_var=nv, rho=rho, Yt=Yt, C=L, Ft=Ft, G=G, Q=Q)
    if not eval_grad:
        return float(llf)
    gamma_ = linalg.solve_triangular(L.T, rho, lower=False).reshape(-1, 1) / v
    Cinv = linalg.cho_solve((L, True), np.eye(n)) / v
    T = v * corr_grad_theta(kind, theta, X, R0)
    T = np.concatenate([T, R0[..., None]], axis=2)
    if mode == "noise_estim":
        T = np.concatenate([T, np.eye(n)[..., None]], axis=2)
    if estimate_trend:
        q = linalg.solve_triangular(L.T, Q, lower=False)
        term = q.dot(q.T)
    grad = np.zeros((n_par, 1))
    for i in range(n_par):
        Cg = T[:, :, i]
        gi = np.sum(Cinv * Cg) - gamma_.T.dot(Cg).dot(gamma_).item()
        if estimate_trend:
            gi -= np.sum(term * Cg)
        grad[i] = -0.5 * gi
    return float(llf), grad


def fit_state(X, y, par, kind="squared_exponential", estimate_trend=True, beta=0.0,
              mode="noiseless", noise_var=0.0, likelihood="concentrated", trend="constant"):
    """Fitted-model state for given hyper-parameters: what `fit` harvests after MLE.

    gpr.py:1087-1101 (final likelihood call with env), :370-383 (attributes),
    :670-674 (compute_beta_gamma: beta = G^-1 Q^T Yt, gamma = L^-T rho).
    """
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    par = np.asarray(par, dtype=np.float64).ravel()
    env = {}
    llf_fn = log_likelihood_restricted if likelihood == "restricted" else log_likelihood_concentrated
    llf = llf_fn(X, y, par, kind, estimate_trend, beta, mode, noise_var, eval_grad=False, env=env,
                 trend=trend)
    if not env:
        raise linalg.LinAlgError("correlation matrix is not positive definite")
    d = X.shape[1]
    # theta = everything before the variance parameters of the mode (gpr.py:712-719, 814, 837, 855);
    # a single theta entry is the isotropic form (kernel.py:312-317)
    extras = {"noiseless": 0, "noisy": 1, "noise_estim": 1}[mode]
    if likelihood == "restricted":
        extras = {"noiseless": 1, "noisy": 1, "noise_estim": 2}[mode]
    theta = par[:par.size - extras]
    st = dict(X=X, y=y, kind=kind, theta=np.array(theta, dtype=np.float64),
              estimate_trend=bool(estimate_trend), mode=mode,
              likelihood=likelihood, sigma2=float(env["sigma2"]), noise_var=float(env["noise_var"]),
              L=env["C"], rho=env["rho"], Yt=env["Yt"], Ft=env["Ft"], G=env["G"], Q=env["Q"],
              llf=llf, par=par, d=d)
    st["trend"] = trend
    p = trend_F(trend, X[:1]).shape[1]
    if estimate_trend:
        bvec = linalg.solve_triangular(st["G"], st["Q"].T.dot(st["Yt"])).reshape(-1, 1)
    else:
        bvec = _beta_vec(beta, p)
    st["beta_vec"] = bvec
    st["beta"] = float(bvec[0, 0]) if p == 1 else bvec.ravel().copy()
    st["gamma"] = linalg.solve_triangular(st["L"].T, st["rho"], lower=False).reshape(-1, 1)
    return st


# --------------------------------------------------------------------------------------
# posterior  (gpr.py:392-483, 511-617)
# --------------------------------------------------------------------------------------
def predict(st, Xs, eval_MSE=True, chunk=2048):
    """Posterior mean (m,) and MSE (m,) (gpr.py:453-479).

    MSE = |sigma2 (1 - sum rt^2 + sum u^2)| -- the reference's sqrt(MSE**2) is an absolute
    value and its `<0 -> 0` clamp is dead code.  Noisy mode uses the unscaled r
    against chol((sigma2 R0 + tau2 I)/(sigma2+tau2)).
    """
    Xs = np.atleast_2d(np.asarray(Xs, dtype=np.float64))
    m = Xs.shape[0]
    X, L = st["X"], st["L"]
    if Xs.shape[1] != X.shape[1]:
        raise ValueError("feature count mismatch")
    mean = np.empty(m)
    mse = np.empty(m) if eval_MSE else None
    for s in range(0, m, chunk):
        xs = Xs[s:s + chunk]
        r = cross_corr(st["kind"], st["theta"], xs, X)
        f = trend_F(st.get("trend", "constant"), xs)
        mean[s:s + chunk] = (f.dot(st["beta_vec"]) + r.dot(st["gamma"])).ravel()
        if eval_MSE:
            rt = linalg.solve_triangular(L, r.T, lower=True)
            if st["estimate_trend"]:
                u = linalg.solve_triangular(st["G"].T, st["Ft"].T.dot(rt) - f.T, lower=True)
                u2 = (u ** 2.0).sum(axis=0)
            else:
                u2 = 0.0
            mse[s:s + chunk] = np.abs(st["sigma2"] * (1.0 - (rt ** 2.0).sum(axis=0) + u2))
    return (mean, mse) if eval_MSE else mean


def corr_dx(st, x):
    """(d, n) Jacobian of r w.r.t. the candidate x (gpr.py:558-617).

    Any numpy warning (0/0 when a Matern candidate coincides with a training point)
    zeroes the whole Jacobian (gpr.py:588-615).
    """
    X, theta, kind = st["X"], st["theta"].reshape(-1, 1), st["kind"]
    diff = (x.reshape(1, -1) - X).T  # (d, n)
    r = corr(kind, st["theta"], np.abs(diff).T).reshape(1, -1)
    if kind == "squared_exponential":
        return -2.0 * r * (theta * diff)
    if kind == "matern":
        D = np.sqrt(np.sum(theta * diff ** 2.0, axis=0))
        if np.any(D == 0.0):
            return np.zeros_like(diff)
        g = diff * theta / D
        g *= -3.0 * D * np.exp(-SQRT3 * D)
        return g
    if kind == "absolute_exponential":
        return -r * theta * np.sign(diff)
    raise ValueError(kind)


def gradient_point(st, x):
    """Literal single-point posterior gradient: (d,1) mean_dx, (d,1) mse_dx (gpr.py:511-554).

    Uses the reference's d+1 forward solves; kept to validate `gradient` below.
    """
    x = np.asarray(x, dtype=np.float64).ravel()
    X, L = st["X"], st["L"]
    r = cross_corr(st["kind"], st["theta"], x.reshape(1, -1), X)
    r_dx = corr_dx(st, x).T  # (n, d)
    trend = st.get("trend", "constant")
    f = trend_F(trend, x.reshape(1, -1)).reshape(-1, 1)
    f_dx = trend_jacobian(trend, x.size)
    y_dx = st["beta_vec"].T.dot(f_dx) + st["gamma"].T.dot(r_dx)     # gpr.py:539
    rt = linalg.solve_triangular(L, r.T, lower=True)
    rt_dx = linalg.solve_triangular(L, r_dx, lower=True)
    mse_dx = -rt.T.dot(rt_dx)
    if st["estimate_trend"]:
        Ft = st["Ft"]
        u = Ft.T.dot(rt) - f
        u_dx = Ft.T.dot(rt_dx) - f_dx
        mse_dx = mse_dx + u.T.dot(linalg.inv(Ft.T.dot(Ft))).dot(u_dx)
    mse_dx = 2.0 * st["sigma2"] * mse_dx
    return y_dx.T, mse_dx.T


def gradient(st, Xs, chunk=512):
    """Batched posterior gradient: mean_dx (m,d), mse_dx (m,d).

    Element-wise application of gpr.py:511-554 using w = L^-T rt and a = L^-T Ft so that
    rt^T L^-1 r_dx = w^T r_dx and Ft^T L^-1 r_dx = a^T r_dx.
    """
    Xs = np.atleast_2d(np.asarray(Xs, dtype=np.float64))
    m, d = Xs.shape
    X, L, kind = st["X"], st["L"], st["kind"]
    theta = _full_theta(st["theta"], d)
    ok = st["estimate_trend"]
    trend = st.get("trend", "constant")
    f_dx = trend_jacobian(trend, d)                                   # (p, d)
    if ok:
        Ft = st["Ft"]
        a = linalg.solve_triangular(L.T, Ft, lower=False)             # (n, p) = L^-T Ft
        Ainv = linalg.inv(Ft.T.dot(Ft))
    mean_dx = np.empty((m, d))
    mse_dx = np.empty((m, d))
    for s in range(0, m, chunk):
        xs = Xs[s:s + chunk]
        diff = xs[:, None, :] - X[None, :, :]  # (c, n, d)
        r = cross_corr(kind, theta, xs, X)  # (c, n)
        if kind == "squared_exponential":
            r_dx = -2.0 * r[..., None] * (theta * diff)
        elif kind == "matern":
            D = np.sqrt(np.sum(theta * diff ** 2.0, axis=-1))
            r_dx = -3.0 * np.exp(-SQRT3 * D)[..., None] * (theta * diff)
            r_dx[np.any(D == 0.0, axis=1)] = 0.0
        elif kind == "absolute_exponential":
            r_dx = -r[..., None] * theta * np.sign(diff)
        else:
            raise ValueError(kind)
        rt = linalg.solve_triangular(L, r.T, lower=True)  # (n, c)
        w = linalg.solve_triangular(L.T, rt, lower=False)  # (n, c)
        mean_dx[s:s + chunk] = np.einsum("j,cji->ci", st["gamma"].ravel(), r_dx) + \
            st["beta_vec"].T.dot(f_dx)
        g = -np.einsum("jc,cji->ci", w, r_dx)
        if ok:
            u = Ft.T.dot(rt) - trend_F(trend, xs).T                   # (p, c)
            kap = Ainv.dot(u)                                         # (p, c)
            g = g + np.einsum("jc,cji->ci", a.dot(kap), r_dx) - kap.T.dot(f_dx)
        mse_dx[s:s + chunk] = 2.0 * st["sigma2"] * g
    return mean_dx, mse_dx


# --------------------------------------------------------------------------------------
# acquisition functions  (mipego/InfillCriteria.py)
# --------------------------------------------------------------------------------------
def _norm_pdf(x):
    return np.exp(-x ** 2 / 2.0) / _NORM_PDF_C


def default_plugin(st, minimize=True):
    """ImprovementBased.plugin default (InfillCriteria.py:66-71), *before* sign flip."""
    return float(np.min(st["y"])) if minimize else float(np.max(st["y"]))


def acquisition(st, Xs, kind, param=None, plugin=None, minimize=True, return_dx=False):
    """Element-wise acquisition value (m,) and optionally gradient (m,d).

    kind: 'EI' (InfillCriteria.py:113-148), 'PI' = EpsilonPI (:150-187; epsilon=param, 0 for the
    reference's un-constructible PI), 'UCB' (:75-111; param=alpha),
    'MGFI' (:196-252; param=t, capped at 22.36).
    `plugin` is the user-facing value (e.g. BO's fopt); it is negated when maximising
    (InfillCriteria.py:64-73); y_hat and y_dx are negated too (:36-47).
    Guards: EI -> 0 when sd/sqrt(sigma2) < 1e-6 (:121-123); MGFI -> 0 when isclose(sd, 0)
    (:220-221) or on overflow / non-finite (:223-236).
    """
    Xs = np.atleast_2d(np.asarray(Xs, dtype=np.float64))
    m, d = Xs.shape
    y_hat, mse = predict(st, Xs, eval_MSE=True)
    sd = np.sqrt(mse)
    if not minimize:
        y_hat = -y_hat
    if kind != "UCB":
        if plugin is None:
            plugin = default_plugin(st, minimize)
        p = plugin if minimize else -1.0 * plugin
    if return_dx:
        y_dx, sd2_dx = gradient(st, Xs)
        if not minimize:
            y_dx = -y_dx
        with np.errstate(all="ignore"):
            sd_dx = sd2_dx / (2.0 * sd)[:, None]

    with np.errstate(all="ignore"):
        if kind == "UCB":
            alpha = 0.5 if param is None else float(param)
            val = y_hat + alpha * sd
            if return_dx:
                dx = y_dx + alpha * sd_dx
        elif kind == "EI":
            dead = sd / math.sqrt(st["sigma2"]) < 1e-6
            xcr_ = p - y_hat
            xcr = xcr_ / sd
            cdf, pdf = ndtr(xcr), _norm_pdf(xcr)
            val = xcr_ * cdf + sd * pdf
            val[dead] = 0.0
            if return_dx:
                dx = -y_dx * cdf[:, None] + sd_dx * pdf[:, None]
                dx[dead] = 0.0
        elif kind == "PI":
            eps = 0.0 if param is None else float(param)
            coef = np.where(y_hat > 0, 1.0 - eps, 1.0 + eps)
            xcr = (p - coef * y_hat) / sd
            val = ndtr(xcr)
            if return_dx:
                dx = -(coef[:, None] * y_dx + xcr[:, None] * sd_dx) * (_norm_pdf(xcr) / sd)[:, None]
        elif kind == "MGFI":
            t = min(1.0 if param is None else float(param), 22.36)
            dead = np.isclose(sd, 0)
            y_hat_p = y_hat - t * sd ** 2.0
            beta_p = (p - y_hat_p) / sd
            term = t * (p - y_hat - 1.0)
            expo = term + t ** 2.0 * sd ** 2.0 / 2.0
            val = ndtr(beta_p) * np.exp(expo)
            bad = dead | ~np.isfinite(val) | (expo > 709.782712893384)
            val[bad] = 0.0
            if return_dx:
                term2 = np.exp(t * (p + t * sd ** 2.0 / 2.0 - y_hat - 1.0))
                m_prime_dx = y_dx - 2.0 * t * (sd[:, None] * sd_dx)
                beta_p_dx = -(m_prime_dx + beta_p[:, None] * sd_dx) / sd[:, None]
                dx = term2[:, None] * (_norm_pdf(beta_p)[:, None] * beta_p_dx
                                       + ndtr(beta_p)[:, None] * ((t ** 2) * sd[:, None] * sd_dx
                                                                  - t * y_dx))
                dx[dead] = 0.0
        else:
            raise ValueError(kind)
    return (val, dx) if return_dx else val
