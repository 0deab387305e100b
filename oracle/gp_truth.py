"""Extended-precision ("truth") restatement of the GP posterior for error-budget arguments.

TEST INFRASTRUCTURE ONLY (same rule as gp_oracle.py: imported by tests/ only, never by the product).

SURVEY H1: with nugget 0 the correlation matrix is nearly singular and two float64 evaluation
orders of the reference's own formulas differ by 4e-9 (mean) / 7e-8 (EI).  A relative gate against
the float64 reference is then a gate against the reference's rounding noise.  This module
evaluates the same formulas (gpr.py:676-697 aux variables, :670-674 beta/gamma, :453-479
posterior) in numpy longdouble (x87 80-bit, eps = 1.1e-19) with hand-written Cholesky and
substitutions (LAPACK has no extended-precision routines), so that a test can state
"the CUDA result is no farther from the truth than the reference is".

Two entry points:
  fit_truth(X, y, theta, ...)   everything from (X, y, theta) in extended precision
  state_truth(L, gamma, ...)    treats a GIVEN float64 fitted state (L, gamma, sigma2, beta) as exact
                                input and evaluates the posterior from it in extended precision
Constant trend only (ordinary kriging beta=None, or simple kriging with the given beta);
noiseless and noisy (fixed nugget) concentrated-likelihood modes.
"""
import numpy as np

LD = np.longdouble
SQRT3 = np.sqrt(LD(3))


def corr(kind, theta, S_or_D, is_S=False):
    """Correlation from S = sum theta_i d_i^2 (SE, Matern) or sum theta_i |d_i| (abs-exp)."""
    S = S_or_D
    if kind == "squared_exponential":
        return np.exp(-S)
    if kind == "matern":
        K = SQRT3 * np.sqrt(S)
        return (LD(1) + K) * np.exp(-K)
    if kind == "absolute_exponential":
        return np.exp(-S)
    raise ValueError(kind)


def weighted_dist(kind, theta, A, B):
    """S[a, b] for rows of A against rows of B, accumulated feature by feature in longdouble."""
    A = np.asarray(A, dtype=LD)
    B = np.asarray(B, dtype=LD)
    th = np.asarray(theta, dtype=LD).ravel()
    if th.size == 1:
        th = np.repeat(th, A.shape[1])
    S = np.zeros((A.shape[0], B.shape[0]), dtype=LD)
    for i in range(A.shape[1]):
        diff = A[:, i:i + 1] - B[None, :, i]
        S += th[i] * (np.abs(diff) if kind == "absolute_exponential" else diff * diff)
    return S


def cholesky(R):
    """Lower Cholesky factor, column by column (left-looking), longdouble."""
    R = np.array(R, dtype=LD)
    n = R.shape[0]
    L = np.zeros_like(R)
    for j in range(n):
        v = R[j:, j] - L[j:, :j].dot(L[j, :j])
        if not v[0] > 0:
            raise np.linalg.LinAlgError("not positive definite at pivot %d" % (j + 1))
        L[j:, j] = v / np.sqrt(v[0])
    return L


def solve_lower(L, B):
    """X with L X = B (forward substitution), B (n, k)."""
    L = np.asarray(L, dtype=LD)
    X = np.array(B, dtype=LD)
    n = L.shape[0]
    for j in range(n):
        X[j] = (X[j] - L[j, :j].dot(X[:j])) / L[j, j]
    return X


def solve_lower_T(L, B):
    """X with L^T X = B (back substitution)."""
    L = np.asarray(L, dtype=LD)
    X = np.array(B, dtype=LD)
    n = L.shape[0]
    for j in range(n - 1, -1, -1):
        X[j] = (X[j] - L[j + 1:, j].dot(X[j + 1:])) / L[j, j]
    return X


def fit_truth(X, y, theta, kind="squared_exponential", estimate_trend=True, beta=0.0, mode="noiseless",
              noise_var=0.0, sigma2_par=None):
    """Fitted state from (X, y, theta) in extended precision (gpr.py:658-697, 814-874, 670-674)."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    yv = np.asarray(y, dtype=LD).reshape(-1, 1)
    R0 = corr(kind, theta, weighted_dist(kind, theta, X, X))
    R0[np.arange(n), np.arange(n)] = LD(1)
    if mode == "noisy":
        s2, tau2 = LD(sigma2_par), LD(noise_var)
        R = (s2 * R0 + tau2 * np.eye(n, dtype=LD)) / (s2 + tau2)     # gpr.py:861-865
    else:
        R = R0
    L = cholesky(R)
    Yt = solve_lower(L, yv)
    one = np.ones((n, 1), dtype=LD)
    if estimate_trend:
        Ft = solve_lower(L, one)
        ftf = (Ft * Ft).sum()
        b = (Ft * Yt).sum() / ftf
        rho = Yt - Ft * b
    else:
        Ft, ftf, b = None, None, LD(beta)
        rho = Yt - solve_lower(L, one * b)
    if mode == "noisy":
        sigma2 = LD(sigma2_par)
    else:
        sigma2 = (rho * rho).sum() / LD(n - (1 if estimate_trend else 0))   # gpr.py:833
    gamma = solve_lower_T(L, rho)
    return dict(X=X, kind=kind, theta=np.asarray(theta, dtype=LD).ravel(), L=L, gamma=gamma, Ft=Ft, ftf=ftf,
                beta=b, sigma2=sigma2, estimate_trend=bool(estimate_trend))


def state_truth(X, theta, L, gamma, sigma2, beta, kind="squared_exponential", estimate_trend=True):
    """Adopt a float64 fitted state as EXACT input (what mgp_set_state receives)."""
    L = np.asarray(L, dtype=LD)
    n = L.shape[0]
    Ft = ftf = None
    if estimate_trend:
        Ft = solve_lower(L, np.ones((n, 1), dtype=LD))
        ftf = (Ft * Ft).sum()
    return dict(X=np.asarray(X, dtype=np.float64), kind=kind, theta=np.asarray(theta, dtype=LD).ravel(), L=L,
                gamma=np.asarray(gamma, dtype=LD).reshape(-1, 1), Ft=Ft, ftf=ftf, beta=LD(beta),
                sigma2=LD(sigma2), estimate_trend=bool(estimate_trend))


def predict(st, Xs):
    """Posterior mean and MSE (gpr.py:453-479) in extended precision; returned as longdouble."""
    Xs = np.atleast_2d(np.asarray(Xs, dtype=np.float64))
    r = corr(st["kind"], st["theta"], weighted_dist(st["kind"], st["theta"], Xs, st["X"]))     # (m, n)
    mean = st["beta"] + r.dot(st["gamma"]).ravel()
    rt = solve_lower(st["L"], r.T)                                                           # (n, m)
    q = (rt * rt).sum(axis=0)
    if st["estimate_trend"]:
        u = (st["Ft"].T.dot(rt).ravel() - LD(1))
        u2 = u * u / st["ftf"]
    else:
        u2 = LD(0)
    mse = np.abs(st["sigma2"] * (LD(1) - q + u2))
    return mean, mse


def ei_from(mean, mse, plugin, sigma2):
    """EI (InfillCriteria.py:114-135) evaluated in float64 from extended-precision mean / MSE: the
    epilogue itself is well conditioned, its own rounding (1e-16) is far below the errors studied."""
    from scipy.special import ndtr
    mean = np.asarray(mean, dtype=np.float64)
    s = np.sqrt(np.asarray(mse, dtype=np.float64))
    out = np.zeros_like(mean)
    ok = s / np.sqrt(float(sigma2)) >= 1e-6
    z = (plugin - mean[ok]) / s[ok]
    out[ok] = (plugin - mean[ok]) * ndtr(z) + s[ok] * np.exp(-0.5 * z * z) / np.sqrt(2 * np.pi)
    return out
