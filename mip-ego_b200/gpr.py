"""GaussianProcess: drop-in for mipego.GaussianProcess.GaussianProcess (gpr.py:80-1149) whose
arithmetic runs on the GPU through libmgp_b200 (include/mgp.h).

Same constructor, `fit / update / predict(eval_MSE) / gradient`, the attributes the BO loop and
the infill criteria read (`is_fitted`, `y`, `sigma2`, `X`, `theta_`, ...), the same exceptions.
Supported configuration: corr in {squared_exponential, matern(nu=1.5),
absolute_exponential},
`constant_trend` mean (beta given = simple kriging, beta=None = ordinary kriging),
likelihood in {'concentrated', 'restricted'}, estimation modes noiseless (nugget falsy), noisy
(nugget > 0) and noise_estim, optimizer in {'BFGS', 'CMA'}.  Anything else raises
NotImplementedError: there is no CPU fallback.
"""
import ctypes

import numpy as np
from numpy import log10

from . import _cabi
from ._mle import multistart_lbfgsb
from .trend import BasisExpansionTrend, constant_trend, linear_trend

MACHINE_EPSILON = np.finfo(np.double).eps


def _check_array(X, name="X"):
    """sklearn.utils.check_array semantics used by the reference (gpr.py:265,431): 2-D float64,
    finite; object-dtype `Solution` arrays are coerced (base.py:516-517)."""
    X = np.asarray(X)
    if X.dtype == object:
        X = np.array(X.tolist(), dtype=np.float64)
    X = np.ascontiguousarray(X, dtype=np.float64)
    if X.ndim == 1:
        raise ValueError("Expected 2D array, got 1D array instead")
    if X.ndim != 2:
        raise ValueError("Found array with dim %d. Expected 2" % X.ndim)
    if not np.isfinite(X).all():
        raise ValueError("Input %s contains NaN or infinity." % name)
    return X


class GaussianProcess(object):
    _optimizer_types = ["BFGS", "CMA"]
    _correlation_types = ("squared_exponential", "matern", "absolute_exponential")
    _likelihood_functions = ["concentrated", "restricted"]

    def __init__(self, mean, corr="squared_exponential", theta0=1e-1, thetaL=None, thetaU=None,
                 sigma2=None, nugget=None, noise_estim=False, optimizer="BFGS",
                 likelihood="concentrated", random_start=1, wait_iter=5, eval_budget=None,
                 random_state=None, verbose=False, device=None, restart_mode="batched"):
        # gpr.py:212-261
        self.mean = mean
        self.corr = corr
        self.sigma2 = sigma2
        self.verbose = verbose
        self.corr_type = corr
        self.is_fitted = False

        self.theta0 = np.array(theta0).flatten()
        self.thetaL = np.array(thetaL).flatten()
        self.thetaU = np.array(thetaU).flatten()
        if not (np.isfinite(self.thetaL.astype(float)).all() and
                np.isfinite(self.thetaU.astype(float)).all()):
            raise ValueError("all bounds are required finite.")

        self.optimizer = optimizer
        self.random_start = random_start
        self.random_state = random_state
        self.wait_iter = wait_iter
        self.eval_budget = eval_budget

        self.noise_var = np.atleast_1d(nugget) if nugget else 0
        self._nugget = self.noise_var
        self.noise_estim = True if noise_estim else False
        self.noisy = True if self.noise_var or self.noise_estim else False
        if not self.noisy:
            self.estimation_mode = "noiseless"
        elif self.noise_estim:
            self.estimation_mode = "noise_estim"
        else:
            self.estimation_mode = "noisy"

        if likelihood not in ("concentrated", "restricted"):
            raise AssertionError(likelihood)
        self.likelihood = likelihood

        if isinstance(self.mean, BasisExpansionTrend):
            self.mean_type = "basis_expansion"
            self.estimate_trend = True if self.mean.beta is None else False
        else:
            raise NotImplementedError("mipego_b200 supports BasisExpansionTrend means only")

        # ---- B200 path ------------------------------------------------------------------
        self.device = device
        self.restart_mode = restart_mode
        self._h = None
        self._host_state = None   # set by __setstate__, uploaded lazily
        self._cache = {}
        self._check_supported()

    # ------------------------------------------------------------------------------------
    def _check_supported(self):
        if callable(self.corr) or self.corr not in self._correlation_types:
            raise NotImplementedError(
                "corr=%r: the CUDA path implements %s" % (self.corr, self._correlation_types))
        if not isinstance(self.mean, (constant_trend, linear_trend)):
            raise NotImplementedError("constant_trend and linear_trend are implemented by the CUDA path")
        if self.optimizer not in self._optimizer_types:
            raise ValueError("optimizer should be one of %s" % self._optimizer_types)

    def _handle(self):
        if self._h is None:
            self._h = _cabi.Handle(self.device)
        return self._h

    def _trend_args(self, beta=None):
        """(trend code, beta array or None) for mgp_set_train."""
        kind = "linear" if isinstance(self.mean, linear_trend) else "constant"
        code = _cabi.TREND[(kind, self.estimate_trend)]
        if self.estimate_trend:
            return code, None
        b = self.mean.beta if beta is None else beta
        b = np.ascontiguousarray(np.ravel(b), dtype=np.float64)
        if b.size != self.mean.n_dim:
            raise Exception("Shapes of beta and F do not match.")
        return code, b

    def _set_train(self, X, y, beta=None):
        h = self._handle()
        n, d = X.shape
        if self.mean.n_feature != d:
            raise Exception("x does not have the right size!")
        code, b = self._trend_args(beta)
        rc = h.lib.mgp_set_train(h.h, n, d, _cabi.dptr(X), _cabi.dptr(y), _cabi.CORR[self.corr], code,
                                 _cabi.dptr(b))
        if rc == _cabi.MGP_EINVAL and b"duplicate rows" in h.lib.mgp_last_error(h.h):
            raise Exception("Multiple input features cannot have the same target value.")
        h.check(rc)
        # one theta for all features (kernel.py:312-317): the parameter vectors get 1 theta entry
        h.set_option("isotropic", 1 if (self.theta0.size == 1 and d > 1) else 0)

    def _beta_from_device(self):
        h = self._handle()
        b = np.zeros(self.mean.n_dim)
        h.check(h.lib.mgp_get_trend(h.h, _cabi.dptr(b)))
        return b.reshape(-1, 1)

    def _check_params(self):
        # gpr.py:1103-1149 (the parts that can fail)
        lth = self.theta0.size
        if self.thetaL is not None and self.thetaU is not None:
            if self.thetaL.size != lth or self.thetaU.size != lth:
                raise ValueError("theta0, thetaL and thetaU must have the same length.")
            if np.any(self.thetaL <= 0) or np.any(self.thetaU < self.thetaL):
                raise ValueError("The bounds must satisfy O < thetaL <= thetaU.")
        self.verbose = bool(self.verbose)
        self.random_start = int(self.random_start)

    def _check_data(self, X, y):
        # gpr.py:263-288
        X = _check_array(X)
        y = np.asarray(y)
        if y.dtype == object:
            y = np.array(y.tolist(), dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64).ravel()
        if y.shape[0] != X.shape[0]:
            raise ValueError("Found input variables with inconsistent numbers of samples: [%d, %d]"
                             % (X.shape[0], y.shape[0]))
        if not np.isfinite(y).all():
            raise ValueError("Input y contains NaN or infinity.")
        self.X, self.y = X, y.reshape(-1, 1)
        self._check_params()
        n, d = X.shape
        if self.theta0.size not in (1, d):
            raise ValueError("Length of theta must be 1 or %s" % d)      # kernel.py:314-315
        self._set_train(X, y)
        self._cache = {}
        self._host_state = None
        if self.estimate_trend:
            self.F = self.mean.F(X)

    # ------------------------------------------------------------------------------------
    # likelihood
    # ------------------------------------------------------------------------------------
    def _mode_nv(self):
        mode = _cabi.MODE[self.estimation_mode]
        if self.likelihood == "restricted":
            mode |= _cabi.LIK_RESTRICTED
        nv = float(np.ravel(self._nugget)[0]) if self.estimation_mode == "noisy" else 0.0
        return mode, nv

    def log_likelihood_batch(self, params, eval_grad=True):
        """(B, n_par) raw parameters -> llf (B,), grad (B, n_par), status (B,) in one launch
        sequence (the batched form of gpr.py:798-946)."""
        h = self._handle()
        P = _cabi.as_f64(params)
        P = P.reshape(1, -1) if P.ndim == 1 else P
        B, n_par = P.shape
        mode, nv = self._mode_nv()
        llf = np.empty(B)
        grad = np.zeros((B, n_par)) if eval_grad else None
        status = np.zeros(B, dtype=np.int32)
        h.check(h.lib.mgp_nll_batch(h.h, B, _cabi.dptr(P), n_par, mode, nv, 1 if eval_grad else 0,
                                    _cabi.dptr(llf), _cabi.dptr(grad),
                                    status.ctypes.data_as(_cabi.c_ip)))
        return llf, grad, status

    def log_likelihood_restricted(self, par, env=None, eval_grad=False):
        """gpr.py:699-796 (REML); par = [theta, sigma2 (, noise_var)]."""
        if self.likelihood != "restricted":
            raise ValueError("this model was built with likelihood=%r" % self.likelihood)
        return self._llf_single(par, env, eval_grad)

    def log_likelihood_concentrated(self, par, env=None, eval_grad=False):
        """Single-vector form with the reference's return types (gpr.py:798-946)."""
        if self.likelihood != "concentrated":
            raise ValueError("this model was built with likelihood=%r" % self.likelihood)
        return self._llf_single(par, env, eval_grad)

    def _llf_single(self, par, env=None, eval_grad=False):
        par = np.asarray(par, dtype=np.float64).ravel()
        llf, grad, status = self.log_likelihood_batch(par, eval_grad=eval_grad)
        if env is not None and status[0] <= 0:
            env.update(self._harvest_env(par))
        if eval_grad:
            return float(llf[0]), grad[0].reshape(-1, 1)
        return float(llf[0])

    def _harvest_env(self, par):
        """The `env` dictionary of a likelihood call (gpr.py:876-887, 785-794) through mgp_nll_env:
        computed in scratch memory, the fitted model (device and Python attributes) is not touched --
        the reference's env harvest has no side effects either."""
        h = self._handle()
        n = self.X.shape[0]
        pcols = self.mean.n_dim if self.estimate_trend else 1
        L, rho, Yt, Ft, sc = np.empty((n, n)), np.empty(n), np.empty(n), np.empty((n, pcols)), np.zeros(5)
        mode, nv = self._mode_nv()
        h.check(h.lib.mgp_nll_env(h.h, _cabi.dptr(par), par.size, mode, nv, _cabi.dptr(L), _cabi.dptr(rho),
                                  _cabi.dptr(Yt), _cabi.dptr(Ft), _cabi.dptr(sc)))
        env = dict(sigma2=self._typed_sigma2(sc[0]), noise_var=self._typed_noise_var(sc[4]),
                   rho=rho.reshape(-1, 1), Yt=Yt.reshape(-1, 1), C=L)
        if self.estimate_trend:
            if pcols == 1:
                G = np.array([[sc[2]]])
                Q = Ft / sc[2]
            else:
                from scipy import linalg
                Q, G = linalg.qr(Ft, mode="economic")
            env.update(Ft=Ft, G=G, Q=Q)
        elif self.likelihood == "concentrated":
            env.update(Ft=None, G=None, Q=None)     # gpr.py:693-697: the names exist, set to None
        return env

    def _typed_sigma2(self, v):
        # sigma2 is a (1,) array where the reference derives it from rho (concentrated
        # noiseless / noise_estim, gpr.py:833,849) and a numpy scalar where it is a parameter
        # (concentrated noisy, restricted)
        derived = self.likelihood == "concentrated" and self.estimation_mode != "noisy"
        return np.array([v]) if derived else np.float64(v)

    def _typed_noise_var(self, v):
        if self.estimation_mode == "noise_estim":
            derived = self.likelihood == "concentrated"
            return np.array([v]) if derived else np.float64(v)
        return self._nugget if self.estimation_mode == "noisy" else 0

    def _hyperparameter_bound(self, par_list):
        # gpr.py:948-962
        bounds = []
        for name in par_list:
            if name == "theta":
                bounds.append(np.c_[self.thetaL, self.thetaU])
            elif name == "sigma2":
                bounds.append(np.atleast_2d([1e-5, self.y.std() ** 2]))
            elif name in ("alpha", "noise_var"):
                bounds.append(np.atleast_2d([1e-10, 1.0 - 1e-10]))
        return np.concatenate(bounds, axis=0)

    def _optimize_hyperparameter(self, solve=None):
        # gpr.py:964-1101.  `solve` replaces the multi-start driver (dist.sharded_fit shards the
        # restarts over ranks); default: multistart_lbfgsb in this process.
        par_list = ["theta"]
        if self.likelihood == "restricted" or self.estimation_mode == "noisy":
            par_list += ["sigma2"]
        if self.estimation_mode == "noise_estim":
            par_list += ["alpha"] if self.likelihood == "concentrated" else ["noise_var"]
        bounds = self._hyperparameter_bound(par_list)
        log10bounds = log10(bounds)
        n_theta = len(self.thetaL)
        if hasattr(self, "theta_"):
            log10theta0 = log10(self.theta_)
        else:
            log10theta0 = log10(self.theta0)
        if self.estimation_mode == "noiseless" and self.likelihood == "concentrated":
            log10param = log10theta0
        else:
            log10param = np.r_[log10theta0, np.random.uniform(log10bounds[n_theta:, 0],
                                                               log10bounds[n_theta:, 1])]
        n_par = len(log10param)
        eval_budget = 200 * n_par if self.eval_budget is None else self.eval_budget
        # restart points are drawn exactly as the sequential loop would (gpr.py:1035-1036)
        starts = [log10param]
        for _ in range(1, self.random_start):
            starts.append(np.random.uniform(log10bounds[:, 0], log10bounds[:, 1]))

        def batch_fn(P):
            llf, grad, _ = self.log_likelihood_batch(10.0 ** P, eval_grad=True)
            return -llf, -grad

        if solve is None and self.optimizer == "CMA":
            solve = self._solve_cma
        if solve is None:
            solve = lambda *a: multistart_lbfgsb(*a, mode=self.restart_mode)  # noqa: E731
        param_opt, llf_opt, n_evals, infos = solve(batch_fn, np.array(starts), log10bounds, eval_budget,
                                                   self.wait_iter)
        self.eval_count = n_evals
        self.opt_info_ = infos
        optimal_param = 10.0 ** param_opt
        llf = self._finalize(optimal_param)
        param = {"theta": optimal_param[:n_theta]}
        for k, name in enumerate(par_list[1:]):
            param[name] = optimal_param[n_theta + k:n_theta + k + 1]
        return param, llf

    def _solve_cma(self, batch_fn, starts, log10bounds, eval_budget, wait_iter):
        """optimizer='CMA' (gpr.py:1066-1085): derivative-free search of the likelihood in the
        log10 box from the first start point, sigma0 = 0.25 x the box width, `eval_budget`
        evaluations.  The reference runs a (1+1)-Cholesky-CMA-ES, one likelihood per step; here a
        (mu/mu_w, lambda)-CMA-ES evaluates its whole population in one batched likelihood launch
        sequence (lambda = 4 + 3 ln n_par rounded up to a multiple of 8)."""
        from .optimizer import cma_es_argmax
        n_par = starts.shape[1]
        lam = int(8 * np.ceil((4 + 3 * np.log(n_par)) / 8.0))

        def values(P):
            llf, _, _ = self.log_likelihood_batch(10.0 ** P, eval_grad=False)
            return np.where(np.isfinite(llf), llf, -1e300)
        x, f, info = cma_es_argmax(values, log10bounds, lambda_=lam, max_eval=max(lam, eval_budget),
                                   sigma0=0.25, x0=starts[0],
                                   seed=np.random.randint(0, 2 ** 31 - 1))
        f0 = values(starts[:1])[0]
        if f0 >= f:
            x, f = starts[0].copy(), f0
            info["funcalls"] += 1
        return x, -f, info["funcalls"], [info]

    def _finalize(self, par):
        """mgp_finalize_fit: the env-harvesting final likelihood call (gpr.py:1087-1101)."""
        h = self._handle()
        par = _cabi.as_f64(par).ravel()
        mode, nv = self._mode_nv()
        out = ctypes.c_double(0.0)
        rc = h.lib.mgp_finalize_fit(h.h, _cabi.dptr(par), par.size, mode, nv, ctypes.byref(out))
        if rc == _cabi.MGP_ENOTPD:
            return -np.inf
        h.check(rc)
        self._cache = {}
        self._host_state = None
        sc = np.zeros(5)
        h.check(h.lib.mgp_get_state(h.h, None, None, None, None, None, None, _cabi.dptr(sc)))
        self.sigma2 = self._typed_sigma2(sc[0])
        self._env_noise_var = self._typed_noise_var(sc[4])
        if self.estimate_trend:
            self.mean.beta = self._beta_from_device()
        self._par_fitted = par.copy()
        return float(out.value)

    # ------------------------------------------------------------------------------------
    def fit(self, X, y, _solve=None):
        """gpr.py:328-385."""
        self._check_data(X, y)
        self.is_fitted = False
        self.par, llf = self._optimize_hyperparameter(_solve)
        if np.isinf(llf):
            # gpr.py:359-366
            print("Doubling theta upper bound. Increasing nugget...")
            self.thetaU = self.thetaU * 2
            self._nugget = self._nugget * 10
            self.par, llf = self._optimize_hyperparameter(_solve)
        if not self._fitted_on_device():
            raise np.linalg.LinAlgError("GP fit failed: correlation matrix is not positive definite")
        scalar_llf = self.likelihood == "restricted" or self.estimation_mode == "noiseless"
        self.log_likelihood_ = llf if scalar_llf else np.array([llf])
        self.theta_ = self.par["theta"]
        self.noise_var = self._env_noise_var     # gpr.py:371
        self.is_fitted = True
        return self

    def fit_fixed(self, X, y, par):
        """Fit at GIVEN hyper-parameters (no MLE): the final likelihood call of
        _optimize_hyperparameter + compute_beta_gamma (gpr.py:1087-1101, 670-674) for
        par = [theta (, sigma2 | alpha) (, noise_var)] in the layout of the chosen likelihood /
        estimation mode.  Used to adopt known hyper-parameters and by the benchmarks."""
        self._check_data(X, y)
        par = np.asarray(par, dtype=np.float64).ravel()
        llf = self._finalize(par)
        if not self._fitted_on_device():
            raise np.linalg.LinAlgError("correlation matrix is not positive definite")
        n_theta = self.theta0.size
        names = self._par_names()[1:]
        self.par = {"theta": par[:n_theta]}
        for k, name in enumerate(names):
            self.par[name] = par[n_theta + k:n_theta + k + 1]
        self.theta_ = self.par["theta"]
        self.log_likelihood_ = llf
        self.noise_var = self._env_noise_var
        self.is_fitted = True
        return self

    def _fitted_on_device(self):
        h = self._handle()
        sc = np.zeros(5)
        return h.lib.mgp_get_state(h.h, None, None, None, None, None, None, _cabi.dptr(sc)) == 0

    def _par_names(self):
        names = ["theta"]
        if self.likelihood == "restricted" or self.estimation_mode == "noisy":
            names.append("sigma2")
        if self.estimation_mode == "noise_estim":
            names.append("alpha" if self.likelihood == "concentrated" else "noise_var")
        return names

    def _par_vector(self):
        """The fitted hyper-parameters as the flat vector of the likelihood's layout."""
        full = getattr(self, "_par_fitted", None)
        if full is not None:
            return np.array(full, dtype=np.float64)
        return np.concatenate([np.ravel(self.par[k]) for k in self._par_names() if k in self.par])

    def update(self, X, y, reoptimize=True):
        """gpr.py:387-390 (`update` = `fit`; its TODO asks for incremental training).
        reoptimize=False keeps the current hyper-parameters.  When (X, y) extends the current
        training set by appended rows, the factorisation is EXTENDED (mgp_append_train: rank-k border
        update of L, L^-1 and the posterior operands, O(n^2 k) instead of O(n^3)); otherwise it is
        rebuilt at the current hyper-parameters (mgp_finalize_fit)."""
        if reoptimize or not self.is_fitted:
            return self.fit(X, y)
        par = self._par_vector()
        n_need = self.theta0.size + len(self._par_names()) - 1
        if par.size != n_need:
            raise ValueError("the adopted state holds %d hyper-parameters, the %s / %s likelihood needs %d: "
                             "pass the full parameter vector to set_fitted_state(par=...)"
                             % (par.size, self.likelihood, self.estimation_mode, n_need))
        self._ensure_device_state()
        Xn = _check_array(X)
        n0 = self.X.shape[0]
        if (self._append_ok and Xn.shape[0] > n0 and Xn.shape[1] == self.X.shape[1]
                and np.array_equal(Xn[:n0], self.X)
                and np.array_equal(np.asarray(y, dtype=np.float64).ravel()[:n0], self.y.ravel())):
            return self._append(Xn, np.asarray(y, dtype=np.float64).ravel(), par)
        return self.fit_fixed(X, y, par)

    _append_ok = True     # update(reoptimize=False) extends the factorisation for appended rows

    def _append(self, X, y, par):
        """Rank-k extension through mgp_append_train (O(n^2 k)); falls back to the rebuild at the
        same hyper-parameters where the library declines (MGP_ESTATE: small models, large batches of
        new rows, non-constant trends, models adopted from outside)."""
        h = self._handle()
        n0 = self.X.shape[0]
        Xa = np.ascontiguousarray(X[n0:], dtype=np.float64)
        ya = np.ascontiguousarray(y[n0:], dtype=np.float64)
        if not np.isfinite(ya).all():
            raise ValueError("Input y contains NaN or infinity.")
        out = ctypes.c_double(0.0)
        rc = h.lib.mgp_append_train(h.h, Xa.shape[0], _cabi.dptr(Xa), _cabi.dptr(ya), ctypes.byref(out))
        if rc == _cabi.MGP_ESTATE or rc == _cabi.MGP_ENOTPD:
            return self.fit_fixed(X, y, par)
        if rc == _cabi.MGP_EINVAL and b"duplicate rows" in h.lib.mgp_last_error(h.h):
            raise Exception("Multiple input features cannot have the same target value.")
        h.check(rc)
        self.X, self.y = X, y.reshape(-1, 1)
        self._cache = {}
        self._host_state = None
        sc = np.zeros(5)
        h.check(h.lib.mgp_get_state(h.h, None, None, None, None, None, None, _cabi.dptr(sc)))
        self.sigma2 = self._typed_sigma2(sc[0])
        self._env_noise_var = self._typed_noise_var(sc[4])
        self.noise_var = self._env_noise_var
        if self.estimate_trend:
            self.mean.beta = self._beta_from_device()
            self.F = self.mean.F(X)
        llf = float(out.value)
        scalar_llf = self.likelihood == "restricted" or self.estimation_mode == "noiseless"
        self.log_likelihood_ = llf if scalar_llf else np.array([llf])
        return self

    def set_fitted_state(self, X, y, theta, L, gamma, sigma2, beta, par=None):
        """Adopt an externally fitted model (mgp_set_state): used by un-pickling, by the multi-GPU
        broadcast and by the parity tests to give the reference and the CUDA path the identical
        model.  `par`: the full hyper-parameter vector [theta, (sigma2 | alpha), (noise_var)] in the
        layout of this model's likelihood / estimation mode; needed for a later
        update(reoptimize=False) in the noisy / noise_estim / REML modes (default: theta only)."""
        X = _check_array(X)
        y = np.ascontiguousarray(y, dtype=np.float64).ravel()
        self.X, self.y = X, y.reshape(-1, 1)
        n, d = X.shape
        h = self._handle()
        bvec = np.ascontiguousarray(np.ravel(beta), dtype=np.float64)
        if bvec.size != self.mean.n_dim:
            raise Exception("Shapes of beta and F do not match.")
        self._set_train(X, y, beta=bvec)
        theta = _cabi.as_f64(theta).ravel()
        if theta.size not in (1, d):
            raise ValueError("Length of theta must be 1 or %s" % d)
        th_full = np.repeat(theta, d) if theta.size == 1 and d > 1 else theta
        if self.theta0.size == 1 and d > 1:          # isotropic model: keep the 1-entry layout
            theta = th_full[:1].copy()
        Lc = _cabi.as_f64(L, (n, n))
        g = _cabi.as_f64(gamma).ravel()
        h.check(h.lib.mgp_set_state(h.h, _cabi.dptr(th_full), _cabi.dptr(Lc), _cabi.dptr(g),
                                    float(np.ravel(sigma2)[0]), _cabi.dptr(bvec)))
        self._cache = {}
        self._host_state = None
        self.theta_ = theta.copy()
        names = self._par_names()
        if par is not None:
            par = np.array(par, dtype=np.float64).ravel()
            if par.size != theta.size + len(names) - 1:
                raise ValueError("par has %d entries, expected %d" % (par.size, theta.size + len(names) - 1))
            self.par = {"theta": self.theta_}
            for k, name in enumerate(names[1:]):
                self.par[name] = par[theta.size + k:theta.size + k + 1]
            self._par_fitted = par
        else:
            self.par = {"theta": self.theta_}
            self._par_fitted = self.theta_.copy() if len(names) == 1 else None
        self.sigma2 = self._typed_sigma2(float(np.ravel(sigma2)[0]))
        if self.estimate_trend:
            self.mean.beta = bvec.reshape(-1, 1).copy()
            self.F = self.mean.F(X)
        self.is_fitted = True
        return self

    # ---- fitted arrays, fetched from the device on first access ---------------------------
    def _state(self, key):
        if self._host_state is not None and key in self._host_state:
            return self._host_state[key]
        if key not in self._cache:
            h = self._handle()
            n, d = self.X.shape
            pcols = self.mean.n_dim if self.estimate_trend else 1
            bufs = dict(theta=np.empty(d), L=np.empty((n, n)), gamma=np.empty(n), rho=np.empty(n),
                        Yt=np.empty(n), Ft=np.empty((n, pcols)), sc=np.empty(5))
            h.check(h.lib.mgp_get_state(h.h, *[_cabi.dptr(bufs[k]) for k in
                                               ("theta", "L", "gamma", "rho", "Yt", "Ft", "sc")]))
            self._cache = bufs
        return self._cache[key]

    @property
    def C(self):
        return self._state("L")

    @property
    def gamma(self):
        return self._state("gamma").reshape(-1, 1)

    @property
    def rho(self):
        return self._state("rho").reshape(-1, 1)

    @property
    def Yt(self):
        return self._state("Yt").reshape(-1, 1)

    @property
    def Ft(self):
        return self._state("Ft").reshape(self.X.shape[0], -1) if self.estimate_trend else None

    def _qr(self):
        """Q, G of Ft (gpr.py:691) for a p > 1 basis.  The posterior kernels work with the
        sign-free quantities (Ft^T Ft and its Cholesky factor); G and Q are only materialised
        for users who read these attributes, with the reference's own LAPACK routine so that the
        Householder signs agree."""
        if "QR" not in self._cache:
            from scipy import linalg
            self._cache["QR"] = linalg.qr(self.Ft, mode="economic")
        return self._cache["QR"]

    @property
    def G(self):
        if not self.estimate_trend:
            return None
        if self.mean.n_dim == 1:
            return np.array([[self._state("sc")[2]]])
        return self._qr()[1]

    @property
    def Q(self):
        if not self.estimate_trend:
            return None
        if self.mean.n_dim == 1:
            return self.Ft / self._state("sc")[2]
        return self._qr()[0]

    # ------------------------------------------------------------------------------------
    # posterior
    # ------------------------------------------------------------------------------------
    def _ensure_device_state(self):
        """Re-upload after un-pickling (the device handle is not part of the pickle)."""
        if self._host_state is not None:
            st = self._host_state
            self._host_state = None
            keep = {k: getattr(self, k) for k in ("par", "_par_fitted", "theta_", "sigma2", "noise_var",
                                                  "_env_noise_var", "log_likelihood_") if hasattr(self, k)}
            self.set_fitted_state(st["X"], st["y"], st["theta"], st["L"], st["gamma"],
                                  st["sigma2"], st["beta"])
            self.__dict__.update(keep)     # the pickled hyper-parameters, types and likelihood stay

    def predict(self, X, eval_MSE=False, batch_size=None):
        """gpr.py:392-483.  `batch_size` is accepted and ignored: the reference's chunking
        branch is broken on Python 3 and chunking happens on the device here."""
        assert hasattr(self, "X")
        X = _check_array(X)
        n_eval, nf = X.shape
        if nf != self.X.shape[1]:
            raise ValueError(("The number of features in X (X.shape[1] = %d) should match the "
                              "number of features used for fit() which is %d.")
                             % (nf, self.X.shape[1]))
        self._ensure_device_state()
        h = self._handle()
        y = np.empty(n_eval)
        mse = np.empty(n_eval) if eval_MSE else None
        h.check(h.lib.mgp_predict(h.h, n_eval, _cabi.dptr(X), _cabi.dptr(y), _cabi.dptr(mse)))
        return (y, mse) if eval_MSE else y

    def predict_device(self, X, eval_MSE=True):
        """Device-resident predict: X a contiguous float64 CUDA tensor (m, d); returns CUDA tensors.
        Enqueued on torch's current stream, not synchronised."""
        import torch
        self._ensure_device_state()
        h = self._handle()
        m = X.shape[0]
        mean = torch.empty(m, dtype=torch.float64, device=X.device)
        mse = torch.empty(m, dtype=torch.float64, device=X.device) if eval_MSE else None
        stream = torch.cuda.current_stream(X.device).cuda_stream
        h.check(h.lib.mgp_predict_dev(h.h, m, X.data_ptr(), mean.data_ptr(),
                                      mse.data_ptr() if eval_MSE else None, stream))
        return (mean, mse) if eval_MSE else mean

    def log_likelihood_batch_device(self, params, eval_grad=True):
        """Device-resident batched likelihood: params a float64 CUDA tensor (B, n_par)."""
        import torch
        h = self._handle()
        B, n_par = params.shape
        mode, nv = self._mode_nv()
        llf = torch.empty(B, dtype=torch.float64, device=params.device)
        grad = torch.zeros((B, n_par), dtype=torch.float64, device=params.device)
        status = torch.zeros(B, dtype=torch.int32, device=params.device)
        stream = torch.cuda.current_stream(params.device).cuda_stream
        h.check(h.lib.mgp_nll_batch_dev(h.h, B, params.data_ptr(), n_par, mode, nv,
                                        1 if eval_grad else 0, llf.data_ptr(), grad.data_ptr(),
                                        status.data_ptr(), stream))
        return llf, grad, status

    def gradient(self, x):
        """gpr.py:511-554: single point; returns ((d,1) mean_dx, (d,1) mse_dx)."""
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        if x.shape[1] != self.X.shape[1]:
            raise Exception("x does not have the right size!")
        if x.shape[0] != 1:
            raise Exception("x must be a vector!")
        a, b = self.gradient_batch(x)
        return a.T, b.T

    def gradient_batch(self, X):
        """Row-wise gradient for (m, d) candidates: mean_dx (m, d), mse_dx (m, d)."""
        X = _check_array(np.atleast_2d(X))
        self._ensure_device_state()
        h = self._handle()
        m, d = X.shape
        a, b = np.empty((m, d)), np.empty((m, d))
        h.check(h.lib.mgp_gradient(h.h, m, _cabi.dptr(X), _cabi.dptr(a), _cabi.dptr(b)))
        return a, b

    # ------------------------------------------------------------------------------------
    # pickling (base.py:575-608 saves the optimiser with dill; joblib ships models to workers)
    # ------------------------------------------------------------------------------------
    def __getstate__(self):
        st = dict(self.__dict__)
        host = None
        if self.is_fitted:
            if self._host_state is not None:
                host = self._host_state
            else:
                host = dict(X=self.X, y=self.y.ravel(), theta=np.array(self._state("theta")),
                            L=np.array(self._state("L")), gamma=np.array(self._state("gamma")),
                            rho=np.array(self._state("rho")), Yt=np.array(self._state("Yt")),
                            Ft=np.array(self._state("Ft")), sc=np.array(self._state("sc")),
                            sigma2=float(np.ravel(self.sigma2)[0]),
                            beta=np.array(np.ravel(self.mean.beta), dtype=np.float64))
        st["_h"] = None
        st["_cache"] = {}
        st["_host_state"] = host
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
