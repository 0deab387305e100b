"""Generate golden fixtures by running the UNMODIFIED reference (container only).

    python oracle/gen_golden.py            # writes tests/golden/*.npz

TEST INFRASTRUCTURE ONLY.  Imports /root/reference/mipego through oracle/ref_shim.py and
records, for seeded inputs, what the reference's own code returns on the GP hot path:
  * log_likelihood_concentrated(par, env, eval_grad=True)       gpr.py:798-946
  * the fitted state harvested from env + compute_beta_gamma     gpr.py:370-383, 670-674
  * predict(X, eval_MSE=True)                                    gpr.py:392-483
  * gradient(x) per point                                        gpr.py:511-554
  * EI / EpsilonPI / UCB / MGFI  __call__(x, return_dx=True) per point   InfillCriteria.py
  * one full fit() (L-BFGS-B MLE)                                gpr.py:328-385, 964-1101
The fixtures are small (n <= 96) and committed; the `-m "not gpu"` suite checks the oracle
against them and the `-m gpu` suite checks the CUDA path against them.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ref_shim import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def make_data(seed, n, d, lb=-5.0, ub=5.0, standardize=True):
    rng = np.random.default_rng(seed)
    X = rng.uniform(lb, ub, (n, d))
    y = (X ** 2).sum(axis=1) + 3.0 * np.sin(X[:, 0])
    if standardize:
        y = (y - y.mean()) / y.std()
    return X, y


def ref_model(kind, d, beta, nugget, theta0, thetaL, thetaU, likelihood="concentrated",
              trend="constant", **kw):
    from mipego.GaussianProcess import GaussianProcess
    from mipego.GaussianProcess.trend import constant_trend, linear_trend
    mean = constant_trend(d, beta=beta) if trend == "constant" else linear_trend(d, beta=beta)
    return GaussianProcess(mean=mean, corr=kind, theta0=theta0,
                           thetaL=thetaL, thetaU=thetaU, nugget=nugget, optimizer="BFGS",
                           likelihood=likelihood, **kw)


def ref_set_par(gp, X, y, par):
    """Do what fit() does after MLE, for a *given* parameter vector (gpr.py:349-383)."""
    from sklearn.utils import check_random_state
    gp.random_state = check_random_state(gp.random_state)
    gp._check_data(X, y)
    env = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fn = gp.log_likelihood_restricted if gp.likelihood == "restricted" else \
            gp.log_likelihood_concentrated
        llf, grad = fn(np.asarray(par, dtype=float), env=None, eval_grad=True)
        llf2 = fn(np.asarray(par, dtype=float), env=env)
    n_theta = len(gp.thetaL)
    gp.par = {"theta": np.asarray(par[:n_theta], dtype=float)}
    gp.theta_ = gp.par["theta"]
    gp.noise_var = env["noise_var"]
    gp.sigma2 = env["sigma2"]
    gp.rho, gp.Yt, gp.C = env["rho"], env["Yt"], env["C"]
    if gp.estimate_trend:
        gp.Ft, gp.G, gp.Q = env["Ft"], env["G"], env["Q"]
    gp.compute_beta_gamma()
    gp.is_fitted = True
    return float(np.ravel(llf)[0]), np.asarray(grad, dtype=float).reshape(-1), float(np.ravel(llf2)[0])


def per_point(fn, Xs, d):
    vals, dxs = [], []
    for x in Xs:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            v, g = fn(x)
        vals.append(float(np.ravel(v)[0]) if np.size(v) else 0.0)
        dxs.append(np.asarray(g, dtype=float).reshape(d))
    return np.array(vals), np.array(dxs)


def case_fixed_par(name, seed, n, d, kind, beta, nugget, m=24, theta_scale=0.12,
                   likelihood="concentrated", noise_estim=False, trend="constant", isotropic=False):
    """Model state at fixed hyper-parameters + all posterior / acquisition outputs."""
    from mipego import InfillCriteria as IC
    X, y = make_data(seed, n, d)
    if trend == "linear":       # give the linear basis something to explain
        y = y + 0.3 * X[:, 1] - 0.2 * X[:, 0]
    rng = np.random.default_rng(seed + 1000)
    theta = (theta_scale / d) * rng.uniform(0.5, 1.5, d) * (4.0 if kind == "matern" else 1.0)
    if isotropic:               # ONE theta for all features (kernel.py:312-317)
        theta = theta[:1]
    nth = theta.size
    noisy = bool(nugget)
    if likelihood == "restricted":          # [theta, sigma2 (, noise_var)]  gpr.py:712-719
        par = np.r_[theta, 0.8, 0.01] if noise_estim else np.r_[theta, 0.8]
    elif noise_estim:                       # [theta, alpha]                 gpr.py:837
        par = np.r_[theta, 0.95]
    else:
        par = np.r_[theta, 0.8] if noisy else theta
    gp = ref_model(kind, d, beta, nugget, theta0=theta, thetaL=1e-5 * np.ones(nth),
                   thetaU=100 * np.ones(nth), likelihood=likelihood, noise_estim=noise_estim,
                   trend=trend)
    llf, grad, llf2 = ref_set_par(gp, X, y, par)
    Xs = rng.uniform(-5, 5, (m, d))
    # a few candidates close to / on training points exercise the small-variance guards
    Xs[0] = X[3] + 1e-9
    Xs[1] = X[5]
    Xs[2] = X[7] + 1e-3
    mean, mse = gp.predict(Xs, eval_MSE=True)
    mdx, sdx = [], []
    for x in Xs:
        a, b = gp.gradient(x)
        mdx.append(np.ravel(a)); sdx.append(np.ravel(b))
    out = dict(X=X, y=y, par=par, theta=theta, kind=kind, nugget=float(nugget or 0.0),
               likelihood=likelihood, mode=gp.estimation_mode, estimate_trend=beta is None,
               trend=trend, beta_in=np.nan if beta is None else np.asarray(beta, dtype=float),
               llf=llf, llf_env=llf2, llf_grad=grad,
               sigma2=float(np.ravel(gp.sigma2)[0]), noise_var=float(np.ravel(gp.noise_var)[0]),
               L=gp.C, rho=gp.rho, Yt=gp.Yt, gamma=gp.gamma,
               beta=float(np.ravel(gp.mean.beta)[0]) if trend == "constant" else np.ravel(gp.mean.beta),
               Xs=Xs, mean=mean, mse=mse, mean_dx=np.array(mdx), mse_dx=np.array(sdx))
    if beta is None:
        out.update(Ft=gp.Ft, G=gp.G, Q=gp.Q)
    plugin_user = float(np.min(y)) + 0.05
    out["plugin"] = plugin_user
    for minimize in (True, False):
        tag = "min" if minimize else "max"
        pl = plugin_user if minimize else float(np.max(y)) - 0.05
        out["plugin_" + tag] = pl
        crits = {
            "EI": IC.EI(model=gp, minimize=minimize, plugin=pl),
            "PI": IC.EpsilonPI(model=gp, minimize=minimize, plugin=pl, epsilon=0.05),
            "UCB": IC.UCB(model=gp, minimize=minimize, alpha=0.7),
            "MGFI": IC.MGFI(model=gp, minimize=minimize, plugin=pl, t=1.7),
        }
        for cname, c in crits.items():
            v, g = per_point(lambda x, c=c: c(x, return_dx=True), Xs, d)
            out["%s_%s" % (cname, tag)] = v
            out["%s_%s_dx" % (cname, tag)] = g
    out["param_PI"] = 0.05
    out["param_UCB"] = 0.7
    out["param_MGFI"] = 1.7
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, "llf=%.6f" % llf)


def case_fit(name, seed, n, d, kind, beta, nugget, random_start=1, isotropic=False):
    """A full reference fit (MLE by L-BFGS-B from theta0; gpr.py:964-1101)."""
    X, y = make_data(seed, n, d)
    lb, ub = -5.0, 5.0
    nth = 1 if isotropic else d
    thetaL = 1e-5 * (ub - lb) * np.ones(nth)
    thetaU = 10 * (ub - lb) * np.ones(nth)
    theta0 = np.full(nth, 0.05)
    np.random.seed(seed)
    gp = ref_model(kind, d, beta, nugget, theta0=theta0, thetaL=thetaL, thetaU=thetaU,
                   random_start=random_start, random_state=seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.random.seed(seed)
        # the noisy mode draws its sigma2 start from the global RNG (gpr.py:1010-1011)
        rs = np.random.get_state()
        gp.fit(X, y)
    np.random.set_state(rs)
    extra = np.random.uniform(0, 1, 4)  # lets a test re-derive the same start draw
    rng = np.random.default_rng(seed + 7)
    Xs = rng.uniform(lb, ub, (32, d))
    mean, mse = gp.predict(Xs, eval_MSE=True)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), X=X, y=y, kind=kind, nugget=float(nugget or 0.0),
        estimate_trend=beta is None, beta_in=np.nan if beta is None else float(beta),
        theta0=theta0, thetaL=thetaL, thetaU=thetaU, theta_=gp.theta_,
        par=np.concatenate([np.ravel(v) for v in gp.par.values()]),
        llf=float(np.ravel(gp.log_likelihood_)[0]), sigma2=float(np.ravel(gp.sigma2)[0]),
        eval_count=int(gp.eval_count), Xs=Xs, mean=mean, mse=mse, seed=seed,
        rng_probe=extra, random_start=random_start)
    print("wrote", name, "theta_=", gp.theta_, "llf=", gp.log_likelihood_, "evals", gp.eval_count)


def case_c1(name="c1_example"):
    """BASELINE config 1: n=50 d=2, y=sum x^2 (example/exmaple_continuous_variables.py:15-17),
    SE + ordinary kriging, nugget 0, fitted by the reference; EI on a 512-subset of the
    10^4 seed-2 candidates (the reference evaluates one point per call)."""
    from mipego import InfillCriteria as IC
    n, d = 50, 2
    rng = np.random.default_rng(1)
    X = rng.uniform(-5, 5, (n, d))
    y = (X ** 2).sum(axis=1)
    thetaL = 1e-10 * 10 * np.ones(d)
    thetaU = 10 * 10 * np.ones(d)
    theta0 = np.full(d, 0.5)
    np.random.seed(1)
    gp = ref_model("squared_exponential", d, None, 0, theta0=theta0, thetaL=thetaL,
                   thetaU=thetaU, random_start=d, random_state=1, wait_iter=3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gp.fit(X, y)
    Xs_all = np.random.default_rng(2).uniform(-5, 5, (10000, d))
    Xs = Xs_all[:512]
    mean, mse = gp.predict(Xs, eval_MSE=True)
    ei = IC.EI(model=gp, minimize=True)
    v, g = per_point(lambda x: ei(x, return_dx=True), Xs, d)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), X=X, y=y, kind="squared_exponential", nugget=0.0,
        estimate_trend=True, theta_=gp.theta_, par=gp.theta_,
        llf=float(np.ravel(gp.log_likelihood_)[0]), sigma2=float(np.ravel(gp.sigma2)[0]),
        L=gp.C, gamma=gp.gamma, rho=gp.rho, Yt=gp.Yt, Ft=gp.Ft, G=gp.G,
        beta=float(np.ravel(gp.mean.beta)[0]), Xs=Xs, mean=mean, mse=mse, EI=v, EI_dx=g,
        plugin=float(np.min(y)))
    print("wrote", name, "theta_=", gp.theta_)


def variants():
    """Estimation-mode / likelihood variants: REML in its three modes
    (gpr.py:699-796) and the concentrated likelihood with an estimated noise share (:837-853)."""
    k = 0
    for kind in ("squared_exponential", "matern"):
        for beta in (None, 0.3):
            tr = "ok" if beta is None else "sk"
            kn = "se" if kind[0] == "s" else "m32"
            for lik, nugget, ne, tag in (("restricted", 0, False, "reml_noiseless"),
                                         ("restricted", 1e-4, False, "reml_noisy"),
                                         ("restricted", 0, True, "reml_noiseestim"),
                                         ("concentrated", 0, True, "conc_noiseestim")):
                k += 1
                case_fixed_par("var_%s_%s_%s" % (kn, tr, tag), seed=300 + k, n=56 + 8 * (k % 3),
                               d=3 + (k % 2), kind=kind, beta=beta, nugget=nugget, m=16,
                               likelihood=lik, noise_estim=ne)


def linear_trends():
    """Universal kriging with the first-order polynomial trend (trend.py:90-110), beta estimated
    (p = d + 1) or given; concentrated (noiseless / noisy) and restricted likelihoods."""
    k = 0
    for kind in ("squared_exponential", "matern"):
        for est in (True, False):
            for lik, nugget, tag in (("concentrated", 0, "noiseless"), ("concentrated", 1e-6, "noisy"),
                                     ("restricted", 1e-4, "reml_noisy")):
                k += 1
                d = 3 + (k % 2)
                beta = None if est else list(np.round(np.linspace(0.2, -0.1, d + 1), 3))
                case_fixed_par("lin_%s_%s_%s" % ("se" if kind[0] == "s" else "m32", "uk" if est else "sk", tag),
                               seed=400 + k, n=56 + 8 * (k % 3), d=d, kind=kind, beta=beta, nugget=nugget,
                               m=16, likelihood=lik, trend="linear")
    linear_trends_absexp()


def linear_trends_absexp():
    """The linear trend with the absolute-exponential kernel (estimated and given coefficients)."""
    case_fixed_par("lin_absexp_uk_noisy", seed=431, n=64, d=3, kind="absolute_exponential", beta=None,
                   nugget=1e-6, m=16, trend="linear")
    case_fixed_par("lin_absexp_uk_reml_noisy", seed=432, n=56, d=4, kind="absolute_exponential", beta=None,
                   nugget=1e-4, m=16, likelihood="restricted", trend="linear")
    case_fixed_par("lin_absexp_sk_noiseless", seed=433, n=56, d=3, kind="absolute_exponential",
                   beta=[0.2, 0.1, 0.0, -0.1], nugget=0, m=16, trend="linear")


def zero_theta():
    """Likelihood + gradient with some theta_i exactly 0 (the CUDA kernels fold sqrt(theta_i) into the
    staged coordinates and take a separate, unscaled path for such vectors; the reference has no
    special case: gpr.py:798-946 with kernel.py:147-317)."""
    out = {}
    for tag, kind in (("se", "squared_exponential"), ("m32", "matern"), ("absexp", "absolute_exponential")):
        d, n = 5, 72
        X, y = make_data(140 + len(tag), n, d)
        rng = np.random.default_rng(1400 + len(tag))
        theta = (0.3 / d) * 10.0 ** rng.uniform(-0.3, 0.3, d)
        theta[2] = 0.0
        if tag == "m32":
            theta[0] = 0.0
        par = np.r_[theta, 0.8]
        gp = ref_model(kind, d, None, 1e-6, theta0=np.full(d, 0.1), thetaL=1e-5 * np.ones(d),
                       thetaU=100 * np.ones(d))
        llf, grad, _ = ref_set_par(gp, X, y, par)
        out.update({tag + "_X": X, tag + "_y": y, tag + "_par": par, tag + "_llf": llf, tag + "_grad": grad})
    np.savez_compressed(os.path.join(OUT, "zero_theta.npz"), **out)
    print("zero_theta", {k: float(v) for k, v in out.items() if k.endswith("_llf")})


def isotropic_cases():
    """ONE theta shared by all features (kernel.py:312-317, 'theta.size == 1').  Note what the
    reference's likelihood gradient does then (gpr.py:907-944): it indexes its (n, n, d (+1))
    derivative tensor with range(n_par), so entry 0 is the derivative w.r.t. the FIRST feature's
    theta only and, in noisy mode, the sigma2 entry receives the second feature's -- frozen here as
    the reference computes it."""
    case_fixed_par("iso_se_ok_noiseless", seed=301, n=64, d=3, kind="squared_exponential", beta=None,
                   nugget=0, theta_scale=0.5, isotropic=True)
    case_fixed_par("iso_se_ok_noisy", seed=302, n=72, d=4, kind="squared_exponential", beta=None,
                   nugget=1e-6, isotropic=True)
    case_fixed_par("iso_m32_sk_noisy", seed=303, n=64, d=3, kind="matern", beta=0.3, nugget=1e-6,
                   isotropic=True)
    case_fixed_par("iso_absexp_ok_noisy", seed=304, n=48, d=3, kind="absolute_exponential", beta=None,
                   nugget=1e-6, isotropic=True)
    case_fixed_par("iso_se_ok_conc_noiseestim", seed=305, n=64, d=3, kind="squared_exponential", beta=None,
                   nugget=0, noise_estim=True, isotropic=True)
    case_fixed_par("iso_se_ok_reml_noisy", seed=306, n=64, d=3, kind="squared_exponential", beta=None,
                   nugget=1e-6, likelihood="restricted", isotropic=True)
    case_fit("ifit_se_ok_noisy", seed=31, n=40, d=3, kind="squared_exponential", beta=None, nugget=1e-6,
             isotropic=True)


def multi_restart_fits():
    """Reference fits with several L-BFGS-B restarts (gpr.py:1030-1064: sequential restarts, shared
    budget, wait_iter early stop).  Modes whose analytic gradient is exact (noisy; given trend --
    SURVEY Q6), so that the end point of a restart is a property of the likelihood surface and not
    of rounding in the line search: the fixture pins "multi-restart fit reaches at least the
    reference's likelihood" (SURVEY H5)."""
    case_fit("mfit_se_ok_noisy_rs4", seed=21, n=60, d=3, kind="squared_exponential", beta=None,
             nugget=1e-6, random_start=4)
    case_fit("mfit_m32_ok_noisy_rs5", seed=22, n=72, d=4, kind="matern", beta=None, nugget=1e-4,
             random_start=5)
    case_fit("mfit_se_sk_noiseless_rs4", seed=23, n=56, d=3, kind="squared_exponential", beta=0.0,
             nugget=0, random_start=4)


def main():
    os.makedirs(OUT, exist_ok=True)
    load_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "variants":
        variants()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "linear":
        linear_trends()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "zero_theta":
        zero_theta()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "linear_absexp":
        linear_trends_absexp()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "isotropic":
        isotropic_cases()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "multi_restart":
        multi_restart_fits()
        return
    variants()
    linear_trends()
    zero_theta()
    multi_restart_fits()
    isotropic_cases()
    k = 0
    for kind in ("squared_exponential", "matern"):
        for beta in (None, 0.3):
            for nugget in (0, 1e-6):
                k += 1
                tr = "ok" if beta is None else "sk"
                nm = "fixed_%s_%s_%s" % ("se" if kind[0] == "s" else "m32", tr,
                                         "noisy" if nugget else "noiseless")
                case_fixed_par(nm, seed=100 + k, n=64 if k % 2 else 96, d=3 + (k % 3), kind=kind,
                               beta=beta, nugget=nugget)
    case_fixed_par("fixed_absexp_ok_noisy", seed=120, n=48, d=3, kind="absolute_exponential",
                   beta=None, nugget=1e-6)
    case_fit("fit_se_ok_noiseless", seed=11, n=40, d=3, kind="squared_exponential", beta=None,
             nugget=0)
    case_fit("fit_se_ok_noisy", seed=12, n=40, d=3, kind="squared_exponential", beta=None,
             nugget=1e-6)
    case_fit("fit_m32_sk_noiseless", seed=13, n=48, d=4, kind="matern", beta=0.0, nugget=0)
    case_c1()


if __name__ == "__main__":
    main()
