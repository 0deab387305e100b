"""Parity of the CUDA path (through the C ABI / drop-in classes) against the golden fixtures
produced by the reference itself and against the numpy oracle on seeded inputs.

Tolerances (BASELINE.json north_star): relative 1e-9 on the posterior mean, 1e-7 on MSE,
acquisition values and gradients, on well-conditioned models (nugget >= 1e-6), each with an
absolute floor scaled by the natural magnitude (sigma2 for the MSE, max |value| for the others).
With nugget 0 the reference's own formulas evaluated in two orders differ by 4e-9 (mean) /
7e-8 (EI): those fixtures are gated at 1e-6 / 2e-5.
"""
import numpy as np
import pytest

from golden_util import FIXED, FITS, MFITS, VARS, LINS, load, mode_of, beta_of, likelihood_of, trend_of

pytestmark = pytest.mark.gpu


def _mods():
    import mipego_b200 as mb
    from oracle import gp_oracle as O
    return mb, O


def tol(g, tight, loose):
    return tight if float(g["nugget"]) > 0 else loose


def make_gp(mb, g, fitted=True):
    d = g["X"].shape[1]
    beta = None if g["estimate_trend"] else float(g["beta_in"])
    nug = float(g["nugget"])
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=beta), corr=g["kind"],
                            theta0=g["theta"] if "theta" in g else g["theta0"],
                            thetaL=1e-5 * np.ones(d), thetaU=100 * np.ones(d),
                            nugget=nug if nug > 0 else 0)
    if fitted:
        gp.set_fitted_state(g["X"], g["y"], g["theta"], g["L"], g["gamma"], float(g["sigma2"]),
                            float(g["beta"]))
    return gp


SEMAT = list(FIXED)   # squared exponential, Matern-3/2 and absolute exponential fixtures


@pytest.mark.parametrize("name", SEMAT)
def test_predict_golden(name):
    mb, O = _mods()
    g = load(name)
    gp = make_gp(mb, g)
    mean, mse = gp.predict(g["Xs"], eval_MSE=True)
    assert np.max(np.abs(mean - g["mean"]) / (np.abs(g["mean"]) + 1e-3)) < tol(g, 1e-9, 1e-6)
    s2 = float(g["sigma2"])
    assert np.max(np.abs(mse - g["mse"]) / (np.abs(g["mse"]) + 1e-6 * s2)) < tol(g, 1e-7, 1e-5)
    only_mean = gp.predict(g["Xs"])
    assert np.array_equal(only_mean, mean)


@pytest.mark.parametrize("name", SEMAT)
def test_gradient_golden(name):
    mb, O = _mods()
    g = load(name)
    gp = make_gp(mb, g)
    mdx, sdx = gp.gradient_batch(g["Xs"])
    sc1, sc2 = np.abs(g["mean_dx"]).max(), np.abs(g["mse_dx"]).max()
    assert np.max(np.abs(mdx - g["mean_dx"])) < tol(g, 1e-9, 1e-6) * sc1
    assert np.max(np.abs(sdx - g["mse_dx"])) < tol(g, 1e-7, 1e-5) * sc2
    a, b = gp.gradient(g["Xs"][4])
    assert a.shape == (g["X"].shape[1], 1) and b.shape == a.shape
    assert np.allclose(a.ravel(), mdx[4], rtol=0, atol=1e-12 * sc1)


@pytest.mark.parametrize("name", SEMAT)
@pytest.mark.parametrize("crit", ["EI", "PI", "UCB", "MGFI"])
@pytest.mark.parametrize("minimize", [True, False])
def test_acquisition_golden(name, crit, minimize):
    mb, O = _mods()
    g = load(name)
    gp = make_gp(mb, g)
    tag = "min" if minimize else "max"
    plugin = float(g["plugin_" + tag])
    if crit == "EI":
        c = mb.EI(model=gp, minimize=minimize, plugin=plugin)
    elif crit == "PI":
        c = mb.EpsilonPI(model=gp, minimize=minimize, plugin=plugin, epsilon=float(g["param_PI"]))
    elif crit == "UCB":
        c = mb.UCB(model=gp, minimize=minimize, alpha=float(g["param_UCB"]))
    else:
        c = mb.MGFI(model=gp, minimize=minimize, plugin=plugin, t=float(g["param_MGFI"]))
    val, dx = c(g["Xs"], return_dx=True)
    ref_v, ref_dx = g["%s_%s" % (crit, tag)], g["%s_%s_dx" % (crit, tag)]
    fin = np.isfinite(ref_v) & np.all(np.isfinite(ref_dx), axis=1)
    noise = g["mse"] < 1e-10 * float(g["sigma2"])   # nugget-0 training points: cancellation noise
    if crit == "EI":
        assert np.all(val[noise] == 0.0)
    fin &= ~noise
    vs = np.abs(ref_v[fin]).max() + 1e-300
    # candidates sitting (almost) on a training point have 1 - |L^-1 r|^2 ~ nugget: the MSE keeps
    # only ~1e-10 sigma2 absolute accuracy (cond(R) eps) and sd = sqrt(MSE) amplifies it by
    # 1/(2 sd).  Those rows (mse < 1e-4 sigma2) are gated at 1e-5, the rest at the headline 1e-7.
    near = g["mse"] < 1e-4 * float(g["sigma2"])
    rel = np.abs(val - ref_v) / (np.abs(ref_v) + 1e-6 * vs)
    assert np.max(rel[fin & ~near]) < tol(g, 1e-7, 2e-5)
    if (fin & near).any():
        assert np.max(rel[fin & near]) < 1e-5
    rows = np.arange(len(ref_v))
    easy = fin & (rows >= 3)
    ds = np.abs(ref_dx[easy]).max() + 1e-300
    assert np.max(np.abs(dx[easy] - ref_dx[easy])) < tol(g, 1e-7, 2e-5) * ds
    # the reference's single-point protocol and return types
    v1 = c(g["Xs"][5])
    v2, d2 = c(g["Xs"][5], return_dx=True)
    if crit == "PI":
        assert isinstance(v1, np.ndarray) and v1.shape == (1,)
    else:
        assert isinstance(v1, float)
    assert d2.shape == (g["X"].shape[1], 1)
    # value-only calls use |L^-1 r|^2, gradient calls use r^T R^-1 r: equal up to rounding
    assert abs(float(np.ravel(v1)[0]) - val[5]) <= 1e-6 * abs(val[5]) + 1e-300
    # (a single point takes the matrix-vector path for <= 8 candidates: same value up to rounding)
    assert abs(float(np.ravel(v2)[0]) - val[5]) <= 1e-7 * abs(val[5]) + 1e-300


@pytest.mark.parametrize("name", SEMAT)
def test_likelihood_golden(name):
    mb, O = _mods()
    g = load(name)
    gp = make_gp(mb, g, fitted=False)
    gp._check_data(g["X"], g["y"])
    par = g["par"]
    rng = np.random.default_rng(5)
    P = np.vstack([par] + [par * rng.uniform(0.5, 2.0, par.size) for _ in range(4)])
    llf, grad, status = gp.log_likelihood_batch(P, eval_grad=True)
    assert abs(llf[0] - float(g["llf"])) <= 1e-9 * abs(float(g["llf"]))
    gs = np.abs(g["llf_grad"]).max()
    assert np.max(np.abs(grad[0] - g["llf_grad"])) <= tol(g, 1e-7, 1e-6) * gs
    for b in range(1, 5):
        ref = O.log_likelihood_concentrated(g["X"], g["y"], P[b], g["kind"], g["estimate_trend"],
                                            beta_of(g), mode_of(g), float(g["nugget"]), eval_grad=True)
        if np.isinf(ref[0]):
            assert np.isinf(llf[b]) and status[b] != 0 and np.all(grad[b] == 0)
            continue
        assert status[b] == 0
        assert abs(llf[b] - ref[0]) <= 1e-9 * abs(ref[0])
        assert np.max(np.abs(grad[b] - ref[1].ravel())) <= tol(g, 1e-7, 1e-6) * np.abs(ref[1]).max()
    # the env harvested by a likelihood call (gpr.py:876-887): no side effect on the model ...
    env = {}
    llf1 = gp.log_likelihood_concentrated(par, env=env)
    assert abs(llf1 - float(g["llf_env"])) <= 1e-9 * abs(float(g["llf_env"]))
    assert not gp.is_fitted and not gp._fitted_on_device()
    assert np.allclose(env["C"], g["L"], rtol=0, atol=1e-11)
    assert np.max(np.abs(env["rho"] - g["rho"])) <= tol(g, 1e-9, 1e-7) * np.abs(g["rho"]).max()
    assert np.max(np.abs(env["Yt"] - g["Yt"])) <= tol(g, 1e-9, 1e-7) * np.abs(g["Yt"]).max()
    assert abs(float(np.ravel(env["sigma2"])[0]) - float(g["sigma2"])) <= 1e-10 * float(g["sigma2"])
    if g["estimate_trend"]:
        assert np.max(np.abs(env["Ft"] - g["Ft"])) <= 1e-9 * np.abs(g["Ft"]).max()
        assert np.max(np.abs(env["Q"] - g["Q"])) <= 1e-9 * np.abs(g["Q"]).max()
    # ... the fitted state is what fit / fit_fixed build (mgp_finalize_fit)
    gp.fit_fixed(g["X"], g["y"], par)
    assert np.allclose(gp.C, g["L"], rtol=0, atol=1e-11)
    assert abs(float(np.ravel(gp.sigma2)[0]) - float(g["sigma2"])) <= 1e-10 * float(g["sigma2"])
    assert abs(float(np.ravel(gp.mean.beta)[0]) - float(g["beta"])) <= 1e-8 * max(1.0, abs(float(g["beta"])))
    assert np.max(np.abs(gp.gamma - g["gamma"])) <= tol(g, 1e-8, 1e-6) * np.abs(g["gamma"]).max()
    assert np.max(np.abs(gp.rho - g["rho"])) <= tol(g, 1e-9, 1e-7) * np.abs(g["rho"]).max()
    if g["estimate_trend"]:
        assert np.max(np.abs(gp.Ft - g["Ft"])) <= 1e-9 * np.abs(g["Ft"]).max()
        assert abs(float(gp.G[0, 0]) - float(g["G"][0, 0])) <= 1e-10 * abs(float(g["G"][0, 0]))


@pytest.mark.parametrize("name", VARS)
def test_variant_modes_golden(name):
    """REML in its three estimation modes (gpr.py:699-796) and the concentrated likelihood with an
    estimated noise share (gpr.py:837-853, 915-930): likelihood, gradient, harvested state,
    posterior and EI against fixtures produced by the reference."""
    mb, O = _mods()
    g = load(name)
    d = g["X"].shape[1]
    lik, mode = likelihood_of(g), mode_of(g)
    beta = None if g["estimate_trend"] else float(g["beta_in"])
    nug = float(g["nugget"])
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=beta), corr=g["kind"], theta0=g["theta"],
                            thetaL=1e-5 * np.ones(d), thetaU=100 * np.ones(d),
                            nugget=nug if nug > 0 else 0, noise_estim=(mode == "noise_estim"),
                            likelihood=lik)
    assert gp.estimation_mode == mode
    gp._check_data(g["X"], g["y"])
    par = g["par"]
    rng = np.random.default_rng(7)
    scale = np.ones((4, par.size))
    scale[:, :d] = rng.uniform(0.6, 1.6, (4, d))
    P = np.vstack([par, par * scale])
    llf, grad, status = gp.log_likelihood_batch(P, eval_grad=True)
    fn = O.log_likelihood_restricted if lik == "restricted" else O.log_likelihood_concentrated
    ref0 = float(g["llf"])
    if np.isinf(ref0):
        assert np.isinf(llf[0]) and status[0] == -1
    else:
        assert status[0] == 0 and abs(llf[0] - ref0) <= 1e-9 * abs(ref0)
    gs = np.abs(g["llf_grad"]).max()
    assert np.max(np.abs(grad[0] - g["llf_grad"])) <= tol(g, 1e-7, 1e-6) * gs, (grad[0], g["llf_grad"])
    for b in range(1, 5):
        rl, rg = fn(g["X"], g["y"], P[b], g["kind"], g["estimate_trend"], beta_of(g), mode, nug,
                    eval_grad=True)
        if np.isinf(rl):
            assert np.isinf(llf[b])
        else:
            assert abs(llf[b] - rl) <= 1e-9 * abs(rl)
        if lik == "restricted" or not np.isinf(rl):
            assert np.max(np.abs(grad[b] - rg.ravel())) <= tol(g, 1e-7, 1e-6) * np.abs(rg).max()
    # harvested state (the env of the final likelihood call) + posterior + EI
    env = {}
    single = gp.log_likelihood_restricted if lik == "restricted" else gp.log_likelihood_concentrated
    l1, g1 = single(par, env=env, eval_grad=True)
    assert g1.shape == (par.size, 1)
    assert np.allclose(env["C"], g["L"], rtol=0, atol=1e-11)
    assert abs(float(np.ravel(env["sigma2"])[0]) - float(g["sigma2"])) <= 1e-10 * float(g["sigma2"])
    assert abs(float(np.ravel(env["noise_var"])[0]) - float(g["noise_var"])) <= \
        1e-10 * max(float(g["noise_var"]), 1e-300)
    assert not gp.is_fitted                      # the env harvest does not adopt the state
    gp.fit_fixed(g["X"], g["y"], par)
    assert np.max(np.abs(gp.gamma - g["gamma"])) <= tol(g, 1e-8, 1e-6) * np.abs(g["gamma"]).max()
    mean, mse = gp.predict(g["Xs"], eval_MSE=True)
    assert np.max(np.abs(mean - g["mean"]) / (np.abs(g["mean"]) + 1e-3)) < tol(g, 1e-9, 1e-6)
    s2 = float(g["sigma2"])
    assert np.max(np.abs(mse - g["mse"]) / (np.abs(g["mse"]) + 1e-6 * s2)) < tol(g, 1e-7, 1e-5)
    ei = mb.EI(model=gp, minimize=True, plugin=float(g["plugin_min"]))
    v, dx = ei(g["Xs"], return_dx=True)
    easy = g["mse"] > 1e-4 * s2
    rv, rdx = g["EI_min"], g["EI_min_dx"]
    assert np.max(np.abs(v[easy] - rv[easy]) / (np.abs(rv[easy]) + 1e-6 * np.abs(rv).max())) < tol(g, 1e-7, 2e-5)
    assert np.max(np.abs(dx[easy] - rdx[easy])) < tol(g, 1e-7, 2e-5) * np.abs(rdx[easy]).max()


@pytest.mark.parametrize("lik,noise_estim,nugget", [("restricted", False, 0), ("restricted", False, 1e-4),
                                                    ("restricted", True, 0), ("concentrated", True, 0)])
def test_variant_modes_fit(lik, noise_estim, nugget):
    """A full fit in each variant: the MLE found through the batched likelihood is a stationary
    point of the oracle's likelihood at least as good as the oracle's value at the start."""
    mb, O = _mods()
    rng = np.random.default_rng(3)
    n, d = 70, 3
    X = rng.uniform(-5, 5, (n, d))
    y = (X ** 2).sum(1) + 3 * np.sin(X[:, 0]) + 0.05 * rng.standard_normal(n)
    y = (y - y.mean()) / y.std()
    np.random.seed(4)
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="squared_exponential",
                            theta0=[0.05] * d, thetaL=[1e-4] * d, thetaU=[10.0] * d, nugget=nugget,
                            noise_estim=noise_estim, likelihood=lik, random_start=3)
    gp.fit(X, y)
    mode = gp.estimation_mode
    par = np.concatenate([np.ravel(v) for v in gp.par.values()])
    fn = O.log_likelihood_restricted if lik == "restricted" else O.log_likelihood_concentrated
    rl = fn(X, y, par, "squared_exponential", True, 0.0, mode, float(nugget))
    assert abs(float(np.ravel(gp.log_likelihood_)[0]) - rl) <= 1e-8 * abs(rl)
    start = np.r_[[0.05] * d, [0.5] * (par.size - d)]
    assert rl >= fn(X, y, start, "squared_exponential", True, 0.0, mode, float(nugget)) - 1e-9
    st = O.fit_state(X, y, par, "squared_exponential", True, 0.0, mode, float(nugget), likelihood=lik)
    Xs = rng.uniform(-5, 5, (200, d))
    rm, rs = O.predict(st, Xs)
    mean, mse = gp.predict(Xs, eval_MSE=True)
    assert np.max(np.abs(mean - rm) / (np.abs(rm) + 1e-3)) < 1e-8
    assert np.max(np.abs(mse - rs) / (np.abs(rs) + 1e-6 * st["sigma2"])) < 1e-6
    assert abs(float(np.ravel(gp.noise_var)[0]) - st["noise_var"]) <= 1e-9 * max(st["noise_var"], 1e-300)


def test_not_positive_definite_and_positive_llf_status():
    mb, O = _mods()
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (60, 2))
    y = (X ** 2).sum(1)
    gp = mb.GaussianProcess(mean=mb.constant_trend(2, beta=None), corr="squared_exponential",
                            theta0=[1e-8, 1e-8], thetaL=[1e-10] * 2, thetaU=[10] * 2, nugget=0)
    gp._check_data(X, y)
    P = np.array([[1e-9, 1e-9], [5.0, 5.0], [60.0, 60.0]])
    llf, grad, status = gp.log_likelihood_batch(P)
    # (a) singular R: Cholesky fails -> (-inf, 0) with the failing pivot as status (gpr.py:820-826)
    ref0 = O.log_likelihood_concentrated(X, y, P[0], eval_grad=True)
    assert np.isinf(ref0[0]) and np.isinf(llf[0]) and status[0] > 0 and np.all(grad[0] == 0)
    # (b) a smooth noiseless fit has llf > 0, which the reference rejects (gpr.py:889-891)
    ref1 = O.log_likelihood_concentrated(X, y, P[1], eval_grad=True)
    assert np.isinf(ref1[0]) and np.isinf(llf[1]) and status[1] == -1 and np.all(grad[1] == 0)
    # (c) a failing batch element must not disturb its neighbours
    ref2 = O.log_likelihood_concentrated(X, y, P[2], eval_grad=True)
    assert np.isfinite(ref2[0]) and status[2] == 0 and abs(llf[2] - ref2[0]) <= 1e-9 * abs(ref2[0])
    assert np.max(np.abs(grad[2] - ref2[1].ravel())) <= 1e-6 * np.abs(ref2[1]).max()


def test_c1_example_config():
    """BASELINE config 1 (n=50, d=2, EI) against the reference's own per-point outputs.  The
    reference's fit ends at theta = (0.5, 0.5), cond(L) = 20: a well-conditioned model although
    nugget = 0, so the north_star gates apply unrelaxed (1e-9 mean, 1e-7 MSE / EI / EI gradient)."""
    mb, O = _mods()
    g = load("c1_example")
    gp = mb.GaussianProcess(mean=mb.constant_trend(2, beta=None), corr="squared_exponential",
                            theta0=g["theta_"], thetaL=[1e-9] * 2, thetaU=[100] * 2, nugget=0)
    gp.set_fitted_state(g["X"], g["y"], g["theta_"], g["L"], g["gamma"], float(g["sigma2"]), float(g["beta"]))
    mean, mse = gp.predict(g["Xs"], eval_MSE=True)
    s2 = float(g["sigma2"])
    e_mean = np.max(np.abs(mean - g["mean"]) / (np.abs(g["mean"]) + 1e-2))
    e_mse = np.max(np.abs(mse - g["mse"]) / (np.abs(g["mse"]) + 1e-6 * s2))
    ei = mb.EI(model=gp, minimize=True)
    v, dx = ei(g["Xs"], return_dx=True)
    e_ei = np.max(np.abs(v - g["EI"])) / np.abs(g["EI"]).max()
    e_dx = np.max(np.abs(dx - g["EI_dx"].reshape(dx.shape))) / np.abs(g["EI_dx"]).max()
    print("C1 rel err: mean %.2e mse %.2e EI %.2e EI_dx %.2e" % (e_mean, e_mse, e_ei, e_dx))
    assert e_mean < 1e-9 and e_mse < 1e-7 and e_ei < 1e-7 and e_dx < 1e-7
    # the same model built by the library itself from (X, y, theta): its own Cholesky
    gp2 = mb.GaussianProcess(mean=mb.constant_trend(2, beta=None), corr="squared_exponential",
                             theta0=g["theta_"], thetaL=[1e-9] * 2, thetaU=[100] * 2, nugget=0)
    gp2.fit_fixed(g["X"], g["y"], g["theta_"])
    mean2, mse2 = gp2.predict(g["Xs"], eval_MSE=True)
    assert np.max(np.abs(mean2 - g["mean"]) / (np.abs(g["mean"]) + 1e-2)) < 1e-9
    assert np.max(np.abs(mse2 - g["mse"]) / (np.abs(g["mse"]) + 1e-6 * s2)) < 1e-7
    assert abs(float(np.ravel(gp2.log_likelihood_)[0]) - float(g["llf"])) <= 1e-9 * abs(float(g["llf"]))


def _stress_problem(kind, scale, n=200, d=4, seed=50, radius=0.05):
    """SURVEY H1 / section 8d stress set: 25 % of the training points within `radius` of one
    centre (what a BO trajectory does around its incumbent), nugget 0."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(-5, 5, (n, d))
    c = rng.uniform(-3, 3, d)
    k = n // 4
    v = rng.standard_normal((k, d))
    v /= np.linalg.norm(v, axis=1)[:, None]
    X[:k] = c + radius * rng.uniform(0, 1, (k, 1)) ** (1.0 / d) * v
    y = (X ** 2).sum(1) + 3 * np.sin(X[:, 0])
    y = (y - y.mean()) / y.std()
    theta = (0.12 / d) * rng.uniform(0.5, 1.5, d) * scale
    Xs = np.r_[rng.uniform(-5, 5, (200, d)), c + 0.2 * rng.standard_normal((100, d))]
    return X, y, theta, Xs


@pytest.mark.parametrize("kind,scale", [("squared_exponential", 30.0), ("squared_exponential", 100.0),
                                        ("matern", 1.0)])
def test_stress_set_against_extended_precision_truth(kind, scale):
    """Ill-conditioned models (cond(R) up to 2e15) cannot be gated at 1e-9 against the float64
    reference: the reference itself is 1e-3 .. 1e-11 away from the exact posterior there (its
    Cholesky loses cond(R) * eps).  Judged instead against an extended-precision evaluation of
    the same formulas (oracle/gp_truth.py, numpy longdouble): the CUDA path -- its own
    factorisation, explicit L^-1, distance GEMM -- must be no farther from the truth than a small
    multiple of the reference's own distance (both are O(cond * eps) with different constants)."""
    mb, O = _mods()
    from oracle import gp_truth as T
    X, y, theta, Xs = _stress_problem(kind, scale)
    d = X.shape[1]
    st = O.fit_state(X, y, theta, kind, True, 0.0, "noiseless", 0.0)        # the reference's float64 path
    ft = T.fit_truth(X, y, theta, kind)
    tm, ts = (np.asarray(a, dtype=np.float64) for a in T.predict(ft, Xs))
    s2 = float(ft["sigma2"])
    rm, rs = O.predict(st, Xs)
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr=kind, theta0=theta,
                            thetaL=1e-6 * np.ones(d), thetaU=100 * np.ones(d), nugget=0)
    gp.fit_fixed(X, y, theta)
    gm, gs = gp.predict(Xs, eval_MSE=True)

    def dist(m, s):
        return (np.max(np.abs(m - tm) / (np.abs(tm) + 1e-3)), np.max(np.abs(s - ts) / (np.abs(ts) + 1e-6 * s2)))
    ref_e, gpu_e = dist(rm, rs), dist(gm, gs)
    e_s2 = (abs(st["sigma2"] - s2) / s2, abs(float(np.ravel(gp.sigma2)[0]) - s2) / s2)
    print("stress %s x%g: cond(L) %.1e | distance to truth  reference: mean %.2e mse %.2e sigma2 %.2e | "
          "CUDA: mean %.2e mse %.2e sigma2 %.2e" % (kind, scale, np.linalg.cond(st["L"]), ref_e[0], ref_e[1],
                                                    e_s2[0], gpu_e[0], gpu_e[1], e_s2[1]))
    # whole pipeline: both sides are dominated by the cond(R) * eps of their Cholesky factorisation.
    # Measured (B200): SE x30 1.7x / 1.9x the reference's distance (mean / MSE), SE x100 0.9x / 1.5x,
    # Matern 6.7x / 6.8x (at the 1e-11 / 1e-9 level)
    K = 20.0
    assert gpu_e[0] <= K * ref_e[0] + 1e-11 and gpu_e[1] <= K * ref_e[1] + 1e-9
    assert e_s2[1] <= K * e_s2[0] + 1e-12
    # the same state handed over (mgp_set_state): only the posterior arithmetic differs now, and the
    # truth is evaluated from that float64 state taken as exact
    gp2 = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr=kind, theta0=theta,
                             thetaL=1e-6 * np.ones(d), thetaU=100 * np.ones(d), nugget=0)
    gp2.set_fitted_state(X, y, theta, st["L"], st["gamma"], st["sigma2"], st["beta"])
    t2 = T.state_truth(X, theta, st["L"], st["gamma"], st["sigma2"], st["beta"], kind)
    tm, ts = (np.asarray(a, dtype=np.float64) for a in T.predict(t2, Xs))
    s2 = float(st["sigma2"])
    ref_e, gpu_e = dist(rm, rs), dist(*gp2.predict(Xs, eval_MSE=True))
    print("   same state: reference mean %.2e mse %.2e | CUDA mean %.2e mse %.2e" % (ref_e + gpu_e))
    # With the factor handed over, the posterior arithmetic alone is compared, and here the CUDA
    # path IS farther from the truth than the reference, by up to 60x on this set (measured: SE x100
    # mean 5.2e-9 vs 8.6e-11, MSE 1.7e-6 vs 3.1e-8; SE x30 9x / 8x; Matern 1x / 6x): (i) T = L^-1 r
    # goes through the explicit inverse, whose entries carry cond(L) * eps, where the reference's
    # forward substitution is backward stable (SURVEY H1 names this trade); (ii) the distance GEMM
    # has absolute error eps * theta |x|^2 in S where component-wise differences keep a relative
    # one, and r . gamma amplifies it by |r| . |gamma| / |mean| (1e5 on this set).  Both are
    # O(cond * eps) effects, invisible (<= 1e-11) on the well-conditioned headline models.
    K_same = 200.0
    assert gpu_e[0] <= K_same * ref_e[0] + 1e-11 and gpu_e[1] <= K_same * ref_e[1] + 1e-9


def synth_model(O, n, d, kind, seed, nugget=1e-6, estimate_trend=True):
    """SURVEY section 8d synthetic model (C2/C3 recipe)."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(-5, 5, (n, d))
    y = (X ** 2).sum(1) + 3 * np.sin(X[:, 0])
    y = (y - y.mean()) / y.std()
    theta = (0.12 / d) * rng.uniform(0.5, 1.5, d) * (4.0 if kind == "matern" else 1.0)
    st = O.fit_state(X, y, np.r_[theta, 0.9], kind, estimate_trend, 0.2, "noisy", nugget)
    return st


@pytest.mark.parametrize("kind,n,d", [("squared_exponential", 500, 10), ("matern", 500, 10),
                                      ("squared_exponential", 1100, 20)])
def test_oracle_parity_medium(kind, n, d):
    """C2-shaped model: full posterior + every criterion against the oracle on 20k candidates."""
    mb, O = _mods()
    st = synth_model(O, n, d, kind, seed=10)
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr=kind, theta0=st["theta"],
                            thetaL=1e-5 * np.ones(d), thetaU=100 * np.ones(d), nugget=1e-6)
    gp.set_fitted_state(st["X"], st["y"], st["theta"], st["L"], st["gamma"], st["sigma2"], st["beta"])
    m = 20000 if n <= 500 else 3000
    Xs = np.random.default_rng(11).uniform(-5, 5, (m, d))
    Xs[:5] = st["X"][:5] + 1e-4
    mean, mse = gp.predict(Xs, eval_MSE=True)
    rm, rs = O.predict(st, Xs)
    assert np.max(np.abs(mean - rm) / (np.abs(rm) + 1e-3)) < 1e-9
    assert np.max(np.abs(mse - rs) / (np.abs(rs) + 1e-6 * st["sigma2"])) < 1e-7
    plugin = float(st["y"].min())
    for crit, cls, kw, par in (("EI", mb.EI, {}, None), ("PI", mb.EpsilonPI, {"epsilon": 0.01}, 0.01),
                               ("UCB", mb.UCB, {"alpha": 0.5}, 0.5), ("MGFI", mb.MGFI, {"t": 2.0}, 2.0)):
        if crit != "UCB":
            kw = dict(kw, plugin=plugin)
        c = cls(model=gp, minimize=True, **kw)
        v = c(Xs)
        rv = O.acquisition(st, Xs, crit, par, plugin, True)
        vs = np.abs(rv).max()
        assert np.max(np.abs(v - rv) / (np.abs(rv) + 1e-6 * vs)) < 1e-7, crit
    # gradients on a subset
    sub = Xs[:512]
    c = mb.EI(model=gp, minimize=True, plugin=plugin)
    v, dx = c(sub, return_dx=True)
    rv, rdx = O.acquisition(st, sub, "EI", None, plugin, True, return_dx=True)
    assert np.max(np.abs(dx - rdx)) < 1e-7 * np.abs(rdx).max()
    # multi-parameter evaluation shares one posterior pass (ParallelBO, BayesOpt.py:56-101)
    ts = np.exp(np.log(2) + 0.5 * np.random.default_rng(22).standard_normal(8))
    mg = mb.MGFI(model=gp, minimize=True, plugin=plugin, t=1.0)
    V = mg.evaluate(sub, params=ts)
    for i, t in enumerate(ts):
        rv = O.acquisition(st, sub, "MGFI", t, plugin, True)
        assert np.max(np.abs(V[i] - rv) / (np.abs(rv) + 1e-6 * np.abs(rv).max())) < 1e-7
    tv, ti = mg.topk(Xs, k=5, params=ts)
    Vfull = mg.evaluate(Xs, params=ts)
    for i in range(len(ts)):
        order = np.lexsort((np.arange(m), -Vfull[i]))[:5]
        assert np.array_equal(ti[i], order)
        assert np.array_equal(tv[i], Vfull[i][order])


@pytest.mark.parametrize("kind,mode,n,d", [("squared_exponential", "noisy", 1100, 20),
                                           ("matern", "noiseless", 1300, 7),
                                           ("squared_exponential", "noiseless", 640, 8),
                                           ("matern", "noisy", 300, 64)])
def test_likelihood_oracle_larger_n(kind, mode, n, d):
    """Batched llf + gradient against the oracle at sizes that exercise several Cholesky panels
    (npad/128 = 9, 11, 5: full and partial panels), the recursive triangular inverse with a ragged
    last pair, and every operand layout of the GEMM kernel; ordinary and simple kriging."""
    mb, O = _mods()
    rng = np.random.default_rng(5)
    X = rng.uniform(-5, 5, (n, d))
    y = (X ** 2).sum(1) + 3 * np.sin(X[:, 0])
    y = (y - y.mean()) / y.std()
    nug = 1e-6 if mode == "noisy" else 0.0
    for estimate_trend in (True, False):
        beta = None if estimate_trend else 0.1
        gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=beta), corr=kind,
                                theta0=0.1 * np.ones(d), thetaL=1e-5 * np.ones(d),
                                thetaU=100 * np.ones(d), nugget=nug if nug > 0 else 0)
        gp._check_data(X, y)
        scale = (0.3 / d) * (4.0 if kind == "matern" else 1.0)
        B = 5
        P = scale * 10.0 ** rng.uniform(-0.4, 0.4, (B, d))
        if mode == "noisy":
            P = np.c_[P, rng.uniform(0.5, 1.0, B)]
        llf, grad, status = gp.log_likelihood_batch(P, eval_grad=True)
        assert np.all(status == 0)
        for b in range(B):
            rl, rg = O.log_likelihood_concentrated(X, y, P[b], kind, estimate_trend, 0.1, mode, nug,
                                                   eval_grad=True)
            rl = float(np.ravel(rl)[0])
            assert abs(llf[b] - rl) < 1e-8 * abs(rl) + 1e-7, (b, llf[b], rl)
            rg = np.ravel(rg)
            assert np.max(np.abs(grad[b] - rg)) < 1e-6 * np.abs(rg).max() + 1e-7, (b, grad[b], rg)


@pytest.mark.parametrize("kind", ["squared_exponential", "matern", "absolute_exponential"])
def test_likelihood_with_zero_theta(kind):
    """A theta_i of exactly 0 cannot be folded into the staged coordinates (the R build and the
    gradient kernel scale them by sqrt(theta_i) or theta_i): such vectors take the unscaled loops.
    Both branches in one batch, llf and the FULL gradient (including d/d theta_i at theta_i = 0)
    against the oracle."""
    mb, O = _mods()
    rng = np.random.default_rng(23)
    n, d = 300, 5
    X = rng.uniform(-5, 5, (n, d))
    y = (X ** 2).sum(1) + 3 * np.sin(X[:, 0])
    y = (y - y.mean()) / y.std()
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr=kind, theta0=0.1 * np.ones(d),
                            thetaL=1e-5 * np.ones(d), thetaU=100 * np.ones(d), nugget=1e-6)
    gp._check_data(X, y)
    scale = 0.3 / d
    P = scale * 10.0 ** rng.uniform(-0.3, 0.3, (4, d))
    P[1, 2] = 0.0
    P[3, 0] = 0.0
    P[3, 4] = 0.0
    P = np.c_[P, rng.uniform(0.5, 1.0, 4)]
    llf, grad, status = gp.log_likelihood_batch(P, eval_grad=True)
    assert np.all(status == 0)
    for b in range(4):
        rl, rg = O.log_likelihood_concentrated(X, y, P[b], kind, True, 0.0, "noisy", 1e-6, eval_grad=True)
        rl = float(np.ravel(rl)[0])
        rg = np.ravel(rg)
        assert abs(llf[b] - rl) < 1e-8 * abs(rl) + 1e-7, (b, llf[b], rl)
        assert np.max(np.abs(grad[b] - rg)) < 1e-6 * np.abs(rg).max() + 1e-7, (b, grad[b], rg)


def test_ragged_and_tiny_inputs():
    mb, O = _mods()
    st = synth_model(O, 130, 3, "squared_exponential", seed=3)
    gp = mb.GaussianProcess(mean=mb.constant_trend(3, beta=None), corr="squared_exponential",
                            theta0=st["theta"], thetaL=1e-5 * np.ones(3), thetaU=100 * np.ones(3), nugget=1e-6)
    gp.set_fitted_state(st["X"], st["y"], st["theta"], st["L"], st["gamma"], st["sigma2"], st["beta"])
    rng = np.random.default_rng(1)
    for m in (1, 7, 127, 128, 129, 1000):
        Xs = rng.uniform(-5, 5, (m, 3))
        mean, mse = gp.predict(Xs, eval_MSE=True)
        rm, rs = O.predict(st, Xs)
        assert np.max(np.abs(mean - rm) / (np.abs(rm) + 1e-3)) < 1e-9
        assert np.max(np.abs(mse - rs) / (np.abs(rs) + 1e-6 * st["sigma2"])) < 1e-7
    assert gp.predict(np.empty((0, 3))).shape == (0,)
    with pytest.raises(ValueError):
        gp.predict(np.zeros((4, 5)))
    # chunking must not change results
    gp._handle().set_option("chunk", 256)
    Xs = rng.uniform(-5, 5, (1000, 3))
    a = gp.predict(Xs, eval_MSE=True)
    gp._handle().set_option("chunk", 0)
    b = gp.predict(Xs, eval_MSE=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_duplicate_rows_and_errors():
    mb, O = _mods()
    X = np.random.default_rng(0).uniform(-1, 1, (20, 2))
    X[7] = X[3]
    gp = mb.GaussianProcess(mean=mb.constant_trend(2, beta=None), corr="matern", theta0=[0.1, 0.1],
                            thetaL=[1e-3] * 2, thetaU=[10] * 2)
    with pytest.raises(Exception, match="Multiple input features"):
        gp.fit(X, X.sum(1))
    with pytest.raises(NotImplementedError):
        mb.GaussianProcess(mean=mb.constant_trend(2), corr="cubic", theta0=[0.1] * 2, thetaL=[1e-3] * 2, thetaU=[1] * 2)
    with pytest.raises(ValueError):
        mb.GaussianProcess(mean=mb.constant_trend(2), theta0=[0.1] * 2, thetaL=[1e-3] * 2, thetaU=[np.inf] * 2)


@pytest.mark.parametrize("name", FITS)
def test_fit_against_reference_fit(name):
    """End-to-end fit(): same start, same optimiser (scipy L-BFGS-B, sequential restarts), against
    a fit() of the unmodified reference.  A rounding-level perturbation of (llf, grad) moves the
    end point of L-BFGS-B by about 1e-11 in likelihood on these problems (measured with the oracle),
    so the fitted likelihood is gated at 1e-6 relative and theta at 1e-3."""
    mb, O = _mods()
    g = load(name)
    d = g["X"].shape[1]
    beta = None if g["estimate_trend"] else float(g["beta_in"])
    nug = float(g["nugget"])
    np.random.seed(int(g["seed"]))
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=beta), corr=g["kind"], theta0=g["theta0"],
                            thetaL=g["thetaL"], thetaU=g["thetaU"], nugget=nug if nug > 0 else 0,
                            random_start=int(g["random_start"]), restart_mode="sequential")
    gp.fit(g["X"], g["y"])
    llf = float(np.ravel(gp.log_likelihood_)[0])
    print("fit %s: llf %.12g (reference %.12g), %d evaluations (reference %d)" %
          (name, llf, float(g["llf"]), gp.eval_count, int(g["eval_count"])))
    # In noiseless mode with an estimated trend the reference's gradient is not the derivative of its
    # likelihood: L-BFGS-B then stops where a line search fails, and rounding decides where.
    # With the oracle (CPU) and a relative perturbation of 1e-15 on (llf, grad), 16 runs of this very fit end at
    # -13.8917 (the reference's end point) or at -0.334 after 130..230 evaluations; the CUDA path, whose
    # likelihood agrees with the oracle to 1e-9 at every point (checked below), ended at -13.8678 with the first
    # diagonal-block kernel and ends at -13.8958 with the second: gated at 2e-3 relative for this fixture only.
    q6 = g["estimate_trend"] and nug == 0
    assert llf >= float(g["llf"]) - (2e-3 if q6 else 1e-6) * abs(float(g["llf"]))
    par_end = np.concatenate([np.ravel(v) for v in gp.par.values()])
    l_or = O.log_likelihood_concentrated(g["X"], g["y"], par_end, g["kind"], g["estimate_trend"], beta_of(g),
                                         mode_of(g), nug)
    assert abs(llf - l_or) <= 1e-9 * abs(l_or)
    # at the reference's own optimum the CUDA likelihood agrees to 1e-9
    l_ref, _, st_ref = gp.log_likelihood_batch(g["par"], eval_grad=False)
    assert st_ref[0] == 0
    assert abs(l_ref[0] - float(g["llf"])) <= 1e-9 * abs(float(g["llf"]))
    # With an exact gradient (noisy mode, given trend) the end points agree to 1e-3 and better; the Q6 fixture
    # (see above) within 1 %.
    assert np.allclose(gp.theta_, g["theta_"], rtol=1e-2 if q6 else 1e-3)
    mean, mse = gp.predict(g["Xs"], eval_MSE=True)
    assert np.max(np.abs(mean - g["mean"]) / (np.abs(g["mean"]) + 1e-2)) < (1e-2 if q6 else 1e-4)


@pytest.mark.parametrize("restart_mode", ["sequential", "batched"])
@pytest.mark.parametrize("name", MFITS)
def test_multi_restart_fit_reaches_reference_likelihood(name, restart_mode):
    """Multi-restart fit() against multi-restart fit() of the unmodified reference
    (oracle/gen_golden.py multi_restart_fits: random_start = 4..5, same numpy seed, hence the same
    restart points, gpr.py:1035-1036).  SURVEY H5: lock-step batching keeps every trajectory but
    not the "which restarts ran" sequence, so the contract is final llf >= reference - 1e-6 |ref|
    in both restart modes."""
    mb, O = _mods()
    g = load(name)
    d = g["X"].shape[1]
    beta = None if g["estimate_trend"] else float(g["beta_in"])
    nug = float(g["nugget"])
    np.random.seed(int(g["seed"]))
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=beta), corr=g["kind"], theta0=g["theta0"],
                            thetaL=g["thetaL"], thetaU=g["thetaU"], nugget=nug if nug > 0 else 0,
                            random_start=int(g["random_start"]), restart_mode=restart_mode)
    gp.fit(g["X"], g["y"])
    llf = float(np.ravel(gp.log_likelihood_)[0])
    print("multi-restart fit %s [%s]: llf %.12g (reference %.12g), %d evaluations (reference %d)" %
          (name, restart_mode, llf, float(g["llf"]), gp.eval_count, int(g["eval_count"])))
    assert llf >= float(g["llf"]) - 1e-6 * abs(float(g["llf"]))
    same_end = abs(llf - float(g["llf"])) <= 1e-6 * abs(float(g["llf"]))
    if restart_mode == "sequential" and same_end:
        # the reference's own loop: same restarts, same end point (the evaluation COUNT may differ by a few
        # line-search steps: 260 / 263, 226 / 206 measured).  A STRICTLY better likelihood is accepted too: on
        # mfit_m32_ok_noisy_rs5 a rounding-level change of the gradient's summation order lets one restart leave
        # the basin the reference stops in (llf -36.47 after 401 evaluations against the reference's -89.82
        # after 263, measured on B200) -- the contract is "no worse than the reference", asserted above.
        assert np.allclose(gp.theta_, g["theta_"], rtol=1e-3)
    mean, mse = gp.predict(g["Xs"], eval_MSE=True)
    if same_end:
        assert np.max(np.abs(mean - g["mean"]) / (np.abs(g["mean"]) + 1e-2)) < 1e-4


def test_batched_restarts_and_pickle():
    mb, O = _mods()
    import pickle
    rng = np.random.default_rng(4)
    n, d = 90, 3
    X = rng.uniform(-5, 5, (n, d))
    y = (X ** 2).sum(1) + rng.normal(0, 0.1, n)
    y = (y - y.mean()) / y.std()
    kw = dict(corr="matern", theta0=[0.05] * d, thetaL=[1e-4] * d, thetaU=[10] * d, nugget=1e-4, random_start=6)
    np.random.seed(3)
    g1 = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), restart_mode="sequential", **kw).fit(X, y)
    np.random.seed(3)
    g2 = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), restart_mode="batched", **kw).fit(X, y)
    l1, l2 = float(np.ravel(g1.log_likelihood_)[0]), float(np.ravel(g2.log_likelihood_)[0])
    assert l2 >= l1 - 1e-8 * abs(l1)
    # the model the oracle builds from the fitted parameters predicts the same
    par = np.r_[g2.theta_, float(np.ravel(g2.par["sigma2"])[0])]
    st = O.fit_state(X, y, par, "matern", True, 0.0, "noisy", 1e-4)
    Xs = rng.uniform(-5, 5, (300, d))
    rm, rs = O.predict(st, Xs)
    mean, mse = g2.predict(Xs, eval_MSE=True)
    assert np.max(np.abs(mean - rm) / (np.abs(rm) + 1e-3)) < 1e-8
    assert np.max(np.abs(mse - rs) / (np.abs(rs) + 1e-6 * st["sigma2"])) < 1e-6
    g3 = pickle.loads(pickle.dumps(g2))
    m3, s3 = g3.predict(Xs, eval_MSE=True)
    assert np.max(np.abs(m3 - mean)) <= 1e-10 * np.abs(mean).max()
    ei = mb.EI(model=g3, minimize=True)
    ei2 = pickle.loads(pickle.dumps(ei))
    assert np.allclose(ei2(Xs[:10]), ei(Xs[:10]), rtol=1e-9, atol=1e-300)


def test_bo_shaped_loop():
    """The call pattern of the reference's BO loop (base.py:375-463 + optimizer/utils.py:25-52):
    fit, then L-BFGS-B restarts on criterion(x, return_dx=True), one point per call."""
    mb, O = _mods()
    from scipy.optimize import fmin_l_bfgs_b
    dim, lb, ub = 3, -1.0, 5.0
    rng = np.random.default_rng(123)

    def fitness(x):
        return float(np.sum(np.asarray(x) ** 2))

    X = rng.uniform(lb, ub, (8, dim))
    y = np.array([fitness(x) for x in X])
    thetaL, thetaU = 1e-10 * (ub - lb) * np.ones(dim), 10 * (ub - lb) * np.ones(dim)
    np.random.seed(123)
    model = mb.GaussianProcess(mean=mb.constant_trend(dim, beta=None), corr="squared_exponential",
                               theta0=rng.uniform(size=dim) * (thetaU - thetaL) + thetaL, thetaL=thetaL,
                               thetaU=thetaU, nugget=1e-6, optimizer="BFGS", wait_iter=3, random_start=dim,
                               likelihood="concentrated", eval_budget=100 * dim)
    assert hasattr(model, "gradient") and hasattr(mb.EI, "plugin") and not hasattr(mb.UCB, "plugin")
    for it in range(4):
        model.fit(X, y)
        assert model.is_fitted
        crit = mb.EI(model=model, minimize=True, plugin=float(y.min()))
        best_x, best_f = None, -np.inf
        for r in range(3):
            x0 = rng.uniform(lb, ub, dim)
            func = lambda x: tuple(map(lambda v: -1.0 * v, crit(x, return_dx=True)))
            xo, fo, info = fmin_l_bfgs_b(func, x0, pgtol=1e-8, factr=1e6, bounds=[(lb, ub)] * dim, maxfun=60)
            if -float(fo) > best_f:
                best_x, best_f = xo, -float(fo)
        X = np.vstack([X, best_x])
        y = np.r_[y, fitness(best_x)]
    assert y.min() <= np.sort(y[:8])[0]


@pytest.mark.parametrize("name", LINS)
def test_linear_trend_golden(name):
    """linear_trend (trend.py:90-110), coefficients estimated (universal kriging, p = d + 1) or
    given: likelihood + gradient, harvested state, posterior, posterior gradient and EI with its
    gradient against fixtures produced by the reference."""
    mb, O = _mods()
    g = load(name)
    d = g["X"].shape[1]
    lik, mode = likelihood_of(g), mode_of(g)
    est = g["estimate_trend"]
    beta = None if est else np.asarray(g["beta_in"], dtype=np.float64)
    nug = float(g["nugget"])
    gp = mb.GaussianProcess(mean=mb.linear_trend(d, beta=beta), corr=g["kind"], theta0=g["theta"],
                            thetaL=1e-5 * np.ones(d), thetaU=100 * np.ones(d),
                            nugget=nug if nug > 0 else 0, likelihood=lik)
    gp._check_data(g["X"], g["y"])
    par = g["par"]
    rng = np.random.default_rng(9)
    scale = np.ones((3, par.size))
    scale[:, :d] = rng.uniform(0.7, 1.5, (3, d))
    P = np.vstack([par, par * scale])
    llf, grad, status = gp.log_likelihood_batch(P, eval_grad=True)
    fn = O.log_likelihood_restricted if lik == "restricted" else O.log_likelihood_concentrated
    for b in range(4):
        rl, rg = fn(g["X"], g["y"], P[b], g["kind"], est, beta_of(g), mode, nug, eval_grad=True,
                    trend="linear")
        if b == 0:
            assert (np.isinf(rl) and np.isinf(float(g["llf"]))) or abs(rl - float(g["llf"])) <= 1e-9 * abs(rl)
        if np.isinf(rl):
            assert np.isinf(llf[b])
        else:
            assert abs(llf[b] - rl) <= 1e-9 * abs(rl), (b, llf[b], rl)
        if np.abs(rg).max() > 0:
            assert np.max(np.abs(grad[b] - rg.ravel())) <= tol(g, 1e-7, 1e-6) * np.abs(rg).max(), (b, grad[b], rg.ravel())
    env = {}
    single = gp.log_likelihood_restricted if lik == "restricted" else gp.log_likelihood_concentrated
    single(par, env=env)
    assert np.allclose(env["C"], g["L"], rtol=0, atol=1e-11)
    assert abs(float(np.ravel(env["sigma2"])[0]) - float(g["sigma2"])) <= 1e-9 * float(g["sigma2"])
    if est:
        assert np.max(np.abs(env["Ft"] - g["Ft"])) <= 1e-9 * np.abs(g["Ft"]).max()
        assert np.max(np.abs(env["G"] - g["G"])) <= 1e-8 * np.abs(g["G"]).max()
    gp.fit_fixed(g["X"], g["y"], par)
    assert np.max(np.abs(np.ravel(gp.mean.beta) - np.ravel(g["beta"]))) <= tol(g, 1e-8, 1e-6) * max(1.0, np.abs(g["beta"]).max())
    assert np.max(np.abs(gp.gamma - g["gamma"])) <= tol(g, 1e-8, 1e-6) * np.abs(g["gamma"]).max()
    assert np.max(np.abs(gp.rho - g["rho"])) <= tol(g, 1e-8, 1e-6) * np.abs(g["rho"]).max()
    if est:
        assert gp.Ft.shape == g["Ft"].shape
        assert np.max(np.abs(gp.Ft - g["Ft"])) <= 1e-9 * np.abs(g["Ft"]).max()
        assert np.max(np.abs(gp.G - g["G"])) <= 1e-8 * np.abs(g["G"]).max()     # same LAPACK QR signs
    mean, mse = gp.predict(g["Xs"], eval_MSE=True)
    assert np.max(np.abs(mean - g["mean"]) / (np.abs(g["mean"]) + 1e-3)) < tol(g, 1e-9, 1e-6)
    s2 = float(g["sigma2"])
    assert np.max(np.abs(mse - g["mse"]) / (np.abs(g["mse"]) + 1e-6 * s2)) < tol(g, 1e-7, 1e-5)
    mdx, sdx = gp.gradient_batch(g["Xs"])
    assert np.max(np.abs(mdx - g["mean_dx"])) < tol(g, 1e-8, 1e-6) * np.abs(g["mean_dx"]).max()
    assert np.max(np.abs(sdx - g["mse_dx"])) < tol(g, 1e-7, 1e-5) * np.abs(g["mse_dx"]).max()
    ei = mb.EI(model=gp, minimize=True, plugin=float(g["plugin_min"]))
    v, dx = ei(g["Xs"], return_dx=True)
    easy = g["mse"] > 1e-4 * s2
    rv, rdx = g["EI_min"], g["EI_min_dx"]
    assert np.max(np.abs(v[easy] - rv[easy]) / (np.abs(rv[easy]) + 1e-6 * np.abs(rv).max())) < tol(g, 1e-7, 2e-5)
    assert np.max(np.abs(dx[easy] - rdx[easy])) < tol(g, 1e-7, 2e-5) * np.abs(rdx[easy]).max()
    # pickle round trip keeps the coefficient vector
    import pickle
    g2 = pickle.loads(pickle.dumps(gp))
    m2, s2_ = g2.predict(g["Xs"], eval_MSE=True)
    assert np.max(np.abs(m2 - mean)) <= 1e-9 * np.abs(mean).max() and np.max(np.abs(s2_ - mse)) <= 1e-7 * np.abs(mse).max() + 1e-12 * s2


def test_linear_trend_fit_and_larger_n():
    """Universal kriging at n = 700, d = 20 (p = 21: three 8-column trend tiles, several training
    blocks): a full fit, then the oracle model for the fitted hyper-parameters."""
    mb, O = _mods()
    rng = np.random.default_rng(12)
    n, d = 700, 20
    X = rng.uniform(-5, 5, (n, d))
    y = (X ** 2).sum(1) + 3 * np.sin(X[:, 0]) + 4 * X[:, 2] - 2 * X[:, 7]
    y = (y - y.mean()) / y.std()
    np.random.seed(2)
    gp = mb.GaussianProcess(mean=mb.linear_trend(d, beta=None), corr="squared_exponential",
                            theta0=[0.01] * d, thetaL=[1e-4] * d, thetaU=[10.0] * d, nugget=1e-6,
                            random_start=2, eval_budget=60)
    gp.fit(X, y)
    par = np.concatenate([np.ravel(v) for v in gp.par.values()])
    st = O.fit_state(X, y, par, "squared_exponential", True, 0.0, "noisy", 1e-6, trend="linear")
    assert abs(float(np.ravel(gp.log_likelihood_)[0]) - st["llf"]) <= 1e-8 * abs(st["llf"])
    assert np.max(np.abs(np.ravel(gp.mean.beta) - np.ravel(st["beta"]))) < 1e-7 * max(1.0, np.abs(st["beta"]).max())
    Xs = rng.uniform(-5, 5, (3000, d))
    rm, rs = O.predict(st, Xs)
    mean, mse = gp.predict(Xs, eval_MSE=True)
    assert np.max(np.abs(mean - rm) / (np.abs(rm) + 1e-3)) < 1e-8
    assert np.max(np.abs(mse - rs) / (np.abs(rs) + 1e-6 * st["sigma2"])) < 1e-6
    sub = Xs[:300]
    ei = mb.EI(model=gp, minimize=True, plugin=float(y.min()))
    v, dx = ei(sub, return_dx=True)
    rv, rdx = O.acquisition(st, sub, "EI", None, float(y.min()), True, return_dx=True)
    assert np.max(np.abs(v - rv) / (np.abs(rv) + 1e-6 * np.abs(rv).max())) < 1e-6
    assert np.max(np.abs(dx - rdx)) < 1e-6 * np.abs(rdx).max()


def test_update_without_reoptimisation():
    """update(X, y, reoptimize=False): same hyper-parameters, new data -> the model the oracle builds
    for those hyper-parameters on the extended data (the refit between MLE rounds of a BO loop)."""
    mb, O = _mods()
    rng = np.random.default_rng(8)
    d = 4
    X = rng.uniform(-5, 5, (150, d))
    y = (X ** 2).sum(1) + 3 * np.sin(X[:, 0])
    np.random.seed(8)
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="matern", theta0=[0.1] * d,
                            thetaL=[1e-4] * d, thetaU=[10.0] * d, nugget=1e-5, random_start=2)
    gp.fit(X[:140], y[:140])
    par = np.concatenate([np.ravel(v) for v in gp.par.values()])
    gp.update(X, y, reoptimize=False)
    assert gp.X.shape[0] == 150 and np.array_equal(np.concatenate([np.ravel(v) for v in gp.par.values()]), par)
    st = O.fit_state(X, y, par, "matern", True, 0.0, "noisy", 1e-5)
    Xs = rng.uniform(-5, 5, (500, d))
    rm, rs = O.predict(st, Xs)
    mean, mse = gp.predict(Xs, eval_MSE=True)
    assert np.max(np.abs(mean - rm) / (np.abs(rm) + 1e-3)) < 1e-9
    assert np.max(np.abs(mse - rs) / (np.abs(rs) + 1e-6 * st["sigma2"])) < 1e-7
    gp.update(X, y)        # default: full refit, like the reference
    assert gp.is_fitted and gp.X.shape[0] == 150


@pytest.mark.parametrize("kind,n,d", [("squared_exponential", 500, 10), ("matern", 1100, 7)])
def test_small_population_path(kind, n, d):
    """<= 8 candidates per call (the reference's one-point-per-call pattern) take a matrix-vector
    path: same results as the MMA path to rounding, and as the oracle within the usual gates."""
    mb, O = _mods()
    st = synth_model(O, n, d, kind, seed=3)
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr=kind, theta0=st["theta"],
                            thetaL=1e-5 * np.ones(d), thetaU=100 * np.ones(d), nugget=1e-6)
    gp.set_fitted_state(st["X"], st["y"], st["theta"], st["L"], st["gamma"], st["sigma2"], st["beta"])
    Xs = np.random.default_rng(4).uniform(-5, 5, (8, d))
    plugin = float(st["y"].min())
    ei = mb.EI(model=gp, minimize=True, plugin=plugin)
    h = gp._handle()
    for m in (1, 3, 8):
        sub = Xs[:m]
        h.set_option("small_path", 1)
        mean, mse = gp.predict(sub, eval_MSE=True)
        v, dx = ei(sub, return_dx=True) if m > 1 else ei.evaluate(sub, return_dx=True)
        v, dx = np.ravel(v), np.reshape(dx, (m, d))
        h.set_option("small_path", 0)
        mean2, mse2 = gp.predict(sub, eval_MSE=True)
        v2, dx2 = ei.evaluate(sub, return_dx=True)
        h.set_option("small_path", 1)
        assert np.array_equal(mean, mean2)
        assert np.max(np.abs(mse - mse2)) <= 1e-10 * st["sigma2"]
        assert np.max(np.abs(v - v2[0])) <= 1e-9 * np.abs(v2).max() + 1e-300
        assert np.max(np.abs(dx - dx2[0])) <= 1e-8 * np.abs(dx2).max() + 1e-300
        rm, rs = O.predict(st, sub)
        assert np.max(np.abs(mean - rm) / (np.abs(rm) + 1e-3)) < 1e-9
        assert np.max(np.abs(mse - rs) / (np.abs(rs) + 1e-6 * st["sigma2"])) < 1e-7
        rv, rdx = O.acquisition(st, sub, "EI", None, plugin, True, return_dx=True)
        assert np.max(np.abs(v - rv) / (np.abs(rv) + 1e-6 * np.abs(rv).max())) < 1e-7
        assert np.max(np.abs(dx - rdx)) < 1e-7 * np.abs(rdx).max()


def test_fit_with_cma_optimizer():
    """optimizer='CMA' (gpr.py:1066-1085; test/test_BO.py:133-165): a derivative-free MLE whose
    populations are evaluated by the batched likelihood; it must reach a likelihood at least as good
    as the start and produce the oracle's model for the parameters it returns."""
    mb, O = _mods()
    rng = np.random.default_rng(21)
    n, d = 90, 3
    X = rng.uniform(-5, 5, (n, d))
    y = (X ** 2).sum(1) + 3 * np.sin(X[:, 0])
    y = (y - y.mean()) / y.std()
    np.random.seed(5)
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="matern", theta0=[0.05] * d,
                            thetaL=[1e-4] * d, thetaU=[10.0] * d, nugget=1e-6, optimizer="CMA",
                            eval_budget=400)
    gp.fit(X, y)
    par = np.concatenate([np.ravel(v) for v in gp.par.values()])
    llf = float(np.ravel(gp.log_likelihood_)[0])
    rl = O.log_likelihood_concentrated(X, y, par, "matern", True, 0.0, "noisy", 1e-6)
    assert abs(llf - rl) <= 1e-8 * abs(rl)
    np.random.seed(5)
    bf = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="matern", theta0=[0.05] * d,
                            thetaL=[1e-4] * d, thetaU=[10.0] * d, nugget=1e-6, optimizer="BFGS",
                            random_start=1).fit(X, y)
    start = O.log_likelihood_concentrated(X, y, np.r_[[0.05] * d, par[-1]], "matern", True, 0.0, "noisy", 1e-6)
    assert llf >= start - 1e-9 and gp.eval_count <= 400 + 16
    assert llf >= float(np.ravel(bf.log_likelihood_)[0]) - 25.0     # same basin or better, loosely
    with pytest.raises(ValueError):
        mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), theta0=[0.1] * d, thetaL=[1e-4] * d,
                           thetaU=[10.0] * d, optimizer="SGD")


def test_edge_shapes_and_values():
    """d = 1 and d = 64 (the supported maximum; 65 is rejected), n = 3 / 128 / 129, candidates on
    training points, far outside the data, non-finite, top-k with k > m and with ties."""
    mb, O = _mods()
    rng = np.random.default_rng(17)
    for n, d in ((3, 1), (128, 1), (129, 2), (200, 64)):
        X = rng.uniform(-5, 5, (n, d))
        y = np.sin(X[:, 0]) + 0.1 * (X ** 2).sum(1) / d
        theta = np.full(d, 0.3 / d)
        par = np.r_[theta, 0.7]
        gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="matern", theta0=theta,
                                thetaL=1e-5 * np.ones(d), thetaU=100 * np.ones(d), nugget=1e-6)
        gp.fit_fixed(X, y, par)
        st = O.fit_state(X, y, par, "matern", True, 0.0, "noisy", 1e-6)
        Xs = np.vstack([rng.uniform(-5, 5, (40, d)), X[:2], X[:1] + 1e3, -1e6 * np.ones((1, d))])
        mean, mse = gp.predict(Xs, eval_MSE=True)
        rm, rs = O.predict(st, Xs)
        assert np.max(np.abs(mean - rm) / (np.abs(rm) + 1e-3)) < 1e-9, (n, d)
        assert np.max(np.abs(mse - rs) / (np.abs(rs) + 1e-6 * st["sigma2"])) < 1e-7, (n, d)
        assert np.all(np.isfinite(mean)) and np.all(np.isfinite(mse))
        ei = mb.EI(model=gp, minimize=True)
        v, dx = ei(Xs, return_dx=True)
        rv, rdx = O.acquisition(st, Xs, "EI", None, None, True, return_dx=True)
        far = np.arange(len(Xs)) >= 42
        assert np.max(np.abs(v - rv) / (np.abs(rv) + 1e-6 * np.abs(rv).max() + 1e-300)) < 1e-6, (n, d)
        assert np.all(np.isfinite(dx))
        # absolute floor: for n = 128, d = 1 every EI value is ~1e-10 (z ~ -6), so the gradients
        # themselves sit at 1e-10 and carry the rounding of the fitted state amplified by |z|
        floor = 1e-13 * np.sqrt(float(np.ravel(st["sigma2"])[0]))
        assert np.max(np.abs(dx[~far] - rdx[~far])) < 1e-5 * np.abs(rdx[~far]).max() + floor, (n, d)
        # top-k: k larger than the population, and exact ties resolved towards the smaller index
        tv, ti = ei.topk(Xs[:5], k=9)
        assert tv.shape == (1, 5) and sorted(ti[0].tolist()) == [0, 1, 2, 3, 4]
        dup = np.vstack([Xs[:3], Xs[:3]])
        tv, ti = ei.topk(dup, k=6)
        vd = ei(dup)
        order = np.lexsort((np.arange(6), -vd))
        assert np.array_equal(ti[0], order)
        with pytest.raises(ValueError):
            gp.predict(np.full((2, d), np.nan))
    with pytest.raises(Exception):
        X = rng.uniform(-1, 1, (80, 65))
        mb.GaussianProcess(mean=mb.constant_trend(65, beta=None), theta0=[0.1] * 65, thetaL=[1e-3] * 65,
                           thetaU=[10.0] * 65).fit_fixed(X, X.sum(1), np.full(65, 0.1))
