"""K2c on the INT8 tensor pipe (option "trmm_i8": Ozaki split of L^-1 and r, tcgen05.mma kind::i8 with
TMEM accumulators, csrc/ozaki.cu) against the oracle at the same tolerances as the FP64 DMMA path:
relative 1e-9 on the mean (which this path does not touch), 1e-7 on MSE and acquisition values with 8
slices per operand; the 7-slice variant (49 bits) is gated at 1e-6.  Also: bitwise repeatability and
chunk invariance (the slice products are exact integer sums), ragged sizes, both trends, and the
switch back to the DMMA path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mods():
    import mipego_b200 as mb
    from oracle import gp_oracle as O
    return mb, O


def _model(mb, O, n, d, kind, ordinary, seed):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-5, 5, (n, d))
    y = (X ** 2).sum(1) + 3 * np.sin(X[:, 0])
    y = (y - y.mean()) / y.std()
    theta = (0.12 / d) * rng.uniform(0.5, 1.5, d) * (4.0 if kind == "matern" else 1.0)
    par = np.r_[theta, 0.9]
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None if ordinary else 0.2), corr=kind, theta0=theta,
                            thetaL=1e-5 * np.ones(d), thetaU=100 * np.ones(d), nugget=1e-6)
    gp.fit_fixed(X, y, par)
    st = O.fit_state(X, y, par, kind, ordinary, 0.2, "noisy", 1e-6)
    return gp, st, X, y, rng


@pytest.mark.parametrize("slices", [8, 7])
@pytest.mark.parametrize("n,d,kind,ordinary", [(500, 10, "squared_exponential", True), (1100, 7, "matern", True),
                                               (333, 3, "matern", False), (1999, 12, "squared_exponential", False),
                                               (64, 2, "squared_exponential", True), (129, 5, "absolute_exponential", True),
                                               # d = 40: the widest K1 that emits the slices itself; d = 44: separate slicing pass
                                               (300, 40, "matern", False), (300, 44, "squared_exponential", True)])
def test_i8_posterior_against_oracle(slices, n, d, kind, ordinary):
    mb, O = _mods()
    gp, st, X, y, rng = _model(mb, O, n, d, kind, ordinary, 1000 + n)
    h = gp._handle()
    m = 3000
    Xs = rng.uniform(-5, 5, (m, d))
    Xs[:40] = X[:40] + 1e-4 * rng.standard_normal((40, d))       # next to training points: 1 - |T|^2 cancels
    Xs[40:44] = X[:4]                                            # on training points
    rm, rs = O.predict(st, Xs)
    h.set_option("trmm_i8", 0)                                   # (a process started with MGP_TRMM_I8=1 has it on)
    mean_f, mse_f = gp.predict(Xs, eval_MSE=True)                # FP64 DMMA path
    h.set_option("trmm_i8", slices)
    assert h.counter("trmm_i8") == slices
    mean, mse = gp.predict(Xs, eval_MSE=True)
    tol = 1e-7 if slices == 8 else 1e-6
    s2 = st["sigma2"]
    e_mean = np.max(np.abs(mean - rm) / (np.abs(rm) + 1e-3))
    e_mse = np.max(np.abs(mse - rs) / (np.abs(rs) + 1e-6 * s2))
    e_vs_dmma = np.max(np.abs(mse - mse_f) / (np.abs(mse_f) + 1e-6 * s2))
    print("i8 S=%d n=%d %s: mean %.2e mse vs oracle %.2e (DMMA path %.2e), vs DMMA %.2e" %
          (slices, n, kind[:6], e_mean, e_mse, np.max(np.abs(mse_f - rs) / (np.abs(rs) + 1e-6 * s2)), e_vs_dmma))
    assert np.array_equal(mean, mean_f)                          # the mean comes from K1, untouched
    assert e_mean < 1e-9 and e_mse < tol
    # two-CTA clusters sharing the candidate slices by multicast: the same exact integer sums, recombined
    # in the same order -> bitwise the single-CTA kernel
    h.set_option("trmm_i8_pair", 1)
    mean_p, mse_p = gp.predict(Xs, eval_MSE=True)
    h.set_option("trmm_i8_pair", 0)
    assert np.array_equal(mse_p, mse) and np.array_equal(mean_p, mean)
    # K1 emitting the slices itself (default where it can: 8 slices, SE / Matern, constant trend) against the
    # separate slicing pass over the FP64 r: the same fixed-point digits, hence bitwise the same result
    assert h.counter("trmm_i8_fuse") == 1
    h.set_option("trmm_i8_fuse", 0)
    mean_s, mse_s = gp.predict(Xs, eval_MSE=True)
    h.set_option("trmm_i8_fuse", 1)
    assert np.array_equal(mse_s, mse) and np.array_equal(mean_s, mean)
    plugin = float(y.min())
    for crit, name, par in ((mb.EI(model=gp, minimize=True, plugin=plugin), "EI", None),
                            (mb.MGFI(model=gp, minimize=True, plugin=plugin, t=2.0), "MGFI", 2.0),
                            (mb.UCB(model=gp, minimize=True, alpha=0.7), "UCB", 0.7)):
        v = crit(Xs)
        rv = O.acquisition(st, Xs, name, par, plugin, True)
        easy = rs > 1e-4 * s2
        assert np.max(np.abs(v[easy] - rv[easy]) / (np.abs(rv[easy]) + 1e-6 * np.abs(rv).max())) < tol, name
    # exact integer slice products: bitwise repeatable, independent of chunking and candidate order
    mse2 = gp.predict(Xs, eval_MSE=True)[1]
    h.set_option("chunk", 512)
    mse3 = gp.predict(Xs, eval_MSE=True)[1]
    perm = rng.permutation(m)
    mse4 = gp.predict(Xs[perm], eval_MSE=True)[1]
    h.set_option("chunk", 0)
    assert np.array_equal(mse, mse2) and np.array_equal(mse, mse3) and np.array_equal(mse4, mse[perm])
    ei = mb.EI(model=gp, minimize=True, plugin=plugin)
    V = ei.evaluate(Xs)[0]
    tv, ti = ei.topk(Xs, k=8)
    order = np.lexsort((np.arange(m), -V))[:8]
    assert np.array_equal(ti[0], order) and np.array_equal(tv[0], V[order])
    # gradients and single points take the FP64 paths; switching off restores the DMMA numbers bit for bit
    v1, dx1 = ei(Xs[:20], return_dx=True)
    h.set_option("trmm_i8", 0)
    v0, dx0 = ei(Xs[:20], return_dx=True)
    assert np.array_equal(dx1, dx0) and np.array_equal(v1, v0)
    assert np.array_equal(gp.predict(Xs, eval_MSE=True)[1], mse_f)


def test_i8_follows_refit_and_append():
    """The packed int8 slices belong to the fitted model: a refit, new data, or appended rows rebuild them."""
    mb, O = _mods()
    gp, st, X, y, rng = _model(mb, O, 700, 6, "squared_exponential", True, 77)
    h = gp._handle()
    h.set_option("trmm_i8", 1)
    Xs = rng.uniform(-5, 5, (2000, 6))
    a = gp.predict(Xs, eval_MSE=True)[1]
    par = np.concatenate([np.ravel(v) for v in gp.par.values()])
    Xn = np.r_[X, rng.uniform(-5, 5, (5, 6))]
    yn = np.r_[y, rng.standard_normal(5)]
    gp.update(Xn, yn, reoptimize=False)                    # rank-5 append
    assert h.counter("appends") == 1
    st2 = O.fit_state(Xn, yn, par, "squared_exponential", True, 0.0, "noisy", 1e-6)
    rs = O.predict(st2, Xs)[1]
    b = gp.predict(Xs, eval_MSE=True)[1]
    assert np.max(np.abs(b - rs) / (np.abs(rs) + 1e-6 * st2["sigma2"])) < 1e-7
    assert not np.array_equal(a, b)
    gp.fit_fixed(X, y, par * np.r_[np.full(6, 1.3), 1.0])
    st3 = O.fit_state(X, y, par * np.r_[np.full(6, 1.3), 1.0], "squared_exponential", True, 0.0, "noisy", 1e-6)
    rs3 = O.predict(st3, Xs)[1]
    c = gp.predict(Xs, eval_MSE=True)[1]
    assert np.max(np.abs(c - rs3) / (np.abs(rs3) + 1e-6 * st3["sigma2"])) < 1e-7
