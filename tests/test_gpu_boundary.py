"""Boundary behaviour of the drop-in on the GPU: isotropic theta against
reference fixtures, stream ordering between the device-pointer and host-pointer entry points,
the env harvest without side effects, pickling in every estimation mode, the NCCL entry points on
a one-rank communicator, the documented limits (d <= 64, n_sets <= 16, k <= 64), and the ctypes
stub of INTEGRATION.md run against the real library."""
import os
import pickle
import re

import numpy as np
import pytest

from golden_util import ISOS, load, mode_of, beta_of, likelihood_of

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mods():
    import mipego_b200 as mb
    from oracle import gp_oracle as O
    return mb, O


def _data(n, d, seed):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-5, 5, (n, d))
    y = (X ** 2).sum(1) + 3 * np.sin(X[:, 0])
    return X, (y - y.mean()) / y.std(), rng


# ---- isotropic theta (kernel.py:312-317) -----------------------------------------------------------
@pytest.mark.parametrize("name", ISOS)
def test_isotropic_theta_golden(name):
    """ONE theta for all features: likelihood, the reference's (quirky) gradient, env, fitted state,
    posterior, gradients and EI against fixtures produced by the unmodified reference."""
    mb, O = _mods()
    g = load(name)
    d = g["X"].shape[1]
    lik, mode = likelihood_of(g), mode_of(g)
    beta = None if g["estimate_trend"] else float(g["beta_in"])
    nug = float(g["nugget"])
    assert g["theta"].size == 1
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=beta), corr=g["kind"], theta0=g["theta"],
                            thetaL=[1e-5], thetaU=[100.0], nugget=nug if nug > 0 else 0,
                            noise_estim=(mode == "noise_estim"), likelihood=lik)
    gp._check_data(g["X"], g["y"])
    par = g["par"]
    assert par.size == 1 + (1 if (lik == "restricted" or mode != "noiseless") else 0)
    P = np.vstack([par, par * np.r_[1.7, np.ones(par.size - 1)], par * np.r_[0.6, np.ones(par.size - 1)]])
    llf, grad, status = gp.log_likelihood_batch(P, eval_grad=True)
    assert status[0] == 0 and abs(llf[0] - float(g["llf"])) <= 1e-9 * abs(float(g["llf"]))
    gs = np.abs(g["llf_grad"]).max()
    assert grad.shape == (3, par.size)
    assert np.max(np.abs(grad[0] - g["llf_grad"])) <= (1e-7 if nug > 0 else 1e-6) * gs, (grad[0], g["llf_grad"])
    fn = O.log_likelihood_restricted if lik == "restricted" else O.log_likelihood_concentrated
    for b in (1, 2):
        rl, rg = fn(g["X"], g["y"], P[b], g["kind"], g["estimate_trend"], beta_of(g), mode, nug, eval_grad=True)
        if np.isfinite(rl):
            assert abs(llf[b] - rl) <= 1e-9 * abs(rl)
            assert np.max(np.abs(grad[b] - rg.ravel())) <= (1e-7 if nug > 0 else 1e-6) * np.abs(rg).max()
    gp.fit_fixed(g["X"], g["y"], par)
    assert gp.theta_.size == 1 and gp.par["theta"].size == 1
    assert np.allclose(gp.C, g["L"], rtol=0, atol=1e-11)
    t = 1e-9 if nug > 0 else 1e-6
    mean, mse = gp.predict(g["Xs"], eval_MSE=True)
    s2 = float(g["sigma2"])
    assert np.max(np.abs(mean - g["mean"]) / (np.abs(g["mean"]) + 1e-3)) < t
    assert np.max(np.abs(mse - g["mse"]) / (np.abs(g["mse"]) + 1e-6 * s2)) < 100 * t
    mdx, sdx = gp.gradient_batch(g["Xs"])
    assert np.max(np.abs(mdx - g["mean_dx"])) < 10 * t * np.abs(g["mean_dx"]).max()
    assert np.max(np.abs(sdx - g["mse_dx"])) < 100 * t * np.abs(g["mse_dx"]).max()
    ei = mb.EI(model=gp, minimize=True, plugin=float(g["plugin_min"]))
    v = ei(g["Xs"])
    easy = g["mse"] > 1e-4 * s2
    rv = g["EI_min"]
    assert np.max(np.abs(v[easy] - rv[easy]) / (np.abs(rv[easy]) + 1e-6 * np.abs(rv).max())) < 100 * t
    # pickle round trip and a refit at the same hyper-parameters keep the one-entry layout
    g2 = pickle.loads(pickle.dumps(gp))
    m2 = g2.predict(g["Xs"])
    assert np.max(np.abs(m2 - mean)) <= 1e-9 * np.abs(mean).max()
    assert g2.theta_.size == 1
    g2.update(g["X"], g["y"], reoptimize=False)
    assert np.max(np.abs(g2.predict(g["Xs"]) - mean)) <= 1e-9 * np.abs(mean).max()


def test_isotropic_fit_against_reference_fit():
    """fit() with one shared theta: same optimiser, same (reference-quirk) gradient, same end point."""
    mb, O = _mods()
    g = load("ifit_se_ok_noisy")
    d = g["X"].shape[1]
    np.random.seed(int(g["seed"]))
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr=g["kind"], theta0=g["theta0"],
                            thetaL=g["thetaL"], thetaU=g["thetaU"], nugget=float(g["nugget"]),
                            random_start=1, restart_mode="sequential")
    gp.fit(g["X"], g["y"])
    llf = float(np.ravel(gp.log_likelihood_)[0])
    assert abs(llf - float(g["llf"])) <= 1e-8 * abs(float(g["llf"]))
    assert np.allclose(gp.theta_, g["theta_"], rtol=1e-6) and gp.eval_count == int(g["eval_count"])
    mean = gp.predict(g["Xs"])
    assert np.max(np.abs(mean - g["mean"]) / (np.abs(g["mean"]) + 1e-2)) < 1e-7
    with pytest.raises(ValueError):
        mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), theta0=[0.1, 0.1], thetaL=[1e-3] * 2,
                           thetaU=[1.0] * 2).fit(g["X"], g["y"])          # neither 1 nor d entries


# ---- the env harvest has no side effects (gpr.py:876-887) -------------------------------------------
def test_env_harvest_leaves_the_fitted_model_alone():
    mb, O = _mods()
    X, y, rng = _data(300, 5, 8)
    theta = (0.3 / 5) * rng.uniform(0.5, 1.5, 5)
    gp = mb.GaussianProcess(mean=mb.constant_trend(5, beta=None), corr="matern", theta0=theta,
                            thetaL=[1e-5] * 5, thetaU=[100.0] * 5, nugget=1e-6)
    gp.fit_fixed(X, y, np.r_[theta, 0.9])
    Xs = rng.uniform(-5, 5, (1000, 5))
    ei = mb.EI(model=gp, minimize=True)
    mean0, mse0 = gp.predict(Xs, eval_MSE=True)
    v0 = ei(Xs)
    L0, gamma0, s0, b0, par0 = gp.C.copy(), gp.gamma.copy(), gp.sigma2, gp.mean.beta.copy(), dict(gp.par)
    other = np.r_[3.0 * theta, 0.4]
    env = {}
    l_other = gp.log_likelihood_concentrated(other, env=env)
    st = O.fit_state(X, y, other, "matern", True, 0.0, "noisy", 1e-6)
    assert abs(l_other - st["llf"]) <= 1e-9 * abs(st["llf"])
    assert np.allclose(env["C"], st["L"], rtol=0, atol=1e-11) and float(env["sigma2"]) == 0.4
    assert np.max(np.abs(env["rho"] - st["rho"])) <= 1e-9 * np.abs(st["rho"]).max()
    # nothing about the fitted model moved: attributes, device state, results
    assert np.array_equal(gp.C, L0) and np.array_equal(gp.gamma, gamma0) and gp.sigma2 == s0
    assert np.array_equal(gp.mean.beta, b0) and all(np.array_equal(gp.par[k], par0[k]) for k in par0)
    mean1, mse1 = gp.predict(Xs, eval_MSE=True)
    assert np.array_equal(mean0, mean1) and np.array_equal(mse0, mse1) and np.array_equal(ei(Xs), v0)


# ---- stream ordering between *_dev and host-pointer calls -------------------------------------------
def test_device_and_host_entry_points_interleave_without_explicit_sync():
    """mgp_acq_dev / mgp_nll_batch_dev enqueue on the caller's stream and share the handle's scratch
    with the host-pointer calls (internal stream): the library orders them itself (event recorded
    after every call, waited on by the next user), so no cudaDeviceSynchronize is needed between."""
    import torch
    mb, O = _mods()
    X, y, rng = _data(1100, 8, 9)
    theta = (0.2 / 8) * rng.uniform(0.5, 1.5, 8)
    gp = mb.GaussianProcess(mean=mb.constant_trend(8, beta=None), corr="squared_exponential", theta0=theta,
                            thetaL=[1e-5] * 8, thetaU=[100.0] * 8, nugget=1e-6)
    gp.fit_fixed(X, y, np.r_[theta, 0.9])
    ei = mb.EI(model=gp, minimize=True)
    Xa = rng.uniform(-5, 5, (200_000, 8))
    Xb = rng.uniform(-5, 5, (30_000, 8))
    va = ei.evaluate(Xa)[0]
    vb = ei.evaluate(Xb)[0]
    P = np.c_[theta * rng.uniform(0.5, 2.0, (4, 8)), np.full(4, 0.8)]
    l_ref, g_ref, _ = gp.log_likelihood_batch(P)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    Xa_d, P_d = torch.from_numpy(Xa).cuda(), torch.from_numpy(P).cuda()
    torch.cuda.synchronize()
    for rep in range(5):
        with torch.cuda.stream(side):
            out_a = ei.evaluate_device(Xa_d)                   # long, on a user stream, not synchronised
        vb2 = ei.evaluate(Xb)[0]                               # host API right behind it: same scratch
        with torch.cuda.stream(side):
            lo = gp.log_likelihood_batch_device(P_d)
        gp.fit_fixed(X, y, np.r_[theta, 0.9])                  # rewrites the model behind a pending dev call
        with torch.cuda.stream(side):
            out_a2 = ei.evaluate_device(Xa_d)
        mean_b = gp.predict(Xb)                                # and again host behind device
        side.synchronize()
        assert np.array_equal(out_a[0].cpu().numpy(), va) and np.array_equal(vb2, vb)
        assert np.array_equal(out_a2[0].cpu().numpy(), va)
        assert np.array_equal(lo[0].cpu().numpy(), l_ref) and np.array_equal(lo[1].cpu().numpy(), g_ref)
        assert np.all(np.isfinite(mean_b))


# ---- small host-pointer calls through the mapped pinned page ---------------------------------------------
@pytest.mark.parametrize("kind,ordinary", [("squared_exponential", True), ("matern", False)])
def test_small_calls_zero_copy_equals_copy_engine_path(kind, ordinary):
    """Calls with <= 4 KB of candidates (the reference optimisers' one-point-per-call pattern,
    optimizer/utils.py:38-42) read their input from, and write their results to, a pinned page mapped into
    the device address space -- no copy engine, no events, one synchronisation.  Every output kind must be
    bitwise what the three-stream staging path returns, at every m up to the limit and one beyond it."""
    mb, O = _mods()
    d = 6
    X, y, rng = _data(700, d, 21)
    theta = (0.2 / d) * rng.uniform(0.5, 1.5, d)
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None if ordinary else 0.1), corr=kind, theta0=theta,
                            thetaL=[1e-5] * d, thetaU=[100.0] * d, nugget=1e-6)
    gp.fit_fixed(X, y, np.r_[theta, 0.9])
    h = gp._handle()
    ei = mb.EI(model=gp, minimize=True)
    mg = mb.MGFI(model=gp, minimize=True, t=1.0)
    ts = [0.5, 1.0, 2.0]

    def everything(Z):
        out = list(gp.predict(np.atleast_2d(Z), eval_MSE=True)) + list(gp.gradient_batch(np.atleast_2d(Z)))
        out += list(ei(Z, return_dx=True)) + [ei(Z)] + [mg.evaluate(np.atleast_2d(Z), params=ts)]
        return [np.asarray(a) for a in out]

    assert h.counter("zero_copy") == 1
    for m in (1, 2, 8, 9, 33, 85, 86):            # 85 x 6 x 8 B = 4080 B: the last size on the page
        Z = rng.uniform(-5, 5, (m, d)) if m > 1 else rng.uniform(-5, 5, d)
        a = everything(Z)
        h.set_option("zero_copy", 0)
        b = everything(Z)
        h.set_option("zero_copy", 1)
        for u, v in zip(a, b):
            assert u.shape == v.shape and np.array_equal(u, v), m
    # against the oracle once, through the page
    st = O.fit_state(X, y, np.r_[theta, 0.9], kind, ordinary, 0.1, "noisy", 1e-6)
    Z = rng.uniform(-5, 5, (5, d))
    rm, rs = O.predict(st, Z)
    mean, mse = gp.predict(Z, eval_MSE=True)
    assert np.max(np.abs(mean - rm) / (np.abs(rm) + 1e-3)) < 1e-9
    assert np.max(np.abs(mse - rs) / (np.abs(rs) + 1e-6 * st["sigma2"])) < 1e-7


# ---- triangular inversion beside the factorisation -----------------------------------------------------------
def test_rowwise_inversion_matches_the_recursive_one():
    """Option trtri_rowwise: L^-1 block row by block row on a second stream while the Cholesky runs.  Same
    likelihood, gradient and status as the recursive inversion after the factorisation (different summation
    order: gated at 1e-10 / 1e-9), eager and replayed from a CUDA graph, one and two batch groups."""
    mb, O = _mods()
    d = 5
    for n, B in ((130, 3), (600, 4), (1100, 5)):
        X, y, rng = _data(n, d, 40 + n)
        theta = (0.5 / d) * rng.uniform(0.5, 1.5, d)
        gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="matern", theta0=theta,
                                thetaL=[1e-5] * d, thetaU=[100.0] * d, nugget=1e-6)
        gp._check_data(X, y)
        h = gp._handle()
        P = np.c_[theta * rng.uniform(0.6, 1.6, (B, d)), rng.uniform(0.5, 1.0, B)]
        P[-1, :d] = 1e-9        # nearly singular correlation matrix: the failure status must agree as well
        out = {}
        for opt in (0, 1):
            h.set_option("trtri_rowwise", opt)
            res = [gp.log_likelihood_batch(P) for _ in range(3)]      # eager, then two graph replays
            for r in res[1:]:
                assert np.array_equal(r[0], res[0][0]) and np.array_equal(r[1], res[0][1])
            out[opt] = res[0]
        assert np.array_equal(out[0][2], out[1][2])
        ok = out[0][2] == 0
        assert ok[:-1].all()
        assert np.max(np.abs(out[0][0][ok] - out[1][0][ok]) / np.abs(out[0][0][ok])) < 1e-10
        g0, g1 = out[0][1][ok], out[1][1][ok]
        assert np.max(np.abs(g0 - g1)) / np.max(np.abs(g0)) < 1e-9
        h.set_option("trtri_rowwise", -1)


# ---- CUDA graphs of the batched likelihood ---------------------------------------------------------------
def test_likelihood_graphs_are_bitwise_the_eager_path():
    """A batch shape is captured in a CUDA graph at its second use and replayed afterwards (host
    and device entry points, one and several batch groups): every value is bitwise what the
    launch-by-launch path gives; new training data, a different batch size or appended rows
    never reuse a stale graph."""
    import torch
    mb, O = _mods()
    d = 6
    X, y, rng = _data(1400, d, 21)
    theta = (0.4 / d) * rng.uniform(0.5, 1.5, d)
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="squared_exponential", theta0=theta,
                            thetaL=[1e-5] * d, thetaU=[100.0] * d, nugget=1e-6)
    gp._check_data(X[:1300], y[:1300])
    h = gp._handle()

    def draw(B):
        return np.c_[theta * rng.uniform(0.6, 1.6, (B, d)), rng.uniform(0.5, 1.0, B)]
    for B in (6, 3):
        h.set_option("nll_graphs", 0)
        Ps = [draw(B) for _ in range(4)]
        eager = [gp.log_likelihood_batch(P) for P in Ps]
        h.set_option("nll_graphs", 1)
        r0 = h.counter("graph_replays")
        for rep in range(2):
            for P, e in zip(Ps, eager):
                llf, grad, status = gp.log_likelihood_batch(P)
                assert np.array_equal(llf, e[0]) and np.array_equal(grad, e[1]) and np.array_equal(status, e[2])
        assert h.counter("graph_replays") - r0 == 7           # first call eager, the rest replayed
        l_ng = gp.log_likelihood_batch(Ps[0], eval_grad=False)[0]     # another shape (no gradient)
        assert np.array_equal(l_ng, eager[0][0])
    # device entry point on a user stream: the caller's tensors change, the graph does not
    side = torch.cuda.Stream()
    Pd = [torch.from_numpy(P).cuda() for P in Ps]
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        outs = [gp.log_likelihood_batch_device(P) for P in Pd for _ in range(2)]
    side.synchronize()
    for k, o in enumerate(outs):
        e = eager[k // 2]
        assert np.array_equal(o[0].cpu().numpy(), e[0]) and np.array_equal(o[1].cpu().numpy(), e[1])
    # a failing element inside a replayed batch still reports its status (noiseless model: theta -> 0
    # makes R singular, gpr.py:820-826)
    nl = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="squared_exponential", theta0=theta,
                            thetaL=[1e-12] * d, thetaU=[100.0] * d)
    nl._check_data(X[:300], y[:300])
    bad = np.vstack([20 * theta, np.full(d, 1e-12), 30 * theta])
    outs = [nl.log_likelihood_batch(bad) for _ in range(3)]
    assert nl._handle().counter("graph_replays") == 2
    for llf, grad, status in outs:
        assert status[1] > 0 and np.isinf(llf[1]) and np.all(grad[1] == 0) and status[0] == 0 and status[2] == 0
        assert np.array_equal(llf, outs[0][0]) and np.array_equal(grad, outs[0][1])
    # new data: same shapes, fresh graphs
    gp._check_data(X, y)
    ref = O.log_likelihood_concentrated(X, y, Ps[0][0], "squared_exponential", True, 0.0, "noisy", 1e-6)
    for _ in range(3):
        llf = gp.log_likelihood_batch(Ps[0])[0]
        assert abs(llf[0] - ref) <= 1e-9 * abs(ref)
    # many groups (n = 4096 uses 8 streams): eager == replay
    Xb, yb, rb = _data(2700, d, 22)
    big = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="matern", theta0=theta, thetaL=[1e-5] * d,
                             thetaU=[100.0] * d, nugget=1e-6)
    big._check_data(Xb, yb)
    P = np.c_[4 * theta * rb.uniform(0.6, 1.6, (8, d)), rb.uniform(0.5, 1.0, 8)]
    first = big.log_likelihood_batch(P)
    for _ in range(3):
        again = big.log_likelihood_batch(P)
        assert np.array_equal(first[0], again[0]) and np.array_equal(first[1], again[1])
    assert big._handle().counter("graph_replays") == 3          # four calls: eager, capture + replay, replay, replay


# ---- pickling keeps the full hyper-parameter vector ---------------------------------------------------
@pytest.mark.parametrize("lik,noise_estim,nugget", [("concentrated", False, 1e-4), ("restricted", False, 0),
                                                    ("restricted", True, 0), ("concentrated", True, 0)])
def test_unpickle_then_update_without_reoptimisation(lik, noise_estim, nugget):
    mb, O = _mods()
    X, y, rng = _data(120, 3, 10)
    y = y + 0.05 * rng.standard_normal(y.size)
    np.random.seed(5)
    gp = mb.GaussianProcess(mean=mb.constant_trend(3, beta=None), corr="squared_exponential", theta0=[0.05] * 3,
                            thetaL=[1e-4] * 3, thetaU=[10.0] * 3, nugget=nugget, noise_estim=noise_estim,
                            likelihood=lik, random_start=2)
    gp.fit(X[:100], y[:100])
    par = np.concatenate([np.ravel(v) for v in gp.par.values()])
    g2 = pickle.loads(pickle.dumps(gp))
    assert type(g2.sigma2) is type(gp.sigma2) and np.array_equal(np.ravel(g2.sigma2), np.ravel(gp.sigma2))
    assert all(np.array_equal(g2.par[k], gp.par[k]) for k in gp.par)
    Xs = rng.uniform(-5, 5, (64, 3))
    assert np.max(np.abs(g2.predict(Xs) - gp.predict(Xs))) <= 1e-9
    g2.update(X, y, reoptimize=False)                     # 20 more rows at the pickled hyper-parameters
    par2 = np.concatenate([np.ravel(v) for v in g2.par.values()])
    assert np.array_equal(par2, par)
    fresh = mb.GaussianProcess(mean=mb.constant_trend(3, beta=None), corr="squared_exponential", theta0=[0.05] * 3,
                               thetaL=[1e-4] * 3, thetaU=[10.0] * 3, nugget=nugget, noise_estim=noise_estim,
                               likelihood=lik).fit_fixed(X, y, par)
    m2, s2 = g2.predict(Xs, eval_MSE=True)
    m3, s3 = fresh.predict(Xs, eval_MSE=True)
    assert np.max(np.abs(m2 - m3)) <= 1e-9 * np.abs(m3).max() and np.max(np.abs(s2 - s3)) <= 1e-7 * np.abs(s3).max() + 1e-12


# ---- incremental training: rank-k append (gpr.py:387-390) ------------------------------------------
@pytest.mark.parametrize("kind,mode,n0,k", [("squared_exponential", "noisy", 700, 3),      # same padded size
                                            ("matern", "noisy", 1150, 8),                 # crosses a 128-row tile
                                            ("squared_exponential", "noiseless", 1536, 1),  # n0 a multiple of 128
                                            ("matern", "noiseless_sk", 2000, 130),         # two new tile rows
                                            ("absolute_exponential", "noisy", 900, 5)])
def test_append_matches_full_rebuild_and_oracle(kind, mode, n0, k):
    """update(X, y, reoptimize=False) with appended rows extends L, L^-1 and the packed operands in
    O(n^2 k) (mgp_append_train).  Against (i) the full rebuild at the same hyper-parameters
    (<= 1e-10 on the factor, 1e-9 / 1e-7 on the posterior) and (ii) the oracle's model on the
    extended data; the likelihood, sigma2 and beta follow the new y."""
    mb, O = _mods()
    d = 6
    X, y, rng = _data(n0 + k, d, 100 + n0)
    scale = {"squared_exponential": 1.0, "matern": 4.0, "absolute_exponential": 1.5}[kind]
    theta = (0.5 / d) * rng.uniform(0.5, 1.5, d) * scale
    sk = mode.endswith("_sk")
    nugget = 1e-6 if mode == "noisy" else 0
    par = np.r_[theta, 0.85] if mode == "noisy" else theta

    def new_gp():      # a fresh trend object each time: fitting stores beta in it (as in the reference)
        return mb.GaussianProcess(mean=mb.constant_trend(d, beta=0.1 if sk else None), corr=kind, theta0=theta,
                                  thetaL=[1e-5] * d, thetaU=[100.0] * d, nugget=nugget)
    gp = new_gp().fit_fixed(X[:n0], y[:n0], par)
    h = gp._handle()
    gp.update(X, y, reoptimize=False)                              # appended rows -> mgp_append_train
    assert gp.X.shape[0] == n0 + k and h.counter("n") == n0 + k
    assert h.counter("appends") == 1                               # the incremental path, not the fallback rebuild
    full = new_gp().fit_fixed(X, y, par)
    Lf, La = full.C, gp.C
    assert np.max(np.abs(La - Lf)) <= 1e-10 * np.abs(Lf).max()
    assert np.max(np.abs(gp.gamma - full.gamma)) <= 1e-8 * np.abs(full.gamma).max()
    assert abs(float(np.ravel(gp.sigma2)[0]) - float(np.ravel(full.sigma2)[0])) <= 1e-10 * float(np.ravel(full.sigma2)[0])
    assert abs(float(np.ravel(gp.log_likelihood_)[0]) - float(np.ravel(full.log_likelihood_)[0])) <= \
        1e-10 * abs(float(np.ravel(full.log_likelihood_)[0]))
    if not sk:
        assert np.max(np.abs(gp.mean.beta - full.mean.beta)) <= 1e-9 * max(1.0, np.abs(full.mean.beta).max())
    Xs = rng.uniform(-5, 5, (2000, d))
    ma, sa = gp.predict(Xs, eval_MSE=True)
    mf, sf = full.predict(Xs, eval_MSE=True)
    s2 = float(np.ravel(full.sigma2)[0])
    assert np.max(np.abs(ma - mf) / (np.abs(mf) + 1e-3)) < 1e-9
    assert np.max(np.abs(sa - sf) / (np.abs(sf) + 1e-6 * s2)) < 1e-7
    st = O.fit_state(X, y, par, kind, not sk, 0.1, "noisy" if mode == "noisy" else "noiseless", float(nugget))
    rm, rs = O.predict(st, Xs)
    assert np.max(np.abs(ma - rm) / (np.abs(rm) + 1e-3)) < 1e-9
    assert np.max(np.abs(sa - rs) / (np.abs(rs) + 1e-6 * st["sigma2"])) < 1e-7
    ei = mb.EI(model=gp, minimize=True)
    v, dx = ei(Xs[:200], return_dx=True)                           # gradient path: R^-1 is rebuilt lazily
    rv, rdx = O.acquisition(st, Xs[:200], "EI", None, None, True, return_dx=True)
    assert np.max(np.abs(v - rv) / (np.abs(rv) + 1e-6 * np.abs(rv).max())) < 1e-7
    assert np.max(np.abs(dx - rdx)) < 1e-7 * np.abs(rdx).max()
    # a second append on top of the first, then the likelihood entry point still sees the new data
    X2 = np.r_[X, rng.uniform(-5, 5, (2, d))]
    y2 = np.r_[y, 0.3, -0.2]
    gp.update(X2, y2, reoptimize=False)
    assert h.counter("appends") == 2
    full2 = new_gp().fit_fixed(X2, y2, par)
    assert np.max(np.abs(gp.C - full2.C)) <= 1e-10 * np.abs(full2.C).max()
    l1, g1, _ = gp.log_likelihood_batch(par.reshape(1, -1))
    l2, g2, _ = full2.log_likelihood_batch(par.reshape(1, -1))
    # (not bitwise: the appended model keeps the centring offsets of the original training set)
    assert np.allclose(l1, l2, rtol=1e-12, atol=0) and np.allclose(g1, g2, rtol=1e-10, atol=1e-10 * np.abs(g2).max())


def test_append_rejections_keep_the_model():
    mb, O = _mods()
    d = 3
    X, y, rng = _data(900, d, 77)
    theta = np.array([0.3, 0.2, 0.4])
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="squared_exponential", theta0=theta,
                            thetaL=[1e-5] * d, thetaU=[100.0] * d, nugget=1e-6).fit_fixed(X, y, np.r_[theta, 0.9])
    Xs = rng.uniform(-5, 5, (300, d))
    before = gp.predict(Xs)
    with pytest.raises(Exception, match="Multiple input features"):
        gp.update(np.r_[X, X[17:18]], np.r_[y, 0.0], reoptimize=False)      # a duplicate of row 17
    assert gp.X.shape[0] == 900 and np.array_equal(gp.predict(Xs), before)
    # too many new rows for the incremental path: silently rebuilt, same answer as a fresh model
    Xn = np.r_[X, rng.uniform(-5, 5, (400, d))]
    yn = np.r_[y, rng.standard_normal(400)]
    gp.update(Xn, yn, reoptimize=False)
    assert gp._handle().counter("appends") == 0
    fresh = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="squared_exponential", theta0=theta,
                               thetaL=[1e-5] * d, thetaU=[100.0] * d, nugget=1e-6).fit_fixed(Xn, yn, np.r_[theta, 0.9])
    assert np.array_equal(gp.predict(Xs), fresh.predict(Xs))


# ---- NCCL entry points on a one-rank communicator ---------------------------------------------------
def test_nccl_entry_points_single_rank():
    """mgp_comm_* / mgp_bcast_model / mgp_allgather_topk with world = 1: NCCL is found and bound at
    run time, the pack / all-gather / merge path reproduces the single-GPU top-k order."""
    import ctypes
    import torch
    mb, O = _mods()
    from mipego_b200 import _cabi
    X, y, rng = _data(200, 4, 11)
    theta = (0.3 / 4) * rng.uniform(0.5, 1.5, 4)
    gp = mb.GaussianProcess(mean=mb.constant_trend(4, beta=None), corr="squared_exponential", theta0=theta,
                            thetaL=[1e-5] * 4, thetaU=[100.0] * 4, nugget=1e-6)
    gp.fit_fixed(X, y, np.r_[theta, 0.9])
    h = gp._handle()
    buf = ctypes.create_string_buffer(128)
    assert h.lib.mgp_comm_unique_id(buf) == 0, h.lib.mgp_comm_info(h.h)
    h.check(h.lib.mgp_comm_init(h.h, 0, 1, buf.raw))
    info = h.lib.mgp_comm_info(h.h).decode()
    assert "world=1" in info and "version=" in info
    print("NCCL:", info)
    Xs = rng.uniform(-5, 5, (5000, 4))
    before = gp.predict(Xs)
    h.check(h.lib.mgp_bcast_model(h.h, 0))                 # root == only rank: the model is unchanged
    assert np.array_equal(gp.predict(Xs), before) and h.counter("bcast_bytes") > 200 * 200 * 8
    ts = np.array([1.0, 2.5, 0.3])
    crit = mb.MGFI(model=gp, minimize=True, t=1.0)
    tv, ti = crit.topk(Xs, k=7, params=ts)
    ti_holes = ti.copy()
    ti_holes[1, 5:] = -1                                   # empty slots of a short shard
    ov, oi = np.empty_like(tv), np.empty_like(ti)
    h.check(h.lib.mgp_allgather_topk(h.h, 3, 7, _cabi.dptr(tv), ti_holes.ctypes.data_as(_cabi.c_lp), 1000,
                                     _cabi.dptr(ov), oi.ctypes.data_as(_cabi.c_lp)))
    assert np.array_equal(oi[0], ti[0] + 1000) and np.array_equal(ov[0], tv[0])
    assert np.array_equal(oi[1, :5], ti[1, :5] + 1000) and np.all(oi[1, 5:] == -1) and np.all(np.isinf(ov[1, 5:]))
    # device-resident variant on torch's stream
    tvd, tid = crit.topk_device(torch.from_numpy(Xs).cuda(), k=7, params=ts)
    ovd, oid = torch.empty_like(tvd), torch.empty_like(tid)
    st = torch.cuda.current_stream().cuda_stream
    h.check(h.lib.mgp_allgather_topk_dev(h.h, 3, 7, tvd.data_ptr(), tid.data_ptr(), 0, ovd.data_ptr(),
                                         oid.data_ptr(), st))
    torch.cuda.synchronize()
    assert np.array_equal(oid.cpu().numpy(), ti) and np.array_equal(ovd.cpu().numpy(), tv)
    h.check(h.lib.mgp_comm_destroy(h.h))
    assert h.lib.mgp_bcast_model(h.h, 0) == _cabi.MGP_ESTATE


# ---- documented limits ----------------------------------------------------------------------------------
def test_limits_at_the_edge():
    """d <= 64 (MGP_MAX_D), n_sets <= 16, k <= 64: the largest admitted value works and agrees with
    the oracle, one more is rejected with MGP_EINVAL (no silent truncation)."""
    mb, O = _mods()
    from mipego_b200 import _cabi
    n, d = 160, 64
    X, y, rng = _data(n, d, 12)
    theta = (0.12 / d) * rng.uniform(0.5, 1.5, d)
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="squared_exponential", theta0=theta,
                            thetaL=[1e-6] * d, thetaU=[100.0] * d, nugget=1e-6)
    par = np.r_[theta, 0.9]
    gp.fit_fixed(X, y, par)
    st = O.fit_state(X, y, par, "squared_exponential", True, 0.0, "noisy", 1e-6)
    Xs = rng.uniform(-5, 5, (300, d))
    mean, mse = gp.predict(Xs, eval_MSE=True)
    rm, rs = O.predict(st, Xs)
    assert np.max(np.abs(mean - rm) / (np.abs(rm) + 1e-3)) < 1e-9
    assert np.max(np.abs(mse - rs) / (np.abs(rs) + 1e-6 * st["sigma2"])) < 1e-7
    llf, grad, status = gp.log_likelihood_batch(par.reshape(1, -1))
    rl, rg = O.log_likelihood_concentrated(X, y, par, "squared_exponential", True, 0.0, "noisy", 1e-6, eval_grad=True)
    assert abs(llf[0] - rl) <= 1e-9 * abs(rl) and np.max(np.abs(grad[0] - rg.ravel())) <= 1e-7 * np.abs(rg).max()
    ei = mb.EI(model=gp, minimize=True)
    v, dx = ei(Xs[:40], return_dx=True)                    # K5 with DP = 64
    rv, rdx = O.acquisition(st, Xs[:40], "EI", None, None, True, return_dx=True)
    assert np.max(np.abs(v - rv) / (np.abs(rv) + 1e-6 * np.abs(rv).max())) < 1e-7
    assert np.max(np.abs(dx - rdx)) < 1e-7 * np.abs(rdx).max()
    ts = np.linspace(0.2, 3.0, 16)                         # n_sets = 16
    crit = mb.MGFI(model=gp, minimize=True, t=1.0)
    V = crit.evaluate(Xs, params=ts)
    for s in (0, 15):
        rv = O.acquisition(st, Xs, "MGFI", float(ts[s]), None, True)
        assert np.max(np.abs(V[s] - rv) / (np.abs(rv) + 1e-6 * np.abs(rv).max())) < 1e-7
    tv, ti = crit.topk(Xs, k=64, params=ts)                # k = 64
    for s in (0, 7, 15):
        order = np.lexsort((np.arange(len(Xs)), -V[s]))[:64]
        assert np.array_equal(ti[s], order) and np.array_equal(tv[s], V[s][order])
    with pytest.raises(_cabi.MgpError) as e:
        crit.evaluate(Xs, params=np.linspace(0.2, 3.0, 17))
    assert e.value.code == _cabi.MGP_EINVAL
    with pytest.raises(_cabi.MgpError) as e:
        crit.topk(Xs, k=65)
    assert e.value.code == _cabi.MGP_EINVAL
    big = mb.GaussianProcess(mean=mb.constant_trend(65, beta=None), corr="squared_exponential", theta0=[0.01] * 65,
                             thetaL=[1e-6] * 65, thetaU=[100.0] * 65)
    with pytest.raises(_cabi.MgpError) as e:
        big.fit_fixed(rng.uniform(-5, 5, (100, 65)), rng.uniform(size=100), np.full(65, 0.01))
    assert e.value.code == _cabi.MGP_EINVAL and "unsupported" in str(e.value)


# ---- the from-scratch ctypes stub, run against the real library ------------------------
def _doc_stub():
    txt = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"<!-- ctypes-stub:begin -->\s*```python\n(.*?)```\s*<!-- ctypes-stub:end -->", txt, flags=re.S)
    assert m, "INTEGRATION.md lost its ctypes stub markers"
    return m.group(1)


def test_integration_doc_stub_runs():
    mb, O = _mods()
    ns = {}
    exec(compile(_doc_stub(), "INTEGRATION.md", "exec"), ns)
    X, y, rng = _data(90, 3, 13)
    Xs = rng.uniform(-5, 5, (500, 3))
    theta = np.array([0.2, 0.3, 0.1])        # a finite likelihood (smoother models have llf > 0, which the reference rejects)
    lib = ns["bind"](os.path.join(ROOT, "mip-ego_b200", "libmgp_b200.so"))
    llf, ei = ns["demo"](lib, X, y, Xs, theta, sigma2=0.9, nugget=1e-6)
    st = O.fit_state(X, y, np.r_[theta, 0.9], "squared_exponential", True, 0.0, "noisy", 1e-6)
    rv = O.acquisition(st, Xs, "EI", None, float(y.min()), True)
    assert np.isfinite(st["llf"]) and abs(llf - st["llf"]) <= 1e-9 * abs(st["llf"])
    assert np.max(np.abs(ei - rv) / (np.abs(rv) + 1e-6 * np.abs(rv).max())) < 1e-7
