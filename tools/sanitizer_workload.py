"""Small pass over every kernel family of the library, for compute-sanitizer (memcheck / racecheck /
initcheck): likelihood batch (all phases, two batch groups), fitted state, posterior on the MMA path
and on the <= 8-candidate path, acquisition with gradient, top-k, rank-k append.
    compute-sanitizer --tool memcheck  python tools/sanitizer_workload.py
    compute-sanitizer --tool racecheck python tools/sanitizer_workload.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mipego_b200 as mb

rng = np.random.default_rng(0)
n, d = 700, 5
X = rng.uniform(-5, 5, (n, d))
y = (X ** 2).sum(1) + 3 * np.sin(X[:, 0])
y = (y - y.mean()) / y.std()
theta = (0.4 / d) * rng.uniform(0.5, 1.5, d)
for kind in ("squared_exponential", "matern", "absolute_exponential"):
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr=kind, theta0=theta, thetaL=[1e-5] * d,
                            thetaU=[100.0] * d, nugget=1e-6)
    gp._check_data(X[:690], y[:690])
    gp._handle().set_option("nll_groups", 2)
    P = np.c_[theta * rng.uniform(0.7, 1.4, (4, d)), np.full(4, 0.8)]
    llf, grad, status = gp.log_likelihood_batch(P)
    assert np.all(status == 0), status
    gp.fit_fixed(X[:690], y[:690], P[0])
    gp.update(X, y, reoptimize=False)                      # rank-10 append
    Xs = rng.uniform(-5, 5, (600, d))
    mean, mse = gp.predict(Xs, eval_MSE=True)
    ei = mb.EI(model=gp, minimize=True)
    v, dx = ei(Xs[:300], return_dx=True)
    v1, dx1 = ei(Xs[0], return_dx=True)                    # <= 8 candidates: matrix-vector path
    tv, ti = mb.MGFI(model=gp, minimize=True, t=2.0).topk(Xs, k=5, params=[1.0, 2.0, 3.0])
    print(kind, "llf", llf[0], "appends", gp._handle().counter("appends"), "EI max", float(v.max()), flush=True)
lin = mb.GaussianProcess(mean=mb.linear_trend(d, beta=None), corr="squared_exponential", theta0=theta,
                         thetaL=[1e-5] * d, thetaU=[100.0] * d, nugget=1e-6, likelihood="restricted")
lin.fit_fixed(X[:300], y[:300], np.r_[theta, 0.8])
print("linear trend REML", lin.predict(X[300:310]), flush=True)
print("SANITIZER_WORKLOAD_OK")
