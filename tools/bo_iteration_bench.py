"""One Bayesian-optimisation iteration of the reference's loop shape (base.py:361-463) on the CUDA
path: model.fit (multi-restart MLE) + arg-max of the acquisition (5 d L-BFGS-B restarts,
optimizer/utils.py:8-87; ParallelBO: n_point MGFI criteria) at n training points in d dimensions."""
import functools
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mipego_b200 as mb

for n, d, R in ((500, 10, 10), (2000, 20, 8)):
    rng = np.random.default_rng(0)
    X = rng.uniform(-5, 5, (n, d))
    y = (X ** 2).sum(1) + 3 * np.sin(X[:, 0])
    y = (y - y.mean()) / y.std()
    sp = mb.BoxSpace([[-5, 5]] * d)
    np.random.seed(1)
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="matern", theta0=[0.05] * d,
                            thetaL=[1e-4] * d, thetaU=[10.0] * d, nugget=1e-6, random_start=R)
    gp.fit(X[:64], y[:64])          # warm the library (allocations, first launches)
    t0 = time.perf_counter()
    gp.fit(X, y)
    t_fit = time.perf_counter() - t0
    ei = mb.EI(model=gp, minimize=True, plugin=float(y.min()))
    mb.argmax_restart(functools.partial(ei, return_dx=True), sp, eval_budget=100 * d, n_restart=2, optimizer="BFGS")
    t0 = time.perf_counter()
    xo, fo = mb.argmax_restart(functools.partial(ei, return_dx=True), sp, eval_budget=100 * d,
                               n_restart=5 * d, optimizer="BFGS")
    t_acq = time.perf_counter() - t0
    ts = np.exp(np.log(2.0) + 0.5 * np.random.default_rng(22).standard_normal(8))
    crits = [functools.partial(mb.MGFI(model=gp, minimize=True, plugin=float(y.min()), t=float(t)), return_dx=True)
             for t in ts]
    t0 = time.perf_counter()
    xs, fs = mb.batch_argmax_restart(crits, sp, eval_budget=100 * d, n_restart=5 * d, optimizer="BFGS")
    t_par = time.perf_counter() - t0
    print("n=%d d=%d: fit (%d restarts, %d likelihood evaluations) %.3f s | EI arg-max (%d restarts) %.3f s | "
          "ParallelBO 8 x MGFI arg-max (%d trajectories) %.3f s" % (n, d, R, gp.eval_count, t_fit, 5 * d, t_acq, 8 * 5 * d, t_par))
