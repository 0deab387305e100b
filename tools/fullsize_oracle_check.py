"""One-off full-size comparison with the CPU oracle (minutes of CPU time, hence not in tests/):
the n = 4096, d = 20 likelihood + gradient (BASELINE configs[3]) and a 512-candidate sample of the
n = 4096 posterior / MGFI values and EI gradients (configs[2])."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mipego_b200 as mb
from oracle import gp_oracle as O

n, d = 4096, 20
rng = np.random.default_rng(30)
X = rng.uniform(-5, 5, (n, d))
y = (X ** 2).sum(1) + 3 * np.sin(X[:, 0])
y = (y - y.mean()) / y.std()
par = np.r_[(0.12 / d) * rng.uniform(0.5, 1.5, d), 0.8]
gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="squared_exponential", theta0=par[:d],
                        thetaL=1e-5 * np.ones(d), thetaU=100 * np.ones(d), nugget=1e-6)
gp._check_data(X, y)
llf, grad, status = gp.log_likelihood_batch(par.reshape(1, -1), eval_grad=True)
t0 = time.time()
rl, rg = O.log_likelihood_concentrated(X, y, par, "squared_exponential", True, 0.0, "noisy", 1e-6, eval_grad=True)
t_or = time.time() - t0
out = {"n": n, "d": d, "llf_gpu": float(llf[0]), "llf_oracle": float(rl), "llf_rel_err": abs(llf[0] - rl) / abs(rl),
       "grad_max_rel_err": float(np.max(np.abs(grad[0] - rg.ravel())) / np.abs(rg).max()), "oracle_seconds": t_or}
gp.fit_fixed(X, y, par)
st = O.fit_state(X, y, par, "squared_exponential", True, 0.0, "noisy", 1e-6)
Xs = rng.uniform(-5, 5, (512, d))
mean, mse = gp.predict(Xs, eval_MSE=True)
rm, rs = O.predict(st, Xs)
out["mean_max_rel_err"] = float(np.max(np.abs(mean - rm) / (np.abs(rm) + 1e-3)))
out["mse_max_rel_err"] = float(np.max(np.abs(mse - rs) / (np.abs(rs) + 1e-6 * st["sigma2"])))
c = mb.MGFI(model=gp, minimize=True, plugin=float(y.min()), t=2.0)
v = c(Xs)
rv = O.acquisition(st, Xs, "MGFI", 2.0, float(y.min()), True)
out["mgfi_max_rel_err"] = float(np.max(np.abs(v - rv) / (np.abs(rv) + 1e-6 * np.abs(rv).max())))
ei = mb.EI(model=gp, minimize=True, plugin=float(y.min()))
v, dx = ei(Xs[:64], return_dx=True)
rv, rdx = O.acquisition(st, Xs[:64], "EI", None, float(y.min()), True, return_dx=True)
out["ei_dx_max_rel_err"] = float(np.max(np.abs(dx - rdx)) / (np.abs(rdx).max() + 1e-300))
print(json.dumps(out))
