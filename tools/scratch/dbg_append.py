import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mipego_b200 as mb
from oracle import gp_oracle as O
from scipy import linalg
rng = np.random.default_rng(800); n0, k, d = 700, 3, 6
X = rng.uniform(-5, 5, (n0 + k, d)); y = (X ** 2).sum(1) + 3 * np.sin(X[:, 0]); y = (y - y.mean()) / y.std()
theta = (0.5 / d) * rng.uniform(0.5, 1.5, d); par = np.r_[theta, 0.85]
def mk(): return mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="squared_exponential", theta0=theta, thetaL=[1e-5]*d, thetaU=[100.0]*d, nugget=1e-6)
gp = mk().fit_fixed(X[:n0], y[:n0], par)
gp.update(X, y, reoptimize=False)
full = mk().fit_fixed(X, y, par)
st = O.fit_state(X, y, par, "squared_exponential", True, 0.0, "noisy", 1e-6)
Xs = rng.uniform(-5, 5, (2000, d))
rm, rs = O.predict(st, Xs)
for name, g in (("append", gp), ("full", full)):
    m, s = g.predict(Xs, eval_MSE=True)
    print(name, "vs oracle: mean", np.max(np.abs(m-rm)/(np.abs(rm)+1e-3)), "mse", np.max(np.abs(s-rs)/(np.abs(rs)+1e-6*st["sigma2"])))
    L = g.C; print("  L err", np.abs(L-st["L"]).max(), "gamma err", np.abs(g.gamma-st["gamma"]).max()/np.abs(st["gamma"]).max())
    # MSE at the training points themselves (new rows last)
    mt, s_t = g.predict(X[-6:], eval_MSE=True); print("  mse at last 6 training rows", s_t)
    # small path (<= 8 candidates) uses row-major dMinv
    ms8, ss8 = g.predict(Xs[:5], eval_MSE=True); print("  small path mse err", np.abs(ss8-rs[:5])/rs[:5])
    # gradient path uses R^-1 = M^T M
    a, b = g.gradient_batch(Xs[:50]); ra, rb = O.gradient(st, Xs[:50]); print("  mse_dx err", np.abs(b-rb).max()/np.abs(rb).max())
