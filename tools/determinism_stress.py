"""Bitwise repeatability of every entry point under recycled device memory: each operation is run
on a fresh handle after big allocations have been made and freed (so that cudaMalloc returns stale
non-zero memory) and must reproduce the first run exactly.  Catches reads of uninitialised device
memory and stream-ordering races (one was found this way: the GEMM scheduler words)."""
import gc
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mipego_b200 as mb


def data(n, d, seed):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-5, 5, (n, d))
    y = (X ** 2).sum(1) + 3 * np.sin(X[:, 0]) + 0.5 * X[:, 1]
    return X, (y - y.mean()) / y.std(), rng


def dirty(it):
    n, d = 3072, 16
    X, y, rng = data(n, d, 100 + it)
    gp = mb.GaussianProcess(mean=mb.constant_trend(d, beta=None), corr="squared_exponential",
                            theta0=0.01 * np.ones(d), thetaL=1e-5 * np.ones(d), thetaU=100 * np.ones(d), nugget=1e-6)
    par = np.r_[(0.15 / d) * rng.uniform(0.5, 1.5, d), 0.8]
    gp.fit_fixed(X, y, par)
    gp.log_likelihood_batch(np.tile(par, (6, 1)) * rng.uniform(0.8, 1.2, (6, d + 1)))
    Xs = rng.uniform(-5, 5, (120000, d))
    c = mb.MGFI(model=gp, minimize=True, t=2.0)
    c.evaluate(Xs, params=[1.0, 2.0, 3.0])
    c.evaluate(Xs[:3000], return_dx=True)
    del gp, c
    gc.collect()


def run_case(case):
    n, d, kind, trend, lik, nug, ne = case
    X, y, rng = data(n, d, 7)
    mean = mb.constant_trend(d, beta=None) if trend == "const" else (
        mb.constant_trend(d, beta=0.1) if trend == "const_fixed" else mb.linear_trend(d, beta=None))
    gp = mb.GaussianProcess(mean=mean, corr=kind, theta0=0.1 * np.ones(d), thetaL=1e-5 * np.ones(d),
                            thetaU=100 * np.ones(d), nugget=nug, noise_estim=ne, likelihood=lik)
    gp._check_data(X, y)
    theta = (0.3 / d) * (4.0 if kind == "matern" else 1.0) * rng.uniform(0.7, 1.4, d)
    extra = []
    if lik == "restricted" or (nug and not ne):
        extra.append(0.8)
    if ne:
        extra.append(0.9 if lik == "concentrated" else 0.01)
    par = np.r_[theta, extra]
    P = par * rng.uniform(0.8, 1.25, (6, par.size))
    out = list(gp.log_likelihood_batch(P, eval_grad=True))
    gp.fit_fixed(X, y, par)
    Xs = rng.uniform(-5, 5, (5000, d))
    out += list(gp.predict(Xs, eval_MSE=True))
    out.append(gp.predict(Xs[:300]))
    ei = mb.EI(model=gp, minimize=True)
    out.append(ei.evaluate(Xs))
    out += list(ei.evaluate(Xs[:700], return_dx=True))
    out += list(ei.evaluate(Xs[:5], return_dx=True))            # matrix-vector path
    out += list(ei.topk(Xs, k=9))
    out += list(mb.MGFI(model=gp, minimize=False, t=1.5).topk(Xs, k=1, params=[0.5, 1.5, 4.0]))
    out += list(gp.gradient_batch(Xs[:64]))
    return [np.array(o) for o in out]


CASES = [(640, 8, "squared_exponential", "const", "concentrated", 0, False),
         (1100, 20, "squared_exponential", "const", "concentrated", 1e-6, False),
         (700, 5, "matern", "const_fixed", "concentrated", 1e-6, False),
         (520, 6, "matern", "const", "restricted", 1e-4, False),
         (900, 7, "squared_exponential", "const", "concentrated", 0, True),
         (800, 9, "squared_exponential", "linear", "concentrated", 1e-6, False),
         (600, 4, "matern", "linear", "restricted", 1e-4, False),
         (300, 3, "absolute_exponential", "const", "concentrated", 1e-6, False)]

def main(reps=4, verbose=True):
    refs = [run_case(c) for c in CASES]
    bad = 0
    for it in range(reps):
        dirty(it)
        for ci, c in enumerate(CASES):
            r = run_case(c)
            for k, (a, b) in enumerate(zip(r, refs[ci])):
                if not np.array_equal(a, b, equal_nan=True):
                    bad += 1
                    if verbose:
                        print("MISMATCH rep", it, "case", ci, c, "output", k, "max abs diff",
                              np.nanmax(np.abs(a.astype(float) - b.astype(float))))
    if verbose:
        print("determinism stress: %d cases x %d reps, mismatches: %d" % (len(CASES), reps, bad))
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(int(sys.argv[1]) if len(sys.argv) > 1 else 4) else 0)
