"""End-to-end BBOB driver in the shape of the reference's benchmark/run_sequential.py:69-106: the
reference's OWN `BO` loop (mipego/base.py) and its vendored COCO test functions
(benchmark/bbobbenchmarks.py:2120 `instantiate`) with the CUDA drop-ins as model, criteria and
acquisition maximiser.  Simple kriging (constant_trend(beta=0)), Matern-3/2, nugget 0, EI, DoE = 10 d,
as run_sequential.py:81-106 configures it.

Needs the reference tree (objective functions + BO loop are the reference's; neither is rebuilt
here): /root/reference in the build container, or the temporary copy of tools/ship_reference_run.sh
on the GPU box.

    python tools/bbob_driver.py [--dim 10] [--fids 1,8,15,21] [--iters 30] [--c5] [--out file.json]

--c5 adds one BO iteration at BASELINE configs[4] scale: n = 8192 evaluated points of a BBOB function
in d = 40, a (budget-capped) MLE refit, and the EI arg-max with the (mu/mu_w, lambda)-CMA-ES population
generator over populations of 4096 candidates (the reference has no CMA-ES acquisition optimiser).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def load():
    from oracle import ref_shim
    if not ref_shim.reference_available():
        raise SystemExit("bbob_driver needs the reference tree (see tools/ship_reference_run.sh)")
    mipego = ref_shim.load_reference()
    sys.path.insert(0, os.path.join(ref_shim.REFERENCE_ROOT, "benchmark"))
    import bbobbenchmarks as bn
    import mipego_b200 as mb
    import mipego.base
    import mipego.BayesOpt
    mipego.base.InfillCriteria = mb.InfillCriteria
    mipego.BayesOpt.InfillCriteria = mb.InfillCriteria
    mipego.base.argmax_restart = mb.argmax_restart       # all restarts of the EI arg-max in lock-step
    return mipego, bn, mb


def make_model(mb, dim, lb, ub, random_start, eval_budget):
    thetaL = 1e-5 * (ub - lb) * np.ones(dim)
    thetaU = 10 * (ub - lb) * np.ones(dim)
    theta0 = np.random.rand(dim) * (thetaU - thetaL) + thetaL
    return mb.GaussianProcess(mean=mb.constant_trend(dim, beta=0), corr="matern", theta0=theta0, thetaL=thetaL,
                              thetaU=thetaU, noise_estim=False, nugget=0, optimizer="BFGS", wait_iter=5,
                              random_start=random_start, eval_budget=eval_budget)


def run_function(mipego, bn, mb, fid, dim, iters, seed):
    from mipego import BO, ContinuousSpace
    np.random.seed(seed)
    f, fopt = bn.instantiate(fid, iinstance=1)
    lb, ub = -5, 5
    n_eval = [0]

    def obj(x):
        n_eval[0] += 1
        return float(f(np.asarray(x, dtype=float)))

    model = make_model(mb, dim, lb, ub, random_start=min(10 * dim, 20), eval_budget=200 * dim)
    opt = BO(search_space=ContinuousSpace([lb, ub]) * dim, obj_fun=obj, model=model, DoE_size=10 * dim,
             max_FEs=10 * dim + iters, verbose=False, n_point=1, minimize=True, acquisition_fun="EI", logger=None)
    t0 = time.perf_counter()
    xopt, fbest, stop = opt.run()
    dt = time.perf_counter() - t0
    h = model._handle()
    return {"fid": fid, "dim": dim, "FEs": n_eval[0], "fbest_minus_fopt": float(fbest - fopt), "seconds": dt,
            "seconds_per_iteration": dt / max(1, iters), "n_final": int(model.X.shape[0]),
            "kernel_launches": int(h.counter("launches")), "appends": int(h.counter("appends"))}


def run_c5(mipego, bn, mb, fid=8, n=8192, dim=40, lam=4096, generations=30, mle_budget=64, seed=40):
    """One BO iteration at configs[4] scale."""
    from mipego_b200.optimizer import cma_es_argmax
    rng = np.random.default_rng(seed)
    f, fopt = bn.instantiate(fid, iinstance=1)
    X = rng.uniform(-5, 5, (n, dim))
    y = np.array([float(f(x)) for x in X])
    np.random.seed(seed)
    model = make_model(mb, dim, -5, 5, random_start=8, eval_budget=mle_budget)
    t0 = time.perf_counter()
    model.fit(X, y)
    t_fit = time.perf_counter() - t0
    ei = mb.EI(model=model, minimize=True, plugin=float(y.min()))
    bounds = np.array([[-5.0, 5.0]] * dim)
    t0 = time.perf_counter()
    xb, fb, info = cma_es_argmax(ei, bounds, lambda_=lam, max_eval=lam * generations, seed=seed,
                                 topk_fn=lambda Z, k: tuple(a[0] for a in ei.topk(Z, k=k)))
    t_acq = time.perf_counter() - t0
    ynew = float(f(xb))
    t0 = time.perf_counter()
    model.update(np.r_[X, xb.reshape(1, -1)], np.r_[y, ynew], reoptimize=False)   # tell: rank-1 append
    t_tell = time.perf_counter() - t0
    return {"fid": fid, "n": n, "dim": dim, "mle": {"seconds": t_fit, "likelihood_evaluations": int(model.eval_count),
                                                       "restarts": 8, "llf": float(np.ravel(model.log_likelihood_)[0])},
            "acquisition": {"optimizer": "(mu/mu_w, lambda)-CMA-ES", "lambda": lam, "candidates": int(info["funcalls"]),
                            "seconds": t_acq, "candidates_per_s": info["funcalls"] / t_acq, "EI_best": float(fb)},
            "tell_append": {"seconds": t_tell, "appends": int(model._handle().counter("appends"))},
            "y_new_minus_ybest": ynew - float(y.min())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dim", type=int, default=10)
    ap.add_argument("--fids", default="1,8,15,21")
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--c5", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    mipego, bn, mb = load()
    out = {"shape": "benchmark/run_sequential.py:69-106 (GP matern, simple kriging, EI, DoE 10 d)", "runs": []}
    for fid in [int(s) for s in args.fids.split(",") if s]:
        r = run_function(mipego, bn, mb, fid, args.dim, args.iters, seed=1000 + fid)
        print(json.dumps(r), flush=True)
        out["runs"].append(r)
    if args.c5:
        out["c5"] = run_c5(mipego, bn, mb)
        print(json.dumps(out["c5"]), flush=True)
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
