"""Batched acquisition maximisers: the callers of the fused posterior + acquisition kernels.

Drop-in for `mipego.optimizer.utils.argmax_restart` (optimizer/utils.py:8-87) with the same
signature and return value, plus the batched front-ends SURVEY section 8f ranks next (N1, N2):

  * `argmax_restart(..., optimizer='BFGS')` -- the reference runs its `n_restart` L-BFGS-B restarts
    one after the other, one candidate per criterion call (optimizer/utils.py:25-52).  Here all
    restarts advance together -- scipy's own compiled L-BFGS-B routine driven as state machines
    (_lockstep.py; verified bit-identical to `fmin_l_bfgs_b` at first use, with a thread-per-restart
    fallback) -- and the pending (value, gradient) requests of all live restarts are served by ONE
    batched criterion call per step ("lock-step").  A candidate's result does not depend on what else is in the batch, so
    each trajectory is bit-for-bit what the sequential loop would produce from the same start.
  * `argmax_restart(..., optimizer='MIES')` -- a vectorised restatement of the continuous-variable
    operators of the reference's MIES (optimizer/mies.py:160-175 recombination, :210-237
    self-adaptive Gaussian mutation, misc.py:116-152 reflection at the box, :174-182 comma
    selection, :256-300 stopping rules); all restarts and all lambda offspring of a generation are
    evaluated in one call (mies.py:306 says "TODO: vectorize this part").
  * `cma_es_argmax` -- a (mu/mu_w, lambda)-CMA-ES population generator for very large lambda
    (BASELINE config 5 names a CMA-ES acquisition optimiser; the reference has none -- its
    OnePlusOne_CMA evaluates one point per step).
  * `batch_argmax_restart` -- ParallelBO's n_point criteria over the same model
    (BayesOpt.py:84-101) maximised together: every batch evaluates all parameter sets from one
    posterior pass (`criterion.evaluate(X, params=[...])`).

`mode='sequential'` keeps the reference's exact control flow (restarts in order, shared evaluation
budget, `wait_iter` early stop); `mode='batched'` (default) runs all restarts to their own
convergence and returns the best, which is at least as good.
Objective functions that are not mipego_b200 criteria are called point by point (that is calling
user code, not a fallback of the GP path).
"""
import functools
import threading

import numpy as np
from numpy.random import rand, randint, randn
from scipy.optimize import fmin_l_bfgs_b

from ._mle import _LockStep

__all__ = ["BoxSpace", "argmax_restart", "batch_argmax_restart", "mies_argmax", "cma_es_argmax",
           "handle_box_constraint", "dynamic_penalty"]


class BoxSpace(object):
    """Minimal continuous search space with the attributes `argmax_restart` and MIES read from
    the reference's ContinuousSpace (SearchSpace.py:278-346): `dim`, `bounds`, `var_name`,
    `C_mask`/`O_mask`/`N_mask`, `id_C`/`id_O`/`id_N`, `sampling(N, method)`.  The reference's own
    ContinuousSpace can be passed instead; this class exists so that the path does not depend on
    the reference package."""

    def __init__(self, bounds, var_name="r"):
        b = np.atleast_2d(np.asarray(bounds, dtype=np.float64))
        if b.shape[1] != 2:
            raise ValueError("bounds must be (dim, 2)")
        if not np.all(b[:, 0] < b[:, 1]):
            raise ValueError("lower bounds must be smaller than upper bounds")
        self.dim = b.shape[0]
        self.bounds = [tuple(r) for r in b.tolist()]
        self._bounds = b.T.copy()
        self.var_name = [var_name + str(i) for i in range(self.dim)]
        self.var_type = ["C"] * self.dim
        self.C_mask = np.ones(self.dim, dtype=bool)
        self.O_mask = np.zeros(self.dim, dtype=bool)
        self.N_mask = np.zeros(self.dim, dtype=bool)
        self.id_C = np.arange(self.dim)
        self.id_O = np.array([], dtype=int)
        self.id_N = np.array([], dtype=int)

    def sampling(self, N=1, method="uniform"):
        lb, ub = self._bounds
        return ((ub - lb) * rand(N, self.dim) + lb).tolist()   # SearchSpace.py:338-342


def _bounds_of(search_space):
    mask = np.nonzero(np.asarray(search_space.C_mask) | np.asarray(search_space.O_mask))[0]
    return np.array([search_space.bounds[i] for i in mask], dtype=np.float64)


def handle_box_constraint(x, lb, ub):
    """Reflection into [lb, ub] (misc.py:116-152, T^r_[a,b] of Li's MIES thesis) for x of shape
    (..., dim)."""
    x = np.asarray(x, dtype=np.float64)
    lb = np.asarray(lb, dtype=np.float64)
    ub = np.asarray(ub, dtype=np.float64)
    y = (x - lb) / (ub - lb)
    fl = np.floor(y)
    even = np.mod(fl, 2) == 0
    yp = np.where(even, np.abs(y - fl), 1.0 - np.abs(y - fl))
    return lb + (ub - lb) * yp


def dynamic_penalty(X, t, equality=None, inquality=None, C=0.5, alpha=1, beta=2, epsilon=0.01,
                    minimize=True):
    """Penalty of the user's constraint functions h(x) = 0 / g(x) <= 0 for a whole population, with
    the reference's keyword names, defaults and formula (mipego/utils.py:13-30, called from
    mies.py:189-193): (C t)^alpha * (sum |h| beyond epsilon + sum max(g, 0)^beta), sign flipped when
    maximising.  The constraint callables are user code and take one point each."""
    pts = list(X)

    def violations(fn):
        return np.array([np.ravel(fn(x)) for x in pts], dtype=np.float64).reshape(len(pts), -1)

    total = np.zeros(len(pts))
    if equality is not None:
        hv = np.abs(violations(equality))
        total += np.where(hv > epsilon, hv, 0.0).sum(axis=1)
    if inquality is not None:
        total += (np.clip(violations(inquality), 0.0, None) ** beta).sum(axis=1)
    scale = (C * t) ** alpha
    return (scale if minimize else -scale) * total


# ------------------------------------------------------------------------------------------------
# objective adapters
# ------------------------------------------------------------------------------------------------
def _unwrap(obj_func):
    """(criterion, return_dx) if obj_func is a mipego_b200 criterion or the functools.partial the
    BO loop builds around one (base.py:554-560); else (None, None)."""
    from .InfillCriteria import InfillCriteria
    kw = {}
    f = obj_func
    while isinstance(f, functools.partial):
        kw = dict(f.keywords or {}, **kw)
        f = f.func
    if isinstance(f, InfillCriteria) and hasattr(f, "evaluate"):
        return f, bool(kw.get("return_dx", kw.get("dx", False)))
    return None, None


class _Objective(object):
    """Batched view of an objective: values(X) -> (m,), values_grads(X) -> ((m,), (m, d)).
    `set_index` selects one of several parameter sets evaluated together."""

    def __init__(self, obj_func, params=None, set_index=0):
        self.obj_func = obj_func
        self.crit, self.return_dx = _unwrap(obj_func)
        self.params, self.set_index = params, set_index
        self.n_calls, self.n_points = 0, 0

    def values(self, X):
        X = np.ascontiguousarray(np.atleast_2d(X), dtype=np.float64)
        self.n_calls += 1
        self.n_points += X.shape[0]
        if self.crit is not None:
            return self.crit.evaluate(X, params=self.params)[self.set_index]
        out = np.empty(X.shape[0])
        for i, x in enumerate(X):
            r = self.obj_func(x.tolist())
            out[i] = float(np.ravel(r[0] if isinstance(r, tuple) else r)[0])
        return out

    def values_grads(self, X):
        X = np.ascontiguousarray(np.atleast_2d(X), dtype=np.float64)
        self.n_calls += 1
        self.n_points += X.shape[0]
        if self.crit is not None:
            v, dx = self.crit.evaluate(X, params=self.params, return_dx=True)
            return v[self.set_index], dx[self.set_index]
        f, g = np.empty(X.shape[0]), np.empty(X.shape)
        for i, x in enumerate(X):
            a, b = self.obj_func(x)
            f[i] = float(np.ravel(a)[0])
            g[i] = np.ravel(b)
        return f, g


# ------------------------------------------------------------------------------------------------
# N2: lock-step multi-start L-BFGS-B
# ------------------------------------------------------------------------------------------------
def _lbfgsb_restarts(objs, starts, bounds, eval_budget, wait_iter, mode, logger=None):
    """objs: one _Objective per problem (all sharing the same criterion object / model);
    starts: (n_problems, n_restart, d).  Returns per problem (xopt list, fopt float)."""
    n_prob, R, d = starts.shape
    results = []
    if mode == "sequential":
        for p, obj in enumerate(objs):
            xs, fs, best, wait, budget = [], [], -np.inf, 0, eval_budget
            for it in range(R):
                def func(x):
                    f, g = obj.values_grads(np.asarray(x, dtype=np.float64).reshape(1, -1))
                    return -float(f[0]), -np.asarray(g[0], dtype=np.float64)
                x_, f_, info = fmin_l_bfgs_b(func, starts[p, it], pgtol=1e-8, factr=1e6,
                                             bounds=bounds, maxfun=budget)
                f_ = -float(f_)
                if logger is not None and info["warnflag"] != 0:
                    logger.debug("L-BFGS-B terminated abnormally with the state: %s" % info)
                if f_ > best:
                    best, wait = f_, 0
                else:
                    wait += 1
                budget -= info["funcalls"]
                xs.append(x_.flatten().tolist())
                fs.append(f_)
                if budget <= 0 or wait >= wait_iter:
                    break
            i = int(np.argsort(fs)[::-1][0])
            results.append((xs[i], fs[i]))
        return results

    # batched: worker id = p * R + it; one batched criterion call serves every pending request
    n_work = n_prob * R
    shared = objs[0].crit is not None and all(o.crit is objs[0].crit for o in objs)

    def batch_fn(P, keys):
        f = np.empty(len(keys))
        g = np.empty((len(keys), d))
        if shared and objs[0].params is not None:
            # all parameter sets from one posterior pass; each request keeps its own set's row
            v, dx = objs[0].crit.evaluate(P, params=objs[0].params, return_dx=True)
            objs[0].n_calls += 1
            objs[0].n_points += len(keys)
            for i, k in enumerate(keys):
                s = objs[k // R].set_index
                f[i], g[i] = v[s, i], dx[s, i]
        else:
            owner = np.array([k // R for k in keys])
            for p in np.unique(owner):
                idx = np.nonzero(owner == p)[0]
                fv, gv = objs[p].values_grads(P[idx])
                f[idx], g[idx] = fv, gv
        return -f, -g

    from . import _lockstep
    if _lockstep.available():
        flat = starts.reshape(n_work, d)
        out, _ = _lockstep.run(batch_fn, flat, bounds, eval_budget, factr=1e6, pgtol=1e-8, with_keys=True)
    else:
        out = _threaded_restarts(batch_fn, starts, bounds, eval_budget, n_work, R)
    for p in range(n_prob):
        best = None
        for it in range(R):
            x_, f_, info = out[p * R + it]
            if logger is not None and info["warnflag"] != 0:
                logger.debug("L-BFGS-B terminated abnormally with the state: %s" % info)
            f_ = -float(f_)
            if best is None or f_ > best[1]:
                best = (x_.flatten().tolist(), f_)
        results.append(best)
    return results


def _threaded_restarts(batch_fn, starts, bounds, eval_budget, n_work, R):
    """Fallback lock-step driver: one thread per restart around the public fmin_l_bfgs_b."""
    ls = _KeyedLockStep(batch_fn, n_work)
    out = [None] * n_work

    def worker(wid):
        try:
            out[wid] = fmin_l_bfgs_b(lambda x: ls.evaluate(wid, x), starts[wid // R, wid % R],
                                     pgtol=1e-8, factr=1e6, bounds=bounds, maxfun=eval_budget)
        except BaseException as e:  # noqa: BLE001 -- re-raised in the caller's thread
            out[wid] = e
        finally:
            ls.retire(wid)

    threads = [threading.Thread(target=worker, args=(w,), daemon=True) for w in range(n_work)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for o in out:
        if isinstance(o, BaseException):
            raise o
    return out


class _KeyedLockStep(_LockStep):
    """_LockStep whose batch function also receives the worker ids of the rows."""

    def _flush_locked(self):
        keys = sorted(self.pending)
        P = np.array([self.pending[k] for k in keys])
        try:
            f, g = self.batch_fn(P, keys)
            for i, k in enumerate(keys):
                self.results[k] = (float(f[i]), np.array(g[i], dtype=np.float64))
        except BaseException as e:  # noqa: BLE001
            self.error = e
            for k in keys:
                self.results[k] = None
        self.pending.clear()
        self.n_batches += 1
        self.cv.notify_all()


# ------------------------------------------------------------------------------------------------
# N1: vectorised MIES (continuous variables)
# ------------------------------------------------------------------------------------------------
def mies_argmax(values_fn, bounds, n_restart=1, max_eval=np.inf, mu_=4, lambda_=10, sigma0=None,
                minimize=False, eq_func=None, ineq_func=None, tolfun=1e-5, x0=None):
    """R = n_restart independent (mu, lambda)-MIES runs on a box, advanced one generation at a
    time with all R * lambda offspring evaluated by ONE call `values_fn(X (m, d)) -> (m,)`.

    Operators follow optimizer/mies.py for continuous variables: uniform initial parents with
    sigma0 = 0.05 (ub - lb) (:72-74, :110-124); parents p1, p2 ~ U{0..mu-1}; intermediate
    recombination of the step sizes and dominant recombination of x where randn > 0.5
    (:160-172); sigma' = sigma exp(tau N + tau' N_i), tau = 1/sqrt(2 d), tau' = 1/sqrt(2 sqrt d)
    (:143-147, :210-218); x' = x + sigma' N_i reflected into the box (:220-225, misc.py:116);
    comma selection of the best mu offspring (:174-182); per-run stops: eval_count > max_eval,
    fitness-history range < tolfun every nbin generations, flat fitness (:256-272).
    Returns (xopt (R, d), fopt (R,), funcalls (R,) ints, stop reasons list)."""
    bounds = np.atleast_2d(np.asarray(bounds, dtype=np.float64))
    lb, ub = bounds[:, 0], bounds[:, 1]
    d, R = len(lb), int(n_restart)
    sigma0 = 0.05 * (ub - lb) if sigma0 is None else np.broadcast_to(np.asarray(sigma0, float), (d,))
    tau, tau_p = 1 / np.sqrt(2 * d), 1 / np.sqrt(2 * np.sqrt(d))
    sign = 1.0 if minimize else -1.0           # internally minimise sign * f
    nbin = int(3 + np.ceil(30.0 * d / lambda_))

    def evaluate(X, it):
        f = np.asarray(values_fn(X), dtype=np.float64).reshape(-1)
        if eq_func is not None or ineq_func is not None:
            f = f + dynamic_penalty(X.tolist(), it + 1, eq_func, ineq_func, minimize=minimize)
        return f

    if x0 is None:
        pop = (ub - lb) * rand(R, mu_, d) + lb
        fit = evaluate(pop.reshape(-1, d), 0).reshape(R, mu_)
    else:
        pop = np.tile(np.asarray(x0, dtype=np.float64).reshape(1, 1, d), (R, mu_, 1))
        f0 = evaluate(pop[:, 0, :], 0)
        fit = np.repeat(f0[:, None], mu_, axis=1)
    sig = np.tile(sigma0, (R, mu_, 1))
    evals = np.full(R, mu_ if x0 is None else 1, dtype=np.int64)
    best_i = np.argmin(sign * fit, axis=1)
    xopt = pop[np.arange(R), best_i].copy()
    fopt = fit[np.arange(R), best_i].copy()
    hist = np.zeros((R, nbin))
    iters = np.zeros(R, dtype=np.int64)
    active = np.ones(R, dtype=bool)
    reasons = [dict() for _ in range(R)]
    first = True
    while True:
        # ---- stop() of every live run (evaluated before each generation, mies.py:303) ----------
        for r in np.nonzero(active)[0]:
            if evals[r] > max_eval:
                reasons[r]["max_eval"] = True
            if not first and iters[r] != 0:
                g = evals[r] / lambda_
                hist[r, int(np.mod(g - 1, nbin))] = last_off[r, 0]
                if np.mod(g, nbin) == 0 and hist[r].max() - hist[r].min() < tolfun:
                    reasons[r]["tolfun"] = True
                j = int(min(np.ceil(0.1 + lambda_ / 4.0), mu_ - 1))
                if last_off[r, 0] == last_off[r, j]:
                    reasons[r]["flatfitness"] = True
            if reasons[r]:
                active[r] = False
        idx = np.nonzero(active)[0]
        if idx.size == 0:
            break
        A = idx.size
        # ---- recombination + mutation, vectorised over (live runs, lambda) ----------------------
        p1 = randint(0, mu_, (A, lambda_))
        p2 = randint(0, mu_, (A, lambda_))
        ar = idx[:, None]
        x1, x2 = pop[ar, p1], pop[ar, p2]
        s1, s2 = sig[ar, p1], sig[ar, p2]
        differ = (p1 != p2)[..., None]
        s = np.where(differ, 0.5 * (s1 + s2), s1)
        take2 = (randn(A, lambda_, d) > 0.5) & differ
        x = np.where(take2, x2, x1)
        if d == 1:
            s = s * np.exp(tau * randn(A, lambda_, 1))
        else:
            s = s * np.exp(tau * randn(A, lambda_, 1) + tau_p * randn(A, lambda_, d))
        xn = handle_box_constraint(x + s * randn(A, lambda_, d), lb, ub)
        f_off = evaluate(xn.reshape(-1, d), int(iters[idx].max())).reshape(A, lambda_)
        evals[idx] += lambda_
        # ---- comma selection ------------------------------------------------------------------
        order = np.argsort(sign * f_off, axis=1, kind="stable")[:, :mu_]
        rows = np.arange(A)[:, None]
        pop[idx] = xn[rows, order]
        sig[idx] = s[rows, order]
        fit[idx] = f_off[rows, order]
        if first:
            last_off = np.zeros((R, lambda_))
            first = False
        # the reference's stop() reads f_offspring AFTER select() without re-sorting: its entry 0 is
        # offspring 0 of the generation, not the best one (mies.py:262-271)
        last_off[idx] = f_off
        better = sign * fit[idx, 0] < sign * fopt[idx]
        upd = idx[better]
        xopt[upd] = pop[upd, 0]
        fopt[upd] = fit[upd, 0]
        iters[idx] += 1
    return xopt, fopt, evals, reasons


# ------------------------------------------------------------------------------------------------
# N1: (mu/mu_w, lambda)-CMA-ES population generator
# ------------------------------------------------------------------------------------------------
def cma_es_argmax(values_fn, bounds, lambda_=4096, max_eval=None, sigma0=0.3, x0=None, tolfun=1e-11,
                  tolx=1e-9, seed=None, topk_fn=None, mu=None):
    """Maximise over a box with a (mu/mu_w, lambda)-CMA-ES (Hansen's standard strategy parameters)
    whose whole population is evaluated by one call per generation.  Candidates are reflected into
    the box before evaluation (the same repair MIES uses).

    values_fn(X (lambda, d)) -> (lambda,) values to maximise.  If `topk_fn(X, k)` is given it must
    return (values (k,), indices (k,)) of the k best rows -- the fused top-k kernel
    (`criterion.topk`, k <= 64) -- so that only mu results travel back to the host; mu then
    defaults to min(lambda/2, 64).
    Returns (xbest (d,), fbest, info dict)."""
    rng = np.random.default_rng(seed)
    bounds = np.atleast_2d(np.asarray(bounds, dtype=np.float64))
    lb, ub = bounds[:, 0], bounds[:, 1]
    d = len(lb)
    lam = int(lambda_)
    mu = int(mu) if mu is not None else (min(lam // 2, 64) if topk_fn is not None else lam // 2)
    w = np.log(mu + 0.5) - np.log(np.arange(1, mu + 1))
    w /= w.sum()
    mueff = 1.0 / np.sum(w ** 2)
    cc = (4 + mueff / d) / (d + 4 + 2 * mueff / d)
    cs = (mueff + 2) / (d + mueff + 5)
    c1 = 2 / ((d + 1.3) ** 2 + mueff)
    cmu = min(1 - c1, 2 * (mueff - 2 + 1 / mueff) / ((d + 2) ** 2 + mueff))
    damps = 1 + 2 * max(0.0, np.sqrt((mueff - 1) / (d + 1)) - 1) + cs
    chiN = np.sqrt(d) * (1 - 1 / (4.0 * d) + 1 / (21.0 * d * d))
    scale = ub - lb
    mean = (lb + scale * rng.random(d)) if x0 is None else np.asarray(x0, dtype=np.float64).copy()
    sigma = float(sigma0)                     # in units of the box width (coordinates are scaled)
    C, pc, ps = np.eye(d), np.zeros(d), np.zeros(d)
    Bm, Dv = np.eye(d), np.ones(d)
    max_eval = 100 * lam if max_eval is None else max_eval
    evals, gen = 0, 0
    xbest, fbest = mean.copy(), -np.inf
    hist = []
    stop = ""
    while evals < max_eval:
        Z = rng.standard_normal((lam, d))
        Y = (Z * Dv) @ Bm.T
        X = handle_box_constraint(mean + sigma * scale * Y, lb, ub)
        if topk_fn is not None:
            fv, order = topk_fn(X, mu)
            order = np.asarray(order, dtype=np.int64)
            fsel = np.asarray(fv, dtype=np.float64)
        else:
            f = np.asarray(values_fn(X), dtype=np.float64).reshape(-1)
            order = np.argsort(-f, kind="stable")[:mu]
            fsel = f[order]
        evals += lam
        gen += 1
        if fsel[0] > fbest:
            fbest, xbest = float(fsel[0]), X[order[0]].copy()
        # the update uses the REPAIRED steps so that the distribution learns the feasible region
        Ysel = (X[order] - mean) / (sigma * scale)
        yw = w @ Ysel
        mean = mean + sigma * scale * yw
        Cinvsqrt_y = Bm @ ((Bm.T @ yw) / Dv)
        ps = (1 - cs) * ps + np.sqrt(cs * (2 - cs) * mueff) * Cinvsqrt_y
        hsig = np.linalg.norm(ps) / np.sqrt(1 - (1 - cs) ** (2 * gen)) / chiN < 1.4 + 2 / (d + 1.0)
        pc = (1 - cc) * pc + hsig * np.sqrt(cc * (2 - cc) * mueff) * yw
        C = (1 - c1 - cmu) * C + c1 * (np.outer(pc, pc) + (1 - hsig) * cc * (2 - cc) * C) \
            + cmu * (Ysel.T * w) @ Ysel
        sigma *= np.exp((cs / damps) * (np.linalg.norm(ps) / chiN - 1))
        C = 0.5 * (C + C.T)
        ev, Bm = np.linalg.eigh(C)
        Dv = np.sqrt(np.maximum(ev, 1e-30))
        hist.append(float(fsel[0]))
        if len(hist) > 10 and max(hist[-10:]) - min(hist[-10:]) < tolfun:
            stop = "tolfun"
            break
        if sigma * Dv.max() < tolx:
            stop = "tolx"
            break
    return xbest, fbest, {"funcalls": evals, "generations": gen, "stop": stop or "max_eval",
                          "sigma": sigma}


# ------------------------------------------------------------------------------------------------
# the reference entry point
# ------------------------------------------------------------------------------------------------
def batch_argmax_restart(obj_funcs, search_space, h=None, g=None, eval_budget=100, n_restart=10,
                         wait_iter=3, optimizer="BFGS", logger=None, mode="batched"):
    """argmax_restart for several criteria over the same model (ParallelBO, BayesOpt.py:84-101).
    Returns (list of xopt, list of fopt).  Start points are drawn in the reference's order: all
    restarts of criterion 0, then criterion 1, ..."""
    objs = [_Objective(f) for f in obj_funcs]
    crits = [o.crit for o in objs]
    # Inside the drivers every batch goes through the MMA path, whatever its size: the library's
    # matrix-vector path for <= 8 candidates rounds differently, and a candidate's value must not
    # depend on how many other restarts happen to be pending ("lock-step" == sequential, bit for bit).
    handles = []
    for c in crits:
        if c is not None:
            hd = c._handle()
            if all(hd is not x for x in handles):
                handles.append(hd)
    previous = [hd.counter("small_path") for hd in handles]      # the user's own setting is restored
    for hd in handles:
        hd.set_option("small_path", 0)
    try:
        return _batch_argmax_restart(objs, crits, search_space, h, g, eval_budget, n_restart, wait_iter,
                                     optimizer, logger, mode)
    finally:
        for hd, old in zip(handles, previous):
            hd.set_option("small_path", old)


def _batch_argmax_restart(objs, crits, search_space, h, g, eval_budget, n_restart, wait_iter, optimizer,
                          logger, mode):
    if all(c is not None for c in crits) and len({id(c._model) for c in crits}) == 1 and \
            len({(type(c), c.minimize, c._plugin_user()) for c in crits}) == 1:
        # same model / class / plugin: one posterior pass evaluates every parameter set
        params = [c._param() for c in crits]
        for i, o in enumerate(objs):
            o.crit, o.params, o.set_index = crits[0], params, i
    n_prob = len(objs)
    bounds = _bounds_of(search_space)
    d = bounds.shape[0]
    starts = np.array([[search_space.sampling(N=1, method="uniform")[0] for _ in range(n_restart)]
                       for _ in range(n_prob)], dtype=object)
    if optimizer == "BFGS":
        if not all(isinstance(v, float) for v in starts.reshape(-1)):
            raise ValueError("BFGS is not supported with mixed variable types.")
        starts = starts.astype(np.float64).reshape(n_prob, n_restart, d)
        res = _lbfgsb_restarts(objs, starts, bounds, eval_budget, wait_iter, mode, logger)
    elif optimizer == "MIES":
        res = []
        for o in objs:
            if mode == "sequential":
                xs, fs, best, wait, budget = [], [], -np.inf, 0, eval_budget
                for it in range(n_restart):
                    x_, f_, ev, _ = mies_argmax(o.values, bounds, 1, max_eval=budget, eq_func=h,
                                                ineq_func=g)
                    if f_[0] > best:
                        best, wait = f_[0], 0
                    else:
                        wait += 1
                    budget -= int(ev[0])
                    xs.append(x_[0].tolist())
                    fs.append(float(f_[0]))
                    if budget <= 0 or wait >= wait_iter:
                        break
                i = int(np.argsort(fs)[::-1][0])
                res.append((xs[i], fs[i]))
            else:
                x_, f_, ev, _ = mies_argmax(o.values, bounds, n_restart, max_eval=eval_budget,
                                            eq_func=h, ineq_func=g)
                i = int(np.argmax(f_))
                res.append((x_[i].tolist(), float(f_[i])))
    else:
        raise ValueError("optimizer should be 'BFGS' or 'MIES'")
    return [r[0] for r in res], [r[1] for r in res]


def argmax_restart(obj_func, search_space, h=None, g=None, eval_budget=100, n_restart=10,
                   wait_iter=3, optimizer="BFGS", logger=None, mode="batched"):
    """optimizer/utils.py:8-87: returns (xopt list, fopt float)."""
    xs, fs = batch_argmax_restart([obj_func], search_space, h, g, eval_budget, n_restart, wait_iter,
                                  optimizer, logger, mode)
    return xs[0], fs[0]
