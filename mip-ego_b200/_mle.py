"""Multi-restart L-BFGS-B hyper-parameter search (GaussianProcess._optimize_hyperparameter,
gpr.py:964-1101) on top of the batched CUDA likelihood.

scipy's `fmin_l_bfgs_b` is the reference's own optimiser (gpr.py:1039) and is kept unmodified:
  * `sequential` mode reproduces the reference loop exactly -- one restart after the other,
    shared evaluation budget, `wait_iter` early stop (gpr.py:1030-1064) -- each objective call
    being a batch-of-1 `mgp_nll_batch`;
  * `batched` mode advances every restart's L-BFGS-B together and gathers the pending objective
    calls of all live restarts into ONE `mgp_nll_batch` call per step ("lock-step"): scipy's
    compiled `setulb` routine driven as state machines (_lockstep.py), or, if that private symbol
    is unavailable, one thread per restart around the public `fmin_l_bfgs_b`.
    Each trajectory is bit-for-bit what scipy would produce alone given the same (f, g), but all
    restarts run to their own convergence (the reference may stop early on budget/wait_iter),
    so the result is at least as good as the reference's.
The log10 parametrisation and the *un-chained* gradient (d llf / d theta handed to an optimiser
working on log10 theta) are the reference's.
"""
import threading

import numpy as np
from scipy.optimize import fmin_l_bfgs_b


class _LockStep:
    """Collects one pending evaluation per live worker thread, evaluates them in one batch."""

    def __init__(self, batch_fn, n_workers):
        self.batch_fn = batch_fn
        self.cv = threading.Condition()
        self.live = n_workers
        self.pending = {}
        self.results = {}
        self.error = None
        self.n_batches = 0

    def _flush_locked(self):
        keys = sorted(self.pending)
        P = np.array([self.pending[k] for k in keys])
        try:
            f, g = self.batch_fn(P)
            for i, k in enumerate(keys):
                self.results[k] = (float(f[i]), np.array(g[i], dtype=np.float64))
        except BaseException as e:  # propagate to every waiter
            self.error = e
            for k in keys:
                self.results[k] = None
        self.pending.clear()
        self.n_batches += 1
        self.cv.notify_all()

    def evaluate(self, wid, p):
        with self.cv:
            self.pending[wid] = np.array(p, dtype=np.float64)
            if len(self.pending) == self.live:
                self._flush_locked()
            while wid not in self.results:
                self.cv.wait()
            r = self.results.pop(wid)
            if r is None:
                raise self.error
            return r

    def retire(self, wid):
        with self.cv:
            self.live -= 1
            if self.live > 0 and len(self.pending) == self.live:
                self._flush_locked()


def multistart_lbfgsb(batch_fn, starts, log10bounds, eval_budget, wait_iter, mode="batched"):
    """Minimise -llf over log10(par).

    batch_fn(P) with P (B, n_par) of log10 parameters must return (-llf (B,), -grad (B, n_par)).
    starts: (R, n_par) log10 start points (row 0 = theta0 / warm start).
    Returns (param_opt log10, f_opt = -llf, n_evals, info list).
    """
    starts = np.atleast_2d(np.asarray(starts, dtype=np.float64))
    R = starts.shape[0]
    n_evals = [0]

    if mode == "sequential" or R == 1:
        def func(p):
            n_evals[0] += 1
            f, g = batch_fn(np.asarray(p, dtype=np.float64).reshape(1, -1))
            return float(f[0]), np.asarray(g[0], dtype=np.float64)

        best_p, best_f, wait, infos = None, np.inf, 0, []
        budget = eval_budget
        for it in range(R):
            p, f, info = fmin_l_bfgs_b(func, starts[it], bounds=log10bounds, maxfun=budget)
            infos.append(info)
            if it == 0:
                best_p, best_f = p, f
            elif f <= best_f:
                best_p, best_f, wait = p, f, 0
            else:
                wait += 1
            budget -= info["funcalls"]
            if budget <= 0 or wait >= wait_iter:
                break
        return best_p, best_f, n_evals[0], infos

    from . import _lockstep
    if _lockstep.available():
        # thread-free engine: scipy's own setulb driven as R state machines (bit-identical to
        # fmin_l_bfgs_b, verified at first use), one batched likelihood call per step
        out, _ = _lockstep.run(batch_fn, starts, log10bounds, eval_budget)
        best = None
        for wid in range(R):
            p, f, info = out[wid]
            n_evals[0] += info["funcalls"]
            if best is None or f <= best[1]:
                best = (p, f)
        return best[0], best[1], n_evals[0], [o[2] for o in out]

    ls = _LockStep(batch_fn, R)
    out = [None] * R

    def worker(wid):
        try:
            def func(p):
                return ls.evaluate(wid, p)
            out[wid] = fmin_l_bfgs_b(func, starts[wid], bounds=log10bounds, maxfun=eval_budget)
        except BaseException as e:
            out[wid] = e
        finally:
            ls.retire(wid)

    threads = [threading.Thread(target=worker, args=(w,), daemon=True) for w in range(R)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for o in out:
        if isinstance(o, BaseException):
            raise o
    best = None
    for wid in range(R):  # first restart wins ties, like the sequential loop's `<=` from row 0
        p, f, info = out[wid]
        n_evals[0] += info["funcalls"]
        if best is None or f <= best[1]:
            best = (p, f)
    return best[0], best[1], n_evals[0], [o[2] for o in out]
