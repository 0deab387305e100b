"""Lock-step multi-start L-BFGS-B without threads.

scipy's `fmin_l_bfgs_b` (the reference's optimiser for both the hyper-parameter MLE, gpr.py:1039, and
the acquisition restarts, optimizer/utils.py:39) is a thin Python loop around the reverse-
communication routine `setulb`: call it, and whenever it answers "need f and g at x", evaluate and
call again (scipy/optimize/_lbfgsb_py.py, `_minimize_lbfgsb`).  `Trajectory` below is that loop
turned into a state machine around the SAME compiled routine, so that R trajectories can be advanced
together and their pending evaluations served by ONE batched GPU call per step.  With threads (the
first implementation, kept as the fallback in _mle.py / optimizer.py) the GIL hand-offs cost 50-200 ms
per step for 100-800 restarts; here a step costs the batched evaluation plus R cheap C calls.

`setulb` is a private scipy symbol.  `available()` therefore replays a few problems through both
`fmin_l_bfgs_b` and this module and requires bit-identical answers (x, f, number of evaluations and
iterations, warning flag, task message) before the engine is used; otherwise callers fall back to
the threaded driver.
"""
import numpy as np

try:  # private module path of scipy >= 1.15 (C translation of L-BFGS-B 3.0)
    from scipy.optimize import _lbfgsb_py as _py
    from scipy.optimize import fmin_l_bfgs_b as _fmin
    _setulb = _py._lbfgsb.setulb
    _INT = np.int64 if getattr(_py, "HAS_ILP64", False) else np.int32
    _STATUS, _TASK = _py.status_messages, _py.task_messages
except Exception:  # noqa: BLE001 -- any layout change disables the engine
    _setulb = None

_checked = None


class Trajectory(object):
    """One L-BFGS-B run, mirroring `_minimize_lbfgsb` (defaults of fmin_l_bfgs_b: m=10, maxls=20)."""

    def __init__(self, x0, bounds, factr=1e7, pgtol=1e-5, maxfun=15000, maxiter=15000, m=10, maxls=20):
        x0 = np.asarray(x0, dtype=np.float64).ravel()
        n = x0.size
        b = np.asarray(bounds, dtype=np.float64).reshape(n, 2)
        lo, up = b[:, 0], b[:, 1]
        if np.any(lo > up):
            raise ValueError("LBFGSB - one of the lower bounds is greater than an upper bound.")
        x0 = np.clip(x0, lo, up)
        self.nbd = np.zeros(n, dtype=_INT)
        self.low = np.zeros(n)
        self.up = np.zeros(n)
        code = {(0, 0): 0, (1, 0): 1, (1, 1): 2, (0, 1): 3}
        for i in range(n):
            fl, fu = int(np.isfinite(lo[i])), int(np.isfinite(up[i]))
            if fl:
                self.low[i] = lo[i]
            if fu:
                self.up[i] = up[i]
            self.nbd[i] = code[(fl, fu)]
        self.m, self.maxls, self.maxfun, self.maxiter = m, maxls, maxfun, maxiter
        self.factr = (factr * np.finfo(float).eps) / np.finfo(float).eps   # ftol round trip of scipy
        self.pgtol = pgtol
        self.x = np.array(x0, dtype=np.float64)
        self.f = np.array(0.0, dtype=np.float64)
        self.g = np.zeros(n)
        self.wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m)
        self.iwa = np.zeros(3 * n, dtype=_INT)
        self.task = np.zeros(2, dtype=_INT)
        self.ln_task = np.zeros(2, dtype=_INT)
        self.lsave = np.zeros(4, dtype=_INT)
        self.isave = np.zeros(44, dtype=_INT)
        self.dsave = np.zeros(29)
        self.nit, self.nfev = 0, 0
        self.last_x, self.last_f, self.last_g = None, None, None   # ScalarFunction's cache
        self.done = False

    def advance(self):
        """Run until an evaluation at self.x is needed (returns True) or the run ends (False)."""
        while True:
            self.g = np.asarray(self.g, dtype=np.float64)
            _setulb(self.m, self.x, self.low, self.up, self.nbd, self.f, self.g, self.factr, self.pgtol,
                    self.wa, self.iwa, self.task, self.lsave, self.isave, self.dsave, self.maxls, self.ln_task)
            if self.task[0] == 3:
                if self.last_x is not None and np.array_equal(self.x, self.last_x):
                    self.f, self.g = self.last_f, self.last_g      # same x: no new evaluation
                    continue
                return True
            elif self.task[0] == 1:
                self.nit += 1
                if self.nit >= self.maxiter:
                    self.task[0], self.task[1] = 5, 504
                elif self.nfev > self.maxfun:
                    self.task[0], self.task[1] = 5, 502
            else:
                self.done = True
                return False

    def feed(self, f, g):
        self.nfev += 1
        self.last_x = self.x.copy()
        self.last_f = float(f)
        self.last_g = np.array(g, dtype=np.float64).ravel()
        self.f, self.g = self.last_f, self.last_g

    def result(self):
        """(x, f, info) exactly as fmin_l_bfgs_b returns them."""
        if self.task[0] == 4:
            warnflag = 0
        elif self.nfev > self.maxfun or self.nit >= self.maxiter:
            warnflag = 1
        else:
            warnflag = 2
        msg = _STATUS[int(self.task[0])] + ": " + _TASK[int(self.task[1])]
        return self.x, float(self.f), {"grad": self.g, "task": msg, "funcalls": self.nfev, "nit": self.nit,
                                        "warnflag": warnflag}


def run(batch_fn, starts, bounds, maxfun, factr=1e7, pgtol=1e-5, with_keys=False):
    """Advance len(starts) trajectories in lock-step.  batch_fn(P) -> (f (k,), g (k, n)) for the k
    pending points (called as batch_fn(P, keys) when with_keys, keys = trajectory indices).
    Returns the list of (x, f, info) in the order of `starts`, and the number of batched calls."""
    trajs = [Trajectory(s, bounds, factr=factr, pgtol=pgtol, maxfun=maxfun) for s in starts]
    live = list(range(len(trajs)))
    n_batches = 0
    while live:
        need = [k for k in live if trajs[k].advance()]
        live = need
        if not need:
            break
        P = np.array([trajs[k].x for k in need])
        f, g = batch_fn(P, need) if with_keys else batch_fn(P)
        n_batches += 1
        for i, k in enumerate(need):
            trajs[k].feed(f[i], g[i])
    return [t.result() for t in trajs], n_batches


def available():
    """True iff the engine reproduces fmin_l_bfgs_b bit for bit on a few probe problems."""
    global _checked
    if _checked is not None:
        return _checked
    _checked = False
    if _setulb is None:
        return False
    try:
        rng = np.random.default_rng(12345)
        ok = True
        for n, maxfun, factr, pgtol in ((3, 200, 1e7, 1e-5), (6, 12, 1e6, 1e-8), (2, 300, 1e6, 1e-8)):
            A = rng.standard_normal((n, n))
            A = A @ A.T + 0.1 * np.eye(n)
            c = rng.standard_normal(n)

            def fg(x, A=A, c=c):
                x = np.asarray(x)
                f = 0.5 * x @ A @ x - c @ x + np.sum(np.cos(2 * x))
                return f, A @ x - c - 2 * np.sin(2 * x)
            b = np.c_[-np.ones(n), 0.7 * np.ones(n)]
            x0 = rng.uniform(-1, 0.7, n)
            xr, fr, dr = _fmin(fg, x0, bounds=b, maxfun=maxfun, factr=factr, pgtol=pgtol)
            (xm, fm, dm), = run(lambda P: tuple(map(np.array, zip(*[fg(p) for p in P]))), [x0], b, maxfun,
                                factr=factr, pgtol=pgtol)[0]
            ok = ok and np.array_equal(xr, xm) and fr == fm and dr["funcalls"] == dm["funcalls"] and \
                dr["nit"] == dm["nit"] and dr["warnflag"] == dm["warnflag"] and dr["task"] == dm["task"] and \
                np.array_equal(dr["grad"], dm["grad"])
        _checked = bool(ok)
    except Exception:  # noqa: BLE001
        _checked = False
    return _checked
