"""Infill criteria: drop-ins for mipego.InfillCriteria (InfillCriteria.py:15-272) evaluated by the
fused posterior + acquisition CUDA kernels (mgp_acq / mgp_acq_topk in include/mgp.h).

Same classes and constructor keywords (`model, minimize, plugin, t | alpha | epsilon`), same
`__call__(X, return_dx=False)` return types: a python float for one point (EpsilonPI: a (1,)
ndarray, as in the reference), `(value, (d,1) ndarray)` with `return_dx`.  Unlike the reference,
whose EI / EpsilonPI / MGFI raise on more than one candidate, an (m, d) input returns
the (m,) values of the single-point rule applied row-wise (and an (m, d) gradient).
A class-level `plugin` property exists on EI/EpsilonPI/PI/MGFI and not on UCB -- the BO loop keys
on that (BayesOpt.py:27-29).
"""
import numpy as np

from . import _cabi


class InfillCriteria(object):
    _kind = None

    def __init__(self, model=None, minimize=True):
        self.model = model
        self.minimize = minimize

    @property
    def model(self):
        return self._model

    @model.setter
    def model(self, model):
        if model is not None:
            self._model = model
            assert hasattr(self._model, "predict")
        else:
            self._model = None

    def check_X(self, X):
        X = np.asarray(X)
        if X.dtype == object:
            X = np.array(X.tolist(), dtype=np.float64)
        return np.ascontiguousarray(np.atleast_2d(X), dtype=np.float64)

    def _param(self):
        return 0.0

    def _plugin_user(self):
        return 0.0

    def _handle(self):
        m = self._model
        if m is None or not hasattr(m, "_ensure_device_state"):
            raise TypeError("mipego_b200 criteria need a mipego_b200.GaussianProcess model "
                            "(no CPU fallback for other surrogates)")
        m._ensure_device_state()
        return m._handle()

    # ---- batched entry points (new capability) ----------------------------------------------
    def evaluate(self, X, params=None, return_dx=False, out=None, out_dx=None):
        """Values (n_sets, m) [and gradients (n_sets, m, d)] for `params` (a list of t / alpha /
        epsilon values sharing one posterior evaluation; default: this object's own parameter).
        `out` / `out_dx`: optional preallocated C-contiguous float64 result arrays (e.g. views of
        pinned host memory, so that the device-to-host copies run asynchronously)."""
        X = self.check_X(X)
        h = self._handle()
        m, d = X.shape
        if d != self._model.X.shape[1]:
            raise ValueError("feature count mismatch")
        P = _cabi.as_f64([self._param()] if params is None else params).ravel()
        ns = P.size

        def _buf(a, shape):
            if a is None:
                return np.empty(shape)
            if a.shape != shape or a.dtype != np.float64 or not a.flags.c_contiguous:
                raise ValueError("output buffer must be a C-contiguous float64 array of shape %r" % (shape,))
            return a
        val = _buf(out, (ns, m))
        dx = _buf(out_dx, (ns, m, d)) if return_dx else None
        h.check(h.lib.mgp_acq(h.h, m, _cabi.dptr(X), _cabi.ACQ[self._kind], ns, _cabi.dptr(P),
                              float(self._plugin_user()), 1 if self.minimize else 0,
                              _cabi.dptr(val), _cabi.dptr(dx)))
        return (val, dx) if return_dx else val

    def topk(self, X, k=1, params=None):
        """k best rows of X per parameter set: values (n_sets, k) descending, indices (n_sets, k)."""
        X = self.check_X(X)
        h = self._handle()
        m, d = X.shape
        P = _cabi.as_f64([self._param()] if params is None else params).ravel()
        ns = P.size
        k = int(min(k, m))
        val = np.empty((ns, k))
        idx = np.empty((ns, k), dtype=np.int64)
        h.check(h.lib.mgp_acq_topk(h.h, m, _cabi.dptr(X), _cabi.ACQ[self._kind], ns, _cabi.dptr(P),
                                   float(self._plugin_user()), 1 if self.minimize else 0, k,
                                   _cabi.dptr(val), idx.ctypes.data_as(_cabi.c_lp)))
        return val, idx

    # ---- device-resident entry points: torch CUDA tensors in, torch CUDA tensors out ----------
    # (torch is plumbing for device memory and streams only; work is enqueued on torch's current
    # stream through mgp_acq_dev / mgp_acq_topk_dev and is NOT synchronised)
    def _dev_args(self, X, params):
        import torch
        h = self._handle()
        if not (X.is_cuda and X.dtype == torch.float64 and X.is_contiguous() and X.dim() == 2):
            raise ValueError("X must be a contiguous 2-D float64 CUDA tensor")
        if X.device.index != h.device:
            raise ValueError("X is on cuda:%d, the model on cuda:%d" % (X.device.index, h.device))
        if X.shape[1] != self._model.X.shape[1]:
            raise ValueError("feature count mismatch")
        P = _cabi.as_f64([self._param()] if params is None else params).ravel()
        stream = torch.cuda.current_stream(X.device).cuda_stream
        return torch, h, P, stream

    def evaluate_device(self, X, params=None, out=None, return_dx=False, out_dx=None):
        """Values (n_sets, m) [and, with return_dx, gradients (n_sets, m, d)] as CUDA tensors."""
        torch, h, P, stream = self._dev_args(X, params)
        m, d = X.shape
        if out is None:
            out = torch.empty((P.size, m), dtype=torch.float64, device=X.device)
        if return_dx and out_dx is None:
            out_dx = torch.empty((P.size, m, d), dtype=torch.float64, device=X.device)
        h.check(h.lib.mgp_acq_dev(h.h, m, X.data_ptr(), _cabi.ACQ[self._kind], P.size, _cabi.dptr(P),
                                  float(self._plugin_user()), 1 if self.minimize else 0,
                                  out.data_ptr(), out_dx.data_ptr() if return_dx else None, stream))
        return (out, out_dx) if return_dx else out

    def topk_device(self, X, k=1, params=None, out=None):
        torch, h, P, stream = self._dev_args(X, params)
        m = X.shape[0]
        k = int(min(k, m))
        if out is None:
            out = (torch.empty((P.size, k), dtype=torch.float64, device=X.device),
                   torch.empty((P.size, k), dtype=torch.int64, device=X.device))
        h.check(h.lib.mgp_acq_topk_dev(h.h, m, X.data_ptr(), _cabi.ACQ[self._kind], P.size,
                                       _cabi.dptr(P), float(self._plugin_user()),
                                       1 if self.minimize else 0, k, out[0].data_ptr(),
                                       out[1].data_ptr(), stream))
        return out

    # ---- the reference's call protocol ---------------------------------------------------------
    def __call__(self, X, return_dx=False, dx=None):
        if dx is not None:  # older spelling (backup/test_acquisition.py:45)
            return_dx = bool(dx)
        X = self.check_X(X)
        n_sample = X.shape[0]
        out = self.evaluate(X, return_dx=return_dx)
        val, grad = (out if return_dx else (out, None))
        val = val[0]
        if n_sample == 1:
            f_value = self._scalar(val)
            if return_dx:
                return f_value, grad[0, 0].reshape(-1, 1)
            return f_value
        if return_dx:
            return val, grad[0]
        return val

    def _scalar(self, val):
        return float(val[0])


class ImprovementBased(InfillCriteria):
    def __init__(self, plugin=None, **kwargs):
        super().__init__(**kwargs)
        self.plugin = plugin

    @property
    def plugin(self):
        return self._plugin

    @plugin.setter
    def plugin(self, plugin):
        # InfillCriteria.py:64-73: stored negated when maximising
        if plugin is None:
            if self._model is not None:
                self._plugin = np.min(self._model.y) if self.minimize \
                    else -1.0 * np.max(self._model.y)
            else:
                self._plugin = None
        else:
            self._plugin = plugin if self.minimize else -1.0 * plugin

    def _plugin_user(self):
        # the C ABI takes the user-facing value and applies the sign itself
        if self._plugin is None:
            self.plugin = None
        return self._plugin if self.minimize else -1.0 * self._plugin


class UCB(InfillCriteria):
    _kind = "UCB"

    def __init__(self, alpha=0.5, **kwargs):
        super().__init__(**kwargs)
        self.alpha = alpha

    @property
    def alpha(self):
        return self._alpha

    @alpha.setter
    def alpha(self, alpha):
        assert alpha > 0
        self._alpha = alpha

    def _param(self):
        return self._alpha


class EI(ImprovementBased):
    _kind = "EI"


class EpsilonPI(ImprovementBased):
    _kind = "PI"

    def __init__(self, epsilon=1e-10, **kwargs):
        super(EpsilonPI, self).__init__(**kwargs)
        self.epsilon = epsilon

    @property
    def epsilon(self):
        return self._epsilon

    @epsilon.setter
    def epsilon(self, eps):
        assert eps > 0
        self._epsilon = eps

    def _param(self):
        return self._epsilon

    def _scalar(self, val):
        return np.array([val[0]])  # the reference returns a (1,) array here (no `sum`)


class PI(EpsilonPI):
    """Probability of improvement = EpsilonPI with epsilon = 0.  The reference's PI cannot be
    constructed (its epsilon setter asserts eps > 0, InfillCriteria.py:164,193-194 -- SURVEY Q3);
    here it can."""

    def __init__(self, **kwargs):
        kwargs.pop("epsilon", None)
        ImprovementBased.__init__(self, **kwargs)
        self._epsilon = 0.0


class MGFI(ImprovementBased):
    _kind = "MGFI"

    def __init__(self, t=1, **kwargs):
        super().__init__(**kwargs)
        self.t = t

    @property
    def t(self):
        return self._t

    @t.setter
    def t(self, t):
        assert t > 0
        self._t = min(t, 22.36)

    def _param(self):
        return self._t


class GEI(ImprovementBased):
    def __init__(self, g=1, **kwargs):
        super().__init__(**kwargs)
        self.g = int(g)

    def __call__(self, X, return_dx=False):
        raise NotImplementedError  # as in the reference (InfillCriteria.py:270-272)
