"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/mgp.h declares; the Python classes mirror the reference's surface; no compute calls."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "mgp.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mgp_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from mipego_b200 import _cabi
    lib = _cabi.load()
    syms = _header_symbols()
    assert len(syms) >= 19
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(_cabi.SIGNATURES) == syms
    assert lib.mgp_abi_version() == _cabi.ABI_VERSION == 4


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    import mipego_b200 as mb
    from mipego_b200 import _cabi
    gp = mb.GaussianProcess(mean=mb.constant_trend(2, beta=None), theta0=[0.1] * 2, thetaL=[1e-3] * 2,
                            thetaU=[1.0] * 2)
    X = np.random.default_rng(0).uniform(size=(10, 2))
    with pytest.raises(_cabi.MgpError):
        gp.fit(X, X.sum(1))


def test_class_surface_matches_reference_protocol():
    import mipego_b200 as mb
    # BayesOpt.py:27-29 keys on a class-level `plugin`; UCB must not have it
    for cls in (mb.EI, mb.PI, mb.EpsilonPI, mb.MGFI):
        assert hasattr(cls, "plugin")
    assert not hasattr(mb.UCB, "plugin")
    # constructible without a model (BayesOpt.py:80-82 reads default t / alpha)
    assert mb.MGFI().t == 1 and mb.UCB().alpha == 0.5
    assert mb.MGFI(t=100).t == 22.36
    assert mb.PI().epsilon == 0.0
    with pytest.raises(AssertionError):
        mb.EpsilonPI(epsilon=0)
    with pytest.raises(AssertionError):
        mb.UCB(alpha=-1)
    with pytest.raises(TypeError):
        mb.UCB(plugin=1.0)
    gp = mb.GaussianProcess(mean=mb.constant_trend(3, beta=None), theta0=[0.1] * 3, thetaL=[1e-3] * 3,
                            thetaU=[1.0] * 3, nugget=1e-6)
    assert gp.is_fitted is False and gp.estimation_mode == "noisy" and gp.estimate_trend
    assert hasattr(gp, "gradient") and hasattr(gp, "predict") and hasattr(gp, "fit") and hasattr(gp, "update")
    assert mb.GaussianProcess(mean=mb.constant_trend(3, beta=0.5), theta0=[0.1] * 3, thetaL=[1e-3] * 3,
                              thetaU=[1.0] * 3).estimation_mode == "noiseless"
    with pytest.raises(ValueError):
        mb.GaussianProcess(mean=mb.constant_trend(3), theta0=[0.1] * 3, thetaL=[1e-3] * 3, thetaU=[np.inf] * 3)


def test_lockstep_driver_matches_sequential_scipy():
    """The threaded lock-step driver feeds each L-BFGS-B the same (f, g) it would get alone."""
    from scipy.optimize import fmin_l_bfgs_b
    from mipego_b200._mle import multistart_lbfgsb

    def f1(p):
        return float(np.sum((p - 0.3) ** 4) + np.sum(np.cos(3 * p))), 4 * (p - 0.3) ** 3 - 3 * np.sin(3 * p)

    calls = []

    def batch_fn(P):
        calls.append(len(P))
        out = [f1(p) for p in P]
        return np.array([o[0] for o in out]), np.array([o[1] for o in out])

    rng = np.random.default_rng(0)
    starts = rng.uniform(-2, 2, (7, 4))
    bounds = np.c_[-2 * np.ones(4), 2 * np.ones(4)]
    p, f, ne, infos = multistart_lbfgsb(batch_fn, starts, bounds, 500, 5, mode="batched")
    singles = [fmin_l_bfgs_b(f1, s, bounds=bounds, maxfun=500) for s in starts]
    best = min(singles, key=lambda r: r[1])
    assert f == best[1] and np.array_equal(p, best[0])
    assert ne == sum(s[2]["funcalls"] for s in singles)
    assert max(calls) == 7 and len(calls) < ne
    p2, f2, ne2, _ = multistart_lbfgsb(batch_fn, starts, bounds, 500, 100, mode="sequential")
    assert f2 == f


def test_integration_doc_stub_signatures():
    """The ctypes stub printed in INTEGRATION.md binds the real library with the same prototypes as
    mipego_b200/_cabi.py (the GPU suite also runs its demo against the oracle)."""
    import ctypes
    txt = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"<!-- ctypes-stub:begin -->\s*```python\n(.*?)```\s*<!-- ctypes-stub:end -->", txt, flags=re.S)
    assert m
    ns = {}
    exec(compile(m.group(1), "INTEGRATION.md", "exec"), ns)
    from mipego_b200 import _cabi
    lib = ns["bind"](_cabi.LIB_PATH)
    for name, (res, args) in ns["PROTOTYPES"].items():
        want_res, want_args = _cabi.SIGNATURES[name]
        assert res is want_res, name
        assert len(args) == len(want_args), name
        for a, b in zip(args, want_args):
            assert ctypes.sizeof(a) == ctypes.sizeof(b), name
            # pointer vs. scalar must agree (a double passed where a pointer is expected was the old doc bug)
            assert issubclass(a, ctypes._Pointer) == issubclass(b, ctypes._Pointer) or \
                {a, b} <= {ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)} or a is b, (name, a, b)
        assert getattr(lib, name).argtypes == args
