"""Container-only loader for the *real* reference (wangronin/MIP-EGO, /root/reference).

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (`mip-ego_b200/`) may import this
module.  It exists so that `oracle/gen_golden.py` can run the unmodified reference
implementation of the GP hot path (mipego/GaussianProcess/gpr.py, mipego/InfillCriteria.py)
in this container and freeze its outputs as golden fixtures under `tests/golden/`.
`/root/reference` does not exist on the GPU box, so nothing that runs there may call this.

The reference targets numpy 1.21 / sklearn 1.0; five in-memory shims
make it import on numpy 2.3 / sklearn 1.9 without touching its sources:
  1. np.int alias (gpr.py:58)
  2. fake `pyDOE.lhs` (SearchSpace.py:10,346) -- DoE only, off the hot path
  3. stub matplotlib modules (GaussianProcess/utils.py:12-16)
  4. stub mipego.GaussianProcess.gpy (needs gpytorch; GaussianProcess/__init__.py:10)
  5. manhattan_distances(sum_over_features=False) re-implemented (gpr.py:317,453,534)
"""
import os
import sys
import types

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_reference():
    """$MIPEGO_REFERENCE_ROOT, else /root/reference (the build container), else the temporary copy
    tools/ship_reference_run.sh places under baseline/_ref/ for ONE run on the GPU box (git-ignored,
    removed afterwards -- reference sources are never part of this repository)."""
    cands = [os.environ.get("MIPEGO_REFERENCE_ROOT"), "/root/reference",
             os.path.join(_ROOT, "baseline", "_ref", "reference")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "mipego")):
            return c
    return cands[0] or "/root/reference"


REFERENCE_ROOT = _find_reference()


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mipego"))


def _lhs(n, samples=None, criterion=None, iterations=None):
    samples = n if samples is None else samples
    rng = np.random
    u = (rng.rand(samples, n) + np.arange(samples)[:, None]) / samples
    for j in range(n):
        u[:, j] = u[rng.permutation(samples), j]
    return u


def _manhattan(X, Y=None, sum_over_features=True):
    X = np.asarray(X, dtype=np.float64)
    Y = X if Y is None else np.asarray(Y, dtype=np.float64)
    D = np.abs(X[:, None, :] - Y[None, :, :])
    if sum_over_features:
        return D.sum(axis=-1)
    return D.reshape(-1, X.shape[1])


def load_reference():
    """Return the imported reference package `mipego` (shimmed)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "mipego" in sys.modules and getattr(sys.modules["mipego"], "_b200_shimmed", False):
        return sys.modules["mipego"]
    if not hasattr(np, "int"):
        np.int = int  # shim 1
    if "pyDOE" not in sys.modules:  # shim 2
        m = types.ModuleType("pyDOE")
        m.lhs = _lhs
        sys.modules["pyDOE"] = m
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.tri",
                 "matplotlib.cm", "mpl_toolkits", "mpl_toolkits.mplot3d"):  # shim 3
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    gpy = types.ModuleType("mipego.GaussianProcess.gpy")  # shim 4
    gpy.PytorchGaussianProcess = None
    sys.modules["mipego.GaussianProcess.gpy"] = gpy
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import mipego  # noqa
    import mipego.GaussianProcess.gpr as gpr
    gpr.manhattan_distances = _manhattan  # shim 5
    mipego._b200_shimmed = True
    return mipego
