"""Helpers shared by the oracle (CPU) and CUDA (GPU) parity tests."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

FIXED = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "fixed_*.npz")))
VARS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "var_*.npz")))
LINS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "lin_*.npz")))
FITS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "fit_*.npz")))
ISOS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "iso_*.npz")))
MFITS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "mfit_*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    for k in ("kind", "likelihood", "mode", "trend"):
        if k in g:
            g[k] = str(g[k])
    for k in ("estimate_trend",):
        g[k] = bool(g[k])
    return g


def mode_of(g):
    if "mode" in g:
        return g["mode"]
    return "noisy" if float(g["nugget"]) > 0 else "noiseless"


def beta_of(g):
    if g["estimate_trend"]:
        return 0.0
    b = np.asarray(g["beta_in"], dtype=np.float64)
    return float(b) if b.ndim == 0 else b


def trend_of(g):
    return g.get("trend", "constant")


def relerr(a, b, floor=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / (np.abs(b) + floor)) if a.size else 0.0


def likelihood_of(g):
    return g.get("likelihood", "concentrated")
