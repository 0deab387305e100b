"""Public names of mipego_b200 -- B200-native Gaussian-process surrogate hot path behind the MIP-EGO API.

Drop-in replacements for `mipego.GaussianProcess.GaussianProcess` and
`mipego.InfillCriteria.{EI, PI, EpsilonPI, UCB, MGFI}` whose arithmetic runs in hand-written
sm_100a CUDA kernels behind the C ABI of include/mgp.h (libmgp_b200.so, bound with ctypes).
There is no CPU fallback: computing without the library or without a CUDA device raises.
"""
from . import _cabi  # noqa: F401
from .trend import constant_trend, linear_trend, BasisExpansionTrend  # noqa: F401
from .gpr import GaussianProcess  # noqa: F401
from . import InfillCriteria  # noqa: F401
from .InfillCriteria import EI, PI, EpsilonPI, UCB, MGFI  # noqa: F401
from . import optimizer  # noqa: F401
from . import dist  # noqa: F401
from .optimizer import argmax_restart, batch_argmax_restart, BoxSpace  # noqa: F401

__version__ = "0.2.0"

__all__ = ["GaussianProcess", "constant_trend", "linear_trend", "BasisExpansionTrend", "InfillCriteria", "EI", "PI",
           "EpsilonPI", "UCB", "MGFI", "optimizer", "dist", "argmax_restart", "batch_argmax_restart", "BoxSpace",
           "__version__"]
