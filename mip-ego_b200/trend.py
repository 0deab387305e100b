"""Prior-mean (trend) objects behind the reference's interface (mipego/GaussianProcess/trend.py).

The class names, `n_feature` / `n_dim` / `beta` attributes and the methods `set_beta`, `F`,
`Jacobian`, `__call__`, `check_input` are the contract: `GaussianProcess` and user scripts construct
`constant_trend(d, beta=...)` / `linear_trend(d, beta=...)` and read those members
(gpr.py:255-261, 281-288, 539).  The arithmetic of the trend itself runs on the device (the basis
matrix, F beta, the normal equations of Ft); these host objects only describe WHICH basis is meant
and carry the coefficients.  `constant_trend`: f(x) = 1, p = 1 (trend.py:69-88); `linear_trend`:
f(x) = [1, x_1 .. x_d], p = d + 1 (trend.py:90-110).  beta given = simple kriging, beta None =
estimated (ordinary / universal kriging).
"""
import numpy as np


class Trend(object):
    pass


class BasisExpansionTrend(Trend):
    """mean(x) = f(x)^T beta over a fixed basis f of `n_dim` functions of `n_feature` inputs."""

    _basis = None      # "constant" | "linear": what the device code is told to build

    def __init__(self, n_feature, n_dim=None, beta=None):
        self.n_feature = int(n_feature)
        self.n_dim = n_dim
        self.beta = None
        if n_dim is not None:
            self.set_beta(beta)

    def set_beta(self, beta):
        """Coefficients as an (n_dim, 1) column; a scalar stands for the same value on every basis
        function; None means "to be estimated" (same conventions as trend.py:27-34)."""
        if beta is None:
            self.beta = None
            return
        col = np.asarray(beta, dtype=np.float64)
        col = np.full(self.n_dim, float(col)) if col.ndim == 0 else col.ravel()
        if col.size != self.n_dim:
            raise Exception("Shapes of beta and F do not match.")
        self.beta = col.reshape(self.n_dim, 1)

    def check_input(self, X):
        """(m, n_feature) view of X; a (n_feature, m) array is transposed, as the reference does."""
        X = np.asarray(X)
        X = X.reshape(1, -1) if X.ndim < 2 else X
        for cand in (X, X.T):
            if cand.shape[1] == self.n_feature:
                return cand
        raise Exception("x does not have the right size!")

    def __call__(self, X):
        if self.beta is None:
            raise Exception("beta is not set!")
        return self.F(X) @ self.beta


class constant_trend(BasisExpansionTrend):
    """Zero-order polynomial trend f(x) = 1 (trend.py:69-88)."""

    _basis = "constant"

    def __init__(self, n_feature, beta=None):
        super().__init__(n_feature, n_dim=1, beta=beta)

    def F(self, X):
        return np.ones((len(self.check_input(X)), 1))

    def Jacobian(self, x):
        self.check_input(x)
        return np.zeros((1, self.n_feature))


class linear_trend(BasisExpansionTrend):
    """First-order polynomial trend f(x) = [1, x_1 .. x_d], p = d + 1 (trend.py:90-110)."""

    _basis = "linear"

    def __init__(self, n_feature, beta=None):
        super().__init__(n_feature, n_dim=int(n_feature) + 1, beta=beta)

    def F(self, X):
        X = self.check_input(X)
        out = np.empty((X.shape[0], self.n_dim))
        out[:, 0] = 1.0
        out[:, 1:] = X
        return out

    def Jacobian(self, x):
        self.check_input(x)
        J = np.zeros((self.n_dim, self.n_feature))
        J[1:] = np.eye(self.n_feature)
        return J
