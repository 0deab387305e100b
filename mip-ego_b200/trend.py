"""Prior-mean (trend) objects with the reference's interface (mipego/GaussianProcess/trend.py).

`constant_trend` (trend.py:69-88) and `linear_trend` (trend.py:90-110) are evaluated by the CUDA
path, with given coefficients (simple kriging) or estimated ones (ordinary / universal kriging).
"""
import numpy as np


class Trend(object):
    pass


class BasisExpansionTrend(Trend):
    def __init__(self, n_feature):
        self.n_feature = n_feature
        self.n_dim = None
        self.beta = None

    def set_beta(self, beta):
        # trend.py:27-34
        if beta is not None:
            if not hasattr(beta, "__iter__"):
                beta = np.array([beta] * self.n_dim)
            beta = np.atleast_2d(beta).reshape(-1, 1)
            if len(beta) != self.n_dim:
                raise Exception("Shapes of beta and F do not match.")
        self.beta = beta

    def __call__(self, X):
        if self.beta is None:
            raise Exception("beta is not set!")
        return self.F(X).dot(self.beta)

    def check_input(self, X):
        X = np.atleast_2d(X)
        if X.shape[1] != self.n_feature:
            X = X.T
        if X.shape[1] != self.n_feature:
            raise Exception("x does not have the right size!")
        return X


class constant_trend(BasisExpansionTrend):
    """Zero-order polynomial trend f(x) = 1 (trend.py:69-88)."""

    def __init__(self, n_feature, beta=None):
        super(constant_trend, self).__init__(n_feature)
        self.n_dim = 1
        self.set_beta(beta)

    def F(self, X):
        X = self.check_input(X)
        return np.ones((X.shape[0], 1))

    def Jacobian(self, x):
        self.check_input(x)
        return np.zeros((1, self.n_feature))


class linear_trend(BasisExpansionTrend):
    """First-order polynomial trend f(x) = [1, x_1..x_d], p = d + 1 (trend.py:90-110)."""

    def __init__(self, n_feature, beta=None):
        super(linear_trend, self).__init__(n_feature)
        self.n_dim = n_feature + 1
        self.set_beta(beta)

    def F(self, X):
        X = self.check_input(X)
        return np.c_[np.ones(X.shape[0]), X]

    def Jacobian(self, x):
        x = self.check_input(x)
        return np.r_[np.zeros((1, self.n_feature)), np.eye(self.n_feature)]
