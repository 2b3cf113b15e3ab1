"""Kernel specifications with GPy's constructor signatures (no arithmetic here: the covariance is
evaluated inside the CUDA kernels, bocf_b200/csrc/kernfn.cuh).

GPy.kern.SE        GPy/kern/src/se.py:14-35      (the fork's default kernel, gpmodel.py:58, gpmodel_fixed_hyps.py:50)
GPy.kern.RBF       GPy/kern/src/rbf.py:21
GPy.kern.Matern52  GPy/kern/src/stationary.py:513
GPy.kern.Matern32  GPy/kern/src/stationary.py:424
"""
import numpy as np


class Kern(object):
    kind = None

    def __init__(self, input_dim, variance=1., lengthscale=None, ARD=False, active_dims=None, name=None):
        self.input_dim = int(input_dim)
        self.ARD = ARD
        if not ARD:
            if lengthscale is None:
                lengthscale = np.ones(1)
            else:
                lengthscale = np.asarray(lengthscale, dtype=float).reshape(-1)
                assert lengthscale.size == 1, "Only 1 lengthscale needed for non-ARD kernel"
        else:
            if lengthscale is not None:
                lengthscale = np.asarray(lengthscale, dtype=float).reshape(-1)
                assert lengthscale.size in [1, input_dim], "Bad number of lengthscales"
                if lengthscale.size != input_dim:
                    lengthscale = np.ones(input_dim) * lengthscale
            else:
                lengthscale = np.ones(self.input_dim)
        self.lengthscale = np.array(lengthscale, dtype=float)
        self.variance = np.array([float(np.asarray(variance).reshape(-1)[0])])

    def lengthscale_vector(self):
        """Per-dimension lengthscales (a non-ARD kernel broadcasts its single lengthscale)."""
        if self.lengthscale.size == self.input_dim:
            return self.lengthscale.copy()
        return np.full(self.input_dim, float(self.lengthscale[0]))


class SE(Kern):
    kind = "se"


class RBF(Kern):
    kind = "rbf"


class Matern52(Kern):
    kind = "matern52"


class Matern32(Kern):
    kind = "matern32"


BY_KIND = {"se": SE, "rbf": RBF, "matern52": Matern52, "matern32": Matern32}
