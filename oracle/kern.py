"""Oracle restatement of the covariance kernels on the path.  Test infrastructure only.

Follows GPy/kern/src/stationary.py (Stationary, Matern32, Matern52),
GPy/kern/src/rbf.py (RBF), GPy/kern/src/se.py (SE, the fork's default kernel)
and the native GPy/kern/src/stationary_utils.c (_grad_X).
"""
import numpy as np
import scipy.spatial.distance

from .linalg import tdot

KINDS = ('se', 'rbf', 'matern52', 'matern32')


class Kern(object):
    """A stationary kernel with variance sigma_f^2 and lengthscale(s).

    kind in {'se','rbf','matern52','matern32'}; 'se' is GPy.kern.SE (exact
    differences via cdist), the others subclass GPy's Stationary (expanded
    ||x||^2+||x'||^2-2xx' distance with clipping).
    """

    def __init__(self, kind, input_dim, variance=1., lengthscale=None, ARD=False):
        assert kind in KINDS
        self.kind = kind
        self.input_dim = int(input_dim)
        self.ARD = ARD
        # stationary.py:58-73 / se.py:15-35
        if not ARD:
            if lengthscale is None:
                lengthscale = np.ones(1)
            else:
                lengthscale = np.asarray(lengthscale, dtype=float).reshape(-1)
                assert lengthscale.size == 1
        else:
            if lengthscale is not None:
                lengthscale = np.asarray(lengthscale, dtype=float).reshape(-1)
                assert lengthscale.size in [1, input_dim]
                if lengthscale.size != input_dim:
                    lengthscale = np.ones(input_dim) * lengthscale
            else:
                lengthscale = np.ones(self.input_dim)
        self.lengthscale = np.array(lengthscale, dtype=float)
        self.variance = np.array([float(np.asarray(variance).reshape(-1)[0])])

    # ---- Stationary family -------------------------------------------------
    def _unscaled_dist(self, X, X2=None):
        # stationary.py:128-146
        if X2 is None:
            Xsq = np.sum(np.square(X), 1)
            r2 = -2. * tdot(X) + (Xsq[:, None] + Xsq[None, :])
            r2[np.diag_indices_from(r2)] = 0.
            r2 = np.clip(r2, 0, np.inf)
            return np.sqrt(r2)
        else:
            X1sq = np.sum(np.square(X), 1)
            X2sq = np.sum(np.square(X2), 1)
            r2 = -2. * np.dot(X, X2.T) + (X1sq[:, None] + X2sq[None, :])
            r2 = np.clip(r2, 0, np.inf)
            return np.sqrt(r2)

    def _scaled_dist(self, X, X2=None):
        # stationary.py:148-166
        if self.ARD:
            if X2 is not None:
                X2 = X2 / self.lengthscale
            return self._unscaled_dist(X / self.lengthscale, X2)
        else:
            return self._unscaled_dist(X, X2) / self.lengthscale

    def K_of_r(self, r):
        if self.kind == 'rbf':        # rbf.py:42-43
            return self.variance * np.exp(-0.5 * r**2)
        if self.kind == 'matern52':   # stationary.py:529-530
            return self.variance * (1 + np.sqrt(5.) * r + 5. / 3 * r**2) * np.exp(-np.sqrt(5.) * r)
        if self.kind == 'matern32':   # stationary.py:440-441
            return self.variance * (1. + np.sqrt(3.) * r) * np.exp(-np.sqrt(3.) * r)
        raise ValueError(self.kind)

    def dK_dr(self, r):
        if self.kind == 'rbf':        # rbf.py:45-46
            return -r * self.K_of_r(r)
        if self.kind == 'matern52':   # stationary.py:532-533
            return self.variance * (10. / 3 * r - 5. * r - 5. * np.sqrt(5.) / 3 * r**2) * np.exp(-np.sqrt(5.) * r)
        if self.kind == 'matern32':   # stationary.py:443-444
            return -3. * self.variance * r * np.exp(-np.sqrt(3.) * r)
        raise ValueError(self.kind)

    def _inv_dist(self, X, X2=None):
        # stationary.py:227-234
        dist = self._scaled_dist(X, X2).copy()
        return 1. / np.where(dist != 0., dist, np.inf)

    # ---- SE ----------------------------------------------------------------
    def _se_scaled_squared_dist(self, X, X2=None):
        # se.py:63-92
        if self.ARD:
            if X2 is not None:
                X2 = X2 / self.lengthscale
            Xs = X / self.lengthscale
            if X2 is None:
                return scipy.spatial.distance.pdist(Xs, 'sqeuclidean')
            return scipy.spatial.distance.cdist(Xs, X2, 'sqeuclidean')
        else:
            if X2 is None:
                return scipy.spatial.distance.pdist(X, 'sqeuclidean') / (self.lengthscale**2)
            return scipy.spatial.distance.cdist(X, X2, 'sqeuclidean') / (self.lengthscale**2)

    def _se_scaled_squared_norm(self, D):
        # se.py:94-100
        if self.ARD:
            return np.sum(np.square(D / self.lengthscale), axis=2)
        return np.sum(np.square(D), axis=2) / (self.lengthscale**2)

    # ---- public ------------------------------------------------------------
    def K(self, X, X2=None):
        if self.kind == 'se':
            # se.py:44-62
            if X2 is None:
                val = scipy.spatial.distance.squareform(
                    self.variance * np.exp(-0.5 * self._se_scaled_squared_dist(X)), checks=False)
                np.fill_diagonal(val, self.variance)
            else:
                val = self.variance * np.exp(-0.5 * self._se_scaled_squared_dist(X, X2))
            return val
        # stationary.py:104-113
        r = self._scaled_dist(X, X2)
        return self.K_of_r(r)

    # ---- hyper-parameter gradients --------------------
    def update_gradients_full(self, dL_dK, X):
        """(d/d variance, d/d lengthscale) of an objective with derivative dL_dK wrt K(X, X).

        Stationary kinds: stationary.py:191-215 with the native loop of stationary_utils.c:34-48
        (_lengthscale_grads); SE: se.py:169-185 (X2 is None branch).  Non-ARD: one shared lengthscale
        (stationary.py:213-215, se.py:184-185)."""
        if self.kind == 'se':
            squared_dist = self._se_scaled_squared_dist(X)
            exp_squared_dist = np.exp(-0.5 * squared_dist)
            tmp = scipy.spatial.distance.squareform(exp_squared_dist, checks=False)
            np.fill_diagonal(tmp, 1.)
            g_var = np.sum(tmp * dL_dK)                                                  # se.py:181
            if not self.ARD:                                                             # se.py:185
                g_len = (self.variance / self.lengthscale) * np.sum(
                    scipy.spatial.distance.squareform(exp_squared_dist * squared_dist) * dL_dK)
                return float(g_var), np.asarray(g_len, dtype=float).reshape(-1)
            g_len = (self.variance * np.sum((tmp * dL_dK)[:, :, None] *
                                            np.square(X[:, None, :] - X[None, :, :]), axis=(0, 1))) / (self.lengthscale**3)
            return float(g_var), np.asarray(g_len, dtype=float)
        g_var = np.sum(self.K(X) * dL_dK) / self.variance                               # stationary.py:197
        r = self._scaled_dist(X)
        dL_dr = self.dK_dr(r) * dL_dK                                                    # :203 (dK_dr_via_X)
        if not self.ARD:                                                                 # :213-215
            return float(np.asarray(g_var).reshape(-1)[0]), np.asarray(-np.sum(dL_dr * r) / self.lengthscale).reshape(-1)
        tmp = dL_dr * self._inv_dist(X)                                                  # :206
        grads = np.array([np.sum(tmp * np.square(X[:, q:q + 1] - X[:, q:q + 1].T)) for q in range(self.input_dim)])
        return float(np.asarray(g_var).reshape(-1)[0]), -grads / self.lengthscale**3    # :232-240

    def Kdiag(self, X):
        # stationary.py:168-171 / se.py:102-111
        ret = np.empty(X.shape[0])
        ret[:] = self.variance
        return ret

    def gradients_X(self, dL_dK, X, X2=None):
        """d/dX of sum(dL_dK * K(X, X2)).  dL_dK may be (1,n) and broadcasts (gp.py:446)."""
        if self.kind == 'se':
            # se.py:135-148
            if X2 is None:
                X2 = X
            aux1 = X[:, None, :] - X2[None, :, :]
            aux2 = (-self.variance) / (self.lengthscale**2)
            if dL_dK is None:                       # se.py:142-144: the per-pair gradient tensor (N, M, d)
                aux3 = np.exp((-0.5) * self._se_scaled_squared_norm(aux1))
                return (aux3[:, :, None] * aux1) * aux2
            aux3 = np.exp((-0.5) * self._se_scaled_squared_norm(aux1)) * dL_dK
            grad = np.sum(aux3[:, :, None] * aux1, axis=1) * aux2
            return grad
        # stationary.py:332-342 (_gradients_X_cython) == :312-330 (_gradients_X_pure)
        invdist = self._inv_dist(X, X2)
        dL_dr = self.dK_dr(self._scaled_dist(X, X2)) * dL_dK
        tmp = invdist * dL_dr
        if X2 is None:
            tmp = tmp + tmp.T
            X2 = X
        grad = grad_X(np.ascontiguousarray(X), np.ascontiguousarray(X2), np.ascontiguousarray(tmp))
        return grad / self.lengthscale**2

    def gradients_X_diag(self, dL_dKdiag, X):
        # stationary.py:344-345 / se.py:150-151
        return np.zeros(X.shape)


def grad_X(X, X2, tmp):
    """stationary_utils.c:1-14  grad[n,d] = sum_m tmp[n,m] * (X[n,d] - X2[m,d]).

    Same summation order over m as the C loop for each (n,d) (sequential in m).
    """
    N, D = X.shape
    M = X2.shape[0]
    grad = np.zeros((N, D))
    # chunk over n to bound the (n, M, D) temporary; cumulative order over m kept by np.add.reduce? -> use
    # an explicit sequential-in-m accumulation identical to the C loop.
    for m in range(M):
        grad += tmp[:, m:m + 1] * (X - X2[m][None, :])
    return grad
