"""Oracle restatement of the exact-GP numerics on the path.  Test infrastructure only.

Follows GPy/models/gp_regression.py:29-36 (GPRegression: Gaussian likelihood,
normalizer=True), GPy/util/normalizer.py:57-72 (fork: mean-centring only, std==1),
GPy/inference/latent_function_inference/exact_gaussian_inference.py:29-65,
GPy/inference/latent_function_inference/posterior.py:171-191,268-320 and the
author-added GPy/core/gp.py:286-326,380-490 and the knowledge-gradient helpers gp.py:493-627.
"""
import numpy as np

from .linalg import pdinv, dpotrs, dpotri, dtrtrs, symmetrify


class GPRegression(object):
    def __init__(self, X, Y, kernel, noise_var=1.):
        self.kern = kernel
        self.noise_var = float(noise_var)       # Gaussian likelihood variance (gaussian.py:61)
        self.set_XY(X, Y)

    # gp.py:191-227 + normalizer.py:57-66
    def set_XY(self, X, Y):
        self.X = np.array(X, dtype=float)
        self.Y = np.array(Y, dtype=float).reshape(self.X.shape[0], -1)
        self.y_mean = self.Y.mean(0)            # Standardize.scale_by; std = 1
        self.Y_normalized = (self.Y - self.y_mean) / 1
        self.parameters_changed()

    # gp.py:247-260 -> exact_gaussian_inference.py:29-65
    def parameters_changed(self):
        K = self.kern.K(self.X)
        Ky = K.copy()
        Ky[np.diag_indices_from(Ky)] += self.noise_var + 1e-8       # :46-47
        Wi, LW, LWi, W_logdet = pdinv(Ky)                             # :49
        alpha, _ = dpotrs(LW, self.Y_normalized, lower=1)             # :51
        self.woodbury_chol = LW
        self.woodbury_vector = alpha
        self._woodbury_inv = None
        self.K_train = K
        # exact_gaussian_inference.py:53-63 (single-output Y: Y.size = n, Y.shape[1] = 1)
        n = self.Y_normalized.shape[0]
        self._log_marginal_likelihood = 0.5 * (-n * np.log(2. * np.pi) - self.Y_normalized.shape[1] * W_logdet
                                               - np.sum(alpha * self.Y_normalized))
        self._dL_dK = 0.5 * (np.dot(alpha, alpha.T) - self.Y_normalized.shape[1] * Wi)

    # gp.py:262-266
    def log_likelihood(self):
        return float(self._log_marginal_likelihood)

    def likelihood_gradients(self):
        """d log p(y) / d (kernel variance, lengthscales, Gaussian noise variance): what GP.parameters_changed leaves
        in kern.variance.gradient, kern.lengthscale.gradient and likelihood.variance.gradient (gp.py:256-258,
        gaussian.py:64-71: the noise gradient is the trace of dL_dK)."""
        g_var, g_len = self.kern.update_gradients_full(self._dL_dK, self.X)
        g_noise = float(np.diag(self._dL_dK).sum())
        return g_var, g_len, g_noise

    @property
    def woodbury_inv(self):
        # posterior.py:171-184: dpotri(L) then symmetrify (dpotri wrapper already symmetrifies)
        if self._woodbury_inv is None:
            self._woodbury_inv, _ = dpotri(self.woodbury_chol, lower=1)
            symmetrify(self._woodbury_inv)
        return self._woodbury_inv

    # posterior.py:299-305 + gp.py:380-400
    def posterior_mean(self, Xnew):
        Kx = self.kern.K(Xnew, self.X)
        mu = np.dot(Kx, self.woodbury_vector)
        if len(mu.shape) == 1:
            mu = mu.reshape(-1, 1)
        return (mu * 1) + self.y_mean            # normalizer.inverse_mean

    # posterior.py:308-320
    def _raw_posterior_variance(self, Xnew):
        Kx = self.kern.K(self.X, Xnew)
        Kxx = self.kern.Kdiag(Xnew)
        tmp = dtrtrs(self.woodbury_chol, Kx)[0]
        var = (Kxx - np.square(tmp).sum(0))[:, None]
        return var

    # gp.py:403-418 + gaussian.py:110-111
    def posterior_variance(self, Xnew):
        return self.noise_var + self._raw_posterior_variance(Xnew)

    # gp.py:421-435
    def posterior_variance_noiseless(self, Xnew):
        return self._raw_posterior_variance(Xnew)

    # gp.py:286-326 -> posterior.py:270-296 -> gaussian.py:94-102
    def predict(self, Xnew):
        Kx = self.kern.K(self.X, Xnew)
        mu = np.dot(Kx.T, self.woodbury_vector)
        if len(mu.shape) == 1:
            mu = mu.reshape(-1, 1)
        Kxx = self.kern.Kdiag(Xnew)
        tmp = dtrtrs(self.woodbury_chol, Kx)[0]
        var = (Kxx - np.square(tmp).sum(0))[:, None]
        var = var + self.noise_var
        mu = (mu * 1) + self.y_mean
        return mu, var

    # gp.py:438-461
    def posterior_mean_gradient(self, X):
        return self.kern.gradients_X(self.woodbury_vector.T, X, self.X)

    # gp.py:464-490.  The first term gradients_X(eye(N), X) is identically zero for
    # these stationary kernels (diagonal distance is exactly 0) and allocates N x N
    # in the reference (quirk q7); it is evaluated only for small N to stay literal.
    def posterior_variance_gradient(self, X):
        if X.shape[0] <= 512:
            dv_dX = self.kern.gradients_X(np.eye(X.shape[0]), X)
        else:
            dv_dX = np.zeros(X.shape)
        alpha = -2. * np.dot(self.kern.K(X, self.X), self.woodbury_inv)
        dv_dX = dv_dX + self.kern.gradients_X(alpha, X, self.X)
        return dv_dX

    # ---- knowledge-gradient helpers (author-added, gp.py:493-627) ----------------------------------------------
    # gp.py:493-501
    def partial_precomputation_for_covariance(self, X):
        self.partial_precomp_cov = np.matmul(self.woodbury_inv, self.kern.K(self.X, X))

    # gp.py:504-512
    def partial_precomputation_for_covariance_gradient(self, x):
        self.partial_precomp_dcov = np.matmul(self.kern.K(x, self.X), self.woodbury_inv)

    # gp.py:515-530
    def partial_precomputation_for_variance_conditioned_on_next_point(self, next_point):
        self.X_next = np.append(self.X, next_point, axis=0)
        K_aux = self.kern.K(self.X_next)
        K_aux[np.diag_indices_from(K_aux)] += self.noise_var + 1e-8            # diag.add(K_aux, noise_var + 1e-8)
        tmp = pdinv(K_aux)
        self.woodbury_inv_conditioned_on_next_point = tmp[0]
        self.woodbury_chol_conditioned_on_next_point = tmp[1]

    # gp.py:533-544
    def posterior_variance_conditioned_on_next_point(self, X):
        Kx = self.kern.K(self.X_next, X)
        Kxx = self.kern.Kdiag(X)
        tmp = dtrtrs(self.woodbury_chol_conditioned_on_next_point, Kx)[0]
        return (Kxx - np.square(tmp).sum(0))[:, None]

    # gp.py:547-575
    def posterior_variance_gradient_conditioned_on_next_point(self, X):
        dv_dX = self.kern.gradients_X(np.eye(X.shape[0]), X)
        alpha = -2. * np.dot(self.kern.K(X, self.X_next), self.woodbury_inv_conditioned_on_next_point)
        dv_dX = dv_dX + self.kern.gradients_X(alpha, X, self.X_next)
        return dv_dX

    # gp.py:577-585 -> posterior.py covariance_between_points: K(X1, X2) - (L^-1 K(X, X1))^T (L^-1 K(X, X2))
    def posterior_covariance_between_points(self, X1, X2):
        tmp1 = dtrtrs(self.woodbury_chol, self.kern.K(self.X, X1))[0]
        tmp2 = dtrtrs(self.woodbury_chol, self.kern.K(self.X, X2))[0]
        return self.kern.K(X1, X2) - tmp1.T.dot(tmp2)

    # gp.py:588-599
    def posterior_covariance_between_points_partially_precomputed(self, X1, X2):
        return self.kern.K(X1, X2) - np.matmul(self.kern.K(self.X, X1).T, self.partial_precomp_cov)

    # gp.py:601-609 (kern.gradients_X(None, ...) returns the per-pair tensor (N, M, d): se.py:142-144)
    def posterior_covariance_gradient(self, X, x2):
        factor = np.matmul(self.kern.K(x2, self.X), self.woodbury_inv)
        return self.kern.gradients_X(None, X, x2) - np.matmul(factor, self.kern.gradients_X(None, X, self.X))

    # gp.py:612-627
    def posterior_covariance_gradient_partially_precomputed(self, X, x2):
        return self.kern.gradients_X(None, X, x2) - np.matmul(self.partial_precomp_dcov, self.kern.gradients_X(None, X, self.X))
