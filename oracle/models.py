"""Oracle restatement of the GPyOpt model wrappers and multi_outputGP.  Test infrastructure only.

Follows GPyOpt/models/gpmodel.py:136-175,259-271 (GPModel: one GPRegression
instance per hyper-sample, selected with set_hyperparameters(i); variance
clipped at 1e-10), GPyOpt/models/gpmodel_fixed_hyps.py:42-112,187-199 and
multi_outputGP.py:97-306.

Hyper-samples are either supplied explicitly (from_hyper_samples) or inferred like
GPModel.updateModel does (ML-II + HMC, gpmodel.py:102-128: GPModelInferred on top of oracle/hmc.py).
"""
import numpy as np

from .kern import Kern
from .gp import GPRegression


class GPModel(object):
    """Single-output model holding H hyper-sample instances (gpmodel.py:31-44)."""
    analytical_gradient_prediction = True

    def __init__(self, kernels, noise_vars):
        # kernels: list of H oracle Kern; noise_vars: list of H floats
        self.kernels = list(kernels)
        self.noise_vars = [float(v) for v in noise_vars]
        self.n_samples = len(self.kernels)
        self.model_instances = [None] * self.n_samples
        self.model = None
        self.current_model = None

    def updateModel(self, X_all, Y_all, X_new=None, Y_new=None):
        # gpmodel.py:102-128 minus the ML-II/HMC fit: every instance gets (X, Y) and is re-factorised
        for i in range(self.n_samples):
            if self.model_instances[i] is None:
                self.model_instances[i] = GPRegression(X_all, Y_all, self.kernels[i], self.noise_vars[i])
            else:
                self.model_instances[i].set_XY(X_all, Y_all)
        self.model = self.model_instances[0]
        self.set_hyperparameters(0)

    def number_of_hyps_samples(self):
        return self.n_samples

    def set_hyperparameters(self, i):
        # gpmodel.py:136-137
        self.current_model = self.model_instances[i]

    def predict(self, X, full_cov=False):
        # gpmodel.py:140-148
        if X.ndim == 1:
            X = X[None, :]
        m, v = self.current_model.predict(X)
        v = np.clip(v, 1e-10, np.inf)
        return m, v

    def posterior_mean(self, X):
        # gpmodel.py:162-167
        if X.ndim == 1:
            X = X[None, :]
        return self.current_model.posterior_mean(X)

    def posterior_variance(self, X):
        # gpmodel.py:169-175
        if X.ndim == 1:
            X = X[None, :]
        return np.clip(self.current_model.posterior_variance(X), 1e-10, np.inf)

    def posterior_variance_noiseless(self, X):
        # gpmodel.py:178-184
        if X.ndim == 1:
            X = X[None, :]
        return np.clip(self.current_model.posterior_variance_noiseless(X), 1e-10, np.inf)

    def posterior_mean_gradient(self, X):
        # gpmodel.py:259-264
        return self.current_model.posterior_mean_gradient(X)

    def posterior_variance_gradient(self, X):
        # gpmodel.py:266-271
        return self.current_model.posterior_variance_gradient(X)

    # knowledge-gradient helpers, gpmodel.py:187-250,273-287: plain dispatch to the current hyper-sample instance
    def partial_precomputation_for_covariance(self, X):
        self.current_model.partial_precomputation_for_covariance(X)

    def partial_precomputation_for_covariance_gradient(self, x):
        self.current_model.partial_precomputation_for_covariance_gradient(x)

    def partial_precomputation_for_variance_conditioned_on_next_point(self, next_point):
        self.current_model.partial_precomputation_for_variance_conditioned_on_next_point(next_point)

    def posterior_variance_conditioned_on_next_point(self, X):
        return self.current_model.posterior_variance_conditioned_on_next_point(X)

    def posterior_variance_gradient_conditioned_on_next_point(self, X):
        return self.current_model.posterior_variance_gradient_conditioned_on_next_point(X)

    def posterior_covariance_between_points(self, X1, X2):
        return self.current_model.posterior_covariance_between_points(X1, X2)

    def posterior_covariance_between_points_partially_precomputed(self, X1, X2):
        return self.current_model.posterior_covariance_between_points_partially_precomputed(X1, X2)

    def posterior_covariance_gradient(self, X, X2):
        return self.current_model.posterior_covariance_gradient(X, X2)

    def posterior_covariance_gradient_partially_precomputed(self, X, x2):
        return self.current_model.posterior_covariance_gradient_partially_precomputed(X, x2)


class GPModelInferred(GPModel):
    """GPModel with its own hyper-parameter inference (gpmodel.py:50-128): every updateModel runs ML-II + HMC on `model`
    and copies the sub-sampled chain states into the n_samples instances (un-fixed entries only, :121-126)."""

    def __init__(self, kind='se', kernel=None, noise_var=None, exact_feval=False, n_samples=10, ARD=False, **sampler):
        from .hmc import GPModelHMC
        self.inference = GPModelHMC(kind=kind, kernel=kernel, noise_var=noise_var, exact_feval=exact_feval,
                                    n_samples=n_samples, ARD=ARD, **sampler)
        self.kind = kind
        self.n_samples = n_samples
        self.model_instances = [None] * n_samples
        self.model = None
        self.current_model = None

    def updateModel(self, X_all, Y_all, X_new=None, Y_new=None):
        self.inference.updateModel(X_all, Y_all)
        d = X_all.shape[1]
        var, ls, noise = self.inference.hyper_samples(d)
        for i in range(self.n_samples):
            kern = Kern(self.kind, d, var[i], ls[i], ARD=True)
            self.model_instances[i] = GPRegression(X_all, Y_all, kern, noise[i])
        self.model = self.model_instances[0]
        self.set_hyperparameters(0)


class GPModelFixedHyps(GPModel):
    """gpmodel_fixed_hyps.py: one model, set_hyperparameters is a no-op (:76-77)."""

    def __init__(self, kernel=None, noise_var=None, input_dim=None):
        if kernel is None:
            # gpmodel_fixed_hyps.py:49-50
            kernel = Kern('se', input_dim, variance=2., lengthscale=0.3)
        noise_var = 1e-10 if noise_var is None else noise_var     # :56
        super(GPModelFixedHyps, self).__init__([kernel], [noise_var])

    def set_hyperparameters(self, i):
        self.current_model = self.model_instances[0]


class multi_outputGP(object):
    """multi_outputGP.py:9-348: m independent single-output models stacked to (m,N) / (m,N,d)."""
    analytical_gradient_prediction = True

    def __init__(self, output_dim, outputs, n_samples):
        self.output_dim = output_dim
        self.output = list(outputs)
        self.n_samples = n_samples

    @classmethod
    def from_hyper_samples(cls, kind, variance, lengthscale, noise, ARD=True, n_samples=None):
        """variance (H,m), lengthscale (H,m,d) [or (H,m,1) if not ARD], noise (H,m); kind: one kernel family or one per
        output (multi_outputGP.py:38-44 builds output j from kernel[j])."""
        variance = np.asarray(variance, dtype=float)
        lengthscale = np.asarray(lengthscale, dtype=float)
        noise = np.asarray(noise, dtype=float)
        H, m = variance.shape
        d = lengthscale.shape[2]
        outs = []
        for j in range(m):
            kj = kind if isinstance(kind, str) else kind[j]
            kerns = [Kern(kj, d, variance[h, j], lengthscale[h, j], ARD=ARD) for h in range(H)]
            outs.append(GPModel(kerns, [noise[h, j] for h in range(H)]))
        return cls(m, outs, H if n_samples is None else n_samples)

    @classmethod
    def inferred(cls, output_dim, kind='se', noise_var=None, exact_feval=None, n_samples=10, ARD=None, **sampler):
        # multi_outputGP.py:23-58 with fixed_hyps=False: one GPModel per output, each with its own ML-II + HMC
        outs = [GPModelInferred(kind=kind, noise_var=None if noise_var is None else noise_var[j],
                                exact_feval=False if exact_feval is None else exact_feval[j], n_samples=n_samples,
                                ARD=True if ARD is None else ARD[j], **sampler) for j in range(output_dim)]
        return cls(output_dim, outs, n_samples)

    @classmethod
    def fixed_hyps(cls, output_dim, input_dim, kernel=None, noise_var=None, n_samples=10):
        # multi_outputGP.py:23-58 with fixed_hyps=True; number_of_hyps_samples() still returns n_samples (q6)
        outs = [GPModelFixedHyps(kernel=None if kernel is None else kernel[j],
                                 noise_var=None if noise_var is None else noise_var[j],
                                 input_dim=input_dim) for j in range(output_dim)]
        return cls(output_dim, outs, n_samples)

    def updateModel(self, X_all, Y_all):
        # multi_outputGP.py:97-102
        for j in range(self.output_dim):
            self.output[j].updateModel(X_all, Y_all[j], None, None)

    def number_of_hyps_samples(self):
        return self.n_samples                                    # :109-110

    def set_hyperparameters(self, n):
        for j in range(self.output_dim):                          # :113-115
            self.output[j].set_hyperparameters(n)

    def get_evaluated_points(self):
        return np.copy(self.output[0].model.X)

    def predict(self, X, full_cov=False):
        # :138-149
        X = np.atleast_2d(X)
        m = np.empty((self.output_dim, X.shape[0]))
        cov = np.empty((self.output_dim, X.shape[0]))
        for j in range(self.output_dim):
            tmp1, tmp2 = self.output[j].predict(X, full_cov)
            m[j, :] = tmp1[:, 0]
            cov[j, :] = tmp2[:, 0]
        return m, cov

    def predict_noiseless(self, X, full_cov=False):
        # :151-162 -> gpmodel.py:150-159 (posterior_mean + clipped noiseless variance)
        X = np.atleast_2d(X)
        return self.posterior_mean(X), self.posterior_variance_noiseless(X)

    def posterior_mean(self, X):
        # :165-173
        m = np.empty((self.output_dim, X.shape[0]))
        for j in range(self.output_dim):
            m[j, :] = self.output[j].posterior_mean(X)[:, 0]
        return m

    def posterior_mean_at_evaluated_points(self):
        return self.posterior_mean(self.output[0].model.X)       # :176-180

    def posterior_variance(self, X):
        # :183-191
        var = np.empty((self.output_dim, X.shape[0]))
        for j in range(self.output_dim):
            var[j, :] = self.output[j].posterior_variance(X)[:, 0]
        return var

    def posterior_variance_noiseless(self, X):
        var = np.empty((self.output_dim, X.shape[0]))
        for j in range(self.output_dim):
            var[j, :] = self.output[j].posterior_variance_noiseless(X)[:, 0]
        return var

    def posterior_mean_gradient(self, X):
        # :284-294
        dmu_dX = np.empty((self.output_dim, X.shape[0], X.shape[1]))
        for j in range(self.output_dim):
            dmu_dX[j, :, :] = self.output[j].posterior_mean_gradient(X)
        return dmu_dX

    def posterior_variance_gradient(self, X):
        # :297-306
        dvar_dX = np.empty((self.output_dim, X.shape[0], X.shape[1]))
        for j in range(self.output_dim):
            dvar_dX[j, :, :] = self.output[j].posterior_variance_gradient(X)
        return dvar_dX

    # knowledge-gradient helpers, multi_outputGP.py:203-281,309-331: loop over outputs, stack
    def partial_precomputation_for_covariance(self, X):
        for j in range(self.output_dim):
            self.output[j].partial_precomputation_for_covariance(X)

    def partial_precomputation_for_covariance_gradient(self, x):
        for j in range(self.output_dim):
            self.output[j].partial_precomputation_for_covariance_gradient(x)

    def partial_precomputation_for_variance_conditioned_on_next_point(self, next_point):
        for j in range(self.output_dim):
            self.output[j].partial_precomputation_for_variance_conditioned_on_next_point(next_point)

    def posterior_variance_conditioned_on_next_point(self, X):
        var = np.empty((self.output_dim, X.shape[0]))
        for j in range(self.output_dim):
            var[j, :] = self.output[j].posterior_variance_conditioned_on_next_point(X)[:, 0]
        return var

    def posterior_variance_gradient_conditioned_on_next_point(self, X):
        dvar_dX = np.empty((self.output_dim, X.shape[0], X.shape[1]))
        for j in range(self.output_dim):
            dvar_dX[j, :, :] = self.output[j].posterior_variance_gradient_conditioned_on_next_point(X)
        return dvar_dX

    def posterior_covariance_between_points(self, X1, X2):
        cov = np.empty((self.output_dim, X1.shape[0], X2.shape[0]))
        for j in range(self.output_dim):
            cov[j, :, :] = self.output[j].posterior_covariance_between_points(X1, X2)
        return cov

    def posterior_covariance_between_points_partially_precomputed(self, X1, X2):
        cov = np.empty((self.output_dim, X1.shape[0], X2.shape[0]))
        for j in range(self.output_dim):
            cov[j, :, :] = self.output[j].posterior_covariance_between_points_partially_precomputed(X1, X2)
        return cov

    def posterior_covariance_gradient(self, X, x2):
        dK_dX = np.empty((self.output_dim, X.shape[0], X.shape[1]))
        for j in range(self.output_dim):
            dK_dX[j, :, :] = self.output[j].posterior_covariance_gradient(X, x2)[:, 0, :]      # multi_outputGP.py:317
        return dK_dX

    def posterior_covariance_gradient_partially_precomputed(self, X, x2):
        dK_dX = np.empty((self.output_dim, X.shape[0], X.shape[1]))
        for j in range(self.output_dim):
            dK_dX[j, :, :] = self.output[j].posterior_covariance_gradient_partially_precomputed(X, x2)[:, 0, :]
        return dK_dX
