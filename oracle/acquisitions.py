"""Oracle restatement of the acquisition classes on the path.  Test infrastructure only.

Follows GPyOpt/acquisitions/base.py:6-74 (AcquisitionBase), uEI_noiseless.py
(EI-CF), uPI.py, maEI.py, maPI.py, EI.py, PI.py.  The ``*_literal`` methods keep
the reference's loop structure (Python triple loop) exactly; the ``vectorised``
twins compute the same sums with numpy broadcasting (used for larger test sizes
and as the strong CPU baseline in bench.py).

Differences from the reference, all explicit so both sides of a parity test
see identical inputs: base samples Z (W_samples) and theta samples are passed
in / settable instead of being drawn from the unseeded global RNG; the pathos
process pool (uEI_noiseless.py:85-97) is replaced by its sequential equivalent
(the pool path computes the same numbers, quirk q1).
"""
import numpy as np
from scipy.special import erfc
from scipy.stats import norm


class AcquisitionBase(object):
    # GPyOpt/acquisitions/base.py:6-74
    analytical_gradient_prediction = False

    def __init__(self, model, space, optimizer, cost_withGradients=None):
        self.model = model
        self.space = space
        self.optimizer = optimizer
        self.analytical_gradient_acq = self.analytical_gradient_prediction and self.model.analytical_gradient_prediction

    def acquisition_function(self, x):
        return -self._compute_acq(x)                      # base.py:33-40

    def acquisition_function_withGradients(self, x):
        f_acqu, df_acqu = self._compute_acq_withGradients(x)
        return -f_acqu, -df_acqu                          # base.py:43-56

    def optimize(self, duplicate_manager=None, x_baseline=None):
        # base.py:58-66
        if not self.analytical_gradient_acq:
            return self.optimizer.optimize(f=self.acquisition_function, duplicate_manager=duplicate_manager,
                                           x_baseline=x_baseline)
        return self.optimizer.optimize(f=self.acquisition_function, f_df=self.acquisition_function_withGradients,
                                       duplicate_manager=duplicate_manager, x_baseline=x_baseline)


class uEI_noiseless(AcquisitionBase):
    """EI-CF.  uEI_noiseless.py:9-175."""
    analytical_gradient_prediction = True

    def __init__(self, model, space=None, optimizer=None, cost_withGradients=None, utility=None,
                 W_samples=None, utility_params_samples=None, grad_utility_params_samples=None, vectorised=False):
        self.utility = utility
        super(uEI_noiseless, self).__init__(model, space, optimizer, cost_withGradients)
        self.n_attributes = self.model.output_dim
        # uEI_noiseless.py:31 draws (25, m) normals from the global RNG; explicit here
        self.W_samples = np.random.normal(size=(25, self.n_attributes)) if W_samples is None else np.asarray(W_samples)
        self.n_hyps_samples = min(10, self.model.number_of_hyps_samples())        # :32
        self.use_full_support = self.utility.parameter_dist.use_full_support      # :33
        if self.use_full_support:
            self.utility_params_samples = self.utility.parameter_dist.support     # :35-36
            self.utility_prob_dist = np.atleast_1d(self.utility.parameter_dist.prob_dist)
        else:
            self.utility_params_samples = (self.utility.parameter_dist.sample(10) if utility_params_samples is None
                                           else utility_params_samples)           # :38
        # :126 draws ONE fresh theta per gradient call (quirk q4); explicit override for parity tests
        self.grad_utility_params_samples = grad_utility_params_samples
        self.vectorised = vectorised

    # ---- value ------------------------------------------------------------------------------------
    def _compute_acq(self, X, parallel=True):
        # uEI_noiseless.py:40-61.  NOTE the two branches are NOT equivalent when H > 1: the pool branch (:44,
        # :85-116) re-evaluates f* inside the helper, i.e. with the hyper-sample of the current pass, while the
        # sequential branch (:46, :63-83) evaluates it once before the h loop (stale hyper-sample, quirk q2).
        X = np.atleast_2d(X)
        if parallel and len(X) > 1:
            if self.vectorised:
                marginal_acqX = self._marginal_acq_vec(X, self.utility_params_samples, per_h_fstar=True)
            else:
                marginal_acqX = self._marginal_acq_parallel(X)
        elif self.vectorised:
            marginal_acqX = self._marginal_acq_vec(X, self.utility_params_samples)
        else:
            marginal_acqX = self._marginal_acq(X, self.utility_params_samples)
        if self.use_full_support:
            acqX = np.matmul(marginal_acqX, self.utility_prob_dist)
        else:
            acqX = np.sum(marginal_acqX, axis=1) / len(self.utility_params_samples)
        return np.reshape(acqX, (X.shape[0], 1))

    def _marginal_acq(self, X, utility_params_samples):
        # uEI_noiseless.py:63-83 (literal)
        fX_evaluated = self.model.posterior_mean_at_evaluated_points()
        n_w = self.W_samples.shape[0]
        L = len(utility_params_samples)
        marginal_acqX = np.zeros((X.shape[0], L))
        for h in range(self.n_hyps_samples):
            self.model.set_hyperparameters(h)
            muX = self.model.posterior_mean(X)
            sigmaX = np.sqrt(self.model.posterior_variance(X))
            for l in range(L):
                max_valX_evaluated = np.max(self.utility.eval_func(utility_params_samples[l], fX_evaluated))
                for W in self.W_samples:
                    for i in range(X.shape[0]):
                        valx = self.utility.eval_func(utility_params_samples[l], muX[:, i] + sigmaX[:, i] * W)
                        marginal_acqX[i, l] += max(valx - max_valX_evaluated, 0)
        marginal_acqX /= (self.n_hyps_samples * n_w)
        return marginal_acqX

    def _marginal_acq_parallel(self, X):
        # uEI_noiseless.py:85-97 with the pathos pool replaced by a serial map (the first, discarded map of
        # :93 is skipped -- quirk q1)
        marginal_acqX = np.zeros((X.shape[0], len(self.utility_params_samples)))
        n_w = self.W_samples.shape[0]
        for h in range(self.n_hyps_samples):
            self.model.set_hyperparameters(h)
            marginal_acqX += np.atleast_2d([self._parallel_acq_helper(x) for x in X])
        marginal_acqX /= (self.n_hyps_samples * n_w)
        return marginal_acqX

    def _improvement(self, valx, max_valX_evaluated):
        return max(valx - max_valX_evaluated, 0)                        # uEI_noiseless.py:114

    def _parallel_acq_helper(self, x):
        # uEI_noiseless.py:99-116
        x = np.atleast_2d(x)
        fX_evaluated = self.model.posterior_mean_at_evaluated_points()
        utility_params_samples = self.utility_params_samples
        L = len(utility_params_samples)
        marginal_acqx = np.zeros(L)
        mux = self.model.posterior_mean(x)[:, 0]
        sigmax = np.sqrt(self.model.posterior_variance(x))[:, 0]
        for l in range(L):
            max_valX_evaluated = np.max(self.utility.eval_func(utility_params_samples[l], fX_evaluated))
            for W in self.W_samples:
                valx = self.utility.eval_func(utility_params_samples[l], mux + sigmax * W)
                marginal_acqx[l] += self._improvement(valx, max_valX_evaluated)
        return marginal_acqx

    def _improvement_vec(self, valx, fstar):
        return np.maximum(valx - fstar, 0)

    def _marginal_acq_vec(self, X, utility_params_samples, per_h_fstar=False):
        fX_evaluated = self.model.posterior_mean_at_evaluated_points()
        n_w = self.W_samples.shape[0]
        L = len(utility_params_samples)
        marginal_acqX = np.zeros((X.shape[0], L))
        for h in range(self.n_hyps_samples):
            self.model.set_hyperparameters(h)
            if per_h_fstar:
                fX_evaluated = self.model.posterior_mean_at_evaluated_points()
            muX = self.model.posterior_mean(X)
            sigmaX = np.sqrt(self.model.posterior_variance(X))
            for l in range(L):
                fstar = np.max(self.utility.eval_func(utility_params_samples[l], fX_evaluated))
                for W in self.W_samples:
                    a = muX + sigmaX * W[:, None]                      # (m, N)
                    valx = self.utility.eval_func(utility_params_samples[l], a)
                    marginal_acqX[:, l] += self._improvement_vec(valx, fstar)
        marginal_acqX /= (self.n_hyps_samples * n_w)
        return marginal_acqX

    # ---- value + gradient ---------------------------------------------------------------------------
    def _compute_acq_withGradients(self, X):
        # uEI_noiseless.py:118-136
        X = np.atleast_2d(X)
        if self.use_full_support:
            utility_params_samples2 = self.utility.parameter_dist.support
        elif self.grad_utility_params_samples is not None:
            utility_params_samples2 = self.grad_utility_params_samples
        else:
            utility_params_samples2 = self.utility.parameter_dist.sample(1)
        if self.vectorised:
            marginal_acqX, marginal_dacq_dX = self._marginal_acq_with_gradient_vec(X, utility_params_samples2)
        else:
            marginal_acqX, marginal_dacq_dX = self._marginal_acq_with_gradient(X, utility_params_samples2)
        if self.use_full_support:
            acqX = np.matmul(marginal_acqX, self.utility_prob_dist)
            dacq_dX = np.tensordot(marginal_dacq_dX, self.utility_prob_dist, 1)
        else:
            acqX = np.sum(marginal_acqX, axis=1) / len(utility_params_samples2)
            dacq_dX = np.sum(marginal_dacq_dX, axis=2) / len(utility_params_samples2)
        acqX = np.reshape(acqX, (X.shape[0], 1))
        dacq_dX = np.reshape(dacq_dX, X.shape)
        return acqX, dacq_dX

    def _marginal_acq_with_gradient(self, X, utility_params_samples):
        # uEI_noiseless.py:138-170 (literal)
        fX_evaluated = self.model.posterior_mean_at_evaluated_points()
        X = np.atleast_2d(X)
        marginal_acqX = np.zeros((X.shape[0], len(utility_params_samples)))
        marginal_dacq_dX = np.zeros((X.shape[0], X.shape[1], len(utility_params_samples)))
        W_samples2 = self.W_samples
        n_w = W_samples2.shape[0]
        for h in range(self.n_hyps_samples):
            self.model.set_hyperparameters(h)
            muX = self.model.posterior_mean(X)
            sigmaX = np.sqrt(self.model.posterior_variance(X))
            dmuX_dX = self.model.posterior_mean_gradient(X)
            dvar_dX = self.model.posterior_variance_gradient(X)
            for l in range(len(utility_params_samples)):
                max_valX_evaluated = np.max(self.utility.eval_func(utility_params_samples[l], fX_evaluated))
                for W in W_samples2:
                    for i in range(X.shape[0]):
                        a = muX[:, i] + sigmaX[:, i] * W
                        valx = self.utility.eval_func(utility_params_samples[l], a)
                        marginal_acqX[i, l] += max(valx - max_valX_evaluated, 0)
                        if valx > max_valX_evaluated:
                            b = np.multiply((0.5 * W / sigmaX[:, i]), dvar_dX[:, i, :].transpose()).transpose()
                            b += dmuX_dX[:, i, :]
                            marginal_dacq_dX[i, :, l] += np.matmul(
                                self.utility.eval_gradient(utility_params_samples[l], a), b)
        marginal_acqX /= (self.n_hyps_samples * n_w)
        marginal_dacq_dX /= (self.n_hyps_samples * n_w)
        return marginal_acqX, marginal_dacq_dX

    def _marginal_acq_with_gradient_vec(self, X, utility_params_samples, dU_vec=None):
        fX_evaluated = self.model.posterior_mean_at_evaluated_points()
        N, d = X.shape
        L = len(utility_params_samples)
        marginal_acqX = np.zeros((N, L))
        marginal_dacq_dX = np.zeros((N, d, L))
        n_w = self.W_samples.shape[0]
        for h in range(self.n_hyps_samples):
            self.model.set_hyperparameters(h)
            muX = self.model.posterior_mean(X)
            sigmaX = np.sqrt(self.model.posterior_variance(X))
            dmuX_dX = self.model.posterior_mean_gradient(X)
            dvar_dX = self.model.posterior_variance_gradient(X)
            for l in range(L):
                theta = utility_params_samples[l]
                fstar = np.max(self.utility.eval_func(theta, fX_evaluated))
                for W in self.W_samples:
                    a = muX + sigmaX * W[:, None]                              # (m, N)
                    valx = self.utility.eval_func(theta, a)                    # (N,)
                    marginal_acqX[:, l] += np.maximum(valx - fstar, 0)
                    act = valx > fstar
                    if not np.any(act):
                        continue
                    name = getattr(self.utility, 'composite_name', None)
                    if name is not None:
                        from .utility import eval_gradient_batch
                        dU = eval_gradient_batch(name, theta, a[:, act])
                    else:
                        dU = np.stack([self.utility.eval_gradient(theta, a[:, i]) for i in np.nonzero(act)[0]], axis=1)
                    b = (0.5 * W[:, None] / sigmaX[:, act])[:, :, None] * dvar_dX[:, act, :] + dmuX_dX[:, act, :]
                    marginal_dacq_dX[act, :, l] += np.einsum('ji,jiq->iq', dU, b)
        marginal_acqX /= (self.n_hyps_samples * n_w)
        marginal_dacq_dX /= (self.n_hyps_samples * n_w)
        return marginal_acqX, marginal_dacq_dX

    def update_Z_samples(self, n_samples=None):
        # uEI_noiseless.py:172-174
        self.W_samples = np.random.normal(size=self.W_samples.shape)


class uPI(uEI_noiseless):
    """uPI.py: indicator of improvement over f* + 1e-6 (:83); no gradient (:19)."""
    analytical_gradient_prediction = False

    def __init__(self, *a, **kw):
        super(uPI, self).__init__(*a, **kw)
        self.jitter = 1e-6                                                     # uPI.py:40

    def _marginal_acq(self, X, utility_params_samples):
        # uPI.py:66-86
        fX_evaluated = self.model.posterior_mean_at_evaluated_points()
        n_w = self.W_samples.shape[0]
        L = len(utility_params_samples)
        marginal_acqX = np.zeros((X.shape[0], L))
        for h in range(self.n_hyps_samples):
            self.model.set_hyperparameters(h)
            muX = self.model.posterior_mean(X)
            sigmaX = np.sqrt(self.model.posterior_variance(X))
            for l in range(L):
                max_valX_evaluated = np.max(self.utility.eval_func(utility_params_samples[l], fX_evaluated))
                for W in self.W_samples:
                    for i in range(X.shape[0]):
                        valx = self.utility.eval_func(utility_params_samples[l], muX[:, i] + sigmaX[:, i] * W)
                        marginal_acqX[i, l] += np.where((valx - (max_valX_evaluated + self.jitter)) > 0., 1., 0.)
        marginal_acqX /= (self.n_hyps_samples * n_w)
        return marginal_acqX

    def _improvement(self, valx, max_valX_evaluated):
        # uPI.py:114 (pool helper) == :83
        return np.where((valx - (max_valX_evaluated + self.jitter)) > 0., 1., 0.)

    def _improvement_vec(self, valx, fstar):
        return np.where((valx - (fstar + self.jitter)) > 0., 1., 0.)

    def _compute_acq_withGradients(self, X):
        raise NotImplementedError('uPI has no analytical gradient (uPI.py:19)')


class maEI(AcquisitionBase):
    """maEI.py: analytic EI of the linear scalarisation theta^T y."""
    analytical_gradient_prediction = True
    jitter = 0.0
    n_theta_draw = 3                                                           # maEI.py:46,65

    def __init__(self, model, space=None, optimizer=None, cost_withGradients=None, utility=None,
                 utility_params_samples=None):
        self.utility = utility
        super(maEI, self).__init__(model, space, optimizer, cost_withGradients)
        self.use_full_support = self.utility.parameter_dist.use_full_support
        self.n_hyps_samples = min(10, self.model.number_of_hyps_samples())
        self._fixed_theta = utility_params_samples      # explicit theta for parity (reference redraws per call)

    def _theta(self):
        if self.use_full_support:
            self.utility_params_samples = self.utility.parameter_dist.support
            self.utility_param_dist = np.atleast_1d(self.utility.parameter_dist.prob_dist)
        elif self._fixed_theta is not None:
            self.utility_params_samples = self._fixed_theta
        else:
            self.utility_params_samples = self.utility.parameter_dist.sample(self.n_theta_draw)

    def _combine(self, X, marginal_acqX, marginal_dacq_dX=None):
        if self.use_full_support:
            acqX = np.matmul(marginal_acqX, self.utility_param_dist)
            if marginal_dacq_dX is not None:
                dacq_dX = np.tensordot(marginal_dacq_dX, self.utility_param_dist, 1)
        else:
            acqX = np.sum(marginal_acqX, axis=1) / len(self.utility_params_samples)
            if marginal_dacq_dX is not None:
                dacq_dX = np.sum(marginal_dacq_dX, axis=2) / len(self.utility_params_samples)
        acqX = np.reshape(acqX, (X.shape[0], 1))
        if marginal_dacq_dX is None:
            return acqX
        return acqX, np.reshape(dacq_dX, X.shape)

    def _compute_acq(self, X):
        # maEI.py:38-54
        self._theta()
        X = np.atleast_2d(X)
        return self._combine(X, self._marginal_acq(X, self.utility_params_samples))

    def _compute_acq_withGradients(self, X):
        # maEI.py:57-78
        self._theta()
        X = np.atleast_2d(X)
        a, g = self._marginal_acq_with_gradient(X, self.utility_params_samples)
        return self._combine(X, a, g)

    def _marginal_acq(self, X, utility_params_samples):
        # maEI.py:81-98
        L = len(utility_params_samples)
        marginal_acqX = np.zeros((X.shape[0], L))
        n_h = self.n_hyps_samples
        for h in range(n_h):
            self.model.set_hyperparameters(h)
            meanX, varX = self.model.predict(X)
            marginal_best_so_far = self._marginal_best_so_far(utility_params_samples)
            for l in range(L):
                current_best = marginal_best_so_far[l]
                for i in range(X.shape[0]):
                    mu = np.dot(utility_params_samples[l], meanX[:, i])
                    sigma = np.sqrt(np.dot(np.square(utility_params_samples[l]), varX[:, i]))
                    phi, Phi, u = self._get_quantiles(current_best, mu, sigma)
                    marginal_acqX[i, l] += self._value(mu, sigma, current_best, phi, Phi, u)
        return marginal_acqX / n_h

    def _value(self, mu, sigma, best, phi, Phi, u):
        return sigma * (u * Phi + phi)                                       # maEI.py:96

    def _marginal_acq_with_gradient(self, X, utility_params_samples):
        # maEI.py:101-126
        L = len(utility_params_samples)
        marginal_acqX = np.zeros((X.shape[0], L))
        marginal_dacq_dX = np.zeros((X.shape[0], X.shape[1], L))
        n_h = self.n_hyps_samples
        for h in range(n_h):
            self.model.set_hyperparameters(h)
            meanX, varX = self.model.predict(X)
            dmean_dX = self.model.posterior_mean_gradient(X)
            dvar_dX = self.model.posterior_variance_gradient(X)
            marginal_best_so_far = self._marginal_best_so_far(utility_params_samples)
            for l in range(L):
                best = marginal_best_so_far[l]
                for i in range(X.shape[0]):
                    mu = np.dot(utility_params_samples[l], meanX[:, i])
                    sigma = np.sqrt(np.dot(np.square(utility_params_samples[l]), varX[:, i]))
                    phi = norm.pdf((mu - best) / sigma)
                    Phi = norm.cdf((mu - best) / sigma)
                    marginal_acqX[i, l] += (mu - best) * Phi + sigma * phi
                    dmu_dX = np.matmul(utility_params_samples[l], dmean_dX[:, i, :])
                    dsigma_dX = 0.5 * np.matmul(np.square(utility_params_samples[l]), dvar_dX[:, i, :]) / sigma
                    marginal_dacq_dX[i, :, l] += dmu_dX * Phi + phi * dsigma_dX
        return marginal_acqX / n_h, marginal_dacq_dX / n_h

    def _marginal_best_so_far(self, utility_params_samples):
        # maEI.py:129-136
        L = len(utility_params_samples)
        marginal_best = np.empty(L)
        muX_eval = self.model.posterior_mean_at_evaluated_points()
        for l in range(L):
            marginal_best[l] = max(np.matmul(utility_params_samples[l], muX_eval))
        return marginal_best

    def _get_quantiles(self, fmax, m, s):
        # maEI.py:147-163 (maPI.py:141-158 adds the jitter)
        if isinstance(s, np.ndarray):
            s = s.copy()
            s[s < 1e-10] = 1e-10
        elif s < 1e-10:
            s = 1e-10
        u = (m - (fmax + self.jitter)) / s
        phi = np.exp(-0.5 * u**2) / np.sqrt(2 * np.pi)
        Phi = 0.5 * erfc(-u / np.sqrt(2))
        return (phi, Phi, u)


class maPI(maEI):
    """maPI.py: Phi(u) with u = (mu - (best + 1e-6)) / sigma; gradient (phi/sigma)(dmu - u dsigma)."""
    jitter = 1e-6                                                              # maPI.py:35
    n_theta_draw = 10                                                          # maPI.py:45 (value); :63 draws 3 for grads

    def _value(self, mu, sigma, best, phi, Phi, u):
        return Phi                                                             # maPI.py:91-92

    def _marginal_acq_with_gradient(self, X, utility_params_samples):
        # maPI.py:97-120
        L = len(utility_params_samples)
        marginal_acqX = np.zeros((X.shape[0], L))
        marginal_dacq_dX = np.zeros((X.shape[0], X.shape[1], L))
        n_h = self.n_hyps_samples
        for h in range(n_h):
            self.model.set_hyperparameters(h)
            meanX, varX = self.model.predict(X)
            dmean_dX = self.model.posterior_mean_gradient(X)
            dvar_dX = self.model.posterior_variance_gradient(X)
            marginal_best_so_far = self._marginal_best_so_far(utility_params_samples)
            for l in range(L):
                current_best = marginal_best_so_far[l]
                for i in range(X.shape[0]):
                    mu = np.dot(utility_params_samples[l], meanX[:, i])
                    sigma = np.sqrt(np.dot(np.square(utility_params_samples[l]), varX[:, i]))
                    phi, Phi, u = self._get_quantiles(current_best, mu, sigma)
                    marginal_acqX[i, l] += Phi
                    dmu_dX = np.matmul(utility_params_samples[l], dmean_dX[:, i, :])
                    dsigma_dX = 0.5 * np.matmul(np.square(utility_params_samples[l]), dvar_dX[:, i, :]) / sigma
                    marginal_dacq_dX[i, :, l] += (phi / sigma) * (dmu_dX - u * dsigma_dX)
        return marginal_acqX / n_h, marginal_dacq_dX / n_h


class EI(maEI):
    """EI.py: single-output (m=1), single hyper-sample (n_hyps_samples=1, EI.py:36) elementwise twin of maEI."""

    def __init__(self, *a, **kw):
        super(EI, self).__init__(*a, **kw)
        self.n_hyps_samples = 1


class PI(maPI):
    """PI.py: single-output, single hyper-sample twin of maPI."""

    def __init__(self, *a, **kw):
        super(PI, self).__init__(*a, **kw)
        self.n_hyps_samples = 1
