"""Oracle restatement of the hyper-parameter inference of GPModel.updateModel: ML-II followed by the HMC chain
whose sub-sampled states become the hyper-sample instances.  Test infrastructure only.

Follows, in the reference checkout:
  GPyOpt/models/gpmodel.py:50-99      _create_model: SE(variance=1, ARD) default kernel, noise 0.01 Var(Y), Gamma.from_EV(2,4)
                                       priors on every kernel parameter and on the noise variance, noise fixed at 1e-6
                                       (exact_feval) / at noise_var / constrained positive
  GPyOpt/models/gpmodel.py:102-128    updateModel: optimize(max_iters=200), multiplicative 1 % perturbation of the WHOLE
                                       param_array, HMC(stepsize), sample(n_burnin + n_samples * subsample_interval,
                                       hmc_iters=leapfrog_steps), ss[n_burnin::subsample_interval] -> instances
  GPy/inference/mcmc/hmc.py:20-66     HMC.__init__, sample, _update, _computeH (identity mass matrix)
  GPy/core/parameterization/priors.py:264-331         Gamma (lnpdf, lnpdf_grad, from_EV)
  GPy/core/parameterization/priorizable.py:49-82      log_prior / _log_prior_gradients incl. the log-Jacobian of the transform

paramz is NOT in the reference checkout (requirements.txt:15 pins paramz==0.9.1).  What this path uses from it is
restated from its published source -- PARITY UNPINNED for these pieces (they cannot be executed here):
  paramz/transformations.py  Logexp: f = log1p(exp(x)), finv = log(expm1(f)), gradfactor = df * -expm1(-f),
                             log_jacobian = log(expm1(f)) - f, log_jacobian_grad = 1/expm1(f), with _lim_val = 36
  paramz/model.py            objective_function = -log_likelihood - log_prior; objective_function_gradients =
                             -(dL + dprior); _transform_gradients (gradfactor, fixed entries dropped); _objective_grads
                             (np.inf and clipped gradients on LinAlgError); optimize -> 'lbfgsb'
  paramz/optimization/optimization.py  opt_lbfgsb: scipy fmin_l_bfgs_b(f_fp, x0, maxfun=max_iters, maxiter=max_iters)
  paramz/core/parameter_core.py        optimizer_array getter / setter (transformed, un-fixed view of param_array)
The HMC algorithm itself, the priors and the likelihood are pinned to the reference's own files
(tests/golden/make_golden_hmc.py runs the real hmc.py / priors.py / inference code on top of this adapter).

Reference behaviour that is kept on purpose:
  * param_array is perturbed in place (gpmodel.py:118) without re-running the inference, so the first Hamiltonian and
    the first leap-frog half step of a chain use the likelihood (and its gradients) of the UN-perturbed optimum while
    the prior terms already see the perturbed values;
  * the perturbation also hits a fixed noise variance (it is an entry of param_array);
  * the chain state persists between updateModel calls (the next ML-II starts from the last accepted state).
"""
import numpy as np
import scipy.optimize
from scipy.special import gammaln

from .gp import GPRegression
from .kern import Kern

_lim_val = 36.0
_log_lim_val = np.log(np.finfo(np.float64).max)


class Logexp(object):
    """paramz.transformations.Logexp (restated; parity unpinned)."""

    @staticmethod
    def f(x):
        return np.where(x > _lim_val, x, np.log1p(np.exp(np.clip(x, -_log_lim_val, _lim_val))))

    @staticmethod
    def finv(f):
        return np.where(f > _lim_val, f, np.log(np.expm1(f)))

    @staticmethod
    def gradfactor(f, df):
        return df * np.where(f > _lim_val, 1., -np.expm1(-f))

    @staticmethod
    def log_jacobian(model_param):
        return np.where(model_param > _lim_val, model_param, np.log(np.expm1(model_param))) - model_param

    @staticmethod
    def log_jacobian_grad(model_param):
        return 1. / (np.expm1(model_param))


class Gamma(object):
    """priors.py:264-331."""

    def __init__(self, a, b):
        self.a = float(a)
        self.b = float(b)
        self.constant = -gammaln(self.a) + a * np.log(b)

    @staticmethod
    def from_EV(E, V):
        return Gamma(np.square(E) / V, E / V)

    def lnpdf(self, x):
        return self.constant + (self.a - 1) * np.log(x) - self.b * x

    def lnpdf_grad(self, x):
        return (self.a - 1.) / x - self.b


class HyperModel(object):
    """The paramz.Model view of one GPRegression configured as in GPModel._create_model (gpmodel.py:50-77).

    param_array = [kernel variance, lengthscale(s), Gaussian noise variance]  (GPRegression links kern before likelihood,
    gp.py:99-101; Stationary/SE link variance before lengthscale, stationary.py:75-79 / se.py:31-35).
    `gp` needs: kern.variance, kern.lengthscale, noise_var, parameters_changed(), log_likelihood(),
    likelihood_gradients() -- the oracle's GPRegression, or the reference's own through the adapter of
    tests/golden/make_golden_hmc.py.  `fix_noise`: Gaussian_noise.constrain_fixed (gpmodel.py:70-75)."""

    def __init__(self, gp, fix_noise=False, prior=None):
        self.gp = gp
        self.fix_noise = bool(fix_noise)
        self.prior = Gamma.from_EV(2., 4.) if prior is None else prior          # gpmodel.py:66-67
        self.n_len = int(np.asarray(gp.kern.lengthscale).size)
        self.size = 2 + self.n_len
        self.param_array = np.concatenate([np.asarray(gp.kern.variance, dtype=float).reshape(-1)[:1],
                                           np.asarray(gp.kern.lengthscale, dtype=float).reshape(-1),
                                           [float(gp.noise_var)]])
        self._fixes_ = np.ones(self.size, dtype=bool)
        self._fixes_[-1] = not self.fix_noise
        self._fail_count = 0
        self._allowed_failures = 10
        self._update()

    # ---- paramz parameter views ----------------------------------------------------------------------------------
    def _update(self):
        """Push param_array into the GP and re-run the inference (paramz: trigger_update -> parameters_changed)."""
        self.gp.kern.variance[...] = self.param_array[0]
        self.gp.kern.lengthscale[...] = self.param_array[1:1 + self.n_len]
        self.gp.noise_var = float(self.param_array[-1])
        self.gp.parameters_changed()
        self._lml = self.gp.log_likelihood()
        g_var, g_len, g_noise = self.gp.likelihood_gradients()
        self._dlml = np.concatenate([[g_var], np.asarray(g_len, dtype=float).reshape(-1), [g_noise]])

    @property
    def optimizer_array(self):
        x = self.param_array.copy()
        x[self._fixes_] = Logexp.finv(self.param_array[self._fixes_])
        return x[self._fixes_]

    @optimizer_array.setter
    def optimizer_array(self, p):
        self.param_array[self._fixes_] = Logexp.f(np.asarray(p, dtype=float))
        self._update()

    @property
    def unfixed_param_array(self):
        return self.param_array[self._fixes_].copy()

    # ---- objective ------------------------------------------------------------------------------------------------
    def log_prior(self):
        # priorizable.py:49-65: every entry carries the Gamma prior; the log-Jacobian only where a Transformation
        # constrains the entry (a fixed noise is not transformed)
        x = self.param_array
        return float(np.sum(self.prior.lnpdf(x)) + np.sum(Logexp.log_jacobian(x[self._fixes_])))

    def _log_prior_gradients(self):
        # priorizable.py:67-82
        x = self.param_array
        ret = self.prior.lnpdf_grad(x) * np.ones(x.size)
        ret[self._fixes_] += Logexp.log_jacobian_grad(x[self._fixes_])
        return ret

    def objective_function(self):
        return -float(self._lml) - self.log_prior()

    def objective_function_gradients(self):
        return -(self._dlml + self._log_prior_gradients())

    def _transform_gradients(self, g):
        g = np.array(g, dtype=float)
        g[self._fixes_] = Logexp.gradfactor(self.param_array[self._fixes_], g[self._fixes_])
        return g[self._fixes_]

    def _objective_grads(self, x):
        try:
            self.optimizer_array = x
            obj_f, obj_grads = self.objective_function(), self._transform_gradients(self.objective_function_gradients())
            self._fail_count = 0
        except (np.linalg.LinAlgError, ZeroDivisionError, ValueError):
            if self._fail_count >= self._allowed_failures:
                raise
            self._fail_count += 1
            obj_f = np.inf
            obj_grads = np.clip(self._transform_gradients(self.objective_function_gradients()), -1e10, 1e10)
        return obj_f, obj_grads

    def optimize(self, max_iters=200):
        """paramz Model.optimize with the default 'lbfgsb' optimiser."""
        res = scipy.optimize.fmin_l_bfgs_b(self._objective_grads, self.optimizer_array, maxfun=max_iters, maxiter=max_iters)
        self.optimizer_array = res[0]
        return res


class HMC(object):
    """hmc.py:7-66 with the default identity mass matrix."""

    def __init__(self, model, M=None, stepsize=1e-1):
        self.model = model
        self.stepsize = stepsize
        self.p = np.empty_like(model.optimizer_array.copy())
        self.M = np.eye(self.p.size) if M is None else M
        self.Minv = np.linalg.inv(self.M)

    def sample(self, num_samples=1000, hmc_iters=20):
        params = np.empty((num_samples, self.p.size))
        for i in range(num_samples):
            self.p[:] = np.random.multivariate_normal(np.zeros(self.p.size), self.M)
            H_old = self._computeH()
            theta_old = self.model.optimizer_array.copy()
            params[i] = self.model.unfixed_param_array
            self._update(hmc_iters)
            H_new = self._computeH()
            if H_old > H_new:
                k = 1.
            else:
                k = np.exp(H_old - H_new)
            if np.random.rand() < k:
                params[i] = self.model.unfixed_param_array
            else:
                self.model.optimizer_array = theta_old
        return params

    def _update(self, hmc_iters):
        for i in range(hmc_iters):
            self.p[:] += -self.stepsize / 2. * self.model._transform_gradients(self.model.objective_function_gradients())
            self.model.optimizer_array = self.model.optimizer_array + self.stepsize * np.dot(self.Minv, self.p)
            self.p[:] += -self.stepsize / 2. * self.model._transform_gradients(self.model.objective_function_gradients())

    def _computeH(self):
        return (self.model.objective_function() + self.p.size * np.log(2 * np.pi) / 2. + np.log(np.linalg.det(self.M)) / 2.
                + np.dot(self.p, np.dot(self.Minv, self.p[:, None])) / 2.)


class GPModelHMC(object):
    """GPModel with its hyper-parameter inference (gpmodel.py:31-128) for ONE output."""

    def __init__(self, kind='se', kernel=None, noise_var=None, exact_feval=False, n_samples=10, n_burnin=100,
                 subsample_interval=10, step_size=1e-1, leapfrog_steps=20, ARD=False, max_iters=200):
        self.kind = kind
        self.kernel = kernel
        self.noise_var = noise_var
        self.exact_feval = exact_feval
        self.n_samples = n_samples
        self.n_burnin = n_burnin
        self.subsample_interval = subsample_interval
        self.step_size = step_size
        self.leapfrog_steps = leapfrog_steps
        self.ARD = ARD
        self.max_iters = max_iters
        self.model = None
        self.hmc_samples = None

    def _create_model(self, X, Y):
        d = X.shape[1]
        kern = Kern(self.kind, d, variance=1., ARD=self.ARD) if self.kernel is None else self.kernel      # :57-58
        noise_var = Y.var() * 0.01 if self.noise_var is None else self.noise_var                             # :64
        fix = False
        if self.exact_feval:                                                                                # :70-71
            noise_var, fix = 1e-6, True
        elif self.noise_var is not None:                                                                    # :72-73
            fix = True
        self.instance_noise = 1e-6 if self.exact_feval else noise_var
        self.model = HyperModel(GPRegression(X, Y, kern, noise_var), fix_noise=fix)

    def updateModel(self, X_all, Y_all):
        if self.model is None:
            self._create_model(X_all, Y_all)
        else:
            self.model.gp.set_XY(X_all, Y_all)
            self.model._update()
        self.model.optimize(max_iters=self.max_iters)                                                        # :117
        self.optimum = self.model.param_array.copy()
        self.model.param_array[:] = self.model.param_array * (1. + np.random.randn(self.model.param_array.size) * 0.01)
        self.hmc = HMC(self.model, stepsize=self.step_size)
        ss = self.hmc.sample(num_samples=self.n_burnin + self.n_samples * self.subsample_interval,
                             hmc_iters=self.leapfrog_steps)
        self.chain = ss
        self.hmc_samples = ss[self.n_burnin::self.subsample_interval]
        return self.hmc_samples

    def hyper_samples(self, d):
        """(variance (H,), lengthscale (H,d), noise (H,)) of the n_samples instances (gpmodel.py:122-126: un-fixed
        entries from the chain; a fixed noise keeps the instance's own value)."""
        s = self.hmc_samples[:self.n_samples]
        n_len = self.model.n_len
        var = s[:, 0].copy()
        ls = np.repeat(s[:, 1:2], d, axis=1) if n_len == 1 else s[:, 1:1 + n_len].copy()
        noise = np.full(len(s), self.instance_noise) if self.model.fix_noise else s[:, -1].copy()
        return var, ls, noise
