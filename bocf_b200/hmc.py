"""Hyper-parameter inference of GPModel.updateModel -- ML-II, then an HMC chain whose sub-sampled states become the
hyper-sample instances -- for all m outputs of a multi_outputGP in LOCKSTEP on the device likelihood.

Reference (one output at a time, m sequential chains of ~4000 Gram + Cholesky + gradient evaluations each):
  GPyOpt/models/gpmodel.py:50-99      priors Gamma.from_EV(2,4) on every kernel parameter and the noise variance; noise
                                       fixed at 1e-6 (exact_feval) / at noise_var / positive
  GPyOpt/models/gpmodel.py:117-126    optimize(max_iters=200); param_array *= 1 + 0.01 randn; HMC(stepsize).sample(
                                       n_burnin + n_samples * subsample_interval, leapfrog_steps);
                                       ss[n_burnin::subsample_interval] -> model_instances
  GPy/inference/mcmc/hmc.py:30-66     HMC.sample / _update / _computeH, identity mass matrix
  GPy/core/parameterization/priorizable.py:49-82, priors.py:264-331     log prior (+ log-Jacobian) and its gradient
  paramz 0.9.1 (not in the checkout; restated from its published source, see oracle/hmc.py header): Logexp transform,
  objective = -(log likelihood + log prior), _transform_gradients, optimizer_array views, lbfgsb with maxfun = maxiter.

Here every likelihood evaluation is ONE device pass over all m outputs (Gram, blocked Cholesky, L^-1, alpha, marginal
likelihood and its gradients: bocf_model_factorize + bocf_model_log_likelihood), so the m chains advance together:
  * ML-II: one scipy L-BFGS-B state machine per output (own thread), their objective calls batched per round;
  * HMC: positions of all outputs move with the same leap-frog index; accept / reject per output.
The random numbers are drawn from numpy's global generator in the reference's order (output-major: perturbation, then
one momentum vector and one uniform per sample), so a seeded run reproduces the reference's sequential chains.

Reference behaviour kept on purpose: the in-place perturbation does not re-run the inference (first Hamiltonian and first
half step use the likelihood terms of the un-perturbed optimum); it also scales a fixed noise variance; the chain state
persists between updateModel calls.

`evaluate(variance (m,), lengthscale (m,d), noise (m,)) -> (lml (m,), g_var (m,), g_len (m,d), g_noise (m,))` is the
only arithmetic dependency and must raise NotPositiveDefiniteError like the library does; multi_outputGP passes its
device pass.  There is no CPU implementation in this package.
"""
import threading

import numpy as np
import scipy.optimize
from scipy.special import gammaln

from ._lib import NotPositiveDefiniteError

_LIM_VAL = 36.0
_LOG_LIM_VAL = np.log(np.finfo(np.float64).max)


# ---- paramz.transformations.Logexp ----------------------------------------------------------------------------------
def logexp_f(x):
    return np.where(x > _LIM_VAL, x, np.log1p(np.exp(np.clip(x, -_LOG_LIM_VAL, _LIM_VAL))))


def logexp_finv(f):
    return np.where(f > _LIM_VAL, f, np.log(np.expm1(f)))


def logexp_gradfactor(f, df):
    return df * np.where(f > _LIM_VAL, 1., -np.expm1(-f))


def logexp_log_jacobian(f):
    return np.where(f > _LIM_VAL, f, np.log(np.expm1(f))) - f


def logexp_log_jacobian_grad(f):
    return 1. / np.expm1(f)


class GammaPrior(object):
    """priors.py:264-331; from_EV(2, 4) -> a = 1, b = 0.5."""

    def __init__(self, E=2., V=4.):
        self.a = float(np.square(E) / V)
        self.b = float(E / V)
        self.constant = -gammaln(self.a) + self.a * np.log(self.b)

    def lnpdf(self, x):
        return self.constant + (self.a - 1) * np.log(x) - self.b * x

    def lnpdf_grad(self, x):
        return (self.a - 1.) / x - self.b


class _Output(object):
    """Parameter bookkeeping of one output: param_array = [variance, lengthscale(s), noise] and its un-fixed mask."""

    def __init__(self, variance, lengthscale, noise, fix_noise, instance_noise):
        self.n_len = int(np.asarray(lengthscale).size)
        self.param_array = np.concatenate([[float(variance)], np.asarray(lengthscale, dtype=float).reshape(-1), [float(noise)]])
        self.free = np.ones(self.param_array.size, dtype=bool)
        self.free[-1] = not fix_noise
        self.fix_noise = bool(fix_noise)
        self.instance_noise = float(instance_noise)
        self.lml = None                 # likelihood terms of the LAST inference (may be stale w.r.t. param_array)
        self.dlml = None
        self.fail_count = 0

    @property
    def optimizer_array(self):
        return logexp_finv(self.param_array[self.free])

    def set_optimizer_array(self, x):
        self.param_array[self.free] = logexp_f(np.asarray(x, dtype=float))


class HyperInference(object):
    def __init__(self, evaluate, input_dim, kernels, noises, fix_noise, instance_noise, n_samples=10, n_burnin=100,
                 subsample_interval=10, step_size=1e-1, leapfrog_steps=20, max_iters=200):
        """kernels: per output (variance, lengthscale array of size 1 (shared) or d); noises: initial noise variances."""
        self.evaluate = evaluate
        self.d = int(input_dim)
        self.out = [_Output(k[0], k[1], nz, fx, inz) for k, nz, fx, inz in zip(kernels, noises, fix_noise, instance_noise)]
        self.m = len(self.out)
        self.n_samples, self.n_burnin, self.subsample_interval = int(n_samples), int(n_burnin), int(subsample_interval)
        self.step_size, self.leapfrog_steps, self.max_iters = float(step_size), int(leapfrog_steps), int(max_iters)
        self.prior = GammaPrior(2., 4.)
        self.allowed_failures = 10
        self.device_passes = 0
        self.chain = None
        self.optimum = None

    # ---- one device pass for all outputs ------------------------------------------------------------------------------
    def _pack(self):
        var = np.array([o.param_array[0] for o in self.out])
        ls = np.stack([np.broadcast_to(o.param_array[1:1 + o.n_len], (self.d,)) if o.n_len == 1
                       else o.param_array[1:1 + o.n_len] for o in self.out])
        nz = np.array([o.param_array[-1] for o in self.out])
        return var, np.ascontiguousarray(ls, dtype=np.float64), nz

    def _store(self, j, lml, gv, gl, gn):
        o = self.out[j]
        o.lml = float(lml[j])
        g_len = [gl[j].sum()] if o.n_len == 1 else gl[j]         # shared lengthscale: stationary.py:213-215, se.py:185
        o.dlml = np.concatenate([[gv[j]], np.asarray(g_len, dtype=float), [gn[j]]])

    def _infer(self, which=None):
        """Re-run the inference at the current param_array of every output; returns the set of outputs whose covariance
        was not positive definite even with jitter (their likelihood terms stay stale, like a failed paramz update)."""
        which = range(self.m) if which is None else which
        try:
            self.device_passes += 1
            res = self.evaluate(*self._pack())
            for j in which:
                self._store(j, *res)
            return set()
        except NotPositiveDefiniteError:
            if self.m == 1:
                return {0}
        # isolate the offending outputs: each requested output alone against known-good stand-ins for the others
        failed = set()
        var, ls, nz = self._pack()
        good = getattr(self, "_last_good", None)
        if good is None:
            raise NotPositiveDefiniteError(-4, "not positive definite, even with jitter.")
        for j in which:
            v2, l2, n2 = good[0].copy(), good[1].copy(), good[2].copy()
            v2[j], l2[j], n2[j] = var[j], ls[j], nz[j]
            try:
                self.device_passes += 1
                res = self.evaluate(v2, l2, n2)
                self._store(j, *res)
            except NotPositiveDefiniteError:
                failed.add(j)
        return failed

    def _remember_good(self):
        self._last_good = self._pack()

    # ---- objective pieces (per output, host) ---------------------------------------------------------------------------
    def _objective(self, o):
        x = o.param_array
        log_prior = float(np.sum(self.prior.lnpdf(x)) + np.sum(logexp_log_jacobian(x[o.free])))
        return -o.lml - log_prior

    def _objective_gradient_t(self, o):
        """_transform_gradients(objective_function_gradients()): gradient w.r.t. the optimizer array."""
        x = o.param_array
        dprior = self.prior.lnpdf_grad(x) * np.ones(x.size)
        dprior[o.free] += logexp_log_jacobian_grad(x[o.free])
        g = -(o.dlml + dprior)
        return logexp_gradfactor(x[o.free], g[o.free])

    # ---- ML-II: one L-BFGS-B run per output, objective calls batched -------------------------------------------------
    def optimize(self):
        m = self.m
        cv = threading.Condition()
        pending, results = {}, {}
        active = [m]
        errors = []

        def flush_locked():
            keys = sorted(pending)
            try:
                for j in keys:
                    self.out[j].set_optimizer_array(pending[j])
                failed = self._infer(keys)
            except BaseException as e:                       # no device / library error: wake every waiting run
                errors.append(e)
                for j in keys:
                    results[j] = None
                pending.clear()
                cv.notify_all()
                return
            if not failed:
                self._remember_good()
            for j in keys:
                o = self.out[j]
                if j in failed:                              # paramz Model._objective_grads, except branch
                    if o.fail_count >= self.allowed_failures:
                        errors.append(NotPositiveDefiniteError(-4, "not positive definite, even with jitter."))
                    o.fail_count += 1
                    results[j] = (np.inf, np.clip(self._objective_gradient_t(o), -1e10, 1e10))
                else:
                    o.fail_count = 0
                    results[j] = (self._objective(o), self._objective_gradient_t(o))
            pending.clear()
            cv.notify_all()

        def call(j, x):
            with cv:
                pending[j] = np.array(x, dtype=float)
                if len(pending) == active[0]:
                    flush_locked()
                else:
                    while j not in results:
                        cv.wait()
                if errors:
                    raise errors[0]
                return results.pop(j)

        x_opt = [None] * m

        def run(j):
            try:
                o = self.out[j]
                if not np.any(o.free):
                    x_opt[j] = o.optimizer_array
                else:
                    res = scipy.optimize.fmin_l_bfgs_b(lambda x: call(j, x), o.optimizer_array, maxfun=self.max_iters,
                                                       maxiter=self.max_iters)
                    x_opt[j] = res[0]
            except BaseException as e:                       # pragma: no cover
                errors.append(e)
            finally:
                with cv:
                    active[0] -= 1
                    if active[0] > 0 and len(pending) == active[0]:
                        flush_locked()

        threads = [threading.Thread(target=run, args=(j,)) for j in range(m)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]
        for j in range(m):                                   # Model.optimize: self.optimizer_array = opt.x_opt
            self.out[j].set_optimizer_array(x_opt[j])
        if self._infer():
            raise NotPositiveDefiniteError(-4, "not positive definite, even with jitter.")
        self._remember_good()
        self.optimum = [o.param_array.copy() for o in self.out]

    # ---- HMC, all outputs with the same leap-frog index ---------------------------------------------------------------
    def sample(self, momenta, uniforms):
        """momenta[j]: (num_samples, P_j); uniforms[j]: (num_samples,).  Returns chain[j] (num_samples, P_j)."""
        num = len(uniforms[0])
        outs = self.out
        chain = [np.empty((num, int(o.free.sum()))) for o in outs]
        eps = self.step_size
        for i in range(num):
            p = [momenta[j][i].copy() for j in range(self.m)]
            H_old, theta_old = [], []
            for j, o in enumerate(outs):
                H_old.append(self._objective(o) + p[j].size * np.log(2 * np.pi) / 2. + np.dot(p[j], p[j]) / 2.)
                theta_old.append(o.optimizer_array.copy())
                chain[j][i] = o.param_array[o.free]
            for _ in range(self.leapfrog_steps):             # hmc.py:58-62
                for j, o in enumerate(outs):
                    p[j] += -eps / 2. * self._objective_gradient_t(o)
                    o.set_optimizer_array(o.optimizer_array + eps * p[j])
                if self._infer():
                    raise NotPositiveDefiniteError(-4, "not positive definite, even with jitter.")
                for j, o in enumerate(outs):
                    p[j] += -eps / 2. * self._objective_gradient_t(o)
            rejected = False
            for j, o in enumerate(outs):
                H_new = self._objective(o) + p[j].size * np.log(2 * np.pi) / 2. + np.dot(p[j], p[j]) / 2.
                k = 1. if H_old[j] > H_new else np.exp(H_old[j] - H_new)
                if uniforms[j][i] < k:
                    chain[j][i] = o.param_array[o.free]
                else:
                    o.set_optimizer_array(theta_old[j])      # hmc.py:56
                    rejected = True
            if rejected and self._infer():
                raise NotPositiveDefiniteError(-4, "not positive definite, even with jitter.")
        return chain

    def draw_randomness(self, num_samples):
        """numpy global-generator draws in the reference's order: for each output, the perturbation of the whole
        param_array (gpmodel.py:118), then per sample one momentum vector (hmc.py:43) and one uniform (hmc.py:53)."""
        perturb, momenta, uniforms = [], [], []
        for o in self.out:
            P = int(o.free.sum())
            perturb.append(np.random.randn(o.param_array.size))
            mom = np.empty((num_samples, P))
            uni = np.empty(num_samples)
            for i in range(num_samples):
                mom[i] = np.random.multivariate_normal(np.zeros(P), np.eye(P))
                uni[i] = np.random.rand()
            momenta.append(mom)
            uniforms.append(uni)
        return perturb, momenta, uniforms

    # ---- GPModel.updateModel ---------------------------------------------------------------------------------------------
    def update(self):
        """Returns (variance (H,m), lengthscale (H,m,d), noise (H,m)) of the n_samples hyper-sample instances."""
        self.optimize()
        num = self.n_burnin + self.n_samples * self.subsample_interval
        perturb, momenta, uniforms = self.draw_randomness(num)
        for o, e in zip(self.out, perturb):
            o.param_array[:] = o.param_array * (1. + e * 0.01)
        self.chain = self.sample(momenta, uniforms)
        H = self.n_samples
        var = np.empty((H, self.m))
        ls = np.empty((H, self.m, self.d))
        nz = np.empty((H, self.m))
        self.hmc_samples = []
        for j, o in enumerate(self.out):
            s = self.chain[j][self.n_burnin::self.subsample_interval][:H]
            self.hmc_samples.append(s)
            var[:, j] = s[:, 0]
            ls[:, j, :] = s[:, 1:2] if o.n_len == 1 else s[:, 1:1 + o.n_len]
            nz[:, j] = o.instance_noise if o.fix_noise else s[:, -1]
        return var, ls, nz
