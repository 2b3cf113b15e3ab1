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


class HyperInference(object):
    """State of all outputs in one padded table: column 0 = kernel variance, columns 1..d = lengthscales (only column 1 for a
    shared lengthscale), column d+1 = noise variance; `valid` marks the entries an output has, `free` the un-fixed ones.
    The compact per-output param_array of the reference is PA[j, valid[j]] (same order)."""

    def __init__(self, evaluate, input_dim, kernels, noises, fix_noise, instance_noise, n_samples=10, n_burnin=100,
                 subsample_interval=10, step_size=1e-1, leapfrog_steps=20, max_iters=200):
        """kernels: per output (variance, lengthscale array of size 1 (shared) or d); noises: initial noise variances."""
        self.evaluate = evaluate
        self.d = d = int(input_dim)
        self.m = m = len(kernels)
        self.PA = np.ones((m, d + 2))
        self.valid = np.zeros((m, d + 2), dtype=bool)
        self.shared = np.zeros(m, dtype=bool)
        for j, (k, nz) in enumerate(zip(kernels, noises)):
            ls = np.asarray(k[1], dtype=float).reshape(-1)
            assert ls.size in (1, d)
            self.shared[j] = (ls.size == 1 and d > 1)
            self.PA[j, 0] = float(k[0])
            self.PA[j, 1:1 + ls.size] = ls
            self.PA[j, d + 1] = float(nz)
            self.valid[j, 0] = self.valid[j, d + 1] = True
            self.valid[j, 1:1 + ls.size] = True
        self.fix_noise = np.array([bool(f) for f in fix_noise])
        self.instance_noise = np.array([float(v) for v in instance_noise])
        self.free = self.valid.copy()
        self.free[self.fix_noise, d + 1] = False
        self.nfree = self.free.sum(1)
        self.lml = np.zeros(m)                   # likelihood terms of the LAST inference (may be stale w.r.t. PA)
        self.dlml = np.zeros((m, d + 2))
        self.fail_count = [0] * m
        self.n_samples, self.n_burnin, self.subsample_interval = int(n_samples), int(n_burnin), int(subsample_interval)
        self.step_size, self.leapfrog_steps, self.max_iters = float(step_size), int(leapfrog_steps), int(max_iters)
        self.prior = GammaPrior(2., 4.)
        self.allowed_failures = 10
        self.device_passes = 0
        self.chain = None
        self.optimum = None

    # ---- parameter views ------------------------------------------------------------------------------------------------
    def param_array(self, j):
        return self.PA[j, self.valid[j]]

    def _free_compact(self, j):
        return self.free[j, self.valid[j]]

    def optimizer_array(self, j):
        return logexp_finv(self.PA[j, self.free[j]])

    def set_optimizer_array(self, j, x):
        self.PA[j, self.free[j]] = logexp_f(np.asarray(x, dtype=float))

    def _optimizer_table(self):
        return np.where(self.free, logexp_finv(self.PA), 0.0)

    def _set_optimizer_table(self, X, rows=None):
        new = np.where(self.free, logexp_f(X), self.PA)
        if rows is None:
            self.PA = new
        else:
            self.PA[rows] = new[rows]

    # ---- one device pass for all outputs ------------------------------------------------------------------------------
    def _pack(self):
        d = self.d
        ls = np.where(self.shared[:, None], self.PA[:, 1:2], self.PA[:, 1:1 + d])
        return self.PA[:, 0].copy(), np.ascontiguousarray(ls, dtype=np.float64), self.PA[:, d + 1].copy()

    def _store(self, rows, lml, gv, gl, gn):
        d = self.d
        rows = np.asarray(list(rows), dtype=int)
        self.lml[rows] = np.asarray(lml)[rows]
        self.dlml[rows, 0] = np.asarray(gv)[rows]
        g = np.asarray(gl)[rows]
        sh = self.shared[rows]
        g = np.where(sh[:, None], 0.0, g)
        g[:, 0] = np.where(sh, np.asarray(gl)[rows].sum(1), g[:, 0])   # shared lengthscale: stationary.py:213-215, se.py:185
        self.dlml[rows, 1:1 + d] = g
        self.dlml[rows, d + 1] = np.asarray(gn)[rows]

    def _infer(self, which=None):
        """Re-run the inference at the current parameters of every output; returns the set of outputs whose covariance was
        not positive definite even with jitter (their likelihood terms stay stale, like a failed paramz update)."""
        which = range(self.m) if which is None else which
        try:
            self.device_passes += 1
            res = self.evaluate(*self._pack())
            self._store(which, *res)
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
                self._store([j], *res)
            except NotPositiveDefiniteError:
                failed.add(j)
        return failed

    def _remember_good(self):
        self._last_good = self._pack()

    # ---- objective pieces ---------------------------------------------------------------------------------------------------
    def _objective(self, j):
        """-log likelihood - log prior of output j (priorizable.py:49-65: Gamma prior on every entry, log-Jacobian of the
        transform on the un-fixed ones)."""
        x = self.param_array(j)
        log_prior = float(np.sum(self.prior.lnpdf(x)) + np.sum(logexp_log_jacobian(x[self._free_compact(j)])))
        return -self.lml[j] - log_prior

    def _objective_gradient_t(self, j):
        """_transform_gradients(objective_function_gradients()) of output j: gradient w.r.t. its optimizer array."""
        x = self.param_array(j)
        fc = self._free_compact(j)
        dprior = self.prior.lnpdf_grad(x) * np.ones(x.size)
        dprior[fc] += logexp_log_jacobian_grad(x[fc])
        g = -(self.dlml[j, self.valid[j]] + dprior)
        return logexp_gradfactor(x[fc], g[fc])

    def _objective_all(self):
        x = self.PA
        lp = np.where(self.valid, self.prior.lnpdf(x), 0.0).sum(1) + np.where(self.free, logexp_log_jacobian(x), 0.0).sum(1)
        return -self.lml - lp

    def _objective_gradient_t_all(self):
        x = self.PA
        dprior = self.prior.lnpdf_grad(x) + np.where(self.free, logexp_log_jacobian_grad(x), 0.0)
        return np.where(self.free, logexp_gradfactor(x, -(self.dlml + dprior)), 0.0)

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
                    self.set_optimizer_array(j, pending[j])
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
                if j in failed:                              # paramz Model._objective_grads, except branch
                    if self.fail_count[j] >= self.allowed_failures:
                        errors.append(NotPositiveDefiniteError(-4, "not positive definite, even with jitter."))
                    self.fail_count[j] += 1
                    results[j] = (np.inf, np.clip(self._objective_gradient_t(j), -1e10, 1e10))
                else:
                    self.fail_count[j] = 0
                    results[j] = (self._objective(j), self._objective_gradient_t(j))
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
                if self.nfree[j] == 0:
                    x_opt[j] = self.optimizer_array(j)
                else:
                    res = scipy.optimize.fmin_l_bfgs_b(lambda x: call(j, x), self.optimizer_array(j), maxfun=self.max_iters,
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
            self.set_optimizer_array(j, x_opt[j])
        if self._infer():
            raise NotPositiveDefiniteError(-4, "not positive definite, even with jitter.")
        self._remember_good()
        self.optimum = [self.param_array(j).copy() for j in range(m)]

    # ---- HMC, all outputs with the same leap-frog index (hmc.py:30-66) -------------------------------------------------
    def sample(self, momenta, uniforms):
        """momenta[j]: (num_samples, P_j); uniforms[j]: (num_samples,).  Returns chain[j] (num_samples, P_j)."""
        num = len(uniforms[0])
        m, eps = self.m, self.step_size
        MOM = np.zeros((num, m, self.d + 2))
        for j in range(m):
            MOM[:, j, self.free[j]] = momenta[j]
        U = np.stack([np.asarray(u, dtype=float) for u in uniforms], axis=1)         # (num, m)
        rec = np.empty((num, m, self.d + 2))
        const = self.nfree * np.log(2 * np.pi) / 2.
        for i in range(num):
            p = MOM[i].copy()
            H_old = self._objective_all() + const + (p * p).sum(1) / 2.
            theta_old = self._optimizer_table()
            rec[i] = self.PA
            for _ in range(self.leapfrog_steps):             # hmc.py:58-62
                p += -eps / 2. * self._objective_gradient_t_all()
                self._set_optimizer_table(self._optimizer_table() + eps * p)
                if self._infer():
                    raise NotPositiveDefiniteError(-4, "not positive definite, even with jitter.")
                p += -eps / 2. * self._objective_gradient_t_all()
            H_new = self._objective_all() + const + (p * p).sum(1) / 2.
            with np.errstate(over="ignore", invalid="ignore"):
                k = np.where(H_old > H_new, 1., np.exp(H_old - H_new))
            accept = U[i] < k
            rec[i][accept] = self.PA[accept]
            if not np.all(accept):
                self._set_optimizer_table(theta_old, rows=~accept)                   # hmc.py:56
                if self._infer():
                    raise NotPositiveDefiniteError(-4, "not positive definite, even with jitter.")
        return [rec[:, j, self.free[j]] for j in range(m)]

    def draw_randomness(self, num_samples):
        """numpy global-generator draws in the reference's order: for each output, the perturbation of the whole
        param_array (gpmodel.py:118), then per sample one momentum vector (hmc.py:43) and one uniform (hmc.py:53)."""
        perturb, momenta, uniforms = [], [], []
        for j in range(self.m):
            P = int(self.nfree[j])
            perturb.append(np.random.randn(int(self.valid[j].sum())))
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
        import time
        t0, p0 = time.perf_counter(), self.device_passes
        self.optimize()
        t1, p1 = time.perf_counter(), self.device_passes
        num = self.n_burnin + self.n_samples * self.subsample_interval
        perturb, momenta, uniforms = self.draw_randomness(num)
        for j, e in enumerate(perturb):
            self.PA[j, self.valid[j]] = self.param_array(j) * (1. + e * 0.01)
        self.chain = self.sample(momenta, uniforms)
        t2 = time.perf_counter()
        # where the wall time of this update went (scripts/bench_hmc.py reports it)
        self.timing = {"mlii_s": t1 - t0, "mlii_passes": p1 - p0, "hmc_s": t2 - t1, "hmc_passes": self.device_passes - p1}
        H, d = self.n_samples, self.d
        var = np.empty((H, self.m))
        ls = np.empty((H, self.m, d))
        nz = np.empty((H, self.m))
        self.hmc_samples = []
        for j in range(self.m):
            s = self.chain[j][self.n_burnin::self.subsample_interval][:H]
            self.hmc_samples.append(s)
            n_len = 1 if (self.shared[j] or d == 1) else d
            var[:, j] = s[:, 0]
            ls[:, j, :] = s[:, 1:2] if n_len == 1 else s[:, 1:1 + n_len]
            nz[:, j] = self.instance_noise[j] if self.fix_noise[j] else s[:, -1]
        return var, ls, nz
