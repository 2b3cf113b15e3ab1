"""Golden vectors for the hyper-parameter inference of GPModel.updateModel (ML-II + HMC).

What runs here is the REFERENCE's own code wherever it exists in the checkout:
  * GPy/inference/mcmc/hmc.py            the HMC class (sample / _update / _computeH), unmodified
  * GPy/core/parameterization/priors.py  Gamma.from_EV(2, 4)
  * GP.parameters_changed -> ExactGaussianInference.inference, kernel update_gradients_full, Gaussian.update_gradients
    (through tests/golden/ref_harness.py, as for the lml_* fixtures)
and the update sequence of GPyOpt/models/gpmodel.py:117-120 is followed statement by statement below.
paramz (absent from the checkout) supplies the Model protocol HMC drives -- optimizer_array, _transform_gradients,
objective_function[_gradients], the Logexp transform and the L-BFGS-B call; that part is the restatement in
oracle/hmc.py:HyperModel (parity unpinned for it, see its header).

    python tests/golden/make_golden_hmc.py        (needs /root/reference; writes tests/golden/hmc_<case>.npz)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402
from oracle.hmc import HyperModel  # noqa: E402

CASES = {
    # name: kind, ARD, exact_feval, noise_var, n, d, seed
    "se_ard": ("se", True, False, None, 28, 3, 5),
    "se_iso_exact": ("se", False, True, None, 24, 2, 6),
    "matern52_ard_fixednoise": ("matern52", True, False, 2e-3, 30, 3, 7),
    "rbf_iso": ("rbf", False, False, None, 26, 2, 8),
    "matern32_ard": ("matern32", True, False, None, 25, 2, 9),
}
N_SAMPLES, N_BURNIN, SUBSAMPLE, STEP, LEAPFROG, MAX_ITERS = 4, 6, 3, 1e-1, 6, 200


class RefGP(object):
    """The reference's GPRegression behind the attribute names HyperModel drives."""

    def __init__(self, g):
        self.g = g
        self.kern = g.kern

    @property
    def noise_var(self):
        return float(np.asarray(self.g.likelihood.variance).reshape(-1)[0])

    @noise_var.setter
    def noise_var(self, v):
        self.g.likelihood.variance[...] = v

    def parameters_changed(self):
        self.g.parameters_changed()

    def log_likelihood(self):
        return float(self.g._log_marginal_likelihood)

    def likelihood_gradients(self):
        g = self.g
        return (float(np.asarray(g.kern.variance.gradient).reshape(-1)[0]),
                np.asarray(g.kern.lengthscale.gradient, dtype=float).reshape(-1),
                float(np.asarray(g.likelihood.variance.gradient).reshape(-1)[0]))


def main():
    ns = rh.install()
    rh._mod("paramz.domains", _REAL="real", _POSITIVE="positive", _NEGATIVE="negative")
    priors = rh._load("GPy.core.parameterization.priors", "GPy/core/parameterization/priors.py")
    hmc_mod = rh._load("GPy.inference.mcmc.hmc", "GPy/inference/mcmc/hmc.py")
    kerns = {"se": ns.GPy.kern.SE, "rbf": ns.GPy.kern.RBF, "matern52": ns.GPy.kern.Matern52, "matern32": ns.GPy.kern.Matern32}
    for name, (kind, ARD, exact, noise_var, n, d, seed) in CASES.items():
        rng = np.random.default_rng(seed)
        X = rng.uniform(size=(n, d))
        Y = np.sin(4.0 * X[:, :1]) + X[:, 1:2] ** 2 + 0.05 * rng.standard_normal((n, 1))
        # gpmodel.py:50-77 (_create_model)
        kern = kerns[kind](d, variance=1., ARD=ARD)
        nv = Y.var() * 0.01 if noise_var is None else noise_var
        fix = False
        if exact:
            nv, fix = 1e-6, True
        elif noise_var is not None:
            fix = True
        g = ns.GPy.models.GPRegression(X, Y, kernel=kern, noise_var=nv)
        model = HyperModel(RefGP(g), fix_noise=fix, prior=priors.Gamma.from_EV(2., 4.))
        # gpmodel.py:117-120
        np.random.seed(seed)
        model.optimize(max_iters=MAX_ITERS)
        optimum = model.param_array.copy()
        model.param_array[:] = model.param_array * (1. + np.random.randn(model.param_array.size) * 0.01)
        hmc = hmc_mod.HMC(model, stepsize=STEP)
        ss = hmc.sample(num_samples=N_BURNIN + N_SAMPLES * SUBSAMPLE, hmc_iters=LEAPFROG)
        hmc_samples = ss[N_BURNIN::SUBSAMPLE]
        out = os.path.join(HERE, "hmc_%s.npz" % name)
        np.savez(out, kind=kind, ARD=ARD, exact_feval=exact, noise_var=np.nan if noise_var is None else noise_var,
                 X=X, Y=Y, seed=seed, n_samples=N_SAMPLES, n_burnin=N_BURNIN, subsample_interval=SUBSAMPLE,
                 step_size=STEP, leapfrog_steps=LEAPFROG, max_iters=MAX_ITERS,
                 optimum=optimum, chain=ss, hmc_samples=hmc_samples, final_param_array=model.param_array.copy())
        acc = int(np.sum(np.any(np.diff(ss, axis=0) != 0, axis=1)))
        print(out, "optimum", optimum.round(4), "moves", acc, "/", len(ss) - 1)


if __name__ == "__main__":
    main()
