"""Hyper-parameter inference of GPModel.updateModel: ML-II + HMC.

CPU: (1) the oracle restatement (oracle/hmc.py) reproduces chains produced by the reference's own HMC class, Gamma prior
and inference code (tests/golden/make_golden_hmc.py); (2) the product's lockstep driver (bocf_b200/hmc.py: batched
L-BFGS-B runs, all outputs' chains advancing together, the reference's random-number order) reproduces the oracle's
sequential per-output chains when both are given the same likelihood function.
GPU: multi_outputGP.updateModel with the device likelihood against the oracle chains; the hyper-samples it loads are the
ones the posterior / acquisition path then uses."""
import glob
import os

import numpy as np
import pytest

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "hmc_*.npz")))
IDS = [os.path.basename(p)[4:-4] for p in GOLDEN]


def _oracle_from_golden(z, **over):
    from oracle.hmc import GPModelHMC
    nv = float(z["noise_var"])
    kw = dict(kind=str(z["kind"]), ARD=bool(z["ARD"]), exact_feval=bool(z["exact_feval"]), noise_var=None if np.isnan(nv) else nv,
              n_samples=int(z["n_samples"]), n_burnin=int(z["n_burnin"]), subsample_interval=int(z["subsample_interval"]),
              step_size=float(z["step_size"]), leapfrog_steps=int(z["leapfrog_steps"]), max_iters=int(z["max_iters"]))
    kw.update(over)
    return GPModelHMC(**kw)


def test_golden_files_present():
    assert len(GOLDEN) >= 5


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_oracle_chain_matches_reference_hmc(path):
    z = np.load(path)
    g = _oracle_from_golden(z)
    np.random.seed(int(z["seed"]))
    s = g.updateModel(z["X"], z["Y"])
    np.testing.assert_allclose(g.optimum, z["optimum"], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(g.chain, z["chain"], rtol=2e-6, atol=1e-12)
    np.testing.assert_allclose(s, z["hmc_samples"], rtol=2e-6, atol=1e-12)
    np.testing.assert_allclose(g.model.param_array, z["final_param_array"], rtol=2e-6, atol=1e-14)


def test_logexp_and_prior_identities():
    from oracle.hmc import Logexp, Gamma
    from bocf_b200 import hmc as H
    x = np.array([-700., -30., -1e-3, 0., 0.5, 20., 35.9, 36.1, 500.])
    f = Logexp.f(x)
    assert np.all(f >= 0) and np.all(np.isfinite(f))
    np.testing.assert_array_equal(H.logexp_f(x), f)
    ok = (x > -30) & (x < 30)
    np.testing.assert_allclose(Logexp.finv(f[ok]), x[ok], rtol=1e-9, atol=1e-12)
    p = np.array([1e-6, 1e-2, 0.7, 3.0, 40.0])
    np.testing.assert_array_equal(H.logexp_finv(p), Logexp.finv(p))
    np.testing.assert_array_equal(H.logexp_gradfactor(p, 2.0 * p), Logexp.gradfactor(p, 2.0 * p))
    np.testing.assert_array_equal(H.logexp_log_jacobian(p), Logexp.log_jacobian(p))
    np.testing.assert_array_equal(H.logexp_log_jacobian_grad(p), Logexp.log_jacobian_grad(p))
    # gradfactor is df/dx at x = finv(f):  d/dx log(1 + e^x) = 1 - e^-f
    e = 1e-6
    fd = (Logexp.f(Logexp.finv(p[:4]) + e) - Logexp.f(Logexp.finv(p[:4]) - e)) / (2 * e)
    np.testing.assert_allclose(Logexp.gradfactor(p[:4], 1.0), fd, rtol=1e-6)
    ga, gb = Gamma.from_EV(2., 4.), H.GammaPrior(2., 4.)
    assert (ga.a, ga.b) == (1.0, 0.5) == (gb.a, gb.b)
    np.testing.assert_array_equal(ga.lnpdf(p), gb.lnpdf(p))
    np.testing.assert_array_equal(ga.lnpdf_grad(p), gb.lnpdf_grad(p) * np.ones_like(p))


# ---- the lockstep driver against the sequential oracle, same likelihood function ---------------------------------
def _problem(m=3, n=26, d=3, seed=3):
    rng = np.random.default_rng(seed)
    X = rng.uniform(size=(n, d))
    Y = [np.sin(3.0 * X[:, :1] + j) + X[:, 1:2] * X[:, 2 % d:2 % d + 1] + 0.05 * rng.standard_normal((n, 1)) + 0.2 * j
         for j in range(m)]
    return X, Y


SETTINGS = dict(n_samples=3, n_burnin=4, subsample_interval=2, step_size=1e-1, leapfrog_steps=5, max_iters=200)


def _oracle_sequential(kind, X, Y, ARD, exact, noise_var, seed, rounds=1, settings=SETTINGS):
    """The reference's order: output after output, each with its own ML-II, perturbation and chain."""
    from oracle.hmc import GPModelHMC
    m = len(Y)
    models = [GPModelHMC(kind=kind, ARD=ARD[j], exact_feval=exact[j], noise_var=noise_var[j], **settings) for j in range(m)]
    np.random.seed(seed)
    for r in range(rounds):
        Xr, Yr = (X, Y) if r == rounds - 1 else (X[:-(rounds - 1 - r)], [y[:-(rounds - 1 - r)] for y in Y])
        for j in range(m):
            models[j].updateModel(Xr, Yr[j])
    return models


def _numpy_evaluator(kind, X, Y, ARD_flags):
    """A likelihood evaluator with the device pass's signature, built on the oracle's GPRegression."""
    from oracle.gp import GPRegression
    from oracle.kern import Kern
    from bocf_b200 import NotPositiveDefiniteError
    d = X.shape[1]

    def evaluate(variance, lengthscale, noise):
        m = len(Y)
        lml, gv, gl, gn = np.zeros(m), np.zeros(m), np.zeros((m, d)), np.zeros(m)
        for j in range(m):
            try:
                g = GPRegression(X, Y[j], Kern(kind, d, variance[j], lengthscale[j], ARD=True), noise[j])
            except np.linalg.LinAlgError:
                raise NotPositiveDefiniteError(-4, "not positive definite, even with jitter.")
            lml[j] = g.log_likelihood()
            gv[j], gl[j], gn[j] = g.likelihood_gradients()
        return lml, gv, gl, gn
    return evaluate


@pytest.mark.parametrize("kind", ["se", "matern52"])
def test_lockstep_driver_matches_sequential_oracle(kind):
    from bocf_b200.hmc import HyperInference
    X, Y = _problem()
    m, d = len(Y), X.shape[1]
    ARD, exact, noise_var = [True, False, True], [False, True, False], [None, None, 3e-3]
    seed = 17
    ref = _oracle_sequential(kind, X, Y, ARD, exact, noise_var, seed)
    kernels = [(1.0, np.ones(d) if ARD[j] else np.ones(1)) for j in range(m)]
    noises = [float(np.var(Y[0])) * 0.01, 1e-6, 3e-3]
    hi = HyperInference(_numpy_evaluator(kind, X, Y, ARD), d, kernels, noises, fix_noise=[False, True, True],
                        instance_noise=[noises[0], 1e-6, 3e-3], **SETTINGS)
    np.random.seed(seed)
    var, ls, nz = hi.update()
    for j in range(m):
        np.testing.assert_allclose(hi.optimum[j], ref[j].optimum, rtol=1e-7, atol=1e-13)
        np.testing.assert_allclose(hi.chain[j], ref[j].chain, rtol=1e-6, atol=1e-13)
        v, l, z = ref[j].hyper_samples(d)
        np.testing.assert_allclose(var[:, j], v, rtol=1e-6)
        np.testing.assert_allclose(ls[:, j], l, rtol=1e-6)
        np.testing.assert_allclose(nz[:, j], z, rtol=1e-6)
    assert nz[0, 1] == 1e-6 and nz[0, 2] == 3e-3           # fixed noise: the instances keep their own value
    # one device pass serves all outputs: far fewer passes than the m sequential chains need inferences
    num = SETTINGS["n_burnin"] + SETTINGS["n_samples"] * SETTINGS["subsample_interval"]
    assert hi.device_passes <= SETTINGS["max_iters"] + 2 + num * (SETTINGS["leapfrog_steps"] + 1)


# ---- GPU: the device likelihood drives the same chains -----------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["se", "rbf", "matern52", "matern32"])
def test_update_model_samples_hyperparameters_like_the_reference(kind, cuda_device):
    import bocf_b200 as B
    X, Y = _problem(m=3, n=40, d=3, seed=4)
    m, d = len(Y), X.shape[1]
    ARD, exact, noise_var = [True, False, True], [False, True, False], [None, None, 3e-3]
    seed = 5
    ref = _oracle_sequential(kind, X, Y, ARD, exact, noise_var, seed)
    K = B.kern.BY_KIND[kind]
    mod = B.multi_outputGP(m, kernel=[K(d, variance=1., ARD=ARD[j]) for j in range(m)], noise_var=noise_var,
                           exact_feval=exact, ARD=ARD, device=cuda_device, **SETTINGS)
    np.random.seed(seed)
    mod.updateModel(X, Y)
    assert mod.last_update == "hmc" and mod.n_hyper_samples_loaded() == SETTINGS["n_samples"]
    var, ls, nz = mod.get_hyperparameters_samples()
    for j in range(m):
        np.testing.assert_allclose(mod._inference.optimum[j], ref[j].optimum, rtol=1e-5, atol=1e-10)
        np.testing.assert_allclose(mod._inference.chain[j], ref[j].chain, rtol=1e-4, atol=1e-10)
        v, l, z = ref[j].hyper_samples(d)
        np.testing.assert_allclose(var[:, j], v, rtol=1e-4)
        np.testing.assert_allclose(ls[:, j], l, rtol=1e-4)
        np.testing.assert_allclose(nz[:, j], z, rtol=1e-4)
    # the loaded instances are what the prediction path uses: posterior of hyper-sample h == oracle model with those hypers
    from oracle.models import multi_outputGP as OM
    om = OM.from_hyper_samples(kind, var, ls, nz)
    om.updateModel(X, Y)
    Xc = np.random.default_rng(0).uniform(size=(50, d))
    for h in range(SETTINGS["n_samples"]):
        mod.set_hyperparameters(h)
        om.set_hyperparameters(h)
        np.testing.assert_allclose(mod.posterior_mean(Xc), om.posterior_mean(Xc), rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(mod.posterior_variance(Xc), om.posterior_variance(Xc), rtol=1e-6)


@pytest.mark.gpu
def test_chain_state_persists_across_updates(cuda_device):
    """Second updateModel (one more observation): ML-II starts from the last chain state, like self.model in the reference."""
    import bocf_b200 as B
    X, Y = _problem(m=2, n=31, d=2, seed=8)
    seed = 9
    ref = _oracle_sequential("se", X, Y, [True, True], [False, False], [None, None], seed, rounds=2)
    mod = B.multi_outputGP(2, device=cuda_device, **SETTINGS)
    np.random.seed(seed)
    mod.updateModel(X[:-1], [y[:-1] for y in Y])
    mod.updateModel(X, Y)
    for j in range(2):
        np.testing.assert_allclose(mod._inference.chain[j], ref[j].chain, rtol=1e-4, atol=1e-10)


@pytest.mark.gpu
def test_hyper_inference_none_keeps_initial_values(cuda_device):
    import bocf_b200 as B
    X, Y = _problem(m=2, n=20, d=2, seed=1)
    mod = B.multi_outputGP(2, device=cuda_device, hyper_inference="none")
    mod.updateModel(X, Y)
    var, ls, nz = mod.get_hyperparameters_samples()
    assert var.shape == (1, 2) and np.all(var == 1.0) and np.all(ls == 1.0)
    np.testing.assert_allclose(nz[0], [0.01 * np.var(Y[0]), 0.01 * np.var(Y[1])])


def test_failed_factorisation_is_isolated_per_output():
    """One output's covariance not positive definite: the other outputs of the shared device pass keep their fresh
    likelihood terms, the failing one reports failure (ML-II then sees +inf there, like paramz's _objective_grads)."""
    from bocf_b200 import NotPositiveDefiniteError
    from bocf_b200.hmc import HyperInference
    m, d = 3, 2
    calls = []

    def evaluate(var, ls, nz):
        calls.append(var.copy())
        if var[1] > 5.0:                                     # output 1 "fails" for large variances
            raise NotPositiveDefiniteError(-4, "not positive definite, even with jitter.")
        return -var, 0.5 * np.ones(m), -0.1 * np.ones((m, d)), 0.2 * np.ones(m)

    hi = HyperInference(evaluate, d, [(1.0, np.ones(d))] * m, [0.01] * m, [False] * m, [0.01] * m)
    assert hi._infer() == set()
    hi._remember_good()
    np.testing.assert_array_equal(hi.lml, [-1.0, -1.0, -1.0])
    hi.PA[:, 0] = [2.0, 9.0, 3.0]
    failed = hi._infer()
    assert failed == {1}
    np.testing.assert_array_equal(hi.lml, [-2.0, -1.0, -3.0])   # output 1 keeps its stale value
    # with nothing known-good to stand in for the other outputs the failure propagates
    hi2 = HyperInference(evaluate, d, [(9.0, np.ones(d))] * m, [0.01] * m, [False] * m, [0.01] * m)
    with pytest.raises(NotPositiveDefiniteError):
        hi2._infer()
    # ML-II survives a failing trial point: L-BFGS-B backs off from the +inf objective
    def evaluate2(var, ls, nz):
        if np.any(var > 50.0):
            raise NotPositiveDefiniteError(-4, "not positive definite, even with jitter.")
        lml = -(np.log(var) - 1.0) ** 2 - ((np.log(ls) - 0.5) ** 2).sum(1) - (np.log(nz) + 3.0) ** 2
        return (lml, -2 * (np.log(var) - 1.0) / var, -2 * (np.log(ls) - 0.5) / ls, -2 * (np.log(nz) + 3.0) / nz)
    hi3 = HyperInference(evaluate2, d, [(40.0, np.ones(d))] * m, [0.01] * m, [False] * m, [0.01] * m, max_iters=100)
    hi3.optimize()
    assert all(np.all(np.isfinite(o)) and o[0] < 50.0 for o in hi3.optimum)
