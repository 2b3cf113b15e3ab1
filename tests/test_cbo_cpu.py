"""BASELINE.json configs[0] on CPU: the CBO loop (bocf_b200/cbo.py + optimization.py, the restated cbo.py / GPyOpt plumbing)
around the ORACLE model and acquisition on a test_1a-shaped problem (test_1a.py:19-96: attributes are GP samples on a
grid, objective = their posterior means, utility -sum_j (y_j - theta_j)^2 with theta taken at a reference point), with a
fixed seed.  No GPU: this covers the host plumbing the GPU loop tests share (tests/test_gpu_cbo.py)."""
import numpy as np
import pytest


def _build(seed, fixed_hyps=True, sampler=None):
    import bocf_b200 as B
    from oracle.gp import GPRegression
    from oracle.kern import Kern
    from oracle.models import multi_outputGP
    from oracle.utility import make_utility, ParameterDistribution
    from oracle.acquisitions import uEI_noiseless
    d, m = 3, 4
    I = np.linspace(0., 1., 5)
    grid = np.array([a.flatten() for a in np.meshgrid(*([I] * d))]).T            # 5^3 points (test_1a.py:21-23 uses 6^4)
    aux = []
    for j in range(m):                                                            # test_1a.py:24-33
        kern = Kern('se', d, variance=2., lengthscale=0.3)
        cov = kern.K(grid)
        Y = np.random.RandomState(j + 7).multivariate_normal(np.zeros(len(grid)), cov).reshape(-1, 1)
        aux.append(GPRegression(grid, Y, kern, noise_var=1e-10))

    def f(X):
        X = np.atleast_2d(X)
        return np.stack([g.posterior_mean(X)[:, 0] for g in aux], axis=0)

    np.random.seed(seed)
    space = B.Design_space(space=[{'name': 'var', 'type': 'continuous', 'domain': (0, 1), 'dimensionality': d}])
    objective = B.MultiObjective(f, as_list=False, output_dim=m)
    acq_opt = B.AcquisitionOptimizer(optimizer='lbfgs2', inner_optimizer='lbfgs2', space=space, n_starting=40, n_anchor=3)
    X_init = B.initial_design('random', space, 2 * (d + 1))
    theta = f(np.array([[0.3, 0.6, 0.5]])).T                                      # attainable target: best utility is 0
    if fixed_hyps:
        model = multi_outputGP.fixed_hyps(m, d, n_samples=1)
    else:
        model = multi_outputGP.inferred(m, kind="se", exact_feval=[True] * m, n_samples=2, **sampler)
    U = make_utility("sumsq_target", ParameterDistribution(support=theta, prob_dist=np.ones(1)))
    acq = uEI_noiseless(model, space, optimizer=acq_opt, utility=U, vectorised=True)
    expU = B.ExpectationUtility(
        lambda th, mu, v: -np.sum(np.square((mu.T - th).T), axis=0) - np.sum(v, axis=0),
        lambda th, mu, v: -np.concatenate((2 * (np.squeeze(mu) - th), np.ones((len(np.squeeze(v)),)))))
    from oracle.cbo import CBO            # host loop of bocf_b200.cbo + the reference's literal marginal-argmax objectives
    bo = CBO(model, space, objective, acq, B.Sequential(acq), X_init, expectation_utility=expU)
    return bo, acq_opt, d


def test_cbo_loop_runs_and_is_deterministic():
    runs = []
    for _ in range(2):
        bo, acq_opt, d = _build(seed=0)
        bo.run_optimization(max_iter=3)
        runs.append(bo)
    a, b = runs
    assert a.X.shape == (2 * (d + 1) + 3, d) and len(a.suggested_points) == 3
    assert np.all(a.X >= 0) and np.all(a.X <= 1)
    assert np.all(np.isfinite(a.historical_optimal_values))
    assert max(a.historical_optimal_values) <= 1e-6                               # utility is -sum of squares
    np.testing.assert_array_equal(a.X, b.X)                                       # fixed seed -> identical trace
    np.testing.assert_array_equal(np.vstack(a.suggested_points), np.vstack(b.suggested_points))
    np.testing.assert_array_equal(np.array(a.historical_optimal_values), np.array(b.historical_optimal_values))


def test_cbo_loop_with_inferred_hyperparameters_cpu():
    """fixed_hyps=False (what every reference script uses): ML-II + HMC inside every updateModel, H = 2 hyper-samples."""
    sampler = dict(n_burnin=2, subsample_interval=2, step_size=1e-1, leapfrog_steps=3, max_iters=50)
    bo, _, d = _build(seed=1, fixed_hyps=False, sampler=sampler)
    bo.run_optimization(max_iter=2)
    assert bo.X.shape == (2 * (d + 1) + 2, d)
    assert np.all(np.isfinite(bo.historical_optimal_values))
    out = bo.model.output[0]
    assert out.n_samples == 2 and out.inference.chain.shape[0] == 2 + 2 * 2
    assert np.all(out.inference.hmc_samples > 0)                                  # Logexp keeps every hyper-parameter positive


def test_negated_wrapper_host_fallback():
    """AcquisitionBase flips the sign on the device inside _run; results that did not come through _run are negated on
    the host, and the pending-sign state never leaks (also not through an exception)."""
    import bocf_b200 as B

    class Fake(B.AcquisitionBase):
        def __init__(self):
            pass

        def _compute_acq(self, x):
            return np.full((len(x), 1), 2.0)

        def _compute_acq_withGradients(self, x):
            if len(x) == 0:
                raise ValueError("empty")
            return np.full((len(x), 1), 2.0), np.ones((len(x), 3))

    a = Fake()
    assert np.all(a.acquisition_function(np.zeros((4, 3))) == -2.0)
    f, g = a.acquisition_function_withGradients(np.zeros((4, 3)))
    assert np.all(f == -2.0) and np.all(g == -1.0)
    try:
        a.acquisition_function_withGradients(np.zeros((0, 3)))
    except ValueError:
        pass
    assert a._sign == 1.0 and getattr(a._tls, "applied", False) is False


def test_product_cbo_refuses_cpu_models():
    """No CPU fallback: the product's marginal-argmax objectives are device calls; CPU models go through oracle.cbo."""
    import bocf_b200 as B
    import pytest
    bo, _, _ = _build(seed=0)
    with pytest.raises(TypeError):
        B.CBO._current_marginal_argmax(bo, np.zeros(4))


def test_theta_matrix_and_utility_consistency():
    import bocf_b200 as B
    import pytest
    from bocf_b200.utility import theta_matrix
    pd = B.ParameterDistribution(support=np.array([0.5, 1.0, 2.0]), prob_dist=np.ones(3) / 3)
    assert theta_matrix(B.Utility(parameter_dist=pd, composite="rosen_composite"), pd.support).shape == (3, 1)
    assert theta_matrix(B.Utility(parameter_dist=pd, composite="sumsq_target"), pd.support).shape == (1, 3)
    B.Utility(func=lambda th, y: -np.sum(np.square((y.T - th).T), axis=0), parameter_dist=pd, composite="sumsq_target")
    with pytest.raises(ValueError):
        B.Utility(func=lambda th, y: np.sum(y, axis=0), parameter_dist=pd, composite="sumsq_target")


def test_kernel_list_is_resolved_per_output():
    # multi_outputGP.py:23,38-44: output j is built from kernel[j]; the product keeps one family name when they all agree
    # and one name per output otherwise (host logic only: no device call before updateModel)
    import bocf_b200 as B
    from bocf_b200 import kern
    d = 3
    mod = B.multi_outputGP(3, kernel=[kern.Matern52(d), None, kern.RBF(d)], fixed_hyps=True)
    assert mod._kernel_kind() == ("matern52", "se", "rbf")
    assert B.multi_outputGP(2, kernel=[kern.RBF(d), kern.RBF(d)], fixed_hyps=True)._kernel_kind() == "rbf"
    assert B.multi_outputGP(2, fixed_hyps=True)._kernel_kind() == "se"
    var, ls, nz = np.ones((1, 3)), np.ones((1, 3, d)), np.full((1, 3), 1e-2)
    mod.set_hyperparameter_samples(var, ls, nz, kind=["se", "se", "se"])
    assert mod._hyp[0] == "se"
    mod.set_hyperparameter_samples(var, ls, nz, kind=("se", "matern32", "se"))
    assert mod._hyp[0] == ("se", "matern32", "se")
    with pytest.raises(AssertionError):
        mod.set_hyperparameter_samples(var, ls, nz, kind=["se", "rbf"])
