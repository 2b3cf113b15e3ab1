"""Config 1 plumbing on the GPU: the CBO loop (cbo.py) + multistart optimiser around the CUDA path, against the
same plumbing driven by the CPU oracle classes with the same seed."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _problem(seed):
    rng = np.random.default_rng(seed)
    d, m = 2, 3
    W = rng.standard_normal((m, d))

    def f(X):
        X = np.atleast_2d(X)
        return np.stack([np.sin(3 * X @ W[j]) + 0.3 * j * X[:, 0] for j in range(m)], axis=0)
    theta = f(np.array([[0.3, 0.7]])).T
    return d, m, f, theta


def _run(side, cuda_device, iters=2, seed=0):
    import bocf_b200 as B
    d, m, f, theta = _problem(3)
    np.random.seed(seed)
    space = B.Design_space(space=[{'name': 'var', 'type': 'continuous', 'domain': (0, 1), 'dimensionality': d}])
    objective = B.MultiObjective(f, as_list=False, output_dim=m)
    acq_opt = B.AcquisitionOptimizer(optimizer='lbfgs2', inner_optimizer='lbfgs2', space=space, n_starting=64, n_anchor=4)
    X_init = B.initial_design('random', space, 2 * (d + 1))
    var, ls, noise = np.full((1, m), 1.0), np.full((1, m, d), 0.4), np.full((1, m), 1e-4)
    if side == "cuda":
        model = B.multi_outputGP(m, n_samples=1, device=cuda_device)
        model.set_hyperparameter_samples(var, ls, noise, kind="matern52")
        pd = B.ParameterDistribution(support=theta, prob_dist=np.ones(1))
        U = B.Utility(parameter_dist=pd, composite="sumsq_target")
        acq = B.uEI_noiseless(model, space, optimizer=acq_opt, utility=U)
    else:
        from oracle.models import multi_outputGP
        from oracle.utility import make_utility, ParameterDistribution
        from oracle.acquisitions import uEI_noiseless
        model = multi_outputGP.from_hyper_samples("matern52", var, ls, noise)
        U = make_utility("sumsq_target", ParameterDistribution(support=theta, prob_dist=np.ones(1)))
        acq = uEI_noiseless(model, space, optimizer=acq_opt, utility=U, vectorised=True)
    expU = B.ExpectationUtility(
        lambda th, mu, v: -np.sum(np.square((mu.T - th).T), axis=0) - np.sum(v, axis=0),
        lambda th, mu, v: -np.concatenate((2 * (np.squeeze(mu) - th), np.ones((len(np.squeeze(v)),)))))
    if side == "cuda":
        bo = B.CBO(model, space, objective, acq, B.Sequential(acq), X_init, expectation_utility=expU)
    else:
        from oracle.cbo import CBO
        bo = CBO(model, space, objective, acq, B.Sequential(acq), X_init, expectation_utility=expU)
    bo.run_optimization(max_iter=iters)
    return bo, acq_opt


def test_cbo_loop_matches_oracle_driven_loop(cuda_device):
    bo_g, opt_g = _run("cuda", cuda_device)
    bo_c, _ = _run("cpu", cuda_device)
    assert len(bo_g.suggested_points) == 2 and bo_g.X.shape == (8, 2)
    # same selected next point (first acquisition: identical model, identical Z, identical candidates)
    np.testing.assert_allclose(bo_g.suggested_points[0], bo_c.suggested_points[0], atol=1e-5)
    np.testing.assert_allclose(bo_g.historical_optimal_values[0], bo_c.historical_optimal_values[0], rtol=1e-4, atol=1e-6)
    assert np.all(np.isfinite(bo_g.historical_optimal_values))
    assert opt_g.last_batches > 0          # the anchors' L-BFGS evaluations were batched into shared launches


def test_batched_multistart_equals_sequential(cuda_device):
    """Batching the anchors' f_df calls must not change any anchor's L-BFGS-B trajectory."""
    import bocf_b200 as B
    from tests.helpers import make_problem, product_model, product_utility
    P = make_problem(m=3, d=3, n=40, H=1, kind="rbf", composite="sumsq_target", N=8, S=64, seed=17)
    model = product_model(P, cuda_device)
    space = B.Design_space(space=[{'name': 'var', 'type': 'continuous', 'domain': (0, 1), 'dimensionality': 3}])
    res = []
    for batched in (True, False):
        opt = B.AcquisitionOptimizer(space, optimizer='lbfgs2', n_starting=50, n_anchor=5, batched=batched)
        acq = B.uEI_noiseless(model, space, optimizer=opt, utility=product_utility(P))
        acq.W_samples = P.Z
        np.random.seed(1)
        res.append(acq.optimize(x_baseline=P.X[:1]))
    np.testing.assert_allclose(res[0][0], res[1][0], atol=1e-9)
    np.testing.assert_allclose(np.asarray(res[0][1]).reshape(-1), np.asarray(res[1][1]).reshape(-1), atol=1e-12)


def _run_inferred(side, cuda_device, iters=2, seed=3):
    """The reference scripts' own configuration (test_1a.py:53, test_2a.py:49): fixed_hyps=False, so every updateModel
    runs ML-II + HMC and the acquisition averages over the n_samples hyper-sample instances."""
    import bocf_b200 as B
    d, m, f, theta = _problem(5)
    sampler = dict(n_burnin=4, subsample_interval=2, step_size=1e-1, leapfrog_steps=5, max_iters=200)
    np.random.seed(seed)
    space = B.Design_space(space=[{'name': 'var', 'type': 'continuous', 'domain': (0, 1), 'dimensionality': d}])
    objective = B.MultiObjective(f, as_list=False, output_dim=m)
    acq_opt = B.AcquisitionOptimizer(optimizer='lbfgs2', inner_optimizer='lbfgs2', space=space, n_starting=64, n_anchor=4)
    X_init = B.initial_design('random', space, 4 * (d + 1))
    if side == "cuda":
        model = B.multi_outputGP(m, exact_feval=[True] * m, n_samples=3, fixed_hyps=False, device=cuda_device, **sampler)
        pd = B.ParameterDistribution(support=theta, prob_dist=np.ones(1))
        U = B.Utility(parameter_dist=pd, composite="sumsq_target")
        acq = B.uEI_noiseless(model, space, optimizer=acq_opt, utility=U)
    else:
        from oracle.models import multi_outputGP
        from oracle.utility import make_utility, ParameterDistribution
        from oracle.acquisitions import uEI_noiseless
        model = multi_outputGP.inferred(m, kind="se", exact_feval=[True] * m, n_samples=3, **sampler)
        U = make_utility("sumsq_target", ParameterDistribution(support=theta, prob_dist=np.ones(1)))
        acq = uEI_noiseless(model, space, optimizer=acq_opt, utility=U, vectorised=True)
    expU = B.ExpectationUtility(
        lambda th, mu, v: -np.sum(np.square((mu.T - th).T), axis=0) - np.sum(v, axis=0),
        lambda th, mu, v: -np.concatenate((2 * (np.squeeze(mu) - th), np.ones((len(np.squeeze(v)),)))))
    if side == "cuda":
        bo = B.CBO(model, space, objective, acq, B.Sequential(acq), X_init, expectation_utility=expU)
    else:
        from oracle.cbo import CBO
        bo = CBO(model, space, objective, acq, B.Sequential(acq), X_init, expectation_utility=expU)
    bo.run_optimization(max_iter=iters)
    return bo, model


def test_cbo_loop_with_inferred_hyperparameters(cuda_device):
    bo_g, mod_g = _run_inferred("cuda", cuda_device)
    bo_c, mod_c = _run_inferred("cpu", cuda_device)
    assert mod_g.last_update == "hmc" and mod_g.n_hyper_samples_loaded() == 3
    assert len(bo_g.suggested_points) == 2 and np.all(np.isfinite(bo_g.historical_optimal_values))
    # first iteration: same random numbers, same chains (up to fp64 rounding amplified by the leap-frog dynamics), same
    # hyper-samples -> same selected point
    np.testing.assert_allclose(bo_g.suggested_points[0], bo_c.suggested_points[0], atol=1e-3)
    np.testing.assert_allclose(bo_g.historical_optimal_values[0], bo_c.historical_optimal_values[0], rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("branch", ["psi", "mc", "linear"])
@pytest.mark.parametrize("composite", ["sumsq_target", "neg_sum_exp", "rosen_composite", "exp_cos"])
def test_marginal_argmax_objectives_match_literal_loops(cuda_device, branch, composite):
    """cbo._current_marginal_argmax (cbo.py:121-235): the device objectives (psi_acq_kernel, mc_acq_kernel<_,3|4>) against
    the reference's per-candidate Python loops (oracle/cbo.py) on the same model, parameter and base samples."""
    import bocf_b200 as B
    from bocf_b200.utility import _host_psi
    from oracle.cbo import CBO as OracleCBO
    from tests.helpers import make_problem, oracle_model, product_model, oracle_utility, product_utility, rel_err, tol
    if branch == "linear":
        if composite != "sumsq_target":
            pytest.skip("one linear case")
        composite = "linear"
    if branch == "psi" and composite == "exp_cos":
        pytest.skip("the reference ships no closed form for exp_cos")
    m = 4
    P = make_problem(m=m, d=3, n=60, H=2, kind="matern52", composite=composite, N=37, S=8, seed=13, noise=1e-3)
    parameter = P.theta[0] if composite in ("sumsq_target", "linear") else np.array([0.8])
    om, pm = oracle_model(P), product_model(P, cuda_device)
    Uo, Up = oracle_utility(P), product_utility(P)
    if composite == "linear":
        Uo.linear = True
    expU = None
    if branch == "psi":
        expU = B.ExpectationUtility(*_host_psi(composite))

    class _Opt(object):                       # captures the objectives instead of optimising them
        def optimize(self, f=None, f_df=None, parallel=False):
            self.f, self.f_df = f, f_df
            return np.zeros((1, 3)), 0.0

    vals = {}
    for side, model, U, cls in (("cpu", om, Uo, OracleCBO), ("cuda", pm, Up, B.CBO)):
        bo = cls.__new__(cls)
        bo.model, bo.utility, bo.expectation_utility = model, U, expU
        bo.n_attributes, bo.n_hyps_samples = m, 2
        bo.acquisition = None
        bo._n_hyps = lambda: 2
        bo.evaluation_optimizer = _Opt()
        np.random.seed(5)                     # the MC branch draws its 50 base samples from numpy's global stream
        bo._current_marginal_argmax(parameter)
        f = bo.evaluation_optimizer.f(P.Xc)
        f2, g = bo.evaluation_optimizer.f_df(P.Xc)
        assert f.shape == (P.N, 1) and g.shape == (P.N, 3)
        np.testing.assert_allclose(f, f2, rtol=1e-12, atol=0)
        vals[side] = (f, g)
    assert rel_err(vals["cuda"][0], vals["cpu"][0]) < tol(1e-8), rel_err(vals["cuda"][0], vals["cpu"][0])
    assert rel_err(vals["cuda"][1], vals["cpu"][1]) < tol(1e-7), rel_err(vals["cuda"][1], vals["cpu"][1])
