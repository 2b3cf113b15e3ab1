"""BASELINE.json configs[0] at the shape the reference names (test_1a.py:19-33,53,60-96,151-152): d = 4, m = 5
attributes, each a GP sample on the 6^4 grid (SE kernel, variance 2, lengthscale 0.3, RandomState(j + 7)); the objective
is the posterior mean of those GPs; utility -sum_j (y_j - theta_j)^2 with theta = f(maximiser of attribute 0); EI-CF
(uEI_noiseless) through the CBO loop with the default optimiser (400 starts, 16 anchors), np.random.seed(seed),
fixed_hyps=True for a deterministic trace.  One builder for both sides of the comparison: the CUDA
product and the CPU oracle get the SAME objective (the oracle's GP posterior mean), initial design and seeds."""
import numpy as np
import scipy.optimize


def objective_function(d=4, m=5):
    from oracle.gp import GPRegression
    from oracle.kern import Kern
    I = np.linspace(0., 1., 6)
    grid = np.array([a.flatten() for a in np.meshgrid(*([I] * d))]).T            # 6^4 points, test_1a.py:21-23
    aux = []
    for j in range(m):                                                            # test_1a.py:24-33
        kern = Kern('se', d, variance=2., lengthscale=0.3)
        cov = kern.K(grid)
        Y = np.random.RandomState(j + 7).multivariate_normal(np.zeros(len(grid)), cov).reshape(-1, 1)
        aux.append(GPRegression(grid, Y, kern, noise_var=1e-10))

    def f(X):
        X = np.atleast_2d(X)
        return np.stack([g.posterior_mean(X)[:, 0] for g in aux], axis=0)
    return f


def build(side, f, seed=0, device="cuda:0", d=4, m=5):
    import bocf_b200 as B
    np.random.seed(seed)
    space = B.Design_space(space=[{'name': 'var', 'type': 'continuous', 'domain': (0, 1), 'dimensionality': d}])
    objective = B.MultiObjective(f, as_list=False, output_dim=m)
    acq_opt = B.AcquisitionOptimizer(optimizer='lbfgs2', inner_optimizer='lbfgs2', space=space)   # 400 starts, 16 anchors
    X_init = B.initial_design('random', space, 2 * (d + 1))                       # test_1a.py:58
    best = (np.inf, None)                                                         # theta: test_1a.py:60-81
    for x0 in np.random.rand(10, d):
        res = scipy.optimize.fmin_l_bfgs_b(lambda x: -f(x)[0, 0], x0, approx_grad=True, bounds=[(0, 1)] * d)
        if res[1] < best[0]:
            best = (res[1], res[0])
    theta = f(best[1]).T
    psi = (lambda th, mu, v: -np.sum(np.square((mu.T - th).T), axis=0) - np.sum(v, axis=0),
           lambda th, mu, v: -np.concatenate((2 * (np.squeeze(mu) - th), np.ones((len(np.squeeze(v)),)))))
    expU = B.ExpectationUtility(*psi)
    if side == "cuda":
        model = B.multi_outputGP(output_dim=m, fixed_hyps=True, device=device)
        U = B.Utility(parameter_dist=B.ParameterDistribution(support=theta, prob_dist=np.ones(1)), composite="sumsq_target")
        acq = B.uEI_noiseless(model, space, optimizer=acq_opt, utility=U)
        return B.CBO(model, space, objective, acq, B.Sequential(acq), X_init, expectation_utility=expU)
    from oracle.cbo import CBO
    from oracle.models import multi_outputGP
    from oracle.utility import make_utility, ParameterDistribution
    from oracle.acquisitions import uEI_noiseless
    model = multi_outputGP.fixed_hyps(m, d)
    U = make_utility("sumsq_target", ParameterDistribution(support=theta, prob_dist=np.ones(1)))
    acq = uEI_noiseless(model, space, optimizer=acq_opt, utility=U, vectorised=True)
    return CBO(model, space, objective, acq, B.Sequential(acq), X_init, expectation_utility=expU)


def run(side, f, iters, seed=0, device="cuda:0"):
    bo = build(side, f, seed=seed, device=device)
    bo.run_optimization(max_iter=iters)
    return {"suggested_points": np.vstack(bo.suggested_points).tolist(),
            "best_value_trace": [float(v) for v in bo.historical_optimal_values],
            "X": np.asarray(bo.X).tolist()}
