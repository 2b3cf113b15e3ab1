#!/usr/bin/env python
"""BASELINE.json configs[0]: the small composite test problem of the reference's test_1a.py, end to end through
the CBO loop with EI-CF (uEI_noiseless, selected at test_1a.py:151), on the CUDA path.

Problem (test_1a.py:19-33): d = 4, m = 5 attributes, each a GP sample on the 6^4 grid (SE kernel, variance 2,
lengthscale 0.3, seeds j+7); the objective is the posterior mean of those GPs; utility -sum_j (y_j - theta_j)^2
with theta = f(x_opt of attribute 0) (:60-96); model = multi_outputGP(output_dim=m, exact_feval=[False]*m,
fixed_hyps=False) as at test_1a.py:53, i.e. every iteration refits the hyper-parameters (ML-II) and draws 10 HMC
hyper-samples per output -- all outputs in lockstep on the device likelihood.  `--fixed-hyps` switches to the
deterministic GPModelFixedHyps variant.  The two dead imports of the script (uKG_SGA, uKG_cf) are dropped.

    python examples/test_1a_cf.py [--iters 10] [--seed 0] [--fixed-hyps]
"""
import argparse
import os
import sys

import numpy as np
import scipy.optimize

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bocf_b200 as B  # noqa: E402


def build(seed=0, n_starting=400, fixed_hyps=False):
    np.random.seed(seed)
    d, m = 4, 5
    I = np.linspace(0., 1., 6)
    grid = np.array([a.flatten() for a in np.meshgrid(I, I, I, I)]).T
    # GP samples on the grid (test_1a.py:24-33); the aux models are one device multi-output GP
    Xs = grid / 0.3
    cov = 2. * np.exp(-0.5 * ((Xs[:, None, :] - Xs[None, :, :]) ** 2).sum(-1))
    Ygrid = [np.random.RandomState(j + 7).multivariate_normal(np.zeros(6 ** d), cov).reshape(-1, 1) for j in range(m)]
    aux = B.multi_outputGP(m, kernel=[B.kern.SE(d, variance=2., lengthscale=0.3)] * m, noise_var=[1e-10] * m,
                           fixed_hyps=True)
    aux.updateModel(grid, Ygrid)

    def f(X):
        return aux.posterior_mean(np.atleast_2d(X))

    objective = B.MultiObjective(f, as_list=False, output_dim=m)
    space = B.Design_space(space=[{'name': 'var', 'type': 'continuous', 'domain': (0, 1), 'dimensionality': d}])
    model = B.multi_outputGP(output_dim=m, exact_feval=[False] * m, fixed_hyps=fixed_hyps)
    acq_opt = B.AcquisitionOptimizer(optimizer='lbfgs2', inner_optimizer='lbfgs2', space=space, n_starting=n_starting)
    X_init = B.initial_design('random', space, 2 * (d + 1))
    # theta: attribute values at the maximiser of attribute 0 (test_1a.py:60-81)
    best = (np.inf, None)
    for x0 in np.random.rand(20, d):
        res = scipy.optimize.fmin_l_bfgs_b(lambda x: -f(x)[0, 0], x0, approx_grad=True, bounds=[(0, 1)] * d)
        if res[1] < best[0]:
            best = (res[1], res[0])
    parameter_support = f(best[1]).T
    pdist = B.ParameterDistribution(continuous=False, support=parameter_support, prob_dist=np.ones((1,)))
    U = B.Utility(parameter_dist=pdist, composite="sumsq_target")
    expU = B.ExpectationUtility(
        lambda th, mu, var: -np.sum(np.square((mu.T - th).T), axis=0) - np.sum(var, axis=0),
        lambda th, mu, var: -np.concatenate((2 * (np.squeeze(mu) - th), np.ones((len(np.squeeze(var)),)))))
    acquisition = B.uEI_noiseless(model, space, optimizer=acq_opt, utility=U)
    evaluator = B.Sequential(acquisition)
    return B.CBO(model, space, objective, acquisition, evaluator, X_init, expectation_utility=expU)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--fixed-hyps", action="store_true")
    a = ap.parse_args()
    bo = build(a.seed, fixed_hyps=a.fixed_hyps)
    bo.run_optimization(max_iter=a.iters, verbosity=True)
    print("suggested points:\n", np.vstack(bo.suggested_points))
    print("best-value trace:", np.array(bo.historical_optimal_values))
