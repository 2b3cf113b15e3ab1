#!/usr/bin/env python
"""The other composite test problems the reference ships as scripts, end to end through the CBO loop on the CUDA path:

    2a  test_2a.py   d = 3, m = 4 GP-sample attributes on an 8^3 grid,  U = -sum_j exp(y_j)               uPI
    3a  test_3a.py   d = 2, m = 5 squared distances to five centres,     U = -sum_j c_j e^{-y_j/pi} cos(pi y_j)   uPI
    5a  test_5a.py   d = 5, m = 8 Rosenbrock pieces,                     U = -sum_j (a - y_j)^2 + 100 y_{j+d-1}^2  uEI_noiseless
    1b  test_1b.py   the test_1a attributes with a LINEAR utility theta . y                        maEI

Each builder follows its script: the objective h, the design space, `multi_outputGP(output_dim=m, exact_feval=[True] * m,
fixed_hyps=False)` (every iteration refits the hyper-parameters: ML-II + HMC, all outputs in lockstep on the device), the
parameter distribution, the utility (by the name of its device composite, with the script's Python callables kept for the
host-side bookkeeping and checked against it), the closed-form expectation where the script has one, and the acquisition
the script selects.  Two differences, both stated in DESIGN.md section 7: the scripts that ask for the CMA acquisition
optimiser get the L-BFGS multistart (the only one built), and `--fixed-hyps` replaces the inference by the deterministic
GPModelFixedHyps variant for quick runs.

    python examples/composite_problems.py --problem 5a [--iters 10] [--seed 0] [--fixed-hyps]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bocf_b200 as B  # noqa: E402


def _grid_gp_attributes(d, m, points_per_axis):
    """Attributes that are GP samples on a regular grid, served by their posterior means (test_1a.py:19-33,
    test_2a.py:18-41): SE kernel, variance 2, lengthscale 0.3, sample j drawn with RandomState(j + 7)."""
    axis = np.linspace(0., 1., points_per_axis)
    grid = np.array([a.flatten() for a in np.meshgrid(*([axis] * d))]).T
    Xs = grid / 0.3
    cov = 2. * np.exp(-0.5 * ((Xs[:, None, :] - Xs[None, :, :]) ** 2).sum(-1))
    n = grid.shape[0]
    Y = [np.random.RandomState(j + 7).multivariate_normal(np.zeros(n), cov).reshape(-1, 1) for j in range(m)]
    aux = B.multi_outputGP(m, kernel=[B.kern.SE(d, variance=2., lengthscale=0.3)] * m, noise_var=[1e-10] * m,
                           fixed_hyps=True)
    aux.updateModel(grid, Y)
    return lambda X: aux.posterior_mean(np.atleast_2d(X))


def problem_2a():
    d, m = 3, 4
    h = _grid_gp_attributes(d, m, 8)
    pdist = B.ParameterDistribution(continuous=False, support=np.ones((1,)), prob_dist=np.ones((1,)))
    U = B.Utility(func=lambda th, y: np.sum(-np.exp(y), axis=0), dfunc=lambda th, y: -np.exp(y), parameter_dist=pdist,
                  composite="neg_sum_exp")                                                   # test_2a.py:60-67
    expU = B.ExpectationUtility(                                                             # test_2a.py:70-83
        lambda th, mu, var: -np.sum(np.exp(np.squeeze(mu) + 0.5 * np.squeeze(var))),
        lambda th, mu, var: -np.concatenate((np.exp(np.squeeze(mu) + 0.5 * np.squeeze(var)),
                                             0.5 * np.exp(np.squeeze(mu) + 0.5 * np.squeeze(var)))))
    return dict(d=d, m=m, h=h, domain=(0, 1), utility=U, expectation=expU, acquisition="uPI")


def problem_3a():
    d, m = 2, 5
    A = np.array([[3., 5., 2., 1., 7.], [5., 2., 4., 1., 9.]])                               # test_3a.py:19

    def h(X):
        X = np.atleast_2d(X)
        return ((X[:, :, None] - A[None, :, :]) ** 2).sum(1).T                               # (m, N)

    c = np.array([1., 2., 5., 2., 3.])

    def U_func(th, y):                                                                       # test_3a.py:53-59
        y = np.squeeze(y)
        return -np.dot(c, np.exp(-y / np.pi) * np.cos(np.pi * y))

    def dU_func(th, y):                                                                      # test_3a.py:62-67
        y = np.squeeze(y)
        return -c * (-np.pi * np.exp(-y / np.pi) * np.sin(np.pi * y) - np.exp(-y / np.pi) * np.cos(np.pi * y) / np.pi)

    pdist = B.ParameterDistribution(continuous=False, support=np.ones((1,)), prob_dist=np.ones((1,)))
    U = B.Utility(func=U_func, dfunc=dU_func, parameter_dist=pdist, composite="exp_cos")
    return dict(d=d, m=m, h=h, domain=(0, 10), utility=U, expectation=None, acquisition="uPI")


def problem_5a():
    d = 5
    m = 2 * (d - 1)

    def h(X):                                                                                # test_5a.py:20-26
        X = np.atleast_2d(X)
        return np.concatenate((X[:, :d - 1].T, (X[:, 1:] - X[:, :d - 1] ** 2).T), axis=0)

    def U_func(a, y):                                                                        # test_5a.py:48-52
        return -np.sum((a - y[:d - 1]) ** 2 + 100. * y[d - 1:] ** 2, axis=0)

    def dU_func(a, y):                                                                       # test_5a.py:54-59
        return np.concatenate((2. * (a - y[:d - 1]), -200. * y[d - 1:]))

    pdist = B.ParameterDistribution(continuous=False, support=np.atleast_1d([1.]), prob_dist=np.ones((1,)))
    U = B.Utility(func=U_func, dfunc=dU_func, parameter_dist=pdist, composite="rosen_composite")

    def psi(a, mean, var):                                                                   # test_5a.py:64-68
        mean, var = np.squeeze(mean), np.squeeze(var)
        return -np.sum((a - mean[:d - 1]) ** 2 + 100. * mean[d - 1:] ** 2 + var[:d - 1] + 100. * var[d - 1:])

    def psi_gradient(a, mean, var):                                                          # test_5a.py:70-77
        mean = np.squeeze(mean)
        return np.concatenate((2. * (a - mean[:d - 1]), -200. * mean[d - 1:], -np.ones(d - 1), -100. * np.ones(d - 1)))

    return dict(d=d, m=m, h=h, domain=(-2, 2), utility=U, expectation=B.ExpectationUtility(psi, psi_gradient),
                acquisition="uEI_noiseless")


def problem_1b():
    d, m = 4, 5
    h = _grid_gp_attributes(d, m, 6)
    # a linear utility with a handful of equally likely weight vectors (test_1b.py:80-95: theta . y, linear=True)
    support = np.random.RandomState(3).dirichlet(np.ones(m), size=4)
    pdist = B.ParameterDistribution(continuous=False, support=support, prob_dist=np.full(4, 0.25))
    U = B.Utility(func=lambda th, y: np.dot(th, y), dfunc=lambda th, y: th, parameter_dist=pdist, linear=True,
                  composite="linear")
    return dict(d=d, m=m, h=h, domain=(0, 1), utility=U, expectation=None, acquisition="maEI")


PROBLEMS = {"2a": problem_2a, "3a": problem_3a, "5a": problem_5a, "1b": problem_1b}


def build(problem, seed=0, fixed_hyps=False, n_starting=400):
    np.random.seed(seed)
    p = PROBLEMS[problem]()
    d, m = p["d"], p["m"]
    objective = B.MultiObjective(p["h"], as_list=False, output_dim=m)
    space = B.Design_space(space=[{'name': 'var', 'type': 'continuous', 'domain': p["domain"], 'dimensionality': d}])
    model = B.multi_outputGP(output_dim=m, exact_feval=[True] * m, fixed_hyps=fixed_hyps)
    acq_opt = B.AcquisitionOptimizer(optimizer='lbfgs2', inner_optimizer='lbfgs2', space=space, n_starting=n_starting)
    X_init = B.initial_design('random', space, 2 * (d + 1))
    acquisition = getattr(B, p["acquisition"])(model, space, optimizer=acq_opt, utility=p["utility"])
    evaluator = B.Sequential(acquisition)
    return B.CBO(model, space, objective, acquisition, evaluator, X_init, expectation_utility=p["expectation"])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--problem", default="5a", choices=sorted(PROBLEMS))
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--fixed-hyps", action="store_true")
    a = ap.parse_args()
    bo = build(a.problem, a.seed, fixed_hyps=a.fixed_hyps)
    bo.run_optimization(max_iter=a.iters, verbosity=True)
    print("suggested points:\n", np.vstack(bo.suggested_points))
    print("best-value trace:", np.array(bo.historical_optimal_values))
