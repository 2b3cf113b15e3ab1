"""Knowledge-gradient posterior helpers (multi_outputGP.py:203-281,309-331 -> gp.py:493-627):
conditioned-on-next-point variance and gradient, posterior covariance between points and its gradient.

CPU: the oracle restatement against goldens produced by the reference's own classes (tests/golden/make_golden_kg.py).
GPU: the product (bocf_posterior_cov_point + rank-one Schur complement, bocf_b200/model.py) against the same goldens
and against the oracle on a larger ragged problem."""
import os

import numpy as np
import pytest

from tests.helpers import Problem, make_problem, oracle_model, product_model, rel_err, assert_close

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["kg_se", "kg_matern52"]


def _load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    P = Problem()
    P.X, P.variance, P.lengthscale, P.noise, P.Xc = z["X"], z["variance"], z["lengthscale"], z["noise"], z["Xc"]
    P.kind = str(z["kind"])
    P.H, P.m = P.variance.shape
    P.n, P.d = P.X.shape
    P.N = P.Xc.shape[0]
    P.Y = [z["Y"][j].reshape(-1, 1) for j in range(P.m)]
    return P, z


def _kg_outputs(model, P, z, h):
    model.set_hyperparameters(h)
    out = {}
    out["cov"] = model.posterior_covariance_between_points(P.Xc, z["X2"])
    model.partial_precomputation_for_covariance(z["X2"])
    out["cov_pp"] = model.posterior_covariance_between_points_partially_precomputed(P.Xc, z["X2"])
    model.partial_precomputation_for_variance_conditioned_on_next_point(z["x_next"])
    out["varcond"] = model.posterior_variance_conditioned_on_next_point(P.Xc)
    out["dvarcond"] = model.posterior_variance_gradient_conditioned_on_next_point(P.Xc)
    if "dcov_h0" in z.files:
        x2 = z["X2"][1:2]
        model.partial_precomputation_for_covariance_gradient(x2)
        out["dcov"] = np.concatenate([model.posterior_covariance_gradient_partially_precomputed(P.Xc[i:i + 1], x2)
                                      for i in range(P.N)], axis=1)
    return out


@pytest.mark.parametrize("name", CASES)
def test_oracle_kg_helpers_match_reference_goldens(name):
    P, z = _load(name)
    om = oracle_model(P)
    for h in range(P.H):
        o = _kg_outputs(om, P, z, h)
        assert rel_err(o["cov"], z["cov_h%d" % h]) < 1e-11 and rel_err(o["cov_pp"], z["cov_h%d" % h]) < 1e-10
        assert rel_err(o["varcond"], z["varcond_h%d" % h]) < 1e-10
        assert rel_err(o["dvarcond"], z["dvarcond_h%d" % h]) < 1e-9
        if "dcov" in o:
            assert rel_err(o["dcov"], z["dcov_h%d" % h]) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_product_kg_helpers_match_reference_goldens(cuda_device, name):
    P, z = _load(name)
    pm = product_model(P, cuda_device)
    for h in range(P.H):
        o = _kg_outputs(pm, P, z, h)
        assert rel_err(o["cov"], z["cov_h%d" % h]) < 1e-8 and np.array_equal(o["cov"], o["cov_pp"])
        # the conditional variance is a difference of O(1) numbers that can reach 1e-8 at the next point itself: absolute bar
        assert np.max(np.abs(o["varcond"] - z["varcond_h%d" % h])) < 2e-7 * float(P.variance.max())
        assert rel_err(o["dvarcond"], z["dvarcond_h%d" % h]) < 2e-6
        if "dcov" in o:
            assert rel_err(o["dcov"], z["dcov_h%d" % h]) < 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["se", "rbf", "matern52", "matern32"])
def test_product_kg_helpers_match_oracle(cuda_device, kind):
    P = make_problem(m=3, d=5, n=131, H=2, kind=kind, N=300, S=4, seed=31)
    om, pm = oracle_model(P), product_model(P, cuda_device)
    X2, x_next = P.Xc[:2], P.X[7:8] + 0.01
    for h in range(P.H):
        om.set_hyperparameters(h)
        pm.set_hyperparameters(h)
        assert_close(pm.posterior_covariance_between_points(P.Xc, X2), om.posterior_covariance_between_points(P.Xc, X2),
                     1e-7, "cov")
        for mod in (om, pm):
            mod.partial_precomputation_for_variance_conditioned_on_next_point(x_next)
        v, v_o = pm.posterior_variance_conditioned_on_next_point(P.Xc), om.posterior_variance_conditioned_on_next_point(P.Xc)
        assert np.max(np.abs(v - v_o)) < 2e-7 * float(P.variance.max())
        assert rel_err(pm.posterior_variance_gradient_conditioned_on_next_point(P.Xc),
                       om.posterior_variance_gradient_conditioned_on_next_point(P.Xc)) < 2e-6
        if kind == "se":      # gradients_X(None, ...) exists only for the fork's SE kernel (se.py:142-144)
            g = pm.posterior_covariance_gradient(P.Xc, X2[0])
            g_o = np.concatenate([om.posterior_covariance_gradient(P.Xc[i:i + 1], X2[0:1]) for i in range(20)], axis=1)
            assert rel_err(g[:, :20], g_o) < 1e-8
