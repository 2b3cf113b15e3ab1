"""Appending one observation with unchanged hyper-parameters: the O(n^2) bordered update of
L, L^-1 and alpha must give the model a full refactorisation gives -- in both contraction modes -- and must fall back
cleanly when the factor buffers are full or the hyper-parameters changed."""
import numpy as np
import pytest

from tests.helpers import make_problem, oracle_model, product_model, rel_err

pytestmark = pytest.mark.gpu


def _sub(P, n):
    import copy
    Q = copy.copy(P)
    Q.X = P.X[:n]
    Q.Y = [y[:n] for y in P.Y]
    Q.n = n
    return Q


@pytest.mark.parametrize("kind", ["se", "matern52"])
def test_append_equals_refactorisation(cuda_device, kind):
    P = make_problem(m=3, d=4, n=110, H=2, kind=kind, N=300, S=8, seed=6)
    pm = product_model(_sub(P, 100), cuda_device)
    assert pm.last_update == "factorize"
    for n in range(101, 111):
        pm.updateModel(P.X[:n], [y[:n] for y in P.Y])
        assert pm.last_update == "append"
    ref = product_model(P, cuda_device)                      # one factorisation of all 110 points
    om = oracle_model(P)
    for h in range(P.H):
        for j in range(P.m):
            L, Li, al = pm.get_factor(h, j)
            L2, Li2, al2 = ref.get_factor(h, j)
            assert rel_err(L, L2) < 1e-10 and rel_err(Li, Li2) < 1e-9 and rel_err(al, al2) < 1e-8
            assert rel_err(L, om.output[j].model_instances[h].woodbury_chol) < 1e-10
        pm.set_hyperparameters(h)
        om.set_hyperparameters(h)
        assert rel_err(pm.posterior_mean(P.Xc), om.posterior_mean(P.Xc)) < 1e-8
        v, v_o = pm.posterior_variance(P.Xc), om.posterior_variance(P.Xc)
        assert np.max(np.abs(v - v_o) / np.abs(v_o)) < 1e-6
        assert rel_err(pm.posterior_variance_gradient(P.Xc), om.posterior_variance_gradient(P.Xc)) < 1e-6
    assert rel_err(pm.log_likelihood(), ref.log_likelihood()) < 1e-10


def test_append_falls_back_when_buffers_are_full_or_hypers_change(cuda_device):
    P = make_problem(m=2, d=3, n=130, H=1, kind="rbf", N=64, S=8, seed=8)
    pm = product_model(_sub(P, 127), cuda_device)
    pm.updateModel(P.X[:128], [y[:128] for y in P.Y])
    assert pm.last_update == "append"                        # 128 fits the 128-row factor buffers
    pm.updateModel(P.X[:129], [y[:129] for y in P.Y])
    assert pm.last_update == "factorize"                     # buffers full: transparent refactorisation
    om = oracle_model(_sub(P, 129))
    assert rel_err(pm.posterior_mean(P.Xc), om.posterior_mean(P.Xc)) < 1e-8
    v, v_o = pm.posterior_variance(P.Xc), om.posterior_variance(P.Xc)
    assert np.max(np.abs(v - v_o) / np.abs(v_o)) < 1e-6
    # new hyper-parameters refactorise at once (set_hyperparameter_samples); the next point is incremental again and
    # the result must match an oracle model built from scratch with the new hyper-parameters
    pm.set_hyperparameter_samples(P.variance * 1.1, P.lengthscale, P.noise, kind=P.kind)
    assert pm.last_update == "factorize"
    pm.updateModel(P.X[:130], [y[:130] for y in P.Y])
    assert pm.last_update == "append"
    P.variance = P.variance * 1.1
    om = oracle_model(P)
    assert rel_err(pm.posterior_mean(P.Xc), om.posterior_mean(P.Xc)) < 1e-8
    v, v_o = pm.posterior_variance(P.Xc), om.posterior_variance(P.Xc)
    assert np.max(np.abs(v - v_o) / np.abs(v_o)) < 1e-6
    # switching the incremental path off forces the refactorisation
    pm2 = product_model(_sub(P, 100), cuda_device)
    pm2.incremental_updates = False
    pm2.updateModel(P.X[:101], [y[:101] for y in P.Y])
    assert pm2.last_update == "factorize"
