"""A kernel family PER OUTPUT (multi_outputGP.py:23,38-44: output j is built from kernel[j]).

The family is a compile-time parameter of every device kernel that evaluates k(.,.), so the library issues one launch
per run of consecutive outputs with the same family (bocf_model_set_kernels).  Every consumer of the family is checked
against the oracle on a model whose outputs mix all four: Gram / factor, posterior and its gradients, EI-CF with
gradient, the likelihood and its gradients, the one-point append and the KG covariance helper."""
import numpy as np
import pytest

from tests.helpers import (assert_close, tol, make_problem, oracle_model, oracle_acq, product_model, product_acq, rel_err)

pytestmark = pytest.mark.gpu

# runs of length 1, 2 and 1 plus a family that comes back later (non-contiguous): 5 launches for 6 outputs
KINDS = ("matern52", "se", "se", "rbf", "matern32", "matern52")


def _problem(**kw):
    args = dict(m=len(KINDS), d=5, n=90, H=2, kind=KINDS, N=300, S=64, seed=21)
    args.update(kw)
    return make_problem(**args)


def test_factor_and_posterior_with_a_kernel_list(cuda_device):
    P = _problem()
    om, pm = oracle_model(P), product_model(P, cuda_device)
    for h in range(P.H):
        om.set_hyperparameters(h)
        pm.set_hyperparameters(h)
        for j in range(P.m):
            gp = om.output[j].model_instances[h]
            L, Linv, alpha = pm.get_factor(h, j)
            assert rel_err(L, gp.woodbury_chol) < 1e-10, (h, j, KINDS[j])
            assert rel_err(alpha, gp.woodbury_vector[:, 0]) < 1e-8
        mu, v = pm.posterior_mean(P.Xc), pm.posterior_variance(P.Xc)
        dm, dv = pm.posterior_mean_gradient(P.Xc), pm.posterior_variance_gradient(P.Xc)
        mu_o, v_o = om.posterior_mean(P.Xc), om.posterior_variance(P.Xc)
        dm_o, dv_o = om.posterior_mean_gradient(P.Xc), om.posterior_variance_gradient(P.Xc)
        for j in range(P.m):                                   # per output: a wrong family on one output must not hide
            assert rel_err(mu[j], mu_o[j]) < tol(1e-8), (j, KINDS[j])
            assert np.max(np.abs(v[j] - v_o[j]) / np.abs(v_o[j])) < tol(1e-7), (j, KINDS[j])
            assert_close(dm[j], dm_o[j], tol(1e-8), "dmean[%d]" % j)
            assert_close(dv[j], dv_o[j], tol(1e-7), "dvar[%d]" % j)


def test_kernel_list_differs_from_single_family(cuda_device):
    # guards against a silently ignored list: the mixed model must not equal the all-matern52 one
    P = _problem(H=1, N=64)
    pm = product_model(P, cuda_device)
    P1 = _problem(H=1, N=64)
    P1.kind = "matern52"
    p1 = product_model(P1, cuda_device)
    mu, mu1 = pm.posterior_mean(P.Xc), p1.posterior_mean(P.Xc)
    assert rel_err(mu[0], mu1[0]) < 1e-12 and rel_err(mu[5], mu1[5]) < 1e-12      # the matern52 outputs agree
    assert rel_err(mu[1], mu1[1]) > 1e-4 and rel_err(mu[3], mu1[3]) > 1e-4        # the others do not


def test_eicf_with_a_kernel_list(cuda_device):
    P = _problem(composite="sumsq_target", L=2)
    a_o, g_o = oracle_acq(P, grad=True)
    a, g = product_acq(P, grad=True, device=cuda_device)
    assert np.mean(a_o > 0) > 0.02, "degenerate test problem"
    assert_close(a, a_o, tol(1e-8), "acq")
    assert_close(g, g_o, tol(1e-7), "grad acq")
    assert np.argmax(a) == np.argmax(a_o)


def test_small_batch_path_with_a_kernel_list(cuda_device):
    # <= 1024 candidates take the K-split K* launches (posterior.cu: kstar_ksplit) -- one per run as well
    P = _problem(H=1, N=17, n=300)
    om, pm = oracle_model(P), product_model(P, cuda_device)
    om.set_hyperparameters(0)
    assert_close(pm.posterior_mean(P.Xc), om.posterior_mean(P.Xc), tol(1e-8), "mean")
    assert_close(pm.posterior_mean_gradient(P.Xc), om.posterior_mean_gradient(P.Xc), tol(1e-8), "dmean")
    assert_close(pm.posterior_variance_gradient(P.Xc), om.posterior_variance_gradient(P.Xc), tol(1e-7), "dvar")


def test_likelihood_with_a_kernel_list(cuda_device):
    P = _problem()
    pm, om = product_model(P, cuda_device), oracle_model(P)
    lml, gv, gl, gn = pm.log_likelihood_and_gradients()
    for h in range(P.H):
        for j in range(P.m):
            g = om.output[j].model_instances[h]
            o_gv, o_gl, o_gn = g.likelihood_gradients()
            o_l = g.log_likelihood()
            assert abs(lml[h, j] - o_l) < 1e-10 * max(1.0, abs(o_l)), (h, j, KINDS[j])
            assert abs(gv[h, j] - o_gv) < 1e-8 * max(1.0, abs(o_gv)) and abs(gn[h, j] - o_gn) < 1e-8 * max(1.0, abs(o_gn))
            assert rel_err(gl[h, j], o_gl) < 1e-8, (h, j, KINDS[j])


def test_append_and_kg_helper_with_a_kernel_list(cuda_device):
    P = _problem(H=1, N=40)
    n0 = P.n - 1
    pm = product_model(P, cuda_device)                                  # reference: factorised on all n points
    import bocf_b200
    inc = bocf_b200.multi_outputGP(P.m, n_samples=P.H, device=cuda_device)
    inc.set_hyperparameter_samples(P.variance, P.lengthscale, P.noise, kind=P.kind)
    inc.updateModel(P.X[:n0], [y[:n0] for y in P.Y])
    inc.updateModel(P.X, P.Y)                                           # one new point: O(n^2) bordered update
    assert inc.last_update == "append"
    for j in range(P.m):
        La, _, aa = inc.get_factor(0, j)
        Lf, _, af = pm.get_factor(0, j)
        assert rel_err(La, Lf) < 1e-9 and rel_err(aa, af) < 1e-7, (j, KINDS[j])
    om = oracle_model(P)
    om.set_hyperparameters(0)
    pm.set_hyperparameters(0)
    X2 = P.Xc[:3]
    c, c_o = pm.posterior_covariance_between_points(P.Xc[3:], X2), om.posterior_covariance_between_points(P.Xc[3:], X2)
    assert np.shape(c) == np.shape(c_o)
    for j in range(P.m):
        assert_close(np.asarray(c)[j], np.asarray(c_o)[j], tol(1e-7), "cov[%d]" % j)
    # the gradient variant of the covariance kernel, through the conditioned-variance helpers (gp.py:493-560; the
    # reference's covariance-gradient helper itself only works for its SE kernel: gradients_X(None, ...) elsewhere)
    x_next = P.Xc[1:2]
    pm.partial_precomputation_for_variance_conditioned_on_next_point(x_next)
    om.partial_precomputation_for_variance_conditioned_on_next_point(x_next)
    vc, vc_o = pm.posterior_variance_conditioned_on_next_point(P.Xc[3:]), om.posterior_variance_conditioned_on_next_point(P.Xc[3:])
    dvc = pm.posterior_variance_gradient_conditioned_on_next_point(P.Xc[3:])
    dvc_o = om.posterior_variance_gradient_conditioned_on_next_point(P.Xc[3:])
    for j in range(P.m):
        assert_close(np.asarray(vc)[j], np.asarray(vc_o)[j], tol(1e-6), "varcond[%d]" % j)
        assert_close(np.asarray(dvc)[j], np.asarray(dvc_o)[j], tol(1e-6), "dvarcond[%d]" % j)


def test_kernel_list_across_output_groups(cuda_device):
    # n = 300 (three 128-blocks), H m = 12: the factorisation and the likelihood pass run as two output groups on two
    # streams (for_each_output_group, chol.cu), each group cutting the family runs at its own border
    P = _problem(n=300, N=64)
    om, pm = oracle_model(P), product_model(P, cuda_device)
    lml, gv, gl, gn = pm.log_likelihood_and_gradients()
    for h in range(P.H):
        om.set_hyperparameters(h)
        pm.set_hyperparameters(h)
        for j in range(P.m):
            gp = om.output[j].model_instances[h]
            L, Linv, alpha = pm.get_factor(h, j)
            assert rel_err(L, gp.woodbury_chol) < 1e-10, (h, j, KINDS[j])
            assert rel_err(Linv @ gp.woodbury_chol, np.eye(P.n)) < 1e-9
            assert rel_err(alpha, gp.woodbury_vector[:, 0]) < 1e-8
            o_gv, o_gl, o_gn = gp.likelihood_gradients()
            assert abs(lml[h, j] - gp.log_likelihood()) < 1e-10 * max(1.0, abs(gp.log_likelihood()))
            assert rel_err(gl[h, j], o_gl) < 1e-8 and abs(gv[h, j] - o_gv) < 1e-8 * max(1.0, abs(o_gv))
        assert_close(pm.posterior_mean(P.Xc), om.posterior_mean(P.Xc), tol(1e-8), "mean")
        v, v_o = pm.posterior_variance(P.Xc), om.posterior_variance(P.Xc)
        assert np.max(np.abs(v - v_o) / np.abs(v_o)) < tol(1e-7)
