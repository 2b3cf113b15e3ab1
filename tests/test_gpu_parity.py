"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on identical
seeded inputs and identical base samples Z.

Tolerances (BASELINE.json north_star): 1e-6 relative on posterior mean / variance in fp64 mode
(the kernels are fp64 end to end, so the gradients and the acquisition are held to the same order);
the EI-CF value / gradient bar of 1e-4 is the mixed-precision bar and is trivially implied here.
"""
import numpy as np
import pytest

from tests.helpers import (assert_close, tol, make_problem, oracle_model, oracle_acq, product_model, product_acq, product_utility,
                           rel_err)

pytestmark = pytest.mark.gpu

TOL_MEANVAR = 1e-6      # north-star fp64 bar on mean and variance
TOL_TIGHT = 1e-8        # what fp64 kernels should actually reach on well-conditioned problems


@pytest.mark.parametrize("kind", ["se", "rbf", "matern52", "matern32"])
@pytest.mark.parametrize("n", [37, 200])
def test_factorisation_matches_lapack(cuda_device, kind, n):
    P = make_problem(m=2, d=5, n=n, H=2, kind=kind, N=8, S=4, seed=n)
    om = oracle_model(P)
    pm = product_model(P, cuda_device)
    assert np.all(pm.jitter_added == 0.0)
    for h in range(P.H):
        for j in range(P.m):
            gp = om.output[j].model_instances[h]
            L, Linv, alpha = pm.get_factor(h, j)
            assert rel_err(L, gp.woodbury_chol) < 1e-10
            assert rel_err(alpha, gp.woodbury_vector[:, 0]) < 1e-8
            assert rel_err(Linv @ gp.woodbury_chol, np.eye(n)) < 1e-9
            assert np.allclose(np.triu(L, 1), 0.0) and np.allclose(np.triu(Linv, 1), 0.0)


@pytest.mark.parametrize("kind", ["se", "rbf", "matern52", "matern32"])
@pytest.mark.parametrize("shape", [(3, 4, 60, 300), (2, 10, 257, 129), (5, 1, 16, 1), (1, 16, 130, 515)])
def test_posterior_matches_oracle(cuda_device, kind, shape):
    m, d, n, N = shape
    P = make_problem(m=m, d=d, n=n, H=2, kind=kind, N=N, S=4, seed=11 * m + d)
    om = oracle_model(P)
    pm = product_model(P, cuda_device)
    for h in range(P.H):
        om.set_hyperparameters(h)
        pm.set_hyperparameters(h)
        mu_o, v_o = om.posterior_mean(P.Xc), om.posterior_variance(P.Xc)
        dm_o, dv_o = om.posterior_mean_gradient(P.Xc), om.posterior_variance_gradient(P.Xc)
        mu, v = pm.posterior_mean(P.Xc), pm.posterior_variance(P.Xc)
        dm, dv = pm.posterior_mean_gradient(P.Xc), pm.posterior_variance_gradient(P.Xc)
        assert mu.shape == (m, N) and v.shape == (m, N) and dm.shape == (m, N, d) and dv.shape == (m, N, d)
        assert rel_err(mu, mu_o) < tol(TOL_MEANVAR) and rel_err(v, v_o) < tol(TOL_MEANVAR)
        assert rel_err(mu, mu_o) < tol(TOL_TIGHT) and np.max(np.abs(v - v_o) / np.abs(v_o)) < tol(1e-7)
        assert rel_err(dm, dm_o) < tol(TOL_TIGHT) and rel_err(dv, dv_o) < tol(1e-7)
        # element-wise (entries above 1 % of the largest: relative; below: 100x tighter than the norm-wise bar)
        assert_close(dm, dm_o, tol(TOL_TIGHT), "dmean")
        assert_close(dv, dv_o, tol(1e-7), "dvar")
        mp, vp = pm.predict(P.Xc)
        mo, vo = om.predict(P.Xc)
        assert rel_err(mp, mo) < tol(TOL_TIGHT) and rel_err(vp, vo) < tol(1e-7)
        vn = pm.posterior_variance_noiseless(P.Xc)
        assert rel_err(vn, om.posterior_variance_noiseless(P.Xc)) < tol(1e-6)


def test_posterior_at_training_points_and_clip(cuda_device):
    # near-noiseless model: variance at the training inputs collapses and must be clipped at 1e-10 (gpmodel.py:174)
    P = make_problem(m=2, d=3, n=40, H=1, kind="se", N=16, S=4, noise=1e-10, seed=5)
    om = oracle_model(P)
    pm = product_model(P, cuda_device)
    v = pm.posterior_variance(P.X)
    vo = om.posterior_variance(P.X)
    assert np.all(v >= 1e-10) and np.all(vo >= 1e-10)
    assert np.max(np.abs(v - vo)) < tol(1e-7)
    f = pm.posterior_mean_at_evaluated_points()
    assert rel_err(f, om.posterior_mean_at_evaluated_points()) < tol(1e-7)


@pytest.mark.parametrize("composite", ["sumsq_target", "neg_sum_exp", "exp_cos", "rosen_composite", "linear"])
@pytest.mark.parametrize("kind", ["se", "matern52"])
def test_eicf_value_and_gradient(cuda_device, composite, kind):
    P = make_problem(m=4, d=6, n=150, H=2, kind=kind, composite=composite, N=400, S=96, L=2, seed=2)
    a_o, g_o = oracle_acq(P, grad=True)
    a, g = product_acq(P, grad=True, device=cuda_device)
    assert np.mean(a_o > 0) > 0.02, "degenerate test problem"
    assert rel_err(a, a_o) < tol(1e-8), rel_err(a, a_o)
    assert rel_err(g, g_o) < tol(1e-7), rel_err(g, g_o)
    assert_close(a, a_o, tol(1e-8), "acq")                     # element-wise as well (tests/helpers.py: elem_err)
    assert_close(g, g_o, tol(1e-7), "grad acq")
    assert np.argmax(a) == np.argmax(a_o)                      # same selected candidate
    av, _ = product_acq(P, grad=False, device=cuda_device)
    avo, _ = oracle_acq(P, grad=False)
    assert rel_err(av, avo) < tol(1e-8)


def test_eicf_matches_literal_reference_loops(cuda_device):
    # the oracle's literal triple loop (uEI_noiseless.py:138-170 line for line), small enough to finish in seconds
    P = make_problem(m=3, d=4, n=50, H=2, kind="rbf", composite="sumsq_target", N=40, S=25, L=1, seed=4)
    a_o, g_o = oracle_acq(P, grad=True, vectorised=False)
    a, g = product_acq(P, grad=True, device=cuda_device)
    assert rel_err(a, a_o) < tol(1e-9) and rel_err(g, g_o) < tol(1e-8)


@pytest.mark.parametrize("S", [1, 25, 1024, 1500])
def test_eicf_sample_counts(cuda_device, S):
    P = make_problem(m=3, d=5, n=64, H=1, kind="matern32", composite="neg_sum_exp", N=130, S=S, seed=S)
    a_o, g_o = oracle_acq(P, grad=True)
    a, g = product_acq(P, grad=True, device=cuda_device)
    assert rel_err(a, a_o) < tol(1e-8) and rel_err(g, g_o) < tol(1e-7)


def test_upi_value_only(cuda_device):
    P = make_problem(m=4, d=6, n=120, H=2, kind="se", composite="sumsq_target", N=300, S=128, L=3, seed=8)
    a_o, _ = oracle_acq(P, grad=False, variant="uPI")
    a, _ = product_acq(P, grad=False, variant="uPI", device=cuda_device)
    assert np.mean(a_o > 0) > 0.02
    assert np.max(np.abs(a - a_o)) < tol(1e-12)         # counts / (H*S): exact up to the final scaling
    with pytest.raises(NotImplementedError):
        product_acq(P, grad=True, variant="uPI", device=cuda_device)


@pytest.mark.parametrize("variant", ["maEI", "maPI"])
def test_analytic_variants(cuda_device, variant):
    P = make_problem(m=5, d=4, n=90, H=3, kind="matern52", composite="linear", N=260, S=4, L=4, seed=6)
    a_o, g_o = oracle_acq(P, grad=True, variant=variant)
    a, g = product_acq(P, grad=True, variant=variant, device=cuda_device)
    assert rel_err(a, a_o) < tol(1e-8) and rel_err(g, g_o) < tol(1e-7)
    assert_close(a, a_o, tol(1e-8), "acq")
    assert_close(g, g_o, tol(1e-7), "grad acq")
    av_o, _ = oracle_acq(P, grad=False, variant=variant)
    av, _ = product_acq(P, grad=False, variant=variant, device=cuda_device)
    assert rel_err(av, av_o) < tol(1e-8)


@pytest.mark.parametrize("variant", ["EI", "PI"])
def test_single_output_ei_pi(cuda_device, variant):
    P = make_problem(m=1, d=3, n=30, H=1, kind="se", composite="linear", N=70, S=4, L=1, seed=9)
    P.theta = np.ones((1, 1))
    a_o, g_o = oracle_acq(P, grad=True, variant=variant)
    a, g = product_acq(P, grad=True, variant=variant, device=cuda_device)
    assert rel_err(a, a_o) < tol(1e-8) and rel_err(g, g_o) < tol(1e-7)


def test_torch_tensors_stay_on_device(cuda_device):
    import torch
    P = make_problem(m=2, d=3, n=33, H=1, kind="rbf", N=50, S=8, seed=1)
    pm = product_model(P, cuda_device)
    Xd = torch.from_numpy(P.Xc).to(cuda_device)
    mu = pm.posterior_mean(Xd)
    assert isinstance(mu, torch.Tensor) and mu.is_cuda and mu.shape == (2, 50)
    assert rel_err(mu.cpu().numpy(), pm.posterior_mean(P.Xc)) == 0.0


def test_jitchol_retry_and_failure(cuda_device):
    import bocf_b200
    # duplicated inputs, zero noise and a huge signal variance: the 1e-8 ridge is absorbed (1e9 + 1e-8 == 1e9), the
    # 40 duplicate pivots are pure rounding noise, so dpotrf fails here and on the device alike and jitchol
    # (linalg.py:52-83: mean(diag)*1e-6 first, x10 per retry) has to add 1e3
    rng = np.random.default_rng(0)
    X = np.repeat(rng.uniform(size=(20, 2)), 3, axis=0)
    Y = [rng.standard_normal((60, 1))]
    mod = bocf_b200.multi_outputGP(1, device=cuda_device)
    mod.set_hyperparameter_samples(np.full((1, 1), 1e9), np.full((1, 1, 2), 50.0), np.zeros((1, 1)), kind="rbf")
    mod.updateModel(X, Y)
    from oracle.linalg import jitchol
    from oracle.kern import Kern
    K = Kern("rbf", 2, 1e9, [50.0, 50.0], ARD=True).K(X)
    K[np.diag_indices_from(K)] += 1e-8
    _, jit_o = jitchol(K, return_jitter=True)
    assert jit_o > 0
    assert mod.jitter_added[0, 0] == pytest.approx(jit_o, rel=1e-12)


def test_topk_matches_argsort(cuda_device):
    import ctypes
    import torch
    from bocf_b200 import _lib
    lib = _lib.load_library()
    rng = np.random.default_rng(3)
    N, d, k = 100003, 5, 16
    acq = rng.standard_normal(N)
    acq[rng.integers(0, N, 50)] = acq.max()          # ties -> smaller index first
    X = rng.uniform(size=(N, d))
    a_d = torch.from_numpy(acq).to(cuda_device)
    x_d = torch.from_numpy(X).to(cuda_device)
    rec = torch.empty((k, 2 + d), dtype=torch.float64, device=cuda_device)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.bocf_topk(ctypes.c_void_p(a_d.data_ptr()), ctypes.c_void_p(x_d.data_ptr()), N, d, k, 1000,
                             ctypes.c_void_p(rec.data_ptr()), st))
    rec = rec.cpu().numpy()
    order = np.lexsort((np.arange(N), -acq))[:k]
    assert np.array_equal(rec[:, 1].astype(np.int64), order + 1000)
    assert np.array_equal(rec[:, 0], acq[order]) and np.array_equal(rec[:, 2:], X[order])


def test_acq_eval_host_entry(cuda_device):
    """The host-buffer C-ABI entry (what bench.py times as e2e) gives the same numbers as the device entry."""
    import ctypes
    import torch
    from bocf_b200 import _lib
    P = make_problem(m=3, d=4, n=48, H=1, kind="matern52", composite="sumsq_target", N=333, S=32, seed=12)
    pm = product_model(P, cuda_device)
    a, g = product_acq(P, grad=True, device=cuda_device, model=pm)
    import bocf_b200
    acq_obj = bocf_b200.uEI_noiseless(pm, None, utility=product_utility(P))
    acq_obj.W_samples = P.Z
    pm.set_hyperparameters(0)
    fstar = acq_obj._fstar(P.theta)
    Zt = acq_obj._Zt()
    out_a = np.empty(P.N)
    out_g = np.empty((P.N, P.d))
    th = np.ascontiguousarray(P.theta)
    w = np.ascontiguousarray(P.prob)
    fs = np.ascontiguousarray(fstar)
    vp = lambda arr: arr.ctypes.data_as(ctypes.c_void_p)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    Xh = np.ascontiguousarray(P.Xc)
    _lib.check(pm._lib.bocf_acq_eval_host(pm._handle, 0, 0, vp(Xh), P.N, ctypes.c_void_p(Zt.data_ptr()), P.S, vp(th), 1,
                                          P.m, vp(w), vp(fs), 1, 0, vp(out_a), vp(out_g), st))
    assert np.array_equal(out_a, a) and np.array_equal(out_g, g)


def test_negated_and_pipelined_host_path(cuda_device):
    """acquisition_function[_withGradients] flip the sign on the device and, for long numpy inputs, stream the candidates
    through in slabs (H2D / D2H on a side stream).  The sign flip is exact; slabs of <= 1024 candidates take the K-split
    form of the K* kernel (posterior.cu), whose mean / mean-gradient sums run in a different order: equal to 1e-13."""
    import torch
    import bocf_b200
    from tests.helpers import make_problem, product_model, product_utility
    P = make_problem(m=3, d=5, n=70, H=2, kind="matern52", composite="sumsq_target", N=1500, S=64, seed=21)
    model = product_model(P, cuda_device)
    acq = bocf_b200.uEI_noiseless(model, None, utility=product_utility(P))
    acq.W_samples = P.Z
    model.set_hyperparameters(0)
    a0, g0 = acq._compute_acq_withGradients(P.Xc)                       # one shot, no sign
    model.set_hyperparameters(0)
    a1, g1 = acq.acquisition_function_withGradients(P.Xc)                # sign on device
    assert np.array_equal(a1, -a0) and np.array_equal(g1, -g0)
    acq.PIPELINE_MIN = 256                                               # force the slab pipeline (4 ragged slabs of 384)
    model.set_hyperparameters(0)
    a2, g2 = acq.acquisition_function_withGradients(P.Xc)
    assert a2.shape == (1500, 1) and g2.shape == (1500, 5)
    assert rel_err(a2, a1) < 1e-13 and rel_err(g2, g1) < 1e-13
    model.set_hyperparameters(0)
    v2 = acq.acquisition_function(P.Xc)
    model.set_hyperparameters(0)
    acq.PIPELINE_MIN = 1 << 30
    v1 = acq.acquisition_function(P.Xc)
    assert rel_err(v2, v1) < 1e-13 and np.all(v1 <= 0)
    # device tensors in -> device tensors out, same values
    model.set_hyperparameters(0)
    at, gt = acq.acquisition_function_withGradients(torch.from_numpy(P.Xc).to(cuda_device))
    assert at.is_cuda and np.array_equal(at.cpu().numpy().reshape(-1, 1), a1) and np.array_equal(gt.cpu().numpy(), g1)
    assert acq._sign == 1.0
