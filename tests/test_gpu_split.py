"""GPU tests of the tcgen05 split-integer contraction mode (bocf_b200/csrc/split_gemm.cu).

The two candidate-side contractions of the posterior variance (posterior.py:312, gp.py:474) run on
`tcgen05.mma.kind::i8` with the fp64 operands split into signed 8-bit digit planes and EXACT int32
accumulation in tensor memory.  A scheme (SA, SB, LMIN) keeps the digit pairs with ta + tb >= LMIN; "splitN" selects
331 / 442 / 554 / 665.  Bars (BASELINE.json north_star):
  * "auto" (554 for the variance, 442 for its gradient on well-conditioned models) and "split5" must meet the
    FP64-mode bar: 1e-6 relative on posterior mean and variance, element-wise -- and the same selected next point;
  * "mixed" (442 / 331) must meet the mixed-precision bar: 1e-4 relative on the EI-CF value and gradient.
The integer GEMM itself must be exact whenever the operands are exactly representable.
"""
import ctypes

import numpy as np
import pytest

from tests.helpers import make_problem, oracle_model, oracle_acq, product_model, product_acq, rel_err, assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture
def cuda_device(built_library):      # every test here names its precision explicitly: no mode parametrisation
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return "cuda:0"


def _split_gemm(A, B, S, tri=0):
    import torch
    from bocf_b200 import _lib
    lib = _lib.load_library()
    dA = torch.from_numpy(np.ascontiguousarray(A, dtype=np.float64)).cuda()
    dB = torch.from_numpy(np.ascontiguousarray(B, dtype=np.float64)).cuda()
    out = torch.zeros((A.shape[0], B.shape[0]), dtype=torch.float64, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(lib.bocf_debug_split_gemm(p(dA), p(dB), A.shape[0], B.shape[0], A.shape[1], S, tri, p(out), None))
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("S", [3, 4, 5, 6])
@pytest.mark.parametrize("shape", [(128, 64, 64), (1, 1, 1), (130, 49, 65), (256, 200, 300), (700, 1000, 1000)])
def test_split_gemm_is_exact_on_integers(cuda_device, S, shape):
    R, N, K = shape
    rng = np.random.default_rng(R + N + K + S)
    A = rng.integers(-100, 101, size=(R, K)).astype(np.float64)
    B = rng.integers(-100, 101, size=(N, K)).astype(np.float64)
    got = _split_gemm(A, B, S)
    assert np.array_equal(got, A @ B.T)


@pytest.mark.parametrize("S", [3, 4, 5, 6])
def test_split_gemm_quantisation_error(cuda_device, S):
    rng = np.random.default_rng(S)
    A = rng.standard_normal((384, 1000)) * np.exp(rng.uniform(-8, 8, size=(384, 1)))     # rows of very different scale
    B = rng.standard_normal((1000, 1000)) * np.exp(rng.uniform(-8, 8, size=(1000, 1)))
    ref = A @ B.T
    got = _split_gemm(A, B, S)
    # per-row / per-column scaling: the error is relative to |A_i| |B_k|, not to the largest entry of the product
    scale = np.abs(A).max(1)[:, None] * np.abs(B).max(1)[None, :] * np.sqrt(A.shape[1])
    assert np.max(np.abs(got - ref) / scale) < 64.0 * 256.0 ** (-S)


@pytest.mark.parametrize("S", [4, 5])
def test_split_gemm_triangular_modes(cuda_device, S):
    rng = np.random.default_rng(7)
    n = 500
    Lm = np.tril(rng.standard_normal((n, n)))
    A = rng.standard_normal((256, n))
    tol = 200 * 256.0 ** (-S)
    assert rel_err(_split_gemm(A, Lm, S, tri=1), A @ Lm.T) < tol           # K index <= column index
    assert rel_err(_split_gemm(A, Lm.T.copy(), S, tri=2), A @ Lm) < tol    # K index >= column index


@pytest.mark.parametrize("kind", ["se", "rbf", "matern52", "matern32"])
@pytest.mark.parametrize("shape", [(3, 4, 60, 300), (2, 10, 257, 129), (5, 1, 16, 1), (1, 16, 130, 515)])
def test_split5_posterior_meets_fp64_bar(cuda_device, kind, shape):
    m, d, n, N = shape
    P = make_problem(m=m, d=d, n=n, H=2, kind=kind, N=N, S=4, seed=11 * m + d)
    om = oracle_model(P)
    pm = product_model(P, cuda_device, precision="split5")
    assert pm.active_slices() == 5
    for h in range(P.H):
        om.set_hyperparameters(h)
        pm.set_hyperparameters(h)
        mu_o, v_o = om.posterior_mean(P.Xc), om.posterior_variance(P.Xc)
        dm_o, dv_o = om.posterior_mean_gradient(P.Xc), om.posterior_variance_gradient(P.Xc)
        mu, v = pm.posterior_mean(P.Xc), pm.posterior_variance(P.Xc)
        dm, dv = pm.posterior_mean_gradient(P.Xc), pm.posterior_variance_gradient(P.Xc)
        assert rel_err(mu, mu_o) < 1e-8 and rel_err(dm, dm_o) < 1e-8          # untouched fp64 path
        assert np.max(np.abs(v - v_o) / np.abs(v_o)) < 1e-6                    # north-star fp64 bar, element-wise
        assert rel_err(dv, dv_o) < 1e-6
        vn = pm.posterior_variance_noiseless(P.Xc)
        assert rel_err(vn, om.posterior_variance_noiseless(P.Xc)) < 1e-6


@pytest.mark.parametrize("prec,tol_v,tol_a", [("split4", 1e-4, 1e-4), ("split5", 1e-6, 1e-6), ("split6", 1e-8, 1e-7)])
def test_split_eicf_value_gradient_and_argmax(cuda_device, prec, tol_v, tol_a):
    P = make_problem(m=4, d=6, n=200, H=1, kind="matern52", composite="sumsq_target", N=2048, S=128, seed=4)
    a_o, g_o = oracle_acq(P, grad=True)
    a, g = product_acq(P, grad=True, device=cuda_device, precision=prec)
    assert rel_err(a, a_o) < tol_a and rel_err(g, g_o) < tol_a
    assert int(np.argmax(a)) == int(np.argmax(a_o))                           # same selected next point
    om, pm = oracle_model(P), product_model(P, cuda_device, precision=prec)
    v, v_o = pm.posterior_variance(P.Xc), om.posterior_variance(P.Xc)
    assert np.max(np.abs(v - v_o) / np.abs(v_o)) < tol_v


def test_split_matches_fp64_mode_across_chunks(cuda_device):
    # several candidate chunks, ragged tail, switching precision on a live handle
    import bocf_b200
    P = make_problem(m=3, d=5, n=333, H=1, kind="rbf", N=3001, S=8, seed=9)
    pm = product_model(P, cuda_device, precision="fp64")
    assert pm.active_slices() == 0
    v64, dv64 = pm.posterior_variance(P.Xc), pm.posterior_variance_gradient(P.Xc)
    pm.set_precision("split5")
    assert pm.active_slices() == 5
    _lib = bocf_b200._lib
    _lib.check(pm._lib.bocf_model_set_scratch_limit(pm._handle, 64 << 20))      # force several chunks
    v5, dv5 = pm.posterior_variance(P.Xc), pm.posterior_variance_gradient(P.Xc)
    assert np.max(np.abs(v5 - v64) / np.abs(v64)) < 1e-6 and rel_err(dv5, dv64) < 1e-6
    pm.set_precision("fp64")
    assert np.array_equal(pm.posterior_variance(P.Xc), v64)


def test_auto_precision_follows_conditioning(cuda_device):
    well = make_problem(m=2, d=6, n=150, H=1, kind="matern52", N=64, S=4, seed=1)                 # noise 1e-2
    pm = product_model(well, cuda_device, precision="auto")
    assert pm.active_slices() in (4, 5)
    om = oracle_model(well)
    v, v_o = pm.posterior_variance(well.Xc), om.posterior_variance(well.Xc)
    assert np.max(np.abs(v - v_o) / np.abs(v_o)) < 1e-6
    ill = make_problem(m=2, d=3, n=120, H=1, kind="se", N=64, S=4, noise=1e-10, seed=2)           # cond ~ 1e10
    pm = product_model(ill, cuda_device, precision="auto")
    assert pm.active_slices() == 0                                                                # stays on fp64 DMMA
    om = oracle_model(ill)
    assert np.max(np.abs(pm.posterior_variance(ill.Xc) - om.posterior_variance(ill.Xc))) < 1e-7


def test_auto_and_mixed_pick_their_schemes(cuda_device):
    # auto: 15 digit pairs for the variance, 13 for its gradient; mixed: 13 / 8 (include/bocf_b200.h)
    P = make_problem(m=3, d=6, n=300, H=1, kind="matern52", composite="sumsq_target", N=1024, S=64, seed=6)
    om = oracle_model(P)
    v_o, dv_o = om.posterior_variance(P.Xc), om.posterior_variance_gradient(P.Xc)
    a_o, g_o = oracle_acq(P, grad=True)
    for prec, schemes, tol_v, tol_dv, tol_a in [("auto", (554, 442), 1e-6, 1e-6, 1e-6), ("split54", (554, 442), 1e-6, 1e-6, 1e-6),
                                               ("mixed", (442, 331), 2e-5, 1e-4, 1e-4), ("split43", (442, 331), 2e-5, 1e-4, 1e-4)]:
        pm = product_model(P, cuda_device, precision=prec)
        v, dv = pm.posterior_variance(P.Xc), pm.posterior_variance_gradient(P.Xc)
        assert pm.active_scheme() == schemes, (prec, pm.active_scheme())
        assert np.max(np.abs(v - v_o) / np.abs(v_o)) < tol_v, (prec, np.max(np.abs(v - v_o) / np.abs(v_o)))
        assert rel_err(dv, dv_o) < tol_dv, (prec, rel_err(dv, dv_o))
        a, g = product_acq(P, grad=True, device=cuda_device, model=pm)
        assert rel_err(a, a_o) < tol_a and rel_err(g, g_o) < tol_a, (prec, rel_err(a, a_o), rel_err(g, g_o))
        assert int(np.argmax(a)) == int(np.argmax(a_o))


@pytest.mark.parametrize("kind", ["rbf", "matern52"])
@pytest.mark.parametrize("noise", [1e-1, 1e-2, 1e-3, 1e-4, 1e-6])
def test_auto_precision_holds_the_fp64_bar_across_conditioning(cuda_device, kind, noise):
    # whatever AUTO picks (4, 5 or 6 digit planes, or the fp64 engine when L^-1 is too large for six planes), the
    # variance must stay within the north-star fp64 bar of the oracle, element-wise
    P = make_problem(m=2, d=4, n=220, H=1, kind=kind, N=512, S=4, noise=noise, seed=int(-np.log10(noise)))
    om = oracle_model(P)
    pm = product_model(P, cuda_device, precision="auto")
    v, v_o = pm.posterior_variance(P.Xc), om.posterior_variance(P.Xc)
    assert np.max(np.abs(v - v_o) / np.abs(v_o)) < 1e-6, (pm.active_slices(), noise)
    dv, dv_o = pm.posterior_variance_gradient(P.Xc), om.posterior_variance_gradient(P.Xc)
    assert rel_err(dv, dv_o) < 1e-6, (pm.active_slices(), noise)
    if noise >= 1e-2:
        assert pm.active_slices() in (4, 5)          # well conditioned: the tensor-core path is the one that runs


@pytest.mark.parametrize("noise", [1e-2, 1e-3, 1e-4])
def test_auto_holds_the_bar_next_to_training_inputs(cuda_device, noise):
    # Candidates next to (and exactly at) training inputs: the variance collapses to the noise level there, so an
    # absolute error that is harmless elsewhere becomes a large RELATIVE one -- and these are the points the acquisition
    # optimiser converges to, where sigma enters the pathwise gradient as 0.5 / sigma.  AUTO sizes the digit planes for
    # it (api.cu apply_precision, bound (ii); CPU model: tests/test_split_numerics.py).
    P = make_problem(m=2, d=4, n=220, H=1, kind="matern52", composite="sumsq_target", N=128, S=64, noise=noise, seed=3)
    rng = np.random.default_rng(1)
    Xc = np.ascontiguousarray(np.concatenate([P.X[:96] + 1e-4 * rng.standard_normal((96, P.d)), P.X[:32]]))
    om, pm = oracle_model(P), product_model(P, cuda_device, precision="auto")
    v, v_o = pm.posterior_variance(Xc), om.posterior_variance(Xc)
    assert v_o.min() < 3.0 * noise                                             # the variance does collapse here
    assert np.max(np.abs(v - v_o) / v_o) < 1e-6, (pm.active_slices(), noise, np.max(np.abs(v - v_o) / v_o))
    vn, vn_o = pm.posterior_variance_noiseless(Xc), om.posterior_variance_noiseless(Xc)
    big = vn_o >= 0.25 * noise                                                 # the noiseless variance has no floor
    assert np.max(np.abs(vn - vn_o)[big] / vn_o[big]) < 1e-6, (pm.active_slices(), noise)
    assert np.max(np.abs(vn - vn_o)) < 2.5e-7 * noise, (pm.active_slices(), noise)
    if noise <= 1e-4:
        assert pm.active_slices() in (0, 6)                                   # five planes would miss the bar here
    a_o, g_o = oracle_acq(P, grad=True, Xc=Xc)
    a, g = product_acq(P, grad=True, device=cuda_device, Xc=Xc, model=pm)
    assert_close(a, a_o, 1e-6, "acq next to the data")
    assert_close(g, g_o, 1e-5, "grad acq next to the data")
