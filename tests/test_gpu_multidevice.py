"""One process driving two GPUs ("one process driving 8 devices"): every kernel that needs more than
48 KB of dynamic shared memory sets its function attribute per device context (posterior.cu, chol.cu, lml.cu,
split_gemm.cu), so a second handle on cuda:1 must work in both contraction modes.  Skipped on one-GPU boxes."""
import numpy as np
import pytest

from tests.helpers import make_problem, oracle_model, product_model, rel_err

pytestmark = pytest.mark.gpu


# m = 3: one output group; m = 8: the factorisation / likelihood pass fork onto the handle's own side stream (two output
# groups, chol.cu: for_each_output_group), which must live on the handle's device
@pytest.mark.parametrize("m", [3, 8])
@pytest.mark.parametrize("precision", ["fp64", "auto"])
def test_two_handles_on_two_devices(built_library, precision, m):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    P = make_problem(m=m, d=5, n=300, H=1, kind="matern52", N=700, S=8, seed=21)
    om = oracle_model(P)
    v_o, dv_o = om.posterior_variance(P.Xc), om.posterior_variance_gradient(P.Xc)
    models = [product_model(P, "cuda:%d" % k, precision=precision) for k in (0, 1)]     # factorise on both first
    for k, pm in enumerate(models):
        v, dv = pm.posterior_variance(P.Xc), pm.posterior_variance_gradient(P.Xc)
        assert np.max(np.abs(v - v_o) / v_o) < 1e-6, k
        assert rel_err(dv, dv_o) < 1e-6, k
        lml = pm.log_likelihood()
        assert np.all(np.isfinite(lml))
    # interleaved calls keep working (the handles switch the current device themselves)
    a = models[0].posterior_mean(P.Xc)
    b = models[1].posterior_mean(P.Xc)
    assert np.array_equal(a, b)
