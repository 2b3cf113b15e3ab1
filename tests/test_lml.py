"""Marginal likelihood and hyper-parameter gradients.

CPU: the oracle restatement against vectors produced by the reference's own code (tests/golden/make_golden_lml.py).
GPU: the device pass (lml.cu through bocf_model_log_likelihood) against the same vectors and against the oracle on
larger problems; ML-II on device reaches the optimum scipy finds with the oracle objective."""
import glob
import os

import numpy as np
import pytest

from tests.helpers import make_problem, oracle_model, product_model, rel_err

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "lml_*.npz")))
IDS = [os.path.basename(p)[4:-4] for p in GOLDEN]


def _oracle_all(om, H, m, d):
    lml, gv, gl, gn = np.zeros((H, m)), np.zeros((H, m)), np.zeros((H, m, d)), np.zeros((H, m))
    for j in range(m):
        for h in range(H):
            g = om.output[j].model_instances[h]
            lml[h, j] = g.log_likelihood()
            gv[h, j], gl[h, j], gn[h, j] = g.likelihood_gradients()
    return lml, gv, gl, gn


def _load(path):
    from tests.helpers import Problem
    z = np.load(path)
    P = Problem()
    P.kind = str(z["kind"])
    P.X, P.variance, P.lengthscale, P.noise = z["X"], z["variance"], z["lengthscale"], z["noise"]
    P.Y = [z["Y"][:, j:j + 1] for j in range(z["Y"].shape[1])]
    P.H, P.m = P.variance.shape
    P.d = P.X.shape[1]
    return P, z


def test_golden_files_present():
    assert len(GOLDEN) == 4


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_oracle_likelihood_matches_reference(path):
    P, z = _load(path)
    lml, gv, gl, gn = _oracle_all(oracle_model(P), P.H, P.m, P.d)
    assert rel_err(lml, z["lml"]) < 1e-12 and rel_err(gv, z["g_var"]) < 1e-11
    assert rel_err(gl, z["g_len"]) < 1e-11 and rel_err(gn, z["g_noise"]) < 1e-11


def test_oracle_gradients_are_gradients():
    # central differences of the oracle's own log likelihood (kernel_tests.py:366-422 pattern)
    P = make_problem(m=1, d=3, n=30, H=1, kind="matern52", N=4, S=4, seed=3)
    base = _oracle_all(oracle_model(P), 1, 1, 3)
    eps = 1e-6
    for name, idx in (("variance", (0, 0)), ("noise", (0, 0)), ("lengthscale", (0, 0, 1))):
        vals = []
        for sgn in (+1, -1):
            arr = getattr(P, name).copy()
            arr[idx] += sgn * eps
            old = getattr(P, name)
            setattr(P, name, arr)
            vals.append(_oracle_all(oracle_model(P), 1, 1, 3)[0][0, 0])
            setattr(P, name, old)
        fd = (vals[0] - vals[1]) / (2 * eps)
        an = {"variance": base[1][0, 0], "noise": base[3][0, 0], "lengthscale": base[2][0, 0, 1]}[name]
        assert abs(fd - an) < 1e-5 * max(1.0, abs(an))


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_cuda_likelihood_matches_reference_golden(cuda_device, path):
    P, z = _load(path)
    pm = product_model(P, cuda_device)
    lml, gv, gl, gn = pm.log_likelihood_and_gradients()
    assert rel_err(lml, z["lml"]) < 1e-10 and rel_err(gv, z["g_var"]) < 1e-9
    assert rel_err(gl, z["g_len"]) < 1e-9 and rel_err(gn, z["g_noise"]) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n,d", [("se", 130, 2), ("rbf", 257, 6), ("matern52", 600, 10), ("matern32", 129, 16)])
def test_cuda_likelihood_matches_oracle(cuda_device, kind, n, d):
    P = make_problem(m=3, d=d, n=n, H=2, kind=kind, N=8, S=4, seed=n + d)
    pm = product_model(P, cuda_device)
    lml, gv, gl, gn = pm.log_likelihood_and_gradients()
    o = _oracle_all(oracle_model(P), P.H, P.m, P.d)
    assert lml.shape == (2, 3) and gl.shape == (2, 3, d)
    assert rel_err(lml, o[0]) < 1e-10 and rel_err(gv, o[1]) < 1e-8
    assert rel_err(gl, o[2]) < 1e-8 and rel_err(gn, o[3]) < 1e-8


@pytest.mark.gpu
def test_ml2_fit_on_device_reaches_the_oracle_optimum(cuda_device):
    import scipy.optimize
    from oracle.models import multi_outputGP as OracleGP
    P = make_problem(m=2, d=2, n=60, H=1, kind="rbf", N=8, S=4, seed=21, noise=2e-2)
    pm = product_model(P, cuda_device)
    before = pm.log_likelihood()[0]
    after = pm.fit_hyperparameters(max_iters=100)
    assert np.all(after > before)
    # the same ML-II problem solved output by output with the oracle objective
    for j in range(P.m):
        def f_df(theta):
            th = np.exp(theta)
            om = OracleGP.from_hyper_samples(P.kind, th[:1][None, :], th[1:3][None, None, :], th[3:4][None, :], ARD=True)
            om.output_dim = 1
            om.updateModel(P.X, [P.Y[j]])
            g = om.output[0].model_instances[0]
            gv, gl, gn = g.likelihood_gradients()
            return -g.log_likelihood(), -np.concatenate([[gv * th[0]], gl * th[1:3], [gn * th[3]]])
        th0 = np.log(np.concatenate([P.variance[0, j:j + 1], P.lengthscale[0, j], P.noise[0, j:j + 1]]))
        res = scipy.optimize.fmin_l_bfgs_b(f_df, th0, maxiter=100)
        assert after[j] > -res[1] - 1e-3 * max(1.0, abs(res[1]))          # device fit is at least as good (same optimum)
    # the fitted model still predicts consistently with an oracle model at the fitted hyper-parameters
    kind, var, ls, nz = pm._hyp
    P.variance, P.lengthscale, P.noise = var, ls, nz
    assert rel_err(pm.posterior_mean(P.Xc), oracle_model(P).posterior_mean(P.Xc)) < 1e-7
