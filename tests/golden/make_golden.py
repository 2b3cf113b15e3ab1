#!/usr/bin/env python
"""Generate golden input/output vectors by running the REFERENCE's own code (see ref_harness.py).

    make -C oracle && python tests/golden/make_golden.py

Writes tests/golden/*.npz.  Each file holds the explicit inputs (X, Y, hyper-samples, candidates, base samples
Z, theta support and probabilities) and the outputs of the reference's real classes:
  multi_outputGP.posterior_mean / posterior_variance / posterior_mean_gradient / posterior_variance_gradient /
  predict per hyper-sample, and the acquisition's _compute_acq / _compute_acq_withGradients.
The composite utilities are the U_func / dU_func definitions extracted verbatim (ast) from the reference's
experiment scripts test_1a.py / test_2a.py / test_3a.py / test_5a.py / test_1b.py.
Needs /root/reference, so it runs only in the build container; the fixtures are committed.
"""
import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.golden import ref_harness as rh      # noqa: E402
from tests.helpers import make_problem          # noqa: E402


def script_functions(script, names, extra_globals=None):
    """Compile the named function definitions of a reference script, verbatim."""
    src = open(os.path.join(rh.REF, script)).read()
    tree = ast.parse(src)
    env = {"np": np}
    env.update(extra_globals or {})
    out = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names and node.name not in out:
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, script, "exec"), env)
            out[node.name] = env[node.name]
    assert set(out) == set(names), (script, out.keys())
    return out


COMPOSITE_SOURCE = {
    "sumsq_target": ("test_1a.py", {}),
    "neg_sum_exp": ("test_2a.py", {}),
    "exp_cos": ("test_3a.py", {}),
    "rosen_composite": ("test_5a.py", None),     # needs globals d, m (m = 2(d-1))
    "linear": ("test_1b.py", {}),
}


def reference_utility(ns, composite, m, theta, prob):
    script, extra = COMPOSITE_SOURCE[composite]
    if extra is None:
        extra = {"d": m // 2 + 1, "m": m}
    f = script_functions(script, ("U_func", "dU_func"), extra)
    pd = ns.parameter_distribution.ParameterDistribution(continuous=False, support=theta, prob_dist=prob)
    return ns.utility.Utility(func=f["U_func"], dfunc=f["dU_func"], parameter_dist=pd,
                              linear=(composite == "linear"))


CASES = [
    # name, acquisition class, problem kwargs
    ("eicf_sumsq_matern52", "uEI_noiseless", dict(m=4, d=5, n=40, H=2, kind="matern52", composite="sumsq_target", N=24, S=16, L=2, seed=101)),
    ("eicf_sumsq_se", "uEI_noiseless", dict(m=3, d=4, n=33, H=2, kind="se", composite="sumsq_target", N=20, S=25, L=1, seed=102)),
    ("eicf_negsumexp_rbf", "uEI_noiseless", dict(m=3, d=3, n=30, H=1, kind="rbf", composite="neg_sum_exp", N=20, S=16, L=1, seed=133)),
    ("eicf_expcos_matern32", "uEI_noiseless", dict(m=5, d=2, n=25, H=2, kind="matern32", composite="exp_cos", N=20, S=16, L=1, seed=104)),
    ("eicf_rosen_se", "uEI_noiseless", dict(m=4, d=3, n=28, H=1, kind="se", composite="rosen_composite", N=20, S=16, L=1, seed=105)),
    ("eicf_linear_rbf", "uEI_noiseless", dict(m=4, d=4, n=30, H=2, kind="rbf", composite="linear", N=20, S=16, L=3, seed=106)),
    ("upi_sumsq_se", "uPI", dict(m=4, d=4, n=30, H=2, kind="se", composite="sumsq_target", N=24, S=32, L=2, seed=107)),
    ("maei_linear_matern52", "maEI", dict(m=5, d=4, n=36, H=3, kind="matern52", composite="linear", N=24, S=4, L=4, seed=108)),
    ("mapi_linear_se", "maPI", dict(m=4, d=3, n=30, H=2, kind="se", composite="linear", N=24, S=4, L=3, seed=109)),
    ("ei_single_se", "EI", dict(m=1, d=3, n=20, H=1, kind="se", composite="linear", N=20, S=4, L=1, seed=110)),
    ("pi_single_rbf", "PI", dict(m=1, d=2, n=20, H=1, kind="rbf", composite="linear", N=20, S=4, L=1, seed=111)),
]


def run_case(ns, name, acq_name, kw):
    P = make_problem(**kw)
    if acq_name in ("EI", "PI"):
        P.theta = np.ones((1, 1))
        P.prob = np.ones(1)
    # test_3a / test_5a / test_2a use a dummy scalar parameter support np.ones((1,)): keep theta 2-d (L, 1)
    model = rh.make_reference_model(ns, P.kind, P.X, P.Y, P.variance, P.lengthscale, P.noise)
    out = dict(X=P.X, Y=np.concatenate(P.Y, axis=1).T, variance=P.variance, lengthscale=P.lengthscale, noise=P.noise,
               Xc=P.Xc, Z=P.Z, theta=P.theta, prob=P.prob, kind=P.kind, composite=P.composite, acq=acq_name)
    for h in range(P.H):
        model.set_hyperparameters(h)
        out["mean_h%d" % h] = model.posterior_mean(P.Xc)
        out["var_h%d" % h] = model.posterior_variance(P.Xc)
        out["dmean_h%d" % h] = model.posterior_mean_gradient(P.Xc)
        out["dvar_h%d" % h] = model.posterior_variance_gradient(P.Xc)
        pm, pv = model.predict(P.Xc)
        out["pmean_h%d" % h], out["pvar_h%d" % h] = pm, pv
        out["varnl_h%d" % h] = model.posterior_variance_noiseless(P.Xc)
    out["mean_train_h0"] = (model.set_hyperparameters(0), model.posterior_mean_at_evaluated_points())[1]
    U = reference_utility(ns, P.composite, P.m, P.theta, P.prob)
    cls = getattr(getattr(ns, acq_name), acq_name)
    np.random.seed(0)
    acq = cls(model, None, optimizer=None, utility=U)
    if hasattr(acq, "W_samples"):
        acq.W_samples = P.Z
    model.set_hyperparameters(0)
    if acq_name in ("uEI_noiseless", "uPI"):
        out["acq_value"] = acq._compute_acq(P.Xc, parallel=False)       # sequential branch :46
        model.set_hyperparameters(0)
        out["acq_value_pool"] = acq._compute_acq(P.Xc, parallel=True)   # pathos branch :44 (serial stand-in pool)
    else:
        out["acq_value"] = acq._compute_acq(P.Xc)
    if acq.analytical_gradient_prediction:
        model.set_hyperparameters(0)
        a, g = acq._compute_acq_withGradients(P.Xc)
        out["acq_grad_value"], out["acq_grad"] = a, g
        model.set_hyperparameters(0)
        f, df = acq.acquisition_function_withGradients(P.Xc)
        assert np.array_equal(f, -a) and np.array_equal(df, -g)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    return path, out


def main():
    ns = rh.install()
    for name, acq_name, kw in CASES:
        path, out = run_case(ns, name, acq_name, kw)
        nz = float(np.mean(out["acq_value"] > 0))
        print("%-24s %-14s acq>0: %.2f  max %.4g  -> %s" % (name, acq_name, nz, out["acq_value"].max(),
                                                           os.path.relpath(path, ROOT)))


if __name__ == "__main__":
    main()
