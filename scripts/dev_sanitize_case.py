"""Tiny split-mode posterior + EI-CF evaluation for compute-sanitizer (memcheck) runs."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import make_problem, product_model, product_acq, oracle_acq, rel_err

P = make_problem(m=2, d=5, n=150, H=1, kind="matern52", N=300, S=32, seed=1)
for prec in ("split5", "split4", "fp64"):
    pm = product_model(P, "cuda:0", precision=prec)
    v = pm.posterior_variance(P.Xc)
    dv = pm.posterior_variance_gradient(P.Xc)
    a, g = product_acq(P, grad=True, device="cuda:0", model=pm)
    lml = pm.log_likelihood()
    print(prec, float(v.sum()), float(np.abs(dv).sum()), float(a.sum()), lml.ravel())
a_o, g_o = oracle_acq(P, grad=True)
print("acq err vs oracle", rel_err(a, a_o), rel_err(g, g_o))
