"""Three likelihood passes (factorise + log likelihood with gradients) at cfg 3's model size, for an ncu launch list:
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/dev_fit_pass_once.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bocf_b200 as B
from tests.helpers import make_problem
m, d, n = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (16, 10, 1000)))
P = make_problem(m=m, d=d, n=n, kind="matern52", N=8, S=8, prior_draw=False)
mod = B.multi_outputGP(m, device="cuda:0", hyper_inference="none")
mod.set_hyperparameter_samples(P.variance, P.lengthscale, P.noise, kind="matern52")
mod.updateModel(P.X, P.Y)
for _ in range(3):
    mod._upload_and_factorize(upload_data=False)
    out = mod.log_likelihood_and_gradients()
print("lml", out[0].ravel()[:4])
