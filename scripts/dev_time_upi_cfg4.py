"""uPI (value only) at BASELINE.json configs[3]: m=8, n=500, 64 parameter samples, S=256, 256k candidates -- per-kernel time."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bocf_b200
from bocf_b200 import _lib
from tests.helpers import make_problem, product_model, product_utility
N = 262144
P = make_problem(m=8, d=8, n=500, H=1, kind="matern52", composite="sumsq_target", N=4096, S=256, L=64, seed=0)
model = product_model(P, "cuda:0")
acq = bocf_b200.uPI(model, None, utility=product_utility(P))
acq.W_samples = P.Z
acq.use_full_support = False                               # the 64 theta samples as an explicit sample set (bench.py)
acq.utility.parameter_dist.sample = lambda k: P.theta
Xd = torch.from_numpy(np.random.default_rng(7).uniform(size=(N, 8))).cuda()
for _ in range(2):
    acq._compute_acq(Xd)
torch.cuda.synchronize()
_lib.profile_enable(True)
t = time.perf_counter()
for _ in range(3):
    acq._compute_acq(Xd)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 3
print("uPI cfg4: %.2f ms per 256k candidates = %.2f M evals/s" % (1e3 * dt, N / dt / 1e6))
print({k: round(v[1] / 3, 2) for k, v in _lib.profile_report().items()})

