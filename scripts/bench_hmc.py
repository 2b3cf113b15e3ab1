"""Wall time of one GPModel.updateModel-style hyper-parameter inference (ML-II + HMC with the reference's default sampler
settings: 100 burn-in + 10 x 10 samples, 20 leap-frog steps) for all m outputs: device lockstep (bocf_b200) next to the
sequential CPU oracle on a bounded number of samples.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bocf_b200 as B  # noqa: E402
from tests.helpers import make_problem  # noqa: E402

m, d, n = int(os.environ.get("HM", 4)), int(os.environ.get("HD", 6)), int(os.environ.get("HN", 200))
kind = os.environ.get("HKIND", "rbf")
cpu_samples = int(os.environ.get("HCPU_SAMPLES", 6))
P = make_problem(m=m, d=d, n=n, kind=kind, N=8, S=8)
K = B.kern.BY_KIND[kind]
mod = B.multi_outputGP(m, kernel=[K(d, variance=1., ARD=True) for _ in range(m)], device="cuda:0")
np.random.seed(0)
t = time.perf_counter()
mod.updateModel(P.X, P.Y)
gpu_s = time.perf_counter() - t
inf = mod._inference
first = dict(inf.timing, total_s=gpu_s, passes=inf.device_passes)
# second call on the same data: what every later BO iteration pays (context, allocations and module load are behind us)
np.random.seed(0)
t = time.perf_counter()
mod.updateModel(P.X, P.Y)
warm_s = time.perf_counter() - t
warm = dict(inf.timing, total_s=warm_s)
num = inf.n_burnin + inf.n_samples * inf.subsample_interval
moves = [int(np.sum(np.any(np.diff(c, axis=0) != 0, axis=1))) for c in inf.chain]

if cpu_samples == 0:                               # HCPU_SAMPLES=0: device side only
    print(json.dumps({"what": "ML-II + HMC hyper-parameter inference, all outputs (GPModel.updateModel defaults)", "m": m,
                      "d": d, "n": n, "kind": kind, "gpu_s": gpu_s, "device_passes": first["passes"],
                      "ms_per_pass": 1e3 * gpu_s / first["passes"], "samples": num, "accepted_moves_per_output": moves,
                      "first_call": first, "second_call": warm}))
    sys.exit(0)

from oracle.hmc import GPModelHMC, HMC  # noqa: E402
np.random.seed(0)
t_opt = t_smp = 0.0
for j in range(m):
    g = GPModelHMC(kind=kind, ARD=True, max_iters=200)
    g._create_model(P.X, P.Y[j])
    t = time.perf_counter()
    g.model.optimize(max_iters=200)
    t_opt += time.perf_counter() - t
    g.model.param_array[:] = g.model.param_array * (1. + np.random.randn(g.model.param_array.size) * 0.01)
    t = time.perf_counter()
    HMC(g.model, stepsize=g.step_size).sample(num_samples=cpu_samples, hmc_iters=g.leapfrog_steps)
    t_smp += time.perf_counter() - t
cpu_total = t_opt + t_smp * num / cpu_samples
print(json.dumps({"what": "ML-II + HMC hyper-parameter inference, all outputs (GPModel.updateModel defaults)", "m": m, "d": d,
                  "n": n, "kind": kind, "gpu_s": gpu_s, "device_passes": first["passes"],
                  "ms_per_pass": 1e3 * gpu_s / first["passes"], "samples": num, "accepted_moves_per_output": moves,
                  "cpu_oracle": {"mlii_s": t_opt, "hmc_s_for_%d_samples" % cpu_samples: t_smp,
                                 "total_s_extrapolated_to_%d_samples" % num: cpu_total, "cores": os.cpu_count(),
                                 "note": "sequential per-output numpy/LAPACK oracle (the reference's structure); ML-II timed "
                                         "in full, the chain on a bounded number of samples and scaled linearly"},
                  "speedup_vs_cpu_oracle": cpu_total / gpu_s}))
