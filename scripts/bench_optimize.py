"""Latency of one full acquisition optimisation at cfg 3's model (acquisition_optimizer.py:114-148,
optimizer.py:319-354,425-466): 400 random starts scored in one call -> 16 anchors + baseline -> batched L-BFGS-B rounds
of N <= 17 candidates.  Reports wall time, rounds, device launches and microseconds per batched f_df round.

    python scripts/bench_optimize.py [--repeat 3] [--samples 1024]
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bocf_b200 as B
from bocf_b200 import _lib
from tests.helpers import make_problem, product_model, product_utility

ap = argparse.ArgumentParser()
ap.add_argument("--repeat", type=int, default=3)
ap.add_argument("--samples", type=int, default=1024)
ap.add_argument("--precision", default="auto")
a = ap.parse_args()
P = make_problem(m=16, d=10, n=1000, H=1, kind="matern52", composite="sumsq_target", N=64, S=a.samples, seed=0)
model = product_model(P, "cuda:0", precision=a.precision)
space = B.Design_space(space=[{'name': 'var', 'type': 'continuous', 'domain': (0, 1), 'dimensionality': P.d}])
opt = B.AcquisitionOptimizer(space, optimizer='lbfgs2', inner_optimizer='lbfgs2')          # 400 starts, 16 anchors
acq = B.uEI_noiseless(model, space, optimizer=opt, utility=product_utility(P))
acq.W_samples = P.Z
stats = {"rounds": 0, "t_fdf": 0.0, "cands": 0, "t_c": 0.0, "c_calls": 0}
orig = acq.acquisition_function_withGradients
_c_entry = model._lib.bocf_acq_eval_host


class _TimedEntry(object):                        # time spent inside the C call (copies + kernels + sync) vs Python
    def __call__(self, *args):
        t0 = time.perf_counter()
        rc = _c_entry(*args)
        stats["t_c"] += time.perf_counter() - t0
        stats["c_calls"] += 1
        return rc


class _LibProxy(object):
    def __init__(self, lib):
        self._lib = lib
        self.bocf_acq_eval_host = _TimedEntry()

    def __getattr__(self, k):
        return getattr(self._lib, k)


model._lib = _LibProxy(model._lib)


def timed_fdf(X):
    t0 = time.perf_counter()
    out = orig(X)
    stats["t_fdf"] += time.perf_counter() - t0
    stats["rounds"] += 1
    stats["cands"] += len(np.atleast_2d(X))
    return out


acq.acquisition_function_withGradients = timed_fdf
res = []
for r in range(a.repeat + 1):
    np.random.seed(r)
    stats.update(rounds=0, t_fdf=0.0, cands=0, t_c=0.0, c_calls=0)
    torch.cuda.synchronize()
    l0, t0 = _lib.launch_count(), time.perf_counter()
    x, fx = acq.optimize(x_baseline=P.X[:1])
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if r == 0:
        continue                                  # warm-up (allocations, first-touch)
    res.append({"wall_ms": 1e3 * wall, "rounds": stats["rounds"], "launches": int(_lib.launch_count() - l0),
                "us_per_fdf_round": 1e6 * stats["t_fdf"] / max(1, stats["rounds"]),
                "candidates_per_round": stats["cands"] / max(1, stats["rounds"]), "fdf_share": stats["t_fdf"] / wall,
                "us_in_c_call_per_call": 1e6 * stats["t_c"] / max(1, stats["c_calls"]), "c_calls": stats["c_calls"],
                "best_value": float(np.asarray(fx).reshape(-1)[0])})
# per-kernel device time of one small batch (library event scopes), separate pass
_lib.profile_enable(True)
Xs = np.random.default_rng(0).uniform(size=(17, P.d))
for _ in range(20):
    orig(Xs)
torch.cuda.synchronize()
prof = _lib.profile_report()
_lib.profile_enable(False)
kernel_us = {k: 1e3 * v[1] / v[0] for k, v in prof.items()}
out = {"kernel_us_per_launch_at_17_candidates": kernel_us, "workload": "cfg3 model (m=16, d=10, n=1000, Matern-5/2), EI-CF with %d base samples, AcquisitionOptimizer default "
                   "(400 starts, 16 anchors + 1 baseline, lbfgs2, batched rounds)" % a.samples,
       "precision": a.precision, "schemes": list(model.active_scheme()), "runs": res,
       "median_wall_ms": float(np.median([r["wall_ms"] for r in res])),
       "median_us_per_fdf_round": float(np.median([r["us_per_fdf_round"] for r in res]))}
print(json.dumps(out))
