"""Is the sweep power-limited?  Times the K* kernel alone (mean-only posterior) and inside the full sweep, sampling SM
clocks through NVML; a slower K* and lower clocks in the mixed loop mean the 1 kW cap, not the kernels, sets the pace."""
import os, sys, time, threading
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bocf_b200 import _lib
from tests.helpers import make_problem, product_model
import pynvml as nv
nv.nvmlInit(); h = nv.nvmlDeviceGetHandleByIndex(0)
P = make_problem(m=16, d=10, n=1000, H=1, kind="matern52", N=4096, S=64, seed=0)
pm = product_model(P, "cuda:0")
X = torch.rand((262144, 10), dtype=torch.float64, device="cuda")
def loop(fn, secs):
    clocks, power, stop = [], [], threading.Event()
    def samp():
        while not stop.is_set():
            clocks.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)); power.append(nv.nvmlDeviceGetPowerUsage(h) / 1e3); time.sleep(0.05)
    th = threading.Thread(target=samp); fn(); torch.cuda.synchronize()
    _lib.profile_enable(True); th.start(); t0 = time.time()
    while time.time() - t0 < secs: fn()
    torch.cuda.synchronize(); stop.set(); th.join()
    prof = _lib.profile_report(); _lib.profile_enable(False)
    return {k: round(v[1] / v[0], 3) for k, v in prof.items()}, int(np.median(clocks)), int(np.median(power))
print("kstar only      :", loop(lambda: pm._posterior(X, want_dmean=True), 3.0))
print("full posterior  :", loop(lambda: pm._posterior(X, want_var=True, want_dmean=True, want_dvar=True), 4.0))
print("kstar only again:", loop(lambda: pm._posterior(X, want_dmean=True), 3.0))
