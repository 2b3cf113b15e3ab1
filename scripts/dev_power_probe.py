"""Where do the joules go?  Runs loops of (a) the K* kernel alone, (b) K* + first contraction, (c) the full posterior
with gradients, sampling SM clock and board power through NVML.  Under the 1 kW cap a step's time is its energy, so
energy per kernel class (power x time share) is the quantity to minimise."""
import os, sys, time, threading
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bocf_b200 import _lib
from tests.helpers import make_problem, product_model
import pynvml as nv
nv.nvmlInit(); h = nv.nvmlDeviceGetHandleByIndex(0)
P = make_problem(m=16, d=10, n=1000, H=1, kind="matern52", N=4096, S=64, seed=0)
pm = product_model(P, "cuda:0", precision=os.environ.get("PROBE_PRECISION", "auto"))
NC = 14080 * 8
X = torch.rand((NC, 10), dtype=torch.float64, device="cuda")
def loop(name, fn, secs):
    clocks, power, stop = [], [], threading.Event()
    def samp():
        while not stop.is_set():
            clocks.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)); power.append(nv.nvmlDeviceGetPowerUsage(h) / 1e3); time.sleep(0.05)
    th = threading.Thread(target=samp); fn(); torch.cuda.synchronize()
    _lib.profile_enable(True); th.start(); t0 = time.time(); calls = 0
    while time.time() - t0 < secs: fn(); calls += 1
    torch.cuda.synchronize(); el = time.time() - t0; stop.set(); th.join()
    prof = _lib.profile_report(); _lib.profile_enable(False)
    ms = {k: round(v[1] / calls, 2) for k, v in prof.items()}
    half = len(power) // 2
    pw, ck = float(np.median(power[half:])), int(np.median(clocks[half:]))
    print("%-34s %6.1f ms/call  clock %4d MHz  power %4.0f W  -> %5.1f J per %d candidates   %s"
          % (name, 1e3 * el / calls, ck, pw, pw * el / calls, NC, ms), flush=True)
idle = [nv.nvmlDeviceGetPowerUsage(h) / 1e3 for _ in range(5)]
print("idle power %.0f W" % np.median(idle))
loop("K* (mean + mean gradient)", lambda: pm._posterior(X, want_dmean=True), 3.0)
loop("K* + first contraction (no grad)", lambda: pm._posterior(X, want_var=True), 3.0)
loop("full posterior with gradients", lambda: pm._posterior(X, want_var=True, want_dmean=True, want_dvar=True), 4.0)
loop("K* (mean + mean gradient) again", lambda: pm._posterior(X, want_dmean=True), 3.0)
