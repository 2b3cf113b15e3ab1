"""Where one likelihood pass of the hyper-parameter inference spends its time (host view)."""
import ctypes, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bocf_b200 as B
from bocf_b200 import _lib
from tests.helpers import make_problem
for (m, d, n) in [(4, 6, 200), (5, 4, 60), (16, 10, 1000)]:
    P = make_problem(m=m, d=d, n=n, kind="rbf", N=8, S=8)
    mod = B.multi_outputGP(m, device="cuda:0", hyper_inference="none")
    mod.set_hyperparameter_samples(P.variance, P.lengthscale, P.noise, kind="rbf")
    mod.updateModel(P.X, P.Y)
    lib = mod._lib
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    v, l, z = [np.ascontiguousarray(a) for a in (P.variance, P.lengthscale, P.noise)]
    jit = np.zeros_like(v); lml = np.zeros_like(v); gv = np.zeros_like(v); gl = np.zeros_like(l); gn = np.zeros_like(v)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    R = 100 if n < 1000 else 10
    t = [0.0] * 4
    for r in range(R + 3):
        if r == 3: t = [0.0] * 4
        a = time.perf_counter(); lib.bocf_model_set_hypers(mod._handle, 1, vp(v), vp(l), vp(z))
        b = time.perf_counter(); lib.bocf_model_factorize(mod._handle, vp(jit), st)
        c = time.perf_counter(); lib.bocf_model_log_likelihood(mod._handle, vp(lml), vp(gv), vp(gl), vp(gn), st)
        e = time.perf_counter(); mod._upload_and_factorize(upload_data=False); mod.log_likelihood_and_gradients()
        f = time.perf_counter()
        t[0] += b - a; t[1] += c - b; t[2] += e - c; t[3] += f - e
    print("m=%d d=%d n=%d  ms per pass: set_hypers %.3f  factorize %.3f  log_likelihood %.3f  | python wrappers total %.3f" % (
        m, d, n, *(1e3 * x / R for x in t)), flush=True)
    _lib.profile_enable(True)
    lib.bocf_model_factorize(mod._handle, vp(jit), st); lib.bocf_model_log_likelihood(mod._handle, vp(lml), vp(gv), vp(gl), vp(gn), st)
    print(_lib.profile_report()); _lib.profile_enable(False)
