"""Timing experiment: MMA-dominated raw split GEMM (large K, full range) with the A operand fetched from shared memory
(normal) or from spare tensor-memory columns (BOCF_SPLIT_EXP=5, results invalid) -- isolates the smem operand port."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bocf_b200 import _lib
lib = _lib.load_library()
rng = np.random.default_rng(0)
R, N, K = 148 * 128 * 8, 1024, 1024
A = torch.from_numpy(rng.standard_normal((R, K))).cuda()
B = torch.from_numpy(rng.standard_normal((N, K))).cuda()
out = torch.zeros((R, N), dtype=torch.float64, device="cuda")
p = lambda t: ctypes.c_void_p(t.data_ptr())
for S in (3, 4, 5):
    for rep in range(2):
        _lib.profile_enable(True)
        _lib.check(lib.bocf_debug_split_gemm(p(A), p(B), R, N, K, S, 0, p(out), None))
        torch.cuda.synchronize()
        prof = _lib.profile_report()
        _lib.profile_enable(False)
    ms = prof["split_raw_kernel"][1]
    nt = {3: 64, 4: 64, 5: 48}[S]
    tiles = (R // 128) * ((N + nt - 1) // nt)
    ksteps = tiles * (K // 32) / 148
    print("S=%d exp=%s  raw kernel %.3f ms  -> %.0f clk per K step at 1.9 GHz (MMA floor %d)" % (
        S, os.environ.get("BOCF_SPLIT_EXP", "0"), ms, ms * 1e-3 * 1.9e9 / ksteps, S * (S + 1) // 2 * nt // 2), flush=True)
