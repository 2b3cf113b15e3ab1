"""Bring-up check of the tcgen05 split-integer GEMM (prints diagnostics; the asserts live in tests/test_gpu_split.py)."""
import ctypes
import sys
import os
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bocf_b200 import _lib


def split_gemm(A, B, S, tri=0):
    lib = _lib.load_library()
    dA = torch.from_numpy(np.ascontiguousarray(A)).cuda()
    dB = torch.from_numpy(np.ascontiguousarray(B)).cuda()
    out = torch.zeros((A.shape[0], B.shape[0]), dtype=torch.float64, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(lib.bocf_debug_split_gemm(p(dA), p(dB), A.shape[0], B.shape[0], A.shape[1], S, tri, p(out), None))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def main():
    rng = np.random.default_rng(0)
    for (R, N, K) in [(128, 64, 64), (128, 48, 64), (128, 64, 128), (256, 200, 300), (1000, 1000, 1000)]:
        for S in (3, 4, 5, 6):
            A = rng.integers(-100, 101, size=(R, K)).astype(np.float64)
            B = rng.integers(-100, 101, size=(N, K)).astype(np.float64)
            ref = A @ B.T
            got = split_gemm(A, B, S)
            err = np.abs(got - ref)
            bad = np.argwhere(err > 0)
            print("int  R=%d N=%d K=%d S=%d  max|err|=%.3e  nbad=%d/%d  first bad %s  got/ref sample %s" % (
                R, N, K, S, err.max(), len(bad), err.size, bad[:3].tolist(),
                [(float(got[tuple(b)]), float(ref[tuple(b)])) for b in bad[:3]]), flush=True)
            A = rng.standard_normal((R, K))
            B = rng.standard_normal((N, K))
            ref = A @ B.T
            got = split_gemm(A, B, S)
            print("real R=%d N=%d K=%d S=%d  rel err %.3e (256^-S = %.1e)" % (
                R, N, K, S, np.abs(got - ref).max() / np.abs(ref).max(), 256.0 ** -S), flush=True)
    # triangular modes
    n = 500
    Lm = np.tril(rng.standard_normal((n, n)))
    A = rng.standard_normal((256, n))
    for S in (4, 5):
        got = split_gemm(A, Lm, S, tri=1)           # out[i,k] = sum_{b<=k} A[i,b] L[k,b]
        ref = A @ Lm.T
        print("tri1 S=%d rel err %.3e" % (S, np.abs(got - ref).max() / np.abs(ref).max()), flush=True)
        got = split_gemm(A, Lm.T.copy(), S, tri=2)  # out[i,b] = sum_{k>=b} A[i,k] L[k,b]
        ref = A @ Lm
        print("tri2 S=%d rel err %.3e" % (S, np.abs(got - ref).max() / np.abs(ref).max()), flush=True)


if __name__ == "__main__":
    main()
