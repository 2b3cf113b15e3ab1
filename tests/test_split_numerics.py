"""CPU model of the split-integer contraction (bocf_b200/csrc/split_gemm.cu): balanced base-256 digit planes, exact integer
pair products, pairs with ta + tb >= S - 1 kept.  Pins the error model the AUTO precision rule relies on without
needing a GPU: the relative variance error stays below 2000 * max|L^-1|^2 * 256^-S (measured 240 ... 1350 across kernels,
sizes and noise levels), and api.cu accepts S only when 400 * max|L^-1|^2 * 256^-S <= 1e-7, i.e. when that bound is
<= 5e-7 -- a factor 2 inside the north-star fp64 bar of 1e-6."""
import numpy as np
import pytest
import scipy.linalg as sl

from tests.helpers import make_problem, _gen_kernel


def digit_planes(x, e, S):
    """Most-significant-last planes d_t in [-128, 127] with x ~= 2^(e - (8S-2)) * sum_t d_t 256^t  (row scale 2^e)."""
    Y = np.rint(np.ldexp(x, 8 * S - 2 - e)).astype(np.int64)
    out = []
    for _ in range(S):
        d = ((Y + 128) & 255) - 128
        out.append(d.astype(np.float64))
        Y = (Y - d) >> 8
    assert np.all(Y == 0)
    return out


def row_exp(x):
    _, e = np.frexp(np.abs(x).max(axis=1) * 1.02)
    return e.astype(np.int64)


def split_matmul(A, eA, B, eB, S):
    """A (M,K) B (N,K) -> A B^T through the digit planes; pairs below level 0 are dropped like the kernel does."""
    Ad, Bd = digit_planes(A, eA[:, None], S), digit_planes(B, eB[:, None], S)
    T = np.zeros((A.shape[0], B.shape[0]))
    for lb in range(S - 1, -1, -1):                       # Horner over the weight levels, most significant first
        C = np.zeros_like(T)
        for ta in range(S):
            tb = S - 1 + lb - ta
            if 0 <= tb < S:
                C += Ad[ta] @ Bd[tb].T                    # exact: |C| < 2^29 << 2^53
        T = T * 256.0 + C
    return np.ldexp(T, (eA[:, None] + eB[None, :]) - 2 * (8 * S - 2) + 8 * (S - 1))


def test_digit_planes_are_exact_for_integers():
    rng = np.random.default_rng(0)
    A = rng.integers(-100, 101, size=(40, 70)).astype(float)
    B = rng.integers(-100, 101, size=(30, 70)).astype(float)
    for S in (3, 4, 5, 6):
        assert np.array_equal(split_matmul(A, row_exp(A), B, row_exp(B), S), A @ B.T)


@pytest.mark.parametrize("kind,noise", [("matern52", 1e-2), ("rbf", 1e-2), ("rbf", 1e-4)])
def test_variance_error_model(kind, noise):
    P = make_problem(m=1, d=6, n=200, H=1, kind=kind, N=256, S=4, noise=noise, seed=5)
    var_f, ls = P.variance[0, 0], P.lengthscale[0, 0]
    K = _gen_kernel(kind, var_f, ls, P.X)
    K[np.diag_indices_from(K)] += noise + 1e-8
    Linv = sl.solve_triangular(np.linalg.cholesky(K), np.eye(P.n), lower=True)
    Xs, Xc = P.X / ls, P.Xc / ls
    r2 = np.maximum(((Xc[:, None, :] - Xs[None, :, :]) ** 2).sum(-1), 0)
    r = np.sqrt(r2)
    Ks = var_f * np.exp(-0.5 * r2) if kind == "rbf" else var_f * (1 + np.sqrt(5) * r + 5 / 3 * r2) * np.exp(-np.sqrt(5) * r)
    V = Ks @ Linv.T
    var = var_f + noise - (V ** 2).sum(1)
    eA = np.full(P.N, np.frexp(var_f * 1.02)[1], dtype=np.int64)          # K* <= sigma_f^2
    amax = np.abs(Linv).max()
    prev = None
    for S in (4, 5, 6):
        V2 = split_matmul(Ks, eA, Linv, row_exp(Linv), S)
        err = np.max(np.abs((var_f + noise - (V2 ** 2).sum(1)) - var) / var)
        assert err < 2000.0 * amax ** 2 * 256.0 ** (-S) + 1e-13, (S, err)   # the bound the AUTO rule assumes
        if 400.0 * amax ** 2 * 256.0 ** (-S) <= 1e-7:
            assert err < 1e-6, (S, err)                                  # whatever AUTO accepts meets the fp64 bar
        if prev is not None and prev > 1e-12:
            assert err < prev / 30.0                                      # every plane buys ~256x
        prev = err
