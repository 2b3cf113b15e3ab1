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


SCHEMES = {3: (3, 3, 1), 4: (4, 4, 2), 5: (5, 5, 4), 6: (6, 6, 5)}     # "splitN" -> (SA, SB, LMIN), split_gemm.cu


def split_matmul(A, eA, B, eB, S, scheme=None):
    """A (M,K) B (N,K) -> A B^T through the digit planes of scheme (SA, SB, LMIN): pairs with ta + tb < LMIN are
    dropped like the kernel does.  `S` alone selects the scheme the library maps "splitS" to."""
    SA, SB, LMIN = scheme if scheme is not None else SCHEMES[S]
    Ad, Bd = digit_planes(A, eA[:, None], SA), digit_planes(B, eB[:, None], SB)
    T = np.zeros((A.shape[0], B.shape[0]))
    for lev in range(SA + SB - 2, LMIN - 1, -1):           # Horner over the weight levels, most significant first
        C = np.zeros_like(T)
        for ta in range(SA):
            tb = lev - ta
            if 0 <= tb < SB:
                C += Ad[ta] @ Bd[tb].T                    # exact: |C| < 2^29 << 2^53
        T = T * 256.0 + C
    return np.ldexp(T, (eA[:, None] + eB[None, :]) - (8 * SA - 2) - (8 * SB - 2) + 8 * LMIN)


def test_digit_planes_are_exact_for_integers():
    rng = np.random.default_rng(0)
    A = rng.integers(-100, 101, size=(40, 70)).astype(float)
    B = rng.integers(-100, 101, size=(30, 70)).astype(float)
    for S in (3, 4, 5, 6):
        assert np.array_equal(split_matmul(A, row_exp(A), B, row_exp(B), S), A @ B.T)


def _posterior_pieces(kind, noise, n=200, d=6, N=256, seed=5):
    P = make_problem(m=1, d=d, n=n, H=1, kind=kind, N=N, S=4, noise=noise, seed=seed)
    var_f, ls = P.variance[0, 0], P.lengthscale[0, 0]
    K = _gen_kernel(kind, var_f, ls, P.X)
    K[np.diag_indices_from(K)] += noise + 1e-8
    Linv = sl.solve_triangular(np.linalg.cholesky(K), np.eye(P.n), lower=True)
    Xs, Xc = P.X / ls, P.Xc / ls
    diff = Xc[:, None, :] - Xs[None, :, :]
    r2 = np.maximum((diff ** 2).sum(-1), 0)
    r = np.sqrt(r2)
    if kind == "rbf":
        Ks = var_f * np.exp(-0.5 * r2)
        G = -Ks                                                           # dK/dr / r
    else:
        e = np.exp(-np.sqrt(5) * r)
        Ks = var_f * (1 + np.sqrt(5) * r + 5 / 3 * r2) * e
        G = var_f * (-5 / 3 - 5 * np.sqrt(5) / 3 * r) * e
    return P, var_f, ls, Linv, Ks, G, diff


# per-scheme constant c of the bound  rel. variance error <= c * max|L^-1|^2 * 256^-SA  that api.cu's AUTO / MIXED rules assume
VAR_BOUND = {4: 150.0, 5: 2000.0, 6: 2000.0}


@pytest.mark.parametrize("kind,noise", [("matern52", 1e-2), ("rbf", 1e-2), ("rbf", 1e-4)])
def test_variance_error_model(kind, noise):
    P, var_f, ls, Linv, Ks, G, diff = _posterior_pieces(kind, noise)
    V = Ks @ Linv.T
    var = var_f + noise - (V ** 2).sum(1)
    eA = np.full(P.N, np.frexp(var_f * 1.02)[1], dtype=np.int64)          # K* <= sigma_f^2
    amax = np.abs(Linv).max()
    prev = None
    for S in (4, 5, 6):
        V2 = split_matmul(Ks, eA, Linv, row_exp(Linv), S)
        err = np.max(np.abs((var_f + noise - (V2 ** 2).sum(1)) - var) / var)
        bound = VAR_BOUND[S] * amax ** 2 * 256.0 ** (-S)
        assert err < bound + 1e-13, (S, err, bound)                      # the bound the AUTO rule assumes
        if bound <= 5e-7:
            assert err < 1e-6, (S, err)                                  # whatever AUTO accepts meets the fp64 bar
        if prev is not None and prev > 1e-12:
            assert err < prev / 5.0                                       # every scheme up buys an order of magnitude
        prev = err


@pytest.mark.parametrize("kind,noise,n,d", [("matern52", 1e-2, 400, 10), ("rbf", 1e-2, 200, 6), ("rbf", 1e-4, 200, 6)])
def test_variance_gradient_error_model(kind, noise, n, d):
    """Second contraction Wt = V Linv -> dvar = gradients_X(-2 Wt): W^-1 K* cancels heavily inside the gradient sum, so
    the DROPPED digit pairs (full-size low digits) dominate the error, not the quantisation: keeping ta + tb >= S - 1
    (round 1) loses two orders of magnitude against keeping one more weight level at the same number of planes."""
    P, var_f, ls, Linv, Ks, G, diff = _posterior_pieces(kind, noise, n=n, d=d)
    V = Ks @ Linv.T
    grad = lambda W: -2 * np.einsum("ib,ibq->iq", W * G, diff) / ls
    dvar = grad(V @ Linv)
    eV = np.full(P.N, np.frexp(np.sqrt(var_f) * 1.02)[1], dtype=np.int64)     # |V| <= sigma_f
    B2 = Linv.T.copy()
    err = {}
    for name, scheme in [("331", (3, 3, 1)), ("332", (3, 3, 2)), ("442", (4, 4, 2)), ("443", (4, 4, 3)), ("554", (5, 5, 4))]:
        d2 = grad(split_matmul(V, eV, B2, row_exp(B2), None, scheme=scheme))
        err[name] = np.max(np.abs(d2 - dvar)) / np.max(np.abs(dvar))
    assert err["442"] < 2e-7 and err["554"] < 2e-7, err           # default mode: gradient contraction on 442
    assert err["331"] < 5e-5, err                                   # mixed mode: 8 pairs, inside the 1e-4 bar
    assert err["331"] < err["443"] * 4 and err["442"] < err["443"] / 30, err     # pair selection beats plane count
    assert err["332"] > 20 * err["331"], err




# ---- candidates next to a training input (the points the acquisition optimiser converges to) -------------------------
# Absolute error of sum V^2 over sigma_f^2 is K_S * 256^-S whatever the conditioning, while `posterior_variance` (noise
# included, the quantity the north-star bar names) drops to the noise variance there: api.cu's AUTO rule asks
# K_S 256^-S sigma_f^2 / noise <= 2.5e-7, a quarter of the bar, so that the NOISELESS variance uEI_noiseless samples with
# (which has no floor: it falls below the noise where many neighbours average) also holds 1e-6 wherever it is >= noise / 4.
# Constants restated from apply_precision.
NEAR_K = {4: 60.0, 5: 1000.0, 6: 1200.0}


def auto_planes(amax, noise_to_signal, target=5e-7):
    """api.cu apply_precision (AUTO): first S in 4..6 with both bounds inside the target; 0 = stay on the fp64 engine."""
    a2 = amax * amax
    away = {4: 150.0 * a2 * 256.0 ** -4, 5: 2000.0 * a2 * 256.0 ** -5, 6: 2000.0 * a2 * 256.0 ** -6}
    for S in (4, 5, 6):
        if away[S] <= target and NEAR_K[S] * 256.0 ** -S / (0.25 * noise_to_signal) <= 2.0 * target:
            return S
    return 0


@pytest.mark.parametrize("kind,n,d", [("matern52", 400, 10), ("rbf", 200, 6), ("matern52", 220, 4)])
@pytest.mark.parametrize("noise", [1e-1, 1e-2, 1e-3, 1e-4, 1e-5, 1e-6])
def test_variance_error_next_to_training_inputs(kind, n, d, noise):
    P, var_f, ls, Linv, _, _, _ = _posterior_pieces(kind, noise, n=n, d=d)
    rng = np.random.default_rng(0)
    Xc = np.concatenate([P.X[:96] + 1e-4 * rng.standard_normal((96, d)), P.X[:32]])      # next to, and exactly at, the data
    r2 = (((Xc / ls)[:, None, :] - (P.X / ls)[None, :, :]) ** 2).sum(-1)
    r = np.sqrt(r2)
    Ks = var_f * np.exp(-0.5 * r2) if kind == "rbf" else var_f * (1 + np.sqrt(5) * r + 5 / 3 * r2) * np.exp(-np.sqrt(5) * r)
    V = Ks @ Linv.T
    ssq = (V ** 2).sum(1)
    var_noiseless = var_f - ssq                                   # what uEI_noiseless samples with (clipped at 1e-10)
    assert var_noiseless.min() < 1.2 * noise * var_f + 1e-6       # the candidates ARE where the variance collapses
    eA = np.full(len(Xc), np.frexp(var_f * 1.02)[1], dtype=np.int64)
    for S in (4, 5, 6):
        ssq_S = (split_matmul(Ks, eA, Linv, row_exp(Linv), S) ** 2).sum(1)
        assert np.max(np.abs(ssq_S - ssq)) / var_f < NEAR_K[S] * 256.0 ** -S, S      # the constant of the rule
    S = auto_planes(np.abs(Linv).max(), (noise + 1e-8) / var_f)
    if noise >= 1e-2:
        assert S in (4, 5)                                        # well conditioned: tensor-core path, 13 or 15 pairs
    if noise == 1e-2 and n >= 400:
        assert S == 5                                             # the benchmark configurations keep 554
    if noise <= 1e-6:
        assert S == 0                                             # nothing on int8 planes holds 1e-6 next to the data
    if S:
        ssq_S = (split_matmul(Ks, eA, Linv, row_exp(Linv), S) ** 2).sum(1)
        v_ref, v_S = var_noiseless + noise, var_f + noise - ssq_S
        assert np.max(np.abs(v_S - v_ref) / v_ref) < 2.5e-7, (S, noise)             # posterior_variance, element-wise
        big = var_noiseless >= 0.25 * noise * var_f
        assert np.max(np.abs(ssq_S - ssq)[big] / var_noiseless[big]) < 1e-6, (S, noise)    # noiseless, where it has a floor
