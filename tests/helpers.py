"""Shared test/bench helpers: seeded synthetic problems (SURVEY.md 8d recipe) and the two sides of a
parity check -- the CPU oracle and the CUDA product -- fed identical explicit inputs."""
import numpy as np

THETA_DIM = {"sumsq_target": "m", "linear": "m", "neg_sum_exp": 1, "exp_cos": 1, "rosen_composite": 1}


class Problem(object):
    pass


def _gen_kernel(kind, variance, ls, X):
    """Plain-numpy covariance used ONLY to draw synthetic observations (independent of oracle/ and of the product)."""
    Xs = X / ls
    r2 = np.maximum(((Xs[:, None, :] - Xs[None, :, :]) ** 2).sum(-1), 0.0)
    r = np.sqrt(r2)
    if kind in ("se", "rbf"):
        return variance * np.exp(-0.5 * r2)
    if kind == "matern52":
        return variance * (1 + np.sqrt(5.) * r + 5. / 3 * r2) * np.exp(-np.sqrt(5.) * r)
    return variance * (1 + np.sqrt(3.) * r) * np.exp(-np.sqrt(3.) * r)


def _gen_score(composite, theta, Y):
    """Utility of the observations (m, n) under theta; only ranks training points for candidate placement."""
    m = Y.shape[0]
    if composite == "sumsq_target":
        return -((Y - theta[:, None]) ** 2).sum(0)
    if composite == "neg_sum_exp":
        return -np.exp(Y).sum(0)
    if composite == "exp_cos":
        c = np.array([1., 2., 5., 2., 3.])[np.arange(m) % 5]
        return -(c[:, None] * np.exp(-Y / np.pi) * np.cos(np.pi * Y)).sum(0)
    if composite == "rosen_composite":
        h = m // 2
        return -((theta[0] - Y[:h]) ** 2 + 100 * Y[h:2 * h] ** 2).sum(0)
    return theta @ Y


def make_problem(m=4, d=6, n=200, H=1, kind="rbf", composite="sumsq_target", N=1024, S=256, L=1, seed=0,
                 noise=1e-2, prior_draw=True, focus=0.75, focus_scale=0.08):
    """Synthetic inputs of SURVEY.md 8(d): X ~ U[0,1]^{n x d}, Y_j a prior-GP draw + noise, lengthscales
    U[0.2,1]*sqrt(d)/2, sigma_f^2 = 1, sigma_n^2 = 1e-2, candidates U[0,1]^{N x d} (seed 7), Z ~ N(0,1) (seed 11).
    ``kind``: one kernel family, or a sequence with one family per output (multi_outputGP.py:23,38-44)."""
    P = Problem()
    rng = np.random.default_rng(seed)
    P.m, P.d, P.n, P.H, P.kind, P.composite, P.N, P.S, P.L = m, d, n, H, kind, composite, N, S, L
    P.X = rng.uniform(size=(n, d))
    P.variance = np.ones((H, m)) if H == 1 else rng.uniform(0.7, 1.4, size=(H, m))
    P.lengthscale = rng.uniform(0.2, 1.0, size=(H, m, d)) * np.sqrt(d) / 2
    P.noise = np.full((H, m), noise)
    P.Y = []
    for j in range(m):
        r = np.random.default_rng(1000 + j + 17 * seed)
        if prior_draw and n <= 2500:
            K = _gen_kernel(kind if isinstance(kind, str) else kind[j], P.variance[0, j], P.lengthscale[0, j], P.X)
            K[np.diag_indices_from(K)] += 1e-8
            Lc = np.linalg.cholesky(K)
            y = Lc @ r.standard_normal(n) + np.sqrt(noise) * r.standard_normal(n)
        else:
            # cheap smooth surrogate for large n (bench only): random Fourier features of the same lengthscales
            W = r.standard_normal((d, 64)) / P.lengthscale[0, j][:, None]
            b = r.uniform(0, 2 * np.pi, 64)
            y = np.sqrt(2.0 / 64) * np.cos(P.X @ W + b) @ r.standard_normal(64) + np.sqrt(noise) * r.standard_normal(n)
        P.Y.append(y.reshape(n, 1) + 0.3 * j)
    P.Xc = np.random.default_rng(7 + seed).uniform(size=(N, d))
    P.Z = np.random.default_rng(11 + seed).standard_normal((S, m))
    tr = np.random.default_rng(23 + seed)
    p = m if THETA_DIM[composite] == "m" else 1
    if composite == "sumsq_target":
        # targets = observations at random training points (test_1a.py:79 pattern), jittered
        idx = tr.integers(0, n, size=L)
        P.theta = np.stack([np.array([P.Y[j][i, 0] for j in range(m)]) for i in idx]) + 0.4 * tr.standard_normal((L, m))
    elif composite == "linear":
        P.theta = tr.dirichlet(np.ones(m), size=L)
    else:
        P.theta = np.ones((L, p))
    P.prob = tr.dirichlet(np.ones(L)) if L > 1 else np.ones(1)
    if focus > 0:
        # uniformly random candidates almost never improve on the incumbent; move a fraction of them next to the
        # best observed points so the parity checks see plenty of non-zero EI values and gradients
        Ymat = np.concatenate(P.Y, axis=1).T                       # (m, n)
        score = _gen_score(composite, P.theta[0], Ymat)
        top = np.argsort(-score)[:8]
        k = int(focus * N)
        cr = np.random.default_rng(31 + seed)
        P.Xc[:k] = np.clip(P.X[top[cr.integers(0, len(top), size=k)]] + focus_scale * cr.standard_normal((k, d)), 0, 1)
    return P


# ---- oracle side ------------------------------------------------------------------------------------------

def oracle_model(P):
    from oracle.models import multi_outputGP
    mod = multi_outputGP.from_hyper_samples(P.kind, P.variance, P.lengthscale, P.noise, ARD=True)
    mod.updateModel(P.X, P.Y)
    return mod


def oracle_utility(P, full_support=True):
    from oracle.utility import make_utility, ParameterDistribution
    pd = ParameterDistribution(support=P.theta, prob_dist=P.prob)
    return make_utility(P.composite, pd)


def oracle_acq(P, grad=True, variant="uEI_noiseless", vectorised=True, Xc=None, model=None, parallel=True):
    from oracle import acquisitions as A
    mod = oracle_model(P) if model is None else model
    U = oracle_utility(P)
    Xc = P.Xc if Xc is None else Xc
    mod.set_hyperparameters(0)
    if variant in ("uEI_noiseless", "uPI"):
        acq = getattr(A, variant)(mod, utility=U, W_samples=P.Z, vectorised=vectorised)
    else:
        acq = getattr(A, variant)(mod, utility=U)
    if grad:
        a, g = acq._compute_acq_withGradients(Xc)
        return a[:, 0], g
    if variant in ("uEI_noiseless", "uPI"):
        return acq._compute_acq(Xc, parallel=parallel)[:, 0], None
    return acq._compute_acq(Xc)[:, 0], None


# ---- product side -----------------------------------------------------------------------------------------

def product_model(P, device="cuda:0", precision=None):
    import bocf_b200
    mod = bocf_b200.multi_outputGP(P.m, n_samples=P.H, device=device, precision=precision)
    mod.set_hyperparameter_samples(P.variance, P.lengthscale, P.noise, kind=P.kind)
    mod.updateModel(P.X, P.Y)
    return mod


def product_utility(P):
    import bocf_b200
    pd = bocf_b200.ParameterDistribution(support=P.theta, prob_dist=P.prob)
    return bocf_b200.Utility(parameter_dist=pd, composite=P.composite)


def product_acq(P, grad=True, variant="uEI_noiseless", device="cuda:0", Xc=None, model=None, parallel=True,
                precision=None):
    import bocf_b200
    mod = product_model(P, device, precision) if model is None else model
    U = product_utility(P)
    Xc = P.Xc if Xc is None else Xc
    mod.set_hyperparameters(0)
    acq = getattr(bocf_b200, variant)(mod, None, utility=U)
    if variant in ("uEI_noiseless", "uPI"):
        acq.W_samples = P.Z
    if grad:
        a, g = acq._compute_acq_withGradients(Xc)
        return a[:, 0], g
    if variant in ("uEI_noiseless", "uPI"):
        return acq._compute_acq(Xc, parallel=parallel)[:, 0], None
    return acq._compute_acq(Xc)[:, 0], None


def tol(fp64_tol, floor=1e-6):
    """Tolerance of a GPU parity check: the fp64-tight value when the contractions run in fp64 (BOCF_PRECISION=fp64),
    otherwise at least `floor` -- the north-star fp64-mode bar of 1e-6 on mean / variance, which the split-integer
    tensor-core mode is also held to for everything derived from the variance (north star allows 1e-4 on EI-CF)."""
    import os
    return fp64_tol if os.environ.get("BOCF_PRECISION", "auto") == "fp64" else max(fp64_tol, floor)


def rel_err(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


def elem_err(a, b, floor=1e-2):
    """Element-wise error metric used next to the norm-wise `rel_err`:  max_k |a_k - b_k| / max(|b_k|, floor * max|b|).
    Entries above floor * max|b| are held to a RELATIVE bar, smaller ones to an absolute bar `floor` times tighter
    than the norm-wise one.  (A purely relative check has no meaning for EI values in the far tail, which are sums of
    a handful of U(theta, a_s) - f* differences that cancel to nothing.)"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.abs(b), floor * max(1e-300, np.max(np.abs(b))))
    return float(np.max(np.abs(a - b) / scale))


def assert_close(a, b, tol_, what="", floor=1e-2):
    """Norm-wise AND element-wise agreement (see elem_err) within tol_."""
    nw, ew = rel_err(a, b), elem_err(a, b, floor)
    assert nw < tol_ and ew < 4 * tol_, "%s: norm-wise %.3e, element-wise %.3e, tol %.1e" % (what, nw, ew, tol_)
