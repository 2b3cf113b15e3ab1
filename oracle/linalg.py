"""Oracle restatement of GPy/util/linalg.py (LAPACK wrappers).  Test infrastructure only.

Calls the same scipy LAPACK entry points the reference calls.
"""
import numpy as np
from scipy import linalg
from scipy.linalg import lapack, blas


class LinAlgError(linalg.LinAlgError):
    pass


def force_F_ordered(A):
    # GPy/util/linalg.py:28-37
    if A.flags['F_CONTIGUOUS']:
        return A
    return np.asfortranarray(A)


def jitchol(A, maxtries=5, return_jitter=False):
    """GPy/util/linalg.py:52-83.  dpotrf; on failure add mean(diag)*1e-6*10^k jitter, k=0..4."""
    A = np.ascontiguousarray(A)
    L, info = lapack.dpotrf(A, lower=1)
    if info == 0:
        return (L, 0.0) if return_jitter else L
    diagA = np.diag(A)
    if np.any(diagA <= 0.):
        raise linalg.LinAlgError("not pd: non-positive diagonal elements")
    jitter = diagA.mean() * 1e-6
    num_tries = 1
    while num_tries <= maxtries and np.isfinite(jitter):
        try:
            L = linalg.cholesky(A + np.eye(A.shape[0]) * jitter, lower=True)
            return (L, jitter) if return_jitter else L
        except Exception:
            jitter *= 10
        finally:
            num_tries += 1
    raise linalg.LinAlgError("not positive definite, even with jitter.")


def dtrtrs(A, B, lower=1, trans=0, unitdiag=0):
    # GPy/util/linalg.py:91-110
    A = np.asfortranarray(A)
    return lapack.dtrtrs(A, B, lower=lower, trans=trans, unitdiag=unitdiag)


def dpotrs(A, B, lower=1):
    # GPy/util/linalg.py:112-121
    A = force_F_ordered(A)
    return lapack.dpotrs(A, B, lower=lower)


def symmetrify(A, upper=False):
    # GPy/util/linalg.py:352-375 (numpy twin of the cython routine)
    triu = np.triu_indices_from(A, k=1)
    if upper:
        A.T[triu] = A[triu]
    else:
        A[triu] = A.T[triu]


def dpotri(A, lower=1):
    # GPy/util/linalg.py:123-141
    A = force_F_ordered(A)
    R, info = lapack.dpotri(A, lower=lower)
    symmetrify(R)
    return R, info


def dtrtri(L):
    # GPy/util/linalg.py:213-223
    L = force_F_ordered(L)
    return lapack.dtrtri(L, lower=1)[0]


def pdinv(A, *args):
    # GPy/util/linalg.py:189-210
    L = jitchol(A, *args)
    logdet = 2. * np.sum(np.log(np.diag(L)))
    Li = dtrtri(L)
    Ai, _ = dpotri(L, lower=1)
    symmetrify(Ai)
    return Ai, L, Li, logdet


def tdot(mat):
    # GPy/util/linalg.py:295-319 (dsyrk: mat mat^T), then symmetrify
    if mat.dtype != 'float64' or len(mat.shape) != 2:
        return np.dot(mat, mat.T)
    nn = mat.shape[0]
    out = np.zeros((nn, nn))
    mat = np.asfortranarray(mat)
    out = blas.dsyrk(alpha=1.0, a=mat, beta=0.0, c=out, overwrite_c=1, trans=0, lower=0)
    symmetrify(out, upper=True)
    return np.ascontiguousarray(out)
