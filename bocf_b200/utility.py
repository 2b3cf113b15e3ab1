"""Utility / ParameterDistribution with the reference's interface (utility.py:6-48,
parameter_distribution.py:5-29).

The reference passes arbitrary Python callables ``func`` / ``dfunc``; the device path needs the
composite as one of the catalogued device functions instead, selected by ``composite=`` (the
U(theta, y) found in the reference's experiment scripts).  ``func``/``dfunc``
stay available as host callables for the parts of the BO loop that evaluate the true objective.
"""
import numpy as np

from ._lib import COMPOSITES


def _host_funcs(name):
    if name == "sumsq_target":       # test_1a.py:89-96
        return (lambda th, y: -np.sum(np.square((np.asarray(y).T - th).T), axis=0),
                lambda th, y: -2 * (np.squeeze(y) - th))
    if name == "neg_sum_exp":        # test_2a.py:60-65
        return (lambda th, y: np.sum(-np.exp(y), axis=0), lambda th, y: -np.exp(y))
    if name == "exp_cos":            # test_3a.py:53-67
        c = np.array([1., 2., 5., 2., 3.])

        def U(th, y):
            y = np.asarray(y)
            return -np.tensordot(c[np.arange(y.shape[0]) % 5], np.exp(-y / np.pi) * np.cos(np.pi * y), axes=1)

        def dU(th, y):
            y = np.squeeze(y)
            aux = -np.pi * np.exp(-y / np.pi) * np.sin(np.pi * y) - np.exp(-y / np.pi) * np.cos(np.pi * y) / np.pi
            return -c[np.arange(len(y)) % 5] * aux
        return U, dU
    if name == "rosen_composite":    # test_5a.py:48-59
        def U(a, y):
            y = np.asarray(y)
            h = y.shape[0] // 2
            a = float(np.asarray(a).reshape(-1)[0])
            return -np.sum((a - y[:h]) ** 2 + 100 * y[h:2 * h] ** 2, axis=0)

        def dU(a, y):
            y = np.squeeze(y)
            h = len(y) // 2
            a = float(np.asarray(a).reshape(-1)[0])
            g = np.zeros(len(y))
            g[:h] = 2 * (a - y[:h])
            g[h:2 * h] = -200 * y[h:2 * h]
            return g
        return U, dU
    if name == "linear":             # test_1b.py:89-93
        return (lambda th, y: np.dot(th, y), lambda th, y: th)
    raise ValueError("unknown composite %r (known: %s)" % (name, sorted(COMPOSITES)))


# length of one utility parameter: "m" = one entry per output, 1 = a scalar
THETA_DIM = {"sumsq_target": "m", "linear": "m", "neg_sum_exp": 1, "exp_cos": 1, "rosen_composite": 1}


def theta_matrix(utility, samples):
    """Utility parameters as an (L, p) matrix.  A 1-D support is L scalar parameters for the composites that take a
    scalar (test_5a.py:41: np.atleast_1d([1.]) -- the reference iterates len(support)), one vector otherwise."""
    th = np.asarray(samples, dtype=np.float64)
    if th.ndim == 0:
        th = th.reshape(1, 1)
    elif th.ndim == 1:
        th = th.reshape(-1, 1) if THETA_DIM.get(utility.composite) == 1 else th.reshape(1, -1)
    return np.ascontiguousarray(th)


def _host_psi(name):
    """Closed-form E[U(theta, y)], y_j ~ N(mu_j, v_j), as (psi, dpsi/d(mu, v)) -- the pairs the reference's scripts pass
    as ExpectationUtility; the device enum of acq.cu (psi_acq_kernel) restates the same formulas."""
    if name == "sumsq_target":       # test_1a.py:100-113
        return (lambda th, mu, v: -np.sum(np.square((np.asarray(mu).T - th).T), axis=0) - np.sum(v, axis=0),
                lambda th, mu, v: -np.concatenate((2 * (np.squeeze(mu) - th), np.ones(len(np.squeeze(v))))))
    if name == "neg_sum_exp":        # test_2a.py:70-83
        return (lambda th, mu, v: -np.sum(np.exp(np.squeeze(mu) + 0.5 * np.squeeze(v))),
                lambda th, mu, v: -np.concatenate((np.exp(np.squeeze(mu) + 0.5 * np.squeeze(v)),
                                                   0.5 * np.exp(np.squeeze(mu) + 0.5 * np.squeeze(v)))))
    if name == "rosen_composite":    # test_5a.py:64-77
        def psi(a, mu, v):
            mu, v = np.squeeze(mu), np.squeeze(v)
            h = len(mu) // 2
            a = float(np.asarray(a).reshape(-1)[0])
            return -np.sum((a - mu[:h]) ** 2 + 100 * mu[h:2 * h] ** 2 + v[:h] + 100 * v[h:2 * h])

        def dpsi(a, mu, v):
            mu = np.squeeze(mu)
            m, h = len(mu), len(mu) // 2
            a = float(np.asarray(a).reshape(-1)[0])
            g = np.zeros(2 * m)
            g[:h] = 2 * (a - mu[:h])
            g[h:2 * h] = -200 * mu[h:2 * h]
            g[m:m + h] = -1.0
            g[m + h:m + 2 * h] = -100.0
            return g
        return psi, dpsi
    if name == "linear":             # cbo.py:126-168: E[theta . y] = theta . mu
        return (lambda th, mu, v: np.dot(th, np.squeeze(mu)),
                lambda th, mu, v: np.concatenate((np.atleast_1d(th), np.zeros(len(np.squeeze(mu))))))
    return None


class Utility(object):
    """utility.py:6-48 plus ``composite``: the name of the device-side U(theta, y)."""

    def __init__(self, func=None, dfunc=None, parameter_dist=None, linear=False, composite=None):
        if composite is None:
            raise ValueError(
                "bocf_b200.Utility needs composite=<one of %s>: arbitrary Python callables cannot run inside the "
                "CUDA kernels (and there is no CPU fallback)." % sorted(COMPOSITES))
        if composite not in COMPOSITES:
            raise ValueError("unknown composite %r" % (composite,))
        self.composite = composite
        hf, hdf = _host_funcs(composite)
        self.func = func if func is not None else hf
        self.dfunc = dfunc if dfunc is not None else hdf
        self.parameter_dist = parameter_dist
        self.linear = linear or composite == "linear"
        if func is not None or dfunc is not None:
            self._check_callables(hf, hdf)

    def _check_callables(self, hf, hdf, m=4):
        """A user-supplied func / dfunc must be the catalogued composite: the device evaluates the enum, the loop's
        reporting path the callable -- two different utilities would silently optimise one and report the other."""
        rng = np.random.default_rng(0)
        y = rng.standard_normal(m)
        th = rng.standard_normal(m) if THETA_DIM[self.composite] == "m" else np.array([0.7])
        try:
            ok = np.allclose(np.asarray(self.func(th, y), dtype=float), np.asarray(hf(th, y), dtype=float), rtol=1e-9, atol=1e-12)
            if self.dfunc is not None:
                ok = ok and np.allclose(np.asarray(self.dfunc(th, y), dtype=float).reshape(-1),
                                        np.asarray(hdf(th, y), dtype=float).reshape(-1), rtol=1e-9, atol=1e-12)
        except Exception:            # callables written for one fixed m (test_5a.py) cannot be probed at m = 4
            return
        if not ok:
            raise ValueError("Utility: func / dfunc disagree with composite=%r (the CUDA kernels evaluate the composite)"
                             % (self.composite,))

    def evaluate_w_gradient(self, parameter, y):
        return self.eval_func(parameter, y), self.eval_gradient(parameter, y)

    def eval_func(self, parameter, y):
        return self.func(parameter, y)

    def eval_gradient(self, parameter, y):
        return self.dfunc(parameter, y)


class ParameterDistribution(object):
    """parameter_distribution.py:5-29 (discrete support + probabilities, or a continuous sampler)."""

    def __init__(self, continuous=False, support=None, prob_dist=None, sample_generator=None):
        if continuous is True and sample_generator is None:
            pass
        else:
            self.continuous = continuous
            self.support = support
            self.prob_dist = prob_dist
            self.sample_generator = sample_generator
        if support is not None and len(support) < 20:
            self.use_full_support = True
        else:
            self.use_full_support = False

    def sample(self, n_samples):
        if self.continuous:
            parameter_samples = self.sample_generator(n_samples)
        else:
            indices = np.random.choice(int(len(self.support)), size=n_samples, p=self.prob_dist)
            parameter_samples = self.support[indices, :]
        return parameter_samples
