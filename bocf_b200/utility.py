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
