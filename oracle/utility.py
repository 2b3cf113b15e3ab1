"""Oracle restatement of utility.py / parameter_distribution.py and the composite catalogue.

Test infrastructure only.  The composites are the U(theta, y) / grad_y U pairs
defined in the reference's experiment scripts.
"""
import numpy as np


class Utility(object):
    # utility.py:6-48
    def __init__(self, func, dfunc=None, parameter_dist=None, linear=False):
        self.func = func
        self.dfunc = dfunc
        self.parameter_dist = parameter_dist
        self.linear = linear

    def eval_func(self, parameter, y):
        return self.func(parameter, y)

    def eval_gradient(self, parameter, y):
        return self.dfunc(parameter, y)


class ParameterDistribution(object):
    # parameter_distribution.py:5-29
    def __init__(self, continuous=False, support=None, prob_dist=None, sample_generator=None, rng=None):
        self.continuous = continuous
        self.support = support
        self.prob_dist = prob_dist
        self.sample_generator = sample_generator
        self.rng = np.random if rng is None else rng
        if support is not None and len(support) < 20:
            self.use_full_support = True
        else:
            self.use_full_support = False

    def sample(self, n_samples):
        if self.continuous:
            parameter_samples = self.sample_generator(n_samples)
        else:
            indices = self.rng.choice(int(len(self.support)), size=n_samples, p=self.prob_dist)
            parameter_samples = self.support[indices, :]
        return parameter_samples


# ---- composite catalogue ---------------------------------------------------------------------------

def sumsq_target_U(parameter, y):
    # test_1a.py:89-92, test_4a.py:84-87
    aux = (y.transpose() - parameter).transpose()
    return -np.sum(np.square(aux), axis=0)


def sumsq_target_dU(parameter, y):
    # test_1a.py:94-96
    y_aux = np.squeeze(y)
    return -2 * (y_aux - parameter)


def neg_sum_exp_U(parameter, y):
    # test_2a.py:60-62
    aux = -np.exp(y)
    return np.sum(aux, axis=0)


def neg_sum_exp_dU(parameter, y):
    # test_2a.py:64-65
    return -np.exp(y)


EXP_COS_C = np.array([1., 2., 5., 2., 3.])


def exp_cos_c(m):
    """c of test_3a.py:54 (m = 5 there); cycled for other m so the composite stays defined."""
    return EXP_COS_C[np.arange(m) % 5]


def exp_cos_U(parameter, y):
    # test_3a.py:53-58 (c is hard-wired to 5 attributes; np.dot contracts the attribute axis)
    aux = np.multiply(np.exp(-y / np.pi), np.cos(np.pi * y))
    return -np.tensordot(exp_cos_c(aux.shape[0]), aux, axes=1)


def exp_cos_dU(parameter, y):
    # test_3a.py:61-67
    y_copy = np.squeeze(y)
    aux = -np.pi * np.multiply(np.exp(-y_copy / np.pi), np.sin(np.pi * y_copy)) \
        - np.multiply(np.exp(-y_copy / np.pi), np.cos(np.pi * y_copy)) / np.pi
    return -np.multiply(exp_cos_c(len(y_copy)), aux)


def rosen_composite_U(a, y):
    # test_5a.py:48-52  (m = 2(dd-1); here dd-1 = m/2)
    h = y.shape[0] // 2
    a = np.asarray(a, dtype=float).reshape(-1)[0]
    val = 0
    for j in range(h):
        val = val - ((a - y[j])**2 + 100 * y[j + h]**2)
    return val


def rosen_composite_dU(a, y):
    # test_5a.py:54-59
    m = y.shape[0]
    h = m // 2
    a = np.asarray(a, dtype=float).reshape(-1)[0]
    gradient = np.empty((m,))
    for j in range(h):
        gradient[j] = 2 * (a - y[j])
        gradient[j + h] = -200 * y[j + h]
    return gradient


def linear_U(parameter, y):
    # test_1b.py:89-90
    return np.dot(parameter, y)


def linear_dU(parameter, y):
    # test_1b.py:92-93
    return parameter


COMPOSITES = {
    'sumsq_target': (sumsq_target_U, sumsq_target_dU, False),
    'neg_sum_exp': (neg_sum_exp_U, neg_sum_exp_dU, False),
    'exp_cos': (exp_cos_U, exp_cos_dU, False),
    'rosen_composite': (rosen_composite_U, rosen_composite_dU, False),
    'linear': (linear_U, linear_dU, True),
}


def make_utility(name, parameter_dist):
    U, dU, linear = COMPOSITES[name]
    ut = Utility(func=U, dfunc=dU, parameter_dist=parameter_dist, linear=linear)
    ut.composite_name = name
    return ut


def eval_gradient_batch(name, parameter, a):
    """grad_y U(theta, a[:, k]) for every column k of a (m, K): the same expressions as the *_dU functions
    above with the attribute axis kept explicit (used only by the vectorised twin of the reference loops)."""
    m = a.shape[0]
    if name == 'sumsq_target':
        return -2 * (a - np.asarray(parameter, dtype=float).reshape(-1)[:, None])
    if name == 'neg_sum_exp':
        return -np.exp(a)
    if name == 'exp_cos':
        aux = -np.pi * np.multiply(np.exp(-a / np.pi), np.sin(np.pi * a)) \
            - np.multiply(np.exp(-a / np.pi), np.cos(np.pi * a)) / np.pi
        return -exp_cos_c(m)[:, None] * aux
    if name == 'rosen_composite':
        h = m // 2
        th = np.asarray(parameter, dtype=float).reshape(-1)[0]
        g = np.zeros_like(a)
        g[:h] = 2 * (th - a[:h])
        g[h:2 * h] = -200 * a[h:2 * h]
        return g
    if name == 'linear':
        return np.broadcast_to(np.asarray(parameter, dtype=float).reshape(-1)[:, None], a.shape)
    raise ValueError(name)
