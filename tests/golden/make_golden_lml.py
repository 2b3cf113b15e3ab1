"""Golden vectors for the marginal likelihood and its hyper-parameter gradients, produced by
the REFERENCE's own code (GP.parameters_changed -> ExactGaussianInference.inference, Stationary / SE
update_gradients_full, the C routine _lengthscale_grads, Gaussian.update_gradients) through tests/golden/ref_harness.py.

    python tests/golden/make_golden_lml.py        (needs /root/reference; writes tests/golden/lml_<kind>.npz)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402


def main():
    ns = rh.install()
    for kind in ("se", "rbf", "matern52", "matern32"):
        rng = np.random.default_rng({"se": 11, "rbf": 12, "matern52": 13, "matern32": 14}[kind])
        n, d, m, H = 45, 3, 2, 2
        X = rng.uniform(size=(n, d))
        Y = [np.sin(3.0 * X[:, :1] + j) + X[:, 1:2] * X[:, 2:3] + 0.05 * rng.standard_normal((n, 1)) + 0.2 * j
             for j in range(m)]
        variance = rng.uniform(0.6, 1.8, size=(H, m))
        lengthscale = rng.uniform(0.25, 0.9, size=(H, m, d))
        noise = rng.uniform(5e-3, 5e-2, size=(H, m))
        model = rh.make_reference_model(ns, kind, X, Y, variance, lengthscale, noise)
        lml = np.zeros((H, m))
        g_var = np.zeros((H, m))
        g_len = np.zeros((H, m, d))
        g_noise = np.zeros((H, m))
        for j in range(m):
            for h in range(H):
                g = model.output[j].model_instances[h]
                lml[h, j] = float(g._log_marginal_likelihood)
                g_var[h, j] = float(np.asarray(g.kern.variance.gradient).reshape(-1)[0])
                g_len[h, j] = np.asarray(g.kern.lengthscale.gradient, dtype=float).reshape(-1)
                g_noise[h, j] = float(np.asarray(g.likelihood.variance.gradient).reshape(-1)[0])
        out = os.path.join(HERE, "lml_%s.npz" % kind)
        np.savez(out, kind=kind, X=X, Y=np.concatenate(Y, axis=1), variance=variance, lengthscale=lengthscale, noise=noise,
                 lml=lml, g_var=g_var, g_len=g_len, g_noise=g_noise)
        print(out, lml.ravel())


if __name__ == "__main__":
    main()
