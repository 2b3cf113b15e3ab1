#!/usr/bin/env python
"""Golden vectors for the knowledge-gradient posterior helpers, produced by the REFERENCE's own classes
(multi_outputGP.py:203-281,309-331 -> GPModel, gpmodel.py:187-287 -> GP, GPy/core/gp.py:493-627) through ref_harness.

    make -C oracle && python tests/golden/make_golden_kg.py

The reference's posterior_covariance_gradient stacks an (N, 1, d) block into an (N, d) slot (multi_outputGP.py:317), which
numpy only accepts for N = 1 -- the way the fork's KG acquisitions call it -- so the gradient goldens are produced one
candidate at a time.  kern.gradients_X(None, ...) exists only for the fork's SE kernel (se.py:142-144): the covariance
gradient cases use SE; the conditioned-variance cases also cover Matern-5/2.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.golden import ref_harness as rh      # noqa: E402
from tests.helpers import make_problem          # noqa: E402

CASES = [("kg_se", dict(m=3, d=3, n=25, H=2, kind="se", N=14, S=4, seed=201), True),
         ("kg_matern52", dict(m=2, d=4, n=30, H=1, kind="matern52", N=12, S=4, seed=202), False)]


def main():
    ns = rh.install()
    for name, kw, with_cov_grad in CASES:
        P = make_problem(**kw)
        model = rh.make_reference_model(ns, P.kind, P.X, P.Y, P.variance, P.lengthscale, P.noise)
        X2 = P.Xc[:3].copy()
        x_next = P.Xc[5:6].copy()
        out = dict(X=P.X, Y=np.concatenate(P.Y, axis=1).T, variance=P.variance, lengthscale=P.lengthscale, noise=P.noise,
                   Xc=P.Xc, X2=X2, x_next=x_next, kind=P.kind)
        for h in range(P.H):
            model.set_hyperparameters(h)
            out["cov_h%d" % h] = model.posterior_covariance_between_points(P.Xc, X2)
            model.partial_precomputation_for_covariance(X2)
            pp = model.posterior_covariance_between_points_partially_precomputed(P.Xc, X2)
            assert np.allclose(pp, out["cov_h%d" % h], rtol=1e-9, atol=1e-12)
            model.partial_precomputation_for_variance_conditioned_on_next_point(x_next)
            out["varcond_h%d" % h] = model.posterior_variance_conditioned_on_next_point(P.Xc)
            out["dvarcond_h%d" % h] = model.posterior_variance_gradient_conditioned_on_next_point(P.Xc)
            if with_cov_grad:
                x2 = X2[1:2]
                model.partial_precomputation_for_covariance_gradient(x2)
                g = np.empty((P.m, P.N, P.d))
                for i in range(P.N):
                    gi = model.posterior_covariance_gradient(P.Xc[i:i + 1], x2)
                    gp = model.posterior_covariance_gradient_partially_precomputed(P.Xc[i:i + 1], x2)
                    assert np.allclose(gi, gp, rtol=1e-9, atol=1e-12)
                    g[:, i, :] = gi[:, 0, :]
                out["dcov_h%d" % h] = g
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "->", os.path.relpath(path, ROOT), {k: v.shape for k, v in out.items() if hasattr(v, "shape") and k.startswith(("cov", "var", "dvar", "dcov"))})


if __name__ == "__main__":
    main()
