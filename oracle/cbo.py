"""Oracle side of cbo._current_marginal_argmax (cbo.py:121-235): the reference's literal per-candidate Python loops.

TEST INFRASTRUCTURE ONLY -- never imported by the product (bocf_b200/), which evaluates these three objectives on the
device (multi_outputGP.expected_utility).  The loop plumbing around it (run_optimization, evaluators, optimiser) is the
host code of bocf_b200.cbo, which has no arithmetic of its own; only the marginal-argmax objectives are restated here:

  * utility.linear                 posterior-mean utility and gradient          cbo.py:126-168
  * expectation_utility given      closed-form psi and psi_gradient, per candidate   cbo.py:170-198
  * otherwise                      50 fresh base samples, U(theta, mu + sigma Z) and its pathwise gradient   cbo.py:203-231
  (none of the three sums is normalised by the number of samples or hyper-samples, cbo.py:171)
"""
import numpy as np

from bocf_b200.cbo import CBO as _HostLoop


class CBO(_HostLoop):
    """The CBO loop driven by CPU (oracle) models, with the reference's own marginal-argmax objectives."""

    def _current_marginal_argmax(self, parameter):
        model = self.model
        n_h = self._n_hyps()
        if self.utility.linear:
            def val_func(X):
                X = np.atleast_2d(X)
                valX = np.zeros((X.shape[0], 1))
                for h in range(n_h):
                    model.set_hyperparameters(h)
                    muX = model.posterior_mean(X)
                    valX += np.reshape(np.matmul(np.atleast_1d(parameter), muX), (X.shape[0], 1))
                return -valX

            def val_func_with_gradient(X):
                X = np.atleast_2d(X)
                valX = np.zeros((X.shape[0], 1))
                dval_dX = np.zeros(X.shape)
                for h in range(n_h):
                    model.set_hyperparameters(h)
                    muX = model.posterior_mean(X)
                    dmu_dX = model.posterior_mean_gradient(X)
                    valX += np.reshape(np.matmul(np.atleast_1d(parameter), muX), (X.shape[0], 1))
                    dval_dX += np.tensordot(np.atleast_1d(parameter), dmu_dX, axes=1)
                return -valX, -dval_dX
        elif self.expectation_utility is not None:
            def val_func(X):
                X = np.atleast_2d(X)
                func_val = np.zeros((X.shape[0], 1))
                for h in range(n_h):
                    model.set_hyperparameters(h)
                    mean, var = model.predict_noiseless(X)
                    for i in range(X.shape[0]):
                        func_val[i, 0] += self.expectation_utility.func(parameter, mean[:, i], var[:, i])
                return -func_val

            def val_func_with_gradient(X):
                X = np.atleast_2d(X)
                func_val = np.zeros((X.shape[0], 1))
                func_gradient = np.zeros(X.shape)
                for h in range(n_h):
                    model.set_hyperparameters(h)
                    mean, var = model.predict_noiseless(X)
                    dmean_dX = model.posterior_mean_gradient(X)
                    dvar_dX = model.posterior_variance_gradient(X)
                    aux = np.concatenate((dmean_dX, dvar_dX))
                    for i in range(X.shape[0]):
                        func_val[i, 0] += self.expectation_utility.func(parameter, mean[:, i], var[:, i])
                        func_gradient[i, :] += np.matmul(self.expectation_utility.gradient(parameter, mean[:, i], var[:, i]),
                                                         aux[:, i])
                return -func_val, -func_gradient
        else:
            Z_samples = np.random.normal(size=(50, self.n_attributes))

            def val_func(X):
                X = np.atleast_2d(X)
                func_val = np.zeros((X.shape[0], 1))
                for h in range(n_h):
                    model.set_hyperparameters(h)
                    mean, var = model.predict_noiseless(X)
                    std = np.sqrt(var)
                    for Z in Z_samples:
                        func_val[:, 0] += np.asarray(self.utility.eval_func(parameter, mean + std * Z[:, None])).reshape(-1)
                return -func_val

            def val_func_with_gradient(X):
                X = np.atleast_2d(X)
                func_val = np.zeros((X.shape[0], 1))
                func_gradient = np.zeros(X.shape)
                for h in range(n_h):
                    model.set_hyperparameters(h)
                    mean, var = model.predict_noiseless(X)
                    std = np.sqrt(var)
                    dmean_dX = model.posterior_mean_gradient(X)
                    dstd_dX = model.posterior_variance_gradient(X) / (2 * std[:, :, None])
                    for i in range(X.shape[0]):
                        for Z in Z_samples:
                            aux1 = mean[:, i] + np.multiply(Z, std[:, i])
                            func_val[i, 0] += self.utility.eval_func(parameter, aux1)
                            aux2 = dmean_dX[:, i, :] + np.multiply(dstd_dX[:, i, :].T, Z).T
                            func_gradient[i, :] += np.matmul(self.utility.eval_gradient(parameter, aux1), aux2)
                return -func_val, -func_gradient

        argmax = self.evaluation_optimizer.optimize(f=val_func, f_df=val_func_with_gradient, parallel=False)[0]
        self.current_argmax = argmax
        return argmax
