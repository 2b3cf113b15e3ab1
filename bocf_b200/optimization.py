"""Host plumbing around the hot path: design space, random design, multistart acquisition optimiser.

Restates (for Python 3 / current numpy, without pathos) the pieces of the vendored GPyOpt the BOCF loop uses:
  Design_space / ContinuousVariable   GPyOpt/core/task/space.py:13-448, variables.py:84-120 (continuous only)
  initial_design('random', ...)       GPyOpt/experiment_design/random_design.py:56-77
  ObjectiveAnchorPointsGenerator      GPyOpt/optimization/anchor_points_generator.py:19-66,87-99
  OptLbfgs / OptLbfgs2                GPyOpt/optimization/optimizer.py:283-354
  apply_optimizer                     GPyOpt/optimization/optimizer.py:425-466
  AcquisitionOptimizer.optimize       GPyOpt/optimization/acquisition_optimizer.py:95-154
  GeneralOptimizer.optimize           GPyOpt/optimization/general_optimizer.py:53-94
  Sequential evaluator                GPyOpt/core/evaluators/sequential.py:7-23
These are CALLERS of the hot path and stay on the host.  The scoring call f(X) over all random candidates is one
device sweep; the per-anchor L-BFGS-B runs keep scipy's state machine but their f_df evaluations are batched into one
device launch per round (`BatchedEvaluations`) instead of the reference's `Pool(4).map` over anchors
(acquisition_optimizer.py:131-133) -- process pools cannot share a CUDA context.
"""
import threading

import numpy as np
import scipy.optimize


# ---- design space ------------------------------------------------------------------------------------------
class Design_space(object):
    """Continuous box domains only (every BOCF script uses a single continuous variable with `dimensionality`)."""

    def __init__(self, space, constraints=None):
        self.config_space = space
        bounds = []
        for var in space:
            if var.get('type', 'continuous') != 'continuous':
                raise NotImplementedError("only continuous variables are supported on this path")
            bounds += [tuple(var['domain'])] * int(var.get('dimensionality', 1))
        self._bounds = bounds
        self.dimensionality = len(bounds)
        self.model_dimensionality = len(bounds)
        self.constraints = constraints

    def get_bounds(self):
        return list(self._bounds)

    def get_continuous_bounds(self):
        return list(self._bounds)

    def input_dim(self):
        return self.dimensionality

    def zip_inputs(self, X):
        return X

    def unzip_inputs(self, X):
        return X

    def round_optimum(self, x):
        # space.py:322-343 + variables.py:100-116: clamp each coordinate to its domain
        x = np.array(x)
        if not ((x.ndim == 1) or (x.ndim == 2 and x.shape[0] == 1)):
            raise ValueError("Unexpected dimentionality of x. Got {}, expected (1, N) or (N,)".format(x.ndim))
        if x.ndim == 2:
            x = x[0]
        lo = np.array([b[0] for b in self._bounds])
        hi = np.array([b[1] for b in self._bounds])
        return np.atleast_2d(np.minimum(np.maximum(x, lo), hi))

    def indicator_constraints(self, x):
        return np.ones((np.atleast_2d(x).shape[0], 1))


def samples_multidimensional_uniform(bounds, points_count):
    # random_design.py:67-77
    dim = len(bounds)
    Z_rand = np.zeros(shape=(points_count, dim))
    for k in range(0, dim):
        Z_rand[:, k] = np.random.uniform(low=bounds[k][0], high=bounds[k][1], size=points_count)
    return Z_rand


def initial_design(design_name, space, init_points_count):
    # experiment_design/__init__.py:7-20 ('random' only; latin/sobol/grid need pyDOE/sobol_seq)
    if design_name != 'random':
        raise NotImplementedError("only the 'random' design is available")
    return samples_multidimensional_uniform(space.get_bounds(), init_points_count)


# ---- local optimisers -----------------------------------------------------------------------------------------
class OptLbfgs(object):
    """optimizer.py:283-316 (maxiter 500, factr 1e6)."""

    def __init__(self, bounds, maxiter=500):
        self.bounds = bounds
        self.maxiter = maxiter
        self.factr, self.pgtol = 1e6, 1e-5

    def optimize(self, x0, f=None, df=None, f_df=None):
        if f_df is None and df is not None:
            f_df = lambda x: (float(f(x)), df(x))     # noqa: E731
        if f_df is None and df is None:
            res = scipy.optimize.fmin_l_bfgs_b(lambda x: float(np.asarray(f(x)).reshape(-1)[0]), x0=x0, bounds=self.bounds,
                                               approx_grad=True, maxiter=self.maxiter, factr=1e3, pgtol=1e-20)
        else:
            res = scipy.optimize.fmin_l_bfgs_b(_scalar_f_df(f_df), x0=x0, bounds=self.bounds, maxiter=self.maxiter,
                                               factr=self.factr, pgtol=self.pgtol)
        if res[2]['task'] in (b'ABNORMAL_TERMINATION_IN_LNSRCH', 'ABNORMAL_TERMINATION_IN_LNSRCH', 'ABNORMAL'):
            result_x = np.atleast_2d(x0)
            result_fx = np.atleast_2d(f(x0))
        else:
            result_x = np.atleast_2d(res[0])
            result_fx = np.atleast_2d(res[1])
        return result_x, result_fx


class OptLbfgs2(OptLbfgs):
    """optimizer.py:319-354 (maxiter 50, factr 1e5, pgtol 1e-15)."""

    def __init__(self, bounds, maxiter=50):
        super(OptLbfgs2, self).__init__(bounds, maxiter)
        self.factr, self.pgtol = 1e5, 1e-15


def _scalar_f_df(f_df):
    def wrapped(x):
        fx, dfx = f_df(np.atleast_2d(x))
        return float(np.asarray(fx).reshape(-1)[0]), np.asarray(dfx, dtype=np.float64).reshape(-1)
    return wrapped


def choose_optimizer(optimizer_name, bounds):
    # optimizer.py:583-613 (lbfgs / lbfgs2 only; DIRECT and CMA need absent packages)
    if optimizer_name == 'lbfgs':
        return OptLbfgs(bounds)
    if optimizer_name == 'lbfgs2':
        return OptLbfgs2(bounds)
    raise NotImplementedError("optimizer %r is not available on this path" % (optimizer_name,))


def apply_optimizer(optimizer, x0, f=None, df=None, f_df=None, duplicate_manager=None, context_manager=None,
                    space=None):
    """optimizer.py:425-466.  The reference runs the local optimiser twice from the same x0 (:452 and :463, the first
    result is discarded, quirk q8) and then re-evaluates f at the optimum (:464); f is deterministic so one run is kept."""
    x0 = np.atleast_2d(x0)
    suggested_x, _ = optimizer.optimize(x0, f, df, f_df)
    suggested_fx = f(suggested_x)
    return suggested_x, suggested_fx


# ---- batching of concurrent f_df evaluations --------------------------------------------------------------------
class BatchedEvaluations(object):
    """Lets several scipy L-BFGS-B runs (one thread per anchor) share device launches.

    Each run calls `f_df(x)` with a single point; the calls of all still-running anchors are collected and evaluated
    together as one (N_active, d) batch.  Every run sees exactly the values it would have seen alone (candidates are
    independent), so the per-anchor trajectories are those of the sequential reference.
    One deviation, only when the utility parameter is SAMPLED (more than 20 support points or a continuous sampler):
    the gradient call draws one fresh theta per call (uEI_noiseless.py:126, quirk q4), so a batched round shares one
    draw across its anchors where the reference would draw one per anchor; with full support (every shipped script)
    there is no draw and the runs are identical."""

    def __init__(self, f_df, n_runs):
        self.f_df = f_df
        self.cv = threading.Condition()
        self.active = n_runs
        self.pending = {}
        self.results = {}
        self.batches = 0

    def _flush_locked(self):
        keys = sorted(self.pending)
        X = np.vstack([self.pending[k] for k in keys])
        fx, dfx = self.f_df(X)
        fx = np.asarray(fx).reshape(-1)
        dfx = np.asarray(dfx).reshape(len(keys), -1)
        for i, k in enumerate(keys):
            self.results[k] = (float(fx[i]), dfx[i].copy())
        self.pending.clear()
        self.batches += 1
        self.cv.notify_all()

    def call(self, run_id, x):
        with self.cv:
            self.pending[run_id] = np.atleast_2d(x)
            if len(self.pending) == self.active:
                self._flush_locked()
            else:
                while run_id not in self.results:
                    self.cv.wait()
            return self.results.pop(run_id)

    def done(self, run_id):
        with self.cv:
            self.active -= 1
            if self.active > 0 and len(self.pending) == self.active:
                self._flush_locked()


def optimize_anchors_batched(optimizer, anchor_points, f, f_df):
    """All anchors' L-BFGS-B runs concurrently, one device launch per round of evaluations."""
    n = len(anchor_points)
    be = BatchedEvaluations(f_df, n)
    out = [None] * n
    errors = []

    def run(i):
        try:
            fdf_i = lambda x: be.call(i, x)          # noqa: E731
            x_opt, _ = optimizer.optimize(np.atleast_2d(anchor_points[i]), f=f, df=None, f_df=fdf_i)
            out[i] = x_opt
        except BaseException as e:                    # pragma: no cover
            errors.append(e)
        finally:
            be.done(i)

    threads = [threading.Thread(target=run, args=(i,)) for i in range(n)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    X_opt = np.vstack(out)
    # optimizer.py:464 re-evaluates f at ONE point per anchor: single-candidate semantics (uEI_noiseless._compute_acq
    # takes its sequential branch -- f* of the current hyper-sample -- for one candidate and the pool branch for more,
    # uEI_noiseless.py:43-46), so the optima are scored one by one, like x_baseline in AcquisitionOptimizer.optimize
    F_opt = np.concatenate([np.asarray(f(X_opt[i:i + 1])).reshape(-1) for i in range(n)])
    return [(np.atleast_2d(X_opt[i]), np.atleast_2d(F_opt[i])) for i in range(n)], be.batches


# ---- anchor points + multistart ----------------------------------------------------------------------------------
class ObjectiveAnchorPointsGenerator(object):
    """anchor_points_generator.py:19-66,87-99: random design -> one batched f(X) -> argsort -> best num_anchor."""

    def __init__(self, space, design_type, objective, num_samples=28):
        self.space, self.design_type, self.objective, self.num_samples = space, design_type, objective, num_samples

    def get(self, num_anchor=8, duplicate_manager=None, unique=False, context_manager=None, get_scores=False):
        X = initial_design(self.design_type, self.space, self.num_samples)
        scores = np.asarray(self.objective(X)).flatten()
        anchor_points = X[np.argsort(scores)[:min(len(scores), num_anchor)], :]
        if get_scores:
            return anchor_points, np.sort(scores)[0:min(len(scores), num_anchor)]
        return anchor_points


class ContextManager(object):
    def __init__(self, space, context=None):
        self.space = space
        self.noncontext_bounds = space.get_bounds()


class AcquisitionOptimizer(object):
    """acquisition_optimizer.py:21-154."""

    def __init__(self, space, optimizer='lbfgs', inner_optimizer='lbfgs2', n_starting=400, n_anchor=16, batched=True,
                 **kwargs):
        self.space = space
        self.optimizer_name = optimizer
        self.inner_optimizer_name = inner_optimizer
        self.n_starting = n_starting
        self.n_anchor = n_anchor
        self.batched = batched
        self.kwargs = kwargs
        self.context_manager = ContextManager(space)
        self.optimizer = choose_optimizer(self.optimizer_name, self.context_manager.noncontext_bounds)
        self.last_batches = 0

    def optimize(self, f=None, df=None, f_df=None, duplicate_manager=None, x_baseline=None):
        self.f, self.df, self.f_df = f, df, f_df
        self.optimizer = choose_optimizer(self.optimizer_name, self.context_manager.noncontext_bounds)
        gen = ObjectiveAnchorPointsGenerator(self.space, 'random', f, self.n_starting)
        anchor_points, anchor_points_values = gen.get(num_anchor=self.n_anchor, duplicate_manager=duplicate_manager,
                                                      context_manager=self.context_manager, get_scores=True)
        if x_baseline is not None:
            f_baseline = np.asarray(f(x_baseline))[:, 0]
            anchor_points = np.vstack((anchor_points, x_baseline))
            anchor_points_values = np.concatenate((anchor_points_values, f_baseline))
        if self.batched and f_df is not None:
            optimized_points, self.last_batches = optimize_anchors_batched(self.optimizer, anchor_points, f, f_df)
        else:
            optimized_points = [apply_optimizer(self.optimizer, a, f=f, df=None, f_df=f_df,
                                                duplicate_manager=duplicate_manager,
                                                context_manager=self.context_manager, space=self.space)
                                for a in anchor_points]
        x_min, fx_min = min(optimized_points, key=lambda t: float(np.asarray(t[1]).reshape(-1)[0]))
        if x_baseline is not None:
            for i in range(x_baseline.shape[0]):
                val = f_baseline[i]
                if val < float(np.asarray(fx_min).reshape(-1)[0]):
                    x_min = np.atleast_2d(x_baseline[i, :])
                    fx_min = val
        return x_min, fx_min


class GeneralOptimizer(AcquisitionOptimizer):
    """general_optimizer.py:19-94 (used by cbo._current_marginal_argmax: 200 random points, 24 anchors, 'lbfgs')."""

    def __init__(self, space, optimizer='lbfgs', inner_optimizer='lbfgs', **kwargs):
        super(GeneralOptimizer, self).__init__(space, optimizer, inner_optimizer, **kwargs)

    def optimize(self, f=None, df=None, f_df=None, parallel=False, duplicate_manager=None, n_starting=200, n_anchor=24):
        self.f, self.df, self.f_df = f, df, f_df
        self.optimizer = choose_optimizer(self.optimizer_name, self.context_manager.noncontext_bounds)
        gen = ObjectiveAnchorPointsGenerator(self.space, 'random', f, n_starting)
        anchor_points, anchor_points_values = gen.get(num_anchor=n_anchor, get_scores=True)
        if self.batched and f_df is not None:
            optimized_points, self.last_batches = optimize_anchors_batched(self.optimizer, anchor_points, f, f_df)
        else:
            optimized_points = [apply_optimizer(self.optimizer, a, f=f, df=None, f_df=f_df, space=self.space)
                                for a in anchor_points]
        x_min, fx_min = min(optimized_points, key=lambda t: float(np.asarray(t[1]).reshape(-1)[0]))
        if float(np.asarray(anchor_points_values[0]).reshape(-1)[0]) < float(np.asarray(fx_min).reshape(-1)[0]):
            fx_min = np.atleast_2d(anchor_points_values[0])       # general_optimizer.py:90-93
            x_min = np.atleast_2d(anchor_points[0])
        return x_min, fx_min


class Sequential(object):
    """GPyOpt/core/evaluators/sequential.py:7-23."""

    def __init__(self, acquisition, batch_size=1):
        self.acquisition = acquisition
        self.batch_size = batch_size

    def compute_batch(self, duplicate_manager=None, context_manager=None, x_baseline=None):
        x, _ = self.acquisition.optimize(duplicate_manager=duplicate_manager, x_baseline=x_baseline)
        return x
