"""Host logic of the acquisition optimiser (GPyOpt/optimization/acquisition_optimizer.py:114-148,
optimizer.py:319-354,425-466, anchor_points_generator.py:19-66) on CPU, with a closed-form objective standing in for the
device acquisition: the batched multistart (one thread per anchor, evaluations gathered into one batch per round) must
follow exactly the per-anchor L-BFGS-B trajectories of the sequential reference loop."""
import threading

import numpy as np
import pytest

from bocf_b200 import optimization as O


def _objective(d, seed=0):
    rng = np.random.default_rng(seed)
    c = rng.uniform(0.2, 0.8, size=d)
    A = rng.uniform(0.5, 2.0, size=d)
    calls = {"f": 0, "f_df": 0, "rows": 0, "max_rows": 0}
    lock = threading.Lock()

    def f(X):
        X = np.atleast_2d(X)
        with lock:
            calls["f"] += 1
        return (A * (X - c) ** 2).sum(1, keepdims=True) + 0.1 * np.sin(5 * X).sum(1, keepdims=True)

    def f_df(X):
        X = np.atleast_2d(X)
        with lock:
            calls["f_df"] += 1
            calls["rows"] += X.shape[0]
            calls["max_rows"] = max(calls["max_rows"], X.shape[0])
        return f(X), 2 * A * (X - c) + 0.5 * np.cos(5 * X)

    return f, f_df, calls


def test_design_space_and_initial_design():
    space = O.Design_space(space=[{'name': 'var', 'type': 'continuous', 'domain': (-2, 2), 'dimensionality': 3}])
    assert space.get_bounds() == [(-2, 2)] * 3
    np.random.seed(0)
    X = O.initial_design('random', space, 8)
    assert X.shape == (8, 3) and X.min() >= -2 and X.max() <= 2
    with pytest.raises(NotImplementedError):
        O.choose_optimizer('CMA', space.get_bounds())


def test_batched_multistart_equals_sequential_on_cpu():
    d, n = 4, 9
    f, f_df, calls = _objective(d)
    bounds = [(0, 1)] * d
    opt = O.choose_optimizer('lbfgs2', bounds)
    anchors = np.random.default_rng(3).uniform(size=(n, d))
    seq = [O.apply_optimizer(opt, a, f=f, df=None, f_df=f_df) for a in anchors]
    rows_seq = calls["rows"]
    calls.update(f=0, f_df=0, rows=0, max_rows=0)
    bat, batches = O.optimize_anchors_batched(opt, anchors, f, f_df)
    assert len(bat) == n
    for (xs, fs), (xb, fb) in zip(seq, bat):
        assert np.array_equal(xs, xb) and np.array_equal(np.asarray(fs).reshape(-1), np.asarray(fb).reshape(-1))
    # the same evaluations, gathered: as many rows as the sequential runs asked for, in far fewer calls
    assert calls["rows"] == rows_seq and calls["f_df"] == batches and batches < rows_seq
    assert calls["max_rows"] == n                                   # the first round carries every anchor


def test_batched_evaluations_survive_runs_of_different_length():
    # runs that stop early must not dead-lock the others (BatchedEvaluations.done re-checks the pending set)
    be = O.BatchedEvaluations(lambda X: (X.sum(1), np.ones_like(X)), 3)
    out = {}

    def run(i, rounds):
        for r in range(rounds):
            out[(i, r)] = be.call(i, np.full((1, 2), float(i + r)))
        be.done(i)

    ts = [threading.Thread(target=run, args=(i, k)) for i, k in enumerate((1, 3, 5))]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=20)
    assert not any(t.is_alive() for t in ts)
    assert out[(2, 4)][0] == 12.0 and out[(0, 0)][0] == 0.0 and len(out) == 9


def test_acquisition_optimizer_picks_the_best_anchor_and_honours_the_baseline():
    d = 3
    f, f_df, _ = _objective(d, seed=5)
    space = O.Design_space(space=[{'name': 'var', 'type': 'continuous', 'domain': (0, 1), 'dimensionality': d}])
    np.random.seed(1)
    ao = O.AcquisitionOptimizer(space, optimizer='lbfgs2', inner_optimizer='lbfgs2', n_starting=64, n_anchor=5)
    x_min, fx_min = ao.optimize(f=f, df=None, f_df=f_df)
    assert x_min.shape == (1, d) and np.all(x_min >= 0) and np.all(x_min <= 1)
    grid = np.random.default_rng(0).uniform(size=(4000, d))
    assert float(np.asarray(fx_min).reshape(-1)[0]) <= f(grid).min() + 1e-6        # at least as good as a dense random search
    # a baseline that is better than every optimum wins (acquisition_optimizer.py:140-148)
    x_star = x_min.copy()
    f_shift = lambda X: f(X) + 5.0 * (np.abs(np.atleast_2d(X) - x_star).sum(1, keepdims=True) > 1e-12)   # noqa: E731
    np.random.seed(1)
    x2, fx2 = ao.optimize(f=f_shift, df=None, f_df=lambda X: (f_shift(X), f_df(X)[1]), x_baseline=x_star)
    assert np.array_equal(x2, np.atleast_2d(x_star))
