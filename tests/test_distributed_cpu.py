"""N > 1 host logic on CPU: candidate sharding and the top-k all-gather/merge, world_size 2 over gloo."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cpu_local_topk(acq, X, k, offset):
    order = np.lexsort((np.arange(len(acq)), -acq))[:k]
    rec = np.full((k, 2 + X.shape[1]), -np.inf)
    rec[:, 1] = -1
    rec[:len(order), 0] = acq[order]
    rec[:len(order), 1] = order + offset
    rec[:len(order), 2:] = X[order]
    return torch.from_numpy(rec)


def _worker(rank, world, port, N, d, k, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from bocf_b200 import distributed as bd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    acq = rng.standard_normal(N)
    acq[[3, N - 2]] = acq.max() + 1.0          # a tie that spans both shards -> smaller global index wins
    X = rng.uniform(size=(N, d))
    lo, hi = bd.shard_bounds(N, world, rank)
    rec = _cpu_local_topk(acq[lo:hi], X[lo:hi], k, lo)
    out = bd.allgather_topk(rec, k)
    q.put((rank, out.numpy()))
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    from bocf_b200.distributed import shard_bounds
    for N in (0, 1, 7, 1000003):
        for W in (1, 2, 3, 8):
            b = [shard_bounds(N, W, r) for r in range(W)]
            assert b[0][0] == 0 and b[-1][1] == N
            assert all(b[i][1] == b[i + 1][0] for i in range(W - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_merge_topk_ties_and_empties():
    from bocf_b200.distributed import merge_topk
    rec = torch.tensor([[1.0, 7, 0.1], [2.0, 9, 0.2], [2.0, 4, 0.3], [float("-inf"), -1, 0.0]], dtype=torch.float64)
    out = merge_topk(rec, 3).numpy()
    assert out[:, 1].tolist() == [4.0, 9.0, 7.0]


def test_allgather_topk_world2_gloo():
    N, d, k, world = 1001, 3, 16, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, d, k, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(5)
    acq = rng.standard_normal(N)
    acq[[3, N - 2]] = acq.max() + 1.0
    X = rng.uniform(size=(N, d))
    order = np.lexsort((np.arange(N), -acq))[:k]
    for r in range(world):
        assert np.array_equal(res[r][:, 1].astype(int), order)
        assert np.array_equal(res[r][:, 0], acq[order]) and np.array_equal(res[r][:, 2:], X[order])
    assert np.array_equal(res[0], res[1])
