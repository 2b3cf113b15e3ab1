"""Candidate sharding and the one exchange step of the path: best-candidate selection across ranks.

Candidates (and multistart anchors) are independent, so each rank scores its own contiguous block
with a replicated model.  The only collective is an all-gather of each rank's local
top-k records (k x (2 + d) fp64 = value, global index, x) over NCCL/NVLink (gloo on CPU in tests),
after which every rank selects the same global top-k (ties break on the smaller global index) --
the anchor selection of GPyOpt/optimization/anchor_points_generator.py:59-64 with f = -acq.
"""
import ctypes

import numpy as np
import torch

from . import _lib


def shard_bounds(N, world_size, rank):
    """Contiguous block [lo, hi) of rank `rank` when N candidates are split over world_size ranks."""
    base, rem = divmod(int(N), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def local_topk(acq, X, k, index_offset=0):
    """Top-k records of this rank's candidates, on the device (bocf_topk).  acq (N,) / (N,1), X (N,d) CUDA."""
    lib = _lib.load_library()
    acq = acq.reshape(-1).contiguous()
    X = X.contiguous()
    N, d = X.shape
    rec = torch.empty((k, 2 + d), dtype=torch.float64, device=X.device)
    with torch.cuda.device(X.device):
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(lib.bocf_topk(ctypes.c_void_p(acq.data_ptr()), ctypes.c_void_p(X.data_ptr()), N, d, k,
                                 int(index_offset), ctypes.c_void_p(rec.data_ptr()), st))
    return rec


def merge_topk(records, k):
    """Global top-k of stacked records (R, 2+d): largest value first, ties -> smaller global index.
    Empty slots carry index -1.  Works on CPU or CUDA tensors; deterministic on every rank."""
    rec = records[records[:, 1] >= 0]
    if rec.shape[0] == 0:
        return rec
    val = rec[:, 0].cpu().numpy()
    idx = rec[:, 1].cpu().numpy()
    order = np.lexsort((idx, -val))[:k]
    return rec[torch.as_tensor(order, device=rec.device)]


def allgather_topk(rec_local, k, group=None):
    """All-gather the per-rank records and select the global top-k (identical result on every rank)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return merge_topk(rec_local, k)
    world = dist.get_world_size(group)
    out = torch.empty((world * rec_local.shape[0], rec_local.shape[1]), dtype=rec_local.dtype,
                      device=rec_local.device)
    dist.all_gather_into_tensor(out, rec_local.contiguous(), group=group)
    return merge_topk(out, k)


def sharded_best_candidates(acquisition, X_local, k=16, index_offset=0, with_gradients=False, group=None):
    """Score this rank's candidates with `acquisition` and return the global top-k records."""
    if with_gradients:
        acq, _ = acquisition._compute_acq_withGradients(X_local)
    else:
        acq = acquisition._compute_acq(X_local)
    rec = local_topk(acq, X_local, k, index_offset)
    return allgather_topk(rec, k, group=group)
