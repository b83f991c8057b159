"""Candidate sharding across GPUs: candidates are independent given the (replicated) model, so
rank r scores the contiguous index range shard_range(M, world, r) with no data-path collective; the only
exchange is the (score, global index) argmax, merged with the lowest-global-index-wins rule so that the result
equals first-occurrence argmax over the whole set.  On GPUs the exchange is ibo_comm_argmax (NCCL all-gather of
16-byte pairs, ibo_b200/csrc/comm.cu); `allreduce_argmax` can also run over any torch.distributed group (gloo in
the CPU tests) so the host logic is testable without NCCL."""
import numpy as np


def shard_range(M, world, rank):
    """[lo, hi) of rank's contiguous candidate block; blocks differ in size by at most one."""
    base, extra = divmod(int(M), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def merge_argmax(scores, indices):
    """max score, lowest index on ties, NaN never wins (the rule of K3/K4 and of ibo_comm_argmax)."""
    best_s, best_i = None, None
    for s, i in zip(scores, indices):
        if s != s:
            continue
        if best_s is None or s > best_s or (s == best_s and i < best_i):
            best_s, best_i = s, i
    if best_s is None:
        return float("nan"), int(min(indices))
    return float(best_s), int(best_i)


def allreduce_argmax(score, index, group_dist=None):
    """all ranks receive the global (score, index)"""
    if group_dist is None:
        import ctypes
        from .. import _lib
        s, i = ctypes.c_double(score), ctypes.c_long(index)
        _lib.check(_lib.lib().ibo_comm_argmax(ctypes.byref(s), ctypes.byref(i)))
        return s.value, i.value
    import torch
    world = group_dist.get_world_size()
    mine = torch.tensor([score, float(index)], dtype=torch.float64)   # indices < 2^53 are exact in f64
    allp = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
    group_dist.all_gather(allp, mine)
    return merge_argmax([float(p[0]) for p in allp], [int(p[1]) for p in allp])


def sharded_batch_objective(batch_fn, group_dist, min_points=0):
    """Wrap a batch objective P (n, d) -> values (n,) so that each rank of `group_dist` (a torch.distributed-like module:
    get_rank / get_world_size / all_gather) evaluates one contiguous slice of ceil(n / world) points and the values are
    all-gathered -- the host-side twin of IBO_FLAG_SHARD in ibo_acqmax (ibo_b200/csrc/direct.cpp: gpu_batch), usable with
    utils.optimize.direct(batch_objective=...).  Every rank must drive the same deterministic DIRECT."""
    import torch
    world, rank = group_dist.get_world_size(), group_dist.get_rank()

    def f(P):
        P = np.asarray(P, dtype=float)
        n = len(P)
        if world == 1 or n < min_points:
            return np.asarray(batch_fn(P), dtype=float).reshape(-1)
        per = (n + world - 1) // world
        lo = min(rank * per, n)
        hi = min(lo + per, n)
        mine = torch.zeros(per, dtype=torch.float64)
        if hi > lo:
            mine[:hi - lo] = torch.from_numpy(np.ascontiguousarray(np.asarray(batch_fn(P[lo:hi]), dtype=float).reshape(-1)))
        parts = [torch.zeros(per, dtype=torch.float64) for _ in range(world)]
        group_dist.all_gather(parts, mine)
        return torch.cat(parts).numpy()[:n].copy()
    return f
