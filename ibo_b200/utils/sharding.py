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
