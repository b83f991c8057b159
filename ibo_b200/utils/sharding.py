"""Candidate sharding across GPUs: candidates are independent given the (replicated) model, so
rank r scores the contiguous index range shard_range(M, world, r) with no data-path collective; the only
exchange is the (score, global index) argmax, merged with the lowest-global-index-wins rule so that the result
equals first-occurrence argmax over the whole set.  On GPUs the exchange is ibo_comm_argmax (NCCL all-gather of
16-byte pairs, ibo_b200/csrc/comm.cu).  The gloo twins used by the CPU tests (the same merge rule and the same batch
slicing over torch.distributed) live in tests/dist_helpers.py: nothing in this package imports torch."""
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


def allreduce_argmax(score, index):
    """all ranks receive the global (score, index): ibo_comm_argmax over the communicator set up with ibo_comm_init"""
    import ctypes
    from .. import _lib
    s, i = ctypes.c_double(score), ctypes.c_long(index)
    _lib.check(_lib.lib().ibo_comm_argmax(ctypes.byref(s), ctypes.byref(i)))
    return s.value, i.value


def batch_slice(n, world, rank):
    """[lo, hi) of the slice of an n-point DIRECT batch that `rank` evaluates under IBO_FLAG_SHARD (ibo_b200/csrc/direct.cpp:
    gpu_batch): ceil(n / world) points per rank, the last ranks may get fewer or none"""
    per = (int(n) + int(world) - 1) // int(world)
    lo = min(rank * per, n)
    return lo, min(lo + per, n)
