"""torch.distributed (gloo) twins of the multi-GPU host logic, for the CPU tests only: the product exchanges over NCCL through
ibo_comm_argmax / ibo_comm_allgather (ibo_b200/csrc/comm.cu) and never imports torch."""
import numpy as np
import torch

from ibo_b200.utils.sharding import batch_slice, merge_argmax


def allreduce_argmax(score, index, group_dist):
    """all ranks receive the global (score, index); same merge rule as ibo_comm_argmax"""
    world = group_dist.get_world_size()
    mine = torch.tensor([score, float(index)], dtype=torch.float64)   # indices < 2^53 are exact in f64
    allp = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
    group_dist.all_gather(allp, mine)
    return merge_argmax([float(p[0]) for p in allp], [int(p[1]) for p in allp])


def sharded_batch_objective(batch_fn, group_dist, min_points=0):
    """Wrap a batch objective P (n, d) -> values (n,) so that each rank of `group_dist` evaluates one contiguous slice of
    ceil(n / world) points and the values are all-gathered -- the host-side twin of IBO_FLAG_SHARD in ibo_acqmax
    (ibo_b200/csrc/direct.cpp: gpu_batch), usable with utils.optimize.direct(batch_objective=...).  Every rank must drive the same
    deterministic DIRECT."""
    world, rank = group_dist.get_world_size(), group_dist.get_rank()

    def f(P):
        P = np.asarray(P, dtype=float)
        n = len(P)
        if world == 1 or n < min_points:
            return np.asarray(batch_fn(P), dtype=float).reshape(-1)
        per = (n + world - 1) // world
        lo, hi = batch_slice(n, world, rank)
        mine = torch.zeros(per, dtype=torch.float64)
        if hi > lo:
            mine[:hi - lo] = torch.from_numpy(np.ascontiguousarray(np.asarray(batch_fn(P[lo:hi]), dtype=float).reshape(-1)))
        parts = [torch.zeros(per, dtype=torch.float64) for _ in range(world)]
        group_dist.all_gather(parts, mine)
        return torch.cat(parts).numpy()[:n].copy()
    return f
