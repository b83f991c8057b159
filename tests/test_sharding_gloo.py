"""CPU test of the N>1 host logic with world_size 2 over gloo: contiguous candidate sharding + the
lowest-global-index-wins (score, index) exchange reproduce first-occurrence argmax of the whole set."""
import os
import socket
import subprocess
import sys

import numpy as np

from ibo_b200.utils.sharding import merge_argmax, shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, %(root)r)
import torch.distributed as dist
from ibo_b200.utils.sharding import shard_range
sys.path.insert(0, os.path.join(%(root)r, 'tests'))
from dist_helpers import allreduce_argmax
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
out = []
for case in range(4):
    rs = np.random.RandomState(100 + case)
    M = [1000, 1001, 7, 4096][case]
    scores = rs.rand(M)
    if case == 1:
        scores[[10, 600, 900]] = 2.0          # ties across shards: the lowest global index must win
    if case == 2:
        scores[:] = np.nan; scores[5] = -1.0  # NaN never wins
    lo, hi = shard_range(M, world, rank)
    loc = scores[lo:hi]
    ok = loc == loc
    if ok.any():
        li = int(np.flatnonzero(ok)[np.argmax(loc[ok])]); ls = float(loc[li])
    else:
        li, ls = 0, float("nan")
    s, i = allreduce_argmax(ls, lo + li, dist)
    out.append([s, i])
print("RESULT" + json.dumps(out))
dist.destroy_process_group()
'''


def test_shard_ranges_partition():
    for M in (0, 1, 7, 1000, 1 << 20):
        for world in (1, 2, 3, 8):
            r = [shard_range(M, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == M
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_merge_rule():
    assert merge_argmax([1.0, 3.0, 3.0], [5, 9, 2]) == (3.0, 2)
    assert merge_argmax([float("nan"), -1.0], [0, 7]) == (-1.0, 7)
    s, i = merge_argmax([float("nan")], [4])
    assert s != s and i == 4


def test_two_rank_gloo_argmax():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER % {"root": ROOT}], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=240) for p in procs]
    import json
    res = []
    for (o, e), p in zip(outs, procs):
        assert p.returncode == 0, e[-2000:]
        res.append(json.loads([ln for ln in o.splitlines() if ln.startswith("RESULT")][0][6:]))
    assert res[0] == res[1] or all((a[0] != a[0] and b[0] != b[0]) or a == b for a, b in zip(res[0], res[1]))
    for case, (s, i) in enumerate(res[0]):
        rs = np.random.RandomState(100 + case)
        M = [1000, 1001, 7, 4096][case]
        scores = rs.rand(M)
        if case == 1:
            scores[[10, 600, 900]] = 2.0
            assert (s, i) == (2.0, 10)
        elif case == 2:
            assert (s, i) == (-1.0, 5)
        else:
            assert i == int(np.argmax(scores)) and s == scores.max()


DIRECT_WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, %(root)r)
import torch.distributed as dist
from ibo_b200.utils.optimize import direct
sys.path.insert(0, os.path.join(%(root)r, 'tests'))
from dist_helpers import sharded_batch_objective
from oracle import ibo_oracle as orc
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rs = np.random.RandomState(7)
X = rs.rand(40, 3); Y = np.sin(3 * X).sum(axis=1)
gp = orc.GPOracle(orc.KernelSpec(orc.K_SE_ARD, [.3, .4, .5], 3), X, Y, 0.1)
calls = []
def negei(P):
    calls.append(len(P))
    mu, s2 = gp.posterior_batch(P)
    return -orc.score(orc.ACQ_EI, "py", mu, s2, Y.max(), 0.01)
f = sharded_batch_objective(negei, dist, min_points=int(os.environ["MINPTS"]))
fmin, xmin = direct(None, [[0., 1.]] * 3, maxiter=12, batch_objective=f)
print("RESULT" + json.dumps([fmin, list(xmin), sum(calls), len(calls)]))
dist.destroy_process_group()
'''


def test_two_rank_gloo_sharded_direct_follows_the_single_rank_trajectory():
    """config #5's multi-GPU scheme on CPU: both ranks drive the same batched DIRECT (host C++ of libibo_b200), each
    evaluates half of every batch (here with the oracle as objective), values are all-gathered; result and trajectory must be
    those of the unsharded run"""
    import json
    from ibo_b200.utils.optimize import direct
    from oracle import ibo_oracle as orc
    rs = np.random.RandomState(7)
    X = rs.rand(40, 3); Y = np.sin(3 * X).sum(axis=1)
    gp = orc.GPOracle(orc.KernelSpec(orc.K_SE_ARD, [.3, .4, .5], 3), X, Y, 0.1)
    npts = []

    def negei(P):
        npts.append(len(P))
        mu, s2 = gp.posterior_batch(P)
        return -orc.score(orc.ACQ_EI, "py", mu, s2, Y.max(), 0.01)
    fmin1, xmin1 = direct(None, [[0., 1.]] * 3, maxiter=12, batch_objective=negei)
    for minpts in (0, 16):
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        procs = []
        for rank in range(2):
            env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), MINPTS=str(minpts))
            procs.append(subprocess.Popen([sys.executable, "-c", DIRECT_WORKER % {"root": ROOT}], env=env, stdout=subprocess.PIPE,
                                          stderr=subprocess.PIPE, text=True))
        outs = [p.communicate(timeout=240) for p in procs]
        res = []
        for (o, e), p in zip(outs, procs):
            assert p.returncode == 0, e[-2000:]
            res.append(json.loads([ln for ln in o.splitlines() if ln.startswith("RESULT")][0][6:]))
        for r in res:
            assert abs(r[0] - fmin1) <= 1e-12 * abs(fmin1) and np.allclose(r[1], xmin1, rtol=0, atol=0)
        assert res[0][3] == len(npts)                                  # same batches (rank 1 skips the 1-point first batch)
        if minpts == 0:
            assert res[0][2] + res[1][2] == sum(npts)                  # every point evaluated exactly once across ranks
