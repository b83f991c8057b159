"""Two-GPU tests (skipped on a one-GPU box): NCCL argmax / all-gather of comm.cu and the sharded batched DIRECT
(IBO_FLAG_SHARD, config #5's scheme) -- the sharded query must follow the single-GPU trajectory bit for bit."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json, ctypes
import numpy as np
sys.path.insert(0, %(root)r)
import torch.distributed as dist
from ibo_b200 import _lib
from ibo_b200.acquisition import maximizeEI, cdirectGP
from ibo_b200.gaussianprocess import GaussianProcess
from ibo_b200.gaussianprocess.kernel import GaussianKernel_ard
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
L = _lib.lib()
uid = [None]
if rank == 0:
    buf = ctypes.create_string_buffer(128)
    _lib.check(L.ibo_comm_unique_id(buf)); uid[0] = buf.raw
dist.broadcast_object_list(uid, src=0)
_lib.check(L.ibo_comm_init(rank, rank, world, uid[0]))
assert L.ibo_comm_rank() == rank and L.ibo_comm_size() == world
# all-gather of per-rank slices
mine = np.arange(5, dtype=float) + 10 * rank
allv = np.zeros(5 * world)
_lib.check(L.ibo_comm_allgather(_lib.dptr(mine), 5, _lib.dptr(allv)))
assert np.array_equal(allv, np.concatenate([np.arange(5.) + 10 * r for r in range(world)]))
# argmax: ties -> lowest global index
s, i = ctypes.c_double(1.5), ctypes.c_long(100 - rank)
_lib.check(L.ibo_comm_argmax(ctypes.byref(s), ctypes.byref(i)))
assert (s.value, i.value) == (1.5, 100 - (world - 1))
# sharded DIRECT == unsharded DIRECT
rs = np.random.RandomState(5)
N, d = 600, 8
X = rs.rand(N, d); Y = np.sin(2 * X).sum(axis=1)
gp = GaussianProcess(GaussianKernel_ard([0.8] * d), X, Y, noise=0.1, device=rank)
b = [[0., 1.]] * d
o1 = maximizeEI(gp, b, xi=0.01, maxiter=60, maxtime=10 ** 6, maxsample=10 ** 9)
n1 = cdirectGP.last["nsamples"]
_lib.set_option("shard_min", 2)          # shard every batch of >= 2 points (default: 64 x ranks)
o2 = maximizeEI(gp, b, xi=0.01, maxiter=60, maxtime=10 ** 6, maxsample=10 ** 9, shard=True)
n2 = cdirectGP.last["nsamples"]
print("RESULT" + json.dumps([o1[0], list(o1[1]), n1, o2[0], list(o2[1]), n2]))
L.ibo_comm_destroy()
dist.destroy_process_group()
'''


def test_two_gpu_collectives_and_sharded_direct():
    from ibo_b200 import _lib
    if _lib.lib().ibo_device_count() < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER % {"root": ROOT}], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=600) for p in procs]
    res = []
    for (o, e), p in zip(outs, procs):
        assert p.returncode == 0, e[-3000:]
        res.append(json.loads([ln for ln in o.splitlines() if ln.startswith("RESULT")][0][6:]))
    assert res[0] == res[1]                                  # both ranks hold the same answer
    o1, x1, n1, o2, x2, n2 = res[0]
    assert o1 == o2 and x1 == x2 and n1 == n2                # sharded == unsharded, bit for bit
