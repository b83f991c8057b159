"""Host / GPU split of one config #5 query (option direct_timing): batched DIRECT, N = 4096, d = 20, 200 iterations."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ibo_b200 import _lib
from ibo_b200.gaussianprocess import GaussianProcess
from ibo_b200.gaussianprocess.kernel import GaussianKernel_ard
from ibo_b200.acquisition import maximizeEI
rs = np.random.RandomState(5)
X5 = rs.rand(4096, 20); Y5 = np.sin(2 * X5).sum(axis=1)
gp5 = GaussianProcess(GaussianKernel_ard([1.0] * 20), X5, Y5, noise=0.1)
gp5.model
for it in range(3):
    if it == 2: _lib.set_option("direct_timing", 1)
    t0 = time.perf_counter()
    maximizeEI(gp5, [[0., 1.]] * 20, xi=0.01, maxiter=200, maxtime=10 ** 6, maxsample=10 ** 9)
    print("query %d: %.2f ms" % (it, 1e3 * (time.perf_counter() - t0)), flush=True)
for mb in (64, 128, 192, 256, 384, 100000):
    _lib.set_option("direct_timing", 0)
    _lib.set_option("i8_min_batch", mb)
    ts = []
    for it in range(3):
        t0 = time.perf_counter()
        maximizeEI(gp5, [[0., 1.]] * 20, xi=0.01, maxiter=200, maxtime=10 ** 6, maxsample=10 ** 9)
        ts.append(1e3 * (time.perf_counter() - t0))
    print("i8_min_batch %d: %.2f ms" % (mb, min(ts)), flush=True)
