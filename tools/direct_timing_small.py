"""Host / GPU split of the N = 2048, d = 6 maximizeEI query of the suite (option direct_timing)."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ibo_b200 import _lib
from ibo_b200.gaussianprocess import GaussianProcess
from ibo_b200.gaussianprocess.kernel import GaussianKernel_ard
from ibo_b200.acquisition import maximizeEI
rs = np.random.RandomState(1)
N, d = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2048, 6)
X = rs.rand(N, d); Y = np.sin(3 * X).sum(axis=1)
gp = GaussianProcess(GaussianKernel_ard([0.5] * d), X, Y, noise=0.1)
gp.model
for it in range(4):
    if it == 3: _lib.set_option("direct_timing", 1)
    t0 = time.perf_counter()
    maximizeEI(gp, [[0., 1.]] * d, xi=0.01, maxiter=50, maxtime=10 ** 6, maxsample=10000)
    print("query %d: %.3f ms" % (it, 1e3 * (time.perf_counter() - t0)), flush=True)
_lib.set_option("direct_timing", 0)
Xs = rs.rand(18, d)
m = gp.model
for M in (1, 8, 18, 32):
    for _ in range(5): m.score(Xs[:M], 0, 1.0, 0.01, flags=_lib.FLAG_MODE_CPP)
    t0 = time.perf_counter()
    for _ in range(200): m.score(Xs[:M], 0, 1.0, 0.01, flags=_lib.FLAG_MODE_CPP)
    print("score() of %d candidates: %.1f us" % (M, 1e6 * (time.perf_counter() - t0) / 200), flush=True)
_lib.set_option("debug_plan", 1)
m.score(Xs[:18], 0, 1.0, 0.01, flags=_lib.FLAG_MODE_CPP)
