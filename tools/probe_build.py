"""model-build timing breakdown and K1 variant comparison (dev tool; run under gpurun)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ibo_b200 import _lib
rs = np.random.RandomState(0)
for N in (50, 512, 2048, 4096, 8192):
    X = rs.rand(N, 6); Y = np.sin(3 * X).sum(axis=1)
    ts = []
    for rep in range(4):
        t0 = time.perf_counter()
        m = _lib.Model(0, [.53, .57, 2.5, .34, .27, .35], X, Y, 0.1)
        ts.append(1e3 * (time.perf_counter() - t0))
        m.close()
    print("build N=%5d  ms: %s" % (N, " ".join("%.2f" % t for t in ts)), flush=True)
