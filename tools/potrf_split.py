"""Model-build time (min of 6) with block columns one at a time and in pairs (option chol_pair)."""
import sys, time, numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from ibo_b200 import _lib
rs = np.random.RandomState(0)
for N in (1024, 2048, 4096, 6144, 8192, 12288):
    X = rs.rand(N, 6); Y = rs.rand(N)
    out = []
    for pair in (0, 1):
        _lib.set_option("chol_pair", pair)
        ts = []
        for _ in range(6):
            t0 = time.perf_counter()
            m = _lib.Model(0, [0.5] * 6, X, Y, 0.1)
            m.close()
            ts.append(time.perf_counter() - t0)
        out.append(1e3 * min(ts))
    print("N=%d: model build %.3f ms single, %.3f ms pairs" % (N, out[0], out[1]), flush=True)
