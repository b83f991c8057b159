"""Where a small score() call goes: kernel times (IBO_FLAG_PROFILE: events around K1 / K2 / K3) against the wall time of the call."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ibo_b200 import _lib
rs = np.random.RandomState(0)
for N, d in ((50, 2), (2048, 6), (4096, 20)):
    X = rs.rand(N, d); Y = np.sin(3 * X).sum(axis=1)
    m = _lib.Model(0, [0.5] * d, X, Y, 0.1)
    for M in (18, 64, 330):
        Xs = rs.rand(M, d)
        for _ in range(5): m.score(Xs, 0, 1.0, 0.01, flags=_lib.FLAG_MODE_CPP)
        t0 = time.perf_counter()
        for _ in range(200): m.score(Xs, 0, 1.0, 0.01, flags=_lib.FLAG_MODE_CPP)
        wall = 1e6 * (time.perf_counter() - t0) / 200
        m.score(Xs, 0, 1.0, 0.01, flags=_lib.FLAG_MODE_CPP | _lib.FLAG_PROFILE)
        p = m.profile()
        print("N=%d d=%d M=%d: call %.1f us; profiled K1 %.1f K2 %.1f K3 %.1f us (%d launches)" % (
            N, d, M, wall, 1e3 * p["k1_ms"], 1e3 * p["k2_ms"], 1e3 * p["k3_ms"], p["launches"]), flush=True)
    m.close()
