"""wall time per ibo_score_batch call for DIRECT-sized batches (run under gpurun)"""
import sys, time, json
import numpy as np
sys.path.insert(0, '.')
from ibo_b200 import _lib
rs = np.random.RandomState(0)
for N, d in ((50, 2), (2048, 6), (4096, 20)):
    X = rs.rand(N, d); Y = np.sin(2 * X).sum(axis=1)
    m = _lib.Model(_lib.KERNEL_SE_ARD, [0.5] * d, X, Y, 0.1)
    for M in (4, 24, 64, 162, 512, 2048):
        Xs = rs.rand(M, d)
        for _ in range(20):
            m.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, _lib.FLAG_MODE_CPP)
        t0 = time.perf_counter()
        R = 300
        for _ in range(R):
            m.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, _lib.FLAG_MODE_CPP)
        t = (time.perf_counter() - t0) / R
        m.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, _lib.FLAG_MODE_CPP | _lib.FLAG_PROFILE)
        p = m.profile()
        print(json.dumps({"N": N, "d": d, "M": M, "wall_us": round(1e6 * t, 1), "k1_us": round(1e3 * p["k1_ms"], 1),
                          "k2_us": round(1e3 * p["k2_ms"], 1), "k3_us": round(1e3 * p["k3_ms"], 1)}), flush=True)
