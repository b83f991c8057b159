"""Host-candidate score() latency of small batches, FP64 latency shapes against the INT8 kernels, by N and batch size:
the measurement behind option i8_min_batch."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ibo_b200 import _lib
rs = np.random.RandomState(0)
d = int(sys.argv[1]) if len(sys.argv) > 1 else 6
for N in (256, 512, 1024, 2048, 4096, 8192):
    X = rs.rand(N, d); Y = np.sin(3 * X).sum(axis=1)
    m = _lib.Model(0, [0.5] * d, X, Y, 0.1)
    line = []
    for M in (32, 64, 96, 128, 192, 256, 512, 1024, 2048):
        Xs = rs.rand(M, d)
        t = []
        for mb in (1, 10 ** 9):
            _lib.set_option("i8_min_batch", mb)
            for _ in range(3): m.score(Xs, 0, 1.0, 0.01)
            t0 = time.perf_counter()
            for _ in range(20): m.score(Xs, 0, 1.0, 0.01)
            t.append(1e6 * (time.perf_counter() - t0) / 20)
        line.append("%d: %.0f/%.0f" % (M, t[0], t[1]))
    print("N=%d d=%d  M: int8/fp64 us   " % (N, d) + "  ".join(line), flush=True)
    m.close()
