import sys, time, os
import numpy as np
sys.path.insert(0, '.')
from ibo_b200 import _lib
rs = np.random.RandomState(0)
for N, d in ((50, 2), (128, 6)):
    X = rs.rand(N, d); Y = np.sin(2 * X).sum(axis=1)
    m = _lib.Model(_lib.KERNEL_SE_ARD, [0.5] * d, X, Y, 0.1)
    Xs = rs.rand(1 << 20, d)
    c = _lib.ResidentCandidates(m, Xs)
    for tiny in ("1", "0"):
        os.environ["IBO_TINY"] = tiny
        c.score(_lib.ACQ_EI, Y.max(), 0.01)
        best, idx, ms = c.score(_lib.ACQ_EI, Y.max(), 0.01)
        print("N=%d d=%d 2^20 candidates resident: IBO_TINY=%s  %.3f ms  (%.1f M evals/s)" % (N, d, tiny, ms, (1 << 20) / ms / 1e3))
    c.close()
