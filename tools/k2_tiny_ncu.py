import sys, numpy as np
sys.path.insert(0, '.')
from ibo_b200 import _lib
rs = np.random.RandomState(0)
N, d = 4096, 20
X = rs.rand(N, d); Y = np.sin(2 * X).sum(axis=1)
m = _lib.Model(_lib.KERNEL_SE_ARD, [0.5] * d, X, Y, 0.1)
Xs = rs.rand(162, d)
for _ in range(6):
    m.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, _lib.FLAG_MODE_CPP)
