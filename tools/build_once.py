import os, sys, numpy as np
sys.path.insert(0, "/root/repo")
from ibo_b200 import _lib
N = 8192
rs = np.random.RandomState(0)
X = rs.rand(N, 6); Y = rs.rand(N)
m = _lib.Model(0, [0.5] * 6, X, Y, 0.1); m.close()
