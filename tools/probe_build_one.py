import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ibo_b200 import _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
rs = np.random.RandomState(0)
X = rs.rand(N, 6); Y = np.sin(3 * X).sum(axis=1)
for rep in range(2):
    m = _lib.Model(0, [.53, .57, 2.5, .34, .27, .35], X, Y, 0.1)
    m.close()
