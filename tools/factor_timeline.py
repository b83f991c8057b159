"""Debug build only (make -C ibo_b200/csrc EXTRA=-DIBO_I8_TRACE): event timeline of the three streams of the factorisation.
Prints, per block step, when each kernel group finished (us since the start of the build)."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ibo_b200 import _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rs = np.random.RandomState(0)
X = rs.rand(N, 6); Y = rs.rand(N)
m = _lib.Model(0, [0.5] * 6, X, Y, 0.1); m.close()       # warm-up
_lib.set_option("debug_plan", 7)
m = _lib.Model(0, [0.5] * 6, X, Y, 0.1); m.close()
