"""config #1 (N = 50, d = 2, 50 DIRECT iterations): per-batch cost with the resident batch server and with one launch per batch."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ibo_oracle as orc
from ibo_b200 import _lib
b = [[-5., 10.], [0., 15.]]
X = np.array(orc.lhc_sample(b, 50, seed=0)); Y = -orc.branin(X) / 100.0
m = _lib.Model(_lib.KERNEL_SE_ARD, [3.4, 10.0], X, Y, 0.1)
lb = np.array([-5., 0.]); ub = np.array([10., 15.])
for srv in (1, 0, 1):
    _lib.set_option("tiny_server", srv)
    for it in range(6):
        if it == 5: _lib.set_option("direct_timing", 1)
        t0 = time.perf_counter()
        r = m.acqmax(lb, ub, _lib.ACQ_EI, float(Y.max()), 0.01, maxiter=50, maxsample=10000)
        dt = 1e3 * (time.perf_counter() - t0)
    _lib.set_option("direct_timing", 0)
    print("tiny_server=%d: %.3f ms, %d samples, opt %.12g" % (srv, dt, r[2], r[0]), flush=True)
