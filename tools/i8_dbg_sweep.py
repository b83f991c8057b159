"""Debug build only (make -C ibo_b200/csrc EXTRA=-DIBO_I8_TRACE): K2 of the INT8 path with one component removed at a time (option
i8_dbg; results are wrong, only the time matters) -- how profiles/r02_int8_k2.md separated issue, feed and tensor time."""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
from ibo_b200 import _lib
from ibo_b200.gaussianprocess import GaussianProcess
from ibo_b200.gaussianprocess.kernel import GaussianKernel_ard
N, d, M = 2048, 6, 1 << 18
rs = np.random.RandomState(0)
X = rs.rand(N, d); Y = np.sin(2 * X).sum(axis=1)
gp = GaussianProcess(GaussianKernel_ard([0.5] * d), X, Y, noise=0.1)
m = gp.model
c = _lib.ResidentCandidates(m, np.ascontiguousarray(rs.rand(M, d)))
for dbg in (0, 1, 2, 2 | 8, 4, 8, 16, 32, 16 | 32, 2 | 8 | 16 | 32, 1 | 16 | 32, 4 | 32, 2 | 4 | 8):
    _lib.set_option("i8_dbg", dbg)
    c.score(_lib.ACQ_EI, 1.0, 0.01, _lib.FLAG_MODE_CPP | _lib.FLAG_PROFILE)
    c.score(_lib.ACQ_EI, 1.0, 0.01, _lib.FLAG_MODE_CPP | _lib.FLAG_PROFILE)
    p = m.profile()
    ksteps = (M / 64) * 544 / 148
    print("dbg %2d: K2 %.3f ms  -> %.0f clk per k-step at 1.965 GHz" % (dbg, p["k2_ms"], p["k2_ms"] * 1e-3 * 1.965e9 / ksteps), flush=True)
