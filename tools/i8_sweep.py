"""Sweep of the INT8 path's launch knobs on resident candidates (device time per step, median of 5 back-to-back steps)."""
import sys
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from ibo_b200 import _lib
from ibo_b200.gaussianprocess import GaussianProcess
from ibo_b200.gaussianprocess.kernel import GaussianKernel_ard
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
d = int(sys.argv[2]) if len(sys.argv) > 2 else 6
M = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
rs = np.random.RandomState(0)
X = rs.rand(N, d); Y = np.sin(2 * X).sum(axis=1)
gp = GaussianProcess(GaussianKernel_ard([0.5] * d), X, Y, noise=0.1)
m = gp.model
c = _lib.ResidentCandidates(m, np.ascontiguousarray(rs.rand(M, d)))
def run():
    c.score(_lib.ACQ_EI, 1.0, 0.01, _lib.FLAG_MODE_CPP)
    return float(np.median([c.score(_lib.ACQ_EI, 1.0, 0.01, _lib.FLAG_MODE_CPP)[2] for _ in range(5)]))
base = run()
print("default: %.3f ms" % base, flush=True)
for ct in (148, 296, 592, 1184):
    _lib.set_option("chunk_tiles", ct)
    for per in (2, 4, 8):
        _lib.set_option("i8_rb_per_cta", per)
        print("chunk_tiles %4d  rb_per_cta %d: %.3f ms" % (ct, per, run()), flush=True)
_lib.set_option("chunk_tiles", 0); _lib.set_option("i8_rb_per_cta", 0)
for pipe in (0, 1):
    _lib.set_option("i8_pipe", pipe)
    print("i8_pipe %d: %.3f ms" % (pipe, run()), flush=True)
