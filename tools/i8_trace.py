"""Debug: per-CTA timeline of trigemm_i8_kernel (library built with `make -C ibo_b200/csrc EXTRA=-DIBO_I8_TRACE`)."""
import ctypes, sys
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from ibo_b200 import _lib
from ibo_b200.gaussianprocess import GaussianProcess
from ibo_b200.gaussianprocess.kernel import GaussianKernel_ard
N, d, M = 2048, 6, 37888
rs = np.random.RandomState(0)
X = rs.rand(N, d); Y = np.sin(2 * X).sum(axis=1)
gp = GaussianProcess(GaussianKernel_ard([0.5] * d), X, Y, noise=0.1)
m = gp.model
Xs = np.ascontiguousarray(rs.rand(M, d))
c = _lib.ResidentCandidates(m, Xs)
for _ in range(3):
    c.score(_lib.ACQ_EI, float(Y.max()), 0.01, _lib.FLAG_MODE_CPP)
L = _lib.lib()
t = np.zeros(4096, dtype=np.int64)
L.ibo_debug_i8_trace(t.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), 4096)
t0 = t[0]
print("CTA total", t[1] - t0)
for rb in range(4):
    print("row-block", rb, "epilogue start", t[2 + 2 * rb] - t0, "end", t[3 + 2 * rb] - t0, "len", t[3 + 2 * rb] - t[2 + 2 * rb])
n = 0
rows = []
prev = None
while 16 + 4 * n + 3 < 1024 and t[16 + 4 * n] != 0:
    a, b, c_, dd = t[16 + 4 * n: 16 + 4 * n + 4] - t0
    rows.append((n, a, b - a, c_ - b, dd - c_, (a - prev) if prev is not None else 0))
    prev = a
    n += 1
print("k-steps", n)
print("  n   start  wait_full  issue10  commit  since_prev_start")
for r in rows[:10] + rows[60:68] + rows[-4:]:
    print("%4d %8d %8d %8d %8d %8d" % r)
arr = np.array(rows)
print("sum wait_full", arr[:, 2].sum(), "issue", arr[:, 3].sum(), "commit", arr[:, 4].sum(), "median step", np.median(arr[1:, 5]))
print("loader set 0 (k-steps 0, 2, 4, ...):  use  start  wait_empty  store+wait::st   since_prev")
prev = None
for u in list(range(0, 8)) + list(range(30, 36)):
    a, b, c_ = t[1024 + 4 * u: 1024 + 4 * u + 3] - t0
    print("%4d %8d %8d %8d %8d" % (u, a, b - a, c_ - b, (a - prev) if prev is not None else 0))
    prev = a
