"""Debug build only (make -C ibo_b200/csrc EXTRA=-DIBO_I8_TRACE): time and clock64 timeline of potrf_diag_kernel alone."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ibo_b200 import _lib
L = _lib.lib()
st = (ctypes.c_longlong * 72)()
us = ctypes.c_double()
rc = L.ibo_debug_potrf(0, st, ctypes.byref(us))
assert rc == 0, rc
s = list(st)
print("potrf_diag alone: %.1f us" % us.value)
print("load %d clk" % (s[1] - s[0]))
for p in range(16):
    b = s[4 + 3 * (p - 1)] if p else s[1]
    print("panel %2d: diag %5d  trsm+sync %5d  trail %5d" % (p, s[2 + 3 * p] - b, s[3 + 3 * p] - s[2 + 3 * p], s[4 + 3 * p] - s[3 + 3 * p]))
print("cholesky total %d clk; store A %d" % (s[49] - s[1], s[50] - s[49]))
print("X_jj %d" % (s[51] - s[50]))
print("merge levels b = 8, 16, 32, 64:", [s[52 + i] - s[51 + i] for i in range(4)])
print("store D %d; total %d clk" % (s[67] - s[55], s[67] - s[0]))
