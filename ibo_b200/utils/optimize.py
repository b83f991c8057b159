"""
DIRECT global optimisation with the reference's two entry points (ego/utils/optimize.py):

  direct(f, bounds, ...)   the pure-Python DIRECT of optimize.py:68-280
  cdirect(f, bounds, ...)  the ctypes wrapper around the C `direct` symbol of optimize.py:310-343

Both are served by the library's host driver (ibo_b200/csrc/direct.cpp), which follows the
deterministic C++ rules of cpp/direct.cpp.  The reference's Python DIRECT keeps its rectangles in a
`set` (optimize.py:207), so its division order is not reproducible; it shares the selection rules,
the epsilon and the 4-samples-per-long-side division with the C++ version, which is the one mirrored
(SURVEY.md 3.4).  One deliberate difference is kept: `direct` stops on `samples >= maxsample`
(optimize.py:270) by passing maxsample-1 to the driver's `>` test.
The matplotlib demos of optimize.py:347-445 are not part of the package.
"""
import ctypes
from ctypes import c_double, c_int, c_long

import numpy as np

from .. import _lib

_BIG = 2 ** 30


def _run(batch_cb, bounds, maxiter, maxtime, maxsample, flags):
    lower = np.array([b[0] for b in bounds], dtype=float)
    upper = np.array([b[1] for b in bounds], dtype=float)
    fmin, xmin = c_double(0), np.empty(len(lower))
    ns, it = c_long(0), c_int(0)
    err = []

    def cb(user, n, ndim, X, y):
        out = np.ctypeslib.as_array(y, shape=(n,))
        try:
            P = np.ctypeslib.as_array(X, shape=(n, ndim))
            out[:] = np.asarray(batch_cb(P), dtype=float).reshape(-1)
        except BaseException as e:     # never unwind through C: NaN tells the driver to stop (IBO_E_OBJECTIVE)
            err.append(e)
            out[:] = np.nan
    rc = _lib.lib().ibo_direct_batched(_lib.BATCH_OBJECTIVE(cb), None, len(lower), _lib.dptr(lower), _lib.dptr(upper),
                                       int(maxiter), int(maxtime), int(maxsample), flags, ctypes.byref(fmin), _lib.dptr(xmin),
                                       ctypes.byref(ns), ctypes.byref(it))
    if err:
        raise err[0]
    _lib.check(rc)
    return fmin.value, xmin, ns.value


def direct(f, bounds, args=None, debug=False, maxiter=None, maxsample=None, maxtime=None, batch_objective=None, pure=False):
    """Minimise f over the box `bounds` (sequence of (min, max)); returns (value, location).

    y = f(x, *args).  At least one of maxiter / maxsample / maxtime must be given (optimize.py:99-100).
    `batch_objective(P)` (optional) evaluates an (n, d) array of points at once; `pure=True` additionally tells the
    driver that the batch objective has no state, so that it may evaluate a few extra points per batch whose values
    it discards (IBO_FLAG_DIRECT_SPECULATE: one batch per iteration even where a child centre depends on the
    division order in its last bit; same result and sample count)."""
    if not (maxiter or maxsample or maxtime):
        raise ValueError("No termination criterion set!")
    args = [] if args is None else args
    if batch_objective is None:
        batch, flags = (lambda P: [f(np.array(p), *args) for p in P]), _lib.FLAG_DIRECT_SEQ
    else:
        batch, flags = batch_objective, (_lib.FLAG_DIRECT_SPECULATE if pure else 0)
    fmin, xmin, _ = _run(batch, bounds, maxiter if maxiter else _BIG, maxtime if maxtime else _BIG,
                         (maxsample - 1) if maxsample else _BIG, flags)
    return fmin, xmin


def cdirect(f, bounds, args=None, maxiter=10, maxtime=10, maxsample=200000, **kwargs):
    """The reference's C `direct` through its own ABI: a CFUNCTYPE scalar callback (optimize.py:310-343)."""
    args = [] if args is None else args
    lower = np.array([b[0] for b in bounds], dtype=float)
    upper = np.array([b[1] for b in bounds], dtype=float)

    def objective(n, x):
        return float(f(np.array([x[i] for i in range(n)]), *args))
    res = _lib.lib().direct(_lib.OBJECTIVE(objective), len(lower), _lib.dptr(lower), _lib.dptr(upper),
                            int(maxiter), int(maxtime), int(maxsample))
    if not res:
        raise _lib.IBOError(_lib.E_BADARG, "direct() failed")
    out = res[0], np.array([res[i + 1] for i in range(len(lower))])
    libc = ctypes.CDLL(None)
    libc.free.argtypes = [ctypes.c_void_p]
    libc.free(res)       # the reference leaks this buffer (optimize.py:333-343)
    return out
