"""CPU tests of the host-side batched DIRECT driver (ibo_b200/csrc/direct.cpp) through the C ABI:
the legacy scalar-callback `direct` symbol and `ibo_direct_batched` must follow the reference's
trajectory exactly -- same FMIN, XMIN, sample count and sample sequence as libego's `direct`
(golden traces in tests/golden/ref_direct.json, generated from the reference itself)."""
import ctypes
import json
import os
from ctypes import c_double, c_int, c_long

import numpy as np
import pytest

from ibo_b200 import _lib
from oracle import ibo_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "ref_direct.json")) as fh:
    GOLD = json.load(fh)["direct"]

FNS = {"shekel5": orc.shekel5, "branin": lambda x: float(orc.branin(x)),
       "quad": lambda x: float(np.sum((x - 0.3) ** 2)), "sin6": lambda x: float(np.sum(np.sin(3 * x) + (x - .4) ** 2))}


@pytest.mark.parametrize("tag", sorted(GOLD.keys()))
def test_legacy_direct_symbol_matches_reference(tag):
    rec = GOLD[tag]
    f = FNS[tag.split("_")[0]]
    b = np.array(rec["bounds"])
    lb = np.ascontiguousarray(b[:, 0]); ub = np.ascontiguousarray(b[:, 1])
    trace = []

    def cb(n, x):
        xx = np.array([x[i] for i in range(n)])
        trace.append(xx)
        return float(f(xx))
    res = _lib.lib().direct(_lib.OBJECTIVE(cb), len(lb), _lib.dptr(lb), _lib.dptr(ub), rec["maxiter"], 100000, rec["maxsample"])
    assert res
    assert res[0] == rec["fmin"]
    assert [res[i + 1] for i in range(len(lb))] == rec["xmin"]
    tr = np.array(trace)
    assert len(tr) == rec["nsamples"]
    # sequential mode reproduces the reference's call order sample by sample
    assert np.array_equal(tr[:40], np.array(rec["trace_head"]))
    assert np.array_equal(tr[-10:], np.array(rec["trace_tail"]))
    libc = ctypes.CDLL(None)
    libc.free.argtypes = [ctypes.c_void_p]
    libc.free(res)            # caller frees, as ego/acquisition/__init__.py:443-447 does


@pytest.mark.parametrize("tag", sorted(GOLD.keys()))
@pytest.mark.parametrize("seq", [False, True])
def test_batched_direct_matches_reference(tag, seq):
    rec = GOLD[tag]
    f = FNS[tag.split("_")[0]]
    b = np.array(rec["bounds"])
    lb = np.ascontiguousarray(b[:, 0]); ub = np.ascontiguousarray(b[:, 1])
    batches = []
    pts = []

    def cb(user, n, ndim, X, y):
        A = np.ctypeslib.as_array(X, shape=(n, ndim)).copy()
        batches.append(n)
        pts.append(A)
        for i in range(n):
            y[i] = float(f(A[i]))
    fmin = c_double(0); xmin = np.empty(len(lb)); ns = c_long(0); it = c_int(0)
    flags = _lib.FLAG_DIRECT_SEQ if seq else 0
    rc = _lib.lib().ibo_direct_batched(_lib.BATCH_OBJECTIVE(cb), None, len(lb), _lib.dptr(lb), _lib.dptr(ub), rec["maxiter"], 100000,
                                       rec["maxsample"], flags, ctypes.byref(fmin), _lib.dptr(xmin), ctypes.byref(ns), ctypes.byref(it))
    assert rc == 0
    assert fmin.value == rec["fmin"]
    assert list(xmin) == rec["xmin"]
    assert ns.value == rec["nsamples"] == sum(batches)
    allpts = np.vstack(pts)
    assert np.allclose(allpts.sum(axis=0), rec["trace_sum"], rtol=1e-12)     # same multiset of samples
    if not seq:
        # two batches per iteration (+ the initial centre and the first division)
        assert len(batches) <= 2 * (it.value + 1) + 1


def test_maxsample_zero_and_degenerate_boxes():
    L = _lib.lib()
    calls = []

    def cb(user, n, ndim, X, y):
        calls.append(n)
        for i in range(n):
            y[i] = float(i)
    lb = np.array([0.0, 1.0]); ub = np.array([1.0, 1.0])      # second dim fixed
    fmin = c_double(0); xmin = np.empty(2); ns = c_long(0); it = c_int(0)
    rc = L.ibo_direct_batched(_lib.BATCH_OBJECTIVE(cb), None, 2, _lib.dptr(lb), _lib.dptr(ub), 5, 1000, 0, 0,
                              ctypes.byref(fmin), _lib.dptr(xmin), ctypes.byref(ns), ctypes.byref(it))
    assert rc == 0 and xmin[1] == 1.0 and ns.value >= 5
    assert L.ibo_direct_batched(_lib.BATCH_OBJECTIVE(cb), None, 0, None, None, 1, 1, 1, 0, None, None, None, None) == _lib.E_BADARG


def test_nested_direct_on_one_thread_is_safe():
    """an objective that itself runs DIRECT (same host thread): the inner run must not disturb the outer one's rectangle store"""
    from ibo_b200.utils.optimize import direct

    def inner(x):
        return float(np.sum((np.asarray(x) - 0.3) ** 2))

    def outer_plain(x):
        return float(np.sum((np.asarray(x) - 0.6) ** 2))

    def outer_nested(x):
        v, _ = direct(inner, [[0., 1.]] * 2, maxiter=5)
        assert abs(v - 2 * (0.5 - 0.3) ** 2) < 0.1
        return outer_plain(x)

    a = direct(outer_plain, [[0., 1.]] * 2, maxiter=15)
    b = direct(outer_nested, [[0., 1.]] * 2, maxiter=15)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])
