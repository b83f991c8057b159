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


# ---------------------------------------------------------------------------------------------------------------
# randomized differential test against the reference's own `direct` (oracle/_ref/libego.so, compiled from the
# untouched cpp/direct.cpp) and against the oracle restatement: random boxes (some dims fixed), objectives built to
# hit the selection rules' corner cases -- exact y ties (quantised / constant objectives), FMIN == 0 (cpp/direct.cpp:
# 430-442), negative and huge values, `maxsample` cuts in the middle of an iteration (:487-492)
# ---------------------------------------------------------------------------------------------------------------
REF_LIBEGO = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libego.so")


def _random_case(seed):
    rs = np.random.RandomState(1000 + seed)
    d = int(rs.randint(1, 7))
    lb = np.round(rs.uniform(-3, 1, d), 2)
    ub = lb + np.round(rs.uniform(0.5, 4, d), 2)
    for i in range(1 if d > 1 else 0, d):        # dim 0 fixed stalls DIRECT by design, covered by the golden traces
        if rs.rand() < 0.15:
            ub[i] = lb[i]
    c = lb + rs.rand(d) * (ub - lb + 1e-9)
    w = rs.uniform(0.2, 3.0, d)
    kind = seed % 6
    if kind == 0:
        f = lambda x: float(np.sum(w * (x - c) ** 2))                              # smooth
    elif kind == 1:
        f = lambda x: float(np.floor(4 * np.sum(w * (x - c) ** 2)) / 4)            # plateaus: many exact ties
    elif kind == 2:
        f = lambda x: 1.25                                                         # everything ties
    elif kind == 3:
        f = lambda x: float(np.sum(np.abs(np.round(x - lb, 12))))                  # reaches FMIN == 0 only at the corner
    elif kind == 4:
        f = lambda x: float(-1e6 * np.exp(-np.sum(w * (x - c) ** 2)) + np.sum(np.sin(5 * x)))   # large negative
    else:
        f = lambda x: float(np.max(np.abs(x - c)) > 0.4)                           # 0/1 valued: FMIN == 0 with ties
    maxiter = int(rs.randint(1, 22))
    maxsample = int(rs.choice([7, 50, 333, 100000]))
    return d, lb, ub, f, maxiter, maxsample


def _run_ours(d, lb, ub, f, maxiter, maxsample, flags):
    pts = []

    def cb(user, n, ndim, X, y):
        A = np.ctypeslib.as_array(X, shape=(n, ndim)).copy()
        pts.append(A)
        for i in range(n):
            y[i] = f(A[i])
    fmin = c_double(0); xmin = np.empty(d); ns = c_long(0); it = c_int(0)
    rc = _lib.lib().ibo_direct_batched(_lib.BATCH_OBJECTIVE(cb), None, d, _lib.dptr(lb), _lib.dptr(ub), maxiter, 100000, maxsample,
                                       flags, ctypes.byref(fmin), _lib.dptr(xmin), ctypes.byref(ns), ctypes.byref(it))
    assert rc == 0
    return fmin.value, xmin.copy(), ns.value, np.vstack(pts)


def _run_reference(d, lb, ub, f, maxiter, maxsample):
    ref = ctypes.CDLL(REF_LIBEGO)
    ref.direct.restype = ctypes.POINTER(c_double)
    trace = []

    def cb(n, x):
        xx = np.array([x[i] for i in range(n)])
        trace.append(xx)
        return f(xx)
    res = ref.direct(_lib.OBJECTIVE(cb), d, _lib.dptr(lb), _lib.dptr(ub), maxiter, 100000, maxsample)
    return res[0], np.array([res[i + 1] for i in range(d)]), len(trace), np.array(trace)


@pytest.mark.skipif(not os.path.exists(REF_LIBEGO), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", range(48))
def test_batched_direct_equals_live_reference_on_random_problems(seed, capfd):
    d, lb, ub, f, maxiter, maxsample = _random_case(seed)
    rf, rx, rn, rtrace = _run_reference(d, lb, ub, f, maxiter, maxsample)
    capfd.readouterr()                      # the reference prints its termination reason
    for flags in (0, _lib.FLAG_DIRECT_SEQ):
        of, ox, on, opts = _run_ours(d, lb, ub, f, maxiter, maxsample, flags)
        assert of == rf and np.array_equal(ox, rx) and on == rn, (seed, flags)
        if flags:
            assert np.array_equal(opts, rtrace)                       # same call sequence, sample by sample
        else:
            # same multiset of samples (the batched driver reorders them inside an iteration)
            assert np.array_equal(np.sort(opts.view([('', opts.dtype)] * d), axis=0), np.sort(rtrace.view([('', rtrace.dtype)] * d), axis=0))


@pytest.mark.parametrize("seed", range(0, 48, 5))
def test_batched_direct_equals_oracle_restatement_on_random_problems(seed):
    """the same cases against oracle.direct_cpp -- runs on the GPU box too, where only the prebuilt oracle/_ref travels"""
    d, lb, ub, f, maxiter, maxsample = _random_case(seed)
    maxiter = min(maxiter, 10)
    rec = []
    rf, rx, rn = orc.direct_cpp(f, lb, ub, maxiter, maxsample, record=rec)
    of, ox, on, opts = _run_ours(d, lb, ub, f, maxiter, maxsample, _lib.FLAG_DIRECT_SEQ)
    assert of == rf and np.array_equal(ox, rx) and on == rn
    assert np.array_equal(opts, np.array(rec))


@pytest.mark.skipif(not os.path.exists(REF_LIBEGO), reason="oracle/_ref not built")
@pytest.mark.parametrize("d", [17, 20, 24])
def test_batched_direct_equals_live_reference_beyond_16_dims(d, capfd):
    """more than 16 equally long sides: std::sort is no longer an insertion sort (cpp/direct.cpp:194), tie order of the
    dimension sort must still be the reference's -- a symmetric objective makes every probe value tie"""
    lb = np.zeros(d); ub = np.ones(d)
    for f in (lambda x: float(np.sum((x - 0.5) ** 2)), lambda x: float(np.sum(np.cos(7 * x)) + x[3])):
        rf, rx, rn, rtrace = _run_reference(d, lb, ub, f, 6, 100000)
        capfd.readouterr()
        of, ox, on, opts = _run_ours(d, lb, ub, f, 6, 100000, _lib.FLAG_DIRECT_SEQ)
        assert of == rf and np.array_equal(ox, rx) and on == rn
        assert np.array_equal(opts, rtrace)
        of, ox, on, opts = _run_ours(d, lb, ub, f, 6, 100000, 0)
        assert of == rf and np.array_equal(ox, rx) and on == rn


@pytest.mark.skipif(not os.path.exists(REF_LIBEGO), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", range(48))
def test_speculating_direct_equals_live_reference(seed, capfd):
    """IBO_FLAG_DIRECT_SPECULATE (what ibo_acqmax runs with): same FMIN / XMIN / sample count as the reference; every
    reference sample is evaluated, in at most a few more points, and -- the point of it -- in fewer batches"""
    d, lb, ub, f, maxiter, maxsample = _random_case(seed)
    rf, rx, rn, rtrace = _run_reference(d, lb, ub, f, maxiter, maxsample)
    capfd.readouterr()
    of, ox, on, opts = _run_ours(d, lb, ub, f, maxiter, maxsample, _lib.FLAG_DIRECT_SPECULATE)
    assert of == rf and np.array_equal(ox, rx) and on == rn, seed
    ref_set = set(map(bytes, np.ascontiguousarray(rtrace)))
    our_set = set(map(bytes, np.ascontiguousarray(opts)))
    assert ref_set <= our_set
    assert len(opts) <= 3 * rn + 8


def test_one_batch_per_iteration():
    """an optimum in a corner of the box: intervals starting at 0 keep full relative precision, so the centre of a middle
    third and the centre of the full side differ in the last place (the 'unclean' sides of divide()); speculation keeps
    such iterations at one batch for up to two unclean sides"""
    def count(d, flags, maxiter=40):
        batches = []

        def cb(user, n, ndim, X, y):
            A = np.ctypeslib.as_array(X, shape=(n, ndim))
            batches.append(n)
            v = np.sum(np.sin(3 * A) + (A - .4) ** 2, axis=1)
            for i in range(n):
                y[i] = v[i]
        lb = np.zeros(d); ub = np.ones(d)
        fmin = c_double(0); xmin = np.empty(d); ns = c_long(0); it = c_int(0)
        rc = _lib.lib().ibo_direct_batched(_lib.BATCH_OBJECTIVE(cb), None, d, _lib.dptr(lb), _lib.dptr(ub), maxiter, 100000, 10 ** 7,
                                           flags, ctypes.byref(fmin), _lib.dptr(xmin), ctypes.byref(ns), ctypes.byref(it))
        assert rc == 0
        return fmin.value, xmin.copy(), ns.value, it.value, batches
    for d in (2, 3, 6):
        plain = count(d, 0)
        spec = count(d, _lib.FLAG_DIRECT_SPECULATE)
        assert plain[0] == spec[0] and np.array_equal(plain[1], spec[1]) and plain[2] == spec[2] and plain[3] == spec[3]
        assert sum(plain[4]) == plain[2]                      # without the flag exactly the reference's samples are evaluated
        assert len(spec[4]) <= len(plain[4]) <= 2 * plain[3] + 3
        if d == 2:
            assert len(spec[4]) == spec[3] + 2                # centre, first division, one batch per iteration
    # interior optimum: already one batch per iteration without speculation (within a few late batches)
    interior = count(4, 0, maxiter=30)
    assert len(interior[4]) <= interior[3] + 2 + 3


def test_objective_failure_stops_the_driver_at_once():
    """a Python objective that raises must not keep DIRECT iterating on made-up values until maxiter (the callback reports NaN,
    the driver returns IBO_E_OBJECTIVE, the exception is re-raised); a NaN returned by the objective itself is an error too"""
    from ibo_b200 import _lib
    from ibo_b200.utils.optimize import direct
    calls = []

    def boom(P):
        calls.append(len(P))
        if len(calls) == 3:
            raise RuntimeError("objective failed")
        return np.sum((np.asarray(P) - 0.3) ** 2, axis=1)

    with pytest.raises(RuntimeError, match="objective failed"):
        direct(None, [(0., 1.)] * 3, maxiter=10 ** 6, batch_objective=boom)
    assert len(calls) == 3
    with pytest.raises(_lib.IBOError) as ei:
        direct(lambda x: float("nan") if x[0] > 0.6 else float(np.sum(x ** 2)), [(0., 1.)] * 2, maxiter=50)
    assert ei.value.code == _lib.E_OBJECTIVE


def test_maxtime_is_checked_after_every_rectangle_for_scalar_callbacks():
    """the legacy `direct` symbol drives a scalar callback rectangle by rectangle and, like the reference
    (cpp/direct.cpp:493-497), looks at the clock after each rectangle: a slow objective overshoots a 1 s budget by at most
    one rectangle's evaluations, not by a whole iteration"""
    import time
    calls = [0]

    def slow(n, x):
        calls[0] += 1
        time.sleep(0.02)
        return float(sum((x[i] - 0.3) ** 2 for i in range(n)))

    d = 6
    lb = np.zeros(d); ub = np.ones(d)
    t0 = time.time()
    res = _lib.lib().direct(_lib.OBJECTIVE(slow), d, _lib.dptr(lb), _lib.dptr(ub), 10 ** 6, 1, 10 ** 6)
    wall = time.time() - t0
    assert res
    libc = ctypes.CDLL(None); libc.free.argtypes = [ctypes.c_void_p]; libc.free(res)
    # time() has one-second resolution: the budget of 1 s ends between 1 and 2 s after the start, plus at most one rectangle
    # (4 d evaluations of 20 ms) -- an iteration of this problem would be several seconds
    assert 1.0 <= wall <= 2.0 + 4 * d * 0.02 + 0.5, wall
