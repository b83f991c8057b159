"""GPU tests of the experimental INT8-tensor-core emulation of K2 (IBO_FLAG_INT8, ibo_b200/csrc/score_i8.cuh):
sigma^2 through 7 x 7-bit Ozaki slices of W and K* multiplied exactly on tcgen05.mma kind::i8 (INT32 accumulators in TMEM,
FP64 assembly) and mu as k* . alpha must stay inside the same 1e-10 parity bound as the FP64 DMMA path, against the oracle
(reference arithmetic: ego/gaussianprocess/__init__.py:169-228, cpp/optimizeGP.cpp:57-215) and against the DMMA path."""
import numpy as np
import pytest

from oracle import ibo_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _rel(a, b, floor):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def _case(N, d, M, kind="se", prior=False, noise=0.1, seed=0):
    from ibo_b200.gaussianprocess import GaussianProcess
    from ibo_b200.gaussianprocess.kernel import GaussianKernel_ard, MaternKernel3
    from ibo_b200.gaussianprocess.prior import RBFNMeanPrior
    rs = np.random.RandomState(100 + seed + N)
    X = rs.rand(N, d)
    Y = np.sin(3 * X).sum(axis=1)
    Xs = rs.rand(M, d)
    p = op = None
    if prior:
        means = rs.rand(5, d); beta = rs.randn(5) * 0.3
        p = RBFNMeanPrior(means=means, beta=beta, theta=4., lowerb=np.zeros(d), width=np.ones(d))
        op = orc.PriorSpec(means, beta, 4., np.zeros(d), np.ones(d))
    if kind == "se":
        theta = list(0.3 + 0.1 * np.arange(d))
        gp = GaussianProcess(GaussianKernel_ard(theta), X, Y, noise=noise, prior=p)
        o = orc.GPOracle(orc.KernelSpec(orc.K_SE_ARD, theta, d), X, Y, noise, prior=op)
    else:
        gp = GaussianProcess(MaternKernel3([0.7, 1.0]), X, Y, noise=noise, prior=p)
        o = orc.GPOracle(orc.KernelSpec(orc.K_MATERN3, [0.7, 1.0], d), X, Y, noise, prior=op)
    return gp, o, Xs, Y


@pytest.mark.parametrize("N,d,M,kind,prior", [(300, 3, 5000, "se", False), (300, 2, 4096, "matern3", False),
                                              (1000, 20, 3000, "se", False), (200, 2, 4100, "se", True),
                                              (129, 1, 2049, "se", False)])
@pytest.mark.parametrize("mode", ["py", "cpp"])
def test_int8_path_matches_oracle(N, d, M, kind, prior, mode):
    from ibo_b200 import _lib
    gp, o, Xs, Y = _case(N, d, M, kind, prior)
    fl = (_lib.FLAG_MODE_PY if mode == "py" else _lib.FLAG_MODE_CPP) | _lib.FLAG_INT8
    sc, mu, s2, best, bidx = gp.model.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, flags=fl, want_posterior=True)
    if mode == "py":
        mu_o, s2_o = o.posterior_batch(Xs)
    else:
        mu_o, sig = o.posterior_cpp(Xs)
        s2_o = sig ** 2
    ei_o = orc.score(orc.ACQ_EI, mode, mu_o, s2_o, Y.max(), 0.01)
    assert _rel(mu, mu_o, 1e-3) <= TOL
    assert _rel(s2, s2_o, 1e-300) <= TOL
    assert _rel(sc, ei_o, 1e-5) <= TOL
    assert bidx == int(np.argmax(ei_o))


def test_int8_path_matches_dmma_path_at_the_headline_shape():
    from ibo_b200 import _lib
    gp, o, Xs, Y = _case(2048, 6, 40000)
    a = gp.model.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, flags=_lib.FLAG_MODE_CPP, want_posterior=True)
    b = gp.model.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, flags=_lib.FLAG_MODE_CPP | _lib.FLAG_INT8, want_posterior=True)
    assert _rel(b[1], a[1], 1e-3) <= TOL and _rel(b[2], a[2], 1e-300) <= TOL and _rel(b[0], a[0], 1e-5) <= TOL
    assert a[4] == b[4]
    # PI and UCB ride on the same (mu, sigma^2)
    for acq, parm in ((_lib.ACQ_PI, 0.05), (_lib.ACQ_UCB, 1.7)):
        a = gp.model.score(Xs, acq, Y.max(), parm, flags=_lib.FLAG_MODE_CPP)
        b = gp.model.score(Xs, acq, Y.max(), parm, flags=_lib.FLAG_MODE_CPP | _lib.FLAG_INT8)
        assert _rel(b[0], a[0], 1e-5) <= TOL and a[4] == b[4]


def test_int8_values_do_not_depend_on_the_batch():
    """a candidate's value is a function of (model, x) only: same bits wherever it sits in whatever batch"""
    from ibo_b200 import _lib
    gp, o, Xs, Y = _case(700, 4, 9000)
    fl = _lib.FLAG_MODE_CPP | _lib.FLAG_INT8
    full = gp.model.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, flags=fl, want_posterior=True)
    perm = np.random.RandomState(5).permutation(len(Xs))[:4000]
    sub = gp.model.score(np.ascontiguousarray(Xs[perm]), _lib.ACQ_EI, Y.max(), 0.01, flags=fl, want_posterior=True)
    assert np.array_equal(sub[0], full[0][perm]) and np.array_equal(sub[1], full[1][perm]) and np.array_equal(sub[2], full[2][perm])
    # exact duplicates: lowest index wins the argmax
    dup = np.vstack([Xs[:3000], Xs[:3000]])
    r = gp.model.score(dup, _lib.ACQ_EI, Y.max(), 0.01, flags=fl)
    assert r[4] == int(np.argmax(full[0][:3000]))


def test_int8_follows_append_and_ignores_small_batches():
    from ibo_b200 import _lib
    gp, o, Xs, Y = _case(250, 3, 3000)
    fl = _lib.FLAG_MODE_PY | _lib.FLAG_INT8
    gp.model.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, flags=fl)
    rs = np.random.RandomState(9)
    Xn = rs.rand(10, 3); Yn = np.sin(3 * Xn).sum(axis=1)
    gp.addData(Xn, Yn)                       # rank-1 appends on the resident model: the slices of W are rebuilt
    o.add_data(Xn, Yn)
    sc, mu, s2, best, bidx = gp.model.score(Xs, _lib.ACQ_EI, max(gp.Y), 0.01, flags=fl, want_posterior=True)
    mu_o, s2_o = o.posterior_batch(Xs)
    assert _rel(mu, mu_o, 1e-3) <= TOL and _rel(s2, s2_o, 1e-300) <= TOL
    # batches of <= 2048 candidates take the FP64 latency shapes whatever the flag says
    a = gp.model.score(Xs[:500], _lib.ACQ_EI, max(gp.Y), 0.01, flags=_lib.FLAG_MODE_PY, want_posterior=True)
    b = gp.model.score(Xs[:500], _lib.ACQ_EI, max(gp.Y), 0.01, flags=fl, want_posterior=True)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])


@pytest.mark.skipif(__import__("os").environ.get("IBO_EXPERIMENTAL_TESTS") != "1",
                    reason="IBO_FLAG_INT8_G9 (eighth accumulator group) has not been run on a device yet; set IBO_EXPERIMENTAL_TESTS=1")
def test_int8_g9_variant_is_tighter():
    from ibo_b200 import _lib
    gp, o, Xs, Y = _case(2048, 6, 20000)
    a = gp.model.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, flags=_lib.FLAG_MODE_CPP, want_posterior=True)
    b = gp.model.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, flags=_lib.FLAG_MODE_CPP | _lib.FLAG_INT8, want_posterior=True)
    c = gp.model.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, flags=_lib.FLAG_MODE_CPP | _lib.FLAG_INT8_G9, want_posterior=True)
    assert _rel(c[2], a[2], 1e-300) <= 0.1 * max(_rel(b[2], a[2], 1e-300), 1e-13)
    assert _rel(c[0], a[0], 1e-5) <= 1e-11 and a[4] == c[4]


@pytest.mark.skipif(__import__("os").environ.get("IBO_EXPERIMENTAL_TESTS") != "1",
                    reason="IBO_FLAG_INT8_D8 (8-bit digits) has not been run on a device yet; set IBO_EXPERIMENTAL_TESTS=1")
@pytest.mark.parametrize("N,d,M,kind,prior", [(300, 3, 5000, "se", False), (300, 2, 4096, "matern3", False), (200, 2, 4100, "se", True),
                                              (2048, 6, 20000, "se", False)])
def test_int8_d8_variant_is_two_digits_tighter(N, d, M, kind, prior):
    """8-bit digits: same 28 products, sigma^2 within ~1e-13 of the FP64 path (CPU model: tests/test_int8_model.py)"""
    from ibo_b200 import _lib
    gp, o, Xs, Y = _case(N, d, M, kind, prior)
    a = gp.model.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, flags=_lib.FLAG_MODE_CPP, want_posterior=True)
    c = gp.model.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, flags=_lib.FLAG_MODE_CPP | _lib.FLAG_INT8_D8, want_posterior=True)
    assert _rel(c[1], a[1], 1e-3) <= 1e-12 and _rel(c[2], a[2], 1e-300) <= 1e-12 and _rel(c[0], a[0], 1e-5) <= 1e-11
    assert a[4] == c[4]


@pytest.mark.skipif(__import__("os").environ.get("IBO_EXPERIMENTAL_TESTS") != "1",
                    reason="IBO_FLAG_INT8_S6 (six 8-bit digits, 21 products) has not been run on a device yet; set IBO_EXPERIMENTAL_TESTS=1")
@pytest.mark.parametrize("N,d,M,kind,prior", [(300, 3, 5000, "se", False), (200, 2, 4100, "se", True), (2048, 6, 20000, "se", False)])
def test_int8_s6_variant_keeps_the_parity_bound(N, d, M, kind, prior):
    from ibo_b200 import _lib
    gp, o, Xs, Y = _case(N, d, M, kind, prior)
    a = gp.model.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, flags=_lib.FLAG_MODE_CPP, want_posterior=True)
    c = gp.model.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, flags=_lib.FLAG_MODE_CPP | _lib.FLAG_INT8_S6, want_posterior=True)
    assert _rel(c[1], a[1], 1e-3) <= TOL and _rel(c[2], a[2], 1e-300) <= TOL and _rel(c[0], a[0], 1e-5) <= TOL
    assert a[4] == c[4]
