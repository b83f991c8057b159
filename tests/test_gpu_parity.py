"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on seeded inputs.

Tolerances (BASELINE.json north_star): mu, sigma and EI within 1e-10 relative (FP64), identical argmax.
EI is compared as |dEI| <= 1e-10 * max(|EI|, EI_FLOOR): below EI_FLOOR the reference formula
ydiff*Phi(Z) + s*phi(Z) itself cancels catastrophically (Phi(Z) = 0.5*(1+erf(.)) keeps only ~1e-16
*absolute* accuracy), so a relative comparison of two correct implementations is meaningless there.
"""
import numpy as np
import pytest

from oracle import ibo_oracle as orc

pytestmark = pytest.mark.gpu

RTOL = 1e-10
EI_FLOOR = 1e-5


def _model(kind, hyper, X, Y, noise, **kw):
    from ibo_b200 import _lib
    return _lib.Model(kind, hyper, X, Y, noise, **kw)


def _check_scores(got, want, floor=EI_FLOOR):
    err = np.abs(got - want) / np.maximum(np.abs(want), floor)
    assert err.max() <= RTOL, "max scaled error %.3e at %d" % (err.max(), err.argmax())


CASES = [
    # (name, kind, hyper, N, d, M, noise)
    ("se_ard_small", orc.K_SE_ARD, [3.4, 10.0], 50, 2, 300, 0.1),
    ("se_iso", orc.K_SE_ISO, [0.3], 130, 4, 517, 0.1),
    ("matern3", orc.K_MATERN3, [1.0, 1.0], 257, 2, 1000, 0.01),
    ("matern5", orc.K_MATERN5, [0.8, 1.0], 300, 3, 700, 0.1),
    ("matern5_ard", orc.K_MATERN5_ARD, [0.5, 0.55, 0.6, 0.65, 0.7, 1.0], 384, 5, 2000, 0.1),
    ("se_ard_hartman", orc.K_SE_ARD, [.53, .57, 2.5, .34, .27, .35], 1024, 6, 4096, 0.1),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("mode", ["py", "cpp"])
def test_posterior_and_ei_match_oracle(case, mode):
    from ibo_b200 import _lib
    name, kind, hyper, N, d, M, noise = case
    rs = np.random.RandomState(abs(hash(name)) % 1000)
    X = rs.rand(N, d)
    Y = np.sin(2 * X).sum(axis=1) if d != 6 else orc.hartman6_neg(X)
    Xs = rs.rand(M, d)
    Xs[:5] = X[:5]                      # candidates on top of training points
    Xs[5:10] = X[5:10] + 1e-9
    kern = orc.KernelSpec(kind, hyper, d)
    gp = orc.GPOracle(kern, X, Y, noise)
    floor = 10e-8 if mode == "py" else 1e-8
    mu_o, s2_o = gp.posterior_batch(Xs, floor=floor)
    ymax, xi = Y.max(), 0.01
    m = _model(kind, hyper, X, Y, noise)
    flags = _lib.FLAG_MODE_PY if mode == "py" else _lib.FLAG_MODE_CPP
    sc, mu, s2, best, bidx = m.score(Xs, _lib.ACQ_EI, ymax, xi, flags, want_posterior=True)
    assert np.max(np.abs(mu - mu_o) / np.maximum(np.abs(mu_o), 1e-3)) <= RTOL
    assert np.max(np.abs(np.sqrt(s2) - np.sqrt(s2_o)) / np.sqrt(s2_o)) <= RTOL
    ei_o = orc.score(orc.ACQ_EI, mode, mu_o, s2_o, ymax, xi)
    _check_scores(sc, ei_o)
    assert bidx == int(np.argmax(ei_o))
    assert best == sc[bidx]
    for acq in (orc.ACQ_PI, orc.ACQ_UCB):
        parm = xi if acq == orc.ACQ_PI else 1.7
        sc2, _, _, _, bidx2 = m.score(Xs, acq, ymax, parm, flags)
        want = orc.score(acq, mode, mu_o, s2_o, ymax, parm)
        _check_scores(sc2, want)
        assert bidx2 == int(np.argmax(want))
    m.close()


def test_factor_matches_numpy_cholesky():
    rs = np.random.RandomState(3)
    N, d = 300, 4
    X = rs.rand(N, d)
    Y = rs.randn(N)
    kern = orc.KernelSpec(orc.K_SE_ARD, [0.4] * d, d)
    gp = orc.GPOracle(kern, X, Y, 0.1)
    m = _model(orc.K_SE_ARD, [0.4] * d, X, Y, 0.1)
    A = m.matrix(0)
    L = m.matrix(1)
    W = m.matrix(2)
    assert np.array_equal(A, A.T)
    assert np.max(np.abs(A - gp.R)) < 1e-14
    assert np.max(np.abs(L - gp.L)) < 1e-12
    assert np.max(np.abs(W.dot(gp.L) - np.eye(N))) < 1e-11
    m.close()


@pytest.mark.parametrize("N", [300, 640, 1100])
@pytest.mark.parametrize("pair", [0, 1])
def test_factor_schedules_match_numpy_cholesky(N, pair):
    """both schedules of the build (block columns one at a time / in pairs with 256-deep updates; option chol_pair) on even and odd
    block counts (3, 5, 9 row-blocks)"""
    from ibo_b200 import _lib
    rs = np.random.RandomState(N)
    d = 3
    X = rs.rand(N, d)
    Y = rs.randn(N)
    gp = orc.GPOracle(orc.KernelSpec(orc.K_MATERN5, [0.5, 1.0], d), X, Y, 0.05)
    _lib.set_option("chol_pair", pair)
    try:
        m = _model(orc.K_MATERN5, [0.5, 1.0], X, Y, 0.05)
        L = m.matrix(1)
        W = m.matrix(2)
        m.close()
    finally:
        _lib.set_option("chol_pair", -1)
    assert np.max(np.abs(L - gp.L)) < 1e-12
    assert np.max(np.abs(W.dot(gp.L) - np.eye(N))) < 1e-10
    assert np.array_equal(np.triu(L, 1), np.zeros_like(L)) and np.array_equal(np.triu(W, 1), np.zeros_like(W))


EDGE = [
    # (name, kind, hyper, N, d, M, noise, box)  -- shapes around the 128-row blocks, d = 1, d > 32 (direct-difference K1),
    # and badly scaled inputs that trip K1's cancellation guard (|x/theta|^2 >> 256 -> per-pair direct differences)
    ("one_observation", orc.K_SE_ISO, [0.7], 1, 2, 40, 0.1, 1.0),
    ("n128_exact", orc.K_SE_ARD, [0.4, 0.6, 0.8], 128, 3, 129, 0.1, 1.0),
    ("n129", orc.K_MATERN3, [0.9, 1.0], 129, 3, 127, 0.1, 1.0),
    ("d1", orc.K_SE_ISO, [0.15], 90, 1, 333, 0.05, 1.0),
    ("d40_direct_kernel", orc.K_SE_ARD, [2.0] * 40, 200, 40, 300, 0.1, 1.0),
    ("d20", orc.K_SE_ARD, [1.0] * 20, 300, 20, 500, 0.1, 1.0),
    ("badly_scaled_se", orc.K_SE_ISO, [0.05], 150, 3, 400, 0.1, 10.0),
    ("badly_scaled_matern5", orc.K_MATERN5, [0.04, 1.0], 150, 2, 400, 0.1, 10.0),
]


@pytest.mark.parametrize("case", EDGE, ids=[c[0] for c in EDGE])
def test_edge_shapes_and_scaling(case):
    from ibo_b200 import _lib
    name, kind, hyper, N, d, M, noise, box = case
    rs = np.random.RandomState(len(name))
    X = rs.rand(N, d) * box
    Y = np.cos(3 * X / box).sum(axis=1)
    Xs = rs.rand(M, d) * box
    k = min(N, M // 2)
    Xs[:k] = X[:k] + rs.randn(k, d) * 0.01 * np.min(hyper)      # near training points so that k* is not ~0
    gp = orc.GPOracle(orc.KernelSpec(kind, hyper, d), X, Y, noise)
    mu_o, s2_o = gp.posterior_batch(Xs, floor=1e-8)
    m = _model(kind, hyper, X, Y, noise)
    sc, mu, s2, best, bidx = m.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, _lib.FLAG_MODE_CPP, want_posterior=True)
    assert np.max(np.abs(mu - mu_o) / np.maximum(np.abs(mu_o), 1e-3)) <= RTOL
    assert np.max(np.abs(np.sqrt(s2) - np.sqrt(s2_o)) / np.sqrt(s2_o)) <= RTOL
    ei_o = orc.score(orc.ACQ_EI, "cpp", mu_o, s2_o, Y.max(), 0.01)
    _check_scores(sc, ei_o)
    assert bidx == int(np.argmax(ei_o))
    m.close()
