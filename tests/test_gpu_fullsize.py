"""Full-size property tests for BASELINE.json configs #4 (N = 8192, Matern-5/2 ARD, d = 10) and #5 (N = 4096, d = 20,
batched DIRECT): size-independent identities that need no O(N^2)-per-candidate oracle pass.

With R = K_offdiag + (1 + nu) I and k(x, x) = 1, the cross-covariance vector of training point i is R e_i - nu e_i, so
    mu(x_i)      = Y_i - nu (inv(R) Y)_i
    sigma^2(x_i) = (1 + nu) - k_i . inv(R) k_i = 2 nu - nu^2 inv(R)_ii
(an encode -> decode round trip through the whole factor), mu is linear in Y, and sigma^2 does not depend on Y."""
import numpy as np
import pytest

from oracle import ibo_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cfg4():
    from ibo_b200 import _lib
    rs = np.random.RandomState(4)
    N, d = 8192, 10
    X = rs.rand(N, d)
    Y = np.sin(2 * X).sum(axis=1)
    hyper = [0.5 + 0.05 * j for j in range(d)] + [1.0]
    return _lib.Model(_lib.KERNEL_MATERN5_ARD, hyper, X, Y, 0.1), X, Y, hyper


def test_config4_training_point_identities(cfg4):
    from ibo_b200 import _lib
    from scipy.linalg import cho_factor, cho_solve
    m, X, Y, hyper = cfg4
    nu = 0.1
    spec = orc.KernelSpec(orc.K_MATERN5_ARD, hyper, 10)
    R = orc.build_R(spec, X, nu)
    cf = cho_factor(R, lower=True)
    RiY = cho_solve(cf, Y)
    idx = np.r_[np.arange(0, 8192, 37), [8191]]
    E = np.zeros((8192, len(idx)))
    E[idx, np.arange(len(idx))] = 1.0
    Rii = cho_solve(cf, E)[idx, np.arange(len(idx))]
    mu, s2 = m.posterior(X[idx], _lib.FLAG_MODE_CPP)
    mu_id = Y[idx] - nu * RiY[idx]
    s2_id = 2 * nu - nu * nu * Rii
    assert np.max(np.abs(mu - mu_id) / np.maximum(np.abs(mu_id), 1e-3)) <= 1e-10
    assert np.max(np.abs(s2 - s2_id) / s2_id) <= 1e-10           # tolerance: north_star's 1e-10 relative (FP64)


def test_config4_linearity_and_y_independence(cfg4):
    from ibo_b200 import _lib
    m, X, Y, hyper = cfg4
    try:
        from scipy.stats import qmc
        Xs = np.ascontiguousarray(qmc.Sobol(d=10, scramble=False).random_base2(12))
    except Exception:
        Xs = np.random.RandomState(1).rand(4096, 10)
    rs = np.random.RandomState(2)
    Y2 = rs.randn(8192)
    m2 = _lib.Model(_lib.KERNEL_MATERN5_ARD, hyper, X, Y2, 0.1)
    m3 = _lib.Model(_lib.KERNEL_MATERN5_ARD, hyper, X, 2.0 * Y - 3.0 * Y2, 0.1)
    mu1, s1 = m.posterior(Xs, _lib.FLAG_MODE_CPP)
    mu2, s2 = m2.posterior(Xs, _lib.FLAG_MODE_CPP)
    mu3, s3 = m3.posterior(Xs, _lib.FLAG_MODE_CPP)
    assert np.array_equal(s1, s2) and np.array_equal(s1, s3)     # the variance never sees Y: same kernels, same bits
    scale = np.abs(2.0 * mu1) + np.abs(3.0 * mu2) + 1e-3
    assert np.max(np.abs(mu3 - (2.0 * mu1 - 3.0 * mu2)) / scale) <= 1e-10
    assert np.all(s1 >= 1e-8) and np.all(s1 <= 1.1 + 1e-12)
    # argmax / scores consistency on the same set
    sc, _, _, best, bidx = m.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, _lib.FLAG_MODE_CPP)
    assert bidx == int(np.argmax(sc)) and best == sc[bidx]
    for o in (m2, m3):
        o.close()


def test_config5_batched_direct_equals_rectangle_by_rectangle():
    """config #5 shape: the batched driver (two GPU batches per iteration) and the reference's call order (one rectangle at a
    time, IBO_FLAG_DIRECT_SEQ) select the same point with the same number of samples"""
    from ibo_b200.acquisition import cdirectGP, maximizeEI
    from ibo_b200.gaussianprocess import GaussianProcess
    from ibo_b200.gaussianprocess.kernel import GaussianKernel_ard
    rs = np.random.RandomState(5)
    N, d = 4096, 20
    X = rs.rand(N, d)
    Y = np.sin(2 * X).sum(axis=1)
    gp = GaussianProcess(GaussianKernel_ard([1.0] * d), X, Y, noise=0.1)
    b = [[0., 1.]] * d
    o1, x1 = maximizeEI(gp, b, xi=0.01, maxiter=30, maxtime=10 ** 6, maxsample=10 ** 9)
    n1 = cdirectGP.last["nsamples"]
    o2, x2 = maximizeEI(gp, b, xi=0.01, maxiter=30, maxtime=10 ** 6, maxsample=10 ** 9, sequential=True)
    n2 = cdirectGP.last["nsamples"]
    assert n1 == n2 and np.array_equal(x1, x2) and abs(o1 - o2) <= 1e-11 * abs(o1)
    # and the returned optimum is the acquisition at the returned point
    from ibo_b200 import _lib
    sc = gp.model.score(np.array([x1]), _lib.ACQ_EI, Y.max(), 0.01, _lib.FLAG_MODE_CPP)[0]
    assert abs(sc[0] - o1) <= 1e-11 * abs(o1)


def test_int8_path_at_its_size_limit_and_fallback_above_it():
    """N = 16384 is the largest model the INT8 path accepts (INT32 head-room of the 8-bit digit sums: 98304 N < 2^31,
    tests/test_int8_model.py): there it must still agree with the FP64 DMMA kernels to rounding; one row-block more and a wide
    batch silently stays on the DMMA kernels (bit-identical to IBO_FLAG_FP64)."""
    from ibo_b200 import _lib
    rs = np.random.RandomState(16)
    d, M = 4, 2600
    Xs = rs.rand(M, d)
    for N, expect_i8 in ((16384, True), (16384 + 128, False)):
        X = rs.rand(N, d)
        Y = np.sin(3 * X).sum(axis=1)
        m = _lib.Model(_lib.KERNEL_SE_ARD, [0.12, 0.15, 0.1, 0.2], X, Y, 0.1)
        a = m.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, _lib.FLAG_MODE_CPP | _lib.FLAG_FP64, want_posterior=True)
        b = m.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, _lib.FLAG_MODE_CPP, want_posterior=True)
        if expect_i8:
            assert not np.array_equal(a[2], b[2])
            assert np.max(np.abs(a[2] - b[2]) / a[2]) <= 1e-11 and np.max(np.abs(a[1] - b[1]) / np.maximum(np.abs(a[1]), 1e-3)) <= 1e-11
            assert np.max(np.abs(a[0] - b[0]) / np.maximum(np.abs(a[0]), 1e-5)) <= 1e-10 and a[4] == b[4]
            # a training point: sigma^2 in [noise, 2 noise], mu pulled towards its observation
            mu, s2 = m.posterior(X[:2600], _lib.FLAG_MODE_CPP)
            assert np.all(s2 >= 0.1 - 1e-9) and np.all(s2 <= 0.2 + 1e-9)
        else:
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
        m.close()
