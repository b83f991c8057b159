"""GPU tests of the device-side rank-1 append (ibo_model_append): the block append of
GaussianProcess.addData (ego/gaussianprocess/__init__.py:300-308) against the oracle's restatement of it and
against a model rebuilt from scratch."""
import numpy as np
import pytest

from oracle import ibo_oracle as orc

pytestmark = pytest.mark.gpu


def _data(N, d, seed):
    rs = np.random.RandomState(seed)
    X = rs.rand(N, d)
    return X, np.sin(3 * X).sum(axis=1) + 0.1 * rs.randn(N)


def _rel(a, b, floor):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))


@pytest.mark.parametrize("N0,adds", [(300, [1, 1, 5]), (127, [1, 1, 1]), (250, [3, 10, 1]), (1, [1, 2]), (640, [64])])
def test_append_matches_oracle_block_append(N0, adds):
    """L, R and the posterior after appends == the oracle's block append == a from-scratch factorisation"""
    from ibo_b200 import _lib
    theta = [0.35, 0.5, 0.45]
    X, Y = _data(N0 + sum(adds), 3, 11)
    m = _lib.Model(_lib.KERNEL_SE_ARD, theta, X[:N0], Y[:N0], 0.1)
    o = orc.GPOracle(orc.KernelSpec(orc.K_SE_ARD, theta, 3), X[:N0], Y[:N0], 0.1)
    n = N0
    for k in adds:
        m.append(X[n:n + k], Y[n:n + k])
        o.add_data(X[n:n + k], Y[n:n + k])
        n += k
    assert m.N == n
    Lfull = np.linalg.cholesky(orc.build_R(o.kernel, X[:n], 0.1))
    Ld = m.matrix(1)
    assert np.max(np.abs(Ld - Lfull)) <= 1e-12
    assert np.max(np.abs(Ld - o.L)) <= 1e-12
    assert np.max(np.abs(m.matrix(0) - o.R)) <= 1e-14
    W = m.matrix(2)
    assert np.max(np.abs(W.dot(Lfull) - np.eye(n))) <= 1e-11
    Xs = np.random.RandomState(5).rand(700, 3)
    mu, s2 = m.posterior(Xs)
    mu_o, s2_o = o.posterior_batch(Xs)
    assert _rel(mu, mu_o, 1e-3) <= 1e-10 and _rel(s2, s2_o, 1e-12) <= 1e-10
    # and against a model built from scratch on the device
    f = _lib.Model(_lib.KERNEL_SE_ARD, theta, X[:n], Y[:n], 0.1)
    mu_f, s2_f = f.posterior(Xs)
    assert _rel(mu, mu_f, 1e-3) <= 1e-11 and _rel(s2, s2_f, 1e-12) <= 1e-11
    sc, _, _, best, bidx = m.score(Xs, _lib.ACQ_EI, Y[:n].max(), 0.01)
    sc_f, _, _, best_f, bidx_f = f.score(Xs, _lib.ACQ_EI, Y[:n].max(), 0.01)
    assert bidx == bidx_f and np.max(np.abs(sc - sc_f)) <= 1e-11


def test_append_matern_and_prior_mean():
    from ibo_b200 import _lib

    class P(object):
        pass
    pr = P()
    rs = np.random.RandomState(2)
    pr.means, pr.beta, pr.theta = rs.rand(5, 2), rs.randn(5), 2.0
    pr.lowerb, pr.width = np.zeros(2), np.ones(2)
    X, Y = _data(400, 2, 3)
    m = _lib.Model(_lib.KERNEL_MATERN3, [0.4, 1.0], X[:390], Y[:390], 0.05, prior=pr)
    for i in range(390, 400):
        m.append(X[i], [Y[i]])
    f = _lib.Model(_lib.KERNEL_MATERN3, [0.4, 1.0], X, Y, 0.05, prior=pr)
    Xs = rs.rand(300, 2)
    a, b = m.posterior(Xs), f.posterior(Xs)
    assert _rel(a[0], b[0], 1e-3) <= 1e-11 and _rel(a[1], b[1], 1e-12) <= 1e-11


def test_append_not_spd_is_reported():
    """prior variance sf2 = 4 above the diagonal 1 + noise (the reference keeps the diagonal at 1 + noise whatever the
    magnitude, App. A): a near-duplicate of a training point makes the enlarged matrix indefinite"""
    from ibo_b200 import _lib
    g = np.arange(6) / 5.0
    X = np.array([[a, b, c] for a in g for b in g for c in g])         # 216 grid points 0.2 apart: off-diagonals ~ 0, SPD
    Y = np.sin(3 * X).sum(axis=1)
    m = _lib.Model(_lib.KERNEL_MATERN3, [0.01, 2.0], X, Y, 0.1)
    with pytest.raises(np.linalg.LinAlgError) as ei:
        m.append(X[:1] + 1e-9, Y[:1])
    assert ei.value.pivot == 217
    with pytest.raises(Exception):
        m.posterior(X[:2])             # the handle was closed


def test_append_rejected_for_laplace_and_legacy_models():
    from ibo_b200 import _lib
    X, Y = _data(40, 2, 1)
    m = _lib.Model(_lib.KERNEL_SE_ISO, [0.3], X, Y, 0.1, Cinv=np.eye(40) * 0.1)
    with pytest.raises(_lib.IBOError):
        m.append(X[:1], Y[:1])


def test_gaussianprocess_add_data_uses_append_for_resident_models():
    """GaussianProcess.addData on a large resident model appends on the device and agrees with batch training"""
    from ibo_b200.gaussianprocess import GaussianProcess
    from ibo_b200.gaussianprocess.kernel import GaussianKernel_ard
    X, Y = _data(520, 4, 9)
    k = GaussianKernel_ard([.3, .4, .5, .6])
    a = GaussianProcess(k, X, Y, noise=0.1)
    b = GaussianProcess(k, X[:500], Y[:500], noise=0.1)
    b.posterior(X[0])                 # make the model resident
    h = b._model
    for i in range(500, 520):
        b.addData(X[i], Y[i])
    assert b._model is h and b._model.N == 520        # appended, not rebuilt
    q = np.random.RandomState(1).rand(50, 4)
    ma, va = a.posteriors(q)
    mb, vb = b.posteriors(q)
    assert _rel(mb, ma, 1e-3) <= 1e-11 and _rel(vb, va, 1e-12) <= 1e-11
    assert np.max(np.abs(a.R - b.R)) <= 1e-14 and np.max(np.abs(a.L - b.L)) <= 1e-12
