"""GPU tests of the device Laplace MAP fit (ibo_pref_fit) against the oracle's restatement of the
reference's functional S and of its optimiser call (ego/gaussianprocess/__init__.py:355-386,441-442)."""
import numpy as np
import pytest

from oracle import ibo_oracle as orc

pytestmark = pytest.mark.gpu


def _problem(npts, nprefs, d, seed, theta):
    rs = np.random.RandomState(seed)
    X = rs.rand(npts, d)
    f = np.sin(3 * X).sum(axis=1)
    prefinds = []
    for _ in range(nprefs):
        a, b = rs.choice(npts, 2, replace=False)
        v, u = (a, b) if f[a] > f[b] else (b, a)
        prefinds.append((int(v), int(u), int(rs.rand() < 0.2)))
    spec = orc.KernelSpec(orc.K_SE_ARD, theta, d)
    L = np.linalg.cholesky(orc.build_R(spec, X, 0.1))
    winners = set(v for v, _, _ in prefinds)
    start = np.array([.5 if i in winners else -.5 for i in range(npts)])
    return X, prefinds, L, start


def _device_fit(X, prefinds, theta, start, **kw):
    from ibo_b200 import _lib
    m = _lib.Model(_lib.KERNEL_SE_ARD, theta, X, np.zeros(len(X)), 0.1)
    v = [p[0] for p in prefinds]; u = [p[1] for p in prefinds]; dg = [float(p[2]) for p in prefinds]
    return m.pref_fit(v, u, dg, start, **kw)


def test_device_fit_matches_reference_bfgs_small():
    """small problem: the reference's own optimiser (BFGS, numerical gradients, gtol 1e-5) and the device Newton fit
    reach the same minimiser within the optimiser's tolerance, and the device S is not larger"""
    theta = [0.4, 0.5]
    X, prefinds, L, start = _problem(14, 12, 2, 3, theta)
    y_ref = orc.pref_fit_bfgs(start, prefinds, L)
    y, S, gnorm, iters = _device_fit(X, prefinds, theta, start)
    assert abs(S - orc.pref_S(y, prefinds, L)) <= 1e-11 * max(1.0, abs(S))       # the device S is the reference's functional
    assert S <= orc.pref_S(y_ref, prefinds, L) + 1e-9
    assert np.max(np.abs(y - y_ref)) <= 2e-4
    assert np.max(np.abs(orc.pref_S_grad(y, prefinds, L))) <= 1e-6
    assert iters <= 30


@pytest.mark.parametrize("npts,nprefs,d", [(130, 200, 3), (300, 150, 4), (600, 300, 4)])
def test_device_fit_is_a_stationary_point_of_S(npts, nprefs, d):
    theta = [0.5] * d
    X, prefinds, L, start = _problem(npts, nprefs, d, 7, theta)
    y, S, gnorm, iters = _device_fit(X, prefinds, theta, start)
    assert gnorm <= 1e-9 and iters <= 40
    assert abs(S - orc.pref_S(y, prefinds, L)) <= 1e-10 * max(1.0, abs(S))
    # the analytic gradient (pdf for dCDF/dz; the Chebyshev erf's own derivative differs by ~1e-7) vanishes
    assert np.max(np.abs(orc.pref_S_grad(y, prefinds, L))) <= 1e-6
    # and an independent optimiser started there cannot improve S
    from scipy.optimize import minimize
    r = minimize(orc.pref_S, y, jac=orc.pref_S_grad, args=(prefinds, L), method="L-BFGS-B", options=dict(maxiter=50, gtol=1e-10))
    assert r.fun >= S - 1e-9 * max(1.0, abs(S))


def test_pref_gp_uses_device_fit_and_respects_preferences():
    from ibo_b200.gaussianprocess import PrefGaussianProcess
    from ibo_b200.gaussianprocess.kernel import GaussianKernel_ard
    bounds = [[0., 10.]] * 4
    P = np.array(orc.lhc_sample(bounds, 120, seed=2))
    prefs = []
    for i in range(60):
        a, b = P[2 * i], P[2 * i + 1]
        prefs.append((a, b, 0) if -orc.shekel5(a) > -orc.shekel5(b) else (b, a, 0))
    gp = PrefGaussianProcess(GaussianKernel_ard([5.146, 4.189, 4.622, 5.843]), prefs, noise=0.1)
    assert gp.fit_info is not None and gp.fit_info["gnorm"] <= 1e-9
    idx = dict((tuple(x), i) for i, x in enumerate(gp.X))
    assert all(gp.Y[idx[tuple(v)]] > gp.Y[idx[tuple(u)]] for v, u, _ in prefs)      # MAP latents order every pair


def test_laplace_matrix_is_assembled_and_inverted_on_the_device():
    """ibo_model_create_pref: C from the preference pairs, inv(C) and chol(R + inv(C)) without any host inverse -- against NumPy
    (ego/gaussianprocess/__init__.py:461-498); the dense-C entry (fromLaplace) must give the same factor; repeated pairs and a
    self-pair are handled; a non-SPD C is reported like the reference's LinAlgError."""
    from ibo_b200 import _lib
    rs = np.random.RandomState(12)
    N, d, P = 300, 4, 700
    X = rs.rand(N, d) * 10; Y = rs.randn(N)
    a = rs.randint(0, N, P).astype(np.int32); b = rs.randint(0, N, P).astype(np.int32)
    a[5], b[5] = a[4], b[4]                     # a repeated pair
    b[6] = a[6]                                 # a pair of a point with itself contributes nothing
    w = rs.rand(P) * 3
    theta = [5.146, 4.189, 4.622, 5.843]
    m = _lib.Model(_lib.KERNEL_SE_ARD, theta, X, Y, 0.1, pref=(a, b, w, 5.0))
    C = np.eye(N) * 5.0
    for p in range(P):
        if a[p] != b[p]:
            C[a[p], a[p]] += w[p]; C[b[p], b[p]] += w[p]; C[a[p], b[p]] -= w[p]; C[b[p], a[p]] -= w[p]
    Cinv = np.linalg.inv(C)
    o = orc.GPOracle(orc.KernelSpec(orc.K_SE_ARD, theta, d), X, Y, 0.1, Cinv=Cinv)
    got = m.matrix(3)
    assert np.array_equal(got, got.T) and np.max(np.abs(got - Cinv)) <= 1e-13
    assert np.max(np.abs(m.matrix(1) - o.L)) <= 1e-11
    Xs = rs.rand(500, d) * 10
    sc, mu, s2, best, bidx = m.score(Xs, _lib.ACQ_UCB, Y.max(), 1.3, _lib.FLAG_MODE_PY, want_posterior=True)
    mu_o, s2_o = o.posterior_batch(Xs)
    assert np.max(np.abs(mu - mu_o) / np.maximum(np.abs(mu_o), 1e-3)) <= 1e-10 and np.max(np.abs(s2 - s2_o) / s2_o) <= 1e-10
    m2 = _lib.Model(_lib.KERNEL_SE_ARD, theta, X, Y, 0.1, C=C)
    assert np.max(np.abs(m2.matrix(1) - m.matrix(1))) <= 1e-12
    with pytest.raises(np.linalg.LinAlgError):
        _lib.Model(_lib.KERNEL_SE_ARD, theta, X, Y, 0.1, pref=(a, b, -w, 0.5))
    m.close(); m2.close()
