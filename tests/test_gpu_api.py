"""GPU tests of the drop-in Python surface (GaussianProcess / PrefGaussianProcess / EI / gallery) and of
size-independent properties at BASELINE.json's full sizes."""
import numpy as np
import pytest

from oracle import ibo_oracle as orc

pytestmark = pytest.mark.gpu


def _gp(kernel, X, Y, **kw):
    from ibo_b200.gaussianprocess import GaussianProcess
    return GaussianProcess(kernel, X, Y, **kw)


def test_posterior_api_matches_numpy_path():
    """GaussianProcess.posterior / posteriors / mu / EI.negf / PI.negf / UCB.negf vs the scalar-faithful oracle"""
    from ibo_b200.acquisition import EI, PI, UCB
    from ibo_b200.gaussianprocess.kernel import GaussianKernel_iso
    bounds = [[0., 10.]] * 4
    X = np.array(orc.lhc_sample(bounds, 10, seed=0))
    Y = np.array([-orc.shekel5(x) for x in X])
    gp = _gp(GaussianKernel_iso([0.3]), X, Y, noise=0.1)
    o = orc.GPOracle(orc.KernelSpec(orc.K_SE_ISO, [0.3], 4), X, Y, 0.1)
    pts = np.array(orc.lhc_sample(bounds, 7, seed=3))
    for x in pts:
        m, v = gp.posterior(x)
        mo, vo = o.posterior_scalar(x)
        assert abs(m - mo) <= 1e-10 * max(abs(mo), 1e-3) and abs(v - vo) <= 1e-10 * vo
        assert gp.mu(x) == m and gp.negmu(x) == -m
        ei = EI(gp, xi=0.01)
        assert abs(-ei.negf(x) - float(orc.ei_py(mo, vo, Y.max(), 0.01))) <= 1e-10 * max(abs(float(orc.ei_py(mo, vo, Y.max(), 0.01))), 1e-5)
        pi = PI(gp, xi=0.05)
        assert abs(pi.f(x) - float(orc.pi_py(mo, vo, Y.max(), 0.05))) <= 1e-10
        ucb = UCB(gp, 4)
        assert abs(ucb.f(x) - float(orc.ucb_py(mo, vo, orc.ucb_sbeta_py(len(Y), 4)))) <= 1e-10
    M, V = gp.posteriors(pts)
    assert M.shape == (7,) and V.shape == (7,)
    assert (M[0], V[0]) == gp.posterior(pts[0])
    # R and L attributes
    assert np.allclose(gp.R, o.R, atol=1e-14) and np.allclose(gp.L, o.L, atol=1e-12)


def test_add_data_batch_equals_sequential():
    """ego/unittest_GP.py:109-156: adding data in one batch or one by one gives the same R and the same posterior"""
    from ibo_b200.gaussianprocess.kernel import GaussianKernel_ard
    rs = np.random.RandomState(4)
    X = rs.rand(40, 3); Y = np.sin(4 * X).sum(axis=1)
    a = _gp(GaussianKernel_ard([.3, .4, .5]), X, Y)
    from ibo_b200.gaussianprocess import GaussianProcess
    b = GaussianProcess(GaussianKernel_ard([.3, .4, .5]))
    assert b.posterior(X[0]) == (0.0, 1.0)
    for x, y in zip(X, Y):
        b.addData(x, y)
    assert np.array_equal(a.R, b.R)
    q = rs.rand(9, 3)
    assert np.array_equal(a.posteriors(q)[0], b.posteriors(q)[0])
    assert np.array_equal(a.posteriors(q)[1], b.posteriors(q)[1])
    assert a.getYfromX(X[3]) == Y[3] and a.getYfromX(np.ones(3) * 7) is None


def test_training_point_bounds():
    """ego/unittest_GP.py:94-98: at training points sigma^2 < 1/(1+noise)... here: below noise-level bound, mu near y"""
    from ibo_b200.gaussianprocess.kernel import GaussianKernel_iso
    bounds = [[0., 10.]] * 4
    X = np.array(orc.lhc_sample(bounds, 30, seed=1))
    Y = np.array([-orc.shekel5(x) for x in X])
    gp = _gp(GaussianKernel_iso([0.3]), X, Y, noise=0.01)
    M, V = gp.posteriors(X)
    assert np.all(V < 1.0 / (1 + 0.01)) and np.all(np.abs(M - Y) < 2 * 0.01 + 0.05)


def test_pref_gp_laplace_factor_and_aug_variance():
    """fitted PrefGaussianProcess (X, Y, C given): L = chol(R + inv(C)); addObservationPoint switches the variance to augL"""
    from scipy.linalg import solve_triangular
    from ibo_b200.gaussianprocess import PrefGaussianProcess
    from ibo_b200.gaussianprocess.kernel import GaussianKernel_ard
    rs = np.random.RandomState(8)
    N, d = 60, 4
    X = rs.rand(N, d) * 10
    Y = rs.randn(N)
    A = rs.randn(N, N) * 0.05
    C = np.eye(N) * 5 + A.dot(A.T)
    theta = [5.146, 4.189, 4.622, 5.843]
    gp = PrefGaussianProcess.fromLaplace(GaussianKernel_ard(theta), X, Y, C, noise=0.1)
    Cinv = np.linalg.inv(C)
    o = orc.GPOracle(orc.KernelSpec(orc.K_SE_ARD, theta, d), X, Y, 0.1, Cinv=Cinv)
    assert np.allclose(gp.L, o.L, atol=1e-12)
    q = rs.rand(50, d) * 10
    mu, s2 = gp.posteriors(q)
    mo, so = o.posterior_batch(q)
    assert np.max(np.abs(mu - mo) / np.maximum(np.abs(mo), 1e-3)) <= 1e-10
    assert np.max(np.abs(s2 - so) / so) <= 1e-10
    # observation points (ego/gaussianprocess/__init__.py:502-519)
    P = rs.rand(3, d) * 10
    gp.addObservationPoint(P[:2]); gp.addObservationPoint(P[2])
    augX = np.r_[X, P]
    augR = orc.build_R(o.kernel, augX, 0.1)
    pad = np.zeros_like(augR); pad[:N, :N] = Cinv
    o.augL = np.linalg.cholesky(augR + pad); o.augX = augX
    mu2, s22 = gp.posteriors(q)
    mo2, so2 = o.posterior_batch(q)
    assert np.array_equal(mu2, mu)
    assert np.max(np.abs(s22 - so2) / so2) <= 1e-10
    assert np.all(s22 <= s2 + 1e-12)
    assert np.allclose(gp.augL, o.augL, atol=1e-12)


def test_pref_gp_fit_orders_latents():
    """ego/unittest_GP.py:398-478 property: the latent means respect the preferences"""
    from ibo_b200.gaussianprocess import PrefGaussianProcess
    from ibo_b200.gaussianprocess.kernel import GaussianKernel_iso
    prefs = [(np.array([.2]), np.array([.5]), 0), (np.array([.5]), np.array([.9]), 0), (np.array([.2]), np.array([.9]), 1)]
    gp = PrefGaussianProcess(GaussianKernel_iso([.3]), prefs)
    assert gp.mu(np.array([.2])) > gp.mu(np.array([.5])) > gp.mu(np.array([.9]))
    assert gp.C.shape == (3, 3) and np.allclose(gp.C, gp.C.T)
    with pytest.raises(NotImplementedError):
        gp.addData([0.1], [1.0])


def test_config3_shape_preference_gallery():
    """BASELINE config #3 at reduced size: PrefGaussianProcess (Laplace) d=4 on Shekel5 preferences, gallery of 4"""
    from ibo_b200.acquisition import fastUCBGallery
    from ibo_b200.gaussianprocess import PrefGaussianProcess
    from ibo_b200.gaussianprocess.kernel import GaussianKernel_ard
    bounds = [[0., 10.]] * 4
    P = np.array(orc.lhc_sample(bounds, 120, seed=2))
    prefs = []
    for i in range(60):
        a, b = P[2 * i], P[2 * i + 1]
        fa, fb = -orc.shekel5(a), -orc.shekel5(b)
        prefs.append((a, b, 0) if fa > fb else (b, a, 0))
    gp = PrefGaussianProcess(GaussianKernel_ard([5.146, 4.189, 4.622, 5.843]), prefs, noise=0.1)
    assert gp.X.shape == (120, 4) and gp.C.shape == (120, 120)
    mu = gp.posteriors(gp.X)[0]
    idx = dict((tuple(x), i) for i, x in enumerate(gp.X))
    ok = sum(mu[idx[tuple(v)]] > mu[idx[tuple(u)]] for v, u, _ in prefs)
    assert ok >= 48                      # the smooth posterior mean respects most preferences (54/60 measured)
    # L is the factor of R + inv(C)
    o = orc.GPOracle(orc.KernelSpec(orc.K_SE_ARD, [5.146, 4.189, 4.622, 5.843], 4), gp.X, gp.Y, 0.1, Cinv=np.linalg.inv(gp.C))
    assert np.allclose(gp.L, o.L, atol=1e-11)
    gal = fastUCBGallery(gp, bounds, 4, seed=3)
    assert len(gal) == 4
    for x in gal:
        assert all(0 <= v <= 10 for v in x)
    for i in range(4):
        for j in range(i):
            assert np.linalg.norm(gal[i] - gal[j]) > .5


def test_not_spd_is_reported():
    from ibo_b200 import _lib
    X = np.array([[0.0], [1.0], [2.0]]); Y = np.zeros(3)
    Cinv = -5.0 * np.eye(3)
    with pytest.raises(np.linalg.LinAlgError) as ei:
        _lib.Model(_lib.KERNEL_SE_ISO, [1.0], X, Y, 0.1, Cinv=Cinv)
    assert ei.value.pivot == 1


def test_fast_ucb_gallery_properties():
    """ego/unittest_IBO.py:844-870: gallery points inside the bounds, fixed dimension pinned"""
    from ibo_b200.acquisition import fastUCBGallery
    from ibo_b200.gaussianprocess.kernel import GaussianKernel_ard
    rs = np.random.RandomState(2)
    bounds = [[0., 0.], [0., 4.], [-1., 3.]]
    X = np.c_[np.zeros(12), rs.rand(12) * 4, rs.rand(12) * 4 - 1]
    Y = np.sin(X[:, 1]) + np.cos(X[:, 2])
    gp = _gp(GaussianKernel_ard([1.0, 1.0, 1.0]), X, Y)
    gal = fastUCBGallery(gp, bounds, 4, seed=3)
    assert len(gal) == 4
    for x in gal:
        assert x[0] == 0.0 and 0 <= x[1] <= 4 and -1 <= x[2] <= 3
    for i in range(4):
        for j in range(i):
            assert np.linalg.norm(gal[i] - gal[j]) > .5
    # deterministic with a seed
    gal2 = fastUCBGallery(gp, bounds, 4, seed=3)
    assert all(np.array_equal(a, b) for a, b in zip(gal, gal2))
    # no data, no prior -> starts from the centre
    from ibo_b200.gaussianprocess import GaussianProcess
    empty = GaussianProcess(GaussianKernel_ard([1.0, 1.0]))
    g3 = fastUCBGallery(empty, [[0., 2.], [0., 2.]], 2, seed=1)
    assert np.array_equal(g3[0], np.array([1.0, 1.0])) and len(g3) == 2


def test_python_direct_route_uses_python_arithmetic():
    from ibo_b200.acquisition import maximizeEI
    from ibo_b200.gaussianprocess.kernel import GaussianKernel_iso
    bounds = [[0., 10.]] * 4
    X = np.array(orc.lhc_sample(bounds, 10, seed=0))
    Y = np.array([-orc.shekel5(x) for x in X])
    gp = _gp(GaussianKernel_iso([0.3]), X, Y, noise=0.1)
    a, ax = maximizeEI(gp, bounds, xi=0.01, maxiter=15, useCDIRECT=True)
    b, bx = maximizeEI(gp, bounds, xi=0.01, maxiter=15, useCDIRECT=False)
    # ego/unittest_IBO.py:427-474: the routes agree to 4 decimals (erf approximations differ by ~3e-7)
    assert abs(a - b) < 5e-5 and np.allclose(ax, bx, atol=1e-4)


# ---- full-size properties (BASELINE.json config #2 shape) ------------------------------------
@pytest.fixture(scope="module")
def big():
    from ibo_b200 import _lib
    rs = np.random.RandomState(0)
    N, d = 2048, 6
    X = rs.rand(N, d)
    Y = orc.hartman6_neg(X)
    theta = [.53, .57, 2.5, .34, .27, .35]
    return _lib.Model(_lib.KERNEL_SE_ARD, theta, X, Y, 0.1), X, Y, theta


def test_full_size_subsample_vs_oracle(big):
    from ibo_b200 import _lib
    m, X, Y, theta = big
    M = 1 << 17
    Xs = np.random.RandomState(1).rand(M, 6)
    sc, mu, s2, best, bidx = m.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, _lib.FLAG_MODE_CPP, want_posterior=True)
    assert best == sc.max() and bidx == int(np.argmax(sc))
    sel = np.r_[np.arange(0, M, 509), [bidx], np.argsort(sc)[-16:]]
    o = orc.GPOracle(orc.KernelSpec(orc.K_SE_ARD, theta, 6), X, Y, 0.1)
    mo, so = o.posterior_batch(Xs[sel], floor=1e-8)
    eo = orc.score(orc.ACQ_EI, "cpp", mo, so, Y.max(), 0.01)
    assert np.max(np.abs(mu[sel] - mo) / np.maximum(np.abs(mo), 1e-3)) <= 1e-10
    assert np.max(np.abs(np.sqrt(s2[sel]) - np.sqrt(so)) / np.sqrt(so)) <= 1e-10
    assert np.max(np.abs(sc[sel] - eo) / np.maximum(np.abs(eo), 1e-5)) <= 1e-10
    assert np.all(s2 >= 1e-8) and np.all(s2 <= 1.1 + 1e-12)


def test_scores_do_not_depend_on_batch_shape(big):
    """a candidate's value is a function of (model, x) only.  Large batches (INT8 tensor-core path by default, or the throughput
    shape of the DMMA K2 with IBO_FLAG_FP64) are bit-identical across batch size / position / chunking / grouping; a small batch
    (M <= 2048) runs a latency shape of the DMMA K2 chosen from its size: there the value is bit-identical across position and
    order within the batch size, and equal to the large-batch value to rounding."""
    from ibo_b200 import _lib
    m, X, Y, theta = big
    Xs = np.random.RandomState(5).rand(70000, 6)
    for fl in (_lib.FLAG_MODE_CPP | _lib.FLAG_FP64, _lib.FLAG_MODE_CPP):
        full = m.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, fl)[0]
        for lo, hi in [(300, 41000), (2000, 4100), (60000, 70000)]:
            part = m.score(Xs[lo:hi], _lib.ACQ_EI, Y.max(), 0.01, fl)[0]
            assert np.array_equal(part, full[lo:hi])
    perm = np.random.RandomState(6).permutation(5000)
    shuf = m.score(Xs[perm], _lib.ACQ_EI, Y.max(), 0.01, _lib.FLAG_MODE_CPP)[0]
    assert np.array_equal(shuf, full[perm])
    for lo, hi in [(0, 1), (5, 133), (1000, 1700), (69000, 70000), (7, 40), (100, 164), (0, 2048)]:
        part = m.score(Xs[lo:hi], _lib.ACQ_EI, Y.max(), 0.01, _lib.FLAG_MODE_CPP)[0]
        assert np.max(np.abs(part - full[lo:hi]) / np.maximum(np.abs(full[lo:hi]), 1e-5)) <= 5e-11
        again = m.score(Xs[lo:hi][::-1].copy(), _lib.ACQ_EI, Y.max(), 0.01, _lib.FLAG_MODE_CPP)[0]
        assert np.array_equal(again[::-1], part)
        if hi - lo > 1:      # same size, different neighbours
            other = m.score(np.r_[Xs[lo:lo + 1], Xs[40000:40000 + hi - lo - 1]], _lib.ACQ_EI, Y.max(), 0.01, _lib.FLAG_MODE_CPP)[0]
            assert other[0] == part[0]


def test_resident_candidates_and_ties(big):
    from ibo_b200 import _lib
    m, X, Y, theta = big
    Xs = np.random.RandomState(9).rand(3000, 6)
    Xs[2000] = Xs[17]; Xs[2999] = Xs[17]            # exact duplicates: lowest index must win any tie
    c = _lib.ResidentCandidates(m, Xs)
    out = np.empty(3000)
    best, bidx, ms = c.score(_lib.ACQ_UCB, Y.max(), 2.0, _lib.FLAG_MODE_CPP, scores_out=out)
    assert out[17] == out[2000] == out[2999]
    assert bidx == int(np.argmax(out)) and best == out[bidx] and ms > 0
    dup = np.repeat(Xs[17:18], 700, axis=0)
    _, _, _, b2, i2 = m.score(dup, _lib.ACQ_EI, Y.max(), 0.01, _lib.FLAG_MODE_CPP)
    assert i2 == 0
    c.close()
    # empty batch
    sc, mu, s2, b, i = m.score(np.zeros((0, 6)), _lib.ACQ_EI, Y.max(), 0.01)
    assert i == -1 and len(sc) == 0


def test_kstar_exp_is_within_one_ulp_of_libdevice():
    """K1's own exp (constant-bank Taylor/Cody-Waite, ibo_b200/csrc/score.cu: exp_nonpos) against libdevice exp"""
    import ctypes
    from ibo_b200 import _lib
    rs = np.random.RandomState(0)
    x = np.concatenate([-rs.rand(2_000_000) * 50.0, -rs.rand(500_000) * 745.0, -np.exp(-rs.rand(500_000) * 40.0),
                        [0.0, -0.0, -1e-300, -706.9, -707.0, -707.1, -745.0, -1e4, -np.inf]])
    fast, ref = np.empty_like(x), np.empty_like(x)
    _lib.check(_lib.lib().ibo_debug_exp(0, _lib.dptr(x), len(x), _lib.dptr(fast), _lib.dptr(ref)))
    keep = x >= -707.0                                  # below: flushed to zero by design (values < 1e-307)
    assert np.all(fast[~keep] == 0.0) and np.all(ref[~keep] < 1e-306)
    ulp = np.spacing(ref[keep])
    assert np.max(np.abs(fast[keep] - ref[keep]) / ulp) <= 1.0
    assert fast[x == 0.0].tolist() == [1.0, 1.0]
    # and against the correctly rounded value for a sample (math.exp is < 1 ulp)
    import math
    idx = rs.choice(np.flatnonzero(keep), 20000, replace=False)
    exact = np.array([math.exp(v) for v in x[idx]])
    assert np.max(np.abs(fast[idx] - exact) / np.spacing(exact)) <= 1.0


# ---- fastUCBGallery against the oracle restatement of ego/acquisition/gallery.py:42-135 ---------------------
@pytest.mark.parametrize("with_prior", [False, True])
def test_fast_ucb_gallery_matches_oracle_restatement(with_prior, monkeypatch):
    """same model, bounds and latin-hypercube seed -> the same gallery, slot by slot: DIRECT (xi=.3) point, the 0.5
    distance rule, the EI(xi=.4) scan over the LHS candidates / prior means, and the hallucinated append."""
    import functools
    import ibo_b200.acquisition as acq
    from ibo_b200.gaussianprocess.kernel import GaussianKernel_ard
    from ibo_b200.gaussianprocess.prior import RBFNMeanPrior
    rs = np.random.RandomState(11)
    bounds = [[0., 4.], [-1., 3.], [0., 2.]]
    X = np.c_[rs.rand(14) * 4, rs.rand(14) * 4 - 1, rs.rand(14) * 2]
    Y = np.sin(X[:, 0]) + np.cos(X[:, 1]) - (X[:, 2] - 1) ** 2
    theta = [1.0, 0.8, 0.6]
    prior = oprior = None
    if with_prior:
        means = rs.rand(4, 3); beta = rs.randn(4) * 0.3
        lowerb = np.array([b[0] for b in bounds]); width = np.array([b[1] - b[0] for b in bounds])
        prior = RBFNMeanPrior(means=means, beta=beta, theta=10., lowerb=lowerb, width=width)
        oprior = orc.PriorSpec(means, beta, 10., lowerb, width)
    gp = _gp(GaussianKernel_ard(theta), X, Y, prior=prior) if with_prior else _gp(GaussianKernel_ard(theta), X, Y)
    # the oracle's DIRECT is pure Python: both sides run 12 iterations instead of maximizeEI's default 50
    monkeypatch.setattr(acq, "maximizeEI", functools.partial(acq.maximizeEI, maxiter=12))
    gal = acq.fastUCBGallery(gp, bounds, 4, samples=100, seed=5)
    ref = orc.fast_ucb_gallery(orc.KernelSpec(orc.K_SE_ARD, theta, 3), X, Y, bounds, 4, prior=oprior, samples=100, seed=5,
                               maxiter=12)
    assert len(gal) == len(ref) == 4
    for a, b in zip(gal, ref):
        assert np.allclose(a, b, rtol=0, atol=1e-9), (gal, ref)


def test_acqmax_many_equals_one_query_at_a_time():
    """independent queries on their own model handles, one host thread each: same points, sample counts and values as running
    them one after the other"""
    from ibo_b200 import _lib
    rs = np.random.RandomState(3)
    d = 5
    models, ymax = [], []
    for q in range(4):
        N = 150 + 130 * q
        X = rs.rand(N, d); Y = np.sin(3 * X).sum(axis=1)
        models.append(_lib.Model(_lib.KERNEL_SE_ARD, [0.4 + 0.05 * q] * d, X, Y, 0.1))
        ymax.append(float(Y.max()))
    lb, ub = np.zeros(d), np.ones(d)
    parm = [0.01, 0.02, 0.0, 0.05]
    seq = [m.acqmax(lb, ub, _lib.ACQ_EI, ymax[q], parm[q], _lib.FLAG_MODE_CPP, 40, 10 ** 6, 10 ** 6) for q, m in enumerate(models)]
    opt, optx, ns, it = _lib.acqmax_many(models, lb, ub, _lib.ACQ_EI, ymax, parm, _lib.FLAG_MODE_CPP, 40, 10 ** 6, 10 ** 6)
    for q in range(4):
        assert opt[q] == seq[q][0] and np.array_equal(optx[q], seq[q][1]) and ns[q] == seq[q][2] and it[q] == seq[q][3]
    with pytest.raises(_lib.IBOError):
        _lib.acqmax_many([models[0], models[0]], lb, ub, _lib.ACQ_EI, ymax[:2], parm[:2])
    for m in models:
        m.close()
