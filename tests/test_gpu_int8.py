"""GPU tests of the INT8 tensor-core path of wide batches (ibo_b200/csrc/score_i8.cuh) -- the DEFAULT for more than 2048
candidates: sigma^2 through 7 x 7 base-256 digits of W and K* multiplied exactly on tcgen05.mma kind::i8 (INT32 accumulators in
TMEM, W digits 1..4 fed through TMEM, FP64 assembly) and mu as k* . alpha must stay inside the same 1e-10 parity bound as the
FP64 DMMA path, against the oracle (reference arithmetic: ego/gaussianprocess/__init__.py:169-228, cpp/optimizeGP.cpp:57-215)
and against the DMMA path; candidates whose sigma^2 is too small for the integer scheme are re-scored by the DMMA kernels."""
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
                                              (129, 1, 2049, "se", False), (2048, 6, 20000, "se", False)])
@pytest.mark.parametrize("mode", ["py", "cpp"])
def test_default_wide_path_matches_oracle(N, d, M, kind, prior, mode):
    """no flag: a batch of more than 2048 candidates is scored by the INT8 path"""
    from ibo_b200 import _lib
    gp, o, Xs, Y = _case(N, d, M, kind, prior)
    fl = _lib.FLAG_MODE_PY if mode == "py" else _lib.FLAG_MODE_CPP
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
    # ... and it is a different arithmetic from the DMMA kernels (the flag really switches paths), equal to rounding
    f = gp.model.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, flags=fl | _lib.FLAG_FP64, want_posterior=True)
    assert not np.array_equal(f[2], s2)
    assert _rel(s2, f[2], 1e-300) <= 1e-12 and _rel(mu, f[1], 1e-3) <= 1e-12 and bidx == f[4]


def test_int8_path_matches_dmma_path_at_the_headline_shape():
    from ibo_b200 import _lib
    gp, o, Xs, Y = _case(2048, 6, 40000)
    a = gp.model.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, flags=_lib.FLAG_MODE_CPP | _lib.FLAG_FP64, want_posterior=True)
    b = gp.model.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, flags=_lib.FLAG_MODE_CPP, want_posterior=True)
    # sigma^2: both paths are within a few 1e-15 of the exact sum v^2 (tests/test_int8_model.py); EI near its 1e-5 floor carries
    # the absolute rounding noise of Phi = (1 + erf) / 2 times |mu - ymax|
    assert _rel(b[1], a[1], 1e-3) <= 1e-12 and _rel(b[2], a[2], 1e-300) <= 1e-12 and _rel(b[0], a[0], 1e-5) <= 5e-11
    assert a[4] == b[4]
    assert gp.model.last_guarded() == 0
    # PI and UCB ride on the same (mu, sigma^2)
    for acq, parm in ((_lib.ACQ_PI, 0.05), (_lib.ACQ_UCB, 1.7)):
        a = gp.model.score(Xs, acq, Y.max(), parm, flags=_lib.FLAG_MODE_CPP | _lib.FLAG_FP64)
        b = gp.model.score(Xs, acq, Y.max(), parm, flags=_lib.FLAG_MODE_CPP)
        assert _rel(b[0], a[0], 1e-5) <= 5e-11 and a[4] == b[4]


def test_int8_values_do_not_depend_on_the_batch():
    """a candidate's value is a function of (model, x) only: same bits wherever it sits in whatever wide batch"""
    from ibo_b200 import _lib
    gp, o, Xs, Y = _case(700, 4, 9000)
    fl = _lib.FLAG_MODE_CPP
    full = gp.model.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, flags=fl, want_posterior=True)
    perm = np.random.RandomState(5).permutation(len(Xs))[:4000]
    sub = gp.model.score(np.ascontiguousarray(Xs[perm]), _lib.ACQ_EI, Y.max(), 0.01, flags=fl, want_posterior=True)
    assert np.array_equal(sub[0], full[0][perm]) and np.array_equal(sub[1], full[1][perm]) and np.array_equal(sub[2], full[2][perm])
    # exact duplicates: lowest index wins the argmax
    dup = np.vstack([Xs[:3000], Xs[:3000]])
    r = gp.model.score(dup, _lib.ACQ_EI, Y.max(), 0.01, flags=fl)
    assert r[4] == int(np.argmax(full[0][:3000]))


def test_int8_follows_append_and_batch_size_classes():
    from ibo_b200 import _lib
    gp, o, Xs, Y = _case(250, 3, 3000)
    fl = _lib.FLAG_MODE_PY
    gp.model.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, flags=fl)
    rs = np.random.RandomState(9)
    Xn = rs.rand(10, 3); Yn = np.sin(3 * Xn).sum(axis=1)
    gp.addData(Xn, Yn)                       # rank-1 appends on the resident model: the digits of W are rebuilt
    o.add_data(Xn, Yn)
    sc, mu, s2, best, bidx = gp.model.score(Xs, _lib.ACQ_EI, max(gp.Y), 0.01, flags=fl, want_posterior=True)
    mu_o, s2_o = o.posterior_batch(Xs)
    assert _rel(mu, mu_o, 1e-3) <= TOL and _rel(s2, s2_o, 1e-300) <= TOL
    assert _lib.get_option("i8_min_batch") == -1
    f = gp.model.score(Xs[:500], _lib.ACQ_EI, max(gp.Y), 0.01, flags=fl | _lib.FLAG_FP64, want_posterior=True)
    # default rule: on a model this small (N = 260) a 500-point batch is below the break-even and stays on the FP64 latency shapes
    g = gp.model.score(Xs[:500], _lib.ACQ_EI, max(gp.Y), 0.01, flags=fl, want_posterior=True)
    assert np.array_equal(g[2], f[2])
    _lib.set_option("i8_min_batch", 192)
    try:
        # batches below a fixed i8_min_batch take the FP64 latency shapes whatever the flag says ...
        a = gp.model.score(Xs[:100], _lib.ACQ_EI, max(gp.Y), 0.01, flags=fl | _lib.FLAG_FP64, want_posterior=True)
        b = gp.model.score(Xs[:100], _lib.ACQ_EI, max(gp.Y), 0.01, flags=fl | _lib.FLAG_INT8, want_posterior=True)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
        # ... mid-size batches (192 .. 2048 points) take the INT8 kernels, and a candidate gets the bits it gets in a wide batch
        c = gp.model.score(Xs[:500], _lib.ACQ_EI, max(gp.Y), 0.01, flags=fl, want_posterior=True)
        assert np.array_equal(c[0], sc[:500]) and np.array_equal(c[1], mu[:500]) and np.array_equal(c[2], s2[:500])
        assert not np.array_equal(f[2], c[2]) and _rel(c[2], f[2], 1e-300) <= 1e-12
        _lib.set_option("i8_min_batch", 0)
        g = gp.model.score(Xs[:500], _lib.ACQ_EI, max(gp.Y), 0.01, flags=fl, want_posterior=True)
        assert np.array_equal(g[2], f[2])
    finally:
        _lib.set_option("i8_min_batch", -1)


def test_mid_size_batches_cross_over_to_int8_by_model_size():
    """default i8_min_batch rule: at N = 4096 a 256-point batch is past the break-even (133) and gets the bits of a wide batch,
    a 64-point batch is not and gets the FP64 bits"""
    from ibo_b200 import _lib
    gp, o, Xs, Y = _case(4096, 5, 3000)
    fl = _lib.FLAG_MODE_CPP
    wide = gp.model.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, flags=fl, want_posterior=True)
    mid = gp.model.score(Xs[:256], _lib.ACQ_EI, Y.max(), 0.01, flags=fl, want_posterior=True)
    assert np.array_equal(mid[0], wide[0][:256]) and np.array_equal(mid[2], wide[2][:256])
    small = gp.model.score(Xs[:64], _lib.ACQ_EI, Y.max(), 0.01, flags=fl, want_posterior=True)
    small64 = gp.model.score(Xs[:64], _lib.ACQ_EI, Y.max(), 0.01, flags=fl | _lib.FLAG_FP64, want_posterior=True)
    assert np.array_equal(small[2], small64[2]) and not np.array_equal(small[2], wide[2][:64])
    mu_o, s2_o = o.posterior_batch(Xs[:256])
    assert _rel(mid[1], mu_o, 1e-3) <= TOL and _rel(mid[2], s2_o, 1e-300) <= TOL


def test_option_int8_off_means_dmma_everywhere():
    from ibo_b200 import _lib
    gp, o, Xs, Y = _case(400, 3, 6000)
    ref = gp.model.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, flags=_lib.FLAG_MODE_CPP | _lib.FLAG_FP64, want_posterior=True)
    assert _lib.get_option("int8") == 1
    _lib.set_option("int8", 0)
    try:
        off = gp.model.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, flags=_lib.FLAG_MODE_CPP, want_posterior=True)
        forced = gp.model.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, flags=_lib.FLAG_MODE_CPP | _lib.FLAG_INT8, want_posterior=True)
    finally:
        _lib.set_option("int8", 1)
    on = gp.model.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, flags=_lib.FLAG_MODE_CPP, want_posterior=True)
    assert np.array_equal(off[0], ref[0]) and np.array_equal(off[2], ref[2])
    assert np.array_equal(forced[2], on[2]) and not np.array_equal(on[2], ref[2])
    with pytest.raises(_lib.IBOError):
        _lib.set_option("no_such_option", 1)


@pytest.mark.parametrize("noise", [1e-4, 1e-6])
def test_guard_rescoring_for_small_noise_models(noise):
    """sigma^2 >= noise for a model built from R, so with noise < 2^-10 candidates at / next to training points fall below the
    guard threshold: they must come back bit-identical to the DMMA path, the others as close to the oracle as the DMMA path is
    (a small-noise model is ill conditioned: cond(R) ~ N / noise bounds what ANY FP64 evaluation can deliver), with the same
    argmax (lowest index wins)."""
    from ibo_b200 import _lib
    from ibo_b200.gaussianprocess import GaussianProcess
    from ibo_b200.gaussianprocess.kernel import GaussianKernel_ard
    N, d, M = 384, 3, 6000
    rs = np.random.RandomState(77)
    X = rs.rand(N, d); Y = np.sin(3 * X).sum(axis=1)
    theta = [0.08, 0.1, 0.12]
    gp = GaussianProcess(GaussianKernel_ard(theta), X, Y, noise=noise)
    o = orc.GPOracle(orc.KernelSpec(orc.K_SE_ARD, theta, d), X, Y, noise)
    Xs = rs.rand(M, d)
    Xs[10:10 + 200] = X[:200]                               # exactly on training points
    Xs[1000:1200] = X[100:300] + 1e-5                       # right next to them
    Xs[5999] = X[7]
    fl = _lib.FLAG_MODE_CPP
    a = gp.model.score(Xs, _lib.ACQ_UCB, Y.max(), 1.5, flags=fl | _lib.FLAG_FP64, want_posterior=True)
    b = gp.model.score(Xs, _lib.ACQ_UCB, Y.max(), 1.5, flags=fl, want_posterior=True)
    ng = gp.model.last_guarded()
    assert 401 <= ng < M
    small = a[2] < 2.0 ** -10 * 0.999
    assert small.sum() >= 401
    assert np.array_equal(b[2][small], a[2][small]) and np.array_equal(b[1][small], a[1][small]) and np.array_equal(b[0][small], a[0][small])
    big = a[2] > 2.0 ** -10 * 1.001
    mu_o, sig = o.posterior_cpp(Xs)
    s2_o = sig ** 2
    e_fp64 = max(_rel(a[2][big], s2_o[big], 1e-300), _rel(a[1][big], mu_o[big], 1e-3))
    e_i8 = max(_rel(b[2][big], s2_o[big], 1e-300), _rel(b[1][big], mu_o[big], 1e-3))
    assert e_i8 <= max(TOL, 3 * e_fp64)
    assert max(_rel(b[2][big], a[2][big], 1e-300), _rel(b[1][big], a[1][big], 1e-3)) <= max(TOL, 3 * e_fp64)
    assert a[4] == b[4] and abs(a[3] - b[3]) <= 1e-12 * abs(a[3])
    # every candidate guarded
    near = np.ascontiguousarray(np.vstack([X] * 7)[:2500] + 1e-7)
    pa = gp.model.score(near, _lib.ACQ_EI, Y.max(), 0.01, flags=fl | _lib.FLAG_FP64, want_posterior=True)
    pb = gp.model.score(near, _lib.ACQ_EI, Y.max(), 0.01, flags=fl, want_posterior=True)
    assert gp.model.last_guarded() == len(near)
    assert np.array_equal(pa[0], pb[0]) and np.array_equal(pa[2], pb[2]) and pa[4] == pb[4] and pa[3] == pb[3]
    # argmax only (no score / posterior arrays requested)
    qa = gp.model.score(Xs, _lib.ACQ_UCB, Y.max(), 1.5, flags=fl | _lib.FLAG_FP64, want_scores=False)
    qb = gp.model.score(Xs, _lib.ACQ_UCB, Y.max(), 1.5, flags=fl, want_scores=False)
    assert qa[4] == qb[4] == a[4] and qb[3] == b[3]
    _lib.set_option("i8_guard", 0)
    try:
        c = gp.model.score(Xs, _lib.ACQ_UCB, Y.max(), 1.5, flags=fl, want_posterior=True)
    finally:
        _lib.set_option("i8_guard", 1)
    assert gp.model.last_guarded() == 0 and not np.array_equal(c[2][small], a[2][small])
