"""GPU tests of the fused small-model kernel (ibo_b200/csrc/tiny.cu, N <= 128): against the oracle, against the general
K1 -> K2 -> K3 path (option tiny = 0), argmax rules, batch-shape independence, big candidate sets."""
import os

import numpy as np
import pytest

from oracle import ibo_oracle as orc

pytestmark = pytest.mark.gpu


def _general(fn):
    from ibo_b200 import _lib
    _lib.set_option("tiny", 0)
    try:
        return fn()
    finally:
        _lib.set_option("tiny", -1)


CASES = [
    ("se_ard_n50_d2", orc.K_SE_ARD, [3.4 / 15, 10.0 / 15], 2, 50, False),
    ("se_iso_n1", orc.K_SE_ISO, [0.3], 3, 1, False),
    ("se_ard_mag_n5", orc.K_SE_ARD, [0.4, 0.5, 0.6, 0.9], 3, 5, False),
    ("matern3_n127", orc.K_MATERN3, [0.6, 1.0], 4, 127, False),
    ("matern5_n128", orc.K_MATERN5, [0.7, 1.0], 2, 128, False),
    ("matern5_ard_n64_d10", orc.K_MATERN5_ARD, [0.5 + 0.05 * j for j in range(10)] + [1.0], 10, 64, False),
    ("se_ard_n100_d20_prior", orc.K_SE_ARD, [1.0] * 20, 20, 100, True),
    ("se_ard_n33_d40", orc.K_SE_ARD, [2.0] * 40, 40, 33, False),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("mode", ["py", "cpp"])
def test_tiny_path_matches_oracle_and_general_path(case, mode):
    from ibo_b200 import _lib
    name, kind, hyper, d, N, with_prior = case
    rs = np.random.RandomState(len(name) + N)
    X = rs.rand(N, d)
    Y = np.sin(3 * X).sum(axis=1) + 0.05 * rs.randn(N)
    prior = None
    pr = None
    if with_prior:
        pr = orc.PriorSpec(rs.rand(6, d), rs.randn(6), 1.7, np.zeros(d), np.ones(d))
        prior = pr
    m = _lib.Model(kind, hyper, X, Y, 0.1, prior=prior)
    o = orc.GPOracle(orc.KernelSpec(kind, hyper, d), X, Y, 0.1, prior=pr)
    flags = _lib.FLAG_MODE_PY if mode == "py" else _lib.FLAG_MODE_CPP
    for M in (1, 7, 8, 9, 200, 5000):
        Xs = rs.rand(M, d)
        Xs[0] = X[0]                                            # a candidate on top of a training point
        mu_o, s2_o = o.posterior_batch(Xs, floor=10e-8 if mode == "py" else 1e-8)
        for acq, parm in ((_lib.ACQ_EI, 0.01), (_lib.ACQ_PI, 0.01), (_lib.ACQ_UCB, 1.3)):
            sc, mu, s2, best, bidx = m.score(Xs, acq, Y.max(), parm, flags, want_posterior=True)
            ref = orc.score(acq, mode, mu_o, s2_o, Y.max(), parm)
            assert np.max(np.abs(mu - mu_o) / np.maximum(np.abs(mu_o), 1e-3)) <= 1e-10
            assert np.max(np.abs(s2 - s2_o) / s2_o) <= 1e-10
            assert np.max(np.abs(sc - ref) / np.maximum(np.abs(ref), 1e-5)) <= 1e-10
            assert bidx == int(np.argmax(sc)) and best == sc[bidx]
            g = _general(lambda: m.score(Xs, acq, Y.max(), parm, flags, want_posterior=True))
            # the two device paths agree to the same tolerance as each agrees with the oracle (1e-10 relative, FP64)
            assert np.max(np.abs(sc - g[0]) / np.maximum(np.abs(g[0]), 1e-5)) <= 1e-10
            assert np.max(np.abs(mu - g[1]) / np.maximum(np.abs(g[1]), 1e-3)) <= 1e-10
            assert np.max(np.abs(s2 - g[2]) / g[2]) <= 1e-10


def test_tiny_values_do_not_depend_on_the_batch():
    """batches of up to 4096 points take the fused kernel: bit-identical whatever the batch size, position or order; larger
    sets take the general path and agree with it to rounding"""
    from ibo_b200 import _lib
    rs = np.random.RandomState(3)
    X = rs.rand(50, 2); Y = np.cos(4 * X).sum(axis=1)
    m = _lib.Model(_lib.KERNEL_SE_ARD, [0.3, 0.4], X, Y, 0.1)
    Xs = rs.rand(300000, 2)
    big = m.score(Xs, _lib.ACQ_EI, Y.max(), 0.01, _lib.FLAG_MODE_CPP)[0]
    full = np.concatenate([m.score(Xs[i:i + 4096], _lib.ACQ_EI, Y.max(), 0.01, _lib.FLAG_MODE_CPP)[0] for i in range(0, 40960, 4096)])
    assert np.max(np.abs(full - big[:40960]) / np.maximum(np.abs(full), 1e-5)) <= 1e-10
    for lo, hi in [(0, 1), (5, 13), (1000, 1700), (40000, 40960), (7, 40), (12345, 13456)]:
        part = m.score(Xs[lo:hi], _lib.ACQ_EI, Y.max(), 0.01, _lib.FLAG_MODE_CPP)[0]
        assert np.array_equal(part, full[lo:hi])
    perm = rs.permutation(4000)
    assert np.array_equal(m.score(Xs[perm], _lib.ACQ_EI, Y.max(), 0.01, _lib.FLAG_MODE_CPP)[0], full[perm])
    # resident candidates + ties: exact duplicates, the lowest index wins
    Xd = Xs[:3000].copy(); Xd[2000] = Xd[17]; Xd[2999] = Xd[17]
    c = _lib.ResidentCandidates(m, Xd)
    out = np.empty(3000)
    best, bidx, ms = c.score(_lib.ACQ_UCB, Y.max(), 2.0, _lib.FLAG_MODE_CPP, scores_out=out)
    assert out[17] == out[2000] == out[2999] and bidx == int(np.argmax(out)) and best == out[bidx]
    dup = np.repeat(Xd[17:18], 700, axis=0)
    assert m.score(dup, _lib.ACQ_EI, Y.max(), 0.01, _lib.FLAG_MODE_CPP)[4] == 0
    c.close()


def test_tiny_direct_matches_general_path_and_reference_goldens_still_hold():
    """maximizeEI on a config-#1-sized model through both paths: same point, same sample count"""
    from ibo_b200.acquisition import cdirectGP, maximizeEI
    from ibo_b200.gaussianprocess import GaussianProcess
    from ibo_b200.gaussianprocess.kernel import GaussianKernel_ard
    b = [[-5., 10.], [0., 15.]]
    X = np.array(orc.lhc_sample(b, 50, seed=0)); Y = -orc.branin(X) / 100.0
    gp = GaussianProcess(GaussianKernel_ard([3.4, 10.0]), X, Y, noise=0.1)
    o1, x1 = maximizeEI(gp, b, xi=0.01, maxiter=50, maxtime=10 ** 6, maxsample=10000)
    n1 = cdirectGP.last["nsamples"]
    o2, x2 = _general(lambda: maximizeEI(gp, b, xi=0.01, maxiter=50, maxtime=10 ** 6, maxsample=10000))
    n2 = cdirectGP.last["nsamples"]
    assert n1 == n2 and np.allclose(x1, x2, rtol=0, atol=1e-12) and abs(o1 - o2) <= 1e-10 * abs(o1)


def test_batch_server_gives_the_bits_of_the_per_batch_launches():
    """DIRECT on a one-row-block model: the resident kernel fed through the mailbox (option tiny_server, default) and one launch
    per batch are the same arithmetic in the same order -- same optimum bit for bit, same samples -- for every acquisition
    function, with a prior mean, and when queries, score() calls and appends alternate on the same model."""
    from ibo_b200 import _lib
    rs = np.random.RandomState(11)
    d, N = 3, 90
    X = rs.rand(N, d); Y = np.sin(3 * X).sum(axis=1)
    pr = orc.PriorSpec(rs.rand(3, d), 0.2 * rs.randn(3), 3.0, np.zeros(d), np.ones(d))
    lb, ub = np.zeros(d), np.ones(d)
    for prior in (None, pr):
        m = _lib.Model(_lib.KERNEL_MATERN3, [0.6, 1.0], X, Y, 0.05, prior=prior)
        try:
            for acq, parm in ((_lib.ACQ_EI, 0.01), (_lib.ACQ_PI, 0.01), (_lib.ACQ_UCB, 1.3)):
                for fl in (_lib.FLAG_MODE_CPP, _lib.FLAG_MODE_PY):
                    a = m.acqmax(lb, ub, acq, float(Y.max()), parm, flags=fl, maxiter=40, maxsample=10 ** 6)
                    _lib.set_option("tiny_server", 0)
                    try:
                        b = m.acqmax(lb, ub, acq, float(Y.max()), parm, flags=fl, maxiter=40, maxsample=10 ** 6)
                    finally:
                        _lib.set_option("tiny_server", 1)
                    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and a[2] == b[2] and a[3] == b[3]
            # queries back to back with other work on the same stream in between
            Xs = rs.rand(37, d)
            ref = m.score(Xs, _lib.ACQ_EI, float(Y.max()), 0.01)[0]
            first = m.acqmax(lb, ub, _lib.ACQ_EI, float(Y.max()), 0.01, maxiter=25, maxsample=10 ** 6)
            for _ in range(20):
                again = m.acqmax(lb, ub, _lib.ACQ_EI, float(Y.max()), 0.01, maxiter=25, maxsample=10 ** 6)
                assert again[0] == first[0] and np.array_equal(again[1], first[1]) and again[2] == first[2]
                assert np.array_equal(m.score(Xs, _lib.ACQ_EI, float(Y.max()), 0.01)[0], ref)
            m.append(rs.rand(2, d), rs.rand(2))
            after = m.acqmax(lb, ub, _lib.ACQ_EI, float(Y.max()), 0.01, maxiter=25, maxsample=10 ** 6)
            _lib.set_option("tiny_server", 0)
            try:
                after0 = m.acqmax(lb, ub, _lib.ACQ_EI, float(Y.max()), 0.01, maxiter=25, maxsample=10 ** 6)
            finally:
                _lib.set_option("tiny_server", 1)
            assert after[0] == after0[0] and np.array_equal(after[1], after0[1]) and after[2] == after0[2]
        finally:
            m.close()


def test_batch_server_side_by_side_queries():
    """ibo_acqmax_many on small models: one resident kernel per model handle, all on one device (at most four at a time, the
    other queries launch per batch)"""
    from ibo_b200 import _lib
    rs = np.random.RandomState(12)
    d = 2
    models, want = [], []
    lb, ub = np.zeros(d), np.ones(d)
    try:
        for q in range(6):
            X = rs.rand(40 + 10 * q, d); Y = np.cos(4 * X).sum(axis=1)
            models.append(_lib.Model(_lib.KERNEL_SE_ARD, [0.3, 0.4], X, Y, 0.1))
            want.append(models[-1].acqmax(lb, ub, _lib.ACQ_EI, float(Y.max()), 0.01, maxiter=30, maxsample=10 ** 6) + (float(Y.max()),))
        opt, optx, ns, it = _lib.acqmax_many(models, lb, ub, _lib.ACQ_EI, [w[4] for w in want], 0.01, maxiter=30, maxsample=10 ** 6)
        for q in range(6):
            assert opt[q] == want[q][0] and np.array_equal(optx[q], want[q][1]) and ns[q] == want[q][2]
    finally:
        for m in models: m.close()
