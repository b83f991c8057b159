"""GPU tests of the device marginal likelihood (ibo_nlml / ibo_kernel_matrix) against the oracle's restatement
of trainhyper.marginalLikelihood and Kernel.derivative (ego/gaussianprocess/trainhyper.py:47-76, kernel.py:92-266) and
against the reference's known answers (ego/unittest_GP.py:160-266)."""
import numpy as np
import pytest

from oracle import ibo_oracle as orc

pytestmark = pytest.mark.gpu

HX = np.array([[.5, .1, .3], [.9, 1.2, .1], [.55, .234, .1], [.234, .547, .675]])
HY = np.array([.5, 1., .5, 2.])


def test_reference_known_answers_through_the_python_surface():
    from ibo_b200.gaussianprocess.kernel import GaussianKernel_ard, SVGaussianKernel_iso, MaternKernel3, MaternKernel5
    from ibo_b200.gaussianprocess.trainhyper import marginalLikelihood
    nl, g = marginalLikelihood(GaussianKernel_ard([2., 2., .1]), HX, HY, 3, noise=0.0)        # unittest_GP.py:204-208
    assert abs(nl - 5.8404) < 5e-5 and np.max(np.abs(g - [0.0039, 0.0302, -0.1733])) < 5e-5
    nl, g = marginalLikelihood(SVGaussianKernel_iso([1.5, 1.1]), HX, HY, 2, noise=0.0)        # :232-236
    assert abs(nl - 7.514) < 5e-4 and np.max(np.abs(g - [11.4659, -10.0714])) < 5e-5
    nl, g = marginalLikelihood(MaternKernel3([1.5, 1.1]), HX, HY, 2, noise=0.0)               # :262-266
    assert abs(nl - 5.1827) < 5e-5 and abs(g[0] - 1.6947897766) < 1e-9 and abs(g[1] + 2.9350) < 5e-5
    nl, g = marginalLikelihood(MaternKernel5([1.5, 1.1]), HX, HY, 2, noise=0.0)               # :268-272
    assert abs(nl - 5.6652) < 5e-5 and abs(g[0] - 4.4782) < 5e-5 and abs(g[1] + 4.8737) < 5e-5
    # useCholesky=False gives the same numbers (:211-214)
    nl2, g2 = marginalLikelihood(GaussianKernel_ard([2., 2., .1]), HX, HY, 3, useCholesky=False, noise=0.0)
    assert abs(nl2 - 5.8404) < 5e-5 and np.max(np.abs(g2 - [0.0039, 0.0302, -0.1733])) < 5e-5
    with pytest.raises(ValueError):
        marginalLikelihood(GaussianKernel_ard([2., 2., .1]), HX, HY, 4, noise=0.0)


def test_derivative_matrices_match_reference_targets():
    from ibo_b200.gaussianprocess.kernel import GaussianKernel_ard, SVGaussianKernel_iso
    k = GaussianKernel_ard([2., 2., .1])
    t2 = np.array([[0, .4561, .54, .012], [.4561, 0, 0, 0], [.54, .0, 0, 0], [.012, 0, 0, 0]])
    assert np.max(np.abs(k.derivative(HX, 2) - t2)) < 1e-4                                    # unittest_GP.py:181
    with pytest.raises(ValueError):
        k.derivative(HX, 3)
    ks = SVGaussianKernel_iso([1.5, 1.1])
    t0 = np.array([[0, .5543, .0321, .2018], [.5543, 0, .449, .4945], [.0321, .449, 0, .2527], [.2018, .4945, .2527, 0]])
    t1 = np.array([[2.42, 1.769, 2.3877, 2.2087], [1.769, 2.42, 1.914, 1.8533], [2.3877, 1.914, 2.42, 2.1519],
                   [2.2087, 1.8533, 2.1519, 2.42]])
    assert np.max(np.abs(ks.derivative(HX, 0) - t0)) < 1e-4                                   # :239-246
    assert np.max(np.abs(ks.derivative(HX, 1) - t1)) < 1e-4
    assert np.max(np.abs(ks.covMatrix(HX) - t1 / 2)) < 1e-4


CASES = [
    (orc.K_SE_ARD, [.4, .6, .9], 3, 300),
    (orc.K_SE_ARD, [.4, .6, .9, .5, .7, 1.3], 5, 257),       # with magnitude
    (orc.K_SE_ISO, [.7], 4, 129),
    (orc.K_MATERN3, [.8, 1.2], 3, 200),
    (orc.K_MATERN5, [.8, 1.2], 3, 128),
    (orc.K_MATERN5_ARD, [.5 + .05 * j for j in range(10)] + [1.1], 10, 400),
    (orc.K_SE_ARD, [1.0] * 20 + [0.9], 20, 333),             # three passes of the gradient reduction (21 hyperparameters)
]


@pytest.mark.parametrize("kind,hyper,d,N", CASES)
@pytest.mark.parametrize("noise", [0.1, 1e-3])
def test_nlml_and_gradient_match_oracle(kind, hyper, d, N, noise):
    from ibo_b200 import _lib
    rs = np.random.RandomState(N + d)
    X = rs.rand(N, d)
    Y = np.sin(2 * X).sum(axis=1) + 0.1 * rs.randn(N)
    spec = orc.KernelSpec(kind, hyper, d)
    nl_o, g_o = orc.marginal_likelihood(spec, X, Y, len(hyper), noise=noise)
    nl, g = _lib.nlml(kind, hyper, X, Y, noise)
    assert abs(nl - nl_o) <= 1e-10 * abs(nl_o)                # tolerance: 1e-10 relative on the value
    # gradient: a sum of N^2 signed terms through inv(K); tolerance relative to the sum of their magnitudes, scaled by
    # the conditioning of K (both sides lose cond(K) * eps in inv(K))
    K = orc.cov_matrix(spec, X) + noise * np.eye(N)
    Ki = np.linalg.inv(K)
    al = Ki.dot(Y)
    G = Ki - np.outer(al, al)
    for h in range(len(hyper)):
        scale = 0.5 * np.sum(np.abs(G * orc.kernel_derivative(spec, X, h)))
        tol = (1e-10 if noise >= 0.1 else 1e-8) * scale
        assert abs(g[h] - g_o[h]) <= tol, (h, g[h], g_o[h], scale)
    # value-only call gives the identical value
    assert _lib.nlml(kind, hyper, X, Y, noise, want_grad=False)[0] == nl


@pytest.mark.parametrize("kind,hyper,d,N", CASES[:6])
def test_kernel_matrix_and_derivatives_match_oracle(kind, hyper, d, N):
    from ibo_b200 import _lib
    rs = np.random.RandomState(3)
    X = rs.rand(min(N, 150), d)
    spec = orc.KernelSpec(kind, hyper, d)
    assert np.max(np.abs(_lib.kernel_matrix(kind, hyper, X) - orc.cov_matrix(spec, X))) < 1e-14
    for h in range(len(hyper)):
        D = _lib.kernel_matrix(kind, hyper, X, h)
        assert np.max(np.abs(D - orc.kernel_derivative(spec, X, h))) < 1e-13
    if kind == orc.K_MATERN3:
        D = _lib.kernel_matrix(kind, hyper, X, 0, flags=_lib.FLAG_GRAD_EXACT)
        assert np.max(np.abs(D - orc.kernel_derivative(spec, X, 0, exact_matern3=True))) < 1e-13


def test_gradient_is_the_derivative_of_the_value():
    """finite differences of the device nlml in log hyperparameters (kinds whose reference derivative is exact)"""
    from ibo_b200 import _lib
    rs = np.random.RandomState(5)
    X = rs.rand(200, 4)
    Y = np.cos(3 * X).sum(axis=1)
    for kind, hyper, flags in ((orc.K_SE_ARD, [.5, .6, .7, .8, 1.2], 0), (orc.K_MATERN5_ARD, [.5, .6, .7, .8, 1.2], 0),
                               (orc.K_MATERN3, [.8, 1.2], _lib.FLAG_GRAD_EXACT)):
        _, g = _lib.nlml(kind, hyper, X, Y, 0.05, flags=flags)
        for h in range(len(hyper)):
            e = np.zeros(len(hyper)); e[h] = 1e-5
            fp = _lib.nlml(kind, np.exp(np.log(hyper) + e), X, Y, 0.05, want_grad=False)[0]
            fm = _lib.nlml(kind, np.exp(np.log(hyper) - e), X, Y, 0.05, want_grad=False)[0]
            assert abs((fp - fm) / 2e-5 - g[h]) < 1e-5 * max(1.0, abs(g[h]))


def test_bfgs_over_log_hyperparameters_reaches_the_reference_answers():
    """the reference's usage (unittest_GP.py:215-217,252-254) with our nlml / dnlml; marginalLikelihood's default noise is
    1e-3 there, the known answers correspond to noise -> 0, hence 2 decimals as in the reference's own assertion"""
    from functools import partial
    from scipy import optimize
    from ibo_b200.gaussianprocess.kernel import SVGaussianKernel_ard, SVGaussianKernel_iso
    from ibo_b200.gaussianprocess.trainhyper import nlml, dnlml
    r = optimize.fmin_bfgs(nlml, np.log([2., 2., .1, 1.]), dnlml, args=(SVGaussianKernel_ard, HX, HY), disp=False)
    assert np.max(np.abs(r[1:] - [0.95405, -0.9769, 0.36469])) < 5e-3
    r = optimize.fmin_bfgs(nlml, np.log([1.5, 1.1]), dnlml, args=(SVGaussianKernel_iso, HX, HY), disp=False)
    assert np.max(np.abs(r - [-0.0893, 0.29])) < 5e-3


def test_not_positive_definite_is_reported():
    from ibo_b200 import _lib
    from ibo_b200.gaussianprocess.kernel import GaussianKernel_ard
    from ibo_b200.gaussianprocess.trainhyper import nlml
    X = np.array([[0.1, 0.2], [0.1, 0.2], [0.5, 0.5]])          # duplicate point, no noise -> singular K
    with pytest.raises(np.linalg.LinAlgError):
        _lib.nlml(_lib.KERNEL_SE_ARD, [1., 1.], X, [1., 2., 3.], 0.0)

    class K0(GaussianKernel_ard):
        pass
    import ibo_b200.gaussianprocess.trainhyper as th
    orig = th.marginalLikelihood
    try:
        th.marginalLikelihood = lambda k, X_, Y_, n, computeGradient=False: orig(k, X_, Y_, n, computeGradient=computeGradient, noise=0.0)
        assert nlml(np.log([1., 1.]), K0, X, [1., 2., 3.]) == 100                   # trainhyper.py:111-114
    finally:
        th.marginalLikelihood = orig
