"""CPU tests: pin the oracle restatement (oracle/ibo_oracle.py) against
 (a) golden vectors produced by the reference's own C++ (tests/golden/, made by make_golden.py),
 (b) the reference library itself when oracle/_ref is present (live, random inputs),
 (c) the known answers of the reference's unit tests (ego/unittest_IBO.py:107-134)."""
import ctypes
import json
import os
from ctypes import POINTER, c_double, c_int, c_long

import numpy as np
import pytest

from oracle import ibo_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "ref_candidates.npz"))
with open(os.path.join(HERE, "golden", "ref_direct.json")) as fh:
    GOLD_DIRECT = json.load(fh)
SCEN = sorted(set(k.split("/")[0] for k in GOLD.files))


def gold_model(name):
    g = lambda k: GOLD["%s/%s" % (name, k)]
    d = g("X").shape[1]
    kern = orc.KernelSpec(int(g("kind")), g("hyper"), d)
    prior = None
    if "%s/p_means" % name in GOLD.files:
        prior = orc.PriorSpec(g("p_means"), g("p_beta"), float(g("p_theta")), g("p_lowerb"), g("p_width"))
    return orc.GPOracle(kern, g("X"), g("Y"), float(g("noise")), prior=prior), g


def rel(a, b, floor):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))


@pytest.mark.parametrize("name", SCEN)
def test_cpp_mode_matches_reference_values(name):
    """explicit-inverse restatement (same invR the reference was given) -> near machine precision"""
    gp, g = gold_model(name)
    mu, sig = gp.posterior_cpp(g("Xs"), invR=g("invR"))
    assert rel(mu, g("mu"), 1e-3) < 1e-12
    assert rel(sig, g("sigma"), 1e-12) < 1e-12
    ymax = g("Y").max()
    assert rel(-orc.ei_cpp(mu, sig, ymax, 0.01), g("negei"), 1e-6) < 1e-10
    assert rel(-orc.ei_cpp(mu, sig, ymax, 0.1), g("negei_xi1"), 1e-6) < 1e-10
    assert rel(-orc.pi_cpp(mu, sig, ymax, 0.01), g("negpi"), 1e-6) < 1e-10
    assert rel(-orc.ucb_cpp(mu, sig, 1.3), g("negucb"), 1e-6) < 1e-12


@pytest.mark.parametrize("name", SCEN)
def test_cholesky_path_matches_reference_values(name):
    """the NumPy (Cholesky) formulation agrees with the reference's explicit-inverse numbers to 1e-10"""
    gp, g = gold_model(name)
    mu, s2 = gp.posterior_batch(g("Xs"), floor=1e-8)
    assert rel(mu, g("mu"), 1e-3) < 1e-10
    assert rel(np.sqrt(s2), g("sigma"), 1e-12) < 1e-10
    ei = orc.score(orc.ACQ_EI, "cpp", mu, s2, g("Y").max(), 0.01)
    assert rel(-ei, g("negei"), 1e-5) < 1e-10
    assert int(np.argmax(ei)) == int(np.argmin(g("negei")))


@pytest.mark.parametrize("name", ["shekel_iso3", "branin_matern3b", "prior_iso"])
def test_scalar_faithful_equals_vectorised(name):
    gp, g = gold_model(name)
    mu, s2 = gp.posterior_batch(g("Xs")[:12])
    for i in range(12):
        m1, v1 = gp.posterior_scalar(g("Xs")[i])
        assert abs(m1 - mu[i]) <= 1e-12 * max(1.0, abs(mu[i]))
        assert abs(v1 - s2[i]) <= 1e-12 * max(1.0, abs(s2[i]))


def test_python_vs_cpp_ei_gap_is_the_documented_erf_gap():
    """Chebyshev erf + truncated constants differ from libm by < 5e-7 absolute."""
    gp, g = gold_model("shekel_iso3")
    mu, s2 = gp.posterior_batch(g("Xs"))
    a = orc.score(orc.ACQ_EI, "py", mu, s2, g("Y").max(), 0.01)
    b = orc.score(orc.ACQ_EI, "cpp", mu, np.maximum(s2, 1e-8), g("Y").max(), 0.01)
    assert np.max(np.abs(a - b)) < 5e-7
    assert np.argmax(a) == np.argmax(b)


def test_erf_py_accuracy_and_sign():
    z = np.linspace(-4, 4, 401)
    import math
    ref = np.array([math.erf(v) for v in z])
    assert np.max(np.abs(orc.erf_py(z) - ref)) < 1.3e-7            # gaussianprocess/__init__.py:52
    nz = z[np.abs(z) > 1e-12]                                      # erf(0) is +1e-9, not 0 (:50-51)
    assert np.array_equal(orc.erf_py(-nz), -orc.erf_py(nz))


def test_training_properties():
    """ego/unittest_GP.py:77-106 style bounds: sigma^2 small and mu near y at training points."""
    gp, g = gold_model("branin_ard50")
    mu, s2 = gp.posterior_batch(g("X"))
    assert np.all(s2 < 1.0 / (1.0 + gp.noise) + gp.noise)
    assert np.all(np.abs(mu - g("Y")) < 2 * gp.noise + 0.25)


@pytest.mark.parametrize("tag", sorted(GOLD_DIRECT["direct"].keys()))
def test_direct_restatement_follows_reference_trajectory(tag):
    rec = GOLD_DIRECT["direct"][tag]
    b = np.array(rec["bounds"])
    fns = {"shekel5": orc.shekel5, "branin": lambda x: float(orc.branin(x)),
           "quad": lambda x: float(np.sum((x - 0.3) ** 2)), "sin6": lambda x: float(np.sum(np.sin(3 * x) + (x - .4) ** 2))}
    f = fns[tag.split("_")[0]]
    trace = []
    fmin, xmin, ns = orc.direct_cpp(f, b[:, 0], b[:, 1], rec["maxiter"], rec["maxsample"], record=trace)
    assert ns == rec["nsamples"]
    assert fmin == rec["fmin"]
    assert np.array_equal(xmin, np.array(rec["xmin"]))
    tr = np.array(trace)
    assert np.array_equal(tr[:40], np.array(rec["trace_head"]))
    assert np.array_equal(tr[-10:], np.array(rec["trace_tail"]))


def test_direct_known_answers():
    """ego/unittest_IBO.py:107-114: Shekel5, 20 iterations -> -10.1532 at (4,4,4,4) to 3 d.p."""
    rec = GOLD_DIRECT["direct"]["shekel5_it20"]
    assert round(rec["fmin"], 3) == -10.153
    assert all(round(v, 2) == 4.0 for v in rec["xmin"])
    fmin, xmin, ns = orc.direct_cpp(orc.shekel5, [0.] * 4, [10.] * 4, 20, 200000)
    assert abs(fmin + 10.1532) < 1e-3 and np.all(np.abs(xmin - 4.0) < 5e-3)


@pytest.mark.parametrize("key", sorted(GOLD_DIRECT["acqmaxGP"].keys()))
def test_oracle_acqmax_matches_reference_acqmaxGP(key):
    """DIRECT restatement over the oracle's C++-mode acquisition reproduces libego's acqmaxGP."""
    name, acq, parm, it, ms = key.split("|")
    acq, parm, it, ms = int(acq[3:]), float(parm[4:]), int(it[2:]), int(ms[2:])
    if name == "hartman_ard200":
        pytest.skip("pure-Python DIRECT too slow for N=200 x 3000 samples; covered by the GPU parity test")
    gp, g = gold_model(name)
    invR = g("invR")
    ymax = g("Y").max()

    def f(x):
        mu, sig = gp.posterior_cpp(x[None, :], invR=invR)
        return -float(orc.score(acq, "cpp", mu, sig ** 2, ymax, parm)[0])
    b = g("bounds")
    fmin, xmin, ns = orc.direct_cpp(f, b[:, 0], b[:, 1], it, ms)
    rec = GOLD_DIRECT["acqmaxGP"][key]
    assert abs(fmin - rec["fmin"]) <= 1e-9 * max(1e-3, abs(rec["fmin"]))
    assert np.allclose(xmin, rec["xmin"], rtol=0, atol=1e-12)


REF_HARNESS = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libego_harness.so")


@pytest.mark.skipif(not os.path.exists(REF_HARNESS), reason="oracle/_ref not built")
def test_live_reference_random_model():
    """The reference's own negei over a fresh random SE-ARD model (N=300, d=5) vs the oracle."""
    pd = POINTER(c_double)
    har = ctypes.CDLL(REF_HARNESS)
    har.ref_set_model.argtypes = [c_int, pd, pd, pd, c_int, c_int, pd, c_int, pd, pd, c_double, pd, pd, c_double, c_double]
    har.ref_eval.argtypes = [c_int, c_long, pd, pd, pd, pd, c_int]
    rs = np.random.RandomState(5)
    N, d, M = 300, 5, 200
    X = np.ascontiguousarray(rs.rand(N, d)); Y = np.ascontiguousarray(np.sin(3 * X).sum(axis=1))
    hyper = np.ascontiguousarray(np.array([.4, .5, .6, .7, .8]))
    gp = orc.GPOracle(orc.KernelSpec(0, hyper, d), X, Y, 0.1)
    invR = np.ascontiguousarray(gp.invR())
    Xs = np.ascontiguousarray(rs.rand(M, d))
    z = np.zeros(1)
    dp = lambda a: a.ctypes.data_as(pd)
    har.ref_set_model(d, dp(invR), dp(X), dp(Y), N, 0, dp(hyper), 0, dp(z), dp(z), 0.0, dp(z), dp(z), 0.01, 0.1)
    v = np.empty(M); mu = np.empty(M); sg = np.empty(M)
    har.ref_eval(0, M, dp(Xs), dp(v), dp(mu), dp(sg), 2)
    mu_o, s2_o = gp.posterior_batch(Xs, floor=1e-8)
    assert rel(mu_o, mu, 1e-3) < 1e-10
    assert rel(np.sqrt(s2_o), sg, 1e-12) < 1e-10
    ei = orc.score(orc.ACQ_EI, "cpp", mu_o, s2_o, Y.max(), 0.01)
    assert rel(-ei, v, 1e-5) < 1e-10
    assert np.argmax(ei) == np.argmin(v)


def test_oracle_block_append_equals_batch_build():
    """ego/gaussianprocess/__init__.py:300-308 restated in GPOracle.add_data: same R (exactly) and same L (rounding)
    as building from all the data at once -- the property ego/unittest_GP.py:109-156 checks"""
    rs = np.random.RandomState(0)
    X = rs.rand(30, 3); Y = rs.randn(30)
    spec = orc.KernelSpec(orc.K_SE_ARD, [.3, .4, .5], 3)
    o = orc.GPOracle(spec, X[:20], Y[:20], 0.1)
    o.add_data(X[20:25], Y[20:25]); o.add_data(X[25], Y[25]); o.add_data(X[26:], Y[26:])
    f = orc.GPOracle(spec, X, Y, 0.1)
    assert np.array_equal(o.R, f.R) and np.max(np.abs(o.L - f.L)) < 1e-14
    q = rs.rand(5, 3)
    assert np.allclose(o.posterior_batch(q)[0], f.posterior_batch(q)[0], rtol=1e-12)


# ---------------------------------------------------------------------------------------------
# hyper-parameter learning: the reference's known answers, ego/unittest_GP.py:160-266
# (values "collected from Carl Rasmussen's MATLAB code"; they correspond to noise = 0)
# ---------------------------------------------------------------------------------------------
HX = np.array([[.5, .1, .3], [.9, 1.2, .1], [.55, .234, .1], [.234, .547, .675]])
HY = np.array([.5, 1., .5, 2.])


def test_marginal_likelihood_known_answers_ard():
    k = orc.KernelSpec(orc.K_SE_ARD, [2., 2., .1], 3)
    t0 = np.array([[0, .0046, .0001, 0], [.0046, 0, .0268, 0], [.0001, .0268, 0, 0], [0, 0, 0, 0]])
    t1 = np.array([[0, .0345, .0006, 0], [.0345, 0, .2044, 0], [.0006, .2044, 0, 0], [0, 0, 0, 0]])
    t2 = np.array([[0, .4561, .54, .012], [.4561, 0, 0, 0], [.54, .0, 0, 0], [.012, 0, 0, 0]])
    for hp, t in enumerate((t0, t1, t2)):                              # unittest_GP.py:179-201, epsilon 1e-4
        assert np.max(np.abs(orc.kernel_derivative(k, HX, hp) - t)) < 1e-4
    nl, g = orc.marginal_likelihood(k, HX, HY, 3, noise=0.0)          # unittest_GP.py:204-208
    assert abs(nl - 5.8404) < 5e-5
    assert np.max(np.abs(g - [0.0039, 0.0302, -0.1733])) < 5e-5
    with pytest.raises(ValueError):                                    # kernel.py:163-166
        orc.kernel_derivative(k, HX, 3)


def test_marginal_likelihood_known_answers_sv_iso():
    # SVGaussianKernel_iso([1.5, 1.1]) == SE-ARD with equal length scales + magnitude; d/dlog theta = sum over dims
    k = orc.KernelSpec(orc.K_SE_ARD, [1.5, 1.5, 1.5, 1.1], 3)
    nl, g = orc.marginal_likelihood(k, HX, HY, 4, noise=0.0)          # unittest_GP.py:232-236
    assert abs(nl - 7.514) < 5e-4
    assert abs(g[:3].sum() - 11.4659) < 5e-5 and abs(g[3] + 10.0714) < 5e-5
    t0 = np.array([[0, .5543, .0321, .2018], [.5543, 0, .449, .4945], [.0321, .449, 0, .2527], [.2018, .4945, .2527, 0]])
    t1 = np.array([[2.42, 1.769, 2.3877, 2.2087], [1.769, 2.42, 1.914, 1.8533], [2.3877, 1.914, 2.42, 2.1519],
                   [2.2087, 1.8533, 2.1519, 2.42]])
    d0 = sum(orc.kernel_derivative(k, HX, h) for h in range(3))
    assert np.max(np.abs(d0 - t0)) < 1e-4                              # unittest_GP.py:239-246
    assert np.max(np.abs(orc.kernel_derivative(k, HX, 3) - t1)) < 1e-4


def test_marginal_likelihood_known_answers_matern():
    k3 = orc.KernelSpec(orc.K_MATERN3, [1.5, 1.1], 3)
    nl, g = orc.marginal_likelihood(k3, HX, HY, 2, noise=0.0)         # unittest_GP.py:262-266
    assert abs(nl - 5.1827) < 5e-5
    assert abs(g[0] - 1.6947897766) < 1e-9                             # pins the reference's unscaled-distance expression
    assert abs(g[1] + 2.9350) < 5e-5
    k5 = orc.KernelSpec(orc.K_MATERN5, [1.5, 1.1], 3)
    nl, g = orc.marginal_likelihood(k5, HX, HY, 2, noise=0.0)         # unittest_GP.py:268-272 (commented out there)
    assert abs(nl - 5.6652) < 5e-5 and abs(g[0] - 4.4782) < 5e-5 and abs(g[1] + 4.8737) < 5e-5


def test_marginal_likelihood_bfgs_known_answers():
    """unittest_GP.py:215-217,252-254: BFGS over the log hyperparameters (the first ARD length scale is a flat direction
    -- the data do not vary along it -- so only the other three coordinates are pinned)"""
    from scipy import optimize
    f = lambda lh: orc.marginal_likelihood(orc.KernelSpec(orc.K_SE_ARD, np.exp(lh), 3), HX, HY, 4, noise=0.0, compute_gradient=False)
    g = lambda lh: orc.marginal_likelihood(orc.KernelSpec(orc.K_SE_ARD, np.exp(lh), 3), HX, HY, 4, noise=0.0)[1]
    r = optimize.fmin_bfgs(f, np.log([2., 2., .1, 1.]), g, disp=False)
    assert np.max(np.abs(r[1:] - [0.95405, -0.9769, 0.36469])) < 5e-4


@pytest.mark.parametrize("kind", [orc.K_SE_ARD, orc.K_SE_ISO, orc.K_MATERN5, orc.K_MATERN5_ARD])
def test_kernel_derivative_is_the_gradient_of_cov(kind):
    """finite differences of covMatrix in log hyperparameters (every kind whose reference derivative is exact)"""
    rs = np.random.RandomState(11)
    X = rs.rand(7, 3)
    hyper = {orc.K_SE_ARD: [.4, .6, .9, 1.3], orc.K_SE_ISO: [.7], orc.K_MATERN5: [.8, 1.2], orc.K_MATERN5_ARD: [.5, .7, .9, 1.1]}[kind]
    k = orc.KernelSpec(kind, hyper, 3)
    for hp in range(len(hyper)):
        e = np.zeros(len(hyper)); e[hp] = 1e-6
        kp = orc.KernelSpec(kind, np.exp(np.log(hyper) + e), 3)
        km = orc.KernelSpec(kind, np.exp(np.log(hyper) - e), 3)
        fd = (orc.cov_matrix(kp, X) - orc.cov_matrix(km, X)) / 2e-6
        assert np.max(np.abs(fd - orc.kernel_derivative(k, X, hp))) < 1e-8


# ---- fastUCBGallery restatement (ego/acquisition/gallery.py:42-135) -------------------------------------------
def _gallery_case():
    rs = np.random.RandomState(11)
    bounds = [[0., 4.], [-1., 3.], [0., 2.]]
    X = np.c_[rs.rand(14) * 4, rs.rand(14) * 4 - 1, rs.rand(14) * 2]
    Y = np.sin(X[:, 0]) + np.cos(X[:, 1]) - (X[:, 2] - 1) ** 2
    return bounds, X, Y, [1.0, 0.8, 0.6]


def test_oracle_gallery_properties():
    """ego/unittest_IBO.py:844-870: nothing out of bounds, a fixed dimension stays pinned; plus the 0.5 distance rule
    (gallery.py:102,113), the seeded determinism, and useBest putting the best in-bounds observation first (:50-63)"""
    bounds, X, Y, theta = _gallery_case()
    k = orc.KernelSpec(orc.K_SE_ARD, theta, 3)
    gal = orc.fast_ucb_gallery(k, X, Y, bounds, 4, samples=100, seed=5, maxiter=12)
    assert len(gal) == 4 and np.array_equal(gal[0], X[np.argmax(Y)])
    for i, x in enumerate(gal):
        assert all(b[0] <= v <= b[1] for v, b in zip(x, bounds))
        assert all(np.linalg.norm(x - gal[j]) > .5 for j in range(i))
    again = orc.fast_ucb_gallery(k, X, Y, bounds, 4, samples=100, seed=5, maxiter=12)
    assert all(np.array_equal(a, b) for a, b in zip(gal, again))
    pinned = [[0., 4.], [-1., 3.], [1., 1.]]                      # unittest_IBO.py:864 pins a dimension
    for x in orc.fast_ucb_gallery(k, X, Y, pinned, 3, samples=60, seed=1, maxiter=8):
        # the seeded best observation (slot 0) may lie outside a pinned box only if it passed the in-bounds filter
        assert all(b[0] <= v <= b[1] for v, b in zip(x, pinned))


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libego.so")),
                    reason="oracle/_ref not built")
def test_oracle_gallery_direct_step_equals_live_acqmaxGP(capfd):
    """the DIRECT step of a gallery slot (maximizeEI(hallucGP, xi=.3), gallery.py:101) is libego's acqmaxGP on inv(R)"""
    import ctypes
    from ctypes import POINTER, c_double, c_int
    bounds, X, Y, theta = _gallery_case()
    gp = orc.GPOracle(orc.KernelSpec(orc.K_SE_ARD, theta, 3), X, Y, noise=0.1)
    opt, optx = orc.acqmax_cpp(gp, bounds, orc.ACQ_EI, .3, maxiter=12)
    ego = ctypes.CDLL(os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libego.so"))
    pd = POINTER(c_double)
    ego.acqmaxGP.restype = pd
    ego.acqmaxGP.argtypes = [c_int, pd, pd, pd, pd, pd, c_int, c_int, c_int, pd, c_int, pd, pd, c_double, pd, pd,
                             c_double, c_double, c_int, c_int, c_int]
    dp = lambda a: a.ctypes.data_as(pd)
    b = np.array(bounds)
    lb = np.ascontiguousarray(b[:, 0]); ub = np.ascontiguousarray(b[:, 1])
    invR = np.ascontiguousarray(gp.invR()); Xc = np.ascontiguousarray(X); Yc = np.ascontiguousarray(Y)
    hyper = np.array(theta); dummy = np.zeros(1)
    res = ego.acqmaxGP(3, dp(lb), dp(ub), dp(invR), dp(Xc), dp(Yc), len(Y), 0, 0, dp(hyper), 0, dp(dummy), dp(dummy), 0.0,
                       dp(dummy), dp(dummy), .3, 0.1, 12, 100000, 10000)
    capfd.readouterr()
    assert abs(-res[0] - opt) <= 1e-12 * max(abs(opt), 1e-3)
    assert np.array_equal(optx, [res[1], res[2], res[3]])
