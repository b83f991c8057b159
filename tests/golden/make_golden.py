"""
Generates the committed golden vectors under tests/golden/ by running the REFERENCE ITSELF:
the unmodified reference C++ (cpp/optimizeGP.cpp, cpp/direct.cpp) compiled by oracle/build_ref.sh
into oracle/_ref/.  Run in the build container (where /root/reference exists):

    bash oracle/build_ref.sh && python tests/golden/make_golden.py

Scenarios follow the reference's own unit tests (ego/unittest_IBO.py:107-134,138-211,427-578):
Shekel5 / Branin LHS fixtures with seed 0, SE-iso / Matern3 / SE-ARD kernels, xi and noise sweeps,
plus the RBF-network prior path.  Inputs are stored together with outputs so that the tests need
neither /root/reference nor oracle/_ref at run time.
"""
import ctypes
import json
import os
import sys
from ctypes import POINTER, c_double, c_int, c_long

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ibo_oracle as orc  # noqa: E402  (fixture *inputs* only: lhc_sample, test functions)

REF = os.path.join(ROOT, "oracle", "_ref")
pd = POINTER(c_double)


def dp(a):
    return a.ctypes.data_as(pd)


def load():
    ego = ctypes.CDLL(os.path.join(REF, "libego.so"))
    har = ctypes.CDLL(os.path.join(REF, "libego_harness.so"))
    ego.acqmaxGP.restype = pd
    ego.acqmaxGP.argtypes = [c_int, pd, pd, pd, pd, pd, c_int, c_int, c_int, pd, c_int, pd, pd, c_double, pd, pd,
                             c_double, c_double, c_int, c_int, c_int]
    ego.direct.restype = pd
    har.ref_set_model.argtypes = [c_int, pd, pd, pd, c_int, c_int, pd, c_int, pd, pd, c_double, pd, pd, c_double, c_double]
    har.ref_eval.argtypes = [c_int, c_long, pd, pd, pd, pd, c_int]
    return ego, har


def scenario(name):
    """name -> dict(kind, hyper, X, Y, noise, bounds, prior)"""
    if name.startswith("shekel_iso"):
        bounds = [[0., 10.]] * 4
        X = np.array(orc.lhc_sample(bounds, 10, seed=0))
        Y = np.array([-orc.shekel5(x) for x in X])            # maximize=True form (unittest_IBO.py:433-437)
        theta = 0.3 if name.endswith("3") else 0.2
        return dict(kind=1, hyper=[theta], X=X, Y=Y, noise=0.1, bounds=bounds, prior=None)
    if name.startswith("branin_matern3"):
        bounds = [[-5., 10.], [0., 15.]]
        X = np.array(orc.lhc_sample(bounds, 10, seed=0))
        Y = -orc.branin(X) / 100.0
        noise = {"a": 1e-4, "b": 0.01, "c": 0.1}[name[-1]]
        return dict(kind=2, hyper=[1.0, 1.0], X=X, Y=Y, noise=noise, bounds=bounds, prior=None)
    if name == "branin_ard50":                                  # BASELINE.json config #1
        bounds = [[-5., 10.], [0., 15.]]
        X = np.array(orc.lhc_sample(bounds, 50, seed=0))
        Y = -orc.branin(X) / 100.0
        return dict(kind=0, hyper=[3.4, 10.0], X=X, Y=Y, noise=0.1, bounds=bounds, prior=None)
    if name == "hartman_ard200":
        rs = np.random.RandomState(0)
        X = rs.rand(200, 6)
        Y = orc.hartman6_neg(X)
        return dict(kind=0, hyper=[.53, .57, 2.5, .34, .27, .35], X=X, Y=Y, noise=0.1, bounds=[[0., 1.]] * 6, prior=None)
    if name == "prior_iso":                                     # unittest_IBO.py:533-578 style
        bounds = [[0., 10.]] * 4
        X = np.array(orc.lhc_sample(bounds, 12, seed=1))
        Y = np.array([-orc.shekel5(x) for x in X])
        rs = np.random.RandomState(7)
        prior = dict(means=rs.rand(5, 4), beta=rs.randn(5), theta=10.0, lowerb=np.zeros(4), width=np.full(4, 10.0))
        return dict(kind=1, hyper=[0.3], X=X, Y=Y, noise=0.1, bounds=bounds, prior=prior)
    if name == "matern5_1d":                                    # the only Matern-5/2 shape the C++ evaluates consistently
        rs = np.random.RandomState(11)
        X = rs.rand(15, 1) * 4
        Y = np.sin(2 * X[:, 0])
        return dict(kind=3, hyper=[0.8, 0.9], X=X, Y=Y, noise=0.1, bounds=[[0., 4.]], prior=None)
    if name in ("matern5_2d", "matern5_4d", "matern5_10d"):
        # Matern-5/2 beyond one dimension.  GP_Maximizer::posterior evaluates the isotropic formula at any d
        # (cpp/optimizeGP.cpp:99-110) and acqmaxGP reads the magnitude from hyperparams[ndim] (:313), so the reference is well
        # defined when it is handed an array of length ndim + 1: [theta, (unused) ..., magnitude] -- `ref_hyper`.
        d = int(name.split("_")[1][:-1])
        rs = np.random.RandomState(20 + d)
        n = {2: 30, 4: 60, 10: 120}[d]
        X = rs.rand(n, d) * 2.0
        Y = np.sin(2 * X).sum(axis=1) / d
        theta, mag = {2: (0.9, 1.0), 4: (1.3, 0.8), 10: (2.1, 0.9)}[d]     # (R keeps 1 + noise on its diagonal whatever the magnitude: mag <= 1)
        ref_hyper = np.zeros(d + 1)
        ref_hyper[0] = theta; ref_hyper[d] = mag
        return dict(kind=3, hyper=[theta, mag], ref_hyper=ref_hyper, X=X, Y=Y, noise=0.1, bounds=[[0., 2.]] * d, prior=None)
    raise KeyError(name)


class quiet_stdout(object):
    """kerneltype 3 streams every kernel value to std::cout (cpp/optimizeGP.cpp:109): silence fd 1 around the reference's acqmaxGP"""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.null)
        os.close(self.saved)


def model_arrays(sc):
    kern = orc.KernelSpec(sc["kind"], sc["hyper"], sc["X"].shape[1])
    pr = None
    if sc["prior"] is not None:
        p = sc["prior"]
        pr = orc.PriorSpec(p["means"], p["beta"], p["theta"], p["lowerb"], p["width"])
    gp = orc.GPOracle(kern, sc["X"], sc["Y"], sc["noise"], prior=pr)
    invR = np.ascontiguousarray(np.linalg.inv(gp.R))           # ego/acquisition/__init__.py:388
    return gp, invR


def prior_args(sc, d):
    if sc["prior"] is None:
        z = np.zeros(1)
        return 0, z, z.copy(), 0.0, z.copy(), z.copy()
    p = sc["prior"]
    return (len(p["beta"]), np.ascontiguousarray(p["means"].reshape(-1)), np.ascontiguousarray(p["beta"]), float(p["theta"]),
            np.ascontiguousarray(p["lowerb"]), np.ascontiguousarray(p["width"]))


def main():
    ego, har = load()
    out = {}
    rs = np.random.RandomState(123)
    # ---- per-candidate values from GP_Maximizer::negei/negpi/negucb/posterior ----
    for name in ["shekel_iso2", "shekel_iso3", "branin_matern3a", "branin_matern3b", "branin_matern3c", "branin_ard50",
                 "hartman_ard200", "prior_iso", "matern5_1d", "matern5_2d", "matern5_4d", "matern5_10d"]:
        sc = scenario(name)
        X = np.ascontiguousarray(sc["X"]); Y = np.ascontiguousarray(sc["Y"])
        N, d = X.shape
        gp, invR = model_arrays(sc)
        hyper = np.ascontiguousarray(np.array(sc["hyper"], dtype=float))
        ref_hyper = np.ascontiguousarray(np.array(sc.get("ref_hyper", sc["hyper"]), dtype=float))
        npb, pm, pb, pt, plb, pw = prior_args(sc, d)
        b = np.array(sc["bounds"])
        M = 64
        Xs = np.ascontiguousarray(b[:, 0] + (b[:, 1] - b[:, 0]) * rs.rand(M, d))
        Xs[:3] = X[:3]
        rec = dict(kind=sc["kind"], hyper=hyper, ref_hyper=ref_hyper, X=X, Y=Y, noise=sc["noise"], invR=invR, Xs=Xs, bounds=b)
        if sc["prior"] is not None:
            rec.update(p_means=sc["prior"]["means"], p_beta=sc["prior"]["beta"], p_theta=sc["prior"]["theta"],
                       p_lowerb=sc["prior"]["lowerb"], p_width=sc["prior"]["width"])
        for acq, parm, tag in [(0, 0.01, "negei"), (0, 0.1, "negei_xi1"), (1, 0.01, "negpi"), (2, 1.3, "negucb")]:
            har.ref_set_model(d, dp(invR), dp(X), dp(Y), N, sc["kind"], dp(ref_hyper), npb, dp(pm), dp(pb), pt, dp(plb), dp(pw), parm, sc["noise"])
            v = np.empty(M); mu = np.empty(M); sg = np.empty(M)
            har.ref_eval(acq, M, dp(Xs), dp(v), dp(mu), dp(sg), 1)
            rec[tag] = v
            rec["mu"] = mu
            rec["sigma"] = sg
        for k, v in rec.items():
            out["%s/%s" % (name, k)] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "ref_candidates.npz"), **out)

    # ---- acqmaxGP end-to-end (DIRECT over the reference objective) ----
    acq_out = {}
    for name, acq, parm, maxiter, maxsample in [
            ("shekel_iso3", 0, 0.0, 20, 10000), ("shekel_iso3", 0, 0.01, 20, 10000), ("shekel_iso3", 0, 0.1, 20, 10000),
            ("shekel_iso2", 1, 0.01, 20, 10000), ("branin_matern3a", 0, 0.01, 20, 10000), ("branin_matern3b", 0, 0.01, 20, 10000),
            ("branin_matern3c", 0, 0.01, 20, 10000), ("branin_ard50", 0, 0.01, 50, 10000), ("branin_ard50", 2, 1.5, 30, 10000),
            ("hartman_ard200", 0, 0.01, 30, 3000), ("prior_iso", 0, 0.01, 20, 10000),
            ("matern5_2d", 0, 0.01, 20, 10000), ("matern5_4d", 0, 0.01, 12, 3000), ("matern5_10d", 2, 1.5, 6, 1500)]:
        sc = scenario(name)
        X = np.ascontiguousarray(sc["X"]); Y = np.ascontiguousarray(sc["Y"])
        N, d = X.shape
        gp, invR = model_arrays(sc)
        hyper = np.ascontiguousarray(np.array(sc.get("ref_hyper", sc["hyper"]), dtype=float))
        npb, pm, pb, pt, plb, pw = prior_args(sc, d)
        b = np.array(sc["bounds"])
        lb = np.ascontiguousarray(b[:, 0]); ub = np.ascontiguousarray(b[:, 1])
        with quiet_stdout():
            res = ego.acqmaxGP(d, dp(lb), dp(ub), dp(invR), dp(X), dp(Y), N, acq, sc["kind"], dp(hyper), npb, dp(pm), dp(pb), pt,
                               dp(plb), dp(pw), parm, sc["noise"], maxiter, 100000, maxsample)
        key = "%s|acq%d|parm%g|it%d|ms%d" % (name, acq, parm, maxiter, maxsample)
        acq_out[key] = dict(fmin=res[0], xmin=[res[i + 1] for i in range(d)])
    # ---- the reference's C DIRECT on analytic functions, full sample trace ----
    OBJ = ctypes.CFUNCTYPE(c_double, c_int, pd)
    ego.direct.argtypes = [OBJ, c_int, pd, pd, c_int, c_int, c_int]
    dir_out = {}

    def run_direct(tag, f, bounds, maxiter, maxsample):
        trace = []

        def cb(n, x):
            xx = np.array([x[i] for i in range(n)])
            trace.append(xx)
            return float(f(xx))
        b = np.array(bounds, dtype=float)
        lb = np.ascontiguousarray(b[:, 0]); ub = np.ascontiguousarray(b[:, 1])
        res = ego.direct(OBJ(cb), len(lb), dp(lb), dp(ub), maxiter, 100000, maxsample)
        tr = np.array(trace)
        dir_out[tag] = dict(bounds=b.tolist(), maxiter=maxiter, maxsample=maxsample, fmin=res[0],
                            xmin=[res[i + 1] for i in range(len(lb))], nsamples=len(trace),
                            trace_head=tr[:40].tolist(), trace_tail=tr[-10:].tolist(),
                            trace_sum=[float(v) for v in tr.sum(axis=0)])

    run_direct("shekel5_it20", orc.shekel5, [[0., 10.]] * 4, 20, 200000)           # unittest_IBO.py:107-114
    run_direct("shekel5_ms300", orc.shekel5, [[0., 10.]] * 4, 1000, 300)
    run_direct("branin_it15", lambda x: float(orc.branin(x)), [[-5., 10.], [0., 15.]], 15, 200000)
    run_direct("quad_fixed_dim2", lambda x: float(np.sum((x - 0.3) ** 2)), [[0., 1.], [0., 1.], [0.5, 0.5]], 12, 200000)
    run_direct("quad_fixed_dim0", lambda x: float(np.sum((x - 0.3) ** 2)), [[0.5, 0.5], [0., 1.], [0., 1.]], 12, 200000)
    run_direct("sin6_it8", lambda x: float(np.sum(np.sin(3 * x) + (x - .4) ** 2)), [[0., 1.]] * 6, 8, 200000)
    with open(os.path.join(HERE, "ref_direct.json"), "w") as fh:
        json.dump(dict(acqmaxGP=acq_out, direct=dir_out), fh, indent=1, sort_keys=True)
    print("wrote ref_candidates.npz (%d arrays), ref_direct.json" % len(out))


if __name__ == "__main__":
    main()
