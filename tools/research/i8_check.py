"""GPU check of the experimental int8-emulated K2 (IBO_FLAG_INT8) against the default DMMA path and the oracle."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from ibo_b200 import _lib
from ibo_b200.gaussianprocess import GaussianProcess
from ibo_b200.gaussianprocess.kernel import GaussianKernel_ard, MaternKernel3
from oracle import ibo_oracle as orc


def one(N, d, M, kernel="se", full=True):
    rs = np.random.RandomState(N + d)
    X = rs.rand(N, d)
    Y = np.sin(3 * X).sum(axis=1)
    Xs = rs.rand(M, d)
    if kernel == "se":
        theta = list(0.3 + 0.1 * np.arange(d))
        gp = GaussianProcess(GaussianKernel_ard(theta), X, Y, noise=0.1)
    else:
        gp = GaussianProcess(MaternKernel3([0.7, 1.0]), X, Y, noise=0.1)
    ymax = Y.max()
    a = gp.model.score(Xs, _lib.ACQ_EI, ymax, 0.01, flags=_lib.FLAG_MODE_CPP, want_posterior=True)
    t0 = time.perf_counter()
    b = gp.model.score(Xs, _lib.ACQ_EI, ymax, 0.01, flags=_lib.FLAG_MODE_CPP | _lib.FLAG_INT8, want_posterior=True)
    t1 = time.perf_counter()
    sc0, mu0, s20 = a[0], a[1], a[2]
    sc1, mu1, s21 = b[0], b[1], b[2]
    rel = lambda x, y, fl: float(np.max(np.abs(x - y) / np.maximum(np.abs(y), fl)))
    print("N=%d d=%d M=%d %s: mu rel %.2e  s2 rel %.2e  EI rel %.2e  argmax %d/%d  (%.1f ms)" % (
        N, d, M, kernel, rel(mu1, mu0, 1e-3), rel(s21, s20, 1e-300), rel(sc1, sc0, 1e-5), b[4], a[4], 1e3 * (t1 - t0)), flush=True)
    if rel(s21, s20, 1e-300) > 1e-9 or rel(mu1, mu0, 1e-3) > 1e-9:
        bad = np.argsort(-np.abs(s21 - s20))[:5]
        print("   worst s2:", [(int(i), s21[i], s20[i]) for i in bad])
        print("   first 8 q=1.1-s2 ratio:", ((1.1 - s21[:8]) / (1.1 - s20[:8])))
        print("   first 8 mu ratio:", mu1[:8] / mu0[:8])
        print("   tile means of |ds2|:", np.abs(s21 - s20)[:1024].reshape(16, 64).mean(axis=1))


if __name__ == "__main__":
    _lib.require_gpu()
    one(100, 3, 2100)
    one(300, 3, 5000)
    one(300, 2, 4096, kernel="matern3")
    one(2048, 6, 40000)
    one(1000, 20, 3000)
