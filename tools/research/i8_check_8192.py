"""config #4 shape (N = 8192, d = 10, Matern-5/2 ARD): INT8-emulated K2 against the FP64 DMMA path -- agreement and device time"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from ibo_b200 import _lib
from ibo_b200.gaussianprocess import GaussianProcess
from ibo_b200.gaussianprocess.kernel import MaternKernel5_ard

_lib.require_gpu()
N, d, M = 8192, 10, 131072
rs = np.random.RandomState(4)
X = rs.rand(N, d)
Y = np.sin(2 * X).sum(axis=1)
theta = [0.5 + 0.05 * j for j in range(d)]
gp = GaussianProcess(MaternKernel5_ard(theta + [1.0]), X, Y, noise=0.1)
Xs = np.random.RandomState(7).rand(M, d)
cands = _lib.ResidentCandidates(gp.model, Xs)
a, b = np.empty(M), np.empty(M)
ra = cands.score(_lib.ACQ_EI, Y.max(), 0.01, _lib.FLAG_MODE_CPP, scores_out=a)
rb = cands.score(_lib.ACQ_EI, Y.max(), 0.01, _lib.FLAG_MODE_CPP | _lib.FLAG_INT8, scores_out=b)
ta = min(cands.score(_lib.ACQ_EI, Y.max(), 0.01, _lib.FLAG_MODE_CPP)[2] for _ in range(2))
tb = min(cands.score(_lib.ACQ_EI, Y.max(), 0.01, _lib.FLAG_MODE_CPP | _lib.FLAG_INT8)[2] for _ in range(3))
pa = gp.model.posterior(Xs[:4096 * 2], flags=_lib.FLAG_MODE_CPP)
pb = gp.model.posterior(Xs[:4096 * 2], flags=_lib.FLAG_MODE_CPP | _lib.FLAG_INT8)
rel = lambda x, y, fl: float(np.max(np.abs(x - y) / np.maximum(np.abs(y), fl)))
print("N=8192 d=10 Matern-5/2 ARD, %d candidates: DMMA %.1f ms (%.3f M evals/s), INT8 %.1f ms (%.3f M evals/s) = %.2fx; "
      "EI rel %.2e, mu rel %.2e, s2 rel %.2e, argmax %d / %d"
      % (M, ta, M / ta / 1e3, tb, M / tb / 1e3, ta / tb, rel(b, a, 1e-5), rel(pb[0], pa[0], 1e-3), rel(pb[1], pa[1], 1e-300), rb[1], ra[1]))
