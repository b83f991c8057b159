"""RBF-network mean prior: evaluation only (ego/gaussianprocess/prior.py:41-73).  `train`
(k-means + ridge fit, prior.py:76-156) is a one-off host fit outside the hot path; a prior is built
from already-fitted (means, beta, theta, lowerb, width).  Batched evaluation for the GP posterior
happens on the GPU (K3 in ibo_b200/csrc/score.cu); `mu` here serves host-side callers such as
fastUCBGallery's no-data branch."""
import numpy as np


class GPMeanPrior(object):
    def mu(self, x):
        raise NotImplementedError('GPMeanPrior-derived class does not have mean function implemented')


class RBFNMeanPrior(GPMeanPrior):
    def __init__(self, means=None, beta=None, theta=10., lowerb=None, width=None):
        super(RBFNMeanPrior, self).__init__()
        self.means = means
        self.beta = beta
        self.theta = theta
        self.lowerb = lowerb
        self.width = width

    def RBF(self, r):
        return np.exp(-self.theta * r ** 2)

    def mu(self, x):
        # x arrives in the source space; the network lives in the unit cube (prior.py:60-63)
        u = (np.asarray(x, dtype=float) - self.lowerb) / self.width
        r = np.array([np.linalg.norm(m - u) for m in np.asarray(self.means, dtype=float)])
        return float(np.sum(np.asarray(self.beta) * self.RBF(r)))

    def negmu(self, x):
        return -self.mu(x)

    def train(self, *args, **kwargs):
        raise NotImplementedError("RBFNMeanPrior.train is a host-side fit outside the acquisition hot path "
                                  "(SURVEY.md 2.1 row 3); construct the prior from fitted parameters")
