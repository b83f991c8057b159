"""
Covariance functions with the reference's class names and `cov` / `covMatrix` / `hyperparams`
surface (ego/gaussianprocess/kernel.py).  The classes are light descriptors: `spec()` tells the
CUDA library which kernel type and hyperparameter vector to use; scalar `cov` exists for
host-side bookkeeping (e.g. a 1x1 query) and documentation of the formula, the batched work is
done on the GPU by K1 (ibo_b200/csrc/score.cu) and the R build (ibo_b200/csrc/model.cu).
`covMatrix` and `derivative` (hyper-parameter learning, ego/gaussianprocess/trainhyper.py) are evaluated on the GPU
by ibo_kernel_matrix (ibo_b200/csrc/hyper.cu); `derivative` follows the reference expressions (kernel.py:92-266),
including its Matern-3/2 quirk, see include/ibo_b200.h: IBO_FLAG_GRAD_EXACT.
"""
import math

import numpy as np

from .. import _lib


class Kernel(object):
    def __init__(self, hyperparams):
        self._hyperparams = np.array(hyperparams, dtype=float)
        self._hyperparams.setflags(write=False)      # read-only, as kernel.py:34-40

    def getHyperparams(self):
        return self._hyperparams

    hyperparams = property(getHyperparams)

    def spec(self, ndim):
        """-> (kernel type id, hyperparameter vector) for the C ABI."""
        raise NotImplementedError

    def cov(self, x1, x2):
        raise NotImplementedError('kernel-derived class does not have cov method')

    def covMatrix(self, X):
        """K[i, j] = cov(X[i], X[j]) (kernel.py:43-52), one device launch instead of N^2/2 interpreted calls"""
        X = np.vstack(X).astype(float)
        kind, hyper = self.spec(X.shape[1])
        return _lib.kernel_matrix(kind, hyper, X, -1)

    def _nhyper_device(self, ndim):
        """(kind, hyper vector, list mapping the reference's hyperparameter index -> device indices to sum)"""
        kind, hyper = self.spec(ndim)
        return kind, hyper, [[h] for h in range(len(hyper))]

    def derivative(self, X, hp):
        """dK / d log hyperparams[hp] as the reference's Kernel.derivative returns it (raises ValueError past the end)"""
        X = np.vstack(X).astype(float)
        kind, hyper, hmap = self._nhyper_device(X.shape[1])
        if hp < 0 or hp >= len(hmap):
            raise ValueError
        out = None
        for h in hmap[hp]:
            D = _lib.kernel_matrix(kind, hyper, X, h)
            out = D if out is None else out + D
        return out


class SVKernel(object):
    """signal-variance mixin (kernel.py:60-68)"""

    def __init__(self, mag):
        self._magnitude = mag
        self._sf2 = math.exp(2.0 * math.log(mag))

    def covScale(self, k):
        return self._sf2 * k


class GaussianKernel_iso(Kernel):
    """exp(-1/2 |x1-x2|^2 / theta^2)  (kernel.py:71-89)"""

    def __init__(self, hyperparams, **kwargs):
        super(GaussianKernel_iso, self).__init__(hyperparams)
        self._itheta2 = 1 / float(self._hyperparams[0]) ** 2

    def spec(self, ndim):
        return _lib.KERNEL_SE_ISO, np.array([self._hyperparams[0]])

    def _nhyper_device(self, ndim):
        kind, hyper = self.spec(ndim)
        return kind, hyper, [[0]]          # kernel.py:92-105: hp == 0 only

    def cov(self, x1, x2):
        diff = np.asarray(x1, dtype=float) - np.asarray(x2, dtype=float)
        return math.exp(-.5 * np.linalg.norm(diff) ** 2 * self._itheta2)


class GaussianKernel_ard(Kernel):
    """exp(-1/2 sum_j (x1_j-x2_j)^2 / theta_j^2), theta clipped to [1e-4, 1e4]  (kernel.py:130-149)"""

    def __init__(self, hyperparams, **kwargs):
        super(GaussianKernel_ard, self).__init__(hyperparams)
        self._theta = np.clip(self._hyperparams, 1e-4, 1e4)
        self._itheta2 = 1.0 / self._theta ** 2

    def spec(self, ndim):
        if len(self._theta) != ndim:
            raise ValueError("ARD kernel has %d length scales for %d dimensions" % (len(self._theta), ndim))
        return _lib.KERNEL_SE_ARD, self._theta.copy()

    def cov(self, x1, x2):
        diff = np.asarray(x1, dtype=float) - np.asarray(x2, dtype=float)
        return math.exp(-.5 * np.sum(self._itheta2 * diff ** 2))


class SVGaussianKernel_iso(SVKernel, GaussianKernel_iso):
    """kernel.py:109-127: hyperparams = [theta, magnitude]"""

    def __init__(self, hyperparams, **kwargs):
        GaussianKernel_iso.__init__(self, hyperparams[:-1])
        SVKernel.__init__(self, hyperparams[-1])
        self._hyperparams = np.array(hyperparams, dtype=float)
        self._hyperparams.setflags(write=False)

    def spec(self, ndim):
        # isotropic SE with a signal variance == ARD with equal length scales + magnitude
        return _lib.KERNEL_SE_ARD, np.array([self._hyperparams[0]] * ndim + [self._magnitude])

    def _nhyper_device(self, ndim):
        kind, hyper = self.spec(ndim)
        return kind, hyper, [list(range(ndim)), [ndim]]      # d/dlog theta = sum of the per-dimension derivatives

    def cov(self, x1, x2):
        return self.covScale(GaussianKernel_iso.cov(self, x1, x2))


class SVGaussianKernel_ard(SVKernel, GaussianKernel_ard):
    """kernel.py:169-188: hyperparams = [theta_1..theta_D, magnitude]"""

    def __init__(self, hyperparams, **kwargs):
        GaussianKernel_ard.__init__(self, hyperparams[:-1])
        SVKernel.__init__(self, hyperparams[-1])
        self._hyperparams = np.array(hyperparams, dtype=float)
        self._hyperparams.setflags(write=False)

    def spec(self, ndim):
        return _lib.KERNEL_SE_ARD, np.concatenate([self._theta, [self._magnitude]])

    def cov(self, x1, x2):
        return self.covScale(GaussianKernel_ard.cov(self, x1, x2))


class MaternKernel3(Kernel):
    """sf2 (1+z) exp(-z), z = sqrt(3) |x1-x2| / theta; hyperparams = [theta, magnitude]  (kernel.py:191-210)"""

    def __init__(self, hyperparams, **kwargs):
        super(MaternKernel3, self).__init__(hyperparams)
        self._theta = float(self._hyperparams[0])
        self._magnitude = float(self._hyperparams[1])
        self._sf2 = math.exp(2.0 * math.log(self._magnitude))
        self.sqrt3 = math.sqrt(3)

    def spec(self, ndim):
        return _lib.KERNEL_MATERN3, np.array([self._theta, self._magnitude])

    def cov(self, x1, x2):
        diff = np.asarray(x1, dtype=float) - np.asarray(x2, dtype=float)
        z = self.sqrt3 * np.linalg.norm(diff) / self._theta
        return self._sf2 * (1.0 + z) * math.exp(-z)


class MaternKernel5(Kernel):
    """sf2 (1 + sqrt5 r/theta + 5 r^2/(3 theta^2)) exp(-sqrt5 r/theta); hyperparams = [theta, magnitude].
    The reference's Python `cov` is broken (prints, returns None: kernel.py:246-249); this is the
    intended formula, identical to cpp/optimizeGP.cpp:99-108."""

    def __init__(self, hyperparams, **kwargs):
        super(MaternKernel5, self).__init__(hyperparams)
        self._theta = float(self._hyperparams[0])
        self._magnitude = float(self._hyperparams[1])
        self._sf2 = math.exp(2.0 * math.log(self._magnitude))

    def spec(self, ndim):
        return _lib.KERNEL_MATERN5, np.array([self._theta, self._magnitude])

    def cov(self, x1, x2):
        diff = np.asarray(x1, dtype=float) - np.asarray(x2, dtype=float)
        z = math.sqrt(5.0) * np.linalg.norm(diff) / self._theta
        return self._sf2 * (1.0 + z + z * z / 3.0) * math.exp(-z)


class MaternKernel5_ard(Kernel):
    """Matern-5/2 with per-dimension length scales (BASELINE.json config #4; no reference class):
    hyperparams = [theta_1..theta_D, magnitude]."""

    def __init__(self, hyperparams, **kwargs):
        super(MaternKernel5_ard, self).__init__(hyperparams)
        self._theta = np.array(self._hyperparams[:-1], dtype=float)
        self._magnitude = float(self._hyperparams[-1])
        self._sf2 = math.exp(2.0 * math.log(self._magnitude))

    def spec(self, ndim):
        if len(self._theta) != ndim:
            raise ValueError("ARD kernel has %d length scales for %d dimensions" % (len(self._theta), ndim))
        return _lib.KERNEL_MATERN5_ARD, np.concatenate([self._theta, [self._magnitude]])

    def cov(self, x1, x2):
        diff = (np.asarray(x1, dtype=float) - np.asarray(x2, dtype=float)) / self._theta
        z = math.sqrt(5.0) * math.sqrt(np.sum(diff ** 2))
        return self._sf2 * (1.0 + z + z * z / 3.0) * math.exp(-z)
