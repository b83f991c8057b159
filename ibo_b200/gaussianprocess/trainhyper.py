"""
Marginal likelihood and its gradient for hyper-parameter learning, with the reference's names and call shapes
(ego/gaussianprocess/trainhyper.py:47-136).  The numerics run on the GPU (ibo_nlml, ibo_b200/csrc/hyper.cu): blocked
DMMA Cholesky, inv(K) = W^T W as a triangular tile GEMM, derivative matrices fused into one reduction -- instead of
the reference's dense solves against the identity and one interpreted N x N derivative matrix per hyperparameter.

Typical use, as in the reference's tests (ego/unittest_GP.py:215,252):
    scipy.optimize.fmin_bfgs(nlml, log(hyper), dnlml, args=(SVGaussianKernel_ard, X, Y))
"""
import numpy as np

from .. import _lib


def marginalLikelihood(kernel, X, Y, nhyper, computeGradient=True, useCholesky=True, noise=1e-3, exact=False):
    """-> nlml or (nlml, dnlml[nhyper]); raises numpy.linalg.LinAlgError when K is not positive definite.

    `useCholesky=False` (the explicit-inverse branch, trainhyper.py:78-95) computes the same quantities; it is served by
    the same Cholesky-based device path.  `exact=True` switches Matern-3/2 to its analytic length-scale derivative.
    """
    X = np.vstack(X).astype(float)
    Y = np.asarray(Y, dtype=float).reshape(-1)
    assert len(X) == len(Y)
    kind, hyper, hmap = kernel._nhyper_device(X.shape[1])
    flags = _lib.FLAG_GRAD_EXACT if exact else 0
    if not computeGradient:
        return _lib.nlml(kind, hyper, X, Y, noise, want_grad=False, flags=flags)[0]
    if nhyper > len(hmap):
        raise ValueError("kernel has %d hyperparameters, gradient asked for %d" % (len(hmap), nhyper))
    val, g = _lib.nlml(kind, hyper, X, Y, noise, want_grad=True, flags=flags)
    return val, np.array([sum(g[h] for h in hmap[i]) for i in range(nhyper)])


def nlml(loghyper, kernel, X, Y, *args):
    """negative log marginal likelihood at exp(loghyper); 100 when K is not positive definite (trainhyper.py:99-115)"""
    k = kernel(np.exp(loghyper))
    try:
        return marginalLikelihood(k, X, Y, len(loghyper), computeGradient=False)
    except np.linalg.LinAlgError as e:
        print(e)
        print('returning nlml = 100')
        return 100


def nlmlMulti(loghyper, kernel, X, Y, *args):
    """sum of the negative log marginal likelihoods over several data sets (trainhyper.py:118-127)"""
    k = kernel(np.exp(loghyper))
    ml = 0.0
    for x, y in zip(X, Y):
        ml += marginalLikelihood(k, x, y, len(loghyper))[0]
    return ml


def dnlml(loghyper, kernel, X, Y):
    """gradient of nlml with respect to the log hyperparameters (trainhyper.py:130-136)"""
    k = kernel(np.exp(loghyper))
    return marginalLikelihood(k, X, Y, len(loghyper), computeGradient=True)[1]
