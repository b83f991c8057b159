"""
fastUCBGallery (ego/acquisition/gallery.py:42-135): a greedy gallery of N points built from
hallucinated observations.  Per slot: maximizeEI(hallucGP, xi=.3) through the batched DIRECT driver,
`samples` latin-hypercube candidates scored with EI(xi=.4) in ONE GPU batch (the reference loops a
Python posterior per sample, :111-116), optionally the prior means, then hallucGP.addData(best, mu(best)).
"""
from copy import deepcopy

import numpy as np
from numpy.linalg import norm

from ..gaussianprocess import GaussianProcess
from ..utils.latinhypercube import lhcSample


def fastUCBGallery(GP, bounds, N, useBest=True, samples=300, useCDIRECT=True, seed=None):
    """`seed` (extension) fixes the latin-hypercube draws, which the reference leaves unseeded (:111)."""
    from . import EI, maximizeEI
    gallery = []
    if len(GP.X) > 0:
        if useBest:
            # best sample already seen that lies within the bounds (:50-63)
            bestY, bestX = -np.inf, None
            for x, y in zip(GP.X, GP.Y):
                if y > bestY and all(b[0] <= v <= b[1] for v, b in zip(x, bounds)):
                    bestY, bestX = y, x
            if bestX is not None:
                gallery.append(bestX)
        # a plain GP with default noise even when the source is a PrefGP; C is dropped (:67)
        hallucGP = GaussianProcess(deepcopy(GP.kernel), deepcopy(GP.X), deepcopy(GP.Y), prior=GP.prior, device=GP.device)
    elif GP.prior is None:
        x = np.array([(b[0] + b[1]) / 2. for b in bounds])
        gallery.append(x)
        hallucGP = GaussianProcess(deepcopy(GP.kernel), [x], [0.0], prior=GP.prior, device=GP.device)
    else:
        # no data: start from the best prior mean (:73-90)
        from scipy.optimize import fmin_bfgs
        bestmu, bestX = -np.inf, None
        for m in GP.prior.means:
            argmin = fmin_bfgs(GP.negmu, m, disp=False)
            argmin = np.array([np.clip(argmin[i], bounds[i][0], bounds[i][1]) for i in range(len(argmin))])
            if GP.mu(argmin) > bestmu:
                bestX, bestmu = argmin, GP.mu(argmin)
        gallery.append(bestX)
        hallucGP = GaussianProcess(deepcopy(GP.kernel), bestX, bestmu, prior=GP.prior, device=GP.device)

    slot = 0
    while len(gallery) < N:
        bestUCB, bestX = -np.inf, None
        ut = EI(hallucGP, xi=.4)
        opt, optx = maximizeEI(hallucGP, bounds, xi=.3, useCDIRECT=useCDIRECT)
        if len(gallery) == 0 or min(norm(optx - gx) for gx in gallery) > .5:
            bestUCB, bestX = opt, optx
        # latin-hypercube candidates, one batch (:111-116)
        cand = np.array(lhcSample(bounds, samples, seed=None if seed is None else seed + slot))
        u = ut.f_batch(cand)
        for x, ux in zip(cand, u):
            if ux > bestUCB and min(norm(x - gx) for gx in gallery) > .5:
                bestUCB, bestX = ux, x
        # prior means (:119-130)
        if hallucGP.prior is not None:
            pm = np.array([[np.clip(x[i], bounds[i][0], bounds[i][1]) for i in range(len(x))] for x in hallucGP.prior.means])
            pm = pm * hallucGP.prior.width + hallucGP.prior.lowerb
            for x, ux in zip(pm, ut.f_batch(pm)):
                if ux > bestUCB and (len(gallery) == 0 or min(norm(x - gx) for gx in gallery) > .5):
                    bestUCB, bestX = ux, x
        gallery.append(bestX)
        hallucGP.addData(bestX, hallucGP.mu(bestX))
        slot += 1
    return gallery
