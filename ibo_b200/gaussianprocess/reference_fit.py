"""The reference's own optimiser for the Laplace latents of PrefGaussianProcess, kept for `reference_exact = True` on small problems:
scipy.optimize.fmin_bfgs on NUMERICAL gradients of the MAP functional S (ego/gaussianprocess/__init__.py:373-386,441-442).  The
default path minimises the same functional on the device (ibo_pref_fit); nothing else in the package imports SciPy."""
import numpy as np

from . import erf


def bfgs_latents(L, vi, ui, dg, start):
    """argmin_x  -sum (d+1) log(CDF((x_v - x_u)/sqrt2) + 1e-10) + |L^-1 x|^2 / 2, L = chol(R) (host copy)"""
    from scipy.linalg import solve_triangular
    from scipy.optimize import fmin_bfgs
    verf = np.vectorize(erf, otypes=[float])

    def S(x):
        z = (x[vi] - x[ui]) / np.sqrt(2)
        cdf = 0.5 * (1 + verf(z * 0.707106))
        Lx = solve_triangular(L, x, lower=True)
        return -np.sum((dg + 1) * np.log(cdf + 1e-10)) + np.dot(Lx, Lx) / 2

    return np.asarray(fmin_bfgs(S, start, disp=0), dtype=float)
