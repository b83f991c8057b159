"""
TEST INFRASTRUCTURE ONLY -- CPU oracle for the IBO acquisition hot path.

Nothing in the product path (ibo_b200/) may import this module.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and only as
the checker.

This is a Python-3 / NumPy restatement of the reference's algorithm (the reference is Python 2.6
and cannot be imported here).  Every function cites the reference file:line
(relative to /root/reference) it follows.  Two arithmetic "modes" exist because the reference
itself has two definitions of the same quantities:

  * mode "py"  -- the NumPy/SciPy path: Cholesky solves, sigma^2 floor 10e-8 (=1e-7),
                  Chebyshev erf with 6-digit constants (ego/gaussianprocess/__init__.py:55-77,169-228)
  * mode "cpp" -- the libego path: explicit inverse, sigma^2 floor 1e-8, libm erf
                  (cpp/optimizeGP.cpp:57-236)

Pinning: tests/test_oracle.py checks this restatement against (a) the reference's own C++
library compiled from the untouched sources (oracle/_ref, built by oracle/build_ref.sh),
(b) golden vectors generated from that library (tests/golden/), and (c) the known answers in the
reference's unit tests (ego/unittest_IBO.py:107-134, ego/unittest_GP.py:77-156).
Matern-5/2 (kernel types 3/4) is "parity unpinned" in its Python form: the reference's Python
MaternKernel5.cov is broken (kernel.py:246-249) -- it is pinned only against the C++ kerneltype 3
at d=1.
"""
import math

import numpy as np

# kernel type ids: 0..3 are the reference's (ego/acquisition/__init__.py:323-333)
K_SE_ARD, K_SE_ISO, K_MATERN3, K_MATERN5, K_MATERN5_ARD = 0, 1, 2, 3, 4
ACQ_EI, ACQ_PI, ACQ_UCB = 0, 1, 2


# ---------------------------------------------------------------------------------------------
# erf / CDF / PDF  (ego/gaussianprocess/__init__.py:55-77)
# ---------------------------------------------------------------------------------------------
_NR = (-1.26551223, 1.00002368, 0.37409196, 0.09678418, -0.18628806, 0.27886807,
       -1.13520398, 1.48851587, -0.82215223, 0.17087277)


def erf_py(z):
    """Numerical-Recipes Chebyshev erf, Horner form, sign flip for z<0 (gaussianprocess/__init__.py:55-71)."""
    z = np.asarray(z, dtype=float)
    t = 1.0 / (1.0 + 0.5 * np.abs(z))
    p = _NR[9]
    for c in _NR[8:0:-1]:
        p = c + t * p
    ans = 1 - t * np.exp(-z * z + _NR[0] + t * p)
    return np.where(z >= 0.0, ans, -ans)


def cdf_py(x):
    """gaussianprocess/__init__.py:73-74 (note the truncated 0.707106)."""
    return 0.5 * (1 + erf_py(np.asarray(x, dtype=float) * 0.707106))


def pdf_py(x):
    """gaussianprocess/__init__.py:76-77 (note the truncated 0.398942)."""
    x = np.asarray(x, dtype=float)
    return np.exp(-(x ** 2 / 2)) * 0.398942


_verf = np.vectorize(math.erf, otypes=[float])


def cdf_cpp(z):
    """cpp/optimizeGP.cpp:202 -- 0.5*(1+erf(Z/sqrt(2))) with libm erf."""
    return 0.5 * (1.0 + _verf(np.asarray(z, dtype=float) / math.sqrt(2.0)))


def pdf_cpp(z):
    """cpp/optimizeGP.cpp:203."""
    z = np.asarray(z, dtype=float)
    return np.exp(-(z * z / 2.0)) / math.sqrt(2.0 * math.pi)


# ---------------------------------------------------------------------------------------------
# kernels  (ego/gaussianprocess/kernel.py:87-89,141-149,207-210,246-248; cpp/optimizeGP.cpp:66-113)
# ---------------------------------------------------------------------------------------------
class KernelSpec(object):
    """kind + hyperparameters, with the derived quantities the reference's kernel classes hold."""

    def __init__(self, kind, hyper, ndim):
        self.kind = int(kind)
        self.hyper = np.array(hyper, dtype=float).reshape(-1)
        self.ndim = int(ndim)
        if kind == K_SE_ARD:
            # kernel.py:141-142 -- clip to [1e-4,1e4] (Python only; the C++ uses raw values)
            self.theta = np.clip(self.hyper[:ndim], 1e-4, 1e4)
            self.sf2 = 1.0 if len(self.hyper) <= ndim else math.exp(2.0 * math.log(self.hyper[ndim]))
        elif kind == K_SE_ISO:
            self.theta = np.full(ndim, self.hyper[0])
            self.sf2 = 1.0 if len(self.hyper) < 2 else 1.0  # iso ignores magnitude (kernel.py:80-83)
        elif kind == K_MATERN3:
            self.theta = np.full(ndim, self.hyper[0])
            self.sf2 = math.exp(2.0 * math.log(self.hyper[1])) if len(self.hyper) > 1 else 1.0  # kernel.py:203-204
        elif kind == K_MATERN5:
            self.theta = np.full(ndim, self.hyper[0])
            self.sf2 = math.exp(2.0 * math.log(self.hyper[1])) if len(self.hyper) > 1 else 1.0  # kernel.py:242-243
        elif kind == K_MATERN5_ARD:
            self.theta = self.hyper[:ndim].copy()
            self.sf2 = math.exp(2.0 * math.log(self.hyper[ndim])) if len(self.hyper) > ndim else 1.0
        else:
            raise ValueError("unknown kernel kind %r" % kind)

    def cov(self, x1, x2):
        """Scalar-faithful k(x1,x2)."""
        x1 = np.asarray(x1, dtype=float)
        x2 = np.asarray(x2, dtype=float)
        if self.kind == K_SE_ARD:      # kernel.py:147-149
            return self.sf2 * math.exp(-.5 * np.sum((1.0 / self.theta ** 2) * (x1 - x2) ** 2))
        if self.kind == K_SE_ISO:      # kernel.py:87-89
            return math.exp(-.5 * np.linalg.norm(x1 - x2) ** 2 * (1 / self.hyper[0] ** 2))
        if self.kind == K_MATERN3:     # kernel.py:207-210
            z = math.sqrt(3) * np.linalg.norm(x1 - x2) / self.hyper[0]
            return self.sf2 * (1.0 + z) * math.exp(-z)
        # Matern 5/2: intended formula of kernel.py:246-248 == cpp/optimizeGP.cpp:99-108 (ARD: r^2 = sum (dx/theta_j)^2)
        r = math.sqrt(np.sum(((x1 - x2) / self.theta) ** 2))
        s5r = math.sqrt(5.0) * r
        return self.sf2 * (1.0 + s5r + 5.0 * r * r / 3.0) * math.exp(-s5r)

    def cross(self, X, Xs):
        """Vectorised k(X_i, Xs_j) -> (N, M); same formulas, direct differences (no expansion)."""
        X = np.asarray(X, dtype=float)
        Xs = np.asarray(Xs, dtype=float)
        r2 = np.zeros((X.shape[0], Xs.shape[0]))
        for j in range(X.shape[1]):
            dj = (X[:, j][:, None] - Xs[:, j][None, :]) / self.theta[j]
            r2 += dj * dj
        if self.kind in (K_SE_ARD, K_SE_ISO):
            return self.sf2 * np.exp(-.5 * r2)
        r = np.sqrt(r2)
        if self.kind == K_MATERN3:
            z = math.sqrt(3) * r
            return self.sf2 * (1.0 + z) * np.exp(-z)
        s5r = math.sqrt(5.0) * r
        return self.sf2 * (1.0 + s5r + 5.0 * r2 / 3.0) * np.exp(-s5r)


# ---------------------------------------------------------------------------------------------
# RBF-network mean prior (ego/gaussianprocess/prior.py:60-73; cpp/optimizeGP.cpp:116-134)
# ---------------------------------------------------------------------------------------------
class PriorSpec(object):
    def __init__(self, means, beta, theta, lowerb, width):
        self.means = np.array(means, dtype=float)
        self.beta = np.array(beta, dtype=float)
        self.theta = float(theta)
        self.lowerb = np.array(lowerb, dtype=float)
        self.width = np.array(width, dtype=float)

    def mu(self, x):
        xs = (np.asarray(x, dtype=float) - self.lowerb) / self.width
        d2 = np.sum((self.means - xs[None, :]) ** 2, axis=1)
        return float(np.sum(self.beta * np.exp(-self.theta * d2)))

    def mu_batch(self, Xs):
        xs = (np.asarray(Xs, dtype=float) - self.lowerb[None, :]) / self.width[None, :]
        d2 = np.sum((xs[:, None, :] - self.means[None, :, :]) ** 2, axis=2)
        return np.sum(self.beta[None, :] * np.exp(-self.theta * d2), axis=1)


# ---------------------------------------------------------------------------------------------
# the model  (ego/gaussianprocess/__init__.py:134-149,169-228,267-308,487-498)
# ---------------------------------------------------------------------------------------------
def build_R(kernel, X, noise):
    """R = K_offdiag + (1+noise) I  (gaussianprocess/__init__.py:134-143: `eye(N)+noise`, off-diagonals overwritten)."""
    X = np.asarray(X, dtype=float)
    R = kernel.cross(X, X)
    R = 0.5 * (R + R.T)  # the reference assigns r[i,j] = r[j,i] = cov(X[i], X[j]) (exactly symmetric)
    np.fill_diagonal(R, 1.0 + noise)
    return R


class GPOracle(object):
    """GaussianProcess / fitted PrefGaussianProcess restated (scalar observations Y given)."""

    def __init__(self, kernel, X, Y, noise=0.1, prior=None, Cinv=None):
        self.kernel = kernel
        self.X = np.array(X, dtype=float, ndmin=2)
        self.Y = np.array(Y, dtype=float).reshape(-1)
        self.noise = float(noise)
        self.prior = prior
        self.R = build_R(kernel, self.X, noise)
        A = self.R if Cinv is None else self.R + np.asarray(Cinv, dtype=float)  # :488 L = chol(R + inv(C))
        self.A = A
        self.L = np.linalg.cholesky(A)                                          # :299 / :488
        self.augL = None
        self.augX = None

    def add_data(self, X, Y):
        """Block append of gaussianprocess/__init__.py:300-308: R grows by [m; r], L by [z^T, chol(r - z^T z)]."""
        X = np.array(X, dtype=float, ndmin=2)
        Y = np.array(Y, dtype=float, ndmin=1).flatten()
        r = build_R(self.kernel, X, self.noise)
        m = self.kernel.cross(self.X, X)     # same vectorised arithmetic as build_R, so R stays exactly what a batch build gives
        self.R = np.r_[np.c_[self.R, m], np.c_[m.T, r]]
        z = np.linalg.solve(self.L, m)
        dd = np.linalg.cholesky(r - np.dot(z.T, z))
        self.L = np.r_[np.c_[self.L, np.zeros(z.shape)], np.c_[z.T, dd]]
        self.A = self.R
        self.X = np.r_[self.X, X]
        self.Y = np.r_[self.Y, Y]

    # -- NumPy path ---------------------------------------------------------------------------
    def posterior_scalar(self, x):
        """Scalar-faithful gaussianprocess/__init__.py:169-228: two *general* linalg.solve against L."""
        x = np.asarray(x, dtype=float).reshape(-1)
        m = 0.0 if self.prior is None else self.prior.mu(x)
        d = self.Y - m
        r = np.array([[self.kernel.cov(xi, x)] for xi in self.X])
        Lr = np.linalg.solve(self.L, r)
        mu = m + np.dot(Lr.T, np.linalg.solve(self.L, d))
        if self.augL is None:
            sigma2 = (1 + self.noise) - np.sum(Lr ** 2, axis=0)
        else:
            ra = np.array([[self.kernel.cov(xi, x)] for xi in self.augX])
            La = np.linalg.solve(self.augL, ra)
            sigma2 = (1 + self.noise) - np.sum(La ** 2, axis=0)
        sigma2 = np.clip(sigma2, 10e-8, 10)
        return float(mu[0]), float(sigma2[0])

    def posterior_batch(self, Xs, floor=10e-8):
        """Vectorised equivalent (triangular solves) -> mu[M], sigma2[M]."""
        from scipy.linalg import solve_triangular
        Xs = np.array(Xs, dtype=float, ndmin=2)
        Ks = self.kernel.cross(self.X, Xs)
        V = solve_triangular(self.L, Ks, lower=True)
        bY = solve_triangular(self.L, self.Y, lower=True)
        if self.prior is None:
            mu = V.T.dot(bY)
        else:
            m = self.prior.mu_batch(Xs)
            b1 = solve_triangular(self.L, np.ones_like(self.Y), lower=True)
            mu = m + V.T.dot(bY) - m * V.T.dot(b1)
        if self.augL is None:
            q = np.sum(V * V, axis=0)
        else:
            Va = solve_triangular(self.augL, self.kernel.cross(self.augX, Xs), lower=True)
            q = np.sum(Va * Va, axis=0)
        s2 = np.clip((1 + self.noise) - q, floor, 10)
        return mu, s2

    # -- libego path ---------------------------------------------------------------------------
    def invR(self):
        """ego/acquisition/__init__.py:385-388."""
        return np.linalg.inv(self.A)

    def posterior_cpp(self, Xs, invR=None):
        """cpp/optimizeGP.cpp:57-170: mu = m + r.invR.(Y-m); sigma = sqrt(clip(1+noise - r.invR.r, 1e-8, 10))."""
        Xs = np.array(Xs, dtype=float, ndmin=2)
        invR = self.invR() if invR is None else invR
        Ks = self.kernel.cross(self.X, Xs)                  # (N, M)
        t = invR.dot(Ks)                                    # aMb: Mb = M*b, then a.Mb  (:174-191)
        if self.prior is None:
            mu = invR.dot(self.Y).dot(Ks)
        else:
            m = self.prior.mu_batch(Xs)
            mu = m + invR.dot(self.Y).dot(Ks) - m * invR.dot(np.ones_like(self.Y)).dot(Ks)
        s2 = np.clip(1.0 + self.noise - np.sum(Ks * t, axis=0), 1e-8, 10.0)
        return mu, np.sqrt(s2)


# ---------------------------------------------------------------------------------------------
# hyper-parameter learning: kernel.derivative and the marginal likelihood
#   (ego/gaussianprocess/kernel.py:92-105,120-127,152-166,181-188,212-228,250-266; ego/gaussianprocess/trainhyper.py:47-95)
# ---------------------------------------------------------------------------------------------
def cov_matrix(kernel, X):
    """Kernel.covMatrix (kernel.py:43-52): K[i,j] = K[j,i] = cov(X[i], X[j]) for j <= i, *including* the diagonal
    (cov(x, x) = sf2) -- unlike R (build_R) there is no 1+noise diagonal here."""
    X = np.array(X, dtype=float, ndmin=2)
    K = kernel.cross(X, X)
    K = 0.5 * (K + K.T)
    np.fill_diagonal(K, kernel.sf2)
    return K


def kernel_derivative(kernel, X, hp, exact_matern3=False):
    """Kernel.derivative(X, hp): the matrix dK / d log(hyper[hp]) as the reference computes it.

    hp indexes the reference's hyperparameter vector: length scales first, then (SV / Matern kernels) the magnitude.
      SE iso  hp=0: K o |xi-xj|^2/theta^2                    (kernel.py:92-101)
      SE ARD  hp<D: K o (xi_hp-xj_hp)^2/theta_hp^2           (kernel.py:152-162); K carries sf2 for the SV classes (:181-185)
      SV      hp=last: 2 K                                   (kernel.py:124-127,186-188)
      Matern3 hp=0: sf2 r^2 exp(-r) with r = |xi-xj| *unscaled by theta* (kernel.py:212-223) -- a reference quirk (the
                    derivative of its own cov is sf2 z^2 exp(-z), z = sqrt3 r/theta: `exact_matern3=True`); hp=1: 2 K
      Matern5 hp=0: sf2 (z + sqrt(z)^3) exp(-sqrt z)/3, z = 5 r^2/theta^2 (kernel.py:250-262); hp=1: 2 K
      Matern5-ARD (no reference class): the analytic derivative sf2 (5/3)(1+s) exp(-s) (dx_hp/theta_hp)^2, s = sqrt5 r.
    """
    X = np.array(X, dtype=float, ndmin=2)
    n, D = X.shape
    K = cov_matrix(kernel, X)
    kind = kernel.kind
    nlen = D if kind in (K_SE_ARD, K_MATERN5_ARD) else 1
    if hp == nlen and kind != K_SE_ISO and len(kernel.hyper) > nlen:
        return 2.0 * K
    if hp < 0 or hp >= nlen:
        raise ValueError("kernel has no hyperparameter %d" % hp)
    diff = X[:, None, :] - X[None, :, :]
    if kind == K_SE_ISO:
        return K * (np.sum(diff ** 2, axis=2) / kernel.hyper[0] ** 2)
    if kind == K_SE_ARD:
        return K * (diff[:, :, hp] ** 2 / kernel.theta[hp] ** 2)
    if kind == K_MATERN3:
        r = np.sqrt(np.sum(diff ** 2, axis=2))
        if exact_matern3:
            z = math.sqrt(3) * r / kernel.hyper[0]
            return kernel.sf2 * z ** 2 * np.exp(-z)
        return kernel.sf2 * r ** 2 * np.exp(-r)
    if kind == K_MATERN5:
        z = np.sum((math.sqrt(5.0) * diff / kernel.hyper[0]) ** 2.0, axis=2)
        return kernel.sf2 * (z + np.sqrt(z) ** 3.0) * np.exp(-np.sqrt(z)) / 3.0
    sc = diff / kernel.theta[None, None, :]
    s = math.sqrt(5.0) * np.sqrt(np.sum(sc ** 2, axis=2))
    return kernel.sf2 * (5.0 / 3.0) * (1.0 + s) * np.exp(-s) * sc[:, :, hp] ** 2


def marginal_likelihood(kernel, X, Y, nhyper, noise=1e-3, compute_gradient=True, exact_matern3=False):
    """trainhyper.marginalLikelihood with useCholesky=True (trainhyper.py:47-76):
    K = covMatrix(X) + noise I; nlml = Y.alpha/2 + sum log diag L + N log(2 pi)/2;
    dnlml[i] = sum((inv(K) - alpha alpha^T) o derivative(X, i)) / 2."""
    X = np.array(X, dtype=float, ndmin=2)
    Y = np.array(Y, dtype=float).reshape(-1)
    NX = len(X)
    K = cov_matrix(kernel, X) + np.eye(NX) * noise
    L = np.linalg.cholesky(K)
    alpha = np.linalg.solve(L.T, np.linalg.solve(L, Y))
    nlml = 0.5 * np.dot(Y, alpha) + np.sum(np.log(np.diag(L))) + 0.5 * NX * math.log(2.0 * math.pi)
    if not compute_gradient:
        return nlml
    W = np.linalg.solve(L.T, np.linalg.solve(L, np.eye(NX))) - np.outer(alpha, alpha)
    dnlml = np.array([np.sum(W * kernel_derivative(kernel, X, i, exact_matern3)) / 2.0 for i in range(nhyper)])
    return nlml, dnlml


# ---------------------------------------------------------------------------------------------
# PrefGaussianProcess: the MAP functional and its minimisation  (ego/gaussianprocess/__init__.py:355-386,441-442)
# ---------------------------------------------------------------------------------------------
def pref_S(x, prefinds, L):
    """S(x) = -sum_p (d+1) log(CDF((x[v]-x[u]) / (sqrt(2) sigma)) + 1e-10) + |inv(L) x|^2 / 2, sigma = 1 (:373-386);
    the reference solves with the general linalg.solve (:381)."""
    x = np.asarray(x, dtype=float)
    logCDFs = 0.
    Z = math.sqrt(2) * 1
    for v, u, d in prefinds:
        logCDFs += (d + 1) * math.log(float(cdf_py((x[v] - x[u]) / Z)) + 1e-10)
    Lx = np.linalg.solve(L, x)
    return -logCDFs + np.dot(Lx, Lx) / 2


def pref_fit_bfgs(start, prefinds, L):
    """the reference's optimiser call, verbatim in meaning: fmin_bfgs on numerical gradients (:442)"""
    from scipy.optimize import fmin_bfgs
    return np.asarray(fmin_bfgs(pref_S, np.asarray(start, dtype=float), args=(prefinds, L), disp=0), dtype=float)


def pref_S_grad(x, prefinds, L):
    """analytic gradient of pref_S (Gaussian pdf for dCDF/dz) -- used by the tests to certify a minimiser"""
    from scipy.linalg import solve_triangular
    x = np.asarray(x, dtype=float)
    g = np.zeros_like(x)
    for v, u, d in prefinds:
        z = (x[v] - x[u]) / math.sqrt(2)
        q = math.exp(-z * z / 2) / math.sqrt(2 * math.pi) / (float(cdf_py(z)) + 1e-10)
        gz = -(d + 1) * q / math.sqrt(2)
        g[v] += gz
        g[u] -= gz
    Lx = solve_triangular(L, x, lower=True)
    return g + solve_triangular(L, Lx, lower=True, trans='T')


def ei_py(mu, s2, ymax, xi):
    """EI.negf negated (ego/acquisition/__init__.py:150-164)."""
    ydiff = mu - ymax - xi
    s = np.sqrt(s2)
    Z = ydiff / s
    return ydiff * cdf_py(Z) + s * pdf_py(Z)


def pi_py(mu, s2, ymax, xi):
    """PI.negf negated (ego/acquisition/__init__.py:100-110)."""
    return cdf_py((mu - (ymax + xi)) / np.sqrt(s2))


def ucb_sbeta_py(nY, NA, delta=0.1):
    """UCB.__init__ (ego/acquisition/__init__.py:60-66); NA/2 is Python-2 integer division."""
    t = nY + 1
    return math.sqrt(2.0 * math.log(t ** (NA // 2 + 2) * math.pi ** 2 / (3.0 * delta)))


def ucb_py(mu, s2, sbeta, scale=0.2):
    """UCB.negf negated (ego/acquisition/__init__.py:68-71)."""
    return mu + math.sqrt(scale * sbeta) * np.sqrt(s2)


def ucb_parm_cpp(nY, NA, delta=0.1, scale=0.2):
    """cdirectGP's multiplier (ego/acquisition/__init__.py:317-319) -- differs from UCB.negf's."""
    t = nY + 1
    return math.sqrt(scale * 2.0 * math.log(t ** (NA // 2 + 2) * math.pi ** 2 / (3.0 * delta)))


def ei_cpp(mu, sigma, maxY, parm):
    """negei negated (cpp/optimizeGP.cpp:194-215)."""
    ydiff = mu - maxY - parm
    Z = ydiff / sigma
    return ydiff * cdf_cpp(Z) + sigma * pdf_cpp(Z)


def pi_cpp(mu, sigma, maxY, parm):
    """negpi negated (cpp/optimizeGP.cpp:217-227)."""
    return cdf_cpp((mu - maxY - parm) / sigma)


def ucb_cpp(mu, sigma, parm):
    """negucb negated (cpp/optimizeGP.cpp:229-236)."""
    return mu + parm * sigma


def score(acq, mode, mu, s2, ymax, parm):
    """Acquisition value (the quantity being *maximised*) from (mu, sigma^2)."""
    if mode == "py":
        if acq == ACQ_EI:
            return ei_py(mu, s2, ymax, parm)
        if acq == ACQ_PI:
            return pi_py(mu, s2, ymax, parm)
        return mu + parm * np.sqrt(s2)
    sig = np.sqrt(s2)
    if acq == ACQ_EI:
        return ei_cpp(mu, sig, ymax, parm)
    if acq == ACQ_PI:
        return pi_cpp(mu, sig, ymax, parm)
    return ucb_cpp(mu, sig, parm)


# ---------------------------------------------------------------------------------------------
# DIRECT, C++ semantics (cpp/direct.cpp:49-65,111-141,146-235,329-581)
# ---------------------------------------------------------------------------------------------
class _Rect(object):
    __slots__ = ("lb", "ub", "center", "d", "y")


def direct_cpp(f, lb, ub, maxiter, maxsample, record=None):
    """Pure-Python restatement of the reference's C++ DIRECT (small cases only).

    f(x) is called at points of the *original* box.  Returns (FMIN, XMIN, nsamples).
    `record`, if a list, receives every sampled point (original coordinates) in call order.
    std::sort in divrec is an insertion sort (stable) for <=16 elements, which Python's sort
    reproduces; for more than 16 equally-long sides tie order may differ from the reference.
    maxtime is not modelled (parity runs use an effectively infinite maxtime).
    """
    lowerb = np.array(lb, dtype=float)
    upperb = np.array(ub, dtype=float)
    N = len(lowerb)
    fixed = [lowerb[i] == upperb[i] for i in range(N)]             # :97-98
    st = {"FMIN": np.finfo(float).max, "XMIN": None, "ns": 0}

    def samplef(x):                                                # :111-141
        s = np.empty(N)
        for i in range(N):
            s[i] = lowerb[i] if fixed[i] else x[i] * (upperb[i] - lowerb[i]) + lowerb[i]
        if record is not None:
            record.append(s.copy())
        y = float(f(s))
        st["ns"] += 1
        if y < st["FMIN"]:
            st["FMIN"] = y
            st["XMIN"] = np.array([lowerb[i] + (upperb[i] - lowerb[i]) * x[i] for i in range(N)])
        return y

    def mkrect(l, u):                                              # :49-65
        r = _Rect()
        r.lb = list(l)
        r.ub = list(u)
        r.center = [0.0] * N
        d = 0.0
        for i in range(N):
            r.center[i] = r.lb[i] + (r.ub[i] - r.lb[i]) / 2.
            d += (r.lb[i] - r.center[i]) ** 2
        r.d = math.sqrt(d)
        r.y = samplef(r.center)
        return r

    def divrec(rec):                                               # :146-235
        maxlength = rec.ub[0] - rec.lb[0]
        for i in range(1, N):
            if not fixed[i] and rec.ub[i] - rec.lb[i] > maxlength:
                maxlength = rec.ub[i] - rec.lb[i]
        I = []
        for i in range(N):
            if not fixed[i] and rec.ub[i] - rec.lb[i] == maxlength:
                s1 = list(rec.center)
                s2 = list(rec.center)
                s1[i] = rec.lb[i] + maxlength / 3.
                s2[i] = rec.lb[i] + 2. * maxlength / 3.
                sf1 = samplef(s1)
                sf2 = samplef(s2)
                I.append((i, sf1 if sf1 < sf2 else sf2))
        I.sort(key=lambda t: t[1])
        old = _Rect()
        old.lb = list(rec.lb); old.ub = list(rec.ub); old.center = list(rec.center); old.y = rec.y; old.d = rec.d
        new = []
        for dd, _ in I:
            dwidth = old.ub[dd] - old.lb[dd]
            split1 = old.lb[dd] + dwidth / 3.
            split2 = old.lb[dd] + 2. * dwidth / 3.
            ub1 = list(old.ub); ub1[dd] = split1
            lb3 = list(old.lb); lb3[dd] = split2
            ub3 = list(old.ub)
            new.append(mkrect(old.lb, ub1))
            old.lb[dd] = split1
            old.ub[dd] = split2
            new.append(mkrect(lb3, ub3))
        d = 0.0
        for i in range(N):
            d += (old.lb[i] - old.center[i]) ** 2
        old.d = math.sqrt(d)
        new.append(old)
        return new

    first = mkrect([0.0] * N, [1.0] * N)
    recs = divrec(first)
    eps = 10e-10
    MIN_D = np.finfo(float).tiny
    MAX_D = np.finfo(float).max
    it = 0
    done = False
    while it < maxiter and not done:
        it += 1
        potopts = []
        for j, Rj in enumerate(recs):                              # :378-456
            maxI1 = MIN_D
            minI2 = MAX_D
            breaked = False
            for i, Ri in enumerate(recs):
                if i == j:
                    continue
                if Ri.d < Rj.d:
                    val = (Rj.y - Ri.y) / (Rj.d - Ri.d)
                    if val > maxI1:
                        maxI1 = val
                elif Ri.d > Rj.d:
                    val = (Ri.y - Rj.y) / (Ri.d - Rj.d)
                    if val < minI2:
                        minI2 = val
                        if minI2 <= 0.:
                            breaked = True
                            break
                else:
                    if Rj.y > Ri.y:
                        breaked = True
                        break
                if maxI1 != MIN_D and minI2 != MAX_D and minI2 < maxI1:
                    breaked = True
                    break
            if not breaked:
                F = st["FMIN"]
                if minI2 == MAX_D:
                    potopts.append(j)
                elif F == 0.0:
                    if Rj.y <= Rj.d * minI2:
                        potopts.append(j)
                elif eps <= (F - Rj.y) / abs(F) + (Rj.d / abs(F)) * minI2:
                    potopts.append(j)
        if not potopts:
            break
        for j in reversed(potopts):                                # :479-498
            new = divrec(recs[j])
            recs.extend(new)
            del recs[j]
            if st["ns"] > maxsample:
                done = True
                break
        if st["ns"] > maxsample:
            break
    return st["FMIN"], st["XMIN"], st["ns"]


# ---------------------------------------------------------------------------------------------
# maximizeEI / cdirectGP and fastUCBGallery
#   (ego/acquisition/__init__.py:174-197,307-447; ego/acquisition/gallery.py:42-135)
# ---------------------------------------------------------------------------------------------
def acqmax_cpp(gp, bounds, acq, parm, maxiter=50, maxsample=10000):
    """cdirectGP + acqmaxGP restated: inv(R) is formed once (ego/acquisition/__init__.py:385-388), DIRECT
    (cpp/direct.cpp) minimises the libego-arithmetic negative acquisition (cpp/optimizeGP.cpp:194-236), and the
    result comes back as (-min, argmin) (:443-447).  maxtime is not modelled."""
    b = np.array(bounds, dtype=float)
    invR = gp.invR()
    ymax = gp.Y.max()

    def f(x):
        mu, sig = gp.posterior_cpp(x[None, :], invR=invR)
        return -float(score(acq, "cpp", mu, sig ** 2, ymax, parm)[0])
    fmin, xmin, ns = direct_cpp(f, b[:, 0], b[:, 1], maxiter, maxsample)
    return -fmin, np.array(xmin, dtype=float)


def fast_ucb_gallery(kernel, X, Y, bounds, N, prior=None, use_best=True, samples=300, seed=None,
                     maxiter=50, maxsample=10000):
    """ego/acquisition/gallery.py:42-135 for a model that has data (the no-data branches, :68-90, need BFGS on the
    prior mean and are host logic only).  hallucGP is a plain GP with the default noise 0.1 whatever the source model
    was (:67); per slot: maximizeEI(xi=.3) through libego's DIRECT (:101), the 0.5 distance rule (:102-105), `samples`
    latin-hypercube points scored with the Python-arithmetic EI(xi=.4) (:98,111-116), the prior means (:119-130),
    then hallucGP.addData(best, hallucGP.mu(best)) (:134).  `seed` pins the reference's unseeded lhcSample the same way
    the product's `seed` extension does (seed + slot index)."""
    gallery = []
    X = np.array(X, dtype=float, ndmin=2)
    Y = np.array(Y, dtype=float).reshape(-1)
    if use_best:                                                                # :50-63
        bestY, bestX = -np.inf, None
        for x, y in zip(X, Y):
            if y > bestY and all(b[0] <= v <= b[1] for v, b in zip(x, bounds)):
                bestY, bestX = y, x
        if bestX is not None:
            gallery.append(bestX)
    halluc = GPOracle(kernel, X, Y, noise=0.1, prior=prior)                     # :67
    slot = 0
    while len(gallery) < N:                                                     # :93
        bestU, bestX = -np.inf, None
        opt, optx = acqmax_cpp(halluc, bounds, ACQ_EI, .3, maxiter, maxsample)  # :101
        if len(gallery) == 0 or min(np.linalg.norm(optx - gx) for gx in gallery) > .5:
            bestU, bestX = opt, optx
        cand = np.array(lhc_sample(bounds, samples, seed=None if seed is None else seed + slot))
        mu, s2 = halluc.posterior_batch(cand)
        u = ei_py(mu, s2, halluc.Y.max(), .4)                                   # :98 ut = EI(hallucGP, xi=.4)
        for x, ux in zip(cand, u):                                              # :111-116
            if ux > bestU and min(np.linalg.norm(x - gx) for gx in gallery) > .5:
                bestU, bestX = ux, x
        if prior is not None:                                                   # :119-130
            pm = np.array([[np.clip(x[i], bounds[i][0], bounds[i][1]) for i in range(len(x))] for x in prior.means])
            pm = pm * prior.width + prior.lowerb
            mu, s2 = halluc.posterior_batch(pm)
            for x, ux in zip(pm, ei_py(mu, s2, halluc.Y.max(), .4)):
                if ux > bestU and (len(gallery) == 0 or min(np.linalg.norm(x - gx) for gx in gallery) > .5):
                    bestU, bestX = ux, x
        gallery.append(bestX)
        mub, _ = halluc.posterior_batch(bestX[None, :])                         # :134
        halluc.add_data(bestX, mub[0])
        slot += 1
    return gallery


# ---------------------------------------------------------------------------------------------
# test functions used as fixture data (ego/utils/testfunctions.py:171-182,244-250,289-304)
# ---------------------------------------------------------------------------------------------
def branin(x):
    """testfunctions.py:244-250 (the Branin class' f; maximised as -f/100 in fixtures)."""
    x = np.asarray(x, dtype=float)
    a = x[..., 1] - (5.1 / (4 * np.pi ** 2)) * x[..., 0] ** 2 + 5 * x[..., 0] / np.pi - 6
    return a ** 2 + 10 * (1 - 1 / (8 * np.pi)) * np.cos(x[..., 0]) + 10


_SHEKEL_A = np.array([[4., 4., 4., 4.], [1., 1., 1., 1.], [8., 8., 8., 8.], [6., 6., 6., 6.], [3., 7., 3., 7.]])
_SHEKEL_C = np.array([.1, .2, .2, .4, .4])


def shekel5(x):
    """Shekel m=5 (testfunctions.py:171-182): f(x) = -sum 1/(||x-a_i||^2 + c_i); min -10.1532 at (4,4,4,4)."""
    x = np.asarray(x, dtype=float)
    return -np.sum(1.0 / (np.sum((x[None, :] - _SHEKEL_A) ** 2, axis=1) + _SHEKEL_C))


_H6_A = np.array([[10, 3, 17, 3.5, 1.7, 8], [.05, 10, 17, .1, 8, 14], [3, 3.5, 1.7, 10, 17, 8], [17, 8, .05, 10, .1, 14]])
_H6_C = np.array([1, 1.2, 3, 3.2])
_H6_P = np.array([[.1312, .1696, .5569, .0124, .8283, .5886], [.2329, .4135, .8307, .3736, .1004, .9991],
                  [.2348, .1451, .3522, .2883, .3047, .6650], [.4047, .8828, .8732, .5743, .1091, .0381]])


def hartman6_neg(X):
    """Negated Hartman-6 (testfunctions.py:289-304), i.e. the maximisation form used as Y."""
    X = np.array(X, dtype=float, ndmin=2)
    e = np.sum(_H6_A[None, :, :] * (X[:, None, :] - _H6_P[None, :, :]) ** 2, axis=2)
    return np.sum(_H6_C[None, :] * np.exp(-e), axis=1)


def lhc_sample(bounds, N, seed=None):
    """ego/utils/latinhypercube.py:27-46 restated (frozen numpy RandomState stream)."""
    rs = np.random.RandomState(seed)
    samp = []
    for bmin, bmax in bounds:
        if bmin == bmax:
            dsamp = np.array([bmin] * N, dtype=float)
        else:
            dsamp = (bmax - bmin) * rs.rand(N) / N + np.arange(bmin, bmax, (bmax - bmin) / N)
        rs.shuffle(dsamp)
        samp.append(dsamp)
    return list(np.vstack(samp).T)
