"""CPU tests of the integer scheme behind the INT8 tensor-core path of wide batches (oracle/int8_model.py restates
ibo_b200/csrc/score_i8.cuh digit for digit): digit ranges, exact reconstruction, INT32 head-room at the largest supported N, and
the error of sigma^2 / EI against the FP64 evaluation the reference performs (ego/gaussianprocess/__init__.py:209-224) and
against an extended-precision evaluation of the same product."""
import numpy as np
from scipy.linalg import solve_triangular

from oracle import ibo_oracle as orc
from oracle import int8_model as i8


def test_digits_reconstruct_the_operands():
    rs = np.random.RandomState(0)
    W = np.tril(rs.randn(64, 64) * np.exp(rs.randn(64, 1) * 3))
    e = i8.row_scale_exponent(W)
    assert np.all(np.max(np.abs(W), axis=1) / 2.0 ** e < 0.25) and np.all(np.max(np.abs(W), axis=1) / 2.0 ** e >= 0.125)
    A = i8.w_digits(W, e)
    assert all(np.all((a >= -128) & (a <= 127)) for a in A[1:]) and np.all(np.abs(A[0]) <= 64)
    rec = sum(A[t - 1] * 2.0 ** (-8 * t) for t in range(1, 8)) * 2.0 ** e[:, None]
    assert np.max(np.abs(rec - W) / 2.0 ** e[:, None]) <= 2.0 ** -57          # half a unit of the last digit
    K = np.r_[rs.rand(1000), [0.0, 1.0, 1 - 2.0 ** -53, 2.0 ** -60]].reshape(-1, 1)
    B = i8.k_digits(K)
    assert all(np.all((b >= -128) & (b <= 127)) for b in B[1:]) and np.all(np.abs(B[0]) <= 64)
    rec = 0.5 + 2.0 * sum(B[u - 1] * 2.0 ** (-8 * u) for u in range(1, 8))
    assert np.max(np.abs(rec - K)) <= 2.0 ** -56


def test_int32_headroom_at_the_largest_model():
    """worst case |D_g| <= (2 x 64 x 128 + 5 x 128 x 128) x N for the seven pairs of group 8: the accumulators are INT32 on the
    device, N <= 16384 supported (use_i8 in score.cu)"""
    N = 16384
    assert (2 * 64 * 128 + 5 * 128 * 128) * N < 2 ** 31
    A = [np.full((2, N), -64, dtype=np.int64)] + [np.full((2, N), -128, dtype=np.int64)] * 6
    B = [np.full((N, 2), -64, dtype=np.int64)] + [np.full((N, 2), -128, dtype=np.int64)] * 6
    D = i8.group_sums(A, B)                 # asserts < 2^31 inside
    assert max(int(np.max(np.abs(d))) for d in D) == (2 * 64 * 128 + 5 * 128 * 128) * N


def _case(N=768, d=6, M=192):
    rs = np.random.RandomState(0)
    X = rs.rand(N, d)
    Y = orc.hartman6_neg(X)
    gp = orc.GPOracle(orc.KernelSpec(orc.K_SE_ARD, [.53, .57, 2.5, .34, .27, .35], d), X, Y, 0.1)
    Xs = np.random.RandomState(1).rand(M, d)
    Ks = gp.kernel.cross(gp.X, Xs)
    W = solve_triangular(gp.L, np.eye(N), lower=True)
    return gp, Y, Ks, W


def test_emulated_sigma2_and_ei_stay_inside_the_parity_bound():
    gp, Y, Ks, W = _case()
    V0 = W @ Ks
    s20 = np.clip(1.1 - np.sum(V0 * V0, axis=0), 1e-8, 10)
    mu = Ks.T @ (W.T @ (W @ Y))             # the device takes mu as k* . alpha in plain FP64
    mu0 = V0.T @ (W @ Y)
    assert np.max(np.abs(mu - mu0) / np.maximum(np.abs(mu0), 1e-3)) <= 1e-11
    ei0 = orc.score(orc.ACQ_EI, "cpp", mu0, s20, Y.max(), 0.01)
    V = i8.emulated_product(W, Ks)
    s2 = np.clip(1.1 - np.sum(V * V, axis=0), 1e-8, 10)
    ei = orc.score(orc.ACQ_EI, "cpp", mu0, s2, Y.max(), 0.01)
    assert np.max(np.abs(s2 - s20) / s20) <= 2e-13
    # (EI near its 1e-5 floor: Phi = (1 + erf) / 2 carries an absolute rounding noise of 1e-16, times |mu - ymax|)
    assert np.max(np.abs(ei - ei0) / np.maximum(np.abs(ei0), 1e-5)) <= 5e-11
    assert int(np.argmax(ei)) == int(np.argmax(ei0))


def test_scheme_error_equals_the_fp64_gemm_error_and_28_products_are_needed():
    """against sum_r V_r^2 of the same W and K* in extended precision: the shipped scheme (7 x 7 digits, t + u <= 8) is as exact
    as the FP64 GEMM it replaces; dropping group 8 (21 products) or the seventh W digit costs two orders of magnitude, the
    seventh K* digit does not matter much (kept for small-noise models, where k* errors are amplified by inv(R))."""
    gp, Y, Ks, W = _case(N=512, M=96)
    Wl, Kl = W.astype(np.longdouble), Ks.astype(np.longdouble)
    V0 = Wl @ Kl
    q0 = np.sum(V0 * V0, axis=0)

    def err(V):
        Vl = V.astype(np.longdouble)
        return float(np.max(np.abs(np.sum(Vl * Vl, axis=0) - q0)))

    e_fp64 = err(W @ Ks)
    e_dev = err(i8.emulated_product(W, Ks))
    assert e_fp64 < 5e-15 and e_dev < 1e-14
    v, n = i8.pairs(W, Ks, 7, 7, 8)
    assert n == 28 and err(v) < 1e-14
    v, n = i8.pairs(W, Ks, 7, 7, 7)
    assert n == 21 and 20 * e_dev < err(v) < 1e-11
    v, n = i8.pairs(W, Ks, 6, 7, 8)
    assert n == 27 and err(v) > 10 * e_dev
    v, n = i8.pairs(W, Ks, 7, 6, 8)
    assert n == 27 and err(v) < 3e-14
