"""CPU tests of the integer scheme behind IBO_FLAG_INT8 (oracle/int8_model.py restates ibo_b200/csrc/score_i8.cuh digit for
digit): digit ranges, exact reconstruction, INT32 head-room at the largest supported N, and the error of sigma^2 / EI against
the FP64 evaluation the reference performs (ego/gaussianprocess/__init__.py:209-224)."""
import numpy as np
from scipy.linalg import solve_triangular

from oracle import ibo_oracle as orc
from oracle import int8_model as i8


def test_digits_reconstruct_the_operands():
    rs = np.random.RandomState(0)
    W = np.tril(rs.randn(64, 64) * np.exp(rs.randn(64, 1) * 3))
    e = i8.row_scale_exponent(W)
    assert np.all(np.max(np.abs(W), axis=1) / 2.0 ** e < 0.5) and np.all(np.max(np.abs(W), axis=1) / 2.0 ** e >= 0.25)
    A = i8.w_digits(W, e)
    assert all(np.all((a >= -64) & (a <= 63)) for a in A[1:]) and np.all(np.abs(A[0]) <= 64)
    rec = sum(A[t - 1] * 2.0 ** (-7 * t) for t in range(1, 8)) * 2.0 ** e[:, None]
    assert np.max(np.abs(rec - W) / 2.0 ** e[:, None]) <= 2.0 ** -50          # half a unit of the last digit
    K = np.r_[rs.rand(1000), [0.0, 1.0, 1 - 2.0 ** -53, 2.0 ** -60]].reshape(-1, 1)
    B = i8.k_digits(K)
    assert all(np.all((b >= 0) & (b <= 127)) for b in B)
    rec = sum(B[u - 1] * 2.0 ** (-7 * u) for u in range(1, 8))
    assert np.max(np.abs(rec - K)) <= 2.0 ** -49                               # k = 1 is clamped one unit below 2^49


def test_int32_headroom_at_the_largest_model():
    """worst case |D_g| <= 7 pairs x N x 64 x 127: the accumulators are INT32 on the device (N <= 32768 supported)"""
    assert 7 * 32768 * 64 * 127 < 2 ** 31
    rs = np.random.RandomState(1)
    A = [np.full((4, 8192), 64, dtype=np.int64)] + [np.full((4, 8192), 63, dtype=np.int64)] * 6
    B = [np.full((8192, 4), 127, dtype=np.int64)] * 7
    D = i8.group_sums(A, B)                 # asserts < 2^31 inside
    assert max(int(np.max(np.abs(d))) for d in D) == (64 + 6 * 63) * 127 * 8192      # group 8: all seven pairs


def test_emulated_sigma2_and_ei_stay_inside_the_parity_bound():
    N, d, M = 768, 6, 192
    rs = np.random.RandomState(0)
    X = rs.rand(N, d)
    Y = orc.hartman6_neg(X)
    gp = orc.GPOracle(orc.KernelSpec(orc.K_SE_ARD, [.53, .57, 2.5, .34, .27, .35], d), X, Y, 0.1)
    Xs = np.random.RandomState(1).rand(M, d)
    Ks = gp.kernel.cross(gp.X, Xs)
    W = solve_triangular(gp.L, np.eye(N), lower=True)
    V0 = W @ Ks
    s20 = np.clip(1.1 - np.sum(V0 * V0, axis=0), 1e-8, 10)
    mu = Ks.T @ (W.T @ (W @ Y))             # the device takes mu as k* . alpha in plain FP64
    mu0 = V0.T @ (W @ Y)
    assert np.max(np.abs(mu - mu0) / np.maximum(np.abs(mu0), 1e-3)) <= 1e-11
    ei0 = orc.score(orc.ACQ_EI, "cpp", mu0, s20, Y.max(), 0.01)
    errs = {}
    for groups in (7, 8):
        V = i8.emulated_product(W, Ks, groups)
        s2 = np.clip(1.1 - np.sum(V * V, axis=0), 1e-8, 10)
        ei = orc.score(orc.ACQ_EI, "cpp", mu0, s2, Y.max(), 0.01)
        errs[groups] = (np.max(np.abs(s2 - s20) / s20), np.max(np.abs(ei - ei0) / np.maximum(np.abs(ei0), 1e-5)))
        assert int(np.argmax(ei)) == int(np.argmax(ei0))
    V8 = i8.emulated_product_d8(W, Ks)                    # 8-bit digits (IBO_FLAG_INT8_D8): same 28 products
    s28 = np.clip(1.1 - np.sum(V8 * V8, axis=0), 1e-8, 10)
    assert np.max(np.abs(s28 - s20) / s20) <= 2e-13 and np.max(np.abs(s28 - s20) / s20) <= 0.05 * errs[7][0]
    V6 = i8.emulated_product_d8(W, Ks, ndig=6)            # six 8-bit digits (IBO_FLAG_INT8_S6): 21 products, today's accuracy
    s26 = np.clip(1.1 - np.sum(V6 * V6, axis=0), 1e-8, 10)
    assert np.max(np.abs(s26 - s20) / s20) <= 1e-11
    assert errs[7][0] <= 1e-11 and errs[7][1] <= 1e-10
    assert errs[8][0] <= 0.2 * errs[7][0]                # the eighth group buys a decimal digit (then the 2^-49 rounding of the operands dominates)
