"""
TEST INFRASTRUCTURE ONLY -- NumPy model of the integer arithmetic of the INT8 tensor-core path of wide batches
(ibo_b200/csrc/score_i8.cuh).  Nothing in the product path imports this module.

There is no reference counterpart: the reference computes v = L^-1 k* in FP64 (ego/gaussianprocess/__init__.py:209-223,
cpp/optimizeGP.cpp:57-191).  This module restates, digit for digit, how the device replaces the FP64 product V = W K* by exact
integer products, so that the scheme's error is pinned on the CPU (tests/test_int8_model.py) independently of the GPU:

  W  (row r) = 2^e_r sum_{t=1..7} 2^(-8t) A_t       A_t: balanced base-256 digits of rint(w 2^(56 - e_r)), |w| 2^-e_r < 1/4   (i8_slice_w_kernel)
  K*         = 1/2 + 2 sum_{u=1..7} 2^(-8u) B_u     B_u: balanced base-256 digits of rint((k - 1/2) 2^55)                      (kstar_i8_kernel)
  V          = 2 * 2^e_r [2^-32 (D_2 2^16 + D_3 2^8 + D_4) + 2^-56 (D_5 2^16 + D_6 2^8 + D_7) + 2^-64 D_8] + 1/2 sum_k W[r, k]
               D_g = sum_{t+u=g} A_t B_u^T  (trigemm_i8_kernel: INT32 accumulators, pairs with t + u > 8 dropped)

`pairs(ndig_w, ndig_k, gmax)` evaluates the same construction for other digit counts / truncations in extended precision: the
study behind the choice of 7 x 7 digits with t + u <= 8 (28 products).
"""
import numpy as np

S = 7
BITS = 8


def row_scale_exponent(W):
    """e_r with max_k |W[r, k]| / 2^e_r in [1/8, 1/4)  (i8_rowscale_kernel: ilogb(max) + 3)"""
    mx = np.max(np.abs(W), axis=1)
    e = np.zeros(len(mx))
    nz = mx > 0
    e[nz] = np.floor(np.log2(mx[nz])) + 3
    return e


def balanced256(q, ndig=S):
    """balanced base-256 digits of the integers q, most significant first: d_t in [-128, 127] for t >= 2, the rest is d_1"""
    q = q.copy()
    out = [None] * ndig
    for t in range(ndig, 1, -1):
        dgt = ((q + 128) & 255) - 128
        q = (q - dgt) >> 8
        out[t - 1] = dgt
    out[0] = q
    return out


def w_digits(W, e, ndig=S):
    return balanced256(np.rint(W / 2.0 ** e[:, None] * 2.0 ** (BITS * ndig)).astype(np.int64), ndig)


def k_digits(K, ndig=S):
    """digits of (k - 1/2) / 2 in [-1/4, 1/4] at 8 ndig fractional bits"""
    return balanced256(np.rint((K - 0.5) * 2.0 ** (BITS * ndig - 1)).astype(np.int64), ndig)


def group_sums(A, B, gmax=S + 1):
    """D_g for g = 2 .. gmax (exact integers; asserted to fit the device's INT32 accumulators)"""
    D = []
    for g in range(2, gmax + 1):
        acc = np.zeros((A[0].shape[0], B[0].shape[1]), dtype=np.int64)
        for t in range(1, len(A) + 1):
            u = g - t
            if 1 <= u <= len(B):
                acc += A[t - 1] @ B[u - 1]
        assert np.max(np.abs(acc)) < 2 ** 31
        D.append(acc)
    return D


def assemble(D, e, W):
    """the epilogue's FP64 assembly: two exact INT64 partial sums, two FMAs, one FMA with the row's scale and shift constant"""
    hi = D[0] * 65536 + D[1] * 256 + D[2]
    mid = D[3] * 65536 + D[4] * 256 + D[5]
    lo = D[6].astype(float) * 2.0 ** -64
    v = hi.astype(float) * 2.0 ** -32 + (mid.astype(float) * 2.0 ** -56 + lo)
    return v * (2.0 * 2.0 ** e[:, None]) + 0.5 * np.sum(W, axis=1)[:, None]


def emulated_product(W, K):
    """V = W K through the device's integer scheme (W: rows x k, lower triangular or not; K: k x candidates, entries in [0, 1])"""
    e = row_scale_exponent(W)
    A, B = w_digits(W, e), k_digits(K)
    assert np.max(np.abs(A[0])) <= 64 and np.max(np.abs(B[0])) <= 64
    return assemble(group_sums(A, B), e, W)


def pairs(W, K, ndig_w, ndig_k, gmax):
    """the same construction with ndig_w / ndig_k digits and the pairs t + u <= gmax, assembled in extended precision;
    returns (V, number of digit products)"""
    e = row_scale_exponent(W)
    A, B = w_digits(W, e, ndig_w), k_digits(K, ndig_k)
    v = np.zeros((W.shape[0], K.shape[1]), dtype=np.longdouble)
    n = 0
    for t in range(1, ndig_w + 1):
        for u in range(1, ndig_k + 1):
            if t + u <= gmax:
                # digit t of W carries 2^-(8 t) of the row scale, digit u of K* 2^-(8 u) of 2
                v += (A[t - 1] @ B[u - 1]).astype(np.longdouble) * np.longdouble(2.0) ** (-8 * (t + u))
                n += 1
    v = v * (2.0 * 2.0 ** e[:, None]).astype(np.longdouble) + (0.5 * np.sum(W.astype(np.longdouble), axis=1))[:, None]
    return v, n
