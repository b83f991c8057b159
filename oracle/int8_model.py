"""
TEST INFRASTRUCTURE ONLY -- NumPy model of the integer arithmetic of the experimental INT8-emulated K2
(ibo_b200/csrc/score_i8.cuh; IBO_FLAG_INT8).  Nothing in the product path imports this module.

There is no reference counterpart: the reference computes v = L^-1 k* in FP64 (ego/gaussianprocess/__init__.py:209-223,
cpp/optimizeGP.cpp:57-191).  This module restates, digit for digit, how the device replaces the FP64 product V = W K* by exact
integer products, so that the scheme's error bound is pinned on the CPU (tests/test_int8_model.py) independently of the GPU:

  W  (row r) = 2^e_r sum_{t=1..7} 2^(-7t) A_t    A_t: balanced digits of rint(w 2^(49 - e_r)), |w| 2^-e_r < 1/2   (i8_slice_w_kernel)
  K*         =       sum_{u=1..7} 2^(-7u) B_u    B_u: unsigned digits of min(rint(k 2^49), 2^49 - 1), k in [0, 1]  (kstar_i8_kernel)
  V          = 2^e_r [ 2^-28 (D_2 2^14 + D_3 2^7 + D_4) + 2^-49 (D_5 2^14 + D_6 2^7 + D_7) + 2^-56 D_8 ],  D_g = sum_{t+u=g} A_t B_u^T
               (trigemm_i8_kernel: INT32 accumulators, pairs with t + u > 8 dropped; `groups=8` keeps t + u = 9 as well)
"""
import numpy as np

S = 7
FRAC = 7 * S


def row_scale_exponent(W):
    """e_r with max_k |W[r, k]| / 2^e_r in [1/4, 1/2)  (i8_rowscale_kernel: ilogb(max) + 2)"""
    mx = np.max(np.abs(W), axis=1)
    e = np.zeros(len(mx))
    nz = mx > 0
    e[nz] = np.floor(np.log2(mx[nz])) + 2
    return e


def w_digits(W, e):
    """balanced digits, most significant first: A[t-1], t = 1..7; d_t in [-64, 63] for t >= 2, |d_1| <= 64"""
    q = np.rint(W / 2.0 ** e[:, None] * 2.0 ** FRAC).astype(np.int64)
    out = [None] * S
    for t in range(S, 1, -1):
        dgt = ((q + 64) & 127) - 64
        q = (q - dgt) >> 7
        out[t - 1] = dgt
    out[0] = q
    return out


def k_digits(K):
    """unsigned digits of min(rint(k 2^49), 2^49 - 1), most significant first"""
    q = np.minimum(np.rint(K * 2.0 ** FRAC).astype(np.int64), 2 ** FRAC - 1)
    return [(q >> (FRAC - 7 * t)) & 127 for t in range(1, S + 1)]


def group_sums(A, B, groups=7):
    """D_g for g = 2 .. groups + 1 (exact integers; asserted to fit the device's INT32 accumulators)"""
    D = []
    for g in range(2, groups + 2):
        acc = np.zeros((A[0].shape[0], B[0].shape[1]), dtype=np.int64)
        for t in range(1, S + 1):
            u = g - t
            if 1 <= u <= S:
                acc += A[t - 1] @ B[u - 1]
        assert np.max(np.abs(acc)) < 2 ** 31
        D.append(acc)
    return D


def assemble(D, e):
    """the epilogue's FP64 assembly: three exact INT64 partial sums, two FMAs, one scaling by the row's power of two"""
    hi = D[0] * 16384 + D[1] * 128 + D[2]
    mid = D[3] * 16384 + D[4] * 128 + D[5]
    lo = D[6].astype(float) * 2.0 ** -56 if len(D) == 7 else (D[6] * 128 + D[7]).astype(float) * 2.0 ** -63
    v = hi.astype(float) * 2.0 ** -28 + (mid.astype(float) * 2.0 ** -49 + lo)
    return v * 2.0 ** e[:, None]


def emulated_product(W, K, groups=7):
    """V = W K through the device's integer scheme (W: rows x k, lower triangular or not; K: k x candidates, entries in [0, 1])"""
    e = row_scale_exponent(W)
    return assemble(group_sums(w_digits(W, e), k_digits(K), groups), e)


# ---------------------------------------------------------------------------------------------
# 8-bit digits (IBO_FLAG_INT8_D8: 7 digits, same 28 products, operands rounded at 2^-56; IBO_FLAG_INT8_S6: 6 digits, 21 products)
#   W  (row r) = 2^e_r sum_t 2^(-8t) A_t          A_t: balanced base-256 digits of rint(w 2^(56 - e_r)), |w| 2^-e_r < 1/4
#   K*         = 1/2 + 2 sum_u 2^(-8u) B_u        B_u: balanced base-256 digits of rint((k - 1/2) 2^55)   (signed: (k - 1/2)/2 in [-1/4, 1/4])
#   V          = 2 * 2^e_r [2^-32 (D_2 2^16 + D_3 2^8 + D_4) + 2^-56 (D_5 2^16 + D_6 2^8 + D_7) + 2^-64 D_8] + 1/2 sum_k W[r, k]
# ---------------------------------------------------------------------------------------------
def _balanced256(q, ndig):
    out = [None] * ndig
    for t in range(ndig, 1, -1):
        dgt = ((q + 128) & 255) - 128
        q = (q - dgt) >> 8
        out[t - 1] = dgt
    out[0] = q
    return out


def emulated_product_d8(W, K, ndig=7):
    """ndig = 7: IBO_FLAG_INT8_D8 (28 pairs, t + u <= 8); ndig = 6: IBO_FLAG_INT8_S6 (21 pairs, t + u <= 7, operands at 2^-48)"""
    mx = np.max(np.abs(W), axis=1)
    e = np.zeros(len(mx))
    e[mx > 0] = np.floor(np.log2(mx[mx > 0])) + 3
    A = _balanced256(np.rint(W / 2.0 ** e[:, None] * 2.0 ** (8 * ndig)).astype(np.int64), ndig)
    B = _balanced256(np.rint((K - 0.5) * 2.0 ** (8 * ndig - 1)).astype(np.int64), ndig)
    assert np.max(np.abs(A[0])) <= 64 and np.max(np.abs(B[0])) <= 64
    D = []
    for g in range(2, ndig + 2):
        acc = np.zeros((W.shape[0], K.shape[1]), dtype=np.int64)
        for t in range(1, ndig + 1):
            u = g - t
            if 1 <= u <= ndig:
                acc += A[t - 1] @ B[u - 1]
        assert np.max(np.abs(acc)) < 2 ** 31
        D.append(acc)
    hi = D[0] * 65536 + D[1] * 256 + D[2]
    mid = D[3] * 65536 + D[4] * 256 + D[5]
    lo = D[6].astype(float) * 2.0 ** -64 if ndig == 7 else 0.0
    v = hi.astype(float) * 2.0 ** -32 + (mid.astype(float) * 2.0 ** -56 + lo)
    return v * (2.0 * 2.0 ** e[:, None]) + 0.5 * np.sum(W, axis=1)[:, None]
