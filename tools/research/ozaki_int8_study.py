"""Accuracy study (CPU, NumPy; research note for the next round, not part of the product or of the tests):
can K2's FP64 triangular GEMM V = W K* be emulated on the B200's INT8 tensor cores (tcgen05 kind::i8, 4.5 POP/s dense
against 37 TF/s of DMMA) within the 1e-10 parity bound on mu, sigma^2 and EI?

Ozaki-style slicing: each row of W and each column of K* is scaled by a power of two and cut into `s` signed 7-bit
slices, W ~ 2^eW sum_t 2^(-7t) A_t, K* ~ 2^eK sum_u 2^(-7u) B_u with int8 A_t, B_u.  The slice products A_t B_u are
exact integer GEMMs (int32 accumulation: |sum| <= N 2^14 < 2^31 for N <= 2^17 / s when the pairs of equal t+u share
an accumulator); pairs with t + u > s + 1 are dropped.  V = 2^(eW+eK) sum_{t+u <= s+1} 2^(-7(t+u)) A_t B_u is then
assembled in FP64 in the epilogue, where the three row reductions of K2 stay as they are.
Cost per candidate: s(s+1)/2 triangular int8 GEMM passes of N^2/2 MACs each, against N^2/2 FP64 FMAs.
"""
import sys
import os
import numpy as np
from scipy.linalg import solve_triangular

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import ibo_oracle as orc


def slices(Ascaled, s):
    """Ascaled in (-1, 1) -> list of s integer arrays in [-127, 127] with Ascaled ~ sum_t 2^(-7t) q_t"""
    out = []
    r = Ascaled.copy()
    for _ in range(s):
        r = r * 128.0
        q = np.trunc(r)
        r = r - q
        out.append(q.astype(np.int64))
    return out


def balanced_slices(Ascaled, s):
    """|Ascaled| < 1/2 -> balanced digits (least significant first), as ibo_b200/csrc/score_i8.cuh slices W"""
    q = np.rint(Ascaled * 2.0 ** (7 * s)).astype(np.int64)
    out = [None] * s
    for t in range(s, 1, -1):
        d = ((q + 64) & 127) - 64
        q = (q - d) >> 7
        out[t - 1] = d
    out[0] = q
    assert np.max(np.abs(q)) <= 64
    return out


def rounded_slices(Ascaled, s):
    """Ascaled in [0, 1] -> unsigned digits of rint(a 2^(7s)) (clamped below 2^(7s))"""
    q = np.minimum(np.rint(Ascaled * 2.0 ** (7 * s)).astype(np.int64), 2 ** (7 * s) - 1)
    return [(q >> (7 * (s - t))) & 127 for t in range(1, s + 1)]


def emulate(W, K, s, balanced=False):
    eW = np.ceil(np.log2(np.max(np.abs(W), axis=1) * (1 + 1e-12)))      # per row of W
    eK = np.ceil(np.log2(np.max(np.abs(K), axis=0) * (1 + 1e-12)))      # per candidate column
    if balanced:
        eW = eW + 1
        eK = np.zeros_like(eK)
        A = balanced_slices(W / 2.0 ** eW[:, None], s)
        B = rounded_slices(K, s)
    else:
        A = slices(W / 2.0 ** eW[:, None], s)
        B = slices(K / 2.0 ** eK[None, :], s)
    V = np.zeros((W.shape[0], K.shape[1]))
    for g in range(s + 1, 1, -1):                 # smallest terms first
        acc = np.zeros((W.shape[0], K.shape[1]), dtype=np.int64)
        for t in range(1, s + 1):
            u = g - t
            if 1 <= u <= s:
                acc += A[t - 1] @ B[u - 1]
        assert np.max(np.abs(acc)) < 2 ** 31
        V += acc.astype(float) * 2.0 ** (-7 * g)
    return V * 2.0 ** eW[:, None] * 2.0 ** eK[None, :]


def main():
    N, d, M = int(sys.argv[1]) if len(sys.argv) > 1 else 2048, 6, 256
    rs = np.random.RandomState(0)
    X = rs.rand(N, d)
    Y = orc.hartman6_neg(X)
    theta = [.53, .57, 2.5, .34, .27, .35]
    gp = orc.GPOracle(orc.KernelSpec(orc.K_SE_ARD, theta, d), X, Y, 0.1)
    Xs = np.random.RandomState(1).rand(M, d)
    Ks = gp.kernel.cross(gp.X, Xs)
    W = solve_triangular(gp.L, np.eye(N), lower=True)
    bY = W @ Y
    V0 = W @ Ks
    mu0 = V0.T @ bY
    s20 = np.clip(1.1 - np.sum(V0 * V0, axis=0), 1e-8, 10)
    ei0 = orc.score(orc.ACQ_EI, "cpp", mu0, s20, Y.max(), 0.01)
    print("N=%d  sigma^2 range [%.3g, %.3g]  EI range [%.3g, %.3g]" % (N, s20.min(), s20.max(), ei0.min(), ei0.max()))
    for s, bal in ((5, False), (6, False), (7, False), (8, False), (6, True), (7, True)):
        V = emulate(W, Ks, s, bal)
        mu = V.T @ bY
        s2 = np.clip(1.1 - np.sum(V * V, axis=0), 1e-8, 10)
        ei = orc.score(orc.ACQ_EI, "cpp", mu, s2, Y.max(), 0.01)
        print(("balanced " if bal else "truncated") + " s=%d  passes=%2d  max|dV|=%.2e  mu rel %.2e  s2 rel %.2e  EI rel(floor 1e-5) %.2e  argmax same: %s"
              % (s, s * (s + 1) // 2, np.max(np.abs(V - V0)), np.max(np.abs(mu - mu0) / np.maximum(np.abs(mu0), 1e-3)),
                 np.max(np.abs(s2 - s20) / s20), np.max(np.abs(ei - ei0) / np.maximum(np.abs(ei0), 1e-5)),
                 int(np.argmax(ei)) == int(np.argmax(ei0))))


if __name__ == "__main__":
    main()
