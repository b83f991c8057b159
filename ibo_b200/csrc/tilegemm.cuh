// 128x128xK tile GEMM on the FP64 tensor pipe (DMMA.8x8x4) from row-major global operands; shared by the model build
// (model.cu), the Laplace fit (laplace.cu) and the marginal-likelihood gradient (hyper.cu).
#pragma once
#include "common.cuh"

namespace ibo {

// ---------------------------------------------------------------------------------------------
// 128x128x(K) tile GEMM on the DMMA pipe from row-major global operands.
//   8 warps, warp tile 64x32 (8 x 4 DMMA tiles, 64 FP64 accumulators per thread),
//   cp.async double-buffered 16-deep k-steps into padded (bank-conflict-free) shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int AS_STRIDE = 20;    // [128][20]: (row*20 + k) mod 16 distinct over a half-warp
constexpr int BN_STRIDE = 132;   // [16][132] for the non-transposed B operand
constexpr int TILE_SMEM_DOUBLES = 2 * (128 * AS_STRIDE) + 2 * (128 * AS_STRIDE);   // A + B(T) buffers (B(N) fits too)

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// acc[mt][nt][2] += op(A)(128 x K) * op(B); TRANSB: B stored [n][k] (row-major, ldb), else [k][n];
// TRANSA: A stored [k][m] (row-major, lda), else [m][k].
template <bool TRANSB, bool TRANSA = false>
__device__ __forceinline__ void tile_gemm_core(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                                               int K, double (&acc)[8][4][2], double* sm) {
    double* sA = sm;                         // [2][128*AS_STRIDE]
    double* sB = sm + 2 * 128 * AS_STRIDE;   // [2][128*AS_STRIDE] or [2][16*BN_STRIDE]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int nk = K / BK;
    auto load_stage = [&](int kb, int buf) {
        // A: 128 rows x 16 doubles = 1024 16-byte chunks
        if (TRANSA) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                int idx = tid + c * 256;
                int r = idx >> 6, ch = idx & 63;
                cp_async16(sA + buf * 16 * BN_STRIDE + r * BN_STRIDE + ch * 2, A + (size_t)(kb * BK + r) * lda + ch * 2);
            }
        } else {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                int idx = tid + c * 256;
                int r = idx >> 3, ch = idx & 7;
                cp_async16(sA + buf * 128 * AS_STRIDE + r * AS_STRIDE + ch * 2, A + (size_t)r * lda + kb * BK + ch * 2);
            }
        }
        if (TRANSB) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                int idx = tid + c * 256;
                int r = idx >> 3, ch = idx & 7;
                cp_async16(sB + buf * 128 * AS_STRIDE + r * AS_STRIDE + ch * 2, B + (size_t)r * ldb + kb * BK + ch * 2);
            }
        } else {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                int idx = tid + c * 256;
                int r = idx >> 6, ch = idx & 63;
                cp_async16(sB + buf * 16 * BN_STRIDE + r * BN_STRIDE + ch * 2, B + (size_t)(kb * BK + r) * ldb + ch * 2);
            }
        }
        cp_async_commit();
    };
    load_stage(0, 0);
    for (int kb = 0; kb < nk; kb++) {
        const int buf = kb & 1;
        if (kb + 1 < nk) { load_stage(kb + 1, buf ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const double* a_s = TRANSA ? sA + buf * 16 * BN_STRIDE + (lane & 3) * BN_STRIDE + wm * 64 + (lane >> 2)
                                   : sA + buf * 128 * AS_STRIDE + (wm * 64 + (lane >> 2)) * AS_STRIDE + (lane & 3);
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
            double af[8], bf[4];
#pragma unroll
            for (int mt = 0; mt < 8; mt++) af[mt] = TRANSA ? a_s[ks * 4 * BN_STRIDE + mt * 8] : a_s[mt * 8 * AS_STRIDE + ks * 4];
            if (TRANSB) {
                const double* b_s = sB + buf * 128 * AS_STRIDE + (wn * 32 + (lane >> 2)) * AS_STRIDE + (lane & 3);
#pragma unroll
                for (int nt = 0; nt < 4; nt++) bf[nt] = b_s[nt * 8 * AS_STRIDE + ks * 4];
            } else {
                const double* b_s = sB + buf * 16 * BN_STRIDE + (ks * 4 + (lane & 3)) * BN_STRIDE + wn * 32 + (lane >> 2);
#pragma unroll
                for (int nt = 0; nt < 4; nt++) bf[nt] = b_s[nt * 8];
            }
#pragma unroll
            for (int mt = 0; mt < 8; mt++)
#pragma unroll
                for (int nt = 0; nt < 4; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
        }
        __syncthreads();
    }
}

}  // namespace ibo
