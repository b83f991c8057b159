// 128x128xK tile GEMM on the FP64 tensor pipe (DMMA.8x8x4) from row-major global operands; shared by the model build
// (model.cu), the Laplace fit (laplace.cu) and the marginal-likelihood gradient (hyper.cu).
#pragma once
#include "common.cuh"

namespace ibo {

// ---------------------------------------------------------------------------------------------
// 128x128x(K) tile GEMM on the DMMA pipe from row-major global operands.
//   8 warps, warp tile 64x32 (8 x 4 DMMA tiles, 64 FP64 accumulators per thread),
//   16-deep k-steps through a 4-stage cp.async ring of padded (bank-conflict-free) shared-memory buffers: three k-steps of
//   loads in flight cover the L2 / HBM latency of a GPU whose 148 SMs all stream tiles at once (with two stages the tile took
//   24 us alone but ~30 us in a full grid), one barrier per k-step.
// ---------------------------------------------------------------------------------------------
constexpr int AS_STRIDE = 20;    // [128][20]: (row*20 + k) mod 16 distinct over a half-warp
constexpr int BN_STRIDE = 132;   // [16][132] for the non-transposed B operand
constexpr int TILE_OPERAND = 128 * AS_STRIDE;                          // one 128-row operand of one stage ([16][BN_STRIDE] fits too)
template <int MROWS> struct TileCfg {
    static constexpr int THREADS = 2 * MROWS;                          // 8 warps for 128 rows, 4 for 64
    static constexpr int STAGES = MROWS == 128 ? 4 : 3;
    static constexpr int STAGE = MROWS * AS_STRIDE + TILE_OPERAND;     // A rows, then B
    static constexpr int SMEM_DOUBLES = STAGES * STAGE;                // 160 KB / 90 KB (two CTAs per SM)
};
constexpr int TILE_SMEM_DOUBLES = TileCfg<128>::SMEM_DOUBLES;

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// acc[mt][nt][2] += op(A)(MROWS x K) * op(B)(K x 128); TRANSB: B stored [n][k] (row-major, ldb), else [k][n];
// TRANSA: A stored [k][m] (row-major, lda), else [m][k] (128 rows only).
// MROWS = 64: half-height tiles from 4-warp CTAs, two of which share an SM, so that one CTA's prologue and read-modify-write
// epilogue run under the other's DMMA loop (the bulk updates of the factorisation: ncu showed the DMMA pipe 72 % busy with one
// 8-warp CTA per SM).
template <bool TRANSB, bool TRANSA = false, int MROWS = 128>
__device__ __forceinline__ void tile_gemm_core(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb,
                                               int K, double (&acc)[8][4][2], double* sm) {
    using Cfg = TileCfg<MROWS>;
    static_assert(MROWS == 128 || !TRANSA, "half-height tiles take A row-major only");
    constexpr int NT = Cfg::THREADS, TILE_STAGES = Cfg::STAGES;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int nk = K / BK;
    auto load_stage = [&](int kb, int buf) {
        double* sA = sm + buf * Cfg::STAGE;
        double* sB = sA + MROWS * AS_STRIDE;
        // A: 128 rows x 16 doubles = 1024 16-byte chunks
        if (TRANSA) {
#pragma unroll
            for (int c = 0; c < 1024 / NT; c++) {
                int idx = tid + c * NT;
                int r = idx >> 6, ch = idx & 63;
                cp_async16(sA + r * BN_STRIDE + ch * 2, A + (size_t)(kb * BK + r) * lda + ch * 2);
            }
        } else {
#pragma unroll
            for (int c = 0; c < 4; c++) {      // MROWS x 8 chunks over 2 MROWS threads
                int idx = tid + c * NT;
                int r = idx >> 3, ch = idx & 7;
                cp_async16(sA + r * AS_STRIDE + ch * 2, A + (size_t)r * lda + kb * BK + ch * 2);
            }
        }
        if (TRANSB) {
#pragma unroll
            for (int c = 0; c < 1024 / NT; c++) {
                int idx = tid + c * NT;
                int r = idx >> 3, ch = idx & 7;
                cp_async16(sB + r * AS_STRIDE + ch * 2, B + (size_t)r * ldb + kb * BK + ch * 2);
            }
        } else {
#pragma unroll
            for (int c = 0; c < 1024 / NT; c++) {
                int idx = tid + c * NT;
                int r = idx >> 6, ch = idx & 63;
                cp_async16(sB + r * BN_STRIDE + ch * 2, B + (size_t)(kb * BK + r) * ldb + ch * 2);
            }
        }
    };
    // one commit group per k-step, empty ones past the end, so that "all but the newest TILE_STAGES - 2 groups" is always k-step kb
#pragma unroll
    for (int s = 0; s < TILE_STAGES - 1; s++) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }
    for (int kb = 0; kb < nk; kb++) {
        const int buf = kb % TILE_STAGES;
        cp_async_wait<TILE_STAGES - 2>();
        __syncthreads();          // k-step kb has landed for everyone; the buffer of k-step kb - 1 is free for the load below
        if (kb + TILE_STAGES - 1 < nk) load_stage(kb + TILE_STAGES - 1, (kb + TILE_STAGES - 1) % TILE_STAGES);
        cp_async_commit();
        const double* sA = sm + buf * Cfg::STAGE;
        const double* sB = sA + MROWS * AS_STRIDE;
        const double* a_s = TRANSA ? sA + (lane & 3) * BN_STRIDE + wm * 64 + (lane >> 2)
                                   : sA + (wm * 64 + (lane >> 2)) * AS_STRIDE + (lane & 3);
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
            double af[8], bf[4];
#pragma unroll
            for (int mt = 0; mt < 8; mt++) af[mt] = TRANSA ? a_s[ks * 4 * BN_STRIDE + mt * 8] : a_s[mt * 8 * AS_STRIDE + ks * 4];
            if (TRANSB) {
                const double* b_s = sB + (wn * 32 + (lane >> 2)) * AS_STRIDE + (lane & 3);
#pragma unroll
                for (int nt = 0; nt < 4; nt++) bf[nt] = b_s[nt * 8 * AS_STRIDE + ks * 4];
            } else {
                const double* b_s = sB + (ks * 4 + (lane & 3)) * BN_STRIDE + wn * 32 + (lane >> 2);
#pragma unroll
                for (int nt = 0; nt < 4; nt++) bf[nt] = b_s[nt * 8];
            }
#pragma unroll
            for (int mt = 0; mt < 8; mt++)
#pragma unroll
                for (int nt = 0; nt < 4; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();              // every read of the operands is complete: callers may overwrite them in place
}

}  // namespace ibo
