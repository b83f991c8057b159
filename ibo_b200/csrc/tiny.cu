// Fused scoring kernel for small models (N <= 128: one 128-row block) -- the reference's own operating point
// (demo.py, the unit tests and interactive galleries work with tens of observations, BASELINE.json config #1).
//
// Replaces, in one launch, GaussianProcess.posterior + EI/PI/UCB.negf per candidate (ego/gaussianprocess/__init__.py:169-228,
// ego/acquisition/__init__.py:60-164) / GP_Maximizer::posterior + negei/negpi/negucb (cpp/optimizeGP.cpp:57-236).
//
// For such models the general path (K1 -> K2 -> K3, score.cu) is pure latency: three dependent launches, an H2D DMA and
// a slab round trip for ~N^2/2 = a few thousand FMAs per candidate.  Here one CTA keeps W = inv(L) transposed in shared
// memory and walks tiles of 8 candidates: k* by direct differences (no expansion, no cancellation), v = W k* with one
// thread per training row (ascending-k FMA chains, 8 candidates in flight per thread), the three row reductions in a
// fixed order, then the epilogue.  For a DIRECT batch the candidates are read straight from mapped pinned host memory
// and the values are written straight back to it: one launch and one stream synchronisation per batch.
// A candidate's value is a function of (model, x) only (fixed summation orders), so DIRECT trajectories stay reproducible.
#include "model.cuh"
#include "scoremath.cuh"
#include <cstring>
#include <mutex>

namespace ibo {
namespace {

constexpr int TC = 8;        // candidates per tile
constexpr int WS = 129;      // row stride of W in shared memory: odd, so that thread r walking row r is conflict free

struct TinySide {            // one factor: the model itself, or the variance model of PrefGaussianProcess.addObservationPoint
    const double *W, *Xt, *center;
    int N;
};

struct TinyParams {
    TinySide side[2];        // [1].N == 0 when there is no variance model
    const double *betaY, *beta1, *invTheta, *cand;
    const double *pmeans, *pbeta, *plb, *pwidth;
    double *score, *mu, *s2;
    double* blkBest; long long* blkIdx;
    long M;
    int d, acq, mode_py, npb, want_argmax;
    double sf2, noise, ymax, parm, ptheta;
};

__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src));
}

template <int KC>
__global__ void __launch_bounds__(128) tiny_fused_kernel(const __grid_constant__ TinyParams P, const __grid_constant__ CandInline I) {
    extern __shared__ double sm[];
    const double* __restrict__ cands = P.cand ? P.cand : I.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = P.d, XS = d | 1;                   // odd row stride of the training inputs
    const int nside = P.side[1].N > 0 ? 2 : 1;
    // shared-memory carve-up: per side W [N][WS] (lower triangle) and X [N][XS]; then k* [TC][128], the raw candidates of
    // the tile [TC][d], 1/theta [d], and the reduction scratch [3][4][TC]
    double* sW[2]; double* sX[2];
    double* ptr = sm;
    for (int s = 0; s < nside; s++) { sW[s] = ptr; ptr += (size_t)P.side[s].N * WS; sX[s] = ptr; ptr += (size_t)P.side[s].N * XS; }
    double* sK = ptr;  ptr += TC * 128;
    double* sC = ptr;  ptr += TC * d;
    double* sT = ptr;  ptr += d;             // 1 / theta
    double* sCtr = ptr; ptr += 2 * d;        // centres of the two sides
    double* sR = ptr;
    // stage the factors with 8-byte async copies: every element is in flight at once (a register-staged copy of W costs one
    // L2 round trip per row and was most of the kernel's time)
    for (int s = 0; s < nside; s++) {
        const int N = P.side[s].N;
        for (int idx = tid; idx < N * 128; idx += 128) {
            const int r = idx >> 7, c = idx & 127;
            if (c <= r) cp_async8(sW[s] + r * WS + c, P.side[s].W + (size_t)r * 128 + c);
        }
        for (int idx = tid; idx < N * d; idx += 128) {
            const int r = idx / d, q = idx - r * d;
            cp_async8(sX[s] + r * XS + q, P.side[s].Xt + idx);
        }
    }
    for (int q = tid; q < d; q += 128) {
        cp_async8(sT + q, P.invTheta + q);
        for (int s = 0; s < nside; s++) cp_async8(sCtr + s * d + q, P.side[s].center + q);
    }
    asm volatile("cp.async.commit_group;");
    // The first tile's candidates are fetched while the copies above are in flight: for a DIRECT batch they come over PCIe
    // (mapped host memory), the longest latency in the kernel -- everything else hides behind it.
    const long ntile = (P.M + TC - 1) / TC;
    for (int idx = tid; idx < TC * d; idx += 128) {
        const int j = idx / d, q = idx - j * d;
        long m = (long)blockIdx.x * TC + j; if (m >= P.M) m = P.M - 1;
        sC[idx] = cands[(size_t)m * d + q];
    }
    const int N0 = P.side[0].N;
    const double by = tid < N0 ? P.betaY[tid] : 0.0;
    const double b1 = (tid < N0 && P.npb > 0) ? P.beta1[tid] : 0.0;
    double best = -INFINITY;
    long long bestIdx = 0x7fffffffffffffffLL;
    asm volatile("cp.async.wait_group 0;");
    for (long t = blockIdx.x; t < ntile; t += gridDim.x) {
        const long m0 = t * TC;
        if (t != blockIdx.x) {
            __syncthreads();                 // previous tile's scratch is free
            for (int idx = tid; idx < TC * d; idx += 128) {
                const int j = idx / d, q = idx - j * d;
                long m = m0 + j; if (m >= P.M) m = P.M - 1;
                sC[idx] = cands[(size_t)m * d + q];
            }
        }
        double q_sum = 0, p_sum = 0, p1_sum = 0;           // of candidate m0 + tid (threads 0..7)
        for (int s = 0; s < nside; s++) {
            const int N = P.side[s].N;
            __syncthreads();
            // k*[j][c], thread c: direct differences between its training row and the scaled, centred candidates
            {
                double r2[TC];
#pragma unroll
                for (int j = 0; j < TC; j++) r2[j] = 0.0;
                if (tid < N) {
                    const double* x = sX[s] + tid * XS;
                    for (int q = 0; q < d; q++) {
                        const double xv = x[q], it = sT[q], ctr = sCtr[s * d + q];     // rows are stored scaled and centred
#pragma unroll
                        for (int j = 0; j < TC; j++) { const double df = xv - (sC[j * d + q] * it - ctr); r2[j] = fma(df, df, r2[j]); }
                    }
                }
#pragma unroll
                for (int j = 0; j < TC; j++) sK[j * 128 + tid] = tid < N ? cov_r2_t<KC>(P.sf2, r2[j]) : 0.0;
            }
            __syncthreads();
            // v_r[j] = sum_{c <= r} W[r][c] k*[j][c], thread r
            double v[TC];
#pragma unroll
            for (int j = 0; j < TC; j++) v[j] = 0.0;
            if (tid < N) {
                const double* w = sW[s] + tid * WS;
                for (int c = 0; c <= tid; c++) {
                    const double wv = w[c];
#pragma unroll
                    for (int j = 0; j < TC; j++) v[j] = fma(wv, sK[j * 128 + c], v[j]);
                }
            }
            // row reductions: xor tree inside the warp, then the four warps in ascending order
#pragma unroll
            for (int j = 0; j < TC; j++) {
                double q = v[j] * v[j], p = v[j] * by, p1 = v[j] * b1;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    q += __shfl_xor_sync(0xffffffffu, q, o);
                    p += __shfl_xor_sync(0xffffffffu, p, o);
                    p1 += __shfl_xor_sync(0xffffffffu, p1, o);
                }
                if (lane == 0) { sR[(0 * 4 + warp) * TC + j] = q; sR[(1 * 4 + warp) * TC + j] = p; sR[(2 * 4 + warp) * TC + j] = p1; }
            }
            __syncthreads();
            if (tid < TC) {
                double q = 0, p = 0, p1 = 0;
#pragma unroll
                for (int w = 0; w < 4; w++) { q += sR[(0 * 4 + w) * TC + tid]; p += sR[(1 * 4 + w) * TC + tid]; p1 += sR[(2 * 4 + w) * TC + tid]; }
                q_sum = q;                                  // the variance comes from the last side (the aug factor if there is one)
                if (s == 0) { p_sum = p; p1_sum = p1; }     // the mean always from the model itself (gaussianprocess/__init__.py:214-223)
            }
        }
        if (tid < TC && m0 + tid < P.M) {
            const long m = m0 + tid;
            double mean0 = 0.0;
            if (P.npb > 0) mean0 = prior_mean(sC + tid * d, d, P.npb, P.pmeans, P.pbeta, P.ptheta, P.plb, P.pwidth);
            const double mu = mean0 + p_sum - mean0 * p1_sum;
            double s2 = (1.0 + P.noise) - q_sum;
            const double floor_ = P.mode_py ? 10e-8 : 1e-8;   // gaussianprocess/__init__.py:224 vs cpp/optimizeGP.cpp:150
            s2 = s2 < floor_ ? floor_ : (s2 > 10.0 ? 10.0 : s2);
            if (P.mu) P.mu[m] = mu;
            if (P.s2) P.s2[m] = s2;
            if (P.acq >= 0) {
                const double val = acq_value(P.acq, P.mode_py, mu, s2, P.ymax, P.parm);
                if (P.score) P.score[m] = val;
                if (val == val) { if (val > best || (val == best && m < bestIdx)) { best = val; bestIdx = m; } }
                else if (bestIdx == 0x7fffffffffffffffLL) bestIdx = m;      // NaN never wins, but an all-NaN set still names an index
            }
        }
    }
    if (P.acq < 0 || !P.want_argmax) return;
    // the tile owners (threads 0..7, all in warp 0) combine: lowest index wins ties
    if (warp == 0) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            const double os = __shfl_xor_sync(0xffffffffu, best, o);
            const long long oi = __shfl_xor_sync(0xffffffffu, bestIdx, o);
            if (os > best || (os == best && oi < bestIdx)) { best = os; bestIdx = oi; }
        }
        if (lane == 0) { P.blkBest[blockIdx.x] = best; P.blkIdx[blockIdx.x] = bestIdx; }
    }
}

__global__ void __launch_bounds__(256) tiny_argmax_kernel(const double* __restrict__ blkBest, const long long* __restrict__ blkIdx,
                                                          int nblk, double* __restrict__ best, long long* __restrict__ bestIdx) {
    __shared__ double ws[256];
    __shared__ long long wi[256];
    double sc = -INFINITY;
    long long idx = 0x7fffffffffffffffLL;
    for (int b = threadIdx.x; b < nblk; b += 256) {
        const double os = blkBest[b]; const long long oi = blkIdx[b];
        if (os > sc || (os == sc && oi < idx)) { sc = os; idx = oi; }
    }
    ws[threadIdx.x] = sc; wi[threadIdx.x] = idx;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            const double os = ws[threadIdx.x + o]; const long long oi = wi[threadIdx.x + o];
            if (os > ws[threadIdx.x] || (os == ws[threadIdx.x] && oi < wi[threadIdx.x])) { ws[threadIdx.x] = os; wi[threadIdx.x] = oi; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { *best = ws[0]; *bestIdx = wi[0]; }
}

cudaError_t set_tiny_attrs() {
    const int maxsm = 200 * 1024;
    cudaError_t e = cudaFuncSetAttribute(tiny_fused_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tiny_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tiny_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxsm);
    return e;
}

size_t tiny_smem(int N, int Nvar, int d) {
    return sizeof(double) * ((size_t)(N + Nvar) * (WS + (d | 1)) + TC * 128 + (size_t)TC * d + 3 * d + 3 * 4 * TC);
}
constexpr size_t TINY_SMEM_MAX = 200 * 1024;

}  // namespace

// models the fused kernel serves: a single row-block (also for the variance model of PrefGP's aug factor, if any) that
// fits shared memory; IBO_TINY=0 disables
// Batches above TINY_MAX_M go through the general path, whose DMMA kernels have the higher throughput (2^20 candidates at
// N = 128: 1.7 ms vs 6.3 ms here); the two paths agree to rounding, and every DIRECT / gallery batch is far below the limit.
constexpr long TINY_MAX_M = 4096;

bool tiny_eligible(const ibo_model* m, long M) {
    if (M > TINY_MAX_M) return false;
    if (m->nb != 1 || m->d > 64) return false;
    const ibo_model* vm = m->var_model;
    if (vm && (vm->nb != 1 || vm->kind != m->kind || vm->sf2 != m->sf2)) return false;
    if (tiny_smem(m->N, vm ? vm->N : 0, m->d) > TINY_SMEM_MAX) return false;
    return get_option(OPT_TINY) != 0;            // option tiny = 0 disables (the tests compare both paths in one process)
}

// Scores M candidates at `cand` (device memory, or mapped pinned host memory) into out = [score | mu | s2][M] (+ the argmax
// pair at out[3M], out[3M+1]); everything is enqueued on m->stream.
// `host_cand`: the same candidates in host memory when the caller has them there (small batches ride in the parameter buffer)
int score_tiny(ibo_model* m, const double* cand, long M, const ScoreReq& rq, double* out, const double* host_cand) {
    const cudaError_t ae = ensure_attrs(m->device, ATTR_TINY, set_tiny_attrs);
    if (ae != cudaSuccess) { set_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(ae)); return IBO_E_CUDA; }
    const int g_tiny_sms = dev_info(m->device).sms;
    cudaStream_t st = m->stream;
    const bool prof = (rq.flags & IBO_FLAG_PROFILE) != 0;
    const long ntile = (M + TC - 1) / TC;
    // one CTA per tile up to a few waves; beyond that CTAs walk tiles so that W is staged once per CTA
    const ibo_model* vm = m->var_model;
    const size_t smem = tiny_smem(m->N, vm ? vm->N : 0, m->d);
    const int perSM = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / smem));
    const int grid = (int)std::min<long>(ntile, (long)g_tiny_sms * perSM);
    if (m->blkCap < (size_t)grid) {
        if (m->dBlkBest) cudaFree(m->dBlkBest);
        if (m->dBlkIdx) cudaFree(m->dBlkIdx);
        m->dBlkBest = nullptr; m->dBlkIdx = nullptr; m->blkCap = 0;
        const size_t cap = (size_t)g_tiny_sms * 8;
        IBO_CUDA_TRY(cudaMalloc(&m->dBlkBest, sizeof(double) * cap));
        IBO_CUDA_TRY(cudaMalloc(&m->dBlkIdx, sizeof(long long) * cap));
        m->blkCap = cap;
    }
    TinyParams P;
    P.side[0] = TinySide{m->dW, m->dXt, m->dCenter, m->N};
    P.side[1] = vm ? TinySide{vm->dW, vm->dXt, vm->dCenter, vm->N} : TinySide{nullptr, nullptr, nullptr, 0};
    P.betaY = m->dBetaY; P.beta1 = m->dBeta1; P.invTheta = m->dInvTheta; P.cand = cand;
    P.pmeans = m->dPmeans; P.pbeta = m->dPbeta; P.plb = m->dPlb; P.pwidth = m->dPwidth;
    P.score = rq.want_score ? out : nullptr;
    P.mu = rq.want_mu ? out + M : nullptr;
    P.s2 = rq.want_s2 ? out + 2 * M : nullptr;
    P.blkBest = m->dBlkBest; P.blkIdx = m->dBlkIdx;
    P.M = M; P.d = m->d; P.acq = rq.acq; P.mode_py = (rq.flags & IBO_FLAG_MODE_PY) ? 1 : 0; P.npb = m->npb;
    P.want_argmax = (rq.acq >= 0 && rq.want_argmax) ? 1 : 0;
    P.sf2 = m->sf2; P.noise = m->noise; P.ymax = rq.ymax; P.parm = rq.parm; P.ptheta = m->ptheta;
    static CandInline I;        // launches copy the parameter buffer synchronously; one host thread per model
    CandInline Ilocal;
    CandInline* Ip = &I;
    if (host_cand && (size_t)M * m->d <= (size_t)CAND_INLINE) {
        std::memcpy(Ilocal.x, host_cand, sizeof(double) * (size_t)M * m->d);
        P.cand = nullptr;
        Ip = &Ilocal;
    }
    if (prof) IBO_CUDA_TRY(cudaEventRecord(m->ev[0], st));
    if (m->kind <= IBO_KERNEL_SE_ISO) tiny_fused_kernel<0><<<grid, 128, smem, st>>>(P, *Ip);
    else if (m->kind == IBO_KERNEL_MATERN3) tiny_fused_kernel<1><<<grid, 128, smem, st>>>(P, *Ip);
    else tiny_fused_kernel<2><<<grid, 128, smem, st>>>(P, *Ip);
    long nlaunch = 1;
    if (P.want_argmax) {
        tiny_argmax_kernel<<<1, 256, 0, st>>>(m->dBlkBest, m->dBlkIdx, grid, out + 3 * M, reinterpret_cast<long long*>(out + 3 * M + 1));
        nlaunch++;
    }
    g_launches += nlaunch;
    if (prof) {
        IBO_CUDA_TRY(cudaEventRecord(m->ev[5], st));
        IBO_CUDA_TRY(cudaEventSynchronize(m->ev[5]));
        float tot; cudaEventElapsedTime(&tot, m->ev[0], m->ev[5]);
        m->prof[0] = 0; m->prof[1] = tot; m->prof[2] = 0; m->prof[3] = tot; m->prof[4] = (double)nlaunch; m->prof[5] = 1;
    }
    IBO_CUDA_TRY(cudaGetLastError());
    return IBO_OK;
}

}  // namespace ibo
