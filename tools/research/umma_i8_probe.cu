// Research probe for the next round (not part of libibo_b200): the B200's INT8 tensor path, tcgen05.mma kind::i8 with
// shared-memory operands in the no-swizzle K-major canonical layout and INT32 accumulators in TMEM.
//   1. correctness of the descriptor encoding (shared-memory matrix descriptor: LBO = byte stride between the two 16-byte
//      K chunks of an MMA, SBO = byte stride between 8-row groups; instruction descriptor: S32 accumulator, signed 8-bit
//      A and B, K-major, N >> 3, M >> 4), checked against a CPU integer GEMM;
//   2. the issue rate of back-to-back MMAs from resident operands for N = 64 / 128 / 256 (whole GPU, one CTA per SM),
//      i.e. the INT8 tensor peak an Ozaki-style FP64 emulation of K2 would be measured against
//      (tools/research/ozaki_int8_study.py).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_i8_probe umma_i8_probe.cu ; run under `timeout`.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, LBO, SBO in 16-byte units, version 1, no swizzle
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;            // version_ = 1 (Blackwell)
    return d;                           // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE / interleave)
}

// instruction descriptor (cute::UMMA::InstrDescriptor) for kind::i8: c_format S32 (2) at bit 4, a/b_format INT8 (1) at bits 7 / 10,
// a_major / b_major K (0), n_dim = N >> 3 at bit 17, m_dim = M >> 4 at bit 24
__host__ __device__ constexpr uint32_t instr_desc_i8(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n\t"
        "D_%=:\n\t}"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// canonical no-swizzle K-major tile: rows x K bytes; 8x16-byte core matrices; core (r/8, k/16) at ((r/8) * (K/16) + k/16) * 128
__device__ __forceinline__ uint32_t canon_off(int r, int k, int K) { return (uint32_t)((((r >> 3) * (K >> 4) + (k >> 4)) << 7) + ((r & 7) << 4) + (k & 15)); }

// ---- 1. correctness: D[128][N] = A[128][K] * B[N][K]^T ----------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(128) probe_correct(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int32_t* __restrict__ D,
                                                     int K, int swap) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t* sA = sm;                         // 128 x K
    uint8_t* sB = sm + 128 * K;               // N x K
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 128 * K; i += 128) sA[canon_off(i / K, i % K, K)] = (uint8_t)A[i];
    for (int i = tid; i < N * K; i += 128) sB[canon_off(i / K, i % K, K)] = (uint8_t)B[i];
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(&tbase, N < 32 ? 32 : N);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t td = tbase;
    if (tid == 0) {
        const uint32_t lbo = swap ? (uint32_t)(K >> 4) * 128u : 128u;     // K-chunk stride
        const uint32_t sbo = swap ? 128u : (uint32_t)(K >> 4) * 128u;     // 8-row-group stride
        const uint32_t idesc = instr_desc_i8(128, N);
        for (int kk = 0; kk < K / 32; kk++) {
            uint64_t da = smem_desc(smem_u32(sA) + kk * 256, lbo, sbo);
            uint64_t db = smem_desc(smem_u32(sB) + kk * 256, lbo, sbo);
            umma_i8(td, da, db, idesc, kk > 0);
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < N; c += 8) {
        uint32_t v[8];
        tmem_ld8(td + ((uint32_t)(warp * 32) << 16) + c, v);
        for (int j = 0; j < 8; j++) D[(warp * 32 + lane) * N + c + j] = (int32_t)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(td, N < 32 ? 32 : N);
}

// ---- 2. issue rate: ITER x (KT/32) MMAs of 128 x N x 32 from resident operands, one CTA per SM ---------------------------
template <int N, int KT>
__global__ void __launch_bounds__(128) probe_rate(int iters, int32_t* __restrict__ sink) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t* sA = sm;
    uint8_t* sB = sm + 128 * KT;
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (128 + N) * KT; i += 128) sm[i] = (uint8_t)((i * 7 + 3) & 3);
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(&tbase, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t td = tbase;
    if (tid == 0) {
        const uint32_t lbo = 128u, sbo = (uint32_t)(KT >> 4) * 128u;
        const uint32_t idesc = instr_desc_i8(128, N);
        for (int it = 0; it < iters; it++) {
            const uint32_t acc = td + (uint32_t)((it & 1) * N) % 512;     // two accumulators alternate
#pragma unroll
            for (int kk = 0; kk < KT / 32; kk++)
                umma_i8(acc, smem_desc(smem_u32(sA) + kk * 256, lbo, sbo), smem_desc(smem_u32(sB) + kk * 256, lbo, sbo), idesc, 1);
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[8];
    tmem_ld8(td + ((uint32_t)(warp * 32) << 16), v);
    if (v[0] == 0x7fffffff) sink[blockIdx.x * 128 + tid] = (int32_t)v[lane & 7];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(td, 512);
}

template <int N>
static int run_correct(int K) {
    std::vector<int8_t> hA(128 * K), hB(N * K);
    srand(1);
    for (auto& x : hA) x = (int8_t)(rand() % 255 - 127);
    for (auto& x : hB) x = (int8_t)(rand() % 255 - 127);
    std::vector<int32_t> ref(128 * N), out(128 * N);
    for (int i = 0; i < 128; i++)
        for (int j = 0; j < N; j++) {
            int32_t s = 0;
            for (int k = 0; k < K; k++) s += (int32_t)hA[i * K + k] * (int32_t)hB[j * K + k];
            ref[i * N + j] = s;
        }
    int8_t *dA, *dB; int32_t* dD;
    CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dD, out.size() * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
    for (int swap = 0; swap < 2; swap++) {
        CK(cudaMemset(dD, 0xff, out.size() * 4));
        probe_correct<N><<<1, 128, (128 + N) * K>>>(dA, dB, dD, K, swap);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("correct N=%d K=%d swap=%d: kernel failed: %s\n", N, K, swap, cudaGetErrorString(e)); return 1; }
        CK(cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost));
        long bad = 0;
        for (size_t i = 0; i < out.size(); i++) bad += out[i] != ref[i];
        printf("correct N=%d K=%d lbo/sbo %s: %ld of %zu elements differ (D[0]=%d ref %d, D[last]=%d ref %d)\n", N, K,
               swap ? "swapped (LBO = row-group stride)" : "as documented (LBO = K-chunk stride)", bad, out.size(), out[0], ref[0],
               out.back(), ref.back());
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return 0;
}

template <int N, int KT>
static int run_rate(int sms) {
    int32_t* sink; CK(cudaMalloc(&sink, sizeof(int32_t) * 128 * sms));
    CK(cudaFuncSetAttribute(probe_rate<N, KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (128 + N) * KT));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 4000;
    probe_rate<N, KT><<<sms, 128, (128 + N) * KT>>>(200, sink);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(e0));
        probe_rate<N, KT><<<sms, 128, (128 + N) * KT>>>(iters, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    double ops = 2.0 * 128 * N * KT * (double)iters * sms;
    printf("rate   N=%3d KT=%3d: %.3f ms for %d iterations on %d SMs -> %.1f TOP/s (int8 MACs x 2)\n", N, KT, best, iters, sms, ops / best / 1e9);
    cudaFree(sink);
    return 0;
}

int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    printf("%s, %d SMs\n", pr.name, pr.multiProcessorCount);
    if (run_correct<64>(32)) return 1;
    if (run_correct<64>(128)) return 1;
    if (run_correct<256>(64)) return 1;
    if (run_rate<64, 128>(pr.multiProcessorCount)) return 1;
    if (run_rate<128, 128>(pr.multiProcessorCount)) return 1;
    if (run_rate<256, 128>(pr.multiProcessorCount)) return 1;
    return 0;
}
