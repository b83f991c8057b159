// Research probe (not part of libibo_b200): issue rate of the INT8 K2's k-step schedule (10 tcgen05.mma kind::i8 instructions
// covering 28 digit pairs, 128 x 64n x 32 each) from RESIDENT operands, one CTA per SM, no data movement:
//   mode 0: every A digit from shared memory (the round-1 kernel)            -> operand fetch 96 KiB per k-step
//   mode 1: A digits 1..4 from TMEM (tcgen05.mma [d], [a], b_desc), 5..7 from shared memory -> 68 KiB per k-step
//   mode 2: as mode 1, and four loader warps rewrite the TMEM A buffer between k-steps with the kernel's handshake
//           (commit -> aempty -> tcgen05.st x4 -> wait::st -> arrive -> afull), data from registers (no global loads)
// Tensor time of a k-step is 28 x 32 = 896 clk at the nominal 8192 MAC/clk/SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_i8_ts_probe umma_i8_ts_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t sdesc(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ constexpr uint32_t idesc(int M, int N) { return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" :: "r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" :: "r"(d), "r"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void st8(uint32_t t, uint32_t v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" :: "r"(t), "r"(v) : "memory");
}

#define MT(t, u0, nn) mma_ts(tb + 64u * ((t) + (u0) - 2), at + 8u * ((t) - 1), sdesc(b0 + ((u0) - 1) * 2048), idesc(128, 64 * (nn)), 1u)
#define MS(t, u0, nn) mma_ss(tb + 64u * ((t) + (u0) - 2), sdesc(a0 + ((t) - 1) * 4096), sdesc(b0 + ((u0) - 1) * 2048), idesc(128, 64 * (nn)), 1u)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ uint4 ldg16(const void* p) { uint4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)); return v; }
__device__ __forceinline__ bool elect_one() { uint32_t p; asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(p)); return p != 0; }

constexpr int AS = 4096, BS = 2048;
// mode 3: the kernel's data movement on top of mode 2.  TMA: a producer warp streams TMAB bytes per k-step from global memory (L2 resident)
// into a 4-deep ring that the MMAs do NOT read (pure bandwidth load on the shared-memory write side); LDG: every loader thread
// reads 128 B per k-step from global memory before its tcgen05.st.
template <int TMAB, int LDG>
__global__ void __launch_bounds__(256) probe3(int iters, long long* clk, int* sink, const uint8_t* __restrict__ gsrc, int rnd) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t* sA = sm;
    uint8_t* sB = sm + 7 * AS;
    uint8_t* ring = sm + 7 * (AS + BS);      // 4 x 28 KiB
    __shared__ uint64_t done, afull[2], aempty[2], rfull[4], rempty[4];
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 7 * (AS + BS); i += 256) sm[i] = rnd ? (uint8_t)(((unsigned)i * 2654435761u) >> 13) : (uint8_t)((i * 7 + 3) & 3);
    if (tid == 0) {
        mbar_init(&done, 1);
        for (int b = 0; b < 2; b++) { mbar_init(&afull[b], 4); mbar_init(&aempty[b], 1); }
        for (int b = 0; b < 4; b++) { mbar_init(&rfull[b], 1); mbar_init(&rempty[b], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tbase)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tbase;
    const uint8_t* mysrc = gsrc + (size_t)(blockIdx.x % 16) * (1 << 20);
    long long t0 = clock64();
    if (warp == 0) {
        if (TMAB > 0)
            for (int n = 0; n < iters; n++) {
                const int s = n & 3;
                mbar_wait(&rempty[s], ((n >> 2) & 1) ^ 1);
                if (elect_one()) { expect_tx(&rfull[s], TMAB); bulk_g2s(ring + s * 28672, mysrc + (size_t)(n & 31) * 28672, TMAB, &rfull[s]); }
                __syncwarp();
            }
    } else if (warp == 1) {
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        for (int n = 0; n < iters; n++) {
            const uint32_t ab = n & 1, at = tb + 448 + 32 * ab;
            if (TMAB > 0) mbar_wait(&rfull[n & 3], (n >> 2) & 1);
            mbar_wait(&afull[ab], (n >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                MT(1, 1, 4); MT(1, 5, 3); MT(2, 1, 4); MT(2, 5, 2); MT(3, 1, 4); MT(3, 5, 1); MT(4, 1, 4);
                MS(5, 1, 3); MS(6, 1, 2); MS(7, 1, 1);
                commit(&aempty[ab]);
                if (TMAB > 0) commit(&rempty[n & 3]);
            }
            __syncwarp();
        }
        if (elect_one()) commit(&done);
        __syncwarp();
    } else if (warp >= 4) {
        const uint32_t tl = tb + ((uint32_t)((warp & 3) * 32) << 16) + 448;
        const uint8_t* lsrc = mysrc + (size_t)((warp & 3) * 32 + lane) * 16;
        uint4 r[8];
        for (int n = 0; n < iters; n++) {
            const uint32_t ab = n & 1;
            if (LDG) { for (int x = 0; x < 8; x++) r[x] = ldg16(lsrc + (size_t)((n * 8 + x) & 255) * 2048); }
            else { for (int x = 0; x < 8; x++) { unsigned h = rnd ? ((unsigned)(n * 8 + x + tid * 977) * 2654435761u) : (unsigned)n; r[x] = make_uint4(h, h * 3u + 1u, h ^ 0x5bd1e995u, h * 7u); } }
            mbar_wait(&aempty[ab], ((n >> 1) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int t = 0; t < 4; t++)
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" :: "r"(tl + 32 * ab + 8 * t),
                             "r"(r[2 * t].x), "r"(r[2 * t].y), "r"(r[2 * t].z), "r"(r[2 * t].w), "r"(r[2 * t + 1].x), "r"(r[2 * t + 1].y), "r"(r[2 * t + 1].z), "r"(r[2 * t + 1].w) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp(); if (lane == 0) mbar_arrive(&afull[ab]);
        }
    }
    mbar_wait(&done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    long long t1 = clock64();
    if (tid == 0) clk[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(512u) : "memory");
}

template <int MODE, int ARRIVE_ALL>
__global__ void __launch_bounds__(192) probe(int iters, long long* clk, int* sink) {
    extern __shared__ __align__(1024) uint8_t sm[];
    uint8_t* sA = sm;               // 7 digits x 4 KiB
    uint8_t* sB = sm + 7 * AS;      // 7 digits x 2 KiB
    __shared__ uint64_t done, afull[2], aempty[2];
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 7 * (AS + BS); i += 192) sm[i] = (uint8_t)((i * 7 + 3) & 3);
    if (tid == 0) {
        mbar_init(&done, 1);
        for (int b = 0; b < 2; b++) { mbar_init(&afull[b], ARRIVE_ALL ? 128 : 4); mbar_init(&aempty[b], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tbase)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tbase;
    if (MODE >= 1 && warp >= 2) {       // initial contents of both A buffers
        const uint32_t tl = tb + ((uint32_t)((warp & 3) * 32) << 16) + 448;
        for (int c = 0; c < 8; c++) st8(tl + 8 * c, 0x01020301u);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    long long t0 = clock64();
    if (warp == 1 && lane == 0) {
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        for (int n = 0; n < iters; n++) {
            const uint32_t ab = n & 1, at = tb + 448 + 32 * ab;
            if (MODE == 2) { mbar_wait(&afull[ab], (n >> 1) & 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
            if (MODE == 0) { MS(1, 1, 4); MS(1, 5, 3); MS(2, 1, 4); MS(2, 5, 2); MS(3, 1, 4); MS(3, 5, 1); MS(4, 1, 4); }
            else           { MT(1, 1, 4); MT(1, 5, 3); MT(2, 1, 4); MT(2, 5, 2); MT(3, 1, 4); MT(3, 5, 1); MT(4, 1, 4); }
            MS(5, 1, 3); MS(6, 1, 2); MS(7, 1, 1);
            if (MODE == 2) commit(&aempty[ab]);
        }
        commit(&done);
    } else if (MODE == 2 && warp >= 2) {
        const uint32_t tl = tb + ((uint32_t)((warp & 3) * 32) << 16) + 448;
        for (int n = 0; n < iters; n++) {
            const uint32_t ab = n & 1;
            mbar_wait(&aempty[ab], ((n >> 1) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int t = 0; t < 4; t++) st8(tl + 32 * ab + 8 * t, 0x01020301u + n);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (ARRIVE_ALL) mbar_arrive(&afull[ab]);
            else { __syncwarp(); if (lane == 0) mbar_arrive(&afull[ab]); }
        }
    }
    mbar_wait(&done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    long long t1 = clock64();
    if (tid == 0) clk[blockIdx.x] = t1 - t0;
    if (warp >= 2) {
        uint32_t v;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(tb + ((uint32_t)((warp & 3) * 32) << 16)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (v == 0x7fffffffu) sink[blockIdx.x * 192 + tid] = (int)v;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(512u) : "memory");
}

template <int MODE, int AA>
static int run(const char* name, int sms) {
    long long* clk; int* sink;
    CK(cudaMalloc(&clk, sizeof(long long) * sms)); CK(cudaMalloc(&sink, sizeof(int) * 192 * sms));
    const int smem = 7 * (AS + BS), iters = 4000;
    CK(cudaFuncSetAttribute(probe<MODE, AA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe<MODE, AA><<<sms, 192, smem>>>(200, clk, sink);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    probe<MODE, AA><<<sms, 192, smem>>>(iters, clk, sink);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    long long h[4]; CK(cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost));
    double tops = 2.0 * 28 * 128 * 64 * 32 * (double)iters * sms / (ms * 1e-3) / 1e12;
    printf("%-58s %.3f ms, %.0f clk per k-step (SM 0), %.2f POP/s\n", name, ms, (double)h[0] / iters, tops / 1e3);
    cudaFree(clk); cudaFree(sink);
    return 0;
}
template <int TMAB, int LDG>
static int run3(const char* name, int sms, int rnd = 0) {
    long long* clk; int* sink; uint8_t* g;
    CK(cudaMalloc(&clk, sizeof(long long) * sms)); CK(cudaMalloc(&sink, sizeof(int) * 256 * sms)); CK(cudaMalloc(&g, 17 << 20)); CK(cudaMemset(g, 1, 17 << 20)); if (rnd) { std::vector<uint8_t> hb(17 << 20); unsigned x = 12345; for (auto& b : hb) { x = x * 1664525u + 1013904223u; b = (uint8_t)(x >> 24); } CK(cudaMemcpy(g, hb.data(), hb.size(), cudaMemcpyHostToDevice)); }
    const int smem = 7 * (AS + BS) + 4 * 28672, iters = 4000;
    CK(cudaFuncSetAttribute(probe3<TMAB, LDG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe3<TMAB, LDG><<<sms, 256, smem>>>(200, clk, sink, g, rnd);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    probe3<TMAB, LDG><<<sms, 256, smem>>>(iters, clk, sink, g, rnd);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    long long h[4]; CK(cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost));
    double tops = 2.0 * 28 * 128 * 64 * 32 * (double)iters * sms / (ms * 1e-3) / 1e12;
    printf("%-58s %.3f ms, %.0f clk per k-step (SM 0), %.2f POP/s\n", name, ms, (double)h[0] / iters, tops / 1e3);
    cudaFree(clk); cudaFree(sink); cudaFree(g);
    return 0;
}
int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    printf("%s, %d SMs\n", pr.name, pr.multiProcessorCount);
    if (run<0, 0>("mode 0: all A digits from shared memory", pr.multiProcessorCount)) return 1;
    if (run<1, 0>("mode 1: A digits 1..4 from TMEM (static)", pr.multiProcessorCount)) return 1;
    if (run<2, 1>("mode 2: + loader handshake, 128 arrivals per buffer", pr.multiProcessorCount)) return 1;
    if (run<2, 0>("mode 2: + loader handshake, 4 arrivals per buffer", pr.multiProcessorCount)) return 1;
    if (run3<0, 0>("mode 3: elect_one issue, no TMA, no LDG", pr.multiProcessorCount)) return 1;
    if (run3<0, 1>("mode 3: + 16 KiB LDG per k-step", pr.multiProcessorCount)) return 1;
    if (run3<26624, 0>("mode 3: + 26 KiB bulk copy per k-step", pr.multiProcessorCount)) return 1;
    if (run3<26624, 1>("mode 3: + 26 KiB bulk copy + 16 KiB LDG per k-step", pr.multiProcessorCount)) return 1;
    if (run3<14336, 1>("mode 3: + 14 KiB bulk copy + 16 KiB LDG per k-step", pr.multiProcessorCount)) return 1;
    if (run3<0, 0>("mode 3: no TMA, no LDG, RANDOM operand bytes", pr.multiProcessorCount, 1)) return 1;
    if (run3<26624, 1>("mode 3: 26 KiB bulk copy + 16 KiB LDG, RANDOM operand bytes", pr.multiProcessorCount, 1)) return 1;
    if (run3<0, 0>("mode 3: no TMA, no LDG, constant bytes again", pr.multiProcessorCount, 0)) return 1;
    return 0;
}
