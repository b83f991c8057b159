// EXPERIMENTAL (IBO_FLAG_INT8 / IBO_INT8=1, wide batches only): K2's FP64 triangular GEMM V = W K* emulated on the INT8 tensor
// cores (tcgen05.mma kind::i8, INT32 accumulators in TMEM) with an Ozaki-style slicing, and the K1 / model-side kernels that
// feed it.  The B200's INT8 tensor rate is 4.5 POP/s (tools/research/umma_i8_probe.cu: 4.49 measured) against 37 TF/s of
// DMMA; 28 slice products per FP64 product leave a ~4x higher ceiling for sigma^2 at the parity bound
// (tools/research/ozaki_int8_study.py: 7 slices of 7 bits -> sigma^2 within 5e-12 relative at N = 1024..2048).
//
//   W  (row r)       = 2^e_r * sum_{t=1..7} 2^(-7t) A_t      A_t balanced digits in [-64, 64] of rint(w 2^(49 - e_r))
//   K* (any element) =         sum_{u=1..7} 2^(-7u) B_u      B_u in [0, 127], digits of rint(k 2^49) (k / sf2 lies in [0, 1])
//   V = 2^e_r * sum_{g=2..8} 2^(-7g) D_g,   D_g = sum_{t+u=g} A_t B_u^T  (exact INT32; pairs with t + u > 8 are dropped)
//
// One MMA covers several pairs: the B slices of a candidate tile are consecutive 64-row blocks of one K-major operand, so
// A_t x [B_u0 .. B_u0+n-1] is a single 128 x 64n x 32 instruction whose 64-column output blocks land on the accumulators of
// groups g = t+u0 .. t+u0+n-1 (TMEM columns 64 (g - 2)): 10 instructions per 32-deep k-step instead of 28.
// sigma^2 only needs sum_r V_r^2 per candidate; the posterior mean is taken as k* . alpha (alpha = W^T W Y, plain FP64 dot
// products inside K1, same partial-sum planes as K2's V . beta), which K3 consumes unchanged.
//
// Operand layout (both in HBM and in shared memory): the no-swizzle K-major canonical layout of the UMMA shared-memory
// descriptor -- 8 x 16-byte core matrices, the two 16-byte k chunks of a row group 128 B apart (LBO), row groups 256 B apart
// (SBO) -- so that a pipeline stage is two contiguous bulk copies (28 KiB of W slices + 14 KiB of K* slices).
#pragma once

namespace ibo {

constexpr int I8_S = 7;                          // slices per operand
constexpr int I8_FRAC = 49;                      // 7 * I8_S fixed-point bits
constexpr int I8_NT = 64;                        // candidates per tile
constexpr int I8_A_SLICE = 128 * 32;             // bytes: 128 rows x 32 k
constexpr int I8_B_SLICE = I8_NT * 32;           // bytes: 64 candidates x 32 k
constexpr int I8_A_STAGE = I8_S * I8_A_SLICE;    // 28672
constexpr int I8_B_STAGE = I8_S * I8_B_SLICE;    // 14336
constexpr int I8_STAGES = 4;
constexpr int I8_THREADS = 192;                  // warp 0: bulk-copy producer, warp 1: TMEM owner + MMA issuer, warps 2-5: epilogue
constexpr int I8_SMEM = I8_STAGES * (I8_A_STAGE + I8_B_STAGE) + 4 * 64 * 8 + 16 * 8;

// byte offset of element (row r, k) in a [rows x 32] canonical tile
__host__ __device__ inline uint32_t i8_canon(int r, int k) { return (uint32_t)((((r >> 3) * 2 + (k >> 4)) << 7) + ((r & 7) << 4) + (k & 15)); }
// first k-step of row-block i in the packed W slices (row-block i owns 4 (i + 1) steps of I8_A_STAGE bytes)
__host__ __device__ inline size_t wi8_base(int i) { return (size_t)4 * ((size_t)i * (i + 1) / 2) * I8_A_STAGE; }

// ---- tcgen05 / TMEM primitives (encodings checked on the device by tools/research/umma_i8_probe.cu) ----------------------
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t addr) {        // LBO = 128 B, SBO = 256 B, version 1, no swizzle
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ constexpr uint32_t umma_idesc_i8(int M, int N) { // S32 accumulator, signed 8-bit A / B, K-major
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {               // arrives on `bar` once every MMA issued so far has completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// model side: per-row power-of-two scale of W and the seven slices, packed per (row-block, k-step)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) i8_rowscale_kernel(const double* __restrict__ W, int Np, int N, double sf2, int headroom,
                                                          double* __restrict__ rowScale) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= Np) return;
    double mx = 0.0, sum = 0.0;
    for (int k = lane; k <= row; k += 32) {
        const double w = W[(size_t)row * Np + k];
        mx = fmax(mx, fabs(w));
        if (k < N) sum += w;
    }
    for (int o = 16; o; o >>= 1) { mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o)); sum += __shfl_xor_sync(0xffffffffu, sum, o); }
    // 2^e with |w| / 2^e < 2^(1 - headroom) for the whole row (balanced digits: the top one must fit its int8).
    //   7-bit digits: headroom 2 (|top| <= 64); slot 1 = scale * sf2 (K* is sliced as k / sf2 in [0, 1]), multiplied back by the epilogue
    //   8-bit digits: headroom 3 (|top| <= 64); K* is sliced as (k / sf2 - 1/2) / 2 in [-1/4, 1/4]: slot 1 = 2 * scale * sf2 and
    //                 slot 2 = sf2 / 2 * sum_{k < N} W[row][k], the constant the shift leaves behind
    if (lane == 0) {
        const double sc = mx > 0.0 ? scalbn(1.0, ilogb(mx) + headroom) : 1.0;
        rowScale[row] = sc;
        rowScale[Np + row] = headroom == 2 ? sc * sf2 : 2.0 * sc * sf2;
        rowScale[2 * Np + row] = 0.5 * sf2 * sum;
    }
}

// BITS: digit width, 7 (validated) or 8 (IBO_FLAG_INT8_D8); S: digits per operand, 7 or (8-bit digits only, IBO_FLAG_INT8_S6) 6.
// The 6-digit variant keeps the 7-slice stage stride and simply leaves the last slice slot unused.
template <int BITS, int S>
__global__ void __launch_bounds__(256) i8_slice_w_kernel(const double* __restrict__ W, const double* __restrict__ rowScale, int Np,
                                                         uint8_t* __restrict__ Wi8) {
    constexpr int FR = BITS * S;                             // fixed-point bits: 49, 56 or 48
    constexpr long long HALF = 1ll << (BITS - 1), MASK = (1ll << BITS) - 1;
    const int j = blockIdx.x, i = blockIdx.y;                 // k-step, row-block
    if (j >= (i + 1) * 4) return;
    const int r = threadIdx.x & 127, c16 = threadIdx.x >> 7;  // row of the block, 16-byte k chunk of the step
    const int row = i * 128 + r, k0 = j * 32 + c16 * 16;
    const double inv = 1.0 / rowScale[row];                   // power of two: exact
    long long q[16];
#pragma unroll
    for (int kk = 0; kk < 16; kk++) {
        const double w = W[(size_t)row * Np + k0 + kk];        // exact zeros above the diagonal
        q[kk] = __double2ll_rn(w * inv * (double)(1ll << FR));  // round to nearest; |w| * inv < 1/2 (7-bit) or 1/4 (8-bit digits)
    }
    uint8_t* dst = Wi8 + wi8_base(i) + (size_t)j * I8_A_STAGE + i8_canon(r, c16 * 16);
    // balanced digits, least significant first: d_t in [-2^(BITS-1), 2^(BITS-1) - 1] for t = 7 .. 2, the rest (|d_1| <= 64) is the top digit.
    // Zero-mean digits make the dropped slice pairs (t + u > 8) a zero-mean error that grows like sqrt(N), not N.
#pragma unroll
    for (int t = S; t >= 1; t--) {
        uint32_t w4[4];
#pragma unroll
        for (int v = 0; v < 4; v++) {
            uint32_t word = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int kk = v * 4 + b;
                long long dgt;
                if (t > 1) { dgt = ((q[kk] + HALF) & MASK) - HALF; q[kk] = (q[kk] - dgt) >> BITS; }
                else dgt = q[kk];
                word |= ((uint32_t)(int)dgt & 0xffu) << (8 * b);
            }
            w4[v] = word;
        }
        *reinterpret_cast<uint4*>(dst + (size_t)(t - 1) * I8_A_SLICE) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
    }
}

// ---------------------------------------------------------------------------------------------
// K1 (int8): one CTA = (64-candidate tile T, row-block i); thread = (candidate c, k-step of the block).  Kernel values from
// direct differences, sliced from a 49-bit fixed-point image, 16 bytes (one core-matrix row) per store; the partial dot
// products k* . alphaY / k* . alpha1 of the block go to planes 1 / 2 of `part` (what K2 writes as V . beta).
// ---------------------------------------------------------------------------------------------
template <int KC, int DMAX, int BITS, int S>
__global__ void __launch_bounds__(256) kstar_i8_kernel(const double* __restrict__ Xt, const double* __restrict__ cand,
                                                       const double* __restrict__ inv_theta, const double* __restrict__ center,
                                                       const double* __restrict__ alphaY, const double* __restrict__ alpha1,
                                                       uint8_t* __restrict__ Ki8, double* __restrict__ part,
                                                       int N, int d, int nb, long M, long m0, long Mpad, double sf2, int want_p1) {
    // DMAX >= d (even): the dimension loops have a compile-time length, the thread's candidate lives in registers and the
    // training rows (zero padded to DMAX) are read with 16-byte broadcast loads
    __shared__ __align__(16) double sX[128 * DMAX];
    __shared__ double sAy[128], sA1[128], red[512];
    const int T = blockIdx.x, i = blockIdx.y, tid = threadIdx.x;
    for (int idx = tid; idx < 128 * DMAX; idx += 256) {
        const int r = idx / DMAX, j = idx - r * DMAX;
        sX[idx] = j < d ? Xt[(size_t)(i * 128 + r) * d + j] : 0.0;
    }
    if (tid < 128) { sAy[tid] = alphaY[i * 128 + tid]; sA1[tid] = alpha1[i * 128 + tid]; }
    const int c = tid & 63, kg = tid >> 6;
    double xc[DMAX];
    {
        long cg = m0 + (long)T * 64 + c;
        if (cg >= M) cg = M - 1;
#pragma unroll
        for (int j = 0; j < DMAX; j++) xc[j] = j < d ? cand[(size_t)cg * d + j] * inv_theta[j] - center[j] : 0.0;
    }
    __syncthreads();
    double sy = 0.0, s1 = 0.0;
    uint8_t* dst0 = Ki8 + ((size_t)T * (nb * 4) + (size_t)i * 4 + kg) * I8_B_STAGE;
#pragma unroll 1
    for (int c16 = 0; c16 < 2; c16++) {
        unsigned long long q[16];
#pragma unroll
        for (int kk = 0; kk < 16; kk++) {
            const int k = kg * 32 + c16 * 16 + kk;
            const double2* xr = reinterpret_cast<const double2*>(sX + k * DMAX);
            double r2a = 0.0, r2b = 0.0;
#pragma unroll
            for (int j2 = 0; j2 < DMAX / 2; j2++) {
                const double2 x2 = xr[j2];
                const double da = x2.x - xc[2 * j2], db = x2.y - xc[2 * j2 + 1];
                r2a = fma(da, da, r2a);
                r2b = fma(db, db, r2b);
            }
            const double v = (i * 128 + k) < N ? cov_r2_t<KC>(1.0, r2a + r2b) : 0.0;   // in [0, 1]
            sy = fma(v, sAy[k], sy);
            s1 = fma(v, sA1[k], s1);
            if (BITS == 7) {
                unsigned long long qq = (unsigned long long)__double2ll_rn(v * 562949953421312.0);      // 2^49, round to nearest
                q[kk] = qq > 562949953421311ull ? 562949953421311ull : qq;                 // v == 1 (candidate on a training point)
            } else {
                // 8-bit digits: (v - 1/2) / 2 in [-1/4, 1/4] at 56 fractional bits, signed; rows beyond N contribute nothing
                q[kk] = (i * 128 + k) < N ? (unsigned long long)__double2ll_rn((v - 0.5) * (double)(1ll << (8 * S - 1))) : 0ull;   // 2^55 (2^47 with 6 digits)
            }
        }
        uint8_t* dst = dst0 + i8_canon(c, c16 * 16);
        if (BITS == 7) {
#pragma unroll
            for (int t = 1; t <= I8_S; t++) {
                uint32_t w4[4];
#pragma unroll
                for (int v = 0; v < 4; v++) {
                    uint32_t word = 0;
#pragma unroll
                    for (int b = 0; b < 4; b++) word |= (uint32_t)((q[v * 4 + b] >> (I8_FRAC - 7 * t)) & 127ull) << (8 * b);
                    w4[v] = word;
                }
                *reinterpret_cast<uint4*>(dst + (size_t)(t - 1) * I8_B_SLICE) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
            }
        } else {
            // balanced base-256 digits, least significant first (as the W slices)
#pragma unroll
            for (int t = S; t >= 1; t--) {
                uint32_t w4[4];
#pragma unroll
                for (int v = 0; v < 4; v++) {
                    uint32_t word = 0;
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        long long qs = (long long)q[v * 4 + b], dgt;
                        if (t > 1) { dgt = ((qs + 128) & 255) - 128; q[v * 4 + b] = (unsigned long long)((qs - dgt) >> 8); }
                        else dgt = qs;
                        word |= ((uint32_t)(int)dgt & 0xffu) << (8 * b);
                    }
                    w4[v] = word;
                }
                *reinterpret_cast<uint4*>(dst + (size_t)(t - 1) * I8_B_SLICE) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
            }
        }
    }
    red[kg * 64 + c] = sy;
    red[256 + kg * 64 + c] = s1;
    __syncthreads();
    if (tid < 64) {
        const size_t plane = (size_t)nb * Mpad;
        const size_t o = (size_t)i * Mpad + (size_t)T * 64 + tid;
        part[plane + o] = sf2 * (((red[tid] + red[64 + tid]) + red[128 + tid]) + red[192 + tid]);
        if (want_p1) part[2 * plane + o] = sf2 * (((red[256 + tid] + red[320 + tid]) + red[384 + tid]) + red[448 + tid]);
    }
}

// ---------------------------------------------------------------------------------------------
// K2 (int8): CTA = (row-block group g of G, 64-candidate tile T).  Row-blocks are dealt to the groups in snake order
// (work ~ i + 1), each one a full sweep over its 4 (i + 1) k-steps into the seven group accumulators, then the epilogue
// warps assemble V in FP64 and reduce sum_r V_r^2 per candidate in a fixed order.
// ---------------------------------------------------------------------------------------------
// NG accumulator groups: 7 (t + u <= 8, the validated default) or 8 (t + u <= 9: IBO_FLAG_INT8_G9); BITS: digit width 7 (validated) or 8
// (IBO_FLAG_INT8_D8: same 28 products, operands rounded at 2^-56 instead of 2^-49; K* sliced as (k - 1/2) / 2 with signed digits, the
// constant the shift leaves behind comes back as rowConst).  The non-default variants are written for the next round and untested.
template <int NG, int BITS, int S>
__global__ void __launch_bounds__(I8_THREADS, 1) trigemm_i8_kernel(const uint8_t* __restrict__ Wi8, const uint8_t* __restrict__ Ki8,
                                                                    const double* __restrict__ rowScaleSf, double* __restrict__ part,
                                                                    int nb, long Mpad) {
    // rowScaleSf[row]: the factor that turns the assembled integer sum into v; rowScaleSf[Np + row] (8-bit digits): the additive constant
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                                          // [stage][7][128 x 32]
    uint8_t* sB = smem + I8_STAGES * I8_A_STAGE;                 // [stage][7][64 x 32]
    double* red = reinterpret_cast<double*>(smem + I8_STAGES * (I8_A_STAGE + I8_B_STAGE));     // [4][64]
    uint64_t* full = reinterpret_cast<uint64_t*>(red + 4 * 64);
    uint64_t* empty = full + I8_STAGES;
    uint64_t* tfull = empty + I8_STAGES;
    uint64_t* tempty = tfull + 1;
    uint32_t* tbase = reinterpret_cast<uint32_t*>(tempty + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x, G = gridDim.x, T = blockIdx.y;
    if (tid == 0) {
        for (int s = 0; s < I8_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tfull, 1);
        mbar_init(tempty, 128);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tbase)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = *tbase;
    const int rounds = (nb + G - 1) / G;

    if (warp == 0) {
        // ---------------- producer: two contiguous bulk copies per k-step ----------------
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            const uint8_t* Bb = Ki8 + (size_t)T * (nb * 4) * I8_B_STAGE;
            for (int r = 0; r < rounds; r++) {
                const int idx = r * G + ((r & 1) ? (G - 1 - g) : g);
                if (idx >= nb) continue;
                const int i = nb - 1 - idx;
                const uint8_t* Ab = Wi8 + wi8_base(i);
                for (int j = 0; j < (i + 1) * 4; j++) {
                    mbar_wait(&empty[s], ph ^ 1);
                    // S slices of each operand (the stage stride stays that of 7 slices)
                    mbar_arrive_expect_tx(&full[s], S * (I8_A_SLICE + I8_B_SLICE));
                    bulk_g2s(sA + s * I8_A_STAGE, Ab + (size_t)j * I8_A_STAGE, S * I8_A_SLICE, &full[s]);
                    bulk_g2s(sB + s * I8_B_STAGE, Bb + (size_t)j * I8_B_STAGE, S * I8_B_SLICE, &full[s]);
                    if (++s == I8_STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (one lane) ----------------
        if (lane == 0) {
            int s = 0; uint32_t ph = 0; int rb = 0;
            for (int r = 0; r < rounds; r++) {
                const int idx = r * G + ((r & 1) ? (G - 1 - g) : g);
                if (idx >= nb) continue;
                const int i = nb - 1 - idx;
                mbar_wait(tempty, (uint32_t)(rb & 1) ^ 1);          // the epilogue has drained the previous row-block's accumulators
                tc_fence_after();
                for (int j = 0; j < (i + 1) * 4; j++) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(sA + s * I8_A_STAGE), b0 = smem_u32(sB + s * I8_B_STAGE);
                    const uint32_t first = j == 0 ? 0u : 1u;
                    // slice t of W against B slices u0 .. u0+n-1: columns 64 (t + u0 - 2), N = 64 n
#define I8_MMA(t, u0, n, acc) umma_i8(tb + 64u * ((t) + (u0) - 2), umma_smem_desc(a0 + ((t) - 1) * I8_A_SLICE), \
                                      umma_smem_desc(b0 + ((u0) - 1) * I8_B_SLICE), umma_idesc_i8(128, 64 * (n)), acc)
                    if (S == 6) {
                        // six 8-bit digits, 21 pairs, t + u <= 7: groups 2..7 on columns 0..383
                        I8_MMA(1, 1, 4, first); I8_MMA(1, 5, 2, first);
                        I8_MMA(2, 1, 4, 1u);    I8_MMA(2, 5, 1, 1u);
                        I8_MMA(3, 1, 4, 1u);
                        I8_MMA(4, 1, 3, 1u);
                        I8_MMA(5, 1, 2, 1u);
                        I8_MMA(6, 1, 1, 1u);
                    } else {
                    I8_MMA(1, 1, 4, first); I8_MMA(1, 5, 3, first);          // the two t = 1 instructions touch groups 2..8 first
                    if (NG == 7) {
                        I8_MMA(2, 1, 4, 1u);    I8_MMA(2, 5, 2, 1u);
                        I8_MMA(3, 1, 4, 1u);    I8_MMA(3, 5, 1, 1u);
                        I8_MMA(4, 1, 4, 1u);
                        I8_MMA(5, 1, 3, 1u);
                        I8_MMA(6, 1, 2, 1u);
                        I8_MMA(7, 1, 1, 1u);
                    } else {
                        // 34 pairs, t + u <= 9; group 9 (columns 448..511) is first touched by (t = 2, u = 7)
                        I8_MMA(2, 1, 4, 1u);    I8_MMA(2, 5, 2, 1u);    I8_MMA(2, 7, 1, first);
                        I8_MMA(3, 1, 4, 1u);    I8_MMA(3, 5, 2, 1u);
                        I8_MMA(4, 1, 4, 1u);    I8_MMA(4, 5, 1, 1u);
                        I8_MMA(5, 1, 4, 1u);
                        I8_MMA(6, 1, 3, 1u);
                        I8_MMA(7, 1, 2, 1u);
                    }
                    }
#undef I8_MMA
                    umma_commit(&empty[s]);                          // the stage is free once these MMAs have read it
                    if (++s == I8_STAGES) { s = 0; ph ^= 1; }
                }
                umma_commit(tfull);                                  // accumulators of row-block i complete
                rb++;
            }
        }
    } else {
        // ---------------- epilogue: warps 2..5 own TMEM lanes 32 (warp % 4) .. + 31 ----------------
        const int qd = warp & 3, et = (warp - 2) * 32 + lane;
        int rb = 0;
        for (int r = 0; r < rounds; r++) {
            const int idx = r * G + ((r & 1) ? (G - 1 - g) : g);
            if (idx >= nb) continue;
            const int i = nb - 1 - idx;
            const double rs = rowScaleSf[i * 128 + qd * 32 + lane];
            const double rc = BITS == 8 ? rowScaleSf[(size_t)nb * 128 + i * 128 + qd * 32 + lane] : 0.0;
            mbar_wait(tfull, (uint32_t)(rb & 1));
            tc_fence_after();
#pragma unroll 1
            for (int c8 = 0; c8 < 8; c8++) {
                uint32_t D[NG][8];
#pragma unroll
                for (int gq = 0; gq < NG; gq++) tmem_ld8_nowait(tb + ((uint32_t)(qd * 32) << 16) + 64u * gq + 8u * c8, D[gq]);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    double v;
                    if (BITS == 7) {
                        const long long hi = (long long)(int)D[0][e] * 16384 + (long long)(int)D[1][e] * 128 + (long long)(int)D[2][e];
                        const long long mid = (long long)(int)D[3][e] * 16384 + (long long)(int)D[4][e] * 128 + (long long)(int)D[5][e];
                        // group 8 at 2^-56 (and group 9 at 2^-63 when it is kept)
                        constexpr int G8 = NG >= 7 ? 6 : NG - 1;            // (NG == 6 exists only with 8-bit digits; keeps the indices in range)
                        const double lo = NG == 7 ? (double)(int)D[G8][e] * 1.387778780781445675529539585113525390625e-17
                                                  : (double)((long long)(int)D[G8][e] * 128 + (long long)(int)D[NG - 1][e]) * 1.084202172485504434007452800869941711425781e-19;
                        v = fma((double)hi, 3.7252902984619140625e-09,                     // 2^-28
                                fma((double)mid, 1.7763568394002504646778106689453125e-15, // 2^-49
                                    lo));
                        v *= rs;
                    } else {
                        // base-256 digits: groups 2..4 at 2^-32, 5..7 at 2^-56, 8 at 2^-64 (9 at 2^-72)
                        const long long hi = (long long)(int)D[0][e] * 65536 + (long long)(int)D[1][e] * 256 + (long long)(int)D[2][e];
                        const long long mid = (long long)(int)D[3][e] * 65536 + (long long)(int)D[4][e] * 256 + (long long)(int)D[5][e];
                        const double lo = NG == 6 ? 0.0
                                        : NG == 7 ? (double)(int)D[NG - 1][e] * 5.42101086242752217003726400434970855712890625e-20
                                                  : (double)((long long)(int)D[NG - 2][e] * 256 + (long long)(int)D[NG - 1][e]) * 2.1175823681357508476708062516990986740112305e-22;
                        v = fma((double)hi, 2.3283064365386962890625e-10,                  // 2^-32
                                fma((double)mid, 1.387778780781445675529539585113525390625e-17,   // 2^-56
                                    lo));
                        v = fma(v, rs, rc);
                    }
                    double sq = v * v;
#pragma unroll
                    for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
                    if (lane == 0) red[qd * 64 + c8 * 8 + e] = sq;
                }
            }
            tc_fence_before();
            mbar_arrive(tempty);                                     // 128 arrivals: the accumulators may be overwritten
            named_bar_sync(1, 128);
            if (et < 64) {
                // TMEM lane quarter q holds rows 32 q .. 32 q + 31: ascending row order
                part[(size_t)i * Mpad + (size_t)T * 64 + et] = ((red[et] + red[64 + et]) + red[128 + et]) + red[192 + et];
            }
            named_bar_sync(1, 128);
            rb++;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(512u) : "memory");
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int NG, int BITS, int S = 7>
static cudaError_t i8_set_k2_attrs() {
    cudaError_t e = cudaFuncSetAttribute(trigemm_i8_kernel<NG, BITS, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, I8_SMEM);
    // K1 CTAs of the next chunk are meant to run next to a resident K2 CTA (174 KiB): ask for the largest shared-memory
    // carve-out so that the 21 KiB a smaller configuration would leave do not limit them to one per SM
    if (e == cudaSuccess) e = cudaFuncSetAttribute(trigemm_i8_kernel<NG, BITS, S>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    return e;
}

// Every K1 variant asks for the same (largest) shared-memory carve-out as K2: an SM has one carve-out at a time, and a K1 CTA that
// configured it smaller would keep the 174 KiB K2 CTA of the other stream off that SM until it drains (and the other way round).
template <int KC, int DMAX, int BITS, int S = 7>
static cudaError_t i8_set_k1_attr() {
    return cudaFuncSetAttribute(kstar_i8_kernel<KC, DMAX, BITS, S>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}
template <int DMAX>
static cudaError_t i8_set_k1_attrs_d() {
    cudaError_t e = i8_set_k1_attr<0, DMAX, 7>();
    if (e == cudaSuccess) e = i8_set_k1_attr<1, DMAX, 7>();
    if (e == cudaSuccess) e = i8_set_k1_attr<2, DMAX, 7>();
    if (e == cudaSuccess) e = i8_set_k1_attr<0, DMAX, 8>();
    if (e == cudaSuccess) e = i8_set_k1_attr<1, DMAX, 8>();
    if (e == cudaSuccess) e = i8_set_k1_attr<2, DMAX, 8>();
    if (e == cudaSuccess) e = i8_set_k1_attr<0, DMAX, 8, 6>();
    if (e == cudaSuccess) e = i8_set_k1_attr<1, DMAX, 8, 6>();
    if (e == cudaSuccess) e = i8_set_k1_attr<2, DMAX, 8, 6>();
    return e;
}
static cudaError_t i8_set_k1_attrs() {
    cudaError_t e = i8_set_k1_attrs_d<2>();
    if (e == cudaSuccess) e = i8_set_k1_attrs_d<4>();
    if (e == cudaSuccess) e = i8_set_k1_attrs_d<6>();
    if (e == cudaSuccess) e = i8_set_k1_attrs_d<8>();
    if (e == cudaSuccess) e = i8_set_k1_attrs_d<12>();
    if (e == cudaSuccess) e = i8_set_k1_attrs_d<16>();
    if (e == cudaSuccess) e = i8_set_k1_attrs_d<24>();
    if (e == cudaSuccess) e = i8_set_k1_attrs_d<32>();
    return e;
}

// digit modes: 0 = 7 x 7-bit (validated), 1 = 7 x 8-bit (IBO_FLAG_INT8_D8), 2 = 6 x 8-bit, 21 slice pairs (IBO_FLAG_INT8_S6)
// builds (once per model state) the packed digit slices of W for the mode, the row scales / constants, alpha
static int ensure_i8(ibo_model* m, int mode) {
    const int Np = m->Np, nb = m->nb;
    cudaStream_t st = m->stream;
    if (!m->dAlphaY || !(m->i8Valid || m->i8Valid8 || m->i8Valid6)) {
        for (double** p : {&m->dAlphaY, &m->dAlpha1}) if (*p) { pool_free(*p); *p = nullptr; }
        IBO_CUDA_TRY(pool_malloc((void**)&m->dAlphaY, sizeof(double) * Np));
        IBO_CUDA_TRY(pool_malloc((void**)&m->dAlpha1, sizeof(double) * Np));
        launch_tri_matvec_t(m->dW, m->dBetaY, m->dAlphaY, Np, Np, st);       // alpha = W^T (W Y)
        launch_tri_matvec_t(m->dW, m->dBeta1, m->dAlpha1, Np, Np, st);
        IBO_CUDA_TRY((i8_set_k2_attrs<7, 7>()));
        IBO_CUDA_TRY((i8_set_k2_attrs<8, 7>()));
        IBO_CUDA_TRY((i8_set_k2_attrs<7, 8>()));
        IBO_CUDA_TRY((i8_set_k2_attrs<8, 8>()));
        IBO_CUDA_TRY((i8_set_k2_attrs<6, 8, 6>()));
        // (A/B switch for the next round: not part of the configuration that was run on a device)
        if (getenv("IBO_I8_K1_CARVEOUT")) { static cudaError_t k1e = i8_set_k1_attrs(); IBO_CUDA_TRY(k1e); }
        for (auto& e : m->evI8) if (!e) IBO_CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    bool& valid = mode == 2 ? m->i8Valid6 : (mode == 1 ? m->i8Valid8 : m->i8Valid);
    if (valid) return IBO_OK;
    double*& slices = mode == 2 ? m->dWi8c : (mode == 1 ? m->dWi8b : m->dWi8);
    double*& scale = mode == 0 ? m->dRowScale : m->dRowScale8;       // the two 8-bit modes share scales and constants
    if (slices) { pool_free(slices); slices = nullptr; }
    IBO_CUDA_TRY(pool_malloc((void**)&slices, wi8_base(nb)));
    const bool haveScale = mode != 0 && (m->i8Valid8 || m->i8Valid6) && scale;
    if (!haveScale) {
        if (scale) { pool_free(scale); scale = nullptr; }
        IBO_CUDA_TRY(pool_malloc((void**)&scale, sizeof(double) * 3 * Np));
        i8_rowscale_kernel<<<(Np + 7) / 8, 256, 0, st>>>(m->dW, Np, m->N, m->sf2, mode == 0 ? 2 : 3, scale);
        g_launches++;
    }
    const dim3 grid(nb * 4, nb);
    if (mode == 2) i8_slice_w_kernel<8, 6><<<grid, 256, 0, st>>>(m->dW, scale, Np, reinterpret_cast<uint8_t*>(slices));
    else if (mode == 1) i8_slice_w_kernel<8, 7><<<grid, 256, 0, st>>>(m->dW, scale, Np, reinterpret_cast<uint8_t*>(slices));
    else i8_slice_w_kernel<7, 7><<<grid, 256, 0, st>>>(m->dW, scale, Np, reinterpret_cast<uint8_t*>(slices));
    g_launches++;
    IBO_CUDA_TRY(cudaGetLastError());
    valid = true;
    return IBO_OK;
}

// IBO_INT8 in the environment forces the path for every wide scoring call: 1 = the validated 7 x 7-bit digits, 8 = 7 x 8-bit digits
// (IBO_FLAG_INT8_D8), 6 = 6 x 8-bit digits (IBO_FLAG_INT8_S6), 9 = eighth accumulator group (IBO_FLAG_INT8_G9), 89 = 8 and 9.
// IBO_I8_PIPE=0 runs K1 and K2 back to back on one stream (A/B of the two-stream pipeline).
static int i8_env() {
    static int env = -1;
    if (env < 0) { const char* e = getenv("IBO_INT8"); env = e ? atoi(e) : 0; }
    return env;
}
static int i8_effective_flags(int flags) {
    const int env = i8_env();
    if (env == 1) flags |= IBO_FLAG_INT8;
    if (env == 8 || env == 89) flags |= IBO_FLAG_INT8_D8;
    if (env == 9 || env == 89) flags |= IBO_FLAG_INT8_G9;
    if (env == 6) flags |= IBO_FLAG_INT8_S6;
    return flags;
}
// d <= 32 (K1's register-resident candidate), no variance model
static bool i8_requested(int flags) {
    return (i8_effective_flags(flags) & (IBO_FLAG_INT8 | IBO_FLAG_INT8_G9 | IBO_FLAG_INT8_D8 | IBO_FLAG_INT8_S6)) != 0;
}
// 8-bit digits: |D_g| <= 7 pairs x N x 128 x 128 must fit INT32
static int i8_mode(const ibo_model* m, int flags) {
    const int f = i8_effective_flags(flags);
    if (m->Np > 16384) return 0;
    return (f & IBO_FLAG_INT8_S6) ? 2 : ((f & IBO_FLAG_INT8_D8) ? 1 : 0);
}
static bool i8_g9(int flags) { return (i8_effective_flags(flags) & IBO_FLAG_INT8_G9) != 0; }
static bool i8_pipe_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("IBO_I8_PIPE"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

template <int DMAX, int BITS, int S>
static void launch_kstar_i8_d(ibo_model* m, const double* dCand, dim3 g1, long M, long m0, long Mpad, uint8_t* Ki8, double* part, cudaStream_t st) {
    const int p1 = m->npb > 0 ? 1 : 0;
    if (m->kind <= IBO_KERNEL_SE_ISO)
        kstar_i8_kernel<0, DMAX, BITS, S><<<g1, 256, 0, st>>>(m->dXt, dCand, m->dInvTheta, m->dCenter, m->dAlphaY, m->dAlpha1, Ki8, part, m->N, m->d, m->nb, M, m0, Mpad, m->sf2, p1);
    else if (m->kind == IBO_KERNEL_MATERN3)
        kstar_i8_kernel<1, DMAX, BITS, S><<<g1, 256, 0, st>>>(m->dXt, dCand, m->dInvTheta, m->dCenter, m->dAlphaY, m->dAlpha1, Ki8, part, m->N, m->d, m->nb, M, m0, Mpad, m->sf2, p1);
    else
        kstar_i8_kernel<2, DMAX, BITS, S><<<g1, 256, 0, st>>>(m->dXt, dCand, m->dInvTheta, m->dCenter, m->dAlphaY, m->dAlpha1, Ki8, part, m->N, m->d, m->nb, M, m0, Mpad, m->sf2, p1);
}

template <int BITS, int S>
static void launch_kstar_i8_b(ibo_model* m, const double* dCand, long tiles, long M, long m0, long Mpad, uint8_t* Ki8, double* part, cudaStream_t st) {
    dim3 g1((unsigned)(tiles * 2), m->nb);
    const int d = m->d;
    if (d <= 2) launch_kstar_i8_d<2, BITS, S>(m, dCand, g1, M, m0, Mpad, Ki8, part, st);
    else if (d <= 4) launch_kstar_i8_d<4, BITS, S>(m, dCand, g1, M, m0, Mpad, Ki8, part, st);
    else if (d <= 6) launch_kstar_i8_d<6, BITS, S>(m, dCand, g1, M, m0, Mpad, Ki8, part, st);
    else if (d <= 8) launch_kstar_i8_d<8, BITS, S>(m, dCand, g1, M, m0, Mpad, Ki8, part, st);
    else if (d <= 12) launch_kstar_i8_d<12, BITS, S>(m, dCand, g1, M, m0, Mpad, Ki8, part, st);
    else if (d <= 16) launch_kstar_i8_d<16, BITS, S>(m, dCand, g1, M, m0, Mpad, Ki8, part, st);
    else if (d <= 24) launch_kstar_i8_d<24, BITS, S>(m, dCand, g1, M, m0, Mpad, Ki8, part, st);
    else launch_kstar_i8_d<32, BITS, S>(m, dCand, g1, M, m0, Mpad, Ki8, part, st);
}

// K1 of one chunk on the int8 path; tiles = 128-candidate tiles of the chunk (the slab holds 2 * tiles 64-candidate tiles)
static void launch_kstar_i8(ibo_model* m, const double* dCand, long tiles, long M, long m0, long Mpad, uint8_t* Ki8, double* part, cudaStream_t st,
                            int mode = 0) {
    if (mode == 2) launch_kstar_i8_b<8, 6>(m, dCand, tiles, M, m0, Mpad, Ki8, part, st);
    else if (mode == 1) launch_kstar_i8_b<8, 7>(m, dCand, tiles, M, m0, Mpad, Ki8, part, st);
    else launch_kstar_i8_b<7, 7>(m, dCand, tiles, M, m0, Mpad, Ki8, part, st);
}

// K2 of one chunk on the int8 path
static void launch_trigemm_i8(ibo_model* m, long tiles, long Mpad, const uint8_t* Ki8, double* part, cudaStream_t st, bool g9 = false, int mode = 0) {
    int G = std::max(1, m->nb / 4);
    if (tiles * 2 * G < g_num_sms) G = (int)std::min<long>(m->nb, (g_num_sms + tiles * 2 - 1) / (tiles * 2));
    const dim3 grid(G, (unsigned)(tiles * 2));
    const uint8_t* Wsl = reinterpret_cast<const uint8_t*>(mode == 2 ? m->dWi8c : (mode == 1 ? m->dWi8b : m->dWi8));
    const double* rs = (mode == 0 ? m->dRowScale : m->dRowScale8) + m->Np;
    if (mode == 2) trigemm_i8_kernel<6, 8, 6><<<grid, I8_THREADS, I8_SMEM, st>>>(Wsl, Ki8, rs, part, m->nb, Mpad);
    else if (mode == 1 && g9) trigemm_i8_kernel<8, 8, 7><<<grid, I8_THREADS, I8_SMEM, st>>>(Wsl, Ki8, rs, part, m->nb, Mpad);
    else if (mode == 1) trigemm_i8_kernel<7, 8, 7><<<grid, I8_THREADS, I8_SMEM, st>>>(Wsl, Ki8, rs, part, m->nb, Mpad);
    else if (g9) trigemm_i8_kernel<8, 7, 7><<<grid, I8_THREADS, I8_SMEM, st>>>(Wsl, Ki8, rs, part, m->nb, Mpad);
    else trigemm_i8_kernel<7, 7, 7><<<grid, I8_THREADS, I8_SMEM, st>>>(Wsl, Ki8, rs, part, m->nb, Mpad);
}

// ---- live INT8 tensor peak (bench.py): back-to-back 128 x 256 x 32 MMAs from resident operands, one CTA per SM ----
__global__ void __launch_bounds__(128) i8_peak_kernel(int iters, int* __restrict__ sink) {
    extern __shared__ __align__(1024) uint8_t smem[];      // A: 128 x 128 B, B: 256 x 128 B (four k-steps of 32)
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 + 256) * 128; i += 128) smem[i] = (uint8_t)((i * 7 + 3) & 3);
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tbase)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t td = tbase;
    if (tid == 0) {
        // 128-byte-deep tiles: the k chunks of a row group are 128 B apart, row groups 8 * 128 B apart
        auto desc = [](uint32_t addr) {
            return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46);
        };
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 128 * 128);
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int kk = 0; kk < 4; kk++) umma_i8(td + (uint32_t)(it & 1) * 256u, desc(a0 + kk * 256), desc(b0 + kk * 256), umma_idesc_i8(128, 256), 1u);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t v[8];
    tmem_ld8_nowait(td + ((uint32_t)(warp * 32) << 16), v);
    tmem_ld_wait();
    if (v[0] == 0x7fffffffu) sink[blockIdx.x * 128 + tid] = (int)v[1];
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(td), "r"(512u) : "memory");
}

}  // namespace ibo
