// FP64 peak microbenchmark for B200 (sm_100a): DFMA vs DMMA.8x8x4 issue-bound throughput.
// Prints one JSON line. Used to establish the measured FP64 roofline denominator that
// MEASURED_PEAKS.json lacks (it only carries HBM GB/s and bf16 TF/s).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) c[i] = threadIdx.x * 1e-6 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a0, double b0) {
    double c0[ILP], c1[ILP];
    double a = a0 + threadIdx.x * 1e-9, b = b0 + threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < ILP; i++) { c0[i] = i; c1[i] = -i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; i++) launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main(int argc, char** argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 16 * 256));
    const int iters = 20000;
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    // DFMA: blocks/SM x ILP sweep
    {
        double best = 0; int bb = 0;
        for (int bps = 1; bps <= 8; bps *= 2) {
            double ms = time_ms([&] { dfma_kernel<8><<<sms * bps, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double tf = 2.0 * 8 * iters * 256.0 * sms * bps / (ms * 1e-3) / 1e12;
            if (tf > best) { best = tf; bb = bps; }
        }
        printf(", \"dfma_tflops\": %.3f, \"dfma_blocks_per_sm\": %d", best, bb);
    }
    // DMMA.8x8x4: 8*8*4*2 = 512 flop per warp-instruction
    {
        double best = 0; int bb = 0, bi = 0;
        for (int bps = 1; bps <= 8; bps *= 2) {
            double ms4 = time_ms([&] { dmma_kernel<4><<<sms * bps, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double tf4 = 512.0 * 4 * iters * 8.0 * sms * bps / (ms4 * 1e-3) / 1e12;
            if (tf4 > best) { best = tf4; bb = bps; bi = 4; }
            double ms8 = time_ms([&] { dmma_kernel<8><<<sms * bps, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double tf8 = 512.0 * 8 * iters * 8.0 * sms * bps / (ms8 * 1e-3) / 1e12;
            if (tf8 > best) { best = tf8; bb = bps; bi = 8; }
        }
        printf(", \"dmma_tflops\": %.3f, \"dmma_blocks_per_sm\": %d, \"dmma_ilp\": %d", best, bb, bi);
        // single warp per SMSP latency-bound point (ILP=1 chain): gives DMMA latency
        double ms1 = time_ms([&] { dmma_kernel<1><<<sms, 128>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf(", \"dmma_chain_ns_per_instr\": %.3f", ms1 * 1e6 / iters);
    }
    // sustained (2 s loop) DMMA
    {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        int n = 0; float ms = 0;
        do {
            for (int i = 0; i < 20; i++) dmma_kernel<8><<<sms * 4, 256>>>(out, iters, 1.0000001, 1e-9);
            n += 20;
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
        } while (ms < 2000.f);
        double tf = 512.0 * 8 * iters * 8.0 * sms * 4 * n / (ms * 1e-3) / 1e12;
        printf(", \"dmma_tflops_sustained\": %.3f", tf);
    }
    printf("}\n");
    return 0;
}
