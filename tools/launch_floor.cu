// Launch / completion latency floor of the small-batch path (run under gpurun):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/launch_floor tools/launch_floor.cu && tools/launch_floor
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstring>
__global__ void k_empty() {}
__global__ void k_touch(const double* in, double* out) { out[threadIdx.x] = in[threadIdx.x] + 1.0; }
__global__ void k_flag(const double* in, double* out, volatile unsigned* flag, unsigned v) {
    out[threadIdx.x] = in[threadIdx.x] + 1.0;
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence_system(); *flag = v; }
}
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    double *hp, *dp; cudaHostAlloc((void**)&hp, 1 << 16, cudaHostAllocMapped); cudaMalloc(&dp, 1 << 16);
    unsigned* hflag; cudaHostAlloc((void**)&hflag, 64, cudaHostAllocMapped); *hflag = 0;
    memset(hp, 0, 1 << 16);
    const int R = 2000;
    auto bench = [&](const char* name, auto fn) {
        for (int i = 0; i < 50; i++) fn(i);
        double t0 = now();
        for (int i = 0; i < R; i++) fn(100 + i);
        printf("%-58s %7.2f us\n", name, 1e6 * (now() - t0) / R);
    };
    bench("1 empty kernel + streamSync", [&](int) { k_empty<<<1, 32, 0, st>>>(); cudaStreamSynchronize(st); });
    bench("3 empty kernels + streamSync", [&](int) { for (int k = 0; k < 3; k++) k_empty<<<1, 32, 0, st>>>(); cudaStreamSynchronize(st); });
    bench("3 kernels device mem + streamSync", [&](int) { for (int k = 0; k < 3; k++) k_touch<<<1, 32, 0, st>>>(dp, dp + 64); cudaStreamSynchronize(st); });
    bench("3 kernels, first reads / last writes mapped host + sync", [&](int) {
        k_touch<<<1, 32, 0, st>>>(hp, dp); k_touch<<<1, 32, 0, st>>>(dp, dp + 64); k_touch<<<1, 32, 0, st>>>(dp + 64, hp + 64); cudaStreamSynchronize(st); });
    bench("H2D 1 KB async + 3 kernels + D2H 1 KB async + sync", [&](int) {
        cudaMemcpyAsync(dp, hp, 1024, cudaMemcpyHostToDevice, st);
        for (int k = 0; k < 3; k++) k_touch<<<1, 32, 0, st>>>(dp, dp + 64);
        cudaMemcpyAsync(hp + 512, dp + 64, 1024, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st); });
    bench("1 kernel mapped in/out + host polls a mapped flag", [&](int i) {
        k_flag<<<1, 32, 0, st>>>(hp, hp + 64, hflag, (unsigned)i + 1);
        while (*(volatile unsigned*)hflag != (unsigned)i + 1) {} });
    bench("3 kernels, last sets mapped flag, host polls", [&](int i) {
        k_touch<<<1, 32, 0, st>>>(hp, dp); k_touch<<<1, 32, 0, st>>>(dp, dp + 64);
        k_flag<<<1, 32, 0, st>>>(dp + 64, hp + 64, hflag, (unsigned)i + 1);
        while (*(volatile unsigned*)hflag != (unsigned)i + 1) {} });
    // CUDA graph of 3 kernels
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    k_touch<<<1, 32, 0, st>>>(hp, dp); k_touch<<<1, 32, 0, st>>>(dp, dp + 64); k_touch<<<1, 32, 0, st>>>(dp + 64, hp + 64);
    cudaStreamEndCapture(st, &g); cudaGraphInstantiate(&ge, g, 0);
    bench("graph(3 kernels mapped in/out) + streamSync", [&](int) { cudaGraphLaunch(ge, st); cudaStreamSynchronize(st); });
    cudaEvent_t ev; cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    bench("3 kernels + eventRecord + eventSynchronize", [&](int) {
        for (int k = 0; k < 3; k++) k_touch<<<1, 32, 0, st>>>(dp, dp + 64); cudaEventRecord(ev, st); cudaEventSynchronize(ev); });
    bench("3 kernels + eventRecord + eventQuery spin", [&](int) {
        for (int k = 0; k < 3; k++) k_touch<<<1, 32, 0, st>>>(dp, dp + 64); cudaEventRecord(ev, st); while (cudaEventQuery(ev) == cudaErrorNotReady) {} });
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
