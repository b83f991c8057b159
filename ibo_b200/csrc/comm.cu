// Multi-GPU plumbing: one process per GPU, one NCCL communicator over the NVLink 5 / NVSwitch domain.
// The acquisition path shards candidates across ranks with NO data-path collective; the only exchange
// is the (score, global index) argmax (16 bytes per rank, all-gathered then reduced locally with the
// lowest-index-wins rule, because NCCL has no argmax op) and an optional broadcast of small factors.
// NCCL is dlopen'ed on first use so that the library loads (and the CPU-side symbol tests run) on
// machines without NCCL or without a GPU.
#include "common.cuh"
#include "../../include/ibo_b200.h"
#include <dlfcn.h>
#include <cstring>
#include <vector>

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclChar = 0, ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct Nccl {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
} N;

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1, device = 0;
    cudaStream_t stream = nullptr;
    unsigned char* dPair = nullptr;    // 16 bytes
    unsigned char* dAll = nullptr;     // 16 * nranks
    // all-gather of per-rank result slices (sharded DIRECT batches): pinned host staging + device buffers, grown on demand
    double* hGather = nullptr;         // [count * (nranks + 1)]: own slice, then everybody's
    double* dGather = nullptr;
    size_t gatherCap = 0;              // doubles per rank
} C;

int load_nccl() {
    if (N.h) return IBO_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    for (int i = 0; names[i] && !N.h; i++) N.h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!N.h) { ibo::set_error(std::string("dlopen(libnccl.so.2) failed: ") + dlerror()); return IBO_E_COMM; }
#define SYM(field, name) *(void**)(&N.field) = dlsym(N.h, name); if (!N.field) { ibo::set_error(std::string("missing NCCL symbol ") + name); return IBO_E_COMM; }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllGather, "ncclAllGather")
    SYM(Broadcast, "ncclBroadcast")
    SYM(AllReduce, "ncclAllReduce")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return IBO_OK;
}

#define NCCL_TRY(expr) do { ncclResult_t r__ = (expr); if (r__ != 0) { ibo::set_error(std::string(#expr) + ": " + N.GetErrorString(r__)); return IBO_E_COMM; } } while (0)

}  // namespace

extern "C" int ibo_comm_unique_id(unsigned char* id128) {
    if (!id128) return IBO_E_BADARG;
    int rc = load_nccl();
    if (rc) return rc;
    ncclUniqueId id;
    NCCL_TRY(N.GetUniqueId(&id));
    std::memcpy(id128, id.internal, 128);
    return IBO_OK;
}

extern "C" int ibo_comm_init(int device, int rank, int nranks, const unsigned char* id128) {
    if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) { ibo::set_error("bad argument"); return IBO_E_BADARG; }
    int rc = load_nccl();
    if (rc) return rc;
    if (C.comm) ibo_comm_destroy();
    IBO_CUDA_TRY(cudaSetDevice(device));
    C.rank = rank; C.nranks = nranks; C.device = device;
    IBO_CUDA_TRY(cudaStreamCreateWithFlags(&C.stream, cudaStreamNonBlocking));
    IBO_CUDA_TRY(cudaMalloc(&C.dPair, 16));
    IBO_CUDA_TRY(cudaMalloc(&C.dAll, 16 * (size_t)nranks));
    ncclUniqueId id;
    std::memcpy(id.internal, id128, 128);
    NCCL_TRY(N.CommInitRank(&C.comm, nranks, id, rank));
    return IBO_OK;
}

extern "C" int ibo_comm_destroy(void) {
    if (C.comm && N.CommDestroy) N.CommDestroy(C.comm);
    C.comm = nullptr;
    if (C.dPair) cudaFree(C.dPair);
    if (C.dAll) cudaFree(C.dAll);
    if (C.stream) cudaStreamDestroy(C.stream);
    if (C.hGather) cudaFreeHost(C.hGather);
    if (C.dGather) cudaFree(C.dGather);
    C.hGather = C.dGather = nullptr; C.gatherCap = 0;
    C.dPair = C.dAll = nullptr; C.stream = nullptr;
    C.rank = 0; C.nranks = 1;
    return IBO_OK;
}

extern "C" int ibo_comm_rank(void) { return C.comm ? C.rank : 0; }
extern "C" int ibo_comm_size(void) { return C.comm ? C.nranks : 1; }

// all[r * count + i] = rank r's mine[i]; every rank passes the same count (host buffers; one H2D, one NCCL all-gather over
// NVLink, one D2H on the communicator's stream)
extern "C" int ibo_comm_allgather(const double* mine, long count, double* all) {
    if (!mine || !all || count < 0) return IBO_E_BADARG;
    if (count == 0) return IBO_OK;
    if (!C.comm) {
        if (C.nranks == 1) { std::memcpy(all, mine, sizeof(double) * (size_t)count); return IBO_OK; }
        ibo::set_error("communicator not initialised"); return IBO_E_COMM;
    }
    IBO_CUDA_TRY(cudaSetDevice(C.device));
    const size_t per = (size_t)count, tot = per * (size_t)(C.nranks + 1);
    if (C.gatherCap < per) {
        if (C.hGather) cudaFreeHost(C.hGather);
        if (C.dGather) cudaFree(C.dGather);
        C.hGather = C.dGather = nullptr; C.gatherCap = 0;
        size_t cap = per < 4096 ? 4096 : per + per / 2;
        IBO_CUDA_TRY(cudaHostAlloc((void**)&C.hGather, sizeof(double) * cap * (size_t)(C.nranks + 1), cudaHostAllocPortable));
        IBO_CUDA_TRY(cudaMalloc((void**)&C.dGather, sizeof(double) * cap * (size_t)(C.nranks + 1)));
        C.gatherCap = cap;
    }
    (void)tot;
    std::memcpy(C.hGather, mine, sizeof(double) * per);
    IBO_CUDA_TRY(cudaMemcpyAsync(C.dGather, C.hGather, sizeof(double) * per, cudaMemcpyHostToDevice, C.stream));
    NCCL_TRY(N.AllGather(C.dGather, C.dGather + per, per, ncclFloat64, C.comm, C.stream));
    IBO_CUDA_TRY(cudaMemcpyAsync(C.hGather + per, C.dGather + per, sizeof(double) * per * (size_t)C.nranks, cudaMemcpyDeviceToHost, C.stream));
    IBO_CUDA_TRY(cudaStreamSynchronize(C.stream));
    std::memcpy(all, C.hGather + per, sizeof(double) * per * (size_t)C.nranks);
    return IBO_OK;
}

// merge rule shared with the CPU-side tests: max score, lowest global index on ties, NaN never wins
static inline bool better(double s, long long i, double bs, long long bi) {
    if (s != s) return false;
    if (bs != bs) return true;
    return s > bs || (s == bs && i < bi);
}

extern "C" int ibo_comm_argmax(double* score, long* index) {
    if (!score || !index) return IBO_E_BADARG;
    if (!C.comm) { if (C.nranks == 1) return IBO_OK; ibo::set_error("communicator not initialised"); return IBO_E_COMM; }
    IBO_CUDA_TRY(cudaSetDevice(C.device));
    unsigned char pair[16];
    long long li = *index;
    std::memcpy(pair, score, 8); std::memcpy(pair + 8, &li, 8);
    IBO_CUDA_TRY(cudaMemcpyAsync(C.dPair, pair, 16, cudaMemcpyHostToDevice, C.stream));
    NCCL_TRY(N.AllGather(C.dPair, C.dAll, 16, ncclChar, C.comm, C.stream));
    std::vector<unsigned char> all(16 * (size_t)C.nranks);
    IBO_CUDA_TRY(cudaMemcpyAsync(all.data(), C.dAll, all.size(), cudaMemcpyDeviceToHost, C.stream));
    IBO_CUDA_TRY(cudaStreamSynchronize(C.stream));
    double bs = 0; long long bi = 0; bool have = false;
    for (int r = 0; r < C.nranks; r++) {
        double s; long long i;
        std::memcpy(&s, &all[16 * r], 8); std::memcpy(&i, &all[16 * r + 8], 8);
        if (!have || better(s, i, bs, bi)) { bs = s; bi = i; have = true; }
    }
    *score = bs; *index = (long)bi;
    return IBO_OK;
}

extern "C" int ibo_comm_bcast(double* buf, long count, int root) {
    if (!buf || count < 0) return IBO_E_BADARG;
    if (!C.comm) { if (C.nranks == 1) return IBO_OK; ibo::set_error("communicator not initialised"); return IBO_E_COMM; }
    if (count == 0) return IBO_OK;
    IBO_CUDA_TRY(cudaSetDevice(C.device));
    double* d = nullptr;
    IBO_CUDA_TRY(cudaMalloc(&d, sizeof(double) * (size_t)count));
    if (C.rank == root) IBO_CUDA_TRY(cudaMemcpyAsync(d, buf, sizeof(double) * (size_t)count, cudaMemcpyHostToDevice, C.stream));
    ncclResult_t r = N.Broadcast(d, d, (size_t)count, ncclFloat64, root, C.comm, C.stream);
    if (r != 0) { cudaFree(d); ibo::set_error(std::string("ncclBroadcast: ") + N.GetErrorString(r)); return IBO_E_COMM; }
    cudaError_t e = cudaMemcpyAsync(buf, d, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, C.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(C.stream);
    cudaFree(d);
    if (e != cudaSuccess) { ibo::set_error(std::string("bcast copy: ") + cudaGetErrorString(e)); return IBO_E_CUDA; }
    return IBO_OK;
}

extern "C" int ibo_comm_barrier(void) {
    if (!C.comm) return IBO_OK;
    IBO_CUDA_TRY(cudaSetDevice(C.device));
    IBO_CUDA_TRY(cudaMemsetAsync(C.dPair, 0, 16, C.stream));
    NCCL_TRY(N.AllReduce(C.dPair, C.dPair, 2, ncclFloat64, ncclSum, C.comm, C.stream));
    IBO_CUDA_TRY(cudaStreamSynchronize(C.stream));
    return IBO_OK;
}
