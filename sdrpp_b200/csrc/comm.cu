// Multi-GPU feed of the IQ stream: one process per GPU, the VFO set sharded across them, every block
// broadcast from the ingest rank over NVLink. The reference's fan-out is dsp::routing::Splitter::run
// (dsp/routing/splitter.h:46-60: one memcpy + blocking swap per consumer); here it is ONE ncclBroadcast of the raw
// (still packed) samples per block, issued by the library itself on a stream of its own, so that the host does no
// per-block collective bookkeeping outside sdrpp_cuda_frontend_submit*.
//
// NCCL is loaded at run time (dlopen of libnccl.so.2, the copy already in the process if there is one), so the library
// has no link-time dependency on it and single-GPU users never touch it.
#include "common.cuh"
#include "comm.h"
#include "../../include/sdrpp_cuda.h"

#include <dlfcn.h>
#include <cstring>
#include <map>
#include <mutex>
#include <string>

namespace sdrpp {

cudaError_t ensure_dynamic_smem(const void* func, size_t bytes, int carveout) {
    static std::mutex mtx;
    static std::map<std::pair<const void*, int>, size_t> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lck(mtx);
    size_t& cur = done[std::make_pair(func, dev)];
    if (bytes <= cur) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    // The SM's shared-memory carve-out is chosen per kernel from what ITS CTAs need (stage 1: 189 KB -> the 196 KB setting),
    // and it cannot change while CTAs are resident: without this preference a tail CTA of the previous block never fits
    // beside the persistent stage-1 CTA although 228 KB would hold both (seen as a tail kernel that takes 75 us instead of 37
    // whenever it overlaps stage 1, tools/timeline_probe.py). Every kernel with opt-in shared memory asks for the largest
    // unless its launcher says otherwise (the spectrum kernels, fft.cu).
    if (e == cudaSuccess) e = cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, carveout);
    if (e == cudaSuccess) cur = bytes;
    return e;
}

namespace {

// The slice of nccl.h this file uses (stable since NCCL 2.0): opaque communicator, 128-byte unique id.
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;   // 0 = ncclSuccess
constexpr int kNcclChar = 0; // ncclInt8 / ncclChar

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string error;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = { getenv("SDRPP_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
        for (const char* n : names) {
            if (!n || !*n) continue;
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.error = std::string("dlopen(libnccl.so.2): ") + dlerror(); return; }
        auto sym = [&](const char* s) { void* p = dlsym(api.handle, s); if (!p && api.error.empty()) api.error = std::string("missing NCCL symbol ") + s; return p; };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
    });
    return api;
}

int nccl_fail(const char* what, ncclResult_t r) {
    NcclApi& a = nccl();
    set_last_error(std::string(what) + ": " + (a.GetErrorString ? a.GetErrorString(r) : "NCCL error"));
    return SDRPP_ERR_CUDA;
}

} // namespace

cudaError_t comm_broadcast(sdrpp_cuda_comm* c, const void* send, void* recv, size_t bytes, int root, cudaStream_t st) {
    NcclApi& a = nccl();
    if (!c || !c->nccl || !a.Broadcast) return cudaErrorNotSupported;
    const ncclResult_t r = a.Broadcast(send, recv, bytes, kNcclChar, root, (ncclComm_t)c->nccl, st);
    if (r != 0) { set_last_error(std::string("ncclBroadcast: ") + (a.GetErrorString ? a.GetErrorString(r) : "NCCL error")); return cudaErrorUnknown; }
    c->broadcasts++;
    c->bytes += (long long)bytes;
    return cudaSuccess;
}

} // namespace sdrpp

using namespace sdrpp;

extern "C" {

int sdrpp_cuda_comm_unique_id(void* id128) {
    if (!id128) { set_last_error("null id"); return SDRPP_ERR_ARG; }
    NcclApi& a = nccl();
    if (!a.error.empty() || !a.GetUniqueId) { set_last_error("NCCL unavailable: " + a.error); return SDRPP_ERR_STATE; }
    ncclUniqueId id;
    const ncclResult_t r = a.GetUniqueId(&id);
    if (r != 0) return nccl_fail("ncclGetUniqueId", r);
    memcpy(id128, &id, sizeof(id));
    return SDRPP_OK;
}

sdrpp_cuda_comm* sdrpp_cuda_comm_create(const void* id128, int rank, int nranks, int device) {
    if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) { set_last_error("bad communicator arguments"); return nullptr; }
    NcclApi& a = nccl();
    if (!a.error.empty() || !a.CommInitRank) { set_last_error("NCCL unavailable: " + a.error); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { set_last_error("cudaSetDevice failed"); cudaGetLastError(); return nullptr; }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    const ncclResult_t r = a.CommInitRank(&comm, nranks, id, rank);
    if (r != 0) { nccl_fail("ncclCommInitRank", r); return nullptr; }
    auto* c = new sdrpp_cuda_comm();
    c->nccl = comm; c->rank = rank; c->nranks = nranks; c->device = device;
    if (a.GetVersion) a.GetVersion(&c->version);
    return c;
}

int sdrpp_cuda_comm_destroy(sdrpp_cuda_comm* c) {
    if (!c) return SDRPP_OK;
    NcclApi& a = nccl();
    cudaSetDevice(c->device);
    if (c->nccl && a.CommDestroy) a.CommDestroy((ncclComm_t)c->nccl);
    delete c;
    return SDRPP_OK;
}

int sdrpp_cuda_comm_info(sdrpp_cuda_comm* c, int* rank, int* nranks, int* nccl_version, long long* broadcasts, long long* bytes) {
    if (!c) { set_last_error("null communicator"); return SDRPP_ERR_ARG; }
    if (rank) *rank = c->rank;
    if (nranks) *nranks = c->nranks;
    if (nccl_version) *nccl_version = c->version;
    if (broadcasts) *broadcasts = c->broadcasts;
    if (bytes) *bytes = c->bytes;
    return SDRPP_OK;
}

} // extern "C"
