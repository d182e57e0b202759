// Communicator object behind sdrpp_cuda_comm_* (comm.cu) and the broadcast the engine issues per block.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

struct sdrpp_cuda_comm {
    void* nccl = nullptr;     // ncclComm_t
    int rank = 0, nranks = 1, device = 0, version = 0;
    long long broadcasts = 0, bytes = 0;
};

namespace sdrpp {
// ncclBroadcast of `bytes` bytes from `send` on rank `root` into `recv` everywhere (send == recv on the root: in place)
cudaError_t comm_broadcast(sdrpp_cuda_comm* c, const void* send, void* recv, size_t bytes, int root, cudaStream_t st);
}
