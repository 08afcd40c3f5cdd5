// kq_comm.h — NCCL plumbing shared by kq_comm.cu (communicator life cycle) and kq_hashagg.cu (the merge
// collectives). libnccl is opened with dlopen on first use: libkqgpu.so itself has no link-time
// dependency on it, and inside a process that already loaded NCCL (torch.distributed) the same copy is used.
#pragma once

#include <nccl.h>

#include "kq_internal.h"

struct KqNccl {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

// The process-wide function table (nullptr + ctx error when libnccl cannot be loaded).
KQ_HIDDEN KqNccl* kq_nccl(kq_ctx* ctx);
KQ_HIDDEN int kq_nccl_fail(kq_ctx* ctx, ncclResult_t r, const char* what);

#define KQ_NCCL(ctx, call)                                               \
    do {                                                                 \
        ncclResult_t _r = (call);                                        \
        if (_r != ncclSuccess) return kq_nccl_fail((ctx), _r, #call);    \
    } while (0)
