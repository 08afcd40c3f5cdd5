// kq_comm.cu — multi-GPU plumbing: one kq_ctx (= one GPU, one process or thread) per rank, NCCL over
// NVLink 5 / NVSwitch between them. This is the "exchange" of the reference's main(): it concatenates
// the partial results of 12 coroutines (Main.kt:1309-1316) and re-aggregates them (Main.kt:1320-1325);
// here partial aggregation tables are merged with collectives (kq_hashagg.cu: kq_hashagg_merge_allreduce,
// kq_hashagg_repartition_alltoall).
#include <dlfcn.h>

#include <mutex>

#include "kq_comm.h"

namespace {
std::mutex g_mu;
KqNccl g_nccl;
bool g_tried = false;
}  // namespace

KqNccl* kq_nccl(kq_ctx* ctx) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!g_tried) {
        g_tried = true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (g_nccl.h) break;
        }
        if (!g_nccl.h) g_nccl.error = "libnccl.so.2 not found (dlopen)";
#define KQ_SYM(field, name)                                                                     \
    if (g_nccl.h) {                                                                             \
        g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.h, name));         \
        if (!g_nccl.field) { g_nccl.error = std::string("libnccl lacks ") + name; g_nccl.h = nullptr; } \
    }
        KQ_SYM(GetUniqueId, "ncclGetUniqueId")
        KQ_SYM(CommInitRank, "ncclCommInitRank")
        KQ_SYM(CommDestroy, "ncclCommDestroy")
        KQ_SYM(AllReduce, "ncclAllReduce")
        KQ_SYM(AllGather, "ncclAllGather")
        KQ_SYM(Send, "ncclSend")
        KQ_SYM(Recv, "ncclRecv")
        KQ_SYM(GroupStart, "ncclGroupStart")
        KQ_SYM(GroupEnd, "ncclGroupEnd")
        KQ_SYM(GetErrorString, "ncclGetErrorString")
#undef KQ_SYM
    }
    if (!g_nccl.h) { kq_fail(ctx, KQ_ERR_NCCL, "%s", g_nccl.error.c_str()); return nullptr; }
    return &g_nccl;
}

int kq_nccl_fail(kq_ctx* ctx, ncclResult_t r, const char* what) {
    return kq_fail(ctx, KQ_ERR_NCCL, "%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error");
}

extern "C" {

static_assert(sizeof(ncclUniqueId) == KQ_COMM_ID_BYTES, "KQ_COMM_ID_BYTES must match ncclUniqueId");

int kq_comm_unique_id(kq_ctx* ctx, uint8_t id[KQ_COMM_ID_BYTES]) {
    if (!ctx || !id) return KQ_ERR_ILLEGAL_ARGUMENT;
    KqNccl* N = kq_nccl(ctx);
    if (!N) return KQ_ERR_NCCL;
    ncclUniqueId u;
    KQ_NCCL(ctx, N->GetUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return KQ_OK;
}

int kq_comm_init(kq_ctx* ctx, const uint8_t id[KQ_COMM_ID_BYTES], int rank, int nranks) {
    if (!ctx || !id || nranks < 1 || rank < 0 || rank >= nranks) return KQ_ERR_ILLEGAL_ARGUMENT;
    if (ctx->comm) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "communicator already initialised");
    KqNccl* N = kq_nccl(ctx);
    if (!N) return KQ_ERR_NCCL;
    cudaSetDevice(ctx->device);
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclComm_t comm = nullptr;
    KQ_NCCL(ctx, N->CommInitRank(&comm, nranks, u, rank));
    ctx->comm = comm; ctx->rank = rank; ctx->nranks = nranks;
    return KQ_OK;
}

int kq_comm_destroy(kq_ctx* ctx) {
    if (!ctx || !ctx->comm) return KQ_OK;
    KqNccl* N = kq_nccl(ctx);
    if (!N) return KQ_ERR_NCCL;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    N->CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr; ctx->rank = 0; ctx->nranks = 1;
    return KQ_OK;
}

// Max over ranks of a float (the device-side timing reduction of bench.py).
int kq_comm_allreduce_max_f32(kq_ctx* ctx, float* inout) {
    if (!ctx || !inout) return KQ_ERR_ILLEGAL_ARGUMENT;
    if (!ctx->comm) return KQ_OK;              // single rank
    KqNccl* N = kq_nccl(ctx);
    if (!N) return KQ_ERR_NCCL;
    cudaSetDevice(ctx->device);
    float* d = nullptr;
    KQ_RET(kq_dev_alloc(ctx, 16, (void**)&d));
    cudaMemcpyAsync(d, inout, 4, cudaMemcpyHostToDevice, ctx->stream);
    ncclResult_t r = N->AllReduce(d, d, 1, ncclFloat32, ncclMax, (ncclComm_t)ctx->comm, ctx->stream);
    cudaMemcpyAsync(inout, d, 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    kq_dev_free(ctx, d);
    if (r != ncclSuccess) return kq_nccl_fail(ctx, r, "ncclAllReduce");
    if (e != cudaSuccess) return kq_cuda_fail(ctx, e, "cudaStreamSynchronize");
    return KQ_OK;
}

int kq_comm_barrier(kq_ctx* ctx) {
    float one = 1.0f;
    return kq_comm_allreduce_max_f32(ctx, &one);
}

}  // extern "C"
