// kq_comm.cu — multi-GPU plumbing (NCCL, loaded lazily with dlopen). Filled in after the single-GPU path.
#include "kq_internal.h"

extern "C" {
int kq_comm_unique_id(kq_ctx* ctx, uint8_t id[KQ_COMM_ID_BYTES]) { (void)id; return kq_fail(ctx, KQ_ERR_NCCL, "comm not built yet"); }
int kq_comm_init(kq_ctx* ctx, const uint8_t id[KQ_COMM_ID_BYTES], int rank, int nranks) { (void)id; (void)rank; (void)nranks; return kq_fail(ctx, KQ_ERR_NCCL, "comm not built yet"); }
int kq_comm_destroy(kq_ctx* ctx) { (void)ctx; return KQ_OK; }
int kq_comm_barrier(kq_ctx* ctx) { return kq_fail(ctx, KQ_ERR_NCCL, "comm not built yet"); }
int kq_comm_allreduce_max_f32(kq_ctx* ctx, float* inout) { (void)inout; return kq_fail(ctx, KQ_ERR_NCCL, "comm not built yet"); }
int kq_hashagg_merge_allreduce(kq_ctx* ctx, kq_hashagg* agg) { (void)agg; return kq_fail(ctx, KQ_ERR_NCCL, "comm not built yet"); }
int kq_hashagg_repartition_alltoall(kq_ctx* ctx, kq_hashagg* agg) { (void)agg; return kq_fail(ctx, KQ_ERR_NCCL, "comm not built yet"); }
}
