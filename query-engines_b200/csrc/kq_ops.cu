// kq_ops.cu — ProjectionExec, FilterExec and the fused filter+project kernel.
//
//   k_project         ProjectionExec.execute for one batch (Main.kt:589-594): every expression of the
//                     projection evaluated in one pass, one 128-bit store per pair of output rows.
//   k_filter_project  FilterExec (absent from the reference) fused with the
//                     projection above it: predicate -> warp ballot/popc ranks -> ordered cross-block
//                     prefix (decoupled look-back) -> compacted stores, all in a single pass over the
//                     input (algorithmic bytes only: each input column read once, each output row
//                     written once).
#include <algorithm>
#include <cstring>

#define KQ_R 4      // rows per thread in the streaming kernels
#include "kq_compile.h"
#include "kq_pipe.cuh"
#include "kq_scan.cuh"

using namespace kq;

namespace {

struct DOut {
    void* data;
    uint32_t* validity;
    int32_t type;
    int32_t _pad;
};

struct OpArgs {
    Program prog;
    int64_t n, ntiles;
    int32_t sel_end;          // instructions [0, sel_end) evaluate the predicate and end in O_SET_SEL
    int32_t nout;
    DOut outs[MAX_OUT];
    unsigned long long* tile_desc;
    unsigned int* ticket;
    unsigned long long* out_count;
    int32_t* selvec;
    uint32_t* err;
    StagePlan sp;
};

// One CTA per SM: 15 consumer warps evaluate 4 rows per thread (thread-level parallelism keeps the
// issue slots busy while each warp walks its dependent dispatch chain), one service warp streams
// tiles in with TMA bulk copies (and, in the filter kernel, resolves the cross-block prefix).
// 512 threads -> 128 registers each; bytes in flight come from the stage ring, not from occupancy.
constexpr int WARPS = 15;
constexpr int BLOCK = WARPS * 32;
constexpr int TILE = WARPS * WARP_ROWS;     // 1920 rows
constexpr int SERVICE_WARP = WARPS;
constexpr int THREADS = BLOCK + 32;

struct SinkBase {
    __device__ __forceinline__ void set_sel(const uint64_t (&)[R], uint32_t, RowCtx&) {}
    __device__ __forceinline__ void emit(int, const uint64_t (&)[R], uint32_t, RowCtx&) {}
    __device__ __forceinline__ void set_key(int, bool, const uint64_t (&)[R], uint32_t, RowCtx&) {}
    __device__ __forceinline__ void set_in(int, bool, const uint64_t (&)[R], uint32_t, RowCtx&) {}
};

// write the 64 row bits of chunk j (rows warp_base + 64j ...) of a bit-packed buffer
__device__ __forceinline__ void store_chunk_bits(uint32_t* bits, const RowCtx& rc, int j, uint32_t m) {
    uint32_t b0 = __ballot_sync(0xffffffffu, (m >> (2 * j)) & 1u);
    uint32_t b1 = __ballot_sync(0xffffffffu, (m >> (2 * j + 1)) & 1u);
    int64_t chunk_base = rc.warp_base + j * 64;
    if (rc.lane == 0 && chunk_base < rc.n) {
        uint2 w = interleave_ballots(b0, b1);
        *reinterpret_cast<uint2*>(bits + (chunk_base >> 5)) = w;
    }
}

struct ProjectSink : SinkBase {
    const DOut* outs;
    __device__ __forceinline__ void emit(int k, const uint64_t (&v)[R], uint32_t ok, RowCtx& rc) {
        const DOut o = outs[k];
        if (o.type == KQ_BOOL) {
            const uint32_t m = (uint32_t)v[0] & rc.inr;      // Bool values are truth masks
#pragma unroll
            for (int j = 0; j < NCHUNK; j++) store_chunk_bits(reinterpret_cast<uint32_t*>(o.data), rc, j, m);
        } else if (o.type == KQ_DATE32 || o.type == KQ_I32) {
#pragma unroll
            for (int j = 0; j < NCHUNK; j++) {
                int64_t r0 = rc.row0(j);
                if (rc.full || r0 < rc.n) stg_v2(reinterpret_cast<uint2*>(o.data) + (r0 >> 1), make_uint2((uint32_t)v[2 * j], (uint32_t)v[2 * j + 1]));
            }
        } else {
#pragma unroll
            for (int j = 0; j < NCHUNK; j++) {
                int64_t r0 = rc.row0(j);
                if (rc.full || r0 < rc.n)
                    stg_v4(reinterpret_cast<uint4*>(o.data) + (r0 >> 1),
                           make_uint4((uint32_t)v[2 * j], (uint32_t)(v[2 * j] >> 32), (uint32_t)v[2 * j + 1], (uint32_t)(v[2 * j + 1] >> 32)));
            }
        }
        if (o.validity) {
#pragma unroll
            for (int j = 0; j < NCHUNK; j++) store_chunk_bits(o.validity, rc, j, ok & rc.inr);
        }
    }
};

struct CompactSink : SinkBase {
    const DOut* outs;
    uint32_t sel;
    int rank[R];
    long long base;
    __device__ __forceinline__ void set_sel(const uint64_t (&v)[R], uint32_t ok, RowCtx& rc) {
        sel = (uint32_t)v[0] & ok & rc.inr;        // truth mask; TRUE only: a null predicate drops the row (rule E3)
    }
    __device__ __forceinline__ void emit(int k, const uint64_t (&v)[R], uint32_t ok, RowCtx&) {
        const DOut o = outs[k];
#pragma unroll
        for (int r = 0; r < R; r++) {
            if ((sel >> r) & 1u) {
                long long pos = base + rank[r];
                if (o.type == KQ_BOOL) {
                    if ((v[0] >> r) & 1u) atomicOr(reinterpret_cast<uint32_t*>(o.data) + (pos >> 5), 1u << (pos & 31));
                } else if (o.type == KQ_DATE32 || o.type == KQ_I32) {
                    reinterpret_cast<uint32_t*>(o.data)[pos] = (uint32_t)v[r];
                } else {
                    reinterpret_cast<uint64_t*>(o.data)[pos] = v[r];
                }
                if (o.validity && ((ok >> r) & 1u)) atomicOr(o.validity + (pos >> 5), 1u << (pos & 31));
            }
        }
    }
};

// ProjectionExec for one batch (Main.kt:589-594).
__global__ void __launch_bounds__(THREADS, 1) k_project(const __grid_constant__ OpArgs A) {
    extern __shared__ __align__(128) unsigned char stages[];
    __shared__ uint64_t full[MAX_STAGES], empty[MAX_STAGES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = A.sp.nstages;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], WARPS); }
        mbar_fence_init();
    }
    __syncthreads();
    if (warp == SERVICE_WARP) {
        if (lane == 0) {
            int k = 0;
            for (int64_t tile = blockIdx.x; tile < A.ntiles; tile += gridDim.x, k++) {
                const int s = k % S;
                mbar_wait(&empty[s], ((k / S) & 1) ^ 1);
                stage_issue(A.sp, stages + (size_t)s * A.sp.stage_bytes, &full[s], tile, TILE, A.n);
            }
        }
        return;
    }
    ProjectSink sink;
    sink.outs = A.outs;
    Vm st;
    int k = 0;
    for (int64_t tile = blockIdx.x; tile < A.ntiles; tile += gridDim.x, k++) {
        const int s = k % S;
        mbar_wait(&full[s], (k / S) & 1);
        RowCtx rc;
        rowctx_init(rc, tile, TILE, A.n, A.err, stages + (size_t)s * A.sp.stage_bytes);
        run(A.prog, 0, A.prog.ninsn, st, rc, sink);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
}

// FilterExec + ProjectionExec, single pass. Per tile k the consumer warps run step A (predicate ->
// selection mask, ballot/popc ranks, per-warp totals) one tile AHEAD of step B (cross-block prefix,
// projection, compacted stores); the tile's columns wait in their shared-memory stage in between.
// The service warp turns the per-warp totals of a tile into its global exclusive prefix (decoupled
// look-back over tile descriptors in HBM) while the consumers are busy with step A of the next
// tile, so the L2 round trips of the look-back stay off the critical path; in between it keeps the
// stage ring full (tickets are taken in look-back order).
__global__ void __launch_bounds__(THREADS, 1) k_filter_project(const __grid_constant__ OpArgs A) {
    extern __shared__ __align__(128) unsigned char stages[];
    __shared__ uint64_t full[MAX_STAGES], empty[MAX_STAGES], agg_ready[MAX_STAGES], prefix_ready[MAX_STAGES];
    __shared__ long long tile_of[MAX_STAGES];
    __shared__ unsigned long long prefix[MAX_STAGES];
    __shared__ int wtot[MAX_STAGES][WARPS];
    __shared__ int tot[MAX_STAGES], arrived[MAX_STAGES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = A.sp.nstages;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) {
            mbar_init(&full[s], 1); mbar_init(&empty[s], WARPS);
            mbar_init(&agg_ready[s], 1); mbar_init(&prefix_ready[s], 1);
            tot[s] = 0; arrived[s] = 0;
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == SERVICE_WARP) {
        int kp = 0, kl = 0;                    // next tile slot to produce / to resolve
        bool prod_done = false, lb_done = false;
        // the next ticket is always requested one step early: the L2 round trip of the atomic overlaps
        // the TMA issue and the look-back of the current step (ticket order = look-back order)
        long long next_ticket = 0;
        if (lane == 0) next_ticket = (long long)atomicAdd(A.ticket, 1u);
        while (!prod_done || !lb_done) {
            bool did = false;
            // (1) start the descriptor loads of the tile waiting for its prefix
            bool lb_pending = false; long long lb_tile = 0; int lb_s = 0;
            unsigned long long d[4] = {0, 0, 0, 0};
            if (!lb_done) {
                lb_s = kl % S;
                int go = 0;
                if (lane == 0) go = mbar_test(&agg_ready[lb_s], (kl / S) & 1) ? 1 : 0;
                go = __shfl_sync(0xffffffffu, go, 0);
                if (go) {
                    lb_tile = tile_of[lb_s];
                    if (lb_tile >= A.ntiles) { lb_done = true; kl++; did = true; }
                    else { lb_pending = true; lb_load(A.tile_desc, lb_tile - 1, d); }
                }
            }
            // (2) keep the stage ring full
            if (!prod_done) {
                const int s = kp % S;
                int go = 0;
                if (lane == 0) go = mbar_test(&empty[s], ((kp / S) & 1) ^ 1) ? 1 : 0;
                go = __shfl_sync(0xffffffffu, go, 0);
                if (go) {
                    int end = 0;
                    if (lane == 0) {
                        const long long tile = next_ticket;
                        tile_of[s] = tile;
                        if (tile >= A.ntiles) { mbar_arrive(&full[s]); end = 1; }
                        else {
                            next_ticket = (long long)atomicAdd(A.ticket, 1u);
                            stage_issue(A.sp, stages + (size_t)s * A.sp.stage_bytes, &full[s], tile, TILE, A.n);
                        }
                    }
                    end = __shfl_sync(0xffffffffu, end, 0);
                    if (end) prod_done = true;
                    kp++; did = true;
                }
            }
            // (3) fold the descriptors; publish the inclusive prefix and hand the exclusive one to step B
            if (lb_pending) {
                unsigned long long excl = 0;
                if (lb_finish(A.tile_desc, lb_tile, d, &excl)) {
                    if (lane == 0) {
                        const unsigned long long t = (unsigned long long)tot[lb_s];
                        if (lb_tile > 0) A.tile_desc[lb_tile] = LB_INCL | (excl + t);
                        prefix[lb_s] = excl;
                        if (lb_tile == A.ntiles - 1) *A.out_count = excl + t;
                        tot[lb_s] = 0; arrived[lb_s] = 0;      // consumers are done with them until the stage is reused
                        mbar_arrive(&prefix_ready[lb_s]);
                    }
                    kl++; did = true;
                }
            }
            if (!did) __nanosleep(40);
        }
        return;
    }

    CompactSink sink;
    sink.outs = A.outs;
    Vm st;
    const uint32_t lt = (1u << lane) - 1u;
    // ranks in row order (chunk, lane, pair element) from a selection mask: ballot + popc
    auto ranks_of = [&](uint32_t sel, int (&rank)[R]) -> int {
        int wt = 0;
#pragma unroll
        for (int j = 0; j < NCHUNK; j++) {
            const uint32_t s0 = (sel >> (2 * j)) & 1u, s1 = (sel >> (2 * j + 1)) & 1u;
            const uint32_t b0 = __ballot_sync(0xffffffffu, s0), b1 = __ballot_sync(0xffffffffu, s1);
            const int below = __popc(b0 & lt) + __popc(b1 & lt);
            rank[2 * j] = wt + below;
            rank[2 * j + 1] = wt + below + (int)s0;
            wt += __popc(b0) + __popc(b1);
        }
        return wt;
    };
    // Step A runs up to LOOKAHEAD tiles ahead of step B; a pending tile is remembered by its selection
    // mask only (ranks are recomputed with ballots), so the look-back latency of tile k hides behind
    // the predicate work of tiles k+1..k+LOOKAHEAD.
    constexpr int LOOKAHEAD = 3;
    uint32_t q_sel[LOOKAHEAD]; long long q_tile[LOOKAHEAD];
#pragma unroll
    for (int i = 0; i < LOOKAHEAD; i++) { q_sel[i] = 0; q_tile[i] = -1; }
    const int D = min(LOOKAHEAD, S - 1);           // tiles in flight between A and B (S >= 2)
    auto step_b = [&](int kb, uint32_t p_sel, long long p_tile) {
        const int p_s = kb % S;
        RowCtx rc;
        rowctx_init(rc, p_tile, TILE, A.n, A.err, stages + (size_t)p_s * A.sp.stage_bytes);
        rc.active = p_sel;        // projection errors only count on surviving rows (FilterExec runs first)
        sink.sel = p_sel;
        ranks_of(p_sel, sink.rank);
        mbar_wait(&prefix_ready[p_s], (kb / S) & 1);
        int woff = 0;
#pragma unroll
        for (int w = 0; w < WARPS; w++) { int x = wtot[p_s][w]; if (w < warp) woff += x; }
        sink.base = (long long)prefix[p_s] + woff;
        if (A.selvec) {
#pragma unroll
            for (int r = 0; r < R; r++)
                if ((p_sel >> r) & 1u) A.selvec[sink.base + sink.rank[r]] = (int32_t)(rc.row0(r >> 1) + (r & 1));
        }
        run(A.prog, A.sel_end, A.prog.ninsn, st, rc, sink);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[p_s]);
    };
    int k = 0;
    for (;; k++) {
        const int s = k % S;
        mbar_wait(&full[s], (k / S) & 1);
        const long long tile = tile_of[s];
        if (tile >= A.ntiles) {
            if (warp == 0 && lane == 0) mbar_arrive(&agg_ready[s]);     // end sentinel: wake the service warp
            break;
        }
        // ---- step A(k): predicate -> selection mask -> warp total -> tile aggregate
        RowCtx rc;
        rowctx_init(rc, tile, TILE, A.n, A.err, stages + (size_t)s * A.sp.stage_bytes);
        sink.sel = 0;
        run(A.prog, 0, A.sel_end, st, rc, sink);
        const uint32_t c_sel = sink.sel;
        int wt = __popc(c_sel);
#pragma unroll
        for (int o = 16; o; o >>= 1) wt += __shfl_xor_sync(0xffffffffu, wt, o);
        if (lane == 0) {
            wtot[s][warp] = wt;
            atomicAdd(&tot[s], wt);
            __threadfence_block();
            if (atomicAdd(&arrived[s], 1) == WARPS - 1) {
                // last warp of the tile: publish the aggregate right away so that no other block's
                // look-back ever waits on this block's service warp
                __threadfence_block();
                const unsigned long long t = (unsigned long long)atomicAdd(&tot[s], 0);
                A.tile_desc[tile] = (tile == 0 ? LB_INCL : LB_PART) | t;
                mbar_arrive(&agg_ready[s]);
            }
        }
        __syncwarp();
        // ---- step B(k - D) if that tile exists; then remember tile k in the queue slot k % LOOKAHEAD
        if (k >= D) {
#pragma unroll
            for (int i = 0; i < LOOKAHEAD; i++)
                if (i == (k - D) % LOOKAHEAD) step_b(k - D, q_sel[i], q_tile[i]);
        }
#pragma unroll
        for (int i = 0; i < LOOKAHEAD; i++)
            if (i == k % LOOKAHEAD) { q_sel[i] = c_sel; q_tile[i] = tile; }
    }
    // drain: tiles k-D .. k-1 still owe their step B
    for (int kb = max(0, k - D); kb < k; kb++) {
#pragma unroll
        for (int i = 0; i < LOOKAHEAD; i++)
            if (i == kb % LOOKAHEAD) step_b(kb, q_sel[i], q_tile[i]);
    }
}

// ---- gathers by selection vector (Utf8 pass-through columns and kq_filter) --------------------------------------
template <typename T>
__global__ void k_gather_fixed(const T* __restrict__ in, const uint32_t* __restrict__ in_valid, const int32_t* __restrict__ sel,
                               const unsigned long long* __restrict__ d_count, T* __restrict__ out, uint32_t* out_valid) {
    long long m = (long long)*d_count;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        int32_t row = sel[i];
        out[i] = in[row];
        if (out_valid && ((in_valid[row >> 5] >> (row & 31)) & 1u)) atomicOr(out_valid + (i >> 5), 1u << (i & 31));
    }
}
__global__ void k_gather_bits(const uint32_t* __restrict__ in, const uint32_t* __restrict__ in_valid, const int32_t* __restrict__ sel,
                              const unsigned long long* __restrict__ d_count, uint32_t* out, uint32_t* out_valid) {
    long long m = (long long)*d_count;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        int32_t row = sel[i];
        if ((in[row >> 5] >> (row & 31)) & 1u) atomicOr(out + (i >> 5), 1u << (i & 31));
        if (out_valid && ((in_valid[row >> 5] >> (row & 31)) & 1u)) atomicOr(out_valid + (i >> 5), 1u << (i & 31));
    }
}

// Utf8 gather: lengths of the selected rows -> exclusive prefix -> output offsets (kq_scan.cuh).
struct GatherLen {
    const int32_t* in_off; const uint32_t* in_valid; const int32_t* sel; uint32_t* out_valid;
    __device__ __forceinline__ int operator()(long long i) const {
        int32_t row = sel[i];
        if (out_valid && ((in_valid[row >> 5] >> (row & 31)) & 1u)) atomicOr(out_valid + (i >> 5), 1u << (i & 31));
        return in_off[row + 1] - in_off[row];
    }
};
__global__ void k_utf8_gather_bytes(const int32_t* __restrict__ in_off, const uint8_t* __restrict__ in_data, const int32_t* __restrict__ sel,
                                    const unsigned long long* __restrict__ d_count, const int32_t* __restrict__ out_off, uint8_t* out_data) {
    long long m = (long long)*d_count;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        int32_t row = sel[i];
        int a = in_off[row], len = in_off[row + 1] - a, o = out_off[i];
        for (int b = 0; b < len; b++) out_data[o + b] = in_data[a + b];
    }
}

// Shared memory available for the stage ring of one CTA when `ctas` CTAs share an SM.
int stage_budget(kq_ctx* ctx, int ctas) {
    return (ctx->max_smem_optin + 1024) / ctas - 1024 - 2048;   // 1 KB/CTA reserved by the driver, 2 KB static
}
int blocks_per_sm(const void* fn, int threads, int smem) {
    int nb = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, threads, smem) != cudaSuccess || nb < 1) nb = 1;
    return nb;
}

int launch_check(kq_ctx* ctx, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return kq_cuda_fail(ctx, e, what);
    ctx->launches++;
    return KQ_OK;
}

int small_grid(kq_ctx* ctx, int64_t items) {
    int64_t g = (items + 255) / 256, cap = (int64_t)ctx->sm_count * 8;
    return (int)std::max<int64_t>(1, std::min(g, cap));
}

// Gather one whole column by a device-resident selection vector with a device-resident count.
int gather_column(kq_ctx* ctx, kq_col* in, const int32_t* sel, kq_lazy_count* lazy, int64_t cap_rows, kq_col** out) {
    kq_col* c = nullptr;
    bool nullable = in->validity != nullptr;
    KQ_RET(kq_col_new(ctx, in->type, cap_rows, nullable, in->type == KQ_UTF8 ? in->data_bytes : 0, &c));
    c->n = -1; c->lazy = lazy; lazy->rc.fetch_add(1);
    if (nullable) cudaMemsetAsync(c->validity, 0, (size_t)((cap_rows + 63) / 64) * 8, ctx->stream);
    int g = small_grid(ctx, cap_rows);
    int st = KQ_OK;
    switch (in->type) {
        case KQ_F64: case KQ_I64:
            k_gather_fixed<uint64_t><<<g, 256, 0, ctx->stream>>>((const uint64_t*)in->data, in->validity, sel, lazy->d_slot, (uint64_t*)c->data, c->validity);
            st = launch_check(ctx, "k_gather_fixed"); break;
        case KQ_DATE32: case KQ_I32:
            k_gather_fixed<uint32_t><<<g, 256, 0, ctx->stream>>>((const uint32_t*)in->data, in->validity, sel, lazy->d_slot, (uint32_t*)c->data, c->validity);
            st = launch_check(ctx, "k_gather_fixed"); break;
        case KQ_BOOL:
            cudaMemsetAsync(c->data, 0, (size_t)((cap_rows + 63) / 64) * 8, ctx->stream);
            k_gather_bits<<<g, 256, 0, ctx->stream>>>((const uint32_t*)in->data, in->validity, sel, lazy->d_slot, (uint32_t*)c->data, c->validity);
            st = launch_check(ctx, "k_gather_bits"); break;
        case KQ_UTF8: {
            int64_t ntiles = (cap_rows + SCAN_TILE - 1) / SCAN_TILE + 1;
            unsigned long long* scratch = nullptr;
            st = kq_dev_alloc(ctx, (size_t)(ntiles + 2) * 8, (void**)&scratch);
            if (st != KQ_OK) break;
            cudaMemsetAsync(scratch, 0, (size_t)(ntiles + 2) * 8, ctx->stream);
            if ((st = kq_dev_alloc(ctx, 8, (void**)&c->d_utf8_bytes)) != KQ_OK) break;
            c->data_bytes = -1;
            int sg = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, (int64_t)ctx->sm_count * 4));
            GatherLen gl{in->offsets, in->validity, sel, c->validity};
            k_exclusive_offsets<GatherLen><<<sg, 256, 0, ctx->stream>>>(gl, lazy->d_slot, c->offsets, scratch + 2, (unsigned int*)scratch, c->d_utf8_bytes);
            st = launch_check(ctx, "k_exclusive_offsets");
            if (st == KQ_OK) {
                k_utf8_gather_bytes<<<g, 256, 0, ctx->stream>>>(in->offsets, (const uint8_t*)in->data, sel, lazy->d_slot, c->offsets, (uint8_t*)c->data);
                st = launch_check(ctx, "k_utf8_gather_bytes");
            }
            kq_dev_free(ctx, scratch);
            break;
        }
        default: st = kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "unknown column type");
    }
    if (st != KQ_OK) { kq_column_free(c); return st; }
    *out = c;
    return KQ_OK;
}

// Shared implementation of kq_project / kq_filter_project / kq_filter / kq_expr_evaluate.
int run_operator(kq_ctx* ctx, kq_expr* pred, kq_expr* const* exprs, int nexprs, kq_batch* input, bool all_columns,
                 kq_batch** out, kq_col** selection) {
    if (!ctx || !input || !out || nexprs < 0) return KQ_ERR_ILLEGAL_ARGUMENT;
    cudaSetDevice(ctx->device);
    int64_t n; KQ_RET(kq_batch_resolve_rows(ctx, input, &n));
    for (kq_col* c : input->cols) KQ_RET(kq_col_resolve_rows(ctx, c, nullptr));

    // expressions of a filter-all: one ColumnExpression per input field
    std::vector<kq_expr*> owned;
    std::vector<kq_expr*> ex(exprs, exprs + nexprs);
    if (all_columns) for (int i = 0; i < (int)input->cols.size(); i++) { owned.push_back(kq_expr_column(i)); ex.push_back(owned.back()); }
    auto cleanup = [&]() { for (kq_expr* e : owned) kq_expr_free(e); };

    KqCompiler cc;
    int st = cc.begin(ctx, input);
    std::vector<kq_col*> outs(ex.size(), nullptr);
    auto fail = [&](int s) { for (kq_col* c : outs) kq_column_free(c); cleanup(); return s; };
    if (st != KQ_OK) return fail(st);

    OpArgs A;
    memset(&A, 0, sizeof A);
    A.n = n; A.ntiles = (n + TILE - 1) / TILE; A.err = ctx->d_err;
    kq_lazy_count* lazy = nullptr;
    if (pred) {
        int t; bool nl;
        if ((st = cc.value(pred, &t, &nl)) != KQ_OK) return fail(st);
        if (t != KQ_BOOL) return fail(kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "filter predicate is not Bool"));
        if ((st = cc.sink(O_SET_SEL, 0)) != KQ_OK) return fail(st);
        A.sel_end = cc.pc();
        lazy = kq_lazy_new(ctx);
        if (!lazy) return fail(kq_fail(ctx, KQ_ERR_OUT_OF_MEMORY, "lazy count"));
    }
    auto fail2 = [&](int s) { if (lazy) kq_lazy_release(ctx, lazy); return fail(s); };

    // classify outputs: gather (Utf8 pass-through below a filter, or any column of kq_filter with
    // too many fused outputs), alias (bare column without a filter, rule R4), or fused VM output.
    std::vector<int> mode(ex.size(), 0);   // 0 = VM, 1 = alias, 2 = gather by selection vector
    int nvm = 0;
    for (size_t k = 0; k < ex.size(); k++) {
        int bc = KqCompiler::bare_column(ex[k]);
        int t; bool nl;
        if ((st = cc.infer(ex[k], &t, &nl)) != KQ_OK) return fail2(st);
        if (bc >= 0 && !pred) mode[k] = 1;
        else if (bc >= 0 && (t == KQ_UTF8 || nvm >= MAX_OUT)) mode[k] = 2;
        else {
            if (t == KQ_UTF8) return fail2(kq_fail(ctx, KQ_ERR_UNSUPPORTED, "expressions cannot produce Utf8 values (only column pass-through)"));
            if (nvm >= MAX_OUT) return fail2(kq_fail(ctx, KQ_ERR_UNSUPPORTED, "more than %d computed outputs in one kernel", MAX_OUT));
            mode[k] = 0;
            int t2; bool n2;
            if ((st = cc.value(ex[k], &t2, &n2)) != KQ_OK) return fail2(st);
            kq_col* c = nullptr;
            if ((st = kq_col_new(ctx, t2, n, n2, 0, &c)) != KQ_OK) return fail2(st);
            outs[k] = c;
            if (pred) {
                c->n = -1; c->lazy = lazy; lazy->rc.fetch_add(1);
                if (c->validity) cudaMemsetAsync(c->validity, 0, (size_t)((n + 63) / 64) * 8, ctx->stream);
                if (t2 == KQ_BOOL) cudaMemsetAsync(c->data, 0, (size_t)((n + 63) / 64) * 8, ctx->stream);
            }
            A.outs[nvm].data = c->data; A.outs[nvm].validity = c->validity; A.outs[nvm].type = t2;
            if ((st = cc.sink(O_EMIT, nvm)) != KQ_OK) return fail2(st);
            nvm++;
        }
    }
    A.nout = nvm;
    cc.plan_stages(stage_budget(ctx, 1), pred ? 3 : 2, TILE, &A.sp);
    A.prog = cc.prog;
    const int smem = A.sp.nstages * A.sp.stage_bytes;

    bool need_sel = selection != nullptr;
    for (int m : mode) need_sel |= (m == 2);
    kq_col* selcol = nullptr;
    if (pred && need_sel) {
        if ((st = kq_col_new(ctx, KQ_I32, n, false, 0, &selcol)) != KQ_OK) return fail2(st);
        selcol->n = -1; selcol->lazy = lazy; lazy->rc.fetch_add(1);
        A.selvec = (int32_t*)selcol->data;
    }
    auto fail3 = [&](int s) { kq_column_free(selcol); return fail2(s); };

    if (pred) {
        unsigned long long* scratch = nullptr;     // [0]: ticket, [2..]: tile descriptors
        if ((st = kq_dev_alloc(ctx, (size_t)(A.ntiles + 2) * 8, (void**)&scratch)) != KQ_OK) return fail3(st);
        cudaMemsetAsync(scratch, 0, (size_t)(A.ntiles + 2) * 8, ctx->stream);
        A.ticket = (unsigned int*)scratch;
        A.tile_desc = scratch + 2;
        A.out_count = lazy->d_slot;
        if (n > 0) {
            cudaFuncSetAttribute(k_filter_project, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            int bps = blocks_per_sm((const void*)k_filter_project, THREADS, smem);
            int grid = (int)std::min<int64_t>(A.ntiles, (int64_t)ctx->sm_count * bps);
            k_filter_project<<<grid, THREADS, smem, ctx->stream>>>(A);
            if ((st = launch_check(ctx, "k_filter_project")) != KQ_OK) { kq_dev_free(ctx, scratch); return fail3(st); }
        }
        kq_dev_free(ctx, scratch);
        for (size_t k = 0; k < ex.size(); k++) {
            if (mode[k] != 2) continue;
            kq_col* in = input->cols[(size_t)KqCompiler::bare_column(ex[k])];
            if ((st = gather_column(ctx, in, (const int32_t*)selcol->data, lazy, n, &outs[k])) != KQ_OK) return fail3(st);
        }
        // the row count travels back asynchronously; it is only waited for when somebody asks
        cudaMemcpyAsync(lazy->h_slot, lazy->d_slot, 8, cudaMemcpyDeviceToHost, ctx->stream);
        cudaEventRecord(lazy->ev, ctx->stream);
    } else {
        for (size_t k = 0; k < ex.size(); k++)
            if (mode[k] == 1) { outs[k] = input->cols[(size_t)KqCompiler::bare_column(ex[k])]; outs[k]->rc.fetch_add(1); }
        if (nvm > 0 && n > 0) {
            cudaFuncSetAttribute(k_project, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            int bps = blocks_per_sm((const void*)k_project, THREADS, smem);
            int grid = (int)std::min<int64_t>(A.ntiles, (int64_t)ctx->sm_count * bps);
            k_project<<<grid, THREADS, smem, ctx->stream>>>(A);
            if ((st = launch_check(ctx, "k_project")) != KQ_OK) return fail3(st);
        }
    }

    kq_batch* b = new kq_batch();
    b->ctx = ctx;
    b->cols = outs;
    if (pred) { b->n = -1; b->lazy = lazy; } else b->n = n;
    *out = b;
    if (selection) *selection = selcol; else kq_column_free(selcol);
    cleanup();
    return KQ_OK;
}

}  // namespace

extern "C" {

int kq_project(kq_ctx* ctx, kq_expr* const* exprs, int nexprs, kq_batch* input, kq_batch** out) {
    return run_operator(ctx, nullptr, exprs, nexprs, input, false, out, nullptr);
}

int kq_filter_project(kq_ctx* ctx, kq_expr* pred, kq_expr* const* exprs, int nexprs, kq_batch* input, kq_batch** out) {
    if (!pred) return kq_fail(ctx, KQ_ERR_ILLEGAL_ARGUMENT, "kq_filter_project needs a predicate");
    return run_operator(ctx, pred, exprs, nexprs, input, false, out, nullptr);
}

int kq_filter(kq_ctx* ctx, kq_expr* pred, kq_batch* input, kq_batch** out, kq_col** selection) {
    if (!pred) return kq_fail(ctx, KQ_ERR_ILLEGAL_ARGUMENT, "kq_filter needs a predicate");
    return run_operator(ctx, pred, nullptr, 0, input, true, out, selection);
}

int kq_expr_evaluate(kq_ctx* ctx, kq_expr* e, kq_batch* input, kq_col** out) {
    if (!out) return KQ_ERR_ILLEGAL_ARGUMENT;
    kq_batch* b = nullptr;
    kq_expr* ex[1] = {e};
    KQ_RET(run_operator(ctx, nullptr, ex, 1, input, false, &b, nullptr));
    *out = b->cols[0];
    (*out)->rc.fetch_add(1);
    kq_batch_free(b);
    return KQ_OK;
}

int kq_filter_project_host(kq_ctx* ctx, kq_expr* pred, kq_expr* const* exprs, int nexprs, int ncols, const int* types,
                           const uint8_t* const* validity, const void* const* data, int64_t n, void* const* out_data,
                           uint8_t* const* out_validity, int64_t* out_rows) {
    // v1: upload -> fused kernel -> download, sequential (chunked overlap is a later optimisation)
    if (!ctx || ncols < 0) return KQ_ERR_ILLEGAL_ARGUMENT;
    std::vector<kq_col*> cols((size_t)ncols, nullptr);
    int st = KQ_OK;
    for (int i = 0; i < ncols && st == KQ_OK; i++) {
        if (types[i] == KQ_UTF8) st = kq_fail(ctx, KQ_ERR_UNSUPPORTED, "kq_filter_project_host: Utf8 columns not supported");
        else st = kq_column_upload(ctx, types[i], n, validity ? validity[i] : nullptr, nullptr, data[i], 0, &cols[(size_t)i]);
    }
    kq_batch *in = nullptr, *res = nullptr;
    if (st == KQ_OK) st = kq_batch_create(ctx, cols.data(), ncols, n, &in);
    if (st == KQ_OK) st = kq_filter_project(ctx, pred, exprs, nexprs, in, &res);
    int64_t m = 0;
    if (st == KQ_OK) st = kq_batch_num_rows(ctx, res, &m);
    for (int k = 0; k < nexprs && st == KQ_OK; k++)
        st = kq_column_download(ctx, res->cols[(size_t)k], out_validity ? out_validity[k] : nullptr, nullptr, out_data[k]);
    if (st == KQ_OK && out_rows) *out_rows = m;
    kq_batch_free(res); kq_batch_free(in);
    for (kq_col* c : cols) kq_column_free(c);
    return st;
}

}  // extern "C"
