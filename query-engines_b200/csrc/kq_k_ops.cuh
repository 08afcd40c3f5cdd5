// kq_k_ops.cuh — kernel skeletons of ProjectionExec and the fused FilterExec + ProjectionExec,
// specialised per query: the generated `struct Q` (kq_codegen.cu) supplies the expression code.
//
//   kq_project         ProjectionExec.execute for one batch (Main.kt:589-594): every expression of the
//                      projection evaluated in one pass, one 128-bit store per pair of output rows.
//   kq_filter_project  FilterExec (absent from the reference) fused with the
//                      projection above it: predicate -> warp ballot/popc ranks -> ordered cross-block
//                      prefix (decoupled look-back) -> compacted stores, all in a single pass over the
//                      input (algorithmic bytes only: each input column read once, each output row
//                      written once).
//
// One CTA per SM: KQ_WARPS consumer warps evaluate R rows per thread out of shared memory, one
// service warp streams tiles in with TMA bulk copies (and, in the filter kernel, resolves the
// cross-block prefix). Bytes in flight come from the stage ring, not from occupancy.
#pragma once

#include "kq_rt.cuh"
#include "kq_scan.cuh"

namespace kq {

constexpr int WARPS = KQ_WARPS;
constexpr int BLOCK = WARPS * 32;
constexpr int TILE = WARPS * WARP_ROWS;
constexpr int SERVICE_WARP = WARPS;
constexpr int THREADS = BLOCK + 32;

// ---- sinks: where the generated projection code hands its results --------------------------------------------------
struct ProjectSink {
    const DOut* outs;
    template <int TYPE>
    __device__ __forceinline__ void emit(int k, const uint64_t (&v)[R], uint32_t ok, bool nullable, const RowCtx& rc) const {
        const DOut o = outs[k];
        if constexpr (TYPE == KQT_DATE32 || TYPE == KQT_I32) {
#pragma unroll
            for (int j = 0; j < NCHUNK; j++) {
                const int64_t r0 = rc.row0(j);
                if (rc.full || r0 < rc.n) stg_v2(reinterpret_cast<uint2*>(o.data) + (r0 >> 1), make_uint2((uint32_t)v[2 * j], (uint32_t)v[2 * j + 1]));
            }
        } else {
#pragma unroll
            for (int j = 0; j < NCHUNK; j++) {
                const int64_t r0 = rc.row0(j);
                if (rc.full || r0 < rc.n)
                    stg_v4(reinterpret_cast<uint4*>(o.data) + (r0 >> 1),
                           make_uint4((uint32_t)v[2 * j], (uint32_t)(v[2 * j] >> 32), (uint32_t)v[2 * j + 1], (uint32_t)(v[2 * j + 1] >> 32)));
            }
        }
        if (nullable) {
#pragma unroll
            for (int j = 0; j < NCHUNK; j++) store_chunk_bits(o.validity, rc, j, ok & rc.inr);
        }
    }
    __device__ __forceinline__ void emit_bool(int k, uint32_t truth, uint32_t ok, bool nullable, const RowCtx& rc) const {
        const DOut o = outs[k];
#pragma unroll
        for (int j = 0; j < NCHUNK; j++) store_chunk_bits(reinterpret_cast<uint32_t*>(o.data), rc, j, truth & rc.inr);
        if (nullable) {
#pragma unroll
            for (int j = 0; j < NCHUNK; j++) store_chunk_bits(o.validity, rc, j, ok & rc.inr);
        }
    }
};

struct CompactSink {
    const DOut* outs;
    uint32_t sel;
    int rank[R];
    long long base;
    template <int TYPE>
    __device__ __forceinline__ void emit(int k, const uint64_t (&v)[R], uint32_t ok, bool nullable, const RowCtx&) const {
        const DOut o = outs[k];
#pragma unroll
        for (int r = 0; r < R; r++) {
            if ((sel >> r) & 1u) {
                const long long pos = base + rank[r];
                if constexpr (TYPE == KQT_DATE32 || TYPE == KQT_I32) reinterpret_cast<uint32_t*>(o.data)[pos] = (uint32_t)v[r];
                else reinterpret_cast<uint64_t*>(o.data)[pos] = v[r];
                if (nullable && ((ok >> r) & 1u)) atomicOr(o.validity + (pos >> 5), 1u << (pos & 31));
            }
        }
    }
    __device__ __forceinline__ void emit_bool(int k, uint32_t truth, uint32_t ok, bool nullable, const RowCtx&) const {
        const DOut o = outs[k];
#pragma unroll
        for (int r = 0; r < R; r++) {
            if ((sel >> r) & 1u) {
                const long long pos = base + rank[r];
                if ((truth >> r) & 1u) atomicOr(reinterpret_cast<uint32_t*>(o.data) + (pos >> 5), 1u << (pos & 31));
                if (nullable && ((ok >> r) & 1u)) atomicOr(o.validity + (pos >> 5), 1u << (pos & 31));
            }
        }
    }
};

#ifdef KQ_KERNEL_PROJECT
// ProjectionExec for one batch (Main.kt:589-594).
extern "C" __global__ void __launch_bounds__(THREADS, 1) kq_project(const __grid_constant__ OpArgs A) {
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
    int k = 0;
    for (int64_t tile = blockIdx.x; tile < A.ntiles; tile += gridDim.x, k++) {
        const int s = k % S;
        mbar_wait(&full[s], (k / S) & 1);
        RowCtx rc;
        rowctx_init(rc, warp, tile, TILE, A.n, A.err, stages + (size_t)s * A.sp.stage_bytes);
        Q::project(A.q, rc, sink);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
}
#endif  // KQ_KERNEL_PROJECT

#ifdef KQ_KERNEL_FILTER
// FilterExec + ProjectionExec, single pass. Per tile k the consumer warps run step A (predicate ->
// selection mask, per-warp totals) ahead of step B (cross-block prefix, projection, compacted
// stores); the tile's columns wait in their shared-memory stage in between. The service warp turns
// the per-warp totals of a tile into its global exclusive prefix (decoupled look-back over tile
// descriptors in HBM) while the consumers are busy with step A of the next tiles, so the L2 round
// trips of the look-back stay off the critical path; in between it keeps the stage ring full
// (tickets are taken in look-back order).
extern "C" __global__ void __launch_bounds__(THREADS, 1) kq_filter_project(const __grid_constant__ OpArgs A) {
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
        rowctx_init(rc, warp, p_tile, TILE, A.n, A.err, stages + (size_t)p_s * A.sp.stage_bytes);
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
        Q::project(A.q, rc, sink);
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
        rowctx_init(rc, warp, tile, TILE, A.n, A.err, stages + (size_t)s * A.sp.stage_bytes);
        const uint32_t c_sel = Q::pred(A.q, rc) & rc.inr;     // TRUE only: a null predicate drops the row (rule E3)
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
#endif  // KQ_KERNEL_FILTER

}  // namespace kq
