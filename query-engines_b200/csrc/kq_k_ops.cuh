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

#ifndef KQ_L2_PREFETCH
#define KQ_L2_PREFETCH 2          // tiles (per CTA) the filter producer asks the L2 to fetch ahead of the stage ring (measured: 0 -> 4.97, 1..4 -> 5.4 TB/s, 8+ thrashes)
#endif
constexpr int WARPS = KQ_WARPS;
constexpr int BLOCK = WARPS * 32;
constexpr int TILE = WARPS * WARP_ROWS;
// Service warps come FIRST in the CTA: the SM's warp arbiter favours high warp ids, so polling service
// warps never take issue slots from the consumer warps that share their scheduler.
constexpr int PRODUCER_WARP = 0;             // TMA producer
#ifdef KQ_KERNEL_FILTER
constexpr int LOOKBACK_WARP = 1;             // first of NLB cross-block prefix resolvers (local tile kl -> warp kl % NLB)
constexpr int NLB = 2;
constexpr int NSERVICE = 1 + NLB;
#else
constexpr int NSERVICE = 1;
#endif
constexpr int THREADS = BLOCK + 32 * NSERVICE;

#ifdef KQ_TRACE
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define KQ_TR(tile, slot) do { if (A.trace) A.trace[(tile) * 8 + (slot)] = gtime(); } while (0)
#else
#define KQ_TR(tile, slot) do { } while (0)
#endif

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

#ifdef KQ_KERNEL_FILTER
// FilterExec: the projection of the surviving rows of a tile is parked, already compacted, in a
// warp-private shared-memory ring until the tile's global output offset is known. A tile takes only
// as many ring rows as it has survivors, so at low selectivity many tiles can be pending.
constexpr int CAP = KQ_STASH_ROWS;            // ring capacity in rows (power of two, >= WARP_ROWS)
struct StashSink {
    uint64_t* val;            // this warp's ring: [NOUT][CAP] values
    uint8_t* flag;            //                   [NOUT][CAP] validity flags (only if some output is nullable)
    uint32_t sel;
    uint32_t head;            // ring position of the tile's first surviving row
    int rank[R];
    template <int TYPE>
    __device__ __forceinline__ void emit(int k, const uint64_t (&v)[R], uint32_t ok, bool nullable, const RowCtx&) const {
#pragma unroll
        for (int r = 0; r < R; r++) {
            if ((sel >> r) & 1u) {
                const uint32_t at = k * CAP + ((head + rank[r]) & (CAP - 1));
                val[at] = v[r];
                if (nullable) flag[at] = (uint8_t)((ok >> r) & 1u);
            }
        }
    }
    __device__ __forceinline__ void emit_bool(int k, uint32_t truth, uint32_t ok, bool nullable, const RowCtx&) const {
#pragma unroll
        for (int r = 0; r < R; r++) {
            if ((sel >> r) & 1u) {
                const uint32_t at = k * CAP + ((head + rank[r]) & (CAP - 1));
                val[at] = (truth >> r) & 1u;
                if (nullable) flag[at] = (uint8_t)((ok >> r) & 1u);
            }
        }
    }
};

// OR `bits` (32 consecutive output rows starting at bit position pos) into a pre-zeroed bitmap
__device__ __forceinline__ void or_bits_at(uint32_t* bitmap, long long pos, uint32_t bits) {
    if (!bits) return;
    const int sh = (int)(pos & 31);
    atomicOr(bitmap + (pos >> 5), bits << sh);
    if (sh && (bits >> (32 - sh))) atomicOr(bitmap + (pos >> 5) + 1, bits >> (32 - sh));
}

// Copy output K of one tile from the warp's ring (rows tail .. tail+cnt) to rows [base, base + cnt) of the output column.
template <int K>
__device__ __forceinline__ void stash_copy_out(const DOut* outs, const uint64_t* val, const uint8_t* flag, uint32_t tail, long long base, int cnt, int lane) {
    if constexpr (K < Q::NOUT) {
        constexpr int TYPE = Q::OUT_TYPE[K];
        constexpr bool NULLABLE = Q::OUT_NULLABLE[K];
        const DOut o = outs[K];
        for (int i0 = 0; i0 < cnt; i0 += 32) {
            const int i = i0 + lane;
            const bool in = i < cnt;
            const uint32_t at = K * CAP + ((tail + i) & (CAP - 1));
            uint64_t v = 0;
            if (in) v = val[at];
            if constexpr (TYPE == KQT_BOOL) {
                const uint32_t bits = __ballot_sync(0xffffffffu, in && (v & 1u));
                if (lane == 0) or_bits_at(reinterpret_cast<uint32_t*>(o.data), base + i0, bits);
            } else if constexpr (TYPE == KQT_DATE32 || TYPE == KQT_I32) {
                if (in) reinterpret_cast<uint32_t*>(o.data)[base + i] = (uint32_t)v;
            } else {
                if (in) reinterpret_cast<uint64_t*>(o.data)[base + i] = v;
            }
            if constexpr (NULLABLE) {
                const uint32_t bits = __ballot_sync(0xffffffffu, in && flag[at]);
                if (lane == 0) or_bits_at(o.validity, base + i0, bits);
            }
        }
        stash_copy_out<K + 1>(outs, val, flag, tail, base, cnt, lane);
    }
}
#endif  // KQ_KERNEL_FILTER

#ifdef KQ_KERNEL_PROJECT
// ProjectionExec for one batch (Main.kt:589-594).
extern "C" __global__ void __launch_bounds__(THREADS, 1) kq_project(const __grid_constant__ OpArgs A) {
    constexpr int S = KQ_STAGES;
    extern __shared__ __align__(128) unsigned char stages[];
    __shared__ uint64_t full[S], empty[S];
    __shared__ uint32_t drain_word;          // stage_release reads it back (always 0)
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warp = wid - NSERVICE;         // consumer warp index
    if (threadIdx.x == 0) {
        drain_word = 0;
        for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], WARPS); }
        mbar_fence_init();
    }
    __syncthreads();
    if (wid == PRODUCER_WARP) {
        if (lane == 0) {
            int k = 0;
            for (int64_t tile = A.tile_begin + blockIdx.x; tile < A.tile_end; tile += gridDim.x, k++) {
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
    for (int64_t tile = A.tile_begin + blockIdx.x; tile < A.tile_end; tile += gridDim.x, k++) {
        const int s = k % S;
        mbar_wait(&full[s], (k / S) & 1);
        RowCtx rc;
        rowctx_init(rc, warp, tile, TILE, A.n, A.err, stages + (size_t)s * A.sp.stage_bytes);
        Q::project(A.q, rc, sink);
        stage_release(&empty[s], &drain_word, 0u, A.err, lane);
    }
}
#endif  // KQ_KERNEL_PROJECT

#ifdef KQ_KERNEL_FILTER
// FilterExec + ProjectionExec, single pass over the input.
//
// Consumer warps, per tile k (tiles are dealt round-robin: CTA b owns tiles b, b+G, b+2G, ...):
//   step A(k)  predicate -> selection mask -> ballot/popc ranks; the warp total goes to the tile's
//              metadata slot (the last warp to arrive publishes the tile aggregate for other CTAs);
//              the projection of the surviving rows is evaluated right away and parked, compacted, in
//              the warp's private stash ring; the input stage goes back to the TMA producer at once,
//              so the whole stage ring is prefetch depth.
//   step B(j)  for every older tile j whose exclusive prefix has arrived meanwhile: copy its rows from
//              the stash to their final position with coalesced stores. A warp only BLOCKS on a prefix
//              when its stash ring or the metadata ring is full, so CTAs can drift apart by many tiles
//              (at 25 % selectivity ~12) before anybody waits.
// The look-back warp turns tile aggregates into exclusive prefixes (decoupled look-back over per-tile
// descriptors in HBM) and hands every consumer warp its own output base.
//
// Ring sizes: S input stages; M metadata slots; a warp keeps at most DMAX = M-S-1 tiles pending — a
// warp can run at most S tiles ahead of the slowest one (stage reuse), so a slot is never rewritten
// while another warp still reads it.
extern "C" __global__ void __launch_bounds__(THREADS, 1) kq_filter_project(const __grid_constant__ OpArgs A) {
    constexpr int S = KQ_STAGES, M = KQ_META, DMAX = M - S - 1;
    constexpr int STASH_WARP = CAP * (8 * Q::NOUT + (Q::ANY_NULLABLE ? Q::NOUT : 0) + (KQ_SELVEC ? 4 : 0));   // bytes per warp
    extern __shared__ __align__(128) unsigned char smem[];      // [S input stages][WARPS stash rings]
    __shared__ uint64_t full[S], empty[S], agg_ready[M], prefix_ready[M];
    __shared__ long long tile_of[S];
    __shared__ long long wbase[M][WARPS];        // per tile: output row of each warp's first survivor
    __shared__ int wtot[M][WARPS];
    __shared__ int tot[M], arrived[M];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warp = wid - NSERVICE;         // consumer warp index
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], WARPS); }
        for (int m = 0; m < M; m++) { mbar_init(&agg_ready[m], 1); mbar_init(&prefix_ready[m], 1); tot[m] = 0; arrived[m] = 0; }
        mbar_fence_init();
    }
    __syncthreads();

    if (wid == PRODUCER_WARP) {
        // ---- TMA producer: keeps the stage ring full. All CTAs move through the table as one wavefront:
        // the predecessors a look-back needs are always tiles the other CTAs process at the same time.
        // (Handing tiles out by an atomic ticket at load-issue time lets fast CTAs bind far-ahead tiles
        // early and scrambles the order in which aggregates appear: measured 3x slower.) The launch is
        // cooperative, so every CTA a look-back may wait for is resident.
        if (lane == 0) {
            int kp = 0;
            for (long long tile = A.tile_begin + blockIdx.x; tile < A.tile_end; tile += gridDim.x, kp++) {
                const int s = kp % S;
                mbar_wait(&empty[s], ((kp / S) & 1) ^ 1);
                tile_of[s] = tile;
                KQ_TR(tile, 0);
                stage_issue(A.sp, smem + (size_t)s * A.sp.stage_bytes, &full[s], tile, TILE, A.n);
                if (KQ_L2_PREFETCH > 0) stage_prefetch_l2(A.sp, tile + (long long)KQ_L2_PREFETCH * gridDim.x, TILE, A.n);
            }
            const int s = kp % S;
            mbar_wait(&empty[s], ((kp / S) & 1) ^ 1);
            tile_of[s] = -1;
            mbar_arrive(&full[s]);
        }
        return;
    }
    if (wid >= LOOKBACK_WARP && wid < LOOKBACK_WARP + NLB) {
        // ---- look-back: tile aggregate -> exclusive prefix -> per-warp output bases. Look-backs of
        // different tiles are independent, so the NLB warps take the CTA's tiles in turn.
        int kl = wid - LOOKBACK_WARP;
        for (long long tile = A.tile_begin + blockIdx.x + (long long)kl * gridDim.x; tile < A.tile_end; tile += (long long)NLB * gridDim.x, kl += NLB) {
            const int m = kl % M;
            mbar_wait(&agg_ready[m], (kl / M) & 1);
            if (lane == 0) KQ_TR(tile, 3);
            const unsigned long long excl = lb_resolve(A.tile_desc, tile);
            int x = lane < WARPS ? wtot[m][lane] : 0, incl = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
            if (lane < WARPS) wbase[m][lane] = (long long)excl + incl - x;
            const unsigned long long t = (unsigned long long)__shfl_sync(0xffffffffu, incl, 31);
            __syncwarp();
            if (lane == 0) {
                KQ_TR(tile, 4);
                if (tile > 0) A.tile_desc[tile] = LB_INCL | (excl + t);
                if (tile == A.ntiles - 1) *A.out_count = excl + t;
                tot[m] = 0; arrived[m] = 0;      // consumers are done with them until the slot is reused
                mbar_arrive(&prefix_ready[m]);
            }
            __syncwarp();
        }
        return;
    }

    unsigned char* const ring = smem + (size_t)S * A.sp.stage_bytes + (size_t)warp * STASH_WARP;
    StashSink sink;
    sink.val = reinterpret_cast<uint64_t*>(ring);
    sink.flag = ring + CAP * 8 * Q::NOUT;
    int32_t* const selring = reinterpret_cast<int32_t*>(ring + CAP * (8 * Q::NOUT + (Q::ANY_NULLABLE ? Q::NOUT : 0)));
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t head = 0, tail = 0;             // stash ring positions (rows, monotonic)
    int kb = 0;                              // oldest tile whose rows are still in the stash
    // step B of local tile kb; `block` = wait for its prefix, otherwise only if it has already arrived
    auto step_b = [&](bool block) -> bool {
        const int m = kb % M;
        if (block) mbar_wait(&prefix_ready[m], (kb / M) & 1);
        else {
            int go = 0;
            if (lane == 0) go = mbar_test(&prefix_ready[m], (kb / M) & 1) ? 1 : 0;
            if (!__shfl_sync(0xffffffffu, go, 0)) return false;
        }
        const int cnt = wtot[m][warp];
        const long long base = wbase[m][warp];
        stash_copy_out<0>(A.outs, sink.val, sink.flag, tail, base, cnt, lane);
        if constexpr (KQ_SELVEC) {
            for (int i = lane; i < cnt; i += 32) A.selvec[base + i] = selring[(tail + i) & (CAP - 1)];
        }
        tail += cnt;
        kb++;
        return true;
    };
    int k = 0;
    for (;; k++) {
        const int s = k % S, m = k % M;
        // make room: one tile's worth of stash rows and a free metadata distance
        while (k - kb >= DMAX || head - tail + WARP_ROWS > CAP) step_b(true);
        mbar_wait(&full[s], (k / S) & 1);
        const long long tile = tile_of[s];
        if (tile < 0) break;
        if (warp == 0 && lane == 0) KQ_TR(tile, 1);
        // ---- step A(k)
        RowCtx rc;
        rowctx_init(rc, warp, tile, TILE, A.n, A.err, smem + (size_t)s * A.sp.stage_bytes);
        const uint32_t sel = Q::pred(A.q, rc) & rc.inr;       // TRUE only: a null predicate drops the row (rule E3)
        // ranks in row order (chunk, lane, pair element): ballot + popc
        int wt = 0;
#pragma unroll
        for (int j = 0; j < NCHUNK; j++) {
            const uint32_t s0 = (sel >> (2 * j)) & 1u, s1 = (sel >> (2 * j + 1)) & 1u;
            const uint32_t b0 = __ballot_sync(0xffffffffu, s0), b1 = __ballot_sync(0xffffffffu, s1);
            const int below = __popc(b0 & lt) + __popc(b1 & lt);
            sink.rank[2 * j] = wt + below;
            sink.rank[2 * j + 1] = wt + below + (int)s0;
            wt += __popc(b0) + __popc(b1);
        }
        if (lane == 0) {
            wtot[m][warp] = wt;
            atomicAdd(&tot[m], wt);
            __threadfence_block();
            if (atomicAdd(&arrived[m], 1) == WARPS - 1) {
                // last warp of the tile: publish the aggregate right away so that no other block's
                // look-back ever waits on this block's look-back warp
                __threadfence_block();
                const unsigned long long t = (unsigned long long)atomicAdd(&tot[m], 0);
                A.tile_desc[tile] = (tile == 0 ? LB_INCL : LB_PART) | t;
                KQ_TR(tile, 2);
                mbar_arrive(&agg_ready[m]);
            }
        }
        // projection of the surviving rows -> stash (errors only count on surviving rows: FilterExec runs first)
        sink.sel = sel;
        sink.head = head;
        rc.active = sel;
        Q::project(A.q, rc, sink);
        if constexpr (KQ_SELVEC) {
#pragma unroll
            for (int r = 0; r < R; r++)
                if ((sel >> r) & 1u) selring[(head + sink.rank[r]) & (CAP - 1)] = (int32_t)(rc.row0(r >> 1) + (r & 1));
        }
        head += wt;
        // The stage is free again. No drain (kq_pipe.cuh stage_release) here: every staged value this kernel uses has been
        // consumed by then — the predicate by the ballots above, the projection's inputs by the stores into the stash — and an
        // instruction that consumes a loaded register cannot issue before the load is back; nothing loaded from the stage is
        // touched after this point. (The drain costs this kernel 1.2 %.)
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        // ---- step B for every pending tile whose prefix has arrived
        while (kb <= k && step_b(false)) {}
    }
    while (kb < k) step_b(true);
}
#endif  // KQ_KERNEL_FILTER

}  // namespace kq
