// kq_k_agg.cuh — kernel skeleton of HashAggregateExec (Main.kt:605-660) and its accumulators
// (Main.kt:514-562), specialised per query: the generated `struct Q` evaluates the optional FilterExec
// predicate, the group-key and the aggregate-input expressions (fused ProjectionExec) for R rows per
// thread and hands them to the sink; the skeleton runs the accumulate step of the drain loop
// (Main.kt:620-632).
//
// Two tiers of state:
//   * a GLOBAL open-addressing table in HBM (kq_aggtable.cuh), the source of truth;
//   * a per-CTA FRONT END in shared memory for the first `fe_groups` distinct keys a CTA meets: a key
//     directory shared by the CTA plus LANE-PRIVATE count/sum accumulators (one copy per lane per
//     warp: no atomics, no bank conflicts) and a CTA-shared MIN/MAX table that is only touched when a
//     value beats the current extreme. Front ends are merged into the global table once, at CTA exit.
//     Low-cardinality GROUP BYs (BASELINE configs 3 and 5) run entirely in the front end; rows whose
//     key does not fit go straight to the global table with atomics (config 4).
//
// Accumulator semantics (oracle: MaxAccumulator etc.): nulls are skipped; a group whose inputs were
// all null yields null, except COUNT; MIN/MAX use a total order in which canonical NaN sorts above
// +inf and -0.0 below +0.0 — where the reference is order-dependent (rule R9/E8) this is the one
// deterministic choice; Float64 sums are reassociated (1e-9 relative tolerance, rule E6).
#pragma once

#include "kq_rt.cuh"
#include "kq_aggtable.cuh"

namespace kq {

// KQ_WARPS consumer warps (the lane-private front end scales with the warp count) + 1 service warp (TMA producer)
constexpr int WARPS = KQ_WARPS;
constexpr int BLOCK = WARPS * 32;
constexpr int TILE = WARPS * WARP_ROWS;
constexpr int SERVICE_WARP = WARPS;
constexpr int THREADS = BLOCK + 32;
constexpr uint32_t DIR_EMPTY = 0, DIR_BUSY = 1, DIR_GLOBAL = 0xFFFFFFFFu;   // FULL = gid + 2

// What the generated code fills per tile: selection, key words and aggregate inputs of the R owned rows.
struct AggSink {
    uint32_t sel;
    uint64_t key[MAX_KEYS][R];
    uint32_t keyok[MAX_KEYS];
    uint64_t in[MAX_INPUTS][R];
    uint32_t inok[MAX_INPUTS];
    template <int K>
    __device__ __forceinline__ void set_key(const uint64_t (&v)[R], uint32_t ok) {
#pragma unroll
        for (int r = 0; r < R; r++) key[K][r] = v[r];
        keyok[K] = ok;
    }
    template <int I>
    __device__ __forceinline__ void set_in(const uint64_t (&v)[R], uint32_t ok) {
#pragma unroll
        for (int r = 0; r < R; r++) in[I][r] = v[r];
        inok[I] = ok;
    }
};

// Look the key up in the CTA directory; returns the front-end group id or -1 (row goes global).
__device__ __forceinline__ int dir_lookup(const AggArgs& A, uint64_t* dirkeys, uint32_t* dirstate, uint32_t* gid2slot,
                                          uint32_t* dir_count, uint64_t h, const uint64_t (&kw)[MAX_KEYS], uint32_t nullmask) {
    const int KW = A.nkeys + 1;
    uint32_t slot = (uint32_t)(h >> 40) & (DIR_SLOTS - 1);
#pragma unroll 1
    for (int probe = 0; probe < 8; probe++) {
        uint32_t st = *reinterpret_cast<volatile uint32_t*>(dirstate + slot);
        if (st == DIR_EMPTY) {
            uint32_t old = atomicCAS(dirstate + slot, DIR_EMPTY, DIR_BUSY);
            if (old == DIR_EMPTY) {
                uint32_t gid = atomicAdd(dir_count, 1u);
                uint64_t* dk = dirkeys + slot * KW;
                dk[0] = nullmask;
#pragma unroll
                for (int k = 0; k < MAX_KEYS; k++) if (k < A.nkeys) dk[1 + k] = kw[k];
                bool fits = gid < (uint32_t)A.fe_groups;
                if (fits) gid2slot[gid] = slot;
                __threadfence_block();
                *reinterpret_cast<volatile uint32_t*>(dirstate + slot) = fits ? gid + 2 : DIR_GLOBAL;
                return fits ? (int)gid : -1;
            }
            st = old;
        }
        if (st == DIR_BUSY) return -1;             // being published: this row takes the global path
        const uint64_t* dk = dirkeys + slot * KW;
        bool eq = dk[0] == (uint64_t)nullmask;
#pragma unroll
        for (int k = 0; k < MAX_KEYS; k++) if (k < A.nkeys) eq &= dk[1 + k] == kw[k];
        if (eq) return st == DIR_GLOBAL ? -1 : (int)(st - 2);
        slot = (slot + 1) & (DIR_SLOTS - 1);
    }
    return -1;
}

extern "C" __global__ void __launch_bounds__(THREADS, 1) kq_hash_aggregate(const __grid_constant__ AggArgs A) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[MAX_STAGES], empty[MAX_STAGES];
    __shared__ long long tile_of[MAX_STAGES];
    __shared__ uint32_t s_dir_count;
    uint64_t* dirkeys = reinterpret_cast<uint64_t*>(smem + A.off_dirkeys);
    uint32_t* dirstate = reinterpret_cast<uint32_t*>(smem + A.off_dirstate);
    uint32_t* gid2slot = reinterpret_cast<uint32_t*>(smem + A.off_gid2slot);
    uint64_t* gslot = reinterpret_cast<uint64_t*>(smem + A.off_gslot);
    uint64_t* mm = reinterpret_cast<uint64_t*>(smem + A.off_mm);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NI = A.ninputs, NS = A.fe_nsum, NM = A.fe_nmm, FG = A.fe_groups;
    const int S = A.sp.nstages;
    // lane-private accumulators of this warp: cnt[(gid*NI + i)*32 + lane], sum[(gid*NS + s)*32 + lane]
    uint32_t* cnt = reinterpret_cast<uint32_t*>(smem + A.off_cnt) + (size_t)(warp % WARPS) * FG * NI * 32;
    uint64_t* sum = reinterpret_cast<uint64_t*>(smem + A.off_sum) + (size_t)(warp % WARPS) * FG * NS * 32;

    for (int i = A.off_fe + threadIdx.x * 4; i < A.smem_bytes; i += THREADS * 4) *reinterpret_cast<uint32_t*>(smem + i) = 0;
    if (threadIdx.x == 0) {
        s_dir_count = 0;
        for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], WARPS); }
        mbar_fence_init();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < FG * NM; i += THREADS) mm[i] = ((A.fe_mm_ismin >> (i % NM)) & 1u) ? ~0ULL : 0ULL;
    __syncthreads();

    if (warp == SERVICE_WARP) {
        if (lane == 0) {
            for (int k = 0;; k++) {
                const int s = k % S;
                mbar_wait(&empty[s], ((k / S) & 1) ^ 1);
                // stop taking tiles once the global table is half full: every ticket taken is processed,
                // so the rows consumed so far are always a prefix of the batch (the host grows and resumes)
                const unsigned long long g = *reinterpret_cast<volatile unsigned long long*>(A.ngroups);
                long long tile = -1;
                if (g <= A.stop_threshold) tile = (long long)atomicAdd(A.ticket, 1u) + A.tile_begin;
                if (tile < 0 || tile >= A.ntiles) { tile_of[s] = -1; mbar_arrive(&full[s]); break; }
                tile_of[s] = tile;
                stage_issue(A.sp, smem + (size_t)s * A.sp.stage_bytes, &full[s], tile, TILE, A.n);
            }
        }
    } else {
    AggSink sink;
    bool bypass = FG == 0;
    for (int k = 0;; k++) {
        const int s = k % S;
        mbar_wait(&full[s], (k / S) & 1);
        const long long tile = tile_of[s];
        if (tile < 0) break;
        RowCtx rc;
        rowctx_init(rc, warp, tile, TILE, A.n, A.err, smem + (size_t)s * A.sp.stage_bytes);
        sink.sel = rc.inr;
        Q::eval(A.q, rc, sink);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);       // everything needed is in registers now

        int fe_hits = 0, rows = 0;
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (!((sink.sel >> r) & 1u)) continue;
            uint64_t kw[MAX_KEYS];
            uint32_t nullmask = 0;
#pragma unroll
            for (int k2 = 0; k2 < MAX_KEYS; k2++) {
                kw[k2] = 0;
                if (k2 < A.nkeys) {
                    if ((sink.keyok[k2] >> r) & 1u) kw[k2] = ((A.key_f64_mask >> k2) & 1u) ? canon_nan(sink.key[k2][r]) : sink.key[k2][r];
                    else nullmask |= 1u << k2;
                }
            }
            const uint64_t h = hash_key(kw, nullmask, A.nkeys);
            int gid = -1;
            if (!bypass) gid = dir_lookup(A, dirkeys, dirstate, gid2slot, &s_dir_count, h, kw, nullmask);
            rows++;
            if (gid >= 0) {
                fe_hits++;
#pragma unroll
                for (int i = 0; i < MAX_INPUTS; i++) {
                    if (i < NI && ((sink.inok[i] >> r) & 1u)) {
                        const AggInput d = A.in[i];
                        const uint64_t v = sink.in[i][r];
                        cnt[(gid * NI + i) * 32 + lane] += 1u;
                        if (d.flags & F_SUM) {
                            uint64_t* p = sum + (gid * NS + d.fe_sum) * 32 + lane;
                            if (d.flags & F_INT) *p += v;
                            else *p = (uint64_t)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)*p), __longlong_as_double((long long)v)));
                        }
                        if (d.flags & (F_MIN | F_MAX)) {
                            const bool is_int = d.flags & F_INT;
                            uint64_t m = order_map(is_int ? v : canon_nan(v), is_int);
                            if (d.flags & F_MIN) { uint64_t* p = mm + gid * NM + d.fe_min; if (m < *reinterpret_cast<volatile uint64_t*>(p)) atomicMin(reinterpret_cast<unsigned long long*>(p), (unsigned long long)m); }
                            if (d.flags & F_MAX) { uint64_t* p = mm + gid * NM + d.fe_max; if (m > *reinterpret_cast<volatile uint64_t*>(p)) atomicMax(reinterpret_cast<unsigned long long*>(p), (unsigned long long)m); }
                        }
                    }
                }
            } else {
                uint64_t* rec = table_find_or_insert(A, h, kw, nullmask);
#pragma unroll
                for (int i = 0; i < MAX_INPUTS; i++)
                    if (i < NI && ((sink.inok[i] >> r) & 1u)) global_accumulate(rec, A.in[i], sink.in[i][r]);
            }
        }
        // once the directory is full and this warp mostly misses it, stop probing it (high cardinality)
        if (!bypass) {
            int hits = fe_hits, tot = rows;
#pragma unroll
            for (int o = 16; o; o >>= 1) { hits += __shfl_xor_sync(0xffffffffu, hits, o); tot += __shfl_xor_sync(0xffffffffu, tot, o); }
            if (tot >= 64 && hits * 8 < tot && *reinterpret_cast<volatile uint32_t*>(&s_dir_count) >= (uint32_t)FG) bypass = true;
        }
    }
    }

    // ---- merge the front end into the global table ---------------------------------------------------
    __syncthreads();
    const int G = min((int)s_dir_count, FG);
    for (int g = threadIdx.x; g < G; g += THREADS) {
        const uint64_t* dk = dirkeys + gid2slot[g] * (A.nkeys + 1);
        uint64_t kw[MAX_KEYS];
#pragma unroll
        for (int k = 0; k < MAX_KEYS; k++) kw[k] = k < A.nkeys ? dk[1 + k] : 0;
        uint32_t nullmask = (uint32_t)dk[0];
        uint64_t* rec = table_find_or_insert(A, hash_key(kw, nullmask, A.nkeys), kw, nullmask);
        gslot[g] = (uint64_t)(rec - A.table);
    }
    __syncthreads();
    for (int g = 0; g < G && warp < WARPS; g++) {
        uint64_t* rec = A.table + gslot[g];
        for (int i = 0; i < NI; i++) {
            unsigned long long c = cnt[(g * NI + i) * 32 + lane];
#pragma unroll
            for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (c == 0) continue;                                   // this warp saw no non-null value of input i in group g
            const AggInput d = A.in[i];
            if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(rec + d.rec_nn), c);
            if (d.flags & F_SUM) {
                uint64_t x = sum[(g * NS + d.fe_sum) * 32 + lane];
                if (d.flags & F_INT) {
#pragma unroll
                    for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
                    if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(rec + d.rec_sum), (unsigned long long)x);
                } else {
                    double f = __longlong_as_double((long long)x);
#pragma unroll
                    for (int o = 16; o; o >>= 1) f = __dadd_rn(f, __shfl_xor_sync(0xffffffffu, f, o));
                    if (lane == 0) atomicAdd(reinterpret_cast<double*>(rec + d.rec_sum), f);
                }
            }
        }
    }
    for (int t = threadIdx.x; t < G * NM; t += THREADS) {
        const int g = t / NM, m = t % NM;
        const uint64_t v = mm[t];
        uint64_t* p = A.table + gslot[g] + A.fe_mm_word[m];
        if ((A.fe_mm_ismin >> m) & 1u) { if (v != ~0ULL) atomicMin(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v); }
        else if (v != 0ULL) atomicMax(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
    }
}


}  // namespace kq
