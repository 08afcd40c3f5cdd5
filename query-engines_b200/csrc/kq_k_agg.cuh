// kq_k_agg.cuh — kernel skeleton of HashAggregateExec (Main.kt:605-660) and its accumulators
// (Main.kt:514-562), specialised per query: the generated `struct Q` evaluates the optional FilterExec
// predicate, the group-key and the aggregate-input expressions (fused ProjectionExec) for R rows per
// thread and hands them to the sink; it also carries the aggregate layout as compile-time constants.
// The skeleton runs the accumulate step of the drain loop (Main.kt:620-632).
//
// Two tiers of state:
//   * a GLOBAL open-addressing table in HBM (kq_aggtable.cuh), the source of truth;
//   * a per-CTA FRONT END in shared memory for the first FG distinct keys a CTA meets: a key directory
//     shared by the CTA plus LANE-PRIVATE count/sum accumulators (one copy per lane per warp: plain
//     read-modify-write, no atomics, no bank conflicts) and a CTA-shared MIN/MAX table that is only
//     written when a value beats the current extreme. Front ends are merged into the global table
//     once, at CTA exit. Low-cardinality GROUP BYs (BASELINE configs 3 and 5) run entirely in the
//     front end; rows whose key does not fit go straight to the global table with atomics (config 4).
//
// Accumulator semantics (oracle: MaxAccumulator etc.): nulls are skipped; a group whose inputs were
// all null yields null, except COUNT; MIN/MAX use a total order in which canonical NaN sorts above
// +inf and -0.0 below +0.0 — where the reference is order-dependent (rule R9/E8) this is the one
// deterministic choice; Float64 sums are reassociated (1e-9 relative tolerance, rule E6).
#pragma once

#include "kq_rt.cuh"
#include "kq_aggtable.cuh"

namespace kq {

// Service warp first (the warp arbiter favours high warp ids: a polling producer must not starve consumers).
#ifndef KQ_L2_PREFETCH
#define KQ_L2_PREFETCH 0          // measured: no effect here (consumers, not HBM latency, bound this kernel)
#endif
constexpr int WARPS = KQ_WARPS;              // consumer warps (the lane-private front end scales with them)
constexpr int PRODUCER_WARP = 0;
constexpr int THREADS = WARPS * 32 + 32;
constexpr int TILE = WARPS * WARP_ROWS;
constexpr int S = KQ_STAGES;
constexpr int FG = KQ_FE_GROUPS;             // front-end capacity in groups (0: no front end)
constexpr int DIR = KQ_DIR_SLOTS;            // directory slots (power of two)
constexpr int ENTRY_WORDS = (Q::NKEYS + 2) / 2 * 2;      // [0] = {state:32, key nullmask:32}, [1..NKEYS] = key words; 16-byte multiple
constexpr uint32_t DIR_EMPTY = 0, DIR_BUSY = 1, DIR_GLOBAL = 0xFFFFFFFFu;   // else gid + 2

// What the generated code fills per tile: selection, key words and aggregate inputs of the R owned rows.
struct AggSink {
    uint32_t sel;
    uint64_t key[Q::NKEYS > 0 ? Q::NKEYS : 1][R];
    uint32_t keyok[Q::NKEYS > 0 ? Q::NKEYS : 1];
    uint64_t in[Q::NIN > 0 ? Q::NIN : 1][R];
    uint32_t inok[Q::NIN > 0 ? Q::NIN : 1];
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

struct FrontEnd {
    uint64_t* dir;            // [DIR][ENTRY_WORDS]
    uint32_t* dir_count;
    uint32_t* gid2slot;       // [FG]
    uint64_t* mm;             // [FG][NMM]  CTA-shared MIN/MAX in order-mapped form
    uint32_t* cnt;            // this warp: [FG][NCNT][32]
    uint64_t* sum;            // this warp: [FG][NSUM][32]
};

// Group id of a key in the CTA directory, inserting it while there is room; -1 = the row goes to the global table.
// cheap 32-bit multiplicative hash for the CTA directory (the 64-bit hash of the global table is only
// computed for rows that actually go there)
__device__ __forceinline__ uint32_t dir_hash(const uint64_t (&kw)[MAX_KEYS], uint32_t nullmask) {
    uint32_t h = nullmask * 0x9E3779B1u;
#pragma unroll
    for (int k = 0; k < Q::NKEYS; k++) h = (h ^ (uint32_t)kw[k] ^ ((uint32_t)(kw[k] >> 32) * 0x85EBCA6Bu)) * 0x9E3779B1u;
    // murmur3 finaliser: small integers and short strings must not share low bits
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h & (DIR - 1);
}

__device__ __forceinline__ int dir_lookup(const FrontEnd& fe, uint32_t slot, const uint64_t (&kw)[MAX_KEYS], uint32_t nullmask) {
#pragma unroll 1
    for (int probe = 0; probe < 8; probe++) {
        uint64_t* e = fe.dir + (size_t)slot * ENTRY_WORDS;
        const uint4 q = lds_v4(e);
        uint32_t st = q.x;
        if (st == DIR_EMPTY) {
            const uint32_t old = atomicCAS(reinterpret_cast<uint32_t*>(e), DIR_EMPTY, DIR_BUSY);
            if (old == DIR_EMPTY) {
                const uint32_t gid = atomicAdd(fe.dir_count, 1u);
                reinterpret_cast<uint32_t*>(e)[1] = nullmask;
#pragma unroll
                for (int k = 0; k < Q::NKEYS; k++) e[1 + k] = kw[k];
                const bool fits = gid < (uint32_t)FG;
                if (fits) fe.gid2slot[gid] = slot;
                __threadfence_block();
                *reinterpret_cast<volatile uint32_t*>(e) = fits ? gid + 2 : DIR_GLOBAL;
                return fits ? (int)gid : -1;
            }
            st = old;
            if (st == DIR_BUSY) return -1;
            // published by somebody else in the meantime: fall through and compare (re-read below)
        }
        if (st == DIR_BUSY) return -1;             // being published: this row takes the global path
        uint32_t e_nm = q.y;
        uint64_t e_k0 = (uint64_t)q.z | ((uint64_t)q.w << 32);
        if (st != q.x) {                           // the entry was published between our load and our CAS: read it again
            e_nm = reinterpret_cast<volatile uint32_t*>(e)[1];
            e_k0 = *reinterpret_cast<volatile uint64_t*>(e + 1);
        }
        bool eq = e_nm == nullmask;
        if constexpr (Q::NKEYS >= 1) eq &= e_k0 == kw[0];
#pragma unroll
        for (int k = 1; k < Q::NKEYS; k++) eq &= *reinterpret_cast<volatile uint64_t*>(e + 1 + k) == kw[k];
        if (eq) return st == DIR_GLOBAL ? -1 : (int)(st - 2);
        slot = (slot + 1) & (DIR - 1);
    }
    return -1;
}

// Front-end accumulate of input I for one row (lane-private slots: plain read-modify-write).
template <int I>
__device__ __forceinline__ void fe_accumulate(const FrontEnd& fe, int gid, int lane, const AggSink& sink, int r) {
    if constexpr (I < Q::NIN) {
        constexpr int FL = Q::IN_FLAGS[I];
        const bool valid = (sink.inok[I] >> r) & 1u;
        if (Q::IN_CNT[I] > 0 && valid) fe.cnt[((gid * Q::NCNT + Q::IN_CNT[I]) << 5) + lane] += 1u;     // slot 0 (all rows) is counted by the caller
        if constexpr ((FL & (F_SUM | F_MIN | F_MAX)) != 0) {
            if (valid) {
                const uint64_t v = sink.in[I][r];
                if constexpr ((FL & F_SUM) != 0) {
                    uint64_t* p = fe.sum + ((gid * Q::NSUM + Q::FE_SUM[I]) << 5) + lane;
                    if constexpr ((FL & F_INT) != 0) *p += v;
                    else *p = as_u64(__dadd_rn(as_f64(*p), as_f64(v)));
                }
                if constexpr ((FL & (F_MIN | F_MAX)) != 0) {
                    constexpr bool is_int = (FL & F_INT) != 0;
                    const uint64_t m = order_map(is_int ? v : canon_nan(v), is_int);
                    if constexpr ((FL & F_MIN) != 0) {
                        uint64_t* p = fe.mm + gid * Q::NMM + Q::FE_MIN[I];
                        if (m < *reinterpret_cast<volatile uint64_t*>(p)) atomicMin(reinterpret_cast<unsigned long long*>(p), (unsigned long long)m);
                    }
                    if constexpr ((FL & F_MAX) != 0) {
                        uint64_t* p = fe.mm + gid * Q::NMM + Q::FE_MAX[I];
                        if (m > *reinterpret_cast<volatile uint64_t*>(p)) atomicMax(reinterpret_cast<unsigned long long*>(p), (unsigned long long)m);
                    }
                }
            }
        }
        fe_accumulate<I + 1>(fe, gid, lane, sink, r);
    }
}

// ---- explicit shared-space accessors (32-bit addresses: no generic-pointer arithmetic on the hot path) -------------
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint64_t lds_u64(uint32_t a) { uint64_t v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint4 lds_u128(uint32_t a) { uint4 r; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a) : "memory"); return r; }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u64(uint32_t a, uint64_t v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
// CTA-shared MIN/MAX slot: a predicated reduction, no branch (the slot is only written when the value beats it)
__device__ __forceinline__ void smem_min_u64(uint32_t a, uint64_t m) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .u64 c;\n\tld.volatile.shared.u64 c, [%0];\n\tsetp.lt.u64 p, %1, c;\n\t@p red.shared.min.u64 [%0], %1;\n\t}"
                 ::"r"(a), "l"(m) : "memory");
}
__device__ __forceinline__ void smem_max_u64(uint32_t a, uint64_t m) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .u64 c;\n\tld.volatile.shared.u64 c, [%0];\n\tsetp.gt.u64 p, %1, c;\n\t@p red.shared.max.u64 [%0], %1;\n\t}"
                 ::"r"(a), "l"(m) : "memory");
}

// Branch-free front-end accumulate of input I for one row. g = the row's group, or FG (a trash group
// nobody reads) when the row is filtered out / takes the slow path; an invalid (null) input goes to the
// trash group as well, so every read-modify-write is unconditional.
template <int I>
__device__ __forceinline__ void fe_accumulate_fast(uint32_t a_cnt, uint32_t a_sum, uint32_t a_mm, int g, int lane, const AggSink& sink, int r) {
    if constexpr (I < Q::NIN) {
        constexpr int FL = Q::IN_FLAGS[I];
        const bool valid = (sink.inok[I] >> r) & 1u;
        const int gi = (Q::IN_CNT[I] > 0 && !valid) ? FG : g;           // statically non-null inputs: always valid
        if constexpr (Q::IN_CNT[I] > 0) {
            const uint32_t a = a_cnt + ((((uint32_t)gi * Q::NCNT + Q::IN_CNT[I]) << 5) + lane) * 4u;
            sts_u32(a, lds_u32(a) + 1u);
        }
        if constexpr ((FL & (F_SUM | F_MIN | F_MAX)) != 0) {
            const uint64_t v = sink.in[I][r];
            if constexpr ((FL & F_SUM) != 0) {
                const uint32_t a = a_sum + ((((uint32_t)gi * Q::NSUM + Q::FE_SUM[I]) << 5) + lane) * 8u;
                if constexpr ((FL & F_INT) != 0) sts_u64(a, lds_u64(a) + v);
                else sts_u64(a, as_u64(__dadd_rn(as_f64(lds_u64(a)), as_f64(v))));
            }
            if constexpr ((FL & (F_MIN | F_MAX)) != 0) {
                constexpr bool is_int = (FL & F_INT) != 0;
                const uint64_t m = order_map(is_int ? v : canon_nan(v), is_int);
                if constexpr ((FL & F_MIN) != 0) smem_min_u64(a_mm + ((uint32_t)gi * Q::NMM + Q::FE_MIN[I]) * 8u, m);
                if constexpr ((FL & F_MAX) != 0) smem_max_u64(a_mm + ((uint32_t)gi * Q::NMM + Q::FE_MAX[I]) * 8u, m);
            }
        }
        fe_accumulate_fast<I + 1>(a_cnt, a_sum, a_mm, g, lane, sink, r);
    }
}

// Accumulate one row straight into its record of the global table.
template <int I>
__device__ __forceinline__ void global_accumulate_all(const AggArgs& A, uint64_t* rec, const AggSink& sink, int r) {
    if constexpr (I < Q::NIN) {
        if ((sink.inok[I] >> r) & 1u) global_accumulate(rec, A.in[I], sink.in[I][r]);
        global_accumulate_all<I + 1>(A, rec, sink, r);
    }
}

// Merge the lane-private slots of front-end group g (this warp's copy) into its global record.
template <int I>
__device__ __forceinline__ void fe_merge_input(const FrontEnd& fe, uint64_t* rec, int g, int lane, const unsigned long long (&c)[Q::NCNT]) {
    if constexpr (I < Q::NIN) {
        constexpr int FL = Q::IN_FLAGS[I];
        const unsigned long long n = c[Q::IN_CNT[I]];
        if (n != 0) {                                       // this warp saw no non-null value of input I in group g otherwise
            if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(rec + Q::REC_NN[I]), n);
            if constexpr ((FL & F_SUM) != 0) {
                uint64_t x = fe.sum[((g * Q::NSUM + Q::FE_SUM[I]) << 5) + lane];
                if constexpr ((FL & F_INT) != 0) {
#pragma unroll
                    for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
                    if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(rec + Q::REC_SUM[I]), (unsigned long long)x);
                } else {
                    double f = as_f64(x);
#pragma unroll
                    for (int o = 16; o; o >>= 1) f = __dadd_rn(f, __shfl_xor_sync(0xffffffffu, f, o));
                    if (lane == 0) atomicAdd(reinterpret_cast<double*>(rec + Q::REC_SUM[I]), f);
                }
            }
        }
        fe_merge_input<I + 1>(fe, rec, g, lane, c);
    }
}

extern "C" __global__ void __launch_bounds__(THREADS, 1) kq_hash_aggregate(const __grid_constant__ AggArgs A) {
    // shared memory: [S stages][directory][gslot][gid2slot][mm][per-warp sums][per-warp counts]
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[S], empty[S], full2[S];     // full2: second-phase copies (Utf8 string bytes)
    __shared__ long long tile_of[S];
    __shared__ long long bbase[S][MAX_COLS];
    __shared__ uint32_t s_dir_count;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warp = wid - 1;                 // consumer warp index
    constexpr int NCNT = Q::NCNT, NSUM = Q::NSUM, NMM = Q::NMM;
    unsigned char* p0 = smem + (size_t)S * A.sp.stage_bytes;
    const size_t fe_begin = (size_t)(p0 - smem);
    FrontEnd fe;
    fe.dir = reinterpret_cast<uint64_t*>(p0);            p0 += (size_t)DIR * ENTRY_WORDS * 8;
    uint64_t* gslot = reinterpret_cast<uint64_t*>(p0);   p0 += (size_t)(FG > 0 ? FG : 1) * 8;
    constexpr int FG1 = FG + 1;               // + one trash group: the branch-free path parks filtered-out / slow-path rows there
    fe.mm = reinterpret_cast<uint64_t*>(p0);             p0 += (size_t)FG1 * NMM * 8;
    uint64_t* sum0 = reinterpret_cast<uint64_t*>(p0);    p0 += (size_t)WARPS * FG1 * NSUM * 32 * 8;
    uint32_t* cnt0 = reinterpret_cast<uint32_t*>(p0);    p0 += (size_t)WARPS * FG1 * NCNT * 32 * 4;
    fe.gid2slot = reinterpret_cast<uint32_t*>(p0);       p0 += (size_t)(FG > 0 ? FG : 1) * 4;
    const size_t fe_end = (size_t)(p0 - smem);
    fe.dir_count = &s_dir_count;
    fe.sum = sum0 + (size_t)(warp < 0 ? 0 : warp) * FG1 * NSUM * 32;
    fe.cnt = cnt0 + (size_t)(warp < 0 ? 0 : warp) * FG1 * NCNT * 32;

    for (size_t i = fe_begin + threadIdx.x * 4; i < fe_end; i += THREADS * 4) *reinterpret_cast<uint32_t*>(smem + i) = 0;
    if (threadIdx.x == 0) {
        s_dir_count = 0;
        for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], WARPS); mbar_init(&full2[s], 1); }
        mbar_fence_init();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < FG1 * NMM; i += THREADS) fe.mm[i] = ((Q::MM_ISMIN >> (i % (NMM > 0 ? NMM : 1))) & 1u) ? ~0ULL : 0ULL;
    __syncthreads();

    if (wid == PRODUCER_WARP) {
        if (lane == 0) {
            // Two cursors over the CTA's tiles: kp = next tile whose fixed-size buffers (phase 1) are issued,
            // kb = next tile whose Utf8 string bytes (phase 2) are issued — the byte range of a tile is only
            // known once its offsets have landed in the stage. Neither waits for the other: the loop polls.
            // The next ticket is always requested one step early so that the L2 round trip of the atomic
            // overlaps the wait for a free stage.
            int kp = 0, kb = 0;
            bool done = false;
            auto take = [&]() -> long long {
                // stop taking tiles once the global table is half full: every ticket taken is processed,
                // so the rows consumed so far are always a prefix of the batch (the host grows and resumes)
                const unsigned long long g = *reinterpret_cast<volatile unsigned long long*>(A.ngroups);
                if (g > A.stop_threshold) return -1;
                const long long t = (long long)atomicAdd(A.ticket, 1u) + A.tile_begin;
                return t < A.ntiles ? t : -1;
            };
            long long next = take();
            while (!done || kb < kp) {
                bool did = false;
                if (KQ_STAGE_BYTES && kb < kp) {
                    const int sb = kb % S;
                    if (mbar_test(&full[sb], (kb / S) & 1)) {
                        stage_issue_bytes(A.sp, smem + (size_t)sb * A.sp.stage_bytes, &full2[sb], tile_of[sb], TILE, A.n, bbase[sb]);
                        kb++; did = true;
                    }
                }
                if (!done) {
                    const int s = kp % S;
                    if (mbar_test(&empty[s], ((kp / S) & 1) ^ 1)) {
                        const long long tile = next;
                        if (tile < 0) { tile_of[s] = -1; mbar_arrive(&full[s]); done = true; }
                        else {
                            tile_of[s] = tile;
                            stage_issue(A.sp, smem + (size_t)s * A.sp.stage_bytes, &full[s], tile, TILE, A.n);
                            next = take();
                            if (KQ_L2_PREFETCH > 0 && next >= 0) stage_prefetch_l2(A.sp, next, TILE, A.n);     // staged one step from now
                            kp++;
                        }
                        did = true;
                    }
                }
                if (!KQ_STAGE_BYTES) kb = kp;
                if (!did) __nanosleep(20);
            }
        }
    } else {
        AggSink sink;
        bool bypass = FG == 0;
        const uint32_t a_dir = smem_u32(fe.dir), a_cnt = smem_u32(fe.cnt), a_sum = smem_u32(fe.sum), a_mm = smem_u32(fe.mm);
        for (int k = 0;; k++) {
            const int s = k % S;
            mbar_wait(&full[s], (k / S) & 1);
            const long long tile = tile_of[s];
            if (tile < 0) break;
            if (KQ_STAGE_BYTES) mbar_wait(&full2[s], (k / S) & 1);
            RowCtx rc;
            rowctx_init(rc, warp, tile, TILE, A.n, A.err, smem + (size_t)s * A.sp.stage_bytes);
            rc.bbase = bbase[s];
            sink.sel = rc.inr;
            Q::eval(A.q, rc, sink);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);       // everything needed is in registers now

            // canonical key words + null masks of the R owned rows
            uint32_t nm[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                uint32_t nullmask = 0;
#pragma unroll
                for (int k2 = 0; k2 < Q::NKEYS; k2++) {
                    uint64_t w = 0;
                    if ((sink.keyok[k2] >> r) & 1u) w = ((Q::KEY_F64_MASK >> k2) & 1u) ? canon_nan(sink.key[k2][r]) : sink.key[k2][r];
                    else nullmask |= 1u << k2;
                    sink.key[k2][r] = w;
                }
                nm[r] = nullmask;
            }
            uint32_t slow = sink.sel;             // rows that still need the general path
            int fe_hits = 0;
            if (!bypass) {
                // pass 1 (branch-free): one directory probe per row; a first-probe hit yields the group id
                int gid[R];
                slow = 0;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    uint64_t kw[MAX_KEYS];
#pragma unroll
                    for (int k2 = 0; k2 < MAX_KEYS; k2++) kw[k2] = k2 < Q::NKEYS ? sink.key[k2][r] : 0;
                    // the key's home slot and its neighbour (linear probing rarely displaces a key further)
                    const uint32_t s0 = dir_hash(kw, nm[r]);
                    const uint32_t e0 = a_dir + s0 * (ENTRY_WORDS * 8u), e1 = a_dir + ((s0 + 1) & (DIR - 1)) * (ENTRY_WORDS * 8u);
                    const uint4 q0 = lds_u128(e0), q1 = lds_u128(e1);
                    bool h0 = q0.x >= 2u && q0.x != DIR_GLOBAL && q0.y == nm[r], h1 = q1.x >= 2u && q1.x != DIR_GLOBAL && q1.y == nm[r];
                    if constexpr (Q::NKEYS >= 1) {
                        h0 &= ((uint64_t)q0.z | ((uint64_t)q0.w << 32)) == kw[0];
                        h1 &= ((uint64_t)q1.z | ((uint64_t)q1.w << 32)) == kw[0];
                    }
#pragma unroll
                    for (int k2 = 1; k2 < Q::NKEYS; k2++) { h0 &= lds_u64(e0 + 8u + 8u * k2) == kw[k2]; h1 &= lds_u64(e1 + 8u + 8u * k2) == kw[k2]; }
                    const bool hit = h0 | h1;
                    const uint4 q = h0 ? q0 : q1;
                    const bool on = (sink.sel >> r) & 1u;
                    gid[r] = (on && hit) ? (int)(q.x - 2u) : FG;
                    slow |= (uint32_t)(on && !hit) << r;
                }
                // pass 2 (branch-free): unconditional read-modify-write of the lane-private slots
#pragma unroll
                for (int r = 0; r < R; r++) {
                    if constexpr (Q::CNT0_USED) { const uint32_t a = a_cnt + ((((uint32_t)gid[r] * NCNT) << 5) + lane) * 4u; sts_u32(a, lds_u32(a) + 1u); }
                    fe_accumulate_fast<0>(a_cnt, a_sum, a_mm, gid[r], lane, sink, r);
                }
                fe_hits = __popc(sink.sel & ~slow);
            }
            // general path for the rest: directory probing with insertion, else the global table
            int rows = __popc(sink.sel);
            if (slow) {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    if (!((slow >> r) & 1u)) continue;
                    uint64_t kw[MAX_KEYS];
#pragma unroll
                    for (int k2 = 0; k2 < MAX_KEYS; k2++) kw[k2] = k2 < Q::NKEYS ? sink.key[k2][r] : 0;
                    int g = -1;
                    if (!bypass) g = dir_lookup(fe, dir_hash(kw, nm[r]), kw, nm[r]);
                    if (g >= 0) {
                        fe_hits++;
                        if constexpr (Q::CNT0_USED) fe.cnt[((g * NCNT) << 5) + lane] += 1u;
                        fe_accumulate<0>(fe, g, lane, sink, r);
                    } else {
                        uint64_t* rec = table_find_or_insert(A, hash_key(kw, nm[r], Q::NKEYS), kw, nm[r]);
                        global_accumulate_all<0>(A, rec, sink, r);
                    }
                }
            }
            // once the directory is full and this warp mostly misses it, stop probing it (high cardinality)
            if (!bypass && *reinterpret_cast<volatile uint32_t*>(&s_dir_count) >= (uint32_t)FG) {
                int hits = fe_hits, tot = rows;
#pragma unroll
                for (int o = 16; o; o >>= 1) { hits += __shfl_xor_sync(0xffffffffu, hits, o); tot += __shfl_xor_sync(0xffffffffu, tot, o); }
                if (tot >= 64 && hits * 8 < tot) bypass = true;
            }
        }
    }

    // ---- merge the front end into the global table ---------------------------------------------------
    __syncthreads();
    const int G = min((int)s_dir_count, FG);
    for (int g = threadIdx.x; g < G; g += THREADS) {
        const uint64_t* e = fe.dir + (size_t)fe.gid2slot[g] * ENTRY_WORDS;
        uint64_t kw[MAX_KEYS];
#pragma unroll
        for (int k = 0; k < MAX_KEYS; k++) kw[k] = k < Q::NKEYS ? e[1 + k] : 0;
        const uint32_t nullmask = (uint32_t)(e[0] >> 32);
        uint64_t* rec = table_find_or_insert(A, hash_key(kw, nullmask, Q::NKEYS), kw, nullmask);
        gslot[g] = (uint64_t)(rec - A.table);
    }
    __syncthreads();
    if (warp >= 0) {
        for (int g = 0; g < G; g++) {
            uint64_t* rec = A.table + gslot[g];
            unsigned long long c[NCNT];
#pragma unroll
            for (int j = 0; j < NCNT; j++) {
                unsigned long long x = fe.cnt[((g * NCNT + j) << 5) + lane];
#pragma unroll
                for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
                c[j] = x;
            }
            fe_merge_input<0>(fe, rec, g, lane, c);
        }
    }
    for (int t = threadIdx.x; t < G * NMM; t += THREADS) {
        const int g = t / (NMM > 0 ? NMM : 1), m = t % (NMM > 0 ? NMM : 1);
        const uint64_t v = fe.mm[t];
        uint64_t* p = A.table + gslot[g] + Q::MM_WORD[m];
        if ((Q::MM_ISMIN >> m) & 1u) { if (v != ~0ULL) atomicMin(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v); }
        else if (v != 0ULL) atomicMax(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
    }
}

}  // namespace kq
