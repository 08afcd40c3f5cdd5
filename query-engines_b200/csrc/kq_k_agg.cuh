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
// all null yields null, except COUNT; MIN/MAX compare like the reference's `value > this.value`
// (Main.kt:552): a NaN never replaces a held value, and a group whose non-null values were all NaN
// yields NaN (k_finalize); among equal-comparing zeros -0.0 orders below +0.0 (the reference keeps
// the first seen, an order dependence); Float64 sums are reassociated (1e-9 relative, rule E6).
#pragma once

#include "kq_rt.cuh"
#include "kq_aggtable.cuh"

namespace kq {

// Service warp first (the warp arbiter favours high warp ids: a polling producer must not starve consumers).
#ifndef KQ_L2_PREFETCH
#define KQ_L2_PREFETCH 0          // measured: no effect one step ahead; several grid rounds ahead: 20 % slower
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
    unsigned long long* mm_bound;   // [NMM] {insert epoch : 32 | bound on the high word over all front-end groups : 32}
    uint32_t* cnt;            // this warp: [FG][NCNT][32]
    uint64_t* sum;            // this warp: [FG][NSUM][32]
};

// Home slot of a key in the CTA directory. Multiplicative (Fibonacci) hashing: the HIGH bits of the product are
// the well-mixed ones, so small integers and short strings spread without a finaliser (the 64-bit hash of the
// global table is only computed for rows that actually go there). The home is an EVEN slot — a two-slot bucket,
// probed linearly from there — so the branch-free first probe reads slots home and home + 1 without wrapping.
__host__ __device__ constexpr int ilog2c(int x) { return x <= 1 ? 0 : 1 + ilog2c(x >> 1); }
__device__ __forceinline__ uint32_t dir_hash(const uint64_t (&kw)[MAX_KEYS], uint32_t nullmask) {
    uint32_t h = Q::KEYS_NULLABLE ? nullmask * 0x9E3779B1u : 0u;
#pragma unroll
    for (int k = 0; k < Q::NKEYS; k++) h = (h ^ (uint32_t)kw[k] ^ ((uint32_t)(kw[k] >> 32) * 0x85EBCA6Bu)) * 0x9E3779B1u;
    return (h >> (33 - ilog2c(DIR))) << 1;
}

// Group id of a key in the CTA directory, inserting it while there is room; -1 = the row goes to the global table.
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
                if (fits) {
                    fe.gid2slot[gid] = slot;
                    // a new group's extremes are the identities: no bound holds any more. The epoch makes a concurrent
                    // recomputation (mm_bound_refresh), which cannot have seen this group, lose its compare-and-swap.
#pragma unroll
                    for (int m = 0; m < Q::NMM; m++)
                        atomicExch(fe.mm_bound + m, ((unsigned long long)(gid + 1u) << 32) | (((Q::MM_ISMIN >> m) & 1u) ? 0xFFFFFFFFull : 0ull));
                }
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
                    const uint64_t m = order_map(v, is_int);
                    const bool cmp = is_int || as_f64(v) == as_f64(v);          // a NaN never replaces a held value (Main.kt:552)
                    if constexpr ((FL & F_MIN) != 0) {
                        uint64_t* p = fe.mm + gid * Q::NMM + Q::FE_MIN[I];
                        if (cmp && m < *reinterpret_cast<volatile uint64_t*>(p)) atomicMin(reinterpret_cast<unsigned long long*>(p), (unsigned long long)m);
                    }
                    if constexpr ((FL & F_MAX) != 0) {
                        uint64_t* p = fe.mm + gid * Q::NMM + Q::FE_MAX[I];
                        if (cmp && m > *reinterpret_cast<volatile uint64_t*>(p)) atomicMax(reinterpret_cast<unsigned long long*>(p), (unsigned long long)m);
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

// Recompute the all-groups bounds from the CTA-shared extremes (one warp, every few tiles). Extremes only tighten, so a
// bound computed from older values stays valid; a group inserted meanwhile changes the epoch and the swap fails.
__device__ __forceinline__ void mm_bound_refresh(const FrontEnd& fe, int lane) {
    if constexpr (Q::NMM > 0 && FG > 0) {
#pragma unroll
        for (int m = 0; m < Q::NMM; m++) {
            const bool ismin = (Q::MM_ISMIN >> m) & 1u;
            const unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(fe.mm_bound + m);
            const int cnt = min((int)*reinterpret_cast<volatile uint32_t*>(fe.dir_count), FG);      // read AFTER the epoch
            uint32_t x = ismin ? 0u : 0xFFFFFFFFu;
            for (int g = lane; g < cnt; g += 32) {
                const uint32_t hi = (uint32_t)(*reinterpret_cast<volatile uint64_t*>(fe.mm + g * Q::NMM + m) >> 32);
                x = ismin ? max(x, hi) : min(x, hi);
            }
            x = ismin ? __reduce_max_sync(0xffffffffu, x) : __reduce_min_sync(0xffffffffu, x);
            if (lane == 0 && cnt > 0) atomicCAS(fe.mm_bound + m, old, (old & 0xFFFFFFFF00000000ull) | x);
        }
    }
}

// Branch-free front-end accumulate of the adjacent row pair (r0, r0 + 1), groups g0 / g1 (FG = trash).
// Every lane-private load of BOTH rows is issued before the first dependent add, so a pair exposes one
// shared-memory latency instead of one per slot. When both rows fall in the same group the second row's add
// chains on the first row's result rather than on the (stale) value it loaded, and its store lands last.
// MIN/MAX: a value is first compared, on the HIGH WORD of its order-mapped form, with a bound that holds for EVERY
// front-end group (mmb: the largest of the groups' minima / the smallest of their maxima, see fe.mm_bound) — two
// register compares per row, no shared-memory access; only a value that beats the bound (or ties, or is a NaN)
// takes the rarely executed branch with the exact 64-bit compare + reduction on its own group's slot.
__device__ __forceinline__ void fe_accumulate_pair(uint32_t a_cnt, uint32_t a_sum, uint32_t a_mm, int g0, int g1, int lane, const AggSink& sink, int r0,
                                                   const uint32_t (&mmb)[Q::NMM > 0 ? Q::NMM : 1]) {
    constexpr int NI = Q::NIN > 0 ? Q::NIN : 1;
    const int r1 = r0 + 1;
    const uint32_t l4 = (uint32_t)lane * 4u, l8 = (uint32_t)lane * 8u;
    uint32_t c0a[2] = {0, 0}, c0v[2] = {0, 0};
    if constexpr (Q::CNT0_USED) {
        c0a[0] = a_cnt + (uint32_t)g0 * (Q::NCNT * 128u) + l4; c0a[1] = a_cnt + (uint32_t)g1 * (Q::NCNT * 128u) + l4;
        c0v[0] = lds_u32(c0a[0]); c0v[1] = lds_u32(c0a[1]);
    }
    uint32_t gi[NI][2], ca[NI][2], cv[NI][2], sa[NI][2];
    uint64_t sv[NI][2];
#pragma unroll
    for (int i = 0; i < Q::NIN; i++) {
        const int FL = Q::IN_FLAGS[i];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const bool valid = (sink.inok[i] >> (r0 + k)) & 1u;
            gi[i][k] = (Q::IN_CNT[i] > 0 && !valid) ? (uint32_t)FG : (uint32_t)(k ? g1 : g0);      // statically non-null inputs: always valid
            if (Q::IN_CNT[i] > 0) { ca[i][k] = a_cnt + (gi[i][k] * Q::NCNT + Q::IN_CNT[i]) * 128u + l4; cv[i][k] = lds_u32(ca[i][k]); }
            if (FL & F_SUM) { sa[i][k] = a_sum + (gi[i][k] * Q::NSUM + Q::FE_SUM[i]) * 256u + l8; sv[i][k] = lds_u64(sa[i][k]); }
        }
    }
    if constexpr (Q::CNT0_USED) {
        const uint32_t y0 = c0v[0] + 1u, y1 = (g0 == g1 ? y0 : c0v[1]) + 1u;
        sts_u32(c0a[0], y0); sts_u32(c0a[1], y1);
    }
#pragma unroll
    for (int i = 0; i < Q::NIN; i++) {
        const int FL = Q::IN_FLAGS[i];
        const bool same = gi[i][0] == gi[i][1];
        if (Q::IN_CNT[i] > 0) {
            const uint32_t y0 = cv[i][0] + 1u, y1 = (same ? y0 : cv[i][1]) + 1u;
            sts_u32(ca[i][0], y0); sts_u32(ca[i][1], y1);
        }
        if (FL & F_SUM) {
            const uint64_t x0 = sink.in[i][r0], x1 = sink.in[i][r1];
            uint64_t y0, y1;
            if (FL & F_INT) { y0 = sv[i][0] + x0; y1 = (same ? y0 : sv[i][1]) + x1; }
            else { y0 = as_u64(__dadd_rn(as_f64(sv[i][0]), as_f64(x0))); y1 = as_u64(__dadd_rn(as_f64(same ? y0 : sv[i][1]), as_f64(x1))); }
            sts_u64(sa[i][0], y0); sts_u64(sa[i][1], y1);
        }
        if (FL & (F_MIN | F_MAX)) {
            const bool is_int = (FL & F_INT) != 0;
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const uint64_t v = sink.in[i][r0 + k];
                const uint32_t hi = (uint32_t)(v >> 32);
                const uint32_t mh = is_int ? hi ^ 0x80000000u : hi ^ ((uint32_t)((int32_t)hi >> 31) | 0x80000000u);     // high word of order_map(v)
                bool exact = false;
                if (FL & F_MIN) exact |= mh <= mmb[Q::FE_MIN[i] < 0 ? 0 : Q::FE_MIN[i]];
                if (FL & F_MAX) exact |= mh >= mmb[Q::FE_MAX[i] < 0 ? 0 : Q::FE_MAX[i]];
                if (exact && (is_int || as_f64(v) == as_f64(v))) {                                                        // a NaN never replaces a held value (Main.kt:552)
                    const uint64_t m = order_map(v, is_int);
                    if (FL & F_MIN) smem_min_u64(a_mm + (gi[i][k] * Q::NMM + Q::FE_MIN[i]) * 8u, m);
                    if (FL & F_MAX) smem_max_u64(a_mm + (gi[i][k] * Q::NMM + Q::FE_MAX[i]) * 8u, m);
                }
            }
        }
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

// ---- partitioned path (high cardinality) ------------------------------------------------------------------------------
// With millions of groups every row of the plain path is a random read-modify-write in HBM/L2 (latency- and
// atomic-bound, a few G rows/s). Instead: PASS 1 (this kernel, KQ_AGG_MODE 1) evaluates the query's expressions
// and scatters one TUPLE per selected row into HBM scratch, bucketed by (hash partition, CTA) — a CTA appends to
// its own buckets through shared-memory cursors, so there are no global atomics and the L2 merges the appends
// into full lines; PASS 2 (kq_agg_partition_reduce) reduces one partition at a time in a shared-memory hash
// table — a partition holds a few thousand distinct keys — and merges the table into the global one, once per
// distinct key instead of once per row. Rows that do not fit (bucket or table full: skewed keys) take the plain
// global path, so the result never depends on the partition geometry.
// Tuple: [key words][meta word if anything is nullable: key null mask | input validity bits << 8][one word per
// input that carries a value (inputs used by COUNT only carry none)].
#ifndef KQ_AGG_MODE
#define KQ_AGG_MODE 0
#endif
#ifndef KQ_PART_L2_HINTS
#define KQ_PART_L2_HINTS 1           // measured: pass 1 2.61 -> 2.44 ms per 100 M rows (fewer partially filled bucket lines evicted)
#endif
constexpr bool TUPLE_META = Q::KEYS_NULLABLE || Q::NCNT > 1;
__host__ __device__ constexpr int tuple_in_word(int i) {
    int w = Q::NKEYS + (TUPLE_META ? 1 : 0);
    for (int j = 0; j < i; j++) if (Q::IN_FLAGS[j] & (F_SUM | F_MIN | F_MAX)) w++;
    return w;
}
constexpr int TW = tuple_in_word(Q::NIN) > 0 ? tuple_in_word(Q::NIN) : 1;
// 32-bit multiplicative hash of a key for the shared-memory table of pass 2 / of the block-shared front end: a different
// function from hash_key, whose top bits are equal for all keys of a partition.
__device__ __forceinline__ uint32_t part_hash(const uint64_t (&kw)[MAX_KEYS], uint32_t nullmask) {
    uint32_t h = Q::KEYS_NULLABLE ? nullmask * 0xC2B2AE35u : 0u;
#pragma unroll
    for (int k = 0; k < Q::NKEYS; k++) h = (h ^ (uint32_t)kw[k] ^ ((uint32_t)(kw[k] >> 32) * 0x85EBCA6Bu)) * 0x9E3779B1u;
    return h;
}

__device__ __forceinline__ uint32_t atoms_inc(uint32_t addr) {
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(addr) : "memory");
    return old;
}

// Scatter the selected rows of this thread into the CTA's buckets; returns the rows whose bucket is full.
__device__ __forceinline__ uint32_t partition_scatter(const AggArgs& A, uint32_t a_cursor, const AggSink& sink, const uint32_t (&nm)[R], uint32_t rows) {
    uint32_t part[R], pos[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        part[r] = 0; pos[r] = 0xFFFFFFFFu;
        if ((rows >> r) & 1u) {
            uint64_t kw[MAX_KEYS];
#pragma unroll
            for (int k2 = 0; k2 < MAX_KEYS; k2++) kw[k2] = k2 < Q::NKEYS ? sink.key[k2][r] : 0;
            part[r] = (uint32_t)(hash_key(kw, nm[r], Q::NKEYS) >> (64 - A.part_log2));      // = top bits of the key's home slot in the global table
            pos[r] = atoms_inc(a_cursor + part[r] * 4u);
        }
    }
    uint32_t overflow = 0;
    const uint64_t st_policy = KQ_PART_L2_HINTS ? l2_policy_evict_last() : 0ULL;
#pragma unroll
    for (int r = 0; r < R; r++) {
        if (!((rows >> r) & 1u)) continue;
        if (pos[r] >= (uint32_t)A.part_cap) { overflow |= 1u << r; continue; }
        uint64_t* t = A.part_scratch + (((uint64_t)part[r] * (uint32_t)A.part_ncta + blockIdx.x) * (uint32_t)A.part_cap + pos[r]) * TW;
        uint64_t w[TW];
#pragma unroll
        for (int k2 = 0; k2 < Q::NKEYS; k2++) w[k2] = sink.key[k2][r];
        if constexpr (TUPLE_META) {
            uint32_t meta = nm[r];
#pragma unroll
            for (int i = 0; i < Q::NIN; i++) meta |= ((sink.inok[i] >> r) & 1u) << (8 + i);
            w[Q::NKEYS] = meta;
        }
#pragma unroll
        for (int i = 0; i < Q::NIN; i++) if (Q::IN_FLAGS[i] & (F_SUM | F_MIN | F_MAX)) w[tuple_in_word(i)] = sink.in[i][r];
        if constexpr (TW % 2 == 0) {
#pragma unroll
            for (int j = 0; j < TW; j += 2) {
                if (KQ_PART_L2_HINTS) asm volatile("st.global.L2::cache_hint.v2.u64 [%0], {%1, %2}, %3;" ::"l"(t + j), "l"(w[j]), "l"(w[j + 1]), "l"(st_policy) : "memory");
                else *reinterpret_cast<ulonglong2*>(t + j) = make_ulonglong2(w[j], w[j + 1]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < TW; j++) t[j] = w[j];
        }
    }
    return overflow;
}

#if KQ_AGG_MODE >= 1
struct PartTable {            // hash table in shared memory: sparse lookup part (SL slots) -> dense accumulator rows (AC)
    uint32_t* state;          // [SL] 0 empty, 1 being published, else (accumulator row + 2) | key null mask << 24
    uint64_t* key;            // [NKEYS][SL]
    uint64_t* sum;            // [NSUM][AC]
    uint64_t* mm;             // [NMM][AC]  order-mapped MIN/MAX
    uint32_t* cnt;            // [NCNT][AC] row 0: rows (statically non-null inputs), else one per nullable input
    uint32_t* nfull;          // accumulator rows handed out
    int SL, AC;
    uint32_t limit;           // stop inserting at this many rows (every thread may have one insert in flight)
};

template <int I>
__device__ __forceinline__ void part_accumulate_input(const PartTable& T, uint32_t slot, uint32_t meta, const uint64_t (&w)[TW]) {
    if constexpr (I < Q::NIN) {
        constexpr int FL = Q::IN_FLAGS[I];
        const bool valid = Q::IN_CNT[I] > 0 ? ((meta >> (8 + I)) & 1u) != 0 : true;
        if (valid) {
            if constexpr (Q::IN_CNT[I] > 0) atomicAdd(T.cnt + (size_t)Q::IN_CNT[I] * T.AC + slot, 1u);
            if constexpr ((FL & (F_SUM | F_MIN | F_MAX)) != 0) {
                const uint64_t v = w[tuple_in_word(I)];
                if constexpr ((FL & F_SUM) != 0) {
                    uint64_t* p = T.sum + (size_t)Q::FE_SUM[I] * T.AC + slot;
                    if constexpr ((FL & F_INT) != 0) atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
                    else atomicAdd(reinterpret_cast<double*>(p), as_f64(v));
                }
                if constexpr ((FL & (F_MIN | F_MAX)) != 0) {
                    constexpr bool is_int = (FL & F_INT) != 0;
                    const uint64_t m = order_map(v, is_int);
                    const bool cmp = is_int || as_f64(v) == as_f64(v);          // a NaN never replaces a held value (Main.kt:552)
                    if constexpr ((FL & F_MIN) != 0) {
                        uint64_t* p = T.mm + (size_t)Q::FE_MIN[I] * T.AC + slot;
                        if (cmp && m < *reinterpret_cast<volatile uint64_t*>(p)) atomicMin(reinterpret_cast<unsigned long long*>(p), (unsigned long long)m);
                    }
                    if constexpr ((FL & F_MAX) != 0) {
                        uint64_t* p = T.mm + (size_t)Q::FE_MAX[I] * T.AC + slot;
                        if (cmp && m > *reinterpret_cast<volatile uint64_t*>(p)) atomicMax(reinterpret_cast<unsigned long long*>(p), (unsigned long long)m);
                    }
                }
            }
        }
        part_accumulate_input<I + 1>(T, slot, meta, w);
    }
}
template <int I>
__device__ __forceinline__ void part_overflow_input(const AggArgs& A, uint64_t* rec, uint32_t meta, const uint64_t (&w)[TW]) {
    if constexpr (I < Q::NIN) {
        const bool valid = Q::IN_CNT[I] > 0 ? ((meta >> (8 + I)) & 1u) != 0 : true;
        if (valid) global_accumulate(rec, A.in[I], (Q::IN_FLAGS[I] & (F_SUM | F_MIN | F_MAX)) ? w[tuple_in_word(I)] : 0ULL);
        part_overflow_input<I + 1>(A, rec, meta, w);
    }
}
// OWNED: the record was claimed by this thread and is not published yet — plain stores instead of atomics.
template <int I, bool OWNED = false>
__device__ __forceinline__ void part_merge_input(const PartTable& T, uint32_t slot, uint64_t* rec) {
    if constexpr (I < Q::NIN) {
        constexpr int FL = Q::IN_FLAGS[I];
        const uint32_t n = T.cnt[(size_t)Q::IN_CNT[I] * T.AC + slot];
        if constexpr (OWNED) {
            rec[Q::REC_NN[I]] = n;
            if constexpr ((FL & F_SUM) != 0) rec[Q::REC_SUM[I]] = n ? T.sum[(size_t)Q::FE_SUM[I] * T.AC + slot] : 0ULL;
            if constexpr ((FL & F_MIN) != 0) rec[Q::MM_WORD[Q::FE_MIN[I]]] = T.mm[(size_t)Q::FE_MIN[I] * T.AC + slot];      // identity ~0 when no value was seen
            if constexpr ((FL & F_MAX) != 0) rec[Q::MM_WORD[Q::FE_MAX[I]]] = T.mm[(size_t)Q::FE_MAX[I] * T.AC + slot];      // identity 0
        } else if (n) {
            atomicAdd(reinterpret_cast<unsigned long long*>(rec + Q::REC_NN[I]), (unsigned long long)n);
            if constexpr ((FL & F_SUM) != 0) {
                const uint64_t x = T.sum[(size_t)Q::FE_SUM[I] * T.AC + slot];
                if constexpr ((FL & F_INT) != 0) atomicAdd(reinterpret_cast<unsigned long long*>(rec + Q::REC_SUM[I]), (unsigned long long)x);
                else atomicAdd(reinterpret_cast<double*>(rec + Q::REC_SUM[I]), as_f64(x));
            }
            if constexpr ((FL & F_MIN) != 0) {
                const uint64_t x = T.mm[(size_t)Q::FE_MIN[I] * T.AC + slot];
                if (x != ~0ULL) atomicMin(reinterpret_cast<unsigned long long*>(rec + Q::MM_WORD[Q::FE_MIN[I]]), (unsigned long long)x);
            }
            if constexpr ((FL & F_MAX) != 0) {
                const uint64_t x = T.mm[(size_t)Q::FE_MAX[I] * T.AC + slot];
                if (x != 0ULL) atomicMax(reinterpret_cast<unsigned long long*>(rec + Q::MM_WORD[Q::FE_MAX[I]]), (unsigned long long)x);
            }
        }
        part_merge_input<I + 1, OWNED>(T, slot, rec);
    }
}

// One tuple per lane into the partition's shared-memory table (find or insert, then shared-memory atomics).
// Called by ALL 32 lanes (`valid` = the lane has a tuple). The probe runs in LOCKSTEP: every unresolved lane takes
// one probe step per iteration and the warp leaves the loop together. Two reasons: (1) a lane that meets a slot being
// published by another lane of its own warp simply looks again next iteration (by then the publisher, which runs
// in the same iteration, is done) — a free-running spin could starve the publisher; (2) with free-running loops
// the compiler emits no reconvergence point and the lanes of a warp drift apart for good (measured: 4 active lanes
// per instruction). The warp pays for its slowest lane, hence the sparse lookup part (probe sequences of 1-4 slots).
__device__ __forceinline__ void part_accumulate(const AggArgs& A, const PartTable& T, const uint64_t (&w)[TW], bool valid) {
    uint64_t kw[MAX_KEYS];
#pragma unroll
    for (int k = 0; k < MAX_KEYS; k++) kw[k] = k < Q::NKEYS ? w[k] : 0;
    const uint32_t meta = TUPLE_META ? (uint32_t)w[Q::NKEYS] : 0u;
    const uint32_t nm = meta & 0xffu;
    // slot hash: independent of the partition bits (top bits of hash_key); finaliser so that low bits are well mixed
    uint32_t h = part_hash(kw, nm);
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    uint32_t slot = h & (uint32_t)(T.SL - 1);
    uint32_t row = 0xFFFFFFFFu;                 // accumulator row once found
    bool probing = valid;
    while (__any_sync(0xffffffffu, probing)) {
        if (probing) {
            uint32_t st = *reinterpret_cast<volatile uint32_t*>(T.state + slot);
            if (st == 0u) {
                if (*reinterpret_cast<volatile uint32_t*>(T.nfull) >= T.limit) probing = false;       // accumulator rows nearly used up: global path
                else {
                    st = atomicCAS(T.state + slot, 0u, 1u);
                    if (st == 0u) {
                        row = atomicAdd(T.nfull, 1u);
#pragma unroll
                        for (int k = 0; k < Q::NKEYS; k++) T.key[(size_t)k * T.SL + slot] = kw[k];
                        __threadfence_block();
                        *reinterpret_cast<volatile uint32_t*>(T.state + slot) = (row + 2u) | (nm << 24);
                        probing = false;
                    }
                }
            }
            if (probing && st > 1u) {           // st == 1: being published, look at the same slot again next iteration
                bool eq = (st >> 24) == nm;
#pragma unroll
                for (int k = 0; k < Q::NKEYS; k++) eq &= *reinterpret_cast<volatile uint64_t*>(T.key + (size_t)k * T.SL + slot) == kw[k];
                if (eq) { row = (st & 0xFFFFFFu) - 2u; probing = false; }
                else slot = (slot + 1) & (uint32_t)(T.SL - 1);
            }
        }
    }
    if (row != 0xFFFFFFFFu) {
        if constexpr (Q::CNT0_USED) atomicAdd(T.cnt + row, 1u);
        part_accumulate_input<0>(T, row, meta, w);
    } else if (valid) {
        uint64_t* rec = table_find_or_insert(A, hash_key(kw, nm, Q::NKEYS), kw, nm);
        part_overflow_input<0>(A, rec, meta, w);
    }
}

// Merge the shared-memory table into the global one: once per distinct key. `fresh` counts the groups this thread added.
// Up to MK table entries per thread are merged TOGETHER: the header loads, the claiming CAS and the accumulator atomics
// of all of them are in flight at once (one entry at a time is a chain of four dependent L2/HBM round trips), and one
// fence covers all the records the thread publishes.
// ONLY for pass 2 of the partitioned path: there a partition's keys belong to one block, so no other thread inserts the
// SAME key while this runs and a slot that is BUSY (being published, necessarily with another key) is treated like any
// occupied slot — nobody waits while holding a claim, which is what makes several claims per thread deadlock-free.
template <int MK>
__device__ __forceinline__ void part_merge_table_batched(const AggArgs& A, const PartTable& T, int tid, int nthreads, uint32_t& fresh) {
    constexpr int NK1 = Q::NKEYS > 0 ? Q::NKEYS : 1;
    uint64_t* const tab_end = A.table + (A.cap_mask + 1) * (uint64_t)A.stride;
    for (int base = 0; base < T.SL; base += nthreads * MK) {
        uint64_t* rec[MK];
        uint64_t key[MK][NK1];
        uint32_t row[MK], nmv[MK];
        uint32_t have = 0;
#pragma unroll
        for (int j = 0; j < MK; j++) {
            const int slot = base + j * nthreads + tid;
            const uint32_t st = slot < T.SL ? T.state[slot] : 0u;
            row[j] = 0; nmv[j] = 0; rec[j] = A.table;
#pragma unroll
            for (int k = 0; k < NK1; k++) key[j][k] = 0;
            if (st >= 2u) {
                have |= 1u << j;
                uint64_t kw[MAX_KEYS];
#pragma unroll
                for (int k = 0; k < MAX_KEYS; k++) kw[k] = k < Q::NKEYS ? T.key[(size_t)k * T.SL + slot] : 0;
#pragma unroll
                for (int k = 0; k < Q::NKEYS; k++) key[j][k] = kw[k];
                nmv[j] = st >> 24; row[j] = (st & 0xFFFFFFu) - 2u;
                rec[j] = A.table + table_home(hash_key(kw, nmv[j], Q::NKEYS), A.cap_mask) * (uint64_t)A.stride;
            }
        }
        // Every probe is ONE atomic round trip: CAS(EMPTY -> BUSY) either claims the slot or returns the header that is there
        // (a separate load first would double the latency of the common case, an empty home slot).
        uint32_t pend = have, mine = 0;
        while (pend) {
            uint64_t hdr[MK];
#pragma unroll
            for (int j = 0; j < MK; j++)        // all CAS in flight before the first result is looked at
                if ((pend >> j) & 1u) hdr[j] = atomicCAS(reinterpret_cast<unsigned long long*>(rec[j]), 0ULL, HDR_BUSY | ((uint64_t)nmv[j] << 32));
#pragma unroll
            for (int j = 0; j < MK; j++) {
                if (!((pend >> j) & 1u)) continue;
                if (hdr[j] == 0ULL) { mine |= 1u << j; pend &= ~(1u << j); continue; }            // claimed: ours to fill in
                bool eq = hdr[j] == (HDR_FULL | ((uint64_t)nmv[j] << 32));
                if (eq) {
#pragma unroll
                    for (int k = 0; k < Q::NKEYS; k++) eq &= __ldcg(rec[j] + 1 + k) == key[j][k];
                }
                if (eq) pend &= ~(1u << j);                                                       // the group exists already
                else { uint64_t* p = rec[j] + A.stride; rec[j] = p == tab_end ? A.table : p; }    // another key, or one being published
            }
        }
        // claimed records: keys and the final accumulator values with plain stores, ONE fence, then the headers
#pragma unroll
        for (int j = 0; j < MK; j++) {
            if (!((mine >> j) & 1u)) continue;
            for (int w = 1 + Q::NKEYS; w < A.stride; w++) rec[j][w] = A.rec_init[w];
#pragma unroll
            for (int k = 0; k < Q::NKEYS; k++) rec[j][1 + k] = key[j][k];
            part_merge_input<0, true>(T, row[j], rec[j]);
        }
        if (mine) asm volatile("fence.acq_rel.gpu;" ::: "memory");
#pragma unroll
        for (int j = 0; j < MK; j++)
            if ((mine >> j) & 1u) { *reinterpret_cast<volatile uint64_t*>(rec[j]) = HDR_FULL | ((uint64_t)nmv[j] << 32); fresh++; }
#pragma unroll
        for (int j = 0; j < MK; j++) if (((have & ~mine) >> j) & 1u) part_merge_input<0>(T, row[j], rec[j]);
    }
}
__device__ __forceinline__ void part_merge_table(const AggArgs& A, const PartTable& T, int tid, int nthreads, uint32_t& fresh) {
    for (int slot = tid; slot < T.SL; slot += nthreads) {
        const uint32_t st = T.state[slot];
        if (st < 2u) continue;
        uint64_t kw[MAX_KEYS];
#pragma unroll
        for (int k = 0; k < MAX_KEYS; k++) kw[k] = k < Q::NKEYS ? T.key[(size_t)k * T.SL + slot] : 0;
        const uint32_t nm = st >> 24;
        uint64_t* rec = table_find_or_insert(A, hash_key(kw, nm, Q::NKEYS), kw, nm, &fresh);
        part_merge_input<0>(T, (st & 0xFFFFFFu) - 2u, rec);
    }
}
// Carve the table out of shared memory at p (16-byte aligned); returns its size in bytes (multiple of 16).
__device__ __forceinline__ size_t part_table_place(PartTable& T, unsigned char* p, int SL, int AC, int nthreads, uint32_t* nfull) {
    T.SL = SL; T.AC = AC; T.nfull = nfull;
    T.limit = AC > nthreads ? (uint32_t)(AC - nthreads) : 0u;      // every thread may have one insert in flight past the check
    T.key = reinterpret_cast<uint64_t*>(p);
    T.sum = T.key + (size_t)Q::NKEYS * SL;
    T.mm = T.sum + (size_t)Q::NSUM * AC;
    T.cnt = reinterpret_cast<uint32_t*>(T.mm + (size_t)Q::NMM * AC);
    T.state = T.cnt + (size_t)Q::NCNT * AC;
    return ((size_t)SL * (8 * Q::NKEYS + 4) + (size_t)AC * (8 * (Q::NSUM + Q::NMM) + 4 * Q::NCNT) + 15) / 16 * 16;
}

#endif  // KQ_AGG_MODE >= 1

extern "C" __global__ void __launch_bounds__(THREADS, 1) kq_hash_aggregate(const __grid_constant__ AggArgs A) {
    // shared memory: [S stages][directory][gslot][gid2slot][mm][per-warp sums][per-warp counts]
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[S], empty[S];
    __shared__ long long tile_of[S];
    __shared__ long long bbase[S][MAX_COLS];
    __shared__ uint32_t s_dir_count;
    __shared__ unsigned long long s_mm_bound[Q::NMM > 0 ? Q::NMM : 1];
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
    p0 = smem + (((size_t)(p0 - smem) + 15) & ~(size_t)15);
    uint32_t* part_cursor = reinterpret_cast<uint32_t*>(p0);   // pass 1 of the partitioned path: tuples appended per partition
    if (KQ_AGG_MODE == 1) p0 += (size_t)A.nparts * 4;
#if KQ_AGG_MODE == 2
    // mid cardinality: ONE block-shared table (sparse lookup part + accumulator rows, shared-memory atomics) instead of
    // the lane-private front end, which only holds a few dozen groups; merged into the global table at block exit
    __shared__ uint32_t s_tab_nfull;
    PartTable T2;
    p0 += part_table_place(T2, p0, A.part_slots, A.part_groups, THREADS, &s_tab_nfull);
    if (threadIdx.x == 0) s_tab_nfull = 0;
#endif
    const size_t fe_end = (size_t)(p0 - smem);
    fe.dir_count = &s_dir_count;
    fe.mm_bound = s_mm_bound;
    fe.sum = sum0 + (size_t)(warp < 0 ? 0 : warp) * FG1 * NSUM * 32;
    fe.cnt = cnt0 + (size_t)(warp < 0 ? 0 : warp) * FG1 * NCNT * 32;

    for (size_t i = fe_begin + threadIdx.x * 4; i < fe_end; i += THREADS * 4) *reinterpret_cast<uint32_t*>(smem + i) = 0;
    if (threadIdx.x == 0) {
        s_dir_count = 0;
        for (int m = 0; m < Q::NMM; m++) s_mm_bound[m] = ((Q::MM_ISMIN >> m) & 1u) ? 0xFFFFFFFFull : 0ull;
        for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], WARPS); }
        mbar_fence_init();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < FG1 * NMM; i += THREADS) fe.mm[i] = ((Q::MM_ISMIN >> (i % (NMM > 0 ? NMM : 1))) & 1u) ? ~0ULL : 0ULL;
#if KQ_AGG_MODE == 2
    for (int i = threadIdx.x; i < NMM * T2.AC; i += THREADS) if ((Q::MM_ISMIN >> (i / T2.AC)) & 1u) T2.mm[i] = ~0ULL;
#endif
    __syncthreads();

    if (wid == PRODUCER_WARP) {
        if (lane == 0) {
            // The next ticket is always requested one step early: the L2 round trip of the atomic and the HBM
            // reads of the tile's Utf8 boundary offsets (stage_bounds_fetch) overlap the wait for a free stage,
            // so string bytes and fixed-size buffers of a tile are issued together on one barrier.
            TileBounds tb = {};
            // partition scatter: the input streams through once (evict first) while the partially filled lines of the
            // buckets should stay in the L2 until they are complete (the stores below ask for evict last)
            const uint64_t in_policy = (KQ_AGG_MODE == 1 && KQ_PART_L2_HINTS) ? l2_policy_evict_first() : 0ULL;
            auto take = [&]() -> long long {
                // stop taking tiles once the global table is half full: every ticket taken is processed,
                // so the rows consumed so far are always a prefix of the batch (the host grows and resumes)
                const unsigned long long g = *reinterpret_cast<volatile unsigned long long*>(A.ngroups);
                if (g > A.stop_threshold) return -1;
                const long long t = (long long)atomicAdd(A.ticket, 1u) + A.tile_begin;
                if (t >= A.ntiles) return -1;
                if (KQ_STAGE_BYTES) stage_bounds_fetch(A.sp, t, TILE, A.n, tb);
                return t;
            };
            long long next = take();
            for (int kp = 0;; kp++) {
                const int s = kp % S;
                while (!mbar_test(&empty[s], ((kp / S) & 1) ^ 1)) __nanosleep(64);      // a spinning producer would take issue slots from the consumer warp on its scheduler
                const long long tile = next;
                tile_of[s] = tile;
                if (tile < 0) { mbar_arrive(&full[s]); break; }
                stage_issue_all(A.sp, smem + (size_t)s * A.sp.stage_bytes, &full[s], tile, TILE, A.n, tb, bbase[s], in_policy);
                next = take();
                if (KQ_L2_PREFETCH > 0 && next >= 0) stage_prefetch_l2(A.sp, next, TILE, A.n);     // staged one step from now
            }
        }
    } else {
        AggSink sink;
        bool bypass = FG == 0;
        int low_tiles = 0;                        // consecutive tiles in which this warp mostly missed a full directory
        const uint32_t a_dir = smem_u32(fe.dir), a_cnt = smem_u32(fe.cnt), a_sum = smem_u32(fe.sum), a_mm = smem_u32(fe.mm);
        for (int k = 0;; k++) {
            const int s = k % S;
            mbar_wait(&full[s], (k / S) & 1);
            const long long tile = tile_of[s];
            if (tile < 0) break;
            RowCtx rc;
            rowctx_init(rc, warp, tile, TILE, A.n, A.err, smem + (size_t)s * A.sp.stage_bytes);
            rc.bbase = bbase[s];
            rc.heap = A.heap.tab ? &A.heap : nullptr;
            sink.sel = rc.inr;
            Q::eval(A.q, rc, sink);
#ifdef KQ_RING_CHECK
            if ((KQ_RING_CHECK & 2) && lane == 0 && A.trace) atomicAdd(reinterpret_cast<unsigned int*>(A.trace) + tile, 1u);      // debugging: warps per tile
#endif
#ifdef KQ_NO_STAGE_DRAIN
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
#else
            stage_release(&empty[s], &tile_of[s], (uint32_t)tile, A.err, lane);       // everything needed is in registers now
#endif

            // canonical key words + null masks of the R owned rows
            uint32_t nm[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                uint32_t nullmask = 0;
#pragma unroll
                for (int k2 = 0; k2 < Q::NKEYS; k2++) {
                    uint64_t w = 0;
                    if (!Q::KEYS_NULLABLE || ((sink.keyok[k2] >> r) & 1u)) w = ((Q::KEY_F64_MASK >> k2) & 1u) ? canon_nan(sink.key[k2][r]) : sink.key[k2][r];
                    else nullmask |= 1u << k2;
                    sink.key[k2][r] = w;
                }
                nm[r] = nullmask;
            }
            uint32_t slow = sink.sel;             // rows that still need the general path
            uint32_t new_groups = 0;              // groups this thread adds to the global table in this tile
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
                    // the key's home bucket: slots s0 (even) and s0 + 1 (linear probing rarely displaces a key further)
                    const uint32_t e0 = a_dir + dir_hash(kw, nm[r]) * (ENTRY_WORDS * 8u), e1 = e0 + ENTRY_WORDS * 8u;
                    const uint4 q0 = lds_u128(e0), q1 = lds_u128(e1);
                    // state word: gid + 2 for a published front-end group (EMPTY, BUSY and GLOBAL all fail the unsigned test)
                    bool h0 = q0.x - 2u < (uint32_t)FG, h1 = q1.x - 2u < (uint32_t)FG;
                    if constexpr (Q::KEYS_NULLABLE) { h0 &= q0.y == nm[r]; h1 &= q1.y == nm[r]; }
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
                // pass 2 (branch-free): unconditional read-modify-write of the lane-private slots, a row pair at a time
                // second chance, only when some lane missed: the NEXT bucket (slots home + 2, home + 3), which is where linear
                // probing puts a key whose home bucket was taken. Without it ONE displaced key among 50 sends half of all
                // warp-rows through the serial general path below (measured: 1.8 -> 0.45 TB/s on BASELINE config 3).
                if (__any_sync(0xffffffffu, slow != 0)) {
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        if (!((slow >> r) & 1u)) continue;
                        uint64_t kw[MAX_KEYS];
#pragma unroll
                        for (int k2 = 0; k2 < MAX_KEYS; k2++) kw[k2] = k2 < Q::NKEYS ? sink.key[k2][r] : 0;
                        const uint32_t e0 = a_dir + ((dir_hash(kw, nm[r]) + 2u) & (uint32_t)(DIR - 1)) * (ENTRY_WORDS * 8u), e1 = e0 + ENTRY_WORDS * 8u;
                        const uint4 q0 = lds_u128(e0), q1 = lds_u128(e1);
                        bool h0 = q0.x - 2u < (uint32_t)FG, h1 = q1.x - 2u < (uint32_t)FG;
                        if constexpr (Q::KEYS_NULLABLE) { h0 &= q0.y == nm[r]; h1 &= q1.y == nm[r]; }
                        if constexpr (Q::NKEYS >= 1) {
                            h0 &= ((uint64_t)q0.z | ((uint64_t)q0.w << 32)) == kw[0];
                            h1 &= ((uint64_t)q1.z | ((uint64_t)q1.w << 32)) == kw[0];
                        }
#pragma unroll
                        for (int k2 = 1; k2 < Q::NKEYS; k2++) { h0 &= lds_u64(e0 + 8u + 8u * k2) == kw[k2]; h1 &= lds_u64(e1 + 8u + 8u * k2) == kw[k2]; }
                        if (h0 | h1) { gid[r] = (int)((h0 ? q0.x : q1.x) - 2u); slow &= ~(1u << r); }
                    }
                }
                // bounds on the extremes of all groups: read AFTER the probes (a group this warp has seen is covered)
                uint32_t mmb[Q::NMM > 0 ? Q::NMM : 1];
#pragma unroll
                for (int m = 0; m < Q::NMM; m++) mmb[m] = (uint32_t)*reinterpret_cast<volatile unsigned long long*>(s_mm_bound + m);
#pragma unroll
                for (int j = 0; j < NCHUNK; j++) fe_accumulate_pair(a_cnt, a_sum, a_mm, gid[2 * j], gid[2 * j + 1], lane, sink, 2 * j, mmb);
                if (Q::NMM > 0 && ((k + warp) & 3) == 0) mm_bound_refresh(fe, lane);
                fe_hits = __popc(sink.sel & ~slow);
            }
            // general path for the rest: directory probing with insertion, else the global table
            int rows = __popc(sink.sel);
            // pass 1 of the partitioned path: rows whose bucket is full (skewed keys) stay in `slow` and take the
            // scalar global path below (table_find_or_insert + global_accumulate_all), like any other overflow row
            if (KQ_AGG_MODE == 1 && slow) { slow = partition_scatter(A, smem_u32(part_cursor), sink, nm, slow); }
#ifdef KQ_PART_DROP_SPILL
            if (KQ_AGG_MODE == 1) slow = 0;           // debugging: rows of full buckets vanish
#endif
#if KQ_AGG_MODE == 2
            if (true) {
#pragma unroll
                for (int r = 0; r < R; r++) {           // all 32 lanes: the table probe runs in lockstep
                    uint64_t w[TW];
#pragma unroll
                    for (int j = 0; j < TW; j++) w[j] = 0;
#pragma unroll
                    for (int k2 = 0; k2 < Q::NKEYS; k2++) w[k2] = sink.key[k2][r];
                    if constexpr (TUPLE_META) {
                        uint32_t meta = nm[r];
#pragma unroll
                        for (int i = 0; i < Q::NIN; i++) meta |= ((sink.inok[i] >> r) & 1u) << (8 + i);
                        w[Q::NKEYS] = meta;
                    }
#pragma unroll
                    for (int i = 0; i < Q::NIN; i++) if (Q::IN_FLAGS[i] & (F_SUM | F_MIN | F_MAX)) w[tuple_in_word(i)] = sink.in[i][r];
                    part_accumulate(A, T2, w, (slow >> r) & 1u);
                    __syncwarp();
                }
            }
#else
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
                        uint64_t* rec = table_find_or_insert(A, hash_key(kw, nm[r], Q::NKEYS), kw, nm[r], &new_groups);
                        global_accumulate_all<0>(A, rec, sink, r);
                    }
                }
            }
#endif
            // one update of the global group count per warp and tile (a single counter bumped by every insert serialises in the L2)
            if (__any_sync(0xffffffffu, new_groups != 0)) {
                const uint32_t tot = __reduce_add_sync(0xffffffffu, new_groups);
                if (lane == 0) atomicAdd(A.ngroups, (unsigned long long)tot);
            }
            // once the directory is full and this warp mostly misses it, stop probing it (high cardinality)
            // (two tiles in a row: while the blocks' first tiles race to fill the directory most rows find their key's slot
            // busy and miss, also when every key ends up fitting — measured as a 10x cliff when FG equals the group count)
            if (!bypass && __any_sync(0xffffffffu, fe_hits < rows) && *reinterpret_cast<volatile uint32_t*>(&s_dir_count) >= (uint32_t)FG) {
                int hits = fe_hits, tot = rows;
#pragma unroll
                for (int o = 16; o; o >>= 1) { hits += __shfl_xor_sync(0xffffffffu, hits, o); tot += __shfl_xor_sync(0xffffffffu, tot, o); }
                if (tot >= 64 && hits * 8 < tot) { if (++low_tiles >= 2) bypass = true; }
                else low_tiles = 0;
            } else low_tiles = 0;
#ifdef KQ_RING_CHECK
            if ((KQ_RING_CHECK & 4) && lane == 0 && A.trace) atomicAdd(reinterpret_cast<unsigned int*>(A.trace) + tile, 1u);      // at the END of the tile: timing of the ring untouched
#endif
        }
    }

    // ---- merge the front end into the global table ---------------------------------------------------
    __syncthreads();
    if (KQ_AGG_MODE == 1) {
        for (int p = threadIdx.x; p < A.nparts; p += THREADS)
            A.part_counts[(size_t)p * A.part_ncta + blockIdx.x] = min(part_cursor[p], (uint32_t)A.part_cap);
    }
#if KQ_AGG_MODE == 2
    {
        uint32_t fresh = 0;
        part_merge_table(A, T2, threadIdx.x, THREADS, fresh);
        const uint32_t tot = __reduce_add_sync(0xffffffffu, fresh);
        if (lane == 0 && tot) atomicAdd(A.ngroups, (unsigned long long)tot);
    }
#endif
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

#if KQ_AGG_MODE == 1
// ---- pass 2 of the partitioned path ------------------------------------------------------------------------------------
constexpr int PR_THREADS = 512, PR_WARPS = PR_THREADS / 32, PR_UNROLL = 4;

extern "C" __global__ void __launch_bounds__(PR_THREADS, 1) kq_agg_partition_reduce(const __grid_constant__ AggArgs A) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ int s_part;
    __shared__ uint32_t s_nfull, s_new;
    const int SL = A.part_slots, AC = A.part_groups;
    PartTable T;
    const size_t table_bytes = part_table_place(T, smem, SL, AC, PR_THREADS, &s_nfull);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto take_part = [&]() -> int {
        if (*reinterpret_cast<volatile unsigned long long*>(A.ngroups) > A.stop_threshold) return -1;
        const long long t = (long long)atomicAdd(A.ticket, 1u) + A.part_begin;
        return t < A.nparts ? (int)t : -1;
    };
    for (;;) {
        __syncthreads();                                    // the previous partition is merged
        if (tid == 0) {
            // same protocol as the tiles of the plain path: partitions are taken in order, and none is taken once the
            // global table is half full (the host grows it and resumes with the next partition)
            s_part = take_part(); s_nfull = 0; s_new = 0;
        }
        for (size_t i = (size_t)tid * 16; i < table_bytes; i += (size_t)PR_THREADS * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
        __syncthreads();
        const int p = s_part;
        if (p < 0) break;
        for (int i = tid; i < Q::NMM * AC; i += PR_THREADS) if ((Q::MM_ISMIN >> (i / AC)) & 1u) T.mm[i] = ~0ULL;
        __syncthreads();
        // The partition's buckets, one warp per bucket, in batches of 32 * PR_UNROLL tuples. The loads of the NEXT batch
        // (possibly of the warp's next bucket: the fill counts of all its buckets are fetched up front, one per lane)
        // are in flight while the current batch goes through the table.
        {
            const int my_seg = warp + PR_WARPS * lane;                  // lane l looks after the warp's l-th bucket
            const uint32_t my_n = my_seg < A.part_ncta ? A.part_counts[(size_t)p * A.part_ncta + my_seg] : 0u;
            const int nseg = (A.part_ncta - warp + PR_WARPS - 1) / PR_WARPS;     // buckets of this warp (<= 32: part_ncta <= 512)
            auto load_batch = [&](int si, uint32_t b0, uint32_t n, uint64_t (&w)[PR_UNROLL][TW]) {
                const uint64_t* base = A.part_scratch + ((uint64_t)p * (uint32_t)A.part_ncta + (uint32_t)(warp + PR_WARPS * si)) * (uint32_t)A.part_cap * TW;
#pragma unroll
                for (int u = 0; u < PR_UNROLL; u++) {
                    const uint32_t i = b0 + lane + 32u * u;
#pragma unroll
                    for (int j = 0; j < TW; j++) w[u][j] = 0;
                    if (i < n) {
                        const uint64_t* t = base + (uint64_t)i * TW;
                        if constexpr (TW % 2 == 0) {
#pragma unroll
                            for (int j = 0; j < TW; j += 2) { const ulonglong2 q = __ldcs(reinterpret_cast<const ulonglong2*>(t + j)); w[u][j] = q.x; w[u][j + 1] = q.y; }
                        } else {
#pragma unroll
                            for (int j = 0; j < TW; j++) w[u][j] = __ldcs(t + j);
                        }
                    }
                }
            };
            // cursor over (bucket, batch): warp-uniform
            int si = 0; uint32_t b0 = 0, n = 0;
            auto seek = [&]() {         // move (si, b0) to the next non-empty batch at or after the current position; false at the end
                while (si < nseg) {
                    n = __shfl_sync(0xffffffffu, my_n, si);
                    if (b0 < n) return true;
                    si++; b0 = 0;
                }
                return false;
            };
            uint64_t cur[PR_UNROLL][TW], nxt[PR_UNROLL][TW];
            bool have = seek();
            if (have) load_batch(si, b0, n, cur);
            while (have) {
                const uint32_t cb0 = b0, cn = n;
                b0 += 32 * PR_UNROLL;
                const bool more = seek();
                if (more) load_batch(si, b0, n, nxt);
#pragma unroll
                for (int u = 0; u < PR_UNROLL; u++) {
                    if (cb0 + 32u * u >= cn) break;                        // warp-uniform
                    part_accumulate(A, T, cur[u], cb0 + lane + 32u * u < cn);
                    __syncwarp();
                }
#pragma unroll
                for (int u = 0; u < PR_UNROLL; u++)
#pragma unroll
                    for (int j = 0; j < TW; j++) cur[u][j] = nxt[u][j];
                have = more;
            }
        }
        __syncthreads();
        // merge the table into the global one: once per distinct key of the partition
        uint32_t fresh = 0;
#ifdef KQ_PART_SCALAR_MERGE
        part_merge_table(A, T, tid, PR_THREADS, fresh);
#else
        part_merge_table_batched<4>(A, T, tid, PR_THREADS, fresh);
#endif
        if (fresh) atomicAdd(&s_new, fresh);
        __syncthreads();
        if (tid == 0 && s_new) atomicAdd(A.ngroups, (unsigned long long)s_new);       // one update per partition, not per group
    }
}
#endif  // KQ_AGG_MODE == 1

}  // namespace kq
