// kq_k_agg_fe.cuh — HashAggregateExec (Main.kt:605-660) for LOW-cardinality GROUP BYs (BASELINE configs 3 and 5): the
// drain loop's per-row work (key list, HashMap.getOrPut, Accumulator.accumulate; Main.kt:620-632) with every group of a
// CTA held in shared memory. Specialised per query like kq_k_agg.cuh (generated `struct Q`, same AggSink contract).
//
// What bounds this kernel is the shared-memory data path (one 128-byte wavefront per cycle and SM), not instruction
// issue and not HBM: at the copy roofline a 14-byte row leaves ~90 bytes of shared-memory traffic per row. So the design
// counts wavefronts:
//   * KEY -> GROUP: a CTA directory with PERFECT placement. slot = (lo * s1 + hi * s2 ...) >> shift; whenever an insert
//     collides with an occupied home slot the directory is rebuilt under new multipliers until every key sits in its
//     home slot (a handful of attempts at <= 6 % load). A probe is therefore ONE 4-byte state load + one 8-byte load per
//     key word, never a second slot; readers validate with a sequence number that rebuilds bump.
//   * ACCUMULATE: lane-private SUM / COUNT slots (one copy per lane per warp: plain read-modify-write, no atomics — the
//     shared-memory atomics of sm_100a are CAS loops for everything but 32-bit integers — and no bank conflicts by
//     construction).
//   * MIN / MAX: kept once per CTA; a row touches them only when its value beats a bound that holds for every group
//     with a value (register compares per row), or when it is the first value its lane sees in that group.
// Inserts and rebuilds are serialised by a CTA lock (whole warps only, a batch of keys per round); first values and bound
// refreshes are lock-free (a validity token frames them; a housekeeping warp does the refreshing). The per-row path takes
// no lock. Rows whose key does not fit the directory go to the global table of
// kq_aggtable.cuh with atomics; the CTA's groups are merged into it once, at exit.
//
// Accumulator semantics (oracle: MaxAccumulator, Main.kt:538-562, and the E5-E7 extensions): nulls are skipped; MIN/MAX
// use IEEE comparisons like the reference's `value > this.value`, so a NaN never replaces a held value; a group whose
// non-null values were all NaN yields NaN (the identity survives, see k_finalize); Float64 sums are reassociated.
#pragma once

#include "kq_rt.cuh"
#include "kq_aggtable.cuh"

namespace kq {

// Branch weights: the rarely taken paths (inserts, exact MIN/MAX updates, refreshes, ragged tiles) are laid out behind the
// hot loop, whose ~700 instructions then sit in a few contiguous kilobytes instead of being spread over the whole kernel
// (measured: instruction-cache misses were the largest single stall reason).
#define KQ_LIKELY(x) (__builtin_expect(!!(x), 1))
#define KQ_UNLIKELY(x) (__builtin_expect(!!(x), 0))

constexpr int WARPS = KQ_WARPS;              // consumer warps
constexpr int PRODUCER_WARP = 0;             // service warp first: the warp arbiter favours high warp ids
// Service warps (lowest ids): the producer, and — when the query has MIN/MAX — a housekeeping warp that keeps the bounds of the
// extremes fresh. (Refreshing from a consumer warp delayed that warp's next stage hand-back by ~1 us every couple of tiles,
// and with a two-stage ring the whole CTA waited for it.)
constexpr int NSERVICE = Q::NMM > 0 ? 2 : 1;
constexpr int THREADS = WARPS * 32 + 32 * NSERVICE;
constexpr int TILE = WARPS * WARP_ROWS;
constexpr int S = KQ_STAGES;
constexpr int FG = KQ_FE_GROUPS;             // directory capacity in groups (<= 254)
constexpr int DIR = KQ_DIR_SLOTS;            // directory slots (power of two, >= 8 * FG)
__host__ __device__ constexpr int ilog2c(int x) { return x <= 1 ? 0 : 1 + ilog2c(x >> 1); }
constexpr int DIR_SHIFT = 32 - ilog2c(DIR);
constexpr int NKW = Q::NKEYS > 0 ? Q::NKEYS : 1;          // key words per group (a global aggregate has one constant word)
constexpr int NMM1 = Q::NMM > 0 ? Q::NMM : 1;
constexpr int GS = Q::NSUM * 256 + Q::NCNT * 128;         // lane-private bytes per group and warp: [NSUM][32] u64, [NCNT][32] u32
constexpr int REBUILD_ATTEMPTS = 256;
// Every wait in this kernel is bounded: a spin that exceeds its budget raises a device error (bits 8..) and gives up, so a
// protocol bug shows up as a failed query with a location code instead of a hung GPU.
constexpr uint32_t ERR_SPIN_LOCK = 0x100u, ERR_SPIN_MM = 0x200u, ERR_SPIN_SLOW = 0x400u, ERR_SPIN_STAGE = 0x800u;
constexpr int SPIN_LIMIT = 1 << 22;

// What the generated code fills per tile: selection, key words and aggregate inputs of the R owned rows.
struct AggSink {
    uint32_t sel;
    uint64_t key[NKW][R];
    uint32_t keyok[NKW];
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

// Control block of the CTA directory (one 16-byte load per tile and warp).
struct __align__(16) DirCtl {
    uint32_t gen;            // sequence number: odd while a rebuild is in progress
    uint32_t s1, s2;         // multipliers of the placement hash
    uint32_t count;          // groups in the directory
};

struct Fe {                  // shared-memory addresses (32-bit) and pointers of the CTA front end
    uint32_t a_ctl, a_meta, a_keys;     // DirCtl; [DIR] u32 state; [DIR][NKW] u64 key words
    uint32_t a_mm;                      // [FG + 1][NMM] u64 order-mapped extremes (CTA-shared; row FG is a trash row)
    uint32_t a_lock, a_bound;           // the CTA lock (directory inserts); [NMM] u64 bounds on the extremes of all groups, as the input's own bits
    uint32_t a_btoken, a_started, a_finished, a_refreshing;      // bound validity token; first-value event counters; refresh flag
    uint32_t a_lane8, a_lane4;          // this warp's lane-private block + lane * 8 / + NSUM*256 + lane * 4
    DirCtl* ctl;
    uint32_t* meta;
    uint64_t* keys;
    uint64_t* gkeys;                    // [FG][NKW] dense list of the groups, in insertion order
    uint32_t* gnm;                      // [FG] their key null masks
    uint64_t* mm;
    uint32_t* limit;                    // no more inserts at this many groups (FG, or fewer after a failed rebuild)
    uint32_t* err;                      // the aggregate's device error word
    unsigned long long* trace;          // debugging (AggArgs::trace)
};
#ifdef KQ_FE_TRACE
#define KQ_FTRACE(code) do { if (fe.trace && blockIdx.x < 16 && (threadIdx.x & 31) == 0) reinterpret_cast<volatile unsigned long long*>(fe.trace)[blockIdx.x * 16 + (threadIdx.x >> 5)] = (unsigned long long)(code); } while (0)
#else
#define KQ_FTRACE(code) do { } while (0)
#endif

__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint64_t lds_u64(uint32_t a) { uint64_t v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds_u128(uint32_t a) { uint4 r; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a)); return r; }

// Shared-space atomics and volatile accesses by 32-bit shared address: no generic-address atomics on shared memory anywhere
// in this kernel (what the compiler emits for those differs per width — a native ATOM.E for 32 bits, a QSPC test plus a
// fallback for 64 — and none of it is needed when the state space is known).
__device__ __forceinline__ uint32_t sh_cas_u32(uint32_t a, uint32_t cmp, uint32_t val) { uint32_t old; asm volatile("atom.shared.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(a), "r"(cmp), "r"(val) : "memory"); return old; }
__device__ __forceinline__ uint32_t sh_exch_u32(uint32_t a, uint32_t val) { uint32_t old; asm volatile("atom.shared.exch.b32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(val) : "memory"); return old; }
__device__ __forceinline__ uint64_t sh_cas_u64(uint32_t a, uint64_t cmp, uint64_t val) { uint64_t old; asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(a), "l"(cmp), "l"(val) : "memory"); return old; }
// 64-bit MIN/MAX as an explicit compare-and-swap loop (the value only ever moves one way, so the loop is bounded by the
// number of competing writers)
__device__ __forceinline__ void sh_min_u64(uint32_t a, uint64_t v) {
    uint64_t cur = 0;
    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(cur) : "r"(a) : "memory");
    for (int i = 0; i < 4096 && v < cur; i++) { const uint64_t old = sh_cas_u64(a, cur, v); if (old == cur) break; cur = old; }
}
__device__ __forceinline__ void sh_max_u64(uint32_t a, uint64_t v) {
    uint64_t cur = 0;
    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(cur) : "r"(a) : "memory");
    for (int i = 0; i < 4096 && v > cur; i++) { const uint64_t old = sh_cas_u64(a, cur, v); if (old == cur) break; cur = old; }
}
__device__ __forceinline__ uint32_t sh_ld_u32(uint32_t a) { uint32_t v; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint64_t sh_ld_u64(uint32_t a) { uint64_t v; asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sh_st_u32(uint32_t a, uint32_t v) { asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sh_st_u64(uint32_t a, uint64_t v) { asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }

// identities of the CTA-shared extremes, in order-mapped form (never produced by a non-NaN value)
__device__ __forceinline__ constexpr uint64_t mm_identity(int m) { return ((Q::MM_ISMIN >> m) & 1u) ? ~0ULL : 0ULL; }
// "no bound": every value takes the exact path
__device__ __forceinline__ uint64_t bound_none(int m, bool is_int) {
    const bool ismin = (Q::MM_ISMIN >> m) & 1u;
    if (is_int) return ismin ? 0x7FFFFFFFFFFFFFFFULL : 0x8000000000000000ULL;       // = the identities: nothing is lost when they tie
    return 0x7FF8000000000000ULL;                                                   // NaN: `!(v >= bound)` / `!(v <= bound)` hold for every v
}
__host__ __device__ constexpr bool mm_is_int(int m) {
    for (int i = 0; i < Q::NIN; i++) if (Q::FE_MIN[i] == m || Q::FE_MAX[i] == m) return (Q::IN_FLAGS[i] & F_INT) != 0;
    return false;
}

// Placement hash of a key under the multipliers (s1, s2): distinct keys are separated by SOME pair of multipliers.
__device__ __forceinline__ uint32_t dir_slot(const uint64_t (&kw)[NKW], uint32_t nm, uint32_t s1, uint32_t s2) {
    uint32_t h = Q::KEYS_NULLABLE ? nm * 0x9E3779B1u : 0u;
#pragma unroll
    for (int k = 0; k < NKW; k++) {
        if (k > 0) h = h * 0x85EBCA6Bu + (h >> 15);
        h += (uint32_t)kw[k] * s1 + (uint32_t)(kw[k] >> 32) * s2;
    }
    return h >> DIR_SHIFT;
}

// The directory control block as ONE value for the whole warp. Lanes of a warp need not run in lockstep: each would read
// the (concurrently changing) block at its own time, and a branch on such a value that encloses warp-synchronising
// operations would split the warp for good. Every shared value that steers warp-level control flow is therefore read by
// lane 0 and broadcast.
__device__ __forceinline__ uint4 ctl_snapshot(const Fe& fe) {
    uint4 c;                                  // {gen, s1, s2, count}: volatile — other warps change it
    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w) : "r"(fe.a_ctl) : "memory");
    c.x = __shfl_sync(0xffffffffu, c.x, 0); c.y = __shfl_sync(0xffffffffu, c.y, 0);
    c.z = __shfl_sync(0xffffffffu, c.z, 0); c.w = __shfl_sync(0xffffffffu, c.w, 0);
    return c;
}
__device__ __forceinline__ uint32_t sh_ld_u32_uniform(uint32_t a) { return __shfl_sync(0xffffffffu, sh_ld_u32(a), 0); }

// One lock-free probe of the home slot: group id, or -1. Valid only if ctl->gen did not move meanwhile (caller checks).
template <bool AGAIN = false>          // AGAIN: a repeated probe (the general path) must not be satisfied from registers
__device__ __forceinline__ int dir_probe(const Fe& fe, const uint64_t (&kw)[NKW], uint32_t nm, uint32_t s1, uint32_t s2) {
    const uint32_t slot = dir_slot(kw, nm, s1, s2);
    const uint32_t m = AGAIN ? sh_ld_u32(fe.a_meta + slot * 4u) : lds_u32(fe.a_meta + slot * 4u);
    bool hit = true;
#pragma unroll
    for (int k = 0; k < NKW; k++) hit &= (AGAIN ? sh_ld_u64(fe.a_keys + slot * (8u * NKW) + 8u * k) : lds_u64(fe.a_keys + slot * (8u * NKW) + 8u * k)) == kw[k];
    const uint32_t x = Q::KEYS_NULLABLE ? m ^ (nm << 8) : m;
    const uint32_t g = x - 1u;                 // state: 0 = empty, else (gid + 1) | null mask << 8
    return (hit && g < (uint32_t)FG) ? (int)g : -1;
}

// The CTA lock is only ever taken by a whole, converged warp (lane 0 spins, the rest wait at the warp barrier): no thread
// of a warp ever waits for a lock that another thread of the same warp holds.
__device__ __forceinline__ void fe_lock(const Fe& fe, int lane, uint32_t site = 1u) {
    if (lane == 0) {
        int spins = 0;
        const uint32_t token = 0x100u + site * 16u + (threadIdx.x >> 5);       // who holds it (debugging: a waiter can tell)
        uint32_t held;
        while ((held = sh_cas_u32(fe.a_lock, 0u, token)) != 0u) {
            __nanosleep(64);
            ++spins;
            if ((spins & 0xFFFF) == 0) KQ_FTRACE(0x800000u | held);
            if (spins > SPIN_LIMIT) { atomicOr(fe.err, ERR_SPIN_LOCK); break; }
        }
    }
    __syncwarp();
    __threadfence_block();
}
__device__ __forceinline__ void fe_unlock(const Fe& fe, int lane) {
    __threadfence_block();
    __syncwarp();
    if (lane == 0) sh_exch_u32(fe.a_lock, 0u);
}

// Place groups [0, n) under the multipliers (s1, s2). Whole warp, under the lock, gen odd. false on any collision.
__device__ __forceinline__ bool dir_place_all(const Fe& fe, int n, uint32_t s1, uint32_t s2, int lane) {
    for (int i = lane; i < DIR; i += 32) fe.meta[i] = 0u;
    __syncwarp();
    bool ok = true;
    for (int g = lane; g < n; g += 32) {
        uint64_t kw[NKW];
#pragma unroll
        for (int k = 0; k < NKW; k++) kw[k] = fe.gkeys[g * NKW + k];
        const uint32_t nm = fe.gnm[g];
        const uint32_t slot = dir_slot(kw, nm, s1, s2);
        if (sh_cas_u32(fe.a_meta + slot * 4u, 0u, (uint32_t)(g + 1) | (nm << 8)) != 0u) ok = false;
        else {
#pragma unroll
            for (int k = 0; k < NKW; k++) fe.keys[slot * NKW + k] = kw[k];
        }
    }
    return __all_sync(0xffffffffu, ok);
}

// Insert up to 32 keys at once — every lane may propose one (`has`), all of them ABSENT from the directory (the caller
// probed under the lock). Whole warp, the warp holding the CTA lock. Lanes that propose the same key elect a leader;
// leaders take consecutive group ids, claim their home slots (BUSY marker -> key words -> state word, so that a lock-free
// reader never sees a state word next to another key's words) and one rebuild settles all collisions of the batch.
// Returns the group id of the lane's key, -1 if the directory could not take it (full, or unplaceable).
// (A CTA's first tile brings all keys of a low-cardinality query at once: three or four of these rounds instead of one
// round per key — measured 60 us less at the start of every kernel, which is most of a small batch.)
constexpr uint32_t META_BUSY = 0xFFFFFFFFu;          // fails dir_probe's `g < FG` test under every null mask
__device__ __forceinline__ int dir_insert_batch(const Fe& fe, bool has, const uint64_t (&kw)[NKW], uint32_t nm, int lane) {
    volatile DirCtl* ctl = fe.ctl;
    const uint32_t hasmask = __ballot_sync(0xffffffffu, has);
    uint32_t grp = 0;                                // lanes proposing the same key as this one
    if (has) {
        grp = __match_any_sync(hasmask, nm);
#pragma unroll
        for (int k = 0; k < NKW; k++) grp &= __match_any_sync(hasmask, (unsigned long long)kw[k]);
    }
    __syncwarp();
    const int head = has ? __ffs(grp) - 1 : 0;
    const bool leader = has && head == lane;
    const uint32_t leaders = __ballot_sync(0xffffffffu, leader);
    const int n = (int)sh_ld_u32_uniform(fe.a_ctl + 12u);
    const int limit = (int)sh_ld_u32_uniform(smem_u32(fe.limit));
    const int room = limit > n ? limit - n : 0;
    const int rank = __popc(leaders & ((1u << lane) - 1u));
    int gid = (leader && rank < room) ? n + rank : -1;
    const int add = min(__popc(leaders), room);
    if (add > 0) {
        if (gid >= 0) {
#pragma unroll
            for (int k = 0; k < NKW; k++) fe.gkeys[gid * NKW + k] = kw[k];
            fe.gnm[gid] = nm;
        }
        __threadfence_block();
        __syncwarp();
        const uint32_t s1 = sh_ld_u32_uniform(fe.a_ctl + 4u), s2 = sh_ld_u32_uniform(fe.a_ctl + 8u);
        bool collided = false;
        if (gid >= 0) {
            const uint32_t slot = dir_slot(kw, nm, s1, s2);
            if (sh_cas_u32(fe.a_meta + slot * 4u, 0u, META_BUSY) != 0u) collided = true;
            else {
#pragma unroll
                for (int k = 0; k < NKW; k++) sh_st_u64(fe.a_keys + slot * (8u * NKW) + 8u * k, kw[k]);
                __threadfence_block();
                sh_st_u32(fe.a_meta + slot * 4u, (uint32_t)(gid + 1) | (nm << 8));
            }
        }
        if (!__any_sync(0xffffffffu, collided)) {
            __threadfence_block();
            __syncwarp();
            if (lane == 0) ctl->count = (uint32_t)(n + add);
        } else {
            // some home slot was taken: new multipliers until every key of the directory (old and new) sits in its home slot
            if (lane == 0) ctl->gen = ctl->gen + 1u;          // odd: probes in flight are void
            __threadfence_block();
            __syncwarp();
            bool placed = false;
            uint32_t t1 = s1, t2 = s2;
            for (int a = 0; a < REBUILD_ATTEMPTS && !placed; a++) {
                t1 = (t1 * 0x2C1B3C6Du + 0x297A2D39u) | 1u;
                t2 = ((t2 ^ (t1 >> 7)) * 0x9E3779B1u + 0x85EBCA6Bu) | 1u;
                placed = dir_place_all(fe, n + add, t1, t2, lane);
                __syncwarp();
            }
            int keep = n + add;
            if (!placed) {
                // keys no multiplier separates (never seen; possible in principle): keep the old directory, which places its n
                // keys, stop inserting, and send the batch to the global table
                dir_place_all(fe, n, s1, s2, lane);
                t1 = s1; t2 = s2; keep = n; gid = -1;
                if (lane == 0) *reinterpret_cast<volatile uint32_t*>(fe.limit) = (uint32_t)n;
            }
            __threadfence_block();
            __syncwarp();
            if (lane == 0) { ctl->s1 = t1; ctl->s2 = t2; ctl->count = (uint32_t)keep; __threadfence_block(); ctl->gen = ctl->gen + 1u; }
        }
    }
    __threadfence_block();
    __syncwarp();
    gid = __shfl_sync(0xffffffffu, gid, head);       // the leader's answer for every lane of its key
    return has ? gid : -1;
}

// ---- the exact MIN/MAX path (rare) --------------------------------------------------------------------------------------
// Sentinels of the CTA-shared extremes (order-mapped): the identity = "no non-null value yet"; a second NaN-patterned value
// = "only NaNs so far" (any real value replaces it; while a group holds it the refreshed bound is a NaN, so every row
// takes the exact path — correct, slow, and only for as long as a group has seen nothing but NaNs).
__device__ __forceinline__ constexpr uint64_t mm_nan_mark(int m) { return ((Q::MM_ISMIN >> m) & 1u) ? ~0ULL - 1ULL : 1ULL; }

// ---- bounds without a lock ----------------------------------------------------------------------------------------------
// A published bound is valid for a row only if it was computed over EVERY group the row may belong to. Two counters frame
// the "first value of a group" events (started before the value is written, finished after), and the bound carries a
// token = (started, groups in the directory) taken while no event was in flight and re-checked after the scan. A reader
// trusts the bound only while the live token still equals the published one: a group inserted or valued since then
// changes the token, and the reader's rows simply take the exact path until the next refresh.
constexpr uint32_t TOKEN_NONE = 0xFFFFFFFFu;
__device__ __forceinline__ uint32_t bound_token(uint32_t started, uint32_t count) { return ((started & 0xFFFFFFu) << 8) | (count & 0xFFu); }

// The exact update of one MIN/MAX slot: shared-memory compare-and-swap loops on the order-mapped values. A NaN never
// replaces a held value (Main.kt:552: `value > this.value`); meeting the identity it leaves the NaN mark. Out of line and
// with scalar arguments only (nothing of the caller's tile goes through local memory): every beaten bound comes here —
// a few rows in ten thousand once the bounds are tight — and the per-tile loop only holds the calls.
static __device__ __noinline__ void fe_exact_slot(uint32_t a, uint32_t a_started, uint32_t a_finished, uint64_t x, bool ismin, bool isnan) {
    const uint64_t identity = ismin ? ~0ULL : 0ULL;
    const uint64_t cur = sh_ld_u64(a);
    const bool first = cur == identity;
    if (first) { asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a_started) : "memory"); __threadfence_block(); }
    if (isnan) { if (first) sh_cas_u64(a, identity, ismin ? ~0ULL - 1ULL : 1ULL); }
    else if (ismin ? x < cur : x > cur) { if (ismin) sh_min_u64(a, x); else sh_max_u64(a, x); }
    if (first) { __threadfence_block(); asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a_finished) : "memory"); }
}
// One row's values (v[i], bit i of okmask = input i is non-null) into the extremes of group gid.
__device__ __forceinline__ void fe_exact_vals(const Fe& fe, uint32_t gid, const uint64_t (&v)[Q::NIN > 0 ? Q::NIN : 1], uint32_t okmask) {
#pragma unroll
    for (int i = 0; i < Q::NIN; i++) {
        const int FL = Q::IN_FLAGS[i];
        if (!(FL & (F_MIN | F_MAX))) continue;
        if (Q::IN_CNT[i] > 0 && !((okmask >> i) & 1u)) continue;
        const bool is_int = (FL & F_INT) != 0;
        const bool isnan = !is_int && as_f64(v[i]) != as_f64(v[i]);
        const uint64_t x = order_map(v[i], is_int);
        // The bounds hold for ALL groups, so most rows that beat one do not beat their own group's extreme (with 50 groups ~4 in
        // 5): one look at the group's current value settles those here, and only a real improvement (or a first value) pays
        // for the call and the compare-and-swap.
        if (FL & F_MIN) {
            const uint32_t a = fe.a_mm + (gid * (uint32_t)NMM1 + (uint32_t)(Q::FE_MIN[i] < 0 ? 0 : Q::FE_MIN[i])) * 8u;
            const uint64_t cur = sh_ld_u64(a);
            if (cur == ~0ULL || (!isnan && x < cur)) fe_exact_slot(a, fe.a_started, fe.a_finished, x, true, isnan);
        }
        if (FL & F_MAX) {
            const uint32_t a = fe.a_mm + (gid * (uint32_t)NMM1 + (uint32_t)(Q::FE_MAX[i] < 0 ? 0 : Q::FE_MAX[i])) * 8u;
            const uint64_t cur = sh_ld_u64(a);
            if (cur == 0ULL || (!isnan && x > cur)) fe_exact_slot(a, fe.a_started, fe.a_finished, x, false, isnan);
        }
    }
}
// All exact rows of a lane's tile (`rows`: R-bit mask; gid[r] valid for those). Per lane: nothing here synchronises the warp.
// One row at a time through a select chain, so that the update code exists once.
__device__ __forceinline__ void fe_exact_rows(const Fe& fe, const uint32_t (&gid)[R], uint32_t rows, const AggSink& sink) {
#pragma unroll 1
    for (uint32_t todo = rows; todo; todo &= todo - 1u) {
        const int r = __ffs(todo) - 1;
        uint64_t v[Q::NIN > 0 ? Q::NIN : 1];
        uint32_t g = 0, okmask = 0;
#pragma unroll
        for (int q = 0; q < R; q++) if (q == r) {
            g = gid[q];
#pragma unroll
            for (int i = 0; i < Q::NIN; i++) { v[i] = sink.in[i][q]; okmask |= ((sink.inok[i] >> q) & 1u) << i; }
        }
        fe_exact_vals(fe, g, v, okmask);
    }
}

// This lane's view of the bounds for the tile it is about to accumulate: the published values if their token is live,
// else "no bound" (every row exact).
__device__ __forceinline__ void bounds_read(const Fe& fe, uint64_t (&bnd)[NMM1]) {
    const uint32_t t1 = sh_ld_u32(fe.a_btoken);
#pragma unroll
    for (int m = 0; m < Q::NMM; m++) bnd[m] = sh_ld_u64(fe.a_bound + 8u * m);
    const uint32_t t2 = sh_ld_u32(fe.a_btoken);
    const uint32_t live = bound_token(sh_ld_u32(fe.a_started), sh_ld_u32(fe.a_ctl + 12u));
    if (KQ_UNLIKELY(t1 != t2 || t1 != live)) {
#pragma unroll
        for (int m = 0; m < Q::NMM; m++) bnd[m] = bound_none(m, mm_is_int(m));
    }
}

// Recompute the bounds (one warp at a time, whoever gets the flag; nobody waits for it).
static __device__ __noinline__ void mm_bound_refresh(Fe fe, int lane) {
    uint32_t got = 1u;
    if (lane == 0) got = sh_cas_u32(fe.a_refreshing, 0u, 1u);
    if (__shfl_sync(0xffffffffu, got, 0) != 0u) return;
    const uint32_t fin = sh_ld_u32_uniform(fe.a_finished), sta = sh_ld_u32_uniform(fe.a_started);
    const int n = (int)sh_ld_u32_uniform(fe.a_ctl + 12u);
    bool ok = fin == sta;                         // no first value in flight
    uint64_t bv[NMM1];
#pragma unroll
    for (int m = 0; m < Q::NMM; m++) {
        const bool ismin = (Q::MM_ISMIN >> m) & 1u, is_int = mm_is_int(m);
        uint64_t b = ismin ? 0ULL : ~0ULL;               // MIN: the largest group minimum; MAX: the smallest group maximum (order-mapped)
        bool none = false;
        for (int g = lane; g < n; g += 32) {
            const uint64_t x = sh_ld_u64(fe.a_mm + (uint32_t)(g * NMM1 + m) * 8u);
            // a group without a value: with a nullable input its rows are caught by the first-row rule (fe_accumulate_row),
            // so it is skipped; otherwise (its first value is on its way, or Int64 where the identity is also a value) no bound
            if (x == mm_identity(m)) { if (is_int || !Q::MM_NULLABLE[m]) none = true; continue; }
            b = ismin ? (x > b ? x : b) : (x < b ? x : b);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const uint64_t y = __shfl_xor_sync(0xffffffffu, b, o);
            b = ismin ? (y > b ? y : b) : (y < b ? y : b);
        }
        none = __any_sync(0xffffffffu, none) || n == 0;
        bv[m] = none ? bound_none(m, is_int) : order_unmap(b, is_int);
    }
    // nothing started, nobody joined the directory while we looked? then the bounds cover every group there is
    ok = ok && sh_ld_u32_uniform(fe.a_started) == sta && (int)sh_ld_u32_uniform(fe.a_ctl + 12u) == n;
    if (lane == 0) {
        if (ok) {
            sh_st_u32(fe.a_btoken, TOKEN_NONE);
            __threadfence_block();
#pragma unroll
            for (int m = 0; m < Q::NMM; m++) sh_st_u64(fe.a_bound + 8u * m, bv[m]);
            __threadfence_block();
            sh_st_u32(fe.a_btoken, bound_token(sta, (uint32_t)n));
        }
        __threadfence_block();
        sh_exch_u32(fe.a_refreshing, 0u);
    }
}

// Accumulate one row into the lane-private slots of group g (FG = trash) — the branch-free per-row path. All loads of
// the row's slots are issued before the first store (the slots of one row never alias; the next row's may).
// Returns true when the row needs the exact MIN/MAX path (beats a bound, or first value of this lane in the group).
__device__ __forceinline__ bool fe_accumulate_row(const Fe& fe, uint32_t g, const AggSink& sink, int r, const uint64_t (&bnd)[NMM1]) {
    constexpr int NI = Q::NIN > 0 ? Q::NIN : 1;
    const uint32_t base = g * (uint32_t)GS;
    const uint32_t a8 = fe.a_lane8 + base, a4 = fe.a_lane4 + base;
    const uint32_t t8 = fe.a_lane8 + (uint32_t)(FG * GS), t4 = fe.a_lane4 + (uint32_t)(FG * GS);      // an invalid row of a nullable input: the trash group
    uint32_t c0 = 0, ca[NI], cv[NI], sa[NI];
    uint64_t sv[NI];
    bool valid[NI];
    if constexpr (Q::CNT0_USED) c0 = lds_u32(a4);
#pragma unroll
    for (int i = 0; i < Q::NIN; i++) {
        const int FL = Q::IN_FLAGS[i];
        const bool nullable = Q::IN_CNT[i] > 0;
        valid[i] = nullable ? ((sink.inok[i] >> r) & 1u) != 0 : true;
        if (nullable) { ca[i] = (valid[i] ? a4 : t4) + (uint32_t)Q::IN_CNT[i] * 128u; cv[i] = lds_u32(ca[i]); }
        if (FL & F_SUM) { sa[i] = (nullable ? (valid[i] ? a8 : t8) : a8) + (uint32_t)Q::FE_SUM[i] * 256u; sv[i] = lds_u64(sa[i]); }
    }
    if constexpr (Q::CNT0_USED) asm volatile("st.shared.u32 [%0], %1;" ::"r"(a4), "r"(c0 + 1u) : "memory");
    bool exact = false;
#pragma unroll
    for (int i = 0; i < Q::NIN; i++) {
        const int FL = Q::IN_FLAGS[i];
        const bool nullable = Q::IN_CNT[i] > 0;
        // first value this lane brings to the group: only a nullable input needs the rule (a group all of whose values were
        // null so far is skipped by the bounds); with a non-nullable input every group has a value from its first row on
        bool first = false;
        if (nullable) {
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(ca[i]), "r"(cv[i] + 1u) : "memory");
            first = cv[i] == 0u;
        }
        const uint64_t v = sink.in[i][r];
        if (FL & F_SUM) {
            const uint64_t y = (FL & F_INT) ? sv[i] + v : as_u64(__dadd_rn(as_f64(sv[i]), as_f64(v)));
            asm volatile("st.shared.u64 [%0], %1;" ::"r"(sa[i]), "l"(y) : "memory");
        }
        if (FL & (F_MIN | F_MAX)) {
            bool e = first;
            if (FL & F_INT) {
                if (FL & F_MIN) e |= (long long)v < (long long)bnd[Q::FE_MIN[i] < 0 ? 0 : Q::FE_MIN[i]];
                if (FL & F_MAX) e |= (long long)v > (long long)bnd[Q::FE_MAX[i] < 0 ? 0 : Q::FE_MAX[i]];
            } else {
                if (FL & F_MIN) e |= !(as_f64(v) >= as_f64(bnd[Q::FE_MIN[i] < 0 ? 0 : Q::FE_MIN[i]]));
                if (FL & F_MAX) e |= !(as_f64(v) <= as_f64(bnd[Q::FE_MAX[i] < 0 ? 0 : Q::FE_MAX[i]]));
            }
            e = e && valid[i];
            // Second look, only for a row that beat a bound (a few per thousand: the bounds hold for ALL groups, so with 50 groups
            // four in five of those rows do not beat their own group's extreme): the group's current values decide whether
            // the row leaves the tile loop for the exact path at all. A first value (identity) and the NaN mark always do.
            if (KQ_UNLIKELY(e) && !first) {
                const bool is_int = (FL & F_INT) != 0;
                const bool isnan = !is_int && as_f64(v) != as_f64(v);
                const uint64_t x = order_map(v, is_int);
                bool need = false;
                if (FL & F_MIN) {
                    const uint64_t cur = sh_ld_u64(fe.a_mm + (g * (uint32_t)NMM1 + (uint32_t)(Q::FE_MIN[i] < 0 ? 0 : Q::FE_MIN[i])) * 8u);
                    need |= cur == ~0ULL || (!isnan && x < cur);
                }
                if (FL & F_MAX) {
                    const uint64_t cur = sh_ld_u64(fe.a_mm + (g * (uint32_t)NMM1 + (uint32_t)(Q::FE_MAX[i] < 0 ? 0 : Q::FE_MAX[i])) * 8u);
                    need |= cur == 0ULL || (!isnan && x > cur);
                }
                e = need;
            }
            exact |= e;
        }
    }
    return exact;
}

template <int I>
__device__ __forceinline__ void global_accumulate_all(const AggArgs& A, uint64_t* rec, const AggSink& sink, int r) {
    if constexpr (I < Q::NIN) {
        if ((sink.inok[I] >> r) & 1u) global_accumulate(rec, A.in[I], sink.in[I][r]);
        global_accumulate_all<I + 1>(A, rec, sink, r);
    }
}

// Merge the lane-private slots of group g — the copies of ALL consumer warps, summed per lane first — into its global
// record: one atomic per group and slot for the whole CTA (every CTA of the grid updates the same few records at exit; with
// one atomic per warp those same-address updates were a quarter of a small batch's kernel time).
template <int I>
__device__ __forceinline__ void fe_merge_input(uint32_t a_blocks, uint64_t* rec, int g, int lane, const unsigned long long (&c)[Q::NCNT]) {
    if constexpr (I < Q::NIN) {
        constexpr int FL = Q::IN_FLAGS[I];
        constexpr uint32_t WB = (uint32_t)((FG + 1) * GS);      // bytes between the blocks of consecutive warps
        const unsigned long long n = c[Q::IN_CNT[I]];
        if (n != 0) {                                       // the CTA saw no non-null value of input I in group g otherwise
            if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(rec + Q::REC_NN[I]), n);
            if constexpr ((FL & F_SUM) != 0) {
                const uint32_t a = a_blocks + (uint32_t)g * GS + (uint32_t)Q::FE_SUM[I] * 256u + (uint32_t)lane * 8u;
                if constexpr ((FL & F_INT) != 0) {
                    uint64_t x = 0;
#pragma unroll
                    for (int w = 0; w < WARPS; w++) x += lds_u64(a + (uint32_t)w * WB);
#pragma unroll
                    for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
                    if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(rec + Q::REC_SUM[I]), (unsigned long long)x);
                } else {
                    double f = 0.0;
#pragma unroll
                    for (int w = 0; w < WARPS; w++) f = __dadd_rn(f, as_f64(lds_u64(a + (uint32_t)w * WB)));
#pragma unroll
                    for (int o = 16; o; o >>= 1) f = __dadd_rn(f, __shfl_xor_sync(0xffffffffu, f, o));
                    if (lane == 0) atomicAdd(reinterpret_cast<double*>(rec + Q::REC_SUM[I]), f);
                }
            }
        }
        fe_merge_input<I + 1>(a_blocks, rec, g, lane, c);
    }
}

// ---- general path: keys that are not in the directory (yet) --------------------------------------------------------------
// Called by the whole warp when some lane has rows (`slow`) whose probe missed. Looks again (another warp may have
// inserted the key), inserts missing keys one at a time under the CTA lock, and sends what the directory cannot take to
// the global table right away. Rows that ended up in the directory come back as group ids: the caller runs them through
// its one accumulate path a second time. Out of line, every argument by value — the caller keeps its registers, and the
// few thousand instructions of inserts, rebuilds and global-table probing stay out of the per-tile loop.
struct NullMasks { uint32_t m[R]; };
struct Resolved {
    uint32_t g[R];             // group of each row resolved to the directory, FG otherwise
    uint32_t new_groups;       // records this lane created in the global table
    int hits;                  // rows of this lane resolved to the directory
};
static __device__ __noinline__ Resolved fe_resolve_slow(const AggArgs* Ap, Fe fe, AggSink sink, NullMasks nms, uint32_t slow, bool bypass, int lane) {
    const AggArgs& A = *Ap;
    Resolved out;
#pragma unroll
    for (int r = 0; r < R; r++) out.g[r] = (uint32_t)FG;
    out.new_groups = 0; out.hits = 0;
    int rounds = 0;
    bool have_lock = false;           // taken before the first insert and kept until every key of this tile is placed
    bool fresh_view = false;          // lock held and probed under it: whatever is still in `slow` is known to be absent from the directory
    while (__any_sync(0xffffffffu, slow != 0)) {
        if (++rounds > 66 * R + 4096) { if (lane == 0) atomicOr(A.err, ERR_SPIN_SLOW); break; }      // each round resolves a row or inserts a key
        KQ_FTRACE(0x200000 + rounds * 256 + (slow & 0xff));
        bool full_dir = true;
        if (!bypass && !fresh_view) {
            // look again: another warp may have inserted the key meanwhile (no lock needed for that)
            const uint4 c = ctl_snapshot(fe);
            int g2[R];
#pragma unroll 1
            for (int r = 0; r < R; r++) {
                g2[r] = -1;
                if ((slow >> r) & 1u) {
                    uint64_t kw[NKW];
#pragma unroll
                    for (int k2 = 0; k2 < NKW; k2++) kw[k2] = sink.key[k2][r];
                    g2[r] = dir_probe<true>(fe, kw, nms.m[r], c.y, c.z);
                }
            }
            const uint32_t gen2 = sh_ld_u32_uniform(fe.a_ctl);
            const bool valid = gen2 == c.x && !(c.x & 1u);
            full_dir = valid && c.w >= sh_ld_u32_uniform(smem_u32(fe.limit));
            if (valid) {
#pragma unroll 1
                for (int r = 0; r < R; r++) {
                    if (g2[r] < 0) continue;
                    out.g[r] = (uint32_t)g2[r];          // group ids never change once given
                    slow &= ~(1u << r);
                    out.hits++;
                }
            }
            if (!__any_sync(0xffffffffu, slow != 0)) break;
            fresh_view = have_lock && valid;      // under the lock nobody else changes the directory
        } else if (!bypass) {
            full_dir = sh_ld_u32_uniform(fe.a_ctl + 12u) >= sh_ld_u32_uniform(smem_u32(fe.limit));
        }
        if (full_dir) {
            // the directory takes no more keys: the rest goes to the global table
#pragma unroll 1
            for (int r = 0; r < R; r++) {
                if (!((slow >> r) & 1u)) continue;
                uint64_t kw[MAX_KEYS];
#pragma unroll
                for (int k2 = 0; k2 < MAX_KEYS; k2++) kw[k2] = k2 < Q::NKEYS ? sink.key[k2][r] : 0;
                uint64_t* rec = table_find_or_insert(A, hash_key(kw, nms.m[r], Q::NKEYS), kw, nms.m[r], &out.new_groups);
                if (rec) global_accumulate_all<0>(A, rec, sink, r);
            }
            slow = 0;
            break;
        }
        // Insert under the CTA lock, a batch per round: the probes are repeated under the lock (only this warp can change the
        // directory now), then every lane proposes the key of its first unresolved row and dir_insert_batch takes them all.
        if (!have_lock) { fe_lock(fe, lane); have_lock = true; continue; }
        if (!fresh_view) continue;                 // a rebuild was in flight during the probe under the lock (cannot be: ours are done) — look again
        // every lane proposes the key of its first unresolved row
        const bool has = slow != 0;
        const int r0 = has ? __ffs(slow) - 1 : 0;
        uint64_t kw[NKW];
#pragma unroll
        for (int k2 = 0; k2 < NKW; k2++) kw[k2] = sink.key[k2][r0];
        const uint32_t knm = nms.m[r0];
        KQ_FTRACE(0x300000 + rounds * 256);
        const int g = dir_insert_batch(fe, has, kw, knm, lane);
        KQ_FTRACE(0x400000 + rounds * 256 + (g & 0xff));
        if (has) {
            if (g >= 0) {
#pragma unroll 1
                for (int r = 0; r < R; r++) {
                    if (!((slow >> r) & 1u)) continue;
                    bool eq = nms.m[r] == knm;
#pragma unroll
                    for (int k2 = 0; k2 < NKW; k2++) eq &= sink.key[k2][r] == kw[k2];
                    if (eq) { out.g[r] = (uint32_t)g; slow &= ~(1u << r); out.hits++; }
                }
            } else {
                // not insertable (directory full or unplaceable): this row goes to the global table now
                uint64_t kg[MAX_KEYS];
#pragma unroll
                for (int k2 = 0; k2 < MAX_KEYS; k2++) kg[k2] = k2 < Q::NKEYS ? kw[k2] : 0;
                uint64_t* rec = table_find_or_insert(A, hash_key(kg, knm, Q::NKEYS), kg, knm, &out.new_groups);
                if (rec) global_accumulate_all<0>(A, rec, sink, r0);
                slow &= ~(1u << r0);
            }
        }
        fresh_view = false;            // other lanes' keys joined the directory: the next round looks again (still under the lock)
    }
    if (have_lock) fe_unlock(fe, lane);
    return out;
}

// debugging: the last checkpoint every warp of blocks 0..15 reached, written to pinned host memory (readable while the kernel hangs)
#ifdef KQ_FE_TRACE
#define KQ_TRACE(code) do { if (A.trace && blockIdx.x < 16 && (threadIdx.x & 31) == 0) reinterpret_cast<volatile unsigned long long*>(A.trace)[blockIdx.x * 16 + (threadIdx.x >> 5)] = (unsigned long long)(code); } while (0)
#else
#define KQ_TRACE(code) do { } while (0)
#endif

#ifndef KQ_CTAS
#define KQ_CTAS 1
#endif
extern "C" __global__ void __launch_bounds__(THREADS, KQ_CTAS) kq_group_aggregate(const __grid_constant__ AggArgs A) {
    // dynamic shared memory: [S stages][directory state][directory keys][dense keys][dense null masks][extremes][gslot]
    //                        [per-warp lane-private blocks of FG + 1 groups]
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[S], empty[S];
    __shared__ long long tile_of[S];
    __shared__ long long bbase[S][MAX_COLS];
    __shared__ DirCtl s_ctl;
    __shared__ uint32_t s_lock, s_limit, s_btoken, s_started, s_finished, s_refreshing, s_done;
    __shared__ uint64_t s_bound[NMM1];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warp = wid - NSERVICE;          // consumer warp index
    unsigned char* p0 = smem + (size_t)S * A.sp.stage_bytes;
    const size_t fe_begin = (size_t)(p0 - smem);
    Fe fe;
    fe.meta = reinterpret_cast<uint32_t*>(p0);            p0 += (size_t)DIR * 4;
    fe.keys = reinterpret_cast<uint64_t*>(p0);            p0 += (size_t)DIR * NKW * 8;
    fe.gkeys = reinterpret_cast<uint64_t*>(p0);           p0 += (size_t)FG * NKW * 8;
    fe.mm = reinterpret_cast<uint64_t*>(p0);              p0 += (size_t)(FG + 1) * NMM1 * 8;
    uint64_t* gslot = reinterpret_cast<uint64_t*>(p0);    p0 += (size_t)FG * 8;
    fe.gnm = reinterpret_cast<uint32_t*>(p0);             p0 += (size_t)((FG + 3) & ~3) * 4;
    unsigned char* lane_blocks = p0;                      p0 += (size_t)WARPS * (FG + 1) * GS;
    const size_t fe_end = (size_t)(p0 - smem);
    fe.ctl = &s_ctl; fe.limit = &s_limit; fe.err = A.err; fe.trace = A.trace;
    fe.a_ctl = smem_u32(&s_ctl); fe.a_meta = smem_u32(fe.meta); fe.a_keys = smem_u32(fe.keys); fe.a_mm = smem_u32(fe.mm);
    fe.a_lock = smem_u32(&s_lock); fe.a_bound = smem_u32(s_bound);
    fe.a_btoken = smem_u32(&s_btoken); fe.a_started = smem_u32(&s_started); fe.a_finished = smem_u32(&s_finished); fe.a_refreshing = smem_u32(&s_refreshing);
    const uint32_t a_warp = smem_u32(lane_blocks) + (uint32_t)(warp < 0 ? 0 : warp) * (uint32_t)((FG + 1) * GS);
    fe.a_lane8 = a_warp + (uint32_t)lane * 8u;
    fe.a_lane4 = a_warp + (uint32_t)Q::NSUM * 256u + (uint32_t)lane * 4u;

    for (size_t i = fe_begin + threadIdx.x * 4; i < fe_end; i += THREADS * 4) *reinterpret_cast<uint32_t*>(smem + i) = 0;
    if (threadIdx.x == 0) {
        s_ctl.gen = 0; s_ctl.s1 = 0x9E3779B1u; s_ctl.s2 = 0x85EBCA6Bu; s_ctl.count = 0;
        s_lock = 0; s_limit = (uint32_t)FG; s_btoken = TOKEN_NONE; s_started = 0; s_finished = 0; s_refreshing = 0; s_done = 0;
        for (int m = 0; m < Q::NMM; m++) s_bound[m] = bound_none(m, mm_is_int(m));
        for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], WARPS); }
        mbar_fence_init();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (FG + 1) * Q::NMM; i += THREADS) fe.mm[i] = mm_identity(i % NMM1);
    __syncthreads();

    if (wid == PRODUCER_WARP) {
        // Tiles are dealt statically: block b owns tiles b, b + grid, b + 2 grid, ... — no ticket atomic on the critical path.
        // A block that stops early (global table past its threshold) records how far it got; the host grows the table and
        // relaunches with the same grid, every block resuming its own sequence.
        //
        // The whole producer WARP issues a tile, one lane per staged buffer: each lane keeps its buffer's descriptor in
        // registers, computes its own byte count and source address, the counts are summed with one warp reduction for the
        // barrier's expect_tx and every lane issues its own bulk copy. (One lane walking all buffers took ~600 dependent
        // instructions = 1.3 us per tile; with a two-stage ring that sat between a stage being freed and its next data being
        // requested, and consumers spent 17 % of their time waiting at the full barrier.) A lane that owns a Utf8 byte
        // buffer reads the two boundary offsets of its column for the tiles PF ahead straight from HBM, so string bytes and
        // fixed-size buffers of a tile go out together on one barrier the moment a stage is free.
        constexpr int PF = 4;
        const long long first = A.progress ? (long long)A.progress[blockIdx.x] : 0;
        auto tile_at = [&](long long j) -> long long { const long long t = A.tile_begin + (long long)blockIdx.x + j * (long long)gridDim.x; return t < A.ntiles ? t : -1; };
        // this lane's buffer
        const bool mine = lane < A.sp.nbuf;
        const StageBuf sb = A.sp.buf[mine ? lane : 0];
        bool is_bytes = mine && sb.kind == SK_BYTES, staged_bytes = false;
        for (int i = 0; i < MAX_BYTES_BUFS; i++) staged_bytes |= is_bytes && i < A.sp.nbytes && A.sp.bytes_buf[i] == lane;
        const int32_t* offp = is_bytes ? reinterpret_cast<const int32_t*>(A.sp.buf[sb.aux & 0xffff].g) : nullptr;
        const int bslot = sb.aux >> 16;
        int32_t blo[PF], bhi[PF];
        long long tq[PF];
        auto bounds_fetch = [&](long long tile, int32_t& lo, int32_t& hi) {
            lo = hi = 0;
            if (staged_bytes && tile >= 0) {
                const long long row0 = tile * TILE;
                const int rows = (int)((A.n - row0) < TILE ? (A.n - row0) : TILE);
                asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(lo) : "l"(offp + row0));
                asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(hi) : "l"(offp + row0 + rows));
            }
        };
#pragma unroll
        for (int u = 0; u < PF; u++) { tq[u] = tile_at(first + u); bounds_fetch(tq[u], blo[u], bhi[u]); }
        long long issued = 0;
        bool more = true;
        // the group count that stops a throttled launch is read one tile ahead (an L2 round trip otherwise in front of every tile)
        const bool throttled = A.stop_threshold != ~0ULL;
        unsigned long long ng = throttled ? *reinterpret_cast<volatile unsigned long long*>(A.ngroups) : 0ULL;
        for (int kp0 = 0; more; kp0 += PF) {
#pragma unroll
            for (int u = 0; u < PF; u++) {
                if (!more) break;
                const int kp = kp0 + u, s = kp % S;
                { int spins = 0; while (!mbar_test(&empty[s], ((kp / S) & 1) ^ 1)) { __nanosleep(32); if (++spins > SPIN_LIMIT) { if (lane == 0) atomicOr(A.err, ERR_SPIN_STAGE); break; } } }
                long long tile = tq[u];
                if (tile >= 0 && throttled && ng > A.stop_threshold) tile = -1;
                if (tile < 0) {
                    if (lane == 0) { tile_of[s] = tile; sh_st_u32(smem_u32(&s_done), 1u); mbar_arrive(&full[s]); }
                    more = false;
                    break;
                }
                // this lane's share of the tile
                const long long row0 = tile * TILE;
                const int rows = (int)((A.n - row0) < TILE ? (A.n - row0) : TILE);
                uint32_t bytes = 0;
                const char* src = sb.g;
                if (mine && !is_bytes) {
                    long long goff;
                    if (sb.kind == SK_W8) { bytes = (uint32_t)rows * 8u; goff = row0 * 8; }
                    else if (sb.kind == SK_W4) { bytes = (uint32_t)rows * 4u; goff = row0 * 4; }
                    else if (sb.kind == SK_W4_PLUS1) { bytes = (uint32_t)(rows + 1) * 4u; goff = row0 * 4; }
                    else { bytes = (uint32_t)(rows + 7) / 8u; goff = row0 / 8; }
                    bytes = (bytes + 15u) & ~15u;
                    src = sb.g + goff;
                } else if (staged_bytes) {
                    const long long lo = (long long)blo[u] & ~15LL, hi = ((long long)bhi[u] + 15LL) & ~15LL;
                    const bool fits = hi - lo + 16 <= (long long)sb.cap;
                    if (fits) { bytes = (uint32_t)(hi - lo + 16); src = sb.g + lo; }       // +16: consumers read whole 8-byte words past the end (buffers are padded)
                    bbase[s][bslot] = fits ? lo : -1LL;
                }
                const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);
                __syncwarp();                                  // the byte lanes' bbase stores before lane 0's release
                if (lane == 0) { tile_of[s] = tile; mbar_arrive_expect_tx(&full[s], total); }
                __syncwarp();
                if (bytes) bulk_g2s(smem + (size_t)s * A.sp.stage_bytes + sb.soff, src, bytes, &full[s]);
                issued++;
                if (throttled) ng = *reinterpret_cast<volatile unsigned long long*>(A.ngroups);
                tq[u] = tile_at(first + kp + PF);
                bounds_fetch(tq[u], blo[u], bhi[u]);
            }
        }
        if (lane == 0) {
            if (A.progress) A.progress[blockIdx.x] = (unsigned int)(first + issued);
            atomicAdd(A.ticket, (unsigned int)issued);           // tiles finished by this launch (the host compares the running total with the batch)
        }
    } else if (NSERVICE == 2 && wid == 1) {
        // housekeeping: recompute the bounds of the extremes all the time (every ~0.3 us while the directory fills, every
        // ~1.5 us later); nobody waits for this warp
        uint32_t it = 0;
        while (sh_ld_u32_uniform(smem_u32(&s_done)) == 0u) {
            mm_bound_refresh(fe, lane);
            __nanosleep(it < 256 ? 200 : 1500);
            it++;
        }
    } else {
        AggSink sink;
        bool bypass = false;                      // high cardinality after all: stop probing a full directory that mostly misses
        int low_tiles = 0;
        for (int k = 0;; k++) {
            const int s = k % S;
            {
                int spins = 0;
                while (!mbar_try_wait(&full[s], (k / S) & 1)) { if (++spins > SPIN_LIMIT) { if (lane == 0) atomicOr(A.err, ERR_SPIN_STAGE); break; } }
                if (spins > SPIN_LIMIT) break;
            }
            const long long tile = tile_of[s];
            KQ_TRACE(0x100000 + k * 16 + 1);
            if (KQ_UNLIKELY(tile < 0)) break;
            RowCtx rc;
            rowctx_init(rc, warp, tile, TILE, A.n, A.err, smem + (size_t)s * A.sp.stage_bytes);
            rc.bbase = bbase[s];
            rc.heap = A.heap.tab ? &A.heap : nullptr;
            sink.sel = rc.inr;
            if constexpr (Q::NKEYS == 0) {
#pragma unroll
                for (int r = 0; r < R; r++) sink.key[0][r] = 0;
                sink.keyok[0] = RMASK;
            }
            Q::eval(A.q, rc, sink);
            stage_release(&empty[s], &tile_of[s], (uint32_t)tile, A.err, lane);       // everything needed is in registers now
            KQ_TRACE(0x100000 + k * 16 + 2);

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

            // ---- key -> group: one probe per row ------------------------------------------------------------------
            uint32_t gsel[R];                         // group id, FG (trash) for rows that are filtered out or unresolved
            uint32_t slow = sink.sel;                 // rows that still need the general path
            if (KQ_LIKELY(!bypass)) {
                // {gen, s1, s2, count}, per lane: nothing below that depends on it synchronises the warp (lanes that read different
                // generations simply disagree on which of their OWN rows missed)
                uint4 c;
                asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w) : "r"(fe.a_ctl) : "memory");
                uint32_t miss = 0;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    uint64_t kw[NKW];
#pragma unroll
                    for (int k2 = 0; k2 < NKW; k2++) kw[k2] = sink.key[k2][r];
                    const int g = dir_probe(fe, kw, nm[r], c.y, c.z);
                    const bool on = (sink.sel >> r) & 1u;
                    gsel[r] = (on && g >= 0) ? (uint32_t)g : (uint32_t)FG;
                    miss |= (uint32_t)(on && g < 0) << r;
                }
                const uint32_t gen2 = sh_ld_u32(fe.a_ctl);
                if (KQ_UNLIKELY(gen2 != c.x || (c.x & 1u))) {      // a rebuild ran meanwhile: nothing probed counts
#pragma unroll
                    for (int r = 0; r < R; r++) gsel[r] = (uint32_t)FG;
                    miss = sink.sel;
                }
                slow = miss;
            } else {
#pragma unroll
                for (int r = 0; r < R; r++) gsel[r] = (uint32_t)FG;
            }

#ifdef KQ_FE_CHECK
#pragma unroll
            for (int r = 0; r < R; r++) if (gsel[r] > (uint32_t)FG) { atomicOr(A.err, 0x1000u); gsel[r] = (uint32_t)FG; }
#endif
            // ---- accumulate (branch-free); bounds are read AFTER the probes: they cover every group probed -----------
            // One copy of the accumulate code serves both the rows whose probe hit (pass 1, nearly always the only one) and the
            // rows the general path resolved to a directory group (pass 2, without bounds: always the exact MIN/MAX compare).
            KQ_TRACE(0x100000 + k * 16 + 3);
            uint64_t bnd[NMM1];
#pragma unroll
            for (int m = 0; m < NMM1; m++) bnd[m] = 0;
            if (Q::NMM > 0) bounds_read(fe, bnd);
            bool went_slow = false;
            uint32_t new_groups = 0;
            const int rows = __popc(sink.sel);
            int fe_hits = rows - __popc(slow);
#pragma unroll 1
            for (;;) {
                if (KQ_LIKELY(!bypass || went_slow)) {
                    uint32_t exact = 0, onmask = 0;
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const bool on = gsel[r] != (uint32_t)FG;
                        onmask |= (uint32_t)on << r;
                        exact |= (uint32_t)(fe_accumulate_row(fe, gsel[r], sink, r, bnd) && on) << r;
                    }
#ifdef KQ_FE_NOEXACT
                    exact = 0;
#endif
                    if (Q::NMM > 0) {
                        // a lane whose row was the first value it brought to a group may have used a bound that does not cover the
                        // group: all its rows take the exact path (nullable inputs only; see fe_accumulate_row)
                        if (Q::ANY_MM_NULLABLE && exact) exact = onmask;
                        if (KQ_UNLIKELY(exact != 0)) fe_exact_rows(fe, gsel, exact, sink);
                    }
                }
                KQ_TRACE(0x100000 + k * 16 + 4);
                if (KQ_LIKELY(went_slow || !__any_sync(0xffffffffu, slow != 0))) break;
                NullMasks nms;
#pragma unroll
                for (int r = 0; r < R; r++) nms.m[r] = nm[r];
                const Resolved rs = fe_resolve_slow(&A, fe, sink, nms, slow, bypass, lane);
                went_slow = true;
                slow = 0;
                new_groups = rs.new_groups;
                fe_hits += rs.hits;
                if (!__any_sync(0xffffffffu, rs.hits != 0)) break;
#pragma unroll
                for (int r = 0; r < R; r++) gsel[r] = rs.g[r];
#pragma unroll
                for (int m = 0; m < Q::NMM; m++) bnd[m] = bound_none(m, mm_is_int(m));
            }
            KQ_TRACE(0x100000 + k * 16 + 5);
            // one update of the global group count per warp and tile (a single counter bumped by every insert serialises in the L2)
            if (KQ_UNLIKELY(went_slow) && __any_sync(0xffffffffu, new_groups != 0)) {
                const uint32_t tot = __reduce_add_sync(0xffffffffu, new_groups);
                if (lane == 0) atomicAdd(A.ngroups, (unsigned long long)tot);
            }
            // once the directory is full and this warp mostly misses it, stop probing it (the hint was wrong: high cardinality)
            // (only a warp that went through the general path can have missed the directory: `went_slow` is warp-uniform)
            if (KQ_UNLIKELY(!bypass && went_slow) && sh_ld_u32_uniform(fe.a_ctl + 12u) >= sh_ld_u32_uniform(smem_u32(&s_limit)) && __any_sync(0xffffffffu, fe_hits < rows)) {
                int hits = fe_hits, tot = rows;
#pragma unroll
                for (int o = 16; o; o >>= 1) { hits += __shfl_xor_sync(0xffffffffu, hits, o); tot += __shfl_xor_sync(0xffffffffu, tot, o); }
                if (tot >= 64 && hits * 8 < tot) { if (++low_tiles >= 2) bypass = true; }
                else low_tiles = 0;
            } else low_tiles = 0;
        }
    }

    // ---- merge the CTA's groups into the global table ---------------------------------------------------------------
    KQ_TRACE(0x500000);
    __syncthreads();
    KQ_TRACE(0x500001);
    int G = (int)s_ctl.count;
#ifdef KQ_FE_NOMERGE
    G = 0;
#endif
#ifdef KQ_FE_CHECK
    if (G > FG) { if (threadIdx.x == 0) atomicOr(A.err, 0x8000u); G = FG; }
#endif
    for (int g = threadIdx.x; g < G; g += THREADS) {
        uint64_t kw[MAX_KEYS];
#pragma unroll
        for (int k = 0; k < MAX_KEYS; k++) kw[k] = k < Q::NKEYS ? fe.gkeys[g * NKW + k] : 0;
        const uint32_t nullmask = fe.gnm[g];
        uint32_t fresh = 0;
        uint64_t* rec = table_find_or_insert(A, hash_key(kw, nullmask, Q::NKEYS), kw, nullmask, &fresh);
        if (fresh) atomicAdd(A.ngroups, 1ULL);
        gslot[g] = rec ? (uint64_t)(rec - A.table) : ~0ULL;
    }
    KQ_TRACE(0x500004);
    __syncthreads();
    KQ_TRACE(0x500003);
    if (warp >= 0) {
        const uint32_t a_blocks = smem_u32(lane_blocks);
        constexpr uint32_t WB = (uint32_t)((FG + 1) * GS);
        for (int g = warp; g < G; g += WARPS) {             // the groups are dealt to the warps; each sums all warps' copies
            if (gslot[g] == ~0ULL) continue;
            uint64_t* rec = A.table + gslot[g];
            unsigned long long c[Q::NCNT];
#pragma unroll
            for (int j = 0; j < Q::NCNT; j++) {
                unsigned long long x = 0;
#pragma unroll
                for (int w = 0; w < WARPS; w++) x += lds_u32(a_blocks + (uint32_t)w * WB + (uint32_t)g * GS + (uint32_t)Q::NSUM * 256u + (uint32_t)j * 128u + (uint32_t)lane * 4u);
#pragma unroll
                for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
                c[j] = x;
            }
            fe_merge_input<0>(a_blocks, rec, g, lane, c);
        }
    }
    KQ_TRACE(0x500002);
    for (int t = threadIdx.x; t < G * Q::NMM; t += THREADS) {
        const int g = t / NMM1, m = t % NMM1;
        if (gslot[g] == ~0ULL) continue;
        const uint64_t v = fe.mm[g * NMM1 + m];
        uint64_t* p = A.table + gslot[g] + Q::MM_WORD[m];
        if ((Q::MM_ISMIN >> m) & 1u) { if (v != ~0ULL) atomicMin(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v); }
        else if (v != 0ULL) atomicMax(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
    }
}

}  // namespace kq
