// kq_args.h — kernel argument blocks shared by the host library (nvcc) and the query kernels that
// are specialised at run time (NVRTC, see kq_jit.cu). Fixed-width integers and raw pointers only;
// the text of this header is embedded verbatim in every generated translation unit.
#pragma once

#ifndef __CUDACC_RTC__
#include <stdint.h>
#endif

namespace kq {

// kq_type values of include/kqgpu.h, usable in device code
constexpr int KQT_F64 = 1, KQT_UTF8 = 2, KQT_I64 = 3, KQT_BOOL = 4, KQT_DATE32 = 5, KQT_I32 = 6;

// comparison truth masks: bit0 = lt, bit1 = eq, bit2 = gt, bit3 = unordered (NaN)
constexpr uint32_t CM_EQ = 0x2, CM_NE = 0xD, CM_LT = 0x1, CM_LE = 0x3, CM_GT = 0x4, CM_GE = 0x6;

constexpr int MAX_COLS = 16;        // distinct input columns one kernel may touch
constexpr int MAX_LIT = 32;         // literal slots (64-bit each)
constexpr int LITPOOL = 256;        // bytes of Utf8 literal text
constexpr int MAX_OUT = 8;          // computed output columns of one kernel
constexpr int MAX_KEYS = 4;         // group-by expressions
constexpr int MAX_INPUTS = 6;       // distinct aggregate input expressions
constexpr int MAX_STAGE_BUFS = 24;
constexpr int MAX_STAGES = 8;
constexpr int MAX_BYTES_BUFS = 4;  // Utf8 byte ranges staged per tile

// One input column (an Arrow FieldVector triple resident in HBM).
struct QCol {
    const void* data;
    const uint32_t* validity;       // NULL: all rows valid
    const int32_t* offsets;         // Utf8 only
};

// Run-time operands of a specialised kernel. Everything that shapes the code (types, nullability,
// operators, staging offsets) is compiled in; pointers and literal VALUES stay arguments so that
// `a > 0.5` and `a > 0.7` share one kernel.
struct QArgs {
    QCol cols[MAX_COLS];
    uint64_t lit[MAX_LIT];          // f64 bits / int64 / date32 (sign-extended) / Utf8: (pool offset << 32) | length
    uint8_t pool[LITPOOL];
};

// Shared-memory stage plan: which column buffers the producer warp moves with TMA bulk copies.
enum StageKind : int32_t { SK_W8 = 0, SK_W4 = 1, SK_W4_PLUS1 = 2, SK_BIT = 3, SK_BYTES = 4 };

struct StageBuf {
    const char* g;                  // global base of the buffer
    int32_t soff;                   // byte offset inside a stage (128-byte aligned)
    int32_t kind;                   // StageKind
    int32_t aux;                    // SK_BYTES: index of the StageBuf holding this column's offsets; column slot in the high half
    int32_t cap;                    // SK_BYTES: bytes reserved per stage
    int32_t slot, role;             // which QArgs column buffer this is: role 0 = data, 1 = validity, 2 = offsets
};


struct StagePlan {
    int32_t nbuf;
    int32_t stage_bytes;            // multiple of 128
    int32_t nstages;
    int32_t nbytes;                 // SK_BYTES buffers (at most MAX_BYTES_BUFS are staged)
    int32_t bytes_buf[4];           // their indices in buf[]
    StageBuf buf[MAX_STAGE_BUFS];
};

// (re)bind the global bases of a stage plan to the column pointers of q
__host__ __device__ inline void stage_plan_bind(StagePlan& sp, const QArgs& q) {
    for (int b = 0; b < sp.nbuf; b++) {
        const QCol& c = q.cols[sp.buf[b].slot];
        sp.buf[b].g = (const char*)(sp.buf[b].role == 0 ? c.data : (sp.buf[b].role == 1 ? (const void*)c.validity : (const void*)c.offsets));
    }
}

struct DOut {
    void* data;
    uint32_t* validity;
};

// ProjectionExec / FilterExec kernels.
struct OpArgs {
    QArgs q;
    int64_t n, ntiles;
    int64_t tile_begin, tile_end;   // this launch covers tiles [tile_begin, tile_end) of the batch (host-streamed input is launched chunk by chunk)
    DOut outs[MAX_OUT];
    unsigned long long* tile_desc;  // decoupled look-back descriptors, one per tile
    unsigned int* ticket;
    unsigned long long* out_count;
    int32_t* selvec;
    unsigned long long* trace;      // debugging (KQ_TRACE builds): 8 timestamps per tile
    uint32_t* err;
    StagePlan sp;
};


// ---- HashAggregateExec -----------------------------------------------------------------------------------------
constexpr int MAX_REC_WORDS = 32;
constexpr int DIR_SLOTS = 256;
constexpr int FE_MAX_GROUPS = 64;

enum : int32_t { F_SUM = 1, F_MIN = 2, F_MAX = 4, F_INT = 8 };

// Utf8 group keys longer than 7 bytes (kq_rt.cuh utf8_intern): the key word of such a string is a 63-bit hash of its bytes
// with the top bit set; the bytes live once in a per-aggregate heap, found through an open-addressing table keyed by the
// word. Every row's bytes are compared with the heap copy, so two different strings with the same hash are an ERROR, never
// a silently merged group.
struct KeyHeap {
    unsigned long long* tab;        // [cap][2]: {word (0 empty, 1 being published), heap offset << 24 | length}; NULL: no long keys in this launch
    uint64_t cap_mask;
    uint8_t* bytes;
    uint64_t bytes_cap;
    unsigned long long* used;       // [0] entries, [1] heap bytes
};

__host__ __device__ inline uint64_t heap_home(uint64_t word, uint64_t cap_mask) { return ((word * 0x9E3779B97F4A7C15ULL) >> 20) & cap_mask; }
// (offset << 24 | length) of a long key word, 0 if absent (finalize)
__host__ __device__ inline unsigned long long heap_find(const KeyHeap& H, uint64_t word) {
    uint64_t slot = heap_home(word, H.cap_mask);
    for (uint64_t probes = 0; probes <= H.cap_mask; probes++) {
        const unsigned long long w = H.tab[2 * slot];
        if (w == word) return H.tab[2 * slot + 1];
        if (w == 0ULL) return 0ULL;
        slot = (slot + 1) & H.cap_mask;
    }
    return 0ULL;
}

#ifdef __CUDACC__
// ---- long Utf8 group keys ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t utf8_hash63(const uint8_t* p, int len) {
    uint64_t h = 0xCBF29CE484222325ULL ^ (uint64_t)(uint32_t)len;
    for (int i = 0; i < len; i++) h = (h ^ __ldg(p + i)) * 0x100000001B3ULL;        // FNV-1a over the bytes ...
    h ^= h >> 32; h *= 0xD6E8FEB86659FD93ULL; h ^= h >> 32;                          // ... and a finaliser
    return h | 0x8000000000000000ULL;                                                // top bit: "long key" (a packed short key has a length <= 7 there)
}
// Key word of the string p[0..len), len > 7: find it in the heap (comparing the bytes) or add it. Out of line: rare path.
static __device__ __noinline__ uint64_t utf8_intern(const KeyHeap* H, const uint8_t* p, int len, uint32_t* err) {
    const uint64_t word = utf8_hash63(p, len);
    uint64_t slot = heap_home(word, H->cap_mask);
    for (uint64_t probes = 0; probes <= H->cap_mask; probes++) {
        unsigned long long* e = H->tab + 2 * slot;
        unsigned long long w = *reinterpret_cast<volatile unsigned long long*>(e);
        if (w == 0ULL) {
            w = atomicCAS(e, 0ULL, 1ULL);
            if (w == 0ULL) {                                                         // ours: copy the bytes, then publish
                const unsigned long long off = atomicAdd(H->used + 1, (unsigned long long)((len + 7) & ~7));
                if (off + (unsigned long long)len > H->bytes_cap) { atomicOr(err, 32u); *reinterpret_cast<volatile unsigned long long*>(e + 1) = 0ULL; }
                else {
                    for (int i = 0; i < len; i++) H->bytes[off + i] = __ldg(p + i);
                    *reinterpret_cast<volatile unsigned long long*>(e + 1) = (off << 24) | (unsigned long long)(uint32_t)len;
                }
                atomicAdd(H->used, 1ULL);
                __threadfence();
                *reinterpret_cast<volatile unsigned long long*>(e) = word;
                return word;
            }
        }
        for (int spins = 0; w == 1ULL && spins < (1 << 20); spins++) w = *reinterpret_cast<volatile unsigned long long*>(e);      // being published
        if (w == word) {
            const unsigned long long meta = *reinterpret_cast<volatile unsigned long long*>(e + 1);
            bool same = (int)(meta & 0xFFFFFFu) == len;
            const uint8_t* q = H->bytes + (meta >> 24);
            for (int i = 0; same && i < len; i++) same = __ldcg(q + i) == __ldg(p + i);       // the heap is written by other SMs: not through this SM's L1
            if (!same) atomicOr(err, 16u);       // two different strings, one 63-bit hash: refuse rather than merge the groups
            return word;
        }
        slot = (slot + 1) & H->cap_mask;
    }
    atomicOr(err, 32u);
    return word;
}
#endif  // __CUDACC__

struct AggInput {
    int32_t flags;
    int32_t rec_nn, rec_sum, rec_min, rec_max;   // record word indices (-1 = absent)
    int32_t fe_sum, fe_min, fe_max;              // front-end slot indices (-1 = absent); the count slot is the input index
};

struct AggArgs {
    QArgs q;
    int64_t n, ntiles, tile_begin;
    int32_t nkeys, ninputs;
    uint32_t key_f64_mask;                 // keys whose NaNs must be canonicalised (Double.equals, rule R7)
    int32_t stride;                        // record stride in 64-bit words
    AggInput in[MAX_INPUTS];
    uint64_t rec_init[MAX_REC_WORDS];
    uint64_t* table;
    uint64_t cap_mask;
    unsigned long long* ngroups;
    unsigned long long stop_threshold;
    unsigned int* ticket;
    uint32_t* err;
    unsigned int* overflow;                // set when an insert finds the table full (optimistically sized table: the host rolls back and grows)
    unsigned int* progress;                // kq_k_agg_fe.cuh: tiles each block has finished in earlier launches over this batch (block b owns tiles b, b + grid, ...)
    KeyHeap heap;                          // long Utf8 group keys
    unsigned long long* trace;             // debugging (KQ_FE_PROGRESS=<pinned host address>): last checkpoint per warp of block 0
    // front end
    int32_t fe_groups, fe_nsum, fe_nmm;
    int32_t geo_r, geo_warps;              // tile geometry the host chose for this launch (host-side bookkeeping)
    int32_t geo_ctas, geo_service;         // CTAs per SM the launch is sized for; service warps per CTA (kq_group_aggregate)
    int32_t fe_sum_word[MAX_INPUTS];       // front-end sum slot -> record word
    uint32_t fe_sum_int;                   // bit s: slot s is an integer sum
    int32_t fe_mm_word[2 * MAX_INPUTS];    // front-end min/max slot -> record word
    uint32_t fe_mm_ismin;
    // partitioned path (high cardinality): pass 1 scatters the evaluated rows into per-(partition, CTA) buckets of
    // HBM scratch, pass 2 reduces one partition at a time in a shared-memory table and merges it into `table`
    uint64_t* part_scratch;                // [nparts][part_ncta][part_cap] tuples
    uint32_t* part_counts;                 // [nparts][part_ncta] tuples per bucket
    int32_t nparts, part_ncta, part_cap, part_begin;
    int32_t part_log2, part_groups;        // nparts = 1 << part_log2; accumulator rows of the pass-2 table (distinct keys it can hold)
    int32_t part_slots, part_tw;           // lookup slots of the pass-2 table (power of two); 64-bit words per tuple
    // shared-memory layout (byte offsets): [stage ring][front end]
    int32_t off_fe, off_dirkeys, off_dirstate, off_gid2slot, off_gslot, off_mm, off_cnt, off_sum, smem_bytes;
    StagePlan sp;
};

}  // namespace kq
