// kq_hashagg.cu — host side of HashAggregateExec (Main.kt:605-660): record layout, table growth,
// the per-batch launch of the specialised aggregate kernel (kq_k_agg.cuh via kq_codegen.cu / kq_jit.cu),
// finalisation into the single output batch (Main.kt:635-650) and the table maintenance kernels.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kq_aggtable.cuh"
#include "kq_codegen.h"
#include "kq_comm.h"
#include "kq_scan.cuh"

using namespace kq;

namespace {

// Tile geometry of the aggregate kernel (compiled in as KQ_R / KQ_WARPS), chosen per query shape from a
// measured sweep: the lane-private front end costs shared memory per consumer warp, so narrow rows (one
// or two aggregate inputs) run best with 7 warps x 4 rows, wide rows (TPC-H Q1 shape) with 6 warps x 8 rows.
struct AggGeometry {
    int r, warps;
    int tile() const { return warps * 32 * r; }
    int threads() const { return warps * 32 + 32; }
};
static AggGeometry agg_geometry(int ninputs) {
    if (const char* e = getenv("KQ_AGG_GEOM")) {       // tuning experiments: "rows,warps"
        int r = 0, w = 0;
        if (sscanf(e, "%d,%d", &r, &w) == 2 && r >= 2 && r <= 16 && r % 2 == 0 && w >= 1 && w <= 24) return AggGeometry{r, w};
    }
    return ninputs >= 3 ? AggGeometry{8, 6} : AggGeometry{4, 7};
}
constexpr int AGG_MAX_STAGES = 4;

// ---- table maintenance ---------------------------------------------------------------------------------------
__global__ void k_rehash(const uint64_t* __restrict__ old_table, uint64_t old_cap, uint64_t* table, uint64_t cap_mask, int stride, int nkeys) {
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < old_cap; s += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t* src = old_table + s * stride;
        if ((uint32_t)src[0] != HDR_FULL) continue;
        uint64_t kw[MAX_KEYS];
#pragma unroll
        for (int k = 0; k < MAX_KEYS; k++) kw[k] = k < nkeys ? src[1 + k] : 0;
        uint64_t slot = table_home(hash_key(kw, (uint32_t)(src[0] >> 32), nkeys), cap_mask);
        while (true) {
            uint64_t* rec = table + slot * stride;
            if (atomicCAS(reinterpret_cast<unsigned long long*>(rec), 0ULL, (unsigned long long)src[0]) == 0ULL) {
                for (int w = 1; w < stride; w++) rec[w] = src[w];
                break;
            }
            slot = (slot + 1) & cap_mask;
        }
    }
}

// Merge records produced elsewhere (other ranks' partials) into this table: the merge step of
// main()'s second query, MAX(max)/MIN(min)/SUM(sum)/SUM(count) (Main.kt:1320).
__global__ void k_merge_records(const AggArgs A, const uint64_t* __restrict__ recs, uint64_t nrecs) {
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < nrecs; s += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t* src = recs + s * A.stride;
        if ((uint32_t)src[0] != HDR_FULL) continue;
        uint64_t kw[MAX_KEYS];
#pragma unroll
        for (int k = 0; k < MAX_KEYS; k++) kw[k] = k < A.nkeys ? src[1 + k] : 0;
        uint32_t nullmask = (uint32_t)(src[0] >> 32);
        uint64_t* rec = table_find_or_insert(A, hash_key(kw, nullmask, A.nkeys), kw, nullmask);
        for (int i = 0; i < A.ninputs; i++) {
            const AggInput d = A.in[i];
            uint64_t c = src[d.rec_nn];
            if (!c) continue;
            atomicAdd(reinterpret_cast<unsigned long long*>(rec + d.rec_nn), (unsigned long long)c);
            if (d.flags & F_SUM) {
                if (d.flags & F_INT) atomicAdd(reinterpret_cast<unsigned long long*>(rec + d.rec_sum), (unsigned long long)src[d.rec_sum]);
                else atomicAdd(reinterpret_cast<double*>(rec + d.rec_sum), __longlong_as_double((long long)src[d.rec_sum]));
            }
            if (d.flags & F_MIN) atomicMin(reinterpret_cast<unsigned long long*>(rec + d.rec_min), (unsigned long long)src[d.rec_min]);
            if (d.flags & F_MAX) atomicMax(reinterpret_cast<unsigned long long*>(rec + d.rec_max), (unsigned long long)src[d.rec_max]);
        }
    }
}

// ---- the one-shot merge of small partial tables (BASELINE configs 3 and 5) ---------------------------------------------
// Block of one rank in the all-gather: [0] FULL records in its table (may exceed the block's room), [1] its status
// (device error bits of the aggregate), [2] cursor, [8..] up to `room` records.
constexpr uint64_t SMALL_MERGE_MAX = 1024;      // partial groups per rank up to which kq_hashagg_merge_allreduce takes the one-shot path
constexpr int MERGE_HDR = 8;
__global__ void k_collect_small(const uint64_t* __restrict__ table, uint64_t cap, int stride, uint64_t* block, uint64_t room, const uint32_t* __restrict__ err) {
    unsigned long long* cursor = reinterpret_cast<unsigned long long*>(block) + 2;
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < cap; s += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t* src = table + s * stride;
        if ((uint32_t)src[0] != HDR_FULL) continue;
        const unsigned long long pos = atomicAdd(cursor, 1ULL);
        if (pos < room) { uint64_t* dst = block + MERGE_HDR + pos * stride; for (int w = 0; w < stride; w++) dst[w] = src[w]; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) block[1] = *err;
}
__global__ void k_collect_small_done(uint64_t* block, const unsigned long long* __restrict__ heap_used) {
    block[0] = block[2];
    block[3] = heap_used ? heap_used[0] : 0;          // long Utf8 keys this rank has interned: entries, bytes
    block[4] = heap_used ? heap_used[1] : 0;
}
// This rank's long Utf8 group keys for the other ranks (a merged group's bytes must exist wherever the group ends up):
// blk[0] strings written, blk[1] 1 = they did not all fit; entries from word 2: {key word, length, bytes padded to 8}.
constexpr uint64_t STRBLK_WORDS = 8192;         // 64 KB per rank
__global__ void k_heap_export(const KeyHeap H, unsigned long long* blk) {
    unsigned long long* cursor = blk + 2;        // words used (starts at 3)
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s <= H.cap_mask; s += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long w = H.tab[2 * s];
        if (w < 2ULL) continue;
        const unsigned long long meta = H.tab[2 * s + 1];
        const uint32_t len = (uint32_t)(meta & 0xFFFFFFu);
        const uint64_t words = 2 + (len + 7) / 8;
        const unsigned long long pos = atomicAdd(cursor, (unsigned long long)words);
        if (pos + words > STRBLK_WORDS) { blk[1] = 1ULL; continue; }
        blk[pos] = w; blk[pos + 1] = len;
        const uint8_t* q = H.bytes + (meta >> 24);
        uint8_t* d = reinterpret_cast<uint8_t*>(blk + pos + 2);
        for (uint32_t i = 0; i < len; i++) d[i] = q[i];
        atomicAdd(blk, 1ULL);
    }
}
__global__ void k_heap_import(const KeyHeap H, const unsigned long long* __restrict__ all, int nranks, int me, uint32_t* err) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nranks || r == me) return;
    const unsigned long long* blk = all + (uint64_t)r * STRBLK_WORDS;
    if (blk[1]) atomicOr(err, 32u);              // that rank's strings did not fit the block: the merged keys would be incomplete
    uint64_t pos = 3;
    for (unsigned long long i = 0; i < blk[0] && pos + 2 <= STRBLK_WORDS; i++) {
        const uint32_t len = (uint32_t)blk[pos + 1];
        const uint64_t word = kq::utf8_intern(&H, reinterpret_cast<const uint8_t*>(blk + pos + 2), (int)len, err);
        if (word != blk[pos]) atomicOr(err, 16u);
        pos += 2 + (len + 7) / 8;
    }
}
// [count, status] of every rank's block -> out[2 * rank]
// [count, status, heap entries, heap bytes] of every rank's block -> out[4 * rank]
__global__ void k_merge_heads(const uint64_t* __restrict__ all, uint64_t pitch, int n, unsigned long long* out) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        out[4 * i] = all[(uint64_t)i * pitch]; out[4 * i + 1] = all[(uint64_t)i * pitch + 1];
        out[4 * i + 2] = all[(uint64_t)i * pitch + 3]; out[4 * i + 3] = all[(uint64_t)i * pitch + 4];
    }
}
// Rebuild this rank's table from the gathered partials, RANK BY RANK in one block: a key occurs at most once per rank,
// so within a rank no two threads touch the same record, and the barrier between ranks fixes the order of every
// Float64 addition — all ranks end with bit-identical tables.
__global__ void __launch_bounds__(1024) k_merge_small(const AggArgs A, const uint64_t* __restrict__ all, uint64_t pitch, int nranks, uint64_t room) {
    __shared__ unsigned int s_new;
    if (threadIdx.x == 0) s_new = 0;
    __syncthreads();
    for (int r = 0; r < nranks; r++) {
        const uint64_t* blk = all + (uint64_t)r * pitch;
        const uint64_t n = blk[0] < room ? blk[0] : room;
        for (uint64_t s = threadIdx.x; s < n; s += blockDim.x) {
            const uint64_t* src = blk + MERGE_HDR + s * A.stride;
            uint64_t kw[MAX_KEYS];
#pragma unroll
            for (int k = 0; k < MAX_KEYS; k++) kw[k] = k < A.nkeys ? src[1 + k] : 0;
            const uint32_t nullmask = (uint32_t)(src[0] >> 32);
            uint32_t fresh = 0;
            uint64_t* rec = table_find_or_insert(A, hash_key(kw, nullmask, A.nkeys), kw, nullmask, &fresh);
            if (fresh) atomicAdd(&s_new, 1u);
            for (int i = 0; i < A.ninputs; i++) {
                const AggInput d = A.in[i];
                const uint64_t c = src[d.rec_nn];
                if (!c) continue;
                rec[d.rec_nn] += c;                     // plain read-modify-write: this thread is the only one on this record in this round
                if (d.flags & F_SUM) {
                    if (d.flags & F_INT) rec[d.rec_sum] += src[d.rec_sum];
                    else rec[d.rec_sum] = (uint64_t)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)rec[d.rec_sum]), __longlong_as_double((long long)src[d.rec_sum])));
                }
                if (d.flags & F_MIN) rec[d.rec_min] = src[d.rec_min] < rec[d.rec_min] ? src[d.rec_min] : rec[d.rec_min];
                if (d.flags & F_MAX) rec[d.rec_max] = src[d.rec_max] > rec[d.rec_max] ? src[d.rec_max] : rec[d.rec_max];
            }
        }
        __threadfence();
        __syncthreads();
    }
    if (threadIdx.x == 0) *A.ngroups = s_new;
}

// ---- kernels of the multi-GPU merges -----------------------------------------------------------------------------
__device__ __forceinline__ int record_part(const uint64_t* src, int nkeys, int nparts) {
    if (nparts <= 1) return 0;
    uint64_t kw[MAX_KEYS];
#pragma unroll
    for (int k = 0; k < MAX_KEYS; k++) kw[k] = k < nkeys ? src[1 + k] : 0;
    return (int)((hash_key(kw, (uint32_t)(src[0] >> 32), nkeys) >> 20) % (uint64_t)nparts);
}
// records per destination rank (hash partitioning of the group keys)
__global__ void k_count_parts(const uint64_t* __restrict__ table, uint64_t cap, int stride, int nkeys, int nparts, unsigned long long* counts) {
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < cap; s += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t* src = table + s * stride;
        if ((uint32_t)src[0] != HDR_FULL) continue;
        atomicAdd(counts + record_part(src, nkeys, nparts), 1ULL);
    }
}
// Compact FULL records into a dense array, partition p starting at record base[p].
__global__ void k_collect_records(const uint64_t* __restrict__ table, uint64_t cap, int stride, int nkeys, uint64_t* out,
                                  unsigned long long* cursors, int nparts, const unsigned long long* __restrict__ base) {
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < cap; s += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t* src = table + s * stride;
        if ((uint32_t)src[0] != HDR_FULL) continue;
        const int part = record_part(src, nkeys, nparts);
        const unsigned long long pos = base[part] + atomicAdd(cursors + part, 1ULL);
        uint64_t* dst = out + pos * stride;
        for (int w = 0; w < stride; w++) dst[w] = src[w];
    }
}

// Union dictionary of the gathered partial records of all ranks: key -> position of its FIRST occurrence in
// the gathered buffer. The buffer is identical on every rank, so every rank derives the same dense index
// without any further exchange. D is a scratch table whose records are {header, keys..., position}.
__global__ void k_dict_build(const AggArgs D, const uint64_t* __restrict__ all, uint64_t nrecs, int stride) {
    const int posw = 1 + D.nkeys;
    for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < nrecs; j += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t* src = all + j * stride;
        if ((uint32_t)src[0] != HDR_FULL) continue;
        uint64_t kw[MAX_KEYS];
#pragma unroll
        for (int k = 0; k < MAX_KEYS; k++) kw[k] = k < D.nkeys ? src[1 + k] : 0;
        const uint32_t nullmask = (uint32_t)(src[0] >> 32);
        uint64_t* rec = table_find_or_insert(D, hash_key(kw, nullmask, D.nkeys), kw, nullmask);
        atomicMin(reinterpret_cast<unsigned long long*>(rec + posw), (unsigned long long)j);
    }
}
__global__ void k_dense_init(uint64_t* dense, uint64_t T, int stride, const AggArgs A) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < T * (uint64_t)stride; i += (uint64_t)gridDim.x * blockDim.x)
        dense[i] = A.rec_init[i / T];          // identity of each accumulator word (0; ~0 for MIN)
}
// own partial records -> dense arrays [word][position]
__global__ void k_dense_scatter(const AggArgs D, const uint64_t* __restrict__ all, uint64_t first, uint64_t count, int stride, uint64_t* dense, uint64_t T) {
    const int posw = 1 + D.nkeys;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t* src = all + (first + i) * stride;
        if ((uint32_t)src[0] != HDR_FULL) continue;
        uint64_t kw[MAX_KEYS];
#pragma unroll
        for (int k = 0; k < MAX_KEYS; k++) kw[k] = k < D.nkeys ? src[1 + k] : 0;
        const uint32_t nullmask = (uint32_t)(src[0] >> 32);
        const uint64_t p = table_find_or_insert(D, hash_key(kw, nullmask, D.nkeys), kw, nullmask)[posw];
        for (int w = 1 + D.nkeys; w < stride; w++) dense[(uint64_t)w * T + p] = src[w];
    }
}
// reduced dense arrays -> this rank's table (every rank ends up with the full merged result)
__global__ void k_dense_writeback(const AggArgs A, const AggArgs D, const uint64_t* __restrict__ all, uint64_t T, const uint64_t* __restrict__ dense) {
    const int posw = 1 + D.nkeys;
    for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < T; j += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t* src = all + j * A.stride;
        if ((uint32_t)src[0] != HDR_FULL) continue;
        uint64_t kw[MAX_KEYS];
#pragma unroll
        for (int k = 0; k < MAX_KEYS; k++) kw[k] = k < A.nkeys ? src[1 + k] : 0;
        const uint32_t nullmask = (uint32_t)(src[0] >> 32);
        const uint64_t h = hash_key(kw, nullmask, A.nkeys);
        if (table_find_or_insert(D, h, kw, nullmask)[posw] != j) continue;       // not the first occurrence of this key
        uint64_t* rec = table_find_or_insert(A, h, kw, nullmask);
        for (int w = 1 + A.nkeys; w < A.stride; w++) rec[w] = dense[(uint64_t)w * T + j];
    }
}

// ---- finalize: one output row per FULL record (Main.kt:639-647) -------------------------------------------------
struct FinKey { void* data; uint32_t* validity; int32_t type; int32_t _pad; };
struct FinAgg { void* data; uint32_t* validity; int32_t kind, word, nn_word, is_int, out_type, _pad; };
struct FinArgs {
    const uint64_t* table; uint64_t cap; int32_t stride, nkeys, naggs, _pad;
    FinKey keys[MAX_KEYS];
    FinAgg aggs[2 * MAX_INPUTS + 4];
    unsigned long long* pos;
    unsigned long long max_rows;            // rows the output columns were allocated for
    uint32_t* err;
};

// Output positions are reserved once per warp and chunk of 1024 table slots (one counter bumped by every warp
// for every 32 slots serialises in the L2 when the table has tens of millions of slots).
template <int CH>                              // slots per lane and chunk: 32 for big tables (few position atomics), 1 for small ones (every lane busy)
__global__ void k_finalize(const __grid_constant__ FinArgs F) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t c0 = warp0 * (32 * CH); c0 < F.cap; c0 += nwarps * (32 * CH)) {
        uint32_t mine = 0;
#pragma unroll (CH < 8 ? CH : 8)
        for (int j = 0; j < CH; j++) {
            const uint64_t s = c0 + (uint64_t)j * 32 + lane;
            if (s < F.cap && (uint32_t)F.table[s * F.stride] == HDR_FULL) mine |= 1u << j;
        }
        uint32_t incl = __popc(mine);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (!total) continue;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(F.pos, (unsigned long long)total);
        base = __shfl_sync(0xffffffffu, base, 0);
        uint64_t row = base + incl - __popc(mine);
        for (; mine; mine &= mine - 1, row++) {
            if (row >= F.max_rows) { atomicOr(F.err, 0x4000u); break; }      // more FULL records than the group counter said: never write past the columns
            const uint64_t* rec = F.table + (c0 + (uint64_t)(__ffs(mine) - 1) * 32 + lane) * F.stride;
            const uint32_t nullmask = (uint32_t)(rec[0] >> 32);
            for (int k = 0; k < F.nkeys; k++) {
                const FinKey o = F.keys[k];
                const uint64_t v = rec[1 + k];
                const bool valid = !((nullmask >> k) & 1u);
                if (o.validity && valid) atomicOr(o.validity + (row >> 5), 1u << (row & 31));
                switch (o.type) {
                    case KQ_F64: case KQ_I64: case KQ_UTF8: reinterpret_cast<uint64_t*>(o.data)[row] = v; break;   // UTF8: packed, expanded later
                    case KQ_DATE32: case KQ_I32: reinterpret_cast<uint32_t*>(o.data)[row] = (uint32_t)v; break;
                    case KQ_BOOL: if (v & 1u) atomicOr(reinterpret_cast<uint32_t*>(o.data) + (row >> 5), 1u << (row & 31)); break;
                }
            }
            for (int a = 0; a < F.naggs; a++) {
                const FinAgg o = F.aggs[a];
                const uint64_t nn = rec[o.nn_word];
                if (o.kind == KQ_AGG_COUNT) { reinterpret_cast<uint64_t*>(o.data)[row] = nn; continue; }   // Int64, never null (rule E7)
                if (nn == 0) continue;                                                                    // all-null group => null (R9)
                atomicOr(o.validity + (row >> 5), 1u << (row & 31));
                uint64_t v = rec[o.word];
                if (o.kind != KQ_AGG_SUM) {
                    v = order_unmap(v, o.is_int);
                    // Float64 MIN/MAX whose identity survived: every non-null value of the group was a NaN (NaNs never replace a held
                    // value) => NaN, as the reference yields when the first value is NaN and nothing compares greater (Main.kt:545-555)
                    if (!o.is_int && (v & 0x7fffffffffffffffULL) > 0x7ff0000000000000ULL) v = 0x7ff8000000000000ULL;
                }
                if (o.out_type == KQ_DATE32) reinterpret_cast<uint32_t*>(o.data)[row] = (uint32_t)v;
                else reinterpret_cast<uint64_t*>(o.data)[row] = v;
            }
        }
    }
}

struct PackedLen {
    const uint64_t* packed; const uint32_t* validity; KeyHeap heap;
    __device__ __forceinline__ int operator()(long long i) const {
        if (validity && !((validity[i >> 5] >> (i & 31)) & 1u)) return 0;
        const uint64_t w = packed[i];
        if (w >> 63) return (int)(heap_find(heap, w) & 0xFFFFFFu);          // a long key: its length is in the key heap
        return (int)(w >> 56);
    }
};
__global__ void k_unpack_utf8(const uint64_t* __restrict__ packed, const int32_t* __restrict__ off, const unsigned long long* __restrict__ nrows, uint8_t* out, const KeyHeap heap) {
    const uint64_t n = *nrows;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        int len = off[i + 1] - off[i];
        uint64_t p = packed[i];
        if (p >> 63) {
            const uint8_t* q = heap.bytes + (heap_find(heap, p) >> 24);
            for (int b = 0; b < len; b++) out[off[i] + b] = q[b];
        } else {
            for (int b = 0; b < len; b++) out[off[i] + b] = (uint8_t)(p >> (8 * b));
        }
    }
}

// ---- long Utf8 group keys (kq_rt.cuh utf8_intern): column statistics and key-heap maintenance ---------------------------------
__global__ void k_utf8_len_stats(const int32_t* __restrict__ off, int64_t n, unsigned long long* out) {      // out: [0] max length, [1] strings > 7 bytes, [2] their bytes
    unsigned long long mx = 0, cnt = 0, bytes = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long len = (unsigned long long)(off[i + 1] - off[i]);
        mx = len > mx ? len : mx;
        if (len > 7) { cnt++; bytes += len; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const unsigned long long m2 = __shfl_xor_sync(0xffffffffu, mx, o);
        mx = m2 > mx ? m2 : mx; cnt += __shfl_xor_sync(0xffffffffu, cnt, o); bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicMax(out, mx); if (cnt) { atomicAdd(out + 1, cnt); atomicAdd(out + 2, bytes); } }
}
__global__ void k_heap_rehash(const unsigned long long* __restrict__ old_tab, uint64_t old_cap, unsigned long long* tab, uint64_t cap_mask) {
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < old_cap; s += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long w = old_tab[2 * s];
        if (w < 2ULL) continue;
        uint64_t slot = heap_home(w, cap_mask);
        while (atomicCAS(tab + 2 * slot, 0ULL, w) != 0ULL) slot = (slot + 1) & cap_mask;
        tab[2 * slot + 1] = old_tab[2 * s + 1];
    }
}

int launch_check(kq_ctx* ctx, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return kq_cuda_fail(ctx, e, what);
    ctx->launches++;
    return KQ_OK;
}
int small_grid(kq_ctx* ctx, uint64_t items) {
    uint64_t g = (items + 255) / 256, cap = (uint64_t)ctx->sm_count * 8;
    return (int)std::max<uint64_t>(1, std::min(g, cap));
}

}  // namespace

// ---- host state ------------------------------------------------------------------------------------------------------
struct kq_hashagg {
    std::atomic<int> rc{1};
    kq_ctx* ctx = nullptr;
    kq_expr* pred = nullptr;
    std::vector<kq_expr*> groups;
    std::vector<int> agg_kinds;
    std::vector<int> agg_input_of;          // aggregate -> distinct input index
    std::vector<kq_expr*> inputs;           // distinct aggregate input expressions
    std::vector<bool> input_count_only;     // input used by COUNT only: validity is enough
    // fixed at the first batch
    bool typed = false;
    std::vector<int> key_types, input_types;
    // record layout (fixed at create)
    int stride = 0;
    AggInput in[MAX_INPUTS];
    uint64_t rec_init[MAX_REC_WORDS];
    // table
    uint64_t* table = nullptr;
    uint64_t capacity = 0;
    unsigned long long* d_counters = nullptr;   // [0] ngroups, [1] ticket (low 32 bits)
    int64_t ngroups_host = 0;
    int64_t expected_groups = 0;
    bool pessimistic = false;               // an optimistically sized table overflowed once: size for the rows in flight from now on
    uint64_t* snapshot = nullptr;           // copy of a small table taken before an optimistic launch (rollback on overflow)
    uint64_t snapshot_cap = 0;
    // Errors raised by this aggregate's kernels (Int64 / 0, malformed number in a cast) live in its OWN device word
    // (d_counters[5]) and are sticky: once seen, every later update / merge / finalize of this aggregate fails with the
    // same status — an unrelated call on the context can neither consume nor hide them.
    uint32_t dev_err = 0;
    bool groups_exact = true;               // ngroups_host is exact (after a merge it is an upper bound until finalize reads the count)
    // long Utf8 group keys: the key heap (kq_args.h KeyHeap); heap.used = d_counters + 6
    KeyHeap heap{};
    uint64_t heap_entries = 0, heap_bytes = 0;      // in use (as of the last counter read)
    // merge buffers (kq_hashagg_merge_allreduce), allocated once per aggregate
    uint64_t* merge_mine = nullptr; uint64_t* merge_all = nullptr; uint64_t merge_words = 0;
};

// d_counters layout (64-bit words): [0] groups in the table, [1] tile/partition tickets, [2] finalize cursor, [3] scratch max,
// [4] table-overflow flag, [5] device error bits of this aggregate, [6] merge scratch, [8..] Utf8 byte totals at finalize
static int sticky_error(kq_ctx* ctx, kq_hashagg* h) { return h->dev_err ? kq_device_error_status(ctx, h->dev_err) : KQ_OK; }
// counters of a finished launch: remembers device errors (sticky), reports an overflow of a conservatively sized table
static int read_counters(kq_ctx* ctx, kq_hashagg* h, uint64_t (&c)[6]) {
    uint64_t c8[8];
    KQ_RET(kq_read_u64(ctx, h->d_counters, 8, c8));
    for (int i = 0; i < 6; i++) c[i] = c8[i];
    h->heap_entries = c8[6]; h->heap_bytes = c8[7];
    if ((uint32_t)c[5]) { h->dev_err |= (uint32_t)c[5]; return sticky_error(ctx, h); }
    return KQ_OK;
}

static bool expr_equal(const kq_expr* a, const kq_expr* b) {
    if (a == b) return true;
    if (!a || !b || a->kind != b->kind) return false;
    switch (a->kind) {
        case KQ_EX_COL: return a->col == b->col;
        case KQ_EX_LIT: return a->type == b->type && a->is_null == b->is_null && a->i == b->i && memcmp(&a->f, &b->f, 8) == 0 && a->s == b->s;
        case KQ_EX_CAST: return a->type == b->type && expr_equal(a->l, b->l);
        case KQ_EX_BIN: return a->op == b->op && expr_equal(a->l, b->l) && expr_equal(a->r, b->r);
    }
    return false;
}

static int table_alloc(kq_ctx* ctx, kq_hashagg* h, uint64_t capacity) {
    size_t bytes = (size_t)(capacity + 1) * h->stride * 8;       // + the dummy record a full table hands out (kq_aggtable.cuh)
    KQ_RET(kq_dev_alloc(ctx, bytes, (void**)&h->table));
    KQ_CUDA(ctx, cudaMemsetAsync(h->table, 0, bytes, ctx->stream));
    h->capacity = capacity;
    return KQ_OK;
}

static int table_grow(kq_ctx* ctx, kq_hashagg* h, uint64_t new_capacity) {
    uint64_t* old = h->table; uint64_t old_cap = h->capacity;
    h->table = nullptr;
    int st = table_alloc(ctx, h, new_capacity);
    if (st != KQ_OK) { h->table = old; h->capacity = old_cap; return st; }
    k_rehash<<<small_grid(ctx, old_cap), 256, 0, ctx->stream>>>(old, old_cap, h->table, new_capacity - 1, h->stride, (int)h->groups.size());
    st = launch_check(ctx, "k_rehash");
    kq_dev_free(ctx, old);
    return st;
}

static void fill_common_args(kq_hashagg* h, AggArgs& A) {
    A.nkeys = (int)h->groups.size();
    A.ninputs = (int)h->inputs.size();
    A.stride = h->stride;
    memcpy(A.in, h->in, sizeof h->in);
    memcpy(A.rec_init, h->rec_init, sizeof h->rec_init);
    A.table = h->table;
    A.cap_mask = h->capacity - 1;
    A.ngroups = h->d_counters;
    A.ticket = (unsigned int*)(h->d_counters + 1);
    A.err = (uint32_t*)(h->d_counters + 5);
    A.overflow = (unsigned int*)(h->d_counters + 4);
    A.heap = h->heap;
    A.trace = getenv("KQ_FE_PROGRESS") ? (unsigned long long*)strtoull(getenv("KQ_FE_PROGRESS"), nullptr, 0) : nullptr;
    A.stop_threshold = ~0ULL;
}

static void hashagg_delete_host(kq_hashagg* h) {
    kq_expr_free(h->pred);
    for (kq_expr* e : h->groups) kq_expr_free(e);
    for (kq_expr* e : h->inputs) kq_expr_free(e);
    delete h;
}

extern "C" {

// Host-only part of construction: expression bookkeeping and the record layout (no CUDA calls).
static int hashagg_new(kq_ctx* ctx, kq_expr* pred, kq_expr* const* group_exprs, int ngroup, const int* agg_kinds,
                       kq_expr* const* agg_inputs, int nagg, int64_t expected_groups, kq_hashagg** out) {
    if (ngroup > MAX_KEYS) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "more than %d group expressions", MAX_KEYS);
    for (int i = 0; i < nagg; i++)
        if (agg_kinds[i] < KQ_AGG_MAX || agg_kinds[i] > KQ_AGG_COUNT)
            return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "Unsupported aggregate function: %d", agg_kinds[i]);   // Main.kt:696
    kq_hashagg* h = new kq_hashagg();
    h->ctx = ctx;
    h->expected_groups = expected_groups;
    if (pred) { pred->rc.fetch_add(1); h->pred = pred; }
    for (int i = 0; i < ngroup; i++) { group_exprs[i]->rc.fetch_add(1); h->groups.push_back(group_exprs[i]); }
    int flags[MAX_INPUTS] = {0};
    for (int i = 0; i < nagg; i++) {
        int idx = -1;
        for (size_t j = 0; j < h->inputs.size(); j++) if (expr_equal(h->inputs[j], agg_inputs[i])) { idx = (int)j; break; }
        if (idx < 0) {
            if (h->inputs.size() >= (size_t)MAX_INPUTS) { hashagg_delete_host(h); return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "more than %d distinct aggregate inputs", MAX_INPUTS); }
            agg_inputs[i]->rc.fetch_add(1);
            h->inputs.push_back(agg_inputs[i]);
            idx = (int)h->inputs.size() - 1;
        }
        h->agg_kinds.push_back(agg_kinds[i]);
        h->agg_input_of.push_back(idx);
        if (agg_kinds[i] == KQ_AGG_SUM) flags[idx] |= F_SUM;
        if (agg_kinds[i] == KQ_AGG_MIN) flags[idx] |= F_MIN;
        if (agg_kinds[i] == KQ_AGG_MAX) flags[idx] |= F_MAX;
    }
    // record layout: [0] header {state:32, key nullmask:32}, [1..K] key words, then per input: nn, sum, min, max
    int w = 1 + ngroup;
    for (int i = 0; i < MAX_REC_WORDS; i++) h->rec_init[i] = 0;
    for (size_t i = 0; i < h->inputs.size(); i++) {
        AggInput& d = h->in[i];
        d.flags = flags[i];
        d.rec_nn = w++;
        d.rec_sum = (flags[i] & F_SUM) ? w++ : -1;
        d.rec_min = (flags[i] & F_MIN) ? w++ : -1;
        d.rec_max = (flags[i] & F_MAX) ? w++ : -1;
        if (d.rec_min >= 0) h->rec_init[d.rec_min] = ~0ULL;
        d.fe_sum = d.fe_min = d.fe_max = -1;
        h->input_count_only.push_back(flags[i] == 0);
    }
    h->stride = (w + 3) / 4 * 4;
    if (h->stride > MAX_REC_WORDS) { hashagg_delete_host(h); return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "aggregate record too wide"); }
    *out = h;
    return KQ_OK;
}

int kq_hashagg_create(kq_ctx* ctx, kq_expr* pred, kq_expr* const* group_exprs, int ngroup, const int* agg_kinds,
                      kq_expr* const* agg_inputs, int nagg, int64_t expected_groups, kq_hashagg** out) {
    if (!ctx || !out || ngroup < 0 || nagg < 0) return KQ_ERR_ILLEGAL_ARGUMENT;
    cudaSetDevice(ctx->device);
    kq_hashagg* h = nullptr;
    KQ_RET(hashagg_new(ctx, pred, group_exprs, ngroup, agg_kinds, agg_inputs, nagg, expected_groups, &h));
    // counters and table come from the ctx's caching allocator: creating an aggregate costs no cudaMalloc after the first query
    int st = kq_dev_alloc(ctx, 8192, (void**)&h->d_counters);      // [0, 256): counters; [1024, 4096): per-block progress; [4096, 8192): its snapshot
    if (st != KQ_OK) { kq_hashagg_free(h); return st; }
    cudaError_t e = cudaMemsetAsync(h->d_counters, 0, 8192, ctx->stream);
    if (e != cudaSuccess) { kq_hashagg_free(h); return kq_cuda_fail(ctx, e, "cudaMemsetAsync"); }
    // sized from the planner's hint: a 50-group query gets a 1024-slot table (64 KB), not one sized for the rows in flight
    uint64_t cap = 1024;
    while ((int64_t)cap < expected_groups * 2) cap <<= 1;
    st = table_alloc(ctx, h, cap);
    if (st != KQ_OK) { kq_hashagg_free(h); return st; }
    *out = h;
    return KQ_OK;
}

int kq_hashagg_free(kq_hashagg* h) {
    if (!h) return KQ_OK;
    if (h->rc.fetch_sub(1) == 1) {
        cudaSetDevice(h->ctx->device);
        kq_expr_free(h->pred);
        for (kq_expr* e : h->groups) kq_expr_free(e);
        for (kq_expr* e : h->inputs) kq_expr_free(e);
        kq_dev_free(h->ctx, h->table);
        kq_dev_free(h->ctx, h->d_counters);
        kq_dev_free(h->ctx, h->snapshot);
        kq_dev_free(h->ctx, h->merge_mine);
        kq_dev_free(h->ctx, h->merge_all);
        kq_dev_free(h->ctx, h->heap.tab);
        kq_dev_free(h->ctx, h->heap.bytes);
        delete h;
    }
    return KQ_OK;
}

}  // extern "C"

// Everything about an aggregate launch that depends on the query SHAPE only: generated source, stage
// plan, front-end layout. Fills the shape-dependent fields of A. No CUDA calls (kq_explain_hashagg).
// mode 0: front end + global table. mode 1: pass 1 of the partitioned path (no front end; `nparts` cursors in shared
// memory) — the same translation unit also holds pass 2 (kq_agg_partition_reduce), whose table geometry is returned in
// A.part_slots / *reduce_smem.
static int plan_agg(kq_ctx* ctx, kq_hashagg* h, kq_batch* input, int smem_optin, AggArgs& A, std::string* defines_out, std::string* gen_out,
                    int mode = 0, int nparts = 0, int* reduce_smem = nullptr) {
    // generate: [predicate -> selection] keys..., inputs...
    KqCodegen cg;
    KQ_RET(cg.begin(ctx, input));
    if (h->pred) {
        KqVal p;
        KQ_RET(cg.value(h->pred, &p));
        if (p.type != KQ_BOOL) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "filter predicate is not Bool");
        cg.line("sink.sel = " + p.v + (p.nullable() ? " & " + p.ok : std::string()) + " & rc.inr;");   // TRUE only (rule E3)
        cg.line("rc.active = sink.sel;");
    }
    std::vector<int> kt, it;
    uint32_t key_f64_mask = 0;
    bool keys_nullable = false;         // statically non-null keys: no null-mask hashing / comparing in the front end
    for (size_t k = 0; k < h->groups.size(); k++) {
        KqVal v;
        KQ_RET(cg.key_value(h->groups[k], &v));
        const KqVal a = cg.as_array(v);
        cg.line("sink.template set_key<" + std::to_string(k) + ">(" + a.v + ", " + a.okx() + ");");
        if (v.type == KQ_F64) key_f64_mask |= 1u << k;
        keys_nullable |= a.nullable();
        kt.push_back(v.type);
    }
    std::vector<int> in_cnt;            // front-end count slot per input: 0 = "every selected row" (input statically non-null)
    int ncnt = 1;
    bool cnt0_used = false;
    for (size_t i = 0; i < h->inputs.size(); i++) {
        int t; bool nl;
        KQ_RET(cg.infer(h->inputs[i], &t, &nl));
        KqVal v;
        if (h->input_count_only[i]) KQ_RET(cg.validity_only(h->inputs[i], &v));
        else {
            int fl = h->in[i].flags;
            // MaxAccumulator throws UnsupportedOperationException for other types (Main.kt:548-550)
            if (t != KQ_F64 && t != KQ_I64 && !(t == KQ_DATE32 && !(fl & F_SUM)))
                return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "%s is not implemented for data type %d", (fl & F_SUM) ? "SUM" : "MIN/MAX", t);
            KQ_RET(cg.value(h->inputs[i], &v));
        }
        const KqVal a = cg.as_array(v);
        cg.line("sink.template set_in<" + std::to_string(i) + ">(" + a.v + ", " + a.okx() + ");");
        it.push_back(t);
        if (a.nullable()) in_cnt.push_back(ncnt++); else { in_cnt.push_back(0); cnt0_used = true; }
    }
    const std::string eval_body = cg.take_body();
    if (!h->typed) { h->key_types = kt; h->input_types = it; h->typed = true; }
    else if (kt != h->key_types || it != h->input_types) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "batch schema differs from earlier batches");
    for (size_t i = 0; i < h->inputs.size(); i++) {
        if (it[i] != KQ_F64) h->in[i].flags |= F_INT; else h->in[i].flags &= ~F_INT;
        if (h->in[i].rec_max >= 0) h->rec_init[h->in[i].rec_max] = 0ULL;
    }
    A.key_f64_mask = key_f64_mask;

    // front-end layout (compile-time constants of the specialised kernel)
    const int NI = (int)h->inputs.size(), NK = (int)h->groups.size();
    int ns = 0, nm = 0;
    uint32_t mm_ismin = 0;
    std::vector<int> mm_word;
    for (int i = 0; i < NI; i++) {
        AggInput& d = h->in[i];
        d.fe_sum = (d.flags & F_SUM) ? ns++ : -1;
        d.fe_min = (d.flags & F_MIN) ? nm++ : -1;
        d.fe_max = (d.flags & F_MAX) ? nm++ : -1;
        if (d.fe_min >= 0) { mm_word.push_back(d.rec_min); mm_ismin |= 1u << d.fe_min; }
        if (d.fe_max >= 0) mm_word.push_back(d.rec_max);
    }
    // Geometry + shared-memory split. Candidates in order of measured speed (more rows per thread = more
    // independent work per warp; the lane-private front end costs shared memory per consumer warp); take the
    // first one whose front end holds the expected number of groups (64 when unknown), else the roomiest.
    static const AggGeometry CANDIDATES[] = {{6, 10}, {10, 7}, {8, 7}, {8, 6}, {6, 6}, {4, 7}, {4, 4}};     // {6,10}: a handful of groups (config 5: 3.77 vs 3.36 TB/s with {10,7})
    const int need_groups = (int)std::min<int64_t>(FE_MAX_GROUPS, h->expected_groups > 0 ? h->expected_groups : FE_MAX_GROUPS);
    const int entry_words = (NK + 2) / 2 * 2;
    AggGeometry geo = CANDIDATES[0];
    std::string stage_defs;
    int fg = -1, dir_slots = 1024;
    bool has_bytes = false;
    auto try_geometry = [&](const AggGeometry& g, StagePlan* sp, std::string* defs, int* dir_out) {
        // stage ring first (2..4 stages within ~64 KB, more only if two stages need it), front end gets the rest
        const int ring_budget = getenv("KQ_AGG_RING_KB") ? atoi(getenv("KQ_AGG_RING_KB")) * 1024 : 64 * 1024;      // tuning experiments
        *defs = cg.plan_stages(ring_budget, 1, g.tile(), sp, true);
        if (sp->nstages < 2) *defs = cg.plan_stages(std::min(2 * sp->stage_bytes, 112 * 1024), 2, g.tile(), sp, true);
        sp->nstages = std::max(1, std::min(sp->nstages, AGG_MAX_STAGES));
        const int ring = sp->nstages * sp->stage_bytes;
        const int per_group = g.warps * 32 * (4 * ncnt + 8 * ns) + 8 + 4 + 8 * nm;     // the kernel adds one trash group (fg + 1)
        int dir = 1024;
        int budget = smem_optin - 3072 - ring;
        while (dir > 256 && dir * entry_words * 8 + (need_groups + 1) * per_group > budget) dir >>= 1;
        budget -= dir * entry_words * 8 + 64;
        *dir_out = dir;
        return std::max(0, std::min(FE_MAX_GROUPS, budget / per_group - 1));
    };
    const char* forced = getenv("KQ_AGG_GEOM");
    int fe_smem = 0, ctas = 1;
    if (mode == 3) {
        // kq_k_agg_fe.cuh: lane-private SUM/COUNT blocks per consumer warp (the shared-memory data path bounds this kernel,
        // so the warp count only has to hide latency), a stage ring with at least ~48 KB in flight beyond the stage being
        // consumed, and a directory sparse enough for perfect placement (>= 8 slots per group).
        const int gs = ns * 256 + ncnt * 128, nmm1 = std::max(nm, 1), nkw = std::max(NK, 1);
        int want = h->expected_groups > 0 ? (int)std::min<int64_t>(FE_MAX_GROUPS, h->expected_groups + 2) : FE_MAX_GROUPS;
        want = std::max(want, 4);
        auto fe_bytes = [&](int g, int warps, int dir) {
            return dir * 4 + dir * nkw * 8 + g * nkw * 8 + (g + 1) * nmm1 * 8 + g * 8 + ((g + 3) & ~3) * 4 + warps * (g + 1) * gs;
        };
        int fr = 0, fw = 0, fs = 0, fc = 0;                           // tuning experiments: KQ_AGG_GEOM="rows,warps[,stages[,CTAs per SM]]"
        if (forced) sscanf(forced, "%d,%d,%d,%d", &fr, &fw, &fs, &fc);
        ctas = fc > 0 ? std::min(fc, 4) : 1;
        // per CTA: its share of the SM's 228 KB (1 KB per resident CTA is the system's), minus the kernel's static shared
        // memory (barriers, tile bookkeeping, directory control: ~1.8 KB)
        const int budget = std::min(smem_optin, (228 * 1024 - ctas * 1024) / ctas) - 2048;
        struct Cand { int r, warps; };
        // measured (B200, configs 3 and 5): the kernel is latency-bound, so consumer warps count most, then rows per thread (8-12; the
        // producer lane's per-tile work is serial); two stages are enough
        static const Cand CAND[] = {{8, 7}, {6, 7}, {12, 6}, {8, 6}, {8, 5}, {6, 6}, {8, 4}, {4, 7}, {4, 6}, {4, 5}, {4, 4}, {2, 6}, {2, 4}, {2, 2}};
        bool found = false;
        // the first candidate (most consumer warps first) that holds `want` groups; a directory of fewer groups only if nothing does
        for (int g = want; g >= 4 && !found; g = g * 3 / 4) {
            for (const Cand& c : CAND) {
                const int r = fr > 0 ? fr : c.r, w = fw > 0 ? fw : c.warps;
                int dir = 256;
                while (dir < 16 * g) dir <<= 1;
                if (const char* e = getenv("KQ_AGG_DIR")) dir = std::max(64, atoi(e));       // tuning experiments
                StagePlan sp;
                const std::string defs = cg.plan_stages(1 << 30, 1, w * 32 * r, &sp, true);
                int fe = fe_bytes(g, w, dir);
                // two stages at least, and ~24 KB in flight beyond the stage being consumed (HBM latency x the SM's bandwidth share)
                int ring_min = std::max(2 * sp.stage_bytes, sp.stage_bytes + 24 * 1024);
                if (fr > 0) ring_min = 2 * sp.stage_bytes;
                if (fe + ring_min > budget) { dir >>= 1; fe = fe_bytes(g, w, dir); }       // 8 slots per group still places within a few attempts
                if (fe + ring_min <= budget) {
                    int st = std::min((budget - fe) / sp.stage_bytes, 6);
                    if (fs > 0) st = std::min(st, std::max(fs, 2));
                    sp.nstages = st;
                    if (getenv("KQ_TRACE_AGG"))
                        fprintf(stderr, "kq fe geometry: %d rows x %d warps, %d stages of %d bytes, directory %d slots for %d groups, front end %d bytes of %d\n", r, w, st, sp.stage_bytes, dir, g, fe, budget);
                    geo = AggGeometry{r, w}; A.sp = sp; stage_defs = defs; fg = g; dir_slots = dir; fe_smem = fe;
                    found = true;
                    break;
                }
                if (fr > 0) break;
            }
        }
        if (!found) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "aggregate too wide for the shared-memory front end");
    }
    if (mode == 2) {
        // mid cardinality: block-shared table; many warps hide the latency of its shared-memory atomics
        geo = forced ? agg_geometry(NI) : AggGeometry{4, 15};
        stage_defs = cg.plan_stages(64 * 1024, 1, geo.tile(), &A.sp, true);
        A.sp.nstages = std::max(1, std::min(A.sp.nstages, AGG_MAX_STAGES));
        fg = 0; dir_slots = 4;
    }
    if (mode == 1) {
        // pass 1 of the partitioned path: no lane-private state, so many warps (latency of the shared-memory cursor
        // atomics and of the scattered stores is hidden by occupancy) and a deep stage ring
        geo = forced ? agg_geometry(NI) : AggGeometry{4, 15};
        stage_defs = cg.plan_stages(96 * 1024, 1, geo.tile(), &A.sp, true);
        A.sp.nstages = std::max(1, std::min(A.sp.nstages, AGG_MAX_STAGES));
        fg = 0; dir_slots = 4;
    }
    for (const AggGeometry& g : CANDIDATES) {
        if (mode != 0) break;
        const AggGeometry cand = forced ? agg_geometry(NI) : g;
        StagePlan sp; std::string defs; int dir;
        const int f = try_geometry(cand, &sp, &defs, &dir);
        if (f > fg) { fg = f; geo = cand; A.sp = sp; stage_defs = defs; dir_slots = dir; }
        if (f >= need_groups || forced) break;
    }
    const int TILE = geo.tile(), WARPS = geo.warps;
    (void)TILE;
    A.q = cg.args;
    for (int b = 0; b < A.sp.nbuf; b++) has_bytes |= A.sp.buf[b].kind == SK_BYTES;
    const int ring = A.sp.nstages * A.sp.stage_bytes;
    A.fe_groups = fg;
    A.geo_r = geo.r; A.geo_warps = geo.warps; A.geo_ctas = ctas; A.geo_service = (mode == 3 && nm > 0) ? 2 : 1;
    A.smem_bytes = ring + dir_slots * entry_words * 8 + std::max(fg, 1) * 12 + (fg + 1) * nm * 8 + WARPS * (fg + 1) * 32 * (8 * ns + 4 * ncnt);
    if (mode == 3) A.smem_bytes = ring + fe_smem;
    if (mode == 1) A.smem_bytes += 16 + nparts * 4;
    A.smem_bytes = (A.smem_bytes + 127) / 128 * 128;
    {   // pass-2 table: a sparse LOOKUP part (state + key words per slot, power-of-two slots, <= 45 % full so that probe
        // sequences stay short) pointing into a dense ACCUMULATOR part (one row of counts / sums / extremes per distinct key)
        const int lookup_bytes = 4 + 8 * NK, acc_bytes = 4 * ncnt + 8 * ns + 8 * nm;
        // mode 2: the table shares the block's shared memory with the stage ring and the (empty) front end
        const int avail = mode == 2 ? smem_optin - 2048 - A.smem_bytes : smem_optin - 4096;
        int sl = 256;
        while (sl < 16384 && sl * 2 * lookup_bytes <= avail * 62 / 100) sl <<= 1;
        const int ac = std::min(sl / 2, (avail - sl * lookup_bytes) / acc_bytes) / 4 * 4;       // accumulator rows: lookup part at most half full
        if (ac < 64) { if (mode == 2) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "aggregate record too wide for the shared-memory table"); }
        A.part_slots = sl;
        A.part_groups = std::max(ac, 0);
        const int table_bytes = (sl * lookup_bytes + std::max(ac, 0) * acc_bytes + 15) / 16 * 16;
        if (reduce_smem) *reduce_smem = table_bytes;
        if (mode == 2) A.smem_bytes = (A.smem_bytes + 32 + table_bytes + 127) / 128 * 128;
        int tw = NK + ((keys_nullable || ncnt > 1) ? 1 : 0);
        for (int i = 0; i < NI; i++) if (h->in[i].flags & (F_SUM | F_MIN | F_MAX)) tw++;
        A.part_tw = std::max(1, tw);
    }

    auto arr = [&](const char* type, const char* name, int nelem, auto get) {
        std::string s = std::string("    static constexpr ") + type + " " + name + "[" + std::to_string(std::max(nelem, 1)) + "] = {";
        for (int i = 0; i < std::max(nelem, 1); i++) s += (i ? ", " : "") + std::to_string(i < nelem ? get(i) : 0);
        return s + "};\n";
    };
    std::string consts;
    consts += "    static constexpr int NKEYS = " + std::to_string(NK) + ", NIN = " + std::to_string(NI) + ", NCNT = " + std::to_string(ncnt) +
              ", NSUM = " + std::to_string(ns) + ", NMM = " + std::to_string(nm) + ";\n";
    consts += std::string("    static constexpr bool CNT0_USED = ") + (cnt0_used ? "true" : "false") + ", KEYS_NULLABLE = " + (keys_nullable ? "true" : "false") + ";\n";
    consts += "    static constexpr uint32_t KEY_F64_MASK = " + std::to_string(key_f64_mask) + "u, MM_ISMIN = " + std::to_string(mm_ismin) + "u;\n";
    consts += arr("int", "IN_FLAGS", NI, [&](int i) { return h->in[i].flags; });
    consts += arr("int", "IN_CNT", NI, [&](int i) { return in_cnt[(size_t)i]; });
    consts += arr("int", "REC_NN", NI, [&](int i) { return h->in[i].rec_nn; });
    consts += arr("int", "REC_SUM", NI, [&](int i) { return h->in[i].rec_sum; });
    consts += arr("int", "FE_SUM", NI, [&](int i) { return h->in[i].fe_sum; });
    consts += arr("int", "FE_MIN", NI, [&](int i) { return h->in[i].fe_min; });
    consts += arr("int", "FE_MAX", NI, [&](int i) { return h->in[i].fe_max; });
    consts += arr("int", "MM_WORD", nm, [&](int i) { return mm_word[(size_t)i]; });
    {   // MIN/MAX slot -> is its input nullable (kq_k_agg_fe.cuh: a group may then hold no value for it)
        std::vector<int> mm_nullable((size_t)std::max(nm, 1), 0);
        bool any = false;
        for (int i = 0; i < NI; i++) {
            const bool nl = in_cnt[(size_t)i] > 0;
            if (h->in[i].fe_min >= 0) { mm_nullable[(size_t)h->in[i].fe_min] = nl; any |= nl; }
            if (h->in[i].fe_max >= 0) { mm_nullable[(size_t)h->in[i].fe_max] = nl; any |= nl; }
        }
        consts += arr("bool", "MM_NULLABLE", nm, [&](int i) { return mm_nullable[(size_t)i]; });
        consts += std::string("    static constexpr bool ANY_MM_NULLABLE = ") + (any ? "true" : "false") + ";\n";
    }
    // one input with both MIN and MAX in adjacent, 16-byte aligned record words: the global path reads them with one load
    const bool mm_paired = NI == 1 && nm == 2 && mm_word[1] == mm_word[0] + 1 && mm_word[0] % 2 == 0 && h->in[0].fe_min == 0;
    consts += std::string("    static constexpr bool MM_PAIRED = ") + (mm_paired ? "true" : "false") + ";\n";
    *gen_out = "namespace kq {\n" + stage_defs + "struct Q {\n" + consts +
                            "    template <class Sink> static __device__ __forceinline__ void eval(const QArgs& q, RowCtx& rc, Sink& sink) {\n" +
                            eval_body + "    }\n};\n}  // namespace kq\n";
    *defines_out = "#define KQ_R " + std::to_string(geo.r) + "\n#define KQ_WARPS " + std::to_string(geo.warps) + "\n#define KQ_STAGES " + std::to_string(A.sp.nstages) + "\n#define KQ_FE_GROUPS " + std::to_string(fg) + "\n#define KQ_DIR_SLOTS " +
                                std::to_string(dir_slots) + (getenv("KQ_L2_PREFETCH") ? "\n#define KQ_L2_PREFETCH " + std::to_string(atoi(getenv("KQ_L2_PREFETCH"))) : std::string()) +
                                "\n#define KQ_CTAS " + std::to_string(ctas) + "\n#define KQ_STAGE_BYTES " + (has_bytes ? "1" : "0") + "\n#define KQ_AGG_MODE " + std::to_string(mode) + "\n" +
                                (getenv("KQ_PART_L2_HINTS") ? "#define KQ_PART_L2_HINTS " + std::to_string(atoi(getenv("KQ_PART_L2_HINTS"))) + "\n" : std::string()) +
                                (getenv("KQ_NO_STAGE_DRAIN") ? "#define KQ_NO_STAGE_DRAIN 1\n" : "") + (getenv("KQ_RING_CHECK") ? "#define KQ_RING_CHECK " + std::to_string(atoi(getenv("KQ_RING_CHECK"))) + "\n" : std::string()) + (getenv("KQ_PART_SCALAR_MERGE") ? "#define KQ_PART_SCALAR_MERGE 1\n" : "") + (getenv("KQ_PART_DROP_SPILL") ? "#define KQ_PART_DROP_SPILL 1\n" : "") + (getenv("KQ_FE_CHECK") ? "#define KQ_FE_CHECK 1\n" : "") + (getenv("KQ_FE_NOEXACT") ? "#define KQ_FE_NOEXACT 1\n" : "") +
                                (getenv("KQ_FE_PROGRESS") ? "#define KQ_FE_TRACE 1\n" : "") + (getenv("KQ_FE_NOMERGE") ? "#define KQ_FE_NOMERGE 1\n" : "");
    return KQ_OK;
}

// ---- the partitioned path (kq_k_agg.cuh, KQ_AGG_MODE 1) ------------------------------------------------------------------
constexpr int64_t PART_MIN_GROUPS = 16384;          // below this the plain path's global table stays L2-resident
constexpr int64_t PART_MIN_ROWS = 1 << 20;          // small batches do not amortise the bucket scratch
constexpr int64_t PART_CHUNK_ROWS = 1LL << 28;      // rows per pass-1 launch (bounds the scratch: ~1.3 tuples per row)

// largest bucket fill of a pass-1 launch (bounds the groups one pass-2 block can add)
__global__ void k_max_u32(const uint32_t* __restrict__ v, uint64_t n, unsigned long long* out) {
    uint32_t m = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) m = max(m, v[i]);
#pragma unroll
    for (int o = 16; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)m);
}

static int hashagg_update_partitioned(kq_ctx* ctx, kq_hashagg* h, kq_batch* input, int64_t n, int64_t g_est) {
    AggArgs A;
    memset(&A, 0, sizeof A);
    std::string defines, gen;
    // shape first (table slots of pass 2 decide the partition count), then the real plan
    int reduce_smem = 0;
    KQ_RET(plan_agg(ctx, h, input, ctx->max_smem_optin, A, &defines, &gen, 1, 512, &reduce_smem));
    int nparts = 512;
    while (nparts < 4096 && (int64_t)nparts * (A.part_groups - 512) * 7 / 10 < g_est) nparts <<= 1;       // partitions ~70 % of what a table holds
    if (const char* e = getenv("KQ_PARTS")) { int v = atoi(e); if (v >= 16 && v <= 8192 && (v & (v - 1)) == 0) nparts = v; }
    memset(&A, 0, sizeof A);
    KQ_RET(plan_agg(ctx, h, input, ctx->max_smem_optin, A, &defines, &gen, 1, nparts, &reduce_smem));
    const AggGeometry geo{A.geo_r, A.geo_warps};
    const int TILE = geo.tile(), THREADS = geo.threads();
    const int64_t ntiles = (n + TILE - 1) / TILE;
    void *k_scatter = nullptr, *k_reduce = nullptr;
    KQ_RET(kq_jit_kernel(ctx, defines, gen, KQ_SKEL_AGG, "kq_hash_aggregate", A.smem_bytes, &k_scatter));
    KQ_RET(kq_jit_kernel(ctx, defines, gen, KQ_SKEL_AGG, "kq_agg_partition_reduce", reduce_smem, &k_reduce));
    int grid = (int)std::min<int64_t>(ntiles, (int64_t)ctx->sm_count);
    if (const char* e = getenv("KQ_PART_GRID")) { int v = atoi(e); if (v >= 1 && v <= grid) grid = v; }      // tuning experiments
    const int64_t chunk_tiles = std::max<int64_t>(1, PART_CHUNK_ROWS / TILE);
    const int64_t chunk_rows = std::min<int64_t>(n, chunk_tiles * TILE);
    // bucket capacity: the mean fill of a (partition, block) bucket plus slack for hash and scheduling imbalance; a
    // bucket that overflows anyway (skewed keys) spills its rows to the plain global path inside pass 1
    const double mean = (double)chunk_rows / ((double)nparts * grid);
    int part_cap = (int)(mean * 1.25 + 4.0 * sqrt(mean) + 16.0);
    if (const char* e = getenv("KQ_PART_CAP")) part_cap = std::max(part_cap, atoi(e));      // debugging: no bucket overflows
    const int tw = A.part_tw;           // tuple words (plan_agg; the kernel derives the same number from the compiled-in layout)
    uint64_t* scratch = nullptr; uint32_t* counts = nullptr;
    const size_t buckets = (size_t)nparts * grid;
    KQ_RET(kq_dev_alloc(ctx, buckets * (size_t)part_cap * tw * 8, (void**)&scratch));
    const bool ring_check = getenv("KQ_RING_CHECK") != nullptr;        // debugging: per-tile consumption counters behind the bucket counts
    int st = kq_dev_alloc(ctx, buckets * 4 + 64 + (ring_check ? (size_t)ntiles * 4 : 0), (void**)&counts);
    if (st != KQ_OK) { kq_dev_free(ctx, scratch); return st; }
    auto done = [&](int s) { kq_dev_free(ctx, scratch); kq_dev_free(ctx, counts); return s; };
    unsigned long long* d_max = h->d_counters + 3;

    // pass 2 needs room for every group a partition in flight may add; keep a floor under the table size
    const uint64_t margin1 = (uint64_t)grid * ((uint64_t)(A.sp.nstages + 1) * TILE);
    int64_t tile_begin = 0;
    while (tile_begin < ntiles) {
        const int64_t tile_end = std::min(ntiles, tile_begin + chunk_tiles);
        // ---- pass 1: scatter tiles [tile_begin, tile_end) (rows of full buckets go to the global table directly)
        {
            const uint64_t rows_here = (uint64_t)(std::min<int64_t>(n, tile_end * TILE) - tile_begin * TILE);
            uint64_t cap = h->capacity;
            while (cap / 4 < std::min(margin1, rows_here) || (uint64_t)h->ngroups_host >= cap / 2) cap <<= 1;
            if (cap != h->capacity && (st = table_grow(ctx, h, cap)) != KQ_OK) return done(st);
        }
        fill_common_args(h, A);
        A.n = n; A.ntiles = tile_end; A.tile_begin = tile_begin;
        A.stop_threshold = h->capacity / 2;
        A.part_scratch = scratch; A.part_counts = counts; A.nparts = nparts; A.part_ncta = grid; A.part_cap = part_cap; A.part_begin = 0;
        A.part_log2 = 0; while ((1 << A.part_log2) < nparts) A.part_log2++;
        if (cudaMemsetAsync(h->d_counters + 1, 0, 8, ctx->stream) != cudaSuccess || cudaMemsetAsync(counts, 0, buckets * 4, ctx->stream) != cudaSuccess)
            return done(kq_cuda_fail(ctx, cudaGetLastError(), "cudaMemsetAsync"));
        if (ring_check) { cudaMemsetAsync(counts + buckets + 16, 0, (size_t)ntiles * 4, ctx->stream); A.trace = reinterpret_cast<unsigned long long*>(counts + buckets + 16); }
        void* kargs[] = {&A};
        cudaError_t ce = cudaLaunchKernel(k_scatter, dim3(grid), dim3(THREADS), kargs, (size_t)A.smem_bytes, ctx->stream);
        if (ce != cudaSuccess) return done(kq_cuda_fail(ctx, ce, "cudaLaunchKernel(kq_hash_aggregate, partition scatter)"));
        ctx->launches++;
        cudaMemsetAsync(d_max, 0, 8, ctx->stream);
        k_max_u32<<<small_grid(ctx, buckets), 256, 0, ctx->stream>>>(counts, buckets, d_max);
        if ((st = launch_check(ctx, "k_max_u32")) != KQ_OK) return done(st);
        uint64_t c[6];
        if ((st = read_counters(ctx, h, c)) != KQ_OK) return done(st);
        if ((uint32_t)c[4] != 0) return done(kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "aggregation table overflow"));
        h->ngroups_host = (int64_t)c[0];
        const int64_t taken = std::min<int64_t>((int64_t)(uint32_t)c[1], tile_end - tile_begin);
        const uint64_t max_bucket = c[3];
        if (ring_check) {
            std::vector<unsigned int> seen((size_t)ntiles);
            cudaMemcpy(seen.data(), counts + buckets + 16, (size_t)ntiles * 4, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int64_t t = tile_begin; t < tile_end; t++)
                if ((int)seen[(size_t)t] != geo.warps && bad++ < 20) fprintf(stderr, "kq ring check: tile %lld consumed by %u warps (want %d)\n", (long long)t, seen[(size_t)t], geo.warps);
            fprintf(stderr, "kq ring check: %d tiles with a wrong consumer count\n", bad);
        }
        if (getenv("KQ_TRACE_AGG"))
            fprintf(stderr, "kq partitioned pass 1: tiles [%lld, %lld) taken %lld, nparts %d grid %d bucket cap %d (max fill %llu), table cap %llu, groups after pass 1 %llu\n", (long long)tile_begin,
                    (long long)tile_end, (long long)taken, nparts, grid, part_cap, (unsigned long long)max_bucket, (unsigned long long)h->capacity, (unsigned long long)c[0]);
        // ---- pass 2: reduce the partitions; a block in flight may add up to one partition's rows as new groups
        const uint64_t margin2 = (uint64_t)std::min<int64_t>(ctx->sm_count, nparts) * std::max<uint64_t>(1, max_bucket * (uint64_t)grid);
        int part_begin = 0;
        while (part_begin < nparts) {
            uint64_t cap = h->capacity;
            while (cap / 4 < margin2 || (uint64_t)h->ngroups_host >= cap / 2) cap <<= 1;
            if (cap != h->capacity && (st = table_grow(ctx, h, cap)) != KQ_OK) return done(st);
            fill_common_args(h, A);
            A.stop_threshold = h->capacity / 2;
            A.part_begin = part_begin;
            if (cudaMemsetAsync(h->d_counters + 1, 0, 8, ctx->stream) != cudaSuccess) return done(kq_cuda_fail(ctx, cudaGetLastError(), "cudaMemsetAsync"));
            ce = cudaLaunchKernel(k_reduce, dim3(std::min(nparts - part_begin, ctx->sm_count)), dim3(512), kargs, (size_t)reduce_smem, ctx->stream);
            if (ce != cudaSuccess) return done(kq_cuda_fail(ctx, ce, "cudaLaunchKernel(kq_agg_partition_reduce)"));
            ctx->launches++;
            if ((st = read_counters(ctx, h, c)) != KQ_OK) return done(st);
            if ((uint32_t)c[4] != 0) return done(kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "aggregation table overflow"));
            h->ngroups_host = (int64_t)c[0];
            if (getenv("KQ_TRACE_AGG"))
                fprintf(stderr, "kq partitioned pass 2: from partition %d, tickets %u, table cap %llu, groups %llu\n", part_begin, (unsigned)c[1], (unsigned long long)h->capacity, (unsigned long long)c[0]);
            part_begin += (int)std::min<int64_t>((int64_t)(uint32_t)c[1], nparts - part_begin);
        }
        tile_begin += taken;
        if (taken == 0 && (st = table_grow(ctx, h, h->capacity * 2)) != KQ_OK) return done(st);   // pass 1 could not start (table at its threshold)
    }
    return done(KQ_OK);
}

// ---- the low-cardinality path (kq_k_agg_fe.cuh) ---------------------------------------------------------------------------
// The global table is sized OPTIMISTICALLY from the planner's hint (a few KB for a 50-group query: nothing to zero, scan
// or collect beyond the groups themselves). Should the hint be wrong by so much that the table fills up, the kernel
// raises the overflow flag instead of spinning; the launch is discarded, the table restored from a snapshot taken
// before it (only a non-empty table needs one) and the batch is redone with the conservative capacity rule.
static int hashagg_update_fe(kq_ctx* ctx, kq_hashagg* h, kq_batch* input, int64_t n) {
    AggArgs A;
    memset(&A, 0, sizeof A);
    std::string defines, gen;
    const bool timing = getenv("KQ_TIME_AGG") != nullptr;        // host-side cost of one update, per phase (stderr)
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return (long long)std::chrono::duration_cast<std::chrono::microseconds>(b - a).count(); };
    const auto t0 = now();
    KQ_RET(plan_agg(ctx, h, input, ctx->max_smem_optin, A, &defines, &gen, 3));
    const auto t1 = now();
    if (n == 0) return KQ_OK;
    const AggGeometry geo{A.geo_r, A.geo_warps};
    const int TILE = geo.tile(), THREADS = geo.threads() + 32 * (A.geo_service - 1);       // + the housekeeping warp of queries with MIN/MAX
    A.n = n; A.ntiles = (n + TILE - 1) / TILE;
    void* kernel = nullptr;
    KQ_RET(kq_jit_kernel(ctx, defines, gen, KQ_SKEL_AGG_FE, "kq_group_aggregate", A.smem_bytes, &kernel));
    const auto t2 = now();
    const int grid = (int)std::min<int64_t>(A.ntiles, (int64_t)ctx->sm_count * std::max(1, A.geo_ctas));
    // rows that may still create groups after a block has decided to continue (conservative rule only)
    const uint64_t margin = (uint64_t)grid * ((uint64_t)(A.sp.nstages + 1) * TILE + FE_MAX_GROUPS);

    // Block b owns tiles b, b + grid, ...; `progress` (device, per block) survives relaunches over this batch, `done` counts
    // the tiles finished so far.
    if (grid > 768) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "more than 768 blocks");
    unsigned int* d_progress = reinterpret_cast<unsigned int*>(h->d_counters + 128);
    unsigned int* d_progress_snap = reinterpret_cast<unsigned int*>(h->d_counters + 512);
    KQ_CUDA(ctx, cudaMemsetAsync(d_progress, 0, (size_t)grid * 4, ctx->stream));
    KQ_CUDA(ctx, cudaMemsetAsync(h->d_counters + 1, 0, 8, ctx->stream));
    int64_t done = 0;
    const int64_t tile_begin = 0;
    while (done < A.ntiles) {
        const uint64_t remaining_rows = (uint64_t)n;              // blocks advance independently: no prefix of the batch is finished until all of it is
        uint64_t cap = h->capacity;
        bool unthrottled = false;
        const bool optimistic = !h->pessimistic;
        if (optimistic) {
            // room for 8x the groups known or promised; the kernel stops taking tiles past half full, overflow is caught
            const uint64_t g = (uint64_t)std::max<int64_t>(std::max<int64_t>(h->expected_groups, h->ngroups_host), 64);
            while (cap < 8 * g) cap <<= 1;
        } else {
            const uint64_t inflight = std::min(margin, remaining_rows);
            while (true) {
                if ((uint64_t)h->ngroups_host + remaining_rows <= cap / 4 * 3) { unthrottled = true; break; }
                if (cap / 4 >= inflight && (uint64_t)h->ngroups_host < cap / 2) break;
                cap <<= 1;
            }
        }
        if (cap != h->capacity) KQ_RET(table_grow(ctx, h, cap));
        const bool need_snapshot = optimistic && h->ngroups_host > 0;
        if (need_snapshot) {
            if (h->snapshot_cap != h->capacity) {
                kq_dev_free(ctx, h->snapshot); h->snapshot = nullptr; h->snapshot_cap = 0;
                KQ_RET(kq_dev_alloc(ctx, (size_t)(h->capacity + 1) * h->stride * 8, (void**)&h->snapshot));
                h->snapshot_cap = h->capacity;
            }
            KQ_CUDA(ctx, cudaMemcpyAsync(h->snapshot, h->table, (size_t)h->capacity * h->stride * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        if (optimistic) KQ_CUDA(ctx, cudaMemcpyAsync(d_progress_snap, d_progress, (size_t)grid * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        const int64_t groups_before = h->ngroups_host;
        fill_common_args(h, A);
        A.tile_begin = tile_begin;
        A.progress = d_progress;
        A.stop_threshold = unthrottled ? ~0ULL : h->capacity / 2;
        void* kargs[] = {&A};
        KQ_CUDA(ctx, cudaLaunchKernel(kernel, dim3(grid), dim3(THREADS), kargs, (size_t)A.smem_bytes, ctx->stream));
        ctx->launches++;
        uint64_t c[6];
        {
            const auto t3 = now();
            const int rst = read_counters(ctx, h, c);
            if (timing) fprintf(stderr, "kq fe update: plan %lld us, kernel lookup %lld us, table + launch %lld us, wait for counters %lld us\n", us(t0, t1), us(t1, t2), us(t2, t3), us(t3, now()));
            if (getenv("KQ_TRACE_AGG"))
                fprintf(stderr, "kq fe launch: done %lld of %lld tiles, grid %d cap %llu optimistic %d -> st %d groups %llu done %u overflow %u err 0x%x\n", (long long)done,
                        (long long)A.ntiles, grid, (unsigned long long)h->capacity, (int)optimistic, rst, (unsigned long long)c[0], (unsigned)c[1], (unsigned)c[4], (unsigned)c[5]);
            KQ_RET(rst);
        }
        if ((uint32_t)c[4] != 0) {
            // the table filled up (the hint was far off): discard this launch and redo its tiles with a table sized for the rows in flight
            if (!optimistic) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "aggregation table overflow");
            if (need_snapshot) KQ_CUDA(ctx, cudaMemcpyAsync(h->table, h->snapshot, (size_t)h->capacity * h->stride * 8, cudaMemcpyDeviceToDevice, ctx->stream));
            else KQ_CUDA(ctx, cudaMemsetAsync(h->table, 0, (size_t)(h->capacity + 1) * h->stride * 8, ctx->stream));
            const unsigned long long restore[2] = {(unsigned long long)groups_before, (unsigned long long)done};
            KQ_CUDA(ctx, cudaMemcpyAsync(h->d_counters, restore, 16, cudaMemcpyHostToDevice, ctx->stream));
            KQ_CUDA(ctx, cudaMemcpyAsync(d_progress, d_progress_snap, (size_t)grid * 4, cudaMemcpyDeviceToDevice, ctx->stream));
            KQ_CUDA(ctx, cudaMemsetAsync(h->d_counters + 4, 0, 8, ctx->stream));
            KQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));          // `restore` lives on this stack frame
            h->pessimistic = true;
            continue;
        }
        h->ngroups_host = (int64_t)c[0];
        const int64_t now_done = (int64_t)(uint32_t)c[1];
        if (now_done < A.ntiles) {
            // stopped early: the table crossed half full. (No progress at all with a table that is not at its threshold would repeat forever.)
            if (now_done == done && (uint64_t)h->ngroups_host <= h->capacity / 2) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "aggregate kernel made no progress");
            KQ_RET(table_grow(ctx, h, h->capacity * 4));
        }
        done = now_done;
    }
    return KQ_OK;
}

// Utf8 group keys longer than 7 bytes: make sure the key heap can take every long string this batch may add. The column's
// length statistics are computed once per column (a pass over its offsets) and kept with it.
static int key_heap_reserve(kq_ctx* ctx, kq_hashagg* h, uint64_t add_entries, uint64_t add_bytes);
static int ensure_key_heap(kq_ctx* ctx, kq_hashagg* h, kq_batch* input) {
    uint64_t n_long = 0, bytes_long = 0, max_len = 0;
    for (kq_expr* g : h->groups) {
        const int bc = KqCodegen::bare_column(g);
        if (bc < 0 || bc >= (int)input->cols.size()) continue;
        kq_col* c = input->cols[(size_t)bc];
        if (c->type != KQ_UTF8) continue;
        if (c->utf8_max_len < 0) {
            int64_t rows; KQ_RET(kq_col_resolve_rows(ctx, c, &rows));
            unsigned long long* d = nullptr;
            KQ_RET(kq_dev_alloc(ctx, 32, (void**)&d));
            cudaMemsetAsync(d, 0, 32, ctx->stream);
            if (rows > 0) { k_utf8_len_stats<<<small_grid(ctx, (uint64_t)rows), 256, 0, ctx->stream>>>(c->offsets, rows, d); ctx->launches++; }
            uint64_t st3[3];
            const int st = kq_read_u64(ctx, d, 3, st3);
            kq_dev_free(ctx, d);
            KQ_RET(st);
            c->utf8_max_len = (int64_t)st3[0]; c->utf8_n_long = (int64_t)st3[1]; c->utf8_bytes_long = (int64_t)st3[2];
        }
        if (c->utf8_max_len > 7) { n_long += (uint64_t)c->utf8_n_long; bytes_long += (uint64_t)c->utf8_bytes_long; max_len = std::max<uint64_t>(max_len, (uint64_t)c->utf8_max_len); }
    }
    if (n_long == 0) return KQ_OK;
    // every long row could be a new string; the planner's hint, when there is one, bounds that far below the row count
    const uint64_t distinct_cap = h->expected_groups > 0 ? std::max<uint64_t>(4 * (uint64_t)h->expected_groups, 1ULL << 16) : n_long;
    const uint64_t add = std::min<uint64_t>(std::min(n_long, distinct_cap), 1ULL << 27);
    return key_heap_reserve(ctx, h, add, std::min<uint64_t>(bytes_long + 8 * n_long, add * (max_len + 8)));
}

// Room for `add_entries` more strings of `add_bytes` bytes in total (8-byte padded) in the aggregate's key heap.
static int key_heap_reserve(kq_ctx* ctx, kq_hashagg* h, uint64_t add_entries, uint64_t add_bytes) {
    const uint64_t need_entries = h->heap_entries + add_entries;
    const uint64_t need_bytes = h->heap_bytes + add_bytes;
    uint64_t cap = h->heap.tab ? h->heap.cap_mask + 1 : 1024;
    while (cap < 2 * need_entries) cap <<= 1;
    uint64_t bcap = h->heap.bytes ? h->heap.bytes_cap : 4096;
    while (bcap < need_bytes) bcap <<= 1;
    if (h->heap.tab && cap == h->heap.cap_mask + 1 && bcap == h->heap.bytes_cap) return KQ_OK;
    unsigned long long* tab = nullptr; uint8_t* bytes = nullptr;
    KQ_RET(kq_dev_alloc(ctx, (size_t)cap * 16, (void**)&tab));
    int st = kq_dev_alloc(ctx, (size_t)bcap + 16, (void**)&bytes);
    if (st != KQ_OK) { kq_dev_free(ctx, tab); return st; }
    cudaMemsetAsync(tab, 0, (size_t)cap * 16, ctx->stream);
    if (h->heap.tab) {
        k_heap_rehash<<<small_grid(ctx, h->heap.cap_mask + 1), 256, 0, ctx->stream>>>(h->heap.tab, h->heap.cap_mask + 1, tab, cap - 1);
        ctx->launches++;
        cudaMemcpyAsync(bytes, h->heap.bytes, (size_t)std::min<uint64_t>(h->heap.bytes_cap, h->heap_bytes + 8), cudaMemcpyDeviceToDevice, ctx->stream);
        kq_dev_free(ctx, h->heap.tab); kq_dev_free(ctx, h->heap.bytes);
    }
    h->heap.tab = tab; h->heap.cap_mask = cap - 1; h->heap.bytes = bytes; h->heap.bytes_cap = bcap;
    h->heap.used = h->d_counters + 6;
    return KQ_OK;
}

extern "C" {

int kq_hashagg_update(kq_ctx* ctx, kq_hashagg* h, kq_batch* input) {
    if (!ctx || !h || !input) return KQ_ERR_ILLEGAL_ARGUMENT;
    cudaSetDevice(ctx->device);
    KQ_RET(sticky_error(ctx, h));
    int64_t n; KQ_RET(kq_batch_resolve_rows(ctx, input, &n));
    for (kq_col* c : input->cols) KQ_RET(kq_col_resolve_rows(ctx, c, nullptr));
    KQ_RET(ensure_key_heap(ctx, h, input));

    AggArgs A;
    memset(&A, 0, sizeof A);
    std::string defines, gen;
    // High cardinality (planner hint, or learnt from earlier batches): the partitioned path.
    const int64_t g_est = std::max<int64_t>(h->expected_groups, h->ngroups_host);
    const int64_t part_min_groups = getenv("KQ_PART_MIN_GROUPS") ? atoll(getenv("KQ_PART_MIN_GROUPS")) : PART_MIN_GROUPS;      // tuning experiments
    if (g_est >= part_min_groups && n >= PART_MIN_ROWS && !getenv("KQ_NO_PARTITION")) return hashagg_update_partitioned(ctx, h, input, n, g_est);
    // Low cardinality (planner hint / what earlier batches showed fits the CTA directory, or nothing is known yet): kq_k_agg_fe.cuh
    if (g_est <= FE_MAX_GROUPS && !getenv("KQ_NO_FE")) return hashagg_update_fe(ctx, h, input, n);
    // Mid cardinality (more groups than the lane-private front end holds, few enough for one shared-memory table per block)
    int mode = 0;
    if (g_est > FE_MAX_GROUPS && !getenv("KQ_NO_SHARED_TABLE")) {
        AggArgs P;
        memset(&P, 0, sizeof P);
        std::string d2, g2;
        if (plan_agg(ctx, h, input, ctx->max_smem_optin, P, &d2, &g2, 2) == KQ_OK && g_est <= (int64_t)(P.part_groups - 32 * (P.geo_warps + 1)) * 9 / 10) mode = 2;
    }
    KQ_RET(plan_agg(ctx, h, input, ctx->max_smem_optin, A, &defines, &gen, mode));
    if (n == 0) return KQ_OK;
    const AggGeometry geo{A.geo_r, A.geo_warps};
    const int TILE = geo.tile(), THREADS = geo.threads();
    A.n = n; A.ntiles = (n + TILE - 1) / TILE;
    void* kernel = nullptr;
    KQ_RET(kq_jit_kernel(ctx, defines, gen, KQ_SKEL_AGG, "kq_hash_aggregate", A.smem_bytes, &kernel));
    int grid = (int)std::min<int64_t>(A.ntiles, (int64_t)ctx->sm_count);
    // rows that may still create groups after a block has decided to continue: one tile per resident
    // block plus its front end
    const uint64_t margin = (uint64_t)grid * ((uint64_t)(A.sp.nstages + 1) * TILE + FE_MAX_GROUPS);     // + the ticket requested one step early

    int64_t tile_begin = 0;
    while (tile_begin < A.ntiles) {
        // Capacity rule. If every remaining row could become a group and the table would still be
        // below 3/4 full, run unthrottled. Otherwise the service warps stop taking tiles once the
        // table is half full; tiles already in flight may add up to `inflight` more groups, which
        // must fit in the next quarter.
        const uint64_t remaining_rows = (uint64_t)(n - tile_begin * TILE);
        const uint64_t inflight = std::min(margin, remaining_rows);
        uint64_t cap = h->capacity;
        bool unthrottled = false;
        while (true) {
            if ((uint64_t)h->ngroups_host + remaining_rows <= cap / 4 * 3) { unthrottled = true; break; }
            if (cap / 4 >= inflight && (uint64_t)h->ngroups_host < cap / 2) break;
            cap <<= 1;
        }
        if (cap != h->capacity) KQ_RET(table_grow(ctx, h, cap));
        fill_common_args(h, A);
        A.tile_begin = tile_begin;
        A.stop_threshold = unthrottled ? ~0ULL : h->capacity / 2;
        KQ_CUDA(ctx, cudaMemsetAsync(h->d_counters + 1, 0, 8, ctx->stream));
        void* kargs[] = {&A};
        KQ_CUDA(ctx, cudaLaunchKernel(kernel, dim3(grid), dim3(THREADS), kargs, (size_t)A.smem_bytes, ctx->stream));
        ctx->launches++;
        uint64_t c[6];
        KQ_RET(read_counters(ctx, h, c));
        if ((uint32_t)c[4] != 0) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "aggregation table overflow");      // cannot happen under the capacity rule above
        h->ngroups_host = (int64_t)c[0];
        int64_t taken = (int64_t)(uint32_t)c[1];
        tile_begin += std::min<int64_t>(taken, A.ntiles - tile_begin);
        if (tile_begin < A.ntiles) {
            // stopped early: the table crossed half full. Grow 4x and continue with the next tile.
            KQ_RET(table_grow(ctx, h, h->capacity * 4));
        }
    }
    return KQ_OK;
}

int kq_explain_hashagg(kq_expr* pred, kq_expr* const* group_exprs, int ngroup, const int* agg_kinds, kq_expr* const* agg_inputs, int nagg,
                       int ncols, const int* types, const int* nullable, int compile, char* source, size_t source_cap) {
    if (ncols < 0 || ngroup < 0 || nagg < 0 || (ncols > 0 && !types)) return KQ_ERR_ILLEGAL_ARGUMENT;
    kq_ctx fake;
    KqSchemaBatch sb(ncols, types, nullable);
    kq_hashagg* h = nullptr;
    const char* eg = getenv("KQ_EXPLAIN_GROUPS");        // tuning aid: the cardinality hint shapes the geometry
    int st = hashagg_new(&fake, pred, group_exprs, ngroup, agg_kinds, agg_inputs, nagg, eg ? atoll(eg) : 0, &h);
    AggArgs A;
    memset(&A, 0, sizeof A);
    std::string defines, gen;
    const char* ep = getenv("KQ_EXPLAIN_PARTS");          // tuning aid: compile the partitioned path's kernels instead (0: the block-shared table mode)
    // default: what kq_hashagg_update would run for this hint — the low-cardinality kernel up to FE_MAX_GROUPS groups
    const int64_t egv = eg ? atoll(eg) : 0;
    const int mode = ep ? (atoi(ep) > 0 ? 1 : 2) : ((egv <= FE_MAX_GROUPS && !getenv("KQ_NO_FE")) ? 3 : 0);
    if (st == KQ_OK) st = plan_agg(&fake, h, &sb.batch, 232448, A, &defines, &gen, mode, ep ? atoi(ep) : 0);
    if (st == KQ_OK && compile) st = kq_jit_compile_only(&fake, defines, gen, mode == 3 ? KQ_SKEL_AGG_FE : KQ_SKEL_AGG);
    kq_copy_text(st == KQ_OK ? defines + gen : fake.last_error, source, source_cap);
    if (h) hashagg_delete_host(h);
    return st;
}

int kq_hashagg_num_groups(kq_ctx* ctx, kq_hashagg* h, int64_t* n) {
    if (!ctx || !h || !n) return KQ_ERR_ILLEGAL_ARGUMENT;
    KQ_RET(sticky_error(ctx, h));
    uint64_t c[6]; KQ_RET(read_counters(ctx, h, c));
    h->ngroups_host = (int64_t)c[0]; h->groups_exact = true;
    *n = (int64_t)c[0];
    return KQ_OK;
}

// One output batch (Main.kt:635-650). ONE host synchronisation: the output columns are allocated for the group count the
// host already knows (exact after an update, an upper bound after a merge), every kernel is queued, then the row count,
// the Utf8 byte totals and this aggregate's error word come back in a single copy.
int kq_hashagg_finalize(kq_ctx* ctx, kq_hashagg* h, kq_batch** out) {
    if (!ctx || !h || !out) return KQ_ERR_ILLEGAL_ARGUMENT;
    cudaSetDevice(ctx->device);
    if (!h->typed) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "finalize before any update: output types unknown");
    KQ_RET(sticky_error(ctx, h));
    const int64_t G = h->ngroups_host;             // rows to allocate (>= the rows that will be written)
    if (G > 2147483647LL) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "more than 2^31-1 groups in one output batch");
    FinArgs F;
    memset(&F, 0, sizeof F);
    F.table = h->table; F.cap = h->capacity; F.stride = h->stride;
    F.nkeys = (int)h->groups.size(); F.naggs = (int)h->agg_kinds.size();
    std::vector<kq_col*> cols;
    std::vector<uint64_t*> packed((size_t)F.nkeys, nullptr);
    std::vector<unsigned long long*> scratches;
    int st = KQ_OK;
    auto fail = [&](int s) {
        for (kq_col* c : cols) kq_column_free(c);
        for (uint64_t* p : packed) kq_dev_free(ctx, p);
        for (unsigned long long* p : scratches) kq_dev_free(ctx, p);
        return s;
    };
    for (int k = 0; k < F.nkeys; k++) {
        kq_col* c = nullptr;
        int t = h->key_types[(size_t)k];
        // Utf8 bytes: 7 per group bound the packed keys; with interned (long) keys in play the exact total is read back first
        const bool long_keys = h->heap.tab != nullptr;
        if ((st = kq_col_new(ctx, t, G, true, t == KQ_UTF8 ? (long_keys ? 0 : G * 7) : 0, &c)) != KQ_OK) return fail(st);
        cols.push_back(c);
        cudaMemsetAsync(c->validity, 0, (size_t)((G + 63) / 64) * 8, ctx->stream);
        if (t == KQ_BOOL) cudaMemsetAsync(c->data, 0, (size_t)((G + 63) / 64) * 8, ctx->stream);
        F.keys[k].validity = c->validity; F.keys[k].type = t;
        if (t == KQ_UTF8) {
            if ((st = kq_dev_alloc(ctx, (size_t)std::max<int64_t>(G, 1) * 8, (void**)&packed[(size_t)k])) != KQ_OK) return fail(st);
            F.keys[k].data = packed[(size_t)k];
        } else F.keys[k].data = c->data;
    }
    for (int a = 0; a < F.naggs; a++) {
        int kind = h->agg_kinds[(size_t)a], i = h->agg_input_of[(size_t)a];
        int in_t = h->input_types[(size_t)i];
        int out_t = kind == KQ_AGG_COUNT ? KQ_I64 : in_t;
        kq_col* c = nullptr;
        if ((st = kq_col_new(ctx, out_t, G, kind != KQ_AGG_COUNT, 0, &c)) != KQ_OK) return fail(st);
        cols.push_back(c);
        if (c->validity) cudaMemsetAsync(c->validity, 0, (size_t)((G + 63) / 64) * 8, ctx->stream);
        FinAgg& o = F.aggs[a];
        o.data = c->data; o.validity = c->validity; o.kind = kind; o.out_type = out_t;
        o.nn_word = h->in[i].rec_nn;
        o.word = kind == KQ_AGG_SUM ? h->in[i].rec_sum : (kind == KQ_AGG_MIN ? h->in[i].rec_min : (kind == KQ_AGG_MAX ? h->in[i].rec_max : h->in[i].rec_nn));
        o.is_int = in_t != KQ_F64;
    }
    unsigned long long* d_pos = h->d_counters + 2;
    cudaMemsetAsync(d_pos, 0, 8, ctx->stream);
    F.pos = d_pos;
    F.max_rows = (unsigned long long)G;
    F.err = (uint32_t*)(h->d_counters + 5);
    if (G > 0) {
        if (h->capacity <= (1u << 18)) k_finalize<1><<<small_grid(ctx, h->capacity), 256, 0, ctx->stream>>>(F);
        else k_finalize<32><<<small_grid(ctx, h->capacity), 256, 0, ctx->stream>>>(F);
        if ((st = launch_check(ctx, "k_finalize")) != KQ_OK) return fail(st);
    }
    // Utf8 keys: packed words -> offsets (device-wide scan over d_pos rows) + bytes; the byte total of key k lands in d_counters[8 + k]
    for (int k = 0; k < F.nkeys; k++) {
        if (h->key_types[(size_t)k] != KQ_UTF8) continue;
        kq_col* c = cols[(size_t)k];
        int64_t ntiles = (G + SCAN_TILE - 1) / SCAN_TILE + 1;
        unsigned long long* scratch = nullptr;
        if ((st = kq_dev_alloc(ctx, (size_t)(ntiles + 4) * 8, (void**)&scratch)) != KQ_OK) return fail(st);
        scratches.push_back(scratch);
        cudaMemsetAsync(scratch, 0, (size_t)(ntiles + 4) * 8, ctx->stream);
        PackedLen pl{packed[(size_t)k], c->validity, h->heap};
        int sg = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, (int64_t)ctx->sm_count * 4));
        k_exclusive_offsets<PackedLen><<<sg, 256, 0, ctx->stream>>>(pl, d_pos, c->offsets, scratch + 4, (unsigned int*)scratch, h->d_counters + 8 + k);
        if ((st = launch_check(ctx, "k_exclusive_offsets")) != KQ_OK) return fail(st);
        if (G > 0 && h->heap.tab == nullptr) {
            k_unpack_utf8<<<small_grid(ctx, (uint64_t)G), 256, 0, ctx->stream>>>(packed[(size_t)k], c->offsets, d_pos, (uint8_t*)c->data, h->heap);
            if ((st = launch_check(ctx, "k_unpack_utf8")) != KQ_OK) return fail(st);
        }
    }
    uint64_t c16[16];
    if ((st = kq_read_u64(ctx, h->d_counters, 8 + MAX_KEYS, c16)) != KQ_OK) return fail(st);
    if ((uint32_t)c16[5]) { h->dev_err |= (uint32_t)c16[5]; return fail(sticky_error(ctx, h)); }
    h->heap_entries = c16[6]; h->heap_bytes = c16[7];
    const int64_t rows = (int64_t)c16[2];
    h->ngroups_host = (int64_t)c16[0]; h->groups_exact = true;
    for (int k = 0; k < F.nkeys; k++) {
        kq_col* c = cols[(size_t)k];
        c->n = rows;
        if (h->key_types[(size_t)k] != KQ_UTF8) continue;
        c->data_bytes = (int64_t)c16[8 + k];
        if (h->heap.tab != nullptr) {
            // interned keys: a group's string can be of any length and several groups may share one (multi-key GROUP BYs), so
            // the bytes are allocated for the total the offsets scan just produced (the one extra round trip of this path)
            if (c->data_bytes > 2147483647LL) return fail(kq_fail(ctx, KQ_ERR_UNSUPPORTED, "more than 2^31-1 bytes of Utf8 group keys in one output batch"));
            kq_dev_free(ctx, c->data); c->data = nullptr;
            if ((st = kq_dev_alloc(ctx, (size_t)std::max<int64_t>(c->data_bytes, 16), &c->data)) != KQ_OK) return fail(st);
            if (rows > 0) {
                k_unpack_utf8<<<small_grid(ctx, (uint64_t)rows), 256, 0, ctx->stream>>>(packed[(size_t)k], c->offsets, d_pos, (uint8_t*)c->data, h->heap);
                if ((st = launch_check(ctx, "k_unpack_utf8")) != KQ_OK) return fail(st);
            }
        }
    }
    for (size_t a = (size_t)F.nkeys; a < cols.size(); a++) cols[a]->n = rows;
    for (uint64_t*& p : packed) { kq_dev_free(ctx, p); p = nullptr; }
    for (unsigned long long*& p : scratches) { kq_dev_free(ctx, p); p = nullptr; }
    st = kq_batch_create(ctx, cols.data(), (int)cols.size(), rows, out);
    for (kq_col* c : cols) kq_column_free(c);
    cols.clear();
    return st;
}

// ---- multi-GPU merge (partial -> merge of main(), Main.kt:1309-1325) --------------------------------------------------
// Word class of a record word: which reduction merges it.
enum { WC_SKIP = 0, WC_SUM_U64, WC_SUM_F64, WC_MIN, WC_MAX };
static void word_classes(kq_hashagg* h, int* cls) {
    for (int w = 0; w < h->stride; w++) cls[w] = WC_SKIP;
    for (size_t i = 0; i < h->inputs.size(); i++) {
        const AggInput& d = h->in[i];
        cls[d.rec_nn] = WC_SUM_U64;
        if (d.rec_sum >= 0) cls[d.rec_sum] = (d.flags & F_INT) ? WC_SUM_U64 : WC_SUM_F64;
        if (d.rec_min >= 0) cls[d.rec_min] = WC_MIN;
        if (d.rec_max >= 0) cls[d.rec_max] = WC_MAX;
    }
}

static int refresh_group_count(kq_ctx* ctx, kq_hashagg* h) {
    uint64_t c; KQ_RET(kq_read_u64(ctx, h->d_counters, 1, &c));
    h->ngroups_host = (int64_t)c;
    return KQ_OK;
}

int kq_hashagg_merge_allreduce(kq_ctx* ctx, kq_hashagg* h) {
    if (!ctx || !h) return KQ_ERR_ILLEGAL_ARGUMENT;
    if (!ctx->comm || ctx->nranks <= 1) return KQ_OK;
    KqNccl* N = kq_nccl(ctx);
    if (!N) return KQ_ERR_NCCL;
    cudaSetDevice(ctx->device);
    ncclComm_t comm = (ncclComm_t)ctx->comm;
    const int nr = ctx->nranks, me = ctx->rank, stride = h->stride, nkeys = (int)h->groups.size();
    (void)me;
    // (a sticky error from an earlier call does not return here: the peers are, or will be, in the all-gather, so this
    //  rank joins it and its status word — d_counters[5] still holds the bits — tells everybody)

    // 0. Small unions (BASELINE configs 3 and 5: tens of groups) are latency-bound, so the all-reduce is done the
    //    latency-optimal way: ONE fixed-size all-gather carrying every rank's group count, status and (when they fit) its
    //    partial records; then every rank rebuilds its table from the gathered partials rank by rank (k_merge_small:
    //    the order of every Float64 addition is fixed, all ranks end bit-identical). One host synchronisation — it also
    //    is where all ranks agree, on the SAME gathered data, whether anybody failed and which path to take, so no rank
    //    leaves while its peers wait in a collective. No allocation after the aggregate's first merge.
    const uint64_t room = (h->expected_groups > 0 && h->expected_groups <= 96) ? 128 : SMALL_MERGE_MAX;       // same on every rank: the plan's hint
    {
        const uint64_t blockw = MERGE_HDR + room * (uint64_t)stride;
        if (h->merge_words != blockw) {
            kq_dev_free(ctx, h->merge_mine); kq_dev_free(ctx, h->merge_all);
            h->merge_mine = h->merge_all = nullptr; h->merge_words = 0;
            // an allocation failure here cannot be reported before the collective without leaving the peers hanging:
            // nothing else has been allocated yet for this merge, so fail the rank hard (the job's launcher aborts the others)
            int st0 = kq_dev_alloc(ctx, (size_t)blockw * 8, (void**)&h->merge_mine);
            if (st0 == KQ_OK) st0 = kq_dev_alloc(ctx, (size_t)blockw * 8 * nr + 4 * 64 * 8, (void**)&h->merge_all);
            if (st0 != KQ_OK) return st0;
            h->merge_words = blockw;
        }
        uint64_t* mine = h->merge_mine; uint64_t* all = h->merge_all;
        unsigned long long* d_heads = reinterpret_cast<unsigned long long*>(all + blockw * nr);
        if (nr > 64) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "more than 64 ranks");
        cudaMemsetAsync(mine, 0, MERGE_HDR * 8, ctx->stream);
        k_collect_small<<<small_grid(ctx, h->capacity), 256, 0, ctx->stream>>>(h->table, h->capacity, stride, mine, room, (const uint32_t*)(h->d_counters + 5));
        k_collect_small_done<<<1, 1, 0, ctx->stream>>>(mine, h->heap.tab ? h->heap.used : nullptr);
        ctx->launches += 2;
        ncclResult_t r0 = N->AllGather(mine, all, (size_t)blockw, ncclUint64, comm, ctx->stream);
        if (r0 != ncclSuccess) return kq_nccl_fail(ctx, r0, "ncclAllGather");
        k_merge_heads<<<1, 64, 0, ctx->stream>>>(all, blockw, nr, d_heads);
        ctx->launches++;
        uint64_t heads[256];
        for (int off = 0; off < 4 * nr; off += 64) KQ_RET(kq_read_u64(ctx, d_heads + off, std::min(64, 4 * nr - off), heads + off));
        uint64_t total = 0, maxn = 0, str_entries = 0, str_bytes = 0; uint32_t status = 0;
        for (int i = 0; i < nr; i++) {
            total += heads[4 * i]; maxn = std::max(maxn, heads[4 * i]); status |= (uint32_t)heads[4 * i + 1];
            if (i != me) { str_entries += heads[4 * i + 2]; str_bytes += heads[4 * i + 3] + 8 * heads[4 * i + 2]; }
        }
        const bool long_keys = str_entries > 0 || heads[4 * me + 2] > 0;          // some rank interned long Utf8 group keys
        if (status) { h->dev_err |= status; return sticky_error(ctx, h); }       // some rank's kernels raised an error: every rank reports it
        if (total == 0) return KQ_OK;
        const bool small = maxn <= room && !getenv("KQ_NO_SMALL_MERGE");
        if (long_keys && !small)       // decided on the same gathered data on every rank
            return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "merging more than %llu partial groups per rank with Utf8 group keys longer than 7 bytes is not implemented", (unsigned long long)room);
        if (long_keys) {
            // the strings behind the long keys go along: every rank ends with all groups, so it needs all their bytes
            unsigned long long *smine = nullptr, *sall = nullptr;
            int st2 = kq_dev_alloc(ctx, (size_t)STRBLK_WORDS * 8, (void**)&smine);
            if (st2 == KQ_OK) st2 = kq_dev_alloc(ctx, (size_t)STRBLK_WORDS * 8 * nr, (void**)&sall);
            if (st2 == KQ_OK) st2 = key_heap_reserve(ctx, h, str_entries, str_bytes);
            if (st2 != KQ_OK) { fprintf(stderr, "kqgpu: rank %d cannot allocate the key exchange buffers inside a collective merge; aborting\n", me); abort(); }
            cudaMemsetAsync(smine, 0, 24, ctx->stream);
            const unsigned long long three = 3;
            cudaMemcpyAsync(smine + 2, &three, 8, cudaMemcpyHostToDevice, ctx->stream);
            if (h->heap.tab && heads[4 * me + 2] > 0) { k_heap_export<<<small_grid(ctx, h->heap.cap_mask + 1), 256, 0, ctx->stream>>>(h->heap, smine); ctx->launches++; }
            ncclResult_t rs = N->AllGather(smine, sall, (size_t)STRBLK_WORDS, ncclUint64, comm, ctx->stream);
            if (rs == ncclSuccess) { k_heap_import<<<1, 64, 0, ctx->stream>>>(h->heap, sall, nr, me, (uint32_t*)(h->d_counters + 5)); ctx->launches++; }
            cudaStreamSynchronize(ctx->stream);          // `three` lives on this frame; the buffers go back to the allocator
            kq_dev_free(ctx, smine); kq_dev_free(ctx, sall);
            if (rs != ncclSuccess) return kq_nccl_fail(ctx, rs, "ncclAllGather");
            h->heap_entries += str_entries; h->heap_bytes += str_bytes;          // upper bounds until the next counter read
        }
        if (small) {
            uint64_t cap = 1024;
            while (cap < 4 * total) cap <<= 1;
            if (cap > h->capacity) {
                uint64_t* old = h->table; const uint64_t old_cap = h->capacity;
                h->table = nullptr;
                const int st1 = table_alloc(ctx, h, cap);              // the new table first: on failure the old one is still there
                if (st1 != KQ_OK) { h->table = old; h->capacity = old_cap; return st1; }
                kq_dev_free(ctx, old);
            } else cudaMemsetAsync(h->table, 0, (size_t)(h->capacity + 1) * stride * 8, ctx->stream);
            AggArgs A;
            memset(&A, 0, sizeof A);
            fill_common_args(h, A);
            k_merge_small<<<1, 1024, 0, ctx->stream>>>(A, all, blockw, nr, room);
            KQ_RET(launch_check(ctx, "k_merge_small"));
            h->ngroups_host = (int64_t)total; h->groups_exact = false;      // an upper bound; finalize reads the exact count
            return KQ_OK;
        }
        // some rank holds more: the general path below (every rank takes it: the decision was made on the same data)
    }
    // ---- general path: larger unions ---------------------------------------------------------------------------------------
    // 1. every rank's group count is known from the heads above; everything this path needs is allocated NOW, and the
    //    ranks agree on whether all of them succeeded before the first data collective (one word per rank, all-gathered):
    //    a rank that cannot allocate must not leave its peers waiting in ncclAllGather.
    uint64_t cnt[64];
    {
        uint64_t heads[256];
        unsigned long long* d_heads = reinterpret_cast<unsigned long long*>(h->merge_all + h->merge_words * nr);
        for (int off = 0; off < 4 * nr; off += 64) KQ_RET(kq_read_u64(ctx, d_heads + off, std::min(64, 4 * nr - off), heads + off));
        for (int i = 0; i < nr; i++) cnt[i] = heads[4 * i];
    }
    const uint64_t G = cnt[me];
    uint64_t maxn = 0;
    for (int i = 0; i < nr; i++) maxn = std::max(maxn, cnt[i]);
    const uint64_t T = maxn * (uint64_t)nr;
    const bool dense_path = T <= (1ULL << 20);
    uint64_t *mine = nullptr, *all = nullptr, *dict = nullptr, *dense = nullptr;
    unsigned long long* d_cur = nullptr;
    auto cleanup = [&](int s) { kq_dev_free(ctx, mine); kq_dev_free(ctx, all); kq_dev_free(ctx, d_cur); kq_dev_free(ctx, dict); kq_dev_free(ctx, dense); return s; };
    AggArgs D;
    memset(&D, 0, sizeof D);
    uint64_t dcap = 1024;
    while (dcap < 2 * T) dcap <<= 1;
    D.nkeys = nkeys; D.stride = (1 + nkeys + 1 + 3) / 4 * 4; D.cap_mask = dcap - 1;
    D.rec_init[1 + nkeys] = ~0ULL;
    int st = kq_dev_alloc(ctx, (size_t)maxn * stride * 8, (void**)&mine);
    if (st == KQ_OK) st = kq_dev_alloc(ctx, (size_t)T * stride * 8, (void**)&all);
    if (st == KQ_OK) st = kq_dev_alloc(ctx, 16, (void**)&d_cur);
    if (st == KQ_OK && dense_path) st = kq_dev_alloc(ctx, (size_t)(dcap + 1) * D.stride * 8 + 64, (void**)&dict);
    if (st == KQ_OK && dense_path) st = kq_dev_alloc(ctx, (size_t)T * stride * 8, (void**)&dense);
    uint64_t* newtab = nullptr;          // the merged table: room for the union of all ranks' keys
    uint64_t need = 1024;
    while (need / 2 < T) need <<= 1;
    if (st == KQ_OK && !dense_path) st = kq_dev_alloc(ctx, (size_t)(need + 1) * stride * 8, (void**)&newtab);
    if (st == KQ_OK && dense_path) {
        // dense path: the reduced values are written back into this rank's table, which must hold the union of all ranks' keys
        uint64_t want_cap = h->capacity;
        while (want_cap / 2 < T + G) want_cap <<= 1;
        if (want_cap != h->capacity) st = table_grow(ctx, h, want_cap);
    }
    {
        const unsigned long long mystatus = st == KQ_OK ? 0ULL : 1ULL;
        unsigned long long* d_st = reinterpret_cast<unsigned long long*>(h->merge_mine);
        unsigned long long* d_all = reinterpret_cast<unsigned long long*>(h->merge_all);
        cudaMemcpyAsync(d_st, &mystatus, 8, cudaMemcpyHostToDevice, ctx->stream);
        ncclResult_t rs = N->AllGather(d_st, d_all, 1, ncclUint64, comm, ctx->stream);
        uint64_t sts[64];
        int st2 = rs == ncclSuccess ? kq_read_u64(ctx, d_all, nr, sts) : kq_nccl_fail(ctx, rs, "ncclAllGather");
        bool peer_failed = false;
        for (int i = 0; st2 == KQ_OK && i < nr; i++) peer_failed |= sts[i] != 0;
        if (st != KQ_OK || st2 != KQ_OK || peer_failed) {
            kq_dev_free(ctx, newtab);
            if (st != KQ_OK) return cleanup(st);
            if (st2 != KQ_OK) return cleanup(st2);
            return cleanup(kq_fail(ctx, KQ_ERR_NCCL, "merge abandoned: another rank could not allocate its merge buffers"));
        }
    }

    // 2. all-gather the partial records themselves (padded to the largest rank; a zero header is skipped)
    ncclResult_t r;
    cudaMemsetAsync(mine, 0, (size_t)maxn * stride * 8, ctx->stream);
    cudaMemsetAsync(d_cur, 0, 16, ctx->stream);      // [0] cursor, [1] base = 0
    k_collect_records<<<small_grid(ctx, h->capacity), 256, 0, ctx->stream>>>(h->table, h->capacity, stride, nkeys, mine, d_cur, 1, d_cur + 1);
    if ((st = launch_check(ctx, "k_collect_records")) != KQ_OK) { kq_dev_free(ctx, newtab); return cleanup(st); }
    if ((r = N->AllGather(mine, all, (size_t)maxn * stride, ncclUint64, comm, ctx->stream)) != ncclSuccess) { kq_dev_free(ctx, newtab); return cleanup(kq_nccl_fail(ctx, r, "ncclAllGather")); }

    if (!dense_path) {
        // too many groups for dense arrays to pay off: every rank rebuilds its table from ALL gathered partials, its own
        // included, in rank order (a key occurs once per rank: the order of every Float64 addition is the same everywhere)
        cudaMemsetAsync(newtab, 0, (size_t)(need + 1) * stride * 8, ctx->stream);
        cudaMemsetAsync(h->d_counters, 0, 8, ctx->stream);
        kq_dev_free(ctx, h->table);
        h->table = newtab; h->capacity = need;
        AggArgs A;
        memset(&A, 0, sizeof A);
        fill_common_args(h, A);
        for (int s2 = 0; s2 < nr && st == KQ_OK; s2++) {
            if (cnt[s2] == 0) continue;
            k_merge_records<<<small_grid(ctx, cnt[s2]), 256, 0, ctx->stream>>>(A, all + (uint64_t)s2 * maxn * stride, cnt[s2]);
            st = launch_check(ctx, "k_merge_records");
        }
        if (st == KQ_OK) st = refresh_group_count(ctx, h);
        return cleanup(st);
    }

    AggArgs A;
    memset(&A, 0, sizeof A);
    fill_common_args(h, A);

    // 3. union dictionary: key -> first position in the gathered buffer (same on every rank)
    cudaMemsetAsync(dict, 0, (size_t)(dcap + 1) * D.stride * 8 + 64, ctx->stream);
    D.table = dict; D.ngroups = (unsigned long long*)(dict + (dcap + 1) * D.stride);
    k_dict_build<<<small_grid(ctx, T), 256, 0, ctx->stream>>>(D, all, T, stride);
    st = launch_check(ctx, "k_dict_build");
    // 4. dense arrays [word][position]: identities everywhere, own partials scattered in, one all-reduce per word
    k_dense_init<<<small_grid(ctx, T * stride), 256, 0, ctx->stream>>>(dense, T, stride, A);
    if (st == KQ_OK) st = launch_check(ctx, "k_dense_init");
    k_dense_scatter<<<small_grid(ctx, maxn), 256, 0, ctx->stream>>>(D, all, (uint64_t)me * maxn, maxn, stride, dense, T);
    if (st == KQ_OK) st = launch_check(ctx, "k_dense_scatter");
    int cls[MAX_REC_WORDS];
    word_classes(h, cls);
    if ((r = N->GroupStart()) != ncclSuccess) return cleanup(kq_nccl_fail(ctx, r, "ncclGroupStart"));
    for (int w = 1 + nkeys; w < stride && r == ncclSuccess; w++) {
        uint64_t* p = dense + (uint64_t)w * T;
        switch (cls[w]) {
            case WC_SUM_U64: r = N->AllReduce(p, p, T, ncclUint64, ncclSum, comm, ctx->stream); break;
            case WC_SUM_F64: r = N->AllReduce(p, p, T, ncclFloat64, ncclSum, comm, ctx->stream); break;
            case WC_MIN: r = N->AllReduce(p, p, T, ncclUint64, ncclMin, comm, ctx->stream); break;     // order-mapped values
            case WC_MAX: r = N->AllReduce(p, p, T, ncclUint64, ncclMax, comm, ctx->stream); break;
            default: break;
        }
    }
    ncclResult_t r2 = N->GroupEnd();
    if (r != ncclSuccess || r2 != ncclSuccess) return cleanup(kq_nccl_fail(ctx, r != ncclSuccess ? r : r2, "ncclAllReduce"));
    if (st != KQ_OK) return cleanup(st);
    // 5. write the reduced values into this rank's table
    k_dense_writeback<<<small_grid(ctx, T), 256, 0, ctx->stream>>>(A, D, all, T, dense);
    if ((st = launch_check(ctx, "k_dense_writeback")) != KQ_OK) return cleanup(st);
    st = refresh_group_count(ctx, h);
    return cleanup(st);
}

int kq_hashagg_repartition_alltoall(kq_ctx* ctx, kq_hashagg* h) {
    if (!ctx || !h) return KQ_ERR_ILLEGAL_ARGUMENT;
    if (!ctx->comm || ctx->nranks <= 1) return KQ_OK;
    KqNccl* N = kq_nccl(ctx);
    if (!N) return KQ_ERR_NCCL;
    cudaSetDevice(ctx->device);
    ncclComm_t comm = (ncclComm_t)ctx->comm;
    const int nr = ctx->nranks, me = ctx->rank, stride = h->stride, nkeys = (int)h->groups.size();
    if (nr > 64) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "more than 64 ranks");
    // Whatever goes wrong on this rank before the exchange (a sticky kernel error, an allocation that fails, a shape this
    // path does not support) is NOT returned right away: the peers are, or will be, inside the all-gather of step 2, so this
    // rank joins it with a status word next to its bucket sizes, and every rank leaves with an error on the same data.
    int lst = sticky_error(ctx, h);
    if (lst == KQ_OK) lst = kq_check_device_errors(ctx);
    if (lst == KQ_OK) lst = refresh_group_count(ctx, h);
    if (lst == KQ_OK && h->heap.tab && h->heap_entries > 0)
        lst = kq_fail(ctx, KQ_ERR_UNSUPPORTED, "repartitioning an aggregate with Utf8 group keys longer than 7 bytes is not implemented (merge_allreduce carries them)");
    const uint64_t G = (uint64_t)h->ngroups_host;
    const int W = nr + 1;                     // words per rank in the exchanged matrix: bucket sizes + status

    // 1. bucket the partial records by hash(key) % nranks: count, then scatter into contiguous buckets
    unsigned long long* d_meta = nullptr;     // [0..nr) counts, [nr..2nr) bases, [2nr..3nr) cursors, [3nr..3nr+nr*W) count matrix (+ status per rank)
    KQ_RET(kq_dev_alloc(ctx, (size_t)(3 * nr + nr * W) * 8, (void**)&d_meta));      // a few hundred bytes from the context's pool
    uint64_t* sendbuf = nullptr; uint64_t* recvbuf = nullptr; uint64_t* newtab = nullptr;
    auto cleanup = [&](int s) { kq_dev_free(ctx, d_meta); kq_dev_free(ctx, sendbuf); kq_dev_free(ctx, recvbuf); if (s != KQ_OK) kq_dev_free(ctx, newtab); return s; };
    cudaMemsetAsync(d_meta, 0, (size_t)(3 * nr + nr * W) * 8, ctx->stream);
    uint64_t cnt[64], base[64];
    for (int p = 0; p < 64; p++) cnt[p] = base[p] = 0;
    if (lst == KQ_OK) {
        k_count_parts<<<small_grid(ctx, h->capacity), 256, 0, ctx->stream>>>(h->table, h->capacity, stride, nkeys, nr, d_meta);
        lst = launch_check(ctx, "k_count_parts");
    }
    if (lst == KQ_OK) lst = kq_read_u64(ctx, d_meta, nr, cnt);
    if (lst == KQ_OK) {
        uint64_t acc = 0;
        for (int p = 0; p < nr; p++) { base[p] = acc; acc += cnt[p]; }
        cudaMemcpyAsync(d_meta + nr, base, (size_t)nr * 8, cudaMemcpyHostToDevice, ctx->stream);
        lst = kq_dev_alloc(ctx, (size_t)std::max<uint64_t>(G, 1) * stride * 8, (void**)&sendbuf);
    }
    if (lst == KQ_OK) {
        k_collect_records<<<small_grid(ctx, h->capacity), 256, 0, ctx->stream>>>(h->table, h->capacity, stride, nkeys, sendbuf, d_meta + 2 * nr, nr, d_meta + nr);
        lst = launch_check(ctx, "k_collect_records");
    }

    // 2. everybody learns everybody's bucket sizes — and whether everybody got this far
    unsigned long long* d_mat = d_meta + 3 * nr;
    {
        uint64_t row[65];
        for (int p = 0; p < nr; p++) row[p] = lst == KQ_OK ? cnt[p] : 0;
        row[nr] = lst == KQ_OK ? 0 : 1;
        cudaMemcpyAsync(d_mat + (size_t)me * W, row, (size_t)W * 8, cudaMemcpyHostToDevice, ctx->stream);
        cudaStreamSynchronize(ctx->stream);            // `row` and `base` live on this frame
    }
    ncclResult_t r = N->AllGather(d_mat + (size_t)me * W, d_mat, (size_t)W, ncclUint64, comm, ctx->stream);
    if (r != ncclSuccess) return cleanup(kq_nccl_fail(ctx, r, "ncclAllGather"));
    std::vector<uint64_t> mat((size_t)nr * W);
    int st = KQ_OK;
    for (int s = 0; s < nr; s++)       // kq_read_u64 moves at most 64 words at a time
        for (int off = 0; off < W; off += 64)
            if ((st = kq_read_u64(ctx, d_mat + (size_t)s * W + off, std::min(64, W - off), mat.data() + (size_t)s * W + off)) != KQ_OK) return cleanup(st);
    bool peer_failed = false;
    for (int s = 0; s < nr; s++) peer_failed |= mat[(size_t)s * W + nr] != 0;
    if (lst != KQ_OK) return cleanup(lst);
    if (peer_failed) return cleanup(kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "repartition abandoned: another rank failed before the exchange"));
    uint64_t R = 0, roff[64];
    for (int s = 0; s < nr; s++) { roff[s] = R; if (s != me) R += mat[(size_t)s * W + me]; }

    // 3. one-shot all-to-all over NVSwitch: grouped send/recv of the buckets
    if ((st = kq_dev_alloc(ctx, (size_t)std::max<uint64_t>(R, 1) * stride * 8, (void**)&recvbuf)) != KQ_OK) return cleanup(st);
    if ((r = N->GroupStart()) != ncclSuccess) return cleanup(kq_nccl_fail(ctx, r, "ncclGroupStart"));
    for (int p = 0; p < nr && r == ncclSuccess; p++) {
        if (p == me) continue;
        if (cnt[p]) r = N->Send(sendbuf + base[p] * stride, (size_t)cnt[p] * stride, ncclUint64, p, comm, ctx->stream);
        const uint64_t rc = mat[(size_t)p * W + me];
        if (r == ncclSuccess && rc) r = N->Recv(recvbuf + roff[p] * stride, (size_t)rc * stride, ncclUint64, p, comm, ctx->stream);
    }
    ncclResult_t r2 = N->GroupEnd();
    if (r != ncclSuccess || r2 != ncclSuccess) return cleanup(kq_nccl_fail(ctx, r != ncclSuccess ? r : r2, "ncclSend/ncclRecv"));

    // 4. this rank's final table: its own bucket merged with the buckets it received
    uint64_t cap = 1ULL << 16;
    while (cap / 2 < cnt[me] + R) cap <<= 1;
    if ((st = kq_dev_alloc(ctx, (size_t)(cap + 1) * stride * 8, (void**)&newtab)) != KQ_OK) return cleanup(st);
    cudaMemsetAsync(newtab, 0, (size_t)(cap + 1) * stride * 8, ctx->stream);
    cudaMemsetAsync(h->d_counters, 0, 8, ctx->stream);
    kq_dev_free(ctx, h->table);
    h->table = newtab; h->capacity = cap;
    AggArgs A;
    memset(&A, 0, sizeof A);
    fill_common_args(h, A);
    if (cnt[me]) {
        k_merge_records<<<small_grid(ctx, cnt[me]), 256, 0, ctx->stream>>>(A, sendbuf + base[me] * stride, cnt[me]);
        if ((st = launch_check(ctx, "k_merge_records")) != KQ_OK) { newtab = nullptr; return cleanup(st); }
    }
    if (R) {
        k_merge_records<<<small_grid(ctx, R), 256, 0, ctx->stream>>>(A, recvbuf, R);
        if ((st = launch_check(ctx, "k_merge_records")) != KQ_OK) { newtab = nullptr; return cleanup(st); }
    }
    st = refresh_group_count(ctx, h);
    newtab = nullptr;      // owned by h now
    return cleanup(st);
}

}  // extern "C"
