// kq_aggtable.cuh — the global aggregation table in HBM: record layout, hashing, find-or-insert and the
// atomic accumulate step. Shared by the specialised aggregate kernel (kq_k_agg.cuh) and the table
// maintenance kernels of kq_hashagg.cu (rehash, merge, collect, finalize).
//
// Open addressing with linear probing, one AoS record per group: [0] header {state:32, key
// nullmask:32}, [1..K] key words, then per aggregate input: non-null count, sum, min, max (only the
// words the query needs), padded to a multiple of four 64-bit words so that a probe and all
// accumulator updates of a row touch one or two 32-byte sectors.
#pragma once

#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#endif
#include "kq_args.h"

namespace kq {

constexpr uint64_t HDR_EMPTY = 0, HDR_BUSY = 1, HDR_FULL = 2;

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {     // splitmix64 finaliser (same as kq_gen.h kq_mix64)
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t order_map(uint64_t bits, bool is_int) {
    if (is_int) return bits ^ 0x8000000000000000ULL;
    return bits ^ ((bits >> 63) ? ~0ULL : 0x8000000000000000ULL);
}
__host__ __device__ __forceinline__ uint64_t order_unmap(uint64_t u, bool is_int) {
    if (is_int) return u ^ 0x8000000000000000ULL;
    return u ^ ((u >> 63) ? 0x8000000000000000ULL : ~0ULL);
}
__device__ __forceinline__ uint64_t canon_nan(uint64_t bits) {
    return ((bits & 0x7fffffffffffffffULL) > 0x7ff0000000000000ULL) ? 0x7ff8000000000000ULL : bits;
}
__device__ __forceinline__ uint64_t hash_key(const uint64_t (&kw)[MAX_KEYS], uint32_t nullmask, int nkeys) {
    uint64_t h = 0x9E3779B97F4A7C15ULL + nullmask;
#pragma unroll
    for (int k = 0; k < MAX_KEYS; k++) if (k < nkeys) h = mix64(h ^ kw[k]) + 0xD1B54A32D192ED03ULL * (k + 1);
    return mix64(h);
}

// Home slot of a hash: its TOP bits. The partitioned path (kq_k_agg.cuh) partitions rows by the top bits of the same
// hash, so the groups of one partition live in one contiguous region of the table (plus what linear probing spills
// into the next): its merge works on an L2-resident slice instead of random HBM lines.
__device__ __forceinline__ uint64_t table_home(uint64_t h, uint64_t cap_mask) { return h >> __clzll((long long)cap_mask); }

// Find the record of (kw, nullmask) in the global table, inserting it if absent. Claim protocol:
// CAS header EMPTY -> BUSY|nullmask, write keys + accumulator identities, fence, publish FULL.
// The host sizes the table so that ngroups stays below capacity/2 plus margin — except when it sized it OPTIMISTICALLY
// from the planner's group-count hint (kq_hashagg.cu): then a probe sequence that has seen every slot raises
// A.overflow and the caller gets the DUMMY record behind the last slot (the table is allocated with capacity + 1
// records), so nothing spins and nothing faults; the host discards the launch, restores the table and grows it.
// `inserted` (optional): count new groups there instead of in A.ngroups — one counter for every insert of the whole
// grid serialises in the L2 (10 M inserts cost milliseconds); callers add their tally to A.ngroups in bulk.
__device__ __forceinline__ uint64_t* table_find_or_insert(const AggArgs& A, uint64_t h, const uint64_t (&kw)[MAX_KEYS], uint32_t nullmask,
                                                          uint32_t* inserted = nullptr) {
    uint64_t slot = table_home(h, A.cap_mask);
    const uint64_t full_hdr = HDR_FULL | ((uint64_t)nullmask << 32);
    uint64_t probes = 0;
    while (true) {
        if (++probes > A.cap_mask + 1 || (A.overflow && (probes & 63u) == 0 && *reinterpret_cast<volatile unsigned int*>(A.overflow))) {
            if (A.overflow) *reinterpret_cast<volatile unsigned int*>(A.overflow) = 1u;
            return A.table + (A.cap_mask + 1) * (uint64_t)A.stride;
        }
        uint64_t* rec = A.table + slot * (uint64_t)A.stride;
        uint64_t hdr = *reinterpret_cast<volatile uint64_t*>(rec);
        uint32_t state = (uint32_t)hdr;
        if (state == HDR_EMPTY) {
            unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(rec), 0ULL, HDR_BUSY | ((uint64_t)nullmask << 32));
            if (old == 0ULL) {
                for (int w = 1 + A.nkeys; w < A.stride; w++) rec[w] = A.rec_init[w];
#pragma unroll
                for (int k = 0; k < MAX_KEYS; k++) if (k < A.nkeys) rec[1 + k] = kw[k];
                asm volatile("fence.acq_rel.gpu;" ::: "memory");      // release: keys and identities before the header (a full fence.sc costs more)
                *reinterpret_cast<volatile uint64_t*>(rec) = full_hdr;
                if (inserted) ++*inserted; else atomicAdd(A.ngroups, 1ULL);
                return rec;
            }
            hdr = old; state = (uint32_t)hdr;
        }
        if (state == HDR_BUSY) { probes--; continue; }          // another thread is publishing this slot: re-read
        if (hdr == full_hdr) {
            bool eq = true;
#pragma unroll
            for (int k = 0; k < MAX_KEYS; k++) if (k < A.nkeys) eq &= (__ldcg(rec + 1 + k) == kw[k]);
            if (eq) return rec;
        }
        slot = (slot + 1) & A.cap_mask;
    }
}

__device__ __forceinline__ void global_accumulate(uint64_t* rec, const AggInput& d, uint64_t v) {
    atomicAdd(reinterpret_cast<unsigned long long*>(rec + d.rec_nn), 1ULL);
    const bool is_int = d.flags & F_INT;
    if (d.flags & F_SUM) {
        if (is_int) atomicAdd(reinterpret_cast<unsigned long long*>(rec + d.rec_sum), (unsigned long long)v);
        else atomicAdd(reinterpret_cast<double*>(rec + d.rec_sum), __longlong_as_double((long long)v));
    }
    // MIN/MAX compare like the reference's `value > this.value` (Main.kt:552): a NaN never replaces a held value. A group
    // whose non-null values were all NaN keeps the identity, which k_finalize turns into NaN.
    if ((d.flags & (F_MIN | F_MAX)) && (is_int || __longlong_as_double((long long)v) == __longlong_as_double((long long)v))) {
        uint64_t m = order_map(v, is_int);
        if ((d.flags & F_MIN) && m < __ldcg(rec + d.rec_min)) atomicMin(reinterpret_cast<unsigned long long*>(rec + d.rec_min), (unsigned long long)m);
        if ((d.flags & F_MAX) && m > __ldcg(rec + d.rec_max)) atomicMax(reinterpret_cast<unsigned long long*>(rec + d.rec_max), (unsigned long long)m);
    }
}


}  // namespace kq
