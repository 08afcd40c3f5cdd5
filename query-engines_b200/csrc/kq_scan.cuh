// kq_scan.cuh — single-pass ordered prefix across thread blocks (decoupled look-back), used by the
// FilterExec stream compaction and by the Utf8 gather (offset prefix sums).
//
// Tiles are handed out by an atomic ticket, so every tile a block waits on is owned by a block that
// is already running: the look-back spin cannot deadlock regardless of residency.
#pragma once

#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#endif

namespace kq {

constexpr unsigned long long LB_PART = 1ULL << 62;
constexpr unsigned long long LB_INCL = 2ULL << 62;
constexpr unsigned long long LB_VMASK = (1ULL << 62) - 1ULL;

// Called by all 32 lanes of ONE warp of the block that owns `tile`. Publishes the tile aggregate,
// walks back over predecessor descriptors 32 at a time and returns the exclusive prefix (same value
// in every lane). A descriptor is one 64-bit word {status:2, value:62}: written and read whole, so
// no fence is needed between status and value.
__device__ __forceinline__ unsigned long long lookback_exclusive(volatile unsigned long long* desc, long long tile,
                                                                 unsigned long long agg) {
    const int lane = threadIdx.x & 31;
    if (lane == 0) desc[tile] = (tile == 0 ? LB_INCL : LB_PART) | agg;
    unsigned long long excl = 0;
    if (tile > 0) {
        long long base = tile - 1;
        while (true) {
            long long idx = base - lane;
            unsigned long long d = LB_INCL;
            if (idx >= 0) { do { d = desc[idx]; } while ((d >> 62) == 0); }
            unsigned incl = __ballot_sync(0xffffffffu, (d >> 62) == 2);
            unsigned long long val = d & LB_VMASK;
            if (incl) { int first = __ffs(incl) - 1; if (lane > first) val = 0; }
#pragma unroll
            for (int o = 16; o; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
            excl += val;
            if (incl) break;
            base -= 32;
        }
        if (lane == 0) desc[tile] = LB_INCL | (excl + agg);
    }
    return excl;
}

// Non-blocking variant used by the filter kernel's service warp, split in two so that the L2 round
// trip of the descriptor loads overlaps other work. The tile's own aggregate has already been
// published (LB_PART) by the consumer warps. lb_load starts reading the 128 predecessor
// descriptors (4 per lane); lb_finish folds them and returns false, without waiting, if a needed
// predecessor has not published yet — the caller retries later.
__device__ __forceinline__ void lb_load(const volatile unsigned long long* desc, long long base, unsigned long long (&d)[4]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        long long idx = base - lane - 32 * j;
        d[j] = 2ULL << 62;                 // out of range: an inclusive zero
        if (idx >= 0) d[j] = desc[idx];
    }
}
__device__ __forceinline__ bool lb_finish(const volatile unsigned long long* desc, long long tile, unsigned long long (&d)[4],
                                          unsigned long long* excl_out) {
    const int lane = threadIdx.x & 31;
    unsigned long long excl = 0;
    long long base = tile - 1;
    while (true) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (__any_sync(0xffffffffu, (d[j] >> 62) == 0)) return false;
            unsigned incl = __ballot_sync(0xffffffffu, (d[j] >> 62) == 2);
            unsigned long long val = d[j] & LB_VMASK;
            if (incl) { int first = __ffs(incl) - 1; if (lane > first) val = 0; }
#pragma unroll
            for (int o = 16; o; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
            excl += val;
            if (incl) { *excl_out = excl; return true; }
        }
        base -= 128;
        lb_load(desc, base, d);
    }
}

// Blocking variant for a dedicated look-back warp. The tile's own aggregate has already been published
// (LB_PART) by the consumer warps. Reads the 128 predecessor descriptors at once (4 per lane, one L2
// round trip), each lane spinning only on descriptors that are not published yet; everything up to
// and including the nearest inclusive prefix is summed with a single butterfly; walks further back
// if no inclusive prefix was among the 128. Returns the exclusive prefix (same value in every lane).
__device__ __forceinline__ unsigned long long lb_resolve(const volatile unsigned long long* desc, long long tile) {
    const int lane = threadIdx.x & 31;
    unsigned long long excl = 0;
    long long base = tile - 1;
    while (base >= 0) {
        unsigned long long d[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const long long idx = base - lane - 32 * j;
            d[j] = LB_INCL;                    // out of range: an inclusive zero
            if (idx >= 0) d[j] = desc[idx];
        }
        // nearest inclusive prefix among the published ones (distance = 32 j + lane); lanes spin only on
        // descriptors that lie in front of it
        int stop = 4 * 32;                     // distance of the nearest inclusive descriptor found so far
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (stop < 4 * 32) break;
            const long long idx = base - lane - 32 * j;
            while ((d[j] >> 62) == 0) { __nanosleep(40); d[j] = desc[idx]; }
            const unsigned incl = __ballot_sync(0xffffffffu, (d[j] >> 62) == 2);
            if (incl) stop = 32 * j + __ffs(incl) - 1;
        }
        unsigned long long val = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) if (32 * j + lane <= stop) val += d[j] & LB_VMASK;
#pragma unroll
        for (int o = 16; o; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        excl += val;
        if (stop < 4 * 32) break;
        base -= 128;
    }
    return excl;
}

// Device-wide exclusive prefix over item lengths -> Arrow int32 offsets. `f(i)` returns the byte
// length of output item i (and may do side effects such as setting its validity bit). The item
// count is device-resident (*d_count) so no host round trip is needed after a filter.
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = 256 * SCAN_ITEMS;
template <class LenFn>
__global__ void __launch_bounds__(256) k_exclusive_offsets(LenFn f, const unsigned long long* __restrict__ d_count, int32_t* out_off,
                                                           unsigned long long* tile_desc, unsigned int* ticket,
                                                           unsigned long long* out_bytes) {
    __shared__ long long s_tile;
    __shared__ int s_w[8];
    __shared__ unsigned long long s_prefix;
    const long long m = (long long)*d_count;
    const long long ntiles = (m + SCAN_TILE - 1) / SCAN_TILE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (m == 0) { if (blockIdx.x == 0 && threadIdx.x == 0) { out_off[0] = 0; if (out_bytes) *out_bytes = 0; } return; }
    while (true) {
        if (threadIdx.x == 0) s_tile = (long long)atomicAdd(ticket, 1u);
        __syncthreads();
        const long long tile = s_tile;
        if (tile >= ntiles) break;
        long long i0 = tile * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
        int len[SCAN_ITEMS]; int tsum = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            long long i = i0 + k; int l = 0;
            if (i < m) l = f(i);
            len[k] = l; tsum += l;
        }
        int incl = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        int woff = 0, ttot = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) { int x = s_w[w]; if (w < warp) woff += x; ttot += x; }
        if (warp == 0) {
            unsigned long long excl = lookback_exclusive(tile_desc, tile, (unsigned long long)ttot);
            if (lane == 0) {
                s_prefix = excl;
                if (tile == ntiles - 1) { if (out_bytes) *out_bytes = excl + ttot; out_off[m] = (int32_t)(excl + ttot); }
            }
        }
        __syncthreads();
        int run = (int)s_prefix + woff + incl - tsum;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) { long long i = i0 + k; if (i < m) out_off[i] = run; run += len[k]; }
    }
}

}  // namespace kq
