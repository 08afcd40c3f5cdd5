// kq_pipe.cuh — the tile pipeline every streaming kernel shares.
//
// A producer warp (one elected lane) moves the next tiles of every referenced column buffer from
// HBM into a ring of shared-memory stages with 1-D TMA bulk copies (cp.async.bulk ... mbarrier::
// complete_tx, SASS UBLKCP); consumer warps wait on the stage's "full" mbarrier, run the expression
// VM out of shared memory and hand the stage back through its "empty" mbarrier. Bytes in flight per
// SM = (stages - 1) x stage bytes, independent of occupancy and of how the VM serialises its loads —
// this is what lets an HBM-bound integer/byte path approach the copy roofline on B200.
#pragma once

#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#endif
#include "kq_args.h"

namespace kq {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Non-blocking probe (try_wait may suspend the thread for a hardware-defined time; test_wait never does).
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}

// Hand a stage back to the producer: called by ALL lanes of a consumer warp after their last read of the stage.
// The shared-memory loads of a warp can still be in flight when its next instruction issues, and nothing holds an
// mbarrier arrive back behind them (the SASS is LDS ...; WARPSYNC; SYNCS.ARRIVE with no scoreboard wait in between).
// Measured on B200 with heavily skewed consumer warps (tests/test_gpu_parity.py, skewed partition keys): without the
// drain below a few warp-tiles in ten thousand were read from a stage that had already been refilled. The drain reads
// one shared word whose value is known (`word` must hold `expect`) and branches on it: loads of a warp return in order,
// so once this one is back every earlier load of the stage is too. A mismatch is a protocol violation and is reported.
constexpr uint32_t ERR_STAGE_PROTOCOL = 0x2000u;
__device__ __forceinline__ void stage_release(uint64_t* empty_bar, const void* word, uint32_t expect, uint32_t* err, int lane) {
    __syncwarp();
    uint32_t d;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(d) : "r"(smem_u32(word)) : "memory");
    if (d != expect) atomicOr(err, ERR_STAGE_PROTOCOL);
    __syncwarp();                 // every lane has its word back (lanes do not run in lockstep: lane 0 must not arrive — and let `word` change — before the others looked)
    if (lane == 0) mbar_arrive(empty_bar);
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on `bar`.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// the same with an L2 eviction policy (createpolicy): streaming inputs that must not displace lines another part of the
// kernel wants to keep in the L2
__device__ __forceinline__ void bulk_g2s_policy(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t l2_policy_evict_last() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Issue the bulk copies of tile `tile` (rows [tile*tile_rows, ...)) into one stage. Called by one thread.
__device__ __forceinline__ void stage_issue(const StagePlan& sp, unsigned char* stage, uint64_t* full, int64_t tile,
                                            int tile_rows, int64_t n) {
    int64_t row0 = tile * tile_rows;
    int rows = (int)((n - row0) < tile_rows ? (n - row0) : tile_rows);
    uint32_t total = 0;
#pragma unroll 1
    for (int b = 0; b < sp.nbuf; b++) {
        const int kind = sp.buf[b].kind;
        if (kind == SK_BYTES) continue;     // second phase (stage_issue_bytes): needs the staged offsets
        uint32_t bytes = kind == SK_W8 ? rows * 8 : (kind == SK_W4 ? rows * 4 : (kind == SK_W4_PLUS1 ? (rows + 1) * 4 : (rows + 7) / 8));
        total += (bytes + 15u) & ~15u;
    }
    mbar_arrive_expect_tx(full, total);
#pragma unroll 1
    for (int b = 0; b < sp.nbuf; b++) {
        const StageBuf sb = sp.buf[b];
        if (sb.kind == SK_BYTES) continue;
        uint32_t bytes; int64_t goff;
        if (sb.kind == SK_W8) { bytes = rows * 8; goff = row0 * 8; }
        else if (sb.kind == SK_W4) { bytes = rows * 4; goff = row0 * 4; }
        else if (sb.kind == SK_W4_PLUS1) { bytes = (rows + 1) * 4; goff = row0 * 4; }
        else { bytes = (rows + 7) / 8; goff = row0 / 8; }
        bytes = (bytes + 15u) & ~15u;       // device buffers are padded (KQ_PAD), whole vectors past the end are readable
        bulk_g2s(stage + sb.soff, sb.g + goff, bytes, full);
    }
}

// Ask the L2 to fetch the fixed-size buffers of a tile that will be staged a few iterations from now: the
// later bulk copy then sees L2 latency instead of HBM latency, so the shared-memory ring has to cover far
// fewer bytes in flight. Called by one thread.
__device__ __forceinline__ void stage_prefetch_l2(const StagePlan& sp, int64_t tile, int tile_rows, int64_t n) {
    const int64_t row0 = tile * tile_rows;
    if (row0 >= n) return;
    const int rows = (int)((n - row0) < tile_rows ? (n - row0) : tile_rows);
#pragma unroll 1
    for (int b = 0; b < sp.nbuf; b++) {
        const StageBuf sb = sp.buf[b];
        if (sb.kind == SK_BYTES) continue;
        uint32_t bytes; int64_t goff;
        if (sb.kind == SK_W8) { bytes = rows * 8; goff = row0 * 8; }
        else if (sb.kind == SK_W4) { bytes = rows * 4; goff = row0 * 4; }
        else if (sb.kind == SK_W4_PLUS1) { bytes = (rows + 1) * 4; goff = row0 * 4; }
        else { bytes = (rows + 7) / 8; goff = row0 / 8; }
        bytes = (bytes + 15u) & ~15u;
        bulk_prefetch_l2(sb.g + goff, bytes);
    }
}

// Boundary offsets of a tile's staged Utf8 columns, read straight from HBM by the producer lane as soon as it
// holds the tile's ticket (one step before the tile gets a stage): with them in registers, the string bytes are
// issued together with the fixed-size buffers (stage_issue_all) instead of a dependent second round trip.
struct TileBounds { int32_t lo[MAX_BYTES_BUFS], hi[MAX_BYTES_BUFS]; };
__device__ __forceinline__ void stage_bounds_fetch(const StagePlan& sp, int64_t tile, int tile_rows, int64_t n, TileBounds& tb) {
    const int64_t row0 = tile * tile_rows;
    const int rows = (int)((n - row0) < tile_rows ? (n - row0) : tile_rows);
#pragma unroll
    for (int i = 0; i < MAX_BYTES_BUFS; i++) {
        tb.lo[i] = tb.hi[i] = 0;
        if (i < sp.nbytes) {
            const int32_t* off = reinterpret_cast<const int32_t*>(sp.buf[sp.buf[sp.bytes_buf[i]].aux & 0xffff].g) + row0;
            asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(tb.lo[i]) : "l"(off));
            asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(tb.hi[i]) : "l"(off + rows));
        }
    }
}

// Issue every bulk copy of a tile — fixed-size buffers and the Utf8 byte ranges given by `tb` — on ONE barrier.
// bbase[column slot] receives the data-buffer offset the staged bytes start at (16-byte aligned), or -1 when the
// range does not fit the stage (consumers then read those bytes from global memory). Called by one thread.
__device__ __forceinline__ void stage_issue_all(const StagePlan& sp, unsigned char* stage, uint64_t* full, int64_t tile,
                                                int tile_rows, int64_t n, const TileBounds& tb, long long* bbase, uint64_t policy = 0) {
    const int64_t row0 = tile * tile_rows;
    const int rows = (int)((n - row0) < tile_rows ? (n - row0) : tile_rows);
    uint32_t total = 0;
#pragma unroll 1
    for (int b = 0; b < sp.nbuf; b++) {
        const int kind = sp.buf[b].kind;
        if (kind == SK_BYTES) continue;
        uint32_t bytes = kind == SK_W8 ? rows * 8 : (kind == SK_W4 ? rows * 4 : (kind == SK_W4_PLUS1 ? (rows + 1) * 4 : (rows + 7) / 8));
        total += (bytes + 15u) & ~15u;
    }
    long long start[MAX_BYTES_BUFS]; uint32_t len[MAX_BYTES_BUFS];
#pragma unroll
    for (int i = 0; i < MAX_BYTES_BUFS; i++) {
        start[i] = -1; len[i] = 0;
        if (i < sp.nbytes) {
            const StageBuf& sb = sp.buf[sp.bytes_buf[i]];
            const long long lo = (long long)tb.lo[i] & ~15LL, hi = ((long long)tb.hi[i] + 15LL) & ~15LL;
            const bool fits = hi - lo + 16 <= (long long)sb.cap;
            if (fits) { start[i] = lo; len[i] = (uint32_t)(hi - lo + 16); }     // +16: consumers read whole 8-byte words past the end (buffers are padded)
            bbase[sb.aux >> 16] = start[i];
            total += len[i];
        }
    }
    mbar_arrive_expect_tx(full, total);
#pragma unroll 1
    for (int b = 0; b < sp.nbuf; b++) {
        const StageBuf sb = sp.buf[b];
        if (sb.kind == SK_BYTES) continue;
        uint32_t bytes; int64_t goff;
        if (sb.kind == SK_W8) { bytes = rows * 8; goff = row0 * 8; }
        else if (sb.kind == SK_W4) { bytes = rows * 4; goff = row0 * 4; }
        else if (sb.kind == SK_W4_PLUS1) { bytes = (rows + 1) * 4; goff = row0 * 4; }
        else { bytes = (rows + 7) / 8; goff = row0 / 8; }
        bytes = (bytes + 15u) & ~15u;
        if (policy) bulk_g2s_policy(stage + sb.soff, sb.g + goff, bytes, full, policy); else bulk_g2s(stage + sb.soff, sb.g + goff, bytes, full);
    }
#pragma unroll
    for (int i = 0; i < MAX_BYTES_BUFS; i++)
        if (i < sp.nbytes && len[i]) {
            const StageBuf& sb = sp.buf[sp.bytes_buf[i]];
            if (policy) bulk_g2s_policy(stage + sb.soff, sb.g + start[i], len[i], full, policy); else bulk_g2s(stage + sb.soff, sb.g + start[i], len[i], full);
        }
}

// Second phase for Utf8 columns: once the tile's offsets have landed in the stage, copy the byte range
// they span. bbase[column slot] receives the data-buffer offset the staged bytes start at (16-byte
// aligned), or -1 when the range does not fit the stage (consumers then read those bytes from global
// memory). Called by one thread after waiting for the first-phase barrier.
__device__ __forceinline__ void stage_issue_bytes(const StagePlan& sp, unsigned char* stage, uint64_t* full2, int64_t tile,
                                                  int tile_rows, int64_t n, long long* bbase) {
    const int64_t row0 = tile * tile_rows;
    const int rows = (int)((n - row0) < tile_rows ? (n - row0) : tile_rows);
    uint32_t total = 0;
    long long start[4]; uint32_t len[4]; int nb = 0;
#pragma unroll 1
    for (int b = 0; b < sp.nbuf && nb < 4; b++) {
        const StageBuf sb = sp.buf[b];
        if (sb.kind != SK_BYTES) continue;
        const int32_t* off = reinterpret_cast<const int32_t*>(stage + sp.buf[sb.aux & 0xffff].soff);
        const long long lo = (long long)off[0] & ~15LL, hi = ((long long)off[rows] + 15LL) & ~15LL;
        const bool fits = hi - lo + 16 <= (long long)sb.cap;
        start[nb] = fits ? lo : -1; len[nb] = fits ? (uint32_t)(hi - lo + 16) : 0u;     // +16: consumers read whole 8-byte words past the end (buffers are padded)
        bbase[sb.aux >> 16] = start[nb];
        total += len[nb];
        nb++;
    }
    mbar_arrive_expect_tx(full2, total);
    nb = 0;
#pragma unroll 1
    for (int b = 0; b < sp.nbuf && nb < 4; b++) {
        const StageBuf sb = sp.buf[b];
        if (sb.kind != SK_BYTES) continue;
        if (len[nb]) bulk_g2s(stage + sb.soff, sb.g + start[nb], len[nb], full2);
        nb++;
    }
}

}  // namespace kq
