// kq_rt.cuh — device-side runtime of the specialised query kernels.
//
// An expression tree (Expression.evaluate, Main.kt:448-450; the reference materialises one Arrow
// vector per node, Main.kt:780-803) is turned by the host (kq_codegen.cu) into straight-line CUDA
// that calls the helpers below, and compiled for sm_100a together with one of the kernel skeletons
// (kq_k_ops.cuh, kq_k_agg.cuh). Types, nullability and the shared-memory staging offsets of every
// column are compile-time constants of the generated code, so a non-nullable Float64 query contains
// no validity handling at all.
//
// Values: every 64-bit value (Float64 bits, Int64, sign-extended Date32) travels as uint64_t v[R]
// plus an R-bit validity mask; Bool values are R-bit truth masks (bit r = row r), so AND / OR and the
// predicate -> selection hand-off are a few bit operations per thread.
//
// Row ownership inside a tile of WARPS*32*R rows: warp w owns rows [w*32R, (w+1)*32R); inside that,
// chunk j (of R/2) covers 64 rows and lane l owns the adjacent pair (2l, 2l+1). One pair of 8-byte
// values is one 128-bit access; one pair of validity bits comes from one 32-bit word.
#pragma once

#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#endif
#include "kq_args.h"
#include "kq_pipe.cuh"
#include "kq_parse.cuh"

namespace kq {

#ifndef KQ_R
#define KQ_R 4
#endif
constexpr int R = KQ_R;              // rows per thread (even)
constexpr int NCHUNK = R / 2;
constexpr int WARP_ROWS = 32 * R;    // rows one warp owns per tile
constexpr uint32_t RMASK = (R >= 32) ? 0xFFFFFFFFu : ((1u << R) - 1u);


// device error bits OR-ed into the ctx status word (kq_internal.h KQ_DEV_ERR_*)
constexpr uint32_t ERR_DIV0 = 1u, ERR_LONG_KEY = 2u, ERR_NUMBER_FORMAT = 4u, ERR_PARSE_RANGE = 8u, ERR_KEY_COLLISION = 16u, ERR_KEY_HEAP_FULL = 32u;

// Per-thread view of the current tile.
struct RowCtx {
    int64_t n;                       // rows in the batch
    int64_t warp_base;               // first row of this warp in the batch
    int lane;
    bool full;                       // whole tile < n: no bounds checks
    uint32_t inr;                    // R-bit mask of owned rows that are < n
    uint32_t active;                 // rows whose errors count (in range and selected)
    uint32_t* err;
    const unsigned char* stage;      // shared-memory stage holding this tile's staged buffers
    const long long* bbase;          // per column slot: data-buffer offset its staged string bytes start at (-1: not staged)
    int wrow;                        // first row of this warp inside the tile
    const KeyHeap* heap;             // long Utf8 group keys (NULL: not available to this kernel)
    __device__ __forceinline__ int64_t row0(int j) const { return warp_base + j * 64 + lane * 2; }
    __device__ __forceinline__ int trow0(int j) const { return wrow + j * 64 + lane * 2; }
};

__device__ __forceinline__ void rowctx_init(RowCtx& rc, int warp, int64_t tile, int tile_rows, int64_t n, uint32_t* err,
                                            const unsigned char* stage) {
    rc.lane = threadIdx.x & 31;
    rc.n = n;
    rc.stage = stage;
    rc.wrow = warp * WARP_ROWS;
    const int64_t tile_base = tile * tile_rows;
    rc.warp_base = tile_base + (int64_t)warp * WARP_ROWS;
    rc.full = tile_base + tile_rows <= n;
    uint32_t m = RMASK;
    if (__builtin_expect(!rc.full, 0)) {
        // closed form (the compiler predicates this block into every tile, so it is kept short): the lane owns the row pairs
        // base + 64 j, base + 64 j + 1; with t = n - 1 - base, pair j is complete for j < t / 64, pair t / 64 holds row t
        const int64_t t = n - 1 - rc.row0(0);
        m = 0;
        if (t >= 0) {
            const int64_t J = t >> 6;
            if (J >= NCHUNK) m = RMASK;
            else m = ((1u << (2 * (int)J)) - 1u) | (1u << (2 * (int)J)) | ((uint32_t)((t & 63) >= 1) << (2 * (int)J + 1));
        }
    }
    rc.inr = m;
    rc.active = m;
    rc.err = err;
    rc.bbase = nullptr;
    rc.heap = nullptr;
}

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_nc_v2(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_v4(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void stg_v2(void* p, uint2 v) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(const void* p) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(smem_u32(p)));
    return r;
}

__device__ __forceinline__ double as_f64(uint64_t x) { return __longlong_as_double((long long)x); }
__device__ __forceinline__ uint64_t as_u64(double x) { return (uint64_t)__double_as_longlong(x); }

__device__ __forceinline__ uint32_t spread16(uint32_t x) {   // bit i -> bit 2i
    x &= 0xFFFFu;
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}
// 64 row bits of a chunk from the two per-lane ballots (b0: rows 2l, b1: rows 2l+1)
__device__ __forceinline__ uint2 interleave_ballots(uint32_t b0, uint32_t b1) {
    uint2 w;
    w.x = spread16(b0) | (spread16(b1) << 1);
    w.y = spread16(b0 >> 16) | (spread16(b1 >> 16) << 1);
    return w;
}

// ---- column loads --------------------------------------------------------------------------------------
// SOFF >= 0: the buffer's tile is staged at byte offset SOFF of the shared-memory stage (filled by
// TMA bulk copies); SOFF < 0: it is read from global memory with coalesced vector loads.
template <int SOFF>
__device__ __forceinline__ uint32_t load_bits(const uint32_t* bits, const RowCtx& rc) {
    uint32_t m = 0;
    if constexpr (SOFF >= 0) {
#pragma unroll
        for (int j = 0; j < NCHUNK; j++) {
            const int t0 = rc.trow0(j);
            const uint32_t w = *reinterpret_cast<const uint32_t*>(rc.stage + SOFF + (t0 >> 5) * 4);
            m |= ((w >> (t0 & 31)) & 3u) << (2 * j);
        }
    } else {
#pragma unroll
        for (int j = 0; j < NCHUNK; j++) {
            const int64_t r0 = rc.row0(j);
            if (rc.full || r0 < rc.n) {
                const uint32_t w = __ldg(bits + (r0 >> 5));
                m |= ((w >> (r0 & 31)) & 3u) << (2 * j);
            }
        }
    }
    return m;
}
template <int SOFF>
__device__ __forceinline__ uint32_t load_valid(const uint32_t* validity, const RowCtx& rc) {
    return load_bits<SOFF>(validity, rc) & rc.inr;
}
template <int SOFF>
__device__ __forceinline__ void load64(const void* data, const RowCtx& rc, uint64_t (&v)[R]) {
#pragma unroll
    for (int j = 0; j < NCHUNK; j++) {
        uint4 q = make_uint4(0, 0, 0, 0);
        if constexpr (SOFF >= 0) q = lds_v4(rc.stage + SOFF + rc.trow0(j) * 8);
        else {
            const int64_t r0 = rc.row0(j);
            if (rc.full || r0 < rc.n) q = ldg_nc_v4(reinterpret_cast<const uint4*>(data) + (r0 >> 1));
        }
        v[2 * j] = (uint64_t)q.x | ((uint64_t)q.y << 32);
        v[2 * j + 1] = (uint64_t)q.z | ((uint64_t)q.w << 32);
    }
}
template <int SOFF>
__device__ __forceinline__ void load32(const void* data, const RowCtx& rc, uint64_t (&v)[R]) {
#pragma unroll
    for (int j = 0; j < NCHUNK; j++) {
        uint2 q = make_uint2(0, 0);
        if constexpr (SOFF >= 0) q = *reinterpret_cast<const uint2*>(rc.stage + SOFF + rc.trow0(j) * 4);
        else {
            const int64_t r0 = rc.row0(j);
            if (rc.full || r0 < rc.n) q = ldg_nc_v2(reinterpret_cast<const uint2*>(data) + (r0 >> 1));
        }
        v[2 * j] = (uint64_t)(int64_t)(int32_t)q.x;
        v[2 * j + 1] = (uint64_t)(int64_t)(int32_t)q.y;
    }
}
// byte range [a, b) of the string in owned row r
template <int SOFF>
__device__ __forceinline__ void utf8_bounds(const int32_t* offsets, const RowCtx& rc, int r, int& a, int& b) {
    if constexpr (SOFF >= 0) {
        const int32_t* o = reinterpret_cast<const int32_t*>(rc.stage + SOFF) + rc.trow0(r >> 1) + (r & 1);
        a = o[0]; b = o[1];
    } else {
        const int64_t row = rc.row0(r >> 1) + (r & 1);
        a = __ldg(offsets + row); b = __ldg(offsets + row + 1);
    }
}

// ---- Utf8 ------------------------------------------------------------------------------------------------
// three-way compare of p[0..pn) with q[0..qn): 0 = lt, 1 = eq, 2 = gt. Unsigned byte order (= code
// point order for valid UTF-8, rule R2 / oracle cmp3).
__device__ __forceinline__ int utf8_cmp3(const uint8_t* p, int pn, const uint8_t* q, int qn) {
    const int m = pn < qn ? pn : qn;
    for (int i = 0; i < m; i++) {
        const uint8_t a = p[i], b = q[i];
        if (a != b) return a < b ? 0 : 2;
    }
    return pn < qn ? 0 : (pn > qn ? 2 : 1);
}
static __device__ __noinline__ uint32_t utf8_cmp_row(const uint8_t* p, int pn, const uint8_t* q, int qn, uint32_t mask) {
    int code;
    if ((mask == CM_EQ || mask == CM_NE) && pn != qn) code = 2;
    else code = utf8_cmp3(p, pn, q, qn);
    return (mask >> code) & 1u;
}
// string column vs literal bytes -> truth mask (rows in `ok` only)
template <int SOFF_OFF>
__device__ __forceinline__ uint32_t utf8_cmp_lit(const QCol& c, const uint8_t* q, int qn, uint32_t mask, uint32_t ok, const RowCtx& rc) {
    const uint8_t* bytes = reinterpret_cast<const uint8_t*>(c.data);
    uint32_t m = 0;
#pragma unroll
    for (int r = 0; r < R; r++)
        if ((ok >> r) & 1u) { int a, b; utf8_bounds<SOFF_OFF>(c.offsets, rc, r, a, b); m |= utf8_cmp_row(bytes + a, b - a, q, qn, mask) << r; }
    return m;
}
template <int SOFF_A, int SOFF_B>
__device__ __forceinline__ uint32_t utf8_cmp_col(const QCol& c, const QCol& d, uint32_t mask, uint32_t ok, const RowCtx& rc) {
    uint32_t m = 0;
#pragma unroll
    for (int r = 0; r < R; r++)
        if ((ok >> r) & 1u) {
            int a, b, a2, b2;
            utf8_bounds<SOFF_A>(c.offsets, rc, r, a, b); utf8_bounds<SOFF_B>(d.offsets, rc, r, a2, b2);
            m |= utf8_cmp_row(reinterpret_cast<const uint8_t*>(c.data) + a, b - a, reinterpret_cast<const uint8_t*>(d.data) + a2, b2 - a2, mask) << r;
        }
    return m;
}
// 64-bit shifts with PTX semantics (an amount >= 64 yields 0; C++ leaves it undefined)
__device__ __forceinline__ uint64_t shl64c(uint64_t x, uint32_t n) { uint64_t r; asm("shl.b64 %0, %1, %2;" : "=l"(r) : "l"(x), "r"(n)); return r; }
__device__ __forceinline__ uint64_t shr64c(uint64_t x, uint32_t n) { uint64_t r; asm("shr.u64 %0, %1, %2;" : "=l"(r) : "l"(x), "r"(n)); return r; }
// the 8 bytes at (unaligned) shared-memory address `addr`, little-endian: three aligned 32-bit loads + two funnel shifts
__device__ __forceinline__ uint64_t lds_window64(uint32_t addr) {
    const uint32_t al = addr & ~3u, sh = addr << 3;          // the funnel shift takes sh mod 32 = (addr & 3) * 8
    uint32_t w0, w1, w2;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(al));
    asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(al));
    asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(w2) : "r"(al));
    return (uint64_t)__funnelshift_r(w0, w1, sh) | ((uint64_t)__funnelshift_r(w1, w2, sh) << 32);
}
// 32-bit shifts with PTX semantics (an amount >= 32 yields 0; C++ leaves it undefined, SASS would wrap)
__device__ __forceinline__ uint32_t shl32c(uint32_t x, uint32_t n) { uint32_t r; asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(n)); return r; }
__device__ __forceinline__ uint32_t shr32c(uint32_t x, uint32_t n) { uint32_t r; asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(n)); return r; }
// first `len` bytes of window w, length in the top byte (len <= 7; longer strings raise ERR_LONG_KEY, their key is unused).
// Masks instead of 64-bit shift pairs: low word keeps its low 8*len bits (all of them from len = 4), high word the low
// 8*(len-4) bits (none up to len = 4) — seven 32-bit instructions.
__device__ __forceinline__ uint64_t pack_key(uint64_t w, int len) {
    const uint32_t nb = 8u * (uint32_t)len;
    const uint32_t lo = (uint32_t)w & ~shl32c(0xFFFFFFFFu, nb);
    const uint32_t hi = ((uint32_t)(w >> 32) & shr32c(0xFFFFFFFFu, 64u - nb)) | ((uint32_t)len << 24);
    return (uint64_t)lo | ((uint64_t)hi << 32);
}
// Staged fast path of utf8_pack. FULL: the whole tile is inside the batch (no per-row range handling). ALLOK: the column
// has no nulls in range, so no key has to be blanked.
// A lane's two adjacent rows are adjacent strings: one 8-byte window at the first string's start yields both
// keys whenever the two fit in it; pairs that do not (long keys) are redone in a rarely taken second loop.
template <int SOFF_OFF, int SB, bool FULL, bool ALLOK>
__device__ __forceinline__ uint32_t utf8_pack_staged(long long base, uint32_t ok, const RowCtx& rc, uint64_t (&out)[R]) {
    const uint32_t sb32 = smem_u32(rc.stage) + (uint32_t)SB - (uint32_t)base;      // shared address of data-buffer offset 0 (mod 2^32)
    const uint32_t so32 = smem_u32(rc.stage) + (uint32_t)SOFF_OFF + (uint32_t)rc.trow0(0) * 4u;
    int mxlen = 0, mxtot = 0;
    // offsets of the lane's row pair in chunk j: o[2l], o[2l+1], o[2l+2]; rows past the end of the batch have
    // none (stale shared memory): give them an empty, in-range string
    auto bounds = [&](int j, int& a0, int& len0, int& len1) {
        const uint32_t oa = so32 + (uint32_t)j * 256u;
        int o0, o1, o2;
        asm volatile("ld.shared.v2.s32 {%0,%1}, [%2];" : "=r"(o0), "=r"(o1) : "r"(oa));
        asm volatile("ld.shared.s32 %0, [%1+8];" : "=r"(o2) : "r"(oa));
        a0 = o0; len0 = o1 - o0; len1 = o2 - o1;
        if constexpr (!FULL) {
            const bool in0 = (rc.inr >> (2 * j)) & 1u, in1 = (rc.inr >> (2 * j + 1)) & 1u;
            a0 = in0 ? a0 : (int)base; len0 = in0 ? len0 : 0; len1 = in1 ? len1 : 0;
        }
    };
#pragma unroll
    for (int j = 0; j < NCHUNK; j++) {
        int a0, len0, len1;
        bounds(j, a0, len0, len1);
        mxlen = max(mxlen, max(len0, len1));
        mxtot = max(mxtot, len0 + len1);
        const uint64_t w0 = lds_window64(sb32 + (uint32_t)a0);
        const uint64_t k0 = pack_key(w0, len0), k1 = pack_key(shr64c(w0, 8u * (uint32_t)len0), len1);
        if constexpr (ALLOK) { out[2 * j] = k0; out[2 * j + 1] = k1; }
        else {
            out[2 * j] = ((ok >> (2 * j)) & 1u) ? k0 : 0ULL;
            out[2 * j + 1] = ((ok >> (2 * j + 1)) & 1u) ? k1 : 0ULL;
        }
    }
    uint32_t too_long = 0;
    if (__builtin_expect(mxtot > 8 || mxlen > 7, 0)) {
#pragma unroll 1
        for (int j = 0; j < NCHUNK; j++) {
            int a0, len0, len1;
            bounds(j, a0, len0, len1);
            too_long |= (uint32_t)(len0 > 7) << (2 * j) | (uint32_t)(len1 > 7) << (2 * j + 1);
            if (len0 + len1 > 8) {
                const uint64_t k1 = pack_key(lds_window64(sb32 + (uint32_t)(a0 + len0)), len1 > 7 ? 7 : len1);
#pragma unroll
                for (int jj = 0; jj < NCHUNK; jj++) if (jj == j) out[2 * jj + 1] = ((ok >> (2 * jj + 1)) & 1u) ? k1 : 0ULL;
            }
        }
    }
    return too_long;
}
// short string (<= 7 bytes) -> packed u64 group key: bytes little-endian, length in the top byte.
// SB >= 0: the tile's string bytes are staged at byte offset SB of the stage (when they fit: rc.bbase[SLOT] >= 0);
// that path is branch-free per row (uniform branches per tile only).
// One string -> its key word, the slow way (out of line, scalar arguments only so that nothing of the caller's tile is
// forced into local memory): bytes read from global memory; up to 7 bytes are packed, longer strings are interned in the
// aggregate's key heap (or raise ERR_LONG_KEY in a kernel that has none).
static __device__ __noinline__ uint64_t utf8_key_slow(const uint8_t* bytes, int a, int len, const KeyHeap* heap, uint32_t* err, bool active) {
    if (len <= 7) {
        uint64_t key = 0;
        for (int i = 0; i < len; i++) key |= (uint64_t)__ldg(bytes + a + i) << (8 * i);
        return key | ((uint64_t)len << 56);
    }
    if (!active) return 0;
    if (heap == nullptr) { atomicOr(err, ERR_LONG_KEY); return 0; }
    return utf8_intern(heap, bytes + a, len, err);
}
template <int SOFF_OFF, int SB, int SLOT>
__device__ __forceinline__ void utf8_pack(const QCol& c, uint32_t ok, const RowCtx& rc, uint64_t (&out)[R]) {
    long long base = -1;
    if constexpr (SB >= 0 && SOFF_OFF >= 0) base = rc.bbase[SLOT];
    uint32_t redo = 0;             // rows whose key the fast path could not produce
    if (base >= 0) {
        if (rc.full) { if (ok == RMASK) redo = utf8_pack_staged<SOFF_OFF, SB, true, true>(base, ok, rc, out); else redo = utf8_pack_staged<SOFF_OFF, SB, true, false>(base, ok, rc, out); }
        else redo = utf8_pack_staged<SOFF_OFF, SB, false, false>(base, ok, rc, out);
        redo &= ok;
    } else {
        // the tile's string bytes did not fit the stage (long strings): every row the slow way
        redo = ok;
#pragma unroll
        for (int r = 0; r < R; r++) out[r] = 0;
    }
    if (redo) {
#pragma unroll 1
        for (uint32_t m = redo; m; m &= m - 1) {
            const int r = __ffs(m) - 1;
            const int32_t* o;
            if constexpr (SOFF_OFF >= 0) o = reinterpret_cast<const int32_t*>(rc.stage + SOFF_OFF) + rc.trow0(r >> 1) + (r & 1);
            else o = c.offsets + rc.row0(r >> 1) + (r & 1);
            const int a = o[0], b = o[1];
            const uint64_t key = utf8_key_slow(reinterpret_cast<const uint8_t*>(c.data), a, b - a, rc.heap, rc.err, (rc.active >> r) & 1u);
#pragma unroll
            for (int rr = 0; rr < R; rr++) if (rr == r) out[rr] = key;
        }
    }
}

// CastExpression Utf8 -> Float64: Java Double.parseDouble grammar (rule R5). Decimal inputs with at
// most 19 significant digits and a decimal exponent in [-22, 22] are converted exactly with one
// IEEE multiply/divide (Clinger's fast path), which is correctly rounded. Everything else the grammar
// accepts (more digits, large exponents, hex floats) takes the exact big-integer tier of kq_parse.cuh.
__device__ __forceinline__ double kq_p10(int e) {
    double p = 1.0;      // exact: every 10^k, k <= 22, is representable
    for (int i = 0; i < e; i++) p = __dmul_rn(p, 10.0);
    return p;
}
// returns 0 = ok (fast path), else: not decided here (malformed, or valid but outside the fast path)
__device__ __forceinline__ int parse_f64_fast(const uint8_t* p, int n, double& out) {
    int b = 0, e = n;
    while (b < e && p[b] <= ' ') b++;
    while (e > b && p[e - 1] <= ' ') e--;
    if (b >= e) return 1;
    bool neg = false;
    if (p[b] == '+' || p[b] == '-') { neg = p[b] == '-'; b++; }
    int len = e - b;
    if (len == 3 && p[b] == 'N' && p[b + 1] == 'a' && p[b + 2] == 'N') { out = __longlong_as_double(0x7ff8000000000000LL); return 0; }
    if (len == 8 && p[b] == 'I' && p[b + 1] == 'n' && p[b + 2] == 'f' && p[b + 3] == 'i' && p[b + 4] == 'n' &&
        p[b + 5] == 'i' && p[b + 6] == 't' && p[b + 7] == 'y') {
        out = neg ? __longlong_as_double(0xfff0000000000000LL) : __longlong_as_double(0x7ff0000000000000LL);
        return 0;
    }
    if (len <= 0) return 1;
    uint8_t last = p[e - 1];
    if (last == 'd' || last == 'D' || last == 'f' || last == 'F') { e--; if (e <= b) return 1; }
    uint64_t mant = 0; int nd = 0, sig = 0, dec_exp = 0; bool seen_nonzero = false;
    int i = b;
    for (; i < e && p[i] >= '0' && p[i] <= '9'; i++) {
        nd++;
        if (p[i] != '0') seen_nonzero = true;
        if (seen_nonzero) { if (sig < 19) { mant = mant * 10 + (p[i] - '0'); sig++; } else return 2; }
    }
    if (i < e && p[i] == '.') {
        i++;
        for (; i < e && p[i] >= '0' && p[i] <= '9'; i++) {
            nd++;
            if (p[i] != '0') seen_nonzero = true;
            if (seen_nonzero) { if (sig < 19) { mant = mant * 10 + (p[i] - '0'); sig++; } else return 2; }
            dec_exp--;
        }
    }
    if (nd == 0) return 1;
    if (i < e && (p[i] == 'e' || p[i] == 'E')) {
        i++;
        bool eneg = false;
        if (i < e && (p[i] == '+' || p[i] == '-')) { eneg = p[i] == '-'; i++; }
        int ed = 0, ev = 0;
        for (; i < e && p[i] >= '0' && p[i] <= '9'; i++) { ed++; if (ev < 100000) ev = ev * 10 + (p[i] - '0'); }
        if (ed == 0) return 1;
        dec_exp += eneg ? -ev : ev;
    }
    if (i != e) return 1;
    double v;
    if (mant == 0) v = 0.0;
    else {
        if (mant > (1ULL << 53) || dec_exp < -22 || dec_exp > 22) return 2;
        v = (double)(int64_t)mant;
        v = dec_exp >= 0 ? __dmul_rn(v, kq_p10(dec_exp)) : __ddiv_rn(v, kq_p10(-dec_exp));
    }
    out = neg ? -v : v;
    return 0;
}
// the exact tier, out of line: ~500 bytes of local memory for the big integer, only on the rows that need it
static __device__ __noinline__ int parse_f64_exact(const uint8_t* p, int n, double& out) {
    uint64_t bits = 0;
    const int st = parse_java_double(p, n, &bits);
    out = __longlong_as_double((long long)bits);
    return st;
}
static __device__ __noinline__ double parse_f64_row(const uint8_t* p, int n, uint32_t* err, bool active) {
    double v = 0.0;
    if (parse_f64_fast(p, n, v) != 0) {
        if (parse_f64_exact(p, n, v) != 0) { v = 0.0; if (active) atomicOr(err, ERR_NUMBER_FORMAT); }      // NumberFormatException (Main.kt:791)
    }
    return v;
}
template <int SOFF_OFF>
__device__ __forceinline__ void utf8_to_f64(const QCol& c, uint32_t ok, const RowCtx& rc, uint64_t (&out)[R]) {
    const uint8_t* bytes = reinterpret_cast<const uint8_t*>(c.data);
#pragma unroll
    for (int r = 0; r < R; r++) {
        double v = 0.0;
        if ((ok >> r) & 1u) { int a, b; utf8_bounds<SOFF_OFF>(c.offsets, rc, r, a, b); v = parse_f64_row(bytes + a, b - a, rc.err, (rc.active >> r) & 1u); }
        out[r] = as_u64(v);
    }
}

// Int64 division with JVM semantics: x / 0 raises (rule E4) on active valid rows, MIN / -1 wraps.
__device__ __forceinline__ void div_i64(const uint64_t (&x)[R], const uint64_t (&y)[R], uint32_t both, const RowCtx& rc, uint64_t (&o)[R]) {
#pragma unroll
    for (int r = 0; r < R; r++) {
        const long long a = (long long)x[r], b = (long long)y[r];
        long long q = 0;
        if (b == 0) {
            if (((both & rc.active) >> r) & 1u) atomicOr(rc.err, ERR_DIV0);
        } else if (b == -1) q = (long long)(0ULL - (unsigned long long)a);
        else q = a / b;
        o[r] = (uint64_t)q;
    }
}

// write the 64 row bits of chunk j (rows warp_base + 64j ...) of a bit-packed buffer
__device__ __forceinline__ void store_chunk_bits(uint32_t* bits, const RowCtx& rc, int j, uint32_t m) {
    const uint32_t b0 = __ballot_sync(0xffffffffu, (m >> (2 * j)) & 1u);
    const uint32_t b1 = __ballot_sync(0xffffffffu, (m >> (2 * j + 1)) & 1u);
    const int64_t chunk_base = rc.warp_base + j * 64;
    if (rc.lane == 0 && chunk_base < rc.n) {
        const uint2 w = interleave_ballots(b0, b1);
        *reinterpret_cast<uint2*>(bits + (chunk_base >> 5)) = w;
    }
}

}  // namespace kq
