// kq_vm.cuh — the fused expression evaluator shared by every operator kernel.
//
// An expression tree (Expression.evaluate, Main.kt:448-450; the reference materialises one Arrow
// vector per node, Main.kt:780-803) is flattened on the host into a postfix program and executed
// by ONE kernel. The program is uniform across the grid, so instruction dispatch is a
// divergence-free branch whose cost is amortised over the KQ_R rows each thread carries. The
// evaluation stack lives in registers: every stack access is indexed by the compile-time stack
// pointer (template parameter SP — the host compiler stamps the stack pointer into each
// instruction), so nothing spills to local memory.
//
// Row ownership inside a tile of TILE = BLOCK*R rows: warp w owns rows [w*32R, (w+1)*32R); inside
// that, chunk j (of R/2) covers 64 rows and lane l owns the adjacent pair (2l, 2l+1). One pair of
// 8-byte values is one 128-bit coalesced load; one pair of validity bits comes from one 32-bit word.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace kq {

constexpr int R = 4;                 // rows per thread (even)
constexpr int NCHUNK = R / 2;
constexpr int BLOCK = 256;
constexpr int WARPS = BLOCK / 32;
constexpr int WARP_ROWS = 32 * R;
constexpr int TILE = BLOCK * R;
constexpr int D = 6;                 // evaluation stack depth
constexpr int MAX_COLS = 16;
constexpr int MAX_INSN = 96;
constexpr int MAX_LIT = 32;
constexpr int LITPOOL = 256;
constexpr int MAX_OUT = 8;
constexpr int MAX_KEYS = 4;
constexpr int MAX_INPUTS = 6;
constexpr uint32_t RMASK = (1u << R) - 1u;

enum Op : uint8_t {
    OP_END = 0,
    OP_PUSH_COL64, OP_PUSH_COL32, OP_PUSH_COLBIT, OP_PUSH_LIT, OP_PUSH_NULL, OP_PUSH_VALID,
    OP_ADD_I64, OP_SUB_I64, OP_MUL_I64, OP_DIV_I64,
    OP_ADD_F64, OP_SUB_F64, OP_MUL_F64, OP_DIV_F64,
    OP_CMP_I64, OP_CMP_F64,          // arg = 4-bit truth mask over {lt, eq, gt, unordered}
    OP_AND, OP_OR,
    OP_UTF8_CMP_LIT,                 // arg = col | lit << 5 | mask << 10
    OP_UTF8_CMP_COL,                 // arg = colA | colB << 5 | mask << 10
    OP_UTF8_PACK,                    // arg = col: short string (<= 7 bytes) -> packed u64 key
    OP_I64_TO_F64,
    OP_UTF8_TO_F64,                  // arg = col: CastExpression Utf8 -> Float64 (Main.kt:772-805)
    OP_PICK,                         // arg = depth below top to copy
    OP_SET_SEL, OP_EMIT, OP_SET_KEY, OP_SET_IN
};

// comparison truth masks: bit0 = lt, bit1 = eq, bit2 = gt, bit3 = unordered (NaN)
constexpr uint32_t CM_EQ = 0x2, CM_NE = 0xD, CM_LT = 0x1, CM_LE = 0x3, CM_GT = 0x4, CM_GE = 0x6;

struct Insn { uint8_t op, sp; uint16_t arg; };

struct DCol {
    const void* data;
    const uint32_t* validity;
    const int32_t* offsets;
};

struct Program {
    int32_t ninsn;
    int32_t ncols;
    Insn insn[MAX_INSN];
    uint64_t lit[MAX_LIT];          // utf8 literal: (pool offset << 32) | length
    uint8_t pool[LITPOOL];
    DCol cols[MAX_COLS];
};

struct Stack {
    uint64_t v[D][R];
    uint32_t ok[D];                  // R-bit validity mask per slot
};

// Per-thread view of the current tile.
struct RowCtx {
    int64_t n;                       // rows in the batch
    int64_t warp_base;               // first row of this warp in the batch
    int lane;
    bool full;                       // whole tile < n: no bounds checks
    uint32_t inr;                    // R-bit mask of owned rows that are < n
    uint32_t active;                 // rows whose errors count (in range and selected)
    uint32_t* err;
    __device__ __forceinline__ int64_t row0(int j) const { return warp_base + j * 64 + lane * 2; }
};

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
__device__ __forceinline__ uint32_t load_bits(const uint32_t* bits, const RowCtx& rc) {
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < NCHUNK; j++) {
        int64_t r0 = rc.row0(j);
        if (rc.full || r0 < rc.n) {
            uint32_t w = __ldg(bits + (r0 >> 5));
            m |= ((w >> (r0 & 31)) & 3u) << (2 * j);
        }
    }
    return m;
}
__device__ __forceinline__ uint32_t load_valid(const DCol& c, const RowCtx& rc) {
    return c.validity ? (load_bits(c.validity, rc) & rc.inr) : rc.inr;
}
__device__ __forceinline__ void load64(const DCol& c, const RowCtx& rc, uint64_t (&v)[R], uint32_t& ok) {
    const uint4* p = reinterpret_cast<const uint4*>(c.data);
#pragma unroll
    for (int j = 0; j < NCHUNK; j++) {
        int64_t r0 = rc.row0(j);
        uint4 q = make_uint4(0, 0, 0, 0);
        if (rc.full || r0 < rc.n) q = ldg_nc_v4(p + (r0 >> 1));
        v[2 * j] = (uint64_t)q.x | ((uint64_t)q.y << 32);
        v[2 * j + 1] = (uint64_t)q.z | ((uint64_t)q.w << 32);
    }
    ok = load_valid(c, rc);
}
__device__ __forceinline__ void load32(const DCol& c, const RowCtx& rc, uint64_t (&v)[R], uint32_t& ok) {
    const uint2* p = reinterpret_cast<const uint2*>(c.data);
#pragma unroll
    for (int j = 0; j < NCHUNK; j++) {
        int64_t r0 = rc.row0(j);
        uint2 q = make_uint2(0, 0);
        if (rc.full || r0 < rc.n) q = ldg_nc_v2(p + (r0 >> 1));
        v[2 * j] = (uint64_t)(int64_t)(int32_t)q.x;
        v[2 * j + 1] = (uint64_t)(int64_t)(int32_t)q.y;
    }
    ok = load_valid(c, rc);
}
__device__ __forceinline__ void loadbit(const DCol& c, const RowCtx& rc, uint64_t (&v)[R], uint32_t& ok) {
    uint32_t m = load_bits(reinterpret_cast<const uint32_t*>(c.data), rc);
#pragma unroll
    for (int r = 0; r < R; r++) v[r] = (m >> r) & 1u;
    ok = load_valid(c, rc);
}

// ---- Utf8 ------------------------------------------------------------------------------------------------
// three-way compare of string row `row` of column c with bytes q[0..qn): 0 = lt, 1 = eq, 2 = gt.
// Unsigned byte order (= code point order for valid UTF-8, rule R2 / oracle cmp3).
__device__ __forceinline__ int utf8_cmp3(const uint8_t* p, int pn, const uint8_t* q, int qn) {
    int m = pn < qn ? pn : qn;
    for (int i = 0; i < m; i++) {
        uint8_t a = p[i], b = q[i];
        if (a != b) return a < b ? 0 : 2;
    }
    return pn < qn ? 0 : (pn > qn ? 2 : 1);
}

// CastExpression Utf8 -> Float64: Java Double.parseDouble grammar (rule R5). Decimal inputs with at
// most 19 significant digits and a decimal exponent in [-22, 22] are converted exactly with one
// IEEE multiply/divide (Clinger's fast path), which is correctly rounded. Anything else (hex
// floats, more digits, large exponents) raises KQ_DEV_ERR_NUMBER_FORMAT — see DESIGN.md.
static __constant__ double KQ_P10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14,
                                          1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
// returns 0 = ok, 1 = malformed (NumberFormatException), 2 = valid but outside the exact fast path
__device__ __forceinline__ int parse_f64(const uint8_t* p, int n, double& out) {
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
        v = dec_exp >= 0 ? __dmul_rn(v, KQ_P10[dec_exp]) : __ddiv_rn(v, KQ_P10[-dec_exp]);
    }
    out = neg ? -v : v;
    return 0;
}

// ---- one instruction at a compile-time stack pointer ---------------------------------------------------------
template <int SP, class Sink>
__device__ __forceinline__ void step(const Program& P, const Insn in, Stack& st, RowCtx& rc, Sink& sink) {
    constexpr int T = SP - 1;      // top
    constexpr int U = SP - 2;      // under top
    switch (in.op) {
        case OP_PUSH_COL64: if constexpr (SP < D) load64(P.cols[in.arg], rc, st.v[SP], st.ok[SP]); break;
        case OP_PUSH_COL32: if constexpr (SP < D) load32(P.cols[in.arg], rc, st.v[SP], st.ok[SP]); break;
        case OP_PUSH_COLBIT: if constexpr (SP < D) loadbit(P.cols[in.arg], rc, st.v[SP], st.ok[SP]); break;
        case OP_PUSH_LIT:
            if constexpr (SP < D) {
                uint64_t x = P.lit[in.arg];
#pragma unroll
                for (int r = 0; r < R; r++) st.v[SP][r] = x;
                st.ok[SP] = rc.inr;
            }
            break;
        case OP_PUSH_NULL:
            if constexpr (SP < D) {
#pragma unroll
                for (int r = 0; r < R; r++) st.v[SP][r] = 0;
                st.ok[SP] = 0;
            }
            break;
        case OP_PUSH_VALID:        // COUNT(col) of any type only needs the validity bit
            if constexpr (SP < D) {
#pragma unroll
                for (int r = 0; r < R; r++) st.v[SP][r] = 1;
                st.ok[SP] = load_valid(P.cols[in.arg], rc);
            }
            break;
        case OP_PICK:
            if constexpr (SP < D && SP >= 1) {
                // copy slot (T - arg) to the top; arg is small, resolved by a uniform switch
                switch (in.arg) {
#define KQ_PICK(k) case k: if constexpr (T - k >= 0) { _Pragma("unroll") for (int r = 0; r < R; r++) st.v[SP][r] = st.v[T - k][r]; st.ok[SP] = st.ok[T - k]; } break;
                    KQ_PICK(0) KQ_PICK(1) KQ_PICK(2) KQ_PICK(3) KQ_PICK(4)
#undef KQ_PICK
                }
            }
            break;
#define KQ_BIN(opname, expr)                                              \
        case opname:                                                      \
            if constexpr (SP >= 2) {                                      \
                _Pragma("unroll") for (int r = 0; r < R; r++) {           \
                    uint64_t a = st.v[U][r], b = st.v[T][r]; (void)a; (void)b; \
                    st.v[U][r] = (expr);                                  \
                }                                                         \
                st.ok[U] &= st.ok[T];                                     \
            }                                                             \
            break;
        KQ_BIN(OP_ADD_I64, a + b)
        KQ_BIN(OP_SUB_I64, a - b)
        KQ_BIN(OP_MUL_I64, a * b)
        // separately rounded IEEE operations: never contracted into FMA (rule E4)
        KQ_BIN(OP_ADD_F64, (uint64_t)__double_as_longlong(__dadd_rn(__longlong_as_double((long long)a), __longlong_as_double((long long)b))))
        KQ_BIN(OP_SUB_F64, (uint64_t)__double_as_longlong(__dsub_rn(__longlong_as_double((long long)a), __longlong_as_double((long long)b))))
        KQ_BIN(OP_MUL_F64, (uint64_t)__double_as_longlong(__dmul_rn(__longlong_as_double((long long)a), __longlong_as_double((long long)b))))
        KQ_BIN(OP_DIV_F64, (uint64_t)__double_as_longlong(__ddiv_rn(__longlong_as_double((long long)a), __longlong_as_double((long long)b))))
#undef KQ_BIN
        case OP_DIV_I64:
            if constexpr (SP >= 2) {
                uint32_t both = st.ok[U] & st.ok[T];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    long long a = (long long)st.v[U][r], b = (long long)st.v[T][r];
                    long long q = 0;
                    if (b == 0) {
                        if ((both & rc.active) >> r & 1u) atomicOr(rc.err, 1u);   // KQ_DEV_ERR_DIV0 (rule E4)
                    } else if (b == -1) q = (long long)(0ULL - (unsigned long long)a);   // JVM ldiv wraps
                    else q = a / b;
                    st.v[U][r] = (uint64_t)q;
                }
                st.ok[U] = both;
            }
            break;
        case OP_CMP_I64:
            if constexpr (SP >= 2) {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    long long a = (long long)st.v[U][r], b = (long long)st.v[T][r];
                    int code = a < b ? 0 : (a == b ? 1 : 2);
                    st.v[U][r] = (in.arg >> code) & 1u;
                }
                st.ok[U] &= st.ok[T];
            }
            break;
        case OP_CMP_F64:
            if constexpr (SP >= 2) {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    double a = __longlong_as_double((long long)st.v[U][r]), b = __longlong_as_double((long long)st.v[T][r]);
                    int code = a < b ? 0 : (a == b ? 1 : (a > b ? 2 : 3));
                    st.v[U][r] = (in.arg >> code) & 1u;
                }
                st.ok[U] &= st.ok[T];
            }
            break;
        case OP_AND:               // SQL three-valued logic (rule E3)
            if constexpr (SP >= 2) {
                uint32_t at = 0, bt = 0;
#pragma unroll
                for (int r = 0; r < R; r++) { at |= (uint32_t)(st.v[U][r] & 1u) << r; bt |= (uint32_t)(st.v[T][r] & 1u) << r; }
                uint32_t oa = st.ok[U], ob = st.ok[T];
                uint32_t f = (oa & ~at) | (ob & ~bt), t = (oa & at) & (ob & bt);
#pragma unroll
                for (int r = 0; r < R; r++) st.v[U][r] = (t >> r) & 1u;
                st.ok[U] = (t | f) & rc.inr;
            }
            break;
        case OP_OR:
            if constexpr (SP >= 2) {
                uint32_t at = 0, bt = 0;
#pragma unroll
                for (int r = 0; r < R; r++) { at |= (uint32_t)(st.v[U][r] & 1u) << r; bt |= (uint32_t)(st.v[T][r] & 1u) << r; }
                uint32_t oa = st.ok[U], ob = st.ok[T];
                uint32_t t = (oa & at) | (ob & bt), f = (oa & ~at) & (ob & ~bt);
#pragma unroll
                for (int r = 0; r < R; r++) st.v[U][r] = (t >> r) & 1u;
                st.ok[U] = (t | f) & rc.inr;
            }
            break;
        case OP_UTF8_CMP_LIT:
            if constexpr (SP < D) {
                const DCol& c = P.cols[in.arg & 31];
                uint64_t L = P.lit[(in.arg >> 5) & 31];
                const uint8_t* q = P.pool + (uint32_t)(L >> 32);
                int qn = (int)(uint32_t)L;
                uint32_t mask = in.arg >> 10;
                uint32_t ok = load_valid(c, rc);
                const uint8_t* bytes = reinterpret_cast<const uint8_t*>(c.data);
#pragma unroll
                for (int r = 0; r < R; r++) {
                    uint64_t res = 0;
                    if ((ok >> r) & 1u) {
                        int64_t row = rc.row0(r >> 1) + (r & 1);
                        int a = __ldg(c.offsets + row), b = __ldg(c.offsets + row + 1);
                        int code;
                        if ((mask == CM_EQ || mask == CM_NE) && (b - a) != qn) code = 2;
                        else code = utf8_cmp3(bytes + a, b - a, q, qn);
                        res = (mask >> code) & 1u;
                    }
                    st.v[SP][r] = res;
                }
                st.ok[SP] = ok;
            }
            break;
        case OP_UTF8_CMP_COL:
            if constexpr (SP < D) {
                const DCol& c = P.cols[in.arg & 31];
                const DCol& d = P.cols[(in.arg >> 5) & 31];
                uint32_t mask = in.arg >> 10;
                uint32_t ok = load_valid(c, rc) & load_valid(d, rc);
#pragma unroll
                for (int r = 0; r < R; r++) {
                    uint64_t res = 0;
                    if ((ok >> r) & 1u) {
                        int64_t row = rc.row0(r >> 1) + (r & 1);
                        int a = __ldg(c.offsets + row), b = __ldg(c.offsets + row + 1);
                        int a2 = __ldg(d.offsets + row), b2 = __ldg(d.offsets + row + 1);
                        int code = utf8_cmp3(reinterpret_cast<const uint8_t*>(c.data) + a, b - a,
                                             reinterpret_cast<const uint8_t*>(d.data) + a2, b2 - a2);
                        res = (mask >> code) & 1u;
                    }
                    st.v[SP][r] = res;
                }
                st.ok[SP] = ok;
            }
            break;
        case OP_UTF8_PACK:
            if constexpr (SP < D) {
                const DCol& c = P.cols[in.arg];
                uint32_t ok = load_valid(c, rc);
                const uint8_t* bytes = reinterpret_cast<const uint8_t*>(c.data);
#pragma unroll
                for (int r = 0; r < R; r++) {
                    uint64_t key = 0;
                    if ((ok >> r) & 1u) {
                        int64_t row = rc.row0(r >> 1) + (r & 1);
                        int a = __ldg(c.offsets + row), b = __ldg(c.offsets + row + 1);
                        int len = b - a;
                        if (len > 7) { if ((rc.active >> r) & 1u) atomicOr(rc.err, 2u); len = 7; }   // KQ_DEV_ERR_LONG_KEY
                        for (int i = 0; i < len; i++) key |= (uint64_t)__ldg(bytes + a + i) << (8 * i);
                        key |= (uint64_t)len << 56;
                    }
                    st.v[SP][r] = key;
                }
                st.ok[SP] = ok;
            }
            break;
        case OP_UTF8_TO_F64:
            if constexpr (SP < D) {
                const DCol& c = P.cols[in.arg];
                uint32_t ok = load_valid(c, rc);
                const uint8_t* bytes = reinterpret_cast<const uint8_t*>(c.data);
#pragma unroll
                for (int r = 0; r < R; r++) {
                    double v = 0.0;
                    if ((ok >> r) & 1u) {
                        int64_t row = rc.row0(r >> 1) + (r & 1);
                        int a = __ldg(c.offsets + row), b = __ldg(c.offsets + row + 1);
                        int pe = parse_f64(bytes + a, b - a, v);
                        if (pe && ((rc.active >> r) & 1u)) atomicOr(rc.err, pe == 1 ? 4u : 8u);   // NumberFormatException / unsupported range
                    }
                    st.v[SP][r] = (uint64_t)__double_as_longlong(v);
                }
                st.ok[SP] = ok;
            }
            break;
        case OP_I64_TO_F64:
            if constexpr (SP >= 1) {
#pragma unroll
                for (int r = 0; r < R; r++) st.v[T][r] = (uint64_t)__double_as_longlong((double)(long long)st.v[T][r]);
            }
            break;
        case OP_SET_SEL: if constexpr (SP >= 1) sink.set_sel(st.v[T], st.ok[T], rc); break;
        case OP_EMIT: if constexpr (SP >= 1) sink.emit(in.arg, st.v[T], st.ok[T], rc); break;
        case OP_SET_KEY: if constexpr (SP >= 1) sink.set_key(in.arg, st.v[T], st.ok[T], rc); break;
        case OP_SET_IN: if constexpr (SP >= 1) sink.set_in(in.arg, st.v[T], st.ok[T], rc); break;
        default: break;
    }
}

// Run instructions [pc, pc_end). in.sp is the stack pointer BEFORE the instruction.
template <class Sink>
__device__ __forceinline__ void run(const Program& P, int pc, int pc_end, Stack& st, RowCtx& rc, Sink& sink) {
#pragma unroll 1
    for (; pc < pc_end; ++pc) {
        const Insn in = P.insn[pc];
        switch (in.sp) {
            case 0: step<0>(P, in, st, rc, sink); break;
            case 1: step<1>(P, in, st, rc, sink); break;
            case 2: step<2>(P, in, st, rc, sink); break;
            case 3: step<3>(P, in, st, rc, sink); break;
            case 4: step<4>(P, in, st, rc, sink); break;
            case 5: step<5>(P, in, st, rc, sink); break;
            case 6: step<6>(P, in, st, rc, sink); break;
            default: break;
        }
    }
}

__device__ __forceinline__ void rowctx_init(RowCtx& rc, int64_t tile, int64_t n, uint32_t* err) {
    int warp = threadIdx.x >> 5;
    rc.lane = threadIdx.x & 31;
    rc.n = n;
    int64_t tile_base = tile * TILE;
    rc.warp_base = tile_base + (int64_t)warp * WARP_ROWS;
    rc.full = tile_base + TILE <= n;
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < NCHUNK; j++) {
        int64_t r0 = rc.row0(j);
        if (r0 < n) m |= 1u << (2 * j);
        if (r0 + 1 < n) m |= 2u << (2 * j);
    }
    rc.inr = rc.full ? RMASK : m;
    rc.active = rc.inr;
    rc.err = err;
}

}  // namespace kq
