// kq_vm.cuh — the fused expression evaluator shared by every operator kernel.
//
// An expression tree (Expression.evaluate, Main.kt:448-450; the reference materialises one Arrow
// vector per node, Main.kt:780-803) is flattened on the host into a short program and executed by
// ONE kernel. The program is uniform across the grid, so instruction dispatch is a divergence-free
// branch whose cost is amortised over the R rows each thread carries.
//
// The machine is an accumulator machine held entirely in registers: ACC (R values + validity mask),
// one operand register TMP, and a small save stack. An instruction is {fetch TMP from a column /
// literal / saved slot / Utf8 operator} followed by {ACC = ACC op TMP}. The host compiler evaluates
// the non-leaf side of a binary node into ACC and feeds the leaf side straight from the column, so
// typical expressions (a*b+c, a>k AND b<m) never touch the save stack, and every operator body
// exists exactly once in the SASS (small code, no spills).
//
// Bool values are kept as an R-bit truth mask in acc[0] / tmp[0] (bit r = row r), so AND / OR / NOT-like
// steps and the predicate -> selection hand-off are a few bit operations per tile, not per row.
//
// Row ownership inside a tile of TILE = BLOCK*R rows: warp w owns rows [w*32R, (w+1)*32R); inside
// that, chunk j (of R/2) covers 64 rows and lane l owns the adjacent pair (2l, 2l+1). One pair of
// 8-byte values is one 128-bit access; one pair of validity bits comes from one 32-bit word.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace kq {

#ifndef KQ_R
#define KQ_R 4
#endif
constexpr int R = KQ_R;              // rows per thread (even); each kernel file picks its own before including this header
constexpr int NCHUNK = R / 2;
constexpr int WARP_ROWS = 32 * R;    // rows one warp owns per tile; a tile is (consumer warps) x WARP_ROWS rows
constexpr int DS = 3;                // save slots (nesting depth of binary nodes with two non-leaf operands)
constexpr int MAX_COLS = 16;
constexpr int MAX_INSN = 96;
constexpr int MAX_LIT = 32;
constexpr int LITPOOL = 256;
constexpr int MAX_OUT = 8;
constexpr int MAX_KEYS = 4;
constexpr int MAX_INPUTS = 6;
constexpr uint32_t RMASK = (1u << R) - 1u;

// Operand sources of the load / Bool instructions
enum Src : uint8_t {
    S_NONE = 0, S_COL64, S_COL32, S_COLBIT, S_LIT, S_NULL, S_VALID, S_STACK,
    S_UTF8_CMP_LIT,                  // a = col, b = lit | mask << 8 : Bool
    S_UTF8_CMP_COL,                  // a = colA, b = colB | mask << 8 : Bool
    S_UTF8_PACK,                     // a = col: short string (<= 7 bytes) -> packed u64 group key
    S_UTF8_F64                       // a = col: CastExpression Utf8 -> Float64 (Main.kt:772-805)
};
// 64-bit binary primitives are specialised per operand mode, like the primitives of a vectorised
// engine: ACC = X op Y with X, Y taken straight from a staged column, a literal, ACC or a save slot.
enum Mode : uint8_t {
    M_ACC_COL = 0, M_ACC_LIT, M_COL_ACC, M_LIT_ACC, M_COL_COL, M_COL_LIT, M_LIT_COL, M_STK_ACC, NMODES
};
enum Bin : uint8_t {
    B_ADD_I64 = 0, B_SUB_I64, B_MUL_I64, B_DIV_I64, B_ADD_F64, B_SUB_F64, B_MUL_F64, B_DIV_F64, B_CMP_I64, B_CMP_F64, NBIN
};
// Opcodes. A binary primitive is op = O_BIN + bin with the operand mode in `src`; a = first
// column/literal/slot, (b & 0xff) = second column/literal, (b >> 8) = comparison truth mask.
enum Op : uint8_t {
    O_END = 0, O_LOAD, O_PUSH,
    O_AND, O_OR, O_CMP_BOOL,         // ACC (truth mask) op operand fetched through `src`
    O_I64_TO_F64,
    O_SET_SEL, O_EMIT, O_SET_KEY, O_SET_IN,
    O_BIN                            // first of NBIN binary primitives
};

// comparison truth masks: bit0 = lt, bit1 = eq, bit2 = gt, bit3 = unordered (NaN)
constexpr uint32_t CM_EQ = 0x2, CM_NE = 0xD, CM_LT = 0x1, CM_LE = 0x3, CM_GT = 0x4, CM_GE = 0x6;

struct Insn { uint8_t op, src; uint16_t a; uint32_t b; };

struct DCol {
    const void* data;
    const uint32_t* validity;
    const int32_t* offsets;
    int32_t s_data, s_valid, s_off, _pad;   // byte offsets of the staged copies inside a pipeline stage (-1: not staged)
};

struct Program {
    int32_t ninsn;
    int32_t ncols;
    Insn insn[MAX_INSN];
    uint64_t lit[MAX_LIT];          // utf8 literal: (pool offset << 32) | length
    uint8_t pool[LITPOOL];
    DCol cols[MAX_COLS];
};

struct Vm {
    uint64_t acc[R]; uint32_t aok;   // accumulator + R-bit validity mask
    // save slots are separate members on purpose: an array indexed by the instruction would be
    // demoted to local memory by the compiler
    uint64_t s0[R], s1[R], s2[R]; uint32_t k0, k1, k2;
};
static_assert(DS == 3, "Vm has exactly three save slots");

// Per-thread view of the current tile.
struct RowCtx {
    int64_t n;                       // rows in the batch
    int64_t warp_base;               // first row of this warp in the batch
    int lane;
    bool full;                       // whole tile < n: no bounds checks
    uint32_t inr;                    // R-bit mask of owned rows that are < n
    uint32_t active;                 // rows whose errors count (in range and selected)
    uint32_t* err;
    const unsigned char* stage;      // shared-memory stage holding this tile's staged buffers (kq_pipe.cuh)
    int wrow;                        // first row of this warp inside the tile
    __device__ __forceinline__ int64_t row0(int j) const { return warp_base + j * 64 + lane * 2; }
    __device__ __forceinline__ int trow0(int j) const { return wrow + j * 64 + lane * 2; }   // row inside the tile
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
// Staged buffers are read from the tile's shared-memory stage (filled by TMA bulk copies); buffers
// that did not fit the stage budget are read from global memory with 128-bit coalesced loads.
static __device__ __noinline__ uint32_t load_bits_g(const uint32_t* bits, const RowCtx& rc) {
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
__device__ __forceinline__ uint32_t load_bits_s(const unsigned char* sbits, const RowCtx& rc) {
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < NCHUNK; j++) {
        int t0 = rc.trow0(j);
        uint32_t w = *reinterpret_cast<const uint32_t*>(sbits + (t0 >> 5) * 4);
        m |= ((w >> (t0 & 31)) & 3u) << (2 * j);
    }
    return m;
}
__device__ __forceinline__ uint32_t load_valid(const DCol& c, const RowCtx& rc) {
    if (!c.validity) return rc.inr;
    return (c.s_valid >= 0 ? load_bits_s(rc.stage + c.s_valid, rc) : load_bits_g(c.validity, rc)) & rc.inr;
}
__device__ __forceinline__ void load64(const DCol& c, const RowCtx& rc, uint64_t (&v)[R], uint32_t& ok) {
    if (c.s_data >= 0) {
        const unsigned char* p = rc.stage + c.s_data;
#pragma unroll
        for (int j = 0; j < NCHUNK; j++) {
            uint4 q = *reinterpret_cast<const uint4*>(p + rc.trow0(j) * 8);
            v[2 * j] = (uint64_t)q.x | ((uint64_t)q.y << 32);
            v[2 * j + 1] = (uint64_t)q.z | ((uint64_t)q.w << 32);
        }
    } else {
        const uint4* p = reinterpret_cast<const uint4*>(c.data);
#pragma unroll
        for (int j = 0; j < NCHUNK; j++) {
            int64_t r0 = rc.row0(j);
            uint4 q = make_uint4(0, 0, 0, 0);
            if (rc.full || r0 < rc.n) q = ldg_nc_v4(p + (r0 >> 1));
            v[2 * j] = (uint64_t)q.x | ((uint64_t)q.y << 32);
            v[2 * j + 1] = (uint64_t)q.z | ((uint64_t)q.w << 32);
        }
    }
    ok = load_valid(c, rc);
}
__device__ __forceinline__ void load32(const DCol& c, const RowCtx& rc, uint64_t (&v)[R], uint32_t& ok) {
    if (c.s_data >= 0) {
        const unsigned char* p = rc.stage + c.s_data;
#pragma unroll
        for (int j = 0; j < NCHUNK; j++) {
            uint2 q = *reinterpret_cast<const uint2*>(p + rc.trow0(j) * 4);
            v[2 * j] = (uint64_t)(int64_t)(int32_t)q.x;
            v[2 * j + 1] = (uint64_t)(int64_t)(int32_t)q.y;
        }
    } else {
        const uint2* p = reinterpret_cast<const uint2*>(c.data);
#pragma unroll
        for (int j = 0; j < NCHUNK; j++) {
            int64_t r0 = rc.row0(j);
            uint2 q = make_uint2(0, 0);
            if (rc.full || r0 < rc.n) q = ldg_nc_v2(p + (r0 >> 1));
            v[2 * j] = (uint64_t)(int64_t)(int32_t)q.x;
            v[2 * j + 1] = (uint64_t)(int64_t)(int32_t)q.y;
        }
    }
    ok = load_valid(c, rc);
}
__device__ __forceinline__ void loadbit(const DCol& c, const RowCtx& rc, uint64_t (&v)[R], uint32_t& ok) {
    v[0] = c.s_data >= 0 ? load_bits_s(rc.stage + c.s_data, rc) : load_bits_g(reinterpret_cast<const uint32_t*>(c.data), rc);
    ok = load_valid(c, rc);
}
// byte range [a, b) of the string in owned row r (offsets from the stage when staged)
__device__ __forceinline__ void utf8_bounds(const DCol& c, const RowCtx& rc, int r, int& a, int& b) {
    if (c.s_off >= 0) {
        const int32_t* o = reinterpret_cast<const int32_t*>(rc.stage + c.s_off) + rc.trow0(r >> 1) + (r & 1);
        a = o[0]; b = o[1];
    } else {
        int64_t row = rc.row0(r >> 1) + (r & 1);
        a = __ldg(c.offsets + row); b = __ldg(c.offsets + row + 1);
    }
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

// ---- Utf8 operators. The per-row bodies are out of line (big and rare: they must not bloat the main
// loop or its register allocation); they take and return scalars so the VM registers stay registers.
static __device__ __noinline__ uint32_t utf8_cmp_row(const uint8_t* p, int pn, const uint8_t* q, int qn, uint32_t mask) {
    int code;
    if ((mask == CM_EQ || mask == CM_NE) && pn != qn) code = 2;
    else code = utf8_cmp3(p, pn, q, qn);
    return (mask >> code) & 1u;
}
static __device__ __noinline__ unsigned long long utf8_pack_row(const uint8_t* p, int len) {
    unsigned long long key = 0;
    for (int i = 0; i < len; i++) key |= (unsigned long long)__ldg(p + i) << (8 * i);
    return key | ((unsigned long long)len << 56);
}
static __device__ __noinline__ double parse_f64_row(const uint8_t* p, int n, uint32_t* err, bool active) {
    double v = 0.0;
    int pe = parse_f64(p, n, v);
    if (pe && active) atomicOr(err, pe == 1 ? 4u : 8u);     // NumberFormatException / unsupported range
    return v;
}
__device__ __forceinline__ void utf8_cmp_lit(const Program& P, const Insn in, const RowCtx& rc, uint64_t (&out)[R], uint32_t& ok_out) {
    const DCol& c = P.cols[in.a];
    const uint64_t L = P.lit[in.b & 0xff];
    const uint8_t* q = P.pool + (uint32_t)(L >> 32);
    const int qn = (int)(uint32_t)L;
    const uint32_t mask = in.b >> 8;
    const uint32_t ok = load_valid(c, rc);
    const uint8_t* bytes = reinterpret_cast<const uint8_t*>(c.data);
    uint32_t m = 0;
#pragma unroll
    for (int r = 0; r < R; r++)
        if ((ok >> r) & 1u) { int a, b; utf8_bounds(c, rc, r, a, b); m |= utf8_cmp_row(bytes + a, b - a, q, qn, mask) << r; }
    out[0] = m;
    ok_out = ok;
}
__device__ __forceinline__ void utf8_cmp_col(const Program& P, const Insn in, const RowCtx& rc, uint64_t (&out)[R], uint32_t& ok_out) {
    const DCol& c = P.cols[in.a];
    const DCol& d = P.cols[in.b & 0xff];
    const uint32_t mask = in.b >> 8;
    const uint32_t ok = load_valid(c, rc) & load_valid(d, rc);
    uint32_t m = 0;
#pragma unroll
    for (int r = 0; r < R; r++)
        if ((ok >> r) & 1u) {
            int a, b, a2, b2; utf8_bounds(c, rc, r, a, b); utf8_bounds(d, rc, r, a2, b2);
            m |= utf8_cmp_row(reinterpret_cast<const uint8_t*>(c.data) + a, b - a, reinterpret_cast<const uint8_t*>(d.data) + a2, b2 - a2, mask) << r;
        }
    out[0] = m;
    ok_out = ok;
}
__device__ __forceinline__ void utf8_pack(const DCol& c, const RowCtx& rc, uint64_t (&out)[R], uint32_t& ok_out) {
    const uint32_t ok = load_valid(c, rc);
    const uint8_t* bytes = reinterpret_cast<const uint8_t*>(c.data);
#pragma unroll
    for (int r = 0; r < R; r++) {
        uint64_t key = 0;
        if ((ok >> r) & 1u) {
            int a, b; utf8_bounds(c, rc, r, a, b);
            int len = b - a;
            if (len > 7) { if ((rc.active >> r) & 1u) atomicOr(rc.err, 2u); len = 7; }   // KQ_DEV_ERR_LONG_KEY
            key = utf8_pack_row(bytes + a, len);
        }
        out[r] = key;
    }
    ok_out = ok;
}
__device__ __forceinline__ void utf8_to_f64(const Program& P, const Insn in, const RowCtx& rc, uint64_t (&out)[R], uint32_t& ok_out) {
    const DCol& c = P.cols[in.a];
    const uint32_t ok = load_valid(c, rc);
    const uint8_t* bytes = reinterpret_cast<const uint8_t*>(c.data);
#pragma unroll
    for (int r = 0; r < R; r++) {
        double v = 0.0;
        if ((ok >> r) & 1u) { int a, b; utf8_bounds(c, rc, r, a, b); v = parse_f64_row(bytes + a, b - a, rc.err, (rc.active >> r) & 1u); }
        out[r] = (uint64_t)__double_as_longlong(v);
    }
    ok_out = ok;
}


// ---- operand fetch for the binary primitives ------------------------------------------------------------
__device__ __forceinline__ void fetch_lit(const Program& P, int idx, const RowCtx& rc, uint64_t (&v)[R], uint32_t& ok) {
    const uint64_t x = P.lit[idx];
#pragma unroll
    for (int r = 0; r < R; r++) v[r] = x;
    ok = rc.inr;
}
__device__ __forceinline__ void fetch_acc(const Vm& vm, uint64_t (&v)[R], uint32_t& ok) {
#pragma unroll
    for (int r = 0; r < R; r++) v[r] = vm.acc[r];
    ok = vm.aok;
}
__device__ __forceinline__ void fetch_stk(const Vm& vm, int slot, uint64_t (&v)[R], uint32_t& ok) {
    switch (slot) {
#define KQ_POP(d, S, K) case d: _Pragma("unroll") for (int r = 0; r < R; r++) v[r] = vm.S[r]; ok = vm.K; break;
        KQ_POP(0, s0, k0) KQ_POP(1, s1, k1)
#undef KQ_POP
        default:
#pragma unroll
            for (int r = 0; r < R; r++) v[r] = vm.s2[r];
            ok = vm.k2;
            break;
    }
}
template <int MODE>
__device__ __forceinline__ void fetch2(const Program& P, const Insn in, const Vm& vm, const RowCtx& rc,
                                       uint64_t (&x)[R], uint32_t& okx, uint64_t (&y)[R], uint32_t& oky) {
    if constexpr (MODE == M_ACC_COL) { fetch_acc(vm, x, okx); load64(P.cols[in.a], rc, y, oky); }
    else if constexpr (MODE == M_ACC_LIT) { fetch_acc(vm, x, okx); fetch_lit(P, in.a, rc, y, oky); }
    else if constexpr (MODE == M_COL_ACC) { load64(P.cols[in.a], rc, x, okx); fetch_acc(vm, y, oky); }
    else if constexpr (MODE == M_LIT_ACC) { fetch_lit(P, in.a, rc, x, okx); fetch_acc(vm, y, oky); }
    else if constexpr (MODE == M_COL_COL) { load64(P.cols[in.a], rc, x, okx); load64(P.cols[in.b & 0xff], rc, y, oky); }
    else if constexpr (MODE == M_COL_LIT) { load64(P.cols[in.a], rc, x, okx); fetch_lit(P, in.b & 0xff, rc, y, oky); }
    else if constexpr (MODE == M_LIT_COL) { fetch_lit(P, in.a, rc, x, okx); load64(P.cols[in.b & 0xff], rc, y, oky); }
    else { fetch_stk(vm, in.a, x, okx); fetch_acc(vm, y, oky); }
}

__device__ __forceinline__ double as_f64(uint64_t x) { return __longlong_as_double((long long)x); }
__device__ __forceinline__ uint64_t as_u64(double x) { return (uint64_t)__double_as_longlong(x); }

template <int BIN>
__device__ __forceinline__ void apply(const Insn in, Vm& vm, RowCtx& rc, const uint64_t (&x)[R], const uint64_t (&y)[R], uint32_t both) {
    if constexpr (BIN == B_CMP_I64 || BIN == B_CMP_F64) {
        // one relation per instruction; the switch is uniform (IEEE: every relation but != is false on NaN)
        uint32_t m = 0;
#define KQ_REL(REL)                                                                                   \
        _Pragma("unroll") for (int r = 0; r < R; r++) {                                               \
            if constexpr (BIN == B_CMP_I64) { const long long a = (long long)x[r], b = (long long)y[r]; if (REL) m |= 1u << r; } \
            else { const double a = as_f64(x[r]), b = as_f64(y[r]); if (REL) m |= 1u << r; }          \
        }
        switch (in.b >> 8) {
            case CM_EQ: KQ_REL(a == b) break;
            case CM_NE: KQ_REL(a != b) break;
            case CM_LT: KQ_REL(a < b) break;
            case CM_LE: KQ_REL(a <= b) break;
            case CM_GT: KQ_REL(a > b) break;
            default: KQ_REL(a >= b) break;
        }
#undef KQ_REL
        vm.acc[0] = m;
    } else if constexpr (BIN == B_DIV_I64) {
#pragma unroll
        for (int r = 0; r < R; r++) {
            const long long a = (long long)x[r], b = (long long)y[r];
            long long q = 0;
            if (b == 0) {
                if ((both & rc.active) >> r & 1u) atomicOr(rc.err, 1u);               // KQ_DEV_ERR_DIV0 (rule E4)
            } else if (b == -1) q = (long long)(0ULL - (unsigned long long)a);        // JVM ldiv wraps
            else q = a / b;
            vm.acc[r] = (uint64_t)q;
        }
    } else {
#pragma unroll
        for (int r = 0; r < R; r++) {
            const uint64_t a = x[r], b = y[r];
            uint64_t o;
            // Float64: separately rounded IEEE operations, never contracted into FMA (rule E4)
            if constexpr (BIN == B_ADD_I64) o = a + b;
            else if constexpr (BIN == B_SUB_I64) o = a - b;
            else if constexpr (BIN == B_MUL_I64) o = a * b;
            else if constexpr (BIN == B_ADD_F64) o = as_u64(__dadd_rn(as_f64(a), as_f64(b)));
            else if constexpr (BIN == B_SUB_F64) o = as_u64(__dsub_rn(as_f64(a), as_f64(b)));
            else if constexpr (BIN == B_MUL_F64) o = as_u64(__dmul_rn(as_f64(a), as_f64(b)));
            else o = as_u64(__ddiv_rn(as_f64(a), as_f64(b)));
            vm.acc[r] = o;
        }
    }
    vm.aok = both;
}

// Run instructions [pc, pc_end). P should live in shared memory (the kernels copy it there once):
// the instruction stream and the column table are then read with uniform LDS instead of indexed
// constant-bank loads.
template <class Sink>
__device__ __forceinline__ void run(const Program& P, int pc, int pc_end, Vm& vm, RowCtx& rc, Sink& sink) {
#pragma unroll 1
    for (; pc < pc_end; ++pc) {
        const Insn in = P.insn[pc];
        if (in.op >= O_BIN) {
            // binary primitive: operands by mode (8 small bodies), then the operation (10 small bodies)
            uint64_t x[R], y[R];
            uint32_t okx, oky;
            switch (in.src) {
                case M_ACC_COL: fetch2<M_ACC_COL>(P, in, vm, rc, x, okx, y, oky); break;
                case M_ACC_LIT: fetch2<M_ACC_LIT>(P, in, vm, rc, x, okx, y, oky); break;
                case M_COL_ACC: fetch2<M_COL_ACC>(P, in, vm, rc, x, okx, y, oky); break;
                case M_LIT_ACC: fetch2<M_LIT_ACC>(P, in, vm, rc, x, okx, y, oky); break;
                case M_COL_COL: fetch2<M_COL_COL>(P, in, vm, rc, x, okx, y, oky); break;
                case M_COL_LIT: fetch2<M_COL_LIT>(P, in, vm, rc, x, okx, y, oky); break;
                case M_LIT_COL: fetch2<M_LIT_COL>(P, in, vm, rc, x, okx, y, oky); break;
                default: fetch2<M_STK_ACC>(P, in, vm, rc, x, okx, y, oky); break;
            }
            const uint32_t both = okx & oky;
            switch (in.op - O_BIN) {
                case B_ADD_I64: apply<B_ADD_I64>(in, vm, rc, x, y, both); break;
                case B_SUB_I64: apply<B_SUB_I64>(in, vm, rc, x, y, both); break;
                case B_MUL_I64: apply<B_MUL_I64>(in, vm, rc, x, y, both); break;
                case B_DIV_I64: apply<B_DIV_I64>(in, vm, rc, x, y, both); break;
                case B_ADD_F64: apply<B_ADD_F64>(in, vm, rc, x, y, both); break;
                case B_SUB_F64: apply<B_SUB_F64>(in, vm, rc, x, y, both); break;
                case B_MUL_F64: apply<B_MUL_F64>(in, vm, rc, x, y, both); break;
                case B_DIV_F64: apply<B_DIV_F64>(in, vm, rc, x, y, both); break;
                case B_CMP_I64: apply<B_CMP_I64>(in, vm, rc, x, y, both); break;
                default: apply<B_CMP_F64>(in, vm, rc, x, y, both); break;
            }
            continue;
        }
        switch (in.op) {
            case O_LOAD:
                switch (in.src) {
                    case S_COL64: load64(P.cols[in.a], rc, vm.acc, vm.aok); break;
                    case S_COL32: load32(P.cols[in.a], rc, vm.acc, vm.aok); break;
                    case S_COLBIT: loadbit(P.cols[in.a], rc, vm.acc, vm.aok); break;
                    case S_LIT: fetch_lit(P, in.a, rc, vm.acc, vm.aok); break;
                    case S_NULL:
#pragma unroll
                        for (int r = 0; r < R; r++) vm.acc[r] = 0;
                        vm.aok = 0;
                        break;
                    case S_VALID:              // COUNT(col) of any type only needs the validity bit
#pragma unroll
                        for (int r = 0; r < R; r++) vm.acc[r] = 1;
                        vm.aok = load_valid(P.cols[in.a], rc);
                        break;
                    case S_UTF8_CMP_LIT: utf8_cmp_lit(P, in, rc, vm.acc, vm.aok); break;
                    case S_UTF8_CMP_COL: utf8_cmp_col(P, in, rc, vm.acc, vm.aok); break;
                    case S_UTF8_PACK: utf8_pack(P.cols[in.a], rc, vm.acc, vm.aok); break;
                    case S_UTF8_F64: utf8_to_f64(P, in, rc, vm.acc, vm.aok); break;
                    default: break;
                }
                break;
            case O_PUSH:
                switch (in.a) {
#define KQ_PUSH(d, S, K) case d: _Pragma("unroll") for (int r = 0; r < R; r++) vm.S[r] = vm.acc[r]; vm.K = vm.aok; break;
                    KQ_PUSH(0, s0, k0) KQ_PUSH(1, s1, k1)
#undef KQ_PUSH
                    default:
#pragma unroll
                        for (int r = 0; r < R; r++) vm.s2[r] = vm.acc[r];
                        vm.k2 = vm.aok;
                        break;
                }
                break;
            case O_AND: case O_OR: case O_CMP_BOOL: {
                // Bool operand (truth mask + validity) through `src`: a Bool column, a literal or a save slot
                uint32_t bt = 0, ob = 0;
                switch (in.src) {
                    case S_COLBIT: { uint64_t t[R]; loadbit(P.cols[in.a], rc, t, ob); bt = (uint32_t)t[0]; break; }
                    case S_LIT: bt = (uint32_t)P.lit[in.a]; ob = rc.inr; break;
                    case S_NULL: bt = 0; ob = 0; break;
                    default: { uint64_t t[R]; fetch_stk(vm, in.a, t, ob); bt = (uint32_t)t[0]; break; }
                }
                const uint32_t at = (uint32_t)vm.acc[0], oa = vm.aok;
                if (in.op == O_AND) {          // SQL three-valued logic (rule E3) on truth masks
                    const uint32_t f = (oa & ~at) | (ob & ~bt), t = (oa & at) & (ob & bt);
                    vm.acc[0] = t; vm.aok = (t | f) & rc.inr;
                } else if (in.op == O_OR) {
                    const uint32_t t = (oa & at) | (ob & bt), f = (oa & ~at) & (ob & ~bt);
                    vm.acc[0] = t; vm.aok = (t | f) & rc.inr;
                } else {                       // ACC ? operand with truth mask b >> 8
                    const uint32_t ltm = ~at & bt, eqm = ~(at ^ bt), gtm = at & ~bt, m = in.b >> 8;
                    vm.acc[0] = (((m & 1u) ? ltm : 0u) | ((m & 2u) ? eqm : 0u) | ((m & 4u) ? gtm : 0u)) & RMASK;
                    vm.aok = oa & ob;
                }
                break;
            }
            case O_I64_TO_F64:
#pragma unroll
                for (int r = 0; r < R; r++) vm.acc[r] = as_u64((double)(long long)vm.acc[r]);
                break;
            case O_SET_SEL: sink.set_sel(vm.acc, vm.aok, rc); break;
            case O_EMIT: sink.emit(in.a, vm.acc, vm.aok, rc); break;
            case O_SET_KEY: sink.set_key(in.a, in.b != 0, vm.acc, vm.aok, rc); break;     // b: ACC is a Bool truth mask
            case O_SET_IN: sink.set_in(in.a, in.b != 0, vm.acc, vm.aok, rc); break;
            default: break;
        }
    }
}

__device__ __forceinline__ void rowctx_init(RowCtx& rc, int64_t tile, int tile_rows, int64_t n, uint32_t* err, const unsigned char* stage) {
    int warp = threadIdx.x >> 5;
    rc.lane = threadIdx.x & 31;
    rc.n = n;
    rc.stage = stage;
    rc.wrow = warp * WARP_ROWS;
    int64_t tile_base = tile * tile_rows;
    rc.warp_base = tile_base + (int64_t)warp * WARP_ROWS;
    rc.full = tile_base + tile_rows <= n;
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
