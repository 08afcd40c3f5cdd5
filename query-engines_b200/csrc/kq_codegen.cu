// kq_codegen.cu — expression handles (kq_expr_*) and the tree -> CUDA source generator.
//
// Type rules follow the oracle's BinaryExpression/CastExpression (rules E2-E4, R5): operand types
// must match, AND/OR need Bool, math needs Int64/Float64, the only cast target is Float64. Type
// errors are IllegalStateException (KQ_ERR_ILLEGAL_STATE), exactly where the reference would throw
// (Main.kt:792, 799).
//
// Generated code conventions (kq_rt.cuh): inside a body `q` is the QArgs block, `rc` the RowCtx;
// 64-bit values are uint64_t[R] arrays, Bool values uint32_t truth masks, validity R-bit masks.
// Column types, nullability and staging offsets are baked into the text; pointers and literal
// VALUES are read from `q`, so one compiled kernel serves every query of the same shape.
#include <algorithm>
#include <cstring>

#include "kq_codegen.h"

using namespace kq;

extern "C" {

kq_expr* kq_expr_column(int i) { auto* e = new kq_expr(); e->kind = KQ_EX_COL; e->col = i; return e; }
static kq_expr* lit(int type) { auto* e = new kq_expr(); e->kind = KQ_EX_LIT; e->type = type; return e; }
kq_expr* kq_expr_literal_f64(double v) { auto* e = lit(KQ_F64); e->f = v; return e; }
kq_expr* kq_expr_literal_i64(int64_t v) { auto* e = lit(KQ_I64); e->i = v; return e; }
kq_expr* kq_expr_literal_bool(int v) { auto* e = lit(KQ_BOOL); e->i = v != 0; return e; }
kq_expr* kq_expr_literal_date32(int32_t v) { auto* e = lit(KQ_DATE32); e->i = v; return e; }
kq_expr* kq_expr_literal_utf8(const char* bytes, int32_t len) { auto* e = lit(KQ_UTF8); e->s.assign(bytes ? bytes : "", (size_t)(len > 0 ? len : 0)); return e; }
kq_expr* kq_expr_literal_null(int type) { auto* e = lit(type); e->is_null = true; return e; }
kq_expr* kq_expr_binary(int op, kq_expr* l, kq_expr* r) {
    if (!l || !r) return nullptr;
    auto* e = new kq_expr(); e->kind = KQ_EX_BIN; e->op = op; e->l = l; e->r = r;
    l->rc.fetch_add(1); r->rc.fetch_add(1);
    return e;
}
kq_expr* kq_expr_cast(kq_expr* x, int type) {
    if (!x) return nullptr;
    auto* e = new kq_expr(); e->kind = KQ_EX_CAST; e->type = type; e->l = x;
    x->rc.fetch_add(1);
    return e;
}
void kq_expr_free(kq_expr* e) {
    if (!e) return;
    if (e->rc.fetch_sub(1) == 1) { kq_expr_free(e->l); kq_expr_free(e->r); delete e; }
}

}  // extern "C"

static std::string S(long long x) { return std::to_string(x); }
static const char* UNROLL = "_Pragma(\"unroll\") for (int r = 0; r < R; r++) ";

int KqCodegen::begin(kq_ctx* c, kq_batch* b) {
    ctx = c; batch = b; ncols = 0; nlit = 0; pool_used = 0; ntmp = 0;
    memset(&args, 0, sizeof args);
    for (int& x : colmap) x = -1;
    for (bool& x : col_bytes_used) x = false;
    body.clear();
    begin_body();
    if (b->cols.size() > 256) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "batches wider than 256 columns are not supported");
    return KQ_OK;
}

void KqCodegen::begin_body() {
    for (int i = 0; i < MAX_COLS; i++) col_loaded[i] = valid_loaded[i] = false;
}

std::string KqCodegen::take_body() {
    std::string s;
    s.swap(body);
    begin_body();
    return s;
}

std::string KqCodegen::tmp(const char* prefix) { return std::string(prefix) + S(ntmp++); }

int KqCodegen::use_col(int bc, int* slot) {
    if (bc < 0 || bc >= (int)batch->cols.size()) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "column index %d out of range (batch has %d fields)", bc, (int)batch->cols.size());
    if (colmap[bc] < 0) {
        if (ncols >= MAX_COLS) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "more than %d distinct columns in one kernel", MAX_COLS);
        kq_col* c = batch->cols[(size_t)bc];
        QCol& d = args.cols[ncols];
        d.data = c->data; d.validity = c->validity; d.offsets = c->offsets;
        slot_col[ncols] = bc;
        colmap[bc] = ncols++;
    }
    *slot = colmap[bc];
    return KQ_OK;
}

int KqCodegen::add_lit(uint64_t v, int* idx) {
    // no de-duplication by value: the slot layout must depend on the query SHAPE only
    if (nlit >= MAX_LIT) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "more than %d literals in one kernel", MAX_LIT);
    args.lit[nlit] = v; *idx = nlit++;
    return KQ_OK;
}
int KqCodegen::add_utf8_lit(const std::string& s, int* idx) {
    if (pool_used + (int)s.size() > LITPOOL) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "Utf8 literals exceed %d bytes", LITPOOL);
    memcpy(args.pool + pool_used, s.data(), s.size());
    uint64_t v = ((uint64_t)pool_used << 32) | (uint64_t)s.size();
    pool_used += (int)s.size();
    return add_lit(v, idx);
}

int KqCodegen::infer(const kq_expr* e, int* type, bool* nullable) {
    if (!e) return kq_fail(ctx, KQ_ERR_ILLEGAL_ARGUMENT, "null expression");
    switch (e->kind) {
        case KQ_EX_COL: {
            if (e->col < 0 || e->col >= (int)batch->cols.size())
                return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "column index %d out of range (batch has %d fields)", e->col, (int)batch->cols.size());
            kq_col* c = batch->cols[(size_t)e->col];
            *type = c->type; *nullable = c->validity != nullptr;
            return KQ_OK;
        }
        case KQ_EX_LIT: *type = e->type; *nullable = e->is_null; return KQ_OK;
        case KQ_EX_CAST: {
            int t; bool n; KQ_RET(infer(e->l, &t, &n));
            if (e->type != KQ_F64) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "Cast to type %d is not supported", e->type);   // Main.kt:799
            *type = e->type; *nullable = n;
            return KQ_OK;
        }
        case KQ_EX_BIN: {
            int lt, rt; bool ln, rn;
            KQ_RET(infer(e->l, &lt, &ln)); KQ_RET(infer(e->r, &rt, &rn));
            *nullable = ln || rn;
            *type = (e->op >= KQ_EQ && e->op <= KQ_OR) ? KQ_BOOL : lt;
            return KQ_OK;
        }
    }
    return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "Unknown expr");   // Main.kt:677
}

static uint32_t cmp_mask(int op) {
    switch (op) { case KQ_EQ: return CM_EQ; case KQ_NE: return CM_NE; case KQ_LT: return CM_LT;
                  case KQ_LE: return CM_LE; case KQ_GT: return CM_GT; default: return CM_GE; }
}
static uint32_t mirror_mask(uint32_t m) { return (m & 0xA) | ((m & 1) << 2) | ((m >> 2) & 1); }
static const char* cmp_sym(int op) {
    switch (op) { case KQ_EQ: return "=="; case KQ_NE: return "!="; case KQ_LT: return "<";
                  case KQ_LE: return "<="; case KQ_GT: return ">"; default: return ">="; }
}

std::string KqCodegen::col_valid(int slot) {
    if (!args.cols[slot].validity) return "";
    const std::string k = "k" + S(slot);
    if (!valid_loaded[slot]) {
        line("const uint32_t " + k + " = load_valid<SV" + S(slot) + ">(q.cols[" + S(slot) + "].validity, rc);");
        valid_loaded[slot] = true;
    }
    return k;
}

int KqCodegen::col_value(const kq_expr* e, KqVal* out) {
    int t; bool nl;
    KQ_RET(infer(e, &t, &nl));
    if (t == KQ_UTF8) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "a Utf8 column cannot be an operand here (only comparisons, casts, COUNT, group keys and pass-through)");
    int slot; KQ_RET(use_col(e->col, &slot));
    const std::string c = "c" + S(slot), ix = S(slot);
    if (!col_loaded[slot]) {
        if (t == KQ_BOOL) line("const uint32_t " + c + " = load_bits<SD" + ix + ">(reinterpret_cast<const uint32_t*>(q.cols[" + ix + "].data), rc);");
        else line("uint64_t " + c + "[R]; " + (t == KQ_F64 || t == KQ_I64 ? "load64" : "load32") + "<SD" + ix + ">(q.cols[" + ix + "].data, rc, " + c + ");");
        col_loaded[slot] = true;
    }
    out->v = c; out->type = t; out->scalar = false;
    out->ok = col_valid(slot);
    return KQ_OK;
}

int KqCodegen::lit_value(const kq_expr* e, KqVal* out) {
    out->type = e->type; out->scalar = true;
    if (e->type == KQ_UTF8 && !e->is_null) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "a Utf8 literal can only be compared with a Utf8 column");
    if (e->is_null) { out->v = e->type == KQ_BOOL ? "0u" : "0ULL"; out->ok = "0u"; return KQ_OK; }
    uint64_t bits;
    if (e->type == KQ_F64) memcpy(&bits, &e->f, 8);
    else if (e->type == KQ_BOOL) bits = e->i ? 0xFFFFFFFFULL : 0ULL;      // Bool values are truth masks
    else bits = (uint64_t)e->i;
    int idx; KQ_RET(add_lit(bits, &idx));
    out->v = e->type == KQ_BOOL ? "((uint32_t)q.lit[" + S(idx) + "])" : "q.lit[" + S(idx) + "]";
    out->ok = "";
    return KQ_OK;
}

static std::string and_ok(const KqVal& a, const KqVal& b) {
    if (a.ok.empty()) return b.ok;
    if (b.ok.empty()) return a.ok;
    if (a.ok == "0u" || b.ok == "0u") return "0u";
    return "(" + a.ok + " & " + b.ok + ")";
}

int KqCodegen::bin_value(const kq_expr* e, KqVal* out) {
    int lt, rt; bool ln, rn;
    KQ_RET(infer(e->l, &lt, &ln)); KQ_RET(infer(e->r, &rt, &rn));
    if (lt != rt) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "binary operand types differ (%d vs %d)", lt, rt);      // rule E2
    const int op = e->op;
    const bool is_cmp = op >= KQ_EQ && op <= KQ_GE, is_logic = op == KQ_AND || op == KQ_OR;
    if (!is_cmp && !is_logic && !(op >= KQ_ADD && op <= KQ_DIV)) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "unknown binary operator %d", op);
    if (is_logic && lt != KQ_BOOL) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "AND/OR need Bool operands");
    out->scalar = false;

    if (is_cmp && lt == KQ_UTF8) {
        out->type = KQ_BOOL;
        uint32_t m = cmp_mask(op);
        const kq_expr *a = e->l, *b = e->r;
        if ((a->kind == KQ_EX_LIT && a->is_null) || (b->kind == KQ_EX_LIT && b->is_null)) { out->v = "0u"; out->ok = "0u"; return KQ_OK; }
        if (a->kind == KQ_EX_LIT && b->kind == KQ_EX_LIT) {
            int c = a->s.compare(b->s); int code = c < 0 ? 0 : (c == 0 ? 1 : 2);
            int idx; KQ_RET(add_lit(((m >> code) & 1u) ? 0xFFFFFFFFULL : 0ULL, &idx));
            out->v = "((uint32_t)q.lit[" + S(idx) + "])"; out->ok = "";
            return KQ_OK;
        }
        if (a->kind == KQ_EX_LIT) { std::swap(a, b); m = mirror_mask(m); }
        int ca = bare_column(a);
        if (ca < 0) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "Utf8 comparison operands must be columns or literals");
        int sa; KQ_RET(use_col(ca, &sa));
        const std::string t = tmp();
        if (b->kind == KQ_EX_LIT) {
            int li; KQ_RET(add_utf8_lit(b->s, &li));
            const std::string ok = col_valid(sa), okx = ok.empty() ? "rc.inr" : ok;
            line("const uint32_t " + t + " = utf8_cmp_lit<SO" + S(sa) + ">(q.cols[" + S(sa) + "], q.pool + (uint32_t)(q.lit[" + S(li) + "] >> 32), (int)(uint32_t)q.lit[" +
                 S(li) + "], " + S(m) + "u, " + okx + ", rc);");
            out->v = t; out->ok = ok;
            return KQ_OK;
        }
        int cb = bare_column(b);
        if (cb < 0) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "Utf8 comparison operands must be columns or literals");
        int sb; KQ_RET(use_col(cb, &sb));
        KqVal va, vb; va.ok = col_valid(sa); vb.ok = col_valid(sb);
        const std::string ok = and_ok(va, vb), okx = ok.empty() ? "rc.inr" : ok;
        line("const uint32_t " + t + " = utf8_cmp_col<SO" + S(sa) + ", SO" + S(sb) + ">(q.cols[" + S(sa) + "], q.cols[" + S(sb) + "], " + S(m) + "u, " + okx + ", rc);");
        out->v = t; out->ok = ok;
        return KQ_OK;
    }
    if (is_cmp && lt == KQ_I32) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "compare on I32");
    if (!is_cmp && !is_logic && lt != KQ_I64 && lt != KQ_F64) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "math needs Int64 or Float64 operands");

    KqVal a, b;
    KQ_RET(value(e->l, &a));
    KQ_RET(value(e->r, &b));
    const std::string t = tmp();

    if (lt == KQ_BOOL) {
        out->type = KQ_BOOL;
        if (is_logic) {
            if (!a.nullable() && !b.nullable()) {
                line("const uint32_t " + t + " = " + a.v + (op == KQ_AND ? " & " : " | ") + b.v + ";");
                out->v = t; out->ok = "";
                return KQ_OK;
            }
            // SQL three-valued logic (rule E3) on truth masks
            const std::string oa = a.okx(), ob = b.okx(), k = tmp("n");
            if (op == KQ_AND) {
                line("const uint32_t " + t + " = (" + oa + " & " + a.v + ") & (" + ob + " & " + b.v + ");");
                line("const uint32_t " + k + " = (" + t + " | (" + oa + " & ~" + a.v + ") | (" + ob + " & ~" + b.v + ")) & rc.inr;");
            } else {
                line("const uint32_t " + t + " = (" + oa + " & " + a.v + ") | (" + ob + " & " + b.v + ");");
                line("const uint32_t " + k + " = (" + t + " | ((" + oa + " & ~" + a.v + ") & (" + ob + " & ~" + b.v + "))) & rc.inr;");
            }
            out->v = t; out->ok = k;
            return KQ_OK;
        }
        // Bool comparison: false < true
        const uint32_t m = cmp_mask(op);
        std::string ex;
        if (m & 1u) ex += "(~" + a.v + " & " + b.v + ")";
        if (m & 2u) ex += std::string(ex.empty() ? "" : " | ") + "~(" + a.v + " ^ " + b.v + ")";
        if (m & 4u) ex += std::string(ex.empty() ? "" : " | ") + "(" + a.v + " & ~" + b.v + ")";
        line("const uint32_t " + t + " = (" + ex + ") & RMASK;");
        out->v = t; out->ok = and_ok(a, b);
        return KQ_OK;
    }

    const bool f64 = lt == KQ_F64;
    const std::string xa = f64 ? "as_f64(" + a.at() + ")" : "(long long)" + a.at();
    const std::string xb = f64 ? "as_f64(" + b.at() + ")" : "(long long)" + b.at();
    out->ok = and_ok(a, b);
    if (is_cmp) {
        out->type = KQ_BOOL;
        line("uint32_t " + t + " = 0;");
        line(std::string(UNROLL) + t + " |= (uint32_t)(" + xa + " " + cmp_sym(op) + " " + xb + ") << r;");
        out->v = t;
        return KQ_OK;
    }
    out->type = lt;
    // result array (operands that are scalars are broadcast by at())
    if (op == KQ_DIV && !f64) {
        const KqVal aa = as_array(a), bb = as_array(b);
        line("uint64_t " + t + "[R]; div_i64(" + aa.v + ", " + bb.v + ", " + (out->ok.empty() ? std::string("rc.inr") : out->ok) + ", rc, " + t + ");");
        out->v = t;
        return KQ_OK;
    }
    std::string ex;
    if (f64) {
        // separately rounded IEEE operations, never contracted into FMA (rule E4)
        const char* fn = op == KQ_ADD ? "__dadd_rn" : (op == KQ_SUB ? "__dsub_rn" : (op == KQ_MUL ? "__dmul_rn" : "__ddiv_rn"));
        ex = "as_u64(" + std::string(fn) + "(" + xa + ", " + xb + "))";
    } else {
        const char* sym = op == KQ_ADD ? "+" : (op == KQ_SUB ? "-" : "*");     // two's-complement wrap like Kotlin Long (rule E1)
        ex = a.at() + " " + sym + " " + b.at();
    }
    line("uint64_t " + t + "[R];");
    line(std::string(UNROLL) + t + "[r] = " + ex + ";");
    out->v = t;
    return KQ_OK;
}

KqVal KqCodegen::as_array(const KqVal& x) {
    if (x.type == KQ_BOOL) {
        KqVal y = x; y.v = tmp(); y.scalar = false;
        line("uint64_t " + y.v + "[R];");
        line(std::string(UNROLL) + y.v + "[r] = (" + x.v + " >> r) & 1u;");
        return y;
    }
    if (!x.scalar) return x;
    KqVal y = x; y.v = tmp(); y.scalar = false;
    line("uint64_t " + y.v + "[R];");
    line(std::string(UNROLL) + y.v + "[r] = " + x.v + ";");
    return y;
}

int KqCodegen::value(const kq_expr* e, KqVal* out) {
    if (!e) return kq_fail(ctx, KQ_ERR_ILLEGAL_ARGUMENT, "null expression");
    switch (e->kind) {
        case KQ_EX_COL: return col_value(e, out);
        case KQ_EX_LIT: return lit_value(e, out);
        case KQ_EX_CAST: {
            if (e->type != KQ_F64) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "Cast to type %d is not supported", e->type);   // Main.kt:799
            int st; bool sn; KQ_RET(infer(e->l, &st, &sn));
            if (st == KQ_UTF8) {
                int bc = bare_column(e->l);
                if (bc < 0) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "Utf8->Float64 cast needs a column operand");
                int slot; KQ_RET(use_col(bc, &slot));
                const std::string ok = col_valid(slot), t = tmp();
                line("uint64_t " + t + "[R]; utf8_to_f64<SO" + S(slot) + ">(q.cols[" + S(slot) + "], " + (ok.empty() ? std::string("rc.inr") : ok) + ", rc, " + t + ");");
                out->v = t; out->ok = ok; out->type = KQ_F64; out->scalar = false;
                return KQ_OK;
            }
            if (st == KQ_I64) {
                KqVal x; KQ_RET(value(e->l, &x));
                const std::string t = tmp();
                if (x.scalar) line("const uint64_t " + t + " = as_u64((double)(long long)" + x.v + ");");
                else { line("uint64_t " + t + "[R];"); line(std::string(UNROLL) + t + "[r] = as_u64((double)(long long)" + x.at() + ");"); }
                *out = x; out->v = t; out->type = KQ_F64;
                return KQ_OK;
            }
            if (st == KQ_F64) return value(e->l, out);
            return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "Cannot cast value to Double");                                          // Main.kt:792
        }
        case KQ_EX_BIN: return bin_value(e, out);
    }
    return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "Unknown expr");
}

int KqCodegen::validity_only(const kq_expr* e, KqVal* out) {
    int bc = bare_column(e);
    if (bc >= 0) {
        int t; bool n; KQ_RET(infer(e, &t, &n));
        int slot; KQ_RET(use_col(bc, &slot));
        out->v = "1ULL"; out->scalar = true; out->type = KQ_I64;
        out->ok = col_valid(slot);
        return KQ_OK;
    }
    return value(e, out);
}

int KqCodegen::key_value(const kq_expr* e, KqVal* out) {
    int t; bool nl;
    KQ_RET(infer(e, &t, &nl));
    if (t == KQ_UTF8) {
        int bc = bare_column(e);
        if (bc < 0) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "Utf8 group keys must be columns");
        int slot; KQ_RET(use_col(bc, &slot));
        col_bytes_used[slot] = true;
        const std::string ok = col_valid(slot), tt = tmp();
        line("uint64_t " + tt + "[R]; utf8_pack<SO" + S(slot) + ", SB" + S(slot) + ", " + S(slot) + ">(q.cols[" + S(slot) + "], " + (ok.empty() ? std::string("rc.inr") : ok) + ", rc, " + tt + ");");
        out->v = tt; out->ok = ok; out->type = KQ_UTF8; out->scalar = false;
        return KQ_OK;
    }
    return value(e, out);
}

std::string KqCodegen::plan_stages(int budget, int min_stages, int tile_rows, StagePlan* sp, bool stage_bytes) {
    const int TILE = tile_rows;
    memset(sp, 0, sizeof *sp);
    int off = 0;
    int cur_slot = 0;
    int bytes_per_row = 4;           // stage room for Utf8 string bytes: 4 per row unless the column's average length says less
    auto add = [&](int kind, const void* g, int role) -> int {
        int bytes = kind == SK_W8 ? TILE * 8 : (kind == SK_W4 ? TILE * 4 : (kind == SK_W4_PLUS1 ? TILE * 4 + 16 : (kind == SK_BYTES ? TILE * bytes_per_row + 64 : TILE / 8)));
        bytes = (bytes + 127) / 128 * 128;
        if (sp->nbuf >= MAX_STAGE_BUFS || (off + bytes) * min_stages > budget) return -1;   // stays on the direct global path
        if (kind == SK_BIT && TILE % 128 != 0) return -1;                                   // a tile's bit range must start 16-byte aligned for the bulk copy
        StageBuf& sb = sp->buf[sp->nbuf];
        sb.g = (const char*)g; sb.soff = off; sb.kind = kind; sb.slot = cur_slot; sb.role = role;
        sp->nbuf++;
        const int at = off;
        off += bytes;
        return at;
    };
    std::string defs;
    for (int i = 0; i < ncols; i++) {
        kq_col* c = batch->cols[(size_t)slot_col[i]];
        int sd = -1, sv = -1, so = -1, sb = -1;
        cur_slot = i;
        switch (c->type) {
            case KQ_F64: case KQ_I64: sd = add(SK_W8, c->data, 0); break;
            case KQ_DATE32: case KQ_I32: sd = add(SK_W4, c->data, 0); break;
            case KQ_BOOL: sd = add(SK_BIT, c->data, 0); break;
            case KQ_UTF8:
                so = add(SK_W4_PLUS1, c->offsets, 2);
                // string bytes: a second-phase copy of the range the tile's offsets span (up to 4 bytes/row on average; longer tiles fall back to global loads)
                if (stage_bytes && so >= 0 && col_bytes_used[i] && sp->nbytes < MAX_BYTES_BUFS) {
                    const int obuf = sp->nbuf - 1;
                    // the column's average string length (known for uploaded and generated columns) + 25 %, at most 4 bytes per
                    // row; a tile whose strings need more falls back to global loads for that tile, nothing else changes
                    bytes_per_row = 4;
                    if (c->n > 0 && c->data_bytes >= 0) {
                        const long long avg125 = (c->data_bytes * 5 + c->n * 4 - 1) / (c->n * 4);
                        bytes_per_row = (int)std::max<long long>(2, std::min<long long>(4, avg125 + 1));
                    }
                    sb = add(SK_BYTES, c->data, 0);
                    if (sb >= 0) { sp->buf[sp->nbuf - 1].aux = obuf | (i << 16); sp->buf[sp->nbuf - 1].cap = TILE * bytes_per_row + 64; sp->bytes_buf[sp->nbytes++] = sp->nbuf - 1; }
                }
                break;
        }
        if (c->validity) sv = add(SK_BIT, c->validity, 1);
        defs += "constexpr int SD" + S(i) + " = " + S(sd) + ", SV" + S(i) + " = " + S(sv) + ", SO" + S(i) + " = " + S(so) + ", SB" + S(i) + " = " + S(sb) + ";\n";
    }
    sp->stage_bytes = off > 0 ? off : 128;
    int ns = budget / sp->stage_bytes;
    sp->nstages = ns > MAX_STAGES ? MAX_STAGES : (ns < 1 ? 1 : ns);
    return defs;
}
