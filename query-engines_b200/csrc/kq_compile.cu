// kq_compile.cu — expression handles (kq_expr_*) and the tree -> postfix-program compiler.
//
// Type rules follow the oracle's BinaryExpression/CastExpression (rules E2-E4, R5): operand types
// must match, AND/OR need Bool, math needs Int64/Float64, the only cast target is Float64. Type
// errors are IllegalStateException (KQ_ERR_ILLEGAL_STATE), exactly where the reference would throw
// (Main.kt:792, 799).
#include <cstring>

#include "kq_compile.h"

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

int KqCompiler::begin(kq_ctx* c, kq_batch* b) {
    ctx = c; batch = b; sp = 0; nlit = 0; pool_used = 0;
    memset(&prog, 0, sizeof prog);
    for (int& x : colmap) x = -1;
    if (b->cols.size() > 256) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "batches wider than 256 columns are not supported");
    return KQ_OK;
}

int KqCompiler::emit(int op, int arg, int delta) {
    if (prog.ninsn >= MAX_INSN) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "expression program too long (> %d instructions)", MAX_INSN);
    if (sp + (delta > 0 ? delta : 0) > D) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "expression too deep (evaluation stack > %d)", D);
    Insn& in = prog.insn[prog.ninsn++];
    in.op = (uint8_t)op; in.sp = (uint8_t)sp; in.arg = (uint16_t)arg;
    sp += delta;
    return KQ_OK;
}

int KqCompiler::use_col(int bc, int* slot) {
    if (bc < 0 || bc >= (int)batch->cols.size()) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "column index %d out of range (batch has %d fields)", bc, (int)batch->cols.size());
    if (colmap[bc] < 0) {
        if (prog.ncols >= MAX_COLS) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "more than %d distinct columns in one kernel", MAX_COLS);
        kq_col* c = batch->cols[(size_t)bc];
        DCol& d = prog.cols[prog.ncols];
        d.data = c->data; d.validity = c->validity; d.offsets = c->offsets;
        colmap[bc] = prog.ncols++;
    }
    *slot = colmap[bc];
    return KQ_OK;
}

int KqCompiler::add_lit(uint64_t v, int* idx) {
    for (int i = 0; i < nlit; i++) if (prog.lit[i] == v) { *idx = i; return KQ_OK; }
    if (nlit >= MAX_LIT) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "more than %d literals in one kernel", MAX_LIT);
    prog.lit[nlit] = v; *idx = nlit++;
    return KQ_OK;
}
int KqCompiler::add_utf8_lit(const std::string& s, int* idx) {
    if (pool_used + (int)s.size() > LITPOOL) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "Utf8 literals exceed %d bytes", LITPOOL);
    memcpy(prog.pool + pool_used, s.data(), s.size());
    uint64_t v = ((uint64_t)pool_used << 32) | (uint64_t)s.size();
    pool_used += (int)s.size();
    if (nlit >= MAX_LIT) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "more than %d literals in one kernel", MAX_LIT);
    prog.lit[nlit] = v; *idx = nlit++;
    return KQ_OK;
}

int KqCompiler::infer(const kq_expr* e, int* type, bool* nullable) {
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

int KqCompiler::value(const kq_expr* e, int* type, bool* nullable) {
    if (!e) return kq_fail(ctx, KQ_ERR_ILLEGAL_ARGUMENT, "null expression");
    switch (e->kind) {
        case KQ_EX_COL: {
            KQ_RET(infer(e, type, nullable));
            int slot; KQ_RET(use_col(e->col, &slot));
            switch (*type) {
                case KQ_F64: case KQ_I64: return emit(OP_PUSH_COL64, slot, +1);
                case KQ_DATE32: case KQ_I32: return emit(OP_PUSH_COL32, slot, +1);
                case KQ_BOOL: return emit(OP_PUSH_COLBIT, slot, +1);
                default: return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "a Utf8 column cannot be an operand here (only comparisons, COUNT, group keys and pass-through)");
            }
        }
        case KQ_EX_LIT: {
            *type = e->type; *nullable = e->is_null;
            if (e->is_null) return emit(OP_PUSH_NULL, 0, +1);
            uint64_t bits;
            switch (e->type) {
                case KQ_F64: memcpy(&bits, &e->f, 8); break;
                case KQ_I64: case KQ_DATE32: case KQ_BOOL: case KQ_I32: bits = (uint64_t)e->i; break;
                default: return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "a Utf8 literal can only be compared with a Utf8 column");
            }
            int idx; KQ_RET(add_lit(bits, &idx));
            return emit(OP_PUSH_LIT, idx, +1);
        }
        case KQ_EX_CAST: {
            if (e->type != KQ_F64) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "Cast to type %d is not supported", e->type);   // Main.kt:799
            int st; bool sn; KQ_RET(infer(e->l, &st, &sn));
            *type = KQ_F64; *nullable = sn;
            if (st == KQ_UTF8) {
                int bc = bare_column(e->l);
                if (bc < 0) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "Utf8->Float64 cast needs a column operand");
                int slot; KQ_RET(use_col(bc, &slot));
                return emit(OP_UTF8_TO_F64, slot, +1);
            }
            if (st == KQ_I64) { int t; bool n; KQ_RET(value(e->l, &t, &n)); return emit(OP_I64_TO_F64, 0, 0); }
            if (st == KQ_F64) { int t; bool n; return value(e->l, &t, &n); }
            return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "Cannot cast value to Double");                                          // Main.kt:792
        }
        case KQ_EX_BIN: {
            int lt, rt; bool ln, rn;
            KQ_RET(infer(e->l, &lt, &ln)); KQ_RET(infer(e->r, &rt, &rn));
            if (lt != rt) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "binary operand types differ (%d vs %d)", lt, rt);      // rule E2
            *nullable = ln || rn;
            int op = e->op;
            bool is_cmp = op >= KQ_EQ && op <= KQ_GE, is_logic = op == KQ_AND || op == KQ_OR;
            if (!is_cmp && !is_logic && !(op >= KQ_ADD && op <= KQ_DIV)) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "unknown binary operator %d", op);
            if (is_logic) {
                if (lt != KQ_BOOL) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "AND/OR need Bool operands");
                int t; bool n;
                KQ_RET(value(e->l, &t, &n)); KQ_RET(value(e->r, &t, &n));
                *type = KQ_BOOL;
                return emit(op == KQ_AND ? OP_AND : OP_OR, 0, -1);
            }
            if (is_cmp) {
                *type = KQ_BOOL;
                uint32_t m = cmp_mask(op);
                if (lt == KQ_UTF8) {
                    const kq_expr *a = e->l, *b = e->r;
                    if ((a->kind == KQ_EX_LIT && a->is_null) || (b->kind == KQ_EX_LIT && b->is_null)) return emit(OP_PUSH_NULL, 0, +1);
                    if (a->kind == KQ_EX_LIT && b->kind == KQ_EX_LIT) {
                        int c = a->s.compare(b->s); int code = c < 0 ? 0 : (c == 0 ? 1 : 2);
                        int idx; KQ_RET(add_lit((m >> code) & 1u, &idx));
                        return emit(OP_PUSH_LIT, idx, +1);
                    }
                    if (a->kind == KQ_EX_LIT) { std::swap(a, b); m = mirror_mask(m); }
                    int ca = bare_column(a);
                    if (ca < 0) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "Utf8 comparison operands must be columns or literals");
                    int sa; KQ_RET(use_col(ca, &sa));
                    if (b->kind == KQ_EX_LIT) {
                        int li; KQ_RET(add_utf8_lit(b->s, &li));
                        return emit(OP_UTF8_CMP_LIT, sa | (li << 5) | (int)(m << 10), +1);
                    }
                    int cb = bare_column(b);
                    if (cb < 0) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "Utf8 comparison operands must be columns or literals");
                    int sb; KQ_RET(use_col(cb, &sb));
                    return emit(OP_UTF8_CMP_COL, sa | (sb << 5) | (int)(m << 10), +1);
                }
                if (lt == KQ_I32) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "compare on I32");
                int t; bool n;
                KQ_RET(value(e->l, &t, &n)); KQ_RET(value(e->r, &t, &n));
                return emit(lt == KQ_F64 ? OP_CMP_F64 : OP_CMP_I64, (int)m, -1);
            }
            if (lt != KQ_I64 && lt != KQ_F64) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "math needs Int64 or Float64 operands");
            int t; bool n;
            KQ_RET(value(e->l, &t, &n)); KQ_RET(value(e->r, &t, &n));
            *type = lt;
            int base = lt == KQ_F64 ? OP_ADD_F64 : OP_ADD_I64;
            return emit(base + (op - KQ_ADD), 0, -1);
        }
    }
    return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "Unknown expr");
}

int KqCompiler::validity_only(const kq_expr* e) {
    int bc = bare_column(e);
    if (bc >= 0) {
        int t; bool n; KQ_RET(infer(e, &t, &n));
        int slot; KQ_RET(use_col(bc, &slot));
        return emit(OP_PUSH_VALID, slot, +1);
    }
    int t; bool n;
    return value(e, &t, &n);
}

int KqCompiler::key_value(const kq_expr* e, int* type, bool* nullable) {
    KQ_RET(infer(e, type, nullable));
    if (*type == KQ_UTF8) {
        int bc = bare_column(e);
        if (bc < 0) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "Utf8 group keys must be columns");
        int slot; KQ_RET(use_col(bc, &slot));
        return emit(OP_UTF8_PACK, slot, +1);
    }
    return value(e, type, nullable);
}

int KqCompiler::sink(int op, int arg) { return emit(op, arg, -1); }
