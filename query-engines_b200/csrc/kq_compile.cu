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

int KqCompiler::emit(int op, int src, int a, uint32_t b) {
    if (prog.ninsn >= MAX_INSN) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "expression program too long (> %d instructions)", MAX_INSN);
    Insn& in = prog.insn[prog.ninsn++];
    in.op = (uint8_t)op; in.src = (uint8_t)src; in.a = (uint16_t)a; in.b = b;
    return KQ_OK;
}

int KqCompiler::use_col(int bc, int* slot) {
    if (bc < 0 || bc >= (int)batch->cols.size()) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "column index %d out of range (batch has %d fields)", bc, (int)batch->cols.size());
    if (colmap[bc] < 0) {
        if (prog.ncols >= MAX_COLS) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "more than %d distinct columns in one kernel", MAX_COLS);
        kq_col* c = batch->cols[(size_t)bc];
        DCol& d = prog.cols[prog.ncols];
        d.data = c->data; d.validity = c->validity; d.offsets = c->offsets;
        d.s_data = d.s_valid = d.s_off = -1;
        slot_col[prog.ncols] = bc;
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

// A leaf can be fetched straight into TMP by the instruction that consumes it.
bool KqCompiler::leaf_src(const kq_expr* e, int* src, int* a) {
    if (e->kind == KQ_EX_COL) {
        if (e->col < 0 || e->col >= (int)batch->cols.size()) return false;
        int t = batch->cols[(size_t)e->col]->type;
        int slot;
        if (t == KQ_UTF8 || use_col(e->col, &slot) != KQ_OK) return false;
        *src = (t == KQ_F64 || t == KQ_I64) ? S_COL64 : (t == KQ_BOOL ? S_COLBIT : S_COL32);
        *a = slot;
        return true;
    }
    if (e->kind == KQ_EX_LIT && e->type != KQ_UTF8) {
        if (e->is_null) { *src = S_NULL; *a = 0; return true; }
        uint64_t bits;
        if (e->type == KQ_F64) memcpy(&bits, &e->f, 8);
        else if (e->type == KQ_BOOL) bits = e->i ? 0xFFFFFFFFULL : 0ULL;     // Bool values are truth masks (kq_vm.cuh)
        else bits = (uint64_t)e->i;
        int idx;
        if (add_lit(bits, &idx) != KQ_OK) return false;
        *src = S_LIT; *a = idx;
        return true;
    }
    return false;
}

// Emit code that leaves the value of e in ACC.
int KqCompiler::value(const kq_expr* e, int* type, bool* nullable) {
    if (!e) return kq_fail(ctx, KQ_ERR_ILLEGAL_ARGUMENT, "null expression");
    switch (e->kind) {
        case KQ_EX_COL: {
            KQ_RET(infer(e, type, nullable));
            if (*type == KQ_UTF8) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "a Utf8 column cannot be an operand here (only comparisons, COUNT, group keys and pass-through)");
            int src, a;
            if (!leaf_src(e, &src, &a)) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "more than %d distinct columns in one kernel", MAX_COLS);
            return emit(O_LOAD, src, a);
        }
        case KQ_EX_LIT: {
            *type = e->type; *nullable = e->is_null;
            if (e->type == KQ_UTF8 && !e->is_null) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "a Utf8 literal can only be compared with a Utf8 column");
            if (e->is_null) return emit(O_LOAD, S_NULL, 0);
            int src, a;
            if (!leaf_src(e, &src, &a)) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "more than %d literals in one kernel", MAX_LIT);
            return emit(O_LOAD, src, a);
        }
        case KQ_EX_CAST: {
            if (e->type != KQ_F64) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "Cast to type %d is not supported", e->type);   // Main.kt:799
            int st; bool sn; KQ_RET(infer(e->l, &st, &sn));
            *type = KQ_F64; *nullable = sn;
            if (st == KQ_UTF8) {
                int bc = bare_column(e->l);
                if (bc < 0) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "Utf8->Float64 cast needs a column operand");
                int slot; KQ_RET(use_col(bc, &slot));
                return emit(O_LOAD, S_UTF8_F64, slot);
            }
            if (st == KQ_I64) { int t; bool n; KQ_RET(value(e->l, &t, &n)); return emit(O_I64_TO_F64, S_NONE, 0); }
            if (st == KQ_F64) { int t; bool n; return value(e->l, &t, &n); }
            return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "Cannot cast value to Double");                                          // Main.kt:792
        }
        case KQ_EX_BIN: {
            int lt, rt; bool ln, rn;
            KQ_RET(infer(e->l, &lt, &ln)); KQ_RET(infer(e->r, &rt, &rn));
            if (lt != rt) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "binary operand types differ (%d vs %d)", lt, rt);      // rule E2
            *nullable = ln || rn;
            const int op = e->op;
            const bool is_cmp = op >= KQ_EQ && op <= KQ_GE, is_logic = op == KQ_AND || op == KQ_OR;
            if (!is_cmp && !is_logic && !(op >= KQ_ADD && op <= KQ_DIV)) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "unknown binary operator %d", op);
            if (is_logic && lt != KQ_BOOL) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "AND/OR need Bool operands");
            if (is_cmp && lt == KQ_UTF8) {
                *type = KQ_BOOL;
                uint32_t m = cmp_mask(op);
                const kq_expr *a = e->l, *b = e->r;
                if ((a->kind == KQ_EX_LIT && a->is_null) || (b->kind == KQ_EX_LIT && b->is_null)) return emit(O_LOAD, S_NULL, 0);
                if (a->kind == KQ_EX_LIT && b->kind == KQ_EX_LIT) {
                    int c = a->s.compare(b->s); int code = c < 0 ? 0 : (c == 0 ? 1 : 2);
                    int idx; KQ_RET(add_lit(((m >> code) & 1u) ? 0xFFFFFFFFULL : 0ULL, &idx));
                    return emit(O_LOAD, S_LIT, idx);
                }
                if (a->kind == KQ_EX_LIT) { std::swap(a, b); m = mirror_mask(m); }
                int ca = bare_column(a);
                if (ca < 0) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "Utf8 comparison operands must be columns or literals");
                int sa; KQ_RET(use_col(ca, &sa));
                if (b->kind == KQ_EX_LIT) {
                    int li; KQ_RET(add_utf8_lit(b->s, &li));
                    return emit(O_LOAD, S_UTF8_CMP_LIT, sa, (uint32_t)li | (m << 8));
                }
                int cb = bare_column(b);
                if (cb < 0) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "Utf8 comparison operands must be columns or literals");
                int sb; KQ_RET(use_col(cb, &sb));
                return emit(O_LOAD, S_UTF8_CMP_COL, sa, (uint32_t)sb | (m << 8));
            }
            if (is_cmp && lt == KQ_I32) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "compare on I32");
            if (!is_cmp && !is_logic && lt != KQ_I64 && lt != KQ_F64) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "math needs Int64 or Float64 operands");
            *type = (is_cmp || is_logic) ? KQ_BOOL : lt;
            int t; bool n;
            if (lt == KQ_BOOL) {
                // Bool operands are truth masks: ACC op operand, operand = Bool column / literal / save slot
                const int bop = is_logic ? (op == KQ_AND ? O_AND : O_OR) : O_CMP_BOOL;
                const uint32_t m = is_cmp ? cmp_mask(op) : 0, mr = is_cmp ? mirror_mask(m) : 0;
                int src, a;
                if (is_plain_leaf(e->r) && leaf_src(e->r, &src, &a)) { KQ_RET(value(e->l, &t, &n)); return emit(bop, src, a, m << 8); }
                if (is_plain_leaf(e->l) && leaf_src(e->l, &src, &a)) { KQ_RET(value(e->r, &t, &n)); return emit(bop, src, a, mr << 8); }
                if (sp >= DS) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "expression too deep (more than %d nested non-leaf operand pairs)", DS);
                KQ_RET(value(e->l, &t, &n));
                KQ_RET(emit(O_PUSH, S_NONE, sp));
                sp++;
                int st = value(e->r, &t, &n);
                sp--;
                KQ_RET(st);
                return emit(bop, S_STACK, sp, mr << 8);       // ACC holds the right operand
            }
            // 64-bit primitives, specialised per operand mode: ACC = X op Y in source order
            const bool f64 = lt == KQ_F64;
            int bin;
            if (is_cmp) bin = f64 ? B_CMP_F64 : B_CMP_I64;
            else bin = (f64 ? B_ADD_F64 : B_ADD_I64) + (op - KQ_ADD);
            const uint32_t mb = is_cmp ? (cmp_mask(op) << 8) : 0;
            auto prim = [&](int mode, int a, uint32_t b2) { return emit(O_BIN + bin, mode, a, b2 | mb); };
            int sl = 0, al = 0, sr = 0, ar = 0;
            const bool ll = is_leaf64(e->l) && leaf_src(e->l, &sl, &al);
            const bool rl = is_leaf64(e->r) && leaf_src(e->r, &sr, &ar);
            if (ll && rl) {
                if (sl == S_COL64 && sr == S_COL64) return prim(M_COL_COL, al, (uint32_t)ar);
                if (sl == S_COL64) return prim(M_COL_LIT, al, (uint32_t)ar);
                if (sr == S_COL64) return prim(M_LIT_COL, al, (uint32_t)ar);
                KQ_RET(emit(O_LOAD, S_LIT, al));
                return prim(M_ACC_LIT, ar, 0);
            }
            if (rl) { KQ_RET(value(e->l, &t, &n)); return prim(sr == S_COL64 ? M_ACC_COL : M_ACC_LIT, ar, 0); }
            if (ll) { KQ_RET(value(e->r, &t, &n)); return prim(sl == S_COL64 ? M_COL_ACC : M_LIT_ACC, al, 0); }
            if (sp >= DS) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "expression too deep (more than %d nested non-leaf operand pairs)", DS);
            KQ_RET(value(e->l, &t, &n));
            KQ_RET(emit(O_PUSH, S_NONE, sp));
            sp++;
            int st = value(e->r, &t, &n);
            sp--;
            KQ_RET(st);
            return prim(M_STK_ACC, sp, 0);
        }
    }
    return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "Unknown expr");
}

// a leaf the 64-bit primitives can read directly: a Float64/Int64 column or a non-null 64-bit literal
bool KqCompiler::is_leaf64(const kq_expr* e) {
    if (e->kind == KQ_EX_LIT) return !e->is_null && e->type != KQ_UTF8 && e->type != KQ_BOOL;
    if (e->kind != KQ_EX_COL || e->col < 0 || e->col >= (int)batch->cols.size()) return false;
    int t = batch->cols[(size_t)e->col]->type;
    return t == KQ_F64 || t == KQ_I64;
}

bool KqCompiler::is_plain_leaf(const kq_expr* e) {
    if (e->kind == KQ_EX_LIT) return e->type != KQ_UTF8;
    if (e->kind != KQ_EX_COL || e->col < 0 || e->col >= (int)batch->cols.size()) return false;
    return batch->cols[(size_t)e->col]->type != KQ_UTF8;
}

int KqCompiler::validity_only(const kq_expr* e) {
    int bc = bare_column(e);
    if (bc >= 0) {
        int t; bool n; KQ_RET(infer(e, &t, &n));
        int slot; KQ_RET(use_col(bc, &slot));
        return emit(O_LOAD, S_VALID, slot);
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
        return emit(O_LOAD, S_UTF8_PACK, slot);
    }
    return value(e, type, nullable);
}

// Hand ACC to a sink (O_SET_SEL / O_EMIT / O_SET_KEY / O_SET_IN); ACC stays valid.
int KqCompiler::sink(int op, int arg, int type) { return emit(op, S_NONE, arg, type == KQ_BOOL ? 1u : 0u); }

void KqCompiler::plan_stages(int budget, int min_stages, int tile_rows, StagePlan* sp) {
    const int TILE = tile_rows;
    memset(sp, 0, sizeof *sp);
    int off = 0;
    auto add = [&](int kind, const void* g, int32_t* soff) {
        int bytes = kind == SK_W8 ? TILE * 8 : (kind == SK_W4 ? TILE * 4 : (kind == SK_W4_PLUS1 ? TILE * 4 + 16 : TILE / 8));
        bytes = (bytes + 127) / 128 * 128;
        if (sp->nbuf >= MAX_STAGE_BUFS || (off + bytes) * min_stages > budget) return;   // stays on the direct global path
        sp->buf[sp->nbuf].g = (const char*)g; sp->buf[sp->nbuf].soff = off; sp->buf[sp->nbuf].kind = kind;
        sp->nbuf++;
        *soff = off;
        off += bytes;
    };
    for (int i = 0; i < prog.ncols; i++) {
        kq_col* c = batch->cols[(size_t)slot_col[i]];
        DCol& d = prog.cols[i];
        switch (c->type) {
            case KQ_F64: case KQ_I64: add(SK_W8, c->data, &d.s_data); break;
            case KQ_DATE32: case KQ_I32: add(SK_W4, c->data, &d.s_data); break;
            case KQ_BOOL: add(SK_BIT, c->data, &d.s_data); break;
            case KQ_UTF8: add(SK_W4_PLUS1, c->offsets, &d.s_off); break;   // string bytes stay in global memory
        }
        if (c->validity) add(SK_BIT, c->validity, &d.s_valid);
    }
    sp->stage_bytes = off > 0 ? off : 128;
    int ns = budget / sp->stage_bytes;
    sp->nstages = ns > MAX_STAGES ? MAX_STAGES : (ns < 1 ? 1 : ns);
}
