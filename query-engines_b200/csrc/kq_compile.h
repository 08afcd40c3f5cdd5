// kq_compile.h — host-side compiler: kq_expr trees -> one postfix kq::Program per kernel launch.
#pragma once

#include "kq_internal.h"
#include "kq_vm.cuh"

struct KqCompiler {
    kq_ctx* ctx = nullptr;
    kq_batch* batch = nullptr;
    kq::Program prog;
    int sp = 0;
    int nlit = 0;
    int pool_used = 0;
    int colmap[256];

    int begin(kq_ctx* c, kq_batch* b);
    // Static type of an expression against this batch (no code emitted). Returns a kq_status.
    int infer(const kq_expr* e, int* type, bool* nullable);
    // Emit code that leaves exactly one value on the stack.
    int value(const kq_expr* e, int* type, bool* nullable);
    // COUNT(expr) only needs validity: cheaper code for bare columns of any type (incl. Utf8).
    int validity_only(const kq_expr* e);
    // Group-key code: 64-bit key word (Utf8 columns are packed, F64 NaNs are left as-is and
    // canonicalised by the aggregate kernel).
    int key_value(const kq_expr* e, int* type, bool* nullable);
    // Pop the top of the stack into a sink (OP_SET_SEL / OP_EMIT / OP_SET_KEY / OP_SET_IN).
    int sink(int op, int arg);
    int pc() const { return prog.ninsn; }
    // index of the batch column if e is a bare ColumnExpression, else -1
    static int bare_column(const kq_expr* e) { return e && e->kind == KQ_EX_COL ? e->col : -1; }

   private:
    int emit(int op, int arg, int delta);
    int use_col(int batch_col, int* slot);
    int add_lit(uint64_t v, int* idx);
    int add_utf8_lit(const std::string& s, int* idx);
};
