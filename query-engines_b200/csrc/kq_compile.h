// kq_compile.h — host-side compiler: kq_expr trees -> one postfix kq::Program per kernel launch.
#pragma once

#include "kq_internal.h"
#include "kq_vm.cuh"
#include "kq_pipe.cuh"

struct KqCompiler {
    kq_ctx* ctx = nullptr;
    kq_batch* batch = nullptr;
    kq::Program prog;
    int sp = 0;
    int nlit = 0;
    int pool_used = 0;
    int colmap[256];
    int slot_col[kq::MAX_COLS];     // program column slot -> batch column index

    int begin(kq_ctx* c, kq_batch* b);
    // Static type of an expression against this batch (no code emitted). Returns a kq_status.
    int infer(const kq_expr* e, int* type, bool* nullable);
    // Emit code that leaves the value of e in the accumulator.
    int value(const kq_expr* e, int* type, bool* nullable);
    // COUNT(expr) only needs validity: cheaper code for bare columns of any type (incl. Utf8).
    int validity_only(const kq_expr* e);
    // Group-key code: 64-bit key word (Utf8 columns are packed, F64 NaNs are left as-is and
    // canonicalised by the aggregate kernel).
    int key_value(const kq_expr* e, int* type, bool* nullable);
    // Hand the accumulator to a sink (O_SET_SEL / O_EMIT / O_SET_KEY / O_SET_IN).
    int sink(int op, int arg, int type = 0);   // type KQ_BOOL: ACC holds a truth mask
    int pc() const { return prog.ninsn; }
    // Decide which column buffers are staged through the shared-memory tile pipeline (kq_pipe.cuh):
    // fills prog.cols[].s_* and the plan. `budget` = bytes of shared memory available for stages,
    // `min_stages` = ring depth that must fit.
    void plan_stages(int budget, int min_stages, int tile_rows, kq::StagePlan* sp);
    // index of the batch column if e is a bare ColumnExpression, else -1
    static int bare_column(const kq_expr* e) { return e && e->kind == KQ_EX_COL ? e->col : -1; }

   private:
    int emit(int op, int src, int a, uint32_t b = 0);
    bool leaf_src(const kq_expr* e, int* src, int* a);
    bool is_plain_leaf(const kq_expr* e);
    bool is_leaf64(const kq_expr* e);
    int use_col(int batch_col, int* slot);
    int add_lit(uint64_t v, int* idx);
    int add_utf8_lit(const std::string& s, int* idx);
};
