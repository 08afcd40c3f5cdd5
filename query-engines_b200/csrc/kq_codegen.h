// kq_codegen.h — host-side query compiler: kq_expr trees + the batch schema -> CUDA source of the
// `struct Q` a kernel skeleton (kq_k_ops.cuh / kq_k_agg.cuh) is specialised with, plus the run-time
// operand block (column pointers, literal values) that goes with it. See kq_jit.cu for how the
// source becomes a kernel.
#pragma once

#include <cstring>
#include <string>
#include <vector>

#include "kq_args.h"
#include "kq_internal.h"

// A value produced by generated code inside the current function body.
struct KqVal {
    std::string v;        // name of a uint64_t[R] array, of a uint64_t scalar (scalar == true) or of a uint32_t truth mask (Bool)
    std::string ok;       // name/expression of the R-bit validity mask; "" = every in-range row valid; "0u" = always null
    int type = 0;         // kq_type
    bool scalar = false;  // literal: one value for all rows
    std::string at(const char* r = "r") const { return scalar ? v : v + "[" + r + "]"; }
    std::string okx() const { return ok.empty() ? std::string("rc.inr") : ok; }
    bool nullable() const { return !ok.empty(); }
};

struct KqCodegen {
    kq_ctx* ctx = nullptr;
    kq_batch* batch = nullptr;
    kq::QArgs args;                  // run-time operands (column pointers filled by use_col, literals by add_lit)
    int ncols = 0, nlit = 0, pool_used = 0;
    int colmap[256];
    int slot_col[kq::MAX_COLS];      // program column slot -> batch column index

    int begin(kq_ctx* c, kq_batch* b);
    // Static type of an expression against this batch (no code emitted). Returns a kq_status.
    int infer(const kq_expr* e, int* type, bool* nullable);
    // Start the body of a new generated function: forgets which columns are already in registers.
    void begin_body();
    std::string take_body();
    // Emit code computing e into the current body.
    int value(const kq_expr* e, KqVal* out);
    // COUNT(expr) only needs validity: cheaper code for bare columns of any type (incl. Utf8).
    int validity_only(const kq_expr* e, KqVal* out);
    // Group-key code: 64-bit key word (Utf8 columns are packed; Bool becomes 0/1).
    int key_value(const kq_expr* e, KqVal* out);
    // Materialise a value as a uint64_t[R] array (broadcast scalars, expand Bool truth masks to 0/1).
    KqVal as_array(const KqVal& x);
    void line(const std::string& s) { body += "        " + s + "\n"; }
    std::string tmp(const char* prefix = "t");

    // Decide which column buffers are staged through the shared-memory tile pipeline (kq_pipe.cuh)
    // and return the `constexpr int SD<i>, SV<i>, SO<i>` definitions the generated code refers to.
    std::string plan_stages(int budget, int min_stages, int tile_rows, kq::StagePlan* sp, bool stage_bytes = false);
    bool col_bytes_used[kq::MAX_COLS];   // Utf8 column slots whose string bytes the generated code can read from the stage

    static int bare_column(const kq_expr* e) { return e && e->kind == KQ_EX_COL ? e->col : -1; }

   private:
    std::string body;
    int ntmp = 0;
    bool col_loaded[kq::MAX_COLS], valid_loaded[kq::MAX_COLS];
    int use_col(int batch_col, int* slot);
    int add_lit(uint64_t v, int* idx);
    int add_utf8_lit(const std::string& s, int* idx);
    std::string col_valid(int slot);                  // "" when the column has no validity buffer
    int col_value(const kq_expr* e, KqVal* out);
    int lit_value(const kq_expr* e, KqVal* out);
    int bin_value(const kq_expr* e, KqVal* out);
};

// A batch that exists as a schema only (types + nullability), for kq_explain_*.
struct KqSchemaBatch {
    kq_batch batch;
    std::vector<kq_col> cols;
    KqSchemaBatch(int ncols, const int* types, const int* nullable) : cols((size_t)ncols) {
        for (int i = 0; i < ncols; i++) {
            cols[(size_t)i].type = types[i];
            cols[(size_t)i].validity = (nullable && nullable[i]) ? reinterpret_cast<uint32_t*>(0x100) : nullptr;   // never dereferenced
            batch.cols.push_back(&cols[(size_t)i]);
        }
    }
};
inline void kq_copy_text(const std::string& s, char* dst, size_t cap) {
    if (!dst || cap == 0) return;
    size_t m = s.size() < cap - 1 ? s.size() : cap - 1;
    memcpy(dst, s.data(), m);
    dst[m] = 0;
}

// ---- kq_jit.cu ----------------------------------------------------------------------------------------------------
// Compile (or fetch from the process-wide cache) the kernel `entry` of: prelude + `generated` + skeleton.
// `defines` are `#define` lines placed before the prelude (KQ_R, KQ_WARPS, KQ_KERNEL_*).
enum KqSkeleton { KQ_SKEL_OPS = 0, KQ_SKEL_AGG = 1, KQ_SKEL_AGG_FE = 2 };
KQ_HIDDEN int kq_jit_kernel(kq_ctx* ctx, const std::string& defines, const std::string& generated, int skeleton,
                            const char* entry, int dynamic_smem, void** kernel_out);
KQ_HIDDEN int kq_jit_compile_only(kq_ctx* ctx, const std::string& defines, const std::string& generated, int skeleton);
