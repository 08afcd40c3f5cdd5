// kq_ops.cu — host side of ProjectionExec, FilterExec and the fused filter+project operator, plus the
// gather kernels for pass-through columns. The streaming kernels themselves live in kq_k_ops.cuh and
// are specialised per query shape at run time (kq_codegen.cu -> kq_jit.cu).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kq_codegen.h"
#include "kq_scan.cuh"

using namespace kq;

namespace {

// Tile geometry of the streaming kernels (compiled into the specialised kernels as KQ_R / KQ_WARPS):
// 16 consumer warps x 32 lanes x 4 rows = 2048 rows per tile, plus the service warps (measured sweep: R x warps of
// 2x28, 4x12, 4x15, 4x16, 6x10, 8x8 all land within 10 % of each other; 4x16 was best).
constexpr int OPS_R = 4;
constexpr int OPS_WARPS = 16;
constexpr int TILE = OPS_WARPS * 32 * OPS_R;
constexpr int THREADS_PROJECT = OPS_WARPS * 32 + 32;     // + TMA producer warp
constexpr int THREADS_FILTER = OPS_WARPS * 32 + 96;      // + TMA producer warp + 2 look-back warps

// ---- gathers by selection vector (Utf8 pass-through columns and kq_filter) --------------------------------------
template <typename T>
__global__ void k_gather_fixed(const T* __restrict__ in, const uint32_t* __restrict__ in_valid, const int32_t* __restrict__ sel,
                               const unsigned long long* __restrict__ d_count, T* __restrict__ out, uint32_t* out_valid) {
    long long m = (long long)*d_count;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        int32_t row = sel[i];
        out[i] = in[row];
        if (out_valid && ((in_valid[row >> 5] >> (row & 31)) & 1u)) atomicOr(out_valid + (i >> 5), 1u << (i & 31));
    }
}
__global__ void k_gather_bits(const uint32_t* __restrict__ in, const uint32_t* __restrict__ in_valid, const int32_t* __restrict__ sel,
                              const unsigned long long* __restrict__ d_count, uint32_t* out, uint32_t* out_valid) {
    long long m = (long long)*d_count;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        int32_t row = sel[i];
        if ((in[row >> 5] >> (row & 31)) & 1u) atomicOr(out + (i >> 5), 1u << (i & 31));
        if (out_valid && ((in_valid[row >> 5] >> (row & 31)) & 1u)) atomicOr(out_valid + (i >> 5), 1u << (i & 31));
    }
}

// Utf8 gather: lengths of the selected rows -> exclusive prefix -> output offsets (kq_scan.cuh).
struct GatherLen {
    const int32_t* in_off; const uint32_t* in_valid; const int32_t* sel; uint32_t* out_valid;
    __device__ __forceinline__ int operator()(long long i) const {
        int32_t row = sel[i];
        if (out_valid && ((in_valid[row >> 5] >> (row & 31)) & 1u)) atomicOr(out_valid + (i >> 5), 1u << (i & 31));
        return in_off[row + 1] - in_off[row];
    }
};
__global__ void k_utf8_gather_bytes(const int32_t* __restrict__ in_off, const uint8_t* __restrict__ in_data, const int32_t* __restrict__ sel,
                                    const unsigned long long* __restrict__ d_count, const int32_t* __restrict__ out_off, uint8_t* out_data) {
    long long m = (long long)*d_count;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        int32_t row = sel[i];
        int a = in_off[row], len = in_off[row + 1] - a, o = out_off[i];
        for (int b = 0; b < len; b++) out_data[o + b] = in_data[a + b];
    }
}

// Shared memory available for the stage ring of one CTA when `ctas` CTAs share an SM.
int stage_budget(kq_ctx* ctx, int ctas) {
    return (ctx->max_smem_optin + 1024) / ctas - 1024 - 2048;   // 1 KB/CTA reserved by the driver, 2 KB static
}
int launch_check(kq_ctx* ctx, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return kq_cuda_fail(ctx, e, what);
    ctx->launches++;
    return KQ_OK;
}

int small_grid(kq_ctx* ctx, int64_t items) {
    int64_t g = (items + 255) / 256, cap = (int64_t)ctx->sm_count * 8;
    return (int)std::max<int64_t>(1, std::min(g, cap));
}

// Gather one whole column by a device-resident selection vector with a device-resident count.
int gather_column(kq_ctx* ctx, kq_col* in, const int32_t* sel, kq_lazy_count* lazy, int64_t cap_rows, kq_col** out) {
    kq_col* c = nullptr;
    bool nullable = in->validity != nullptr;
    KQ_RET(kq_col_new(ctx, in->type, cap_rows, nullable, in->type == KQ_UTF8 ? in->data_bytes : 0, &c));
    c->n = -1; c->lazy = lazy; lazy->rc.fetch_add(1);
    if (nullable) cudaMemsetAsync(c->validity, 0, (size_t)((cap_rows + 63) / 64) * 8, ctx->stream);
    int g = small_grid(ctx, cap_rows);
    int st = KQ_OK;
    switch (in->type) {
        case KQ_F64: case KQ_I64:
            k_gather_fixed<uint64_t><<<g, 256, 0, ctx->stream>>>((const uint64_t*)in->data, in->validity, sel, lazy->d_slot, (uint64_t*)c->data, c->validity);
            st = launch_check(ctx, "k_gather_fixed"); break;
        case KQ_DATE32: case KQ_I32:
            k_gather_fixed<uint32_t><<<g, 256, 0, ctx->stream>>>((const uint32_t*)in->data, in->validity, sel, lazy->d_slot, (uint32_t*)c->data, c->validity);
            st = launch_check(ctx, "k_gather_fixed"); break;
        case KQ_BOOL:
            cudaMemsetAsync(c->data, 0, (size_t)((cap_rows + 63) / 64) * 8, ctx->stream);
            k_gather_bits<<<g, 256, 0, ctx->stream>>>((const uint32_t*)in->data, in->validity, sel, lazy->d_slot, (uint32_t*)c->data, c->validity);
            st = launch_check(ctx, "k_gather_bits"); break;
        case KQ_UTF8: {
            int64_t ntiles = (cap_rows + SCAN_TILE - 1) / SCAN_TILE + 1;
            unsigned long long* scratch = nullptr;
            st = kq_dev_alloc(ctx, (size_t)(ntiles + 2) * 8, (void**)&scratch);
            if (st != KQ_OK) break;
            cudaMemsetAsync(scratch, 0, (size_t)(ntiles + 2) * 8, ctx->stream);
            if ((st = kq_dev_alloc(ctx, 8, (void**)&c->d_utf8_bytes)) != KQ_OK) break;
            c->data_bytes = -1;
            int sg = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, (int64_t)ctx->sm_count * 4));
            GatherLen gl{in->offsets, in->validity, sel, c->validity};
            k_exclusive_offsets<GatherLen><<<sg, 256, 0, ctx->stream>>>(gl, lazy->d_slot, c->offsets, scratch + 2, (unsigned int*)scratch, c->d_utf8_bytes);
            st = launch_check(ctx, "k_exclusive_offsets");
            if (st == KQ_OK) {
                k_utf8_gather_bytes<<<g, 256, 0, ctx->stream>>>(in->offsets, (const uint8_t*)in->data, sel, lazy->d_slot, c->offsets, (uint8_t*)c->data);
                st = launch_check(ctx, "k_utf8_gather_bytes");
            }
            kq_dev_free(ctx, scratch);
            break;
        }
        default: st = kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "unknown column type");
    }
    if (st != KQ_OK) { kq_column_free(c); return st; }
    *out = c;
    return KQ_OK;
}

// Everything about one kernel launch that depends on the query SHAPE only: the generated source, the
// stage plan and the shared-memory budget. No CUDA calls (kq_explain_* runs it without a device).
struct OpsPlan {
    std::vector<int> out_type;        // per fused output
    std::vector<char> out_nullable;
    std::string defines, gen;
    const char* entry = "";
    StagePlan sp;
    int smem = 0;
    bool fits = true;                 // false: too many fused outputs for the stash, split the launch
};

// `ex`: the expressions this launch computes (at most MAX_OUT); `selvec`: also emit the selection vector.
int plan_ops(kq_ctx* ctx, KqCodegen& cg, kq_expr* pred, const std::vector<kq_expr*>& ex, bool selvec, int smem_optin, OpsPlan* P) {
    std::string pred_body, proj_body;
    if (pred) {
        KqVal p;
        KQ_RET(cg.value(pred, &p));
        if (p.type != KQ_BOOL) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "filter predicate is not Bool");
        cg.line("return " + p.v + (p.nullable() ? " & " + p.ok : std::string()) + ";");     // TRUE only (rule E3)
        pred_body = cg.take_body();
    }
    if ((int)ex.size() > MAX_OUT) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "more than %d computed outputs in one kernel", MAX_OUT);
    bool any_nullable = false;
    std::string types, nulls;
    for (size_t k = 0; k < ex.size(); k++) {
        KqVal v;
        KQ_RET(cg.value(ex[k], &v));
        if (v.type == KQ_UTF8) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "expressions cannot produce Utf8 values (only column pass-through)");
        P->out_type.push_back(v.type); P->out_nullable.push_back(v.nullable());
        any_nullable |= v.nullable();
        const std::string nul = v.nullable() ? "true" : "false";
        types += (k ? ", " : "") + std::to_string(v.type); nulls += (k ? ", " : "") + nul;
        if (v.type == KQ_BOOL) cg.line("sink.emit_bool(" + std::to_string(k) + ", " + v.v + ", " + v.okx() + ", " + nul + ", rc);");
        else {
            const KqVal a = cg.as_array(v);
            cg.line("sink.template emit<" + std::to_string(v.type) + ">(" + std::to_string(k) + ", " + a.v + ", " + a.okx() + ", " + nul + ", rc);");
        }
    }
    proj_body = cg.take_body();
    if (ex.empty()) { types = "0"; nulls = "false"; }

    // shared memory: [stage ring][per-warp stash rings (filter only)]; 1 KB/CTA is reserved by the driver,
    // up to 9 KB static (barriers, per-tile metadata)
    const int avail = smem_optin - 10240;
    const int nout = (int)ex.size();
    int cap = 32 * OPS_R, stash_total = 0;
    std::string stage_defs;
    if (pred) {
        const int row_bytes = 8 * nout + (any_nullable ? nout : 0) + (selvec ? 4 : 0);
        cg.plan_stages(1 << 30, 1, TILE, &P->sp);
        const int stage_all = P->sp.stage_bytes;           // one stage with every referenced buffer staged
        // the largest stash ring (rows per warp, power of two) that leaves three input stages, else two
        bool found = false;
        for (int want = 3; want >= 2 && !found; want--)
            for (int c = 8 * 32 * OPS_R; c >= 32 * OPS_R; c >>= 1)
                if (OPS_WARPS * c * row_bytes + want * stage_all <= avail) { cap = c; found = true; break; }
        if (!found) {
            cap = 32 * OPS_R;
            if (nout > 1 || OPS_WARPS * cap * row_bytes + 16 * 1024 > avail) { P->fits = false; return KQ_OK; }
        }
        if (const char* e = getenv("KQ_OPS_CAP")) cap = std::max(32 * OPS_R, std::min(cap, atoi(e)));     // tuning experiments
        stash_total = (OPS_WARPS * cap * row_bytes + 127) / 128 * 128;
    }
    stage_defs = cg.plan_stages(avail - stash_total, 2, TILE, &P->sp);
    P->smem = P->sp.nstages * P->sp.stage_bytes + stash_total;
    P->gen = "namespace kq {\n" + stage_defs + "struct Q {\n";
    P->gen += "    static constexpr int NOUT = " + std::to_string(nout) + ";\n";
    P->gen += "    static constexpr int OUT_TYPE[" + std::to_string(std::max(nout, 1)) + "] = {" + types + "};\n";
    P->gen += "    static constexpr bool OUT_NULLABLE[" + std::to_string(std::max(nout, 1)) + "] = {" + nulls + "};\n";
    P->gen += std::string("    static constexpr bool ANY_NULLABLE = ") + (any_nullable ? "true" : "false") + ";\n";
    if (pred) P->gen += "    static __device__ __forceinline__ uint32_t pred(const QArgs& q, const RowCtx& rc) {\n" + pred_body + "    }\n";
    P->gen += "    template <class Sink> static __device__ __forceinline__ void project(const QArgs& q, const RowCtx& rc, const Sink& sink) {\n" + proj_body + "    }\n";
    P->gen += "};\n}  // namespace kq\n";
    P->defines = "#define KQ_R " + std::to_string(OPS_R) + "\n#define KQ_WARPS " + std::to_string(OPS_WARPS) + "\n#define KQ_STAGES " +
                 std::to_string(P->sp.nstages) + "\n";
    if (getenv("KQ_TRACE_FILE")) P->defines += "#define KQ_TRACE 1\n";
    if (pred) P->defines += "#define KQ_KERNEL_FILTER\n#define KQ_STASH_ROWS " + std::to_string(cap) + "\n#define KQ_META 32\n#define KQ_SELVEC " + (selvec ? "1" : "0") + "\n";
    else P->defines += "#define KQ_KERNEL_PROJECT\n";
    P->entry = pred ? "kq_filter_project" : "kq_project";
    return KQ_OK;
}

// Shared implementation of kq_project / kq_filter_project / kq_filter / kq_expr_evaluate.
int run_operator(kq_ctx* ctx, kq_expr* pred, kq_expr* const* exprs, int nexprs, kq_batch* input, bool all_columns,
                 kq_batch** out, kq_col** selection) {
    if (!ctx || !input || !out || nexprs < 0) return KQ_ERR_ILLEGAL_ARGUMENT;
    cudaSetDevice(ctx->device);
    int64_t n; KQ_RET(kq_batch_resolve_rows(ctx, input, &n));
    for (kq_col* c : input->cols) KQ_RET(kq_col_resolve_rows(ctx, c, nullptr));

    // expressions of a filter-all: one ColumnExpression per input field
    std::vector<kq_expr*> owned;
    std::vector<kq_expr*> ex(exprs, exprs + nexprs);
    if (all_columns) for (int i = 0; i < (int)input->cols.size(); i++) { owned.push_back(kq_expr_column(i)); ex.push_back(owned.back()); }

    std::vector<kq_col*> outs(ex.size(), nullptr);
    kq_lazy_count* lazy = nullptr;
    kq_col* selcol = nullptr;
    auto fail = [&](int s) {
        for (kq_col* c : outs) kq_column_free(c);
        kq_column_free(selcol);
        if (lazy) kq_lazy_release(ctx, lazy);
        for (kq_expr* e : owned) kq_expr_free(e);
        return s;
    };

    // classify outputs: alias (bare column without a filter, rule R4), gather by selection vector (Utf8
    // pass-through below a filter), or computed by the fused kernel
    std::vector<int> mode(ex.size(), 0);   // 0 = fused, 1 = alias, 2 = gather
    std::vector<size_t> fused;
    bool need_sel = selection != nullptr;
    {
        KqCodegen probe;
        int st = probe.begin(ctx, input);
        if (st != KQ_OK) return fail(st);
        if (pred) { int t; bool nl; if ((st = probe.infer(pred, &t, &nl)) != KQ_OK) return fail(st); }
        for (size_t k = 0; k < ex.size(); k++) {
            int bc = KqCodegen::bare_column(ex[k]);
            int t; bool nl;
            if ((st = probe.infer(ex[k], &t, &nl)) != KQ_OK) return fail(st);
            if (bc >= 0 && !pred) mode[k] = 1;
            else if (bc >= 0 && t == KQ_UTF8) { mode[k] = 2; need_sel = true; }
            else fused.push_back(k);
        }
    }
    if (pred) {
        lazy = kq_lazy_new(ctx);
        if (!lazy) return fail(kq_fail(ctx, KQ_ERR_OUT_OF_MEMORY, "lazy count"));
        if (need_sel) {
            int st = kq_col_new(ctx, KQ_I32, n, false, 0, &selcol);
            if (st != KQ_OK) return fail(st);
            selcol->n = -1; selcol->lazy = lazy; lazy->rc.fetch_add(1);
        }
    }

    // One launch computes a chunk of the fused outputs; a filter whose outputs do not fit the stash is
    // split into several launches that repeat the predicate.
    std::vector<std::vector<size_t>> chunks;
    if (!fused.empty() || pred) chunks.push_back(fused);
    for (size_t ci = 0; ci < chunks.size(); ci++) {
        const bool first = ci == 0;
        std::vector<kq_expr*> cex;
        for (size_t k : chunks[ci]) cex.push_back(ex[k]);
        KqCodegen cg;
        OpsPlan P;
        int st = cg.begin(ctx, input);
        if (st == KQ_OK && !pred && cex.size() > (size_t)MAX_OUT) P.fits = false;
        else if (st == KQ_OK) st = plan_ops(ctx, cg, pred, cex, first && need_sel, ctx->max_smem_optin, &P);
        if (st != KQ_OK) return fail(st);
        if (!P.fits) {
            if (chunks[ci].size() < 2) return fail(kq_fail(ctx, KQ_ERR_UNSUPPORTED, "projection does not fit the shared-memory budget"));
            std::vector<size_t> lo(chunks[ci].begin(), chunks[ci].begin() + chunks[ci].size() / 2), hi(chunks[ci].begin() + chunks[ci].size() / 2, chunks[ci].end());
            chunks[ci] = lo;
            chunks.insert(chunks.begin() + (long)ci + 1, hi);
            ci--;
            continue;
        }
        OpArgs A;
        memset(&A, 0, sizeof A);
        A.n = n; A.ntiles = (n + TILE - 1) / TILE; A.err = ctx->d_err;
        A.q = cg.args; A.sp = P.sp;
        for (size_t j = 0; j < chunks[ci].size(); j++) {
            const size_t k = chunks[ci][j];
            kq_col* c = nullptr;
            if ((st = kq_col_new(ctx, P.out_type[j], n, P.out_nullable[j], 0, &c)) != KQ_OK) return fail(st);
            outs[k] = c;
            if (pred) {
                c->n = -1; c->lazy = lazy; lazy->rc.fetch_add(1);
                if (c->validity) cudaMemsetAsync(c->validity, 0, (size_t)((n + 63) / 64) * 8, ctx->stream);
                if (P.out_type[j] == KQ_BOOL) cudaMemsetAsync(c->data, 0, (size_t)((n + 63) / 64) * 8, ctx->stream);
            }
            A.outs[j].data = c->data; A.outs[j].validity = c->validity;
        }
        if (n == 0) continue;
        void* kernel = nullptr;
        if ((st = kq_jit_kernel(ctx, P.defines, P.gen, KQ_SKEL_OPS, P.entry, P.smem, &kernel)) != KQ_OK) return fail(st);
        void* kargs[] = {&A};
        const int grid = (int)std::min<int64_t>(A.ntiles, (int64_t)ctx->sm_count);
        unsigned long long* scratch = nullptr;     // [0]: ticket, [2..]: tile descriptors
        if (pred) {
            if ((st = kq_dev_alloc(ctx, (size_t)(A.ntiles + 2) * 8, (void**)&scratch)) != KQ_OK) return fail(st);
            cudaMemsetAsync(scratch, 0, (size_t)(A.ntiles + 2) * 8, ctx->stream);
            A.ticket = (unsigned int*)scratch;
            A.tile_desc = scratch + 2;
            A.out_count = lazy->d_slot;
            if (first && need_sel) A.selvec = (int32_t*)selcol->data;
            if (getenv("KQ_TRACE_FILE")) {       // debugging aid: per-tile timestamps of the pipeline (needs a KQ_TRACE build)
                kq_dev_alloc(ctx, (size_t)A.ntiles * 64, (void**)&A.trace);
                cudaMemsetAsync(A.trace, 0, (size_t)A.ntiles * 64, ctx->stream);
            }
        }
        // the filter kernel's CTAs wait for one another (cross-block prefix): cooperative launch = all resident
        cudaError_t e = pred ? cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(THREADS_FILTER), kargs, (size_t)P.smem, ctx->stream)
                             : cudaLaunchKernel(kernel, dim3(grid), dim3(THREADS_PROJECT), kargs, (size_t)P.smem, ctx->stream);
        kq_dev_free(ctx, scratch);
        if (A.trace) {
            std::vector<unsigned long long> h((size_t)A.ntiles * 8);
            cudaMemcpyAsync(h.data(), A.trace, h.size() * 8, cudaMemcpyDeviceToHost, ctx->stream);
            cudaStreamSynchronize(ctx->stream);
            if (FILE* f = fopen(getenv("KQ_TRACE_FILE"), "wb")) { fwrite(h.data(), 8, h.size(), f); fclose(f); }
            kq_dev_free(ctx, A.trace);
        }
        if (e != cudaSuccess) return fail(kq_cuda_fail(ctx, e, P.entry));
        ctx->launches++;
    }

    if (pred) {
        for (size_t k = 0; k < ex.size(); k++) {
            if (mode[k] != 2) continue;
            kq_col* in = input->cols[(size_t)KqCodegen::bare_column(ex[k])];
            int st = gather_column(ctx, in, (const int32_t*)selcol->data, lazy, n, &outs[k]);
            if (st != KQ_OK) return fail(st);
        }
        // the row count travels back asynchronously; it is only waited for when somebody asks
        cudaMemcpyAsync(lazy->h_slot, lazy->d_slot, 8, cudaMemcpyDeviceToHost, ctx->stream);
        cudaEventRecord(lazy->ev, ctx->stream);
    } else {
        for (size_t k = 0; k < ex.size(); k++)
            if (mode[k] == 1) { outs[k] = input->cols[(size_t)KqCodegen::bare_column(ex[k])]; outs[k]->rc.fetch_add(1); }
    }

    kq_batch* b = new kq_batch();
    b->ctx = ctx;
    b->cols = outs;
    if (pred) { b->n = -1; b->lazy = lazy; } else b->n = n;
    *out = b;
    if (selection) *selection = selcol; else kq_column_free(selcol);
    for (kq_expr* e : owned) kq_expr_free(e);
    return KQ_OK;
}


}  // namespace

extern "C" {

int kq_project(kq_ctx* ctx, kq_expr* const* exprs, int nexprs, kq_batch* input, kq_batch** out) {
    return run_operator(ctx, nullptr, exprs, nexprs, input, false, out, nullptr);
}

int kq_filter_project(kq_ctx* ctx, kq_expr* pred, kq_expr* const* exprs, int nexprs, kq_batch* input, kq_batch** out) {
    if (!pred) return kq_fail(ctx, KQ_ERR_ILLEGAL_ARGUMENT, "kq_filter_project needs a predicate");
    return run_operator(ctx, pred, exprs, nexprs, input, false, out, nullptr);
}

int kq_filter(kq_ctx* ctx, kq_expr* pred, kq_batch* input, kq_batch** out, kq_col** selection) {
    if (!pred) return kq_fail(ctx, KQ_ERR_ILLEGAL_ARGUMENT, "kq_filter needs a predicate");
    return run_operator(ctx, pred, nullptr, 0, input, true, out, selection);
}

int kq_expr_evaluate(kq_ctx* ctx, kq_expr* e, kq_batch* input, kq_col** out) {
    if (!out) return KQ_ERR_ILLEGAL_ARGUMENT;
    kq_batch* b = nullptr;
    kq_expr* ex[1] = {e};
    KQ_RET(run_operator(ctx, nullptr, ex, 1, input, false, &b, nullptr));
    *out = b->cols[0];
    (*out)->rc.fetch_add(1);
    kq_batch_free(b);
    return KQ_OK;
}

int kq_explain_filter_project(kq_expr* pred, kq_expr* const* exprs, int nexprs, int ncols, const int* types, const int* nullable,
                              int compile, char* source, size_t source_cap) {
    if (ncols < 0 || nexprs < 0 || (ncols > 0 && !types)) return KQ_ERR_ILLEGAL_ARGUMENT;
    kq_ctx fake;
    KqSchemaBatch sb(ncols, types, nullable);
    KqCodegen cg;
    OpsPlan P;
    std::vector<kq_expr*> ex(exprs, exprs + nexprs);
    int st = cg.begin(&fake, &sb.batch);
    if (st == KQ_OK) st = plan_ops(&fake, cg, pred, ex, false, 232448, &P);
    if (st == KQ_OK && !P.fits) st = kq_fail(&fake, KQ_ERR_UNSUPPORTED, "outputs do not fit one launch (the operator would split them)");
    if (st == KQ_OK && compile) st = kq_jit_compile_only(&fake, P.defines, P.gen, KQ_SKEL_OPS);
    kq_copy_text(st == KQ_OK ? P.defines + P.gen : fake.last_error, source, source_cap);
    return st;
}

int kq_filter_project_host(kq_ctx* ctx, kq_expr* pred, kq_expr* const* exprs, int nexprs, int ncols, const int* types,
                           const uint8_t* const* validity, const void* const* data, int64_t n, void* const* out_data,
                           uint8_t* const* out_validity, int64_t* out_rows) {
    // v1: upload -> fused kernel -> download, sequential (chunked overlap is a later optimisation)
    if (!ctx || ncols < 0) return KQ_ERR_ILLEGAL_ARGUMENT;
    std::vector<kq_col*> cols((size_t)ncols, nullptr);
    int st = KQ_OK;
    for (int i = 0; i < ncols && st == KQ_OK; i++) {
        if (types[i] == KQ_UTF8) st = kq_fail(ctx, KQ_ERR_UNSUPPORTED, "kq_filter_project_host: Utf8 columns not supported");
        else st = kq_column_upload(ctx, types[i], n, validity ? validity[i] : nullptr, nullptr, data[i], 0, &cols[(size_t)i]);
    }
    kq_batch *in = nullptr, *res = nullptr;
    if (st == KQ_OK) st = kq_batch_create(ctx, cols.data(), ncols, n, &in);
    if (st == KQ_OK) st = kq_filter_project(ctx, pred, exprs, nexprs, in, &res);
    int64_t m = 0;
    if (st == KQ_OK) st = kq_batch_num_rows(ctx, res, &m);
    for (int k = 0; k < nexprs && st == KQ_OK; k++)
        st = kq_column_download(ctx, res->cols[(size_t)k], out_validity ? out_validity[k] : nullptr, nullptr, out_data[k]);
    if (st == KQ_OK && out_rows) *out_rows = m;
    kq_batch_free(res); kq_batch_free(in);
    for (kq_col* c : cols) kq_column_free(c);
    return st;
}

}  // extern "C"
