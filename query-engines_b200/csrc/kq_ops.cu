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
    if ((int)ex.size() > MAX_OUT) { P->fits = false; return KQ_OK; }       // the caller splits the projection into several launches
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
    if (const char* e = getenv("KQ_L2_PREFETCH")) P->defines += "#define KQ_L2_PREFETCH " + std::to_string(atoi(e)) + "\n";     // tuning experiments
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
        A.tile_begin = 0; A.tile_end = A.ntiles;
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

// FilterExec + ProjectionExec end to end from HOST Arrow buffers to HOST result buffers. The input is
// streamed chunk by chunk: pinned H2D copies on one side stream (double-buffered device chunks), the
// fused kernel on the compute stream over the chunk's tile range of ONE logical batch (the cross-block
// prefix simply continues from the previous launch, so the compacted output stays globally ordered),
// D2H of the rows each chunk produced on a second side stream. PCIe is full duplex, so the step costs
// about max(H2D, D2H) instead of their sum plus the kernel.
int kq_filter_project_host(kq_ctx* ctx, kq_expr* pred, kq_expr* const* exprs, int nexprs, int ncols, const int* types,
                           const uint8_t* const* validity, const void* const* data, int64_t n, void* const* out_data,
                           uint8_t* const* out_validity, int64_t* out_rows) {
    if (!ctx || ncols < 0 || nexprs < 0 || n < 0 || (ncols > 0 && (!types || !data))) return KQ_ERR_ILLEGAL_ARGUMENT;
    if (!pred) return kq_fail(ctx, KQ_ERR_ILLEGAL_ARGUMENT, "kq_filter_project_host needs a predicate");
    if (n > 2147483647LL) return kq_fail(ctx, KQ_ERR_ILLEGAL_ARGUMENT, "batch larger than 2^31-1 rows");
    cudaSetDevice(ctx->device);
    std::vector<int> nullable((size_t)ncols, 0);
    for (int i = 0; i < ncols; i++) {
        if (types[i] == KQ_UTF8) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "kq_filter_project_host: Utf8 columns are not supported");
        nullable[(size_t)i] = validity && validity[i];
    }
    // plan + kernel (schema only)
    KqSchemaBatch sb(ncols, types, nullable.data());
    KqCodegen cg;
    OpsPlan P;
    std::vector<kq_expr*> ex(exprs, exprs + nexprs);
    KQ_RET(cg.begin(ctx, &sb.batch));
    KQ_RET(plan_ops(ctx, cg, pred, ex, false, ctx->max_smem_optin, &P));
    if (!P.fits) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "kq_filter_project_host: too many outputs for one launch");
    if (out_rows) *out_rows = 0;
    if (n == 0) return KQ_OK;
    void* kernel = nullptr;
    KQ_RET(kq_jit_kernel(ctx, P.defines, P.gen, KQ_SKEL_OPS, P.entry, P.smem, &kernel));

    const int64_t CH = (int64_t)TILE * 2048;          // rows per chunk (4 Mi): a multiple of the tile and of 64 (bitmap words)
    const int64_t nchunks = (n + CH - 1) / CH, ntiles = (n + TILE - 1) / TILE;
    auto width = [](int t) { return t == KQ_DATE32 || t == KQ_I32 ? 4 : 8; };

    // device memory: double-buffered input chunks, full-size outputs, look-back descriptors
    std::vector<void*> allocs;
    cudaEvent_t h2d_done[2] = {nullptr, nullptr}, k_done[2] = {nullptr, nullptr}, cnt_ev[2] = {nullptr, nullptr};
    uint64_t* h_cnt = nullptr;
    auto cleanup = [&](int s) {
        cudaStreamSynchronize(ctx->copy_stream[0]); cudaStreamSynchronize(ctx->copy_stream[1]); cudaStreamSynchronize(ctx->stream);
        for (void* p : allocs) kq_dev_free(ctx, p);
        for (int i = 0; i < 2; i++) { if (h2d_done[i]) cudaEventDestroy(h2d_done[i]); if (k_done[i]) cudaEventDestroy(k_done[i]); if (cnt_ev[i]) cudaEventDestroy(cnt_ev[i]); }
        if (h_cnt) cudaFreeHost(h_cnt);
        return s;
    };
    auto dalloc = [&](size_t bytes, void** p) { int s = kq_dev_alloc(ctx, bytes, p); if (s == KQ_OK) allocs.push_back(*p); return s; };
    int st = KQ_OK;
    std::vector<void*> in_d[2], in_v[2];
    for (int s = 0; s < 2 && st == KQ_OK; s++) {
        in_d[s].assign((size_t)ncols, nullptr); in_v[s].assign((size_t)ncols, nullptr);
        for (int c = 0; c < ncols && st == KQ_OK; c++) {
            st = dalloc(types[c] == KQ_BOOL ? (size_t)CH / 8 : (size_t)CH * width(types[c]), &in_d[s][(size_t)c]);
            if (st == KQ_OK && nullable[(size_t)c]) st = dalloc((size_t)CH / 8, &in_v[s][(size_t)c]);
        }
    }
    std::vector<void*> o_d((size_t)nexprs, nullptr), o_v((size_t)nexprs, nullptr);
    for (int k = 0; k < nexprs && st == KQ_OK; k++) {
        const size_t bytes = P.out_type[(size_t)k] == KQ_BOOL ? (size_t)((n + 63) / 64) * 8 : (size_t)n * width(P.out_type[(size_t)k]);
        st = dalloc(bytes, &o_d[(size_t)k]);
        if (st == KQ_OK && P.out_type[(size_t)k] == KQ_BOOL) cudaMemsetAsync(o_d[(size_t)k], 0, bytes, ctx->stream);
        if (st == KQ_OK && P.out_nullable[(size_t)k]) {
            st = dalloc((size_t)((n + 63) / 64) * 8, &o_v[(size_t)k]);
            if (st == KQ_OK) cudaMemsetAsync(o_v[(size_t)k], 0, (size_t)((n + 63) / 64) * 8, ctx->stream);
        }
    }
    unsigned long long* scratch = nullptr;     // [0]: out_count, [2..]: tile descriptors
    if (st == KQ_OK) st = dalloc((size_t)(ntiles + 2) * 8, (void**)&scratch);
    if (st != KQ_OK) return cleanup(st);
    cudaMemsetAsync(scratch, 0, (size_t)(ntiles + 2) * 8, ctx->stream);
    for (int i = 0; i < 2; i++) {
        cudaEventCreateWithFlags(&h2d_done[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&k_done[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&cnt_ev[i], cudaEventDisableTiming);
    }
    if (cudaMallocHost(&h_cnt, 2 * sizeof(uint64_t)) != cudaSuccess) return cleanup(kq_fail(ctx, KQ_ERR_OUT_OF_MEMORY, "pinned count slots"));
    // the side streams start after everything queued so far on the compute stream (allocator order, memsets)
    cudaEventRecord(ctx->copy_done, ctx->stream);
    cudaStreamWaitEvent(ctx->copy_stream[0], ctx->copy_done, 0);
    cudaStreamWaitEvent(ctx->copy_stream[1], ctx->copy_done, 0);

    OpArgs A;
    memset(&A, 0, sizeof A);
    A.n = n; A.ntiles = ntiles; A.err = ctx->d_err; A.sp = P.sp;
    A.q = cg.args;
    A.tile_desc = scratch + 2; A.out_count = scratch; A.ticket = (unsigned int*)(scratch + 1);
    for (int k = 0; k < nexprs; k++) { A.outs[k].data = o_d[(size_t)k]; A.outs[k].validity = (uint32_t*)o_v[(size_t)k]; }
    void* kargs[] = {&A};

    // D2H of output rows [from, to) (bit-packed buffers move whole 32-bit words: boundary words are rewritten by later chunks)
    auto download = [&](int64_t from, int64_t to) {
        if (to <= from) return;
        for (int k = 0; k < nexprs; k++) {
            const int t = P.out_type[(size_t)k];
            const int64_t w0 = from / 32, w1 = (to + 31) / 32;
            if (t == KQ_BOOL) cudaMemcpyAsync((char*)out_data[k] + w0 * 4, (char*)o_d[(size_t)k] + w0 * 4, (size_t)(w1 - w0) * 4, cudaMemcpyDeviceToHost, ctx->copy_stream[1]);
            else cudaMemcpyAsync((char*)out_data[k] + from * width(t), (char*)o_d[(size_t)k] + from * width(t), (size_t)(to - from) * width(t), cudaMemcpyDeviceToHost, ctx->copy_stream[1]);
            if (out_validity && out_validity[k]) {
                if (o_v[(size_t)k]) cudaMemcpyAsync(out_validity[k] + w0 * 4, (char*)o_v[(size_t)k] + w0 * 4, (size_t)(w1 - w0) * 4, cudaMemcpyDeviceToHost, ctx->copy_stream[1]);
                else memset(out_validity[k] + from / 8, 0xFF, (size_t)((to + 7) / 8 - from / 8));       // no nulls: all-ones (kq_column_download convention)
            }
        }
    };

    int64_t done_rows = 0;          // output rows already handed to the D2H stream
    for (int64_t i = 0; i <= nchunks; i++) {
        if (i < nchunks) {
            const int s = (int)(i & 1);
            const int64_t row0 = i * CH, rows = std::min(CH, n - row0);
            if (i >= 2) cudaStreamWaitEvent(ctx->copy_stream[0], k_done[s], 0);       // the chunk buffer is free once its last kernel ran
            for (int c = 0; c < ncols; c++) {
                const size_t bytes = types[c] == KQ_BOOL ? (size_t)((rows + 7) / 8) : (size_t)rows * width(types[c]);
                const size_t off = types[c] == KQ_BOOL ? (size_t)(row0 / 8) : (size_t)row0 * width(types[c]);
                cudaMemcpyAsync(in_d[s][(size_t)c], (const char*)data[c] + off, bytes, cudaMemcpyHostToDevice, ctx->copy_stream[0]);
                if (nullable[(size_t)c]) cudaMemcpyAsync(in_v[s][(size_t)c], validity[c] + row0 / 8, (size_t)((rows + 7) / 8), cudaMemcpyHostToDevice, ctx->copy_stream[0]);
            }
            cudaEventRecord(h2d_done[s], ctx->copy_stream[0]);
            cudaStreamWaitEvent(ctx->stream, h2d_done[s], 0);
            // the kernel indexes rows of the whole batch: bias the chunk pointers back by the chunk's first row
            for (int q = 0; q < cg.ncols; q++) {
                const int c = cg.slot_col[q];
                const size_t off = types[c] == KQ_BOOL ? (size_t)(row0 / 8) : (size_t)row0 * width(types[c]);
                A.q.cols[q].data = (const char*)in_d[s][(size_t)c] - off;
                A.q.cols[q].validity = nullable[(size_t)c] ? (const uint32_t*)((const char*)in_v[s][(size_t)c] - row0 / 8) : nullptr;
            }
            stage_plan_bind(A.sp, A.q);          // the TMA producer reads through the same biased bases
            A.tile_begin = row0 / TILE; A.tile_end = (row0 + rows + TILE - 1) / TILE;
            const int grid = (int)std::min<int64_t>(A.tile_end - A.tile_begin, (int64_t)ctx->sm_count);
            cudaError_t e = cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(THREADS_FILTER), kargs, (size_t)P.smem, ctx->stream);
            if (e != cudaSuccess) return cleanup(kq_cuda_fail(ctx, e, "kq_filter_project"));
            ctx->launches++;
            cudaEventRecord(k_done[s], ctx->stream);
            // running total = inclusive prefix of the chunk's last tile
            cudaMemcpyAsync(&h_cnt[s], A.tile_desc + (A.tile_end - 1), 8, cudaMemcpyDeviceToHost, ctx->stream);
            cudaEventRecord(cnt_ev[s], ctx->stream);
        }
        if (i >= 1) {       // results of chunk i-1: its kernel is queued behind nothing but chunk i's copies
            const int s = (int)((i - 1) & 1);
            cudaError_t e = cudaEventSynchronize(cnt_ev[s]);
            if (e != cudaSuccess) return cleanup(kq_cuda_fail(ctx, e, "cudaEventSynchronize"));
            const int64_t total = (int64_t)(h_cnt[s] & ((1ULL << 62) - 1ULL));
            cudaStreamWaitEvent(ctx->copy_stream[1], k_done[s], 0);
            download(done_rows, total);
            done_rows = total;
        }
    }
    cudaError_t e = cudaStreamSynchronize(ctx->copy_stream[1]);
    if (e != cudaSuccess) return cleanup(kq_cuda_fail(ctx, e, "D2H"));
    if (out_rows) *out_rows = done_rows;
    st = kq_check_device_errors(ctx);
    return cleanup(st);
}

}  // extern "C"
