// Host harness for csrc/kq_csv.cu (tests/test_csv_host.py): the few helpers of kq_core.cu that the CSV scan and reader use,
// restated over malloc, plus accessors for the test. Built with g++ against tests/host_shim/ (no CUDA). Test infrastructure
// only. The allocator poisons fresh memory and keeps guard bytes behind every block, so a read of unwritten bytes or a write
// past the end shows up as a mismatch or as kqh_check_guards() != 0.
#include <cstdarg>
#include <cstdio>
#include <map>

#include "kq_internal.h"

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

namespace {
constexpr size_t GUARD = 64;
std::map<const uint8_t*, size_t> g_blocks;          // every live "device" block: start -> logical size
int g_guard_errors = 0;
}

cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* at, const void* p) {
    at->type = cudaMemoryTypeUnregistered;
    auto it = g_blocks.upper_bound((const uint8_t*)p);
    if (it != g_blocks.begin()) { --it; if ((const uint8_t*)p < it->first + it->second + GUARD) at->type = cudaMemoryTypeDevice; }
    return cudaSuccess;
}

int kq_fail(kq_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->last_error = buf;
    return code;
}
int kq_cuda_fail(kq_ctx* ctx, cudaError_t e, const char* what) { return kq_fail(ctx, KQ_ERR_CUDA, "CUDA error %d at %s", (int)e, what); }

int kq_dev_alloc(kq_ctx*, size_t bytes, void** out) {
    // like the real allocator: 256-byte aligned; the real one pads by KQ_PAD zeroed bytes, here the bytes behind the block are guards
    uint8_t* p = (uint8_t*)aligned_alloc(256, ((bytes + GUARD + 255) / 256) * 256);
    memset(p, 0xA5, bytes);
    memset(p + bytes, 0x5C, GUARD);
    g_blocks[p] = bytes;
    *out = p;
    return KQ_OK;
}
void kq_dev_free(kq_ctx*, void* p) {
    if (!p) return;
    auto it = g_blocks.find((const uint8_t*)p);
    if (it == g_blocks.end()) { g_guard_errors++; return; }
    for (size_t i = 0; i < GUARD; i++) if (it->first[it->second + i] != 0x5C) { g_guard_errors++; break; }
    g_blocks.erase(it);
    free(p);
}
int kq_col_new(kq_ctx* ctx, int type, int64_t n, bool with_validity, int64_t utf8_bytes, kq_col** out) {
    kq_col* c = new kq_col();
    c->ctx = ctx; c->type = type; c->n = n; c->capacity_rows = n;
    c->data_bytes = type == KQ_UTF8 ? utf8_bytes : 0;
    kq_dev_alloc(ctx, type == KQ_UTF8 ? (size_t)utf8_bytes : (size_t)n * 8, &c->data);
    if (with_validity) kq_dev_alloc(ctx, (size_t)((n + 63) / 64) * 8, (void**)&c->validity);
    if (type == KQ_UTF8) kq_dev_alloc(ctx, (size_t)(n + 1) * 4, (void**)&c->offsets);
    *out = c;
    return KQ_OK;
}
int kq_read_u64(kq_ctx* ctx, const void* d_ptr, int count, uint64_t* out) {
    memcpy(ctx->h_scratch, d_ptr, (size_t)count * 8);
    for (int i = 0; i < count; i++) out[i] = ctx->h_scratch[i];
    return KQ_OK;
}

extern "C" {

int kq_column_free(kq_col* col) {
    if (!col) return KQ_OK;
    if (col->rc.fetch_sub(1) == 1) {
        kq_dev_free(col->ctx, col->data); kq_dev_free(col->ctx, col->validity); kq_dev_free(col->ctx, col->offsets);
        delete col;
    }
    return KQ_OK;
}
int kq_batch_free(kq_batch* b) {
    if (!b) return KQ_OK;
    if (b->rc.fetch_sub(1) == 1) { for (kq_col* c : b->cols) kq_column_free(c); delete b; }
    return KQ_OK;
}
const char* kq_last_error(kq_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "no context"; }

// ---- accessors for the test ------------------------------------------------------------------------------------------
KQ_API kq_ctx* kqh_ctx_new() {
    kq_ctx* c = new kq_ctx();
    c->h_scratch = (uint64_t*)calloc(64, 8);
    return c;
}
KQ_API void kqh_ctx_free(kq_ctx* c) { free(c->h_scratch); delete c; }
KQ_API void* kqh_device_buffer(kq_ctx* c, const uint8_t* bytes, int64_t n, int misalign) {       // a text "resident in HBM"
    void* p = nullptr;
    kq_dev_alloc(c, (size_t)n + 16, &p);
    memcpy((uint8_t*)p + misalign, bytes, (size_t)n);
    return p;
}
KQ_API void kqh_device_free(kq_ctx* c, void* p) { kq_dev_free(c, p); }
KQ_API int64_t kqh_batch_rows(kq_batch* b) { return b->n; }
KQ_API int kqh_batch_cols(kq_batch* b) { return (int)b->cols.size(); }
KQ_API int64_t kqh_col_bytes(kq_batch* b, int c) { return b->cols[(size_t)c]->data_bytes; }
KQ_API void kqh_col_read(kq_batch* b, int c, int32_t* offsets, uint8_t* data) {
    kq_col* col = b->cols[(size_t)c];
    memcpy(offsets, col->offsets, (size_t)(col->n + 1) * 4);
    if (col->data_bytes) memcpy(data, col->data, (size_t)col->data_bytes);
}
KQ_API int64_t kqh_live_blocks() { return (int64_t)g_blocks.size(); }
KQ_API int kqh_guard_errors() { return g_guard_errors; }

}  // extern "C"
