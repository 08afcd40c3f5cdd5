// kq_core.cu — context, device memory, columns, batches, synthetic tables.
#include <cstring>

#include "kq_internal.h"

// ---- errors ---------------------------------------------------------------------------------------------
int kq_fail(kq_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->last_error = buf;
    return code;
}
int kq_cuda_fail(kq_ctx* ctx, cudaError_t e, const char* what) {
    int code = (e == cudaErrorMemoryAllocation) ? KQ_ERR_OUT_OF_MEMORY : KQ_ERR_CUDA;
    return kq_fail(ctx, code, "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
}

extern "C" {

const char* kq_version(void) { return "kqgpu 0.1 (sm_100a)"; }

const char* kq_status_name(int s) {
    switch (s) {
        case KQ_OK: return "OK";
        case KQ_ERR_ILLEGAL_STATE: return "IllegalStateException";
        case KQ_ERR_UNSUPPORTED: return "UnsupportedOperationException";
        case KQ_ERR_ILLEGAL_ARGUMENT: return "IllegalArgumentException";
        case KQ_ERR_SQL: return "SQLException";
        case KQ_ERR_NUMBER_FORMAT: return "NumberFormatException";
        case KQ_ERR_ARITHMETIC: return "ArithmeticException";
        case KQ_ERR_OUT_OF_MEMORY: return "OutOfMemoryError";
        case KQ_ERR_CUDA: return "CudaError";
        case KQ_ERR_NCCL: return "NcclError";
        case KQ_ERR_NO_DEVICE: return "NoCudaDevice";
        default: return "Unknown";
    }
}

int kq_device_count(int* count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    if (count) *count = n;
    return KQ_OK;
}

int kq_ctx_create(int device, kq_ctx** out) {
    if (!out) return KQ_ERR_ILLEGAL_ARGUMENT;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return KQ_ERR_NO_DEVICE; }   // no CPU fallback
    if (device < 0 || device >= n) return KQ_ERR_ILLEGAL_ARGUMENT;
    kq_ctx* c = new kq_ctx();
    c->device = device;
    auto bail = [&](cudaError_t e, const char* what) { fprintf(stderr, "kq_ctx_create: %s: %s\n", what, cudaGetErrorString(e)); delete c; return KQ_ERR_CUDA; };
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail(e, "cudaGetDeviceProperties");
    c->sm_count = prop.multiProcessorCount;
    c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "stream");
    for (int i = 0; i < 2; i++)
        if ((e = cudaStreamCreateWithFlags(&c->copy_stream[i], cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "copy stream");
    if ((e = cudaEventCreateWithFlags(&c->copy_done, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "event");
    for (int i = 0; i < 2; i++)
        if ((e = cudaEventCreate(&c->timer[i])) != cudaSuccess) return bail(e, "timer event");
    if ((e = cudaMalloc(&c->d_err, 256)) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMemset(c->d_err, 0, 256)) != cudaSuccess) return bail(e, "cudaMemset");
    if ((e = cudaHostAlloc(&c->h_err, 64, cudaHostAllocDefault)) != cudaSuccess) return bail(e, "cudaHostAlloc");
    if ((e = cudaHostAlloc(&c->h_scratch, 64 * sizeof(uint64_t), cudaHostAllocDefault)) != cudaSuccess) return bail(e, "cudaHostAlloc");
    *c->h_err = 0;
    const int nslots = 4096;
    if ((e = cudaHostAlloc(&c->pinned_slots, nslots * 16, cudaHostAllocDefault)) != cudaSuccess) return bail(e, "cudaHostAlloc");
    for (int i = nslots - 1; i >= 0; i--) c->free_slots.push_back(i);
    *out = c;
    return KQ_OK;
}

int kq_ctx_destroy(kq_ctx* ctx) {
    if (!ctx) return KQ_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (kq_lazy_count* l : ctx->pending_lazy) { cudaEventDestroy(l->ev); delete l; }
    ctx->pending_lazy.clear();
    if (ctx->d_flush) cudaFree(ctx->d_flush);
    for (auto& kv : ctx->block_size) cudaFree(kv.first);
    for (cudaEvent_t ev : ctx->free_events) cudaEventDestroy(ev);
    cudaFreeHost(ctx->pinned_slots);
    cudaFree(ctx->d_err);
    cudaFreeHost(ctx->h_err);
    cudaFreeHost(ctx->h_scratch);
    for (int i = 0; i < 2; i++) { cudaStreamDestroy(ctx->copy_stream[i]); cudaEventDestroy(ctx->timer[i]); }
    cudaEventDestroy(ctx->copy_done);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return KQ_OK;
}

const char* kq_last_error(kq_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "no context"; }
void* kq_ctx_stream(kq_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int64_t kq_ctx_launch_count(kq_ctx* ctx) { return ctx ? ctx->launches : 0; }

int kq_ctx_sync(kq_ctx* ctx) {
    if (!ctx) return KQ_ERR_ILLEGAL_ARGUMENT;
    KQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return kq_check_device_errors(ctx);
}

int kq_timer_begin(kq_ctx* ctx) {
    KQ_CUDA(ctx, cudaEventRecord(ctx->timer[0], ctx->stream));
    return KQ_OK;
}
int kq_timer_end(kq_ctx* ctx, float* ms) {
    KQ_CUDA(ctx, cudaEventRecord(ctx->timer[1], ctx->stream));
    KQ_CUDA(ctx, cudaEventSynchronize(ctx->timer[1]));
    KQ_CUDA(ctx, cudaEventElapsedTime(ms, ctx->timer[0], ctx->timer[1]));
    return KQ_OK;
}

int kq_flush_l2(kq_ctx* ctx, size_t bytes) {
    if (bytes > ctx->flush_bytes) {
        if (ctx->d_flush) cudaFree(ctx->d_flush);
        ctx->d_flush = nullptr; ctx->flush_bytes = 0;
        KQ_CUDA(ctx, cudaMalloc(&ctx->d_flush, bytes));
        ctx->flush_bytes = bytes;
    }
    KQ_CUDA(ctx, cudaMemsetAsync(ctx->d_flush, 0x5a, bytes, ctx->stream));
    return KQ_OK;
}

int kq_host_alloc(kq_ctx* ctx, size_t bytes, void** out) {
    KQ_CUDA(ctx, cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return KQ_OK;
}
int kq_host_free(kq_ctx* ctx, void* p) { KQ_CUDA(ctx, cudaFreeHost(p)); return KQ_OK; }
int kq_host_register(kq_ctx* ctx, void* p, size_t bytes) { KQ_CUDA(ctx, cudaHostRegister(p, bytes, cudaHostRegisterDefault)); return KQ_OK; }
int kq_host_unregister(kq_ctx* ctx, void* p) { KQ_CUDA(ctx, cudaHostUnregister(p)); return KQ_OK; }

}  // extern "C"

// ---- internal helpers --------------------------------------------------------------------------------------
int kq_check_device_errors(kq_ctx* ctx) {
    // d_err is only ever OR-ed by kernels; read it back and clear it.
    KQ_CUDA(ctx, cudaMemcpyAsync(ctx->h_err, ctx->d_err, 4, cudaMemcpyDeviceToHost, ctx->stream));
    KQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    uint32_t e = *ctx->h_err;
    if (!e) return KQ_OK;
    KQ_CUDA(ctx, cudaMemsetAsync(ctx->d_err, 0, 4, ctx->stream));
    return kq_device_error_status(ctx, e);
}

// kq_status (and message) for the error bits a kernel raised.
int kq_device_error_status(kq_ctx* ctx, uint32_t e) {
    if (!e) return KQ_OK;
    if (e & 0xFF00u) return kq_fail(ctx, KQ_ERR_CUDA, "aggregate kernel gave up (internal protocol error, location bits 0x%x)", e & 0xFF00u);
    if (e & 16u) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "two different Utf8 group keys share a 63-bit hash: refusing to merge their groups");
    if (e & 32u) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "more distinct long Utf8 group keys than the key heap was sized for (raise expected_groups)");
    if (e & KQ_DEV_ERR_DIV0) return kq_fail(ctx, KQ_ERR_ARITHMETIC, "/ by zero");
    if (e & KQ_DEV_ERR_NUMBER_FORMAT) return kq_fail(ctx, KQ_ERR_NUMBER_FORMAT, "For input string: cannot parse as double");
    if (e & KQ_DEV_ERR_LONG_KEY) return kq_fail(ctx, KQ_ERR_UNSUPPORTED, "Utf8 group key longer than 7 bytes in a kernel without a key heap");
    return kq_fail(ctx, KQ_ERR_CUDA, "unknown device error bits 0x%x", e);
}

int kq_dev_alloc(kq_ctx* ctx, size_t bytes, void** out) {
    size_t padded = ((bytes + 255) / 256) * 256 + KQ_PAD;
    // size classes: round up to 1/8 of the leading power of two so freed blocks get reused
    size_t cls = 4096;
    while (cls < padded) cls <<= 1;
    size_t step = cls / 8;
    size_t want = step ? ((padded + step - 1) / step) * step : padded;
    void* p = nullptr;
    auto it = ctx->free_blocks.lower_bound(want);
    if (it != ctx->free_blocks.end() && it->first <= want + want / 4) {
        p = it->second;
        ctx->bytes_cached -= it->first;
        want = it->first;
        ctx->free_blocks.erase(it);
    } else {
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaErrorMemoryAllocation) {
            // give everything cached back to the driver and retry once
            cudaGetLastError();
            cudaStreamSynchronize(ctx->stream);
            for (auto& kv : ctx->free_blocks) { ctx->block_size.erase(kv.second); cudaFree(kv.second); }
            ctx->free_blocks.clear(); ctx->bytes_cached = 0;
            e = cudaMalloc(&p, want);
        }
        if (e != cudaSuccess) return kq_cuda_fail(ctx, e, "cudaMalloc");
        ctx->block_size[p] = want;
    }
    ctx->bytes_live += want;
    *out = p;
    // zero the slack so that whole-vector tail reads see defined bytes
    size_t tail_from = (bytes / 256) * 256;
    KQ_CUDA(ctx, cudaMemsetAsync((char*)p + tail_from, 0, padded - tail_from, ctx->stream));
    return KQ_OK;
}
void kq_dev_free(kq_ctx* ctx, void* p) {
    if (!p) return;
    auto it = ctx->block_size.find(p);
    if (it == ctx->block_size.end()) { cudaFree(p); return; }
    ctx->free_blocks.emplace(it->second, p);
    ctx->bytes_cached += it->second;
    ctx->bytes_live -= it->second;
}

int kq_type_width(int type) {
    switch (type) {
        case KQ_F64: case KQ_I64: return 8;
        case KQ_DATE32: case KQ_I32: return 4;
        default: return 0;
    }
}

static size_t kq_data_bytes_for(int type, int64_t n, int64_t utf8_bytes) {
    if (type == KQ_UTF8) return (size_t)utf8_bytes;
    if (type == KQ_BOOL) return (size_t)((n + 63) / 64) * 8;
    return (size_t)n * kq_type_width(type);
}

int kq_col_new(kq_ctx* ctx, int type, int64_t n, bool with_validity, int64_t utf8_bytes, kq_col** out) {
    kq_col* c = new kq_col();
    c->ctx = ctx; c->type = type; c->n = n; c->capacity_rows = n;
    c->data_bytes = type == KQ_UTF8 ? utf8_bytes : 0;
    int st = kq_dev_alloc(ctx, kq_data_bytes_for(type, n, utf8_bytes), &c->data);
    if (st == KQ_OK && with_validity) st = kq_dev_alloc(ctx, (size_t)((n + 63) / 64) * 8, (void**)&c->validity);
    if (st == KQ_OK && type == KQ_UTF8) st = kq_dev_alloc(ctx, (size_t)(n + 1) * 4, (void**)&c->offsets);
    if (st != KQ_OK) { kq_column_free(c); return st; }
    *out = c;
    return KQ_OK;
}

// Slots whose owner went away before the asynchronous count copy landed are parked here and recycled
// once their event has fired, so dropping an unread filter result never blocks the host.
static void lazy_recycle(kq_ctx* ctx, kq_lazy_count* l) {
    ctx->free_events.push_back(l->ev);
    ctx->free_slots.push_back(l->slot);
    kq_dev_free(ctx, l->d_slot);
    delete l;
}
static void lazy_reap(kq_ctx* ctx, bool wait_for_one) {
    for (size_t i = 0; i < ctx->pending_lazy.size();) {
        kq_lazy_count* l = ctx->pending_lazy[i];
        if (cudaEventQuery(l->ev) == cudaSuccess) { lazy_recycle(ctx, l); ctx->pending_lazy.erase(ctx->pending_lazy.begin() + (long)i); }
        else i++;
    }
    if (wait_for_one && ctx->free_slots.empty() && !ctx->pending_lazy.empty()) {
        kq_lazy_count* l = ctx->pending_lazy.front();
        cudaEventSynchronize(l->ev);
        lazy_recycle(ctx, l);
        ctx->pending_lazy.erase(ctx->pending_lazy.begin());
    }
}

kq_lazy_count* kq_lazy_new(kq_ctx* ctx) {
    if (!ctx->pending_lazy.empty()) lazy_reap(ctx, true);
    kq_lazy_count* l = new kq_lazy_count();
    if (ctx->free_slots.empty()) { delete l; return nullptr; }
    l->slot = ctx->free_slots.back(); ctx->free_slots.pop_back();
    l->h_slot = ctx->pinned_slots + 2 * l->slot;
    if (kq_dev_alloc(ctx, 16, (void**)&l->d_slot) != KQ_OK) { ctx->free_slots.push_back(l->slot); delete l; return nullptr; }
    cudaMemsetAsync(l->d_slot, 0, 16, ctx->stream);
    if (!ctx->free_events.empty()) { l->ev = ctx->free_events.back(); ctx->free_events.pop_back(); }
    else cudaEventCreateWithFlags(&l->ev, cudaEventDisableTiming);
    return l;
}
void kq_lazy_release(kq_ctx* ctx, kq_lazy_count* l) {
    if (!l) return;
    if (l->rc.fetch_sub(1) == 1) {
        // the async D2H into the pinned slot must have landed before the slot is reused
        if (!l->resolved && cudaEventQuery(l->ev) != cudaSuccess) ctx->pending_lazy.push_back(l);
        else lazy_recycle(ctx, l);
    }
}
static int kq_lazy_resolve(kq_ctx* ctx, kq_lazy_count* l, int64_t* n) {
    if (!l->resolved) {
        KQ_CUDA(ctx, cudaEventSynchronize(l->ev));
        l->value = (int64_t)l->h_slot[0];
        l->resolved = true;
    }
    *n = l->value;
    return KQ_OK;
}
int kq_col_resolve_rows(kq_ctx* ctx, kq_col* c, int64_t* n) {
    if (c->n < 0) {
        int64_t v; KQ_RET(kq_lazy_resolve(ctx, c->lazy, &v));
        c->n = v;
    }
    if (c->type == KQ_UTF8 && c->data_bytes < 0) {
        uint64_t b; KQ_RET(kq_read_u64(ctx, c->d_utf8_bytes, 1, &b));
        c->data_bytes = (int64_t)b;
    }
    if (n) *n = c->n;
    return KQ_OK;
}
int kq_batch_resolve_rows(kq_ctx* ctx, kq_batch* b, int64_t* n) {
    if (b->n < 0) {
        int64_t v; KQ_RET(kq_lazy_resolve(ctx, b->lazy, &v));
        b->n = v;
        for (kq_col* c : b->cols) if (c->n < 0) c->n = v;
    }
    if (n) *n = b->n;
    return KQ_OK;
}

int kq_read_u64(kq_ctx* ctx, const void* d_ptr, int count, uint64_t* out) {
    if (count > 64) return kq_fail(ctx, KQ_ERR_ILLEGAL_ARGUMENT, "kq_read_u64: count too large");
    KQ_CUDA(ctx, cudaMemcpyAsync(ctx->h_scratch, d_ptr, (size_t)count * 8, cudaMemcpyDeviceToHost, ctx->stream));
    KQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < count; i++) out[i] = ctx->h_scratch[i];
    return KQ_OK;
}

// ---- kernels used by column plumbing -----------------------------------------------------------------------
__global__ void k_popcount_valid(const uint32_t* __restrict__ bits, int64_t n, unsigned long long* out) {
    int64_t nwords = (n + 31) / 32;
    unsigned long long local = 0;
    for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < nwords; w += (int64_t)gridDim.x * blockDim.x) {
        uint32_t x = bits[w];
        if (w == nwords - 1 && (n & 31)) x &= (1u << (n & 31)) - 1u;
        local += __popc(x);
    }
    for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, local);
}
__global__ void k_fill_ones(uint32_t* bits, int64_t nwords) {
    for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < nwords; w += (int64_t)gridDim.x * blockDim.x) bits[w] = 0xffffffffu;
}
__global__ void k_rebase_offsets(int32_t* off, int64_t n1, int32_t base) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n1; i += (int64_t)gridDim.x * blockDim.x) off[i] -= base;
}

static int grid_for(kq_ctx* ctx, int64_t items, int block) {
    int64_t g = (items + block - 1) / block;
    int64_t cap = (int64_t)ctx->sm_count * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

extern "C" {

// ---- columns ----------------------------------------------------------------------------------------------------
int kq_column_upload(kq_ctx* ctx, int type, int64_t n, const uint8_t* validity, const int32_t* offsets,
                     const void* data, int64_t data_bytes, kq_col** out) {
    if (!ctx || !out) return KQ_ERR_ILLEGAL_ARGUMENT;
    if (type < KQ_F64 || type > KQ_I32) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "unknown column type %d", type);   // Main.kt:195
    if (n < 0 || n > 2147483647LL) return kq_fail(ctx, KQ_ERR_ILLEGAL_ARGUMENT, "row count %lld out of range", (long long)n);
    if (type == KQ_UTF8 && n > 0 && !offsets) return kq_fail(ctx, KQ_ERR_ILLEGAL_ARGUMENT, "UTF8 column without offsets");
    cudaSetDevice(ctx->device);
    int64_t payload_off = 0, payload = 0;
    if (type == KQ_UTF8) {
        if (n > 0) { payload_off = offsets[0]; payload = (int64_t)offsets[n] - offsets[0]; }
        if (payload < 0 || (data_bytes > 0 && payload_off + payload > data_bytes))
            return kq_fail(ctx, KQ_ERR_ILLEGAL_ARGUMENT, "UTF8 offsets exceed data buffer");
    }
    kq_col* c = nullptr;
    KQ_RET(kq_col_new(ctx, type, n, validity != nullptr, payload, &c));
    // allocations were stream-ordered on the compute stream; make the copy stream wait for them
    cudaStream_t cs = ctx->copy_stream[0];
    cudaError_t e = cudaEventRecord(ctx->copy_done, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(cs, ctx->copy_done, 0);
    size_t bytes = type == KQ_UTF8 ? (size_t)payload : (type == KQ_BOOL ? (size_t)((n + 7) / 8) : (size_t)n * kq_type_width(type));
    if (e == cudaSuccess && bytes)
        e = cudaMemcpyAsync(c->data, (const char*)data + payload_off, bytes, cudaMemcpyHostToDevice, cs);
    if (e == cudaSuccess && validity && n)
        e = cudaMemcpyAsync(c->validity, validity, (size_t)((n + 7) / 8), cudaMemcpyHostToDevice, cs);
    if (e == cudaSuccess && type == KQ_UTF8) {
        if (n > 0) e = cudaMemcpyAsync(c->offsets, offsets, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, cs);
        else e = cudaMemsetAsync(c->offsets, 0, 4, cs);
    }
    // the caller's buffers must be reusable on return (they are pageable or pinned Arrow memory the
    // library does not own): wait for the copies, then order the compute stream after them.
    if (e == cudaSuccess) e = cudaEventRecord(ctx->copy_done, cs);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->copy_done, 0);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cs);
    if (e != cudaSuccess) { kq_column_free(c); return kq_cuda_fail(ctx, e, "column upload"); }
    if (type == KQ_UTF8 && n > 0 && payload_off != 0) {
        k_rebase_offsets<<<grid_for(ctx, n + 1, 256), 256, 0, ctx->stream>>>(c->offsets, n + 1, (int32_t)payload_off);
        ctx->launches++;
    }
    *out = c;
    return KQ_OK;
}

int kq_column_sizes(kq_ctx* ctx, kq_col* col, int64_t* n, int64_t* data_bytes, int64_t* null_count) {
    if (!ctx || !col) return KQ_ERR_ILLEGAL_ARGUMENT;
    cudaSetDevice(ctx->device);
    int64_t rows; KQ_RET(kq_col_resolve_rows(ctx, col, &rows));
    if (n) *n = rows;
    if (data_bytes) {
        if (col->type == KQ_UTF8) *data_bytes = col->data_bytes;
        else if (col->type == KQ_BOOL) *data_bytes = (rows + 7) / 8;
        else *data_bytes = rows * kq_type_width(col->type);
    }
    if (null_count) {
        if (!col->validity || rows == 0) *null_count = 0;
        else {
            unsigned long long* d;
            KQ_RET(kq_dev_alloc(ctx, 8, (void**)&d));
            KQ_CUDA(ctx, cudaMemsetAsync(d, 0, 8, ctx->stream));
            k_popcount_valid<<<grid_for(ctx, (rows + 31) / 32, 256), 256, 0, ctx->stream>>>(col->validity, rows, d);
            ctx->launches++;
            uint64_t v; KQ_RET(kq_read_u64(ctx, d, 1, &v));
            kq_dev_free(ctx, d);
            *null_count = rows - (int64_t)v;
        }
    }
    return KQ_OK;
}

int kq_column_type(kq_col* col) { return col ? col->type : 0; }

int kq_column_download(kq_ctx* ctx, kq_col* col, uint8_t* validity, int32_t* offsets, void* data) {
    if (!ctx || !col) return KQ_ERR_ILLEGAL_ARGUMENT;
    cudaSetDevice(ctx->device);
    int64_t rows; KQ_RET(kq_col_resolve_rows(ctx, col, &rows));
    size_t bytes = col->type == KQ_UTF8 ? (size_t)col->data_bytes
                 : (col->type == KQ_BOOL ? (size_t)((rows + 7) / 8) : (size_t)rows * kq_type_width(col->type));
    if (data && bytes) KQ_CUDA(ctx, cudaMemcpyAsync(data, col->data, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (validity && rows) {
        if (col->validity) KQ_CUDA(ctx, cudaMemcpyAsync(validity, col->validity, (size_t)((rows + 7) / 8), cudaMemcpyDeviceToHost, ctx->stream));
        else memset(validity, 0xff, (size_t)((rows + 7) / 8));
    }
    if (offsets && col->type == KQ_UTF8) KQ_CUDA(ctx, cudaMemcpyAsync(offsets, col->offsets, (size_t)(rows + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    KQ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return kq_check_device_errors(ctx);
}

int kq_column_device_ptrs(kq_ctx* ctx, kq_col* col, void** validity, void** offsets, void** data) {
    if (!ctx || !col) return KQ_ERR_ILLEGAL_ARGUMENT;
    if (validity) *validity = col->validity;
    if (offsets) *offsets = col->offsets;
    if (data) *data = col->data;
    return KQ_OK;
}

int kq_column_retain(kq_col* col) { if (col) col->rc.fetch_add(1); return KQ_OK; }

int kq_column_free(kq_col* col) {
    if (!col) return KQ_OK;
    if (col->rc.fetch_sub(1) == 1) {
        kq_ctx* ctx = col->ctx;
        cudaSetDevice(ctx->device);
        kq_dev_free(ctx, col->data);
        kq_dev_free(ctx, col->validity);
        kq_dev_free(ctx, col->offsets);
        kq_dev_free(ctx, col->d_utf8_bytes);
        kq_lazy_release(ctx, col->lazy);
        delete col;
    }
    return KQ_OK;
}

// ---- batches ------------------------------------------------------------------------------------------------------
int kq_batch_create(kq_ctx* ctx, kq_col* const* cols, int ncols, int64_t n_rows, kq_batch** out) {
    if (!ctx || !out || ncols < 0) return KQ_ERR_ILLEGAL_ARGUMENT;
    if (ncols == 0 && n_rows < 0) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "rowCount of a batch without columns");   // fields.first() on empty list
    kq_batch* b = new kq_batch();
    b->ctx = ctx;
    int64_t n = n_rows;
    for (int i = 0; i < ncols; i++) {
        int64_t ci;
        int st = kq_col_resolve_rows(ctx, cols[i], &ci);
        if (st == KQ_OK && n >= 0 && ci != n && !(i == 0 && n_rows < 0))
            st = kq_fail(ctx, KQ_ERR_ILLEGAL_ARGUMENT, "column %d has %lld rows, expected %lld", i, (long long)ci, (long long)n);
        if (st != KQ_OK) { kq_batch_free(b); return st; }
        if (n < 0) n = ci;
        cols[i]->rc.fetch_add(1);
        b->cols.push_back(cols[i]);
    }
    b->n = n;
    *out = b;
    return KQ_OK;
}

int kq_batch_num_rows(kq_ctx* ctx, kq_batch* batch, int64_t* n) {
    if (!ctx || !batch || !n) return KQ_ERR_ILLEGAL_ARGUMENT;
    cudaSetDevice(ctx->device);
    return kq_batch_resolve_rows(ctx, batch, n);
}
int kq_batch_num_columns(kq_batch* batch) { return batch ? (int)batch->cols.size() : 0; }

int kq_batch_column(kq_ctx* ctx, kq_batch* batch, int i, kq_col** out) {
    if (!ctx || !batch || !out) return KQ_ERR_ILLEGAL_ARGUMENT;
    if (i < 0 || i >= (int)batch->cols.size()) return kq_fail(ctx, KQ_ERR_ILLEGAL_STATE, "field index %d out of range", i);
    batch->cols[(size_t)i]->rc.fetch_add(1);
    *out = batch->cols[(size_t)i];
    return KQ_OK;
}

int kq_batch_free(kq_batch* b) {
    if (!b) return KQ_OK;
    if (b->rc.fetch_sub(1) == 1) {
        for (kq_col* c : b->cols) kq_column_free(c);
        kq_lazy_release(b->ctx, b->lazy);
        delete b;
    }
    return KQ_OK;
}

}  // extern "C"

// ---- synthetic tables ----------------------------------------------------------------------------------------------
struct GenCol {
    int kind, col_id;
    int64_t ilo, ihi;
    double flo, fhi;
    int null_per_10k, dict_width, dict_count;
    void* data; uint32_t* validity; int32_t* offsets;
    unsigned char dict[128];
};

// One thread generates 64 consecutive rows of one column so that validity / bool words are written whole.
__global__ void __launch_bounds__(256) k_generate(GenCol g, uint64_t seed, int64_t row_begin, int64_t n) {
    int64_t nblk = (n + 63) / 64;
    for (int64_t blk = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; blk < nblk; blk += (int64_t)gridDim.x * blockDim.x) {
        int64_t i0 = blk * 64;
        uint64_t vbits = 0, bbits = 0;
        int cnt = (int)((n - i0) < 64 ? (n - i0) : 64);
        for (int k = 0; k < cnt; k++) {
            int64_t i = i0 + k, row = row_begin + i;
            uint64_t h = kq_gen_hash(seed, g.col_id, row);
            if (g.validity && !kq_gen_is_null(seed, g.col_id, row, g.null_per_10k)) vbits |= 1ULL << k;
            switch (g.kind) {
                case KQ_GEN_I64_UNIFORM: ((int64_t*)g.data)[i] = kq_gen_i64(h, g.ilo, g.ihi); break;
                case KQ_GEN_F64_UNIFORM: ((double*)g.data)[i] = kq_gen_f64_uniform(h, g.flo, g.fhi); break;
                case KQ_GEN_F64_INT: ((double*)g.data)[i] = (double)kq_gen_i64(h, g.ilo, g.ihi); break;
                case KQ_GEN_F64_STEP: ((double*)g.data)[i] = kq_gen_f64_step(h, g.ilo, g.ihi, g.fhi); break;
                case KQ_GEN_UTF8_DICT: {
                    int code = (int)(h % (uint64_t)g.dict_count);
                    for (int b = 0; b < g.dict_width; b++) ((unsigned char*)g.data)[i * g.dict_width + b] = g.dict[code * g.dict_width + b];
                    g.offsets[i] = (int32_t)(i * g.dict_width);
                    if (i == n - 1) g.offsets[n] = (int32_t)(n * g.dict_width);
                    break;
                }
                case KQ_GEN_DATE32_UNIFORM: ((int32_t*)g.data)[i] = (int32_t)kq_gen_i64(h, g.ilo, g.ihi); break;
                case KQ_GEN_BOOL: if ((int64_t)(h % 10000ULL) < g.ilo) bbits |= 1ULL << k; break;
            }
        }
        if (g.validity) ((uint64_t*)g.validity)[blk] = vbits;
        if (g.kind == KQ_GEN_BOOL) ((uint64_t*)g.data)[blk] = bbits;
    }
}

extern "C" int kq_generate(kq_ctx* ctx, const kq_gen_spec* specs, int ncols, uint64_t seed,
                           int64_t row_begin, int64_t row_end, kq_batch** out) {
    if (!ctx || !out || ncols < 0 || row_end < row_begin) return KQ_ERR_ILLEGAL_ARGUMENT;
    cudaSetDevice(ctx->device);
    int64_t n = row_end - row_begin;
    if (n > 2147483647LL) return kq_fail(ctx, KQ_ERR_ILLEGAL_ARGUMENT, "batch larger than 2^31-1 rows");
    std::vector<kq_col*> cols;
    int st = KQ_OK;
    for (int i = 0; i < ncols && st == KQ_OK; i++) {
        const kq_gen_spec& s = specs[i];
        int type;
        switch (s.kind) {
            case KQ_GEN_I64_UNIFORM: type = KQ_I64; break;
            case KQ_GEN_F64_UNIFORM: case KQ_GEN_F64_INT: case KQ_GEN_F64_STEP: type = KQ_F64; break;
            case KQ_GEN_UTF8_DICT: type = KQ_UTF8; break;
            case KQ_GEN_DATE32_UNIFORM: type = KQ_DATE32; break;
            case KQ_GEN_BOOL: type = KQ_BOOL; break;
            default: st = kq_fail(ctx, KQ_ERR_ILLEGAL_ARGUMENT, "unknown generator kind %d", s.kind); continue;
        }
        if (s.kind != KQ_GEN_BOOL && s.kind != KQ_GEN_F64_UNIFORM && s.ihi <= s.ilo) { st = kq_fail(ctx, KQ_ERR_ILLEGAL_ARGUMENT, "empty integer range"); continue; }
        GenCol g{};
        g.kind = s.kind; g.col_id = s.col_id; g.ilo = s.ilo; g.ihi = s.ihi; g.flo = s.flo; g.fhi = s.fhi;
        g.null_per_10k = s.null_per_10k; g.dict_width = s.dict_width; g.dict_count = s.dict_count;
        int64_t utf8_bytes = 0;
        if (type == KQ_UTF8) {
            if (!s.dict || s.dict_width <= 0 || s.dict_count <= 0 || (size_t)s.dict_width * s.dict_count > sizeof g.dict) {
                st = kq_fail(ctx, KQ_ERR_ILLEGAL_ARGUMENT, "bad UTF8 dictionary"); continue;
            }
            if (n * s.dict_width > 2147483647LL) { st = kq_fail(ctx, KQ_ERR_ILLEGAL_ARGUMENT, "UTF8 data exceeds int32 offsets"); continue; }
            memcpy(g.dict, s.dict, (size_t)s.dict_width * s.dict_count);
            utf8_bytes = n * s.dict_width;
        }
        kq_col* c = nullptr;
        st = kq_col_new(ctx, type, n, s.null_per_10k > 0, utf8_bytes, &c);
        if (st != KQ_OK) continue;
        cols.push_back(c);
        g.data = c->data; g.validity = c->validity; g.offsets = c->offsets;
        if (n > 0) {
            k_generate<<<grid_for(ctx, (n + 63) / 64, 256), 256, 0, ctx->stream>>>(g, seed, row_begin, n);
            ctx->launches++;
        } else if (type == KQ_UTF8) cudaMemsetAsync(c->offsets, 0, 4, ctx->stream);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) st = kq_cuda_fail(ctx, e, "k_generate");
    }
    if (st == KQ_OK) st = kq_batch_create(ctx, cols.data(), (int)cols.size(), n, out);
    for (kq_col* c : cols) kq_column_free(c);
    return st;
}
