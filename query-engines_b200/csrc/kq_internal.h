// kq_internal.h — host-side structures behind the opaque handles of include/kqgpu.h.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/kqgpu.h"

#define KQ_HIDDEN __attribute__((visibility("hidden")))

// ---- device buffers -------------------------------------------------------------------------------
// Every device buffer is 256-byte aligned and padded by KQ_PAD bytes of slack so that kernels may
// read or write whole 16-byte vectors / 64-row validity words past the logical end.
constexpr size_t KQ_PAD = 512;

struct kq_ctx {
    int device = 0;
    int sm_count = 148;
    int max_smem_optin = 0;
    cudaStream_t stream = nullptr;       // compute stream: every kernel of this ctx
    cudaStream_t copy_stream[2] = {nullptr, nullptr};  // H2D / D2H side streams
    cudaEvent_t copy_done = nullptr;     // orders side-stream uploads before compute
    cudaEvent_t timer[2] = {nullptr, nullptr};
    std::string last_error;
    int64_t launches = 0;
    uint32_t* d_err = nullptr;           // device status word (deferred kernel errors)
    uint32_t* h_err = nullptr;           // pinned mirror
    void* d_flush = nullptr;             // L2 flush scratch
    size_t flush_bytes = 0;
    // pinned staging for small D2H reads (counts)
    uint64_t* h_scratch = nullptr;       // 64 x u64 pinned
    void* comm = nullptr;                // ncclComm_t when kq_comm_init was called
    int rank = 0, nranks = 1;
    // Caching device allocator. Every use of a block is ordered on `stream` (uploads on the copy
    // stream wait for an event recorded on `stream` first), so a freed block can be handed out again
    // immediately: stream order protects it. Blocks go back to the driver only at kq_ctx_destroy.
    std::multimap<size_t, void*> free_blocks;
    std::unordered_map<void*, size_t> block_size;
    size_t bytes_cached = 0, bytes_live = 0;
    // pool of pinned 16-byte slots + events for lazily resolved row counts
    uint64_t* pinned_slots = nullptr;
    std::vector<int> free_slots;
    std::vector<cudaEvent_t> free_events;
    std::vector<struct kq_lazy_count*> pending_lazy;   // released before their count arrived (kq_lazy_release)
};

struct kq_lazy_count;

struct kq_col {
    std::atomic<int> rc{1};
    kq_ctx* ctx = nullptr;
    int type = 0;
    int64_t n = 0;                // rows; -1 while lazily unknown (output of a filter)
    uint32_t* validity = nullptr; // device bitmap, NULL => all valid
    int32_t* offsets = nullptr;   // UTF8: n+1 (device)
    void* data = nullptr;         // device
    int64_t data_bytes = 0;       // UTF8 payload bytes (-1 while lazily unknown)
    int64_t capacity_rows = 0;    // rows the buffers can hold
    kq_lazy_count* lazy = nullptr; // shared, refcounted: resolves n (and nothing else)
    unsigned long long* d_utf8_bytes = nullptr;  // device slot holding data_bytes while lazy
    // Utf8 length statistics, computed once per column on first use as a group key (kq_hashagg.cu): longest string, and how many
    // strings / bytes exceed the 7 bytes a packed key word holds
    int64_t utf8_max_len = -1, utf8_n_long = 0, utf8_bytes_long = 0;
};

// Row count that becomes known when an event fires (stream compaction output).
struct kq_lazy_count {
    std::atomic<int> rc{1};
    uint64_t* h_slot = nullptr;   // pinned (ctx->pinned_slots pool)
    int slot = -1;
    unsigned long long* d_slot = nullptr;
    cudaEvent_t ev = nullptr;
    bool resolved = false;
    int64_t value = -1;
};

struct kq_batch {
    std::atomic<int> rc{1};
    kq_ctx* ctx = nullptr;
    std::vector<kq_col*> cols;
    int64_t n = 0;                 // -1 => lazy
    kq_lazy_count* lazy = nullptr;
};

enum { KQ_EX_COL = 1, KQ_EX_LIT = 2, KQ_EX_BIN = 3, KQ_EX_CAST = 4 };

struct kq_expr {
    std::atomic<int> rc{1};
    int kind = 0;
    int col = -1;
    int type = 0;        // literal type / cast target
    bool is_null = false;
    double f = 0;
    int64_t i = 0;
    std::string s;
    int op = 0;
    kq_expr* l = nullptr;
    kq_expr* r = nullptr;
};

// ---- error plumbing --------------------------------------------------------------------------------
KQ_HIDDEN int kq_fail(kq_ctx* ctx, int code, const char* fmt, ...);
KQ_HIDDEN int kq_cuda_fail(kq_ctx* ctx, cudaError_t e, const char* what);

#define KQ_CUDA(ctx, call)                                              \
    do {                                                                \
        cudaError_t _e = (call);                                        \
        if (_e != cudaSuccess) return kq_cuda_fail((ctx), _e, #call);   \
    } while (0)

#define KQ_RET(call)                     \
    do {                                 \
        int _s = (call);                 \
        if (_s != KQ_OK) return _s;      \
    } while (0)

// device error bits (OR-ed into ctx->d_err by kernels)
enum { KQ_DEV_ERR_DIV0 = 1u, KQ_DEV_ERR_LONG_KEY = 2u, KQ_DEV_ERR_NUMBER_FORMAT = 4u };

// ---- helpers implemented in kq_core.cu ----------------------------------------------------------------
KQ_HIDDEN int kq_dev_alloc(kq_ctx* ctx, size_t bytes, void** out);      // stream-ordered, padded
KQ_HIDDEN void kq_dev_free(kq_ctx* ctx, void* p);
KQ_HIDDEN int kq_col_new(kq_ctx* ctx, int type, int64_t n, bool with_validity, int64_t utf8_bytes, kq_col** out);
KQ_HIDDEN int kq_type_width(int type);   // bytes per row for fixed-width types (0 for UTF8/BOOL)
KQ_HIDDEN int kq_batch_resolve_rows(kq_ctx* ctx, kq_batch* b, int64_t* n);
KQ_HIDDEN int kq_col_resolve_rows(kq_ctx* ctx, kq_col* c, int64_t* n);
KQ_HIDDEN int kq_check_device_errors(kq_ctx* ctx);
KQ_HIDDEN int kq_device_error_status(kq_ctx* ctx, uint32_t bits);   // kq_status for error bits raised by a kernel (0 -> KQ_OK)
KQ_HIDDEN kq_lazy_count* kq_lazy_new(kq_ctx* ctx);
KQ_HIDDEN void kq_lazy_release(kq_ctx* ctx, kq_lazy_count* l);
KQ_HIDDEN int kq_read_u64(kq_ctx* ctx, const void* d_ptr, int count, uint64_t* out);  // sync small D2H
