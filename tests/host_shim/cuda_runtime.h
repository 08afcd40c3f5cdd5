// A stand-in for <cuda_runtime.h> that lets the HOST compiler build csrc/kq_csv.cu for the CPU test-suite
// (tests/test_csv_host.py): kernels become plain functions run once per (block, thread) in sequence, "device" memory is
// host memory, streams and events do nothing (every call completes before it returns). It exists so that the host
// orchestration of the CSV scan and reader (piece cutting, carries, projections, error paths) and the kernels' index
// arithmetic run against the oracle where there is no GPU. Test infrastructure only: nothing in the product includes it,
// and it cannot see stream-ordering mistakes (the GPU tests do).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#define KQ_HOST_SHIM 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __noinline__ /* empty: libstdc++ spells __attribute__((__noinline__)) */

struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct alignas(16) uint4 { unsigned x, y, z, w; };
extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

template <class T> inline T __ldg(const T* p) { return *p; }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
typedef struct kq_shim_stream* cudaStream_t;
typedef struct kq_shim_event* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void* devicePointer; void* hostPointer; };
enum { cudaEventDisableTiming = 2, cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaHostRegisterDefault = 0 };

cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* at, const void* p);      // the harness keeps a registry of "device" buffers
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "host shim"; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t)malloc(1); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }

// kernel<<<grid, block, smem, stream>>>(args) is rewritten to KQ_LAUNCH(kernel, grid, block, args) by tests/test_csv_host.py
#define KQ_LAUNCH(kernel, g, b, ...)                                                    \
    do {                                                                                \
        gridDim = dim3((unsigned)(g)); blockDim = dim3((unsigned)(b));                  \
        for (unsigned _b = 0; _b < gridDim.x; _b++)                                     \
            for (unsigned _t = 0; _t < blockDim.x; _t++) {                              \
                blockIdx.x = _b; threadIdx.x = _t;                                      \
                kernel(__VA_ARGS__);                                                    \
            }                                                                           \
    } while (0)
