// Host stand-in for csrc/kq_scan.cuh (see cuda_runtime.h in this directory): the device-wide look-back scan is replaced
// by a sequential exclusive sum with the same contract — out_off[i] = sum of f(j) for j < i, out_off[m] = *out_bytes = the
// total, tickets and tile descriptors untouched. The functors it is instantiated with are the code under test.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace kq {
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = 256 * SCAN_ITEMS;
template <class LenFn>
void k_exclusive_offsets(LenFn f, const unsigned long long* d_count, int32_t* out_off, unsigned long long* tile_desc, unsigned int* ticket,
                         unsigned long long* out_bytes) {
    (void)tile_desc; (void)ticket;
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const long long m = (long long)*d_count;
    long long run = 0;
    for (long long i = 0; i < m; i++) { out_off[i] = (int32_t)run; run += f(i); }
    out_off[m] = (int32_t)run;
    if (out_bytes) *out_bytes = (unsigned long long)run;
}
}  // namespace kq
