// Host build of csrc/kq_parse.cuh (the device's String.toDouble) for tests/test_parse_cpu.py: the same integer-only code
// the GPU runs, compiled with g++ so that its rounding can be checked against the CPU oracle without a device.
#include <stdint.h>
#include "kq_parse.cuh"
extern "C" int kq_test_parse(const char* s, int n, uint64_t* bits) { return kq::parse_java_double((const uint8_t*)s, n, bits); }
