// kq_parse.cuh — String.toDouble() (Main.kt:791, = Double.parseDouble) for EVERY input the Java grammar accepts,
// correctly rounded (round-half-even), in plain integer arithmetic so that it runs inside the fused expression kernel.
//
// Two tiers (kq_rt.cuh parse_f64 picks):
//   * fast: <= 19 significant digits, |decimal exponent| <= 22, mantissa <= 2^53 — one exact IEEE multiply or divide
//     (Clinger); nearly every CSV cell takes it;
//   * exact (this file): the decimal is turned into a big integer (up to 768 significant digits — no double's rounding
//     boundary has more — plus a sticky flag for what was dropped), scaled by the power of ten with exact multi-word
//     arithmetic (multiply by 10^9 chunks, or shift left and divide by 5^13 chunks while the remainder feeds the sticky
//     flag), and the leading 64 bits + sticky are rounded once. Hex floats ("0x1.8p3") take the same final rounding.
// The functions are __host__ __device__ and use nothing but integers: tests/test_parse_cpu.py compiles this header with
// g++ and checks a million random strings bit for bit against the CPU oracle without a GPU.
#pragma once

#ifndef __CUDACC_RTC__
#include <stdint.h>
#endif

#ifndef KQ_HD
#ifdef __CUDACC__
#define KQ_HD __host__ __device__
#else
#define KQ_HD
#endif
#endif

namespace kq {

constexpr int PARSE_LIMBS = 120;          // 3840 bits: 768 digits (2552 bits) x 10^310, or 5^1100 + 66 bits of quotient
constexpr int PARSE_MAX_DIGITS = 768;

struct BigNat {
    uint32_t w[PARSE_LIMBS];
    int n;                                // limbs in use (w[n-1] != 0), 0 = zero
};

KQ_HD inline void big_mul_add(BigNat& b, uint32_t m, uint32_t a) {          // b = b * m + a
    uint64_t carry = a;
    for (int i = 0; i < b.n; i++) {
        const uint64_t t = (uint64_t)b.w[i] * m + carry;
        b.w[i] = (uint32_t)t; carry = t >> 32;
    }
    if (carry && b.n < PARSE_LIMBS) b.w[b.n++] = (uint32_t)carry;
}
KQ_HD inline uint32_t big_div_small(BigNat& b, uint32_t d) {                 // b = b / d, returns the remainder
    uint64_t rem = 0;
    for (int i = b.n - 1; i >= 0; i--) {
        const uint64_t t = (rem << 32) | b.w[i];
        b.w[i] = (uint32_t)(t / d); rem = t % d;
    }
    while (b.n > 0 && b.w[b.n - 1] == 0) b.n--;
    return (uint32_t)rem;
}
KQ_HD inline int clz32_(uint32_t x) { int n = 0; if (!x) return 32; while (!(x & 0x80000000u)) { x <<= 1; n++; } return n; }
KQ_HD inline int big_bitlen(const BigNat& b) { return b.n == 0 ? 0 : 32 * b.n - clz32_(b.w[b.n - 1]); }
KQ_HD inline void big_shl(BigNat& b, int s) {
    if (b.n == 0 || s <= 0) return;
    const int ws = s >> 5, bs = s & 31;
    int nn = b.n + ws + 1;
    if (nn > PARSE_LIMBS) nn = PARSE_LIMBS;
    for (int i = nn - 1; i >= 0; i--) {
        const int j = i - ws;
        uint32_t lo = (j >= 0 && j < b.n) ? b.w[j] : 0u, lo1 = (j - 1 >= 0 && j - 1 < b.n) ? b.w[j - 1] : 0u;
        b.w[i] = bs ? (lo << bs) | (lo1 >> (32 - bs)) : lo;
    }
    b.n = nn;
    while (b.n > 0 && b.w[b.n - 1] == 0) b.n--;
}

// Round (top: the value's leading 64 bits, MSB at bit 63; eb: the unbiased binary exponent of that MSB; sticky: anything
// nonzero below the 64 bits) to the nearest double, ties to even; overflow -> +inf, underflow -> subnormals / 0.
KQ_HD inline uint64_t round_to_f64_bits(uint64_t top, int eb, bool sticky) {
    if (top == 0) return 0;
    if (eb > 1023) return 0x7FF0000000000000ULL;
    int p = 53;                                          // significand bits that fit
    if (eb < -1022) p = eb + 1075;                       // subnormal: fewer
    if (p < 0) return 0;
    uint64_t mant = p > 0 ? top >> (64 - p) : 0;
    const bool guard = (top >> (63 - p)) & 1u;
    const bool rest = (top & ((1ULL << (63 - p)) - 1ULL)) != 0 || sticky;
    if (guard && (rest || (mant & 1u))) mant++;
    if (eb >= -1022) {
        if (mant == (1ULL << 53)) { mant >>= 1; eb++; }
        if (eb > 1023) return 0x7FF0000000000000ULL;
        return ((uint64_t)(eb + 1023) << 52) | (mant & ((1ULL << 52) - 1ULL));
    }
    return mant;                                          // subnormal (a carry into bit 52 is the smallest normal: same encoding)
}
KQ_HD inline uint64_t big_round(const BigNat& b, int e2, bool sticky) {     // value = b * 2^e2
    const int L = big_bitlen(b);
    if (L == 0) return 0;
    // leading 64 bits
    uint64_t top = 0;
    const int hi = L - 1;                                 // index of the MSB
    for (int k = 0; k < 64; k++) {
        const int bit = hi - k;
        if (bit < 0) break;
        top |= (uint64_t)((b.w[bit >> 5] >> (bit & 31)) & 1u) << (63 - k);
    }
    if (L > 64) {
        const int low = L - 64;                           // bits [0, low) lie below
        for (int i = 0; i < (low >> 5) && !sticky; i++) sticky |= b.w[i] != 0;
        if (!sticky && (low & 31)) sticky |= (b.w[low >> 5] & ((1u << (low & 31)) - 1u)) != 0;
    }
    return round_to_f64_bits(top, hi + e2, sticky);
}

// digits: the characters [p, p + n) hold decimal digits and at most one '.', already validated; exp10: the explicit
// exponent. Returns the IEEE bits of the correctly rounded magnitude.
KQ_HD inline uint64_t decimal_to_f64_bits(const uint8_t* p, int n, long long exp10) {
    BigNat b; b.n = 0;
    int used = 0;                      // significant digits taken into b
    long long point_pos = 0;           // significant digits (incl. dropped ones) before the decimal point; negative: zeros after the point
    bool started = false, seen_point = false, sticky = false;
    uint32_t chunk = 0; int cd = 0;
    for (int i = 0; i < n; i++) {
        const uint8_t c = p[i];
        if (c == '.') { seen_point = true; continue; }
        const uint32_t d = (uint32_t)(c - '0');
        if (!started) {
            if (d == 0) { if (seen_point) point_pos--; continue; }
            started = true;
        }
        if (!seen_point) point_pos++;
        if (used < PARSE_MAX_DIGITS) {
            chunk = chunk * 10u + d; cd++; used++;
            if (cd == 9) { big_mul_add(b, 1000000000u, chunk); chunk = 0; cd = 0; }
        } else sticky |= d != 0;
    }
    if (!started) return 0;
    if (cd) { uint32_t m = 1; for (int i = 0; i < cd; i++) m *= 10u; big_mul_add(b, m, chunk); }
    // value = 0.d1d2... x 10^(point_pos + exp10): quick range cuts keep every loop below bounded
    if (exp10 > 100000) exp10 = 100000;
    if (exp10 < -100000) exp10 = -100000;
    const long long dexp = point_pos + exp10;
    if (dexp > 310) return 0x7FF0000000000000ULL;
    if (dexp < -326) return 0;
    long long e10 = dexp - used;       // value = b x 10^e10 (+ sticky)
    if (e10 >= 0) {
        while (e10 >= 9) { big_mul_add(b, 1000000000u, 0); e10 -= 9; }
        if (e10) { uint32_t m = 1; for (int i = 0; i < (int)e10; i++) m *= 10u; big_mul_add(b, m, 0); }
        return big_round(b, 0, sticky);
    }
    // b x 10^-m = (b x 2^s / 5^m) x 2^(-s - m): shift so that the quotient keeps >= 66 significant bits
    const int m = (int)(-e10);
    const int need = (int)((m * 2322LL + 999) / 1000) + 66;            // bit length of 5^m < 2.322 m
    int s = need - big_bitlen(b);
    if (s < 0) s = 0;
    big_shl(b, s);
    int left = m;
    while (left >= 13) { sticky |= big_div_small(b, 1220703125u) != 0; left -= 13; }
    if (left) { uint32_t d = 1; for (int i = 0; i < left; i++) d *= 5u; sticky |= big_div_small(b, d) != 0; }
    return big_round(b, -s - m, sticky);
}

KQ_HD inline int hex_digit_(uint8_t c) {
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    return -1;
}
// hex significand [p, p + n): hex digits with at most one '.', validated; exp2: the 'p' exponent
KQ_HD inline uint64_t hex_to_f64_bits(const uint8_t* p, int n, long long exp2) {
    uint64_t m = 0; bool sticky = false, seen_point = false;
    long long e = 0;                   // value = m x 2^e (before exp2)
    for (int i = 0; i < n; i++) {
        const uint8_t c = p[i];
        if (c == '.') { seen_point = true; continue; }
        const uint64_t d = (uint64_t)hex_digit_(c);
        if (m >> 60) { sticky |= d != 0; if (!seen_point) e += 4; }     // no room for another digit: drop it
        else { m = (m << 4) | d; if (seen_point) e -= 4; }
    }
    if (m == 0) return 0;
    if (exp2 > 100000) exp2 = 100000;
    if (exp2 < -100000) exp2 = -100000;
    int lz = 0; while (!(m >> 63)) { m <<= 1; lz++; }
    const long long eb = 63 - lz + e + exp2;
    if (eb > 2000) return 0x7FF0000000000000ULL;
    if (eb < -2000) return 0;
    return round_to_f64_bits(m, (int)eb, sticky);
}

// Double.parseDouble (FloatingDecimal.readJavaFormatString) on [p, p + n): 0 = ok, 1 = NumberFormatException.
// *bits receives the IEEE bits. `fast_only` callers are told 2 instead of running the exact tier.
KQ_HD inline int parse_java_double(const uint8_t* p, int n, uint64_t* bits) {
    int b = 0, e = n;
    while (b < e && p[b] <= ' ') b++;
    while (e > b && p[e - 1] <= ' ') e--;
    if (b >= e) return 1;
    uint64_t sign = 0;
    if (p[b] == '+' || p[b] == '-') { sign = p[b] == '-' ? 0x8000000000000000ULL : 0; b++; }
    int len = e - b;
    if (len <= 0) return 1;
    if (len == 3 && p[b] == 'N' && p[b + 1] == 'a' && p[b + 2] == 'N') { *bits = 0x7ff8000000000000ULL; return 0; }
    if (len == 8 && p[b] == 'I' && p[b + 1] == 'n' && p[b + 2] == 'f' && p[b + 3] == 'i' && p[b + 4] == 'n' &&
        p[b + 5] == 'i' && p[b + 6] == 't' && p[b + 7] == 'y') { *bits = sign | 0x7ff0000000000000ULL; return 0; }
    bool hex = len > 2 && p[b] == '0' && (p[b + 1] == 'x' || p[b + 1] == 'X');
    // one optional trailing d/D/f/F (in a hex literal they are digits unless they follow the p-exponent)
    const uint8_t last = p[e - 1];
    if (last == 'd' || last == 'D' || last == 'f' || last == 'F') {
        bool has_p = false;
        if (hex) for (int i = b; i < e; i++) has_p |= p[i] == 'p' || p[i] == 'P';
        if (!hex || has_p) { e--; if (e <= b) return 1; }
        hex = (e - b) > 2 && p[b] == '0' && (p[b + 1] == 'x' || p[b + 1] == 'X');
    }
    int i = hex ? b + 2 : b;
    const int dig0 = i;
    int nd = 0;
    bool seen_point = false;
    for (; i < e; i++) {
        const uint8_t c = p[i];
        if (c == '.') { if (seen_point) break; seen_point = true; continue; }
        if (hex ? hex_digit_(c) < 0 : (c < '0' || c > '9')) break;
        nd++;
    }
    if (nd == 0) return 1;
    const int dig1 = i;
    long long ex = 0; bool has_exp = false;
    if (i < e && (hex ? (p[i] == 'p' || p[i] == 'P') : (p[i] == 'e' || p[i] == 'E'))) {
        has_exp = true; i++;
        bool eneg = false;
        if (i < e && (p[i] == '+' || p[i] == '-')) { eneg = p[i] == '-'; i++; }
        int ed = 0;
        for (; i < e && p[i] >= '0' && p[i] <= '9'; i++) { ed++; if (ex < 1000000) ex = ex * 10 + (p[i] - '0'); }
        if (ed == 0) return 1;
        if (eneg) ex = -ex;
    }
    if (i != e) return 1;
    if (hex && !has_exp) return 1;
    *bits = sign | (hex ? hex_to_f64_bits(p + dig0, dig1 - dig0, ex) : decimal_to_f64_bits(p + dig0, dig1 - dig0, ex));
    return 0;
}

}  // namespace kq
