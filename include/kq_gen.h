/*
 * kq_gen.h — deterministic synthetic-table generator shared by host and device.
 *
 * Every cell is a pure function of (seed, column id, global row index), so a table is
 * identical on the CPU oracle and on the GPU, and identical at 1/2/4/8 GPUs (each rank
 * generates its own contiguous row range [row_begin, row_end) of the same global table).
 * "value = f(splitmix64(seed ^ column_id*phi ^ row))".
 *
 * Bit-exactness host<->device: only integer arithmetic plus IEEE-754 double
 * mul/add/div with round-to-nearest and NO fused multiply-add (the device side uses
 * __dmul_rn/__dadd_rn/__ddiv_rn explicitly; compile host users with -ffp-contract=off).
 *
 * This header is plain C99 / CUDA C++; it has no dependency on the oracle or on the
 * GPU library and is included by both.
 */
#ifndef KQ_GEN_H
#define KQ_GEN_H

#include <stdint.h>

#if defined(__CUDACC__)
#define KQ_HD __host__ __device__ __forceinline__
#else
#define KQ_HD static inline
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* Column generator kinds. */
typedef enum kq_gen_kind {
    KQ_GEN_I64_UNIFORM = 1,  /* int64: ilo + h % (ihi - ilo)                        (ihi > ilo) */
    KQ_GEN_F64_UNIFORM = 2,  /* float64: flo + u * (fhi - flo), u = (h >> 11) * 2^-53 in [0,1)   */
    KQ_GEN_F64_INT     = 3,  /* float64 integer-valued: (double)(ilo + h % (ihi - ilo))          */
    KQ_GEN_F64_STEP    = 4,  /* float64: (double)(ilo + h % (ihi - ilo)) / fhi   (e.g. 0.00..0.10) */
    KQ_GEN_UTF8_DICT   = 5,  /* Utf8: fixed-width code number (h % dict_count) from a dictionary  */
    KQ_GEN_DATE32_UNIFORM = 6, /* date32 (int32 days): ilo + h % (ihi - ilo)                     */
    KQ_GEN_BOOL        = 7   /* bool: (h % 10000) < ilo   (ilo = true-probability in 1/10000)    */
} kq_gen_kind;

/* One column of a synthetic table. `col_id` keys the hash stream, so two specs with the
 * same col_id and seed produce correlated columns on purpose (not used by the configs). */
typedef struct kq_gen_spec {
    int32_t  kind;          /* kq_gen_kind */
    int32_t  col_id;        /* hash stream id */
    int64_t  ilo, ihi;      /* integer range [ilo, ihi) */
    double   flo, fhi;      /* float range / divisor */
    int32_t  null_per_10k;  /* rows with (h2 % 10000) < null_per_10k are null; 0 => no validity buffer */
    int32_t  dict_width;    /* UTF8_DICT: bytes per code */
    int32_t  dict_count;    /* UTF8_DICT: number of codes */
    int32_t  _pad;
    const char* dict;       /* UTF8_DICT: dict_count * dict_width bytes (host pointer) */
} kq_gen_spec;

/* splitmix64 finaliser (Steele, Lea, Flood 2014; public domain constants). */
KQ_HD uint64_t kq_mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

/* Value hash for (seed, column, row). */
KQ_HD uint64_t kq_gen_hash(uint64_t seed, int32_t col_id, int64_t row) {
    uint64_t s = seed ^ ((uint64_t)(uint32_t)(col_id + 1) * 0x9E3779B97F4A7C15ULL);
    return kq_mix64(s + (uint64_t)row * 0xD1B54A32D192ED03ULL);
}

/* Independent hash stream for the null decision. */
KQ_HD uint64_t kq_gen_hash_null(uint64_t seed, int32_t col_id, int64_t row) {
    return kq_mix64(kq_gen_hash(seed, col_id, row) ^ 0xA5A5A5A55A5A5A5AULL);
}

KQ_HD int kq_gen_is_null(uint64_t seed, int32_t col_id, int64_t row, int32_t null_per_10k) {
    if (null_per_10k <= 0) return 0;
    return (int)(kq_gen_hash_null(seed, col_id, row) % 10000ULL) < null_per_10k;
}

KQ_HD int64_t kq_gen_i64(uint64_t h, int64_t ilo, int64_t ihi) {
    return ilo + (int64_t)(h % (uint64_t)(ihi - ilo));
}

KQ_HD double kq_gen_unit(uint64_t h) { /* exact: 53-bit integer times 2^-53 */
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}

KQ_HD double kq_gen_f64_uniform(uint64_t h, double flo, double fhi) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(flo, __dmul_rn(kq_gen_unit(h), __dsub_rn(fhi, flo)));
#else
    volatile double span = fhi - flo;       /* volatile: forbid contraction / x87 excess precision */
    volatile double prod = kq_gen_unit(h) * span;
    return flo + prod;
#endif
}

KQ_HD double kq_gen_f64_step(uint64_t h, int64_t ilo, int64_t ihi, double div) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn((double)kq_gen_i64(h, ilo, ihi), div);
#else
    return (double)kq_gen_i64(h, ilo, ihi) / div;
#endif
}

#ifdef __cplusplus
}
#endif
#endif /* KQ_GEN_H */
