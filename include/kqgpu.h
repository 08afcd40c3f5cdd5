/*
 * kqgpu.h — C ABI of libkqgpu.so, the B200-native execution layer for the KQuery-style
 * physical plan of folkol/query-engines (kquerydiy/src/Main.kt, cited below as Main.kt:N).
 *
 * The reference has no FFI; its operator API is a set of Kotlin interfaces. Each entry point
 * here replaces the body of one of them, so that a Kotlin shim can keep
 * PhysicalPlan / Expression / Accumulator unchanged and forward to this library through
 * Panama FFM or JNI:
 *
 *   kq_column_upload / kq_column_download ... ArrowFieldVector + ColumnVector.getValue (Main.kt:24-27, 176-202)
 *   kq_batch_create ......................... RecordBatch(schema, fields)               (Main.kt:56-61)
 *   kq_expr_column .......................... ColumnExpression                          (Main.kt:452-460)
 *   kq_expr_cast ............................ CastExpression                            (Main.kt:772-805)
 *   kq_expr_literal_*, kq_expr_binary ....... Literal/BinaryExpression — ABSENT from the reference
 *                                             (SURVEY.md §8 a12); defined under rules E1-E8
 *   kq_expr_evaluate ........................ Expression.evaluate(RecordBatch): ColumnVector (Main.kt:448-450)
 *   kq_project .............................. ProjectionExec.execute, one out batch per in batch (Main.kt:589-594)
 *   kq_filter, kq_filter_project ............ FilterExec — ABSENT from the reference (a12)
 *   kq_hashagg_create/update/finalize ....... HashAggregateExec.execute drain loop + single-batch emit
 *                                             (Main.kt:615-651); KQ_AGG_MAX = MaxAccumulator (Main.kt:538-562);
 *                                             SUM/MIN/COUNT are extensions (a12)
 *   kq_comm_*, kq_hashagg_merge_* ........... partition -> partial aggregate -> merge of main() (Main.kt:1306-1342)
 *   kq_csv_header, kq_csv_scan .............. CsvDataSource.schema()/inferSchema and scan() + ReaderIterator.createBatch
 *                                             (Main.kt:251-273, 276-357): CSV text -> one batch of Utf8 columns
 *   kq_csv_reader_open/next/close ........... the same scan as the Sequence<RecordBatch> the reference's ReaderIterator
 *                                             yields (Main.kt:239-249): files of any size, H2D copies under the scan
 *   kq_generate ............................. bench-only synthetic tables (no reference counterpart)
 *
 * Conventions
 *  - Plain pointers and sizes only. Host buffers use the Arrow columnar layout that Arrow Java's
 *    FieldVector exposes (validity bitmap LSB-first, int32 offsets (n+1), raw data bytes).
 *  - Every function returns a kq_status (0 = OK). The message for the last failure on a context is
 *    kq_last_error(ctx). The status codes map 1:1 to the exception classes the reference throws
 *     so the shim can rethrow the same class.
 *  - No global mutable state: one kq_ctx per plan/thread (Main.kt:1309-1313 runs 12 plans
 *    concurrently, each with its own ExecutionContext). A ctx may be used from any thread, one at a time.
 *  - The library never retains or frees caller host pointers past the call that received them (one exception, by
 *    design: an open kq_csv_reader reads the caller's text until kq_csv_reader_close).
 *  - All handles (kq_col, kq_batch, kq_expr, kq_hashagg) are reference counted; *_free drops one reference.
 *  - There is NO CPU fallback: without a CUDA device kq_ctx_create fails with KQ_ERR_NO_DEVICE.
 */
#ifndef KQGPU_H
#define KQGPU_H

#include <stddef.h>
#include <stdint.h>
#include "kq_gen.h"

#ifdef __cplusplus
extern "C" {
#endif

#define KQ_API __attribute__((visibility("default")))

typedef struct kq_ctx kq_ctx;
typedef struct kq_col kq_col;
typedef struct kq_batch kq_batch;
typedef struct kq_expr kq_expr;
typedef struct kq_hashagg kq_hashagg;
typedef struct kq_csv_reader kq_csv_reader;

typedef enum kq_status {
    KQ_OK = 0,
    KQ_ERR_ILLEGAL_STATE = 1,    /* IllegalStateException      (Main.kt:195,469,497,501,677,696,704,792,799) */
    KQ_ERR_UNSUPPORTED = 2,      /* UnsupportedOperationException (Main.kt:548,767) */
    KQ_ERR_ILLEGAL_ARGUMENT = 3, /* IllegalArgumentException   (Main.kt:49) */
    KQ_ERR_SQL = 4,              /* SQLException               (Main.kt:79,666) */
    KQ_ERR_NUMBER_FORMAT = 5,    /* NumberFormatException from String.toDouble() (Main.kt:791) */
    KQ_ERR_ARITHMETIC = 6,       /* ArithmeticException: Int64 division by zero (rule E4) */
    KQ_ERR_OUT_OF_MEMORY = 7,
    KQ_ERR_CUDA = 8,
    KQ_ERR_NCCL = 9,
    KQ_ERR_NO_DEVICE = 10
} kq_status;

/* Column types. F64 and UTF8 are the reference's only types (Main.kt:19-22); the others are
 * extensions (rule E1). Physical layout = Arrow: F64/I64 8 B, DATE32/I32 4 B, BOOL bit-packed. */
typedef enum kq_type {
    KQ_F64 = 1,
    KQ_UTF8 = 2,
    KQ_I64 = 3,
    KQ_BOOL = 4,
    KQ_DATE32 = 5,
    KQ_I32 = 6      /* selection vectors */
} kq_type;

typedef enum kq_binop {
    KQ_EQ = 1, KQ_NE = 2, KQ_LT = 3, KQ_LE = 4, KQ_GT = 5, KQ_GE = 6,
    KQ_AND = 7, KQ_OR = 8,
    KQ_ADD = 9, KQ_SUB = 10, KQ_MUL = 11, KQ_DIV = 12
} kq_binop;

typedef enum kq_aggkind {
    KQ_AGG_MAX = 1,   /* MaxAccumulator, Main.kt:538-562 */
    KQ_AGG_MIN = 2,   /* rule E5 */
    KQ_AGG_SUM = 3,   /* rule E6 */
    KQ_AGG_COUNT = 4  /* rule E7: non-null rows, Int64, never null */
} kq_aggkind;

/* ---- library / context ---------------------------------------------------------------- */
KQ_API const char* kq_version(void);
KQ_API const char* kq_status_name(int status);
KQ_API int kq_device_count(int* count);
KQ_API int kq_ctx_create(int device, kq_ctx** out);
KQ_API int kq_ctx_destroy(kq_ctx* ctx);
KQ_API const char* kq_last_error(kq_ctx* ctx);
/* Wait for all queued work; surfaces deferred device-side errors (e.g. Int64 division by zero). */
KQ_API int kq_ctx_sync(kq_ctx* ctx);
/* The cudaStream_t all kernels of this ctx are launched on (for caller-side CUDA-event timing). */
KQ_API void* kq_ctx_stream(kq_ctx* ctx);
/* Number of kernels this ctx has launched so far (bench.py's gpu_launches). */
KQ_API int64_t kq_ctx_launch_count(kq_ctx* ctx);
/* Device-side timer on the ctx stream: begin/end bracket a region, elapsed in milliseconds. */
KQ_API int kq_timer_begin(kq_ctx* ctx);
KQ_API int kq_timer_end(kq_ctx* ctx, float* elapsed_ms);
/* Writes `bytes` of device scratch (evicts L2 between timed iterations). */
KQ_API int kq_flush_l2(kq_ctx* ctx, size_t bytes);

/* Pinned host memory for H2D staging (Arrow off-heap buffers can also be registered in place). */
KQ_API int kq_host_alloc(kq_ctx* ctx, size_t bytes, void** out);
KQ_API int kq_host_free(kq_ctx* ctx, void* p);
KQ_API int kq_host_register(kq_ctx* ctx, void* p, size_t bytes);
KQ_API int kq_host_unregister(kq_ctx* ctx, void* p);

/* ---- columns and batches (ArrowFieldVector / RecordBatch) -------------------------------- */
/* Stage an Arrow FieldVector triple into a device-resident column (cudaMemcpyAsync on a side
 * stream, ordered before the compute stream). validity == NULL => all rows valid. offsets is
 * used for KQ_UTF8 only ((n+1) int32, offsets[0] may be non-zero); data_bytes is the data-buffer
 * length for KQ_UTF8 and ignored otherwise. */
KQ_API int kq_column_upload(kq_ctx* ctx, int type, int64_t n, const uint8_t* validity,
                            const int32_t* offsets, const void* data, int64_t data_bytes,
                            kq_col** out);
/* n = valueCount (ColumnVector.size, Main.kt:199-201); data_bytes = bytes kq_column_download will
 * write to `data`; null_count = number of null rows. Any out pointer may be NULL. */
KQ_API int kq_column_sizes(kq_ctx* ctx, kq_col* col, int64_t* n, int64_t* data_bytes,
                           int64_t* null_count);
KQ_API int kq_column_type(kq_col* col);
/* Copy back into caller buffers: validity (n+7)/8 bytes (all-ones written when the column has
 * no nulls; may be NULL to skip), offsets (n+1) int32 for UTF8 (rebased to 0), data. */
KQ_API int kq_column_download(kq_ctx* ctx, kq_col* col, uint8_t* validity, int32_t* offsets,
                              void* data);
/* Raw device pointers (validity may come back NULL = all valid) for zero-copy consumers. */
KQ_API int kq_column_device_ptrs(kq_ctx* ctx, kq_col* col, void** validity, void** offsets,
                                 void** data);
KQ_API int kq_column_retain(kq_col* col);
KQ_API int kq_column_free(kq_col* col);

/* RecordBatch(schema, fields): rowCount() = fields.first().size() (Main.kt:57). All columns must
 * have the same length. ncols may be 0 only together with an explicit row count (n_rows >= 0);
 * pass n_rows = -1 to take it from the first column. */
KQ_API int kq_batch_create(kq_ctx* ctx, kq_col* const* cols, int ncols, int64_t n_rows,
                           kq_batch** out);
KQ_API int kq_batch_num_rows(kq_ctx* ctx, kq_batch* batch, int64_t* n);
KQ_API int kq_batch_num_columns(kq_batch* batch);
/* RecordBatch.field(i) (Main.kt:58-60); returns a new reference. */
KQ_API int kq_batch_column(kq_ctx* ctx, kq_batch* batch, int i, kq_col** out);
KQ_API int kq_batch_free(kq_batch* batch);

/* ---- physical expressions (Expression, Main.kt:448-450) ---------------------------------- */
KQ_API kq_expr* kq_expr_column(int i);                         /* ColumnExpression(i) */
KQ_API kq_expr* kq_expr_literal_f64(double v);
KQ_API kq_expr* kq_expr_literal_i64(int64_t v);
KQ_API kq_expr* kq_expr_literal_bool(int v);
KQ_API kq_expr* kq_expr_literal_date32(int32_t days);
KQ_API kq_expr* kq_expr_literal_utf8(const char* bytes, int32_t len);
KQ_API kq_expr* kq_expr_literal_null(int type);
KQ_API kq_expr* kq_expr_binary(int op, kq_expr* l, kq_expr* r); /* retains l and r */
KQ_API kq_expr* kq_expr_cast(kq_expr* e, int type);             /* CastExpression; retains e */
KQ_API void kq_expr_free(kq_expr* e);
/* Expression.evaluate(input): one fused kernel over the whole tree. A bare column reference
 * returns the input column itself (alias, Main.kt:453-455). */
KQ_API int kq_expr_evaluate(kq_ctx* ctx, kq_expr* e, kq_batch* input, kq_col** out);

/* ---- operators --------------------------------------------------------------------------- */
/* ProjectionExec.execute for one batch (Main.kt:589-594): same row count, one column per expr. */
KQ_API int kq_project(kq_ctx* ctx, kq_expr* const* exprs, int nexprs, kq_batch* input,
                      kq_batch** out);
/* FilterExec (extension): keeps rows whose predicate is TRUE (null => dropped, rule E3), in input
 * order, gathering every column. `selection` (optional) receives the KQ_I32 selection vector. */
KQ_API int kq_filter(kq_ctx* ctx, kq_expr* pred, kq_batch* input, kq_batch** out,
                     kq_col** selection);
/* Fused FilterExec + ProjectionExec: single pass, ordered stream compaction. */
KQ_API int kq_filter_project(kq_ctx* ctx, kq_expr* pred, kq_expr* const* exprs, int nexprs,
                             kq_batch* input, kq_batch** out);
/* Same, end to end from HOST Arrow buffers to HOST result buffers: chunks the input, overlaps
 * pinned H2D copies on side streams with compute and D2H. Host columns are described by parallel
 * arrays (type/validity/offsets/data per column). Output buffers (one F64/I64 column per expr)
 * must hold n rows; *out_rows receives the selected row count. */
KQ_API int kq_filter_project_host(kq_ctx* ctx, kq_expr* pred, kq_expr* const* exprs, int nexprs,
                                  int ncols, const int* types, const uint8_t* const* validity,
                                  const void* const* data, int64_t n,
                                  void* const* out_data, uint8_t* const* out_validity,
                                  int64_t* out_rows);

/* Diagnostics (EXPLAIN): the CUDA source of the query-specific part of the kernel kq_project /
 * kq_filter_project would run for this query shape over a batch whose columns have the given types
 * and nullability; with compile != 0 the full kernel is also compiled for sm_100a (no device
 * needed). On failure `source` receives the error text. pred may be NULL (projection only). */
KQ_API int kq_explain_filter_project(kq_expr* pred, kq_expr* const* exprs, int nexprs, int ncols,
                                     const int* types, const int* nullable, int compile,
                                     char* source, size_t source_cap);

/* ---- hash aggregate ---------------------------------------------------------------------- */
/* HashAggregateExec(input, groupExpr, aggregateExpr) (Main.kt:605-610). `pred` (nullable) fuses a
 * FilterExec below the aggregate; group/aggregate input expressions are evaluated in the same
 * kernel (fused ProjectionExec). expected_groups is a sizing hint (0 = unknown). */
KQ_API int kq_hashagg_create(kq_ctx* ctx, kq_expr* pred, kq_expr* const* group_exprs, int ngroup,
                             const int* agg_kinds, kq_expr* const* agg_inputs, int nagg,
                             int64_t expected_groups, kq_hashagg** out);
/* One iteration of the drain loop (Main.kt:617-634): state persists across batches (rule R11). Group keys may be Utf8
 * strings of any length (Main.kt:621-627). Errors raised by the aggregate's kernels are sticky: every later call on the
 * aggregate reports them again. */
KQ_API int kq_hashagg_update(kq_ctx* ctx, kq_hashagg* agg, kq_batch* input);
/* Emit the single output batch (Main.kt:635-650): group columns then aggregate columns. Row
 * order is unspecified (the reference's is HashMap iteration order, rule R10). */
KQ_API int kq_hashagg_finalize(kq_ctx* ctx, kq_hashagg* agg, kq_batch** out);
KQ_API int kq_hashagg_num_groups(kq_ctx* ctx, kq_hashagg* agg, int64_t* n);
KQ_API int kq_hashagg_free(kq_hashagg* agg);

/* EXPLAIN for the aggregate kernel (see kq_explain_filter_project). */
KQ_API int kq_explain_hashagg(kq_expr* pred, kq_expr* const* group_exprs, int ngroup,
                              const int* agg_kinds, kq_expr* const* agg_inputs, int nagg,
                              int ncols, const int* types, const int* nullable, int compile,
                              char* source, size_t source_cap);

/* ---- multi-GPU merge (partial -> merge of main(), Main.kt:1309-1325) ----------------------- */
#define KQ_COMM_ID_BYTES 128
/* Rank 0 creates the id, the caller distributes it (any channel), all ranks call init. */
KQ_API int kq_comm_unique_id(kq_ctx* ctx, uint8_t id[KQ_COMM_ID_BYTES]);
KQ_API int kq_comm_init(kq_ctx* ctx, const uint8_t id[KQ_COMM_ID_BYTES], int rank, int nranks);
KQ_API int kq_comm_destroy(kq_ctx* ctx);
KQ_API int kq_comm_barrier(kq_ctx* ctx);
/* Max over ranks of a float (device-side timing reduction). */
KQ_API int kq_comm_allreduce_max_f32(kq_ctx* ctx, float* inout);
/* Low-cardinality merge; afterwards every rank holds the full merged result, bit-identical on all ranks. Small unions
 * (tens to a thousand partial groups per rank): one fixed-size ncclAllGather of the partial records (and of the strings
 * behind Utf8 group keys longer than 7 bytes), one host synchronisation, a rank-by-rank rebuild. Larger unions: all-gather
 * -> union dictionary -> dense arrays -> ncclAllReduce (sum / min / max). COLLECTIVE: every rank of the communicator must
 * call it; a rank whose aggregate carries an error joins the collective and all ranks return that error. */
KQ_API int kq_hashagg_merge_allreduce(kq_ctx* ctx, kq_hashagg* agg);
/* High-cardinality merge: partials bucketed by hash(key) % nranks, exchanged with a grouped
 * ncclSend/ncclRecv all-to-all, merged locally. Afterwards each key lives on exactly one rank. COLLECTIVE; the ranks
 * exchange a status word with the bucket sizes, so an error on one rank (incl. Utf8 group keys longer than 7 bytes,
 * which this merge does not carry) is returned by all of them instead of leaving the peers waiting. */
KQ_API int kq_hashagg_repartition_alltoall(kq_ctx* ctx, kq_hashagg* agg);

/* ---- scan side: CsvDataSource (Main.kt:276-357) -------------------------------------------- */
/* CsvDataSource.inferSchema (Main.kt:328-356), host only: detects the delimiter and line separator as the reference's
 * parser settings ask (Main.kt:289-296) and reads the first record. `names` receives the column names, one per line
 * ('\n'-terminated): the header fields if has_headers, else field_1.. (Main.kt:345-349). All columns are Utf8. */
KQ_API int kq_csv_header(const uint8_t* text, int64_t nbytes, int has_headers, char* names, size_t names_cap,
                         int* ncols, char* delimiter);
/* CsvDataSource.scan(projection) (Main.kt:304-326) fused with ReaderIterator.createBatch (Main.kt:251-273): the
 * whole text of a CSV file (< 2 GiB; a HOST buffer, or a DEVICE pointer when the file is already in HBM) -> ONE batch of Utf8 columns in HBM, every value trimmed
 * (Main.kt:263), never null. `projection`: file column indices in output order (NULL/0 = all columns); the caller
 * maps names to indices as Schema.select does (Main.kt:47-52). Rules C1-C9: csrc/kq_csv.cu. */
KQ_API int kq_csv_scan(kq_ctx* ctx, const uint8_t* text, int64_t nbytes, int has_headers, const int* projection,
                       int nproj, kq_batch** out);
/* The same scan as a Sequence<RecordBatch> (CsvDataSource.scan returns ReaderIterator batches, Main.kt:239-249, 323-325):
 * the text (HOST or DEVICE memory, any size, valid until close) streams through two device buffers in pieces of
 * `piece_bytes` (0 = 64 MiB; at most 512 MiB); every piece is cut at its last complete record and the copy of the next
 * piece runs under the scan of the current one (pinned or registered host memory: kq_host_register). kq_csv_reader_next
 * hands out one batch per piece — the concatenation of all batches equals kq_csv_scan's batch — and *out = NULL after the
 * last one; batches without rows are not yielded (Main.kt:245-247). A record that does not fit a piece is
 * KQ_ERR_UNSUPPORTED; an unbalanced quote surfaces with the last piece (KQ_ERR_ILLEGAL_STATE). */
KQ_API int kq_csv_reader_open(kq_ctx* ctx, const uint8_t* text, int64_t nbytes, int has_headers, const int* projection,
                              int nproj, int64_t piece_bytes, kq_csv_reader** out);
KQ_API int kq_csv_reader_next(kq_csv_reader* reader, kq_batch** out);
KQ_API int kq_csv_reader_close(kq_csv_reader* reader);

/* ---- bench-only: device-side synthetic tables -------------------------------------------- */
KQ_API int kq_generate(kq_ctx* ctx, const kq_gen_spec* specs, int ncols, uint64_t seed,
                       int64_t row_begin, int64_t row_end, kq_batch** out);

#ifdef __cplusplus
}
#endif
#endif /* KQGPU_H */
