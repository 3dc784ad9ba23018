/*
 * chdb_gpu.h -- C ABI of the B200-native filter / projection / compaction path of
 * ChapterhouseDB (alekLukanen/ChapterhouseQE).
 *
 * Drop-in boundary.  The reference's filter and materialize producer tasks call two pure,
 * synchronous functions (paths relative to the reference's
 * src/handlers/operator_handler/operators/):
 *
 *   record_utils/filter_record.rs:21-25
 *       filter_record(rec: Arc<RecordBatch>, table_aliases: &Vec<Vec<String>>, expr: &Expr)
 *           -> anyhow::Result<RecordBatch>               called at filter_tasks/filter_task.rs:99-103
 *   record_utils/record_projection.rs:16-20
 *       project_record(fields: &Vec<SelectItem>, record: Arc<RecordBatch>, table_aliases)
 *           -> anyhow::Result<RecordBatch>               called at materialize_tasks/materialize_files_task.rs:110-114
 *   record_utils/compute_value.rs:57-61
 *       compute_value(rec, table_aliases, expr) -> Result<ArrayDatum>   (the evaluator under both)
 *
 * A Rust `chdb-gpu` crate binds exactly the functions below (see INTEGRATION.md for the
 * extern "C" block and the TaskBuilder that replaces filter_task.rs:99).  Conventions:
 *
 *   - Batches cross as the Arrow C Data Interface (struct-typed ArrowArray + ArrowSchema:
 *     arrow-rs 53 `arrow::ffi::{to_ffi, from_ffi}` on a StructArray, pyarrow `_export_to_c`).
 *   - Expression trees cross as the serde-JSON of sqlparser 0.52 `Expr` / `Vec<SelectItem>`,
 *     i.e. `serde_json::to_string(&filter_config.expr)`; the planner is unchanged.
 *   - table_aliases crosses as JSON `[["alias"], [], ...]` (one list per input column), or NULL.
 *   - Every fallible call returns a chdb_code (0 = ok) and fills the optional chdb_status;
 *     codes mirror the reference's error enums (compute_value.rs:12-32, filter_record.rs:11-15,
 *     record_projection.rs:10-14) and the ArrowError variants its arrow calls can raise.
 *   - No thread-local state.  A chdb_ctx is single-owner (one operator instance = one GPU +
 *     one stream) but may be used from any OS thread (tokio tasks migrate between threads);
 *     different ctxs are independent.  Programs are immutable and shareable between ctxs of
 *     any device.
 *   - Inputs are borrowed for the duration of the call.  Outputs are owned by the caller and
 *     freed through ArrowArray.release / chdb_device_batch_release.
 *   - There is no CPU fallback: without a CUDA device every ctx call fails with CHDB_ERR_CUDA.
 */
#ifndef CHDB_GPU_H
#define CHDB_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Arrow C Data Interface (https://arrow.apache.org/docs/format/CDataInterface.html) ---- */
#ifndef ARROW_C_DATA_INTERFACE
#define ARROW_C_DATA_INTERFACE
#define ARROW_FLAG_DICTIONARY_ORDERED 1
#define ARROW_FLAG_NULLABLE 2
#define ARROW_FLAG_MAP_KEYS_SORTED 4
struct ArrowSchema {
  const char* format;
  const char* name;
  const char* metadata;
  int64_t flags;
  int64_t n_children;
  struct ArrowSchema** children;
  struct ArrowSchema* dictionary;
  void (*release)(struct ArrowSchema*);
  void* private_data;
};
struct ArrowArray {
  int64_t length;
  int64_t null_count;
  int64_t offset;
  int64_t n_buffers;
  int64_t n_children;
  const void** buffers;
  struct ArrowArray** children;
  struct ArrowArray* dictionary;
  void (*release)(struct ArrowArray*);
  void* private_data;
};
#endif

/* ---- status ---- */
typedef enum chdb_code {
  CHDB_OK = 0,
  /* ComputeValueError (compute_value.rs:12-32) */
  CHDB_ERR_VALUE_TYPE_NOT_IMPLEMENTED = 1,
  CHDB_ERR_EXPRESSION_TYPE_NOT_IMPLEMENTED = 2,
  CHDB_ERR_BINARY_OPERATOR_NOT_IMPLEMENTED = 3,
  CHDB_ERR_BINARY_OPERATION_CAST_FAILED = 4,
  CHDB_ERR_FAILED_TO_PARSE_AS_AN_INTEGER = 5,
  CHDB_ERR_FAILED_TO_PARSE_AS_A_FLOAT = 6,
  CHDB_ERR_COLUMN_NOT_FOUND = 7,
  CHDB_ERR_IDENTIFIER_NOT_FOUND = 8,
  CHDB_ERR_UNSUPPORTED_TYPE_COERSION = 9,
  /* FilterRecordError (filter_record.rs:11-15) */
  CHDB_ERR_CAST_TO_BOOLEAN_ARRAY_FAILED = 10,
  /* ProjectRecordError (record_projection.rs:10-14) and unsupported inputs of this library */
  CHDB_ERR_NOT_IMPLEMENTED = 11,
  /* ArrowError variants raised by the arrow kernels on this path */
  CHDB_ERR_ARITHMETIC_OVERFLOW = 12,
  CHDB_ERR_DIVIDE_BY_ZERO = 13,
  CHDB_ERR_COMPUTE_ERROR = 14,
  CHDB_ERR_INVALID_ARGUMENT = 15,
  /* this library */
  CHDB_ERR_CUDA = 16,
  CHDB_ERR_BAD_JSON = 17,
  CHDB_ERR_PANIC = 18 /* the reference would hit an .expect() (compute_value.rs:293) */
} chdb_code;

typedef struct chdb_status {
  int32_t code;
  char message[508];
} chdb_status;

/* Name of a code as the reference spells the enum variant, e.g. "ArithmeticOverflow". */
const char* chdb_code_name(int32_t code);
/* Library version string, and the SM architecture the kernels were compiled for ("sm_100a"). */
const char* chdb_version(void);
const char* chdb_compiled_arch(void);

/* ---- context: one per operator instance (one GPU, one stream) ---- */
typedef struct chdb_ctx chdb_ctx;
int32_t chdb_ctx_create(int32_t device, chdb_ctx** out, chdb_status* st);
void chdb_ctx_destroy(chdb_ctx* ctx);
/* cudaStream_t the ctx launches on (for callers that record their own CUDA events). */
void* chdb_ctx_stream(chdb_ctx* ctx);
int32_t chdb_ctx_device(chdb_ctx* ctx);
int32_t chdb_ctx_synchronize(chdb_ctx* ctx, chdb_status* st);
/* Number of kernel launches issued by this ctx so far (bench.py's gpu_launches). */
int64_t chdb_ctx_launch_count(chdb_ctx* ctx);
/* ...and how many of them ran a kernel specialised for the program at run time (see below). */
int64_t chdb_ctx_jit_launch_count(chdb_ctx* ctx);
/* Device allocations that missed the ctx's block cache and went to cudaMallocAsync (0 in a warmed-up steady state). */
int64_t chdb_ctx_alloc_miss_count(chdb_ctx* ctx);
/* Launch sets whose select kernel ran on the ctx's second stream, next to the previous launch set's gather kernel
 * (device-resident batches created at least one launch set earlier; CHDB_OVERLAP=0 turns it off). */
int64_t chdb_ctx_overlapped_count(chdb_ctx* ctx);

/* ---- run-time specialisation ----
 * Long scans run the kernels' own source compiled by NVRTC with the program's bytecode baked
 * in as constants (same code, dispatch folded away).  Controlled by the CHDB_JIT environment
 * variable: "0" interpreter kernel only, unset / "1" batches of >= 2^18 rows, "always" every launch.
 * Without libnvrtc the interpreter kernel is used; results are identical either way. */
int32_t chdb_jit_available(char* why, size_t cap);

/* ---- programs: an expression tree lowered to register bytecode for one input schema ----
 * compile_filter  : predicate = `expr`,  outputs = every input column       (filter_record)
 * compile_project : no predicate,        outputs = `select_items`           (project_record)
 * compile_filter_project : both fused in one pass; identical results to
 *                   project_record(filter_record(rec)) incl. "errors only on surviving rows".
 * Compilation never touches the GPU. */
typedef struct chdb_program chdb_program;
int32_t chdb_program_compile_filter(const char* expr_json, const struct ArrowSchema* in_schema,
                                    const char* table_aliases_json, chdb_program** out, chdb_status* st);
int32_t chdb_program_compile_project(const char* select_items_json, const struct ArrowSchema* in_schema,
                                     const char* table_aliases_json, chdb_program** out, chdb_status* st);
int32_t chdb_program_compile_filter_project(const char* expr_json, const char* select_items_json,
                                            const struct ArrowSchema* in_schema,
                                            const char* table_aliases_json, chdb_program** out,
                                            chdb_status* st);
/* SQL nodes beyond the reference's compute_value (README.md:44-74 lists them as missing).  Off by
 * default: every such node then returns the reference's own error (BinaryOperatorNotImplemented for `-`,
 * ExpressionTypeNotImplemented for the others).  The mask is read when a program is compiled; returns the previous mask.
 *   CHDB_EXT_OPERATORS: binary `-` (numeric::sub: checked integers, IEEE floats), unary `-` (numeric::neg: checked on signed
 *                       integers, sign flip on floats, error on unsigned) and `+`, NOT (compute::not over the Boolean cast,
 *                       nulls stay null), IS NULL / IS NOT NULL (never null);
 *   CHDB_EXT_KLEENE:    AND / OR as SQL three-valued logic (and_kleene / or_kleene) instead of the reference's and / or. */
#define CHDB_EXT_OPERATORS 1u
#define CHDB_EXT_KLEENE 2u
uint32_t chdb_set_sql_extensions(uint32_t mask);
uint32_t chdb_get_sql_extensions(void);
void chdb_program_release(chdb_program* prog);
/* Human-readable bytecode listing; returns the number of bytes needed (excluding NUL). */
size_t chdb_program_disassemble(const chdb_program* prog, char* buf, size_t cap);
int32_t chdb_program_num_instructions(const chdb_program* prog);
/* The generated specialisation prologue (constants only; device_code.cuh follows it). */
size_t chdb_program_jit_source(const chdb_program* prog, char* buf, size_t cap);
/* Compiles the specialised kernel for sm_100a without a GPU (build / test check). */
int32_t chdb_program_jit_check(const chdb_program* prog, int64_t* cubin_bytes, char* log, size_t cap, chdb_status* st);

/* ---- host batches: same contract as the reference's functions ----
 * `in` / `in_schema`: a struct array whose children are the batch columns.
 * `out` / `out_schema`: filled with a new struct array (caller releases both). */
int32_t chdb_filter_record(chdb_ctx* ctx, const chdb_program* prog, const struct ArrowArray* in,
                           const struct ArrowSchema* in_schema, struct ArrowArray* out,
                           struct ArrowSchema* out_schema, chdb_status* st);
int32_t chdb_project_record(chdb_ctx* ctx, const chdb_program* prog, const struct ArrowArray* in,
                            const struct ArrowSchema* in_schema, struct ArrowArray* out,
                            struct ArrowSchema* out_schema, chdb_status* st);
/* Non-blocking forms for callers that must not park their thread (the reference calls filter_record inline on a
 * tokio worker, filter_task.rs:86-125): *_async enqueues upload + kernels and returns at once; `in` stays
 * borrowed until chdb_poll() has returned 1.  chdb_poll never blocks: 0 = not ready, 1 = ready, < 0 = -chdb_code
 * (see *st).  chdb_pending_result hands the output over (it waits if called early); release the handle after. */
typedef struct chdb_pending chdb_pending;
int32_t chdb_filter_record_async(chdb_ctx* ctx, const chdb_program* prog, const struct ArrowArray* in,
                                 const struct ArrowSchema* in_schema, chdb_pending** out, chdb_status* st);
int32_t chdb_project_record_async(chdb_ctx* ctx, const chdb_program* prog, const struct ArrowArray* in,
                                  const struct ArrowSchema* in_schema, chdb_pending** out, chdb_status* st);
int32_t chdb_poll(chdb_pending* pending, chdb_status* st);
int32_t chdb_pending_result(chdb_pending* pending, struct ArrowArray* out, struct ArrowSchema* out_schema,
                            chdb_status* st);
void chdb_pending_release(chdb_pending* pending);
/* One-shot forms taking the expression each call (exactly the reference signatures). */
int32_t chdb_filter_record_expr(chdb_ctx* ctx, const struct ArrowArray* in, const struct ArrowSchema* in_schema,
                                const char* table_aliases_json, const char* expr_json,
                                struct ArrowArray* out, struct ArrowSchema* out_schema, chdb_status* st);
int32_t chdb_project_record_items(chdb_ctx* ctx, const char* select_items_json, const struct ArrowArray* in,
                                  const struct ArrowSchema* in_schema, const char* table_aliases_json,
                                  struct ArrowArray* out, struct ArrowSchema* out_schema, chdb_status* st);
/* compute_value(rec, aliases, expr): a one-column batch named "value"; *is_scalar as ArrayDatum. */
int32_t chdb_compute_value(chdb_ctx* ctx, const struct ArrowArray* in, const struct ArrowSchema* in_schema,
                           const char* table_aliases_json, const char* expr_json, struct ArrowArray* out,
                           struct ArrowSchema* out_schema, int32_t* is_scalar, chdb_status* st);

/* ---- device-resident batches (what the exchanges carry between read_files, filter and
 *      materialize so only the final result crosses PCIe) ---- */
typedef struct chdb_device_batch chdb_device_batch;
int32_t chdb_upload(chdb_ctx* ctx, const struct ArrowArray* in, const struct ArrowSchema* in_schema,
                    chdb_device_batch** out, chdb_status* st);
/* Wrap device buffers the caller already owns (no copy; they must outlive the batch).
 * Per column c: values[c] (Utf8: value bytes), validity[c] or NULL, offsets[c] (Utf8 only, int32[n+1]).
 * Every buffer must be 16-byte aligned and readable for 32 bytes past its logical end.
 * Their contents must be complete in ctx-stream order at the time of this call (filled on the ctx stream, on a stream
 * the ctx stream has waited for, or before a synchronisation), and must not change while the batch lives. */
int32_t chdb_device_batch_wrap(chdb_ctx* ctx, const struct ArrowSchema* schema, int64_t num_rows,
                               const void* const* values, const void* const* validity,
                               const void* const* offsets, chdb_device_batch** out, chdb_status* st);
/* Asynchronous on the ctx stream; the result's row count is known after the stream is
 * synchronised (chdb_device_batch_num_rows does that). Runs whatever the program holds
 * (filter, project or fused). */
int32_t chdb_run_device(chdb_ctx* ctx, const chdb_program* prog, const chdb_device_batch* in,
                        chdb_device_batch** out, chdb_status* st);
/* The same over `count` batches of one schema in ONE launch set: the reference's native records are at most
 * 10 000 rows (src/planner/physical_planner.rs:323, read_files_task.rs:249-252), far too small to fill a GPU one
 * at a time.  out[i] is the result of in[i] -- one output batch per record, so the record_id <-> rec_<id>.parquet
 * mapping of materialize_files_task.rs:119 is unchanged.  Errors of the data (overflow, divide by zero) stay
 * per batch (chdb_device_batch_status). */
int32_t chdb_run_device_many(chdb_ctx* ctx, const chdb_program* prog, const chdb_device_batch* const* in,
                             int32_t count, chdb_device_batch** out, chdb_status* st);
/* 1 if the run that produces `b` has finished (row count, status and buffers can be read without waiting),
 * 0 if not; never blocks (cudaEventQuery). */
int32_t chdb_device_batch_ready(chdb_ctx* ctx, const chdb_device_batch* b, chdb_status* st);
/* Checks the device error word of a finished run (ArithmeticOverflow / DivideByZero). */
int32_t chdb_device_batch_status(chdb_ctx* ctx, const chdb_device_batch* b, chdb_status* st);
int64_t chdb_device_batch_num_rows(chdb_ctx* ctx, const chdb_device_batch* b, chdb_status* st);
int32_t chdb_device_batch_num_columns(const chdb_device_batch* b);
/* Device pointers and byte sizes of column `col` (after sync): values (Utf8: the value bytes the
 * offsets index, i.e. buffer start + offsets[0]), validity bitmap or NULL, Utf8 offsets or NULL.
 * For zero-copy hand-off to other device code (peer gather at materialize, a GPU Parquet encoder). */
int32_t chdb_device_batch_column(chdb_ctx* ctx, const chdb_device_batch* b, int32_t col, const void** values,
                                 int64_t* values_bytes, const void** validity, int64_t* validity_bytes,
                                 const void** offsets, int64_t* offsets_bytes, chdb_status* st);
/* Bytes of HBM written for this batch's columns (values + validity + offsets), after sync. */
int64_t chdb_device_batch_nbytes(chdb_ctx* ctx, const chdb_device_batch* b, chdb_status* st);
int32_t chdb_download(chdb_ctx* ctx, const chdb_device_batch* b, struct ArrowArray* out,
                      struct ArrowSchema* out_schema, chdb_status* st);
/* Packs every buffer of `count` finished batches into ONE contiguous device buffer (pieces 256-byte aligned; order:
 * batch, column, {validity, offsets, values}) with device-to-device copies on the ctx stream: the unit a
 * materialize-side gather between processes sends over NVLink -- one message per GPU instead of one per buffer.
 * sizes_out: int64[count * num_columns * 3] byte sizes of the pieces (0 = absent), or NULL.  dst == NULL: only the
 * sizes and *total_bytes are computed (to size the destination). */
int32_t chdb_device_batches_pack(chdb_ctx* ctx, const chdb_device_batch* const* batches, int32_t count, void* dst,
                                 int64_t capacity, int64_t* sizes_out, int64_t* total_bytes, chdb_status* st);
/* Copy a finished batch's buffers to another ctx's GPU (cudaMemcpyPeerAsync over NVLink; peer access is
 * enabled on first use): the materialize-side gather of per-GPU results.  Asynchronous on the destination
 * ctx's stream, ordered after the source ctx's stream; the source may be released right after the call. */
int32_t chdb_peer_copy(chdb_ctx* dst_ctx, chdb_ctx* src_ctx, const chdb_device_batch* src,
                       chdb_device_batch** out, chdb_status* st);
/* Batch handles are reference counted: a new handle starts at 1, _retain adds a reference, _release drops one
 * (the buffers go back to the ctx's block cache with the last one). */
void chdb_device_batch_retain(chdb_device_batch* b);
void chdb_device_batch_release(chdb_device_batch* b);
void chdb_device_batch_release_many(chdb_device_batch* const* batches, int32_t count);

/* ---- device-resident record pool: the GPU-aware exchange's RecordPool ----
 * The reference's exchange keeps `records: HashMap<u64, Arc<RecordBatch>>` with a per-consumer-operator queue
 * and drops a record once every consumer operator has completed it (exchange_operator.rs:566-777, :727-733);
 * memory management of that pool is a listed TODO (DEV_NOTES.md:133-140).  This is that pool for device batches:
 * records are held by reference, handed out by reference, and -- beyond `budget_bytes` of HBM held -- spilled to
 * pinned host memory, least recently used first (records a consumer currently holds are never spilled); a spilled
 * record is uploaded again by _get.  budget_bytes <= 0: no limit.  Thread-safe. */
typedef struct chdb_record_pool chdb_record_pool;
int32_t chdb_record_pool_create(chdb_ctx* ctx, int64_t budget_bytes, chdb_record_pool** out, chdb_status* st);
void chdb_record_pool_destroy(chdb_record_pool* pool);
/* add_record (exchange_operator.rs:596-619): the pool takes its own reference; `consumers` = consumer operators. */
int32_t chdb_record_pool_add(chdb_record_pool* pool, uint64_t record_id, chdb_device_batch* batch, int32_t consumers,
                             chdb_status* st);
/* get_next_record (:621-667): a new reference to the record's device batch (caller releases it). */
int32_t chdb_record_pool_get(chdb_record_pool* pool, uint64_t record_id, chdb_device_batch** out, chdb_status* st);
/* operator completed the record (:727-733): the record is dropped after the last consumer. */
int32_t chdb_record_pool_complete(chdb_record_pool* pool, uint64_t record_id, chdb_status* st);
void chdb_record_pool_stats(chdb_record_pool* pool, int64_t* records, int64_t* device_bytes, int64_t* spilled_records,
                            int64_t* spilled_bytes);

/* ---- Parquet -> device decode: the step immediately upstream of the filter ----
 * Replaces the CPU decode of read_files (table_func_tasks/read_files_task.rs:233-282: ParquetRecordBatchStreamBuilder
 * with_batch_size(max_rows_per_batch)) for the files the reference writes (parquet-rs default writer properties,
 * src/bin/create_sample_data.rs:221-224: uncompressed, data page v1, dictionary encoding with PLAIN fallback).  The host
 * reads the footer, the page headers and the RLE run headers; a row group's column chunks cross PCIe as stored and are
 * decoded by kernels into a device batch (one batch per row group; the filter runs on it in place).
 * Supported: flat schemas; BOOLEAN, INT32 (+ INT_8/16, UINT_8/16/32), INT64 (+ UINT_64), FLOAT, DOUBLE, BYTE_ARRAY/UTF8;
 * PLAIN, PLAIN_DICTIONARY / RLE_DICTIONARY, RLE booleans; data pages v1 and v2; REQUIRED / OPTIONAL; UNCOMPRESSED.
 * Anything else: CHDB_ERR_NOT_IMPLEMENTED.  `file` (the whole file's bytes; pinned memory makes the copy asynchronous)
 * stays borrowed until chdb_parquet_close.  chdb_parquet_open never touches the GPU. */
typedef struct chdb_parquet chdb_parquet;
int32_t chdb_parquet_open(const void* file, int64_t len, chdb_parquet** out, chdb_status* st);
void chdb_parquet_close(chdb_parquet* f);
int32_t chdb_parquet_num_row_groups(const chdb_parquet* f);
int32_t chdb_parquet_num_columns(const chdb_parquet* f);
int64_t chdb_parquet_num_rows(const chdb_parquet* f);
int64_t chdb_parquet_row_group_num_rows(const chdb_parquet* f, int32_t row_group);
/* name / Arrow C format string / nullability of column `col` (pointers valid until close). */
int32_t chdb_parquet_column(const chdb_parquet* f, int32_t col, const char** name, const char** arrow_format,
                            int32_t* nullable);
/* Host-only walk of a row group's page and run headers: CHDB_OK if chdb_parquet_decode_row_group supports every page
 * of it (else the reason, e.g. a compression codec); *pages / *runs = data pages / hybrid runs found. */
int32_t chdb_parquet_check_row_group(const chdb_parquet* f, int32_t row_group, int64_t* pages, int64_t* runs,
                                     chdb_status* st);
/* Decodes one row group on the ctx stream and returns when the batch is complete (one synchronise: the string
 * buffers are sized from the decoded lengths). */
int32_t chdb_parquet_decode_row_group(chdb_ctx* ctx, const chdb_parquet* f, int32_t row_group,
                                      chdb_device_batch** out, chdb_status* st);

/* Row groups [first, first + count) into out[0 .. count): the same results, with row group i+1's bytes crossing PCIe
 * (on a second stream of the ctx, when `file` is pinned) while row group i's kernels run. */
int32_t chdb_parquet_decode_row_groups(chdb_ctx* ctx, const chdb_parquet* f, int32_t first, int32_t count,
                                       chdb_device_batch** out, chdb_status* st);

/* ---- device batches -> Parquet, with batch coalescing: the step immediately downstream of the projection ----
 * Replaces the write half of materialize_files_task.rs:116-141 (AsyncArrowWriter over one <= 10 000-row record per file)
 * and implements the compaction DEV_NOTES.md:117-122 lists as a TODO: the first `*consumed` of `count` finished device
 * batches of one schema become ONE Parquet file image in pinned host memory; consecutive batches are coalesced into row
 * groups of at most max_rows_per_row_group rows (<= 0: 1 Mi, parquet-rs's default; a batch is never split), at most
 * max_row_groups of them (<= 0: no limit; the caller encodes the remaining batches into the next file).  Every batch is
 * one data page per column.  Written: data page v1, PLAIN, RLE / bit-packed definition levels (the Arrow validity bitmap
 * as one bit-packed run), UNCOMPRESSED; the decoder's types; nullable columns OPTIONAL.  Kernels squeeze null slots out,
 * widen narrow integers, re-pack booleans and interleave BYTE_ARRAY length prefixes; the host writes page headers and
 * footer.  *file_out (length *len_out) is freed with chdb_parquet_image_free. */
int32_t chdb_parquet_encode(chdb_ctx* ctx, const chdb_device_batch* const* batches, int32_t count,
                            int64_t max_rows_per_row_group, int32_t max_row_groups, void** file_out, int64_t* len_out,
                            int32_t* consumed, int32_t* row_groups, chdb_status* st);
void chdb_parquet_image_free(void* file);

#ifdef __cplusplus
}
#endif
#endif /* CHDB_GPU_H */
