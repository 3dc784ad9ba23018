//! Raw bindings of include/chdb_gpu.h -- every function the header declares, in header order.
//! SOURCE ONLY (no Rust toolchain in the build image); tests/test_library_cpu.py checks that the
//! names declared here are exported by libchdb_gpu.so with the header's signatures.
use arrow::ffi::{FFI_ArrowArray, FFI_ArrowSchema};
use std::os::raw::{c_char, c_void};

#[repr(C)]
pub struct ChdbStatus {
    pub code: i32,
    pub message: [c_char; 508],
}

macro_rules! opaque {
    ($name:ident) => {
        #[repr(C)]
        pub struct $name {
            _private: [u8; 0],
        }
    };
}
opaque!(ChdbCtx);
opaque!(ChdbProgram);
opaque!(ChdbDeviceBatch);
opaque!(ChdbPending);
opaque!(ChdbRecordPool);
opaque!(ChdbParquet);

extern "C" {
    // ---- status ----
    pub fn chdb_code_name(code: i32) -> *const c_char;
    pub fn chdb_version() -> *const c_char;
    pub fn chdb_compiled_arch() -> *const c_char;

    // ---- context: one per operator instance (one GPU, one stream) ----
    pub fn chdb_ctx_create(device: i32, out: *mut *mut ChdbCtx, st: *mut ChdbStatus) -> i32;
    pub fn chdb_ctx_destroy(ctx: *mut ChdbCtx);
    pub fn chdb_ctx_stream(ctx: *mut ChdbCtx) -> *mut c_void;
    pub fn chdb_ctx_device(ctx: *mut ChdbCtx) -> i32;
    pub fn chdb_ctx_synchronize(ctx: *mut ChdbCtx, st: *mut ChdbStatus) -> i32;
    pub fn chdb_ctx_launch_count(ctx: *mut ChdbCtx) -> i64;
    pub fn chdb_ctx_jit_launch_count(ctx: *mut ChdbCtx) -> i64;
    pub fn chdb_ctx_alloc_miss_count(ctx: *mut ChdbCtx) -> i64;
    pub fn chdb_ctx_overlapped_count(ctx: *mut ChdbCtx) -> i64;

    // ---- Parquet -> device decode (read_files_task.rs:233-282) ----
    pub fn chdb_parquet_open(file: *const c_void, len: i64, out: *mut *mut ChdbParquet, st: *mut ChdbStatus) -> i32;
    pub fn chdb_parquet_close(f: *mut ChdbParquet);
    pub fn chdb_parquet_num_row_groups(f: *const ChdbParquet) -> i32;
    pub fn chdb_parquet_num_columns(f: *const ChdbParquet) -> i32;
    pub fn chdb_parquet_num_rows(f: *const ChdbParquet) -> i64;
    pub fn chdb_parquet_row_group_num_rows(f: *const ChdbParquet, row_group: i32) -> i64;
    pub fn chdb_parquet_column(f: *const ChdbParquet, col: i32, name: *mut *const c_char, arrow_format: *mut *const c_char,
                               nullable: *mut i32) -> i32;
    pub fn chdb_parquet_check_row_group(f: *const ChdbParquet, row_group: i32, pages: *mut i64, runs: *mut i64,
                                        st: *mut ChdbStatus) -> i32;
    pub fn chdb_parquet_decode_row_groups(ctx: *mut ChdbCtx, f: *const ChdbParquet, first: i32, count: i32,
                                          out: *mut *mut ChdbDeviceBatch, st: *mut ChdbStatus) -> i32;
    pub fn chdb_parquet_decode_row_group(ctx: *mut ChdbCtx, f: *const ChdbParquet, row_group: i32,
                                         out: *mut *mut ChdbDeviceBatch, st: *mut ChdbStatus) -> i32;
    // ---- device batches -> Parquet with record coalescing (materialize_files_task.rs:116-141, DEV_NOTES.md:117-122) ----
    pub fn chdb_parquet_encode(ctx: *mut ChdbCtx, batches: *const *const ChdbDeviceBatch, count: i32, max_rows_per_row_group: i64,
                               max_row_groups: i32, file_out: *mut *mut c_void, len_out: *mut i64, consumed: *mut i32,
                               row_groups: *mut i32, st: *mut ChdbStatus) -> i32;
    pub fn chdb_parquet_image_free(file: *mut c_void);
    pub fn chdb_jit_available(why: *mut c_char, cap: usize) -> i32;
    pub fn chdb_set_sql_extensions(mask: u32) -> u32;
    pub fn chdb_get_sql_extensions() -> u32;

    // ---- programs ----
    pub fn chdb_program_compile_filter(
        expr_json: *const c_char,
        in_schema: *const FFI_ArrowSchema,
        table_aliases_json: *const c_char,
        out: *mut *mut ChdbProgram,
        st: *mut ChdbStatus,
    ) -> i32;
    pub fn chdb_program_compile_project(
        select_items_json: *const c_char,
        in_schema: *const FFI_ArrowSchema,
        table_aliases_json: *const c_char,
        out: *mut *mut ChdbProgram,
        st: *mut ChdbStatus,
    ) -> i32;
    pub fn chdb_program_compile_filter_project(
        expr_json: *const c_char,
        select_items_json: *const c_char,
        in_schema: *const FFI_ArrowSchema,
        table_aliases_json: *const c_char,
        out: *mut *mut ChdbProgram,
        st: *mut ChdbStatus,
    ) -> i32;
    pub fn chdb_program_release(prog: *mut ChdbProgram);
    pub fn chdb_program_disassemble(prog: *const ChdbProgram, buf: *mut c_char, cap: usize) -> usize;
    pub fn chdb_program_num_instructions(prog: *const ChdbProgram) -> i32;
    pub fn chdb_program_jit_source(prog: *const ChdbProgram, buf: *mut c_char, cap: usize) -> usize;
    pub fn chdb_program_jit_check(
        prog: *const ChdbProgram,
        cubin_bytes: *mut i64,
        log: *mut c_char,
        cap: usize,
        st: *mut ChdbStatus,
    ) -> i32;

    // ---- host batches: the reference's contract ----
    pub fn chdb_filter_record(
        ctx: *mut ChdbCtx,
        prog: *const ChdbProgram,
        input: *const FFI_ArrowArray,
        in_schema: *const FFI_ArrowSchema,
        out: *mut FFI_ArrowArray,
        out_schema: *mut FFI_ArrowSchema,
        st: *mut ChdbStatus,
    ) -> i32;
    pub fn chdb_project_record(
        ctx: *mut ChdbCtx,
        prog: *const ChdbProgram,
        input: *const FFI_ArrowArray,
        in_schema: *const FFI_ArrowSchema,
        out: *mut FFI_ArrowArray,
        out_schema: *mut FFI_ArrowSchema,
        st: *mut ChdbStatus,
    ) -> i32;
    pub fn chdb_filter_record_async(
        ctx: *mut ChdbCtx,
        prog: *const ChdbProgram,
        input: *const FFI_ArrowArray,
        in_schema: *const FFI_ArrowSchema,
        out: *mut *mut ChdbPending,
        st: *mut ChdbStatus,
    ) -> i32;
    pub fn chdb_project_record_async(
        ctx: *mut ChdbCtx,
        prog: *const ChdbProgram,
        input: *const FFI_ArrowArray,
        in_schema: *const FFI_ArrowSchema,
        out: *mut *mut ChdbPending,
        st: *mut ChdbStatus,
    ) -> i32;
    pub fn chdb_poll(pending: *mut ChdbPending, st: *mut ChdbStatus) -> i32;
    pub fn chdb_pending_result(
        pending: *mut ChdbPending,
        out: *mut FFI_ArrowArray,
        out_schema: *mut FFI_ArrowSchema,
        st: *mut ChdbStatus,
    ) -> i32;
    pub fn chdb_pending_release(pending: *mut ChdbPending);
    pub fn chdb_filter_record_expr(
        ctx: *mut ChdbCtx,
        input: *const FFI_ArrowArray,
        in_schema: *const FFI_ArrowSchema,
        table_aliases_json: *const c_char,
        expr_json: *const c_char,
        out: *mut FFI_ArrowArray,
        out_schema: *mut FFI_ArrowSchema,
        st: *mut ChdbStatus,
    ) -> i32;
    pub fn chdb_project_record_items(
        ctx: *mut ChdbCtx,
        select_items_json: *const c_char,
        input: *const FFI_ArrowArray,
        in_schema: *const FFI_ArrowSchema,
        table_aliases_json: *const c_char,
        out: *mut FFI_ArrowArray,
        out_schema: *mut FFI_ArrowSchema,
        st: *mut ChdbStatus,
    ) -> i32;
    pub fn chdb_compute_value(
        ctx: *mut ChdbCtx,
        input: *const FFI_ArrowArray,
        in_schema: *const FFI_ArrowSchema,
        table_aliases_json: *const c_char,
        expr_json: *const c_char,
        out: *mut FFI_ArrowArray,
        out_schema: *mut FFI_ArrowSchema,
        is_scalar: *mut i32,
        st: *mut ChdbStatus,
    ) -> i32;

    // ---- device-resident batches ----
    pub fn chdb_upload(
        ctx: *mut ChdbCtx,
        input: *const FFI_ArrowArray,
        in_schema: *const FFI_ArrowSchema,
        out: *mut *mut ChdbDeviceBatch,
        st: *mut ChdbStatus,
    ) -> i32;
    pub fn chdb_device_batch_wrap(
        ctx: *mut ChdbCtx,
        schema: *const FFI_ArrowSchema,
        num_rows: i64,
        values: *const *const c_void,
        validity: *const *const c_void,
        offsets: *const *const c_void,
        out: *mut *mut ChdbDeviceBatch,
        st: *mut ChdbStatus,
    ) -> i32;
    pub fn chdb_run_device(
        ctx: *mut ChdbCtx,
        prog: *const ChdbProgram,
        input: *const ChdbDeviceBatch,
        out: *mut *mut ChdbDeviceBatch,
        st: *mut ChdbStatus,
    ) -> i32;
    pub fn chdb_run_device_many(
        ctx: *mut ChdbCtx,
        prog: *const ChdbProgram,
        input: *const *const ChdbDeviceBatch,
        count: i32,
        out: *mut *mut ChdbDeviceBatch,
        st: *mut ChdbStatus,
    ) -> i32;
    pub fn chdb_device_batch_ready(ctx: *mut ChdbCtx, b: *const ChdbDeviceBatch, st: *mut ChdbStatus) -> i32;
    pub fn chdb_device_batch_status(ctx: *mut ChdbCtx, b: *const ChdbDeviceBatch, st: *mut ChdbStatus) -> i32;
    pub fn chdb_device_batch_num_rows(ctx: *mut ChdbCtx, b: *const ChdbDeviceBatch, st: *mut ChdbStatus) -> i64;
    pub fn chdb_device_batch_num_columns(b: *const ChdbDeviceBatch) -> i32;
    pub fn chdb_device_batch_column(
        ctx: *mut ChdbCtx,
        b: *const ChdbDeviceBatch,
        col: i32,
        values: *mut *const c_void,
        values_bytes: *mut i64,
        validity: *mut *const c_void,
        validity_bytes: *mut i64,
        offsets: *mut *const c_void,
        offsets_bytes: *mut i64,
        st: *mut ChdbStatus,
    ) -> i32;
    pub fn chdb_device_batch_nbytes(ctx: *mut ChdbCtx, b: *const ChdbDeviceBatch, st: *mut ChdbStatus) -> i64;
    pub fn chdb_download(
        ctx: *mut ChdbCtx,
        b: *const ChdbDeviceBatch,
        out: *mut FFI_ArrowArray,
        out_schema: *mut FFI_ArrowSchema,
        st: *mut ChdbStatus,
    ) -> i32;
    pub fn chdb_device_batches_pack(
        ctx: *mut ChdbCtx,
        batches: *const *const ChdbDeviceBatch,
        count: i32,
        dst: *mut c_void,
        capacity: i64,
        sizes_out: *mut i64,
        total_bytes: *mut i64,
        st: *mut ChdbStatus,
    ) -> i32;
    pub fn chdb_peer_copy(
        dst_ctx: *mut ChdbCtx,
        src_ctx: *mut ChdbCtx,
        src: *const ChdbDeviceBatch,
        out: *mut *mut ChdbDeviceBatch,
        st: *mut ChdbStatus,
    ) -> i32;
    pub fn chdb_device_batch_retain(b: *mut ChdbDeviceBatch);
    pub fn chdb_device_batch_release(b: *mut ChdbDeviceBatch);
    pub fn chdb_device_batch_release_many(batches: *const *mut ChdbDeviceBatch, count: i32);

    // ---- device-resident record pool (what a GPU-aware exchange keeps instead of Arc<RecordBatch>) ----
    pub fn chdb_record_pool_create(ctx: *mut ChdbCtx, budget_bytes: i64, out: *mut *mut ChdbRecordPool, st: *mut ChdbStatus) -> i32;
    pub fn chdb_record_pool_destroy(pool: *mut ChdbRecordPool);
    pub fn chdb_record_pool_add(
        pool: *mut ChdbRecordPool,
        record_id: u64,
        batch: *mut ChdbDeviceBatch,
        consumers: i32,
        st: *mut ChdbStatus,
    ) -> i32;
    pub fn chdb_record_pool_get(
        pool: *mut ChdbRecordPool,
        record_id: u64,
        out: *mut *mut ChdbDeviceBatch,
        st: *mut ChdbStatus,
    ) -> i32;
    pub fn chdb_record_pool_complete(pool: *mut ChdbRecordPool, record_id: u64, st: *mut ChdbStatus) -> i32;
    pub fn chdb_record_pool_stats(
        pool: *mut ChdbRecordPool,
        records: *mut i64,
        device_bytes: *mut i64,
        spilled_records: *mut i64,
        spilled_bytes: *mut i64,
    );
}
