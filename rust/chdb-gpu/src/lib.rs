//! chdb-gpu: the B200 filter / projection path behind the signatures of
//! `record_utils::{filter_record, project_record}` (record_utils/mod.rs:13-15 of ChapterhouseDB),
//! plus the two `TaskBuilder`s (`tasks.rs`) a GPU worker registers instead of the CPU ones.
//!
//! SOURCE ONLY -- not compiled in the build image (no Rust toolchain). The compiled, tested host
//! layer is the C++ library behind include/chdb_gpu.h; this crate is the thin wrapper over it.
pub mod ffi;
#[cfg(feature = "chapterhouse-tasks")]
pub mod tasks;

use anyhow::{anyhow, Result};
use arrow::array::{Array, RecordBatch, StructArray};
use arrow::ffi::{from_ffi, to_ffi, FFI_ArrowArray, FFI_ArrowSchema};
use std::ffi::{CStr, CString};
use std::sync::Arc;
use std::time::Duration;

/// One GPU context per operator instance (one device, one stream). `Send`: tokio may move the
/// owning task between worker threads; the library keeps no thread-local state.
pub struct GpuContext(*mut ffi::ChdbCtx);
unsafe impl Send for GpuContext {}

impl GpuContext {
    pub fn new(device: i32) -> Result<Self> {
        let mut ctx = std::ptr::null_mut();
        let mut st = new_status();
        check(unsafe { ffi::chdb_ctx_create(device, &mut ctx, &mut st) }, &st)?;
        Ok(GpuContext(ctx))
    }
    pub fn device(&self) -> i32 {
        unsafe { ffi::chdb_ctx_device(self.0) }
    }
}
impl Drop for GpuContext {
    fn drop(&mut self) {
        unsafe { ffi::chdb_ctx_destroy(self.0) }
    }
}

/// An `Expr` / `Vec<SelectItem>` lowered to bytecode for one input schema. Compiled on the first
/// batch (the schema is only known then) and reused for every later batch of the exchange.
pub struct GpuProgram(*mut ffi::ChdbProgram);
unsafe impl Send for GpuProgram {}
unsafe impl Sync for GpuProgram {}
impl Drop for GpuProgram {
    fn drop(&mut self) {
        unsafe { ffi::chdb_program_release(self.0) }
    }
}

/// A device-resident RecordBatch: what a GPU-aware exchange carries in the `record` slot of
/// `ExchangeRequests::{SendRecordRequest, GetNextRecordResponseRecord}` (messages/exchange.rs:63-89)
/// between operators of one worker; `download()` is the fallback for consumers on another worker.
pub struct DeviceBatch {
    raw: *mut ffi::ChdbDeviceBatch,
    ctx: *mut ffi::ChdbCtx,
}
unsafe impl Send for DeviceBatch {}
impl Drop for DeviceBatch {
    fn drop(&mut self) {
        unsafe { ffi::chdb_device_batch_release(self.raw) }
    }
}

fn new_status() -> ffi::ChdbStatus {
    ffi::ChdbStatus { code: 0, message: [0; 508] }
}

fn check(rc: i32, st: &ffi::ChdbStatus) -> Result<()> {
    if rc == 0 {
        return Ok(());
    }
    let kind = unsafe { CStr::from_ptr(ffi::chdb_code_name(rc)) }.to_string_lossy();
    let msg = unsafe { CStr::from_ptr(st.message.as_ptr()) }.to_string_lossy();
    // same text the reference's thiserror enums / ArrowError would carry through anyhow
    Err(anyhow!("{kind}: {msg}"))
}

fn export(rec: &RecordBatch) -> Result<(FFI_ArrowArray, FFI_ArrowSchema)> {
    let s: StructArray = rec.clone().into();
    Ok(to_ffi(&s.to_data())?)
}

fn import(out: FFI_ArrowArray, out_schema: &FFI_ArrowSchema) -> Result<RecordBatch> {
    let data = unsafe { from_ffi(out, out_schema) }?;
    Ok(RecordBatch::from(StructArray::from(data)))
}

fn aliases_json(table_aliases: &Vec<Vec<String>>) -> Result<CString> {
    Ok(CString::new(serde_json::to_string(table_aliases)?)?)
}

pub fn compile_filter(rec: &RecordBatch, table_aliases: &Vec<Vec<String>>, expr: &sqlparser::ast::Expr) -> Result<GpuProgram> {
    let (_, schema) = export(rec)?;
    let expr = CString::new(serde_json::to_string(expr)?)?;
    let al = aliases_json(table_aliases)?;
    let (mut prog, mut st) = (std::ptr::null_mut(), new_status());
    check(unsafe { ffi::chdb_program_compile_filter(expr.as_ptr(), &schema, al.as_ptr(), &mut prog, &mut st) }, &st)?;
    Ok(GpuProgram(prog))
}

pub fn compile_project(
    fields: &Vec<sqlparser::ast::SelectItem>,
    rec: &RecordBatch,
    table_aliases: &Vec<Vec<String>>,
) -> Result<GpuProgram> {
    let (_, schema) = export(rec)?;
    let items = CString::new(serde_json::to_string(fields)?)?;
    let al = aliases_json(table_aliases)?;
    let (mut prog, mut st) = (std::ptr::null_mut(), new_status());
    check(unsafe { ffi::chdb_program_compile_project(items.as_ptr(), &schema, al.as_ptr(), &mut prog, &mut st) }, &st)?;
    Ok(GpuProgram(prog))
}

/// Filter and projection of one query fused into a single pass (identical results to
/// `project_record(filter_record(rec))`, errors on surviving rows only).
pub fn compile_filter_project(
    expr: &sqlparser::ast::Expr,
    fields: &Vec<sqlparser::ast::SelectItem>,
    rec: &RecordBatch,
    table_aliases: &Vec<Vec<String>>,
) -> Result<GpuProgram> {
    let (_, schema) = export(rec)?;
    let expr = CString::new(serde_json::to_string(expr)?)?;
    let items = CString::new(serde_json::to_string(fields)?)?;
    let al = aliases_json(table_aliases)?;
    let (mut prog, mut st) = (std::ptr::null_mut(), new_status());
    check(
        unsafe { ffi::chdb_program_compile_filter_project(expr.as_ptr(), items.as_ptr(), &schema, al.as_ptr(), &mut prog, &mut st) },
        &st,
    )?;
    Ok(GpuProgram(prog))
}

fn run(ctx: &GpuContext, prog: &GpuProgram, rec: &RecordBatch, project: bool) -> Result<RecordBatch> {
    let (array, schema) = export(rec)?;
    let mut out = FFI_ArrowArray::empty();
    let mut out_schema = FFI_ArrowSchema::empty();
    let mut st = new_status();
    let rc = unsafe {
        if project {
            ffi::chdb_project_record(ctx.0, prog.0, &array, &schema, &mut out, &mut out_schema, &mut st)
        } else {
            ffi::chdb_filter_record(ctx.0, prog.0, &array, &schema, &mut out, &mut out_schema, &mut st)
        }
    };
    check(rc, &st)?;
    import(out, &out_schema)
}

/// Drop-in for `record_utils::filter_record` (filter_record.rs:21-39). Blocks the calling thread
/// until the result is on the host, like the reference's inline call.
pub fn filter_record(ctx: &GpuContext, prog: &GpuProgram, rec: Arc<RecordBatch>) -> Result<RecordBatch> {
    run(ctx, prog, &rec, false)
}

/// Drop-in for `record_utils::project_record` (record_projection.rs:16-76).
pub fn project_record(ctx: &GpuContext, prog: &GpuProgram, rec: Arc<RecordBatch>) -> Result<RecordBatch> {
    run(ctx, prog, &rec, true)
}

/// The same without parking a tokio worker: upload + kernels are enqueued, then the task yields
/// between `chdb_poll` calls (an event query, never a stream synchronise).
pub async fn run_async(ctx: &GpuContext, prog: &GpuProgram, rec: Arc<RecordBatch>) -> Result<RecordBatch> {
    let (array, schema) = export(&rec)?; // `rec` (and `array`) stay alive until the result is taken
    let mut pending = std::ptr::null_mut();
    let mut st = new_status();
    check(unsafe { ffi::chdb_filter_record_async(ctx.0, prog.0, &array, &schema, &mut pending, &mut st) }, &st)?;
    struct Guard(*mut ffi::ChdbPending);
    impl Drop for Guard {
        fn drop(&mut self) {
            unsafe { ffi::chdb_pending_release(self.0) }
        }
    }
    let guard = Guard(pending);
    let mut wait = Duration::from_micros(20);
    loop {
        let r = unsafe { ffi::chdb_poll(guard.0, &mut st) };
        if r == 1 {
            break;
        }
        if r < 0 {
            check(-r, &st)?;
        }
        tokio::time::sleep(wait).await;
        wait = (wait * 2).min(Duration::from_micros(500));
    }
    let mut out = FFI_ArrowArray::empty();
    let mut out_schema = FFI_ArrowSchema::empty();
    check(unsafe { ffi::chdb_pending_result(guard.0, &mut out, &mut out_schema, &mut st) }, &st)?;
    import(out, &out_schema)
}

impl DeviceBatch {
    pub fn upload(ctx: &GpuContext, rec: &RecordBatch) -> Result<DeviceBatch> {
        let (array, schema) = export(rec)?;
        let (mut raw, mut st) = (std::ptr::null_mut(), new_status());
        check(unsafe { ffi::chdb_upload(ctx.0, &array, &schema, &mut raw, &mut st) }, &st)?;
        Ok(DeviceBatch { raw, ctx: ctx.0 })
    }
    /// Asynchronous on the ctx stream: the output handle can be pushed to the exchange at once.
    pub fn run(&self, ctx: &GpuContext, prog: &GpuProgram) -> Result<DeviceBatch> {
        let (mut raw, mut st) = (std::ptr::null_mut(), new_status());
        check(unsafe { ffi::chdb_run_device(ctx.0, prog.0, self.raw, &mut raw, &mut st) }, &st)?;
        Ok(DeviceBatch { raw, ctx: ctx.0 })
    }
    /// ONE launch set over many records (the reference's records are at most 10 000 rows).
    pub fn run_many(ctx: &GpuContext, prog: &GpuProgram, recs: &[&DeviceBatch]) -> Result<Vec<DeviceBatch>> {
        let ins: Vec<*const ffi::ChdbDeviceBatch> = recs.iter().map(|b| b.raw as *const _).collect();
        let mut outs: Vec<*mut ffi::ChdbDeviceBatch> = vec![std::ptr::null_mut(); recs.len()];
        let mut st = new_status();
        check(
            unsafe { ffi::chdb_run_device_many(ctx.0, prog.0, ins.as_ptr(), ins.len() as i32, outs.as_mut_ptr(), &mut st) },
            &st,
        )?;
        Ok(outs.into_iter().map(|raw| DeviceBatch { raw, ctx: ctx.0 }).collect())
    }
    pub fn ready(&self) -> Result<bool> {
        let mut st = new_status();
        let r = unsafe { ffi::chdb_device_batch_ready(self.ctx, self.raw, &mut st) };
        if r < 0 {
            check(st.code, &st)?;
        }
        Ok(r == 1)
    }
    pub fn num_rows(&self) -> Result<i64> {
        let mut st = new_status();
        let n = unsafe { ffi::chdb_device_batch_num_rows(self.ctx, self.raw, &mut st) };
        check(st.code, &st)?;
        Ok(n)
    }
    /// The single device -> host crossing of a query: at materialize, or when the consumer lives on another worker.
    pub fn download(&self) -> Result<RecordBatch> {
        let mut out = FFI_ArrowArray::empty();
        let mut out_schema = FFI_ArrowSchema::empty();
        let mut st = new_status();
        check(unsafe { ffi::chdb_download(self.ctx, self.raw, &mut out, &mut out_schema, &mut st) }, &st)?;
        import(out, &out_schema)
    }
    /// Materialize-side gather: copy this batch to another GPU of the box over NVLink.
    pub fn peer_copy(&self, src_ctx: &GpuContext, dst_ctx: &GpuContext) -> Result<DeviceBatch> {
        let (mut raw, mut st) = (std::ptr::null_mut(), new_status());
        check(unsafe { ffi::chdb_peer_copy(dst_ctx.0, src_ctx.0, self.raw, &mut raw, &mut st) }, &st)?;
        Ok(DeviceBatch { raw, ctx: dst_ctx.0 })
    }
}

/// A Parquet file image in pinned host memory written by the device encoder (`chdb_parquet_encode`): the GPU build's
/// materialize step (materialize_files_task.rs:116-141) with the record compaction of DEV_NOTES.md:117-122.
pub struct ParquetImage {
    ptr: *mut std::ffi::c_void,
    len: usize,
    /// how many of the batches handed to `encode_parquet` went into this file
    pub consumed: usize,
    pub row_groups: usize,
}
unsafe impl Send for ParquetImage {}
impl ParquetImage {
    pub fn as_bytes(&self) -> &[u8] {
        unsafe { std::slice::from_raw_parts(self.ptr as *const u8, self.len) }
    }
}
impl Drop for ParquetImage {
    fn drop(&mut self) {
        unsafe { ffi::chdb_parquet_image_free(self.ptr) }
    }
}

/// Device batches of one schema -> one Parquet file image; consecutive batches are coalesced into row groups of at most
/// `max_rows_per_row_group` rows (<= 0: 1 Mi), at most `max_row_groups` of them (<= 0: no limit; see `consumed`).
pub fn encode_parquet(ctx: &GpuContext, recs: &[&DeviceBatch], max_rows_per_row_group: i64, max_row_groups: i32) -> Result<ParquetImage> {
    let ins: Vec<*const ffi::ChdbDeviceBatch> = recs.iter().map(|b| b.raw as *const _).collect();
    let (mut ptr, mut len, mut consumed, mut groups, mut st) = (std::ptr::null_mut(), 0i64, 0i32, 0i32, new_status());
    check(
        unsafe {
            ffi::chdb_parquet_encode(ctx.0, ins.as_ptr(), ins.len() as i32, max_rows_per_row_group, max_row_groups, &mut ptr,
                                     &mut len, &mut consumed, &mut groups, &mut st)
        },
        &st,
    )?;
    Ok(ParquetImage { ptr, len: len as usize, consumed: consumed as usize, row_groups: groups as usize })
}

/// SQL nodes beyond the reference's compute_value (`-`, unary minus, NOT, IS [NOT] NULL: bit 0; Kleene AND / OR: bit 1); read
/// when a program is compiled, off by default.  Returns the previous mask.
pub fn set_sql_extensions(mask: u32) -> u32 {
    unsafe { ffi::chdb_set_sql_extensions(mask) }
}
