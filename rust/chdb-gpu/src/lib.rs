//! chdb-gpu: the B200 filter / projection path behind the signatures of
//! `record_utils::{filter_record, project_record}` (record_utils/mod.rs:13-15 of ChapterhouseDB).
//!
//! SOURCE ONLY -- not compiled in the build image (no Rust toolchain). The compiled, tested host
//! layer is the C++ library behind include/chdb_gpu.h; this file is the thin wrapper over it.
pub mod ffi;

use anyhow::{anyhow, Result};
use arrow::array::{Array, RecordBatch, StructArray};
use arrow::ffi::{from_ffi, to_ffi, FFI_ArrowArray, FFI_ArrowSchema};
use std::ffi::{CStr, CString};
use std::sync::Arc;

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
    let data = unsafe { from_ffi(out, &out_schema) }?;
    Ok(RecordBatch::from(StructArray::from(data)))
}

/// Drop-in for `record_utils::filter_record` (filter_record.rs:21-39).
pub fn filter_record(ctx: &GpuContext, prog: &GpuProgram, rec: Arc<RecordBatch>) -> Result<RecordBatch> {
    run(ctx, prog, &rec, false)
}

/// Drop-in for `record_utils::project_record` (record_projection.rs:16-76).
pub fn project_record(ctx: &GpuContext, prog: &GpuProgram, rec: Arc<RecordBatch>) -> Result<RecordBatch> {
    run(ctx, prog, &rec, true)
}
