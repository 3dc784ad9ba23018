//! Raw bindings of include/chdb_gpu.h (only what the filter / materialize tasks need).
use arrow::ffi::{FFI_ArrowArray, FFI_ArrowSchema};
use std::os::raw::c_char;

#[repr(C)]
pub struct ChdbStatus {
    pub code: i32,
    pub message: [c_char; 508],
}

#[repr(C)]
pub struct ChdbCtx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct ChdbProgram {
    _private: [u8; 0],
}

extern "C" {
    pub fn chdb_code_name(code: i32) -> *const c_char;
    pub fn chdb_ctx_create(device: i32, out: *mut *mut ChdbCtx, st: *mut ChdbStatus) -> i32;
    pub fn chdb_ctx_destroy(ctx: *mut ChdbCtx);
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
    pub fn chdb_program_release(prog: *mut ChdbProgram);
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
}
