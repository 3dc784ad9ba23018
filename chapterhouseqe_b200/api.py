"""ctypes binding of include/chdb_gpu.h + the reference-shaped Python entry points."""
from __future__ import annotations

import ctypes
import json
import os
from dataclasses import dataclass

import pyarrow as pa

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class ChdbError(RuntimeError):
    """An error returned through the C ABI. `.kind` is the reference's enum variant name
    (e.g. "ArithmeticOverflow", "ColumnNotFound"); `.code` the chdb_code."""

    def __init__(self, code: int, kind: str, message: str):
        super().__init__(f"{kind}: {message}")
        self.code, self.kind, self.message = code, kind, message


class _Status(ctypes.Structure):
    _fields_ = [("code", ctypes.c_int32), ("message", ctypes.c_char * 508)]


class _ArrowSchema(ctypes.Structure):
    pass


class _ArrowArray(ctypes.Structure):
    pass


_ArrowSchema._fields_ = [("format", ctypes.c_char_p), ("name", ctypes.c_char_p), ("metadata", ctypes.c_char_p),
                         ("flags", ctypes.c_int64), ("n_children", ctypes.c_int64),
                         ("children", ctypes.POINTER(ctypes.POINTER(_ArrowSchema))),
                         ("dictionary", ctypes.POINTER(_ArrowSchema)),
                         ("release", ctypes.CFUNCTYPE(None, ctypes.POINTER(_ArrowSchema))),
                         ("private_data", ctypes.c_void_p)]
_ArrowArray._fields_ = [("length", ctypes.c_int64), ("null_count", ctypes.c_int64), ("offset", ctypes.c_int64),
                        ("n_buffers", ctypes.c_int64), ("n_children", ctypes.c_int64),
                        ("buffers", ctypes.POINTER(ctypes.c_void_p)),
                        ("children", ctypes.POINTER(ctypes.POINTER(_ArrowArray))),
                        ("dictionary", ctypes.POINTER(_ArrowArray)),
                        ("release", ctypes.CFUNCTYPE(None, ctypes.POINTER(_ArrowArray))),
                        ("private_data", ctypes.c_void_p)]


def lib_path() -> str:
    # CHDB_LIB: an alternative build of the same library (kernel-shape experiments)
    return os.environ.get("CHDB_LIB") or os.path.join(_HERE, "libchdb_gpu.so")


def load_library():
    """Loads libchdb_gpu.so. Raises if it has not been built (python -m chapterhouseqe_b200.build):
    there is deliberately no fallback implementation."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -m chapterhouseqe_b200.build` "
                          "(nvcc, sm_100a). chapterhouseqe_b200 has no CPU fallback.")
    L = ctypes.CDLL(path)
    vp, cp, i32, i64 = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int32, ctypes.c_int64
    stp = ctypes.POINTER(_Status)
    pvp = ctypes.POINTER(ctypes.c_void_p)

    def sig(name, restype, *argtypes):
        f = getattr(L, name)
        f.restype, f.argtypes = restype, list(argtypes)

    sig("chdb_code_name", cp, i32)
    sig("chdb_version", cp)
    sig("chdb_compiled_arch", cp)
    sig("chdb_set_sql_extensions", ctypes.c_uint32, ctypes.c_uint32)
    sig("chdb_get_sql_extensions", ctypes.c_uint32)
    sig("chdb_ctx_create", i32, i32, pvp, stp)
    sig("chdb_ctx_destroy", None, vp)
    sig("chdb_ctx_stream", vp, vp)
    sig("chdb_ctx_device", i32, vp)
    sig("chdb_ctx_synchronize", i32, vp, stp)
    sig("chdb_ctx_launch_count", i64, vp)
    sig("chdb_ctx_jit_launch_count", i64, vp)
    sig("chdb_ctx_alloc_miss_count", i64, vp)
    sig("chdb_ctx_overlapped_count", i64, vp)
    sig("chdb_jit_available", i32, ctypes.c_char_p, ctypes.c_size_t)
    sig("chdb_program_jit_source", ctypes.c_size_t, vp, ctypes.c_char_p, ctypes.c_size_t)
    sig("chdb_program_jit_check", i32, vp, ctypes.POINTER(i64), ctypes.c_char_p, ctypes.c_size_t, stp)
    sig("chdb_program_compile_filter", i32, cp, vp, cp, pvp, stp)
    sig("chdb_program_compile_project", i32, cp, vp, cp, pvp, stp)
    sig("chdb_program_compile_filter_project", i32, cp, cp, vp, cp, pvp, stp)
    sig("chdb_program_release", None, vp)
    sig("chdb_program_disassemble", ctypes.c_size_t, vp, ctypes.c_char_p, ctypes.c_size_t)
    sig("chdb_program_num_instructions", i32, vp)
    sig("chdb_filter_record", i32, vp, vp, vp, vp, vp, vp, stp)
    sig("chdb_project_record", i32, vp, vp, vp, vp, vp, vp, stp)
    sig("chdb_filter_record_expr", i32, vp, vp, vp, cp, cp, vp, vp, stp)
    sig("chdb_project_record_items", i32, vp, cp, vp, vp, cp, vp, vp, stp)
    sig("chdb_compute_value", i32, vp, vp, vp, cp, cp, vp, vp, ctypes.POINTER(i32), stp)
    sig("chdb_upload", i32, vp, vp, vp, pvp, stp)
    sig("chdb_device_batch_wrap", i32, vp, vp, i64, pvp, pvp, pvp, pvp, stp)
    sig("chdb_run_device", i32, vp, vp, vp, pvp, stp)
    sig("chdb_run_device_many", i32, vp, vp, pvp, i32, pvp, stp)
    sig("chdb_device_batch_ready", i32, vp, vp, stp)
    sig("chdb_filter_record_async", i32, vp, vp, vp, vp, pvp, stp)
    sig("chdb_project_record_async", i32, vp, vp, vp, vp, pvp, stp)
    sig("chdb_poll", i32, vp, stp)
    sig("chdb_pending_result", i32, vp, vp, vp, stp)
    sig("chdb_pending_release", None, vp)
    sig("chdb_device_batch_status", i32, vp, vp, stp)
    sig("chdb_device_batch_num_rows", i64, vp, vp, stp)
    sig("chdb_device_batch_num_columns", i32, vp)
    sig("chdb_device_batch_nbytes", i64, vp, vp, stp)
    sig("chdb_device_batch_column", i32, vp, vp, i32, pvp, ctypes.POINTER(i64), pvp, ctypes.POINTER(i64), pvp,
        ctypes.POINTER(i64), stp)
    sig("chdb_download", i32, vp, vp, vp, vp, stp)
    sig("chdb_peer_copy", i32, vp, vp, vp, pvp, stp)
    sig("chdb_device_batches_pack", i32, vp, pvp, i32, vp, i64, ctypes.POINTER(i64), ctypes.POINTER(i64), stp)
    sig("chdb_device_batch_retain", None, vp)
    sig("chdb_device_batch_release", None, vp)
    sig("chdb_device_batch_release_many", None, pvp, i32)
    sig("chdb_record_pool_create", i32, vp, i64, pvp, stp)
    sig("chdb_record_pool_destroy", None, vp)
    sig("chdb_record_pool_add", i32, vp, ctypes.c_uint64, vp, i32, stp)
    sig("chdb_record_pool_get", i32, vp, ctypes.c_uint64, pvp, stp)
    sig("chdb_record_pool_complete", i32, vp, ctypes.c_uint64, stp)
    sig("chdb_record_pool_stats", None, vp, ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(i64))
    sig("chdb_parquet_open", i32, vp, i64, pvp, stp)
    sig("chdb_parquet_close", None, vp)
    sig("chdb_parquet_num_row_groups", i32, vp)
    sig("chdb_parquet_num_columns", i32, vp)
    sig("chdb_parquet_num_rows", i64, vp)
    sig("chdb_parquet_row_group_num_rows", i64, vp, i32)
    sig("chdb_parquet_column", i32, vp, i32, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(i32))
    sig("chdb_parquet_decode_row_group", i32, vp, vp, i32, pvp, stp)
    sig("chdb_parquet_decode_row_groups", i32, vp, vp, i32, i32, pvp, stp)
    sig("chdb_parquet_check_row_group", i32, vp, i32, ctypes.POINTER(i64), ctypes.POINTER(i64), stp)
    sig("chdb_parquet_encode", i32, vp, vp, i32, i64, i32, pvp, ctypes.POINTER(i64), ctypes.POINTER(i32), ctypes.POINTER(i32), stp)
    sig("chdb_parquet_image_free", None, vp)
    _LIB = L
    return L


EXPORTED_SYMBOLS = [
    "chdb_set_sql_extensions", "chdb_get_sql_extensions", "chdb_code_name", "chdb_version", "chdb_compiled_arch", "chdb_ctx_create", "chdb_ctx_destroy", "chdb_ctx_stream",
    "chdb_ctx_device", "chdb_ctx_synchronize", "chdb_ctx_launch_count", "chdb_ctx_jit_launch_count", "chdb_ctx_alloc_miss_count",
    "chdb_ctx_overlapped_count", "chdb_parquet_open", "chdb_parquet_close", "chdb_parquet_num_row_groups",
    "chdb_parquet_num_columns", "chdb_parquet_num_rows", "chdb_parquet_row_group_num_rows", "chdb_parquet_column",
    "chdb_parquet_decode_row_group", "chdb_parquet_check_row_group", "chdb_parquet_encode", "chdb_parquet_image_free",
    "chdb_parquet_decode_row_groups",
    "chdb_jit_available", "chdb_program_jit_source", "chdb_program_jit_check", "chdb_program_compile_filter",
    "chdb_program_compile_project", "chdb_program_compile_filter_project", "chdb_program_release",
    "chdb_program_disassemble", "chdb_program_num_instructions", "chdb_filter_record", "chdb_project_record",
    "chdb_filter_record_expr", "chdb_project_record_items", "chdb_compute_value", "chdb_upload",
    "chdb_device_batch_wrap", "chdb_run_device", "chdb_run_device_many", "chdb_device_batch_ready",
    "chdb_filter_record_async", "chdb_project_record_async", "chdb_poll", "chdb_pending_result", "chdb_pending_release",
    "chdb_device_batch_status", "chdb_device_batch_num_rows",
    "chdb_device_batch_num_columns", "chdb_device_batch_column", "chdb_device_batch_nbytes", "chdb_download",
    "chdb_peer_copy", "chdb_device_batches_pack", "chdb_device_batch_retain",
    "chdb_device_batch_release", "chdb_device_batch_release_many", "chdb_record_pool_create", "chdb_record_pool_destroy", "chdb_record_pool_add",
    "chdb_record_pool_get", "chdb_record_pool_complete", "chdb_record_pool_stats",
]


def jit_available() -> tuple[bool, str]:
    buf = ctypes.create_string_buffer(512)
    ok = load_library().chdb_jit_available(buf, len(buf))
    return bool(ok), buf.value.decode()


def _check(rc: int, st: _Status):
    if rc != 0:
        L = load_library()
        raise ChdbError(rc, L.chdb_code_name(rc).decode(), st.message.decode(errors="replace"))


def _json(x) -> bytes | None:
    if x is None:
        return None
    if isinstance(x, bytes):
        return x
    if isinstance(x, str):
        return x.encode()
    return json.dumps(x, separators=(",", ":")).encode()


def _export_schema(schema: pa.Schema) -> _ArrowSchema:
    c = _ArrowSchema()
    schema._export_to_c(ctypes.addressof(c))
    return c


def _release_schema(c: _ArrowSchema):
    if c.release:
        c.release(ctypes.byref(c))


def _export_batch(rb: pa.RecordBatch):
    a, s = _ArrowArray(), _ArrowSchema()
    rb._export_to_c(ctypes.addressof(a), ctypes.addressof(s))
    return a, s


def _import_batch(a: _ArrowArray, s: _ArrowSchema) -> pa.RecordBatch:
    return pa.RecordBatch._import_from_c(ctypes.addressof(a), ctypes.addressof(s))


# ---------------------------------------------------------------------------------------------
class Context:
    """One operator instance's GPU context (one device, one stream)."""

    def __init__(self, device: int = 0):
        L = load_library()
        h, st = ctypes.c_void_p(), _Status()
        _check(L.chdb_ctx_create(device, ctypes.byref(h), ctypes.byref(st)), st)
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            load_library().chdb_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    @property
    def stream(self) -> int:
        """cudaStream_t as an integer (wrap with torch.cuda.ExternalStream to record events)."""
        return int(load_library().chdb_ctx_stream(self._h) or 0)

    @property
    def launch_count(self) -> int:
        return int(load_library().chdb_ctx_launch_count(self._h))

    @property
    def jit_launch_count(self) -> int:
        """Launches that ran an NVRTC-specialised kernel (subset of launch_count)."""
        return int(load_library().chdb_ctx_jit_launch_count(self._h))

    @property
    def alloc_misses(self) -> int:
        """Device allocations that missed the block cache (cudaMallocAsync calls)."""
        return int(load_library().chdb_ctx_alloc_miss_count(self._h))

    @property
    def overlapped_launch_sets(self) -> int:
        """Launch sets whose select kernel ran on the second stream, next to the previous set's gather kernel."""
        return int(load_library().chdb_ctx_overlapped_count(self._h))

    def synchronize(self):
        st = _Status()
        _check(load_library().chdb_ctx_synchronize(self._h, ctypes.byref(st)), st)


_DEFAULT_CTX: dict[int, Context] = {}


EXT_OPERATORS, EXT_KLEENE = 1, 2


class sql_extensions:
    """`with sql_extensions(EXT_OPERATORS | EXT_KLEENE): prog = Program.compile_filter(...)` -- nodes beyond the reference's
    compute_value (chdb_set_sql_extensions); the mask is read when a program is compiled and restored on exit."""

    def __init__(self, mask: int):
        self.mask = mask

    def __enter__(self):
        self.prev = int(load_library().chdb_set_sql_extensions(self.mask))
        return self

    def __exit__(self, *exc):
        load_library().chdb_set_sql_extensions(self.prev)
        return False


def default_context(device: int = 0) -> Context:
    if device not in _DEFAULT_CTX:
        _DEFAULT_CTX[device] = Context(device)
    return _DEFAULT_CTX[device]


class Program:
    """An expression tree lowered to bytecode for one input schema (no GPU needed to compile)."""

    def __init__(self, handle, mode: str):
        self._h, self.mode = handle, mode

    @staticmethod
    def _compile(mode, expr, items, schema: pa.Schema, table_aliases) -> "Program":
        L = load_library()
        cs = _export_schema(schema)
        # a RecordBatch schema crosses as a struct-typed ArrowSchema
        h, st = ctypes.c_void_p(), _Status()
        try:
            al = _json(table_aliases)
            if mode == "filter":
                rc = L.chdb_program_compile_filter(_json(expr), ctypes.addressof(cs), al, ctypes.byref(h), ctypes.byref(st))
            elif mode == "project":
                rc = L.chdb_program_compile_project(_json(items), ctypes.addressof(cs), al, ctypes.byref(h), ctypes.byref(st))
            else:
                rc = L.chdb_program_compile_filter_project(_json(expr), _json(items), ctypes.addressof(cs), al,
                                                           ctypes.byref(h), ctypes.byref(st))
        finally:
            _release_schema(cs)
        _check(rc, st)
        return Program(h, mode)

    @staticmethod
    def compile_filter(expr, schema: pa.Schema, table_aliases=None) -> "Program":
        return Program._compile("filter", expr, None, schema, table_aliases)

    @staticmethod
    def compile_project(select_items, schema: pa.Schema, table_aliases=None) -> "Program":
        return Program._compile("project", None, select_items, schema, table_aliases)

    @staticmethod
    def compile_filter_project(expr, select_items, schema: pa.Schema, table_aliases=None) -> "Program":
        return Program._compile("filter_project", expr, select_items, schema, table_aliases)

    def disassemble(self) -> str:
        L = load_library()
        n = L.chdb_program_disassemble(self._h, None, 0)
        buf = ctypes.create_string_buffer(n + 1)
        L.chdb_program_disassemble(self._h, buf, n + 1)
        return buf.value.decode()

    @property
    def num_instructions(self) -> int:
        return int(load_library().chdb_program_num_instructions(self._h))

    def jit_source(self) -> str:
        """The constants NVRTC sees in front of device_code.cuh when this program is specialised."""
        L = load_library()
        n = L.chdb_program_jit_source(self._h, None, 0)
        buf = ctypes.create_string_buffer(n + 1)
        L.chdb_program_jit_source(self._h, buf, n + 1)
        return buf.value.decode()

    def jit_check(self) -> tuple[int, str]:
        """Compile the specialised kernel for sm_100a (no GPU needed). Returns (cubin bytes, log)."""
        L = load_library()
        n, st = ctypes.c_int64(0), _Status()
        log = ctypes.create_string_buffer(1 << 16)
        _check(L.chdb_program_jit_check(self._h, ctypes.byref(n), log, len(log), ctypes.byref(st)), st)
        return int(n.value), log.value.decode(errors="replace")

    def run(self, rb: pa.RecordBatch, ctx: Context | None = None) -> pa.RecordBatch:
        """Host batch in, host batch out (upload -> select, scan, gather kernels -> download)."""
        L = load_library()
        ctx = ctx or default_context()
        a, s = _export_batch(rb)
        oa, os_, st = _ArrowArray(), _ArrowSchema(), _Status()
        try:
            rc = L.chdb_filter_record(ctx._h, self._h, ctypes.addressof(a), ctypes.addressof(s), ctypes.addressof(oa),
                                      ctypes.addressof(os_), ctypes.byref(st))
        finally:
            if a.release:
                a.release(ctypes.byref(a))
            _release_schema(s)
        _check(rc, st)
        return _import_batch(oa, os_)

    def run_async(self, rb: pa.RecordBatch, ctx: Context | None = None) -> "Pending":
        """Enqueues upload + kernels and returns at once (chdb_filter_record_async); poll() / result() on the handle."""
        L = load_library()
        ctx = ctx or default_context()
        a, s = _export_batch(rb)
        h, st = ctypes.c_void_p(), _Status()
        try:
            rc = L.chdb_filter_record_async(ctx._h, self._h, ctypes.addressof(a), ctypes.addressof(s), ctypes.byref(h),
                                            ctypes.byref(st))
        except BaseException:
            if a.release:
                a.release(ctypes.byref(a))
            _release_schema(s)
            raise
        if rc != 0:
            if a.release:
                a.release(ctypes.byref(a))
            _release_schema(s)
        _check(rc, st)
        return Pending(h, a, s, rb)

    def close(self):
        if getattr(self, "_h", None):
            load_library().chdb_program_release(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class Pending:
    """A host-batch call in flight (chdb_filter_record_async): the input batch stays borrowed until poll() is true."""

    def __init__(self, handle, arr, sch, keep):
        self._h, self._a, self._s, self._keep = handle, arr, sch, keep

    def poll(self) -> bool:
        """Never blocks. True once the result can be taken."""
        st = _Status()
        r = load_library().chdb_poll(self._h, ctypes.byref(st))
        if r < 0:
            _check(-r, st)
        return r == 1

    def result(self) -> pa.RecordBatch:
        L = load_library()
        oa, os_, st = _ArrowArray(), _ArrowSchema(), _Status()
        rc = L.chdb_pending_result(self._h, ctypes.addressof(oa), ctypes.addressof(os_), ctypes.byref(st))
        _check(rc, st)
        out = _import_batch(oa, os_)
        self.close()
        return out

    def close(self):
        if getattr(self, "_h", None):
            load_library().chdb_pending_release(self._h)
            self._h = None
            if self._a.release:
                self._a.release(ctypes.byref(self._a))
            _release_schema(self._s)
            self._keep = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class ParquetFile:
    """A Parquet file's bytes, decoded row group by row group ON THE DEVICE (chdb_parquet_*): the GPU build's
    read_files (table_func_tasks/read_files_task.rs:233-282).  Opening parses the footer on the host and never
    touches the GPU; `data` (bytes / bytearray / memoryview / numpy uint8 array) stays borrowed."""

    _FORMATS = {"b": pa.bool_(), "c": pa.int8(), "s": pa.int16(), "i": pa.int32(), "l": pa.int64(), "C": pa.uint8(),
                "S": pa.uint16(), "I": pa.uint32(), "L": pa.uint64(), "f": pa.float32(), "g": pa.float64(), "u": pa.utf8()}

    def __init__(self, data):
        L = load_library()
        self._view = memoryview(data).cast("B")
        self._buf = (ctypes.c_char * len(self._view)).from_buffer_copy(self._view) if self._view.readonly else \
            (ctypes.c_char * len(self._view)).from_buffer(self._view)
        h, st = ctypes.c_void_p(), _Status()
        _check(L.chdb_parquet_open(ctypes.addressof(self._buf), len(self._view), ctypes.byref(h), ctypes.byref(st)), st)
        self._h = h

    @property
    def num_row_groups(self) -> int:
        return int(load_library().chdb_parquet_num_row_groups(self._h))

    @property
    def num_rows(self) -> int:
        return int(load_library().chdb_parquet_num_rows(self._h))

    def row_group_num_rows(self, i: int) -> int:
        return int(load_library().chdb_parquet_row_group_num_rows(self._h, i))

    @property
    def schema(self) -> pa.Schema:
        L = load_library()
        fields = []
        for c in range(int(L.chdb_parquet_num_columns(self._h))):
            name, fmt, nullable = ctypes.c_char_p(), ctypes.c_char_p(), ctypes.c_int32()
            L.chdb_parquet_column(self._h, c, ctypes.byref(name), ctypes.byref(fmt), ctypes.byref(nullable))
            fields.append(pa.field(name.value.decode(), self._FORMATS[fmt.value.decode()], bool(nullable.value)))
        return pa.schema(fields)

    def check_row_group(self, i: int) -> tuple[int, int]:
        """Host-only: (data pages, hybrid runs) of row group i; raises if the decoder does not support it."""
        pages, runs, st = ctypes.c_int64(), ctypes.c_int64(), _Status()
        _check(load_library().chdb_parquet_check_row_group(self._h, i, ctypes.byref(pages), ctypes.byref(runs), ctypes.byref(st)), st)
        return int(pages.value), int(runs.value)

    def decode_row_group(self, i: int, ctx: "Context | None" = None) -> "DeviceBatch":
        L = load_library()
        ctx = ctx or default_context()
        h, st = ctypes.c_void_p(), _Status()
        _check(L.chdb_parquet_decode_row_group(ctx._h, self._h, i, ctypes.byref(h), ctypes.byref(st)), st)
        return DeviceBatch(h, ctx)

    def decode_row_groups(self, first: int = 0, count: "int | None" = None, ctx: "Context | None" = None) -> "list[DeviceBatch]":
        """Row groups [first, first + count): the next row group's bytes cross PCIe while this one's kernels run."""
        L = load_library()
        ctx = ctx or default_context()
        count = self.num_row_groups - first if count is None else count
        outs = (ctypes.c_void_p * max(count, 1))()
        st = _Status()
        _check(L.chdb_parquet_decode_row_groups(ctx._h, self._h, first, count, outs, ctypes.byref(st)), st)
        return [DeviceBatch(ctypes.c_void_p(outs[i]), ctx) for i in range(count)]

    def close(self):
        if getattr(self, "_h", None):
            load_library().chdb_parquet_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class DeviceBatch:
    """A device-resident RecordBatch: what the exchanges carry between read_files, filter and
    materialize in the GPU build."""

    def __init__(self, handle, ctx: Context, keepalive=None):
        self._h, self.ctx, self._keepalive = handle, ctx, keepalive

    @staticmethod
    def upload(rb: pa.RecordBatch, ctx: Context | None = None) -> "DeviceBatch":
        L = load_library()
        ctx = ctx or default_context()
        a, s = _export_batch(rb)
        h, st = ctypes.c_void_p(), _Status()
        try:
            rc = L.chdb_upload(ctx._h, ctypes.addressof(a), ctypes.addressof(s), ctypes.byref(h), ctypes.byref(st))
        finally:
            if a.release:
                a.release(ctypes.byref(a))
            _release_schema(s)
        _check(rc, st)
        return DeviceBatch(h, ctx)

    @staticmethod
    def wrap(schema: pa.Schema, num_rows: int, values, validity=None, offsets=None, ctx: Context | None = None,
             keepalive=None) -> "DeviceBatch":
        """Wrap device pointers (ints) the caller owns, e.g. torch tensors' data_ptr(). No copy."""
        L = load_library()
        ctx = ctx or default_context()
        n = len(schema)
        arr = ctypes.c_void_p * n
        vals = arr(*[ctypes.c_void_p(v) for v in values])
        vald = arr(*[ctypes.c_void_p(v or None) for v in (validity or [None] * n)])
        offs = arr(*[ctypes.c_void_p(v or None) for v in (offsets or [None] * n)])
        cs = _export_schema(schema)
        h, st = ctypes.c_void_p(), _Status()
        try:
            rc = L.chdb_device_batch_wrap(ctx._h, ctypes.addressof(cs), num_rows, vals, vald, offs, ctypes.byref(h),
                                          ctypes.byref(st))
        finally:
            _release_schema(cs)
        _check(rc, st)
        return DeviceBatch(h, ctx, keepalive)

    def run(self, prog: Program) -> "DeviceBatch":
        """Asynchronous on the ctx stream."""
        L = load_library()
        h, st = ctypes.c_void_p(), _Status()
        _check(L.chdb_run_device(self.ctx._h, prog._h, self._h, ctypes.byref(h), ctypes.byref(st)), st)
        return DeviceBatch(h, self.ctx, (self, self._keepalive))

    @staticmethod
    def run_many(prog: Program, batches) -> "DeviceBatchList":
        """ONE launch set over many batches of one schema (chdb_run_device_many); one output batch per input.
        `batches`: a list of DeviceBatch, or a DeviceBatchList (its handle array is passed as is: no per-record work
        on the Python side, like a Rust caller handing over a slice of pointers)."""
        if not isinstance(batches, DeviceBatchList):
            batches = DeviceBatchList.from_batches(batches)
        n = len(batches)
        if n == 0:
            return DeviceBatchList((ctypes.c_void_p * 0)(), 0, batches.ctx, None)
        L = load_library()
        outs = (ctypes.c_void_p * n)()
        st = _Status()
        _check(L.chdb_run_device_many(batches.ctx._h, prog._h, batches._arr, n, outs, ctypes.byref(st)), st)
        return DeviceBatchList(outs, n, batches.ctx, batches)

    @property
    def ready(self) -> bool:
        """Whether the run producing this batch has finished (never blocks)."""
        st = _Status()
        r = load_library().chdb_device_batch_ready(self.ctx._h, self._h, ctypes.byref(st))
        _check(st.code, st)
        return r == 1

    def check(self):
        st = _Status()
        _check(load_library().chdb_device_batch_status(self.ctx._h, self._h, ctypes.byref(st)), st)

    @property
    def num_rows(self) -> int:
        st = _Status()
        n = load_library().chdb_device_batch_num_rows(self.ctx._h, self._h, ctypes.byref(st))
        _check(st.code, st)
        return int(n)

    @property
    def num_columns(self) -> int:
        return int(load_library().chdb_device_batch_num_columns(self._h))

    @property
    def nbytes(self) -> int:
        st = _Status()
        n = load_library().chdb_device_batch_nbytes(self.ctx._h, self._h, ctypes.byref(st))
        _check(st.code, st)
        return int(n)

    def column_buffers(self, col: int) -> dict:
        """Device pointers / sizes of one column: {"values": (ptr, nbytes), "validity": .., "offsets": ..}."""
        L = load_library()
        v, va, o = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        nv, nva, no = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        st = _Status()
        _check(L.chdb_device_batch_column(self.ctx._h, self._h, col, ctypes.byref(v), ctypes.byref(nv), ctypes.byref(va),
                                          ctypes.byref(nva), ctypes.byref(o), ctypes.byref(no), ctypes.byref(st)), st)
        return {"values": (v.value or 0, nv.value), "validity": (va.value or 0, nva.value), "offsets": (o.value or 0, no.value)}

    def download(self) -> pa.RecordBatch:
        L = load_library()
        oa, os_, st = _ArrowArray(), _ArrowSchema(), _Status()
        _check(L.chdb_download(self.ctx._h, self._h, ctypes.addressof(oa), ctypes.addressof(os_), ctypes.byref(st)), st)
        return _import_batch(oa, os_)

    def peer_copy(self, dst_ctx: Context) -> "DeviceBatch":
        L = load_library()
        h, st = ctypes.c_void_p(), _Status()
        _check(L.chdb_peer_copy(dst_ctx._h, self.ctx._h, self._h, ctypes.byref(h), ctypes.byref(st)), st)
        return DeviceBatch(h, dst_ctx)

    def close(self):
        if getattr(self, "_h", None):
            load_library().chdb_device_batch_release(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class DeviceBatchList:
    """n device batches behind one ctypes array of handles; a DeviceBatch wrapper is only made for the ones asked for."""

    def __init__(self, arr, n: int, ctx: Context, keepalive=None):
        self._arr, self._n, self.ctx, self._keepalive = arr, n, ctx, keepalive

    @staticmethod
    def from_batches(batches) -> "DeviceBatchList":
        batches = list(batches)
        ctx = batches[0].ctx if batches else default_context()
        L = load_library()
        arr = (ctypes.c_void_p * len(batches))(*[b._h for b in batches])
        for b in batches:
            L.chdb_device_batch_retain(b._h)
        return DeviceBatchList(arr, len(batches), ctx, batches)

    def __len__(self) -> int:
        return self._n

    def __getitem__(self, i: int) -> DeviceBatch:
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        h = ctypes.c_void_p(self._arr[i])
        load_library().chdb_device_batch_retain(h)
        return DeviceBatch(h, self.ctx, self._keepalive)

    def __iter__(self):
        return (self[i] for i in range(self._n))

    def pack(self, dst_ptr: int = 0, capacity: int = 0):
        """chdb_device_batches_pack: every buffer of every batch into one contiguous device buffer (asynchronous on
        the ctx stream).  Returns (total bytes, sizes) -- sizes: int64[n * columns * 3]; dst_ptr == 0 only sizes them."""
        L = load_library()
        ncols = int(L.chdb_device_batch_num_columns(ctypes.c_void_p(self._arr[0]))) if self._n else 0
        sizes = (ctypes.c_int64 * max(self._n * ncols * 3, 1))()
        total, st = ctypes.c_int64(0), _Status()
        _check(L.chdb_device_batches_pack(self.ctx._h, self._arr, self._n, ctypes.c_void_p(dst_ptr or None), capacity, sizes,
                                          ctypes.byref(total), ctypes.byref(st)), st)
        return int(total.value), sizes

    def close(self):
        if getattr(self, "_arr", None) is not None:
            load_library().chdb_device_batch_release_many(self._arr, self._n)
            self._arr = None
            self._n = 0

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


class ParquetImage:
    """A Parquet file image in pinned host memory, written by the device encoder (chdb_parquet_encode): the GPU build's
    materialize (materialize_files_task.rs:116-141) with the record compaction of DEV_NOTES.md:117-122.
    `view` is a zero-copy numpy uint8 view valid until close(); `consumed` = batches that went into this file."""

    def __init__(self, ptr: int, nbytes: int, consumed: int, row_groups: int):
        self._ptr, self.nbytes, self.consumed, self.row_groups = ptr, nbytes, consumed, row_groups

    @property
    def view(self):
        import numpy as np
        return np.ctypeslib.as_array((ctypes.c_uint8 * self.nbytes).from_address(self._ptr))

    def to_bytes(self) -> bytes:
        return ctypes.string_at(self._ptr, self.nbytes)

    def close(self):
        if getattr(self, "_ptr", None):
            load_library().chdb_parquet_image_free(ctypes.c_void_p(self._ptr))
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def encode_parquet(batches, max_rows_per_row_group: int = 0, max_row_groups: int = 0, ctx: "Context | None" = None) -> ParquetImage:
    """Device batches of one schema -> one Parquet file image; consecutive batches are coalesced into row groups of at most
    `max_rows_per_row_group` rows (0: 1 Mi), at most `max_row_groups` of them (0: no limit; see `.consumed`)."""
    L = load_library()
    if isinstance(batches, DeviceBatchList):
        arr, n, ctx = batches._arr, len(batches), ctx or batches.ctx
    else:
        batches = list(batches)
        arr, n = (ctypes.c_void_p * max(len(batches), 1))(*[b._h for b in batches]), len(batches)
        ctx = ctx or (batches[0].ctx if batches else default_context())
    ptr, nbytes, consumed, groups, st = ctypes.c_void_p(), ctypes.c_int64(), ctypes.c_int32(), ctypes.c_int32(), _Status()
    _check(L.chdb_parquet_encode(ctx._h, arr, n, max_rows_per_row_group, max_row_groups, ctypes.byref(ptr), ctypes.byref(nbytes),
                                 ctypes.byref(consumed), ctypes.byref(groups), ctypes.byref(st)), st)
    return ParquetImage(ptr.value, int(nbytes.value), int(consumed.value), int(groups.value))


class RecordPool:
    """Device-resident RecordPool of a GPU-aware exchange (exchange_operator.rs:566-777): records by reference,
    dropped after the last consumer operator completes them, spilled to pinned host memory beyond the budget."""

    def __init__(self, ctx: Context | None = None, budget_bytes: int = 0):
        L = load_library()
        self.ctx = ctx or default_context()
        h, st = ctypes.c_void_p(), _Status()
        _check(L.chdb_record_pool_create(self.ctx._h, budget_bytes, ctypes.byref(h), ctypes.byref(st)), st)
        self._h = h

    def add(self, record_id: int, batch: DeviceBatch, consumers: int = 1):
        st = _Status()
        _check(load_library().chdb_record_pool_add(self._h, record_id, batch._h, consumers, ctypes.byref(st)), st)

    def get(self, record_id: int) -> DeviceBatch:
        h, st = ctypes.c_void_p(), _Status()
        _check(load_library().chdb_record_pool_get(self._h, record_id, ctypes.byref(h), ctypes.byref(st)), st)
        return DeviceBatch(h, self.ctx)

    def complete(self, record_id: int):
        st = _Status()
        _check(load_library().chdb_record_pool_complete(self._h, record_id, ctypes.byref(st)), st)

    def stats(self) -> dict:
        vals = [ctypes.c_int64() for _ in range(4)]
        load_library().chdb_record_pool_stats(self._h, *[ctypes.byref(v) for v in vals])
        return dict(zip(("records", "device_bytes", "spilled_records", "spilled_bytes"), [v.value for v in vals]))

    def close(self):
        if getattr(self, "_h", None):
            load_library().chdb_record_pool_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


# ---------------------------------------------------------------------------------------------
# the reference's three functions
# ---------------------------------------------------------------------------------------------
@dataclass
class ArrayDatum:
    """compute_value.rs:34-55"""
    array: pa.Array
    is_scalar: bool


def _host_call(fn_name: str, rb: pa.RecordBatch, ctx: Context | None, *mid_args):
    L = load_library()
    ctx = ctx or default_context()
    a, s = _export_batch(rb)
    oa, os_, st = _ArrowArray(), _ArrowSchema(), _Status()
    try:
        yield_args = (ctx, a, s, oa, os_, st)
        rc = mid_args[0](L, *yield_args)
    finally:
        if a.release:
            a.release(ctypes.byref(a))
        _release_schema(s)
    _check(rc, st)
    return _import_batch(oa, os_)


def filter_record(rec: pa.RecordBatch, table_aliases, expr, ctx: Context | None = None) -> pa.RecordBatch:
    """filter_record.rs:21-39 -- `expr` is the serde-JSON form of sqlparser's Expr (dict or str)."""
    def call(L, ctx, a, s, oa, os_, st):
        return L.chdb_filter_record_expr(ctx._h, ctypes.addressof(a), ctypes.addressof(s), _json(table_aliases), _json(expr),
                                         ctypes.addressof(oa), ctypes.addressof(os_), ctypes.byref(st))
    return _host_call("filter", rec, ctx, call)


def project_record(fields, record: pa.RecordBatch, table_aliases, ctx: Context | None = None) -> pa.RecordBatch:
    """record_projection.rs:16-76 -- `fields` is the serde-JSON form of Vec<SelectItem>."""
    def call(L, ctx, a, s, oa, os_, st):
        return L.chdb_project_record_items(ctx._h, _json(fields), ctypes.addressof(a), ctypes.addressof(s),
                                           _json(table_aliases), ctypes.addressof(oa), ctypes.addressof(os_),
                                           ctypes.byref(st))
    return _host_call("project", record, ctx, call)


def compute_value(rec: pa.RecordBatch, table_aliases, expr, ctx: Context | None = None) -> ArrayDatum:
    """compute_value.rs:57-344"""
    flag = ctypes.c_int32(0)

    def call(L, ctx, a, s, oa, os_, st):
        return L.chdb_compute_value(ctx._h, ctypes.addressof(a), ctypes.addressof(s), _json(table_aliases), _json(expr),
                                    ctypes.addressof(oa), ctypes.addressof(os_), ctypes.byref(flag), ctypes.byref(st))
    out = _host_call("value", rec, ctx, call)
    return ArrayDatum(out.column(0), bool(flag.value))


def filter_project_record(expr, fields, rec: pa.RecordBatch, table_aliases, ctx: Context | None = None) -> pa.RecordBatch:
    """project_record(fields, filter_record(rec, aliases, expr), aliases) as ONE fused pass."""
    prog = Program.compile_filter_project(expr, fields, rec.schema, table_aliases)
    try:
        return prog.run(rec, ctx)
    finally:
        prog.close()


def get_record_table_aliases(op_type: dict, record: pa.RecordBatch) -> list[list[str]]:
    """record_aliases.rs:12-59 -- op_type is the serde form of planner::OperatorType."""
    (_, body), = op_type.items()
    (task_name, task), = body["task"].items()
    if task_name not in ("TableFunc", "Table"):
        raise ChdbError(11, "OperatorTaskTypeDoesNotHaveAnAliasField",
                        f"operator task type does not have an alias field: OperatorTask::{task_name}")
    alias = task.get("alias")
    return [[alias] if alias is not None else [] for _ in range(record.num_columns)]
