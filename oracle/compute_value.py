"""oracle/compute_value.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's record_utils hot path (paths relative to
/root/reference/src/handlers/operator_handler/operators/record_utils/):

  compute_value.rs:57-344    -> compute_value()      (tree walk, one materialised array per node)
  compute_value.rs:219-265   -> _literal()           (f32-first / i32-first literal typing)
  compute_value.rs:350-431   -> get_common_type()    (coercion lattice)
  compute_value.rs:433-461   -> cast_to_common_type()
  filter_record.rs:21-39     -> filter_record()
  record_projection.rs:16-76 -> project_record()
  record_aliases.rs:12-59    -> get_record_table_aliases()

The per-array arithmetic the reference delegates to the un-vendored `arrow`
crate (53.x) lives in oracle/arrow_kernels.c; this module is the Rust-side
control flow around it: recursion, scalar tracking (ArrayDatum.is_scalar),
type coercion, error kinds.  Expression trees are the serde-JSON form of
sqlparser 0.52 `Expr` / `SelectItem` (what the reference ships between workers,
handlers/message_handler/messages/query.rs:424-431).

Pinned against the reference's own golden vectors in tests/test_oracle_golden.py.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

import ctypes
import os
import re
import subprocess
from fractions import Fraction

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# --------------------------------------------------------------------------
# C kernels
# --------------------------------------------------------------------------
TYPE_ID = {"bool": 0, "int8": 1, "int16": 2, "int32": 3, "int64": 4, "uint8": 5, "uint16": 6,
           "uint32": 7, "uint64": 8, "float32": 9, "float64": 10, "utf8": 11}
NP_DTYPE = {"int8": np.int8, "int16": np.int16, "int32": np.int32, "int64": np.int64,
            "uint8": np.uint8, "uint16": np.uint16, "uint32": np.uint32, "uint64": np.uint64,
            "float32": np.float32, "float64": np.float64}
ARITH = {"Plus": 0, "Multiply": 1, "Divide": 2, "Modulo": 3}
# SQL nodes beyond the reference's compute_value, mirroring chdb_set_sql_extensions: off by default, so the
# oracle raises the reference's errors for them.  Semantics = the arrow-rs kernels a Rust implementation would call
# (numeric::sub / neg, boolean::not, is_null / is_not_null, and_kleene / or_kleene).  Parity for these is UNPINNED by the
# reference (it has no such code); the checker beside this restatement is pyarrow.compute in tests/test_gpu_extensions.py.
EXT_OPERATORS, EXT_KLEENE = 1, 2
EXTENSIONS = 0


class extensions:
    def __init__(self, mask):
        self.mask = mask

    def __enter__(self):
        global EXTENSIONS
        self.prev, EXTENSIONS = EXTENSIONS, self.mask
        return self

    def __exit__(self, *exc):
        global EXTENSIONS
        EXTENSIONS = self.prev
        return False
CMP = {"Eq": 0, "NotEq": 1, "Lt": 2, "LtEq": 3, "Gt": 4, "GtEq": 5}
# arrow DataType Display names, used in error messages like the reference's
ARROW_NAME = {"bool": "Boolean", "int8": "Int8", "int16": "Int16", "int32": "Int32", "int64": "Int64",
              "uint8": "UInt8", "uint16": "UInt16", "uint32": "UInt32", "uint64": "UInt64",
              "float32": "Float32", "float64": "Float64", "utf8": "Utf8"}

_lib = None


def build_lib(march: str = "x86-64-v3", out: str | None = None) -> str:
    out = out or os.path.join(_HERE, "libchdb_oracle.so")
    cmd = ["gcc", "-O3", f"-march={march}", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-std=c11",
           "-shared", "-o", out, os.path.join(_HERE, "arrow_kernels.c"), "-lm"]
    subprocess.check_call(cmd)
    return out


def lib(path: str | None = None):
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.path.join(_HERE, "libchdb_oracle.so")
    src = os.path.join(_HERE, "arrow_kernels.c")
    if not os.path.exists(p) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(p)):
        build_lib(out=p)
    L = ctypes.CDLL(p)
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int
    L.ora_popcount.restype = i64
    L.ora_popcount.argtypes = [i64, vp]
    L.ora_bitmap_and.argtypes = [i64, vp, vp, vp]
    L.ora_bitmap_or.argtypes = [i64, vp, vp, vp]
    L.ora_arith.restype = i32
    L.ora_arith.argtypes = [i32, i32, i64, vp, i32, vp, i32, vp, vp, ctypes.POINTER(i64)]
    L.ora_cmp.restype = i32
    L.ora_cmp.argtypes = [i32, i32, i64, vp, i32, vp, i32, vp]
    L.ora_cmp_bool.restype = i32
    L.ora_cmp_bool.argtypes = [i32, i64, vp, i32, vp, i32, vp]
    L.ora_cmp_utf8.restype = i32
    L.ora_cmp_utf8.argtypes = [i32, i64, vp, vp, i32, vp, vp, i32, vp]
    L.ora_cast.restype = i32
    L.ora_cast.argtypes = [i32, i32, i64, vp, vp]
    L.ora_cast_to_bool.restype = i32
    L.ora_cast_to_bool.argtypes = [i32, i64, vp, vp]
    L.ora_filter_fixed.restype = i64
    L.ora_filter_fixed.argtypes = [i32, i64, vp, vp, vp]
    L.ora_filter_bits.restype = i64
    L.ora_filter_bits.argtypes = [i64, vp, vp, vp]
    L.ora_filter_utf8.restype = i64
    L.ora_filter_utf8.argtypes = [i64, vp, vp, vp, vp, vp]
    if path is None:
        _lib = L
    return L


def _p(a: np.ndarray | None):
    return None if a is None else a.ctypes.data


# --------------------------------------------------------------------------
# errors (kinds mirror the reference's thiserror enums and ArrowError variants)
# --------------------------------------------------------------------------
class OracleError(Exception):
    def __init__(self, kind: str, msg: str):
        super().__init__(f"{kind}: {msg}")
        self.kind = kind
        self.msg = msg


# --------------------------------------------------------------------------
# arrays (Arrow layout: typed values, LSB-first packed validity / booleans)
# --------------------------------------------------------------------------
def nbytes_bits(n: int) -> int:
    return (n + 7) // 8


def pack_bits(b: np.ndarray) -> np.ndarray:
    out = np.packbits(np.asarray(b, dtype=bool), bitorder="little")
    return np.ascontiguousarray(out)


def unpack_bits(bits: np.ndarray, n: int, offset: int = 0) -> np.ndarray:
    return np.unpackbits(bits, bitorder="little")[offset:offset + n].astype(bool)


class Array:
    """One Arrow array: values buffer (+ offsets for utf8) + optional validity bitmap."""

    __slots__ = ("dtype", "length", "values", "offsets", "validity")

    def __init__(self, dtype, length, values, offsets=None, validity=None):
        self.dtype = dtype
        self.length = int(length)
        self.values = values
        self.offsets = offsets
        self.validity = validity

    @property
    def null_count(self) -> int:
        if self.validity is None:
            return 0
        return self.length - int(lib().ora_popcount(self.length, _p(self.validity)))

    def valid_mask(self) -> np.ndarray:
        if self.validity is None:
            return np.ones(self.length, dtype=bool)
        return unpack_bits(self.validity, self.length)

    def slice(self, start: int, n: int) -> "Array":
        if start == 0 and n == self.length:
            return self
        vm = None if self.validity is None else pack_bits(unpack_bits(self.validity, self.length)[start:start + n])
        if self.dtype == "bool":
            return Array("bool", n, pack_bits(unpack_bits(self.values, self.length)[start:start + n]), None, vm)
        if self.dtype == "utf8":
            off = self.offsets[start:start + n + 1]
            data = np.ascontiguousarray(self.values[off[0]:off[-1]])
            return Array("utf8", n, data, np.ascontiguousarray(off - off[0]).astype(np.int32), vm)
        return Array(self.dtype, n, np.ascontiguousarray(self.values[start:start + n]), None, vm)

    # -- python-level views, for tests ------------------------------------------------
    def to_pylist(self):
        vm = self.valid_mask()
        if self.dtype == "bool":
            v = unpack_bits(self.values, self.length)
            return [bool(v[i]) if vm[i] else None for i in range(self.length)]
        if self.dtype == "utf8":
            d = self.values.tobytes()
            return [d[self.offsets[i]:self.offsets[i + 1]].decode() if vm[i] else None for i in range(self.length)]
        return [self.values[i].item() if vm[i] else None for i in range(self.length)]

    @staticmethod
    def from_pylist(dtype: str, items) -> "Array":
        n = len(items)
        vm = np.array([x is not None for x in items], dtype=bool)
        validity = None if vm.all() else pack_bits(vm)
        if dtype == "bool":
            return Array("bool", n, pack_bits(np.array([bool(x) for x in items], dtype=bool)) if n else
                         np.zeros(0, np.uint8), None, validity)
        if dtype == "utf8":
            bs = [(x or "").encode() if not isinstance(x, bytes) else x for x in items]
            off = np.zeros(n + 1, dtype=np.int32)
            if n:
                off[1:] = np.cumsum([len(b) for b in bs])
            data = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if n else np.zeros(0, np.uint8)
            return Array("utf8", n, data, off, validity)
        vals = np.array([0 if x is None else x for x in items], dtype=NP_DTYPE[dtype])
        return Array(dtype, n, vals, None, validity)


class ArrayDatum:
    """compute_value.rs:34-55"""

    __slots__ = ("array", "is_scalar")

    def __init__(self, array: Array, is_scalar: bool):
        self.array = array
        self.is_scalar = is_scalar


class Field:
    __slots__ = ("name", "dtype", "nullable")

    def __init__(self, name, dtype, nullable):
        self.name, self.dtype, self.nullable = name, dtype, bool(nullable)

    def __eq__(self, o):
        return (self.name, self.dtype, self.nullable) == (o.name, o.dtype, o.nullable)

    def __repr__(self):
        return f"Field({self.name!r}, {self.dtype}, nullable={self.nullable})"


class Batch:
    """arrow RecordBatch: schema (fields) + equal-length columns."""

    def __init__(self, fields: list[Field], columns: list[Array], num_rows: int | None = None):
        if num_rows is None:
            if not columns:
                raise OracleError("InvalidArgumentError",
                                  "must either specify a row count or at least one column")
            num_rows = columns[0].length
        for c in columns:
            if c.length != num_rows:
                raise OracleError("InvalidArgumentError",
                                  "all columns in a record batch must have the same length")
        self.fields, self.columns, self.num_rows = fields, columns, num_rows

    def column_by_name(self, name: str):
        # arrow: first field with that name (test_arrow_compute_behavior.rs:81-108)
        for f, c in zip(self.fields, self.columns):
            if f.name == name:
                return c
        return None

    def slice(self, start, n) -> "Batch":
        return Batch(self.fields, [c.slice(start, n) for c in self.columns], n)


# --------------------------------------------------------------------------
# pyarrow bridge (test convenience)
# --------------------------------------------------------------------------
def _pa_type_name(t) -> str:
    import pyarrow as pa
    m = {pa.bool_(): "bool", pa.int8(): "int8", pa.int16(): "int16", pa.int32(): "int32", pa.int64(): "int64",
         pa.uint8(): "uint8", pa.uint16(): "uint16", pa.uint32(): "uint32", pa.uint64(): "uint64",
         pa.float32(): "float32", pa.float64(): "float64", pa.utf8(): "utf8"}
    if t not in m:
        raise OracleError("NotImplemented", f"arrow type {t}")
    return m[t]


def array_from_arrow(arr) -> Array:
    dtype = _pa_type_name(arr.type)
    n, off = len(arr), arr.offset
    bufs = arr.buffers()
    validity = None
    if bufs[0] is not None and arr.null_count > 0:
        validity = pack_bits(unpack_bits(np.frombuffer(bufs[0], dtype=np.uint8), n, off))
    if dtype == "bool":
        vals = pack_bits(unpack_bits(np.frombuffer(bufs[1], dtype=np.uint8), n, off)) if n else np.zeros(0, np.uint8)
        return Array("bool", n, vals, None, validity)
    if dtype == "utf8":
        o = np.frombuffer(bufs[1], dtype=np.int32)[off:off + n + 1] if bufs[1] is not None and bufs[1].size else \
            np.zeros(1, np.int32)
        data = np.frombuffer(bufs[2], dtype=np.uint8) if bufs[2] is not None and bufs[2].size else np.zeros(0, np.uint8)
        data = np.ascontiguousarray(data[o[0]:o[-1]])
        return Array("utf8", n, data, np.ascontiguousarray(o - o[0]).astype(np.int32), validity)
    vals = np.frombuffer(bufs[1], dtype=NP_DTYPE[dtype])[off:off + n].copy() if n else np.zeros(0, NP_DTYPE[dtype])
    return Array(dtype, n, vals, None, validity)


def batch_from_arrow(rb) -> Batch:
    fields = [Field(f.name, _pa_type_name(f.type), f.nullable) for f in rb.schema]
    return Batch(fields, [array_from_arrow(c) for c in rb.columns], rb.num_rows)


def array_to_arrow(a: Array):
    import pyarrow as pa
    t = {"bool": pa.bool_(), "utf8": pa.utf8()}.get(a.dtype) or getattr(pa, a.dtype)()
    vb = None if a.validity is None else pa.py_buffer(a.validity.tobytes())
    if a.dtype == "utf8":
        return pa.Array.from_buffers(t, a.length, [vb, pa.py_buffer(a.offsets.tobytes()),
                                                   pa.py_buffer(a.values.tobytes())])
    return pa.Array.from_buffers(t, a.length, [vb, pa.py_buffer(a.values.tobytes())])


def batch_to_arrow(b: Batch):
    import pyarrow as pa
    arrs = [array_to_arrow(c) for c in b.columns]
    schema = pa.schema([pa.field(f.name, a.type, f.nullable) for f, a in zip(b.fields, arrs)])
    return pa.RecordBatch.from_arrays(arrs, schema=schema)


# --------------------------------------------------------------------------
# literal typing (compute_value.rs:219-265)
# --------------------------------------------------------------------------
_RUST_FLOAT = re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)([eE][+-]?\d+)?$")
_RUST_INT = re.compile(r"^[+-]?\d+$")


def parse_f32(s: str) -> np.float32:
    """Rust `str::parse::<f32>()`: correctly rounded decimal -> binary32 (no double rounding)."""
    fr = Fraction(s)
    with np.errstate(over="ignore"):
        return _round_f32(fr, s)


def _round_f32(fr: Fraction, s: str) -> np.float32:
    f = np.float32(float(fr))
    if not np.isfinite(f) or fr == 0:
        if fr == 0 and s.lstrip().startswith("-"):
            return np.float32(-0.0)
        return f
    lo = np.nextafter(f, np.float32(-np.inf))
    hi = np.nextafter(f, np.float32(np.inf))
    best, best_d = None, None
    for c in (lo, f, hi):
        if not np.isfinite(c):
            continue
        d = abs(Fraction(float(c)) - fr)
        even = (int(np.float32(c).view(np.uint32)) & 1) == 0
        if best is None or d < best_d or (d == best_d and even):
            best, best_d = c, d
    fmax = Fraction(float(np.finfo(np.float32).max))
    if abs(fr) >= fmax + Fraction(2) ** 103:  # beyond max + half ulp -> inf
        return np.float32(np.inf if fr > 0 else -np.inf)
    return np.float32(best)


def _literal(val) -> ArrayDatum:
    if isinstance(val, dict) and "Number" in val:
        num_val, is_long = val["Number"]
        if is_long:
            raise OracleError("ValueTypeNotImplemented", repr(val))
        if "." in num_val:
            if _RUST_FLOAT.match(num_val):
                f = parse_f32(num_val)  # f32 parse never fails on a well-formed decimal => Float32
                return ArrayDatum(Array("float32", 1, np.array([f], dtype=np.float32)), True)
            raise OracleError("FailedToParseAsAFloat", num_val)
        if _RUST_INT.match(num_val):
            i = int(num_val)
            if -2**31 <= i < 2**31:
                return ArrayDatum(Array("int32", 1, np.array([i], dtype=np.int32)), True)
            if -2**63 <= i < 2**63:
                return ArrayDatum(Array("int64", 1, np.array([i], dtype=np.int64)), True)
        raise OracleError("FailedToParseAsAnInteger", num_val)
    if isinstance(val, dict) and "Boolean" in val:
        return ArrayDatum(Array.from_pylist("bool", [bool(val["Boolean"])]), True)
    if isinstance(val, dict) and "SingleQuotedString" in val:
        return ArrayDatum(Array.from_pylist("utf8", [val["SingleQuotedString"]]), True)
    raise OracleError("ValueTypeNotImplemented", repr(val))


# --------------------------------------------------------------------------
# coercion (compute_value.rs:350-461)
# --------------------------------------------------------------------------
_SIGNED = ["int8", "int16", "int32", "int64"]
_UNSIGNED = ["uint8", "uint16", "uint32", "uint64"]


def get_common_type(left: str, right: str) -> str:
    if left == right:
        return left
    pair = {left, right}
    if left in _SIGNED and right in _SIGNED:
        return _SIGNED[max(_SIGNED.index(left), _SIGNED.index(right))]
    if left in _UNSIGNED and right in _UNSIGNED:
        return _UNSIGNED[max(_UNSIGNED.index(left), _UNSIGNED.index(right))]
    # mixed signed/unsigned: only (uN, i2N or wider) is listed (:375-383)
    for u, s in ((left, right), (right, left)):
        if u in _UNSIGNED and s in _SIGNED:
            if _SIGNED.index(s) > _UNSIGNED.index(u) and u != "uint64":
                return s
            raise OracleError("UnsupportedTypeCoersion", f"{ARROW_NAME[left]} and {ARROW_NAME[right]}")
    if pair == {"float32", "float64"}:
        return "float64"
    for f, i in ((left, right), (right, left)):
        if f == "float32" and i in ("int8", "int16", "int32", "uint8", "uint16", "uint32"):
            return "float32"
        if f == "float64" and (i in _SIGNED or i in _UNSIGNED):
            return "float64"
    raise OracleError("UnsupportedTypeCoersion", f"{ARROW_NAME.get(left, left)} and {ARROW_NAME.get(right, right)}")


def _cast(a: Array, to: str) -> Array:
    """arrow compute::cast on the coercion lattice (validity preserved)."""
    out = np.empty(a.length, dtype=NP_DTYPE[to])
    rc = lib().ora_cast(TYPE_ID[a.dtype], TYPE_ID[to], a.length, _p(a.values), _p(out))
    if rc != 0:
        raise OracleError("CastError", f"{a.dtype} -> {to}")
    return Array(to, a.length, out, None, a.validity)


def cast_to_common_type(left: ArrayDatum, right: ArrayDatum):
    common = get_common_type(left.array.dtype, right.array.dtype)
    la = left.array if left.array.dtype == common else _cast(left.array, common)
    ra = right.array if right.array.dtype == common else _cast(right.array, common)
    return ArrayDatum(la, left.is_scalar), ArrayDatum(ra, right.is_scalar)


def _cast_to_bool(a: Array) -> Array:
    """compute::cast(x, Boolean) (compute_value.rs:72-73, :95-96)."""
    if a.dtype == "bool":
        return a
    if a.dtype == "utf8":
        # arrow can parse 'true'/'false' strings here; not a path any reference query reaches.
        raise OracleError("NotImplemented", "cast Utf8 to Boolean")
    out = np.zeros(nbytes_bits(a.length), dtype=np.uint8)
    lib().ora_cast_to_bool(TYPE_ID[a.dtype], a.length, _p(a.values), _p(out))
    return Array("bool", a.length, out, None, a.validity)


def _union_validity(n: int, a: Array, a_scalar: bool, b: Array, b_scalar: bool):
    """NullBuffer::union; literals are never null so a scalar side contributes nothing."""
    va = None if a_scalar else a.validity
    vb = None if b_scalar else b.validity
    if a_scalar and a.validity is not None or b_scalar and b.validity is not None:
        raise OracleError("NotImplemented", "null scalar")
    if va is None:
        return vb
    if vb is None:
        return va
    out = np.empty(nbytes_bits(n), dtype=np.uint8)
    lib().ora_bitmap_and(n, _p(va), _p(vb), _p(out))
    return out


def _result_len(l: ArrayDatum, r: ArrayDatum, what: str, kind: str) -> int:
    if l.is_scalar and r.is_scalar:
        return 1
    if l.is_scalar:
        return r.array.length
    if r.is_scalar:
        return l.array.length
    if l.array.length != r.array.length:
        raise OracleError(kind, f"{what} arrays of different length, got {l.array.length} vs {r.array.length}")
    return l.array.length


# --------------------------------------------------------------------------
# compute_value (compute_value.rs:57-344)
# --------------------------------------------------------------------------
def _boolean_kernel(l: ArrayDatum, r: ArrayDatum, is_and: bool) -> ArrayDatum:
    lb, rb = _cast_to_bool(l.array), _cast_to_bool(r.array)
    if lb.length != rb.length:  # arrow-arith boolean.rs binary_boolean_kernel
        raise OracleError("ComputeError", "Cannot perform bitwise operation on arrays of different length")
    n = lb.length
    out = np.empty(nbytes_bits(n), dtype=np.uint8)
    (lib().ora_bitmap_and if is_and else lib().ora_bitmap_or)(n, _p(lb.values), _p(rb.values), _p(out))
    validity = _union_validity(n, lb, False, rb, False)
    # new_binary_op(&BooleanArray, &BooleanArray, ..): a bare array's Datum flag is false
    return ArrayDatum(Array("bool", n, out, None, validity), False)


def _arith(op: str, l: ArrayDatum, r: ArrayDatum) -> ArrayDatum:
    l, r = cast_to_common_type(l, r)
    t = l.array.dtype
    if t in ("bool", "utf8"):
        raise OracleError("InvalidArgumentError", f"Invalid arithmetic operation: {ARROW_NAME[t]} {op} {ARROW_NAME[t]}")
    n = _result_len(l, r, "Cannot perform binary operation on", "ComputeError")
    validity = _union_validity(n, l.array, l.is_scalar, r.array, r.is_scalar)
    out = np.empty(n, dtype=NP_DTYPE[t])
    err_row = ctypes.c_int64(-1)
    rc = lib().ora_arith(ARITH.get(op, 4), TYPE_ID[t], n, _p(l.array.values), int(l.is_scalar), _p(r.array.values),
                         int(r.is_scalar), _p(validity), _p(out), ctypes.byref(err_row))
    if rc == 1:
        raise OracleError("ArithmeticOverflow", f"Overflow happened on row {err_row.value}")
    if rc == 2:
        raise OracleError("DivideByZero", f"Divide by zero error (row {err_row.value})")
    if rc != 0:
        raise OracleError("ComputeError", "bad arithmetic arguments")
    return ArrayDatum(Array(t, n, out, None, validity), l.is_scalar and r.is_scalar)


def _compare(op: str, l: ArrayDatum, r: ArrayDatum) -> ArrayDatum:
    l, r = cast_to_common_type(l, r)
    t = l.array.dtype
    n = _result_len(l, r, "Cannot compare", "InvalidArgumentError")
    validity = _union_validity(n, l.array, l.is_scalar, r.array, r.is_scalar)
    out = np.empty(nbytes_bits(n), dtype=np.uint8)
    la, ra, L = l.array, r.array, lib()
    if t == "bool":
        L.ora_cmp_bool(CMP[op], n, _p(la.values), int(l.is_scalar), _p(ra.values), int(r.is_scalar), _p(out))
    elif t == "utf8":
        L.ora_cmp_utf8(CMP[op], n, _p(la.offsets), _p(la.values), int(l.is_scalar), _p(ra.offsets), _p(ra.values),
                       int(r.is_scalar), _p(out))
    else:
        L.ora_cmp(CMP[op], TYPE_ID[t], n, _p(la.values), int(l.is_scalar), _p(ra.values), int(r.is_scalar), _p(out))
    return ArrayDatum(Array("bool", n, out, None, validity), l.is_scalar and r.is_scalar)


def _boolean_kernel_kleene(l: ArrayDatum, r: ArrayDatum, is_and: bool) -> ArrayDatum:
    """arrow-arith boolean.rs and_kleene / or_kleene: a false (true) side decides an AND (OR) whatever the other side is."""
    lb, rb = _cast_to_bool(l.array), _cast_to_bool(r.array)
    if lb.length != rb.length:
        raise OracleError("ComputeError", "Cannot perform bitwise operation on arrays of different length")
    n = lb.length
    av, bv = lb.valid_mask(), rb.valid_mask()
    a, b = unpack_bits(lb.values, n), unpack_bits(rb.values, n)
    at, bt, af, bf = a & av, b & bv, ~a & av, ~b & bv
    if is_and:
        val, valid = at & bt, (av & bv) | af | bf
    else:
        val, valid = at | bt, (av & bv) | at | bt
    validity = None if valid.all() else np.packbits(valid, bitorder="little")
    return ArrayDatum(Array("bool", n, np.packbits(val, bitorder="little") if n else np.zeros(0, np.uint8), None, validity), False)


def _unary(op: str, x: ArrayDatum) -> ArrayDatum:
    a = x.array
    if op == "Plus":
        return x
    if op == "Not":  # boolean::not: values flipped, validity kept
        b = _cast_to_bool(a)
        vals = np.packbits(~unpack_bits(b.values, b.length), bitorder="little") if b.length else np.zeros(0, np.uint8)
        return ArrayDatum(Array("bool", b.length, vals, None, b.validity), x.is_scalar)
    if op != "Minus":
        raise OracleError("ExpressionTypeNotImplemented", "UnaryOp " + op)
    if a.dtype in ("int8", "int16", "int32", "int64"):  # numeric::neg = neg_checked on valid slots
        zero = ArrayDatum(Array(a.dtype, 1, np.zeros(1, dtype=NP_DTYPE[a.dtype])), True)
        return _arith("Minus", zero, x)
    if a.dtype in ("float32", "float64"):  # neg_wrapping: the sign bit, NaN included
        u = a.values.view(np.uint32 if a.dtype == "float32" else np.uint64)
        out = (u ^ (u.dtype.type(1) << u.dtype.type(u.dtype.itemsize * 8 - 1))).view(a.values.dtype)
        return ArrayDatum(Array(a.dtype, a.length, out, None, a.validity), x.is_scalar)
    raise OracleError("InvalidArgumentError", f"Invalid arithmetic operation: -{ARROW_NAME[a.dtype]}")


def _is_null(x: ArrayDatum, negate: bool) -> ArrayDatum:
    a = x.array
    v = a.valid_mask()
    bits = v if negate else ~v
    return ArrayDatum(Array("bool", a.length, np.packbits(bits, bitorder="little") if a.length else np.zeros(0, np.uint8)), x.is_scalar)


def compute_value(rec: Batch, table_aliases: list[list[str]], expr) -> ArrayDatum:
    if isinstance(expr, dict) and len(expr) == 1:
        (tag, body), = expr.items()
    else:
        tag, body = (expr if isinstance(expr, str) else "?"), None

    if tag == "Nested":  # :63-65
        return compute_value(rec, table_aliases, body)
    if tag == "BinaryOp":  # :66-218
        left = compute_value(rec, table_aliases, body["left"])
        right = compute_value(rec, table_aliases, body["right"])
        op = body["op"] if isinstance(body["op"], str) else next(iter(body["op"]))
        if op in ("And", "Or"):
            return (_boolean_kernel_kleene if EXTENSIONS & EXT_KLEENE else _boolean_kernel)(left, right, op == "And")
        if op in ARITH or (op == "Minus" and EXTENSIONS & EXT_OPERATORS):
            return _arith(op, left, right)
        if op in CMP:
            return _compare(op, left, right)
        raise OracleError("BinaryOperatorNotImplemented", op)  # :210-216 (Minus lands here)
    if tag == "Value":  # :219-265
        return _literal(body)
    if tag == "Identifier":  # :266-274
        col = rec.column_by_name(body["value"])
        if col is None:
            raise OracleError("ColumnNotFound", body["value"])
        return ArrayDatum(col, False)
    if tag == "CompoundIdentifier":  # :275-337
        idents = body
        if len(idents) == 1:
            col = rec.column_by_name(idents[0]["value"])
            if col is None:
                raise OracleError("ColumnNotFound", idents[0]["value"])
            return ArrayDatum(col, False)
        if len(idents) == 2:
            alias, name = idents[0]["value"], idents[1]["value"]
            for idx, f in enumerate(rec.fields):
                if f.name == name:
                    if idx >= len(table_aliases):
                        raise OracleError("Panic", "table aliases vec has incorrect length")
                    if alias in table_aliases[idx]:
                        return ArrayDatum(rec.columns[idx], False)
        raise OracleError("IdentifierNotFound", ".".join(i["value"] for i in idents))
    if EXTENSIONS & EXT_OPERATORS and tag == "UnaryOp":
        return _unary(body["op"], compute_value(rec, table_aliases, body["expr"]))
    if EXTENSIONS & EXT_OPERATORS and tag in ("IsNull", "IsNotNull"):
        return _is_null(compute_value(rec, table_aliases, body), tag == "IsNotNull")
    raise OracleError("ExpressionTypeNotImplemented", tag)  # :338-342


# --------------------------------------------------------------------------
# filter_record (filter_record.rs:21-39) + arrow-select filter_record_batch
# --------------------------------------------------------------------------
def _selection(mask: Array) -> tuple[np.ndarray, int]:
    """predicate with nulls -> values & validity (NULL = drop); count = popcount."""
    n = mask.length
    if mask.validity is None:
        sel = mask.values
    else:
        sel = np.empty(nbytes_bits(n), dtype=np.uint8)
        lib().ora_bitmap_and(n, _p(mask.values), _p(mask.validity), _p(sel))
    return sel, int(lib().ora_popcount(n, _p(sel)))


def _filter_array(a: Array, n: int, sel: np.ndarray, count: int) -> Array:
    L = lib()
    validity = None
    if a.validity is not None and a.null_count > 0:
        vb = np.empty(nbytes_bits(n), dtype=np.uint8)
        nset = L.ora_filter_bits(n, _p(sel), _p(a.validity), _p(vb))
        if nset != count:  # dropped entirely when no nulls survive
            validity = np.ascontiguousarray(vb[:nbytes_bits(count)])
    if a.dtype == "bool":
        out = np.empty(nbytes_bits(n), dtype=np.uint8)
        L.ora_filter_bits(n, _p(sel), _p(a.values), _p(out))
        return Array("bool", count, np.ascontiguousarray(out[:nbytes_bits(count)]), None, validity)
    if a.dtype == "utf8":
        off = np.empty(count + 1, dtype=np.int32)
        data = np.empty(max(int(a.offsets[n]) if n else 0, 1), dtype=np.uint8)
        nb = L.ora_filter_utf8(n, _p(sel), _p(a.offsets), _p(a.values), _p(off), _p(data))
        return Array("utf8", count, np.ascontiguousarray(data[:nb]), off, validity)
    out = np.empty(count, dtype=a.values.dtype)
    L.ora_filter_fixed(a.values.dtype.itemsize, n, _p(sel), _p(a.values), _p(out))
    return Array(a.dtype, count, out, None, validity)


def filter_record_batch(rec: Batch, mask: Array) -> Batch:
    n = mask.length
    for c in rec.columns:
        if n > c.length:
            raise OracleError("InvalidArgumentError",
                              f"Filter predicate of length {n} is larger than target array of length {c.length}")
    sel, count = _selection(mask)
    if count == 0:  # IterationStrategy::None
        return rec.slice(0, 0)
    if count == n:  # IterationStrategy::All -> values.slice(0, count)
        return rec.slice(0, count)
    return Batch(rec.fields, [_filter_array(c, n, sel, count) for c in rec.columns], count)


def filter_record(rec: Batch, table_aliases, expr) -> Batch:
    res = compute_value(rec, table_aliases, expr)
    if res.array.dtype != "bool":
        raise OracleError("CastToBooleanArrayFailedForArrayType", ARROW_NAME[res.array.dtype])
    return filter_record_batch(rec, res.array)


# --------------------------------------------------------------------------
# project_record (record_projection.rs:16-76)
# --------------------------------------------------------------------------
def project_record(fields, record: Batch, table_aliases) -> Batch:
    unnamed_idx = 0
    proj_fields, proj_arrays = [], []
    for item in fields:
        tag, body = (item, None) if isinstance(item, str) else next(iter(item.items()))
        if tag == "Wildcard":
            for f, c in zip(record.fields, record.columns):
                proj_fields.append(f)
                proj_arrays.append(c)
        elif tag == "QualifiedWildcard":
            raise OracleError("NotImplemented", "SelectItem::QualifiedWildcard")
        elif tag == "UnnamedExpr":
            res = compute_value(record, table_aliases, body)
            if isinstance(body, dict) and "Identifier" in body:
                name = body["Identifier"]["value"]
            else:
                name = f"unnamed_{unnamed_idx}"
            proj_fields.append(Field(name, res.array.dtype, res.array.null_count != 0))
            proj_arrays.append(res.array)
            unnamed_idx += 1
        elif tag == "ExprWithAlias":
            res = compute_value(record, table_aliases, body["expr"])
            proj_fields.append(Field(body["alias"]["value"], res.array.dtype, res.array.null_count != 0))
            proj_arrays.append(res.array)
        else:
            raise OracleError("NotImplemented", f"SelectItem::{tag}")
    return Batch(proj_fields, proj_arrays)


# --------------------------------------------------------------------------
# get_record_table_aliases (record_aliases.rs:12-59)
# --------------------------------------------------------------------------
def get_record_table_aliases(op_type: dict, record: Batch) -> list[list[str]]:
    """op_type: serde form of planner::OperatorType, e.g.
    {"Producer": {"task": {"TableFunc": {"alias": "t", ...}}, ...}}"""
    (_, body), = op_type.items()
    (task_name, task), = body["task"].items()
    if task_name not in ("TableFunc", "Table"):
        raise OracleError("OperatorTaskTypeDoesNotHaveAnAliasField", f"OperatorTask::{task_name}")
    alias = task.get("alias")
    return [[alias] if alias is not None else [] for _ in range(len(record.columns))]


# --------------------------------------------------------------------------
# comparison rule (SURVEY.md section 8c "bit-exact")
# --------------------------------------------------------------------------
def arrays_equal(a: Array, b: Array) -> tuple[bool, str]:
    if a.dtype != b.dtype:
        return False, f"dtype {a.dtype} != {b.dtype}"
    if a.length != b.length:
        return False, f"length {a.length} != {b.length}"
    va, vb = a.valid_mask(), b.valid_mask()
    if not np.array_equal(va, vb):
        i = int(np.nonzero(va != vb)[0][0])
        return False, f"validity differs at row {i}"
    if a.length == 0:
        return True, ""
    if a.dtype == "bool":
        x, y = unpack_bits(a.values, a.length), unpack_bits(b.values, b.length)
    elif a.dtype == "utf8":
        if not np.array_equal(a.offsets.astype(np.int64) - int(a.offsets[0]),
                              b.offsets.astype(np.int64) - int(b.offsets[0])):
            # lengths under null slots are part of the offsets buffer, compare only valid rows
            la, lb = np.diff(a.offsets), np.diff(b.offsets)
            bad = np.nonzero((la != lb) & va)[0]
            if bad.size:
                return False, f"utf8 length differs at row {int(bad[0])}"
        for i in np.nonzero(va)[0]:
            sa = a.values[a.offsets[i]:a.offsets[i + 1]]
            sb = b.values[b.offsets[i]:b.offsets[i + 1]]
            if not np.array_equal(sa, sb):
                return False, f"utf8 bytes differ at row {int(i)}"
        return True, ""
    else:
        w = a.values.dtype.itemsize
        ut = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[w]
        x, y = a.values.view(ut), b.values.view(ut)  # floats compare as bits
    neq = (x != y) & va
    if neq.any():
        i = int(np.nonzero(neq)[0][0])
        return False, f"value differs at row {i}: {x[i]} vs {y[i]}"
    return True, ""


def batches_equal(a: Batch, b: Batch, check_nullable: bool = True) -> tuple[bool, str]:
    if a.num_rows != b.num_rows:
        return False, f"num_rows {a.num_rows} != {b.num_rows}"
    if len(a.columns) != len(b.columns):
        return False, f"num_columns {len(a.columns)} != {len(b.columns)}"
    for i, (fa, fb) in enumerate(zip(a.fields, b.fields)):
        if fa.name != fb.name or fa.dtype != fb.dtype or (check_nullable and fa.nullable != fb.nullable):
            return False, f"field {i}: {fa} != {fb}"
    for i, (ca, cb) in enumerate(zip(a.columns, b.columns)):
        ok, why = arrays_equal(ca, cb)
        if not ok:
            return False, f"column {i} ({a.fields[i].name}): {why}"
    return True, ""
