"""Shared test plumbing: build Arrow batches from KAT specs, run them through an
implementation (the CPU oracle or the CUDA library through its C-ABI) and compare with
the bit-exact rule of SURVEY.md section 8c."""
from __future__ import annotations

import struct

import numpy as np
import pyarrow as pa

from chapterhouseqe_b200 import sqlparser_lite as sp
from oracle import compute_value as O

PA_TYPE = {"bool": pa.bool_(), "int8": pa.int8(), "int16": pa.int16(), "int32": pa.int32(), "int64": pa.int64(),
           "uint8": pa.uint8(), "uint16": pa.uint16(), "uint32": pa.uint32(), "uint64": pa.uint64(),
           "float32": pa.float32(), "float64": pa.float64(), "utf8": pa.utf8()}


def _bits_to_float(dtype: str, bits: int) -> float:
    if dtype == "float32":
        return struct.unpack("<f", struct.pack("<I", bits))[0]
    return struct.unpack("<d", struct.pack("<Q", bits))[0]


def make_array(dtype: str, items) -> pa.Array:
    """python list (None = null, ("bits", u) = raw float bits) -> pyarrow array, bit-preserving."""
    if dtype in ("float32", "float64"):
        npt = np.float32 if dtype == "float32" else np.float64
        ut = np.uint32 if dtype == "float32" else np.uint64
        raw = np.zeros(len(items), dtype=ut)
        mask = np.zeros(len(items), dtype=bool)
        for i, x in enumerate(items):
            if x is None:
                mask[i] = True
            elif isinstance(x, tuple):
                raw[i] = x[1]
            else:
                raw[i] = np.array([x], dtype=npt).view(ut)[0]
        return pa.array(raw.view(npt), type=PA_TYPE[dtype], mask=mask if mask.any() else None, from_pandas=False)
    return pa.array(items, type=PA_TYPE[dtype])


def make_batch(schema, cols) -> pa.RecordBatch:
    arrays = [make_array(t, c) for (_, t, _), c in zip(schema, cols)]
    fields = [pa.field(n, PA_TYPE[t], nullable) for (n, t, nullable) in schema]
    return pa.RecordBatch.from_arrays(arrays, schema=pa.schema(fields))


def expected_array(dtype: str, items) -> O.Array:
    return O.array_from_arrow(make_array(dtype, items))


class OracleImpl:
    """The CPU oracle behind the same three entry points the reference exports
    (record_utils/mod.rs:13-15 + compute_value)."""

    name = "oracle"

    def compute_value(self, rb: pa.RecordBatch, aliases, expr: dict) -> O.Array:
        return O.compute_value(O.batch_from_arrow(rb), aliases, expr).array

    def filter_record(self, rb, aliases, expr) -> O.Batch:
        return O.filter_record(O.batch_from_arrow(rb), aliases, expr)

    def project_record(self, fields, rb, aliases) -> O.Batch:
        return O.project_record(fields, O.batch_from_arrow(rb), aliases)

    def filter_project_record(self, expr, fields, rb, aliases) -> O.Batch:
        b = O.batch_from_arrow(rb)
        return O.project_record(fields, O.filter_record(b, aliases, expr), aliases)


def error_kind(exc: Exception) -> str:
    return getattr(exc, "kind", type(exc).__name__)


def run_case(impl, case):
    """Returns ("ok", result) or ("error", kind). result: O.Array for kind=value else O.Batch."""
    schema, cols = case["schema"], case["cols"]
    rb = make_batch(schema, cols)
    aliases = case.get("aliases")
    if aliases is None:
        aliases = [[] for _ in schema]
    kind = case["kind"]
    try:
        if kind == "value":
            return "ok", impl.compute_value(rb, aliases, sp.parse_expr(case["sql"]))
        if kind == "filter":
            return "ok", impl.filter_record(rb, aliases, sp.parse_expr(case["sql"]))
        sel = sp.parse_select(case["sql"])
        if kind == "project":
            return "ok", impl.project_record(sel["projection"], rb, aliases)
        if kind == "filter_project":
            return "ok", impl.filter_project_record(sel["selection"], sel["projection"], rb, aliases)
        raise ValueError(kind)
    except Exception as e:  # noqa: BLE001 - error kinds are part of the contract
        if not hasattr(e, "kind"):
            raise
        return "error", error_kind(e)


def check_case(impl, case):
    status, got = run_case(impl, case)
    want_status, want = case["expect"]
    assert status == want_status, f"{case['name']}: got {status} {got!r}, want {want_status} {want!r}"
    if status == "error":
        assert got == want, f"{case['name']}: error kind {got} != {want}"
        return
    kind = case["kind"]
    if kind == "value":
        dtype, items = want
        ok, why = O.arrays_equal(got, expected_array(dtype, items))
        assert ok, f"{case['name']}: {why}; got {got.to_pylist()}"
    elif kind == "filter":
        assert len(got.columns) == len(want)
        for (n, t, nullable), f, col, items in zip(case["schema"], got.fields, got.columns, want):
            assert (f.name, f.dtype, f.nullable) == (n, t, nullable), f"{case['name']}: schema changed: {f}"
            ok, why = O.arrays_equal(col, expected_array(t, items))
            assert ok, f"{case['name']}: column {n}: {why}; got {col.to_pylist()}"
        assert got.num_rows == (len(want[0]) if want else 0)
    else:
        assert len(got.columns) == len(want), f"{case['name']}: {len(got.columns)} columns"
        for f, col, (n, t, nullable, items) in zip(got.fields, got.columns, want):
            assert (f.name, f.dtype, f.nullable) == (n, t, nullable), f"{case['name']}: field {f} != {(n, t, nullable)}"
            ok, why = O.arrays_equal(col, expected_array(t, items))
            assert ok, f"{case['name']}: column {n}: {why}; got {col.to_pylist()}"
