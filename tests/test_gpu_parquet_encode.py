"""Device batches -> Parquet (csrc/parquet_encode.inc, chdb_parquet_encode), SURVEY.md 8f row f3: the GPU build's materialize
(materialize_files_task.rs:116-141) with the record compaction of DEV_NOTES.md:117-122.  The written image is read back with
pyarrow's reader (Arrow C++, an independent implementation of the format) and with this library's own device decoder, and
compared with the batches that went in."""
import io

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

import chapterhouseqe_b200 as C
import parquet_cases as PC
from chapterhouseqe_b200 import sqlparser_lite as sp
from oracle import compute_value as O

pytestmark = pytest.mark.gpu


def as_batch(t: pa.Table) -> pa.RecordBatch:
    t = t.combine_chunks()
    return pa.RecordBatch.from_arrays([c.chunk(0) if c.num_chunks else pa.array([], type=c.type) for c in t.columns], schema=t.schema)


def cut(t: pa.Table, sizes) -> list:
    """The table as consecutive record batches of the given row counts (the last takes the rest)."""
    out, at = [], 0
    for s in sizes:
        out.append(as_batch(t.slice(at, s)))
        at += s
    if at < t.num_rows:
        out.append(as_batch(t.slice(at)))
    return out


def assert_same(got: pa.Table, want: pa.Table, what=""):
    assert got.schema.names == want.schema.names
    assert got.num_rows == want.num_rows, what
    for name in want.schema.names:
        g, w = got.column(name).combine_chunks(), want.column(name).combine_chunks()
        assert g.type == w.type, f"{what} column {name}: {g.type} != {w.type}"
        assert g.null_count == w.null_count, f"{what} column {name}: null count {g.null_count} != {w.null_count}"
        assert g.equals(w), f"{what} column {name} differs"


def encode_and_check(t: pa.Table, sizes, max_rows=0, max_groups=0):
    parts = cut(t, sizes)
    devs = [C.DeviceBatch.upload(rb) for rb in parts]
    img = C.encode_parquet(devs, max_rows, max_groups)
    data = img.to_bytes()
    assert len(data) == img.nbytes and data[:4] == b"PAR1" and data[-4:] == b"PAR1"
    f = pq.ParquetFile(io.BytesIO(data))
    want = pa.Table.from_batches(parts[:img.consumed], schema=t.schema) if img.consumed else t.slice(0, 0)
    assert f.metadata.num_row_groups == img.row_groups
    for field in t.schema:   # nullable columns are OPTIONAL, the others REQUIRED
        assert f.schema_arrow.field(field.name).nullable == field.nullable, field.name
    assert_same(f.read(), want, "pyarrow")
    # and through this library's own decoder, row group by row group
    mine = C.ParquetFile(data)
    assert mine.num_rows == want.num_rows and mine.num_row_groups == img.row_groups
    if img.row_groups:
        back = pa.Table.from_batches([d.download() for d in mine.decode_row_groups()])
        assert_same(back, want, "own decoder")
    return img, f


@pytest.mark.parametrize("n,sizes", [(1, [1]), (31, [7, 8, 9]), (5000, [2048, 1, 2047]), (70001, [10000] * 6)])
def test_every_type_round_trips(n, sizes):
    encode_and_check(PC.sample_table(n, seed=n), sizes)


def test_no_nulls_and_wide_strings():
    encode_and_check(PC.sample_table(20000, seed=5, nulls=False, wide=100), [7000, 7000])


def test_wide_strings_with_nulls():
    encode_and_check(PC.sample_table(9000, seed=6, wide=100).select(["id", "value1", "flag"]), [4001, 13])


def test_records_are_coalesced_into_row_groups():
    """25 reference-native records of 10 000 rows (physical_planner.rs:323) -> row groups of at most 65 536 rows: six records
    per row group, each record one page per column; max_row_groups cuts the file and reports what it consumed."""
    t = PC.sample_table(250_000, seed=9).select(["id", "value1", "value2", "flag"])
    img, f = encode_and_check(t, [10_000] * 25, max_rows=65_536)
    assert img.consumed == 25 and img.row_groups == 5
    assert [f.metadata.row_group(i).num_rows for i in range(5)] == [60_000] * 4 + [10_000]
    img, f = encode_and_check(t, [10_000] * 25, max_rows=65_536, max_groups=2)
    assert img.consumed == 12 and img.row_groups == 2 and f.metadata.num_rows == 120_000
    # a record larger than the limit gets a row group of its own
    img, f = encode_and_check(t, [100_000, 10_000, 10_000], max_rows=65_536)
    assert [f.metadata.row_group(i).num_rows for i in range(f.metadata.num_row_groups)] == [100_000, 20_000, 130_000]


def test_empty_records_contribute_nothing():
    t = PC.sample_table(300, seed=3)
    img, f = encode_and_check(t, [0, 100, 0, 0, 200])
    assert img.consumed == 5 and img.row_groups == 1
    img, f = encode_and_check(t.slice(0, 0), [0])
    assert img.row_groups == 0 and f.metadata.num_rows == 0


def test_filter_then_project_then_encode_without_leaving_the_device():
    """filter -> projection -> materialize on device batches: only the Parquet image crosses PCIe.  The image holds exactly
    what the oracle's filter_record + project_record produce (per-batch nullability of the projected columns folded into
    one file schema: OPTIONAL if any record says nullable)."""
    t = PC.sample_table(120_000, seed=12).select(["id", "value1", "value2", "d", "k"])
    parts = cut(t, [30_000] * 4)
    sel = sp.parse_select("select id, value1, value2 + 10 as v, d from t where (id % 2 = 0 and value2 > 10.0) or d < 0.5")
    expr, items = sel["selection"], sel["projection"]
    prog = C.Program.compile_filter_project(expr, items, t.schema)
    al = [[] for _ in t.schema]
    outs = [C.DeviceBatch.upload(rb).run(prog) for rb in parts]
    wants = [O.project_record(items, O.filter_record(O.batch_from_arrow(rb), al, expr), al) for rb in parts]
    img = C.encode_parquet(outs, max_rows_per_row_group=50_000)
    f = pq.ParquetFile(io.BytesIO(img.to_bytes()))
    got = f.read()
    assert img.consumed == 4 and sum(w.num_rows for w in wants) == got.num_rows
    at = 0
    for i, w in enumerate(wants):
        piece = as_batch(got.slice(at, w.num_rows))
        ok, why = O.batches_equal(O.batch_from_arrow(piece), w, check_nullable=False)
        assert ok, f"record {i}: {why}"
        at += w.num_rows


def test_mismatched_schemas_are_rejected():
    a = C.DeviceBatch.upload(as_batch(PC.sample_table(10, seed=1).select(["id", "value1"])))
    b = C.DeviceBatch.upload(as_batch(PC.sample_table(10, seed=1).select(["id", "value2"])))
    c = C.DeviceBatch.upload(as_batch(PC.sample_table(10, seed=1).select(["id"])))
    with pytest.raises(C.ChdbError):
        C.encode_parquet([a, b])
    with pytest.raises(C.ChdbError):
        C.encode_parquet([a, c])
    with pytest.raises(C.ChdbError):
        C.encode_parquet([])
