"""Parquet -> device decode (csrc/parquet.inc, chdb_parquet_*) against pyarrow's reader (Arrow C++, an independent
implementation of the format) on files written in the reference's format and its supported neighbours; then the
decoded device batch straight into the filter, against the oracle."""
import numpy as np
import pyarrow as pa
import pytest

import chapterhouseqe_b200 as C
import parquet_cases as PC
from chapterhouseqe_b200 import sqlparser_lite as sp
from oracle import compute_value as O

pytestmark = pytest.mark.gpu


def as_batch(t: pa.Table) -> pa.RecordBatch:
    t = t.combine_chunks()
    return pa.RecordBatch.from_arrays([c.chunk(0) if c.num_chunks else pa.array([], type=c.type) for c in t.columns], schema=t.schema)


def check_file(data: bytes, table: pa.Table):
    f = C.ParquetFile(data)
    assert f.schema.equals(table.schema)
    rows = 0
    for i in range(f.num_row_groups):
        want = as_batch(PC.read_row_group(data, i))
        dev = f.decode_row_group(i)
        assert dev.num_rows == want.num_rows
        got = dev.download()
        assert got.schema.names == want.schema.names
        for name in want.schema.names:
            g, w = got.column(name), want.column(name)
            assert g.type == w.type, name
            assert g.null_count == w.null_count, f"row group {i} column {name}: null count {g.null_count} != {w.null_count}"
            assert g.equals(w), f"row group {i} column {name} differs"
        rows += want.num_rows
    assert rows == table.num_rows


@pytest.mark.parametrize("variant", sorted(PC.WRITER_VARIANTS))
@pytest.mark.parametrize("n,rg", [(1, 100), (31, 100), (5000, 2048), (70001, 30000)])
def test_decode_matches_pyarrow(variant, n, rg):
    t = PC.sample_table(n, seed=n)
    check_file(PC.write(t, row_group_size=rg, **PC.WRITER_VARIANTS[variant]), t)


@pytest.mark.parametrize("variant", ["reference_like", "plain", "small_pages"])
def test_decode_without_nulls_and_wide_strings(variant):
    t = PC.sample_table(20000, seed=5, nulls=False, wide=100)
    check_file(PC.write(t, row_group_size=8000, **PC.WRITER_VARIANTS[variant]), t)


def test_reference_sample_schema_one_million_rows():
    """create_sample_data.rs: id Int32, value1 Utf8, value2 Float32, no nulls, one row group per file."""
    n = 1_000_000
    rng = np.random.default_rng(3)
    t = pa.table({"id": pa.array(np.arange(n, dtype=np.int32)),
                  "value1": pa.array(np.char.add("v", rng.integers(0, 10**7, n).astype(str))),
                  "value2": pa.array(rng.uniform(0, 100, n).astype(np.float32))})
    check_file(PC.write(t), t)
    check_file(PC.write(t, use_dictionary=False), t)


def test_decoded_batch_feeds_the_filter():
    """read_files -> filter without leaving the device: only the filtered result is downloaded."""
    t = PC.sample_table(50000, seed=11).select(["id", "value1", "value2", "d", "k"])
    data = PC.write(t, row_group_size=20000)
    f = C.ParquetFile(data)
    expr = sp.parse_expr("(id % 2 = 0 and value2 > 10.0) or d < 0.5")
    prog = C.Program.compile_filter(expr, f.schema)
    al = [[] for _ in f.schema]
    for i in range(f.num_row_groups):
        want = O.filter_record(O.batch_from_arrow(as_batch(PC.read_row_group(data, i))), al, expr)
        got = O.batch_from_arrow(f.decode_row_group(i).run(prog).download())
        ok, why = O.batches_equal(got, want)
        assert ok, f"row group {i}: {why}"


def test_truncated_page_does_not_crash():
    t = PC.sample_table(3000, seed=2).select(["id", "value1"])
    data = bytearray(PC.write(t, use_dictionary=False))
    # corrupt one string length prefix in the middle of the value1 chunk: the walk stops there, nothing reads out of bounds
    pos = data.find(b"chapterhouse")
    data[pos - 4:pos] = (2**31 - 1).to_bytes(4, "little")
    f = C.ParquetFile(bytes(data))
    dev = f.decode_row_group(0)
    assert dev.num_rows == 3000
    dev.download()
