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


def test_pipelined_decode_of_all_row_groups_matches_one_by_one():
    t = PC.sample_table(90000, seed=21)
    data = PC.write(t, row_group_size=11000)
    f = C.ParquetFile(data)
    outs = f.decode_row_groups()
    assert len(outs) == f.num_row_groups == 9
    for i, dev in enumerate(outs):
        want = as_batch(PC.read_row_group(data, i))
        got = dev.download()
        for name in want.schema.names:
            assert got.column(name).equals(want.column(name)), f"row group {i} column {name}"
    part = f.decode_row_groups(3, 2)
    assert [p.num_rows for p in part] == [11000, 11000]
    assert f.decode_row_groups(4, 0) == []
    with pytest.raises(C.ChdbError):
        f.decode_row_groups(7, 5)


def test_big_pages_with_nulls_are_cut_into_segments():
    """Pages far larger than one segment (4096 rows), with nulls: the per-segment ranks come from the host's popcount of
    the definition levels."""
    n = 300_000
    rng = np.random.default_rng(8)
    t = pa.table({"a": pa.array(rng.integers(0, 1 << 30, n).astype(np.int32), mask=rng.random(n) < 0.3),
                  "s": pa.array(np.char.add("k", rng.integers(0, 50, n).astype(str)), mask=rng.random(n) < 0.5),
                  "b": pa.array(rng.random(n) < 0.5, mask=rng.random(n) < 0.01),
                  "runs": pa.array(np.repeat(np.arange(n // 1000, dtype=np.int64), 1000), mask=np.repeat(rng.random(n // 1000) < 0.3, 1000))})
    for kw in (dict(), dict(use_dictionary=False), dict(data_page_version="2.0")):
        check_file(PC.write(t, data_page_size=8 << 20, **kw), t)


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
