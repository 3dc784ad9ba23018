"""Host side of the Parquet -> device decoder (csrc/parquet.inc): footer, page headers and run headers are parsed on
the CPU and need no GPU."""
import pyarrow as pa
import pytest

import chapterhouseqe_b200 as C
import parquet_cases as PC


@pytest.mark.parametrize("variant", sorted(PC.WRITER_VARIANTS))
def test_footer_and_page_walk(variant):
    t = PC.sample_table(30000, seed=1)
    data = PC.write(t, row_group_size=12000, **PC.WRITER_VARIANTS[variant])
    f = C.ParquetFile(data)
    assert f.num_rows == 30000 and f.num_row_groups == 3
    assert [f.row_group_num_rows(i) for i in range(3)] == [12000, 12000, 6000]
    assert f.schema.equals(t.schema)
    for i in range(3):
        pages, runs = f.check_row_group(i)
        assert pages >= t.num_columns
        assert runs >= 1   # definition levels of the OPTIONAL columns at least


def test_empty_table_and_bad_files():
    f = C.ParquetFile(PC.write(PC.sample_table(0)))
    assert f.num_rows == 0 and f.schema.names[0] == "id"
    with pytest.raises(C.ChdbError, match="not a Parquet file"):
        C.ParquetFile(b"PAR1 nope")
    data = bytearray(PC.write(PC.sample_table(100)))
    data[-8:-4] = (2**31 - 1).to_bytes(4, "little")
    with pytest.raises(C.ChdbError, match="footer length"):
        C.ParquetFile(bytes(data))


def test_unsupported_features_are_not_implemented():
    t = PC.sample_table(1000)
    f = C.ParquetFile(PC.write(t, compression="snappy"))
    with pytest.raises(C.ChdbError, match="compression codec") as e:
        f.check_row_group(0)
    assert e.value.kind == "NotImplemented"
    nested = pa.table({"l": pa.array([[1, 2], [3]], type=pa.list_(pa.int32()))})
    with pytest.raises(C.ChdbError, match="nested") as e:
        C.ParquetFile(PC.write(nested))
    assert e.value.kind == "NotImplemented"
    ts = pa.table({"t": pa.array([1, 2], type=pa.timestamp("us"))})
    with pytest.raises(C.ChdbError, match="not supported"):
        C.ParquetFile(PC.write(ts))
    delta = pa.table({"x": pa.array(range(1000), type=pa.int32())})
    f = C.ParquetFile(PC.write(delta, use_dictionary=False, column_encoding={"x": "DELTA_BINARY_PACKED"}))
    with pytest.raises(C.ChdbError, match="data page encoding"):
        f.check_row_group(0)
