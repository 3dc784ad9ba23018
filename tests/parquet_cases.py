"""Parquet files for the decoder tests, written in memory with pyarrow (Arrow C++'s writer) in the formats the reference's
writer produces (parquet-rs defaults, src/bin/create_sample_data.rs:221-224: uncompressed, data page v1, dictionary
encoding with PLAIN fallback) and the neighbouring ones the decoder supports (v2 pages, no dictionary, small pages)."""
import io

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq


def sample_table(n: int, seed: int = 0, nulls: bool = True, wide: int = 0) -> pa.Table:
    """The reference's sample schema (id, value1, value2 -- create_sample_data.rs) plus one column of every supported type."""
    rng = np.random.default_rng(seed)

    def mask(p):
        return (rng.random(n) < p) if (nulls and n) else None

    words = ["", "a", "bb", "hello", "world", "chapterhouse", "x" * 40, "tab\tsep", "ünïcode", "0123456789abcdef"]
    if wide:
        value1 = ["".join(chr(97 + (i * 7 + k) % 26) for k in range(wide)) for i in range(n)]
    else:
        value1 = [words[int(k)] for k in rng.integers(0, len(words), n)]
    unique = [f"row-{i}-{'y' * int(k)}" for i, k in enumerate(rng.integers(0, 30, n))]
    cols = {
        "id": pa.array(np.arange(n, dtype=np.int32)),
        "value1": pa.array(value1, type=pa.utf8(), mask=mask(0.1)),
        "value2": pa.array(rng.uniform(0, 100, n).astype(np.float32), mask=mask(0.1)),
        "d": pa.array(rng.normal(size=n), mask=mask(0.05)),
        "k": pa.array(rng.integers(-2**62, 2**62, n, dtype=np.int64), mask=mask(0.2)),
        "small": pa.array(rng.integers(0, 7, n, dtype=np.int64)),
        "flag": pa.array(rng.random(n) < 0.3, mask=mask(0.15)),
        "u8": pa.array(rng.integers(0, 256, n).astype(np.uint8)),
        "i16": pa.array(rng.integers(-30000, 30000, n).astype(np.int16), mask=mask(0.5)),
        "u32": pa.array(rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)),
        "u64": pa.array(rng.integers(0, 2**63, n, dtype=np.uint64) * 2 + 1, mask=mask(0.01)),
        "unique": pa.array(unique, type=pa.utf8(), mask=mask(0.3)),
        "allnull": pa.array([None] * n, type=pa.int32()),
    }
    fields = [pa.field(name, arr.type, nullable=(name not in ("id", "small", "u8", "u32") or False)) for name, arr in cols.items()]
    return pa.Table.from_arrays(list(cols.values()), schema=pa.schema(fields))


WRITER_VARIANTS = {
    "reference_like": dict(),                                   # dictionary + v1 pages, uncompressed
    "plain": dict(use_dictionary=False),
    "v2": dict(data_page_version="2.0"),
    "v2_plain": dict(data_page_version="2.0", use_dictionary=False),
    "small_pages": dict(data_page_size=2048),
    "small_pages_plain": dict(data_page_size=2048, use_dictionary=False),
    "dict_fallback": dict(dictionary_pagesize_limit=4096),      # dictionary pages overflow -> later pages PLAIN
}


def write(table: pa.Table, row_group_size: int = 1 << 20, **kw) -> bytes:
    buf = io.BytesIO()
    args = dict(compression="NONE", row_group_size=row_group_size, write_statistics=True)
    args.update(kw)
    pq.write_table(table, buf, **args)
    return buf.getvalue()


def read_row_group(data: bytes, i: int) -> pa.Table:
    return pq.ParquetFile(io.BytesIO(data)).read_row_group(i)
