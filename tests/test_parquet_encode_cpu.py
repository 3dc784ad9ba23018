"""Host side of the device Parquet encoder (csrc/parquet_encode.inc) without a GPU: tests/host/pqe_host_test.cu builds a file
with the encoder's own page-header / footer writer (values packed on the host the way the kernels pack them) and pyarrow's reader
(Arrow C++) must read back exactly what went in: REQUIRED Int32, OPTIONAL Float64 with a validity-bitmap definition-level run,
OPTIONAL Utf8 with an RLE run, two row groups of two pages per column, 21 rows per page (not a multiple of 8)."""
import os
import subprocess

import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from chapterhouseqe_b200 import build as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_binary(tmp_path_factory):
    B.build()
    out = tmp_path_factory.mktemp("pqe")
    exe, obj = str(out / "pqe_host_test"), str(out / "pqe_host_test.o")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    arch = ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call([nvcc, *arch, "-std=c++17", "-O1", "-fmad=false", "-w", "-x", "cu", "-c",
                           os.path.join(ROOT, "tests", "host", "pqe_host_test.cu"), "-o", obj])
    objs = [os.path.join(B.HERE, "build", n + ".o") for n in ("kernels", "lower", "jit")]
    subprocess.check_call([nvcc, *arch, obj, *objs, "-ldl", "-o", exe])
    return exe, str(out / "written.parquet")


def test_thrift_writer_and_layout_read_back_by_arrow_cpp(host_binary):
    exe, path = host_binary
    r = subprocess.run([exe, path], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "own reader: 2 row groups, 84 rows, 3 cols" in r.stdout
    f = pq.ParquetFile(path)
    assert f.metadata.num_row_groups == 2 and f.metadata.num_rows == 84
    assert f.metadata.created_by == "chapterhouseqe_b200 device encoder"
    assert f.schema_arrow == pa.schema([pa.field("id", pa.int32(), False), pa.field("d", pa.float64()), pa.field("s", pa.utf8())])
    t = f.read()
    n = 21
    want_id, want_d, want_s = [], [], []
    for g in range(2):
        for b in range(2):
            want_id += [g * 1000 + b * 100 + i for i in range(n)]
            want_d += [(i * 0.5 + b) if i % 3 else None for i in range(n)]
            want_s += [f"row{i * 7 + b}" for i in range(n)]
    assert t.column("id").to_pylist() == want_id
    assert t.column("d").to_pylist() == want_d
    assert t.column("s").to_pylist() == want_s
    assert f.read_row_group(1).num_rows == 42
    col = f.metadata.row_group(0).column(1)
    assert col.compression == "UNCOMPRESSED" and "PLAIN" in col.encodings and col.num_values == 42
