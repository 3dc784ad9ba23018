"""Pins the CPU oracle against the reference's own golden vectors
(record_utils/test_compute_value.rs, test_filter_record.rs, test_arrow_compute_behavior.rs)
and against the hand-derived KATs in tests/kats.py.  CPU only."""
import numpy as np
import pyarrow as pa
import pytest

import harness as H
import kats
import refcases
from chapterhouseqe_b200 import sqlparser_lite as sp
from oracle import compute_value as O

ORACLE = H.OracleImpl()


@pytest.mark.parametrize("case", refcases.GOLDEN, ids=[c[0] for c in refcases.GOLDEN])
def test_reference_golden(case):
    name, cite, schema, columns, aliases, kind, sql, expected = case
    spec = dict(name=name, schema=schema, cols=[columns[n] for n, _, _ in schema], aliases=aliases or None,
                kind=kind, sql=sql,
                expect=("ok", expected if kind == "value" else [expected[n] for n, _, _ in schema]))
    H.check_case(ORACLE, spec)


def test_reference_table_alias():
    c = refcases.TABLE_ALIAS_CASE
    H.check_case(ORACLE, dict(name="test_table_alias", schema=c["schema"], cols=c["columns"],
                              aliases=c["table_aliases"], kind="value", sql=c["query"],
                              expect=("ok", c["expected"])))


def test_scalar_plus_scalar_is_len1_scalar_datum():
    # test_arrow_compute_behavior.rs:48-64 + compute_value.rs:43-48 (new_binary_op recomputes is_scalar)
    rb = H.make_batch([("x", "int32", False)], [[5, 6, 7]])
    d = O.compute_value(O.batch_from_arrow(rb), [[]], sp.parse_expr("1 + 2"))
    assert d.array.length == 1 and d.array.to_pylist() == [3] and d.is_scalar is True


def test_scalar_plus_array():
    # test_arrow_compute_behavior.rs:67-78
    rb = H.make_batch([("x", "int32", False)], [[5, 4, 3, 2, 1, 0]])
    d = O.compute_value(O.batch_from_arrow(rb), [[]], sp.parse_expr("5 + x"))
    assert d.array.to_pylist() == [10, 9, 8, 7, 6, 5] and d.is_scalar is False


def test_duplicate_column_names_first_wins():
    # test_arrow_compute_behavior.rs:81-108
    rb = H.make_batch([("t", "utf8", False), ("t", "utf8", False)], [["hello", "x"], ["a", "b"]])
    d = O.compute_value(O.batch_from_arrow(rb), [[], []], sp.parse_expr("t = 'hello'"))
    assert d.array.to_pylist() == [True, False]


def test_u32_to_f32_cast_roundoff():
    # test_arrow_compute_behavior.rs:111-126
    src, want = refcases.U32_TO_F32
    a = O.Array.from_pylist("uint32", src)
    got = O._cast(a, "float32")
    assert got.values.tolist() == want
    # and the same through pyarrow's unchecked cast (independent implementation)
    assert pa.array(src, pa.uint32()).cast(pa.float32(), safe=False).to_pylist() == want


@pytest.mark.parametrize("case", kats.KATS, ids=[c["name"] for c in kats.KATS])
def test_kats(case):
    H.check_case(ORACLE, case)


def test_parse_f32_is_correctly_rounded():
    # Rust's f32 parser rounds the decimal once; going through f64 first can double-round.
    assert O.parse_f32("0.1").view(np.uint32) == 0x3DCCCCCD
    assert O.parse_f32("16777217.0") == np.float32(16777216.0)
    # 1 + 2^-24 + 2^-60 is just above the tie, so correctly rounded is 1 + 2^-23; via f64 the
    # 2^-60 is lost first and the tie then rounds to even (1.0).
    s = "1.0000000596046447762489318847656250008673617379884"
    assert float(s) == 1.0 + 2.0 ** -24
    assert O.parse_f32(s).view(np.uint32) == 0x3F800001
    assert O.parse_f32("340282350000000000000000000000000000000.0") == np.finfo(np.float32).max
    assert np.isinf(O.parse_f32("3402823700000000000000000000000000000000.0"))


def test_sample_queries_parse_and_run():
    rng = np.random.default_rng(0xC4DB0001)
    n = 100
    ids = np.arange(n, dtype=np.int32)
    v1 = ["".join(chr(97 + c) for c in rng.integers(0, 26, 8)) for _ in range(n)]
    v2 = rng.uniform(0, 100, n).astype(np.float32)
    rb = pa.RecordBatch.from_arrays([pa.array(ids), pa.array(v1), pa.array(v2)],
                                    schema=pa.schema([pa.field("id", pa.int32(), False),
                                                      pa.field("value1", pa.utf8(), False),
                                                      pa.field("value2", pa.float32(), False)]))
    b = O.batch_from_arrow(rb)
    al = [[], [], []]
    for name, sql in refcases.SAMPLE_QUERIES.items():
        sel = sp.parse_select(sql)
        out = O.project_record(sel["projection"], O.filter_record(b, al, sel["selection"]), al)
        if name == "simple_q1":
            assert out.num_rows == 25 and [f.name for f in out.fields] == ["id", "value1", "value2"]
        if name == "simple_q5":
            assert out.columns[0].to_pylist() == list(range(0, 100, 2))
        if name == "simple_q4":
            assert out.num_rows == 74
            assert [f.name for f in out.fields] == ["id", "value1", "id_plus_10", "value2", "value3", "value4",
                                                    "value5"]
            assert [f.dtype for f in out.fields] == ["int32", "utf8", "float32", "float32", "float32", "float32",
                                                     "int32"]
        if name == "readme":
            assert out.num_rows == int((v2 > np.float32(10.0)).sum())


def test_oracle_vs_pyarrow_differential():
    """Independent implementation (Arrow C++). NaN / signed-zero free inputs only: pyarrow compares
    floats the IEEE way while arrow-rs uses totalOrder."""
    import pyarrow.compute as pc
    rng = np.random.default_rng(7)
    n = 5000
    ids = rng.integers(-1000, 1000, n).astype(np.int32)
    k = rng.integers(-2**40, 2**40, n).astype(np.int64)
    f = (rng.uniform(1, 100, n)).astype(np.float32)
    d = rng.normal(0, 1, n) + 3.0
    fmask = rng.random(n) < 0.1
    dmask = rng.random(n) < 0.05
    s = ["".join(chr(97 + c) for c in rng.integers(0, 26, rng.integers(0, 12))) for _ in range(n)]
    rb = pa.RecordBatch.from_arrays(
        [pa.array(ids), pa.array(k), pa.array(f, mask=fmask), pa.array(d, mask=dmask), pa.array(s)],
        schema=pa.schema([pa.field("id", pa.int32(), False), pa.field("k", pa.int64(), False),
                          pa.field("f", pa.float32(), True), pa.field("d", pa.float64(), True),
                          pa.field("s", pa.utf8(), False)]))
    b = O.batch_from_arrow(rb)
    al = [[]] * 5
    C = {n_: rb.column(i) for i, n_ in enumerate(rb.schema.names)}
    f32 = lambda x: pa.scalar(np.float32(x), pa.float32())  # noqa: E731
    checks = {
        "id + 7": pc.add_checked(C["id"], pa.scalar(7, pa.int32())),
        "id * id": pc.multiply_checked(C["id"], C["id"]),
        "k + id": pc.add_checked(C["k"], C["id"].cast(pa.int64())),
        "f * 2.5": pc.multiply(C["f"], f32(2.5)),
        "f / d": pc.divide(C["f"].cast(pa.float64()), C["d"]),
        "id + f": pc.add(C["id"].cast(pa.float32(), safe=False), C["f"]),
        "f > 10.0": pc.greater(C["f"], f32(10.0)),
        "d <= 3.0": pc.less_equal(C["d"], pa.scalar(float(np.float32(3.0)))),
        "id = 5": pc.equal(C["id"], pa.scalar(5, pa.int32())),
        "s < 'm'": pc.less(C["s"], pa.scalar("m")),
        "s = 'a'": pc.equal(C["s"], pa.scalar("a")),
        "f > 10.0 and d < 3.0": pc.and_(pc.greater(C["f"], f32(10.0)),
                                        pc.less(C["d"], pa.scalar(float(np.float32(3.0))))),
        "id > 0 or f < 50.0": pc.or_(pc.greater(C["id"], pa.scalar(0, pa.int32())), pc.less(C["f"], f32(50.0))),
    }
    for sql, want in checks.items():
        got = O.compute_value(b, al, sp.parse_expr(sql)).array
        ok, why = O.arrays_equal(got, O.array_from_arrow(want))
        assert ok, f"{sql}: {why}"
    # filter: Table.filter drops NULL mask rows and rebuilds utf8 offsets from 0, like arrow-rs
    for sql, mask in [("id % 2 = 0", pc.equal(pc.subtract(C["id"], pc.multiply(pc.divide(C["id"], 2), 2)), 0)),
                      ("f > 10.0 and d < 3.0", checks["f > 10.0 and d < 3.0"]),
                      ("s < 'm'", checks["s < 'm'"])]:
        got = O.filter_record(b, al, sp.parse_expr(sql))
        want = O.batch_from_arrow(rb.filter(mask))
        ok, why = O.batches_equal(got, want)
        assert ok, f"filter {sql}: {why}"


def test_get_record_table_aliases():
    rb = H.make_batch([("a", "int32", False), ("b", "int32", False)], [[1], [2]])
    b = O.batch_from_arrow(rb)
    op = {"Producer": {"task": {"TableFunc": {"alias": "t", "func_name": "read_files", "args": [],
                                              "max_rows_per_batch": 10000}},
                       "outbound_exchange_id": "x", "inbound_exchange_ids": []}}
    assert O.get_record_table_aliases(op, b) == [["t"], ["t"]]
    op["Producer"]["task"]["TableFunc"]["alias"] = None
    assert O.get_record_table_aliases(op, b) == [[], []]
    with pytest.raises(O.OracleError) as ei:
        O.get_record_table_aliases({"Producer": {"task": {"Filter": {"expr": {}}}}}, b)
    assert ei.value.kind == "OperatorTaskTypeDoesNotHaveAnAliasField"


def test_golden_fixture_file_matches_refcases():
    """tests/golden/reference_vectors.json is the serialised form of refcases.py (made by make_reference_vectors.py)."""
    import json
    import os
    import refcases
    doc = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))
    by_name = {c["name"]: c for c in doc["cases"]}
    assert len(by_name) == len(refcases.GOLDEN) + 1
    for name, cite, schema, cols, aliases, kind, query, expected in refcases.GOLDEN:
        c = by_name[name]
        assert c["reference"] == cite and c["query"] == query and c["kind"] == kind and c["columns"] == cols
        want = {"dtype": expected[0], "values": expected[1]} if kind == "value" else expected
        assert c["expected"] == want
