"""CUDA parity tests: every call goes through the C ABI (libchdb_gpu.so) and is compared with
(1) the reference's own golden vectors, (2) the hand-derived KATs, (3) the CPU oracle on seeded
random inputs, bit for bit (SURVEY.md 8c comparison rule)."""
import numpy as np
import pyarrow as pa
import pytest

import chapterhouseqe_b200 as C
import harness as H
import kats
import refcases
from chapterhouseqe_b200 import sqlparser_lite as sp
from oracle import compute_value as O

pytestmark = pytest.mark.gpu


class GpuImpl:
    name = "gpu"

    def compute_value(self, rb, aliases, expr) -> O.Array:
        d = C.compute_value(rb, aliases, expr)
        return O.array_from_arrow(d.array)

    def filter_record(self, rb, aliases, expr) -> O.Batch:
        return O.batch_from_arrow(C.filter_record(rb, aliases, expr))

    def project_record(self, fields, rb, aliases) -> O.Batch:
        return O.batch_from_arrow(C.project_record(fields, rb, aliases))

    def filter_project_record(self, expr, fields, rb, aliases) -> O.Batch:
        return O.batch_from_arrow(C.filter_project_record(expr, fields, rb, aliases))


GPU = GpuImpl()
ORACLE = H.OracleImpl()


def test_native_library_is_what_runs():
    lib = C.load_library()
    assert lib.chdb_compiled_arch() == b"sm_100a"
    ctx = C.default_context()
    before = ctx.launch_count
    rb = H.make_batch([("id", "int32", False)], [[1, 2, 3, 4]])
    out = C.filter_record(rb, [[]], sp.parse_expr("id % 2 = 0"))
    assert out.column(0).to_pylist() == [2, 4]
    assert ctx.launch_count == before + 2   # workspace zeroing + the (fused, at this size) stream kernel


@pytest.mark.parametrize("case", refcases.GOLDEN, ids=[c[0] for c in refcases.GOLDEN])
def test_reference_golden(case):
    name, cite, schema, columns, aliases, kind, sql, expected = case
    spec = dict(name=name, schema=schema, cols=[columns[n] for n, _, _ in schema], aliases=aliases or None, kind=kind,
                sql=sql, expect=("ok", expected if kind == "value" else [expected[n] for n, _, _ in schema]))
    H.check_case(GPU, spec)


def test_reference_table_alias():
    c = refcases.TABLE_ALIAS_CASE
    H.check_case(GPU, dict(name="test_table_alias", schema=c["schema"], cols=c["columns"], aliases=c["table_aliases"],
                           kind="value", sql=c["query"], expect=("ok", c["expected"])))


def test_scalar_datum_flags():
    rb = H.make_batch([("x", "int32", False)], [[5, 6, 7]])
    d = C.compute_value(rb, [[]], sp.parse_expr("1 + 2"))
    assert d.array.to_pylist() == [3] and d.is_scalar is True
    d = C.compute_value(rb, [[]], sp.parse_expr("5 + x"))
    assert d.array.to_pylist() == [10, 11, 12] and d.is_scalar is False
    d = C.compute_value(rb, [[]], sp.parse_expr("x"))
    assert d.array.to_pylist() == [5, 6, 7] and d.is_scalar is False


@pytest.mark.parametrize("case", kats.KATS, ids=[c["name"] for c in kats.KATS])
def test_kats(case):
    H.check_case(GPU, case)


# ------------------------------------------------------------------------------------------------
# randomised differential tests against the oracle
# ------------------------------------------------------------------------------------------------
def _rand_strings(rng, n, lo, hi):
    lens = rng.integers(lo, hi + 1, n)
    data = rng.integers(97, 123, int(lens.sum()), dtype=np.uint8)
    offs = np.zeros(n + 1, dtype=np.int32)
    offs[1:] = np.cumsum(lens)
    return pa.Array.from_buffers(pa.utf8(), n, [None, pa.py_buffer(offs.tobytes()), pa.py_buffer(data.tobytes())])


def _with_nulls(rng, arr, frac):
    if frac <= 0:
        return arr
    mask = rng.random(len(arr)) < frac
    return pa.array(arr.to_numpy(zero_copy_only=False) if arr.type != pa.utf8() else arr.to_pylist(), type=arr.type,
                    mask=mask)


def make_mixed_batch(n, seed, null_frac=0.1, str_lo=0, str_hi=12):
    """large_simple-shaped synthetic rows with the extra column types of BASELINE config 2."""
    rng = np.random.default_rng(seed)
    cols = {
        "id": pa.array(np.arange(n, dtype=np.int32)),
        "k": pa.array(rng.integers(-2**31, 2**31, n).astype(np.int64)),
        "value2": _with_nulls(rng, pa.array(rng.uniform(0, 100, n).astype(np.float32)), null_frac),
        "d": _with_nulls(rng, pa.array(rng.normal(0, 1, n)), null_frac / 2),
        "value1": _rand_strings(rng, n, str_lo, str_hi),
        "flag": _with_nulls(rng, pa.array(rng.random(n) < 0.5), null_frac),
        "small": pa.array(rng.integers(-100, 100, n).astype(np.int16)),
        "u": pa.array(rng.integers(0, 100, n).astype(np.uint8)),
        "s2": _with_nulls(rng, _rand_strings(rng, n, 0, 5), null_frac),
    }
    fields = [pa.field(k, v.type, k in ("value2", "d", "flag", "s2")) for k, v in cols.items()]
    return pa.RecordBatch.from_arrays(list(cols.values()), schema=pa.schema(fields))


PREDICATES = [
    "id % 2 = 0",
    "(id % 2 = 0 and value2 > 10.0) or d < 0.5",
    "value2 > 90.0",
    "id > 25",
    "id < 0",
    "value1 < 'm'",
    "value1 = s2",
    "flag = true and small * 3 > u",
    "k % 7 = 3 or small / 5 = 2",
    "d * d + 1.5 > value2 / 7",
    "flag or id % 3 = 0",
    "u + small >= 50 and k > id",
]


@pytest.mark.parametrize("n", [1, 3, 127, 2048, 2049, 6151, 40000])
@pytest.mark.parametrize("pi", range(len(PREDICATES)))
def test_filter_matches_oracle(n, pi):
    rb = make_mixed_batch(n, seed=1000 + n)
    al = [[] for _ in rb.schema]
    expr = sp.parse_expr(PREDICATES[pi])
    want = O.filter_record(O.batch_from_arrow(rb), al, expr)
    got = O.batch_from_arrow(C.filter_record(rb, al, expr))
    ok, why = O.batches_equal(got, want)
    assert ok, f"n={n} {PREDICATES[pi]!r}: {why}"


PROJECTIONS = [
    "select id, value1, id + 10.0 as id_plus_10, (value2 + 10) / 100 as value2, 1.0 / id as value3, "
    "1.0 / (small * small) as value4, small * small as value5 from t",
    "select *, k * 3 as k3, d / value2 as ratio, value2 > 50.0 as big, flag and id % 2 = 0 as both from t",
    "select k % 1000 as km, k / 3 as kd, d % 0.25 as dm, u + u as uu, small + u as su, value1 < s2 as lt from t",
    "select (id + 1) * (small + 2) as a, (d + 1.0) / (value2 + 1.0) as b from t",
]


@pytest.mark.parametrize("n", [1, 5, 2048, 4100, 30011])
@pytest.mark.parametrize("qi", range(len(PROJECTIONS)))
def test_project_matches_oracle(n, qi):
    rb = make_mixed_batch(n, seed=2000 + n)
    al = [[] for _ in rb.schema]
    items = sp.parse_select(PROJECTIONS[qi])["projection"]
    want = O.project_record(items, O.batch_from_arrow(rb), al)
    got = O.batch_from_arrow(C.project_record(items, rb, al))
    ok, why = O.batches_equal(got, want)
    assert ok, f"n={n} q{qi}: {why}"


@pytest.mark.parametrize("n", [1, 2049, 30011])
@pytest.mark.parametrize("qi", range(len(PROJECTIONS)))
@pytest.mark.parametrize("pred", ["id % 2 = 0", "(id % 2 = 0 and value2 > 10.0) or d < 0.5", "value1 < 'c'"])
def test_fused_filter_project_matches_oracle(n, qi, pred):
    rb = make_mixed_batch(n, seed=3000 + n)
    al = [[] for _ in rb.schema]
    sel = sp.parse_select(PROJECTIONS[qi] + " where " + pred)
    b = O.batch_from_arrow(rb)
    want = O.project_record(sel["projection"], O.filter_record(b, al, sel["selection"]), al)
    got = O.batch_from_arrow(C.filter_project_record(sel["selection"], sel["projection"], rb, al))
    ok, why = O.batches_equal(got, want)
    assert ok, f"n={n} q{qi} {pred!r}: {why}"


@pytest.mark.parametrize("L,n", [(100, 1), (100, 5000), (100, 20000), (8, 50000), (37, 9999), (1000, 3000)])
@pytest.mark.parametrize("pred", ["id > 25", "id % 2 = 0", "id % 7 = 1"])
def test_wide_strings(L, n, pred):
    """only_wide_strings_query.sql-shaped: fixed-length L strings (create_sample_data.rs:157-230)."""
    rng = np.random.default_rng(L * 7 + n)
    rb = pa.RecordBatch.from_arrays(
        [pa.array(np.arange(n, dtype=np.int32)), _rand_strings(rng, n, L, L),
         pa.array(rng.uniform(0, 100, n).astype(np.float32))],
        schema=pa.schema([pa.field("id", pa.int32(), False), pa.field("value1", pa.utf8(), False),
                          pa.field("value2", pa.float32(), False)]))
    al = [[], [], []]
    expr = sp.parse_expr(pred)
    want = O.filter_record(O.batch_from_arrow(rb), al, expr)
    got = O.batch_from_arrow(C.filter_record(rb, al, expr))
    ok, why = O.batches_equal(got, want)
    assert ok, f"L={L} n={n} {pred!r}: {why}"


def test_odd_length_strings_and_sliced_input():
    rng = np.random.default_rng(5)
    n = 10000
    full = pa.RecordBatch.from_arrays(
        [pa.array(np.arange(n, dtype=np.int32)), _with_nulls(rng, _rand_strings(rng, n, 0, 33), 0.2),
         _with_nulls(rng, pa.array(rng.random(n) < 0.3), 0.1)],
        schema=pa.schema([pa.field("id", pa.int32(), False), pa.field("s", pa.utf8(), True),
                          pa.field("b", pa.bool_(), True)]))
    for start, length in [(0, n), (3, 5000), (1001, 4099), (9999, 1)]:
        rb = full.slice(start, length)   # non-zero Arrow offsets, bit offsets not multiples of 8
        al = [[], [], []]
        for pred in ["id % 3 = 0", "b", "s < 'k'"]:
            expr = sp.parse_expr(pred)
            want = O.filter_record(O.batch_from_arrow(rb), al, expr)
            got = O.batch_from_arrow(C.filter_record(rb, al, expr))
            ok, why = O.batches_equal(got, want)
            assert ok, f"slice({start},{length}) {pred!r}: {why}"


def test_empty_batch():
    rb = make_mixed_batch(0, seed=1)
    al = [[] for _ in rb.schema]
    out = C.filter_record(rb, al, sp.parse_expr("id % 2 = 0"))
    assert out.num_rows == 0 and out.schema.names == rb.schema.names
    out = C.project_record(sp.parse_select("select id + 1 as x, value1 from t")["projection"], rb, al)
    assert out.num_rows == 0 and out.schema.names == ["x", "value1"]


def test_float_specials_match_oracle():
    """NaN payloads, signed zeros, infinities, denormals through arithmetic and totalOrder compares."""
    specials32 = [0x00000000, 0x80000000, 0x7F800000, 0xFF800000, 0x7FC00000, 0xFFC00000, 0x7F800001, 0xFFC12345,
                  0x00000001, 0x807FFFFF, 0x3F800000, 0xBF800000, 0x7F7FFFFF, 0x00800000]
    a = [("bits", x) for x in specials32 for _ in specials32]
    b = [("bits", y) for _ in specials32 for y in specials32]
    rb = H.make_batch([("a", "float32", False), ("b", "float32", False)], [a, b])
    al = [[], []]
    for sql in ["a + b", "a * b", "a / b", "a % b", "a < b", "a = b", "a >= b", "a <> b", "a + b > a * b", "a and b",
                "a / b < 5.0"]:
        want = O.compute_value(O.batch_from_arrow(rb), al, sp.parse_expr(sql)).array
        got = O.array_from_arrow(C.compute_value(rb, al, sp.parse_expr(sql)).array)
        ok, why = O.arrays_equal(got, want)
        assert ok, f"{sql}: {why}"
    specials64 = [0x0, 0x8000000000000000, 0x7FF0000000000000, 0xFFF0000000000000, 0x7FF8000000000000,
                  0xFFF8000000000000, 0x7FF0000000000001, 0x1, 0x3FF0000000000000, 0xBFF0000000000000,
                  0x7FEFFFFFFFFFFFFF]
    a = [("bits", x) for x in specials64 for _ in specials64]
    b = [("bits", y) for _ in specials64 for y in specials64]
    rb = H.make_batch([("a", "float64", False), ("b", "float64", False), ("f", "float32", False)],
                      [a, b, [("bits", specials32[i % len(specials32)]) for i in range(len(a))]])
    al = [[], [], []]
    for sql in ["a + b", "a * b", "a / b", "a % b", "a < b", "a = b", "a >= b", "a + f", "f < a", "a or f"]:
        want = O.compute_value(O.batch_from_arrow(rb), al, sp.parse_expr(sql)).array
        got = O.array_from_arrow(C.compute_value(rb, al, sp.parse_expr(sql)).array)
        ok, why = O.arrays_equal(got, want)
        assert ok, f"{sql}: {why}"


def test_device_resident_pipeline_and_program_reuse():
    """read_files -> [exchange] -> filter -> [exchange] -> materialize with device-resident batches:
    the filter output feeds the projection without crossing PCIe; one compiled program serves many batches."""
    ctx = C.default_context()
    sel = sp.parse_select(refcases.SAMPLE_QUERIES["simple_q4"])
    schema = make_mixed_batch(1, 0).select(["id", "value1", "value2"]).schema
    fprog = C.Program.compile_filter(sel["selection"], schema)
    pprog = C.Program.compile_project(sel["projection"], schema)
    for rec_id, n in enumerate([100, 4096, 10000, 33]):
        rb = make_mixed_batch(n, seed=rec_id, null_frac=0.0, str_lo=8, str_hi=8).select(["id", "value1", "value2"])
        rb = rb.slice(0, min(n, 46000))
        al = [[], [], []]
        dev_in = C.DeviceBatch.upload(rb, ctx)
        dev_filtered = dev_in.run(fprog)
        dev_out = dev_filtered.run(pprog)
        got = O.batch_from_arrow(dev_out.download())
        b = O.batch_from_arrow(rb)
        want = O.project_record(sel["projection"], O.filter_record(b, al, sel["selection"]), al)
        ok, why = O.batches_equal(got, want)
        assert ok, f"record {rec_id}: {why}"
        assert dev_filtered.num_rows == want.num_rows


def test_overflow_error_reports_through_device_batch():
    ctx = C.default_context()
    rb = H.make_batch([("id", "int32", False)], [list(range(100000, 100010))])
    prog = C.Program.compile_project(sp.parse_select("select id * id as sq from t")["projection"], rb.schema)
    out = C.DeviceBatch.upload(rb, ctx).run(prog)
    with pytest.raises(C.ChdbError) as ei:
        out.check()
    assert ei.value.kind == "ArithmeticOverflow"
    with pytest.raises(C.ChdbError):
        out.download()


def test_round_trip_properties_at_scale():
    """BASELINE-sized shape through size-independent properties: a 4M-row batch (2^22, the device-native
    batch size) where id % 2 = 0 keeps exactly the even ids in order and the other columns stay aligned."""
    n = 1 << 22
    rng = np.random.default_rng(42)
    ids = np.arange(n, dtype=np.int32)
    v2 = rng.uniform(0, 100, n).astype(np.float32)
    s = _rand_strings(rng, n, 8, 8)
    rb = pa.RecordBatch.from_arrays([pa.array(ids), s, pa.array(v2)],
                                    schema=pa.schema([pa.field("id", pa.int32(), False),
                                                      pa.field("value1", pa.utf8(), False),
                                                      pa.field("value2", pa.float32(), False)]))
    out = C.filter_record(rb, [[], [], []], sp.parse_expr("id % 2 = 0"))
    assert out.num_rows == n // 2
    oid = out.column(0).to_numpy()
    assert np.array_equal(oid, ids[::2])
    assert np.array_equal(out.column(2).to_numpy().view(np.uint32), v2[::2].view(np.uint32))
    offs = np.frombuffer(out.column(1).buffers()[1], dtype=np.int32)[: n // 2 + 1]
    assert np.array_equal(offs, np.arange(n // 2 + 1, dtype=np.int32) * 8)
    src = np.frombuffer(s.buffers()[2], dtype=np.uint8).reshape(n, 8)
    dst = np.frombuffer(out.column(1).buffers()[2], dtype=np.uint8)[: n // 2 * 8].reshape(n // 2, 8)
    assert np.array_equal(dst, src[::2])
    # idempotence: filtering the result again with the same predicate changes nothing
    again = C.filter_record(out, [[], [], []], sp.parse_expr("id % 2 = 0"))
    assert again.equals(out)


@pytest.mark.parametrize("jit", ["0", "always"])
def test_many_tiles_per_cta(jit, monkeypatch):
    """700k rows = 684 tiles on at most 148 persistent CTAs: every CTA walks its shared-memory ring several
    times (stage reuse, mbarrier phase flips), with nulls, Utf8 and a fused projection, against the oracle."""
    monkeypatch.setenv("CHDB_JIT", jit)
    rb = make_mixed_batch(700_000, seed=991)
    al = [[] for _ in rb.schema]
    expr = sp.parse_expr("(id % 2 = 0 and value2 > 10.0) or d < 0.5")
    want = O.filter_record(O.batch_from_arrow(rb), al, expr)
    ok, why = O.batches_equal(O.batch_from_arrow(C.filter_record(rb, al, expr)), want)
    assert ok, why
    sel = sp.parse_select(PROJECTIONS[1] + " where id % 3 = 0")
    b = O.batch_from_arrow(rb)
    want = O.project_record(sel["projection"], O.filter_record(b, al, sel["selection"]), al)
    got = O.batch_from_arrow(C.filter_project_record(sel["selection"], sel["projection"], rb, al))
    ok, why = O.batches_equal(got, want)
    assert ok, why


@pytest.mark.parametrize("jit", ["0", "always"])
def test_schema_wider_than_the_stage_ring(jit, monkeypatch):
    """20 Int64 columns (+ validity on half of them) need 164 KB per 1024-row stage: a ring of at least four stages does
    not fit in shared memory, so the planner leaves the largest buffers in global memory (read through the same
    column table, prefetched into L2 by the producer warps).  Results must not depend on what is staged."""
    monkeypatch.setenv("CHDB_JIT", jit)
    rng = np.random.default_rng(5)
    n = 150_000
    cols, fields = [], []
    for c in range(20):
        v = rng.integers(-1000, 1000, n, dtype=np.int64)
        nullable = c % 2 == 1
        cols.append(pa.array(v, mask=(rng.random(n) < 0.2) if nullable else None))
        fields.append(pa.field(f"c{c}", pa.int64(), nullable))
    rb = pa.RecordBatch.from_arrays(cols, schema=pa.schema(fields))
    al = [[] for _ in rb.schema]
    expr = sp.parse_expr("c0 % 3 = 0 or c1 > 10")
    want = O.filter_record(O.batch_from_arrow(rb), al, expr)
    ok, why = O.batches_equal(O.batch_from_arrow(C.filter_record(rb, al, expr)), want)
    assert ok, why
