"""SQL extension nodes on the device: binary / unary minus, NOT, IS [NOT] NULL and Kleene AND / OR,
through the C ABI against the oracle's restatement of the arrow-rs kernels (itself checked against pyarrow.compute in
tests/test_extensions_cpu.py) -- interpreter kernel on small batches, run-time specialised select + gather pair on large ones."""
import pyarrow as pa
import pytest

import chapterhouseqe_b200 as C
import ext_cases as X
from chapterhouseqe_b200 import sqlparser_lite as sp
from oracle import compute_value as O

pytestmark = pytest.mark.gpu


def aliases(rb):
    return [[] for _ in rb.schema]


@pytest.mark.parametrize("sql,mask", X.VALUE_CASES)
def test_compute_value_matches_oracle(sql, mask):
    rb = X.table(3000, seed=7)
    expr = sp.parse_expr(sql)
    with O.extensions(mask):
        want = O.compute_value(O.batch_from_arrow(rb), aliases(rb), expr)
    with C.sql_extensions(mask):
        got = C.compute_value(rb, aliases(rb), expr)
    assert got.is_scalar == want.is_scalar, sql
    ok, why = O.arrays_equal(O.array_from_arrow(got.array), want.array)
    assert ok, f"{sql!r}: {why}"


@pytest.mark.parametrize("n", [1, 5000, 200_000])
@pytest.mark.parametrize("sql,mask", X.FILTER_CASES)
def test_filter_matches_oracle(sql, mask, n):
    rb = X.table(n, seed=n)
    expr = sp.parse_expr(sql)
    with O.extensions(mask):
        want = O.filter_record(O.batch_from_arrow(rb), aliases(rb), expr)
    with C.sql_extensions(mask):
        got = C.filter_record(rb, aliases(rb), expr)
    ok, why = O.batches_equal(O.batch_from_arrow(got), want)
    assert ok, f"{sql!r} n={n}: {why}"


@pytest.mark.parametrize("n", [4000, 150_000])
def test_projection_with_extension_nodes_matches_oracle(n):
    rb = X.table(n, seed=n + 1)
    sel = sp.parse_select("select id, -f as nf, a - 1 as am, s is null as sn, not p as np, p or q as pq, -d - f as dd from t "
                          "where a is not null and not (q and p) and k - a > 0")
    with O.extensions(3):
        want = O.project_record(sel["projection"], O.filter_record(O.batch_from_arrow(rb), aliases(rb), sel["selection"]), aliases(rb))
    with C.sql_extensions(3):
        got = C.filter_project_record(sel["selection"], sel["projection"], rb, aliases(rb))
    ok, why = O.batches_equal(O.batch_from_arrow(got), want)
    assert ok, why


def test_checked_negation_and_subtraction_errors_only_on_live_rows():
    rb = pa.RecordBatch.from_arrays([pa.array([5, -2**31, 7, None], type=pa.int32()), pa.array([1, 1, 0, 1], type=pa.int32())], names=["a", "w"])
    with C.sql_extensions(1):
        for sql in ("-a", "a - 1", "0 - a"):
            with pytest.raises(C.ChdbError) as e:
                C.compute_value(rb, aliases(rb), sp.parse_expr(sql))
            assert e.value.kind == "ArithmeticOverflow", sql
        # fused filter + projection: the overflowing row is filtered out, so no error (errors on surviving rows only)
        sel = sp.parse_select("select -a as na from t where a > 0")
        got = C.filter_project_record(sel["selection"], sel["projection"], rb, aliases(rb))
        assert got.column(0).to_pylist() == [-5, -7]


@pytest.mark.parametrize("sql,mask,kind", X.ERROR_CASES)
def test_error_kinds(sql, mask, kind):
    rb = X.table(64, seed=4)
    with C.sql_extensions(mask):
        with pytest.raises(C.ChdbError) as e:
            C.filter_record(rb, aliases(rb), sp.parse_expr(sql))
    assert e.value.kind == kind
