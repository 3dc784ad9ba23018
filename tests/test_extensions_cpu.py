"""SQL nodes beyond the reference's compute_value (the reference's README.md:44-74 lists them as
missing): the oracle's restatement of the arrow-rs kernels against pyarrow.compute (Arrow C++: subtract_checked,
negate_checked, invert, is_null, and_kleene, or_kleene), and the host lowering with and without the extension mask.  No GPU."""
import ctypes

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

import chapterhouseqe_b200 as C
import ext_cases as X
from chapterhouseqe_b200 import sqlparser_lite as sp
from oracle import compute_value as O


def oracle_value(rb, sql, mask):
    with O.extensions(mask):
        return O.compute_value(O.batch_from_arrow(rb), [[] for _ in rb.schema], sp.parse_expr(sql)).array


def same(arr: O.Array, want: pa.Array):
    got = O.array_to_arrow(arr) if hasattr(O, "array_to_arrow") else None
    if got is None:
        pytest.skip("oracle has no Arrow export")
    assert got.type == want.type, (got.type, want.type)
    assert got.null_count == want.null_count
    # floats bit for bit (NaN sign included), everything else by equality
    if pa.types.is_floating(want.type):
        ut = np.uint32 if want.type == pa.float32() else np.uint64
        v = np.asarray(want.is_valid())
        g = got.to_numpy(zero_copy_only=False).view(ut)[v]
        w = want.to_numpy(zero_copy_only=False).view(ut)[v]
        assert (g == w).all()
    else:
        assert got.equals(want)


def test_oracle_extension_kernels_match_arrow_cpp():
    rb = X.table(5000, seed=1)
    col = {n: rb.column(n) for n in rb.schema.names}
    checks = [
        ("a - 7", 1, pc.subtract_checked(col["a"], pa.scalar(7, pa.int32()))),
        ("k - a", 1, pc.subtract_checked(col["k"], pc.cast(col["a"], pa.int64()))),
        ("d - f", 1, pc.subtract(col["d"], pc.cast(col["f"], pa.float64()))),
        ("-a", 1, pc.negate_checked(col["a"])),
        ("-k", 1, pc.negate_checked(col["k"])),
        ("-f", 1, pc.negate(col["f"])),
        ("-d", 1, pc.negate(col["d"])),
        ("not p", 1, pc.invert(col["p"])),
        ("a is null", 1, pc.is_null(col["a"])),
        ("s is not null", 1, pc.is_valid(col["s"])),
        ("p and q", 2, pc.and_kleene(col["p"], col["q"])),
        ("p or q", 2, pc.or_kleene(col["p"], col["q"])),
        ("p and q", 0, pc.and_(col["p"], col["q"])),
    ]
    for sql, mask, want in checks:
        got = oracle_value(rb, sql, mask)
        if pa.types.is_floating(want.type) and "-" in sql and sql.count("-") == 1 and not sql.startswith("-"):
            # subtraction with NaN operands: arrow-rs on x86 and Arrow C++ agree on which NaN propagates; compare values only
            g = O.array_to_arrow(got)
            assert g.null_count == want.null_count
            np.testing.assert_array_equal(g.to_numpy(zero_copy_only=False), want.to_numpy(zero_copy_only=False))
            continue
        same(got, want)


def test_oracle_checked_sub_and_neg_overflow():
    rb = pa.RecordBatch.from_arrays([pa.array([1, -2**31, 5], type=pa.int32()), pa.array([None, -2**63, 1], type=pa.int64())], names=["a", "k"])
    for sql in ("-a", "a - 1", "-k", "0 - a"):
        with pytest.raises(O.OracleError) as e:
            oracle_value(rb, sql, 1)
        assert e.value.kind == "ArithmeticOverflow", sql
    # a null slot is not checked
    rb2 = pa.RecordBatch.from_arrays([pa.array([None, 3], type=pa.int32())], names=["a"])
    assert O.array_to_arrow(oracle_value(rb2, "-a", 1)).to_pylist() == [None, -3]


@pytest.mark.parametrize("sql,mask,kind", X.ERROR_CASES)
def test_errors_are_the_references_without_the_mask(sql, mask, kind):
    rb = X.table(16, seed=2)
    with pytest.raises(O.OracleError) as e:
        oracle_value(rb, sql, mask)
    assert e.value.kind == kind
    with C.sql_extensions(mask):
        with pytest.raises(C.ChdbError) as e2:
            C.Program.compile_filter(sp.parse_expr(sql), rb.schema)
    assert e2.value.kind == kind, (sql, e2.value.kind)


def test_lowering_emits_the_extension_opcodes_only_under_the_mask():
    rb = X.table(16, seed=3)
    assert C.load_library().chdb_get_sql_extensions() == 0
    with C.sql_extensions(C.EXT_OPERATORS | C.EXT_KLEENE):
        assert C.load_library().chdb_get_sql_extensions() == 3
        text = C.Program.compile_filter(sp.parse_expr("(not p) or (s is null and -f - 1.0 > d) or a - 1 > 0"), rb.schema).disassemble()
    assert C.load_library().chdb_get_sql_extensions() == 0
    for word in ("not", "isnull", "neg", "sub", "or_kleene", "and_kleene"):
        assert word in text, (word, text)
    plain = C.Program.compile_filter(sp.parse_expr("p or (q and a > 0)"), rb.schema).disassemble()
    assert "kleene" not in plain
    # constants fold on the host
    with C.sql_extensions(C.EXT_OPERATORS):
        text = C.Program.compile_filter(sp.parse_expr("a > -5 - 2 and 3 is not null and not false"), rb.schema).disassemble()
    assert "neg" not in text and "isnull" not in text and "sub" not in text, text


def test_specialised_kernel_compiles_with_extension_opcodes():
    ok, why = C.jit_available()
    if not ok:
        pytest.skip(why)
    rb = X.table(16, seed=3)
    with C.sql_extensions(3):
        prog = C.Program.compile_filter(sp.parse_expr("(not p) or (s is null and -f - 1.0 > d) or -k - a > 0"), rb.schema)
    assert prog.jit_check()[0] > 0
