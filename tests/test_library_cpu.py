"""CPU-only checks of the C-ABI library: it loads, exports every symbol the header declares,
lowers expressions without a GPU, reports the reference's error kinds at compile time, and
refuses to run without a CUDA device (no CPU fallback)."""
import os
import re

import pytest

import chapterhouseqe_b200 as C
import harness as H
import kats
from chapterhouseqe_b200 import api
from chapterhouseqe_b200 import sqlparser_lite as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False


def test_library_exports_every_header_symbol():
    import ctypes
    hdr = open(os.path.join(ROOT, "include", "chdb_gpu.h")).read()
    declared = sorted(set(re.findall(r"^(?:const char\*|int32_t|uint32_t|int64_t|size_t|void\*?)\s+(chdb_[a-z0-9_]+)\(", hdr, re.M)))
    assert len(declared) >= 25
    L = ctypes.CDLL(C.lib_path())
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/chdb_gpu.h but not exported"
    assert sorted(api.EXPORTED_SYMBOLS) == declared
    lib = C.load_library()
    assert lib.chdb_compiled_arch() == b"sm_100a"
    assert lib.chdb_code_name(12) == b"ArithmeticOverflow"
    assert lib.chdb_code_name(9) == b"UnsupportedTypeCoersion"


@pytest.mark.skipif(_has_gpu(), reason="checks behaviour WITHOUT a CUDA device")
def test_no_cpu_fallback():
    with pytest.raises(C.ChdbError) as ei:
        C.Context(0)
    assert ei.value.kind == "Cuda"
    rb = H.make_batch([("id", "int32", False)], [[1, 2, 3]])
    with pytest.raises(C.ChdbError) as ei:
        C.filter_record(rb, [[]], sp.parse_expr("id > 1"))
    assert ei.value.kind == "Cuda"


SCHEMA = [("id", "int32", False), ("value1", "utf8", False), ("value2", "float32", False)]


def _schema():
    return H.make_batch(SCHEMA, [[1], ["a"], [1.0]]).schema


def test_lowering_large_simple_predicate():
    # sample_queries/large_simple.sql:4  `where id % 2 = 0`
    p = C.Program.compile_filter(sp.parse_expr("id % 2 = 0"), _schema())
    d = p.disassemble()
    assert "container=u32" in d
    assert "0: load.Int32 col0" in d and "1: rem.Int32 imm=0x2" in d and "2: cmp.eq.Int32 imm=0x0" in d
    assert d.count("(compacted)") == 3 and p.num_instructions == 3


def test_lowering_scalar_folding_and_f32_literals():
    # simple.sql:24 `where id > 25 + 0.0`: Int32 25 -> Float32, folded, id cast to Float32 (RNE)
    p = C.Program.compile_filter(sp.parse_expr("id > 25 + 0.0"), _schema())
    d = p.disassemble()
    assert "load.Float32 col0 (Int32)" in d and "cmp.gt.Float32 imm=0x41c80000" in d
    # README.md:86 `value2 > 10.0`: the literal is Float32 10.0 = 0x41200000
    d = C.Program.compile_filter(sp.parse_expr("value2 > 10.0"), _schema()).disassemble()
    assert "cmp.gt.Float32 imm=0x41200000" in d


def test_lowering_projection_names_and_sharing():
    sel = sp.parse_select("select id, id + 1, value2, value2 * 2.0, id as ident, * from t")
    d = C.Program.compile_project(sel["projection"], _schema()).disassemble()
    for name in ("'id'", "'unnamed_1'", "'value2'", "'unnamed_3'", "'ident'", "'value1'"):
        assert name in d
    assert "(shared)" in d and "(compacted)" not in d   # pure projection never copies pass-through columns


def test_lowering_spills_only_for_two_complex_children():
    d = C.Program.compile_filter(sp.parse_expr("(id + 1) * (id + 2) > 10"), _schema()).disassemble()
    assert "push.Int32 spill0" in d and "mul.Int32 spill0" in d   # integer * commutes: no swap flag
    d = C.Program.compile_filter(sp.parse_expr("id * id + id > 10"), _schema()).disassemble()
    assert "push" not in d


def test_single_row_constraint_is_flagged():
    d = C.Program.compile_filter(sp.parse_expr("id > 1 and true"), _schema()).disassemble()
    assert "requires_single_row" in d


_COMPILE_KINDS = {"FailedToParseAsAnInteger", "BinaryOperatorNotImplemented", "ExpressionTypeNotImplemented",
                  "ValueTypeNotImplemented", "ColumnNotFound", "IdentifierNotFound", "UnsupportedTypeCoersion",
                  "CastToBooleanArrayFailedForArrayType", "NotImplemented"}


@pytest.mark.parametrize("case", kats.KATS, ids=[c["name"] for c in kats.KATS])
def test_compile_time_error_kinds(case):
    status, want = case["expect"]
    rb = H.make_batch(case["schema"], case["cols"])
    al = case.get("aliases") or [[] for _ in case["schema"]]
    try:
        if case["kind"] == "value":
            C.Program.compile_project([{"UnnamedExpr": sp.parse_expr(case["sql"])}], rb.schema, al)
        elif case["kind"] == "filter":
            C.Program.compile_filter(sp.parse_expr(case["sql"]), rb.schema, al)
        else:
            s = sp.parse_select(case["sql"])
            if case["kind"] == "project":
                C.Program.compile_project(s["projection"], rb.schema, al)
            else:
                C.Program.compile_filter_project(s["selection"], s["projection"], rb.schema, al)
        got = "ok"
    except C.ChdbError as e:
        got = e.kind
    if status == "error" and want in _COMPILE_KINDS:
        assert got == want
    else:
        assert got == "ok", f"{case['name']}: lowering failed with {got}"


def test_bad_json_and_bad_arguments():
    with pytest.raises(C.ChdbError) as ei:
        C.Program.compile_filter("{not json", _schema())
    assert ei.value.kind == "BadJson"
    with pytest.raises(C.ChdbError) as ei:
        C.Program.compile_filter({"BinaryOp": {"left": {"Value": {"Number": ["1", False]}}}}, _schema())
    assert ei.value.kind == "BadJson"
    with pytest.raises(C.ChdbError) as ei:   # table_aliases shorter than the column list -> the reference's .expect()
        C.Program.compile_filter(sp.parse_expr("t.value2 > 1.0"), _schema(), [["t"]])
    assert ei.value.kind == "Panic"


def test_reference_serde_shapes_are_accepted():
    # exactly what serde_json emits for sqlparser 0.52 (Cargo.toml:29): Ident{value,quote_style},
    # Value::Number(String,bool), unit variants as strings
    expr = ('{"BinaryOp":{"left":{"Nested":{"BinaryOp":{"left":{"CompoundIdentifier":[{"value":"t","quote_style":null},'
            '{"value":"id","quote_style":null}]},"op":"Modulo","right":{"Value":{"Number":["2",false]}}}}},"op":"Eq",'
            '"right":{"Value":{"Number":["0",false]}}}}')
    items = ('[{"Wildcard":{"opt_ilike":null,"opt_exclude":null,"opt_except":null,"opt_replace":null,"opt_rename":null}},'
             '{"ExprWithAlias":{"expr":{"Identifier":{"value":"id","quote_style":"\\""}},"alias":{"value":"x","quote_style":null}}}]')
    p = C.Program.compile_filter_project(expr, items, _schema(), [["t"], ["t"], ["t"]])
    d = p.disassemble()
    assert "rem.Int32" in d and "out 3 'x' Int32" in d


def test_sqlparser_lite_precedence_matches_sqlparser():
    e = sp.parse_expr("a+1.0/(2.0+c)*b")     # test_compute_value.rs:127-175
    assert e["BinaryOp"]["op"] == "Plus"
    mul = e["BinaryOp"]["right"]["BinaryOp"]
    assert mul["op"] == "Multiply" and mul["left"]["BinaryOp"]["op"] == "Divide"
    assert "Nested" in mul["left"]["BinaryOp"]["right"]
    e = sp.parse_expr("a = 1 or b = 2 and c = 3")
    assert e["BinaryOp"]["op"] == "Or" and e["BinaryOp"]["right"]["BinaryOp"]["op"] == "And"
    e = sp.parse_expr("id > -1")
    assert "UnaryOp" in e["BinaryOp"]["right"]
    s = sp.parse_select("select *, t.*, x as y, z w from read_files('a/*.parquet') t where id < 25;")
    assert [next(iter(i)) for i in s["projection"]] == ["Wildcard", "QualifiedWildcard", "ExprWithAlias", "ExprWithAlias"]
    assert s["alias"] == "t" and s["from"] == "read_files('a/*.parquet')"


def test_specialised_kernel_builds_offline():
    """The NVRTC specialisation of a program compiles to an sm_100a cubin without a GPU
    (same device source as the interpreter kernel, bytecode baked in as constants)."""
    ok, why = api.jit_available()
    if not ok:
        pytest.skip(f"NVRTC unavailable: {why}")
    p = C.Program.compile_filter(sp.parse_expr("id % 2 = 0"), _schema())
    src = p.jit_source()
    assert "#define CHDB_JIT 1" in src and "kInstrs[]" in src and "kOutMeta[]" in src
    nbytes, log = p.jit_check()
    assert nbytes > 10_000, log
