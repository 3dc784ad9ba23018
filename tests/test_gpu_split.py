"""The two-launch form (select kernel + gather kernel, CHDB_SPLIT=always) against the CPU oracle, with the
bytecode interpreter kernels (CHDB_JIT=0) and the NVRTC-specialised ones (CHDB_JIT=always).  Large batches take
this path by default (runtime.cu kSplitAutoRows); forcing it here runs the small known-answer cases and the
randomised differential cases through it as well."""
import os

import numpy as np
import pyarrow as pa
import pytest

import chapterhouseqe_b200 as C
import harness as H
import kats
from chapterhouseqe_b200 import sqlparser_lite as sp
from oracle import compute_value as O
from test_gpu_parity import GPU, PREDICATES, PROJECTIONS, _rand_strings, make_mixed_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["0", "always"], ids=["interp", "jit"])
def split_always(request):
    if request.param == "always":
        ok, why = C.api.jit_available()
        if not ok:
            pytest.skip(f"NVRTC unavailable: {why}")
    old = {k: os.environ.get(k) for k in ("CHDB_SPLIT", "CHDB_JIT")}
    os.environ["CHDB_SPLIT"] = "always"
    os.environ["CHDB_JIT"] = request.param
    yield
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def test_split_runs_three_launches():
    ctx = C.default_context()
    rb = H.make_batch([("id", "int32", False)], [[1, 2, 3, 4]])
    before = ctx.launch_count
    out = C.filter_record(rb, [[]], sp.parse_expr("id % 2 = 0"))
    assert out.column(0).to_pylist() == [2, 4]
    assert ctx.launch_count == before + 3   # workspace zeroing + select + gather


FILTER_KATS = [c for c in kats.KATS if c["kind"] != "value"]


@pytest.mark.parametrize("case", FILTER_KATS, ids=[c["name"] for c in FILTER_KATS])
def test_split_kats(case):
    H.check_case(GPU, case)


@pytest.mark.parametrize("n", [1, 2049, 70000])
@pytest.mark.parametrize("pi", range(len(PREDICATES)))
def test_split_filter_matches_oracle(n, pi):
    rb = make_mixed_batch(n, seed=5000 + n)
    al = [[] for _ in rb.schema]
    expr = sp.parse_expr(PREDICATES[pi])
    want = O.filter_record(O.batch_from_arrow(rb), al, expr)
    got = O.batch_from_arrow(C.filter_record(rb, al, expr))
    ok, why = O.batches_equal(got, want)
    assert ok, f"n={n} {PREDICATES[pi]!r}: {why}"


@pytest.mark.parametrize("n", [2049, 30011])
@pytest.mark.parametrize("qi", range(len(PROJECTIONS)))
@pytest.mark.parametrize("pred", ["(id % 2 = 0 and value2 > 10.0) or d < 0.5", "value1 < 'c'", "id < 0"])
def test_split_filter_project_matches_oracle(n, qi, pred):
    rb = make_mixed_batch(n, seed=6000 + n)
    al = [[] for _ in rb.schema]
    sel = sp.parse_select(PROJECTIONS[qi] + " where " + pred)
    b = O.batch_from_arrow(rb)
    want = O.project_record(sel["projection"], O.filter_record(b, al, sel["selection"]), al)
    got = O.batch_from_arrow(C.filter_project_record(sel["selection"], sel["projection"], rb, al))
    ok, why = O.batches_equal(got, want)
    assert ok, f"n={n} q{qi} {pred!r}: {why}"


@pytest.mark.parametrize("L,n", [(100, 5000), (8, 50000), (37, 9999), (1000, 3000)])
@pytest.mark.parametrize("pred", ["id > 25", "id % 7 = 1"])
def test_split_wide_strings(L, n, pred):
    rng = np.random.default_rng(L * 11 + n)
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


def test_split_many_tile_groups():
    """More than kGroupTiles * 32 tiles: the gather kernel's prefix sums loop over the group totals."""
    n = 1024 * 64 * 33 + 777
    rng = np.random.default_rng(9)
    rb = pa.RecordBatch.from_arrays(
        [pa.array(np.arange(n, dtype=np.int32)), pa.array(rng.uniform(0, 100, n).astype(np.float32), mask=rng.random(n) < 0.1)],
        schema=pa.schema([pa.field("id", pa.int32(), False), pa.field("v", pa.float32(), True)]))
    al = [[], []]
    expr = sp.parse_expr("id % 3 = 0 or v > 50.0")
    want = O.filter_record(O.batch_from_arrow(rb), al, expr)
    got = O.batch_from_arrow(C.filter_record(rb, al, expr))
    ok, why = O.batches_equal(got, want)
    assert ok, why


def test_select_overlaps_previous_gather_and_stays_exact():
    """Device-resident batches of different contents run back to back: from the second launch set on, the select
    kernel runs on the ctx's second stream next to the previous set's gather kernel (runtime.cu launch_set).  Outputs
    are released in an irregular pattern so that workspace blocks come back to the cache at every distance; every
    result is compared with the oracle."""
    ctx = C.Context(0)
    al = None
    expr = sp.parse_expr("(id % 3 = 0 and value2 > 10.0) or d < 0.5")
    sizes = [70000, 70000, 131072, 70000, 99999, 70000, 70000, 131072, 70000, 70000, 99999, 70000]
    rbs = [make_mixed_batch(n, seed=9100 + i) for i, n in enumerate(sizes)]
    al = [[] for _ in rbs[0].schema]
    prog = C.Program.compile_filter(expr, rbs[0].schema)
    wants = [O.filter_record(O.batch_from_arrow(rb), al, expr) for rb in rbs]
    devs = [C.DeviceBatch.upload(rb, ctx) for rb in rbs]
    before = ctx.overlapped_launch_sets
    for rounds in range(3):
        outs = []
        for i, d in enumerate(devs):
            outs.append(d.run(prog))
            if i % 3 == 2:      # drop some outputs right away, keep others until the end of the round
                outs[i - 1] = None
        for i, o in enumerate(outs):
            if o is None:
                continue
            ok, why = O.batches_equal(O.batch_from_arrow(o.download()), wants[i])
            assert ok, f"round {rounds} batch {i}: {why}"
        outs = None
    if os.environ.get("CHDB_OVERLAP") != "0":
        assert ctx.overlapped_launch_sets > before, "no launch set used the second stream"
    # a batch consumed right after it was produced must NOT overlap (its producer is the previous launch set)
    chained = devs[2].run(prog)
    n0 = ctx.overlapped_launch_sets
    again = chained.run(prog)
    assert ctx.overlapped_launch_sets == n0
    ok, why = O.batches_equal(O.batch_from_arrow(again.download()), wants[2])   # (filtering twice = filtering once)
    assert ok, why
