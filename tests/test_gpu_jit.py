"""The NVRTC-specialised kernels (CHDB_JIT=always) against the CPU oracle: same device source as the
interpreter kernel with the program baked in, so results must be bit-identical."""
import os

import numpy as np
import pytest

import chapterhouseqe_b200 as C
import harness as H
import kats
from chapterhouseqe_b200 import sqlparser_lite as sp
from oracle import compute_value as O
from test_gpu_parity import PROJECTIONS, make_mixed_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def jit_always():
    ok, why = C.api.jit_available()
    if not ok:
        pytest.skip(f"NVRTC unavailable: {why}")
    old = os.environ.get("CHDB_JIT")
    os.environ["CHDB_JIT"] = "always"
    yield
    if old is None:
        os.environ.pop("CHDB_JIT", None)
    else:
        os.environ["CHDB_JIT"] = old


def _check_filter(rb, pred):
    al = [[] for _ in rb.schema]
    expr = sp.parse_expr(pred)
    ctx = C.default_context()
    before = ctx.jit_launch_count
    got = O.batch_from_arrow(C.filter_record(rb, al, expr))
    assert ctx.jit_launch_count >= before + 1, "the specialised kernel did not run"
    want = O.filter_record(O.batch_from_arrow(rb), al, expr)
    ok, why = O.batches_equal(got, want)
    assert ok, f"{pred!r}: {why}"


@pytest.mark.parametrize("pred", ["(id % 2 = 0 and value2 > 10.0) or d < 0.5", "value1 < 'm' and flag",
                                  "k % 7 = 3 or small / 5 = 2"])
@pytest.mark.parametrize("n", [1, 6151, 40000])
def test_jit_filter_matches_oracle(pred, n):
    _check_filter(make_mixed_batch(n, seed=4000 + n), pred)


def test_jit_wide_strings():
    import pyarrow as pa
    from test_gpu_parity import _rand_strings
    rng = np.random.default_rng(11)
    n = 20000
    rb = pa.RecordBatch.from_arrays(
        [pa.array(np.arange(n, dtype=np.int32)), _rand_strings(rng, n, 100, 100),
         pa.array(rng.uniform(0, 100, n).astype(np.float32))],
        schema=pa.schema([pa.field("id", pa.int32(), False), pa.field("value1", pa.utf8(), False),
                          pa.field("value2", pa.float32(), False)]))
    _check_filter(rb, "id % 2 = 0")


def test_jit_fused_filter_project():
    rb = make_mixed_batch(30011, seed=77)
    al = [[] for _ in rb.schema]
    sel = sp.parse_select(PROJECTIONS[1] + " where id % 3 = 0")
    b = O.batch_from_arrow(rb)
    want = O.project_record(sel["projection"], O.filter_record(b, al, sel["selection"]), al)
    ctx = C.default_context()
    before = ctx.jit_launch_count
    got = O.batch_from_arrow(C.filter_project_record(sel["selection"], sel["projection"], rb, al))
    assert ctx.jit_launch_count >= before + 1
    ok, why = O.batches_equal(got, want)
    assert ok, why


def test_jit_error_kinds():
    by_name = {c["name"]: c for c in kats.KATS}
    for name in ["first_error_node_wins", "first_error_node_wins_2", "fused_errors_only_on_surviving_rows",
                 "div_min_by_minus1", "checked_add_skips_null_slots"]:
        H.check_case(__import__("test_gpu_parity").GPU, by_name[name])
