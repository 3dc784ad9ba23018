"""CUDA tests of the entry points added around the hot path: many-batch launches for the reference's native
10 000-row records (physical_planner.rs:323), the non-blocking host calls (filter_task.rs:86-125 runs inside a
tokio task), the benchmarked C2 configuration at its full batch size, and the NVLink peer copy.  Everything goes
through the C ABI and is compared with the CPU oracle bit for bit."""
import os
import sys
import time

import numpy as np
import pyarrow as pa
import pytest

import chapterhouseqe_b200 as C
from chapterhouseqe_b200 import sqlparser_lite as sp
from oracle import compute_value as O
from test_gpu_parity import PROJECTIONS, make_mixed_batch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

pytestmark = pytest.mark.gpu

C2_PRED = "(id % 2 = 0 and value2 > 10.0) or d < 0.5"


def _oracle(rb, sel):
    al = [[] for _ in rb.schema]
    out = O.filter_record(O.batch_from_arrow(rb), al, sel["selection"])
    if not (len(sel["projection"]) == 1 and "Wildcard" in sel["projection"][0]):
        out = O.project_record(sel["projection"], out, al)
    return out


@pytest.mark.parametrize("jit", ["0", "always"])
def test_run_many_1000_mixed_size_records(jit, monkeypatch):
    """1 000 records of 1 .. 10 000 rows (the reference's batches are *at most* 10 000 rows; the last batch of a
    file is ragged) through chdb_run_device_many in launches of 256: one output per record, each against the oracle."""
    monkeypatch.setenv("CHDB_JIT", jit)
    ctx = C.default_context()
    rng = np.random.default_rng(77)
    sizes = [10_000] * 600 + [int(x) for x in rng.integers(1, 10_000, 396)] + [1, 1023, 1024, 1025]
    rng.shuffle(sizes)
    big = make_mixed_batch(sum(sizes), seed=4242).select(["id", "k", "value2", "d", "value1"])
    sel = sp.parse_select("select * from t where " + C2_PRED)
    prog = C.Program.compile_filter(sel["selection"], big.schema)
    recs, at = [], 0
    for n in sizes:
        recs.append(big.slice(at, n))
        at += n
    dev = [C.DeviceBatch.upload(r, ctx) for r in recs]   # (sliced views of one big batch: offsets are applied on upload)
    before = ctx.launch_count
    outs = []
    for i in range(0, len(dev), 256):
        outs.extend(C.DeviceBatch.run_many(prog, dev[i:i + 256]))
    groups = (len(dev) + 255) // 256
    # zero + stream kernel per 256 records (+ a launch set of its own for a record whose validity layout differs, e.g. a
    # 1-row record without nulls)
    assert 2 * groups <= ctx.launch_count - before <= 2 * groups + 2 * 16
    assert len(outs) == len(recs)
    for i, (o, r) in enumerate(zip(outs, recs)):
        ok, why = O.batches_equal(O.batch_from_arrow(o.download()), _oracle(r, sel))
        assert ok, f"record {i} ({r.num_rows} rows): {why}"


def test_run_many_fused_projection_and_per_record_errors():
    """Fused filter + project over many records; a checked-integer overflow in one record fails that record only."""
    ctx = C.default_context()
    sel = sp.parse_select("select id, id * id as sq, value1 from t where id % 3 = 0")
    recs = []
    for i in range(40):
        n = 3000 + 17 * i
        base = 0 if i != 7 else 60_000   # record 7: ids above 46 340 -> id * id overflows Int32 for surviving rows
        rb = make_mixed_batch(n, seed=i).select(["id", "value1", "value2"])
        ids = pa.array(np.arange(base, base + n, dtype=np.int32))
        recs.append(pa.RecordBatch.from_arrays([ids, rb.column(1), rb.column(2)], schema=rb.schema))
    prog = C.Program.compile_filter_project(sel["selection"], sel["projection"], recs[0].schema)
    outs = C.DeviceBatch.run_many(prog, [C.DeviceBatch.upload(r, ctx) for r in recs])
    for i, (o, r) in enumerate(zip(outs, recs)):
        if i == 7:
            with pytest.raises(C.ChdbError) as ei:
                o.check()
            assert ei.value.kind == "ArithmeticOverflow"
            continue
        o.check()
        ok, why = O.batches_equal(O.batch_from_arrow(o.download()), _oracle(r, sel))
        assert ok, f"record {i}: {why}"


def test_run_many_mixed_validity_layouts_and_empty_records():
    """Records whose columns do / do not carry validity bitmaps, and empty records, in one call."""
    ctx = C.default_context()
    sel = sp.parse_select("select * from t where " + C2_PRED)
    recs = []
    for i in range(12):
        n = [5000, 0, 777, 10_000][i % 4]
        rb = make_mixed_batch(n, seed=100 + i, null_frac=0.0 if i % 3 == 0 else 0.2).select(["id", "k", "value2", "d", "value1"])
        recs.append(rb)
    prog = C.Program.compile_filter(sel["selection"], recs[0].schema)
    outs = C.DeviceBatch.run_many(prog, [C.DeviceBatch.upload(r, ctx) for r in recs])
    for i, (o, r) in enumerate(zip(outs, recs)):
        ok, why = O.batches_equal(O.batch_from_arrow(o.download()), _oracle(r, sel))
        assert ok, f"record {i}: {why}"


def test_async_host_call_polls_to_completion():
    """chdb_filter_record_async returns before the GPU is done; chdb_poll never blocks; the result equals the
    synchronous call's."""
    ctx = C.default_context()
    rb = make_mixed_batch(1 << 20, seed=5).select(["id", "k", "value2", "d", "value1"])
    sel = sp.parse_select("select * from t where " + C2_PRED)
    prog = C.Program.compile_filter(sel["selection"], rb.schema)
    want = prog.run(rb, ctx)
    pend = [prog.run_async(rb, ctx) for _ in range(3)]
    polls = 0
    t0 = time.time()
    while not all(p.poll() for p in pend):
        polls += 1
        assert time.time() - t0 < 60
        time.sleep(0.0005)
    for p in pend:
        assert p.poll()
        got = p.result()
        assert got.equals(want)
    ok, why = O.batches_equal(O.batch_from_arrow(want), _oracle(rb, sel))
    assert ok, why
    # an error of the data surfaces through poll / result
    bad = pa.RecordBatch.from_arrays([pa.array(np.arange(100000, 100100, dtype=np.int32))], names=["id"])
    pp = C.Program.compile_project(sp.parse_select("select id * id as sq from t")["projection"], bad.schema)
    p = pp.run_async(bad, ctx)
    with pytest.raises(C.ChdbError) as ei:
        t0 = time.time()
        while not p.poll():
            assert time.time() - t0 < 60
        p.result()
    assert ei.value.kind == "ArithmeticOverflow"


def test_device_batch_ready_flag():
    ctx = C.default_context()
    rb = make_mixed_batch(1 << 21, seed=6).select(["id", "k", "value2", "d", "value1"])
    prog = C.Program.compile_filter(sp.parse_expr(C2_PRED), rb.schema)
    dev = C.DeviceBatch.upload(rb, ctx)
    out = dev.run(prog)
    t0 = time.time()
    while not out.ready:
        assert time.time() - t0 < 60
    assert out.num_rows > 0


def test_benchmarked_c2_batch_against_oracle():
    """The exact configuration bench.py times (C2 schema, predicate, 10 % / 5 % nulls, 2^22 rows, JIT auto),
    generated by bench.py's own generator, against the oracle."""
    import torch

    import bench
    cfg = bench.CONFIGS["C2"]
    dev = torch.device("cuda:0")
    t = bench.gen_batch_torch(cfg, 3 << 22, 1 << 22, cfg["seed"] + 3, dev)
    ctx = C.default_context()
    sel = sp.parse_select(cfg["sql"])
    prog = C.Program.compile_filter(sel["selection"], bench.schema(cfg))
    before = ctx.jit_launch_count
    out = bench.wrap_device_batch(C, cfg, t, ctx).run(prog)
    got = O.batch_from_arrow(out.download())
    if C.jit_available()[0]:
        assert ctx.jit_launch_count >= before + 1, "the benchmarked batch size must run the specialised kernel"
    rb, _keep = bench.to_host_batch(cfg, t, pin=False)
    ok, why = O.batches_equal(got, _oracle(rb, sel))
    assert ok, why
    assert 0.70 < got.num_rows / (1 << 22) < 0.72


@pytest.mark.parametrize("config", ["C3", "C4", "C4H", "C5"])
def test_other_bench_configs_against_oracle(config):
    """One batch of every other configuration bench.py can time, against the oracle."""
    import torch

    import bench
    cfg = bench.CONFIGS[config]
    n = 1 << 20
    t = bench.gen_batch_torch(cfg, 12345, n, cfg["seed"], torch.device("cuda:0"))
    ctx = C.default_context()
    sel = sp.parse_select(cfg["sql"])
    star = len(sel["projection"]) == 1 and "Wildcard" in sel["projection"][0]
    sch = bench.schema(cfg)
    prog = (C.Program.compile_filter(sel["selection"], sch) if star
            else C.Program.compile_filter_project(sel["selection"], sel["projection"], sch))
    got = O.batch_from_arrow(bench.wrap_device_batch(C, cfg, t, ctx).run(prog).download())
    rb, _keep = bench.to_host_batch(cfg, t, pin=False)
    ok, why = O.batches_equal(got, _oracle(rb, sel))
    assert ok, why


def test_peer_copy_between_two_gpus():
    """chdb_peer_copy: a filtered batch on GPU 0 copied to GPU 1 over NVLink equals the oracle's result; the source
    is released right after the call (the copy must not read recycled blocks)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctx0, ctx1 = C.Context(0), C.Context(1)
    sel = sp.parse_select("select * from t where " + C2_PRED)
    rb = make_mixed_batch(1 << 20, seed=9)
    prog = C.Program.compile_filter(sel["selection"], rb.schema)
    want = _oracle(rb, sel)
    copies = []
    for _ in range(4):
        out = C.DeviceBatch.upload(rb, ctx0).run(prog)
        copies.append(out.peer_copy(ctx1))
        out.close()                         # blocks go back to ctx0's cache ...
        C.DeviceBatch.upload(rb, ctx0).run(prog).close()   # ... and the next launch on ctx0 reuses them
    for c in copies:
        assert c.ctx.device == 1
        ok, why = O.batches_equal(O.batch_from_arrow(c.download()), want)
        assert ok, why
    # one compiled program serves both devices, specialised kernels included
    out1 = C.DeviceBatch.upload(rb, ctx1).run(prog)
    ok, why = O.batches_equal(O.batch_from_arrow(out1.download()), want)
    assert ok, why


def test_record_pool_budget_spill_and_consumers():
    """The device RecordPool of a GPU-aware exchange: records held by reference, a record a consumer holds is never
    spilled, least-recently-used records beyond the byte budget go to pinned host memory and come back on get(),
    a record is dropped after its last consumer operator completes it (exchange_operator.rs:727-733)."""
    ctx = C.default_context()
    sel = sp.parse_select("select * from t where " + C2_PRED)
    recs = [make_mixed_batch(50_000, seed=300 + i).select(["id", "k", "value2", "d", "value1"]) for i in range(6)]
    prog = C.Program.compile_filter(sel["selection"], recs[0].schema)
    one = C.DeviceBatch.upload(recs[0], ctx)
    pool = C.RecordPool(ctx, budget_bytes=0)
    pool.add(0, one, consumers=1)
    per_record = pool.stats()["device_bytes"]
    assert per_record > 50_000 * 30
    pool.complete(0)
    assert pool.stats() == {"records": 0, "device_bytes": 0, "spilled_records": 0, "spilled_bytes": 0}
    pool.close()
    one.close()

    pool = C.RecordPool(ctx, budget_bytes=int(per_record * 2.5))    # room for two records
    for i, r in enumerate(recs):
        d = C.DeviceBatch.upload(r, ctx)
        pool.add(i, d, consumers=2 if i == 1 else 1)
        d.close()                                                    # the pool's reference keeps it alive
    s = pool.stats()
    assert s["records"] == 6 and s["spilled_records"] == 4 and s["device_bytes"] <= per_record * 2.5, s
    held = pool.get(5)                                               # resident, and now referenced by a consumer
    for i in range(6):                                               # walks the whole pool: spilled ones are uploaded again
        b = pool.get(i)
        ok, why = O.batches_equal(O.batch_from_arrow(b.run(prog).download()), _oracle(recs[i], sel))
        assert ok, f"record {i}: {why}"
        b.close()
    assert pool.stats()["device_bytes"] <= per_record * 3.5          # record 5 is pinned by `held`, never spilled
    ok, why = O.batches_equal(O.batch_from_arrow(held.download()), O.batch_from_arrow(recs[5]))
    assert ok, why
    held.close()
    pool.complete(1)
    assert pool.stats()["records"] == 6                              # second consumer of record 1 still to come
    for i in range(6):
        pool.complete(i)
    assert pool.stats()["records"] == 0 and pool.stats()["device_bytes"] == 0
    with pytest.raises(C.ChdbError):
        pool.get(3)
    pool.close()
