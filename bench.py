#!/usr/bin/env python
"""bench.py -- filter + project + compaction throughput on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2|C3|C4|C4H|C5|C2S]   our arm (CUDA)
  python bench.py --impl reference [...]                                          the CPU restatement, all host threads

Workloads (BASELINE.json configs).  Every input column is referenced in every config.
  C2  (default, the configuration the metric is quoted on): large_simple.sql-shaped, 100 M rows per GPU,
      id Int32 | k Int64 | value2 Float32 (10 % null) | d Float64 (5 % null) | value1 Utf8[8],
      `select * where (id % 2 = 0 AND value2 > 10.0) OR d < 0.5` (the reference's non-Kleene AND/OR), 2^22-row batches.
  C2S same data in the reference's native 10 000-row batches (physical_planner.rs:323), processed through the
      many-batch entry point (one launch set per 512 records).
  C3  simple.sql query 4: 7-column projection with arithmetic and implicit casts over `where id > 25 + 0.0`
      (fused filter + project), base schema, 16 M rows.
  C4  only_wide_strings_query.sql: base schema with 100-byte strings, `select * where id > 25`, 10 M rows;
  C4H the same data with `where id % 2 = 0`.
  C5  huge_simple.sql scaled: base schema (id Int32 | value1 Utf8[8] | value2 Float32), `select * where id % 2 = 0`,
      500 M rows per GPU (4 B rows at 8 GPUs), sharded by file / row group.
One step = `--passes` passes of the hot path over the GPU's resident rows (default: enough for >= 50 ms per step).
Multi-GPU: one process per GPU (torchrun), records shard by record_id, no collective on the data path ("weak").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "filter_project_rows_per_s"
UNIT = "rows/s"

BASE_COLS = lambda L: [("id", "i32", "seq", 0.0), ("value1", "utf8", L, 0.0), ("value2", "f32", "u100", 0.0)]  # noqa: E731
CONFIGS = {
    "C2": dict(seed=0xC4DB0002, rows=100_000_000, batch_rows=1 << 22, passes=30,
               cols=[("id", "i32", "seq", 0.0), ("k", "i64", "u31", 0.0), ("value2", "f32", "u100", 0.10),
                     ("d", "f64", "normal", 0.05), ("value1", "utf8", 8, 0.0)],
               sql="select * from read_files('large_simple/*.parquet') where (id % 2 = 0 and value2 > 10.0) or d < 0.5",
               title="C2 large_simple-shaped synthetic"),
    "C3": dict(seed=0xC4DB0003, rows=16_000_000, batch_rows=1 << 22, passes=96, cols=BASE_COLS(8), id_mod=46340,
               sql="select id, value1, id + 10.0 as id_plus_10, (value2 + 10) / 100 as value2, 1.0 / id as value3, "
                   "1.0 / (id * id) as value4, id * id as value5 from read_files('simple/*.parquet') where id > 25 + 0.0",
               title="C3 simple.sql q4-shaped projection with arithmetic and casts (ids in [1, 46340])"),
    "C4": dict(seed=0xC4DB0004, rows=10_000_000, batch_rows=1 << 21, passes=24, cols=BASE_COLS(100),
               sql="select * from read_files('simple_wide_string/*.parquet') where id > 25",
               title="C4 only_wide_strings-shaped (100-byte strings), ~100 % selected"),
    "C4H": dict(seed=0xC4DB0004, rows=10_000_000, batch_rows=1 << 21, passes=24, cols=BASE_COLS(100),
                sql="select * from read_files('simple_wide_string/*.parquet') where id % 2 = 0",
                title="C4 only_wide_strings-shaped (100-byte strings), 50 % selected"),
    "C5": dict(seed=0xC4DB0005, rows=500_000_000, batch_rows=1 << 22, passes=6, cols=BASE_COLS(8),
               sql="select * from read_files('huge_simple/*.parquet') where id % 2 = 0",
               title="C5 huge_simple-scaled (4 B rows at 8 GPUs)"),
}
CONFIGS["C2S"] = dict(CONFIGS["C2"], batch_rows=10_000, rows=20_000_000, passes=8, many=512,
                      title="C2 in reference-native 10 000-row records (many-batch launches of 512 records)")
WIDTH = {"i32": 4, "i64": 8, "f32": 4, "f64": 8}

if os.environ.get("CHDB_BENCH_WATCHDOG"):   # debugging aid: dump every thread's stack if the run is still going after N seconds
    import faulthandler
    faulthandler.dump_traceback_later(float(os.environ["CHDB_BENCH_WATCHDOG"]), exit=True)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--config", choices=sorted(CONFIGS), default="C2")
    ap.add_argument("--rows", type=int, default=None, help="rows per GPU (default: the config's)")
    ap.add_argument("--batch-rows", type=int, default=None)
    ap.add_argument("--passes", type=int, default=None, help="passes over the resident rows per step")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-instances", type=int, default=3, help="filter operator instances (ctx + stream each) for e2e")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="budget for each cpu_baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="skip the N>1 materialize-gather measurement")
    ap.add_argument("--gather-steps", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of three benchmarked batches")
    ap.add_argument("--parquet", action="store_true",
                    help="SURVEY 8f row f1 instead of the filter: Parquet -> device decode of the reference's sample schema")
    ap.add_argument("--parquet-plain", action="store_true", help="--parquet: write the file without dictionary encoding")
    ap.add_argument("--parquet-encode", action="store_true",
                    help="f3: device batches -> Parquet file image in pinned host memory (chdb_parquet_encode), records of --batch-rows rows "
                         "coalesced into 1 Mi-row row groups, against pyarrow's writer on one thread")
    ap.add_argument("--parquet-filter", action="store_true",
                    help="--parquet: read_files -> filter (id %% 2 = 0) -> download of the result, the decoded batches never leave HBM")
    a = ap.parse_args()
    cfg = CONFIGS[a.config]
    a.rows = a.rows or cfg["rows"]
    a.batch_rows = a.batch_rows or cfg["batch_rows"]
    a.passes = a.passes or cfg["passes"]
    return a


# ---------------------------------------------------------------------------------------------
# synthetic data (deterministic; generated on the device when there is one)
# ---------------------------------------------------------------------------------------------
def batch_sizes(rows: int, batch_rows: int):
    out, done = [], 0
    while done < rows:
        n = min(batch_rows, rows - done)
        out.append((done, n))
        done += n
    return out


def gen_batch_torch(cfg, start: int, n: int, seed: int, device):
    """One batch as torch tensors per column: {"name": (values, validity | None, offsets | None, nulls)}; every
    buffer is readable 64 bytes past its logical end (TMA bulk copies move whole 16-byte units)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    pad = 64
    n8 = (n + 7) // 8 * 8
    w = torch.tensor([1, 2, 4, 8, 16, 32, 64, 128], dtype=torch.int32, device=device)

    def bitmap(p_null):
        valid = (torch.rand(n8, generator=g, device=device) >= p_null)
        valid[n:] = False
        packed = (valid.view(-1, 8).to(torch.int32) * w).sum(dim=1).to(torch.uint8)
        out = torch.zeros(n8 // 8 + pad, dtype=torch.uint8, device=device)
        out[: n8 // 8] = packed
        return out, n - int(valid.sum().item())

    cols = {}
    for name, typ, dist, p_null in cfg["cols"]:
        offsets = None
        if typ == "utf8":
            L = int(dist)
            vals = torch.zeros(n * L + pad, dtype=torch.uint8, device=device)
            vals[: n * L] = torch.randint(97, 123, (n * L,), generator=g, device=device, dtype=torch.int64).to(torch.uint8)
            offsets = torch.zeros(n + 1 + pad, dtype=torch.int32, device=device)
            offsets[: n + 1] = torch.arange(0, (n + 1) * L, L, dtype=torch.int64, device=device).to(torch.int32)
        else:
            dt = {"i32": torch.int32, "i64": torch.int64, "f32": torch.float32, "f64": torch.float64}[typ]
            vals = torch.zeros(n + pad, dtype=dt, device=device)
            if dist == "seq":
                ids = torch.arange(start, start + n, dtype=torch.int64, device=device)
                if cfg.get("id_mod"):
                    ids = ids % cfg["id_mod"] + 1
                vals[:n] = ids.to(dt)
            elif dist == "u31":
                vals[:n] = torch.randint(-2**31, 2**31, (n,), generator=g, device=device, dtype=torch.int64).to(dt)
            elif dist == "u100":
                vals[:n] = (torch.rand(n, generator=g, device=device) * 100.0).to(dt)
            elif dist == "normal":
                vals[:n] = torch.randn(n, generator=g, device=device, dtype=dt)
            else:
                raise ValueError(dist)
        validity, nulls = (None, 0)
        if p_null > 0:
            validity, nulls = bitmap(p_null)
        cols[name] = (vals, validity, offsets, nulls)
    return dict(n=n, cols=cols)


def schema(cfg):
    import pyarrow as pa
    t = {"i32": pa.int32(), "i64": pa.int64(), "f32": pa.float32(), "f64": pa.float64(), "utf8": pa.utf8()}
    return pa.schema([pa.field(name, t[typ], p_null > 0) for name, typ, _, p_null in cfg["cols"]])


def in_bytes_per_row(cfg) -> float:
    """Algorithmic input bytes per row: value bytes (Utf8: 4-byte offset + string bytes) + 1/8 per
    validity bitmap, every referenced column once (all columns are referenced in every config)."""
    b = 0.0
    for _, typ, dist, p_null in cfg["cols"]:
        b += (4 + int(dist)) if typ == "utf8" else WIDTH[typ]
        b += 0.125 if p_null > 0 else 0.0
    return b


def wrap_device_batch(C, cfg, t, ctx):
    vals, vald, offs = [], [], []
    for name, *_ in cfg["cols"]:
        v, val, off, _ = t["cols"][name]
        vals.append(v.data_ptr())
        vald.append(val.data_ptr() if val is not None else 0)
        offs.append(off.data_ptr() if off is not None else 0)
    return C.DeviceBatch.wrap(schema(cfg), t["n"], vals, vald, offs, ctx=ctx, keepalive=t)


def to_host_batch(cfg, t, pin: bool):
    """torch tensors -> pyarrow RecordBatch over (optionally pinned) host memory, zero-copy."""
    import pyarrow as pa
    import torch
    n = t["n"]
    keep, arrays = [], []

    def host(x, count):
        h = torch.empty(count, dtype=x.dtype, pin_memory=pin)
        h.copy_(x[:count])
        keep.append(h)
        return pa.py_buffer(h.numpy())

    sch = schema(cfg)
    for f, (name, typ, dist, _) in zip(sch, cfg["cols"]):
        v, val, off, nulls = t["cols"][name]
        vb = host(val, (n + 7) // 8) if val is not None else None
        if typ == "utf8":
            arrays.append(pa.Array.from_buffers(f.type, n, [vb, host(off, n + 1), host(v, n * int(dist))], null_count=nulls))
        else:
            arrays.append(pa.Array.from_buffers(f.type, n, [vb, host(v, n)], null_count=nulls))
    return pa.RecordBatch.from_arrays(arrays, schema=sch), keep


def batch_nbytes(rb) -> int:
    return sum(b.size for col in rb.columns for b in col.buffers() if b is not None)


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons of one GPU during the timed region (NVML).

    An NVML query -- from a thread of this process or from `nvidia-smi -lms` next to it -- stalls this process's
    CUDA calls for milliseconds now and then, which starves the GPU in a timed region, so the samples are taken
    by the benchmark thread itself after it has enqueued all timed steps and before it synchronises: the GPU is
    still working through the queue (under load, inside the timed region)."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self.nv = [], set(), None, None
        if os.environ.get("CHDB_BENCH_NO_CLOCKS"):
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)   # first query (slow) outside the timed region
        except Exception:  # noqa: BLE001
            self.nv = None

    def sample_while(self, busy, max_samples: int = 64):
        """Samples until busy() turns false (at least once)."""
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while True:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            if len(self.samples) >= max_samples or not busy():
                break
            time.sleep(0.005)

    def result(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "sampled": "NVML, by the benchmark thread between enqueueing the last timed step and synchronising"}


def measured_traffic(config: str, batch_rows: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the stream kernel from the committed `ncu --set full`
    capture of this workload (profiles/traffic.json: {config: {batch_rows, dram_bytes_per_launch}}); None if never captured."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(config)
        if t and int(t.get("batch_rows", -1)) == int(batch_rows):
            return t["dram_bytes_per_launch"]
    except Exception:  # noqa: BLE001
        pass
    return None


def kernel_description(many, jit: bool, split: bool) -> str:
    how = "device_code.cuh specialised for the program by NVRTC" if jit else "bytecode interpreter"
    if many:
        return (f"chdb_jit_stream_many ({how}): the fused single-pass kernel, one launch per {many} records, preceded by "
                "zero_kernel")
    if split:
        return (f"chdb_jit_select + chdb_jit_gather ({how}): per record one zero_kernel, one select launch (predicate -> "
                "selection bitmap + tile totals) and one gather launch (TMA-staged tiles -> compacted columns); the gather "
                "kernel is the dominant one (about three quarters of the launch set); zero + select of record k run on the "
                "ctx's second stream next to the gather of record k-1 (overlapped_launch_sets)")
    return f"chdb_jit_stream ({how}): the fused single-pass kernel, one launch per record, preceded by zero_kernel"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------
# CPU legs (the oracle is the checker / baseline here, never the product path)
# ---------------------------------------------------------------------------------------------
def oracle_step_fn(sel):
    """The reference's per-record work (filter_record, then project_record when the query has a select list other
    than `*`) on an already converted oracle Batch: the reference receives an Arc<RecordBatch> by pointer
    (exchange_operator.rs:621-667), so the Arrow -> oracle-array conversion stays outside every timed loop."""
    from oracle import compute_value as O
    star = len(sel["projection"]) == 1 and "Wildcard" in sel["projection"][0]

    def run(b):
        al = [[] for _ in b.fields]
        out = O.filter_record(b, al, sel["selection"])
        if not star:
            out = O.project_record(sel["projection"], out, al)
        return out
    return run


def pyarrow_step_fn(config: str):
    """Independent SIMD CPU reference (Arrow C++ through pyarrow.compute).  `%` has no pyarrow kernel:
    x % 2 is computed as x - (x / 2) * 2 (truncating integer division, same result)."""
    import pyarrow.compute as pc

    def mod2_is_0(x):
        return pc.equal(pc.subtract(x, pc.multiply(pc.divide(x, 2), 2)), 0)

    if config in ("C2", "C2S"):
        return lambda rb: rb.filter(pc.or_(pc.and_(mod2_is_0(rb["id"]), pc.greater(rb["value2"], 10.0)),
                                           pc.less(rb["d"], 0.5)))
    if config in ("C4H", "C5"):
        return lambda rb: rb.filter(mod2_is_0(rb["id"]))
    if config == "C4":
        return lambda rb: rb.filter(pc.greater(rb["id"], 25))
    return None


def timed_sample(fn, items, budget_s: float):
    t0 = time.perf_counter()
    rows = used = 0
    for it, n in items:
        fn(it)
        rows += n
        used += 1
        if time.perf_counter() - t0 >= budget_s:
            break
    dt = time.perf_counter() - t0
    return rows / dt, used, rows, dt


def cpu_baselines(args, cfg, sel, host_batches):
    """Single-thread oracle (mirrors the reference's one inline filter instance, filter_task.rs:99) and single-thread
    pyarrow.compute, each on a bounded sample of the same workload."""
    from oracle import compute_value as O
    O.lib()
    fn = oracle_step_fn(sel)
    conv = [(O.batch_from_arrow(rb), rb.num_rows) for rb in host_batches]   # outside the timed region
    fn(conv[0][0])
    v, used, rows, dt = timed_sample(fn, conv, args.cpu_seconds)
    out = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
           "sample": f"{used} record(s) = {rows} rows of the same workload, oracle (arrow_kernels.c + tree walk) on 1 thread, "
                     f"{dt:.2f} s; Arrow->oracle conversion outside the timed loop"}
    pa_fn = pyarrow_step_fn(args.config)
    if pa_fn is not None:
        import pyarrow as pa
        pa.set_cpu_count(1)
        pa_fn(host_batches[0])
        v2, used2, rows2, dt2 = timed_sample(pa_fn, [(rb, rb.num_rows) for rb in host_batches], args.cpu_seconds)
        out["pyarrow_compute"] = {"value": v2, "unit": UNIT, "cores": 1,
                                  "sample": f"{used2} record(s) = {rows2} rows, pyarrow {pa.__version__} compute + RecordBatch.filter, "
                                            f"1 thread, {dt2:.2f} s"}
    return out


def run_reference(args, cfg, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port; the Rust reference cannot be built here) on all
    host threads, one record per thread at a time -- N filter instances pulling from one exchange."""
    if rank != 0:
        return
    import concurrent.futures as cf

    import torch

    from chapterhouseqe_b200 import sqlparser_lite as sp
    from oracle import compute_value as O
    O.lib()
    sel = sp.parse_select(cfg["sql"])
    fn = oracle_step_fn(sel)
    cores = os.cpu_count() or 1
    # bounded sample of the workload: >= 2 records per thread, at most ~64 M rows, generated like our arm's data
    sizes = batch_sizes(args.rows, args.batch_rows)
    n_batches = min(len(sizes), max(8, 2 * cores, min(4096, (32_000_000 // args.batch_rows))))
    dev = torch.device("cuda:0") if torch.cuda.is_available() else torch.device("cpu")
    conv = []
    for i, (start, n) in enumerate(sizes[:n_batches]):
        t = gen_batch_torch(cfg, start, n, cfg["seed"] + i, dev)
        rb, keep = to_host_batch(cfg, t, pin=False)
        conv.append(O.batch_from_arrow(rb))   # the reference holds RecordBatches already: converted once, outside the timed loop
        del t, rb, keep
    sample_rows = sum(b.num_rows for b in conv)
    pool = cf.ThreadPoolExecutor(max_workers=cores)

    def step():
        list(pool.map(fn, conv))

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = sample_rows * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "i32/f32/f64", "data": "synthetic",
        "config": workload_config(args, cfg),
        "sample_rows_per_step": sample_rows,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n_batches} records = {sample_rows} rows per step, one record per thread on {cores} threads, "
                                   "inputs already in memory as arrays (no conversion in the timed loop); oracle port of the "
                                   "arrow-rs path -- the Rust reference cannot be built in this image"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, cfg):
    cols = " | ".join(f"{n} {t}" + (f"[{d}]" if t == "utf8" else "") + (f" ({int(p * 100)}% null)" if p else "")
                      for n, t, d, p in cfg["cols"])
    return {"workload": f"{cfg['title']}: {cfg['sql']}", "name": args.config, "rows_per_gpu": args.rows,
            "batch_rows": args.batch_rows, "passes_per_step": args.passes, "schema": cols,
            "null_semantics": "non-Kleene (arrow compute::and/or)", "parallelism": f"shard-by-record x{args.gpus}, no collective",
            "l2": "resident inputs per GPU exceed the 126 MB L2 and every pass streams all of them; no flush needed"}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args, cfg, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import chapterhouseqe_b200 as C
    from chapterhouseqe_b200 import multigpu
    from chapterhouseqe_b200 import sqlparser_lite as sp

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; chapterhouseqe_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    C.load_library()
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    ctx = C.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=device)
    sel = sp.parse_select(cfg["sql"])
    star = len(sel["projection"]) == 1 and "Wildcard" in sel["projection"][0]
    sch = schema(cfg)
    prog = (C.Program.compile_filter(sel["selection"], sch) if star
            else C.Program.compile_filter_project(sel["selection"], sel["projection"], sch))
    many = int(cfg.get("many", 0))

    # ---- resident inputs ----
    sizes = batch_sizes(args.rows, args.batch_rows)
    base_row = rank * args.rows
    tensors, dev_batches = [], []
    for i, (start, n) in enumerate(sizes):
        t = gen_batch_torch(cfg, base_row + start, n, cfg["seed"] + rank * 100003 + i, device)
        tensors.append(t)
        dev_batches.append(wrap_device_batch(C, cfg, t, ctx))
    torch.cuda.synchronize(device)
    total_rows = sum(n for _, n in sizes)

    def barrier():
        if world > 1:
            dist.barrier()

    if many:
        # handle arrays built once: a launch set is one C call over a slice of pointers (what a Rust caller passes)
        groups = [C.DeviceBatchList.from_batches(dev_batches[i:i + many]) for i in range(0, len(dev_batches), many)]

        def one_pass():
            return [C.DeviceBatch.run_many(prog, g) for g in groups]
    else:
        def one_pass():
            return [b.run(prog) for b in dev_batches]

    def run_passes(n):
        """n passes back to back; the previous pass's outputs are released while the next one is enqueued (their blocks
        go back to the ctx's block cache), exactly like a consumer that drains the outbound exchange."""
        prev = None
        for i in range(n):
            cur = one_pass()
            prev = cur
            if os.environ.get("CHDB_BENCH_TRACE"):
                sys.stderr.write(f"[bench] pass {i}: block cache misses so far {ctx.alloc_misses}\n")
        return prev

    sampler = ClockSampler(local_rank)   # NVML initialised (and queried once) before the warm-up
    # warm-up: the same enqueue pattern as the timed region, so the block cache holds every buffer the steady state
    # needs (a cache miss is a cudaMallocAsync that may take the driver's slow path: tens of milliseconds)
    outs = run_passes(max(args.warmup, 3))
    ctx.synchronize()
    torch.cuda.synchronize(device)
    def flat(results):
        return [o for r in results for o in r] if many else results

    rows_out = bytes_out = 0
    for o in flat(outs):
        o.check()
        rows_out += o.num_rows
        bytes_out += o.nbytes
    o = None   # (the loop variable would keep the last batch -- and its blocks -- out of the cache)

    # ---- the benchmarked configuration against the oracle: first, middle and last record of this rank ----
    parity = None
    if not args.no_parity and rank == 0:
        from oracle import compute_value as O
        O.lib()
        fn = oracle_step_fn(sel)
        checked = []
        flat_outs = flat(outs)
        for i in sorted({0, len(dev_batches) // 2, len(dev_batches) - 1}):
            got = O.batch_from_arrow(flat_outs[i].download())
            want = fn(O.batch_from_arrow(to_host_batch(cfg, tensors[i], pin=False)[0]))
            ok, why = O.batches_equal(got, want)
            if not ok:
                raise SystemExit(f"bench.py: record {i} of the benchmarked run differs from the oracle: {why}")
            checked.append(i)
        flat_outs = None
        parity = {"parity_checked_batches": len(checked), "records": checked, "rows_each": [sizes[i][1] for i in checked],
                  "against": "oracle (bit-exact: values, validity, offsets, schema)"}
    outs = None

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0, jit0, overlapped0 = ctx.launch_count, ctx.jit_launch_count, ctx.overlapped_launch_sets
    barrier()
    torch.cuda.synchronize(device)
    # a cyclic-GC pause in the middle of the enqueue loop (tens of ms with torch / pyarrow loaded) starves the GPU,
    # which is only a few ms behind the host: collect now, not during the timed region
    import gc
    gc.collect()
    gc.disable()
    misses0 = ctx.alloc_misses
    ev0.record(stream)
    t_host0 = time.perf_counter()
    prev = run_passes(args.steps * args.passes)
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3
    ev1.record(stream)
    sampler.sample_while(lambda: not ev1.query())   # the GPU is still inside the timed region, the host is done enqueueing
    ctx.synchronize()
    torch.cuda.synchronize(device)
    barrier()
    gc.enable()
    clocks = sampler.result()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count - launches0
    jit_launches = ctx.jit_launch_count - jit0
    rows_timed = 0
    for o in flat(prev):   # the LAST timed pass's outputs: device error word + row counts must match the warm-up's
        o.check()
        rows_timed += o.num_rows
    o = None
    assert rows_timed == rows_out, "timed pass produced a different row count than the warm-up"
    prev = None

    # whole-job numbers: MAX over ranks of the device time, SUM of rows / launches (no data-path collective)
    elapsed_ms, (all_rows, all_rows_out, all_launches, jit_launches) = multigpu.reduce_step(
        elapsed_ms, [total_rows, rows_out, launches, jit_launches], device)

    selectivity = rows_out / total_rows
    # algorithmic bytes: every referenced input byte once + every output byte once (the output bytes are
    # what the result batches hold: values + offsets + validity of the surviving rows)
    bytes_per_row = in_bytes_per_row(cfg) + bytes_out / total_rows
    n_passes = args.steps * args.passes
    records = len(sizes)
    secs = elapsed_ms / 1e3
    value = all_rows * n_passes / secs
    per_gpu_gbs = total_rows * n_passes * bytes_per_row / secs / 1e9
    peak, peak_src = measured_peak()
    stream_launches_per_pass = len(groups) if many else records
    # (zero + select + gather = 3 launches per record in the two-launch form, zero + stream = 2 in the fused form)
    split_launches = (not many) and launches >= 3 * records * (args.steps * args.passes)
    avg_launch_us = elapsed_ms * 1e3 / max(stream_launches_per_pass * n_passes, 1)
    algo_bytes_per_launch = total_rows * bytes_per_row / stream_launches_per_pass

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "i32/f32/f64", "data": f"synthetic (torch Philox on device, seed {cfg['seed']:#x}+rank)",
        "config": workload_config(args, cfg),
        "selectivity": selectivity, "rows_out_per_pass": all_rows_out,
        "hbm_gbs_per_gpu": per_gpu_gbs, "pct_of_8TBs": per_gpu_gbs / 8000.0 * 100.0,
        "clocks": clocks, "gpu_launches": int(all_launches), "host_enqueue_ms_per_step": host_enqueue_ms / args.steps,
        "specialised_launches": int(jit_launches), "block_cache_misses_in_timed_region": ctx.alloc_misses - misses0,
        "overlapped_launch_sets": ctx.overlapped_launch_sets - overlapped0,
        "roofline": {"bound": "hbm",
                     "kernel": kernel_description(many, bool(jit_launches), split_launches),
                     "achieved": per_gpu_gbs, "peak": peak, "unit": "GB/s", "frac": per_gpu_gbs / peak,
                     "traffic": measured_traffic(args.config, args.batch_rows),
                     "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of ALL kernels of one record's launch set "
                                     "(select + gather, or the fused stream kernel) under ncu, cold L2 and serialised: the "
                                     "gather kernel's re-read of the predicate columns counts in full there, and part of the "
                                     "output is still dirty in L2 when the last kernel ends",
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_row": bytes_per_row, "algorithmic_bytes_per_launch": algo_bytes_per_launch,
                     "avg_launch_us": avg_launch_us, "stream_launches_per_pass": stream_launches_per_pass,
                     "timing": "CUDA events on the ctx stream around all timed passes (every kernel of every launch set: "
                               "zero_kernel + select + gather, or zero_kernel + the fused stream kernel)"},
    }
    if parity:
        line.update(parity)

    # ---- materialize-side gather (N > 1): every rank's compacted batches travel to rank 0 over NVLink ----
    if world > 1 and not args.no_gather:
        def gather_step(verify=False):
            outs_ = one_pass()
            torch.cuda.current_stream(device).wait_stream(stream)
            got = multigpu.gather_batches(flat(outs_), dst=0, verify=verify)
            return outs_, got
        gather_step(verify=True)   # (checksums of every slab compared across the link once, outside the timed steps)
        gather_step()
        barrier()
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        for _ in range(args.gather_steps):
            keep = gather_step()
        torch.cuda.synchronize(device)
        barrier()
        dtg = time.perf_counter() - t0
        gathered = multigpu.gathered_bytes(keep[1]) if rank == 0 else 0
        del keep
        dtg_ms, (gathered,) = multigpu.reduce_step(dtg * 1e3, [gathered], device)
        line["gather"] = {"value": all_rows * args.gather_steps / (dtg_ms / 1e3), "unit": UNIT, "steps": args.gather_steps,
                          "to_rank": 0, "ingest_gbs_rank0": gathered * args.gather_steps / (dtg_ms / 1e3) / 1e9,
                          "transport": multigpu.GATHER_TRANSPORT,
                          "note": "one pass of filter + variable-size gather of every result buffer to rank 0; the reference's "
                                  "per-record result files allow per-GPU materialize instead, which is what `value` measures"}

    # ---- e2e: host batches (pinned) -> chdb_filter_record -> host batches, copies inside the timed region ----
    if not args.no_e2e:
        import concurrent.futures as cf
        e2e_records = min(len(tensors), 2000) if many else len(tensors)
        host = [to_host_batch(cfg, t, pin=True) for t in tensors[:e2e_records]]
        e2e_rows = sum(rb.num_rows for rb, _ in host)
        n_inst = max(1, args.e2e_instances)
        ctxs = [ctx] + [C.Context(local_rank) for _ in range(n_inst - 1)]
        pool = cf.ThreadPoolExecutor(max_workers=n_inst)

        def worker(w):
            rows, nbytes = 0, 0
            for i in range(w, len(host), n_inst):
                out = prog.run(host[i][0], ctxs[w])
                rows += out.num_rows
                nbytes += batch_nbytes(out)
            return rows, nbytes

        def e2e_step():
            res = list(pool.map(worker, range(n_inst)))
            return sum(r for r, _ in res), sum(b for _, b in res)

        e2e_step()   # warm-up: pinned output pools, mempools
        e2e_step()
        barrier()
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        d2h = 0
        for _ in range(args.e2e_steps):
            _, d2h = e2e_step()
        torch.cuda.synchronize(device)
        dt = time.perf_counter() - t0
        dt, (all_e2e_rows,) = multigpu.reduce_step(dt * 1e3, [e2e_rows], device)
        h2d = sum(batch_nbytes(rb) for rb, _ in host)
        line["e2e"] = {"value": all_e2e_rows * args.e2e_steps / (dt / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": d2h, "steps": args.e2e_steps, "instances": n_inst, "rows_per_step": e2e_rows,
                       "path": "Program.run -> chdb_filter_record (Arrow C Data Interface, pinned host buffers)"}
        if rank == 0 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baselines(args, cfg, sel, [rb for rb, _ in host])
        pool.shutdown()
    elif rank == 0 and not args.no_cpu:
        host = [to_host_batch(cfg, t, pin=False)[0] for t in tensors[:max(4, 32_000_000 // args.batch_rows // 64)]]
        line["cpu_baseline"] = cpu_baselines(args, cfg, sel, host)

    if rank == 0:
        emit(line)


_REAL_STDOUT = None


def emit(line: dict):
    """The one JSON line, on the process's real stdout (see main())."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


# ---------------------------------------------------------------------------------------------
# f1: Parquet -> device decode (read_files_task.rs:233-282), the step upstream of the filter
# ---------------------------------------------------------------------------------------------
def run_parquet(args):
    """One step = decoding every row group of one in-memory Parquet file of the reference's sample schema
    (create_sample_data.rs: id Int32, value1 Utf8, value2 Float32; uncompressed, dictionary + v1 pages like parquet-rs's
    defaults) from PINNED host memory into device batches.  The file bytes start on the host, so the H2D copy is inside
    every timed step: value == e2e.  cpu_baseline: pyarrow's reader (Arrow C++) on one thread over the same bytes."""
    import io

    import numpy as np
    import pyarrow as pa
    import pyarrow.parquet as pq
    import torch

    import chapterhouseqe_b200 as C
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; chapterhouseqe_b200 has no CPU fallback")
    C.load_library()
    torch.cuda.set_device(0)
    ctx = C.Context(0)
    n = args.rows or 16_000_000
    rg_rows = args.batch_rows or (1 << 20)     # parquet-rs DEFAULT_MAX_ROW_GROUP_SIZE
    rng = np.random.default_rng(0xF1)
    table = pa.table({"id": pa.array(np.arange(n, dtype=np.int32)),
                      "value1": pa.array(np.char.add("v", rng.integers(0, 10**7, n).astype(str)) if args.parquet_plain
                                         else np.char.add("word", rng.integers(0, 1000, n).astype(str))),
                      "value2": pa.array(rng.uniform(0, 100, n).astype(np.float32))})
    buf = io.BytesIO()
    pq.write_table(table, buf, compression="NONE", row_group_size=rg_rows, use_dictionary=not args.parquet_plain)
    raw = buf.getvalue()
    pinned = torch.empty(len(raw), dtype=torch.uint8, pin_memory=True)
    pinned.numpy()[:] = np.frombuffer(raw, dtype=np.uint8)
    f = C.ParquetFile(pinned.numpy())
    arrow_bytes = sum(c.nbytes for c in table.columns)

    prog = None
    if args.parquet_filter:
        from chapterhouseqe_b200 import sqlparser_lite as sp
        prog = C.Program.compile_filter(sp.parse_expr("id % 2 = 0"), f.schema)

    def step():
        decoded = f.decode_row_groups(0, None, ctx)   # (row group i+1's H2D copy runs next to row group i's kernels)
        if prog is None:
            return decoded
        results = [b.run(prog) for b in decoded]      # filter on the decoded device batches, in place
        if args.parquet_encode:                       # ... and materialize on the device: only the result FILE crosses PCIe back
            return C.encode_parquet(results, 1 << 20, 0, ctx)
        return [r.download() for r in results]        # only the filtered rows cross PCIe back

    if args.parquet_encode and prog is None:
        raise SystemExit("bench.py --parquet --parquet-encode needs --parquet-filter (the whole query: read_files -> filter -> materialize)")
    outs = None
    for _ in range(max(args.warmup, 3)):
        if args.parquet_encode and outs is not None:
            outs.close()
        outs = step()
    if args.parquet_encode:
        return run_parquet_query_tail(args, ctx, f, raw, n, step, outs, arrow_bytes)
    # parity of what is timed: first and last row group against pyarrow's reader
    pf = pq.ParquetFile(io.BytesIO(raw))
    import pyarrow.compute as pc
    for i in sorted({0, f.num_row_groups - 1}):
        want = pf.read_row_group(i).combine_chunks()
        if prog is not None:   # (ids are non-negative: id % 2 = 0 <=> the low bit is clear)
            want = want.filter(pc.equal(pc.bit_wise_and(want.column("id"), 1), 0)).combine_chunks()
        got = outs[i] if prog is not None else outs[i].download()
        for name in want.schema.names:
            if not got.column(name).equals(want.column(name).chunk(0)):
                raise SystemExit(f"bench.py --parquet: row group {i} column {name} differs from pyarrow's reader")
    outs = None
    launches0 = ctx.launch_count
    ctx.synchronize()
    sampler = ClockSampler(0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        outs = step()
    ctx.synchronize()
    secs = time.perf_counter() - t0
    sampler.sample_while(lambda: False)
    launches = ctx.launch_count - launches0
    # CPU baseline: pyarrow reader, one thread, bounded sample
    d2h = 16 * 3 * f.num_row_groups + (sum(rb.nbytes for rb in outs) if prog is not None else 0)
    outs = None
    t1 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t1 < min(args.cpu_seconds, 10.0) or reps == 0:
        tb = pq.read_table(io.BytesIO(raw), use_threads=False)
        if prog is not None:
            tb = tb.filter(pc.equal(pc.bit_wise_and(tb.column("id"), 1), 0))
        reps += 1
    cpu_secs = (time.perf_counter() - t1) / reps
    peak, peak_src = measured_peak()
    value = n * args.steps / secs
    moved = (len(raw) + arrow_bytes) * args.steps / secs / 1e9
    e2e = {"value": value, "unit": "rows/s", "h2d_bytes_per_step": len(raw), "d2h_bytes_per_step": d2h,
           "path": "ParquetFile.decode_row_groups -> chdb_parquet_decode_row_groups from pinned host memory"
                   + (" -> DeviceBatch.run (chdb_run_device, filter id % 2 = 0) -> download of the filtered batches" if prog is not None
                      else " (result stays in HBM; per row group the null counts and string totals are read back)")}
    emit({"metric": "parquet_decode_filter_rows_per_s" if prog is not None else "parquet_decode_rows_per_s", "value": value, "unit": "rows/s", "n_gpus": 1, "steps": args.steps,
          "warmup": max(args.warmup, 3), "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
          "vs_baseline": None, "dtype": "u8/i32/f32", "data": "synthetic",
          "config": {"workload": f"f1 Parquet -> device decode, reference sample schema, {n} rows in {f.num_row_groups} row groups, "
                                 f"{'PLAIN' if args.parquet_plain else 'dictionary'} value1, uncompressed, v1 pages (pyarrow writer)",
                     "file_bytes": len(raw), "arrow_bytes": arrow_bytes,
                     "l2": "every step streams the whole file (larger than L2) from host memory"},
          "timing": "host wall clock around all steps (each row group's decode ends in a stream synchronise), H2D inside",
          "e2e": e2e, "gpu_launches": int(launches), "parity_checked_row_groups": 2, "clocks": sampler.result(),
          "roofline": {"bound": "hbm", "kernel": "pq_decode_rows / pq_walk_byte_arrays / pq_scan_* / pq_copy_utf8 + the H2D copy",
                       "achieved": moved, "peak": peak, "unit": "GB/s", "frac": moved / peak, "traffic": None, "peak_source": peak_src,
                       "note": "file bytes in + Arrow bytes out per second of the WHOLE call (PCIe copy and per-row-group "
                               "synchronise included): far from the HBM bound by construction; the PCIe link bounds it first"},
          "cpu_baseline": {"value": n / cpu_secs, "unit": "rows/s", "cores": 1, "kind": "port",
                           "sample": f"pyarrow {pa.__version__} parquet.read_table(use_threads=False) of the same bytes"
                                     + (" + Table.filter(id & 1 == 0)" if prog is not None else "") + f", {reps} run(s)"}})

# ---------------------------------------------------------------------------------------------
# f3: device batches -> Parquet with record coalescing (materialize_files_task.rs:116-141, DEV_NOTES.md:117-122)
# ---------------------------------------------------------------------------------------------
def run_parquet_encode(args):
    """One step = encoding device-resident records of the reference's sample schema (id Int32, value1 Utf8 with 5 % nulls,
    value2 Float32) into ONE Parquet file image in pinned host memory: records of --batch-rows rows (default 10 000, the
    reference's record size) coalesced into row groups of at most 1 Mi rows.  The image crosses PCIe inside every timed
    step: value == e2e.  cpu_baseline: pyarrow's writer (Arrow C++), one thread, PLAIN, uncompressed, same row groups."""
    import io

    import numpy as np
    import pyarrow as pa
    import pyarrow.parquet as pq
    import torch

    import chapterhouseqe_b200 as C
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; chapterhouseqe_b200 has no CPU fallback")
    C.load_library()
    torch.cuda.set_device(0)
    ctx = C.Context(0)
    n = args.rows or 16_000_000
    rec_rows = args.batch_rows or 10_000
    rng = np.random.default_rng(0xF3)
    table = pa.table({"id": pa.array(np.arange(n, dtype=np.int32)),
                      "value1": pa.array(np.char.add("v", rng.integers(0, 10**7, n).astype(str)), mask=rng.random(n) < 0.05),
                      "value2": pa.array(rng.uniform(0, 100, n).astype(np.float32))})
    recs = table.to_batches(max_chunksize=rec_rows)
    devs = C.DeviceBatchList.from_batches([C.DeviceBatch.upload(rb, ctx) for rb in recs])
    arrow_bytes = sum(c.nbytes for c in table.columns)
    img = None
    for _ in range(max(args.warmup, 3)):
        if img is not None:
            img.close()
        img = C.encode_parquet(devs, 1 << 20, 0, ctx)
    # parity of what is timed: the whole image through pyarrow's reader
    back = pq.read_table(io.BytesIO(img.to_bytes()), use_threads=True)
    for name in table.schema.names:
        if not back.column(name).combine_chunks().equals(table.column(name).combine_chunks()):
            raise SystemExit(f"bench.py --parquet-encode: column {name} read back by pyarrow differs from the records encoded")
    file_bytes, row_groups = img.nbytes, img.row_groups
    launches0 = ctx.launch_count
    ctx.synchronize()
    sampler = ClockSampler(0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        img.close()
        img = C.encode_parquet(devs, 1 << 20, 0, ctx)
    ctx.synchronize()
    secs = time.perf_counter() - t0
    sampler.sample_while(lambda: False)
    launches = ctx.launch_count - launches0
    img.close()
    t1 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t1 < min(args.cpu_seconds, 10.0) or reps == 0:
        sink = io.BytesIO()
        pq.write_table(table, sink, compression="NONE", use_dictionary=False, row_group_size=1 << 20, write_statistics=False)
        reps += 1
    cpu_secs = (time.perf_counter() - t1) / reps
    peak, peak_src = measured_peak()
    value = n * args.steps / secs
    moved = (arrow_bytes + file_bytes) * args.steps / secs / 1e9
    emit({"metric": "parquet_encode_rows_per_s", "value": value, "unit": "rows/s", "n_gpus": 1, "steps": args.steps,
          "warmup": max(args.warmup, 3), "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
          "vs_baseline": None, "dtype": "u8/i32/f32", "data": "synthetic",
          "config": {"workload": f"f3 device batches -> Parquet, reference sample schema (value1 5 % nulls), {n} rows in {len(recs)} records of "
                                 f"<= {rec_rows} rows coalesced into {row_groups} row groups, PLAIN, uncompressed, v1 pages",
                     "file_bytes": file_bytes, "arrow_bytes": arrow_bytes,
                     "l2": "every step streams all records (larger than L2) and writes the whole image to host memory"},
          "timing": "host wall clock around all steps (each encode ends in a stream synchronise), D2H of the image inside",
          "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": file_bytes,
                  "path": "encode_parquet -> chdb_parquet_encode: device-resident records -> file image in pinned host memory"},
          "gpu_launches": int(launches), "parity_checked": "whole image read back by pyarrow", "clocks": sampler.result(),
          "roofline": {"bound": "hbm", "kernel": "pqe_sizes / pqe_scan / pqe_write + device-to-device copies + the D2H copy",
                       "achieved": moved, "peak": peak, "unit": "GB/s", "frac": moved / peak, "traffic": None, "peak_source": peak_src,
                       "note": "Arrow bytes in + file bytes out per second of the WHOLE call (D2H copy and two synchronises included): "
                               "the PCIe link bounds it first"},
          "cpu_baseline": {"value": n / cpu_secs, "unit": "rows/s", "cores": 1, "kind": "port",
                           "sample": f"pyarrow {pa.__version__} parquet.write_table(use_dictionary=False, compression=NONE) of the same table, "
                                     f"{reps} run(s)"}})

def run_parquet_query_tail(args, ctx, f, raw, n, step, img, arrow_bytes):
    """--parquet --parquet-filter --parquet-encode: the whole query `select * from read_files(..) where id % 2 = 0`
    materialized to Parquet, file bytes in -> file bytes out, everything in between on the device (f1 -> filter -> f3)."""
    import io

    import pyarrow as pa
    import pyarrow.compute as pc
    import pyarrow.parquet as pq
    want = pq.read_table(io.BytesIO(raw))
    want = want.filter(pc.equal(pc.bit_wise_and(want.column("id"), 1), 0))
    got = pq.read_table(io.BytesIO(img.to_bytes()))
    for name in want.schema.names:
        if not got.column(name).combine_chunks().equals(want.column(name).combine_chunks()):
            raise SystemExit(f"bench.py --parquet query: column {name} of the written file differs from pyarrow read + filter")
    out_bytes, out_groups, out_rows = img.nbytes, img.row_groups, got.num_rows
    launches0 = ctx.launch_count
    ctx.synchronize()
    sampler = ClockSampler(0)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        img.close()
        img = step()
    ctx.synchronize()
    secs = time.perf_counter() - t0
    sampler.sample_while(lambda: False)
    launches = ctx.launch_count - launches0
    img.close()
    t1 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t1 < min(args.cpu_seconds, 10.0) or reps == 0:
        tb = pq.read_table(io.BytesIO(raw), use_threads=False)
        tb = tb.filter(pc.equal(pc.bit_wise_and(tb.column("id"), 1), 0))
        pq.write_table(tb, io.BytesIO(), compression="NONE", use_dictionary=False, row_group_size=1 << 20, write_statistics=False)
        reps += 1
    cpu_secs = (time.perf_counter() - t1) / reps
    peak, peak_src = measured_peak()
    value = n * args.steps / secs
    moved = (len(raw) + 2 * arrow_bytes + out_bytes) * args.steps / secs / 1e9
    emit({"metric": "parquet_query_rows_per_s", "value": value, "unit": "rows/s", "n_gpus": 1, "steps": args.steps,
          "warmup": max(args.warmup, 3), "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
          "vs_baseline": None, "dtype": "u8/i32/f32", "data": "synthetic",
          "config": {"workload": f"select * from read_files(..) where id % 2 = 0 -> Parquet: {n} input rows in {f.num_row_groups} row groups "
                                 f"(reference sample schema, {'PLAIN' if args.parquet_plain else 'dictionary'} value1) -> {out_rows} rows in "
                                 f"{out_groups} row groups; decode, filter and encode on the device",
                     "file_bytes": len(raw), "out_file_bytes": out_bytes,
                     "l2": "every step streams the whole file (larger than L2) from host memory"},
          "timing": "host wall clock around all steps, H2D of the input file and D2H of the output file inside",
          "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": len(raw), "d2h_bytes_per_step": out_bytes,
                  "path": "ParquetFile.decode_row_groups -> DeviceBatch.run (filter) -> encode_parquet: chdb_parquet_decode_row_groups -> "
                          "chdb_run_device -> chdb_parquet_encode; only file bytes cross PCIe"},
          "gpu_launches": int(launches), "parity_checked": "whole output file read back by pyarrow against pyarrow read + filter",
          "clocks": sampler.result(),
          "roofline": {"bound": "hbm", "kernel": "pq_* decode + select / gather + pqe_* encode + both PCIe copies",
                       "achieved": moved, "peak": peak, "unit": "GB/s", "frac": moved / peak, "traffic": None, "peak_source": peak_src,
                       "note": "bytes of the three stages per second of the WHOLE call: PCIe and per-row-group synchronises bound it"},
          "cpu_baseline": {"value": n / cpu_secs, "unit": "rows/s", "cores": 1, "kind": "port",
                           "sample": f"pyarrow {pa.__version__} read_table(use_threads=False) + Table.filter(id & 1 == 0) + write_table "
                                     f"(PLAIN, uncompressed), {reps} run(s)"}})


def main():
    global _REAL_STDOUT
    args = parse_args()
    # stdout carries exactly one JSON line: everything else a library may print there (NCCL's version banner, ...)
    # is sent to stderr by pointing file descriptor 1 at stderr for the whole run
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.parquet_encode and not args.parquet:
        if rank == 0:
            run_parquet_encode(args)
        return
    if args.parquet:
        if rank == 0:
            run_parquet(args)
        return
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own output (its version banner at NCCL_DEBUG >= VERSION) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, cfg, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
