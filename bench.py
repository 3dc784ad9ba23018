#!/usr/bin/env python
"""bench.py -- filter + project + compaction throughput on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]           our arm (CUDA, device-resident + e2e)
  python bench.py --impl reference [...]                         the CPU restatement on all host threads

Workload (BASELINE.json configs[1], SURVEY.md 8d "C2"): large_simple.sql-shaped synthetic rows,
100 M rows per GPU, columns id Int32 | k Int64 | value2 Float32 (10 % null) | d Float64 (5 % null) |
value1 Utf8 (8 bytes), predicate `(id % 2 = 0 AND value2 > 10.0) OR d < 0.5` with the reference's
non-Kleene null semantics, `select *`, in device-native batches of 2^22 rows.  One step = one pass of
the fused kernel over every batch.  Multi-GPU: one process per GPU (torchrun), batches shard by
file / row group, so there is no collective on the data path ("weak" scaling: 100 M rows per GPU).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PREDICATE = "(id % 2 = 0 and value2 > 10.0) or d < 0.5"
SQL = f"select * from read_files('large_simple/*.parquet') where {PREDICATE}"
METRIC = "filter_project_rows_per_s"
UNIT = "rows/s"
SEED = 0xC4DB0002
NULL_V2, NULL_D = 0.10, 0.05
STR_LEN = 8
# algorithmic bytes per input row: id 4 + k 8 + value2 (4 + 1/8) + d (8 + 1/8) + value1 (4 offset + 8 bytes)
IN_BYTES_PER_ROW = 4 + 8 + (4 + 0.125) + (8 + 0.125) + (4 + STR_LEN)


if os.environ.get("CHDB_BENCH_WATCHDOG"):   # debugging aid: dump every thread's stack if the run is still going after N seconds
    import faulthandler
    faulthandler.dump_traceback_later(float(os.environ["CHDB_BENCH_WATCHDOG"]), exit=True)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--rows", type=int, default=100_000_000, help="rows per GPU")
    ap.add_argument("--batch-rows", type=int, default=1 << 22)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-instances", type=int, default=3, help="filter operator instances (ctx + stream each) for e2e")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget for the single-thread cpu_baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="skip the N>1 materialize-gather measurement")
    ap.add_argument("--gather-steps", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


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


def gen_batch_torch(start: int, n: int, seed: int, device):
    """One batch as torch tensors (values padded so every buffer is readable 64 bytes past its end)."""
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
        return out, int(valid.sum().item())

    ids = torch.zeros(n + pad, dtype=torch.int32, device=device)
    ids[:n] = torch.arange(start, start + n, dtype=torch.int64, device=device).to(torch.int32)
    k = torch.zeros(n + pad, dtype=torch.int64, device=device)
    k[:n] = torch.randint(-2**31, 2**31, (n,), generator=g, device=device, dtype=torch.int64)
    v2 = torch.zeros(n + pad, dtype=torch.float32, device=device)
    v2[:n] = torch.rand(n, generator=g, device=device) * 100.0
    v2_valid, v2_nvalid = bitmap(NULL_V2)
    d = torch.zeros(n + pad, dtype=torch.float64, device=device)
    d[:n] = torch.randn(n, generator=g, device=device, dtype=torch.float64)
    d_valid, d_nvalid = bitmap(NULL_D)
    s = torch.zeros(n * STR_LEN + pad, dtype=torch.uint8, device=device)
    s[: n * STR_LEN] = torch.randint(97, 123, (n * STR_LEN,), generator=g, device=device, dtype=torch.int64).to(torch.uint8)
    offs = torch.zeros(n + 1 + pad, dtype=torch.int32, device=device)
    offs[: n + 1] = torch.arange(0, (n + 1) * STR_LEN, STR_LEN, dtype=torch.int64, device=device).to(torch.int32)
    return dict(n=n, id=ids, k=k, value2=v2, value2_valid=v2_valid, value2_nulls=n - v2_nvalid, d=d, d_valid=d_valid,
                d_nulls=n - d_nvalid, value1=s, value1_offsets=offs)


def schema():
    import pyarrow as pa
    return pa.schema([pa.field("id", pa.int32(), False), pa.field("k", pa.int64(), False),
                      pa.field("value2", pa.float32(), True), pa.field("d", pa.float64(), True),
                      pa.field("value1", pa.utf8(), False)])


def to_host_batch(t, pin: bool):
    """torch tensors -> pyarrow RecordBatch over (optionally pinned) host memory, zero-copy."""
    import pyarrow as pa
    import torch
    n = t["n"]

    def host(x, count):
        h = torch.empty(count, dtype=x.dtype, pin_memory=pin)
        h.copy_(x[:count])
        return h

    keep = {}
    keep["id"] = host(t["id"], n)
    keep["k"] = host(t["k"], n)
    keep["value2"] = host(t["value2"], n)
    keep["d"] = host(t["d"], n)
    keep["value1"] = host(t["value1"], n * STR_LEN)
    keep["offs"] = host(t["value1_offsets"], n + 1)
    keep["v2v"] = host(t["value2_valid"], (n + 7) // 8)
    keep["dv"] = host(t["d_valid"], (n + 7) // 8)
    buf = lambda h: pa.py_buffer(h.numpy())  # noqa: E731
    arrays = [
        pa.Array.from_buffers(pa.int32(), n, [None, buf(keep["id"])]),
        pa.Array.from_buffers(pa.int64(), n, [None, buf(keep["k"])]),
        pa.Array.from_buffers(pa.float32(), n, [buf(keep["v2v"]), buf(keep["value2"])], null_count=t["value2_nulls"]),
        pa.Array.from_buffers(pa.float64(), n, [buf(keep["dv"]), buf(keep["d"])], null_count=t["d_nulls"]),
        pa.Array.from_buffers(pa.utf8(), n, [None, buf(keep["offs"]), buf(keep["value1"])]),
    ]
    rb = pa.RecordBatch.from_arrays(arrays, schema=schema())
    return rb, keep


def batch_nbytes(rb) -> int:
    return sum(b.size for col in rb.columns for b in col.buffers() if b is not None)


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons of one GPU during the timed region (NVML).

    An NVML query -- from a thread of this process or from `nvidia-smi -lms` next to it -- stalls this process's
    CUDA calls for milliseconds to tens of milliseconds now and then (measured: host enqueue time per step 0.8 ms
    -> 2-13 ms with a 20 ms sampling period), which starves the GPU in a timed region that is itself only tens
    of milliseconds long.  So the samples are taken by the benchmark thread itself after it has enqueued all
    timed steps and before it synchronises: the GPU is still working through the queue (under load, inside the
    timed region) and the host has nothing left to enqueue that a stall could delay."""

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
            time.sleep(0.002)

    def result(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "sampled": "NVML, by the benchmark thread between enqueueing the last timed step and synchronising"}


def measured_traffic(batch_rows: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture of this same workload (profiles/traffic.json); None if never captured."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(p))
        if int(t.get("batch_rows", -1)) == int(batch_rows):
            return t["dram_bytes_per_launch"]
    except Exception:  # noqa: BLE001
        pass
    return None


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
def cpu_filter(rb, expr):
    from oracle import compute_value as O
    b = O.batch_from_arrow(rb)
    return O.filter_record(b, [[] for _ in b.fields], expr)


def cpu_baseline_single_thread(host_batches, expr, budget_s: float):
    t0 = time.perf_counter()
    rows = 0
    used = 0
    for rb in host_batches:
        cpu_filter(rb, expr)
        rows += rb.num_rows
        used += 1
        if time.perf_counter() - t0 >= budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": rows / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{used} batch(es) = {rows} rows of the same workload, oracle/arrow_kernels.c + tree walk, 1 thread, {dt:.2f} s"}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port; the Rust reference cannot be built
    here) on all host threads, one batch per thread like N filter instances pulling from one exchange."""
    if rank != 0:
        return
    import concurrent.futures as cf

    import torch

    from chapterhouseqe_b200 import sqlparser_lite as sp
    from oracle import compute_value as O
    O.lib()
    expr = sp.parse_expr(PREDICATE)
    cores = os.cpu_count() or 1
    # bounded sample of the workload: 2 batches per thread (at least 8), generated like our arm's data
    n_batches = min(len(batch_sizes(args.rows, args.batch_rows)), max(8, 2 * cores))
    dev = torch.device("cuda:0") if torch.cuda.is_available() else torch.device("cpu")
    host = []
    for i, (start, n) in enumerate(batch_sizes(args.rows, args.batch_rows)[:n_batches]):
        t = gen_batch_torch(start, n, SEED + i, dev)
        rb, keep = to_host_batch(t, pin=False)
        host.append((rb, keep))
        del t
    sample_rows = sum(rb.num_rows for rb, _ in host)
    pool = cf.ThreadPoolExecutor(max_workers=cores)

    def step():
        list(pool.map(lambda x: cpu_filter(x[0], expr), host))

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
        "config": workload_config(args, sample_rows=sample_rows),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n_batches} batches = {sample_rows} rows per step, one batch per thread on {cores} threads "
                                   "(oracle port of the arrow-rs path; the Rust reference cannot be built in this image)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, sample_rows=None):
    cfg = {"workload": "C2 large_simple-shaped synthetic: " + SQL, "rows_per_gpu": args.rows,
           "batch_rows": args.batch_rows, "schema": "id i32 | k i64 | value2 f32 (10% null) | d f64 (5% null) | value1 utf8[8]",
           "null_semantics": "non-Kleene (arrow compute::and/or)", "parallelism": f"shard-by-batch x{args.gpus}, no collective",
           "l2": "inputs (3.6 GB per GPU, 152 MB per batch) exceed the 126 MB L2; no flush needed"}
    if sample_rows is not None:
        cfg["sample_rows_per_step"] = sample_rows
    return cfg


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
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
    sel = sp.parse_select(SQL)
    prog = C.Program.compile_filter(sel["selection"], schema())

    # ---- resident inputs ----
    sizes = batch_sizes(args.rows, args.batch_rows)
    base_row = rank * args.rows
    tensors, dev_batches = [], []
    for i, (start, n) in enumerate(sizes):
        t = gen_batch_torch(base_row + start, n, SEED + rank * 100003 + i, device)
        tensors.append(t)
        vals = [t["id"].data_ptr(), t["k"].data_ptr(), t["value2"].data_ptr(), t["d"].data_ptr(), t["value1"].data_ptr()]
        vald = [0, 0, t["value2_valid"].data_ptr(), t["d_valid"].data_ptr(), 0]
        offs = [0, 0, 0, 0, t["value1_offsets"].data_ptr()]
        dev_batches.append(C.DeviceBatch.wrap(schema(), n, vals, vald, offs, ctx=ctx, keepalive=t))
    torch.cuda.synchronize(device)
    total_rows = sum(n for _, n in sizes)

    def barrier():
        if world > 1:
            dist.barrier()

    def one_step():
        return [b.run(prog) for b in dev_batches]

    sampler = ClockSampler(local_rank)   # NVML initialised (and queried once) before the warm-up
    outs = None
    for _ in range(max(args.warmup, 0)):
        outs = one_step()
    ctx.synchronize()
    torch.cuda.synchronize(device)
    rows_out = sum(o.num_rows for o in outs) if outs else 0
    bytes_out = sum(o.nbytes for o in outs) if outs else 0
    for o in outs or []:
        o.check()
    outs = None

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0, jit0 = ctx.launch_count, ctx.jit_launch_count
    barrier()
    torch.cuda.synchronize(device)
    # a cyclic-GC pause in the middle of the enqueue loop (tens of ms with torch / pyarrow loaded) starves the GPU,
    # which is only a few ms behind the host: collect now, not during the timed region
    import gc
    gc.collect()
    gc.disable()
    ev0.record(stream)
    prev = None
    t_host0 = time.perf_counter()
    trace = [] if os.environ.get("CHDB_BENCH_TRACE") else None
    for _ in range(args.steps):
        if trace is not None:
            cur = []
            for b in dev_batches:
                t0 = time.perf_counter()
                cur.append(b.run(prog))
                trace.append(("run", (time.perf_counter() - t0) * 1e6))
            t0 = time.perf_counter()
            prev = cur
            trace.append(("release", (time.perf_counter() - t0) * 1e6))
        else:
            cur = one_step()
            prev = cur   # the previous step's outputs are released here (they go back to the ctx's block cache)
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3
    if trace is not None:
        sys.stderr.write("[bench trace] " + " ".join(f"{k}:{v:.0f}" for k, v in trace) + "\n")
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
    if rows_out == 0 and prev:
        rows_out = sum(o.num_rows for o in prev)
        bytes_out = sum(o.nbytes for o in prev)
    prev = None

    # whole-job numbers: MAX over ranks of the device time, SUM of rows / launches (no data-path collective)
    elapsed_ms, (all_rows, all_rows_out, all_launches, jit_launches) = multigpu.reduce_step(
        elapsed_ms, [total_rows, rows_out, launches, jit_launches], device)

    selectivity = rows_out / total_rows
    # algorithmic bytes: every referenced input byte once + every output byte once
    out_bytes_per_row = IN_BYTES_PER_ROW   # select *: same columns (output validity kept: nulls survive)
    bytes_per_row = IN_BYTES_PER_ROW + selectivity * out_bytes_per_row
    launches_per_step = len(sizes)
    secs = elapsed_ms / 1e3
    value = all_rows * args.steps / secs
    per_gpu_gbs = total_rows * args.steps * bytes_per_row / secs / 1e9
    peak, peak_src = measured_peak()
    avg_launch_us = elapsed_ms * 1e3 / max(launches_per_step * args.steps, 1)   # one batch = select + scan + gather
    algo_bytes_per_launch = total_rows * bytes_per_row / launches_per_step

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "i32/f32/f64", "data": "synthetic (torch Philox on device, seed 0xC4DB0002+rank)",
        "config": workload_config(args),
        "selectivity": selectivity, "rows_out_per_step": all_rows_out,
        "hbm_gbs_per_gpu": per_gpu_gbs, "pct_of_8TBs": per_gpu_gbs / 8000.0 * 100.0,
        "clocks": clocks, "gpu_launches": int(all_launches), "host_enqueue_ms_per_step": host_enqueue_ms / args.steps,
        "specialised_launches": int(jit_launches),
        "roofline": {"bound": "hbm",
                     "kernel": ("chdb_jit_select + scan_kernel + chdb_jit_gather per batch (device_code.cuh specialised for the "
                                "program by NVRTC); gather dominates" if jit_launches
                                else "select_kernel + scan_kernel + gather_kernel per batch (bytecode interpreter)"),
                     "achieved": per_gpu_gbs, "peak": peak,
                     "unit": "GB/s", "frac": per_gpu_gbs / peak, "traffic": measured_traffic(args.batch_rows),
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_row": bytes_per_row, "algorithmic_bytes_per_launch": algo_bytes_per_launch,
                     "avg_batch_us": avg_launch_us, "batches_per_step": launches_per_step, "kernels_per_batch": 3},
    }

    # ---- materialize-side gather (N > 1): every rank's compacted batches travel to rank 0 over NVLink ----
    if world > 1 and not args.no_gather:
        def gather_step():
            outs_ = one_step()
            bufs = [t_ for o in outs_ for t_ in multigpu.device_batch_buffers(o)]
            torch.cuda.current_stream(device).wait_stream(stream)
            got = multigpu.gather_buffers(bufs, dst=0)
            return outs_, got
        gather_step()
        barrier()
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        for _ in range(args.gather_steps):
            keep = gather_step()
        torch.cuda.synchronize(device)
        barrier()
        dtg = time.perf_counter() - t0
        del keep
        dtg_ms, _ = multigpu.reduce_step(dtg * 1e3, [0.0], device)
        line["gather"] = {"value": all_rows * args.gather_steps / (dtg_ms / 1e3), "unit": UNIT, "steps": args.gather_steps,
                          "to_rank": 0, "transport": "torch.distributed isend/irecv (NCCL p2p over NVLink)",
                          "note": "filter + variable-size gather of every output buffer to rank 0; the reference's per-record "
                                  "result files allow per-GPU materialize instead, which is what `value` measures"}

    # ---- e2e: host batches (pinned) -> chdb_filter_record -> host batches, copies inside the timed region ----
    if not args.no_e2e:
        import concurrent.futures as cf
        host = [to_host_batch(t, pin=True) for t in tensors]
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
        dt = multigpu.reduce_step(dt * 1e3, [0.0], device)[0] / 1e3
        h2d = sum(batch_nbytes(rb) for rb, _ in host)
        line["e2e"] = {"value": all_rows * args.e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": d2h, "steps": args.e2e_steps, "instances": n_inst,
                       "path": "Program.run -> chdb_filter_record (Arrow C Data Interface, pinned host buffers)"}
        if rank == 0 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_single_thread([rb for rb, _ in host], sel["selection"], args.cpu_seconds)
        pool.shutdown()
    elif rank == 0 and not args.no_cpu:
        host = [to_host_batch(t, pin=False)[0] for t in tensors[:4]]
        line["cpu_baseline"] = cpu_baseline_single_thread(host, sel["selection"], args.cpu_seconds)

    if rank == 0:
        print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own output (its version banner at NCCL_DEBUG >= VERSION) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
