"""Multi-GPU plumbing: one process per GPU (torch.distributed), batches shard by record id.

The reference hands every record (one Parquet row-group chunk, `record_id` assigned in
read_files_task.rs:284-288) to exactly one puller per consumer operator
(exchange_operator.rs:621-667), so N filter instances on N GPUs drain one queue and the hot path
needs NO collective.  What does move between GPUs is the materialize-side gather of
the compacted per-GPU results (variable-size point-to-point over NVLink).  These helpers are the
host-side logic of both; they run on NCCL (GPU) and gloo (CPU tests) alike.
"""
from __future__ import annotations

from typing import Sequence


def assign_records(record_ids: Sequence[int], world: int) -> list[list[int]]:
    """Static round-robin placement g = record_id mod G of records on GPUs."""
    if world < 1:
        raise ValueError("world must be >= 1")
    out: list[list[int]] = [[] for _ in range(world)]
    for r in record_ids:
        out[r % world].append(r)
    return out


def reduce_step(elapsed_ms: float, sums: Sequence[float], device=None):
    """Whole-job numbers from per-rank ones: MAX over ranks of the device time, SUM of everything else."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(elapsed_ms), [float(x) for x in sums]
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    s = torch.tensor([float(x) for x in sums], dtype=torch.float64, device=device)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return float(t.item()), [float(x) for x in s.tolist()]


GATHER_TRANSPORT = ("one ncclGroupStart/End per step (torch.distributed.batch_isend_irecv): every result buffer of every "
                    "rank is one grouped NCCL send/recv over NVLink, received into one slab per source rank")


def gather_buffers(buffers: Sequence, dst: int = 0, packed: bool | None = None):
    """Variable-size gather of byte buffers (1-D uint8 tensors, one list per rank) onto rank `dst`.

    Sizes travel in one all_gather; the payloads as ONE group of point-to-point operations
    (ncclGroupStart/End through batch_isend_irecv), so NCCL runs them concurrently on all its channels instead
    of serialising one kernel per buffer; the receiver lands every source rank's buffers in one slab (one
    allocation per source, 256-byte aligned pieces).  packed=True first concatenates a rank's buffers on the
    sender (one extra HBM pass, one message per rank).  Returns on dst a list (per source rank) of lists of
    tensors; elsewhere None."""
    import os

    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    if packed is None:
        packed = os.environ.get("CHDB_GATHER_PACKED", "0") == "1"
    dev = buffers[0].device if buffers else torch.device("cpu")
    sizes = torch.tensor([int(b.numel()) for b in buffers], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    up = lambda n: (n + 255) // 256 * 256  # noqa: E731
    ops, out = [], None
    if rank == dst:
        out = []
        for src in range(world):
            if src == dst:
                out.append(list(buffers))
                continue
            ns = [int(n) for n in all_sizes[src].tolist()]
            if packed:
                slab = torch.empty(sum(ns), dtype=torch.uint8, device=dev)
                views, at = [], 0
                for n in ns:
                    views.append(slab[at:at + n])
                    at += n
                if slab.numel():
                    ops.append(dist.P2POp(dist.irecv, slab, src))
            else:
                slab = torch.empty(sum(up(n) for n in ns), dtype=torch.uint8, device=dev)
                views, at = [], 0
                for n in ns:
                    v = slab[at:at + n]
                    views.append(v)
                    at += up(n)
                    if n:
                        ops.append(dist.P2POp(dist.irecv, v, src))
            out.append(views)
    elif packed:
        nonempty = [b for b in buffers if b.numel()]
        if nonempty:
            ops.append(dist.P2POp(dist.isend, torch.cat(nonempty), dst))
    else:
        ops = [dist.P2POp(dist.isend, b, dst) for b in buffers if b.numel()]
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    return out


def gather_batches(batches: Sequence, dst: int = 0):
    """The materialize-side gather: every buffer of every result DeviceBatch of every rank onto rank `dst`."""
    bufs = [t for b in batches for t in device_batch_buffers(b)]
    return gather_buffers(bufs, dst=dst)


def gathered_bytes(got) -> int:
    """Bytes rank `dst` received from the other ranks (its own buffers are not counted)."""
    import torch.distributed as dist
    if got is None:
        return 0
    rank = dist.get_rank()
    return sum(int(t.numel()) for src, bufs in enumerate(got) if src != rank for t in bufs)


class _CudaView:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can wrap it without a copy."""

    def __init__(self, ptr: int, nbytes: int, keepalive=None):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
        self._keepalive = keepalive


def device_batch_buffers(batch) -> list:
    """Every buffer of a DeviceBatch (values / validity / offsets per column) as 1-D uint8 CUDA tensors, zero-copy."""
    import torch
    out = []
    for c in range(batch.num_columns):
        bufs = batch.column_buffers(c)
        for key in ("values", "validity", "offsets"):
            ptr, n = bufs[key]
            if ptr and n:
                out.append(torch.as_tensor(_CudaView(ptr, n, batch), device=f"cuda:{batch.ctx.device}"))
            else:
                out.append(torch.empty(0, dtype=torch.uint8, device=f"cuda:{batch.ctx.device}"))
    return out
