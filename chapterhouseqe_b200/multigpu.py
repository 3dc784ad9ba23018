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


_SLABS: dict = {}


def _slab(key, nbytes: int, device):
    """A reusable device buffer (grown, never shrunk): gather traffic lands in / leaves from the same allocation every step."""
    import torch
    t = _SLABS.get(key)
    if t is None or t.numel() < nbytes or t.device != device:
        t = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _SLABS[key] = t
    return t


def gather_batches(batches, dst: int = 0, verify: bool = False):
    """The materialize-side gather: every result batch of every rank onto rank `dst`.

    Each rank packs all buffers of its batches into ONE slab (chdb_device_batches_pack: device-to-device copies on
    the ctx stream) and sends it as ONE NCCL message; rank `dst` posts one receive per source rank, all inside one
    ncclGroupStart/End.  The pieces' sizes travel next to the payload, so `dst` can take the slab apart again
    (order: batch, column, {validity, offsets, values}; pieces 256-byte aligned).  Returns on dst a list per source
    rank of (slab[:total] tensor, sizes tensor); elsewhere None.  verify: also compares a checksum of every slab
    across the link."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from .api import DeviceBatchList
    world, rank = dist.get_world_size(), dist.get_rank()
    bl = batches if isinstance(batches, DeviceBatchList) else DeviceBatchList.from_batches(list(batches))
    device = torch.device("cuda", bl.ctx.device)
    total, sizes = bl.pack()                      # waits for the batches' counts, sizes the slab
    out_slab = _slab(("send",), total, device)
    if total:
        bl.pack(out_slab.data_ptr(), out_slab.numel())
    # NCCL runs on torch's stream: order it behind the copies the library enqueued on the ctx stream
    torch.cuda.current_stream(device).wait_stream(torch.cuda.ExternalStream(bl.ctx.stream, device=device))
    sizes_t = torch.from_numpy(np.frombuffer(sizes, dtype=np.int64).copy()).to(device)
    meta = torch.tensor([total, sizes_t.numel()], dtype=torch.int64, device=device)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    ops, out = [], None
    if rank == dst:
        out = []
        for src in range(world):
            if src == dst:
                out.append((out_slab[:total], sizes_t))
                continue
            n, m = (int(x) for x in metas[src].tolist())
            slab = _slab(("recv", src), n, device)
            st = torch.empty(m, dtype=torch.int64, device=device)
            if n:
                ops.append(dist.P2POp(dist.irecv, slab[:n], src))
            if m:
                ops.append(dist.P2POp(dist.irecv, st, src))
            out.append((slab[:n], st))
    else:
        if total:
            ops.append(dist.P2POp(dist.isend, out_slab[:total], dst))
        if sizes_t.numel():
            ops.append(dist.P2POp(dist.isend, sizes_t, dst))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    if verify:
        mine = out_slab[:total].to(torch.int64).sum().reshape(1) if total else torch.zeros(1, dtype=torch.int64, device=device)
        sums = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(sums, mine)
        if rank == dst:
            for src, (slab, _) in enumerate(out):
                got = int(slab.to(torch.int64).sum()) if slab.numel() else 0
                if got != int(sums[src]):
                    raise AssertionError(f"gather: slab of rank {src} arrived with checksum {got}, sent {int(sums[src])}")
    return out


GATHER_TRANSPORT = ("per rank ONE packed slab (chdb_device_batches_pack, device-to-device) sent as ONE NCCL message; rank 0 posts one "
                    "receive per source inside one ncclGroupStart/End (torch.distributed.batch_isend_irecv) over NVLink")


def gathered_bytes(got) -> int:
    """Bytes rank `dst` received from the other ranks (its own slab is not counted)."""
    import torch.distributed as dist
    if got is None:
        return 0
    rank = dist.get_rank()
    return sum(int(slab.numel()) for src, (slab, _) in enumerate(got) if src != rank)


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
