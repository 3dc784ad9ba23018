"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: record placement, the whole-job
reduction bench.py reports, and the variable-size materialize gather."""
import os
import subprocess
import sys
import textwrap

from chapterhouseqe_b200 import multigpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_assign_records_round_robin():
    ids = list(range(10))
    assert multigpu.assign_records(ids, 1) == [ids]
    parts = multigpu.assign_records(ids, 4)
    assert parts == [[0, 4, 8], [1, 5, 9], [2, 6], [3, 7]]
    assert sorted(x for p in parts for x in p) == ids


def test_reduce_and_gather_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import torch, torch.distributed as dist
        from chapterhouseqe_b200 import multigpu
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        assert world == 2
        # per-rank shard of 7 records
        mine = multigpu.assign_records(list(range(7)), world)[rank]
        ms, sums = multigpu.reduce_step(10.0 + rank, [len(mine), 100.0 * (rank + 1)])
        assert ms == 11.0 and sums == [7.0, 300.0], (ms, sums)
        # variable-size gather: rank r sends buffers of r+1 and 3 bytes, and an empty one
        bufs = [torch.full((rank + 1,), rank + 1, dtype=torch.uint8), torch.full((3,), 9 - rank, dtype=torch.uint8),
                torch.empty(0, dtype=torch.uint8)]
        got = multigpu.gather_buffers(bufs, dst=0)
        if rank == 0:
            assert [[b.tolist() for b in src] for src in got] == [[[1], [9, 9, 9], []], [[2, 2], [8, 8, 8], []]], got
        else:
            assert got is None
        dist.barrier()
        dist.destroy_process_group()
        open(os.path.join({str(tmp_path)!r}, f"ok{{rank}}"), "w").write("ok")   # (stdout of two ranks can interleave)
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29653", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
