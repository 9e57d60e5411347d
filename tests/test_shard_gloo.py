"""CPU tier, world_size 2 over gloo: the N>1 plumbing of the path.  The path shards as independent
chunks of row tiles with NO data-path collective (multi-GPU); what the ranks must agree on is
the chunk plan and a disjoint, complete, balanced ownership -- checked here with the product's own host
logic (gkm_plan_chunks / gkm_chunk_owner) and the same max-over-ranks reduction bench.py uses."""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gkmqc_b200 import capi
    lib = capi.load()

    class Chunk(ctypes.Structure):
        _fields_ = [("row_begin", ctypes.c_int), ("row_end", ctypes.c_int), ("col_begin", ctypes.c_int),
                    ("col_end", ctypes.c_int), ("entries", ctypes.c_longlong)]
    buf = (Chunk * 4096)()
    lib.gkm_plan_chunks.argtypes = [ctypes.c_int] * 6 + [ctypes.c_longlong, ctypes.POINTER(Chunk), ctypes.c_int]
    budget = max(4 << 20, n * n // 2 * 8 // (16 * world))
    nchunks = lib.gkm_plan_chunks(0, n, 0, n, 1, 16, budget, buf, 4096)
    mine = [c for c in range(nchunks) if lib.gkm_chunk_owner(c, nchunks, world) == rank]
    owned_rows = torch.zeros(n, dtype=torch.int32)
    entries = 0
    for c in mine:
        owned_rows[buf[c].row_begin:buf[c].row_end] += 1
        entries += buf[c].entries
    # what bench.py does: sum of per-rank row ownership must be exactly one everywhere, entries add up,
    # and the timed quantity is the max over ranks
    dist.all_reduce(owned_rows, op=dist.ReduceOp.SUM)
    tot = torch.tensor([float(entries)], dtype=torch.float64)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    mx = torch.tensor([float(entries)], dtype=torch.float64)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        q.put((int(owned_rows.min()), int(owned_rows.max()), float(tot.item()), float(mx.item()), nchunks))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [10000, 14144])
def test_two_ranks_partition_the_triangle(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lo, hi, total, mx, nchunks = res
    assert (lo, hi) == (1, 1), "every row owned by exactly one rank"
    assert total == n * (n - 1) // 2
    assert nchunks >= 16
    assert mx <= 0.56 * total, "ownership is balanced to within one chunk (max share %.3f)" % (mx / total)
