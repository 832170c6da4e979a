"""Host-side logic of the multi-GPU path on CPU: row-block dealing and the all-gather (gloo, world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gpsig_b200 import parallel as P


@pytest.mark.parametrize("n,ws", [(4096, 8), (4096, 2), (1000, 4), (100, 3), (7, 4), (1, 2)])
def test_partition_covers_rows_and_balances_the_triangle(n, ws):
    parts = P.partition(n, ws)
    rows = np.sort(np.concatenate([P.rows_of(p) for p in parts]))
    assert (rows == np.arange(n)).all()
    if n >= 64 * ws:
        work = [sum(int((n - np.arange(b, e)).sum()) for b, e in p) for p in parts]
        assert max(work) / min(work) < 1.02


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, n, n2, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    full = torch.arange(n * n2, dtype=torch.float32).reshape(n, n2)
    parts = P.partition(n, ws)
    rows = P.rows_of(parts[rank])
    out = P.gather_rows(full[rows], n, parts)
    q.put((rank, bool(torch.equal(out, full))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,n2", [(50, 50), (17, 5)])
def test_gather_rows_gloo_world_size_2(n, n2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, n2, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


@pytest.mark.parametrize("n,ws", [(4096, 8), (10, 3), (2, 4), (0, 2)])
def test_column_shards_cover_the_sequence_axis(n, ws):
    sh = P.column_shards(n, ws)
    assert len(sh) == ws and sh[0][0] == 0 and sh[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
    sizes = [e - b for b, e in sh]
    assert max(sizes) - min(sizes) <= 1


def _gather_cols_worker(rank, ws, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    n, rows = 11, 3
    full = torch.arange(rows * n, dtype=torch.float32).reshape(rows, n)
    sh = P.column_shards(n, ws)
    b, e = sh[rank]
    got = P.gather_columns(full[:, b:e].contiguous(), sh)
    q.put((rank, bool(torch.equal(got, full))))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_columns_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_cols_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)], res
