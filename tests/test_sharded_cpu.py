"""CPU tests (gloo, world_size 2) of the host side of strip-sharded execution: the strip
partition and the one all-gather that carries the order-r tails (recfilter_b200/sharded.py).
The kernels on either side of the exchange are covered on the GPU by the virtual-strip tests of
tests/test_fused_gpu.py; the real NCCL path by bench.py --gpus N."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from recfilter_b200.sharded import exchange_tails, gather_carry_chunks, scatter_tail_chunks, strip_bounds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, elems, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # tails of rank r, image b: value encodes (r, b, element)
        mine = torch.arange(elems, dtype=torch.float32).repeat(batch, 1)
        mine += 1000.0 * rank + 100.0 * torch.arange(batch, dtype=torch.float32).unsqueeze(1)
        g = exchange_tails(mine, world)
        ok = list(g.shape) == [batch, world, elems]
        for b in range(batch):
            for r in range(world):
                expect = torch.arange(elems, dtype=torch.float32) + 1000.0 * r + 100.0 * b
                ok = ok and bool(torch.equal(g[b, r], expect))
        q.put((rank, ok))
    except Exception as exc:            # report instead of leaving the parent waiting on the queue
        q.put((rank, f"{type(exc).__name__}: {exc}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch,elems", [(1, 7), (4, 48)])
def test_tail_exchange_layout_gloo_world2(batch, elems):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, elems, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in results) == [0, 1]
    assert all(ok is True for _, ok in results), results


def test_single_rank_exchange_is_identity():
    t = torch.rand(3, 5)
    g = exchange_tails(t, 1)
    assert g.shape == (3, 1, 5) and torch.equal(g[:, 0], t)


def test_strip_bounds():
    assert strip_bounds(8192, 4, 0) == (0, 2048)
    assert strip_bounds(8192, 4, 3) == (6144, 8192)
    assert strip_bounds(512, 8, 5, multiple=64) == (320, 384)
    with pytest.raises(ValueError):
        strip_bounds(1000, 3, 0)
    with pytest.raises(ValueError):
        strip_bounds(8192, 8, 0, multiple=2048)


def _chunk_worker(rank, world, port, vectors, lines, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # tails of shard r: value encodes (shard, vector, line)
        def tails_of(r):
            v = torch.arange(vectors, dtype=torch.float32).unsqueeze(1) * 10000.0
            return (v + torch.arange(lines, dtype=torch.float32).unsqueeze(0) + 1e6 * r).reshape(-1)
        c = lines // world
        recv = scatter_tail_chunks(tails_of(rank), vectors, world)          # [shard][vector][my chunk of the lines]
        ok = list(recv.shape) == [world, vectors, c]
        for g in range(world):
            ok = ok and bool(torch.equal(recv[g], tails_of(g).view(vectors, lines)[:, rank * c:(rank + 1) * c]))
        # a stand-in resolver: the carry entering shard g is the sum of the tails of the shards before it
        ext_all = torch.stack([recv[:g].sum(0) if g else torch.zeros_like(recv[0]) for g in range(world)])
        ext = gather_carry_chunks(ext_all)
        want = sum((tails_of(g) for g in range(rank)), torch.zeros(vectors * lines))
        ok = ok and bool(torch.equal(ext, want))
        q.put((rank, ok))
    except Exception as exc:
        q.put((rank, f"{type(exc).__name__}: {exc}"))
    finally:
        dist.destroy_process_group()


def test_column_chunked_exchange_gloo_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_chunk_worker, args=(r, world, port, 6, 32, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in results) == [0, 1]
    assert all(ok is True for _, ok in results), results
