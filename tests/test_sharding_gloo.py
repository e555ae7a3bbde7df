"""World-size-2 test (gloo, CPU) of the host logic of the multi-GPU path: shard bounds, env start mapping and the
per-step gather of packed rewards / dones / flags."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pypownet_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = sharding.shard_bounds(n_total, rank, world)
    B = hi - lo
    g = torch.arange(lo, hi, dtype=torch.float64)
    reward = torch.stack([g * k for k in range(1, 6)], dim=1)          # a function of the global env index
    done = (torch.arange(lo, hi) % 3 == 0).to(torch.uint8)
    flag = (torch.arange(lo, hi) % 5).to(torch.int32)
    packed = sharding.pack_results(reward, done, flag)
    out = sharding.gather_results(packed, world)
    r, d, f = sharding.unpack_results(out)
    q.put((rank, lo, hi, r.numpy(), d.numpy(), f.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_the_batch():
    for n, w in ((4096, 1), (4096, 8), (10, 3), (7, 8)):
        spans = [sharding.shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_env_starts_do_not_depend_on_the_number_of_shards():
    whole = sharding.env_starts(12, 720, 0, 64)
    parts = [sharding.env_starts(12, 720, *sharding.shard_bounds(64, r, 4)) for r in range(4)]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), whole[0])
    assert np.array_equal(np.concatenate([p[1] for p in parts]), whole[1])


def test_gather_of_packed_results_world_size_2():
    world, n_total = 2, 12
    port = _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = np.arange(n_total, dtype=np.float64)
    for rank, lo, hi, r, d, f in res:
        assert np.array_equal(r, np.stack([g * k for k in range(1, 6)], axis=1))
        assert np.array_equal(d, (np.arange(n_total) % 3 == 0).astype(np.uint8))
        assert np.array_equal(f, (np.arange(n_total) % 5).astype(np.int32))


def test_round_robin_shards_cover_the_global_batch_once():
    ids = np.concatenate([sharding.strided_env_ids(16, r, 4) for r in range(4)])
    assert sorted(ids.tolist()) == list(range(64))
    c, r = sharding.env_starts_of(12, 720, sharding.strided_env_ids(16, 1, 4))
    c0, r0 = sharding.env_starts(12, 720, 0, 64)
    assert np.array_equal(c, c0[1::4]) and np.array_equal(r, r0[1::4])


def _pipeline_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    B = 5
    pg = sharding.PipelinedGather(B, world)
    seen = []
    for t in range(7):
        pg.buffer().copy_(torch.full((B, sharding.PACK_WIDTH), float(100 * t + rank), dtype=torch.float64))   # "step t"
        prev = pg.launch()
        if prev is not None:
            seen.append(prev.clone())
    seen.append(pg.wait_all().clone())
    q.put((rank, [s.numpy() for s in seen]))
    dist.barrier()
    dist.destroy_process_group()


def test_pipelined_gather_delivers_every_step_in_order():
    """The double-buffered, asynchronous all-gather of bench.py's sharded loop: every step's rows of every rank arrive,
    in step order, and no buffer is overwritten while its gather is in flight."""
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pipeline_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, seen in res:
        assert len(seen) == 7
        for t, g in enumerate(seen):
            assert g.shape == (world * 5, sharding.PACK_WIDTH)
            for r in range(world):
                assert np.all(g[5 * r:5 * (r + 1)] == 100 * t + r), (rank, t, r)
