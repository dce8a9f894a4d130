"""CPU (gloo, world_size 2): the multi-GPU host logic -- world sharding, barrier, max-over-ranks timing
and whole-job aggregation -- without GPUs."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from odeb200 import scenes, sharding


def test_shard_range_partitions_exactly():
    for n in (1, 7, 8, 8192, 8191):
        for ws in (1, 2, 3, 4, 8):
            covered = []
            for r in range(ws):
                f, c = sharding.shard_range(n, r, ws)
                covered += list(range(f, f + c))
            assert covered == list(range(n))
            sizes = [sharding.shard_range(n, r, ws)[1] for r in range(ws)]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, n_worlds, q):
    import zlib
    import torch.distributed as dist
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world_size), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, lr, ws = sharding.init_process_group("gloo")
    assert (r, ws) == (rank, world_size)
    first, count = sharding.shard_range(n_worlds, rank, world_size)
    sc = scenes.batched_worlds_scene(count, seed=4, first_world=first)
    crc = zlib.crc32(sc["bodies"]["pos"].tobytes()) ^ zlib.crc32(sc["geoms"]["dims"][1:].tobytes())
    sharding.barrier()
    t_max = sharding.all_reduce_max(10.0 + rank)            # slowest rank defines the step time
    total = sharding.all_reduce_sum(len(sc["bodies"]["pos"]))  # whole-job body count
    q.put((rank, first, count, crc, t_max, total))
    dist.destroy_process_group()


def test_two_rank_world_sharding_gloo():
    n_worlds, ws = 6, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, n_worlds, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(ws))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    import zlib
    full = scenes.batched_worlds_scene(n_worlds, seed=4)
    for rank, first, count, crc, t_max, total in res:
        pos = full["bodies"]["pos"][first * 128:(first + count) * 128]
        dims = full["geoms"]["dims"][1 + first * 128:1 + (first + count) * 128]
        assert crc == zlib.crc32(np.ascontiguousarray(pos).tobytes()) ^ zlib.crc32(np.ascontiguousarray(dims).tobytes())
        assert t_max == 11.0 and total == n_worlds * 128
