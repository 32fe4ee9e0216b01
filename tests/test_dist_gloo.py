"""world_size-2 gloo test (CPU) of the multi-GPU host logic: disjoint scan shards, barrier, max/sum over ranks."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ndnet_b200.dist import barrier, max_over_ranks, scan_seeds, shard_range, sum_over_ranks


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = list(shard_range(13, rank, world))
        seeds = list(scan_seeds(rank, 4, 1))
        barrier()
        slowest = max_over_ranks(10.0 + rank)           # rank-dependent "step time"
        total = sum_over_ranks(len(mine))
        gathered = [None] * world
        dist.all_gather_object(gathered, (mine, seeds))
        if rank == 0:
            out.put((slowest, total, gathered))
    finally:
        dist.destroy_process_group()


def test_two_ranks_shard_and_reduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    slowest, total, gathered = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert slowest == 11.0 and total == 13
    shards = [g[0] for g in gathered]
    assert sorted(shards[0] + shards[1]) == list(range(13)) and abs(len(shards[0]) - len(shards[1])) <= 1
    assert not set(gathered[0][1]) & set(gathered[1][1])


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 64, 257):
        for world in (1, 2, 3, 8):
            items = [i for r in range(world) for i in shard_range(n, r, world)]
            assert items == list(range(n))
