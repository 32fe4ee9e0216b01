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


def _grad_worker(rank, world, port, out):
    from ndnet_b200.train import allreduce_gradients
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.BatchNorm1d(7), torch.nn.Linear(7, 3))
        net[2].bias.requires_grad_(False)                       # parameters without a gradient are skipped
        x = torch.full((4, 5), float(rank + 1))
        (net(x) ** 2).sum().backward()
        local = [p.grad.clone() for p in net.parameters() if p.grad is not None]
        n = allreduce_gradients(net)
        gathered = [None] * world
        dist.all_gather_object(gathered, [g.numpy() for g in local])
        if rank == 0:
            out.put((n, [p.grad.numpy() for p in net.parameters() if p.grad is not None], gathered))
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_averages_over_two_ranks():
    """The training path's only collective: one flat all-reduce, mean over ranks."""
    import numpy as np
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    n, reduced, gathered = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert n == sum(g.size for g in reduced)
    for k, g in enumerate(reduced):
        assert np.allclose(g, (gathered[0][k] + gathered[1][k]) / 2, rtol=1e-6, atol=1e-7)


def _bn_worker(rank, world, port, out):
    from ndnet_b200.train import allreduce_gradients, sync_batchnorm_buffers
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.BatchNorm1d(7), torch.nn.Linear(7, 3)).train()
        for _ in range(rank + 1):                                  # the ranks see different data and step counts
            net(torch.randn(6, 5) * (rank + 1) + rank)
        before = [b.clone() for b in net.buffers()]
        n = sync_batchnorm_buffers(net)
        # a rank whose last layer produced no gradient still takes part in the same collective (missing = zeros)
        if rank == 0:
            (net[0](torch.ones(2, 5)) ** 2).sum().backward()       # only layer 0 has gradients on rank 0
        else:
            (net(torch.ones(4, 5)) ** 2).sum().backward()
        had = [p.grad is not None for p in net.parameters()]
        local = [p.grad.clone() if p.grad is not None else torch.zeros_like(p) for p in net.parameters()]
        m = allreduce_gradients(net)
        gathered = [None] * world
        dist.all_gather_object(gathered, ([b.numpy() for b in before], [g.numpy() for g in local], had))
        if rank == 0:
            out.put((n, m, [b.numpy() for b in net.buffers()], [p.grad.numpy() for p in net.parameters()], gathered))
    finally:
        dist.destroy_process_group()


def test_batchnorm_buffers_are_averaged_and_missing_gradients_count_as_zeros():
    import numpy as np
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bn_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    n, m, buffers, grads, gathered = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    b0, b1 = gathered[0][0], gathered[1][0]
    assert n == sum(b.size for b in buffers)
    for k, b in enumerate(buffers):
        if b.dtype.kind == "f":
            assert np.allclose(b, (b0[k] + b1[k]) / 2, rtol=1e-6, atol=1e-7)       # running_mean / running_var: mean over ranks
        else:
            assert b == max(b0[k], b1[k])                                          # num_batches_tracked: the larger count
    assert gathered[0][2] != gathered[1][2]                                         # rank 0 really lacked some gradients
    assert m == sum(g.size for g in grads)
    for k, g in enumerate(grads):
        assert np.allclose(g, (gathered[0][1][k] + gathered[1][1][k]) / 2, rtol=1e-6, atol=1e-7)
