"""The N>1 data-parallel host logic on CPU: world_size=2, gloo (flat arena, one all-reduce, identical update)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _sgd_reference(p, g, m, lr, mu, wd, gscale, first):
    gg = g * gscale + wd * p
    m.copy_(gg if first else mu * m + gg)
    p.sub_(lr * m)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from deeplearning_b200.classification.mnist.models.network import mnist_fcn
        from deeplearning_b200.engine.trainer import FlatArena

        torch.manual_seed(100 + rank)  # deliberately different init per rank: broadcast must fix it
        model = mnist_fcn(10)
        arena = FlatArena(model.parameters())
        arena.broadcast(model.buffers())
        assert all(p.data_ptr() == arena.flat_p[o:].data_ptr() for p, o in zip(arena.params, arena.offsets))
        g = torch.Generator().manual_seed(7)
        x = torch.randn(8, 3, 28, 28, generator=g)
        y = torch.randint(0, 10, (8,), generator=g)
        xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]
        for step in range(2):
            arena.flat_g.zero_()
            loss = F.cross_entropy(model(xs), ys)
            grads = torch.autograd.grad(loss, arena.params)
            for p, gr in zip(arena.params, grads):
                arena.grad_view(p).copy_(gr)
            arena.all_reduce_grads()
            _sgd_reference(arena.flat_p, arena.flat_g, arena.flat_m, 0.1, 0.9, 5e-5, arena.grad_scale, step == 0)
        out[rank] = (arena.flat_p.clone(), arena.flat_g.clone() * arena.grad_scale)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_equals_single_process_full_batch():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    p0, g0 = out[0]
    p1, g1 = out[1]
    assert torch.equal(p0, p1) and torch.equal(g0, g1)  # replicas stay bit-identical
    # single process, full batch, same init as rank 0 (BN-free model => mean-loss gradient == average of shard gradients)
    from deeplearning_b200.classification.mnist.models.network import mnist_fcn

    torch.manual_seed(100)
    model = mnist_fcn(10)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(8, 3, 28, 28, generator=g)
    y = torch.randint(0, 10, (8,), generator=g)
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-5)
    for _ in range(2):
        opt.zero_grad()
        F.cross_entropy(model(x), y).backward()
        opt.step()
    flat = torch.cat([torch.nn.functional.pad(p.detach().reshape(-1), (0, (-p.numel()) % 4)) for p in model.parameters()])
    assert torch.allclose(flat, p0, rtol=1e-5, atol=1e-6)


def _overlap_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from deeplearning_b200.classification.mnist.models.network import mnist_fcn
        from deeplearning_b200.engine.trainer import FlatArena

        torch.manual_seed(5)
        model = mnist_fcn(10)
        arena = FlatArena(model.parameters(), bucket_mb=0.02)     # ~5000-element buckets: several per backward
        g = torch.Generator().manual_seed(40 + rank)
        local = torch.randn(arena.flat_g.numel(), generator=g)
        # (a) gradients become final in reverse parameter order (what the engines do), (b) in a scrambled order
        for order in (list(reversed(arena.params)), [arena.params[i] for i in torch.randperm(len(arena.params), generator=torch.Generator().manual_seed(3)).tolist()]):
            arena.flat_g.copy_(local)
            arena.begin_backward()
            for p in order:
                arena.notify(p)
            launched = arena.buckets_launched
            arena.finish_backward()
            out[(rank, len(out))] = (arena.flat_g.clone(), launched, arena.buckets_launched)
    finally:
        dist.destroy_process_group()


def test_overlapped_bucket_allreduce_equals_one_allreduce():
    """FlatArena.begin_backward / notify / finish_backward (the bucketed reduce the GPU step overlaps with its backward pass)
    sums exactly what a single all-reduce of the arena sums, whatever order the gradients complete in."""
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_overlap_worker, args=(world, port, out), nprocs=world, join=True)
    vals = list(out.values())
    n = vals[0][0].numel()
    expect = torch.randn(n, generator=torch.Generator().manual_seed(40)) + torch.randn(n, generator=torch.Generator().manual_seed(41))
    for flat, launched_mid, launched_all in vals:
        assert torch.allclose(flat, expect, rtol=0, atol=1e-6)
        assert launched_all >= launched_mid >= 1
    # in reverse parameter order the buckets go out while "the backward" is still running, not all at the end
    assert max(v[1] for v in vals) >= 2
