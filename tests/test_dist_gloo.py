"""Data-parallel plumbing on CPU with the gloo backend, world_size 2 (the N>1 path of bench.py without GPUs).
The message-passing model itself is CUDA-only, so the plumbing is exercised with a plain torch module: what is
checked is the flat gradient bucket, the single averaged all-reduce, parameter broadcast and sharding."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cwn_b200.dist import FlatGradBucket, broadcast_parameters, shard, train_step


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _net():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.Tanh(), torch.nn.Linear(8, 1))


class _Batch(object):
    def __init__(self, x, y):
        self.x, self.y = x, y


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        model = _net()
        if rank == 1:  # diverge on purpose: broadcast must repair it
            with torch.no_grad():
                for p in model.parameters():
                    p.add_(1.0)
        broadcast_parameters(model, src=0)
        bucket = FlatGradBucket(model)
        g = torch.Generator().manual_seed(1)
        x, y = torch.randn(8, 6, generator=g), torch.randn(8, generator=g)
        idx = shard(list(range(8)), rank, world)
        loss = train_step(lambda b: model(b.x), _Batch(x[idx], y[idx]),
                          lambda o, t: torch.nn.functional.mse_loss(o.view(-1), t), bucket)
        out[rank] = (bucket.flat.clone(), [p.detach().clone() for p in model.parameters()], float(loss))
    finally:
        dist.destroy_process_group()


def test_flat_bucket_all_reduce_equals_single_process_gradient():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    # single process on the union batch (equal shard sizes: mean of shard-mean losses == mean loss)
    model = _net()
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(8, 6, generator=g), torch.randn(8, generator=g)
    torch.nn.functional.mse_loss(model(x).view(-1), y).backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    for r in range(world):
        flat, params, _ = out[r]
        assert torch.allclose(flat, ref, rtol=1e-5, atol=1e-6)
        for p, q in zip(params, model.parameters()):
            assert torch.equal(p, q.detach())  # rank 1 was re-synchronised to rank 0's weights
    assert torch.equal(out[0][0], out[1][0])


def test_bucket_views_and_shard():
    model = _net()
    bucket = FlatGradBucket(model)
    assert bucket.flat.numel() == sum(p.numel() for p in model.parameters())
    model(torch.ones(2, 6)).sum().backward()
    assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in model.parameters())  # grads live in the bucket
    assert float(bucket.flat.abs().sum()) > 0
    bucket.zero()
    assert all(float(p.grad.abs().sum()) == 0 for p in model.parameters())
    assert bucket.all_reduce() is None  # no process group: no-op
    assert shard(list(range(10)), 0, 4) == [0, 1, 2] and shard(list(range(10)), 3, 4) == [9]
