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


# ------------------------------------------------------------------------------- the real model on the data plane
def _cwn_worker(rank, world, port, out, backend, use_cuda):
    """Every rank: EmbedSparseCIN (graph_norm 'id': BatchNorm statistics are shard-local by design, so equality with a
    single process on the union holds without it) on its shard of 8 molecules -> FlatGradBucket -> all-reduce."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, 'oracle'), os.path.join(root, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    if use_cuda:
        torch.cuda.set_device(rank)
    dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        from cwn_b200.data import synthetic
        from cwn_b200.data.complex import ComplexBatch
        from cwn_b200.mp.molec_models import EmbedSparseCIN
        if not use_cuda:  # TEST INFRASTRUCTURE: torch restatements of the six ops entry points (the product is CUDA-only)
            import cpu_ops_shim
            from cwn_b200 import ops
            for name in ('gather_scatter', 'gather_rows', 'scatter_rows', 'cob_pass', 'segment_pool', 'prepare_plans'):
                setattr(ops, name, getattr(cpu_ops_shim, name))
        dev = torch.device('cuda', rank) if use_cuda else torch.device('cpu')
        torch.manual_seed(0)
        model = EmbedSparseCIN(**_CWN_CFG).to(dev).train()
        broadcast_parameters(model, src=0)
        bucket = FlatGradBucket(model)
        comps = shard(synthetic.zinc_like_complexes(8, seed=11), rank, world)
        batch = ComplexBatch.from_complex_list(comps).to(dev)
        loss = train_step(model, batch, lambda o, y: torch.nn.functional.l1_loss(o, y.view(-1, 1)), bucket)
        out[rank] = (bucket.flat.cpu().clone(), float(loss))
    finally:
        dist.destroy_process_group()


_CWN_CFG = dict(atom_types=28, bond_types=4, out_size=1, num_layers=2, hidden=16, dropout_rate=0.0, max_dim=2,
                embed_edge=True, use_coboundaries=True, graph_norm='id', nonlinearity='elu')


def _union_reference():
    """Mean of the per-shard gradients == gradient of the mean of the shard losses (equal shard sizes), from the
    CPU oracle-equivalent host path in ONE process."""
    import cpu_ops_shim
    from cwn_b200 import ops
    from cwn_b200.data import synthetic
    from cwn_b200.data.complex import ComplexBatch
    from cwn_b200.mp.molec_models import EmbedSparseCIN
    saved = {n: getattr(ops, n) for n in ('gather_scatter', 'gather_rows', 'scatter_rows', 'cob_pass', 'segment_pool', 'prepare_plans')}
    try:
        for n in saved:
            setattr(ops, n, getattr(cpu_ops_shim, n))
        torch.manual_seed(0)
        model = EmbedSparseCIN(**_CWN_CFG).train()
        comps = synthetic.zinc_like_complexes(8, seed=11)
        total = 0.0
        for r in range(2):
            b = ComplexBatch.from_complex_list(shard(comps, r, 2))
            total = total + torch.nn.functional.l1_loss(model(b), b.y.view(-1, 1)) / 2
        total.backward()
        return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1)
                          for p in model.parameters() if p.requires_grad])
    finally:
        for n, f in saved.items():
            setattr(ops, n, f)


def test_cwn_model_gradients_average_over_ranks_gloo():
    """World 2 over gloo with the cochain model itself on the data plane: the all-reduced flat bucket of every rank ==
    the mean of the per-shard gradients computed in one process."""
    world, port = 2, _free_port()
    out = mp.Manager().dict()
    mp.spawn(_cwn_worker, args=(world, port, out, 'gloo', False), nprocs=world, join=True)
    ref = _union_reference()
    for r in range(world):
        assert torch.allclose(out[r][0], ref, rtol=1e-5, atol=1e-6), float((out[r][0] - ref).abs().max())
    assert torch.equal(out[0][0], out[1][0])
