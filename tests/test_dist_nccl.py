"""World-size-2 NCCL check of the data-parallel step on the real data plane (needs 2 GPUs: `gpurun --gpus 2`; skipped
otherwise): NCCL-averaged gradients of the CUDA model == the mean of the per-shard gradients (one CPU process)."""
import pytest
import torch
import torch.multiprocessing as mp

from test_dist_gloo import _cwn_worker, _free_port, _union_reference

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_cwn_model_gradients_average_over_ranks_nccl():
    world, port = 2, _free_port()
    out = mp.Manager().dict()
    mp.spawn(_cwn_worker, args=(world, port, out, 'nccl', True), nprocs=world, join=True)
    ref = _union_reference()
    for r in range(world):
        err = float((out[r][0] - ref).abs().max())
        assert torch.allclose(out[r][0], ref, rtol=1e-4, atol=1e-5), err
    assert torch.equal(out[0][0], out[1][0])



def _fused_worker(rank, world, port, out):
    """Three training steps of the CUDA model per rank, once with NCCL all-reduce + FlatAdam, once with the gradients
    in symmetric memory and the all-reduce fused into the Adam kernel over NVLink peer memory."""
    import os
    import sys
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        from test_dist_gloo import _CWN_CFG
        from cwn_b200.data import synthetic
        from cwn_b200.data.complex import ComplexBatch
        from cwn_b200.dist import FlatGradBucket, SymmetricGradBucket, broadcast_parameters, shard
        from cwn_b200.mp.molec_models import EmbedSparseCIN
        from cwn_b200.optim import FlatAdam
        dev = torch.device('cuda', rank)
        results = {}
        for kind in ('nccl', 'fused'):
            torch.manual_seed(0)
            model = EmbedSparseCIN(**_CWN_CFG).to(dev).train()
            broadcast_parameters(model, src=0)
            bucket = SymmetricGradBucket(model) if kind == 'fused' else FlatGradBucket(model)
            if kind == 'fused':
                bucket.self_test()
            opt = FlatAdam(model, bucket, lr=1e-2)
            assert opt.fuses_allreduce == (kind == 'fused')
            n = sum(p.numel() for p in bucket.params)
            for it in range(3):
                comps = shard(synthetic.zinc_like_complexes(8, seed=11 + it), rank, world)
                batch = ComplexBatch.from_complex_list(comps).to(dev)
                loss = torch.nn.functional.l1_loss(model(batch), batch.y.view(-1, 1))
                loss.backward()
                if kind == 'nccl':
                    bucket.all_reduce()
                opt.step()
            torch.cuda.synchronize()
            if kind == 'fused':
                bucket.check()
                assert float(bucket.flat.abs().max()) == 0.0  # the kernel cleared the bucket it consumed
            results[kind] = opt.flat_param[:n].cpu().clone()
        out[rank] = results
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_fused_allreduce_adam_over_peer_memory_matches_nccl_then_adam():
    """cwn_allreduce_adam_step_f32 (two-shot average over NVLink peer memory + Adam, one launch) against
    ncclAllReduce(AVG) + cwn_adam_step_f32: parameters after three steps agree to summation-order noise of the average,
    and the ranks hold bit-identical parameters (every rank receives the SAME averaged bits)."""
    world, port = 2, _free_port()
    out = mp.Manager().dict()
    mp.spawn(_fused_worker, args=(world, port, out), nprocs=world, join=True)
    assert torch.equal(out[0]['fused'], out[1]['fused'])
    for r in range(world):
        a, b = out[r]['fused'], out[r]['nccl']
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6), float((a - b).abs().max())
