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

