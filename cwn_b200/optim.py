"""`FlatAdam`: torch.optim.Adam semantics (the reference's optimizer, `exp/run_exp.py:343`) as ONE kernel launch over
flat buffers. Parameters are re-homed into one contiguous buffer (they stay ordinary `nn.Parameter`s, now views),
gradients are the `FlatGradBucket`'s flat buffer, the moments are flat too; the step counter is a device scalar so
the update can be captured in a CUDA graph, and the kernel zeroes the gradients it consumed (no separate memset)."""
import torch

from cwn_b200 import _lib, ops
from cwn_b200.dist import FlatGradBucket, SymmetricGradBucket


class FlatAdam(object):
    def __init__(self, module: torch.nn.Module, bucket: FlatGradBucket = None, lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
                 weight_decay=0.0, zero_grad=True):
        self.bucket = bucket if bucket is not None else FlatGradBucket(module)
        params = self.bucket.params
        if not self.bucket.flat.is_cuda:
            raise RuntimeError('cwn_b200: FlatAdam is CUDA-only')
        self.lr, self.betas, self.eps, self.weight_decay, self.zero_grad = lr, betas, eps, weight_decay, zero_grad
        dev = self.bucket.flat.device
        self.flat_param = torch.empty_like(self.bucket.flat)
        off = 0
        for p in params:  # same order / offsets as the gradient bucket
            n = p.numel()
            self.flat_param[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + n].view_as(p)
            off += n
        self.exp_avg = torch.zeros_like(self.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.flat_param)
        self._step = torch.zeros(1, dtype=torch.int32, device=dev)
        self._counter = torch.zeros(1, dtype=torch.int32, device=dev)
        # with a SymmetricGradBucket the step ALSO averages the gradients over the ranks (one kernel over NVLink peer
        # memory): callers must not all-reduce the bucket themselves (CapturedStep checks this flag)
        self.fuses_allreduce = isinstance(self.bucket, SymmetricGradBucket) and self.bucket.world > 1
        if isinstance(self.bucket, SymmetricGradBucket):
            self.flat_param[self.bucket.numel:].zero_()  # padding

    @property
    def num_steps(self):
        return int(self._step.item())

    def step(self):
        lib = _lib.load()
        n = self.flat_param.numel()
        if self.fuses_allreduce:
            b = self.bucket
            with torch.cuda.device(self.flat_param.device):
                ops._call('allreduce_adam_step', 4 * (7 + 2 * b.world) * n, lib.cwn_allreduce_adam_step_f32,
                          self.flat_param.data_ptr(), b.grad_handle.buffer_ptrs_dev, b.pad_handle.buffer_ptrs_dev, b.rank,
                          b.world, b.n_ctas, self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), n, float(self.lr),
                          float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay),
                          self._step.data_ptr(), self._counter.data_ptr(), int(self.zero_grad), b.error.data_ptr(),
                          torch.cuda.current_stream().cuda_stream)
            return
        with torch.cuda.device(self.flat_param.device):
            ops._call('adam_step', 4 * 7 * n, lib.cwn_adam_step_f32, self.flat_param.data_ptr(),
                      self.bucket.flat.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), n,
                      float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                      float(self.weight_decay), self._step.data_ptr(), self._counter.data_ptr(),
                      int(self.zero_grad), torch.cuda.current_stream().cuda_stream)

    def zero_grad_(self):
        self.bucket.zero()
