"""Data-parallel plumbing around the hot path (the reference has none: single process, single device,
`exp/run_exp.py:22-23`).

Complexes never exchange messages (block-diagonal adjacency, `data/complex.py:148-169`), so the only collective is
the gradient all-reduce: one process per GPU, each with its own batch of complexes and its own CSR plans, the
model replicated. All gradients live in ONE flat fp32 bucket (every `param.grad` is a view into it, so backward
writes straight into the bucket) that is all-reduced with a single NCCL call over NVLink/NVSwitch — ~1.7 MB at
hidden 64 / 4 layers, i.e. latency-bound; bucketing for bandwidth would be pointless.

BatchNorm statistics stay per-rank (shard-local), exactly like running the reference on each shard; numerical
equality with a single-process run on the union batch therefore holds for `graph_norm='id'`/'ln' or in eval mode,
and that is what the multi-process tests check.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's env (RANK, WORLD_SIZE, MASTER_ADDR, MASTER_PORT).
    Returns (rank, world_size, local_rank). World size 1 without env => no process group."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kwargs = {}
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
            kwargs['device_id'] = torch.device('cuda', local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local_rank


class FlatGradBucket(object):
    """All gradients of a module in one contiguous buffer; `all_reduce()` averages it across ranks in one call."""

    def __init__(self, module: torch.nn.Module):
        self.params = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError('module has no trainable parameters')
        dev, dtype = self.params[0].device, self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=dtype, device=dev)
        off = 0
        for p in self.params:
            if p.device != dev or p.dtype != dtype:
                raise ValueError('FlatGradBucket needs all parameters on one device with one dtype')
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce(self, async_op=False):
        """Average over ranks (sum then / world_size). No-op without a process group."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return None
        world = dist.get_world_size()
        if dist.get_backend() == 'nccl':
            return dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, async_op=async_op)
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)
        if async_op:
            work.wait()
        self.flat.div_(world)
        return None


class SymmetricGradBucket(FlatGradBucket):
    """The flat gradient bucket in SYMMETRIC memory (torch.distributed._symmetric_memory): every rank of the node maps
    every other rank's bucket, so `FlatAdam` can average the gradients and apply the update in ONE kernel over NVLink
    peer memory (`cwn_allreduce_adam_step_f32`) — no NCCL call on the step. The buffer is padded to a multiple of four
    floats (128-bit peer loads). `all_reduce()` remains available (NCCL on the same buffer) for callers that keep the
    optimizer separate. Single node only; raises if symmetric memory cannot be set up (no silent fallback)."""

    MAX_CTAS = 128

    def __init__(self, module: torch.nn.Module, group=None):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError('SymmetricGradBucket needs an initialised process group')
        import torch.distributed._symmetric_memory as symm
        self.params = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError('module has no trainable parameters')
        dev, dtype = self.params[0].device, self.params[0].dtype
        if dtype != torch.float32 or dev.type != 'cuda':
            raise ValueError('SymmetricGradBucket: CUDA float32 parameters only')
        total = sum(p.numel() for p in self.params)
        padded = (total + 3) // 4 * 4
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n_ctas = self.MAX_CTAS
        with torch.cuda.device(dev):
            self._storage = symm.empty(padded, dtype=dtype, device=dev)
            self._pad = symm.empty(self.n_ctas * self.world, dtype=torch.int32, device=dev)
            self._storage.zero_()
            self._pad.zero_()
            self.grad_handle = symm.rendezvous(self._storage, group.group_name)
            self.pad_handle = symm.rendezvous(self._pad, group.group_name)
        self.flat = self._storage  # [padded]; the tail beyond `total` stays zero
        self.numel = total
        self.error = torch.zeros(1, dtype=torch.int32, device=dev)
        off = 0
        for p in self.params:
            if p.device != dev or p.dtype != dtype:
                raise ValueError('FlatGradBucket needs all parameters on one device with one dtype')
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        torch.cuda.synchronize(dev)
        dist.barrier(group)  # every rank's pad is zero before anyone signals

    def self_test(self):
        """One fused launch on known data (learning rate 0: parameters untouched, the bucket keeps the average) checked
        against the closed form; raises if the peers did not meet or the average is wrong. Collective: every rank calls
        it. Leaves the bucket zeroed. Costs one launch (or ~2 s if a peer never arrives: the kernel gives up)."""
        from cwn_b200 import _lib
        n = self.flat.numel()
        dev = self.flat.device
        with torch.cuda.device(dev):
            base = torch.arange(n, dtype=torch.float32, device=dev).remainder_(97.0).add_(1.0)
            self.flat.copy_(base * float(self.rank + 1))
            scratch = torch.zeros(3, n, dtype=torch.float32, device=dev)
            step = torch.zeros(2, dtype=torch.int32, device=dev)
            torch.cuda.synchronize(dev)
            dist.barrier()
            _lib.check(_lib.load().cwn_allreduce_adam_step_f32(
                scratch[0].data_ptr(), self.grad_handle.buffer_ptrs_dev, self.pad_handle.buffer_ptrs_dev, self.rank,
                self.world, self.n_ctas, scratch[1].data_ptr(), scratch[2].data_ptr(), n, 0.0, 0.9, 0.999, 1e-8, 0.0,
                step[0:1].data_ptr(), step[1:2].data_ptr(), 0, self.error.data_ptr(),
                torch.cuda.current_stream().cuda_stream), 'cwn_allreduce_adam_step_f32 (self test)')
            torch.cuda.synchronize(dev)
            code = int(self.error.item())
            expect = base * (sum(range(1, self.world + 1)) / self.world)
            err = float((self.flat - expect).abs().max())
            # the verdict must be the same on every rank (a rank that fell back to NCCL alone would deadlock the others)
            bad = torch.tensor([float(code != 0 or not err <= 1e-3)], device=dev)
            dist.all_reduce(bad, op=dist.ReduceOp.MAX)
            self.flat.zero_()
            self.error.zero_()
            torch.cuda.synchronize(dev)
            dist.barrier()
        if float(bad.item()) != 0.0:
            raise RuntimeError(f'cwn_b200: fused all-reduce self test failed on some rank (here: barrier code {code}, '
                               f'average off by {err})')

    def check(self):
        """Raise if a fused step ever gave up waiting for a peer (host sync: call it outside the hot loop)."""
        code = int(self.error.item())
        if code:
            raise RuntimeError(f'cwn_b200: peer barrier of the fused all-reduce + Adam timed out (code {code})')


def broadcast_parameters(module: torch.nn.Module, src: int = 0):
    """Make every rank start from rank `src`'s weights and buffers."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src)


def shard(items, rank: int, world: int):
    """Contiguous shard of a list of complexes for `rank` (independent units, no halo)."""
    per = (len(items) + world - 1) // world
    return items[rank * per:(rank + 1) * per]


def train_step(model, batch, loss_fn, bucket: FlatGradBucket, optimizer=None):
    """One data-parallel step: zero bucket -> forward -> loss -> backward -> all-reduce -> optimizer.
    DDP-aware twin of the inner loop of the reference's `train()` (`exp/train_utils.py:57-72`), minus its
    per-step `loss.item()` host sync."""
    bucket.zero()
    out = model(batch)
    loss = loss_fn(out, batch.y)
    loss.backward()
    bucket.all_reduce()
    if optimizer is not None:
        optimizer.step()
    return loss
