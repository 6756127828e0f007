"""Whole-step CUDA-graph capture of the hot path.

At the reference's real-data shape (128 molecules, ~6.5k cells) one adjacency pass moves 1-2 MB: the training step
is bound by launch latency and host-side dispatch, not by HBM ("two regimes"). The B200 answer is to
take the host out of the loop: the CSR plan builds, every fused message-passing kernel, the dense update nets, the
loss, the whole backward pass and the optimizer are captured ONCE into a CUDA graph over static, packed batch
buffers; a training step is then `load_packed_` (one copy per dtype into the static buffers, from pinned host
memory or from HBM) + one graph launch.

A graph is specific to a packed layout (`Complex.packed_signature`: the cell and message counts of every
dimension). Batches of identically shaped complexes (the synthetic benchmark) share one graph; `run()` refuses a
batch with another layout (`load_packed_` raises). Real (ragged) batches are padded to one fixed-capacity layout by
`cwn_b200.bucketed.BucketedStep`, which replays them through a single graph as well.
"""
import copy

import torch

from cwn_b200 import ops


def _index_tensors(batch):
    out = []
    for d in range(batch.dimension + 1):
        c = batch.cochains[d]
        out += [c.upper_index, c.lower_index, c.boundary_index, c.batch, c.shared_coboundaries, c.shared_boundaries]
    return [t for t in out if t is not None]


class CapturedStep(object):
    """forward -> loss -> backward (-> optimizer) on batches of one packed layout, as a single CUDA graph.

    Args:
        model, loss_fn(out, y): the training closure pieces.
        bucket: `cwn_b200.dist.FlatGradBucket` (gradients are written into its static flat buffer).
        optimizer: captured into the graph when `optimizer_in_graph` (single GPU; needs `capturable=True`);
            with data parallelism either `allreduce_in_graph=True` (the NCCL all-reduce is captured too: one graph), or
            the all-reduce runs between the graph and a (separately captured) optimizer.
    """

    def __init__(self, model, loss_fn, bucket, optimizer=None, optimizer_in_graph=True, warmup=3,
                 allreduce_in_graph=False):
        self.model, self.loss_fn, self.bucket, self.optimizer = model, loss_fn, bucket, optimizer
        self.optimizer_in_graph = optimizer_in_graph and optimizer is not None
        self.warmup = warmup
        # data parallel: the NCCL all-reduce of the flat bucket is captured INTO the graph, between backward and the
        # optimizer, so a step stays one graph launch (three host-issued pieces cost +28 / +40 / +50 us per step at
        # 2 / 4 / 8 GPUs in round 1). Needs optimizer_in_graph for the optimizer to follow it inside the graph.
        self.allreduce_in_graph = allreduce_in_graph
        # (Issuing it layer by layer during backward — async NCCL works from tensor hooks, captured as a parallel branch —
        # was measured at 2 GPUs: 0.8675 vs 0.8712 ms per step, i.e. 0.4 %, and the captured async works kept
        # destroy_process_group() from returning. Not kept: profiles/README.md, round 2.)
        self.graph = self.opt_graph = self.static = self.loss = None

    def _optimizer_clears_grads(self):
        """True only for an in-graph `FlatAdam(zero_grad=True)`, whose kernel zeroes the gradients it consumed. Every
        other optimizer (torch.optim.*: `zero_grad` is a bound method there, always truthy), a FlatAdam that runs in its
        own graph after the all-reduce, or no optimizer at all leaves the flat bucket as it is — autograd ACCUMULATES
        into the `.grad` views, so the bucket must be cleared at the top of every replayed step."""
        from cwn_b200.optim import FlatAdam
        return isinstance(self.optimizer, FlatAdam) and bool(self.optimizer.zero_grad)

    def _forward_loss(self, b):
        return self.loss_fn(self.model(b), b.y)

    def _body(self):
        b = self.static
        for d, x in enumerate(self._inputs):  # the forward overwrites cochain.x with hidden features (set_xs)
            b.cochains[d]._x = x
        ops.clear_plan_cache(*self._indices)  # plans are part of the step: every step is a new batch
        if not self._optimizer_clears_grads():
            self.bucket.zero()
        loss = self._forward_loss(b)
        loss.backward()
        if self.allreduce_in_graph and not getattr(self.optimizer, 'fuses_allreduce', False):
            self.bucket.all_reduce()  # (a FlatAdam over a SymmetricGradBucket averages inside its own kernel)
        if self.optimizer_in_graph:
            self.optimizer.step()
        return loss

    # ---- the warm-up passes run REAL steps (lazy initialisations must happen before capture): everything they touch is
    #      put back afterwards, in place (the graph has captured the addresses), so that step 1 starts from the caller's state
    def _snapshot(self):
        snap = {'model': {k: v.detach().clone() for k, v in self.model.state_dict().items()}} \
            if hasattr(self.model, 'state_dict') else {'model': {}}
        opt = self.optimizer
        if opt is None:
            return snap
        from cwn_b200.optim import FlatAdam
        if isinstance(opt, FlatAdam):
            snap['flat'] = [t.clone() for t in (opt.flat_param, opt.exp_avg, opt.exp_avg_sq, opt._step)]
        elif hasattr(opt, 'state_dict'):
            snap['opt'] = copy.deepcopy(opt.state_dict()['state'])
        return snap

    def _restore(self, snap):
        with torch.no_grad():
            opt = self.optimizer
            if 'flat' in snap:
                for t, saved in zip((opt.flat_param, opt.exp_avg, opt.exp_avg_sq, opt._step), snap['flat']):
                    t.copy_(saved)
            elif 'opt' in snap:
                index = {}
                for gi, group in enumerate(opt.param_groups):
                    for p in group['params']:
                        index[id(p)] = len(index)
                for p, st in opt.state.items():
                    before = snap['opt'].get(index.get(id(p)), {})
                    for k, v in st.items():
                        if torch.is_tensor(v):
                            v.copy_(before[k]) if k in before and torch.is_tensor(before[k]) else v.zero_()
            live = self.model.state_dict() if snap['model'] else {}
            for k, saved in snap['model'].items():
                live[k].copy_(saved)

    def capture(self, example):
        """`example`: a packed batch ON THE DEVICE; it becomes the graph's static input (do not reuse it)."""
        if example.packed_signature is None:
            raise ValueError('CapturedStep.capture needs a packed batch (ComplexBatch.pack_() / .to(device))')
        self.static = example
        self._inputs = [example.cochains[d].x for d in range(example.dimension + 1)]
        self._indices = _index_tensors(example)
        snap = self._snapshot()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # lazy initialisations (cuBLAS workspaces, optimizer state) before capture
            for _ in range(self.warmup):
                self._body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        # thread_local: other threads (NCCL's watchdog under data parallelism) keep making CUDA calls during capture
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode='thread_local'):
            self.loss = self._body()
        if self.optimizer is not None and not self.optimizer_in_graph:
            self.opt_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.opt_graph, capture_error_mode='thread_local'):
                self.optimizer.step()
        self._restore(snap)
        self.bucket.zero()  # gradients accumulated by the warm-up / capture passes must not leak into step 1
        return self

    @property
    def signature(self):
        return None if self.static is None else self.static.packed_signature

    def run(self, batch=None):
        """Load `batch` (packed, same layout; host-pinned or device) into the static buffers and replay.
        Returns the static loss tensor (valid until the next run)."""
        if batch is not None:
            self.static.load_packed_(batch)
        self.graph.replay()
        if self.opt_graph is not None:
            self.bucket.all_reduce()
            self.opt_graph.replay()
        return self.loss
