"""Ragged batches at CUDA-graph speed: pad every batch to ONE fixed-capacity layout and replay a single graph.

The reference's training loop (`exp/train_utils.py:35-72`, `data/data_loading.py:44-82`) collates a different set of
molecules every step, so cell and message counts change from batch to batch and a graph captured for one packed layout
(`cwn_b200.graph.CapturedStep`) cannot be replayed. Here a batch of B complexes is completed with DUMMY complexes up to
fixed capacities (cells per dimension, messages per adjacency, complexes):

  * the dummies are ordinary complexes (zero features, label 0) appended after the real ones, so the real cells keep
    the leading rows of every matrix; all padding messages run from a dummy cell to a dummy cell — the block-diagonal
    structure that makes complexes independent (`data/complex.py:148-169`) keeps them away from the real cells (they
    are spread evenly over the dummy's cells: no long CSR row);
  * BatchNorm is the one place where rows of different complexes meet: the dense kernels read the number of LIVE rows
    from a device scalar that travels with the batch (`cwn_linear_desc::n_rows_live`), so batch statistics, running
    statistics and the BatchNorm backward see the real cells only;
  * the loss is a weighted mean with weight 0 on the dummy complexes, so no gradient ever enters a padding row (and the
    backward kernels zero g_z of padding rows, which would otherwise pick up the BatchNorm mean terms).

Results on the real complexes are those of the unpadded batch (same kernels, same summation order inside a row; the
BatchNorm tile partials are merged over the same 64-row tiles) — `tests/test_gpu_parity.py::test_padded_batch_*`.
A batch that does not fit the capacities falls back to the eager path.
"""
import math

import torch

from cwn_b200 import fused
from cwn_b200.data.complex import Cochain, Complex, ComplexBatch
from cwn_b200.graph import CapturedStep


class Overflow(ValueError):
    """The batch has more cells / messages / complexes than the padded layout holds."""


def _counts(comp):
    """(cells per dim [3], upper messages per dim [2], boundary messages per dim 1..2 [2]) of one complex."""
    cells, up, bnd = [0, 0, 0], [0, 0], [0, 0]
    for d in range(min(comp.dimension, 2) + 1):
        c = comp.cochains[d]
        if c.lower_index is not None:
            raise ValueError('padded batches support upper / boundary adjacencies only (the SparseCIN family)')
        cells[d] = int(c.num_cells or 0)
        if d < 2 and c.upper_index is not None:
            up[d] = int(c.upper_index.size(1))
        if d > 0 and c.boundary_index is not None:
            bnd[d - 1] = int(c.boundary_index.size(1))
    return cells, up, bnd


class Capacity(object):
    """Fixed sizes of the padded layout. `complexes` = real slots (one more is always taken by a dummy)."""

    def __init__(self, cells, up, bnd, complexes):
        self.cells, self.up, self.bnd, self.complexes = list(cells), list(up), list(bnd), int(complexes)

    @classmethod
    def from_dataset(cls, complexes, batch_size, sigmas=5.0, cell_multiple=64, message_multiple=256):
        """Capacities that hold a random batch of `batch_size` complexes of this dataset with overwhelming probability:
        batch_size * mean + sigmas * sqrt(batch_size) * std of every count (the sums of a shuffled batch concentrate),
        rounded up to the kernels' tile sizes. The slack costs nothing at this size: the step is latency-bound."""
        rows = [sum(_counts(c), []) for c in complexes]
        t = torch.tensor(rows, dtype=torch.float64)
        mean, std = t.mean(0), t.std(0, unbiased=False)
        need = mean * batch_size + sigmas * std * math.sqrt(batch_size) + 1  # (+1: the dummy's own cell)
        worst = t.max(0).values  # a batch of one complex must fit as well
        need = torch.maximum(need, worst + 1)

        def up_to(v, m):
            return int(math.ceil(float(v) / m) * m)
        n = [up_to(v, cell_multiple) for v in need[:3]]
        m = [up_to(v, message_multiple) for v in need[3:]]
        return cls(n, m[:2], m[2:], batch_size)

    def as_tuple(self):
        return tuple(self.cells), tuple(self.up), tuple(self.bnd), self.complexes

    def __repr__(self):
        return f'Capacity(cells={self.cells}, up={self.up}, bnd={self.bnd}, complexes={self.complexes})'


def _dummy(template, cells, up, bnd):
    """A complex with exactly these counts: zero features shaped like `template`'s; padding message i runs from dummy
    cell i mod n_src to dummy cell i mod n_dst (through dummy coboundary i mod n_up). Spread out on purpose: thousands
    of messages into ONE destination row would be one lane group walking one CSR row sequentially (measured: a padded
    step 2.8x slower than the unpadded one when every padding message targeted cell 0)."""
    def zeros_like_rows(x, n):
        return None if x is None else x.new_zeros((n,) + tuple(x.shape[1:]))

    def spread_index(n, n_src, n_dst):
        if n <= 0:
            return None
        i = torch.arange(n, dtype=torch.long)
        return torch.stack([i % n_src, i % n_dst])

    def spread(n, m):
        return torch.arange(n, dtype=torch.long) % m if n > 0 else None
    tv, te = template.cochains[0], template.cochains.get(1)
    v = Cochain(dim=0, x=zeros_like_rows(tv.x, cells[0]), upper_index=spread_index(up[0], cells[0], cells[0]),
                shared_coboundaries=spread(up[0], cells[1]), num_cells=cells[0], num_cells_up=cells[1])
    e = Cochain(dim=1, x=zeros_like_rows(te.x if te is not None else None, cells[1]),
                boundary_index=spread_index(bnd[0], cells[0], cells[1]), upper_index=spread_index(up[1], cells[1], cells[1]),
                shared_coboundaries=spread(up[1], cells[2]), num_cells=cells[1], num_cells_down=cells[0],
                num_cells_up=cells[2])
    r = Cochain(dim=2, x=None, boundary_index=spread_index(bnd[1], cells[1], cells[2]), num_cells=cells[2],
                num_cells_down=cells[1], num_cells_up=0)
    y = None if template.y is None else torch.zeros_like(template.y)
    return Complex(v, e, r, y=y)


def pad_complexes(complexes, cap: Capacity, max_dim: int = 2):
    """`ComplexBatch` of `complexes` + dummies with exactly the capacities' shapes. Extra attributes (packed with the
    batch): `cochains[d].live_rows` int32 [1] = real cells of dimension d, `cochains[0].complex_weight` float32
    [complexes + 1] = 1 for the real complexes. Raises `Overflow` if the batch does not fit."""
    if max_dim != 2:
        raise ValueError('padded batches are built for 2-complexes (max_dim = 2)')
    n_real = len(complexes)
    if n_real < 1 or n_real > cap.complexes:
        raise Overflow(f'{n_real} complexes, capacity {cap.complexes}')
    tot_c, tot_u, tot_b = [0, 0, 0], [0, 0], [0, 0]
    template = None
    for comp in complexes:
        c, u, b = _counts(comp)
        tot_c = [a + x for a, x in zip(tot_c, c)]
        tot_u = [a + x for a, x in zip(tot_u, u)]
        tot_b = [a + x for a, x in zip(tot_b, b)]
        if template is None and comp.dimension >= 1:
            template = comp
    template = template or complexes[0]
    n_dummy = cap.complexes + 1 - n_real
    # every dummy owns one cell per dimension; the first one absorbs the rest of the slack and all padding messages
    big_c = [cap.cells[d] - tot_c[d] - (n_dummy - 1) for d in range(3)]
    big_u = [cap.up[d] - tot_u[d] for d in range(2)]
    big_b = [cap.bnd[d] - tot_b[d] for d in range(2)]
    if min(big_c) < 1 or min(big_u) < 0 or min(big_b) < 0:
        raise Overflow(f'cells {tot_c} / upper {tot_u} / boundary {tot_b} do not fit {cap}')
    dummies = [_dummy(template, big_c, big_u, big_b)] + [_dummy(template, [1, 1, 1], [0, 0], [0, 0]) for _ in range(n_dummy - 1)]
    batch = ComplexBatch.from_complex_list(list(complexes) + dummies, max_dim=max_dim)
    for d in range(3):
        # `ptr` has one entry per member that owns cells of the dimension (data/complex.py:427-432): its LENGTH varies
        # with how many molecules have rings. Nothing on the hot path reads it, so the padded batch drops it.
        batch.cochains[d].ptr = None
        batch.cochains[d].live_rows = torch.tensor([tot_c[d]], dtype=torch.int32)
    w = torch.zeros(cap.complexes + 1, dtype=torch.float32)
    w[:n_real] = 1.0
    batch.cochains[0].complex_weight = w
    return batch


def masked_l1(out, y, weight):
    """L1 loss over the real complexes = the reference's `torch.nn.L1Loss` on the unpadded batch (`exp/run_exp.py:165`)."""
    per_complex = (out - y.view_as(out)).abs().reshape(out.size(0), -1).mean(1)
    return (per_complex * weight).sum() / weight.sum()


class PaddedModel(torch.nn.Module):
    """`model(batch)` with the batch's live row counts installed for the fused dense kernels."""

    def __init__(self, model):
        super(PaddedModel, self).__init__()
        self.model = model

    def forward(self, batch):
        live = [getattr(batch.cochains[d], 'live_rows', None) for d in range(batch.dimension + 1)]
        if any(t is None for t in live):
            return self.model(batch)
        with fused.live_rows(live):
            return self.model(batch)


class BucketedStep(CapturedStep):
    """One CUDA graph for every batch of a dataset: forward -> masked loss -> backward (-> optimizer) over the padded
    layout `capacity`.

        step = BucketedStep(model, masked_l1, bucket, optimizer, capacity).capture(example_complexes)
        loss = step.run(step.pad(complexes))          # or step.step(complexes): pad, replay, eager fallback on Overflow

    `loss_fn(out, y, weight)` receives the per-complex weights (1 real / 0 dummy)."""

    def __init__(self, model, loss_fn, bucket, optimizer=None, capacity: Capacity = None, optimizer_in_graph=True,
                 warmup=3, max_dim=2):
        super(BucketedStep, self).__init__(model, loss_fn, bucket, optimizer, optimizer_in_graph, warmup)
        self.capacity, self.max_dim = capacity, max_dim
        self.padded_model = PaddedModel(model)
        self.fallbacks = 0

    def _forward_loss(self, b):
        out = self.padded_model(b)
        return self.loss_fn(out, b.y, b.cochains[0].complex_weight)

    def pad(self, complexes, pin_memory=True):
        """Padded, packed (pinned) host batch ready for `run()`."""
        return pad_complexes(complexes, self.capacity, self.max_dim).pack_(pin_memory=pin_memory)

    def capture(self, example):
        """`example`: a list of complexes (padded here) or an already padded batch; it becomes the static input."""
        if isinstance(example, (list, tuple)):
            example = self.pad(example)
        dev = self.bucket.flat.device
        if not any(True for f in getattr(example, '_flat', {}).values() if f.is_cuda):
            example = example.to(dev)
        return super(BucketedStep, self).capture(example)

    def step(self, complexes):
        """Pad + replay; a batch beyond the capacities runs eagerly on its own (unpadded) layout."""
        try:
            return self.run(self.pad(complexes))
        except Overflow:
            self.fallbacks += 1
            return self.eager(complexes)

    def eager(self, complexes):
        dev = self.bucket.flat.device
        batch = ComplexBatch.from_complex_list(list(complexes), max_dim=self.max_dim).to(dev)
        if not self._optimizer_clears_grads():
            self.bucket.zero()
        out = self.model(batch)
        w = torch.ones(out.size(0), dtype=torch.float32, device=dev)
        loss = self.loss_fn(out, batch.y, w)
        loss.backward()
        if self.optimizer is not None:
            if not self.optimizer_in_graph:
                self.bucket.all_reduce()
            self.optimizer.step()
        return loss
