"""`PackedComplexDataset`: a dataset of complexes resident in HBM as flat arrays, collated into `ComplexBatch`es BY THE
GPU (`csrc/collate.cu`) instead of by Python loops on the CPU.

Reference path being replaced: `Collater` -> `ComplexBatch.from_complex_list` -> `CochainBatch.from_cochain_list`
(`data/data_loading.py:44-82`, `data/complex.py:323-458, 690-728`): per key a Python loop over the complexes that adds
the running offsets of `Cochain.__inc__` (:148-169) and `torch.cat`s — milliseconds per batch of 128, i.e. more than
the whole training step takes here. The packed dataset keeps, per dimension, every tensor of every complex
concatenated once (at dataset-build time) plus per-complex segment pointers; collating a batch is then
  host : look up the segment sizes of the chosen ids (numpy), prefix-sum them, upload ONE small table,
  GPU  : ONE kernel that copies every segment of every tensor to its place, adding the per-segment offsets,
and the result is a *packed* `ComplexBatch` (flat buffer per dtype) with exactly the layout
`ComplexBatch.from_complex_list(...).pack_()` would have produced — so it can be `load_packed_` into the static
buffers of a captured CUDA graph, or used directly.

Supported attributes: x, upper_index, lower_index, boundary_index, shared_boundaries, shared_coboundaries (+ the
derived batch / ptr) and the complex label y. Cell-level labels and orientations are not packed.
"""
import numpy as np
import torch

from cwn_b200 import _lib, ops
from cwn_b200.data.complex import Cochain, CochainBatch, Complex, ComplexBatch

_ELEMENT_SIZE = {torch.float32: 4, torch.long: 8, torch.float64: 8, torch.int32: 4}
_INDEX_KEYS = ('upper_index', 'lower_index', 'boundary_index')
_VECTOR_KEYS = ('shared_boundaries', 'shared_coboundaries')
# the order in which ComplexBatch.pack_ meets the tensors of a cochain (Cochain.__init__ order, then batch / ptr)
_SLOT_ORDER = ('x', 'upper_index', 'lower_index', 'boundary_index', 'shared_boundaries', 'shared_coboundaries',
               'batch', 'ptr')


class PackedComplexDataset(object):
    def __init__(self, complexes, max_dim: int = 2, device='cuda'):
        self.max_dim = max_dim
        self.device = torch.device(device)
        self.n = len(complexes)
        self.dim_of = np.array([min(c.dimension, max_dim) for c in complexes], dtype=np.int64)
        top = int(self.dim_of.max())
        self.dims = list(range(top + 1))
        self.sizes, self.ptrs, self.flat = {}, {}, {}
        self.has_y = all(c.y is not None for c in complexes)
        for d in self.dims:
            cochains = [c.cochains.get(d) if d <= min(c.dimension, max_dim) else None for c in complexes]
            for c in cochains:
                if c is not None and (c.y is not None or c.upper_orient is not None or c.lower_orient is not None):
                    pass  # cell labels / orientations are simply not carried into the batch
            n_cells = np.array([0 if c is None else (c.num_cells or 0) for c in cochains], dtype=np.int64)
            self.sizes[(d, 'cells')] = n_cells
            self.ptrs[(d, 'cells')] = np.concatenate([[0], np.cumsum(n_cells)])
            xs = [c.x for c in cochains if c is not None and c.x is not None]
            if xs:
                if len(xs) != sum(1 for c in cochains if c is not None and (c.num_cells or 0) > 0):
                    raise ValueError(f'PackedComplexDataset: dimension {d} has features on some complexes only')
                x = torch.cat([t if t.dim() == 2 else t.unsqueeze(-1) for t in xs], 0)
                self.flat[(d, 'x')] = x.contiguous().to(self.device)
            for key in _INDEX_KEYS + _VECTOR_KEYS:
                items = [None if c is None else getattr(c, key) for c in cochains]
                lens = np.array([0 if t is None else t.size(-1) for t in items], dtype=np.int64)
                if lens.sum() == 0:
                    continue
                self.sizes[(d, key)] = lens
                self.ptrs[(d, key)] = np.concatenate([[0], np.cumsum(lens)])
                cat = torch.cat([t for t in items if t is not None], -1).contiguous()
                self.flat[(d, key)] = cat.to(self.device)  # [2, total] for index keys, [total] for vector keys
        if self.has_y:
            ys = [c.y.reshape(-1) for c in complexes]
            lens = np.array([t.numel() for t in ys], dtype=np.int64)
            self.sizes['y'] = lens
            self.ptrs['y'] = np.concatenate([[0], np.cumsum(lens)])
            self.flat['y'] = torch.cat(ys).contiguous().to(self.device)

    def __len__(self):
        return self.n

    # -------------------------------------------------------------------------------------------------- collate
    def collate(self, ids, out: ComplexBatch = None) -> ComplexBatch:
        """Batch of the complexes `ids` (host sequence of ints), built on the GPU. `out`: an existing packed batch of
        the SAME layout to write into (e.g. the static batch of a captured CUDA graph) instead of allocating.
        = `launch(prepare(ids), out)`; the two halves are separate so that a worker thread can prepare the next batch
        (`cwn_b200.data.data_loading.prefetched`)."""
        return self.launch(self.prepare(ids), out=out)

    def prepare(self, ids):
        """Host half of a collation: segment sizes, prefix sums, the batch layout and the segment table (numpy only, no
        CUDA call — safe on a worker thread)."""
        ids = np.asarray(ids, dtype=np.int64)
        B = len(ids)
        dimension = int(min(self.dim_of[ids].max(), self.max_dim))
        dims = list(range(dimension + 1))
        cell_cnt = {d: self.sizes[(d, 'cells')][ids] for d in dims}
        cell_off = {d: np.concatenate([[0], np.cumsum(cell_cnt[d])]) for d in dims}  # [B+1]
        zeros = np.zeros(B + 1, dtype=np.int64)

        # ---- what the batch contains: (dim, key, dtype, shape) in pack_ order, plus how to fill it
        slots = []  # (dim, key, dtype, shape, fill spec)
        for d in dims:
            n_d = int(cell_off[d][-1])
            for key in _SLOT_ORDER:
                if key == 'x':
                    if (d, 'x') in self.flat and n_d > 0:
                        x = self.flat[(d, 'x')]
                        if x.dtype == torch.float32:
                            spec = ('rows', x, self.ptrs[(d, 'cells')][ids], cell_off[d], None)
                        elif x.dtype == torch.long:  # integer feature columns (ogbg-mol*): w int64 elements per cell
                            w = x.size(1)
                            spec = ('index1', x, self.ptrs[(d, 'cells')][ids] * w, cell_off[d] * w, None)
                        else:
                            raise TypeError(f'PackedComplexDataset: unsupported feature dtype {x.dtype}')
                        slots.append((d, 'x', x.dtype, (n_d, x.size(1)), spec))
                elif key in _INDEX_KEYS or key in _VECTOR_KEYS:
                    if (d, key) not in self.flat:
                        continue
                    lens = self.sizes[(d, key)][ids]
                    total = int(lens.sum())
                    if total == 0:
                        continue
                    dst = np.concatenate([[0], np.cumsum(lens)])
                    src = self.ptrs[(d, key)][ids]
                    if key in ('upper_index', 'lower_index'):
                        adds = (cell_off[d][:-1], cell_off[d][:-1])
                    elif key == 'boundary_index':
                        adds = (cell_off[d - 1][:-1] if d > 0 else zeros[:-1], cell_off[d][:-1])
                    elif key == 'shared_boundaries':
                        adds = (cell_off[d - 1][:-1] if d > 0 else zeros[:-1],)
                    else:  # shared_coboundaries
                        adds = (cell_off[d + 1][:-1] if (d + 1) in cell_off else zeros[:-1],)
                    shape = (2, total) if key in _INDEX_KEYS else (total,)
                    slots.append((d, key, torch.long, shape, ('index', self.flat[(d, key)], src, dst, adds)))
                elif key == 'batch' and n_d > 0:
                    slots.append((d, 'batch', torch.long, (n_d,), ('fill', None, None, cell_off[d], None)))
                elif key == 'ptr' and n_d > 0:
                    present = cell_cnt[d] > 0  # complexes without this dimension contribute no ptr entry
                    ptr = np.concatenate([[0], np.cumsum(cell_cnt[d][present])])
                    slots.append((d, 'ptr', torch.long, (len(ptr),), ('table', ptr, None, None, None)))
        if self.has_y:
            y = self.flat['y']
            lens = self.sizes['y'][ids]
            dst = np.concatenate([[0], np.cumsum(lens)])
            kind = 'rows' if y.dtype == torch.float32 else 'index1'
            slots.append((None, 'y', y.dtype, (int(lens.sum()),), (kind, y, self.ptrs['y'][ids], dst, None)))

        # ---- layout: one flat buffer per dtype, 16-byte aligned sub-ranges, entries grouped by dtype in order of
        #      first appearance (exactly the rule of Complex.pack_, so that the signatures coincide)
        order = []
        for slot in slots:
            if slot[2] not in order:
                order.append(slot[2])
        slots = [slot for dt in order for slot in slots if slot[2] == dt]
        layout, totals = [], {}
        for d, key, dtype, shape, _ in slots:
            align = max(1, 16 // _ELEMENT_SIZE[dtype])
            off = totals.get(dtype, 0)
            numel = 1
            for extent in shape:
                numel *= int(extent)
            layout.append((d, key, dtype, off, tuple(shape)))
            totals[dtype] = off + (numel + align - 1) // align * align
        layout = tuple(layout)
        # ---- one host table with every segment array; destinations as (dtype, byte offset into that flat buffer)
        table, jobs_spec, table_len = [], [], [0]

        def put(arr):
            start = table_len[0]
            arr = np.asarray(arr, dtype=np.int64)
            table.append(arr)
            table_len[0] = start + len(arr)
            return start

        for (d, key, dtype, off, shape), (_, _, _, _, spec) in zip(layout, slots):
            dst_at = (dtype, off * _ELEMENT_SIZE[dtype])
            kind = spec[0]
            if kind == 'rows':
                _, src_t, src, dst, _ = spec
                width = src_t.size(1) if src_t.dim() == 2 else 1
                jobs_spec.append((src_t.data_ptr(), dst_at, 0, put(src), put(dst), None, B, 1, width, int(dst[-1])))
            elif kind == 'index1':  # 1-D int64 segments without offsets (integer labels)
                _, src_t, src, dst, _ = spec
                jobs_spec.append((src_t.data_ptr(), dst_at, 0, put(src), put(dst), None, B, 0, 1, int(dst[-1])))
            elif kind == 'index':
                _, src_t, src, dst, adds = spec
                s_off, d_off = put(src), put(dst)
                total_src = src_t.size(-1)
                for row, add in enumerate(adds):
                    jobs_spec.append((src_t.data_ptr() + 8 * row * total_src, dst_at, 8 * row * int(dst[-1]),
                                      s_off, d_off, put(add), B, 0, 1, int(dst[-1])))
            elif kind == 'fill':
                dst = spec[3]
                jobs_spec.append((None, dst_at, 0, None, put(dst), None, B, 2, 1, int(dst[-1])))
            else:  # 'table': host-known values (ptr) copied out of the uploaded table itself
                ptr = spec[1]
                pos = put(ptr)
                jobs_spec.append(('table', dst_at, 0, put([pos]), put([0, len(ptr)]), None, 1, 0, 1, len(ptr)))
        return {'B': B, 'dimension': dimension, 'dims': dims, 'cell_cnt': cell_cnt, 'cell_off': cell_off,
                'layout': layout, 'totals': totals, 'table': np.concatenate(table), 'jobs': jobs_spec}

    def launch(self, prep, out: ComplexBatch = None) -> ComplexBatch:
        """Device half: one H2D copy of the segment table + one kernel; assembles the `ComplexBatch` around the views."""
        B, dimension, dims = prep['B'], prep['dimension'], prep['dims']
        cell_cnt, cell_off, layout, totals = prep['cell_cnt'], prep['cell_off'], prep['layout'], prep['totals']
        if out is not None:
            if out.packed_signature != layout:
                raise ValueError('collate(out=...): the batch layout differs from the destination (different cell / '
                                 'message counts); collate without `out` and re-capture, or run eagerly')
            flat = out._flat
        else:
            flat = {dt: torch.zeros(max(n, 1), dtype=dt, device=self.device) for dt, n in totals.items()}
        flat_table = prep['table']
        host, dev_table = self._staging(len(flat_table))
        host[:len(flat_table)].copy_(torch.from_numpy(flat_table))
        dev_table.copy_(host, non_blocking=True)
        base = dev_table.data_ptr()
        jobs = []
        for src, (dtype, dst_bytes), extra, s_off, d_off, a_off, nseg, kind, width, n_out in prep['jobs']:
            if src == 'table':  # the source IS the table; src_start holds the position of the values inside it
                src = base
            jobs.append(_lib.CollateJob(src, flat[dtype].data_ptr() + dst_bytes + extra,
                                        None if s_off is None else base + 8 * s_off, base + 8 * d_off,
                                        None if a_off is None else base + 8 * a_off, nseg, kind, width, n_out))
        lib = _lib.load()
        arr = (_lib.CollateJob * len(jobs))(*jobs)
        with torch.cuda.device(self.device):
            nbytes = sum(2 * j.n_out * (8 if j.kind != 1 else 4 * j.row_elems) for j in jobs) if ops._profile else 0
            ops._call('collate', nbytes, lib.cwn_collate, arr, len(jobs), torch.cuda.current_stream().cuda_stream)
        self._last_slot[2].record()
        self._keepalive = dev_table  # (ring slot: stays valid for the next 8 collations)
        self._table_bytes = len(flat_table) * 8
        if out is not None:
            return out

        # ---- assemble the ComplexBatch object around the views
        views = {}
        for d, key, dtype, off, shape in layout:
            views[(d, key)] = flat[dtype][off:off + int(np.prod(shape))].view(shape)
        cochains = []
        for d in dims:
            cb = CochainBatch(d)
            for key in _SLOT_ORDER:
                if (d, key) in views:
                    cb._assign(key, views[(d, key)])
            n_d = int(cell_off[d][-1])
            cb._num_cells = n_d
            cb._num_cells_up = int(cell_off[d + 1][-1]) if (d + 1) in cell_off else 0
            if d > 0:
                cb._num_cells_down = int(cell_off[d - 1][-1])
            cb._num_cochains = B
            cb._num_cells_list = [int(v) if v > 0 else None for v in cell_cnt[d]]
            cb._ptr_host = [0] + list(np.cumsum(cell_cnt[d][cell_cnt[d] > 0]))
            cochains.append(cb)
        batch = ComplexBatch(*cochains, y=views.get((None, 'y')), num_complexes=B, dimension=dimension)
        batch._flat, batch._layout = flat, list(layout)
        batch._h2d_bytes = len(flat_table) * 8
        return batch

    def _staging(self, n):
        """A (pinned host, device) pair of int64 buffers for the segment table, from a small ring: allocating pinned
        memory per batch would cost more than the collation. A slot is reused only after 8 later collations."""
        ring = getattr(self, '_ring', None)
        if ring is None:
            ring = self._ring = {'slots': [None] * 8, 'pos': 0}
        i = ring['pos']
        ring['pos'] = (i + 1) % len(ring['slots'])
        slot = ring['slots'][i]
        if slot is None or slot[0].numel() < n:
            cap = max(1024, 2 * n)
            slot = (torch.empty(cap, dtype=torch.long).pin_memory(), torch.empty(cap, dtype=torch.long, device=self.device),
                    torch.cuda.Event())
            ring['slots'][i] = slot
        else:
            slot[2].synchronize()  # the copy issued from this slot 8 collations ago must have left the host buffer
        self._last_slot = slot
        return slot[0], slot[1]
