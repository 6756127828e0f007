"""Cochain / CochainBatch / Complex / ComplexBatch — the data API of the hot path.

Drop-in for the reference's `data/complex.py` (Cochain :36, CochainBatch :296, Complex :490, ComplexBatch :670):
same constructor signatures, attribute names, `None` conventions, index-offset rules and `get_cochain_params`
contract, re-expressed without torch_sparse and designed around a device-resident batch:

  * every cell count a kernel launch needs (`num_cells`, `num_cells_up`, `num_cells_down`, `num_complexes`) is
    known on the HOST after batching, so no `.max()+1` device->host syncs are ever required downstream;
  * `ComplexBatch.to(device)` moves the whole batch with ONE pinned staging buffer and ONE H2D copy per dtype
    (the reference does ~10 small copies per dimension, `data/complex.py:276-283,539-546`);
  * `get_cochain_params` does not materialise `up_attr`/`down_attr` (`data/complex.py:579-580,587-588`): it hands
    out a `LazyRows(x, index)` that the fused kernels consume directly and that turns into a real tensor the
    moment anybody treats it as one (so `params.kwargs['up_attr']` keeps working, `data/test_data.py:10-13`).

Host-side plumbing only: nothing here launches a message-passing kernel.
"""
import copy
import os
import logging
from typing import Dict, List, Optional, Sequence

import torch
from torch import Tensor

from cwn_b200.mp.params import CochainMessagePassingParams, LazyRows

# key -> which neighbouring cell count offsets it when cochains are concatenated (reference `__inc__`, :148-169)
_INDEX_KEYS = ('upper_index', 'lower_index', 'shared_boundaries', 'shared_coboundaries', 'boundary_index')


_CHECK_INDICES = os.environ.get('CWN_B200_CHECK_INDICES', '1') != '0'


class Cochain(object):
    """Vector-valued signal over the k-cells of a complex plus the adjacencies of those cells.

    Args mirror the reference (`data/complex.py:59-61`): dim, x [num_cells, F], upper_index / lower_index /
    boundary_index int64 [2, E] (row 0 = source, row 1 = destination), shared_boundaries / shared_coboundaries
    int64 [E], mapping, upper_orient / lower_orient, y; kwargs `num_cells`, `num_cells_up`, `num_cells_down`
    set explicit counts, any other kwarg becomes a batched attribute.
    """

    def __init__(self, dim: int, x: Tensor = None, upper_index: Tensor = None, lower_index: Tensor = None,
                 shared_boundaries: Tensor = None, shared_coboundaries: Tensor = None, mapping: Tensor = None,
                 boundary_index: Tensor = None, upper_orient=None, lower_orient=None, y=None, **kwargs):
        if dim == 0:
            assert lower_index is None
            assert shared_boundaries is None
            assert boundary_index is None
        self._dim = dim
        self._x = x
        self._mapping = mapping
        self._num_cells = None
        self._num_cells_up = None
        self._num_cells_down = None
        self._oriented = False
        self.upper_index = upper_index
        self.lower_index = lower_index
        self.boundary_index = boundary_index
        self.y = y
        self.shared_boundaries = shared_boundaries
        self.shared_coboundaries = shared_coboundaries
        self.upper_orient = upper_orient
        self.lower_orient = lower_orient
        for key, item in kwargs.items():
            if key == 'num_cells':
                self._num_cells = item
            elif key == 'num_cells_down':
                self._num_cells_down = item
            elif key == 'num_cells_up':
                self._num_cells_up = item
            else:
                setattr(self, key, item)

    # ------------------------------------------------------------------ basic accessors
    @property
    def dim(self):
        return self._dim

    @property
    def x(self):
        return self._x

    @x.setter
    def x(self, new_x):
        # reference :107-114 — models push every layer's output through this setter (`set_xs`)
        if new_x is None:
            logging.warning("Cochain features were set to None. ")
        else:
            assert self.num_cells == len(new_x)
        self._x = new_x

    @property
    def mapping(self):
        return self._mapping

    @property
    def keys(self):
        """Names of the public attributes that are set (features are reported as 'x')."""
        out = ['x'] if self._x is not None else []
        out += [k for k, v in self.__dict__.items() if not k.startswith('_') and v is not None]
        return out

    def __getitem__(self, key):
        return getattr(self, key, None)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def __contains__(self, key):
        return key in self.keys

    def __call__(self, *keys):
        for key in sorted(self.keys) if not keys else keys:
            if key in self:
                yield key, self[key]

    def __cat_dim__(self, key, value):
        return -1 if key in _INDEX_KEYS else 0

    def __inc__(self, key, value=None):
        """Offset added to `key` of the NEXT cochain when batching (reference :148-169)."""
        if key in ('upper_index', 'lower_index'):
            inc = self.num_cells
        elif key == 'shared_boundaries':
            inc = self.num_cells_down
        elif key == 'shared_coboundaries':
            inc = self.num_cells_up
        elif key == 'boundary_index':
            return [[self.num_cells_down or 0], [self.num_cells or 0]]
        else:
            inc = 0
        return 0 if inc is None else inc

    # ------------------------------------------------------------------ cell counts (host integers)
    @property
    def num_cells(self):
        # resolution order of the reference (:182-192): explicit -> x rows -> boundary destinations -> None
        if self._num_cells is not None:
            return self._num_cells
        if self._x is not None:
            return self._x.size(0)
        if self.boundary_index is not None:
            return int(self.boundary_index[1, :].max()) + 1
        assert self.upper_index is None and self.lower_index is None
        return None

    @num_cells.setter
    def num_cells(self, value):
        self._num_cells = value

    @property
    def num_cells_up(self):
        if self._num_cells_up is not None:
            return self._num_cells_up
        if self.shared_coboundaries is not None:
            assert self.upper_index is not None
            return int(self.shared_coboundaries.max()) + 1
        assert self.upper_index is None
        return 0

    @num_cells_up.setter
    def num_cells_up(self, value):
        self._num_cells_up = value

    @property
    def num_cells_down(self):
        if self.dim == 0:
            return None
        if self._num_cells_down is not None:
            return self._num_cells_down
        if self.lower_index is None:
            return 0
        raise ValueError('Cannot infer the number of cells in the cochain below.')

    @num_cells_down.setter
    def num_cells_down(self, value):
        self._num_cells_down = value

    @property
    def num_features(self):
        if self._x is None:
            return 0
        return 1 if self._x.dim() == 1 else self._x.size(1)

    # ------------------------------------------------------------------ tensor plumbing
    def _tensor_items(self):
        if self._x is not None:
            yield 'x', self._x
        for k, v in self.__dict__.items():
            if not k.startswith('_') and torch.is_tensor(v):
                yield k, v

    def _assign(self, key, value):
        if key == 'x':
            self._x = value
        else:
            setattr(self, key, value)

    def apply(self, func, *keys):
        for key, item in list(self._tensor_items()):
            if not keys or key in keys:
                self._assign(key, func(item))
        return self

    def contiguous(self, *keys):
        return self.apply(lambda t: t.contiguous(), *keys)

    def to(self, device, *keys, **kwargs):
        return self.apply(lambda t: t.to(device, **kwargs), *keys)

    def clone(self):
        new = copy.copy(self)
        new.__dict__ = {k: (v.clone() if torch.is_tensor(v) else copy.deepcopy(v)) for k, v in self.__dict__.items()}
        return new


class CochainBatch(Cochain):
    """Many cochains of one dimension stored as a single block-diagonal cochain (`batch`: cell -> member id,
    `ptr`: member boundaries). Reference: `data/complex.py:296-458`."""

    def __init__(self, dim, batch=None, ptr=None, **kwargs):
        super(CochainBatch, self).__init__(dim, **kwargs)
        self.batch = batch
        self.ptr = ptr
        self._num_cochains = None
        self._num_cells_list = None

    @classmethod
    def from_cochain_list(cls, data_list: Sequence[Cochain], follow_batch=()):
        """Concatenate cochains: features along dim 0, indices along the last dim after adding the running
        offsets of `Cochain.__inc__`. `None` members of a key are skipped but still advance the offset
        (reference :353-410); `batch`/`ptr` are built from `num_cells` (:427-432)."""
        dim = data_list[0].dim
        out = cls(dim)
        keys = []
        for data in data_list:
            for k in data.keys:
                if k not in keys:
                    keys.append(k)
        assert 'batch' not in keys and 'ptr' not in keys

        pieces: Dict[str, list] = {k: [] for k in keys}
        follow: Dict[str, list] = {}
        running: Dict[str, object] = {k: 0 for k in keys}
        batch_vec, ptr = [], [0]
        counts = {'cells': [], 'up': [], 'down': []}
        device = None
        for i, data in enumerate(data_list):
            for k in keys:
                item = data[k]
                if item is not None:
                    cum = running[k]
                    nonzero = (not isinstance(cum, int)) or cum != 0
                    if torch.is_tensor(item):
                        if item.dtype != torch.bool and nonzero:
                            item = item + cum
                        if item.dim() == 0:
                            item = item.unsqueeze(0)
                        device = item.device
                        if k in follow_batch:
                            follow.setdefault(k, []).append(torch.full(
                                (item.size(data.__cat_dim__(k, item)),), i, dtype=torch.long, device=device))
                    elif isinstance(item, (int, float)) and not isinstance(item, bool):
                        item = item + cum
                    pieces[k].append(item)
                inc = data.__inc__(k, item)
                if isinstance(inc, (list, tuple)):
                    inc = torch.tensor(inc)
                running[k] = running[k] + inc
            n = data.num_cells
            counts['cells'].append(n)
            counts['up'].append(data.num_cells_up)
            counts['down'].append(data.num_cells_down)
            if n is not None:
                batch_vec.append((i, n))
                ptr.append(ptr[-1] + n)

        ref = data_list[0]
        for k in keys:
            items = pieces[k]
            if len(items) == 0:
                continue
            if torch.is_tensor(items[0]):
                out._assign(k, torch.cat(items, ref.__cat_dim__(k, items[0])).contiguous())
            elif isinstance(items[0], (int, float)):
                out._assign(k, torch.tensor(items))
            else:
                out._assign(k, items)
        for k, vecs in follow.items():
            setattr(out, f'{k}_batch', torch.cat(vecs, 0))
        if batch_vec:
            ids = torch.tensor([i for i, _ in batch_vec], dtype=torch.long, device=device)
            reps = torch.tensor([n for _, n in batch_vec], dtype=torch.long, device=device)
            out.batch = torch.repeat_interleave(ids, reps)
            out.ptr = torch.tensor(ptr)
        # host-known totals: the reason no kernel launch downstream needs a device->host sync
        known = [n for n in counts['cells'] if n is not None]
        out._num_cells = sum(known) if known else None
        out._num_cells_up = sum(n or 0 for n in counts['up'])
        if dim > 0:
            out._num_cells_down = sum(n or 0 for n in counts['down'])
        out._num_cochains = len(data_list)
        out._num_cells_list = counts['cells']
        out._ptr_host = ptr
        return out

    def __getitem__(self, idx):
        if isinstance(idx, str):
            return super(CochainBatch, self).__getitem__(idx)
        raise NotImplementedError

    def to_cochain_list(self) -> List[Cochain]:
        raise NotImplementedError

    @property
    def num_cochains(self) -> int:
        if self._num_cochains is not None:
            return self._num_cochains
        return self.ptr.numel() + 1


class Complex(object):
    """A cochain complex: one `Cochain` per dimension 0..dimension plus an optional complex-level label.
    Reference: `data/complex.py:490-667`."""

    def __init__(self, *cochains: Cochain, y: Tensor = None, dimension: int = None):
        if len(cochains) == 0:
            raise ValueError('At least one cochain is required.')
        if dimension is None:
            dimension = len(cochains) - 1
        if len(cochains) < dimension + 1:
            raise ValueError(f'Not enough cochains passed, expected {dimension + 1}, received {len(cochains)}')
        self.dimension = dimension
        self.cochains = {i: cochains[i] for i in range(dimension + 1)}
        self.nodes = cochains[0]
        self.edges = cochains[1] if dimension >= 1 else None
        self.two_cells = cochains[2] if dimension >= 2 else None
        self.y = y
        self._consolidate()

    def _consolidate(self):
        # neighbour counts always come from the neighbouring cochain (reference :518-537)
        for dim in range(self.dimension + 1):
            cochain = self.cochains[dim]
            assert cochain.dim == dim
            if dim < self.dimension:
                n_up = self.cochains[dim + 1].num_cells
                assert n_up is not None
                cochain.num_cells_up = n_up
            if dim > 0:
                n_down = self.cochains[dim - 1].num_cells
                assert n_down is not None
                cochain.num_cells_down = n_down

    # ------------------------------------------------------------------ packed storage + device transfer
    def _slots(self):
        """(owner cochain or None for the complex label, attribute name, tensor) of every tensor of the complex."""
        slots = []
        for dim in range(self.dimension + 1):
            for key, t in self.cochains[dim]._tensor_items():
                slots.append((self.cochains[dim], key, t))
        if self.y is not None and torch.is_tensor(self.y):
            slots.append((None, 'y', self.y))
        return slots

    def _rebind(self, owner, key, value):
        if owner is None:
            self.y = value
        else:
            owner._assign(key, value)

    def pack_(self, pin_memory: bool = False):
        """Re-home every tensor into ONE flat buffer per dtype (16-byte aligned sub-ranges; attributes become
        views), optionally pinned. A packed batch crosses PCIe/NVLink-C2C as one copy per dtype and can be loaded
        into the static buffers of a captured CUDA graph (`load_packed_`). Returns self."""
        slots = self._slots()
        if slots and len({t.device for _, _, t in slots}) != 1:
            raise ValueError('pack_: all tensors must live on one device')
        by_dtype: Dict[torch.dtype, list] = {}
        for slot in slots:
            by_dtype.setdefault(slot[2].dtype, []).append(slot)
        self._flat, self._layout = {}, []
        for dtype, group in by_dtype.items():
            esz = torch.empty((), dtype=dtype).element_size()
            align = max(1, 16 // esz)
            offs, total = [], 0
            for _, _, t in group:
                offs.append(total)
                total += (t.numel() + align - 1) // align * align
            dev = group[0][2].device
            flat = torch.zeros(max(total, 1), dtype=dtype, device=dev,
                               pin_memory=bool(pin_memory and dev.type == 'cpu'))
            for (owner, key, t), o in zip(group, offs):
                view = flat[o:o + t.numel()].view(t.shape)
                view.copy_(t)
                self._rebind(owner, key, view)
                self._layout.append((None if owner is None else owner.dim, key, dtype, o, tuple(t.shape)))
            self._flat[dtype] = flat
        return self

    @property
    def packed_signature(self):
        """Hashable description of the packed layout (shapes + offsets); equal signatures => buffers are
        interchangeable element for element."""
        if getattr(self, '_layout', None) is None:
            return None
        return tuple(self._layout)

    @property
    def packed_nbytes(self):
        return sum(f.numel() * f.element_size() for f in getattr(self, '_flat', {}).values())

    def load_packed_(self, other, non_blocking=True):
        """Overwrite this (packed) complex's tensors with those of `other` (packed, identical signature) — one
        copy per dtype, no allocation: the input side of a replayed CUDA graph."""
        if self.packed_signature is None or self.packed_signature != other.packed_signature:
            raise ValueError('load_packed_: layouts differ (different cell/message counts); re-capture or run eagerly')
        for dtype, flat in self._flat.items():
            flat.copy_(other._flat[dtype], non_blocking=non_blocking)
        return self

    def to(self, device, **kwargs):
        """Move every tensor of the complex. CPU -> CUDA packs (pinned) if needed and then issues ONE async H2D
        copy per dtype (the reference copies attribute by attribute, `data/complex.py:276-283,539-546`)."""
        device = torch.device(device)
        slots = self._slots()
        packable = device.type == 'cuda' and not kwargs and all(t.device.type == 'cpu' for _, _, t in slots)
        if packable and _CHECK_INDICES:
            self.check_indices()
        if not packable:
            for owner, key, t in slots:
                self._rebind(owner, key, t.to(device, **kwargs))
            self._flat = self._layout = None
            return self
        if getattr(self, '_layout', None) is None or any(not f.is_pinned() for f in self._flat.values()):
            self.pack_(pin_memory=True)
        by_key = {(None if o is None else o.dim, k): o for o, k, _ in slots}
        dev_flat = {dtype: flat.to(device, non_blocking=True) for dtype, flat in self._flat.items()}
        self._h2d_bytes = self.packed_nbytes
        for dim, key, dtype, off, shape in self._layout:
            n = 1
            for sdim in shape:
                n *= sdim
            self._rebind(by_key[(dim, key)], key, dev_flat[dtype][off:off + n].view(shape))
        self._flat = dev_flat
        return self

    def check_indices(self):
        """Every index of every adjacency points at an existing cell. The kernels read `x[idx]` unchecked (the reference
        relies on torch's device-side assert); this is the host-side equivalent, run once per batch while the tensors
        are still on the CPU (`CWN_B200_CHECK_INDICES=0` turns it off)."""
        for dim in range(self.dimension + 1):
            c = self.cochains[dim]
            n, n_up, n_down = c.num_cells or 0, c.num_cells_up or 0, c.num_cells_down or 0
            checks = [('upper_index', c.upper_index, n), ('lower_index', c.lower_index, n),
                      ('shared_coboundaries', c.shared_coboundaries, n_up), ('shared_boundaries', c.shared_boundaries, n_down)]
            if c.boundary_index is not None and c.boundary_index.numel():
                checks += [('boundary_index[0]', c.boundary_index[0], n_down), ('boundary_index[1]', c.boundary_index[1], n)]
            for name, idx, bound in checks:
                if idx is None or idx.numel() == 0 or idx.device.type != 'cpu':
                    continue
                lo, hi = int(idx.min()), int(idx.max())
                if lo < 0 or hi >= bound:
                    raise IndexError(f'cwn_b200: {name} of dimension {dim} has entries in [{lo}, {hi}], valid range is '
                                     f'[0, {bound})')
        return self

    def get_cochain_params(self, dim: int, max_dim: int = 2, include_top_features=True,
                           include_down_features=True, include_boundary_features=True
                           ) -> CochainMessagePassingParams:
        """Inputs of `propagate` for the `dim`-cells (reference :548-602): x, upper/lower/boundary index,
        `up_attr = x_{dim+1}[shared_coboundaries]` (only if `dim < max_dim or include_top_features`),
        `down_attr = x_{dim-1}[shared_boundaries]`, `boundary_attr = x_{dim-1}`. The two gathered operands are
        handed out lazily (`LazyRows`)."""
        if dim not in self.cochains:
            raise NotImplementedError('Dim {} is not present in the complex or not yet supported.'.format(dim))
        cells = self.cochains[dim]
        upper_index, upper_features = None, None
        if cells.upper_index is not None and (dim + 1) in self.cochains:
            upper_index = cells.upper_index
            x_up = self.cochains[dim + 1].x
            if x_up is not None and (dim < max_dim or include_top_features):
                upper_features = LazyRows(x_up, cells.shared_coboundaries)
        lower_index, lower_features = None, None
        if include_down_features and cells.lower_index is not None:
            lower_index = cells.lower_index
            if dim > 0 and self.cochains[dim - 1].x is not None:
                lower_features = LazyRows(self.cochains[dim - 1].x, cells.shared_boundaries)
        boundary_index, boundary_features = None, None
        if include_boundary_features and cells.boundary_index is not None:
            boundary_index = cells.boundary_index
            if dim > 0 and self.cochains[dim - 1].x is not None:
                boundary_features = self.cochains[dim - 1].x
        params = CochainMessagePassingParams(cells.x, upper_index, lower_index,
                                             up_attr=upper_features, down_attr=lower_features,
                                             boundary_attr=boundary_features, boundary_index=boundary_index)
        # host-known row count of this cochain (lets InitReduceConv skip the reference's `.max() + 1` sync)
        params.num_cells = cells._num_cells if cells._num_cells is not None else (
            cells.x.size(0) if cells.x is not None else None)
        return params

    def get_all_cochain_params(self, max_dim: int = 2, include_top_features=True, include_down_features=True,
                               include_boundary_features=True) -> List[CochainMessagePassingParams]:
        return [self.get_cochain_params(dim, max_dim=max_dim, include_top_features=include_top_features,
                                        include_down_features=include_down_features,
                                        include_boundary_features=include_boundary_features)
                for dim in range(min(max_dim, self.dimension) + 1)]

    def get_labels(self, dim=None):
        if dim is None:
            return self.y
        if dim in self.cochains:
            return self.cochains[dim].y
        raise NotImplementedError('Dim {} is not present in the complex or not yet supported.'.format(dim))

    def set_xs(self, xs: List[Tensor]):
        assert (self.dimension + 1) >= len(xs)
        for i, x in enumerate(xs):
            self.cochains[i].x = x

    @property
    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None and not k.startswith('_')]

    def __getitem__(self, key):
        return getattr(self, key, None)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def __contains__(self, key):
        return key in self.keys


class ComplexBatch(Complex):
    """A batch of complexes stored as one complex of `CochainBatch`es. Reference: `data/complex.py:670-728`."""

    def __init__(self, *cochains: CochainBatch, dimension: int, y: Tensor = None, num_complexes: int = None):
        super(ComplexBatch, self).__init__(*cochains, y=y)
        self.num_complexes = num_complexes
        self.dimension = dimension

    @classmethod
    def from_complex_list(cls, data_list: List[Complex], follow_batch=(), max_dim: int = 2):
        dimension = min(max(data.dimension for data in data_list), max_dim)
        per_dim = [[] for _ in range(dimension + 1)]
        labels, all_labelled = [], True
        for comp in data_list:
            for dim in range(dimension + 1):
                if dim in comp.cochains:
                    per_dim[dim].append(comp.cochains[dim])
                else:
                    # pad with an empty cochain that still carries how many (dim-1)-cells lie below it, so the
                    # boundary offsets of the following members stay right (reference :710-716)
                    pad = Cochain(dim=dim)
                    if dim - 1 in comp.cochains:
                        pad.num_cells_down = comp.cochains[dim - 1].num_cells
                    per_dim[dim].append(pad)
            all_labelled &= comp.y is not None
            if all_labelled:
                labels.append(comp.y)
        batched = [CochainBatch.from_cochain_list(lst, follow_batch=follow_batch) for lst in per_dim]
        y = torch.cat(labels, 0) if all_labelled else None
        return cls(*batched, y=y, num_complexes=len(data_list), dimension=dimension)
