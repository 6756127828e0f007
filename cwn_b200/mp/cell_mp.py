"""`CochainMessagePassing` — the operator API of the hot path (reference `mp/cell_mp.py:41-524`).

Same contract as the reference: `propagate(up_index, down_index, boundary_index, up_size, down_size,
boundary_size, **kwargs)` runs up to three adjacency passes (upper, lower, boundary) and hands their aggregated
outputs to `update`. Each pass is `message_{adj}` followed by `aggregate_{adj}`; hook arguments are resolved BY
NAME from `kwargs` — an argument ending in `_j` / `_i` with the adjacency prefix (`up_`, `down_`, `boundary_`)
is the matrix of that name gathered at the source (`index[0]`) / destination (`index[1]`) of every message
(`boundary_*_j` gathers from `boundary_attr`), flow `source_to_target`.

What is different is WHERE the arithmetic runs. Per pass:
  * hooks left at their defaults (identity message, `aggr` in add/mean/max): ONE fused sm_100a kernel on a CSR
    plan grouped by destination — no `[E, F]` temporary, no atomics, deterministic (ops.gather_scatter);
  * a subclass may offer `fused_message_and_aggregate_{adj}` (e.g. the coboundary MLP of SparseCINConv) which is
    tried next;
  * any other override keeps the reference's observable behaviour: operands are gathered by the CUDA gather
    kernel, the user's `message_*` runs on them, and `aggregate_*` reduces with the CUDA segmented kernel.
There is no CPU path: tensors that are not on a CUDA device raise.
"""
import inspect
from typing import Dict, List, Optional, Set

import torch
from torch import Tensor

from cwn_b200 import ops
from cwn_b200.mp.params import CochainMessagePassingParams, LazyRows, as_tensor  # noqa: F401 (re-export)

_EMPTY = inspect.Parameter.empty
_ADJACENCIES = ('up', 'down', 'boundary')


class HookSignatures(object):
    """Parameter tables of the overridable hooks (the role PyG's `Inspector` plays for the reference,
    `mp/cell_mp_inspector.py:34-49`): which names a hook wants, whether a subclass overrides it, and how a dict of
    collected values is dealt onto a hook's parameters."""

    def __init__(self, owner):
        self.owner = owner
        self.params: Dict[str, Dict[str, inspect.Parameter]] = {}

    def inspect(self, func, pop_first_n: int = 0):
        params = list(inspect.signature(func).parameters.items())[pop_first_n:]
        self.params[func.__name__] = dict(params)

    def keys(self, func_names: Optional[List[str]] = None) -> Set[str]:
        names = set()
        for fn in func_names or list(self.params):
            names.update(self.params[fn].keys())
        return names

    def implements(self, func_name: str) -> bool:
        """True iff some class below CochainMessagePassing in the owner's MRO defines `func_name`."""
        for klass in type(self.owner).__mro__:
            if klass.__name__ == 'CochainMessagePassing':
                return False
            if func_name in klass.__dict__:
                return True
        return False

    def distribute(self, func_name: str, collected: dict) -> dict:
        out = {}
        for key, param in self.params[func_name].items():
            data = collected.get(key, _EMPTY)
            if data is _EMPTY:
                if param.default is _EMPTY:
                    raise TypeError(f'Required parameter {key} is empty.')
                data = param.default
            out[key] = data
        return out


class CochainMessagePassing(torch.nn.Module):
    """Base class for message passing on cochain complexes (boundary, upper and lower adjacencies).

    Args (reference `mp/cell_mp.py:81-91`): up_msg_size, down_msg_size, aggr_up / aggr_down / aggr_boundary in
    {"add", "mean", "max", None}, flow, node_dim, boundary_msg_size (defaults to down_msg_size), use_down_msg,
    use_boundary_msg.
    """

    special_args: Set[str] = set(
        f'{adj}_{suffix}' for adj in _ADJACENCIES
        for suffix in ('index', 'adj_t', 'index_i', 'index_j', 'size', 'size_i', 'size_j', 'ptr', 'dim_size')
    ) | {f'agg_{adj}_index' for adj in _ADJACENCIES}

    def __init__(self, up_msg_size, down_msg_size, aggr_up: Optional[str] = "add",
                 aggr_down: Optional[str] = "add", aggr_boundary: Optional[str] = "add",
                 flow: str = "source_to_target", node_dim: int = -2, boundary_msg_size=None,
                 use_down_msg=True, use_boundary_msg=True):
        super(CochainMessagePassing, self).__init__()
        self.up_msg_size = up_msg_size
        self.down_msg_size = down_msg_size
        self.use_boundary_msg = use_boundary_msg
        self.use_down_msg = use_down_msg
        self.boundary_msg_size = down_msg_size if boundary_msg_size is None else boundary_msg_size
        self.aggr_up = aggr_up
        self.aggr_down = aggr_down
        self.aggr_boundary = aggr_boundary
        assert self.aggr_up in ['add', 'mean', 'max', None]
        assert self.aggr_down in ['add', 'mean', 'max', None]
        self.flow = flow
        assert self.flow in ['source_to_target', 'target_to_source']
        self.node_dim = node_dim
        if node_dim not in (-2, 0):
            raise NotImplementedError('cwn_b200: cells must be indexed along dim 0 of a [num_cells, F] matrix')

        sig = self.inspector = HookSignatures(self)
        for adj in _ADJACENCIES:
            sig.inspect(getattr(self, f'message_{adj}'))
            sig.inspect(getattr(self, f'aggregate_{adj}'), pop_first_n=1)
            sig.inspect(getattr(self, f'message_and_aggregate_{adj}'), pop_first_n=1)
        sig.inspect(self.update, pop_first_n=3)
        self.__user_args__ = sig.keys([f'{kind}_{adj}' for kind in ('message', 'aggregate')
                                       for adj in _ADJACENCIES]).difference(self.special_args)
        self.__fused_user_args__ = sig.keys([f'message_and_aggregate_{adj}'
                                             for adj in _ADJACENCIES]).difference(self.special_args)
        self.__update_user_args__ = sig.keys(['update']).difference(self.special_args)
        self.fuse_up = sig.implements('message_and_aggregate_up')
        self.fuse_down = sig.implements('message_and_aggregate_down')
        self.fuse_boundary = sig.implements('message_and_aggregate_boundary')
        # which passes may take the single-kernel identity path (hooks untouched by subclasses)
        self._default_hooks = {adj: not (sig.implements(f'message_{adj}') or sig.implements(f'aggregate_{adj}'))
                               for adj in _ADJACENCIES}

    # ------------------------------------------------------------------ input checks (reference :146-193)
    def __check_input_together__(self, index_up, index_down, size_up, size_down):
        if (index_up is not None and index_down is not None
                and size_up is not None and size_down is not None):
            assert size_up[0] == size_down[0]
            assert size_up[1] == size_down[1]

    def __check_input_separately__(self, index, size):
        the_size: List[Optional[int]] = [None, None]
        if isinstance(index, Tensor):
            assert index.dtype == torch.long
            assert index.dim() == 2
            assert index.size(0) == 2
            if size is not None:
                the_size[0] = size[0]
                the_size[1] = size[1]
            return the_size
        if index is None:
            return the_size
        raise ValueError('`MessagePassing.propagate` only supports `torch.LongTensor` of shape '
                         '`[2, num_messages]` for argument `edge_index` (SparseTensor adjacencies are never '
                         'constructed by the reference and are not supported).')

    def __set_size__(self, size: List[Optional[int]], dim: int, src: Tensor):
        the_size = size[dim]
        if the_size is None:
            size[dim] = src.size(0)
        elif the_size != src.size(0):
            raise ValueError(f'Encountered tensor with size {src.size(0)} in dimension {self.node_dim}, '
                             f'but expected size {the_size}.')

    def __lift__(self, src, index, dim):
        return ops.gather_rows(as_tensor(src), index[dim])

    # ------------------------------------------------------------------ operand resolution (reference :209-282)
    def _operand(self, arg: str, adjacency: str, kwargs: dict):
        """(matrix to gather from, matrix whose row count sizes that side) for a `{adjacency}_<name>_{i|j}` arg,
        or None if the argument belongs to another adjacency."""
        prefix = adjacency + '_'
        if not arg.startswith(prefix):
            return None
        name = arg[len(prefix):-2]
        if adjacency == 'boundary' and arg.endswith('_j'):
            return kwargs.get('boundary_attr', _EMPTY), kwargs.get(name, _EMPTY)
        data = kwargs.get(name, _EMPTY)
        return data, data

    def __collect__(self, args, index, size, adjacency, kwargs):
        i, j = (1, 0) if self.flow == 'source_to_target' else (0, 1)
        assert adjacency in _ADJACENCIES
        out = {}
        for arg in args:
            if arg[-2:] not in ('_i', '_j'):
                out[arg] = as_tensor(kwargs.get(arg, _EMPTY))
            elif index is not None:
                operand = self._operand(arg, adjacency, kwargs)
                if operand is None:
                    continue
                data, size_data = operand
                if isinstance(data, (tuple, list)):
                    raise ValueError('This format is not supported for cellular message passing')
                data = as_tensor(data)
                if isinstance(data, Tensor):
                    dim = 0 if arg.endswith('_j') else 1
                    self.__set_size__(size, dim, as_tensor(size_data))
                    data = self.__lift__(data, index, j if arg.endswith('_j') else i)
                out[arg] = data
        if isinstance(index, Tensor):
            out[f'{adjacency}_adj_t'] = None
            out[f'{adjacency}_ptr'] = None
            out[f'{adjacency}_index'] = index
            out[f'{adjacency}_index_i'] = index[i]
            out[f'{adjacency}_index_j'] = index[j]
            out[f'agg_{adjacency}_index'] = out[f'{adjacency}_index_i']
        out[f'{adjacency}_size'] = size
        out[f'{adjacency}_size_i'] = size[1] or size[0]
        out[f'{adjacency}_size_j'] = size[0] or size[1]
        out[f'{adjacency}_dim_size'] = out[f'{adjacency}_size_i']
        return out

    def get_msg_and_agg_func(self, adjacency):
        return getattr(self, f'message_and_aggregate_{adjacency}', None) if adjacency in _ADJACENCIES else None

    def get_msg_func(self, adjacency):
        return getattr(self, f'message_{adjacency}', None) if adjacency in _ADJACENCIES else None

    def get_agg_func(self, adjacency):
        return getattr(self, f'aggregate_{adjacency}', None) if adjacency in _ADJACENCIES else None

    def get_fuse_boolean(self, adjacency):
        return getattr(self, f'fuse_{adjacency}', None) if adjacency in _ADJACENCIES else None

    # ------------------------------------------------------------------ one adjacency pass
    def _identity_pass(self, index, adjacency, size, kwargs):
        """Default hooks: one fused gather->reduce kernel. Returns NotImplemented when the generic path must run
        (so that its error behaviour — missing operands, size mismatches — stays that of the reference)."""
        if self.flow != 'source_to_target' or not isinstance(index, Tensor):
            return NotImplemented
        wanted = self.inspector.params[f'message_{adjacency}']
        for name in wanted:  # e.g. message_up(up_x_j, up_attr): `up_attr` must have been passed (may be None)
            if not name.endswith('_j') and name not in kwargs and wanted[name].default is _EMPTY:
                return NotImplemented
        x = kwargs.get('x', _EMPTY)
        src = kwargs.get('boundary_attr', _EMPTY) if adjacency == 'boundary' else x
        if not isinstance(x, Tensor) or not isinstance(src, Tensor):
            return NotImplemented
        self.__set_size__(size, 0, x)  # same sizing rule as __collect__ (both sides sized by `x`)
        n_dst = size[1] or size[0]
        aggr = getattr(self, f'aggr_{adjacency}')
        if aggr is None:
            return NotImplemented
        return ops.gather_scatter(src, index, n_dst, reduce=aggr)

    def __message_and_aggregate__(self, index, adjacency: str, size: List[Optional[int]] = None, **kwargs):
        assert adjacency in _ADJACENCIES
        if self._default_hooks[adjacency]:
            out = self._identity_pass(index, adjacency, size, kwargs)
            if out is not NotImplemented:
                return out
        fused = getattr(self, f'fused_message_and_aggregate_{adjacency}', None)
        if fused is not None:
            out = fused(index, size, kwargs)
            if out is not NotImplemented:
                return out
        for key, param in self.inspector.params[f'message_{adjacency}'].items():
            # fail on a never-passed plain hook argument BEFORE any gather is launched (same TypeError as the
            # reference's Inspector.distribute)
            if key[-2:] not in ('_i', '_j') and key not in kwargs and key not in self.special_args \
                    and param.default is _EMPTY:
                raise TypeError(f'Required parameter {key} is empty.')
        coll_dict = self.__collect__(self.__user_args__, index, size, adjacency, kwargs)
        msg_kwargs = self.inspector.distribute(f'message_{adjacency}', coll_dict)
        out = self.get_msg_func(adjacency)(**msg_kwargs)
        aggr_kwargs = self.inspector.distribute(f'aggregate_{adjacency}', coll_dict)
        return self.get_agg_func(adjacency)(out, **aggr_kwargs)

    def propagate(self, up_index: Optional[Tensor], down_index: Optional[Tensor],
                  boundary_index: Optional[Tensor], up_size=None, down_size=None, boundary_size=None, **kwargs):
        """The initial call to start propagating messages (reference :357-392)."""
        up_size = self.__check_input_separately__(up_index, up_size)
        down_size = self.__check_input_separately__(down_index, down_size)
        boundary_size = self.__check_input_separately__(boundary_index, boundary_size)
        self.__check_input_together__(up_index, down_index, up_size, down_size)

        up_out, down_out, boundary_out = None, None, None
        if up_index is not None:
            up_out = self.__message_and_aggregate__(up_index, 'up', up_size, **kwargs)
        if self.use_down_msg and down_index is not None:
            down_out = self.__message_and_aggregate__(down_index, 'down', down_size, **kwargs)
        if self.use_boundary_msg and 'boundary_attr' in kwargs and kwargs['boundary_attr'] is not None:
            boundary_out = self.__message_and_aggregate__(boundary_index, 'boundary', boundary_size, **kwargs)

        coll_dict = {}
        for arg in self.__update_user_args__:
            if arg[-2:] not in ('_i', '_j'):
                coll_dict[arg] = as_tensor(kwargs.get(arg, _EMPTY))
        update_kwargs = self.inspector.distribute('update', coll_dict)
        return self.update(up_out, down_out, boundary_out, **update_kwargs)

    # ------------------------------------------------------------------ overridable hooks
    def message_up(self, up_x_j: Tensor, up_attr: Tensor) -> Tensor:
        """Message from upper-adjacent cell j to cell i for every column of `up_index`; `up_attr` holds the
        features of the shared coboundary. Default: the neighbour's features."""
        return up_x_j

    def message_down(self, down_x_j: Tensor, down_attr: Tensor) -> Tensor:
        """Message from lower-adjacent cell j; `down_attr` holds the features of the shared boundary."""
        return down_x_j

    def message_boundary(self, boundary_x_j: Tensor):
        """Message from boundary cell j (a row of `boundary_attr`) to the cell it bounds."""
        return boundary_x_j

    def _aggregate(self, inputs, index, ptr, dim_size, aggr):
        if ptr is not None:
            raise NotImplementedError('cwn_b200: CSR `ptr` aggregation belongs to the SparseTensor path, which '
                                      'the reference never exercises')
        return ops.scatter_rows(inputs, index, dim_size, reduce=aggr)

    def aggregate_up(self, inputs: Tensor, agg_up_index: Tensor, up_ptr: Optional[Tensor] = None,
                     up_dim_size: Optional[int] = None) -> Tensor:
        return self._aggregate(inputs, agg_up_index, up_ptr, up_dim_size, self.aggr_up)

    def aggregate_down(self, inputs: Tensor, agg_down_index: Tensor, down_ptr: Optional[Tensor] = None,
                       down_dim_size: Optional[int] = None) -> Tensor:
        return self._aggregate(inputs, agg_down_index, down_ptr, down_dim_size, self.aggr_down)

    def aggregate_boundary(self, inputs: Tensor, agg_boundary_index: Tensor,
                           boundary_ptr: Optional[Tensor] = None,
                           boundary_dim_size: Optional[int] = None) -> Tensor:
        return self._aggregate(inputs, agg_boundary_index, boundary_ptr, boundary_dim_size, self.aggr_boundary)

    def message_and_aggregate_up(self, up_adj_t) -> Tensor:
        raise NotImplementedError

    def message_and_aggregate_down(self, down_adj_t) -> Tensor:
        raise NotImplementedError

    def message_and_aggregate_boundary(self, boundary_adj_t) -> Tensor:
        raise NotImplementedError

    def update(self, up_inputs: Optional[Tensor], down_inputs: Optional[Tensor],
               boundary_inputs: Optional[Tensor], x: Tensor):
        """Absent passes become zeros of width `{up,down,boundary}_msg_size` (reference :511-524)."""
        if up_inputs is None:
            up_inputs = torch.zeros(x.size(0), self.up_msg_size, device=x.device)
        if down_inputs is None:
            down_inputs = torch.zeros(x.size(0), self.down_msg_size, device=x.device)
        if boundary_inputs is None:
            boundary_inputs = torch.zeros(x.size(0), self.boundary_msg_size, device=x.device)
        return up_inputs, down_inputs, boundary_inputs
