"""`CochainMessagePassingParams` (reference `mp/cell_mp.py:527-550`) and `LazyRows`, the deferred row gather that
lets the data API keep handing out `up_attr` / `down_attr` (reference `data/complex.py:576-588`) without paying
for an `[E, F]` materialisation when the consumer is a fused kernel."""
import torch
from torch import Tensor


class LazyRows(object):
    """`source[index]` (row gather along dim 0), not yet computed.

    Fused kernels read `.source` / `.index` and gather on the fly. Anything that treats the object as a tensor
    (a torch function, `.size()`, arithmetic, indexing) triggers `materialize()` once and is forwarded to the
    real tensor, so reference-style code such as `torch.equal(expected, params.kwargs['up_attr'])` is unchanged.
    """

    __slots__ = ('source', 'index', '_dense')

    def __init__(self, source: Tensor, index: Tensor):
        self.source = source
        self.index = index
        self._dense = None

    def materialize(self) -> Tensor:
        if self._dense is None:
            if self.source.is_cuda:
                from cwn_b200 import ops
                self._dense = ops.gather_rows(self.source, self.index)
            else:
                # host-side data plumbing (collation workers, fixtures); message passing itself is CUDA-only
                self._dense = torch.index_select(self.source, 0, self.index)
        return self._dense

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}

        def dense(a):
            if isinstance(a, LazyRows):
                return a.materialize()
            if isinstance(a, (list, tuple)):
                return type(a)(dense(v) for v in a)
            return a

        return func(*dense(args), **{k: dense(v) for k, v in kwargs.items()})

    def __getattr__(self, name):
        return getattr(self.materialize(), name)

    def __getitem__(self, item):
        return self.materialize()[item]

    def __len__(self):
        return self.index.numel()

    def _binary(name):
        def op(self, other):
            other = other.materialize() if isinstance(other, LazyRows) else other
            return getattr(self.materialize(), name)(other)
        return op

    __add__ = _binary('__add__')
    __radd__ = _binary('__radd__')
    __sub__ = _binary('__sub__')
    __rsub__ = _binary('__rsub__')
    __mul__ = _binary('__mul__')
    __rmul__ = _binary('__rmul__')
    __truediv__ = _binary('__truediv__')
    __eq__ = _binary('__eq__')
    __hash__ = None
    del _binary

    def __repr__(self):
        return f'LazyRows(source={tuple(self.source.shape)}, index={tuple(self.index.shape)})'


def as_tensor(value):
    """Dense view of a maybe-lazy operand (None passes through)."""
    return value.materialize() if isinstance(value, LazyRows) else value


class CochainMessagePassingParams:
    """Inputs of `propagate` for one cochain: `x`, `up_index`, `down_index` and keyword operands
    (`up_attr`, `down_attr`, `boundary_attr`, `boundary_index`), exactly as in the reference."""

    def __init__(self, x: Tensor, up_index: Tensor = None, down_index: Tensor = None, **kwargs):
        self.x = x
        self.up_index = up_index
        self.down_index = down_index
        self.kwargs = kwargs
        self.boundary_index = self.kwargs.get('boundary_index', None)
        self.boundary_attr = self.kwargs.get('boundary_attr', None)
