"""Cochain convolution layers on top of `CochainMessagePassing` (reference `mp/layers.py`).

Module structure, constructor signatures and parameter names follow the reference so that a reference
state_dict loads unchanged (`mp_levels.{d}.update_up_nn.0.weight`, `...msg_up_nn.1.weight`, `eps1`, ...).
The compute differs: `SparseCINCochainConv.forward` recognises the nets `SparseCINConv` builds by default and then
issues at most TWO fused kernels per cochain (upper pass with or without coboundary MLP, boundary pass), each
with the GIN residual `(1 + eps) * x` folded in; anything it does not recognise (user-supplied callables,
subclasses overriding a hook) goes through the generic hook protocol of `CochainMessagePassing.propagate`.
"""
from abc import ABC, abstractmethod
from typing import Callable, Optional

import torch
import torch.nn.functional as F
from torch import Tensor
from torch.nn import Linear, Sequential, BatchNorm1d as BN

from cwn_b200 import ops
from cwn_b200.streams import run_concurrently
from cwn_b200.mp.cell_mp import CochainMessagePassing, CochainMessagePassingParams
from cwn_b200.mp.nn import activation_name, reset
from cwn_b200.mp.params import LazyRows, as_tensor


# ----------------------------------------------------------------------------------------------- test vehicles
class DummyCochainMessagePassing(CochainMessagePassing):
    """Parameter-free layer used by the known-answer tests: message = neighbour + shared (co)boundary features
    (reference `mp/layers.py:14-40`)."""

    def __init__(self, up_msg_size, down_msg_size, boundary_msg_size=None, use_boundary_msg=False,
                 use_down_msg=True):
        super(DummyCochainMessagePassing, self).__init__(up_msg_size, down_msg_size,
                                                         boundary_msg_size=boundary_msg_size,
                                                         use_boundary_msg=use_boundary_msg,
                                                         use_down_msg=use_down_msg)

    def message_up(self, up_x_j: Tensor, up_attr: Tensor) -> Tensor:
        return up_x_j + up_attr

    def message_down(self, down_x_j: Tensor, down_attr: Tensor) -> Tensor:
        return down_x_j + down_attr

    def forward(self, cochain: CochainMessagePassingParams):
        up_out, down_out, boundary_out = self.propagate(cochain.up_index, cochain.down_index,
                                                        cochain.boundary_index, x=cochain.x,
                                                        up_attr=cochain.kwargs['up_attr'],
                                                        down_attr=cochain.kwargs['down_attr'],
                                                        boundary_attr=cochain.kwargs['boundary_attr'])
        return cochain.x + up_out + down_out + boundary_out


class _PerDimension(torch.nn.Module):
    """Runs `mp_levels[d]` on the parameters of dimension d (shared `forward` of the *Conv containers)."""

    concurrent_dims = True  # False when the dimensions share nn objects (their BatchNorm buffers would race)

    def forward(self, *cochain_params: CochainMessagePassingParams, start_to_process=0):
        assert len(cochain_params) <= self.max_dim + 1

        def level(dim):
            if dim < start_to_process:
                return lambda: cochain_params[dim].x
            return lambda: self.mp_levels[dim].forward(cochain_params[dim])

        x0 = cochain_params[0].x if len(cochain_params) else None
        device = x0.device if torch.is_tensor(x0) else None
        # the dimensions of a layer are independent: one concurrent branch each
        return run_concurrently([level(dim) for dim in range(len(cochain_params))],
                                device if self.concurrent_dims else None)


class DummyCellularMessagePassing(_PerDimension):
    def __init__(self, input_dim=1, max_dim: int = 2, use_boundary_msg=False, use_down_msg=True):
        super(DummyCellularMessagePassing, self).__init__()
        self.max_dim = max_dim
        self.mp_levels = torch.nn.ModuleList(
            DummyCochainMessagePassing(input_dim, input_dim, boundary_msg_size=input_dim,
                                       use_boundary_msg=use_boundary_msg, use_down_msg=use_down_msg)
            for _ in range(max_dim + 1))


# ----------------------------------------------------------------------------------------------- dense CIN
class CINCochainConv(CochainMessagePassing):
    """Dense CIN cochain layer: upper and lower messages through an MLP over `cat[x_j, attr]`, one update MLP
    (reference `mp/layers.py:62-103`). The message nets are arbitrary callables (and contain BatchNorm over the
    message population in `CIN0`), so the passes run through the generic gather -> hook -> reduce kernels."""

    def __init__(self, up_msg_size: int, down_msg_size: int, msg_up_nn: Callable, msg_down_nn: Callable,
                 update_nn: Callable, eps: float = 0., train_eps: bool = False):
        super(CINCochainConv, self).__init__(up_msg_size, down_msg_size, use_boundary_msg=False)
        self.msg_up_nn = msg_up_nn
        self.msg_down_nn = msg_down_nn
        self.update_nn = update_nn
        self.initial_eps = eps
        if train_eps:
            self.eps = torch.nn.Parameter(torch.Tensor([eps]))
        else:
            self.register_buffer('eps', torch.Tensor([eps]))
        self.reset_parameters()

    fuse_messages = True  # K5: message MLP + BatchNorm over the messages inside the passes (ops.cin_message_pass)

    @staticmethod
    def _message_form(nn_):
        """(Linear, activation name, BatchNorm1d or None) of a message net `Sequential(Linear, act[, BatchNorm1d])`
        (the nets of CIN0 / EdgeCIN0, reference mp/models.py:40-47), else None."""
        if not isinstance(nn_, Sequential) or len(nn_) not in (2, 3) or not isinstance(nn_[0], Linear) or nn_[0].bias is None:
            return None
        act = activation_name(nn_[1])
        if act is None:
            return None
        bn = nn_[2] if len(nn_) == 3 else None
        if bn is not None and not (isinstance(bn, BN) and bn.affine):
            return None
        return nn_[0], act, bn

    def _can_fuse(self, nn_, x, index, attr):
        """Is this adjacency inside K5's closed form? (Decided for BOTH adjacencies before either runs: a half-fused
        layer falling back to `propagate` would update shared BatchNorm buffers twice.)"""
        form = self._message_form(nn_)
        if form is None or not isinstance(attr, LazyRows) or not (x.is_cuda and x.dtype == torch.float32):
            return False
        lin, act, bn = form
        if attr.source.dim() != 2 or x.size(1) + attr.source.size(1) != lin.in_features \
                or attr.source.dtype != torch.float32:
            return False
        if bn is not None and bn.training and index.size(1) < 2:
            return False  # torch raises "Expected more than 1 value per channel": keep that behaviour
        return True

    def _fused_pass(self, nn_, x, index, attr):
        """One adjacency through K5 (ops.cin_message_pass); the split-weight products through the grouped linear."""
        lin, act, bn = self._message_form(nn_)
        fx = x.size(1)
        from cwn_b200 import fused
        P, Q = fused.grouped_linear([(x, lin.weight, 0, None), (attr.source, lin.weight, fx, lin.bias)])
        return ops.cin_message_pass(P, Q, index, attr.index, x.size(0), act, bn)

    def _hooks_untouched(self):
        klass = type(self)
        return all(getattr(klass, name) is getattr(CINCochainConv, name)
                   for name in ('message_up', 'message_down', 'aggregate_up', 'aggregate_down', 'update', 'propagate')) \
            and self.aggr_up == 'add' and self.aggr_down == 'add' and self.flow == 'source_to_target'

    def forward(self, cochain: CochainMessagePassingParams):
        x = cochain.x
        up_attr, down_attr = cochain.kwargs['up_attr'], cochain.kwargs['down_attr']
        fusable = self.fuse_messages and self._hooks_untouched() and isinstance(x, Tensor) and x.dim() == 2
        if fusable:  # absent adjacency: zeros [N, msg_size] in the reference, i.e. only the residual (widths must agree)
            up_ok = (self.up_msg_size == x.size(1)) if cochain.up_index is None else \
                (up_attr is not None and self._can_fuse(self.msg_up_nn, x, cochain.up_index, up_attr))
            down_ok = (self.down_msg_size == x.size(1)) if cochain.down_index is None else \
                self._can_fuse(self.msg_down_nn, x, cochain.down_index, down_attr)
            fusable = up_ok and down_ok
        if fusable:
            # in the order of the reference: the upper net first, then the lower one (they may share BatchNorm buffers
            # with other dimensions, whose running statistics are updated sequentially)
            out_up = None if cochain.up_index is None else self._fused_pass(self.msg_up_nn, x, cochain.up_index, up_attr)
            out_down = None if cochain.down_index is None else \
                self._fused_pass(self.msg_down_nn, x, cochain.down_index, down_attr)
        else:
            out_up, out_down, _ = self.propagate(cochain.up_index, cochain.down_index, None, x=x, up_attr=up_attr,
                                                 down_attr=down_attr)
        res = (1 + self.eps) * x
        out_up = res if out_up is None else out_up + res
        out_down = res if out_down is None else out_down + res
        return self.update_nn(out_up + out_down)

    def reset_parameters(self):
        reset(self.msg_up_nn)
        reset(self.msg_down_nn)
        reset(self.update_nn)
        self.eps.data.fill_(self.initial_eps)

    def message_up(self, up_x_j: Tensor, up_attr: Tensor) -> Tensor:
        if up_attr is not None:
            return self.msg_up_nn(torch.cat([up_x_j, up_attr], dim=-1))
        return self.msg_up_nn(up_x_j)

    def message_down(self, down_x_j: Tensor, down_attr: Tensor) -> Tensor:
        return self.msg_down_nn(torch.cat([down_x_j, down_attr], dim=-1))


class CINConv(_PerDimension):
    """One `CINCochainConv` per dimension, all sharing the SAME nn objects (reference `mp/layers.py:106-124`)."""

    concurrent_dims = False  # the SAME BatchNorm objects serve every dimension, in order

    def __init__(self, up_msg_size: int, down_msg_size: int, msg_up_nn: Callable, msg_down_nn: Callable,
                 update_nn: Callable, eps: float = 0., train_eps: bool = False, max_dim: int = 2):
        super(CINConv, self).__init__()
        self.max_dim = max_dim
        self.mp_levels = torch.nn.ModuleList(
            CINCochainConv(up_msg_size, down_msg_size, msg_up_nn, msg_down_nn, update_nn, eps, train_eps)
            for _ in range(max_dim + 1))

    def forward(self, *cochain_params: CochainMessagePassingParams):
        return super(CINConv, self).forward(*cochain_params)


class EdgeCINConv(torch.nn.Module):
    """CIN layer that passes messages only up to the edges (reference `mp/layers.py:127-151`): a vertex level with upper
    messages and an edge level with upper and lower messages, each a `CINCochainConv` with its own nets. Vertices have
    no lower adjacency, so their down net is the reference's `lambda *args: None`."""

    def __init__(self, up_msg_size: int, down_msg_size: int, v_msg_up_nn: Callable, e_msg_down_nn: Callable,
                 e_msg_up_nn: Callable, v_update_nn: Callable, e_update_nn: Callable, eps: float = 0.,
                 train_eps=False):
        super(EdgeCINConv, self).__init__()
        self.max_dim = 1
        self.mp_levels = torch.nn.ModuleList()
        v_mp = CINCochainConv(up_msg_size, down_msg_size, v_msg_up_nn, lambda *args: None, v_update_nn, eps, train_eps)
        e_mp = CINCochainConv(up_msg_size, down_msg_size, e_msg_up_nn, e_msg_down_nn, e_update_nn, eps, train_eps)
        self.mp_levels.extend([v_mp, e_mp])

    def forward(self, *cochain_params: CochainMessagePassingParams):
        assert len(cochain_params) <= self.max_dim + 1
        device = cochain_params[0].x.device if len(cochain_params) else None
        # the two levels own separate nets: independent branches
        return run_concurrently([(lambda d=d: self.mp_levels[d].forward(cochain_params[d]))
                                 for d in range(len(cochain_params))], device)

    def reset_parameters(self):
        for level in self.mp_levels:
            level.reset_parameters()


class OrientedConv(CochainMessagePassing):
    """Orientation-equivariant edge convolution (reference `mp/layers.py:430-470`): messages are the neighbour's
    features times the relative orientation (+-1) of the two cells, carried through `up_attr` / `down_attr`;
    out = act(update(x) + update_up(SUM_up) + update_down(SUM_down)). Takes a `Cochain` (not params), as the reference."""

    def __init__(self, dim: int, up_msg_size: int, down_msg_size: int, update_up_nn: Optional[Callable],
                 update_down_nn: Optional[Callable], update_nn: Optional[Callable], act_fn, orient=True):
        super(OrientedConv, self).__init__(up_msg_size, down_msg_size, use_boundary_msg=False)
        self.dim = dim
        self.update_up_nn = update_up_nn
        self.update_down_nn = update_down_nn
        self.update_nn = update_nn
        self.act_fn = act_fn
        self.orient = orient

    def forward(self, cochain):
        assert len(cochain.upper_orient) == cochain.upper_index.size(1)
        assert len(cochain.lower_orient) == cochain.lower_index.size(1)
        # (the reference also asserts `index.max() < len(x)`: a device->host sync per layer; out-of-range indices are
        #  caught by cwn_check_index_range in debug mode instead)
        out_up, out_down, _ = self.propagate(cochain.upper_index, cochain.lower_index, None, x=cochain.x,
                                             up_attr=cochain.upper_orient.view(-1, 1),
                                             down_attr=cochain.lower_orient.view(-1, 1))
        out_up = self.update_up_nn(out_up)
        out_down = self.update_down_nn(out_down)
        x = self.update_nn(cochain.x)
        return self.act_fn(x + out_up + out_down)

    def reset_parameters(self):
        reset(self.update_up_nn)
        reset(self.update_down_nn)
        reset(self.update_nn)

    def message_up(self, up_x_j: Tensor, up_attr: Tensor) -> Tensor:
        if self.orient:
            return up_x_j * up_attr
        return up_x_j

    def message_down(self, down_x_j: Tensor, down_attr: Tensor) -> Tensor:
        if self.orient:
            return down_x_j * down_attr
        return down_x_j


# ----------------------------------------------------------------------------------------------- sparse CIN
class Catter(torch.nn.Module):
    def forward(self, x):
        return torch.cat([as_tensor(v) for v in x], dim=-1)


def take_first(xs):
    """Default upper message without coboundaries: the neighbour's features (reference `:295`)."""
    return xs[0]


def identity(x):
    """Default boundary message (reference `:299`)."""
    return x


class SparseCINCochainConv(CochainMessagePassing):
    """CIN cochain layer over boundaries and upper-adjacent cells (reference `mp/layers.py:154-214`):
        u = SUM_up msg_up(x_j, y_cob) + (1+eps1) x ;  b = SUM_bnd msg_b(x_{d-1,j}) + (1+eps2) x
        out = combine_nn(cat[update_up_nn(u), update_boundaries_nn(b)])
    """

    def __init__(self, dim: int, up_msg_size: int, down_msg_size: int, boundary_msg_size: Optional[int],
                 msg_up_nn: Callable, msg_boundaries_nn: Callable, update_up_nn: Callable,
                 update_boundaries_nn: Callable, combine_nn: Callable, eps: float = 0.,
                 train_eps: bool = False):
        super(SparseCINCochainConv, self).__init__(up_msg_size, down_msg_size,
                                                   boundary_msg_size=boundary_msg_size, use_down_msg=False)
        self.dim = dim
        self.msg_up_nn = msg_up_nn
        self.msg_boundaries_nn = msg_boundaries_nn
        self.update_up_nn = update_up_nn
        self.update_boundaries_nn = update_boundaries_nn
        self.combine_nn = combine_nn
        self.initial_eps = eps
        if train_eps:
            self.eps1 = torch.nn.Parameter(torch.Tensor([eps]))
            self.eps2 = torch.nn.Parameter(torch.Tensor([eps]))
        else:
            self.register_buffer('eps1', torch.Tensor([eps]))
            self.register_buffer('eps2', torch.Tensor([eps]))
        self.reset_parameters()

    # -------------------------------------------------------------- recognition of the closed-form messages
    def _hooks_untouched(self):
        klass = type(self)
        return all(getattr(klass, name) is getattr(SparseCINCochainConv, name)
                   for name in ('message_up', 'message_boundary', 'aggregate_up', 'aggregate_boundary',
                                'update', 'propagate')) \
            and self.aggr_up == 'add' and self.aggr_boundary == 'add' and self.flow == 'source_to_target'

    def _up_message_form(self):
        """('identity',) | ('cob', Linear, act_name) | None (unrecognised -> generic hook path)."""
        nn_ = self.msg_up_nn
        if nn_ is take_first:
            return ('identity',)
        if (isinstance(nn_, Sequential) and len(nn_) == 3 and isinstance(nn_[0], Catter)
                and isinstance(nn_[1], Linear)):
            act = activation_name(nn_[2])
            if act is not None:
                return ('cob', nn_[1], act)
        return None

    def _fused_forward(self, cochain: CochainMessagePassingParams):
        """(upper branch, boundary branch) thunks running the fused kernels, residuals included; NotImplemented
        if the layer's nets / hooks are not the recognised closed forms."""
        x = cochain.x
        form = self._up_message_form()
        if form is None or self.msg_boundaries_nn is not identity or not self._hooks_untouched():
            return NotImplemented
        if not isinstance(x, Tensor) or x.dim() != 2:
            return NotImplemented
        n = x.size(0)
        up_index, up_attr = cochain.up_index, cochain.kwargs['up_attr']
        b_index, b_attr = cochain.boundary_index, cochain.kwargs['boundary_attr']
        if up_index is not None and form[0] == 'cob':
            lin = form[1]
            if not (isinstance(up_attr, LazyRows) and x.size(1) + up_attr.source.size(1) == lin.in_features):
                return NotImplemented  # dense / missing coboundary features: let the hook protocol decide

        if up_index is None and self.up_msg_size != x.size(1):
            return NotImplemented
        if b_attr is not None and b_index is None:
            return NotImplemented
        if b_attr is None and self.boundary_msg_size != x.size(1):
            return NotImplemented

        def up_branch(pq=None):  # upper adjacencies
            if up_index is None:
                return (1 + self.eps1) * x
            if form[0] == 'identity':
                return ops.gather_scatter(x, up_index, n, 'add', x_res=x, eps=self.eps1)
            _, lin, act = form
            # W [x_j ; y_cob] + b  ==  (x W1^T)[src] + (y W2^T + b)[cob]: two per-CELL GEMMs instead of one per
            # message, then a memory-bound fused pass
            if pq is None:
                fx = x.size(1)
                pq = F.linear(x, lin.weight[:, :fx]), F.linear(up_attr.source, lin.weight[:, fx:], lin.bias)
            return ops.cob_pass(pq[0], pq[1], up_index, up_attr.index, n, act=act, x_res=x, eps=self.eps1)

        # operands of the two split-weight products, so that a container can batch them over dimensions
        up_branch.split_linear = None
        if up_index is not None and form[0] == 'cob':
            fx = x.size(1)
            up_branch.split_linear = [(x, form[1].weight, 0, None), (up_attr.source, form[1].weight, fx, form[1].bias)]

        eps_b = self._boundary_eps()

        def boundary_branch():  # the pass only runs when boundary features exist (reference mp/cell_mp.py:381)
            if b_attr is not None:
                return ops.gather_scatter(as_tensor(b_attr), b_index, n, 'add', x_res=x, eps=eps_b)
            return (1 + eps_b) * x

        return up_branch, boundary_branch

    def _boundary_eps(self):
        return self.eps2

    def dense_tail(self, agg_up, agg_boundaries):
        """update nets + combine on the aggregated (residual included) messages, through the torch modules."""
        out_up, out_boundaries = run_concurrently([lambda: self.update_up_nn(agg_up),
                                                   lambda: self.update_boundaries_nn(agg_boundaries)],
                                                  agg_up.device)
        return self.combine_nn(torch.cat([out_up, out_boundaries], dim=-1))

    def forward(self, cochain: CochainMessagePassingParams):
        fused = self._fused_forward(cochain)
        if fused is NotImplemented:
            agg_up, _, agg_b = self.propagate(cochain.up_index, cochain.down_index, cochain.boundary_index,
                                              x=cochain.x, up_attr=cochain.kwargs['up_attr'],
                                              boundary_attr=cochain.kwargs['boundary_attr'])
            up_branch = lambda: agg_up + (1 + self.eps1) * cochain.x          # noqa: E731
            boundary_branch = lambda: agg_b + (1 + self.eps2) * cochain.x     # noqa: E731
        else:
            up_branch, boundary_branch = fused
        # the two branches (aggregation + update MLP) are independent until combine_nn
        out_up, out_boundaries = run_concurrently([lambda: self.update_up_nn(up_branch()),
                                                   lambda: self.update_boundaries_nn(boundary_branch())],
                                                  cochain.x.device)
        return self.combine_nn(torch.cat([out_up, out_boundaries], dim=-1))

    def reset_parameters(self):
        reset(self.msg_up_nn)
        reset(self.msg_boundaries_nn)
        reset(self.update_up_nn)
        reset(self.update_boundaries_nn)
        reset(self.combine_nn)
        self.eps1.data.fill_(self.initial_eps)
        self.eps2.data.fill_(self.initial_eps)

    def message_up(self, up_x_j: Tensor, up_attr: Tensor) -> Tensor:
        return self.msg_up_nn((up_x_j, up_attr))

    def message_boundary(self, boundary_x_j: Tensor) -> Tensor:
        return self.msg_boundaries_nn(boundary_x_j)


class SparseCINConv(_PerDimension):
    """Cellular GIN with messages from upper neighbours and boundaries, one `SparseCINCochainConv` per dimension
    with its own nets unless `passed_*` callables are supplied (reference `mp/layers.py:271-342`).

    kwargs: `layer_dim`, `hidden`, `act_module`.

    When every level carries the default nets, a layer runs as: all aggregation passes of all dimensions as
    concurrent branches, then the update/combine nets of all dimensions as ONE fused autograd node
    (`cwn_b200.fused`); otherwise each level runs on its own (torch modules for the dense nets)."""

    fuse_dense = True
    fuse_aggregation = True  # all aggregation passes of the layer as one autograd node (fused._LayerAggregate)

    def _forms(self, fused):
        """`fused.recognise` of every level, cached on the identity of the nets (replacing a net invalidates it)."""
        key = tuple(id(m) for level in self.mp_levels
                    for m in (level.update_up_nn, level.update_boundaries_nn, level.combine_nn,
                              getattr(level, 'update_down_nn', None), level.msg_up_nn))
        cached = getattr(self, '_dense_forms', None)
        if cached is None or cached[0] != key:
            cached = self._dense_forms = (key, [fused.recognise(level) for level in self.mp_levels])
        return cached[1]

    def forward(self, *cochain_params: CochainMessagePassingParams, start_to_process=0):
        assert len(cochain_params) <= self.max_dim + 1
        n = len(cochain_params)
        if self.fuse_dense and start_to_process == 0 and n > 0:
            from cwn_b200 import fused
            forms = self._forms(fused)
            if self.fuse_aggregation and all(f is not None for f in forms[:n]) and type(self) is SparseCINConv:
                # every pass of the layer (+ the split-weight products) as one autograd node: see fused._LayerAggregate
                agg = fused.layer_aggregate(self.mp_levels[:n], cochain_params)
                if agg is not NotImplemented:
                    us, bs = agg
                    if fused.applicable(forms[:n], us, bs, self.mp_levels[0].training):
                        return fused.sparse_cin_dense(forms[:n], us, bs, self.mp_levels[0].training)
                    fused.padded_batch_needs_fused_path('this SparseCINConv configuration')
                    return [self.mp_levels[d].dense_tail(us[d], bs[d]) for d in range(n)]
            branches = [self.mp_levels[d]._fused_forward(cochain_params[d]) for d in range(n)]
            if all(f is not None for f in forms[:n]) and all(b is not NotImplemented for b in branches):
                device = cochain_params[0].x.device
                # the split-weight products of every dimension's coboundary message net in one grouped launch
                reqs = [(d, branches[d][0].split_linear) for d in range(n) if branches[d][0].split_linear]
                if reqs and all(prob[0].is_cuda and prob[0].dtype == torch.float32 for _, r in reqs for prob in r):
                    prods = fused.grouped_linear([prob for _, r in reqs for prob in r])
                    for i, (d, _) in enumerate(reqs):
                        up, pq = branches[d][0], (prods[2 * i], prods[2 * i + 1])
                        branches[d] = ((lambda up=up, pq=pq: up(pq)), branches[d][1])
                aggs = run_concurrently([t for pair in branches for t in pair], device)
                us, bs = aggs[0::2], aggs[1::2]
                if fused.applicable(forms[:n], us, bs, self.mp_levels[0].training):
                    return fused.sparse_cin_dense(forms[:n], us, bs, self.mp_levels[0].training)
                fused.padded_batch_needs_fused_path('this SparseCINConv configuration')
                return [self.mp_levels[d].dense_tail(us[d], bs[d]) for d in range(n)]
            fused.padded_batch_needs_fused_path('this SparseCINConv configuration')
        return super(SparseCINConv, self).forward(*cochain_params, start_to_process=start_to_process)

    def __init__(self, up_msg_size: int, down_msg_size: int, boundary_msg_size: Optional[int],
                 passed_msg_up_nn: Optional[Callable], passed_msg_boundaries_nn: Optional[Callable],
                 passed_update_up_nn: Optional[Callable], passed_update_boundaries_nn: Optional[Callable],
                 eps: float = 0., train_eps: bool = False, max_dim: int = 2, graph_norm=BN,
                 use_coboundaries=False, **kwargs):
        super(SparseCINConv, self).__init__()
        self.max_dim = max_dim
        self.mp_levels = torch.nn.ModuleList()
        # passed nets are the SAME objects for every dimension: their BatchNorm buffers must not be updated from
        # concurrent streams (same rule as CINConv)
        self.concurrent_dims = not any(net is not None for net in (passed_msg_up_nn, passed_msg_boundaries_nn,
                                                                   passed_update_up_nn, passed_update_boundaries_nn))

        def update_mlp():
            return Sequential(Linear(kwargs['layer_dim'], kwargs['hidden']), graph_norm(kwargs['hidden']),
                              kwargs['act_module'](),
                              Linear(kwargs['hidden'], kwargs['hidden']), graph_norm(kwargs['hidden']),
                              kwargs['act_module']())

        for dim in range(max_dim + 1):
            msg_up_nn = passed_msg_up_nn
            if msg_up_nn is None:
                if use_coboundaries:
                    msg_up_nn = Sequential(Catter(), Linear(kwargs['layer_dim'] * 2, kwargs['layer_dim']),
                                           kwargs['act_module']())
                else:
                    msg_up_nn = take_first
            msg_boundaries_nn = identity if passed_msg_boundaries_nn is None else passed_msg_boundaries_nn
            update_up_nn = update_mlp() if passed_update_up_nn is None else passed_update_up_nn
            update_boundaries_nn = update_mlp() if passed_update_boundaries_nn is None \
                else passed_update_boundaries_nn
            combine_nn = Sequential(Linear(kwargs['hidden'] * 2, kwargs['hidden']), graph_norm(kwargs['hidden']),
                                    kwargs['act_module']())
            self.mp_levels.append(SparseCINCochainConv(
                dim, up_msg_size, down_msg_size, boundary_msg_size=boundary_msg_size, msg_up_nn=msg_up_nn,
                msg_boundaries_nn=msg_boundaries_nn, update_up_nn=update_up_nn,
                update_boundaries_nn=update_boundaries_nn, combine_nn=combine_nn, eps=eps, train_eps=train_eps))


# ----------------------------------------------------------------------------------------------- CIN++
class CINppCochainConv(SparseCINCochainConv):
    """CIN++ cochain layer (reference `mp/layers.py:216-260`): a third update branch for lower-adjacent messages and a
    combine over three branches. As in the reference, the layer is built with `use_down_msg=False` (it goes through
    SparseCINCochainConv's constructor, `:223-226`, `:167-168`) and its models never request lower adjacencies
    (`include_down_features=False`), so the "down" branch sees `(1 + eps2) * x` only — that behaviour is kept.
    Residual epsilons: up = eps1, down = eps2, boundaries = eps3 (`:250-252`)."""

    def __init__(self, dim: int, up_msg_size: int, down_msg_size: int, boundary_msg_size: int, msg_up_nn: Callable,
                 msg_boundaries_nn: Callable, msg_down_nn: Callable, update_up_nn: Callable,
                 update_boundaries_nn: Callable, update_down_nn: Callable, combine_nn: Callable, eps: float = 0,
                 train_eps: bool = False):
        super(CINppCochainConv, self).__init__(dim, up_msg_size, down_msg_size, boundary_msg_size, msg_up_nn,
                                               msg_boundaries_nn, update_up_nn, update_boundaries_nn, combine_nn, eps,
                                               train_eps)
        self.msg_down_nn = msg_down_nn
        self.update_down_nn = update_down_nn
        if train_eps:
            self.eps3 = torch.nn.Parameter(torch.Tensor([eps]))
        else:
            self.register_buffer('eps3', torch.Tensor([eps]))
        reset(self.msg_down_nn)
        reset(self.update_down_nn)
        self.eps3.data.fill_(self.initial_eps)

    def _boundary_eps(self):
        return self.eps3

    def message_down(self, down_x_j: Tensor, down_attr: Tensor) -> Tensor:
        return self.msg_down_nn((down_x_j, down_attr))

    def forward(self, cochain: CochainMessagePassingParams):
        x = cochain.x
        fused = self._fused_forward(cochain)
        if fused is NotImplemented:
            agg_up, agg_down, agg_b = self.propagate(cochain.up_index, cochain.down_index, cochain.boundary_index,
                                                     x=x, up_attr=cochain.kwargs['up_attr'],
                                                     boundary_attr=cochain.kwargs['boundary_attr'])
            up_branch = lambda: agg_up + (1 + self.eps1) * x              # noqa: E731
            down_branch = lambda: agg_down + (1 + self.eps2) * x          # noqa: E731
            boundary_branch = lambda: agg_b + (1 + self.eps3) * x         # noqa: E731
        else:
            up_branch, boundary_branch = fused
            if self.down_msg_size != x.size(1):  # zeros [N, down_msg_size] + (1+eps) x must broadcast as torch would
                down_branch = lambda: torch.zeros(x.size(0), self.down_msg_size, device=x.device) + (1 + self.eps2) * x  # noqa: E731
            else:
                down_branch = lambda: (1 + self.eps2) * x                 # noqa: E731
        out_up, out_down, out_boundaries = run_concurrently(
            [lambda: self.update_up_nn(up_branch()), lambda: self.update_down_nn(down_branch()),
             lambda: self.update_boundaries_nn(boundary_branch())], x.device)
        return self.combine_nn(torch.cat([out_up, out_down, out_boundaries], dim=-1))


class CINppConv(SparseCINConv):
    """One `CINppCochainConv` per dimension (reference `mp/layers.py:344-427`). Like the reference it first runs
    SparseCINConv's constructor (whose levels are then discarded), so a seeded construction consumes the random
    stream identically."""

    fuse_dense = True  # the three update branches + combine as one fused autograd node (cwn_b200.fused)

    def forward(self, *cochain_params: CochainMessagePassingParams, start_to_process=0):
        """All aggregation passes of the layer as one node (`fused._LayerAggregate`, boundary epsilon = eps3), the down
        branch's input `(1 + eps2) x` (no lower messages are ever propagated: see CINppCochainConv), then the three
        update MLPs and the 3-block combine of every dimension as ONE fused dense node; anything outside that closed
        form runs level by level through the torch modules."""
        n = len(cochain_params)
        assert n <= self.max_dim + 1
        if self.fuse_dense and self.fuse_aggregation and start_to_process == 0 and n > 0:
            from cwn_b200 import fused
            forms = self._forms(fused)
            if all(f is not None and f[3] is not None for f in forms[:n]) and \
                    all(self.mp_levels[d].down_msg_size == cochain_params[d].x.size(1) for d in range(n)
                        if isinstance(cochain_params[d].x, Tensor) and cochain_params[d].x.dim() == 2):
                agg = fused.layer_aggregate(self.mp_levels[:n], cochain_params)
                if agg is not NotImplemented:
                    us, bs = agg
                    ds = [(1 + self.mp_levels[d].eps2) * cochain_params[d].x for d in range(n)]
                    if fused.applicable(forms[:n], us, bs, self.mp_levels[0].training, ds):
                        return fused.sparse_cin_dense(forms[:n], us, bs, self.mp_levels[0].training, ds)
                    levels = self.mp_levels
                    outs = []
                    for d in range(n):  # aggregated already: finish through the torch modules
                        ou, od, ob = levels[d].update_up_nn(us[d]), levels[d].update_down_nn(ds[d]), levels[d].update_boundaries_nn(bs[d])
                        outs.append(levels[d].combine_nn(torch.cat([ou, od, ob], dim=-1)))
                    return outs
        return _PerDimension.forward(self, *cochain_params, start_to_process=start_to_process)

    def __init__(self, up_msg_size: int, down_msg_size: int, boundary_msg_size: Optional[int],
                 passed_msg_up_nn: Optional[Callable], passed_msg_down_nn: Optional[Callable],
                 passed_msg_boundaries_nn: Optional[Callable], passed_update_up_nn: Optional[Callable],
                 passed_update_down_nn: Optional[Callable], passed_update_boundaries_nn: Optional[Callable],
                 eps: float = 0., train_eps: bool = False, max_dim: int = 2, graph_norm=BN, use_coboundaries=False,
                 **kwargs):
        super(CINppConv, self).__init__(up_msg_size, down_msg_size, boundary_msg_size, passed_msg_up_nn,
                                        passed_msg_boundaries_nn, passed_update_up_nn, passed_update_boundaries_nn,
                                        eps, train_eps, max_dim, graph_norm, use_coboundaries, **kwargs)
        self.max_dim = max_dim
        self.mp_levels = torch.nn.ModuleList()
        self.concurrent_dims = not any(net is not None for net in (
            passed_msg_up_nn, passed_msg_down_nn, passed_msg_boundaries_nn, passed_update_up_nn, passed_update_down_nn,
            passed_update_boundaries_nn))

        def message_mlp():
            return Sequential(Catter(), Linear(kwargs['layer_dim'] * 2, kwargs['layer_dim']), kwargs['act_module']())

        def update_mlp():
            return Sequential(Linear(kwargs['layer_dim'], kwargs['hidden']), graph_norm(kwargs['hidden']),
                              kwargs['act_module'](),
                              Linear(kwargs['hidden'], kwargs['hidden']), graph_norm(kwargs['hidden']),
                              kwargs['act_module']())

        for dim in range(max_dim + 1):
            msg_up_nn = passed_msg_up_nn
            if msg_up_nn is None:
                msg_up_nn = message_mlp() if use_coboundaries else take_first
            msg_down_nn = passed_msg_down_nn
            if msg_down_nn is None:
                msg_down_nn = message_mlp() if use_coboundaries else take_first
            msg_boundaries_nn = identity if passed_msg_boundaries_nn is None else passed_msg_boundaries_nn
            update_up_nn = update_mlp() if passed_update_up_nn is None else passed_update_up_nn
            update_down_nn = update_mlp() if passed_update_down_nn is None else passed_update_down_nn
            update_boundaries_nn = update_mlp() if passed_update_boundaries_nn is None \
                else passed_update_boundaries_nn
            combine_nn = Sequential(Linear(kwargs['hidden'] * 3, kwargs['hidden']), graph_norm(kwargs['hidden']),
                                    kwargs['act_module']())
            self.mp_levels.append(CINppCochainConv(
                dim, up_msg_size, down_msg_size, boundary_msg_size=boundary_msg_size, msg_up_nn=msg_up_nn,
                msg_down_nn=msg_down_nn, msg_boundaries_nn=msg_boundaries_nn, update_up_nn=update_up_nn,
                update_down_nn=update_down_nn, update_boundaries_nn=update_boundaries_nn, combine_nn=combine_nn,
                eps=eps, train_eps=train_eps))


# ----------------------------------------------------------------------------------------------- initialisation
class InitReduceConv(torch.nn.Module):
    """Initial features of d-cells as a reduction of their boundary features (reference `mp/layers.py:473-487`).

    `out_size` (host int) avoids the reference's `boundary_index[1].max() + 1` device sync; when omitted the
    reference behaviour (sync) is kept."""

    def __init__(self, reduce='add'):
        super(InitReduceConv, self).__init__()
        if reduce not in ops.REDUCE_CODES:
            raise NotImplementedError(f'cwn_b200: InitReduceConv reduce={reduce!r} is not supported')
        self.reduce = reduce

    def forward(self, boundary_x, boundary_index, out_size: Optional[int] = None):
        if out_size is None:
            out_size = int(boundary_index[1, :].max()) + 1
        return ops.gather_scatter(boundary_x, boundary_index, out_size, reduce=self.reduce)


class AbstractEmbedVEWithReduce(torch.nn.Module, ABC):
    """Embeds vertex (and optionally edge) integer features and initialises the features of higher cells by
    reducing over boundaries; rings are reduced from the REDUCED edge features and halved
    (reference `mp/layers.py:490-547`)."""

    def __init__(self, v_embed_layer: Callable, e_embed_layer: Optional[Callable], init_reduce: InitReduceConv):
        super(AbstractEmbedVEWithReduce, self).__init__()
        self.v_embed_layer = v_embed_layer
        self.e_embed_layer = e_embed_layer
        self.init_reduce = init_reduce

    @abstractmethod
    def _prepare_v_inputs(self, v_params):
        pass

    @abstractmethod
    def _prepare_e_inputs(self, e_params):
        pass

    def forward(self, *cochain_params: CochainMessagePassingParams):
        assert 1 <= len(cochain_params) <= 3
        v_params = cochain_params[0]
        e_params = cochain_params[1] if len(cochain_params) >= 2 else None
        c_params = cochain_params[2] if len(cochain_params) == 3 else None

        vx = self._embed(self.v_embed_layer, self._prepare_v_inputs(v_params))
        out = [vx]
        if e_params is None:
            assert c_params is None
            return out

        reduced_ex = self.init_reduce(vx, e_params.boundary_index, getattr(e_params, 'num_cells', None))
        ex = reduced_ex
        if e_params.x is not None:
            ex = self._embed(self.e_embed_layer, self._prepare_e_inputs(e_params))
            assert ex.size(1) == vx.size(1)
        out.append(ex)

        if c_params is not None:
            cx = self.init_reduce(reduced_ex, c_params.boundary_index, getattr(c_params, 'num_cells', None)) / 2.
            out.append(cx)
        return out

    @staticmethod
    def _embed(layer, inputs):
        """A plain `nn.Embedding` lookup is a row gather: run it (and its gradient, a segmented reduction over a
        CSR plan of the type ids instead of torch's sort-based embedding backward) on the cwn kernels."""
        if (type(layer) is torch.nn.Embedding and layer.padding_idx is None and layer.max_norm is None
                and not layer.sparse and not layer.scale_grad_by_freq and inputs.dim() == 1 and inputs.is_cuda
                and layer.weight.dtype == torch.float32):
            return ops.gather_rows(layer.weight, inputs)
        return layer(inputs)

    def reset_parameters(self):
        reset(self.v_embed_layer)
        reset(self.e_embed_layer)


class EmbedVEWithReduce(AbstractEmbedVEWithReduce):
    """`nn.Embedding` over scalar integer features stored as `[N, 1]` floats (ZINC; reference `:550-570`)."""

    def __init__(self, v_embed_layer: torch.nn.Embedding, e_embed_layer: Optional[torch.nn.Embedding],
                 init_reduce: InitReduceConv):
        super(EmbedVEWithReduce, self).__init__(v_embed_layer, e_embed_layer, init_reduce)

    def _prepare_v_inputs(self, v_params):
        assert v_params.x is not None
        assert v_params.x.dim() == 2
        assert v_params.x.size(1) == 1
        return v_params.x.squeeze(1).to(dtype=torch.long)

    def _prepare_e_inputs(self, e_params):
        assert self.e_embed_layer is not None
        assert e_params.x.dim() == 2
        assert e_params.x.size(1) == 1
        return e_params.x.squeeze(1).to(dtype=torch.long)


class OGBEmbedVEWithReduce(AbstractEmbedVEWithReduce):
    """OGB Atom/Bond encoders over multi-column integer features (reference `:573-593`)."""

    def __init__(self, v_embed_layer, e_embed_layer, init_reduce: InitReduceConv):
        super(OGBEmbedVEWithReduce, self).__init__(v_embed_layer, e_embed_layer, init_reduce)

    def _prepare_v_inputs(self, v_params):
        assert v_params.x is not None
        assert v_params.x.dim() == 2
        return v_params.x.to(dtype=torch.long)

    def _prepare_e_inputs(self, e_params):
        assert self.e_embed_layer is not None
        assert e_params.x.dim() == 2
        return e_params.x.to(dtype=torch.long)
