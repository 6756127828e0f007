"""`SparseCIN` and `CIN0` — the float-feature models of the hot path (reference `mp/models.py:12-257`).
Constructor signatures, module names (state_dict keys) and forward semantics follow the reference; message
passing, readout and their gradients run in the cwn_b200 CUDA kernels."""
import torch
import torch.nn.functional as F
from torch.nn import Linear, Sequential, BatchNorm1d as BN

from cwn_b200 import ops
from cwn_b200.data.complex import ComplexBatch
from cwn_b200.mp.layers import CINConv, CINppConv, EdgeCINConv, OrientedConv, SparseCINConv
from cwn_b200.mp.nn import (JumpingKnowledge, get_graph_norm, get_nonlinearity, get_pooling_fn,
                            num_complexes_of, pool_complex)


def _readout_head(model, xs, act, res, include_partial):
    """Shared tail of the SparseCIN-family forwards (reference `mp/models.py:230-254`,
    `mp/molec_models.py:137-161`): per-dimension lin1 + act, sum/mean over dimensions, dropout, lin2."""
    new_xs = []
    for i, x in enumerate(xs):
        if model.apply_dropout_before == 'lin1':
            x = F.dropout(x, p=model.dropout_rate, training=model.training)
        new_xs.append(act(model.lin1s[model.readout_dims[i]](x)))
    x = torch.stack(new_xs, dim=0)
    if model.apply_dropout_before == 'final_readout':
        x = F.dropout(x, p=model.dropout_rate, training=model.training)
    if model.final_readout == 'mean':
        x = x.mean(0)
    elif model.final_readout == 'sum':
        x = x.sum(0)
    else:
        raise NotImplementedError
    if model.apply_dropout_before not in ['lin1', 'final_readout']:
        x = F.dropout(x, p=model.dropout_rate, training=model.training)
    x = model.lin2(x)
    if include_partial:
        res['out'] = x
        return x, res
    return x


class _JumpMixin(object):
    def jump_complex(self, jump_xs):
        return [self.jump(jumpx) for jumpx in jump_xs]

    def _remember_layer(self, history, xs):
        """Per-dimension list of every layer's output, for JumpingKnowledge (None when the model does not jump)."""
        if self.jump_mode is None:
            return None
        history = history if history is not None else [[] for _ in xs]
        for per_dim, x in zip(history, xs):
            per_dim.append(x)
        return history

    def _summed_head(self, xs, data, history):
        """Readout shared by the dense models: jump, per-complex pooling, sum over dimensions, lin1 + act, dropout,
        lin2 (reference `mp/models.py:95-106, 404-416`)."""
        if history is not None:
            xs = self.jump_complex(history)
        x = self.pool_complex(xs, data).sum(dim=0)
        x = get_nonlinearity(self.nonlinearity, return_module=False)(self.lin1(x))
        return self.lin2(F.dropout(x, p=self.dropout_rate, training=self.training))


class CIN0(torch.nn.Module, _JumpMixin):
    """Dense cellular GIN: upper + lower messages through shared MLPs (reference `mp/models.py:12-109`)."""

    def __init__(self, num_input_features, num_classes, num_layers, hidden, dropout_rate: float = 0.5,
                 max_dim: int = 2, jump_mode=None, nonlinearity='relu', readout='sum'):
        super(CIN0, self).__init__()
        self.max_dim = max_dim
        self.dropout_rate = dropout_rate
        self.jump_mode = jump_mode
        self.convs = torch.nn.ModuleList()
        self.nonlinearity = nonlinearity
        self.readout = readout
        self.pooling_fn = get_pooling_fn(readout)
        conv_nonlinearity = get_nonlinearity(nonlinearity, return_module=True)
        for i in range(num_layers):
            layer_dim = num_input_features if i == 0 else hidden
            conv_update = Sequential(Linear(layer_dim, hidden), conv_nonlinearity(), Linear(hidden, hidden),
                                     conv_nonlinearity(), BN(hidden))
            conv_up = Sequential(Linear(layer_dim * 2, layer_dim), conv_nonlinearity(), BN(layer_dim))
            conv_down = Sequential(Linear(layer_dim * 2, layer_dim), conv_nonlinearity(), BN(layer_dim))
            self.convs.append(CINConv(layer_dim, layer_dim, conv_up, conv_down, conv_update, train_eps=False,
                                      max_dim=self.max_dim))
        self.jump = JumpingKnowledge(jump_mode) if jump_mode is not None else None
        self.lin1 = Linear(num_layers * hidden if jump_mode == 'cat' else hidden, hidden)
        self.lin2 = Linear(hidden, num_classes)

    def reset_parameters(self):
        for conv in self.convs:
            conv.reset_parameters()
        if self.jump_mode is not None:
            self.jump.reset_parameters()
        self.lin1.reset_parameters()
        self.lin2.reset_parameters()

    def pool_complex(self, xs, data):
        return pool_complex(xs, data, self.max_dim, self.readout)

    def forward(self, data: ComplexBatch):
        xs, history = None, None
        for conv in self.convs:
            xs = conv(*data.get_all_cochain_params(max_dim=self.max_dim))
            data.set_xs(xs)
            history = self._remember_layer(history, xs)
        return self._summed_head(xs, data, history)

    def __repr__(self):
        return self.__class__.__name__


class SparseCIN(torch.nn.Module, _JumpMixin):
    """Cellular GIN over boundaries and upper adjacencies (reference `mp/models.py:112-257`)."""

    def __init__(self, num_input_features, num_classes, num_layers, hidden, dropout_rate: float = 0.5,
                 max_dim: int = 2, jump_mode=None, nonlinearity='relu', readout='sum', train_eps=False,
                 final_hidden_multiplier: int = 2, use_coboundaries=False, readout_dims=(0, 1, 2),
                 final_readout='sum', apply_dropout_before='lin2', graph_norm='bn'):
        super(SparseCIN, self).__init__()
        self.max_dim = max_dim
        if readout_dims is not None:
            self.readout_dims = tuple([dim for dim in readout_dims if dim <= max_dim])
        else:
            self.readout_dims = list(range(max_dim + 1))
        self.final_readout = final_readout
        self.dropout_rate = dropout_rate
        self.apply_dropout_before = apply_dropout_before
        self.jump_mode = jump_mode
        self.convs = torch.nn.ModuleList()
        self.nonlinearity = nonlinearity
        self.readout = readout
        self.use_coboundaries = use_coboundaries
        self.pooling_fn = get_pooling_fn(readout)
        self.graph_norm = get_graph_norm(graph_norm)
        act_module = get_nonlinearity(nonlinearity, return_module=True)
        for i in range(num_layers):
            layer_dim = num_input_features if i == 0 else hidden
            self.convs.append(
                SparseCINConv(up_msg_size=layer_dim, down_msg_size=layer_dim, boundary_msg_size=layer_dim,
                              passed_msg_boundaries_nn=None, passed_msg_up_nn=None, passed_update_up_nn=None,
                              passed_update_boundaries_nn=None, train_eps=train_eps, max_dim=self.max_dim,
                              hidden=hidden, act_module=act_module, layer_dim=layer_dim,
                              graph_norm=self.graph_norm, use_coboundaries=use_coboundaries))
        self.jump = JumpingKnowledge(jump_mode) if jump_mode is not None else None
        self.lin1s = torch.nn.ModuleList()
        for _ in range(max_dim + 1):
            if jump_mode == 'cat':
                # bias-free: a dimension absent from a complex must contribute exactly zero
                self.lin1s.append(Linear(num_layers * hidden, final_hidden_multiplier * hidden, bias=False))
            else:
                self.lin1s.append(Linear(hidden, final_hidden_multiplier * hidden))
        self.lin2 = Linear(final_hidden_multiplier * hidden, num_classes)

    def reset_parameters(self):
        for conv in self.convs:
            conv.reset_parameters()
        if self.jump_mode is not None:
            self.jump.reset_parameters()
        self.lin1s.reset_parameters()
        self.lin2.reset_parameters()

    fuse_readout = True  # pool + lin1s + act + lin2 as one kernel (cwn_b200.fused.readout_head) when its closed form applies

    def pool_complex(self, xs, data):
        pooled = pool_complex(xs, data, self.max_dim, self.readout)
        return [pooled[i] for i in range(self.max_dim + 1)]

    def forward(self, data: ComplexBatch, include_partial=False):
        act = get_nonlinearity(self.nonlinearity, return_module=False)
        xs, jump_xs = None, None
        res = {}
        # every CSR plan of this batch (all dimensions, forward + backward, readout) in one kernel launch
        ops.prepare_plans(data, self.max_dim, self.use_coboundaries, backward=torch.is_grad_enabled())
        for c, conv in enumerate(self.convs):
            params = data.get_all_cochain_params(max_dim=self.max_dim, include_down_features=False)
            xs = conv(*params, start_to_process=0)
            data.set_xs(xs)
            if include_partial:
                for k in range(len(xs)):
                    res[f"layer{c}_{k}"] = xs[k]
            if self.jump_mode is not None:
                if jump_xs is None:
                    jump_xs = [[] for _ in xs]
                for i, x in enumerate(xs):
                    jump_xs[i] += [x]
        if self.jump_mode is not None:
            xs = self.jump_complex(jump_xs)
        if not include_partial and self.fuse_readout:
            from cwn_b200 import fused
            out = fused.readout_head(self, xs, data, self.nonlinearity if self.nonlinearity in ops.ACT_CODES else None)
            if out is not NotImplemented:
                return out
        xs = self.pool_complex(xs, data)
        xs = [xs[i] for i in self.readout_dims]
        if include_partial:
            for k in range(len(xs)):
                res[f"pool_{k}"] = xs[k]
        return _readout_head(self, xs, act, res, include_partial)

    def __repr__(self):
        return self.__class__.__name__


class CINpp(SparseCIN):
    """CIN++ on float features (reference `mp/models.py:259-283`): SparseCIN whose layers are `CINppConv`."""

    def __init__(self, num_input_features, num_classes, num_layers, hidden, dropout_rate: float = 0.5,
                 max_dim: int = 2, jump_mode=None, nonlinearity='relu', readout='sum', train_eps=False,
                 final_hidden_multiplier: int = 2, use_coboundaries=False, readout_dims=(0, 1, 2),
                 final_readout='sum', apply_dropout_before='lin2', graph_norm='bn'):
        super(CINpp, self).__init__(num_input_features, num_classes, num_layers, hidden, dropout_rate, max_dim,
                                    jump_mode, nonlinearity, readout, train_eps, final_hidden_multiplier,
                                    use_coboundaries, readout_dims, final_readout, apply_dropout_before, graph_norm)
        self.convs = torch.nn.ModuleList()
        act_module = get_nonlinearity(nonlinearity, return_module=True)
        for i in range(num_layers):
            layer_dim = num_input_features if i == 0 else hidden
            self.convs.append(
                CINppConv(up_msg_size=layer_dim, down_msg_size=layer_dim, boundary_msg_size=layer_dim,
                          passed_msg_boundaries_nn=None, passed_msg_up_nn=None, passed_msg_down_nn=None,
                          passed_update_up_nn=None, passed_update_down_nn=None, passed_update_boundaries_nn=None,
                          train_eps=train_eps, max_dim=self.max_dim, hidden=hidden, act_module=act_module,
                          layer_dim=layer_dim, graph_norm=self.graph_norm, use_coboundaries=use_coboundaries))


class EdgeCIN0(torch.nn.Module, _JumpMixin):
    """CIN0 operating up to the edges; two-cell features may feed the edges' upper messages and be refreshed by their
    own MLP between layers (reference `mp/models.py:286-419`)."""

    def __init__(self, num_input_features, num_classes, num_layers, hidden, dropout_rate: float = 0.5, jump_mode=None,
                 nonlinearity='relu', include_top_features=True, update_top_features=True, readout='sum'):
        super(EdgeCIN0, self).__init__()
        self.max_dim = 1
        self.include_top_features = include_top_features
        self.update_top_features = include_top_features and update_top_features
        self.dropout_rate = dropout_rate
        self.jump_mode = jump_mode
        self.convs = torch.nn.ModuleList()
        self.update_top_nns = torch.nn.ModuleList()
        self.nonlinearity = nonlinearity
        self.readout = readout
        self.pooling_fn = get_pooling_fn(readout)
        act = get_nonlinearity(nonlinearity, return_module=True)

        def update_mlp(layer_dim):
            return Sequential(Linear(layer_dim, hidden), act(), Linear(hidden, hidden), act(), BN(hidden))

        for i in range(num_layers):
            layer_dim = num_input_features if i == 0 else hidden
            v_conv_update, e_conv_update = update_mlp(layer_dim), update_mlp(layer_dim)
            v_conv_up = Sequential(Linear(layer_dim * 2, layer_dim), act(), BN(layer_dim))
            e_conv_down = Sequential(Linear(layer_dim * 2, layer_dim), act(), BN(layer_dim))
            e_conv_inp_dim = layer_dim * 2 if include_top_features else layer_dim
            e_conv_up = Sequential(Linear(e_conv_inp_dim, layer_dim), act(), BN(layer_dim))
            self.convs.append(EdgeCINConv(layer_dim, layer_dim, v_conv_up, e_conv_down, e_conv_up, v_conv_update,
                                          e_conv_update, train_eps=False))
            if self.update_top_features and i < num_layers - 1:
                self.update_top_nns.append(update_mlp(layer_dim))
        self.jump = JumpingKnowledge(jump_mode) if jump_mode is not None else None
        self.lin1 = Linear(num_layers * hidden if jump_mode == 'cat' else hidden, hidden)
        self.lin2 = Linear(hidden, num_classes)

    def reset_parameters(self):
        for conv in self.convs:
            conv.reset_parameters()
        if self.jump_mode is not None:
            self.jump.reset_parameters()
        self.lin1.reset_parameters()
        self.lin2.reset_parameters()
        for net in self.update_top_nns:
            for m in net:
                if hasattr(m, 'reset_parameters'):
                    m.reset_parameters()

    def pool_complex(self, xs, data):
        return pool_complex(xs, data, self.max_dim, self.readout)

    def forward(self, data: ComplexBatch):
        xs, history = None, None
        last = len(self.convs) - 1
        for c, conv in enumerate(self.convs):
            xs = conv(*data.get_all_cochain_params(max_dim=self.max_dim,
                                                   include_top_features=self.include_top_features))
            refreshed = list(xs)
            # between layers the two-cell features (inputs of the edges' upper messages) go through their own MLP,
            # provided the batch has two-cells at all
            if self.update_top_features and c < last and 2 in data.cochains:
                refreshed.append(self.update_top_nns[c](data.cochains[2].x))
            data.set_xs(refreshed)
            history = self._remember_layer(history, xs)
        return self._summed_head(xs, data, history)

    def __repr__(self):
        return self.__class__.__name__


class _OrientedEdgeModel(torch.nn.Module):
    """Shared body of `EdgeOrient` / `EdgeMPNN` (reference `mp/models.py:474-608`): a stack of `OrientedConv` on ONE
    cochain batch of edges, |.| for orientation invariance, per-complex readout, lin1 + ReLU, lin2."""

    def _build(self, num_input_features, num_classes, num_layers, hidden, dropout_rate, jump_mode, nonlinearity,
               readout, fully_invar, with_up):
        self.max_dim = 1
        self.fully_invar = fully_invar
        orient = not self.fully_invar
        self.dropout_rate = dropout_rate
        self.jump_mode = jump_mode
        self.convs = torch.nn.ModuleList()
        self.nonlinearity = nonlinearity
        self.readout = readout
        self.pooling_fn = get_pooling_fn(readout)
        for i in range(num_layers):
            layer_dim = num_input_features if i == 0 else hidden
            # biases must stay off: with them the layer is not orientation-equivariant (reference :489)
            update_up = Linear(layer_dim, hidden, bias=False) if with_up else (lambda x: 0)
            update_down = Linear(layer_dim, hidden, bias=False)
            update = Linear(layer_dim, hidden, bias=False)
            self.convs.append(OrientedConv(dim=1, up_msg_size=layer_dim, down_msg_size=layer_dim,
                                           update_up_nn=update_up, update_down_nn=update_down, update_nn=update,
                                           act_fn=get_nonlinearity(nonlinearity, return_module=False), orient=orient))
        self.jump = JumpingKnowledge(jump_mode) if jump_mode is not None else None
        self.lin1 = Linear(hidden, hidden)
        self.lin2 = Linear(hidden, num_classes)

    def reset_parameters(self):
        for conv in self.convs:
            conv.reset_parameters()
        if self.jump_mode is not None:
            self.jump.reset_parameters()
        self.lin1.reset_parameters()
        self.lin2.reset_parameters()

    def forward(self, data, include_partial=False):
        if self.fully_invar:
            data.x = torch.abs(data.x)
        x = None
        for conv in self.convs:
            x = conv(data)
            data.x = x
        cell_pred = x
        batch_size = getattr(data, 'num_cochains', None)
        if batch_size is None:
            batch_size = int(data.batch.max()) + 1
        if not self.fully_invar:
            x = torch.abs(x)
        x = self.pooling_fn(x, data.batch, size=batch_size)
        # invariance holds from here on: any non-linearity will do; the reference picks ReLU
        x = torch.relu(self.lin1(x))
        x = F.dropout(x, p=self.dropout_rate, training=self.training)
        x = self.lin2(x)
        if include_partial:
            return x, cell_pred
        return x

    def __repr__(self):
        return self.__class__.__name__


class EdgeOrient(_OrientedEdgeModel):
    """Edge-signal model that takes edge orientation into account (reference `mp/models.py:474-545`)."""

    def __init__(self, num_input_features, num_classes, num_layers, hidden, dropout_rate: float = 0.0, jump_mode=None,
                 nonlinearity='id', readout='sum', fully_invar=False):
        super(EdgeOrient, self).__init__()
        self._build(num_input_features, num_classes, num_layers, hidden, dropout_rate, jump_mode, nonlinearity, readout,
                    fully_invar, with_up=True)


class EdgeMPNN(_OrientedEdgeModel):
    """MPNN on the line graph: lower adjacencies only (reference `mp/models.py:548-608`)."""

    def __init__(self, num_input_features, num_classes, num_layers, hidden, dropout_rate: float = 0.0, jump_mode=None,
                 nonlinearity='relu', readout='sum', fully_invar=True):
        super(EdgeMPNN, self).__init__()
        self._build(num_input_features, num_classes, num_layers, hidden, dropout_rate, jump_mode, nonlinearity, readout,
                    fully_invar, with_up=False)
