"""`EmbedSparseCIN` (ZINC) and `OGBEmbedSparseCIN` (ogbg-mol*) — reference `mp/molec_models.py:12-164, 201-353`."""
import torch
import torch.nn.functional as F
from torch.nn import Embedding, Linear

from cwn_b200 import ops
from cwn_b200.data.complex import ComplexBatch
from cwn_b200.mp.encoders import AtomEncoder, BondEncoder
from cwn_b200.mp.layers import (CINppConv, EmbedVEWithReduce, InitReduceConv, OGBEmbedVEWithReduce,
                                SparseCINConv)
from cwn_b200.mp.models import _JumpMixin, _readout_head
from cwn_b200.mp.nn import JumpingKnowledge, get_graph_norm, get_nonlinearity, pool_complex


class _EmbedSparseCINBase(torch.nn.Module, _JumpMixin):
    """Everything the two molecular models share: conv stack, jump, per-dimension readout head."""

    fuse_readout = True  # pool + lin1s + act + lin2 as one kernel (cwn_b200.fused.readout_head) when its closed form applies

    def _build_trunk(self, out_size, num_layers, hidden, max_dim, jump_mode, nonlinearity, readout, train_eps,
                     final_hidden_multiplier, readout_dims, final_readout, apply_dropout_before, embed_dim,
                     use_coboundaries, graph_norm):
        self.final_readout = final_readout
        self.apply_dropout_before = apply_dropout_before
        self.jump_mode = jump_mode
        self.convs = torch.nn.ModuleList()
        self.nonlinearity = nonlinearity
        self.readout = readout
        self.use_coboundaries = use_coboundaries
        self.graph_norm = get_graph_norm(graph_norm)
        act_module = get_nonlinearity(nonlinearity, return_module=True)
        for i in range(num_layers):
            layer_dim = embed_dim if i == 0 else hidden
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
                self.lin1s.append(Linear(num_layers * hidden, final_hidden_multiplier * hidden, bias=False))
            else:
                self.lin1s.append(Linear(hidden, final_hidden_multiplier * hidden))
        self.lin2 = Linear(final_hidden_multiplier * hidden, out_size)

    @staticmethod
    def _readout_dims(readout_dims, max_dim):
        if readout_dims is not None:
            return tuple([dim for dim in readout_dims if dim <= max_dim])
        return list(range(max_dim + 1))

    def reset_parameters(self):
        for conv in self.convs:
            conv.reset_parameters()
        if self.jump_mode is not None:
            self.jump.reset_parameters()
        self.init_conv.reset_parameters()
        self.lin1s.reset_parameters()
        self.lin2.reset_parameters()

    def _forward(self, data, include_partial, in_dropout, conv_dropout):
        act = get_nonlinearity(self.nonlinearity, return_module=False)
        xs, jump_xs = None, None
        res = {}
        # every CSR plan of this batch (all dimensions, forward + backward, readout) in one kernel launch
        ops.prepare_plans(data, self.max_dim, self.use_coboundaries, backward=torch.is_grad_enabled())
        # embed vertices (+edges) and populate the higher dimensions by boundary reduction
        params = data.get_all_cochain_params(max_dim=self.max_dim, include_down_features=False)
        xs = list(self.init_conv(*params))
        for i, x in enumerate(xs):
            xs[i] = F.dropout(xs[i], p=in_dropout, training=self.training)
        data.set_xs(xs)

        for c, conv in enumerate(self.convs):
            params = data.get_all_cochain_params(max_dim=self.max_dim, include_down_features=False)
            xs = conv(*params, start_to_process=0)
            if conv_dropout is not None:
                for i, x in enumerate(xs):
                    xs[i] = F.dropout(xs[i], p=conv_dropout, training=self.training)
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
        xs = pool_complex(xs, data, self.max_dim, self.readout)
        xs = [xs[i] for i in self.readout_dims]
        if include_partial:
            for k in range(len(xs)):
                res[f"pool_{k}"] = xs[k]
        return _readout_head(self, xs, act, res, include_partial)

    def __repr__(self):
        return self.__class__.__name__


class EmbedSparseCIN(_EmbedSparseCINBase):
    """Cellular GIN for ZINC-style molecules with integer atom / bond types (reference `:12-164`)."""

    def __init__(self, atom_types, bond_types, out_size, num_layers, hidden, dropout_rate: float = 0.5,
                 max_dim: int = 2, jump_mode=None, nonlinearity='relu', readout='sum', train_eps=False,
                 final_hidden_multiplier: int = 2, readout_dims=(0, 1, 2), final_readout='sum',
                 apply_dropout_before='lin2', init_reduce='sum', embed_edge=False, embed_dim=None,
                 use_coboundaries=False, graph_norm='bn'):
        super(EmbedSparseCIN, self).__init__()
        self.max_dim = max_dim
        self.readout_dims = self._readout_dims(readout_dims, max_dim)
        if embed_dim is None:
            embed_dim = hidden
        self.v_embed_init = Embedding(atom_types, embed_dim)
        self.e_embed_init = None
        if embed_edge:
            self.e_embed_init = Embedding(bond_types, embed_dim)
        self.reduce_init = InitReduceConv(reduce=init_reduce)
        self.init_conv = EmbedVEWithReduce(self.v_embed_init, self.e_embed_init, self.reduce_init)
        self.dropout_rate = dropout_rate
        self._build_trunk(out_size, num_layers, hidden, max_dim, jump_mode, nonlinearity, readout, train_eps,
                          final_hidden_multiplier, readout_dims, final_readout, apply_dropout_before, embed_dim,
                          use_coboundaries, graph_norm)

    def forward(self, data: ComplexBatch, include_partial=False):
        # input node/edge features are scalars (integer types stored as floats)
        assert data.cochains[0].x.size(-1) == 1
        if 1 in data.cochains and data.cochains[1].x is not None:
            assert data.cochains[1].x.size(-1) == 1
        return self._forward(data, include_partial, in_dropout=self.dropout_rate, conv_dropout=None)


class OGBEmbedSparseCIN(_EmbedSparseCINBase):
    """Cellular GIN for ogbg-mol* with OGB atom / bond encoders (reference `:201-353`). `atom_feature_dims` /
    `bond_feature_dims` (extension) override the ogb vocabularies."""

    def __init__(self, out_size, num_layers, hidden, dropout_rate: float = 0.5, indropout_rate: float = 0.0,
                 max_dim: int = 2, jump_mode=None, nonlinearity='relu', readout='sum', train_eps=False,
                 final_hidden_multiplier: int = 2, readout_dims=(0, 1, 2), final_readout='sum',
                 apply_dropout_before='lin2', init_reduce='sum', embed_edge=False, embed_dim=None,
                 use_coboundaries=False, graph_norm='bn', atom_feature_dims=None, bond_feature_dims=None):
        super(OGBEmbedSparseCIN, self).__init__()
        self.max_dim = max_dim
        self.readout_dims = self._readout_dims(readout_dims, max_dim)
        if embed_dim is None:
            embed_dim = hidden
        self.v_embed_init = AtomEncoder(embed_dim) if atom_feature_dims is None \
            else AtomEncoder(embed_dim, atom_feature_dims)
        self.e_embed_init = None
        if embed_edge:
            self.e_embed_init = BondEncoder(embed_dim) if bond_feature_dims is None \
                else BondEncoder(embed_dim, bond_feature_dims)
        self.reduce_init = InitReduceConv(reduce=init_reduce)
        self.init_conv = OGBEmbedVEWithReduce(self.v_embed_init, self.e_embed_init, self.reduce_init)
        self.dropout_rate = dropout_rate
        self.in_dropout_rate = indropout_rate
        self._build_trunk(out_size, num_layers, hidden, max_dim, jump_mode, nonlinearity, readout, train_eps,
                          final_hidden_multiplier, readout_dims, final_readout, apply_dropout_before, embed_dim,
                          use_coboundaries, graph_norm)

    def forward(self, data: ComplexBatch, include_partial=False):
        return self._forward(data, include_partial, in_dropout=self.in_dropout_rate,
                             conv_dropout=self.dropout_rate)


def _cinpp_convs(model, num_layers, hidden, embed_dim, nonlinearity, train_eps, use_coboundaries):
    convs = torch.nn.ModuleList()
    act_module = get_nonlinearity(nonlinearity, return_module=True)
    if embed_dim is None:
        embed_dim = hidden
    for i in range(num_layers):
        layer_dim = embed_dim if i == 0 else hidden
        convs.append(
            CINppConv(up_msg_size=layer_dim, down_msg_size=layer_dim, boundary_msg_size=layer_dim,
                      passed_msg_boundaries_nn=None, passed_msg_up_nn=None, passed_msg_down_nn=None,
                      passed_update_up_nn=None, passed_update_down_nn=None, passed_update_boundaries_nn=None,
                      train_eps=train_eps, max_dim=model.max_dim, hidden=hidden, act_module=act_module,
                      layer_dim=layer_dim, graph_norm=model.graph_norm, use_coboundaries=use_coboundaries))
    return convs


class EmbedCINpp(EmbedSparseCIN):
    """CIN++ for ZINC-style molecules (reference `mp/molec_models.py:167-199`): EmbedSparseCIN with `CINppConv` layers."""

    def __init__(self, atom_types, bond_types, out_size, num_layers, hidden, dropout_rate: float = 0.5,
                 max_dim: int = 2, jump_mode=None, nonlinearity='relu', readout='sum', train_eps=False,
                 final_hidden_multiplier: int = 2, readout_dims=(0, 1, 2), final_readout='sum',
                 apply_dropout_before='lin2', init_reduce='sum', embed_edge=False, embed_dim=None,
                 use_coboundaries=False, graph_norm='bn'):
        super(EmbedCINpp, self).__init__(atom_types, bond_types, out_size, num_layers, hidden, dropout_rate, max_dim,
                                         jump_mode, nonlinearity, readout, train_eps, final_hidden_multiplier,
                                         readout_dims, final_readout, apply_dropout_before, init_reduce, embed_edge,
                                         embed_dim, use_coboundaries, graph_norm)
        self.convs = _cinpp_convs(self, num_layers, hidden, embed_dim, nonlinearity, train_eps, use_coboundaries)


class OGBEmbedCINpp(OGBEmbedSparseCIN):
    """CIN++ for ogbg-mol* (reference `mp/molec_models.py:355-385`)."""

    def __init__(self, out_size, num_layers, hidden, dropout_rate: float = 0.5, indropout_rate: float = 0,
                 max_dim: int = 2, jump_mode=None, nonlinearity='relu', readout='sum', train_eps=False,
                 final_hidden_multiplier: int = 2, readout_dims=(0, 1, 2), final_readout='sum',
                 apply_dropout_before='lin2', init_reduce='sum', embed_edge=False, embed_dim=None,
                 use_coboundaries=False, graph_norm='bn', atom_feature_dims=None, bond_feature_dims=None):
        super(OGBEmbedCINpp, self).__init__(out_size, num_layers, hidden, dropout_rate, indropout_rate, max_dim,
                                            jump_mode, nonlinearity, readout, train_eps, final_hidden_multiplier,
                                            readout_dims, final_readout, apply_dropout_before, init_reduce,
                                            embed_edge, embed_dim, use_coboundaries, graph_norm, atom_feature_dims,
                                            bond_feature_dims)
        self.convs = _cinpp_convs(self, num_layers, hidden, embed_dim, nonlinearity, train_eps, use_coboundaries)
