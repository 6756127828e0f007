"""CPU ORACLE of the cochain message-passing hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/` (including the analysis script tests/analysis_tf32_error_budget.py and the test-only ops shim
tests/cpu_ops_shim.py), `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may import
this file; nothing under `cwn_b200/` does. It is a dependency-free (torch-only, CPU or any device) restatement of what
the reference computes on this path, written functionally over a `state_dict` so that the same weights can be fed
to it and to the CUDA implementation:

  propagate                mp/cell_mp.py:357-392 (+ __collect__/__lift__ :195-282, aggregate_* :423-479, update :511-524)
  get_cochain_params       data/complex.py:548-626
  sparse_cin_cochain_conv  mp/layers.py:184-214 with the default nets of SparseCINConv :286-325
  cinpp_cochain_conv       mp/layers.py:243-260 (CIN++; models mp/models.py:259-283, mp/molec_models.py:167-199, :355-385)
  cin_cochain_conv         mp/layers.py:78-103 with the nets of CIN0, mp/models.py:34-50
  init_reduce / embed_ve   mp/layers.py:484-487, :516-543
  pool_complex             mp/nn.py:50-60
  sparse_cin / embed_sparse_cin / ogb_embed_sparse_cin / cin0   mp/models.py:194-254, mp/molec_models.py:90-161,
                           :281-350, mp/models.py:84-106
  edge_cin0                mp/models.py:286-419 over EdgeCINConv mp/layers.py:127-151
  oriented_edge_model      EdgeOrient / EdgeMPNN mp/models.py:474-608 over OrientedConv mp/layers.py:430-470

Third-party semantics restated (their sources are not under /root/reference): torch_scatter 2.0.5
`scatter(reduce=add|mean|max)` = zeros(dim_size) + scatter_add_ (mean: / clamp(count,1); max: empty rows 0);
PyG `global_add_pool/global_mean_pool` = scatter over `batch`; `JumpingKnowledge('cat'|'max')`; ogb encoders =
sum of per-column embeddings.

Parity pin: checked in tests/test_oracle_golden.py against (a) the literal known answers of the reference's own
tests (mp/test_cell_mp.py, mp/test_layers.py, data/test_data.py) and (b) golden vectors produced by running the
reference's unmodified mp/ + data/complex.py code in the build container with the third-party stand-ins of
oracle/ref_shims (tests/golden/make_golden.py). Unpinned (the restatement IS the definition): BatchNorm
training-mode numerics and ogb vocabularies, which live in torch / ogb, not in the reference.
"""
import torch
import torch.nn.functional as F

_ACT = {'relu': F.relu, 'elu': F.elu, 'id': lambda v: v, 'sigmoid': torch.sigmoid, 'tanh': torch.tanh}


# ------------------------------------------------------------------------------------------ third-party semantics
def scatter(src, index, dim_size, reduce='add'):
    """torch_scatter.scatter(src, index, dim=0, dim_size=dim_size, reduce) for 2-D `src`."""
    out = torch.zeros(dim_size, src.size(1), dtype=src.dtype, device=src.device)
    idx = index.unsqueeze(-1).expand_as(src)
    if reduce in ('add', 'sum'):
        return out.scatter_add_(0, idx, src)
    if reduce == 'mean':
        out.scatter_add_(0, idx, src)
        count = torch.zeros(dim_size, dtype=src.dtype, device=src.device)
        count.scatter_add_(0, index, torch.ones_like(index, dtype=src.dtype))
        return out / count.clamp_(min=1).unsqueeze(-1)
    if reduce == 'max':
        return out.scatter_reduce_(0, idx, src, reduce='amax', include_self=False)
    raise ValueError(reduce)


# ------------------------------------------------------------------------------------------ data API restatement
class Params(object):
    """What `Complex.get_cochain_params` returns (fields of CochainMessagePassingParams)."""

    def __init__(self, x, up_index, down_index, up_attr, down_attr, boundary_attr, boundary_index):
        self.x, self.up_index, self.down_index = x, up_index, down_index
        self.up_attr, self.down_attr = up_attr, down_attr
        self.boundary_attr, self.boundary_index = boundary_attr, boundary_index


def get_cochain_params(cochains, dim, max_dim=2, include_top_features=True, include_down_features=True,
                       include_boundary_features=True):
    """`cochains`: dict dim -> object with attributes x, upper_index, lower_index, boundary_index,
    shared_boundaries, shared_coboundaries (data/complex.py:568-598)."""
    cells = cochains[dim]
    up_index = up_attr = None
    if cells.upper_index is not None and (dim + 1) in cochains:
        up_index = cells.upper_index
        if cochains[dim + 1].x is not None and (dim < max_dim or include_top_features):
            up_attr = cochains[dim + 1].x.index_select(0, cells.shared_coboundaries)
    down_index = down_attr = None
    if include_down_features and cells.lower_index is not None:
        down_index = cells.lower_index
        if dim > 0 and cochains[dim - 1].x is not None:
            down_attr = cochains[dim - 1].x.index_select(0, cells.shared_boundaries)
    b_index = b_attr = None
    if include_boundary_features and cells.boundary_index is not None:
        b_index = cells.boundary_index
        if dim > 0 and cochains[dim - 1].x is not None:
            b_attr = cochains[dim - 1].x
    return Params(cells.x, up_index, down_index, up_attr, down_attr, b_attr, b_index)


def get_all_cochain_params(data, max_dim=2, include_top_features=True, include_down_features=True,
                           include_boundary_features=True):
    return [get_cochain_params(data.cochains, d, max_dim, include_top_features, include_down_features,
                               include_boundary_features) for d in range(min(max_dim, data.dimension) + 1)]


class Snapshot(object):
    """Detached, device-moved copy of the tensors of a Complex/ComplexBatch-like object, so that the oracle can
    overwrite `x` layer after layer without touching the object under test."""

    class _C(object):
        pass

    def __init__(self, data, device='cpu'):
        self.dimension = data.dimension
        self.num_complexes = getattr(data, 'num_complexes', None)
        self.cochains = {}
        for d, c in data.cochains.items():
            s = Snapshot._C()
            for name in ('x', 'upper_index', 'lower_index', 'boundary_index', 'shared_boundaries',
                         'shared_coboundaries', 'batch'):
                v = getattr(c, name, None)
                setattr(s, name, None if v is None else v.detach().to(device))
            self.cochains[d] = s
        y = getattr(data, 'y', None)
        self.y = None if y is None else y.detach().to(device)


# ------------------------------------------------------------------------------------------ propagate
def propagate(x, up_index=None, down_index=None, boundary_index=None, up_attr=None, down_attr=None,
              boundary_attr=None, up_msg_size=None, down_msg_size=None, boundary_msg_size=None,
              use_down_msg=True, use_boundary_msg=True, aggr_up='add', aggr_down='add', aggr_boundary='add',
              message_up=None, message_down=None, message_boundary=None):
    """(up_out, down_out, boundary_out). Flow source_to_target: gather by index[0], scatter to index[1]; every
    output has x.size(0) rows; an absent pass is zeros of width `*_msg_size`."""
    n = x.size(0)
    boundary_msg_size = down_msg_size if boundary_msg_size is None else boundary_msg_size
    message_up = message_up or (lambda x_j, attr: x_j)
    message_down = message_down or (lambda x_j, attr: x_j)
    message_boundary = message_boundary or (lambda x_j: x_j)
    up_out = down_out = boundary_out = None
    if up_index is not None:
        msg = message_up(x.index_select(0, up_index[0]), up_attr)
        up_out = scatter(msg, up_index[1], n, aggr_up)
    if use_down_msg and down_index is not None:
        msg = message_down(x.index_select(0, down_index[0]), down_attr)
        down_out = scatter(msg, down_index[1], n, aggr_down)
    if use_boundary_msg and boundary_attr is not None:
        msg = message_boundary(boundary_attr.index_select(0, boundary_index[0]))
        boundary_out = scatter(msg, boundary_index[1], n, aggr_boundary)
    if up_out is None:
        up_out = torch.zeros(n, up_msg_size, device=x.device)
    if down_out is None:
        down_out = torch.zeros(n, down_msg_size, device=x.device)
    if boundary_out is None:
        boundary_out = torch.zeros(n, boundary_msg_size, device=x.device)
    return up_out, down_out, boundary_out


def dummy_cochain_mp(p, use_boundary_msg=False, use_down_msg=True, size=1):
    """DummyCochainMessagePassing.forward (mp/layers.py:23-40): message = x_j + attr."""
    up, down, bnd = propagate(p.x, p.up_index, p.down_index, p.boundary_index, p.up_attr, p.down_attr,
                              p.boundary_attr, size, size, size, use_down_msg, use_boundary_msg,
                              message_up=lambda x_j, a: x_j + a, message_down=lambda x_j, a: x_j + a)
    return p.x + up + down + bnd


# ------------------------------------------------------------------------------------------ nets over a state_dict
def _norm(sd, prefix, v, kind, training):
    if kind == 'id':
        return v
    if kind == 'ln':
        return F.layer_norm(v, (v.size(-1),), sd[prefix + 'weight'], sd[prefix + 'bias'])
    rm, rv = sd[prefix + 'running_mean'], sd[prefix + 'running_var']
    if training:
        nbt = sd.get(prefix + 'num_batches_tracked')
        if nbt is not None:
            nbt += 1
    return F.batch_norm(v, rm, rv, sd[prefix + 'weight'], sd[prefix + 'bias'], training, 0.1, 1e-5)


def _update_mlp(sd, prefix, v, cfg, training):
    """Linear -> norm -> act -> Linear -> norm -> act (mp/layers.py:303-321)."""
    act = _ACT[cfg['nonlinearity']]
    v = F.linear(v, sd[prefix + '0.weight'], sd[prefix + '0.bias'])
    v = act(_norm(sd, prefix + '1.', v, cfg['graph_norm'], training))
    v = F.linear(v, sd[prefix + '3.weight'], sd[prefix + '3.bias'])
    return act(_norm(sd, prefix + '4.', v, cfg['graph_norm'], training))


def sparse_cin_cochain_conv(sd, prefix, p, cfg, training, layer_dim):
    """SparseCINCochainConv.forward (mp/layers.py:184-199) with SparseCINConv's default nets."""
    act = _ACT[cfg['nonlinearity']]
    if cfg['use_coboundaries']:
        def message_up(x_j, attr):  # Catter -> Linear(2F, F) -> act (:290-293)
            return act(F.linear(torch.cat((x_j, attr), dim=-1), sd[prefix + 'msg_up_nn.1.weight'],
                                sd[prefix + 'msg_up_nn.1.bias']))
    else:
        message_up = None
    out_up, _, out_b = propagate(p.x, p.up_index, p.down_index, p.boundary_index, p.up_attr, None,
                                 p.boundary_attr, layer_dim, layer_dim, layer_dim, use_down_msg=False,
                                 message_up=message_up)
    out_up = out_up + (1 + sd[prefix + 'eps1']) * p.x
    out_b = out_b + (1 + sd[prefix + 'eps2']) * p.x
    out_up = _update_mlp(sd, prefix + 'update_up_nn.', out_up, cfg, training)
    out_b = _update_mlp(sd, prefix + 'update_boundaries_nn.', out_b, cfg, training)
    v = F.linear(torch.cat([out_up, out_b], dim=-1), sd[prefix + 'combine_nn.0.weight'],
                 sd[prefix + 'combine_nn.0.bias'])
    return act(_norm(sd, prefix + 'combine_nn.1.', v, cfg['graph_norm'], training))


def cinpp_cochain_conv(sd, prefix, p, cfg, training, layer_dim):
    """CINppCochainConv.forward (mp/layers.py:243-260). The layer is constructed with use_down_msg=False
    (mp/layers.py:223-226 -> :167-168) and its models pass include_down_features=False, so the down branch only ever
    sees zeros + (1+eps2) x; residual epsilons: up eps1, down eps2, boundaries eps3."""
    act = _ACT[cfg['nonlinearity']]
    message_up = None
    if cfg['use_coboundaries']:
        def message_up(x_j, attr):
            return act(F.linear(torch.cat((x_j, attr), dim=-1), sd[prefix + 'msg_up_nn.1.weight'],
                                sd[prefix + 'msg_up_nn.1.bias']))
    out_up, out_down, out_b = propagate(p.x, p.up_index, p.down_index, p.boundary_index, p.up_attr, None,
                                        p.boundary_attr, layer_dim, layer_dim, layer_dim, use_down_msg=False,
                                        message_up=message_up)
    out_up = out_up + (1 + sd[prefix + 'eps1']) * p.x
    out_down = out_down + (1 + sd[prefix + 'eps2']) * p.x
    out_b = out_b + (1 + sd[prefix + 'eps3']) * p.x
    out_up = _update_mlp(sd, prefix + 'update_up_nn.', out_up, cfg, training)
    out_down = _update_mlp(sd, prefix + 'update_down_nn.', out_down, cfg, training)
    out_b = _update_mlp(sd, prefix + 'update_boundaries_nn.', out_b, cfg, training)
    v = F.linear(torch.cat([out_up, out_down, out_b], dim=-1), sd[prefix + 'combine_nn.0.weight'],
                 sd[prefix + 'combine_nn.0.bias'])
    return act(_norm(sd, prefix + 'combine_nn.1.', v, cfg['graph_norm'], training))


def sparse_cin_conv(sd, prefix, params, cfg, training, layer_dim):
    level = cinpp_cochain_conv if cfg.get('cinpp') else sparse_cin_cochain_conv
    return [level(sd, f'{prefix}mp_levels.{d}.', p, cfg, training, layer_dim) for d, p in enumerate(params)]


def _cin_msg(sd, prefix, v, cfg, training):
    """Linear(2F, F) -> act -> BN over the message population (mp/models.py:40-47)."""
    v = _ACT[cfg['nonlinearity']](F.linear(v, sd[prefix + '0.weight'], sd[prefix + '0.bias']))
    return _norm(sd, prefix + '2.', v, 'bn', training)


def cin_cochain_conv(sd, prefix, p, cfg, training, layer_dim):
    """CINCochainConv.forward (mp/layers.py:78-103) with CIN0's nets (mp/models.py:34-50)."""
    act = _ACT[cfg['nonlinearity']]

    def message_up(x_j, attr):
        v = torch.cat([x_j, attr], dim=-1) if attr is not None else x_j
        return _cin_msg(sd, prefix + 'msg_up_nn.', v, cfg, training)

    def message_down(x_j, attr):
        return _cin_msg(sd, prefix + 'msg_down_nn.', torch.cat([x_j, attr], dim=-1), cfg, training)

    out_up, out_down, _ = propagate(p.x, p.up_index, p.down_index, None, p.up_attr, p.down_attr, None,
                                    layer_dim, layer_dim, None, use_boundary_msg=False,
                                    message_up=message_up, message_down=message_down)
    eps = sd[prefix + 'eps']
    v = (out_up + (1 + eps) * p.x) + (out_down + (1 + eps) * p.x)
    up = prefix + 'update_nn.'
    v = act(F.linear(v, sd[up + '0.weight'], sd[up + '0.bias']))
    v = act(F.linear(v, sd[up + '2.weight'], sd[up + '2.bias']))
    return _norm(sd, up + '4.', v, 'bn', training)


# ------------------------------------------------------------------------------------------ init / readout
def init_reduce(boundary_x, boundary_index, reduce='add', out_size=None):
    """InitReduceConv.forward (mp/layers.py:484-487)."""
    out_size = int(boundary_index[1].max()) + 1 if out_size is None else out_size
    return scatter(boundary_x.index_select(0, boundary_index[0]), boundary_index[1], out_size, reduce)


def embed_ve_with_reduce(sd, params, cfg, ogb=False):
    """AbstractEmbedVEWithReduce.forward (mp/layers.py:516-543)."""
    def embed(prefix, x, list_name):
        if not ogb:
            return F.embedding(x.squeeze(1).long(), sd[prefix + 'weight'])
        x = x.long()
        return sum(F.embedding(x[:, i], sd[f'{prefix}{list_name}.{i}.weight']) for i in range(x.size(1)))

    vx = embed('v_embed_init.', params[0].x, 'atom_embedding_list')
    out = [vx]
    if len(params) == 1:
        return out
    reduce = cfg.get('init_reduce', 'sum')
    reduced_ex = init_reduce(vx, params[1].boundary_index, reduce)
    ex = reduced_ex
    if params[1].x is not None:
        ex = embed('e_embed_init.', params[1].x, 'bond_embedding_list')
    out.append(ex)
    if len(params) == 3:
        out.append(init_reduce(reduced_ex, params[2].boundary_index, reduce) / 2.)
    return out


def pool_complex(xs, data, max_dim, readout, batch_size=None):
    """[max_dim+1, B, F] per-complex readout; absent dimensions stay zero (mp/nn.py:50-60)."""
    if batch_size is None:
        batch_size = int(data.cochains[0].batch.max()) + 1
    pooled = torch.zeros(max_dim + 1, batch_size, xs[0].size(-1), device=xs[0].device)
    for i in range(len(xs)):
        pooled[i] = scatter(xs[i], data.cochains[i].batch, batch_size, 'mean' if readout == 'mean' else 'add')
    return pooled


def _jump(xs_per_layer, mode):
    if mode == 'cat':
        return torch.cat(xs_per_layer, dim=-1)
    return torch.stack(xs_per_layer, dim=-1).max(dim=-1)[0]


def _set_xs(data, xs):
    for i, x in enumerate(xs):
        data.cochains[i].x = x


# ------------------------------------------------------------------------------------------ model forwards
DEFAULTS = dict(dropout_rate=0.5, indropout_rate=0.0, max_dim=2, jump_mode=None, nonlinearity='relu', readout='sum',
                final_hidden_multiplier=2, readout_dims=(0, 1, 2), final_readout='sum',
                apply_dropout_before='lin2', init_reduce='sum', use_coboundaries=False, graph_norm='bn')


def _cfg(cfg):
    out = dict(DEFAULTS)
    out.update(cfg)
    out['graph_norm'] = {'bn': 'bn', 'ln': 'ln', 'id': 'id'}[out['graph_norm']]
    return out


def _sparse_family(sd, cfg, data, training, include_partial, embed=None, first_dim=None):
    """SparseCIN / EmbedSparseCIN / OGBEmbedSparseCIN forwards. Dropout must be inactive (p = 0 or eval): the
    oracle has no access to the other implementation's random mask."""
    cfg = _cfg(cfg)
    act = _ACT[cfg['nonlinearity']]
    max_dim = cfg['max_dim']
    assert not training or (cfg['dropout_rate'] == 0 and cfg['indropout_rate'] == 0), 'dropout breaks parity'
    readout_dims = tuple(d for d in cfg['readout_dims'] if d <= max_dim) if cfg['readout_dims'] is not None \
        else tuple(range(max_dim + 1))
    res = {}
    if embed is not None:
        params = get_all_cochain_params(data, max_dim, include_down_features=False)
        _set_xs(data, embed_ve_with_reduce(sd, params, cfg, ogb=(embed == 'ogb')))
    jump_xs, xs = None, None
    for c in range(cfg['num_layers']):
        params = get_all_cochain_params(data, max_dim, include_down_features=False)
        layer_dim = first_dim if c == 0 else cfg['hidden']
        xs = sparse_cin_conv(sd, f'convs.{c}.', params, cfg, training, layer_dim)
        _set_xs(data, xs)
        if include_partial:
            for k, x in enumerate(xs):
                res[f'layer{c}_{k}'] = x
        if cfg['jump_mode'] is not None:
            if jump_xs is None:
                jump_xs = [[] for _ in xs]
            for i, x in enumerate(xs):
                jump_xs[i].append(x)
    if cfg['jump_mode'] is not None:
        xs = [_jump(j, cfg['jump_mode']) for j in jump_xs]
    pooled = pool_complex(xs, data, max_dim, cfg['readout'], getattr(data, 'num_complexes', None))
    xs = [pooled[i] for i in readout_dims]
    if include_partial:
        for k, x in enumerate(xs):
            res[f'pool_{k}'] = x
    new_xs = []
    for i, x in enumerate(xs):
        d = readout_dims[i]
        new_xs.append(act(F.linear(x, sd[f'lin1s.{d}.weight'], sd.get(f'lin1s.{d}.bias'))))
    x = torch.stack(new_xs, dim=0)
    x = x.mean(0) if cfg['final_readout'] == 'mean' else x.sum(0)
    x = F.linear(x, sd['lin2.weight'], sd['lin2.bias'])
    if include_partial:
        res['out'] = x
        return x, res
    return x


def sparse_cin(sd, cfg, data, training=False, include_partial=False):
    """SparseCIN.forward (mp/models.py:194-254). cfg: num_input_features, num_layers, hidden + DEFAULTS keys."""
    return _sparse_family(sd, cfg, data, training, include_partial, None, cfg['num_input_features'])


def embed_sparse_cin(sd, cfg, data, training=False, include_partial=False):
    """EmbedSparseCIN.forward (mp/molec_models.py:90-161). cfg: num_layers, hidden, embed_dim (optional)."""
    return _sparse_family(sd, cfg, data, training, include_partial, 'zinc', cfg.get('embed_dim') or cfg['hidden'])


def ogb_embed_sparse_cin(sd, cfg, data, training=False, include_partial=False):
    """OGBEmbedSparseCIN.forward (mp/molec_models.py:281-350)."""
    return _sparse_family(sd, cfg, data, training, include_partial, 'ogb', cfg.get('embed_dim') or cfg['hidden'])


def cinpp(sd, cfg, data, training=False, include_partial=False):
    """CINpp.forward = SparseCIN.forward over CINppConv layers (mp/models.py:259-283)."""
    return sparse_cin(sd, dict(cfg, cinpp=True), data, training, include_partial)


def embed_cinpp(sd, cfg, data, training=False, include_partial=False):
    """EmbedCINpp (mp/molec_models.py:167-199)."""
    return embed_sparse_cin(sd, dict(cfg, cinpp=True), data, training, include_partial)


def ogb_embed_cinpp(sd, cfg, data, training=False, include_partial=False):
    """OGBEmbedCINpp (mp/molec_models.py:355-385)."""
    return ogb_embed_sparse_cin(sd, dict(cfg, cinpp=True), data, training, include_partial)


def cin0(sd, cfg, data, training=False):
    """CIN0.forward (mp/models.py:84-106). cfg: num_input_features, num_layers, hidden, max_dim, jump_mode, ..."""
    cfg = _cfg(cfg)
    act = _ACT[cfg['nonlinearity']]
    max_dim = cfg['max_dim']
    assert not training or cfg['dropout_rate'] == 0, 'dropout breaks parity'
    jump_xs, xs = None, None
    for c in range(cfg['num_layers']):
        params = get_all_cochain_params(data, max_dim)
        layer_dim = cfg['num_input_features'] if c == 0 else cfg['hidden']
        xs = [cin_cochain_conv(sd, f'convs.{c}.mp_levels.{d}.', p, cfg, training, layer_dim)
              for d, p in enumerate(params)]
        _set_xs(data, xs)
        if cfg['jump_mode'] is not None:
            if jump_xs is None:
                jump_xs = [[] for _ in xs]
            for i, x in enumerate(xs):
                jump_xs[i].append(x)
    if cfg['jump_mode'] is not None:
        xs = [_jump(j, cfg['jump_mode']) for j in jump_xs]
    x = pool_complex(xs, data, max_dim, cfg['readout'], getattr(data, 'num_complexes', None)).sum(dim=0)
    x = act(F.linear(x, sd['lin1.weight'], sd['lin1.bias']))
    return F.linear(x, sd['lin2.weight'], sd['lin2.bias'])


# ------------------------------------------------------------------------------------------ edge-level models
def edge_cin0(sd, cfg, data, training=False):
    """EdgeCIN0.forward (mp/models.py:387-416) over EdgeCINConv (mp/layers.py:127-151): CINCochainConv levels for
    vertices and edges with their own nets; optional refresh of the two-cell features between layers."""
    cfg = dict(cfg)
    cfg.setdefault('nonlinearity', 'relu')
    cfg.setdefault('readout', 'sum')
    cfg.setdefault('jump_mode', None)
    cfg.setdefault('dropout_rate', 0.5)
    include_top = cfg.get('include_top_features', True)
    update_top = include_top and cfg.get('update_top_features', True)
    act = _ACT[cfg['nonlinearity']]
    assert not training or cfg['dropout_rate'] == 0, 'dropout breaks parity'
    L = cfg['num_layers']
    jump_xs, xs = None, None
    for c in range(L):
        params = get_all_cochain_params(data, 1, include_top_features=include_top)
        layer_dim = cfg['num_input_features'] if c == 0 else cfg['hidden']
        xs = [cin_cochain_conv(sd, f'convs.{c}.mp_levels.{d}.', p, cfg, training, layer_dim)
              for d, p in enumerate(params)]
        if update_top and c < L - 1 and 2 in data.cochains:
            up = f'update_top_nns.{c}.'
            v = data.cochains[2].x
            v = act(F.linear(v, sd[up + '0.weight'], sd[up + '0.bias']))
            v = act(F.linear(v, sd[up + '2.weight'], sd[up + '2.bias']))
            _set_xs(data, xs + [_norm(sd, up + '4.', v, 'bn', training)])
        else:
            _set_xs(data, xs)
        if cfg['jump_mode'] is not None:
            if jump_xs is None:
                jump_xs = [[] for _ in xs]
            for i, x in enumerate(xs):
                jump_xs[i].append(x)
    if cfg['jump_mode'] is not None:
        xs = [_jump(j, cfg['jump_mode']) for j in jump_xs]
    x = pool_complex(xs, data, 1, cfg['readout'], getattr(data, 'num_complexes', None)).sum(dim=0)
    x = act(F.linear(x, sd['lin1.weight'], sd['lin1.bias']))
    return F.linear(x, sd['lin2.weight'], sd['lin2.bias'])


def oriented_edge_model(sd, cfg, x, upper_index, lower_index, upper_orient, lower_orient, batch, num_cochains,
                        with_up=True, training=False):
    """EdgeOrient (with_up=True, mp/models.py:474-545) / EdgeMPNN (with_up=False, :548-608) over OrientedConv
    (mp/layers.py:430-470): message = x_j * orientation (unless fully_invar), out = act(U x + U_up SUM_up + U_down
    SUM_down) with bias-free linears; |.|, per-complex readout, lin1 + ReLU, lin2. Returns (out, cell_pred)."""
    fully_invar = cfg.get('fully_invar', not with_up)
    act = _ACT[cfg.get('nonlinearity', 'id' if with_up else 'relu')]
    assert not training or cfg.get('dropout_rate', 0.0) == 0, 'dropout breaks parity'
    orient = not fully_invar
    if fully_invar:
        x = torch.abs(x)
    for c in range(cfg['num_layers']):
        size = x.size(1)
        msg_up = (lambda x_j, a: x_j * a) if orient else (lambda x_j, a: x_j)
        out_up, out_down, _ = propagate(x, upper_index, lower_index, None, upper_orient.view(-1, 1),
                                        lower_orient.view(-1, 1), None, size, size, None, use_boundary_msg=False,
                                        message_up=msg_up, message_down=msg_up)
        pre = f'convs.{c}.'
        v = F.linear(x, sd[pre + 'update_nn.weight']) + F.linear(out_down, sd[pre + 'update_down_nn.weight'])
        if with_up:
            # reference order of the sum: x + out_up + out_down (mp/layers.py:452)
            v = F.linear(x, sd[pre + 'update_nn.weight']) + F.linear(out_up, sd[pre + 'update_up_nn.weight']) \
                + F.linear(out_down, sd[pre + 'update_down_nn.weight'])
        x = act(v)
    cell_pred = x
    if not fully_invar:
        x = torch.abs(x)
    x = scatter(x, batch, num_cochains, {'sum': 'add', 'mean': 'mean'}[cfg.get('readout', 'sum')])
    x = torch.relu(F.linear(x, sd['lin1.weight'], sd['lin1.bias']))
    return F.linear(x, sd['lin2.weight'], sd['lin2.bias']), cell_pred
