"""The Python host layer (hook protocol, layers, models, data API) against the golden vectors recorded from the
reference, in the GPU-less build container: the six `cwn_b200.ops` entry points are replaced by torch restatements
(`tests/cpu_ops_shim.py`, test infrastructure) for the duration of each test, so everything ABOVE the C ABI runs exactly
as it does on the GPU box — only the kernels are substituted. The CUDA kernels themselves are covered by `-m gpu`."""
import pytest
import torch

import cpu_ops_shim
from helpers import assert_close, batch_of, golden


def _klass(name):
    from cwn_b200.mp import models as M, molec_models as MM
    table = [('ogb_embed_sparse_cin', MM.OGBEmbedSparseCIN), ('ogb_embed_cinpp', MM.OGBEmbedCINpp),
             ('embed_sparse_cin', MM.EmbedSparseCIN),
             ('embed_cinpp', MM.EmbedCINpp), ('cinpp', M.CINpp), ('sparse_cin', M.SparseCIN), ('cin0', M.CIN0),
             ('edge_cin0', M.EdgeCIN0)]
    return next(k for prefix, k in table if name.startswith(prefix))


def _loss(name, out, y):
    if name.startswith('ogb'):
        return torch.nn.functional.binary_cross_entropy_with_logits(out, (y.view(-1, 1) > 0).float())
    return torch.nn.functional.l1_loss(out, y.view(-1, 1))


@pytest.mark.parametrize('name', ['sparse_cin_train', 'embed_sparse_cin_train', 'embed_sparse_cin_train_nocob',
                                  'ogb_embed_sparse_cin_train', 'cin0_train', 'cinpp_train', 'embed_cinpp_train', 'ogb_embed_cinpp_train',
                                  'edge_cin0_train', 'edge_cin0_notop_train'])
def test_model_host_logic_matches_reference_train_step(name, monkeypatch):
    cpu_ops_shim.install(monkeypatch)
    m = golden()['models'][name]
    kwargs = dict(m['cfg'])
    if name.startswith('ogb'):
        from cwn_b200.mp.encoders import ATOM_FEATURE_DIMS, BOND_FEATURE_DIMS
        kwargs.setdefault('atom_feature_dims', ATOM_FEATURE_DIMS)
        kwargs.setdefault('bond_feature_dims', BOND_FEATURE_DIMS)
    model = _klass(name)(**kwargs)
    model.load_state_dict(m['state_dict'])
    model.train()
    batch = batch_of(m['inputs'], max_dim=m['cfg'].get('max_dim', 2))
    out = model(batch)
    loss = _loss(name, out, batch.y)
    loss.backward()
    assert_close(out, m['output'], rtol=1e-5, atol=1e-5, what=name)
    assert_close(loss, m['loss'], rtol=1e-5, atol=1e-6, what=name + ':loss')
    got = dict(model.named_parameters())
    for k, ref in m['grads'].items():
        assert_close(got[k].grad, ref, rtol=1e-4, atol=2e-6 + 2e-5 * float(ref.abs().max()), what=f'{name}:grad:{k}')


@pytest.mark.parametrize('name', ['edge_orient_train', 'edge_mpnn_train'])
def test_oriented_model_host_logic_matches_reference_train_step(name, monkeypatch):
    cpu_ops_shim.install(monkeypatch)
    from cwn_b200.data.complex import Cochain, CochainBatch
    from cwn_b200.mp.models import EdgeMPNN, EdgeOrient
    m = golden()['models'][name]
    model = (EdgeOrient if name.startswith('edge_orient') else EdgeMPNN)(**m['cfg'])
    model.load_state_dict(m['state_dict'])
    model.train()
    cochains = [Cochain(dim=1, x=c['x'], upper_index=c['upper_index'], lower_index=c['lower_index'],
                        upper_orient=c['upper_orient'], lower_orient=c['lower_orient'], y=c['y'],
                        num_cells_up=1, num_cells_down=c['x'].size(0) + 1) for c in m['inputs']]
    batch = CochainBatch.from_cochain_list(cochains)
    out, cell_pred = model(batch, include_partial=True)
    loss = torch.nn.functional.l1_loss(out, batch.y.view(-1, 1))
    loss.backward()
    assert_close(out, m['output'], rtol=1e-5, atol=1e-5, what=name)
    assert_close(cell_pred, m['cell_pred'], rtol=1e-5, atol=1e-5, what=name + ':cell_pred')
    got = dict(model.named_parameters())
    for k, ref in m['grads'].items():
        assert_close(got[k].grad, ref, rtol=1e-4, atol=2e-6 + 2e-5 * float(ref.abs().max()), what=f'{name}:grad:{k}')


# ---------------------------------------------------------------------------------------------------------------
# The reference's own known-answer / invariance tests, run through the host layer on the CPU: the
# bodies live in tests/test_gpu_parity.py (where they run on the CUDA kernels); here the device is switched to 'cpu'
# and the ops entry points to their torch restatements, so the SAME assertions cover hook protocol, layers and models
# in the GPU-less tier.
import test_gpu_parity as G  # noqa: E402

_HOST_TWINS = [
    (G.test_reference_known_answers_on_gpu, ()),                       # mp/test_cell_mp.py:13-111, :179-270
    (G.test_isolated_cells_and_empty_index, ()),                       # mp/test_cell_mp.py:114-176
    (G.test_init_reduce_known_answers, ()),                            # mp/test_layers.py:135-149
    (G.test_user_overridden_hooks_are_honoured, ()),
    (G.test_batched_equals_unbatched, ()),                             # mp/test_models.py:139-185
    (G.test_mean_max_aggregations_equal_reference, ('mean',)),
    (G.test_mean_max_aggregations_equal_reference, ('max',)),
] + [(G.test_propagate_and_dummy_layers_equal_reference_on_every_fixture, (n,)) for n in golden()['fixtures']] \
  + [(G.test_eval_models_match_reference_outputs, (n,))
     for n in ('sparse_cin_eval', 'sparse_cin_eval_dim1', 'embed_sparse_cin_eval', 'cin0_eval')]


@pytest.mark.parametrize('fn,args', _HOST_TWINS, ids=[f'{f.__name__}{list(a)}' for f, a in _HOST_TWINS])
def test_reference_test_suite_through_the_host_layer(fn, args, monkeypatch):
    cpu_ops_shim.install(monkeypatch)
    monkeypatch.setattr(G, 'DEV', 'cpu')
    fn(*args)


@pytest.mark.parametrize('use_cob', [False, True])
def test_baseline_config0_sparse_cin_conv_on_the_house_complex(use_cob, monkeypatch):
    """BASELINE.json configs[0]: `SparseCINConv` forward on the house complex of data/dummy_complexes.py (5 vertices,
    6 edges, 1 two-cell, F = 1), CPU, world size 1 — the reference's own CPU-runnable plumbing case. Host layer (ops
    substituted) against the oracle's restatement of mp/layers.py:184-199 fed the same weights, eval-mode BatchNorm."""
    import cwn_oracle as O
    from cwn_b200.mp.layers import SparseCINConv
    from cwn_b200.mp.nn import get_graph_norm, get_nonlinearity
    from helpers import fixture
    cpu_ops_shim.install(monkeypatch)
    torch.manual_seed(0)
    conv = SparseCINConv(1, 1, 1, None, None, None, None, layer_dim=1, hidden=4, act_module=get_nonlinearity('relu'),
                         graph_norm=get_graph_norm('bn'), use_coboundaries=use_cob).eval()
    house = fixture('house')
    params = house.get_all_cochain_params(max_dim=2, include_down_features=False)
    with torch.no_grad():
        outs = conv(*params)
    sd = {k: v.detach().clone() for k, v in conv.state_dict().items()}
    cfg = dict(nonlinearity='relu', graph_norm='bn', use_coboundaries=use_cob)
    snap = O.Snapshot(fixture('house'))
    ref = O.sparse_cin_conv(sd, '', O.get_all_cochain_params(snap, 2, include_down_features=False), cfg, False, 1)
    assert [tuple(o.shape) for o in outs] == [(5, 4), (6, 4), (1, 4)]
    for d, (o, r) in enumerate(zip(outs, ref)):
        assert_close(o, r, rtol=1e-6, atol=1e-6, what=f'house dim {d}')


def test_padded_batches_share_one_layout():
    """cwn_b200.bucketed.pad_complexes: ragged batches (and a short last batch) completed with dummy complexes all have
    ONE packed layout; the real cells keep the leading rows, the dummies are disconnected from them, and the live-row
    scalars / complex weights describe the real part."""
    from cwn_b200.bucketed import Capacity, Overflow, pad_complexes
    from cwn_b200.data import synthetic
    from cwn_b200.data.complex import ComplexBatch
    pool = synthetic.zinc_like_complexes(300, seed=3, ragged=True)
    cap = Capacity.from_dataset(pool, 64)
    sigs = set()
    for lo, hi in [(0, 64), (64, 128), (128, 192), (192, 200), (7, 8)]:
        comps = pool[lo:hi]
        ref = ComplexBatch.from_complex_list(comps)
        pb = pad_complexes(comps, cap)
        sigs.add(pb.pack_().packed_signature)
        assert pb.num_complexes == 65 and float(pb.cochains[0].complex_weight.sum()) == len(comps)
        for d in range(3):
            n = ref.cochains[d].num_cells if d <= ref.dimension else 0
            c = pb.cochains[d]
            assert int(c.live_rows) == n and c.num_cells == cap.cells[d]
            if n and ref.cochains[d].x is not None:
                assert torch.equal(c.x[:n], ref.cochains[d].x)
            assert int((c.batch[:n] >= len(comps)).sum()) == 0 and int((c.batch[n:] < len(comps)).sum()) == 0
            for key, src_n in (('upper_index', n), ('boundary_index', ref.cochains[d - 1].num_cells if d else 0)):
                idx = getattr(c, key)
                if idx is None:
                    continue
                real_dst, real_src = idx[1] < n, idx[0] < src_n
                assert torch.equal(real_dst, real_src), 'a message crosses between real and padding cells'
    assert len(sigs) == 1
    with pytest.raises(Overflow):
        pad_complexes(pool[:65], cap)


def test_ogb_encoders_one_gather_over_the_concatenated_tables(monkeypatch):
    """The fused form of ogb's Atom/BondEncoder (`mp/encoders.py::_lookup_concat`: one row gather over the concatenated
    embedding tables, summed over the feature columns) equals the per-column lookups the reference performs
    (ogb mol_encoder.py via `mp/molec_models.py:7`), values and every table's gradient; the gather goes through the
    CPU stand-in of `ops.gather_rows` here (the CUDA guard is what `forward` checks on the GPU)."""
    import cpu_ops_shim
    from cwn_b200 import ops
    from cwn_b200.mp.encoders import ATOM_FEATURE_DIMS, BOND_FEATURE_DIMS, AtomEncoder, BondEncoder
    monkeypatch.setattr(ops, 'gather_rows', cpu_ops_shim.gather_rows)
    g = torch.Generator().manual_seed(0)
    for cls, dims in ((AtomEncoder, ATOM_FEATURE_DIMS), (BondEncoder, BOND_FEATURE_DIMS)):
        torch.manual_seed(1)
        enc = cls(12)
        x = torch.stack([torch.randint(0, v, (70,), generator=g) for v in dims], 1)
        w = torch.randn(70, 12, generator=g)
        ref = enc(x)  # CPU tensors: the per-column path
        (ref * w).sum().backward()
        ref_grads = [t.weight.grad.clone() for t in getattr(enc, enc._list_name)]
        enc.zero_grad()
        out = enc._lookup_concat(x)
        (out * w).sum().backward()
        assert torch.allclose(out, ref, rtol=1e-6, atol=1e-6)
        for t, rg in zip(getattr(enc, enc._list_name), ref_grads):
            assert torch.allclose(t.weight.grad, rg, rtol=1e-6, atol=1e-6)
