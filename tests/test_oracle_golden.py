"""The CPU oracle against (a) the literal known answers of the reference's own tests and (b) golden vectors
recorded by running the reference (tests/golden/make_golden.py). CPU only."""
import pytest
import torch

import cwn_oracle as O
from helpers import assert_close, batch_of, fixture, golden, oracle_state, share_cin0_levels


def _params(comp, dim):
    return O.get_cochain_params(O.Snapshot(comp).cochains, dim)


def _propagate(p, **kw):
    return O.propagate(p.x, p.up_index, p.down_index, p.boundary_index, p.up_attr, p.down_attr, p.boundary_attr,
                       1, 1, 1, **kw)


def T(v):
    return torch.tensor(v, dtype=torch.float).view(-1, 1)


def test_house_edge_level_known_answers():  # reference mp/test_cell_mp.py:13-35
    up, down, bnd = _propagate(_params(fixture('house'), 1))
    assert torch.equal(down, T([6, 10, 17, 9, 13, 10]))
    assert torch.equal(up, T([0, 0, 11, 0, 9, 8]))
    assert torch.equal(bnd, T([3, 5, 7, 5, 9, 8]))


def test_house_vertex_level_known_answers():  # mp/test_cell_mp.py:38-62
    up, down, bnd = _propagate(_params(fixture('house'), 0))
    assert torch.equal(up, T([6, 4, 11, 9, 7]))
    assert torch.equal(down, torch.zeros(5, 1))
    assert torch.equal(bnd, torch.zeros(5, 1))


def test_house_two_cell_level_known_answers():  # mp/test_cell_mp.py:65-88
    up, down, bnd = _propagate(_params(fixture('house'), 2))
    assert torch.equal(up, torch.zeros(1, 1))
    assert torch.equal(down, torch.zeros(1, 1))
    assert torch.equal(bnd, T([14]))


def test_replicated_adjacencies_are_summed_with_multiplicity():  # mp/test_cell_mp.py:179-270
    comp = fixture('bridged')
    up, _, _ = _propagate(_params(comp, 1))
    assert torch.equal(up, T([4 + 5 + 6 + 2 + 3 + 4, 3 + 5 + 6 + 1 + 3 + 4, 2 + 5 + 6 + 1 + 2 + 4,
                              1 + 5 + 6 + 1 + 2 + 3, 1 + 4 + 6 + 2 + 3 + 6, 1 + 4 + 5 + 2 + 3 + 5]))
    _, down, bnd = _propagate(_params(comp, 2))
    assert torch.equal(down, T([2 + 2 + 3 + 3, 1 + 1 + 3 + 3, 1 + 1 + 2 + 2]))
    assert torch.equal(bnd, T([1 + 6 + 5 + 4, 2 + 3 + 5 + 6, 1 + 2 + 3 + 4]))


def test_isolated_and_empty():  # mp/test_cell_mp.py:114-176
    p = _params(fixture('square_dot'), 0)
    up, down, _ = _propagate(p)
    assert torch.equal(up[4], torch.zeros(1)) and bool((up[:4] != 0).all())
    assert torch.equal(down, torch.zeros_like(down))
    for name in ('fullstop', 'colon'):
        p = _params(fixture(name), 0)
        up, _, _ = _propagate(p)
        assert torch.equal(up, torch.zeros_like(p.x))


def test_dummy_layers_known_answers():  # mp/test_layers.py:11-69
    house = O.Snapshot(fixture('house'))
    ps = [O.get_cochain_params(house.cochains, d) for d in range(3)]
    out = [O.dummy_cochain_mp(p) for p in ps]
    assert torch.equal(out[0], T([12, 9, 25, 25, 23]))
    assert torch.equal(out[1], T([10, 20, 47, 22, 42, 37]))
    assert torch.equal(out[2], T([1]))
    out = [O.dummy_cochain_mp(p, use_boundary_msg=True, use_down_msg=False) for p in ps]
    assert torch.equal(out[1], T([4, 7, 23, 9, 25, 24]))
    assert torch.equal(out[2], T([15]))
    mol = O.Snapshot(fixture('molecular'))
    out = [O.dummy_cochain_mp(O.get_cochain_params(mol.cochains, d), True, True) for d in range(3)]
    assert torch.equal(out[0], T([12, 24, 24, 15, 25, 31, 47, 24]))
    assert torch.equal(out[1], T([35, 79, 41, 27, 66, 70, 92, 82, 53]))
    assert torch.equal(out[2], T([15, 33]))


def test_init_reduce_known_answers():  # mp/test_layers.py:135-149
    house = O.Snapshot(fixture('house'))
    assert torch.equal(O.init_reduce(house.cochains[0].x, house.cochains[1].boundary_index), T([3, 5, 7, 5, 9, 8]))
    assert torch.equal(O.init_reduce(house.cochains[1].x, house.cochains[2].boundary_index), T([14]))


def test_up_down_attr_known_answers():  # data/test_data.py:6-54
    house = O.Snapshot(fixture('house'))
    v = O.get_cochain_params(house.cochains, 0)
    assert torch.equal(v.up_attr, T([1, 1, 4, 4, 2, 2, 3, 3, 6, 6, 5, 5]))
    e = O.get_cochain_params(house.cochains, 1)
    assert torch.equal(e.up_attr, T([1] * 6))
    assert torch.equal(e.down_attr, T([2, 2, 1, 1, 3, 3, 3, 3, 4, 4, 4, 4, 3, 3, 4, 4, 5, 5]))
    t = O.get_cochain_params(house.cochains, 2)
    assert t.up_attr is None and t.down_attr is None
    e = O.get_cochain_params(house.cochains, 1, max_dim=1, include_top_features=False)
    assert e.up_attr is None and e.up_index is not None and e.up_index.size(1) == 6


@pytest.mark.parametrize('name', list(golden()['fixtures']))
def test_propagate_matches_reference_on_every_fixture(name):
    kat = golden()['kat'][name]
    comp = fixture(name)
    for d in range(comp.dimension + 1):
        got = _propagate(_params(comp, d))
        for g, r in zip(got, kat['propagate'][d]):
            assert torch.equal(g, r)
    snap = O.Snapshot(comp)
    for b, dn in [(False, True), (True, False), (True, True)]:
        ref = kat[f'dummy_b{int(b)}_d{int(dn)}']
        if isinstance(ref, str):
            continue
        got = [O.dummy_cochain_mp(O.get_cochain_params(snap.cochains, d), b, dn)
               for d in range(min(2, comp.dimension) + 1)]
        for g, r in zip(got, ref):
            assert torch.equal(g, r)


@pytest.mark.parametrize('reduce', ['mean', 'max'])
def test_mean_max_aggregation_matches_reference(reduce):
    comp = fixture('house')
    for d in range(3):
        got = _propagate(_params(comp, d), aggr_up=reduce, aggr_down=reduce, aggr_boundary=reduce)
        for g, r in zip(got, golden()['kat']['house'][f'propagate_{reduce}'][d]):
            assert torch.equal(g, r)


FORWARDS = {'sparse_cin': O.sparse_cin, 'embed_sparse_cin': O.embed_sparse_cin,
            'ogb_embed_sparse_cin': O.ogb_embed_sparse_cin, 'cinpp': O.cinpp, 'embed_cinpp': O.embed_cinpp,
            'ogb_embed_cinpp': O.ogb_embed_cinpp}


@pytest.mark.parametrize('name', ['sparse_cin_eval', 'sparse_cin_eval_dim1', 'embed_sparse_cin_eval', 'cin0_eval'])
def test_eval_models_match_reference(name):
    m = golden()['models'][name]
    sd = oracle_state(m['state_dict'])
    for chunk, ref in zip(m['chunks'], m['outputs']):
        batch = batch_of(chunk, max_dim=m['max_dim'])
        if m['strip']:
            for d in (1, 2):
                if d in batch.cochains:
                    batch.cochains[d]._x = None
        snap = O.Snapshot(batch)
        with torch.no_grad():
            if name.startswith('cin0'):
                assert_close(O.cin0(sd, m['cfg'], snap), ref, atol=1e-5, what=name)
            else:
                fwd = FORWARDS['embed_sparse_cin' if name.startswith('embed') else 'sparse_cin']
                out, res = fwd(sd, m['cfg'], snap, include_partial=True)
                assert_close(out, ref[0], atol=1e-6, what=name)
                assert set(res) == set(ref[1])
                for k in res:
                    assert_close(res[k], ref[1][k], atol=1e-6, what=f'{name}:{k}')


def _loss(name, out, y):
    if name.startswith('ogb'):
        return torch.nn.functional.binary_cross_entropy_with_logits(out, (y.view(-1, 1) > 0).float())
    return torch.nn.functional.l1_loss(out, y.view(-1, 1))


@pytest.mark.parametrize('name', ['sparse_cin_train', 'embed_sparse_cin_train', 'embed_sparse_cin_train_nocob',
                                  'ogb_embed_sparse_cin_train', 'cin0_train', 'cinpp_train', 'embed_cinpp_train', 'ogb_embed_cinpp_train'])
def test_train_models_match_reference_forward_backward(name):
    m = golden()['models'][name]
    sd = oracle_state(m['state_dict'], requires_grad=True)
    if name.startswith('cin0'):  # CINConv hands the SAME nn objects to every dimension (mp/layers.py:113-116)
        share_cin0_levels(sd)
    batch = batch_of(m['inputs'], max_dim=m['cfg'].get('max_dim', 2))
    snap = O.Snapshot(batch)
    if name.startswith('cin0'):
        out = O.cin0(sd, m['cfg'], snap, training=True)
    else:
        key = next(k for k in ('ogb_embed_sparse_cin', 'ogb_embed_cinpp', 'embed_sparse_cin', 'embed_cinpp', 'cinpp', 'sparse_cin')
                   if name.startswith(k))
        out = FORWARDS[key](sd, m['cfg'], snap, training=True)
    assert_close(out, m['output'], atol=1e-6, what=name)
    loss = _loss(name, out, snap.y)
    assert_close(loss, m['loss'], atol=1e-6, what=name + ':loss')
    leaves = {k: v for k, v in sd.items() if v.requires_grad}
    uniq = {id(v): v for v in leaves.values()}
    grads = dict(zip(uniq.keys(), torch.autograd.grad(loss, list(uniq.values()), allow_unused=True)))
    for k, ref in m['grads'].items():
        g = grads[id(sd[k])]
        assert g is not None, k
        assert_close(g, ref, rtol=1e-4, atol=1e-6, what=f'{name}:grad:{k}')
    for k, ref in m['state_dict_after'].items():  # BatchNorm running statistics were updated identically
        if 'running' in k or 'num_batches' in k:
            assert_close(sd[k].float(), ref.float(), atol=1e-6, what=f'{name}:{k}')


@pytest.mark.parametrize('name', ['edge_cin0_train', 'edge_cin0_notop_train'])
def test_edge_cin0_matches_reference_forward_backward(name):
    """EdgeCIN0 / EdgeCINConv (reference mp/models.py:286-419, mp/layers.py:127-151) — SURVEY 8(f) rank 3."""
    m = golden()['models'][name]
    sd = oracle_state(m['state_dict'], requires_grad=True)
    snap = O.Snapshot(batch_of(m['inputs'], max_dim=2))
    out = O.edge_cin0(sd, m['cfg'], snap, training=True)
    assert_close(out, m['output'], atol=1e-6, what=name)
    loss = torch.nn.functional.l1_loss(out, snap.y.view(-1, 1))
    assert_close(loss, m['loss'], atol=1e-6, what=name + ':loss')
    leaves = {k: v for k, v in sd.items() if v.requires_grad}
    grads = dict(zip(leaves.keys(), torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)))
    for k, ref in m['grads'].items():
        assert grads[k] is not None, k
        assert_close(grads[k], ref, rtol=1e-4, atol=1e-6, what=f'{name}:grad:{k}')


def _oriented_inputs(m):
    """Batched edge cochains of the recorded inputs: (x, upper_index, lower_index, upper_orient, lower_orient, batch, y)."""
    xs, ups, los, uo, lo, batch, ys, off = [], [], [], [], [], [], [], 0
    for i, c in enumerate(m['inputs']):
        n = c['x'].size(0)
        xs.append(c['x'])
        if c['upper_index'] is not None:
            ups.append(c['upper_index'] + off)
            uo.append(c['upper_orient'])
        los.append(c['lower_index'] + off)
        lo.append(c['lower_orient'])
        batch.append(torch.full((n,), i, dtype=torch.long))
        ys.append(c['y'])
        off += n
    return (torch.cat(xs), torch.cat(ups, dim=1), torch.cat(los, dim=1), torch.cat(uo), torch.cat(lo), torch.cat(batch),
            torch.cat(ys))


@pytest.mark.parametrize('name', ['edge_orient_train', 'edge_mpnn_train'])
def test_oriented_edge_models_match_reference_forward_backward(name):
    """EdgeOrient / EdgeMPNN over OrientedConv (reference mp/models.py:474-608, mp/layers.py:430-470)."""
    m = golden()['models'][name]
    sd = oracle_state(m['state_dict'], requires_grad=True)
    x, up, low, uo, lo, batch, y = _oriented_inputs(m)
    out, cell_pred = O.oriented_edge_model(sd, m['cfg'], x, up, low, uo, lo, batch, len(m['inputs']),
                                           with_up=name.startswith('edge_orient'), training=True)
    assert_close(out, m['output'], atol=1e-6, what=name)
    assert_close(cell_pred, m['cell_pred'], atol=1e-6, what=name + ':cell_pred')
    loss = torch.nn.functional.l1_loss(out, y.view(-1, 1))
    assert_close(loss, m['loss'], atol=1e-6, what=name + ':loss')
    leaves = {k: v for k, v in sd.items() if v.requires_grad}
    grads = dict(zip(leaves.keys(), torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)))
    for k, ref in m['grads'].items():
        assert grads[k] is not None, k
        assert_close(grads[k], ref, rtol=1e-4, atol=1e-6, what=f'{name}:grad:{k}')
