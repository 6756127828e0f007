"""Data API (Cochain / CochainBatch / Complex / ComplexBatch) against the reference's own batching, run on the
reference's fixtures (golden). CPU only — this is host-side plumbing."""
import pytest
import torch

from cwn_b200.data import dummy_complexes as dc
from cwn_b200.data.complex import Cochain, Complex, ComplexBatch
from cwn_b200.mp.params import LazyRows
from helpers import COCHAIN_KEYS, batch_of, fixture, golden


def _multiset(idx, shared=None):
    if idx is None:
        return None
    cols = idx if shared is None else torch.cat([idx, shared.view(1, -1)], 0)
    return sorted(map(tuple, cols.t().tolist()))


@pytest.mark.parametrize('name', list(golden()['fixtures']))
def test_derived_fixtures_equal_reference_fixtures(name):
    ref = golden()['fixtures'][name]
    mine = getattr(dc, f'get_{name}_complex')()
    assert mine.dimension == ref['dimension'] and torch.equal(mine.y, ref['y'])
    for d, rc in ref['cochains'].items():
        mc = mine.cochains[d]
        assert torch.equal(mc.x, rc['x']) and torch.equal(mc.y, rc['y'])
        assert _multiset(mc.upper_index, mc.shared_coboundaries) == _multiset(rc['upper_index'], rc['shared_coboundaries'])
        assert _multiset(mc.lower_index, mc.shared_boundaries) == _multiset(rc['lower_index'], rc['shared_boundaries'])
        assert _multiset(mc.boundary_index) == _multiset(rc['boundary_index'])
        assert (mc.num_cells, mc.num_cells_up, mc.num_cells_down) == (rc['num_cells'], rc['num_cells_up'], rc['num_cells_down'])


@pytest.mark.parametrize('key', list(golden()['batching']))
def test_batching_reproduces_reference_tensors_exactly(key):
    max_dim, bs = key
    names = golden()['testing_list']
    chunks = [names[i:i + bs] for i in range(0, len(names), bs)]
    for chunk, ref in zip(chunks, golden()['batching'][key]):
        batch = batch_of(chunk, max_dim=max_dim)
        assert batch.dimension == ref['dimension'] and batch.num_complexes == ref['num_complexes']
        assert torch.equal(batch.y, ref['y'])
        for d, rc in ref['cochains'].items():
            mc = batch.cochains[d]
            for k in COCHAIN_KEYS + ['batch']:
                a, b = getattr(mc, k), rc[k]
                assert (a is None) == (b is None), (d, k)
                if a is not None:
                    assert torch.equal(a, b), (d, k)
            assert torch.equal(mc.ptr, rc['ptr'])
            assert (mc.num_cells, mc.num_cells_up, mc.num_cells_down) == (rc['num_cells'], rc['num_cells_up'], rc['num_cells_down'])


def test_house_pair_batching_known_answer():  # reference data/test_batching.py (two houses)
    batch = ComplexBatch.from_complex_list([fixture('house'), fixture('house')])
    v, e, t = (batch.cochains[d] for d in range(3))
    assert torch.equal(v.x, torch.tensor([1, 2, 3, 4, 5] * 2, dtype=torch.float).view(-1, 1))
    h = fixture('house')
    assert torch.equal(v.upper_index, torch.cat([h.cochains[0].upper_index, h.cochains[0].upper_index + 5], 1))
    assert torch.equal(v.shared_coboundaries, torch.cat([h.cochains[0].shared_coboundaries,
                                                         h.cochains[0].shared_coboundaries + 6]))
    off = torch.tensor([[5], [6]])
    assert torch.equal(e.boundary_index, torch.cat([h.cochains[1].boundary_index, h.cochains[1].boundary_index + off], 1))
    assert torch.equal(e.shared_boundaries, torch.cat([h.cochains[1].shared_boundaries, h.cochains[1].shared_boundaries + 5]))
    assert torch.equal(t.boundary_index, torch.tensor([[2, 4, 5, 8, 10, 11], [0, 0, 0, 1, 1, 1]]))
    assert torch.equal(v.batch, torch.tensor([0] * 5 + [1] * 5)) and torch.equal(t.batch, torch.tensor([0, 1]))
    assert batch.num_complexes == 2 and (v.num_cells, e.num_cells, t.num_cells) == (10, 12, 2)


def test_cochain_params_known_answers():  # reference data/test_data.py:6-54
    house = dc.get_house_complex()
    v = house.get_cochain_params(dim=0)
    up_attr = v.kwargs['up_attr']
    assert isinstance(up_attr, LazyRows)
    got = sorted(zip(v.up_index.t().tolist(), up_attr.view(-1).tolist()))
    ref_idx = torch.tensor([[0, 1, 0, 3, 1, 2, 2, 3, 2, 4, 3, 4], [1, 0, 3, 0, 2, 1, 3, 2, 4, 2, 4, 3]])
    ref = sorted(zip(ref_idx.t().tolist(), [1., 1., 4., 4., 2., 2., 3., 3., 6., 6., 5., 5.]))
    assert got == ref
    # on the reference's own tensors the values are literally those of the reference test
    h = fixture('house')
    assert torch.equal(h.get_cochain_params(0).kwargs['up_attr'],
                       torch.tensor([[1], [1], [4], [4], [2], [2], [3], [3], [6], [6], [5], [5]], dtype=torch.float))
    e = h.get_cochain_params(dim=1)
    assert torch.equal(torch.tensor([[1.]] * 6), e.kwargs['up_attr'])
    assert torch.equal(e.kwargs['down_attr'], torch.tensor(
        [2, 2, 1, 1, 3, 3, 3, 3, 4, 4, 4, 4, 3, 3, 4, 4, 5, 5], dtype=torch.float).view(-1, 1))
    t = h.get_cochain_params(dim=2)
    assert t.kwargs['up_attr'] is None and t.kwargs['down_attr'] is None
    params = h.get_all_cochain_params(max_dim=1, include_top_features=False)
    assert len(params) == 2 and params[1].kwargs['up_attr'] is None and params[1].up_index.size(1) == 6
    assert params[1].boundary_attr is h.cochains[0].x


def test_set_xs_and_counts():  # reference data/test_batching.py:530-546
    batch = batch_of(['house', 'square', 'kite'])
    xs = [torch.zeros(c.num_cells, 7) for c in (batch.cochains[d] for d in range(3))]
    batch.set_xs(xs)
    assert all(batch.cochains[d].x is xs[d] for d in range(3))
    with pytest.raises(AssertionError):
        batch.set_xs([torch.zeros(3, 7)])
    assert batch.cochains[0].num_cells_up == batch.cochains[1].num_cells
    assert batch.cochains[2].num_cells_down == batch.cochains[1].num_cells


def test_padding_cochains_keep_offsets():
    # fullstop has no edges: the edge batch must still offset the next complex's boundaries by its vertex count
    batch = batch_of(['fullstop', 'square'])
    e = batch.cochains[1]
    assert torch.equal(e.boundary_index[0], fixture('square').cochains[1].boundary_index[0] + 1)
    assert torch.equal(e.batch, torch.ones(4, dtype=torch.long))
    assert batch.cochains[0].num_cells == 5 and e.num_cells == 4 and e.num_cells_down == 5


def test_lazy_rows_behaves_like_a_tensor():
    src = torch.arange(12.).view(6, 2)
    idx = torch.tensor([5, 0, 0, 3])
    lazy = LazyRows(src, idx)
    assert lazy.size(0) == 4 and lazy.shape == (4, 2) and len(lazy) == 4
    assert torch.equal(lazy, src[idx]) and torch.equal(src[idx], lazy)
    assert torch.equal(lazy + 1, src[idx] + 1) and torch.equal(torch.cat([lazy, lazy], -1), torch.cat([src[idx]] * 2, -1))


def test_complex_constructor_errors():
    with pytest.raises(ValueError):
        Complex()
    with pytest.raises(ValueError):
        Complex(Cochain(dim=0, x=torch.ones(2, 1)), dimension=1)
    with pytest.raises(AssertionError):
        Cochain(dim=0, boundary_index=torch.zeros(2, 1, dtype=torch.long))
