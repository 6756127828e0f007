"""Dependency-free ring lifting (`cwn_b200/data/lifting.py`) against the known answers of the
reference's own tests (`data/test_utils.py:215-420`, literal expected tensors restated here) and a brute-force
induced-cycle enumeration. CPU only."""
import itertools

import numpy as np
import pytest
import torch

from cwn_b200.data import synthetic
from cwn_b200.data.complex import ComplexBatch
from cwn_b200.data.lifting import compute_ring_2complex, convert_graph_dataset_with_rings, find_rings

HOUSE = torch.tensor([[0, 0, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4],
                      [1, 3, 0, 2, 1, 3, 4, 0, 2, 4, 2, 3]], dtype=torch.long)  # data/test_utils.py:35-37
T = lambda v, dt=torch.long: torch.tensor(v, dtype=dt)  # noqa: E731


def _check_house_structure(c):
    assert c.nodes.num_cells_down is None and c.nodes.num_cells_up == 6 and c.nodes.boundary_index is None
    assert c.edges.num_cells_down == 5 and c.edges.num_cells_up == 2
    assert list(c.edges.boundary_index.size()) == [2, 12]
    assert c.cochains[2].num_cells == 2 and c.cochains[2].num_cells_down == 6 and c.cochains[2].num_cells_up == 0
    assert list(c.cochains[2].boundary_index.size()) == [2, 7]
    v = c.get_cochain_params(dim=0)
    assert v.down_index is None
    assert torch.equal(v.up_index, T([[0, 1, 0, 3, 1, 2, 2, 3, 2, 4, 3, 4], [1, 0, 3, 0, 2, 1, 3, 2, 4, 2, 4, 3]]))
    assert v.kwargs['down_attr'] is None and v.kwargs['boundary_attr'] is None
    e = c.get_cochain_params(dim=1)
    assert torch.equal(e.up_index, T([[0, 1, 0, 2, 0, 3, 1, 2, 1, 3, 2, 3, 3, 4, 3, 5, 4, 5],
                                      [1, 0, 2, 0, 3, 0, 2, 1, 3, 1, 3, 2, 4, 3, 5, 3, 5, 4]]))
    assert torch.equal(e.down_index, T([[0, 1, 0, 2, 2, 3, 2, 4, 3, 4, 1, 3, 1, 5, 3, 5, 4, 5],
                                        [1, 0, 2, 0, 3, 2, 4, 2, 4, 3, 3, 1, 5, 1, 5, 3, 5, 4]]))
    assert torch.equal(torch.as_tensor(e.kwargs['down_attr']),
                       T([[0], [0], [1], [1], [2], [2], [2], [2], [2], [2], [3], [3], [3], [3], [3], [3], [4], [4]],
                         torch.float))
    assert torch.equal(e.kwargs['boundary_index'], T([[0, 1, 0, 3, 1, 2, 2, 3, 2, 4, 3, 4],
                                                      [0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5]]))
    t = c.get_cochain_params(dim=2)
    assert torch.equal(t.down_index, T([[0, 1], [1, 0]])) and t.up_index is None
    assert torch.equal(t.kwargs['boundary_index'], T([[0, 1, 2, 3, 3, 4, 5], [0, 0, 0, 0, 1, 1, 1]]))
    return v, e, t


def test_ring_2complex_of_the_house_graph():  # data/test_utils.py:215-289
    x = torch.arange(0, 5, dtype=torch.float).view(5, 1)
    c = compute_ring_2complex(x, HOUSE, None, 5, max_k=4, y=torch.tensor([1]), init_rings=True)
    v, e, t = _check_house_structure(c)
    assert torch.equal(v.x, x)
    assert torch.equal(torch.as_tensor(v.kwargs['up_attr']),
                       T([[1], [1], [3], [3], [3], [3], [5], [5], [6], [6], [7], [7]], torch.float))
    assert torch.equal(e.x, T([[1], [3], [3], [5], [6], [7]], torch.float))
    assert torch.equal(torch.as_tensor(e.kwargs['up_attr']), T([[6]] * 12 + [[9]] * 6, torch.float))
    assert torch.equal(e.kwargs['boundary_attr'], x)
    assert torch.equal(t.x, T([[6], [9]], torch.float))
    assert torch.equal(torch.as_tensor(t.kwargs['down_attr']), T([[5], [5]], torch.float))
    assert torch.equal(t.kwargs['boundary_attr'], e.x)
    assert torch.equal(c.y, torch.tensor([1]))


def test_ring_2complex_with_edge_features():  # data/test_utils.py:292-398
    edge_attr = torch.FloatTensor([[0, 1], [0, 3], [0, 1], [1, 2], [1, 2], [2, 3], [2, 4], [0, 3], [2, 3], [3, 4],
                                   [2, 4], [3, 4]])
    x = torch.arange(0, 5, dtype=torch.float).view(5, 1)
    c = compute_ring_2complex(x, HOUSE, edge_attr, 5, max_k=4, y=torch.tensor([1]), init_rings=False)
    v, e, t = _check_house_structure(c)
    ex = torch.FloatTensor([[0, 1], [0, 3], [1, 2], [2, 3], [2, 4], [3, 4]])
    assert torch.equal(e.x, ex)
    assert torch.equal(torch.as_tensor(v.kwargs['up_attr']), ex.repeat_interleave(2, dim=0))
    assert e.kwargs['up_attr'] is None and t.x is None
    assert torch.equal(torch.as_tensor(t.kwargs['down_attr']), torch.FloatTensor([[2, 3], [2, 3]]))
    assert torch.equal(t.kwargs['boundary_attr'], ex)


def test_larger_max_k_changes_nothing_and_smaller_drops_rings():  # data/test_utils.py:401-420
    x = torch.arange(0, 5, dtype=torch.float).view(5, 1)
    a = compute_ring_2complex(x, HOUSE, None, 5, max_k=4, y=torch.tensor([1]), init_rings=True)
    b = compute_ring_2complex(x, HOUSE, None, 5, max_k=10, y=torch.tensor([1]), init_rings=True)
    for d in range(3):
        for k in ('x', 'upper_index', 'lower_index', 'boundary_index', 'shared_boundaries', 'shared_coboundaries'):
            u, w = getattr(a.cochains[d], k), getattr(b.cochains[d], k)
            assert (u is None and w is None) or torch.equal(u, w), (d, k)
    only_triangle = compute_ring_2complex(x, HOUSE, None, 5, max_k=3, init_rings=True)
    assert only_triangle.cochains[2].num_cells == 1
    assert torch.equal(only_triangle.cochains[2].x, T([[9]], torch.float))


def _brute_force_rings(n, edges, max_k):
    adj = [[False] * n for _ in range(n)]
    for a, b in edges:
        adj[a][b] = adj[b][a] = True
    out = set()
    for k in range(3, max_k + 1):
        for sub in itertools.combinations(range(n), k):
            if any(sum(adj[v][w] for w in sub) != 2 for v in sub):
                continue
            seen, stack = {sub[0]}, [sub[0]]      # 2-regular induced subgraph: a ring iff connected
            while stack:
                v = stack.pop()
                for w in sub:
                    if adj[v][w] and w not in seen:
                        seen.add(w)
                        stack.append(w)
            if len(seen) == k:
                out.add(sub)
    return out


@pytest.mark.parametrize('seed', range(6))
def test_find_rings_equals_brute_force_on_random_graphs(seed):
    rng = np.random.default_rng(seed)
    for _ in range(25):
        n = int(rng.integers(3, 10))
        edges = [(a, b) for a in range(n) for b in range(a + 1, n) if rng.random() < rng.uniform(0.15, 0.6)]
        ei = np.array(edges + [(b, a) for a, b in edges], dtype=np.int64).reshape(-1, 2).T
        for max_k in (3, 5, 7):
            got = find_rings(ei, max_k)
            assert len(got) == len(set(got))
            for r in got:  # cycle order: consecutive vertices adjacent, smallest vertex first
                assert r[0] == min(r) and all((min(r[i], r[(i + 1) % len(r)]), max(r[i], r[(i + 1) % len(r)])) in set(edges)
                                              for i in range(len(r)))
            assert {tuple(sorted(r)) for r in got} == _brute_force_rings(n, edges, max_k)


def test_lifting_recovers_the_synthetic_generator_and_batches():
    """The complexes `synthetic.zinc_like_complexes` builds from KNOWN rings equal what lifting the bare graphs finds,
    up to the numbering of the rings; and lifted complexes batch like any other."""
    rng = np.random.default_rng(3)
    comps = []
    for _ in range(12):
        sizes = tuple(int(s) for s in rng.integers(5, 7, size=int(rng.integers(0, 4))))
        n, edges, rings = synthetic.molecule_graph(sizes, int(rng.integers(2, 9)), rng)
        ei = torch.tensor(edges + [(b, a) for a, b in edges], dtype=torch.long).t()
        x = torch.randn(n, 3)
        c = compute_ring_2complex(x, ei, None, n, max_k=6, y=torch.randn(1), include_down_adj=False)
        ref = synthetic.make_complex(n, edges, sorted(rings, key=lambda r: tuple(sorted(r))), x, None)
        assert c.dimension == ref.dimension
        for d in range(c.dimension + 1):
            for k in ('upper_index', 'boundary_index', 'shared_coboundaries'):
                u, w = getattr(c.cochains[d], k), getattr(ref.cochains[d], k)
                assert (u is None and w is None) or torch.equal(u, w), (d, k)
        comps.append(c)
    batch = ComplexBatch.from_complex_list(comps, max_dim=2)
    assert batch.num_complexes == 12 and batch.cochains[0].x.size(1) == 3


def test_convert_graph_dataset_with_rings():
    class G(object):
        pass
    ds = []
    for seed in range(4):
        n, edges, _ = synthetic.molecule_graph((6, 5), 3, np.random.default_rng(seed))
        g = G()
        g.x, g.num_nodes, g.y = torch.randn(n, 2), n, torch.randn(1)
        g.edge_index = torch.tensor(edges + [(b, a) for a, b in edges], dtype=torch.long).t()
        g.edge_attr = None
        ds.append(g)
    complexes, dim, nf = convert_graph_dataset_with_rings(ds, max_ring_size=6, init_rings=True)
    assert len(complexes) == 4 and dim == 2 and nf == [2, 2, 2]


def test_models_are_invariant_to_the_numbering_of_the_rings(monkeypatch):
    """The reference numbers rings in CPython-set order; `find_rings` numbers them lexicographically. Any numbering is a
    relabelling of the 2-cells, and the models must not care: SparseCIN (host layer on the CPU, ops substituted) gives
    the same output for the lifted complex and for the same complex with its rings in reversed order."""
    import cpu_ops_shim
    from cwn_b200.mp.models import SparseCIN
    cpu_ops_shim.install(monkeypatch)
    torch.manual_seed(0)
    model = SparseCIN(num_input_features=3, num_classes=2, num_layers=2, hidden=8, max_dim=2,
                      use_coboundaries=True).eval()
    rng = np.random.default_rng(11)
    outs = []
    graphs = [synthetic.molecule_graph((6, 5, 6), 4, rng) for _ in range(3)]
    xs = [torch.randn(n, 3) for n, _, _ in graphs]
    for flip in (False, True):
        comps = []
        for (n, edges, _), x in zip(graphs, xs):
            ei = torch.tensor(edges + [(b, a) for a, b in edges], dtype=torch.long).t()
            rings = find_rings(ei, 6)
            rings = list(reversed(rings)) if flip else rings
            ex = torch.stack([x[a] + x[b] for a, b in sorted({(min(a, b), max(a, b)) for a, b in edges})])
            rx = torch.stack([x[list(r)].sum(0) for r in rings])
            c = synthetic.make_complex(n, edges, rings, x, ex)
            c.cochains[2].x = rx
            comps.append(c)
        with torch.no_grad():
            outs.append(model(ComplexBatch.from_complex_list(comps, max_dim=2)))
    assert torch.allclose(outs[0], outs[1], rtol=1e-5, atol=1e-5)
