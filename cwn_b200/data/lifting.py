"""Ring lifting of a graph to a 2-complex WITHOUT graph-tool / gudhi.

Reference: `compute_ring_2complex` and its helpers, `data/utils.py:300-498` (rings = induced cycles found by graph-tool
subgraph isomorphism against cycle patterns, `:300-330`; vertices/edges from a gudhi simplex tree, `:45-65`; adjacencies
`build_adj` `:100-138`; features `construct_features` `:141-156`; cochains `generate_cochain` `:176-226`). Neither
library is available offline, and both are only used as enumeration engines, so the enumeration is restated:

* edges: the undirected edges in lexicographic order of (min, max) — gudhi's simplex-tree order;
* rings: the chordless cycles of length 3..max_k. A depth-first search grows chordless paths from their smallest
  vertex; a path closes into a ring when its last vertex is adjacent to the start, and is abandoned as soon as a new
  vertex touches any earlier path vertex (that edge would be a chord). Molecular graphs have a handful of small rings,
  so the search is tiny; pure preprocessing, off the GPU path.

Ring ids: the reference keeps rings in a Python `set` of graph-tool's isomorphism tuples and numbers them in that
set's iteration order — a function of CPython's tuple hashing and graph-tool's vertex mapping, not of the graph. Here
rings are numbered by the lexicographic order of their sorted vertex tuples, which reproduces the reference's own
known-answer test (`data/test_utils.py:215-289`: square before triangle on the house graph). Any other numbering is a
relabelling of the 2-cells, to which the models are invariant (`tests/test_lifting.py`).
"""
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
from torch import Tensor

from cwn_b200.data.complex import Cochain, Complex
from cwn_b200.data.synthetic import lift_rings


def find_rings(edge_index: Union[Tensor, np.ndarray], max_k: int = 7) -> List[Tuple[int, ...]]:
    """Chordless cycles with 3..max_k vertices of the undirected simple graph behind `edge_index` ([2, E]; self loops
    and parallel / reversed duplicates are ignored, as `data/utils.py:307-308` does). Each ring is a vertex tuple in
    cycle order starting at its smallest vertex; rings are sorted by their sorted vertex tuples."""
    ei = edge_index.numpy() if isinstance(edge_index, Tensor) else np.asarray(edge_index)
    adj = {}
    for a, b in ei.T.tolist():
        if a == b:
            continue
        adj.setdefault(a, set()).add(b)
        adj.setdefault(b, set()).add(a)
    rings = []

    def extend(path, blocked):
        """`path` = chordless path from its smallest vertex s; `blocked` = vertices adjacent to path[1:-1] (a chord)."""
        s, last = path[0], path[-1]
        for u in sorted(adj[last]):
            if u <= s or u in blocked or u in path:
                continue
            if len(path) >= 2 and u in adj[s]:
                # closes the cycle; orientation fixed by path[1] < u so that each ring is found once
                if path[1] < u:
                    rings.append(tuple(path + [u]))
                continue  # going on through u would keep the chord (u, s)
            if len(path) + 1 < max_k:
                extend(path + [u], blocked | (adj[last] - {u}) if len(path) >= 2 else blocked)

    for s in sorted(adj):
        for v in sorted(adj[s]):
            if v > s:
                # neighbours of s other than the path's second vertex may only appear as the closing vertex, which
                # the `u in adj[s]` branch handles; vertices adjacent to interior path vertices are blocked
                extend([s, v], set())
    return sorted(rings, key=lambda r: tuple(sorted(r)))


def _cell_features(vx: Tensor, cells: Sequence[Sequence[int]], init_method: str) -> Tensor:
    """`construct_features` for one dimension: sum / mean / ... of the member vertices' features."""
    out = torch.zeros(len(cells), vx.size(1), dtype=vx.dtype)
    if init_method in ('sum', 'add', 'mean'):
        for c, cell in enumerate(cells):
            v = vx[list(cell)].sum(dim=0)
            out[c] = v / len(cell) if init_method == 'mean' else v
        return out
    if init_method == 'max':
        for c, cell in enumerate(cells):
            out[c] = vx[list(cell)].max(dim=0)[0]
        return out
    if init_method == 'min':
        for c, cell in enumerate(cells):
            out[c] = vx[list(cell)].min(dim=0)[0]
        return out
    raise ValueError(f'unknown init_method {init_method!r}')


def compute_ring_2complex(x: Union[Tensor, np.ndarray], edge_index: Union[Tensor, np.ndarray],
                          edge_attr: Optional[Union[Tensor, np.ndarray]], size: int,
                          y: Optional[Union[Tensor, np.ndarray]] = None, max_k: int = 7, include_down_adj=True,
                          init_method: str = 'sum', init_edges=True, init_rings=False) -> Complex:
    """Ring 2-complex of a graph; same signature and semantics as the reference's (`data/utils.py:414-498`)."""
    assert x is not None
    assert isinstance(edge_index, (np.ndarray, Tensor))
    x = torch.as_tensor(x)
    edge_index = torch.as_tensor(edge_index)
    edge_attr = None if edge_attr is None else torch.as_tensor(edge_attr)
    y = None if y is None else torch.as_tensor(y)

    pairs = [(int(a), int(b)) for a, b in edge_index.t().tolist() if a != b]
    rings = find_rings(edge_index, max_k=max_k) if pairs else []
    L = lift_rings(size, pairs, rings, include_down_adj=include_down_adj)
    edges = L['edges']
    n_e, n_r = len(edges), L['num_rings']
    complex_dim = 2 if n_r > 0 else (1 if n_e > 0 else 0)
    t = lambda a: None if a is None else torch.from_numpy(a)  # noqa: E731

    xs = [x, None, None]
    if init_rings and n_r > 0:
        xs[2] = _cell_features(x, rings, init_method)
    if init_edges and n_e > 0:
        if edge_attr is None:
            xs[1] = _cell_features(x, edges, init_method)
        else:
            if edge_attr.dim() == 1:
                edge_attr = edge_attr.view(-1, 1)
            edge_id = {e: i for i, e in enumerate(edges)}
            ex = {}
            for e, (a, b) in enumerate(edge_index.t().tolist()):
                i = edge_id[(min(a, b), max(a, b))]
                if i in ex:
                    assert torch.equal(ex[i], edge_attr[e]), 'edge features must be undirected'
                else:
                    ex[i] = edge_attr[e]
            xs[1] = torch.stack([ex[i] for i in range(n_e)], dim=0)

    v_y = complex_y = None
    if y is not None:
        if y.size(0) == 1:
            complex_y = y
        else:
            assert y.size(0) == size
            v_y = y

    cochains = [Cochain(dim=0, x=xs[0], upper_index=t(L['upper_index0']),
                        shared_coboundaries=t(L['shared_coboundaries0']), y=v_y, num_cells_down=None,
                        num_cells_up=n_e if complex_dim >= 1 else 0)]
    if complex_dim >= 1:
        cochains.append(Cochain(dim=1, x=xs[1], boundary_index=t(L['boundary_index1']),
                                upper_index=t(L['upper_index1']), shared_coboundaries=t(L['shared_coboundaries1']),
                                lower_index=t(L['lower_index1']), shared_boundaries=t(L['shared_boundaries1']),
                                num_cells_down=size, num_cells_up=n_r if complex_dim == 2 else 0))
    if complex_dim == 2:
        cochains.append(Cochain(dim=2, x=xs[2], boundary_index=t(L['boundary_index2']),
                                lower_index=t(L['lower_index2']), shared_boundaries=t(L['shared_boundaries2']),
                                num_cells_down=n_e, num_cells_up=0))
    return Complex(*cochains, y=complex_y, dimension=complex_dim)


def convert_graph_dataset_with_rings(dataset, max_ring_size=7, include_down_adj=False, init_method: str = 'sum',
                                     init_edges=True, init_rings=False):
    """`data/utils.py:501-560` without joblib: `dataset` yields objects with x, edge_index, edge_attr, y, num_nodes.
    Returns (complexes, dimension, num_features per dimension)."""
    dimension, complexes, num_features = -1, [], [None, None, None]
    for data in dataset:
        comp = compute_ring_2complex(data.x, data.edge_index, getattr(data, 'edge_attr', None), data.num_nodes,
                                     y=data.y, max_k=max_ring_size, include_down_adj=include_down_adj,
                                     init_method=init_method, init_edges=init_edges, init_rings=init_rings)
        dimension = max(dimension, comp.dimension)
        for dim in range(comp.dimension + 1):
            nf = comp.cochains[dim].num_features
            if num_features[dim] is None:
                num_features[dim] = nf
            else:
                assert num_features[dim] == nf
        complexes.append(comp)
    return complexes, dimension, num_features[:dimension + 1]
