"""Small hand-checkable cell complexes (house, kite, square, pyramid, ...) used by the known-answer tests.

Same complexes, features and labels as the reference's `data/dummy_complexes.py` fixtures (house :45-102,
bridged :105-173, fullstop :176, colon :191, square :208, square-dot :249, kite :290, pyramid :353,
filled-square :458, molecular :511), but DERIVED from each complex's boundary lists instead of being typed in as
index tensors: `cell_complex` computes boundary / upper / lower adjacencies and the shared (co)boundary columns
with the pair-enumeration convention of the lifting code (`data/utils.py:103-138`). The column ORDER of an
adjacency may therefore differ from the hand-written reference tensors; as multisets of
(source, destination, shared cell) they are identical (tests/test_data_api.py checks this against the golden
copies of the reference fixtures), and message passing sums over columns, so every known answer is unchanged.
"""
from itertools import combinations
from typing import List, Optional, Sequence

import torch

from cwn_b200.data.complex import Cochain, Complex


def _pairs(groups):
    src, dst, shared = [], [], []
    for cell, members in groups:
        for a, b in combinations(members, 2):
            src += [a, b]
            dst += [b, a]
            shared += [cell, cell]
    if not src:
        return None, None
    return torch.tensor([src, dst], dtype=torch.long), torch.tensor(shared, dtype=torch.long)


def cell_complex(num_vertices: int, boundaries: Sequence[Sequence[Sequence[int]]] = (), labels=None, y=None,
                 listing: Optional[dict] = None) -> Complex:
    """`boundaries[d-1][c]` = ids of the (d-1)-cells bounding d-cell `c`. Features of every dimension are
    1..num_cells (as floats, one column), cell labels `labels[d]` (default: the dimension), complex label = number
    of vertices. `listing[d]` optionally gives the order in which d-cells appear in their boundary_index."""
    counts = [num_vertices] + [len(b) for b in boundaries]
    top = len(counts) - 1
    cochains = []
    for d, n in enumerate(counts):
        kw = {}
        if d < top:  # upper adjacency: two d-cells on the boundary of the same (d+1)-cell
            kw['upper_index'], kw['shared_coboundaries'] = _pairs(list(enumerate(boundaries[d])))
        if d > 0:
            cells = boundaries[d - 1]
            order = (listing or {}).get(d, range(n))
            kw['boundary_index'] = torch.tensor([[b for c in order for b in cells[c]],
                                                 [c for c in order for _ in cells[c]]], dtype=torch.long)
            cofaces = [[] for _ in range(counts[d - 1])]
            for c, members in enumerate(cells):
                for b in members:
                    cofaces[b].append(c)
            kw['lower_index'], kw['shared_boundaries'] = _pairs(list(enumerate(cofaces)))
        lab = d if labels is None else labels[d]
        cochains.append(Cochain(dim=d, x=torch.arange(1, n + 1, dtype=torch.float).view(-1, 1),
                                y=torch.full((n,), lab, dtype=torch.long), **kw))
    y = torch.LongTensor([num_vertices]) if y is None else y
    return Complex(*cochains, y=y)


def get_house_complex():
    """Square 0-1-2-3 with the filled triangle 2-3-4 on top: 5 vertices, 6 edges, one 2-cell."""
    return cell_complex(5, [[[0, 1], [1, 2], [2, 3], [0, 3], [3, 4], [2, 4]], [[2, 4, 5]]])


def get_bridged_complex():
    """Square 0-1-2-3 with the path 3-4-1 across it; rings 0-1-4-3, 1-2-3-4 and 0-1-2-3 are all filled, so pairs
    of edges share TWO rings (replicated adjacencies must be summed with multiplicity)."""
    return cell_complex(5, [[[0, 1], [1, 2], [2, 3], [0, 3], [3, 4], [1, 4]],
                            [[0, 3, 4, 5], [1, 2, 4, 5], [0, 1, 2, 3]]])


def get_fullstop_complex():
    return cell_complex(1)


def get_colon_complex():
    return cell_complex(2)


def get_square_complex():
    return cell_complex(4, [[[0, 1], [1, 2], [2, 3], [0, 3]]])


def get_square_dot_complex():
    """The square plus an isolated vertex."""
    return cell_complex(5, [[[0, 1], [1, 2], [2, 3], [0, 3]]])


def get_kite_complex():
    """Triangles 0-1-2 and 1-2-3 (both filled) with the tail 3-4."""
    return cell_complex(5, [[[0, 1], [1, 2], [0, 2], [1, 3], [2, 3], [3, 4]], [[0, 1, 2], [1, 3, 4]]])


def get_pyramid_complex():
    """Filled tetrahedron: 4 vertices, 6 edges, 4 triangles, one 3-cell."""
    return cell_complex(4, [[[0, 1], [1, 2], [0, 2], [1, 3], [2, 3], [0, 3]],
                            [[0, 1, 2], [0, 3, 5], [1, 3, 4], [2, 4, 5]], [[0, 1, 2, 3]]], labels=[3, 1, 2, 3])


def get_filled_square_complex():
    return cell_complex(4, [[[0, 1], [1, 2], [2, 3], [0, 3]], [[0, 1, 2, 3]]])


def get_molecular_complex():
    """A 4-ring 0-1-2-3 and a 5-ring 1-2-4-5-6 sharing the bond 1-2, plus the pendant atom 7 (edge listing of
    the boundary index is deliberately not in id order, as in the reference fixture)."""
    return cell_complex(8, [[[0, 1], [1, 2], [2, 3], [0, 3], [2, 4], [4, 5], [5, 6], [1, 6], [6, 7]],
                            [[0, 1, 2, 3], [1, 4, 5, 6, 7]]], listing={1: [0, 1, 2, 3, 7, 4, 5, 6, 8]})


def get_testing_complex_list():
    """Mixed-dimension list with many edge cases (reference `data/dummy_complexes.py:28-34`)."""
    g = globals()
    return [g[f'get_{n}_complex']() for n in (
        'fullstop', 'pyramid', 'house', 'kite', 'square', 'square_dot', 'square', 'fullstop', 'house', 'kite',
        'pyramid', 'bridged', 'square_dot', 'colon', 'filled_square', 'molecular', 'fullstop', 'colon', 'bridged',
        'colon', 'fullstop', 'fullstop', 'colon')]


def get_mol_testing_complex_list():
    g = globals()
    return [g[f'get_{n}_complex']() for n in (
        'house', 'kite', 'square', 'fullstop', 'bridged', 'square_dot', 'square', 'filled_square', 'colon',
        'bridged', 'kite', 'square_dot', 'colon', 'molecular', 'bridged', 'filled_square', 'molecular',
        'fullstop', 'colon')]
