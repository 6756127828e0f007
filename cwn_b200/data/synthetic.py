"""Synthetic ring-lifted molecular complexes with the index layout of the reference's lifting code.

There is no network (no ZINC / ogbg-molhiv download) and no graph-tool / gudhi here, so the benchmark inputs are
generated: molecule-like graphs whose induced cycles are known by construction, lifted to 2-complexes with
exactly the conventions of `data/utils.py` (`build_tables` :45-65, `build_adj` :103-138, `generate_cochain`
:177-221, `compute_ring_2complex` :400-498; layout pinned by `data/test_utils.py:215-289`):

  * edges are the undirected edges in lexicographic (min, max) order;
  * `boundary_index` of a dimension is cell-major with non-decreasing row 1 (edge -> its 2 vertices, ring of
    size k -> its k edges, ascending);
  * `upper_index` of dimension d: for every (d+1)-cell in id order, for every unordered pair (a, b) of its
    boundary cells in combination order, the columns (a, b), (b, a); `shared_coboundaries` = that cell twice
    (so edge-level upper adjacencies are grouped by ring, destinations unsorted);
  * `lower_index` (optional): for every (d-1)-cell, pairs of its cofaces, `shared_boundaries` = that cell.

Default shape: V = 23, E = 25, rings {6, 6, 5}: three disjoint cycles joined by two bridge
edges plus six pendant/chain atoms; vertex ids are shuffled with a seeded permutation so that the edge order is
not trivially banded. ZINC vocabularies: 28 atom types, 4 bond types (`data/datasets/zinc.py:29-30`).
"""
from itertools import combinations
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from cwn_b200.data.complex import Cochain, Complex


def lift_rings(num_nodes: int, edges: Sequence[Tuple[int, int]], rings: Sequence[Sequence[int]],
               include_down_adj: bool = False):
    """Index tensors of the 2-complex of a graph whose rings (vertex cycles) are given.

    Returns a dict with, per dimension d in (0, 1, 2), `boundary_index{d}`, `upper_index{d}`,
    `shared_coboundaries{d}`, `lower_index{d}`, `shared_boundaries{d}` (int64 numpy arrays or None) and the
    sorted edge list."""
    edge_list = sorted({(min(a, b), max(a, b)) for a, b in edges})
    edge_id = {e: i for i, e in enumerate(edge_list)}
    out = {'edges': edge_list}

    def pair_columns(groups):
        """For each (cell id, ascending members): columns (a,b),(b,a) for every pair, tagged with the cell id."""
        src, dst, shared = [], [], []
        for cell, members in groups:
            for a, b in combinations(members, 2):
                src += [a, b]
                dst += [b, a]
                shared += [cell, cell]
        if not src:
            return None, None
        return np.array([src, dst], dtype=np.int64), np.array(shared, dtype=np.int64)

    # dimension 0: vertices are upper-adjacent through edges
    out['upper_index0'], out['shared_coboundaries0'] = pair_columns(
        [(i, list(e)) for i, e in enumerate(edge_list)])
    # dimension 1: boundaries = end points; upper adjacency through rings
    if edge_list:
        flat = np.array(edge_list, dtype=np.int64).reshape(-1)
        out['boundary_index1'] = np.stack([flat, np.repeat(np.arange(len(edge_list), dtype=np.int64), 2)])
    else:
        out['boundary_index1'] = None
    ring_edges = []
    for ring in rings:
        k = len(ring)
        ids = sorted(edge_id[(min(ring[i], ring[(i + 1) % k]), max(ring[i], ring[(i + 1) % k]))]
                     for i in range(k))
        ring_edges.append(ids)
    out['upper_index1'], out['shared_coboundaries1'] = pair_columns(list(enumerate(ring_edges)))
    if ring_edges:
        out['boundary_index2'] = np.stack([
            np.array([e for ids in ring_edges for e in ids], dtype=np.int64),
            np.array([r for r, ids in enumerate(ring_edges) for _ in ids], dtype=np.int64)])
    else:
        out['boundary_index2'] = None
    out['lower_index1'] = out['shared_boundaries1'] = out['lower_index2'] = out['shared_boundaries2'] = None
    if include_down_adj:
        incident = [[] for _ in range(num_nodes)]
        for i, (a, b) in enumerate(edge_list):
            incident[a].append(i)
            incident[b].append(i)
        out['lower_index1'], out['shared_boundaries1'] = pair_columns(list(enumerate(incident)))
        cofaces = [[] for _ in edge_list]
        for r, ids in enumerate(ring_edges):
            for e in ids:
                cofaces[e].append(r)
        out['lower_index2'], out['shared_boundaries2'] = pair_columns(list(enumerate(cofaces)))
    out['num_rings'] = len(ring_edges)
    return out


def molecule_graph(ring_sizes: Sequence[int] = (6, 6, 5), num_pendant: int = 6,
                   rng: Optional[np.random.Generator] = None):
    """A connected molecule-like graph: disjoint cycles of the given sizes chained by single bridge edges, plus
    `num_pendant` extra atoms attached one by one to random existing atoms (tree edges only, so the induced cycles
    are exactly the rings). Returns (num_nodes, edges, rings) with vertex ids shuffled by `rng`."""
    rng = rng or np.random.default_rng(0)
    edges, rings, n = [], [], 0
    for k in ring_sizes:
        ring = list(range(n, n + k))
        rings.append(ring)
        edges += [(ring[i], ring[(i + 1) % k]) for i in range(k)]
        n += k
    for a, b in zip(rings[:-1], rings[1:]):
        edges.append((a[len(a) // 2], b[0]))
    if n == 0:
        n = 1
    for _ in range(num_pendant):
        edges.append((int(rng.integers(0, n)), n))
        n += 1
    perm = rng.permutation(n)
    edges = [(int(perm[a]), int(perm[b])) for a, b in edges]
    rings = [[int(perm[v]) for v in ring] for ring in rings]
    return n, edges, rings


def make_complex(num_nodes, edges, rings, atom_x: torch.Tensor, bond_x: Optional[torch.Tensor], y=None,
                 include_down_adj=False, max_dim=2) -> Complex:
    """Assemble a `Complex` from a lifted graph (rings get no input features, as in the ZINC/OGB datasets:
    `init_rings=False`, `data/datasets/zinc.py:63-69`)."""
    L = lift_rings(num_nodes, edges, rings, include_down_adj)
    t = lambda a: None if a is None else torch.from_numpy(a)  # noqa: E731
    n_e, n_r = len(L['edges']), L['num_rings']
    v = Cochain(dim=0, x=atom_x, upper_index=t(L['upper_index0']),
                shared_coboundaries=t(L['shared_coboundaries0']), num_cells_up=n_e)
    cochains = [v]
    if n_e > 0 and max_dim >= 1:
        with_rings = n_r > 0 and max_dim >= 2
        e = Cochain(dim=1, x=bond_x, boundary_index=t(L['boundary_index1']),
                    upper_index=t(L['upper_index1']) if with_rings else None,
                    shared_coboundaries=t(L['shared_coboundaries1']) if with_rings else None,
                    lower_index=t(L['lower_index1']), shared_boundaries=t(L['shared_boundaries1']),
                    num_cells=n_e, num_cells_down=num_nodes, num_cells_up=n_r if with_rings else 0)
        cochains.append(e)
        if with_rings:
            r = Cochain(dim=2, x=None, boundary_index=t(L['boundary_index2']), lower_index=t(L['lower_index2']),
                        shared_boundaries=t(L['shared_boundaries2']), num_cells=n_r, num_cells_down=n_e,
                        num_cells_up=0)
            cochains.append(r)
    return Complex(*cochains, y=y)


def zinc_like_complexes(num: int, seed: int = 0, ring_sizes=(6, 6, 5), num_pendant: int = 6, atom_types: int = 28,
                        bond_types: int = 4, edge_features: bool = True, ragged: bool = False,
                        include_down_adj: bool = False, ogb_features: bool = False, ring_count_range=(0, 4),
                        pendant_range=(2, 12)) -> List[Complex]:
    """`num` seeded synthetic molecules. `ragged=True` varies ring count/sizes and chain length per molecule
    (ring sizes 5..6 as with max_ring=6; ring count in [ring_count_range), chain atoms in [pendant_range));
    `ogb_features=True` emits 9 integer atom columns / 3 bond columns in the ogbg-mol* vocabularies instead of the
    scalar ZINC types."""
    from cwn_b200.mp.encoders import ATOM_FEATURE_DIMS, BOND_FEATURE_DIMS
    rng = np.random.default_rng(seed)
    gen = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(num):
        sizes, pend = tuple(ring_sizes), num_pendant
        if ragged:
            sizes = tuple(int(s) for s in rng.integers(5, 7, size=int(rng.integers(*ring_count_range))))
            pend = int(rng.integers(*pendant_range))
        n, edges, rings = molecule_graph(sizes, pend, rng)
        n_e = len({(min(a, b), max(a, b)) for a, b in edges})
        if ogb_features:
            atom_x = torch.stack([torch.randint(0, d, (n,), generator=gen) for d in ATOM_FEATURE_DIMS], 1)
            bond_x = torch.stack([torch.randint(0, d, (n_e,), generator=gen) for d in BOND_FEATURE_DIMS], 1) \
                if edge_features and n_e > 0 else None
        else:
            atom_x = torch.randint(0, atom_types, (n, 1), generator=gen).float()
            bond_x = torch.randint(0, bond_types, (n_e, 1), generator=gen).float() \
                if edge_features and n_e > 0 else None
        y = torch.randn(1, generator=gen)
        out.append(make_complex(n, edges, rings, atom_x, bond_x, y=y, include_down_adj=include_down_adj))
    return out


def float_feature_complexes(num: int, num_features: int, seed: int = 0, **kwargs) -> List[Complex]:
    """Same molecules with N(0,1) float features on every dimension (inputs of `SparseCIN` / `CIN0`)."""
    gen = torch.Generator().manual_seed(seed + 1)
    comps = zinc_like_complexes(num, seed=seed, **kwargs)
    for comp in comps:
        for d in range(comp.dimension + 1):
            c = comp.cochains[d]
            c._x = torch.randn(c.num_cells, num_features, generator=gen)
    return comps


# ------------------------------------------------------------------------------------------------ kernel sweep
def tiled_adjacency(kind: str, n_units: int, seed: int = 0):
    """Block-diagonal adjacency of `n_units` copies of the unit molecule, built directly as tensors (config 5:
    10k..1M cells per dimension). kind in {'edge_boundary', 'ring_boundary', 'vertex_up', 'edge_up'}.

    Returns (index int64 [2, E], cob int64 [E] or None, n_src, n_dst, n_cob)."""
    rng = np.random.default_rng(seed)
    n, edges, rings = molecule_graph((6, 6, 5), 6, rng)
    L = lift_rings(n, edges, rings)
    n_v, n_e, n_r = n, len(L['edges']), L['num_rings']
    table = {
        'edge_boundary': (L['boundary_index1'], None, n_v, n_e, 0),
        'ring_boundary': (L['boundary_index2'], None, n_e, n_r, 0),
        'vertex_up': (L['upper_index0'], L['shared_coboundaries0'], n_v, n_v, n_e),
        'edge_up': (L['upper_index1'], L['shared_coboundaries1'], n_e, n_e, n_r),
    }
    idx, cob, us, ud, uc = table[kind]
    idx = torch.from_numpy(idx)
    units = torch.arange(n_units, dtype=torch.long)
    off = torch.stack([units * us, units * ud]).unsqueeze(-1)           # [2, n_units, 1]
    index = (idx.unsqueeze(1) + off).reshape(2, -1).contiguous()
    cob_t = None
    if cob is not None:
        cob_t = (torch.from_numpy(cob).unsqueeze(0) + (units * uc).unsqueeze(-1)).reshape(-1).contiguous()
    return index, cob_t, us * n_units, ud * n_units, uc * n_units


def random_adjacency(n_src: int, n_dst: int, E: int, n_cob: int = 0, seed: int = 0):
    """Uniform-random stress layout (secondary layout of config 5)."""
    gen = torch.Generator().manual_seed(seed)
    index = torch.stack([torch.randint(0, n_src, (E,), generator=gen), torch.randint(0, n_dst, (E,), generator=gen)])
    cob = torch.randint(0, n_cob, (E,), generator=gen) if n_cob else None
    return index, cob, n_src, n_dst, n_cob
