"""Generate the golden vectors of tests/golden/*.pt by RUNNING THE REFERENCE in the build container.

    PYTHONPATH=oracle/ref_shims:/root/reference python tests/golden/make_golden.py

The reference's own, unmodified `mp/*.py`, `data/complex.py`, `data/dummy_complexes.py` are imported from
/root/reference; its missing third-party dependencies are replaced by the stand-ins of oracle/ref_shims (see
oracle/ref_shims/README.md for exactly what is restated). /root/reference does not exist on the GPU box, so the
outputs are committed as small fixtures and this script documents how they were made. Inputs that are not the
reference's fixtures come from cwn_b200.data.synthetic (seeded).
"""
import os
import sys
import warnings

import torch

warnings.filterwarnings('ignore')
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'ref_shims'))
sys.path.insert(0, '/root/reference')

import data.dummy_complexes as ref_fix                                    # noqa: E402  (reference)
from data.complex import Cochain as RefCochain, Complex as RefComplex, ComplexBatch as RefBatch  # noqa: E402
from mp.cell_mp import CochainMessagePassing as RefCMP                    # noqa: E402
from mp.layers import DummyCellularMessagePassing as RefDummy, InitReduceConv as RefInitReduce  # noqa: E402
from mp.models import SparseCIN as RefSparseCIN, CIN0 as RefCIN0, CINpp as RefCINpp  # noqa: E402
from mp.models import EdgeCIN0 as RefEdgeCIN0, EdgeOrient as RefEdgeOrient, EdgeMPNN as RefEdgeMPNN  # noqa: E402
from data.complex import CochainBatch as RefCochainBatch  # noqa: E402
from mp.molec_models import (EmbedSparseCIN as RefEmbed, OGBEmbedSparseCIN as RefOGB,  # noqa: E402
                             EmbedCINpp as RefEmbedCINpp)

from cwn_b200.data import synthetic                                        # noqa: E402

FIXTURES = ['house', 'bridged', 'fullstop', 'colon', 'square', 'square_dot', 'kite', 'pyramid', 'filled_square',
            'molecular']
TESTING_LIST = ['fullstop', 'pyramid', 'house', 'kite', 'square', 'square_dot', 'square', 'fullstop', 'house',
                'kite', 'pyramid', 'bridged', 'square_dot', 'colon', 'filled_square', 'molecular', 'fullstop',
                'colon', 'bridged', 'colon', 'fullstop', 'fullstop', 'colon']  # data/dummy_complexes.py:28-34
COCHAIN_KEYS = ['x', 'upper_index', 'lower_index', 'boundary_index', 'shared_boundaries', 'shared_coboundaries', 'y']


def ref_complex(name):
    return getattr(ref_fix, f'get_{name}_complex')()


def dump_cochain(c):
    d = {k: getattr(c, k) for k in COCHAIN_KEYS}
    d['num_cells'], d['num_cells_up'], d['num_cells_down'] = c.num_cells, c.num_cells_up, c.num_cells_down
    for k in ('batch', 'ptr'):
        if hasattr(c, k):
            d[k] = getattr(c, k)
    return d


def dump_complex(c):
    return {'dimension': c.dimension, 'y': c.y, 'num_complexes': getattr(c, 'num_complexes', None),
            'cochains': {d: dump_cochain(c.cochains[d]) for d in range(c.dimension + 1)}}


def to_ref(comp):
    """cwn_b200 Complex (synthetic generator) -> reference Complex, attribute by attribute."""
    cochains = []
    for d in range(comp.dimension + 1):
        c = comp.cochains[d]
        kw = {k: getattr(c, k) for k in COCHAIN_KEYS}
        cochains.append(RefCochain(dim=d, num_cells=c.num_cells, num_cells_up=c.num_cells_up,
                                   num_cells_down=c.num_cells_down, **kw))
    return RefComplex(*cochains, y=comp.y)


def propagate_all(cmp_factory, comp):
    out = {}
    for d in range(comp.dimension + 1):
        p = comp.get_cochain_params(dim=d)
        cmp = cmp_factory()
        out[d] = cmp.propagate(p.up_index, p.down_index, p.boundary_index, x=p.x, up_attr=p.kwargs['up_attr'],
                               down_attr=p.kwargs['down_attr'], boundary_attr=p.kwargs['boundary_attr'])
    return out


def chunks(lst, n):
    return [lst[i:i + n] for i in range(0, len(lst), n)]


def grads_of(model, loss):
    model.zero_grad()
    loss.backward()
    return {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}


def main():
    torch.manual_seed(0)
    gold = {}

    # ---- A. the reference's hand-built fixtures
    gold['fixtures'] = {n: dump_complex(ref_complex(n)) for n in FIXTURES}
    gold['testing_list'] = TESTING_LIST

    # ---- B. known answers of the bare propagate and of the parameter-free Dummy layers on every fixture
    kat = {}
    for n in FIXTURES:
        entry = {'propagate': propagate_all(lambda: RefCMP(up_msg_size=1, down_msg_size=1), ref_complex(n))}
        for flags in [(False, True), (True, False), (True, True)]:
            comp = ref_complex(n)
            params = comp.get_all_cochain_params()
            layer = RefDummy(use_boundary_msg=flags[0], use_down_msg=flags[1])
            try:
                entry[f'dummy_b{int(flags[0])}_d{int(flags[1])}'] = layer.forward(*params)
            except Exception as e:  # e.g. missing features in a dimension
                entry[f'dummy_b{int(flags[0])}_d{int(flags[1])}'] = f'{type(e).__name__}'
        kat[n] = entry
    house = ref_complex('house')
    v, e, t = (house.get_cochain_params(dim=d) for d in range(3))
    kat['house']['init_reduce'] = [RefInitReduce('add').forward(v.x, e.boundary_index),
                                   RefInitReduce('add').forward(e.x, t.boundary_index)]
    for reduce in ('mean', 'max'):
        kat['house'][f'propagate_{reduce}'] = propagate_all(
            lambda: RefCMP(1, 1, aggr_up=reduce, aggr_down=reduce, aggr_boundary=reduce), ref_complex('house'))
    gold['kat'] = kat

    # ---- C. batching of the testing list
    batching = {}
    for max_dim in (1, 2, 3):
        for bs in (3, 5, 23):
            batching[(max_dim, bs)] = [dump_complex(RefBatch.from_complex_list(
                [ref_complex(n) for n in chunk], max_dim=max_dim)) for chunk in chunks(TESTING_LIST, bs)]
    gold['batching'] = batching

    # ---- D. models: weights, inputs, outputs (and gradients in training mode)
    models = {}

    def run_eval(name, model, cfg, lists, max_dim=2, strip_edge_ring_x=False, partial=True):
        model.eval()
        outs = []
        for chunk in lists:
            batch = RefBatch.from_complex_list([ref_complex(n) for n in chunk], max_dim=max_dim)
            if strip_edge_ring_x:
                for d in (1, 2):
                    if d in batch.cochains:
                        batch.cochains[d].x = None
            with torch.no_grad():
                outs.append(model.forward(batch, include_partial=True) if partial else model.forward(batch))
        models[name] = {'cfg': cfg, 'state_dict': {k: v.clone() for k, v in model.state_dict().items()},
                        'chunks': lists, 'max_dim': max_dim, 'strip': strip_edge_ring_x, 'outputs': outs}

    lists = chunks(TESTING_LIST, 4)
    cfg = dict(num_input_features=1, num_classes=3, num_layers=3, hidden=5, jump_mode='cat', max_dim=2)
    run_eval('sparse_cin_eval', RefSparseCIN(**cfg), cfg, lists)
    cfg = dict(num_input_features=1, num_classes=3, num_layers=2, hidden=6, jump_mode='max', max_dim=1,
               readout='mean', final_readout='mean', nonlinearity='tanh', graph_norm='ln')
    run_eval('sparse_cin_eval_dim1', RefSparseCIN(**cfg), cfg, lists, max_dim=1)
    cfg = dict(atom_types=32, bond_types=4, out_size=3, num_layers=3, hidden=5, jump_mode='cat', max_dim=2)
    run_eval('embed_sparse_cin_eval', RefEmbed(**cfg), cfg, lists, strip_edge_ring_x=True)
    cfg = dict(num_input_features=1, num_classes=3, num_layers=3, hidden=5, jump_mode='cat', max_dim=2)
    run_eval('cin0_eval', RefCIN0(**cfg), cfg, lists, partial=False)

    def run_train(name, model, cfg, comps, loss_fn, batch_hook=None):
        model.train()
        sd0 = {k: v.clone() for k, v in model.state_dict().items()}
        batch = RefBatch.from_complex_list([to_ref(c) for c in comps], max_dim=cfg.get('max_dim', 2))
        if batch_hook:
            batch_hook(batch)
        out = model.forward(batch)
        loss = loss_fn(out, batch.y)
        grads = grads_of(model, loss)
        models[name] = {'cfg': cfg, 'state_dict': sd0, 'inputs': [dump_complex(to_ref(c)) for c in comps],
                        'output': out.detach().clone(), 'loss': loss.detach().clone(), 'grads': grads,
                        'state_dict_after': {k: v.clone() for k, v in model.state_dict().items()}}

    l1 = lambda out, y: torch.nn.functional.l1_loss(out, y.view(-1, 1))  # noqa: E731  (exp/train_utils.py:25-26)

    cfg = dict(num_input_features=8, num_classes=1, num_layers=2, hidden=16, dropout_rate=0.0, max_dim=2,
               nonlinearity='elu', readout='mean', train_eps=True, use_coboundaries=True)
    run_train('sparse_cin_train', RefSparseCIN(**cfg), cfg,
              synthetic.float_feature_complexes(5, 8, seed=3, ragged=True), l1)

    cfg = dict(atom_types=28, bond_types=4, out_size=1, num_layers=2, hidden=16, dropout_rate=0.0, max_dim=2,
               embed_edge=True, use_coboundaries=True)
    run_train('embed_sparse_cin_train', RefEmbed(**cfg), cfg, synthetic.zinc_like_complexes(6, seed=1), l1)

    cfg = dict(atom_types=28, bond_types=4, out_size=1, num_layers=2, hidden=16, dropout_rate=0.0, max_dim=2,
               embed_edge=False, use_coboundaries=False, jump_mode='cat', readout='mean')
    run_train('embed_sparse_cin_train_nocob', RefEmbed(**cfg), cfg,
              synthetic.zinc_like_complexes(6, seed=2, ragged=True, edge_features=False), l1)

    cfg = dict(out_size=1, num_layers=2, hidden=16, dropout_rate=0.0, indropout_rate=0.0, max_dim=2,
               embed_edge=True, use_coboundaries=True, readout='mean')
    bce = lambda out, y: torch.nn.functional.binary_cross_entropy_with_logits(  # noqa: E731
        out, (y.view(-1, 1) > 0).float())                                       # exp/train_utils.py:23-24
    run_train('ogb_embed_sparse_cin_train', RefOGB(**cfg), cfg,
              synthetic.zinc_like_complexes(6, seed=4, ogb_features=True), bce)

    cfg = dict(num_input_features=4, num_classes=1, num_layers=2, hidden=8, dropout_rate=0.0, max_dim=2,
               jump_mode='cat')
    run_train('cin0_train', RefCIN0(**cfg), cfg,
              synthetic.float_feature_complexes(5, 4, seed=5, include_down_adj=True), l1)

    # CIN++ (appended last so that the random stream of everything above is unchanged)
    cfg = dict(num_input_features=8, num_classes=1, num_layers=2, hidden=16, dropout_rate=0.0, max_dim=2,
               nonlinearity='relu', train_eps=True, use_coboundaries=True)
    run_train('cinpp_train', RefCINpp(**cfg), cfg, synthetic.float_feature_complexes(5, 8, seed=6, ragged=True), l1)
    cfg = dict(atom_types=28, bond_types=4, out_size=1, num_layers=2, hidden=16, dropout_rate=0.0, max_dim=2,
               embed_edge=True, use_coboundaries=True, readout='mean')
    run_train('embed_cinpp_train', RefEmbedCINpp(**cfg), cfg, synthetic.zinc_like_complexes(6, seed=7), l1)

    # edge-level models, appended after everything else (random stream above unchanged)
    cfg = dict(num_input_features=4, num_classes=1, num_layers=3, hidden=8, dropout_rate=0.0, jump_mode='cat')
    run_train('edge_cin0_train', RefEdgeCIN0(**cfg), cfg,
              synthetic.float_feature_complexes(5, 4, seed=9, include_down_adj=True), l1)
    cfg = dict(num_input_features=4, num_classes=1, num_layers=2, hidden=8, dropout_rate=0.0,
               include_top_features=False)
    run_train('edge_cin0_notop_train', RefEdgeCIN0(**cfg), cfg,
              synthetic.float_feature_complexes(5, 4, seed=10, include_down_adj=True), l1)

    def run_oriented(name, model, cfg, comps, seed):
        """EdgeOrient / EdgeMPNN on the batched EDGE cochains of `comps`, with seeded +-1 relative orientations on
        every upper / lower adjacency message and a signed edge signal (the flow datasets' setting)."""
        gen = torch.Generator().manual_seed(seed)
        cochains = []
        for comp in comps:
            e = to_ref(comp).cochains[1]
            sign = lambda n: (torch.randint(0, 2, (n,), generator=gen) * 2 - 1).float()  # noqa: E731
            up_o = sign(e.upper_index.size(1)) if e.upper_index is not None else None
            lo_o = sign(e.lower_index.size(1))
            cochains.append(RefCochain(dim=1, x=e.x, upper_index=e.upper_index, lower_index=e.lower_index,
                                       shared_boundaries=e.shared_boundaries, shared_coboundaries=e.shared_coboundaries,
                                       upper_orient=up_o, lower_orient=lo_o, y=comp.y,
                                       num_cells_down=e.num_cells_down, num_cells_up=e.num_cells_up))
        dumps = [{k: getattr(c, k) for k in ('x', 'upper_index', 'lower_index', 'upper_orient', 'lower_orient', 'y')}
                 for c in cochains]
        model.train()
        sd0 = {k: v.clone() for k, v in model.state_dict().items()}
        batch = RefCochainBatch.from_cochain_list(cochains)
        out, cell_pred = model.forward(batch, include_partial=True)
        loss = l1(out, batch.y)
        models[name] = {'cfg': cfg, 'state_dict': sd0, 'inputs': dumps, 'output': out.detach().clone(),
                        'cell_pred': cell_pred.detach().clone(), 'loss': loss.detach().clone(),
                        'grads': grads_of(model, loss)}

    # every edge of these complexes lies on a ring or next to one; edges without upper adjacency keep upper_index None
    cfg = dict(num_input_features=3, num_classes=1, num_layers=2, hidden=6, nonlinearity='tanh')
    run_oriented('edge_orient_train', RefEdgeOrient(**cfg), cfg,
                 synthetic.float_feature_complexes(4, 3, seed=11, include_down_adj=True), 21)
    cfg = dict(num_input_features=3, num_classes=1, num_layers=2, hidden=6)
    run_oriented('edge_mpnn_train', RefEdgeMPNN(**cfg), cfg,
                 synthetic.float_feature_complexes(4, 3, seed=12, include_down_adj=True), 22)

    # OGBEmbedCINpp (mp/molec_models.py:355-385), appended after everything else (random stream above unchanged)
    from mp.molec_models import OGBEmbedCINpp as RefOGBCINpp
    cfg = dict(out_size=1, num_layers=2, hidden=16, dropout_rate=0.0, indropout_rate=0.0, max_dim=2,
               embed_edge=True, use_coboundaries=True, readout='mean')
    run_train('ogb_embed_cinpp_train', RefOGBCINpp(**cfg), cfg,
              synthetic.zinc_like_complexes(6, seed=13, ogb_features=True), bce)

    gold['models'] = models
    path = os.path.join(HERE, 'reference_golden.pt')
    torch.save(gold, path)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
