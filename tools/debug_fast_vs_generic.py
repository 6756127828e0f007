"""Max |fast - generic| of every output / gradient of one SparseCINConv layer (debug aid)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cwn_b200 import _lib  # noqa: E402
from cwn_b200.data import synthetic  # noqa: E402
from cwn_b200.data.complex import ComplexBatch  # noqa: E402
from cwn_b200.mp.layers import SparseCINConv  # noqa: E402
from cwn_b200.mp.nn import get_graph_norm, get_nonlinearity  # noqa: E402

DEV = torch.device('cuda', 0)


def main(layer_dim=32, hidden=128, n_complexes=300):
    torch.manual_seed(2)
    conv = SparseCINConv(layer_dim, layer_dim, layer_dim, None, None, None, None, layer_dim=layer_dim, hidden=hidden,
                         act_module=get_nonlinearity('relu'), graph_norm=get_graph_norm('bn'), use_coboundaries=True,
                         train_eps=True).to(DEV).train()
    state = {k: v.clone() for k, v in conv.state_dict().items()}
    results = []
    for generic in (0, 1):
        _lib.check(_lib.load().cwn_debug_force_generic_dense(generic))
        conv.load_state_dict(state)
        conv.zero_grad(set_to_none=True)
        batch = ComplexBatch.from_complex_list(
            synthetic.float_feature_complexes(n_complexes, layer_dim, seed=5, ragged=True)).to(DEV)
        for d in range(3):
            batch.cochains[d]._x = batch.cochains[d].x.clone().requires_grad_(True)
        outs = conv(*batch.get_all_cochain_params(max_dim=2, include_down_features=False))
        g = torch.Generator(device=DEV).manual_seed(3)
        sum((o * torch.randn(o.shape, device=DEV, generator=g)).sum() for o in outs).backward()
        r = {f'out{d}': outs[d].detach().clone() for d in range(3)}
        r.update({f'gx{d}': batch.cochains[d].x.grad.clone() for d in range(3)})
        r.update({k: p.grad.clone() for k, p in conv.named_parameters() if p.grad is not None})
        results.append(r)
    a, b = results
    print('rows per dim:', [a[f'out{d}'].shape[0] for d in range(3)])
    for k in a:
        err = (a[k] - b[k]).abs().max().item()
        print(f'{k:50s} max|fast-generic| {err:.3e}   max|generic| {b[k].abs().max().item():.3e}')


if __name__ == '__main__':
    main(*[int(v) for v in sys.argv[1:]])
