"""Shared test utilities: golden loading, conversion of golden dumps to cwn_b200 data objects, oracle plumbing."""
import os

import torch

from cwn_b200.data.complex import Cochain, Complex, ComplexBatch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_golden.pt')
_cache = {}
COCHAIN_KEYS = ['x', 'upper_index', 'lower_index', 'boundary_index', 'shared_boundaries', 'shared_coboundaries', 'y']


def golden():
    if 'g' not in _cache:
        _cache['g'] = torch.load(GOLDEN, weights_only=False)
    return _cache['g']


def complex_from_dump(d) -> Complex:
    """Golden dump of a reference Complex -> cwn_b200 Complex with the reference's exact tensors."""
    cochains = []
    for dim in range(d['dimension'] + 1):
        c = d['cochains'][dim]
        kw = {k: (None if c[k] is None else c[k].clone()) for k in COCHAIN_KEYS}
        cochains.append(Cochain(dim=dim, num_cells=c['num_cells'], num_cells_up=c['num_cells_up'],
                                num_cells_down=c['num_cells_down'], **kw))
    return Complex(*cochains, y=None if d['y'] is None else d['y'].clone())


def fixture(name) -> Complex:
    return complex_from_dump(golden()['fixtures'][name])


def batch_of(names_or_dumps, max_dim=2) -> ComplexBatch:
    comps = [fixture(n) if isinstance(n, str) else complex_from_dump(n) for n in names_or_dumps]
    return ComplexBatch.from_complex_list(comps, max_dim=max_dim)


def oracle_state(sd, device='cpu', requires_grad=False):
    """Clone a state_dict for the oracle; entries that alias one tensor (shared modules of CIN0) keep aliasing."""
    seen, out = {}, {}
    for k, v in sd.items():
        key = (v.data_ptr(), tuple(v.shape), v.dtype) if v.numel() else (id(v),)
        if key not in seen:
            t = v.detach().clone().to(device)
            if requires_grad and t.is_floating_point() and not any(
                    s in k for s in ('running_mean', 'running_var', 'num_batches_tracked')):
                t.requires_grad_(True)
            seen[key] = t
        out[k] = seen[key]
    return out


def assert_close(a, b, rtol=1e-5, atol=1e-5, what=''):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    err = (a - b).abs()
    bound = atol + rtol * b.abs()
    assert bool((err <= bound).all()), f'{what}: max abs err {err.max().item():.3e}, ' \
                                       f'max excess {(err - bound).max().item():.3e}'


def share_cin0_levels(sd):
    """Re-alias the per-dimension entries of a (cloned) CIN0 state_dict: all `mp_levels.{d}` share one set of nets."""
    import re
    for k in list(sd):
        k0 = re.sub(r'mp_levels\.\d+\.', 'mp_levels.0.', k)
        if k0 != k and not k.endswith('.eps'):
            sd[k] = sd[k0]
    return sd
