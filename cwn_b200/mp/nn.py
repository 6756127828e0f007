"""Activation / pooling / normalisation factories and the per-complex readout (reference `mp/nn.py:7-60`)."""
import torch
import torch.nn.functional as F
from torch.nn import BatchNorm1d as BN, LayerNorm as LN, Identity

from cwn_b200 import ops

_ACTIVATIONS = {
    'relu': (torch.nn.ReLU, F.relu),
    'elu': (torch.nn.ELU, F.elu),
    'id': (torch.nn.Identity, lambda x: x),
    'sigmoid': (torch.nn.Sigmoid, torch.sigmoid),
    'tanh': (torch.nn.Tanh, torch.tanh),
}


def get_nonlinearity(nonlinearity, return_module=True):
    if nonlinearity not in _ACTIVATIONS:
        raise NotImplementedError('Nonlinearity {} is not currently supported.'.format(nonlinearity))
    module, function = _ACTIVATIONS[nonlinearity]
    return module if return_module else function


def activation_name(module) -> str:
    """Kernel activation code name of an activation module instance, or None if it is not one of the five."""
    for name, (klass, _) in _ACTIVATIONS.items():
        if type(module) is klass:
            if name == 'elu' and (module.alpha != 1.0):
                return None
            return name
    return None


def global_add_pool(x, batch, size=None):
    """Sum readout per complex; `size` (number of complexes, host int) avoids a device sync."""
    size = int(batch.max()) + 1 if size is None else int(size)
    return ops.segment_pool(x, batch, size, mean=False)


def global_mean_pool(x, batch, size=None):
    size = int(batch.max()) + 1 if size is None else int(size)
    return ops.segment_pool(x, batch, size, mean=True)


def get_pooling_fn(readout):
    if readout == 'sum':
        return global_add_pool
    if readout == 'mean':
        return global_mean_pool
    raise NotImplementedError('Readout {} is not currently supported.'.format(readout))


def get_graph_norm(norm):
    if norm == 'bn':
        return BN
    if norm == 'ln':
        return LN
    if norm == 'id':
        return Identity
    raise ValueError(f'Graph Normalisation {norm} not currently supported')


def num_complexes_of(data) -> int:
    """Batch size known on the host (ComplexBatch.num_complexes); the reference syncs on `batch.max() + 1`."""
    n = getattr(data, 'num_complexes', None)
    if n is None:
        n = int(data.cochains[0].batch.max()) + 1
    return int(n)


def pool_complex(xs, data, max_dim, readout_type):
    """[max_dim+1, num_complexes, F] readout; dimensions absent from the batch stay zero (reference :50-60)."""
    pooling_fn = get_pooling_fn(readout_type)
    batch_size = num_complexes_of(data)
    pooled = [pooling_fn(xs[i], data.cochains[i].batch, size=batch_size) for i in range(len(xs))]
    zero = None
    while len(pooled) < max_dim + 1:
        if zero is None:
            zero = torch.zeros(batch_size, xs[0].size(-1), device=xs[0].device)
        pooled.append(zero)
    return torch.stack(pooled, dim=0)


class JumpingKnowledge(torch.nn.Module):
    """'cat': concatenate the per-layer features; 'max': elementwise maximum over layers."""

    def __init__(self, mode):
        super(JumpingKnowledge, self).__init__()
        self.mode = mode.lower()
        if self.mode not in ('cat', 'max'):
            raise NotImplementedError(f'JumpingKnowledge mode {mode} is not supported')

    def reset_parameters(self):
        pass

    def forward(self, xs):
        if self.mode == 'cat':
            return torch.cat(xs, dim=-1)
        return torch.stack(xs, dim=-1).max(dim=-1)[0]


def reset(nn):
    """Call `reset_parameters` on a module's children (or on the module itself if it has none)."""
    if nn is None:
        return
    children = list(nn.children()) if hasattr(nn, 'children') else []
    for item in (children if children else [nn]):
        if hasattr(item, 'reset_parameters'):
            item.reset_parameters()
