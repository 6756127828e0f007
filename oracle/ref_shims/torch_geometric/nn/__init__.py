"""JumpingKnowledge ('cat' = concat over layers on the last dim, 'max' = elementwise max over layers),
global_add_pool / global_mean_pool (= scatter over the `batch` vector with dim_size=size)."""
import torch
from torch_scatter import scatter


class JumpingKnowledge(torch.nn.Module):
    def __init__(self, mode, channels=None, num_layers=None):
        super().__init__()
        self.mode = mode.lower()
        assert self.mode in ("cat", "max")

    def reset_parameters(self):
        pass

    def forward(self, xs):
        assert isinstance(xs, (list, tuple))
        if self.mode == "cat":
            return torch.cat(xs, dim=-1)
        return torch.stack(xs, dim=-1).max(dim=-1)[0]


def global_add_pool(x, batch, size=None):
    size = int(batch.max().item() + 1) if size is None else size
    return scatter(x, batch, dim=0, dim_size=size, reduce="add")


def global_mean_pool(x, batch, size=None):
    size = int(batch.max().item() + 1) if size is None else size
    return scatter(x, batch, dim=0, dim_size=size, reduce="mean")


class _Unavailable(torch.nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("stand-in: not on the hot path")


GINEConv = GINConv = _Unavailable
