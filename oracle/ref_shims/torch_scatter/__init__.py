"""Stand-in for torch_scatter==2.0.5 restating its published semantics (scatter.py of that release):
  scatter(src, index, dim, out=None, dim_size=None, reduce) with
    'sum'/'add': index broadcast to src, out = zeros(dim_size along dim), out.scatter_add_(dim, index, src)
    'mean'     : sum, then divided by clamp(count, min=1)
    'max'/'min': per-destination extremum, destinations without any message are 0
Call sites in the reference: mp/cell_mp.py:437-440,456-459,476-479; mp/layers.py:487."""
import torch


def _broadcast(index, src, dim):
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(0, dim):
            index = index.unsqueeze(0)
    for _ in range(index.dim(), src.dim()):
        index = index.unsqueeze(-1)
    return index.expand_as(src)


def _out_size(src, index, dim, dim_size):
    size = list(src.size())
    if dim_size is not None:
        size[dim] = int(dim_size)
    elif index.numel() == 0:
        size[dim] = 0
    else:
        size[dim] = int(index.max()) + 1
    return size


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    index = _broadcast(index, src, dim)
    if out is None:
        out = torch.zeros(_out_size(src, index, dim, dim_size), dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, index, src)


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    out = scatter_sum(src, index, dim, out, dim_size)
    dim_size = out.size(dim)
    index_dim = dim
    if index_dim < 0:
        index_dim = index_dim + src.dim()
    if index.dim() <= index_dim:
        index_dim = index.dim() - 1
    ones = torch.ones(index.size(), dtype=src.dtype, device=src.device)
    count = scatter_sum(ones, index, index_dim, None, dim_size)
    count.clamp_(1)
    count = _broadcast(count, out, dim)
    if torch.is_floating_point(out):
        out.true_divide_(count)
    else:
        out.floor_divide_(count)
    return out


def _scatter_extremum(src, index, dim, dim_size, kind):
    index = _broadcast(index, src, dim)
    out = torch.zeros(_out_size(src, index, dim, dim_size), dtype=src.dtype, device=src.device)
    return out.scatter_reduce_(dim, index, src, reduce=kind, include_self=False)


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    if reduce in ("sum", "add"):
        return scatter_sum(src, index, dim, out, dim_size)
    if reduce == "mean":
        return scatter_mean(src, index, dim, out, dim_size)
    if reduce == "max":
        return _scatter_extremum(src, index, dim, dim_size, "amax")
    if reduce == "min":
        return _scatter_extremum(src, index, dim, dim_size, "amin")
    raise ValueError(reduce)


def segment_csr(*args, **kwargs):
    raise NotImplementedError("stand-in: segment_csr is dead code in the reference (SparseTensor path)")


def gather_csr(*args, **kwargs):
    raise NotImplementedError("stand-in: gather_csr is dead code in the reference (SparseTensor path)")
