"""TEST INFRASTRUCTURE ONLY — torch restatements of the six `cwn_b200.ops` entry points the Python host layer calls, so
that the HOST LOGIC (hook protocol, layer / model wiring, data API) can be exercised in the GPU-less build container.
The product has no CPU path: `cwn_b200.ops` raises on non-CUDA tensors; this module is imported by tests only and is
installed with `monkeypatch` for the duration of one test. Semantics follow `oracle/cwn_oracle.py::scatter`."""
import torch

import cwn_oracle as O


def gather_scatter(x_src, index, n_dst, reduce='add', x_res=None, eps=None):
    out = O.scatter(x_src.index_select(0, index[0]), index[1], n_dst, 'add' if reduce == 'sum' else reduce)
    if x_res is not None:
        out = out + (1 + (eps if eps is not None else 0)) * x_res
    return out


def gather_rows(x, idx, scale=1.0):
    out = x.index_select(0, idx)
    return out if scale == 1.0 else scale * out


def scatter_rows(msg, dst, n_dst, reduce='add'):
    if msg.dim() == 1:
        msg = msg.unsqueeze(-1)
    return O.scatter(msg, dst, int(n_dst), 'add' if reduce == 'sum' else reduce)


def cob_pass(P, Q, index, cob, n_dst, act='relu', x_res=None, eps=None):
    out = O.scatter(O._ACT[act](P.index_select(0, index[0]) + Q.index_select(0, cob)), index[1], n_dst, 'add')
    if x_res is not None:
        out = out + (1 + (eps if eps is not None else 0)) * x_res
    return out


def segment_pool(x, batch, size, mean=False):
    return O.scatter(x, batch, int(size), 'mean' if mean else 'add')


def prepare_plans(*args, **kwargs):
    return 0


def install(monkeypatch):
    from cwn_b200 import ops
    for name in ('gather_scatter', 'gather_rows', 'scatter_rows', 'cob_pass', 'segment_pool', 'prepare_plans'):
        monkeypatch.setattr(ops, name, globals()[name])
