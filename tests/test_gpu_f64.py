"""float64 path (reference exp/run_exp.py:41-43 sets the default dtype to double for the SR
isomorphism experiments): the `*_f64` CSR entry points through the public operators, against float64 torch on the CPU,
and a whole model in double against the CPU oracle in double. Tolerances are those of double arithmetic with a different
summation order (1e-12); additive aggregation is bit-exact against a sequential CPU index_add_."""
import pytest
import torch

import cwn_oracle as O
from cwn_b200 import ops
from cwn_b200.data import synthetic
from cwn_b200.data.complex import ComplexBatch
from cwn_b200.mp.models import SparseCIN
from helpers import oracle_state

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _problem(F, seed=0, n_src=700, n_dst=900, E=4000):
    g = torch.Generator().manual_seed(seed)
    idx = torch.stack([torch.randint(0, n_src, (E,), generator=g), torch.randint(0, n_dst - 20, (E,), generator=g)])
    x = torch.randn(n_src, F, generator=g, dtype=torch.float64)
    return idx, x, n_src, n_dst, g


@pytest.mark.parametrize('F', [1, 5, 64])
def test_f64_additive_aggregation_is_bit_exact_with_sequential_cpu(F):
    idx, x, n_src, n_dst, g = _problem(F)
    ref = torch.zeros(n_dst, F, dtype=torch.float64).index_add_(0, idx[1], x[idx[0]])
    out = ops.gather_scatter(x.to(DEV), idx.to(DEV), n_dst)
    assert out.dtype == torch.float64 and torch.equal(out.cpu(), ref)
    res = torch.randn(n_dst, F, generator=g, dtype=torch.float64)
    eps = torch.tensor([0.3], dtype=torch.float64)
    out = ops.gather_scatter(x.to(DEV), idx.to(DEV), n_dst, x_res=res.to(DEV), eps=eps.to(DEV))
    assert torch.equal(out.cpu(), ref + (1 + eps) * res)


@pytest.mark.parametrize('reduce', ['add', 'mean', 'max'])
def test_f64_gather_scatter_forward_backward(reduce):
    idx, x, n_src, n_dst, g = _problem(20, seed=3)
    w = torch.randn(n_dst, 20, generator=g, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    ref = O.scatter(xr.index_select(0, idx[0]), idx[1], n_dst, reduce)
    xg = x.to(DEV).requires_grad_(True)
    if reduce == 'max':
        out = ops.gather_scatter(xg.detach(), idx.to(DEV), n_dst, reduce)
        assert torch.allclose(out.cpu(), ref.detach(), rtol=0, atol=0)
        with pytest.raises(NotImplementedError):
            ops.gather_scatter(xg, idx.to(DEV), n_dst, reduce)
        return
    (ref * w).sum().backward()
    out = ops.gather_scatter(xg, idx.to(DEV), n_dst, reduce)
    (out * w.to(DEV)).sum().backward()
    assert torch.allclose(out.detach().cpu(), ref.detach(), rtol=1e-13, atol=1e-13)
    assert torch.allclose(xg.grad.cpu(), xr.grad, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize('act', ['relu', 'elu', 'tanh', 'sigmoid', 'id'])
def test_f64_coboundary_pass_forward_backward(act):
    F, n, n_cob, E = 12, 500, 130, 2600
    g = torch.Generator().manual_seed(7)
    idx = torch.stack([torch.randint(0, n, (E,), generator=g), torch.randint(0, n - 9, (E,), generator=g)])
    cob = torch.randint(0, n_cob, (E,), generator=g)
    P0, Q0 = torch.randn(n, F, generator=g, dtype=torch.float64), torch.randn(n_cob, F, generator=g, dtype=torch.float64)
    res0, w = torch.randn(n, F, generator=g, dtype=torch.float64), torch.randn(n, F, generator=g, dtype=torch.float64)
    eps = torch.tensor([0.25], dtype=torch.float64)
    fn = O._ACT[act]
    P, Q, res = (t.clone().requires_grad_(True) for t in (P0, Q0, res0))
    ref = O.scatter(fn(P.index_select(0, idx[0]) + Q.index_select(0, cob)), idx[1], n) + (1 + eps) * res
    (ref * w).sum().backward()
    Pg, Qg, rg = (t.to(DEV).requires_grad_(True) for t in (P0, Q0, res0))
    out = ops.cob_pass(Pg, Qg, idx.to(DEV), cob.to(DEV), n, act=act, x_res=rg, eps=eps.to(DEV))
    (out * w.to(DEV)).sum().backward()
    assert out.dtype == torch.float64
    assert torch.allclose(out.detach().cpu(), ref.detach(), rtol=1e-12, atol=1e-12)
    for got, want in ((Pg, P), (Qg, Q), (rg, res)):
        assert torch.allclose(got.grad.cpu(), want.grad, rtol=1e-11, atol=1e-11)


def test_f64_row_gather_and_message_scatter():
    idx, x, n_src, n_dst, g = _problem(9, seed=5)
    xr = x.clone().requires_grad_(True)
    w = torch.randn(idx.size(1), 9, generator=g, dtype=torch.float64)
    (2.5 * xr.index_select(0, idx[0]) * w).sum().backward()
    xg = x.to(DEV).requires_grad_(True)
    out = ops.gather_rows(xg, idx[0].to(DEV), 2.5)
    (out * w.to(DEV)).sum().backward()
    assert torch.equal(out.detach().cpu(), 2.5 * x.index_select(0, idx[0]))
    assert torch.allclose(xg.grad.cpu(), xr.grad, rtol=1e-12, atol=1e-12)
    msg = torch.randn(idx.size(1), 9, generator=g, dtype=torch.float64)
    for reduce in ('add', 'mean'):
        mr = msg.clone().requires_grad_(True)
        ref = O.scatter(mr, idx[1], n_dst, reduce)
        ref.pow(2).sum().backward()
        mg = msg.to(DEV).requires_grad_(True)
        got = ops.scatter_rows(mg, idx[1].to(DEV), n_dst, reduce)
        got.pow(2).sum().backward()
        assert torch.allclose(got.detach().cpu(), ref.detach(), rtol=1e-13, atol=1e-13)
        assert torch.allclose(mg.grad.cpu(), mr.grad, rtol=1e-12, atol=1e-12)
    with pytest.raises(TypeError):
        ops.gather_scatter(x.to(DEV), idx.to(DEV), n_dst, x_res=torch.zeros(n_dst, 9, device=DEV))  # mixed dtypes


def _double_model_step(dev):
    """What the reference does for SR: torch.set_default_dtype(torch.float64), then build and run (exp/run_exp.py:41-43)."""
    cfg = dict(num_input_features=3, num_classes=2, num_layers=2, hidden=16, dropout_rate=0.0, max_dim=2,
               nonlinearity='elu', readout='sum', use_coboundaries=True)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        torch.manual_seed(1)
        model = SparseCIN(**cfg)
        sd = oracle_state(model.state_dict(), requires_grad=True)

        def mk():
            b = ComplexBatch.from_complex_list(synthetic.float_feature_complexes(6, 3, seed=4, ragged=True))
            for c in b.cochains.values():
                c.x = c.x.double()
            return b
        ref = O.sparse_cin(sd, cfg, O.Snapshot(mk()), training=True)
        ref.pow(2).sum().backward()
        out = None
        if dev is not None:
            model.to(dev).train()
            out = model(mk().to(dev))
            out.pow(2).sum().backward()
        return model, sd, ref, out
    finally:
        torch.set_default_dtype(prev)


def test_f64_model_training_step_against_the_oracle_in_double():
    """SparseCIN with the default dtype set to double, as the reference does for the SR experiments: message passing on
    the f64 kernels, dense nets on torch (the fused dense kernels are float32-only and step aside), against the CPU
    oracle in double at 1e-10."""
    model, sd, ref, out = _double_model_step(DEV)
    assert out.dtype == torch.float64 and ref.dtype == torch.float64
    assert torch.allclose(out.detach().cpu(), ref.detach(), rtol=1e-10, atol=1e-10)
    checked = 0
    for k, p in model.named_parameters():
        if sd[k].grad is None:
            continue
        assert torch.allclose(p.grad.cpu(), sd[k].grad, rtol=1e-8, atol=1e-9), k
        checked += 1
    assert checked > 10
