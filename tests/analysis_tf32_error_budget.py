"""Would tensor-core products meet the parity gate? CPU emulation of TF32 and 3xTF32 ("split") GEMMs inside the oracle.

The dense update / combine nets of the hot path run as fp32 FFMA tiles (csrc/dense.cu). The north star allows tensor
cores for exactly that part, and profiles/README.md shows those tiles are issue-bound, so tcgen05 / mma.sync with
kind::tf32 is the obvious next step — IF the 1e-5 rtol gate on cochain features survives it. This tool replaces every
`F.linear` of the torch-only oracle by an emulation of
    tf32   : both operands rounded to 10 mantissa bits (cvt.rna), products accumulated in fp32
    3xtf32 : a = a_hi + a_lo, b = b_hi + b_lo (each part TF32), a_lo*b_hi + a_hi*b_lo + a_hi*b_hi
and measures, on the benchmark configuration (EmbedSparseCIN, 4 layers, hidden 64, 128 ZINC-shaped complexes, BatchNorm
in training mode), the deviation of the model output, of the per-layer cochain features and of the parameter gradients
from a float64 run of the same oracle — next to the deviation of plain fp32, which is what "parity" already tolerates.

    python tests/analysis_tf32_error_budget.py          (CPU, ~1 minute; lives under tests/ because it drives the oracle,
                                                        which is test infrastructure)
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import cwn_oracle as O  # noqa: E402
from cwn_b200.data import synthetic  # noqa: E402
from cwn_b200.data.complex import ComplexBatch  # noqa: E402
from cwn_b200.mp.molec_models import EmbedSparseCIN  # noqa: E402

CFG = dict(atom_types=28, bond_types=4, out_size=1, num_layers=4, hidden=64, dropout_rate=0.0, max_dim=2,
           embed_edge=True, use_coboundaries=True)


def tf32(x):
    """fp32 -> TF32 (10 explicit mantissa bits), round to nearest, ties away from zero (cvt.rna.tf32.f32)."""
    bits = x.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


_REAL_LINEAR = F.linear  # captured before `run` swaps F.linear for the emulation


def make_linear(mode):
    real = _REAL_LINEAR

    def linear(x, w, b=None):
        if mode == 'fp32' or x.dtype != torch.float32:
            return real(x, w, b)
        xh, wh = tf32(x), tf32(w)
        if mode == 'tf32':
            out = real(xh, wh)
        else:
            xl, wl = tf32(x - xh), tf32(w - wh)
            out = real(xl, wh) + real(xh, wl) + real(xh, wh)
        return out if b is None else out + b
    return linear


class _EmulatedLinear(torch.autograd.Function):
    """y = x W^T + b with the chosen product emulation in forward AND in both backward GEMMs."""

    @staticmethod
    def forward(ctx, x, w, b, mode):
        ctx.mode = mode
        ctx.save_for_backward(x, w)
        ctx.has_b = b is not None
        return make_linear(mode)(x, w, b)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        lin = make_linear(ctx.mode)
        gx = lin(g, w.t().contiguous())          # g W
        gw = lin(g.t().contiguous(), x.t().contiguous())  # g^T x
        return gx, gw, (g.sum(0) if ctx.has_b else None), None


def run(mode, dtype, sd0, batch):
    sd = {k: (v.detach().clone().to(dtype) if v.is_floating_point() else v.detach().clone()) for k, v in sd0.items()}
    for k, v in sd.items():
        if v.is_floating_point() and 'running' not in k:
            v.requires_grad_(True)
    snap = O.Snapshot(batch)
    for c in snap.cochains.values():
        if c.x is not None and c.x.is_floating_point():
            c.x = c.x.to(dtype)
    snap.y = snap.y.to(dtype)
    real = _REAL_LINEAR
    torch.set_default_dtype(dtype)  # the oracle creates its zero / count tensors in the default dtype
    if dtype == torch.float32 and mode != 'fp32':
        F.linear = lambda x, w, b=None: _EmulatedLinear.apply(x, w, b, mode)
    try:
        out, res = O.embed_sparse_cin(sd, CFG, snap, training=True, include_partial=True)
        loss = F.l1_loss(out, snap.y.view(-1, 1))
        leaves = {k: v for k, v in sd.items() if v.requires_grad}
        grads = dict(zip(leaves, torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)))
    finally:
        F.linear = real
        torch.set_default_dtype(torch.float32)
    return out.detach().double(), {k: v.detach().double() for k, v in res.items()}, \
        {k: g.detach().double() for k, g in grads.items() if g is not None}


def rel_excess(a, ref, rtol=1e-5, atol=1e-5):
    """max over elements of |a - ref| / (atol + rtol |ref|): <= 1 passes the parity gate."""
    return float(((a - ref).abs() / (atol + rtol * ref.abs())).max())


def main():
    torch.manual_seed(0)
    model = EmbedSparseCIN(**CFG)
    sd0 = model.state_dict()
    batch = ComplexBatch.from_complex_list(synthetic.zinc_like_complexes(128, seed=0))
    ref = run('fp32', torch.float64, sd0, batch)       # ground truth
    base = run('fp32', torch.float32, sd0, batch)      # what parity already tolerates
    G = max(float(g.abs().max()) for g in ref[2].values())

    def report(tag, got, against):
        out, res, g = got
        r_out, r_res, r_g = against
        feat = max(rel_excess(res[k], r_res[k]) for k in r_res if k.startswith('layer'))
        grad = max(float((g[k] - r_g[k]).abs().max()) for k in r_g) / G
        print(f'{tag:22s} {float((out - r_out).abs().max()):12.3e} {rel_excess(out, r_out):8.2f} | {feat:12.2f} | {grad:10.2e}')

    print(f'{"":22s} {"out: max|d|":>12s} {"gate x":>8s} | {"features x":>12s} | {"grads /G":>10s}')
    report('fp32    vs float64', base, ref)
    for mode in ('3xtf32', 'tf32'):
        got = run(mode, torch.float32, sd0, batch)
        report(f'{mode:7s} vs float64', got, ref)
        report(f'{mode:7s} vs fp32', got, base)
    print('gate x / features x = max |d| / (1e-5 + 1e-5 |ref|) over the model output / over every layer\'s cochain '
          'features: <= 1 is inside the parity tolerance; grads /G = worst parameter-gradient deviation over the largest '
          'gradient entry of the model')


if __name__ == '__main__':
    main()
