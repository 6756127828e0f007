"""Per-parameter gradient error of the ZINC-shaped training step against the CPU oracle (debug aid).
usage: python tools/dbg_grad_err.py [n_complexes] [seed ...]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle')); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import cwn_oracle as O
from helpers import oracle_state
from cwn_b200.data import synthetic
from cwn_b200.data.complex import ComplexBatch
from cwn_b200.mp.molec_models import EmbedSparseCIN

DEV = torch.device('cuda', 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
seeds = [int(s) for s in sys.argv[2:]] or [0]
cfg = dict(atom_types=28, bond_types=4, out_size=1, num_layers=4, hidden=64, dropout_rate=0.0, max_dim=2,
           embed_edge=True, use_coboundaries=True, nonlinearity=os.environ.get('ACT', 'relu'))
for seed in seeds:
    torch.manual_seed(seed)
    model = EmbedSparseCIN(**cfg)
    sd = oracle_state(model.state_dict(), requires_grad=True)
    snap = O.Snapshot(ComplexBatch.from_complex_list(synthetic.zinc_like_complexes(n, seed=seed)))
    ref = O.embed_sparse_cin(sd, cfg, snap, training=True)
    torch.nn.functional.l1_loss(ref, snap.y.view(-1, 1)).backward()
    model.to(DEV).train()
    batch = ComplexBatch.from_complex_list(synthetic.zinc_like_complexes(n, seed=seed)).to(DEV)
    out = model(batch)
    torch.nn.functional.l1_loss(out, batch.y.view(-1, 1)).backward()
    worst = []
    for k, p in model.named_parameters():
        if sd[k].grad is None:
            continue
        e = (p.grad.cpu().double() - sd[k].grad.double()).abs()
        bound = 1e-5 + 1e-4 * sd[k].grad.double().abs()
        worst.append((float((e - bound).max()), float(e.max()), float(sd[k].grad.abs().max()), k))
    worst.sort(reverse=True)
    oe = float((out.cpu() - ref).abs().max())
    print(f'act {cfg["nonlinearity"]} seed {seed} n {n} TC5={os.environ.get("CWN_B200_DENSE_TC5", "1")}: out err {oe:.2e}; params over bound: '
          f'{sum(w[0] > 0 for w in worst)} of {len(worst)}')
    for w in worst[:6]:
        print(f'    excess {w[0]:.2e} err {w[1]:.2e} max|grad| {w[2]:.2e}  {w[3]}')
