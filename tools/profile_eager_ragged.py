"""cProfile of the eager (no CUDA graph) training step on ragged batches: where the host time goes."""
import cProfile
import pstats
import sys
import os
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cwn_b200.data import synthetic
from cwn_b200.data.complex import ComplexBatch
from cwn_b200.dist import FlatGradBucket
from cwn_b200.mp import molec_models
from cwn_b200.optim import FlatAdam
import bench

dev = torch.device('cuda', 0)
sys.argv = sys.argv[:1]
args = bench.parse()
model_name, model_cfg, _, gen, loss_fn = bench.workload(args)
model = getattr(molec_models, model_name)(**model_cfg).to(dev).train()
bucket = FlatGradBucket(model)
opt = FlatAdam(model, bucket, lr=1e-3)
pool = synthetic.zinc_like_complexes(8 * 128, seed=7000, **bench.RAGGED_GEN)
batches = [ComplexBatch.from_complex_list(pool[i * 128:(i + 1) * 128]).to(dev) for i in range(8)]
inputs = [[b.cochains[d].x for d in range(3)] for b in batches]


def eager(i):
    b = batches[i % 8]
    for d, x in enumerate(inputs[i % 8]):
        b.cochains[d]._x = x
    loss = bench.l1(model(b), b.y)
    loss.backward()
    opt.step()


for i in range(3):
    eager(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(10):
    eager(i)
torch.cuda.synchronize()
print('eager ms/step', 1e3 * (time.perf_counter() - t0) / 10)
pr = cProfile.Profile()
pr.enable()
for i in range(10):
    eager(i)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(35)
