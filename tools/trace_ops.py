"""Which aten operators still launch torch kernels inside ONE eager training step (forward, loss, backward, Adam), with the
Python source line that issued them: the to-do list for launch-count reduction.

    python tools/trace_ops.py > gpurun_out/trace_ops.txt
"""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cwn_b200 import ops  # noqa: E402
from cwn_b200.dist import FlatGradBucket  # noqa: E402
from cwn_b200.mp.molec_models import EmbedSparseCIN  # noqa: E402
from cwn_b200.optim import FlatAdam  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    torch.manual_seed(0)
    model = EmbedSparseCIN(**bench.MODEL_CFG).to(dev).train()
    bucket = FlatGradBucket(model)
    opt = FlatAdam(model, bucket, lr=1e-3)
    batches = [b.to(dev) for b in bench.make_batches(2, 128, 1000)]
    inputs = [[b.cochains[d].x for d in range(3)] for b in batches]

    def step(i):
        b = batches[i % 2]
        ops.clear_plan_cache(*[t for d in range(3) for t in (b.cochains[d].upper_index, b.cochains[d].boundary_index,
                                                             b.cochains[d].batch)])
        for d, x in enumerate(inputs[i % 2]):
            b.cochains[d]._x = x
        loss = bench.l1(model(b), b.y)
        loss.backward()
        opt.step()

    for i in range(4):
        step(i)
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA],
                                with_stack=True) as prof:
        step(0)
        torch.cuda.synchronize()
    # aten ops that launched at least one kernel themselves (self device time > 0)
    rows = collections.defaultdict(lambda: [0, 0.0])
    for e in prof.events():
        if e.device_type != torch.autograd.DeviceType.CPU or not e.name.startswith('aten::'):
            continue
        kernels = [k for k in e.kernels]
        if not kernels:
            continue
        # attribute to the innermost op only: skip if a child aten op owns the same kernels
        if any(c.name.startswith('aten::') and c.kernels for c in e.cpu_children):
            continue
        where = next((s for s in (e.stack or []) if 'cwn_b200' in s or 'bench.py' in s), '(autograd engine)')
        key = (e.name, where.split('/root/repo/')[-1].split(os.sep + 'repo' + os.sep)[-1][:90])
        rows[key][0] += len(kernels)
        rows[key][1] += sum(k.duration for k in kernels)
    total = sum(v[0] for v in rows.values())
    print(f'{total} torch kernel launches in one eager step, by aten op and issuing line')
    for (name, where), (n, us) in sorted(rows.items(), key=lambda kv: -kv[1][0]):
        print(f'{n:4d} launches {us:8.1f} us  {name:32s} {where}')


if __name__ == '__main__':
    main()
