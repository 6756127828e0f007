"""Kernel timeline of ONE replayed training step (CUPTI via torch.profiler): per-kernel warm durations inside the CUDA
graph, how much of the step some kernel is running, and how much is idle (dependency / launch latency).

    python tools/trace_step.py > gpurun_out/trace_step.txt
"""
import collections
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cwn_b200.dist import FlatGradBucket  # noqa: E402
from cwn_b200.mp.molec_models import EmbedSparseCIN  # noqa: E402
from cwn_b200.optim import FlatAdam  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    torch.manual_seed(0)
    model = EmbedSparseCIN(**bench.MODEL_CFG).to(dev).train()
    bucket = FlatGradBucket(model)
    opt = FlatAdam(model, bucket, lr=1e-3)
    # CUPTI reports a replayed CUDA graph as one opaque activity, so the step is traced in EAGER mode: the kernel
    # durations are the warm-cache ones of the real step, the gaps between them are host dispatch (absent in the graph)
    from cwn_b200 import ops
    batches = [b.to(dev) for b in bench.make_batches(4, 128, 1000)]
    inputs = [[b.cochains[d].x for d in range(3)] for b in batches]

    def step(i):
        b = batches[i % 4]
        ops.clear_plan_cache(*[t for d in range(3) for t in (b.cochains[d].upper_index, b.cochains[d].boundary_index,
                                                             b.cochains[d].batch)])
        for d, x in enumerate(inputs[i % 4]):
            b.cochains[d]._x = x
        loss = bench.l1(model(b), b.y)
        loss.backward()
        opt.step()

    for i in range(5):
        step(i)
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        step(0)
        torch.cuda.synchronize()
    step = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    step.sort(key=lambda e: e.time_range.start)
    t0, t1 = step[0].time_range.start, max(e.time_range.end for e in step)
    print(f'one eager step: {len(step)} device activities, span {(t1 - t0):.1f} us')
    if os.environ.get('CWN_TRACE_TIMELINE'):
        for e in step:
            print(f'  {e.time_range.start - t0:9.1f} +{e.time_range.end - e.time_range.start:7.2f}  '
                  + re.sub(r'<.*', '', e.name).replace('void ', '')[:60])
    tot, cnt = collections.Counter(), collections.Counter()
    for e in step:
        name = re.sub(r'<.*', '', e.name).replace('void ', '')[:70]
        tot[name] += e.time_range.end - e.time_range.start
        cnt[name] += 1
    busy = 0.0
    cur_end = t0
    for e in step:  # union of intervals = time at least one kernel is running
        s, en = e.time_range.start, e.time_range.end
        if en <= cur_end:
            continue
        busy += en - max(s, cur_end)
        cur_end = en
    print(f'some kernel running: {busy:.1f} us ({100 * busy / (t1 - t0):.1f} %), idle: {(t1 - t0) - busy:.1f} us; '
          f'sum of kernel durations {sum(tot.values()):.1f} us')
    for k, v in tot.most_common(40):
        print(f'{v:9.1f} us {cnt[k]:4d} x {v / cnt[k]:6.2f}  {k}')


if __name__ == '__main__':
    main()
