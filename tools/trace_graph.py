"""Kernel timeline of ONE replay of the captured training step (CUPTI via torch.profiler): which kernels run
concurrently, where the device idles, and the critical path as the graph actually executes it.

    python tools/trace_graph.py > gpurun_out/trace_graph.txt      (CWN_TRACE_TIMELINE=1 adds the full list)
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
from cwn_b200.graph import CapturedStep  # noqa: E402
from cwn_b200.mp.molec_models import EmbedSparseCIN  # noqa: E402
from cwn_b200.optim import FlatAdam  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    torch.manual_seed(0)
    model = EmbedSparseCIN(**bench.MODEL_CFG).to(dev).train()
    bucket = FlatGradBucket(model)
    opt = FlatAdam(model, bucket, lr=1e-3)
    batches = [b.pack_().to(dev) for b in bench.make_batches(3, 128, 1000)]
    cap = CapturedStep(model, bench.l1, bucket, opt, optimizer_in_graph=True).capture(batches[0])
    for i in range(5):
        cap.run(batches[1 + i % 2])
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        cap.run(batches[1])
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
    print(f'one replay: {len(evs)} device activities, span {(t1 - t0):.1f} us')
    tot, cnt = collections.Counter(), collections.Counter()
    busy, cur_end = 0.0, t0
    for e in evs:
        name = re.sub(r'<.*', '', e.name).replace('void ', '')[:60]
        tot[name] += e.time_range.end - e.time_range.start
        cnt[name] += 1
        s, en = e.time_range.start, e.time_range.end
        if en > cur_end:
            busy += en - max(s, cur_end)
            cur_end = en
    print(f'some kernel running: {busy:.1f} us ({100 * busy / (t1 - t0):.1f} %), idle {(t1 - t0) - busy:.1f} us; '
          f'sum of kernel durations {sum(tot.values()):.1f} us')
    for k, v in tot.most_common(30):
        print(f'{v:9.1f} us {cnt[k]:4d} x {v / cnt[k]:6.2f}  {k}')
    if os.environ.get('CWN_TRACE_TIMELINE'):
        for e in evs:
            print(f'  {e.time_range.start - t0:9.1f} +{e.time_range.end - e.time_range.start:7.2f}  '
                  + re.sub(r'<.*', '', e.name).replace('void ', '')[:60])


if __name__ == '__main__':
    main()
