"""Group an ncu launch list (`--metrics gpu__time_duration.sum --csv`) by kernel for ONE training step.

A step is delimited by consecutive `cwn::adam_kernel` launches (Adam is the last kernel of a step).
usage: python tools/summarize_launches.py gpurun_out/launches.csv [step_index] > profiles/...summary.txt
"""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r'^void ', '', name)
    name = re.sub(r'at::native::(\(anonymous namespace\)::)?', 'at::', name)
    return re.sub(r'<.*', '', name)[:90]


def main(path, step=-1):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    head, rows = rows[0], rows[1:]
    k, v = head.index('Kernel Name'), head.index('Metric Value')
    launches = [(short(r[k]), float(r[v].replace(',', '')) / 1e3) for r in rows]
    ends = [i for i, (n, _) in enumerate(launches) if n.startswith('cwn::adam_kernel')]
    assert len(ends) >= 2, 'need two Adam launches to delimit a step'
    lo, hi = ends[step - 1] + 1, ends[step] + 1
    agg = defaultdict(lambda: [0.0, 0])
    for n, us in launches[lo:hi]:
        agg[n][0] += us
        agg[n][1] += 1
    total = sum(a[0] for a in agg.values())
    print(f'# ONE training step = launches {lo}..{hi - 1} of the capture; cold-cache serialised times: compare SHARES')
    print(f'# total {total:.1f} us over {hi - lo} launches')
    for n, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f'{us:10.1f} us  {100 * us / total:5.1f}%  {c:5d} launches  {n}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else -1)
