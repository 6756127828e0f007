"""DRAM bytes per launch of the step's dense kernels from an `ncu --set full` report, in the form bench.py reads for
`roofline.traffic`:  python tools/ncu_dram_traffic.py REPORT.ncu-rep OUT.json"""
import csv
import json
import subprocess
import sys

NAMES = {'unit_bwd_tc5_kernel': 'unit_bwd_grouped', 'linear_fwd_tc5_kernel': 'linear_fwd_grouped',
         'unit_bwd_reduce_fast_kernel': 'unit_bwd_reduce_grouped', 'unit_bwd_fast_kernel': 'unit_bwd_grouped',
         'linear_fwd_fast_kernel': 'linear_fwd_grouped'}
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def main(path, out):
    txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    acc = {}
    for r in rows[2:]:
        name = r[col['Kernel Name']]
        key = next((v for k, v in NAMES.items() if k in name), None)
        if key is None:
            continue
        total = 0.0
        for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            total += float(r[col[m]].replace(',', '')) * UNIT[units[col[m]]]
        us = float(r[col['gpu__time_duration.sum']].replace(',', ''))
        if units[col['gpu__time_duration.sum']] in ('nsecond', 'ns'):
            us /= 1e3
        a = acc.setdefault(key, [0.0, 0, 0.0])
        a[0] += total
        a[1] += 1
        a[2] += us
    res = {k: {'dram_bytes_per_launch': v[0] / v[1], 'launches_captured': v[1], 'avg_us_under_ncu': v[2] / v[1],
               'source': f'{path} (ncu --set full, cold-cache replays: DRAM reads include operands a warm step finds in '
                         f'L2; writes stay in the 126 MB L2 and show as 0)'} for k, v in acc.items()}
    json.dump(res, open(out, 'w'), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
