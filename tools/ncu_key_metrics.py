"""Key metrics of every kernel in an ncu report (`--set full`): python tools/ncu_key_metrics.py REPORT.ncu-rep"""
import csv
import subprocess
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
STALLS = 'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio'


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        for w in WANT:
            if w in d:
                print(f'{w:80s} {d[w][:110]:>24s} {units[hdr.index(w)]}')
        stalls = sorted(((float(v.replace(',', '')), k.split('stalled_')[1].split('_per_issue')[0]) for k, v in d.items()
                         if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio') and v),
                        reverse=True)
        print('stalls (warps per issue):', ', '.join(f'{n} {v:.2f}' for v, n in stalls[:8]))
        print()


if __name__ == '__main__':
    main(sys.argv[1])
