"""Stall samples of one kernel of an ncu report, aggregated per CUDA source line.

    python tools/ncu_by_line.py REPORT.ncu-rep KERNEL_REGEX path/to/lib.so [top_n]

ncu's CLI prints per-SASS-instruction samples (`--page source --csv`); nvdisasm -g gives the source line of every SASS
instruction of the same binary. The two are joined by instruction index."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, regex, lib = sys.argv[1:4]
    top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{regex}'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'rows': []}
            blocks.append(cur)
        elif r and r[0] == 'Address':
            cur['hdr'] = r
        elif cur is not None and r:
            cur['rows'].append(r)
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
    sass = ''
    for f in os.listdir(tmp):
        if f.endswith('.cubin'):
            sass += subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, f)], capture_output=True, text=True).stdout
    src_lines = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'cwn_b200', 'csrc', 'dense.cu')).read().splitlines()
    seen = set()
    for b in blocks:
        if not re.search(regex, b['name']) or (b['name'], len(b['rows'])) in seen:
            continue
        seen.add((b['name'], len(b['rows'])))
        h = {k: i for i, k in enumerate(b['hdr'])}
        samples = [int(r[h['# Samples']] or 0) for r in b['rows']]
        first = b['rows'][0][h['Source']].strip()
        # locate the function in the nvdisasm output whose instruction count matches
        best = None
        funcs, name, lines, cur_line = {}, None, None, None
        for ln in sass.splitlines():
            m = re.match(r'^(_ZN3cwn\w+):\s*$', ln)
            if m:
                name, lines, cur_line = m.group(1), [], None
                funcs[name] = lines
                continue
            if name is None:
                continue
            mm = re.search(r'//## File ".*?", line (\d+)', ln)
            if mm:
                cur_line = int(mm.group(1))
            elif re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln):
                lines.append(cur_line)
        for fname, flines in funcs.items():
            if len(flines) == len(samples):
                best = (fname, flines)
                break
        if best is None:
            print('no matching function (is the library the one that was profiled?)', len(samples))
            continue
        per_line = collections.Counter()
        for s, ln in zip(samples, best[1]):
            per_line[ln] += s
        total = sum(samples)
        print(f'{b["name"][:80]}  ->  {best[0]}\n{len(samples)} instructions, {total} samples')
        for ln, s in per_line.most_common(top_n):
            text = src_lines[ln - 1].strip() if ln and ln <= len(src_lines) else ''
            print(f'{s:6d} {100 * s / max(total, 1):5.1f}%  line {ln}: {text[:110]}')


if __name__ == '__main__':
    main()
