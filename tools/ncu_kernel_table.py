#!/usr/bin/env python
"""Per-kernel table from an ncu launch list (CSV of `--metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum`): launches, mean duration, mean DRAM bytes, achieved DRAM GB/s against the measured HBM peak.

    python tools/ncu_kernel_table.py gpurun_out/r2_ncu_all_kernels.csv [--skip-first-half] > profiles/r02_ncu_all_kernels.txt
ncu times are cold-cache and serialised: use the SHARE of a kernel, and the bytes, not the absolute duration."""
import csv
import json
import os
import re
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path = sys.argv[1]
    peak = 6537.0
    mp = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(mp):
        peak = float(json.load(open(mp)).get('hbm_gbs', peak))
    rows = [r for r in csv.reader(l for l in open(path, errors='replace') if l.startswith('"'))]
    hdr = rows[0]
    ix = {n: hdr.index(n) for n in ('ID', 'Kernel Name', 'Metric Name', 'Metric Unit', 'Metric Value')}
    per = OrderedDict()
    for r in rows[1:]:
        if len(r) <= ix['Metric Value']:
            continue
        kid = int(r[ix['ID']])
        d = per.setdefault(kid, {'name': r[ix['Kernel Name']]})
        try:
            v = float(r[ix['Metric Value']].replace(',', ''))
        except ValueError:
            continue
        unit = r[ix['Metric Unit']]
        m = r[ix['Metric Name']]
        if m.startswith('gpu__time_duration'):
            v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(unit, 1.0)          # -> us
        else:
            v *= {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1.0)  # -> bytes
        d[m] = v
    ids = sorted(per)
    if '--skip-first-half' in sys.argv:
        ids = ids[len(ids) // 2:]
    agg = OrderedDict()
    for k in ids:
        d = per[k]
        name = re.sub(r'\(.*', '', d['name'])
        name = re.sub(r'^void ', '', name)
        a = agg.setdefault(name, {'n': 0, 'us': 0.0, 'rd': 0.0, 'wr': 0.0})
        a['n'] += 1
        a['us'] += d.get('gpu__time_duration.sum', 0.0)
        a['rd'] += d.get('dram__bytes_read.sum', 0.0)
        a['wr'] += d.get('dram__bytes_write.sum', 0.0)
    total = sum(a['us'] for a in agg.values())
    print(f'# {os.path.basename(path)}: {sum(a["n"] for a in agg.values())} launches, {total / 1e3:.2f} ms under ncu '
          f'(cold-cache, serialised); HBM peak {peak:.0f} GB/s (MEASURED_PEAKS.json)')
    print(f'{"kernel":72s} {"launches":>8s} {"mean us":>9s} {"share":>6s} {"DRAM MB/launch":>15s} {"GB/s":>8s} {"of HBM":>7s}')
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
        mb = (a['rd'] + a['wr']) / a['n'] / 1e6
        gbs = (a['rd'] + a['wr']) / (a['us'] * 1e-6) / 1e9 if a['us'] else 0.0
        print(f'{name[:72]:72s} {a["n"]:8d} {a["us"] / a["n"]:9.1f} {a["us"] / total:6.1%} {mb:15.2f} {gbs:8.0f} {gbs / peak:7.2f}')


if __name__ == '__main__':
    main()
