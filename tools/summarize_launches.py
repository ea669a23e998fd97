#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.txt
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if 'Kernel Name' in r:
            hdr, start = r, i + 1
            break
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        name = r[ki].split('(')[0][:80]
        v = float(r[vi].replace(',', ''))
        v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f'# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.1f} us total (ncu, serialised, cold cache)')
    print(f'{"us":>10s} {"share":>6s} {"n":>5s}  kernel')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{v[1]:10.1f} {100 * v[1] / tot:5.1f}% {v[0]:5d}  {k}')


if __name__ == '__main__':
    main(sys.argv[1])
