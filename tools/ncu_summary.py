#!/usr/bin/env python
"""Compact summary of `ncu --set full` reports: python tools/ncu_summary.py a.ncu-rep [b.ncu-rep ...]"""
import csv
import subprocess
import sys

METRICS = [
    'gpu__time_duration.sum',
    'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
    'sm__cycles_elapsed.max',
    'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
    'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed',
    'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_sector_hit_rate.pct',
    'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum',
    'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            name = vals[hdr.index('Kernel Name')][:90]
            print(f'== {path}: {name}')
            for m in METRICS:
                hits = [i for i, h in enumerate(hdr) if h == m or h.endswith('.' + m)]
                for i in hits[:1]:
                    print(f'   {m:75s} {vals[i]:>16s} {units[i]}')


if __name__ == '__main__':
    main()
