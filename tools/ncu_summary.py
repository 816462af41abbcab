#!/usr/bin/env python
"""Summarise ncu output for profiles/.

    tools/ncu_summary.py full  <report.ncu-rep>     # key counters per captured launch (from --set full)
    tools/ncu_summary.py list  <launches.csv>       # per-kernel totals / shares of a gpu__time_duration launch list
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    'gpu__time_duration.sum',
    'launch__grid_size', 'launch__block_size', 'launch__cluster_size', 'launch__registers_per_thread',
    'launch__shared_mem_per_block_dynamic',
    'sm__cycles_elapsed.avg.per_second',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
    'sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg',
    'sm__inst_executed_pipe_uniform.sum', 'sm__inst_executed_pipe_tensor.sum',
    'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__cycles_active.avg', 'sm__cycles_active.avg',
]


def short(name):
    name = re.sub(r'^void\s+', '', name)
    name = re.sub(r'\(.*$', '', name)
    if len(name) > 90:
        name = name[:87] + '...'
    return name


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print('## launch id %s  %s' % (r[col['ID']], short(r[col['Kernel Name']])))
        for k in KEYS:
            hit = [h for h in hdr if h == k or h.endswith('.' + k)]
            for h in hit:
                print('  %-80s %s %s' % (k, r[col[h]], units[col[h]]))
        print()


def launch_list(path):
    with open(path) as fh:
        lines = [l for l in fh if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    tot = OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(',', ''))
        if r[ui] == 'us':
            v *= 1e3
        elif r[ui] == 'ms':
            v *= 1e6
        k = short(r[ki])
        n, t = tot.get(k, (0, 0.0))
        tot[k] = (n + 1, t + v)
    allt = sum(t for _, t in tot.values())
    print('%-92s %7s %12s %10s %7s' % ('kernel', 'calls', 'total_us', 'avg_us', 'share'))
    for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print('%-92s %7d %12.1f %10.2f %6.1f%%' % (k, n, t / 1e3, t / n / 1e3, 100 * t / allt))
    print('%-92s %7d %12.1f' % ('TOTAL', sum(n for n, _ in tot.values()), allt / 1e3))


if __name__ == '__main__':
    {'full': full, 'list': launch_list}[sys.argv[1]](sys.argv[2])
