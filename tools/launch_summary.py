#!/usr/bin/env python3
'''Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
python tools/launch_summary.py FILE.csv [first_launch [last_launch]]'''
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    rows = []
    for row in r:
        v = float(row[vi].replace(',', ''))
        v = v / 1e3 if row[ui] == 'ns' else v * 1e3 if row[ui] == 'ms' else v
        rows.append((re.sub(r'\(.*', '', row[ki]), v))
    return rows


if __name__ == '__main__':
    rows = load(sys.argv[1])
    a = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    b = int(sys.argv[3]) if len(sys.argv) > 3 else len(rows)
    win = rows[a:b]
    agg = collections.OrderedDict()
    for k, v in win:
        e = agg.setdefault(k, [0, 0.0])
        e[0] += 1
        e[1] += v
    tot = sum(v for _, v in win)
    print('launches %d..%d of %d: %d launches, %.1f us' % (a, b, len(rows), len(win), tot))
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print('%-72s n=%5d %10.1f us %5.1f%%  avg %7.1f' % (k[:72], n, t, 100 * t / tot, t / n))
