#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
launch count, total device time and share.  Usage: tools/ncu_summary.py launches.csv [title]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name, val, unit = row.get('Kernel Name'), row.get('Metric Value'), row.get('Metric Unit')
        if not name or not val:
            continue
        v = float(val.replace(',', '')) * {'ns': 1.0, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(unit, 1.0)
        short = re.sub(r'\(.*', '', name).replace('void ', '')[:80]
        agg[short][0] += 1
        agg[short][1] += v
    tot = sum(v[1] for v in agg.values())
    print('# %s\n' % title)
    print('ncu `gpu__time_duration.sum`, `--clock-control none` (cold-cache, serialised launches: compare shares).')
    print('%d launches, %.3f ms total.\n' % (sum(v[0] for v in agg.values()), tot / 1e6))
    print('| kernel | launches | total ms | share | avg us |')
    print('|---|---:|---:|---:|---:|')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| `%s` | %d | %.3f | %.1f%% | %.1f |' % (k, v[0], v[1] / 1e6, 100 * v[1] / tot, v[1] / v[0] / 1e3))


if __name__ == '__main__':
    main()
