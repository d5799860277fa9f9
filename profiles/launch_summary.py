#!/usr/bin/env python
''' Per-kernel totals of an ncu launch list (gpu__time_duration.sum CSV):  python profiles/launch_summary.py launches.csv '''
import collections
import csv
import gzip
import re
import sys
path = sys.argv[1]
op = gzip.open if path.endswith('.gz') else open
with op(path, 'rt') as f:
    lines = [l for l in f if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v = v / 1e3 if unit == 'ns' else (v * 1e3 if unit == 'ms' else v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print(f'{k[:72]:72s} n={n:5d} total={t / 1e3:9.2f} ms  mean={t / n:8.1f} us  share={100 * t / tot:5.1f}%')
