#!/usr/bin/env python
''' Per CUDA source line totals from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name K > f.csv`:
    python profiles/src_lines.py f.csv [top]   -> executed warp instructions and stall samples per source line '''
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
agg = {}
fname = ''
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]
        continue
    if r[0] == 'Line No':
        hdr = r
        ia = hdr.index('Instructions Executed'); iss = hdr.index('Warp Stall Sampling (All Samples)')
        continue
    if hdr is None or len(r) <= ia:
        continue
    if r[0].isdigit() and r[2] == '-':           # a CUDA source line (the SASS lines under it carry an address)
        try:
            n, s = int(r[ia]), int(r[iss])
        except ValueError:
            continue
        key = (fname, int(r[0]), r[1].strip())
        a = agg.setdefault(key, [0, 0])
        a[0] += n; a[1] += s
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print(f'total warp instructions {tot}, stall samples {tots}')
for (f, ln, src), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'{n:10d} {100 * n / tot:5.1f}%  stalls {100 * s / max(tots, 1):5.1f}%  {f}:{ln}: {src[:100]}')
