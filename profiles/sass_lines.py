'''
Per-source-line instruction counts of one profiled kernel: joins the SASS listing of an ncu report (per-instruction "Instructions
Executed" / stall samples; `ncu -i X.ncu-rep --page source --csv --kernel-name regex:NAME`) with the line table of the same kernel
in the library's cubin (`nvdisasm -g`), instruction by instruction.

    python profiles/sass_lines.py gpurun_out/fused_a.ncu-rep day_begin '(bool)1, (bool)1, (bool)1, (bool)1' day_begin_kernelILb1ELb1ELb1ELb1
'''
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kregex, name_part, mangled_part = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, 'covasim_b200', 'libcovasim_b200.so')
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', f'regex:{kregex}'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
first = next(k for k in range(len(starts) - 1) if name_part in rows[starts[k]][1])
rows = rows[starts[first]:starts[first + 1]]            # the first launch whose name contains name_part
print('profiled launch:', rows[0][1][:120])
h = next(i for i, r in enumerate(rows) if 'Source' in r and 'Address' in r)
hdr = rows[h]
ii, si, sa = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
ti = hdr.index('Thread Instructions Executed')
sass = [(r[si].strip(), int(r[ii] or 0), int(r[sa] or 0), int(r[ti] or 0)) for r in rows[h + 1:] if len(r) > ii]
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', lib], cwd=tmp, capture_output=True)
lines = None
for f in os.listdir(tmp):
    if not f.endswith('.cubin'):
        continue
    syms = subprocess.run(['cuobjdump', '-elf', os.path.join(tmp, f)], capture_output=True, text=True).stdout
    m = re.search(r'\.text\.(\S*' + re.escape(mangled_part) + r'\S*)', syms)
    if not m:
        continue
    dis = subprocess.run(['nvdisasm', '-g', os.path.join(tmp, f)], capture_output=True, text=True).stdout
    lines, inside = [], False
    cur = ('?', 0)
    for ln in dis.splitlines():
        if ln.startswith('.text.'):
            inside = ln.startswith('.text.' + m.group(1) + ':')
            continue
        if not inside:
            continue
        if ln.lstrip().startswith('.section'):
            inside = False
            continue
        mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if mm:
            cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s', ln):
            lines.append(cur)
    break
assert lines is not None, 'kernel not found in the library'
assert len(lines) == len(sass), (len(lines), len(sass))
agg = {}
for (src, n, smp, tn), key in zip(sass, lines):
    a = agg.setdefault(key, [0, 0, 0, 0])
    a[0] += n; a[1] += smp; a[2] += 1; a[3] += tn
tot = sum(a[0] for a in agg.values())
tots = sum(a[1] for a in agg.values())
print(f'kernel {kregex}: {len(sass)} SASS instructions, {tot} warp instructions executed, {tots} stall samples')
text = {}
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1 if os.environ.get('BY') == 'smp' else 0])[:int(os.environ.get('TOP', 45))]:
    if f not in text:
        p = os.path.join(ROOT, 'covasim_b200', 'csrc', f)
        text[f] = open(p).read().splitlines() if os.path.exists(p) else []
    srcline = text[f][l - 1].strip()[:110] if 0 < l <= len(text[f]) else ''
    print(f'{100 * a[0] / tot:5.1f}% inst {100 * a[1] / max(tots, 1):5.1f}% smp  {a[2]:5d} sass  thr/inst {a[3] / max(a[0], 1):4.1f}  {f}:{l}  {srcline}')
