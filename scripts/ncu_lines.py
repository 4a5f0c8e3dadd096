#!/usr/bin/env python
"""Hottest CUDA source lines (warp-stall samples, executed instructions) of a kernel in an .ncu-rep.
The report's SASS page is joined with `nvdisasm -g` line info of the in-tree library by instruction offset, so the
.so must be the build the report was captured from.
usage: python scripts/ncu_lines.py X.ncu-rep <unit: extract|grid|pose|raster> <kernel substring> [topN]"""
import csv, glob, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, unit, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.join(ROOT, 'mov-slam_b200', 'lib', 'libmovfe.so')], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, unit + '*.cubin'))[0]
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout.splitlines()
line_of, cur, inside = {}, None, False
for l in dis:
    if l.startswith('.text.'):
        inside = kern in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File ".*?", line (\d+)', l)
    if m:
        cur = int(m.group(1))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m:
        line_of[int(m.group(1), 16)] = cur
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
st, ie = hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
base = None
agg = {}
for r in rows[hi + 1:]:
    if len(r) <= max(st, ie) or not r[0].startswith('0x'):
        continue
    a = int(r[0], 16)
    base = a if base is None else base
    ln = line_of.get(a - base)
    s, n = int(r[st] or 0), int(r[ie] or 0)
    x = agg.setdefault(ln, [0, 0])
    x[0] += s
    x[1] += n
src = open(os.path.join(ROOT, 'mov-slam_b200', 'csrc', unit + '.cu')).read().splitlines()
tot = sum(v[0] for v in agg.values()) or 1
toti = sum(v[1] for v in agg.values()) or 1
print('total samples %d, warp instructions %d' % (tot, toti))
for ln, (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = src[ln - 1].strip()[:110] if ln and ln <= len(src) else '?'
    print('%6d %5.1f%% | inst %5.1f%% | L%s: %s' % (s, 100 * s / tot, 100 * n / toti, ln, text))
