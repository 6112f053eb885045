"""Aggregate an ncu source-page CSV (SASS view) by CUDA source line, using nvdisasm -g line info of the cubin.

    ncu -i rep.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all lib.o ; nvdisasm -g x.cubin > g.txt
    python tools/ncu_by_line.py src.csv g.txt kernel_name_substring [n]
"""
import csv
import re
import sys

src_csv, gtxt, kname = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
cur, infunc, m = None, False, {}
for ln in open(gtxt):
    if '.text.' in ln and ln.lstrip().startswith('.section'):
        infunc = kname in ln
    t = ln.strip()
    mm = re.match(r'//## File "([^"]+)", line (\d+)(.*)', t)
    if mm:
        cur = (mm.group(1).split('/')[-1], int(mm.group(2)))
        continue
    mm = re.match(r'/\*([0-9a-f]{4,})\*/\s+(.*?);', t)
    if mm and infunc:
        m[int(mm.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
h = rows[1]
ix = {k: i for i, k in enumerate(h)}
stalls = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
addrs = [int(r[ix['Address']], 16) for r in rows[2:] if r and r[ix['Address']]]
base = min(addrs)
agg, tot = {}, 0
for r in rows[2:]:
    try:
        n = int(r[ix['# Samples']])
    except Exception:
        continue
    off = int(r[ix['Address']], 16) - base
    key = m.get(off, ('?', 0))
    a = agg.setdefault(key, [0, {}])
    a[0] += n
    tot += n
    for k in stalls:
        v = int(r[ix[k]] or 0)
        if v:
            a[1][k[6:]] = a[1].get(k[6:], 0) + v
print('samples', tot)
srcs = {}
for (f, l), (n, st) in sorted(agg.items(), key=lambda x: -x[1][0])[:topn]:
    if f not in srcs:
        try:
            srcs[f] = open('/root/repo/fastllama_b200/csrc/' + f).read().split('\n')
        except Exception:
            srcs[f] = []
    text = srcs[f][l - 1].strip()[:80] if 0 < l <= len(srcs[f]) else ''
    top = sorted(st.items(), key=lambda x: -x[1])[:3]
    print(f'{n:6d} {100 * n / tot:5.1f}%  {f[:20]:20s}:{l:4d}  {text:80s} {top}')
