"""Splits a kernel's SASS (ncu source page) into regions delimited by BAR.SYNC / tcgen05 / SYNCS markers and prints, per
region, executed warp-instructions and stall samples with an opcode digest. Usage: sass_phases.py rep KERNEL_REGEX"""
import csv, io, subprocess, sys
from collections import Counter
rep, rx = sys.argv[1], sys.argv[2]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
hdr = rows[hi]; col = {h: i for i, h in enumerate(hdr)}
ex, st, src = col['Instructions Executed'], col['Warp Stall Sampling (All Samples)'], col['Source']
regions = []; cur = dict(inst=0, stall=0, ops=Counter(), n=0, first=None, mark='start')
for r in rows[hi + 1:]:
    if len(r) <= ex: continue
    try: e, s = int(r[ex] or 0), int(r[st] or 0)
    except ValueError: continue
    txt = r[src]
    toks = txt.split()
    op = toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '?')
    cur['inst'] += e; cur['stall'] += s; cur['ops'][op.split('.')[0]] += e; cur['n'] += 1
    if cur['first'] is None: cur['first'] = r[col['Address']]
    if op.startswith('BAR') :
        regions.append(cur); cur = dict(inst=0, stall=0, ops=Counter(), n=0, first=None, mark=txt[:40])
regions.append(cur)
ti = sum(x['inst'] for x in regions); ts = sum(x['stall'] for x in regions)
print(f'total warp-inst {ti}  stall samples {ts}  regions {len(regions)}')
for i, x in enumerate(regions):
    if x['inst'] == 0 and x['stall'] == 0: continue
    top = ' '.join(f'{k}:{100*v//max(x["inst"],1)}' for k, v in x['ops'].most_common(7))
    print(f'[{i:2d}] sass {x["n"]:5d}  inst {100*x["inst"]/ti:5.1f}%  stall {100*x["stall"]/max(ts,1):5.1f}%  | {top}')
