#!/usr/bin/env python
"""Register-move instructions (MOV / IMAD.MOV / CS2R) executed per source line of a `--page source --print-source
cuda,sass --csv` dump.  Usage: python tools/ncu_movs.py dump.csv units"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1]))); units = float(sys.argv[2])
cur = line = None; cnt = collections.Counter(); allc = collections.Counter(); txt = {}; tot = 0; total = 0
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No": ie = r.index("Instructions Executed"); continue
    if len(r) < 6: continue
    if r[0].isdigit():
        line = (cur, int(r[0])); txt[line] = r[1].strip()[:90]; continue
    if r[2] == '...': continue
    op = r[3].split()
    if not op: continue
    o = op[1] if op[0].startswith('@') else op[0]
    try: n = int(r[ie])
    except Exception: continue
    total += n; allc[line] += n
    if o.startswith('IMAD.MOV') or o == 'MOV' or o.startswith('CS2R'):
        cnt[line] += n; tot += n
print(f"all instructions per unit {total / units:.0f}, register moves {tot / units:.0f}")
for k, v in cnt.most_common(14): print(k, round(v / units, 1), "of", round(allc[k] / units, 1), txt[k])
