#!/usr/bin/env python
"""Aggregate an .ncu-rep's per-line stats over named line ranges of one file (and whole other files).
Usage: python tools/ncu_regions.py rep file.cuh name:lo-hi name:lo-hi ..."""
import csv, io, subprocess, sys
rep, fname = sys.argv[1], sys.argv[2]
regions = []
for a in sys.argv[3:]:
    nm, rg = a.split(":"); lo, hi = rg.split("-"); regions.append((nm, int(lo), int(hi)))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
agg = {}
cur = ci = None
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No":
        ci = {}
        for i, n in enumerate(r): ci.setdefault(n, i)
        continue
    if ci is None or len(r) < 10 or not r[0].isdigit(): continue
    def g(n):
        try: return int(r[ci[n]])
        except Exception: return 0
    l = int(r[0])
    key = cur
    if cur == fname:
        key = "other:" + fname
        for nm, lo, hi in regions:
            if lo <= l <= hi: key = nm; break
    a = agg.setdefault(key, [0, 0, 0])
    a[0] += g("# Samples"); a[1] += g("Instructions Executed"); a[2] += g("L1 Wavefronts Shared")
ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values()); tw = max(1, sum(a[2] for a in agg.values()))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:28s} samples {100*a[0]/ts:5.1f}%  inst {100*a[1]/ti:5.1f}%  wavefronts {100*a[2]/tw:5.1f}%")
