#!/usr/bin/env python
"""Warp instructions per source line of an .ncu-rep, per `unit` (e.g. entity x evaluations), sorted by count.
Usage: python tools/ncu_inst_lines.py file.ncu-rep units [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; units = float(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
out, cur, ci = {}, None, None
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
    key = (cur, int(r[0]))
    o = out.setdefault(key, [r[1].strip()[:95], 0, 0])
    o[1] += g("Instructions Executed"); o[2] += g("# Samples")
ti = sum(o[1] for o in out.values()); ts = sum(o[2] for o in out.values())
print(f"warp instructions per unit: {ti / units:.0f}")
for (f, l), o in sorted(out.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f[:16]:16s} {l:4d} {o[1] / units:8.1f} ({100 * o[1] / ti:4.1f}%) samp {100 * o[2] / ts:4.1f}% | {o[0]}")
