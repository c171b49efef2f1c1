#!/usr/bin/env python
"""Per-source-line breakdown of an .ncu-rep (samples, warp instructions, shared-memory wavefronts vs ideal).
Usage: python tools/ncu_lines.py file.ncu-rep [min_pct]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
out, cur, ci = {}, None, None
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No":
        ci = {}
        for i, n in enumerate(r):
            ci.setdefault(n, i)
        continue
    if ci is None or len(r) < 10 or not r[0].isdigit():
        continue
    def g(n):
        try: return int(r[ci[n]])
        except Exception: return 0
    key = (cur, int(r[0]))
    o = out.setdefault(key, [r[1].strip()[:88], 0, 0, 0, 0])
    o[1] += g("# Samples"); o[2] += g("Instructions Executed"); o[3] += g("L1 Wavefronts Shared"); o[4] += g("L1 Wavefronts Shared Ideal")
ts = sum(o[1] for o in out.values()); ti = sum(o[2] for o in out.values()); tw = max(1, sum(o[3] for o in out.values()))
print(f"samples {ts}  warp-inst {ti}  smem wavefronts {tw} (ideal {sum(o[4] for o in out.values())})")
per_file = {}
for (f, l), o in out.items():
    a = per_file.setdefault(f, [0, 0, 0]); a[0] += o[1]; a[1] += o[2]; a[2] += o[3]
for f, a in per_file.items():
    print(f"  {f:18s} samp {100*a[0]/ts:5.1f}%  inst {100*a[1]/ti:5.1f}%  wavefronts {100*a[2]/tw:5.1f}%")
for (f, l), o in sorted(out.items()):
    if 100*o[1]/ts >= minp or 100*o[2]/ti >= minp or 100*o[3]/tw >= minp:
        print(f"{f:16s}{l:4d} s{100*o[1]/ts:5.1f} i{100*o[2]/ti:5.1f} w{100*o[3]/tw:5.1f} (ideal {100*o[4]/tw:4.1f}) | {o[0]}")
if len(sys.argv) > 3:
    print("== top by instructions ==")
    for (f, l), o in sorted(out.items(), key=lambda kv: -kv[1][2])[:int(sys.argv[3])]:
        print(f"{f:16s}{l:4d} i{100*o[2]/ti:5.1f} s{100*o[1]/ts:5.1f} | {o[0]}")
