#!/usr/bin/env python
"""Dynamic SASS opcode histogram of an .ncu-rep (warp instructions executed per opcode).
Usage: python tools/ncu_opcodes.py file.ncu-rep [top_n]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; ci = {}
for i, n in enumerate(h): ci.setdefault(n, i)
cnt = collections.Counter(); samp = collections.Counter()
for r in rows[hi + 1:]:
    if len(r) < 5: continue
    try:
        n = int(r[ci["Instructions Executed"]]); sm = int(r[ci["# Samples"]])
    except Exception: continue
    txt = r[ci["Source"]].strip()
    parts = txt.split()
    op = parts[0]
    if op.startswith("@") and len(parts) > 1: op = parts[1]
    op = ".".join(op.split(".")[:2]) if op.startswith(("IMAD", "LDS", "STS", "SHFL", "F2F", "LDG", "STG", "LDL", "STL")) else op.split(".")[0]
    cnt[op] += n; samp[op] += sm
tot = sum(cnt.values()); ts = sum(samp.values())
print("total warp instructions", tot)
for op, n in cnt.most_common(top):
    print(f"{op:14s} {100*n/tot:5.1f}%  samples {100*samp[op]/max(ts,1):5.1f}%")
