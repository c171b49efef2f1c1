#!/usr/bin/env python
"""Summary of ONE kernel launch of an .ncu-rep: headline metrics, stall reasons per issue, hottest source lines.
Usage: python tools/ncu_kernel_summary.py REPORT KERNEL_REGEX [launch_skip] [top_n]"""
import csv
import io
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
flt = ["--kernel-name", f"regex:{kre}", "--launch-skip", skip, "--launch-count", "1"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + flt, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sectors.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max"]
print("== headline ==")
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f"{h:75s} {v[:60]} {u}")
print("== stall reasons per issue ==")
for h, u, v in zip(hdr, units, vals):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        try:
            if float(v) >= 0.05:
                print(f"  {h.split('stalled_')[1].split('_per_issue')[0]:25s} {float(v):.3f}")
        except ValueError:
            pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + flt,
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
try:
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
except StopIteration:
    sys.exit(0)
h = rows[hi]
ci = {name: i for i, name in enumerate(h)}
tot_s = tot_i = 0
lines = []
for r in rows[hi + 1:]:
    if len(r) < 10 or r[0] == "":
        continue
    try:
        s = int(r[4]); inst = int(r[ci["Instructions Executed"]])
    except ValueError:
        continue
    lines.append((int(r[0]), r[1].strip()[:110], s, inst))
    tot_s += s; tot_i += inst
print(f"== per source line (samples {tot_s}, warp instructions {tot_i}) ==")
for l in sorted(lines, key=lambda x: -x[2])[:top]:
    print(f"{l[0]:5d} samp {100*l[2]/max(tot_s,1):5.1f}%  inst {100*l[3]/max(tot_i,1):5.1f}%  {l[1]}")
