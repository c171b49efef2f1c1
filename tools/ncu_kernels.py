#!/usr/bin/env python
"""One line per kernel launch of an .ncu-rep: duration, DRAM bytes, DRAM %, issue %, warps %, instructions, L2 sectors,
L1 global-load sectors, L1 hit %, shared bank conflicts, registers, grid.  Usage: python tools/ncu_kernels.py rep [name-filter]"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"), ("smsp__inst_executed.sum", "inst"), ("lts__t_sectors.sum", "l2_sectors"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1_ld_sectors"), ("l1tex__t_sector_hit_rate.pct", "l1hit%"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]
idx = [(h.index(w), n) for w, n in want if w in h]
units = rows[1]
print(" | ".join(f"{n}[{units[i]}]" if units[i] else n for i, n in idx))
flt = sys.argv[2] if len(sys.argv) > 2 else ""
seen = set()
for r in rows[2:]:
    if flt and flt not in r[idx[0][0]]:
        continue
    key = r[idx[0][0]][:24]
    if key in seen:
        continue
    seen.add(key)
    print(" | ".join(r[i][:26] for i, _ in idx))
