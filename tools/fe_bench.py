#!/usr/bin/env python
"""Fixed-effect objective/gradient pass (gdmix_fe_loss_grad) on synthetic rows of the C2 shape: rows x 32 nnz over
D = 100 000 features (Zipf-ish popularity), GB/s against the 8k+16 B/row algorithmic traffic.
Usage: python tools/fe_bench.py [rows] [D] [k] [iters] [tile_rows]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gdmix_b200 import _capi as capi

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
D = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 32
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 10
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
# Zipf(1.1)-like feature popularity through an inverse-CDF on a power law
u = torch.rand(rows * k, device=dev, generator=g)
col = ((D ** (1.0 - u.double())) - 1.0).clamp_(0, D - 1).to(torch.int32)
val = torch.randn(rows * k, device=dev, generator=g)
rowptr = torch.arange(rows + 1, device=dev, dtype=torch.int64) * k
label = (torch.rand(rows, device=dev, generator=g) < 0.5).float()


class R:  # duck-typed DeviceFeRows
    pass


r = capi.DeviceFeRows.__new__(capi.DeviceFeRows)
r.rowptr, r.col, r.val, r.label, r.weight, r.offset = rowptr, col, val, label, None, None
r.n_rows, r.nnz, r.n_features, r.linear_regression, r.num_workers = rows, rows * k, D, False, 1
opts = capi.make_opts(l2=1.0, regularize_bias=True)
x = torch.randn(D + 1, device=dev, dtype=torch.float64) * 0.01
fg = torch.empty(D + 2, device=dev, dtype=torch.float64)
for _ in range(3):
    capi.fe_loss_grad_device(r, opts, x, fg=fg)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(iters):
    capi.fe_loss_grad_device(r, opts, x, fg=fg)
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / iters
alg = rows * (8 * k + 16)
print(json.dumps({"kernel": "fe_loss_grad_kernel", "rows": rows, "D": D, "k": k, "ms": ms,
                  "rows_per_s": rows / ms * 1e3, "algorithmic_GBps": alg / ms / 1e6, "fg0": float(fg[0].item())}))
tile_rows = int(sys.argv[5]) if len(sys.argv) > 5 else 0
plan = capi.DeviceFeTilePlan(r, tile_rows=tile_rows)
fg2 = torch.empty_like(fg)
for _ in range(3):
    capi.fe_loss_grad_device(r, opts, x, fg=fg2, plan=plan)
torch.cuda.synchronize()
ev0.record()
for _ in range(iters):
    capi.fe_loss_grad_device(r, opts, x, fg=fg2, plan=plan)
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / iters
rel = float(((fg2 - fg).abs().max() / fg.abs().max()).item())
print(json.dumps({"kernel": "fe tiled (z + g + cold + finish; columns NOT ranked here)", "ms": ms, "rows_per_s": rows / ms * 1e3,
                  "algorithmic_GBps": alg / ms / 1e6, "max_rel_diff_vs_atomic": rel, "tile_rows": plan.tile_rows, "tiles": plan.n_tiles,
                  "hz": plan.hz, "hg": plan.hg}))
lg = capi.fe_score_device(r, opts, x)
torch.cuda.synchronize()
ev0.record()
for _ in range(iters):
    capi.fe_score_device(r, opts, x)
ev1.record(); torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / iters
print(json.dumps({"kernel": "fe_score_kernel", "ms": ms, "algorithmic_GBps": rows * (8 * k + 16) / ms / 1e6}))
