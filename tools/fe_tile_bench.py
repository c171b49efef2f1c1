#!/usr/bin/env python
"""One objective / gradient evaluation of the tiled fixed-effect path on a Zipf-like shard (bench shape).
Usage: python tools/fe_tile_bench.py [rows] [iters] [hz] [hg] [tile_rows]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gdmix_b200 import _capi as capi
from gdmix_b200.fe_solver import FixedEffectSolver
from tools import subbench

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
pa = {}
for k, i in (("hz", 3), ("hg", 4), ("tile_rows", 5)):
    if len(sys.argv) > i and int(sys.argv[i]) > 0:
        pa[k] = int(sys.argv[i])
dev = torch.device("cuda", 0)
shard = subbench.zipf_rows(rows, 100_000, 32, 100, dev)
opts = capi.make_opts(l2=1.0, regularize_bias=True)
s = FixedEffectSolver(shard, opts, 100_000, plan_args=pa)
s._prepare()
torch.cuda.synchronize()
ms = subbench._timed(s._partial, iters, warm=2)
p = s.plan
print(json.dumps({"rows": rows, "ms": ms, "algorithmic_GBps": rows * 272 / ms / 1e6, "hz": p.hz, "hg": p.hg, "tile_rows": p.tile_rows,
                  "tiles": p.n_tiles, "cold_z_frac": p.n_cold_z / (rows * 32), "cold_g_frac": p.n_cold_g / (rows * 32),
                  "plan_bytes_per_row": p.bytes / rows}))
