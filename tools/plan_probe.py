#!/usr/bin/env python
"""Seconds to build the tiled fixed-effect plan of a C2-sized shard, three times in a row (allocation state varies).
Usage: python tools/plan_probe.py [rows]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gdmix_b200 import _capi as capi
from gdmix_b200.fe_solver import FixedEffectSolver
from tools import subbench
dev = torch.device("cuda", 0)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 62_500_000
shard = subbench.zipf_rows(rows, 100_000, 32, 100, dev)
opts = capi.make_opts(l2=1.0, regularize_bias=True)
for rep in range(3):
    s = FixedEffectSolver(shard, opts, 100_000)
    torch.cuda.synchronize(); t = time.perf_counter()
    s._prepare()
    torch.cuda.synchronize(); print("prepare", round(time.perf_counter() - t, 3), "s", "free GB", torch.cuda.mem_get_info()[0] / 1e9)
    del s
