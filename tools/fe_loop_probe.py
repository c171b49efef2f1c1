#!/usr/bin/env python
"""Where the wall time of a fixed-effect solve goes besides the kernels (one GPU).  Usage: python tools/fe_loop_probe.py [rows]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gdmix_b200 import _capi as capi
from gdmix_b200.fe_solver import FixedEffectSolver
from tools import subbench

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 62_500_000
dev = torch.device("cuda", 0)
shard = subbench.zipf_rows(rows, 100_000, 32, 100, dev)
opts = capi.make_opts(l2=1.0, regularize_bias=True, max_iter=20)
s = FixedEffectSolver(shard, opts, 100_000, profile=False)
s._prepare()
opts.max_iter = 2; s.fit(); opts.max_iter = 20
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    lb = capi.DeviceLbfgs(s._x_dev, s._fg_dev, opts)
    t1 = time.perf_counter()
    s._x_dev.zero_(); lb.reset()
    marks = []
    task = 1
    while task != 0:
        a = time.perf_counter()
        if task == 1:
            s._evaluate()
        b = time.perf_counter()
        lb.step()
        c = time.perf_counter()
        info = lb.poll(); task = info["task"]
        d = time.perf_counter()
        marks.append((b - a, c - b, d - c))
    t2 = time.perf_counter()
    lb.close()
    t3 = time.perf_counter()
    import numpy as np
    m = np.array(marks) * 1e3
    print(f"create {1e3*(t1-t0):.2f} ms, loop {1e3*(t2-t1):.2f} ms over {len(marks)} evals ({1e3*(t2-t1)/len(marks):.2f} ms/eval), close {1e3*(t3-t2):.2f} ms")
    print("per eval host ms: enqueue eval", m[:, 0].round(3).tolist())
    print("enqueue step", m[:, 1].round(3).tolist())
    print("poll", m[:, 2].round(2).tolist())
t0 = time.perf_counter(); x, info = s.fit(); torch.cuda.synchronize(); print("fit()", time.perf_counter() - t0, info)
