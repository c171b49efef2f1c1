#!/usr/bin/env python
"""BASELINE.json configs[3] on one GPU at a rank's share of the rows: fixed effect -> per-user random effect ->
per-item random effect, each stage's fp32 scores the next stage's offsets, everything on the device (FE solve,
scoring, group-by-user radix sort + entity-local indexing, batched RE solve, scoring, regroup by item, RE solve,
AUC after every coordinate).  Seconds per stage; no oracle here (tests/test_chain_gpu.py checks the same chain
against the CPU oracle at test scale).
Usage: python tools/chain_bench.py [rows] [users] [items] [fe_iters]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gdmix_b200 import _capi as capi, partition as P
from gdmix_b200.fe_solver import FixedEffectSolver

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000
U = int(sys.argv[2]) if len(sys.argv) > 2 else 500_000
I = int(sys.argv[3]) if len(sys.argv) > 3 else 50_000
fe_iters = int(sys.argv[4]) if len(sys.argv) > 4 else 20
D, k, Du, ku, Di, ki = 100_000, 32, 64, 8, 64, 8
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(3)


def bag(nrows, Dn, kn, zipf):
    """Sorted unique columns per row: a random start inside each of kn strata of the feature range."""
    stride = Dn // kn
    if zipf:
        u = torch.rand(nrows, kn, device=dev, generator=g)
        within = ((stride ** (1.0 - u.double())) - 1.0).clamp_(0, stride - 1).to(torch.int32)
    else:
        within = torch.randint(0, stride, (nrows, kn), dtype=torch.int32, device=dev, generator=g)
    col = (torch.arange(kn, dtype=torch.int32, device=dev) * stride)[None, :] + within
    val = torch.randn(nrows, kn, device=dev, generator=g)
    return torch.arange(nrows + 1, dtype=torch.int64, device=dev) * kn, col.reshape(-1).contiguous(), val.reshape(-1).contiguous()


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


t0 = sync()
user = (torch.rand(n, device=dev, generator=g) ** 2 * U).to(torch.int64)          # power-law user activity
item = ((I ** torch.rand(n, device=dev, generator=g).double()) - 1.0).clamp_(0, I - 1).to(torch.int64)  # Zipf-like items
g_rp, g_col, g_val = bag(n, D, k, True)
u_rp, u_col, u_val = bag(n, Du, ku, False)
i_rp, i_col, i_val = bag(n, Di, ki, False)
y = (torch.rand(n, device=dev, generator=g) < 0.45).float()
t_gen = sync() - t0
out = {"workload": f"c3 share: {n} rows, {U} users, {I} items, FE D={D} k={k}, per-user d={Du} k={ku}, per-item d={Di} k={ki}",
       "generation_s": t_gen}

# ---- stage 1: fixed effect (fe_iters L-BFGS iterations; C2 / C3 stop on max_iter at scale)
opts_fe = capi.make_opts(l2=1.0, regularize_bias=True, max_iter=fe_iters)
rows = capi.DeviceFeRows.__new__(capi.DeviceFeRows)
rows.rowptr, rows.col, rows.val, rows.label, rows.weight, rows.offset = g_rp, g_col, g_val, y, None, None
rows.n_rows, rows.nnz, rows.n_features, rows.linear_regression, rows.num_workers = n, n * k, D, False, 1
solver = FixedEffectSolver(rows, opts_fe, D)
t0 = sync()
solver._prepare()
t_plan = sync() - t0
t0 = sync()
x, info = solver.fit()
t_fe = sync() - t0
xd = torch.from_numpy(x).to(dev)
t0 = sync()
s0, _ = capi.fe_score_device(rows, opts_fe, xd)
t_score0 = sync() - t0
out["fe"] = {"plan_s": t_plan, "fit_s": t_fe, "nit": info["nit"], "nfev": info["nfev"], "evals_per_s": info["nfev"] / t_fe,
             "score_s": t_score0, "auc": P.auc(s0, y)}
del solver, rows, g_rp, g_col, g_val
torch.cuda.empty_cache()


def re_stage(name, keys, rp, col, val, off, Dn):
    t0 = sync()
    gb = P.GroupedBatch(P.regroup_batch(keys, rp, col, val, y, off, None, num_features=Dn))
    t_group = sync() - t0
    opts = capi.make_opts(l2=1.0, regularize_bias=False)
    t0 = sync()
    fit = capi.re_fit_device(gb, opts)
    t_fit = sync() - t0
    plan = capi.last_plan()
    t0 = sync()
    logit, _ = capi.re_score_device(gb, opts, fit["theta"])
    s = gb.scatter_to_input_order(logit)
    t_score = sync() - t0
    st = fit["status"]
    E = st.numel()
    cnt = fit["workspace"][:32].view(torch.int32).cpu().numpy()
    out[name] = {"entities": E, "max_rows": gb.host.max_rows, "group_s": t_group, "fit_s": t_fit, "entities_per_s": E / t_fit,
                 "rows_per_s": n / t_fit, "score_s": t_score, "converged_frac": float((st == 0).double().mean().item()),
                 "rejected": int((st < 0).sum().item()), "mean_nit": float(fit["nit"].float().mean().item()),
                 "plan": plan, "deferred_typical": int(cnt[7]), "deferred_to_general": int(cnt[2]),
                 "deferred_to_global_x": int(cnt[5]), "auc": P.auc(s, y)}
    return s


s1 = re_stage("per_user", user, u_rp, u_col, u_val, s0, Du)
s2 = re_stage("per_item", item, i_rp, i_col, i_val, s1, Di)
print(json.dumps(out))
