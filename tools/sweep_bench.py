#!/usr/bin/env python
"""L2 sweep on the C4 shape (BASELINE.json configs[4]: entities x avg 64 samples, 64 local features, 16 nnz per
sample, l2 in {0.1, 1, 10, 100} solved from ONE staged copy of every entity block) on one B200, next to the
same four weights as four separate gdmix_re_fit calls.  Reports (entity, l2) models/s and entities/s.
Usage: python tools/sweep_bench.py [entities] [iters]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gdmix_b200 import _capi as capi
from gdmix_b200.synthetic import make_device_batch

E = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n, d, k = 64, 64, 16
L2 = np.array([0.1, 1.0, 10.0, 100.0])
dev = torch.device("cuda", 0)
data = make_device_batch(E, n, d, k, seed=4, device=dev, ragged=True)
cb = capi.ReBatch(E, data["n_rows"], data["nnz"], data["ent_rowptr"].data_ptr(), data["rowptr"].data_ptr(),
                  data["col"].data_ptr(), data["val"].data_ptr(), data["label"].data_ptr(), None,
                  data["offset"].data_ptr(), data["theta_ptr"].data_ptr(), data["max_rows"], data["max_nnz"],
                  data["max_coef"], 0)
opts = capi.make_opts(l2=1.0)
ws = torch.empty(max(capi.re_workspace_size(cb, opts), 256), dtype=torch.uint8, device=dev)
nc = data["n_coef"]
theta = torch.empty((len(L2), nc), dtype=torch.float64, device=dev)
f = torch.empty((len(L2), E), dtype=torch.float64, device=dev)
nit = torch.empty((len(L2), E), dtype=torch.int32, device=dev)
nfev = torch.empty_like(nit)
status = torch.empty_like(nit)
st = torch.cuda.current_stream()
P = lambda t: C.c_void_p(t.data_ptr())


def sweep():
    capi.check(capi.lib.gdmix_re_fit_sweep(C.byref(cb), C.byref(opts), L2.ctypes.data_as(C.c_void_p), C.c_int32(len(L2)),
                                           None, P(theta), C.c_int64(nc), P(f), P(nit), P(nfev), P(status), P(ws),
                                           C.c_size_t(ws.numel()), C.c_void_p(st.cuda_stream)))


theta1 = torch.empty((len(L2), nc), dtype=torch.float64, device=dev)


def separate():
    for j, l2 in enumerate(L2):
        o = capi.make_opts(l2=float(l2))
        capi.check(capi.lib.gdmix_re_fit(C.byref(cb), C.byref(o), None, P(theta1[j]), None, None, None, None, None,
                                         P(ws), C.c_size_t(ws.numel()), C.c_void_p(st.cuda_stream)))


def timed(fn):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


ms_sweep = timed(sweep)
ms_sep = timed(separate)
same = bool(torch.equal(theta, theta1))
alg = 8 * data["nnz"] + 16 * data["n_rows"] + len(L2) * 8 * nc + 4 * (nc - E)
print(json.dumps({"workload": f"c4 shape: {E} entities x avg {n} samples (ragged 8..1024) x {d} local features, {k} nnz/sample, "
                              f"l2 in {L2.tolist()}", "plan": capi.last_plan(),
                  "sweep_ms": ms_sweep, "models_per_s": len(L2) * E / ms_sweep * 1e3, "entities_per_s": E / ms_sweep * 1e3,
                  "separate_fits_ms": ms_sep, "separate_models_per_s": len(L2) * E / ms_sep * 1e3,
                  "speedup_vs_separate": ms_sep / ms_sweep, "bitwise_equal_to_separate_fits": same,
                  "algorithmic_GBps": alg / ms_sweep / 1e6,
                  "converged_frac": float((status == 0).float().mean().item()),
                  "mean_nit_per_l2": nit.float().mean(1).tolist()}))
