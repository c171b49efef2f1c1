#!/usr/bin/env python
"""configs[1] solve alone: E synthetic C1 entities resident on the device, K timed launches of gdmix_re_fit.
   python tools/c1_probe.py [--entities E] [--steps K] [--profile]   (--profile: cudaProfilerStart around ONE launch,
   for `ncu --profile-from-start off`)"""
import argparse, ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gdmix_b200 import _capi as capi
from gdmix_b200.synthetic import make_device_batch
from tools import subbench

ap = argparse.ArgumentParser()
ap.add_argument("--entities", type=int, default=444 * 600)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--profile", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
w = bench.WORKLOAD
data = make_device_batch(a.entities, w["n"], w["d"], w["k"], seed=w["seed"], device=dev)
cb = subbench._re_batch(data)
opts = capi.make_opts(l2=w["l2"], regularize_bias=False, has_intercept=True)
ws = torch.empty(max(capi.re_workspace_size(cb, opts), 256), dtype=torch.uint8, device=dev)
theta = torch.empty(data["n_coef"], dtype=torch.float64, device=dev)
nit = torch.empty(a.entities, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream

def launch():
    capi.check(capi.lib.gdmix_re_fit(C.byref(cb), C.byref(opts), None, C.c_void_p(theta.data_ptr()), None,
                                     C.c_void_p(nit.data_ptr()), None, None, None, C.c_void_p(ws.data_ptr()),
                                     C.c_size_t(ws.numel()), C.c_void_p(st)))

launch(); launch()
torch.cuda.synchronize()
if a.profile:
    torch.cuda.profiler.start()
    launch()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
else:
    ms = []
    for _ in range(a.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); launch(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    med = ms[len(ms) // 2]
    print(json.dumps({"entities": a.entities, "ms": round(med, 3), "M_entities_per_s": round(a.entities / med / 1e3, 4),
                      "plan": capi.last_plan(), "mean_nit": float(nit.float().mean()),
                      "theta_sum": float(theta.sum()), "theta_abs": float(theta.abs().sum())}))
