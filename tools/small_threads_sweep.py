#!/usr/bin/env python
"""configs[3] per-user shape against the threads-per-entity override (0 = the planner: 32 threads is its choice and the
best: 14.7 / 14.8 / 14.1 / 9.9 M entities/s for auto / 32 / 64 / 128).  Usage: python tools/small_threads_sweep.py"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ctypes as C
from gdmix_b200 import _capi as capi
from gdmix_b200.synthetic import make_device_batch
from tools import subbench
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
E = 2_000_000
data = make_device_batch(E, 32, 64, 8, seed=7, device=dev)
cb = subbench._re_batch(data)
for tpe in (0, 32, 64, 128):
    opts = capi.make_opts(l2=1.0, threads_per_entity=tpe)
    ws = torch.empty(max(capi.re_workspace_size(cb, opts), 256), dtype=torch.uint8, device=dev)
    theta = torch.empty(data["n_coef"], dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream()
    def run():
        capi.check(capi.lib.gdmix_re_fit(C.byref(cb), C.byref(opts), None, C.c_void_p(theta.data_ptr()), None, None, None, None, None,
                                         C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel()), C.c_void_p(st.cuda_stream)))
    ms = subbench._timed(run, 3)
    print(tpe, round(E / ms / 1e3, 2), "M entities/s", capi.last_plan())
