#!/usr/bin/env python
"""One profiled launch of the configs[3] per-user shape (32 samples x 64 features x 8 nnz): for ncu --profile-from-start off."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gdmix_b200 import _capi as capi
from gdmix_b200.synthetic import make_device_batch
from tools import subbench
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
E = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
data = make_device_batch(E, 32, 64, 8, seed=7, device=dev)
cb = subbench._re_batch(data)
opts = capi.make_opts(l2=1.0)
ws = torch.empty(max(capi.re_workspace_size(cb, opts), 256), dtype=torch.uint8, device=dev)
theta = torch.empty(data["n_coef"], dtype=torch.float64, device=dev)
def launch():
    capi.check(capi.lib.gdmix_re_fit(C.byref(cb), C.byref(opts), None, C.c_void_p(theta.data_ptr()), None, None, None, None,
                                     None, C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel()),
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)))
launch(); torch.cuda.synchronize()
torch.cuda.profiler.start(); launch(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("plan", capi.last_plan())
