#!/usr/bin/env python
"""Phase cycles of re_fast_kernel on C1 from a library built with -DGDMIX_FAST_TIMING (kernel experiments):
   GDMIX_LIB_OUT=gdmix_b200/lib/exp/timing.so GDMIX_NVCC_FLAGS=-DGDMIX_FAST_TIMING python gdmix_b200/build.py
   GDMIX_B200_LIB=$PWD/gdmix_b200/lib/exp/timing.so python tools/c1_phases.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gdmix_b200 import _capi as capi
from gdmix_b200.synthetic import make_device_batch
from tools import subbench
E = int(sys.argv[1]) if len(sys.argv) > 1 else 444 * 100
dev = torch.device("cuda", 0)
w = bench.WORKLOAD
data = make_device_batch(E, w["n"], w["d"], w["k"], seed=w["seed"], device=dev)
cb = subbench._re_batch(data)
opts = capi.make_opts(l2=w["l2"], regularize_bias=False, has_intercept=True)
ws = torch.empty(max(capi.re_workspace_size(cb, opts), 256), dtype=torch.uint8, device=dev)
theta = torch.empty(data["n_coef"], dtype=torch.float64, device=dev)
out = (C.c_ulonglong * 12)()
for i in range(2):
    capi.check(capi.lib.gdmix_re_fit(C.byref(cb), C.byref(opts), None, C.c_void_p(theta.data_ptr()), None, None, None,
                                     None, None, C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel()),
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    capi.lib.gdmix_debug_fast_cycles(out, 1)
v = [out[k] / E for k in range(12)]
names = ["pop+stage", "B1->B2 rows", "B2->B3 cols", "B3->B4 dots", "B4->mxm decide+pair", "mxm (warp 0)", "mxm->B5", "B5->B1 direction",
         "emit+loop top"]
tot = sum(v[:9])
for n, x in zip(names, v[:9]):
    print(f"{n:22s} {x:10.0f} cycles/entity  {100 * x / tot:5.1f}%")
print(f"total {tot:.0f} cycles/entity, iterations with an update {v[9]:.2f}/entity, mxm {v[5] / max(v[9], 1e-9):.0f} cycles each")
