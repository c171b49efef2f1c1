#!/usr/bin/env python
"""gdmix_re_fit throughput on an arbitrary synthetic entity shape (device-resident), for every threads-per-entity
choice next to the planner's own.  Usage: python tools/shape_bench.py E n d k [ragged]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from gdmix_b200 import _capi as capi
from gdmix_b200.synthetic import make_device_batch

E, n, d, k = (int(x) for x in sys.argv[1:5])
ragged = len(sys.argv) > 5 and sys.argv[5] == "ragged"
dev = torch.device("cuda", 0)
data = make_device_batch(E, n, d, k, seed=7, device=dev, ragged=ragged)
cb = capi.ReBatch(E, data["n_rows"], data["nnz"], data["ent_rowptr"].data_ptr(), data["rowptr"].data_ptr(),
                  data["col"].data_ptr(), data["val"].data_ptr(), data["label"].data_ptr(), None,
                  data["offset"].data_ptr(), data["theta_ptr"].data_ptr(), data["max_rows"], data["max_nnz"],
                  data["max_coef"], 0)
theta = torch.empty(data["n_coef"], dtype=torch.float64, device=dev)
nit = torch.empty(E, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream()
P = lambda t: C.c_void_p(t.data_ptr())
for tpe in (0, 32, 64, 128, 256):
    opts = capi.make_opts(l2=1.0, threads_per_entity=tpe)
    try:
        ws = torch.empty(max(capi.re_workspace_size(cb, opts), 256), dtype=torch.uint8, device=dev)
        run = lambda: capi.check(capi.lib.gdmix_re_fit(C.byref(cb), C.byref(opts), None, P(theta), None, P(nit), None, None,
                                                       None, P(ws), C.c_size_t(ws.numel()), C.c_void_p(st.cuda_stream)))
        run(); run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(json.dumps({"threads_per_entity": tpe or "auto", "ms": ms, "entities_per_s": E / ms * 1e3,
                          "mean_nit": float(nit.float().mean().item()), "plan": capi.last_plan()}))
    except capi.GdmixError as ex:
        print(json.dumps({"threads_per_entity": tpe, "error": str(ex)}))
