#!/usr/bin/env python
"""gdmix_re_fit_host / gdmix_re_score_host on 100 K C1 entities held in page-locked arrays (what the plugin path hands
them): seconds per call, and the C call alone.  Usage: python tools/score_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gdmix_b200 import _capi as capi
from gdmix_b200.synth_arrays import make_arrays
from gdmix_b200.synthetic import make_batch
hb = make_batch(100000, 128, 256, 32, seed=3)
# move the big arrays into pinned memory as the plugin path has them
def pin(a):
    out = capi.pinned_empty(a.shape[0], a.dtype); out[:] = a; return out
for name in ("val", "label", "offset", "rowptr"):
    if getattr(hb, name) is not None: setattr(hb, name, pin(getattr(hb, name)))
opts = capi.make_opts(l2=1.0)
for rep in range(3):
    t = time.perf_counter(); fit = capi.re_fit_host(hb, opts); t1 = time.perf_counter()
    lg = capi.re_score_host(hb, opts, fit["theta"], None); t2 = time.perf_counter()
    print(f"fit {t1-t:.3f} s  score {t2-t1:.3f} s")
import ctypes as C
# inside score: time the C call alone
cb = hb.c_struct(); th = fit["theta"]
logit = capi.pinned_empty(hb.n_rows, np.float32); per = capi.pinned_empty(hb.n_rows, np.float32)
for rep in range(3):
    t = time.perf_counter()
    capi.check(capi.lib.gdmix_re_score_host(C.byref(cb), C.byref(opts), capi._np_ptr(th), None, capi._np_ptr(logit), capi._np_ptr(per)))
    print(f"C call {time.perf_counter()-t:.3f} s")
