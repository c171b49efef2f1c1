#!/usr/bin/env python
"""Throughput of the partitioner kernels on the C3 shape: n rows keyed by entity id (power-law sizes), radix
sort + group-by, AUC.  GB/s against the bytes each must move at least once (sort: 12 B in + 12 B out per pass)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gdmix_b200 import partition as P

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64_000_000
g = torch.Generator(device="cuda"); g.manual_seed(0)
ent = (torch.rand(n, device="cuda", generator=g) ** 3 * 10_000_000).to(torch.int64)
for _ in range(2):
    P.group_by_entity(ent, key_bits=24)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    perm, seg, key = P.group_by_entity(ent, key_bits=24)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(json.dumps({"op": "group_by_entity (3 radix passes + segments)", "rows": n, "entities": int(key.numel()), "ms": ms,
                  "rows_per_s": n / ms * 1e3, "algorithmic_GBps": 3 * 24 * n / ms / 1e6}))
s = torch.randn(n, device="cuda", generator=g)
y = (torch.rand(n, device="cuda", generator=g) < torch.sigmoid(s)).float()
P.auc(s, y)
torch.cuda.synchronize()
e0.record()
for _ in range(3):
    a = P.auc(s, y)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(json.dumps({"op": "auc (4 radix passes + tie groups)", "rows": n, "auc": a, "ms": ms, "rows_per_s": n / ms * 1e3,
                  "algorithmic_GBps": (4 * 24 + 8) * n / ms / 1e6}))
