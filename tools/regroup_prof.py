#!/usr/bin/env python
"""Where regroup_batch (group rows by entity + entity-local indexing) spends its time at the C3 per-user shape."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gdmix_b200 import partition as P
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(3)
n, U, Dn, kn = 40_000_000, 1_250_000, 64, 8
user = (torch.rand(n, device=dev, generator=g) ** 2 * U).to(torch.int64)
stride = Dn // kn
col = ((torch.arange(kn, dtype=torch.int32, device=dev) * stride)[None, :] + torch.randint(0, stride, (n, kn), dtype=torch.int32, device=dev, generator=g)).reshape(-1).contiguous()
val = torch.randn(n * kn, device=dev, generator=g)
rp = torch.arange(n + 1, dtype=torch.int64, device=dev) * kn
y = torch.rand(n, device=dev, generator=g)
def T(f, name):
    torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize(); print(f"{name:34s} {1e3*(time.perf_counter()-t):8.2f} ms"); return r
T(lambda: P.group_by_entity(user), "warm")
perm, ent_rowptr, ids = T(lambda: P.group_by_entity(user), "group_by_entity(40M keys)")
E = ids.numel()
rp2, gc, va = T(lambda: P.gather_rows(rp, col, val, perm), "gather_rows")
T(lambda: P.gather_f32(y, perm), "gather_f32")
ent_of_row = T(lambda: torch.repeat_interleave(torch.arange(E, device=dev), ent_rowptr[1:] - ent_rowptr[:-1]), "repeat_interleave rows")
ent_of_nnz = T(lambda: torch.repeat_interleave(ent_of_row, rp2[1:] - rp2[:-1]), "repeat_interleave nnz")
pair = T(lambda: (ent_of_nnz << 6) | gc.to(torch.int64), "pair keys")
pperm, pseg, pkey = T(lambda: P.group_by_entity(pair, key_bits=6 + 21), "group_by (320M pairs, 27 bits)")
T(lambda: torch.repeat_interleave(torch.arange(pkey.numel(), device=dev), pseg[1:] - pseg[:-1]), "repeat_interleave groups")
T(lambda: P.regroup_batch(user, rp, col, val, y, None, None, num_features=64), "regroup_batch total")
