"""Seeded synthetic random-effect workloads with the shapes BASELINE.json names.

Every entity gets n_e samples with k non-zeros drawn from d entity-local features (one per stratum of
d/k columns, so columns are sorted and unique inside a row), fp32 N(0,1) values, an fp32 N(0,1) offset,
and labels drawn from a planted per-entity model theta* ~ N(0, 0.5^2).  The numpy generator feeds the
parity tests (identical bytes go to the oracle and to the GPU); the torch generator builds the full-size
bench workloads directly in HBM.
"""
import numpy as np

from ._capi import HostBatch

CONFIGS = {
    # name: (avg samples, local features, nnz per sample)
    "c1": (128, 256, 32),      # BASELINE.json configs[1]: 1M entities, avg 128 samples x 256 sparse features
    "c3_user": (32, 64, 8),    # configs[3] per-user stage
    "c4": (64, 64, 16),        # configs[4]
    "movielens_user": (106, 20, 6),  # configs[0] look-alike marginals
}


from .synth_arrays import _sample_counts, make_arrays  # noqa: F401


def make_batch(E, n_mean=128, d=256, k=32, seed=20240601, ragged=False, weights=False, has_intercept=True,
               return_truth=False):
    """-> HostBatch (numpy).  Deterministic in (E, shape, seed)."""
    a, theta_star = make_arrays(E, n_mean, d, k, seed, ragged, weights, has_intercept)
    hb = HostBatch(a["ent_rowptr"], a["rowptr"], a["col"], a["val"], a["label"], a["weight"], a["offset"],
                   a["theta_ptr"], has_intercept)
    return (hb, theta_star) if return_truth else hb


def slice_batch(hb, e0, e1):
    """Entities [e0, e1) of a HostBatch as a new, rebased HostBatch."""
    r0, r1 = int(hb.ent_rowptr[e0]), int(hb.ent_rowptr[e1])
    q0, q1 = int(hb.rowptr[r0]), int(hb.rowptr[r1])
    return HostBatch(hb.ent_rowptr[e0:e1 + 1] - r0, hb.rowptr[r0:r1 + 1] - q0, hb.col[q0:q1], hb.val[q0:q1],
                     hb.label[r0:r1], None if hb.weight is None else hb.weight[r0:r1],
                     None if hb.offset is None else hb.offset[r0:r1], hb.theta_ptr[e0:e1 + 1] - hb.theta_ptr[e0])


def make_device_batch(E, n=128, d=256, k=32, seed=20240601, device="cuda", ragged=False, has_intercept=True):
    """Builds the workload in HBM with torch (uniform n unless ragged).  Returns a dict of CUDA tensors plus the
    scalar bounds gdmix_re_batch needs; nothing of size O(nnz) ever touches the host."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    k = min(k, d)
    if ragged:
        raw = torch.empty(E, device=device).log_normal_(0.0, 0.5, generator=g) * (n / float(np.exp(0.125)))
        n_e = raw.round().clamp_(min(8, n), 1024).to(torch.int64)
    else:
        n_e = torch.full((E,), n, dtype=torch.int64, device=device)
    ent_rowptr = torch.zeros(E + 1, dtype=torch.int64, device=device)
    ent_rowptr[1:] = torch.cumsum(n_e, 0)
    N = int(ent_rowptr[-1].item())
    rowptr = torch.arange(N + 1, dtype=torch.int64, device=device) * k
    stride = d // k
    p = d + (1 if has_intercept else 0)
    col = torch.empty(N * k, dtype=torch.int32, device=device)
    val = torch.empty(N * k, dtype=torch.float32, device=device)
    off = torch.empty(N, dtype=torch.float32, device=device)
    y = torch.empty(N, dtype=torch.float32, device=device)
    theta_star = torch.empty(E, d + 1, dtype=torch.float32, device=device).normal_(0.0, 0.5, generator=g)
    ent_of_row = torch.repeat_interleave(torch.arange(E, device=device), n_e)
    base = (torch.arange(k, dtype=torch.int32, device=device) * stride)[None, :]
    chunk = max(1, (64 << 20) // max(k, 1))  # rows per generation chunk: bounds temporaries to ~1 GB
    for r0 in range(0, N, chunk):
        r1 = min(N, r0 + chunk)
        m = r1 - r0
        c = base + torch.randint(0, stride, (m, k), dtype=torch.int32, device=device, generator=g)
        v = torch.empty(m, k, dtype=torch.float32, device=device).normal_(generator=g)
        o = torch.empty(m, dtype=torch.float32, device=device).normal_(generator=g)
        ents = ent_of_row[r0:r1]
        th = theta_star[ents[:, None], (1 + c).long()]
        z = (v * th).sum(1) + theta_star[ents, 0] + o
        yy = (torch.rand(m, device=device, generator=g) < torch.sigmoid(z)).float()
        col[r0 * k:r1 * k] = c.reshape(-1)
        val[r0 * k:r1 * k] = v.reshape(-1)
        off[r0:r1] = o
        y[r0:r1] = yy
        del c, v, o, th, z, yy
    theta_ptr = torch.arange(E + 1, dtype=torch.int64, device=device) * p
    return {"ent_rowptr": ent_rowptr, "rowptr": rowptr, "col": col, "val": val, "label": y, "offset": off,
            "weight": None, "theta_ptr": theta_ptr, "n_entities": E, "n_rows": N, "nnz": N * k,
            "max_rows": int(n_e.max().item()), "max_nnz": int(n_e.max().item()) * k, "max_coef": p, "n_coef": E * p}
