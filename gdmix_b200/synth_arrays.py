"""Pure-numpy generator of the synthetic random-effect workloads -- no package imports, so that the
CPU arm of bench.py can load this file by path without importing gdmix_b200 (which loads the CUDA library).
gdmix_b200.synthetic.make_batch wraps make_arrays() into a HostBatch."""
import numpy as np


def _sample_counts(rng, E, n_mean, ragged, n_min=8, n_max=1024):
    if not ragged:
        return np.full(E, n_mean, np.int64)
    raw = rng.lognormal(mean=0.0, sigma=0.5, size=E)
    raw *= n_mean / np.exp(0.125)  # E[lognormal(0, .5)] = exp(.125)
    return np.clip(np.rint(raw), min(n_min, n_mean), n_max).astype(np.int64)


def make_arrays(E, n_mean=128, d=256, k=32, seed=20240601, ragged=False, weights=False, has_intercept=True):
    """-> (dict of numpy arrays in the gdmix_re_batch layout, theta_star).  Deterministic in (E, shape, seed)."""
    rng = np.random.default_rng(seed)
    k = min(k, d)
    n_e = _sample_counts(rng, E, n_mean, ragged)
    ent_rowptr = np.zeros(E + 1, np.int64)
    np.cumsum(n_e, out=ent_rowptr[1:])
    N = int(ent_rowptr[-1])
    rowptr = np.arange(N + 1, dtype=np.int64) * k
    stride = d // k
    col = (np.arange(k, dtype=np.int32)[None, :] * stride +
           rng.integers(0, stride, size=(N, k), dtype=np.int32)).reshape(-1)
    val = rng.standard_normal(N * k, dtype=np.float32)
    off = rng.standard_normal(N, dtype=np.float32)
    theta_star = (0.5 * rng.standard_normal((E, d + 1))).astype(np.float64)
    ent_of_row = np.repeat(np.arange(E), n_e)
    z = (val.reshape(N, k).astype(np.float64) *
         theta_star[ent_of_row[:, None], 1 + col.reshape(N, k)]).sum(axis=1) + theta_star[ent_of_row, 0] + off
    y = (rng.random(N) < 1.0 / (1.0 + np.exp(-z))).astype(np.float32)
    w = rng.uniform(0.5, 2.0, N).astype(np.float32) if weights else None
    p = d + (1 if has_intercept else 0)
    theta_ptr = np.arange(E + 1, dtype=np.int64) * p
    return {"ent_rowptr": ent_rowptr, "rowptr": rowptr, "col": col, "val": val, "label": y, "weight": w,
            "offset": off, "theta_ptr": theta_ptr}, theta_star


