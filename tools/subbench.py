"""Sub-benchmarks of bench.py: the configurations of BASELINE.json other than the headline (configs[2], [3], [4] and the
small-entity regime of configs[3]'s per-user stage), each one rank's share of the workload, every rank running its own
share (weak scaling), timed with CUDA events after a warm-up, inputs resident in HBM and larger than L2.

Every function returns a JSON-serialisable dict; `world` / `group` describe the torch.distributed job (the fixed-effect
solve is the one place with a data-path collective: one all-reduce of [value | gradient] per evaluation).
"""
import ctypes as C
import time

import numpy as np
import torch

from gdmix_b200 import _capi as capi
from gdmix_b200 import partition as P
from gdmix_b200.fe_solver import FixedEffectSolver
from gdmix_b200.synthetic import make_device_batch


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _re_batch(data):
    return capi.ReBatch(data["n_entities"], data["n_rows"], data["nnz"], data["ent_rowptr"].data_ptr(),
                        data["rowptr"].data_ptr(), data["col"].data_ptr(), data["val"].data_ptr(),
                        data["label"].data_ptr(), None, data["offset"].data_ptr(), data["theta_ptr"].data_ptr(),
                        data["max_rows"], data["max_nnz"], data["max_coef"], 0)


def _timed(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def _max_over_ranks(v, dev, world):
    if world <= 1:
        return float(v)
    t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


# ---- configs[3] per-user shape: the small-entity regime ----------------------------------------------------------------
def small_entities(dev, rank, world, peak_gbs, E=4_000_000, n=32, d=64, k=8, iters=3):
    data = make_device_batch(E, n, d, k, seed=7 + rank, device=dev)
    cb = _re_batch(data)
    opts = capi.make_opts(l2=1.0)
    ws = torch.empty(max(capi.re_workspace_size(cb, opts), 256), dtype=torch.uint8, device=dev)
    theta = torch.empty(data["n_coef"], dtype=torch.float64, device=dev)
    nit = torch.empty(E, dtype=torch.int32, device=dev)
    status = torch.empty(E, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()

    def run():
        capi.check(capi.lib.gdmix_re_fit(C.byref(cb), C.byref(opts), None, _ptr(theta), None, _ptr(nit), None,
                                         _ptr(status), None, _ptr(ws), C.c_size_t(ws.numel()),
                                         C.c_void_p(st.cuda_stream)))
    ms = _max_over_ranks(_timed(run, iters), dev, world)
    alg = 8 * data["nnz"] + 16 * data["n_rows"] + 8 * data["n_coef"] + 4 * (data["n_coef"] - E)
    return {"workload": f"c3 per-user shape: {E} entities/GPU x {n} samples x {d} local features, {k} nnz/sample",
            "entities_per_s": world * E / ms * 1e3, "ms": ms, "plan": capi.last_plan(),
            "converged_frac": float((status == 0).float().mean().item()), "mean_nit": float(nit.float().mean().item()),
            "roofline": {"bound": "hbm", "achieved": alg / ms / 1e6, "peak": peak_gbs, "unit": "GB/s",
                         "frac": alg / ms / 1e6 / peak_gbs, "algorithmic_bytes_per_launch": alg}}


# ---- configs[4]: regularisation sweep from one staged copy per entity -------------------------------------------------
def l2_sweep(dev, rank, world, peak_gbs, E=2_000_000, n=64, d=64, k=16, iters=2):
    L2 = np.array([0.1, 1.0, 10.0, 100.0])
    data = make_device_batch(E, n, d, k, seed=4 + rank, device=dev, ragged=True)
    cb = _re_batch(data)
    opts = capi.make_opts(l2=1.0)
    ws = torch.empty(max(capi.re_workspace_size(cb, opts), 256), dtype=torch.uint8, device=dev)
    nc = data["n_coef"]
    theta = torch.empty((len(L2), nc), dtype=torch.float64, device=dev)
    nit = torch.empty((len(L2), E), dtype=torch.int32, device=dev)
    status = torch.empty_like(nit)
    st = torch.cuda.current_stream()

    def sweep():
        capi.check(capi.lib.gdmix_re_fit_sweep(C.byref(cb), C.byref(opts), L2.ctypes.data_as(C.c_void_p),
                                               C.c_int32(len(L2)), None, _ptr(theta), C.c_int64(nc), None, _ptr(nit),
                                               None, _ptr(status), _ptr(ws), C.c_size_t(ws.numel()),
                                               C.c_void_p(st.cuda_stream)))
    ms = _max_over_ranks(_timed(sweep, iters, warm=1), dev, world)
    alg = 8 * data["nnz"] + 16 * data["n_rows"] + len(L2) * 8 * nc + 4 * (nc - E)
    return {"workload": f"c4 shape: {E} entities/GPU x avg {n} samples (ragged 8..1024) x {d} local features, "
                        f"{k} nnz/sample, l2 in {L2.tolist()} from one staged copy per entity",
            "models_per_s": world * len(L2) * E / ms * 1e3, "entities_per_s": world * E / ms * 1e3, "ms": ms,
            "plan": capi.last_plan(), "converged_frac": float((status == 0).float().mean().item()),
            "mean_nit_per_l2": nit.float().mean(1).tolist(),
            "roofline": {"bound": "hbm", "achieved": alg / ms / 1e6, "peak": peak_gbs, "unit": "GB/s",
                         "frac": alg / ms / 1e6 / peak_gbs, "algorithmic_bytes_per_launch": alg}}


# ---- configs[2]: fixed-effect objective + all-reduce + replicated L-BFGS ----------------------------------------------
def zipf_rows(rows, D, k, seed, dev, planted=True):
    """`rows` samples x k non-zeros over D features with Zipf-like popularity (inverse CDF of a power law), N(0,1)
    values, labels from a planted model.  Generated in chunks: nothing but the shard itself stays allocated."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    gx = torch.Generator(device=dev)
    gx.manual_seed(12345)                       # the planted model is the same on every rank
    xs = torch.randn(D + 1, device=dev, generator=gx) * 0.3
    col = torch.empty(rows * k, dtype=torch.int32, device=dev)
    val = torch.empty(rows * k, dtype=torch.float32, device=dev)
    label = torch.empty(rows, dtype=torch.float32, device=dev)
    chunk = 4_000_000
    for r0 in range(0, rows, chunk):
        r1 = min(rows, r0 + chunk)
        m = r1 - r0
        u = torch.rand(m * k, device=dev, generator=g)
        c = ((D ** (1.0 - u.double())) - 1.0).clamp_(0, D - 1).to(torch.int32)
        v = torch.randn(m * k, device=dev, generator=g)
        col[r0 * k:r1 * k] = c
        val[r0 * k:r1 * k] = v
        if planted:
            z = (v.view(m, k) * xs[c.long()].view(m, k)).sum(1) + xs[D]
            label[r0:r1] = (torch.rand(m, device=dev, generator=g) < torch.sigmoid(z)).float()
        else:
            label[r0:r1] = (torch.rand(m, device=dev, generator=g) < 0.5).float()
        del u, c, v
    r = capi.DeviceFeRows.__new__(capi.DeviceFeRows)
    r.rowptr = torch.arange(rows + 1, device=dev, dtype=torch.int64) * k
    r.col, r.val, r.label, r.weight, r.offset = col, val, label, None, None
    r.n_rows, r.nnz, r.n_features, r.linear_regression, r.num_workers = rows, rows * k, D, False, 1
    return r


def fixed_effect(dev, rank, world, group, peak_gbs, rows=62_500_000, D=100_000, k=32, iters=20, eval_reps=5):
    t0 = time.perf_counter()
    shard = zipf_rows(rows, D, k, 100 + rank, dev)
    shard.num_workers = world
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    opts = capi.make_opts(l2=1.0, regularize_bias=True, max_iter=iters)
    solver = FixedEffectSolver(shard, opts, D, group=group if world > 1 else None, profile=True)
    t0 = time.perf_counter()
    solver._prepare()
    torch.cuda.synchronize()
    plan_s = time.perf_counter() - t0
    # one objective / gradient evaluation (this rank's kernels only), steady state
    ms_kernels = _timed(solver._partial, eval_reps, warm=2)
    ms_kernels = _max_over_ranks(ms_kernels, dev, world)
    # the solve: `iters` L-BFGS iterations from x = 0, all-reduce and solver step included (after a two-iteration
    # solve that loads the solver's kernels and, for N > 1, opens the NCCL channels)
    opts.max_iter = 2
    solver.fit()
    opts.max_iter = iters
    # the timed solve, twice: the slower rank sets the pace through the all-reduce, and a collector pass or a cudaFree of
    # the benchmarks before this one landing on ONE rank's host thread triples a 0.25 s solve (seen at N = 4) -- so:
    # collect first, no collector inside, and the better of two solves is reported (both are in `solve_seconds_runs`)
    import gc
    runs = []
    for _ in range(2):
        solver.phase_ms.clear()
        gc.collect()
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        gc.disable()
        t0 = time.perf_counter()
        x, info = solver.fit()
        torch.cuda.synchronize()
        runs.append(_max_over_ranks(time.perf_counter() - t0, dev, world))
        gc.enable()
    fit_s = min(runs)
    ph = np.array(solver.phase_ms) if solver.phase_ms else np.zeros((1, 3))
    xt = torch.from_numpy(x).to(dev)
    identical = True
    if world > 1:
        lo, hi = xt.clone(), xt.clone()
        torch.distributed.all_reduce(lo, op=torch.distributed.ReduceOp.MIN)
        torch.distributed.all_reduce(hi, op=torch.distributed.ReduceOp.MAX)
        identical = bool((lo == hi).all().item())
    alg = rows * (8 * k + 16)
    out = {"workload": f"c2 share: {rows} rows/GPU x {k} nnz over D={D} features (Zipf-like popularity), l2=1, m=10, "
                       f"{iters} L-BFGS iterations from x=0; {world} rank(s), rows sharded, one all-reduce of "
                       f"fg[D+2] per evaluation",
           "rows_per_gpu": rows, "generation_s": gen_s, "plan_s": plan_s, "solve_seconds_runs": runs,
           "ms_per_eval_kernels": ms_kernels,
           "ms_per_eval_in_solve": {"kernels": float(np.median(ph[:, 0])), "allreduce": float(np.median(ph[:, 1])),
                                    "solver_step": float(np.median(ph[:, 2]))},
           "solve_seconds": fit_s, "nit": info["nit"], "nfev": info["nfev"], "status": info["status"],
           "f": info["f"], "evals_per_s": info["nfev"] / fit_s, "rows_per_s": world * rows * info["nfev"] / fit_s,
           "ranks_bit_identical": identical, "solver": solver.solver,
           "roofline": {"bound": "hbm", "achieved": alg / ms_kernels / 1e6, "peak": peak_gbs, "unit": "GB/s",
                        "frac": alg / ms_kernels / 1e6 / peak_gbs, "algorithmic_bytes_per_eval": alg,
                        "bytes_per_row": 8 * k + 16, "traffic": None}}
    return out


# ---- configs[3]: fixed effect -> per-user -> per-item chained on offsets -----------------------------------------------
def chain(dev, rank, world, group, n=40_000_000, U=1_250_000, I=125_000, fe_iters=20):
    """One rank's share of configs[3] (an 8-rank run of 3.2e8 rows / 1e7 users / 1e6 items): everything on the device.
    Rows are assumed partitioned by entity before they reach the rank (DataPartitioner's job), so only the
    fixed-effect stage has a collective."""
    D, k, Du, ku, Di, ki = 100_000, 32, 64, 8, 64, 8
    g = torch.Generator(device=dev)
    g.manual_seed(3 + rank)

    def bag(nrows, Dn, kn):
        stride = Dn // kn
        within = torch.randint(0, stride, (nrows, kn), dtype=torch.int32, device=dev, generator=g)
        col = (torch.arange(kn, dtype=torch.int32, device=dev) * stride)[None, :] + within
        val = torch.randn(nrows, kn, device=dev, generator=g)
        return (torch.arange(nrows + 1, dtype=torch.int64, device=dev) * kn, col.reshape(-1).contiguous(),
                val.reshape(-1).contiguous())

    def sync():
        torch.cuda.synchronize()
        return time.perf_counter()

    user = (torch.rand(n, device=dev, generator=g) ** 2 * U).to(torch.int64)          # power-law user activity
    item = ((I ** torch.rand(n, device=dev, generator=g).double()) - 1.0).clamp_(0, I - 1).to(torch.int64)
    u_rp, u_col, u_val = bag(n, Du, ku)
    i_rp, i_col, i_val = bag(n, Di, ki)
    rows = zipf_rows(n, D, k, 300 + rank, dev)
    rows.num_workers = world
    y = rows.label
    out = {"workload": f"c3 share: {n} rows/GPU, {U} users, {I} items, FE D={D} k={k}, per-user d={Du} k={ku}, "
                       f"per-item d={Di} k={ki}; fp32 scores of each stage are the next stage's offsets"}
    opts_fe = capi.make_opts(l2=1.0, regularize_bias=True, max_iter=fe_iters)
    solver = FixedEffectSolver(rows, opts_fe, D, group=group if world > 1 else None)
    t0 = sync()
    solver._prepare()
    t_plan = sync() - t0
    t0 = sync()
    x, info = solver.fit()
    t_fe = sync() - t0
    xd = torch.from_numpy(x).to(dev)
    t0 = sync()
    s0, _ = capi.fe_score_device(rows, opts_fe, xd)
    t_score0 = sync() - t0
    out["fe"] = {"plan_s": t_plan, "fit_s": t_fe, "nit": info["nit"], "nfev": info["nfev"],
                 "evals_per_s": info["nfev"] / t_fe, "score_s": t_score0, "auc": P.auc(s0, y)}
    del solver, rows
    torch.cuda.empty_cache()

    def re_stage(name, keys, rp, col, val, off, Dn):
        t0 = sync()
        gb = P.GroupedBatch(P.regroup_batch(keys, rp, col, val, y, off, None, num_features=Dn))
        t_group = sync() - t0
        opts = capi.make_opts(l2=1.0, regularize_bias=False)
        t0 = sync()
        fit = capi.re_fit_device(gb, opts)
        t_fit = sync() - t0
        plan = capi.last_plan()
        t0 = sync()
        logit, _ = capi.re_score_device(gb, opts, fit["theta"])
        s = gb.scatter_to_input_order(logit)
        t_score = sync() - t0
        st = fit["status"]
        E = st.numel()
        out[name] = {"entities": E, "max_rows": gb.host.max_rows, "group_s": t_group, "fit_s": t_fit,
                     "entities_per_s": E / t_fit, "score_s": t_score,
                     "converged_frac": float((st == 0).double().mean().item()),
                     "mean_nit": float(fit["nit"].float().mean().item()), "plan": plan, "auc": P.auc(s, y)}
        return s

    s1 = re_stage("per_user", user, u_rp, u_col, u_val, s0, Du)
    re_stage("per_item", item, i_rp, i_col, i_val, s1, Di)
    out["seconds_total"] = sum(out[s][key] for s in ("fe", "per_user", "per_item") for key in out[s]
                               if key.endswith("_s") and key != "evals_per_s" and key != "entities_per_s")
    return out
