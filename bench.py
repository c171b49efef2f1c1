#!/usr/bin/env python
"""bench.py -- random-effect entities converged/sec on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--entities E]

Ours: a "step" is one pass of the hot path (gdmix_re_fit: stage + L-BFGS-B solve of every entity) over the
workload BASELINE.json configs[1] names -- 1M synthetic entities x 128 samples x 256 local features, 32 nnz per
sample, l2=1, cold start -- resident in HBM (38 GB, far larger than L2).  `value` = entities solved per second
with device-resident inputs, timed with CUDA events on the launching stream, max over ranks.  `e2e` = the same
metric through the host-buffer C-ABI call the plugin classes make (gdmix_re_fit_host): pinned host CSR in,
coefficients out, H2D and D2H inside the timed region.  For N>1 every rank solves its own 1M entities (entities
shard with no data-path collective: weak scaling); launched by torch.distributed.run, NCCL only for the barrier
and the max-over-ranks.

Reference arm (--impl reference): the reference's CPU path restated with the same scipy/numpy calls
(oracle/scipy_port.py; /root/reference cannot travel to the GPU box) on all host cores, on a bounded sample of
the same workload per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOAD = dict(name="c1", n=128, d=256, k=32, l2=1.0, seed=20240601)
METRIC = "re_entities_converged_per_sec"
UNIT = "entities/s"


def _peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def _oracle_batch_dict(hb):
    return {"ent_rowptr": hb.ent_rowptr, "rowptr": hb.rowptr, "col": hb.col, "val": hb.val, "y": hb.label,
            "w": hb.weight if hb.weight is not None else np.ones(hb.n_rows, np.float32),
            "off": hb.offset if hb.offset is not None else np.zeros(hb.n_rows, np.float32), "theta_ptr": hb.theta_ptr}


def cpu_baseline(sample_entities):
    """Reference CPU path (scipy port) on all host cores over a bounded sample of the workload, plus the plain-C
    oracle port on one core for scale."""
    from gdmix_b200.synthetic import make_batch
    from oracle import oracle as O
    from oracle import scipy_port as SP
    kw = dict(l2=WORKLOAD["l2"], regularize_bias=False, has_intercept=True)
    if sample_entities <= 0:
        # size the sample for ~15 s of CPU work on this box: probe the rate on a small slice first
        probe = make_batch(64 * (os.cpu_count() or 1), WORKLOAD["n"], WORKLOAD["d"], WORKLOAD["k"],
                           seed=WORKLOAD["seed"])
        rate = SP.timed_fit(_oracle_batch_dict(probe), probe.n_entities, **kw)["entities_per_sec"]
        sample_entities = int(min(max(rate * 15.0, 2048), 400000))
    hb = make_batch(sample_entities, WORKLOAD["n"], WORKLOAD["d"], WORKLOAD["k"], seed=WORKLOAD["seed"])
    b = _oracle_batch_dict(hb)
    r = SP.timed_fit(b, sample_entities, **kw)
    nc = min(sample_entities, 2000)
    t0 = time.perf_counter()
    O.re_fit_batch(b, O.make_opts(l2=WORKLOAD["l2"]), e0=0, e1=nc)
    c1 = nc / (time.perf_counter() - t0)
    return {"value": r["entities_per_sec"], "unit": UNIT, "cores": r["cores"], "kind": "port",
            "sample": f"first {sample_entities} entities of the c1 workload (identical generator/seed), scipy "
                      f"{__import__('scipy').__version__} fmin_l_bfgs_b + scipy.sparse loss/gradient restated from "
                      "the reference (oracle/scipy_port.py), multiprocessing over all cores, TF reader / queue / "
                      "Avro excluded",
            "seconds": r["seconds"], "mean_nit": r["mean_nit"],
            "c_oracle_1core_entities_per_sec": c1}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gdmix_b200.synthetic import make_batch
    from oracle import scipy_port as SP
    cores = os.cpu_count() or 1
    kw = dict(l2=WORKLOAD["l2"], regularize_bias=False, has_intercept=True)
    probe = make_batch(64 * cores, WORKLOAD["n"], WORKLOAD["d"], WORKLOAD["k"], seed=WORKLOAD["seed"])
    rate = SP.timed_fit(_oracle_batch_dict(probe), probe.n_entities, **kw)["entities_per_sec"]
    per_step = int(min(max(rate * 6.0, 1024), 200000))  # ~6 s of CPU work per step
    hb = make_batch(per_step, WORKLOAD["n"], WORKLOAD["d"], WORKLOAD["k"], seed=WORKLOAD["seed"])
    b = _oracle_batch_dict(hb)
    for _ in range(args.warmup):
        SP.timed_fit(b, min(per_step, 4 * cores), **kw)
    secs, done = 0.0, 0
    for _ in range(args.steps):
        r = SP.timed_fit(b, per_step, **kw)
        secs += r["seconds"]; done += r["entities"]
    v = done / secs
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.entities), "entities_per_gpu": args.entities,
                       "sample_entities_per_step": per_step,
                       "note": "each step solves a bounded sample of the same workload (same generator and seed)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{per_step} c1 entities per step through oracle/scipy_port.py "
                                       "(reference call sequence on scipy/numpy), all host cores"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_name(E):
    w = WORKLOAD
    return (f"c1: {E} synthetic entities x {w['n']} samples x {w['d']} local features, {w['k']} nnz/sample, "
            f"l2={w['l2']}, bias unregularised, m=10, tol=1e-12, cold start (BASELINE.json configs[1])")


def run_ours(args):
    import torch
    import torch.distributed as dist
    from gdmix_b200 import _capi as capi
    from gdmix_b200.synthetic import make_device_batch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner there at any debug level
        os.environ.pop("NCCL_DEBUG", None)
        if os.environ.get("GDMIX_NCCL_DEBUG"):
            os.environ["NCCL_DEBUG"] = os.environ["GDMIX_NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    E = args.entities
    w = WORKLOAD
    t0 = time.time()
    data = make_device_batch(E, w["n"], w["d"], w["k"], seed=w["seed"] + rank, device=dev)
    torch.cuda.synchronize()
    gen_s = time.time() - t0

    cb = capi.ReBatch(E, data["n_rows"], data["nnz"], data["ent_rowptr"].data_ptr(), data["rowptr"].data_ptr(),
                      data["col"].data_ptr(), data["val"].data_ptr(), data["label"].data_ptr(), None,
                      data["offset"].data_ptr(), data["theta_ptr"].data_ptr(), data["max_rows"], data["max_nnz"],
                      data["max_coef"], 0)
    opts = capi.make_opts(l2=w["l2"], regularize_bias=False, has_intercept=True,
                          threads_per_entity=args.threads_per_entity)
    ws = torch.empty(max(capi.re_workspace_size(cb, opts), 256), dtype=torch.uint8, device=dev)
    theta = torch.empty(data["n_coef"], dtype=torch.float64, device=dev)
    f = torch.empty(E, dtype=torch.float64, device=dev)
    nit = torch.empty(E, dtype=torch.int32, device=dev)
    nfev = torch.empty(E, dtype=torch.int32, device=dev)
    status = torch.empty(E, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    import ctypes as C

    def step():
        capi.check(capi.lib.gdmix_re_fit(C.byref(cb), C.byref(opts), None, C.c_void_p(theta.data_ptr()),
                                         C.c_void_p(f.data_ptr()), C.c_void_p(nit.data_ptr()),
                                         C.c_void_p(nfev.data_ptr()), C.c_void_p(status.data_ptr()), None,
                                         C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel()),
                                         C.c_void_p(stream.cuda_stream)))

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = capi.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record(stream)
    for i in range(args.steps):
        step()
        ev[i + 1].record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = capi.launch_count() - launches0
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[-1])
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * E * args.steps / (total_ms_max / 1e3)

    # sanity of the solve itself (not timed): all converged, iteration statistics
    st = status.cpu().numpy()
    nit_h = nit.cpu().numpy()
    nfev_h = nfev.cpu().numpy()
    converged = float((st == 0).mean())

    # roofline of the dominant (only) kernel: algorithmic bytes per launch / average launch duration
    alg_bytes = 8 * data["nnz"] + 16 * data["n_rows"] + 8 * data["n_coef"] + 4 * (data["n_coef"] - E)
    kern_s = (sum(step_ms) / len(step_ms)) / 1e3
    peak, peak_src = _peaks()
    achieved = alg_bytes / kern_s / 1e9
    streaming_bytes = float(nfev_h.astype(np.float64).sum()) * (8 * w["n"] * w["k"] + 16 * w["n"])
    # fp64 work actually issued per entity: 2 sparse matvecs per evaluation + two-loop per iteration
    flops = float((nfev_h.astype(np.float64) * (4.0 * w["n"] * w["k"])).sum() +
                  (nit_h.astype(np.float64) * (8.0 * 10 * (w["d"] + 1))).sum())
    plan = capi.last_plan()
    kernel = (f"re_fast_kernel<{plan['threads']},{plan['ept']}>" if plan["fast"] else
              f"re_solver_kernel<{plan['threads']}>")
    # DRAM bytes per entity of this kernel on this workload from the committed ncu --set full capture
    # (profiles/r1_ncu_full_re_fast_v8.txt: dram__bytes_read.sum + dram__bytes_write.sum over 30 000 entities)
    ncu_dram_bytes_per_entity = (1.047181e9 + 87.874816e6) / 30000.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_dram_bytes_per_entity * E if w["name"] == "c1" else None,
                "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per entity "
                                  "(profiles/r1_ncu_full_re_fast_v8.txt) x entities per launch",
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel": kernel, "plan": plan,
                "kernel_ms_per_launch": kern_s * 1e3,
                "streaming_model_gbs": streaming_bytes / kern_s / 1e9,
                "fp64_gflops": flops / kern_s / 1e9,
                "note": "each entity is read from HBM once and solved on chip (about 15.8 f/g evaluations and 14.4 "
                        "L-BFGS updates out of shared memory and registers), so the kernel is bound by fp64 issue "
                        "latency and the shared-memory pipe, not by HBM (DESIGN.md section 4)"}

    # ---- e2e: host CSR in pinned memory -> gdmix_re_fit_host -> coefficients on the host -------------------
    e2e = None
    Ee = min(args.e2e_entities, E)
    if world >= 4:
        Ee = min(Ee, 262144)   # keeps the pinned host staging of an 8-rank run under 50 GB in total
    rows_e, nnz_e, coef_e = Ee * w["n"], Ee * w["n"] * w["k"], Ee * (w["d"] + 1)

    def pinned(t):
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t)
        return h
    h_ent = pinned(data["ent_rowptr"][:Ee + 1]); h_row = pinned(data["rowptr"][:rows_e + 1])
    # local column indices cross PCIe as one byte each (every entity has 256 local features: gdmix_re_batch.col8)
    # and are widened on the device
    h_col = pinned(data["col"][:nnz_e].to(torch.uint8)); h_val = pinned(data["val"][:nnz_e])
    h_lab = pinned(data["label"][:rows_e]); h_off = pinned(data["offset"][:rows_e])
    h_tp = pinned(data["theta_ptr"][:Ee + 1])
    h_theta = torch.empty(coef_e, dtype=torch.float64, pin_memory=True)
    h_f = torch.empty(Ee, dtype=torch.float64, pin_memory=True)
    h_nit = torch.empty(Ee, dtype=torch.int32, pin_memory=True)
    h_nfev = torch.empty(Ee, dtype=torch.int32, pin_memory=True)
    h_st = torch.empty(Ee, dtype=torch.int32, pin_memory=True)
    hcb = capi.ReBatch(Ee, rows_e, nnz_e, h_ent.data_ptr(), h_row.data_ptr(), None, h_val.data_ptr(),
                       h_lab.data_ptr(), None, h_off.data_ptr(), h_tp.data_ptr(), w["n"], w["n"] * w["k"],
                       w["d"] + 1, 0, None, h_col.data_ptr())

    def e2e_step():
        capi.check(capi.lib.gdmix_re_fit_host(C.byref(hcb), C.byref(opts), None, C.c_void_p(h_theta.data_ptr()),
                                              C.c_void_p(h_f.data_ptr()), C.c_void_p(h_nit.data_ptr()),
                                              C.c_void_p(h_nfev.data_ptr()), C.c_void_p(h_st.data_ptr()), None,
                                              C.c_int64(args.e2e_chunk)))
    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    barrier()
    t1 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        e2e_step()  # synchronous: returns with the results in host memory
    torch.cuda.synchronize()
    dt = time.perf_counter() - t1
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    h2d = 8 * (Ee + 1) * 2 + 8 * (rows_e + 1) + 5 * nnz_e + 8 * rows_e
    d2h = 8 * coef_e + 8 * Ee + 12 * Ee
    e2e = {"value": world * Ee * e2e_steps / float(t.item()), "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "entities_per_step": Ee, "steps": e2e_steps,
           "api": "gdmix_re_fit_host (pinned host CSR in -- fp32 values, uint8 local columns -- host coefficients out)",
           "host_theta_checksum": float(h_theta.sum().item())}
    capi.lib.gdmix_host_release()

    if rank == 0:
        cpu = cpu_baseline(args.cpu_sample) if (world == 1 and not args.no_cpu_baseline) else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(E), "entities_per_gpu": E,
                           "l2_cache": "inputs (38 KB/entity x E) larger than L2 for E >= 4096; no flush needed",
                           "threads_per_entity": args.threads_per_entity or "auto",
                           "parallelism": f"entities sharded over {world} GPU(s), no data-path collective",
                           "generation_seconds": gen_s},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "solve": {"converged_frac": converged, "mean_nit": float(nit_h.mean()),
                          "mean_nfev": float(nfev_h.mean()), "max_nit": int(nit_h.max())}}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--entities", type=int, default=1_000_000, help="entities per GPU")
    ap.add_argument("--e2e-entities", type=int, default=524288)
    ap.add_argument("--e2e-chunk", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=0, help="entities for the cpu_baseline leg (0 = ~15 s of work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--threads-per-entity", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
