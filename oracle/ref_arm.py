"""CPU arm of bench.py (`--impl reference`, `cpu_baseline`) -- TEST INFRASTRUCTURE.

Times the reference's per-entity training path on the host cores: for every entity what TrainingJobConsumer.__call__
does (gdmix-trainer/src/gdmix/models/custom/scipy/job_consumers.py:36-63) -- a scipy COO matrix in entity-local
index space (job_consumers.py:247), BinaryLogisticRegressionTrainer.fit (binary_logistic_regression.py:191-239,
scipy fmin_l_bfgs_b), threshold_coefficients (util/model_utils.py:4-12) -- in a fork pool over all cores.

kind "reference": the reference's OWN class, staged unmodified under oracle/_ref by oracle/build_ref.py.
kind "port":      oracle/scipy_port.py, the same call sequence restated (when oracle/_ref is absent).
The TF reader, the Manager queue (two pickles per job) and the Avro writer of the reference are not included, so
either kind over-states the reference's throughput.  Nothing here imports gdmix_b200 or loads its library.
"""
import os
import time

import numpy as np

from . import build_ref

EPS = float(np.finfo(float).eps)
_STATE = {}


def kind():
    return "reference" if build_ref.available() else "port"


def _fit_range_reference(batch, e0, e1, kw):
    import scipy.sparse
    Trainer, threshold = _STATE["ref"]
    hi = 1 if kw.get("has_intercept", True) else 0
    tr = Trainer(lambda_l2=kw.get("l2", 1.0), precision=kw.get("tol", 1e-12) / EPS,
                 num_lbfgs_corrections=kw.get("m", 10), max_iter=kw.get("max_iter", 100),
                 regularize_bias=kw.get("regularize_bias", False), has_intercept=bool(hi))
    nit = 0
    out = []
    for e in range(e0, e1):
        r0, r1 = batch["ent_rowptr"][e], batch["ent_rowptr"][e + 1]
        q0, q1 = batch["rowptr"][r0], batch["rowptr"][r1]
        n, d = int(r1 - r0), int(batch["theta_ptr"][e + 1] - batch["theta_ptr"][e]) - hi
        rows = np.repeat(np.arange(n), np.diff(batch["rowptr"][r0:r1 + 1]))
        X = scipy.sparse.coo_matrix((batch["val"][q0:q1], (rows, batch["col"][q0:q1])), shape=(n, d))
        (theta, f, info), _ = tr.fit(X=X, y=batch["y"][r0:r1].astype(np.int64), weights=batch["w"][r0:r1],
                                     offsets=batch["off"][r0:r1], theta_initial=None, variance_mode=None)
        out.append(threshold(theta, 1e-4))
        nit += int(info["nit"])
    return e1 - e0, nit, out


def _init(batch):
    _STATE["batch"] = batch
    if build_ref.available():
        _STATE["ref"] = build_ref.load()
    try:   # one BLAS/OpenMP thread per worker: the pool already uses every core
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:  # pragma: no cover
        pass


def _job(args):
    e0, e1, kw = args
    if "ref" in _STATE:
        n, nit, _ = _fit_range_reference(_STATE["batch"], e0, e1, kw)
    else:
        from . import scipy_port
        n, nit, _ = scipy_port._fit_range((_STATE["batch"], e0, e1, kw))
    return n, nit


def fit_entities(batch, e0, e1, **kw):
    """Single-process: coefficient vectors of entities [e0, e1) (tests pin this against the golden vectors)."""
    _init(batch)
    if "ref" in _STATE:
        return _fit_range_reference(batch, e0, e1, kw)[2]
    from . import scipy_port
    return scipy_port._fit_range((batch, e0, e1, kw))[2]


def timed_fit(batch, n_entities, cores=None, grain=16, **kw):
    """Solve entities [0, n_entities) of `batch` on `cores` processes (default: all).  Workers inherit the batch by
    fork, so no per-job pickling is charged."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    jobs = [(e, min(n_entities, e + grain), kw) for e in range(0, n_entities, grain)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_init, initargs=(batch,)) as pool:
        pool.map(_job, [(0, 1, kw)] * cores)  # spin the workers up outside the timed region
        t0 = time.perf_counter()
        res = pool.map(_job, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    done = sum(r[0] for r in res)
    return {"entities": done, "seconds": dt, "entities_per_sec": done / dt, "cores": cores, "kind": kind(),
            "mean_nit": sum(r[1] for r in res) / max(done, 1)}
